// arqs_kernels.cu -- one step of the sequential loops of ARQS (src/flows/spline/arqs.py:53-76 and :93-116):
//   out = cur with column `col` replaced by rational_quadratic_spline(v[:, col]; params_row) (the public [0,1]
//   spline, rational_quadratic_spline.py:4-104), ld_out = ld_in + log|dy/dx|.
// params is the [B, 3K-1] block of the conditioner output that belongs to dimension `col` (row pitch ldp): the host
// evaluates only those 3K-1 rows of MADE's last masked linear.  The log-det vector is float32 whatever the data
// dtype, as in the reference (`torch.zeros(B)` accumulated in place, arqs.py:52,75); a float64 run therefore adds in
// double and rounds to float32 at every step, like ATen's type-promoting in-place add.
// Streaming, HBM-bound: the cost of an ARQS pass is the D conditioner evaluations, not this kernel.
#include "nf_common.cuh"

namespace nf {

static inline int arqs_grid(int64_t n) {
    int64_t need = cdiv(n, 256);
    int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

template <typename T, int KMAX>
__global__ void __launch_bounds__(256)
arqs_step_fwd_kernel(const T* __restrict__ cur, const T* __restrict__ v, const T* __restrict__ params, int64_t ldp,
                     const float* __restrict__ ld_in, T* __restrict__ out, float* __restrict__ ld_out, int64_t B, int D,
                     int col, int K, int inverse, RqsCfg<T> c) {
    const int64_t n = B * D, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t row = i / D;
        const int d = (int)(i - row * D);
        if (d != col) { out[i] = cur[i]; continue; }
        const T* pp = params + row * ldp;
        T uw[KMAX], uh[KMAX], ud[KMAX];
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) {
            uw[j] = (j < K) ? pp[j] : T(0);
            uh[j] = (j < K) ? pp[K + j] : T(0);
            ud[j] = (j < K - 1) ? pp[2 * K + j] : T(0);
        }
        T o, lad;
        rqs_eval<T, KMAX, false>(v[i], uw, uh, ud, K, inverse != 0, c, o, lad);
        out[i] = o;
        ld_out[row] = (float)((T)(ld_in ? ld_in[row] : 0.f) + lad);
    }
}

// reverse mode of one step: gcur = gout with column `col` zeroed, gv = d/dv (zero outside `col`), gparams [B, 3K-1]
template <typename T, int KMAX>
__global__ void __launch_bounds__(128)
arqs_step_bwd_kernel(const T* __restrict__ v, const T* __restrict__ params, int64_t ldp, const T* __restrict__ gout,
                     const float* __restrict__ gld, T* __restrict__ gcur, T* __restrict__ gv, T* __restrict__ gparams,
                     int64_t B, int D, int col, int K, int inverse, RqsCfg<T> c) {
    const int64_t n = B * D, stride = (int64_t)gridDim.x * blockDim.x;
    const int P = 3 * K - 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t row = i / D;
        const int d = (int)(i - row * D);
        if (d != col) { gcur[i] = gout[i]; gv[i] = T(0); continue; }
        const T* pp = params + row * ldp;
        T uw[KMAX], uh[KMAX], ud[KMAX], guw[KMAX], guh[KMAX], gud[KMAX];
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) {
            uw[j] = (j < K) ? pp[j] : T(0);
            uh[j] = (j < K) ? pp[K + j] : T(0);
            ud[j] = (j < K - 1) ? pp[2 * K + j] : T(0);
            guw[j] = T(0); guh[j] = T(0); gud[j] = T(0);
        }
        T g = T(0);
        rqs_eval_bwd<T, KMAX, false>(v[i], uw, uh, ud, K, inverse != 0, c, gout[i], (T)(gld ? gld[row] : 0.f), g, guw, guh, gud);
        gcur[i] = T(0);
        gv[i] = g;
        T* gp = gparams + row * P;
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) if (j < K) { gp[j] = guw[j]; gp[K + j] = guh[j]; }
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) if (j < K - 1) gp[2 * K + j] = gud[j];
    }
}

template <typename T>
static int arqs_fwd_launch(const void* cur, const void* v, const void* params, int64_t ldp, const void* ld_in, void* out,
                           void* ld_out, int64_t B, int D, int col, int K, int inverse, double mw, double mh, double md,
                           cudaStream_t st) {
    const RqsCfg<T> c = make_rqs_cfg<T>(false, K, 0.0, mw, mh, md);
    const int grid = arqs_grid(B * D);
#define NF_AF(KM) arqs_step_fwd_kernel<T, KM><<<grid, 256, 0, st>>>((const T*)cur, (const T*)v, (const T*)params, ldp, \
        (const float*)ld_in, (T*)out, (float*)ld_out, B, D, col, K, inverse, c)
    if (K <= 8) NF_AF(8); else NF_AF(16);
#undef NF_AF
    return NF_OK;
}

template <typename T>
static int arqs_bwd_launch(const void* v, const void* params, int64_t ldp, const void* gout, const void* gld, void* gcur,
                           void* gv, void* gparams, int64_t B, int D, int col, int K, int inverse, double mw, double mh,
                           double md, cudaStream_t st) {
    const RqsCfg<T> c = make_rqs_cfg<T>(false, K, 0.0, mw, mh, md);
    const int grid = arqs_grid(B * D);
#define NF_AB(KM) arqs_step_bwd_kernel<T, KM><<<grid, 128, 0, st>>>((const T*)v, (const T*)params, ldp, (const T*)gout, \
        (const float*)gld, (T*)gcur, (T*)gv, (T*)gparams, B, D, col, K, inverse, c)
    if (K <= 8) NF_AB(8); else NF_AB(16);
#undef NF_AB
    return NF_OK;
}

}  // namespace nf

using namespace nf;
#define NF_REQ(p) do { if ((p) == nullptr) return NF_ERR_NULL; } while (0)

extern "C" int nf_arqs_step_forward(const void* cur, const void* v, const void* params, int64_t ldp, const void* ld_in,
                                    void* out, void* ld_out, int64_t B, int D, int col, int num_bins, int inverse,
                                    double min_bin_width, double min_bin_height, double min_derivative, int dtype,
                                    nf_stream_t stream) {
    if (B < 0 || D < 1 || col < 0 || col >= D || num_bins < 2 || ldp < 3 * (int64_t)num_bins - 1) return NF_ERR_BAD_SHAPE;
    if (num_bins > 16) return NF_ERR_UNSUPPORTED;
    if (B == 0) return NF_OK;
    NF_REQ(cur); NF_REQ(v); NF_REQ(params); NF_REQ(out); NF_REQ(ld_out);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NF_F32) arqs_fwd_launch<float>(cur, v, params, ldp, ld_in, out, ld_out, B, D, col, num_bins, inverse, min_bin_width, min_bin_height, min_derivative, st);
    else if (dtype == NF_F64) arqs_fwd_launch<double>(cur, v, params, ldp, ld_in, out, ld_out, B, D, col, num_bins, inverse, min_bin_width, min_bin_height, min_derivative, st);
    else return NF_ERR_UNSUPPORTED;
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

extern "C" int nf_arqs_step_backward(const void* v, const void* params, int64_t ldp, const void* gout, const void* gld,
                                     void* gcur, void* gv, void* gparams, int64_t B, int D, int col, int num_bins,
                                     int inverse, double min_bin_width, double min_bin_height, double min_derivative,
                                     int dtype, nf_stream_t stream) {
    if (B < 0 || D < 1 || col < 0 || col >= D || num_bins < 2 || ldp < 3 * (int64_t)num_bins - 1) return NF_ERR_BAD_SHAPE;
    if (num_bins > 16) return NF_ERR_UNSUPPORTED;
    if (B == 0) return NF_OK;
    NF_REQ(v); NF_REQ(params); NF_REQ(gout); NF_REQ(gcur); NF_REQ(gv); NF_REQ(gparams);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NF_F32) arqs_bwd_launch<float>(v, params, ldp, gout, gld, gcur, gv, gparams, B, D, col, num_bins, inverse, min_bin_width, min_bin_height, min_derivative, st);
    else if (dtype == NF_F64) arqs_bwd_launch<double>(v, params, ldp, gout, gld, gcur, gv, gparams, B, D, col, num_bins, inverse, min_bin_width, min_bin_height, min_derivative, st);
    else return NF_ERR_UNSUPPORTED;
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}
