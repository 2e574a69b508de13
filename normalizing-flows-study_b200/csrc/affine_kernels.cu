// affine_kernels.cu -- HBM-bound affine transforms (conditioner outputs already in HBM):
//   nf_affine_coupling_*     a1/a2 CouplingLayer transform          (coupling_layer.py:47-66,76-94)
//   nf_affine_ar_*           a11/a13 MAF.inverse / IAF.forward      (masked_autoregressive_flow.py:24-42,
//                                                                    inverse_autoregressive_flow.py:36-61)
#include "nf_common.cuh"

namespace nf {

// ------------------------------------------------------------------------------------------------
// a1/a2 affine coupling and a11/a13 affine autoregressive transforms: G lanes per row.
// ------------------------------------------------------------------------------------------------
template <typename T, int G>
__global__ void __launch_bounds__(256)
affine_coupling_fwd_kernel(const T* __restrict__ x, const T* __restrict__ s, const T* __restrict__ b,
                           const T* __restrict__ mask, T* __restrict__ y, T* __restrict__ ld, int64_t B, int D,
                           int inverse) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, g = lane % G;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nblk = (B + RPW - 1) / RPW;
    for (int64_t blk = warp; blk < nblk; blk += nwarps) {
        const int64_t row = blk * RPW + lane / G;
        const bool valid = row < B;
        T acc = T(0);
        if (valid) {
            for (int dd = g; dd < D; dd += G) {
                const int64_t o = row * D + dd;
                T out, t;
                affine_coupling_elem<T>(ld_stream(x + o), __ldg(mask + dd), ld_stream(s + o), ld_stream(b + o),
                                        inverse != 0, out, t);
                st_stream(y + o, scrub0(out));
                acc += t;
            }
        }
        acc = group_sum<T, G>(acc);
        if (valid && g == 0) ld[row] = scrub0(acc);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
affine_coupling_bwd_kernel(const T* __restrict__ x, const T* __restrict__ s, const T* __restrict__ b,
                           const T* __restrict__ mask, const T* __restrict__ gy, const T* __restrict__ gld,
                           T* __restrict__ gx, T* __restrict__ gs, T* __restrict__ gb, int64_t B, int D, int inverse) {
    const int64_t n = B * D;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += stride) {
        const int64_t row = o / D;
        const int dd = (int)(o - row * D);
        const T xv = x[o], sv = s[o], bv = b[o], m = __ldg(mask + dd);
        T out, t;
        affine_coupling_elem<T>(xv, m, sv, bv, inverse != 0, out, t);
        const T go = is_finite(out) ? gy[o] : T(0);
        T a, c, e;
        affine_coupling_elem_bwd<T>(xv, m, sv, bv, inverse != 0, go, gld[row], a, c, e);
        gx[o] = a; gs[o] = c; gb[o] = e;
    }
}

template <typename T, int G>
__global__ void __launch_bounds__(256)
affine_ar_fwd_kernel(const T* __restrict__ v, const T* __restrict__ params, T* __restrict__ out, T* __restrict__ ld,
                     int64_t B, int D, int mode) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, g = lane % G;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nblk = (B + RPW - 1) / RPW;
    const T lim = (mode == AR_IAF_FWD) ? T(50) : T(100);
    for (int64_t blk = warp; blk < nblk; blk += nwarps) {
        const int64_t row = blk * RPW + lane / G;
        const bool valid = row < B;
        T acc = T(0);
        if (valid) {
            for (int dd = g; dd < D; dd += G) {
                const T vv = ld_stream(v + row * D + dd);
                T o, t;
                affine_ar_elem<T>(mode, vv, ld_stream(params + row * 2 * D + dd),
                                  ld_stream(params + row * 2 * D + D + dd), o, t);
                if (!is_finite(o)) o = (mode == AR_IAF_FWD) ? vv : T(0);   // IAF scrubs to the input (:53)
                st_stream(out + row * D + dd, o);
                acc += t;
            }
        }
        acc = group_sum<T, G>(acc);
        if (valid && g == 0) ld[row] = clamp_mm(scrub0(acc), -lim, lim);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
affine_ar_bwd_kernel(const T* __restrict__ v, const T* __restrict__ params, const T* __restrict__ ld_saved,
                     const T* __restrict__ gout, const T* __restrict__ gld, T* __restrict__ gv,
                     T* __restrict__ gparams, int64_t B, int D, int mode) {
    const int64_t n = B * D;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const T lim = (mode == AR_IAF_FWD) ? T(50) : T(100);
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += stride) {
        const int64_t row = o / D;
        const int dd = (int)(o - row * D);
        const T vv = v[o], mu = params[row * 2 * D + dd], al = params[row * 2 * D + D + dd];
        T out, t;
        affine_ar_elem<T>(mode, vv, mu, al, out, t);
        const bool fin = is_finite(out);
        const T go = fin ? gout[o] : T(0);
        const T lds = ld_saved[row];
        const T gl = (lds > -lim && lds < lim) ? gld[row] : T(0);   // clamp(+-lim) passes strictly inside
        T a, c, e;
        affine_ar_elem_bwd<T>(mode, vv, mu, al, go, gl, a, c, e);
        if (!fin && mode == AR_IAF_FWD) a += gout[o];                // scrubbed x takes z's place
        gv[o] = a;
        gparams[row * 2 * D + dd] = c;
        gparams[row * 2 * D + D + dd] = e;
    }
}

static inline int grid_for(int64_t work_items, int per_block, int blocks_per_sm) {
    int64_t need = cdiv(work_items, per_block);
    int64_t cap = (int64_t)kNumSMs * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}
static inline int pick_group(int n) { int g = 1; while (g < n && g < 32) g <<= 1; return g; }

template <typename T>
static int affine_coupling_fwd_launch(const void* x, const void* s, const void* b, const void* mask, void* y, void* ld,
                                      int64_t B, int D, int inverse, cudaStream_t st) {
    const int G = pick_group(D);
    const int grid = grid_for(cdiv(B, 32 / G), 8, 32);
#define NF_AC(GG) affine_coupling_fwd_kernel<T, GG><<<grid, 256, 0, st>>>((const T*)x, (const T*)s, (const T*)b, \
                                                                          (const T*)mask, (T*)y, (T*)ld, B, D, inverse)
    switch (G) { case 1: NF_AC(1); break; case 2: NF_AC(2); break; case 4: NF_AC(4); break; case 8: NF_AC(8); break;
                 case 16: NF_AC(16); break; default: NF_AC(32); break; }
#undef NF_AC
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

template <typename T>
static int affine_ar_fwd_launch(const void* v, const void* params, void* out, void* ld, int64_t B, int D, int mode,
                                cudaStream_t st) {
    const int G = pick_group(D);
    const int grid = grid_for(cdiv(B, 32 / G), 8, 32);
#define NF_AR(GG) affine_ar_fwd_kernel<T, GG><<<grid, 256, 0, st>>>((const T*)v, (const T*)params, (T*)out, (T*)ld, B, D, mode)
    switch (G) { case 1: NF_AR(1); break; case 2: NF_AR(2); break; case 4: NF_AR(4); break; case 8: NF_AR(8); break;
                 case 16: NF_AR(16); break; default: NF_AR(32); break; }
#undef NF_AR
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

}  // namespace nf

using namespace nf;
#define NF_REQ(p) do { if ((p) == nullptr) return NF_ERR_NULL; } while (0)

extern "C" int nf_affine_coupling_forward(const void* x, const void* s_raw, const void* b_raw, const void* mask,
                                          void* y, void* ld, int64_t B, int D, int inverse, int dtype,
                                          nf_stream_t stream) {
    if (B < 0 || D < 1) return NF_ERR_BAD_SHAPE;
    if (B == 0) return NF_OK;
    NF_REQ(x); NF_REQ(s_raw); NF_REQ(b_raw); NF_REQ(mask); NF_REQ(y); NF_REQ(ld);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NF_F32) return affine_coupling_fwd_launch<float>(x, s_raw, b_raw, mask, y, ld, B, D, inverse, st);
    if (dtype == NF_F64) return affine_coupling_fwd_launch<double>(x, s_raw, b_raw, mask, y, ld, B, D, inverse, st);
    return NF_ERR_UNSUPPORTED;
}

extern "C" int nf_affine_coupling_backward(const void* x, const void* s_raw, const void* b_raw, const void* mask,
                                           const void* gy, const void* gld, void* gx, void* gs, void* gb, int64_t B,
                                           int D, int inverse, int dtype, nf_stream_t stream) {
    if (B < 0 || D < 1) return NF_ERR_BAD_SHAPE;
    if (B == 0) return NF_OK;
    NF_REQ(x); NF_REQ(s_raw); NF_REQ(b_raw); NF_REQ(mask); NF_REQ(gy); NF_REQ(gld); NF_REQ(gx); NF_REQ(gs); NF_REQ(gb);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(B * D, 256, 32);
    if (dtype == NF_F32)
        affine_coupling_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (const float*)s_raw,
            (const float*)b_raw, (const float*)mask, (const float*)gy, (const float*)gld, (float*)gx, (float*)gs,
            (float*)gb, B, D, inverse);
    else if (dtype == NF_F64)
        affine_coupling_bwd_kernel<double><<<grid, 256, 0, st>>>((const double*)x, (const double*)s_raw,
            (const double*)b_raw, (const double*)mask, (const double*)gy, (const double*)gld, (double*)gx,
            (double*)gs, (double*)gb, B, D, inverse);
    else return NF_ERR_UNSUPPORTED;
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

extern "C" int nf_affine_ar_forward(const void* v, const void* params, void* out, void* ld, int64_t B, int D,
                                    int mode, int dtype, nf_stream_t stream) {
    if (B < 0 || D < 1) return NF_ERR_BAD_SHAPE;
    if (mode != NF_AR_MAF_INVERSE && mode != NF_AR_IAF_FORWARD) return NF_ERR_UNSUPPORTED;
    if (B == 0) return NF_OK;
    NF_REQ(v); NF_REQ(params); NF_REQ(out); NF_REQ(ld);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NF_F32) return affine_ar_fwd_launch<float>(v, params, out, ld, B, D, mode, st);
    if (dtype == NF_F64) return affine_ar_fwd_launch<double>(v, params, out, ld, B, D, mode, st);
    return NF_ERR_UNSUPPORTED;
}

extern "C" int nf_affine_ar_backward(const void* v, const void* params, const void* ld_saved, const void* gout,
                                     const void* gld, void* gv, void* gparams, int64_t B, int D, int mode, int dtype,
                                     nf_stream_t stream) {
    if (B < 0 || D < 1) return NF_ERR_BAD_SHAPE;
    if (mode != NF_AR_MAF_INVERSE && mode != NF_AR_IAF_FORWARD) return NF_ERR_UNSUPPORTED;
    if (B == 0) return NF_OK;
    NF_REQ(v); NF_REQ(params); NF_REQ(ld_saved); NF_REQ(gout); NF_REQ(gld); NF_REQ(gv); NF_REQ(gparams);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(B * D, 256, 32);
    if (dtype == NF_F32)
        affine_ar_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)v, (const float*)params,
            (const float*)ld_saved, (const float*)gout, (const float*)gld, (float*)gv, (float*)gparams, B, D, mode);
    else if (dtype == NF_F64)
        affine_ar_bwd_kernel<double><<<grid, 256, 0, st>>>((const double*)v, (const double*)params,
            (const double*)ld_saved, (const double*)gout, (const double*)gld, (double*)gv, (double*)gparams, B, D, mode);
    else return NF_ERR_UNSUPPORTED;
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}
