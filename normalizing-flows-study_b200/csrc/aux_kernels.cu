// aux_kernels.cu -- the remaining [B,D]-sized pieces of the layered (training / float64) route:
//   nf_feature_affine_*   per-feature affine  y = (x - sub[d]) / div[d] * mul[d] + add[d]
//                         = between-layer BatchNorm on running statistics and its inverse
//                           (normalizing_flow_model.py:67-85, :110-128) and the spline layer's
//                           data_min/data_max rescale of the conditioner input (spline_coupling_layer.py:78-94)
//   nf_col_stats          per-feature batch mean / biased variance (running-stat update, normalizing_flow_model.py:74-79)
//   nf_ar_step_*          one step of the sequential autoregressive loop
//                         (masked_autoregressive_flow.py:55-67, inverse_autoregressive_flow.py:76-91)
//   nf_ar_finish_*        the NaN/Inf scrubs and log-det clamp that close that loop (:69-76 / :93-101)
// All are streaming, HBM-bound kernels; reductions over rows accumulate in double.
#include "nf_common.cuh"

namespace nf {

static inline int ew_grid(int64_t n) {
    int64_t need = cdiv(n, 256);
    int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

template <typename T>
__global__ void __launch_bounds__(256)
feature_affine_fwd_kernel(const T* __restrict__ x, const T* __restrict__ sub, const T* __restrict__ div,
                          const T* __restrict__ mul, const T* __restrict__ add, T add_scalar, T* __restrict__ y,
                          int64_t B, int D) {
    const int64_t n = B * D, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int d = (int)(i % D);
        T v = x[i];
        if (sub) v = v - sub[d];
        if (div) v = v / div[d];
        if (mul) v = v * mul[d];
        v = v + (add ? add[d] : add_scalar);
        y[i] = v;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
feature_affine_bwd_x_kernel(const T* __restrict__ div, const T* __restrict__ mul, const T* __restrict__ gy,
                            T* __restrict__ gx, int64_t B, int D) {
    const int64_t n = B * D, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int d = (int)(i % D);
        T g = gy[i];
        if (mul) g = g * mul[d];
        if (div) g = g / div[d];
        gx[i] = g;
    }
}

// partial column sums S1[d] = sum_b w[b,d], S2[d] = sum_b w[b,d]*(x[b,d]-sub[d]); w = gy (or 1 for statistics).
// block = 32 columns x 8 row lanes; grid.y row chunks combine with double atomics into acc[2*D].
template <typename T>
__global__ void __launch_bounds__(256)
col_moments_kernel(const T* __restrict__ x, const T* __restrict__ sub, const T* __restrict__ w, double* __restrict__ acc,
                   int64_t B, int D, int64_t rows_per_chunk, int square) {
    __shared__ double s1[8][33], s2[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cx;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t r1 = (B < r0 + rows_per_chunk) ? B : r0 + rows_per_chunk;
    double a = 0.0, b = 0.0;
    if (col < D) {
        const double sb = sub ? (double)sub[col] : 0.0;
        for (int64_t r = r0 + ry; r < r1; r += 8) {
            const double xv = (double)x[r * D + col] - sb;
            const double wv = w ? (double)w[r * D + col] : 1.0;
            if (square) { a += xv; b += xv * xv; }      // statistics mode: sum x, sum x^2
            else        { a += wv; b += wv * xv; }
        }
    }
    s1[ry][cx] = a; s2[ry][cx] = b;
    __syncthreads();
    if (ry == 0 && col < D) {
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { ta += s1[i][cx]; tb += s2[i][cx]; }
        atomicAdd(acc + col, ta);
        atomicAdd(acc + D + col, tb);
    }
}

template <typename T>
__global__ void feature_affine_bwd_finish_kernel(const double* __restrict__ acc, const T* __restrict__ div,
                                                 const T* __restrict__ mul, T* __restrict__ gsub, T* __restrict__ gdiv,
                                                 T* __restrict__ gmul, T* __restrict__ gadd, int D) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const double S1 = acc[d], S2 = acc[D + d];
    const double dv = div ? (double)div[d] : 1.0, ml = mul ? (double)mul[d] : 1.0;
    if (gadd) gadd[d] = (T)S1;
    if (gmul) gmul[d] = (T)(S2 / dv);
    if (gsub) gsub[d] = (T)(-S1 * ml / dv);
    if (gdiv) gdiv[d] = (T)(-S2 * ml / (dv * dv));
}

template <typename T>
__global__ void col_stats_finish_kernel(const double* __restrict__ acc, T* __restrict__ mean, T* __restrict__ var,
                                        int64_t B, int D) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const double m = acc[d] / (double)B;
    double v = acc[D + d] / (double)B - m * m;
    if (v < 0.0) v = 0.0;
    mean[d] = (T)m;
    var[d] = (T)v;
}

static inline void chunking(int64_t B, int& chunks, int64_t& rpc) {
    int64_t c = B / 2048;
    if (c < 1) c = 1;
    if (c > 4 * kNumSMs) c = 4 * kNumSMs;
    rpc = cdiv(B, c);
    chunks = (int)cdiv(B, rpc);
}

// ------------------------------------------------------------------------------------------------
// sequential autoregressive loop body and epilogue
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
ar_step_fwd_kernel(const T* __restrict__ cur, const T* __restrict__ v, const T* __restrict__ params,
                   const T* __restrict__ ld_in, T* __restrict__ out, T* __restrict__ ld_out, int64_t B, int D, int col,
                   int mode) {
    const int64_t n = B * D, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t row = i / D;
        const int d = (int)(i - row * D);
        if (d != col) { out[i] = cur[i]; continue; }
        T o, t;
        affine_ar_elem<T>(mode, v[i], params[row * 2 * D + d], params[row * 2 * D + D + d], o, t);
        out[i] = o;
        ld_out[row] = (ld_in ? ld_in[row] : T(0)) + t;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
ar_step_bwd_kernel(const T* __restrict__ v, const T* __restrict__ params, const T* __restrict__ gout,
                   const T* __restrict__ gld, T* __restrict__ gcur, T* __restrict__ gv, T* __restrict__ gparams,
                   int64_t B, int D, int col, int mode) {
    const int64_t n = B * D, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t row = i / D;
        const int d = (int)(i - row * D);
        T a = T(0), c = T(0), e = T(0);
        if (d == col) {
            affine_ar_elem_bwd<T>(mode, v[i], params[row * 2 * D + d], params[row * 2 * D + D + d], gout[i], gld[row],
                                  a, c, e);
            gcur[i] = T(0);
        } else {
            gcur[i] = gout[i];
        }
        gv[i] = a;
        gparams[row * 2 * D + d] = c;
        gparams[row * 2 * D + D + d] = e;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
ar_finish_fwd_kernel(const T* __restrict__ cur, const T* __restrict__ v, const T* __restrict__ ld_sum,
                     T* __restrict__ out, T* __restrict__ ld, int64_t B, int D, int mode) {
    const int64_t n = B * D, stride = (int64_t)gridDim.x * blockDim.x;
    const T lim = (mode == AR_IAF_INV) ? T(50) : T(100);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T c = cur[i];
        out[i] = is_finite(c) ? c : ((mode == AR_IAF_INV) ? v[i] : T(0));
        if (i < B) ld[i] = clamp_mm(scrub0(ld_sum[i]), -lim, lim);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
ar_finish_bwd_kernel(const T* __restrict__ cur, const T* __restrict__ ld_sum, const T* __restrict__ gout,
                     const T* __restrict__ gld, T* __restrict__ gcur, T* __restrict__ gv, T* __restrict__ gld_sum,
                     int64_t B, int D, int mode) {
    const int64_t n = B * D, stride = (int64_t)gridDim.x * blockDim.x;
    const T lim = (mode == AR_IAF_INV) ? T(50) : T(100);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool fin = is_finite(cur[i]);
        gcur[i] = fin ? gout[i] : T(0);
        gv[i] = (!fin && mode == AR_IAF_INV) ? gout[i] : T(0);
        if (i < B) {
            const T s = ld_sum[i];
            gld_sum[i] = (is_finite(s) && pass_mm(s, -lim, lim)) ? gld[i] : T(0);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Flow.log_prob head for a standard-normal base (flow.py:56-73): lp = sum_d(-z^2/2) - D/2 log(2 pi) + log_det
// ------------------------------------------------------------------------------------------------
template <typename T, int G>
__global__ void __launch_bounds__(256)
std_normal_log_prob_fwd_kernel(const T* __restrict__ z, const T* __restrict__ ld, T* __restrict__ lp, int64_t B, int D,
                               T norm_const) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, g = lane % G;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nblk = (B + RPW - 1) / RPW;
    for (int64_t blk = warp; blk < nblk; blk += nwarps) {
        const int64_t row = blk * RPW + lane / G;
        const bool valid = row < B;
        T acc = T(0);
        if (valid)
            for (int d = g; d < D; d += G) { const T v = ld_stream(z + row * D + d); acc += T(-0.5) * v * v; }
        acc = group_sum<T, G>(acc);
        if (valid && g == 0) lp[row] = acc - norm_const + (ld ? ld[row] : T(0));
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
std_normal_log_prob_bwd_kernel(const T* __restrict__ z, const T* __restrict__ glp, T* __restrict__ gz, int64_t B, int D) {
    const int64_t n = B * D, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) gz[i] = -z[i] * glp[i / D];
}

template <typename T>
static int feature_affine_bwd(const void* x, const void* sub, const void* div, const void* mul, const void* gy, void* gx,
                              void* gsub, void* gdiv, void* gmul, void* gadd, void* ws, int64_t B, int D,
                              cudaStream_t st) {
    feature_affine_bwd_x_kernel<T><<<ew_grid(B * D), 256, 0, st>>>((const T*)div, (const T*)mul, (const T*)gy, (T*)gx, B, D);
    count_launch();
    NF_LAUNCH_CHECK();
    if (!gsub && !gdiv && !gmul && !gadd) return NF_OK;
    NF_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * D, st));
    int chunks; int64_t rpc;
    chunking(B, chunks, rpc);
    dim3 grid((unsigned)cdiv(D, 32), (unsigned)chunks);
    col_moments_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (const T*)sub, (const T*)gy, (double*)ws, B, D, rpc, 0);
    count_launch();
    NF_LAUNCH_CHECK();
    feature_affine_bwd_finish_kernel<T><<<(D + 127) / 128, 128, 0, st>>>((const double*)ws, (const T*)div, (const T*)mul,
                                                                        (T*)gsub, (T*)gdiv, (T*)gmul, (T*)gadd, D);
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

template <typename T>
static int col_stats(const void* x, void* mean, void* var, void* ws, int64_t B, int D, cudaStream_t st) {
    NF_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * D, st));
    int chunks; int64_t rpc;
    chunking(B, chunks, rpc);
    dim3 grid((unsigned)cdiv(D, 32), (unsigned)chunks);
    col_moments_kernel<T><<<grid, 256, 0, st>>>((const T*)x, nullptr, nullptr, (double*)ws, B, D, rpc, 1);
    count_launch();
    NF_LAUNCH_CHECK();
    col_stats_finish_kernel<T><<<(D + 127) / 128, 128, 0, st>>>((const double*)ws, (T*)mean, (T*)var, B, D);
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

}  // namespace nf

using namespace nf;
#define NF_REQ(p) do { if ((p) == nullptr) return NF_ERR_NULL; } while (0)
#define NF_DISPATCH(expr_f, expr_d)                         \
    do {                                                    \
        if (dtype == NF_F32) { expr_f; }                    \
        else if (dtype == NF_F64) { expr_d; }               \
        else return NF_ERR_UNSUPPORTED;                     \
    } while (0)

extern "C" int nf_feature_affine_forward(const void* x, const void* sub, const void* div, const void* mul,
                                         const void* add, double add_scalar, void* y, int64_t B, int D, int dtype,
                                         nf_stream_t stream) {
    if (B < 0 || D < 1) return NF_ERR_BAD_SHAPE;
    if (B == 0) return NF_OK;
    NF_REQ(x); NF_REQ(y);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ew_grid(B * D);
    NF_DISPATCH(
        (feature_affine_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (const float*)sub, (const float*)div,
            (const float*)mul, (const float*)add, (float)add_scalar, (float*)y, B, D)),
        (feature_affine_fwd_kernel<double><<<grid, 256, 0, st>>>((const double*)x, (const double*)sub, (const double*)div,
            (const double*)mul, (const double*)add, add_scalar, (double*)y, B, D)));
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

extern "C" int nf_feature_affine_backward(const void* x, const void* sub, const void* div, const void* mul,
                                          const void* gy, void* gx, void* gsub, void* gdiv, void* gmul, void* gadd,
                                          void* workspace, int64_t B, int D, int dtype, nf_stream_t stream) {
    if (B < 0 || D < 1) return NF_ERR_BAD_SHAPE;
    const bool need_red = gsub || gdiv || gmul || gadd;
    if (need_red) NF_REQ(workspace);
    if (B == 0) {
        if (need_red) {
            const size_t es = dtype == NF_F64 ? 8 : 4;
            cudaStream_t st0 = (cudaStream_t)stream;
            if (gsub) NF_CUDA(cudaMemsetAsync(gsub, 0, es * D, st0));
            if (gdiv) NF_CUDA(cudaMemsetAsync(gdiv, 0, es * D, st0));
            if (gmul) NF_CUDA(cudaMemsetAsync(gmul, 0, es * D, st0));
            if (gadd) NF_CUDA(cudaMemsetAsync(gadd, 0, es * D, st0));
        }
        return NF_OK;
    }
    NF_REQ(x); NF_REQ(gy); NF_REQ(gx);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NF_F32) return feature_affine_bwd<float>(x, sub, div, mul, gy, gx, gsub, gdiv, gmul, gadd, workspace, B, D, st);
    if (dtype == NF_F64) return feature_affine_bwd<double>(x, sub, div, mul, gy, gx, gsub, gdiv, gmul, gadd, workspace, B, D, st);
    return NF_ERR_UNSUPPORTED;
}

extern "C" int nf_col_stats(const void* x, void* mean, void* var, void* workspace, int64_t B, int D, int dtype,
                            nf_stream_t stream) {
    if (B < 1 || D < 1) return NF_ERR_BAD_SHAPE;
    NF_REQ(x); NF_REQ(mean); NF_REQ(var); NF_REQ(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NF_F32) return col_stats<float>(x, mean, var, workspace, B, D, st);
    if (dtype == NF_F64) return col_stats<double>(x, mean, var, workspace, B, D, st);
    return NF_ERR_UNSUPPORTED;
}

extern "C" int nf_ar_step_forward(const void* cur, const void* v, const void* params, const void* ld_in, void* out,
                                  void* ld_out, int64_t B, int D, int col, int mode, int dtype, nf_stream_t stream) {
    if (B < 0 || D < 1 || col < 0 || col >= D) return NF_ERR_BAD_SHAPE;
    if (mode != NF_AR_MAF_FORWARD && mode != NF_AR_IAF_INVERSE) return NF_ERR_UNSUPPORTED;
    if (B == 0) return NF_OK;
    NF_REQ(cur); NF_REQ(v); NF_REQ(params); NF_REQ(out); NF_REQ(ld_out);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ew_grid(B * D);
    NF_DISPATCH(
        (ar_step_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)cur, (const float*)v, (const float*)params,
            (const float*)ld_in, (float*)out, (float*)ld_out, B, D, col, mode)),
        (ar_step_fwd_kernel<double><<<grid, 256, 0, st>>>((const double*)cur, (const double*)v, (const double*)params,
            (const double*)ld_in, (double*)out, (double*)ld_out, B, D, col, mode)));
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

extern "C" int nf_ar_step_backward(const void* v, const void* params, const void* gout, const void* gld, void* gcur,
                                   void* gv, void* gparams, int64_t B, int D, int col, int mode, int dtype,
                                   nf_stream_t stream) {
    if (B < 0 || D < 1 || col < 0 || col >= D) return NF_ERR_BAD_SHAPE;
    if (mode != NF_AR_MAF_FORWARD && mode != NF_AR_IAF_INVERSE) return NF_ERR_UNSUPPORTED;
    if (B == 0) return NF_OK;
    NF_REQ(v); NF_REQ(params); NF_REQ(gout); NF_REQ(gld); NF_REQ(gcur); NF_REQ(gv); NF_REQ(gparams);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ew_grid(B * D);
    NF_DISPATCH(
        (ar_step_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)v, (const float*)params, (const float*)gout,
            (const float*)gld, (float*)gcur, (float*)gv, (float*)gparams, B, D, col, mode)),
        (ar_step_bwd_kernel<double><<<grid, 256, 0, st>>>((const double*)v, (const double*)params, (const double*)gout,
            (const double*)gld, (double*)gcur, (double*)gv, (double*)gparams, B, D, col, mode)));
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

extern "C" int nf_ar_finish_forward(const void* cur, const void* v, const void* ld_sum, void* out, void* ld, int64_t B,
                                    int D, int mode, int dtype, nf_stream_t stream) {
    if (B < 0 || D < 1) return NF_ERR_BAD_SHAPE;
    if (mode != NF_AR_MAF_FORWARD && mode != NF_AR_IAF_INVERSE) return NF_ERR_UNSUPPORTED;
    if (B == 0) return NF_OK;
    NF_REQ(cur); NF_REQ(v); NF_REQ(ld_sum); NF_REQ(out); NF_REQ(ld);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ew_grid(B * D);
    NF_DISPATCH(
        (ar_finish_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)cur, (const float*)v, (const float*)ld_sum,
            (float*)out, (float*)ld, B, D, mode)),
        (ar_finish_fwd_kernel<double><<<grid, 256, 0, st>>>((const double*)cur, (const double*)v, (const double*)ld_sum,
            (double*)out, (double*)ld, B, D, mode)));
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

extern "C" int nf_ar_finish_backward(const void* cur, const void* ld_sum, const void* gout, const void* gld, void* gcur,
                                     void* gv, void* gld_sum, int64_t B, int D, int mode, int dtype,
                                     nf_stream_t stream) {
    if (B < 0 || D < 1) return NF_ERR_BAD_SHAPE;
    if (mode != NF_AR_MAF_FORWARD && mode != NF_AR_IAF_INVERSE) return NF_ERR_UNSUPPORTED;
    if (B == 0) return NF_OK;
    NF_REQ(cur); NF_REQ(ld_sum); NF_REQ(gout); NF_REQ(gld); NF_REQ(gcur); NF_REQ(gv); NF_REQ(gld_sum);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ew_grid(B * D);
    NF_DISPATCH(
        (ar_finish_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)cur, (const float*)ld_sum, (const float*)gout,
            (const float*)gld, (float*)gcur, (float*)gv, (float*)gld_sum, B, D, mode)),
        (ar_finish_bwd_kernel<double><<<grid, 256, 0, st>>>((const double*)cur, (const double*)ld_sum, (const double*)gout,
            (const double*)gld, (double*)gcur, (double*)gv, (double*)gld_sum, B, D, mode)));
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

extern "C" int nf_std_normal_log_prob_forward(const void* z, const void* ld, void* lp, int64_t B, int D, int dtype,
                                              nf_stream_t stream) {
    if (B < 0 || D < 1) return NF_ERR_BAD_SHAPE;
    if (B == 0) return NF_OK;
    NF_REQ(z); NF_REQ(lp);
    cudaStream_t st = (cudaStream_t)stream;
    const double nc = 0.5 * (double)D * 1.8378770664093453;   // D/2 * log(2 pi)
    int G = 1; while (G < D && G < 32) G <<= 1;
    int64_t need = cdiv(cdiv(B, 32 / G), 8), cap = (int64_t)kNumSMs * 32;
    const int grid = (int)(need < 1 ? 1 : (need < cap ? need : cap));
#define NF_LP(T, GG) std_normal_log_prob_fwd_kernel<T, GG><<<grid, 256, 0, st>>>((const T*)z, (const T*)ld, (T*)lp, B, D, (T)nc)
#define NF_LP_G(T) switch (G) { case 1: NF_LP(T, 1); break; case 2: NF_LP(T, 2); break; case 4: NF_LP(T, 4); break; \
                                case 8: NF_LP(T, 8); break; case 16: NF_LP(T, 16); break; default: NF_LP(T, 32); break; }
    NF_DISPATCH(NF_LP_G(float), NF_LP_G(double));
#undef NF_LP_G
#undef NF_LP
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

extern "C" int nf_std_normal_log_prob_backward(const void* z, const void* glp, void* gz, int64_t B, int D, int dtype,
                                               nf_stream_t stream) {
    if (B < 0 || D < 1) return NF_ERR_BAD_SHAPE;
    if (B == 0) return NF_OK;
    NF_REQ(z); NF_REQ(glp); NF_REQ(gz);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ew_grid(B * D);
    NF_DISPATCH(
        (std_normal_log_prob_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)z, (const float*)glp, (float*)gz, B, D)),
        (std_normal_log_prob_bwd_kernel<double><<<grid, 256, 0, st>>>((const double*)z, (const double*)glp, (double*)gz, B, D)));
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}
