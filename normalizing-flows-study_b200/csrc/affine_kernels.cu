// affine_kernels.cu -- HBM-bound affine transforms (conditioner outputs already in HBM):
//   nf_affine_coupling_*     a1/a2 CouplingLayer transform          (coupling_layer.py:47-66,76-94)
//   nf_affine_ar_*           a11/a13 MAF.inverse / IAF.forward      (masked_autoregressive_flow.py:24-42,
//                                                                    inverse_autoregressive_flow.py:36-61)
#include <initializer_list>
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

// ------------------------------------------------------------------------------------------------
// 128-bit paths.  Vec<T> is the 16-byte vector of T (float4 / double2); rows whose length is a multiple of the vector
// width are walked by G lanes in 16-byte steps, so every LDG/STG is a full 128-bit access and a warp covers whole
// 512-byte segments.  OP 0: affine coupling (s, b: two [B,D] arrays + column mask), OP 1: affine autoregressive
// (params [B,2D] = [mu | alpha]).
// ------------------------------------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float>  { using type = float4;  static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };
template <typename T> __device__ __forceinline__ void vload(const T* p, T (&o)[Vec<T>::N]) {
    const typename Vec<T>::type v = __ldcs(reinterpret_cast<const typename Vec<T>::type*>(p));
    if constexpr (Vec<T>::N == 4) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; } else { o[0] = v.x; o[1] = v.y; }
}
template <typename T> __device__ __forceinline__ void vstore(T* p, const T (&o)[Vec<T>::N]) {
    typename Vec<T>::type v;
    if constexpr (Vec<T>::N == 4) { v.x = o[0]; v.y = o[1]; v.z = o[2]; v.w = o[3]; } else { v.x = o[0]; v.y = o[1]; }
    __stcs(reinterpret_cast<typename Vec<T>::type*>(p), v);
}

template <typename T, int G, int OP>
__global__ void __launch_bounds__(256)
affine_rows_vec_fwd_kernel(const T* __restrict__ x, const T* __restrict__ p0, const T* __restrict__ p1,
                           const T* __restrict__ mask, T* __restrict__ y, T* __restrict__ ld, int64_t B, int D, int mode) {
    constexpr int V = Vec<T>::N;
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
            for (int c = g * V; c < D; c += G * V) {
                T xv[V], a[V], b[V], o[V];
                vload<T>(x + row * D + c, xv);
                if constexpr (OP == 0) { vload<T>(p0 + row * D + c, a); vload<T>(p1 + row * D + c, b); }
                else                   { vload<T>(p0 + row * 2 * D + c, a); vload<T>(p0 + row * 2 * D + D + c, b); }
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    T t;
                    if constexpr (OP == 0) {
                        affine_coupling_elem<T>(xv[e], __ldg(mask + c + e), a[e], b[e], mode != 0, o[e], t);
                        o[e] = scrub0(o[e]);
                    } else {
                        affine_ar_elem<T>(mode, xv[e], a[e], b[e], o[e], t);
                        if (!is_finite(o[e])) o[e] = (mode == AR_IAF_FWD) ? xv[e] : T(0);
                    }
                    acc += t;
                }
                vstore<T>(y + row * D + c, o);
            }
        }
        acc = group_sum<T, G>(acc);
        if (valid && g == 0) ld[row] = (OP == 0) ? scrub0(acc) : clamp_mm(scrub0(acc), -lim, lim);
    }
}

// float, D == 2: one thread owns two consecutive rows (one float4 of x / s / b, one float2 of log-dets)
template <int OP>
__global__ void __launch_bounds__(256)
affine_rows_d2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ p0, const float* __restrict__ p1,
                          const float* __restrict__ mask, float* __restrict__ y, float* __restrict__ ld, int64_t B, int mode) {
    const int64_t npair = B >> 1, stride = (int64_t)gridDim.x * blockDim.x;
    const float m0 = (OP == 0) ? __ldg(mask) : 0.f, m1 = (OP == 0) ? __ldg(mask + 1) : 0.f;
    const float lim = (mode == AR_IAF_FWD) ? 50.f : 100.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += stride) {
        float xv[4], a[4], b[4], o[4], t[4];
        vload<float>(x + 4 * i, xv);
        if constexpr (OP == 0) { vload<float>(p0 + 4 * i, a); vload<float>(p1 + 4 * i, b); }
        else {   // params rows: [mu0 mu1 al0 al1]
            float r0[4], r1[4];
            vload<float>(p0 + 8 * i, r0); vload<float>(p0 + 8 * i + 4, r1);
            a[0] = r0[0]; a[1] = r0[1]; b[0] = r0[2]; b[1] = r0[3];
            a[2] = r1[0]; a[3] = r1[1]; b[2] = r1[2]; b[3] = r1[3];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if constexpr (OP == 0) {
                affine_coupling_elem<float>(xv[e], (e & 1) ? m1 : m0, a[e], b[e], mode != 0, o[e], t[e]);
                o[e] = scrub0(o[e]);
            } else {
                affine_ar_elem<float>(mode, xv[e], a[e], b[e], o[e], t[e]);
                if (!is_finite(o[e])) o[e] = (mode == AR_IAF_FWD) ? xv[e] : 0.f;
            }
        }
        vstore<float>(y + 4 * i, o);
        float2 l;
        l.x = (OP == 0) ? scrub0(t[0] + t[1]) : clamp_mm(scrub0(t[0] + t[1]), -lim, lim);
        l.y = (OP == 0) ? scrub0(t[2] + t[3]) : clamp_mm(scrub0(t[2] + t[3]), -lim, lim);
        __stcs(reinterpret_cast<float2*>(ld + 2 * i), l);
    }
}

// reverse mode of the affine coupling in 16-byte steps over the flattened [B*D] arrays (D % V == 0 or D == 2)
template <typename T>
__global__ void __launch_bounds__(256)
affine_coupling_vec_bwd_kernel(const T* __restrict__ x, const T* __restrict__ s, const T* __restrict__ b,
                               const T* __restrict__ mask, const T* __restrict__ gy, const T* __restrict__ gld,
                               T* __restrict__ gx, T* __restrict__ gs, T* __restrict__ gb, int64_t B, int D, int inverse) {
    constexpr int V = Vec<T>::N;
    const int64_t nvec = B * D / V, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const int64_t o = i * V;
        T xv[V], sv[V], bv[V], g[V], a[V], c[V], e[V];
        vload<T>(x + o, xv); vload<T>(s + o, sv); vload<T>(b + o, bv); vload<T>(gy + o, g);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int64_t row = (o + k) / D;
            const int dd = (int)((o + k) - row * D);
            const T m = __ldg(mask + dd);
            T out, t;
            affine_coupling_elem<T>(xv[k], m, sv[k], bv[k], inverse != 0, out, t);
            const T go = is_finite(out) ? g[k] : T(0);
            affine_coupling_elem_bwd<T>(xv[k], m, sv[k], bv[k], inverse != 0, go, __ldg(gld + row), a[k], c[k], e[k]);
        }
        vstore<T>(gx + o, a); vstore<T>(gs + o, c); vstore<T>(gb + o, e);
    }
}

static inline int grid_for(int64_t work_items, int per_block, int blocks_per_sm) {
    int64_t need = cdiv(work_items, per_block);
    int64_t cap = (int64_t)kNumSMs * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}
static inline int pick_group(int n) { int g = 1; while (g < n && g < 32) g <<= 1; return g; }

static inline bool aligned16_all(std::initializer_list<const void*> ps) {
    for (const void* p : ps) if (p && !aligned16(p)) return false;
    return true;
}
static inline int pick_group_vec(int nvec) { int g = 1; while (g < nvec && g < 32) g <<= 1; return g; }

// OP 0: coupling (p0 = s_raw, p1 = b_raw), OP 1: autoregressive (p0 = params, p1 unused)
template <typename T, int OP>
static int affine_rows_fwd_launch(const void* x, const void* p0, const void* p1, const void* mask, void* y, void* ld,
                                  int64_t B, int D, int mode, cudaStream_t st) {
    constexpr int V = Vec<T>::N;
    const bool al = aligned16_all({x, p0, p1, y});
    if (al && sizeof(T) == 4 && D == 2 && (B % 2) == 0 && (reinterpret_cast<uintptr_t>(ld) & 7) == 0) {
        if constexpr (sizeof(T) == 4) {
            const int grid = grid_for(B / 2, 256, 16);
            affine_rows_d2_fwd_kernel<OP><<<grid, 256, 0, st>>>((const float*)x, (const float*)p0, (const float*)p1,
                                                               (const float*)mask, (float*)y, (float*)ld, B, mode);
        }
    } else if (al && D % V == 0) {
        const int G = pick_group_vec(D / V);
        const int grid = grid_for(cdiv(B, 32 / G), 8, 16);
#define NF_AV(GG) affine_rows_vec_fwd_kernel<T, GG, OP><<<grid, 256, 0, st>>>((const T*)x, (const T*)p0, (const T*)p1, \
                                                                              (const T*)mask, (T*)y, (T*)ld, B, D, mode)
        switch (G) { case 1: NF_AV(1); break; case 2: NF_AV(2); break; case 4: NF_AV(4); break; case 8: NF_AV(8); break;
                     case 16: NF_AV(16); break; default: NF_AV(32); break; }
#undef NF_AV
    } else {
        const int G = pick_group(D);
        const int grid = grid_for(cdiv(B, 32 / G), 8, 32);
#define NF_AS(GG)                                                                                                      \
        do {                                                                                                           \
            if constexpr (OP == 0)                                                                                     \
                affine_coupling_fwd_kernel<T, GG><<<grid, 256, 0, st>>>((const T*)x, (const T*)p0, (const T*)p1,       \
                                                                        (const T*)mask, (T*)y, (T*)ld, B, D, mode);    \
            else                                                                                                       \
                affine_ar_fwd_kernel<T, GG><<<grid, 256, 0, st>>>((const T*)x, (const T*)p0, (T*)y, (T*)ld, B, D, mode); \
        } while (0)
        switch (G) { case 1: NF_AS(1); break; case 2: NF_AS(2); break; case 4: NF_AS(4); break; case 8: NF_AS(8); break;
                     case 16: NF_AS(16); break; default: NF_AS(32); break; }
#undef NF_AS
    }
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

template <typename T>
static int affine_coupling_fwd_launch(const void* x, const void* s, const void* b, const void* mask, void* y, void* ld,
                                      int64_t B, int D, int inverse, cudaStream_t st) {
    return affine_rows_fwd_launch<T, 0>(x, s, b, mask, y, ld, B, D, inverse, st);
}

template <typename T>
static int affine_ar_fwd_launch(const void* v, const void* params, void* out, void* ld, int64_t B, int D, int mode,
                                cudaStream_t st) {
    return affine_rows_fwd_launch<T, 1>(v, params, nullptr, nullptr, out, ld, B, D, mode, st);
}

template <typename T>
static int affine_coupling_bwd_launch(const void* x, const void* s, const void* b, const void* mask, const void* gy,
                                      const void* gld, void* gx, void* gs, void* gb, int64_t B, int D, int inverse,
                                      cudaStream_t st) {
    constexpr int V = Vec<T>::N;
    if (aligned16_all({x, s, b, gy, gx, gs, gb}) && (B * D) % V == 0) {
        const int grid = grid_for(B * D / V, 256, 16);
        affine_coupling_vec_bwd_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (const T*)s, (const T*)b, (const T*)mask,
                                                                (const T*)gy, (const T*)gld, (T*)gx, (T*)gs, (T*)gb, B, D, inverse);
    } else {
        const int grid = grid_for(B * D, 256, 32);
        affine_coupling_bwd_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (const T*)s, (const T*)b, (const T*)mask,
                                                            (const T*)gy, (const T*)gld, (T*)gx, (T*)gs, (T*)gb, B, D, inverse);
    }
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
    if (dtype == NF_F32) return affine_coupling_bwd_launch<float>(x, s_raw, b_raw, mask, gy, gld, gx, gs, gb, B, D, inverse, st);
    if (dtype == NF_F64) return affine_coupling_bwd_launch<double>(x, s_raw, b_raw, mask, gy, gld, gx, gs, gb, B, D, inverse, st);
    return NF_ERR_UNSUPPORTED;
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
