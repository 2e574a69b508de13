// skinny_kernels.cu -- dense products with one very small dimension (data_dim = 2..8 next to hidden_dim 64+):
// the first and last Linear of the conditioners of low-dimensional flows (coupling_layer.py:18-35,
// spline_coupling_layer.py:55-62, made.py:81-134) and their gradients.  A 128x128-tile GEMM wastes > 90 % of its
// work on them and, for the weight gradients, leaves a handful of CTAs walking the whole batch (measured: 201 us per
// launch at 5 000 rows, 55 % of a RealNVP(2,8,64) training step).  These kernels are streaming, HBM-bound:
//   skinny_reduce   C = sum_b Wd[b,:]^T (x) Sk[b,:]      one operand <= 32 columns wide, reduction over the batch
//                   (dW of the first / last Linear; bias gradients go through col_sum_small in dense_kernels.cu)
//   skinny_k        C[m,:] = sum_{k<=32} A[m,k] B[k,:]   (first Linear forward, input gradient of the last Linear / spline head)
//   skinny_n        C[m,n<=8] = A[m,:] . B[:,n]          (last Linear forward)
// Called from nf_gemm's dispatcher (dense_kernels.cu); same semantics as gemm_kernel (bias, ReLU, strides).
#include "nf_common.cuh"

namespace nf {

// ---- C[w,s] (or C[s,w]) = sum_b Wd[b*ldw + w] * Sk[b*lds + s];  S <= 8 -------------------------------------------------
// block = 8 warps sharing 128 columns of Wd (lane + 32 j), rows strided over the warps and over blockIdx.y chunks;
// per-thread accumulators [4][S]; block reduction in shared memory; chunks combine with atomics (C zeroed by the host)
template <typename T, int S, int JW>
__global__ void __launch_bounds__(256)
skinny_reduce_kernel(const T* __restrict__ Wd, const T* __restrict__ Sk, T* __restrict__ C, int64_t B, int W, int64_t ldw,
                     int64_t lds, int64_t so_w, int64_t so_s, int s_live, int64_t rows_per_chunk) {
    __shared__ T red[4][32 * JW * S + 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32 * JW;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t r1 = (B < r0 + rows_per_chunk) ? B : r0 + rows_per_chunk;
    T acc[JW][S];
#pragma unroll
    for (int j = 0; j < JW; ++j)
#pragma unroll
        for (int s = 0; s < S; ++s) acc[j][s] = T(0);
    // four rows per warp in flight when the accumulators leave room (one row per iteration: 36 % of HBM on the
    // [2^20, 64]^T x [2^20, 2] weight gradient of a 2-D conditioner's first Linear); same order of additions per row
    constexpr int U = (JW * S <= 16) ? 4 : 1;
    int64_t r = r0 + warp;
    for (; r + (U - 1) * 8 < r1; r += U * 8) {
        T sk[U][S], v[U][JW];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int s = 0; s < S; ++s) sk[u][s] = (s < s_live) ? Sk[(r + u * 8) * lds + s] : T(0);
#pragma unroll
            for (int j = 0; j < JW; ++j) { const int c = c0 + lane + 32 * j; v[u][j] = (c < W) ? Wd[(r + u * 8) * ldw + c] : T(0); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < JW; ++j)
#pragma unroll
                for (int s = 0; s < S; ++s) acc[j][s] += v[u][j] * sk[u][s];
    }
    for (; r < r1; r += 8) {
        T sk[S];
#pragma unroll
        for (int s = 0; s < S; ++s) sk[s] = (s < s_live) ? Sk[r * lds + s] : T(0);
        T v[JW];
#pragma unroll
        for (int j = 0; j < JW; ++j) { const int c = c0 + lane + 32 * j; v[j] = (c < W) ? Wd[r * ldw + c] : T(0); }
#pragma unroll
        for (int j = 0; j < JW; ++j)
#pragma unroll
            for (int s = 0; s < S; ++s) acc[j][s] += v[j] * sk[s];
    }
    if (warp >= 4) {
#pragma unroll
        for (int j = 0; j < JW; ++j)
#pragma unroll
            for (int s = 0; s < S; ++s) red[warp - 4][(lane + 32 * j) * S + s] = acc[j][s];
    }
    __syncthreads();
    if (warp < 4) {
#pragma unroll
        for (int j = 0; j < JW; ++j)
#pragma unroll
            for (int s = 0; s < S; ++s) red[warp][(lane + 32 * j) * S + s] += acc[j][s];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * JW * S; i += 256) {
        const int cl = i / S, s = i - cl * S, c = c0 + cl;
        if (c >= W || s >= s_live) continue;
        T t = T(0);
#pragma unroll
        for (int w = 0; w < 4; ++w) t += red[w][i];
        T* dst = C + (int64_t)c * so_w + (int64_t)s * so_s;
        if (gridDim.y == 1) *dst = t; else atomicAdd(dst, t);
    }
}

// ---- C[m, n] = act(sum_{k<K<=32} A[m*sam + k] * Bm[k*sbk + n*sbn] + bias[n]) ----------------------------------------------
// A rows are contiguous in k.  The block keeps the weights in shared memory as [k][n]; a thread owns one column and
// R = 4 rows per pass, so a k step is 1 LDS + 4 broadcast loads (the 4 rows' A values, shared by the row's threads through
// L1) + 4 FMAs with immediate offsets.  (First version: one output per thread, a 64-bit division and ~400 instructions
// per 32 outputs: 0.55 TB/s at K = 2 and 1.0 ms for a 2^20 x 64 x 23 product.)
template <typename T>
__global__ void __launch_bounds__(256)
skinny_k_kernel(const T* __restrict__ A, const T* __restrict__ Bm, T* __restrict__ C, const T* __restrict__ bias, int64_t M,
                int N, int K, int64_t sam, int64_t sbk, int64_t sbn, int64_t ldc, int relu, int accumulate, int tx) {
    extern __shared__ __align__(16) unsigned char skinny_smem[];
    T* ws = reinterpret_cast<T*>(skinny_smem);                   // [K][N] then bias [N]
    T* bs = ws + (size_t)K * N;
    for (int i = threadIdx.x; i < K * N; i += 256) { const int k = i / N, n = i - k * N; ws[i] = Bm[k * sbk + n * sbn]; }
    for (int i = threadIdx.x; i < N; i += 256) bs[i] = bias ? bias[i] : T(0);
    __syncthreads();
    constexpr int R = 4;
    const int ty = 256 / tx;                                     // tx: power of two >= min(N, 256)
    const int cx = threadIdx.x & (tx - 1), ry = threadIdx.x / tx;
    for (int64_t m0 = ((int64_t)blockIdx.x * ty + ry) * R; m0 < M; m0 += (int64_t)gridDim.x * ty * R) {
        const T* a0 = A + m0 * sam;
        const bool v1 = m0 + 1 < M, v2 = m0 + 2 < M, v3 = m0 + 3 < M;
        const T* a1 = v1 ? a0 + sam : a0;
        const T* a2 = v2 ? a0 + 2 * sam : a0;
        const T* a3 = v3 ? a0 + 3 * sam : a0;
        for (int n = cx; n < N; n += tx) {
            T c0 = bs[n], c1 = c0, c2 = c0, c3 = c0;
            const T* wp = ws + n;
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const T w = wp[k * N];
                c0 += a0[k] * w; c1 += a1[k] * w; c2 += a2[k] * w; c3 += a3[k] * w;
            }
            T* cp = C + m0 * ldc + n;
            if (accumulate) { c0 += cp[0]; if (v1) c1 += cp[ldc]; if (v2) c2 += cp[2 * ldc]; if (v3) c3 += cp[3 * ldc]; }
            if (relu) { c0 = relu_nan(c0); c1 = relu_nan(c1); c2 = relu_nan(c2); c3 = relu_nan(c3); }
            cp[0] = c0;
            if (v1) cp[ldc] = c1;
            if (v2) cp[2 * ldc] = c2;
            if (v3) cp[3 * ldc] = c3;
        }
    }
}

// ---- C[m, n<N<=8] = act(sum_k A[m*sam + k] * Bm[k*sbk + n*sbn] + bias[n]);  8 lanes per row, coalesced row reads ----------
template <typename T, int NS>
__global__ void __launch_bounds__(256)
skinny_n_kernel(const T* __restrict__ A, const T* __restrict__ Bm, T* __restrict__ C, const T* __restrict__ bias, int64_t M,
                int N, int K, int64_t sam, int64_t sbk, int64_t sbn, int64_t ldc, int relu, int accumulate) {
    const int sub = threadIdx.x & 7;
    const int64_t rows_per_pass = ((int64_t)gridDim.x * blockDim.x) >> 3;
    // every lane of a warp runs the same number of iterations (shuffles below): loop on the warp's first row
    for (int64_t m0 = (((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) >> 3); m0 < M; m0 += rows_per_pass) {
        const int64_t m = m0 + ((threadIdx.x & 31) >> 3);
        T acc[NS];
#pragma unroll
        for (int n = 0; n < NS; ++n) acc[n] = T(0);
        if (m < M) {
            const T* ar = A + m * sam;
            for (int k = sub; k < K; k += 8) {
                const T a = ar[k];
#pragma unroll
                for (int n = 0; n < NS; ++n) if (n < N) acc[n] += a * __ldg(Bm + k * sbk + n * sbn);
            }
        }
#pragma unroll
        for (int n = 0; n < NS; ++n) acc[n] = group_sum<T, 8>(acc[n]);
        if (m < M && sub == 0) {
#pragma unroll
            for (int n = 0; n < NS; ++n) if (n < N) {
                T v = acc[n] + (bias ? __ldg(bias + n) : T(0));
                T* cp = C + m * ldc + n;
                if (accumulate) v += *cp;
                if (relu) v = relu_nan(v);
                *cp = v;
            }
        }
    }
}

// ---- float32, 128-bit versions of the two forward kernels ---------------------------------------------------------------
// skinny_k above writes 16 bytes per thread between two dependent broadcast loads and skinny_n reads 4 bytes per lane and
// iteration: 2.5 TB/s = 38 % of HBM on [2^20, 2] x [2, 256] and [2^20, 256] x [256, 2] (the first / last Linear of
// RealNVP(2, 8, 256), profiles/r02ag_nb_realnvp256_launch_summary.txt).  Here a thread of the K <= 8 kernel owns FOUR
// columns and EIGHT rows per pass (eight 128-bit stores per thread behind one round of row loads), and a lane of the
// N <= 8 kernel reads 16 bytes per load with every load of its row slice in flight before the first FMA.
template <int KS, int R>
__global__ void __launch_bounds__(256)
skinny_k_vec4_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C,
                     const float* __restrict__ bias, int64_t M, int N, int K, int64_t sam, int64_t sbk, int64_t sbn, int64_t ldc,
                     int relu, int tx) {
    extern __shared__ __align__(16) unsigned char skinny_smem[];
    float* ws = reinterpret_cast<float*>(skinny_smem);           // [K][N] then bias [N]
    float* bs = ws + (size_t)K * N;
    for (int i = threadIdx.x; i < K * N; i += 256) { const int k = i / N, n = i - k * N; ws[i] = Bm[k * sbk + n * sbn]; }
    for (int i = threadIdx.x; i < N; i += 256) bs[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int ty = 256 / tx;                                     // tx: power of two >= min(N / 4, 256)
    const int cx = threadIdx.x & (tx - 1), ry = threadIdx.x / tx;
    const int N4 = N >> 2;
    for (int64_t m0 = ((int64_t)blockIdx.x * ty + ry) * R; m0 < M; m0 += (int64_t)gridDim.x * ty * R) {
        float a[R][KS];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int64_t m = (m0 + i < M) ? m0 + i : M - 1;    // rows beyond the end recompute the last row (not stored)
#pragma unroll
            for (int k = 0; k < KS; ++k) a[i][k] = (k < K) ? __ldg(A + m * sam + k) : 0.f;
        }
        for (int n4 = cx; n4 < N4; n4 += tx) {
            const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * n4);
            float4 c[R];
#pragma unroll
            for (int i = 0; i < R; ++i) c[i] = b4;
#pragma unroll
            for (int k = 0; k < KS; ++k) if (k < K) {
                const float4 w = *reinterpret_cast<const float4*>(ws + (size_t)k * N + 4 * n4);
#pragma unroll
                for (int i = 0; i < R; ++i) {                    // two packed FMAs per row and k (same operations)
                    const float2 aa = make_float2(a[i][k], a[i][k]);
                    const float2 lo = __ffma2_rn(aa, make_float2(w.x, w.y), make_float2(c[i].x, c[i].y));
                    const float2 hi = __ffma2_rn(aa, make_float2(w.z, w.w), make_float2(c[i].z, c[i].w));
                    c[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
                }
            }
#pragma unroll
            for (int i = 0; i < R; ++i) if (m0 + i < M) {
                float4 v = c[i];
                if (relu) { v.x = relu_nan(v.x); v.y = relu_nan(v.y); v.z = relu_nan(v.z); v.w = relu_nan(v.w); }
                *reinterpret_cast<float4*>(C + (m0 + i) * ldc + 4 * n4) = v;
            }
        }
    }
}

// K in (8, 32] (the input gradient of a 23 / 29-wide spline head: dX[B, 64] = dY[B, 23] W): a thread owns ONE ROW -- its K
// values sit in registers, the weights arrive as warp-uniform 128-bit broadcasts from shared memory, and the row leaves as
// full 32-byte sectors.  (A thread per four columns needs 4 x K row values per thread: one CTA per SM with ~6 KB of loads
// in flight, 20 % of HBM.)  Columns in blocks of NT.
template <int KS, int NT>
__global__ void __launch_bounds__(256, 2)
skinny_k_rows_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C,
                     const float* __restrict__ bias, int64_t M, int N, int K, int64_t sbk, int64_t sbn, int64_t ldc, int relu,
                     int wide_st) {
    extern __shared__ __align__(16) unsigned char skinny_smem[];
    float* ws = reinterpret_cast<float*>(skinny_smem);           // [K][N], bias [N], then the CTA's 256 rows of A [256][K]
    float* bs = ws + (size_t)K * N;
    float* sA = bs + N;
    for (int i = threadIdx.x; i < K * N; i += 256) { const int k = i / N, n = i - k * N; ws[i] = Bm[k * sbk + n * sbn]; }
    for (int i = threadIdx.x; i < N; i += 256) bs[i] = bias ? bias[i] : 0.f;
    for (int64_t m0 = (int64_t)blockIdx.x * 256; m0 < M; m0 += (int64_t)gridDim.x * 256) {
        __syncthreads();                                         // weights staged / the previous rows are consumed
        // the block's rows are one contiguous run of A (row pitch == K): coalesced copy, 128-bit when the run allows
        const int nrow = (int)((M - m0) < 256 ? (M - m0) : 256);
        const int nfl = nrow * K;
        const float* src = A + m0 * K;
        if (((reinterpret_cast<uintptr_t>(src) | (uintptr_t)(nfl * 4)) & 15) == 0) {
            for (int i = threadIdx.x * 4; i < nfl; i += 1024)
                *reinterpret_cast<float4*>(sA + i) = __ldcs(reinterpret_cast<const float4*>(src + i));
        } else {
            for (int i = threadIdx.x; i < nfl; i += 256) sA[i] = __ldcs(src + i);
        }
        __syncthreads();
        if ((int)threadIdx.x >= nrow) continue;
        float a[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) a[k] = (k < K) ? sA[threadIdx.x * K + k] : 0.f;     // pitch K: odd or not, <= 2-way
        float* crow = C + (m0 + threadIdx.x) * ldc;
        for (int n0 = 0; n0 < N; n0 += NT) {                     // N % NT == 0 (launcher)
            float2 c[NT / 2];
#pragma unroll
            for (int j = 0; j < NT / 4; ++j) {
                const float4 b4 = *reinterpret_cast<const float4*>(bs + n0 + 4 * j);
                c[2 * j] = make_float2(b4.x, b4.y); c[2 * j + 1] = make_float2(b4.z, b4.w);
            }
#pragma unroll
            for (int k = 0; k < KS; ++k) if (k < K) {
                const float2 aa = make_float2(a[k], a[k]);
                const float* wr = ws + (size_t)k * N + n0;
#pragma unroll
                for (int j = 0; j < NT / 4; ++j) {
                    const float4 w = *reinterpret_cast<const float4*>(wr + 4 * j);
                    c[2 * j] = __ffma2_rn(aa, make_float2(w.x, w.y), c[2 * j]);
                    c[2 * j + 1] = __ffma2_rn(aa, make_float2(w.z, w.w), c[2 * j + 1]);
                }
            }
#pragma unroll
            for (int j = 0; j < NT / 8; ++j) {
                float o[8] = {c[4 * j].x, c[4 * j].y, c[4 * j + 1].x, c[4 * j + 1].y, c[4 * j + 2].x, c[4 * j + 2].y, c[4 * j + 3].x, c[4 * j + 3].y};
                if (relu) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) o[q] = relu_nan(o[q]);
                }
                if (wide_st) {                                   // one full 32-byte sector per store
                    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(crow + n0 + 8 * j), "f"(o[0]), "f"(o[1]),
                                 "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
                } else {
                    *reinterpret_cast<float4*>(crow + n0 + 8 * j) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4*>(crow + n0 + 8 * j + 4) = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
        }
    }
}

// C[m, n < N <= NS] = act(A[m, :] . W[:, n] + bias[n]): 8 lanes per row, weights transposed in shared memory ([n][K], one
// 128-bit broadcast-free read per lane and output), U 16-byte row loads in flight per lane
template <int NS, int LPR>          // LPR lanes per row: 8, or 4 for short rows (K <= 64: keeps four loads per lane in flight)
__global__ void __launch_bounds__(256)
skinny_n_vec4_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C,
                     const float* __restrict__ bias, int64_t M, int N, int K, int64_t sam, int64_t sbk, int64_t sbn, int64_t ldc,
                     int relu, int accumulate) {
    extern __shared__ __align__(16) unsigned char skinny_smem[];
    float* wt = reinterpret_cast<float*>(skinny_smem);           // [NS][K]
    for (int i = threadIdx.x; i < NS * K; i += 256) { const int n = i / K, k = i - n * K; wt[i] = (n < N) ? Bm[k * sbk + n * sbn] : 0.f; }
    __syncthreads();
    constexpr int U = 4;
    constexpr int SH = (LPR == 8) ? 3 : 2;
    const int sub = threadIdx.x & (LPR - 1);
    const int K4 = K >> 2;
    const int64_t rows_per_pass = ((int64_t)gridDim.x * blockDim.x) >> SH;
    for (int64_t m0 = (((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) >> SH); m0 < M; m0 += rows_per_pass) {
        const int64_t m = m0 + ((threadIdx.x & 31) >> SH);
        float acc[NS];
#pragma unroll
        for (int n = 0; n < NS; ++n) acc[n] = 0.f;
        if (m < M) {
            const float4* ar = reinterpret_cast<const float4*>(A + m * sam);
            for (int c0 = sub; c0 < K4; c0 += LPR * U) {
                float4 av[U];
#pragma unroll
                for (int u = 0; u < U; ++u) { const int c = c0 + LPR * u; av[u] = (c < K4) ? __ldcs(ar + c) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + LPR * u;
                    if (c < K4) {
#pragma unroll
                        for (int n = 0; n < NS; ++n) {
                            const float4 w = *reinterpret_cast<const float4*>(wt + (size_t)n * K + 4 * c);
                            acc[n] += av[u].x * w.x; acc[n] += av[u].y * w.y; acc[n] += av[u].z * w.z; acc[n] += av[u].w * w.w;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int n = 0; n < NS; ++n) acc[n] = group_sum<float, LPR>(acc[n]);
        if (m < M && sub == 0) {
#pragma unroll
            for (int n = 0; n < NS; ++n) if (n < N) {
                float v = acc[n] + (bias ? __ldg(bias + n) : 0.f);
                float* cp = C + m * ldc + n;
                if (accumulate) v += *cp;
                if (relu) v = relu_nan(v);
                *cp = v;
            }
        }
    }
}

// column sums of a[rows, cols], cols <= CM: every thread walks whole rows with `cols` accumulators (a warp reads 32
// consecutive rows = one contiguous run); block reduction, then atomics across blocks (out zeroed by the host)
template <typename T, int CM>
__global__ void __launch_bounds__(256)
col_sum_small_kernel(const T* __restrict__ a, T* __restrict__ out, int64_t rows, int cols, int64_t rows_per_block) {
    __shared__ double red[8][CM];
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = (rows < r0 + rows_per_block) ? rows : r0 + rows_per_block;
    double acc[CM];
#pragma unroll
    for (int c = 0; c < CM; ++c) acc[c] = 0.0;
    // four rows per thread in flight (one row per iteration left the kernel latency-bound: 38 % of HBM at cols = 2)
    int64_t r = r0 + threadIdx.x;
    if constexpr (CM <= 8)
    for (; r + 3 * 256 < r1; r += 4 * 256) {
        T v[4][CM];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < CM; ++c) v[u][c] = (c < cols) ? a[(r + u * 256) * cols + c] : T(0);
#pragma unroll
        for (int c = 0; c < CM; ++c) acc[c] += ((double)v[0][c] + (double)v[1][c]) + ((double)v[2][c] + (double)v[3][c]);
    }
    for (; r < r1; r += 256) {
#pragma unroll
        for (int c = 0; c < CM; ++c) if (c < cols) acc[c] += (double)a[r * cols + c];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < CM; ++c) {
        const double t = warp_sum<double>(acc[c]);
        if (lane == 0) red[warp][c] = t;
    }
    __syncthreads();
    if (threadIdx.x < cols) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        if (gridDim.x == 1) out[threadIdx.x] = (T)t; else atomicAdd(out + threadIdx.x, (T)t);
    }
}

// ---- dispatcher used by nf_gemm ------------------------------------------------------------------------------------------
// returns 1 when the product was handled here (launch issued), 0 when the shape is not skinny, < 0 on error
template <typename T>
int skinny_gemm_try(const void* A, const void* Bm, void* C, const void* bias, int64_t M, int64_t N, int64_t K, int64_t sam,
                    int64_t sak, int64_t sbk, int64_t sbn, int64_t ldc, int relu, int accumulate, const int32_t* k_extent,
                    cudaStream_t st) {
    if (k_extent) return 0;
    // (1) reduction over a long batch with one narrow operand: C[M,N] = sum_k A[m + k*sak] * Bm[k*sbk + n]
    if (!bias && !relu && !accumulate && sam == 1 && sbn == 1 && K >= 256 && (M <= 32 || N <= 32) && M * N <= (1 << 20)) {
        const bool narrow_b = N <= M;                             // Sk = Bm (S = N), Wd = A (W = M); else the transpose
        const T* Wd = (const T*)(narrow_b ? A : Bm);
        const T* Sk = (const T*)(narrow_b ? Bm : A);
        const int W = (int)(narrow_b ? M : N), S = (int)(narrow_b ? N : M);
        const int64_t ldw = narrow_b ? sak : sbk, lds = narrow_b ? sbk : sak;
        const int64_t so_w = narrow_b ? ldc : 1, so_s = narrow_b ? 1 : ldc;
        const int JW = S <= 8 ? 4 : 1;                            // columns of Wd per lane (accumulators: JW x S)
        const int64_t cblocks = cdiv(W, 32 * JW);
        int64_t ch = cdiv((int64_t)kNumSMs * 4, cblocks);
        const int64_t chmax = K / 64 > 1 ? K / 64 : 1;           // at least 64 rows (8 per warp) per chunk
        if (ch > chmax) ch = chmax;
        const int64_t rpc = cdiv(K, ch);
        ch = cdiv(K, rpc);
        if (ch > 65535) return 0;
        if (ch > 1) {
            if (ldc == N) NF_CUDA(cudaMemsetAsync(C, 0, sizeof(T) * M * N, st));
            else NF_CUDA(cudaMemset2DAsync(C, sizeof(T) * ldc, 0, sizeof(T) * N, M, st));
        }
        dim3 grid((unsigned)cblocks, (unsigned)ch);
#define NF_SR(SS, JJ) skinny_reduce_kernel<T, SS, JJ><<<grid, 256, 0, st>>>(Wd, Sk, (T*)C, K, W, ldw, lds, so_w, so_s, S, rpc)
        if (S <= 2) NF_SR(2, 4); else if (S <= 4) NF_SR(4, 4); else if (S <= 8) NF_SR(8, 4);
        else if (S <= 16) NF_SR(16, 1); else if (S <= 24) NF_SR(24, 1); else NF_SR(32, 1);
#undef NF_SR
        return 1;
    }
    // (2) small reduction dimension, contiguous A rows: a thread per output column, 4 rows per pass
    if constexpr (sizeof(T) == 4) {
        // K in (8, 32] (the 23 / 29-wide spline heads' input gradient): a row per thread
        if (K > 8 && K <= 32 && sak == 1 && sam == K && M >= 256 && !accumulate && (N % 32) == 0 && (ldc % 4) == 0 && aligned16(C) &&
            (size_t)((K + 1) * N + 256 * K) * sizeof(T) <= 96 * 1024) {
            const size_t smem = (size_t)((K + 1) * N + 256 * K) * sizeof(T);
            int64_t g = cdiv(M, 256);
            const int64_t cap = (int64_t)kNumSMs * 2;
            const int grid = (int)(g < cap ? g : cap);
            const int wide_st = ((ldc % 8) == 0 && aligned32(C)) ? 1 : 0;
#define NF_SKR(KSv)                                                                                                           \
            do {                                                                                                              \
                if (smem > 48 * 1024) NF_CUDA(cudaFuncSetAttribute(skinny_k_rows_kernel<KSv, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                skinny_k_rows_kernel<KSv, 32><<<grid, 256, smem, st>>>((const float*)A, (const float*)Bm, (float*)C, (const float*)bias, \
                                                                        M, (int)N, (int)K, sbk, sbn, ldc, relu, wide_st);          \
            } while (0)
            if (K <= 16) NF_SKR(16); else if (K <= 24) NF_SKR(24); else NF_SKR(32);
#undef NF_SKR
            return 1;
        }
        // K <= 8: eight rows per thread, four columns
        if (K >= 1 && K <= 8 && sak == 1 && M >= 64 && !accumulate && (N % 4) == 0 && (ldc % 4) == 0 && aligned16(C) &&
            (size_t)(K + 1) * N * sizeof(T) <= 40 * 1024) {
            int tx = 1;
            while (tx < N / 4 && tx < 256) tx <<= 1;
            const int ty = 256 / tx;
            const int Rv = 8;
            int64_t g = cdiv(M, (int64_t)ty * Rv * 2);
            if (g < 1) g = 1;
            const size_t smem = (size_t)(K + 1) * N * sizeof(T);
            // one wave of resident CTAs
#define NF_SKV(KSv, RRv)                                                                                                  \
            do {                                                                                                          \
                int per_sm = 0;                                                                                           \
                NF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, skinny_k_vec4_kernel<KSv, RRv>, 256, smem)); \
                const int64_t cap = (int64_t)kNumSMs * (per_sm < 1 ? 1 : per_sm);                                          \
                skinny_k_vec4_kernel<KSv, RRv><<<(int)(g < cap ? g : cap), 256, smem, st>>>(                               \
                    (const float*)A, (const float*)Bm, (float*)C, (const float*)bias, M, (int)N, (int)K, sam, sbk, sbn, ldc, relu, tx); \
            } while (0)
            if (K <= 2) NF_SKV(2, 8); else if (K <= 4) NF_SKV(4, 8); else NF_SKV(8, 8);
#undef NF_SKV
            return 1;
        }
    }
    if (K >= 1 && K <= 32 && sak == 1 && M >= 64 && (K <= 8 || (K % 4) != 0) && (size_t)(K + 1) * N * sizeof(T) <= 40 * 1024) {
        int tx = 1;
        while (tx < N && tx < 256) tx <<= 1;
        const int ty = 256 / tx;
        int64_t g = cdiv(M, (int64_t)ty * 4 * 4), cap = (int64_t)kNumSMs * 8;
        if (g < 1) g = 1;
        const size_t smem = (size_t)(K + 1) * N * sizeof(T);
        skinny_k_kernel<T><<<(int)(g < cap ? g : cap), 256, smem, st>>>((const T*)A, (const T*)Bm, (T*)C, (const T*)bias, M, (int)N,
                                                                         (int)K, sam, sbk, sbn, ldc, relu, accumulate, tx);
        return 1;
    }
    // (3) narrow output: per-row dot products, contiguous A rows
    if (N <= 8 && sak == 1 && M >= 64 && K >= 16) {
        int64_t g = cdiv(M * 8, 256), cap = (int64_t)kNumSMs * 16;
        const int grid = (int)(g < cap ? g : cap);
        if constexpr (sizeof(T) == 4) {
            if ((K % 4) == 0 && (sam % 4) == 0 && aligned16(A) && (size_t)8 * K * sizeof(T) <= 48 * 1024) {
                const int NSv = N <= 2 ? 2 : (N <= 4 ? 4 : 8);
                const size_t smem = (size_t)NSv * K * sizeof(T);
                const int64_t g4 = cdiv(M * 4, 256);
                const int grid4 = (int)(g4 < cap ? g4 : cap);
#define NF_SNV(NS)                                                                                                            \
                do {                                                                                                          \
                    if (K <= 64) skinny_n_vec4_kernel<NS, 4><<<grid4, 256, smem, st>>>((const float*)A, (const float*)Bm, (float*)C, \
                                     (const float*)bias, M, (int)N, (int)K, sam, sbk, sbn, ldc, relu, accumulate);             \
                    else skinny_n_vec4_kernel<NS, 8><<<grid, 256, smem, st>>>((const float*)A, (const float*)Bm, (float*)C,     \
                                     (const float*)bias, M, (int)N, (int)K, sam, sbk, sbn, ldc, relu, accumulate);             \
                } while (0)
                if (N <= 2) NF_SNV(2); else if (N <= 4) NF_SNV(4); else NF_SNV(8);
#undef NF_SNV
                return 1;
            }
        }
#define NF_SN(NS) skinny_n_kernel<T, NS><<<grid, 256, 0, st>>>((const T*)A, (const T*)Bm, (T*)C, (const T*)bias, M, (int)N, (int)K, \
                                                             sam, sbk, sbn, ldc, relu, accumulate)
        if (N <= 2) NF_SN(2); else if (N <= 4) NF_SN(4); else NF_SN(8);
#undef NF_SN
        return 1;
    }
    return 0;
}

template int skinny_gemm_try<float>(const void*, const void*, void*, const void*, int64_t, int64_t, int64_t, int64_t, int64_t,
                                    int64_t, int64_t, int64_t, int, int, const int32_t*, cudaStream_t);
template int skinny_gemm_try<double>(const void*, const void*, void*, const void*, int64_t, int64_t, int64_t, int64_t, int64_t,
                                     int64_t, int64_t, int64_t, int, int, const int32_t*, cudaStream_t);

template <typename T>
int col_sum_small_launch(const void* a, void* out, int64_t rows, int cols, cudaStream_t st) {
    int64_t blocks = cdiv(rows, 2048);
    const int64_t cap = (int64_t)kNumSMs * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const int64_t rpb = cdiv(rows, blocks);
    blocks = cdiv(rows, rpb);
    if (blocks > 1) NF_CUDA(cudaMemsetAsync(out, 0, sizeof(T) * cols, st));
    if (cols <= 8) col_sum_small_kernel<T, 8><<<(int)blocks, 256, 0, st>>>((const T*)a, (T*)out, rows, cols, rpb);
    else col_sum_small_kernel<T, 32><<<(int)blocks, 256, 0, st>>>((const T*)a, (T*)out, rows, cols, rpb);
    return NF_OK;
}
template int col_sum_small_launch<float>(const void*, void*, int64_t, int, cudaStream_t);
template int col_sum_small_launch<double>(const void*, void*, int64_t, int, cudaStream_t);

}  // namespace nf
