// transform_impl.cuh -- HBM-bound spline transform kernels (parameters already in HBM):
//   rqs_unit_*            a7    rational_quadratic_spline            (rational_quadratic_spline.py:4-104)
//   spline_transform_*    a5/a6 SplineCouplingLayer transform         (spline_coupling_layer.py:96-309)
// Streaming kernels: every byte is touched once, the roofline is HBM.
// Included by the per-dtype translation units (rqs_unit_*.cu, spline_tf_*.cu) for the register path
// (num_bins<=16, bin loops fully unrolled) and by transform_generic.cu for num_bins in (16,32].
#pragma once
#include "nf_common.cuh"

namespace nf {


// ------------------------------------------------------------------------------------------------
// vectorised row loads: N contiguous T starting at p into registers.  VEC = widest power-of-two
// byte width (<=16) dividing N*sizeof(T) *and* guaranteed by the caller to divide the row address.
// ------------------------------------------------------------------------------------------------
template <typename T, int N, int VEC>
__device__ __forceinline__ void load_row(const T* __restrict__ p, T* out) {
    constexpr int E = VEC / (int)sizeof(T);   // elements per vector
    if constexpr (E >= 4 && sizeof(T) == 4) {
NF_UNROLL
        for (int i = 0; i < N / 4; ++i) {
            float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
            out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
        }
    } else if constexpr (E == 2 && sizeof(T) == 4) {
NF_UNROLL
        for (int i = 0; i < N / 2; ++i) {
            float2 v = __ldg(reinterpret_cast<const float2*>(p) + i);
            out[2 * i] = v.x; out[2 * i + 1] = v.y;
        }
    } else if constexpr (E == 2 && sizeof(T) == 8) {
NF_UNROLL
        for (int i = 0; i < N / 2; ++i) {
            double2 v = __ldg(reinterpret_cast<const double2*>(p) + i);
            out[2 * i] = v.x; out[2 * i + 1] = v.y;
        }
    } else {
NF_UNROLL
        for (int i = 0; i < N; ++i) out[i] = __ldg(p + i);
    }
}

template <typename T, int KMAX>
__device__ __forceinline__ void load_row_rt(const T* __restrict__ p, int n, T* out) {
NF_UNROLL
    for (int i = 0; i < KMAX; ++i) out[i] = (i < n) ? __ldg(p + i) : T(0);
}

__host__ __device__ constexpr int vec_bytes(int n_elems, int elem_size) {
    int b = n_elems * elem_size;
    return (b % 16 == 0) ? 16 : (b % 8 == 0) ? 8 : elem_size;
}

// ------------------------------------------------------------------------------------------------
// a7 forward.  One thread per element; widths/heights rows are fetched with the widest aligned
// vector loads (K=8: 2x LDG.128 per array), neighbouring lanes share 128B lines through L1, so DRAM
// sees every byte once.  Algorithmic bytes per element: 4*(3K-1) params + 4 in + 8 out.
// ------------------------------------------------------------------------------------------------
template <typename T, int KMAX, bool SK>
__global__ void __launch_bounds__(256)
rqs_unit_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ h,
                    const T* __restrict__ d, T* __restrict__ y, T* __restrict__ ld, int64_t n, int Krt, int inverse,
                    RqsCfg<T> c) {
    const int K = SK ? KMAX : Krt;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T uw[KMAX], uh[KMAX], ud[KMAX];
        if constexpr (SK) {
            load_row<T, KMAX, vec_bytes(KMAX, sizeof(T))>(w + i * KMAX, uw);
            load_row<T, KMAX, vec_bytes(KMAX, sizeof(T))>(h + i * KMAX, uh);
            load_row<T, KMAX - 1, (int)sizeof(T)>(d + i * (KMAX - 1), ud);
            ud[KMAX - 1] = T(0);
        } else {
            load_row_rt<T, KMAX>(w + i * K, K, uw);
            load_row_rt<T, KMAX>(h + i * K, K, uh);
            load_row_rt<T, KMAX>(d + i * (K - 1), K - 1, ud);
        }
        T out, lad;
        rqs_eval<T, KMAX, false>(ld_stream(x + i), uw, uh, ud, K, inverse != 0, c, out, lad);
        st_stream(y + i, out);
        st_stream(ld + i, lad);
    }
}

template <typename T, int KMAX, bool SK>
__global__ void __launch_bounds__(128)
rqs_unit_bwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ h,
                    const T* __restrict__ d, const T* __restrict__ gy, const T* __restrict__ gld,
                    T* __restrict__ gx, T* __restrict__ gw, T* __restrict__ gh, T* __restrict__ gd, int64_t n,
                    int Krt, int inverse, RqsCfg<T> c) {
    const int K = SK ? KMAX : Krt;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T uw[KMAX], uh[KMAX], ud[KMAX];
        load_row_rt<T, KMAX>(w + i * K, K, uw);
        load_row_rt<T, KMAX>(h + i * K, K, uh);
        load_row_rt<T, KMAX>(d + i * (K - 1), K - 1, ud);
        T guw[KMAX], guh[KMAX], gud[KMAX];
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) { guw[j] = T(0); guh[j] = T(0); gud[j] = T(0); }
        T gv = T(0);
        rqs_eval_bwd<T, KMAX, false>(x[i], uw, uh, ud, K, inverse != 0, c, gy[i], gld[i], gv, guw, guh, gud);
        gx[i] = gv;
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) if (j < K) { gw[i * K + j] = guw[j]; gh[i * K + j] = guh[j]; }
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) if (j < K - 1) gd[i * (K - 1) + j] = gud[j];
    }
}

// a7 backward with shared-memory staging: a warp handles 32 consecutive elements; their width / height / derivative
// rows are three contiguous runs of 32K, 32K and 32(K-1) values, fetched with coalesced cp.async into the warp's
// slab, and the three gradient runs leave through the same slab with coalesced stores (the register-path kernel
// above issues 3K-1 strided 4-byte loads and stores per element: 12 % of HBM).
template <typename T, int KMAX, bool SK>
__global__ void __launch_bounds__(128)
rqs_unit_bwd_slab_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ h,
                         const T* __restrict__ d, const T* __restrict__ gy, const T* __restrict__ gld,
                         T* __restrict__ gx, T* __restrict__ gw, T* __restrict__ gh, T* __restrict__ gd, int64_t n,
                         int Krt, int inverse, RqsCfg<T> c) {
    extern __shared__ __align__(16) unsigned char slab_raw[];
    const int K = SK ? KMAX : Krt;
    const int P = 3 * K - 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    T* sw = reinterpret_cast<T*>(slab_raw) + (size_t)wib * 32 * P;
    T* sh = sw + 32 * K;
    T* sd = sh + 32 * K;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nblk = (n + 31) / 32;
    for (int64_t blk = warp; blk < nblk; blk += nwarps) {
        const int64_t i0 = blk * 32;
        const int cnt = (int)((n - i0) < 32 ? (n - i0) : 32);
        const int nk = cnt * K, nd = cnt * (K - 1);
        __syncwarp();
        if constexpr (sizeof(T) == 4) {
            for (int i = lane; i < nk; i += 32) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(sw + i))), "l"(w + i0 * K + i));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(sh + i))), "l"(h + i0 * K + i));
            }
            for (int i = lane; i < nd; i += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(sd + i))), "l"(d + i0 * (K - 1) + i));
            asm volatile("cp.async.commit_group;\n" ::);
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        } else {
            for (int i = lane; i < nk; i += 32) { sw[i] = __ldcs(w + i0 * K + i); sh[i] = __ldcs(h + i0 * K + i); }
            for (int i = lane; i < nd; i += 32) sd[i] = __ldcs(d + i0 * (K - 1) + i);
        }
        __syncwarp();
        if (lane < cnt) {
            const int64_t i = i0 + lane;
            T uw[KMAX], uh[KMAX], ud[KMAX];
NF_UNROLL
            for (int j = 0; j < KMAX; ++j) {
                uw[j] = (j < K) ? sw[lane * K + j] : T(0);
                uh[j] = (j < K) ? sh[lane * K + j] : T(0);
                ud[j] = (j < K - 1) ? sd[lane * (K - 1) + j] : T(0);
            }
            T guw[KMAX], guh[KMAX], gud[KMAX];
NF_UNROLL
            for (int j = 0; j < KMAX; ++j) { guw[j] = T(0); guh[j] = T(0); gud[j] = T(0); }
            T gv = T(0);
            rqs_eval_bwd<T, KMAX, false>(ld_stream(x + i), uw, uh, ud, K, inverse != 0, c, ld_stream(gy + i), ld_stream(gld + i), gv, guw, guh, gud);
            st_stream(gx + i, gv);
NF_UNROLL
            for (int j = 0; j < KMAX; ++j) if (j < K) { sw[lane * K + j] = guw[j]; sh[lane * K + j] = guh[j]; }
NF_UNROLL
            for (int j = 0; j < KMAX; ++j) if (j < K - 1) sd[lane * (K - 1) + j] = gud[j];
        }
        __syncwarp();
        for (int i = lane; i < nk; i += 32) { st_stream(gw + i0 * K + i, sw[i]); st_stream(gh + i0 * K + i, sh[i]); }
        for (int i = lane; i < nd; i += 32) st_stream(gd + i0 * (K - 1) + i, sd[i]);
    }
}

// ------------------------------------------------------------------------------------------------
// a5/a6: spline coupling transform from interleaved params [B, D*P].  G lanes cooperate on one row
// (G = pow2 >= min(Dt,32)); each lane walks transformed dims t = g, g+G, ...; row log-det via
// shuffles.  Identity dims are copied (with the layer-level NaN/Inf scrub) by the same lanes.
// ------------------------------------------------------------------------------------------------
template <typename T, int KMAX, bool SK, int G>
__global__ void __launch_bounds__(128)
spline_transform_fwd_kernel(const T* __restrict__ x, const T* __restrict__ params, const T* __restrict__ mask,
                            const int32_t* __restrict__ tidx, T* __restrict__ y, T* __restrict__ ld, int64_t B, int D,
                            int Dt, int Krt, int inverse, RqsCfg<T> c, const T* __restrict__ r_in,
                            const T* __restrict__ r_lo, const T* __restrict__ r_out, int compact) {
    const int K = SK ? KMAX : Krt;
    const int P = 3 * K - 1;
    const int PD = compact ? Dt : D;                   // parameter blocks per row
    constexpr int RPW = 32 / G;                        // rows per warp
    const int lane = threadIdx.x & 31, g = lane % G;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nblk = (B + RPW - 1) / RPW;
    for (int64_t blk = warp; blk < nblk; blk += nwarps) {
        const int64_t row = blk * RPW + lane / G;
        const bool valid = row < B;
        T acc = T(0);
        if (valid) {
            const T* xr = x + row * D;
            T* yr = y + row * D;
            for (int dd = g; dd < D; dd += G)
                if (__ldg(mask + dd) != T(0)) yr[dd] = scrub0(xr[dd]);
            for (int t = g; t < Dt; t += G) {
                const int dim = __ldg(tidx + t);
                T v = xr[dim];
                if (r_in) v = r_in[dim] * (v - r_lo[dim]) - c.hi;
                const T* pp = params + (row * PD + (compact ? t : dim)) * P;
                T uw[KMAX], uh[KMAX], ud[KMAX];
                load_row_rt<T, KMAX>(pp, K, uw);
                load_row_rt<T, KMAX>(pp + K, K, uh);
                load_row_rt<T, KMAX>(pp + 2 * K, K - 1, ud);
                T out, lad;
                rqs_eval<T, KMAX, true>(v, uw, uh, ud, K, inverse != 0, c, out, lad);
                if (r_in) out = (out + c.hi) * r_out[dim] + r_lo[dim];
                yr[dim] = scrub0(out);
                acc += lad;
            }
        }
        acc = group_sum<T, G>(acc);
        if (valid && g == 0) ld[row] = scrub0(acc);
    }
}

// ------------------------------------------------------------------------------------------------
// a5/a6 forward, compact parameter layout [B, Dt*(3K-1)]: the parameter blocks of consecutive (row, t) elements are
// contiguous, so a warp fetches the blocks of its 32 elements with fully coalesced loads (one 128-byte line per
// instruction) into a per-warp shared-memory slab and every lane then reads its own block at stride 3K-1 (odd:
// bank-conflict free).  The register-path kernel above issues one strided LDG per parameter instead (3K-1 LDGs
// touching ~24 lines each), which makes it L1-wavefront bound at ~1/3 of this kernel's rate.
// G lanes per row (G = pow2 >= min(Dt,32)); Dt > 32 walks the row in chunks of 32 elements.
// ------------------------------------------------------------------------------------------------
template <typename T, int KMAX, bool SK, int G>
__global__ void __launch_bounds__(256)
spline_transform_compact_fwd_kernel(const T* __restrict__ x, const T* __restrict__ params, const T* __restrict__ mask,
                                    const int32_t* __restrict__ tidx, T* __restrict__ y, T* __restrict__ ld, int64_t B,
                                    int D, int Dt, int Krt, int inverse, RqsCfg<T> c, const T* __restrict__ r_in,
                                    const T* __restrict__ r_lo, const T* __restrict__ r_out) {
    extern __shared__ __align__(16) unsigned char slab_raw[];
    const int K = SK ? KMAX : Krt;
    const int P = 3 * K - 1;
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = lane % G, rsub = lane / G;
    T* slab = reinterpret_cast<T*>(slab_raw) + (size_t)wib * 32 * P;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nblk = (B + RPW - 1) / RPW;
    for (int64_t blk = warp; blk < nblk; blk += nwarps) {
        const int64_t row0 = blk * RPW;
        const int64_t row = row0 + rsub;
        const bool valid = row < B;
        const int nrows = (int)((B - row0) < RPW ? (B - row0) : RPW);
        if (valid)
            for (int dd = g; dd < D; dd += G)
                if (__ldg(mask + dd) != T(0)) y[row * D + dd] = scrub0(x[row * D + dd]);
        T acc = T(0);
        for (int t0 = 0; t0 < Dt; t0 += G) {
            // blocks of this chunk: rows row0..row0+nrows-1, dims t0..min(t0+G,Dt)-1 -- contiguous when G >= Dt
            // (several whole rows) or RPW == 1 (a slice of one row)
            const int tn = (Dt - t0) < G ? (Dt - t0) : G;
            const int nblocks = (G >= Dt) ? nrows * Dt : tn;
            const T* src = params + ((G >= Dt) ? row0 * Dt : row0 * Dt + t0) * (int64_t)P;
            const int nfl = nblocks * P;
            __syncwarp();
            if constexpr (sizeof(T) == 4) {
                // cp.async: all coalesced line fetches of the warp are in flight at once, no register staging; 16-byte
                // copies when the run is 16-byte aligned (32 blocks are a multiple of 128 bytes), 4-byte otherwise
                if ((((uintptr_t)src | (uintptr_t)(nfl * 4)) & 15) == 0) {
                    for (int i = lane * 4; i < nfl; i += 128) {
                        const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(slab + i));
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + i));
                    }
                } else {
                    for (int i = lane; i < nfl; i += 32) {
                        const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(slab + i));
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(src + i));
                    }
                }
                asm volatile("cp.async.commit_group;\n" ::);
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            } else {
                for (int i = lane; i < nfl; i += 32) slab[i] = __ldcs(src + i);
            }
            __syncwarp();
            const int t = t0 + g;
            if (valid && t < Dt) {
                const int dim = __ldg(tidx + t);
                T v = x[row * D + dim];
                if (r_in) v = r_in[dim] * (v - r_lo[dim]) - c.hi;
                const T* pp = slab + (size_t)((G >= Dt) ? rsub * Dt + t : g) * P;
                T uw[KMAX], uh[KMAX], ud[KMAX];
NF_UNROLL
                for (int j = 0; j < KMAX; ++j) {
                    uw[j] = (j < K) ? pp[j] : T(0);
                    uh[j] = (j < K) ? pp[K + j] : T(0);
                    ud[j] = (j < K - 1) ? pp[2 * K + j] : T(0);
                }
                T out, lad;
                rqs_eval<T, KMAX, true>(v, uw, uh, ud, K, inverse != 0, c, out, lad);
                if (r_in) out = (out + c.hi) * r_out[dim] + r_lo[dim];
                y[row * D + dim] = scrub0(out);
                acc += lad;
            }
        }
        acc = group_sum<T, G>(acc);
        if (valid && g == 0) ld[row] = scrub0(acc);
    }
}

template <typename T, int KMAX, bool SK, int G>
__global__ void __launch_bounds__(128)
spline_transform_bwd_kernel(const T* __restrict__ x, const T* __restrict__ params, const T* __restrict__ mask,
                            const int32_t* __restrict__ tidx, const T* __restrict__ gy, const T* __restrict__ gld,
                            T* __restrict__ gx, T* __restrict__ gparams, int64_t B, int D, int Dt, int Krt,
                            int inverse, RqsCfg<T> c, const T* __restrict__ r_in, const T* __restrict__ r_lo,
                            const T* __restrict__ r_out, int compact) {
    const int K = SK ? KMAX : Krt;
    const int P = 3 * K - 1;
    const int PD = compact ? Dt : D;
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, g = lane % G;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nblk = (B + RPW - 1) / RPW;
    for (int64_t blk = warp; blk < nblk; blk += nwarps) {
        const int64_t row = blk * RPW + lane / G;
        if (row >= B) continue;
        const T* xr = x + row * D;
        const T gl = gld[row];
        for (int dd = g; dd < D; dd += G)
            if (__ldg(mask + dd) != T(0)) gx[row * D + dd] = is_finite(xr[dd]) ? gy[row * D + dd] : T(0);
        for (int t = g; t < Dt; t += G) {
            const int dim = __ldg(tidx + t);
            const T xin = xr[dim];
            T v = xin;
            if (r_in) v = r_in[dim] * (v - r_lo[dim]) - c.hi;
            const int64_t pblk = row * PD + (compact ? t : dim);
            const T* pp = params + pblk * P;
            T uw[KMAX], uh[KMAX], ud[KMAX];
            load_row_rt<T, KMAX>(pp, K, uw);
            load_row_rt<T, KMAX>(pp + K, K, uh);
            load_row_rt<T, KMAX>(pp + 2 * K, K - 1, ud);
            T go = gy[row * D + dim];
            {   // layer-level scrub of y (:130): recompute the output to see whether it was replaced by 0
                T out, lad;
                rqs_eval<T, KMAX, true>(v, uw, uh, ud, K, inverse != 0, c, out, lad);
                if (r_in) out = (out + c.hi) * r_out[dim] + r_lo[dim];
                if (!is_finite(out)) go = T(0);
            }
            if (r_in) go *= r_out[dim];
            T guw[KMAX], guh[KMAX], gud[KMAX];
NF_UNROLL
            for (int j = 0; j < KMAX; ++j) { guw[j] = T(0); guh[j] = T(0); gud[j] = T(0); }
            T gv = T(0);
            rqs_eval_bwd<T, KMAX, true>(v, uw, uh, ud, K, inverse != 0, c, go, gl, gv, guw, guh, gud);
            if (r_in) gv *= r_in[dim];
            gx[row * D + dim] = gv;
            T* gp = gparams + pblk * P;
NF_UNROLL
            for (int j = 0; j < KMAX; ++j) if (j < K) { gp[j] = guw[j]; gp[K + j] = guh[j]; }
NF_UNROLL
            for (int j = 0; j < KMAX; ++j) if (j < K - 1) gp[2 * K + j] = gud[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// a5/a6 backward, compact parameter layout: same staging as spline_transform_compact_fwd_kernel.  A warp fetches the
// parameter blocks of its 32 elements with coalesced cp.async into its shared-memory slab, every lane recomputes its
// element and overwrites its own block with the parameter gradients, and the slab goes back to gparams with
// coalesced 128-byte stores (the register-path kernel above issues 3K-1 strided loads and 3K-1 strided stores per
// element and is L1-wavefront bound at 8-15 % of HBM).
// ------------------------------------------------------------------------------------------------
template <typename T, int KMAX, bool SK, int G>
__global__ void __launch_bounds__(128)
spline_transform_compact_bwd_kernel(const T* __restrict__ x, const T* __restrict__ params, const T* __restrict__ mask,
                                    const int32_t* __restrict__ tidx, const T* __restrict__ gy,
                                    const T* __restrict__ gld, T* __restrict__ gx, T* __restrict__ gparams, int64_t B,
                                    int D, int Dt, int Krt, int inverse, RqsCfg<T> c, const T* __restrict__ r_in,
                                    const T* __restrict__ r_lo, const T* __restrict__ r_out) {
    extern __shared__ __align__(16) unsigned char slab_raw[];
    const int K = SK ? KMAX : Krt;
    const int P = 3 * K - 1;
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = lane % G, rsub = lane / G;
    T* slab = reinterpret_cast<T*>(slab_raw) + (size_t)wib * 32 * P;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nblk = (B + RPW - 1) / RPW;
    for (int64_t blk = warp; blk < nblk; blk += nwarps) {
        const int64_t row0 = blk * RPW;
        const int64_t row = row0 + rsub;
        const bool valid = row < B;
        const int nrows = (int)((B - row0) < RPW ? (B - row0) : RPW);
        T gl = T(0);
        if (valid) {
            gl = gld[row];
            for (int dd = g; dd < D; dd += G)
                if (__ldg(mask + dd) != T(0)) gx[row * D + dd] = is_finite(x[row * D + dd]) ? gy[row * D + dd] : T(0);
        }
        for (int t0 = 0; t0 < Dt; t0 += G) {
            const int tn = (Dt - t0) < G ? (Dt - t0) : G;
            const int nblocks = (G >= Dt) ? nrows * Dt : tn;
            const int64_t off = ((G >= Dt) ? row0 * Dt : row0 * Dt + t0) * (int64_t)P;
            const T* src = params + off;
            const int nfl = nblocks * P;
            T* dst = gparams + off;
            const bool vec16 = sizeof(T) == 4 && (((uintptr_t)src | (uintptr_t)dst | (uintptr_t)(nfl * 4)) & 15) == 0;
            __syncwarp();
            if constexpr (sizeof(T) == 4) {
                if (vec16) {
                    for (int i = lane * 4; i < nfl; i += 128) {
                        const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(slab + i));
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + i));
                    }
                } else {
                    for (int i = lane; i < nfl; i += 32) {
                        const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(slab + i));
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(src + i));
                    }
                }
                asm volatile("cp.async.commit_group;\n" ::);
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            } else {
                for (int i = lane; i < nfl; i += 32) slab[i] = __ldcs(src + i);
            }
            __syncwarp();
            const int t = t0 + g;
            if (valid && t < Dt) {
                const int dim = __ldg(tidx + t);
                T v = x[row * D + dim];
                if (r_in) v = r_in[dim] * (v - r_lo[dim]) - c.hi;
                T* pp = slab + (size_t)((G >= Dt) ? rsub * Dt + t : g) * P;
                T uw[KMAX], uh[KMAX], ud[KMAX];
NF_UNROLL
                for (int j = 0; j < KMAX; ++j) {
                    uw[j] = (j < K) ? pp[j] : T(0);
                    uh[j] = (j < K) ? pp[K + j] : T(0);
                    ud[j] = (j < K - 1) ? pp[2 * K + j] : T(0);
                }
                T go = gy[row * D + dim];
                {   // layer-level scrub of y (:130): recompute the output to see whether it was replaced by 0
                    T out, lad;
                    rqs_eval<T, KMAX, true>(v, uw, uh, ud, K, inverse != 0, c, out, lad);
                    if (r_in) out = (out + c.hi) * r_out[dim] + r_lo[dim];
                    if (!is_finite(out)) go = T(0);
                }
                if (r_in) go *= r_out[dim];
                T guw[KMAX], guh[KMAX], gud[KMAX];
NF_UNROLL
                for (int j = 0; j < KMAX; ++j) { guw[j] = T(0); guh[j] = T(0); gud[j] = T(0); }
                T gv = T(0);
                rqs_eval_bwd<T, KMAX, true>(v, uw, uh, ud, K, inverse != 0, c, go, gl, gv, guw, guh, gud);
                if (r_in) gv *= r_in[dim];
                gx[row * D + dim] = gv;
NF_UNROLL
                for (int j = 0; j < KMAX; ++j) if (j < K) { pp[j] = guw[j]; pp[K + j] = guh[j]; }
NF_UNROLL
                for (int j = 0; j < KMAX; ++j) if (j < K - 1) pp[2 * K + j] = gud[j];
            }
            __syncwarp();
            if constexpr (sizeof(T) == 4) {
                if (vec16) {
                    for (int i = lane * 4; i < nfl; i += 128)
                        __stcs(reinterpret_cast<float4*>(dst + i), *reinterpret_cast<const float4*>(slab + i));
                } else {
                    for (int i = lane; i < nfl; i += 32) st_stream(dst + i, slab[i]);
                }
            } else {
                for (int i = lane; i < nfl; i += 32) st_stream(dst + i, slab[i]);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// host-side launchers (explicitly instantiated per translation unit)
// ------------------------------------------------------------------------------------------------
static inline int grid_for(int64_t work_items, int per_block, int blocks_per_sm) {
    int64_t need = cdiv(work_items, per_block);
    int64_t cap = (int64_t)kNumSMs * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

template <typename T, bool GENERIC>
int rqs_unit_fwd_launch(const void* x, const void* w, const void* h, const void* d, void* y, void* ld, int64_t n,
                        int K, int inverse, RqsCfg<T> c, cudaStream_t st) {
    const int grid = grid_for(n, 256, 8 * 4);   // up to 4 waves of 8 CTAs/SM, grid-stride beyond
#define NF_RQS_FWD(KM, SKF)                                                                                    \
    rqs_unit_fwd_kernel<T, KM, SKF><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, (const T*)h, (const T*)d,   \
                                                          (T*)y, (T*)ld, n, K, inverse, c)
    if constexpr (GENERIC) {
        NF_RQS_FWD(32, false);
    } else {
        const bool al = aligned16(w) && aligned16(h);
        if (al && K == 8) NF_RQS_FWD(8, true);
        else if (al && K == 10) NF_RQS_FWD(10, true);
        else if (K <= 8) NF_RQS_FWD(8, false);
        else NF_RQS_FWD(16, false);
    }
#undef NF_RQS_FWD
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

template <typename T, bool GENERIC>
int rqs_unit_bwd_launch(const void* x, const void* w, const void* h, const void* d, const void* gy, const void* gld,
                        void* gx, void* gw, void* gh, void* gd, int64_t n, int K, int inverse, RqsCfg<T> c,
                        cudaStream_t st) {
    const int grid = grid_for(n, 128, 8);
    const size_t smem = (size_t)4 * 32 * (3 * K - 1) * sizeof(T);
#define NF_RQS_BWD(KM, SKF)                                                                                     \
    do {                                                                                                        \
        auto kern = rqs_unit_bwd_slab_kernel<T, KM, SKF>;                                                       \
        if (smem > 48 * 1024) NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, 128, smem, st>>>((const T*)x, (const T*)w, (const T*)h, (const T*)d, (const T*)gy, (const T*)gld, \
                                      (T*)gx, (T*)gw, (T*)gh, (T*)gd, n, K, inverse, c);                        \
    } while (0)
    if constexpr (GENERIC) { NF_RQS_BWD(32, false); }
    else {
        if (K == 8) NF_RQS_BWD(8, true);
        else if (K == 10) NF_RQS_BWD(10, true);
        else if (K < 8) NF_RQS_BWD(8, false);
        else NF_RQS_BWD(16, false);
    }
#undef NF_RQS_BWD
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

// lanes per row: 1 (Dt<=2), 4 (Dt<=16), 32 otherwise
static inline int pick_group3(int n) { return n <= 2 ? 1 : (n <= 16 ? 4 : 32); }

template <typename T>
struct SplineTfArgs {
    const T* x; const T* params; const T* mask; const int32_t* tidx;
    T* y; T* ld;                       // forward outputs
    const T* gy; const T* gld; T* gx; T* gparams;   // backward
    int64_t B; int D, Dt, K, inverse;
    RqsCfg<T> c;
    const T* r_in; const T* r_lo; const T* r_out;
    int compact;                       // params / gparams rows hold only the Dt transformed dims' blocks
};

static inline int pick_group_pow2(int n) { int g = 1; while (g < n && g < 32) g <<= 1; return g; }

template <typename T, bool GENERIC>
int spline_transform_compact_fwd_launch(const SplineTfArgs<T>& a, cudaStream_t st) {
    if constexpr (sizeof(T) == 4 && !GENERIC) {
        SplineStreamArgs sa{(const float*)a.x, (const float*)a.params, (const float*)a.mask, a.tidx, (float*)a.y, (float*)a.ld,
                            (const float*)a.gy, (const float*)a.gld, (float*)a.gx, (float*)a.gparams, a.B, a.D, a.Dt,
                            reinterpret_cast<const RqsCfg<float>&>(a.c), (const float*)a.r_in, (const float*)a.r_lo,
                            (const float*)a.r_out};
        const int rc = spline_stream_launch(sa, a.K, false, a.inverse, st);
        if (rc != NF_ERR_UNSUPPORTED) return rc;
    }
    const int G = pick_group_pow2(a.Dt);
    const int P = 3 * a.K - 1;
    const size_t smem = (size_t)8 * 32 * P * sizeof(T);
    const int grid = grid_for(cdiv(a.B, 32 / G), 8, 8);
#define NF_SC(KM, SKF, GG)                                                                                            \
    do {                                                                                                              \
        auto kern = spline_transform_compact_fwd_kernel<T, KM, SKF, GG>;                                              \
        if (smem > 48 * 1024) NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, 256, smem, st>>>(a.x, a.params, a.mask, a.tidx, a.y, a.ld, a.B, a.D, a.Dt, a.K, a.inverse, a.c,  \
                                      a.r_in, a.r_lo, a.r_out);                                                       \
    } while (0)
#define NF_SC_G(KM, SKF)                                                                                              \
    do {                                                                                                              \
        switch (G) { case 1: NF_SC(KM, SKF, 1); break; case 2: NF_SC(KM, SKF, 2); break; case 4: NF_SC(KM, SKF, 4); break; \
                     case 8: NF_SC(KM, SKF, 8); break; case 16: NF_SC(KM, SKF, 16); break; default: NF_SC(KM, SKF, 32); break; } \
    } while (0)
    if constexpr (GENERIC) { NF_SC_G(32, false); }
    else {
        if (a.K == 8) NF_SC_G(8, true);
        else if (a.K == 10) NF_SC_G(10, true);
        else if (a.K < 8) NF_SC_G(8, false);
        else NF_SC_G(16, false);
    }
#undef NF_SC_G
#undef NF_SC
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

template <typename T, bool GENERIC>
int spline_transform_compact_bwd_launch(const SplineTfArgs<T>& a, cudaStream_t st) {
    if constexpr (sizeof(T) == 4 && !GENERIC) {
        SplineStreamArgs sa{(const float*)a.x, (const float*)a.params, (const float*)a.mask, a.tidx, (float*)a.y, (float*)a.ld,
                            (const float*)a.gy, (const float*)a.gld, (float*)a.gx, (float*)a.gparams, a.B, a.D, a.Dt,
                            reinterpret_cast<const RqsCfg<float>&>(a.c), (const float*)a.r_in, (const float*)a.r_lo,
                            (const float*)a.r_out};
        const int rc = spline_stream_launch(sa, a.K, true, a.inverse, st);
        if (rc != NF_ERR_UNSUPPORTED) return rc;
    }
    const int G = pick_group_pow2(a.Dt);
    const int P = 3 * a.K - 1;
    const size_t smem = (size_t)4 * 32 * P * sizeof(T);
    const int grid = grid_for(cdiv(a.B, 32 / G), 4, 8);
#define NF_SB(KM, SKF, GG)                                                                                            \
    do {                                                                                                              \
        auto kern = spline_transform_compact_bwd_kernel<T, KM, SKF, GG>;                                              \
        if (smem > 48 * 1024) NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, 128, smem, st>>>(a.x, a.params, a.mask, a.tidx, a.gy, a.gld, a.gx, a.gparams, a.B, a.D, a.Dt, a.K, \
                                      a.inverse, a.c, a.r_in, a.r_lo, a.r_out);                                       \
    } while (0)
#define NF_SB_G(KM, SKF)                                                                                              \
    do {                                                                                                              \
        switch (G) { case 1: NF_SB(KM, SKF, 1); break; case 2: NF_SB(KM, SKF, 2); break; case 4: NF_SB(KM, SKF, 4); break; \
                     case 8: NF_SB(KM, SKF, 8); break; case 16: NF_SB(KM, SKF, 16); break; default: NF_SB(KM, SKF, 32); break; } \
    } while (0)
    if constexpr (GENERIC) { NF_SB_G(32, false); }
    else {
        if (a.K == 8) NF_SB_G(8, true);
        else if (a.K == 10) NF_SB_G(10, true);
        else if (a.K < 8) NF_SB_G(8, false);
        else NF_SB_G(16, false);
    }
#undef NF_SB_G
#undef NF_SB
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

template <typename T, bool BWD, bool GENERIC>
int spline_transform_launch(const SplineTfArgs<T>& a, cudaStream_t st) {
    if constexpr (!BWD) {
        if (a.compact && a.Dt > 0) return spline_transform_compact_fwd_launch<T, GENERIC>(a, st);
    } else {
        if (a.compact && a.Dt > 0) return spline_transform_compact_bwd_launch<T, GENERIC>(a, st);
    }
    const int G = pick_group3(a.Dt);
    const int grid = grid_for(cdiv(a.B, 32 / G), 4, 64);
#define NF_ST(KM, GG)                                                                                              \
    do {                                                                                                           \
        if constexpr (BWD)                                                                                         \
            spline_transform_bwd_kernel<T, KM, false, GG><<<grid, 128, 0, st>>>(a.x, a.params, a.mask, a.tidx, a.gy, \
                a.gld, a.gx, a.gparams, a.B, a.D, a.Dt, a.K, a.inverse, a.c, a.r_in, a.r_lo, a.r_out, a.compact);    \
        else                                                                                                       \
            spline_transform_fwd_kernel<T, KM, false, GG><<<grid, 128, 0, st>>>(a.x, a.params, a.mask, a.tidx, a.y,  \
                a.ld, a.B, a.D, a.Dt, a.K, a.inverse, a.c, a.r_in, a.r_lo, a.r_out, a.compact);                      \
    } while (0)
#define NF_ST_G(KM)                                        \
    do {                                                   \
        if (G == 1) NF_ST(KM, 1);                          \
        else if (G == 4) NF_ST(KM, 4);                     \
        else NF_ST(KM, 32);                                \
    } while (0)
    if constexpr (GENERIC) { NF_ST_G(32); }
    else { if (a.K <= 8) NF_ST_G(8); else NF_ST_G(16); }
#undef NF_ST_G
#undef NF_ST
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

}  // namespace nf
