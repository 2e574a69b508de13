// ar_blocked.cu -- sequential directions of the affine autoregressive flows (MAF.forward / IAF.inverse,
// masked_autoregressive_flow.py:46-78 / inverse_autoregressive_flow.py:65-103) as a *blocked* triangular evaluation.
//
// The reference re-evaluates the whole MADE D times (D x 4 dense GEMMs).  Hidden units sorted by degree make every
// masked weight block-lower-triangular, so the D dependent steps are grouped into blocks of `gb` consecutive degrees:
//   * contributions of all PREVIOUS blocks to the block's pre-activations are plain dense products over the whole
//     batch -> nf_linear_tc (tcgen05, 3xTF32, TMA) on column slices of the activation buffers:
//         pre_l[:, blk] = act_{l-1}[:, :u0] * W_l[blk, :u0]^T        (l = 1..3),   preo[:, dims] = act3[:, :u0] * W3[dims, :u0]^T
//   * the IN-BLOCK part (gb dependent steps over <= ~72 units per layer) runs in ar_block_warp_kernel (below: a warp
//     owns 32 rows for the whole block, no CTA barrier); ar_block_kernel is the first version (32 rows per CTA, four
//     __syncthreads per degree), kept as the fallback for blocks whose weights do not fit in shared memory.
// Every hidden unit is still evaluated exactly once (total work = one masked MADE pass), ~75 % of it on the tensor pipe.
#include "nf_common.cuh"

extern "C" int nf_linear_tc(const void*, const void*, const void*, const void*, void*, int64_t, int64_t, int64_t, int64_t,
                            int64_t, int64_t, int, const int32_t*, nf_stream_t);
extern "C" int nf_ar_finish_forward(const void*, const void*, const void*, void*, void*, int64_t, int, int, int, nf_stream_t);

namespace nf {

int gemm_tc2_launch(const void* x, const void* w_hi, const void* w_lo, const void* bias, void* y, int64_t M, int64_t N, int64_t K,
                    int64_t ldx, int64_t ldw, int64_t ldy, int relu, const int32_t* k_begin, const int32_t* k_extent,
                    cudaStream_t st, int accumulate);

// Y[M,N] (+)= X[M,K] W[N,K]^T for K <= 128 through the persistent direct-epilogue kernel of gemm_tc2.cu
static int linear_tc_push(const float* x, const float* w_hi, const float* w_lo, float* y, int64_t M, int64_t N, int64_t K,
                          int64_t ldx, int64_t ldw, int64_t ldy, int accumulate, cudaStream_t st) {
    if (!aligned16(x) || !aligned16(w_hi) || !aligned16(w_lo) || (ldx % 4) != 0 || (ldw % 4) != 0) return NF_ERR_UNSUPPORTED;
    const int rc = gemm_tc2_launch(x, w_hi, w_lo, nullptr, y, M, N, K, ldx, ldw, ldy, 0, nullptr, nullptr, st, accumulate);
    if (rc != NF_OK) return rc;
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

constexpr int kBlkRows = 32;
constexpr int kBlkWarps = 4;
constexpr int kBlkPad = 33;          // [unit][row] tiles padded to 33 rows: conflict-free transposed fills and lane reads

// dot over v in [0,n) of w[v] * act[v][lane]; w warp-uniform (global, read-only path), act [*][kBlkPad] in shared memory
__device__ __forceinline__ float dot_tile(const float* __restrict__ w, const float* __restrict__ act, int n, int lane) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int v = 0;
    for (; v + 4 <= n; v += 4) {
        s0 = fmaf(__ldg(w + v + 0), act[(v + 0) * kBlkPad + lane], s0);
        s1 = fmaf(__ldg(w + v + 1), act[(v + 1) * kBlkPad + lane], s1);
        s2 = fmaf(__ldg(w + v + 2), act[(v + 2) * kBlkPad + lane], s2);
        s3 = fmaf(__ldg(w + v + 3), act[(v + 3) * kBlkPad + lane], s3);
    }
    for (; v < n; ++v) s0 = fmaf(__ldg(w + v), act[v * kBlkPad + lane], s0);
    return (s0 + s1) + (s2 + s3);
}

// coalesced [rows x n] slice of a row-major [B, ld] array -> transposed shared tile [n][kBlkPad] (zeros when src == nullptr)
__device__ __forceinline__ void fill_tile(float* tile, const float* __restrict__ src, int64_t r0, int nrow, int ld, int c0, int n) {
    for (int i = threadIdx.x; i < kBlkRows * n; i += blockDim.x) {
        const int r = i / n, c = i - r * n;
        tile[c * kBlkPad + r] = (src && r < nrow) ? src[(r0 + r) * ld + c0 + c] : 0.f;
    }
}
__device__ __forceinline__ void drain_tile(const float* tile, float* __restrict__ dst, int64_t r0, int nrow, int ld, int c0, int n) {
    for (int i = threadIdx.x; i < nrow * n; i += blockDim.x) {
        const int r = i / n, c = i - r * n;
        dst[(r0 + r) * ld + c0 + c] = tile[c * kBlkPad + r];
    }
}

// dims [g0,g1), hidden units [u0,u1) (= degrees g0..g1-1).  pre*/preo: partial sums from previous blocks (nullptr for
// the first block).  xcur: outputs so far, unscrubbed (NaN/Inf must keep poisoning later dims like the dense reference).
__global__ void __launch_bounds__(kBlkRows * kBlkWarps)
ar_block_kernel(const float* __restrict__ vin, float* __restrict__ xcur, const float* __restrict__ pre1,
                const float* __restrict__ pre2, const float* __restrict__ pre3, const float* __restrict__ preo,
                float* __restrict__ act1, float* __restrict__ act2, float* __restrict__ act3,
                const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
                const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                const float* __restrict__ w3, const float* __restrict__ b3, const int32_t* __restrict__ gstart,
                float* __restrict__ ldacc, int* __restrict__ bad, int64_t B, int D, int H, int g0, int g1, int u0, int u1,
                int mode) {
    extern __shared__ __align__(16) float sm[];
    const int nd = g1 - g0, nu = u1 - u0;
    float* sx = sm;                              // [nd][pad] inputs of the block's dims, replaced by outputs
    float* a1 = sx + nd * kBlkPad;               // [nu][pad] layer-1 pre-activation partials -> activations
    float* a2 = a1 + nu * kBlkPad;
    float* a3 = a2 + nu * kBlkPad;
    float* sld = a3 + nu * kBlkPad;              // [32] running log-det
    int* sbad = reinterpret_cast<int*>(sld + kBlkRows);   // [32] poison flag
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * kBlkRows;
    const int nrow = (int)((B - r0) < kBlkRows ? (B - r0) : kBlkRows);
    const bool first = (g0 == 0);

    fill_tile(sx, vin, r0, nrow, D, g0, nd);
    fill_tile(a1, pre1, r0, nrow, H, u0, nu);
    fill_tile(a2, pre2, r0, nrow, H, u0, nu);
    fill_tile(a3, pre3, r0, nrow, H, u0, nu);
    if (threadIdx.x < kBlkRows) {
        const bool ok = threadIdx.x < nrow;
        sld[threadIdx.x] = (!first && ok) ? ldacc[r0 + threadIdx.x] : 0.f;
        sbad[threadIdx.x] = (!first && ok) ? bad[r0 + threadIdx.x] : 0;
    }
    __syncthreads();

    for (int g = g0; g < g1; ++g) {
        const int ub0 = gstart[g] - u0, ub1 = gstart[g + 1] - u0;     // in-block units of degree g
        // (A) parameters of dim g: previous blocks (preo) + in-block layer-3 units of degree < g
        if (warp == 0) {
            float mu = __ldg(b3 + g) + dot_tile(w3 + (size_t)g * H + u0, a3, ub0, lane);
            float al = __ldg(b3 + D + g) + dot_tile(w3 + (size_t)(D + g) * H + u0, a3, ub0, lane);
            if (preo && lane < nrow) { mu += preo[(r0 + lane) * 2 * D + 2 * g]; al += preo[(r0 + lane) * 2 * D + 2 * g + 1]; }
            float o, t;
            affine_ar_elem<float>(mode, sx[(g - g0) * kBlkPad + lane], mu, al, o, t);
            if (sbad[lane]) { o = __int_as_float(0x7fc00000); t = o; }
            if (!is_finite(o)) sbad[lane] = 1;      // 0*NaN of the dense reference poisons every later dim
            sx[(g - g0) * kBlkPad + lane] = o;
            sld[lane] += t;
        }
        __syncthreads();
        if (g == D - 1) break;
        // (B) hidden units of degree g, layer by layer
        for (int u = ub0 + warp; u < ub1; u += kBlkWarps) {
            float s = a1[u * kBlkPad + lane] + __ldg(b0 + u0 + u);
            s += dot_tile(w0 + (size_t)(u0 + u) * D + g0, sx, g - g0 + 1, lane);
            a1[u * kBlkPad + lane] = relu_nan(s);
        }
        __syncthreads();
        for (int u = ub0 + warp; u < ub1; u += kBlkWarps) {
            const float s = a2[u * kBlkPad + lane] + __ldg(b1 + u0 + u) + dot_tile(w1 + (size_t)(u0 + u) * H + u0, a1, ub1, lane);
            a2[u * kBlkPad + lane] = relu_nan(s);
        }
        __syncthreads();
        for (int u = ub0 + warp; u < ub1; u += kBlkWarps) {
            const float s = a3[u * kBlkPad + lane] + __ldg(b2 + u0 + u) + dot_tile(w2 + (size_t)(u0 + u) * H + u0, a2, ub1, lane);
            a3[u * kBlkPad + lane] = relu_nan(s);
        }
        __syncthreads();
    }
    drain_tile(sx, xcur, r0, nrow, D, g0, nd);
    if (g1 < D) {
        drain_tile(a1, act1, r0, nrow, H, u0, nu);
        drain_tile(a2, act2, r0, nrow, H, u0, nu);
        drain_tile(a3, act3, r0, nrow, H, u0, nu);
    }
    if (threadIdx.x < nrow) { ldacc[r0 + threadIdx.x] = sld[threadIdx.x]; bad[r0 + threadIdx.x] = sbad[threadIdx.x]; }
}


// ------------------------------------------------------------------------------------------------------------------
// Second version of the in-block kernel.  Every tile access above is [unit][lane]: a lane only ever touches its own
// row, so nothing in the step loop needs a CTA barrier -- the first version still split the units of a degree across
// four warps and paid four __syncthreads per degree (plus two loads per FMA in dot_tile): 0.83 ms per block at
// 262 144 rows.  Here a warp owns 32 rows for the whole block (no barrier after the weight staging), the in-block
// weights sit in shared memory TRANSPOSED ([v][unit]) so that the weights of four consecutive units are one
// broadcast LDS.128, and up to three 4-unit chunks share every activation load: 1 LDS.32 + 3 LDS.128 per 12 FMAs.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kWarpChunks = 3;       // 4-unit chunks evaluated together (12 independent FMA chains per lane)

// [32 rows x n] slice of a row-major [B, ld] array -> transposed tile [n][kBlkPad] of the calling warp.  4-byte cp.async:
// every element of the slice is in flight at once (the warp is alone on its scheduler: a register-staged loop would
// expose the full global-memory latency per element); rows are walked in the outer loop, so no integer division.
__device__ __forceinline__ void warp_fill_tile(float* tile, const float* __restrict__ src, int64_t r0, int nrow, int ld, int c0,
                                               int n, int lane) {
    if (!src) {
        for (int c = 0; c < n; ++c) tile[c * kBlkPad + lane] = 0.f;
        return;
    }
    // lanes along the columns (coalesced 128-byte row segments), rows in the unrolled inner loop: one LDGSTS and one
    // pointer bump per element (the first version recomputed both addresses per element: 25 instructions per copy)
    for (int c = lane; c < n; c += 32) {
        const float* gp = src + r0 * (int64_t)ld + c0 + c;
        const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(tile + c * kBlkPad));
        if (nrow == kBlkRows) {
#pragma unroll 8
            for (int r = 0; r < kBlkRows; ++r, gp += ld)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa + 4u * r), "l"(gp));
        } else {
            for (int r = 0; r < kBlkRows; ++r, gp += ld) {
                if (r < nrow) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa + 4u * r), "l"(gp));
                else tile[c * kBlkPad + r] = 0.f;
            }
        }
    }
}
__device__ __forceinline__ void warp_drain_tile(const float* tile, float* __restrict__ dst, int64_t r0, int nrow, int ld, int c0,
                                                int n, int lane) {
    for (int c = lane; c < n; c += 32) {
        float* gp = dst + r0 * (int64_t)ld + c0 + c;
        const float* sp = tile + c * kBlkPad;
        if (nrow == kBlkRows) {
#pragma unroll 8
            for (int r = 0; r < kBlkRows; ++r, gp += ld) *gp = sp[r];
        } else {
            for (int r = 0; r < nrow; ++r, gp += ld) *gp = sp[r];
        }
    }
}

template <int NCH, int NUP>
__device__ __forceinline__ void warp_units_chunk(const float* __restrict__ wt, const float* __restrict__ bias,
                                                 const float* __restrict__ in, int nv, float* __restrict__ out, int cu, int ub0,
                                                 int ub1, int lane) {
    // accumulators as register pairs: packed fp32 FMAs (FFMA2), two units per instruction -- the weights of a pair are
    // adjacent in the broadcast LDS.128, the lane's activation is duplicated into a pair; same operations and order
    float2 acc2[NCH][2];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { acc2[c][0] = make_float2(0.f, 0.f); acc2[c][1] = make_float2(0.f, 0.f); }
    // compile-time row pitches (NUP, kBlkPad): the unrolled body addresses everything with immediate offsets
    const float* ap = in + lane;
    const float* wp = wt + cu;
    // groups of four inputs, software-pipelined by hand: the operands of group i+1 are requested before the FMAs of group
    // i issue.  A warp shares its scheduler with at most one other warp here (shared memory bounds the CTA at 5-6 warps),
    // so nothing else covers the ~30-cycle LDS latency: the straight loop sat on the short scoreboard for 40 % of its
    // samples (profiles/r02l_c3_ar_block_warp_ncu.txt).  Same FMAs in the same order.
    auto load = [&](int grp, float (&a)[4], float4 (&w)[4][NCH]) {
        const float* a_ = ap + grp * 4 * kBlkPad;
        const float* w_ = wp + grp * 4 * NUP;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[k] = a_[k * kBlkPad];
#pragma unroll
            for (int c = 0; c < NCH; ++c) w[k][c] = *reinterpret_cast<const float4*>(w_ + k * NUP + 4 * c);   // rows padded by 4*kWarpChunks floats
        }
    };
    auto fma = [&](const float (&a)[4], const float4 (&w)[4][NCH]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 aa = make_float2(a[k], a[k]);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                acc2[c][0] = __ffma2_rn(make_float2(w[k][c].x, w[k][c].y), aa, acc2[c][0]);
                acc2[c][1] = __ffma2_rn(make_float2(w[k][c].z, w[k][c].w), aa, acc2[c][1]);
            }
        }
    };
    const int ngrp = nv >> 2;
    if (ngrp > 0) {
        float aA[4], aB[4];
        float4 wA[4][NCH], wB[4][NCH];
        load(0, aA, wA);
        int g = 0;
        for (; g + 2 <= ngrp; g += 2) {
            load(g + 1, aB, wB);
            fma(aA, wA);
            if (g + 2 < ngrp) load(g + 2, aA, wA);
            fma(aB, wB);
        }
        if (g < ngrp) fma(aA, wA);
    }
    ap += ngrp * 4 * kBlkPad; wp += ngrp * 4 * NUP;
    for (int v = ngrp * 4; v < nv; ++v, ap += kBlkPad, wp += NUP) {
        const float a = *ap;
        const float2 aa = make_float2(a, a);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(wp + 4 * c);
            acc2[c][0] = __ffma2_rn(make_float2(w.x, w.y), aa, acc2[c][0]);
            acc2[c][1] = __ffma2_rn(make_float2(w.z, w.w), aa, acc2[c][1]);
        }
    }
    float acc[NCH][4];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { acc[c][0] = acc2[c][0].x; acc[c][1] = acc2[c][0].y; acc[c][2] = acc2[c][1].x; acc[c][3] = acc2[c][1].y; }
    float* op = out + cu * kBlkPad + lane;
    const float* bp = bias + cu;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int u = cu + 4 * c + j;
            if (u >= ub0 && u < ub1) op[(4 * c + j) * kBlkPad] = relu_nan(op[(4 * c + j) * kBlkPad] + bp[4 * c + j] + acc[c][j]);
        }
}

template <int NUP>
__device__ __forceinline__ void warp_units(const float* __restrict__ wt, const float* __restrict__ bias,
                                           const float* __restrict__ in, int nv, float* __restrict__ out, int ub0, int ub1,
                                           int lane) {
    int cu = ub0 & ~3;
    while (cu < ub1) {
        const int left = (ub1 - cu + 3) >> 2;          // 4-unit chunks still to do (warp-uniform)
        if (left >= 3) { warp_units_chunk<3, NUP>(wt, bias, in, nv, out, cu, ub0, ub1, lane); cu += 12; }
        else if (left == 2) { warp_units_chunk<2, NUP>(wt, bias, in, nv, out, cu, ub0, ub1, lane); cu += 8; }
        else { warp_units_chunk<1, NUP>(wt, bias, in, nv, out, cu, ub0, ub1, lane); cu += 4; }
    }
}

constexpr int kBlkMaxDeg = 8;        // degrees per block the warp kernel takes (compile-time pitch of its W3 tile)

// Layer 3 of the in-block step: the units [ub0, ub1) of the current degree, evaluated like warp_units_chunk (inputs: the
// layer-2 tile), but their values never go to shared memory.  Nothing in the block reads a layer-3 activation except the
// output layer, so each finished unit is PUSHED straight into the (mu, alpha) partial sums of the block's dims, held in
// registers (par[d]; the output weights of dims the unit does not feed are exact zeros of the folded mask), and written
// to the global activation buffer for the later blocks' pull / push products.  The partial pre-activation (previous
// blocks' pull product) comes straight from global memory -- requested before the contraction loop, used after it.
// Without the layer-3 tile a warp's tile is a third smaller: eight warps per SM instead of five or six.
template <int NCH, int NUP>
__device__ __forceinline__ void warp_units_l3_chunk(const float* __restrict__ wt, const float* __restrict__ bias,
                                                    const float* __restrict__ in, int nv, const float* __restrict__ pre_row,
                                                    float* __restrict__ act_row, const float* __restrict__ w3t,
                                                    float2 (&par)[kBlkMaxDeg], int cu, int ub0, int ub1, int lane) {
    float4 pre[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
        pre[c] = pre_row ? *reinterpret_cast<const float4*>(pre_row + cu + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float2 acc2[NCH][2];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { acc2[c][0] = make_float2(0.f, 0.f); acc2[c][1] = make_float2(0.f, 0.f); }
    const float* ap = in + lane;
    const float* wp = wt + cu;
    auto load = [&](int grp, float (&a)[4], float4 (&w)[4][NCH]) {
        const float* a_ = ap + grp * 4 * kBlkPad;
        const float* w_ = wp + grp * 4 * NUP;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[k] = a_[k * kBlkPad];
#pragma unroll
            for (int c = 0; c < NCH; ++c) w[k][c] = *reinterpret_cast<const float4*>(w_ + k * NUP + 4 * c);
        }
    };
    auto fma = [&](const float (&a)[4], const float4 (&w)[4][NCH]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 aa = make_float2(a[k], a[k]);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                acc2[c][0] = __ffma2_rn(make_float2(w[k][c].x, w[k][c].y), aa, acc2[c][0]);
                acc2[c][1] = __ffma2_rn(make_float2(w[k][c].z, w[k][c].w), aa, acc2[c][1]);
            }
        }
    };
    const int ngrp = nv >> 2;
    if (ngrp > 0) {
        float aA[4], aB[4];
        float4 wA[4][NCH], wB[4][NCH];
        load(0, aA, wA);
        int g = 0;
        for (; g + 2 <= ngrp; g += 2) {
            load(g + 1, aB, wB);
            fma(aA, wA);
            if (g + 2 < ngrp) load(g + 2, aA, wA);
            fma(aB, wB);
        }
        if (g < ngrp) fma(aA, wA);
    }
    ap += ngrp * 4 * kBlkPad; wp += ngrp * 4 * NUP;
    for (int v = ngrp * 4; v < nv; ++v, ap += kBlkPad, wp += NUP) {
        const float a = *ap;
        const float2 aa = make_float2(a, a);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(wp + 4 * c);
            acc2[c][0] = __ffma2_rn(make_float2(w.x, w.y), aa, acc2[c][0]);
            acc2[c][1] = __ffma2_rn(make_float2(w.z, w.w), aa, acc2[c][1]);
        }
    }
    const float* bp = bias + cu;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const float accv[4] = {acc2[c][0].x, acc2[c][0].y, acc2[c][1].x, acc2[c][1].y};
        const float prev[4] = {pre[c].x, pre[c].y, pre[c].z, pre[c].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int u = cu + 4 * c + j;
            if (u >= ub0 && u < ub1) {                               // warp-uniform
                const float h = relu_nan(prev[j] + bp[4 * c + j] + accv[j]);
                if (act_row) act_row[u] = h;
                const float2 hh = make_float2(h, h);
                const float4* wr = reinterpret_cast<const float4*>(w3t + u * (2 * kBlkMaxDeg));
#pragma unroll
                for (int q = 0; q < kBlkMaxDeg / 2; ++q) {
                    const float4 w = wr[q];                           // (mu, alpha) weights of dims 2q, 2q + 1
                    par[2 * q] = __ffma2_rn(make_float2(w.x, w.y), hh, par[2 * q]);
                    par[2 * q + 1] = __ffma2_rn(make_float2(w.z, w.w), hh, par[2 * q + 1]);
                }
            }
        }
    }
}

template <int NUP>
__device__ __forceinline__ void warp_units_l3(const float* __restrict__ wt, const float* __restrict__ bias,
                                              const float* __restrict__ in, int nv, const float* __restrict__ pre_row,
                                              float* __restrict__ act_row, const float* __restrict__ w3t,
                                              float2 (&par)[kBlkMaxDeg], int ub0, int ub1, int lane) {
    int cu = ub0 & ~3;
    while (cu < ub1) {
        const int left = (ub1 - cu + 3) >> 2;
        if (left >= 3) { warp_units_l3_chunk<3, NUP>(wt, bias, in, nv, pre_row, act_row, w3t, par, cu, ub0, ub1, lane); cu += 12; }
        else if (left == 2) { warp_units_l3_chunk<2, NUP>(wt, bias, in, nv, pre_row, act_row, w3t, par, cu, ub0, ub1, lane); cu += 8; }
        else { warp_units_l3_chunk<1, NUP>(wt, bias, in, nv, pre_row, act_row, w3t, par, cu, ub0, ub1, lane); cu += 4; }
    }
}

// PERSISTENT: the block's weights are staged once per CTA (one CTA per SM) and every warp then walks over row tiles on
// its own -- no CTA barrier after the staging, so one warp's tile fill / drain overlaps the other warps' step loops.  (As
// one CTA per 6 row tiles, each of the ~9 CTAs an SM ran in turn re-staged ~58 KB of weights and sat alone on the SM
// while they and the tiles arrived: a third of the kernel's stall samples, profiles/r02l_c3_ar_block_warp_ncu.txt.)
// preo: [B, 2D] with (mu, alpha) of a dim adjacent -- the layout the push GEMMs of nf_ar_blocked_forward accumulate into.
template <int NUP>
__global__ void __launch_bounds__(256)
ar_block_warp_kernel(const float* __restrict__ vin, float* __restrict__ xcur, const float* __restrict__ pre1,
                     const float* __restrict__ pre2, const float* __restrict__ pre3, const float* __restrict__ preo,
                     float* __restrict__ act1, float* __restrict__ act2, float* __restrict__ act3,
                     const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
                     const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                     const float* __restrict__ w3, const float* __restrict__ b3, const int32_t* __restrict__ gstart,
                     float* __restrict__ ldacc, int* __restrict__ bad, int64_t B, int D, int H, int g0, int g1, int u0, int u1,
                     int mode) {
    extern __shared__ __align__(16) float sm[];
    const int nd = g1 - g0, nu = u1 - u0;
    constexpr int nup = NUP;                                     // padded row length of the transposed weight tiles
    constexpr int p3 = 2 * kBlkMaxDeg;                           // row pitch of the W3 tile
    const int nwarps = blockDim.x >> 5;
    // shared weights: W0t [nd][nup], W1t / W2t [nu][nup], W3t [nu][p3] (mu, alpha interleaved per dim), biases [3][nu],
    // output biases [p3] (mu, alpha interleaved), in-block unit boundaries of the block's degrees [kBlkMaxDeg + 2]
    float* W0t = sm;
    float* W1t = W0t + nd * nup;
    float* W2t = W1t + nu * nup;
    float* W3t = W2t + nu * nup;
    float* bs = W3t + nu * p3;
    float* b3s = bs + ((3 * nu + 3) & ~3);
    int* gsm = reinterpret_cast<int*>(b3s + p3);
    float* tiles = b3s + p3 + 16;
    const int tile_floats = (nd + 2 * nu) * kBlkPad;             // inputs/outputs of the dims, layer-1 and layer-2 units
    for (int i = threadIdx.x; i < (nd + 2 * nu) * nup + nu * p3; i += blockDim.x) W0t[i] = 0.f;      // padding columns must be finite
    __syncthreads();
    {
        const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5, nw_ = blockDim.x >> 5;
#define NF_CPA4(dst, srcp) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(dst))), "l"(srcp))
        for (int u = warp_; u < nu; u += nw_) {                   // one unit (weight row) per warp, lanes along v; all copies in flight
            const float* r0p = w0 + (size_t)(u0 + u) * D + g0;
            for (int d = lane_; d < nd; d += 32) NF_CPA4(W0t + d * nup + u, r0p + d);
            const float* r1p = w1 + (size_t)(u0 + u) * H + u0;
            const float* r2p = w2 + (size_t)(u0 + u) * H + u0;
            for (int v = lane_; v < nu; v += 32) { NF_CPA4(W1t + v * nup + u, r1p + v); NF_CPA4(W2t + v * nup + u, r2p + v); }
        }
        for (int d = warp_; d < nd; d += nw_) {
            const float* rm = w3 + (size_t)(g0 + d) * H + u0;
            const float* ra = w3 + (size_t)(D + g0 + d) * H + u0;
            for (int v = lane_; v < nu; v += 32) { NF_CPA4(W3t + v * p3 + 2 * d, rm + v); NF_CPA4(W3t + v * p3 + 2 * d + 1, ra + v); }
        }
#undef NF_CPA4
    }
    for (int i = threadIdx.x; i < nu; i += blockDim.x) { bs[i] = b0[u0 + i]; bs[nu + i] = b1[u0 + i]; bs[2 * nu + i] = b2[u0 + i]; }
    if (threadIdx.x < nd) { b3s[2 * threadIdx.x] = b3[g0 + threadIdx.x]; b3s[2 * threadIdx.x + 1] = b3[D + g0 + threadIdx.x]; }
    if (threadIdx.x <= nd + 1) gsm[threadIdx.x] = (g0 + (int)threadIdx.x <= D ? gstart[g0 + threadIdx.x] : gstart[D]) - u0;
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();                                 // the weights (all warps' copies) have landed; no CTA barrier below

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool first = (g0 == 0);
    float* sx = tiles + (size_t)warp * tile_floats;
    float* a1 = sx + nd * kBlkPad;
    float* a2 = a1 + nu * kBlkPad;
    const int64_t ntiles = (B + kBlkRows - 1) / kBlkRows;
    for (int64_t tile = (int64_t)blockIdx.x * nwarps + warp; tile < ntiles; tile += (int64_t)gridDim.x * nwarps) {
        const int64_t r0 = tile * kBlkRows;
        const int nrow = (int)((B - r0) < kBlkRows ? (B - r0) : kBlkRows);
        warp_fill_tile(sx, vin, r0, nrow, D, g0, nd, lane);
        warp_fill_tile(a1, pre1, r0, nrow, H, u0, nu, lane);
        warp_fill_tile(a2, pre2, r0, nrow, H, u0, nu, lane);
        asm volatile("cp.async.commit_group;\n" ::);
        const bool ok = lane < nrow;
        float ld = (!first && ok) ? ldacc[r0 + lane] : 0.f;
        int poisoned = (!first && ok) ? bad[r0 + lane] : 0;
        // previous blocks' contributions to the parameters of this block's dims, up front: inside the step loop every
        // global load would sit on the critical path of a warp that is nearly alone on its scheduler
        const float4* prow = (preo && ok) ? reinterpret_cast<const float4*>(preo + (r0 + lane) * 2 * D + 2 * g0) : nullptr;
        float2 par[kBlkMaxDeg];                      // (mu, alpha) partial sums of the block's dims
#pragma unroll
        for (int q = 0; q < kBlkMaxDeg / 2; ++q) {
            const float4 pv = (prow && 2 * q < nd) ? prow[q] : make_float4(0.f, 0.f, 0.f, 0.f);      // nd is a multiple of 4
            par[2 * q] = make_float2(pv.x, pv.y); par[2 * q + 1] = make_float2(pv.z, pv.w);
        }
        const float* pre3_row = (pre3 && ok) ? pre3 + (r0 + lane) * (int64_t)H + u0 : nullptr;
        float* act3_row = (g1 < D && ok) ? act3 + (r0 + lane) * (int64_t)H + u0 : nullptr;
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncwarp();

        for (int g = g0; g < g1; ++g) {
            const int gi = g - g0;
            const int ub0 = gsm[gi], ub1 = gsm[gi + 1];                       // in-block units of degree g
            // (A) parameters of dim g: bias + previous blocks (preo) + the in-block layer-3 units of degree < g, all
            //     already folded into par[gi] by the pushes of the earlier steps
            {
                float2 pq = par[0];
#pragma unroll
                for (int d = 1; d < kBlkMaxDeg; ++d) if (gi == d) pq = par[d];
                const float2 bq = *reinterpret_cast<const float2*>(b3s + 2 * gi);
                const float mu = bq.x + pq.x, al = bq.y + pq.y;
                float o, t;
                affine_ar_elem<float>(mode, sx[gi * kBlkPad + lane], mu, al, o, t);
                if (poisoned) { o = __int_as_float(0x7fc00000); t = o; }
                if (!is_finite(o)) poisoned = 1;        // 0*NaN of the dense reference poisons every later dim
                sx[gi * kBlkPad + lane] = o;
                ld += t;
            }
            if (g == D - 1) break;
            // (B) hidden units of degree g, layer by layer (all lane-private: no synchronisation)
            warp_units<NUP>(W0t, bs, sx, gi + 1, a1, ub0, ub1, lane);
            warp_units<NUP>(W1t, bs + nu, a1, ub1, a2, ub0, ub1, lane);
            warp_units_l3<NUP>(W2t, bs + 2 * nu, a2, ub1, pre3_row, act3_row, W3t, par, ub0, ub1, lane);
        }
        __syncwarp();
        warp_drain_tile(sx, xcur, r0, nrow, D, g0, nd, lane);
        if (g1 < D) {
            warp_drain_tile(a1, act1, r0, nrow, H, u0, nu, lane);
            warp_drain_tile(a2, act2, r0, nrow, H, u0, nu, lane);
        }
        if (ok) { ldacc[r0 + lane] = ld; bad[r0 + lane] = poisoned; }
        __syncwarp();                                // the drain's reads of the tile are done before the next fill lands
    }
}

int g_ar_block_variant = 1;          // nf_set_option(3, v): 0 = first in-block kernel (CTA barriers), 1 = warp-private tiles

}  // namespace nf

using namespace nf;
#define NF_REQ(p) do { if ((p) == nullptr) return NF_ERR_NULL; } while (0)

extern "C" int64_t nf_ar_blocked_workspace_floats(int64_t B, int D, int H) {
    // act1..3, pre1..3: 6*B*H; preo: 2*B*D; xcur: B*D; ldacc: B; bad: B (ints)
    return 6 * B * H + 3 * B * D + 2 * B;
}

extern "C" int nf_ar_blocked_forward(const void* v, const void* const* w, const void* const* w_hi, const void* const* w_lo,
                                     const void* const* b, const int32_t* gstart_dev, const int32_t* gstart_host,
                                     void* workspace, void* out, void* ld, int64_t B, int D, int H, int mode,
                                     int block_degrees, nf_stream_t stream) {
    if (B < 0 || D < 1 || H < 1 || block_degrees < 1) return NF_ERR_BAD_SHAPE;
    if (mode != NF_AR_MAF_FORWARD && mode != NF_AR_IAF_INVERSE) return NF_ERR_UNSUPPORTED;
    if (B == 0) return NF_OK;
    NF_REQ(v); NF_REQ(w); NF_REQ(w_hi); NF_REQ(w_lo); NF_REQ(b); NF_REQ(gstart_dev); NF_REQ(gstart_host);
    NF_REQ(workspace); NF_REQ(out); NF_REQ(ld);
    if ((D % 4) != 0 || (H % 4) != 0 || (block_degrees % 4) != 0) return NF_ERR_UNSUPPORTED;      // TMA row pitches of the slice GEMMs
    // every block of hidden units starts at a multiple of 4 (packing.blocked_made_pack pads the blocks with dead units):
    // 16-byte TMA bases of the push products' column slices, 128-bit row stores of the pull products.  (Multiples of 8
    // would allow 256-bit stores, but a 65-unit block padded to 72 costs the in-block kernel one of its six warps.)
    for (int g0 = 0; g0 < D; g0 += block_degrees)
        if (gstart_host[g0] % 4) return NF_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace;
    const size_t BH = (size_t)B * H, BD = (size_t)B * D;
    float* act1 = ws; float* act2 = act1 + BH; float* act3 = act2 + BH;
    float* pre1 = act3 + BH; float* pre2 = pre1 + BH; float* pre3 = pre2 + BH;
    float* preo = pre3 + BH; float* xcur = preo + 2 * BD; float* ldacc = xcur + BD;
    int* bad = reinterpret_cast<int*>(ldacc + B);
    const float* w0 = (const float*)w[0]; const float* w1 = (const float*)w[1];
    const float* w2 = (const float*)w[2]; const float* w3 = (const float*)w[3];
    const int grid = (int)cdiv(B, kBlkRows);
    bool pushed = false;
    for (int g0 = 0; g0 < D; g0 += block_degrees) {
        const int g1 = (g0 + block_degrees < D) ? g0 + block_degrees : D;
        const int u0 = gstart_host[g0], u1 = gstart_host[g1];
        const int nd = g1 - g0, nu = u1 - u0;
        const size_t smem = sizeof(float) * ((size_t)(nd + 3 * nu) * kBlkPad + 2 * kBlkRows);
        if (smem > 160 * 1024) return NF_ERR_UNSUPPORTED;
        const bool prev = g0 > 0;
        int rc;
        if (prev && nu > 0) {
            // PULL: contributions of dims < g0 / units < u0 (all final) to the block's hidden pre-activations, dense over
            // the whole batch (the output layer's share arrives by PUSH, below)
            rc = nf_linear_tc(xcur, (const float*)w_hi[0] + (size_t)u0 * D, (const float*)w_lo[0] + (size_t)u0 * D, nullptr,
                              pre1 + u0, B, nu, g0, D, D, H, 0, nullptr, stream);
            if (rc) return rc;
            if (u0 > 0) {
                rc = nf_linear_tc(act1, (const float*)w_hi[1] + (size_t)u0 * H, (const float*)w_lo[1] + (size_t)u0 * H, nullptr,
                                  pre2 + u0, B, nu, u0, H, H, H, 0, nullptr, stream);
                if (rc) return rc;
                rc = nf_linear_tc(act2, (const float*)w_hi[2] + (size_t)u0 * H, (const float*)w_lo[2] + (size_t)u0 * H, nullptr,
                                  pre3 + u0, B, nu, u0, H, H, H, 0, nullptr, stream);
                if (rc) return rc;
            }
        }
        const bool hp = prev && u0 > 0;
        const bool have_preo = pushed;               // some earlier block pushed into preo (it covers every later dim)
        bool launched = false;
        {
            // warp-private variant: shared transposed weights + one tile per warp; as many warps as fit in 220 KB
            const int need = ((nu + 3) & ~3) + 4 * kWarpChunks;
            const int nup = need <= 48 ? 48 : (need <= 80 ? 80 : (need <= 96 ? 96 : 160));   // compile-time pitches of ar_block_warp_kernel
            const size_t wfl = (size_t)(nd + 2 * nu) * nup + (size_t)nu * 2 * kBlkMaxDeg + ((3 * nu + 3) & ~3) + 2 * kBlkMaxDeg + 16;
            const size_t tfl = (size_t)(nd + 2 * nu) * kBlkPad;
            int nw = wfl * sizeof(float) < 200 * 1024 ? (int)((220 * 1024 / sizeof(float) - wfl) / tfl) : 0;
            if (nw > 8) nw = 8;
            if (g_ar_block_variant == 1 && nw >= 2 && need <= 160 && nd <= kBlkMaxDeg && (nd % 4) == 0) {
                const size_t smem2 = sizeof(float) * (wfl + (size_t)nw * tfl);
                const int64_t ctas = cdiv(cdiv(B, (int64_t)kBlkRows), (int64_t)nw);
                const int grid2 = (int)(ctas < kNumSMs ? ctas : kNumSMs);              // persistent: one CTA per SM
#define NF_ABW(NUPV)                                                                                                          \
                do {                                                                                                          \
                    NF_CUDA(cudaFuncSetAttribute(ar_block_warp_kernel<NUPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
                    ar_block_warp_kernel<NUPV><<<grid2, 32 * nw, smem2, st>>>(                                                 \
                        (const float*)v, xcur, prev ? pre1 : nullptr, hp ? pre2 : nullptr, hp ? pre3 : nullptr,                \
                        have_preo ? preo : nullptr, act1, act2, act3, w0, (const float*)b[0], w1, (const float*)b[1], w2,      \
                        (const float*)b[2], w3, (const float*)b[3], gstart_dev, ldacc, bad, B, D, H, g0, g1, u0, u1, mode);    \
                } while (0)
                if (nup == 48) NF_ABW(48); else if (nup == 80) NF_ABW(80); else if (nup == 96) NF_ABW(96); else NF_ABW(160);
#undef NF_ABW
                count_launch();
                NF_LAUNCH_CHECK();
                launched = true;
            }
        }
        if (!launched) {
            if (smem > 48 * 1024)
                NF_CUDA(cudaFuncSetAttribute(ar_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ar_block_kernel<<<grid, kBlkRows * kBlkWarps, smem, st>>>(
                (const float*)v, xcur, prev ? pre1 : nullptr, hp ? pre2 : nullptr, hp ? pre3 : nullptr, have_preo ? preo : nullptr,
                act1, act2, act3, w0, (const float*)b[0], w1, (const float*)b[1], w2, (const float*)b[2], w3, (const float*)b[3],
                gstart_dev, ldacc, bad, B, D, H, g0, g1, u0, u1, mode);
            count_launch();
            NF_LAUNCH_CHECK();
        }
        // PUSH: the block's layer-3 units are final -> their share of the output-layer pre-activations of ALL later dims in
        // one narrow-K product, accumulated into preo ([B, 2D], (mu, alpha) of a dim adjacent; w_hi[3] / w_lo[3] hold the
        // output layer's rows in that order).  Pulling per block instead re-read act3[:, :u0] for 2 x 8 outputs every time:
        // 1.85 ms of the 8.3 ms pass at [262144, 64] x 512 (profiles/r01z_c3_launches.csv).
        if (g1 < D && nu > 0) {
            rc = linear_tc_push(act3 + u0, (const float*)w_hi[3] + (size_t)2 * g1 * H + u0, (const float*)w_lo[3] + (size_t)2 * g1 * H + u0,
                                preo + 2 * g1, B, 2 * (D - g1), nu, H, H, 2 * D, pushed ? 1 : 0, st);
            if (rc) return rc;
            pushed = true;
        }
    }
    return nf_ar_finish_forward(xcur, v, ldacc, out, ld, B, D, mode, NF_F32, stream);
}
