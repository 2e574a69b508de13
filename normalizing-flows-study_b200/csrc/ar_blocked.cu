// ar_blocked.cu -- sequential directions of the affine autoregressive flows (MAF.forward / IAF.inverse,
// masked_autoregressive_flow.py:46-78 / inverse_autoregressive_flow.py:65-103) as a *blocked* triangular evaluation.
//
// The reference re-evaluates the whole MADE D times (D x 4 dense GEMMs).  Hidden units sorted by degree make every
// masked weight block-lower-triangular, so the D dependent steps are grouped into blocks of `gb` consecutive degrees:
//   * contributions of all PREVIOUS blocks to the block's pre-activations are plain dense products over the whole
//     batch -> nf_linear_tc (tcgen05, 3xTF32, TMA) on column slices of the activation buffers:
//         pre_l[:, blk] = act_{l-1}[:, :u0] * W_l[blk, :u0]^T        (l = 1..3),   preo[:, dims] = act3[:, :u0] * W3[dims, :u0]^T
//   * the IN-BLOCK part (gb dependent steps over <= ~72 units per layer) runs in ar_block_kernel: 32 rows per CTA,
//     ~30 KB of shared memory, so 6-7 CTAs per SM hide the step-to-step barrier latency that bounds the one-launch
//     incremental kernel (made_kernels.cu: one CTA per SM at D=64, H=512).
// Every hidden unit is still evaluated exactly once (total work = one masked MADE pass), ~75 % of it on the tensor pipe.
#include "nf_common.cuh"

extern "C" int nf_linear_tc(const void*, const void*, const void*, const void*, void*, int64_t, int64_t, int64_t, int64_t,
                            int64_t, int64_t, int, const int32_t*, nf_stream_t);
extern "C" int nf_ar_finish_forward(const void*, const void*, const void*, void*, void*, int64_t, int, int, int, nf_stream_t);

namespace nf {

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
            if (preo && lane < nrow) { mu += preo[(r0 + lane) * 2 * D + g]; al += preo[(r0 + lane) * 2 * D + D + g]; }
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
    if ((D % 4) != 0 || (H % 4) != 0) return NF_ERR_UNSUPPORTED;      // TMA row pitches of the slice GEMMs
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
    for (int g0 = 0; g0 < D; g0 += block_degrees) {
        const int g1 = (g0 + block_degrees < D) ? g0 + block_degrees : D;
        const int u0 = gstart_host[g0], u1 = gstart_host[g1];
        const int nd = g1 - g0, nu = u1 - u0;
        const size_t smem = sizeof(float) * ((size_t)(nd + 3 * nu) * kBlkPad + 2 * kBlkRows);
        if (smem > 160 * 1024) return NF_ERR_UNSUPPORTED;
        const bool prev = g0 > 0;
        int rc;
        if (prev) {
            // contributions of dims < g0 / units < u0 (all final) to this block, dense over the whole batch
            if (nu > 0) {
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
            if (u0 > 0) {
                rc = nf_linear_tc(act3, (const float*)w_hi[3] + (size_t)g0 * H, (const float*)w_lo[3] + (size_t)g0 * H, nullptr,
                                  preo + g0, B, nd, u0, H, H, 2 * D, 0, nullptr, stream);
                if (rc) return rc;
                rc = nf_linear_tc(act3, (const float*)w_hi[3] + (size_t)(D + g0) * H, (const float*)w_lo[3] + (size_t)(D + g0) * H,
                                  nullptr, preo + D + g0, B, nd, u0, H, H, 2 * D, 0, nullptr, stream);
                if (rc) return rc;
            }
        }
        const bool hp = prev && u0 > 0;
        if (smem > 48 * 1024)
            NF_CUDA(cudaFuncSetAttribute(ar_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ar_block_kernel<<<grid, kBlkRows * kBlkWarps, smem, st>>>(
            (const float*)v, xcur, prev ? pre1 : nullptr, hp ? pre2 : nullptr, hp ? pre3 : nullptr, hp ? preo : nullptr,
            act1, act2, act3, w0, (const float*)b[0], w1, (const float*)b[1], w2, (const float*)b[2], w3, (const float*)b[3],
            gstart_dev, ldacc, bad, B, D, H, g0, g1, u0, u1, mode);
        count_launch();
        NF_LAUNCH_CHECK();
    }
    return nf_ar_finish_forward(xcur, v, ldacc, out, ld, B, D, mode, NF_F32, stream);
}
