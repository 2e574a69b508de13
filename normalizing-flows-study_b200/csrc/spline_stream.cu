// spline_stream.cu -- a5/a6 SplineCouplingLayer transform (spline_coupling_layer.py:96-309), float32, compact parameter
// layout [B, Dt*(3K-1)], K in {8, 10}: the HBM-bound forward / inverse / backward passes of the layered (training) route.
//
// Second version of spline_transform_compact_{fwd,bwd}_kernel (transform_impl.cuh), which stay as the fallback for every
// other shape.  What the first version paid per element (profiles/r02p_spline_tf_ncu.txt: 744 instructions per element
// forward against 526 for the same arithmetic in rqs_unit_fwd_kernel; 2 360 backward, 71 % / 55 % of the issue slots
// busy at 56 % / 26 % of the HBM rate):
//   * the parameter blocks of a warp's 32 elements were fetched with six cp.async per lane plus their address
//     arithmetic, waited for on the spot, and (backward) written back with six LDS.128 + STG.128 per lane
//       -> ONE bulk copy through the TMA unit per warp and chunk (cp.async.bulk, mbarrier completion), issued by one
//          lane a whole chunk ahead (two slabs per warp); the gradient slab leaves through one bulk store;
//   * x / y / gy went through per-dimension mask and index loads with 64-bit address arithmetic per access
//       -> data_dim == 2 (the C1 / C2 family): one 8-byte load and one 8-byte store per row; the next chunk's inputs
//          are requested before the current chunk is evaluated;
//   * the runtime `inverse` flag kept both branches of the bin evaluation live -> template parameter;
//   * backward: the element was evaluated three times with IEEE divisions -> rqs_eval_grad (nf_math.cuh).
#include "nf_common.cuh"
#include "tc_common.cuh"

namespace nf {

using tc::bulk_g2s;
using tc::fence_mbar_init;
using tc::fence_proxy_async_smem;
using tc::mbar_arrive_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

int g_spline_stream = 1;          // nf_set_option(11, v): 0 = first-version kernels only

__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }


// Work decomposition.  A warp walks over ITEMS and, inside an item, over CHUNKS of up to 32 parameter blocks (one per
// lane).  MODE 0 (G >= Dt, G lanes per row): an item = RPW = 32 / G rows = one chunk, the rows' Dt blocks each being
// contiguous in memory.  MODE 1: MODE 0 for data_dim == 2 (one 8-byte load / store per row, no mask or index loads).
// MODE 2 (G == 32 < Dt): an item = one row, chunks of 32 consecutive transformed dims.  All per-chunk bookkeeping is
// incremental (running row number and parameter pointer): the first version of this file recomputed 64-bit offsets
// three times per chunk and spent ~80 of its 722 instructions per element on it.
enum { kModeGroup = 0, kModeD2 = 1, kModeWide = 2 };

struct ChunkState {
    int row;                // first row of the item (MODE 0 / 1) or the row (MODE 2); the launcher bounds B below 2^31

    const float* src;       // the chunk's parameter blocks
    int ch;                 // chunk within the row (MODE 2)
    int nfl;                // floats in the chunk
};

template <int P, int RPW, int MODE>
struct ChunkWalk {
    int64_t src_stride;
    int B, row_stride, Dt, nch, full_fl;
    __device__ __forceinline__ bool live(const ChunkState& c) const { return c.row < B; }
    __device__ __forceinline__ void start(ChunkState& c, const float* params, int64_t warp) const {
        c.row = (int)warp * RPW; c.ch = 0;
        c.src = params + warp * (MODE == kModeWide ? (int64_t)Dt * P : (int64_t)full_fl);
        size(c);
    }
    __device__ __forceinline__ void size(ChunkState& c) const {
        if (MODE == kModeWide) { const int tn = Dt - c.ch * 32; c.nfl = (tn < 32 ? tn : 32) * P; }
        else { const int left = B - c.row; c.nfl = left >= RPW ? full_fl : left * Dt * P; }
    }
    __device__ __forceinline__ void advance(ChunkState& c) const {
        if (MODE == kModeWide) {
            if (++c.ch == nch) { c.ch = 0; c.row += row_stride; c.src += src_stride - (int64_t)(nch - 1) * 32 * P; }
            else c.src += 32 * P;
        } else { c.row += row_stride; c.src += src_stride; }
        size(c);
    }
};

// rqs_eval<float, K, true> with the direction chosen at compile time and the raw derivative parameters of the two
// knots read from the slab by index (two LDS instead of two 7-select multiplexer trees); same operations otherwise,
// bit-identical results
template <int KMAX, bool INV>
__device__ __forceinline__ void eval_elem(float v, const float* pp, const RqsCfg<float>& c, float& out, float& lad) {
#if defined(__CUDA_ARCH__)           // rqs_knots_pair exists in the device pass only
    constexpr int K = KMAX;
    const bool inside = (v >= c.lo) && (v <= c.hi);
    if (!inside) { out = v; lad = 0.f; return; }
    float uw[KMAX], uh[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) { uw[j] = pp[j]; uh[j] = pp[K + j]; }
    float wn[KMAX], hn[KMAX], cw[KMAX + 1], ch[KMAX + 1];
    rqs_knots_pair<KMAX, true>(uw, uh, K, c, wn, hn, cw, ch);
    const float* sk = INV ? ch : cw;
    int cnt[KMAX];
    cnt[0] = 0;
#pragma unroll
    for (int j = 1; j < KMAX; ++j) cnt[j] = (sk[j] <= v) ? 1 : 0;
    const int k = tree_sum<int, KMAX>(cnt);
    const float xk = mux<float, KMAX>(cw, k), xk1 = mux<float, KMAX>(cw + 1, k);
    const float yk = mux<float, KMAX>(ch, k), yk1 = mux<float, KMAX>(ch + 1, k);
    const float udk = pp[2 * K + (k >= 1 ? k - 1 : 0)], udk1 = pp[2 * K + (k < K - 1 ? k : K - 2)];
    float spk, spk1;
    softplus_x2(udk, udk1, spk, spk1);
    const float dk = (k == 0) ? 1.f : clamp_min(c.min_d + spk, c.eps);
    const float dk1 = (k == K - 1) ? 1.f : clamp_min(c.min_d + spk1, c.eps);
    const float wk = clamp_min(xk1 - xk, c.eps), hk = clamp_min(yk1 - yk, c.eps);
    rqs_bin_eval<float, true>(v, xk, yk, wk, hk, dk, dk1, INV, c.eps, out, lad);
    if (!is_finite(out)) out = v;
    if (!is_finite(lad)) lad = 0.f;
#endif
}

constexpr int kIdU = 4;       // identity dims per lane in flight at once

template <int KMAX, int G, bool INV, int MODE>
__global__ void __launch_bounds__(128, 8)
spline_stream_fwd_kernel(const SplineStreamArgs a) {
    constexpr int K = KMAX, P = 3 * K - 1, RPW = 32 / G, SLAB = 32 * P;
    constexpr bool D2 = MODE == kModeD2, WIDE = MODE == kModeWide;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwc = blockDim.x >> 5;
    float* slab = reinterpret_cast<float*>(smem_raw) + (size_t)wib * 2 * SLAB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nwc * 2 * SLAB * sizeof(float)) + 2 * wib;
    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    __syncwarp();
    const int g = lane % G, rsub = lane / G;
    const int D = a.D, Dt = a.Dt;
    const int B = (int)a.B;
    const int64_t warp = (int64_t)blockIdx.x * nwc + wib, nwarps = (int64_t)gridDim.x * nwc;
    ChunkWalk<P, RPW, MODE> walk;
    walk.B = B; walk.Dt = Dt; walk.nch = WIDE ? (Dt + 31) >> 5 : 1; walk.full_fl = RPW * Dt * P;
    walk.row_stride = (int)nwarps * RPW; walk.src_stride = nwarps * (WIDE ? (int64_t)Dt * P : (int64_t)walk.full_fl);
    const RqsCfg<float> c = a.c;
    const bool resc = a.r_in != nullptr;
    auto issue = [&](const ChunkState& cs, uint32_t q) {             // one lane: bulk copy of the chunk into slab q & 1
        if (walk.live(cs) && (cs.nfl & 3) == 0 && lane == 0) {
            mbar_arrive_expect_tx(&bars[q & 1], (uint32_t)cs.nfl * 4u);
            bulk_g2s(slab + (q & 1) * SLAB, cs.src, (uint32_t)cs.nfl * 4u, &bars[q & 1]);
        }
    };
    struct In { float2 x; int dim; };
    const int dim_fixed = D2 ? __ldg(a.tidx) : ((!WIDE && g < Dt) ? __ldg(a.tidx + g) : 0);
    auto fetch = [&](const ChunkState& cs, In& in) {                 // this lane's inputs of the chunk, requested early
        in.x = make_float2(0.f, 0.f); in.dim = dim_fixed;
        const int row = cs.row + rsub;
        if (!walk.live(cs) || row >= B) return;
        if (D2) { in.x = __ldcs(reinterpret_cast<const float2*>(a.x) + row); return; }
        const float* xr = a.x + (int64_t)row * D;
        const int t = WIDE ? cs.ch * 32 + g : g;
        if (t < Dt) { if (WIDE) in.dim = __ldg(a.tidx + t); in.x.x = xr[in.dim]; }
    };

    ChunkState cur, nxt;
    walk.start(cur, a.params, warp);
    nxt = cur;
    issue(cur, 0);
    In in;
    fetch(cur, in);
    float acc = 0.f;
    for (uint32_t q = 0; walk.live(cur); ++q) {
        walk.advance(nxt);
        __syncwarp();                                    // every lane is done with the slab the next copy lands in
        issue(nxt, q + 1);
        In inn;
        fetch(nxt, inn);
        const int row = cur.row + rsub;
        const bool vrow = row < B;
        if (!D2 && vrow && (!WIDE || cur.ch == 0)) {     // identity dims of the row (layer-level scrub, :130), the row's
            const float* xr = a.x + (int64_t)row * D;    // G lanes side by side, four dims per lane in flight
            float* yr = a.y + (int64_t)row * D;
            for (int dd0 = g; dd0 < D; dd0 += kIdU * G) {
                float xi[kIdU];
                bool on[kIdU];
#pragma unroll
                for (int u = 0; u < kIdU; ++u) {
                    const int dd = dd0 + u * G;
                    on[u] = dd < D && __ldg(a.mask + dd) != 0.f;
                    xi[u] = on[u] ? xr[dd] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < kIdU; ++u) if (on[u]) yr[dd0 + u * G] = scrub0(xi[u]);
            }
        }
        float* sl = slab + (q & 1) * SLAB;
        if ((cur.nfl & 3) == 0) {
            mbar_wait(&bars[q & 1], (q >> 1) & 1);
        } else {                                         // ragged last item: plain loads
            for (int i = lane; i < cur.nfl; i += 32) sl[i] = __ldcs(cur.src + i);
            __syncwarp();
        }
        const int t = WIDE ? cur.ch * 32 + g : g;
        if (vrow && t < Dt) {
            const int e = WIDE ? g : rsub * Dt + g;
            float v = D2 ? (dim_fixed ? in.x.y : in.x.x) : in.x.x;
            if (resc) v = a.r_in[in.dim] * (v - a.r_lo[in.dim]) - c.hi;
            float out, lad;
            eval_elem<KMAX, INV>(v, sl + e * P, c, out, lad);
            if (resc) out = (out + c.hi) * a.r_out[in.dim] + a.r_lo[in.dim];
            out = scrub0(out);
            if (D2) {
                const float keep = scrub0(dim_fixed ? in.x.x : in.x.y);
                __stcs(reinterpret_cast<float2*>(a.y) + row, dim_fixed ? make_float2(keep, out) : make_float2(out, keep));
            } else {
                a.y[(int64_t)row * D + in.dim] = out;
            }
            acc += lad;
        }
        if (!WIDE || cur.ch == walk.nch - 1) {
            acc = group_sum<float, G>(acc);
            if (vrow && g == 0) a.ld[row] = scrub0(acc);
            acc = 0.f;
        }
        cur = nxt; in = inn;
    }
}

// upstream-gradient adjustment of the layer: the element's output after the spline's own scrub (:306-307), the optional
// rescale and the layer-level scrub of y (:130) -- a replaced output receives no gradient
struct LayerAdj {
    float v, r_out, r_lo, hi; bool resc;
    __device__ __forceinline__ void operator()(float out, float lad, float& g_out, float& g_lad, float& gv_direct) const {
        float o = is_finite(out) ? out : v;
        if (resc) o = (o + hi) * r_out + r_lo;
        float go = is_finite(o) ? g_out : 0.f;
        if (resc) go *= r_out;
        if (!is_finite(lad)) g_lad = 0.f;
        if (!is_finite(out)) { gv_direct = go; go = 0.f; }
        g_out = go;
    }
};

template <int KMAX, int G, bool INV, int MODE>
__global__ void __launch_bounds__(128, 5)
spline_stream_bwd_kernel(const SplineStreamArgs a) {
    constexpr int K = KMAX, P = 3 * K - 1, RPW = 32 / G, SLAB = 32 * P;
    constexpr bool D2 = MODE == kModeD2, WIDE = MODE == kModeWide;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwc = blockDim.x >> 5;
    float* slab = reinterpret_cast<float*>(smem_raw) + (size_t)wib * 3 * SLAB;      // two input slabs + the gradient slab
    float* oslab = slab + 2 * SLAB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nwc * 3 * SLAB * sizeof(float)) + 2 * wib;
    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    __syncwarp();
    const int g = lane % G, rsub = lane / G;
    const int D = a.D, Dt = a.Dt;
    const int B = (int)a.B;
    const int64_t warp = (int64_t)blockIdx.x * nwc + wib, nwarps = (int64_t)gridDim.x * nwc;
    ChunkWalk<P, RPW, MODE> walk;
    walk.B = B; walk.Dt = Dt; walk.nch = WIDE ? (Dt + 31) >> 5 : 1; walk.full_fl = RPW * Dt * P;
    walk.row_stride = (int)nwarps * RPW; walk.src_stride = nwarps * (WIDE ? (int64_t)Dt * P : (int64_t)walk.full_fl);
    const RqsCfg<float> c = a.c;
    const bool resc = a.r_in != nullptr;
    const int64_t gdelta = a.gparams - a.params;         // the gradient block sits where the parameter block does

    auto issue = [&](const ChunkState& cs, uint32_t q) {
        if (walk.live(cs) && (cs.nfl & 3) == 0 && lane == 0) {
            mbar_arrive_expect_tx(&bars[q & 1], (uint32_t)cs.nfl * 4u);
            bulk_g2s(slab + (q & 1) * SLAB, cs.src, (uint32_t)cs.nfl * 4u, &bars[q & 1]);
        }
    };
    struct In { float2 x, gy; float gl; int dim; };
    const int dim_fixed = D2 ? __ldg(a.tidx) : ((!WIDE && g < Dt) ? __ldg(a.tidx + g) : 0);
    auto fetch = [&](const ChunkState& cs, In& in) {
        in.x = make_float2(0.f, 0.f); in.gy = in.x; in.gl = 0.f; in.dim = dim_fixed;
        const int row = cs.row + rsub;
        if (!walk.live(cs) || row >= B) return;
        in.gl = __ldg(a.gld + row);
        if (D2) {
            in.x = __ldcs(reinterpret_cast<const float2*>(a.x) + row);
            in.gy = __ldcs(reinterpret_cast<const float2*>(a.gy) + row);
            return;
        }
        const float* xr = a.x + (int64_t)row * D;
        const float* gr = a.gy + (int64_t)row * D;
        const int t = WIDE ? cs.ch * 32 + g : g;
        if (t < Dt) { if (WIDE) in.dim = __ldg(a.tidx + t); in.x.x = xr[in.dim]; in.gy.x = gr[in.dim]; }
    };

    ChunkState cur, nxt;
    walk.start(cur, a.params, warp);
    nxt = cur;
    issue(cur, 0);
    In in;
    fetch(cur, in);
    for (uint32_t q = 0; walk.live(cur); ++q) {
        walk.advance(nxt);
        __syncwarp();
        issue(nxt, q + 1);
        In inn;
        fetch(nxt, inn);
        const int row = cur.row + rsub;
        const bool vrow = row < B;
        if (!D2 && vrow && (!WIDE || cur.ch == 0)) {     // identity dims: the layer-level scrub zeroes non-finite inputs
            const float* xr = a.x + (int64_t)row * D;
            const float* gr = a.gy + (int64_t)row * D;
            float* gxr = a.gx + (int64_t)row * D;
            for (int dd0 = g; dd0 < D; dd0 += kIdU * G) {
                float xi[kIdU], gi[kIdU];
                bool on[kIdU];
#pragma unroll
                for (int u = 0; u < kIdU; ++u) {
                    const int dd = dd0 + u * G;
                    on[u] = dd < D && __ldg(a.mask + dd) != 0.f;
                    xi[u] = on[u] ? xr[dd] : 0.f;
                    gi[u] = on[u] ? gr[dd] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < kIdU; ++u) if (on[u]) gxr[dd0 + u * G] = is_finite(xi[u]) ? gi[u] : 0.f;
            }
        }
        const bool bulk = (cur.nfl & 3) == 0;
        float* sl = slab + (q & 1) * SLAB;
        if (bulk) {
            mbar_wait(&bars[q & 1], (q >> 1) & 1);
        } else {
            for (int i = lane; i < cur.nfl; i += 32) sl[i] = __ldcs(cur.src + i);
            __syncwarp();
        }
        const int t = WIDE ? cur.ch * 32 + g : g;
        const bool act = vrow && t < Dt;
        const int e = WIDE ? g : rsub * Dt + g;
        float gv = 0.f, guw[KMAX], guh[KMAX], gud[KMAX];
        if (act) {
            const float* pp = sl + e * P;
            float uw[KMAX], uh[KMAX], ud[KMAX];
#pragma unroll
            for (int j = 0; j < KMAX; ++j) { uw[j] = pp[j]; uh[j] = pp[K + j]; ud[j] = (j < K - 1) ? pp[2 * K + j] : 0.f; }
            float v = D2 ? (dim_fixed ? in.x.y : in.x.x) : in.x.x;
            const float go = D2 ? (dim_fixed ? in.gy.y : in.gy.x) : in.gy.x;
            LayerAdj adj{0.f, 1.f, 0.f, c.hi, resc};
            if (resc) { v = a.r_in[in.dim] * (v - a.r_lo[in.dim]) - c.hi; adj.r_out = a.r_out[in.dim]; adj.r_lo = a.r_lo[in.dim]; }
            adj.v = v;
            rqs_eval_grad<float, KMAX, true>(v, uw, uh, ud, K, INV, c, go, in.gl, adj, gv, guw, guh, gud);
            if (resc) gv *= a.r_in[in.dim];
        }
        if (lane == 0) bulk_wait_read0();                // the previous chunk's bulk store has read the gradient slab
        __syncwarp();
        if (act) {
            float* op = oslab + e * P;
#pragma unroll
            for (int j = 0; j < KMAX; ++j) { op[j] = guw[j]; op[K + j] = guh[j]; if (j < K - 1) op[2 * K + j] = gud[j]; }
            if (D2) {
                const float o = dim_fixed ? in.x.x : in.x.y, go_o = dim_fixed ? in.gy.x : in.gy.y;
                const float keep = is_finite(o) ? go_o : 0.f;
                __stcs(reinterpret_cast<float2*>(a.gx) + row, dim_fixed ? make_float2(keep, gv) : make_float2(gv, keep));
            } else {
                a.gx[(int64_t)row * D + in.dim] = gv;
            }
        }
        float* dst = const_cast<float*>(cur.src) + gdelta;
        if (bulk) {
            fence_proxy_async_smem();                    // this lane's st.shared -> visible to the bulk store
            __syncwarp();
            if (lane == 0) { bulk_s2g(dst, oslab, (uint32_t)cur.nfl * 4u); bulk_commit(); }
        } else {
            __syncwarp();
            for (int i = lane; i < cur.nfl; i += 32) __stcs(dst + i, oslab[i]);
        }
        cur = nxt; in = inn;
    }
    if (lane == 0) bulk_wait_read0();                    // shared memory stays valid until the last store has read it
}

template <int KMAX, int G>
static int stream_launch_g(const SplineStreamArgs& a, bool bwd, int inverse, cudaStream_t st) {
    constexpr int P = 3 * KMAX - 1, RPW = 32 / G;
    const bool d2 = (a.D == 2 && a.Dt == 1 && G == 1 && (reinterpret_cast<uintptr_t>(a.x) & 7) == 0 &&
                     (bwd ? ((reinterpret_cast<uintptr_t>(a.gx) | reinterpret_cast<uintptr_t>(a.gy)) & 7) == 0
                          : (reinterpret_cast<uintptr_t>(a.y) & 7) == 0));
    const bool wide = (G == 32 && a.Dt > 32);
    const int nwc = 4;
    const size_t smem = (size_t)nwc * (bwd ? 3 : 2) * 32 * P * sizeof(float) + (size_t)nwc * 2 * sizeof(uint64_t);
    const int64_t nitems = cdiv(a.B, RPW);
    const int64_t ctas = cdiv(nitems, nwc);
    // persistent warps: exactly one wave of resident CTAs (registers: 8 / 5 per SM; shared memory: 7 / 4 with 10 bins --
    // a grid of 8 / 5 per SM there ran a second, quarter-full wave: 57 % instead of 66 % of HBM at D = 16, K = 10)
#define NF_SS(KERN, INV, MODE)                                                                                         \
    do {                                                                                                               \
        auto kern = KERN<KMAX, G, INV, MODE>;                                                                          \
        NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
        NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        int per_sm = 0;                                                                                                \
        NF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * nwc, smem));                         \
        if (per_sm < 1) return NF_ERR_UNSUPPORTED;                                                                     \
        const int64_t cap = (int64_t)kNumSMs * per_sm;                                                                 \
        const int grid = (int)(ctas < cap ? ctas : cap);                                                               \
        kern<<<grid, 32 * nwc, smem, st>>>(a);                                                                         \
    } while (0)
#define NF_SS_DIR(KERN, MODE) do { if (inverse) NF_SS(KERN, true, MODE); else NF_SS(KERN, false, MODE); } while (0)
    constexpr int kNarrow = (G == 1) ? kModeD2 : kModeGroup, kWide = (G == 32) ? kModeWide : kModeGroup;
    if (bwd) {
        if (d2) NF_SS_DIR(spline_stream_bwd_kernel, kNarrow);
        else if (wide) NF_SS_DIR(spline_stream_bwd_kernel, kWide);
        else NF_SS_DIR(spline_stream_bwd_kernel, kModeGroup);
    } else {
        if (d2) NF_SS_DIR(spline_stream_fwd_kernel, kNarrow);
        else if (wide) NF_SS_DIR(spline_stream_fwd_kernel, kWide);
        else NF_SS_DIR(spline_stream_fwd_kernel, kModeGroup);
    }
#undef NF_SS_DIR
#undef NF_SS
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

// NF_ERR_UNSUPPORTED = not taken (the caller falls back to the first-version kernels)
int spline_stream_launch(const SplineStreamArgs& a, int K, bool bwd, int inverse, cudaStream_t st) {
    if (!g_spline_stream || (K != 8 && K != 10) || a.Dt < 1 || a.B > (int64_t)0x7ff00000) return NF_ERR_UNSUPPORTED;
    if (!aligned16(a.params) || (bwd && !aligned16(a.gparams))) return NF_ERR_UNSUPPORTED;
    int G = 1;
    while (G < a.Dt && G < 32) G <<= 1;
    // every full chunk must be a whole number of 16-byte units (the bulk copies' granularity); P is odd
    if (a.Dt > 32 ? (a.Dt % 4) != 0 : (((32 / G) * a.Dt) % 4) != 0) return NF_ERR_UNSUPPORTED;
#define NF_SG(KM)                                                                                     \
    switch (G) {                                                                                      \
        case 1: return stream_launch_g<KM, 1>(a, bwd, inverse, st);                                   \
        case 2: return stream_launch_g<KM, 2>(a, bwd, inverse, st);                                   \
        case 4: return stream_launch_g<KM, 4>(a, bwd, inverse, st);                                   \
        case 8: return stream_launch_g<KM, 8>(a, bwd, inverse, st);                                   \
        case 16: return stream_launch_g<KM, 16>(a, bwd, inverse, st);                                 \
        default: return stream_launch_g<KM, 32>(a, bwd, inverse, st);                                 \
    }
    if (K == 8) { NF_SG(8) } else { NF_SG(10) }
#undef NF_SG
}

}  // namespace nf
