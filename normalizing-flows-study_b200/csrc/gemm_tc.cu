// gemm_tc.cu -- dense layer on the 5th-gen tensor cores, one TMEM accumulation chain per output tile (the first
// kernel; nf_linear_tc* now routes through gemm_tc2.cu by default and falls back here):  Y[M,N] = act(X[M,K] * W[N,K]^T + bias)
//   (F.linear of MaskedLinear / MADE, masked_linear.py:14-18, made.py:136-140, and of the coupling / spline
//    conditioner MLPs, coupling_layer.py:18-35, spline_coupling_layer.py:55-62)
//
// 3xTF32: X is split on the fly (hi = top 19 bits, lo = X - hi), W is pre-split into W_hi / W_lo [N,K] arrays by
// nf_split_tf32 when the weights are folded; D += X_hi*W_hi + X_lo*W_hi + X_hi*W_lo accumulates in TMEM (fp32).
//
// One CTA computes a 128 x BN output tile (BN = 128; two CTAs co-reside per SM: 256 of 512 TMEM columns and ~97 KB of
// shared memory each) with warp-specialised roles over K blocks of 32 floats (one 128-byte swizzle atom):
//   warp 0      TMA producer: X tile [128x32] and W_hi / W_lo tiles [BNx32] -> shared memory (SWIZZLE_128B) per stage
//   warp 1      MMA issuer: 4 k-steps x 3 split passes = 12 tcgen05.mma (kind::tf32, M=128, N=BN) per K block,
//               A operand from TMEM, B operand from shared memory; tcgen05.commit releases the stages
//   warps 2-5   converters: thread r reads row r of the X tile from shared memory, splits it and stores the hi / lo
//               halves into the TMEM A stage (lane = row, column = k); after the K loop the same warps run the
//               epilogue: tcgen05.ld -> bias + ReLU -> 256-bit stores into the thread's own output row.
// Zero-tile skipping: k_extent[n/64] bounds the K loop of an output tile (block-lower-triangular MADE masks).
#include <cuda.h>
#include "nf_common.cuh"
#include "tc_common.cuh"

namespace nf {

#ifndef NF_GEMM_BN
#define NF_GEMM_BN 128
#endif
#ifndef NF_GEMM_STAGES
#define NF_GEMM_STAGES 2
#endif
constexpr int kGemmBM = 128, kGemmBN = NF_GEMM_BN, kGemmBK = 32;
constexpr int kGemmStages = NF_GEMM_STAGES;                 // shared-memory stages per CTA (x2 CTAs per SM)
constexpr int kGemmAStages = 2;                // TMEM A-operand stages
constexpr int kGemmThreads = 192;              // 6 warps
constexpr int kGemmTmemCols = 256;             // D: 128 | A stage 0: hi 32 + lo 32 | A stage 1: hi 32 + lo 32
constexpr int kColD = 0, kColA = 128;
constexpr uint32_t kXBytes = kGemmBM * kGemmBK * 4, kWBytes = kGemmBN * kGemmBK * 4;
constexpr uint32_t kStageBytes = kXBytes + 2 * kWBytes;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(tc::smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_wh,
               const __grid_constant__ CUtensorMap tm_wl, float* __restrict__ Y, const float* __restrict__ bias,
               int M, int N, int K, int64_t ldc, int relu, const int32_t* __restrict__ k_extent,
               const int32_t* __restrict__ k_begin, int passes, int vec) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // passes: 3 = 3xTF32, 1 = one TF32 pass (reduced-precision mode, nf_set_option(7, .), see gemm_tc2.cu)
    // [stage: X | W_hi | W_lo] x kGemmStages, then barriers
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGemmStages * kStageBytes);
    uint64_t* full = bars;                              // [S]  TMA landed
    uint64_t* empty = bars + kGemmStages;               // [S]  stage consumed (MMA commit)
    uint64_t* a_full = bars + 2 * kGemmStages;          // [A]  converters wrote the TMEM A stage
    uint64_t* a_empty = a_full + kGemmAStages;          // [A]  MMAs consumed the TMEM A stage
    uint64_t* d_full = a_empty + kGemmAStages;          // accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // 1-D grid, N tiles fastest: the CTAs sharing an X row-block run back to back and find it in L2
    const int n_tiles = (N + kGemmBN - 1) / kGemmBN;
    const int n0 = (int)(blockIdx.x % n_tiles) * kGemmBN, m0 = (int)(blockIdx.x / n_tiles) * kGemmBM;
    int k_end = K;
    if (k_extent) {
        int e = 0;
        for (int c = n0 / 64; c <= (n0 + kGemmBN - 1) / 64 && c * 64 < N; ++c) e = max(e, k_extent[c]);
        k_end = min(K, e);
    }
    // k_begin[n/64]: columns k < k_begin contribute exact zeros to these outputs (upper-triangular blocks of a
    // transposed MADE weight in the input-gradient product): the K loop starts at that block
    int kb_first = 0;
    if (k_begin) {
        int b = K;
        for (int c = n0 / 64; c <= (n0 + kGemmBN - 1) / 64 && c * 64 < N; ++c) b = min(b, k_begin[c]);
        kb_first = max(0, min(b, k_end)) / kGemmBK;
    }
    const int nkb = max(0, (k_end + kGemmBK - 1) / kGemmBK - kb_first);

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kGemmStages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < kGemmAStages; ++i) { tc::mbar_init(&a_full[i], 128); tc::mbar_init(&a_empty[i], 1); }
        tc::mbar_init(d_full, 1);
        tc::fence_mbar_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, kGemmTmemCols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kGemmStages;
                if (kb >= kGemmStages) tc::mbar_wait(&empty[s], ((kb / kGemmStages) - 1) & 1);
                uint8_t* st = smem + s * kStageBytes;
                tc::mbar_arrive_expect_tx(&full[s], passes == 1 ? kXBytes + kWBytes : kStageBytes);
                tma_load_2d(st, &tm_x, (kb_first + kb) * kGemmBK, m0, &full[s]);
                tma_load_2d(st + kXBytes, &tm_wh, (kb_first + kb) * kGemmBK, n0, &full[s]);
                if (passes != 1) tma_load_2d(st + kXBytes + kWBytes, &tm_wl, (kb_first + kb) * kGemmBK, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (whole warp, convergent; one elected lane issues) ----------------
        const uint32_t idesc = tc::idesc_tf32_m128((uint32_t)kGemmBN);
        const bool leader = tc::elect_one();
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kGemmStages, a = kb % kGemmAStages;
            tc::mbar_wait(&full[s], (kb / kGemmStages) & 1);
            tc::mbar_wait(&a_full[a], (kb / kGemmAStages) & 1);
            tc::fence_after_sync();
            const uint32_t st = tc::smem_u32(smem + s * kStageBytes);
            const uint64_t d_hi = tc::smem_desc_k_sw128(st + kXBytes), d_lo = tc::smem_desc_k_sw128(st + kXBytes + kWBytes);
            const uint32_t a_hi = tb + kColA + a * 64, a_lo = a_hi + 32;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                if (pass >= passes) break;
                const uint32_t ac = (pass == 1) ? a_lo : a_hi;
                const uint64_t wd = (pass == 2) ? d_lo : d_hi;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (leader) tc::mma_tf32_ts(tb + kColD, ac + k * 8, wd + (uint64_t)(k * 2), idesc, (kb | pass | k) != 0 ? 1u : 0u);
                }
            }
            if (leader) { tc::mma_commit(&empty[s]); tc::mma_commit(&a_empty[a]); }
            __syncwarp();
        }
        if (leader) tc::mma_commit(d_full);
        __syncwarp();
    } else {
        // ---------------- converters, then epilogue (warps 2..5; TMEM lane quadrant = warp % 4) ----------------
        const int q = warp & 3;                              // TMEM lanes [32q, 32q+32)
        const int r = q * 32 + lane;                         // tile row owned by this thread
        const uint32_t lane_addr = tb + ((uint32_t)(q * 32) << 16);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kGemmStages, a = kb % kGemmAStages;
            tc::mbar_wait(&full[s], (kb / kGemmStages) & 1);
            if (kb >= kGemmAStages) tc::mbar_wait(&a_empty[a], ((kb / kGemmAStages) - 1) & 1);
            tc::fence_after_sync();
            const uint8_t* xrow = smem + s * kStageBytes + (r >> 3) * 1024 + (r & 7) * 128;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(xrow + ((c ^ (r & 7)) << 4));
                tc::split_tf32(v.x, hi[4 * c + 0], lo[4 * c + 0]);
                tc::split_tf32(v.y, hi[4 * c + 1], lo[4 * c + 1]);
                tc::split_tf32(v.z, hi[4 * c + 2], lo[4 * c + 2]);
                tc::split_tf32(v.w, hi[4 * c + 3], lo[4 * c + 3]);
            }
            const uint32_t a_hi = lane_addr + kColA + a * 64, a_lo = a_hi + 32;
            {
                uint32_t t0[16], t1[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { t0[j] = hi[j]; t1[j] = hi[16 + j]; }
                tc::tmem_st16(a_hi, t0); tc::tmem_st16(a_hi + 16, t1);
                if (passes != 1) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) { t0[j] = lo[j]; t1[j] = lo[16 + j]; }
                    tc::tmem_st16(a_lo, t0); tc::tmem_st16(a_lo + 16, t1);
                }
            }
            tc::wait_st();
            tc::fence_before_sync();
            tc::mbar_arrive(&a_full[a]);
        }
        // epilogue: all stages are free once d_full fires; stage memory doubles as the transpose buffer
        tc::mbar_wait(d_full, 0);
        tc::fence_after_sync();
        // NOTE on accuracy: the tensor core truncates its fp32 accumulator on every MMA (measured, scripts/gemm_accuracy.py:
        // rms error 3e-6 / 7e-6 of rms(y) at K = 512 / 1024 against 4e-7 / 6e-7 for an FFMA GEMM; a bias of -T * 2^-25 for
        // same-sign sums).  Scaling the result by the expected loss was tried and dropped: it over-corrects mixed-sign sums
        // (C4 log-det bias +4.4e-4 -> -3.3e-4).  Shorter chains (rotating accumulators) are the fix; see DESIGN.md.
        // The thread owns row m0 + r: bias + ReLU + one 256-bit store per 8 columns (tc::epilogue_store8).
        const int row = m0 + r;
        float* yrow = Y + (int64_t)row * ldc;
        // every stage is free once d_full has fired: bias tile (aligned outputs) or per-warp transpose buffers (unaligned)
        float* bias_s = reinterpret_cast<float*>(smem);
        float* tbuf = reinterpret_cast<float*>(smem) + (size_t)q * 32 * 33;
        if (vec) tc::stage_bias_tile(bias_s, bias, n0, N, (warp - 2) * 32 + lane, 1);
#pragma unroll
        for (int c = 0; c < kGemmBN / 32; ++c) {
            uint32_t v0[16], v1[16];
            if (nkb > 0) {
                tc::tmem_ld16(lane_addr + kColD + c * 32, v0);
                tc::tmem_ld16(lane_addr + kColD + c * 32 + 16, v1);
                tc::wait_ld();
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) { v0[j] = 0u; v1[j] = 0u; }
            }
            if (vec) {
                if (row < M) tc::epilogue_store32(yrow, n0 + c * 32, N, v0, v1, bias_s + c * 32, relu, true);
            } else {
                tc::epilogue_store32_transposed(Y, ldc, m0 + q * 32, M, n0 + c * 32, N, v0, v1, bias, relu, tbuf, lane);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tb, kGemmTmemCols);
}

// W ~= hi + lo: hi = W rounded to the nearest TF32, lo = (W - hi) rounded to the nearest TF32 (|W - hi - lo| <= 2^-24 |W|)
__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t h, l;
        tc::split_tf32_weight(w[i], h, l);
        hi[i] = __uint_as_float(h);
        lo[i] = __uint_as_float(l);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major fp32 [rows, cols] (leading dimension ld elements), box = [box_rows x 32 floats], SWIZZLE_128B
static bool make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern int g_gemm_tc_variant;
extern int g_gemm_tc_small_k;
extern int g_tc_passes;
int gemm_tc2_launch(const void* x, const void* w_hi, const void* w_lo, const void* bias, void* y, int64_t M, int64_t N, int64_t K,
                    int64_t ldx, int64_t ldw, int64_t ldy, int relu, const int32_t* k_begin, const int32_t* k_extent,
                    cudaStream_t st, int accumulate);

}  // namespace nf

using namespace nf;
#define NF_REQ(p) do { if ((p) == nullptr) return NF_ERR_NULL; } while (0)

extern "C" int nf_split_tf32(const void* w, void* w_hi, void* w_lo, int64_t n, nf_stream_t stream) {
    if (n < 0) return NF_ERR_BAD_SHAPE;
    if (n == 0) return NF_OK;
    NF_REQ(w); NF_REQ(w_hi); NF_REQ(w_lo);
    int64_t need = cdiv(n, 256), cap = (int64_t)kNumSMs * 8;
    split_tf32_kernel<<<(int)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>((const float*)w, (float*)w_hi, (float*)w_lo, n);
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}

extern "C" int nf_linear_tc_range(const void* x, const void* w_hi, const void* w_lo, const void* bias, void* y, int64_t M,
                                  int64_t N, int64_t K, int64_t ldx, int64_t ldw, int64_t ldy, int relu,
                                  const int32_t* k_begin, const int32_t* k_extent, nf_stream_t stream);

extern "C" int nf_linear_tc(const void* x, const void* w_hi, const void* w_lo, const void* bias, void* y, int64_t M,
                            int64_t N, int64_t K, int64_t ldx, int64_t ldw, int64_t ldy, int relu,
                            const int32_t* k_extent, nf_stream_t stream) {
    return nf_linear_tc_range(x, w_hi, w_lo, bias, y, M, N, K, ldx, ldw, ldy, relu, nullptr, k_extent, stream);
}

extern "C" int nf_linear_tc_range(const void* x, const void* w_hi, const void* w_lo, const void* bias, void* y, int64_t M,
                                  int64_t N, int64_t K, int64_t ldx, int64_t ldw, int64_t ldy, int relu,
                                  const int32_t* k_begin, const int32_t* k_extent, nf_stream_t stream) {
    if (M < 0 || N < 1 || K < 1 || ldx < K || ldw < K || ldy < N) return NF_ERR_BAD_SHAPE;
    if (M == 0) return NF_OK;
    NF_REQ(x); NF_REQ(w_hi); NF_REQ(w_lo); NF_REQ(y);
    // TMA: 16-byte aligned bases and row pitches
    if (!aligned16(x) || !aligned16(w_hi) || !aligned16(w_lo) || (ldx % 4) != 0 || (ldw % 4) != 0) return NF_ERR_UNSUPPORTED;
    if (M > 2147483647LL - 128 || N > 2147483647LL - 128 || K > 2147483647LL - 64) return NF_ERR_BAD_SHAPE;
    // contractions longer than 128 go to the short-chain kernel (gemm_tc2.cu: chains of 4 K blocks folded into registers
    // with round-to-nearest adds); up to K = 128 the single chain here is just as short and two CTAs per SM are faster
    if (g_gemm_tc_variant == 1 && (K > 128 || g_gemm_tc_small_k)) {
        const int rc = gemm_tc2_launch(x, w_hi, w_lo, bias, y, M, N, K, ldx, ldw, ldy, relu, k_begin, k_extent, (cudaStream_t)stream, 0);
        if (rc == NF_OK) { count_launch(); NF_LAUNCH_CHECK(); return NF_OK; }
        if (rc != NF_ERR_UNSUPPORTED) return rc;
    }
    alignas(64) CUtensorMap tx, twh, twl;
    if (!make_map(&tx, x, M, K, ldx, kGemmBM) || !make_map(&twh, w_hi, N, K, ldw, kGemmBN) || !make_map(&twl, w_lo, N, K, ldw, kGemmBN))
        return NF_ERR_UNSUPPORTED;
    const size_t smem = (size_t)kGemmStages * kStageBytes + 256;
    NF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int64_t nblocks = cdiv(N, kGemmBN) * cdiv(M, kGemmBM);
    if (nblocks > 2147483647LL) return NF_ERR_BAD_SHAPE;
    gemm_tc_kernel<<<(unsigned)nblocks, kGemmThreads, smem, (cudaStream_t)stream>>>(tx, twh, twl, (float*)y, (const float*)bias, (int)M, (int)N,
                                                                       (int)K, ldy, relu, k_extent, k_begin, g_tc_passes,
                                                                       (aligned32(y) && (ldy % 8) == 0) ? 1 : 0);
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}
