// gemm_bf16_probe.cu -- stand-alone probe (NOT part of libnfb200): what a bf16-operand dense layer can reach on B200.
//
//   Y[M,N] (fp32) = X[M,K] (bf16, row-major) * W[N,K]^T (bf16, row-major), fp32 accumulation in TMEM.
//
// Why: the shipped dense layers keep fp32 activations in HBM and are bound by staging them (converter warps: shared
// memory -> registers -> TMEM, DESIGN.md "Reduced-precision mode"; the converter-free SS form of the one-pass TF32 mode
// already reaches 459 TFLOP/s).  With bf16 activations written by the producing layer's epilogue both operands go
// TMA -> shared memory -> tcgen05.mma (SS form, kind::f16) with half the bytes and twice the MMA rate.  This probe
// measures that ceiling before the product path is changed (next round).
//
// Structure (same barrier protocol as csrc/gemm_tc2.cu, minus converters / chains):
//   persistent CTA per SM, tile 128 x 256, K block = 64 bf16 (one 128-byte swizzle row), 4 stages of 48 KB,
//   warp 0 = TMA producer, warp 1 = MMA issuer (4 x tcgen05.mma M128 N256 K16 per K block), warps 2-5 = epilogue
//   (tcgen05.ld -> 256-bit row stores), two 256-column TMEM accumulators so that the epilogue of tile i overlaps the
//   main loop of tile i+1.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o gemm_bf16_probe gemm_bf16_probe.cu -lcuda
// run:   ./gemm_bf16_probe            (prints max error against a host reference on sampled outputs and TFLOP/s)
// STATUS: compiled for sm_100a (UTCHMMA in the SASS) at the end of round 1 when the GPU budget was spent -- NOT YET RUN.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../normalizing-flows-study_b200/csrc/tc_common.cuh"

using namespace nf;

constexpr int BM = 128, BN = 256, BK = 64;          // BK in bf16 elements = 128 bytes
constexpr int STAGES = 4;
constexpr int THREADS = 192;
constexpr uint32_t X_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2, STAGE_BYTES = X_BYTES + W_BYTES;
constexpr int TMEM_COLS = 512;                       // two 256-column accumulators

// kind::f16 instruction descriptor: D = f32 (1 << 4), A = bf16 (1 << 7), B = bf16 (1 << 10), both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24 (same field layout as tc::idesc_tf32_m128)
__host__ __device__ constexpr uint32_t idesc_bf16_m128(uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(tc::smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, float* __restrict__ Y,
                 int M, int N, int K, int64_t ldc, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                  // [S] TMA landed
    uint64_t* empty = full + STAGES;        // [S] MMAs consumed the stage
    uint64_t* d_full = empty + STAGES;      // [2] accumulator complete
    uint64_t* d_empty = d_full + 2;         // [2] epilogue read the accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (N + BN - 1) / BN;
    const int nkb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&d_full[i], 1); tc::mbar_init(&d_empty[i], 128); }
        tc::fence_mbar_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0; bool first = true;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (!first) tc::mbar_wait(&empty[s], ph ^ 1u);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    tc::mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
                    tma_load_2d(st, &tm_x, kb * BK, m0, &full[s]);
                    tma_load_2d(st + X_BYTES, &tm_w, kb * BK, n0, &full[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1u; first = false; }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = idesc_bf16_m128((uint32_t)BN);
        const bool leader = tc::elect_one();
        int s = 0, acc = 0; uint32_t ph = 0, dph = 0; bool first_acc = true;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            if (!first_acc) tc::mbar_wait(&d_empty[acc], dph ^ 1u);
            tc::fence_after_sync();
            const uint32_t dcol = tb + acc * BN;
            for (int kb = 0; kb < nkb; ++kb) {
                tc::mbar_wait(&full[s], ph);
                tc::fence_after_sync();
                const uint32_t st = tc::smem_u32(smem + s * STAGE_BYTES);
                const uint64_t dx = tc::smem_desc_k_sw128(st), dw = tc::smem_desc_k_sw128(st + X_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {         // 16 bf16 = 32 bytes per MMA along K
                    if (leader) mma_bf16_ss(dcol, dx + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                }
                if (leader) {
                    tc::mma_commit(&empty[s]);
                    if (kb == nkb - 1) tc::mma_commit(&d_full[acc]);
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
            if (++acc == 2) { acc = 0; dph ^= 1u; first_acc = false; }
        }
    } else {
        const int q = warp & 3;
        const uint32_t lane_addr = tb + ((uint32_t)(q * 32) << 16);
        int acc = 0; uint32_t dph = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
            tc::mbar_wait(&d_full[acc], dph);
            tc::fence_after_sync();
            const int row = m0 + q * 32 + lane;
            float* yrow = Y + (int64_t)row * ldc;
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v0[16], v1[16];
                tc::tmem_ld16(lane_addr + acc * BN + c * 32, v0);
                tc::tmem_ld16(lane_addr + acc * BN + c * 32 + 16, v1);
                tc::wait_ld();
                if (c == BN / 32 - 1) { tc::fence_before_sync(); tc::mbar_arrive(&d_empty[acc]); }
                if (row < M) tc::epilogue_store32(yrow, n0 + c * 32, N, v0, v1, nullptr, 0, true);
            }
            if (++acc == 2) { acc = 0; dph ^= 1u; }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tb, TMEM_COLS);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int box_rows) {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
        fn = reinterpret_cast<EncodeFn>(p);
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__global__ void fill_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        p[i] = __float2bfloat16(((float)(h & 0xffff) / 32768.f - 1.f) * scale);
    }
}

static int run(int64_t M, int64_t N, int64_t K) {
    __nv_bfloat16 *x, *w; float* y;
    cudaMalloc(&x, (size_t)M * K * 2); cudaMalloc(&w, (size_t)N * K * 2); cudaMalloc(&y, (size_t)M * N * 4);
    fill_bf16<<<1184, 256>>>(x, (size_t)M * K, 1u, 1.f);
    fill_bf16<<<1184, 256>>>(w, (size_t)N * K, 7u, 1.f / sqrtf((float)K));
    cudaMemset(y, 0xff, (size_t)M * N * 4);
    alignas(64) CUtensorMap tx, tw;
    if (!make_map(&tx, x, M, K, BM) || !make_map(&tw, w, N, K, BN)) { printf("tensor map failed\n"); return 1; }
    const int tiles = (int)(((M + BM - 1) / BM) * ((N + BN - 1) / BN));
    const int grid = tiles < 148 ? tiles : 148;
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 256;
    cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    auto launch = [&]() { gemm_bf16_kernel<<<grid, THREADS, smem>>>(tx, tw, y, (int)M, (int)N, (int)K, N, tiles); };
    launch();
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    // sampled check against a host reference
    std::vector<__nv_bfloat16> hx((size_t)M * K), hw((size_t)N * K);
    cudaMemcpy(hx.data(), x, hx.size() * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(hw.data(), w, hw.size() * 2, cudaMemcpyDeviceToHost);
    double max_err = 0;
    for (int smp = 0; smp < 256; ++smp) {
        const int64_t r = (smp * 7919LL + (smp % 3 == 0 ? M - 1 - smp : 0)) % M, c = (smp * 104729LL + (smp % 5 == 0 ? N - 1 - smp : 0)) % N;
        double ref = 0, scale = 0;
        for (int64_t k = 0; k < K; ++k) {
            const double a = __bfloat162float(hx[r * K + k]), b = __bfloat162float(hw[c * K + k]);
            ref += a * b; scale += fabs(a * b);
        }
        float got;
        cudaMemcpy(&got, y + r * N + c, 4, cudaMemcpyDeviceToHost);
        const double e = fabs(got - ref) / (scale + 1e-30);
        if (!(e <= max_err)) max_err = e;
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        for (int i = 0; i < 4; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 4;
        if (ms < best) best = ms;
    }
    printf("[%lld x %lld x %lld] bf16 SS 128x256 tiles: %.3f ms  %.1f TFLOP/s   max |err| / sum|a||b| on 256 samples = %.3e  (%s)\n",
           (long long)M, (long long)N, (long long)K, best, 2.0 * M * N * K / best / 1e9, max_err, max_err < 1e-4 ? "ok" : "MISMATCH");
    cudaFree(x); cudaFree(w); cudaFree(y);
    return max_err < 1e-4 ? 0 : 1;       // bf16 products are exact in fp32; what remains is the TMEM accumulation (<= ~2^-25 per MMA)
}

int main() {
    int rc = 0;
    rc |= run(262144, 512, 512);       // MADE(64,512) / coupling(256,512) hidden layer (C3 / C5)
    rc |= run(4096, 22736, 1024);      // RealNVPSpline(784,16,1024) head (C4); N % 256 != 0
    rc |= run(65536, 1024, 1024);      // 4 x MAF(256,1024) hidden layer (C5)
    rc |= run(8192, 8192, 8192);       // the shape MEASURED_PEAKS.json's bf16 figure is quoted on
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return rc;
}
