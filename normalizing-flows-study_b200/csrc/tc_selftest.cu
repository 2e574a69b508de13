// tc_selftest.cu -- unit-test hook for the tcgen05 tile primitive used by the fused stacks (tc_common.cuh):
//   D[128][N] = A[128][64] * W[N][64]^T  with A split hi/lo into TMEM by its owning threads, W pre-packed on the
//   host as two K-major SWIZZLE_128B images (hi, lo), 3xTF32 accumulation in TMEM, result read back with tcgen05.ld.
// tests/test_gpu_tensorcore.py compares it with a float64 product.
#include "nf_common.cuh"
#include "tc_common.cuh"

namespace nf {

__global__ void __launch_bounds__(128)
tc_gemm128_kernel(const float* __restrict__ A, const float* __restrict__ Wimg, float* __restrict__ D, int N, int passes,
                  long long* __restrict__ timing, int nacc) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float* sW = reinterpret_cast<float*>(smem_raw);          // hi image [N*64] then lo image [N*64]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid * 4; i < 2 * N * 64; i += 128 * 4) cp_async16(sW + i, Wimg + i);
    cp_async_commit();
    cp_async_wait<0>();
    tc::fence_proxy_async_smem();
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tb = tmem_base_s;
    const uint32_t lane_addr = tb + ((uint32_t)(warp * 32) << 16);

#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) tc::split_tf32(A[tid * 64 + c * 16 + j], hi[j], lo[j]);
        tc::tmem_st16(lane_addr + 0 + c * 16, hi);
        tc::tmem_st16(lane_addr + 64 + c * 16, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    long long t0 = clock64(), t1 = t0;
    if (warp == 0) {
        tc::fence_after_sync();
        const uint32_t w_hi = tc::smem_u32(sW), w_lo = tc::smem_u32(sW + N * 64);
        if (passes == 3 && nacc == 1) {
            tc::warp_issue_gemm_k64_3xtf32(tb, 128, 0, 64, w_hi, w_lo, (uint32_t)N, &bar);       // the product path
        } else {
            // timing / single-pass variants: extra passes repeat pass 0; nacc > 1 rotates the accumulator block so that
            // consecutive MMAs are independent (the checked result is accumulator 0 with nacc == 1)
            const uint32_t idesc = tc::idesc_tf32_m128((uint32_t)N);
            const uint64_t d_hi = tc::smem_desc_k_sw128(w_hi), d_lo = tc::smem_desc_k_sw128(w_lo);
            const uint32_t atom16 = ((uint32_t)N * 128u) >> 4;
            const bool leader = tc::elect_one();
            uint32_t acc = 0;
            for (int pass = 0; pass < passes; ++pass) {
                const uint32_t a_col = (pass == 1) ? 64u : 0u;
                const uint64_t wd = (pass == 2) ? d_lo : d_hi;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t bd = wd + (uint64_t)((uint32_t)(k >> 2) * atom16 + (uint32_t)(k & 3) * 2u);
                    if (leader) tc::mma_tf32_ts(tb + 128 + acc * 128u, tb + a_col + k * 8, bd, idesc, (pass | k) != 0 ? 1u : 0u);
                    acc = (acc + 1 == (uint32_t)nacc) ? 0u : acc + 1;
                }
            }
            if (leader) tc::mma_commit(&bar);
            __syncwarp();
        }
        t1 = clock64();
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    if (tid == 0 && timing) { timing[0] = t1 - t0; timing[1] = clock64() - t0; }
    for (int c = 0; c < N / 16; ++c) {
        uint32_t v[16];
        tc::tmem_ld16(lane_addr + 128 + c * 16, v);
        tc::wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) D[tid * N + c * 16 + j] = __uint_as_float(v[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 512);
}

}  // namespace nf

using namespace nf;

extern "C" int nf_debug_tc_gemm128(const void* a, const void* w_images, void* d, int N, int passes, void* timing, int nacc,
                                   nf_stream_t stream) {
    if (N < 16 || N > 128 || N % 16 != 0 || passes < 1 || passes > 64 || nacc < 1 || nacc > 3) return NF_ERR_BAD_SHAPE;
    if (!a || !w_images || !d) return NF_ERR_NULL;
    if (!aligned16(w_images)) return NF_ERR_MISALIGNED;
    const size_t smem = (size_t)2 * N * 64 * sizeof(float);
    NF_CUDA(cudaFuncSetAttribute(tc_gemm128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm128_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const float*)a, (const float*)w_images, (float*)d, N, passes,
                                                                (long long*)timing, nacc);
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}
