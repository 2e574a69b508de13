// umma_probe.cu -- which shared-memory words does a tcgen05.mma (kind::tf32, A in TMEM) read for a given B descriptor?
// A[m][k] = (k == kk0), smem word i holds (i % 32) [mode 0] or the 128-byte row index i / 32 [mode 1], so
// D[m][n] identifies the word read as B(n, kk0).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "../../normalizing-flows-study_b200/csrc/tc_common.cuh"
using namespace nf;

__global__ void __launch_bounds__(128) probe(float* D, uint64_t desc_fields, uint32_t idesc, int kk0, int N, int mode, int ss) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* w = reinterpret_cast<float*>(smem);
    for (int i = tid; i < 16384; i += 128) w[i] = mode == 0 ? (float)((i & 31) + 1) : (float)((i >> 5) + 1);
    // SS form: A tile [128 x 32] K-major SWIZZLE_128B at byte 32768, A[m][k] = (k == kk0)
    for (int i = tid; i < 4096; i += 128) w[8192 + i] = 0.f;
    __syncthreads();
    w[8192 + tc::sw128_offset((uint32_t)tid, (uint32_t)kk0, 128) / 4] = 1.0f;
    tc::fence_proxy_async_smem();
    if (warp == 0) tc::tmem_alloc(&slot, 512);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tb = slot;
    const uint32_t lane_addr = tb + ((uint32_t)(warp * 32) << 16);
    uint32_t a[16];
    for (int j = 0; j < 16; ++j) a[j] = (j == kk0) ? __float_as_uint(1.0f) : 0u;
    tc::tmem_st16(lane_addr + 0, a);
    uint32_t z[16];
    for (int j = 0; j < 16; ++j) z[j] = __float_as_uint(-7.0f);
    for (int c = 0; c < 16; ++c) tc::tmem_st16(lane_addr + 128 + c * 16, z);      // D pre-filled with -7: untouched output is visible
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        tc::fence_after_sync();
        const uint64_t desc = desc_fields | (uint64_t)((tc::smem_u32(smem) >> 4) & 0x3fffu);
        const bool leader = tc::elect_one();
        if (leader) {
            if (ss) tc::mma_tf32_ss(tb + 128, tc::smem_desc_k_sw128(tc::smem_u32(smem) + 32768), desc, idesc, 0u);
            else tc::mma_tf32_ts(tb + 128, tb + 0, desc, idesc, 0u);
            tc::mma_commit(&bar);
        }
        __syncwarp();
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c = 0; c < N / 16; ++c) {
        uint32_t v[16];
        tc::tmem_ld16(lane_addr + 128 + c * 16, v);
        tc::wait_ld();
        for (int j = 0; j < 16; ++j) D[tid * 256 + c * 16 + j] = __uint_as_float(v[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 512);
}

int main() {
    float* D; cudaMalloc(&D, 128 * 256 * 4);
    std::vector<float> h0(128 * 256), h1(128 * 256);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    struct Cfg { const char* name; int transpose; uint32_t lbo, sbo; int layout; int N; int ss; };
    Cfg cfgs[] = {
        {"TS MN sw128_base32 lbo4096 sbo512", 1, 4096, 512, 1, 128, 0},
        {"TS MN sw128_base32 lbo4096 sbo1024", 1, 4096, 1024, 1, 128, 0},
        {"SS MN sw128_base32 lbo4096 sbo512", 1, 4096, 512, 1, 128, 1},
    };
    for (auto& c : cfgs) {
        uint64_t f = ((uint64_t)((c.lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((c.sbo >> 4) & 0x3fff) << 32) | (1ull << 46) | ((uint64_t)c.layout << 61);
        uint32_t idesc = tc::idesc_tf32_m128((uint32_t)c.N) | (c.transpose == 1 ? (1u << 16) : c.transpose == 2 ? (1u << 15) : 0u);
        printf("=== %s  idesc=%08x\n", c.name, idesc);
        for (int kk0 : {0, 1, 2, 3, 4, 5, 7}) {
            probe<<<1, 128, 65536>>>(D, f, idesc, kk0, c.N, 0, c.ss); cudaMemcpy(h0.data(), D, h0.size() * 4, cudaMemcpyDeviceToHost);
            probe<<<1, 128, 65536>>>(D, f, idesc, kk0, c.N, 1, c.ss); cudaMemcpy(h1.data(), D, h1.size() * 4, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            printf(" kk0=%d (row:word) n=0..:", kk0);
            for (int n = 0; n < c.N; ++n) { if (n < 34 || (n >= 30 && n < 36) || (n >= 62 && n < 66) || n == c.N - 1) printf(" %g:%g", h1[n], h0[n]); else if (n == 34 || n == 36 || n == 66) printf(" .."); }
            printf("   | row 77:"); for (int n = 0; n < 4; ++n) printf(" %g:%g", h1[77 * 256 + n], h0[77 * 256 + n]);
            printf("\n");
        }
    }
    return 0;
}
