// Probe: latency and throughput of the register-level mma.sync.m16n8k8 TF32 path on sm_100a (is it fast enough for the
// dependent in-block steps of the blocked sequential direction, where N = 8 units and K = 8..72 per step?).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_tmp/mma_sync_probe scripts/dbg/mma_sync_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int CH>
__global__ void probe(float* out, long long* cyc, int iters) {
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
    float c[CH][4];
    for (int k = 0; k < CH; ++k) for (int i = 0; i < 4; ++i) c[k][i] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < CH; ++k) mma_tf32(c[k], a, b);
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int k = 0; k < CH; ++k) for (int i = 0; i < 4; ++i) s += c[k][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void run(int warps, int iters) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<CH><<<148, warps * 32>>>(out, cyc, iters);
    cudaEventRecord(e0);
    probe<CH><<<148, warps * 32>>>(out, cyc, iters);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double mmas = 148.0 * warps * iters * CH;
    printf("chains=%d warps/SM=%d: %.1f cycles per iteration (%.2f per MMA per warp), %.3f ms, %.1f TFLOP/s tf32 (m16n8k8), %.2f MMA/clk/SM\n",
           CH, warps, (double)h / iters, (double)h / iters / CH, ms, mmas * 2048 / (ms * 1e-3) / 1e12, (double)warps * iters * CH / h);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(1, 4096); run<2>(1, 4096); run<4>(1, 4096); run<8>(1, 4096);
    run<1>(4, 4096); run<4>(4, 4096); run<8>(4, 4096);
    run<4>(8, 4096); run<8>(8, 4096); run<6>(16, 4096); run<8>(16, 2048); run<4>(32, 2048);
    cudaError_t e = cudaGetLastError(); printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
