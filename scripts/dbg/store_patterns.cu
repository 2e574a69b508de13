// Store-pattern probe: 262144 rows, `cols` floats written per row at a row pitch of `ld` floats.
//   mode 0: thread = row, one st.global.v8.f32 (32 B) per 8 columns           (tensor-core epilogue, round 1u)
//   mode 1: thread = row, two st.global.v4.f32 per 8 columns
//   mode 2: warp = row group, lanes along columns: 128-byte coalesced lines   (transposed epilogue, round 1r)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_row(float* y, int rows, int cols, int ld, int mode) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    float* p = y + (size_t)row * ld;
    const float v = (float)row;
    for (int c = 0; c < cols; c += 8) {
        if (mode == 0)
            asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + c), "f"(v));
        else {
            asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p + c), "f"(v));
            asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p + c + 4), "f"(v));
        }
    }
}
__global__ void k_line(float* y, int rows, int cols, int ld) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int r0 = warp * 32;
    for (int c = 0; c < cols; c += 32)
        for (int rr = 0; rr < 32; ++rr) {
            const int row = r0 + rr;
            if (row < rows && c + lane < cols) y[(size_t)row * ld + c + lane] = (float)row;
        }
}
int main() {
    const int rows = 262144;
    float* y;
    cudaMalloc(&y, (size_t)rows * 512 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int shapes[][2] = {{64, 512}, {512, 512}, {8, 128}, {64, 64}, {128, 512}};
    for (auto& s : shapes) {
        const int cols = s[0], ld = s[1];
        for (int mode = 0; mode < 3; ++mode) {
            float best = 1e9f;
            for (int it = 0; it < 5; ++it) {
                cudaEventRecord(e0);
                if (mode < 2) k_row<<<rows / 128, 128>>>(y, rows, cols, ld, mode);
                else k_line<<<rows / 128, 128>>>(y, rows, cols, ld);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (it > 0 && ms < best) best = ms;
            }
            printf("cols=%3d ld=%3d mode=%d (%s): %.3f ms  %.2f TB/s\n", cols, ld, mode,
                   mode == 0 ? "row v8" : mode == 1 ? "row 2xv4" : "coalesced lines", best, (double)rows * cols * 4 / best / 1e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
