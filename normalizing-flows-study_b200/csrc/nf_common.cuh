// nf_common.cuh -- device/host utilities shared by the kernels of libnfb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nfb200.h"
#include "nf_math.cuh"

namespace nf {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

#define NF_LAUNCH_CHECK()                                             \
    do {                                                              \
        cudaError_t e__ = cudaGetLastError();                         \
        if (e__ != cudaSuccess) return nf_set_cuda_error(e__);        \
    } while (0)

#define NF_CUDA(call)                                                 \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return nf_set_cuda_error(e__);        \
    } while (0)

int nf_set_cuda_error(cudaError_t e);   // records the last CUDA error string; returns NF_ERR_CUDA
void count_launch();                    // bumps nf_launch_count()
void count_launches(int n);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over aligned groups of G lanes (G power of two <= 32)
template <typename T, int G>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming loads/stores: data touched exactly once, keep it out of L1
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <typename T>
__host__ inline RqsCfg<T> make_rqs_cfg(bool bounded, int K, double bound, double min_w, double min_h, double min_d) {
    RqsCfg<T> c;
    if (bounded) { c.lo = (T)(-bound); c.hi = (T)bound; c.span = (T)(2.0 * bound); c.eps = (T)1e-8; }
    else         { c.lo = (T)0; c.hi = (T)1; c.span = (T)1; c.eps = (T)1e-6; }
    c.min_w = (T)min_w; c.min_h = (T)min_h; c.min_d = (T)min_d;
    c.scale_w = (T)(1.0 - min_w * K);     // python-double scalar, then cast (spline_coupling_layer.py:205)
    c.scale_h = (T)(1.0 - min_h * K);
    return c;
}

// spline_stream.cu: TMA-staged float32 kernels of the compact spline coupling transform (K = 8 / 10).  Returns
// NF_ERR_UNSUPPORTED when the shape is outside its envelope (the caller then takes the first-version kernels).
struct SplineStreamArgs {
    const float* x; const float* params; const float* mask; const int32_t* tidx;
    float* y; float* ld;                                              // forward outputs
    const float* gy; const float* gld; float* gx; float* gparams;     // backward
    int64_t B; int D, Dt;
    RqsCfg<float> c;
    const float* r_in; const float* r_lo; const float* r_out;
};
int spline_stream_launch(const SplineStreamArgs& a, int K, bool bwd, int inverse, cudaStream_t st);

}  // namespace nf
