// made_kernels.cu -- MADE-conditioned affine autoregressive flows.
//   nf_made_affine_forward     parallel directions: MADE.forward (made.py:136-140) + MAF.inverse
//                              (masked_autoregressive_flow.py:18-44) / IAF.forward (inverse_autoregressive_flow.py:30-63)
//   nf_ar_sequential_forward   sequential directions: MAF.forward (:46-78) / IAF.inverse (:65-103), computed
//                              incrementally -- the reference re-evaluates the full MADE D times; here every hidden
//                              unit is evaluated once, as soon as the inputs its degree allows are known.
#include "nf_common.cuh"

namespace nf {

constexpr int kSeqRows = 32;      // rows per CTA (lane = row)
constexpr int kSeqWarps = 8;

// dot over v in [0,n) of w[v] * act[v][lane]; w is warp-uniform (global, read-only path), act is [*][32] in smem
__device__ __forceinline__ float dot_bcast(const float* __restrict__ w, const float* __restrict__ act, int n, int lane) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int v = 0;
    if ((reinterpret_cast<uintptr_t>(w) & 15) == 0) {
        for (; v + 8 <= n; v += 8) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(w + v));
            const float4 b = __ldg(reinterpret_cast<const float4*>(w + v + 4));
            s0 = fmaf(a.x, act[(v + 0) * kSeqRows + lane], s0);
            s1 = fmaf(a.y, act[(v + 1) * kSeqRows + lane], s1);
            s2 = fmaf(a.z, act[(v + 2) * kSeqRows + lane], s2);
            s3 = fmaf(a.w, act[(v + 3) * kSeqRows + lane], s3);
            s0 = fmaf(b.x, act[(v + 4) * kSeqRows + lane], s0);
            s1 = fmaf(b.y, act[(v + 5) * kSeqRows + lane], s1);
            s2 = fmaf(b.z, act[(v + 6) * kSeqRows + lane], s2);
            s3 = fmaf(b.w, act[(v + 7) * kSeqRows + lane], s3);
        }
    }
    for (; v < n; ++v) s0 = fmaf(__ldg(w + v), act[v * kSeqRows + lane], s0);
    return (s0 + s1) + (s2 + s3);
}

// mode: AR_MAF_FWD (2) or AR_IAF_INV (3)
__global__ void __launch_bounds__(kSeqRows * kSeqWarps, 1)
ar_sequential_kernel(const float* __restrict__ vin, const float* __restrict__ w0, const float* __restrict__ b0,
                     const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                     const float* __restrict__ b2, const float* __restrict__ w3, const float* __restrict__ b3,
                     const int32_t* __restrict__ gstart, float* __restrict__ out, float* __restrict__ ld, int64_t B,
                     int D, int H, int mode) {
    extern __shared__ __align__(16) float smem[];
    float* sx = smem;                         // [D][32]  inputs, progressively replaced by outputs
    float* a1 = sx + (size_t)D * kSeqRows;    // [H][32]
    float* a2 = a1 + (size_t)H * kSeqRows;
    float* a3 = a2 + (size_t)H * kSeqRows;
    float* red = a3 + (size_t)H * kSeqRows;   // [kSeqWarps][2][32] partial mu/alpha
    float* sld = red + kSeqWarps * 2 * kSeqRows;  // [32] running log-det
    int* sbad = reinterpret_cast<int*>(sld + kSeqRows);   // [32] poison flag

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = (B + kSeqRows - 1) / kSeqRows;
    const float lim = (mode == AR_IAF_INV) ? 50.f : 100.f;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * kSeqRows;
        const int nrow = (int)((B - r0) < kSeqRows ? (B - r0) : kSeqRows);
        __syncthreads();
        for (int i = threadIdx.x; i < kSeqRows * D; i += blockDim.x) {     // coalesced tile load, transposed into sx
            const int r = i / D, d = i - r * D;
            sx[d * kSeqRows + r] = (r < nrow) ? vin[r0 * D + i] : 0.f;
        }
        if (threadIdx.x < kSeqRows) { sld[threadIdx.x] = 0.f; sbad[threadIdx.x] = 0; }
        __syncthreads();

        for (int g = 0; g < D; ++g) {
            // (A) parameters of dim g from a3 units of degree < g
            const int nprev = gstart[g];
            {
                const int per = (nprev + kSeqWarps - 1) / kSeqWarps;
                const int per4 = (per + 3) & ~3;                       // keep 16B alignment of the weight slices
                const int v0 = warp * per4 < nprev ? warp * per4 : nprev;
                const int v1 = (v0 + per4) < nprev ? (v0 + per4) : nprev;
                const float pm = dot_bcast(w3 + (size_t)g * H + v0, a3 + (size_t)v0 * kSeqRows, v1 - v0, lane);
                const float pa = dot_bcast(w3 + (size_t)(D + g) * H + v0, a3 + (size_t)v0 * kSeqRows, v1 - v0, lane);
                red[(warp * 2 + 0) * kSeqRows + lane] = pm;
                red[(warp * 2 + 1) * kSeqRows + lane] = pa;
            }
            __syncthreads();
            if (warp == 0) {
                float mu = __ldg(b3 + g), al = __ldg(b3 + D + g);
#pragma unroll
                for (int w = 0; w < kSeqWarps; ++w) { mu += red[(w * 2) * kSeqRows + lane]; al += red[(w * 2 + 1) * kSeqRows + lane]; }
                float o, t;
                affine_ar_elem<float>(mode, sx[g * kSeqRows + lane], mu, al, o, t);
                if (sbad[lane]) { o = __int_as_float(0x7fc00000); t = o; }
                if (!is_finite(o)) sbad[lane] = 1;      // 0*NaN of the dense reference poisons every later dim
                sx[g * kSeqRows + lane] = o;
                sld[lane] += t;
            }
            __syncthreads();
            if (g == D - 1) break;
            // (B) hidden units whose degree is g
            const int u0 = gstart[g], u1 = gstart[g + 1];
            for (int u = u0 + warp; u < u1; u += kSeqWarps) {
                float s = __ldg(b0 + u);
                for (int j = 0; j <= g; ++j) s = fmaf(__ldg(w0 + (size_t)u * D + j), sx[j * kSeqRows + lane], s);
                a1[u * kSeqRows + lane] = relu_nan(s);
            }
            __syncthreads();
            for (int u = u0 + warp; u < u1; u += kSeqWarps)
                a2[u * kSeqRows + lane] = relu_nan(__ldg(b1 + u) + dot_bcast(w1 + (size_t)u * H, a1, u1, lane));
            __syncthreads();
            for (int u = u0 + warp; u < u1; u += kSeqWarps)
                a3[u * kSeqRows + lane] = relu_nan(__ldg(b2 + u) + dot_bcast(w2 + (size_t)u * H, a2, u1, lane));
            __syncthreads();
        }
        // epilogue: scrubs + coalesced store
        for (int i = threadIdx.x; i < nrow * D; i += blockDim.x) {
            const int r = i / D, d = i - r * D;
            float o = sx[d * kSeqRows + r];
            if (!is_finite(o)) o = (mode == AR_IAF_INV) ? vin[r0 * D + i] : 0.f;
            out[r0 * D + i] = o;
        }
        if (threadIdx.x < nrow) ld[r0 + threadIdx.x] = clamp_mm(scrub0(sld[threadIdx.x]), -lim, lim);
    }
}

}  // namespace nf

using namespace nf;
#define NF_REQ(p) do { if ((p) == nullptr) return NF_ERR_NULL; } while (0)

extern "C" int nf_made_affine_forward(const void* v, const void* w0, const void* b0, const void* w1, const void* b1,
                                      const void* w2, const void* b2, const void* w3, const void* b3,
                                      const int32_t* kext1, const int32_t* kext2, const int32_t* kext3,
                                      void* workspace, void* out, void* ld, int64_t B, int D, int H, int mode,
                                      int dtype, nf_stream_t stream) {
    if (B < 0 || D < 1 || H < 1) return NF_ERR_BAD_SHAPE;
    if (mode != NF_AR_MAF_INVERSE && mode != NF_AR_IAF_FORWARD) return NF_ERR_UNSUPPORTED;
    if (dtype != NF_F32 && dtype != NF_F64) return NF_ERR_UNSUPPORTED;
    if (B == 0) return NF_OK;
    NF_REQ(v); NF_REQ(w0); NF_REQ(w1); NF_REQ(w2); NF_REQ(w3); NF_REQ(workspace); NF_REQ(out); NF_REQ(ld);
    const size_t es = dtype == NF_F64 ? 8 : 4;
    // workspace: two ping-pong activation buffers of B*max(H,2D) elements; the [B,2D] parameter block lands in the second
    const size_t wcols = (size_t)(H > 2 * D ? H : 2 * D);
    char* ha = (char*)workspace;
    char* hb = ha + es * (size_t)B * wcols;
    int rc;
    rc = nf_gemm(v, w0, ha, b0, B, H, D, D, 1, 1, D, H, 1, 0, nullptr, dtype, stream);          if (rc) return rc;
    rc = nf_gemm(ha, w1, hb, b1, B, H, H, H, 1, 1, H, H, 1, 0, kext1, dtype, stream);           if (rc) return rc;
    rc = nf_gemm(hb, w2, ha, b2, B, H, H, H, 1, 1, H, H, 1, 0, kext2, dtype, stream);           if (rc) return rc;
    rc = nf_gemm(ha, w3, hb, b3, B, 2 * D, H, H, 1, 1, H, 2 * D, 0, 0, kext3, dtype, stream);   if (rc) return rc;
    return nf_affine_ar_forward(v, hb, out, ld, B, D, mode, dtype, stream);
}

extern "C" int nf_ar_sequential_forward(const void* v, const void* w0, const void* b0, const void* w1, const void* b1,
                                        const void* w2, const void* b2, const void* w3, const void* b3,
                                        const int32_t* gstart, void* out, void* ld, int64_t B, int D, int H, int mode,
                                        nf_stream_t stream) {
    if (B < 0 || D < 1 || H < 1) return NF_ERR_BAD_SHAPE;
    if (mode != NF_AR_MAF_FORWARD && mode != NF_AR_IAF_INVERSE) return NF_ERR_UNSUPPORTED;
    if (B == 0) return NF_OK;
    NF_REQ(v); NF_REQ(w0); NF_REQ(b0); NF_REQ(w1); NF_REQ(b1); NF_REQ(w2); NF_REQ(b2); NF_REQ(w3); NF_REQ(b3);
    NF_REQ(gstart); NF_REQ(out); NF_REQ(ld);
    const size_t smem = sizeof(float) * ((size_t)(D + 3 * (size_t)H) * kSeqRows + kSeqWarps * 2 * kSeqRows + 2 * kSeqRows);
    if (smem > 227 * 1024) return NF_ERR_UNSUPPORTED;   // caller falls back to D dense passes (reference algorithm)
    cudaStream_t st = (cudaStream_t)stream;
    NF_CUDA(cudaFuncSetAttribute(ar_sequential_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NF_CUDA(cudaFuncSetAttribute(ar_sequential_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int64_t ntiles = cdiv(B, kSeqRows);
    int per_sm = 1;
    NF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_sequential_kernel, kSeqRows * kSeqWarps, smem));
    if (per_sm < 1) return NF_ERR_UNSUPPORTED;
    const int64_t cap = (int64_t)kNumSMs * per_sm;
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    ar_sequential_kernel<<<grid, kSeqRows * kSeqWarps, smem, st>>>(
        (const float*)v, (const float*)w0, (const float*)b0, (const float*)w1, (const float*)b1, (const float*)w2,
        (const float*)b2, (const float*)w3, (const float*)b3, gstart, (float*)out, (float*)ld, B, D, H, mode);
    count_launch();
    NF_LAUNCH_CHECK();
    return NF_OK;
}
