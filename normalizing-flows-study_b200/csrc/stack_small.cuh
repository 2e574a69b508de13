// stack_small.cuh -- packed-weight layout of the fused small-data_dim inference stacks.
// Shared by stack_small.cu (device) and mirrored by the Python packer (nfb200/packing.py), which asks
// nf_*_stack_packed_floats() for sizes and fills the buffer following the offsets documented here.
//
// All words are 32-bit; integers are stored as int32 bit patterns inside the float buffer.
//
//   header (NF_STACK_HDR words):
//     [0] magic   [1] D   [2] H (true hidden_dim)   [3] HP (padded: 64 or 128)   [4] K (spline) / 0 (affine)
//     [5] L       [6] W1S (words per hidden unit in W1k: 4 if D<=3 else 12)      [7] NO (padded #outputs of the head)
//     [8] layer_stride (words)   [9] bn_between (0/1)   [10] bound   [11] min_w  [12] min_h  [13] min_d
//     [14] scale_w = 1-min_w*K   [15] scale_h
//   per layer (layer_stride words, 16-byte aligned sections):
//     mask[8] | tdim[8] (int) | meta[8]: {Dt, rescale, bn_on, bn_ld(float), 0..} | r_in[8] | r_lo[8] | r_out[8]
//     | bn_mean[8] | bn_sd[8] | bn_gamma[8] | bn_beta[8]                      (NF_LAYER_HDR = 80 words)
//     then `nets` conditioner blocks (1 for spline: param_net; 2 for affine: s_net, b_net), each:
//       W1k [HP][W1S]   : {W1[k][0..D-1], zero pad.., b1[k] at slot W1S-1}
//       W2t [HP][HP]    : W2t[k][j] = W2[j][k]
//       b2  [HP]
//       W3c [NO/4][HP][4]: W3c[c][j][q] = W3[out(c*4+q)][j]
//       b3  [NO]
//   spline head outputs are ordered t*P + p for transformed dim t (P = 3K-1); affine head outputs are d.
#pragma once
#include <stdint.h>

#define NF_STACK_MAGIC_SPLINE 0x4e465331  /* 'NFS1' */
#define NF_STACK_MAGIC_AFFINE 0x4e464131  /* 'NFA1' */
#define NF_STACK_MAGIC_SPLINE_TC 0x4e465332  /* 'NFS2': tensor-core layout, see stack_tc.cu */
#define NF_STACK_MAGIC_AFFINE_TC 0x4e464132  /* 'NFA2' */
#define NF_STACK_MAGIC_MADE_TC 0x4e464d32    /* 'NFM2': MADE / MAF / IAF stack on the tensor cores, see stack_tc.cu */
#define NF_STACK_HDR 16
#define NF_LAYER_HDR 80
#define NF_STACK_DMAX 8
// `inverse` argument of the fused stack entry points is a flag word:
//   bit 0  direction (1 = inverse, x -> z)
//   bit 1  log-prob head: the per-row output is log N(z; 0, I) + log_det (Flow.log_prob, flow.py:56-73) instead of
//          log_det -- the standard-normal head evaluated on the row while it is still in registers
//   bit 2  do not store the transformed rows (y may be NULL): Flow.log_prob only returns the per-row value
#define NF_STACK_INVERSE 1
#define NF_STACK_LOG_PROB_HEAD 2
#define NF_STACK_SKIP_Y 4

static inline int nf_stack_hp(int H) { return H <= 64 ? 64 : 128; }
static inline int nf_stack_w1s(int D) { return D <= 3 ? 4 : 12; }
static inline int64_t nf_stack_net_words(int HP, int W1S, int NO) {
    return (int64_t)HP * W1S + (int64_t)HP * HP + HP + (int64_t)NO * HP + NO;
}

#ifdef __CUDACC__
// log N(z; 0, I) + log_det for one row held in registers (same arithmetic as std_normal_log_prob_fwd_kernel)
template <int DM>
__device__ __forceinline__ float nf_stack_row_head(const float (&z)[DM], int D, float log_det) {
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < DM; ++d) if (d < D) acc += -0.5f * z[d] * z[d];
    return acc - (float)(0.5 * (double)D * 1.8378770664093453) + log_det;
}
#endif
