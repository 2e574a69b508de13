// nf_math.cuh -- scalar math of the flow transforms, shared by every kernel.
//
// Every function is __host__ __device__ so that tests/hostcheck can compile the *same* source
// for the CPU and compare it with the oracle / torch autograd before any GPU time is spent.
// The product library (libnfb200.so) only instantiates the __device__ side.
//
// Semantics follow the reference (paths relative to the reference repo root):
//   * torch.clamp / torch.relu propagate NaN  -> clamp_* / relu_nan below never use fmin/fmax
//   * NaN/Inf scrubs after every transform     (coupling_layer.py:61-66, spline_coupling_layer.py:306-307,
//                                               masked_autoregressive_flow.py:35-42, inverse_autoregressive_flow.py:53-61)
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NF_HD __host__ __device__ __forceinline__
#else
#define NF_HD inline
#endif

// Loop unrolling of the per-bin loops.  The register kernels unroll fully (arrays stay in registers);
// transform_generic.cu compiles the same source with `unroll 1` for the rarely used num_bins>16 path.
#ifndef NF_UNROLL
#define NF_UNROLL _Pragma("unroll")
#endif

namespace nf {

// ------------------------------------------------------------------------------------------
// type-generic libm
// ------------------------------------------------------------------------------------------
NF_HD float  t_exp(float x)    { return expf(x); }
NF_HD double t_exp(double x)   { return exp(x); }
NF_HD float  t_log(float x)    { return logf(x); }
NF_HD double t_log(double x)   { return log(x); }
NF_HD float  t_log1p(float x)  { return log1pf(x); }
NF_HD double t_log1p(double x) { return log1p(x); }
NF_HD float  t_sqrt(float x)   { return sqrtf(x); }
NF_HD double t_sqrt(double x)  { return sqrt(x); }
NF_HD float  t_abs(float x)    { return fabsf(x); }
NF_HD double t_abs(double x)   { return fabs(x); }

// exp() of the softmax terms (argument <= 0 after the max subtraction).  Device float: exact range reduction
// (n = rint(x*log2e), f = x*log2e - n through an FMA) + ex2.approx on |f| <= 0.5 + exponent add: 8 instructions,
// relative error <= 2^-22 + 1 ulp (the libm expf costs 22).  NaN propagates; arguments below -87 give e^-87.
NF_HD double sm_exp(double x) { return exp(x); }
NF_HD float sm_exp(float x) {
#if defined(__CUDA_ARCH__)
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(x) : "f"(x), "f"(-87.0f));
    const float magic = 12582912.0f;                                  // 1.5 * 2^23: integer part lands in the low mantissa bits
    const float tm = fmaf(x, 1.4426950408889634f, magic);
    const float n = tm - magic;
    float f = fmaf(x, 1.4426950408889634f, -n);                       // fractional part, one rounding
    f = fmaf(x, 1.9259629911e-8f, f);                                 // low word of log2(e)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f));
    return __int_as_float(__float_as_int(r) + (__float_as_int(tm) << 23));
#else
    return expf(x);
#endif
}
#if defined(__CUDA_ARCH__)
// two exponentials through one packed-fp32 instruction stream (FFMA2 / FADD2; ex2 and the exponent add stay scalar):
// the same operations on each half as sm_exp(float), bit-identical results, 10 instructions per pair instead of 16
__device__ __forceinline__ float2 sm_exp_x2(float2 x) {
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(x.x) : "f"(x.x), "f"(-87.0f));
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(x.y) : "f"(x.y), "f"(-87.0f));
    const float magic = 12582912.0f;
    const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
    const float2 tm = __ffma2_rn(x, l2e, make_float2(magic, magic));
    const float2 nn = __ffma2_rn(tm, make_float2(-1.0f, -1.0f), make_float2(magic, magic));     // -(tm - magic), exact
    float2 f = __ffma2_rn(x, l2e, nn);
    f = __ffma2_rn(x, make_float2(1.9259629911e-8f, 1.9259629911e-8f), f);
    float2 r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(f.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(f.y));
    r.x = __int_as_float(__float_as_int(r.x) + (__float_as_int(tm.x) << 23));
    r.y = __int_as_float(__float_as_int(r.y) + (__float_as_int(tm.y) << 23));
    return r;
}
#endif
// reciprocal used to normalise the softmax (one division per row of bins instead of one per bin)
// Device float: rcp.approx + one Newton step (<= 1 ulp; the argument is a sum of K exponentials in [1, K]) -- the
// IEEE division 1.0f / x costs ten instructions with its range check.
NF_HD float t_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
    return 1.0f / x;
#endif
}
NF_HD double t_rcp(double x) { return 1.0 / x; }
// division / logarithm inside the spline bin evaluation.  Device float: a * rcp.approx(b) (2 instructions, <= 2 ulp;
// __fdividef's subnormal handling costs six) and lg2.approx
// (absolute error < 4e-7 on the log-det terms, tolerance 1e-4); both cut the dependent-instruction chain of the
// per-row spline from ~40 to ~10 cycles per operation.  Host and double: exact.
NF_HD double t_div(double a, double b) { return a / b; }
NF_HD float t_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    float r;                          // every denominator on this path is clamped to >= eps: no subnormals to scale
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return a * r;
#else
    return a / b;
#endif
}
NF_HD double t_logf(double x) { return log(x); }
NF_HD float t_logf(float x) {
#if defined(__CUDA_ARCH__)
    return __logf(x);
#else
    return logf(x);
#endif
}

template <typename T> NF_HD bool is_finite(T x) { return (x - x) == T(0); }       // false for NaN and +-Inf
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ bool is_finite(float x) { return fabsf(x) < __int_as_float(0x7f800000); }   // one FSETP
#endif
template <typename T> NF_HD T clamp_min(T x, T lo) { return x < lo ? lo : x; }    // NaN stays NaN (torch.clamp)
template <typename T> NF_HD T clamp_mm(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }
#if defined(__CUDA_ARCH__)
// device float: the NaN-propagating min / max of the FMNMX unit, one instruction per bound instead of a compare + select
// pair (same value for every input; the only visible difference is clamp(-0.0, 0, .) = +0.0 instead of -0.0)
__device__ __forceinline__ float clamp_min(float x, float lo) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(lo));
    return r;
}
__device__ __forceinline__ float clamp_mm(float x, float lo, float hi) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(lo));
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(hi));
    return r;
}
#endif
template <typename T> NF_HD T relu_nan(T x) { return x < T(0) ? T(0) : x; }       // torch.relu(NaN) = NaN
template <typename T> NF_HD bool pass_min(T x, T lo) { return x >= lo; }          // clamp backward mask
template <typename T> NF_HD bool pass_mm(T x, T lo, T hi) { return x >= lo && x <= hi; }
template <typename T> NF_HD T scrub0(T x) { return is_finite(x) ? x : T(0); }
// F.softplus, beta=1, threshold=20
template <typename T> NF_HD T softplus(T x) { return x > T(20) ? x : t_log1p(sm_exp(x)); }
template <typename T> NF_HD void sm_exp_pair(T a, T b, T& za, T& zb) { za = sm_exp(a); zb = sm_exp(b); }
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void sm_exp_pair(float a, float b, float& za, float& zb) {
    const float2 z = sm_exp_x2(make_float2(a, b));
    za = z.x; zb = z.y;
}
#endif
// min_d + softplus of the two derivative parameters of a bin (knots k and k+1) -- device float: the exponentials paired
template <typename T> NF_HD void softplus_x2(T a, T b, T& sa, T& sb) { sa = softplus(a); sb = softplus(b); }
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void softplus_x2(float a, float b, float& sa, float& sb) {
    const float2 z = sm_exp_x2(make_float2(a, b));
    sa = a > 20.f ? a : log1pf(z.x);
    sb = b > 20.f ? b : log1pf(z.y);
}
#endif
template <typename T> NF_HD T softplus_grad(T x) {
    if (x > T(20)) return T(1);
    T z = t_exp(x);
    return z / (z + T(1));
}

// ------------------------------------------------------------------------------------------
// a1/a2  affine coupling, one element          (src/flows/coupling/coupling_layer.py:47-58,76-86)
//   m is the float mask value of this column; s_raw/b_raw are the un-clamped net outputs.
// ------------------------------------------------------------------------------------------
template <typename T>
NF_HD void affine_coupling_elem(T x, T m, T s_raw, T b_raw, bool inverse, T& y, T& ld_term) {
    const T s = clamp_mm(s_raw, T(-10), T(10));
    const T b = clamp_mm(b_raw, T(-10), T(10));
    const T xa = x * m;
    const T om = T(1) - m;
    if (!inverse) {
        y = xa + om * (x * t_exp(s) + b);
        ld_term = om * s;
    } else {
        y = xa + om * ((x - b) * t_exp(-s));
        ld_term = om * (-s);
    }
}

// reverse mode of the above (gy already zeroed where the output was scrubbed, gld likewise)
template <typename T>
NF_HD void affine_coupling_elem_bwd(T x, T m, T s_raw, T b_raw, bool inverse, T gy, T gld,
                                    T& gx, T& gs_raw, T& gb_raw) {
    const T s = clamp_mm(s_raw, T(-10), T(10));
    const T b = clamp_mm(b_raw, T(-10), T(10));
    const T om = T(1) - m;
    const bool ps = pass_mm(s_raw, T(-10), T(10)), pb = pass_mm(b_raw, T(-10), T(10));
    if (!inverse) {
        const T e = t_exp(s);
        gx = gy * (m + om * e);
        gs_raw = ps ? (gy * om * x * e + gld * om) : T(0);
        gb_raw = pb ? gy * om : T(0);
    } else {
        const T e = t_exp(-s);
        gx = gy * (m + om * e);
        gs_raw = ps ? (-(gy * om * (x - b) * e) - gld * om) : T(0);
        gb_raw = pb ? -(gy * om * e) : T(0);
    }
}

// ------------------------------------------------------------------------------------------
// a11/a13  affine autoregressive transform, one element (parallel directions)
//   mode 0: MAF.inverse  (masked_autoregressive_flow.py:24-33)  z=(x-mu)exp(clamp(-clamp(a,3),5)), ld-=a
//   mode 1: IAF.forward  (inverse_autoregressive_flow.py:36-50)  x=z exp(clamp(clamp(a,2),3))+clamp(mu,10), ld+=a
//   mode 2: MAF.forward step (:59-67)                            x=z exp(clamp(clamp(a,3),5))+mu, ld+=a
//   mode 3: IAF.inverse step (:80-91)                            z=(x-clamp(mu,10)) exp(clamp(-clamp(a,2),3)), ld-=a
// ------------------------------------------------------------------------------------------
enum { AR_MAF_INV = 0, AR_IAF_FWD = 1, AR_MAF_FWD = 2, AR_IAF_INV = 3 };

template <typename T>
NF_HD void affine_ar_elem(int mode, T v, T mu_raw, T al_raw, T& out, T& ld_term) {
    const bool iaf = (mode == AR_IAF_FWD || mode == AR_IAF_INV);
    const T ca = iaf ? T(2) : T(3), cs = iaf ? T(3) : T(5);
    const T al = clamp_mm(al_raw, -ca, ca);
    const T mu = iaf ? clamp_mm(mu_raw, T(-10), T(10)) : mu_raw;
    if (mode == AR_MAF_INV || mode == AR_IAF_INV) {
        out = (v - mu) * t_exp(clamp_mm(-al, -cs, cs));
        ld_term = -al;
    } else {
        out = v * t_exp(clamp_mm(al, -cs, cs)) + mu;
        ld_term = al;
    }
}

template <typename T>
NF_HD void affine_ar_elem_bwd(int mode, T v, T mu_raw, T al_raw, T gout, T gld,
                              T& gv, T& gmu_raw, T& gal_raw) {
    const bool iaf = (mode == AR_IAF_FWD || mode == AR_IAF_INV);
    const T ca = iaf ? T(2) : T(3), cs = iaf ? T(3) : T(5);
    const T al = clamp_mm(al_raw, -ca, ca);
    const bool pa = pass_mm(al_raw, -ca, ca);
    const bool pm = iaf ? pass_mm(mu_raw, T(-10), T(10)) : true;
    const T mu = iaf ? clamp_mm(mu_raw, T(-10), T(10)) : mu_raw;
    if (mode == AR_MAF_INV || mode == AR_IAF_INV) {
        const T ls = -al;
        const T e = t_exp(clamp_mm(ls, -cs, cs));
        const bool pl = pass_mm(ls, -cs, cs);
        gv = gout * e;
        gmu_raw = pm ? -(gout * e) : T(0);
        T g_al = -gld;
        if (pl) g_al -= gout * (v - mu) * e;
        gal_raw = pa ? g_al : T(0);
    } else {
        const T e = t_exp(clamp_mm(al, -cs, cs));
        const bool pl = pass_mm(al, -cs, cs);
        gv = gout * e;
        gmu_raw = pm ? gout : T(0);
        T g_al = gld;
        if (pl) g_al += gout * v * e;
        gal_raw = pa ? g_al : T(0);
    }
}

// ------------------------------------------------------------------------------------------
// Rational-quadratic splines.
//   BOUNDED=true : SplineCouplingLayer._rational_quadratic_spline   (spline_coupling_layer.py:182-309)
//                  domain [-B,B], identity tails, eps=1e-8, end knots pinned, widths re-differenced.
//   BOUNDED=false: public rational_quadratic_spline                 (rational_quadratic_spline.py:4-104)
//                  domain [0,1], eps=1e-6, no tails, no pinning, per-bin width = normalised width.
// Raw parameters of one element: uw[K], uh[K], ud[K-1].  KMAX is the register-array extent; K<=KMAX
// is the live bin count (K==KMAX lets every guard fold at compile time).
// ------------------------------------------------------------------------------------------
template <typename T>
struct RqsCfg {
    T lo, hi;        // domain: (-B, B) or (0, 1)
    T span;          // 2*B (bounded) or 1
    T min_w, min_h, min_d;
    T scale_w, scale_h;   // 1 - min_w*K, 1 - min_h*K (computed on the host in double, like the python scalars)
    T eps;
};

// compile-time unrolled pairwise reductions / inclusive scan over register arrays (depth log2 instead of K)
template <typename T> NF_HD T max2(T l, T r) { return (r > l) ? r : l; }
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float max2(float l, float r) {         // one FMNMX; a NaN among the bins reaches the softmax
    float m;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(m) : "f"(l), "f"(r));
    return m;
}
#endif
template <typename T, int N> NF_HD T tree_max(const T* a) {
    if constexpr (N == 1) return a[0];
    else { const T l = tree_max<T, N / 2>(a), r = tree_max<T, N - N / 2>(a + N / 2); return max2(l, r); }
}
template <typename T, int N> NF_HD T tree_sum(const T* a) {
    if constexpr (N == 1) return a[0];
    else return tree_sum<T, N / 2>(a) + tree_sum<T, N - N / 2>(a + N / 2);
}
template <typename T, int N> NF_HD void inclusive_scan(T* a) {      // Hillis-Steele
#pragma unroll
    for (int off = 1; off < N; off <<= 1) {
        T t[N];
#pragma unroll
        for (int j = 0; j < N; ++j) t[j] = (j >= off) ? a[j] + a[j - off] : a[j];
#pragma unroll
        for (int j = 0; j < N; ++j) a[j] = t[j];
    }
}

// softmax -> floor -> clamp -> cumulative knots.  wn: normalised widths (post clamp); kn: K+1 knots.
// KMAX <= 16: max / sum / cumulative sum as trees (the per-row dependent chain is what bounds the fused stack
// kernels; summation order differs from torch.cumsum by a few ulp).  KMAX > 16 (generic path): sequential loops.
template <typename T, int KMAX, bool BOUNDED>
NF_HD void rqs_knots(const T* u, int K, T floor_, T scale, const RqsCfg<T>& c, T* wn, T* kn, T* smx = nullptr) {
    if constexpr (KMAX <= 16) {
        T m[KMAX];
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) m[j] = (j < K) ? u[j] : u[0];
        const T mx = tree_max<T, KMAX>(m);
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) wn[j] = (j < K) ? sm_exp(u[j] - mx) : T(0);
        const T inv = t_rcp(tree_sum<T, KMAX>(wn));
        T run[KMAX];
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) {
            const T p = wn[j] * inv;                       // softmax value (kept for the reverse sweep when asked for)
            if (smx) smx[j] = p;
            T w = floor_ + scale * p;
            w = clamp_min(w, c.eps);
            wn[j] = w;
            run[j] = (j < K) ? w : T(0);
        }
        inclusive_scan<T, KMAX>(run);
        kn[0] = BOUNDED ? c.lo : T(0);
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) kn[j + 1] = BOUNDED ? ((j + 1 == K) ? c.hi : (c.span * run[j] + c.lo)) : run[j];
    } else {
        T mx = u[0];
NF_UNROLL
        for (int j = 1; j < KMAX; ++j) if (j < K) mx = (u[j] > mx) ? u[j] : mx;
        T sum = T(0);
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) if (j < K) { wn[j] = sm_exp(u[j] - mx); sum += wn[j]; }
        const T inv = t_rcp(sum);
        T run = T(0);
        kn[0] = BOUNDED ? c.lo : T(0);
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) if (j < K) {
            const T p = wn[j] * inv;
            if (smx) smx[j] = p;
            T w = floor_ + scale * p;
            w = clamp_min(w, c.eps);
            wn[j] = w;
            run += w;
            kn[j + 1] = BOUNDED ? (c.span * run + c.lo) : run;
        }
        if (BOUNDED) {
NF_UNROLL
            for (int j = 0; j < KMAX; ++j) if (j + 1 == K) kn[j + 1] = c.hi;
        }
    }
}

#if defined(__CUDA_ARCH__)
// float, KMAX <= 16, device: widths AND heights of one element through one instruction stream on the packed fp32 pipe
// (Blackwell FADD2 / FMUL2 / FFMA2), (width_j, height_j) as a float2.  Same operations in the same order as two
// rqs_knots<float> calls, so the results are bit-identical; max, exp, reciprocal and the clamps have no packed form and
// stay scalar.  (The fused stack kernels and the streaming spline kernels are issue-bound; this halves ~45 of the ~70
// instructions per array.)
template <int N> __device__ __forceinline__ float2 tree_sum2(const float2* a) {
    if constexpr (N == 1) return a[0];
    else return __fadd2_rn(tree_sum2<N / 2>(a), tree_sum2<N - N / 2>(a + N / 2));
}
template <int KMAX, bool BOUNDED>
__device__ __forceinline__ void rqs_knots_pair(const float* uw, const float* uh, int K, const RqsCfg<float>& c,
                                               float* wn, float* hn, float* cw, float* ch, float* smw = nullptr,
                                               float* smh = nullptr) {
    float mw[KMAX], mh[KMAX];
NF_UNROLL
    for (int j = 0; j < KMAX; ++j) { mw[j] = (j < K) ? uw[j] : uw[0]; mh[j] = (j < K) ? uh[j] : uh[0]; }
    const float2 nmx = make_float2(-tree_max<float, KMAX>(mw), -tree_max<float, KMAX>(mh));
    float2 e[KMAX];
NF_UNROLL
    for (int j = 0; j < KMAX; ++j) {
        const float2 d = __fadd2_rn(make_float2(uw[j], uh[j]), nmx);
        e[j] = (j < K) ? sm_exp_x2(d) : make_float2(0.f, 0.f);
    }
    const float2 sum = tree_sum2<KMAX>(e);
    const float2 inv = make_float2(t_rcp(sum.x), t_rcp(sum.y));
    const float2 floor2 = make_float2(c.min_w, c.min_h), scale2 = make_float2(c.scale_w, c.scale_h);
    float2 run[KMAX];
NF_UNROLL
    for (int j = 0; j < KMAX; ++j) {
        const float2 p = __fmul2_rn(e[j], inv);
        if (smw) { smw[j] = p.x; smh[j] = p.y; }
        float2 w = __ffma2_rn(scale2, p, floor2);
        w.x = clamp_min(w.x, c.eps); w.y = clamp_min(w.y, c.eps);
        wn[j] = w.x; hn[j] = w.y;
        run[j] = (j < K) ? w : make_float2(0.f, 0.f);
    }
NF_UNROLL
    for (int off = 1; off < KMAX; off <<= 1) {              // Hillis-Steele inclusive scan, as inclusive_scan<>
        float2 t[KMAX];
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) t[j] = (j >= off) ? __fadd2_rn(run[j], run[j - off]) : run[j];
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) run[j] = t[j];
    }
    cw[0] = BOUNDED ? c.lo : 0.f; ch[0] = cw[0];
    const float2 span2 = make_float2(c.span, c.span), lo2 = make_float2(c.lo, c.lo);
NF_UNROLL
    for (int j = 0; j < KMAX; ++j) {
        float2 k = BOUNDED ? __ffma2_rn(span2, run[j], lo2) : run[j];
        if (BOUNDED && j + 1 == K) k = make_float2(c.hi, c.hi);
        cw[j + 1] = k.x; ch[j + 1] = k.y;
    }
}
#endif

// Quantities of the bin an input falls in.
template <typename T>
struct RqsBin {
    int k;
    T xk, xk1, yk, yk1;   // knots bracketing the bin on both axes
    T wn, hn;             // normalised width/height of the bin (used by the unit variant)
    T udk, udk1;          // raw derivative params at the two knots (ignored when edge flag set)
    bool lo_edge, hi_edge;// derivative pinned to 1 at k==0 / k==K-1
};

// value of a[k] for a register array through a binary multiplexer tree (depth log2 N, no local memory)
template <typename T, int N> NF_HD T mux(const T* a, int k) {
    if constexpr (N == 1) return a[0];
    else {
        constexpr int H = (N + 1) / 2;
        const T lo = mux<T, H>(a, k), hi = mux<T, N - H>(a + H, k - H);
        return (k >= H) ? hi : lo;
    }
}

// searchsorted(right=True)-1, clamped to [0,K-1]: the last j<K with knot_j <= v (j=0 always taken).  The knots are
// non-decreasing (cumulative sums of positive bin sizes), so that index equals the number of knots 1..K-1 that are
// <= v; NaN knots compare false and give bin 0, as the sequential scan did.
template <typename T, int KMAX>
NF_HD void rqs_select(T v, const T* sk /*search knots*/, const T* cw, const T* ch, const T* wn, const T* hn,
                      const T* ud, int K, RqsBin<T>& b) {
    if constexpr (KMAX <= 16) {
        int cnt[KMAX];
        cnt[0] = 0;
NF_UNROLL
        for (int j = 1; j < KMAX; ++j) cnt[j] = (j < K && sk[j] <= v) ? 1 : 0;
        const int k = tree_sum<int, KMAX>(cnt);
        b.k = k;
        b.xk = mux<T, KMAX>(cw, k); b.xk1 = mux<T, KMAX>(cw + 1, k);
        b.yk = mux<T, KMAX>(ch, k); b.yk1 = mux<T, KMAX>(ch + 1, k);
        b.wn = mux<T, KMAX>(wn, k); b.hn = mux<T, KMAX>(hn, k);
        // ud[j] = raw derivative parameter of knot j+1: knot k -> ud[k-1] (k >= 1), knot k+1 -> ud[k] (k <= K-2)
        b.udk1 = mux<T, KMAX>(ud, (k < KMAX - 1) ? k : KMAX - 1);
        b.udk = (k >= 1) ? mux<T, KMAX>(ud, k - 1) : T(0);
    } else {
        b.k = 0; b.xk = cw[0]; b.xk1 = cw[1]; b.yk = ch[0]; b.yk1 = ch[1]; b.wn = wn[0]; b.hn = hn[0];
        b.udk = T(0); b.udk1 = ud[0];
NF_UNROLL
        for (int j = 1; j < KMAX; ++j) if (j < K) {
            if (sk[j] <= v) {
                b.k = j; b.xk = cw[j]; b.xk1 = cw[j + 1]; b.yk = ch[j]; b.yk1 = ch[j + 1];
                b.wn = wn[j]; b.hn = hn[j]; b.udk = ud[j - 1];
                b.udk1 = (j < KMAX - 1) ? ud[j] : T(0);     // unused when j==K-1 (hi_edge)
            }
        }
    }
    b.lo_edge = (b.k == 0);
    b.hi_edge = (b.k == K - 1);
}

// value + log|derivative| inside the selected bin.  (x_k, y_k, w_k, h_k, d_k, d_k1) as in the reference.
template <typename T, bool BOUNDED>
NF_HD void rqs_bin_eval(T v, T xk, T yk, T wk, T hk, T dk, T dk1, bool inverse, T eps, T& out, T& lad) {
    const T wkc = clamp_min(wk, eps);
    const T s = t_div(hk, wkc);
    if (!inverse) {
        const T xi = clamp_mm(t_div(v - xk, wkc), T(0), T(1));
        const T om = T(1) - xi;
        const T A = dk1 + dk - T(2) * s;
        const T den0 = s + A * xi * om;
        const T denc = clamp_min(den0, eps);
        const T N1 = s * (xi * xi) + dk * xi * om;
        out = yk + t_div(hk * N1, denc);
        const T num = (s * s) * (dk1 * (xi * xi) + T(2) * s * xi * om + dk * (om * om));
        const T den2 = BOUNDED ? denc * denc : den0 * den0;
        const T der = t_div(num, clamp_min(den2, eps));
        lad = t_logf(clamp_min(der, eps));
    } else {
        const T t0 = v - yk;
        const T A = dk + dk1 - T(2) * s;
        const T t = t0 * A;
        const T a = BOUNDED ? (t + hk * (s - dk)) : (hk * (s - dk) + t);
        const T b = hk * dk - t;
        const T c = -s * t0;
        const T disc = clamp_min(b * b - T(4) * a * c, T(0));
        T q = -b - t_sqrt(disc);
        if (BOUNDED) { if (t_abs(q) < eps) q = eps; }
        const T xi = clamp_mm(t_div(T(2) * c, q), T(0), T(1));
        out = xi * wk + xk;
        const T om = T(1) - xi;
        const T Q = dk1 * (xi * xi) + (BOUNDED ? T(2) * s * xi * om : T(2) * s * (xi * om)) + dk * (om * om);
        const T num = (s * s) * Q;
        if (BOUNDED) {
            const T den0 = s + (dk1 + dk - T(2) * s) * xi * om;
            lad = -t_logf(clamp_min(num, eps)) + T(2) * t_logf(clamp_min(den0, eps));
        } else {
            const T den0 = s + A * (xi * om);
            const T der = t_div(num, clamp_min(den0 * den0, eps));
            lad = -t_logf(clamp_min(der, eps));
        }
    }
}

// Full element evaluation from raw parameters.
template <typename T, int KMAX, bool BOUNDED>
NF_HD void rqs_eval(T v, const T* uw, const T* uh, const T* ud, int K, bool inverse, const RqsCfg<T>& c,
                    T& out, T& lad) {
    if (BOUNDED) {
        const bool inside = (v >= c.lo) && (v <= c.hi);
        if (!inside) { out = v; lad = T(0); return; }       // identity tails (:192-197); NaN lands here too
    }
    T wn[KMAX], hn[KMAX], cw[KMAX + 1], ch[KMAX + 1];
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4 && KMAX <= 16) {
        rqs_knots_pair<KMAX, BOUNDED>(reinterpret_cast<const float*>(uw), reinterpret_cast<const float*>(uh), K,
                                      reinterpret_cast<const RqsCfg<float>&>(c), reinterpret_cast<float*>(wn),
                                      reinterpret_cast<float*>(hn), reinterpret_cast<float*>(cw), reinterpret_cast<float*>(ch));
    } else
#endif
    {
        rqs_knots<T, KMAX, BOUNDED>(uw, K, c.min_w, c.scale_w, c, wn, cw);
        rqs_knots<T, KMAX, BOUNDED>(uh, K, c.min_h, c.scale_h, c, hn, ch);
    }
    RqsBin<T> b;
    rqs_select<T, KMAX>(v, inverse ? ch : cw, cw, ch, wn, hn, ud, K, b);
    T spk, spk1;
    softplus_x2(b.udk, b.udk1, spk, spk1);
    const T dk  = b.lo_edge ? T(1) : clamp_min(c.min_d + spk, c.eps);
    const T dk1 = b.hi_edge ? T(1) : clamp_min(c.min_d + spk1, c.eps);
    T wk, hk;
    if (BOUNDED) { wk = clamp_min(b.xk1 - b.xk, c.eps); hk = clamp_min(b.yk1 - b.yk, c.eps); }
    else         { wk = b.wn; hk = b.hn; }
    rqs_bin_eval<T, BOUNDED>(v, b.xk, b.yk, wk, hk, dk, dk1, inverse, c.eps, out, lad);
    if (BOUNDED) {                                          // :306-307
        if (!is_finite(out)) out = v;
        if (!is_finite(lad)) lad = T(0);
    }
}

// ------------------------------------------------------------------------------------------
// Reverse mode.  Given upstream (g_out, g_lad) of one element, produce g_v and the raw-parameter gradients
// guw[K], guh[K], gud[K-1].  Forward intermediates are recomputed ONCE (the element is 3K-1 parameters; nothing
// is stashed): the knots keep their softmax values for the softmax backward, the bin evaluation yields the
// element's own output (needed for the scrubs) and its reverse sweep from the same intermediates, every
// division is one reciprocal shared by its uses.  (The first version evaluated the element three times --
// scrub check, forward recomputation, softmax backward -- with IEEE divisions: 2 360 instructions per element.)
// ------------------------------------------------------------------------------------------
template <typename T>
struct RqsBinGrad { T gv, gxk, gyk, gwk, ghk, gdk, gdk1; };

// How the element's own forward result (out, lad of the selected bin, BEFORE the scrubs of :306-307) adjusts the
// upstream gradients: a replaced output passes its gradient straight to the input (gv_direct), a replaced log-det
// term receives none.  Kernels with a layer-level scrub / rescale supply their own functor.
template <typename T, bool BOUNDED>
struct RqsScrubAdj {
    NF_HD void operator()(T out, T lad, T& g_out, T& g_lad, T& gv_direct) const {
        if (BOUNDED) {
            if (!is_finite(lad)) g_lad = T(0);
            if (!is_finite(out)) { gv_direct = g_out; g_out = T(0); }
        }
    }
};

// value, log|derivative| and reverse sweep inside the selected bin.  Same formulas as rqs_bin_eval; a / b is
// written a * (1 / b) with the reciprocal shared (float device: rcp.approx, 1 ulp; double / host: exact).
template <typename T, bool BOUNDED, typename Adj>
NF_HD void rqs_bin_fwdbwd(T v, T xk, T yk, T wk, T hk, T dk, T dk1, bool inverse, T eps, T g_out, T g_lad,
                          const Adj& adj, RqsBinGrad<T>& g, T& gv_direct) {
    const T wkc = clamp_min(wk, eps);
    const bool p_wkc = pass_min(wk, eps);
    const T rw = t_div(T(1), wkc);
    const T s = hk * rw;
    const T A = dk1 + dk - T(2) * s;
    T g_s, g_xi, g_om, g_A, g_wkc;
    gv_direct = T(0);
    if (!inverse) {
        const T xi0 = (v - xk) * rw;
        const T xi = clamp_mm(xi0, T(0), T(1));
        const T om = T(1) - xi;
        const T xo = xi * om, xx = xi * xi, oo = om * om;
        const T den0 = s + A * xo;
        const T denc = clamp_min(den0, eps);
        const T rden = t_div(T(1), denc);
        const T N1 = s * xx + dk * xo;
        const T Q = dk1 * xx + T(2) * s * xo + dk * oo;
        const T num = (s * s) * Q;
        const T den2 = BOUNDED ? denc * denc : den0 * den0;
        const T den2c = clamp_min(den2, eps);
        const T rden2 = t_div(T(1), den2c);
        const T der = num * rden2;
        const T derc = clamp_min(der, eps);
        const T hr = hk * rden;                         // d out / d N1
        const T nr = N1 * rden;                         // d out / d hk
        adj(yk + hk * nr, t_logf(derc), g_out, g_lad, gv_direct);
        // lad = log(derc), der = num / den2c
        const T g_der = pass_min(der, eps) ? g_lad * t_div(T(1), derc) : T(0);
        const T g_num = g_der * rden2;
        const T g_den2 = pass_min(den2, eps) ? -(g_der * der) * rden2 : T(0);
        T g_denc = T(0), g_den0 = T(0);
        if (BOUNDED) g_denc = T(2) * denc * g_den2; else g_den0 = T(2) * den0 * g_den2;
        // num = s^2 Q
        const T g_Q = (s * s) * g_num;
        g_s = T(2) * s * Q * g_num + T(2) * xo * g_Q;
        g.gdk1 = xx * g_Q;
        g.gdk = oo * g_Q;
        g_xi = T(2) * (dk1 * xi + s * om) * g_Q;
        g_om = T(2) * (s * xi + dk * om) * g_Q;
        // out = yk + hk * N1 / denc
        g.gyk = g_out;
        g.ghk = nr * g_out;
        const T g_N1 = hr * g_out;
        g_denc -= hr * nr * g_out;
        g_s += xx * g_N1;
        g_xi += (T(2) * s * xi + dk * om) * g_N1;
        g.gdk += xo * g_N1;
        g_om += dk * xi * g_N1;
        // denc = clamp(den0), den0 = s + A xi om
        if (pass_min(den0, eps)) g_den0 += g_denc;
        g_s += g_den0;
        g_A = xo * g_den0;
        g_xi += A * om * g_den0;
        g_om += A * xi * g_den0;
        g_xi -= g_om;
        const T g_xi0 = pass_mm(xi0, T(0), T(1)) ? g_xi : T(0);
        g.gv = g_xi0 * rw;
        g.gxk = -g.gv;
        g_wkc = -xi0 * g.gv;
        g.gwk = T(0);
    } else {
        const T t0 = v - yk;
        const T t = t0 * A;
        const T a = t + hk * (s - dk);
        const T b = hk * dk - t;
        const T c = -s * t0;
        const T disc0 = b * b - T(4) * a * c;
        const T disc = clamp_min(disc0, T(0));
        const T sq = t_sqrt(disc);
        const T q0 = -b - sq;
        const bool q_small = BOUNDED && (t_abs(q0) < eps);
        const T q = q_small ? eps : q0;
        const T rq = t_div(T(1), q);
        const T xi0 = (T(2) * c) * rq;
        const T xi = clamp_mm(xi0, T(0), T(1));
        const T om = T(1) - xi;
        const T xo = xi * om, xx = xi * xi, oo = om * om;
        const T Q = dk1 * xx + T(2) * s * xo + dk * oo;
        const T num = (s * s) * Q;
        const T den0 = s + A * xo;
        T g_num, g_den0;
        if (BOUNDED) {
            const T numc = clamp_min(num, eps), denc = clamp_min(den0, eps);
            adj(xi * wk + xk, -t_logf(numc) + T(2) * t_logf(denc), g_out, g_lad, gv_direct);
            g_num = pass_min(num, eps) ? -g_lad * t_div(T(1), numc) : T(0);
            g_den0 = pass_min(den0, eps) ? T(2) * g_lad * t_div(T(1), denc) : T(0);
        } else {
            const T den2 = den0 * den0;
            const T den2c = clamp_min(den2, eps);
            const T rden2 = t_div(T(1), den2c);
            const T der = num * rden2;
            const T derc = clamp_min(der, eps);
            adj(xi * wk + xk, -t_logf(derc), g_out, g_lad, gv_direct);
            const T g_der = pass_min(der, eps) ? -g_lad * t_div(T(1), derc) : T(0);
            g_num = g_der * rden2;
            g_den0 = pass_min(den2, eps) ? T(2) * den0 * (-(g_der * der) * rden2) : T(0);
        }
        const T g_Q = (s * s) * g_num;
        g_s = T(2) * s * Q * g_num + T(2) * xo * g_Q + g_den0;
        g.gdk1 = xx * g_Q;
        g.gdk = oo * g_Q;
        g_xi = T(2) * (dk1 * xi + s * om) * g_Q + A * om * g_den0;
        g_om = T(2) * (s * xi + dk * om) * g_Q + A * xi * g_den0;
        g_A = xo * g_den0;
        // out = xi * wk + xk
        g_xi += wk * g_out;
        g.gwk = xi * g_out;
        g.gxk = g_out;
        g_xi -= g_om;
        const T g_xi0 = pass_mm(xi0, T(0), T(1)) ? g_xi : T(0);
        // xi0 = 2 c / q, q = -b - sqrt(disc)
        T g_c = T(2) * rq * g_xi0;
        const T g_q0 = q_small ? T(0) : -xi0 * rq * g_xi0;
        const T g_disc = (sq > T(0)) ? -g_q0 * t_div(T(0.5), sq) : T(0);
        const T g_disc0 = pass_min(disc0, T(0)) ? g_disc : T(0);
        const T g_b = -g_q0 + T(2) * b * g_disc0;
        const T g_a = -T(4) * c * g_disc0;
        g_c -= T(4) * a * g_disc0;
        g_s -= t0 * g_c;
        // b = hk dk - t, a = t + hk (s - dk), t = t0 A
        g.ghk = dk * g_b + (s - dk) * g_a;
        g.gdk += hk * (g_b - g_a);
        g_s += hk * g_a;
        const T g_t = g_a - g_b;
        const T g_t0 = A * g_t - s * g_c;
        g_A += t0 * g_t;
        g.gv = g_t0;
        g.gyk = -g_t0;
        g_wkc = T(0);
    }
    // A = dk1 + dk - 2 s;  s = hk / wkc
    g.gdk1 += g_A; g.gdk += g_A; g_s -= T(2) * g_A;
    g.ghk += g_s * rw;
    g_wkc -= s * rw * g_s;
    if (p_wkc) g.gwk += g_wkc;
}

// push d(loss)/d(normalised bin sizes) through floor+clamp and the softmax:  gu = J^T gw, with the softmax values sm
// kept from the knot computation
template <typename T, int KMAX>
NF_HD void rqs_softmax_bwd(const T* sm, int K, T floor_, T scale, T eps, const T* gw, T* gu) {
    T dot = T(0);
    T gs[KMAX];
NF_UNROLL
    for (int j = 0; j < KMAX; ++j) if (j < K) {
        const T w = floor_ + scale * sm[j];
        gs[j] = pass_min(w, eps) ? scale * gw[j] : T(0);
        dot += gs[j] * sm[j];
    }
NF_UNROLL
    for (int j = 0; j < KMAX; ++j) gu[j] = (j < K) ? sm[j] * (gs[j] - dot) : T(0);
}

// ASSIGNS g_v, guw[0..KMAX), guh[0..KMAX), gud[0..KMAX-1) (zeros beyond the live bins).
template <typename T, int KMAX, bool BOUNDED, typename Adj>
NF_HD void rqs_eval_grad(T v, const T* uw, const T* uh, const T* ud, int K, bool inverse, const RqsCfg<T>& c,
                         T g_out, T g_lad, const Adj& adj, T& g_v, T* guw, T* guh, T* gud) {
    if (BOUNDED) {
        const bool inside = (v >= c.lo) && (v <= c.hi);
        if (!inside) {                                      // identity tails: out = v, lad = 0
            T gvd = T(0);
            adj(v, T(0), g_out, g_lad, gvd);
            g_v = g_out + gvd;
NF_UNROLL
            for (int j = 0; j < KMAX; ++j) { guw[j] = T(0); guh[j] = T(0); }
NF_UNROLL
            for (int j = 0; j < KMAX - 1; ++j) gud[j] = T(0);
            return;
        }
    }
    T wn[KMAX], hn[KMAX], cw[KMAX + 1], ch[KMAX + 1], smw[KMAX], smh[KMAX];
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4 && KMAX <= 16) {
        rqs_knots_pair<KMAX, BOUNDED>(reinterpret_cast<const float*>(uw), reinterpret_cast<const float*>(uh), K,
                                      reinterpret_cast<const RqsCfg<float>&>(c), reinterpret_cast<float*>(wn),
                                      reinterpret_cast<float*>(hn), reinterpret_cast<float*>(cw), reinterpret_cast<float*>(ch),
                                      reinterpret_cast<float*>(smw), reinterpret_cast<float*>(smh));
    } else
#endif
    {
        rqs_knots<T, KMAX, BOUNDED>(uw, K, c.min_w, c.scale_w, c, wn, cw, smw);
        rqs_knots<T, KMAX, BOUNDED>(uh, K, c.min_h, c.scale_h, c, hn, ch, smh);
    }
    RqsBin<T> b;
    rqs_select<T, KMAX>(v, inverse ? ch : cw, cw, ch, wn, hn, ud, K, b);
    // derivatives at the two knots: softplus and its slope from one exponential each
    T zk, zk1;
    sm_exp_pair(b.udk > T(20) ? T(20) : b.udk, b.udk1 > T(20) ? T(20) : b.udk1, zk, zk1);
    const T dk_pre  = c.min_d + (b.udk > T(20) ? b.udk : t_log1p(zk));
    const T dk1_pre = c.min_d + (b.udk1 > T(20) ? b.udk1 : t_log1p(zk1));
    const T dk  = b.lo_edge ? T(1) : clamp_min(dk_pre, c.eps);
    const T dk1 = b.hi_edge ? T(1) : clamp_min(dk1_pre, c.eps);
    T wk, hk;
    if (BOUNDED) { wk = clamp_min(b.xk1 - b.xk, c.eps); hk = clamp_min(b.yk1 - b.yk, c.eps); }
    else         { wk = b.wn; hk = b.hn; }
    RqsBinGrad<T> g;
    T gv_direct;
    rqs_bin_fwdbwd<T, BOUNDED, Adj>(v, b.xk, b.yk, wk, hk, dk, dk1, inverse, c.eps, g_out, g_lad, adj, g, gv_direct);
    g_v = g.gv + gv_direct;
    // derivative parameters: ud[j] belongs to knot j + 1
    const T sk  = (!b.lo_edge && pass_min(dk_pre, c.eps)) ? g.gdk * (b.udk > T(20) ? T(1) : zk * t_div(T(1), zk + T(1))) : T(0);
    const T sk1 = (!b.hi_edge && pass_min(dk1_pre, c.eps)) ? g.gdk1 * (b.udk1 > T(20) ? T(1) : zk1 * t_div(T(1), zk1 + T(1))) : T(0);
NF_UNROLL
    for (int j = 0; j < KMAX - 1; ++j) gud[j] = (j == b.k - 1 ? sk : T(0)) + (j == b.k ? sk1 : T(0));
    // bin sizes: map (gxk, gwk) / (gyk, ghk) onto d/d(normalised widths / heights)
    T gw[KMAX], gh[KMAX];
    if (BOUNDED) {
        // wk = clamp(xk1 - xk): gxk1 = +gwk', gxk -= gwk';  knots j (1..K-1) = span*cumsum_{i<j} + lo
        T gxk = g.gxk, gxk1 = T(0), gyk = g.gyk, gyk1 = T(0);
        if (pass_min(b.xk1 - b.xk, c.eps)) { gxk1 += g.gwk; gxk -= g.gwk; }
        if (pass_min(b.yk1 - b.yk, c.eps)) { gyk1 += g.ghk; gyk -= g.ghk; }
        const T aw = (b.k >= 1) ? c.span * gxk : T(0);           // knot k interior?
        const T bw = (b.k + 1 <= K - 1) ? c.span * gxk1 : T(0);   // knot k+1 interior?
        const T ah = (b.k >= 1) ? c.span * gyk : T(0);
        const T bh = (b.k + 1 <= K - 1) ? c.span * gyk1 : T(0);
        const T abw = aw + bw, abh = ah + bh;
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) {
            gw[j] = (j < b.k) ? abw : (j == b.k ? bw : T(0));
            gh[j] = (j < b.k) ? abh : (j == b.k ? bh : T(0));
        }
    } else {
NF_UNROLL
        for (int j = 0; j < KMAX; ++j) {
            gw[j] = (j < b.k) ? g.gxk : (j == b.k ? g.gwk : T(0));
            gh[j] = (j < b.k) ? g.gyk : (j == b.k ? g.ghk : T(0));
        }
    }
    rqs_softmax_bwd<T, KMAX>(smw, K, c.min_w, c.scale_w, c.eps, gw, guw);
    rqs_softmax_bwd<T, KMAX>(smh, K, c.min_h, c.scale_h, c.eps, gh, guh);
}

// accumulating form (caller zero-initialises guw / guh / gud), default scrub handling
template <typename T, int KMAX, bool BOUNDED>
NF_HD void rqs_eval_bwd(T v, const T* uw, const T* uh, const T* ud, int K, bool inverse, const RqsCfg<T>& c,
                        T g_out, T g_lad, T& g_v, T* guw, T* guh, T* gud) {
    T a[KMAX], b[KMAX], d[KMAX];
    rqs_eval_grad<T, KMAX, BOUNDED>(v, uw, uh, ud, K, inverse, c, g_out, g_lad, RqsScrubAdj<T, BOUNDED>(), g_v, a, b, d);
NF_UNROLL
    for (int j = 0; j < KMAX; ++j) if (j < K) { guw[j] += a[j]; guh[j] += b[j]; }
NF_UNROLL
    for (int j = 0; j < KMAX - 1; ++j) if (j < K - 1) gud[j] += d[j];
}

}  // namespace nf
