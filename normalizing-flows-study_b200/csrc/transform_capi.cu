// transform_capi.cu -- C-ABI entry points of the spline transforms (dispatch on dtype and bin count).
#include "transform_impl.cuh"

namespace nf {
#define NF_DECL(T, G)                                                                                                  \
    extern template int rqs_unit_fwd_launch<T, G>(const void*, const void*, const void*, const void*, void*, void*,   \
                                                  int64_t, int, int, RqsCfg<T>, cudaStream_t);                         \
    extern template int rqs_unit_bwd_launch<T, G>(const void*, const void*, const void*, const void*, const void*,    \
                                                  const void*, void*, void*, void*, void*, int64_t, int, int,          \
                                                  RqsCfg<T>, cudaStream_t);                                            \
    extern template int spline_transform_launch<T, false, G>(const SplineTfArgs<T>&, cudaStream_t);                   \
    extern template int spline_transform_launch<T, true, G>(const SplineTfArgs<T>&, cudaStream_t);
NF_DECL(float, false) NF_DECL(float, true) NF_DECL(double, false) NF_DECL(double, true)
#undef NF_DECL

template <typename T>
static int rqs_unit_fwd_any(const void* x, const void* w, const void* h, const void* d, void* y, void* ld, int64_t n,
                            int K, int inverse, double mw, double mh, double md, cudaStream_t st) {
    auto c = make_rqs_cfg<T>(false, K, 0, mw, mh, md);
    return K <= 16 ? rqs_unit_fwd_launch<T, false>(x, w, h, d, y, ld, n, K, inverse, c, st)
                   : rqs_unit_fwd_launch<T, true>(x, w, h, d, y, ld, n, K, inverse, c, st);
}
template <typename T>
static int rqs_unit_bwd_any(const void* x, const void* w, const void* h, const void* d, const void* gy,
                            const void* gld, void* gx, void* gw, void* gh, void* gd, int64_t n, int K, int inverse,
                            double mw, double mh, double md, cudaStream_t st) {
    auto c = make_rqs_cfg<T>(false, K, 0, mw, mh, md);
    return K <= 16 ? rqs_unit_bwd_launch<T, false>(x, w, h, d, gy, gld, gx, gw, gh, gd, n, K, inverse, c, st)
                   : rqs_unit_bwd_launch<T, true>(x, w, h, d, gy, gld, gx, gw, gh, gd, n, K, inverse, c, st);
}
template <typename T, bool BWD>
static int spline_tf_any(const void* x, const void* params, const void* mask, const int32_t* tidx, void* y, void* ld,
                         const void* gy, const void* gld, void* gx, void* gparams, int64_t B, int D, int Dt, int K,
                         int inverse, double bound, double mw, double mh, double md, const void* r_in,
                         const void* r_lo, const void* r_out, int compact, cudaStream_t st) {
    SplineTfArgs<T> a;
    a.x = (const T*)x; a.params = (const T*)params; a.mask = (const T*)mask; a.tidx = tidx;
    a.y = (T*)y; a.ld = (T*)ld; a.gy = (const T*)gy; a.gld = (const T*)gld; a.gx = (T*)gx; a.gparams = (T*)gparams;
    a.B = B; a.D = D; a.Dt = Dt; a.K = K; a.inverse = inverse;
    a.c = make_rqs_cfg<T>(true, K, bound, mw, mh, md);
    a.r_in = (const T*)r_in; a.r_lo = (const T*)r_lo; a.r_out = (const T*)r_out;
    a.compact = compact;
    return K <= 16 ? spline_transform_launch<T, BWD, false>(a, st) : spline_transform_launch<T, BWD, true>(a, st);
}
}  // namespace nf

using namespace nf;
#define NF_REQ(p) do { if ((p) == nullptr) return NF_ERR_NULL; } while (0)

extern "C" int nf_rqs_unit_forward(const void* x, const void* w, const void* h, const void* d, void* y, void* ld,
                                   int64_t n, int num_bins, int inverse, double min_w, double min_h, double min_d,
                                   int dtype, nf_stream_t stream) {
    if (n < 0 || num_bins < 1 || num_bins > 32) return NF_ERR_BAD_SHAPE;
    if (n == 0) return NF_OK;
    NF_REQ(x); NF_REQ(w); NF_REQ(h); NF_REQ(y); NF_REQ(ld);
    if (num_bins > 1) NF_REQ(d);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NF_F32) return rqs_unit_fwd_any<float>(x, w, h, d, y, ld, n, num_bins, inverse, min_w, min_h, min_d, st);
    if (dtype == NF_F64) return rqs_unit_fwd_any<double>(x, w, h, d, y, ld, n, num_bins, inverse, min_w, min_h, min_d, st);
    return NF_ERR_UNSUPPORTED;
}

extern "C" int nf_rqs_unit_backward(const void* x, const void* w, const void* h, const void* d, const void* gy,
                                    const void* gld, void* gx, void* gw, void* gh, void* gd, int64_t n, int num_bins,
                                    int inverse, double min_w, double min_h, double min_d, int dtype,
                                    nf_stream_t stream) {
    if (n < 0 || num_bins < 1 || num_bins > 32) return NF_ERR_BAD_SHAPE;
    if (n == 0) return NF_OK;
    NF_REQ(x); NF_REQ(w); NF_REQ(h); NF_REQ(gy); NF_REQ(gld); NF_REQ(gx); NF_REQ(gw); NF_REQ(gh);
    if (num_bins > 1) { NF_REQ(d); NF_REQ(gd); }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NF_F32)
        return rqs_unit_bwd_any<float>(x, w, h, d, gy, gld, gx, gw, gh, gd, n, num_bins, inverse, min_w, min_h, min_d, st);
    if (dtype == NF_F64)
        return rqs_unit_bwd_any<double>(x, w, h, d, gy, gld, gx, gw, gh, gd, n, num_bins, inverse, min_w, min_h, min_d, st);
    return NF_ERR_UNSUPPORTED;
}

extern "C" int nf_spline_transform_forward(const void* x, const void* params, const void* mask, const int32_t* tidx,
                                           void* y, void* ld, int64_t B, int D, int Dt, int num_bins, int inverse,
                                           double bound, double min_w, double min_h, double min_d,
                                           const void* r_in, const void* r_lo, const void* r_out, int params_compact,
                                           int dtype, nf_stream_t stream) {
    if (B < 0 || D < 1 || Dt < 0 || Dt > D || num_bins < 1 || num_bins > 32) return NF_ERR_BAD_SHAPE;
    if (B == 0) return NF_OK;
    NF_REQ(x); NF_REQ(mask); NF_REQ(y); NF_REQ(ld);
    if (Dt > 0) { NF_REQ(params); NF_REQ(tidx); }
    if ((r_in != nullptr) != (r_lo != nullptr) || (r_in != nullptr) != (r_out != nullptr)) return NF_ERR_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NF_F32)
        return spline_tf_any<float, false>(x, params, mask, tidx, y, ld, nullptr, nullptr, nullptr, nullptr, B, D, Dt,
                                           num_bins, inverse, bound, min_w, min_h, min_d, r_in, r_lo, r_out, params_compact, st);
    if (dtype == NF_F64)
        return spline_tf_any<double, false>(x, params, mask, tidx, y, ld, nullptr, nullptr, nullptr, nullptr, B, D, Dt,
                                            num_bins, inverse, bound, min_w, min_h, min_d, r_in, r_lo, r_out, params_compact, st);
    return NF_ERR_UNSUPPORTED;
}

extern "C" int nf_spline_transform_backward(const void* x, const void* params, const void* mask, const int32_t* tidx,
                                            const void* gy, const void* gld, void* gx, void* gparams, int64_t B, int D,
                                            int Dt, int num_bins, int inverse, double bound, double min_w,
                                            double min_h, double min_d, const void* r_in, const void* r_lo,
                                            const void* r_out, int params_compact, int dtype, nf_stream_t stream) {
    if (B < 0 || D < 1 || Dt < 0 || Dt > D || num_bins < 1 || num_bins > 32) return NF_ERR_BAD_SHAPE;
    if (B == 0) return NF_OK;
    NF_REQ(x); NF_REQ(mask); NF_REQ(gy); NF_REQ(gld); NF_REQ(gx);
    if (Dt > 0) { NF_REQ(params); NF_REQ(tidx); NF_REQ(gparams); }
    if ((r_in != nullptr) != (r_lo != nullptr) || (r_in != nullptr) != (r_out != nullptr)) return NF_ERR_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NF_F32)
        return spline_tf_any<float, true>(x, params, mask, tidx, nullptr, nullptr, gy, gld, gx, gparams, B, D, Dt,
                                          num_bins, inverse, bound, min_w, min_h, min_d, r_in, r_lo, r_out, params_compact, st);
    if (dtype == NF_F64)
        return spline_tf_any<double, true>(x, params, mask, tidx, nullptr, nullptr, gy, gld, gx, gparams, B, D, Dt,
                                           num_bins, inverse, bound, min_w, min_h, min_d, r_in, r_lo, r_out, params_compact, st);
    return NF_ERR_UNSUPPORTED;
}
