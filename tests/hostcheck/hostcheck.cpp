// Test-only harness: compiles the product's scalar transform math (csrc/nf_math.cuh, which is
// __host__ __device__) for the CPU so that tests can compare it with the oracle and with torch
// autograd without a GPU.  Never linked into libnfb200.so; never used by the product.
#include <cstring>
#include "nf_math.cuh"

using namespace nf;
static const int KM = 32;

template <typename T, bool BOUNDED>
static void run_rqs(const T* x, const T* w, const T* h, const T* d, long n, int K, int inverse, const double* cfg,
                    T* y, T* ld, const T* gy, const T* gld, T* gx, T* gw, T* gh, T* gd) {
    RqsCfg<T> c;
    c.lo = (T)cfg[0]; c.hi = (T)cfg[1]; c.span = (T)cfg[2]; c.min_w = (T)cfg[3]; c.min_h = (T)cfg[4];
    c.min_d = (T)cfg[5]; c.scale_w = (T)cfg[6]; c.scale_h = (T)cfg[7]; c.eps = (T)cfg[8];
    for (long i = 0; i < n; ++i) {
        T uw[KM] = {0}, uh[KM] = {0}, ud[KM] = {0};
        for (int j = 0; j < K; ++j) { uw[j] = w[i * K + j]; uh[j] = h[i * K + j]; }
        for (int j = 0; j < K - 1; ++j) ud[j] = d[i * (K - 1) + j];
        rqs_eval<T, KM, BOUNDED>(x[i], uw, uh, ud, K, inverse != 0, c, y[i], ld[i]);
        if (gy) {
            T guw[KM] = {0}, guh[KM] = {0}, gud[KM] = {0};
            rqs_eval_bwd<T, KM, BOUNDED>(x[i], uw, uh, ud, K, inverse != 0, c, gy[i], gld[i], gx[i], guw, guh, gud);
            for (int j = 0; j < K; ++j) { gw[i * K + j] = guw[j]; gh[i * K + j] = guh[j]; }
            for (int j = 0; j < K - 1; ++j) gd[i * (K - 1) + j] = gud[j];
        }
    }
}

extern "C" {
void hc_rqs_f32(int bounded, const float* x, const float* w, const float* h, const float* d, long n, int K, int inverse,
                const double* cfg, float* y, float* ld, const float* gy, const float* gld, float* gx, float* gw,
                float* gh, float* gd) {
    if (bounded) run_rqs<float, true>(x, w, h, d, n, K, inverse, cfg, y, ld, gy, gld, gx, gw, gh, gd);
    else run_rqs<float, false>(x, w, h, d, n, K, inverse, cfg, y, ld, gy, gld, gx, gw, gh, gd);
}
void hc_rqs_f64(int bounded, const double* x, const double* w, const double* h, const double* d, long n, int K,
                int inverse, const double* cfg, double* y, double* ld, const double* gy, const double* gld, double* gx,
                double* gw, double* gh, double* gd) {
    if (bounded) run_rqs<double, true>(x, w, h, d, n, K, inverse, cfg, y, ld, gy, gld, gx, gw, gh, gd);
    else run_rqs<double, false>(x, w, h, d, n, K, inverse, cfg, y, ld, gy, gld, gx, gw, gh, gd);
}
// affine coupling / affine AR element ops on flat arrays (f64 for gradient checks)
void hc_affine_coupling_f64(const double* x, const double* m, const double* s, const double* b, long n, int inverse,
                            double* y, double* ldt, const double* gy, const double* gld, double* gx, double* gs,
                            double* gb) {
    for (long i = 0; i < n; ++i) {
        affine_coupling_elem<double>(x[i], m[i], s[i], b[i], inverse != 0, y[i], ldt[i]);
        if (gy) affine_coupling_elem_bwd<double>(x[i], m[i], s[i], b[i], inverse != 0, gy[i], gld[i], gx[i], gs[i], gb[i]);
    }
}
void hc_affine_ar_f64(int mode, const double* v, const double* mu, const double* al, long n, double* out, double* ldt,
                      const double* gout, const double* gld, double* gv, double* gmu, double* gal) {
    for (long i = 0; i < n; ++i) {
        affine_ar_elem<double>(mode, v[i], mu[i], al[i], out[i], ldt[i]);
        if (gout) affine_ar_elem_bwd<double>(mode, v[i], mu[i], al[i], gout[i], gld[i], gv[i], gmu[i], gal[i]);
    }
}
}
