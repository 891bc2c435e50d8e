// host/golden_plf.cpp -- CPU golden model for the host's verification phase (see golden_plf.h).
// Strict fp32: products and sums rounded separately, sums left to right from +0.0f -- the
// arithmetic of app/src/plf.cpp:29-64.  Build with -ffp-contract=off.
#include "golden_plf.h"

#include <cmath>

namespace plfhost {

namespace {

struct Mat4 {
    const float *m;   // row-major [row][col]
    // y[row] = sum_col m[row][col] * v[col], col ascending, starting from +0.0f
    void mul_vec(const float *v, float *y) const
    {
        for (int row = 0; row < 4; ++row) {
            float s = 0.0f;
            for (int col = 0; col < 4; ++col) s += v[col] * m[4 * row + col];
            y[row] = s;
        }
    }
    // y[col] = sum_row v[row] * m[row][col], row ascending, starting from +0.0f
    void vec_mul(const float *v, float *y) const
    {
        for (int col = 0; col < 4; ++col) {
            float s = 0.0f;
            for (int row = 0; row < 4; ++row) s += v[row] * m[4 * row + col];
            y[col] = s;
        }
    }
};

}  // namespace

void golden_plf(const float *x1, const float *x2, float *x3, const float *ev, size_t n,
                const float *left, const float *right, const int *wgt, long long &scaler_increment,
                unsigned char *scaler)
{
    const float tiny = std::ldexp(1.0f, -32);
    const float huge = std::ldexp(1.0f, 32);
    const Mat4 EV{ev};
    long long inc = 0;
    for (size_t site = 0; site < n; ++site) {
        float *dst = x3 + 16 * site;
        bool underflow = true;
        for (int cat = 0; cat < 4; ++cat) {
            float a[4], b[4], p[4];
            Mat4{left + 16 * cat}.mul_vec(x1 + 16 * site + 4 * cat, a);
            Mat4{right + 16 * cat}.mul_vec(x2 + 16 * site + 4 * cat, b);
            for (int k = 0; k < 4; ++k) p[k] = a[k] * b[k];
            EV.vec_mul(p, dst + 4 * cat);
            for (int l = 0; l < 4; ++l) underflow = underflow && (std::fabs(dst[4 * cat + l]) < tiny);
        }
        if (underflow) {
            for (int e = 0; e < 16; ++e) dst[e] *= huge;
            inc += wgt ? wgt[site] : 1;
        }
        if (scaler) scaler[site] = underflow ? 1 : 0;
    }
    scaler_increment = inc;
}

void golden_plf_states(unsigned S, const float *x1, const float *x2, float *x3, const float *ev, size_t n,
                       const float *left, const float *right, const int *wgt, long long &scaler_increment,
                       unsigned char *scaler)
{
    const float tiny = std::ldexp(1.0f, -32);
    const float huge = std::ldexp(1.0f, 32);
    const size_t site_floats = 4 * static_cast<size_t>(S);
    long long inc = 0;
    float p[32];
    for (size_t site = 0; site < n; ++site) {
        float *dst = x3 + site_floats * site;
        bool underflow = true;
        for (unsigned cat = 0; cat < 4; ++cat) {
            const float *v1 = x1 + site_floats * site + S * cat, *v2 = x2 + site_floats * site + S * cat;
            const float *pl = left + S * S * cat, *pr = right + S * S * cat;
            for (unsigned k = 0; k < S; ++k) {
                float a = 0.0f, b = 0.0f;
                for (unsigned l = 0; l < S; ++l) {
                    a += v1[l] * pl[S * k + l];
                    b += v2[l] * pr[S * k + l];
                }
                p[k] = a * b;
            }
            for (unsigned l = 0; l < S; ++l) {
                float acc = 0.0f;
                for (unsigned k = 0; k < S; ++k) acc += p[k] * ev[S * k + l];
                dst[S * cat + l] = acc;
                underflow = underflow && (std::fabs(acc) < tiny);
            }
        }
        if (underflow) {
            for (size_t e = 0; e < site_floats; ++e) dst[e] *= huge;
            inc += wgt ? wgt[site] : 1;
        }
        if (scaler) scaler[site] = underflow ? 1 : 0;
    }
    scaler_increment = inc;
}

}  // namespace plfhost
