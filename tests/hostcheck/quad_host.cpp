// TEST INFRASTRUCTURE ONLY.  Compiles rapt_b200/csrc/rapt_quad.cuh (the field-independent numerics of the
// bounce-centre path: spline, brentq, QUADPACK QAGS, Simpson) for the HOST so that tests/test_quad_host.py can pin
// the restatement against scipy itself without a GPU.  Nothing in rapt_b200/ loads this library and it is not
// part of librapt_b200.so: the product path is the CUDA build of the same header.
#define RAPT_NS rapt_hostcheck
#include "../../rapt_b200/csrc/rapt_quad.cuh"
#include <vector>
using namespace rapt_hostcheck;

struct TestFn {
    int id; double p;
    double operator()(double x) const
    {
        switch (id) {
        case 0: return sqrt(x);
        case 1: return 1 / sqrt(x);
        case 2: return log(x) / sqrt(x);
        case 3: return 1 / sqrt(fabs(x - p));
        case 4: return cos(p * x) * exp(-x);
        case 5: return 1 / (1 + p * x * x);
        case 6: return pow(x, p);
        case 7: return sqrt(fabs(sin(p * x)));
        case 8: return x * x * x - p;
        case 9: return cos(x) - p * x;
        default: return x;
        }
    }
};

static std::vector<double> curve_of(const double *s, const double *b, long long n)
{
    std::vector<double> cv(5 * n, 0.0);
    for (long long k = 0; k < n; k++) { cv[5 * k] = s[k]; cv[5 * k + 4] = b[k]; }
    return cv;
}

extern "C" {
void hc_qags(int id, double p, double a, double b, double epsabs, double epsrel, double *out)
{
    TestFn f = {id, p};
    QagsOut o = qags(f, a, b, epsabs, epsrel);
    out[0] = o.result; out[1] = o.abserr; out[2] = o.neval; out[3] = o.ier; out[4] = o.last;
}
double hc_brentq(int id, double p, double a, double b, int *calls)
{
    TestFn f = {id, p};
    return brentq(f, a, b, calls);
}
void hc_spline(const double *s, const double *b, int m, const double *x, int nx, double *out, double *coef)
{
    std::vector<double> cv = curve_of(s, b, m), w(4 * m);
    SplineView sp = {cv.data(), w.data(), 0, m};
    sp.build();
    for (int i = 0; i < nx; i++) out[i] = sp.at(x[i]);
    if (coef) for (int j = 0; j < m; j++) coef[j] = sp.c(j);
}
double hc_halfbounce(const double *s, const double *b, long long n, double Bm, int quadpack)
{
    std::vector<double> cv = curve_of(s, b, n), w(4 * n);
    return halfbouncepath_curve(cv.data(), w.data(), n, Bm, quadpack);
}
double hc_eye(const double *s, const double *b, long long n, double Bm, int *err)
{
    std::vector<double> cv = curve_of(s, b, n), w(4 * n);
    return eye_curve(cv.data(), w.data(), n, Bm, err);
}
double hc_simpson(const double *y, const double *x, int N)
{
    auto yy = [&](int j) { return y[j]; };
    auto xx = [&](int j) { return x[j]; };
    return simpson_irregular(yy, xx, N);
}
}
