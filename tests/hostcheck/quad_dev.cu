// TEST INFRASTRUCTURE ONLY.  The DEVICE build of rapt_b200/csrc/rapt_quad.cuh behind the same C entry points as
// quad_host.cpp (one thread per call), so that tests/test_gpu_quad.py can run the scipy comparisons of
// tests/test_quad_host.py against the code the kernels actually execute.  Not part of librapt_b200.so.
#include <cuda_runtime.h>
#include <vector>
#define RAPT_NS rapt_devcheck
#include "../../rapt_b200/csrc/rapt_quad.cuh"
using namespace rapt_devcheck;

struct TestFn {
    int id; double p;
    __host__ __device__ double operator()(double x) const
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

__global__ void k_qags(int id, double p, double a, double b, double epsabs, double epsrel, double *out)
{
    TestFn f = {id, p};
    QagsOut o = qags(f, a, b, epsabs, epsrel);
    out[0] = o.result; out[1] = o.abserr; out[2] = o.neval; out[3] = o.ier; out[4] = o.last;
}
__global__ void k_brentq(int id, double p, double a, double b, double *out)
{
    TestFn f = {id, p};
    int calls = 0;
    out[0] = brentq(f, a, b, &calls);
    out[1] = calls;
}
// what = 0: halfbouncepath (quadpack), 1: halfbouncepath (closed form), 2: eye; out[1] = err flag
__global__ void k_curve(const double *cv, double *w, long long n, double Bm, int what, double *out)
{
    int err = 0;
    if (what == 2) out[0] = eye_curve(cv, w, n, Bm, &err);
    else out[0] = halfbouncepath_curve(cv, w, n, Bm, what == 0);
    out[1] = err;
}
// the mirror points and the first Gauss-Kronrod pass, for diagnosis: out = sm1, sm2, calls1, calls2, qk21 result/abserr/resabs/resasc
__global__ void k_parts(const double *cv, double *w, long long i1, int m, double Bm, int kind, double *out)
{
    SplineView sp = {cv, w, i1, m};
    sp.build();
    MirrorIntegrand root = {sp, Bm, 2};
    int c1 = 0, c2 = 0;
    out[0] = brentq(root, sp.s(0), sp.s(1), &c1);
    out[1] = brentq(root, sp.s(m - 2), sp.s(m - 1), &c2);
    out[2] = c1; out[3] = c2;
    MirrorIntegrand g = {sp, Bm, kind};
    qk21(g, out[0], out[1], out[4], out[5], out[6], out[7]);
    QagsOut o = qags(g, out[0], out[1], 1.49e-8, 1e-4, out + 13);
    out[8] = o.result; out[9] = o.abserr; out[10] = o.neval; out[11] = o.ier; out[12] = o.last;
}

template <class K, class... A> static int run1(double *host_out, int nout, K kern, A... args)
{
    double *d = nullptr;
    if (cudaMalloc(&d, nout * sizeof(double)) != cudaSuccess) return -1;
    kern<<<1, 1>>>(args..., d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(host_out, d, nout * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" {
int dv_qags(int id, double p, double a, double b, double epsabs, double epsrel, double *out)
{ return run1(out, 5, k_qags, id, p, a, b, epsabs, epsrel); }
int dv_brentq(int id, double p, double a, double b, double *out)
{ return run1(out, 2, k_brentq, id, p, a, b); }
int dv_curve(const double *s, const double *b, long long n, double Bm, int what, double *out)
{
    std::vector<double> cv(5 * n, 0.0);
    for (long long k = 0; k < n; k++) { cv[5 * k] = s[k]; cv[5 * k + 4] = b[k]; }
    double *dcv = nullptr, *dw = nullptr;
    cudaMalloc(&dcv, 5 * n * sizeof(double)); cudaMalloc(&dw, 4 * n * sizeof(double));
    cudaMemcpy(dcv, cv.data(), 5 * n * sizeof(double), cudaMemcpyHostToDevice);
    int rc = run1(out, 2, k_curve, (const double *)dcv, dw, n, Bm, what);
    cudaFree(dcv); cudaFree(dw);
    return rc;
}
int dv_parts(const double *s, const double *b, long long n, long long i1, int m, double Bm, int kind, double *out)
{
    std::vector<double> cv(5 * n, 0.0);
    for (long long k = 0; k < n; k++) { cv[5 * k] = s[k]; cv[5 * k + 4] = b[k]; }
    double *dcv = nullptr, *dw = nullptr;
    cudaMalloc(&dcv, 5 * n * sizeof(double)); cudaMalloc(&dw, 4 * n * sizeof(double));
    cudaMemcpy(dcv, cv.data(), 5 * n * sizeof(double), cudaMemcpyHostToDevice);
    int rc = run1(out, 13 + 24 * 50, k_parts, (const double *)dcv, dw, i1, m, Bm, kind);
    cudaFree(dcv); cudaFree(dw);
    return rc;
}
// the same on the host (quad_host.cpp has no k_parts counterpart)
void hs_parts(const double *s, const double *b, long long n, long long i1, int m, double Bm, int kind, double *out)
{
    std::vector<double> cv(5 * n, 0.0), w(4 * n);
    for (long long k = 0; k < n; k++) { cv[5 * k] = s[k]; cv[5 * k + 4] = b[k]; }
    SplineView sp = {cv.data(), w.data(), i1, m};
    sp.build();
    MirrorIntegrand root = {sp, Bm, 2};
    int c1 = 0, c2 = 0;
    out[0] = brentq(root, sp.s(0), sp.s(1), &c1);
    out[1] = brentq(root, sp.s(m - 2), sp.s(m - 1), &c2);
    out[2] = c1; out[3] = c2;
    MirrorIntegrand g = {sp, Bm, kind};
    qk21(g, out[0], out[1], out[4], out[5], out[6], out[7]);
    QagsOut o = qags(g, out[0], out[1], 1.49e-8, 1e-4, out + 13);
    out[8] = o.result; out[9] = o.abserr; out[10] = o.neval; out[11] = o.ier; out[12] = o.last;
}
}
