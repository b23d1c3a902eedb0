// rapt_quad.cuh -- the field-independent numerics behind flutils.halfbouncepath / flutils.eye
// (rapt/flutils.py:65-151, 254-316): everything the reference delegates to scipy on a traced field line,
// restated so that one thread can do it for its own curve:
//   * scipy.interpolate.interp1d(kind='quadratic')  = make_interp_spline(k=2): quadratic B-spline with knots at
//     the interior midpoints, tridiagonal collocation                                  (SplineView)
//   * scipy.optimize.brentq (xtol 2e-12, rtol 4 eps, maxiter 100)                       (brentq_spline)
//   * scipy.integrate.quad(..., epsrel=1e-4) = QUADPACK QAGS: 21-point Gauss-Kronrod, bisection of the
//     worst interval, Wynn epsilon extrapolation; epsabs 1.49e-8, limit 50            (qags)
//   * scipy.integrate.simpson(y, x=x) for irregular spacing, numpy's pairwise summation (simpson_irregular)
// The QUADPACK and Brent routines are third-party to the reference (scipy, not vendored); they follow the
// published algorithms (Piessens et al., QUADPACK 1983: dqagse/dqk21/dqelg/dqpsrt; Brent 1973 as coded in
// scipy/optimize/Zeros/brentq.c) and are pinned against scipy itself in tests/test_quad_host.py, which
// compiles THIS header for the host (tests/hostcheck/) -- test infrastructure only, never linked into
// librapt_b200.so.
//
// A curve is rows of stride 5: s at [0], x,y,z at [1..3], |B| at [4] (k_bounce_setup's layout).
#pragma once
#ifndef __CUDACC_RTC__
#include <math.h>
#include <float.h>
#endif

#ifndef RAPT_NS
#define RAPT_NS rapt_fast
#endif
#ifndef RAPT_HD
#if defined(__CUDACC_RTC__)
#define RAPT_HD __device__
#elif defined(__CUDACC__)
#define RAPT_HD __host__ __device__
#else
#define RAPT_HD
#endif
#endif

namespace RAPT_NS {

// The QUADPACK routines are real functions on the device, one copy each, with their own stack frames: inlined into
// one frame, nvcc 12.9 let qk21's fv1/fv2 share storage with the live extrapolation table rlist2 of qags (measured on
// sm_100a: rlist2[1..2] came back holding integrand samples; the host build of the same source is ASan/UBSan clean).
#if defined(__CUDACC__)
#define RAPT_QUAD_NOINLINE __noinline__
#else
#define RAPT_QUAD_NOINLINE
#endif

#define RAPT_QUAD_PI 3.141592653589793
#define RAPT_QUAD_LIMIT 50

RAPT_HD inline double quad_nan() { return sqrt(-1.0); }

// ------------------------------------------------------------------------------------------------
// make_interp_spline(s, b, k=2) on the kept part of a curve
// ------------------------------------------------------------------------------------------------
struct SplineView {
    const double *cv;     // curve rows (stride 5): s at [0], |B| at [4]
    double *w;            // work rows (stride 4): [0] spline coefficient c, [1], [2] Thomas scratch
    long long i1;         // first kept point
    int m;                // number of kept points
    RAPT_HD double s(int j) const { return cv[5 * (i1 + j)]; }
    RAPT_HD double b(int j) const { return cv[5 * (i1 + j) + 4]; }
    // knot vector of make_interp_spline(k=2): s0 x3, interior midpoints (without first and last), s_{m-1} x3
    RAPT_HD double knot(int k) const
    {
        if (k <= 2) return s(0);
        if (k >= m) return s(m - 1);
        return 0.5 * (s(k - 2) + s(k - 1));
    }
    RAPT_HD double &c(int j) const { return w[4 * j]; }
    // the three non-zero quadratic B-splines of span l = p + 2 at x (columns p, p+1, p+2)
    RAPT_HD void basis(int p, double x, double &nlo, double &nmid, double &nhi) const
    {
        const int l = p + 2;
        const double tl1 = knot(l - 1), tl = knot(l), tr = knot(l + 1), tr2 = knot(l + 2);
        nlo = (tr - x) * (tr - x) / ((tr - tl1) * (tr - tl));
        nhi = (x - tl) * (x - tl) / ((tr2 - tl) * (tr - tl));
        nmid = 1 - nlo - nhi;
    }
    RAPT_HD double eval(int p, double x) const
    {
        double a, bm, h; basis(p, x, a, bm, h);
        return a * c(p) + bm * c(p + 1) + h * c(p + 2);
    }
    // span p in [0, m-3] with knot(p+2) <= x < knot(p+3) (BSpline's interval rule; the last span is closed)
    RAPT_HD int span(double x) const
    {
        int lo = 0, hi = m - 3;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (knot(mid + 2) <= x) lo = mid; else hi = mid - 1;
        }
        return lo;
    }
    RAPT_HD double at(double x) const { return eval(span(x), x); }
    // tridiagonal collocation system (row j: columns j-1, j, j+1), Thomas algorithm
    RAPT_HD void build() const
    {
        w[1] = 0.0; w[2] = b(0);                                         // row 0: c_0 = b_0
        for (int j = 1; j <= m - 2; j++) {
            double lo, mid, hi; basis(j - 1, s(j), lo, mid, hi);
            const double den = mid - lo * w[4 * (j - 1) + 1];
            w[4 * j + 1] = hi / den;
            w[4 * j + 2] = (b(j) - lo * w[4 * (j - 1) + 2]) / den;
        }
        c(m - 1) = b(m - 1);                                             // last row: c_{m-1} = b_{m-1}
        for (int j = m - 2; j >= 0; j--) c(j) = w[4 * j + 2] - w[4 * j + 1] * c(j + 1);
    }
};

// integrands of the two quadratures: kind 0 = sqrt(1 - B(s)/Bm) (eye, flutils.py:149),
// kind 1 = 1/sqrt(1 - B(s)/Bm) (halfbouncepath, flutils.py:314), kind 2 = B(s) - Bm (the brentq target)
struct MirrorIntegrand {
    SplineView sp;
    double Bm;
    int kind;
    RAPT_HD double operator()(double x) const
    {
        const double B = sp.at(x);
        if (kind == 2) return B - Bm;
        const double u = sqrt(1 - B / Bm);
        return kind ? 1 / u : u;
    }
};

// ------------------------------------------------------------------------------------------------
// scipy.optimize.brentq(f, xa, xb) with the Python-level defaults xtol = 2e-12, rtol = 4*eps, maxiter = 100.
// Returns NaN when f(xa), f(xb) have the same sign (ValueError in scipy).
// ------------------------------------------------------------------------------------------------
template <class Fn>
RAPT_HD double brentq(const Fn &f, double xa, double xb, int *funcalls = nullptr)
{
    const double xtol = 2e-12, rtol = 8.881784197001252e-16;
    double xpre = xa, xcur = xb, xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis, delta, stry, dpre, dblk;
    int calls = 2;
    fpre = f(xpre); fcur = f(xcur);
    if (funcalls) *funcalls = calls;
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) return quad_nan();
    for (int i = 0; i < 100; i++) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) { xblk = xpre; fblk = fpre; spre = scur = xcur - xpre; }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) break;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) stry = -fcur * (xcur - xpre) / (fcur - fpre);                      // secant
            else {                                                                                // inverse quadratic
                dpre = (fpre - fcur) / (xpre - xcur);
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else { spre = sbis; scur = sbis; }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = f(xcur);
        calls++;
    }
    if (funcalls) *funcalls = calls;
    return xcur;
}

// ------------------------------------------------------------------------------------------------
// QUADPACK, double precision.  Arrays are 1-based as in the Fortran so the index arithmetic reads the same.
// ------------------------------------------------------------------------------------------------
struct QagsOut { double result, abserr; int neval, ier, last; };

#define RAPT_QK_EPMACH 2.220446049250313e-16
#define RAPT_QK_UFLOW 2.2250738585072014e-308
#define RAPT_QK_OFLOW 1.7976931348623157e+308

// dqk21: 21-point Gauss-Kronrod rule on [a, b]
template <class Fn>
RAPT_HD RAPT_QUAD_NOINLINE void qk21(const Fn &f, double a, double b, double &result, double &abserr, double &resabs, double &resasc)
{
    const double wg[5] = {0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
                          0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
                          0.295524224714752870173815619188769};
    const double xgk[11] = {0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
                            0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
                            0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
                            0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
                            0.294392862701460198131126603103866, 0.148874338981631210884826001129720, 0.0};
    const double wgk[11] = {0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
                            0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
                            0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
                            0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
                            0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
                            0.149445554002916905664936468389821};
    double fv1[10], fv2[10];
    const double centr = 0.5 * (a + b), hlgth = 0.5 * (b - a), dhlgth = fabs(hlgth);
    double resg = 0.0;
    const double fc = f(centr);
    double resk = wgk[10] * fc;
    resabs = fabs(resk);
    for (int j = 0; j < 5; j++) {
        const int jtw = 2 * j + 1;
        const double absc = hlgth * xgk[jtw];
        const double fval1 = f(centr - absc), fval2 = f(centr + absc);
        fv1[jtw] = fval1; fv2[jtw] = fval2;
        const double fsum = fval1 + fval2;
        resg += wg[j] * fsum;
        resk += wgk[jtw] * fsum;
        resabs += wgk[jtw] * (fabs(fval1) + fabs(fval2));
    }
    for (int j = 0; j < 5; j++) {
        const int jtwm1 = 2 * j;
        const double absc = hlgth * xgk[jtwm1];
        const double fval1 = f(centr - absc), fval2 = f(centr + absc);
        fv1[jtwm1] = fval1; fv2[jtwm1] = fval2;
        const double fsum = fval1 + fval2;
        resk += wgk[jtwm1] * fsum;
        resabs += wgk[jtwm1] * (fabs(fval1) + fabs(fval2));
    }
    const double reskh = resk * 0.5;
    resasc = wgk[10] * fabs(fc - reskh);
    for (int j = 0; j < 10; j++) resasc += wgk[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
    result = resk * hlgth;
    resabs *= dhlgth;
    resasc *= dhlgth;
    abserr = fabs((resk - resg) * hlgth);
    if (resasc != 0.0 && abserr != 0.0) abserr = resasc * fmin(1.0, pow(200.0 * abserr / resasc, 1.5));
    if (resabs > RAPT_QK_UFLOW / (50.0 * RAPT_QK_EPMACH)) abserr = fmax((RAPT_QK_EPMACH * 50.0) * resabs, abserr);
}

// dqelg: Wynn's epsilon algorithm on the table epstab[1..n] (room for n + 2)
RAPT_HD RAPT_QUAD_NOINLINE inline void qelg(int &n, double *epstab, double &result, double &abserr, double *res3la, int &nres)
{
    const int limexp = 50;
    nres++;
    abserr = RAPT_QK_OFLOW;
    result = epstab[n];
    if (n >= 3) {
        epstab[n + 2] = epstab[n];
        const int newelm = (n - 1) / 2;
        epstab[n] = RAPT_QK_OFLOW;
        const int num = n;
        int k1 = n;
        bool converged = false;
        for (int i = 1; i <= newelm; i++) {
            const int k2 = k1 - 1, k3 = k1 - 2;
            double res = epstab[k1 + 2];
            const double e0 = epstab[k3], e1 = epstab[k2], e2 = res;
            const double e1abs = fabs(e1), delta2 = e2 - e1, err2 = fabs(delta2), tol2 = fmax(fabs(e2), e1abs) * RAPT_QK_EPMACH;
            const double delta3 = e1 - e0, err3 = fabs(delta3), tol3 = fmax(e1abs, fabs(e0)) * RAPT_QK_EPMACH;
            if (!(err2 > tol2 || err3 > tol3)) {
                // e0, e1 and e2 equal to within machine accuracy: convergence
                result = res;
                abserr = err2 + err3;
                abserr = fmax(abserr, 5.0 * RAPT_QK_EPMACH * fabs(result));
                converged = true;
                break;
            }
            const double e3 = epstab[k1];
            epstab[k1] = e1;
            const double delta1 = e1 - e3, err1 = fabs(delta1), tol1 = fmax(e1abs, fabs(e3)) * RAPT_QK_EPMACH;
            if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) { n = i + i - 1; break; }
            const double ss = 1.0 / delta1 + 1.0 / delta2 - 1.0 / delta3;
            const double epsinf = fabs(ss * e1);
            if (!(epsinf > 1e-4)) { n = i + i - 1; break; }
            res = e1 + 1.0 / ss;
            epstab[k1] = res;
            k1 -= 2;
            const double error = err2 + fabs(res - e2) + err3;
            if (!(error > abserr)) { abserr = error; result = res; }
        }
        if (converged) return;                     // label 100 was already applied above
        if (n == limexp) n = 2 * (limexp / 2) - 1;
        int ib = ((num / 2) * 2 == num) ? 2 : 1;
        const int ie = newelm + 1;
        for (int i = 1; i <= ie; i++) { const int ib2 = ib + 2; epstab[ib] = epstab[ib2]; ib = ib2; }
        if (num != n) {
            int indx = num - n + 1;
            for (int i = 1; i <= n; i++) { epstab[i] = epstab[indx]; indx++; }
        }
        if (nres < 4) { res3la[nres] = result; abserr = RAPT_QK_OFLOW; }
        else {
            abserr = fabs(result - res3la[3]) + fabs(result - res3la[2]) + fabs(result - res3la[1]);
            res3la[1] = res3la[2]; res3la[2] = res3la[3]; res3la[3] = result;
        }
    }
    abserr = fmax(abserr, 5.0 * RAPT_QK_EPMACH * fabs(result));
}

// dqpsrt: keep iord[] ordered by decreasing error estimate
RAPT_HD RAPT_QUAD_NOINLINE inline void qpsrt(int limit, int last, int &maxerr, double &ermax, const double *elist, int *iord, int &nrmax)
{
    if (last <= 2) { iord[1] = 1; iord[2] = 2; }
    else {
        const double errmax = elist[maxerr];
        if (nrmax != 1) {
            const int ido = nrmax - 1;
            for (int i = 1; i <= ido; i++) {
                const int isucc = iord[nrmax - 1];
                if (errmax <= elist[isucc]) break;
                iord[nrmax] = isucc;
                nrmax--;
            }
        }
        int jupbn = last;
        if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
        const double errmin = elist[last];
        const int jbnd = jupbn - 1, ibeg = nrmax + 1;
        int i = ibeg;
        bool found = false;
        for (; i <= jbnd; i++) {
            const int isucc = iord[i];
            if (errmax >= elist[isucc]) { found = true; break; }
            iord[i - 1] = isucc;
        }
        if (!found) { iord[jbnd] = maxerr; iord[jupbn] = last; }
        else {
            iord[i - 1] = maxerr;
            int k = jbnd;
            bool placed = false;
            for (int j = i; j <= jbnd; j++) {
                const int isucc = iord[k];
                if (errmin < elist[isucc]) { iord[k + 1] = last; placed = true; break; }
                iord[k + 1] = isucc;
                k--;
            }
            if (!placed) iord[i] = last;
        }
    }
    maxerr = iord[nrmax];
    ermax = elist[maxerr];
}

// dqagse with limit = 50 (scipy.integrate.quad's default); epsabs / epsrel as given
template <class Fn>
RAPT_HD RAPT_QUAD_NOINLINE QagsOut qags(const Fn &f, double a, double b, double epsabs, double epsrel, double *dbg = nullptr)
{
    const int limit = RAPT_QUAD_LIMIT;
    double alist[RAPT_QUAD_LIMIT + 1], blist[RAPT_QUAD_LIMIT + 1], rlist[RAPT_QUAD_LIMIT + 1], elist[RAPT_QUAD_LIMIT + 1];
    double rlist2[53], res3la[4];
    int iord[RAPT_QUAD_LIMIT + 1];
    QagsOut o; o.result = 0; o.abserr = 0; o.neval = 0; o.ier = 0; o.last = 0;
    double result = 0.0, abserr = 0.0;
    int ier = 0, last = 0;
    alist[1] = a; blist[1] = b; rlist[1] = 0.0; elist[1] = 0.0;
    if (epsabs <= 0.0 && epsrel < fmax(50.0 * RAPT_QK_EPMACH, 0.5e-28)) { o.ier = 6; return o; }
    int ierro = 0;
    double defabs, resabs;
    qk21(f, a, b, result, abserr, defabs, resabs);
    double dres = fabs(result);
    double errbnd = fmax(epsabs, epsrel * dres);
    last = 1;
    rlist[1] = result; elist[1] = abserr; iord[1] = 1;
    if (abserr <= 100.0 * RAPT_QK_EPMACH * defabs && abserr > errbnd) ier = 2;
    if (limit == 1) ier = 1;
    if (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0) {
        o.result = result; o.abserr = abserr; o.ier = ier; o.last = last; o.neval = 42 * last - 21;
        return o;
    }
    rlist2[1] = result;
    double errmax = abserr, area = result, errsum = abserr;
    int maxerr = 1, nrmax = 1, nres = 0, numrl2 = 2, ktmin = 0, iroff1 = 0, iroff2 = 0, iroff3 = 0;
    bool extrap = false, noext = false;
    abserr = RAPT_QK_OFLOW;
    int ksgn = -1;
    if (dres >= (1.0 - 50.0 * RAPT_QK_EPMACH) * defabs) ksgn = 1;
    double small = 0, erlarg = 0, ertest = 0, correc = 0, erlast, reseps, abseps;
    int exit_code = 0;       // 0: loop exhausted / label 100, 115: sum the pieces
    for (last = 2; last <= limit; last++) {
        const double a1 = alist[maxerr], b1 = 0.5 * (alist[maxerr] + blist[maxerr]), a2 = b1, b2 = blist[maxerr];
        erlast = errmax;
        double area1, error1, defab1, area2, error2, defab2;
        qk21(f, a1, b1, area1, error1, resabs, defab1);
        qk21(f, a2, b2, area2, error2, resabs, defab2);
        const double area12 = area1 + area2, erro12 = error1 + error2;
        errsum = errsum + erro12 - errmax;
        area = area + area12 - rlist[maxerr];
        if (!(defab1 == error1 || defab2 == error2)) {
            if (!(fabs(rlist[maxerr] - area12) > 1e-5 * fabs(area12) || erro12 < 0.99 * errmax)) {
                if (extrap) iroff2++; else iroff1++;
            }
            if (last > 10 && erro12 > errmax) iroff3++;
        }
        rlist[maxerr] = area1;
        rlist[last] = area2;
        errbnd = fmax(epsabs, epsrel * fabs(area));
        if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ier = 2;
        if (iroff2 >= 5) ierro = 3;
        if (last == limit) ier = 1;
        if (fmax(fabs(a1), fabs(b2)) <= (1.0 + 100.0 * RAPT_QK_EPMACH) * (fabs(a2) + 1000.0 * RAPT_QK_UFLOW)) ier = 4;
        if (error2 > error1) {
            alist[maxerr] = a2; alist[last] = a1; blist[last] = b1;
            rlist[maxerr] = area2; rlist[last] = area1;
            elist[maxerr] = error2; elist[last] = error1;
        } else {
            alist[last] = a2; blist[maxerr] = b1; blist[last] = b2;
            elist[maxerr] = error1; elist[last] = error2;
        }
        qpsrt(limit, last, maxerr, errmax, elist, iord, nrmax);
        if (dbg) { double *g = dbg + 12 * (last - 2); g[0] = a1; g[1] = b2; g[2] = area1; g[3] = area2; g[4] = error1; g[5] = error2;
                   g[6] = errsum; g[7] = errbnd; g[8] = maxerr; g[9] = nrmax; g[10] = extrap; g[11] = ier; }
        if (errsum <= errbnd) { exit_code = 115; break; }
        if (ier != 0) break;
        if (last == 2) {
            small = fabs(b - a) * 0.375;
            erlarg = errsum;
            ertest = errbnd;
            rlist2[2] = area;
            continue;
        }
        if (noext) continue;
        erlarg -= erlast;
        if (fabs(b1 - a1) > small) erlarg += erro12;
        if (!extrap) {
            // is the interval to be bisected next the smallest one?
            if (fabs(blist[maxerr] - alist[maxerr]) > small) continue;
            extrap = true;
            nrmax = 2;
        }
        if (ierro != 3 && erlarg > ertest) {
            // the smallest interval has the largest error: first work on the larger ones
            const int id = nrmax;
            int jupbnd = last;
            if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
            bool again = false;
            for (int k = id; k <= jupbnd; k++) {
                maxerr = iord[nrmax];
                errmax = elist[maxerr];
                if (fabs(blist[maxerr] - alist[maxerr]) > small) { again = true; break; }
                nrmax++;
            }
            if (again) continue;
        }
        // extrapolate
        numrl2++;
        rlist2[numrl2] = area;
        if (dbg) { double *g = dbg + 12 * 50 + 12 * (last - 2); g[0] = numrl2; g[1] = rlist2[numrl2]; g[2] = rlist2[numrl2 - 1]; g[3] = rlist2[numrl2 > 2 ? numrl2 - 2 : 1]; g[7] = erlarg; g[8] = ertest; g[9] = small; }
        qelg(numrl2, rlist2, reseps, abseps, res3la, nres);
        if (dbg) { double *g = dbg + 12 * 50 + 12 * (last - 2); g[4] = reseps; g[5] = abseps; g[6] = nres; g[10] = abserr; g[11] = numrl2; }
        ktmin++;
        if (ktmin > 5 && abserr < 1e-3 * errsum) ier = 5;
        if (abseps < abserr) {
            ktmin = 0;
            abserr = abseps;
            result = reseps;
            correc = erlarg;
            ertest = fmax(epsabs, epsrel * fabs(reseps));
            if (abserr <= ertest) break;
        }
        if (numrl2 == 1) noext = true;
        if (ier == 5) break;
        maxerr = iord[1];
        errmax = elist[maxerr];
        nrmax = 1;
        extrap = false;
        small *= 0.5;
        erlarg = errsum;
    }
    if (last > limit) last = limit;          // the Fortran DO variable after a completed loop is not used past here
    bool sum_pieces = (exit_code == 115);
    if (!sum_pieces) {
        // label 100
        if (abserr == RAPT_QK_OFLOW) sum_pieces = true;
        else {
            bool to110 = (ier + ierro == 0);
            bool to130 = false;
            if (!to110) {
                if (ierro == 3) abserr += correc;
                if (ier == 0) ier = 3;
                if (result != 0.0 && area != 0.0) {
                    if (abserr / fabs(result) > errsum / fabs(area)) sum_pieces = true; else to110 = true;
                } else {
                    if (abserr > errsum) sum_pieces = true;
                    else if (area == 0.0) to130 = true;
                    else to110 = true;
                }
            }
            if (!sum_pieces && to110 && !to130) {
                if (ksgn == -1 && fmax(fabs(result), fabs(area)) <= defabs * 0.01) { /* -> 130 */ }
                else if (0.01 > (result / area) || (result / area) > 100.0 || errsum > fabs(area)) ier = 6;
            }
        }
    }
    if (sum_pieces) {
        result = 0.0;
        for (int k = 1; k <= last; k++) result += rlist[k];
        abserr = errsum;
    }
    if (ier > 2) ier--;
    o.result = result; o.abserr = abserr; o.ier = ier; o.last = last; o.neval = 42 * last - 21;
    return o;
}

// ------------------------------------------------------------------------------------------------
// numpy's pairwise summation of term(0..n-1) (numpy/_core/src/umath/loops_utils.h.src, blocks of 128,
// eight running sums) -- what np.sum does inside scipy.integrate.simpson
// ------------------------------------------------------------------------------------------------
template <class Term>
RAPT_HD double np_block_sum(const Term &term, int off, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res += term(off + i);
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = term(off + j);
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] += term(off + i + j);
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += term(off + i);
    return res;
}

template <class Term>
RAPT_HD double np_pairwise_sum(const Term &term, int n)
{
    if (n <= 128) return np_block_sum(term, 0, n);
    // explicit stack for the recursion sum(a, n) = sum(a, n2) + sum(a + n2, n - n2), n2 = n/2 rounded down to 8
    int off[24], len[24], state[24];
    double left[24];
    int sp = 0;
    off[0] = 0; len[0] = n; state[0] = 0;
    double ret = 0;
    while (sp >= 0) {
        if (len[sp] <= 128) { ret = np_block_sum(term, off[sp], len[sp]); sp--; continue; }
        int n2 = len[sp] / 2; n2 -= n2 % 8;
        if (state[sp] == 0) { state[sp] = 1; off[sp + 1] = off[sp]; len[sp + 1] = n2; state[sp + 1] = 0; sp++; }
        else if (state[sp] == 1) { left[sp] = ret; state[sp] = 2; off[sp + 1] = off[sp] + n2; len[sp + 1] = len[sp] - n2; state[sp + 1] = 0; sp++; }
        else { ret = left[sp] + ret; sp--; }
    }
    return ret;
}

// scipy.integrate.simpson(y, x=x) (scipy/integrate/_quadrature.py, scipy >= 1.11: Cartwright's correction of
// the last interval when the number of points is even), y and x given as functors of the point index
template <class Y, class X>
RAPT_HD double simpson_irregular(const Y &y, const X &x, int N)
{
    auto basic = [&](int stop) {       // _basic_simpson(y, 0, stop, x): panels starting at 0, 2, ..., < stop
        const int np_ = (stop + 1) / 2;
        auto term = [&](int k) {
            const int i = 2 * k;
            const double h0 = x(i + 1) - x(i), h1 = x(i + 2) - x(i + 1);
            const double hsum = h0 + h1, hprod = h0 * h1;
            const double h0divh1 = (h1 != 0) ? h0 / h1 : 0.;
            return hsum / 6.0 * (y(i) * (2.0 - ((h0divh1 != 0) ? 1.0 / h0divh1 : 0.))
                                 + y(i + 1) * (hsum * ((hprod != 0) ? hsum / hprod : 0.))
                                 + y(i + 2) * (2.0 - h0divh1));
        };
        return np_pairwise_sum(term, np_);
    };
    if (N % 2 == 0) {
        double val = 0.0, result = 0.0;
        if (N == 2) val = 0.5 * (x(1) - x(0)) * (y(1) + y(0));
        else {
            result = basic(N - 3);
            const double h0 = x(N - 2) - x(N - 3), h1 = x(N - 1) - x(N - 2);
            double num = 2 * (h1 * h1) + 3 * h0 * h1, den = 6 * (h1 + h0);
            const double alpha = (den != 0) ? num / den : 0.;
            num = (h1 * h1) + 3.0 * h0 * h1; den = 6 * h0;
            const double beta = (den != 0) ? num / den : 0.;
            num = 1 * (h1 * h1 * h1); den = 6 * h0 * (h0 + h1);
            const double eta = (den != 0) ? num / den : 0.;
            result += alpha * y(N - 1) + beta * y(N - 2) - eta * y(N - 3);
        }
        result += val;
        return result;
    }
    return basic(N - 2);
}

// ------------------------------------------------------------------------------------------------
// flutils.halfbouncepath (flutils.py:274-316) on a traced curve of n points.
//   quadpack = 1: the reference's own route -- interpolating spline, brentq for the two mirror points,
//                 QAGS with epsrel 1e-4 (agrees with the reference to round-off);
//   quadpack = 0: same spline, mirror points and integral in closed form (no quadrature error; differs from
//                 the reference by QUADPACK's own error, 1e-7 typical).
// Returns NaN when the trace does not bracket both mirror points.
// ------------------------------------------------------------------------------------------------

// primitive of 1/sqrt(al + be u + ga u^2)
RAPT_HD inline double invsqrt_quadratic_primitive(double al, double be, double ga, double u)
{
    const double q = fmax(al + be * u + ga * u * u, 0.0);
    if (ga == 0) return 2 * sqrt(q) / be;
    const double disc = be * be - 4 * al * ga;
    if (ga < 0) return -asin(fmax(-1.0, fmin(1.0, (2 * ga * u + be) / sqrt(disc)))) / sqrt(-ga);
    return log(fabs(2 * ga * u + be + 2 * sqrt(ga) * sqrt(q))) / sqrt(ga);
}

RAPT_HD inline double halfbouncepath_closed_form(const SplineView &sp, double Bm)
{
    const int m = sp.m;
    // mirror points: B(s) = Bm in [s_0, s_1] and in [s_{m-2}, s_{m-1}]  (flutils.py:311-313)
    double sm[2] = {quad_nan(), quad_nan()};
    for (int side = 0; side < 2; side++) {
        const double lo = side ? sp.s(m - 2) : sp.s(0), hi = side ? sp.s(m - 1) : sp.s(1);
        for (int p = 0; p <= m - 3; p++) {
            const double sl = sp.knot(p + 2), sr = sp.knot(p + 3);
            const double L = fmax(sl, lo), R = fmin(sr, hi);
            if (!(L < R)) continue;
            const double hh = sr - sl, f0 = sp.eval(p, sl), f1 = sp.eval(p, 0.5 * (sl + sr)), f2 = sp.eval(p, sr);
            const double cc = 2 * (f2 - 2 * f1 + f0) / (hh * hh), bb = (f2 - f0) / hh - cc * hh, aa = f0 - Bm;
            double r0, r1;
            if (cc == 0) { r0 = r1 = -aa / bb; }
            else {
                const double d = bb * bb - 4 * cc * aa;
                if (d < 0) continue;
                const double qq = -0.5 * (bb + copysign(sqrt(d), bb));
                r0 = qq / cc; r1 = (qq != 0) ? aa / qq : quad_nan();
            }
            const double eps = 1e-9 * hh;
            if (r0 >= L - sl - eps && r0 <= R - sl + eps) { sm[side] = sl + r0; break; }
            if (r1 >= L - sl - eps && r1 <= R - sl + eps) { sm[side] = sl + r1; break; }
        }
    }
    if (!(sm[0] == sm[0]) || !(sm[1] == sm[1])) return quad_nan();
    // S_b = integral_{sm1}^{sm2} ds / sqrt(1 - B(s)/Bm), span by span in closed form
    double tot = 0;
    for (int p = 0; p <= m - 3; p++) {
        const double sl = sp.knot(p + 2), sr = sp.knot(p + 3);
        const double L = fmax(sl, sm[0]), R = fmin(sr, sm[1]);
        if (!(L < R)) continue;
        const double hh = sr - sl, f0 = sp.eval(p, sl), f1 = sp.eval(p, 0.5 * (sl + sr)), f2 = sp.eval(p, sr);
        const double cc = 2 * (f2 - 2 * f1 + f0) / (hh * hh), bb = (f2 - f0) / hh - cc * hh;
        const double al = 1 - f0 / Bm, be = -bb / Bm, ga = -cc / Bm;
        tot += invsqrt_quadratic_primitive(al, be, ga, R - sl) - invsqrt_quadratic_primitive(al, be, ga, L - sl);
    }
    return tot;
}

RAPT_HD inline double halfbouncepath_curve(const double *cv, double *work, long long n, double Bm, int quadpack)
{
    long long first = -1, last = -1;
    for (long long k = 0; k < n; k++) if (cv[5 * k + 4] <= Bm) { if (first < 0) first = k; last = k; }
    long long i1, i2;
    if (first < 0) { i1 = (n - 3) / 2; i2 = (n + 1) / 2; }          // flutils.py:281-283
    else { i1 = first - 1; i2 = last + 1; }
    if (i1 < 0 || i2 > n - 1) return quad_nan();
    const int m = (int)(i2 - i1 + 1);
    if (m < 3) return quad_nan();
    if (m == 3) {                                                    // flutils.py:295-305
        const double s1 = cv[5 * i1], s2 = cv[5 * (i1 + 1)], s3 = cv[5 * (i1 + 2)];
        const double B1 = cv[5 * i1 + 4], B2 = cv[5 * (i1 + 1) + 4], B3 = cv[5 * (i1 + 2) + 4];
        const double s12 = s1 - s2, s23 = s2 - s3, s13 = s1 - s3;
        const double B2s = 2 * (B1 * s23 - B2 * s13 + B3 * s12) / (s12 * s13 * s23);
        return RAPT_QUAD_PI * sqrt(2 * Bm / B2s);
    }
    SplineView sp = {cv, work, i1, m};
    sp.build();
    if (!quadpack) return halfbouncepath_closed_form(sp, Bm);
    MirrorIntegrand root = {sp, Bm, 2};
    const double sm1 = brentq(root, sp.s(0), sp.s(1));                      // flutils.py:311
    const double sm2 = brentq(root, sp.s(m - 2), sp.s(m - 1));              // flutils.py:313
    if (!(sm1 == sm1) || !(sm2 == sm2)) return quad_nan();
    MirrorIntegrand g = {sp, Bm, 1};
    return qags(g, sm1, sm2, 1.49e-8, 1e-4).result;                         // flutils.py:314
}

// ------------------------------------------------------------------------------------------------
// flutils.eye (flutils.py:65-151) on a traced curve: the second invariant I = integral of sqrt(1 - B/Bm) ds
// between the mirror points.  Equatorial pitch angle >= 70 deg: spline + brentq + QAGS.  Below 70 deg the
// reference calls `simps`, a name it never imports (flutils.py:130: NameError); as in GuidingCenter.geteye
// here, that branch is evaluated with scipy.integrate.simpson(y, x=x), which is what the import line
// (flutils.py:18) provides.  *err = 1 where the reference's `assert` (flutils.py:117) would fail.
// ------------------------------------------------------------------------------------------------
RAPT_HD inline double eye_curve(const double *cv, double *work, long long n, double Bm, int *err)
{
    *err = 0;
    double Bmin = cv[4];
    for (long long k = 1; k < n; k++) Bmin = fmin(Bmin, cv[5 * k + 4]);
    if (Bmin > Bm) return 0;
    if (fabs(Bmin - Bm) / Bm < 1e-12) return 0;
    long long first = -1, last = -1;
    for (long long k = 0; k < n; k++) if (cv[5 * k + 4] < Bm) { if (first < 0) first = k; last = k; }
    // np.delete(range(0, inside[0]-1) + range(inside[-1]+2, n)): keep [max(first-1, 0), min(last+1, n-1)]
    const long long i1 = first - 1 > 0 ? first - 1 : 0, i2 = last + 1 < n - 1 ? last + 1 : n - 1;
    const int m = (int)(i2 - i1 + 1);
    auto S = [&](int j) { return cv[5 * (i1 + j)]; };
    auto Bv = [&](int j) { return cv[5 * (i1 + j) + 4]; };
    if (m < 4 || !(Bv(0) > Bm && Bv(1) < Bm && Bv(m - 2) < Bm && Bv(m - 1) > Bm)) { *err = 1; return quad_nan(); }
    const double eqpa = asin(sqrt(Bmin / Bm)) * 180 / RAPT_QUAD_PI;
    if (eqpa < 70) {
        const double sm1 = (Bm - Bv(0)) * (S(1) - S(0)) / (Bv(1) - Bv(0)) + S(0);
        const double sm2 = (Bm - Bv(m - 2)) * (S(m - 1) - S(m - 2)) / (Bv(m - 1) - Bv(m - 2)) + S(m - 2);
        auto yy = [&](int j) { return sqrt(1 - Bv(j + 1) / Bm); };
        auto xx = [&](int j) { return S(j + 1); };
        double I = simpson_irregular(yy, xx, m - 2);
        double d = sm2 - S(m - 2);
        I += (2.0 / 3.0) * d * sqrt((Bm - Bv(m - 2)) / Bm);
        d = S(1) - sm1;
        I += (2.0 / 3.0) * d * sqrt((Bm - Bv(1)) / Bm);
        return I;
    }
    SplineView sp = {cv, work, i1, m};
    sp.build();
    MirrorIntegrand root = {sp, Bm, 2};
    const double sm1 = brentq(root, sp.s(0), sp.s(1));
    double sm2;
    if (sp.at(sp.s(m - 2)) == Bm) sm2 = sp.s(m - 2);
    else sm2 = brentq(root, sp.s(m - 2), sp.s(m - 1));
    if (!(sm1 == sm1) || !(sm2 == sm2)) { *err = 1; return quad_nan(); }
    MirrorIntegrand g = {sp, Bm, 0};
    return qags(g, sm1, sm2, 1.49e-8, 1e-4).result;
}

}  // namespace RAPT_NS
