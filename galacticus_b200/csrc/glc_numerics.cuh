// glc_numerics.cuh -- device numerics used inside the RHS.
//
// GPU-first restructuring: every routine evaluates its user function from exactly ONE call
// site inside a small state machine, so the (large, inlined) physics integrands are instantiated
// once per use instead of once per textual call of the CPU algorithms.  The sequence of points
// evaluated and every decision are those of:
//   rootFinder%find            source/numerical/root_finder.F90:587-1075 (Brent branch)
//   gsl_root_fsolver_brent     libgsl 2.6 roots/brent.c ; gsl_root_test_interval roots/convergence.c
//   gsl_integration_qag, GK15  libgsl 2.6 integration/{qag.c,qk.c,qk15.c} (QUADPACK dqage/dqk15)
//   table1DLinearLinear        source/objects/tables/_module.F90:1237-1405
//   fastExponentiator          source/math/exponentiation.F90:57-104
#pragma once

#include <float.h>

namespace glc {

enum : int { EXPAND_NONE = 0, EXPAND_ADDITIVE = 1, EXPAND_MULTIPLICATIVE = 2 };
enum : int { SIGN_NONE = 0, SIGN_NEGATIVE = -1, SIGN_POSITIVE = 1 };

struct RootOptions {
    double tolAbs, tolRel;
    int expandType;
    double expandUpward, expandDownward;
    int signExpectUpward, signExpectDownward;
};

GLC_DEVICE_INLINE double fsign1(double x) { return signbit(x) ? -1.0 : 1.0; }

// Warp-synchronous calling discipline.  Every routine below that contains a data-dependent loop must be
// called by ALL lanes of the warp from uniform control flow; `on` tells whether this lane really wants the
// result.  The loops run `while (GLC_ANY(lane still busy))`: the vote is an explicit reconvergence point
// in every iteration, lanes that are finished idle through the body.  (Left to itself the compiler does not
// re-converge the multi-exit state machines reliably: measured 3 active lanes per warp instruction.)

// returns the root; status 0 = ok, 2 = could not bracket, 3 = bad bracket, 4 = no convergence
template <class F>
GLC_DEVICE_INLINE double root_find(F &&f, bool on, const RootOptions &o, double xLow, double xHigh,
                                            bool haveValues, double fLow, double fHigh, int &status) {
    enum : int { ST_FLO, ST_FHI, ST_BRACKET, ST_EXP_UP, ST_EXP_DOWN, ST_BRENT };
    int state = haveValues ? ST_BRACKET : ST_FLO;
    bool lowerOk = false, upperOk = false, rangeChanged = false, first = true;
    // Brent state (GSL brent_state_t)
    double a = 0, b = 0, c = 0, d = 0, e = 0, fa = 0, fb = 0, fc = 0;
    double xl = 0, xh = 0, root = 0;
    int iteration = 0;
    bool busy = on;
    double result = 0.0;
    status = 0;
    while (GLC_ANY(busy)) {
        double x = 0.0;
        bool evaluate = false;
        // ---- advance the state machine to the next point at which the function is needed (short, no heavy work)
        while (busy && !evaluate) {
            evaluate = true;
            if (state == ST_FLO) {
                x = xLow;
            } else if (state == ST_FHI) {
                x = xHigh;
            } else if (state == ST_BRACKET) {
                evaluate = false;
                if (first) {
                    if (xHigh == xLow) fHigh = fLow;
                    first = false;
                }
                if (fsign1(fLow) * fsign1(fHigh) > 0.0 && fLow != 0.0 && fHigh != 0.0) {
                    lowerOk = o.signExpectDownward == SIGN_NEGATIVE   ? (fLow < 0.0)
                              : o.signExpectDownward == SIGN_POSITIVE ? (fLow > 0.0)
                                                                      : false;
                    upperOk = o.signExpectUpward == SIGN_NEGATIVE   ? (fHigh < 0.0)
                              : o.signExpectUpward == SIGN_POSITIVE ? (fHigh > 0.0)
                                                                    : false;
                    rangeChanged = false;
                    state = ST_EXP_UP;
                } else {
                    // brent_init (function values at the bracket ends are already known)
                    a = xLow;
                    fa = fLow;
                    b = xHigh;
                    fb = fHigh;
                    c = xHigh;
                    fc = fHigh;
                    d = xHigh - xLow;
                    e = xHigh - xLow;
                    if ((fLow < 0.0 && fHigh < 0.0) || (fLow > 0.0 && fHigh > 0.0)) {
                        status = 3;
                        result = 0.0;
                        busy = false;
                    }
                    state = ST_BRENT;
                }
            } else if (state == ST_EXP_UP) {
                bool move;
                if (o.expandType == EXPAND_ADDITIVE)
                    move = o.expandUpward > 0.0 && !upperOk;
                else if (o.expandType == EXPAND_MULTIPLICATIVE)
                    move = ((o.expandUpward > 1.0 && xHigh > 0.0) || (o.expandUpward < 1.0 && xHigh < 0.0)) && !upperOk;
                else
                    move = false;
                if (move) {
                    if (lowerOk) {
                        xLow = xHigh;
                        fLow = fHigh;
                    }
                    xHigh = (o.expandType == EXPAND_ADDITIVE) ? xHigh + o.expandUpward : xHigh * o.expandUpward;
                    x = xHigh;
                    rangeChanged = true;
                } else {
                    evaluate = false;
                    state = ST_EXP_DOWN;
                }
            } else if (state == ST_EXP_DOWN) {
                bool move;
                if (o.expandType == EXPAND_ADDITIVE)
                    move = o.expandDownward < 0.0 && !lowerOk;
                else if (o.expandType == EXPAND_MULTIPLICATIVE)
                    move = ((o.expandDownward < 1.0 && xLow > 0.0) || (o.expandDownward > 1.0 && xLow < 0.0)) && !lowerOk;
                else
                    move = false;
                if (move) {
                    if (upperOk) {
                        xHigh = xLow;
                        fHigh = fLow;
                    }
                    xLow = (o.expandType == EXPAND_ADDITIVE) ? xLow + o.expandDownward : xLow * o.expandDownward;
                    x = xLow;
                    rangeChanged = true;
                } else {
                    evaluate = false;
                    if (!rangeChanged) {
                        status = 2;
                        result = 0.0;
                        busy = false;
                    }
                    state = ST_BRACKET;
                }
            } else {  // ST_BRENT: brent_iterate up to the point where f(b) is needed
                double tol, m;
                bool acEqual = false;
                iteration++;
                if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) {
                    acEqual = true;
                    c = a;
                    fc = fa;
                    d = b - a;
                    e = b - a;
                }
                if (fabs(fc) < fabs(fb)) {
                    acEqual = true;
                    a = b;
                    b = c;
                    c = a;
                    fa = fb;
                    fb = fc;
                    fc = fa;
                }
                tol = 0.5 * DBL_EPSILON * fabs(b);
                m = 0.5 * (c - b);
                bool done = false;
                if (fb == 0) {
                    root = b;
                    xl = b;
                    xh = b;
                    done = true;
                } else if (fabs(m) <= tol) {
                    root = b;
                    if (b < c) {
                        xl = b;
                        xh = c;
                    } else {
                        xl = c;
                        xh = b;
                    }
                    done = true;
                }
                if (done) {
                    evaluate = false;
                    // convergence test happens only from the second iteration on (root_finder.F90:1029)
                    if (iteration > 1) {
                        const double al = fabs(xl), au = fabs(xh);
                        const double minAbs = ((xl > 0.0 && xh > 0.0) || (xl < 0.0 && xh < 0.0)) ? fmin(al, au) : 0.0;
                        if (fabs(xh - xl) < o.tolAbs + o.tolRel * minAbs) {
                            result = root;
                            busy = false;
                        }
                    }
                    if (busy && iteration > 1000) {  // iterationMaximum, root_finder.F90:1028
                        result = root;
                        busy = false;
                    }
                } else {
                    if (fabs(e) < tol || fabs(fa) <= fabs(fb)) {
                        d = m;
                        e = m;
                    } else {
                        double p, q, r;
                        const double s = fb / fa;
                        if (acEqual) {
                            p = 2 * m * s;
                            q = 1 - s;
                        } else {
                            q = fa / fc;
                            r = fb / fc;
                            p = s * (2 * m * q * (q - r) - (b - a) * (r - 1));
                            q = (q - 1) * (r - 1) * (s - 1);
                        }
                        if (p > 0)
                            q = -q;
                        else
                            p = -p;
                        if (2 * p < fmin(3 * m * q - fabs(tol * q), fabs(e * q))) {
                            e = d;
                            d = p / q;
                        } else {
                            d = m;
                            e = m;
                        }
                    }
                    a = b;
                    fa = fb;
                    if (fabs(d) > tol)
                        b += d;
                    else
                        b += (m > 0 ? +tol : -tol);
                    x = b;
                }
            }
        }
        if (busy && evaluate) {
            const double fx = f(x);  // ---- the single call site (straight-line integrands only)

            if (state == ST_FLO) {
                fLow = fx;
                state = ST_FHI;
            } else if (state == ST_FHI) {
                fHigh = fx;
                state = ST_BRACKET;
            } else if (state == ST_EXP_UP) {
                fHigh = fx;
                state = ST_EXP_DOWN;
            } else if (state == ST_EXP_DOWN) {
                fLow = fx;
                state = ST_BRACKET;
            } else {
                fb = fx;
                root = b;
                double cc = c;
                if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) cc = a;
                if (b < cc) {
                    xl = b;
                    xh = cc;
                } else {
                    xl = cc;
                    xh = b;
                }
                if (iteration > 1) {
                    const double al = fabs(xl), au = fabs(xh);
                    const double minAbs = ((xl > 0.0 && xh > 0.0) || (xl < 0.0 && xh < 0.0)) ? fmin(al, au) : 0.0;
                    if (fabs(xh - xl) < o.tolAbs + o.tolRel * minAbs) {
                        result = root;
                        busy = false;
                    }
                }
                if (busy && iteration > 1000) {
                    result = root;
                    busy = false;
                }
            }
        }
    }
    return result;
}

// ---------------------------------------------------------------- Gauss-Kronrod 15 / QAG
// node j of the 15-point rule in the evaluation order of qk.c: centre, then the Gauss nodes
// (odd xgk indices), then the remaining Kronrod nodes (even indices), each as a -/+ pair.
__constant__ double c_xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                                0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                                0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
__constant__ double c_wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                               0.381830050505118944950369775488975, 0.417959183673469387755102040816327};
__constant__ double c_wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                                0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                                0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
// xgk index visited by pair k = 0..6 (qk.c loops: jtw = 1,3,5 then jtwm1 = 0,2,4,6)
__constant__ int c_qk_order[7] = {1, 3, 5, 0, 2, 4, 6};
// position p of xgk index k in that visiting order (result_asc is summed in xgk order)
__constant__ int c_qk_pos[7] = {3, 0, 4, 1, 5, 2, 6};

GLC_DEVICE_INLINE double rescale_error(double err, double resultAbs, double resultAsc) {
    err = fabs(err);
    if (resultAsc != 0 && err != 0) {
        const double scale = dm_pow((200 * err / resultAsc), 1.5);
        err = (scale < 1) ? resultAsc * scale : resultAsc;
    }
    if (resultAbs > DBL_MIN / (50 * DBL_EPSILON)) {
        const double minErr = 50 * DBL_EPSILON * resultAbs;
        if (minErr > err) err = minErr;
    }
    return err;
}

constexpr int kQagLimitDevice = 24;  // intervals kept per thread; the reference allows 1000 and aborts beyond

// gsl_integration_qag(key = GAUSS15). status: 0 ok, 11 interval budget exhausted, 18/21 round-off/singular.
// Warp-synchronous (see root_find): one pass of the loop = one 15-point rule on one (sub)interval.
template <class F>
GLC_DEVICE_INLINE double qag15(F &&f, bool on, double a, double b, double epsabs, double epsrel, int &status) {
    double alist[kQagLimitDevice], blist[kQagLimitDevice], rlist[kQagLimitDevice], elist[kQagLimitDevice];
    int size = 0, iteration = 0, errorType = 0, roundoff1 = 0, roundoff2 = 0;
    double area = 0, errsum = 0, tolerance = 0;
    status = 0;
    // work list: first the whole interval, then (a1,b1),(a2,b2) of each bisection
    double ia = a, ib = b;
    int phase = 0;  // 0: initial interval, 1: first half, 2: second half
    int iMax = 0;
    double a1 = 0, b1 = 0, a2 = 0, b2 = 0, rI = 0, eI = 0;
    double area1 = 0, error1 = 0, resasc1 = 0;
    bool busy = on, summed = false;
    double answer = 0.0;
    while (GLC_ANY(busy)) {
        if (busy) {
            // ---- qk15 on (ia, ib): one call site of f
            double result, abserr, resabs, resasc;
            {
                double fv[15];
                const double center = 0.5 * (ia + ib);
                const double halfLength = 0.5 * (ib - ia);
                const double absHalfLength = fabs(halfLength);
#pragma unroll 1
                for (int j = 0; j < 15; j++) {
                    double x;
                    if (j == 0)
                        x = center;
                    else {
                        const int k = c_qk_order[(j - 1) >> 1];
                        const double absc = halfLength * c_xgk[k];
                        x = ((j - 1) & 1) ? center + absc : center - absc;
                    }
                    fv[j] = f(x);
                }
                const double fCenter = fv[0];
                double resultGauss = fCenter * c_wg[3];
                double resultKronrod = fCenter * c_wgk[7];
                double resultAbs = fabs(resultKronrod);
#pragma unroll
                for (int p = 0; p < 7; p++) {
                    const int k = c_qk_order[p];
                    const double f1 = fv[1 + 2 * p], f2 = fv[2 + 2 * p];
                    if (p < 3) resultGauss += c_wg[p] * (f1 + f2);
                    resultKronrod += c_wgk[k] * (f1 + f2);
                    resultAbs += c_wgk[k] * (fabs(f1) + fabs(f2));
                }
                const double mean = resultKronrod * 0.5;
                double resultAsc = c_wgk[7] * fabs(fCenter - mean);
#pragma unroll
                for (int k = 0; k < 7; k++) {
                    const int p = c_qk_pos[k];
                    resultAsc += c_wgk[k] * (fabs(fv[1 + 2 * p] - mean) + fabs(fv[2 + 2 * p] - mean));
                }
                const double err = (resultKronrod - resultGauss) * halfLength;
                resultKronrod *= halfLength;
                resultAbs *= absHalfLength;
                resultAsc *= absHalfLength;
                result = resultKronrod;
                resabs = resultAbs;
                resasc = resultAsc;
                abserr = rescale_error(err, resultAbs, resultAsc);
            }
            bool bisect = true;
            if (phase == 0) {
                alist[0] = a;
                blist[0] = b;
                rlist[0] = result;
                elist[0] = abserr;
                size = 1;
                tolerance = fmax(epsabs, epsrel * fabs(result));
                const double roundOff = 50 * DBL_EPSILON * resabs;
                if (abserr <= roundOff && abserr > tolerance) {
                    status = 18;
                    answer = result;
                    busy = false;
                } else if ((abserr <= tolerance && abserr != resasc) || abserr == 0.0) {
                    answer = result;
                    busy = false;
                }
                area = result;
                errsum = abserr;
                iteration = 1;
            } else if (phase == 1) {
                area1 = result;
                error1 = abserr;
                resasc1 = resasc;
                ia = a2;
                ib = b2;
                phase = 2;
                bisect = false;
            } else {
                const double area2 = result, error2 = abserr, resasc2 = resasc;
                const double area12 = area1 + area2, error12 = error1 + error2;
                errsum += (error12 - eI);
                area += area12 - rI;
                if (resasc1 != error1 && resasc2 != error2) {
                    const double delta = rI - area12;
                    if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * eI) roundoff1++;
                    if (iteration >= 10 && error12 > eI) roundoff2++;
                }
                tolerance = fmax(epsabs, epsrel * fabs(area));
                if (errsum > tolerance) {
                    if (roundoff1 >= 6 || roundoff2 >= 20) errorType = 2;
                    const double tmp = (1 + 100 * DBL_EPSILON) * (fabs(a2) + 1000 * DBL_MIN);
                    if (fabs(a1) <= tmp && fabs(b2) <= tmp) errorType = 3;
                }
                if (error2 > error1) {
                    alist[iMax] = a2;
                    rlist[iMax] = area2;
                    elist[iMax] = error2;
                    alist[size] = a1;
                    blist[size] = b1;
                    rlist[size] = area1;
                    elist[size] = error1;
                } else {
                    blist[iMax] = b1;
                    rlist[iMax] = area1;
                    elist[iMax] = error1;
                    alist[size] = a2;
                    blist[size] = b2;
                    rlist[size] = area2;
                    elist[size] = error2;
                }
                size++;
                iteration++;
                if (!(iteration < 1000 && !errorType && errsum > tolerance)) {
                    busy = false;
                    summed = true;
                } else if (size >= kQagLimitDevice) {
                    status = 11;
                    busy = false;
                    summed = true;
                }
            }
            if (busy && bisect) {
                // ---- bisect the interval with the largest error
                iMax = 0;
                for (int i = 1; i < size; i++)
                    if (elist[i] > elist[iMax]) iMax = i;
                rI = rlist[iMax];
                eI = elist[iMax];
                a1 = alist[iMax];
                b1 = 0.5 * (alist[iMax] + blist[iMax]);
                a2 = b1;
                b2 = blist[iMax];
                ia = a1;
                ib = b1;
                phase = 1;
            }
        }
    }
    if (summed) {
        double sum = 0;
        for (int i = 0; i < size; i++) sum += rlist[i];
        if (status == 0 && errsum > tolerance) status = (errorType == 2) ? 18 : ((errorType == 3) ? 21 : 11);
        answer = sum;
    }
    return answer;
}

// value of a table1DLinearLinear with n points on [xmin,xmax] populated by g (evaluated on the fly)
template <class G>
GLC_DEVICE_INLINE double linear_table_eval(G &&g, double xmin, double xmax, int n, double x,
                                                    bool extrapFix) {
    const double dx = (xmax - xmin) / (double)(n - 1);
    const double inverseDx = 1.0 / ((xmin + dx) - xmin);
    double xe = x;
    if (extrapFix) xe = fmin(fmax(xe, xmin), xmax);
    int i;
    if (xe < xmin)
        i = 1;
    else if (xe >= xmax)
        i = n - 1;
    else
        i = max(min((int)((xe - xmin) * inverseDx) + 1, n - 1), 1);
    const double xi = xmin + dx * (double)(i - 1);
    const double xi1 = (i == n - 1) ? xmax : xmin + dx * (double)i;
    const double h = (xe - xi) * inverseDx;
    return g(xi) * (1.0 - h) + g(xi1) * h;
}

// fastExponentiator (math/exponentiation.F90:57-104): linear interpolation in a table of x^exponent with `density`
// points per unit x on [rangeMin, rangeMax], exact pow outside.  `table` holds the n lattice values (built once on the
// host by build_pow_table with the same dm_pow).
GLC_DEVICE_INLINE double fast_exponentiate(const double *__restrict__ table, int n, double rangeMin, double rangeMax,
                                           double exponent, double x) {
    if (x < rangeMin || x > rangeMax) return dm_pow(x, exponent);
    const double dx = (rangeMax - rangeMin) / (double)(n - 1);
    const double inverseDx = 1.0 / ((rangeMin + dx) - rangeMin);
    int i;
    if (x >= rangeMax)
        i = n - 1;
    else
        i = max(min((int)((x - rangeMin) * inverseDx) + 1, n - 1), 1);
    const double xi = rangeMin + dx * (double)(i - 1);
    const double h = (x - xi) * inverseDx;
    return GLC_LDG(table + i - 1) * (1.0 - h) + GLC_LDG(table + i) * h;
}

}  // namespace glc
