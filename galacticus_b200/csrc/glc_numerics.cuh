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

// Re-entrant form of rootFinder%find (Brent branch): the caller owns the loop.
//   brent_begin(...)            initialise;   (the RootOptions are passed to every call: compile-time constants)
//   brent_advance(B) -> bool    run the state machine to the next abscissa B.x at which the function is needed
//                               (false: nothing to evaluate -- finished, B.busy == 0, B.result/B.status are set);
//   brent_feed(B, f(B.x))       digest the value.
// status 0 = ok, 2 = could not bracket, 3 = bad bracket
struct BrentState {
    double xLow, xHigh, fLow, fHigh;
    double a, b, c, d, e, fa, fb, fc, xl, xh, root, x, result;
    int state, iteration, status;
    int lowerOk, upperOk, rangeChanged, first, busy;
};
enum : int { ST_FLO, ST_FHI, ST_BRACKET, ST_EXP_UP, ST_EXP_DOWN, ST_BRENT };

GLC_DEVICE_INLINE void brent_begin(BrentState &B, bool on, double xLow, double xHigh, bool haveValues, double fLow,
                                   double fHigh) {
    B.xLow = xLow;
    B.xHigh = xHigh;
    B.fLow = fLow;
    B.fHigh = fHigh;
    B.a = B.b = B.c = B.d = B.e = B.fa = B.fb = B.fc = B.xl = B.xh = B.root = B.x = B.result = 0.0;
    B.state = haveValues ? ST_BRACKET : ST_FLO;
    B.iteration = 0;
    B.status = 0;
    B.lowerOk = B.upperOk = B.rangeChanged = 0;
    B.first = 1;
    B.busy = on ? 1 : 0;
}

// GLC_NOINLINE_BRENT (experiment, profiles/r02ad): one copy of the Brent state machine per kernel instead of one per root problem
#ifdef GLC_NOINLINE_BRENT
#define GLC_DEVICE_BRENT GLC_DEVICE_NOINLINE
#else
#define GLC_DEVICE_BRENT GLC_DEVICE_INLINE
#endif
GLC_DEVICE_BRENT bool brent_advance(BrentState &B, const RootOptions &o) {
    double x = 0.0;
    bool evaluate = false;
    while (B.busy && !evaluate) {
            evaluate = true;
            if (B.state == ST_FLO) {
                x = B.xLow;
            } else if (B.state == ST_FHI) {
                x = B.xHigh;
            } else if (B.state == ST_BRACKET) {
                evaluate = false;
                if (B.first) {
                    if (B.xHigh == B.xLow) B.fHigh = B.fLow;
                    B.first = false;
                }
                if (fsign1(B.fLow) * fsign1(B.fHigh) > 0.0 && B.fLow != 0.0 && B.fHigh != 0.0) {
                    B.lowerOk = o.signExpectDownward == SIGN_NEGATIVE   ? (B.fLow < 0.0)
                              : o.signExpectDownward == SIGN_POSITIVE ? (B.fLow > 0.0)
                                                                      : false;
                    B.upperOk = o.signExpectUpward == SIGN_NEGATIVE   ? (B.fHigh < 0.0)
                              : o.signExpectUpward == SIGN_POSITIVE ? (B.fHigh > 0.0)
                                                                    : false;
                    B.rangeChanged = false;
                    B.state = ST_EXP_UP;
                } else {
                    // brent_init (function values at the bracket ends are already known)
                    B.a = B.xLow;
                    B.fa = B.fLow;
                    B.b = B.xHigh;
                    B.fb = B.fHigh;
                    B.c = B.xHigh;
                    B.fc = B.fHigh;
                    B.d = B.xHigh - B.xLow;
                    B.e = B.xHigh - B.xLow;
                    if ((B.fLow < 0.0 && B.fHigh < 0.0) || (B.fLow > 0.0 && B.fHigh > 0.0)) {
                        B.status = 3;
                        B.result = 0.0;
                        B.busy = false;
                    }
                    B.state = ST_BRENT;
                }
            } else if (B.state == ST_EXP_UP) {
                bool move;
                if (o.expandType == EXPAND_ADDITIVE)
                    move = o.expandUpward > 0.0 && !B.upperOk;
                else if (o.expandType == EXPAND_MULTIPLICATIVE)
                    move = ((o.expandUpward > 1.0 && B.xHigh > 0.0) || (o.expandUpward < 1.0 && B.xHigh < 0.0)) && !B.upperOk;
                else
                    move = false;
                if (move) {
                    if (B.lowerOk) {
                        B.xLow = B.xHigh;
                        B.fLow = B.fHigh;
                    }
                    B.xHigh = (o.expandType == EXPAND_ADDITIVE) ? B.xHigh + o.expandUpward : B.xHigh * o.expandUpward;
                    x = B.xHigh;
                    B.rangeChanged = true;
                } else {
                    evaluate = false;
                    B.state = ST_EXP_DOWN;
                }
            } else if (B.state == ST_EXP_DOWN) {
                bool move;
                if (o.expandType == EXPAND_ADDITIVE)
                    move = o.expandDownward < 0.0 && !B.lowerOk;
                else if (o.expandType == EXPAND_MULTIPLICATIVE)
                    move = ((o.expandDownward < 1.0 && B.xLow > 0.0) || (o.expandDownward > 1.0 && B.xLow < 0.0)) && !B.lowerOk;
                else
                    move = false;
                if (move) {
                    if (B.upperOk) {
                        B.xHigh = B.xLow;
                        B.fHigh = B.fLow;
                    }
                    B.xLow = (o.expandType == EXPAND_ADDITIVE) ? B.xLow + o.expandDownward : B.xLow * o.expandDownward;
                    x = B.xLow;
                    B.rangeChanged = true;
                } else {
                    evaluate = false;
                    if (!B.rangeChanged) {
                        B.status = 2;
                        B.result = 0.0;
                        B.busy = false;
                    }
                    B.state = ST_BRACKET;
                }
            } else {  // ST_BRENT: brent_iterate up to the point where f(B.b) is needed
                double tol, m;
                bool acEqual = false;
                B.iteration++;
                if ((B.fb < 0 && B.fc < 0) || (B.fb > 0 && B.fc > 0)) {
                    acEqual = true;
                    B.c = B.a;
                    B.fc = B.fa;
                    B.d = B.b - B.a;
                    B.e = B.b - B.a;
                }
                if (fabs(B.fc) < fabs(B.fb)) {
                    acEqual = true;
                    B.a = B.b;
                    B.b = B.c;
                    B.c = B.a;
                    B.fa = B.fb;
                    B.fb = B.fc;
                    B.fc = B.fa;
                }
                tol = 0.5 * DBL_EPSILON * fabs(B.b);
                m = 0.5 * (B.c - B.b);
                bool done = false;
                if (B.fb == 0) {
                    B.root = B.b;
                    B.xl = B.b;
                    B.xh = B.b;
                    done = true;
                } else if (fabs(m) <= tol) {
                    B.root = B.b;
                    if (B.b < B.c) {
                        B.xl = B.b;
                        B.xh = B.c;
                    } else {
                        B.xl = B.c;
                        B.xh = B.b;
                    }
                    done = true;
                }
                if (done) {
                    evaluate = false;
                    // convergence test happens only from the second B.iteration on (root_finder.F90:1029)
                    if (B.iteration > 1) {
                        const double al = fabs(B.xl), au = fabs(B.xh);
                        const double minAbs = ((B.xl > 0.0 && B.xh > 0.0) || (B.xl < 0.0 && B.xh < 0.0)) ? fmin(al, au) : 0.0;
                        if (fabs(B.xh - B.xl) < o.tolAbs + o.tolRel * minAbs) {
                            B.result = B.root;
                            B.busy = false;
                        }
                    }
                    if (B.busy && B.iteration > 1000) {  // iterationMaximum, root_finder.F90:1028
                        B.result = B.root;
                        B.busy = false;
                    }
                } else {
                    if (fabs(B.e) < tol || fabs(B.fa) <= fabs(B.fb)) {
                        B.d = m;
                        B.e = m;
                    } else {
                        double p, q, r;
                        const double s = B.fb / B.fa;
                        if (acEqual) {
                            p = 2 * m * s;
                            q = 1 - s;
                        } else {
                            q = B.fa / B.fc;
                            r = B.fb / B.fc;
                            p = s * (2 * m * q * (q - r) - (B.b - B.a) * (r - 1));
                            q = (q - 1) * (r - 1) * (s - 1);
                        }
                        if (p > 0)
                            q = -q;
                        else
                            p = -p;
                        if (2 * p < fmin(3 * m * q - fabs(tol * q), fabs(B.e * q))) {
                            B.e = B.d;
                            B.d = p / q;
                        } else {
                            B.d = m;
                            B.e = m;
                        }
                    }
                    B.a = B.b;
                    B.fa = B.fb;
                    if (fabs(B.d) > tol)
                        B.b += B.d;
                    else
                        B.b += (m > 0 ? +tol : -tol);
                    x = B.b;
                }
            }
        }
    B.x = x;
    return B.busy && evaluate;
}

GLC_DEVICE_BRENT void brent_feed(BrentState &B, const RootOptions &o, double fx) {
            if (B.state == ST_FLO) {
                B.fLow = fx;
                B.state = ST_FHI;
            } else if (B.state == ST_FHI) {
                B.fHigh = fx;
                B.state = ST_BRACKET;
            } else if (B.state == ST_EXP_UP) {
                B.fHigh = fx;
                B.state = ST_EXP_DOWN;
            } else if (B.state == ST_EXP_DOWN) {
                B.fLow = fx;
                B.state = ST_BRACKET;
            } else {
                B.fb = fx;
                B.root = B.b;
                double cc = B.c;
                if ((B.fb < 0 && B.fc < 0) || (B.fb > 0 && B.fc > 0)) cc = B.a;
                if (B.b < cc) {
                    B.xl = B.b;
                    B.xh = cc;
                } else {
                    B.xl = cc;
                    B.xh = B.b;
                }
                if (B.iteration > 1) {
                    const double al = fabs(B.xl), au = fabs(B.xh);
                    const double minAbs = ((B.xl > 0.0 && B.xh > 0.0) || (B.xl < 0.0 && B.xh < 0.0)) ? fmin(al, au) : 0.0;
                    if (fabs(B.xh - B.xl) < o.tolAbs + o.tolRel * minAbs) {
                        B.result = B.root;
                        B.busy = false;
                    }
                }
                if (B.busy && B.iteration > 1000) {
                    B.result = B.root;
                    B.busy = false;
                }
            }
}

// Warp-synchronous calling discipline.  Every routine below that contains a data-dependent loop must be
// called by ALL lanes of the warp from uniform control flow; `on` tells whether this lane really wants the
// result.  The loops run `while (GLC_ANY(lane still busy))`: the vote is an explicit reconvergence point
// in every iteration, lanes that are finished idle through the body.
template <class F>
GLC_DEVICE_INLINE double root_find(F &&f, bool on, const RootOptions &o, double xLow, double xHigh,
                                            bool haveValues, double fLow, double fHigh, int &status) {
    BrentState B;
    brent_begin(B, on, xLow, xHigh, haveValues, fLow, fHigh);
    while (GLC_ANY(B.busy != 0)) {
        if (brent_advance(B, o)) brent_feed(B, o, f(B.x));  // ---- the single call site (straight-line integrands only)
    }
    status = B.status;
    return B.result;
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

#ifndef GLC_QAG_UNROLL
#define GLC_QAG_UNROLL 1
#endif
constexpr int kQagUnroll = GLC_QAG_UNROLL;
constexpr int kQagLimitDevice = 24;  // intervals kept per thread; the reference allows 1000 and aborts beyond

// Re-entrant form of gsl_integration_qag(key = GAUSS15): the caller owns the loop.
//   qag_begin(...)      initialise;   qag_pass(Q, f)  one 15-point rule on the pending (sub)interval + bookkeeping
//   (requires Q.busy);  qag_finish(Q) the integral once Q.busy == 0.
// status: 0 ok, 11 interval budget exhausted, 18/21 round-off/singular
struct QagState {
    double alist[kQagLimitDevice], blist[kQagLimitDevice], rlist[kQagLimitDevice], elist[kQagLimitDevice];
    double a, b, epsabs, epsrel;
    double area, errsum, tolerance, ia, ib, a1, b1, a2, b2, rI, eI, area1, error1, resasc1, answer;
    int size, iteration, errorType, roundoff1, roundoff2, phase, iMax, status, busy, summed;
};

GLC_DEVICE_INLINE void qag_begin(QagState &Q, bool on, double a, double b, double epsabs, double epsrel) {
    Q.a = a;
    Q.b = b;
    Q.epsabs = epsabs;
    Q.epsrel = epsrel;
    Q.area = Q.errsum = Q.tolerance = 0.0;
    Q.ia = a;
    Q.ib = b;
    Q.a1 = Q.b1 = Q.a2 = Q.b2 = Q.rI = Q.eI = Q.area1 = Q.error1 = Q.resasc1 = Q.answer = 0.0;
    Q.size = Q.iteration = Q.errorType = Q.roundoff1 = Q.roundoff2 = 0;
    Q.phase = 0;  // 0: initial interval, 1: first half, 2: second half
    Q.iMax = 0;
    Q.status = 0;
    Q.busy = on ? 1 : 0;
    Q.summed = 0;
}

// abscissa j of the 15-point rule in the order the values are taken: the centre, then the pairs centre -/+ halfLength x_k
GLC_DEVICE_INLINE double qag_abscissa(double center, double halfLength, int j) {
    if (j == 0) return center;
    const int k = c_qk_order[(j - 1) >> 1];
    const double absc = halfLength * c_xgk[k];
    return ((j - 1) & 1) ? center + absc : center - absc;
}

// second half of a pass: qk15's sums over the 15 values fv[] taken on (Q.ia, Q.ib), then qag's bookkeeping
GLC_DEVICE_INLINE void qag_digest(QagState &Q, const double (&fv)[15]) {
    double result, abserr, resabs, resasc;
    {
        const double halfLength = 0.5 * (Q.ib - Q.ia);
        const double absHalfLength = fabs(halfLength);
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
    if (Q.phase == 0) {
        Q.alist[0] = Q.a;
        Q.blist[0] = Q.b;
        Q.rlist[0] = result;
        Q.elist[0] = abserr;
        Q.size = 1;
        Q.tolerance = fmax(Q.epsabs, Q.epsrel * fabs(result));
        const double roundOff = 50 * DBL_EPSILON * resabs;
        if (abserr <= roundOff && abserr > Q.tolerance) {
            Q.status = 18;
            Q.answer = result;
            Q.busy = false;
        } else if ((abserr <= Q.tolerance && abserr != resasc) || abserr == 0.0) {
            Q.answer = result;
            Q.busy = false;
        }
        Q.area = result;
        Q.errsum = abserr;
        Q.iteration = 1;
    } else if (Q.phase == 1) {
        Q.area1 = result;
        Q.error1 = abserr;
        Q.resasc1 = resasc;
        Q.ia = Q.a2;
        Q.ib = Q.b2;
        Q.phase = 2;
        bisect = false;
    } else {
        const double area2 = result, error2 = abserr, resasc2 = resasc;
        const double area12 = Q.area1 + area2, error12 = Q.error1 + error2;
        Q.errsum += (error12 - Q.eI);
        Q.area += area12 - Q.rI;
        if (Q.resasc1 != Q.error1 && resasc2 != error2) {
            const double delta = Q.rI - area12;
            if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * Q.eI) Q.roundoff1++;
            if (Q.iteration >= 10 && error12 > Q.eI) Q.roundoff2++;
        }
        Q.tolerance = fmax(Q.epsabs, Q.epsrel * fabs(Q.area));
        if (Q.errsum > Q.tolerance) {
            if (Q.roundoff1 >= 6 || Q.roundoff2 >= 20) Q.errorType = 2;
            const double tmp = (1 + 100 * DBL_EPSILON) * (fabs(Q.a2) + 1000 * DBL_MIN);
            if (fabs(Q.a1) <= tmp && fabs(Q.b2) <= tmp) Q.errorType = 3;
        }
        if (error2 > Q.error1) {
            Q.alist[Q.iMax] = Q.a2;
            Q.rlist[Q.iMax] = area2;
            Q.elist[Q.iMax] = error2;
            Q.alist[Q.size] = Q.a1;
            Q.blist[Q.size] = Q.b1;
            Q.rlist[Q.size] = Q.area1;
            Q.elist[Q.size] = Q.error1;
        } else {
            Q.blist[Q.iMax] = Q.b1;
            Q.rlist[Q.iMax] = Q.area1;
            Q.elist[Q.iMax] = Q.error1;
            Q.alist[Q.size] = Q.a2;
            Q.blist[Q.size] = Q.b2;
            Q.rlist[Q.size] = area2;
            Q.elist[Q.size] = error2;
        }
        Q.size++;
        Q.iteration++;
        if (!(Q.iteration < 1000 && !Q.errorType && Q.errsum > Q.tolerance)) {
            Q.busy = false;
            Q.summed = true;
        } else if (Q.size >= kQagLimitDevice) {
            Q.status = 11;
            Q.busy = false;
            Q.summed = true;
        }
    }
    if (Q.busy && bisect) {
        // ---- bisect the interval with the largest error
        Q.iMax = 0;
        for (int i = 1; i < Q.size; i++)
            if (Q.elist[i] > Q.elist[Q.iMax]) Q.iMax = i;
        Q.rI = Q.rlist[Q.iMax];
        Q.eI = Q.elist[Q.iMax];
        Q.a1 = Q.alist[Q.iMax];
        Q.b1 = 0.5 * (Q.alist[Q.iMax] + Q.blist[Q.iMax]);
        Q.a2 = Q.b1;
        Q.b2 = Q.blist[Q.iMax];
        Q.ia = Q.a1;
        Q.ib = Q.b1;
        Q.phase = 1;
    }
}

// one pass by one lane: the 15 values one after the other (the micro-task machine's QAG unit, the CPU build)
template <class F>
GLC_DEVICE_INLINE void qag_pass(QagState &Q, F &&f) {
    double fv[15];
    const double center = 0.5 * (Q.ia + Q.ib);
    const double halfLength = 0.5 * (Q.ib - Q.ia);
    // the 15 abscissae are independent: unrolling by GLC_QAG_UNROLL interleaves their dependency chains (a lone lane is
    // bound by the latency of dependent FP64 instructions, not by issue slots)
#if defined(__CUDACC__)
#pragma unroll kQagUnroll
#endif
    for (int j = 0; j < 15; j++) fv[j] = f(qag_abscissa(center, halfLength, j));
    qag_digest(Q, fv);
}

GLC_DEVICE_INLINE double qag_finish(QagState &Q) {
    if (Q.summed) {
        double sum = 0;
        for (int i = 0; i < Q.size; i++) sum += Q.rlist[i];
        if (Q.status == 0 && Q.errsum > Q.tolerance) Q.status = (Q.errorType == 2) ? 18 : ((Q.errorType == 3) ? 21 : 11);
        Q.answer = sum;
    }
    return Q.answer;
}

// Warp-synchronous (see root_find): one pass of the loop = one 15-point rule on one (sub)interval.
template <class F>
GLC_DEVICE_INLINE double qag15(F &&f, bool on, double a, double b, double epsabs, double epsrel, int &status) {
    QagState Q;
    qag_begin(Q, on, a, b, epsabs, epsrel);
    while (GLC_ANY(Q.busy != 0)) {
        if (Q.busy) qag_pass(Q, f);
    }
    const double r = qag_finish(Q);
    status = Q.status;
    return r;
}

#if defined(__CUDACC__)
// The same integral with the 15 values of a pass taken by 15 LANES at once.  In the warp-synchronous kernels few lanes of a
// warp are inside an integral at the same time (4 of 32 in a dense drain pass, 1 in a lone-lane pass, profiles/r02ac), and
// the values are independent: two busy lanes ("owners") are served per round, lanes 0-14 take the first owner's abscissae
// and lanes 16-30 the second's, with the owner's problem (params: plain doubles; centre, half length) fetched by shuffles
// and the values shuffled back in rule order.  Every value is computed by the same instructions from the same inputs and
// summed in the same order by its owner (qag_digest), so the result is bit-identical to qag15; with more than kQagCoopMax
// busy lanes the rounds would outnumber the 15 sequential evaluations and the warp takes the plain pass.
constexpr int kQagCoopMax = 16;
template <class T>
GLC_DEVICE_INLINE T shfl_words(const T &v, int src) {
    static_assert(sizeof(T) % sizeof(double) == 0, "plain doubles only");
    T out;
    const double *in = reinterpret_cast<const double *>(&v);
    double *o = reinterpret_cast<double *>(&out);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / sizeof(double)); i++) o[i] = __shfl_sync(0xffffffffu, in[i], src);
    return out;
}
template <class P, class F>
GLC_DEVICE_INLINE double qag15_coop(const P &params, F &&f, bool on, double a, double b, double epsabs, double epsrel,
                                    int &status) {
    QagState Q;
    qag_begin(Q, on, a, b, epsabs, epsrel);
    const int lane = threadIdx.x & 31, group = lane >> 4, j = lane & 15;
    for (;;) {
        unsigned int busy = __ballot_sync(0xffffffffu, Q.busy != 0);
        if (!busy) break;
        if (__popc(busy) > kQagCoopMax) {
            if (Q.busy) qag_pass(Q, [&](double x) { return f(params, x); });
            continue;
        }
        double fv[15];
        const double center = 0.5 * (Q.ia + Q.ib);
        const double halfLength = 0.5 * (Q.ib - Q.ia);
        while (busy) {  // (warp-uniform)
            const int owner0 = __ffs(busy) - 1;
            busy &= busy - 1;
            const int owner1 = busy ? __ffs(busy) - 1 : -1;
            busy &= busy - 1;  // (0 stays 0)
            const int owner = group == 0 ? owner0 : owner1;
            const int src = owner < 0 ? lane : owner;
            const P p = shfl_words(params, src);
            const double c = __shfl_sync(0xffffffffu, center, src), h = __shfl_sync(0xffffffffu, halfLength, src);
            double v = 0.0;
            if (owner >= 0 && j < 15) v = f(p, qag_abscissa(c, h, j));
            const int base = lane == owner1 ? 16 : 0;
            const bool mine = lane == owner0 || lane == owner1;
#pragma unroll
            for (int k = 0; k < 15; k++) {
                const double t = __shfl_sync(0xffffffffu, v, base + k);
                if (mine) fv[k] = t;
            }
        }
        if (Q.busy) qag_digest(Q, fv);
    }
    const double r = qag_finish(Q);
    status = Q.status;
    return r;
}
#endif

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
template <bool kStaged = false>
GLC_DEVICE_INLINE double fast_exponentiate(const double *__restrict__ table, int n, double dx, double inverseDx,
                                           double rangeMin, double rangeMax, double exponent, double x) {
    // dx = (rangeMax - rangeMin) / (n - 1) and inverseDx = 1 / ((rangeMin + dx) - rangeMin) are table constants
    // (pow_table_spacing, glc_tables_host.h): the same two IEEE divisions, done once instead of per call
    if (x < rangeMin || x > rangeMax) return dm_pow(x, exponent);
    int i;
    if (x >= rangeMax)
        i = n - 1;
    else
        i = max(min((int)((x - rangeMin) * inverseDx) + 1, n - 1), 1);
    const double xi = rangeMin + dx * (double)(i - 1);
    const double h = (x - xi) * inverseDx;
    if (kStaged) return table[i - 1] * (1.0 - h) + table[i] * h;  // a staged copy (shared memory): not the read-only global path
    return GLC_LDG(table + i - 1) * (1.0 - h) + GLC_LDG(table + i) * h;
}

}  // namespace glc
