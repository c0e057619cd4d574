// ltk_fitpack_core.cuh -- FITPACK's periodic interpolating cubic spline, one candidate per thread.
//
//   reference: Path.__init__ -> scipy.interpolate.splprep(controls, u=dists, k=3, s=0, per=1) (path.py:25)
//              Path.curvature -> splev(u, tck, der=1|2)                                        (path.py:51-54)
//
// The default K1a solves the classical cyclic tridiagonal system for the same spline; its curvature agrees
// with FITPACK's to ~1e-13, which TBR18's friction-circle cancellation (vehicle.py:29-35) amplifies past the
// 1e-9 lap-time tolerance on a tail of candidates.  This file is the `LTK_SPLINE_FITPACK` mode: Dierckx's own
// algorithm (clocur -> fpclos, s = 0, odd degree: a row-by-row Givens QR of the banded periodic collocation
// matrix, fpbacp back substitution; splder: de Boor derivative coefficients + fpbspl evaluation) with every
// floating-point operation in FITPACK's order, so that coefficients, derivatives and curvature come out with
// the bits SciPy produces.  oracle/fitpack_port.c is the plain-C statement of the same routines (pinned
// against SciPy); tests/test_host.py compiles THIS header for the host and compares the two bit for bit.
//
// What is restructured (same operations, same operands, different schedule):
//   * the collocation row of data site i is non-zero in columns i, i+1, i+2 only, and rows arrive in order,
//     so the band rows being rotated are a three-row window kept in registers: row j is final as soon as
//     data row j has been rotated into it;
//   * the two wrapping data rows (the last two sites, whose B-splines reach across the period) are rotated
//     through row j right after it became final instead of in two later passes over the whole factor --
//     they only ever meet final rows, in the same order, so the operands are identical;
//   * fpbspl at a data site is evaluated with the terms that FITPACK multiplies by an exact zero left out.
// Everything is a function of one candidate: arrays are addressed as p[i * stride] so that the same code
// runs on lane-minor shared memory, candidate-minor global memory (stride Bp) or a host array (stride 1).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LTK_HD __host__ __device__ __forceinline__
#else
#define LTK_HD inline
#ifndef __align__
#define __align__(n) alignas(n)
#endif
#endif

namespace ltk {
namespace fit {

LTK_HD double fdiv(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return ddiv<false>(a, b);
#else
    return a / b;
#endif
}
LTK_HD void prefetch(const double* p)  // into L1, several rows ahead of a dependent loop (no-op on the host)
{
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
LTK_HD double fsqrt(double x)
{
#if defined(__CUDA_ARCH__)
    return dsqrt<false>(x);
#else
    return sqrt(x);
#endif
}

// fpgivs: rotation that annihilates piv against the diagonal element ww (ww > 0 or ww == 0 on first touch)
LTK_HD void givens(double piv, double& ww, double& c, double& s)
{
    const double store = fabs(piv);
    double dd;
    if (store >= ww) {
        const double r = fdiv(ww, piv);
        dd = store * fsqrt(1.0 + r * r);
    } else {
        const double r = fdiv(piv, ww);
        dd = ww * fsqrt(1.0 + r * r);
    }
    c = fdiv(ww, dd);
    s = fdiv(piv, dd);
    ww = dd;
}

// Two independent divisions / square roots written stage by stage: a warp issues in order, so two dependency
// chains only overlap as far as the instruction stream alternates between them (same trick as dsqrt_pair in
// ltk_sweep_fused.cuh).  Same operations, same results as two scalar calls.
LTK_HD void fdiv2(double a0, double b0, double a1, double b1, double& q0, double& q1)
{
#if defined(__CUDA_ARCH__)
    double y0 = rcp_seed(b0), y1 = rcp_seed(b1);
    double e0 = fma(y0, -b0, 1.0), e1 = fma(y1, -b1, 1.0);
    e0 = fma(e0, e0, e0); e1 = fma(e1, e1, e1);
    y0 = fma(y0, e0, y0); y1 = fma(y1, e1, y1);
    e0 = fma(y0, -b0, 1.0); e1 = fma(y1, -b1, 1.0);
    y0 = fma(y0, e0, y0); y1 = fma(y1, e1, y1);
    const double p0 = y0 * a0, p1 = y1 * a1;
    const double r0 = fma(p0, -b0, a0), r1 = fma(p1, -b1, a1);
    q0 = fma(y0, r0, p0); q1 = fma(y1, r1, p1);
#else
    q0 = a0 / b0; q1 = a1 / b1;
#endif
}
// c = w / d and s = p / d for two rotations: one refined reciprocal per divisor, four quotients
LTK_HD void fdiv2x2(double w0, double p0, double d0, double w1, double p1, double d1, double& c0, double& s0,
                    double& c1, double& s1)
{
#if defined(__CUDA_ARCH__)
    double y0 = rcp_seed(d0), y1 = rcp_seed(d1);
    double e0 = fma(y0, -d0, 1.0), e1 = fma(y1, -d1, 1.0);
    e0 = fma(e0, e0, e0); e1 = fma(e1, e1, e1);
    y0 = fma(y0, e0, y0); y1 = fma(y1, e1, y1);
    e0 = fma(y0, -d0, 1.0); e1 = fma(y1, -d1, 1.0);
    y0 = fma(y0, e0, y0); y1 = fma(y1, e1, y1);
    const double qc0 = y0 * w0, qc1 = y1 * w1, qs0 = y0 * p0, qs1 = y1 * p1;
    const double rc0 = fma(qc0, -d0, w0), rc1 = fma(qc1, -d1, w1), rs0 = fma(qs0, -d0, p0), rs1 = fma(qs1, -d1, p1);
    c0 = fma(y0, rc0, qc0); c1 = fma(y1, rc1, qc1); s0 = fma(y0, rs0, qs0); s1 = fma(y1, rs1, qs1);
#else
    c0 = w0 / d0; s0 = p0 / d0; c1 = w1 / d1; s1 = p1 / d1;
#endif
}
LTK_HD void fsqrt2(double x0, double x1, double& r0, double& r1)
{
#if defined(__CUDA_ARCH__)
    dsqrt_pair(x0, x1, r0, r1);
#else
    r0 = sqrt(x0); r1 = sqrt(x1);
#endif
}

// Two independent fpgivs set-ups (both pivots non-zero), branch-free and interleaved.
LTK_HD void givens2(double p0, double& w0, double& c0, double& s0, double p1, double& w1, double& c1, double& s1)
{
    const double a0 = fabs(p0), a1 = fabs(p1);
    const bool g0 = a0 >= w0, g1 = a1 >= w1;
    double r0, r1;
    fdiv2(g0 ? w0 : p0, g0 ? p0 : w0, g1 ? w1 : p1, g1 ? p1 : w1, r0, r1);
    double t0, t1;
    fsqrt2(1.0 + r0 * r0, 1.0 + r1 * r1, t0, t1);
    const double d0 = (g0 ? a0 : w0) * t0, d1 = (g1 ? a1 : w1) * t1;
    fdiv2x2(w0, p0, d0, w1, p1, d1, c0, s0, c1, s1);
    w0 = d0; w1 = d1;
}

// fprota
LTK_HD void rota(double c, double s, double& a, double& b)
{
    const double s1 = a, s2 = b;
    b = c * s2 + s * s1;
    a = c * s1 - s * s2;
}

// fpbspl(k = 3) at x = t(l): the three non-zero cubic B-splines at a data site, from t(l-2) .. t(l+2).
LTK_HD void bspl_at_knot(double tm2, double tm1, double t0, double tp1, double tp2, double& h1, double& h2,
                         double& h3)
{
    const double a = tp1 - t0;  // t(l+1) - x
    double f = fdiv(1.0, a);
    double g1 = f * a;                                   // degree 1: (g1, 0)
    f = fdiv(g1, tp1 - tm1);
    g1 = f * a;                                          // degree 2: (g1, g2, 0)
    double g2 = f * (t0 - tm1);
    double fa, fb;
    fdiv2(g1, tp1 - tm2, g2, tp2 - tm1, fa, fb);
    h1 = fa * a;                                         // degree 3
    h2 = fa * (t0 - tm2);
    h2 = h2 + fb * (tp2 - t0);
    h3 = fb * (t0 - tm1);
}

// The same in two halves, so that the three dependent divisions of the NEXT data row can be issued inside the
// basic blocks of the current row's two rotation stages (a warp issues in order: only instructions of one
// block interleave).  bspl_begin: degrees 1 and 2; bspl_end: degree 3.
struct BsplMid {
    double a, g1, g2;
};
LTK_HD BsplMid bspl_begin(double tm1, double t0, double tp1)
{
    BsplMid m;
    m.a = tp1 - t0;
    double f = fdiv(1.0, m.a);
    m.g1 = f * m.a;
    f = fdiv(m.g1, tp1 - tm1);
    m.g1 = f * m.a;
    m.g2 = f * (t0 - tm1);
    return m;
}
LTK_HD void bspl_end(const BsplMid& m, double tm2, double tm1, double t0, double tp1, double tp2, double& h1, double& h2,
                     double& h3)
{
    double fa, fb;
    fdiv2(m.g1, tp1 - tm2, m.g2, tp2 - tm1, fa, fb);
    h1 = fa * m.a;
    h2 = fa * (t0 - tm2);
    h2 = h2 + fb * (tp2 - t0);
    h3 = fb * (t0 - tm1);
}

struct Row {  // one row of the triangular factor: band part a1(j, 1..3), periodic part a2(j, 1..2), z(j) per dim
    double a, b, c, e1, e2, zx, zy;
};

struct WrapRow {  // a wrapping data row being rotated through the factor
    double h1[3], h2[2], x, y;
};

// the part of "rotation with the rows 1,2,...n10 of matrix a" (fpclos) that follows fpgivs, row j (1-based)
LTK_HD void wrap_apply(WrapRow& S, Row& R, int j, int n10, double c, double s)
{
    rota(c, s, S.x, R.zx);
    rota(c, s, S.y, R.zy);
    rota(c, s, S.h2[0], R.e1);
    rota(c, s, S.h2[1], R.e2);
    if (j == n10) return;
    rota(c, s, S.h1[1], R.b);
    S.h1[0] = S.h1[1];
    if (n10 - j >= 2) {
        rota(c, s, S.h1[2], R.c);
        S.h1[1] = S.h1[2];
        S.h1[2] = 0.0;
    } else {
        S.h1[1] = 0.0;
    }
}
LTK_HD void wrap_rotate(WrapRow& S, Row& R, int j, int n10)
{
    const double piv = S.h1[0];
    if (piv == 0.0) {
        S.h1[0] = S.h1[1]; S.h1[1] = S.h1[2]; S.h1[2] = 0.0;
        return;
    }
    double c, s;
    givens(piv, R.a, c, s);
    wrap_apply(S, R, j, n10, c, s);
}

// "rotation with the rows n10+1,...n7": R0 = row n10+1 (a2 entries in a, b), R1 = row n10+2 (a2 entry in a)
LTK_HD void wrap_tail(WrapRow& S, Row& R0, Row& R1)
{
    double c, s;
    double piv = S.h2[0];
    if (piv != 0.0) {
        givens(piv, R0.a, c, s);
        rota(c, s, S.x, R0.zx);
        rota(c, s, S.y, R0.zy);
        rota(c, s, S.h2[1], R0.b);
    }
    piv = S.h2[1];
    if (piv != 0.0) {
        givens(piv, R1.a, c, s);
        rota(c, s, S.x, R1.zx);
        rota(c, s, S.y, R1.zy);
    }
}

// Addressing of one candidate's arrays.
struct Io {
    double* t; long st;                            // knots [N + 7]; entries 3 .. N+3 hold the chord-length knots on entry
    double* rows; long sr;                         // scratch: 7 N entries
    double* cx; double* cy; long sc;               // B-spline coefficients [N + 3] each
    double* w1x; double* w1y; double* w2x; double* w2y; long sw;  // derivative coefficients [N+2], [N+2], [N+1], [N+1]
};

#define LTK_T(l) io.t[(long)((l) - 1) * io.st]          /* 1-based knot index as in FITPACK */
#define LTK_R(j, e) io.rows[(long)(((j) - 1) * 7 + (e)) * io.sr]
#define LTK_STORE_ROW(j, W)                                                                                   \
    do {                                                                                                      \
        LTK_R(j, 0) = (W).a; LTK_R(j, 1) = (W).b; LTK_R(j, 2) = (W).c; LTK_R(j, 3) = (W).e1;                   \
        LTK_R(j, 4) = (W).e2; LTK_R(j, 5) = (W).zx; LTK_R(j, 6) = (W).zy;                                      \
    } while (0)

// The whole solve for one candidate with N >= 5 unique control points.  Control point j (0-based, unique points
// only) comes in two steps so that its loads can be issued a whole data row before the values are needed:
// pts.fetch(j, raw) only loads, pts.finish(raw, x, y) does the arithmetic.
//
// Schedule of the forward pass.  Per data row `it` FITPACK performs three band rotations (into rows it, it+1,
// it+2) and, in our fused order, the two wrapping rows A and B are rotated through row `it` once it is final.
// The third band rotation always meets a fresh row (diagonal 0): fpgivs then gives c = 0, s = +-1 exactly, so it is
// a copy.  The remaining four form two dependency chains per row -- band1(it) -> band2(it) -> band1(it+1) and
// band1(it) -> A(it) -> B(it) -- which are issued as the pairs  band1(it) || B(it-1)  and  band2(it) || A(it).
template <class Points>
LTK_HD void solve(int N, const Io& io, Points pts)
{
    auto point = [&](int j, double& x, double& y) {
        typename Points::Raw raw;
        pts.fetch(j, raw);
        pts.finish(raw, x, y);
    };
    const int n10 = N - 2;
    // periodic knot extension: t(4 - j) = t(N + 4 - j) - per, t(N + 4 + j) = t(4 + j) + per
    const double per = LTK_T(N + 4) - LTK_T(4);
    for (int j = 1; j <= 3; ++j) {
        LTK_T(N + 4 + j) = LTK_T(4 + j) + per;
        LTK_T(4 - j) = LTK_T(N + 4 - j) - per;
    }
    // the two wrapping data rows: sites N-1 and N (1-based), l = site + 3
    WrapRow A, B;
    {
        double h1, h2, h3;
        int l = N + 2;
        bspl_at_knot(LTK_T(l - 2), LTK_T(l - 1), LTK_T(l), LTK_T(l + 1), LTK_T(l + 2), h1, h2, h3);
        A.h2[0] = 0.0 + h1; A.h2[1] = 0.0 + h2; A.h1[0] = h3; A.h1[1] = 0.0; A.h1[2] = 0.0;
        point(N - 2, A.x, A.y);
        l = N + 3;
        bspl_at_knot(LTK_T(l - 2), LTK_T(l - 1), LTK_T(l), LTK_T(l + 1), LTK_T(l + 2), h1, h2, h3);
        B.h2[0] = 0.0; B.h2[1] = 0.0 + h1; B.h1[0] = h2; B.h1[1] = h3; B.h1[2] = 0.0;
        point(N - 1, B.x, B.y);
    }
    Row W0 = {0, 0, 0, 0, 0, 0, 0}, W1 = W0, W2 = W0, Wp = W0;  // rows it, it+1, it+2; Wp = row it-1 awaiting B
    // software pipeline: the B-spline row and the control point of data row it+1 are produced during row it
    double tm2 = LTK_T(2), tm1 = LTK_T(3), t0 = LTK_T(4), tp1 = LTK_T(5), tp2 = LTK_T(6);
    double hn1, hn2, hn3, xn, yn;
    bspl_at_knot(tm2, tm1, t0, tp1, tp2, hn1, hn2, hn3);
    point(0, xn, yn);
    for (int it = 1; it <= n10; ++it) {
        double h1 = hn1, h2 = hn2, h3 = hn3, x = xn, y = yn, c, s, cw, sw;
        // operands of data row it+1 (row n10+1 = A exists as well: the values are simply not used after the loop);
        // the control point's loads are issued here and consumed at the end of the iteration
        tm2 = tm1; tm1 = t0; t0 = tp1; tp1 = tp2; tp2 = LTK_T(it + 6);
        typename Points::Raw raw;
        pts.fetch(it, raw);
        BsplMid mid;
        // ---- band1(it) || B(it-1)  [+ first half of the next B-spline row]
        if (it >= 2 && h1 != 0.0 && B.h1[0] != 0.0) {
            givens2(h1, W0.a, c, s, B.h1[0], Wp.a, cw, sw);
            mid = bspl_begin(tm1, t0, tp1);
            rota(c, s, x, W0.zx);
            rota(c, s, y, W0.zy);
            rota(c, s, h2, W0.b);
            rota(c, s, h3, W0.c);
            wrap_apply(B, Wp, it - 1, n10, cw, sw);
        } else {
            mid = bspl_begin(tm1, t0, tp1);
            if (h1 != 0.0) {
                givens(h1, W0.a, c, s);
                rota(c, s, x, W0.zx);
                rota(c, s, y, W0.zy);
                rota(c, s, h2, W0.b);
                rota(c, s, h3, W0.c);
            }
            if (it >= 2) wrap_rotate(B, Wp, it - 1, n10);
        }
        if (it >= 2) LTK_STORE_ROW(it - 1, Wp);
        // row `it` is final for the band part; its entries beyond column n10 belong to the periodic block
        if (it == n10 - 1) W0.e1 = W0.c;
        if (it == n10) { W0.e1 = W0.b; W0.e2 = W0.c; }
        // ---- band2(it) || A(it)  [+ second half of the next B-spline row]
        if (h2 != 0.0 && A.h1[0] != 0.0) {
            givens2(h2, W1.a, c, s, A.h1[0], W0.a, cw, sw);
            bspl_end(mid, tm2, tm1, t0, tp1, tp2, hn1, hn2, hn3);
            rota(c, s, x, W1.zx);
            rota(c, s, y, W1.zy);
            rota(c, s, h3, W1.b);
            wrap_apply(A, W0, it, n10, cw, sw);
        } else {
            bspl_end(mid, tm2, tm1, t0, tp1, tp2, hn1, hn2, hn3);
            if (h2 != 0.0) {
                givens(h2, W1.a, c, s);
                rota(c, s, x, W1.zx);
                rota(c, s, y, W1.zy);
                rota(c, s, h3, W1.b);
            }
            wrap_rotate(A, W0, it, n10);
        }
        // ---- band3(it): row it+2 is fresh (a = 0, z = 0), so fpgivs yields dd = |h3|, c = 0, s = h3 / |h3|
        if (h3 != 0.0) {
            const double sg = (h3 < 0.0) ? -1.0 : 1.0;
            W2.a = fabs(h3);
            W2.zx = 0.0 + sg * x;
            W2.zy = 0.0 + sg * y;
        }
        Wp = W0; W0 = W1; W1 = W2;
        W2.a = W2.b = W2.c = W2.e1 = W2.e2 = W2.zx = W2.zy = 0.0;
        pts.finish(raw, xn, yn);
    }
    wrap_rotate(B, Wp, n10, n10);
    LTK_STORE_ROW(n10, Wp);
    // W0 = row N-1: a2(N-1, 1..2) = (a, b); W1 = row N: a2(N, 2) = a
    wrap_tail(A, W0, W1);
    wrap_tail(B, W0, W1);
    // fpbacp; the coefficients go to cx, cy (c(N + q) = c(q), q = 1..3)
#define LTK_CX(i) io.cx[(long)((i) - 1) * io.sc]
#define LTK_CY(i) io.cy[(long)((i) - 1) * io.sc]
    const double cNx = fdiv(W1.zx, W1.a), cNy = fdiv(W1.zy, W1.a);
    const double cMx = fdiv(W0.zx - cNx * W0.b, W0.a), cMy = fdiv(W0.zy - cNy * W0.b, W0.a);
    LTK_CX(N) = cNx; LTK_CY(N) = cNy; LTK_CX(N - 1) = cMx; LTK_CY(N - 1) = cMy;
    double c1x = 0, c2x = 0, c1y = 0, c2y = 0;  // c(i+1), c(i+2)
    // The factor rows come back from the global scratch (written by this thread during the forward pass, by now in
    // L2 or DRAM: ~1,000 cycles) while one step of the recurrence is ~150 cycles, and a warp issues in order: the
    // loads of BU rows are issued together, then the BU dependent steps run (ncu: 20 % of the kernel's stall samples
    // were these loads when each row was fetched one step ahead).
    constexpr int BU = 6;
    for (int i0 = n10; i0 >= 1; i0 -= BU) {
        double ra[BU], rb[BU], rc[BU], re1[BU], re2[BU], rzx[BU], rzy[BU];
#pragma unroll
        for (int u = 0; u < BU; ++u) {
            const int i = (i0 - u >= 1) ? i0 - u : 1;  // clamped: the surplus loads of the last block are not used
            ra[u] = LTK_R(i, 0); rb[u] = LTK_R(i, 1); rc[u] = LTK_R(i, 2); re1[u] = LTK_R(i, 3);
            re2[u] = LTK_R(i, 4); rzx[u] = LTK_R(i, 5); rzy[u] = LTK_R(i, 6);
        }
#pragma unroll
        for (int u = 0; u < BU; ++u) {
            const int i = i0 - u;
            if (i >= 1) {
                double sx = rzx[u], sy = rzy[u];
                sx = sx - cMx * re1[u]; sx = sx - cNx * re2[u];
                sy = sy - cMy * re1[u]; sy = sy - cNy * re2[u];
                if (i <= n10 - 1) { sx = sx - c1x * rb[u]; sy = sy - c1y * rb[u]; }
                if (i <= n10 - 2) { sx = sx - c2x * rc[u]; sy = sy - c2y * rc[u]; }
                fdiv2(sx, ra[u], sy, ra[u], sx, sy);
                c2x = c1x; c2y = c1y; c1x = sx; c1y = sy;
                LTK_CX(i) = sx; LTK_CY(i) = sy;
            }
        }
    }
    for (int q = 1; q <= 3; ++q) { LTK_CX(N + q) = LTK_CX(q); LTK_CY(N + q) = LTK_CY(q); }
    // splder: wrk1(i) = 3 (c(i+1) - c(i)) / (t(i+4) - t(i+1)),  wrk2(i) = 2 (wrk1(i+1) - wrk1(i)) / (t(i+4) - t(i+2))
    // same blocking for the coefficient loads of the two derivative passes (c(N + q) = c(q) were just stored)
    double px_ = LTK_CX(1), py_ = LTK_CY(1), w1px = 0, w1py = 0;
    constexpr int WU = 8;
    for (int ib = 1; ib <= N + 2; ib += WU) {
        double nxs[WU], nys[WU];
#pragma unroll
        for (int u = 0; u < WU; ++u) {
            const int i = (ib + u <= N + 2) ? ib + u : N + 2;
            nxs[u] = LTK_CX(i + 1); nys[u] = LTK_CY(i + 1);
        }
#pragma unroll
        for (int u = 0; u < WU; ++u) {
            const int i = ib + u;
            if (i <= N + 2) {
                const double cx1 = nxs[u], cy1 = nys[u];
                const double fac = LTK_T(i + 4) - LTK_T(i + 1);
                double w1x, w1y;
                fdiv2(3.0 * (cx1 - px_), fac, 3.0 * (cy1 - py_), fac, w1x, w1y);
                io.w1x[(long)(i - 1) * io.sw] = w1x;
                io.w1y[(long)(i - 1) * io.sw] = w1y;
                if (i >= 2) {  // wrk2(i-1) from wrk1(i-1), wrk1(i): knots t(i+3), t(i+1)
                    const double fac2 = LTK_T(i + 3) - LTK_T(i + 1);
                    double w2x, w2y;
                    fdiv2(2.0 * (w1x - w1px), fac2, 2.0 * (w1y - w1py), fac2, w2x, w2y);
                    io.w2x[(long)(i - 2) * io.sw] = w2x;
                    io.w2y[(long)(i - 2) * io.sw] = w2y;
                }
                w1px = w1x; w1py = w1y; px_ = cx1; py_ = cy1;
            }
        }
    }
#undef LTK_CX
#undef LTK_CY
}
#undef LTK_T
#undef LTK_R
#undef LTK_STORE_ROW

// One spline interval t(l) <= x < t(l+1) as the sample loop needs it (splder + fpbspl of degree 2 and 1).
struct __align__(16) FitInterval {
    double t0, tp1;           // t(l), t(l+1): the interval
    double tm1, tp2;          // t(l-1), t(l+2)
    double inv01;             // 1 / (t(l+1) - t(l))
    double d1, r1;            // t(l+1) - t(l-1) and its reciprocal
    double d2, r2;            // t(l+2) - t(l)   and its reciprocal
    double w1x[3], w1y[3];    // wrk1(l-3 .. l-1)
    double w2x[2], w2y[2];    // wrk2(l-3 .. l-2)
    double pad;
};

// q = a / d with r = RN(1/d): one product and one exact-residual correction (Markstein).  The corrected value is
// the correctly rounded quotient unless a/d lies within ~2^-53 ulp of a rounding boundary.
LTK_HD double div_by(double a, double d, double r)
{
#if defined(__CUDA_ARCH__)
    const double q = a * r;
    const double rem = fma(-q, d, a);
    return fma(rem, r, q);
#else
    (void)r;
    return a / d;
#endif
}

// x**1.5 rounded to nearest: square root and product carried in double-double (numpy's `** (3/2)`, path.py:58).
// The correction e = (x - s*s) / (2 s) is a 2^-53-relative term, so the refined reciprocal square root of the
// square-root sequence itself (good to ~2^-50) serves as 1/s: the sum below rounds like oracle/lap_oracle.c's
// lto_pow15, which divides (the two could only differ if x**1.5 lay within 2^-100 of a rounding boundary).
LTK_HD double pow15(double x)
{
#if defined(__CUDA_ARCH__)
    double y = rsqrt_seed(x);
    const double e0 = fma(x, -(y * y), 1.0);
    const double q = fma(e0, 0.375, 0.5);
    y = fma(y, q * e0, y);
    const double g = x * y;
    const double hy = half_of(y);
    const double s = fma(fma(g, -g, x), hy, g);  // = dsqrt<false>(x)
    const double e = fma(-s, s, x) * hy;
#else
    const double s = sqrt(x);
    const double e = fma(-s, s, x) / (s + s);
#endif
    const double p = x * s;
    const double pe = fma(x, s, -p);
    return p + fma(x, e, pe);
}

// splev(x, tck, der=1) and der=2 in both coordinates, then the curvature of path.py:58,61 in numpy's order.
LTK_HD double curvature_at(const FitInterval& v, double x, double& dx, double& dy, double& ddx, double& ddy)
{
    const double a = v.tp1 - x, b = x - v.t0;
    const double g1 = v.inv01 * a, g2 = v.inv01 * b;  // degree-1 B-splines
    ddx = v.w2x[0] * g1; ddx = ddx + v.w2x[1] * g2;
    ddy = v.w2y[0] * g1; ddy = ddy + v.w2y[1] * g2;
    const double f1 = div_by(g1, v.d1, v.r1);
    const double h1 = f1 * a;
    double h2 = f1 * (x - v.tm1);
    const double f2 = div_by(g2, v.d2, v.r2);
    h2 = h2 + f2 * (v.tp2 - x);
    const double h3 = f2 * b;
    dx = v.w1x[0] * h1; dx = dx + v.w1x[1] * h2; dx = dx + v.w1x[2] * h3;
    dy = v.w1y[0] * h1; dy = dy + v.w1y[1] * h2; dy = dy + v.w1y[2] * h3;
    const double cross = dx * ddy - dy * ddx;
    const double n2 = dx * dx + dy * dy;
    return fabs(fdiv(cross, pow15(n2)));
}

// ---- general evaluation (Path facade): splev / splder at one point of any cubic spline (t, c) ---------------------
// fpbspl for degree k <= 3 at t(l) <= x < t(l+1); t is contiguous, FITPACK's t(l) = t[l - 1]; h[0 .. k]
LTK_HD void bspl(const double* t, int k, double x, int l, double* h)
{
    double hh[4];
    h[0] = 1.0;
    for (int j = 1; j <= k; ++j) {
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 1; i <= j; ++i) {
            const double tli = t[l + i - 1], tlj = t[l + i - j - 1];
            if (tli == tlj) {
                h[i] = 0.0;
                continue;
            }
            const double f = fdiv(hh[i - 1], tli - tlj);
            h[i - 1] = h[i - 1] + f * (tli - x);
            h[i] = f * (x - tlj);
        }
    }
}

// splder's coefficient passes for nu = 1 and 2: w1 [n - 5], w2 [n - 6] from c [n - 4]
LTK_HD void der_coeffs(const double* t, int n, const double* c, double* w1, double* w2)
{
    const int nk1 = n - 4;
    for (int i = 1; i <= nk1 - 1; ++i) {
        const double fac = t[i + 4 - 1] - t[i + 1 - 1];
        w1[i - 1] = (fac <= 0.0) ? c[i - 1] : fdiv(3.0 * (c[i] - c[i - 1]), fac);
    }
    for (int i = 1; i <= nk1 - 2; ++i) {
        const double fac = t[i + 4 - 1] - t[i + 2 - 1];
        w2[i - 1] = (fac <= 0.0) ? w1[i - 1] : fdiv(2.0 * (w1[i] - w1[i - 1]), fac);
    }
}

// knot interval of x as splev / splder pick it: the largest l in [4, n - 4] with t(l) <= x (l = 4 below the span)
LTK_HD int find_interval(const double* t, int n, double x)
{
    int lo = 4, hi = n - 4;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t[mid - 1] <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// value of the nu-th derivative at x in interval l; coef = c (nu = 0), w1 (nu = 1) or w2 (nu = 2)
LTK_HD double splev_at(const double* t, const double* coef, int nu, double x, int l)
{
    double h[4];
    bspl(t, 3 - nu, x, l, h);
    double sp = 0.0;
    for (int j = 0; j <= 3 - nu; ++j) sp = sp + coef[l - 4 + j] * h[j];
    return sp;
}

// Open interpolating cubic spline through m >= 4 points (splprep(..., per=0, s=0): fppara with the not-a-knot knot
// vector, banded Givens QR, fpback); reference path.py:25 with closed = False.  u [m], px, py [m] -> t [m + 4],
// cx, cy [m]; scratch a [4 m], z [2 m].  Contiguous arrays, one thread.
LTK_HD void solve_open(int m, const double* u, const double* px, const double* py, double* t, double* a, double* z,
                       double* cx, double* cy)
{
    const int n = m + 4, nk1 = m;
    for (int i = 0; i < 4; ++i) { t[i] = u[0]; t[nk1 + i] = u[m - 1]; }
    for (int q = 0; q < m - 4; ++q) t[4 + q] = u[2 + q];
    for (int i = 0; i < 4 * m; ++i) a[i] = 0.0;
    for (int i = 0; i < 2 * m; ++i) z[i] = 0.0;
#define LTK_A(j, e) a[((j) - 1) * 4 + (e) - 1]
    int l = 4;
    for (int it = 1; it <= m; ++it) {
        const double ui = u[it - 1];
        double x = px[it - 1], y = py[it - 1];
        while (!(ui < t[l] || l == nk1)) ++l;
        double h[4];
        bspl(t, 3, ui, l, h);
        int j = l - 4;
        for (int i = 1; i <= 4; ++i) {
            ++j;
            const double piv = h[i - 1];
            if (piv == 0.0) continue;
            double c, s;
            givens(piv, LTK_A(j, 1), c, s);
            rota(c, s, x, z[j - 1]);
            rota(c, s, y, z[m + j - 1]);
            if (i == 4) break;
            int i2 = 1;
            for (int i1 = i + 1; i1 <= 4; ++i1) {
                ++i2;
                rota(c, s, h[i1 - 1], LTK_A(j, i2));
            }
        }
    }
    for (int d = 0; d < 2; ++d) {  // fpback
        const double* zz = z + d * m;
        double* cc = d ? cy : cx;
        cc[nk1 - 1] = fdiv(zz[nk1 - 1], LTK_A(nk1, 1));
        int i = nk1 - 1;
        for (int j = 2; j <= nk1; ++j) {
            double store = zz[i - 1];
            const int i1 = j <= 3 ? j - 1 : 3;
            int mm = i;
            for (int q = 1; q <= i1; ++q) {
                ++mm;
                store = store - cc[mm - 1] * LTK_A(i, q + 1);
            }
            cc[i - 1] = fdiv(store, LTK_A(i, 1));
            --i;
        }
    }
#undef LTK_A
    (void)n;
}

}  // namespace fit
}  // namespace ltk
