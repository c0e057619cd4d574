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
    f = fdiv(g1, tp1 - tm2);
    h1 = f * a;                                          // degree 3
    h2 = f * (t0 - tm2);
    f = fdiv(g2, tp2 - tm1);
    h2 = h2 + f * (tp2 - t0);
    h3 = f * (t0 - tm1);
}

struct Row {  // one row of the triangular factor: band part a1(j, 1..3), periodic part a2(j, 1..2), z(j) per dim
    double a, b, c, e1, e2, zx, zy;
};

struct WrapRow {  // a wrapping data row being rotated through the factor
    double h1[3], h2[2], x, y;
};

// one step of "rotation with the rows 1,2,...n10 of matrix a" (fpclos) for row j (1-based), n10 = N - 2
LTK_HD void wrap_rotate(WrapRow& S, Row& R, int j, int n10)
{
    const double piv = S.h1[0];
    if (piv == 0.0) {
        S.h1[0] = S.h1[1]; S.h1[1] = S.h1[2]; S.h1[2] = 0.0;
        return;
    }
    double c, s;
    givens(piv, R.a, c, s);
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
    const double* px; const double* py; long sp;  // unique control points [N]
    double* t; long st;                            // knots [N + 7]; entries 3 .. N+3 hold the chord-length knots on entry
    double* rows; long sr;                         // scratch: 7 N entries
    double* cx; double* cy; long sc;               // B-spline coefficients [N + 3] each (nullptr: not wanted)
    double* w1x; double* w1y; double* w2x; double* w2y; long sw;  // derivative coefficients [N+2], [N+2], [N+1], [N+1]
};

#define LTK_T(l) io.t[(long)((l) - 1) * io.st]          /* 1-based knot index as in FITPACK */
#define LTK_R(j, e) io.rows[(long)(((j) - 1) * 7 + (e)) * io.sr]

// The whole solve for one candidate with N >= 5 unique control points.
LTK_HD void solve(int N, const Io& io)
{
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
        A.x = io.px[(long)(N - 2) * io.sp]; A.y = io.py[(long)(N - 2) * io.sp];
        l = N + 3;
        bspl_at_knot(LTK_T(l - 2), LTK_T(l - 1), LTK_T(l), LTK_T(l + 1), LTK_T(l + 2), h1, h2, h3);
        B.h2[0] = 0.0; B.h2[1] = 0.0 + h1; B.h1[0] = h2; B.h1[1] = h3; B.h1[2] = 0.0;
        B.x = io.px[(long)(N - 1) * io.sp]; B.y = io.py[(long)(N - 1) * io.sp];
    }
    Row W0 = {0, 0, 0, 0, 0, 0, 0}, W1 = W0, W2 = W0;
    double tm2 = LTK_T(2), tm1 = LTK_T(3), t0 = LTK_T(4), tp1 = LTK_T(5);
    for (int it = 1; it <= n10; ++it) {
        const double tp2 = LTK_T(it + 5);
        double h1, h2, h3, c, s;
        bspl_at_knot(tm2, tm1, t0, tp1, tp2, h1, h2, h3);
        tm2 = tm1; tm1 = t0; t0 = tp1; tp1 = tp2;
        double x = io.px[(long)(it - 1) * io.sp], y = io.py[(long)(it - 1) * io.sp];
        if (h1 != 0.0) {
            givens(h1, W0.a, c, s);
            rota(c, s, x, W0.zx);
            rota(c, s, y, W0.zy);
            rota(c, s, h2, W0.b);
            rota(c, s, h3, W0.c);
        }
        if (h2 != 0.0) {
            givens(h2, W1.a, c, s);
            rota(c, s, x, W1.zx);
            rota(c, s, y, W1.zy);
            rota(c, s, h3, W1.b);
        }
        if (h3 != 0.0) {
            givens(h3, W2.a, c, s);
            rota(c, s, x, W2.zx);
            rota(c, s, y, W2.zy);
        }
        // row `it` is final for the band part; its entries beyond column n10 belong to the periodic block
        if (it == n10 - 1) W0.e1 = W0.c;
        if (it == n10) { W0.e1 = W0.b; W0.e2 = W0.c; }
        wrap_rotate(A, W0, it, n10);
        wrap_rotate(B, W0, it, n10);
        LTK_R(it, 0) = W0.a; LTK_R(it, 1) = W0.b; LTK_R(it, 2) = W0.c; LTK_R(it, 3) = W0.e1;
        LTK_R(it, 4) = W0.e2; LTK_R(it, 5) = W0.zx; LTK_R(it, 6) = W0.zy;
        W0 = W1; W1 = W2;
        W2.a = W2.b = W2.c = W2.e1 = W2.e2 = W2.zx = W2.zy = 0.0;
    }
    // W0 = row N-1: a2(N-1, 1..2) = (a, b); W1 = row N: a2(N, 2) = a
    wrap_tail(A, W0, W1);
    wrap_tail(B, W0, W1);
    // fpbacp
    const double cNx = fdiv(W1.zx, W1.a), cNy = fdiv(W1.zy, W1.a);
    const double cMx = fdiv(W0.zx - cNx * W0.b, W0.a), cMy = fdiv(W0.zy - cNy * W0.b, W0.a);
    double c1x = 0, c2x = 0, c1y = 0, c2y = 0;  // c(i+1), c(i+2)
    for (int i = n10; i >= 1; --i) {
        const double a = LTK_R(i, 0), b = LTK_R(i, 1), cc = LTK_R(i, 2), e1 = LTK_R(i, 3), e2 = LTK_R(i, 4);
        double sx = LTK_R(i, 5), sy = LTK_R(i, 6);
        sx = sx - cMx * e1; sx = sx - cNx * e2;
        sy = sy - cMy * e1; sy = sy - cNy * e2;
        if (i <= n10 - 1) { sx = sx - c1x * b; sy = sy - c1y * b; }
        if (i <= n10 - 2) { sx = sx - c2x * cc; sy = sy - c2y * cc; }
        sx = fdiv(sx, a); sy = fdiv(sy, a);
        c2x = c1x; c2y = c1y; c1x = sx; c1y = sy;
        LTK_R(i, 5) = sx; LTK_R(i, 6) = sy;  // c(i) takes the place of z(i)
    }
#define LTK_CX(i) ((i) > N ? LTK_CX0((i) - N) : LTK_CX0(i))
#define LTK_CX0(i) ((i) == N ? cNx : (i) == N - 1 ? cMx : LTK_R(i, 5))
#define LTK_CY(i) ((i) > N ? LTK_CY0((i) - N) : LTK_CY0(i))
#define LTK_CY0(i) ((i) == N ? cNy : (i) == N - 1 ? cMy : LTK_R(i, 6))
    if (io.cx) {
        for (int i = 1; i <= N + 3; ++i) {
            io.cx[(long)(i - 1) * io.sc] = LTK_CX(i);
            io.cy[(long)(i - 1) * io.sc] = LTK_CY(i);
        }
    }
    // splder: wrk1(i) = 3 (c(i+1) - c(i)) / (t(i+4) - t(i+1)),  wrk2(i) = 2 (wrk1(i+1) - wrk1(i)) / (t(i+4) - t(i+2))
    double px_ = LTK_CX(1), py_ = LTK_CY(1), w1px = 0, w1py = 0;
    for (int i = 1; i <= N + 2; ++i) {
        const double nx = LTK_CX(i + 1), ny = LTK_CY(i + 1);
        const double t4 = LTK_T(i + 4);
        const double fac = t4 - LTK_T(i + 1);
        const double w1x = fdiv(3.0 * (nx - px_), fac), w1y = fdiv(3.0 * (ny - py_), fac);
        io.w1x[(long)(i - 1) * io.sw] = w1x;
        io.w1y[(long)(i - 1) * io.sw] = w1y;
        if (i >= 2) {  // wrk2(i-1) from wrk1(i-1), wrk1(i): knots t(i+3), t(i+1)
            const double fac2 = LTK_T(i + 3) - LTK_T(i + 1);
            io.w2x[(long)(i - 2) * io.sw] = fdiv(2.0 * (w1x - w1px), fac2);
            io.w2y[(long)(i - 2) * io.sw] = fdiv(2.0 * (w1y - w1py), fac2);
        }
        w1px = w1x; w1py = w1y; px_ = nx; py_ = ny;
    }
#undef LTK_CX
#undef LTK_CX0
#undef LTK_CY
#undef LTK_CY0
}
#undef LTK_T
#undef LTK_R

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

}  // namespace fit
}  // namespace ltk
