/* ORACLE (test infrastructure, not product code) -- level 2b: FITPACK's periodic interpolating cubic
 * spline, restated in plain C from Dierckx's published algorithm.
 *
 * The reference builds its racing line with scipy.interpolate.splprep(controls, u=dists, k=3, s=0, per=1)
 * (/root/reference/src/path.py:25) and differentiates it with splev(u, tck, der=1|2) (path.py:51-54).
 * SciPy is a third-party dependency (requirements.txt:14 pins scipy==1.13.0; the build container and the
 * GPU box carry 1.18.1) and is absent from /root/reference, so this file restates the routines that call
 * reaches -- clocur -> fpclos (s = 0 branch) with fpbspl / fpgivs / fprota / fpbacp, and splder -- after
 * P. Dierckx, "Algorithms for smoothing data with periodic and parametric splines", CGIP 20 (1982) and
 * "Curve and surface fitting with splines", OUP 1993, in the operation order of the netlib FITPACK
 * routines of those names.
 *
 * Parity status: PINNED by tests/test_oracle.py::test_fitpack_port_* -- knots and B-spline coefficients
 * np.array_equal to scipy.interpolate.splprep's tck, derivative values np.array_equal to splev, on the
 * control polygons of every golden case and on random polygons (live SciPy is importable wherever the
 * tests run).
 *
 * Arrays are 1-based inside (index 0 unused) to keep Dierckx's subscripts.
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* fpbspl: the k+1 non-zero B-splines of degree k at t(l) <= x < t(l+1), de Boor-Cox recurrence. */
static void fpbspl(const double *t, int k, double x, int l, double *h /* 1..k+1 */)
{
    double hh[7];
    h[1] = 1.0;
    for (int j = 1; j <= k; ++j) {
        for (int i = 1; i <= j; ++i) hh[i] = h[i];
        h[1] = 0.0;
        for (int i = 1; i <= j; ++i) {
            int li = l + i, lj = li - j;
            if (t[li] == t[lj]) {
                h[i + 1] = 0.0;
                continue;
            }
            double f = hh[i] / (t[li] - t[lj]);
            h[i] = h[i] + f * (t[li] - x);
            h[i + 1] = f * (x - t[lj]);
        }
    }
}

/* fpgivs: parameters of a Givens rotation that annihilates piv against the diagonal element ww. */
static void fpgivs(double piv, double *ww, double *c, double *s)
{
    double store = fabs(piv), dd, r;
    if (store >= *ww) {
        r = *ww / piv;
        dd = store * sqrt(1.0 + r * r);
    } else {
        r = piv / *ww;
        dd = *ww * sqrt(1.0 + r * r);
    }
    *c = *ww / dd;
    *s = piv / dd;
    *ww = dd;
}

/* fprota: apply the rotation to the pair (a, b). */
static void fprota(double c, double s, double *a, double *b)
{
    double s1 = *a, s2 = *b;
    *b = c * s2 + s * s1;
    *a = c * s1 - s * s2;
}

/* Periodic interpolating spline of degree k = 3 through m points (the last equals the first) of a curve in
 * idim dimensions: the s = 0, odd-k branch of fpclos.  u[0..m-1] parameter values, x[(i*idim)+d] data
 * (point-major, as splprep passes ravel(transpose(x))).  Outputs: t[0..n-1] with n = m + 6, c[d*n + i]
 * (n per dimension, the last k+1 of each are the periodic copies / padding as in FITPACK).
 * Returns n, or -1 on bad input.  Unit weights (splprep's default) -- the products with w are exact. */
int fpk_clocur(const double *u0, const double *x0, int m, int idim, double *t0, double *c0)
{
    const int k = 3, k1 = 4;
    if (m < 5 || idim < 1 || idim > 4) return -1; /* need n10 = m - 3 >= 2 */
    int n = m + 2 * k, nk1 = n - k1, nk2 = nk1 + 1, m1 = m - 1;
    int kk = k - 1, kk1 = k; /* interpolation at the knots: only k B-splines are non-zero per row */
    int n7 = nk1 - k, n10 = n7 - kk;
    const double *u = u0 - 1, *x = x0 - 1;
    double *t = t0 - 1;
    double per = u[m] - u[1];
    /* knots: interior knots at the data sites, periodic extension on both sides */
    for (int i = 2; i <= m1; ++i) t[i + k] = u[i];
    t[k1] = u[1];
    t[nk2] = u[m];
    for (int j = 1; j <= k; ++j) {
        t[nk2 + j] = t[k1 + j] + per;
        t[k1 - j] = t[nk2 - j] - per;
    }
    /* work arrays (1-based): a1[nk1+1][kk1+1], a2[n7+1][kk+1], z[idim*n + 1] */
    double(*a1)[5] = calloc((size_t)nk1 + 2, sizeof *a1);
    double(*a2)[4] = calloc((size_t)n7 + 2, sizeof *a2);
    double *z = calloc((size_t)idim * n + 2, sizeof *z);
    double *c = c0 - 1;
    for (int i = 1; i <= idim * n; ++i) c[i] = 0.0;
    double h[8], h1[8], h2[8], xi[4];
    int jper = 0, l = k1, jj = 0;
    for (int it = 1; it <= m1; ++it) {
        double ui = u[it];
        for (int j = 0; j < idim; ++j) xi[j] = x[++jj];
        while (!(ui < t[l + 1])) ++l;
        fpbspl(t, k, ui, l, h);
        int l5 = l - k1;
        if (l5 < n10) {
            /* the row touches the band part only */
            int j = l5;
            for (int i = 1; i <= kk1; ++i) {
                ++j;
                double piv = h[i];
                if (piv == 0.0) continue;
                double co, si;
                fpgivs(piv, &a1[j][1], &co, &si);
                for (int d = 0, j1 = j; d < idim; ++d, j1 += n) fprota(co, si, &xi[d], &z[j1]);
                if (i == kk1) break;
                int i2 = 1;
                for (int i1 = i + 1; i1 <= kk1; ++i1) {
                    ++i2;
                    fprota(co, si, &h[i1], &a1[j][i2]);
                }
            }
            continue;
        }
        if (!jper) {
            /* first wrapping row: move the columns of a1 that lie beyond n10 into a2 */
            for (int i = 1; i <= n7; ++i)
                for (int j = 1; j <= kk; ++j) a2[i][j] = 0.0;
            int jk = n10 + 1;
            for (int i = 1; i <= kk; ++i) {
                int ik = jk;
                for (int j = 1; j <= kk1; ++j) {
                    if (ik <= 0) break;
                    a2[ik][i] = a1[ik][j];
                    --ik;
                }
                ++jk;
            }
            jper = 1;
        }
        /* split the row into its a1 part (h1) and its a2 part (h2) using the periodicity condition */
        for (int i = 1; i <= kk; ++i) h1[i] = h2[i] = 0.0;
        h1[kk1] = 0.0;
        int j = l5 - n10;
        for (int i = 1; i <= kk1; ++i) {
            ++j;
            int l0 = j, l1;
            for (;;) {
                l1 = l0 - kk;
                if (l1 <= 0) { h2[l0] = h2[l0] + h[i]; break; }
                if (l1 <= n10) { h1[l1] = h[i]; break; }
                l0 = l1 - n10;
            }
        }
        /* rotate through rows 1..n10 */
        int done = 0;
        for (j = 1; j <= n10 && !done; ++j) {
            double piv = h1[1];
            if (piv == 0.0) {
                for (int i = 1; i <= kk; ++i) h1[i] = h1[i + 1];
                h1[kk1] = 0.0;
                continue;
            }
            double co, si;
            fpgivs(piv, &a1[j][1], &co, &si);
            for (int d = 0, j1 = j; d < idim; ++d, j1 += n) fprota(co, si, &xi[d], &z[j1]);
            for (int i = 1; i <= kk; ++i) fprota(co, si, &h2[i], &a2[j][i]);
            if (j == n10) { done = 1; break; }
            int i2 = n10 - j < kk ? n10 - j : kk, i1 = 1;
            for (int i = 1; i <= i2; ++i) {
                i1 = i + 1;
                fprota(co, si, &h1[i1], &a1[j][i1]);
                h1[i] = h1[i1];
            }
            h1[i1] = 0.0;
        }
        /* rotate through rows n10+1..n7 */
        for (j = 1; j <= kk; ++j) {
            int ij = n10 + j;
            if (ij <= 0) continue;
            double piv = h2[j];
            if (piv == 0.0) continue;
            double co, si;
            fpgivs(piv, &a2[ij][j], &co, &si);
            for (int d = 0, j1 = ij; d < idim; ++d, j1 += n) fprota(co, si, &xi[d], &z[j1]);
            if (j == kk) break;
            for (int i = j + 1; i <= kk; ++i) fprota(co, si, &h2[i], &a2[ij][i]);
        }
    }
    /* fpbacp: back substitution, per dimension */
    for (int d = 0; d < idim; ++d) {
        double *zz = z + d * n, *cc = c + d * n;
        int nn = n7, n2 = nn - kk, ll = nn;
        for (int i = 1; i <= kk; ++i) {
            double store = zz[ll];
            int j = kk + 2 - i;
            if (i != 1) {
                int l0 = ll;
                for (int l1 = j; l1 <= kk; ++l1) {
                    ++l0;
                    store = store - cc[l0] * a2[ll][l1];
                }
            }
            cc[ll] = store / a2[ll][j - 1];
            --ll;
            if (ll == 0) break;
        }
        for (int i = 1; i <= n2; ++i) {
            double store = zz[i];
            ll = n2;
            for (int j = 1; j <= kk; ++j) {
                ++ll;
                store = store - cc[ll] * a2[i][j];
            }
            cc[i] = store;
        }
        int i = n2;
        cc[i] = cc[i] / a1[i][1];
        for (int j = 2; j <= n2; ++j) {
            --i;
            double store = cc[i];
            int i1 = j <= kk ? j - 1 : kk;
            ll = i;
            for (int l0 = 1; l0 <= i1; ++l0) {
                ++ll;
                store = store - cc[ll] * a1[i][l0 + 1];
            }
            cc[i] = store / a1[i][1];
        }
        /* periodicity: the last k coefficients repeat the first k */
        for (int q = 1; q <= k; ++q) cc[q + n7] = cc[q];
    }
    free(a1);
    free(a2);
    free(z);
    return n;
}

/* splder: derivative of order nu (0..k-1) of the spline (t, c, k = 3) at x[0..m-1] (inside the knot span).
 * wrk0 must hold n doubles. */
void fpk_splder(const double *t0, int n, const double *c0, int nu, const double *x, int m, double *y,
                double *wrk0)
{
    const int k = 3, k1 = 4;
    const double *t = t0 - 1, *c = c0 - 1;
    double *wrk = wrk0 - 1;
    int nk1 = n - k1, kk = k, l = 1;
    for (int i = 1; i <= nk1; ++i) wrk[i] = c[i];
    int nk2 = nk1;
    for (int j = 1; j <= nu; ++j) {
        double ak = kk;
        --nk2;
        int l1 = l;
        for (int i = 1; i <= nk2; ++i) {
            ++l1;
            int l2 = l1 + kk;
            double fac = t[l2] - t[l1];
            if (fac <= 0.0) continue;
            wrk[i] = ak * (wrk[i + 1] - wrk[i]) / fac;
        }
        ++l;
        --kk;
    }
    l = k1;
    int k2 = k1 - nu;
    double h[8];
    for (int i = 0; i < m; ++i) {
        double arg = x[i];
        while (!(arg < t[l + 1] || l == nk1)) ++l;
        while (arg < t[l] && l > k1) --l;
        fpbspl(t, kk, arg, l, h);
        double sp = 0.0;
        int ll = l - k1;
        for (int j = 1; j <= k2; ++j) {
            ++ll;
            sp = sp + wrk[ll] * h[j];
        }
        y[i] = sp;
    }
}

/* Open interpolating spline of degree 3 through m points (splprep(..., per=0, s=0): parcur -> fppara with the
 * not-a-knot knot vector t = [u_1 x4, u_3 .. u_{m-2}, u_m x4]; /root/reference/src/path.py:25 with closed = False,
 * reached from trajectory.py:189-192).  Row-by-row Givens QR of the banded collocation matrix, fpback.
 * Outputs t[0..n-1], n = m + 4, and c[d*n + i].  Returns n, or -1. */
int fpk_parcur_open(const double *u0, const double *x0, int m, int idim, double *t0, double *c0)
{
    const int k = 3, k1 = 4;
    if (m < 4 || idim < 1 || idim > 4) return -1;
    int n = m + k1, nk1 = n - k1;
    const double *u = u0 - 1, *x = x0 - 1;
    double *t = t0 - 1, *c = c0 - 1;
    for (int i = 1; i <= k1; ++i) {
        t[i] = u[1];
        t[nk1 + i] = u[m];
    }
    for (int l = 1, i = k + 2, j = k / 2 + 2; l <= m - k1; ++l) t[i++] = u[j++];
    double(*a)[5] = calloc((size_t)nk1 + 2, sizeof *a);
    double *z = calloc((size_t)idim * n + 2, sizeof *z);
    for (int i = 1; i <= idim * n; ++i) c[i] = 0.0;
    double h[8], xi[4];
    int l = k1, jj = 0;
    for (int it = 1; it <= m; ++it) {
        double ui = u[it];
        for (int d = 0; d < idim; ++d) xi[d] = x[++jj];
        while (!(ui < t[l + 1] || l == nk1)) ++l;
        fpbspl(t, k, ui, l, h);
        int j = l - k1;
        for (int i = 1; i <= k1; ++i) {
            ++j;
            double piv = h[i];
            if (piv == 0.0) continue;
            double co, si;
            fpgivs(piv, &a[j][1], &co, &si);
            for (int d = 0, j1 = j; d < idim; ++d, j1 += n) fprota(co, si, &xi[d], &z[j1]);
            if (i == k1) break;
            int i2 = 1;
            for (int i1 = i + 1; i1 <= k1; ++i1) {
                ++i2;
                fprota(co, si, &h[i1], &a[j][i2]);
            }
        }
    }
    for (int d = 0; d < idim; ++d) { /* fpback */
        double *zz = z + d * n, *cc = c + d * n;
        cc[nk1] = zz[nk1] / a[nk1][1];
        int i = nk1 - 1;
        for (int j = 2; j <= nk1; ++j) {
            double store = zz[i];
            int i1 = j <= k ? j - 1 : k, mm = i;
            for (int q = 1; q <= i1; ++q) {
                ++mm;
                store = store - cc[mm] * a[i][q + 1];
            }
            cc[i] = store / a[i][1];
            --i;
        }
    }
    free(a);
    free(z);
    return n;
}
