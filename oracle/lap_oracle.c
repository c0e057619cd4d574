/* ORACLE (test infrastructure, not product code) -- level 2: plain-C restatement, batch capable.
 *
 * What it is: the reference's lap-time path restated in scalar C so that full-size populations
 * (65,536+ candidates) can be checked in seconds.  The spline is NOT FITPACK (SciPy is a third-party
 * dependency of the reference, `scipy==1.13.0` requirements.txt:14, absent from /root/reference): it is
 * the classical periodic interpolating cubic spline through the same knots, solved as a cyclic
 * tridiagonal system (SURVEY.md section 8(a) "equivalent closed form"), which is the same piecewise
 * cubic up to rounding.  Everything after the curvature follows the reference operation by operation.
 *
 * Parity status: PINNED by tests/test_oracle.py --
 *   - lto_sweeps() fed the golden (reference) curvature reproduces the golden v_acclim/v_declim/v/lap
 *     (bit-for-bit except where libm pow(x,2) != x*x, see LTO_USE_POW);
 *   - lto_eval() reproduces golden curvature to <=1e-11 (relative to max) and golden laps to <=1e-9.
 *
 * Reference lines followed (relative to /root/reference/src):
 *   control points      track.py:82-94            chord knots    path.py:11-14
 *   closure quirk       trajectory_bayesian_nonlinear.py:58-69 (splprep per=1 writes c[:, -1] = c[:, 0])
 *   sampling            trajectory_bayesian_nonlinear.py:71 / trajectory.py:45 (np.linspace)
 *   curvature           path.py:51-61             v_local        velocity.py:28-29
 *   forward sweep       velocity.py:31-53         backward sweep velocity.py:55-76
 *   engine / traction   vehicle.py:25-35, vehicleMX5.py:19-37
 *   lap time            trajectory_bayesian_nonlinear.py:51-54
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no fast-math).
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define LTO_MAX_MAP 16

typedef struct {
    int kind;      /* 0: tabulated engine map (vehicle.py); 1: MX5 polynomial (vehicleMX5.py) */
    int n_map;
    double mass;
    double mu_g;     /* friction_coef * GRAV, velocity.py:29 */
    double f_max;    /* vehicle.py:30  (mu*m)*g   | vehicleMX5.py:33 (lam*D)*(m*g) */
    double f_max_sq; /* f_max**2 as Python computes it */
    double map_v[LTO_MAX_MAP];
    double map_f[LTO_MAX_MAP];
    double e0;  /* MX5: (T*C_m) - Cr_0 */
    double cr2; /* MX5: Cr_2 */
} lto_vehicle;

/* 1: square with libm pow(x, 2.0) exactly like CPython / numpy scalars do (bit-identical to the
 * reference on the same libm); 0: x*x (what the CUDA kernels do). Differs in ~0.08 % of calls by 1 ulp. */
static int g_use_pow = 0;
/* 0: numpy pairwise order for the lap sum (what the reference does);
 * 1: the order of the CUDA K3 kernel (sequential from sample p-1 downwards, wrapping, sample p last);
 * 2: the order of the fused CUDA sweep kernel K23: with rows r = (q - p) mod n, h = (n-1)/2 rounded down,
 *    lap_f = rows n-h .. n-1 ascending, lap_b = rows h .. 1 descending,
 *    lap = ((lap_f + lap_b) [+ middle row h+1 when n-1 is odd]) + row 0.
 * 1 and 2 exist for bit-for-bit checks of the device kernels. */
static int g_sum_mode = 0;
void lto_set_sum_mode(int m) { g_sum_mode = m; }
/* 0: the classical cyclic-tridiagonal periodic spline below (what the default CUDA K1a solves);
 * 1: FITPACK's own arithmetic (oracle/fitpack_port.c: fpclos Givens QR + splder), i.e. the bits
 *    scipy.interpolate.splprep / splev produce for the reference (path.py:25, :51-54). */
static int g_spline_mode = 0;
void lto_set_spline_mode(int m) { g_spline_mode = m; }
int fpk_clocur(const double *u, const double *x, int m, int idim, double *t, double *c);
void fpk_splder(const double *t, int n, const double *c, int nu, const double *x, int m, double *y,
                double *wrk);
static volatile double g_two = 2.0; /* volatile: stops gcc folding pow(x, 2.0) into x*x */
void lto_set_use_pow(int on) { g_use_pow = on; }
static inline double sq(double x) { return g_use_pow ? pow(x, g_two) : x * x; }

/* np.interp for one abscissa (numpy/_core/src/multiarray/compiled_base.c arr_interp): clamp outside,
 * exact node value on a hit, otherwise slope*(x - xp[j]) + fp[j] with slope formed per segment. */
static double engine_table(const lto_vehicle *v, double x)
{
    int n = v->n_map;
    if (x <= v->map_v[0]) return v->map_f[0];
    if (x >= v->map_v[n - 1]) return v->map_f[n - 1];
    int j = 0;
    while (j + 2 < n && x >= v->map_v[j + 1]) ++j;
    if (x == v->map_v[j]) return v->map_f[j];
    double slope = (v->map_f[j + 1] - v->map_f[j]) / (v->map_v[j + 1] - v->map_v[j]);
    return slope * (x - v->map_v[j]) + v->map_f[j];
}

static double engine_force(const lto_vehicle *v, double vel)
{
    if (v->kind == 0) return engine_table(v, vel);
    return v->e0 - v->cr2 * sq(vel); /* vehicleMX5.py:21 */
}

static double traction(const lto_vehicle *v, double vel, double k)
{
    double f_lat = (v->kind == 0) ? (v->mass * sq(vel)) * k   /* vehicle.py:31 */
                                  : ((v->mass * vel) * vel) * k; /* vehicleMX5.py:34 */
    if (v->f_max <= f_lat) return 0.0;
    return sqrt(v->f_max_sq - sq(f_lat));
}

double lto_pairwise_sum(const double *a, long n);

/* np.linspace(0, L, ns)[i]: i*step with step = L/(ns-1); the last element is L itself. */
static inline double sample_at(long i, double step, double L, int ns)
{
    return (i == ns - 1) ? L : (double)i * step;
}

/* Three-pass profile + lap time for one candidate. k[n], n = ns-1 samples; closed path of period L.
 * Outputs (any may be NULL): v_local, v_acc, v_dec, v (each n).  Returns the lap time.
 * `work` must hold 3*n doubles. */
static double sweeps_one(const lto_vehicle *veh, const double *k, int ns, double L, int closed,
                         double *o_vlocal, double *o_vacc, double *o_vdec, double *o_v, double *work)
{
    int n = ns - 1;
    double step = L / (double)(ns - 1);
    double *vl = work, *va = work + n, *vd = work + 2 * n;
    int p = 0;
    for (int i = 0; i < n; ++i) {
        vl[i] = sqrt(veh->mu_g / k[i]); /* velocity.py:29 */
        if (vl[i] < vl[p]) p = i;       /* np.argmin: first minimum */
    }
    memcpy(va, vl, sizeof(double) * n);
    memcpy(vd, vl, sizeof(double) * n);
    /* forward, velocity.py:40-50 */
    for (int i = 0; i < n; ++i) {
        int q = p + i; if (q >= n) q -= n;
        int prev = q == 0 ? n - 1 : q - 1;
        int wrap = q == 0;
        if (wrap && !closed) continue;
        if (va[q] > va[prev]) {
            double tr = traction(veh, va[prev], k[prev]);
            double en = engine_force(veh, va[prev]);
            double force = en < tr ? en : tr; /* Python min(en, tr): tr unless en < tr; same value on ties */
            double accel = force / veh->mass;
            double ds = wrap ? L - sample_at(prev, step, L, ns)
                             : sample_at(q, step, L, ns) - sample_at(prev, step, L, ns);
            double vlim = sqrt(sq(va[prev]) + 2 * accel * ds);
            if (vlim < va[q]) va[q] = vlim;
        }
    }
    /* backward, velocity.py:64-73 */
    for (int i = 0; i < n; ++i) {
        int q = p - i; if (q < 0) q += n;
        int nxt = q == n - 1 ? 0 : q + 1;
        int wrap = q == n - 1;
        if (wrap && !closed) continue;
        if (vd[q] > vd[nxt]) {
            double tr = traction(veh, vd[nxt], k[nxt]);
            double decel = tr / veh->mass;
            double ds = wrap ? L - sample_at(q, step, L, ns)
                             : sample_at(nxt, step, L, ns) - sample_at(q, step, L, ns);
            double vlim = sqrt(sq(vd[nxt]) + 2 * decel * ds);
            if (vlim < vd[q]) vd[q] = vlim;
        }
    }
    /* v = minimum, lap = sum(diff(s)/v) with numpy's pairwise summation (blocks of 8 accumulators
     * below 128 elements, recursive halving above; numpy/_core/src/umath/loops_utils.h.src) */
    double *term = vl; /* reuse */
    for (int i = 0; i < n; ++i) {
        double v = va[i] < vd[i] ? va[i] : vd[i];
        if (o_vlocal) o_vlocal[i] = vl[i];
        if (o_v) o_v[i] = v;
        double ds = sample_at(i + 1, step, L, ns) - sample_at(i, step, L, ns);
        term[i] = ds / v;
    }
    if (o_vacc) memcpy(o_vacc, va, sizeof(double) * n);
    if (o_vdec) memcpy(o_vdec, vd, sizeof(double) * n);
    if (g_sum_mode == 2) {
        int rows = n - 1, h = rows / 2, has_mid = rows & 1;
        double lap_f = 0.0, lap_b = 0.0;
#define ROWQ(r) (((p + (r)) >= n) ? (p + (r)) - n : (p + (r)))
        for (int r = n - h; r <= n - 1; ++r) lap_f = lap_f + term[ROWQ(r)];
        for (int r = h; r >= 1; --r) lap_b = lap_b + term[ROWQ(r)];
        double lap = lap_f + lap_b;
        if (has_mid) lap = lap + term[ROWQ(h + 1)];
#undef ROWQ
        return lap + term[p];
    }
    if (g_sum_mode == 1) {
        double lap = 0.0;
        for (int i = 1; i < n; ++i) {
            int q = p - i; if (q < 0) q += n;
            lap = lap + term[q];
        }
        return lap + term[p];
    }
    return lto_pairwise_sum(term, n);
}

double lto_pairwise_sum(const double *a, long n)
{
    if (n < 8) {
        double r = 0.0;
        for (long i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        long i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    long n2 = n / 2;
    n2 -= n2 % 8;
    return lto_pairwise_sum(a, n2) + lto_pairwise_sum(a + n2, n - n2);
}

/* Periodic cubic spline through P_0..P_{N-1} (closure P_N = P_0), knots = cumulative chord length.
 * Produces per interval j: u[j], and for x and y: c1 (S' at the knot), c2 = M_j, c3 = (M_{j+1}-M_j)/h_j.
 * Cyclic tridiagonal  h_{j-1} M_{j-1} + 2(h_{j-1}+h_j) M_j + h_j M_{j+1} = 6 (d_j - d_{j-1}),
 * Sherman-Morrison on top of one Thomas factorisation shared by three right-hand sides.
 * Arrays: px,py [N]; u [N+1]; coef [6][N]; work [8*N]. Returns the period L (= u[N]). */
static double spline_build(const double *px, const double *py, int N, double *u, double *coef,
                           double *work)
{
    double *h = work, *dx = work + N, *dy = work + 2 * N, *inv = work + 3 * N, *cp = work + 4 * N;
    double *rx = work + 5 * N, *ry = work + 6 * N, *rz = work + 7 * N;
    u[0] = 0.0;
    for (int j = 0; j < N; ++j) {
        int jn = j + 1 == N ? 0 : j + 1;
        double ex = px[jn] - px[j], ey = py[jn] - py[j];
        dx[j] = ex;
        dy[j] = ey;
        u[j + 1] = u[j] + sqrt(ex * ex + ey * ey); /* np.linalg.norm(axis=0) then np.cumsum (path.py:13) */
    }
    for (int j = 0; j < N; ++j) {
        /* the spline only ever sees the knots (splprep gets u=dists), so the interval widths are knot
         * differences, not the chord lengths they were accumulated from */
        h[j] = u[j + 1] - u[j];
        dx[j] = dx[j] / h[j];
        dy[j] = dy[j] / h[j];
    }
    /* rows: a_j = h_{j-1}, b_j = 2(h_{j-1}+h_j), c_j = h_j; corners a_0 (col N-1) and c_{N-1} (col 0) */
    double hl = h[N - 1];
    double b0 = 2.0 * (hl + h[0]);
    double gamma = -b0;
    /* modified diagonal: b0' = b0 - gamma, b_{N-1}' = b_{N-1} - c_{N-1} a_0 / gamma */
    double bfirst = b0 - gamma;
    double blast = 2.0 * (h[N - 2] + hl) - hl * hl / gamma;
    /* forward elimination */
    inv[0] = 1.0 / bfirst;
    cp[0] = h[0] * inv[0];
    rx[0] = 6.0 * (dx[0] - dx[N - 1]) * inv[0];
    ry[0] = 6.0 * (dy[0] - dy[N - 1]) * inv[0];
    rz[0] = gamma * inv[0];
    for (int j = 1; j < N; ++j) {
        double a = h[j - 1];
        double b = (j == N - 1) ? blast : 2.0 * (h[j - 1] + h[j]);
        double den = b - a * cp[j - 1];
        inv[j] = 1.0 / den;
        cp[j] = h[j] * inv[j];
        double fx = 6.0 * (dx[j] - dx[j - 1]);
        double fy = 6.0 * (dy[j] - dy[j - 1]);
        double fz = (j == N - 1) ? hl : 0.0; /* u vector: (gamma, 0, ..., 0, c_{N-1}) */
        rx[j] = (fx - a * rx[j - 1]) * inv[j];
        ry[j] = (fy - a * ry[j - 1]) * inv[j];
        rz[j] = (fz - a * rz[j - 1]) * inv[j];
    }
    /* back substitution (in place) */
    for (int j = N - 2; j >= 0; --j) {
        rx[j] -= cp[j] * rx[j + 1];
        ry[j] -= cp[j] * ry[j + 1];
        rz[j] -= cp[j] * rz[j + 1];
    }
    /* Sherman-Morrison: v = (1, 0, ..., 0, a_0/gamma) */
    double vN = hl / gamma;
    double denom = 1.0 + (rz[0] + vN * rz[N - 1]);
    double fxs = (rx[0] + vN * rx[N - 1]) / denom;
    double fys = (ry[0] + vN * ry[N - 1]) / denom;
    for (int j = 0; j < N; ++j) {
        rx[j] -= fxs * rz[j]; /* M_x */
        ry[j] -= fys * rz[j]; /* M_y */
    }
    for (int j = 0; j < N; ++j) {
        int jn = j + 1 == N ? 0 : j + 1;
        coef[0 * N + j] = dx[j] - h[j] * (2.0 * rx[j] + rx[jn]) / 6.0;
        coef[1 * N + j] = rx[j];
        coef[2 * N + j] = (rx[jn] - rx[j]) / h[j];
        coef[3 * N + j] = dy[j] - h[j] * (2.0 * ry[j] + ry[jn]) / 6.0;
        coef[4 * N + j] = ry[j];
        coef[5 * N + j] = (ry[jn] - ry[j]) / h[j];
    }
    return u[N];
}

/* curvature at the ns-1 samples s_i = i*L/(ns-1), path.py:58/61 */
static void spline_curvature(const double *u, const double *coef, int N, int ns, double L, double *k,
                             double *o_dx, double *o_dy, double *o_ddx, double *o_ddy)
{
    double step = L / (double)(ns - 1);
    int j = 0;
    for (int i = 0; i < ns - 1; ++i) {
        double s = (double)i * step;
        while (j + 1 < N && s >= u[j + 1]) ++j;
        double t = s - u[j];
        double c1x = coef[0 * N + j], c2x = coef[1 * N + j], c3x = coef[2 * N + j];
        double c1y = coef[3 * N + j], c2y = coef[4 * N + j], c3y = coef[5 * N + j];
        /* Horner form with explicit fused multiply-adds: the same sequence the CUDA K1 kernel issues */
        double ddx = fma(c3x, t, c2x), ddy = fma(c3y, t, c2y);
        double dx = fma(t, fma(0.5 * c3x, t, c2x), c1x);
        double dy = fma(t, fma(0.5 * c3y, t, c2y), c1y);
        double cross = fma(dx, ddy, -(dy * ddx));
        double n2 = fma(dx, dx, dy * dy);
        k[i] = fabs(cross / (n2 * sqrt(n2)));
        if (o_dx) { o_dx[i] = dx; o_dy[i] = dy; o_ddx[i] = ddx; o_ddy[i] = ddy; }
    }
}

/* x**1.5 rounded to nearest: sqrt in double-double, product in double-double.  numpy evaluates
 * `(dx**2 + dy**2) ** (3/2)` (path.py:58) with libm pow -- or, on AVX512 hosts, with its vendored SVML pow,
 * which is off by one ulp in ~5 % of the arguments and cannot be restated; the correctly rounded value is
 * what both approximate and what the CUDA kernel computes with the same four operations. */
double lto_pow15(double x)
{
    double s = sqrt(x);
    double r = fma(-s, s, x);  /* x - s*s exactly */
    double e = r / (s + s);    /* sqrt(x) = s + e */
    double p = x * s;
    double pe = fma(x, s, -p); /* x*s = p + pe exactly */
    return p + fma(x, e, pe);
}

/* FITPACK mode: knots = cumulative chord lengths of the closed polygon (path.py:11-14), periodic
 * interpolating B-spline (path.py:25), derivative values at the samples (path.py:51-54), curvature in
 * numpy's operation order (path.py:58, :61).  work: (N+1) + 2(N+1) + (N+7) + 2(N+7) + (N+7) + 4n doubles. */
static double fitpack_curvature(const double *px, const double *py, int N, int ns, double *k, double *work,
                                double *o_dx, double *o_dy, double *o_ddx, double *o_ddy)
{
    int m = N + 1, nk = m + 6, n = ns - 1;
    double *u = work, *xy = u + m, *t = xy + 2 * m, *c = t + nk, *wrk = c + 2 * nk, *s = wrk + nk;
    double *d1x = s + n, *d1y = d1x + n, *d2x = d1y + n;
    u[0] = 0.0;
    for (int j = 0; j < N; ++j) {
        int jn = j + 1 == N ? 0 : j + 1;
        double ex = px[jn] - px[j], ey = py[jn] - py[j];
        u[j + 1] = u[j] + sqrt(ex * ex + ey * ey);
        xy[2 * j] = px[j];
        xy[2 * j + 1] = py[j];
    }
    xy[2 * N] = px[0];
    xy[2 * N + 1] = py[0];
    fpk_clocur(u, xy, m, 2, t, c);
    double L = u[N], step = L / (double)(ns - 1);
    for (int i = 0; i < n; ++i) s[i] = (double)i * step;
    double *d2y = k; /* reuse the output as scratch for the last derivative */
    fpk_splder(t, nk, c, 1, s, n, d1x, wrk);
    fpk_splder(t, nk, c + nk, 1, s, n, d1y, wrk);
    fpk_splder(t, nk, c, 2, s, n, d2x, wrk);
    fpk_splder(t, nk, c + nk, 2, s, n, d2y, wrk);
    for (int i = 0; i < n; ++i) {
        double dx = d1x[i], dy = d1y[i], ddx = d2x[i], ddy = d2y[i];
        if (o_dx) { o_dx[i] = dx; o_dy[i] = dy; o_ddx[i] = ddx; o_ddy[i] = ddy; }
        double cross = dx * ddy - dy * ddx;
        double n2 = dx * dx + dy * dy;
        k[i] = fabs(cross / lto_pow15(n2));
    }
    return L;
}

/* ---- exported entry points (ctypes) ------------------------------------------------------------ */

/* Sweeps only: one candidate, curvature supplied (e.g. the reference's own FITPACK curvature). */
double lto_sweeps(const lto_vehicle *veh, const double *k, int ns, double L, int closed,
                  double *o_vlocal, double *o_vacc, double *o_vdec, double *o_v)
{
    double *work = (double *)malloc(sizeof(double) * 3 * (size_t)(ns - 1));
    double lap = sweeps_one(veh, k, ns, L, closed, o_vlocal, o_vacc, o_vdec, o_v, work);
    free(work);
    return lap;
}

/* Full path for a batch: alphas[B][N] -> lap[B].  left_xy/diff_xy are [2][N] (x row then y row) for
 * the N unique control points (closure implied).  Optional per-candidate dumps for candidate
 * `dump_index` (pass -1 for none): k, v_local, v_acc, v_dec, v (each ns-1) and length[1].
 * `n_threads` POSIX threads split the candidates.  Returns 0, or -1 on bad arguments. */
typedef struct {
    const double *left_xy, *diff_xy, *alphas;
    const lto_vehicle *veh;
    int N, ns;
    long b0, b1, dump_index;
    double *lap, *o_k, *o_vlocal, *o_vacc, *o_vdec, *o_v, *o_length;
} lto_job;

static void *eval_range(void *arg)
{
    lto_job *jb = (lto_job *)arg;
    int N = jb->N, ns = jb->ns, n = ns - 1;
    double *px = (double *)malloc(sizeof(double) * (size_t)(2 * N + (N + 1) + 6 * N + 8 * N + 4 * n));
    double *fw = g_spline_mode ? (double *)malloc(sizeof(double) * (size_t)(8 * (N + 7) + 4 * n)) : 0;
    double *py = px + N, *u = py + N, *coef = u + N + 1, *work = coef + 6 * N;
    double *k = work + 8 * N, *sw = k + n;
    for (long b = jb->b0; b < jb->b1; ++b) {
        const double *a = jb->alphas + b * N;
        for (int j = 0; j < N; ++j) { /* track.py:87 / :94 */
            px[j] = jb->left_xy[j] + a[j] * jb->diff_xy[j];
            py[j] = jb->left_xy[N + j] + a[j] * jb->diff_xy[N + j];
        }
        double L;
        if (g_spline_mode) {
            L = fitpack_curvature(px, py, N, ns, k, fw, 0, 0, 0, 0);
        } else {
            L = spline_build(px, py, N, u, coef, work);
            spline_curvature(u, coef, N, ns, L, k, 0, 0, 0, 0);
        }
        int d = (b == jb->dump_index);
        jb->lap[b] = sweeps_one(jb->veh, k, ns, L, 1, d ? jb->o_vlocal : 0, d ? jb->o_vacc : 0,
                                d ? jb->o_vdec : 0, d ? jb->o_v : 0, sw);
        if (d) {
            if (jb->o_k) memcpy(jb->o_k, k, sizeof(double) * n);
            if (jb->o_length) *jb->o_length = L;
        }
    }
    free(px);
    free(fw);
    return 0;
}

int lto_eval(const double *left_xy, const double *diff_xy, int N, const lto_vehicle *veh, int ns,
             const double *alphas, long B, double *lap, int n_threads, long dump_index, double *o_k,
             double *o_vlocal, double *o_vacc, double *o_vdec, double *o_v, double *o_length)
{
    if (N < 3 || ns < 3 || B < 0) return -1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if ((long)n_threads > B) n_threads = B > 0 ? (int)B : 1;
    lto_job jobs[256];
    pthread_t tid[256];
    for (int t = 0; t < n_threads; ++t) {
        lto_job j = {left_xy, diff_xy, alphas, veh, N, ns, B * t / n_threads, B * (t + 1) / n_threads,
                     dump_index, lap, o_k, o_vlocal, o_vacc, o_vdec, o_v, o_length};
        jobs[t] = j;
    }
    for (int t = 1; t < n_threads; ++t) pthread_create(&tid[t], 0, eval_range, &jobs[t]);
    eval_range(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) pthread_join(tid[t], 0);
    return 0;
}

/* Spline only: derivatives and curvature of one candidate's path (for Path-level tests). */
int lto_path(const double *px, const double *py, int N, int ns, double *o_length, double *o_k,
             double *o_dx, double *o_dy, double *o_ddx, double *o_ddy)
{
    if (N < 3 || ns < 3) return -1;
    double *u = (double *)malloc(sizeof(double) * (size_t)((N + 1) + 6 * N + 8 * N + 8 * (N + 7) + 4 * ns));
    double *coef = u + N + 1, *work = coef + 6 * N;
    double L;
    if (g_spline_mode) {
        L = fitpack_curvature(px, py, N, ns, o_k, work + 8 * N, o_dx, o_dy, o_ddx, o_ddy);
    } else {
        L = spline_build(px, py, N, u, coef, work);
        spline_curvature(u, coef, N, ns, L, o_k, o_dx, o_dy, o_ddx, o_ddy);
    }
    if (o_length) *o_length = L;
    free(u);
    return 0;
}
