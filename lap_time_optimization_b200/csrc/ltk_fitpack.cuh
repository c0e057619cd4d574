// ltk_fitpack.cuh -- K1a-F: alphas -> control points -> chord-length knots -> FITPACK's periodic interpolating
// B-spline (fpclos, s = 0) -> derivative coefficients (splder), one thread per candidate.
//
//   reference: Track.control_points* (track.py:82-94), cumulative_distances (path.py:11-14),
//              splprep(controls, u=dists, k=3, s=0, per=1) (path.py:25), splev(der=1|2) set-up (path.py:51-54)
//
// The arithmetic is ltk_fitpack_core.cuh (fit::solve).  One warp per CTA, lane = candidate.  Only the knot vector
// lives in shared memory (N + 7 rows of 32 doubles, lane-minor: 12.8 KB at N = 43, so that all 2,048 warps of a
// 65,536-candidate population are resident at once -- the solve is a serial chain of ~80 Givens rotations per
// candidate and nothing but resident warps hides its latency); control points are recomputed from the alpha row
// where they are needed (the row stays in L1), the triangular factor goes to a candidate-minor global scratch
// (written once, read once by the back substitution; coalesced 256-byte rows), and the knots, B-spline
// coefficients and derivative coefficients are left candidate-minor for K1b (k1b_samples<..., FIT = true>).
#pragma once

namespace ltk {

constexpr int K1AF_THREADS = 32;
__host__ __device__ inline size_t k1af_smem_bytes(int N) { return (size_t)(N + 7) * 32 * sizeof(double); }
// doubles per candidate of the FITPACK-mode hand-off / scratch region of the workspace
__host__ __device__ inline size_t fit_region_doubles(int N)
{
    return (size_t)(N + 7) + 7 * (size_t)N + 2 * (size_t)(N + 3) + 2 * (size_t)(N + 2) + 2 * (size_t)(N + 1);
}

// control point j of candidate b in two steps (fit::solve): loads, then track.py:87,:94 / the caller's points
struct K1Points {
    struct Raw { double al, lx, ly, dx, dy; };
    const K1Args& a;
    long long b;
    __device__ __forceinline__ void fetch(int j, Raw& r) const
    {
        if (a.mode == 0) {
            r.al = a.alphas[b * a.N + j];
            r.lx = a.left[j]; r.dx = a.diff[j];
            r.ly = a.left[a.N + j]; r.dy = a.diff[a.N + j];
        } else {
            r.lx = a.xy[(b * 2 + 0) * a.m + j];
            r.ly = a.xy[(b * 2 + 1) * a.m + j];
            r.al = r.dx = r.dy = 0.0;
        }
    }
    __device__ __forceinline__ void finish(const Raw& r, double& x, double& y) const
    {
        if (a.mode == 0) { x = r.lx + r.al * r.dx; y = r.ly + r.al * r.dy; }
        else { x = r.lx; y = r.ly; }
    }
};
struct SmemPoints {  // the one-path facade kernel: points already in shared memory
    struct Raw { double x, y; };
    const double *px, *py;
    __device__ __forceinline__ void fetch(int j, Raw& r) const { r.x = px[j]; r.y = py[j]; }
    __device__ __forceinline__ void finish(const Raw& r, double& x, double& y) const { x = r.x; y = r.y; }
};

__global__ void __launch_bounds__(K1AF_THREADS) k1a_fitpack(K1Args a, FitArgs f)
{
    extern __shared__ __align__(16) double smf[];
    const int N = a.N, lane = threadIdx.x;
    double* TK = smf;  // [N + 7][32]; row l-1 holds FITPACK's t(l)
    const long long b = (long long)blockIdx.x * 32 + lane;
    const long long bb = (b < a.B) ? b : a.B - 1;  // padding lanes repeat the last candidate
    const K1Points pts{a, bb};
    // knots: np.cumsum of the chord lengths of the closed polygon (path.py:13-14); t(4 + j) = u_j.
    // The loads of point j+2 are in flight while the chord j -> j+1 is measured.
    {
        double acc = 0.0, x0, y0, xp, yp, xn, yn;
        K1Points::Raw r0, r1, r2;
        pts.fetch(0, r0);
        pts.fetch(1, r1);
        pts.finish(r0, x0, y0);
        TK[3 * 32 + lane] = 0.0;
        xp = x0; yp = y0;
        for (int j = 0; j < N; ++j) {
            if (j + 2 < N) pts.fetch(j + 2, r2);
            if (j + 1 < N) pts.finish(r1, xn, yn);
            else { xn = x0; yn = y0; }
            const double ex = xn - xp, ey = yn - yp;
            acc = acc + dsqrt<false>(ex * ex + ey * ey);
            TK[(j + 4) * 32 + lane] = acc;
            xp = xn; yp = yn;
            r1 = r2;
        }
    }
    fit::Io io;
    io.t = TK + lane; io.st = 32;
    io.rows = f.rows + b; io.sr = (long)a.Bp;
    io.cx = f.cx + b; io.cy = f.cy + b; io.sc = (long)a.Bp;
    const int hrows = 5 * N + 13;
    double* hand = f.hand + hand_index(b, hrows, 0);
    io.w1x = hand + (size_t)(N + 7) * HAND_G; io.w1y = io.w1x + (size_t)(N + 2) * HAND_G;
    io.w2x = io.w1y + (size_t)(N + 2) * HAND_G; io.w2y = io.w2x + (size_t)(N + 1) * HAND_G; io.sw = HAND_G;
    fit::solve(N, io, pts);
    for (int l = 0; l < N + 7; ++l) hand[(size_t)l * HAND_G] = TK[l * 32 + lane];
}

// ------------------------------------------------------------------------------------------------
// Path facade in FITPACK arithmetic (path.py:17-77 for ONE path): splprep(controls, u=dists, k=3, s=0, per=closed)
// by thread 0 -- the periodic fpclos solve (fit::solve) or the open not-a-knot one (fit::solve_open) -- then
// splev(der = 0 | 1 | 2) at the caller's parameters by all threads, curvature in numpy's operation order.
// Everything lives in shared memory: t | cx | cy | w1x | w1y | w2x | w2y | scratch.
// ------------------------------------------------------------------------------------------------
struct PathFitArgs {
    const double* xy;     // [2][m]
    const double* knots;  // [m]
    int m, closed;
    const double* u;
    long long n;
    double *x, *y, *dx, *dy, *ddx, *ddy, *k, *gamma2;
    double* t_out;  // [m + 6] closed | [m + 4] open, or nullptr
    double* c_out;  // [2][m + 2] closed | [2][m] open, or nullptr
};

__host__ __device__ inline size_t path_fit_smem_doubles(int m, int closed)
{
    const size_t n = closed ? m + 6 : m + 4;
    return n + 6 * n + (closed ? 7 * (size_t)(m - 1) + 2 * (size_t)m : 6 * (size_t)m);
}

__global__ void __launch_bounds__(256) path_fit_kernel(PathFitArgs a)
{
    extern __shared__ __align__(16) double smp[];
    __shared__ double red[256];
    const int m = a.m, tid = threadIdx.x;
    const int n = a.closed ? m + 6 : m + 4, nc = n - 4;
    double* T = smp;
    double* CX = T + n;
    double* CY = CX + n;
    double* W1X = CY + n;
    double* W1Y = W1X + n;
    double* W2X = W1Y + n;
    double* W2Y = W2X + n;
    double* SCR = W2Y + n;
    if (tid == 0) {
        if (a.closed) {
            const int N = m - 1;
            double* PX = SCR + 7 * N;
            double* PY = PX + m;
            for (int j = 0; j < N; ++j) { PX[j] = a.xy[j]; PY[j] = a.xy[m + j]; }
            for (int j = 0; j <= N; ++j) T[j + 3] = a.knots[j];
            fit::Io io;
            io.t = T; io.st = 1;
            io.rows = SCR; io.sr = 1;
            io.cx = CX; io.cy = CY; io.sc = 1;
            io.w1x = W1X; io.w1y = W1Y; io.w2x = W2X; io.w2y = W2Y; io.sw = 1;
            fit::solve(N, io, SmemPoints{PX, PY});
        } else {
            double* A = SCR;         // [4 m]
            double* Z = A + 4 * m;   // [2 m]
            fit::solve_open(m, a.knots, a.xy, a.xy + m, T, A, Z, CX, CY);
            fit::der_coeffs(T, n, CX, W1X, W2X);
            fit::der_coeffs(T, n, CY, W1Y, W2Y);
        }
    }
    __syncthreads();
    if (a.t_out) for (int i = tid; i < n; i += blockDim.x) a.t_out[i] = T[i];
    if (a.c_out) for (int i = tid; i < nc; i += blockDim.x) { a.c_out[i] = CX[i]; a.c_out[nc + i] = CY[i]; }
    double g2 = 0.0;
    for (long long e = tid; e < a.n; e += blockDim.x) {
        const double s = a.u[e];
        const int l = fit::find_interval(T, n, s);
        const double dx = fit::splev_at(T, W1X, 1, s, l), dy = fit::splev_at(T, W1Y, 1, s, l);
        const double ddx = fit::splev_at(T, W2X, 2, s, l), ddy = fit::splev_at(T, W2Y, 2, s, l);
        const double cross = dx * ddy - dy * ddx;       // path.py:58 in numpy's operation order
        const double n2 = dx * dx + dy * dy;
        const double k = cross / fit::pow15(n2);
        if (a.x) a.x[e] = fit::splev_at(T, CX, 0, s, l);
        if (a.y) a.y[e] = fit::splev_at(T, CY, 0, s, l);
        if (a.dx) a.dx[e] = dx;
        if (a.dy) a.dy[e] = dy;
        if (a.ddx) a.ddx[e] = ddx;
        if (a.ddy) a.ddy[e] = ddy;
        if (a.k) a.k[e] = k;
        g2 += k * k;
    }
    red[tid] = g2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0 && a.gamma2) a.gamma2[0] = red[0];
}

}  // namespace ltk
