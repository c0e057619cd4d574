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

__global__ void __launch_bounds__(K1AF_THREADS) k1a_fitpack(K1Args a, FitArgs f)
{
    extern __shared__ __align__(16) double smf[];
    const int N = a.N, lane = threadIdx.x;
    double* TK = smf;  // [N + 7][32]; row l-1 holds FITPACK's t(l)
    const long long b = (long long)blockIdx.x * 32 + lane;
    const long long bb = (b < a.B) ? b : a.B - 1;  // padding lanes repeat the last candidate
    auto point = [&](int j, double& x, double& y) { control_point(a, bb, j, x, y); };  // track.py:87,:94
    // knots: np.cumsum of the chord lengths of the closed polygon (path.py:13-14); t(4 + j) = u_j
    {
        double acc = 0.0, x0, y0;
        point(0, x0, y0);
        TK[3 * 32 + lane] = 0.0;
        double xp = x0, yp = y0;
        for (int j = 0; j < N; ++j) {
            double xn = x0, yn = y0;
            if (j + 1 < N) point(j + 1, xn, yn);
            const double ex = xn - xp, ey = yn - yp;
            acc = acc + dsqrt<false>(ex * ex + ey * ey);
            TK[(j + 4) * 32 + lane] = acc;
            xp = xn; yp = yn;
        }
    }
    fit::Io io;
    io.t = TK + lane; io.st = 32;
    io.rows = f.rows + b; io.sr = (long)a.Bp;
    io.cx = f.cx + b; io.cy = f.cy + b; io.sc = (long)a.Bp;
    io.w1x = f.w1x + b; io.w1y = f.w1y + b; io.w2x = f.w2x + b; io.w2y = f.w2y + b; io.sw = (long)a.Bp;
    fit::solve(N, io, point);
    for (int l = 0; l < N + 7; ++l) f.t[(size_t)l * a.Bp + b] = TK[l * 32 + lane];
}

}  // namespace ltk
