// ltk_fitpack.cuh -- K1a-F: alphas -> control points -> chord-length knots -> FITPACK's periodic interpolating
// B-spline (fpclos, s = 0) -> derivative coefficients (splder), one thread per candidate.
//
//   reference: Track.control_points* (track.py:82-94), cumulative_distances (path.py:11-14),
//              splprep(controls, u=dists, k=3, s=0, per=1) (path.py:25), splev(der=1|2) set-up (path.py:51-54)
//
// The arithmetic is ltk_fitpack_core.cuh (fit::solve); this kernel stages one warp's 32 candidates in shared
// memory, lane-minor (control points and the knot vector: 3N + 7 rows of 32 doubles, conflict-free), runs the
// solve with the triangular factor in a candidate-minor global scratch (written once, read once by the back
// substitution; coalesced 256-byte rows) and leaves knots and derivative coefficients candidate-minor for
// K1b (k1b_samples<..., FIT = true>).
#pragma once

namespace ltk {

constexpr int K1AF_THREADS = 32;
__host__ __device__ inline size_t k1af_smem_bytes(int N) { return (size_t)(3 * N + 7) * 32 * sizeof(double); }
// doubles per candidate of the FITPACK-mode hand-off / scratch region of the workspace
__host__ __device__ inline size_t fit_region_doubles(int N)
{
    return (size_t)(N + 7) + 7 * (size_t)N + 2 * (size_t)(N + 3) + 2 * (size_t)(N + 2) + 2 * (size_t)(N + 1);
}

__global__ void __launch_bounds__(K1AF_THREADS) k1a_fitpack(K1Args a, FitArgs f)
{
    extern __shared__ __align__(16) double smf[];
    const int N = a.N, lane = threadIdx.x;
    double* PX = smf;                // [N][32]
    double* PY = PX + N * 32;
    double* TK = PY + N * 32;        // [N + 7][32]; row l-1 holds FITPACK's t(l)
    const long long b0 = (long long)blockIdx.x * 32;
    // control points (track.py:87,:94): consecutive lanes along a candidate's alpha row
    for (int idx = lane; idx < N * 32; idx += 32) {
        const int g = idx / N, j = idx - g * N;
        long long b = b0 + g;
        b = (b < a.B) ? b : a.B - 1;  // padding lanes repeat the last candidate
        double x, y;
        control_point(a, b, j, x, y);
        PX[j * 32 + g] = x;
        PY[j * 32 + g] = y;
    }
    __syncwarp();
    // knots: np.cumsum of the chord lengths of the closed polygon (path.py:13-14); t(4 + j) = u_j
    {
        double acc = 0.0;
        TK[3 * 32 + lane] = 0.0;
        double xp = PX[lane], yp = PY[lane];
        const double x0 = xp, y0 = yp;
        for (int j = 0; j < N; ++j) {
            const double xn = (j + 1 == N) ? x0 : PX[(j + 1) * 32 + lane];
            const double yn = (j + 1 == N) ? y0 : PY[(j + 1) * 32 + lane];
            const double ex = xn - xp, ey = yn - yp;
            acc = acc + dsqrt<false>(ex * ex + ey * ey);
            TK[(j + 4) * 32 + lane] = acc;
            xp = xn; yp = yn;
        }
    }
    const long long b = b0 + lane;
    fit::Io io;
    io.px = PX + lane; io.py = PY + lane; io.sp = 32;
    io.t = TK + lane; io.st = 32;
    io.rows = f.rows + b; io.sr = (long)a.Bp;
    io.cx = f.cx ? f.cx + b : nullptr; io.cy = f.cy ? f.cy + b : nullptr; io.sc = (long)a.Bp;
    io.w1x = f.w1x + b; io.w1y = f.w1y + b; io.w2x = f.w2x + b; io.w2y = f.w2y + b; io.sw = (long)a.Bp;
    fit::solve(N, io);
    for (int l = 0; l < N + 7; ++l) f.t[(size_t)l * a.Bp + b] = TK[l * 32 + lane];
}

}  // namespace ltk
