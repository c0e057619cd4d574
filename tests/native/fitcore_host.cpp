// TEST INFRASTRUCTURE: host build of lap_time_optimization_b200/csrc/ltk_fitpack_core.cuh.
//
// The FITPACK-mode arithmetic of the CUDA kernels (fit::solve, fit::curvature_at) is written as
// __host__ __device__ code; this file compiles the same header with g++ (-ffp-contract=off, the host
// counterpart of nvcc -fmad=false) so that tests/test_host.py can compare it with SciPy and with
// oracle/fitpack_port.c bit for bit without a GPU.  Nothing in the product links this.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../lap_time_optimization_b200/csrc/ltk_fitpack_core.cuh"

using namespace ltk::fit;

extern "C" {

// px, py [N] unique control points; u [N+1] chord-length knots.  Outputs: t [N+7], cx, cy [N+3],
// w1x, w1y [N+2], w2x, w2y [N+1].
int fitcore_solve(int N, const double* px, const double* py, const double* u, double* t, double* cx, double* cy,
                  double* w1x, double* w1y, double* w2x, double* w2y)
{
    if (N < 5) return -1;
    std::vector<double> rows((size_t)7 * N, 0.0);
    for (int j = 0; j <= N; ++j) t[j + 3] = u[j];
    Io io;
    io.t = t; io.st = 1;
    io.rows = rows.data(); io.sr = 1;
    io.cx = cx; io.cy = cy; io.sc = 1;
    io.w1x = w1x; io.w1y = w1y; io.w2x = w2x; io.w2y = w2y; io.sw = 1;
    struct Points {
        struct Raw { double x, y; };
        const double *px, *py;
        void fetch(int j, Raw& r) const { r.x = px[j]; r.y = py[j]; }
        void finish(const Raw& r, double& x, double& y) const { x = r.x; y = r.y; }
    } pts{px, py};
    solve(N, io, pts);
    return 0;
}

// curvature and derivatives at x[0..m-1] (ascending, inside [t(4), t(N+4)) ) from the outputs of fitcore_solve
int fitcore_curvature(int N, const double* t, const double* w1x, const double* w1y, const double* w2x,
                      const double* w2y, const double* x, int m, double* k, double* dx, double* dy, double* ddx,
                      double* ddy)
{
    int j = 0;
    for (int i = 0; i < m; ++i) {
        while (j + 1 < N && x[i] >= t[j + 4]) ++j;
        FitInterval v;
        v.tm1 = t[j + 2]; v.t0 = t[j + 3]; v.tp1 = t[j + 4]; v.tp2 = t[j + 5];
        v.inv01 = 1.0 / (v.tp1 - v.t0);
        v.d1 = v.tp1 - v.tm1; v.r1 = 1.0 / v.d1;
        v.d2 = v.tp2 - v.t0;  v.r2 = 1.0 / v.d2;
        for (int q = 0; q < 3; ++q) { v.w1x[q] = w1x[j + q]; v.w1y[q] = w1y[j + q]; }
        for (int q = 0; q < 2; ++q) { v.w2x[q] = w2x[j + q]; v.w2y[q] = w2y[j + q]; }
        v.pad = 0.0;
        k[i] = curvature_at(v, x[i], dx[i], dy[i], ddx[i], ddy[i]);
    }
    return 0;
}

// general evaluation: derivative nu (0..2) of the spline (t [n], cx, cy [n - 4]) at x [m]
int fitcore_splev(int n, const double* t, const double* cx, const double* cy, int nu, const double* x, int m,
                  double* ox, double* oy)
{
    std::vector<double> w1x(n), w1y(n), w2x(n), w2y(n);
    der_coeffs(t, n, cx, w1x.data(), w2x.data());
    der_coeffs(t, n, cy, w1y.data(), w2y.data());
    const double* kx = nu == 0 ? cx : nu == 1 ? w1x.data() : w2x.data();
    const double* ky = nu == 0 ? cy : nu == 1 ? w1y.data() : w2y.data();
    for (int i = 0; i < m; ++i) {
        const int l = find_interval(t, n, x[i]);
        ox[i] = splev_at(t, kx, nu, x[i], l);
        oy[i] = splev_at(t, ky, nu, x[i], l);
    }
    return 0;
}

// open (not-a-knot) interpolating spline: u, px, py [m] -> t [m + 4], cx, cy [m]
int fitcore_solve_open(int m, const double* u, const double* px, const double* py, double* t, double* cx, double* cy)
{
    if (m < 4) return -1;
    std::vector<double> a((size_t)4 * m), z((size_t)2 * m);
    solve_open(m, u, px, py, t, a.data(), z.data(), cx, cy);
    return 0;
}

}  // extern "C"
