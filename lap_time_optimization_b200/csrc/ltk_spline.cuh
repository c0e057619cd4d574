// ltk_spline.cuh -- K1: alphas -> periodic cubic spline -> curvature at the samples.
//
//   reference: Track.control_points* (track.py:82-94), Path.__init__ -> splprep(k=3, s=0, per=1)
//   (path.py:11-26), np.linspace sampling (tbn.py:71), Path.curvature (path.py:36-61), and the arg-min of
//   v_local that both sweeps start from (velocity.py:34,58).
//
// Two kernels:
//
//   K1a k1a_solve     control points -> chord-length knots -> cyclic tridiagonal solve (Thomas +
//                     Sherman-Morrison) for the second derivatives M_x, M_y.  A CTA is three warps over
//                     the same 32 candidates: lane = candidate, warp = right-hand side (x | y | corner
//                     vector).  The three warps repeat the pivot recurrence, so the serial part needs no
//                     communication; scratch is five [N][32] arrays in shared memory (55 KB at N = 43,
//                     four CTAs per SM).  Output: knots [N+1][Bp], M_x, M_y [N][Bp], candidate-minor.
//   K1b k1b_samples   a CTA owns G = 4 candidates (one 32-byte sector of every 128-byte tile row):
//                     per-interval cubic coefficients and the first sample index of every interval
//                     (thread per interval), curvature (each thread walks a contiguous chunk of samples;
//                     the interval switch is an integer test on the sample index), the tile in shared
//                     memory, first maximum, write-out rotated so that row 0 of every candidate is its
//                     own slowest sample.
//
// (Both were tried as ONE kernel with the solve on 3G lanes of each CTA: the ~15,000-cycle serial part
// then idles the other warps of its CTA and the kernel took 0.51 ms against 0.10 + 0.30 ms for the
// previous pair; see profiles/.)
//
// Arithmetic is the fixed sequence oracle/lap_oracle.c mirrors: IEEE + - * / sqrt, fma only where written.
#pragma once

namespace ltk {

#ifndef LTK_K1B_TRIM
#define LTK_K1B_TRIM 31  // bit set of the trims below (A/B): 1 packed record, 2 guessed chunk start, 4 write-out, 8 step,
#endif                   // 16 record phase (first-sample guess by a multiplication, / 6 as a multiplication + one correction)
// Interval record of k1b_samples (tridiagonal mode), 64 bytes = four 16-byte shared loads:
// S'(t) = c1 + t (c2 + t h),  S''(t) = c2 + c3 t;  h = c3 / 2 is formed on load (exact).
struct __align__(16) IntervalP {
    double u, c1x, c1y, c2x, c2y, c3x, c3y;
    int inext, pad;
    double pad2[2];  // 80-byte stride: the G records of a row then fall in distinct banks (64 bytes: two-way conflicts)
};
static_assert(sizeof(IntervalP) == 80, "IntervalP must be 80 bytes");
constexpr bool K1B_TRIM = (LTK_K1B_TRIM & 1) != 0;      // packed record
constexpr bool K1B_GUESS = (LTK_K1B_TRIM & 2) != 0;     // chunk start interval from a proportional guess
constexpr bool K1B_WLIN = (LTK_K1B_TRIM & 4) != 0;      // write-out with linear cursors
constexpr bool K1B_STEP = (LTK_K1B_TRIM & 8) != 0;      // sampling step once per candidate
constexpr bool K1B_CREC = (LTK_K1B_TRIM & 16) != 0;     // cheaper record phase
#ifndef LTK_K1B_PREFETCH
#define LTK_K1B_PREFETCH 0  // > 0: every CTA bulk-prefetches into L2 the inputs of the CTA that many blocks ahead
#endif
#ifndef LTK_K1B_TIES
#define LTK_K1B_TIES 1  // 0 (A/B only): rotation = first maximum of the curvature, plateaus not re-examined
#endif
constexpr bool K1B_TIES = LTK_K1B_TIES != 0;

__host__ __device__ inline int k1_chunk(int n, int cpt)
{
    int c = (n + cpt - 1) / cpt;
    return c | 1;  // odd: the per-thread rows of a warp then fall in distinct shared-memory banks
}

// doubles of scratch that alias the curvature tile (see the kernel for the carve-up)
__host__ __device__ inline size_t k1f_scratch_doubles(int G, int N, bool fitpack = false)
{
    return (size_t)(fitpack ? 5 * N + 13 : 5 * N + 1) * G;
}

__host__ __device__ inline size_t k1f_smem_bytes(int G, int threads, int N, int ns, bool fitpack = false,
                                                 bool staged = true)
{
    size_t tile = staged ? (size_t)(ns - 1) * G : 0, scr = k1f_scratch_doubles(G, N, fitpack);
    size_t bytes = (size_t)N * G * (fitpack ? sizeof(fit::FitInterval) : K1B_TRIM ? sizeof(IntervalP) : sizeof(Interval));  // interval records [N][G]
    bytes += (tile > scr ? tile : scr) * sizeof(double);             // curvature tile | scratch
    bytes += (size_t)(N + 1) * G * sizeof(int) + 8;                  // first sample index of each interval (+ alignment)
    bytes += (size_t)G * (4 * sizeof(double) + sizeof(int));         // length, sampling step, its reciprocal, largest curvature, rotation
    bytes += (size_t)(threads / 32) * G * (sizeof(double) + 2 * sizeof(int));  // arg-max partials
    return (bytes + 15) / 16 * 16;
}

// ------------------------------------------------------------------------------------------------
// K1a
// ------------------------------------------------------------------------------------------------
constexpr int K1A_THREADS = 96;  // warp = right-hand side, lane = candidate
__host__ __device__ inline size_t k1a_smem_bytes(int N) { return (size_t)5 * N * 32 * sizeof(double); }

__global__ void __launch_bounds__(K1A_THREADS, 4) k1a_solve(K1Args a)
{
    extern __shared__ __align__(16) double sm1[];
    const int N = a.N, NL = N * 32;
    double* RX = sm1;         // [N][32] control point x, then solution x, then M_x
    double* RY = RX + NL;
    double* RZ = RY + NL;     // corner-vector solution
    double* H = RZ + NL;      // chord lengths, then interval widths as the spline sees them
    double* CP = H + NL;      // modified super-diagonal
    const int tid = threadIdx.x, lane = tid & 31, rhs = tid >> 5;
    const long long b0 = (long long)blockIdx.x * 32;
#define S1(A, j) A[(j) * 32 + lane]
    // control points (track.py:87,:94): consecutive threads along a candidate's alpha row
    for (int idx = tid; idx < NL; idx += K1A_THREADS) {
        const int g = idx / N, j = idx - g * N;
        long long b = b0 + g;
        b = (b < a.B) ? b : a.B - 1;  // padding lanes repeat the last candidate
        double x, y;
        control_point(a, b, j, x, y);
        RX[j * 32 + g] = x;
        RY[j * 32 + g] = y;
    }
    __syncthreads();
    for (int idx = tid; idx < NL; idx += K1A_THREADS) {  // chord lengths (path.py:13)
        const int j = idx >> 5, g = idx & 31;
        const int jn = (j + 1 == N) ? 0 : j + 1;
        const double ex = RX[jn * 32 + g] - RX[idx], ey = RY[jn * 32 + g] - RY[idx];
        H[idx] = dsqrt<false>(ex * ex + ey * ey);
    }
    __syncthreads();
    if (rhs == 0) {  // knots: np.cumsum order (path.py:14); interval widths are knot differences
        double* kn = a.hand + hand_index(b0 + lane, 3 * N + 1, 0);
        double acc = 0.0;
        kn[0] = 0.0;
        for (int j = 0; j < N; ++j) {
            const double prev = acc;
            acc = acc + S1(H, j);
            kn[(size_t)(j + 1) * HAND_G] = acc;
            S1(H, j) = acc - prev;
        }
    }
    __syncthreads();
    {
        // forward elimination of this warp's right-hand side; row 0's corner term lives in the
        // Sherman-Morrison vector
        double* R = (rhs == 0) ? RX : (rhs == 1) ? RY : RZ;
        const double hl = S1(H, N - 1), h0 = S1(H, 0);
        const double b0d = 2.0 * (hl + h0);
        const double gamma = -b0d;
        const double p0 = (rhs < 2) ? S1(R, 0) : 0.0;
        const double dl = (rhs < 2) ? ddiv<false>(p0 - S1(R, N - 1), hl) : 0.0;  // closing chord slope
        // rows 0 and N-1 are peeled: they carry the only special cases (no sub-diagonal in row 0, the
        // Sherman-Morrison term and the closing slope in row N-1), so the N-2 rows between run branch-free
        double pj = p0, dprev = dl, hprev, cp, r;
        {   // row 0
            const double hj = h0;
            double f;
            if (rhs < 2) {
                const double pn = S1(R, 1);
                const double dj = ddiv<false>(pn - pj, hj);
                pj = pn;
                f = 6.0 * (dj - dprev);
                dprev = dj;
            } else {
                f = gamma;
            }
            const double inv = ddiv<false>(1.0, b0d - gamma);
            cp = hj * inv;
            r = f * inv;
            S1(CP, 0) = cp;
            S1(R, 0) = r;
            hprev = hj;
        }
        for (int j = 1; j < N - 1; ++j) {
            const double hj = S1(H, j);
            double f = 0.0;
            if (rhs < 2) {
                const double pn = S1(R, j + 1);
                const double dj = ddiv<false>(pn - pj, hj);
                pj = pn;
                f = 6.0 * (dj - dprev);
                dprev = dj;
            }
            const double aa = hprev;
            const double inv = ddiv<false>(1.0, 2.0 * (aa + hj) - aa * cp);
            cp = hj * inv;
            r = (f - aa * r) * inv;
            S1(CP, j) = cp;  // the three warps store identical values
            S1(R, j) = r;
            hprev = hj;
        }
        {   // row N-1
            const double f = (rhs < 2) ? 6.0 * (dl - dprev) : hl;
            const double aa = hprev;
            const double inv = ddiv<false>(1.0, (2.0 * (hprev + hl) - hl * hl / gamma) - aa * cp);
            cp = hl * inv;
            r = (f - aa * r) * inv;
            S1(CP, N - 1) = cp;
            S1(R, N - 1) = r;
        }
        for (int j = N - 2; j >= 0; --j) {  // back substitution
            r = S1(R, j) - S1(CP, j) * r;
            S1(R, j) = r;
        }
    }
    __syncthreads();
    if (rhs < 2) {  // Sherman-Morrison correction, store M_x / M_y
        const long long b = b0 + lane;
        const double* R = (rhs == 0) ? RX : RY;
        double* out = a.hand + hand_index(b, 3 * N + 1, (rhs == 0) ? N + 1 : 2 * N + 1);
        const double hl = S1(H, N - 1);
        const double gamma = -(2.0 * (hl + S1(H, 0)));
        const double vN = hl / gamma;
        const double denom = 1.0 + (S1(RZ, 0) + vN * S1(RZ, N - 1));
        const double fs = (S1(R, 0) + vN * S1(R, N - 1)) / denom;
        for (int j = 0; j < N; ++j) out[(size_t)j * HAND_G] = S1(R, j) - fs * S1(RZ, j);
    }
#undef S1
}

// ------------------------------------------------------------------------------------------------
// K1b
// ------------------------------------------------------------------------------------------------
#ifdef LTK_K1B_CLOCK  // developer build: phase time stamps of every CTA (scripts/k1b_phase_probe.py)
constexpr int K1B_CLOCK_SLOTS = 8;
__device__ long long g_k1b_clock[K1B_CLOCK_SLOTS * 65536];
__device__ __forceinline__ void k1b_stamp(int slot)
{
    if (threadIdx.x == 0 && blockIdx.x < 65536) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_k1b_clock[(size_t)blockIdx.x * K1B_CLOCK_SLOTS + slot] = t;
    }
}
#define K1B_STAMP(slot) k1b_stamp(slot)
#else
#define K1B_STAMP(slot)
#endif
// FIT = true: the records and the per-sample arithmetic of the FITPACK mode (ltk_fitpack_core.cuh); the
// hand-off arrays are then K1a-F's knots t [N+7] and derivative coefficients wrk1 [N+2], wrk2 [N+1] per coordinate.
// STAGED = false: two passes over the samples instead of the shared-memory tile -- the first finds the rotation,
// the second recomputes the curvatures and writes them straight to their rotated rows.  Twice the arithmetic,
// but full occupancy when the tile would not fit (ns = 10,001: 8.0 ms with the tile at G = 2 and one CTA per SM).
// Run at G = 8: half-line (64-byte) row segments; at G = 4 the 32-byte segments of four CTAs reached DRAM as
// partial lines (ncu: 0.56 GB of DRAM reads for a kernel that only writes).
// F32K = true (optional fp32 variant, ltk_set_sweep_precision(ctx, 32)): the sample parameter t = s - u_j stays fp64,
// the cubic's derivatives and the curvature are evaluated in fp32 (FFMA + one MUFU.RSQ instead of 33 FP64
// instructions with an fp64 square root and division), only the fp32 curvature array is written.
template <int G, int T, int MINB, bool FIT, bool STAGED, bool F32K = false>
__global__ void __launch_bounds__(T, MINB) k1b_samples(K1Args a, FitArgs fa)
{
    static_assert(!F32K || (!FIT && STAGED), "the fp32 evaluation exists for the staged tridiagonal kernel");
    static_assert(32 % G == 0 && T % 32 == 0, "lanes split evenly over the candidates of a CTA");
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ __align__(8) unsigned long long mbar;  // completion of the hand-off bulk copy
    const int N = a.N, n = a.ns - 1, NG = N * G;
    constexpr int NW = T / 32, CPT = T / G;
    using Rec = typename std::conditional<FIT, fit::FitInterval,
                                          typename std::conditional<K1B_TRIM, IntervalP, Interval>::type>::type;
    constexpr bool PACKED = !FIT && K1B_TRIM;  // the record carries the next interval's first sample index
    Rec* REC = reinterpret_cast<Rec*>(smraw);
    double* KT = reinterpret_cast<double*>(REC + NG);
    const size_t tile = STAGED ? (size_t)n * G : 0, scr = k1f_scratch_doubles(G, N, FIT);
    int* IB = reinterpret_cast<int*>(KT + ((STAGED && tile > scr) ? tile : scr));   // [N+1][G]
    double* LEN = reinterpret_cast<double*>(IB + (N + 1) * G + (((N + 1) * G) & 1));
    double* STEP = LEN + G;                                             // [G] np.linspace step (tbn.py:71)
    double* ISTEP = STEP + G;                                           // [G] ~ 1 / step: starting guesses only
    double* KTHR = ISTEP + G;                                           // [G] largest curvature (see the arg-min below)
    double* RV = KTHR + G;                                              // [NW][G]
    int* RI = reinterpret_cast<int*>(RV + NW * G);                      // [NW][G]
    int* RH = RI + NW * G;                                              // [NW][G]
    int* ROT = RH + NW * G;                                             // [G]
    // scratch inside the tile region (dead before the first curvature is stored)
    double* PX = KT;                 // [N][G] control points
    double* PY = PX + NG;
    double* U = PY + NG;             // [N+1][G] knots
    double* RX = U + NG + G;         // [N][G] second derivatives M_x, M_y
    double* RY = RX + NG;
    // FIT: knots t [N+7][G] (U = t + 3 rows), wrk1 x|y [N+2][G], wrk2 x|y [N+1][G]
    double* TK = KT;
    double* W1X = TK + (N + 7) * G;
    double* W1Y = W1X + (N + 2) * G;
    double* W2X = W1Y + (N + 2) * G;
    double* W2Y = W2X + (N + 1) * G;
    if (FIT) U = TK + 3 * G;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long b0 = (long long)blockIdx.x * G;
    K1B_STAMP(0);

    if constexpr (FIT) {
        // ---- L (FITPACK mode): everything comes from K1a-F: one bulk copy of the group's packed block (G = 4),
        //      which has exactly the layout of the scratch arrays TK | W1X | W1Y | W2X | W2Y ----------------------
        constexpr int ROWS_PER = 5;  // 5 N + 13 rows
        const int rows = ROWS_PER * N + 13;
        if constexpr (G == HAND_G) {
            const unsigned bytes = (unsigned)rows * HAND_G * sizeof(double);
            if (tid == 0) mbar_init(&mbar, 1);
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(&mbar, bytes);
                bulk_g2s(TK, fa.hand + (size_t)blockIdx.x * rows * HAND_G, bytes, &mbar);
            }
            mbar_wait(&mbar, 0);
        } else {
            for (int idx = tid; idx < rows * G; idx += T) {
                const int r = idx / G, g = idx - r * G;
                TK[idx] = fa.hand[hand_index(b0 + g, rows, r)];
            }
            __syncthreads();
        }
        if (tid < G) {
            LEN[tid] = U[NG + tid];
            STEP[tid] = U[NG + tid] / (double)(a.ns - 1);
            ISTEP[tid] = (double)(a.ns - 1) / U[NG + tid];
        }
        __syncthreads();
        // ---- C (FITPACK mode): per-interval record (splder / fpbspl operands) and first sample index ---------
        for (int idx = tid; idx < NG; idx += T) {
            const int j = idx / G, g = idx - j * G;
            fit::FitInterval rec;
            rec.tm1 = TK[(j + 2) * G + g]; rec.t0 = TK[(j + 3) * G + g];
            rec.tp1 = TK[(j + 4) * G + g]; rec.tp2 = TK[(j + 5) * G + g];
            rec.inv01 = ddiv<false>(1.0, rec.tp1 - rec.t0);
            rec.d1 = rec.tp1 - rec.tm1; rec.r1 = ddiv<false>(1.0, rec.d1);
            rec.d2 = rec.tp2 - rec.t0;  rec.r2 = ddiv<false>(1.0, rec.d2);
#pragma unroll
            for (int q = 0; q < 3; ++q) { rec.w1x[q] = W1X[(j + q) * G + g]; rec.w1y[q] = W1Y[(j + q) * G + g]; }
#pragma unroll
            for (int q = 0; q < 2; ++q) { rec.w2x[q] = W2X[(j + q) * G + g]; rec.w2y[q] = W2Y[(j + q) * G + g]; }
            rec.pad = 0.0;
            REC[idx] = rec;
            const double u0 = rec.t0;
            const double step = K1B_STEP ? STEP[g] : LEN[g] / (double)(a.ns - 1);  // np.linspace step (tbn.py:71)
            int i = 0;
            if (j > 0) {
                i = (int)ddiv<false>(u0, step);
                i = max(0, min(i, n));
                while (i < n && (double)i * step < u0) ++i;
                while (i > 0 && (double)(i - 1) * step >= u0) --i;
            }
            IB[idx] = i;
            if (j == N - 1) IB[NG + g] = n;
        }
        __syncthreads();
    } else {

        // ---- L: control points from the alphas; knots and second derivatives from K1a: one bulk copy of the
        //      group's packed block (G = 4), whose layout is that of the scratch arrays U | RX | RY; it is in flight
        //      while the control points are computed ----------------------------------------------------------
        const int rows = 3 * N + 1;
        if constexpr (G == HAND_G) {
            if (tid == 0) mbar_init(&mbar, 1);
            __syncthreads();
            if (tid == 0) {
                const unsigned bytes = (unsigned)rows * HAND_G * sizeof(double);
                mbar_expect_tx(&mbar, bytes);
                bulk_g2s(U, a.hand + (size_t)blockIdx.x * rows * HAND_G, bytes, &mbar);
            }
#if LTK_K1B_PREFETCH
            // the CTA that will run on this slot a couple of generations from now starts with two cold reads (its
            // packed hand-off block and its four alpha rows): pull them into L2 now
            if (tid == 32 && a.mode == 0) {
                const long long nb = (long long)blockIdx.x + LTK_K1B_PREFETCH;
                if ((nb + 1) * G <= a.B) {
                    bulk_prefetch_l2(a.hand + (size_t)nb * rows * HAND_G, (unsigned)rows * HAND_G * sizeof(double));
                    const double* al = a.alphas + (size_t)nb * G * N;  // 4 rows of N doubles, 32 N bytes: 16-byte multiple
                    bulk_prefetch_l2(al, (unsigned)(G * N * sizeof(double)));
                }
            }
#endif
        }
        {   // T / G threads per candidate, j fastest: a candidate's alpha row is contiguous (no division by N)
            const int g = tid / CPT;
            long long b = b0 + g;
            b = (b < a.B) ? b : a.B - 1;             // padding lanes repeat the last candidate
            for (int j = tid - g * CPT; j < N; j += CPT) {
                double x, y;
                control_point(a, b, j, x, y);
                PX[j * G + g] = x;
                PY[j * G + g] = y;
            }
        }
        if constexpr (G == HAND_G) {
            mbar_wait(&mbar, 0);
        } else {
            for (int idx = tid; idx < rows * G; idx += T) {
                const int r = idx / G, g = idx - r * G;
                U[idx] = a.hand[hand_index(b0 + g, rows, r)];
            }
            __syncthreads();
        }
        if (tid < G) {
            LEN[tid] = U[NG + tid];
            STEP[tid] = U[NG + tid] / (double)(a.ns - 1);
            ISTEP[tid] = (double)(a.ns - 1) / U[NG + tid];
        }
        __syncthreads();
        K1B_STAMP(1);
        // ---- C: per-interval coefficients  S'(t) = c1 + t (c2 + t h),  S''(t) = c2 + c3 t,  h = c3/2, and the
        //      first sample index of every interval: IB[j] = min{ i : fl(i*step) >= U[j] } ------------------
        for (int idx = tid; idx < NG; idx += T) {
            const int j = idx / G, g = idx - j * G;
            const int jn = (j + 1 == N) ? 0 : j + 1;
            const double u0 = U[idx], u1 = U[idx + G], h = u1 - u0;  // the spline only sees the knots
            const double mx = RX[idx], mxn = RX[jn * G + g];
            const double my = RY[idx], myn = RY[jn * G + g];
            const double c3x = ddiv<false>(mxn - mx, h), c3y = ddiv<false>(myn - my, h);
            Rec rec;
            rec.u = u0;
            // x / 6 with RN(1/6) and one residual correction is the correctly rounded quotient (Markstein; checked
            // against IEEE division on 4e8 operands): 3 FP64 instructions instead of 9 and a MUFU
            auto sixth = [](double x) {
                return K1B_CREC ? div_by_const<false>(x, 6.0, 1.0 / 6.0) : ddiv<false>(x, 6.0);
            };
            rec.c1x = ddiv<false>(PX[jn * G + g] - PX[idx], h) - sixth(h * (2.0 * mx + mxn));
            rec.c1y = ddiv<false>(PY[jn * G + g] - PY[idx], h) - sixth(h * (2.0 * my + myn));
            rec.c2x = mx; rec.c2y = my;
            rec.c3x = c3x; rec.c3y = c3y;
            if constexpr (PACKED) {
                rec.inext = 0; rec.pad = 0;  // (the next interval's first sample comes from IB, see narrow())
            } else {
                rec.unext = u1;
                rec.hx = 0.5 * c3x; rec.hy = 0.5 * c3y;
            }
            REC[idx] = rec;
            const double step = K1B_STEP ? STEP[g] : LEN[g] / (double)(a.ns - 1);  // np.linspace step (tbn.py:71)
            int i = 0;
            if (j > 0) {
                // any starting guess will do: the two loops below enforce the definition exactly
                i = K1B_CREC ? (int)(u0 * ISTEP[g]) : (int)ddiv<false>(u0, step);
                i = max(0, min(i, n));
                while (i < n && (double)i * step < u0) ++i;
                while (i > 0 && (double)(i - 1) * step >= u0) --i;
            }
            IB[idx] = i;
            if (j == N - 1) IB[NG + g] = n;
        }
        __syncthreads();  // scratch is dead from here on: the tile region now takes curvatures
    }
    K1B_STAMP(2);

    // ---- K: curvature at the samples ----------------------------------------------------------------------
    const int g = tid % G, c = tid / G;
    // balanced chunks: every thread of a candidate gets n / CPT samples, the first n % CPT one more
    const int cbase = n / CPT, cextra = n - cbase * CPT;
    const int i0 = c * cbase + min(c, cextra), i1 = i0 + cbase + (c < cextra ? 1 : 0);
    const double step = K1B_STEP ? STEP[g] : LEN[g] / (double)(a.ns - 1);
    // One flat loop per run of samples: every lane of the warp runs the same number of iterations (nested
    // per-interval loops diverge -- interval boundaries differ from lane to lane -- and ran at 19 of
    // 32 lanes); the interval switch is an integer test that fires about once per 20 samples, the wrap of the
    // sample index (second pass of the two-pass variant: the lanes' rotations differ) at most once per run.
    auto walk = [&](int qa, int count, auto&& sink) {  // samples qa, qa+1, ... (mod n) of candidate g
        if (count <= 0) return;
        int q = (qa >= n) ? qa - n : qa;
        int j;  // interval of sample q: largest j with IB[j] <= q
        if constexpr (K1B_GUESS) {
            // the knots follow the track's own spacing: start from the proportional guess and walk (a step or two)
            j = min(N - 1, (int)(((long long)q * N) / n));
            while (j + 1 < N && IB[(j + 1) * G + g] <= q) ++j;
            while (j > 0 && IB[j * G + g] > q) --j;
        } else {
            int lo = 0, hi = N - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (IB[mid * G + g] <= q) lo = mid; else hi = mid - 1;
            }
            j = lo;
        }
        Rec v = REC[j * G + g];
        int inext;
        [[maybe_unused]] double hx = 0, hy = 0;
        [[maybe_unused]] float f1x = 0, f2x = 0, f3x = 0, fhx = 0, f1y = 0, f2y = 0, f3y = 0, fhy = 0;
        auto narrow = [&]() {
            inext = IB[(j + 1) * G + g];
            if constexpr (PACKED) {
                hx = 0.5 * v.c3x; hy = 0.5 * v.c3y;
            } else if constexpr (!FIT) {
                hx = v.hx; hy = v.hy;
            }
            if constexpr (F32K) {
                f1x = (float)v.c1x; f2x = (float)v.c2x; f3x = (float)v.c3x; fhx = (float)hx;
                f1y = (float)v.c1y; f2y = (float)v.c2y; f3y = (float)v.c3y; fhy = (float)hy;
            }
        };
        narrow();
        for (int r = 0; r < count; ++r) {
            while (q >= inext) {
                ++j;
                v = REC[j * G + g];
                narrow();
            }
            const double s = (double)q * step;
            double k;
            if constexpr (FIT) {
                double dx, dy, ddx, ddy;
                k = fit::curvature_at(v, s, dx, dy, ddx, ddy);
            } else if constexpr (F32K) {
                const float t = (float)(s - v.u);
                const float ddx = fmaf(f3x, t, f2x), ddy = fmaf(f3y, t, f2y);
                const float dx = fmaf(t, fmaf(fhx, t, f2x), f1x), dy = fmaf(t, fmaf(fhy, t, f2y), f1y);
                const float cross = fabsf(fmaf(dx, ddy, -(dy * ddx)));
                const float n2 = fmaf(dx, dx, dy * dy);
                float rs;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(n2));
                k = (double)(cross * rs * (rs * rs));  // exact widening: the tile and the arg-max stay fp64
            } else {
                // |x'y'' - y'x''| / (x'^2 + y'^2)^(3/2)   (path.py:58,61), Horner form, explicit FMAs
                const double t = s - v.u;
                const double ddx = fma(v.c3x, t, v.c2x), ddy = fma(v.c3y, t, v.c2y);
                const double dx = fma(t, fma(hx, t, v.c2x), v.c1x);
                const double dy = fma(t, fma(hy, t, v.c2y), v.c1y);
                const double cross = fabs(fma(dx, ddy, -(dy * ddx)));
                const double n2 = fma(dx, dx, dy * dy);
                k = ddiv<false>(cross, n2 * dsqrt<false>(n2));
            }
            sink(r, q, k);
            if (!STAGED && ++q == n) {  // wrap: back to the first interval
                q = 0; j = 0;
                v = REC[g];
                narrow();
            }
            if (STAGED) ++q;
        }
    };
    // The sweeps start at np.argmin(v_local), v_local = sqrt(mu g / k) (velocity.py:28-29, :34, :58): the FIRST sample
    // whose v_local is minimal.  That is the first maximum of the curvature unless an EARLIER sample, a few ulps
    // below the maximum, rounds to the same v_local (plateaus: circular arcs, symmetric candidates).  Such a sample
    // shares the maximum's upper 32 bits (or sits one below), so each partial result carries, next to the first
    // maximum, the largest upper word seen BEFORE it -- one select per sample -- and a CTA that finds the two words
    // within one of each other (about one candidate in a thousand on real tracks) looks again, exactly.
    double best = -1.0;
    int bi = 0, bh = 0;  // bh: upper word of the largest curvature among the samples preceding bi (0: none)
    walk(i0, i1 - i0, [&](int, int q, double k) {
        if constexpr (STAGED) KT[(size_t)q * G + g] = k;
        const bool gt = k > best;
        if (K1B_TIES) bh = gt ? __double2hiint(best) : bh;
        best = gt ? k : best;
        bi = gt ? q : bi;
    });
    K1B_STAMP(3);
    // Lanes l, l+G, l+2G, ... hold consecutive chunks of one candidate: fold the upper lanes into the lower ones,
    // lower chunk first
#pragma unroll
    for (int o = G; o < 32; o <<= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oh = __shfl_down_sync(0xffffffffu, bh, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if ((lane % (2 * o)) < o && lane + o < 32 && ob > best) {
            bh = max(__double2hiint(best), oh);  // curvatures are >= 0 or the -1 start value: signed order = value order
            best = ob; bi = oi;
        }
    }
    if (lane < G) { RV[warp * G + lane] = best; RH[warp * G + lane] = bh; RI[warp * G + lane] = bi; }
    __syncthreads();
    bool tie = false;
    if (tid < G) {
        double bb = -1.0;
        int bbi = 0, bbh = 0;
        for (int w = 0; w < NW; ++w) {
            const double v = RV[w * G + tid];
            if (v > bb) {
                bbh = max(__double2hiint(bb), RH[w * G + tid]);
                bb = v; bbi = RI[w * G + tid];
            }
        }
        ROT[tid] = bbi;
        KTHR[tid] = bb;  // the maximum (read again below if some candidate of the CTA has a tie to examine)
        tie = K1B_TIES && bbh > 0 && bbh >= __double2hiint(bb) - 1;
    }
    if (__syncthreads_or(tie)) {  // rare: find the first sample of the plateau (all threads, CTA-uniform branch)
        const double kmax = KTHR[g];
        const double vmin = sqrt(a.mu_g / kmax), kthr = kmax * (1.0 - 0x1p-48);  // one ulp of k moves v_local by half an ulp
        int first = 0x7fffffff;
        auto look = [&](int q, double k) {
            if (first == 0x7fffffff && k >= kthr && sqrt(a.mu_g / k) == vmin) first = q;
        };
        if constexpr (STAGED) {
            for (int q = i0; q < i1; ++q) look(q, KT[(size_t)q * G + g]);
        } else {
            walk(i0, i1 - i0, [&](int, int q, double k) { look(q, k); });
        }
#pragma unroll
        for (int o = G; o < 32; o <<= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        if (lane < G) RI[warp * G + lane] = first;
        __syncthreads();
        if (tid < G) {
            int f = 0x7fffffff;
            for (int w = 0; w < NW; ++w) f = min(f, RI[w * G + tid]);
            if (f < ROT[tid]) ROT[tid] = f;
        }
        __syncthreads();
    }
    if (tid < G) {
        a.rot[b0 + tid] = ROT[tid];
        a.len[b0 + tid] = LEN[tid];
    }
    K1B_STAMP(4);
    // ---- W: write-out, rotated: row i of candidate g is sample (i + rot_g) mod n ----------------------------
    {
        const int q0 = ROT[g];
        double* dst = a.kap + tile_base(b0 + g, n);
        float* dst32 = a.kap32 ? a.kap32 + tile_base(b0 + g, n) : nullptr;
        if constexpr (STAGED && K1B_WLIN) {
            // rotated row i = c, c + CPT, ... takes tile row i + q0 (- n past the wrap): the four candidates of a row
            // segment stay in the same iteration (one full 32-byte sector per store), the source cursor steps back
            // by n rows once; the fp32 copy has its own loop (no predicated-off stores in the fp64 one)
            const double* src = KT + (size_t)(c + q0) * G + g;
            const int iw = n - q0;  // first rotated row past the wrap
            if (F32K || dst32) {
                for (int i = c; i < n; i += CPT) {
                    const double k = src[(size_t)(i - c) * G - ((i >= iw) ? (size_t)n * G : 0)];
                    if (!F32K) dst[(size_t)i * TILE] = k;
                    dst32[(size_t)i * TILE] = (float)k;
                }
            } else {
                double* d = dst + (size_t)c * TILE;
#pragma unroll 4
                for (int i = c; i < n; i += CPT, d += (size_t)CPT * TILE)
                    *d = src[(i - c) * G - ((i >= iw) ? n * G : 0)];
            }
        } else if constexpr (STAGED) {
            for (int i = c; i < n; i += CPT) {
                int q = i + q0;
                q = (q >= n) ? q - n : q;
                const double k = KT[(size_t)q * G + g];
                if (!F32K) dst[(size_t)i * TILE] = k;
                if (dst32) dst32[(size_t)i * TILE] = (float)k;
            }
        } else {
            // second pass: this thread's rows i0 .. i1-1 are the samples i0+q0 .. i1+q0-1 (mod n); the G
            // candidates of a row are written by adjacent lanes (64 contiguous bytes at G = 8)
            walk(i0 + q0, i1 - i0, [&](int r, int, double k) {
                const size_t row = (size_t)(i0 + r) * TILE;
                dst[row] = k;
                if (dst32) dst32[row] = (float)k;
            });
        }
    }
    K1B_STAMP(5);
}

}  // namespace ltk
