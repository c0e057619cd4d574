/* ltk.h -- C ABI of the B200-native batched lap-time kernels ("ltk").
 *
 * Drop-in boundary for the one data-parallel hot path of bruno-maruszczak/lap-time-optimization:
 * scoring candidate racing lines (alpha vectors) to lap times.  The reference has no FFI of its own --
 * its boundary is the Python class surface Path / VelocityProfile / Trajectory.lap_time /
 * TrajectoryBayesianNonlinear.calcMinTime -- so each entry point below names the reference lines it
 * replaces (paths relative to the reference's src/).  The Python facade in
 * lap_time_optimization_b200/ binds these with ctypes; INTEGRATION.md shows the stub a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - plain C types only; every array argument named d_* is a CALLER-OWNED DEVICE pointer (e.g.
 *     torch.Tensor.data_ptr()); h_* are host pointers read during the call only.  Nothing is retained
 *     or freed by the library except what ltk_create allocates inside the context.
 *   - work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*; NULL = the
 *     legacy default stream) and the call returns without synchronising unless stated.
 *   - return value: 0 = OK, negative = error (LTK_E_*); ltk_last_error(ctx) gives the message.
 *   - a context is bound to one device; contexts are not thread-safe; no exceptions cross the ABI.
 *   - arithmetic is IEEE fp64 throughout (no fast-math); there is no CPU fallback.
 */
#ifndef LTK_H
#define LTK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTK_MAX_ENGINE_MAP 16

#define LTK_OK 0
#define LTK_E_ARG -1       /* bad argument */
#define LTK_E_CUDA -2      /* CUDA runtime error (see ltk_last_error) */
#define LTK_E_WORKSPACE -3 /* workspace too small */
#define LTK_E_UNSUPPORTED -4

/* Vehicle constants, pre-folded on the host in the reference's own operation order.
 * kind 0: tabulated engine map -- reference vehicle.py:11-35 (TBR18)
 * kind 1: polynomial engine    -- reference vehicleMX5.py:19-37 (MX-5)                      */
typedef struct ltk_vehicle {
    int32_t kind;
    int32_t n_map;                       /* kind 0: number of engine-map nodes (2..16) */
    double mass;                         /* vehicle.py:16 / vehicleMX5.py:54 */
    double mu_g;                         /* friction_coef * 9.81          (velocity.py:29) */
    double f_max;                        /* (mu*m)*g  (vehicle.py:30)  |  (lam*D)*(m*g) (vehicleMX5.py:28-33) */
    double f_max_sq;                     /* f_max**2 */
    double map_v[LTK_MAX_ENGINE_MAP];    /* engineMap.v, ascending (vehicle.py:18-21) */
    double map_f[LTK_MAX_ENGINE_MAP];    /* engineMap.f */
    double e0;                           /* kind 1: (T*C_m) - Cr_0        (vehicleMX5.py:21) */
    double cr2;                          /* kind 1: Cr_2 */
} ltk_vehicle;

typedef struct ltk_ctx ltk_ctx;

/* Create an evaluator for one (track, vehicle, sampling) triple.
 *   h_left_xy, h_diff_xy : [2][n_ctrl] row-major (x row, y row) -- inner boundary point and
 *       inner->outer vector of each UNIQUE control point (the closing duplicate is implied).
 *       Replaces the constants behind Track.control_points / control_points_bayesian
 *       (track.py:82-94) including the closure quirk of trajectory_bayesian_nonlinear.py:58-69.
 *   ns : samples per lap including the end point (Trajectory.ns, trajectory.py:35); ns-1 are swept.
 *   Only closed tracks are supported by the batched path (all four reference tracks are closed). */
int ltk_create(ltk_ctx **out, int device, const double *h_left_xy, const double *h_diff_xy,
               int n_ctrl, const ltk_vehicle *vehicle, int ns);
void ltk_destroy(ltk_ctx *ctx);
const char *ltk_last_error(const ltk_ctx *ctx); /* ctx may be NULL: last creation error */

/* Change the sampling density (Trajectory.ns is a plain attribute in the reference). */
int ltk_set_ns(ltk_ctx *ctx, int ns);

/* Optional fp32 variant of the velocity sweeps (BASELINE.json north_star: 1e-9 for the fp64 kernels,
 * 1e-4 for an optional fp32 variant): bits = 32 runs velocity.py:14-76 in IEEE fp32 on the fp64
 * curvature of the spline kernels, lap sum in fp64; bits = 64 (the default) restores the fp64 sweeps.
 * Affects ltk_eval_alphas / ltk_eval_controls; ltk_profile always uses fp64. */
int ltk_set_sweep_precision(ltk_ctx *ctx, int bits);

/* Spline arithmetic (Path.__init__ / Path.curvature, path.py:25, :51-61).  Both modes build the periodic
 * interpolating cubic spline through the control points with chord-length knots; they differ in how.
 *   LTK_SPLINE_TRIDIAGONAL (default): cyclic tridiagonal solve for the second derivatives (Thomas +
 *       Sherman-Morrison), Horner evaluation.  Curvature agrees with SciPy's to ~1e-13 of its maximum; TBR18's
 *       friction-circle cancellation (vehicle.py:29-35) turns that into lap-time differences of median 2e-11,
 *       but > 1e-9 on about one candidate in 2,400 (27 of 65,536, worst 6.5e-9).
 *   LTK_SPLINE_FITPACK: SciPy FITPACK's own algorithm and operation order -- clocur/fpclos (Givens QR of the
 *       periodic collocation matrix, fpbacp) for splprep(k=3, s=0, per=1), splder/fpbspl for splev(der=1|2):
 *       B-spline coefficients and derivative values bit-equal to SciPy's, curvature with a correctly rounded
 *       x**1.5 (numpy's pow is libm's, or an SVML routine one ulp off on AVX512 hosts -- the reference's own TBR18
 *       lap times move by up to 6.5e-9 between the two).  Lap times then stay within 1e-9 of the reference's on
 *       every candidate (measured: 0 of 65,536 beyond it, worst 1.9e-10, median 3e-16, against the reference on
 *       numpy's baseline dispatch; against an AVX512 host they inherit that host's own distance).  Costs ~4 KB more
 *       workspace per candidate (ask ltk_workspace_bytes AFTER setting the mode) and a slower K1. */
#define LTK_SPLINE_TRIDIAGONAL 0
#define LTK_SPLINE_FITPACK 1
int ltk_set_spline_mode(ltk_ctx *ctx, int mode);
int ltk_spline_mode(const ltk_ctx *ctx);

/* Tracing: after ltk_trace_begin(ctx, n) every kernel launch of ltk_eval_* on this context records a CUDA
 * event before and after itself on its stream (up to n launches; n = 0 switches tracing off).
 * ltk_trace_read waits for the recorded launches and returns kind (LTK_TRACE_*), start and end in
 * milliseconds relative to the FIRST record of `base` (another traced context of the same device, or
 * NULL for ctx itself), so that contexts working side by side can be put on one time line.  The
 * reference only brackets whole optimiser runs with time.time() (trajectory.py:67-170). */
#define LTK_TRACE_K1A 1
#define LTK_TRACE_K1B 2
#define LTK_TRACE_K23 3
int ltk_trace_begin(ltk_ctx *ctx, int max_records);
int ltk_trace_read(ltk_ctx *ctx, const ltk_ctx *base, int max_records, int *h_kind, float *h_start_ms,
                   float *h_end_ms, int *n_out);

/* Scheduling knob, no effect on results.  on (default): a population that fills the sweep kernel's warp
 * slots unevenly (65,536 candidates = 3.46 warps per scheduler) is split -- whole layers of one warp per
 * scheduler to the two-chain kernel, the remainder to the one-chain kernel on a second stream of the
 * context (0.527 -> 0.487 ms).  off: one sweep launch; callers that keep several populations in flight
 * on different contexts switch it off (the extra concurrent kernel costs more than it gains there). */
int ltk_set_sweep_split(ltk_ctx *ctx, int on);

/* Bytes of device scratch ltk_eval_* needs for a batch of B candidates. */
int ltk_workspace_bytes(const ltk_ctx *ctx, int64_t B, size_t *out_bytes);

/* alphas -> lap times.  d_alphas [B][n_ctrl] row-major fp64, d_lap [B].
 * Replaces, per candidate: TrajectoryBayesianNonlinear.calcMinTime(updateAlphas(a))
 * (trajectory_bayesian_nonlinear.py:58-80) and Trajectory.update/update_velocity/lap_time
 * (trajectory.py:40-58): Track.control_points* (track.py:82-94), Path.__init__ (path.py:20-26),
 * np.linspace sampling, Path.curvature (path.py:36-61), VelocityProfile (velocity.py:14-76),
 * Vehicle.engine_force/traction (vehicle.py:25-35, vehicleMX5.py:19-37), lap_time (:51-54).
 * Kernels launched: K1a spline solve, K1b curvature (rotated write-out), K23 forward + backward sweeps
 * with the lap-time sum. */
int ltk_eval_alphas(ltk_ctx *ctx, const double *d_alphas, int64_t B, double *d_lap,
                    void *d_workspace, size_t workspace_bytes, void *stream);

/* Host in, host out: h_alphas [B][n_ctrl] and h_lap [B] are ordinary host arrays; returns when h_lap is
 * filled.  This is the call the reference's optimiser loops make once per objective evaluation --
 * scipy's COBYLA through calcMinTime (trajectory_bayesian_nonlinear.py:207-227) and L-BFGS-B through
 * Trajectory.lap_time (trajectory.py:128-146) -- batched over the finite-difference points or the
 * lock-step starts.  The context owns the pinned staging, device buffers and stream; for B <= 16,384 the
 * upload, K1a, K1b, K23 and the download are one instantiated CUDA graph per batch size (up to 8 sizes
 * cached, rebuilt after ltk_set_*), so a call costs one graph launch and one stream synchronisation.
 * Not for concurrent use on one context. */
int ltk_eval_alphas_host(ltk_ctx *ctx, const double *h_alphas, int64_t B, double *h_lap);

/* Measurement hook: same work as ltk_eval_alphas, with CUDA events recorded on `stream` around each
 * kernel; synchronises and writes the durations in milliseconds to h_ms[4] = {K1a, K1b, K23, -1}. */
int ltk_eval_alphas_timed(ltk_ctx *ctx, const double *d_alphas, int64_t B, double *d_lap,
                          void *d_workspace, size_t workspace_bytes, void *stream, float *h_ms);

/* alphas -> curvature objectives of the same spline: d_gamma2[B] = sum of squared curvatures at ALL ns
 * samples (Path.gamma2(self.s), path.py:63-77 as called from trajectory.py:60-97), d_length[B] =
 * Path.length (path.py:26).  Either output may be NULL.  Replaces, per candidate, the objective of
 * Trajectory.minimise_curvature / minimise_compromise.  Kernels: K1a, K1b, one reduction. */
int ltk_eval_objectives(ltk_ctx *ctx, const double *d_alphas, int64_t B, double *d_gamma2,
                        double *d_length, void *d_workspace, size_t workspace_bytes, void *stream);

/* Candidate generation on the device: d_out[i] = low + (high - low) * u_(first + i), u the stream of
 * numpy's Generator(Philox(key=[key0, key1])).random() -- so that
 *     Generator(Philox(key=[key0, key1])).uniform(low, high, n)[first : first + count]
 * on a host equals the device array bit for bit.  Replaces the per-element np.random.uniform(0, 0.99)
 * of the population stages (trajectory_bayesian_nonlinear.py:142, :244); ranks generate disjoint
 * slices of one stream by their `first`. */
int ltk_random_uniform(int device, uint64_t key0, uint64_t key1, int64_t first, int64_t count, double low,
                       double high, double *d_out, void *stream);

/* control points -> lap times: the calcMinTime(controls) surface
 * (trajectory_bayesian_nonlinear.py:65-80).  d_xy [B][2][m] row-major with m = n_ctrl + 1 columns
 * (the last column is the closing duplicate and is ignored, exactly as splprep(per=1) overwrites it). */
int ltk_eval_controls(ltk_ctx *ctx, const double *d_xy, int m, int64_t B, double *d_lap,
                      void *d_workspace, size_t workspace_bytes, void *stream);

/* Full profile of ONE candidate in natural sample order, for the facade attributes
 * (VelocityProfile.v / v_local / v_acclim / v_declim, Trajectory.s; velocity.py:20-26).
 * Each output is a device array of ns-1 doubles (d_s: ns doubles) and may be NULL; d_scalars[2] =
 * {lap time, path length}.  Synchronises the stream. */
int ltk_profile(ltk_ctx *ctx, const double *d_alpha, double *d_s, double *d_k, double *d_vlocal,
                double *d_vacc, double *d_vdec, double *d_v, double *d_scalars, void *stream);

/* Stable ascending top-k of lap times: what `sorted(zip(taus, alphas), key=tau)[0:10]` selects
 * (trajectory_bayesian_nonlinear.py:253-257).  Ties keep the lower index; NaN sorts last.
 * d_best_idx holds index_base + position.  k <= 64. */
int ltk_topk(ltk_ctx *ctx, const double *d_lap, int64_t B, int64_t index_base, int k,
             double *d_best_lap, int64_t *d_best_idx, void *stream);

/* alphas -> lap times AND the population's k best in one call: what the random-population stage does with its
 * scores, `sorted(zip(taus, alphas), key=tau)[0:10]` after the scoring loop
 * (trajectory_bayesian_nonlinear.py:236-257).  Same results as ltk_eval_alphas followed by ltk_topk on the same
 * stream; the selection runs in the epilogue of the sweep kernel (every CTA selects its own k best from the lap
 * times it holds in registers, the CTA that finishes last merges them) -- three kernel launches per population
 * instead of four, no second pass over d_lap.  Falls back to the separate selection kernel where the fused
 * merge does not fit (k > 16, populations beyond ~600,000 candidates per call, the fp32 sweep variant). */
int ltk_eval_alphas_topk(ltk_ctx *ctx, const double *d_alphas, int64_t B, double *d_lap, void *d_workspace,
                         size_t workspace_bytes, int64_t index_base, int k, double *d_best_lap,
                         int64_t *d_best_idx, void *stream);

/* Path facade (path.py:17-77): evaluate the periodic spline through m-1 unique points
 * d_xy [2][m] (closed; last column ignored) with knots d_knots [m] (Path.dists) at n parameters d_u.
 * Any of the outputs may be NULL. d_gamma2[1] receives sum(k^2) (path.py:63-77). */
int ltk_path_eval(int device, const double *d_xy, const double *d_knots, int m, const double *d_u,
                  int64_t n, double *d_x, double *d_y, double *d_dx, double *d_dy, double *d_ddx,
                  double *d_ddy, double *d_k_signed, double *d_gamma2, void *stream);

/* The same facade in SciPy FITPACK's own arithmetic (see LTK_SPLINE_FITPACK), closed or open:
 * splprep(d_xy, u=d_knots, k=3, s=0, per=closed) (path.py:25) -- fpclos for a closed path (m-1 unique points, last
 * column ignored; m >= 6), fppara with the not-a-knot knot vector for an open one (m >= 4; path.py:25 with
 * closed = False, reached from trajectory.py:189-192) -- and splev(d_u, tck, der=0|1|2) (path.py:33, :51-54).
 * Knots, coefficients, positions and derivatives come out bit-equal to SciPy's; the curvature uses a correctly
 * rounded x**1.5.  d_t [m + 6 closed | m + 4 open] and d_c [2][m + 2 closed | m open] receive the tck that
 * Path.spline holds in the reference (either may be NULL; n may be 0). */
int ltk_path_eval_fitpack(int device, const double *d_xy, const double *d_knots, int m, int closed,
                          const double *d_u, int64_t n, double *d_x, double *d_y, double *d_dx, double *d_dy,
                          double *d_ddx, double *d_ddy, double *d_k_signed, double *d_gamma2, double *d_t,
                          double *d_c, void *stream);

/* VelocityProfile facade (velocity.py:9-76) for caller-supplied samples: d_s, d_k [n];
 * s_max < 0 means an open path (s_max=None in the reference). Outputs [n] each; d_vacc and d_vdec are
 * required (they are the sweep state), d_vlocal and d_v may be NULL. */
int ltk_velocity_profile(int device, const ltk_vehicle *vehicle, const double *d_s, const double *d_k,
                         int64_t n, double s_max, double *d_vlocal, double *d_vacc, double *d_vdec,
                         double *d_v, void *stream);

/* Merge step of the multi-GPU top-k: stable ascending top-k of explicit (lap, index) pairs, e.g. the
 * all-gathered per-rank winners. Entries with index < 0 are padding. */
int ltk_topk_pairs(ltk_ctx *ctx, const double *d_lap, const int64_t *d_idx, int64_t count, int k,
                   double *d_best_lap, int64_t *d_best_idx, void *stream);

/* Library/ABI version and a count of kernel launches issued through this library since load
 * (bench.py reports it as gpu_launches). */
/* The cross-rank half of the multi-GPU selection in one launch: d_gathered is the all-gathered buffer
 * [world][2][k_in] of int64 -- per rank, k_in lap times as bit patterns followed by k_in global indices, i.e. what
 * ltk_eval_alphas_topk leaves on a rank when it is given d_best_lap = buf and d_best_idx = buf + k_in -- and the
 * result is the stable top-k of all world * k_in pairs (entries with index < 0 are padding).  One collective
 * (all-gather of 16 k_in bytes per rank) and this call replace `sorted(results)[0:10]` over the ranks' lists
 * (trajectory_bayesian_nonlinear.py:253-257).  world * k_in <= 1,024. */
int ltk_topk_gathered(ltk_ctx *ctx, const int64_t *d_gathered, int world, int k_in, int k, double *d_best_lap,
                      int64_t *d_best_idx, void *stream);

int ltk_version(void);
int64_t ltk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LTK_H */
