// ltk_sweep_fused.cuh -- K23: forward and backward sweeps fused into one kernel.
//
// One thread per candidate runs BOTH recurrences at once: the forward chain visits rotated rows
// 1, 2, ..., n-1 while the backward chain visits rows n-1, n-2, ..., 1 (velocity.py:31-53 / :55-76; the
// two passes of the reference are independent of each other -- only the final minimum couples them).
// Two independent dependency chains per thread double the instruction-level parallelism of a kernel
// that is otherwise latency-bound at the ~3.5 warps per scheduler a 65,536-candidate population gives.
//
//   phase 1 (first half of the steps): each chain parks its result in the staging array S
//           (forward -> rows 1..h, backward -> rows n-1..n-h; disjoint);
//   middle  (only when the number of rows n-1 is odd): both chains land on the same row;
//   phase 2 (second half): each chain meets the rows the other one parked, takes the minimum
//           (velocity.py:26) and accumulates ds/v (tbn.py:51-54).
//
// HBM traffic is exactly that of the separate kernels: every curvature row read twice, every staged
// value written once and read once.
//
// The kernel is bound by the FP64 pipe (and by issue slots next to it), so the regular path is written
// to issue as few instructions as the reference's arithmetic allows without changing a single bit:
//   * min before sqrt.  velocity.py:44-50 computes  v = v_local  if v_local <= v_prev  else
//     min(v_local, sqrt(v_prev^2 + 2 a ds)),  v_local = sqrt(mu g / k).  With a >= 0 this equals
//     sqrt(min(mu g / k, v_prev^2 + 2 a ds)) bit for bit: correctly rounded sqrt is monotone, so
//     min(sqrt x, sqrt y) = sqrt(min(x, y)); and when v_local <= v_prev the squared limit
//     w = RN(v_prev^2 + ..) >= RN(v_local^2) gives sqrt(min(.)) in [sqrt(RN(v_local^2)), v_local] =
//     {v_local}.  One square root and one comparison per step disappear.  (a >= 0: traction is >= 0 by
//     construction; the engine force is >= 0 on the regular path, see VehDev::k_lo_hi.)
//   * engine map through a cell table indexed by the leading bits of v (EngineLut) instead of one
//     comparison per node.
//   * the np.linspace clock: the wrap-around of the sample index happens once per chain and lane, so
//     blocks in which no lane of the warp wraps run a select-free clock (warp vote).
#pragma once

namespace ltk {

constexpr int FUSED_THREADS = 64;
#ifndef LTK_FUSED_UNROLL
#define LTK_FUSED_UNROLL 2  // row pairs per block of the regular path (A/B: 3, 4 -- more registers, fewer block headers)
#endif
constexpr int FUSED_UNROLL = LTK_FUSED_UNROLL;
// bytes of readable memory the workspace keeps in front of the curvature array (ws_layout): the look-ahead
// loads of the sweep run 2 * FUSED_UNROLL rows past the rows a chain uses, unconditionally
constexpr size_t SWEEP_SLACK = 4096;
#ifndef LTK_SWEEP_PREFETCH
#define LTK_SWEEP_PREFETCH 2  // blocks of FUSED_UNROLL rows fetched into L1 ahead of the register look-ahead (0: none)
#endif
#ifndef LTK_SWEEP_LOOKAHEAD
#define LTK_SWEEP_LOOKAHEAD 1  // 1: the next block's rows are loaded into registers while the current block runs;
#endif                         // 0: every block loads its own rows (L1 hits behind the prefetch), no look-ahead registers
#ifndef LTK_SWEEP_MINB
#define LTK_SWEEP_MINB 8       // resident CTAs per SM the register budget is cut for
#endif
constexpr int SWEEP_LA = LTK_SWEEP_LOOKAHEAD ? 1 : 0;
static_assert(SWEEP_SLACK >= (2 + LTK_SWEEP_PREFETCH) * FUSED_UNROLL * TILE * sizeof(double),
              "look-ahead rows must fit the slack");

// The register look-ahead is one block (FUSED_UNROLL rows) deep: more would cost registers the kernel does not
// have.  A prefetch into L1 one block further ahead costs none: the register loads then hit L1.
__device__ __forceinline__ void prefetch_l1(const double* p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
constexpr unsigned FULL_MASK = 0xffffffffu;

struct FusedArgs {
    const double* kap;  // [n][tile-blocked] rotated curvature
    double* stage;      // [n][tile-blocked] parking array (v_acc of the first half, v_dec of the second)
    const int* rot;
    const double* len;
    double* lap;        // [B]
    double* vacc_d;     // optional dumps, [n][tile-blocked]
    double* vdec_d;
    double* vmin_d;
    const int4* lut;    // engine cell table (EngineLut), device memory; nullptr when ENG != 0
    int ns;
    long long B, Bp;
    long long first, last;  // candidates [first, last) of the padded population are this launch's (multiples of 32)
    TopkFuse tk;            // tk.k > 0: the CTAs select the population's k best in their epilogue (ltk_topk_fused.cuh)
};

// ---- engine map by cell table ---------------------------------------------------------------------
// Cell c = clamp((hi32(v) >> shift) - base, 0, cells-1) holds at most one node of the map: record
// (thr.lo, thr.hi, j0, -) with j = j0 + (bits(v) >= thr) = #{m : v >= map_v[m]}  (np.interp's segment,
// vehicle.py:25-27).  Cell 0 collects everything below the first node's cell, the last cell everything
// above the last node's cell.  Integer comparisons on the bit patterns are exact for v >= 0.
struct __align__(16) EngineSeg {
    double s, b, f, pad;  // slope, left abscissa, left ordinate of extended segment j (see VehDev)
};
struct FusedShared {
    EngineSeg seg[LTK_MAX_ENGINE_MAP + 1];
    int4 cell[LTK_LUT_MAX_CELLS];
};

template <int ENG>
__device__ __forceinline__ double engine_fast(const VehDev& V, const FusedShared& S, double x)
{
    int j;
    if (ENG == 0) {
        int c = (__double2hiint(x) >> V.lut_shift) - V.lut_base;
        c = max(c, 0);
        c = min(c, V.lut_top);
        const int4 r = S.cell[c];
        const long long thr = (long long)(((unsigned long long)(unsigned)r.y << 32) | (unsigned)r.x);
        j = r.z + ((__double_as_longlong(x) >= thr) ? 1 : 0);
    } else {
        const long long xi = __double_as_longlong(x);
        j = 0;
#pragma unroll
        for (int m = 0; m < ENG; ++m) j += (xi >= V.thr[m]) ? 1 : 0;
    }
    const EngineSeg g = S.seg[j];
    return g.s * (x - g.b) + g.f;
}

// curvature usable by the unguarded sequences AND small enough a local limit that the engine force
// stays non-negative (VehDev::k_lo_hi / k_span_hi are chosen on the host)
__device__ __forceinline__ bool kappa_regular(const VehDev& V, double k)
{
    return (unsigned)(__double2hiint(k) - V.k_lo_hi) < V.k_span_hi;
}

// velocity.py:44-50 as sqrt(min(.)) -- see the header comment.  wl = mu g / k of the row being entered.
template <int KIND, int ENG>
__device__ __forceinline__ double forward_fast(const VehDev& V, const FusedShared& S, double v_prev,
                                               double k_prev, double wl, double ds)
{
    double w = v_prev * v_prev;
    double tr = traction_from<false>(V, lateral_force<KIND>(V, v_prev, w, k_prev));
    double en = (KIND == 0) ? engine_fast<ENG>(V, S, v_prev) : V.e0 - V.cr2 * w;
    double force = lt_nonneg<false>(en, tr) ? en : tr;
    double accel2 = div_by_const<false>(force, V.half_mass, V.inv_half_mass);
    double wlim = w + accel2 * ds;
    return dsqrt<false>(lt_nonneg<false>(wlim, wl) ? wlim : wl);
}

// velocity.py:68-73, same transformation
template <int KIND>
__device__ __forceinline__ double backward_fast(const VehDev& V, double v_next, double k_next, double wl, double ds)
{
    double w = v_next * v_next;
    double tr = traction_from<false>(V, lateral_force<KIND>(V, v_next, w, k_next));
    double decel2 = div_by_const<false>(tr, V.half_mass, V.inv_half_mass);
    double wlim = w + decel2 * ds;
    return dsqrt<false>(lt_nonneg<false>(wlim, wl) ? wlim : wl);
}

// state of the two chains of one candidate
struct Chains {
    double vf, kf, vb, kb;  // velocity and curvature of the row each chain just left
    double ds_f;            // np.diff(s) of the interval the forward chain crosses next
    double lap_f, lap_b;
    int qf, qb;             // clocks: see GridClock (forward: current sample; backward: interval to enter)
    double sf, sb;
    double step, L;
    int n;
};

template <bool WRAP>
__device__ __forceinline__ double clock_advance(Chains& c)
{
    int k1 = c.qf + 1;
    double s1 = (double)k1 * c.step;
    if (WRAP) {
        bool wrap = (k1 == c.n);
        s1 = wrap ? c.L : s1;
        double ds = s1 - c.sf;
        c.qf = wrap ? 0 : k1;
        c.sf = wrap ? 0.0 : s1;
        return ds;
    }
    double ds = s1 - c.sf;
    c.qf = k1;
    c.sf = s1;
    return ds;
}
template <bool WRAP>
__device__ __forceinline__ double clock_retreat(Chains& c)
{
    double s_lo = (double)c.qb * c.step;
    double ds = c.sb - s_lo;
    if (WRAP) {
        bool wrap = (c.qb == 0);
        c.sb = wrap ? c.L : s_lo;
        c.qb = wrap ? c.n - 1 : c.qb - 1;
        return ds;
    }
    c.sb = s_lo;
    c.qb = c.qb - 1;
    return ds;
}

template <int KIND, int ENG>
__device__ __forceinline__ void step_pair_fast(const VehDev& V, const FusedShared& S, double vf, double kf, double wlf,
                                               double ds_f, double vb, double kb, double wlb, double ds_b,
                                               double& va, double& vd)
{
    const double wf = vf * vf, wb = vb * vb;
    const double en = (KIND == 0) ? engine_fast<ENG>(V, S, vf) : V.e0 - V.cr2 * wf;  // forward chain only
    const double lf = lateral_force<KIND>(V, vf, wf, kf), lb = lateral_force<KIND>(V, vb, wb, kb);
    const double xf = V.f_max_sq - lf * lf, xb = V.f_max_sq - lb * lb;  // vehicle.py:35
    double tf, tb;
    dsqrt_pair(xf, xb, tf, tb);
    tf = le_nonneg<false>(V.f_max, lf) ? 0.0 : tf;                       // vehicle.py:33-34
    tb = le_nonneg<false>(V.f_max, lb) ? 0.0 : tb;
    const double force = lt_nonneg<false>(en, tf) ? en : tf;
    const double af = div_by_const<false>(force, V.half_mass, V.inv_half_mass);
    const double ab = div_by_const<false>(tb, V.half_mass, V.inv_half_mass);
    const double limf = wf + af * ds_f, limb = wb + ab * ds_b;
    const double nf = lt_nonneg<false>(limf, wlf) ? limf : wlf, nb = lt_nonneg<false>(limb, wlb) ? limb : wlb;
    dsqrt_pair(nf, nb, va, vd);
}

// U row pairs on the regular path.  PHASE 1 parks, PHASE 2 meets the parked values and accumulates.
template <int KIND, int ENG, int PHASE, bool WRAP>
__device__ __forceinline__ void fused_block(const VehDev& V, const FusedShared& S, Chains& c,
                                            const double (&fc)[FUSED_UNROLL], const double (&bc)[FUSED_UNROLL],
                                            const double (&fo)[FUSED_UNROLL], const double (&bo)[FUSED_UNROLL],
                                            double* sfp, double* sbp)
{
    constexpr int U = FUSED_UNROLL;
    constexpr size_t P = TILE;
    double wlf[U], wlb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {  // independent of the recurrences: issued ahead of them
        wlf[u] = ddiv<false>(V.mu_g, fc[u]);
        wlb[u] = ddiv<false>(V.mu_g, bc[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const double ds_b = clock_retreat<WRAP>(c);
        double va, vd;
        step_pair_fast<KIND, ENG>(V, S, c.vf, c.kf, wlf[u], c.ds_f, c.vb, c.kb, wlb[u], ds_b, va, vd);
        c.ds_f = clock_advance<WRAP>(c);
        if (PHASE == 1) {
            sfp[(size_t)u * P] = va;
            *(sbp - (size_t)u * P) = vd;
        } else {
            double v1 = lt_nonneg<false>(va, fo[u]) ? va : fo[u];  // velocity.py:26
            double v2 = lt_nonneg<false>(bo[u], vd) ? bo[u] : vd;
            c.lap_f = c.lap_f + ddiv<false>(c.ds_f, v1);  // tbn.py:53
            c.lap_b = c.lap_b + ddiv<false>(ds_b, v2);
        }
        c.vf = va; c.kf = fc[u];
        c.vb = vd; c.kb = bc[u];
    }
}

// the two sweeps of candidate b (every lane of the warp is inside the padded population); returns the lap time
template <int KIND, int ENG>
__device__ __forceinline__ double k23_candidate(const FusedArgs& a, const VehDev& V, const FusedShared& S,
                                                const EngineTable& T, const long long b)
{
    constexpr int U = FUSED_UNROLL;
    constexpr size_t P = TILE;  // row pitch in doubles
    constexpr int NPAD = (ENG == 16) ? 16 : 8;  // comparison count of the library-operator path
    const int n = a.ns - 1;
    const size_t base = tile_base(b, n);
    const int p = a.rot[b];
    const bool dump = a.vdec_d != nullptr;

    Chains c;
    c.L = a.len[b];
    c.step = c.L / (double)(a.ns - 1);
    c.n = n;
    GridClock cf, cb;  // forward / backward position on the np.linspace grid
    cf.L = cb.L = c.L; cf.step = cb.step = c.step; cf.n = cb.n = n;
    cf.k = p; cf.s_k = (double)p * c.step;
    if (p == 0) { cb.k = n - 1; cb.s_k = c.L; } else { cb.k = p - 1; cb.s_k = (double)p * c.step; }

    // row 0 = the slowest sample: both chains start from v_local there (velocity.py:34-36, :58-61)
    const double k0 = a.kap[base];
    const double v0 = sqrt(V.mu_g / k0);
    c.vf = v0; c.kf = k0;
    c.vb = v0; c.kb = k0;
    c.ds_f = cf.advance();
    const double term0 = c.ds_f / v0;
    if (dump) { a.vacc_d[base] = v0; a.vdec_d[base] = v0; a.vmin_d[base] = v0; }
    c.qf = cf.k; c.sf = cf.s_k; c.qb = cb.k; c.sb = cb.s_k;
    c.lap_f = 0.0; c.lap_b = 0.0;

    const int rows = n - 1;           // rows 1 .. n-1
    const int h = rows / 2;           // steps per phase
    const bool has_mid = (rows & 1);  // middle row h+1 when the row count is odd

    const double* kfp = a.kap + base + P;                      // forward cursor: row 1 upwards
    const double* kbp = a.kap + base + (size_t)(n - 1) * P;    // backward cursor: row n-1 downwards
    double* sfp = a.stage + base + P;
    double* sbp = a.stage + base + (size_t)(n - 1) * P;
#define LTK_RF ((size_t)(kfp - a.kap))  // the cursors as element offsets (dump arrays share the layout)
#define LTK_RB ((size_t)(kbp - a.kap))

    // ---- generic single step (library operators, reference branch structure); used for tails, the
    //      middle row, irregular blocks and dumps ----------------------------------------------------
    auto step_safe = [&](int phase) {
        double va, vd;
        {
            double kc = *kfp;
            va = forward_step<KIND, NPAD, true>(V, T, c.vf, c.kf, local_limit<true>(V, kc), c.ds_f);
            c.vf = va; c.kf = kc;
            c.ds_f = clock_advance<true>(c);
        }
        double ds_b;
        {
            double kc = *kbp;
            ds_b = clock_retreat<true>(c);
            vd = backward_step<KIND, true>(V, c.vb, c.kb, local_limit<true>(V, kc), ds_b);
            c.vb = vd; c.kb = kc;
        }
        if (phase == 1) {
            *sfp = va;
            *sbp = vd;
            if (dump) { a.vacc_d[LTK_RF] = va; a.vdec_d[LTK_RB] = vd; }
        } else {
            double o = *sfp;  // v_dec parked by the backward chain
            double v = (va < o) ? va : o;
            c.lap_f = c.lap_f + c.ds_f / v;
            if (dump) { a.vacc_d[LTK_RF] = va; a.vmin_d[LTK_RF] = v; }
            o = *sbp;  // v_acc parked by the forward chain
            v = (o < vd) ? o : vd;
            c.lap_b = c.lap_b + ds_b / v;
            if (dump) { a.vdec_d[LTK_RB] = vd; a.vmin_d[LTK_RB] = v; }
        }
        kfp += P; sfp += P;
        kbp -= P; sbp -= P;
    };

    // The regular path keeps the chain state regular (v^2 = min(w + a ds, mu g / k) with k inside the window
    // and a >= 0), so the state is only re-examined after steps on the library-operator path.
    auto chains_regular = [&]() {
        return is_regular(c.vf) && kappa_regular(V, c.kf) && is_regular(c.vb) && kappa_regular(V, c.kb);
    };
    bool state_ok = chains_regular();

    // ---- phase 1 -----------------------------------------------------------------------------------
    int t = 0;  // steps done in this phase
    if (!dump) {
        double fc[U], fn[U], bc[U], bn[U];
        if (SWEEP_LA) {
#pragma unroll
            for (int u = 0; u < U; ++u) {  // look-ahead loads are unconditional: see SWEEP_SLACK
                fc[u] = kfp[(size_t)u * P];
                bc[u] = *(kbp - (size_t)u * P);
            }
        } else if (LTK_SWEEP_PREFETCH) {  // the first blocks' lines
#pragma unroll
            for (int u = 0; u < LTK_SWEEP_PREFETCH * U; ++u) {
                prefetch_l1(kfp + (size_t)u * P);
                prefetch_l1(kbp - (size_t)u * P);
            }
        }
        for (; t + U <= h; t += U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (SWEEP_LA) {
                    fn[u] = kfp[(size_t)(U + u) * P];
                    bn[u] = *(kbp - (size_t)(U + u) * P);
                } else {
                    fc[u] = kfp[(size_t)u * P];
                    bc[u] = *(kbp - (size_t)u * P);
                }
                if (LTK_SWEEP_PREFETCH) {
                    prefetch_l1(kfp + (size_t)((SWEEP_LA + LTK_SWEEP_PREFETCH) * U + u) * P);
                    prefetch_l1(kbp - (size_t)((SWEEP_LA + LTK_SWEEP_PREFETCH) * U + u) * P);
                }
            }
            bool regular = state_ok;
#pragma unroll
            for (int u = 0; u < U; ++u) regular = regular && kappa_regular(V, fc[u]) && kappa_regular(V, bc[u]);
            const bool nowrap = (c.qf + U < n) && (c.qb >= U);
            const bool all_regular = __all_sync(FULL_MASK, regular);
            if (all_regular) {
                if (__all_sync(FULL_MASK, nowrap)) fused_block<KIND, ENG, 1, false>(V, S, c, fc, bc, fc, bc, sfp, sbp);
                else fused_block<KIND, ENG, 1, true>(V, S, c, fc, bc, fc, bc, sfp, sbp);
                kfp += (size_t)U * P; sfp += (size_t)U * P;
                kbp -= (size_t)U * P; sbp -= (size_t)U * P;
            } else {  // zero / inf / nan curvature somewhere in this block of this warp
#pragma unroll 1
                for (int u = 0; u < U; ++u) step_safe(1);
                state_ok = chains_regular();
            }
            if (SWEEP_LA) {
#pragma unroll
                for (int u = 0; u < U; ++u) { fc[u] = fn[u]; bc[u] = bn[u]; }
            }
        }
    }
#pragma unroll 1
    for (; t < h; ++t) step_safe(1);

    // ---- middle row (odd row count): both chains arrive at row h+1 ------------------------------------
    double term_mid = 0.0;
    if (has_mid) {
        double kc = *kfp;
        double vl = local_limit<true>(V, kc);
        double va = forward_step<KIND, NPAD, true>(V, T, c.vf, c.kf, vl, c.ds_f);
        c.ds_f = clock_advance<true>(c);
        double ds_b = clock_retreat<true>(c);
        double vd = backward_step<KIND, true>(V, c.vb, c.kb, vl, ds_b);
        double v = (va < vd) ? va : vd;
        term_mid = ds_b / v;
        if (dump) { a.vacc_d[LTK_RF] = va; a.vdec_d[LTK_RF] = vd; a.vmin_d[LTK_RF] = v; }
        c.vf = va; c.kf = kc; c.vb = vd; c.kb = kc;
        kfp += P; sfp += P;
        kbp -= P; sbp -= P;
    }

    // ---- phase 2 -----------------------------------------------------------------------------------
    state_ok = chains_regular();  // the middle row ran on the library-operator path
    t = 0;
    if (!dump) {
        double fc[U], fn[U], bc[U], bn[U], fo[U], fon[U], bo[U], bon[U];
        if (SWEEP_LA) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                fc[u] = kfp[(size_t)u * P];
                fo[u] = sfp[(size_t)u * P];
                bc[u] = *(kbp - (size_t)u * P);
                bo[u] = *(sbp - (size_t)u * P);
            }
        } else if (LTK_SWEEP_PREFETCH) {
#pragma unroll
            for (int u = 0; u < LTK_SWEEP_PREFETCH * U; ++u) {
                prefetch_l1(kfp + (size_t)u * P);
                prefetch_l1(sfp + (size_t)u * P);
                prefetch_l1(kbp - (size_t)u * P);
                prefetch_l1(sbp - (size_t)u * P);
            }
        }
        for (; t + U <= h; t += U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (SWEEP_LA) {
                    fn[u] = kfp[(size_t)(U + u) * P];
                    fon[u] = sfp[(size_t)(U + u) * P];
                    bn[u] = *(kbp - (size_t)(U + u) * P);
                    bon[u] = *(sbp - (size_t)(U + u) * P);
                } else {
                    fc[u] = kfp[(size_t)u * P];
                    fo[u] = sfp[(size_t)u * P];
                    bc[u] = *(kbp - (size_t)u * P);
                    bo[u] = *(sbp - (size_t)u * P);
                }
                if (LTK_SWEEP_PREFETCH) {
                    prefetch_l1(kfp + (size_t)((SWEEP_LA + LTK_SWEEP_PREFETCH) * U + u) * P);
                    prefetch_l1(sfp + (size_t)((SWEEP_LA + LTK_SWEEP_PREFETCH) * U + u) * P);
                    prefetch_l1(kbp - (size_t)((SWEEP_LA + LTK_SWEEP_PREFETCH) * U + u) * P);
                    prefetch_l1(sbp - (size_t)((SWEEP_LA + LTK_SWEEP_PREFETCH) * U + u) * P);
                }
            }
            bool regular = state_ok;
#pragma unroll
            for (int u = 0; u < U; ++u)
                regular = regular && kappa_regular(V, fc[u]) && kappa_regular(V, bc[u]) && is_regular(fo[u]) && is_regular(bo[u]);
            const bool nowrap = (c.qf + U < n) && (c.qb >= U);
            const bool all_regular = __all_sync(FULL_MASK, regular);
            if (all_regular) {
                if (__all_sync(FULL_MASK, nowrap)) fused_block<KIND, ENG, 2, false>(V, S, c, fc, bc, fo, bo, sfp, sbp);
                else fused_block<KIND, ENG, 2, true>(V, S, c, fc, bc, fo, bo, sfp, sbp);
                kfp += (size_t)U * P; sfp += (size_t)U * P;
                kbp -= (size_t)U * P; sbp -= (size_t)U * P;
            } else {  // zero / inf / nan curvature somewhere in this block of this warp
#pragma unroll 1
                for (int u = 0; u < U; ++u) step_safe(2);
                state_ok = chains_regular();
            }
            if (SWEEP_LA) {
#pragma unroll
                for (int u = 0; u < U; ++u) { fc[u] = fn[u]; bc[u] = bn[u]; fo[u] = fon[u]; bo[u] = bon[u]; }
            }
        }
    }
#pragma unroll 1
    for (; t < h; ++t) step_safe(2);

#undef LTK_RF
#undef LTK_RB
    double lap = c.lap_f + c.lap_b;
    if (has_mid) lap = lap + term_mid;
    return lap + term0;
}

template <int KIND, int ENG>
__global__ void __launch_bounds__(FUSED_THREADS, LTK_SWEEP_MINB) k23_sweep(FusedArgs a, VehDev V)
{
    __shared__ FusedShared S;
    __shared__ EngineTable T;  // library-operator path (tails, irregular blocks, dumps)
    if (KIND == 0) {
        load_engine_table(T, V, threadIdx.x, FUSED_THREADS);
        for (int i = threadIdx.x; i <= LTK_MAX_ENGINE_MAP; i += FUSED_THREADS) {
            S.seg[i].s = V.ext_s[i]; S.seg[i].b = V.ext_b[i]; S.seg[i].f = V.ext_f[i]; S.seg[i].pad = 0.0;
        }
        if (ENG == 0)
            for (int i = threadIdx.x; i <= V.lut_top; i += FUSED_THREADS) S.cell[i] = a.lut[i];
        __syncthreads();
    }
    const long long b = a.first + (long long)blockIdx.x * FUSED_THREADS + threadIdx.x;
    // a whole warp beyond this launch's candidates has nothing to sweep; lanes in [B, Bp) sweep the padding copies
    // K1 wrote (so that warp votes see a full warp)
    const bool live = b - (threadIdx.x & 31) < a.last;
    double lap = 0.0;
    if (live) {
        lap = k23_candidate<KIND, ENG>(a, V, S, T, b);
        if (b < a.B) a.lap[b] = lap;
    }
    if (a.tk.k > 0) topk_epilogue(a.tk, live && b < a.B, lap, b);
}

}  // namespace ltk
