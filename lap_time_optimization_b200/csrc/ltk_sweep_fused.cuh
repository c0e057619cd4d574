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
#pragma once

namespace ltk {

constexpr int FUSED_THREADS = 64;
constexpr int FUSED_UNROLL = 4;

struct FusedArgs {
    const double* kap;  // [n][tile-blocked] rotated curvature
    double* stage;      // [n][tile-blocked] parking array (v_acc of the first half, v_dec of the second)
    const int* rot;
    const double* len;
    double* lap;        // [B]
    double* vacc_d;     // optional dumps, [n][tile-blocked]
    double* vdec_d;
    double* vmin_d;
    int ns;
    long long B;
};

template <int KIND, int NPAD>
__global__ void __launch_bounds__(FUSED_THREADS, 8) k23_sweep(FusedArgs a, VehDev V)
{
    constexpr int U = FUSED_UNROLL;
    constexpr size_t P = TILE;  // row pitch in doubles
    __shared__ EngineTable T;
    if (KIND == 0) {
        load_engine_table(T, V, threadIdx.x, FUSED_THREADS);
        __syncthreads();
    }
    const long long b = (long long)blockIdx.x * FUSED_THREADS + threadIdx.x;
    if (b >= a.B) return;
    const int n = a.ns - 1;
    const size_t base = tile_base(b, n);
    const int p = a.rot[b];
    const bool dump = a.vdec_d != nullptr;

    const double L = a.len[b];
    const double step = L / (double)(a.ns - 1);
    GridClock cf, cb;  // forward / backward position on the np.linspace grid
    cf.L = cb.L = L; cf.step = cb.step = step; cf.n = cb.n = n;
    cf.k = p; cf.s_k = (double)p * step;
    if (p == 0) { cb.k = n - 1; cb.s_k = L; } else { cb.k = p - 1; cb.s_k = (double)p * step; }

    // row 0 = the slowest sample: both chains start from v_local there (velocity.py:34-36, :58-61)
    const double k0 = a.kap[base];
    const double v0 = sqrt(V.mu_g / k0);
    double vf = v0, kf = k0;  // forward state: v_acc and curvature of the row just left
    double vb = v0, kb = k0;  // backward state
    double ds_f = cf.advance();  // np.diff(s) of the interval the forward chain crosses next
    const double term0 = ds_f / v0;
    if (dump) { a.vacc_d[base] = v0; a.vdec_d[base] = v0; a.vmin_d[base] = v0; }

    const int rows = n - 1;           // rows 1 .. n-1
    const int h = rows / 2;           // steps per phase
    const bool has_mid = (rows & 1);  // middle row h+1 when the row count is odd

    const double* kfp = a.kap + base + P;                      // forward cursor: row 1 upwards
    const double* kbp = a.kap + base + (size_t)(n - 1) * P;    // backward cursor: row n-1 downwards
    double* sfp = a.stage + base + P;
    double* sbp = a.stage + base + (size_t)(n - 1) * P;
    size_t rf = base + P, rb = base + (size_t)(n - 1) * P;     // same cursors as offsets (dumps)

    double lap_f = 0.0, lap_b = 0.0;

    // ---- generic single step (library operators); used for tails, the middle row, irregular blocks ----
    auto step_safe = [&](bool do_f, bool do_b, int phase) {
        double va = 0.0, vd = 0.0;
        if (do_f) {
            double kc = *kfp;
            va = forward_step<KIND, NPAD, true>(V, T, vf, kf, local_limit<true>(V, kc), ds_f);
            vf = va; kf = kc;
            ds_f = cf.advance();
        }
        double ds_b = 0.0;
        if (do_b) {
            double kc = *kbp;
            ds_b = cb.retreat();
            vd = backward_step<KIND, true>(V, vb, kb, local_limit<true>(V, kc), ds_b);
            vb = vd; kb = kc;
        }
        if (phase == 1) {
            if (do_f) *sfp = va;
            if (do_b) *sbp = vd;
            if (dump) { if (do_f) a.vacc_d[rf] = va; if (do_b) a.vdec_d[rb] = vd; }
        } else if (phase == 2) {
            if (do_f) {
                double o = *sfp;  // v_dec parked by the backward chain
                double v = (va < o) ? va : o;
                lap_f = lap_f + ds_f / v;
                if (dump) { a.vacc_d[rf] = va; a.vmin_d[rf] = v; }
            }
            if (do_b) {
                double o = *sbp;  // v_acc parked by the forward chain
                double v = (o < vd) ? o : vd;
                lap_b = lap_b + ds_b / v;
                if (dump) { a.vdec_d[rb] = vd; a.vmin_d[rb] = v; }
            }
        }
        if (do_f) { kfp += P; sfp += P; rf += P; }
        if (do_b) { kbp -= P; sbp -= P; rb -= P; }
    };

    // ---- phase 1 -----------------------------------------------------------------------------------
    int t = 0;  // steps done in this phase
    if (!dump) {
        double fc[U], fn[U], bc[U], bn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            bool in = (u < h);
            fc[u] = in ? kfp[(size_t)u * P] : 1.0;
            bc[u] = in ? *(kbp - (size_t)u * P) : 1.0;
        }
        for (; t + U <= h; t += U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                bool in = (t + U + u < h);
                fn[u] = in ? kfp[(size_t)(U + u) * P] : 1.0;
                bn[u] = in ? *(kbp - (size_t)(U + u) * P) : 1.0;
            }
            bool regular = is_regular(vf) && is_regular(kf) && is_regular(vb) && is_regular(kb);
#pragma unroll
            for (int u = 0; u < U; ++u) regular = regular && is_regular(fc[u]) && is_regular(bc[u]);
            if (regular) {
                double vlf[U], vlb[U];
#pragma unroll
                for (int u = 0; u < U; ++u) { vlf[u] = local_limit<false>(V, fc[u]); vlb[u] = local_limit<false>(V, bc[u]); }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    double va = forward_step<KIND, NPAD, false>(V, T, vf, kf, vlf[u], ds_f);
                    ds_f = cf.advance();
                    double ds_b = cb.retreat();
                    double vd = backward_step<KIND, false>(V, vb, kb, vlb[u], ds_b);
                    sfp[(size_t)u * P] = va;
                    *(sbp - (size_t)u * P) = vd;
                    vf = va; kf = fc[u];
                    vb = vd; kb = bc[u];
                }
                kfp += (size_t)U * P; sfp += (size_t)U * P; rf += (size_t)U * P;
                kbp -= (size_t)U * P; sbp -= (size_t)U * P; rb -= (size_t)U * P;
            } else {
#pragma unroll 1
                for (int u = 0; u < U; ++u) step_safe(true, true, 1);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { fc[u] = fn[u]; bc[u] = bn[u]; }
        }
    }
#pragma unroll 1
    for (; t < h; ++t) step_safe(true, true, 1);

    // ---- middle row (odd row count): both chains arrive at row h+1 ------------------------------------
    double term_mid = 0.0;
    if (has_mid) {
        double kc = *kfp;
        double vl = local_limit<true>(V, kc);
        double va = forward_step<KIND, NPAD, true>(V, T, vf, kf, vl, ds_f);
        ds_f = cf.advance();
        double ds_b = cb.retreat();
        double vd = backward_step<KIND, true>(V, vb, kb, vl, ds_b);
        double v = (va < vd) ? va : vd;
        term_mid = ds_b / v;
        if (dump) { a.vacc_d[rf] = va; a.vdec_d[rf] = vd; a.vmin_d[rf] = v; }
        vf = va; kf = kc; vb = vd; kb = kc;
        kfp += P; sfp += P; rf += P;
        kbp -= P; sbp -= P; rb -= P;
    }

    // ---- phase 2 -----------------------------------------------------------------------------------
    t = 0;
    if (!dump) {
        double fc[U], fn[U], bc[U], bn[U], fo[U], fon[U], bo[U], bon[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            bool in = (u < h);
            fc[u] = in ? kfp[(size_t)u * P] : 1.0;
            fo[u] = in ? sfp[(size_t)u * P] : 1.0;
            bc[u] = in ? *(kbp - (size_t)u * P) : 1.0;
            bo[u] = in ? *(sbp - (size_t)u * P) : 1.0;
        }
        for (; t + U <= h; t += U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                bool in = (t + U + u < h);
                fn[u] = in ? kfp[(size_t)(U + u) * P] : 1.0;
                fon[u] = in ? sfp[(size_t)(U + u) * P] : 1.0;
                bn[u] = in ? *(kbp - (size_t)(U + u) * P) : 1.0;
                bon[u] = in ? *(sbp - (size_t)(U + u) * P) : 1.0;
            }
            bool regular = is_regular(vf) && is_regular(kf) && is_regular(vb) && is_regular(kb);
#pragma unroll
            for (int u = 0; u < U; ++u)
                regular = regular && is_regular(fc[u]) && is_regular(bc[u]) && is_regular(fo[u]) && is_regular(bo[u]);
            if (regular) {
                double vlf[U], vlb[U];
#pragma unroll
                for (int u = 0; u < U; ++u) { vlf[u] = local_limit<false>(V, fc[u]); vlb[u] = local_limit<false>(V, bc[u]); }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    double va = forward_step<KIND, NPAD, false>(V, T, vf, kf, vlf[u], ds_f);
                    ds_f = cf.advance();
                    double ds_b = cb.retreat();
                    double vd = backward_step<KIND, false>(V, vb, kb, vlb[u], ds_b);
                    double v1 = lt_nonneg<false>(va, fo[u]) ? va : fo[u];  // velocity.py:26
                    double v2 = lt_nonneg<false>(bo[u], vd) ? bo[u] : vd;
                    lap_f = lap_f + ddiv<false>(ds_f, v1);  // tbn.py:53
                    lap_b = lap_b + ddiv<false>(ds_b, v2);
                    vf = va; kf = fc[u];
                    vb = vd; kb = bc[u];
                }
                kfp += (size_t)U * P; sfp += (size_t)U * P; rf += (size_t)U * P;
                kbp -= (size_t)U * P; sbp -= (size_t)U * P; rb -= (size_t)U * P;
            } else {
#pragma unroll 1
                for (int u = 0; u < U; ++u) step_safe(true, true, 2);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { fc[u] = fn[u]; bc[u] = bn[u]; fo[u] = fon[u]; bo[u] = bon[u]; }
        }
    }
#pragma unroll 1
    for (; t < h; ++t) step_safe(true, true, 2);

    double lap = lap_f + lap_b;
    if (has_mid) lap = lap + term_mid;
    a.lap[b] = lap + term0;
}

}  // namespace ltk
