// ltk_sweep_roles.cuh -- K23r: the two sweeps of a candidate on two lanes of two different warps.
//
// Same data flow as the fused kernel (ltk_sweep_fused.cuh): the forward chain walks rotated rows
// 1 .. n-1, the backward chain rows n-1 .. 1; in the first half each parks its values, in the second
// half it meets what the other one parked, takes the minimum (velocity.py:26) and accumulates ds/v
// (tbn.py:51-54).  Same arithmetic, same summation order, same HBM traffic -- bit-identical results.
//
// What changes is who runs the chains.  Measured on B200 (tools/ubench): a dependent FP64 instruction
// issues 8.35 cycles after its producer, the pipe accepts one warp instruction every 2 cycles, and a
// warp issues in order.  With both chains in ONE thread the interleaving of the two dependency chains
// is whatever the compiler's static schedule achieves; at the 3.5 warps per scheduler of a
// 65,536-candidate population the FP64 pipe sat at 51 % with "wait" (fixed-latency dependency) the
// top stall.  Here a CTA is two warps over the same 32 candidates: warp 0 runs the forward chains,
// warp 1 the backward chains.  That is 7 single-chain warps per scheduler instead of 3.5 double-chain
// ones; the hardware scheduler interleaves them instruction by instruction, and a thread needs half
// the registers (14 CTAs of 64 threads per SM, the whole population resident in one wave).
// The two warps meet at one __syncthreads() between the halves (the parked rows cross over there).
//
// Measured (B200, Buckmore/TBR18): equal to the fused kernel at 65,536 candidates (0.56 ms), 17 % faster
// at <= 28,416 (latency-bound: 0.217 vs 0.261 ms at 8,192, 0.265 vs 0.290 at 24,576), 13 % slower on multi-wave populations (a
// forward step costs ~19 % more FP64-pipe time than a backward one, so the backward warp idles at the
// barrier and at the end while holding its registers).  Swapping the chains between the warps at half
// time was tried and does not help: the barrier makes each half take max(forward, backward) either way.
// The host picks this kernel for small batches only (ltk_api.cu).
#pragma once
#include <type_traits>

namespace ltk {

constexpr int ROLES_THREADS = 64;   // warp 0: forward chains, warp 1: backward chains of 32 candidates
constexpr int ROLES_UNROLL = 4;

// one chain's state
struct Chain {
    double v, k;      // velocity and curvature of the row this chain just left
    double ds;        // forward only: np.diff(s) of the interval crossed next
    double lap;
    int q;            // clock, see GridClock (forward: current sample; backward: interval to enter)
    double s;
    double step, L;
    int n;
};

template <bool WRAP>
__device__ __forceinline__ double chain_advance(Chain& c)  // forward clock
{
    int k1 = c.q + 1;
    double s1 = (double)k1 * c.step;
    if (WRAP) {
        bool wrap = (k1 == c.n);
        s1 = wrap ? c.L : s1;
        double ds = s1 - c.s;
        c.q = wrap ? 0 : k1;
        c.s = wrap ? 0.0 : s1;
        return ds;
    }
    double ds = s1 - c.s;
    c.q = k1;
    c.s = s1;
    return ds;
}
template <bool WRAP>
__device__ __forceinline__ double chain_retreat(Chain& c)  // backward clock
{
    double s_lo = (double)c.q * c.step;
    double ds = c.s - s_lo;
    if (WRAP) {
        bool wrap = (c.q == 0);
        c.s = wrap ? c.L : s_lo;
        c.q = wrap ? c.n - 1 : c.q - 1;
        return ds;
    }
    c.s = s_lo;
    c.q = c.q - 1;
    return ds;
}

// U rows of one chain on the regular path.  ROLE 0 forward / 1 backward; PHASE 1 parks, PHASE 2 meets
// the parked values `oc` and accumulates.  `sp` points at the staging row of step 0; rows advance by
// +TILE (forward) or -TILE (backward) doubles.
template <int KIND, int ENG, int ROLE, int PHASE, bool WRAP>
__device__ __forceinline__ void chain_block(const VehDev& V, const FusedShared& S, Chain& c,
                                            const double (&kc)[ROLES_UNROLL], const double (&oc)[ROLES_UNROLL],
                                            double* sp)
{
    constexpr int U = ROLES_UNROLL;
    constexpr ptrdiff_t D = (ROLE == 0) ? (ptrdiff_t)TILE : -(ptrdiff_t)TILE;
    double wl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) wl[u] = ddiv<false>(V.mu_g, kc[u]);  // off the recurrence: issued ahead
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (ROLE == 0) {
            double va = forward_fast<KIND, ENG>(V, S, c.v, c.k, wl[u], c.ds);
            c.ds = chain_advance<WRAP>(c);
            if (PHASE == 1) {
                sp[u * D] = va;
            } else {
                double v1 = lt_nonneg<false>(va, oc[u]) ? va : oc[u];         // velocity.py:26
                c.lap = c.lap + ddiv<false>(c.ds, v1);         // tbn.py:53
            }
            c.v = va;
        } else {
            double ds = chain_retreat<WRAP>(c);
            double vd = backward_fast<KIND>(V, c.v, c.k, wl[u], ds);
            if (PHASE == 1) {
                sp[u * D] = vd;
            } else {
                double v2 = lt_nonneg<false>(oc[u], vd) ? oc[u] : vd;
                c.lap = c.lap + ddiv<false>(ds, v2);
            }
            c.v = vd;
        }
        c.k = kc[u];
    }
}

template <int KIND, int ENG>
__global__ void __launch_bounds__(ROLES_THREADS, 14) k23_roles(FusedArgs a, VehDev V)
{
    constexpr int U = ROLES_UNROLL;
    constexpr ptrdiff_t P = TILE;
    constexpr int NPAD = (ENG == 16) ? 16 : 8;  // comparison count of the library-operator path
    __shared__ FusedShared S;
    __shared__ EngineTable T;  // library-operator path (tails, irregular blocks, dumps)
    __shared__ double x_lap[32], x_mid[32];  // backward warp -> forward warp
    if (KIND == 0) {
        load_engine_table(T, V, threadIdx.x, ROLES_THREADS);
        for (int i = threadIdx.x; i <= LTK_MAX_ENGINE_MAP; i += ROLES_THREADS) {
            S.seg[i].s = V.ext_s[i]; S.seg[i].b = V.ext_b[i]; S.seg[i].f = V.ext_f[i]; S.seg[i].pad = 0.0;
        }
        if (ENG == 0)
            for (int i = threadIdx.x; i <= V.lut_top; i += ROLES_THREADS) S.cell[i] = a.lut[i];
        __syncthreads();
    }
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;  // warp 0 forward, warp 1 backward
    // one CTA per 32 candidates of the padded population (lanes in [B, Bp) sweep K1's padding copies)
    const long long b = a.first + (long long)blockIdx.x * 32 + lane;
    const int n = a.ns - 1;
    const size_t base = tile_base(b, n);
    const int p = a.rot[b];
    const bool dump = a.vdec_d != nullptr;

    Chain c;
    c.L = a.len[b];
    c.step = c.L / (double)(a.ns - 1);
    c.n = n;
    c.lap = 0.0;
    // row 0 = the slowest sample: both chains start from v_local there (velocity.py:34-36, :58-61)
    const double k0 = a.kap[base];
    const double v0 = sqrt(V.mu_g / k0);
    c.v = v0; c.k = k0;
    double term0 = 0.0;
    if (role == 0) {
        c.q = p; c.s = (double)p * c.step;
        c.ds = chain_advance<true>(c);
        term0 = c.ds / v0;
        if (dump) { a.vacc_d[base] = v0; a.vmin_d[base] = v0; }
    } else {
        if (p == 0) { c.q = n - 1; c.s = c.L; } else { c.q = p - 1; c.s = (double)p * c.step; }
        c.ds = 0.0;
        if (dump) a.vdec_d[base] = v0;
    }

    const int rows = n - 1;           // rows 1 .. n-1
    const int h = rows / 2;           // steps per phase
    const bool has_mid = (rows & 1);  // middle row h+1 when the row count is odd
    const ptrdiff_t D = (role == 0) ? P : -P;
    ptrdiff_t r = (ptrdiff_t)base + (ptrdiff_t)((role == 0) ? 1 : n - 1) * P;  // cursor (offset of the row entered next)

    // ---- one step with library operators and the reference's branch structure: tails, the middle
    //      row, irregular blocks, dumps.  phase 1 parks, phase 2 meets + accumulates, phase 3 = the
    //      forward chain parking the middle row, phase 4 = the backward chain meeting it --------------
    double term_mid = 0.0;
    auto step_safe = [&](int phase) {
        const double kc = a.kap[r];
        const double vl = local_limit<true>(V, kc);
        double vn, ds_here;  // new value; np.diff(s) of the interval STARTING at the row entered
        if (role == 0) {
            vn = forward_step<KIND, NPAD, true>(V, T, c.v, c.k, vl, c.ds);
            c.ds = chain_advance<true>(c);
            ds_here = c.ds;
        } else {
            ds_here = chain_retreat<true>(c);
            vn = backward_step<KIND, true>(V, c.v, c.k, vl, ds_here);
        }
        c.v = vn; c.k = kc;
        if (phase == 1 || phase == 3) {
            a.stage[r] = vn;
            if (dump) { if (role == 0) a.vacc_d[r] = vn; else a.vdec_d[r] = vn; }
        } else {
            const double o = a.stage[r];  // parked by the other chain
            const double v = (role == 0) ? ((vn < o) ? vn : o) : ((o < vn) ? o : vn);
            if (phase == 4) term_mid = ds_here / v;
            else c.lap = c.lap + ds_here / v;
            if (dump) {
                if (role == 0) a.vacc_d[r] = vn; else a.vdec_d[r] = vn;
                a.vmin_d[r] = v;
            }
        }
        r += D;
    };

    // ---- the block loop of one phase ------------------------------------------------------------------
    auto run_phase = [&](auto role_tag, auto phase_tag) {
        constexpr int ROLE = decltype(role_tag)::value;
        constexpr int PHASE = decltype(phase_tag)::value;
        constexpr ptrdiff_t D = (ROLE == 0) ? P : -P;
        int t = 0;
        if (!dump) {
            const double* kp = a.kap + r;
            double* sp = a.stage + r;
            double kc[U], kn[U], oc[U], on[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                bool in = (u < h);
                kc[u] = in ? kp[u * D] : 1.0;
                oc[u] = (PHASE == 2 && in) ? sp[u * D] : 1.0;
            }
            for (; t + U <= h; t += U) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    bool in = (t + U + u < h);
                    kn[u] = in ? kp[(U + u) * D] : 1.0;
                    on[u] = (PHASE == 2 && in) ? sp[(U + u) * D] : 1.0;
                }
                bool regular = is_regular(c.v) && kappa_regular(V, c.k);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    regular = regular && kappa_regular(V, kc[u]);
                    if (PHASE == 2) regular = regular && is_regular(oc[u]);
                }
                const bool nowrap = (ROLE == 0) ? (c.q + U < n) : (c.q >= U);
                if (__all_sync(FULL_MASK, regular)) {
                    if (__all_sync(FULL_MASK, nowrap)) chain_block<KIND, ENG, ROLE, PHASE, false>(V, S, c, kc, oc, sp);
                    else chain_block<KIND, ENG, ROLE, PHASE, true>(V, S, c, kc, oc, sp);
                    r += U * D;
                } else {  // zero / inf / nan curvature somewhere in this block of this warp
#pragma unroll 1
                    for (int u = 0; u < U; ++u) step_safe(PHASE);
                }
                kp += U * D; sp += U * D;
#pragma unroll
                for (int u = 0; u < U; ++u) { kc[u] = kn[u]; oc[u] = on[u]; }
            }
        }
#pragma unroll 1
        for (; t < h; ++t) step_safe(PHASE);
    };

    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    if (role == 0) {
        run_phase(I0{}, I1{});
        if (has_mid) step_safe(3);  // forward chain parks the middle row
    } else {
        run_phase(I1{}, I1{});
    }
    __syncthreads();  // half time: the parked rows cross over
    if (role == 0) {
        run_phase(I0{}, I2{});
    } else {
        if (has_mid) step_safe(4);  // backward chain meets the middle row
        run_phase(I1{}, I2{});
    }

    if (role == 1) { x_lap[lane] = c.lap; x_mid[lane] = term_mid; }
    __syncthreads();
    double lap = 0.0;
    if (role == 0) {
        lap = c.lap + x_lap[lane];
        if (has_mid) lap = lap + x_mid[lane];
        lap = lap + term0;
        if (b < a.B) a.lap[b] = lap;
    }
    if (a.tk.k > 0) topk_epilogue(a.tk, role == 0 && b < a.B, lap, b);
}

}  // namespace ltk
