// ltk_sweep_f32.cuh -- K23f: the optional fp32 variant of the velocity sweeps (BASELINE.json north_star:
// "lap times within 1e-9 relative in the fp64 kernels, 1e-4 for an optional fp32 variant").
//
// Same data flow as k23_sweep: one thread per candidate runs the forward chain (velocity.py:31-53) over
// rotated rows 1 .. n-1 and the backward chain (velocity.py:55-76) over rows n-1 .. 1 at once, parks
// each chain's first half, meets the other chain's parked values in the second half, takes the
// minimum (velocity.py:26) and accumulates ds / v (tbn.py:51-54).  What changes:
//   * the recurrences run in fp32 with the reference's own branch structure as selects; square roots,
//     reciprocal square roots and reciprocals are the hardware approximations (MUFU, ~2^-22 relative:
//     `sqrt.approx`, `rsqrt.approx`, `rcp.approx`) -- with correctly rounded sqrtf and division the
//     kernel was no faster than the fp64 one (0.94 ms against 0.53 ms, measured) and the 1e-4 budget is
//     four orders of magnitude above these errors;
//   * the curvature is still computed in fp64 by K1b, which also writes it rounded to fp32 for this
//     kernel; parked velocities are fp32: 8 n + 8 n bytes per candidate here instead of 32 n
//     (+ 4 n for K1b's second copy);
//   * np.diff(s) is taken as the constant step L / (ns - 1) (its fp64 values differ from it by 1e-13);
//   * the lap-time sum is accumulated in fp64.
// Measured error against the fp64 kernels: see tests/test_gpu_parity.py::test_fp32_sweep_variant.
#pragma once
#include <type_traits>

namespace ltk {

constexpr int F32_THREADS = 128;
constexpr int F32_UNROLL = 4;
#ifndef LTK_F32_PREFETCH
#define LTK_F32_PREFETCH 1
#endif
constexpr bool F32_PREFETCH = LTK_F32_PREFETCH != 0;  // L1 prefetch one block ahead of the register look-ahead

__device__ __forceinline__ void prefetch_l1_any(const void* p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

struct F32Args {
    const double* kap;   // [n][tile-blocked] rotated curvature (fp64, from K1b)
    const float* kap32;  // the same rounded to fp32 by K1b (K32 kernels read this one: half the bytes)
    float* stage;       // [n][tile-blocked] parking array, fp32 (aliases the fp64 staging array)
    const double* len;
    double* lap;        // [B]
    int ns;
    long long B, Bp;
};

struct VehF32 {
    int kind, n_map;
    float mass, mu_g, f_max, f_max_sq, e0, cr2;
    float thr[LTK_MAX_ENGINE_MAP];                                  // node abscissae, padded with +inf
    float ext_b[LTK_MAX_ENGINE_MAP + 1], ext_f[LTK_MAX_ENGINE_MAP + 1], ext_s[LTK_MAX_ENGINE_MAP + 1];
};

__device__ __forceinline__ float sqrt_fast(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_fast(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_fast(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int KIND, int NPAD>
__device__ __forceinline__ float engine_f32(const VehF32& V, const float (&sb)[LTK_MAX_ENGINE_MAP + 1],
                                            const float (&sf)[LTK_MAX_ENGINE_MAP + 1],
                                            const float (&ss)[LTK_MAX_ENGINE_MAP + 1], float v, float w)
{
    if (KIND == 1) return V.e0 - V.cr2 * w;  // vehicleMX5.py:21
    int j = 0;                               // np.interp segment (vehicle.py:25-27)
#pragma unroll
    for (int m = 0; m < NPAD; ++m) j += (v >= V.thr[m]) ? 1 : 0;
    return ss[j] * (v - sb[j]) + sf[j];
}

template <int KIND>
__device__ __forceinline__ float traction_f32(const VehF32& V, float v, float w, float k)
{
    const float fl = (KIND == 0) ? (V.mass * w) * k : ((V.mass * v) * v) * k;  // vehicle.py:31 / vehicleMX5.py:34
    return (V.f_max <= fl) ? 0.0f : sqrt_fast(V.f_max_sq - fl * fl);           // vehicle.py:33-35
}

// K32: the curvature comes from K1b's fp32 copy (same values as rounding the fp64 one here)
template <int KIND, int NPAD, bool K32>
__global__ void __launch_bounds__(F32_THREADS) k23_f32(F32Args a, VehF32 V)
{
    using KT_ = typename std::conditional<K32, float, double>::type;
    const KT_* kap = K32 ? reinterpret_cast<const KT_*>(a.kap32) : reinterpret_cast<const KT_*>(a.kap);
    constexpr int U = F32_UNROLL;
    constexpr size_t P = TILE;
    __shared__ float sb[LTK_MAX_ENGINE_MAP + 1], sf[LTK_MAX_ENGINE_MAP + 1], ss[LTK_MAX_ENGINE_MAP + 1];
    if (threadIdx.x <= LTK_MAX_ENGINE_MAP) {
        sb[threadIdx.x] = V.ext_b[threadIdx.x]; sf[threadIdx.x] = V.ext_f[threadIdx.x]; ss[threadIdx.x] = V.ext_s[threadIdx.x];
    }
    __syncthreads();
    const long long b = (long long)blockIdx.x * F32_THREADS + threadIdx.x;
    if (b >= a.B) return;
    const int n = a.ns - 1;
    const size_t base = tile_base(b, n);
    const float ds = (float)(a.len[b] / (double)(a.ns - 1));
    const float two_ds_m = 2.0f * ds / V.mass;   // 2 ds / m: v^2 + 2 (F / m) ds  (velocity.py:47-49)
    const float root_mug = sqrtf(V.mu_g);        // v_local = sqrt(mu g / k) = sqrt(mu g) * rsqrt(k)

    // row 0 = the slowest sample: both chains start from v_local there (velocity.py:34-36, :58-61)
    const float k0 = (float)kap[base];
    const float v0 = root_mug * rsqrt_fast(k0);
    float vf = v0, kf = k0, vb = v0, kb = k0;
    double lap = (double)(ds * rcp_fast(v0));

    auto fwd = [&](float kc) {  // velocity.py:44-50
        const float vl = root_mug * rsqrt_fast(kc);
        const float w = vf * vf;
        const float tr = traction_f32<KIND>(V, vf, w, kf);
        const float en = engine_f32<KIND, NPAD>(V, sb, sf, ss, vf, w);
        const float vlim = sqrt_fast(w + two_ds_m * fminf(en, tr));
        const float v = (vl > vf) ? fminf(vl, vlim) : vl;
        vf = v; kf = kc;
        return v;
    };
    auto bwd = [&](float kc) {  // velocity.py:68-73
        const float vl = root_mug * rsqrt_fast(kc);
        const float w = vb * vb;
        const float vlim = sqrt_fast(w + two_ds_m * traction_f32<KIND>(V, vb, w, kb));
        const float v = (vl > vb) ? fminf(vl, vlim) : vl;
        vb = v; kb = kc;
        return v;
    };

    const int rows = n - 1, h = rows / 2;
    const bool has_mid = rows & 1;
    const KT_* kfp = kap + base + P;
    const KT_* kbp = kap + base + (size_t)(n - 1) * P;
    // the fp32 staging array uses the fp64 array's element index (its second half stays unused)
    float* sfp = a.stage + base + P;
    float* sbp = a.stage + base + (size_t)(n - 1) * P;

    // ---- phase 1: park (loads run one block of U rows ahead of their use) -------------------------------
    int t = 0;
    {
        KT_ fc[U], bc[U], fn[U], bn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool in = (u < h);
            fc[u] = in ? kfp[(size_t)u * P] : (KT_)1;
            bc[u] = in ? *(kbp - (size_t)u * P) : (KT_)1;
        }
        for (; t + U <= h; t += U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool in = (t + U + u < h);
                fn[u] = in ? kfp[(size_t)(U + u) * P] : (KT_)1;
                bn[u] = in ? *(kbp - (size_t)(U + u) * P) : (KT_)1;
            }
            if (F32_PREFETCH) {  // one 128-byte line holds two fp32 rows of a tile
#pragma unroll
                for (int u = 0; u < U; u += 2) {
                    prefetch_l1_any(kfp + (size_t)(2 * U + u) * P);
                    prefetch_l1_any(kbp - (size_t)(2 * U + u) * P);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                sfp[(size_t)u * P] = fwd((float)fc[u]);
                *(sbp - (size_t)u * P) = bwd((float)bc[u]);
            }
            kfp += U * P; sfp += U * P; kbp -= U * P; sbp -= U * P;
#pragma unroll
            for (int u = 0; u < U; ++u) { fc[u] = fn[u]; bc[u] = bn[u]; }
        }
    }
    for (; t < h; ++t) {
        *sfp = fwd((float)*kfp);
        *sbp = bwd((float)*kbp);
        kfp += P; sfp += P; kbp -= P; sbp -= P;
    }
    if (has_mid) {  // both chains arrive at the middle row
        const float kc = (float)*kfp;
        const float va = fwd(kc), vd = bwd(kc);
        lap += (double)(ds * rcp_fast(fminf(va, vd)));
        kfp += P; sfp += P; kbp -= P; sbp -= P;
    }
    // ---- phase 2: meet the parked values, minimum, lap sum -------------------------------------------------
    t = 0;
    {
        KT_ fc[U], bc[U], fn[U], bn[U];
        float fo[U], bo[U], fon[U], bon[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool in = (u < h);
            fc[u] = in ? kfp[(size_t)u * P] : (KT_)1;
            bc[u] = in ? *(kbp - (size_t)u * P) : (KT_)1;
            fo[u] = in ? sfp[(size_t)u * P] : 1.0f;
            bo[u] = in ? *(sbp - (size_t)u * P) : 1.0f;
        }
        for (; t + U <= h; t += U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool in = (t + U + u < h);
                fn[u] = in ? kfp[(size_t)(U + u) * P] : (KT_)1;
                bn[u] = in ? *(kbp - (size_t)(U + u) * P) : (KT_)1;
                fon[u] = in ? sfp[(size_t)(U + u) * P] : 1.0f;
                bon[u] = in ? *(sbp - (size_t)(U + u) * P) : 1.0f;
            }
            if (F32_PREFETCH) {
#pragma unroll
                for (int u = 0; u < U; u += 2) {
                    prefetch_l1_any(kfp + (size_t)(2 * U + u) * P);
                    prefetch_l1_any(kbp - (size_t)(2 * U + u) * P);
                    prefetch_l1_any(sfp + (size_t)(2 * U + u) * P);
                    prefetch_l1_any(sbp - (size_t)(2 * U + u) * P);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float va = fwd((float)fc[u]), vd = bwd((float)bc[u]);
                lap += (double)(ds * rcp_fast(fminf(va, fo[u])) + ds * rcp_fast(fminf(bo[u], vd)));  // velocity.py:26, tbn.py:53
            }
            kfp += U * P; sfp += U * P; kbp -= U * P; sbp -= U * P;
#pragma unroll
            for (int u = 0; u < U; ++u) { fc[u] = fn[u]; bc[u] = bn[u]; fo[u] = fon[u]; bo[u] = bon[u]; }
        }
    }
    for (; t < h; ++t) {
        const float va = fwd((float)*kfp), vd = bwd((float)*kbp);
        lap += (double)(ds * rcp_fast(fminf(va, *sfp)) + ds * rcp_fast(fminf(*sbp, vd)));
        kfp += P; sfp += P; kbp -= P; sbp -= P;
    }
    a.lap[b] = lap;
}

}  // namespace ltk
