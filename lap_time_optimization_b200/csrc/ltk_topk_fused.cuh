// ltk_topk_fused.cuh -- the selection key of the top-k kernels and the top-k stage the sweep kernels run in
// their epilogue.   reference: sorted(zip(laps, index))[0:10]  (trajectory_bayesian_nonlinear.py:253-257)
#pragma once

namespace ltk {

// stable ascending order by (lap, index)
struct Key {
    double lap;
    long long idx;
};
__device__ __forceinline__ bool key_less(const Key& x, const Key& y)
{
    return (x.lap < y.lap) || (x.lap == y.lap && x.idx < y.idx);
}
__device__ __forceinline__ Key key_min(Key x, Key y) { return key_less(y, x) ? y : x; }

// ------------------------------------------------------------------------------------------------
// top-k fused into the sweep kernels' epilogue (SURVEY kernel table: "K3 ... + per-block top-k").
//
// Every sweep CTA holds the lap times of its candidates in registers when it finishes: it selects its own k best
// (k rounds of block minimum on shuffles), publishes them to slot `slot` of the scratch and takes a ticket of its
// group of `group` consecutive slots; the CTA that draws a group's last ticket merges the group's winners and
// takes a global ticket; the CTA that draws the last global ticket merges the group winners into the final
// result and leaves every counter at zero.  No extra launch, no second pass over the lap array; the result is
// the one topk_select produces (same key order: lap, then index; NaN last).
// Limits (checked on the host, which otherwise launches topk_select): k <= FUSE_K_MAX, group * k and
// n_groups * k <= FUSE_THREADS * FUSE_E.
// ------------------------------------------------------------------------------------------------
constexpr int FUSE_THREADS = 64;   // both sweep kernels run 64-thread CTAs
constexpr int FUSE_E = 16;         // keys per thread in a merge stage
constexpr int FUSE_K_MAX = 16;
struct TopkFuse {
    double* mid_lap;        // [(n_slots + n_groups) * k]: CTA winners, then group winners
    long long* mid_idx;
    unsigned* tickets;      // [1 + n_groups], zero between launches: global counter, then one per group
    double* out_lap;        // [k]
    long long* out_idx;
    long long index_base;
    int k;                  // 0: no fused selection
    int slot_base;          // first slot of this launch (a population split over two kernels shares the scratch)
    int n_slots, group, n_groups;
};

// k rounds of block-wide minimum over NK keys per thread (FUSE_THREADS threads); winners in order to out_*[0..k)
template <int NK>
__device__ __forceinline__ void fuse_rounds(Key (&key)[NK], int k, double* out_lap, long long* out_idx,
                                            Key (&wbest)[2][FUSE_THREADS / 32])
{
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const long long NONE = 0x7fffffffffffffffLL;
    for (int r = 0; r < k; ++r) {
        Key best = key[0];
#pragma unroll
        for (int j = 1; j < NK; ++j) best = key_min(best, key[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Key other{__shfl_xor_sync(0xffffffffu, best.lap, o), __shfl_xor_sync(0xffffffffu, best.idx, o)};
            best = key_min(best, other);
        }
        if ((threadIdx.x & 31) == 0) wbest[r & 1][threadIdx.x >> 5] = best;
        __syncthreads();
        const Key win = key_min(wbest[r & 1][0], wbest[r & 1][1]);
        const bool found = win.idx != NONE;
        if (threadIdx.x == 0) {
            out_lap[r] = found ? win.lap : INF;
            out_idx[r] = found ? win.idx : -1;
        }
#pragma unroll
        for (int j = 0; j < NK; ++j)
            if (key[j].idx == win.idx) { key[j].lap = INF; key[j].idx = NONE; }
    }
}

// merge `count` published keys into out_*[0..k)  (a call: its 16 keys per thread stay out of the sweeps' register budget)
__device__ __noinline__ void fuse_merge(const double* mid_lap, const long long* mid_idx, int k, int count,
                                       double* out_lap, long long* out_idx, Key (&wbest)[2][FUSE_THREADS / 32])
{
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const long long NONE = 0x7fffffffffffffffLL;
    Key key[FUSE_E];
#pragma unroll
    for (int j = 0; j < FUSE_E; ++j) {
        const int e = threadIdx.x + j * FUSE_THREADS;
        key[j].lap = INF;
        key[j].idx = NONE;
        if (e < count) {
            const long long ix = __ldcg(mid_idx + e);
            if (ix >= 0) { key[j].lap = __ldcg(mid_lap + e); key[j].idx = ix; }
        }
    }
    __syncthreads();  // wbest of the previous selection is no longer read
    fuse_rounds<FUSE_E>(key, k, out_lap, out_idx, wbest);
}

// Called by EVERY thread of a sweep CTA once its lap times are final.  `mine`: this thread holds a candidate of
// the population (global index t.index_base + b); every other thread contributes nothing.
__device__ __forceinline__ void topk_epilogue(const TopkFuse& t, bool mine, double lap, long long b)
{
    __shared__ Key wbest[2][FUSE_THREADS / 32];
    __shared__ unsigned drawn;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const long long NONE = 0x7fffffffffffffffLL;
    const int k = t.k;
    const int slot = t.slot_base + (int)blockIdx.x;
    Key key[1];
    key[0].lap = INF;
    key[0].idx = NONE;
    if (mine) { key[0].lap = (lap != lap) ? INF : lap; key[0].idx = t.index_base + b; }
    fuse_rounds<1>(key, k, t.mid_lap + (long long)slot * k, t.mid_idx + (long long)slot * k, wbest);
    // ---- group stage --------------------------------------------------------------------------------
    const int g = slot / t.group;
    const int g_first = g * t.group;
    const int g_size = min(t.group, t.n_slots - g_first);
    __threadfence();  // this CTA's winners are visible before its ticket is
    __syncthreads();
    if (threadIdx.x == 0) drawn = atomicAdd(t.tickets + 1 + g, 1u);
    __syncthreads();
    if (drawn != (unsigned)(g_size - 1)) return;
    __threadfence();
    const long long g_out = ((long long)t.n_slots + g) * k;
    fuse_merge(t.mid_lap + (long long)g_first * k, t.mid_idx + (long long)g_first * k, k, g_size * k,
               t.mid_lap + g_out, t.mid_idx + g_out, wbest);
    // ---- final stage --------------------------------------------------------------------------------
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        t.tickets[1 + g] = 0u;  // every member of the group has drawn
        drawn = atomicAdd(t.tickets, 1u);
    }
    __syncthreads();
    if (drawn != (unsigned)(t.n_groups - 1)) return;
    __threadfence();
    fuse_merge(t.mid_lap + (long long)t.n_slots * k, t.mid_idx + (long long)t.n_slots * k, k, t.n_groups * k,
               t.out_lap, t.out_idx, wbest);
    if (threadIdx.x == 0) t.tickets[0] = 0u;
}

}  // namespace ltk
