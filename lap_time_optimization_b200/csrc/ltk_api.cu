// ltk_api.cu -- C ABI (include/ltk.h) over the kernels in ltk_kernels.cuh.  sm_100a only.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "ltk_kernels.cuh"

using namespace ltk;

#define LTK_HOST_GRAPHS 8          // cached graphs per context (least recently used one is replaced)
#define LTK_HOST_GRAPH_MAX_B 16384  // larger host batches are launched directly (launch cost no longer matters)

static std::atomic<long long> g_launches{0};
static char g_create_err[256] = "";

struct ltk_ctx {
    int device;
    int N, ns;
    int sm_count;
    size_t smem_optin;
    double* d_left;  // [2][N]
    double* d_diff;  // [2][N]
    VehDev veh;
    VehF32 veh32;
    int sweep_bits;  // 64 (default) or 32: the optional fp32 sweep variant
    int spline_mode; // LTK_SPLINE_TRIDIAGONAL (default) or LTK_SPLINE_FITPACK
    // scratch owned by the context
    int4* d_lut;  // engine cell table of the fused sweep (nullptr: none)
    double* d_topk_lap[2];  // ping-pong scratch of the top-k stages, grown on demand
    long long* d_topk_idx[2];
    long long topk_cap;     // entries per scratch buffer
    unsigned* d_ticket;     // [TICKET_CAP] "last block finishes" counters (zero between launches): [0] global, then
                            // one per group of the selection fused into the sweeps (ltk_topk_fused.cuh)
    cudaStream_t aux_stream;  // the remainder sweep (see run_pipeline) runs next to the main one
    cudaEvent_t aux_ev[2];
    int sweep_remainder;      // 0 disables the split (LTK_SWEEP_REMAINDER=0)
    // tracing (ltk_trace_*): one (start, end) event pair per kernel launch of the pipeline
    cudaEvent_t* trace_ev;    // [2 * trace_cap]
    int* trace_kind;          // LTK_TRACE_K1A ...
    int trace_cap, trace_n;
    void* d_profile_ws;
    size_t profile_ws_bytes;
    // host-in / host-out small-batch path (ltk_eval_alphas_host): pinned staging, device buffers, a stream
    // of its own and one instantiated CUDA graph (upload, K1a, K1b, K23, download) per batch size
    cudaStream_t host_stream;
    double *h_in, *h_out, *d_hin, *d_hout;
    char* d_hws;
    long long host_cap;      // candidates the buffers hold
    size_t host_ws_bytes;
    struct HostGraph { long long B; unsigned long long epoch; cudaGraphExec_t exec; } host_graph[LTK_HOST_GRAPHS];
    unsigned long long epoch;  // bumped by every setter that changes what the pipeline launches
    unsigned long long host_tick;
    unsigned long long host_used[LTK_HOST_GRAPHS];
    int k1_g_override, k1_staged_override, k1_threads_override, sweep_mode, k1_mode;
    char err[256];
};

namespace {


struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int fail(ltk_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess)
{
    char* dst = ctx ? ctx->err : g_create_err;
    if (e != cudaSuccess) snprintf(dst, 256, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(dst, 256, "%s", what);
    return code;
}

#define LTK_CUDA(ctx, call)                                                   \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return fail(ctx, LTK_E_CUDA, #call, e_);       \
    } while (0)

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

VehDev make_vehdev(const ltk_vehicle& v)
{
    VehDev d;
    memset(&d, 0, sizeof(d));
    d.kind = v.kind;
    d.n_map = v.n_map;
    d.mass = v.mass;
    d.half_mass = 0.5 * v.mass;           // exact
    d.inv_half_mass = 1.0 / d.half_mass;  // correctly rounded; see div_by_const / forward_step
    d.mu_g = v.mu_g;
    d.f_max = v.f_max;
    d.f_max_sq = v.f_max_sq;
    d.e0 = v.e0;
    d.cr2 = v.cr2;
    for (int i = 0; i < LTK_MAX_ENGINE_MAP; ++i) d.thr[i] = 0x7fffffffffffffffLL;
    if (v.kind == 0) {
        const int n = v.n_map;
        for (int i = 0; i < n; ++i) memcpy(&d.thr[i], &v.map_v[i], sizeof(double));
        // extended segment table, see VehDev; slopes exactly as np.interp forms them
        d.ext_b[0] = v.map_v[0]; d.ext_f[0] = v.map_f[0]; d.ext_s[0] = 0.0;
        for (int j = 1; j < n; ++j) {
            d.ext_b[j] = v.map_v[j - 1];
            d.ext_f[j] = v.map_f[j - 1];
            d.ext_s[j] = (v.map_f[j] - v.map_f[j - 1]) / (v.map_v[j] - v.map_v[j - 1]);
        }
        d.ext_b[n] = v.map_v[n - 1]; d.ext_f[n] = v.map_f[n - 1]; d.ext_s[n] = 0.0;
    }
    // curvature window of the regular path: 2^-511 <= k < 2^513 ...
    d.k_lo_hi = 0x20000000;
    d.k_span_hi = 0x40000000u;
    bool engine_nonneg = true;
    if (v.kind == 0) {
        for (int i = 0; i < v.n_map; ++i) engine_nonneg = engine_nonneg && (v.map_f[i] >= 0.0);
    } else if (v.cr2 > 0.0) {
        // ... and, for the polynomial engine e0 - cr2 v^2, k large enough that v^2 <= mu g / k keeps it >= 0
        if (v.e0 > 0.0 && v.mu_g > 0.0) {
            double k_min = v.mu_g * v.cr2 / (0.99 * v.e0);
            int hi;
            long long bits;
            memcpy(&bits, &k_min, sizeof(bits));
            hi = (int)(bits >> 32) + 1;  // round the window's lower end up
            if (hi > d.k_lo_hi) {
                unsigned top = (unsigned)d.k_lo_hi + d.k_span_hi;
                d.k_span_hi = ((unsigned)hi < top) ? top - (unsigned)hi : 0u;
                d.k_lo_hi = hi;
            }
        } else {
            engine_nonneg = false;
        }
    } else {
        engine_nonneg = (v.e0 >= 0.0);
    }
    if (!engine_nonneg) d.k_span_hi = 0u;  // never take the regular path (it assumes accel >= 0)
    // the regular path also assumes mu g / k, f^2 - f_lat^2 and 2 F / m stay far from the ends of the exponent
    // range for every k in the window: true for any physical vehicle, checked so that an odd one is only slow
    auto moderate = [](double x) { return x >= 0x1p-100 && x <= 0x1p100; };
    if (!moderate(v.mu_g) || !moderate(v.f_max) || !moderate(v.mass)) d.k_span_hi = 0u;
    d.lut_shift = 0; d.lut_base = 0; d.lut_top = -1;
    return d;
}

VehF32 make_vehf32(const ltk_vehicle& v)
{
    VehF32 d;
    memset(&d, 0, sizeof(d));
    d.kind = v.kind; d.n_map = v.n_map;
    d.mass = (float)v.mass; d.mu_g = (float)v.mu_g; d.f_max = (float)v.f_max; d.f_max_sq = (float)v.f_max_sq;
    d.e0 = (float)v.e0; d.cr2 = (float)v.cr2;
    for (int i = 0; i < LTK_MAX_ENGINE_MAP; ++i) d.thr[i] = __builtin_inff();
    if (v.kind == 0) {
        const int n = v.n_map;
        for (int i = 0; i < n; ++i) d.thr[i] = (float)v.map_v[i];
        d.ext_b[0] = (float)v.map_v[0]; d.ext_f[0] = (float)v.map_f[0]; d.ext_s[0] = 0.0f;
        for (int j = 1; j < n; ++j) {
            d.ext_b[j] = (float)v.map_v[j - 1];
            d.ext_f[j] = (float)v.map_f[j - 1];
            d.ext_s[j] = (float)((v.map_f[j] - v.map_f[j - 1]) / (v.map_v[j] - v.map_v[j - 1]));
        }
        d.ext_b[n] = (float)v.map_v[n - 1]; d.ext_f[n] = (float)v.map_f[n - 1]; d.ext_s[n] = 0.0f;
    }
    return d;
}

// Engine cell table (EngineLut in ltk_sweep_fused.cuh).  Picks the coarsest split of the high word that
// leaves at most one map node per cell; returns the number of cells (0: no table fits).
int build_engine_lut(const ltk_vehicle& v, VehDev* d, int4* cells)
{
    if (v.kind != 0) return 0;
    const int n = v.n_map;
    int hi[LTK_MAX_ENGINE_MAP];
    long long bits[LTK_MAX_ENGINE_MAP];
    for (int i = 0; i < n; ++i) {
        memcpy(&bits[i], &v.map_v[i], sizeof(double));
        hi[i] = (int)(bits[i] >> 32);
    }
    for (int shift = 20; shift >= 0; --shift) {
        bool distinct = true;
        for (int i = 1; i < n && distinct; ++i) distinct = (hi[i] >> shift) != (hi[i - 1] >> shift);
        if (!distinct) continue;
        const int first = hi[0] >> shift, last = hi[n - 1] >> shift;
        const long long ncell = (long long)last - first + 3;  // one catch-all cell on either side
        if (ncell > LTK_LUT_MAX_CELLS) return 0;              // finer shifts only need more cells
        const int base = first - 1;
        const long long never = 0x7fffffffffffffffLL;
        int node = 0;  // nodes below the current cell
        for (int c = 0; c < (int)ncell; ++c) {
            const int cell_id = base + c;
            long long thr = never;
            if (node < n && (hi[node] >> shift) == cell_id && c > 0 && c < (int)ncell - 1) thr = bits[node];
            cells[c].x = (int)(unsigned)(thr & 0xffffffffLL);
            cells[c].y = (int)(thr >> 32);
            cells[c].z = node;
            cells[c].w = 0;
            if (thr != never) ++node;
        }
        d->lut_shift = shift;
        d->lut_base = base;
        d->lut_top = (int)ncell - 1;
        return (int)ncell;
    }
    return 0;
}

int check_vehicle(const ltk_vehicle* v)
{
    if (!v) return 0;
    if (v->kind != 0 && v->kind != 1) return 0;
    if (v->kind == 0 && (v->n_map < 2 || v->n_map > LTK_MAX_ENGINE_MAP)) return 0;
    if (v->kind == 0)
        for (int i = 0; i < v->n_map; ++i) {
            if (!(v->map_v[i] >= 0.0)) return 0;                      // integer-ordered comparison needs x >= 0
            if (i > 0 && !(v->map_v[i] > v->map_v[i - 1])) return 0;  // np.interp needs increasing abscissae
        }
    if (!(v->mass > 0.0)) return 0;
    return 1;
}

struct WsLayout {
    long long Bp;
    size_t kap_off, vacc_off, len_off, rot_off, mx_off, my_off, knots_off, vaccd_off, vdec_off, vmin_off, fit_off, total;
};

WsLayout ws_layout(int ns, int N, long long B, bool dumps, bool fitpack = false)
{
    WsLayout w;
    w.Bp = round_up(B < 1 ? 1 : B, 32);  // multiple of TILE and of the warp size
    size_t rows = (size_t)(ns - 1);
    size_t arr = rows * (size_t)w.Bp * sizeof(double);
    // K23's look-ahead loads are unconditional: up to 2 * FUSED_UNROLL rows before the first / after the last
    // row of a tile are read (never used).  Inside the arrays that is a neighbouring tile; the front of the
    // curvature array gets SWEEP_SLACK bytes, the staging array is followed by the small per-candidate arrays.
    size_t off = SWEEP_SLACK;
    w.kap_off = off; off += arr;
    w.vacc_off = off; off += arr;
    w.len_off = off; off += (size_t)w.Bp * sizeof(double);
    w.rot_off = off; off += round_up((long long)w.Bp * sizeof(int), 256);
    w.mx_off = off; off += (size_t)N * (size_t)w.Bp * sizeof(double);
    w.my_off = off; off += (size_t)N * (size_t)w.Bp * sizeof(double);
    w.knots_off = off; off += (size_t)(N + 1) * (size_t)w.Bp * sizeof(double);
    w.vdec_off = w.vmin_off = w.vaccd_off = w.fit_off = 0;
    if (fitpack) {  // K1a-F scratch and hand-off (ltk_fitpack.cuh)
        w.fit_off = off; off += fit_region_doubles(N) * (size_t)w.Bp * sizeof(double);
    }
    if (dumps) {
        w.vaccd_off = off; off += arr;
        w.vdec_off = off; off += arr;
        w.vmin_off = off; off += arr;
    }
    w.total = off;
    return w;
}

WsLayout ctx_layout(const ltk_ctx* ctx, long long B, bool dumps)
{
    return ws_layout(ctx->ns, ctx->N, B, dumps, ctx->spline_mode == LTK_SPLINE_FITPACK);
}

FitArgs fit_args(const ltk_ctx* ctx, char* ws, const WsLayout& w)
{
    FitArgs f;
    const size_t Bp = (size_t)w.Bp, N = (size_t)ctx->N;
    double* p = reinterpret_cast<double*>(ws + w.fit_off);
    f.hand = p; p += (5 * N + 13) * Bp;  // first: the packed blocks stay 16-byte aligned (bulk copy)
    f.rows = p; p += 7 * N * Bp;
    f.cx = p; p += (N + 3) * Bp;
    f.cy = p;
    return f;
}

struct K1Config {
    int G, threads, staged;
    size_t smem;
};

int k1a_threads(const ltk_ctx* ctx);

bool pick_k1(const ltk_ctx* ctx, K1Config* out)
{
    const int cand_g[2] = {8, 16};  // measured: two co-resident half-tile CTAs per SM beat one full-tile CTA
    for (int staged = 1; staged >= 0; --staged) {
        if (ctx->k1_staged_override >= 0 && staged != ctx->k1_staged_override) continue;
        for (int gi = 0; gi < 2; ++gi) {
            int G = cand_g[gi];
            if (ctx->k1_g_override > 0 && G != ctx->k1_g_override) continue;
            int threads = ctx->k1_threads_override > 0 ? ctx->k1_threads_override : 512;
            size_t s = k1_smem_bytes(G, threads, ctx->N, ctx->ns, staged);
            if (s <= ctx->smem_optin && k1a_threads(ctx) >= 32) {
                out->G = G; out->threads = threads; out->staged = staged; out->smem = s;
                return true;
            }
        }
    }
    return false;
}

// threads per CTA of the solve kernel: its scratch is 4 doubles per (row, thread)
int k1a_threads(const ltk_ctx* ctx)
{
    size_t per_thread = (size_t)4 * ctx->N * sizeof(double);
    long long t = (long long)(ctx->smem_optin / per_thread) / 32 * 32;
    if (t > 128) t = 128;
    return (int)t;  // 0: does not fit
}

cudaError_t launch_k1a(const ltk_ctx* ctx, const K1Args& a, cudaStream_t st)
{
    int T = k1a_threads(ctx);
    size_t smem = (size_t)4 * ctx->N * T * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(k1a_spline_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    unsigned grid = (unsigned)((a.Bp + T - 1) / T);
    k1a_spline_solve<<<grid, T, smem, st>>>(a);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

template <int G, int THREADS>
cudaError_t launch_k1b(const K1Args& a, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k1b_curvature<G, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    unsigned grid = (unsigned)(a.Bp / G);
    k1b_curvature<G, THREADS><<<grid, THREADS, smem, st>>>(a);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

// K1b (ltk_spline.cuh): G candidates per CTA, T threads
struct K1FConfig {
    int G, threads, staged;
    size_t smem;
};

bool pick_k1f(const ltk_ctx* ctx, K1FConfig* out)
{
    const bool fitp = ctx->spline_mode == LTK_SPLINE_FITPACK;
    if (ctx->k1_mode == 1 && !fitp) return false;  // LTK_K1=old: the previous K1a + K1b pair (A/B reference)
    if ((fitp ? k1af_smem_bytes(ctx->N) : k1a_smem_bytes(ctx->N)) > ctx->smem_optin) return false;
    // measured order.  The shared-memory tile (staged) wins while four candidates' rows fit; beyond that
    // (ns > ~6,600) the two-pass variant at G = 4 does (profiles/README.md, ns = 10,001).  G = 1, 2 tiles are
    // only reachable through LTK_K1_G.
    const int cand[8][3] = {{4, 256, 1}, {4, 128, 1}, {8, 256, 1}, {8, 512, 0}, {2, 128, 1}, {2, 64, 1}, {1, 256, 1}, {1, 128, 1}};
    for (int i = 0; i < 8; ++i) {
        int G = cand[i][0], T = cand[i][1], staged = cand[i][2];
        if (ctx->k1_g_override > 0 && G != ctx->k1_g_override) continue;
        if (ctx->k1_threads_override > 0 && T != ctx->k1_threads_override) continue;
        if (ctx->k1_staged_override >= 0 && staged != ctx->k1_staged_override) continue;
        size_t s = k1f_smem_bytes(G, T, ctx->N, ctx->ns, fitp, staged != 0);
        if (s <= ctx->smem_optin) { out->G = G; out->threads = T; out->staged = staged; out->smem = s; return true; }
    }
    return false;
}

template <int G, int T, int MINB, bool FIT, bool STAGED = true, bool F32K = false>
cudaError_t launch_k1f(const K1Args& a, const FitArgs& fa, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k1b_samples<G, T, MINB, FIT, STAGED, F32K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k1b_samples<G, T, MINB, FIT, STAGED, F32K><<<(unsigned)(a.Bp / G), T, smem, st>>>(a, fa);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

template <bool FIT>
cudaError_t launch_k1f_cfg(const K1FConfig& c, const K1Args& a, const FitArgs& fa, cudaStream_t st)
{
    if (!c.staged) return launch_k1f<8, 512, 2, FIT, false>(a, fa, c.smem, st);
    if (c.G == 1 && c.threads == 256) return launch_k1f<1, 256, 2, FIT>(a, fa, c.smem, st);
    if (c.G == 1) return launch_k1f<1, 128, 4, FIT>(a, fa, c.smem, st);
    if (c.G == 2 && c.threads == 128) return launch_k1f<2, 128, 8, FIT>(a, fa, c.smem, st);
    if (c.G == 2) return launch_k1f<2, 64, 12, FIT>(a, fa, c.smem, st);
    if (c.G == 4 && c.threads == 128) return launch_k1f<4, 128, 5, FIT>(a, fa, c.smem, st);
    if (c.G == 4) {
        // CTAs per SM: 4 (64 registers) for the tridiagonal mode; the FITPACK mode's 160-byte interval record spills
        // there and runs faster at 3 (80 registers): 0.51 -> 0.41 ms; the tridiagonal mode loses (0.283 -> 0.306).
        // LTK_K1B_MINB overrides (A/B).
        static const int minb = getenv("LTK_K1B_MINB") ? atoi(getenv("LTK_K1B_MINB")) : 0;
        const int m = minb ? minb : (FIT ? 3 : 4);
        if constexpr (!FIT) {
            // fp32 variant: curvature evaluated in fp32 when only the fp32 copy is wanted (LTK_K1B_F32=0: fp64 + copy)
            static const bool f32k = !(getenv("LTK_K1B_F32") && atoi(getenv("LTK_K1B_F32")) == 0);
            if (a.kap32 && f32k) return launch_k1f<4, 256, 4, false, true, true>(a, fa, c.smem, st);
        }
        if (m == 5) return launch_k1f<4, 256, 5, FIT>(a, fa, c.smem, st);
        if (m == 3) return launch_k1f<4, 256, 3, FIT>(a, fa, c.smem, st);
        if (m == 2) return launch_k1f<4, 256, 2, FIT>(a, fa, c.smem, st);
        return launch_k1f<4, 256, 4, FIT>(a, fa, c.smem, st);
    }
    return launch_k1f<8, 256, 2, FIT>(a, fa, c.smem, st);
}

cudaError_t launch_k1a_fitpack(const ltk_ctx* ctx, const K1Args& a, const FitArgs& fa, cudaStream_t st)
{
    size_t smem = k1af_smem_bytes(ctx->N);
    cudaError_t e = cudaFuncSetAttribute(k1a_fitpack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k1a_fitpack<<<(unsigned)(a.Bp / 32), K1AF_THREADS, smem, st>>>(a, fa);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

cudaError_t launch_k1a_solve(const ltk_ctx* ctx, const K1Args& a, cudaStream_t st)
{
    size_t smem = k1a_smem_bytes(ctx->N);
    cudaError_t e = cudaFuncSetAttribute(k1a_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k1a_solve<<<(unsigned)(a.Bp / 32), K1A_THREADS, smem, st>>>(a);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

cudaError_t launch_k1b_cfg(const K1Config& c, const K1Args& a, cudaStream_t st)
{
    if (c.G == 16) {
        if (c.threads == 1024) return launch_k1b<16, 1024>(a, c.smem, st);
        if (c.threads == 256) return launch_k1b<16, 256>(a, c.smem, st);
        return launch_k1b<16, 512>(a, c.smem, st);
    }
    if (c.threads == 1024) return launch_k1b<8, 1024>(a, c.smem, st);
    if (c.threads == 256) return launch_k1b<8, 256>(a, c.smem, st);
    return launch_k1b<8, 512>(a, c.smem, st);
}

// top-k of the population the pipeline scores, wanted together with the lap times: run_pipeline fuses the
// selection into the sweep kernels' epilogue where it can and says so in `fused`
struct TopkReq {
    int k;
    long long index_base;
    double* best_lap;
    long long* best_idx;
    bool fused;
};
constexpr int TICKET_CAP = 1 + 1024;
int ensure_topk_scratch(ltk_ctx* ctx, long long entries, cudaStream_t st);

int run_pipeline(ltk_ctx* ctx, const double* d_alphas, const double* d_xy, int m, long long B, double* d_lap, char* ws,
                 const WsLayout& w, bool dumps, cudaStream_t st, cudaEvent_t* ev, bool k1_only, TopkReq* topk = nullptr);

// the pipeline launches on a laid-out workspace
int run_pipeline(ltk_ctx* ctx, const double* d_alphas, const double* d_xy, int m, long long B,
                 double* d_lap, char* ws, const WsLayout& w, bool dumps, cudaStream_t st,
                 cudaEvent_t* ev = nullptr)
{
    return run_pipeline(ctx, d_alphas, d_xy, m, B, d_lap, ws, w, dumps, st, ev, false);
}

int run_pipeline(ltk_ctx* ctx, const double* d_alphas, const double* d_xy, int m, long long B,
                 double* d_lap, char* ws, const WsLayout& w, bool dumps, cudaStream_t st,
                 cudaEvent_t* ev, bool k1_only, TopkReq* topk)
{
    if (topk) topk->fused = false;
    K1Config cfg;
    K1FConfig fcfg;
    const bool k1_one_kernel = pick_k1f(ctx, &fcfg);
    if (!k1_one_kernel && (ctx->spline_mode == LTK_SPLINE_FITPACK || !pick_k1(ctx, &cfg)))
        return fail(ctx, LTK_E_UNSUPPORTED, "control-point count or sampling density too large for shared memory");
    K1Args a;
    a.alphas = d_alphas; a.xy = d_xy; a.mode = d_xy ? 1 : 0; a.m = m;
    a.left = ctx->d_left; a.diff = ctx->d_diff; a.N = ctx->N; a.ns = ctx->ns;
    a.B = B; a.Bp = w.Bp;
    a.mu_g = ctx->veh.mu_g;
    a.kap = reinterpret_cast<double*>(ws + w.kap_off);
    // fp32 sweeps read an fp32 copy of the curvature, kept in the upper half of the staging array (its lower
    // half holds the fp32 parked velocities); only the current K1b writes it
    a.kap32 = nullptr;
    a.rot = reinterpret_cast<int*>(ws + w.rot_off);
    a.len = reinterpret_cast<double*>(ws + w.len_off);
    a.mx = reinterpret_cast<double*>(ws + w.mx_off);
    a.my = reinterpret_cast<double*>(ws + w.my_off);
    a.knots = reinterpret_cast<double*>(ws + w.knots_off);
    a.hand = a.mx;  // (3 N + 1) Bp doubles: the packed hand-off of the k1a_solve / k1b_samples pair
    auto trace_open = [&](int kind, cudaStream_t s_) {
        if (ctx->trace_ev && ctx->trace_n < ctx->trace_cap) {
            ctx->trace_kind[ctx->trace_n] = kind;
            cudaEventRecord(ctx->trace_ev[2 * ctx->trace_n], s_);
        }
    };
    auto trace_close = [&](cudaStream_t s_) {
        if (ctx->trace_ev && ctx->trace_n < ctx->trace_cap) {
            cudaEventRecord(ctx->trace_ev[2 * ctx->trace_n + 1], s_);
            ++ctx->trace_n;
        }
    };
    if (ev) LTK_CUDA(ctx, cudaEventRecord(ev[0], st));
    if (k1_one_kernel) {
        a.staged = 1;
        if (ctx->sweep_bits == 32 && !dumps && !k1_only)
            a.kap32 = reinterpret_cast<float*>(ws + w.vacc_off) + (size_t)(ctx->ns - 1) * (size_t)w.Bp;
        const bool fitp = ctx->spline_mode == LTK_SPLINE_FITPACK;
        FitArgs fa;
        memset(&fa, 0, sizeof(fa));
        if (fitp) fa = fit_args(ctx, ws, w);
        trace_open(LTK_TRACE_K1A, st);
        if (fitp) LTK_CUDA(ctx, launch_k1a_fitpack(ctx, a, fa, st));
        else LTK_CUDA(ctx, launch_k1a_solve(ctx, a, st));
        trace_close(st);
        if (ev) LTK_CUDA(ctx, cudaEventRecord(ev[1], st));
        trace_open(LTK_TRACE_K1B, st);
        if (fitp) LTK_CUDA(ctx, launch_k1f_cfg<true>(fcfg, a, fa, st));
        else LTK_CUDA(ctx, launch_k1f_cfg<false>(fcfg, a, fa, st));
        trace_close(st);
    } else {
        a.staged = cfg.staged;
        LTK_CUDA(ctx, launch_k1a(ctx, a, st));
        if (ev) LTK_CUDA(ctx, cudaEventRecord(ev[1], st));
        LTK_CUDA(ctx, launch_k1b_cfg(cfg, a, st));
    }
    if (ev) LTK_CUDA(ctx, cudaEventRecord(ev[2], st));
    if (k1_only) return LTK_OK;

    if (ctx->sweep_bits == 32 && !dumps) {  // optional fp32 sweeps (ltk_sweep_f32.cuh)
        F32Args f;
        f.kap = a.kap;
        f.stage = reinterpret_cast<float*>(ws + w.vacc_off);
        f.len = a.len; f.lap = d_lap; f.ns = ctx->ns; f.B = B; f.Bp = w.Bp;
        f.kap32 = a.kap32;
        unsigned g32 = (unsigned)((B + F32_THREADS - 1) / F32_THREADS);
        trace_open(LTK_TRACE_K23, st);
        if (f.kap32) {
            if (ctx->veh.kind == 0 && ctx->veh.n_map <= 8) k23_f32<0, 8, true><<<g32, F32_THREADS, 0, st>>>(f, ctx->veh32);
            else if (ctx->veh.kind == 0) k23_f32<0, 16, true><<<g32, F32_THREADS, 0, st>>>(f, ctx->veh32);
            else k23_f32<1, 8, true><<<g32, F32_THREADS, 0, st>>>(f, ctx->veh32);
        } else {
            if (ctx->veh.kind == 0 && ctx->veh.n_map <= 8) k23_f32<0, 8, false><<<g32, F32_THREADS, 0, st>>>(f, ctx->veh32);
            else if (ctx->veh.kind == 0) k23_f32<0, 16, false><<<g32, F32_THREADS, 0, st>>>(f, ctx->veh32);
            else k23_f32<1, 8, false><<<g32, F32_THREADS, 0, st>>>(f, ctx->veh32);
        }
        trace_close(st);
        if (ev) { LTK_CUDA(ctx, cudaEventRecord(ev[3], st)); LTK_CUDA(ctx, cudaEventRecord(ev[4], st)); }
        g_launches.fetch_add(1);
        LTK_CUDA(ctx, cudaGetLastError());
        return LTK_OK;
    }
    {
        FusedArgs f;
        f.lut = ctx->d_lut; f.Bp = w.Bp; f.first = 0; f.last = w.Bp;
        f.kap = a.kap;
        f.stage = reinterpret_cast<double*>(ws + w.vacc_off);
        f.rot = a.rot; f.len = a.len; f.lap = d_lap;
        f.vacc_d = dumps ? reinterpret_cast<double*>(ws + w.vaccd_off) : nullptr;
        f.vdec_d = dumps ? reinterpret_cast<double*>(ws + w.vdec_off) : nullptr;
        f.vmin_d = dumps ? reinterpret_cast<double*>(ws + w.vmin_off) : nullptr;
        f.ns = ctx->ns; f.B = B;
        unsigned gridf = (unsigned)((w.Bp + FUSED_THREADS - 1) / FUSED_THREADS);
        unsigned gridr = (unsigned)(w.Bp / 32);
        // Small batches are latency-bound: one chain per thread, two warps per 32 candidates (K23r).  Both kernels'
        // times step with the warps per scheduler (one two-chain warp per scheduler = 32 * 4 * SMs candidates, one
        // "layer"): measured K23 / K23r at 8,192: 0.261 / 0.217 ms, 20,480: 0.289 / 0.263, 24,576: 0.290 / 0.265,
        // 32,768: 0.292 / 0.318, 40,960: 0.398 / 0.399, 49,152: 0.398 / 0.479 -- one chain per thread up to one
        // and a half layers (28,416 candidates on 148 SMs), two chains per thread above.
        const long long layer = 32LL * 4 * ctx->sm_count;  // candidates in one two-chain warp per scheduler
        const bool roles = (ctx->sweep_mode == 2) || (ctx->sweep_mode == 0 && !dumps && 2 * (long long)B <= 3 * layer);
        // A population that fills the two-chain kernel's warp slots unevenly -- e.g. 65,536 candidates =
        // 3.46 warps per scheduler, so most schedulers carry 4 warps and the rest idle a quarter of the
        // time -- is split: whole layers of one warp per scheduler go to the two-chain kernel, the
        // remainder to the one-chain kernel (half-size warps, two per 32 candidates) on a second stream,
        // so that the busiest schedulers carry 3.5 warps' worth instead of 4.
        const long long whole = (w.Bp / layer) * layer;
        const long long rest = w.Bp - whole;
        const bool split = !roles && !dumps && ctx->sweep_mode == 0 && ctx->sweep_remainder && ctx->aux_stream &&
                           whole > 0 && rest > 0 && 2 * rest <= layer + layer / 8;
        // the selection of the k best rides in the sweep CTAs' epilogue when the two merge stages fit one CTA each
        f.tk.k = 0;
        static const bool fuse_off = getenv("LTK_TOPK_FUSED") && atoi(getenv("LTK_TOPK_FUSED")) == 0;
        if (topk && !dumps && !fuse_off && topk->k >= 1 && topk->k <= FUSE_K_MAX) {
            const long long n_slots = roles ? (long long)gridr : split ? whole / FUSED_THREADS + rest / 32 : (long long)gridf;
            const long long cap = (long long)FUSE_THREADS * FUSE_E / topk->k;  // keys one merge stage takes
            long long group = 1;
            while (group * group < n_slots) ++group;
            const long long n_groups = (n_slots + group - 1) / group;
            if (group <= cap && n_groups <= cap && 1 + n_groups <= TICKET_CAP) {
                int rc = ensure_topk_scratch(ctx, (n_slots + n_groups) * topk->k, st);
                if (rc != LTK_OK) return rc;
                f.tk.mid_lap = ctx->d_topk_lap[0]; f.tk.mid_idx = ctx->d_topk_idx[0];
                f.tk.tickets = ctx->d_ticket;
                f.tk.out_lap = topk->best_lap; f.tk.out_idx = topk->best_idx;
                f.tk.index_base = topk->index_base; f.tk.k = topk->k;
                f.tk.slot_base = 0;
                f.tk.n_slots = (int)n_slots; f.tk.group = (int)group; f.tk.n_groups = (int)n_groups;
                topk->fused = true;
            }
        }
        if (split) {
            f.first = whole; f.last = w.Bp;
            f.tk.slot_base = (int)(whole / FUSED_THREADS);  // the remainder's CTAs take the slots after the main launch's
            LTK_CUDA(ctx, cudaEventRecord(ctx->aux_ev[0], st));
            LTK_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[0], 0));
            unsigned gr = (unsigned)(rest / 32);
            if (ctx->veh.kind == 0) {
                if (ctx->veh.lut_top >= 0) k23_roles<0, 0><<<gr, ROLES_THREADS, 0, ctx->aux_stream>>>(f, ctx->veh);
                else if (ctx->veh.n_map <= 8) k23_roles<0, 8><<<gr, ROLES_THREADS, 0, ctx->aux_stream>>>(f, ctx->veh);
                else k23_roles<0, 16><<<gr, ROLES_THREADS, 0, ctx->aux_stream>>>(f, ctx->veh);
            } else {
                k23_roles<1, 8><<<gr, ROLES_THREADS, 0, ctx->aux_stream>>>(f, ctx->veh);
            }
            g_launches.fetch_add(1);
            LTK_CUDA(ctx, cudaEventRecord(ctx->aux_ev[1], ctx->aux_stream));
            f.first = 0; f.last = whole;
            f.tk.slot_base = 0;
            gridf = (unsigned)(whole / FUSED_THREADS);
        }
        trace_open(LTK_TRACE_K23, st);
        if (roles) {
            if (ctx->veh.kind == 0) {
                if (ctx->veh.lut_top >= 0) k23_roles<0, 0><<<gridr, ROLES_THREADS, 0, st>>>(f, ctx->veh);
                else if (ctx->veh.n_map <= 8) k23_roles<0, 8><<<gridr, ROLES_THREADS, 0, st>>>(f, ctx->veh);
                else k23_roles<0, 16><<<gridr, ROLES_THREADS, 0, st>>>(f, ctx->veh);
            } else {
                k23_roles<1, 8><<<gridr, ROLES_THREADS, 0, st>>>(f, ctx->veh);
            }
        } else if (ctx->veh.kind == 0) {
            if (ctx->veh.lut_top >= 0) k23_sweep<0, 0><<<gridf, FUSED_THREADS, 0, st>>>(f, ctx->veh);
            else if (ctx->veh.n_map <= 8) k23_sweep<0, 8><<<gridf, FUSED_THREADS, 0, st>>>(f, ctx->veh);
            else k23_sweep<0, 16><<<gridf, FUSED_THREADS, 0, st>>>(f, ctx->veh);
        } else {
            k23_sweep<1, 8><<<gridf, FUSED_THREADS, 0, st>>>(f, ctx->veh);
        }
        if (split) LTK_CUDA(ctx, cudaStreamWaitEvent(st, ctx->aux_ev[1], 0));
        trace_close(st);
        if (ev) { LTK_CUDA(ctx, cudaEventRecord(ev[3], st)); LTK_CUDA(ctx, cudaEventRecord(ev[4], st)); }
        g_launches.fetch_add(1);
    }
    LTK_CUDA(ctx, cudaGetLastError());
    return LTK_OK;
}

// top-k of `count` keys.  Up to TOPK_BLOCK_KEYS / k blocks: ONE launch (the block that finishes last
// merges every block's winners); beyond that, stages chained until few enough blocks are left.
// scratch of the top-k stages (two ping-pong buffers of `entries` keys) and the ticket counters; grows on demand
// (rare: first call / larger population -- the stream is drained before the old buffers are freed)
int ensure_topk_scratch(ltk_ctx* ctx, long long entries, cudaStream_t st)
{
    if (entries > ctx->topk_cap) {
        LTK_CUDA(ctx, cudaDeviceSynchronize());  // a split sweep of an earlier call may still use the scratch on the aux stream
        for (int i = 0; i < 2; ++i) {
            cudaFree(ctx->d_topk_lap[i]); cudaFree(ctx->d_topk_idx[i]);
            ctx->d_topk_lap[i] = nullptr; ctx->d_topk_idx[i] = nullptr;
        }
        ctx->topk_cap = 0;
        long long cap = 16384;
        while (cap < entries) cap *= 2;
        for (int i = 0; i < 2; ++i) {
            LTK_CUDA(ctx, cudaMalloc(&ctx->d_topk_lap[i], sizeof(double) * cap));
            LTK_CUDA(ctx, cudaMalloc(&ctx->d_topk_idx[i], sizeof(long long) * cap));
        }
        ctx->topk_cap = cap;
    }
    if (!ctx->d_ticket) {
        LTK_CUDA(ctx, cudaMalloc(&ctx->d_ticket, sizeof(unsigned) * TICKET_CAP));
        LTK_CUDA(ctx, cudaMemsetAsync(ctx->d_ticket, 0, sizeof(unsigned) * TICKET_CAP, st));
    }
    return LTK_OK;
}

int run_topk(ltk_ctx* ctx, const double* d_lap, const long long* d_idx, long long count, long long index_base, int k,
             double* d_best_lap, long long* d_best_idx, cudaStream_t st)
{
    long long blocks = (count + TOPK_BLOCK_KEYS - 1) / TOPK_BLOCK_KEYS;
    if (blocks < 1) blocks = 1;
    if (blocks > 1) {
        int rc = ensure_topk_scratch(ctx, blocks * (long long)k, st);
        if (rc != LTK_OK) return rc;
    }
    int pp = 0;
    while (true) {
        const bool single = (blocks == 1);
        const bool fused = !single && blocks * k <= TOPK_BLOCK_KEYS;  // last block merges in the same launch
        double* o_lap = (single || fused) ? d_best_lap : ctx->d_topk_lap[pp];
        long long* o_idx = (single || fused) ? d_best_idx : ctx->d_topk_idx[pp];
        topk_select<<<(unsigned)blocks, TOPK_THREADS, 0, st>>>(d_lap, d_idx, count, index_base, k, o_lap, o_idx,
                                                               ctx->d_topk_lap[pp], ctx->d_topk_idx[pp],
                                                               fused ? ctx->d_ticket : nullptr);
        g_launches.fetch_add(1);
        LTK_CUDA(ctx, cudaGetLastError());
        if (single || fused) break;
        d_lap = o_lap; d_idx = o_idx; count = blocks * k; index_base = 0;
        blocks = (count + TOPK_BLOCK_KEYS - 1) / TOPK_BLOCK_KEYS;
        pp ^= 1;
    }
    return LTK_OK;
}

}  // namespace

extern "C" {

#ifdef LTK_K1B_CLOCK
// developer build only (not in include/ltk.h): the K1b phase time stamps of the last launch
int ltk_debug_k1b_clocks(long long* h_out, long long count)
{
    return cudaMemcpyFromSymbol(h_out, g_k1b_clock, sizeof(long long) * count) == cudaSuccess ? 0 : -2;
}
#endif

int ltk_version(void) { return 100; }
int64_t ltk_launch_count(void) { return (int64_t)g_launches.load(); }

const char* ltk_last_error(const ltk_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

int ltk_create(ltk_ctx** out, int device, const double* h_left_xy, const double* h_diff_xy, int n_ctrl,
               const ltk_vehicle* vehicle, int ns)
{
    if (!out || !h_left_xy || !h_diff_xy) return fail(nullptr, LTK_E_ARG, "null argument");
    if (n_ctrl < 3) return fail(nullptr, LTK_E_ARG, "need at least 3 unique control points");
    if (ns < 3) return fail(nullptr, LTK_E_ARG, "ns must be >= 3");
    if (!check_vehicle(vehicle)) return fail(nullptr, LTK_E_ARG, "bad vehicle description");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return fail(nullptr, LTK_E_CUDA, "cudaGetDeviceCount", e);
    if (device < 0 || device >= ndev) return fail(nullptr, LTK_E_ARG, "no such CUDA device");
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, LTK_E_CUDA, "cudaGetDeviceProperties", e);
    if (prop.major != 10) return fail(nullptr, LTK_E_UNSUPPORTED, "ltk kernels are built for sm_100a (B200) only");

    ltk_ctx* ctx = new (std::nothrow) ltk_ctx;
    if (!ctx) return fail(nullptr, LTK_E_ARG, "out of host memory");
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->N = n_ctrl;
    ctx->ns = ns;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    ctx->veh = make_vehdev(*vehicle);
    ctx->veh32 = make_vehf32(*vehicle);
    ctx->sweep_bits = 64;
    ctx->spline_mode = LTK_SPLINE_TRIDIAGONAL;
    ctx->k1_g_override = 0;
    ctx->k1_staged_override = -1;
    if (const char* s = getenv("LTK_K1_G")) ctx->k1_g_override = atoi(s);
    if (const char* s = getenv("LTK_K1_STAGED")) ctx->k1_staged_override = atoi(s);
    ctx->k1_mode = 0;
    if (const char* s = getenv("LTK_K1")) ctx->k1_mode = (strcmp(s, "old") == 0) ? 1 : 0;
    ctx->sweep_mode = 0;
    if (const char* s = getenv("LTK_SWEEP")) {
        ctx->sweep_mode = (strcmp(s, "fused") == 0) ? 1 : (strcmp(s, "roles") == 0) ? 2 : 0;
    }
    ctx->k1_threads_override = 0;
    if (const char* s = getenv("LTK_K1_THREADS")) {
        int t = atoi(s);
        if (t == 64 || t == 128 || t == 256 || t == 512 || t == 1024) ctx->k1_threads_override = t;
    }
    int4 lut_cells[LTK_LUT_MAX_CELLS];
    const int n_cells = getenv("LTK_NO_ENGINE_LUT") ? 0 : build_engine_lut(*vehicle, &ctx->veh, lut_cells);
    if (n_cells > 0) {
        if ((e = cudaMalloc(&ctx->d_lut, sizeof(int4) * n_cells)) != cudaSuccess ||
            (e = cudaMemcpy(ctx->d_lut, lut_cells, sizeof(int4) * n_cells, cudaMemcpyHostToDevice)) != cudaSuccess) {
            fail(nullptr, LTK_E_CUDA, "engine table allocation", e);
            ltk_destroy(ctx);
            return LTK_E_CUDA;
        }
    }
    size_t bytes = sizeof(double) * 2 * (size_t)n_ctrl;
    if ((e = cudaMalloc(&ctx->d_left, bytes)) != cudaSuccess || (e = cudaMalloc(&ctx->d_diff, bytes)) != cudaSuccess ||
        (e = cudaMemcpy(ctx->d_left, h_left_xy, bytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(ctx->d_diff, h_diff_xy, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) {
        fail(nullptr, LTK_E_CUDA, "context allocation", e);
        ltk_destroy(ctx);
        return LTK_E_CUDA;
    }
    ctx->sweep_remainder = 1;
    if (const char* s = getenv("LTK_SWEEP_REMAINDER")) ctx->sweep_remainder = atoi(s);
    if (cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking) != cudaSuccess) ctx->aux_stream = nullptr;
    if (ctx->aux_stream) {
        cudaEventCreateWithFlags(&ctx->aux_ev[0], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->aux_ev[1], cudaEventDisableTiming);
    }
    K1Config cfg;
    K1FConfig fcfg0;
    if (!pick_k1f(ctx, &fcfg0) && !pick_k1(ctx, &cfg)) {
        fail(nullptr, LTK_E_UNSUPPORTED, "control-point count too large for shared memory");
        ltk_destroy(ctx);
        return LTK_E_UNSUPPORTED;
    }
    *out = ctx;
    return LTK_OK;
}

void ltk_destroy(ltk_ctx* ctx)
{
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    cudaFree(ctx->d_left);
    cudaFree(ctx->d_diff);
    cudaFree(ctx->d_lut);
    for (int i = 0; i < 2; ++i) { cudaFree(ctx->d_topk_lap[i]); cudaFree(ctx->d_topk_idx[i]); }
    cudaFree(ctx->d_ticket);
    if (ctx->trace_ev) {
        for (int i = 0; i < 2 * ctx->trace_cap; ++i) cudaEventDestroy(ctx->trace_ev[i]);
        free(ctx->trace_ev); free(ctx->trace_kind);
    }
    if (ctx->aux_stream) { cudaStreamDestroy(ctx->aux_stream); cudaEventDestroy(ctx->aux_ev[0]); cudaEventDestroy(ctx->aux_ev[1]); }
    cudaFree(ctx->d_profile_ws);
    for (int i = 0; i < LTK_HOST_GRAPHS; ++i)
        if (ctx->host_graph[i].exec) cudaGraphExecDestroy(ctx->host_graph[i].exec);
    if (ctx->host_stream) cudaStreamDestroy(ctx->host_stream);
    cudaFreeHost(ctx->h_in); cudaFreeHost(ctx->h_out);
    cudaFree(ctx->d_hin); cudaFree(ctx->d_hout); cudaFree(ctx->d_hws);
    delete ctx;
}

int ltk_set_ns(ltk_ctx* ctx, int ns)
{
    if (!ctx) return LTK_E_ARG;
    if (ns < 3) return fail(ctx, LTK_E_ARG, "ns must be >= 3");
    int old = ctx->ns;
    ctx->ns = ns;
    K1Config cfg;
    K1FConfig fcfg0;
    if (!pick_k1f(ctx, &fcfg0) && (ctx->spline_mode == LTK_SPLINE_FITPACK || !pick_k1(ctx, &cfg))) {
        ctx->ns = old;
        return fail(ctx, LTK_E_UNSUPPORTED, "no K1 configuration fits shared memory");
    }
    ++ctx->epoch;
    return LTK_OK;
}

int ltk_set_sweep_precision(ltk_ctx* ctx, int bits)
{
    if (!ctx) return LTK_E_ARG;
    if (bits != 64 && bits != 32) return fail(ctx, LTK_E_ARG, "sweep precision must be 64 or 32");
    ctx->sweep_bits = bits;
    ++ctx->epoch;
    return LTK_OK;
}

int ltk_set_spline_mode(ltk_ctx* ctx, int mode)
{
    if (!ctx) return LTK_E_ARG;
    if (mode != LTK_SPLINE_TRIDIAGONAL && mode != LTK_SPLINE_FITPACK) return fail(ctx, LTK_E_ARG, "unknown spline mode");
    if (mode == LTK_SPLINE_FITPACK && ctx->N < 5) return fail(ctx, LTK_E_UNSUPPORTED, "the FITPACK mode needs at least 5 unique control points");
    const int old = ctx->spline_mode;
    ctx->spline_mode = mode;
    K1FConfig fcfg0;
    if (mode == LTK_SPLINE_FITPACK && !pick_k1f(ctx, &fcfg0)) {
        ctx->spline_mode = old;
        return fail(ctx, LTK_E_UNSUPPORTED, "no FITPACK-mode K1 configuration fits shared memory");
    }
    ++ctx->epoch;
    return LTK_OK;
}

int ltk_spline_mode(const ltk_ctx* ctx) { return ctx ? ctx->spline_mode : LTK_E_ARG; }

int ltk_trace_begin(ltk_ctx* ctx, int max_records)
{
    if (!ctx || max_records < 0) return LTK_E_ARG;
    DeviceGuard guard(ctx->device);
    if (ctx->trace_ev) {
        for (int i = 0; i < 2 * ctx->trace_cap; ++i) cudaEventDestroy(ctx->trace_ev[i]);
        free(ctx->trace_ev); free(ctx->trace_kind);
        ctx->trace_ev = nullptr; ctx->trace_kind = nullptr; ctx->trace_cap = ctx->trace_n = 0;
    }
    if (max_records == 0) return LTK_OK;
    ctx->trace_ev = static_cast<cudaEvent_t*>(calloc(2 * (size_t)max_records, sizeof(cudaEvent_t)));
    ctx->trace_kind = static_cast<int*>(calloc((size_t)max_records, sizeof(int)));
    if (!ctx->trace_ev || !ctx->trace_kind) return fail(ctx, LTK_E_ARG, "out of host memory");
    for (int i = 0; i < 2 * max_records; ++i) LTK_CUDA(ctx, cudaEventCreate(&ctx->trace_ev[i]));
    ctx->trace_cap = max_records;
    ctx->trace_n = 0;
    return LTK_OK;
}

int ltk_trace_read(ltk_ctx* ctx, const ltk_ctx* base, int max_records, int* h_kind, float* h_start_ms, float* h_end_ms,
                   int* n_out)
{
    if (!ctx || !n_out || (max_records > 0 && (!h_kind || !h_start_ms || !h_end_ms))) return LTK_E_ARG;
    const ltk_ctx* b = base ? base : ctx;
    if (!ctx->trace_ev || !b->trace_ev || b->trace_n < 1) { *n_out = 0; return LTK_OK; }
    DeviceGuard guard(ctx->device);
    int n = ctx->trace_n < max_records ? ctx->trace_n : max_records;
    for (int i = 0; i < n; ++i) {
        LTK_CUDA(ctx, cudaEventSynchronize(ctx->trace_ev[2 * i + 1]));
        h_kind[i] = ctx->trace_kind[i];
        LTK_CUDA(ctx, cudaEventElapsedTime(&h_start_ms[i], b->trace_ev[0], ctx->trace_ev[2 * i]));
        LTK_CUDA(ctx, cudaEventElapsedTime(&h_end_ms[i], b->trace_ev[0], ctx->trace_ev[2 * i + 1]));
    }
    *n_out = n;
    return LTK_OK;
}

int ltk_set_sweep_split(ltk_ctx* ctx, int on)
{
    if (!ctx) return LTK_E_ARG;
    ctx->sweep_remainder = on ? 1 : 0;
    ++ctx->epoch;
    return LTK_OK;
}

int ltk_workspace_bytes(const ltk_ctx* ctx, int64_t B, size_t* out_bytes)
{
    if (!ctx || !out_bytes || B < 0) return LTK_E_ARG;
    *out_bytes = ctx_layout(ctx, B, false).total;
    return LTK_OK;
}

int ltk_eval_alphas(ltk_ctx* ctx, const double* d_alphas, int64_t B, double* d_lap, void* d_workspace,
                    size_t workspace_bytes, void* stream)
{
    if (!ctx) return LTK_E_ARG;
    if (B == 0) return LTK_OK;
    if (!d_alphas || !d_lap || !d_workspace || B < 0) return fail(ctx, LTK_E_ARG, "null or negative argument");
    WsLayout w = ctx_layout(ctx, B, false);
    if (workspace_bytes < w.total) return fail(ctx, LTK_E_WORKSPACE, "workspace too small (see ltk_workspace_bytes)");
    DeviceGuard guard(ctx->device);
    return run_pipeline(ctx, d_alphas, nullptr, 0, B, d_lap, static_cast<char*>(d_workspace), w, false,
                        static_cast<cudaStream_t>(stream));
}

int ltk_eval_alphas_topk(ltk_ctx* ctx, const double* d_alphas, int64_t B, double* d_lap, void* d_workspace,
                         size_t workspace_bytes, int64_t index_base, int k, double* d_best_lap, int64_t* d_best_idx,
                         void* stream)
{
    if (!ctx) return LTK_E_ARG;
    if (!d_alphas || !d_lap || !d_workspace || !d_best_lap || !d_best_idx || B < 1)
        return fail(ctx, LTK_E_ARG, "null argument or empty population");
    if (k < 1 || k > TOPK_MAX) return fail(ctx, LTK_E_ARG, "k must be in 1..64");
    WsLayout w = ctx_layout(ctx, B, false);
    if (workspace_bytes < w.total) return fail(ctx, LTK_E_WORKSPACE, "workspace too small (see ltk_workspace_bytes)");
    DeviceGuard guard(ctx->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TopkReq req{k, (long long)index_base, d_best_lap, reinterpret_cast<long long*>(d_best_idx), false};
    int rc = run_pipeline(ctx, d_alphas, nullptr, 0, B, d_lap, static_cast<char*>(d_workspace), w, false, st, nullptr,
                          false, &req);
    if (rc != LTK_OK || req.fused) return rc;
    return run_topk(ctx, d_lap, nullptr, B, index_base, k, d_best_lap, reinterpret_cast<long long*>(d_best_idx), st);
}

// Host in, host out, one call: the latency path of the optimiser loops (finite-difference rounds of
// N + 1 candidates, lock-step COBYLA rounds of one candidate per start).  The
// upload, the three kernels and the download of one batch size are ONE instantiated CUDA graph, launched
// on the context's own stream; the caller's arrays are ordinary (pageable) host memory.
static int host_reserve(ltk_ctx* ctx, long long B)
{
    if (!ctx->host_stream) LTK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->host_stream, cudaStreamNonBlocking));
    const WsLayout w = ctx_layout(ctx, B, false);
    if (B <= ctx->host_cap && w.total <= ctx->host_ws_bytes) return LTK_OK;
    LTK_CUDA(ctx, cudaStreamSynchronize(ctx->host_stream));
    for (int i = 0; i < LTK_HOST_GRAPHS; ++i)  // the graphs hold the old addresses
        if (ctx->host_graph[i].exec) { cudaGraphExecDestroy(ctx->host_graph[i].exec); ctx->host_graph[i].exec = nullptr; }
    if (B > ctx->host_cap) {
        long long cap = ctx->host_cap ? ctx->host_cap : 64;
        while (cap < B) cap *= 2;
        cudaFreeHost(ctx->h_in); cudaFreeHost(ctx->h_out); cudaFree(ctx->d_hin); cudaFree(ctx->d_hout);
        ctx->h_in = ctx->h_out = ctx->d_hin = ctx->d_hout = nullptr;
        ctx->host_cap = 0;
        const size_t in_bytes = sizeof(double) * (size_t)cap * (size_t)ctx->N;
        LTK_CUDA(ctx, cudaMallocHost(&ctx->h_in, in_bytes));
        LTK_CUDA(ctx, cudaMallocHost(&ctx->h_out, sizeof(double) * (size_t)cap));
        LTK_CUDA(ctx, cudaMalloc(&ctx->d_hin, in_bytes));
        LTK_CUDA(ctx, cudaMalloc(&ctx->d_hout, sizeof(double) * (size_t)cap));
        ctx->host_cap = cap;
    }
    const size_t need = ctx_layout(ctx, ctx->host_cap, false).total;
    if (need > ctx->host_ws_bytes) {
        cudaFree(ctx->d_hws);
        ctx->d_hws = nullptr; ctx->host_ws_bytes = 0;
        LTK_CUDA(ctx, cudaMalloc(&ctx->d_hws, need));
        ctx->host_ws_bytes = need;
    }
    return LTK_OK;
}

static int host_enqueue(ltk_ctx* ctx, long long B)
{
    const WsLayout w = ctx_layout(ctx, B, false);
    cudaStream_t st = ctx->host_stream;
    LTK_CUDA(ctx, cudaMemcpyAsync(ctx->d_hin, ctx->h_in, sizeof(double) * (size_t)B * (size_t)ctx->N, cudaMemcpyHostToDevice, st));
    int rc = run_pipeline(ctx, ctx->d_hin, nullptr, 0, B, ctx->d_hout, ctx->d_hws, w, false, st);
    if (rc != LTK_OK) return rc;
    LTK_CUDA(ctx, cudaMemcpyAsync(ctx->h_out, ctx->d_hout, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, st));
    return LTK_OK;
}

int ltk_eval_alphas_host(ltk_ctx* ctx, const double* h_alphas, int64_t B, double* h_lap)
{
    if (!ctx) return LTK_E_ARG;
    if (B == 0) return LTK_OK;
    if (!h_alphas || !h_lap || B < 0) return fail(ctx, LTK_E_ARG, "null or negative argument");
    DeviceGuard guard(ctx->device);
    int rc = host_reserve(ctx, B);
    if (rc != LTK_OK) return rc;
    memcpy(ctx->h_in, h_alphas, sizeof(double) * (size_t)B * (size_t)ctx->N);
    const bool use_graph = B <= LTK_HOST_GRAPH_MAX_B && !ctx->trace_ev && !getenv("LTK_NO_GRAPH");
    if (use_graph) {
        int slot = -1, victim = 0;
        for (int i = 0; i < LTK_HOST_GRAPHS; ++i) {
            if (ctx->host_graph[i].exec && ctx->host_graph[i].B == B && ctx->host_graph[i].epoch == ctx->epoch) { slot = i; break; }
            if (ctx->host_used[i] < ctx->host_used[victim]) victim = i;
        }
        if (slot < 0) {
            slot = victim;
            if (ctx->host_graph[slot].exec) { cudaGraphExecDestroy(ctx->host_graph[slot].exec); ctx->host_graph[slot].exec = nullptr; }
            cudaGraph_t graph = nullptr;
            LTK_CUDA(ctx, cudaStreamBeginCapture(ctx->host_stream, cudaStreamCaptureModeThreadLocal));
            const long long before = g_launches.load();
            rc = host_enqueue(ctx, B);
            cudaError_t e = cudaStreamEndCapture(ctx->host_stream, &graph);
            g_launches.store(before);  // captured, not launched
            if (rc != LTK_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail(ctx, LTK_E_CUDA, "cudaStreamEndCapture", e);
            e = cudaGraphInstantiate(&ctx->host_graph[slot].exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { ctx->host_graph[slot].exec = nullptr; return fail(ctx, LTK_E_CUDA, "cudaGraphInstantiate", e); }
            ctx->host_graph[slot].B = B;
            ctx->host_graph[slot].epoch = ctx->epoch;
        }
        ctx->host_used[slot] = ++ctx->host_tick;
        LTK_CUDA(ctx, cudaGraphLaunch(ctx->host_graph[slot].exec, ctx->host_stream));
        g_launches.fetch_add(3);  // K1a, K1b, K23 inside the graph
    } else {
        rc = host_enqueue(ctx, B);
        if (rc != LTK_OK) return rc;
    }
    LTK_CUDA(ctx, cudaStreamSynchronize(ctx->host_stream));
    memcpy(h_lap, ctx->h_out, sizeof(double) * (size_t)B);
    return LTK_OK;
}

int ltk_eval_alphas_timed(ltk_ctx* ctx, const double* d_alphas, int64_t B, double* d_lap, void* d_workspace,
                          size_t workspace_bytes, void* stream, float* h_ms)
{
    if (!ctx) return LTK_E_ARG;
    if (!d_alphas || !d_lap || !d_workspace || !h_ms || B < 1) return fail(ctx, LTK_E_ARG, "null or non-positive argument");
    WsLayout w = ctx_layout(ctx, B, false);
    if (workspace_bytes < w.total) return fail(ctx, LTK_E_WORKSPACE, "workspace too small (see ltk_workspace_bytes)");
    DeviceGuard guard(ctx->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaEvent_t ev[5];
    for (int i = 0; i < 5; ++i) LTK_CUDA(ctx, cudaEventCreate(&ev[i]));
    int rc = run_pipeline(ctx, d_alphas, nullptr, 0, B, d_lap, static_cast<char*>(d_workspace), w, false, st, ev);
    if (rc == LTK_OK) {
        cudaError_t e = cudaEventSynchronize(ev[4]);
        for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventElapsedTime(&h_ms[i], ev[i], ev[i + 1]);
        if (e != cudaSuccess) rc = fail(ctx, LTK_E_CUDA, "event timing", e);
        h_ms[3] = -1.0f;  // one sweep kernel: no separate backward launch
    }
    for (int i = 0; i < 5; ++i) cudaEventDestroy(ev[i]);
    return rc;
}

int ltk_eval_objectives(ltk_ctx* ctx, const double* d_alphas, int64_t B, double* d_gamma2, double* d_length,
                        void* d_workspace, size_t workspace_bytes, void* stream)
{
    if (!ctx) return LTK_E_ARG;
    if (B == 0) return LTK_OK;
    if (!d_alphas || (!d_gamma2 && !d_length) || !d_workspace || B < 0) return fail(ctx, LTK_E_ARG, "null or negative argument");
    WsLayout w = ctx_layout(ctx, B, false);
    if (workspace_bytes < w.total) return fail(ctx, LTK_E_WORKSPACE, "workspace too small (see ltk_workspace_bytes)");
    DeviceGuard guard(ctx->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(d_workspace);
    int rc = run_pipeline(ctx, d_alphas, nullptr, 0, B, nullptr, ws, w, false, st, nullptr, true);
    if (rc != LTK_OK) return rc;
    curvature_objectives<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(
        reinterpret_cast<double*>(ws + w.kap_off), reinterpret_cast<int*>(ws + w.rot_off),
        reinterpret_cast<double*>(ws + w.len_off), ctx->ns, B, d_gamma2, d_length);
    g_launches.fetch_add(1);
    LTK_CUDA(ctx, cudaGetLastError());
    return LTK_OK;
}

int ltk_topk_pairs(ltk_ctx* ctx, const double* d_lap, const int64_t* d_idx, int64_t count, int k, double* d_best_lap,
                   int64_t* d_best_idx, void* stream)
{
    if (!ctx) return LTK_E_ARG;
    if (!d_lap || !d_idx || !d_best_lap || !d_best_idx || count < 0) return fail(ctx, LTK_E_ARG, "null or negative argument");
    if (k < 1 || k > TOPK_MAX) return fail(ctx, LTK_E_ARG, "k must be in 1..64");
    DeviceGuard guard(ctx->device);
    return run_topk(ctx, d_lap, reinterpret_cast<const long long*>(d_idx), count, 0, k, d_best_lap,
                    reinterpret_cast<long long*>(d_best_idx), static_cast<cudaStream_t>(stream));
}

int ltk_topk_gathered(ltk_ctx* ctx, const int64_t* d_gathered, int world, int k_in, int k, double* d_best_lap,
                      int64_t* d_best_idx, void* stream)
{
    if (!ctx) return LTK_E_ARG;
    if (!d_gathered || !d_best_lap || !d_best_idx || world < 1 || k_in < 1) return fail(ctx, LTK_E_ARG, "null or non-positive argument");
    if (k < 1 || k > TOPK_MAX) return fail(ctx, LTK_E_ARG, "k must be in 1..64");
    if ((long long)world * k_in > TOPK_BLOCK_KEYS) return fail(ctx, LTK_E_UNSUPPORTED, "world * k_in exceeds one merge block (1,024 keys)");
    DeviceGuard guard(ctx->device);
    topk_merge_gathered<<<1, TOPK_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(d_gathered), world, k_in, k, d_best_lap, reinterpret_cast<long long*>(d_best_idx));
    g_launches.fetch_add(1);
    LTK_CUDA(ctx, cudaGetLastError());
    return LTK_OK;
}

int ltk_eval_controls(ltk_ctx* ctx, const double* d_xy, int m, int64_t B, double* d_lap, void* d_workspace,
                      size_t workspace_bytes, void* stream)
{
    if (!ctx) return LTK_E_ARG;
    if (B == 0) return LTK_OK;
    if (!d_xy || !d_lap || !d_workspace || B < 0) return fail(ctx, LTK_E_ARG, "null or negative argument");
    if (m != ctx->N + 1) return fail(ctx, LTK_E_ARG, "controls must have n_ctrl + 1 columns");
    WsLayout w = ctx_layout(ctx, B, false);
    if (workspace_bytes < w.total) return fail(ctx, LTK_E_WORKSPACE, "workspace too small (see ltk_workspace_bytes)");
    DeviceGuard guard(ctx->device);
    return run_pipeline(ctx, nullptr, d_xy, m, B, d_lap, static_cast<char*>(d_workspace), w, false,
                        static_cast<cudaStream_t>(stream));
}

int ltk_profile(ltk_ctx* ctx, const double* d_alpha, double* d_s, double* d_k, double* d_vlocal, double* d_vacc,
                double* d_vdec, double* d_v, double* d_scalars, void* stream)
{
    if (!ctx) return LTK_E_ARG;
    if (!d_alpha) return fail(ctx, LTK_E_ARG, "null alpha");
    DeviceGuard guard(ctx->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    WsLayout w = ws_layout(ctx->ns, ctx->N, 1, true);
    size_t need = w.total + 256;
    if (ctx->profile_ws_bytes < need) {
        LTK_CUDA(ctx, cudaStreamSynchronize(st));
        cudaFree(ctx->d_profile_ws);
        ctx->d_profile_ws = nullptr;
        ctx->profile_ws_bytes = 0;
        LTK_CUDA(ctx, cudaMalloc(&ctx->d_profile_ws, need));
        ctx->profile_ws_bytes = need;
    }
    char* ws = static_cast<char*>(ctx->d_profile_ws);
    double* d_lap = reinterpret_cast<double*>(ws + w.total);
    int rc = run_pipeline(ctx, d_alpha, nullptr, 0, 1, d_lap, ws, w, true, st);
    if (rc != LTK_OK) return rc;
    unrotate_profile<<<8, 256, 0, st>>>(reinterpret_cast<double*>(ws + w.kap_off), reinterpret_cast<double*>(ws + w.vaccd_off),
                                        reinterpret_cast<double*>(ws + w.vdec_off), reinterpret_cast<double*>(ws + w.vmin_off),
                                        reinterpret_cast<int*>(ws + w.rot_off), reinterpret_cast<double*>(ws + w.len_off),
                                        ctx->ns, w.Bp, ctx->veh.mu_g, d_s, d_k, d_vlocal, d_vacc, d_vdec, d_v);
    g_launches.fetch_add(1);
    LTK_CUDA(ctx, cudaGetLastError());
    if (d_scalars) {
        LTK_CUDA(ctx, cudaMemcpyAsync(d_scalars, d_lap, sizeof(double), cudaMemcpyDeviceToDevice, st));
        LTK_CUDA(ctx, cudaMemcpyAsync(d_scalars + 1, ws + w.len_off, sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    LTK_CUDA(ctx, cudaStreamSynchronize(st));
    return LTK_OK;
}

int ltk_topk(ltk_ctx* ctx, const double* d_lap, int64_t B, int64_t index_base, int k, double* d_best_lap,
             int64_t* d_best_idx, void* stream)
{
    if (!ctx) return LTK_E_ARG;
    if (!d_lap || !d_best_lap || !d_best_idx || B < 0) return fail(ctx, LTK_E_ARG, "null or negative argument");
    if (k < 1 || k > TOPK_MAX) return fail(ctx, LTK_E_ARG, "k must be in 1..64");
    DeviceGuard guard(ctx->device);
    return run_topk(ctx, d_lap, nullptr, B, index_base, k, d_best_lap, reinterpret_cast<long long*>(d_best_idx),
                    static_cast<cudaStream_t>(stream));
}

int ltk_random_uniform(int device, uint64_t key0, uint64_t key1, int64_t first, int64_t count, double low, double high,
                       double* d_out, void* stream)
{
    if (count == 0) return LTK_OK;
    if (!d_out || first < 0 || count < 0) return fail(nullptr, LTK_E_ARG, "null or negative argument");
    DeviceGuard guard(device);
    const long long nblk = (first + count + 3) / 4 - first / 4;
    long long grid = (nblk + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    philox_uniform<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(key0, key1, first, count, low, high - low, d_out);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nullptr, LTK_E_CUDA, "philox_uniform", e);
    return LTK_OK;
}

int ltk_path_eval(int device, const double* d_xy, const double* d_knots, int m, const double* d_u, int64_t n,
                  double* d_x, double* d_y, double* d_dx, double* d_dy, double* d_ddx, double* d_ddy,
                  double* d_k_signed, double* d_gamma2, void* stream)
{
    if (!d_xy || !d_knots || (!d_u && n > 0) || n < 0) return fail(nullptr, LTK_E_ARG, "null or negative argument");
    if (m < 4) return fail(nullptr, LTK_E_ARG, "a closed path needs at least 3 unique points");
    DeviceGuard guard(device);
    int N = m - 1;
    size_t smem = sizeof(double) * (size_t)(11 * N + N + 1);
    cudaError_t e = cudaFuncSetAttribute(path_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(nullptr, LTK_E_UNSUPPORTED, "path too long for shared memory", e);
    PathArgs a{d_xy, d_knots, m, d_u, n, d_x, d_y, d_dx, d_dy, d_ddx, d_ddy, d_k_signed, d_gamma2};
    path_eval_kernel<<<1, 256, smem, static_cast<cudaStream_t>(stream)>>>(a);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nullptr, LTK_E_CUDA, "path_eval_kernel", e);
    return LTK_OK;
}

int ltk_path_eval_fitpack(int device, const double* d_xy, const double* d_knots, int m, int closed, const double* d_u,
                          int64_t n, double* d_x, double* d_y, double* d_dx, double* d_dy, double* d_ddx, double* d_ddy,
                          double* d_k_signed, double* d_gamma2, double* d_t, double* d_c, void* stream)
{
    if (!d_xy || !d_knots || (!d_u && n > 0) || n < 0) return fail(nullptr, LTK_E_ARG, "null or negative argument");
    if (closed ? m < 6 : m < 4)
        return fail(nullptr, LTK_E_UNSUPPORTED, "FITPACK arithmetic needs 5 unique points (closed) or 4 points (open)");
    DeviceGuard guard(device);
    size_t smem = sizeof(double) * path_fit_smem_doubles(m, closed);
    cudaError_t e = cudaFuncSetAttribute(path_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(nullptr, LTK_E_UNSUPPORTED, "path too long for shared memory", e);
    PathFitArgs a{d_xy, d_knots, m, closed ? 1 : 0, d_u, n, d_x, d_y, d_dx, d_dy, d_ddx, d_ddy, d_k_signed, d_gamma2, d_t, d_c};
    path_fit_kernel<<<1, 256, smem, static_cast<cudaStream_t>(stream)>>>(a);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nullptr, LTK_E_CUDA, "path_fit_kernel", e);
    return LTK_OK;
}

int ltk_velocity_profile(int device, const ltk_vehicle* vehicle, const double* d_s, const double* d_k, int64_t n,
                         double s_max, double* d_vlocal, double* d_vacc, double* d_vdec, double* d_v, void* stream)
{
    if (!check_vehicle(vehicle)) return fail(nullptr, LTK_E_ARG, "bad vehicle description");
    if (!d_s || !d_k || n < 1) return fail(nullptr, LTK_E_ARG, "null or empty samples");
    if (!d_vacc || !d_vdec) return fail(nullptr, LTK_E_ARG, "d_vacc and d_vdec are required (they are the sweep state)");
    DeviceGuard guard(device);
    VehDev V = make_vehdev(*vehicle);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (V.kind == 0) velocity_profile_kernel<0><<<1, 32, 0, st>>>(V, d_s, d_k, n, s_max, d_vlocal, d_vacc, d_vdec, d_v);
    else velocity_profile_kernel<1><<<1, 32, 0, st>>>(V, d_s, d_k, n, s_max, d_vlocal, d_vacc, d_vdec, d_v);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nullptr, LTK_E_CUDA, "velocity_profile_kernel", e);
    return LTK_OK;
}

}  // extern "C"
