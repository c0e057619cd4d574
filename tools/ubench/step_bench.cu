// step_bench.cu -- the sweep recurrences without memory traffic: cycles per step of one forward /
// backward chain (single warp = latency; 7 warps per scheduler = throughput), and of the bare
// unguarded sqrt / div sequences.  Build: nvcc -O3 -fmad=false -I../../include -I../../lap_time_optimization_b200/csrc
#include <cstdio>
#include <cstring>
#include "ltk_kernels.cuh"
using namespace ltk;

template <int MODE>
__global__ void bench(VehDev V, const int4* lut, double* out, int iters, long long* cyc, double k_in, double ds)
{
    __shared__ FusedShared S;
    for (int i = threadIdx.x; i <= LTK_MAX_ENGINE_MAP; i += blockDim.x) {
        S.seg[i].s = V.ext_s[i]; S.seg[i].b = V.ext_b[i]; S.seg[i].f = V.ext_f[i]; S.seg[i].pad = 0.0;
    }
    for (int i = threadIdx.x; i <= V.lut_top; i += blockDim.x) S.cell[i] = lut[i];
    __syncthreads();
    double v = 12.0 + 0.01 * (threadIdx.x & 31), k = k_in * (1.0 + 1e-3 * (threadIdx.x & 7));
    double acc = 0.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // forward step incl. the local-limit division
            double wl = ddiv<false>(V.mu_g, k);
            v = forward_fast<0, 0>(V, S, v, k, wl, ds);
        } else if (MODE == 1) {  // backward step
            double wl = ddiv<false>(V.mu_g, k);
            v = backward_fast<0>(V, v, k, wl, ds);
        } else if (MODE == 2) {  // bare sqrt chain
            v = dsqrt<false>(v + ds);
        } else if (MODE == 3) {  // bare division chain
            v = ddiv<false>(V.mu_g, v) ;
        } else if (MODE == 4) {  // forward step + lap term (phase 2)
            double wl = ddiv<false>(V.mu_g, k);
            v = forward_fast<0, 0>(V, S, v, k, wl, ds);
            acc = acc + ddiv<false>(ds, v);
        } else if (MODE == 5) {  // engine lookup only
            v = engine_fast<0>(V, S, v) * 1e-3 + 10.0;
        }
        k = k + 1e-9;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = v + acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// host copies of make_vehdev / build_engine_lut live in ltk_api.cu; rebuild the TBR18 tables here
static void tbr18(VehDev* d, int4* cells)
{
    memset(d, 0, sizeof(*d));
    const double mv[7] = {5, 10, 15, 20, 25, 30, 35}, mf[7] = {5000, 4700, 3500, 2800, 2300, 1900, 1600};
    d->kind = 0; d->n_map = 7; d->mass = 200; d->half_mass = 100; d->inv_half_mass = 0.01;
    d->mu_g = 1.5 * 9.81; d->f_max = 1.5 * 200 * 9.81; d->f_max_sq = d->f_max * d->f_max;
    d->ext_b[0] = mv[0]; d->ext_f[0] = mf[0]; d->ext_s[0] = 0;
    for (int j = 1; j < 7; ++j) { d->ext_b[j] = mv[j - 1]; d->ext_f[j] = mf[j - 1]; d->ext_s[j] = (mf[j] - mf[j - 1]) / (mv[j] - mv[j - 1]); }
    d->ext_b[7] = mv[6]; d->ext_f[7] = mf[6]; d->ext_s[7] = 0;
    int hi[7]; long long bits[7];
    for (int i = 0; i < 7; ++i) { memcpy(&bits[i], &mv[i], 8); hi[i] = (int)(bits[i] >> 32); }
    const int shift = 18, first = hi[0] >> shift, last = hi[6] >> shift, ncell = last - first + 3, base = first - 1;
    int node = 0;
    for (int c = 0; c < ncell; ++c) {
        long long thr = 0x7fffffffffffffffLL;
        if (node < 7 && (hi[node] >> shift) == base + c && c > 0 && c < ncell - 1) thr = bits[node];
        cells[c].x = (int)(unsigned)(thr & 0xffffffffLL); cells[c].y = (int)(thr >> 32); cells[c].z = node; cells[c].w = 0;
        if (thr != 0x7fffffffffffffffLL) ++node;
    }
    d->lut_shift = shift; d->lut_base = base; d->lut_top = ncell - 1;
    d->k_lo_hi = 0x20000000; d->k_span_hi = 0x40000000u;
}

template <int MODE>
void run(const char* name, const VehDev& V, const int4* d_lut, int warps_per_sm)
{
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    for (int r = 0; r < 2; ++r) { bench<MODE><<<148, 32 * warps_per_sm>>>(V, d_lut, out, iters, cyc, 0.02, 1.0); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-22s warps/SMSP=%.2f : %8.1f cycles per step per warp, %7.1f SMSP-cycles per step\n", name, warps_per_sm / 4.0,
           (double)h / iters, (double)h / iters / (warps_per_sm / 4.0));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    VehDev V; int4 cells[LTK_LUT_MAX_CELLS]; tbr18(&V, cells);
    int4* d_lut; cudaMalloc(&d_lut, sizeof(cells)); cudaMemcpy(d_lut, cells, sizeof(cells), cudaMemcpyHostToDevice);
    const int ws[5] = {4, 8, 16, 28, 32};
    for (int i = 0; i < 5; ++i) run<0>("forward step", V, d_lut, ws[i]);
    for (int i = 0; i < 5; ++i) run<1>("backward step", V, d_lut, ws[i]);
    for (int i = 0; i < 5; ++i) run<4>("forward step + lap", V, d_lut, ws[i]);
    for (int i = 0; i < 5; ++i) run<2>("dsqrt chain", V, d_lut, ws[i]);
    for (int i = 0; i < 5; ++i) run<3>("ddiv chain", V, d_lut, ws[i]);
    for (int i = 0; i < 5; ++i) run<5>("engine lookup", V, d_lut, ws[i]);
    return 0;
}
