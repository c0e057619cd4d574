// fp64_lat.cu -- B200 FP64 pipe micro-benchmarks: dependent-issue latency and throughput of the
// instructions the sweep kernels are made of.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ double op(double x, double a, double b)
{
    if (OP == 0) return fma(x, a, b);                       // DFMA
    if (OP == 1) return x * a;                              // DMUL
    if (OP == 2) return x + a;                              // DADD
    if (OP == 3) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }  // MUFU.RSQ64H
    if (OP == 4) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
    if (OP == 5) return (x < a) ? x : b;                    // DSETP + 2 FSEL
    if (OP == 6) return sqrt(x);                            // library sqrt
    if (OP == 7) return a / x;                              // library div
    if (OP == 8) return (double)(__double2hiint(x) & 1023) * a + b;  // I2F + DFMA
    return x;
}

// CH independent dependent chains per thread, ITER steps
template <int OP, int CH>
__global__ void chain(double* out, double a, double b, int iters, long long* cyc)
{
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = 1.0 + 1e-3 * (threadIdx.x + c);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) x[c] = op<OP>(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int CH>
void run(const char* name, int warps_per_sm, double a, double b)
{
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    chain<OP, CH><<<148, 32 * warps_per_sm>>>(out, a, b, iters, cyc);
    cudaDeviceSynchronize();
    chain<OP, CH><<<148, 32 * warps_per_sm>>>(out, a, b, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_step = (double)h / iters;  // cycles per step of CH ops per warp
    double warps_per_smsp = warps_per_sm / 4.0;
    printf("%-10s chains/thread %d warps/SMSP %.2f : %.2f cyc/step  -> %.3f warp-instr/cyc/SMSP\n", name, CH,
           warps_per_smsp, per_step, CH * warps_per_smsp / per_step);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    printf("== dependent latency (1 chain, 1 warp/SMSP) ==\n");
    run<0, 1>("DFMA", 4, 0.999999, 1e-7);
    run<1, 1>("DMUL", 4, 0.999999, 0);
    run<2, 1>("DADD", 4, 1e-9, 0);
    run<3, 1>("RSQ64H", 4, 0, 0);
    run<4, 1>("RCP64H", 4, 0, 0);
    run<5, 1>("DSETP+SEL", 4, 2.0, 1.0);
    run<6, 1>("sqrt", 4, 0, 0);
    run<7, 1>("div", 4, 1.5, 0);
    run<8, 1>("I2F+DFMA", 4, 1e-3, 1.0);
    printf("== DFMA throughput vs independent chains ==\n");
    run<0, 2>("DFMA", 4, 0.999999, 1e-7);
    run<0, 4>("DFMA", 4, 0.999999, 1e-7);
    run<0, 8>("DFMA", 4, 0.999999, 1e-7);
    run<0, 1>("DFMA", 8, 0.999999, 1e-7);
    run<0, 1>("DFMA", 16, 0.999999, 1e-7);
    run<0, 2>("DFMA", 16, 0.999999, 1e-7);
    run<0, 4>("DFMA", 16, 0.999999, 1e-7);
    run<0, 8>("DFMA", 32, 0.999999, 1e-7);
    printf("== MUFU / mixed throughput ==\n");
    run<3, 4>("RSQ64H", 16, 0, 0);
    run<3, 8>("RSQ64H", 32, 0, 0);
    run<5, 4>("DSETP+SEL", 16, 2.0, 1.0);
    run<6, 4>("sqrt", 16, 0, 0);
    run<7, 4>("div", 16, 1.5, 0);
    run<8, 4>("I2F+DFMA", 16, 1e-3, 1.0);
    run<6, 2>("sqrt", 14, 0, 0);
    return 0;
}
