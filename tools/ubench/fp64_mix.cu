// fp64_mix.cu -- do the XU (MUFU.RSQ64H / I2F.F64) and ALU (FSEL, ISETP) pipes overlap with the FP64 pipe on B200?
// Each thread runs CH independent chains; one step of a chain = NF dependent DFMAs + NX "other" ops of kind OP.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ double other(double x, double a)
{
    if (OP == 0) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
    if (OP == 1) return (double)(__double2loint(x) & 1023);                 // I2F.F64
    if (OP == 2) return __hiloint2double(__double2hiint(x) + 1, __double2loint(x));  // ALU (VIADD)
    if (OP == 3) return (__double2hiint(x) > 0x3ff80000) ? a : x;           // ISETP + 2 SEL
    if (OP == 4) return (x < a) ? x : a;                                    // DSETP + 2 FSEL
    if (OP == 5) { float f = __int_as_float(__double2hiint(x)); f = fmaf(f, 1.0001f, 0.5f); return __hiloint2double(__float_as_int(f), __double2loint(x)); } // FFMA
    return x;
}

template <int OP, int NF, int NX, int CH>
__global__ void mix(double* out, double a, double b, int iters, long long* cyc)
{
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = 1.0 + 1e-3 * (threadIdx.x + c);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
#pragma unroll
            for (int k = 0; k < NF; ++k) x[c] = fma(x[c], a, b);
#pragma unroll
            for (int k = 0; k < NX; ++k) x[c] = other<OP>(x[c], a);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int NF, int NX, int CH>
void run(const char* name, int warps_per_sm)
{
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    for (int r = 0; r < 2; ++r) { mix<OP, NF, NX, CH><<<148, 32 * warps_per_sm>>>(out, 0.999999, 1e-7, iters, cyc); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double w = warps_per_sm / 4.0;
    double per = (double)h / iters / (CH * w);  // SMSP cycles per chain-step
    printf("%-8s NF=%d NX=%d CH=%d warps/SMSP=%.1f : %.2f SMSP-cycles per step (FP64 alone would be %d)\n", name, NF, NX, CH, w, per, 2 * NF);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0, 8, 0, 4>("dfma", 16);
    run<0, 8, 1, 4>("rsq64h", 16);
    run<0, 8, 2, 4>("rsq64h", 16);
    run<0, 0, 1, 4>("rsq64h", 16);
    run<1, 8, 1, 4>("i2f", 16);
    run<1, 8, 2, 4>("i2f", 16);
    run<1, 0, 1, 4>("i2f", 16);
    run<2, 8, 4, 4>("viadd", 16);
    run<2, 8, 8, 4>("viadd", 16);
    run<2, 0, 8, 4>("viadd", 16);
    run<3, 8, 2, 4>("isetp+sel", 16);
    run<3, 8, 4, 4>("isetp+sel", 16);
    run<3, 0, 4, 4>("isetp+sel", 16);
    run<4, 8, 2, 4>("dsetp+sel", 16);
    run<4, 0, 2, 4>("dsetp+sel", 16);
    run<5, 8, 4, 4>("ffma", 16);
    run<5, 8, 8, 4>("ffma", 16);
    run<5, 0, 8, 4>("ffma", 16);
    return 0;
}
