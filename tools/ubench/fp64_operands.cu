// fp64_operands.cu -- FP64 pipe occupancy per DFMA as a function of how many DISTINCT vector registers it reads.
#include <cstdio>
#include <cuda_runtime.h>
#define CH 8
template <int OP>
__global__ void k(double* out, const double* in, int iters, long long* cyc, double ua, double ub)
{
    double x[CH], y[CH], z[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { x[c] = in[c] + threadIdx.x; y[c] = in[CH + c] + 1e-9 * threadIdx.x; z[c] = in[2 * CH + c] + 1e-9 * threadIdx.x; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (OP == 0) x[c] = fma(x[c], y[c], z[c]);   // 3 distinct vector registers
            if (OP == 1) x[c] = fma(x[c], x[c], y[c]);   // 2 distinct (a == b)
            if (OP == 2) x[c] = fma(x[c], y[c], y[c]);   // 2 distinct (b == c)
            if (OP == 3) x[c] = fma(x[c], y[c], x[c]);   // 2 distinct (a == c)
            if (OP == 4) x[c] = fma(x[c], y[c], 0.5);    // 2 vector + immediate
            if (OP == 5) x[c] = fma(x[c], y[c], ub);     // 2 vector + uniform/constant
            if (OP == 6) x[c] = fma(x[c], ua, ub);       // 1 vector
            if (OP == 7) x[c] = x[c] * y[c];             // DMUL 2 vector
            if (OP == 8) x[c] = fma(y[c], z[c], x[c]);   // 3 distinct, accumulate form
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c] + y[c] + z[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name)
{
    double *out, *in; long long* cyc; long long h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&cyc, 8); cudaMalloc(&in, 8 * 64);
    double hin[64]; for (int i = 0; i < 64; ++i) hin[i] = 0.5 + 0.001 * i;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    const int iters = 2048;
    for (int r = 0; r < 2; ++r) { k<OP><<<148, 512>>>(out, in, iters, cyc, 0.999999, 1e-7); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s : %.2f SMSP-cycles per instruction\n", name, (double)h / iters / (CH * 4));
}
int main()
{
    run<0>("DFMA x=fma(x,y,z) 3 distinct");
    run<8>("DFMA x=fma(y,z,x) 3 distinct");
    run<1>("DFMA x=fma(x,x,y) 2 distinct");
    run<2>("DFMA x=fma(x,y,y) 2 distinct");
    run<3>("DFMA x=fma(x,y,x) 2 distinct");
    run<4>("DFMA x=fma(x,y,imm)");
    run<5>("DFMA x=fma(x,y,uniform)");
    run<6>("DFMA x=fma(x,uniform,uniform)");
    run<7>("DMUL x=x*y");
    return 0;
}
