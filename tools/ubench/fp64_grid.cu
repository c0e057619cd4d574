// fp64_ops.cu -- does a DFMA with three distinct register operands still issue every 2 cycles on B200?
// (register-file read bandwidth), and DMUL / DADD / DSETP rates.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int CH>
__global__ void k(double* out, const double* in, int iters, long long* cyc)
{
    double x[CH], y[CH], z[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { x[c] = in[c] + threadIdx.x; y[c] = in[CH + c] * 0.999999; z[c] = in[2 * CH + c] * 1e-7; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (OP == 0) x[c] = fma(x[c], y[c], z[c]);                 // 3 distinct register operands
            if (OP == 1) x[c] = fma(x[c], y[0], z[0]);                 // shared multiplier/addend registers
            if (OP == 2) x[c] = x[c] * y[c];
            if (OP == 3) x[c] = x[c] + z[c];
            if (OP == 4) { x[c] = fma(x[c], y[c], z[c]); y[c] = fma(y[c], x[(c + 1) % CH], z[c]); }  // operands change every time
            if (OP == 5) x[c] = fma(x[c], 0.999999, 1e-7);            // immediates / constant bank
            if (OP == 6) { bool p = x[c] < y[c]; x[c] = p ? x[c] + z[c] : x[c]; }  // DSETP + predicated DADD
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c] + y[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int CH>
void run(const char* name, int warps_per_sm, int per_step)
{
    double *out, *in; long long* cyc; long long h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&cyc, 8); cudaMalloc(&in, 8 * 64);
    double hin[64]; for (int i = 0; i < 64; ++i) hin[i] = 1.0 + 0.01 * i;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    const int iters = 2048;
    for (int r = 0; r < 2; ++r) { k<OP, CH><<<148, 32 * warps_per_sm>>>(out, in, iters, cyc); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double w = warps_per_sm / 4.0;
    printf("%-28s CH=%d warps/SMSP=%.0f : %.2f SMSP-cycles per FP64 instruction\n", name, CH, w, (double)h / iters / (CH * w * per_step));
    cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main()
{
    const int ws[8] = {4, 8, 12, 16, 20, 24, 28, 32};
    for (int i = 0; i < 8; ++i) run<0, 1>("DFMA dep chain x1", ws[i], 1);
    for (int i = 0; i < 8; ++i) run<0, 2>("DFMA dep chain x2", ws[i], 1);
    for (int i = 0; i < 8; ++i) run<0, 4>("DFMA dep chain x4", ws[i], 1);
    for (int i = 0; i < 8; ++i) run<0, 8>("DFMA dep chain x8", ws[i], 1);
    return 0;
}
