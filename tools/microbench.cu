// tools/microbench.cu — instruction-issue microbenchmarks used to budget the fused kernel
// (results recorded in profiles/). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNR 8

template <int OP>
__global__ void __launch_bounds__(256) k(float *out, const float *in, int n)
{
    float a[UNR], b = in[threadIdx.x & 31], c = in[(threadIdx.x + 1) & 31];
    float2 a2[UNR];
#pragma unroll
    for (int i = 0; i < UNR; ++i) { a[i] = in[(threadIdx.x + i) & 63]; a2[i] = make_float2(a[i], a[i] + 1.f); }
    int ia[UNR];
#pragma unroll
    for (int i = 0; i < UNR; ++i) ia[i] = threadIdx.x + i;
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = in[i & 63];
    __syncthreads();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < UNR; ++i) {
            if (OP == 0) a[i] = fmaf(a[i], b, c);                                   // FFMA
            if (OP == 1) a2[i] = __ffma2_rn(a2[i], make_float2(b, b), make_float2(c, c));   // FFMA2
            if (OP == 2) a[i] = a[i] + b;                                            // FADD
            if (OP == 3) a2[i] = __fadd2_rn(a2[i], make_float2(b, c));               // FADD2
            if (OP == 4) a[i] = floorf(a[i]) + b;                                    // FRND + FADD
            if (OP == 5) { ia[i] = (int)a[i]; a[i] = a[i] + (float)ia[i]; }          // F2I + I2F + FADD
            if (OP == 6) a[i] = sm[(ia[i] + (int)a[i]) & 2047] + a[i] * 0.5f;       // LDS (+F2I, FFMA)
            if (OP == 7) ia[i] = ia[i] * 3 + it;                                     // IMAD
            if (OP == 8) a[i] = (a[i] > b) ? a[i] - c : a[i] + c;                    // FSETP+FSEL-ish
            if (OP == 9) { a[i] = fmaf(a[i], b, c); ia[i] = (ia[i] + it) ^ i; }      // FFMA + ALU mix
            if (OP == 10) a[i] = __ldg(in + ((ia[i] + it) & 63)) + a[i];            // LDG L1-hit
            if (OP == 11) { ia[i] = (ia[i] + 1) & 2047; a[i] += sm[ia[i]]; }         // LDS consecutive
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < UNR; ++i) s += a[i] + a2[i].x + a2[i].y + (float)ia[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, float *out, float *in, double ops_per_iter)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    dim3 g(sms * 8), b(256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<g, b>>>(out, in, 64);
    cudaEventRecord(e0);
    k<OP><<<g, b>>>(out, in, ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)g.x * 8 /*warps*/ * ITERS * UNR * ops_per_iter;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %7.2f Gwarp-instr/s  => %.2f warp-instr/clk/SM @%d MHz(max)\n", name, ms,
           warp_instr / ms / 1e6, warp_instr / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
}

int main()
{
    float *out, *in;
    cudaMalloc(&out, 148 * 8 * 256 * 4 * 4); cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0, 4096 * 4);
    run<0>("FFMA", out, in, 1);
    run<1>("FFMA2 (count as 1)", out, in, 1);
    run<2>("FADD", out, in, 1);
    run<3>("FADD2 (count as 1)", out, in, 1);
    run<4>("FRND+FADD (2)", out, in, 2);
    run<5>("F2I+I2F+FADD (3)", out, in, 3);
    run<6>("LDS+F2I+IADD+LOP+FFMA (5)", out, in, 5);
    run<7>("IMAD", out, in, 1);
    run<8>("FSETP+FADD+FADD+SEL (4?)", out, in, 4);
    run<9>("FFMA+IADD+LOP (3)", out, in, 3);
    run<10>("LDG(L1)+IADD+LOP+FADD (4)", out, in, 4);
    run<11>("LDS+IADD+LOP+FADD (4)", out, in, 4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
