// tools/microbench2.cu — register-file pressure of packed fp32 (FFMA2/FADD2) on sm_100a: does an FFMA2 whose three
// operands are three DISTINCT register pairs issue as fast as one with broadcast/reused operands?  (round 2; results in
// profiles/r02_microbench_ffma2_operands.txt).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb2 tools/microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define UNR 8
template <int OP>
__global__ void __launch_bounds__(512, 1) k(float *out, const float *in, int n)
{
    float2 acc[UNR], x[UNR], y[UNR];
    float a[2 * UNR], xs[2 * UNR], ys[2 * UNR];
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
        acc[i] = make_float2(in[threadIdx.x & 31], in[(threadIdx.x + i) & 63]);
        x[i] = make_float2(in[(threadIdx.x + 2 * i) & 63], in[(threadIdx.x + 3 * i) & 63]);
        y[i] = make_float2(in[(threadIdx.x + 5 * i) & 63], in[(threadIdx.x + 7 * i) & 63]);
    }
#pragma unroll
    for (int i = 0; i < 2 * UNR; ++i) { a[i] = in[(threadIdx.x + i) & 63]; xs[i] = in[(threadIdx.x + 2 * i + 1) & 63]; ys[i] = in[(threadIdx.x + 3 * i + 1) & 63]; }
    const float b = in[threadIdx.x & 15], c = in[threadIdx.x & 7];
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < UNR; ++i) {
            if (OP == 0) acc[i] = __ffma2_rn(x[i], y[i], acc[i]);                         // 3 distinct pairs
            if (OP == 1) acc[i] = __ffma2_rn(x[0], y[i], acc[i]);                         // one pair shared by all (reuse)
            if (OP == 2) acc[i] = __ffma2_rn(make_float2(b, b), y[i], acc[i]);            // scalar broadcast + 2 pairs
            if (OP == 3) acc[i] = __ffma2_rn(make_float2(b, b), make_float2(c, c), acc[i]); // 2 broadcasts
            if (OP == 4) { a[2 * i] = fmaf(xs[2 * i], ys[2 * i], a[2 * i]); a[2 * i + 1] = fmaf(xs[2 * i + 1], ys[2 * i + 1], a[2 * i + 1]); }  // 2 scalar FFMA, distinct regs
            if (OP == 5) acc[i] = __fadd2_rn(acc[i], x[i]);                               // FADD2 2 distinct pairs
            if (OP == 6) acc[i] = __ffma2_rn(x[i], x[i], acc[i]);                         // square: 2 distinct pairs
            if (OP == 7) { acc[i] = __ffma2_rn(x[0], y[i], acc[i]); x[i] = __ffma2_rn(x[0], y[i], x[i]); }   // pairs of FFMA2 sharing two operands (moment-update pattern)
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < UNR; ++i) s += acc[i].x + acc[i].y + x[i].x + x[i].y + y[i].x;
#pragma unroll
    for (int i = 0; i < 2 * UNR; ++i) s += a[i] + xs[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent chain: latency of FFMA2 / FADD2 (one warp per SMSP)
template <int OP>
__global__ void __launch_bounds__(128, 1) lat(float *out, const float *in, int n)
{
    float2 v = make_float2(in[threadIdx.x & 31], in[(threadIdx.x + 1) & 31]);
    const float2 m = make_float2(in[2], in[3]), d = make_float2(in[4], in[5]);
    long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) v = __ffma2_rn(v, m, d);
            if (OP == 1) v = __fadd2_rn(v, d);
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = v.x + v.y;
    if (threadIdx.x == 0) ((long long *)out)[64] = t1 - t0;
}
template <int OP>
void run(const char *name, float *out, float *in, double per)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    dim3 g(sms), b(512);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<g, b>>>(out, in, 64);
    cudaEventRecord(e0);
    k<OP><<<g, b>>>(out, in, ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double instr_per_smsp = 4.0 /*warps per SMSP*/ * ITERS * UNR * per;
    printf("%-46s %8.3f ms  => %.2f cycles per instruction per SMSP @%d MHz\n", name, ms, ms * 1e-3 * clk * 1e3 / instr_per_smsp, clk / 1000);
}
int main()
{
    float *out, *in;
    cudaMalloc(&out, 148 * 512 * 4 * 4); cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0, 4096 * 4);
    run<0>("FFMA2 three distinct pairs", out, in, 1);
    run<1>("FFMA2 one pair shared (reuse)", out, in, 1);
    run<2>("FFMA2 scalar broadcast + 2 pairs", out, in, 1);
    run<3>("FFMA2 two broadcasts", out, in, 1);
    run<4>("FFMA scalar, distinct regs (per FFMA)", out, in, 2);
    run<5>("FADD2 two distinct pairs", out, in, 1);
    run<6>("FFMA2 x*x+acc (2 distinct pairs)", out, in, 1);
    run<7>("FFMA2 pairs sharing two operands (per FFMA2)", out, in, 2);
    for (int op = 0; op < 2; ++op) {
        if (op == 0) lat<0><<<1, 128>>>(out, in, 1000); else lat<1><<<1, 128>>>(out, in, 1000);
        cudaDeviceSynchronize();
        long long cyc; cudaMemcpy(&cyc, (long long *)out + 64, 8, cudaMemcpyDeviceToHost);
        printf("dependent %s chain: %.2f cycles per instruction\n", op == 0 ? "FFMA2" : "FADD2", (double)cyc / 16000.0);
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
