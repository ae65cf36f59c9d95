// common.cuh — shared device helpers for libtrb_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/trb.h"

#ifndef __CUDA_ARCH__
#define TRB_HOST_PASS 1
#endif

namespace trb {

constexpr int kThreads = 256;          // threads per CTA for the reduction kernels
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);

// ---- loads ---------------------------------------------------------------
__device__ __forceinline__ float ldg_f(const float *p) { return __ldg(p); }

// streaming (read-once) load: keep it out of L1 so the gathered volume owns the cache
__device__ __forceinline__ float ld_stream_f(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Row-wise voxel loop: a 256-thread CTA takes two rows (z,y) at a time, 128 threads along x each, so the only
// integer division is one 32-bit div per row (the per-voxel 64-bit idx % W, idx / W of the first version cost
// more than the whole interpolation).  f(idx, x, y, z).
template <int UNROLL = 1, typename F>
__device__ __forceinline__ void for_each_voxel(int D, int H, int W, F f)
{
    const int rows = D * H;
    const int rx = threadIdx.x & 127, ry = threadIdx.x >> 7;
    for (int row = 2 * blockIdx.x + ry; row < rows; row += 2 * gridDim.x) {
        const int z = row / H, y = row - z * H;
        const size_t base = (size_t)row * W;
        // UNROLL 2 keeps two voxels in flight per thread: pays for the light warp kernels (+15 %), costs the
        // statistics / gradient kernels 20 registers and occupancy (-12 %)
#pragma unroll(UNROLL)
        for (int x = rx; x < W; x += 128) f(base + x, x, y, z);
    }
}

// ---- similarity coefficients from the five global moments -------------------
// dL/dw_v = cw*w_v + ct*t_v + c0   (reference utils.py:197-205,
// nn.MSELoss at warpings.py:37,124; weighted sum warpings.py:78-79,144-145)
struct LossCoef {
    double loss, cw, ct, c0;
};
__device__ __forceinline__ LossCoef loss_coefficients(double n, double St, double Sw, double Stt,
                                                      double Sww, double Stw, double w_mse, double w_ncc)
{
    LossCoef o;
    double L = 0.0, gm = 0.0, ga = 0.0, gb = 0.0;
    if (w_mse != 0.0) {
        L += w_mse * ((Stt - 2.0 * Stw + Sww) / n);
        gm = 2.0 * w_mse / n;
    }
    if (w_ncc != 0.0) {
        double A = Stt - St * St / n, B = Sww - Sw * Sw / n, C = Stw - St * Sw / n;
        double S = sqrt(A * B + 1e-10);
        L += w_ncc * 100.0 * (1.0 - C / S);
        ga = -100.0 * w_ncc / S;
        gb = 100.0 * w_ncc * C * A / (S * S * S);
    }
    o.loss = L;
    o.cw = gm + gb;
    o.ct = ga - gm;
    o.c0 = -(ga * St / n + gb * Sw / n);
    return o;
}

}  // namespace trb
