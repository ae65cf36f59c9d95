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

// ---- flow sample positions and zero-padded (tri|bi)linear sampling (flow.cu, flow_direct.cu) ----------------
// The reference normalises voxel position + flow to [-1,1] (utils.py:354-356) and grid_sample(align_corners=True)
// un-normalises again; the fp32 round trip is kept op for op.  Its IEEE division by the loop-invariant S-1 is done
// with the hoisted correctly rounded reciprocal: q0 = RN(a*r), q = RN(q0 + (a - d*q0)*r) is the correctly rounded
// quotient (Markstein), 3 instructions instead of the ~14 of the generic division sequence.
struct AxisMap {
    float d, r;      // S-1 and RN(1/(S-1))
};
__device__ __forceinline__ AxisMap axis_map(int S)
{
    AxisMap a;
    a.d = (float)(S - 1);
    a.r = __frcp_rn(a.d);
    return a;
}
__device__ __forceinline__ float flow_pos(const AxisMap &m, int i, float f)
{
    const float loc = (float)i + f;
    const float q0 = __fmul_rn(loc, m.r);
    const float q = __fmaf_rn(__fmaf_rn(-m.d, q0, loc), m.r, q0);
    const float nrm = 2.f * (q - 0.5f);
    return __fmul_rn(__fmul_rn(__fadd_rn(nrm, 1.f), 0.5f), m.d);
}

// floor and fraction without the quarter-rate conversion pipe: adding 1.5*2^23 rounding down leaves floor(p) in the
// low mantissa bits (exact for |p| < 2^22; anything further out — inf and NaN included — lands on an index no volume
// reaches, i.e. on the zero padding).
__device__ __forceinline__ void floor_frac(float p, int &i0, float &t, float &fl)
{
    const float m = __fadd_rd(p, 12582912.f);
    i0 = __float_as_int(m) - 0x4B400000;
    fl = m - 12582912.f;
    t = p - fl;
}
__device__ __forceinline__ void floor_frac(float p, int &i0, float &t)
{
    float fl;
    floor_frac(p, i0, t, fl);
}

// cells that straddle the volume boundary (rare: kept out of line so the common path stays small)
static __device__ __noinline__ void gather_border3(const float *__restrict__ m, int D, int H, int W, int x0, int y0, int z0,
                                            float (&c)[8])
{
    const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
    const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
    const bool vz0 = (unsigned)z0 < (unsigned)D, vz1 = (unsigned)(z0 + 1) < (unsigned)D;
    const long long HW = (long long)H * W;
    const float *q = m + (((long long)z0 * H + y0) * W + x0);
    c[0] = (vz0 & vy0 & vx0) ? __ldg(q) : 0.f;
    c[1] = (vz0 & vy0 & vx1) ? __ldg(q + 1) : 0.f;
    c[2] = (vz0 & vy1 & vx0) ? __ldg(q + W) : 0.f;
    c[3] = (vz0 & vy1 & vx1) ? __ldg(q + W + 1) : 0.f;
    c[4] = (vz1 & vy0 & vx0) ? __ldg(q + HW) : 0.f;
    c[5] = (vz1 & vy0 & vx1) ? __ldg(q + HW + 1) : 0.f;
    c[6] = (vz1 & vy1 & vx0) ? __ldg(q + HW + W) : 0.f;
    c[7] = (vz1 & vy1 & vx1) ? __ldg(q + HW + W + 1) : 0.f;
}

template <int NDIM>
struct Sample {
    float val;
    float g[NDIM];      // d val / d (x, y[, z]) in voxel units
};

// The 8 corners of the cell around (px,py,pz) + the fractions: the loads can be issued early and blended later.
struct Cell3 {
    float c[8];
    float tx, ty, tz;
    float fx, fy, fz;       // floor of the position (the cell's origin), for sample_near
};
// BRANCH_FREE: predicated loads, no control flow — the compiler can overlap the gathers of neighbouring voxels and
// phases (what latency-bound kernels want: warp_flow -22 %, fused direct-flow epoch -10 %); otherwise interior cells
// take an unpredicated fast path and boundary cells an out-of-line one (fewer instructions: flow node -11 %).
template <bool BRANCH_FREE = true>
__device__ __forceinline__ Cell3 gather_cell3(const float *__restrict__ m, int D, int H, int W, float px, float py, float pz)
{
    Cell3 k;
    int x0, y0, z0;
    floor_frac(px, x0, k.tx, k.fx);
    floor_frac(py, y0, k.ty, k.fy);
    floor_frac(pz, z0, k.tz, k.fz);
    if constexpr (BRANCH_FREE) {
        const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
        const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
        const bool vz0 = (unsigned)z0 < (unsigned)D, vz1 = (unsigned)(z0 + 1) < (unsigned)D;
        const int HW = H * W;
        const float *q = m + ((z0 * H + y0) * W + x0);
        k.c[0] = (vz0 & vy0 & vx0) ? __ldg(q) : 0.f;
        k.c[1] = (vz0 & vy0 & vx1) ? __ldg(q + 1) : 0.f;
        k.c[2] = (vz0 & vy1 & vx0) ? __ldg(q + W) : 0.f;
        k.c[3] = (vz0 & vy1 & vx1) ? __ldg(q + W + 1) : 0.f;
        k.c[4] = (vz1 & vy0 & vx0) ? __ldg(q + HW) : 0.f;
        k.c[5] = (vz1 & vy0 & vx1) ? __ldg(q + HW + 1) : 0.f;
        k.c[6] = (vz1 & vy1 & vx0) ? __ldg(q + HW + W) : 0.f;
        k.c[7] = (vz1 & vy1 & vx1) ? __ldg(q + HW + W + 1) : 0.f;
        return k;
    }
    if ((unsigned)x0 < (unsigned)(W - 1) && (unsigned)y0 < (unsigned)(H - 1) && (unsigned)z0 < (unsigned)(D - 1)) {
        const float *q = m + ((z0 * H + y0) * W + x0);          // the whole cell is inside: no predicates
        k.c[0] = __ldg(q); k.c[1] = __ldg(q + 1); k.c[2] = __ldg(q + W); k.c[3] = __ldg(q + W + 1);
        q += H * W;
        k.c[4] = __ldg(q); k.c[5] = __ldg(q + 1); k.c[6] = __ldg(q + W); k.c[7] = __ldg(q + W + 1);
    } else {
        float b[8];                                             // only this rare path touches local memory
        gather_border3(m, D, H, W, x0, y0, z0, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) k.c[i] = b[i];
    }
    return k;
}
template <bool WANT_GRAD>
__device__ __forceinline__ Sample<3> blend_cell3(const Cell3 &k)
{
    Sample<3> s;
    const float d00 = k.c[1] - k.c[0], d01 = k.c[3] - k.c[2], d10 = k.c[5] - k.c[4], d11 = k.c[7] - k.c[6];
    const float v00 = fmaf(k.tx, d00, k.c[0]), v01 = fmaf(k.tx, d01, k.c[2]);
    const float v10 = fmaf(k.tx, d10, k.c[4]), v11 = fmaf(k.tx, d11, k.c[6]);
    const float e0 = v01 - v00, e1 = v11 - v10;
    const float w0 = fmaf(k.ty, e0, v00), w1 = fmaf(k.ty, e1, v10);
    const float gz = w1 - w0;
    s.val = fmaf(k.tz, gz, w0);
    if (WANT_GRAD) {
        s.g[2] = gz;
        s.g[1] = fmaf(k.tz, e1 - e0, e0);
        const float dx0 = fmaf(k.ty, d01 - d00, d00), dx1 = fmaf(k.ty, d11 - d10, d10);
        s.g[0] = fmaf(k.tz, dx1 - dx0, dx0);
    }
    return s;
}

// Value at a second position that usually lies in the cell already gathered (a gradient step moves a sample by a
// fraction of a voxel): same corners, new fractions — bit-identical to a fresh gather, without its 8 loads.
__device__ __forceinline__ float sample_near(const Cell3 &k, const float *__restrict__ m, int D, int H, int W,
                                             float qx, float qy, float qz);

// grid_sample(mode='bilinear', padding_mode='zeros') at voxel coordinates (px,py,pz).  Needs D*H*W < 2^31 (hosts check).
template <int NDIM, bool WANT_GRAD, bool BRANCH_FREE = true>
__device__ __forceinline__ Sample<NDIM> sample_zero_pad(const float *__restrict__ m, int D, int H, int W,
                                                        float px, float py, float pz)
{
    if constexpr (NDIM == 3) {
        return blend_cell3<WANT_GRAD>(gather_cell3<BRANCH_FREE>(m, D, H, W, px, py, pz));
    } else {
        Sample<NDIM> s;
        int x0, y0;
        float tx, ty;
        floor_frac(px, x0, tx);
        floor_frac(py, y0, ty);
        const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
        const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
        const float *q = m + ((long long)y0 * W + x0);
        const float c00 = (vy0 & vx0) ? __ldg(q) : 0.f;
        const float c01 = (vy0 & vx1) ? __ldg(q + 1) : 0.f;
        const float c10 = (vy1 & vx0) ? __ldg(q + W) : 0.f;
        const float c11 = (vy1 & vx1) ? __ldg(q + W + 1) : 0.f;
        const float d0 = c01 - c00, d1 = c11 - c10;
        const float v0 = fmaf(tx, d0, c00), v1 = fmaf(tx, d1, c10);
        const float gy = v1 - v0;
        s.val = fmaf(ty, gy, v0);
        if (WANT_GRAD) {
            s.g[1] = gy;
            s.g[0] = fmaf(ty, d1 - d0, d0);
        }
        return s;
    }
}

__device__ __forceinline__ float sample_near(const Cell3 &k, const float *__restrict__ m, int D, int H, int W,
                                             float qx, float qy, float qz)
{
    const float tx = qx - k.fx, ty = qy - k.fy, tz = qz - k.fz;
    if (tx >= 0.f && tx < 1.f && ty >= 0.f && ty < 1.f && tz >= 0.f && tz < 1.f) {
        Cell3 c2 = k;
        c2.tx = tx; c2.ty = ty; c2.tz = tz;
        return blend_cell3<false>(c2).val;
    }
    return sample_zero_pad<3, false>(m, D, H, W, qx, qy, qz).val;
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
    // one reciprocal of n and one of S instead of eight divisions: this runs on the critical path between two epochs
    // of a single pair (a dependent fp64 division is ~150 cycles)
    const double inv_n = 1.0 / n;
    if (w_mse != 0.0) {
        L += w_mse * ((Stt - 2.0 * Stw + Sww) * inv_n);
        gm = 2.0 * w_mse * inv_n;
    }
    if (w_ncc != 0.0) {
        double A = Stt - St * St * inv_n, B = Sww - Sw * Sw * inv_n, C = Stw - St * Sw * inv_n;
        double S = sqrt(A * B + 1e-10);
        const double inv_S = 1.0 / S;
        L += w_ncc * 100.0 * (1.0 - C * inv_S);
        ga = -100.0 * w_ncc * inv_S;
        gb = 100.0 * w_ncc * C * A * (inv_S * inv_S * inv_S);
    }
    o.loss = L;
    o.cw = gm + gb;
    o.ct = ga - gm;
    o.c0 = -(ga * St + gb * Sw) * inv_n;
    return o;
}

}  // namespace trb
