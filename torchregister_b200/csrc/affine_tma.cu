// affine_tma.cu — TMA-staged, warp-specialised 3-D rigid/affine registration epoch (sm_100a).
//
// Same contract as affine_moments_kernel<3,*> in affine.cu (reference call sites listed there),
// re-organised around what bounds it on B200: 8 B/voxel of HBM traffic against ~60 fp32
// pipe operations per voxel.  Design:
//   * persistent CTAs (one per SM): 16 consumer warps + 1 producer warp;
//   * work unit = a z-run of ZC output tiles of 32x16x8 voxels; lane <-> x, warp <-> y, so a
//     thread keeps its x and y base coordinates for the whole unit and steps along z;
//   * the producer maps each output tile through theta, takes the bounding box of its source
//     footprint and, if it fits the fixed TMA box (BX x BY x BZ incl. halo), fetches that box of the
//     moving volume and the target tile with two cp.async.bulk.tensor loads into a 4-stage smem
//     ring (mbarrier full/empty).  TMA's out-of-bounds zero fill IS grid_sample's zeros padding.
//     Tiles whose footprint does not fit (large rotations) are gathered from global memory with
//     explicit bounds checks instead — same arithmetic, slower, always correct;
//   * consumers process two voxels (z, z+1) per thread with packed f32x2 arithmetic
//     (FFMA2/FADD2): half the issue slots for the same fp32 pipe work; floor() is a round-down
//     add of 1.5*2^23 (FADD.RM) instead of FRND/F2I (those run at 16 lanes/clk/SM on the XU pipe,
//     profiles/r01_microbench_issue_rates.txt); the smem index is formed in fp32 and extracted
//     from the mantissa, so the only integer work per voxel is one subtract and one LEA;
//   * moments are kept per thread over a unit (x,y constant -> only the z-weighted sums need a
//     per-voxel FMA), folded with x,y once per unit, reduced per CTA, and finished by the last
//     CTA exactly like the non-TMA kernel (same epilogue code).
#include "common.cuh"
#include "affine_shared.cuh"
#include <cuda.h>

namespace trb {

constexpr int TX = 32, TY = 16, TZ = 8;          // output tile (voxels)
constexpr int kConsumerWarps = TY;               // warp <-> y row of the tile
constexpr int kTmaThreads = (kConsumerWarps + 1) * 32;
constexpr float kMagic = 12582912.f;             // 1.5 * 2^23
constexpr int kMagicBits = 0x4B400000;

// ---- PTX helpers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }

struct TmaParams {
    AffineParams a;
    int n_pairs;
    int tiles_x, tiles_y, tiles_z;     // output tiles per axis
    int zc;                            // tiles per unit (z-run)
    int units_per_col;                 // ceil(tiles_z / zc)
    int n_units;                       // n_pairs * tiles_y * tiles_x * units_per_col
};

struct TileMeta { int ox, oy, oz, fits; };

// per-axis affine map voxel index -> un-normalised source coordinate (grid_sample align_corners=False
// folded into affine_grid): i_r = A[r][0]*xv + A[r][1]*yv + A[r][2]*zv + C[r], xv/yv/zv the base coordinates
struct Coef { float A[3][3], C[3]; };

__device__ __forceinline__ Coef make_coef(const float *th, int D, int H, int W)
{
    Coef k;
    const float h[3] = {0.5f * W, 0.5f * H, 0.5f * D};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) k.A[r][c] = th[r * 4 + c] * h[r];
        k.C[r] = fmaf(th[r * 4 + 3] + 1.f, h[r], -0.5f);
    }
    return k;
}

template <int BX, int BY, int BZ>
struct SmemLayout {
    static constexpr int kBoxFloats = BX * BY * BZ;
    static constexpr int kTgtFloats = TX * TY * TZ;
    static constexpr int kStageBytes = ((kBoxFloats + kTgtFloats) * 4 + 127) / 128 * 128;
};

// one pair of voxels (same x,y; z and z+1).  SECOND=false masks the second voxel (odd tail).
template <int BX, int BY, bool SECOND>
__device__ __forceinline__ void pair_step(const float *__restrict__ box, float2 Mrel, float2 ix, float2 iy, float2 iz,
                                          float2 t, float2 zf, float2 (&s)[5], float2 (&P)[3][3], float2 (&Q)[3][3])
{
    const float2 M = f2(kMagic), nM = f2(-kMagic);
    const float2 flx = __fadd2_rd(ix, M), fly = __fadd2_rd(iy, M), flz = __fadd2_rd(iz, M);
    const float2 fx = __fadd2_rn(flx, nM), fy = __fadd2_rn(fly, nM), fz = __fadd2_rn(flz, nM);
    const float2 tx = sub2(ix, fx), ty = sub2(iy, fy), tz = sub2(iz, fz);
    const float2 fidx = __ffma2_rn(f2((float)(BX * BY)), fz, __ffma2_rn(f2((float)BX), fy, fx));
    const float2 tb = __fadd2_rn(fidx, Mrel);
    const float *qa = box + (__float_as_int(tb.x) - kMagicBits);
    const float *qb = box + (__float_as_int(tb.y) - kMagicBits);
    float2 c000 = make_float2(qa[0], qb[0]), c001 = make_float2(qa[1], qb[1]);
    float2 c010 = make_float2(qa[BX], qb[BX]), c011 = make_float2(qa[BX + 1], qb[BX + 1]);
    float2 c100 = make_float2(qa[BX * BY], qb[BX * BY]), c101 = make_float2(qa[BX * BY + 1], qb[BX * BY + 1]);
    float2 c110 = make_float2(qa[BX * BY + BX], qb[BX * BY + BX]), c111 = make_float2(qa[BX * BY + BX + 1], qb[BX * BY + BX + 1]);
    const float2 d00 = sub2(c001, c000), d01 = sub2(c011, c010), d10 = sub2(c101, c100), d11 = sub2(c111, c110);
    const float2 v00 = __ffma2_rn(tx, d00, c000), v01 = __ffma2_rn(tx, d01, c010);
    const float2 v10 = __ffma2_rn(tx, d10, c100), v11 = __ffma2_rn(tx, d11, c110);
    const float2 e0 = sub2(v01, v00), e1 = sub2(v11, v10);
    const float2 w0 = __ffma2_rn(ty, e0, v00), w1 = __ffma2_rn(ty, e1, v10);
    float2 G[3];
    G[2] = sub2(w1, w0);
    float2 val = __ffma2_rn(tz, G[2], w0);
    G[1] = __ffma2_rn(tz, sub2(e1, e0), e0);
    const float2 dx0 = __ffma2_rn(ty, sub2(d01, d00), d00), dx1 = __ffma2_rn(ty, sub2(d11, d10), d10);
    G[0] = __ffma2_rn(tz, sub2(dx1, dx0), dx0);
    if (!SECOND) {                       // odd tail: the duplicate voxel contributes nothing
        const float2 m = make_float2(1.f, 0.f);
        val = __fmul2_rn(val, m); t = __fmul2_rn(t, m);
#pragma unroll
        for (int r = 0; r < 3; ++r) G[r] = __fmul2_rn(G[r], m);
    }
    s[0] = __fadd2_rn(s[0], t);
    s[1] = __fadd2_rn(s[1], val);
    s[2] = __ffma2_rn(t, t, s[2]);
    s[3] = __ffma2_rn(val, val, s[3]);
    s[4] = __ffma2_rn(t, val, s[4]);
    const float2 tzf = __fmul2_rn(t, zf), wzf = __fmul2_rn(val, zf);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        P[0][r] = __fadd2_rn(P[0][r], G[r]);
        P[1][r] = __ffma2_rn(t, G[r], P[1][r]);
        P[2][r] = __ffma2_rn(val, G[r], P[2][r]);
        Q[0][r] = __ffma2_rn(zf, G[r], Q[0][r]);
        Q[1][r] = __ffma2_rn(tzf, G[r], Q[1][r]);
        Q[2][r] = __ffma2_rn(wzf, G[r], Q[2][r]);
    }
}

// fallback for tiles whose source footprint does not fit the TMA box: one voxel, global gathers
__device__ __forceinline__ void voxel_direct(const float *__restrict__ mov, int D, int H, int W, float ix, float iy, float iz,
                                             float t, float zf, float2 (&s)[5], float2 (&P)[3][3], float2 (&Q)[3][3])
{
    ix = fminf(fmaxf(ix, -4.f), (float)W + 4.f);        // keeps the magic-number floor in range
    iy = fminf(fmaxf(iy, -4.f), (float)H + 4.f);
    iz = fminf(fmaxf(iz, -4.f), (float)D + 4.f);
    const float fx = __fadd_rd(ix, kMagic) - kMagic, fy = __fadd_rd(iy, kMagic) - kMagic, fz = __fadd_rd(iz, kMagic) - kMagic;
    const float tx = ix - fx, ty = iy - fy, tz = iz - fz;
    const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
    const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
    const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
    const bool vz0 = (unsigned)z0 < (unsigned)D, vz1 = (unsigned)(z0 + 1) < (unsigned)D;
    const long long HW = (long long)H * W, o = ((long long)z0 * H + y0) * W + x0;
    const float c000 = (vz0 & vy0 & vx0) ? __ldg(mov + o) : 0.f, c001 = (vz0 & vy0 & vx1) ? __ldg(mov + o + 1) : 0.f;
    const float c010 = (vz0 & vy1 & vx0) ? __ldg(mov + o + W) : 0.f, c011 = (vz0 & vy1 & vx1) ? __ldg(mov + o + W + 1) : 0.f;
    const float c100 = (vz1 & vy0 & vx0) ? __ldg(mov + o + HW) : 0.f, c101 = (vz1 & vy0 & vx1) ? __ldg(mov + o + HW + 1) : 0.f;
    const float c110 = (vz1 & vy1 & vx0) ? __ldg(mov + o + HW + W) : 0.f, c111 = (vz1 & vy1 & vx1) ? __ldg(mov + o + HW + W + 1) : 0.f;
    const float d00 = c001 - c000, d01 = c011 - c010, d10 = c101 - c100, d11 = c111 - c110;
    const float v00 = fmaf(tx, d00, c000), v01 = fmaf(tx, d01, c010), v10 = fmaf(tx, d10, c100), v11 = fmaf(tx, d11, c110);
    const float e0 = v01 - v00, e1 = v11 - v10;
    const float w0 = fmaf(ty, e0, v00), w1 = fmaf(ty, e1, v10);
    float G[3];
    G[2] = w1 - w0;
    const float val = fmaf(tz, G[2], w0);
    G[1] = fmaf(tz, e1 - e0, e0);
    G[0] = fmaf(tz, fmaf(ty, d11 - d10, d10) - fmaf(ty, d01 - d00, d00), fmaf(ty, d01 - d00, d00));
    s[0].x += t; s[1].x += val; s[2].x = fmaf(t, t, s[2].x); s[3].x = fmaf(val, val, s[3].x); s[4].x = fmaf(t, val, s[4].x);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float g = G[r], tg = t * g, wg = val * g;
        P[0][r].x += g; P[1][r].x += tg; P[2][r].x += wg;
        Q[0][r].x = fmaf(zf, g, Q[0][r].x); Q[1][r].x = fmaf(zf, tg, Q[1][r].x); Q[2][r].x = fmaf(zf, wg, Q[2][r].x);
    }
}

template <int BX, int BY, int BZ, int NSTAGE, bool FUSED>
__global__ void __launch_bounds__(kTmaThreads, 1)
affine3d_tma_kernel(const TmaParams p, const __grid_constant__ CUtensorMap map_mov, const __grid_constant__ CUtensorMap map_tgt)
{
    using L = SmemLayout<BX, BY, BZ>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NSTAGE * L::kStageBytes);
    uint64_t *empty_bar = full_bar + NSTAGE;
    TileMeta *meta = reinterpret_cast<TileMeta *>(empty_bar + NSTAGE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = p.a.W, H = p.a.H, D = p.a.D;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, kConsumerWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int units_per_pair = p.tiles_y * p.tiles_x * p.units_per_col;
    const int G = gridDim.x;
    const float inv_d2 = 2.f / (float)D, zoff = 1.f / (float)D - 1.f;      // zv(z) = (2z+1)/D - 1

    if (warp == kConsumerWarps) {
        // ===================== producer: one lane walks the same unit/tile sequence ==========
        if (lane == 0) {
            int it = 0;
            for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
                const int pair = u / units_per_pair;
                int rem = u - pair * units_per_pair;
                const int zc_i = rem / (p.tiles_y * p.tiles_x);
                rem -= zc_i * (p.tiles_y * p.tiles_x);
                const int ty_i = rem / p.tiles_x, tx_i = rem - ty_i * p.tiles_x;
                const int x0 = tx_i * TX, y0 = ty_i * TY;
                const float *st = p.a.state + (size_t)pair * TRB_STATE_FLOATS + TRB_STATE_THETA;
                float th[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) th[i] = __ldcg(st + i);
                const Coef k = make_coef(th, D, H, W);
                const float xa = __ldg(p.a.xb + x0), xe = __ldg(p.a.xb + min(x0 + TX - 1, W - 1));
                const float ya = __ldg(p.a.yb + y0), ye = __ldg(p.a.yb + min(y0 + TY - 1, H - 1));
                const int t_begin = zc_i * p.zc, t_end = min(t_begin + p.zc, p.tiles_z);
                for (int tz_i = t_begin; tz_i < t_end; ++tz_i, ++it) {
                    const int stage = it % NSTAGE;
                    const unsigned phase = (unsigned)(it / NSTAGE) & 1u;
                    mbar_wait(empty_bar + stage, phase ^ 1u);
                    const int z0 = p.a.s_begin + tz_i * TZ;
                    const float za = fmaf(inv_d2, (float)z0, zoff), ze = fmaf(inv_d2, (float)min(z0 + TZ - 1, p.a.s_end - 1), zoff);
                    int o[3];
                    bool fits = true;
                    const int B[3] = {BX, BY, BZ};
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float base = k.A[r][0] * xa + k.A[r][1] * ya + k.A[r][2] * za + k.C[r];
                        const float dx = k.A[r][0] * (xe - xa), dy = k.A[r][1] * (ye - ya), dz = k.A[r][2] * (ze - za);
                        const float lo = base + fminf(dx, 0.f) + fminf(dy, 0.f) + fminf(dz, 0.f) - 0.02f;
                        const float hi = base + fmaxf(dx, 0.f) + fmaxf(dy, 0.f) + fmaxf(dz, 0.f) + 0.02f;
                        // keep the float->int conversions defined for wild thetas
                        const float loc = fminf(fmaxf(lo, -1.0e6f), 1.0e6f), hic = fminf(fmaxf(hi, -1.0e6f), 1.0e6f);
                        // TMA needs the box start 16-byte aligned along x (tools/tma_probe.cu: an unaligned
                        // innermost coordinate raises an illegal-instruction fault); y and z are free
                        o[r] = r == 0 ? 4 * (int)floorf(loc * 0.25f) : (int)floorf(loc);
                        fits = fits && ((int)floorf(hic) + 1 <= o[r] + B[r] - 1);
                    }
                    // the fp32 index trick needs |x + BX*y + BX*BY*z| < 2^22
                    fits = fits && (fabsf((float)o[0]) + BX * fabsf((float)o[1]) + (float)(BX * BY) * fabsf((float)o[2]) < 3.0e6f);
                    TileMeta m;
                    m.ox = o[0]; m.oy = o[1]; m.oz = o[2]; m.fits = fits ? 1 : 0;
                    meta[stage] = m;
                    unsigned char *stg = smem_raw + (size_t)stage * L::kStageBytes;
                    const unsigned tgt_bytes = L::kTgtFloats * 4, box_bytes = L::kBoxFloats * 4;
                    mbar_arrive_expect_tx(full_bar + stage, fits ? (tgt_bytes + box_bytes) : tgt_bytes);
                    if (fits) tma_load_4d(stg, &map_mov, full_bar + stage, o[0], o[1], o[2], pair);
                    tma_load_4d(stg + L::kBoxFloats * 4, &map_tgt, full_bar + stage, x0, y0, z0, pair);
                }
            }
        }
    } else {
        // ===================== consumers ======================================================
        // per-pair accumulators of this thread, folded with its x,y,z base coordinates
        float S[5], T1[3][3], Tx[3][3], Ty[3][3], Tz[3][3];
        auto zero_totals = [&]() {
#pragma unroll
            for (int i = 0; i < 5; ++i) S[i] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                for (int r = 0; r < 3; ++r) T1[kk][r] = Tx[kk][r] = Ty[kk][r] = Tz[kk][r] = 0.f;
        };
        // a CTA's units are pair-major, so it finishes one pair before touching the next: hand the
        // pair's CTA total to the grid-level reduction when the pair changes
        auto flush_pair = [&](int pr) {
            float acc[TRB_MOMENTS];
#pragma unroll
            for (int i = 0; i < 5; ++i) acc[i] = S[i];
#pragma unroll
            for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int b = 5 + kk * 12 + r * 4;
                    acc[b + 0] = Tx[kk][r]; acc[b + 1] = Ty[kk][r]; acc[b + 2] = Tz[kk][r]; acc[b + 3] = T1[kk][r];
                }
            const int count = min(G, units_per_pair);
            const int first = (int)(((long long)pr * units_per_pair) % G);
            reduce_and_finish<3, FUSED, kConsumerWarps>(acc, p.a, pr, blockIdx.x, G, first, count, 1, threadIdx.x);
        };
        zero_totals();
        int cur_pair = -1;
        int it = 0;
        for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
            const int pair = u / units_per_pair;
            if (pair != cur_pair) {
                if (cur_pair >= 0) { flush_pair(cur_pair); zero_totals(); }
                cur_pair = pair;
            }
            int rem = u - pair * units_per_pair;
            const int zc_i = rem / (p.tiles_y * p.tiles_x);
            rem -= zc_i * (p.tiles_y * p.tiles_x);
            const int ty_i = rem / p.tiles_x, tx_i = rem - ty_i * p.tiles_x;
            const int x = tx_i * TX + lane, y = ty_i * TY + warp;
            const bool valid = (x < W) && (y < H);
            const float *st = p.a.state + (size_t)pair * TRB_STATE_FLOATS + TRB_STATE_THETA;
            float th[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) th[i] = __ldcg(st + i);
            const Coef k = make_coef(th, D, H, W);
            const float xv = __ldg(p.a.xb + min(x, W - 1)), yv = __ldg(p.a.yb + min(y, H - 1));
            float pxy[3], sz[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                pxy[r] = fmaf(k.A[r][0], xv, fmaf(k.A[r][1], yv, fmaf(k.A[r][2], zoff, k.C[r])));
                sz[r] = k.A[r][2] * inv_d2;
            }
            const float *__restrict__ mov = p.a.moving + (size_t)pair * p.a.pair_stride;
            float2 s[5], P[3][3], Q[3][3];
#pragma unroll
            for (int i = 0; i < 5; ++i) s[i] = f2(0.f);
#pragma unroll
            for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                for (int r = 0; r < 3; ++r) P[kk][r] = Q[kk][r] = f2(0.f);

            const int t_begin = zc_i * p.zc, t_end = min(t_begin + p.zc, p.tiles_z);
            for (int tz_i = t_begin; tz_i < t_end; ++tz_i, ++it) {
                const int stage = it % NSTAGE;
                const unsigned phase = (unsigned)(it / NSTAGE) & 1u;
                mbar_wait(full_bar + stage, phase);
                const TileMeta m = meta[stage];
                const float *box = reinterpret_cast<const float *>(smem_raw + (size_t)stage * L::kStageBytes);
                const float *tg = box + L::kBoxFloats + warp * TX + lane;
                const int z0 = p.a.s_begin + tz_i * TZ;
                const int nz = min(TZ, p.a.s_end - z0);
                if (valid) {
                    if (m.fits) {
                        const float2 Mrel = f2(kMagic - (float)(m.ox + BX * m.oy + BX * BY * m.oz));
#pragma unroll 2
                        for (int zz = 0; zz < nz; zz += 2) {
                            const bool second = zz + 1 < nz;
                            const float zf0 = (float)(z0 + zz);
                            const float2 zf = make_float2(zf0, second ? zf0 + 1.f : zf0);
                            const float2 ix = __ffma2_rn(f2(sz[0]), zf, f2(pxy[0]));
                            const float2 iy = __ffma2_rn(f2(sz[1]), zf, f2(pxy[1]));
                            const float2 iz = __ffma2_rn(f2(sz[2]), zf, f2(pxy[2]));
                            const float2 t = make_float2(tg[zz * (TX * TY)], tg[(second ? zz + 1 : zz) * (TX * TY)]);
                            if (second) pair_step<BX, BY, true>(box, Mrel, ix, iy, iz, t, zf, s, P, Q);
                            else pair_step<BX, BY, false>(box, Mrel, ix, iy, iz, t, zf, s, P, Q);
                        }
                    } else {
                        for (int zz = 0; zz < nz; ++zz) {
                            const float zf = (float)(z0 + zz);
                            voxel_direct(mov, D, H, W, fmaf(sz[0], zf, pxy[0]), fmaf(sz[1], zf, pxy[1]), fmaf(sz[2], zf, pxy[2]),
                                         tg[zz * (TX * TY)], zf, s, P, Q);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar + stage);
            }
            // fold the unit's sums with this thread's base coordinates (x, y constant over the unit)
            if (valid) {
#pragma unroll
                for (int i = 0; i < 5; ++i) S[i] += s[i].x + s[i].y;
#pragma unroll
                for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float pp = P[kk][r].x + P[kk][r].y, qq = Q[kk][r].x + Q[kk][r].y;
                        T1[kk][r] += pp;
                        Tx[kk][r] = fmaf(xv, pp, Tx[kk][r]);
                        Ty[kk][r] = fmaf(yv, pp, Ty[kk][r]);
                        Tz[kk][r] += fmaf(inv_d2, qq, zoff * pp);
                    }
            }
        }
        if (cur_pair >= 0) flush_pair(cur_pair);
    }

}


// ---- host side ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 4-D map over [pair][D][H][W] fp32 with a (bx,by,bz,1) box; out-of-range elements read as zero
static int make_map(CUtensorMap *map, const float *base, int n_pairs, long long pair_stride, int D, int H, int W,
                    int bx, int by, int bz)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return TRB_ERR_UNSUPPORTED; }
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)n_pairs};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4,
                                   (cuuint64_t)(n_pairs > 1 ? pair_stride : (long long)W * H * D) * 4};
    const cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return TRB_ERR_UNSUPPORTED; }
    return TRB_OK;
}

constexpr int kBX = 40, kBY = 20, kBZ = 12, kStages = 4;

bool tma_path_eligible(int ndim, const AffineParams &a, int n_pairs)
{
    if (ndim != 3) return false;
    if (a.W % 4 != 0 || a.W < TX || a.H < TY || (a.s_end - a.s_begin) < 1) return false;
    if (((uintptr_t)a.moving & 15) || ((uintptr_t)a.target & 15)) return false;
    if (n_pairs > 1 && (a.pair_stride % 4 != 0)) return false;
    // fp32 index trick range: |x + BX*y + BX*BY*z| < 2^22 with some headroom
    if ((double)a.W + (double)kBX * a.H + (double)kBX * kBY * a.D > 2.9e6) return false;
    return encode_fn() != nullptr;
}

// Enqueue n_launch epochs (FUSED) or one moments pass (!FUSED) of the TMA kernel.
int launch_affine3d_tma(AffineParams a, int n_pairs, bool fused, int epoch0, int n_launch, cudaStream_t stream)
{
    using L = SmemLayout<kBX, kBY, kBZ>;
    CUtensorMap map_mov, map_tgt;
    int rc = make_map(&map_mov, a.moving, n_pairs, a.pair_stride, a.D, a.H, a.W, kBX, kBY, kBZ);
    if (rc) return rc;
    rc = make_map(&map_tgt, a.target, n_pairs, a.pair_stride, a.D, a.H, a.W, TX, TY, TZ);
    if (rc) return rc;
    TmaParams p;
    p.a = a;
    p.n_pairs = n_pairs;
    p.tiles_x = (a.W + TX - 1) / TX;
    p.tiles_y = (a.H + TY - 1) / TY;
    p.tiles_z = (a.s_end - a.s_begin + TZ - 1) / TZ;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long cols = (long long)n_pairs * p.tiles_x * p.tiles_y;
    // units of zc tiles along z: aim for >= 8 units per CTA so the static round-robin balances
    long long zc = cols * p.tiles_z / ((long long)sms * 8);
    if (zc < 1) zc = 1;
    if (zc > p.tiles_z) zc = p.tiles_z;
    p.zc = (int)zc;
    p.units_per_col = (p.tiles_z + p.zc - 1) / p.zc;
    const long long n_units = cols * p.units_per_col;
    if (n_units > 0x7fffffffLL) { set_error("too many work units"); return TRB_ERR_UNSUPPORTED; }
    p.n_units = (int)n_units;
    int grid = sms;
    if (grid > p.n_units) grid = p.n_units;
    if (grid > kMaxSlots) grid = kMaxSlots;
    const size_t smem = (size_t)kStages * L::kStageBytes + 2 * kStages * sizeof(uint64_t) + kStages * sizeof(TileMeta);
    auto kf = affine3d_tma_kernel<kBX, kBY, kBZ, kStages, true>;
    auto ku = affine3d_tma_kernel<kBX, kBY, kBZ, kStages, false>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ku, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(tma kernel)");
        attr_set = true;
    }
    for (int e = 0; e < n_launch; ++e) {
        p.a.epoch = epoch0 + e;
        if (fused) kf<<<grid, kTmaThreads, smem, stream>>>(p, map_mov, map_tgt);
        else ku<<<grid, kTmaThreads, smem, stream>>>(p, map_mov, map_tgt);
    }
    return check_cuda(cudaGetLastError(), "affine3d_tma");
}

}  // namespace trb
