// tools/tma_probe.cu — which start coordinates does a tiled TMA load accept? (answers a design question
// for affine_tma.cu; results in profiles/).  nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
constexpr int BX = 40, BY = 20, BZ = 12;
__global__ void k(const __grid_constant__ CUtensorMap map, float *out, int c0, int c1, int c2, int c3)
{
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *bar = (uint64_t *)(sm + BX * BY * BZ * 4);
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(BX * BY * BZ * 4));
        asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6}], [%2];"
                     ::"r"(d), "l"(&map), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < BX * BY * BZ; i += blockDim.x) out[i] = ((float *)sm)[i];
}
int main(int argc, char **argv)
{
    const int W = 44, H = 36, D = 40, P = 2;
    size_t n = (size_t)W * H * D * P;
    float *h = (float *)malloc(n * 4);
    for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 100003);
    float *g, *out; cudaMalloc(&g, n * 4); cudaMalloc(&out, BX * BY * BZ * 4);
    cudaMemcpy(g, h, n * 4, cudaMemcpyHostToDevice);
    void *fnp; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fnp;
    CUtensorMap map;
    cuuint64_t dims[4] = {W, H, D, P}, str[3] = {W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4};
    cuuint32_t box[4] = {BX, BY, BZ, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    int smem = BX * BY * BZ * 4 + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float *ho = (float *)malloc(BX * BY * BZ * 4);
    int tests[][4] = {{0, 0, 0, 0}, {4, 0, 0, 0}, {8, 3, 5, 1}, {-4, 0, 0, 0}, {-8, -2, -1, 1}, {-4, -7, 3, 0}, {12, 20, 33, 1}, {28, 30, 35, 0}, {40, 1, 8, 0}, {-40, 0, 0, 0}, {44, 0, 0, 0}};
    for (auto &t : tests) {
        k<<<1, 128, smem>>>(map, out, t[0], t[1], t[2], t[3]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("coords (%d,%d,%d,%d): %s\n", t[0], t[1], t[2], t[3], cudaGetErrorString(e)); return 1; }
        cudaMemcpy(ho, out, BX * BY * BZ * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int z = 0; z < BZ; ++z) for (int y = 0; y < BY; ++y) for (int x = 0; x < BX; ++x) {
            int gx = t[0] + x, gy = t[1] + y, gz = t[2] + z;
            float exp = 0.f;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H && gz >= 0 && gz < D) exp = h[(((size_t)t[3] * D + gz) * H + gy) * W + gx];
            if (ho[(z * BY + y) * BX + x] != exp) ++bad;
        }
        printf("coords (%d,%d,%d,%d): ok, mismatches=%d\n", t[0], t[1], t[2], t[3], bad);
    }
    return 0;
}
