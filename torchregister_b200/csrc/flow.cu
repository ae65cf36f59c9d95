// flow.cu — flow-field warp, its VJP, and the fused warp+similarity+gradient node (sm_100a).
//
// Replaces, from the reference (paths relative to /root/reference/src/TorchRegister/):
//   SpatialTransformer.forward                         utils.py:350-365
//   the similarity forward + backward down to the flow warpings.py:213-215 (MSELoss, NCCLoss utils.py:197-205)
//   flow_register.deform / Register.__call__           warpings.py:238-242, torchregister.py:124-125
//
// Layout: src/target fp32 [D][H][W]; flow fp32 [ndim][D][H][W] (planar, channel i displaces
// spatial axis i).  One thread per output voxel, x fastest: flow/target/dflow/warped accesses are
// fully coalesced streams, the 8 (4) gathered corners hit neighbouring lines of `src`.
// Algorithmic traffic of the fused node: 32 B/voxel (moving 4 + target 4 + flow 12 read,
// dflow 12 written) + 20 B/voxel for the statistics pre-pass when the NCC term is on.
#include "common.cuh"

namespace trb {

// voxel (x,y,z) -> the sample position displaced by the flow
struct FlowAxes {
    AxisMap x, y, z;
};
__device__ __forceinline__ FlowAxes flow_axes(int D, int H, int W)
{
    FlowAxes a;
    a.x = axis_map(W); a.y = axis_map(H); a.z = axis_map(D > 1 ? D : 2);
    return a;
}
template <int NDIM>
__device__ __forceinline__ void flow_position(const float *__restrict__ flow, size_t vol, size_t idx, int x, int y, int z,
                                              const FlowAxes &ax, float &px, float &py, float &pz)
{
    if (NDIM == 3) {
        pz = flow_pos(ax.z, z, ld_stream_f(flow + idx));
        py = flow_pos(ax.y, y, ld_stream_f(flow + vol + idx));
        px = flow_pos(ax.x, x, ld_stream_f(flow + 2 * vol + idx));
    } else {
        pz = 0.f;
        py = flow_pos(ax.y, y, ld_stream_f(flow + idx));
        px = flow_pos(ax.x, x, ld_stream_f(flow + vol + idx));
    }
}

template <int NDIM>
__global__ void __launch_bounds__(256, 4) warp_flow_kernel(const float *__restrict__ src, const float *__restrict__ flow,
                                                         float *__restrict__ out, int n_channels, int D, int H, int W)
{
    const size_t vol = (size_t)(NDIM == 3 ? D : 1) * H * W;
    const FlowAxes ax = flow_axes(D, H, W);
    for_each_voxel<2>(NDIM == 3 ? D : 1, H, W, [&](size_t idx, int x, int y, int z) {
        float px, py, pz;
        flow_position<NDIM>(flow, vol, idx, x, y, z, ax, px, py, pz);
        for (int c = 0; c < n_channels; ++c)
            out[(size_t)c * vol + idx] = sample_zero_pad<NDIM, false>(src + (size_t)c * vol, D, H, W, px, py, pz).val;
    });
}

// dflow channel a (spatial axis a) receives the derivative along sampling coordinate NDIM-1-a
template <int NDIM>
__device__ __forceinline__ void store_dflow(float *__restrict__ dflow, size_t vol, size_t idx, float r, const float *g)
{
#pragma unroll
    for (int a = 0; a < NDIM; ++a) dflow[(size_t)a * vol + idx] = r * g[NDIM - 1 - a];
}

template <int NDIM>
__global__ void __launch_bounds__(256) warp_flow_vjp_kernel(const float *__restrict__ src, const float *__restrict__ flow,
                                                             const float *__restrict__ gout, float *__restrict__ dflow,
                                                             int D, int H, int W)
{
    const size_t vol = (size_t)(NDIM == 3 ? D : 1) * H * W;
    const FlowAxes ax = flow_axes(D, H, W);
    for_each_voxel(NDIM == 3 ? D : 1, H, W, [&](size_t idx, int x, int y, int z) {
        float px, py, pz;
        flow_position<NDIM>(flow, vol, idx, x, y, z, ax, px, py, pz);
        const Sample<NDIM> s = sample_zero_pad<NDIM, true>(src, D, H, W, px, py, pz);
        store_dflow<NDIM>(dflow, vol, idx, ld_stream_f(gout + idx), s.g);
    });
}

// workspace layout (doubles): [0..3] cw, ct, c0, loss ; [4] ticket (as unsigned) ; [8..] partials[blocks][5]
constexpr int kFlowMaxBlocks = 2048;

template <int NDIM>
__global__ void __launch_bounds__(256) flow_stats_kernel(const float *__restrict__ moving, const float *__restrict__ target,
                                                          const float *__restrict__ flow, float *__restrict__ warped,
                                                          int D, int H, int W, double w_mse, double w_ncc,
                                                          double *ws, float *loss_out)
{
    const size_t vol = (size_t)(NDIM == 3 ? D : 1) * H * W;
    const FlowAxes ax = flow_axes(D, H, W);
    // a thread sums ~50-60 voxels (grid = 8 CTAs per SM): fp32 partials are exact enough (values in [0,1],
    // relative error < 4e-6 worst case) and keep the register count low enough for full occupancy; everything
    // above the thread level is fp64
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for_each_voxel(NDIM == 3 ? D : 1, H, W, [&](size_t idx, int x, int y, int z) {
        float px, py, pz;
        flow_position<NDIM>(flow, vol, idx, x, y, z, ax, px, py, pz);
        const float w = sample_zero_pad<NDIM, false, false>(moving, D, H, W, px, py, pz).val;
        const float t = ld_stream_f(target + idx);
        if (warped) warped[idx] = w;
        s[0] += t; s[1] += w;
        s[2] = fmaf(t, t, s[2]); s[3] = fmaf(w, w, s[3]); s[4] = fmaf(t, w, s[4]);
    });
    double acc[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) acc[i] = (double)s[i];
    __shared__ double red[8][5];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    double *partials = ws + 8;
    unsigned *ticket = (unsigned *)(ws + 4);
    if (threadIdx.x < 5) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        __stcg(partials + (size_t)blockIdx.x * 5 + threadIdx.x, v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // fixed-order reduction over blocks: warp i < 5 sums moment i
    if (warp < 5) {
        double v = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) v += __ldcg(partials + (size_t)b * 5 + warp);
        v = warp_sum(v);
        if (lane == 0) red[0][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const LossCoef lc = loss_coefficients((double)vol, red[0][0], red[0][1], red[0][2], red[0][3], red[0][4], w_mse, w_ncc);
        ws[0] = lc.cw; ws[1] = lc.ct; ws[2] = lc.c0; ws[3] = lc.loss;
        if (loss_out) *loss_out = (float)lc.loss;
        *ticket = 0u;
    }
}

template <int NDIM>
__global__ void __launch_bounds__(256) flow_grad_kernel(const float *__restrict__ moving, const float *__restrict__ target,
                                                         const float *__restrict__ flow, float *__restrict__ dflow,
                                                         int D, int H, int W, const double *__restrict__ ws)
{
    const size_t vol = (size_t)(NDIM == 3 ? D : 1) * H * W;
    const FlowAxes ax = flow_axes(D, H, W);
    const float cw = (float)ws[0], ct = (float)ws[1], c0 = (float)ws[2];
    for_each_voxel(NDIM == 3 ? D : 1, H, W, [&](size_t idx, int x, int y, int z) {
        float px, py, pz;
        flow_position<NDIM>(flow, vol, idx, x, y, z, ax, px, py, pz);
        const Sample<NDIM> s = sample_zero_pad<NDIM, true, false>(moving, D, H, W, px, py, pz);
        const float t = ld_stream_f(target + idx);
        const float r = fmaf(cw, s.val, fmaf(ct, t, c0));
        store_dflow<NDIM>(dflow, vol, idx, r, s.g);
    });
}

// MSE only (w_ncc == 0): dL/dw_v = 2 w_mse (w_v - t_v) / n needs no global moments, so the statistics pass is not needed —
// ONE pass samples, writes d loss / d flow (and optionally the warped volume) and reduces sum (w - t)^2 for the loss.
template <int NDIM>
__global__ void __launch_bounds__(256) flow_mse_fused_kernel(const float *__restrict__ moving, const float *__restrict__ target,
                                                              const float *__restrict__ flow, float *__restrict__ dflow,
                                                              float *__restrict__ warped, int D, int H, int W, double w_mse,
                                                              double *__restrict__ ws, float *__restrict__ loss_out)
{
    const size_t vol = (size_t)(NDIM == 3 ? D : 1) * H * W;
    const FlowAxes ax = flow_axes(D, H, W);
    const float gm = (float)(2.0 * w_mse / (double)vol);
    float sd = 0.f;
    for_each_voxel(NDIM == 3 ? D : 1, H, W, [&](size_t idx, int x, int y, int z) {
        float px, py, pz;
        flow_position<NDIM>(flow, vol, idx, x, y, z, ax, px, py, pz);
        const Sample<NDIM> s = sample_zero_pad<NDIM, true, false>(moving, D, H, W, px, py, pz);
        const float t = ld_stream_f(target + idx);
        const float d = s.val - t;
        if (warped) warped[idx] = s.val;
        sd = fmaf(d, d, sd);
        store_dflow<NDIM>(dflow, vol, idx, gm * d, s.g);
    });
    __shared__ double red[8];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double v = warp_sum((double)sd);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double *partials = ws + 8;
    unsigned *ticket = (unsigned *)(ws + 4);
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        __stcg(partials + blockIdx.x, t);
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (warp == 0) {                               // fixed-order reduction over blocks
        double t = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(partials + b);
        t = warp_sum(t);
        if (lane == 0) {
            const double loss = w_mse * t / (double)vol;
            ws[3] = loss;
            if (loss_out) *loss_out = (float)loss;
            *ticket = 0u;
        }
    }
}

static unsigned flow_grid(size_t vol)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t nb = (vol + 255) / 256;      // enough CTAs for every SM to hold 8; each strides over row pairs
    const size_t cap = (size_t)sms * 8;
    if (nb > cap) nb = cap;
    if (nb > kFlowMaxBlocks) nb = kFlowMaxBlocks;
    return (unsigned)(nb < 1 ? 1 : nb);
}

static int validate_flow(int ndim, int D, int H, int W)
{
    if (ndim != 2 && ndim != 3) { set_error("ndim must be 2 or 3 (got %d)", ndim); return TRB_ERR_ARG; }
    if (H < 2 || W < 2 || (ndim == 3 && D < 2)) { set_error("flow warp needs every axis >= 2 (got %dx%dx%d)", D, H, W); return TRB_ERR_ARG; }
    if ((unsigned long long)(ndim == 3 ? D : 1) * H * W >= (1ull << 31)) { set_error("flow kernels index volumes with 32 bits: D*H*W must stay below 2^31"); return TRB_ERR_ARG; }
    return TRB_OK;
}

}  // namespace trb

using namespace trb;

extern "C" size_t trb_flow_workspace_bytes(void) { return (size_t)(8 + kFlowMaxBlocks * 5) * sizeof(double); }

extern "C" int trb_warp_flow(int ndim, const float *src_dev, const float *flow_dev, float *out_dev, int n_channels,
                             int D, int H, int W, void *stream)
{
    int rc = validate_flow(ndim, D, H, W);
    if (rc) return rc;
    if (!src_dev || !flow_dev || !out_dev || n_channels < 1) { set_error("null pointer / n_channels"); return TRB_ERR_ARG; }
    const size_t vol = (size_t)(ndim == 3 ? D : 1) * H * W;
    cudaStream_t s = (cudaStream_t)stream;
    if (ndim == 3) warp_flow_kernel<3><<<flow_grid(vol), 256, 0, s>>>(src_dev, flow_dev, out_dev, n_channels, D, H, W);
    else warp_flow_kernel<2><<<flow_grid(vol), 256, 0, s>>>(src_dev, flow_dev, out_dev, n_channels, 1, H, W);
    return check_cuda(cudaGetLastError(), "warp_flow");
}

extern "C" int trb_warp_flow_vjp(int ndim, const float *src_dev, const float *flow_dev, const float *gout_dev,
                                 float *dflow_dev, int D, int H, int W, void *stream)
{
    int rc = validate_flow(ndim, D, H, W);
    if (rc) return rc;
    if (!src_dev || !flow_dev || !gout_dev || !dflow_dev) { set_error("null pointer"); return TRB_ERR_ARG; }
    const size_t vol = (size_t)(ndim == 3 ? D : 1) * H * W;
    cudaStream_t s = (cudaStream_t)stream;
    if (ndim == 3) warp_flow_vjp_kernel<3><<<flow_grid(vol), 256, 0, s>>>(src_dev, flow_dev, gout_dev, dflow_dev, D, H, W);
    else warp_flow_vjp_kernel<2><<<flow_grid(vol), 256, 0, s>>>(src_dev, flow_dev, gout_dev, dflow_dev, 1, H, W);
    return check_cuda(cudaGetLastError(), "warp_flow_vjp");
}

extern "C" int trb_flow_loss_grad(int ndim, const float *moving_dev, const float *target_dev, const float *flow_dev,
                                  int D, int H, int W, float w_mse, float w_ncc, float *loss_dev, float *dflow_dev,
                                  float *warped_dev_or_null, void *workspace_dev, size_t workspace_bytes, void *stream)
{
    int rc = validate_flow(ndim, D, H, W);
    if (rc) return rc;
    if (!moving_dev || !target_dev || !flow_dev || !loss_dev || !dflow_dev) { set_error("null pointer"); return TRB_ERR_ARG; }
    if (!workspace_dev || workspace_bytes < trb_flow_workspace_bytes()) {
        set_error("workspace too small: need %zu bytes", trb_flow_workspace_bytes());
        return TRB_ERR_WORKSPACE;
    }
    const size_t vol = (size_t)(ndim == 3 ? D : 1) * H * W;
    cudaStream_t s = (cudaStream_t)stream;
    double *ws = (double *)workspace_dev;
    const unsigned g = flow_grid(vol);
    if (w_ncc == 0.f) {                 // no global moments needed: one fused pass (32 B/voxel instead of 52)
        if (ndim == 3) flow_mse_fused_kernel<3><<<g, 256, 0, s>>>(moving_dev, target_dev, flow_dev, dflow_dev, warped_dev_or_null, D, H, W, w_mse, ws, loss_dev);
        else flow_mse_fused_kernel<2><<<g, 256, 0, s>>>(moving_dev, target_dev, flow_dev, dflow_dev, warped_dev_or_null, 1, H, W, w_mse, ws, loss_dev);
        return check_cuda(cudaGetLastError(), "flow_loss_grad");
    }
    if (ndim == 3) {
        flow_stats_kernel<3><<<g, 256, 0, s>>>(moving_dev, target_dev, flow_dev, warped_dev_or_null, D, H, W, w_mse, w_ncc, ws, loss_dev);
        flow_grad_kernel<3><<<g, 256, 0, s>>>(moving_dev, target_dev, flow_dev, dflow_dev, D, H, W, ws);
    } else {
        flow_stats_kernel<2><<<g, 256, 0, s>>>(moving_dev, target_dev, flow_dev, warped_dev_or_null, 1, H, W, w_mse, w_ncc, ws, loss_dev);
        flow_grad_kernel<2><<<g, 256, 0, s>>>(moving_dev, target_dev, flow_dev, dflow_dev, 1, H, W, ws);
    }
    return check_cuda(cudaGetLastError(), "flow_loss_grad");
}
