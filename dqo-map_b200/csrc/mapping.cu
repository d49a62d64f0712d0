// Mapping-step kernels around the rasterizer: masked L1 colour/depth loss with image gradients, fused
// multi-tensor Adam, per-Gaussian error scatter.
//
//   dqo_masked_l1_loss   <- Mapping.loss_update, SLAM/multiprocess/mapper.py:830-875 (l1_loss utils/loss_utils.py:27-31)
//   dqo_adam_step        <- torch.optim.Adam(l, lr=0.0, eps=1e-15).step() over GaussianPointCloud.parametrize groups
//                           (SLAM/gaussian_pointcloud.py:331-378, mapper.py:548,906) + confidence bump mapper.py:909-910
//   dqo_accumulate_error <- accumulate_gaussian_error_impl, submodules/cuda_utils/map_process.cu:33-245
#include "common.cuh"

namespace dqo {

// ------------------------------------------------------------------------------------------------
// masked L1 loss
// ------------------------------------------------------------------------------------------------
#define LOSS_BLOCKS 592 // 4 per SM
#define LOSS_THREADS 256

struct LossArgs {
    int W, H;
    const float *image, *depth, *gt_color, *gt_depth;
    const int *hit;
    const uint8_t *mask;
    float color_w, depth_w, depth_thr;
    float *dimg, *ddepth, *loss_out;
    int *counts_out;
    double *partial; // [LOSS_BLOCKS][4]: colour sum, colour pixel count, depth sum, depth count
    // optional (fused step): the tile mask of the render.  The backward blend reads the gradient images only in tiles that
    // were rendered, so the zero gradients of masked-out tiles need not be written.
    const int *tile_mask;
    int tiles_x;
};

__device__ __forceinline__ bool depth_valid(const LossArgs &a, size_t p, bool m, float *err) {
    const float d = a.depth[p], g = a.gt_depth[p];
    const float e = d - g;
    *err = e;
    return m && a.hit[p] != -1 && g > 0.f && e < a.depth_thr;
}

// Four consecutive pixels of every per-pixel array in 128-bit loads (the images are H*W-contiguous planes; gt_color is
// [H,W,3], so four pixels are three float4).  `vec` is decided on the host: every base pointer 16-byte aligned and
// N % 4 == 0 (true for the 16-pixel-aligned images of the benchmarks); otherwise the scalar path runs.
struct LossPix4 {
    float img[3][4], gt[3][4], d[4], g[4];
    int hit[4];
    bool m[4];
};
__device__ __forceinline__ uchar4 load_mask4(const LossArgs &a, size_t q) {
    return a.mask ? reinterpret_cast<const uchar4 *>(a.mask)[q] : make_uchar4(1, 1, 1, 1);
}
__device__ __forceinline__ void load_pix4(const LossArgs &a, size_t N, size_t q, bool need_depth, uchar4 mk, LossPix4 &x) {
    const float4 i0 = reinterpret_cast<const float4 *>(a.image)[q];
    const float4 i1 = reinterpret_cast<const float4 *>(a.image + N)[q];
    const float4 i2 = reinterpret_cast<const float4 *>(a.image + 2 * N)[q];
    const float4 g0 = reinterpret_cast<const float4 *>(a.gt_color)[3 * q];
    const float4 g1 = reinterpret_cast<const float4 *>(a.gt_color)[3 * q + 1];
    const float4 g2 = reinterpret_cast<const float4 *>(a.gt_color)[3 * q + 2];
    const float iv[3][4] = {{i0.x, i0.y, i0.z, i0.w}, {i1.x, i1.y, i1.z, i1.w}, {i2.x, i2.y, i2.z, i2.w}};
    const float gf[12] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x, g2.y, g2.z, g2.w};
    const uint8_t mv[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x.m[k] = mv[k] != 0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            x.img[c][k] = iv[c][k];
            x.gt[c][k] = gf[3 * k + c];
        }
    }
    if (need_depth) {
        const float4 d = reinterpret_cast<const float4 *>(a.depth)[q];
        const float4 g = reinterpret_cast<const float4 *>(a.gt_depth)[q];
        const int4 h = reinterpret_cast<const int4 *>(a.hit)[q];
        x.d[0] = d.x; x.d[1] = d.y; x.d[2] = d.z; x.d[3] = d.w;
        x.g[0] = g.x; x.g[1] = g.y; x.g[2] = g.z; x.g[3] = g.w;
        x.hit[0] = h.x; x.hit[1] = h.y; x.hit[2] = h.z; x.hit[3] = h.w;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_partial_kernel(LossArgs a, int vec) {
    pdl_enter();
    const size_t N = (size_t)a.W * a.H;
    double cs = 0.0, ds = 0.0;
    double cn = 0.0, dn = 0.0;
    if (vec) {
        const bool need_depth = a.depth_w > 0.f;
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < N / 4; q += (size_t)gridDim.x * blockDim.x) {
            const uchar4 mk = load_mask4(a, q);
            if (!(mk.x | mk.y | mk.z | mk.w)) continue; // no masked-in pixel: contributes to neither term (object masks
                                                        // cover a small part of the image: most quads stop here)
            LossPix4 x;
            load_pix4(a, N, q, need_depth, mk, x);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (x.m[k]) {
                    cs += (double)(fabsf(x.img[0][k] - x.gt[0][k]) + fabsf(x.img[1][k] - x.gt[1][k]) +
                                   fabsf(x.img[2][k] - x.gt[2][k]));
                    cn += 1.0;
                }
                if (need_depth) {
                    const float e = x.d[k] - x.g[k];
                    if (x.m[k] && x.hit[k] != -1 && x.g[k] > 0.f && e < a.depth_thr) {
                        ds += (double)fabsf(e);
                        dn += 1.0;
                    }
                }
            }
        }
    } else {
        for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (size_t)gridDim.x * blockDim.x) {
            const bool m = a.mask ? (a.mask[p] != 0) : true;
            if (m) {
                const float e0 = a.image[p] - a.gt_color[3 * p];
                const float e1 = a.image[N + p] - a.gt_color[3 * p + 1];
                const float e2 = a.image[2 * N + p] - a.gt_color[3 * p + 2];
                cs += (double)(fabsf(e0) + fabsf(e1) + fabsf(e2));
                cn += 1.0;
            }
            if (a.depth_w > 0.f) {
                float e;
                if (depth_valid(a, p, m, &e)) {
                    ds += (double)fabsf(e);
                    dn += 1.0;
                }
            }
        }
    }
    __shared__ double sh[4][LOSS_THREADS / 32];
    double v[4] = {cs, cn, ds, dn};
#pragma unroll
    for (int k = 0; k < 4; k++)
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xFFFFFFFFu, v[k], o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 4; k++) sh[k][warp] = v[k];
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < LOSS_THREADS / 32; w++) t += sh[threadIdx.x][w];
        a.partial[4 * blockIdx.x + threadIdx.x] = t;
    }
}

__device__ __forceinline__ float sgn_scaled(float e, float g) { return (e > 0.f) ? g : ((e < 0.f) ? -g : 0.f); }

__global__ void __launch_bounds__(LOSS_THREADS) loss_grad_kernel(LossArgs a, int nparts, int vec) {
    pdl_enter();
    __shared__ double tot[4];
    { // fixed-pattern (deterministic) final reduction, redundantly per block: warp q sums quantity q
        const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (q < 4) {
            double t = 0.0;
            for (int b = lane; b < nparts; b += 32) t += a.partial[4 * b + q];
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
            if (lane == 0) tot[q] = t;
        }
    }
    __syncthreads();
    const double cs = tot[0], cn = tot[1], ds = tot[2], dn = tot[3];
    const float color_loss = (float)(cs / (3.0 * cn)); // NaN when the selection is empty, like torch.mean
    const float depth_loss = (a.depth_w > 0.f) ? (float)(ds / dn) : 0.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.loss_out[0] = a.depth_w * depth_loss + a.color_w * color_loss;
        a.loss_out[1] = color_loss;
        a.loss_out[2] = depth_loss;
        a.loss_out[3] = 0.f;
        a.counts_out[0] = (int)cn;
        a.counts_out[1] = (int)dn;
    }
    const float gc = (cn > 0.0) ? (float)((double)a.color_w / (3.0 * cn)) : 0.f;
    const float gd = (a.depth_w > 0.f && dn > 0.0) ? (float)((double)a.depth_w / dn) : 0.f;
    const size_t N = (size_t)a.W * a.H;
    if (vec) {
        const bool need_depth = a.depth_w > 0.f;
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < N / 4; q += (size_t)gridDim.x * blockDim.x) {
            const uchar4 mk = load_mask4(a, q);
            if (!(mk.x | mk.y | mk.z | mk.w)) { // all four gradients are zero; nothing else has to be read
                if (a.tile_mask && (a.W & 3) == 0) { // ... and in a tile that is not rendered nobody reads them either
                    const size_t p = 4 * q;
                    const int py = (int)(p / a.W), px = (int)(p - (size_t)py * a.W);
                    if (a.tile_mask[(py >> 4) * a.tiles_x + (px >> 4)] == 0) continue;
                }
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                reinterpret_cast<float4 *>(a.dimg)[q] = z;
                reinterpret_cast<float4 *>(a.dimg + N)[q] = z;
                reinterpret_cast<float4 *>(a.dimg + 2 * N)[q] = z;
                reinterpret_cast<float4 *>(a.ddepth)[q] = z;
                continue;
            }
            LossPix4 x;
            load_pix4(a, N, q, need_depth, mk, x);
            float g[3][4], gdp[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
#pragma unroll
                for (int c = 0; c < 3; c++) g[c][k] = x.m[k] ? sgn_scaled(x.img[c][k] - x.gt[c][k], gc) : 0.f;
                gdp[k] = 0.f;
                if (need_depth) {
                    const float e = x.d[k] - x.g[k];
                    if (x.m[k] && x.hit[k] != -1 && x.g[k] > 0.f && e < a.depth_thr) gdp[k] = sgn_scaled(e, gd);
                }
            }
            reinterpret_cast<float4 *>(a.dimg)[q] = make_float4(g[0][0], g[0][1], g[0][2], g[0][3]);
            reinterpret_cast<float4 *>(a.dimg + N)[q] = make_float4(g[1][0], g[1][1], g[1][2], g[1][3]);
            reinterpret_cast<float4 *>(a.dimg + 2 * N)[q] = make_float4(g[2][0], g[2][1], g[2][2], g[2][3]);
            reinterpret_cast<float4 *>(a.ddepth)[q] = make_float4(gdp[0], gdp[1], gdp[2], gdp[3]);
        }
        return;
    }
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (size_t)gridDim.x * blockDim.x) {
        const bool m = a.mask ? (a.mask[p] != 0) : true;
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, gdp = 0.f;
        if (m) {
            g0 = sgn_scaled(a.image[p] - a.gt_color[3 * p], gc);
            g1 = sgn_scaled(a.image[N + p] - a.gt_color[3 * p + 1], gc);
            g2 = sgn_scaled(a.image[2 * N + p] - a.gt_color[3 * p + 2], gc);
        }
        if (a.depth_w > 0.f) {
            float e;
            if (depth_valid(a, p, m, &e)) gdp = sgn_scaled(e, gd);
        }
        a.dimg[p] = g0;
        a.dimg[N + p] = g1;
        a.dimg[2 * N + p] = g2;
        a.ddepth[p] = gdp;
    }
}

// ------------------------------------------------------------------------------------------------
// fused multi-tensor Adam
// ------------------------------------------------------------------------------------------------
#define ADAM_CHUNK 4096 // elements per block
struct AdamPack {
    float *param[DQO_ADAM_MAX_TENSORS];
    const float *grad[DQO_ADAM_MAX_TENSORS];
    float *m[DQO_ADAM_MAX_TENSORS];
    float *v[DQO_ADAM_MAX_TENSORS];
    long long numel[DQO_ADAM_MAX_TENSORS];
    int chunk_start[DQO_ADAM_MAX_TENSORS + 1];
    float step_size[DQO_ADAM_MAX_TENSORS]; // lr / bias_correction1
    int n;
    float beta1, beta2, one_minus_beta1, one_minus_beta2, bc2_sqrt, eps;
};

__device__ __forceinline__ void adam_elem(float &p, float g, float &m, float &v, const AdamPack &k, float step_size) {
    // torch/optim/adam.py _single_tensor_adam: lerp_, mul_/addcmul_, sqrt/div/add_, addcdiv_
    m = m + k.one_minus_beta1 * (g - m);
    v = v * k.beta2 + k.one_minus_beta2 * g * g;
    const float denom = sqrtf(v) / k.bc2_sqrt + k.eps;
    p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(AdamPack k) {
    pdl_enter();
    int t = 0;
    const int chunk = blockIdx.x;
    while (t + 1 < k.n && chunk >= k.chunk_start[t + 1]) t++;
    const long long base = (long long)(chunk - k.chunk_start[t]) * ADAM_CHUNK;
    const long long n = k.numel[t];
    float *__restrict__ P = k.param[t];
    const float *__restrict__ G = k.grad[t];
    float *__restrict__ Mo = k.m[t];
    float *__restrict__ Vo = k.v[t];
    const float ss = k.step_size[t];
    const bool aligned = ((((uintptr_t)P) | ((uintptr_t)G) | ((uintptr_t)Mo) | ((uintptr_t)Vo)) & 15) == 0;
    if (aligned && base + ADAM_CHUNK <= n) {
#pragma unroll
        for (int r = 0; r < ADAM_CHUNK / (256 * 4); r++) {
            const long long i = base + (long long)(r * 256 + threadIdx.x) * 4;
            float4 p = *reinterpret_cast<float4 *>(P + i);
            const float4 g = *reinterpret_cast<const float4 *>(G + i);
            float4 m = *reinterpret_cast<float4 *>(Mo + i);
            float4 v = *reinterpret_cast<float4 *>(Vo + i);
            adam_elem(p.x, g.x, m.x, v.x, k, ss);
            adam_elem(p.y, g.y, m.y, v.y, k, ss);
            adam_elem(p.z, g.z, m.z, v.z, k, ss);
            adam_elem(p.w, g.w, m.w, v.w, k, ss);
            *reinterpret_cast<float4 *>(P + i) = p;
            *reinterpret_cast<float4 *>(Mo + i) = m;
            *reinterpret_cast<float4 *>(Vo + i) = v;
        }
    } else {
        const long long end = (base + ADAM_CHUNK < n) ? base + ADAM_CHUNK : n;
        for (long long i = base + threadIdx.x; i < end; i += 256) {
            float p = P[i], m = Mo[i], v = Vo[i];
            adam_elem(p, G[i], m, v, k, ss);
            P[i] = p;
            Mo[i] = m;
            Vo[i] = v;
        }
    }
}

__global__ void __launch_bounds__(256) confidence_kernel(long long rows, int width, const float *__restrict__ grad, float *conf) {
    pdl_enter();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    bool any = false;
    for (int c = 0; c < width; c++) any |= (fabsf(grad[i * width + c]) != 0.f);
    if (any) conf[i] += 1.0f;
}

// ------------------------------------------------------------------------------------------------
// per-Gaussian error scatter
// ------------------------------------------------------------------------------------------------
struct AccErrArgs {
    int W, H, P;
    const float *ce, *de, *ne;
    const int *ci, *di;
    float cthr, dthr, nthr;
    int check_max;
    float *gce, *gde, *gne, *resc;
    int *cc, *dc, *nc;
};
__device__ __forceinline__ void scatter_max(float *addr, float val) {
    // the reference's CAS loop (map_process.cu:8-18) stores max(current, val) for val > current with the
    // arrays initialised to 0, i.e. only positive values ever land: an integer atomicMax on the bits is identical
    if (val > 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(val));
}
__global__ void __launch_bounds__(256) acc_error_kernel(AccErrArgs a) {
    pdl_enter();
    const size_t N = (size_t)a.W * a.H;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float ce = a.ce[p], de = a.de[p], ne = a.ne[p];
    const int ci = a.ci[p], di = a.di[p];
    if (ci >= 0 && ci < a.P) {
        if (a.check_max)
            scatter_max(&a.gce[ci], ce);
        else
            atomicAdd(&a.gce[ci], ce);
        atomicAdd(&a.cc[ci], 1);
        if (ce > a.cthr) atomicAdd(&a.resc[ci], 1.0f);
    }
    if (di >= 0 && di < a.P) {
        if (a.check_max) {
            scatter_max(&a.gde[di], de);
            scatter_max(&a.gne[di], ne);
        } else {
            atomicAdd(&a.gde[di], de);
            atomicAdd(&a.gne[di], ne);
        }
        atomicAdd(&a.dc[di], 1);
        atomicAdd(&a.nc[di], 1);
        if (de > a.dthr) atomicAdd(&a.resc[di], 1.0f);
        if (ne > a.nthr) atomicAdd(&a.resc[di], 1.0f);
    }
}
__global__ void __launch_bounds__(256) acc_error_mean_kernel(AccErrArgs a) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.P) return;
    const int c = a.cc[i], d = a.dc[i], n = a.nc[i];
    if (c > 0) a.gce[i] = a.gce[i] / c;
    if (d > 0) a.gde[i] = a.gde[i] / d;
    if (n > 0) a.gne[i] = a.gne[i] / n;
}

} // namespace dqo

using namespace dqo;

extern "C" size_t dqo_loss_workspace_bytes(int32_t W, int32_t H) {
    (void)W;
    (void)H;
    return (size_t)LOSS_BLOCKS * 4 * sizeof(double);
}

namespace dqo {
int masked_l1_loss_impl(int32_t W, int32_t H, const float *image, const float *depth, const int32_t *hit_depth,
                        const float *gt_color, const float *gt_depth, const uint8_t *render_mask, float color_weight,
                        float depth_weight, float depth_err_thres, float *dL_dimage, float *dL_ddepth, float *loss_out,
                        int32_t *counts_out, void *workspace, const int32_t *tile_mask, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (W <= 0 || H <= 0 || !image || !depth || !hit_depth || !gt_color || !gt_depth || !dL_dimage || !dL_ddepth ||
        !loss_out || !counts_out || !workspace) {
        set_error("dqo_masked_l1_loss: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    LossArgs a;
    a.W = W; a.H = H; a.image = image; a.depth = depth; a.gt_color = gt_color; a.gt_depth = gt_depth;
    a.hit = hit_depth; a.mask = render_mask; a.color_w = color_weight; a.depth_w = depth_weight;
    a.depth_thr = depth_err_thres; a.dimg = dL_dimage; a.ddepth = dL_ddepth; a.loss_out = loss_out;
    a.counts_out = counts_out; a.partial = (double *)workspace;
    a.tile_mask = tile_mask; a.tiles_x = (W + 15) / 16;
    const size_t N = (size_t)W * H;
    int blocks = (int)((N + LOSS_THREADS - 1) / LOSS_THREADS);
    if (blocks > LOSS_BLOCKS) blocks = LOSS_BLOCKS;
    // 128-bit path: every array 16-byte aligned and a whole number of 4-pixel groups
    int vec = (N % 4 == 0);
    for (const void *q : {(const void *)image, (const void *)depth, (const void *)hit_depth, (const void *)gt_color,
                          (const void *)gt_depth, (const void *)dL_dimage, (const void *)dL_ddepth})
        vec &= ((uintptr_t)q % 16 == 0);
    vec &= ((uintptr_t)render_mask % 4 == 0) && ((N * sizeof(float)) % 16 == 0);
    if (vec) {
        blocks = (int)((N / 4 + LOSS_THREADS - 1) / LOSS_THREADS);
        if (blocks > LOSS_BLOCKS) blocks = LOSS_BLOCKS;
        if (blocks < 1) blocks = 1;
    }
    launch_pdl(loss_partial_kernel, dim3(blocks), dim3(LOSS_THREADS), 0, stream, a, vec);
    launch_pdl(loss_grad_kernel, dim3(blocks), dim3(LOSS_THREADS), 0, stream, a, blocks, vec);
    DQO_LAUNCH_CHECK("masked l1 loss", 0, stream);
    return DQO_OK;
}
} // namespace dqo

extern "C" int dqo_masked_l1_loss(int32_t W, int32_t H, const float *image, const float *depth,
                                  const int32_t *hit_depth, const float *gt_color, const float *gt_depth,
                                  const uint8_t *render_mask, float color_weight, float depth_weight,
                                  float depth_err_thres, float *dL_dimage, float *dL_ddepth, float *loss_out,
                                  int32_t *counts_out, void *workspace, void *stream_) {
    return masked_l1_loss_impl(W, H, image, depth, hit_depth, gt_color, gt_depth, render_mask, color_weight, depth_weight,
                               depth_err_thres, dL_dimage, dL_ddepth, loss_out, counts_out, workspace, nullptr, stream_);
}

extern "C" int dqo_adam_step(const dqo_adam_tensor *tensors, int32_t n_tensors, int32_t step, double beta1, double beta2,
                             double eps, float *confidence, int32_t conf_tensor, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!tensors || n_tensors < 0 || n_tensors > DQO_ADAM_MAX_TENSORS || step < 1) {
        set_error("dqo_adam_step: invalid argument (n_tensors=%d, step=%d)", n_tensors, step);
        return DQO_ERR_INVALID_ARG;
    }
    AdamPack k;
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    k.beta1 = (float)beta1; k.beta2 = (float)beta2;
    k.one_minus_beta1 = (float)(1.0 - beta1);
    k.one_minus_beta2 = (float)(1.0 - beta2);
    k.bc2_sqrt = (float)sqrt(bc2);
    k.eps = (float)eps;
    int n = 0, chunks = 0;
    for (int i = 0; i < n_tensors; i++) {
        const dqo_adam_tensor &t = tensors[i];
        if (t.numel <= 0 || !t.grad) continue; // torch skips parameters without a gradient
        if (!t.param || !t.exp_avg || !t.exp_avg_sq) {
            set_error("dqo_adam_step: tensor %d has null state", i);
            return DQO_ERR_INVALID_ARG;
        }
        k.param[n] = t.param; k.grad[n] = t.grad; k.m[n] = t.exp_avg; k.v[n] = t.exp_avg_sq;
        k.numel[n] = t.numel;
        k.step_size[n] = (float)(t.lr / bc1);
        k.chunk_start[n] = chunks;
        chunks += (int)((t.numel + ADAM_CHUNK - 1) / ADAM_CHUNK);
        n++;
    }
    k.chunk_start[n] = chunks;
    k.n = n;
    if (chunks > 0) {
        launch_pdl(adam_kernel, dim3(chunks), dim3(256), 0, stream, k);
        DQO_LAUNCH_CHECK("adam", 0, stream);
    }
    if (confidence && conf_tensor >= 0 && conf_tensor < n_tensors) {
        const dqo_adam_tensor &t = tensors[conf_tensor];
        if (t.grad && t.row_width > 0 && t.numel > 0) {
            const long long rows = t.numel / t.row_width;
            launch_pdl(confidence_kernel, dim3((unsigned)((rows + 255) / 256)), dim3(256), 0, stream, rows, t.row_width, t.grad, confidence);
            DQO_LAUNCH_CHECK("confidence", 0, stream);
        }
    }
    return DQO_OK;
}

extern "C" int dqo_accumulate_error(int32_t W, int32_t H, int32_t P, const float *color_err, const float *depth_err,
                                    const float *normal_err, const int32_t *color_index, const int32_t *depth_index,
                                    float color_thr, float depth_thr, float normal_thr, int32_t check_max,
                                    float *gs_color_error, float *gs_depth_error, float *gs_normal_error,
                                    int32_t *color_counter, int32_t *depth_counter, int32_t *normal_counter,
                                    float *rescale_counter, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (W < 0 || H < 0 || P < 0) {
        set_error("dqo_accumulate_error: negative size");
        return DQO_ERR_INVALID_ARG;
    }
    if (P == 0) return DQO_OK;
    if (!gs_color_error || !gs_depth_error || !gs_normal_error || !color_counter || !depth_counter || !normal_counter ||
        !rescale_counter) {
        set_error("dqo_accumulate_error: null output");
        return DQO_ERR_INVALID_ARG;
    }
    const size_t pb = (size_t)P * 4;
    DQO_CUDA_CHECK(cudaMemsetAsync(gs_color_error, 0, pb, stream));
    DQO_CUDA_CHECK(cudaMemsetAsync(gs_depth_error, 0, pb, stream));
    DQO_CUDA_CHECK(cudaMemsetAsync(gs_normal_error, 0, pb, stream));
    DQO_CUDA_CHECK(cudaMemsetAsync(color_counter, 0, pb, stream));
    DQO_CUDA_CHECK(cudaMemsetAsync(depth_counter, 0, pb, stream));
    DQO_CUDA_CHECK(cudaMemsetAsync(normal_counter, 0, pb, stream));
    DQO_CUDA_CHECK(cudaMemsetAsync(rescale_counter, 0, pb, stream));
    const size_t N = (size_t)W * H;
    if (N == 0) return DQO_OK;
    if (!color_err || !depth_err || !normal_err || !color_index || !depth_index) {
        set_error("dqo_accumulate_error: null input");
        return DQO_ERR_INVALID_ARG;
    }
    AccErrArgs a;
    a.W = W; a.H = H; a.P = P; a.ce = color_err; a.de = depth_err; a.ne = normal_err; a.ci = color_index;
    a.di = depth_index; a.cthr = color_thr; a.dthr = depth_thr; a.nthr = normal_thr; a.check_max = check_max;
    a.gce = gs_color_error; a.gde = gs_depth_error; a.gne = gs_normal_error; a.resc = rescale_counter;
    a.cc = color_counter; a.dc = depth_counter; a.nc = normal_counter;
    launch_pdl(acc_error_kernel, dim3((unsigned)((N + 255) / 256)), dim3(256), 0, stream, a);
    if (!check_max) launch_pdl(acc_error_mean_kernel, dim3((P + 255) / 256), dim3(256), 0, stream, a);
    DQO_LAUNCH_CHECK("accumulate error", 0, stream);
    return DQO_OK;
}
