// Mask builders and error maps of the mapper: the small image-space passes that sit between two
// renders of the mapping loop.  The reference runs each as a chain of torch ops (pad, avg_pool2d, compare, topk, index
// scatter, interpolate ...); here each is one pass over the image.
//   transmission2tilemask   SLAM/utils.py:752-763
//   colorerror2tilemask     SLAM/utils.py:765-796
//   evaluate_render_range   SLAM/multiprocess/mapper.py:930-987
//   error maps of error_gaussians_remove   SLAM/multiprocess/mapper.py:1008-1025
// Tiles are the rasterizer's 16x16 blocks; callers always pass stride 16 (mapper.py:956,968,985).
#include "common.cuh"

namespace dqo {

constexpr int TILE = 16;

// One block per tile.  SRC = 0: pixel mask given as bytes; SRC = 1: mask derived as T_map != 1 and written out.
// avg_pool2d over a zero-padded 0/1 image is count/256 exactly, so the compare is done on that float.
template <int SRC>
__global__ void __launch_bounds__(256) tilemask_kernel(int W, int H, const void *src, float ratio, uint8_t *render_mask,
                                                       int32_t *tile_mask, int32_t *render_count) {
    const int tx = blockIdx.x, ty = blockIdx.y;
    const int x = tx * TILE + (threadIdx.x & 15), y = ty * TILE + (threadIdx.x >> 4);
    int m = 0;
    if (x < W && y < H) {
        const size_t p = (size_t)y * W + x;
        if (SRC == 0) {
            m = static_cast<const uint8_t *>(src)[p] != 0;
        } else {
            m = static_cast<const float *>(src)[p] != 1.0f;
            if (render_mask) render_mask[p] = (uint8_t)m;
        }
    }
    const int cnt = __syncthreads_count(m);
    if (threadIdx.x == 0) {
        tile_mask[ty * gridDim.x + tx] = ((float)cnt * (1.0f / 256.0f) > ratio) ? 1 : 0;
        if (render_count && cnt) atomicAdd(render_count, cnt);
    }
}

// color_error = sum_c |render - gt| with pixels whose rendered colour sums to 0 forced to 0 (mapper.py:948-955);
// CHW inputs, HW output.  The channel sums follow torch's order (c0 + c1) + c2.
__global__ void __launch_bounds__(256) color_error_kernel(size_t N, const float *render, const float *gt, float *err) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float r0 = render[p], r1 = render[N + p], r2 = render[2 * N + p];
    const float e = fadd(fadd(fabsf(fsub(r0, gt[p])), fabsf(fsub(r1, gt[N + p]))), fabsf(fsub(r2, gt[2 * N + p])));
    err[p] = (fadd(fadd(r0, r1), r2) == 0.0f) ? 0.0f : e;
}

// Tile means in avg_pool2d's own accumulation order (row by row, left to right, float accumulator, padded zeros last,
// then / 256) so that the ranking below sees the same floats as torch.topk does in the reference.
__global__ void __launch_bounds__(128) tile_mean_kernel(int W, int H, int tw, int th, const float *err, float *mean) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tw * th) return;
    const int x0 = (t % tw) * TILE, y0 = (t / tw) * TILE;
    float acc = 0.0f;
    for (int dy = 0; dy < TILE; dy++) {
        const int y = y0 + dy;
        if (y >= H) break;
        const float *row = err + (size_t)y * W;
        for (int dx = 0; dx < TILE; dx++) {
            const int x = x0 + dx;
            if (x < W) acc = fadd(acc, row[x]);
        }
    }
    mean[t] = acc / 256.0f;
}

// k largest tile means -> mask.  rank(i) = #{j : v_j > v_i or (v_j == v_i and j < i)}; n is a few thousand, the
// quadratic count over a shared-memory copy is a few microseconds and needs no sort workspace.  (torch.topk leaves the
// order among equal values unspecified; lower tile index wins here.)
__global__ void __launch_bounds__(256) topk_mask_kernel(int n, int k, const float *mean, int32_t *tile_mask, int or_into) {
    extern __shared__ float s_v[];
    for (int j = threadIdx.x; j < n; j += blockDim.x) s_v[j] = mean[j];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = s_v[i];
    int rank = 0;
    for (int j = 0; j < n; j++) {
        const float u = s_v[j];
        rank += (u > v) || (u == v && j < i);
    }
    const int sel = rank < k;
    tile_mask[i] = or_into ? (tile_mask[i] | sel) : sel;
}

// nearest-neighbour x16 upsampling of the tile mask cropped to the image (mapper.py:971-980)
__global__ void __launch_bounds__(256) tile_to_pixel_kernel(int W, int H, int tw, const int32_t *tile_mask, uint8_t *render_mask) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)W * H) return;
    const int x = (int)(p % W), y = (int)(p / W);
    render_mask[p] = tile_mask[(y >> 4) * tw + (x >> 4)] != 0;
}

// Error maps of error_gaussians_remove (mapper.py:1008-1025): gt maps are HWC (frame_map), renders are CHW.
struct ErrMapArgs {
    int W, H;
    const float *render_color, *render_depth, *gt_color, *gt_depth;
    const int32_t *depth_index;
    float *color_error, *depth_error, *normal_error;
};
__global__ void __launch_bounds__(256) error_maps_kernel(ErrMapArgs a) {
    const size_t N = (size_t)a.W * a.H;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float gd = a.gt_depth[p], d = a.render_depth[p];
    const float diff = fsub(gd, d);
    const bool invalid = (gd == 0.0f) || (a.depth_index[p] == -1);
    a.depth_error[p] = (invalid || diff < 0.0f) ? 0.0f : fabsf(diff);
    const float e0 = fabsf(fsub(a.gt_color[3 * p], a.render_color[p]));
    const float e1 = fabsf(fsub(a.gt_color[3 * p + 1], a.render_color[N + p]));
    const float e2 = fabsf(fsub(a.gt_color[3 * p + 2], a.render_color[2 * N + p]));
    a.color_error[p] = (gd == 0.0f) ? 0.0f : fadd(fadd(e0, e1), e2);
    if (a.normal_error) a.normal_error[p] = 0.0f;
}

static bool bad_image(int W, int H) { return W <= 0 || H <= 0; }

} // namespace dqo

using namespace dqo;

extern "C" int dqo_render_range(int32_t W, int32_t H, const float *T_map, float tile_mask_ratio, uint8_t *render_mask,
                                int32_t *tile_mask, int32_t *render_count, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bad_image(W, H) || !T_map || !tile_mask) {
        set_error("dqo_render_range: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    if (render_count) cudaMemsetAsync(render_count, 0, sizeof(int32_t), stream);
    dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE);
    tilemask_kernel<1><<<grid, 256, 0, stream>>>(W, H, T_map, tile_mask_ratio, render_mask, tile_mask, render_count);
    DQO_LAUNCH_CHECK("render range", 0, stream);
    return DQO_OK;
}

extern "C" int dqo_pixelmask_to_tilemask(int32_t W, int32_t H, const uint8_t *pixelmask, float tile_mask_ratio,
                                         int32_t *tile_mask, int32_t *pixel_count, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bad_image(W, H) || !pixelmask || !tile_mask) {
        set_error("dqo_pixelmask_to_tilemask: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    if (pixel_count) cudaMemsetAsync(pixel_count, 0, sizeof(int32_t), stream);
    dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE);
    tilemask_kernel<0><<<grid, 256, 0, stream>>>(W, H, pixelmask, tile_mask_ratio, nullptr, tile_mask, pixel_count);
    DQO_LAUNCH_CHECK("pixel mask to tile mask", 0, stream);
    return DQO_OK;
}

extern "C" int dqo_color_error_map(int32_t W, int32_t H, const float *render, const float *gt, float *color_error,
                                   void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bad_image(W, H) || !render || !gt || !color_error) {
        set_error("dqo_color_error_map: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    const size_t N = (size_t)W * H;
    color_error_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(N, render, gt, color_error);
    DQO_LAUNCH_CHECK("color error map", 0, stream);
    return DQO_OK;
}

extern "C" size_t dqo_topk_tilemask_workspace_bytes(int32_t W, int32_t H) {
    if (W <= 0 || H <= 0) return 0;
    return align_up((size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE) * sizeof(float), 256);
}

extern "C" int dqo_topk_tilemask(int32_t W, int32_t H, const float *error, int32_t k, int32_t or_into,
                                 int32_t *tile_mask, uint8_t *render_mask, void *workspace, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bad_image(W, H) || !error || !tile_mask || !workspace || k < 0) {
        set_error("dqo_topk_tilemask: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    const int tw = (W + TILE - 1) / TILE, th = (H + TILE - 1) / TILE, n = tw * th;
    const size_t smem = (size_t)n * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("dqo_topk_tilemask: %d tiles exceed the shared-memory ranking limit", n);
        return DQO_ERR_INVALID_ARG;
    }
    float *mean = (float *)workspace;
    tile_mean_kernel<<<(n + 127) / 128, 128, 0, stream>>>(W, H, tw, th, error, mean);
    DQO_LAUNCH_CHECK("tile mean", 0, stream);
    if (smem > 48 * 1024) cudaFuncSetAttribute(topk_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    topk_mask_kernel<<<(n + 255) / 256, 256, smem, stream>>>(n, k, mean, tile_mask, or_into);
    DQO_LAUNCH_CHECK("top-k tile mask", 0, stream);
    if (render_mask) {
        const size_t N = (size_t)W * H;
        tile_to_pixel_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(W, H, tw, tile_mask, render_mask);
        DQO_LAUNCH_CHECK("tile mask to pixel mask", 0, stream);
    }
    return DQO_OK;
}

extern "C" int dqo_tilemask_to_pixelmask(int32_t W, int32_t H, const int32_t *tile_mask, uint8_t *render_mask,
                                         void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bad_image(W, H) || !tile_mask || !render_mask) {
        set_error("dqo_tilemask_to_pixelmask: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    const size_t N = (size_t)W * H;
    tile_to_pixel_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(W, H, (W + TILE - 1) / TILE, tile_mask, render_mask);
    DQO_LAUNCH_CHECK("tile mask to pixel mask", 0, stream);
    return DQO_OK;
}

extern "C" int dqo_render_error_maps(int32_t W, int32_t H, const float *render_color, const float *render_depth,
                                     const float *gt_color_hwc, const float *gt_depth, const int32_t *depth_index,
                                     float *color_error, float *depth_error, float *normal_error, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bad_image(W, H) || !render_color || !render_depth || !gt_color_hwc || !gt_depth || !depth_index || !color_error ||
        !depth_error) {
        set_error("dqo_render_error_maps: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    ErrMapArgs a;
    a.W = W; a.H = H; a.render_color = render_color; a.render_depth = render_depth; a.gt_color = gt_color_hwc;
    a.gt_depth = gt_depth; a.depth_index = depth_index; a.color_error = color_error; a.depth_error = depth_error;
    a.normal_error = normal_error;
    const size_t N = (size_t)W * H;
    error_maps_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(a);
    DQO_LAUNCH_CHECK("render error maps", 0, stream);
    return DQO_OK;
}
