// Map-geometry kernels around distCUDA2:
//   dqo_bbox_mask      bbox_filter (SLAM/utils.py:801-808): which points of a cloud lie strictly inside the padded
//                      bounding box of another cloud; the bounding box is reduced and consumed on the device;
//   dqo_gaussian_radius GaussianPointCloud.get_radius (SLAM/gaussian_pointcloud.py:739-743) from the log-scales;
//   dqo_scale_init     the arithmetic of GaussianPointCloud.update_geometry (gaussian_pointcloud.py:540-569) after the
//                      kNN: distances to the 3 neighbours minus 3 x their radii, RMS, clip, x scale_factor x xyz_factor,
//                      log; plus the "delete" mask (any distance < 0) and the number of survivors -- one pass instead of
//                      ~25 torch kernels with three [P,3] gathers;
//   dqo_knn_cross3     the K = 3 nearest points of a REFERENCE cloud for every point of a QUERY cloud, as
//                      Mapping.temp_points_filter needs them (mapper.py:1351-1380; pytorch3d.ops.knn_points in the
//                      reference -- a third-party op that is not part of /root/reference, restated from its documented
//                      contract: squared L2 distances, ascending, with the reference-cloud indices), and
//   dqo_inside_mask    temp_points_filter's test `(sqrt(d2) < 0.6 * radius[idx]).any(-1)` on that result.
// The cross-cloud search reuses the Morton machinery of knn.cu's design: both clouds are ordered along the same
// Morton curve, 256 consecutive queries share a block, candidate boxes of 1024 reference points are staged in shared
// memory and culled by exact box distance, 32-point sub-boxes are culled again.
#include "common.cuh"
#include "sort.cuh"
#include <float.h>
#include <limits.h>

namespace dqo {

__device__ __forceinline__ uint32_t g_enc(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float g_dec(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

// order-encoded min / max of a cloud; mm[0..2] = min (init +inf), mm[3..5] = max (init -inf)
__global__ void geo_minmax_init_kernel(uint32_t *mm) {
    if (threadIdx.x < 3) mm[threadIdx.x] = g_enc(INFINITY);
    else if (threadIdx.x < 6) mm[threadIdx.x] = g_enc(-INFINITY);
}
__global__ void __launch_bounds__(256) geo_minmax_kernel(int n, const float *__restrict__ pts, uint32_t *mm) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float v = pts[3 * (size_t)i + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xFFFFFFFFu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xFFFFFFFFu, mx[c], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            atomicMin(&mm[c], g_enc(mn[c]));
            atomicMax(&mm[3 + c], g_enc(mx[c]));
        }
    }
}
// (total > local_min - padding).all() & (total < local_max + padding).all()   (SLAM/utils.py:802-806)
__global__ void __launch_bounds__(256) bbox_mask_kernel(int n, const float *__restrict__ pts, const uint32_t *__restrict__ mm,
                                                        float padding, uint8_t *mask, int *count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    if (i < n) {
        in = true;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float lo = fsub(g_dec(mm[c]), padding), hi = fadd(g_dec(mm[3 + c]), padding);
            const float v = pts[3 * (size_t)i + c];
            in = in && (v > lo) && (v < hi);
        }
        mask[i] = in ? 1 : 0;
    }
    if (count) {
        const unsigned b = __ballot_sync(0xFFFFFFFFu, in);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
    }
}

// radius = (sum(exp(s)) - min(exp(s))) / 2   (gaussian_pointcloud.py:739-743)
__global__ void __launch_bounds__(256) radius_kernel(int n, const float *__restrict__ log_scales, float *radius) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = expf(log_scales[3 * (size_t)i]), b = expf(log_scales[3 * (size_t)i + 1]), c = expf(log_scales[3 * (size_t)i + 2]);
    const float mn = fminf(fminf(a, b), c);
    radius[i] = fmul(fsub(fadd(fadd(a, b), c), mn), 0.5f);
}

struct ScaleInitArgs {
    int n_new, n_total;
    const float *xyz, *radius;
    const int *knn_idx;
    float min_radius, max_radius, scale_factor, fx, fy, fz;
    float *log_scales;
    uint8_t *invalid;
    int *valid_count;
};
__global__ void __launch_bounds__(256) scale_init_kernel(ScaleInitArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    if (i < a.n_new) {
        const float px = a.xyz[3 * (size_t)i], py = a.xyz[3 * (size_t)i + 1], pz = a.xyz[3 * (size_t)i + 2];
        float sum2 = 0.f;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int j = a.knn_idx[3 * (size_t)i + k];
            if (j < 0 || j >= a.n_total) { // fewer than 3 neighbours exist (the reference would index out of range)
                bad = true;
                continue;
            }
            const float dx = fsub(px, a.xyz[3 * (size_t)j]), dy = fsub(py, a.xyz[3 * (size_t)j + 1]),
                        dz = fsub(pz, a.xyz[3 * (size_t)j + 2]);
            // torch.norm(p=2, dim=1) then minus 3 x the neighbour's radius (gaussian_pointcloud.py:541-552)
            const float d = fsub(fsqrt(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz))), fmul(3.0f, a.radius[j]));
            bad = bad || (d < 0.f);
            sum2 = fadd(sum2, fmul(d, d));
        }
        float s = fsqrt(fdiv(sum2, 3.0f));
        s = fminf(fmaxf(s, a.min_radius), a.max_radius);
        a.log_scales[3 * (size_t)i] = logf(fmul(a.scale_factor, fmul(s, a.fx)));
        a.log_scales[3 * (size_t)i + 1] = logf(fmul(a.scale_factor, fmul(s, a.fy)));
        a.log_scales[3 * (size_t)i + 2] = logf(fmul(a.scale_factor, fmul(s, a.fz)));
        a.invalid[i] = bad ? 1 : 0;
        valid = !bad;
    }
    if (a.valid_count) {
        const unsigned b = __ballot_sync(0xFFFFFFFFu, valid);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(a.valid_count, __popc(b));
    }
}

// ------------------------------------------------------------------------------------------------
// cross-cloud 3-NN
// ------------------------------------------------------------------------------------------------
#define GX_BOX 1024
#define GX_SUB 32

struct CrossLayout {
    size_t mm, r_codes, r_codes2, r_ids, r_ids2, q_codes, q_codes2, q_ids, q_ids2, r_sp, q_sp, boxes, subboxes, sort_temp, total;
};
static size_t gbump(size_t &cur, size_t bytes) {
    size_t off = align_up(cur, 256);
    cur = off + bytes;
    return off;
}
static void make_cross_layout(int nq, int nr, CrossLayout *L) {
    size_t cur = 0;
    const size_t q = (size_t)(nq > 0 ? nq : 1), r = (size_t)(nr > 0 ? nr : 1);
    L->mm = gbump(cur, 32);
    L->r_codes = gbump(cur, r * 4);
    L->r_codes2 = gbump(cur, r * 4);
    L->r_ids = gbump(cur, r * 4);
    L->r_ids2 = gbump(cur, r * 4);
    L->q_codes = gbump(cur, q * 4);
    L->q_codes2 = gbump(cur, q * 4);
    L->q_ids = gbump(cur, q * 4);
    L->q_ids2 = gbump(cur, q * 4);
    L->r_sp = gbump(cur, r * 16);
    L->q_sp = gbump(cur, q * 16);
    L->boxes = gbump(cur, ((r + GX_BOX - 1) / GX_BOX) * 24);
    L->subboxes = gbump(cur, ((r + GX_SUB - 1) / GX_SUB) * 24);
    SortTemp T;
    make_sort_temp((int64_t)(q > r ? q : r), 30, &T);
    L->sort_temp = gbump(cur, T.total);
    L->total = align_up(cur, 256);
}

__device__ __forceinline__ uint32_t g_spread(uint32_t x) {
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}
// Morton code on the grid of the (joint) bounding box mm; any consistent curve works -- it only orders the search
__global__ void __launch_bounds__(256) cross_morton_kernel(int n, const float *__restrict__ pts, const uint32_t *__restrict__ mm,
                                                           uint32_t *codes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t m[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float mn = g_dec(mm[c]), mx = g_dec(mm[3 + c]);
        const float ext = fmaxf(mx - mn, 1e-30f);
        float t = (pts[3 * (size_t)i + c] - mn) / ext * 1023.0f;
        t = fminf(fmaxf(t, 0.f), 1023.f);
        m[c] = g_spread((uint32_t)t);
    }
    codes[i] = m[0] | (m[1] << 1) | (m[2] << 2);
}
__global__ void __launch_bounds__(256) cross_gather_kernel(int n, const float *__restrict__ pts, const uint32_t *__restrict__ ids,
                                                           float4 *sp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = ids[i];
    sp[i] = make_float4(pts[3 * (size_t)id], pts[3 * (size_t)id + 1], pts[3 * (size_t)id + 2], __uint_as_float(id));
}
__global__ void __launch_bounds__(1024) cross_boxes_kernel(int n, const float4 *__restrict__ sp, float *boxes, float *subboxes) {
    __shared__ float s_mn[32][3], s_mx[32][3];
    const int i = blockIdx.x * GX_BOX + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        const float4 p = sp[i];
        mn[0] = mx[0] = p.x;
        mn[1] = mx[1] = p.y;
        mn[2] = mx[2] = p.z;
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xFFFFFFFFu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xFFFFFFFFu, mx[c], o));
        }
    if (lane == 0) {
        const int sb = blockIdx.x * (GX_BOX / GX_SUB) + warp;
        if ((size_t)sb * GX_SUB < (size_t)n) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                subboxes[6 * (size_t)sb + c] = mn[c];
                subboxes[6 * (size_t)sb + 3 + c] = mx[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            s_mn[warp][c] = mn[c];
            s_mx[warp][c] = mx[c];
        }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float a = s_mn[lane][c], b = s_mx[lane][c];
            for (int o = 16; o > 0; o >>= 1) {
                a = fminf(a, __shfl_xor_sync(0xFFFFFFFFu, a, o));
                b = fmaxf(b, __shfl_xor_sync(0xFFFFFFFFu, b, o));
            }
            if (lane == 0) {
                boxes[6 * (size_t)blockIdx.x + c] = a;
                boxes[6 * (size_t)blockIdx.x + 3 + c] = b;
            }
        }
    }
}
__device__ __forceinline__ float g_box_dist(const float *__restrict__ b, float px, float py, float pz) {
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (px < b[0]) dx = b[0] - px; else if (px > b[3]) dx = px - b[3];
    if (py < b[1]) dy = b[1] - py; else if (py > b[4]) dy = py - b[4];
    if (pz < b[2]) dz = b[2] - pz; else if (pz > b[5]) dz = pz - b[5];
    // a lower bound of every member's squared distance AS EVALUATED by g_update: the same operations in the same order,
    // and float subtraction / multiplication / addition are monotone in each operand
    return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}
__device__ __forceinline__ void g_update(float px, float py, float pz, const float4 c, float best[3], int bidx[3]) {
    const float dx = px - c.x, dy = py - c.y, dz = pz - c.z;
    float dist = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
    int id = (int)__float_as_uint(c.w);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        // ties: lower reference index first (deterministic whatever the visiting order)
        if (best[j] > dist || (best[j] == dist && bidx[j] > id)) {
            const float t = best[j];
            best[j] = dist;
            dist = t;
            const int ti = bidx[j];
            bidx[j] = id;
            id = ti;
        }
    }
}
// position of the first sorted reference code >= code
__device__ __forceinline__ int g_lower_bound(const uint32_t *__restrict__ codes, int n, uint32_t code) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (codes[mid] < code) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__global__ void __launch_bounds__(256)
    cross_search_kernel(int nq, int nr, const float4 *__restrict__ q_sp, const uint32_t *__restrict__ q_codes,
                        const float4 *__restrict__ r_sp, const uint32_t *__restrict__ r_codes, const float *__restrict__ boxes,
                        const float *__restrict__ subboxes, int nboxes, float *dist2, int *idx) {
    __shared__ float4 s_pts[GX_BOX];
    __shared__ float s_sub[(GX_BOX / GX_SUB) * 6];
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = qi < nq;
    float4 me = make_float4(0, 0, 0, 0);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    int bidx[3] = {INT_MAX, INT_MAX, INT_MAX};
    if (live) {
        me = q_sp[qi];
        // seed bound from the reference points around the query's place on the curve (discarded afterwards: every
        // candidate is then visited exactly once, in box order)
        const int pos = g_lower_bound(r_codes, nr, q_codes[qi]);
        for (int i = max(0, pos - 4); i < min(nr, pos + 4); i++) g_update(me.x, me.y, me.z, r_sp[i], best, bidx);
    }
    const float reject = best[2];
    best[0] = best[1] = best[2] = FLT_MAX;
    bidx[0] = bidx[1] = bidx[2] = INT_MAX;
    for (int b = 0; b < nboxes; b++) {
        bool need = false;
        if (live) {
            const float d = g_box_dist(boxes + 6 * (size_t)b, me.x, me.y, me.z);
            need = !(d > reject || d > best[2]);
        }
        if (!__syncthreads_or(need)) continue;
        const int base = b * GX_BOX;
        const int count = min(GX_BOX, nr - base);
        for (int t = threadIdx.x; t < count; t += blockDim.x) s_pts[t] = r_sp[base + t];
        const int nsub = (count + GX_SUB - 1) / GX_SUB;
        for (int t = threadIdx.x; t < nsub * 6; t += blockDim.x) s_sub[t] = subboxes[(size_t)(base / GX_SUB) * 6 + t];
        __syncthreads();
        if (need) {
            for (int sb = 0; sb < nsub; sb++) {
                const float d = g_box_dist(&s_sub[6 * sb], me.x, me.y, me.z);
                if (d > best[2] || d > reject) continue;
                const int lo = sb * GX_SUB, hi = min(count, lo + GX_SUB);
                for (int t = lo; t < hi; t++) g_update(me.x, me.y, me.z, s_pts[t], best, bidx);
            }
        }
        __syncthreads();
    }
    if (live) {
        const uint32_t orig = __float_as_uint(me.w);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            // fewer than 3 reference points: pytorch3d pads distances and indices with zeros (knn_points' documented
            // contract); with an empty reference cloud there is no index 0 either: -1
            const bool missing = bidx[k] == INT_MAX;
            dist2[3 * (size_t)orig + k] = missing ? 0.f : best[k];
            idx[3 * (size_t)orig + k] = missing ? (nr > 0 ? 0 : -1) : bidx[k];
        }
    }
}
// (sqrt(d2) < ratio * radius[idx]).any(-1)   (mapper.py:1376-1377)
__global__ void __launch_bounds__(256) inside_mask_kernel(int n, const float *__restrict__ dist2, const int *__restrict__ idx,
                                                          const float *__restrict__ radius, float ratio, uint8_t *mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool in = false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int j = idx[3 * (size_t)i + k];
        if (j >= 0) in = in || (fsqrt(dist2[3 * (size_t)i + k]) < fmul(radius[j], ratio));
    }
    mask[i] = in ? 1 : 0;
}

} // namespace dqo

using namespace dqo;

extern "C" int dqo_bbox_mask(int32_t n_local, const float *local_xyz, int32_t n_total, const float *total_xyz, float padding,
                             uint8_t *mask, int32_t *count, void *workspace32, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_local < 0 || n_total < 0 || (n_total > 0 && (!total_xyz || !mask)) || (n_local > 0 && !local_xyz) || !workspace32) {
        set_error("dqo_bbox_mask: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    if (count) DQO_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), stream));
    if (n_total == 0) return DQO_OK;
    uint32_t *mm = (uint32_t *)workspace32;
    geo_minmax_init_kernel<<<1, 32, 0, stream>>>(mm);
    if (n_local > 0) geo_minmax_kernel<<<min((n_local + 255) / 256, 148 * 8), 256, 0, stream>>>(n_local, local_xyz, mm);
    bbox_mask_kernel<<<(n_total + 255) / 256, 256, 0, stream>>>(n_total, total_xyz, mm, padding, mask, count);
    DQO_LAUNCH_CHECK("bbox mask", 0, stream);
    note_launch(2);
    return DQO_OK;
}

extern "C" int dqo_gaussian_radius(int32_t n, const float *log_scales, float *radius, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || (n > 0 && (!log_scales || !radius))) {
        set_error("dqo_gaussian_radius: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    if (n == 0) return DQO_OK;
    radius_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, log_scales, radius);
    DQO_LAUNCH_CHECK("gaussian radius", 0, stream);
    return DQO_OK;
}

extern "C" int dqo_scale_init(int32_t n_new, int32_t n_total, const float *xyz_total, const float *radius_total,
                              const int32_t *knn_idx, float min_radius, float max_radius, float scale_factor,
                              float xyz_factor_x, float xyz_factor_y, float xyz_factor_z, float *log_scales, uint8_t *invalid,
                              int32_t *valid_count, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_new < 0 || n_total < n_new || (n_new > 0 && (!xyz_total || !radius_total || !knn_idx || !log_scales || !invalid))) {
        set_error("dqo_scale_init: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    if (valid_count) DQO_CUDA_CHECK(cudaMemsetAsync(valid_count, 0, sizeof(int), stream));
    if (n_new == 0) return DQO_OK;
    ScaleInitArgs a;
    a.n_new = n_new; a.n_total = n_total; a.xyz = xyz_total; a.radius = radius_total; a.knn_idx = knn_idx;
    a.min_radius = min_radius; a.max_radius = max_radius; a.scale_factor = scale_factor;
    a.fx = xyz_factor_x; a.fy = xyz_factor_y; a.fz = xyz_factor_z;
    a.log_scales = log_scales; a.invalid = invalid; a.valid_count = valid_count;
    scale_init_kernel<<<(n_new + 255) / 256, 256, 0, stream>>>(a);
    DQO_LAUNCH_CHECK("scale init", 0, stream);
    return DQO_OK;
}

extern "C" size_t dqo_knn_cross3_workspace_bytes(int32_t n_query, int32_t n_ref) {
    CrossLayout L;
    make_cross_layout(n_query, n_ref, &L);
    return L.total;
}

extern "C" int dqo_knn_cross3(int32_t n_query, const float *query, int32_t n_ref, const float *ref, float *dist2, int32_t *idx,
                              void *workspace, size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_query < 0 || n_ref < 0 || (n_query > 0 && (!query || !dist2 || !idx)) || (n_ref > 0 && !ref) || !workspace) {
        set_error("dqo_knn_cross3: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    if (n_query == 0) return DQO_OK;
    CrossLayout L;
    make_cross_layout(n_query, n_ref, &L);
    if (workspace_bytes < L.total) {
        set_error("dqo_knn_cross3: workspace too small (%zu < %zu)", workspace_bytes, L.total);
        return DQO_ERR_WORKSPACE;
    }
    char *ws = (char *)workspace;
    uint32_t *mm = (uint32_t *)(ws + L.mm);
    uint32_t *r_codes = (uint32_t *)(ws + L.r_codes), *r_codes2 = (uint32_t *)(ws + L.r_codes2);
    uint32_t *r_ids = (uint32_t *)(ws + L.r_ids), *r_ids2 = (uint32_t *)(ws + L.r_ids2);
    uint32_t *q_codes = (uint32_t *)(ws + L.q_codes), *q_codes2 = (uint32_t *)(ws + L.q_codes2);
    uint32_t *q_ids = (uint32_t *)(ws + L.q_ids), *q_ids2 = (uint32_t *)(ws + L.q_ids2);
    float4 *r_sp = (float4 *)(ws + L.r_sp), *q_sp = (float4 *)(ws + L.q_sp);
    float *boxes = (float *)(ws + L.boxes), *subboxes = (float *)(ws + L.subboxes);
    const int qb = (n_query + 255) / 256, rb = (n_ref + 255) / 256;
    geo_minmax_init_kernel<<<1, 32, 0, stream>>>(mm);
    geo_minmax_kernel<<<min(qb, 148 * 8), 256, 0, stream>>>(n_query, query, mm);
    if (n_ref > 0) geo_minmax_kernel<<<min(rb, 148 * 8), 256, 0, stream>>>(n_ref, ref, mm);
    cross_morton_kernel<<<qb, 256, 0, stream>>>(n_query, query, mm, q_codes);
    if (n_ref > 0) cross_morton_kernel<<<rb, 256, 0, stream>>>(n_ref, ref, mm, r_codes);
    DQO_LAUNCH_CHECK("cross morton", 0, stream);
    note_launch(4);
    // 4 digit passes each: sorted codes end up in the (a) key buffer, sorted ids in the (a) value buffer
    int rc = radix_sort_pairs<uint32_t>(q_codes, q_codes2, q_ids, q_ids2, true, nullptr, nullptr, n_query, 30, ws + L.sort_temp,
                                        stream);
    if (rc) return rc;
    if (n_ref > 0) {
        rc = radix_sort_pairs<uint32_t>(r_codes, r_codes2, r_ids, r_ids2, true, nullptr, nullptr, n_ref, 30, ws + L.sort_temp,
                                        stream);
        if (rc) return rc;
    }
    cross_gather_kernel<<<qb, 256, 0, stream>>>(n_query, query, q_ids, q_sp);
    const int nboxes = n_ref > 0 ? (n_ref + GX_BOX - 1) / GX_BOX : 0;
    if (n_ref > 0) {
        cross_gather_kernel<<<rb, 256, 0, stream>>>(n_ref, ref, r_ids, r_sp);
        cross_boxes_kernel<<<nboxes, GX_BOX, 0, stream>>>(n_ref, r_sp, boxes, subboxes);
    }
    cross_search_kernel<<<qb, 256, 0, stream>>>(n_query, n_ref, q_sp, q_codes, r_sp, r_codes, boxes, subboxes, nboxes, dist2, idx);
    DQO_LAUNCH_CHECK("cross search", 0, stream);
    note_launch(3);
    return DQO_OK;
}

extern "C" int dqo_inside_mask(int32_t n, const float *dist2, const int32_t *idx, const float *ref_radius, float ratio,
                               uint8_t *mask, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || (n > 0 && (!dist2 || !idx || !ref_radius || !mask))) {
        set_error("dqo_inside_mask: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    if (n == 0) return DQO_OK;
    inside_mask_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, dist2, idx, ref_radius, ratio, mask);
    DQO_LAUNCH_CHECK("inside mask", 0, stream);
    return DQO_OK;
}
