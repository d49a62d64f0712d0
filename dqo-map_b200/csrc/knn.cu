// distCUDA2: Morton sort + box-culled exact 3-NN (reference KNN/simple_knn.cu:45-252).
//
// Result-identical re-design:
//   * bounding box (with the reference's {0,0,0} init, simple_knn.cu:222-231) is reduced on the device and
//     consumed on the device: no host round trips, no per-call cudaMalloc;
//   * points are gathered once into Morton order (float4 with the original index in .w) so every later
//     access is coalesced;
//   * a 256-thread block searches 256 consecutive sorted points; candidate boxes (1024 sorted points,
//     simple_knn.cu:78-117) are staged in shared memory once per block instead of being re-read through two
//     dependent global loads per candidate per thread (simple_knn.cu:202-208);
//   * inside a staged box, 32-point sub-boxes are culled with the same exact box-distance test.  Skipping a
//     (sub-)box whose distance exceeds the current third-best never changes the result because the box
//     distance is a monotone lower bound of every member's distance in the same float arithmetic.
// Candidates are visited in the reference order (box ascending, sorted position ascending) with the same
// strict '>' insertion, so ties resolve identically.
#include "common.cuh"
#include "sort.cuh"
#include <float.h>
#include <limits.h>

namespace dqo {

#define KNN_BOX 1024
#define KNN_SUB 32

struct KnnLayout {
    size_t mm;        // u32[6] order-encoded min xyz, max xyz
    size_t codes, codes_sorted, ids, ids_sorted; // u32[P]
    size_t sp;        // float4[P] sorted points (w = original index bits)
    size_t boxes;     // float[6] per 1024-box
    size_t subboxes;  // float[6] per 32-sub-box
    size_t sort_temp, total;
};

static size_t kbump(size_t &cur, size_t bytes) {
    size_t off = align_up(cur, 256);
    cur = off + bytes;
    return off;
}
static int make_knn_layout(int P, KnnLayout *L) {
    size_t cur = 0;
    size_t n = (size_t)(P > 0 ? P : 1);
    L->mm = kbump(cur, 32);
    L->codes = kbump(cur, n * 4);
    L->codes_sorted = kbump(cur, n * 4);
    L->ids = kbump(cur, n * 4);
    L->ids_sorted = kbump(cur, n * 4);
    L->sp = kbump(cur, n * 16);
    L->boxes = kbump(cur, ((n + KNN_BOX - 1) / KNN_BOX) * 24);
    L->subboxes = kbump(cur, ((n + KNN_SUB - 1) / KNN_SUB) * 24);
    SortTemp T;
    make_sort_temp((int64_t)n, 30, &T);
    L->sort_temp = kbump(cur, T.total);
    L->total = align_up(cur, 256);
    return 0;
}

__device__ __forceinline__ uint32_t enc_float(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void knn_init_kernel(uint32_t *mm) {
    if (threadIdx.x < 6) mm[threadIdx.x] = 0x80000000u; // enc(+0.0f): reduce init {0,0,0}
}

__global__ void __launch_bounds__(256) knn_minmax_kernel(int P, const float *__restrict__ pts, uint32_t *mm) {
    float mn[3] = {0.f, 0.f, 0.f}, mx[3] = {0.f, 0.f, 0.f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float v = pts[3 * i + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xFFFFFFFFu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xFFFFFFFFu, mx[c], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            atomicMin(&mm[c], enc_float(mn[c]));
            atomicMax(&mm[3 + c], enc_float(mx[c]));
        }
    }
}

__device__ __forceinline__ uint32_t prep_morton(uint32_t x) { // simple_knn.cu:45-52
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void __launch_bounds__(256)
    knn_morton_kernel(int P, const float *__restrict__ pts, const uint32_t *__restrict__ mm, uint32_t *codes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t m[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float mn = dec_float(mm[c]), mx = dec_float(mm[3 + c]);
        const float t = fmul(fdiv(fsub(pts[3 * i + c], mn), fsub(mx, mn)), 1023.0f);
        m[c] = prep_morton((uint32_t)t);
    }
    codes[i] = m[0] | (m[1] << 1) | (m[2] << 2);
}

__global__ void __launch_bounds__(256)
    knn_gather_kernel(int P, const float *__restrict__ pts, const uint32_t *__restrict__ ids_sorted, float4 *sp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t id = ids_sorted[i];
    sp[i] = make_float4(pts[3 * id], pts[3 * id + 1], pts[3 * id + 2], __uint_as_float(id));
}

// AABB of every 32 and every 1024 consecutive sorted points (simple_knn.cu:78-117)
__global__ void __launch_bounds__(1024) knn_boxes_kernel(int P, const float4 *__restrict__ sp, float *boxes, float *subboxes) {
    __shared__ float s_mn[32][3], s_mx[32][3];
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < P) {
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
        const int sb = blockIdx.x * (KNN_BOX / KNN_SUB) + warp;
        if ((size_t)sb * KNN_SUB < (size_t)P) {
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

__device__ __forceinline__ float box_dist(const float *__restrict__ b, float px, float py, float pz) { // simple_knn.cu:119-129
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (px < b[0] || px > b[3]) dx = fminf(fabsf(fsub(px, b[0])), fabsf(fsub(px, b[3])));
    if (py < b[1] || py > b[4]) dy = fminf(fabsf(fsub(py, b[1])), fabsf(fsub(py, b[4])));
    if (pz < b[2] || pz > b[5]) dz = fminf(fabsf(fsub(pz, b[2])), fabsf(fsub(pz, b[5])));
    return dot3_ref(dx, dx, dy, dy, dz, dz);
}

__device__ __forceinline__ void update_best(float px, float py, float pz, const float4 c, float best[3], int bidx[3]) {
    const float dx = fsub(c.x, px), dy = fsub(c.y, py), dz = fsub(c.z, pz); // simple_knn.cu:148-167
    float dist = dot3_ref(dx, dx, dy, dy, dz, dz);
    int id = (int)__float_as_uint(c.w);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (best[j] > dist) {
            const float t = best[j];
            best[j] = dist;
            dist = t;
            const int ti = bidx[j];
            bidx[j] = id;
            id = ti;
        }
    }
}

__global__ void __launch_bounds__(256)
    knn_search_kernel(int P, const float4 *__restrict__ sp, const float *__restrict__ boxes,
                      const float *__restrict__ subboxes, int nboxes, float *dists, int *knn_idx) {
    __shared__ float4 s_pts[KNN_BOX];
    __shared__ float s_sub[(KNN_BOX / KNN_SUB) * 6];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < P;
    float4 me = make_float4(0, 0, 0, 0);
    if (live) me = sp[idx];
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    int bidx[3] = {INT_MAX, INT_MAX, INT_MAX};
    if (live) {
        for (int i = max(0, idx - 3); i <= min(P - 1, idx + 3); i++) {
            if (i == idx) continue;
            update_best(me.x, me.y, me.z, sp[i], best, bidx);
        }
    }
    const float reject = best[2];
    best[0] = best[1] = best[2] = FLT_MAX;
    bidx[0] = bidx[1] = bidx[2] = INT_MAX;

    for (int b = 0; b < nboxes; b++) {
        bool need = false;
        if (live) {
            const float d = box_dist(boxes + 6 * (size_t)b, me.x, me.y, me.z);
            need = !(d > reject || d > best[2]);
        }
        if (!__syncthreads_or(need)) continue;
        const int base = b * KNN_BOX;
        const int count = min(KNN_BOX, P - base);
        for (int t = threadIdx.x; t < count; t += blockDim.x) s_pts[t] = sp[base + t];
        const int nsub = (count + KNN_SUB - 1) / KNN_SUB;
        for (int t = threadIdx.x; t < nsub * 6; t += blockDim.x) s_sub[t] = subboxes[(size_t)(base / KNN_SUB) * 6 + t];
        __syncthreads();
        if (need) {
            for (int sb = 0; sb < nsub; sb++) {
                const float d = box_dist(&s_sub[6 * sb], me.x, me.y, me.z);
                if (d > best[2]) continue;
                const int lo = sb * KNN_SUB, hi = min(count, lo + KNN_SUB);
                for (int t = lo; t < hi; t++) {
                    if (base + t == idx) continue;
                    update_best(me.x, me.y, me.z, s_pts[t], best, bidx);
                }
            }
        }
        __syncthreads();
    }
    if (live) {
        const uint32_t orig = __float_as_uint(me.w);
        dists[orig] = fdiv(fadd(fadd(best[0], best[1]), best[2]), 3.0f);
        knn_idx[3 * (size_t)orig] = bidx[0];
        knn_idx[3 * (size_t)orig + 1] = bidx[1];
        knn_idx[3 * (size_t)orig + 2] = bidx[2];
    }
}

} // namespace dqo

using namespace dqo;

extern "C" size_t dqo_knn_workspace_bytes(int32_t P) {
    KnnLayout L;
    if (make_knn_layout(P, &L)) return 0;
    return L.total;
}

extern "C" int dqo_knn3(int32_t P, const float *points, float *mean_dist2, int32_t *knn_idx, void *workspace,
                        size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P < 0) {
        set_error("dqo_knn3: negative P");
        return DQO_ERR_INVALID_ARG;
    }
    if (P == 0) return DQO_OK;
    if (!points || !mean_dist2 || !knn_idx || !workspace) {
        set_error("dqo_knn3: null pointer argument");
        return DQO_ERR_INVALID_ARG;
    }
    KnnLayout L;
    if (make_knn_layout(P, &L)) return DQO_ERR_WORKSPACE;
    if (workspace_bytes < L.total) {
        set_error("dqo_knn3: workspace too small (%zu < %zu)", workspace_bytes, L.total);
        return DQO_ERR_WORKSPACE;
    }
    char *ws = (char *)workspace;
    uint32_t *mm = (uint32_t *)(ws + L.mm);
    uint32_t *codes = (uint32_t *)(ws + L.codes), *codes_sorted = (uint32_t *)(ws + L.codes_sorted);
    uint32_t *ids = (uint32_t *)(ws + L.ids), *ids_sorted = (uint32_t *)(ws + L.ids_sorted);
    float4 *sp = (float4 *)(ws + L.sp);
    float *boxes = (float *)(ws + L.boxes), *subboxes = (float *)(ws + L.subboxes);
    const int nb256 = (P + 255) / 256;
    knn_init_kernel<<<1, 32, 0, stream>>>(mm);
    knn_minmax_kernel<<<min(nb256, 148 * 8), 256, 0, stream>>>(P, points, mm);
    knn_morton_kernel<<<nb256, 256, 0, stream>>>(P, points, mm, codes);
    DQO_LAUNCH_CHECK("knn morton", 0, stream);
    // stable sort of the 30-bit Morton codes (cub::DeviceRadixSort::SortPairs in the reference, simple_knn.cu:241-244; equal codes
    // keep index order either way).  4 digit passes: the sorted ids end up in the (a) value buffer = ids_sorted; the
    // values are implicit (value = index), `ids` is only the ping-pong scratch.
    {
        const int rc = radix_sort_pairs<uint32_t>(codes, codes_sorted, ids_sorted, ids, true, nullptr, nullptr, P, 30,
                                                  ws + L.sort_temp, stream);
        if (rc) return rc;
    }
    knn_gather_kernel<<<nb256, 256, 0, stream>>>(P, points, ids_sorted, sp);
    const int nboxes = (P + KNN_BOX - 1) / KNN_BOX;
    knn_boxes_kernel<<<nboxes, KNN_BOX, 0, stream>>>(P, sp, boxes, subboxes);
    knn_search_kernel<<<nb256, 256, 0, stream>>>(P, sp, boxes, subboxes, nboxes, mean_dist2, knn_idx);
    DQO_LAUNCH_CHECK("knn search", 0, stream);
    return DQO_OK;
}
