// Forward rasterization: preprocess -> depth-rank sort -> instance emission -> tile sort -> ranges -> blend.
//
// Replaces CudaRasterizer::Rasterizer::forward (reference RAST/cuda_rasterizer/rasterizer_impl.cu:205-441,
// forward.cu:238-354 and 636-866).  Design differences (results identical):
//   * no host synchronisation: R and the compact tile list stay on the device (reference syncs twice,
//     rasterizer_impl.cu:307,349-364);
//   * the 64-bit (tile|depth) sort of R instances is split into a 32-bit depth sort of the P Gaussians
//     followed by a stable tile-id sort of the instances emitted in depth-rank order.  The resulting
//     point_list is identical (stable LSD sort, ties in Gaussian-index order) at ~1/5 of the traffic;
//   * per-Gaussian splat data is packed in one 48-byte record gathered once per (tile, Gaussian);
//   * the plane / normal math of forward.cu:779-791 is evaluated only at the single "first opaque hit"
//     event per pixel (N3 in SURVEY.md);
//   * an exact-safe early reject (power < ln(1/(255 o)) - margin) skips expf for pairs that cannot pass
//     the alpha >= 1/255 test;
//   * fill values for non-rendered tiles are written by the blend kernel itself;
//   * occlusion-aware two-phase binning (front_instances > 0): only a depth-rank prefix of the Gaussians is binned for
//     every tile, the rest only for tiles that have not terminated (render_forward_kernel<1> / <2>);
//   * the latency-bound depth sort and the tile-list compaction run on a per-device high-priority side stream, forked
//     from and joined back into the caller's stream with events (everything is ordered on the caller's stream again
//     before the call returns).
#include "common.cuh"
#include "sort.cuh"
#include <math.h>
#include <stdio.h>

namespace dqo {

// ------------------------------------------------------------------------------------------------
// layouts
// ------------------------------------------------------------------------------------------------
static size_t bump(size_t &cur, size_t bytes) {
    size_t off = align_up(cur, 256);
    cur = off + bytes;
    return off;
}

int make_geom_layout(int P, GeomLayout *L) {
    size_t cur = 0;
    size_t n = (size_t)(P > 0 ? P : 1);
    L->rec = bump(cur, n * 48);
    L->depth = bump(cur, n * 4);
    L->depth_key = bump(cur, n * 4);
    L->depth_key2 = bump(cur, n * 4);
    L->ids = bump(cur, n * 4);
    L->order = bump(cur, n * 4);
    L->tiles = bump(cur, n * 4);
    L->offsets = bump(cur, n * 4);
    L->rect = bump(cur, n * 8);
    L->clamped = bump(cur, n);
    L->gacc = bump(cur, n * DQO_GACC_FLOATS * 8);
    L->touched = bump(cur, n);
    L->out_nz = bump(cur, n);
    // look-back words + tickets of the two emission kernels (front / single phase, back phase): cleared together
    L->tiles_b = bump(cur, n * 4);
    L->emit_blocks = (int)((n + 255) / 256);
    L->emit_groups = (L->emit_blocks + 63) / 64;
    // per phase (front / single, back): u32 block totals [emit_blocks] + u32 group totals [emit_groups]; the group totals
    // are accumulated with atomics, so the region is cleared at the start of every forward
    L->sums_stride = align_up((size_t)(L->emit_blocks + L->emit_groups) * 4, 256);
    L->sums = bump(cur, 2 * L->sums_stride);
    SortTemp T;
    make_sort_temp((int64_t)n, 32, &T);
    L->sort_temp = bump(cur, T.total);
    L->total = align_up(cur, 256);
    return 0;
}

int make_bin_layout(int64_t C, BinLayout *L) {
    size_t cur = 0;
    size_t n = (size_t)(C > 0 ? C : 1);
    L->keys_a = bump(cur, n * 4);
    L->keys_b = bump(cur, n * 4);
    L->vals_a = bump(cur, n * 4);
    L->vals_b = bump(cur, n * 4);
    SortTemp T;
    make_sort_temp((int64_t)n, 32, &T);
    L->sort_temp = bump(cur, T.total);
    L->total = align_up(cur, 256);
    return 0;
}

void make_img_layout(int W, int H, ImgLayout *L) {
    L->tiles_x = (W + DQO_TILE - 1) / DQO_TILE;
    L->tiles_y = (H + DQO_TILE - 1) / DQO_TILE;
    L->T = L->tiles_x * L->tiles_y;
    size_t cur = 0;
    size_t T = (size_t)(L->T > 0 ? L->T : 1);
    L->ranges = bump(cur, T * 8);
    L->n_contrib = bump(cur, T * DQO_TILE_PIX * 4);
    L->final_T = bump(cur, T * DQO_TILE_PIX * 4);
    L->hit_geo = bump(cur, T * DQO_TILE_PIX * 4 * 6);
    L->mask_words = (L->tiles_x + 31) / 32;
    L->mask_bits = bump(cur, (size_t)(L->tiles_y > 0 ? L->tiles_y : 1) * L->mask_words * 4);
    L->ranges_b = bump(cur, T * 8);
    L->unfinished = bump(cur, T * 4);
    L->mask_bits_b = bump(cur, (size_t)(L->tiles_y > 0 ? L->tiles_y : 1) * L->mask_words * 4);
    L->row_any_b = bump(cur, (size_t)((L->tiles_y + 31) / 32 + 1) * 4);
    L->state = bump(cur, T * DQO_TILE_PIX * 4 * 4);
    L->sat = bump(cur, (size_t)(L->tiles_y + 1) * (L->tiles_x + 1) * 4);
    L->sat_b = bump(cur, (size_t)(L->tiles_y + 1) * (L->tiles_x + 1) * 4);
    L->total = align_up(cur, 256);
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
struct PreArgs {
    int P, D, M, W, H;
    float color_sigma, scale_modifier;
    float tanfovx, tanfovy, focal_x, focal_y, cx, cy;
    int grid_x, grid_y;
    int prefiltered;
    int lazy_color; // two-phase binning: SH colours are evaluated by lazy_color_kernel for the Gaussians that get binned
    const float *means3D, *scales, *rotations, *opacities, *shs, *cov3D_precomp, *colors_precomp;
    const float *f_rest; // split-SH mode: shs = f_dc [P,3], f_rest [P,45]
    const float *view, *proj, *campos;
    const uint32_t *mask_bits;
    const uint32_t *sat; // summed-area table of the tile mask, (grid_y + 1) x (grid_x + 1)
    int mask_words;
    int *radii;
    int *n_touched;
    float4 *rec;
    float *depth;
    uint32_t *depth_key, *ids, *tiles;
    uint2 *rect;
    uint8_t *clamped;
    int *status;
};

__device__ __constant__ float SH_C0 = 0.28209479177387814f;
__device__ __constant__ float SH_C1 = 0.4886025119029199f;
__device__ __constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                          -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                          0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                          -0.5900435899266435f};

// SH -> RGB (forward.cu:104-155).  `sh` points at this Gaussian's [M][3] block.
template <typename SH>
__device__ __forceinline__ float3 sh_to_rgb(int deg, const SH sh, float3 pos, float3 campos, uint8_t *clamp_bits) {
    float dx = fsub(pos.x, campos.x), dy = fsub(pos.y, campos.y), dz = fsub(pos.z, campos.z);
    float len = fsqrt(dot3_ref(dx, dx, dy, dy, dz, dz));
    float x = fdiv(dx, len), y = fdiv(dy, len), z = fdiv(dz, len);
    float res[3];
#pragma unroll
    for (int c = 0; c < 3; c++) res[c] = fmul(SH_C0, sh[c]);
    if (deg > 0) {
        const float c1y = fmul(SH_C1, y), c1z = fmul(SH_C1, z), c1x = fmul(SH_C1, x);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float r = ffma(-c1y, sh[3 + c], res[c]);
            r = ffma(c1z, sh[6 + c], r);
            res[c] = ffma(-c1x, sh[9 + c], r);
        }
        if (deg > 1) {
            // operation order of the compiled reference (forward.cu:122-150, SASS of preprocessCUDA): squares and
            // products rounded separately, 2*zz as zz + zz, every coefficient = (polynomial) * constant rounded before
            // the FFMA with the SH value; in the degree-3 block 3*xx - yy, 4*zz - xx, 2*zz - 3*xx, .. - 3*yy and
            // xx - 3*yy are fused
            const float xx = fmul(x, x), yy = fmul(y, y), zz = fmul(z, z);
            const float xy = fmul(x, y), yz = fmul(y, z), xz = fmul(x, z);
            const float zz2 = fadd(zz, zz);
            const float xx_yy = fadd(xx, -yy);
            const float k0 = fmul(xy, SH_C2[0]);
            const float k1 = fmul(yz, SH_C2[1]);
            const float k2 = fmul(fadd(-yy, fadd(-xx, zz2)), SH_C2[2]);
            const float k3 = fmul(xz, SH_C2[3]);
            const float k4 = fmul(xx_yy, SH_C2[4]);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float r = ffma(k0, sh[12 + c], res[c]);
                r = ffma(k1, sh[15 + c], r);
                r = ffma(k2, sh[18 + c], r);
                r = ffma(k3, sh[21 + c], r);
                res[c] = ffma(k4, sh[24 + c], r);
            }
            if (deg > 2) {
                const float e4 = fadd(-yy, ffma(zz, 4.0f, -xx));                  // 4zz - xx - yy
                const float t0 = fmul(fmul(y, SH_C3[0]), ffma(xx, 3.0f, -yy));
                const float t1 = fmul(fmul(xy, SH_C3[1]), z);
                const float t2 = fmul(fmul(y, SH_C3[2]), e4);
                const float t3 = fmul(fmul(z, SH_C3[3]), ffma(yy, -3.0f, ffma(xx, -3.0f, zz2)));
                const float t4 = fmul(e4, fmul(x, SH_C3[4]));
                const float t5 = fmul(xx_yy, fmul(z, SH_C3[5]));
                const float t6 = fmul(fmul(x, SH_C3[6]), ffma(yy, -3.0f, xx));
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float r = ffma(t0, sh[27 + c], res[c]);
                    r = ffma(t1, sh[30 + c], r);
                    r = ffma(t2, sh[33 + c], r);
                    r = ffma(t3, sh[36 + c], r);
                    r = ffma(t4, sh[39 + c], r);
                    r = ffma(t5, sh[42 + c], r);
                    res[c] = ffma(t6, sh[45 + c], r);
                }
            }
        }
    }
    uint8_t bits = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        res[c] = fadd(res[c], 0.5f);
        if (res[c] < 0.f) bits |= (1u << c);
        res[c] = fmaxf(res[c], 0.0f);
    }
    *clamp_bits = bits;
    return make_float3(res[0], res[1], res[2]);
}

// frustum test shared with mark_visible (auxiliary.h:139-165)
__device__ __forceinline__ bool frustum_test(float px, float py, float pz, const float *__restrict__ view,
                                             const float *__restrict__ proj, float *pview_z, float *projx,
                                             float *projy) {
    float hx = xform_row(proj, 0, px, py, pz);
    float hy = xform_row(proj, 1, px, py, pz);
    float hw = xform_row(proj, 3, px, py, pz);
    float p_w = frcp(fadd(hw, 0.0000001f));
    float ppx = fmul(hx, p_w), ppy = fmul(hy, p_w);
    float vz = xform_row(view, 2, px, py, pz);
    *pview_z = vz;
    *projx = ppx;
    *projy = ppy;
    // the reference compares against the double literals -1.3 / 1.3
    if (vz <= 0.2f || (double)ppx < -1.3 || (double)ppx > 1.3 || (double)ppy < -1.3 || (double)ppy > 1.3) return false;
    return true;
}

// Depth-sort keys from the positions alone: float bits of the view depth, 0xFFFFFFFF for Gaussians outside the frustum.
// (Gaussians that pass the frustum test but emit no instance keep their depth key: they carry zero tiles, so the
// relative order of the emitting ones -- all that matters -- is unchanged.)  Being independent of the rest of the
// preprocess, the depth sort runs on a side stream concurrently with it.
__global__ void __launch_bounds__(256) depth_key_kernel(int P, const float *__restrict__ means, const float *__restrict__ view,
                                                        const float *__restrict__ proj, uint32_t *depth_key) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float vz, ppx, ppy;
    const bool ok = frustum_test(means[3 * idx], means[3 * idx + 1], means[3 * idx + 2], view, proj, &vz, &ppx, &ppy);
    depth_key[idx] = ok ? __float_as_uint(vz) : 0xFFFFFFFFu;
}

// Everything the forward pass needs zeroed (tile ranges, status words, prefix-sum totals, the scratch headers of its three
// sorts) in ONE launch at its start instead of up to seven memset nodes spread over the chain: a memset between two
// kernels is a full serialisation point (no programmatic overlap) and one more node to dispatch.
#define DQO_CLEAR_MAX 8
struct ClearArgs {
    void *ptr[DQO_CLEAR_MAX];
    unsigned long long bytes[DQO_CLEAR_MAX]; // multiples of 4, pointers 4-byte aligned
    int n;
};
__global__ void __launch_bounds__(256) clear_regions_kernel(ClearArgs a) {
    pdl_enter();
    for (int r = 0; r < a.n; r++) {
        uint32_t *p = (uint32_t *)a.ptr[r];
        const unsigned long long words = a.bytes[r] / 4;
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < words;
             i += (unsigned long long)gridDim.x * blockDim.x)
            p[i] = 0u;
    }
}

// tile_mask != 0 as one bitmap row per tile row: the per-Gaussian tile count and the instance emission then cost
// O(rows x words) instead of one global load per tile of the rectangle
__global__ void mask_bits_kernel(int gx, int gy, int words, const int *__restrict__ tile_mask, uint32_t *bits) {
    pdl_enter();
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= gy * words) return;
    const int y = w / words, word = w % words, x = word * 32 + (threadIdx.x & 31);
    const bool on = (x < gx) && (tile_mask[y * gx + x] != 0);
    const unsigned b = __ballot_sync(0xFFFFFFFFu, on);
    if ((threadIdx.x & 31) == 0) bits[w] = b;
}

// number of set mask bits in columns [minx, maxx) of tile row y
__device__ __forceinline__ uint32_t mask_row_count(const uint32_t *__restrict__ bits, int words, uint32_t y, uint32_t minx,
                                                   uint32_t maxx) {
    uint32_t c = 0;
    for (uint32_t w = minx >> 5; w <= (maxx - 1) >> 5; w++) {
        uint32_t m = __ldg(&bits[y * words + w]);
        const uint32_t lo = (w == (minx >> 5)) ? (minx & 31) : 0;
        const uint32_t hi = (w == ((maxx - 1) >> 5)) ? ((maxx - 1) & 31) : 31;
        m &= (0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (31 - hi));
        c += __popc(m);
    }
    return c;
}

// Summed-area table of a tile bitmap: sat[y * (gx + 1) + x] = number of set bits in rows [0, y) x columns [0, x).
// The number of masked tiles in a rectangle is then four loads, whatever its size (the row-by-row popcount it replaces
// was a chain of dependent loads per Gaussian: 15 % of the preprocess).  Built by ONE block: row prefixes from the
// bitmap words, then a running sum down every column.
#define SAT_SMEM_ENTRIES 9216 // 36 KB of shared memory: images up to 1920x1080 (121 x 68 entries); larger ones take the slow path
__device__ __forceinline__ void build_sat(int gx, int gy, int words, const uint32_t *bits, uint32_t *sat) {
    __shared__ uint32_t s_pre[SAT_SMEM_ENTRIES];
    const int stride = gx + 1;
    if (gy * stride <= SAT_SMEM_ENTRIES) {
        // all (row, column) prefixes in parallel from the bitmap words, then one thread per column sums down the rows
        for (int e = threadIdx.x; e < gy * stride; e += blockDim.x) {
            const int y = e / stride, x = e - y * stride; // bits of row y in columns [0, x)
            const uint32_t *row = bits + y * words;
            uint32_t c = 0;
            for (int w = 0; w < (x >> 5); w++) c += __popc(row[w]);
            if (x & 31) c += __popc(row[x >> 5] & ((1u << (x & 31)) - 1u));
            s_pre[e] = c;
        }
        __syncthreads();
        for (int x = threadIdx.x; x < stride; x += blockDim.x) {
            uint32_t run = 0;
            sat[x] = 0;
            for (int y = 0; y < gy; y++) {
                run += s_pre[y * stride + x];
                sat[(y + 1) * stride + x] = run;
            }
        }
        return;
    }
    // one thread per column x: running sum over the rows of "set bits of row y in columns [0, x)", read straight from
    // the bitmap words (a few hundred bytes, cache-resident); the table is only ever written, never read back
    for (int x = threadIdx.x; x < stride; x += blockDim.x) {
        const int full = x >> 5;
        const uint32_t part = (x & 31) ? ((1u << (x & 31)) - 1u) : 0u;
        uint32_t run = 0;
        sat[x] = 0;
        for (int y = 0; y < gy; y++) {
            const uint32_t *row = bits + y * words;
            uint32_t c = 0;
            for (int w = 0; w < full; w++) c += __popc(row[w]);
            if (part) c += __popc(row[full] & part);
            run += c;
            sat[(y + 1) * stride + x] = run;
        }
    }
}
__device__ __forceinline__ uint32_t sat_count(const uint32_t *__restrict__ sat, int gx, uint32_t minx, uint32_t maxx,
                                              uint32_t miny, uint32_t maxy) {
    const uint32_t stride = (uint32_t)gx + 1u;
    return __ldg(&sat[maxy * stride + maxx]) - __ldg(&sat[miny * stride + maxx]) - __ldg(&sat[maxy * stride + minx]) +
           __ldg(&sat[miny * stride + minx]);
}
__global__ void __launch_bounds__(1024) mask_sat_kernel(int gx, int gy, int words, const uint32_t *bits, uint32_t *sat) {
    pdl_enter();
    build_sat(gx, gy, words, bits, sat);
}

struct ShRegs { // 48 SH floats of one Gaussian held in registers (flat [k][c] order)
    float v[48];
    __device__ __forceinline__ float operator[](int i) const { return v[i]; }
};
struct ShPtr {
    const float *p;
    __device__ __forceinline__ float operator[](int i) const { return p[i]; }
};

#define PRE_THREADS 128
#define SH_ROW_Q 13 // float4 per staged SH row (12 used + 1 pad: conflict-free for both access patterns)

// SHMODE 1 (M == 16, 16-byte aligned SH): each warp pulls its 32 x 192 B of coefficients with coalesced 128-bit loads
// through shared memory instead of 32 strided streams.  SHMODE 2: the same for the reference's split parameter
// tensors f_dc [P,1,3] / f_rest [P,15,3] (GaussianPointCloud.get_features' torch.cat is never materialised); the
// 45-float rows of f_rest are conflict-free for per-thread scalar reads as they lie.  SHMODE 0: plain pointer.
template <int SHMODE>
__global__ void __launch_bounds__(PRE_THREADS) preprocess_kernel(PreArgs a) {
    pdl_enter();
    extern __shared__ float4 s_sh[];
    constexpr bool STAGED = SHMODE == 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *wbuf = s_sh + warp * 32 * SH_ROW_Q;
    float *wrest = reinterpret_cast<float *>(wbuf), *wdc = wrest + 1440;
    if (SHMODE == 2 && !a.lazy_color) {
        const int base_g = blockIdx.x * blockDim.x + warp * 32;
        const int nrow = min(32, a.P - base_g);
        if (nrow == 32) {
            const float4 *g = reinterpret_cast<const float4 *>(a.f_rest + (size_t)base_g * 45);
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const int q = i * 32 + lane;
                if (q < 360) wbuf[q] = __ldg(g + q);
            }
            const float4 *gd = reinterpret_cast<const float4 *>(a.shs + (size_t)base_g * 3);
            if (lane < 24) reinterpret_cast<float4 *>(wdc)[lane] = __ldg(gd + lane);
        } else if (nrow > 0) {
            for (int q = lane; q < nrow * 45; q += 32) wrest[q] = a.f_rest[(size_t)base_g * 45 + q];
            for (int q = lane; q < nrow * 3; q += 32) wdc[q] = a.shs[(size_t)base_g * 3 + q];
        }
        __syncwarp();
    }
    if (STAGED && !a.lazy_color) {
        const int base_g = blockIdx.x * blockDim.x + warp * 32;
        const int nrow = min(32, a.P - base_g);
        if (nrow > 0) {
            const float4 *g = reinterpret_cast<const float4 *>(a.shs) + (size_t)base_g * 12;
            const int nq = nrow * 12;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const int q = i * 32 + lane;
                if (q < nq) wbuf[(q / 12) * SH_ROW_Q + (q % 12)] = __ldg(g + q);
            }
        }
        __syncwarp();
    }
    if (idx >= a.P) return;
    int radius_out = 0;
    uint32_t tiles_out = 0;
    uint8_t flags_out = 0; // bits 0-2: clamped channels, bit 7: splat record valid
    if (a.n_touched) a.n_touched[idx] = 0;

    const float px = a.means3D[3 * idx], py = a.means3D[3 * idx + 1], pz = a.means3D[3 * idx + 2];
    float vz, ppx, ppy;
    bool ok = frustum_test(px, py, pz, a.view, a.proj, &vz, &ppx, &ppy);
    if (!ok && a.prefiltered) {
        printf("Point is filtered although prefiltered is set. This shouldn't happen!");
        __trap();
    }
    if (ok) {
        float cov3[6];
        if (a.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) cov3[k] = a.cov3D_precomp[6 * idx + k];
        } else {
            const float4 q = reinterpret_cast<const float4 *>(a.rotations)[idx];
            cov3d_from_scale_rot(a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2], a.scale_modifier, q.x,
                                 q.y, q.z, q.w, cov3);
        }
        // EWA 2D covariance (forward.cu:158-197)
        const float *v = a.view;
        float tx = xform_row(v, 0, px, py, pz), ty = xform_row(v, 1, px, py, pz);
        const float tz = vz;
        const float limx = fmul(1.3f, a.tanfovx), limy = fmul(1.3f, a.tanfovy);
        const float txtz = fdiv(tx, tz), tytz = fdiv(ty, tz);
        tx = fmul(fminf(limx, fmaxf(-limx, txtz)), tz);
        ty = fmul(fminf(limy, fmaxf(-limy, tytz)), tz);
        const float tz2 = fmul(tz, tz);
        const float J00 = fdiv(a.focal_x, tz), J11 = fdiv(a.focal_y, tz);
        const float J02 = fdiv(fmul(-tx, a.focal_x), tz2), J12 = fdiv(fmul(-ty, a.focal_y), tz2);
        // T = W * J ; W[0][r] = v[4r], W[1][r] = v[4r+1], W[2][r] = v[4r+2]
        float T0[3], T1[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            T0[r] = ffma(J02, v[4 * r + 2], fmul(v[4 * r], J00));
            T1[r] = ffma(J12, v[4 * r + 2], fmul(v[4 * r + 1], J11));
        }
        const float V0[3] = {cov3[0], cov3[1], cov3[2]}, V1[3] = {cov3[1], cov3[3], cov3[4]},
                    V2[3] = {cov3[2], cov3[4], cov3[5]};
        // A[c][r] = T[r] . V[c]
        const float A00 = dot3_ref(T0[0], V0[0], T0[1], V0[1], T0[2], V0[2]);
        const float A01 = dot3_ref(T1[0], V0[0], T1[1], V0[1], T1[2], V0[2]);
        const float A10 = dot3_ref(T0[0], V1[0], T0[1], V1[1], T0[2], V1[2]);
        const float A11 = dot3_ref(T1[0], V1[0], T1[1], V1[1], T1[2], V1[2]);
        const float A20 = dot3_ref(T0[0], V2[0], T0[1], V2[1], T0[2], V2[2]);
        const float A21 = dot3_ref(T1[0], V2[0], T1[1], V2[1], T1[2], V2[2]);
        // cov[c][r] = A[0][r]*T[c][0] + A[1][r]*T[c][1] + A[2][r]*T[c][2]
        const float cov_x = fadd(dot3_ref(T0[0], A00, T0[1], A10, T0[2], A20), 0.3f);
        const float cov_y = dot3_ref(T0[0], A01, T0[1], A11, T0[2], A21);
        const float cov_z = fadd(dot3_ref(T1[0], A01, T1[1], A11, T1[2], A21), 0.3f);

        const float det = ffma(cov_x, cov_z, -fmul(cov_y, cov_y));
        if (det != 0.0f) {
            const float det_inv = frcp(det);
            const float conic_x = fmul(cov_z, det_inv), conic_y = fmul(cov_y, -det_inv), conic_z = fmul(cov_x, det_inv);
            const float mid = fmul(fadd(cov_x, cov_z), 0.5f);
            const float sq = fsqrt(fmaxf(0.1f, ffma(mid, mid, -det)));
            const float lam = fmaxf(fadd(mid, sq), fsub(mid, sq));
            const float rad_f = ceilf(fmul(a.color_sigma, fsqrt(lam)));
            const int my_radius = (int)rad_f;
            // ndc2Pix(v, S, c) = v * S * 0.5 + c evaluated in double (auxiliary.h:44-47)
            const float pix_x = (float)fma((double)fmul(ppx, (float)a.W), 0.5, (double)a.cx);
            const float pix_y = (float)fma((double)fmul(ppy, (float)a.H), 0.5, (double)a.cy);
            // getRect (auxiliary.h:49-57)
            const float rf = (float)my_radius;
            int rx0 = (int)fmul(fsub(pix_x, rf), 0.0625f), ry0 = (int)fmul(fsub(pix_y, rf), 0.0625f);
            int rx1 = (int)fmul(fadd(fadd(fadd(pix_x, rf), 16.0f), -1.0f), 0.0625f);
            int ry1 = (int)fmul(fadd(fadd(fadd(pix_y, rf), 16.0f), -1.0f), 0.0625f);
            const uint32_t minx = min((uint32_t)a.grid_x, (uint32_t)max(0, rx0));
            const uint32_t miny = min((uint32_t)a.grid_y, (uint32_t)max(0, ry0));
            const uint32_t maxx = min((uint32_t)a.grid_x, (uint32_t)max(0, rx1));
            const uint32_t maxy = min((uint32_t)a.grid_y, (uint32_t)max(0, ry1));
            if ((maxx - minx) * (maxy - miny) != 0) {
                float3 rgb = make_float3(0.f, 0.f, 0.f);
                uint8_t cl = 0;
                if (a.lazy_color) {
                    // filled in by lazy_color_kernel if this Gaussian is binned
                } else if (a.colors_precomp == nullptr) {
                    const float3 cam = make_float3(a.campos[0], a.campos[1], a.campos[2]);
                    if (STAGED) {
                        ShRegs sh;
#pragma unroll
                        for (int i = 0; i < 12; i++) {
                            const float4 t = wbuf[lane * SH_ROW_Q + i];
                            sh.v[4 * i] = t.x; sh.v[4 * i + 1] = t.y; sh.v[4 * i + 2] = t.z; sh.v[4 * i + 3] = t.w;
                        }
                        rgb = sh_to_rgb(a.D, sh, make_float3(px, py, pz), cam, &cl);
                    } else if (SHMODE == 2) {
                        ShRegs sh;
#pragma unroll
                        for (int c = 0; c < 3; c++) sh.v[c] = wdc[lane * 3 + c];
#pragma unroll
                        for (int k = 0; k < 45; k++) sh.v[3 + k] = wrest[lane * 45 + k];
                        rgb = sh_to_rgb(a.D, sh, make_float3(px, py, pz), cam, &cl);
                    } else {
                        rgb = sh_to_rgb(a.D, ShPtr{a.shs + (size_t)idx * a.M * 3}, make_float3(px, py, pz), cam, &cl);
                    }
                } else {
                    rgb = make_float3(a.colors_precomp[3 * idx], a.colors_precomp[3 * idx + 1],
                                      a.colors_precomp[3 * idx + 2]);
                }
                const float opacity = a.opacities[idx];
                // conservative reject bound: alpha = o*exp(power) < 1/255 whenever power < ln(1/(255 o)) - margin
                float thr = logf(1.0f / (255.0f * opacity));
                thr = thr - (1e-4f + 1e-5f * fabsf(thr));
                if (!(thr == thr)) thr = -INFINITY; // NaN (negative / NaN opacity): never early-reject
                // conservative extent of {pixels whose float-evaluated power can reach thr}: bounding box of the ellipse
                // d^T Q d <= tau' of the conic actually used, tau' inflated for the rounding noise of the fp32 power
                // (noise <= delta * (A dx^2 + C dy^2) <= delta * kappa * d^T Q d, kappa = 1 / (1 - |rho|))
                float ex = INFINITY, ey = INFINITY;
                {
                    const double A = conic_x, B = conic_y, C = conic_z;
                    const double detq = A * C - B * B;
                    const double rho = fabs(B) / sqrt(A * C);
                    const double shrink = 1.0 - 2.0 * 2e-6 / (1.0 - rho);
                    if (detq > 0.0 && A > 0.0 && C > 0.0 && rho < 1.0 && shrink > 0.5) {
                        const double tau = -2.0 * (double)thr / shrink;
                        if (tau < 0.0) {
                            ex = ey = -1.0f; // no pixel can pass the alpha test (opacity < 1/255)
                        } else {
                            ex = (float)(sqrt(tau * C / detq) * 1.001 + 0.05);
                            ey = (float)(sqrt(tau * A / detq) * 1.001 + 0.05);
                        }
                    }
                    if (!(ex == ex) || !(ey == ey)) ex = ey = INFINITY;
                }
                a.rec[3 * idx + 0] = make_float4(pix_x, pix_y, conic_x, conic_y);
                a.rec[3 * idx + 1] = make_float4(conic_z, opacity, thr, ey);
                a.rec[3 * idx + 2] = make_float4(rgb.x, rgb.y, rgb.z, ex);
                a.depth[idx] = vz;
                a.rect[idx] = make_uint2(minx | (maxx << 16), miny | (maxy << 16));
                flags_out = cl | 0x80;
                radius_out = my_radius;
                tiles_out = sat_count(a.sat, a.grid_x, minx, maxx, miny, maxy);
            }
        }
    }
    a.radii[idx] = radius_out;
    a.tiles[idx] = tiles_out;
    a.clamped[idx] = flags_out;
    const unsigned vis = __ballot_sync(__activemask(), radius_out > 0);
    if ((threadIdx.x & 31) == 0 && vis) atomicAdd(&a.status[DQO_ST_NUM_VISIBLE], __popc(vis));
}

// SH colours on demand (two-phase binning).  With occlusion-aware binning only the nearest ~15-20 % of the Gaussians of a
// large map are ever binned, so evaluating computeColorFromSH (forward.cu:106-163) for all of them in the preprocess
// reads 192 B and spends ~350 instructions per Gaussian on colours nobody looks at.  The preprocess then leaves the rgb
// of the splat record empty and this kernel fills it in -- same function, same inputs, bit-identical values -- for
// exactly the ranks a binning phase emitted:
//   phase 1  ranks whose inclusive instance offset fits the front region (what emit_kernel<1> wrote);
//   phase 2  ranks with unfinished tiles in their rectangle (tiles_rank of rank_sums_kernel<2>); disjoint from phase 1.
// It runs on the side stream beside the tile sort of its phase; only the blend of that phase waits for it.
// One warp per 32 consecutive ranks; the needed SH rows (scattered: rank order is not memory order) are fetched
// row by row with coalesced loads through shared memory.
struct ColorArgs {
    int P, D, phase;
    int64_t front;
    const uint32_t *order, *offsets, *tiles, *tiles_rank; // tiles: tiles_touched in RANK order (rank_sums_kernel<0>)
    const float *means3D, *shs, *f_rest, *campos;
    float4 *rec;
    uint8_t *clamped;
};
#define LC_THREADS 128
template <int SHMODE> // 1: merged SH [P,16,3], 2: split f_dc [P,3] / f_rest [P,45]
__global__ void __launch_bounds__(LC_THREADS) lazy_color_kernel(ColorArgs a) {
    pdl_enter();
    __shared__ float s_rows[LC_THREADS / 32][32 * 49]; // 48 floats per row + 1 pad
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * LC_THREADS + threadIdx.x;
    bool need = false;
    uint32_t id = 0;
    if (r < a.P) {
        id = a.order[r];
        need = (a.phase == 1) ? (a.tiles[r] != 0 && (int64_t)a.offsets[r] <= a.front) : (a.tiles_rank[r] != 0);
    }
    const unsigned rows = __ballot_sync(0xFFFFFFFFu, need);
    if (!rows) return;
    float *wrow = s_rows[warp];
    // four rows per trip: all their loads are issued before the first shared-memory store waits for one of them
    for (unsigned m = rows; m;) {
        int k[4];
        uint32_t gid[4];
        float v0[4], v1[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            k[q] = m ? (__ffs(m) - 1) : -1;
            m &= m - 1;
            gid[q] = __shfl_sync(0xFFFFFFFFu, id, k[q] < 0 ? 0 : k[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            v0[q] = v1[q] = 0.f;
            if (k[q] < 0) continue;
            if (SHMODE == 2) {
                const float *src = a.f_rest + (size_t)gid[q] * 45;
                v0[q] = __ldg(src + lane);
                if (lane < 13) v1[q] = __ldg(src + 32 + lane);
                else if (lane < 16) v1[q] = __ldg(a.shs + (size_t)gid[q] * 3 + (lane - 13));
            } else {
                const float *src = a.shs + (size_t)gid[q] * 48;
                v0[q] = __ldg(src + lane);
                if (lane < 16) v1[q] = __ldg(src + 32 + lane);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (k[q] < 0) continue;
            if (SHMODE == 2) {
                wrow[k[q] * 49 + 3 + lane] = v0[q];
                if (lane < 13) wrow[k[q] * 49 + 3 + 32 + lane] = v1[q];
                else if (lane < 16) wrow[k[q] * 49 + (lane - 13)] = v1[q];
            } else {
                wrow[k[q] * 49 + lane] = v0[q];
                if (lane < 16) wrow[k[q] * 49 + 32 + lane] = v1[q];
            }
        }
    }
    __syncwarp();
    if (!need) return;
    ShRegs sh;
#pragma unroll
    for (int k = 0; k < 48; k++) sh.v[k] = wrow[lane * 49 + k];
    const float3 pos = make_float3(a.means3D[3 * (size_t)id], a.means3D[3 * (size_t)id + 1], a.means3D[3 * (size_t)id + 2]);
    uint8_t cl = 0;
    const float3 rgb = sh_to_rgb(a.D, sh, pos, make_float3(a.campos[0], a.campos[1], a.campos[2]), &cl);
    float *dst = reinterpret_cast<float *>(a.rec + 3 * (size_t)id + 2);
    dst[0] = rgb.x;
    dst[1] = rgb.y;
    dst[2] = rgb.z;
    a.clamped[id] = cl | 0x80;
}

__global__ void mark_visible_kernel(int P, const float *__restrict__ means, const float *__restrict__ view,
                                    const float *__restrict__ proj, uint8_t *present) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float vz, ppx, ppy;
    present[idx] = frustum_test(means[3 * idx], means[3 * idx + 1], means[3 * idx + 2], view, proj, &vz, &ppx, &ppy) ? 1 : 0;
}

// Number of instances rank i emits: tiles_touched of its Gaussian (MODE 0 / 1), or the number of UNFINISHED tiles in its
// rectangle when the front phase did not bin it (MODE 2: offsets[rank] > front).
template <int MODE>
__device__ __forceinline__ uint32_t rank_count(int64_t i, uint32_t id, int64_t front, const uint32_t *__restrict__ tiles,
                                               const uint32_t *offsets, const uint2 *__restrict__ rect,
                                               const uint32_t *__restrict__ sat_b, int grid_x,
                                               const uint32_t *__restrict__ rank_tiles) {
    if (MODE != 2) return tiles[id];
    uint32_t n = 0;
    if ((int64_t)offsets[i] > front && rank_tiles[i]) { // (tiles_touched in rank order: no second gather)
        const uint2 rc = rect[id];
        const uint32_t minx = rc.x & 0xFFFF, maxx = rc.x >> 16, miny = rc.y & 0xFFFF, maxy = rc.y >> 16;
        n = sat_count(sat_b, grid_x, minx, maxx, miny, maxy);
    }
    return n;
}

// Step 1 of the prefix sum of the per-rank instance counts (rasterizer_impl.cu:303, in depth-rank order): the total of
// every block of 256 ranks, and the totals of groups of 64 blocks.  The emission kernel then finds its exclusive prefix
// by adding at most P/16384 group totals and 63 block totals -- no chain of blocks waiting for each other (a single-pass
// look-back scan had every one of the ~4000 simultaneously resident blocks wait for its predecessors: 3x slower).
template <int MODE>
__global__ void __launch_bounds__(256)
    rank_sums_kernel(int P, int64_t front, const uint32_t *__restrict__ order, const uint32_t *__restrict__ tiles,
                     const uint32_t *offsets, const uint2 *__restrict__ rect, const uint32_t *__restrict__ sat_b, int grid_x,
                     uint32_t *tiles_rank, uint32_t *rank_tiles, uint32_t *sums, uint32_t *group_sums) {
    pdl_enter();
    __shared__ uint32_t s_w[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i = (int64_t)blockIdx.x * 256 + tid;
    uint32_t n = 0;
    if (i < P) {
        n = rank_count<MODE>(i, order[i], front, tiles, offsets, rect, sat_b, grid_x, rank_tiles);
        if (MODE == 2) tiles_rank[i] = n;
        else rank_tiles[i] = n; // tiles_touched in depth-rank order: the emission and the back phase read it coalesced
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
    if (lane == 0) s_w[warp] = n;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += s_w[w];
        sums[blockIdx.x] = t;
        if (t) atomicAdd(&group_sums[blockIdx.x >> 6], t);
    }
}

// Step 2: emission of one (tile, gaussian) pair per masked tile of the rectangle (duplicateWithKeys,
// rasterizer_impl.cu:70-115), walking the Gaussians in depth-rank order so that a stable sort by tile id alone
// reproduces the reference order.  A block owns 256 consecutive ranks: exclusive prefix from the block / group totals,
// block-wide inclusive scan of the counts, then one warp serves 32 consecutive ranks -- the run of each Gaussian is
// written by all lanes together (coalesced) when no tile of its rectangle is masked out, otherwise by its owner lane
// walking the mask bitmap.  R never visits the host and nothing is padded: the sort that follows reads its count from
// `status`.
//   MODE 0 (single phase): every rank; writes beyond `capacity` are dropped and DQO_ST_OVERFLOW is raised by the last rank.
//   MODE 1 (front phase):  only the ranks whose inclusive offset fits into `capacity` (= front_instances), i.e. the
//          nearest Gaussians; R_front is reported.  `offsets` still receives the full scan (R, and the back phase's
//          "not binned yet" test).
//   MODE 2 (back phase):   the count of a rank is the number of unfinished tiles in its rectangle (tiles_rank, written by
//          rank_sums_kernel<2>); overflow when R_back > capacity.
template <typename KeyT, int MODE>
__global__ void __launch_bounds__(256)
    emit_kernel(int P, int64_t capacity, const uint32_t *__restrict__ order, const uint32_t *__restrict__ tiles,
                const uint32_t *__restrict__ tiles_rank, const uint32_t *__restrict__ rank_tiles, uint32_t *offsets,
                const uint2 *__restrict__ rect,
                const uint32_t *__restrict__ mask_bits, const uint32_t *__restrict__ row_any, int mask_words, int grid_x,
                KeyT *__restrict__ keys, uint32_t *__restrict__ vals, const uint32_t *__restrict__ sums,
                const uint32_t *__restrict__ group_sums, int *status) {
    pdl_enter();
    __shared__ uint32_t s_n[257];
    __shared__ uint32_t s_warp_tot[8], s_warp_pre[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int64_t i = (int64_t)tile * 256 + tid;
    // exclusive prefix of this block: whole groups of 64 blocks, then the blocks of its own group
    uint32_t pre = 0;
    for (int j = tid; j < (tile >> 6); j += 256) pre += group_sums[j];
    {
        const int j = ((tile >> 6) << 6) + tid;
        if (tid < 64 && j < tile) pre += sums[j];
    }
    uint32_t id = 0, n = 0;
    uint2 rc = make_uint2(0, 0);
    if (i < P) {
        id = order[i];
        n = (MODE == 2) ? tiles_rank[i] : rank_tiles[i];
        if (n) rc = rect[id];
    }
    s_n[tid] = n;
    if (MODE == 1 && tid == 255) s_n[256] = (i + 1 < P) ? rank_tiles[i + 1] : 0u;
    // block-wide inclusive scan of n, block-wide sum of pre
    uint32_t incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xFFFFFFFFu, pre, o);
    if (lane == 31) s_warp_tot[warp] = incl;
    if (lane == 0) s_warp_pre[warp] = pre;
    __syncthreads();
    uint32_t woff = 0, block_excl = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        if (w < warp) woff += s_warp_tot[w];
        block_excl += s_warp_pre[w];
    }
    const uint32_t s_excl = block_excl;
    const uint32_t end = s_excl + woff + incl; // inclusive offset of rank i
    uint32_t off = end - n;
    if (i < P) {
        if (MODE != 2) offsets[i] = end;
        if (MODE == 0) {
            if (i == P - 1) {
                status[DQO_ST_NUM_RENDERED] = (int)end;
                status[DQO_ST_OVERFLOW] = ((int64_t)end > capacity) ? 1 : 0;
            }
        } else if (MODE == 1) {
            if (i == P - 1) status[DQO_ST_NUM_RENDERED] = (int)end;
            if ((int64_t)end > capacity)
                n = 0;
            else if (i == P - 1 || (int64_t)end + (int64_t)s_n[tid + 1] > capacity)
                status[DQO_ST_R_FRONT] = (int)end;
        } else {
            if (i == P - 1) {
                status[DQO_ST_R_BACK] = (int)end;
                if ((int64_t)end > capacity) status[DQO_ST_OVERFLOW] = 1;
            }
        }
    }
    unsigned todo = __ballot_sync(0xFFFFFFFFu, n != 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t s_id = __shfl_sync(0xFFFFFFFFu, id, src), sn = __shfl_sync(0xFFFFFFFFu, n, src);
        const uint32_t s_off = __shfl_sync(0xFFFFFFFFu, off, src);
        const uint32_t rx = __shfl_sync(0xFFFFFFFFu, rc.x, src), ry = __shfl_sync(0xFFFFFFFFu, rc.y, src);
        const uint32_t minx = rx & 0xFFFF, maxx = rx >> 16, miny = ry & 0xFFFF, maxy = ry >> 16;
        const uint32_t w = maxx - minx;
        if (sn == w * (maxy - miny)) { // nothing masked inside the rectangle
            for (uint32_t k = lane; k < sn; k += 32) {
                const uint32_t dy = k / w, dx = k - dy * w;
                if ((int64_t)s_off + k < capacity) {
                    keys[s_off + k] = (KeyT)((miny + dy) * grid_x + minx + dx);
                    vals[s_off + k] = s_id;
                }
            }
        } else if (lane == src) {
            uint32_t o = s_off;
            for (uint32_t y = miny; y < maxy; y++) {
                if (MODE == 2 && !((__ldg(&row_any[y >> 5]) >> (y & 31)) & 1)) continue; // no unfinished tile in this row
                for (uint32_t wd = minx >> 5; wd <= (maxx - 1) >> 5; wd++) {
                    uint32_t m = __ldg(&mask_bits[y * mask_words + wd]);
                    const uint32_t lo = (wd == (minx >> 5)) ? (minx & 31) : 0;
                    const uint32_t hi = (wd == ((maxx - 1) >> 5)) ? ((maxx - 1) & 31) : 31;
                    m &= (0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (31 - hi));
                    while (m) {
                        const uint32_t x = wd * 32 + (__ffs(m) - 1);
                        m &= m - 1;
                        if ((int64_t)o < capacity) {
                            keys[o] = (KeyT)(y * grid_x + x);
                            vals[o] = s_id;
                        }
                        o++;
                    }
                }
            }
        }
    }
}

// Back phase, step 1 (one block): bitmap of the tiles that are both masked in and unfinished after the front phase,
// plus one summary bit per tile row so that the per-Gaussian passes below skip rows (and, mostly, whole Gaussians)
// without touching the bitmap.
__global__ void __launch_bounds__(1024)
    mask_unfinished_kernel(int tiles_x, int tiles_y, int mask_words, const uint32_t *__restrict__ mask_bits,
                           const int *__restrict__ unfinished, uint32_t *mask_bits_b, uint32_t *row_any, uint32_t *sat_b) {
    pdl_enter();
    __shared__ uint32_t s_any[64];
    const int row_words = (tiles_y + 31) / 32;
    for (int k = threadIdx.x; k < 64; k += blockDim.x) s_any[k] = 0;
    __syncthreads();
    // one warp per bitmap word: 32 flags -> one ballot
    const int lane = threadIdx.x & 31;
    for (int w = threadIdx.x >> 5; w < tiles_y * mask_words; w += blockDim.x >> 5) {
        const int y = w / mask_words, x = (w % mask_words) * 32 + lane;
        const bool u = x < tiles_x && unfinished[y * tiles_x + x] != 0;
        const uint32_t bits = __ballot_sync(0xFFFFFFFFu, u) & mask_bits[w];
        if (lane == 0) {
            mask_bits_b[w] = bits;
            if (bits && (y >> 5) < 64) atomicOr(&s_any[y >> 5], 1u << (y & 31));
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < row_words; k += blockDim.x) row_any[k] = (k < 64) ? s_any[k] : 0xFFFFFFFFu;
    __syncthreads(); // mask_bits_b was written by this block
    build_sat(tiles_x, tiles_y, mask_words, mask_bits_b, sat_b);
}

// per-tile [start, end) in the sorted list (rasterizer_impl.cu:120-142); each thread checks 8 consecutive keys
template <typename KeyT>
__global__ void __launch_bounds__(256)
    tile_ranges_kernel(int64_t capacity, const KeyT *__restrict__ keys, const int *__restrict__ status, int count_word,
                       uint2 *ranges) {
    pdl_enter();
    const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const int64_t L = status[DQO_ST_OVERFLOW] ? 0 : status[count_word];
    if (base >= L) return;
    uint32_t prev = (base > 0) ? (uint32_t)keys[base - 1] : 0xFFFFFFFFu;
    KeyT k[8];
    if (base + 8 <= L && sizeof(KeyT) == 2) {
        const uint4 v = *reinterpret_cast<const uint4 *>(keys + base); // capacity-sized arrays are 256-byte aligned
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            k[2 * q] = (KeyT)(w[q] & 0xFFFF);
            k[2 * q + 1] = (KeyT)(w[q] >> 16);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 8; q++) k[q] = (base + q < L) ? keys[base + q] : (KeyT)0;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int64_t i = base + q;
        if (i >= L) break;
        const uint32_t cur = k[q];
        if (i == 0)
            ranges[cur].x = 0;
        else if (cur != prev) {
            ranges[prev].y = (uint32_t)i;
            ranges[cur].x = (uint32_t)i;
        }
        if (i == L - 1) ranges[cur].y = (uint32_t)L;
        prev = cur;
    }
}

// compact list of non-empty tiles in row-major order (rasterizer_impl.cu:348-365 on the host in the reference)
__global__ void __launch_bounds__(1024) compact_tiles_kernel(int T, const uint2 *__restrict__ ranges,
                                                             const uint2 *__restrict__ ranges_b, int *tile_indices,
                                                             int *status) {
    pdl_enter();
    __shared__ int warp_sums[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < T; start += 1024) {
        const int t = start + threadIdx.x;
        int flag = 0;
        if (t < T) {
            const uint2 r = ranges[t];
            flag = (r.x != r.y) ? 1 : 0;
            if (ranges_b) {
                const uint2 rb = ranges_b[t];
                flag |= (rb.x != rb.y) ? 1 : 0;
            }
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, flag);
        const int within = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) warp_sums[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
        for (int w = 0; w < 32; w++) {
            const int c = warp_sums[w];
            if (w < warp) woff += c;
            total += c;
        }
        if (flag) tile_indices[base + woff + within] = t;
        __syncthreads();
        if (threadIdx.x == 0) base += total;
        __syncthreads();
    }
    for (int t = base + threadIdx.x; t < T; t += 1024) tile_indices[t] = -1;
    if (threadIdx.x == 0) status[DQO_ST_TILE_NUM] = base;
}

struct __align__(16) SplatS {
    float4 r0, r1, c; // {x, y, conic.x, conic.y} {conic.z, opacity, power_reject, -} {r, g, b, Gaussian id bits}
};

struct RenderArgs {
    int W, H, grid_x;
    float fx, fy, cx, cy, scale_mod;
    float opaque_thr, depth_thr, normal_thr, T_thr;
    const uint2 *ranges;
    const uint32_t *point_list;
    const float4 *rec;
    const float *depth;
    const float *view, *means3D, *scales, *rotations, *bg;
    uint32_t *n_contrib;
    float *final_T;
    float *hit_geo;
    size_t plane; // T*256
    float *out_color, *out_depth, *out_hit_cw, *out_hit_dw, *out_T;
    int *out_hit_depth, *out_hit_color, *n_touched;
    // optional (persistent outputs, single-phase): tile_filled[t] != 0 <=> the output pixels of tile t currently hold the
    // fill values of a tile that is not rendered; such a tile is not written again until it is rendered
    uint8_t *tile_filled;
    // two-phase binning (PHASE 1 / 2)
    const uint2 *ranges_b;
    const uint32_t *point_list_b;
    int *unfinished;
    float *state;
    int *status;
};

// Front-to-back blend of one 16x16 tile (forward.cu:636-866).  Warp w owns the 8x4 pixel sub-block
// (bx = (w&1)*8, by = (w>>1)*4).  Every 256-entry batch of the tile's list is staged in shared memory together
// with an 8-bit mask of the sub-blocks each splat can reach; each warp then compacts the batch into its own
// index list and only walks the splats that can contribute to its pixels.  Skipped splats are exactly those the
// reference rejects for all 32 pixels (alpha < 1/255), so every output is unchanged; `contributor` is the list
// position, recovered from the batch index instead of being counted.
//   PHASE 0: the tile's whole list (single-phase binning).
//   PHASE 1: the front list only (nearest Gaussians).  A tile whose pixels have not all terminated when the list ends
//            is flagged unfinished and parks the running (T, C) of its pixels; every other per-pixel quantity is
//            already in the output images.
//   PHASE 2: unfinished tiles resume from the parked state and walk their back list; list positions continue at the
//            length of the front list, so n_contrib and the blend order are those of the concatenated (= reference) list.
// shared-memory accessors of the blend loops: the window address is computed once and kept opaque, so the compiler
// cannot rematerialise it (S2R + LEA) inside the loop as it does for indexed __shared__ arrays under register pressure
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16x2(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

#define RF_LIST_STRIDE 264 // u16 entries per warp list: 256 + sentinel padding, multiple of 4

template <int PHASE, bool NT>
__global__ void __launch_bounds__(256, PHASE == 2 ? 2 : 4) render_forward_kernel(RenderArgs a) {
    pdl_enter();
    __shared__ SplatS s_sp[257]; // one 48-byte record per staged splat (+ an all-zero sentinel that never contributes)
    __shared__ uint8_t s_mask[256];
    __shared__ __align__(16) uint16_t s_list[8][RF_LIST_STRIDE];

    const int tile = blockIdx.x;
    const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    const uint32_t pix_x = tile_x * DQO_TILE + lx, pix_y = tile_y * DQO_TILE + ly;
    const bool inside = pix_x < (uint32_t)a.W && pix_y < (uint32_t)a.H;
    const size_t pix_id = (size_t)a.W * pix_y + pix_x;
    const size_t HW = (size_t)a.W * a.H;
    const size_t sp = (size_t)tile * 256 + ly * 16 + lx;
    uint2 range = a.ranges[tile];
    const uint32_t *__restrict__ point_list = a.point_list;
    int base = 0; // list position of the first entry walked by this launch
    if (PHASE == 2) {
        base = (int)(range.y - range.x);
        range = a.ranges_b[tile];
        if (range.x == range.y) return; // finished in the front phase (or nothing behind it)
        point_list = a.point_list_b;
    }

    if (PHASE != 2 && range.x == range.y) { // tile not rendered: reference fill values (rasterize_points.cu:79-86)
        if (PHASE == 1 && tid == 0) a.unfinished[tile] = 1; // nothing in front: the back phase may still reach it
        if (PHASE == 0 && a.tile_filled) { // object steps render a few tiles of a 1080p image: the rest stays as it is
            const bool filled = a.tile_filled[tile] != 0;
            __syncthreads();
            if (filled) return;
            if (tid == 0) a.tile_filled[tile] = 1;
        }
        if (inside) {
            a.out_color[pix_id] = 0.f;
            a.out_color[HW + pix_id] = 0.f;
            a.out_color[2 * HW + pix_id] = 0.f;
            a.out_depth[pix_id] = 0.f;
            a.out_hit_depth[pix_id] = 0;
            a.out_hit_color[pix_id] = 0;
            a.out_hit_cw[pix_id] = 0.f;
            a.out_hit_dw[pix_id] = 0.f;
            a.out_T[pix_id] = 1.f;
        }
        return;
    }
    const float pixfx = (float)pix_x, pixfy = (float)pix_y;
    const float tile_px = (float)(tile_x * DQO_TILE), tile_py = (float)(tile_y * DQO_TILE);
    const int total = (int)(range.y - range.x);
    const int rounds = (total + 255) / 256;
    if (PHASE == 0 && a.tile_filled && tid == 0) a.tile_filled[tile] = 0;
    if (tid == 0) s_sp[256].r0 = s_sp[256].r1 = s_sp[256].c = make_float4(0.f, 0.f, 0.f, 0.f); // opacity 0: alpha = 0

    // Per-pixel state.  Two of the reference's flags live inside other values so that the walk needs no predicate <->
    // register traffic: `done` is the SIGN BIT of T (T itself is never negative), and the first-opaque-hit state is one
    // integer: -1 no hit yet, -2 hit found by an earlier launch (PHASE 2 resume), >= 0 the Gaussian hit in this launch.
    float T = 1.0f, end_T = 1.0f;
    int ncontrib = 0;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f;
    float depth_ = 0.f;
    int hit_state = -1;
    int hit_id = -1, hit_color_id = -1;
    float cw_max = -1.f, hit_dw = 0.f;
    bool was_done = false;
    if (PHASE == 2 && base > 0 && inside) { // resume: parked (T, C) + what the front phase wrote to the outputs
        T = a.state[sp];
        C0 = a.state[a.plane + sp];
        C1 = a.state[2 * a.plane + sp];
        C2 = a.state[3 * a.plane + sp];
        was_done = T < 0.f; // parked as -1: terminated in the front phase, nothing left to do
        ncontrib = (int)a.n_contrib[sp];
        end_T = a.final_T[sp];
        hit_id = a.out_hit_depth[pix_id];
        hit_state = hit_id >= 0 ? -2 : -1;
        hit_color_id = a.out_hit_color[pix_id];
        cw_max = hit_color_id >= 0 ? a.out_hit_cw[pix_id] : -1.f;
        hit_dw = a.out_hit_dw[pix_id];
        depth_ = a.out_depth[pix_id];
    }
    if (!inside) T = -1.0f;
    const float opaque_thr = a.opaque_thr, T_thr = a.T_thr;
    const uint32_t sp_base = smem_addr(s_sp), list_base = smem_addr(&s_list[warp][0]);

    // The walk keeps the warp converged: every lane executes every list entry of its warp and folds the reference's
    // early-outs (forward.cu:766-848) into predicates, so the body is straight-line code.  Lists are padded with the
    // sentinel to an even length and walked two entries per trip (one all-terminated vote per trip).  The geometry of the
    // first opaque hit (forward.cu:785-812) depends only on the hit Gaussian and the pixel: the loop records the event,
    // the plane / normal math runs once per pixel after the walk.
    float hit_w = 0.f; // alpha * T at the hit event
    // PHASE 2 walks the long tails of a few hundred unfinished tiles (thousands of entries each, most pixels already
    // terminated): one or two resident blocks per SM, every round a chain of two dependent memory round trips (list
    // entry -> splat record) with little blending behind it.  There the next round's entry is fetched into registers
    // while the current round is compacted and blended (software pipelining; the other phases have the occupancy to
    // hide the latency and no registers to spare).
    int pf_id = 0;
    float4 pf0 = make_float4(0.f, 0.f, 0.f, 0.f), pf1 = pf0, pf2 = pf0;
    if (PHASE == 2 && tid < total) {
        pf_id = (int)point_list[range.x + tid];
        pf0 = __ldg(&a.rec[3 * (size_t)pf_id]);
        pf1 = __ldg(&a.rec[3 * (size_t)pf_id + 1]);
        pf2 = __ldg(&a.rec[3 * (size_t)pf_id + 2]);
    }
    int i = 0;
    for (; i < rounds; i++) {
        if (__syncthreads_count(__float_as_int(T) < 0) == 256) break;
        const int progress = i * 256 + tid;
        const int n = min(256, total - i * 256);
        if (progress < total) {
            int id;
            float4 r0, r1, r2;
            if (PHASE == 2) {
                id = pf_id; r0 = pf0; r1 = pf1; r2 = pf2;
            } else {
                id = (int)point_list[range.x + progress];
                r0 = __ldg(&a.rec[3 * (size_t)id]);
                r1 = __ldg(&a.rec[3 * (size_t)id + 1]);
                r2 = __ldg(&a.rec[3 * (size_t)id + 2]);
            }
            s_sp[tid].r0 = r0;
            s_sp[tid].r1 = r1;
            s_sp[tid].c = make_float4(r2.x, r2.y, r2.z, __int_as_float(id));
            s_mask[tid] = (uint8_t)subblock_mask(r0.x, r0.y, r0.z, r0.w, r1.x, r2.w, r1.w, tile_px, tile_py);
        }
        __syncthreads();
        if (PHASE == 2 && progress + 256 < total) {
            pf_id = (int)point_list[range.x + progress + 256];
            pf0 = __ldg(&a.rec[3 * (size_t)pf_id]);
            pf1 = __ldg(&a.rec[3 * (size_t)pf_id + 1]);
            pf2 = __ldg(&a.rec[3 * (size_t)pf_id + 2]);
        }
        // per-warp compaction of the batch (order preserved)
        int cnt = 0;
        if (!__all_sync(0xFFFFFFFFu, __float_as_int(T) < 0)) {
            for (int b = 0; b < n; b += 32) {
                const int j = b + lane;
                const bool m = (j < n) && ((s_mask[j] >> warp) & 1);
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, m);
                if (m) s_list[warp][cnt + __popc(bal & ((1u << lane) - 1))] = (uint16_t)j;
                cnt += __popc(bal);
            }
            if (lane == 0) s_list[warp][cnt] = 256; // sentinel: pads an odd list
            __syncwarp();
        }
        int last_j = -1; // batch slot of the last entry accumulated by this pixel
        for (int k = 0; k < cnt; k += 2) {
            if (__all_sync(0xFFFFFFFFu, __float_as_int(T) < 0)) break;
            const uint32_t jj = lds_u16x2(list_base + 2 * k);
            unsigned nt_m[2] = {0u, 0u};
            int nt_id[2] = {0, 0};
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int j = h ? (int)(jj >> 16) : (int)(jj & 0xFFFFu);
                const uint32_t e = sp_base + (uint32_t)j * 48u;
                const float4 r0 = lds128(e);
                const float4 r1 = lds128(e + 16);
                const float4 r2 = lds128(e + 32);
                const float dx = fsub(r0.x, pixfx), dy = fsub(r0.y, pixfy);
                const float power = ffma(ffma(dx, fmul(dx, r0.z), fmul(dy, fmul(dy, r1.x))), -0.5f, -fmul(dy, fmul(dx, r0.w)));
                const float alpha = fminf(0.99f, fmul(r1.y, expf(power)));
                // live: the pixel is still being blended and the pair passes the reference's rejections (power > 0,
                // alpha < 1/255; the staged power_reject bound is implied by the alpha test)
                const bool live = (__float_as_int(T) >= 0) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                const float w = fmul(alpha, T);
                const float test_T = fmul(T, fsub(1.0f, alpha));
                const int id = __float_as_int(r2.w);
                const bool opq = live && (hit_state == -1) && (alpha >= opaque_thr); // first opaque hit of this pixel
                hit_state = opq ? id : hit_state;
                hit_w = opq ? w : hit_w;
                const bool acc = live && (test_T >= T_thr);
                const bool stop = live && (test_T < T_thr) && (hit_state != -1);
                C0 = acc ? ffma(r2.x, w, C0) : C0;
                C1 = acc ? ffma(r2.y, w, C1) : C1;
                C2 = acc ? ffma(r2.z, w, C2) : C2;
                const bool better = acc && (w > cw_max);
                cw_max = better ? w : cw_max;
                hit_color_id = better ? id : hit_color_id;
                if (NT) { // pairs with T > 0.5 (forward.cu:836-839): one vote per entry, one rare branch per trip
                    nt_m[h] = __ballot_sync(0xFFFFFFFFu, acc && test_T > 0.5f);
                    nt_id[h] = id;
                }
                last_j = acc ? j : last_j;
                end_T = acc ? test_T : end_T;
                // a terminated pixel keeps its T (the output needs it) with the sign bit set
                T = live ? (stop ? -T : test_T) : T;
            }
            if (NT && (nt_m[0] | nt_m[1]) != 0 && lane == 0) {
                if (nt_m[0]) atomicAdd(&a.n_touched[nt_id[0]], __popc(nt_m[0]));
                if (nt_m[1]) atomicAdd(&a.n_touched[nt_id[1]], __popc(nt_m[1]));
            }
        }
        if (last_j >= 0) ncontrib = base + i * 256 + last_j + 1;
    }
    const bool done = __float_as_int(T) < 0;
    T = fabsf(T);
    if (hit_state >= 0) { // geometry of the first opaque hit (forward.cu:785-812)
        const int id = hit_state;
        const float sx = a.scales[3 * id], sy = a.scales[3 * id + 1], sz = a.scales[3 * id + 2];
        const float4 q = reinterpret_cast<const float4 *>(a.rotations)[id];
        const QuatMat R = quat_to_glm(q.x, q.y, q.z, q.w);
        const int ax = arg_min3(sx, sy, sz);
        const float nx = ax == 0 ? R.c0[0] : (ax == 1 ? R.c0[1] : R.c0[2]);
        const float ny = ax == 0 ? R.c1[0] : (ax == 1 ? R.c1[1] : R.c1[2]);
        const float nz = ax == 0 ? R.c2[0] : (ax == 1 ? R.c2[1] : R.c2[2]);
        const float smax = fmul(fmaxf(fmaxf(sx, sy), sz), a.scale_mod);
        const float *v = a.view;
        const float ncx = xform_row3(v, 0, nx, ny, nz), ncy = xform_row3(v, 1, nx, ny, nz),
                    ncz = xform_row3(v, 2, nx, ny, nz);
        const float wx = a.means3D[3 * id], wy = a.means3D[3 * id + 1], wz = a.means3D[3 * id + 2];
        const float pcx = xform_row(v, 0, wx, wy, wz), pcy = xform_row(v, 1, wx, wy, wz),
                    pcz = xform_row(v, 2, wx, wy, wz);
        const float3 ray = pixel_ray(pix_x, pix_y, a.fx, a.fy, a.cx, a.cy);
        const float num = dot3_ref(pcx, ncx, pcy, ncy, pcz, ncz);
        const float den = dot3_ref(ray.x, ncx, ray.y, ncy, ray.z, ncz);
        const float t = (float)((double)num / ((double)den + 1e-8));
        const float hx = fmul(t, ray.x), hy = fmul(t, ray.y), hz = fmul(t, ray.z);
        const float depth_distance = fabsf(fsub(hz, pcz));
        const float angle_distance = fabsf(den);
        hit_id = id;
        hit_dw = hit_w;
        if (depth_distance <= fmul(smax, a.depth_thr) && angle_distance >= a.normal_thr)
            depth_ = hz;
        else
            depth_ = a.depth[id];
        a.hit_geo[sp] = ncx;
        a.hit_geo[a.plane + sp] = ncy;
        a.hit_geo[2 * a.plane + sp] = ncz;
        a.hit_geo[3 * a.plane + sp] = hx;
        a.hit_geo[4 * a.plane + sp] = hy;
        a.hit_geo[5 * a.plane + sp] = hz;
    }
    const float hit_cw = hit_color_id >= 0 ? cw_max : 0.f;
    if (tid == 0 && a.status) atomicAdd(&a.status[DQO_ST_WALKED], min(total, i * 256));
    if (PHASE == 1) {
        const bool unfinished = __syncthreads_count(done) < 256;
        if (tid == 0) {
            a.unfinished[tile] = unfinished ? 1 : 0;
            if (unfinished) atomicAdd(&a.status[DQO_ST_UNFINISHED], 1);
        }
        if (unfinished && inside) {
            a.state[sp] = done ? -1.0f : T;
            a.state[a.plane + sp] = C0;
            a.state[2 * a.plane + sp] = C1;
            a.state[3 * a.plane + sp] = C2;
        }
    }
    if (inside && !was_done) {
        a.final_T[sp] = end_T;
        a.n_contrib[sp] = (uint32_t)ncontrib;
        a.out_color[pix_id] = ffma(T, a.bg[0], C0);
        a.out_color[HW + pix_id] = ffma(T, a.bg[1], C1);
        a.out_color[2 * HW + pix_id] = ffma(T, a.bg[2], C2);
        a.out_depth[pix_id] = depth_;
        a.out_hit_depth[pix_id] = hit_id;
        a.out_hit_color[pix_id] = hit_color_id;
        a.out_hit_cw[pix_id] = hit_cw;
        a.out_hit_dw[pix_id] = hit_dw;
        a.out_T[pix_id] = end_T;
    }
}

// Re-blend of an already binned view with another colour per Gaussian.  The reference renders its
// semantic / instance images by running the whole rasterizer again with colors_precomp (SLAM/render.py:227-262):
// preprocess, duplicate, sort and ranges are repeated although only the colour differs.  This kernel walks the lists of
// the existing forward state (front list, then back list) with exactly the forward's alpha / hit / termination logic
// (forward.cu:760-848), so the image is bit-identical to what a second full forward would return in `out_color`.
struct ExtraArgs {
    int W, H, grid_x;
    float opaque_thr, T_thr;
    const uint2 *ranges, *ranges_b;
    const uint32_t *point_list, *point_list_b;
    const float4 *rec;
    const float *colors; // [P,3]
    const float *bg;
    float *out_color;
};
__global__ void __launch_bounds__(256) blend_extra_kernel(ExtraArgs a) {
    pdl_enter();
    __shared__ SplatS s_sp[256];
    __shared__ uint8_t s_mask[256];
    __shared__ uint8_t s_list[8][256];

    const int tile = blockIdx.x;
    const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    const uint32_t pix_x = tile_x * DQO_TILE + lx, pix_y = tile_y * DQO_TILE + ly;
    const bool inside = pix_x < (uint32_t)a.W && pix_y < (uint32_t)a.H;
    const size_t pix_id = (size_t)a.W * pix_y + pix_x;
    const size_t HW = (size_t)a.W * a.H;
    const uint2 range_a = a.ranges[tile];
    uint2 range_b = make_uint2(0, 0);
    if (a.ranges_b) range_b = a.ranges_b[tile];
    if (range_a.x == range_a.y && range_b.x == range_b.y) { // tile not rendered: fill value (rasterize_points.cu:79)
        if (inside) a.out_color[pix_id] = a.out_color[HW + pix_id] = a.out_color[2 * HW + pix_id] = 0.f;
        return;
    }
    const float pixfx = (float)pix_x, pixfy = (float)pix_y;
    const float tile_px = (float)(tile_x * DQO_TILE), tile_py = (float)(tile_y * DQO_TILE);
    bool done = !inside, hit = false, stop = false;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    for (int seg = 0; seg < 2 && !stop; seg++) {
        const uint2 range = seg ? range_b : range_a;
        const uint32_t *__restrict__ list = seg ? a.point_list_b : a.point_list;
        const int total = (int)(range.y - range.x);
        const int rounds = (total + 255) / 256;
        for (int i = 0; i < rounds; i++) {
            if (__syncthreads_count(done) == 256) {
                stop = true;
                break;
            }
            const int progress = i * 256 + tid;
            const int n = min(256, total - i * 256);
            if (progress < total) {
                const int id = (int)list[range.x + progress];
                const float4 r0 = __ldg(&a.rec[3 * (size_t)id]);
                const float4 r1 = __ldg(&a.rec[3 * (size_t)id + 1]);
                const float4 r2 = __ldg(&a.rec[3 * (size_t)id + 2]);
                s_sp[tid].r0 = r0;
                s_sp[tid].r1 = r1;
                s_sp[tid].c = make_float4(a.colors[3 * (size_t)id], a.colors[3 * (size_t)id + 1], a.colors[3 * (size_t)id + 2], 0.f);
                s_mask[tid] = (uint8_t)subblock_mask(r0.x, r0.y, r0.z, r0.w, r1.x, r2.w, r1.w, tile_px, tile_py);
            }
            __syncthreads();
            int cnt = 0;
            if (!__all_sync(0xFFFFFFFFu, done)) {
                for (int b = 0; b < n; b += 32) {
                    const int j = b + lane;
                    const bool m = (j < n) && ((s_mask[j] >> warp) & 1);
                    const unsigned bal = __ballot_sync(0xFFFFFFFFu, m);
                    if (m) s_list[warp][cnt + __popc(bal & ((1u << lane) - 1))] = (uint8_t)j;
                    cnt += __popc(bal);
                }
                __syncwarp();
            }
            for (int k = 0; !done && k < cnt; k++) {
                const int j = s_list[warp][k];
                const float4 r0 = s_sp[j].r0;
                const float4 r1 = s_sp[j].r1;
                const float dx = fsub(r0.x, pixfx), dy = fsub(r0.y, pixfy);
                const float power = ffma(ffma(dx, fmul(dx, r0.z), fmul(dy, fmul(dy, r1.x))), -0.5f, -fmul(dy, fmul(dx, r0.w)));
                if (power > 0.0f || power < r1.z) continue;
                const float alpha = fminf(0.99f, fmul(r1.y, expf(power)));
                if (alpha < 1.0f / 255.0f) continue;
                if (alpha >= a.opaque_thr) hit = true;
                const float test_T = fmul(T, fsub(1.0f, alpha));
                if (test_T < a.T_thr && hit) {
                    done = true;
                    continue;
                }
                if (test_T >= a.T_thr) {
                    const float w = fmul(alpha, T);
                    const float4 c = s_sp[j].c;
                    C0 = ffma(c.x, w, C0);
                    C1 = ffma(c.y, w, C1);
                    C2 = ffma(c.z, w, C2);
                }
                T = test_T;
            }
            __syncthreads();
        }
    }
    if (inside) {
        a.out_color[pix_id] = ffma(T, a.bg[0], C0);
        a.out_color[HW + pix_id] = ffma(T, a.bg[1], C1);
        a.out_color[2 * HW + pix_id] = ffma(T, a.bg[2], C2);
    }
}

// ------------------------------------------------------------------------------------------------
// state export for parity tests
// ------------------------------------------------------------------------------------------------
template <typename KeyT>
__global__ void export_instances_kernel(int64_t capacity, const int *__restrict__ status,
                                        const KeyT *__restrict__ keys, const uint32_t *__restrict__ vals,
                                        const float *__restrict__ depth, uint64_t *out_keys, uint32_t *out_list) {
    pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= capacity) return;
    const int64_t L = status[DQO_ST_OVERFLOW] ? 0 : status[DQO_ST_NUM_RENDERED];
    if (i < L) {
        const uint32_t id = vals[i];
        const uint32_t dbits = __float_as_uint(depth[id]);
        if (out_keys) out_keys[i] = ((uint64_t)keys[i] << 32) | dbits;
        if (out_list) out_list[i] = id;
    } else {
        if (out_keys) out_keys[i] = 0;
        if (out_list) out_list[i] = 0;
    }
}
__global__ void export_pixels_kernel(int W, int H, int grid_x, const uint32_t *__restrict__ n_contrib,
                                     const float *__restrict__ final_T, const uint2 *__restrict__ ranges,
                                     const uint2 *__restrict__ ranges_b, uint32_t *out_nc, float *out_T) {
    pdl_enter();
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int lx = tid & 15, ly = tid >> 4;
    const int px = (tile % grid_x) * 16 + lx, py = (tile / grid_x) * 16 + ly;
    if (px >= W || py >= H) return;
    const uint2 r = ranges[tile];
    bool rendered = r.x != r.y;
    if (ranges_b) {
        const uint2 rb = ranges_b[tile];
        rendered |= rb.x != rb.y;
    }
    const size_t sp = (size_t)tile * 256 + tid;
    if (out_nc) out_nc[(size_t)py * W + px] = rendered ? n_contrib[sp] : 0u;
    if (out_T) out_T[(size_t)py * W + px] = rendered ? final_T[sp] : 1.0f;
}
__global__ void export_gauss_kernel(int P, const float4 *__restrict__ rec, const uint32_t *__restrict__ tiles,
                                    const uint8_t *__restrict__ flags, const float *__restrict__ depth_in, float *means2D,
                                    float *depths, float *conic_o, float *rgb, uint32_t *tiles_out) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const bool valid = (flags[i] & 0x80) != 0;
    float4 r0 = make_float4(0, 0, 0, 0), r1 = r0, r2 = r0;
    if (valid) {
        r0 = rec[3 * (size_t)i];
        r1 = rec[3 * (size_t)i + 1];
        r2 = rec[3 * (size_t)i + 2];
    }
    if (means2D) {
        means2D[2 * i] = r0.x;
        means2D[2 * i + 1] = r0.y;
    }
    if (depths) depths[i] = valid ? depth_in[i] : 0.f;
    if (conic_o) {
        conic_o[4 * i] = r0.z;
        conic_o[4 * i + 1] = r0.w;
        conic_o[4 * i + 2] = r1.x;
        conic_o[4 * i + 3] = r1.y;
    }
    if (rgb) {
        rgb[3 * i] = r2.x;
        rgb[3 * i + 1] = r2.y;
        rgb[3 * i + 2] = r2.z;
    }
    if (tiles_out) tiles_out[i] = tiles[i];
}
} // namespace dqo

using namespace dqo;

extern "C" size_t dqo_rast_geom_bytes(int32_t P) {
    GeomLayout L;
    if (make_geom_layout(P, &L)) return 0;
    return L.total;
}
extern "C" size_t dqo_rast_binning_bytes(int64_t C) {
    BinLayout L;
    if (make_bin_layout(C, &L)) return 0;
    return L.total;
}
extern "C" size_t dqo_rast_image_bytes(int32_t W, int32_t H) {
    ImgLayout L;
    make_img_layout(W, H, &L);
    return L.total;
}

extern "C" int dqo_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                                uint8_t *present, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !projmatrix || !present))) {
        set_error("dqo_mark_visible: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    if (P == 0) return DQO_OK;
    launch_pdl(mark_visible_kernel, dim3((P + 255) / 256), dim3(256), 0, stream, P, means3D, viewmatrix, projmatrix, present);
    DQO_LAUNCH_CHECK("mark_visible", 0, stream);
    return DQO_OK;
}

namespace dqo {
int rast_forward_impl(const dqo_rast_settings *s, const float *background, const float *means3D, const float *shs,
                      const float *f_rest, const float *colors_precomp, const float *opacities, const float *scales,
                      const float *rotations, const float *cov3D_precomp, const float *viewmatrix,
                      const float *projmatrix, const float *campos, const int32_t *tile_mask, void *geom_buffer,
                      void *binning_buffer, int64_t capacity, void *image_buffer, int32_t *tile_indices, float *out_color,
                      float *out_depth, int32_t *out_hit_depth, int32_t *out_hit_color, float *out_hit_color_weight,
                      float *out_hit_depth_weight, float *out_T, int32_t *radii, int32_t *n_touched, int32_t *status,
                      void *stream_, void (*pre_hook)(void *, void *), void *hook_ctx, uint8_t *tile_filled);
}

extern "C" int dqo_rast_forward(const dqo_rast_settings *s, const float *background, const float *means3D,
                                const float *shs, const float *colors_precomp, const float *opacities,
                                const float *scales, const float *rotations, const float *cov3D_precomp,
                                const float *viewmatrix, const float *projmatrix, const float *campos,
                                const int32_t *tile_mask, void *geom_buffer, void *binning_buffer,
                                int64_t capacity, void *image_buffer, int32_t *tile_indices, float *out_color,
                                float *out_depth, int32_t *out_hit_depth, int32_t *out_hit_color,
                                float *out_hit_color_weight, float *out_hit_depth_weight, float *out_T, int32_t *radii,
                                int32_t *n_touched, int32_t *status, void *stream_) {
    if (!tile_indices) {
        set_error("dqo_rast_forward: tile_indices is NULL");
        return DQO_ERR_INVALID_ARG;
    }
    return rast_forward_impl(s, background, means3D, shs, nullptr, colors_precomp, opacities, scales, rotations,
                             cov3D_precomp, viewmatrix, projmatrix, campos, tile_mask, geom_buffer, binning_buffer,
                             capacity, image_buffer, tile_indices, out_color, out_depth, out_hit_depth, out_hit_color,
                             out_hit_color_weight, out_hit_depth_weight, out_T, radii, n_touched, status, stream_, nullptr,
                             nullptr, nullptr);
}

int dqo::rast_forward_impl(const dqo_rast_settings *s, const float *background, const float *means3D,
                                const float *shs, const float *f_rest, const float *colors_precomp, const float *opacities,
                                const float *scales, const float *rotations, const float *cov3D_precomp,
                                const float *viewmatrix, const float *projmatrix, const float *campos,
                                const int32_t *tile_mask, void *geom_buffer, void *binning_buffer,
                                int64_t capacity, void *image_buffer, int32_t *tile_indices, float *out_color,
                                float *out_depth, int32_t *out_hit_depth, int32_t *out_hit_color,
                                float *out_hit_color_weight, float *out_hit_depth_weight, float *out_T, int32_t *radii,
                                int32_t *n_touched, int32_t *status, void *stream_, void (*pre_hook)(void *, void *),
                                void *hook_ctx, uint8_t *tile_filled) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!s || s->P < 0 || s->W <= 0 || s->H <= 0 || !status) {
        set_error("dqo_rast_forward: invalid settings");
        return DQO_ERR_INVALID_ARG;
    }
    if (!background || !viewmatrix || !projmatrix || !campos || !tile_mask || !image_buffer ||
        !out_color || !out_depth || !out_hit_depth || !out_hit_color || !out_hit_color_weight ||
        !out_hit_depth_weight || !out_T) {
        set_error("dqo_rast_forward: null pointer argument");
        return DQO_ERR_INVALID_ARG;
    }
    const int P = s->P;
    if (P > 0) {
        if (!means3D || !opacities || !geom_buffer || !binning_buffer || !radii || capacity <= 0) {
            set_error("dqo_rast_forward: null pointer argument");
            return DQO_ERR_INVALID_ARG;
        }
        if ((shs == nullptr) == (colors_precomp == nullptr)) {
            set_error("Please provide excatly one of either SHs or precomputed colors!");
            return DQO_ERR_INVALID_ARG;
        }
        if (!scales || !rotations) {
            // the reference blend dereferences scales/rotations unconditionally (forward.cu:780)
            set_error("scale/rotation pair is required by the depth rasterizer (cov3D_precomp alone is not supported)");
            return DQO_ERR_INVALID_ARG;
        }
        if (shs && (s->M <= 0 || (s->D + 1) * (s->D + 1) > s->M || s->D > 3 || s->D < 0)) {
            set_error("dqo_rast_forward: SH degree %d incompatible with %d coefficients", s->D, s->M);
            return DQO_ERR_INVALID_ARG;
        }
    }
    const int debug = s->debug;
    ImgLayout IL;
    make_img_layout(s->W, s->H, &IL);
    char *img = (char *)image_buffer;
    uint2 *ranges = (uint2 *)(img + IL.ranges);
    const int T = IL.T;
    if ((int64_t)T >= (1ll << 31)) {
        set_error("image too large");
        return DQO_ERR_INVALID_ARG;
    }
    pdl_scope(s->P);
    stage_mark(stream, ST_BEGIN_FWD);
    nvtx_push("dqo_rast_forward");
    ClearArgs clr;
    clr.n = 0;
    auto clear_add = [&](void *ptr, size_t bytes) {
        clr.ptr[clr.n] = ptr;
        clr.bytes[clr.n] = bytes;
        clr.n++;
    };
    clear_add(ranges, (size_t)T * sizeof(uint2));
    clear_add(status, DQO_ST_WORDS * sizeof(int));

    const float focal_y = s->H / (2.0f * s->tanfovy);
    const float focal_x = s->W / (2.0f * s->tanfovx);
    GeomLayout GL;
    BinLayout BL;
    const float4 *rec = nullptr;
    const float *depth = nullptr;
    const bool keys16 = T < 65535;
    const bool two_phase = P > 0 && s->front_instances > 0;
    const int64_t front = two_phase ? s->front_instances : 0, back = two_phase ? s->back_instances : 0;
    if (two_phase && (front % 256 != 0 || back <= 0 || front + back > capacity)) {
        set_error("dqo_rast_forward: front_instances must be a multiple of 256 and front + back instances must fit the "
                  "instance capacity");
        nvtx_pop();
        return DQO_ERR_INVALID_ARG;
    }
    uint2 *ranges_b = two_phase ? (uint2 *)(img + IL.ranges_b) : nullptr;
    if (two_phase) clear_add(ranges_b, (size_t)T * sizeof(uint2));
    if (two_phase && tile_filled) clear_add(tile_filled, ((size_t)T + 3) / 4 * 4); // the two-phase blends do not keep the flags
    uint32_t *vals_a = nullptr, *vals_b = nullptr;
    char *keys_a = nullptr, *keys_b = nullptr, *sort_temp = nullptr;
    const uint32_t *d_order = nullptr, *d_tiles = nullptr, *d_mask_bits = nullptr;
    uint32_t *d_offsets = nullptr;
    const uint2 *d_rect = nullptr;
    uint32_t *d_sums = nullptr, *d_tiles_b = nullptr, *d_rank_tiles = nullptr;
    size_t sums_stride = 0;
    int emit_blocks = 0;
    ForkJoin *fj = debug ? nullptr : fork_join(stream); // side stream + events owned by (device, caller stream)
    if (P <= 0) {
        launch_pdl(clear_regions_kernel, dim3(64), dim3(256), 0, stream, clr);
        DQO_LAUNCH_CHECK("clear", debug, stream);
    }
    int lazy_shmode = 0;
    ColorArgs ca = {};
    if (P > 0) {
        if (make_geom_layout(P, &GL)) return DQO_ERR_WORKSPACE;
        if (make_bin_layout(capacity, &BL)) return DQO_ERR_WORKSPACE;
        char *geom = (char *)geom_buffer;
        char *bin = (char *)binning_buffer;
        uint32_t *mask_bits = (uint32_t *)(img + IL.mask_bits);
        clear_add(geom + GL.sums, 2 * GL.sums_stride);
        { // scratch headers of the depth sort and of the tile sorts (front and back share one scratch, disjoint pass slots)
            void *q;
            size_t nb;
            sort_clear_region(geom + GL.sort_temp, P, 32, &q, &nb);
            clear_add(q, nb);
            sort_clear_region(bin + BL.sort_temp, capacity > 0 ? capacity : 1, 32, &q, &nb);
            clear_add(q, nb);
        }
        launch_pdl(clear_regions_kernel, dim3(64), dim3(256), 0, stream, clr);
        DQO_LAUNCH_CHECK("clear", debug, stream);
        // fork: depth keys + the (depth, id) sort of the Gaussians (stable LSD sort on the depth bits) on the side
        // stream, concurrently with the rest of the preprocess
        uint32_t *order = (uint32_t *)(geom + GL.order);
        launch_pdl(depth_key_kernel, dim3((P + 255) / 256), dim3(256), 0, stream, P, means3D, viewmatrix, projmatrix, (uint32_t *)(geom + GL.depth_key));
        DQO_LAUNCH_CHECK("depth keys", debug, stream);
        cudaStream_t sort_stream = stream;
        if (fj) {
            DQO_CUDA_CHECK(cudaEventRecord(fj->ev[0], stream));
            DQO_CUDA_CHECK(cudaStreamWaitEvent(fj->side, fj->ev[0], 0));
            sort_stream = fj->side;
        }
        // The sort chain (12 short kernels) is the critical path of this stage, so it is enqueued first; the preprocess
        // fills the SMs around its high-priority kernels.  4 passes (even): the sorted ids end up in `order` (vals_a),
        // values are implicit (value = index) so no iota array is read.
        {
            const int rc = radix_sort_pairs<uint32_t>((uint32_t *)(geom + GL.depth_key), (uint32_t *)(geom + GL.depth_key2), order, (uint32_t *)(geom + GL.ids), true,
                                                      nullptr, nullptr, P, 32, geom + GL.sort_temp, sort_stream, 0, false, P);
            if (rc) return rc;
            if (debug) DQO_CUDA_CHECK(cudaStreamSynchronize(sort_stream));
        }
        // work of the caller that only has to precede the preprocess (the fused step's activation kernel): enqueued
        // behind the fork so that the sort chain, which only needs the positions, starts first
        if (pre_hook) pre_hook(hook_ctx, stream_);
        {
            const int nw = IL.tiles_y * IL.mask_words;
            launch_pdl(mask_bits_kernel, dim3((nw + 7) / 8), dim3(256), 0, stream, IL.tiles_x, IL.tiles_y, IL.mask_words, tile_mask, mask_bits);
            DQO_LAUNCH_CHECK("mask bits", debug, stream);
            launch_pdl(mask_sat_kernel, dim3(1), dim3(1024), 0, stream, IL.tiles_x, IL.tiles_y, IL.mask_words, mask_bits,
                       (uint32_t *)(img + IL.sat));
            DQO_LAUNCH_CHECK("mask summed-area table", debug, stream);
        }
        PreArgs pa;
        pa.P = P; pa.D = s->D; pa.M = s->M; pa.W = s->W; pa.H = s->H;
        pa.color_sigma = s->color_sigma; pa.scale_modifier = s->scale_modifier;
        pa.tanfovx = s->tanfovx; pa.tanfovy = s->tanfovy; pa.focal_x = focal_x; pa.focal_y = focal_y;
        pa.cx = s->cx; pa.cy = s->cy; pa.grid_x = IL.tiles_x; pa.grid_y = IL.tiles_y;
        pa.prefiltered = s->prefiltered;
        // SH colours on demand: only with two-phase binning (else every rank is emitted) and the staged SH layouts
        lazy_shmode = 0;
        if (two_phase && !colors_precomp && shs && s->M == 16 && (uintptr_t)shs % 16 == 0)
            lazy_shmode = f_rest ? ((uintptr_t)f_rest % 16 == 0 ? 2 : 0) : 1;
        pa.lazy_color = lazy_shmode != 0;
        pa.means3D = means3D; pa.scales = scales; pa.rotations = rotations; pa.opacities = opacities;
        pa.shs = shs; pa.f_rest = f_rest; pa.cov3D_precomp = cov3D_precomp; pa.colors_precomp = colors_precomp;
        pa.view = viewmatrix; pa.proj = projmatrix; pa.campos = campos;
        pa.mask_bits = mask_bits; pa.mask_words = IL.mask_words; pa.sat = (const uint32_t *)(img + IL.sat);
        pa.radii = radii; pa.n_touched = n_touched;
        pa.rec = (float4 *)(geom + GL.rec);
        pa.depth = (float *)(geom + GL.depth);
        pa.depth_key = (uint32_t *)(geom + GL.depth_key);
        pa.ids = (uint32_t *)(geom + GL.ids);
        pa.tiles = (uint32_t *)(geom + GL.tiles);
        pa.rect = (uint2 *)(geom + GL.rect);
        pa.clamped = (uint8_t *)(geom + GL.clamped);
        pa.status = status;
        rec = pa.rec;
        depth = pa.depth;
        const bool staged = shs && !f_rest && s->M == 16 && ((uintptr_t)shs % 16 == 0);
        const int pre_blocks = (P + PRE_THREADS - 1) / PRE_THREADS;
        const size_t smem = pa.lazy_color ? 0 : (size_t)(PRE_THREADS / 32) * 32 * SH_ROW_Q * sizeof(float4);
        if (f_rest) {
            if (s->M != 16 || (uintptr_t)shs % 16 || (uintptr_t)f_rest % 16) {
                set_error("split SH input requires M == 16 and 16-byte aligned f_dc / f_rest");
                return DQO_ERR_INVALID_ARG;
            }
            launch_pdl(preprocess_kernel<2>, dim3(pre_blocks), dim3(PRE_THREADS), smem, stream, pa);
        } else if (staged) {
            launch_pdl(preprocess_kernel<1>, dim3(pre_blocks), dim3(PRE_THREADS), smem, stream, pa);
        } else {
            launch_pdl(preprocess_kernel<0>, dim3(pre_blocks), dim3(PRE_THREADS), 0, stream, pa);
        }
        DQO_LAUNCH_CHECK("preprocess", debug, stream);
        stage_mark(stream, ST_PREPROCESS);
        if (fj) { // join
            DQO_CUDA_CHECK(cudaEventRecord(fj->ev[1], fj->side));
            DQO_CUDA_CHECK(cudaStreamWaitEvent(stream, fj->ev[1], 0));
        }
        stage_mark(stream, ST_DEPTH_SORT);

        vals_a = (uint32_t *)(bin + BL.vals_a);
        vals_b = (uint32_t *)(bin + BL.vals_b);
        keys_a = bin + BL.keys_a;
        keys_b = bin + BL.keys_b;
        sort_temp = bin + BL.sort_temp;
        d_order = order;
        d_tiles = pa.tiles;
        d_offsets = (uint32_t *)(geom + GL.offsets);
        d_rect = pa.rect;
        d_mask_bits = mask_bits;
        d_sums = (uint32_t *)(geom + GL.sums);
        d_tiles_b = (uint32_t *)(geom + GL.tiles_b);
        d_rank_tiles = (uint32_t *)(geom + GL.depth_key2); // the depth sort's ping-pong buffer is free once it has joined
        sums_stride = GL.sums_stride / 4;
        emit_blocks = GL.emit_blocks;
        ca.P = P; ca.D = s->D; ca.front = front;
        ca.order = d_order; ca.offsets = d_offsets; ca.tiles = d_rank_tiles; ca.tiles_rank = d_tiles_b;
        ca.means3D = means3D; ca.shs = shs; ca.f_rest = f_rest; ca.campos = campos;
        ca.rec = pa.rec; ca.clamped = pa.clamped;
    }
    // lazy colours of one binning phase on the side stream (fork here, join in front of that phase's blend)
    auto color_fork = [&](int phase) -> int {
        if (!lazy_shmode) return DQO_OK;
        cudaStream_t cs = stream;
        if (fj) {
            DQO_CUDA_CHECK(cudaEventRecord(fj->ev[phase == 1 ? 4 : 6], stream));
            DQO_CUDA_CHECK(cudaStreamWaitEvent(fj->side, fj->ev[phase == 1 ? 4 : 6], 0));
            cs = fj->side;
        }
        ca.phase = phase;
        const unsigned blocks = (unsigned)((P + LC_THREADS - 1) / LC_THREADS);
        if (lazy_shmode == 2) launch_pdl(lazy_color_kernel<2>, dim3(blocks), dim3(LC_THREADS), 0, cs, ca);
        else launch_pdl(lazy_color_kernel<1>, dim3(blocks), dim3(LC_THREADS), 0, cs, ca);
        DQO_LAUNCH_CHECK("lazy colours", debug, cs);
        if (fj) DQO_CUDA_CHECK(cudaEventRecord(fj->ev[phase == 1 ? 5 : 7], cs));
        return DQO_OK;
    };
    auto color_join = [&](int phase) -> int {
        if (lazy_shmode && fj) DQO_CUDA_CHECK(cudaStreamWaitEvent(stream, fj->ev[phase == 1 ? 5 : 7], 0));
        return DQO_OK;
    };

    const int sort_bits = tile_sort_bits(T);
    const bool in_a = bin_sorted_in_a(T);
    uint32_t *point_list = in_a ? vals_a : vals_b;
    RenderArgs ra;
    ra.W = s->W; ra.H = s->H; ra.grid_x = IL.tiles_x;
    ra.fx = focal_x; ra.fy = focal_y; ra.cx = s->cx; ra.cy = s->cy; ra.scale_mod = s->scale_modifier;
    ra.opaque_thr = s->opaque_threshold; ra.depth_thr = s->depth_threshold; ra.normal_thr = s->normal_threshold;
    ra.T_thr = s->T_threshold;
    ra.ranges = ranges; ra.point_list = point_list; ra.rec = rec; ra.depth = depth;
    ra.view = viewmatrix; ra.means3D = means3D; ra.scales = scales; ra.rotations = rotations; ra.bg = background;
    ra.n_contrib = (uint32_t *)(img + IL.n_contrib);
    ra.final_T = (float *)(img + IL.final_T);
    ra.hit_geo = (float *)(img + IL.hit_geo);
    ra.plane = (size_t)T * 256;
    ra.out_color = out_color; ra.out_depth = out_depth; ra.out_hit_cw = out_hit_color_weight;
    ra.out_hit_dw = out_hit_depth_weight; ra.out_T = out_T; ra.out_hit_depth = out_hit_depth;
    ra.out_hit_color = out_hit_color; ra.n_touched = s->need_n_touched ? n_touched : nullptr;
    ra.ranges_b = ranges_b; ra.point_list_b = point_list ? point_list + front : nullptr;
    ra.tile_filled = two_phase ? nullptr : tile_filled;
    ra.unfinished = (int *)(img + IL.unfinished);
    ra.state = (float *)(img + IL.state);
    ra.status = status;

    const uint32_t *row_any_b = (const uint32_t *)(img + IL.row_any_b);
    const size_t ksz = keys16 ? 2 : 4;
    // one binning phase: scan + emit -> stable sort by tile id (count read on the device) -> ranges.
    // `n` = slots of this phase's region, which starts at instance `at` of the binning arrays.
    auto bin_phase = [&](int mode, int64_t n, int64_t at, const uint32_t *bits, int count_word, uint2 *out_ranges) -> int {
        void *ka = keys_a + at * ksz, *kb = keys_b + at * ksz;
        uint32_t *va = vals_a + at, *vb = vals_b + at;
        uint32_t *sums = d_sums + (mode == 2 ? sums_stride : 0), *group_sums = sums + emit_blocks;
        if (mode == 2)
            launch_pdl(rank_sums_kernel<2>, dim3(emit_blocks), dim3(256), 0, stream, P, front, d_order, d_tiles, d_offsets, d_rect,
                       (const uint32_t *)(img + IL.sat_b), IL.tiles_x, d_tiles_b, d_rank_tiles, sums, group_sums);
        else
            launch_pdl(rank_sums_kernel<0>, dim3(emit_blocks), dim3(256), 0, stream, P, front, d_order, d_tiles, d_offsets, d_rect,
                       (const uint32_t *)nullptr, IL.tiles_x, (uint32_t *)nullptr, d_rank_tiles, sums, group_sums);
        DQO_LAUNCH_CHECK("rank sums", debug, stream);
        if (mode == 2) {
            const int rcc = color_fork(2);
            if (rcc) return rcc;
        }
#define DQO_EMIT(KT, MODE)                                                                                             \
    launch_pdl(emit_kernel<KT, MODE>, dim3(emit_blocks), dim3(256), 0, stream, P, n, d_order, d_tiles, d_tiles_b, d_rank_tiles, d_offsets, d_rect, bits,  \
                                                          row_any_b, IL.mask_words, IL.tiles_x, (KT *)ka, va, sums,    \
                                                          group_sums, status)
        if (keys16) {
            if (mode == 0) DQO_EMIT(uint16_t, 0); else if (mode == 1) DQO_EMIT(uint16_t, 1); else DQO_EMIT(uint16_t, 2);
        } else {
            if (mode == 0) DQO_EMIT(uint32_t, 0); else if (mode == 1) DQO_EMIT(uint32_t, 1); else DQO_EMIT(uint32_t, 2);
        }
#undef DQO_EMIT
        DQO_LAUNCH_CHECK("scan + emit", debug, stream);
        if (mode == 1) {
            const int rcc = color_fork(1);
            if (rcc) return rcc;
        }
        if (mode != 2) stage_mark(stream, ST_DUPLICATE);
        int rc;
        // the scratch header was cleared at the start of the pass; the back-phase sort uses the pass slots behind the front
        // sort's (with more than two passes per sort -- images of >= 65535 tiles -- it clears for itself)
        const int sort_passes = radix_passes(sort_bits);
        const bool own_clear = mode == 2 && 2 * sort_passes > RS_MAX_PASSES;
        const int slot0 = (mode == 2 && !own_clear) ? sort_passes : 0;
        if (keys16)
            rc = radix_sort_pairs<uint16_t>((uint16_t *)ka, (uint16_t *)kb, va, vb, false, status + count_word,
                                            status + DQO_ST_OVERFLOW, n, sort_bits, sort_temp, stream, slot0, own_clear,
                                            capacity);
        else
            rc = radix_sort_pairs<uint32_t>((uint32_t *)ka, (uint32_t *)kb, va, vb, false, status + count_word,
                                            status + DQO_ST_OVERFLOW, n, sort_bits, sort_temp, stream, slot0, own_clear,
                                            capacity);
        if (rc) return rc;
        if (debug) DQO_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (mode != 2) stage_mark(stream, ST_TILE_SORT);
        const void *ks = in_a ? ka : kb;
        const unsigned rb = (unsigned)((n + 2047) / 2048);
        if (keys16)
            launch_pdl(tile_ranges_kernel<uint16_t>, dim3(rb), dim3(256), 0, stream, n, (const uint16_t *)ks, status, count_word, out_ranges);
        else
            launch_pdl(tile_ranges_kernel<uint32_t>, dim3(rb), dim3(256), 0, stream, n, (const uint32_t *)ks, status, count_word, out_ranges);
        DQO_LAUNCH_CHECK("tile ranges", debug, stream);
        if (mode != 2) stage_mark(stream, ST_RANGES);
        return DQO_OK;
    };

    // The list of non-empty tiles (rasterizer_impl.cu:348-365) is an output only -- the blend reads the ranges -- so
    // its single-block kernel runs on the side stream beside the final blend instead of in front of it.
    auto compact_fork = [&](const uint2 *rb) -> int {
        if (!tile_indices) return DQO_OK; // internal callers that never read the tile list (the fused mapping step)
        cudaStream_t cs = stream;
        if (fj) {
            DQO_CUDA_CHECK(cudaEventRecord(fj->ev[2], stream));
            DQO_CUDA_CHECK(cudaStreamWaitEvent(fj->side, fj->ev[2], 0));
            cs = fj->side;
        }
        launch_pdl(compact_tiles_kernel, dim3(1), dim3(1024), 0, cs, T, ranges, rb, tile_indices, status);
        DQO_LAUNCH_CHECK("compact tiles", debug, stream);
        if (fj) DQO_CUDA_CHECK(cudaEventRecord(fj->ev[3], cs));
        stage_mark(stream, ST_COMPACT);
        return DQO_OK;
    };
    auto compact_join = [&]() -> int {
        if (fj && tile_indices) DQO_CUDA_CHECK(cudaStreamWaitEvent(stream, fj->ev[3], 0));
        return DQO_OK;
    };

    if (!two_phase) {
        if (P > 0) {
            const int rc = bin_phase(0, capacity, 0, d_mask_bits, DQO_ST_NUM_RENDERED, ranges);
            if (rc) return rc;
        }
        int rc = compact_fork(nullptr);
        if (rc) return rc;
        if (ra.n_touched) launch_pdl(render_forward_kernel<0, true>, dim3(T), dim3(256), 0, stream, ra);
        else launch_pdl(render_forward_kernel<0, false>, dim3(T), dim3(256), 0, stream, ra);
        DQO_LAUNCH_CHECK("render forward", debug, stream);
        if ((rc = compact_join())) return rc;
        stage_mark(stream, ST_RENDER_FWD);
        nvtx_pop();
        return DQO_OK;
    }

    // two-phase: nearest Gaussians first, the rest only into the tiles that are still unfinished
    uint32_t *mask_bits_b = (uint32_t *)(img + IL.mask_bits_b);
    int rc = bin_phase(1, front, 0, d_mask_bits, DQO_ST_R_FRONT, ranges);
    if (rc) return rc;
    if ((rc = color_join(1))) return rc;
    if (ra.n_touched) launch_pdl(render_forward_kernel<1, true>, dim3(T), dim3(256), 0, stream, ra);
    else launch_pdl(render_forward_kernel<1, false>, dim3(T), dim3(256), 0, stream, ra);
    DQO_LAUNCH_CHECK("render forward (front)", debug, stream);
    stage_mark(stream, ST_RENDER_FRONT);
    launch_pdl(mask_unfinished_kernel, dim3(1), dim3(1024), 0, stream, IL.tiles_x, IL.tiles_y, IL.mask_words, d_mask_bits, ra.unfinished,
                                                   mask_bits_b, (uint32_t *)(img + IL.row_any_b), (uint32_t *)(img + IL.sat_b));
    DQO_LAUNCH_CHECK("unfinished mask", debug, stream);
    rc = bin_phase(2, back, front, mask_bits_b, DQO_ST_R_BACK, ranges_b);
    if (rc) return rc;
    stage_mark(stream, ST_BACK_BIN);
    if ((rc = compact_fork(ranges_b))) return rc;
    if ((rc = color_join(2))) return rc;
    if (ra.n_touched) launch_pdl(render_forward_kernel<2, true>, dim3(T), dim3(256), 0, stream, ra);
    else launch_pdl(render_forward_kernel<2, false>, dim3(T), dim3(256), 0, stream, ra);
    DQO_LAUNCH_CHECK("render forward (back)", debug, stream);
    if ((rc = compact_join())) return rc;
    stage_mark(stream, ST_RENDER_FWD);
    nvtx_pop();
    return DQO_OK;
}

extern "C" int dqo_rast_blend_extra(const dqo_rast_settings *s, const float *background, const float *colors,
                                    const void *geom_buffer, const void *binning_buffer, int64_t capacity,
                                    const void *image_buffer, const int32_t *status, float *out_color, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!s || s->W <= 0 || s->H <= 0 || s->P < 0 || !background || !image_buffer || !status || !out_color ||
        (s->P > 0 && (!colors || !geom_buffer || !binning_buffer))) {
        set_error("dqo_rast_blend_extra: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    ImgLayout IL;
    make_img_layout(s->W, s->H, &IL);
    const char *img = (const char *)image_buffer;
    ExtraArgs a;
    a.W = s->W; a.H = s->H; a.grid_x = IL.tiles_x; a.opaque_thr = s->opaque_threshold; a.T_thr = s->T_threshold;
    a.ranges = (const uint2 *)(img + IL.ranges);
    a.ranges_b = nullptr; a.point_list = nullptr; a.point_list_b = nullptr; a.rec = nullptr;
    if (s->P > 0) {
        GeomLayout GL;
        BinLayout BL;
        if (make_geom_layout(s->P, &GL) || make_bin_layout(capacity, &BL)) return DQO_ERR_WORKSPACE;
        a.rec = (const float4 *)((const char *)geom_buffer + GL.rec);
        a.point_list = (const uint32_t *)((const char *)binning_buffer + bin_point_list(BL, IL.T));
        if (s->front_instances > 0) {
            a.ranges_b = (const uint2 *)(img + IL.ranges_b);
            a.point_list_b = a.point_list + s->front_instances;
        }
    }
    a.colors = colors; a.bg = background; a.out_color = out_color;
    launch_pdl(blend_extra_kernel, dim3(IL.T), dim3(256), 0, stream, a);
    DQO_LAUNCH_CHECK("blend extra colours", s->debug, stream);
    return DQO_OK;
}

extern "C" int dqo_rast_export_state(const dqo_rast_settings *s, const void *geom_buffer, const void *binning_buffer,
                                     int64_t capacity, const void *image_buffer, const int32_t *status,
                                     uint64_t *sorted_keys, uint32_t *point_list, uint32_t *ranges_out,
                                     uint32_t *n_contrib, float *final_T, float *means2D, float *depths,
                                     float *conic_opacity, float *rgb, uint32_t *tiles_touched, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!s || !image_buffer || !status) return DQO_ERR_INVALID_ARG;
    if (s->front_instances > 0 && (sorted_keys || point_list || ranges_out)) {
        set_error("dqo_rast_export_state: the full sorted instance list only exists in single-phase mode (front_instances = 0)");
        return DQO_ERR_INVALID_ARG;
    }
    ImgLayout IL;
    make_img_layout(s->W, s->H, &IL);
    const char *img = (const char *)image_buffer;
    const int P = s->P;
    if (ranges_out)
        DQO_CUDA_CHECK(cudaMemcpyAsync(ranges_out, img + IL.ranges, (size_t)IL.T * 8, cudaMemcpyDeviceToDevice, stream));
    if (n_contrib || final_T)
        launch_pdl(export_pixels_kernel, dim3(IL.T), dim3(256), 0, stream, s->W, s->H, IL.tiles_x, (const uint32_t *)(img + IL.n_contrib),
                                                       (const float *)(img + IL.final_T),
                                                       (const uint2 *)(img + IL.ranges),
                                                       s->front_instances > 0 ? (const uint2 *)(img + IL.ranges_b) : nullptr,
                                                       n_contrib, final_T);
    if (P > 0) {
        GeomLayout GL;
        BinLayout BL;
        if (make_geom_layout(P, &GL) || make_bin_layout(capacity, &BL)) return DQO_ERR_WORKSPACE;
        const char *geom = (const char *)geom_buffer;
        const char *bin = (const char *)binning_buffer;
        if (sorted_keys || point_list) {
            const unsigned nb = (unsigned)((capacity + 255) / 256);
            if (IL.T < 65535)
                launch_pdl(export_instances_kernel<uint16_t>, dim3(nb), dim3(256), 0, stream, 
                    capacity, status, (const uint16_t *)(bin + bin_sorted_keys(BL, IL.T)), (const uint32_t *)(bin + bin_point_list(BL, IL.T)),
                    (const float *)(geom + GL.depth), sorted_keys, point_list);
            else
                launch_pdl(export_instances_kernel<uint32_t>, dim3(nb), dim3(256), 0, stream, 
                    capacity, status, (const uint32_t *)(bin + bin_sorted_keys(BL, IL.T)), (const uint32_t *)(bin + bin_point_list(BL, IL.T)),
                    (const float *)(geom + GL.depth), sorted_keys, point_list);
        }
        if (means2D || depths || conic_opacity || rgb || tiles_touched) {
            launch_pdl(export_gauss_kernel, dim3((P + 255) / 256), dim3(256), 0, stream, 
                P, (const float4 *)(geom + GL.rec), (const uint32_t *)(geom + GL.tiles),
                (const uint8_t *)(geom + GL.clamped), (const float *)(geom + GL.depth), means2D, depths, conic_opacity, rgb,
                tiles_touched);
        }
    }
    DQO_LAUNCH_CHECK("export", 0, stream);
    return DQO_OK;
}
