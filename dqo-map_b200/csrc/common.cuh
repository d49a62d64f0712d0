// Shared declarations for the B200-native DQO-MAP hot path (sm_100a only).
//
// Arithmetic policy: every value that feeds an integer decision of the reference (depth sort key, pixel
// centre, radius, tile rectangle, alpha / transmittance thresholds) is computed with explicit
// __fmul_rn/__fadd_rn/__fmaf_rn in the exact operation order that nvcc 12.9 emits for the reference
// sources on sm_100a (decoded from SASS, see DESIGN.md "Bit-exactness").  The compiler is therefore not
// free to re-contract them and oracle/ can restate them with fmaf() on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/dqo_b200.h"

#define DQO_TILE 16
#define DQO_TILE_PIX 256
#define DQO_ABI_VERSION 14

namespace dqo {

void set_error(const char *fmt, ...);
void note_launch(int n = 1);                      // counts this library's own kernel launches (dqo_launch_count)
void stage_mark(cudaStream_t stream, int stage);  // records a CUDA event when stage profiling is enabled (per thread)
void nvtx_push(const char *name);                 // NVTX range around a C-ABI entry point
void nvtx_pop();
// Helper stream + fork / join events owned by (current device, caller's stream); created on the first call on that
// stream, then reused: no stream / event creation on the hot path, no sharing between callers on different streams.
struct ForkJoin {
    cudaStream_t side = nullptr;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
ForkJoin *fork_join(cudaStream_t caller); // nullptr: run unforked on the caller's stream

// stage ids of dqo_profile_read()
enum {
    ST_BEGIN_FWD = 0, ST_PREPROCESS, ST_DEPTH_SORT, ST_SCAN, ST_DUPLICATE, ST_TILE_SORT, ST_RANGES,
    ST_RENDER_FRONT, ST_BACK_BIN, // two-phase binning only: front blend, then count/scan/duplicate/sort/ranges of the back phase
    ST_COMPACT, ST_RENDER_FWD, ST_BEGIN_BWD, ST_RENDER_BWD, ST_GAUSS_BWD, ST_COUNT
};

#define DQO_CUDA_CHECK(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            dqo::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return (int)_e;                                                                   \
        }                                                                                     \
    } while (0)

#define DQO_LAUNCH_CHECK(name, debug, stream)                                                 \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        dqo::note_launch();                                                                   \
        if (_e == cudaSuccess && (debug)) _e = cudaStreamSynchronize(stream);                 \
        if (_e != cudaSuccess) {                                                              \
            dqo::set_error("kernel %s failed: %s", name, cudaGetErrorString(_e));             \
            return (int)_e;                                                                   \
        }                                                                                     \
    } while (0)

// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// The step is a chain of ~45 short kernels on one stream; with plain stream order every kernel boundary costs the drain
// of the previous grid plus the launch / block-scheduling latency of the next.  Kernels that start with pdl_enter() and
// are launched through launch_pdl() let the next grid become resident while the previous one is finishing: the
// dependent's blocks are scheduled as soon as every block of the predecessor has STARTED (launch_dependents is the first
// instruction) and then block in griddepcontrol.wait until the predecessor has COMPLETED and its writes are visible, so
// no kernel touches memory earlier than under stream order.  Rule: launch_pdl() only for kernels that call pdl_enter()
// before their first global-memory access (a kernel without the wait could finish before its predecessor and break the
// transitive ordering of the chain).  Memsets / event waits between kernels simply fall back to full serialisation.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled(); // api.cu: on unless the environment says DQO_PDL=0 (A/B measurements)
// Whether the launches of the current C-ABI call (this thread) use the programmatic attribute.  A waiting dependent grid
// occupies registers / shared memory on the SMs; for one large scene on one stream that is free, but when many small
// scenes are mapped concurrently on several streams (the object-sharded workload) it takes the slots the other streams'
// kernels need (measured: 6100 -> 3700 objects*iters/s).  So every entry point declares its problem size first.
#define DQO_PDL_MIN_GAUSSIANS 500000
void pdl_scope(long long n_gaussians);
bool pdl_active();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = pdl_active() ? 1 : 0;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- exact-op helpers -------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fadd_rn(a, -b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float frcp(float a) { return __frcp_rn(a); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

// row `c` of a column-major 4x4 applied to (x,y,z,1): m[c]*x + m[c+4]*y + m[c+8]*z (+ m[c+12])
// reference order (auxiliary.h:59-77 as compiled): FMUL(y), FFMA(x), FFMA(z), FADD(w)
__device__ __forceinline__ float xform_row(const float *__restrict__ m, int c, float x, float y, float z) {
    float t = fmul(y, m[c + 4]);
    t = ffma(x, m[c], t);
    t = ffma(z, m[c + 8], t);
    return fadd(t, m[c + 12]);
}
__device__ __forceinline__ float xform_row3(const float *__restrict__ m, int c, float x, float y, float z) {
    float t = fmul(y, m[c + 4]);
    t = ffma(x, m[c], t);
    return ffma(z, m[c + 8], t);
}

// a0*b0 + a1*b1 + a2*b2 as the reference compiles it: FMUL(term1), FFMA(term0), FFMA(term2)
__device__ __forceinline__ float dot3_ref(float a0, float b0, float a1, float b1, float a2, float b2) {
    float t = fmul(a1, b1);
    t = ffma(a0, b0, t);
    return ffma(a2, b2, t);
}

// Quaternion (r,x,y,z), NOT normalised (forward.cu:211).  Entries of the GLM matrix
// R = mat3(col0 | col1 | col2) of forward.cu:218-221, rounded exactly as compiled.
struct QuatMat {
    float c0[3], c1[3], c2[3]; // GLM columns: R[col][row]
};
__device__ __forceinline__ QuatMat quat_to_glm(float r, float x, float y, float z) {
    const float xz = fmul(x, z), rx = fmul(r, x), rz = fmul(r, z);
    const float yy = fmul(y, y), zz = fmul(z, z);
    const float xz_p_ry = ffma(r, y, xz), xz_m_ry = ffma(-r, y, xz);
    const float yz_m_rx = ffma(y, z, -rx), yz_p_rx = ffma(y, z, rx);
    const float xy_m_rz = ffma(x, y, -rz), xy_p_rz = ffma(x, y, rz);
    float s0 = fadd(yy, zz), s1 = ffma(x, x, zz), s2 = ffma(x, x, yy);
    QuatMat q;
    q.c0[0] = fadd(-fadd(s0, s0), 1.f);
    q.c0[1] = fadd(xy_m_rz, xy_m_rz);
    q.c0[2] = fadd(xz_p_ry, xz_p_ry);
    q.c1[0] = fadd(xy_p_rz, xy_p_rz);
    q.c1[1] = fadd(-fadd(s1, s1), 1.f);
    q.c1[2] = fadd(yz_m_rx, yz_m_rx);
    q.c2[0] = fadd(xz_m_ry, xz_m_ry);
    q.c2[1] = fadd(yz_p_rx, yz_p_rx);
    q.c2[2] = fadd(-fadd(s2, s2), 1.f);
    return q;
}

// 3D covariance (forward.cu:202-235): M = S*R, Sigma = M^T M, upper triangle.
__device__ __forceinline__ void cov3d_from_scale_rot(float sx, float sy, float sz, float mod, float r, float x,
                                                     float y, float z, float *cov) {
    const QuatMat q = quat_to_glm(r, x, y, z);
    const float s[3] = {fmul(mod, sx), fmul(mod, sy), fmul(mod, sz)};
    float M0[3], M1[3], M2[3]; // M[col][row] = s[row] * R[col][row]
#pragma unroll
    for (int k = 0; k < 3; k++) {
        M0[k] = fmul(s[k], q.c0[k]);
        M1[k] = fmul(s[k], q.c1[k]);
        M2[k] = fmul(s[k], q.c2[k]);
    }
    cov[0] = dot3_ref(M0[0], M0[0], M0[1], M0[1], M0[2], M0[2]);
    cov[1] = dot3_ref(M1[0], M0[0], M1[1], M0[1], M1[2], M0[2]);
    cov[2] = dot3_ref(M2[0], M0[0], M2[1], M0[1], M2[2], M0[2]);
    cov[3] = dot3_ref(M1[0], M1[0], M1[1], M1[1], M1[2], M1[2]);
    cov[4] = dot3_ref(M2[0], M1[0], M2[1], M1[1], M2[2], M1[2]);
    cov[5] = dot3_ref(M2[0], M2[0], M2[1], M2[1], M2[2], M2[2]);
}

// argMin / argMax with the reference's tie rules (forward.cu:20-52)
__device__ __forceinline__ int arg_min3(float a, float b, float c) {
    if (a <= b && a <= c) return 0;
    if (b <= a && b <= c) return 1;
    return 2;
}
__device__ __forceinline__ int arg_max3(float a, float b, float c) {
    if (a >= b && a >= c) return 0;
    if (b >= a && b >= c) return 1;
    return 2;
}

// Pixel ray (forward.cu:92-100)
__device__ __forceinline__ float3 pixel_ray(uint32_t px, uint32_t py, float fx, float fy, float cx, float cy) {
    float rx = fdiv(fsub((float)px, cx), fx);
    float ry = fdiv(fsub((float)py, cy), fy);
    float n2 = fadd(ffma(rx, rx, fmul(ry, ry)), 1.0f);
    float inv = frcp(fsqrt(n2));
    return make_float3(fmul(rx, inv), fmul(ry, inv), inv); // 1.0 * inv is exact
}

__host__ __device__ inline uint32_t higher_msb(uint32_t n) { // rasterizer_impl.cu:35-50 == bit length
    uint32_t b = 0;
    while (n >> b && b < 32) b++;
    return b;
}

// ---- private workspace layouts ------------------------------------------------------------------
// geometry buffer: per-Gaussian splat records + depth-sort scratch
// A second colour set blended over the lists of the main render (semantic image of loss_update): inputs and outputs of
// its backward pass (rast_backward_impl)
struct ExtraBlendGrad {
    const float *colors;  // [P,3]
    const float *dL_dpix; // [3,H,W] gradient of the loss w.r.t. the extra image
    double *cacc;         // f64[4P] colour-gradient accumulators, zero on entry, left zero
    float *dL_dcolors;    // [P,3] out
    int only;             // 1: the backward of the extra image alone (no main image, no depth, no SH gradients)
};
struct GeomLayout {
    size_t rec;        // float4[3P]: {x,y,conic.x,conic.y} {conic.z,opacity,power_reject,extent.y} {r,g,b,extent.x}
    size_t depth;      // f32[P] view-space depth (forward.cu:339)
    size_t depth_key;  // u32[P] float bits of view depth, 0xFFFFFFFF when the Gaussian is outside the frustum
    size_t depth_key2; // u32[P] sort ping-pong buffer
    size_t ids;        // u32[P] sort ping-pong buffer (values)
    size_t order;      // u32[P] Gaussian ids in (depth, id) order
    size_t tiles;      // u32[P] tiles_touched
    size_t offsets;    // u32[P] inclusive scan of tiles_touched in depth-rank order
    size_t rect;       // uint2[P] {min.x | max.x<<16, min.y | max.y<<16}
    size_t clamped;    // u8[P] bit c = colour channel c clamped (forward.cu:151-153)
    size_t gacc;       // f64[16P] gradient accumulators of the backward blend (see DQO_GACC_FLOATS)
    size_t touched;    // u8[P] set by the backward blend when it adds to a Gaussian's record, cleared by the consumer
    size_t out_nz;     // u8[P] geom_clean == 2: 1 where the previous backward wrote non-zero gradient rows
    size_t tiles_b;    // u32[P] two-phase: unfinished tiles in the rectangle of every rank the front phase left out
    size_t sums;       // per phase: u32[emit_blocks] totals of 256 ranks + u32[emit_groups] totals of 64 such blocks
    size_t sums_stride;
    int emit_blocks, emit_groups;
    size_t sort_temp;  // histograms / look-back words of the depth sort (sort.cuh: SortTemp)
    size_t total;
};
// binning buffer: instance lists.  The radix sort ping-pongs between the (a) and (b) arrays; where the sorted list ends
// up depends only on the number of digit passes, i.e. on the tile count of the image (bin_sorted_in_a).
struct BinLayout {
    size_t keys_a, keys_b; // tile id per instance: u16[C] when the image has < 65535 tiles, else u32[C]
    size_t vals_a, vals_b; // u32[C] Gaussian id per instance; the sorted one == reference point_list
    size_t sort_temp;
    size_t total;
};
// image buffer: per-tile ranges and tile-major per-pixel state for the backward pass
struct ImgLayout {
    size_t ranges;     // uint2[T]
    size_t n_contrib;  // u32[T*256]
    size_t final_T;    // f32[T*256]
    size_t hit_geo;    // f32[6][T*256]: hit_normal_c.xyz, hit_point_c.xyz (forward.cu:807-808)
    size_t mask_bits;  // u32[tiles_y][mask_words]: tile_mask != 0 as a bitmap
    size_t ranges_b;   // uint2[T] two-phase: ranges of the back-phase lists (relative to the back region)
    size_t unfinished; // i32[T] two-phase: 1 when the front phase left some pixel of the tile unterminated
    size_t mask_bits_b; // u32[tiles_y][mask_words]: mask_bits & unfinished
    size_t row_any_b;  // u32[ceil(tiles_y/32)]: bit y set <=> row y of mask_bits_b has any bit set
    size_t state;      // f32[4][T*256] two-phase: running T (-1 = pixel terminated) and colour accumulators
    size_t sat, sat_b; // u32[(tiles_y+1)*(tiles_x+1)] summed-area tables of mask_bits / mask_bits_b
    int mask_words;
    size_t total;
    int tiles_x, tiles_y, T;
};

// The 8 sub-blocks (8x4 pixels, one per warp) of a 16x16 tile that a splat can reach.  Bit b set <=> some pixel
// centre of sub-block b = (bx = (b&1)*8, by = (b>>1)*4) may satisfy power >= power_reject, i.e. lies inside the
// ellipse d^T Q d <= tau (Q = conic).  Two conservative tests: the ellipse's axis-aligned extent (ex, ey), then the
// exact minimum of the quadratic form over the sub-block rectangle (interior, else the best point of the 4 edges).
// tau is recovered from ex (which already carries the rounding-noise inflation and margins of the preprocess);
// any NaN / inf in the chain keeps the sub-block.
__device__ __forceinline__ float edge_min_qf(float A, float B, float C, float u, float vlo, float vhi, float invC) {
    // min over v in [vlo, vhi] of A u^2 + 2 B u v + C v^2
    const float v = fminf(fmaxf(-B * u * invC, vlo), vhi);
    return A * u * u + 2.f * B * u * v + C * v * v;
}
__device__ __forceinline__ unsigned subblock_mask(float mx, float my, float A, float B, float C, float ex, float ey,
                                                  float tile_px, float tile_py) {
    unsigned m = 0;
    const bool x0 = (mx + ex >= tile_px) && (mx - ex <= tile_px + 7.f);
    const bool x1 = (mx + ex >= tile_px + 8.f) && (mx - ex <= tile_px + 15.f);
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const float y0 = tile_py + 4.f * r;
        const bool yr = (my + ey >= y0) && (my - ey <= y0 + 3.f);
        if (yr && x0) m |= 1u << (2 * r);
        if (yr && x1) m |= 1u << (2 * r + 1);
    }
    if (m == 0) return 0;
    const float tau = ex * ex * (A * C - B * B) / C * 1.0001f;
    if (!(tau < 3.0e38f)) return m; // inf / NaN: extent test only
    const float invA = 1.f / A, invC = 1.f / C;
    // The form is convex with its minimum at the splat centre, so over a rectangle that does not contain the centre the
    // minimum lies on an edge that faces it: at most one vertical and one horizontal edge need the 1-D minimisation.
    // Only the candidate sub-blocks are visited (a warp iterates as often as its lane with the most candidates).
    unsigned keep = 0;
    for (unsigned todo = m; todo; todo &= todo - 1) {
        const int b = __ffs(todo) - 1;
        const float dxl = tile_px + 8.f * (b & 1) - mx, dxh = dxl + 7.f;
        const float dyl = tile_py + 4.f * (b >> 1) - my, dyh = dyl + 3.f;
        const bool in_x = dxl <= 0.f && dxh >= 0.f, in_y = dyl <= 0.f && dyh >= 0.f;
        float f = (in_x && in_y) ? 0.f : 3.4e38f;
        const float fx = edge_min_qf(A, B, C, dxl > 0.f ? dxl : dxh, dyl, dyh, invC);
        const float fy = edge_min_qf(C, B, A, dyl > 0.f ? dyl : dyh, dxl, dxh, invA);
        if (!in_x) f = fminf(f, fx);
        if (!in_y) f = fminf(f, fy);
        if (!(f > tau)) keep |= 1u << b;
    }
    return keep;
}

int make_geom_layout(int P, GeomLayout *L);
int make_bin_layout(int64_t C, BinLayout *L);
void make_img_layout(int W, int H, ImgLayout *L);

// number of key bits the tile sort looks at for an image of T tiles (rasterizer_impl.cu:327: getHigherMsb(tiles))
inline int tile_sort_bits(int T) {
    const int bit = (int)higher_msb((uint32_t)T);
    return (T < 65535) ? (bit < 16 ? bit : 16) : bit;
}
// true: after the tile sort the sorted keys / point list live in keys_a / vals_a, false: in keys_b / vals_b
inline bool bin_sorted_in_a(int T) { return ((tile_sort_bits(T) < 1 ? 1 : tile_sort_bits(T)) + 7) / 8 % 2 == 0; }
inline size_t bin_point_list(const BinLayout &BL, int T) { return bin_sorted_in_a(T) ? BL.vals_a : BL.vals_b; }
inline size_t bin_sorted_keys(const BinLayout &BL, int T) { return bin_sorted_in_a(T) ? BL.keys_a : BL.keys_b; }

// per-Gaussian gradient accumulator written by the backward blend: 16 DOUBLES (128 B).  Every warp adds its fp32 partial
// sums with fp64 atomics, so the order in which warps and tiles arrive no longer shows in the result (the reference's
// float atomics make its gradients differ run to run by up to 1e-3 for ill-conditioned Gaussians)
// {dmean2D.x, dmean2D.y, dconic.x, dconic.y | dconic.w, dopacity, dcolor.r, dcolor.g | dcolor.b, dmean3D.xyz | drot.rxyz}
#define DQO_GACC_FLOATS 16

} // namespace dqo
