// Backward rasterization: reverse blend with warp-level pre-reduction, then one fused per-Gaussian kernel.
//
// Replaces CudaRasterizer::Rasterizer::backward (reference RAST/cuda_rasterizer/rasterizer_impl.cu:445-564):
//   BACKWARD::renderCUDA_flat  backward.cu:808-1066  -> render_backward_kernel
//   computeCov2DCUDA           backward.cu:273-422   \
//   BACKWARD::preprocessCUDA   backward.cu:492-548    > gaussian_backward_kernel (one pass, no dL_dcov3D round trip)
//   computeColorFromSH (bwd)   backward.cu:152-268   /
//   computeCov3D (bwd)         backward.cu:426-487  /
// The reference issues 9 global float atomics per contributing (pixel, Gaussian) pair.  Here the 9 partial
// gradients of a warp are summed with a 12-shuffle recursive-halving reduction and written with one
// 9-lane RED into a 64-byte per-Gaussian accumulator record; the per-Gaussian kernel then produces every
// output tensor densely (zeros included), so no gradient tensor has to be pre-zeroed by the caller.
#include "common.cuh"

namespace dqo {

struct RenderBwdArgs {
    int W, H, grid_x;
    float fx, fy, cx, cy;
    float depth_thr, normal_thr;
    const uint2 *ranges;
    const uint32_t *point_list;
    const uint2 *ranges_b;          // two-phase binning: back lists (nullptr in single-phase mode)
    const uint32_t *point_list_b;
    const float4 *rec;
    const float *view, *means3D, *scales, *rotations, *bg;
    const uint32_t *n_contrib;
    const float *final_T;
    const float *hit_geo;
    size_t plane;
    const float *dL_dpix, *dL_ddepth;
    const int *hit_image;
    double *gacc;
    uint8_t *touched; // touched[id] = 1 whenever a record receives a contribution (plain store, every writer stores 1)
    // EXTRA instantiation (a second colour set blended over the same lists, e.g. the semantic image of loss_update): the
    // colours come from extra_colors [P,3] instead of the splat record, their gradient goes to cacc (f64[4] per Gaussian),
    // the geometry gradients into the same records as the main pass (gradients of the two images add), no depth path
    const float *extra_colors;
    double *cacc;
};

// Sum 9 per-lane values over the warp; on return lanes with (lane & 1) == 0 whose slot index < 9 hold the
// total of value `slot` (see slot_of_lane).  12 shuffles instead of 45.
__device__ __forceinline__ float warp_reduce9(const float v[9], int lane) {
    const unsigned FULL = 0xFFFFFFFFu;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float w[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const float lo = v[k], hi = (k + 5 < 9) ? v[k + 5] : 0.f;
        const float send = b4 ? lo : hi;
        const float keep = b4 ? hi : lo;
        w[k] = keep + __shfl_xor_sync(FULL, send, 16);
    }
    float u[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float lo = w[k], hi = (k + 3 < 5) ? w[k + 3] : 0.f;
        const float send = b3 ? lo : hi;
        const float keep = b3 ? hi : lo;
        u[k] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    float t[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const float lo = u[k], hi = (k + 2 < 3) ? u[k + 2] : 0.f;
        const float send = b2 ? lo : hi;
        const float keep = b2 ? hi : lo;
        t[k] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    float r = (b1 ? t[1] : t[0]) + __shfl_xor_sync(FULL, b1 ? t[0] : t[1], 2);
    r += __shfl_xor_sync(FULL, r, 1);
    return r;
}
__device__ __forceinline__ int slot_of_lane(int lane) {
    // five-slot index after the first round, three-slot after the second, two-slot after the third
    const int s2 = (lane & 2) ? 1 : 0;
    const int s3 = ((lane & 4) ? 2 : 0) + s2;       // valid if < 3
    const int s5 = ((lane & 8) ? 3 : 0) + s3;       // valid if < 5
    const int s9 = ((lane & 16) ? 5 : 0) + s5;      // valid if < 9
    const bool ok = (s3 < 3) && (s5 < 5) && (s9 < 9) && !((lane & 4) && s2) && ((lane & 1) == 0);
    return ok ? s9 : -1;
}

// shared-memory accessors of the blend loop: the window address is computed once and kept opaque, so the compiler
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
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}

#ifndef RB_BATCH
#define RB_BATCH 256 // 80-byte records + mask + four u16 lists: 23 KB per block, eight blocks per SM
#endif
#ifndef RB_OCC
#define RB_OCC 8 // 64 registers: the loop is latency-bound (one dependent chain per warp), warps are what hides it
#endif
#define RB_THREADS 128 // 4 warps x (8x8 pixels); every thread owns the pixels (x, y) and (x, y + 4)

// depth gradient to the single hit Gaussian of a pixel (backward.cu:998-1065)
__device__ __noinline__ void bwd_depth_path(const float *scales, const float *rotations, const float *means3D,
                                            const float *view, const float *hit_geo, size_t plane, size_t sp, double *gacc, uint8_t *touched,
                                            int gid, float g, uint32_t pix_x, uint32_t pix_y, float fx, float fy, float cx,
                                            float cy, float depth_thr, float normal_thr) {
    const float3 ray = pixel_ray(pix_x, pix_y, fx, fy, cx, cy);
    const float sx = scales[3 * gid], sy = scales[3 * gid + 1], sz = scales[3 * gid + 2];
    const float scale_max = fmaxf(fmaxf(sx, sy), sz);
    const float ncx = hit_geo[sp], ncy = hit_geo[plane + sp], ncz = hit_geo[2 * plane + sp];
    const float hz = hit_geo[5 * plane + sp];
    const float *v = view;
    const float wx = means3D[3 * gid], wy = means3D[3 * gid + 1], wz = means3D[3 * gid + 2];
    const float pcx = xform_row(v, 0, wx, wy, wz), pcy = xform_row(v, 1, wx, wy, wz), pcz = xform_row(v, 2, wx, wy, wz);
    const float ndotr = dot3_ref(ncx, ray.x, ncy, ray.y, ncz, ray.z);
    const float angle_distance = fabsf(ndotr);
    const float depth_distance = fabsf(fsub(hz, pcz));
    double *acc = gacc + (size_t)gid * DQO_GACC_FLOATS;
    touched[gid] = 1;
    if (depth_distance <= fmul(depth_thr, scale_max) && angle_distance >= normal_thr) {
        const float nr = (float)((double)ndotr + 1e-8);
        const float inv_nr = 1.f / nr;
        const float inv_nr2 = inv_nr * inv_nr;
        const float np = ncx * pcx + ncy * pcy + ncz * pcz;
        const float dpx = ray.z * ncx * inv_nr, dpy = ray.z * ncy * inv_nr, dpz = ray.z * ncz * inv_nr;
        atomicAdd(&acc[9], (double)(g * (dpx * v[0] + dpy * v[1] + dpz * v[2])));
        atomicAdd(&acc[10], (double)(g * (dpx * v[4] + dpy * v[5] + dpz * v[6])));
        atomicAdd(&acc[11], (double)(g * (dpx * v[8] + dpy * v[9] + dpz * v[10])));
        const int axis = arg_min3(sx, sy, sz);
        const float n1c = ray.z * (nr * pcx - np * ray.x) * inv_nr2;
        const float n2c = ray.z * (nr * pcy - np * ray.y) * inv_nr2;
        const float n3c = ray.z * (nr * pcz - np * ray.z) * inv_nr2;
        const float n1w = n1c * v[0] + n2c * v[1] + n3c * v[2];
        const float n2w = n1c * v[4] + n2c * v[5] + n3c * v[6];
        const float n3w = n1c * v[8] + n2c * v[9] + n3c * v[10];
        const float4 q = reinterpret_cast<const float4 *>(rotations)[gid];
        const float q0 = q.x, q1 = q.y, q2 = q.z, q3 = q.w;
        float d0[3], d1[3], d2[3], d3[3]; // d normal / d q_k (backward.cu:100-148)
        if (axis == 0) {
            d0[0] = 0; d0[1] = 2 * q3; d0[2] = -2 * q2;
            d1[0] = 0; d1[1] = 2 * q2; d1[2] = 2 * q3;
            d2[0] = -4 * q2; d2[1] = 2 * q1; d2[2] = -2 * q0;
            d3[0] = -4 * q3; d3[1] = 2 * q0; d3[2] = 2 * q1;
        } else if (axis == 1) {
            d0[0] = -2 * q3; d0[1] = 0; d0[2] = 2 * q1;
            d1[0] = 2 * q2; d1[1] = -4 * q1; d1[2] = 2 * q0;
            d2[0] = 2 * q1; d2[1] = 0; d2[2] = 2 * q3;
            d3[0] = -2 * q0; d3[1] = -4 * q3; d3[2] = 2 * q2;
        } else {
            d0[0] = 2 * q2; d0[1] = -2 * q1; d0[2] = 0;
            d1[0] = 2 * q3; d1[1] = -2 * q0; d1[2] = -4 * q1;
            d2[0] = 2 * q0; d2[1] = 2 * q3; d2[2] = -4 * q2;
            d3[0] = 2 * q1; d3[1] = 2 * q2; d3[2] = 0;
        }
        atomicAdd(&acc[12], (double)(g * (n1w * d0[0] + n2w * d0[1] + n3w * d0[2])));
        atomicAdd(&acc[13], (double)(g * (n1w * d1[0] + n2w * d1[1] + n3w * d1[2])));
        atomicAdd(&acc[14], (double)(g * (n1w * d2[0] + n2w * d2[1] + n3w * d2[2])));
        atomicAdd(&acc[15], (double)(g * (n1w * d3[0] + n2w * d3[1] + n3w * d3[2])));
    } else {
        atomicAdd(&acc[9], (double)(g * v[2]));
        atomicAdd(&acc[10], (double)(g * v[6]));
        atomicAdd(&acc[11], (double)(g * v[10]));
    }
}

// Reverse blend of one 16x16 tile.  Same per-warp culled lists as the forward kernel; a warp owns an 8x8 pixel block and
// every thread the two pixels (x, y) and (x, y + 4).
//
// The loop is bound by instruction issue, not by any pipe, so the two pixels of a thread are evaluated as the two
// halves of Blackwell's packed FP32 instructions (fma / mul / add .f32x2, __ffma2_rn & co.): one issue slot per pair of
// IEEE-rounded operations (measured: 8 FFMA + 16 integer ops per iteration take 0.67 ms, 4 FFMA2 + 16 integer ops
// 0.41 ms, tests/microbench/ffma2.cu).  Every per-lane result is the same correctly rounded value a scalar
// instruction gives.  The staged record keeps each field DUPLICATED ({v, v}) so that a 128-bit shared-memory load
// yields ready-made operand pairs; conic.y is stored negated (sign flips commute with rounding), which turns the
// subtractions of the power and of dG/ddelta into plain FMAs.  exp (FFMA.SAT / .RM sequence), min and the IEEE
// division stay scalar.
struct __align__(16) SplatB {
    float2 x, y, cx, ncy, cz, op, r, g, b; // duplicated fields: {v, v}
    float id_bits, pad;                    // 80 bytes
};
typedef float2 f2;
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 dup(float v) { return make_float2(v, v); }
__device__ __forceinline__ f2 lo_hi(float4 q, int h) { return h ? make_float2(q.z, q.w) : make_float2(q.x, q.y); }

template <bool EXTRA>
__global__ void __launch_bounds__(RB_THREADS, RB_OCC) render_backward_kernel(RenderBwdArgs a) {
    pdl_enter();
    // RB_BATCH entries are staged per round; most tiles need a single round (max n_contrib is a few hundred): the four
    // warps then walk their own lists without meeting at a barrier, where the fast ones would wait for the slowest.
    __shared__ SplatB s_sp[RB_BATCH];
    __shared__ uint8_t s_mask[RB_BATCH];
    // per-warp list entry: batch slot (9 bits) | reach bit of the upper pixel row block << 14 | of the lower << 15
    __shared__ uint16_t s_list[RB_THREADS / 32][RB_BATCH];
    __shared__ int s_max;

    const int tile = blockIdx.x;
    const uint2 range = a.ranges[tile];
    // the tile's list is the front list followed by the back list (two-phase binning); position p lives in the
    // front list when p < len_a
    const int len_a = (int)(range.y - range.x);
    uint32_t start_b = 0;
    bool has_b = false;
    if (a.ranges_b) {
        const uint2 rb = a.ranges_b[tile];
        start_b = rb.x;
        has_b = rb.x != rb.y;
    }
    if (len_a == 0 && !has_b) return;
    const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int lx = (warp & 1) * 8 + (lane & 7), ly0 = (warp >> 1) * 8 + (lane >> 3);
    const uint32_t pix_x = tile_x * DQO_TILE + lx;
    const size_t HW = (size_t)a.W * a.H;
    const float tile_px = (float)(tile_x * DQO_TILE), tile_py = (float)(tile_y * DQO_TILE);
    const int b_lo = 2 * (2 * (warp >> 1)) + (warp & 1), b_hi = b_lo + 2;
    const unsigned warp_bits = (1u << b_lo) | (1u << b_hi);

    // per-pixel state of the reverse traversal (backward.cu:881-900), pixel h in half h of every pair
    float Tv[2], Tfin[2], dl0[2], dl1[2], dl2[2], bgd[2];
    int last_c[2];
    int warp_max = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int ly = ly0 + 4 * h;
        const uint32_t py = tile_y * DQO_TILE + ly;
        const bool inside = pix_x < (uint32_t)a.W && py < (uint32_t)a.H;
        const size_t pid = (size_t)a.W * py + pix_x, sp = (size_t)tile * 256 + ly * 16 + lx;
        Tfin[h] = inside ? a.final_T[sp] : 0.f;
        Tv[h] = Tfin[h];
        last_c[h] = inside ? (int)a.n_contrib[sp] : 0;
        dl0[h] = dl1[h] = dl2[h] = 0.f;
        if (inside) {
            dl0[h] = a.dL_dpix[pid];
            dl1[h] = a.dL_dpix[HW + pid];
            dl2[h] = a.dL_dpix[2 * HW + pid];
        }
        bgd[h] = a.bg[0] * dl0[h] + a.bg[1] * dl1[h] + a.bg[2] * dl2[h];
        warp_max = max(warp_max, last_c[h]);
    }
    f2 T2 = make_float2(Tv[0], Tv[1]);
    const f2 dLp0 = make_float2(dl0[0], dl0[1]), dLp1 = make_float2(dl1[0], dl1[1]), dLp2 = make_float2(dl2[0], dl2[1]);
    f2 acc0 = dup(0.f), acc1 = dup(0.f), acc2 = dup(0.f);
    const bool any_bg = bgd[0] != 0.f || bgd[1] != 0.f; // background term (backward.cu:975), 0 for bg = 0
    const f2 npx = dup(-(float)pix_x);
    const f2 npy = make_float2(-(float)(tile_y * DQO_TILE + ly0), -(float)(tile_y * DQO_TILE + ly0 + 4));
    if (tid == 0) s_max = 0;
    __syncthreads();
    // entries at positions >= warp_max contribute to no pixel of this warp, >= max_c to no pixel of the tile
    for (int o = 16; o > 0; o >>= 1) warp_max = max(warp_max, __shfl_xor_sync(0xFFFFFFFFu, warp_max, o));
    if (lane == 0) atomicMax(&s_max, warp_max);
    __syncthreads();
    const int max_c = s_max;
    const f2 ddel = make_float2(0.5f * a.W, 0.5f * a.H); // {ddelx_dx, ddely_dy}
    const int my_slot = slot_of_lane(lane);
    const f2 one2 = dup(1.f), mone2 = dup(-1.f), mhalf2 = dup(-0.5f);

    const uint32_t sp_base = smem_addr(s_sp), list_base = smem_addr(&s_list[warp][0]);
    const int rounds = (max_c + RB_BATCH - 1) / RB_BATCH;
    for (int i = 0; i < rounds; i++) {
        __syncthreads();
        const int n = min(RB_BATCH, max_c - i * RB_BATCH);
#pragma unroll
        for (int e = 0; e < RB_BATCH / RB_THREADS; e++) {
            const int slot = e * RB_THREADS + tid;
            if (slot < n) {
                const int pos = max_c - 1 - (i * RB_BATCH + slot);
                const int id = (int)(pos < len_a ? a.point_list[range.x + pos] : a.point_list_b[start_b + (pos - len_a)]);
                const float4 r0 = __ldg(&a.rec[3 * (size_t)id]);
                const float4 r1 = __ldg(&a.rec[3 * (size_t)id + 1]);
                float4 r2 = __ldg(&a.rec[3 * (size_t)id + 2]);
                if constexpr (EXTRA) {
                    r2.x = __ldg(&a.extra_colors[3 * (size_t)id]);
                    r2.y = __ldg(&a.extra_colors[3 * (size_t)id + 1]);
                    r2.z = __ldg(&a.extra_colors[3 * (size_t)id + 2]);
                }
                float4 *dst = reinterpret_cast<float4 *>(&s_sp[slot]);
                dst[0] = make_float4(r0.x, r0.x, r0.y, r0.y);
                dst[1] = make_float4(r0.z, r0.z, -r0.w, -r0.w);
                dst[2] = make_float4(r1.x, r1.x, r1.y, r1.y);
                dst[3] = make_float4(r2.x, r2.x, r2.y, r2.y);
                dst[4] = make_float4(r2.z, r2.z, __int_as_float(id), 0.f);
                s_mask[slot] = (uint8_t)subblock_mask(r0.x, r0.y, r0.z, r0.w, r1.x, r2.w, r1.w, tile_px, tile_py);
            }
        }
        __syncthreads();
        // this warp's entries of the batch, in processing (back-to-front) order
        int cnt = 0;
        for (int b = 0; b < n; b += 32) {
            const int j = b + lane;
            const int posj = max_c - 1 - (i * RB_BATCH + j);
            const unsigned mk = (j < n) ? s_mask[j] : 0u;
            const bool m = (posj < warp_max) && (mk & warp_bits);
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, m);
            if (m) s_list[warp][cnt + __popc(bal & ((1u << lane) - 1))] =
                       (uint16_t)(j | (((mk >> b_lo) & 1u) << 14) | (((mk >> b_hi) & 1u) << 15));
            cnt += __popc(bal);
        }
        __syncwarp();
        const int pos_top = max_c - 1 - i * RB_BATCH; // list position of batch slot 0
        for (int k = 0; k < cnt; k++) {
            const uint32_t e = lds_u16(list_base + 2 * k);
            const int j = (int)(e & 0x3FFFu);
            const uint32_t rec = sp_base + (uint32_t)j * 80u;
            const float4 q0 = lds128(rec), q1 = lds128(rec + 16), q2 = lds128(rec + 32);
            const f2 x2 = lo_hi(q0, 0), y2 = lo_hi(q0, 1), cx2 = lo_hi(q1, 0), ncy2 = lo_hi(q1, 1), cz2 = lo_hi(q2, 0),
                     op2 = lo_hi(q2, 1);
            const int posj = pos_top - j;
            // power (backward.cu:931-934), both pixels at once, in the forward's operation order
            const f2 dx2 = add2(x2, npx), dy2 = add2(y2, npy);
            const f2 inner = fma2(dx2, mul2(dx2, cx2), mul2(dy2, mul2(dy2, cz2)));
            const f2 power = fma2(inner, mhalf2, mul2(dy2, mul2(dx2, ncy2)));
            const f2 G2 = make_float2(expf(power.x), expf(power.y));
            const f2 al = mul2(op2, G2);
            const f2 alpha = make_float2(fminf(0.99f, al.x), fminf(0.99f, al.y));
            const bool live0 = (e & 0x4000u) && (posj < last_c[0]) && !(power.x > 0.0f) && !(alpha.x < 1.0f / 255.0f);
            const bool live1 = (e & 0x8000u) && (posj < last_c[1]) && !(power.y > 0.0f) && !(alpha.y < 1.0f / 255.0f);
            if (!__any_sync(0xFFFFFFFFu, live0 || live1)) continue;
            const float4 q3 = lds128(rec + 32 + 16), q4 = lds128(rec + 64);
            const f2 cr = lo_hi(q3, 0), cg = lo_hi(q3, 1), cb = lo_hi(q4, 0);
            const f2 om = fma2(alpha, mone2, one2); // 1 - alpha
            // IEEE division as in the reference (backward.cu:947): the conic -> cov3D -> rotation chain downstream amplifies
            // a 1-ulp change of T to ~1e-3 in dL_drotations, so approximations here are not free
            const f2 Tn = make_float2(T2.x / om.x, T2.y / om.y);
            // accum_rec of the reference (backward.cu:953-961) is last_alpha * last_color + (1 - last_alpha) * accum_rec,
            // evaluated when the NEXT contributor is visited; the same expression is evaluated here one visit earlier, so
            // last_alpha / last_color need not be carried
            f2 dLa = mul2(fma2(acc0, mone2, cr), dLp0);
            dLa = fma2(fma2(acc1, mone2, cg), dLp1, dLa);
            dLa = fma2(fma2(acc2, mone2, cb), dLp2, dLa);
            dLa = mul2(dLa, Tn);
            const f2 n0 = fma2(om, acc0, mul2(alpha, cr)), n1 = fma2(om, acc1, mul2(alpha, cg)),
                     n2 = fma2(om, acc2, mul2(alpha, cb));
            if (any_bg) {
                dLa.x += (-Tfin[0] / om.x) * bgd[0];
                dLa.y += (-Tfin[1] / om.y) * bgd[1];
            }
            // a rejected pair keeps its pixel's state and contributes zeros: the three roots of the nine products below
            T2 = make_float2(live0 ? Tn.x : T2.x, live1 ? Tn.y : T2.y);
            acc0 = make_float2(live0 ? n0.x : acc0.x, live1 ? n0.y : acc0.y);
            acc1 = make_float2(live0 ? n1.x : acc1.x, live1 ? n1.y : acc1.y);
            acc2 = make_float2(live0 ? n2.x : acc2.x, live1 ? n2.y : acc2.y);
            dLa = make_float2(live0 ? dLa.x : 0.f, live1 ? dLa.y : 0.f);
            const f2 G = make_float2(live0 ? G2.x : 0.f, live1 ? G2.y : 0.f);
            const f2 aT = mul2(alpha, Tn);
            const f2 dch = make_float2(live0 ? aT.x : 0.f, live1 ? aT.y : 0.f); // dchannel_dcolor
            const f2 dL_dG = mul2(op2, dLa);
            const f2 gdx = mul2(G, dx2), gdy = mul2(G, dy2);
            // dG_ddelx = -gdx * conic.x - gdy * conic.y, dG_ddely = -gdy * conic.z - gdx * conic.y
            const f2 dGx = fma2(gdy, ncy2, mul2(mul2(gdx, mone2), cx2));
            const f2 dGy = fma2(gdx, ncy2, mul2(mul2(gdy, mone2), cz2));
            const f2 hx = mul2(mul2(gdx, mhalf2), dL_dG), hy = mul2(mul2(gdy, mhalf2), dL_dG);
            const f2 p0 = mul2(mul2(dL_dG, dGx), dup(ddel.x)), p1 = mul2(mul2(dL_dG, dGy), dup(ddel.y));
            const f2 p2 = mul2(hx, dx2), p3 = mul2(hx, dy2), p4 = mul2(hy, dy2), p5 = mul2(G, dLa);
            const f2 p6 = mul2(dch, dLp0), p7 = mul2(dch, dLp1), p8 = mul2(dch, dLp2);
            float v[9] = {p0.x + p0.y, p1.x + p1.y, p2.x + p2.y, p3.x + p3.y, p4.x + p4.y,
                          p5.x + p5.y, p6.x + p6.y, p7.x + p7.y, p8.x + p8.y};
            const float total = warp_reduce9(v, lane);
            if constexpr (EXTRA) {
                if (my_slot >= 6) atomicAdd(&a.cacc[(size_t)__float_as_int(q4.z) * 4 + (my_slot - 6)], (double)total);
                else if (my_slot >= 0) atomicAdd(&a.gacc[(size_t)__float_as_int(q4.z) * DQO_GACC_FLOATS + my_slot], (double)total);
            } else {
                if (my_slot >= 0) atomicAdd(&a.gacc[(size_t)__float_as_int(q4.z) * DQO_GACC_FLOATS + my_slot], (double)total);
            }
            if (lane == 0) a.touched[__float_as_int(q4.z)] = 1;
        }
    }
    if constexpr (!EXTRA) { // (the extra image carries no depth output)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int ly = ly0 + 4 * h;
            const uint32_t py = tile_y * DQO_TILE + ly;
            if (!(pix_x < (uint32_t)a.W && py < (uint32_t)a.H)) continue;
            const size_t pid = (size_t)a.W * py + pix_x;
            const int gid = a.hit_image[pid];
            if (gid >= 0)
                bwd_depth_path(a.scales, a.rotations, a.means3D, a.view, a.hit_geo, a.plane, (size_t)tile * 256 + ly * 16 + lx,
                               a.gacc, a.touched, gid, a.dL_ddepth[pid], pix_x, py, a.fx, a.fy, a.cx, a.cy, a.depth_thr,
                               a.normal_thr);
        }
    }
}

struct GaussBwdArgs {
    int P, D, M;
    float scale_modifier, tanfovx, tanfovy, focal_x, focal_y;
    const float *means3D, *scales, *rotations, *shs, *cov3D_precomp;
    const float *f_rest; // split-SH mode: shs = f_dc [P,3], f_rest [P,45]
    const float *view, *proj, *campos;
    const int *radii;
    const uint8_t *clamped;
    double *gacc; // read, and the records that were non-zero cleared again (dqo_rast_settings.geom_clean)
    uint8_t *touched; // one byte instead of the 128-byte record for the Gaussians nothing was added to
    float *dL_dmeans2D, *dL_dconic, *dL_dopacity, *dL_dcolors, *dL_dmeans3D, *dL_dcov3D, *dL_dsh, *dL_dscales, *dL_drot;
    // Fused mapping step only (nullptr otherwise): ever[i] != 0 once Gaussian i has received a non-zero gradient.  A
    // Gaussian that never has is a fixed point of Adam (zero gradient on zero moments), so its gradients are neither
    // written here nor read by the optimiser kernels: one byte instead of ~240 B written and ~600 B read per step.
    uint8_t *ever;
    uint32_t *ever_list; // optional compact list of the Gaussians whose flag is set (append order), with its length
    int *ever_count;
    // dqo_rast_settings.geom_clean == 2 (nullptr otherwise): out_nz[i] != 0 where the previous call wrote a non-zero row
    // into the caller's gradient tensors; rows that were zero and stay zero are not written again.
    uint8_t *out_nz;
    // extra colour set (see RenderBwdArgs): accumulators f64[4] per Gaussian, consumed and cleared like gacc; the gradient
    // w.r.t. the extra colours [P,3] is written under the same rules as the other rows (nullptr: none)
    double *cacc;
    float *dL_dextra;
    // 1: every block takes the thread-per-Gaussian branch (callers without persistent outputs: every row is written, zeros
    // included, through the coalesced staged stores)
    int force_dense;
    int chunks; // 128-Gaussian chunks per block, 1 .. GB_MAX_CHUNKS
};

__device__ __constant__ float B_SH_C0 = 0.28209479177387814f;
__device__ __constant__ float B_SH_C1 = 0.4886025119029199f;
__device__ __constant__ float B_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                            -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float B_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                            0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                            -0.5900435899266435f};

#define GB_THREADS 128
#define GB_ROW_Q 13 // float4 per staged SH row (12 used + 1 pad)

// STAGED (M == 16, 16-byte aligned shs / dL_dsh): each warp moves its 32 x 192 B of SH coefficients in and its
// 32 x 192 B of SH gradients out with coalesced 128-bit accesses through a padded shared-memory tile.
// SHMODE 1: staged merged SH, 2: staged split f_dc / f_rest inputs (gradients are still written merged), 0: plain.
//
// One Gaussian per thread.  COMPACT = false: thread t of the warp owns Gaussian base_g + t (consecutive rows: the staged
// paths can move whole 32-row blocks).  COMPACT = true: the warp's lanes own arbitrary (touched) Gaussians handed in through
// `idx`, live lanes first; `nrow` live lanes; the staged paths then move single rows, addressed through a shuffle of idx.
// Inlined at exactly ONE call site of the one kernel every caller launches (operator path, RasterPipeline, fused step), so a
// Gaussian's gradients do not depend on which form processed it: separately inlined copies were optimised separately and
// differed in FMA contraction (last-ulp differences between the pipeline and the operator path), and a non-inlined
// function reads the kernel arguments through generic loads (-15 % on the small clouds of the object workload).
template <int SHMODE>
__device__ __forceinline__ void gaussian_backward_body(const GaussBwdArgs &a, const int idx, const bool live, const int base_g,
                                                    const int nrow, float4 *s_row, const bool COMPACT) {
    constexpr bool STAGED = SHMODE != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *wbuf = s_row + warp * 32 * GB_ROW_Q;
    float *wrest = reinterpret_cast<float *>(wbuf), *wdc = wrest + 1440;
    // only Gaussians the blend added to (a visible Gaussian that no pixel with a gradient saw has an all-zero record)
    const bool active = live && (COMPACT || a.touched[idx] != 0);
    const int M = a.M;
    float g[DQO_GACC_FLOATS];
    if (active) {
        const double2 *gp = reinterpret_cast<const double2 *>(a.gacc + (size_t)idx * DQO_GACC_FLOATS);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double2 t = gp[k];
            g[2 * k] = (float)t.x; g[2 * k + 1] = (float)t.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < DQO_GACC_FLOATS; k++) g[k] = 0.f;
    }
    // Every output is a sum of products with the accumulator record, so a Gaussian whose record is all zero (most of a
    // large map: it is in the frustum but contributes to no pixel that carries a gradient) has all-zero gradients and
    // needs neither its parameters (236 B) nor the chain below: only Gaussians with `need` are loaded and evaluated.
    bool nz = false;
#pragma unroll
    for (int k = 0; k < DQO_GACC_FLOATS; k++) nz |= (g[k] != 0.f);
    float ex[3] = {0.f, 0.f, 0.f};
    if (a.cacc && active) {
        double2 *cp = reinterpret_cast<double2 *>(a.cacc + (size_t)idx * 4);
        const double2 c0 = cp[0], c1 = cp[1];
        ex[0] = (float)c0.x; ex[1] = (float)c0.y; ex[2] = (float)c1.x;
        if (ex[0] != 0.f || ex[1] != 0.f || ex[2] != 0.f) {
            nz = true;
            cp[0] = make_double2(0.0, 0.0);
            cp[1] = make_double2(0.0, 0.0);
        }
    }
    const bool need = active && nz;
    if (active) a.touched[idx] = 0;
    if (need) { // leave the accumulator clean for the next backward pass
        double2 *gz = reinterpret_cast<double2 *>(a.gacc + (size_t)idx * DQO_GACC_FLOATS);
#pragma unroll
        for (int k = 0; k < 8; k++) gz[k] = make_double2(0.0, 0.0);
    }
    const unsigned need_rows = __ballot_sync(0xFFFFFFFFu, need);
    const bool whole_block = !COMPACT && nrow == 32 && __popc(need_rows) >= 12; // dense: the warp's block in 12 coalesced 128-bit loads
    if (SHMODE == 2 && need_rows) {
        if (whole_block) {
            const float4 *g4 = reinterpret_cast<const float4 *>(a.f_rest + (size_t)base_g * 45);
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const int q = i * 32 + lane;
                if (q < 360) wbuf[q] = __ldg(g4 + q);
            }
            const float4 *gd4 = reinterpret_cast<const float4 *>(a.shs + (size_t)base_g * 3);
            if (lane < 24) reinterpret_cast<float4 *>(wdc)[lane] = __ldg(gd4 + lane);
        } else if (COMPACT) { // scattered rows: every lane fetches its own (48 independent loads in flight per lane; a
            if (need) {       // row-by-row loop pays one memory latency per row)
                const float *src = a.f_rest + (size_t)idx * 45;
#pragma unroll
                for (int k = 0; k < 45; k++) wrest[lane * 45 + k] = __ldg(src + k);
#pragma unroll
                for (int c = 0; c < 3; c++) wdc[lane * 3 + c] = __ldg(a.shs + (size_t)idx * 3 + c);
            }
        } else { // sparse: only the needed rows, 180 + 12 contiguous bytes each
            for (unsigned m = need_rows; m; m &= m - 1) {
                const int r = __ffs(m) - 1;
                const size_t row = (size_t)(base_g + r);
                const float *src = a.f_rest + row * 45;
                wrest[r * 45 + lane] = __ldg(src + lane);
                if (lane < 13) wrest[r * 45 + 32 + lane] = __ldg(src + 32 + lane);
                else if (lane < 16) wdc[r * 3 + (lane - 13)] = __ldg(a.shs + row * 3 + (lane - 13));
            }
        }
        __syncwarp();
    }
    if (SHMODE == 1 && need_rows) {
        const float4 *gsh = reinterpret_cast<const float4 *>(a.shs) + (size_t)(COMPACT ? 0 : base_g) * 12;
        if (whole_block) {
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const int q = i * 32 + lane;
                wbuf[(q / 12) * GB_ROW_Q + (q % 12)] = __ldg(gsh + q);
            }
        } else if (COMPACT) {
            if (need) {
#pragma unroll
                for (int i = 0; i < 12; i++) wbuf[lane * GB_ROW_Q + i] = __ldg(gsh + (size_t)idx * 12 + i);
            }
        } else {
            for (unsigned m = need_rows; m; m &= m - 1) {
                const int r = __ffs(m) - 1;
                const size_t row = (size_t)r;
                if (lane < 12) wbuf[r * GB_ROW_Q + lane] = __ldg(gsh + row * 12 + lane);
            }
        }
        __syncwarp();
    }
    float dmean[3] = {g[9], g[10], g[11]};
    float drot[4] = {g[12], g[13], g[14], g[15]};
    float dscale[3] = {0.f, 0.f, 0.f};
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bool write_out = live;
    if (a.out_nz && live) {
        const bool was = a.out_nz[idx] != 0;
        if (was != nz) a.out_nz[idx] = nz ? 1 : 0;
        write_out = nz || was;
    }
    if (a.ever && live) {
        const bool was = a.ever[idx] != 0;
        if (nz && !was) {
            a.ever[idx] = 1;
            if (a.ever_list) a.ever_list[atomicAdd(a.ever_count, 1)] = (uint32_t)idx; // (the compiler aggregates per warp)
        }
        write_out = nz || was;
    }
    float *dsh = (a.dL_dsh && live) ? a.dL_dsh + (size_t)idx * M * 3 : nullptr;
    float shg[48]; // staged path: this Gaussian's SH gradients (flat [k][c])
    if (STAGED) {
#pragma unroll
        for (int k = 0; k < 48; k++) shg[k] = 0.f;
    }

    if (need) {
        const float mx = a.means3D[3 * idx], my = a.means3D[3 * idx + 1], mz = a.means3D[3 * idx + 2];
        const float *v = a.view;
        float cov3[6];
        float sx = 0, sy = 0, sz = 0;
        float4 q = make_float4(0, 0, 0, 0);
        if (a.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) cov3[k] = a.cov3D_precomp[6 * idx + k];
        } else {
            sx = a.scales[3 * idx]; sy = a.scales[3 * idx + 1]; sz = a.scales[3 * idx + 2];
            q = reinterpret_cast<const float4 *>(a.rotations)[idx];
            cov3d_from_scale_rot(sx, sy, sz, a.scale_modifier, q.x, q.y, q.z, q.w, cov3);
        }
        // ---- conic -> cov2D -> cov3D, mean (backward.cu:273-422) ----
        // This chain is ill-conditioned in fp32 (denom = a*c - b*b and the dL_da/db/dc sums cancel for
        // edge-on surfels), so it is written in the exact operation order of the reference's compiled
        // computeCov2DCUDA; any other association differs from the reference by up to ~1e-3 relative.
        {
            const float dcx = g[2], dcy = g[3], dcz = g[4];
            float tx = xform_row(v, 0, mx, my, mz), ty = xform_row(v, 1, mx, my, mz);
            const float tz = xform_row(v, 2, mx, my, mz);
            const float limx = fmul(1.3f, a.tanfovx), limy = fmul(1.3f, a.tanfovy);
            const float txtz = fdiv(tx, tz), tytz = fdiv(ty, tz);
            tx = fmul(fminf(limx, fmaxf(-limx, txtz)), tz);
            ty = fmul(fminf(limy, fmaxf(-limy, tytz)), tz);
            const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
            const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
            const float hx = a.focal_x, hy = a.focal_y;
            const float tzz = fmul(tz, tz);
            const float J00 = fdiv(hx, tz), J11 = fdiv(hy, tz);
            const float J02 = fdiv(fmul(tx, -hx), tzz), J12 = fdiv(fmul(ty, -hy), tzz);
            float T0[3], T1[3]; // T[0][r], T[1][r]
#pragma unroll
            for (int r = 0; r < 3; r++) {
                T0[r] = ffma(J02, v[4 * r + 2], fmul(v[4 * r], J00));
                T1[r] = ffma(J12, v[4 * r + 2], fmul(v[4 * r + 1], J11));
            }
            const float V[3][3] = {{cov3[0], cov3[1], cov3[2]}, {cov3[1], cov3[3], cov3[4]}, {cov3[2], cov3[4], cov3[5]}};
            float TV0[3], TV1[3]; // (T0 . V[k]), (T1 . V[k])
#pragma unroll
            for (int k = 0; k < 3; k++) {
                TV0[k] = dot3_ref(T0[0], V[k][0], T0[1], V[k][1], T0[2], V[k][2]);
                TV1[k] = dot3_ref(T1[0], V[k][0], T1[1], V[k][1], T1[2], V[k][2]);
            }
            // here the compiled reference multiplies the FIRST term and fuses the other two
            const float ca = fadd(ffma(T0[2], TV0[2], ffma(T0[1], TV0[1], fmul(T0[0], TV0[0]))), 0.3f);
            const float cb = ffma(T0[2], TV1[2], ffma(T0[1], TV1[1], fmul(T0[0], TV1[0])));
            const float cc = fadd(ffma(T1[2], TV1[2], ffma(T1[1], TV1[1], fmul(T1[0], TV1[0]))), 0.3f);
            const float ac = fmul(ca, cc);
            const float denom = ffma(-cb, cb, ac);
            float dL_da = 0, dL_db = 0, dL_dc = 0;
            const float denom2inv = frcp(ffma(denom, denom, 0.0000001f));
            if (denom2inv != 0) {
                const float b2 = fadd(cb, cb), a2 = fadd(ca, ca);
                const float dmac = fadd(denom, -ac);
                dL_da = fmul(ffma(dcz, dmac, ffma(dcy, fmul(cc, b2), -fmul(dcx, fmul(cc, cc)))), denom2inv);
                dL_dc = fmul(ffma(dcx, dmac, ffma(dcy, fmul(cb, a2), -fmul(dcz, fmul(ca, ca)))), denom2inv);
                dL_db = fmul(fadd(denom2inv, denom2inv),
                             ffma(dcz, fmul(cb, ca), ffma(dcx, fmul(cb, cc), -fmul(dcy, ffma(cb, b2, denom)))));
                dcov[0] = ffma(dL_dc, fmul(T1[0], T1[0]), ffma(dL_da, fmul(T0[0], T0[0]), fmul(dL_db, fmul(T0[0], T1[0]))));
                dcov[3] = ffma(dL_dc, fmul(T1[1], T1[1]), ffma(dL_da, fmul(T0[1], T0[1]), fmul(dL_db, fmul(T0[1], T1[1]))));
                dcov[5] = ffma(dL_dc, fmul(T1[2], T1[2]), ffma(dL_da, fmul(T0[2], T0[2]), fmul(dL_db, fmul(T0[2], T1[2]))));
                const float T00x2 = fadd(T0[0], T0[0]), T10x2 = fadd(T1[0], T1[0]);
                const float T02x2 = fadd(T0[2], T0[2]), T11x2 = fadd(T1[1], T1[1]);
                dcov[1] = ffma(dL_dc, fmul(T1[1], T10x2),
                               ffma(dL_da, fmul(T0[1], T00x2), fmul(dL_db, ffma(T0[0], T1[1], fmul(T0[1], T1[0])))));
                dcov[2] = ffma(dL_dc, fmul(T1[2], T10x2),
                               ffma(dL_da, fmul(T0[2], T00x2), fmul(dL_db, ffma(T0[0], T1[2], fmul(T0[2], T1[0])))));
                dcov[4] = ffma(dL_dc, fmul(T1[2], T11x2),
                               ffma(dL_da, fmul(T0[1], T02x2), fmul(dL_db, ffma(T0[1], T1[2], fmul(T0[2], T1[1])))));
            }
            float dT0[3], dT1[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                dT0[k] = ffma(fadd(TV0[k], TV0[k]), dL_da, fmul(TV1[k], dL_db));
                dT1[k] = ffma(TV0[k], dL_db, fmul(fadd(TV1[k], TV1[k]), dL_dc));
            }
            const float dJ00 = dot3_ref(v[0], dT0[0], v[4], dT0[1], v[8], dT0[2]);
            const float dJ02 = dot3_ref(v[2], dT0[0], v[6], dT0[1], v[10], dT0[2]);
            const float dJ11 = dot3_ref(v[1], dT1[0], v[5], dT1[1], v[9], dT1[2]);
            const float dJ12 = dot3_ref(v[2], dT1[0], v[6], dT1[1], v[10], dT1[2]);
            const float itz = frcp(tz), itz2 = fmul(itz, itz), itz3 = fmul(itz2, itz);
            const float dtx = fmul(dJ02, fmul(itz2, fmul(x_grad_mul, -hx)));
            const float dty = fmul(dJ12, fmul(itz2, fmul(y_grad_mul, -hy)));
            float dtz = ffma(dJ00, fmul(itz2, -hx), -fmul(dJ11, fmul(itz2, hy)));
            dtz = ffma(dJ02, fmul(itz3, fmul(tx, fadd(hx, hx))), dtz);
            dtz = ffma(dJ12, fmul(itz3, fmul(ty, fadd(hy, hy))), dtz);
            dmean[0] = fadd(dot3_ref(dtx, v[0], dty, v[1], dtz, v[2]), dmean[0]);
            dmean[1] = fadd(dot3_ref(dtx, v[4], dty, v[5], dtz, v[6]), dmean[1]);
            dmean[2] = fadd(dot3_ref(dtx, v[8], dty, v[9], dtz, v[10]), dmean[2]);
        }
        // ---- mean2D -> mean3D through the projection (backward.cu:516-533) ----
        {
            const float *p = a.proj;
            const float hw = p[3] * mx + p[7] * my + p[11] * mz + p[15];
            const float m_w = 1.0f / (hw + 0.0000001f);
            const float mul1 = (p[0] * mx + p[4] * my + p[8] * mz + p[12]) * m_w * m_w;
            const float mul2 = (p[1] * mx + p[5] * my + p[9] * mz + p[13]) * m_w * m_w;
            const float d2x = g[0], d2y = g[1];
            dmean[0] += (p[0] * m_w - p[3] * mul1) * d2x + (p[1] * m_w - p[3] * mul2) * d2y;
            dmean[1] += (p[4] * m_w - p[7] * mul1) * d2x + (p[5] * m_w - p[7] * mul2) * d2y;
            dmean[2] += (p[8] * m_w - p[11] * mul1) * d2x + (p[9] * m_w - p[11] * mul2) * d2y;
        }
        // ---- colour -> SH (+ view direction -> mean) (backward.cu:152-268) ----
        if (a.shs) {
            float sh[48];
            if (SHMODE == 1) {
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    const float4 t = wbuf[lane * GB_ROW_Q + i];
                    sh[4 * i] = t.x; sh[4 * i + 1] = t.y; sh[4 * i + 2] = t.z; sh[4 * i + 3] = t.w;
                }
            } else if (SHMODE == 2) {
#pragma unroll
                for (int c = 0; c < 3; c++) sh[c] = wdc[lane * 3 + c];
#pragma unroll
                for (int k = 0; k < 45; k++) sh[3 + k] = wrest[lane * 45 + k];
            } else {
                const float *gp = a.shs + (size_t)idx * M * 3;
                const int nload = (a.D + 1) * (a.D + 1) * 3;
#pragma unroll
                for (int k = 0; k < 48; k++) sh[k] = (k < nload) ? gp[k] : 0.f;
            }
            const uint8_t cl = a.clamped[idx];
            const float dRGB[3] = {(cl & 1) ? 0.f : g[6], (cl & 2) ? 0.f : g[7], (cl & 4) ? 0.f : g[8]};
            const float ox = mx - a.campos[0], oy = my - a.campos[1], oz = mz - a.campos[2];
            const float inv_len = 1.0f / sqrtf(ox * ox + oy * oy + oz * oz);
            const float x = ox * inv_len, y = oy * inv_len, z = oz * inv_len;
            float w[16]; // dRGB / dsh_k
#pragma unroll
            for (int k = 0; k < 16; k++) w[k] = 0.f;
            float dx_[3] = {0, 0, 0}, dy_[3] = {0, 0, 0}, dz_[3] = {0, 0, 0}; // dRGB/d dir per channel
            w[0] = B_SH_C0;
            const int deg = a.D;
            if (deg > 0) {
                w[1] = -B_SH_C1 * y; w[2] = B_SH_C1 * z; w[3] = -B_SH_C1 * x;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    dx_[c] = -B_SH_C1 * sh[9 + c];
                    dy_[c] = -B_SH_C1 * sh[3 + c];
                    dz_[c] = B_SH_C1 * sh[6 + c];
                }
                if (deg > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    w[4] = B_SH_C2[0] * xy; w[5] = B_SH_C2[1] * yz; w[6] = B_SH_C2[2] * (2.f * zz - xx - yy);
                    w[7] = B_SH_C2[3] * xz; w[8] = B_SH_C2[4] * (xx - yy);
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        dx_[c] += B_SH_C2[0] * y * sh[12 + c] + B_SH_C2[2] * 2.f * -x * sh[18 + c] + B_SH_C2[3] * z * sh[21 + c] +
                                  B_SH_C2[4] * 2.f * x * sh[24 + c];
                        dy_[c] += B_SH_C2[0] * x * sh[12 + c] + B_SH_C2[1] * z * sh[15 + c] + B_SH_C2[2] * 2.f * -y * sh[18 + c] +
                                  B_SH_C2[4] * 2.f * -y * sh[24 + c];
                        dz_[c] += B_SH_C2[1] * y * sh[15 + c] + B_SH_C2[2] * 2.f * 2.f * z * sh[18 + c] + B_SH_C2[3] * x * sh[21 + c];
                    }
                    if (deg > 2) {
                        w[9] = B_SH_C3[0] * y * (3.f * xx - yy); w[10] = B_SH_C3[1] * xy * z;
                        w[11] = B_SH_C3[2] * y * (4.f * zz - xx - yy);
                        w[12] = B_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                        w[13] = B_SH_C3[4] * x * (4.f * zz - xx - yy); w[14] = B_SH_C3[5] * z * (xx - yy);
                        w[15] = B_SH_C3[6] * x * (xx - 3.f * yy);
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            dx_[c] += B_SH_C3[0] * sh[27 + c] * 3.f * 2.f * xy + B_SH_C3[1] * sh[30 + c] * yz +
                                      B_SH_C3[2] * sh[33 + c] * -2.f * xy + B_SH_C3[3] * sh[36 + c] * -3.f * 2.f * xz +
                                      B_SH_C3[4] * sh[39 + c] * (-3.f * xx + 4.f * zz - yy) + B_SH_C3[5] * sh[42 + c] * 2.f * xz +
                                      B_SH_C3[6] * sh[45 + c] * 3.f * (xx - yy);
                            dy_[c] += B_SH_C3[0] * sh[27 + c] * 3.f * (xx - yy) + B_SH_C3[1] * sh[30 + c] * xz +
                                      B_SH_C3[2] * sh[33 + c] * (-3.f * yy + 4.f * zz - xx) +
                                      B_SH_C3[3] * sh[36 + c] * -3.f * 2.f * yz + B_SH_C3[4] * sh[39 + c] * -2.f * xy +
                                      B_SH_C3[5] * sh[42 + c] * -2.f * yz + B_SH_C3[6] * sh[45 + c] * -3.f * 2.f * xy;
                            dz_[c] += B_SH_C3[1] * sh[30 + c] * xy + B_SH_C3[2] * sh[33 + c] * 4.f * 2.f * yz +
                                      B_SH_C3[3] * sh[36 + c] * 3.f * (2.f * zz - xx - yy) +
                                      B_SH_C3[4] * sh[39 + c] * 4.f * 2.f * xz + B_SH_C3[5] * sh[42 + c] * (xx - yy);
                        }
                    }
                }
            }
            if (STAGED) {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    shg[3 * k] = w[k] * dRGB[0];
                    shg[3 * k + 1] = w[k] * dRGB[1];
                    shg[3 * k + 2] = w[k] * dRGB[2];
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k < M) {
                        dsh[3 * k] = w[k] * dRGB[0];
                        dsh[3 * k + 1] = w[k] * dRGB[1];
                        dsh[3 * k + 2] = w[k] * dRGB[2];
                    }
                }
                for (int k = 48; k < 3 * M; k++) dsh[k] = 0.f;
            }
            const float ddx = dx_[0] * dRGB[0] + dx_[1] * dRGB[1] + dx_[2] * dRGB[2];
            const float ddy = dy_[0] * dRGB[0] + dy_[1] * dRGB[1] + dy_[2] * dRGB[2];
            const float ddz = dz_[0] * dRGB[0] + dz_[1] * dRGB[1] + dz_[2] * dRGB[2];
            // dnormvdv (auxiliary.h:107-117)
            const float sum2 = ox * ox + oy * oy + oz * oz;
            const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dmean[0] += ((+sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * invsum32;
            dmean[1] += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * invsum32;
            dmean[2] += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * invsum32;
        }
        // ---- cov3D -> scale, rotation (backward.cu:426-487) ----
        if (!a.cov3D_precomp) {
            const QuatMat R = quat_to_glm(q.x, q.y, q.z, q.w);
            const float s[3] = {a.scale_modifier * sx, a.scale_modifier * sy, a.scale_modifier * sz};
            const float Rc[3][3] = {{R.c0[0], R.c0[1], R.c0[2]}, {R.c1[0], R.c1[1], R.c1[2]}, {R.c2[0], R.c2[1], R.c2[2]}};
            float Mm[3][3]; // M[c][r] = s[r] * R[c][r]
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int r = 0; r < 3; r++) Mm[c][r] = s[r] * Rc[c][r];
            const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            float dM[3][3]; // dL_dM[c][r] = 2 * sum_k M[k][r] * dS[c][k]
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int r = 0; r < 3; r++)
                    dM[c][r] = 2.0f * (Mm[0][r] * dS[c][0] + Mm[1][r] * dS[c][1] + Mm[2][r] * dS[c][2]);
#pragma unroll
            for (int r = 0; r < 3; r++) dscale[r] = Rc[0][r] * dM[0][r] + Rc[1][r] * dM[1][r] + Rc[2][r] * dM[2][r];
            float Mt[3][3]; // dL_dMt[r][c] = s[r] * dM[c][r]
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) Mt[r][c] = s[r] * dM[c][r];
            const float r_ = q.x, x = q.y, y = q.z, z = q.w;
            drot[0] += 2 * z * (Mt[0][1] - Mt[1][0]) + 2 * y * (Mt[2][0] - Mt[0][2]) + 2 * x * (Mt[1][2] - Mt[2][1]);
            drot[1] += 2 * y * (Mt[1][0] + Mt[0][1]) + 2 * z * (Mt[2][0] + Mt[0][2]) + 2 * r_ * (Mt[1][2] - Mt[2][1]) -
                       4 * x * (Mt[2][2] + Mt[1][1]);
            drot[2] += 2 * x * (Mt[1][0] + Mt[0][1]) + 2 * r_ * (Mt[2][0] - Mt[0][2]) + 2 * z * (Mt[1][2] + Mt[2][1]) -
                       4 * y * (Mt[2][2] + Mt[0][0]);
            drot[3] += 2 * r_ * (Mt[0][1] - Mt[1][0]) + 2 * x * (Mt[2][0] + Mt[0][2]) + 2 * y * (Mt[1][2] + Mt[2][1]) -
                       4 * z * (Mt[1][1] + Mt[0][0]);
        }
    } else if (!STAGED && dsh) {
        for (int k = 0; k < 3 * M; k++) dsh[k] = 0.f;
    }
    if (!STAGED && need && !a.shs && dsh) {
        for (int k = 0; k < 3 * M; k++) dsh[k] = 0.f;
    }
    if (STAGED && COMPACT) { // scattered rows: every lane stores its own 192 B
        if (write_out) {
            float4 *d = reinterpret_cast<float4 *>(a.dL_dsh) + (size_t)idx * 12;
#pragma unroll
            for (int i = 0; i < 12; i++) d[i] = make_float4(shg[4 * i], shg[4 * i + 1], shg[4 * i + 2], shg[4 * i + 3]);
        }
    } else if (STAGED) { // own row -> shared -> coalesced 128-bit stores of the warp's contiguous 32 x 192 B block
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 12; i++)
            wbuf[lane * GB_ROW_Q + i] = make_float4(shg[4 * i], shg[4 * i + 1], shg[4 * i + 2], shg[4 * i + 3]);
        __syncwarp();
        const unsigned rows_out = __ballot_sync(0xFFFFFFFFu, write_out);
        if (nrow > 0) {
            float4 *gd = reinterpret_cast<float4 *>(a.dL_dsh) + (size_t)base_g * 12;
            const int nq = nrow * 12;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const int q = i * 32 + lane;
                if (q < nq && ((rows_out >> (q / 12)) & 1)) gd[q] = wbuf[(q / 12) * GB_ROW_Q + (q % 12)];
            }
        }
    }
    if (!write_out) return;

    if (a.dL_dextra) {
        a.dL_dextra[3 * idx] = ex[0];
        a.dL_dextra[3 * idx + 1] = ex[1];
        a.dL_dextra[3 * idx + 2] = ex[2];
    }
    if (a.dL_dmeans2D) {
        a.dL_dmeans2D[3 * idx] = g[0];
        a.dL_dmeans2D[3 * idx + 1] = g[1];
        a.dL_dmeans2D[3 * idx + 2] = 0.f;
    }
    if (a.dL_dconic) reinterpret_cast<float4 *>(a.dL_dconic)[idx] = make_float4(g[2], g[3], 0.f, g[4]);
    if (a.dL_dopacity) a.dL_dopacity[idx] = g[5];
    if (a.dL_dcolors) {
        a.dL_dcolors[3 * idx] = g[6];
        a.dL_dcolors[3 * idx + 1] = g[7];
        a.dL_dcolors[3 * idx + 2] = g[8];
    }
    a.dL_dmeans3D[3 * idx] = dmean[0];
    a.dL_dmeans3D[3 * idx + 1] = dmean[1];
    a.dL_dmeans3D[3 * idx + 2] = dmean[2];
    if (a.dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; k++) a.dL_dcov3D[6 * idx + k] = dcov[k];
    }
    if (a.dL_dscales) {
        a.dL_dscales[3 * idx] = dscale[0];
        a.dL_dscales[3 * idx + 1] = dscale[1];
        a.dL_dscales[3 * idx + 2] = dscale[2];
    }
    if (a.dL_drot) reinterpret_cast<float4 *>(a.dL_drot)[idx] = make_float4(drot[0], drot[1], drot[2], drot[3]);
}

// A Gaussian the blend did not touch: all-zero gradients.  With the persistent-output rules (out_nz / ever, see
// GaussBwdArgs) its rows are rewritten -- with zeros -- only if they held something.
template <int SHMODE>
__device__ __forceinline__ void gaussian_backward_untouched(const GaussBwdArgs &a, const int idx) {
    bool write_out = true;
    if (a.out_nz) {
        write_out = a.out_nz[idx] != 0;
        if (write_out) a.out_nz[idx] = 0;
    }
    if (a.ever) write_out = a.ever[idx] != 0;
    if (!write_out) return;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.dL_dsh) {
        if (SHMODE != 0) { // M == 16, 16-byte aligned rows
            float4 *d = reinterpret_cast<float4 *>(a.dL_dsh) + (size_t)idx * 12;
#pragma unroll
            for (int i = 0; i < 12; i++) d[i] = z4;
        } else {
            float *d = a.dL_dsh + (size_t)idx * a.M * 3;
            for (int k = 0; k < 3 * a.M; k++) d[k] = 0.f;
        }
    }
    if (a.dL_dextra) a.dL_dextra[3 * idx] = a.dL_dextra[3 * idx + 1] = a.dL_dextra[3 * idx + 2] = 0.f;
    if (a.dL_dmeans2D) a.dL_dmeans2D[3 * idx] = a.dL_dmeans2D[3 * idx + 1] = a.dL_dmeans2D[3 * idx + 2] = 0.f;
    if (a.dL_dconic) reinterpret_cast<float4 *>(a.dL_dconic)[idx] = z4;
    if (a.dL_dopacity) a.dL_dopacity[idx] = 0.f;
    if (a.dL_dcolors) a.dL_dcolors[3 * idx] = a.dL_dcolors[3 * idx + 1] = a.dL_dcolors[3 * idx + 2] = 0.f;
    a.dL_dmeans3D[3 * idx] = a.dL_dmeans3D[3 * idx + 1] = a.dL_dmeans3D[3 * idx + 2] = 0.f;
    if (a.dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; k++) a.dL_dcov3D[6 * idx + k] = 0.f;
    }
    if (a.dL_dscales) a.dL_dscales[3 * idx] = a.dL_dscales[3 * idx + 1] = a.dL_dscales[3 * idx + 2] = 0.f;
    if (a.dL_drot) reinterpret_cast<float4 *>(a.dL_drot)[idx] = z4;
}

// The per-Gaussian pass.  One view touches about 1 % of a 1 M-Gaussian map: with a thread per Gaussian a third of all warps
// ran the whole chain for one or two live lanes.  A block owns `chunks` x 128 consecutive Gaussians, collects the touched
// ones in shared memory and runs the chain on the compacted list (one warp-chain per 32 touched Gaussians); the untouched
// ones cost their flag bytes.  Blocks in which most Gaussians are touched (small object clouds) and callers without
// persistent outputs (force_dense) keep the thread-per-Gaussian form with its whole-block staged SH loads and stores.
#define GB_MAX_CHUNKS 16
template <int SHMODE>
__global__ void __launch_bounds__(GB_THREADS, 4) gaussian_backward_kernel(const GaussBwdArgs a) {
    pdl_enter();
    extern __shared__ float4 s_row[];
    __shared__ int s_ids[GB_THREADS * GB_MAX_CHUNKS];
    __shared__ int s_n;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int chunks = a.chunks; // 128-Gaussian chunks per block (host: sized so that the grid is about one wave)
    const int first = blockIdx.x * (GB_THREADS * chunks);
    if (tid == 0) s_n = 0;
    __shared__ int s_total;
    if (tid == 0) s_total = 0;
    __syncthreads();
    unsigned tmask = 0;
#pragma unroll 4
    for (int c = 0; c < chunks; c++) { // independent loads: all in flight together
        const int i = first + c * GB_THREADS + tid;
        const bool t = i < a.P && a.touched[i] != 0;
        tmask |= (t ? 1u : 0u) << c;
    }
    {
        int cnt = __popc(tmask);
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        if ((tid & 31) == 0 && cnt) atomicAdd(&s_total, cnt);
    }
    __syncthreads();
    const int total = s_total;
    if (total == 0 && !a.force_dense) { // nothing touched: flags only
        for (int c = 0; c < chunks; c++) {
            const int i = first + c * GB_THREADS + tid;
            if (i < a.P) gaussian_backward_untouched<SHMODE>(a, i);
        }
        return;
    }
    const bool dense = a.force_dense || 2 * total >= GB_THREADS * chunks;
    int n = GB_THREADS * chunks;
    if (!dense) {
        for (int c = 0; c < chunks; c++) {
            const int i = first + c * GB_THREADS + tid;
            if ((tmask >> c) & 1u) s_ids[atomicAdd(&s_n, 1)] = i;
            else if (i < a.P) gaussian_backward_untouched<SHMODE>(a, i);
        }
        __syncthreads();
        n = s_n;
    }
    // one loop, one inlined copy of the chain: over the block's chunks (dense) or over the compacted list (sparse)
#pragma unroll 1
    for (int base = 0; base < n; base += GB_THREADS) {
        int idx, base_g, nrow;
        bool live;
        if (dense) {
            idx = first + base + tid;
            live = idx < a.P;
            base_g = first + base + warp * 32;
            nrow = min(32, a.P - base_g);
        } else {
            const int j = base + tid;
            live = j < n;
            idx = live ? s_ids[j] : 0;
            base_g = 0;
            nrow = max(0, min(32, n - (base + warp * 32)));
        }
        gaussian_backward_body<SHMODE>(a, idx, live, base_g, nrow, s_row, !dense);
        __syncwarp();
    }
}

} // namespace dqo

using namespace dqo;

namespace dqo {
int rast_backward_impl(const dqo_rast_settings *s, const float *background, const float *means3D, const float *shs,
                       const float *f_rest, const float *colors_precomp, const float *scales, const float *rotations,
                       const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix, const float *campos,
                       const int32_t *radii, void *geom_buffer, const void *binning_buffer, int64_t capacity,
                       const void *image_buffer, const int32_t *status, const float *dL_dout_color,
                       const float *dL_dout_depth, const int32_t *hit_image, float *dL_dmeans2D, float *dL_dconic,
                       float *dL_dopacity, float *dL_dcolors, float *dL_dmeans3D, float *dL_dcov3D, float *dL_dsh,
                       float *dL_dscales, float *dL_drotations, uint8_t *ever, uint32_t *ever_list,
                       int32_t *ever_count, void *stream_, const ExtraBlendGrad *extra);
}

extern "C" int dqo_rast_geom_init(int32_t P, void *geom_buffer, void *stream_) {
    if (P < 0 || (P > 0 && !geom_buffer)) {
        set_error("dqo_rast_geom_init: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    if (P == 0) return DQO_OK;
    GeomLayout GL;
    if (make_geom_layout(P, &GL)) return DQO_ERR_WORKSPACE;
    DQO_CUDA_CHECK(cudaMemsetAsync((char *)geom_buffer + GL.gacc, 0, (size_t)P * DQO_GACC_FLOATS * sizeof(double),
                                   (cudaStream_t)stream_));
    DQO_CUDA_CHECK(cudaMemsetAsync((char *)geom_buffer + GL.touched, 0, (size_t)P, (cudaStream_t)stream_));
    DQO_CUDA_CHECK(cudaMemsetAsync((char *)geom_buffer + GL.out_nz, 0, (size_t)P, (cudaStream_t)stream_));
    return DQO_OK;
}

extern "C" int dqo_rast_backward(const dqo_rast_settings *s, const float *background, const float *means3D,
                                 const float *shs, const float *colors_precomp, const float *scales,
                                 const float *rotations, const float *cov3D_precomp, const float *viewmatrix,
                                 const float *projmatrix, const float *campos, const int32_t *radii,
                                 void *geom_buffer, const void *binning_buffer, int64_t capacity,
                                 const void *image_buffer, const int32_t *status, const float *dL_dout_color,
                                 const float *dL_dout_depth, const int32_t *hit_image, float *dL_dmeans2D,
                                 float *dL_dconic, float *dL_dopacity, float *dL_dcolors, float *dL_dmeans3D,
                                 float *dL_dcov3D, float *dL_dsh, float *dL_dscales, float *dL_drotations,
                                 void *stream_) {
    return rast_backward_impl(s, background, means3D, shs, nullptr, colors_precomp, scales, rotations, cov3D_precomp,
                              viewmatrix, projmatrix, campos, radii, geom_buffer, binning_buffer, capacity, image_buffer,
                              status, dL_dout_color, dL_dout_depth, hit_image, dL_dmeans2D, dL_dconic, dL_dopacity,
                              dL_dcolors, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, nullptr, nullptr, nullptr,
                              stream_, nullptr);
}

int dqo::rast_backward_impl(const dqo_rast_settings *s, const float *background, const float *means3D,
                                 const float *shs, const float *f_rest, const float *colors_precomp, const float *scales,
                                 const float *rotations, const float *cov3D_precomp, const float *viewmatrix,
                                 const float *projmatrix, const float *campos, const int32_t *radii,
                                 void *geom_buffer, const void *binning_buffer, int64_t capacity,
                                 const void *image_buffer, const int32_t *status, const float *dL_dout_color,
                                 const float *dL_dout_depth, const int32_t *hit_image, float *dL_dmeans2D,
                                 float *dL_dconic, float *dL_dopacity, float *dL_dcolors, float *dL_dmeans3D,
                                 float *dL_dcov3D, float *dL_dsh, float *dL_dscales, float *dL_drotations,
                                 uint8_t *ever, uint32_t *ever_list, int32_t *ever_count, void *stream_,
                                 const ExtraBlendGrad *extra) {
    cudaStream_t stream = (cudaStream_t)stream_;
    pdl_scope(s ? s->P : 0);
    if (!s || s->P < 0) {
        set_error("dqo_rast_backward: invalid settings");
        return DQO_ERR_INVALID_ARG;
    }
    const int P = s->P;
    if (P == 0) return DQO_OK;
    if (!background || !means3D || !scales || !rotations || !viewmatrix || !projmatrix || !campos || !radii ||
        !geom_buffer || !binning_buffer || !image_buffer || !status || !dL_dout_color || !dL_dout_depth || !hit_image ||
        !dL_dmeans3D) {
        set_error("dqo_rast_backward: null pointer argument");
        return DQO_ERR_INVALID_ARG;
    }
    if (s->M > 0 && shs && !dL_dsh) {
        set_error("dqo_rast_backward: dL_dsh is required when SH coefficients are given");
        return DQO_ERR_INVALID_ARG;
    }
    (void)colors_precomp;
    GeomLayout GL;
    BinLayout BL;
    ImgLayout IL;
    if (make_geom_layout(P, &GL) || make_bin_layout(capacity, &BL)) return DQO_ERR_WORKSPACE;
    make_img_layout(s->W, s->H, &IL);
    char *geom = (char *)geom_buffer;
    const char *bin = (const char *)binning_buffer;
    const char *img = (const char *)image_buffer;
    double *gacc = (double *)(geom + GL.gacc);
    stage_mark(stream, ST_BEGIN_BWD);
    if (!s->geom_clean) {
        DQO_CUDA_CHECK(cudaMemsetAsync(gacc, 0, (size_t)P * DQO_GACC_FLOATS * sizeof(double), stream));
        DQO_CUDA_CHECK(cudaMemsetAsync(geom + GL.touched, 0, (size_t)P, stream));
    }

    const float focal_y = s->H / (2.0f * s->tanfovy);
    const float focal_x = s->W / (2.0f * s->tanfovx);
    RenderBwdArgs ra;
    ra.W = s->W; ra.H = s->H; ra.grid_x = IL.tiles_x;
    ra.fx = focal_x; ra.fy = focal_y; ra.cx = s->cx; ra.cy = s->cy;
    ra.depth_thr = s->depth_threshold; ra.normal_thr = s->normal_threshold;
    ra.ranges = (const uint2 *)(img + IL.ranges);
    ra.point_list = (const uint32_t *)(bin + bin_point_list(BL, IL.T));
    const bool two_phase = s->front_instances > 0;
    ra.ranges_b = two_phase ? (const uint2 *)(img + IL.ranges_b) : nullptr;
    ra.point_list_b = two_phase ? ra.point_list + s->front_instances : nullptr;
    ra.rec = (const float4 *)(geom + GL.rec);
    ra.view = viewmatrix; ra.means3D = means3D; ra.scales = scales; ra.rotations = rotations; ra.bg = background;
    ra.n_contrib = (const uint32_t *)(img + IL.n_contrib);
    ra.final_T = (const float *)(img + IL.final_T);
    ra.hit_geo = (const float *)(img + IL.hit_geo);
    ra.plane = (size_t)IL.T * 256;
    ra.dL_dpix = dL_dout_color; ra.dL_ddepth = dL_dout_depth; ra.hit_image = hit_image; ra.gacc = gacc;
    ra.touched = (uint8_t *)(geom + GL.touched);
    ra.extra_colors = nullptr; ra.cacc = nullptr;
    if (!(extra && extra->only)) {
        launch_pdl(render_backward_kernel<false>, dim3(IL.T), dim3(RB_THREADS), 0, stream, ra);
        DQO_LAUNCH_CHECK("render backward", s->debug, stream);
    }
    if (extra) { // second colour set over the same lists: its image gradient, its colours, its own colour accumulators
        if (!extra->colors || !extra->dL_dpix || !extra->cacc || !extra->dL_dcolors) {
            set_error("dqo_rast_backward: incomplete extra colour set");
            return DQO_ERR_INVALID_ARG;
        }
        RenderBwdArgs rx = ra;
        rx.dL_dpix = extra->dL_dpix; rx.extra_colors = extra->colors; rx.cacc = extra->cacc;
        launch_pdl(render_backward_kernel<true>, dim3(IL.T), dim3(RB_THREADS), 0, stream, rx);
        DQO_LAUNCH_CHECK("render backward (extra colours)", s->debug, stream);
    }
    stage_mark(stream, ST_RENDER_BWD);

    GaussBwdArgs ga;
    ga.P = P; ga.D = s->D; ga.M = s->M;
    ga.scale_modifier = s->scale_modifier; ga.tanfovx = s->tanfovx; ga.tanfovy = s->tanfovy;
    ga.focal_x = focal_x; ga.focal_y = focal_y;
    ga.means3D = means3D; ga.scales = scales; ga.rotations = rotations; ga.shs = shs; ga.f_rest = f_rest; ga.cov3D_precomp = cov3D_precomp;
    ga.view = viewmatrix; ga.proj = projmatrix; ga.campos = campos; ga.radii = radii;
    ga.clamped = (const uint8_t *)(geom + GL.clamped);
    ga.gacc = gacc;
    ga.touched = (uint8_t *)(geom + GL.touched);
    ga.dL_dmeans2D = dL_dmeans2D; ga.dL_dconic = dL_dconic; ga.dL_dopacity = dL_dopacity; ga.dL_dcolors = dL_dcolors;
    ga.dL_dmeans3D = dL_dmeans3D; ga.dL_dcov3D = dL_dcov3D; ga.dL_dsh = (s->M > 0) ? dL_dsh : nullptr;
    ga.dL_dscales = dL_dscales; ga.dL_drot = dL_drotations;
    ga.cacc = extra ? extra->cacc : nullptr;
    ga.dL_dextra = extra ? extra->dL_dcolors : nullptr;
    ga.ever = ever;
    ga.out_nz = (s->geom_clean == 2 && !ever) ? (uint8_t *)(geom + GL.out_nz) : nullptr;
    ga.ever_list = (ever && ever_list && ever_count) ? ever_list : nullptr;
    ga.ever_count = ever_count;
    const bool staged = shs && !f_rest && dL_dsh && s->M == 16 && ((uintptr_t)shs % 16 == 0) && ((uintptr_t)dL_dsh % 16 == 0);
    // persistent outputs (rows that stay zero are not rewritten): blocks compact their touched Gaussians; otherwise every
    // row is written anyway and the blocks keep the thread-per-Gaussian form
    ga.force_dense = (ga.ever != nullptr || ga.out_nz != nullptr) ? 0 : 1;
    // chunks per block: about one wave of blocks (148 SMs x 4) on large maps, where a block finds a few touched Gaussians
    // per chunk and its latency is one chain; one chunk per block on small clouds, where the chains should run side by side
    int chunks = (P + GB_THREADS * 592 - 1) / (GB_THREADS * 592);
    chunks = chunks < 1 ? 1 : (chunks > GB_MAX_CHUNKS ? GB_MAX_CHUNKS : chunks);
    if (ga.force_dense) chunks = 1;
    ga.chunks = chunks;
    const int per_block = GB_THREADS * chunks;
    const int gb_blocks = (P + per_block - 1) / per_block;
    const size_t gb_smem = (size_t)(GB_THREADS / 32) * 32 * GB_ROW_Q * sizeof(float4);
    if (f_rest) {
        if (s->M != 16 || !dL_dsh || (uintptr_t)shs % 16 || (uintptr_t)f_rest % 16 || (uintptr_t)dL_dsh % 16) {
            set_error("split SH input requires M == 16 and 16-byte aligned f_dc / f_rest / dL_dsh");
            return DQO_ERR_INVALID_ARG;
        }
        launch_pdl(gaussian_backward_kernel<2>, dim3(gb_blocks), dim3(GB_THREADS), gb_smem, stream, ga);
    } else if (staged)
        launch_pdl(gaussian_backward_kernel<1>, dim3(gb_blocks), dim3(GB_THREADS), gb_smem, stream, ga);
    else
        launch_pdl(gaussian_backward_kernel<0>, dim3(gb_blocks), dim3(GB_THREADS), 0, stream, ga);
    DQO_LAUNCH_CHECK("gaussian backward", s->debug, stream);
    stage_mark(stream, ST_GAUSS_BWD);
    return DQO_OK;
}

// Backward of dqo_rast_blend_extra: the gradient of a loss on the extra image w.r.t. its colours and -- through alpha and the
// 2-D geometry -- w.r.t. the geometric inputs of the view held by the workspaces, which autograd adds to the main render's.
extern "C" int dqo_rast_blend_extra_backward(const dqo_rast_settings *s, const float *background, const float *colors,
                                             const float *means3D, const float *scales, const float *rotations,
                                             const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix,
                                             const float *campos, const int32_t *radii, void *geom_buffer,
                                             const void *binning_buffer, int64_t capacity, const void *image_buffer,
                                             const int32_t *status, const float *dL_dextra_image, void *color_acc,
                                             float *dL_dcolors, float *dL_dmeans2D, float *dL_dconic, float *dL_dopacity,
                                             float *dL_dmeans3D, float *dL_dcov3D, float *dL_dscales, float *dL_drotations,
                                             void *stream_) {
    if (!s || !colors || !dL_dextra_image || !color_acc || !dL_dcolors || !status) {
        set_error("dqo_rast_blend_extra_backward: null pointer argument");
        return DQO_ERR_INVALID_ARG;
    }
    ExtraBlendGrad xg;
    xg.colors = colors; xg.dL_dpix = dL_dextra_image; xg.cacc = (double *)color_acc; xg.dL_dcolors = dL_dcolors;
    xg.only = 1;
    // (the main image's gradient, the depth gradient and the hit image are not read in this mode: placeholders)
    return rast_backward_impl(s, background, means3D, nullptr, nullptr, nullptr, scales, rotations, cov3D_precomp, viewmatrix,
                              projmatrix, campos, radii, geom_buffer, binning_buffer, capacity, image_buffer, status,
                              dL_dextra_image, dL_dextra_image, status, dL_dmeans2D, dL_dconic, dL_dopacity, nullptr,
                              dL_dmeans3D, dL_dcov3D, nullptr, dL_dscales, dL_drotations, nullptr, nullptr, nullptr,
                              stream_, &xg);
}
