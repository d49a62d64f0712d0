// Fused mapping iteration: one C-ABI call enqueues the whole step of the reference's hot loop
// (SLAM/multiprocess/mapper.py:568-599 + loss_update :799-928: masked L1 colour + depth loss, the attach term :810-829,
// the SSIM term of the mask-less global pass :839-841 and the semantic colour term :877-880; not the normal term, whose
// weight is 0 in every shipped config, nor the gradient-free instance term) on one stream, with no host synchronisation
// and no intermediate tensor owned by the host:
//
//   raw parameters --activate--> rasterize forward --> masked L1 colour/depth loss + image gradients
//                  <-- Adam on the raw parameters <-- activation backward <-- rasterize backward
//
// This is SURVEY.md §8f row 1 ("activation + concat fusion") built on top of the parity-checked kernels: the
// reference's exp / sigmoid / normalize / torch.cat graph (mapper.py:1810-1840, gaussian_pointcloud.py:724-826) and
// the autograd nodes behind it are replaced by two small kernels, the spherical-harmonics tensors f_dc / f_rest are
// read in place by the staged loaders (SHMODE 2) and updated in place by an Adam kernel that walks the merged
// gradient.  The plain operator path (dqo_rast_forward / dqo_rast_backward / dqo_adam_step) stays the parity baseline.
#include "common.cuh"
#include <math.h>

namespace dqo {
int rast_forward_impl(const dqo_rast_settings *s, const float *background, const float *means3D, const float *shs,
                      const float *f_rest, const float *colors_precomp, const float *opacities, const float *scales,
                      const float *rotations, const float *cov3D_precomp, const float *viewmatrix,
                      const float *projmatrix, const float *campos, const int32_t *tile_mask, void *geom_buffer,
                      void *binning_buffer, int64_t capacity, void *image_buffer, int32_t *tile_indices, float *out_color,
                      float *out_depth, int32_t *out_hit_depth, int32_t *out_hit_color, float *out_hit_color_weight,
                      float *out_hit_depth_weight, float *out_T, int32_t *radii, int32_t *n_touched, int32_t *status,
                      void *stream_, void (*pre_hook)(void *, void *), void *hook_ctx, uint8_t *tile_filled);
int masked_l1_loss_impl(int32_t W, int32_t H, const float *image, const float *depth, const int32_t *hit_depth,
                        const float *gt_color, const float *gt_depth, const uint8_t *render_mask, float color_weight,
                        float depth_weight, float depth_err_thres, float *dL_dimage, float *dL_ddepth, float *loss_out,
                        int32_t *counts_out, void *workspace, const int32_t *tile_mask, void *stream_);
int rast_backward_impl(const dqo_rast_settings *s, const float *background, const float *means3D, const float *shs,
                       const float *f_rest, const float *colors_precomp, const float *scales, const float *rotations,
                       const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix, const float *campos,
                       const int32_t *radii, void *geom_buffer, const void *binning_buffer, int64_t capacity,
                       const void *image_buffer, const int32_t *status, const float *dL_dout_color,
                       const float *dL_dout_depth, const int32_t *hit_image, float *dL_dmeans2D, float *dL_dconic,
                       float *dL_dopacity, float *dL_dcolors, float *dL_dmeans3D, float *dL_dcov3D, float *dL_dsh,
                       float *dL_dscales, float *dL_drotations, uint8_t *ever, uint32_t *ever_list, int32_t *ever_count,
                       void *stream_, const ExtraBlendGrad *extra);
int ssim_loss_impl(int32_t W, int32_t H, const float *image, const float *gt_color, float weight, float *dL_dimage,
                   int32_t accumulate, float *loss_out, void *workspace, void *stream_);

struct StepLayout {
    size_t act_opacity, act_scales, act_rot;                    // activated copies [P], [P,3], [P,4]
    size_t geom, binning, image;                                // rasterizer workspaces
    size_t color, depth, hit_depth, hit_color, hit_cw, hit_dw, T, radii, n_touched, tile_indices;
    size_t g_img, g_depth, loss_ws;                             // loss gradients + reduction scratch
    size_t g_means3D, g_sh, g_opacity, g_scales, g_rot;         // activated-space parameter gradients
    size_t adam;                                                // AdamScalars of this step, written on the device
    // optional loss terms: SSIM scratch; semantic image, its gradient image, a dummy depth-gradient image, loss scratch,
    // f64[4] colour-gradient accumulators per Gaussian (kept zero between steps) and the semantic colour gradient [P,3]
    size_t ssim_ws, extra_loss, sem_img, g_sem, g_sem_depth, sem_loss_ws, cacc, g_semantics;
    size_t total;
};
static size_t sbump(size_t &cur, size_t bytes) {
    size_t off = align_up(cur, 256);
    cur = off + bytes;
    return off;
}
static int make_step_layout(int P, int M, int W, int H, int64_t capacity, int terms, StepLayout *L) {
    const size_t n = (size_t)(P > 0 ? P : 1), N = (size_t)W * H;
    const size_t tiles = (size_t)((W + 15) / 16) * ((H + 15) / 16);
    size_t cur = 0;
    L->act_opacity = sbump(cur, n * 4);
    L->act_scales = sbump(cur, n * 12);
    L->act_rot = sbump(cur, n * 16);
    const size_t gb = dqo_rast_geom_bytes(P), bb = dqo_rast_binning_bytes(capacity), ib = dqo_rast_image_bytes(W, H);
    if (!gb || !bb || !ib) return -1;
    L->geom = sbump(cur, gb);
    L->binning = sbump(cur, bb);
    L->image = sbump(cur, ib);
    L->color = sbump(cur, N * 12);
    L->depth = sbump(cur, N * 4);
    L->hit_depth = sbump(cur, N * 4);
    L->hit_color = sbump(cur, N * 4);
    L->hit_cw = sbump(cur, N * 4);
    L->hit_dw = sbump(cur, N * 4);
    L->T = sbump(cur, N * 4);
    L->radii = sbump(cur, n * 4);
    L->n_touched = sbump(cur, n * 4);
    L->tile_indices = sbump(cur, tiles * 4);
    L->g_img = sbump(cur, N * 12);
    L->g_depth = sbump(cur, N * 4);
    L->loss_ws = sbump(cur, dqo_loss_workspace_bytes(W, H));
    L->g_means3D = sbump(cur, n * 12);
    L->g_sh = sbump(cur, n * (size_t)(M > 0 ? M : 1) * 12);
    L->g_opacity = sbump(cur, n * 4);
    L->g_scales = sbump(cur, n * 12);
    L->g_rot = sbump(cur, n * 16);
    L->adam = sbump(cur, 256);
    // the optional terms' regions exist only in workspaces sized for them (DQO_STEP_TERM_*)
    const bool t_ssim = terms & DQO_STEP_TERM_SSIM, t_sem = terms & DQO_STEP_TERM_SEMANTIC;
    L->extra_loss = sbump(cur, 256);
    L->ssim_ws = sbump(cur, t_ssim ? dqo_ssim_workspace_bytes(W, H) : 0);
    L->sem_img = sbump(cur, t_sem ? N * 12 : 0);
    L->g_sem = sbump(cur, t_sem ? N * 12 : 0);
    L->g_sem_depth = sbump(cur, t_sem ? N * 4 : 0);
    L->sem_loss_ws = sbump(cur, t_sem ? dqo_loss_workspace_bytes(W, H) : 0);
    L->cacc = sbump(cur, t_sem ? n * 32 : 0);
    L->g_semantics = sbump(cur, t_sem ? n * 12 : 0);
    L->total = align_up(cur, 256);
    return 0;
}

// exp / sigmoid / normalize of the raw parameters (gaussian_pointcloud.py:20-30: torch.exp, torch.sigmoid,
// torch.nn.functional.normalize with eps = 1e-12)
__global__ void __launch_bounds__(256) activate_kernel(int P, const float *__restrict__ opacity_raw,
                                                       const float *__restrict__ scaling_log,
                                                       const float *__restrict__ rotation_raw, float *opacity, float *scales,
                                                       float *rot) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    opacity[i] = 1.0f / (1.0f + expf(-opacity_raw[i]));
#pragma unroll
    for (int c = 0; c < 3; c++) scales[3 * i + c] = expf(scaling_log[3 * i + c]);
    const float4 r = reinterpret_cast<const float4 *>(rotation_raw)[i];
    const float nrm = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
    const float d = fmaxf(nrm, 1e-12f);
    reinterpret_cast<float4 *>(rot)[i] = make_float4(r.x / d, r.y / d, r.z / d, r.w / d);
}

struct AdamScalars {
    float beta1, beta2, omb1, omb2, bc2_sqrt, eps;
    float step_size[6]; // xyz, f_dc, f_rest, opacity, scaling, rotation
    // attach term (mapper.py:810-829): d/dp of 1000 * mean((p - p0)^2) over the masked rows = attach_grad * (p - p0),
    // and the value itself = attach_val * sum((p - p0)^2); index 0 xyz, 1 scaling, 2 rotation; 0 when disabled
    float attach_grad[3], attach_val[3];
    float attach_logit_thr; // a Gaussian is anchored when sigmoid(init_opacity) < thr
    int skip;               // the forward flagged an instance overflow: gradients are invalid, no update
    float step_size_sem;    // semantic colours
};
__device__ __forceinline__ void adam_update(float &p, float g, float &m, float &v, const AdamScalars &k, float step_size) {
    m = m + k.omb1 * (g - m);
    v = v * k.beta2 + k.omb2 * g * g;
    const float denom = sqrtf(v) / k.bc2_sqrt + k.eps;
    p = p - step_size * (m / denom);
}

// One thread: the step number lives on the device (step_state[0] counts the updates that really happened, step_state[1]
// the steps skipped because the forward overflowed its instance capacity -- sticky until the host resets it), so the
// call takes no per-step host argument and the whole iteration can be replayed from a CUDA graph.
struct PrepareArgs {
    const int *status;
    int *step_state; // may be NULL: host_step is used
    int host_step;
    double beta1, beta2, eps, lr[6];
    double attach_weight;
    float attach_thr;
    const int *attach_count; // NULL: no attach term
    AdamScalars *out;
    // optional loss terms computed into scratch by their own kernels: folded into the report here
    float *loss_out;          // float[8] of the step
    const float *ssim_val;    // {1 - ssim, weight * (1 - ssim)} or NULL
    const float *sem_val;     // {weight * L1, L1} or NULL
    double lr_semantics;
};
__global__ void adam_prepare_kernel(PrepareArgs a) {
    pdl_enter();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int overflow = a.status[DQO_ST_OVERFLOW];
    int step = a.host_step;
    if (a.step_state) {
        if (overflow) {
            a.step_state[1] += 1;
            step = a.step_state[0] + 1;
        } else {
            step = a.step_state[0] + 1;
            a.step_state[0] = step;
        }
    }
    AdamScalars k;
    const double bc1 = 1.0 - pow(a.beta1, (double)step), bc2 = 1.0 - pow(a.beta2, (double)step);
    k.beta1 = (float)a.beta1; k.beta2 = (float)a.beta2; k.omb1 = (float)(1.0 - a.beta1); k.omb2 = (float)(1.0 - a.beta2);
    k.bc2_sqrt = (float)sqrt(bc2); k.eps = (float)a.eps;
    for (int t = 0; t < 6; t++) k.step_size[t] = (float)(a.lr[t] / bc1);
    const int widths[3] = {3, 3, 4};
    const int n_attach = a.attach_count ? *a.attach_count : 0;
    for (int t = 0; t < 3; t++) {
        const double denom = (double)n_attach * widths[t];
        k.attach_grad[t] = n_attach > 0 ? (float)(2.0 * a.attach_weight / denom) : 0.f;
        k.attach_val[t] = n_attach > 0 ? (float)(a.attach_weight / denom) : 0.f;
    }
    k.attach_logit_thr = a.attach_thr;
    k.skip = overflow;
    k.step_size_sem = (float)(a.lr_semantics / bc1);
    *a.out = k;
    float total = a.loss_out[0];
    a.loss_out[4] = a.ssim_val ? a.ssim_val[0] : 0.f;
    if (a.ssim_val) total += a.ssim_val[1];
    a.loss_out[5] = a.sem_val ? a.sem_val[1] : 0.f;
    if (a.sem_val) total += a.sem_val[0];
    a.loss_out[0] = total;
    a.loss_out[6] = a.loss_out[7] = 0.f;
}
// number of Gaussians whose initial opacity is below the attach threshold (mapper.py:810-812)
__global__ void __launch_bounds__(256) attach_count_kernel(int P, const float *__restrict__ init_opacity, float thr, int *count) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = (i < P) && (1.0f / (1.0f + expf(-init_opacity[i])) < thr);
    const unsigned b = __ballot_sync(0xFFFFFFFFu, in);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}
__device__ __forceinline__ bool anchored(const float *__restrict__ init_opacity, long long i, float thr) {
    return (1.0f / (1.0f + expf(-init_opacity[i]))) < thr;
}

// Activation backward + Adam for the 11 geometric parameters of a Gaussian, plus the confidence bump
// (mapper.py:909-910: confidence += 1 where any f_dc gradient is non-zero) and the attach term (mapper.py:810-829):
// Gaussians whose INITIAL opacity is below 0.9 are anchored to their initial xyz / log-scale / raw rotation by
// 1000 * (mse + mse + mse); its gradient 2000 / (n_anchored * width) * (p - p0) is added to the rasterizer's gradient.
struct SmallArgs {
    int P, M;
    float *xyz, *opacity, *scaling, *rotation, *f_dc;
    float *m_xyz, *v_xyz, *m_op, *v_op, *m_sc, *v_sc, *m_rot, *v_rot, *m_dc, *v_dc;
    const float *act_opacity, *act_scales;
    const float *g_means3D, *g_opacity, *g_scales, *g_rot, *g_sh;
    float *confidence;
    const uint8_t *ever; // per-Gaussian "has ever had a non-zero gradient" (nullptr: decide from the values instead)
    const float *init_xyz, *init_scaling, *init_rotation, *init_opacity; // attach reference (all nullptr: no attach term)
    float *attach_out;   // loss_out[3]: value of the attach term (accumulated, zeroed by the loss kernel)
    const AdamScalars *kd; // this step's scalars, written by adam_prepare_kernel
};
__device__ __forceinline__ void block_add(float v, float *out) { // sum over the block, one atomic per block
    __shared__ float s_part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_part[w];
        if (t != 0.f) atomicAdd(out, t);
    }
}
// One parameter group (role) of Gaussian i: 0 xyz, 1 opacity, 2 scaling, 3 rotation, 4 f_dc (+ confidence bump);
// returns its share of the attach value
__device__ __forceinline__ float adam_role(const SmallArgs &a, const AdamScalars &k, int i, int role) {
    float att = 0.f;
    const bool anch = (role == 0 || role == 2 || role == 3) && a.init_opacity && anchored(a.init_opacity, i, k.attach_logit_thr);
    if (role == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int e = 3 * i + c;
            float p = a.xyz[e], m = a.m_xyz[e], v = a.v_xyz[e];
            float g = a.g_means3D[e];
            if (anch) {
                const float d = p - a.init_xyz[e];
                g += k.attach_grad[0] * d;
                att += k.attach_val[0] * d * d;
            }
            adam_update(p, g, m, v, k, k.step_size[0]);
            a.xyz[e] = p; a.m_xyz[e] = m; a.v_xyz[e] = v;
        }
    } else if (role == 1) { // opacity: o = sigmoid(x), dL/dx = g * o * (1 - o)
        const float o = a.act_opacity[i];
        const float g = a.g_opacity[i] * (o * (1.0f - o));
        float p = a.opacity[i], m = a.m_op[i], v = a.v_op[i];
        if (!(g == 0.f && m == 0.f && v == 0.f)) { // zero gradient on zero moments: fixed point (lr_opacity is 0 anyway)
            adam_update(p, g, m, v, k, k.step_size[3]);
            a.opacity[i] = p; a.m_op[i] = m; a.v_op[i] = v;
        }
    } else if (role == 2) {
#pragma unroll
        for (int c = 0; c < 3; c++) { // scale: s = exp(x), dL/dx = g * s
            const int e = 3 * i + c;
            float g = a.g_scales[e] * a.act_scales[e];
            float p = a.scaling[e], m = a.m_sc[e], v = a.v_sc[e];
            if (anch) {
                const float d = p - a.init_scaling[e];
                g += k.attach_grad[1] * d;
                att += k.attach_val[1] * d * d;
            }
            adam_update(p, g, m, v, k, k.step_size[4]);
            a.scaling[e] = p; a.m_sc[e] = m; a.v_sc[e] = v;
        }
    } else if (role == 3) { // rotation: q = r / max(|r|, eps); dL/dr = (g - q (q . g)) / max(|r|, eps)
        float4 r = reinterpret_cast<float4 *>(a.rotation)[i];
        const float4 g = reinterpret_cast<const float4 *>(a.g_rot)[i];
        const float nrm = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
        const float d = fmaxf(nrm, 1e-12f);
        const float qx = r.x / d, qy = r.y / d, qz = r.z / d, qw = r.w / d;
        const float qg = qx * g.x + qy * g.y + qz * g.z + qw * g.w;
        const float inv = (nrm > 1e-12f) ? 1.0f / d : 0.0f;
        float gr[4] = {(g.x - qx * qg) * inv, (g.y - qy * qg) * inv, (g.z - qz * qg) * inv, (g.w - qw * qg) * inv};
        float pr[4] = {r.x, r.y, r.z, r.w};
        if (anch) {
            const float4 r0 = reinterpret_cast<const float4 *>(a.init_rotation)[i];
            const float d0[4] = {r.x - r0.x, r.y - r0.y, r.z - r0.z, r.w - r0.w};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                gr[c] += k.attach_grad[2] * d0[c];
                att += k.attach_val[2] * d0[c] * d0[c];
            }
        }
        float4 m4 = reinterpret_cast<float4 *>(a.m_rot)[i], v4 = reinterpret_cast<float4 *>(a.v_rot)[i];
        float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int c = 0; c < 4; c++) adam_update(pr[c], gr[c], mm[c], vv[c], k, k.step_size[5]);
        reinterpret_cast<float4 *>(a.rotation)[i] = make_float4(pr[0], pr[1], pr[2], pr[3]);
        reinterpret_cast<float4 *>(a.m_rot)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        reinterpret_cast<float4 *>(a.v_rot)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else { // f_dc: gradient = first coefficient of the merged SH gradient
        const float *gs = a.g_sh + (size_t)i * a.M * 3;
        const float g3[3] = {gs[0], gs[1], gs[2]};
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int e = 3 * i + c;
            float p = a.f_dc[e], m = a.m_dc[e], v = a.v_dc[e];
            adam_update(p, g3[c], m, v, k, k.step_size[1]);
            a.f_dc[e] = p; a.m_dc[e] = m; a.v_dc[e] = v;
        }
        if (a.confidence && (fabsf(g3[0]) != 0.f || fabsf(g3[1]) != 0.f || fabsf(g3[2]) != 0.f)) a.confidence[i] += 1.0f;
    }
    return att;
}
// all 14 geometric parameters of Gaussian i
__device__ __forceinline__ float adam_gaussian(const SmallArgs &a, const AdamScalars &k, int i) {
    float att = 0.f;
#pragma unroll
    for (int role = 0; role < 5; role++) att += adam_role(a, k, i, role);
    return att;
}
// one thread per Gaussian over the whole cloud (unaligned tensors, or no `ever` bookkeeping)
__global__ void __launch_bounds__(256) adam_geometry_kernel(SmallArgs a) {
    pdl_enter();
    const AdamScalars k = *a.kd;
    if (k.skip) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float att = 0.f;
    if (i < a.P && !(a.ever && !a.ever[i])) att = adam_gaussian(a, k, i);
    if (a.init_opacity) block_add(att, a.attach_out);
}
// The optimiser over the COMPACT LIST of Gaussians that have ever received a gradient (dqo_map_params.ever_list,
// appended to by the backward pass): for a single keyframe that is 10-20 % of a 1 M-Gaussian map, and walking the whole
// cloud just to skip the rest (one flag byte per Gaussian and tensor) cost more than the updates themselves.  Persistent
// grid: the list length is only known on the device.
struct ListArgs {
    SmallArgs s;
    const uint32_t *list;
    const int *count;
    float *f_rest, *m_rest, *v_rest;
};
__global__ void __launch_bounds__(256) adam_list_kernel(ListArgs a) {
    pdl_enter();
    const AdamScalars k = *a.s.kd;
    if (k.skip) return;
    const int n = *a.count;
    float att = 0.f;
    // one work item per (parameter group, listed Gaussian), group-major: a warp works on one group of 32 consecutive list
    // entries, five times the loads in flight of a thread-per-Gaussian loop and no load queued behind another group's stores
    const long long items = 5ll * n;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += (long long)gridDim.x * blockDim.x) {
        const int role = (int)(w / n), l = (int)(w - (long long)role * n);
        att += adam_role(a.s, k, (int)a.list[l], role);
    }
    if (a.s.init_opacity) block_add(att, a.s.attach_out);
}
// f_rest [P,45] of the listed Gaussians: consecutive threads walk the 45 coefficients of one Gaussian (180-byte runs);
// the gradient comes from the merged [P,16,3] layout (row stride 48, offset 3)
__global__ void __launch_bounds__(256) adam_rest_list_kernel(ListArgs a) {
    pdl_enter();
    const AdamScalars k = *a.s.kd;
    if (k.skip) return;
    const unsigned total = (unsigned)(*a.count) * 45u;
    const float ss = k.step_size[2];
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const unsigned l = e / 45u, c = e - l * 45u;
        const size_t i = a.list[l];
        const size_t pe = i * 45 + c;
        const float g = __ldg(&a.s.g_sh[i * 48 + 3 + c]);
        float m = a.m_rest[pe], v = a.v_rest[pe];
        if (g == 0.f && m == 0.f && v == 0.f) continue;
        float p = a.f_rest[pe];
        adam_update(p, g, m, v, k, ss);
        a.f_rest[pe] = p; a.m_rest[pe] = m; a.v_rest[pe] = v;
    }
}

// Same update with every flat tensor streamed in 128-bit accesses: blocks are split into five roles by index range.
//   role 0  xyz      [3P]  Adam on g_means3D (+ attach)                   float4 chunks of the flat array
//   role 1  scaling  [3P]  g * exp(x) (the saved activation) (+ attach)   float4 chunks
//   role 2  opacity  [P]   g * o (1 - o)                                  float4 chunks
//   role 3  rotation [P,4] normalisation backward (+ attach), one Gaussian per thread (one float4 per tensor)
//   role 4  f_dc     [P,3] gradient gathered from the merged SH gradient (stride 3M), + confidence bump
// Requires 16-byte aligned tensors (checked by the caller; adam_geometry_kernel is the fallback).
struct FlatArgs {
    SmallArgs s;
    long long n4_vec3, n4_scalar; // float4 chunks of a [3P] / [P] array
    unsigned nb_vec3, nb_scalar, nb_gauss;
};
// zero gradient on zero moments is a fixed point of Adam (m' = v' = 0, p' = p - step * 0 / eps = p): such elements
// need no parameter read, no arithmetic (IEEE sqrt / div take their slow paths on zeros) and no write-back
__device__ __forceinline__ bool all_zero(const float4 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f && a.w == 0.f; }
__device__ __forceinline__ void adam4(float4 &p, const float4 g, float4 &m, float4 &v, const AdamScalars &k, float ss) {
    adam_update(p.x, g.x, m.x, v.x, k, ss);
    adam_update(p.y, g.y, m.y, v.y, k, ss);
    adam_update(p.z, g.z, m.z, v.z, k, ss);
    adam_update(p.w, g.w, m.w, v.w, k, ss);
}
// flat Adam over [from, n) for the (< 4) elements a float4 sweep leaves over; returns the attach value of those elements
__device__ __forceinline__ float adam_tail(float *p, float *m, float *v, const float *g, const float *act, int mode,
                                           long long from, long long n, const AdamScalars &k, float ss, const uint8_t *ever,
                                           int row_width, const float *init, const float *init_opacity, int attach_slot) {
    float att = 0.f;
    for (long long e = from + threadIdx.x; e < n; e += blockDim.x) {
        if (ever && !ever[e / row_width]) continue;
        float gg = g[e];
        if (mode == 1) gg *= act[e];
        if (mode == 2) gg *= act[e] * (1.0f - act[e]);
        float pp = p[e], mm = m[e], vv = v[e];
        if (init && anchored(init_opacity, e / row_width, k.attach_logit_thr)) {
            const float d = pp - init[e];
            gg += k.attach_grad[attach_slot] * d;
            att += k.attach_val[attach_slot] * d * d;
        }
        adam_update(pp, gg, mm, vv, k, ss);
        p[e] = pp; m[e] = mm; v[e] = vv;
    }
    return att;
}
// `b` is a virtual block index (one role per block).  Returns this thread's share of the attach value.
__device__ __forceinline__ float adam_flat_body(const FlatArgs &fa, const AdamScalars &k, unsigned b) {
    const SmallArgs &a = fa.s;
    float att = 0.f;
    if (b < 2 * fa.nb_vec3) { // roles 0 / 1
        const bool sc = b >= fa.nb_vec3;
        if (sc) b -= fa.nb_vec3;
        float *P_ = sc ? a.scaling : a.xyz, *M_ = sc ? a.m_sc : a.m_xyz, *V_ = sc ? a.v_sc : a.v_xyz;
        const float *G_ = sc ? a.g_scales : a.g_means3D;
        const float *I_ = a.init_opacity ? (sc ? a.init_scaling : a.init_xyz) : nullptr;
        const int slot = sc ? 1 : 0;
        const float ss = k.step_size[sc ? 4 : 0];
        const long long q = (long long)b * blockDim.x + threadIdx.x;
        if (q < fa.n4_vec3) {
            const long long ra = (4 * q) / 3, rb = (4 * q + 3) / 3;
            bool ea = true, eb = true;
            if (a.ever) { // gradients of never-touched Gaussians were not written: neither read nor used
                ea = a.ever[ra] != 0;
                eb = a.ever[rb] != 0;
            }
            if (ea || eb) {
                bool e[4], an[4] = {false, false, false, false};
#pragma unroll
                for (int c = 0; c < 4; c++) e[c] = ((4 * q + c) / 3 == ra) ? ea : eb;
                float4 g = reinterpret_cast<const float4 *>(G_)[q];
                g.x = e[0] ? g.x : 0.f; g.y = e[1] ? g.y : 0.f; g.z = e[2] ? g.z : 0.f; g.w = e[3] ? g.w : 0.f;
                float4 m = reinterpret_cast<float4 *>(M_)[q], v = reinterpret_cast<float4 *>(V_)[q];
                bool any_an = false;
                if (I_) {
                    const bool aa = ea && anchored(a.init_opacity, ra, k.attach_logit_thr);
                    const bool ab = eb && (rb == ra ? aa : anchored(a.init_opacity, rb, k.attach_logit_thr));
#pragma unroll
                    for (int c = 0; c < 4; c++) an[c] = ((4 * q + c) / 3 == ra) ? aa : ab;
                    any_an = aa || ab;
                }
                // a parameter only leaves its initial value through an update, which leaves non-zero moments behind: with
                // zero gradient AND zero moments the attach gradient is zero as well
                if (!(all_zero(g) && all_zero(m) && all_zero(v))) {
                    if (sc) {
                        const float4 ex = reinterpret_cast<const float4 *>(a.act_scales)[q];
                        g.x *= ex.x; g.y *= ex.y; g.z *= ex.z; g.w *= ex.w;
                    }
                    float4 p = reinterpret_cast<float4 *>(P_)[q];
                    if (any_an) {
                        const float4 p0 = reinterpret_cast<const float4 *>(I_)[q];
                        const float d[4] = {p.x - p0.x, p.y - p0.y, p.z - p0.z, p.w - p0.w};
                        float ga[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int c = 0; c < 4; c++)
                            if (an[c]) {
                                ga[c] = k.attach_grad[slot] * d[c];
                                att += k.attach_val[slot] * d[c] * d[c];
                            }
                        g.x += ga[0]; g.y += ga[1]; g.z += ga[2]; g.w += ga[3];
                    }
                    adam4(p, g, m, v, k, ss);
                    reinterpret_cast<float4 *>(P_)[q] = p;
                    reinterpret_cast<float4 *>(M_)[q] = m;
                    reinterpret_cast<float4 *>(V_)[q] = v;
                }
            }
        }
        if (b == 0)
            att += adam_tail(P_, M_, V_, G_, a.act_scales, sc ? 1 : 0, fa.n4_vec3 * 4, 3ll * a.P, k, ss, a.ever, 3, I_,
                             a.init_opacity, slot);
        return att;
    }
    b -= 2 * fa.nb_vec3;
    if (b < fa.nb_scalar) { // role 2
        const float ss = k.step_size[3];
        const long long q = (long long)b * blockDim.x + threadIdx.x;
        if (q < fa.n4_scalar) {
            uint32_t e4 = 0x01010101u;
            if (a.ever) e4 = reinterpret_cast<const uint32_t *>(a.ever)[q];
            if (e4 != 0) {
                float4 g = reinterpret_cast<const float4 *>(a.g_opacity)[q];
                g.x = (e4 & 0xFFu) ? g.x : 0.f; g.y = (e4 & 0xFF00u) ? g.y : 0.f;
                g.z = (e4 & 0xFF0000u) ? g.z : 0.f; g.w = (e4 & 0xFF000000u) ? g.w : 0.f;
                float4 m = reinterpret_cast<float4 *>(a.m_op)[q], v = reinterpret_cast<float4 *>(a.v_op)[q];
                if (!(all_zero(g) && all_zero(m) && all_zero(v))) {
                    const float4 o = reinterpret_cast<const float4 *>(a.act_opacity)[q];
                    g.x *= o.x * (1.0f - o.x); g.y *= o.y * (1.0f - o.y); g.z *= o.z * (1.0f - o.z); g.w *= o.w * (1.0f - o.w);
                    float4 p = reinterpret_cast<float4 *>(a.opacity)[q];
                    adam4(p, g, m, v, k, ss);
                    reinterpret_cast<float4 *>(a.opacity)[q] = p;
                    reinterpret_cast<float4 *>(a.m_op)[q] = m;
                    reinterpret_cast<float4 *>(a.v_op)[q] = v;
                }
            }
        }
        if (b == 0)
            adam_tail(a.opacity, a.m_op, a.v_op, a.g_opacity, a.act_opacity, 2, fa.n4_scalar * 4, a.P, k, ss, a.ever, 1, nullptr,
                      nullptr, 0);
        return 0.f;
    }
    b -= fa.nb_scalar;
    const bool dc = b >= fa.nb_gauss;
    if (dc) b -= fa.nb_gauss;
    const int i = (int)(b * blockDim.x + threadIdx.x);
    if (i >= a.P) return 0.f;
    if (a.ever && !a.ever[i]) return 0.f;
    if (!dc) { // role 3: rotation, q = r / max(|r|, eps); dL/dr = (g - q (q . g)) / max(|r|, eps)
        const float4 g = reinterpret_cast<const float4 *>(a.g_rot)[i];
        float4 m = reinterpret_cast<float4 *>(a.m_rot)[i], v = reinterpret_cast<float4 *>(a.v_rot)[i];
        if (all_zero(g) && all_zero(m) && all_zero(v)) return 0.f; // normalisation backward of a zero gradient is zero
        float4 r = reinterpret_cast<float4 *>(a.rotation)[i];
        const float nrm = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
        const float d = fmaxf(nrm, 1e-12f);
        const float qx = r.x / d, qy = r.y / d, qz = r.z / d, qw = r.w / d;
        const float qg = qx * g.x + qy * g.y + qz * g.z + qw * g.w;
        const float inv = (nrm > 1e-12f) ? 1.0f / d : 0.0f;
        float4 gr = make_float4((g.x - qx * qg) * inv, (g.y - qy * qg) * inv, (g.z - qz * qg) * inv, (g.w - qw * qg) * inv);
        if (a.init_opacity && anchored(a.init_opacity, i, k.attach_logit_thr)) {
            const float4 r0 = reinterpret_cast<const float4 *>(a.init_rotation)[i];
            const float d0[4] = {r.x - r0.x, r.y - r0.y, r.z - r0.z, r.w - r0.w};
            gr.x += k.attach_grad[2] * d0[0]; gr.y += k.attach_grad[2] * d0[1];
            gr.z += k.attach_grad[2] * d0[2]; gr.w += k.attach_grad[2] * d0[3];
            att = k.attach_val[2] * (d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2] + d0[3] * d0[3]);
        }
        adam4(r, gr, m, v, k, k.step_size[5]);
        reinterpret_cast<float4 *>(a.rotation)[i] = r;
        reinterpret_cast<float4 *>(a.m_rot)[i] = m;
        reinterpret_cast<float4 *>(a.v_rot)[i] = v;
        return att;
    }
    // role 4: f_dc (first coefficient of the merged SH gradient) + confidence
    const float *gs = a.g_sh + (size_t)i * a.M * 3;
    const float g3[3] = {gs[0], gs[1], gs[2]};
    float m3[3], v3[3];
    bool zero = true;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        m3[c] = a.m_dc[3 * i + c];
        v3[c] = a.v_dc[3 * i + c];
        zero &= (g3[c] == 0.f && m3[c] == 0.f && v3[c] == 0.f);
    }
    if (zero) return 0.f; // also no confidence bump: every f_dc gradient is zero
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int e = 3 * i + c;
        float p = a.f_dc[e];
        adam_update(p, g3[c], m3[c], v3[c], k, k.step_size[1]);
        a.f_dc[e] = p; a.m_dc[e] = m3[c]; a.v_dc[e] = v3[c];
    }
    if (a.confidence && (fabsf(g3[0]) != 0.f || fabsf(g3[1]) != 0.f || fabsf(g3[2]) != 0.f)) a.confidence[i] += 1.0f;
    return 0.f;
}
__global__ void __launch_bounds__(256) adam_flat_kernel(FlatArgs fa) {
    pdl_enter();
    const AdamScalars k = *fa.s.kd;
    if (k.skip) return;
    const float att = adam_flat_body(fa, k, blockIdx.x);
    if (fa.s.init_opacity) block_add(att, fa.s.attach_out);
}

// Adam for f_rest [P,45]: 128-bit accesses on the parameter and its two moments (6 of the 7 streams), the gradient is
// gathered from the merged [P,16,3] layout (row stride 48, offset 3)
struct RestAdamArgs {
    long long n4; // number of float4 chunks of f_rest
    float *f_rest, *m_rest, *v_rest;
    const float *g_sh;
    const uint8_t *ever;
    const AdamScalars *kd;
};
constexpr int ADAM_REST_CHUNKS = 8; // float4 chunks per thread: most are skipped after reading one flag byte
template <typename IndexT>
__device__ __forceinline__ void adam_rest_body(const RestAdamArgs &a, const AdamScalars &k, IndexT q) {
    const IndexT e0 = q * 4;
    bool ea = true, eb = true;
    const IndexT ra = e0 / 45;
    if (a.ever) {
        ea = a.ever[ra] != 0;
        eb = a.ever[(e0 + 3) / 45] != 0;
        if (!(ea || eb)) return;
    }
    float4 m = reinterpret_cast<float4 *>(a.m_rest)[q];
    float4 v = reinterpret_cast<float4 *>(a.v_rest)[q];
    float g[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const IndexT e = e0 + c;
        const IndexT row = e / 45;
        const int col = (int)(e - row * 45);
        g[c] = ((row == ra) ? ea : eb) ? __ldg(&a.g_sh[(size_t)row * 48 + 3 + col]) : 0.f;
    }
    // zero gradient on zero moments is a fixed point of Adam (m' = v' = 0, p' = p - step * 0 / eps = p): nothing to
    // read or write for Gaussians that no keyframe has touched yet (most of the map for any single view)
    if (g[0] == 0.f && g[1] == 0.f && g[2] == 0.f && g[3] == 0.f && m.x == 0.f && m.y == 0.f && m.z == 0.f && m.w == 0.f &&
        v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f)
        return;
    float4 p = reinterpret_cast<float4 *>(a.f_rest)[q];
    const float ss = k.step_size[2];
    adam_update(p.x, g[0], m.x, v.x, k, ss);
    adam_update(p.y, g[1], m.y, v.y, k, ss);
    adam_update(p.z, g[2], m.z, v.z, k, ss);
    adam_update(p.w, g[3], m.w, v.w, k, ss);
    reinterpret_cast<float4 *>(a.f_rest)[q] = p;
    reinterpret_cast<float4 *>(a.m_rest)[q] = m;
    reinterpret_cast<float4 *>(a.v_rest)[q] = v;
}
__global__ void __launch_bounds__(256) adam_rest_kernel(RestAdamArgs a) {
    pdl_enter();
    const AdamScalars k = *a.kd;
    if (k.skip) return;
    const long long base = (long long)blockIdx.x * (256 * ADAM_REST_CHUNKS) + threadIdx.x;
    const bool small = a.n4 * 4 < (1ll << 31); // 32-bit index arithmetic (the / 45 is the hot instruction of a skipped chunk)
#pragma unroll 1
    for (int c = 0; c < ADAM_REST_CHUNKS; c++) {
        const long long q = base + (long long)c * 256;
        if (q >= a.n4) return;
        if (small)
            adam_rest_body<unsigned>(a, k, (unsigned)q);
        else
            adam_rest_body<long long>(a, k, q);
    }
}
__global__ void adam_rest_tail_kernel(long long begin, long long end, RestAdamArgs a) {
    pdl_enter();
    const long long e = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const AdamScalars k = *a.kd;
    if (e >= end || k.skip) return;
    const long long row = e / 45;
    if (a.ever && !a.ever[row]) return;
    const int col = (int)(e - row * 45);
    float p = a.f_rest[e], m = a.m_rest[e], v = a.v_rest[e];
    adam_update(p, a.g_sh[row * 48 + 3 + col], m, v, k, k.step_size[2]);
    a.f_rest[e] = p; a.m_rest[e] = m; a.v_rest[e] = v;
}

// Adam for the semantic colours [P,3] (parameter group "semantics_color", gaussian_pointcloud.py:371-378): over the
// compact list of ever-touched Gaussians when there is one, else over the cloud with the flag byte
struct SemAdamArgs {
    int P;
    float *p, *m, *v;
    const float *g;
    const uint8_t *ever;
    const uint32_t *list;
    const int *count;
    const AdamScalars *kd;
};
__global__ void __launch_bounds__(256) adam_semantics_kernel(SemAdamArgs a) {
    pdl_enter();
    const AdamScalars k = *a.kd;
    if (k.skip) return;
    const long long n = a.list ? 3ll * (*a.count) : 3ll * a.P;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (long long)gridDim.x * blockDim.x) {
        const long long l = w / 3;
        const int c = (int)(w - 3 * l);
        const long long i = a.list ? (long long)a.list[l] : l;
        if (!a.list && a.ever && !a.ever[i]) continue;
        const long long e = 3 * i + c;
        const float g = a.g[e];
        float m = a.m[e], v = a.v[e];
        if (g == 0.f && m == 0.f && v == 0.f) continue;
        float p = a.p[e];
        adam_update(p, g, m, v, k, k.step_size_sem);
        a.p[e] = p; a.m[e] = m; a.v[e] = v;
    }
}

} // namespace dqo

using namespace dqo;

extern "C" size_t dqo_mapping_step_workspace_bytes(int32_t P, int32_t M, int32_t W, int32_t H, int64_t capacity,
                                                   int32_t terms) {
    StepLayout L;
    if (make_step_layout(P, M, W, H, capacity, terms, &L)) return 0;
    return L.total;
}

extern "C" int dqo_rast_geom_init(int32_t P, void *geom_buffer, void *stream);
extern "C" int dqo_mapping_step_workspace_init(int32_t P, int32_t M, int32_t W, int32_t H, int64_t capacity, int32_t terms,
                                               void *workspace, void *stream) {
    StepLayout L;
    if (!workspace || make_step_layout(P, M, W, H, capacity, terms, &L)) {
        set_error("dqo_mapping_step_workspace_init: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    const size_t tiles = (size_t)((W + 15) / 16) * ((H + 15) / 16);
    DQO_CUDA_CHECK(cudaMemsetAsync((char *)workspace + L.tile_indices, 0, tiles * 4, (cudaStream_t)stream));
    if (terms & DQO_STEP_TERM_SEMANTIC)
        DQO_CUDA_CHECK(cudaMemsetAsync((char *)workspace + L.cacc, 0, (size_t)(P > 0 ? P : 1) * 32, (cudaStream_t)stream));
    return dqo_rast_geom_init(P, (char *)workspace + L.geom, stream);
}

extern "C" int dqo_mapping_step(const dqo_rast_settings *s, const dqo_map_params *p, const dqo_keyframe *kf, int32_t step,
                                double beta1, double beta2, double eps, void *workspace, int64_t capacity,
                                float *loss_out, int32_t *counts_out, int32_t *status, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!s || !p || !kf || !workspace || !loss_out || !counts_out || !status || (step < 1 && !p->step_state) || s->P <= 0) {
        set_error("dqo_mapping_step: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    pdl_scope(s->P);
    nvtx_push("dqo_mapping_step");
    if (!(s->M == 16 || s->M == 1)) {
        set_error("dqo_mapping_step supports M == 16 (SH degree 3 storage) or M == 1");
        return DQO_ERR_INVALID_ARG;
    }
    for (int t = 0; t < 6; t++)
        if (!p->param[t] || !p->exp_avg[t] || !p->exp_avg_sq[t]) {
            if (t == 2 && s->M == 1) continue; // no f_rest at degree 0
            set_error("dqo_mapping_step: parameter tensor %d or its Adam state is NULL", t);
            return DQO_ERR_INVALID_ARG;
        }
    const int P = s->P, M = s->M, W = s->W, H = s->H;
    StepLayout L;
    if (make_step_layout(P, M, W, H, capacity, p->workspace_terms, &L)) return DQO_ERR_WORKSPACE;
    char *ws = (char *)workspace;
    float *act_op = (float *)(ws + L.act_opacity), *act_sc = (float *)(ws + L.act_scales), *act_rot = (float *)(ws + L.act_rot);
    float *xyz = p->param[0], *f_dc = p->param[1], *f_rest = (M == 16) ? p->param[2] : nullptr;
    const int nb = (P + 255) / 256;
    // the activations are launched by the forward pass right behind the fork of its depth sort (which needs positions only)
    struct ActivateCtx {
        int P, nb;
        const float *op, *sc, *rot;
        float *act_op, *act_sc, *act_rot;
    } actx = {P, nb, p->param[3], p->param[4], p->param[5], act_op, act_sc, act_rot};
    auto activate_hook = [](void *ctx, void *st) {
        const ActivateCtx *c = (const ActivateCtx *)ctx;
        launch_pdl(activate_kernel, dim3(c->nb), dim3(256), 0, (cudaStream_t)st, c->P, c->op, c->sc, c->rot, c->act_op, c->act_sc,
                   c->act_rot);
        note_launch();
    };

    float *color = (float *)(ws + L.color), *depth = (float *)(ws + L.depth);
    int32_t *hit_depth = (int32_t *)(ws + L.hit_depth), *radii = (int32_t *)(ws + L.radii);
    int rc = rast_forward_impl(s, kf->background, xyz, f_dc, f_rest, nullptr, act_op, act_sc, act_rot, nullptr, kf->viewmatrix,
                               kf->projmatrix, kf->campos, kf->tile_mask, ws + L.geom, ws + L.binning, capacity,
                               ws + L.image, /* tile list: nobody reads it in the step */ nullptr, color, depth, hit_depth,
                               (int32_t *)(ws + L.hit_color), (float *)(ws + L.hit_cw), (float *)(ws + L.hit_dw),
                               (float *)(ws + L.T), radii, (int32_t *)(ws + L.n_touched), status, stream_, activate_hook,
                               &actx, (uint8_t *)(ws + L.tile_indices) /* per-tile "holds fill values" flags */);
    if (rc) return rc;
    float *g_img = (float *)(ws + L.g_img), *g_depth = (float *)(ws + L.g_depth);
    rc = masked_l1_loss_impl(W, H, color, depth, hit_depth, kf->gt_color, kf->gt_depth, kf->render_mask, kf->color_weight,
                             kf->depth_weight, kf->depth_err_thres, g_img, g_depth, loss_out, counts_out, ws + L.loss_ws,
                             kf->tile_mask, stream_);
    if (rc) return rc;
    // optional terms.  SSIM: like the reference only in the mask-less pass (mapper.py:839-841); its gradient is added to
    // the colour-gradient image the backward blend reads.
    float *extra_loss = (float *)(ws + L.extra_loss);
    const bool use_ssim = kf->ssim_weight > 0.f && !kf->render_mask;
    if (use_ssim && !(p->workspace_terms & DQO_STEP_TERM_SSIM)) {
        set_error("dqo_mapping_step: the workspace was not sized for the SSIM term (dqo_map_params.workspace_terms)");
        return DQO_ERR_WORKSPACE;
    }
    if (use_ssim) {
        rc = ssim_loss_impl(W, H, color, kf->gt_color, kf->ssim_weight, g_img, 1, extra_loss, ws + L.ssim_ws, stream_);
        if (rc) return rc;
    }
    // semantic term (mapper.py:877-880): the semantic image is one more blend over the lists of the main render
    const bool use_sem = kf->gt_semantic != nullptr && kf->semantic_weight > 0.f;
    ExtraBlendGrad xg;
    if (use_sem && !(p->workspace_terms & DQO_STEP_TERM_SEMANTIC)) {
        set_error("dqo_mapping_step: the workspace was not sized for the semantic term (dqo_map_params.workspace_terms)");
        return DQO_ERR_WORKSPACE;
    }
    if (use_sem) {
        if (!p->semantics || !p->semantics_exp_avg || !p->semantics_exp_avg_sq) {
            set_error("dqo_mapping_step: the semantic term needs dqo_map_params.semantics and its Adam state");
            return DQO_ERR_INVALID_ARG;
        }
        float *sem_img = (float *)(ws + L.sem_img), *g_sem = (float *)(ws + L.g_sem);
        rc = dqo_rast_blend_extra(s, kf->background, p->semantics, ws + L.geom, ws + L.binning, capacity, ws + L.image, status,
                                  sem_img, stream_);
        if (rc) return rc;
        rc = masked_l1_loss_impl(W, H, sem_img, depth, hit_depth, kf->gt_semantic, kf->gt_depth, kf->render_mask,
                                 kf->semantic_weight, 0.f, kf->depth_err_thres, g_sem, (float *)(ws + L.g_sem_depth),
                                 extra_loss + 4, (int32_t *)(extra_loss + 8), ws + L.sem_loss_ws, kf->tile_mask, stream_);
        if (rc) return rc;
        xg.colors = p->semantics; xg.dL_dpix = g_sem; xg.cacc = (double *)(ws + L.cacc);
        xg.dL_dcolors = (float *)(ws + L.g_semantics);
        xg.only = 0;
    }
    float *g_means3D = (float *)(ws + L.g_means3D), *g_sh = (float *)(ws + L.g_sh), *g_op = (float *)(ws + L.g_opacity);
    float *g_sc = (float *)(ws + L.g_scales), *g_rot = (float *)(ws + L.g_rot);
    rc = rast_backward_impl(s, kf->background, xyz, f_dc, f_rest, nullptr, act_sc, act_rot, nullptr, kf->viewmatrix,
                            kf->projmatrix, kf->campos, radii, ws + L.geom, ws + L.binning, capacity, ws + L.image, status,
                            g_img, g_depth, hit_depth, nullptr, nullptr, g_op, nullptr, g_means3D, nullptr, g_sh, g_sc,
                            g_rot, p->ever, p->ever_list, p->ever_count, stream_, use_sem ? &xg : nullptr);
    if (rc) return rc;

    // step number, bias corrections, attach scales and the overflow decision: one thread on the device
    AdamScalars *kd = (AdamScalars *)(ws + L.adam);
    const bool attach = p->init_opacity != nullptr;
    if (attach && (!p->init_xyz || !p->init_scaling || !p->init_rotation || !p->attach_count)) {
        set_error("dqo_mapping_step: the attach term needs init_xyz / init_scaling / init_rotation / init_opacity and attach_count");
        return DQO_ERR_INVALID_ARG;
    }
    {
        PrepareArgs pa;
        pa.status = status; pa.step_state = p->step_state; pa.host_step = step;
        pa.beta1 = beta1; pa.beta2 = beta2; pa.eps = eps;
        for (int t = 0; t < 6; t++) pa.lr[t] = p->lr[t];
        pa.attach_weight = attach ? (double)p->attach_weight : 0.0;
        pa.attach_thr = p->attach_opacity_thres;
        pa.attach_count = attach ? p->attach_count : nullptr;
        pa.out = kd;
        pa.loss_out = loss_out;
        pa.ssim_val = use_ssim ? extra_loss : nullptr;
        pa.sem_val = use_sem ? extra_loss + 4 : nullptr;
        pa.lr_semantics = use_sem ? p->lr_semantics : 0.0;
        launch_pdl(adam_prepare_kernel, dim3(1), dim3(32), 0, stream, pa);
        DQO_LAUNCH_CHECK("adam prepare", s->debug, stream);
    }
    SmallArgs sa;
    sa.P = P; sa.M = M; sa.xyz = xyz; sa.opacity = p->param[3]; sa.scaling = p->param[4]; sa.rotation = p->param[5];
    sa.f_dc = f_dc; sa.m_dc = p->exp_avg[1]; sa.v_dc = p->exp_avg_sq[1];
    sa.m_xyz = p->exp_avg[0]; sa.v_xyz = p->exp_avg_sq[0]; sa.m_op = p->exp_avg[3]; sa.v_op = p->exp_avg_sq[3];
    sa.m_sc = p->exp_avg[4]; sa.v_sc = p->exp_avg_sq[4]; sa.m_rot = p->exp_avg[5]; sa.v_rot = p->exp_avg_sq[5];
    sa.act_opacity = act_op; sa.act_scales = act_sc; sa.g_means3D = g_means3D; sa.g_opacity = g_op; sa.g_scales = g_sc;
    sa.g_rot = g_rot; sa.g_sh = g_sh; sa.confidence = p->confidence; sa.ever = p->ever; sa.kd = kd;
    sa.init_xyz = attach ? p->init_xyz : nullptr; sa.init_scaling = attach ? p->init_scaling : nullptr;
    sa.init_rotation = attach ? p->init_rotation : nullptr; sa.init_opacity = attach ? p->init_opacity : nullptr;
    sa.attach_out = loss_out + 3;
    bool aligned = true;
    {
        const void *ptrs[] = {sa.xyz, sa.m_xyz, sa.v_xyz, sa.scaling, sa.m_sc, sa.v_sc, sa.opacity, sa.m_op, sa.v_op,
                              sa.rotation, sa.m_rot, sa.v_rot, g_means3D, g_sc, g_op, g_rot, act_op, act_sc,
                              sa.init_xyz, sa.init_scaling, sa.init_rotation};
        for (const void *q : ptrs) aligned &= ((uintptr_t)q % 16 == 0);
        aligned &= ((uintptr_t)p->ever % 4 == 0);
    }
    const bool use_list = p->ever && p->ever_list && p->ever_count && (long long)P * 45 < (1ll << 32);
    if (use_sem) {
        SemAdamArgs se;
        se.P = P; se.p = p->semantics; se.m = p->semantics_exp_avg; se.v = p->semantics_exp_avg_sq;
        se.g = (const float *)(ws + L.g_semantics); se.ever = p->ever;
        se.list = use_list ? p->ever_list : nullptr; se.count = p->ever_count; se.kd = kd;
        launch_pdl(adam_semantics_kernel, dim3(148 * 4), dim3(256), 0, stream, se);
        DQO_LAUNCH_CHECK("adam (semantic colours)", s->debug, stream);
    }
    if (use_list) {
        ListArgs la;
        la.s = sa; la.list = p->ever_list; la.count = p->ever_count;
        la.f_rest = f_rest; la.m_rest = p->exp_avg[2]; la.v_rest = p->exp_avg_sq[2];
        const int blocks = 148 * 8;
        launch_pdl(adam_list_kernel, dim3(blocks), dim3(256), 0, stream, la);
        if (M == 16) launch_pdl(adam_rest_list_kernel, dim3(148 * 8), dim3(256), 0, stream, la);
        DQO_LAUNCH_CHECK("adam (listed Gaussians)", s->debug, stream);
        note_launch(M == 16 ? 1 : 0);
        nvtx_pop();
        return DQO_OK;
    }
    if (aligned) {
        FlatArgs fa;
        fa.s = sa;
        fa.n4_vec3 = 3ll * P / 4;
        fa.n4_scalar = P / 4;
        fa.nb_vec3 = (unsigned)((fa.n4_vec3 + 255) / 256);
        if (fa.nb_vec3 == 0) fa.nb_vec3 = 1; // the tail loop lives in block 0 of each role
        fa.nb_scalar = (unsigned)((fa.n4_scalar + 255) / 256);
        if (fa.nb_scalar == 0) fa.nb_scalar = 1;
        fa.nb_gauss = (unsigned)nb;
        const unsigned vblocks = 2 * fa.nb_vec3 + fa.nb_scalar + 2 * fa.nb_gauss;
        launch_pdl(adam_flat_kernel, dim3(vblocks), dim3(256), 0, stream, fa);
    } else {
        launch_pdl(adam_geometry_kernel, dim3(nb), dim3(256), 0, stream, sa);
    }
    DQO_LAUNCH_CHECK("adam geometry", s->debug, stream);
    if (M == 16) {
        RestAdamArgs ra;
        const long long total = (long long)P * 45;
        ra.n4 = total / 4; ra.f_rest = f_rest; ra.m_rest = p->exp_avg[2]; ra.v_rest = p->exp_avg_sq[2]; ra.g_sh = g_sh;
        ra.ever = p->ever; ra.kd = kd;
        if (ra.n4 > 0)
            launch_pdl(adam_rest_kernel, dim3((unsigned)((ra.n4 + 256 * ADAM_REST_CHUNKS - 1) / (256 * ADAM_REST_CHUNKS))), dim3(256), 0, stream, ra);
        if (total % 4) launch_pdl(adam_rest_tail_kernel, dim3(1), dim3(32), 0, stream, ra.n4 * 4, total, ra);
        DQO_LAUNCH_CHECK("adam f_rest", s->debug, stream);
    }
    nvtx_pop();
    return DQO_OK;
}

// Number of Gaussians the attach term anchors (sigmoid(init_opacity) < thres), counted once per optimisation window
// (the reference recomputes `attach_mask.sum()` every iteration from the same init_stat: mapper.py:810-812).
extern "C" int dqo_attach_count(int32_t P, const float *init_opacity, float opacity_thres, int32_t *count, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P < 0 || !count || (P > 0 && !init_opacity)) {
        set_error("dqo_attach_count: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    DQO_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), stream));
    if (P == 0) return DQO_OK;
    launch_pdl(attach_count_kernel, dim3((P + 255) / 256), dim3(256), 0, stream, P, init_opacity, opacity_thres, count);
    DQO_LAUNCH_CHECK("attach count", 0, stream);
    return DQO_OK;
}

// Read-only views into the step workspace for callers that want the rendered images of the last step
extern "C" int dqo_mapping_step_outputs(int32_t P, int32_t M, int32_t W, int32_t H, int64_t capacity, void *workspace,
                                        float **color, float **depth, int32_t **hit_depth, float **T_map) {
    StepLayout L; // (the optional regions lie behind everything this returns)
    if (!workspace || make_step_layout(P, M, W, H, capacity, 0, &L)) return DQO_ERR_WORKSPACE;
    char *ws = (char *)workspace;
    if (color) *color = (float *)(ws + L.color);
    if (depth) *depth = (float *)(ws + L.depth);
    if (hit_depth) *hit_depth = (int32_t *)(ws + L.hit_depth);
    if (T_map) *T_map = (float *)(ws + L.T);
    return DQO_OK;
}
