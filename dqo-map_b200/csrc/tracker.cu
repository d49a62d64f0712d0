// Tracker front-end: depth pyramid, vertex / normal maps and the point-to-plane ICP
// Gauss-Newton iterations that consume them.  Reference: SLAM/icp.py:16-130 (ICP), :230-330 (skew, damping, se(3)
// exponential, solve), :342-360 (ImagePyramids), SLAM/utils.py:65-125 (compute_vertex_map, feature_gradient,
// compute_normal_map), :542-559 (pyramids).  The reference runs each ICP iteration as ~40 torch kernels plus a
// device->host->device round trip for the 6x6 inverse (icp.py:313-326); here an iteration is two launches (one fused
// association + Jacobian + 27-sum reduction pass over the image, one single-thread solve / pose update) and the pose
// never leaves the device.
#include "common.cuh"

namespace dqo {

// ---- depth pyramid: MaxPool2d(1 << level, 1 << level) of the full-resolution depth (icp.py:346-349) -----------------
__global__ void __launch_bounds__(256) depth_maxpool_kernel(int W, int H, int shift, const float *__restrict__ depth, float *out) {
    const int Wo = W >> shift, Ho = H >> shift;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Wo * Ho) return;
    const int xo = p % Wo, yo = p / Wo, k = 1 << shift;
    float m = -INFINITY;
    for (int dy = 0; dy < k; dy++)
        for (int dx = 0; dx < k; dx++) m = fmaxf(m, depth[(size_t)(yo * k + dy) * W + xo * k + dx]);
    out[p] = m;
}

// ---- vertex and normal maps ---------------------------------------------------------------------------------------
// min / max of the depth (compute_normal_map masks depth <= min and depth >= max, utils.py:120-121); depth >= 0, so the
// float bit patterns order like unsigned integers
__global__ void __launch_bounds__(256) depth_minmax_kernel(int N, const float *__restrict__ depth, int stride, unsigned *mm) {
    unsigned lo = 0xFFFFFFFFu, hi = 0u;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        const unsigned b = __float_as_uint(fmaxf(depth[(size_t)p * stride], 0.0f));
        lo = min(lo, b);
        hi = max(hi, b);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[0], lo);
        atomicMax(&mm[1], hi);
    }
}
__device__ __forceinline__ float3 vertex_at(const float *__restrict__ depth, int W, int H, int x, int y, float fx, float fy,
                                            float cx, float cy) {
    x = min(max(x, 0), W - 1); // replicate padding of feature_gradient (utils.py:91)
    y = min(max(y, 0), H - 1);
    const float d = depth[(size_t)y * W + x];
    return make_float3(fmul(fdiv(fsub((float)x, cx), fx), d), fmul(fdiv(fsub((float)y, cy), fy), d), d);
}
__device__ __forceinline__ float3 vertex_load(const float *__restrict__ vmap, int W, int H, int x, int y) {
    x = min(max(x, 0), W - 1);
    y = min(max(y, 0), H - 1);
    const size_t o = 3 * ((size_t)y * W + x);
    return make_float3(vmap[o], vmap[o + 1], vmap[o + 2]);
}
// FROM_DEPTH: vertices are computed from the depth image (and written out); otherwise read from `vertex_in`.
template <bool FROM_DEPTH>
__global__ void __launch_bounds__(256)
    vertex_normal_kernel(int W, int H, const float *__restrict__ depth, const float *__restrict__ vertex_in, float fx, float fy,
                         float cx, float cy, const unsigned *__restrict__ mm, float *vertex, float *normal) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int x = p % W, y = p / W;
    float3 v[3][3];
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++)
            v[dy][dx] = FROM_DEPTH ? vertex_at(depth, W, H, x + dx - 1, y + dy - 1, fx, fy, cx, cy)
                                   : vertex_load(vertex_in, W, H, x + dx - 1, y + dy - 1);
    const float3 c = v[1][1];
    if (FROM_DEPTH) {
        vertex[3 * (size_t)p] = c.x;
        vertex[3 * (size_t)p + 1] = c.y;
        vertex[3 * (size_t)p + 2] = c.z;
    }
    if (!normal) return;
    // Sobel responses (utils.py:86-93): wx = [[-1,0,1],[-2,0,2],[-1,0,1]], wy = [[-1,-2,-1],[0,0,0],[1,2,1]]
#define SOBX(f) ((v[0][2].f - v[0][0].f) + 2.0f * (v[1][2].f - v[1][0].f) + (v[2][2].f - v[2][0].f))
#define SOBY(f) ((v[2][0].f - v[0][0].f) + 2.0f * (v[2][1].f - v[0][1].f) + (v[2][2].f - v[0][2].f))
    const float3 gx = make_float3(SOBX(x), SOBX(y), SOBX(z)), gy = make_float3(SOBY(x), SOBY(y), SOBY(z));
#undef SOBX
#undef SOBY
    // normal = cross(img_dy, img_dx) / (|.| + 1e-8) (utils.py:113-117)
    float3 n = make_float3(gy.y * gx.z - gy.z * gx.y, gy.z * gx.x - gy.x * gx.z, gy.x * gx.y - gy.y * gx.x);
    const float inv = 1.0f / (sqrtf(n.x * n.x + n.y * n.y + n.z * n.z) + 1e-8f);
    const float dmin = __uint_as_float(mm[0]), dmax = __uint_as_float(mm[1]);
    const bool invalid = (c.z <= dmin) || (c.z >= dmax);
    normal[3 * (size_t)p] = invalid ? 0.f : n.x * inv;
    normal[3 * (size_t)p + 1] = invalid ? 0.f : n.y * inv;
    normal[3 * (size_t)p + 2] = invalid ? 0.f : n.z * inv;
}

// ---- ICP ----------------------------------------------------------------------------------------------------------------
// workspace: double acc[28] (21 upper-triangle JtJ, 6 JtR, 1 valid count) followed by float pose scratch
constexpr int ICP_ACC = 28;
struct IcpArgs {
    int W, H;
    const float *v0, *v1, *n0, *n1; // [H,W,3]
    const float *pose;              // [4,4] row-major, device
    float fx, fy, cx, cy, dist_thr, normal_thr;
    double *acc;
};

// One pass: transform the template vertices / normals with the current pose, projective data association (nearest
// sample of frame 1 = grid_sample(nearest, align_corners=True, border), icp.py:128-145), point-to-plane residual and
// Jacobian [v x n, n] (icp.py:83-96), validity mask (icp.py:99-102), and the sums J^T J, J^T r (icp.py:106-121).
__global__ void __launch_bounds__(256) icp_accumulate_kernel(IcpArgs a) {
    __shared__ float s_part[8][ICP_ACC];
    const int N = a.W * a.H;
    float R[9], t[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 3; c++) R[3 * r + c] = a.pose[4 * r + c];
        t[r] = a.pose[4 * r + 3];
    }
    float s[ICP_ACC];
#pragma unroll
    for (int k = 0; k < ICP_ACC; k++) s[k] = 0.f;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        const float3 v0 = make_float3(a.v0[3 * (size_t)p], a.v0[3 * (size_t)p + 1], a.v0[3 * (size_t)p + 2]);
        const float3 n0 = make_float3(a.n0[3 * (size_t)p], a.n0[3 * (size_t)p + 1], a.n0[3 * (size_t)p + 2]);
        const float3 q = make_float3(R[0] * v0.x + R[1] * v0.y + R[2] * v0.z + t[0], R[3] * v0.x + R[4] * v0.y + R[5] * v0.z + t[1],
                                     R[6] * v0.x + R[7] * v0.y + R[8] * v0.z + t[2]);
        const float3 m = make_float3(R[0] * n0.x + R[1] * n0.y + R[2] * n0.z, R[3] * n0.x + R[4] * n0.y + R[5] * n0.z,
                                     R[6] * n0.x + R[7] * n0.y + R[8] * n0.z);
        const float u = (q.x / q.z) * a.fx + a.cx, v = (q.y / q.z) * a.fy + a.cy;
        const bool inview = (u > 0.f) && (u < (float)(a.W - 1)) && (v > 0.f) && (v < (float)(a.H - 1));
        // nearest sample with border clamp; NaN coordinates (z == 0) land on pixel 0 and are rejected by the masks
        int xi = (int)nearbyintf(fminf(fmaxf(u, 0.f), (float)(a.W - 1)));
        int yi = (int)nearbyintf(fminf(fmaxf(v, 0.f), (float)(a.H - 1)));
        if (!(u == u)) xi = 0;
        if (!(v == v)) yi = 0;
        const size_t o = 3 * ((size_t)yi * a.W + xi);
        const float3 v1 = make_float3(a.v1[o], a.v1[o + 1], a.v1[o + 2]);
        const float3 n1 = make_float3(a.n1[o], a.n1[o + 1], a.n1[o + 2]);
        const float3 d = make_float3(q.x - v1.x, q.y - v1.y, q.z - v1.z);
        const bool ok = inview && !(sqrtf(d.x * d.x + d.y * d.y + d.z * d.z) > a.dist_thr) && (v0.z > 0.f) && (v1.z > 0.f) &&
                        (m.x * n1.x + m.y * n1.y + m.z * n1.z > a.normal_thr);
        if (!ok) continue;
        const float r = n1.x * d.x + n1.y * d.y + n1.z * d.z;
        const float J[6] = {q.y * n1.z - q.z * n1.y, q.z * n1.x - q.x * n1.z, q.x * n1.y - q.y * n1.x, n1.x, n1.y, n1.z};
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = i; j < 6; j++) s[k++] += J[i] * J[j];
#pragma unroll
        for (int i = 0; i < 6; i++) s[21 + i] += J[i] * r;
        s[27] += 1.0f;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < ICP_ACC; k++) {
        float x = s[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
        if (lane == 0) s_part[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < ICP_ACC) {
        double x = 0.0;
        for (int w = 0; w < 8; w++) x += (double)s_part[w][threadIdx.x];
        atomicAdd(&a.acc[threadIdx.x], x);
    }
}

// Single thread: damping (lev_mar_H, icp.py:248-256), solve H xi = -J^T r (icp.py:313-335; Gaussian elimination with
// partial pivoting in double instead of a host-side torch.inverse), se(3) exponential (icp.py:272-310), pose update
// pose <- exp(xi) pose; then clears the accumulators for the next iteration.
__global__ void icp_solve_kernel(double *acc, float damping, float *pose, float *valid_ratio, int n_pixels) {
    double A[6][7];
    int k = 0;
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) {
            A[i][j] = A[j][i] = acc[k++];
        }
    double trace = 0.0;
    for (int i = 0; i < 6; i++) trace += A[i][i];
    for (int i = 0; i < 6; i++) {
        A[i][i] += trace * (double)damping;
        A[i][6] = -acc[21 + i];
    }
    if (valid_ratio) *valid_ratio = (float)(acc[27] / (double)n_pixels);
    for (int i = 0; i < ICP_ACC; i++) acc[i] = 0.0;
    bool singular = false;
    for (int c = 0; c < 6; c++) {
        int piv = c;
        for (int r = c + 1; r < 6; r++)
            if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (fabs(A[piv][c]) < 1e-300) {
            singular = true;
            break;
        }
        if (piv != c)
            for (int j = 0; j < 7; j++) {
                const double tmp = A[c][j];
                A[c][j] = A[piv][j];
                A[piv][j] = tmp;
            }
        for (int r = c + 1; r < 6; r++) {
            const double f = A[r][c] / A[c][c];
            for (int j = c; j < 7; j++) A[r][j] -= f * A[c][j];
        }
    }
    if (singular) return; // no valid correspondence: the pose is left unchanged
    double xi[6];
    for (int r = 5; r >= 0; r--) {
        double x = A[r][6];
        for (int j = r + 1; j < 6; j++) x -= A[r][j] * xi[j];
        xi[r] = x / A[r][r];
    }
    const float w0 = (float)xi[0], w1 = (float)xi[1], w2 = (float)xi[2];
    const float vx = (float)xi[3], vy = (float)xi[4], vz = (float)xi[5];
    const float Wh[9] = {0.f, -w2, w1, w2, 0.f, -w0, -w1, w0, 0.f};
    float W2[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) W2[3 * r + c] = Wh[3 * r] * Wh[c] + Wh[3 * r + 1] * Wh[3 + c] + Wh[3 * r + 2] * Wh[6 + c];
    const float theta = sqrtf(w0 * w0 + w1 * w1 + w2 * w2);
    float E[9], Jm[9];
    for (int i = 0; i < 9; i++) E[i] = Jm[i] = (i % 4 == 0) ? 1.f : 0.f;
    if (theta > 1e-8f) {
        const float t2 = theta * theta, t3 = t2 * theta, st = sinf(theta), ct = cosf(theta);
        const float k1 = (1.f - ct) / t2, k2 = (theta - st) / t3;
        for (int i = 0; i < 9; i++) {
            E[i] += Wh[i] * st / theta + W2[i] * (1.f - ct) / t2;
            Jm[i] += k1 * Wh[i] + k2 * W2[i];
        }
    }
    float T[16] = {E[0], E[1], E[2], Jm[0] * vx + Jm[1] * vy + Jm[2] * vz,
                   E[3], E[4], E[5], Jm[3] * vx + Jm[4] * vy + Jm[5] * vz,
                   E[6], E[7], E[8], Jm[6] * vx + Jm[7] * vy + Jm[8] * vz,
                   0.f, 0.f, 0.f, 1.f};
    float P[16];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            float x = 0.f;
            for (int q = 0; q < 4; q++) x += T[4 * r + q] * pose[4 * q + c];
            P[4 * r + c] = x;
        }
    for (int i = 0; i < 16; i++) pose[i] = P[i];
}

static bool bad_img(int W, int H) { return W <= 0 || H <= 0; }

} // namespace dqo

using namespace dqo;

extern "C" int dqo_depth_maxpool(int32_t W, int32_t H, int32_t level, const float *depth, float *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bad_img(W, H) || level < 0 || level > 8 || !depth || !out || (W >> level) <= 0 || (H >> level) <= 0) {
        set_error("dqo_depth_maxpool: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    const int n = (W >> level) * (H >> level);
    depth_maxpool_kernel<<<(n + 255) / 256, 256, 0, stream>>>(W, H, level, depth, out);
    DQO_LAUNCH_CHECK("depth maxpool", 0, stream);
    return DQO_OK;
}

extern "C" size_t dqo_vertex_normal_workspace_bytes(void) { return 256; }

extern "C" int dqo_vertex_normal_map(int32_t W, int32_t H, const float *depth, float fx, float fy, float cx, float cy,
                                     float *vertex, float *normal, void *workspace, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bad_img(W, H) || !depth || !vertex || (normal && !workspace)) {
        set_error("dqo_vertex_normal_map: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    unsigned *mm = (unsigned *)workspace;
    const int N = W * H;
    if (normal) {
        const unsigned init[2] = {0xFFFFFFFFu, 0u};
        DQO_CUDA_CHECK(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, stream));
        depth_minmax_kernel<<<min((N + 255) / 256, 592), 256, 0, stream>>>(N, depth, 1, mm);
        DQO_LAUNCH_CHECK("depth min/max", 0, stream);
    }
    vertex_normal_kernel<true><<<(N + 255) / 256, 256, 0, stream>>>(W, H, depth, nullptr, fx, fy, cx, cy, mm, vertex, normal);
    DQO_LAUNCH_CHECK("vertex / normal map", 0, stream);
    return DQO_OK;
}

extern "C" int dqo_normal_map(int32_t W, int32_t H, const float *vertex, float *normal, void *workspace, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bad_img(W, H) || !vertex || !normal || !workspace) {
        set_error("dqo_normal_map: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    unsigned *mm = (unsigned *)workspace;
    const int N = W * H;
    const unsigned init[2] = {0xFFFFFFFFu, 0u};
    DQO_CUDA_CHECK(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    depth_minmax_kernel<<<min((N + 255) / 256, 592), 256, 0, stream>>>(N, vertex + 2, 3, mm);
    DQO_LAUNCH_CHECK("depth min/max", 0, stream);
    vertex_normal_kernel<false><<<(N + 255) / 256, 256, 0, stream>>>(W, H, nullptr, vertex, 0.f, 0.f, 0.f, 0.f, mm, nullptr, normal);
    DQO_LAUNCH_CHECK("normal map", 0, stream);
    return DQO_OK;
}

extern "C" size_t dqo_icp_workspace_bytes(void) { return align_up(ICP_ACC * sizeof(double), 256); }

extern "C" int dqo_icp_level(int32_t W, int32_t H, const float *vertex0, const float *vertex1, const float *normal0,
                             const float *normal1, float fx, float fy, float cx, float cy, float distance_threshold,
                             float normal_threshold_cos, float damping, int32_t iterations, float *pose10,
                             float *valid_ratio, void *workspace, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bad_img(W, H) || !vertex0 || !vertex1 || !normal0 || !normal1 || !pose10 || !workspace || iterations < 0) {
        set_error("dqo_icp_level: invalid argument");
        return DQO_ERR_INVALID_ARG;
    }
    IcpArgs a;
    a.W = W; a.H = H; a.v0 = vertex0; a.v1 = vertex1; a.n0 = normal0; a.n1 = normal1; a.pose = pose10;
    a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.dist_thr = distance_threshold; a.normal_thr = normal_threshold_cos;
    a.acc = (double *)workspace;
    DQO_CUDA_CHECK(cudaMemsetAsync(a.acc, 0, ICP_ACC * sizeof(double), stream));
    const int N = W * H;
    const int blocks = min((N + 255) / 256, 148 * 4);
    for (int it = 0; it < iterations; it++) {
        icp_accumulate_kernel<<<blocks, 256, 0, stream>>>(a);
        DQO_LAUNCH_CHECK("icp accumulate", 0, stream);
        icp_solve_kernel<<<1, 1, 0, stream>>>(a.acc, damping, pose10, valid_ratio, N);
        DQO_LAUNCH_CHECK("icp solve", 0, stream);
    }
    return DQO_OK;
}
