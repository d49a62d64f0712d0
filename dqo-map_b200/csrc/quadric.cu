// Batched dual-quadric kernels (reference SLAM/multiprocess/quadrics.py).
//
// The reference handles objects one at a time in Python: numpy for construction/projection and, for the
// refinement, ~40 tiny CUDA launches plus a torch.linalg.eig call per Adam iteration per object
// (quadrics.py:2245-2295).  Here one thread owns one object and a single launch runs the whole batch
// (all 20 iterations of the IoU-Adam loop included) with closed-form 2x2 eigen-analysis and analytic gradients.
#include "common.cuh"
#include <math.h>

namespace dqo {

// ---- Object.__init__ single-view construction (quadrics.py:451-487), fp64 like numpy ----
__global__ void quadric_init_kernel(int n, const double *__restrict__ bboxes, const double *__restrict__ dstat,
                                    const double *__restrict__ K, const double *__restrict__ Rts, double *axes,
                                    double *Rout, double *center) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *bb = bboxes + 4 * i;
    const double avg = dstat[2 * i], diff = dstat[2 * i + 1];
    const double *Rt = Rts + 12 * i; // [3,4] row-major
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double bcx = (bb[0] + bb[2]) / 2, bcy = (bb[1] + bb[3]) / 2;
    const double u = (bcx - cx) / fx, v = (bcy - cy) / fy;
    const double cc[3] = {u * avg, v * avg, avg};
    // center_world = Rcw^T c_cam - Rcw^T tcw
    double cw[3];
    for (int r = 0; r < 3; r++) {
        double a = 0, b = 0;
        for (int k = 0; k < 3; k++) {
            a += Rt[4 * k + r] * cc[k];
            b += Rt[4 * k + r] * Rt[4 * k + 3];
        }
        cw[r] = a + (-b);
    }
    const double nrm = sqrt(cc[0] * cc[0] + cc[1] * cc[1] + cc[2] * cc[2]);
    const double zc[3] = {cc[0] / nrm, cc[1] / nrm, cc[2] / nrm};
    // xc = cross(-up, zc), up = (0,-1,0)
    double xc[3] = {1.0 * zc[2] - 0.0 * zc[1], 0.0 * zc[0] - 0.0 * zc[2], 0.0 * zc[1] - 1.0 * zc[0]};
    const double xn = sqrt(xc[0] * xc[0] + xc[1] * xc[1] + xc[2] * xc[2]);
    xc[0] /= xn; xc[1] /= xn; xc[2] /= xn;
    const double yc[3] = {zc[1] * xc[2] - zc[2] * xc[1], zc[2] * xc[0] - zc[0] * xc[2], zc[0] * xc[1] - zc[1] * xc[0]};
    const double rc[3][3] = {{xc[0], yc[0], zc[0]}, {xc[1], yc[1], zc[1]}, {xc[2], yc[2], zc[2]}};
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double a = 0;
            for (int k = 0; k < 3; k++) a += Rt[4 * k + r] * rc[k][c];
            Rout[9 * i + 3 * r + c] = a;
        }
    const double w_img = bb[2] - bb[0], h_img = bb[3] - bb[1];
    axes[3 * i] = w_img * avg / fx * 0.5;
    axes[3 * i + 1] = h_img * avg / fy * 0.5;
    axes[3 * i + 2] = diff * 0.5;
    center[3 * i] = cw[0];
    center[3 * i + 1] = cw[1];
    center[3 * i + 2] = cw[2];
}

// ---- dual quadric, projection and bounding box (quadrics.py:388-425,148-248) ----
template <typename T>
__device__ __forceinline__ void build_dual_quadric(const T *ax, const T *R, const T *c, T Q[4][4]) {
    const T D[3] = {ax[0] * ax[0], ax[1] * ax[1], ax[2] * ax[2]};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            T s = 0;
            for (int k = 0; k < 3; k++) s += R[3 * i + k] * D[k] * R[3 * j + k];
            Q[i][j] = s - c[i] * c[j];
        }
    for (int i = 0; i < 3; i++) {
        Q[i][3] = -c[i];
        Q[3][i] = -c[i];
    }
    Q[3][3] = -1;
}
template <typename T>
__device__ __forceinline__ void project_quadric(const T Q[4][4], const T *P, T C[3][3]) {
    T PQ[3][4];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) {
            T s = 0;
            for (int k = 0; k < 4; k++) s += P[4 * i + k] * Q[k][j];
            PQ[i][j] = s;
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            T s = 0;
            for (int k = 0; k < 4; k++) s += PQ[i][k] * P[4 * j + k];
            C[i][j] = s;
        }
    for (int i = 0; i < 3; i++)
        for (int j = i + 1; j < 3; j++) {
            const T m = (T)0.5 * (C[i][j] + C[j][i]);
            C[i][j] = C[j][i] = m;
        }
}

__global__ void quadric_project_kernel(int n, const double *__restrict__ axes, const double *__restrict__ R,
                                       const double *__restrict__ center, const double *__restrict__ Ps, double *bbox,
                                       double *ellipse) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double Q[4][4], C[3][3];
    build_dual_quadric<double>(axes + 3 * i, R + 9 * i, center + 3 * i, Q);
    project_quadric<double>(Q, Ps + 12 * i, C);
    const double nrm = -C[2][2];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) C[r][c] /= nrm;
    const double cx = -C[0][2], cy = -C[1][2];
    const double a = C[0][0] + cx * cx, b = C[0][1] + cx * cy, d = C[1][1] + cy * cy;
    // symmetric 2x2 eigen-decomposition, eigenvalues ascending like numpy.linalg.eigh
    const double m = 0.5 * (a + d), del = 0.5 * (a - d);
    const double r = sqrt(del * del + b * b);
    const double l0 = m - r, l1 = m + r;
    // eigenvector of l0
    double vx, vy;
    if (r == 0.0) {
        vx = 1.0;
        vy = 0.0;
    } else if (fabs(b) > 0.0) {
        vx = b;
        vy = l0 - a;
        const double nn = sqrt(vx * vx + vy * vy);
        vx /= nn;
        vy /= nn;
    } else if (a <= d) {
        vx = 1.0;
        vy = 0.0;
    } else {
        vx = 0.0;
        vy = 1.0;
    }
    if (vx < 0.0 || (vx == 0.0 && vy < 0.0)) {
        vx = -vx;
        vy = -vy;
    }
    const double ax0 = sqrt(fabs(l0)), ax1 = sqrt(fabs(l1));
    const double angle = atan2(vy, vx);
    const double co = cos(angle), si = sin(angle);
    const double xmax = sqrt(ax0 * ax0 * co * co + ax1 * ax1 * si * si);
    const double ymax = sqrt(ax0 * ax0 * si * si + ax1 * ax1 * co * co);
    if (bbox) {
        bbox[4 * i] = cx - xmax;
        bbox[4 * i + 1] = cy - ymax;
        bbox[4 * i + 2] = cx + xmax;
        bbox[4 * i + 3] = cy + ymax;
    }
    if (ellipse) {
        ellipse[5 * i] = ax0;
        ellipse[5 * i + 1] = ax1;
        ellipse[5 * i + 2] = angle;
        ellipse[5 * i + 3] = cx;
        ellipse[5 * i + 4] = cy;
    }
}

// ---- IoU-Adam refinement (quadrics.py:2144-2298) in fp32 ----
// extent^2 along x and y of the centred dual conic [[a,b],[b,d]] with |eigenvalue| semantics
// (Ellipse_tensor: axes = sqrt(|eig|), quadrics.py:2053-2075) and its gradient.
__device__ __forceinline__ void extent2_and_grad(float a, float b, float d, float *X, float *Y, float gX[3], float gY[3]) {
    const float m = 0.5f * (a + d), del = 0.5f * (a - d);
    const float r = sqrtf(del * del + b * b);
    const float lp = m + r, lm = m - r;
    // derivatives of m, del, r w.r.t. (a, b, d)
    const float dm[3] = {0.5f, 0.f, 0.5f}, dd[3] = {0.5f, 0.f, -0.5f};
    float dr[3] = {0.f, 0.f, 0.f};
    if (r > 0.f) {
        dr[0] = del * 0.5f / r;
        dr[1] = b / r;
        dr[2] = -del * 0.5f / r;
    }
    if (lp >= 0.f && lm >= 0.f) { // ellipse: X = a, Y = d
        *X = m + del;
        *Y = m - del;
        for (int k = 0; k < 3; k++) {
            gX[k] = dm[k] + dd[k];
            gY[k] = dm[k] - dd[k];
        }
    } else if (lp < 0.f && lm < 0.f) {
        *X = -(m + del);
        *Y = -(m - del);
        for (int k = 0; k < 3; k++) {
            gX[k] = -(dm[k] + dd[k]);
            gY[k] = -(dm[k] - dd[k]);
        }
    } else { // mixed signs: X = r + del*m/r, Y = r - del*m/r
        const float q = (r > 0.f) ? del * m / r : 0.f;
        *X = r + q;
        *Y = r - q;
        for (int k = 0; k < 3; k++) {
            const float dq = (r > 0.f) ? (dd[k] * m + del * dm[k]) / r - del * m / (r * r) * dr[k] : 0.f;
            gX[k] = dr[k] + dq;
            gY[k] = dr[k] - dq;
        }
    }
}

struct RefineArgs {
    int n, iters, max_views;
    const int *n_views;
    const float *obs, *Ps;
    const int *choice;
    double lr[3]; // axes, center, R
    float *axes, *R, *center, *last_loss;
};

__global__ void quadric_refine_kernel(RefineArgs A) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= A.n) return;
    float p[15]; // axes[3], center[3], R[9]
    for (int k = 0; k < 3; k++) p[k] = A.axes[3 * o + k];
    for (int k = 0; k < 3; k++) p[3 + k] = A.center[3 * o + k];
    for (int k = 0; k < 9; k++) p[6 + k] = A.R[9 * o + k];
    float mom[15], var[15];
    for (int k = 0; k < 15; k++) mom[k] = var[k] = 0.f;
    int step = 0;
    float loss_v = 0.f;
    const int nv = A.n_views[o];
    const double beta1 = 0.9, beta2 = 0.999;
    const float b2f = (float)beta2, omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2), eps = 1e-15f;
    for (int it = 0; it < A.iters; it++) {
        int k = A.choice[(size_t)o * A.iters + it];
        if (k < 0) k += nv; // python negative index: -1 = latest observation
        if (k < 0 || k >= nv) continue;
        const float *P = A.Ps + ((size_t)o * A.max_views + k) * 12;
        const float *ob = A.obs + ((size_t)o * A.max_views + k) * 4;
        // ---------------- forward ----------------
        float Q[4][4], C[3][3];
        build_dual_quadric<float>(p, p + 6, p + 3, Q);
        project_quadric<float>(Q, P, C);
        const float c22 = C[2][2];
        const float cx = C[0][2] / c22, cy = C[1][2] / c22;
        const float a = -C[0][0] / c22 + cx * cx, b = -C[0][1] / c22 + cx * cy, d = -C[1][1] / c22 + cy * cy;
        float X, Y, gX[3], gY[3];
        extent2_and_grad(a, b, d, &X, &Y, gX, gY);
        const float xmax = sqrtf(X), ymax = sqrtf(Y);
        const float bx0 = cx - xmax, by0 = cy - ymax, bx1 = cx + xmax, by1 = cy + ymax;
        // bboxes_iou(obs, pred) with python min/max tie rules (quadrics.py:285-290)
        const bool min_x_pred = bx1 < ob[2], max_x_pred = bx0 > ob[0];
        const bool min_y_pred = by1 < ob[3], max_y_pred = by0 > ob[1];
        const float iw_raw = (min_x_pred ? bx1 : ob[2]) - (max_x_pred ? bx0 : ob[0]);
        const float ih_raw = (min_y_pred ? by1 : ob[3]) - (max_y_pred ? by0 : ob[1]);
        const bool wz = 0.f > iw_raw, hz = 0.f > ih_raw;
        const float iw = wz ? 0.f : iw_raw, ih = hz ? 0.f : ih_raw;
        const float inter = iw * ih;
        const float area_o = (ob[2] - ob[0]) * (ob[3] - ob[1]);
        const float area_p = (bx1 - bx0) * (by1 - by0);
        const float uni = area_o + area_p - inter;
        const float iou = inter / uni;
        loss_v = 1.0f - iou;
        if (loss_v == 1.0f) continue; // "Loss is 1": iteration skipped, optimiser state untouched
        // ---------------- backward ----------------
        // d loss / d (inter, area_p)
        const float dinter = -(1.f / uni + inter / (uni * uni));
        const float darea_p = inter / (uni * uni);
        float g_bx0 = 0, g_by0 = 0, g_bx1 = 0, g_by1 = 0;
        const float diw = wz ? 0.f : dinter * ih, dih = hz ? 0.f : dinter * iw;
        if (min_x_pred) g_bx1 += diw;
        if (max_x_pred) g_bx0 -= diw;
        if (min_y_pred) g_by1 += dih;
        if (max_y_pred) g_by0 -= dih;
        g_bx1 += darea_p * (by1 - by0);
        g_bx0 -= darea_p * (by1 - by0);
        g_by1 += darea_p * (bx1 - bx0);
        g_by0 -= darea_p * (bx1 - bx0);
        float g_cx = g_bx0 + g_bx1, g_cy = g_by0 + g_by1;
        const float g_xmax = g_bx1 - g_bx0, g_ymax = g_by1 - g_by0;
        const float g_X = g_xmax * 0.5f / xmax, g_Y = g_ymax * 0.5f / ymax;
        const float g_a = g_X * gX[0] + g_Y * gY[0], g_b = g_X * gX[1] + g_Y * gY[1], g_d = g_X * gX[2] + g_Y * gY[2];
        g_cx += g_a * 2.f * cx + g_b * cy;
        g_cy += g_d * 2.f * cy + g_b * cx;
        // unique-variable gradients of the symmetric C: c00,c01,c02,c11,c12,c22
        const float g00 = -g_a / c22, g01 = -g_b / c22, g11 = -g_d / c22;
        const float g02 = g_cx / c22, g12 = g_cy / c22;
        const float g22 = (g_a * C[0][0] + g_b * C[0][1] + g_d * C[1][1]) / (c22 * c22) -
                          (g_cx * C[0][2] + g_cy * C[1][2]) / (c22 * c22);
        const float G[3][3] = {{g00, 0.5f * g01, 0.5f * g02}, {0.5f * g01, g11, 0.5f * g12}, {0.5f * g02, 0.5f * g12, g22}};
        // H = P^T G P  (4x4)
        float GP[3][4], H[4][4];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 4; j++) {
                float s = 0;
                for (int q = 0; q < 3; q++) s += G[i][q] * P[4 * q + j];
                GP[i][j] = s;
            }
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                float s = 0;
                for (int q = 0; q < 3; q++) s += P[4 * q + i] * GP[q][j];
                H[i][j] = s;
            }
        float grad[15];
        const float *ax = p, *cen = p + 3, *R = p + 6;
        // axes: 2 a_k (R^T H3 R)_kk ; R: 2 H3 R D ; centre: -2 H3 c - 2 H[:3,3]
        float HR[3][3];
        for (int i = 0; i < 3; i++)
            for (int kk = 0; kk < 3; kk++) {
                float s = 0;
                for (int j = 0; j < 3; j++) s += H[i][j] * R[3 * j + kk];
                HR[i][kk] = s;
            }
        for (int kk = 0; kk < 3; kk++) {
            float s = 0;
            for (int i = 0; i < 3; i++) s += R[3 * i + kk] * HR[i][kk];
            grad[kk] = 2.f * ax[kk] * s;
        }
        for (int i = 0; i < 3; i++) {
            float s = 0;
            for (int j = 0; j < 3; j++) s += H[i][j] * cen[j];
            grad[3 + i] = -2.f * s - 2.f * H[i][3];
        }
        for (int i = 0; i < 3; i++)
            for (int kk = 0; kk < 3; kk++) grad[6 + 3 * i + kk] = 2.f * HR[i][kk] * ax[kk] * ax[kk];
        // ---------------- Adam (torch semantics, eps = 1e-15) ----------------
        step++;
        const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
        const float bc2s = (float)sqrt(bc2);
        for (int q = 0; q < 15; q++) {
            const double lr = (q < 3) ? A.lr[0] : ((q < 6) ? A.lr[1] : A.lr[2]);
            const float step_size = (float)(lr / bc1);
            const float g = grad[q];
            mom[q] = mom[q] + omb1 * (g - mom[q]);
            var[q] = var[q] * b2f + omb2 * g * g;
            const float denom = sqrtf(var[q]) / bc2s + eps;
            p[q] = p[q] - step_size * (mom[q] / denom);
        }
    }
    for (int k = 0; k < 3; k++) A.axes[3 * o + k] = p[k];
    for (int k = 0; k < 3; k++) A.center[3 * o + k] = p[3 + k];
    for (int k = 0; k < 9; k++) A.R[9 * o + k] = p[6 + k];
    if (A.last_loss) A.last_loss[o] = loss_v;
}

} // namespace dqo

using namespace dqo;

extern "C" int dqo_quadric_init(int32_t n, const double *bboxes, const double *depth_stats, const double *K,
                                const double *Rts, double *axes, double *R, double *center, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0) return DQO_ERR_INVALID_ARG;
    if (n == 0) return DQO_OK;
    if (!bboxes || !depth_stats || !K || !Rts || !axes || !R || !center) {
        set_error("dqo_quadric_init: null pointer argument");
        return DQO_ERR_INVALID_ARG;
    }
    quadric_init_kernel<<<(n + 63) / 64, 64, 0, stream>>>(n, bboxes, depth_stats, K, Rts, axes, R, center);
    DQO_LAUNCH_CHECK("quadric init", 0, stream);
    return DQO_OK;
}

extern "C" int dqo_quadric_project(int32_t n, const double *axes, const double *R, const double *center,
                                   const double *P, double *bbox, double *ellipse, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0) return DQO_ERR_INVALID_ARG;
    if (n == 0) return DQO_OK;
    if (!axes || !R || !center || !P) {
        set_error("dqo_quadric_project: null pointer argument");
        return DQO_ERR_INVALID_ARG;
    }
    quadric_project_kernel<<<(n + 63) / 64, 64, 0, stream>>>(n, axes, R, center, P, bbox, ellipse);
    DQO_LAUNCH_CHECK("quadric project", 0, stream);
    return DQO_OK;
}

extern "C" int dqo_quadric_refine(int32_t n, int32_t iters, int32_t max_views, const int32_t *n_views,
                                  const float *obs_bboxes, const float *Ps, const int32_t *view_choice, float lr_axes,
                                  float lr_center, float lr_R, float *axes, float *R, float *center, float *last_loss,
                                  void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || iters < 0 || max_views <= 0) {
        set_error("dqo_quadric_refine: invalid sizes");
        return DQO_ERR_INVALID_ARG;
    }
    if (n == 0) return DQO_OK;
    if (!n_views || !obs_bboxes || !Ps || !view_choice || !axes || !R || !center) {
        set_error("dqo_quadric_refine: null pointer argument");
        return DQO_ERR_INVALID_ARG;
    }
    RefineArgs A;
    A.n = n; A.iters = iters; A.max_views = max_views; A.n_views = n_views; A.obs = obs_bboxes; A.Ps = Ps;
    A.choice = view_choice; A.lr[0] = lr_axes; A.lr[1] = lr_center; A.lr[2] = lr_R;
    A.axes = axes; A.R = R; A.center = center; A.last_loss = last_loss;
    quadric_refine_kernel<<<(n + 31) / 32, 32, 0, stream>>>(A);
    DQO_LAUNCH_CHECK("quadric refine", 0, stream);
    return DQO_OK;
}
