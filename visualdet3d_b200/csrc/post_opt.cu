// Post-optimisation of the 3-D box orientation by hill climbing:
//   R/lib/fast_utils/hill_climbing.py:24-122 (post_optimization / hill_climb / test_projection),
//   R/lib/fast_utils/bbox3d.py:19-82 (project_3d), R/lib/fast_utils/bbox2d.py:4-66 (bbox2d_area, iou_2d).
// For every car detection deeper than 3 m (detection_3d_head.py:303) the yaw is moved in +-step_r steps (0.4 rad, halved when
// neither direction improves, until step_r <= r_lim) so that the 2-D hull of the projected 3-D box overlaps the detected 2-D box best.
//
// One routine, float64 like the reference's numba code, compiled for the host (vd3d_post_opt_host: what Mono3D detectors call on the
// few kept rows after the result copy; the reference does the same on the CPU with a D2H `.item()` per detection) and for the device
// (post_opt_kernel / vd3d_post_opt: one thread per detection on the fixed-capacity NMS output, no host round trip).  This translation unit
// is compiled with -fmad=false (__graft_entry__.py): the device then evaluates every a * b + c of the search in two IEEE roundings like the
// host compiler and the reference's numba code do, so the discrete accept / halve decisions agree.
#include "common.cuh"
#include <math.h>

namespace vd3d {

struct HillIn {
    double p2[16], p2_inv[16];      // 4x4 projection (P2 in rows 0..2, row 3 = 0 0 0 1) and its inverse, row-major
    double img_w, img_h;            // the reference clips the projected hull to 1280 x 288 (hill_climbing.py:98-103)
    double step_r_init, r_lim;
};

// IoU of the detected box (float32 corners, area formed in float32 like bbox2d_area on a float32 array) with the hull of the projected box
__host__ __device__ inline double hull_iou(const HillIn& q, const float* box, double cx, double cy, double z, double w3d, double h3d, double l3d, double ry) {
    // centre in camera coordinates: p2_inv . (cx z, cy z, z, 1)
    const double v0 = cx * z, v1 = cy * z, v2 = z, v3 = 1.0;
    double c3[3];
    for (int r = 0; r < 3; ++r) c3[r] = ((q.p2_inv[4 * r] * v0 + q.p2_inv[4 * r + 1] * v1) + q.p2_inv[4 * r + 2] * v2) + q.p2_inv[4 * r + 3] * v3;
    const double cs = cos(ry), sn = sin(ry);
    const double xs[8] = {0.0, l3d, l3d, l3d, l3d, 0.0, 0.0, 0.0};
    const double ys[8] = {0.0, 0.0, h3d, h3d, 0.0, 0.0, h3d, h3d};
    const double zs[8] = {0.0, 0.0, 0.0, w3d, w3d, w3d, w3d, 0.0};
    double xmin = 0, ymin = 0, xmax = 0, ymax = 0;
    for (int i = 0; i < 8; ++i) {
        const double xc = xs[i] + (-l3d / 2), yc = ys[i] + (-h3d / 2), zc = zs[i] + (-w3d / 2);
        // R . corner  (R = [[c, 0, s], [0, 1, 0], [-s, 0, c]]), then translate
        const double X = ((cs * xc + 0.0 * yc) + sn * zc) + c3[0];
        const double Y = ((0.0 * xc + 1.0 * yc) + 0.0 * zc) + c3[1];
        const double Z = ((-sn * xc + 0.0 * yc) + cs * zc) + c3[2];
        const double u = ((q.p2[0] * X + q.p2[1] * Y) + q.p2[2] * Z) + q.p2[3];
        const double v = ((q.p2[4] * X + q.p2[5] * Y) + q.p2[6] * Z) + q.p2[7];
        const double d = ((q.p2[8] * X + q.p2[9] * Y) + q.p2[10] * Z) + q.p2[11];
        const double px = u / d, py = v / d;
        if (i == 0) { xmin = xmax = px; ymin = ymax = py; }
        else { xmin = fmin(xmin, px); xmax = fmax(xmax, px); ymin = fmin(ymin, py); ymax = fmax(ymax, py); }
    }
    const double x_new = fmax(0.0, xmin), y_new = fmax(0.0, ymin), x2_new = fmin(xmax, q.img_w), y2_new = fmin(ymax, q.img_h);
    // iou_2d(b1 float32, b2 float64)
    float dx0 = box[2] - box[0], dy0 = box[3] - box[1];
    if (dx0 < 0.f) dx0 = 0.f;
    if (dy0 < 0.f) dy0 = 0.f;
    const float area0 = dx0 * dy0;
    double dx1 = x2_new - x_new, dy1 = y2_new - y_new;
    if (dx1 < 0.0) dx1 = 0.0;
    if (dy1 < 0.0) dy1 = 0.0;
    const double area1 = dx1 * dy1;
    const double ix1 = fmax((double)box[0], x_new), ix2 = fmin((double)box[2], x2_new);
    const double iy1 = fmax((double)box[1], y_new), iy2 = fmin((double)box[3], y2_new);
    const double dx = ix2 - ix1, dy = iy2 - iy1;
    if (dx <= 0.0 || dy <= 0.0) return 0.0;
    const double area = dx * dy;
    return area / (((double)area0 + area1) - area);
}

// hill_climb (hill_climbing.py:52-80) with min_ol_dif = 0; returns the wrapped yaw and the best IoU
__host__ __device__ inline void hill_climb(const HillIn& q, const float* box, double cx, double cy, double z, double w3d, double h3d, double l3d,
                                           double ry, double* ry_out, double* iou_out) {
    double step_r = q.step_r_init;
    double best = hull_iou(q, box, cx, cy, z, w3d, h3d, l3d, ry);
    while (step_r > q.r_lim) {
        const double neg = hull_iou(q, box, cx, cy, z, w3d, h3d, l3d, ry - step_r);
        const double pos = hull_iou(q, box, cx, cy, z, w3d, h3d, l3d, ry + step_r);
        const bool invalid = ((pos - best) <= 0.0) && ((neg - best) <= 0.0);
        if (invalid) step_r = step_r * 0.5;
        else if ((pos - best) > 0.0 && pos > neg) { ry += step_r; best = pos; }
        else if ((neg - best) > 0.0) { ry -= step_r; best = neg; }
        else step_r = step_r * 0.5;
    }
    while (ry > 3.14) ry -= 3.14 * 2;
    while (ry < -3.14) ry += 3.141592653589793 * 2;
    *ry_out = ry; *iou_out = best;
}

// device form: boxes [B][cap][11] = (x1, y1, x2, y2, cx, cy, z, w, h, l, alpha) after NMS, cls [B][cap], count [B]; P2 [B][3][4];
// theta0 / theta_out in the reference's convention (convertAlpha2Rot / convertRot2Alpha with the pixel abscissa): alpha +- atan2(cx - P2[0,2], P2[0,0])
__global__ void post_opt_kernel(float* __restrict__ boxes, const long long* __restrict__ cls, const int* __restrict__ count, const float* __restrict__ P2,
                                int B, int cap, double img_w, double img_h, double step_r_init, double r_lim, float min_depth, int label) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * cap) return;
    const int b = idx / cap, k = idx - b * cap;
    if (k >= count[b]) return;
    float* bx = boxes + (long long)idx * 11;
    if (cls[idx] != label) return;
    const float* P = P2 + b * 12;
    HillIn q;
    for (int i = 0; i < 12; ++i) q.p2[i] = (double)P[i];
    q.p2[12] = q.p2[13] = q.p2[14] = 0.0; q.p2[15] = 1.0;
    // inverse of [[fx 0 cx tx], [0 fy cy ty], [0 0 1 tz], [0 0 0 1]] (KITTI P2 has this shape; general matrices go through the host entry)
    const double fx = q.p2[0], fy = q.p2[5], cxp = q.p2[2], cyp = q.p2[6], tx = q.p2[3], ty = q.p2[7], tz = q.p2[11];
    const double inv[16] = {1.0 / fx, 0.0, -cxp / fx, (cxp * tz - tx) / fx,   0.0, 1.0 / fy, -cyp / fy, (cyp * tz - ty) / fy,   0.0, 0.0, 1.0, -tz,   0.0, 0.0, 0.0, 1.0};
    for (int i = 0; i < 16; ++i) q.p2_inv[i] = inv[i];
    q.img_w = img_w; q.img_h = img_h; q.step_r_init = step_r_init; q.r_lim = r_lim;
    const float cx = bx[4], cy = bx[5], z = bx[6];
    // depth test on the back-projected z (BackProjection keeps z): detection_3d_head.py:303
    if (!(z > min_depth)) return;
    // float32 atan2 of the reference's numpy call (convertAlpha2Rot): evaluated in double and rounded once == a correctly rounded atan2f
    const float off = (float)atan2((double)(cx - P[2]), (double)P[0]);
    float th0 = bx[10] + off;
    if (th0 > 3.14159274f) th0 -= 6.28318548f;
    if (th0 <= -3.14159274f) th0 += 6.28318548f;
    double ry, iou;
    hill_climb(q, bx, (double)cx, (double)cy, (double)z, (double)bx[7], (double)bx[8], (double)bx[9], (double)th0, &ry, &iou);
    double alpha = ry - (double)off;
    if (alpha > 3.141592653589793) alpha -= 2 * 3.141592653589793;
    if (alpha <= -3.141592653589793) alpha += 2 * 3.141592653589793;
    bx[10] = (float)alpha;
}

}  // namespace vd3d

using namespace vd3d;

extern "C" int vd3d_post_opt_host(const double* p2, const double* p2_inv, int n, const float* box2d, const double* cx, const double* cy,
                                  const float* z, const float* w, const float* h, const float* l, const float* theta0,
                                  double img_w, double img_h, double step_r_init, double r_lim, double* theta_out, double* iou_out) {
    VD3D_REQUIRE(p2 && p2_inv && (n == 0 || (box2d && cx && cy && z && w && h && l && theta0 && theta_out)), "post_opt_host: null pointer");
    VD3D_REQUIRE(n >= 0 && step_r_init > 0.0 && r_lim >= 0.0, "post_opt_host: n >= 0, step_r_init > 0 and r_lim >= 0 are required");
    HillIn q;
    for (int i = 0; i < 16; ++i) { q.p2[i] = p2[i]; q.p2_inv[i] = p2_inv[i]; }
    q.img_w = img_w; q.img_h = img_h; q.step_r_init = step_r_init; q.r_lim = r_lim;
    for (int i = 0; i < n; ++i) {
        double ry, iou;
        hill_climb(q, box2d + 4 * i, cx[i], cy[i], (double)z[i], (double)w[i], (double)h[i], (double)l[i], (double)theta0[i], &ry, &iou);
        theta_out[i] = ry;
        if (iou_out) iou_out[i] = iou;
    }
    return VD3D_OK;
}

extern "C" int vd3d_post_opt(float* boxes, const long long* cls, const int* count, const float* P2, int B, int cap,
                             float img_w, float img_h, float step_r_init, float r_lim, float min_depth, int label, void* stream) {
    VD3D_REQUIRE(boxes && cls && count && P2 && B > 0 && cap > 0, "post_opt: bad args");
    VD3D_REQUIRE(step_r_init > 0.f && r_lim > 0.f, "post_opt: step_r_init > 0 and r_lim > 0 are required");
    const int total = B * cap;
    post_opt_kernel<<<cdiv(total, 64), 64, 0, (cudaStream_t)stream>>>(boxes, cls, count, P2, B, cap, (double)img_w, (double)img_h, (double)step_r_init,
                                                                      (double)r_lim, min_depth, label);
    VD3D_CHECK_LAUNCH("post_opt");
    return VD3D_OK;
}
