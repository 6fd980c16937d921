// Post-forward geometry on the device: what `test_one` does between the detector output and the result file
//   BackProjection        R/networks/utils/utils.py:256-278     (u, v, z) -> (x, y, z) in the camera frame
//   alpha2theta_3d        visualDet3D/utils/utils.py:47-62      theta = alpha + atan2(x + tx / fx, z)
//   BBox3dProjector       R/networks/utils/utils.py:198-254     8 corners in the camera frame and in the image (optional outputs)
//   2-D rescale           R/networks/pipelines/evaluators.py:118-127   boxes of the network input -> pixels of the original frame
// One thread per kept row of the fixed-capacity NMS output; float32, the reference's operations in the reference's order with explicit
// round-to-nearest intrinsics (no FMA contraction), so x / y / the rescaled boxes are bit-identical to the reference's tensors; theta and
// the corners go through atan2 / cos / sin (evaluated in double and rounded once: within an ulp of any float32 libm).
// The all-gather record block then carries these columns (vd3d_pack_records_geo), and the host only formats text.
#include "common.cuh"
#include <math.h>

namespace vd3d {

__global__ void post_forward_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ count, const float* __restrict__ P2,
                                    const float* __restrict__ origP, int B, int cap, float* __restrict__ box3d, float* __restrict__ theta_out,
                                    float* __restrict__ box2d, float* __restrict__ corners, float* __restrict__ homo) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * cap) return;
    const int b = idx / cap, k = idx - b * cap;
    const int n = count[b];
    float* o3 = box3d + (long long)idx * 7;
    float* o2 = box2d + (long long)idx * 4;
    if (k >= n) {                                  // rows past the count: defined (zero) so that the record block is deterministic
        for (int i = 0; i < 7; ++i) o3[i] = 0.f;
        for (int i = 0; i < 4; ++i) o2[i] = 0.f;
        theta_out[idx] = 0.f;
        if (corners) for (int i = 0; i < 24; ++i) corners[(long long)idx * 24 + i] = 0.f;
        if (homo) for (int i = 0; i < 24; ++i) homo[(long long)idx * 24 + i] = 0.f;
        return;
    }
    const float* s = boxes + (long long)idx * 11;
    const float* P = P2 + b * 12;
    const float fx = P[0], fy = P[5], cx = P[2], cy = P[6], tx = P[3], ty = P[7];
    // BackProjection: x3d = (u * z - cx * z - tx) / fx
    const float u = s[4], v = s[5], z = s[6];
    const float x = __fdiv_rn(__fsub_rn(__fsub_rn(__fmul_rn(u, z), __fmul_rn(cx, z)), tx), fx);
    const float y = __fdiv_rn(__fsub_rn(__fsub_rn(__fmul_rn(v, z), __fmul_rn(cy, z)), ty), fy);
    o3[0] = x; o3[1] = y; o3[2] = z; o3[3] = s[7]; o3[4] = s[8]; o3[5] = s[9]; o3[6] = s[10];
    // alpha2theta_3d
    const float offset = __fdiv_rn(tx, fx);
    const float th = __fadd_rn(s[10], (float)atan2((double)__fadd_rn(x, offset), (double)z));
    theta_out[idx] = th;
    // 2-D boxes back to the original frame: += shift, *= scale (evaluators.py:118-127); without original_P the boxes are copied
    if (origP) {
        const float* O = origP + b * 12;
        const float sx = __fdiv_rn(O[0], fx), sy = __fdiv_rn(O[5], fy);
        const float shl = __fsub_rn(__fdiv_rn(O[2], sx), cx), sht = __fsub_rn(__fdiv_rn(O[6], sy), cy);
        o2[0] = __fmul_rn(__fadd_rn(s[0], shl), sx);
        o2[1] = __fmul_rn(__fadd_rn(s[1], sht), sy);
        o2[2] = __fmul_rn(__fadd_rn(s[2], shl), sx);
        o2[3] = __fmul_rn(__fadd_rn(s[3], sht), sy);
    } else {
        o2[0] = s[0]; o2[1] = s[1]; o2[2] = s[2]; o2[3] = s[3];
    }
    if (!corners && !homo) return;
    // BBox3dProjector: corner signs (x: w, y: h, z: l)
    const float sgn[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {1, 1, 1}, {1, -1, 1}, {-1, -1, 1}, {-1, 1, 1}, {-1, 1, -1}};
    const float c = (float)cos((double)th), sn = (float)sin((double)th);
    for (int i = 0; i < 8; ++i) {
        const float rx0 = __fmul_rn(0.5f * sgn[i][0], s[7]), ry0 = __fmul_rn(0.5f * sgn[i][1], s[8]), rz0 = __fmul_rn(0.5f * sgn[i][2], s[9]);
        const float X = __fadd_rn(__fadd_rn(__fmul_rn(rz0, c), __fmul_rn(rx0, sn)), x);
        const float Y = __fadd_rn(ry0, y);
        const float Z = __fadd_rn(__fadd_rn(__fmul_rn(-rz0, sn), __fmul_rn(rx0, c)), z);
        if (corners) { float* q = corners + (long long)idx * 24 + i * 3; q[0] = X; q[1] = Y; q[2] = Z; }
        if (homo) {
            float cam[3];
            for (int r = 0; r < 3; ++r)
                cam[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[4 * r], X), __fmul_rn(P[4 * r + 1], Y)), __fmul_rn(P[4 * r + 2], Z)), P[4 * r + 3]);
            const float d = __fadd_rn(cam[2], 1e-6f);
            float* q = homo + (long long)idx * 24 + i * 3;
            q[0] = __fdiv_rn(cam[0], d); q[1] = __fdiv_rn(cam[1], d); q[2] = __fdiv_rn(cam[2], d);
        }
    }
}

// record block with the post-forward columns: rec[b] = [count | kmax x (11 box floats, score, class, x3d, y3d, theta, 4 rescaled 2-D box floats)]
__global__ void pack_records_geo_kernel(const float* __restrict__ scores, const float* __restrict__ boxes, const int64_t* __restrict__ cls,
                                        const int32_t* __restrict__ count, const float* __restrict__ box3d, const float* __restrict__ theta,
                                        const float* __restrict__ box2d, int cap, int kmax, float* __restrict__ rec, const int* __restrict__ range_flag) {
    constexpr int R = 20;
    const int b = blockIdx.x;
    int n = count[b];
    if (range_flag && *range_flag) n = -2;                            // fp16-range guard (see pack_records_kernel)
    float* r = rec + (long long)b * (1 + kmax * R);
    if (threadIdx.x == 0) r[0] = (n > kmax) ? -1.0f : (float)n;
    const int m = n < 0 ? 0 : (n > kmax ? 0 : n);
    for (int i = threadIdx.x; i < kmax * R; i += blockDim.x) {
        const int k = i / R, q = i - k * R;
        float v = 0.f;
        if (k < m) {
            const long long row = (long long)b * cap + k;
            if (q < 11) v = boxes[row * 11 + q];
            else if (q == 11) v = scores[row];
            else if (q == 12) v = (float)cls[row];
            else if (q < 15) v = box3d[row * 7 + (q - 13)];
            else if (q == 15) v = theta[row];
            else v = box2d[row * 4 + (q - 16)];
        }
        r[1 + i] = v;
    }
}

}  // namespace vd3d

using namespace vd3d;

extern "C" int vd3d_post_forward(const float* boxes, const int32_t* count, const float* P2, const float* original_P, int B, int cap,
                                 float* box3d, float* theta, float* box2d, float* corners, float* homo, void* stream) {
    VD3D_REQUIRE(boxes && count && P2 && box3d && theta && box2d && B > 0 && cap > 0, "post_forward: bad args");
    const int total = B * cap;
    post_forward_kernel<<<cdiv(total, 128), 128, 0, (cudaStream_t)stream>>>(boxes, count, P2, original_P, B, cap, box3d, theta, box2d, corners, homo);
    VD3D_CHECK_LAUNCH("post_forward");
    return VD3D_OK;
}

extern "C" int vd3d_pack_records_geo(const float* scores, const float* boxes, const int64_t* cls, const int32_t* count, const float* box3d,
                                     const float* theta, const float* box2d, int B, int cap, int kmax, float* rec, void* stream) {
    VD3D_REQUIRE(scores && boxes && cls && count && box3d && theta && box2d && rec && B > 0 && cap > 0 && kmax > 0, "pack_records_geo: bad args");
    pack_records_geo_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(scores, boxes, cls, count, box3d, theta, box2d, cap, kmax, rec, fp16_range_flag());
    VD3D_CHECK_LAUNCH("pack_records_geo");
    return VD3D_OK;
}
