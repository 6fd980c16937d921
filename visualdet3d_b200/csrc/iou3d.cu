// Rotated bird's-eye-view box overlap / IoU and NMS (R/lib/ops/iou3d): boxes are [x1, y1, x2, y2, ry].
//
// Algorithm (follows iou3d_kernel.cu:34-221 step for step so results agree to rounding): rotate the 4 corners of both
// rectangles about their centres, collect (a) the proper intersections of the 4 x 4 edge pairs and (b) every corner of one
// rectangle lying inside the other (1e-5 margin), order the points by angle about their mean, and take the shoelace area of
// the fan from the first point.  NMS is fully on the device: a 64 x 64-tile suppression bitmask kernel followed by a
// single-warp greedy sweep (the reference copies the mask to the host and sweeps there, iou3d.cpp:87-116).
#include "common.cuh"

namespace vd3d {

constexpr float IOU_EPS = 1e-8f;

struct P2 { float x, y; };

__device__ __forceinline__ float cross3(const P2& p1, const P2& p2, const P2& p0) {
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

// proper intersection of segments (p0,p1) and (q0,q1); returns false when they do not strictly cross
__device__ __forceinline__ bool seg_intersect(const P2& p1, const P2& p0, const P2& q1, const P2& q0, P2& out) {
    // bounding-box rejection (inclusive)
    bool bb = fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
              fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y);
    if (!bb) return false;
    float s1 = cross3(q0, p1, p0);
    float s2 = cross3(p1, q1, p0);
    float s3 = cross3(p0, q1, q0);
    float s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0.f && s3 * s4 > 0.f)) return false;
    float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > IOU_EPS) {
        out.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        out.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        out.x = (b0 * c1 - b1 * c0) / D;
        out.y = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}

__device__ __forceinline__ bool inside_box(const float* box, const P2& p) {
    const float MARGIN = 1e-5f;
    float cx = (box[0] + box[2]) / 2, cy = (box[1] + box[3]) / 2;
    float ac = cosf(-box[4]), as = sinf(-box[4]);
    float rx = (p.x - cx) * ac + (p.y - cy) * as + cx;
    float ry = -(p.x - cx) * as + (p.y - cy) * ac + cy;
    return rx > box[0] - MARGIN && rx < box[2] + MARGIN && ry > box[1] - MARGIN && ry < box[3] + MARGIN;
}

__device__ __forceinline__ void rotated_corners(const float* box, P2 (&c)[5]) {
    float x1 = box[0], y1 = box[1], x2 = box[2], y2 = box[3];
    float cx = (x1 + x2) / 2, cy = (y1 + y2) / 2;
    float ac = cosf(box[4]), as = sinf(box[4]);
    const float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[k].x = (px[k] - cx) * ac + (py[k] - cy) * as + cx;
        c[k].y = -(px[k] - cx) * as + (py[k] - cy) * ac + cy;
    }
    c[4] = c[0];
}

__device__ float rotated_overlap(const float* a, const float* b) {
    P2 ca[5], cb[5];
    rotated_corners(a, ca);
    rotated_corners(b, cb);
    P2 pts[16];
    float sx = 0.f, sy = 0.f;
    int n = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            P2 x;
            if (seg_intersect(ca[i + 1], ca[i], cb[j + 1], cb[j], x)) { pts[n++] = x; sx += x.x; sy += x.y; }
        }
    for (int k = 0; k < 4; ++k) {
        if (inside_box(a, cb[k])) { pts[n++] = cb[k]; sx += cb[k].x; sy += cb[k].y; }
        if (inside_box(b, ca[k])) { pts[n++] = ca[k]; sx += ca[k].x; sy += ca[k].y; }
    }
    float mx = sx / n, my = sy / n;          // n == 0 -> NaN centre, never used (no points)
    // order by angle about the mean: stable exchange sort (same ordering as the reference's bubble sort with a strict '>')
    float ang[16];
    for (int i = 0; i < n; ++i) ang[i] = atan2f(pts[i].y - my, pts[i].x - mx);
    for (int j = 0; j < n - 1; ++j)
        for (int i = 0; i < n - j - 1; ++i)
            if (ang[i] > ang[i + 1]) {
                float t = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = t;
                P2 tp = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = tp;
            }
    float area = 0.f;
    for (int k = 0; k < n - 1; ++k) {
        float ax = pts[k].x - pts[0].x, ay = pts[k].y - pts[0].y;
        float bx = pts[k + 1].x - pts[0].x, by = pts[k + 1].y - pts[0].y;
        area += ax * by - ay * bx;
    }
    return fabsf(area) / 2.0f;
}

__device__ __forceinline__ float rotated_iou(const float* a, const float* b) {
    float sa = (a[2] - a[0]) * (a[3] - a[1]), sb = (b[2] - b[0]) * (b[3] - b[1]);
    float so = rotated_overlap(a, b);
    return so / fmaxf(sa + sb - so, IOU_EPS);
}

__device__ __forceinline__ float aligned_iou(const float* a, const float* b) {
    float l = fmaxf(a[0], b[0]), r = fminf(a[2], b[2]), t = fmaxf(a[1], b[1]), bt = fminf(a[3], b[3]);
    float w = fmaxf(r - l, 0.f), h = fmaxf(bt - t, 0.f);
    float inter = w * h;
    float sa = (a[2] - a[0]) * (a[3] - a[1]), sb = (b[2] - b[0]) * (b[3] - b[1]);
    return inter / fmaxf(sa + sb - inter, IOU_EPS);
}

template <bool IOU>
__global__ void pairwise_kernel(const float* __restrict__ A, int M, const float* __restrict__ Bx, int N, float* __restrict__ out) {
    int a = blockIdx.y * 16 + threadIdx.y, b = blockIdx.x * 16 + threadIdx.x;
    if (a >= M || b >= N) return;
    float ba[5], bb[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { ba[i] = A[a * 5 + i]; bb[i] = Bx[b * 5 + i]; }
    out[(long long)a * N + b] = IOU ? rotated_iou(ba, bb) : rotated_overlap(ba, bb);
}

template <bool ROTATED>
__global__ void nms_mask_kernel(const float* __restrict__ boxes, int N, float thr, unsigned long long* __restrict__ mask) {
    const int row = blockIdx.y, col = blockIdx.x;
    const int rows = min(N - row * 64, 64), cols = min(N - col * 64, 64);
    __shared__ float sb[64 * 5];
    if ((int)threadIdx.x < cols)
        for (int i = 0; i < 5; ++i) sb[threadIdx.x * 5 + i] = boxes[(col * 64 + threadIdx.x) * 5 + i];
    __syncthreads();
    if ((int)threadIdx.x < rows) {
        const int cur = row * 64 + threadIdx.x;
        float cb[5];
        for (int i = 0; i < 5; ++i) cb[i] = boxes[cur * 5 + i];
        unsigned long long t = 0;
        int start = (row == col) ? threadIdx.x + 1 : 0;
        for (int i = start; i < cols; ++i) {
            float v = ROTATED ? rotated_iou(cb, sb + i * 5) : aligned_iou(cb, sb + i * 5);
            if (v > thr) t |= 1ULL << i;
        }
        mask[(long long)cur * gridDim.x + col] = t;
    }
}

// greedy sweep over the suppression mask, one warp; keep[] gets the kept indices in order, *count their number
__global__ void nms_sweep_kernel(const unsigned long long* __restrict__ mask, int N, int col_blocks, long long* __restrict__ keep, int* __restrict__ count) {
    extern __shared__ unsigned long long remv[];
    for (int j = threadIdx.x; j < col_blocks; j += 32) remv[j] = 0;
    __syncwarp();
    int n = 0;
    for (int i = 0; i < N; ++i) {
        int nb = i / 64, ib = i % 64;
        if (!(remv[nb] & (1ULL << ib))) {          // uniform across the warp (shared value)
            if (threadIdx.x == 0) keep[n] = i;
            ++n;
            for (int j = nb + threadIdx.x; j < col_blocks; j += 32) remv[j] |= mask[(long long)i * col_blocks + j];
            __syncwarp();
        }
    }
    if (threadIdx.x == 0) *count = n;
}

}  // namespace vd3d

using namespace vd3d;

extern "C" int vd3d_boxes_overlap_bev(const float* a, int M, const float* b, int N, float* out, void* stream) {
    VD3D_REQUIRE(a && b && out && M >= 0 && N >= 0, "boxes_overlap_bev: bad args");
    if (M == 0 || N == 0) return VD3D_OK;
    dim3 grid(cdiv(N, 16), cdiv(M, 16)), block(16, 16);
    pairwise_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(a, M, b, N, out);
    VD3D_CHECK_LAUNCH("boxes_overlap_bev");
    return VD3D_OK;
}

extern "C" int vd3d_boxes_iou_bev(const float* a, int M, const float* b, int N, float* out, void* stream) {
    VD3D_REQUIRE(a && b && out && M >= 0 && N >= 0, "boxes_iou_bev: bad args");
    if (M == 0 || N == 0) return VD3D_OK;
    dim3 grid(cdiv(N, 16), cdiv(M, 16)), block(16, 16);
    pairwise_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(a, M, b, N, out);
    VD3D_CHECK_LAUNCH("boxes_iou_bev");
    return VD3D_OK;
}

extern "C" long long vd3d_nms_bev_workspace(int N) { return (long long)N * cdiv(N > 0 ? N : 1, 64) * 8 + 64; }

extern "C" int vd3d_nms_bev(const float* boxes, int N, float thresh, int rotated, void* ws, long long* keep, int* count, void* stream) {
    VD3D_REQUIRE(boxes && ws && keep && count && N >= 0, "nms_bev: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) { VD3D_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st)); return VD3D_OK; }
    int cb = cdiv(N, 64);
    VD3D_REQUIRE(cb * 8 <= 48 * 1024, "nms_bev: too many boxes (%d)", N);
    unsigned long long* mask = (unsigned long long*)ws;
    dim3 grid(cb, cb);
    if (rotated) nms_mask_kernel<true><<<grid, 64, 0, st>>>(boxes, N, thresh, mask);
    else nms_mask_kernel<false><<<grid, 64, 0, st>>>(boxes, N, thresh, mask);
    VD3D_CHECK_LAUNCH("nms_mask");
    nms_sweep_kernel<<<1, 32, cb * sizeof(unsigned long long), st>>>(mask, N, cb, keep, count);
    VD3D_CHECK_LAUNCH("nms_sweep");
    return VD3D_OK;
}
