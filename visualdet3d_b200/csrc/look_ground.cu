// Ground-Aware Convolution sampling stage (LookGround.forward, R/lib/look_ground.py:24-71), NHWC.
//
// For every feature-map pixel (b, h, w):
//   P2' = P2 with rows 0..1 divided by 16                                               (:29-30)
//   disp      = 0.1 * (0.05 t + 0.95 t),  t = tanh(conv3x3(x))  [conv done by the conv engine, channel 0 of `dconv`]   (:32-33)
//   disparity = relu(fy * baseline * (h - cy) / (|fy * elev + Ty| + 1e-10))              (:48-49)   (depends on b, h only)
//   gx = linspace(-1, 1, W)[w] ;  gy = linspace(-1, 1, H)[h] + relu(1.535 (h - cy) / (2 (elev - 0.7675))) / (H/2) + disp   (:52-63)
//   sampled = grid_sample(cat[disparity, x], (gx, gy), bilinear, padding_mode=border, align_corners=True)                 (:66-69)
// Output layout for the 1x1 `extract` conv that follows on the tensor-core engine: [sampled x (C) | sampled disparity | zeros],
// plus its lo companion.  HBM-bound gather: x is read ~once (rows reused through L2), C+32 floats (+lo) written per pixel.
#include "common.cuh"

namespace vd3d {

__device__ __forceinline__ float lin_space(int i, int n) {
    // torch.linspace(-1, 1, n) in float32: start + i*step for the first half, end - (n-1-i)*step for the second half
    float step = (1.0f - (-1.0f)) / (float)(n - 1);
    return (i < n / 2) ? (-1.0f + step * (float)i) : (1.0f - step * (float)(n - 1 - i));
}

__global__ void __launch_bounds__(256) look_ground_sample_kernel(
    const float* __restrict__ x, int B, int H, int W, int C, int x_cs, int x_co,
    const float* __restrict__ dconv, int d_cs, int d_co, const float* __restrict__ P2,
    float baseline, float elev, float ydiv, float* __restrict__ out, float* __restrict__ out_lo, int o_cs) {
    // one warp per pixel; lanes sweep channel quads
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const long long npix = (long long)B * H * W;
    if (warp >= npix) return;
    const int w = warp % W; int r = warp / W; const int h = r % H; const int b = r / H;
    const float* P = P2 + 12 * b;
    const float fy = P[5] / 16.0f, cy = P[6] / 16.0f, Ty = P[7] / 16.0f;
    const float yy = (float)h;
    // learned offset
    float t = tanhf(__ldg(dconv + (long long)warp * d_cs + d_co));
    float disp = 0.1f * (0.05f * t + 0.95f * t);
    // sampling coordinates (normalised), then grid_sample's unnormalise (align_corners=True) + border clip
    const float h_mean = 1.535f;
    float ysb = fmaxf(h_mean * (yy - cy) / ydiv, 0.f) / ((float)H * 0.5f);     // ydiv = float(2 * (elev - 0.5 * 1.535)), host double
    float gy = lin_space(h, H) + (ysb + disp);
    float gx = lin_space(w, W);
    float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
    float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
    ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
    const int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
    const int x1 = x0 + 1, y1 = y0 + 1;
    const float wnw = ((float)x1 - ix) * ((float)y1 - iy), wne = (ix - (float)x0) * ((float)y1 - iy);
    const float wsw = ((float)x1 - ix) * (iy - (float)y0), wse = (ix - (float)x0) * (iy - (float)y0);
    const bool bx0 = x0 >= 0 && x0 < W, bx1 = x1 >= 0 && x1 < W, by0 = y0 >= 0 && y0 < H, by1 = y1 >= 0 && y1 < H;
    const long long base = (long long)b * H * W;
    const float* pnw = x + (base + (long long)y0 * W + x0) * x_cs + x_co;
    const float* pne = pnw + x_cs;
    const float* psw = pnw + (long long)W * x_cs;
    const float* pse = psw + x_cs;
    float* op = out + (long long)warp * o_cs;
    float* ol = out_lo ? out_lo + (long long)warp * o_cs : nullptr;
    for (int c = 4 * lane; c < C; c += 128) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (by0 && bx0) { float4 v = ldg4(pnw + c); a.x += v.x * wnw; a.y += v.y * wnw; a.z += v.z * wnw; a.w += v.w * wnw; }
        if (by0 && bx1) { float4 v = ldg4(pne + c); a.x += v.x * wne; a.y += v.y * wne; a.z += v.z * wne; a.w += v.w * wne; }
        if (by1 && bx0) { float4 v = ldg4(psw + c); a.x += v.x * wsw; a.y += v.y * wsw; a.z += v.z * wsw; a.w += v.w * wsw; }
        if (by1 && bx1) { float4 v = ldg4(pse + c); a.x += v.x * wse; a.y += v.y * wse; a.z += v.z * wse; a.w += v.w * wse; }
        *reinterpret_cast<float4*>(op + c) = a;
        if (ol) {
            float4 l;
            l.x = a.x - __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u);
            l.y = a.y - __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u);
            l.z = a.z - __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u);
            l.w = a.w - __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u);
            *reinterpret_cast<float4*>(ol + c) = l;
        }
    }
    // disparity plane channel: a function of the row only -> bilinear over the two rows (x weights of in-bounds corners)
    if (lane == 0) {
        const float den = fabsf(fy * elev + Ty) + 1e-10f;
        float d0 = fmaxf(fy * baseline * ((float)y0 - cy) / den, 0.f);
        float d1 = fmaxf(fy * baseline * ((float)y1 - cy) / den, 0.f);
        float v = 0.f;
        if (by0 && bx0) v += d0 * wnw;
        if (by0 && bx1) v += d0 * wne;
        if (by1 && bx0) v += d1 * wsw;
        if (by1 && bx1) v += d1 * wse;
        op[C] = v;
        if (ol) ol[C] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
    // zero padding channels (C+1 .. o_cs) are written once by the host (buffer is zero-initialised and never touched)
}

}  // namespace vd3d

using namespace vd3d;

extern "C" int vd3d_look_ground_sample(const float* x, int B, int H, int W, int C, int x_cs, int x_co,
                                       const float* dconv, int d_cs, int d_co, const float* P2,
                                       float baseline, float relative_elevation,
                                       float* out, float* out_lo, int out_cs, void* stream) {
    VD3D_REQUIRE(x && dconv && P2 && out, "look_ground_sample: null pointer");
    VD3D_REQUIRE(B > 0 && H > 1 && W > 1 && C > 0 && C % 4 == 0 && x_cs % 4 == 0 && x_co % 4 == 0, "look_ground_sample: bad shape");
    VD3D_REQUIRE(out_cs >= C + 1 && out_cs % 4 == 0, "look_ground_sample: out pitch must hold C+1 channels");
    const double relative_elevation_d = (double)relative_elevation == (double)1.65f ? 1.65 : (double)relative_elevation;
    long long npix = (long long)B * H * W;
    long long threads = npix * 32;
    look_ground_sample_kernel<<<cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, x_cs, x_co, dconv, d_cs, d_co, P2,
                                                                                 baseline, relative_elevation,
                                                                                 (float)(2.0 * ((double)relative_elevation_d - 0.5 * 1.535)), out, out_lo, out_cs);
    VD3D_CHECK_LAUNCH("look_ground_sample");
    return VD3D_OK;
}
