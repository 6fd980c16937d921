// Backward of the deformable convolutions (training side): the reference's two scatter / gather kernels per op
//   modulated_deformable_col2im_gpu_kernel        R/lib/ops/dcn/src/cuda/deform_conv_cuda_kernel.cu:635-686   grad_input  (atomic scatter)
//   modulated_deformable_col2im_coord_gpu_kernel  :688-767                                                     grad_offset, grad_mask
//   deformable_col2im_gpu_kernel / deformable_col2im_coord_gpu_kernel  :279-436   the DCNv1 forms (mask == 1, no grad_mask)
// as ONE pass over the column gradients on NHWC data.  colgrad[pix][k*C + c] = sum_o W[o, c, k] * grad_out[pix][o] is a plain GEMM (the
// reference calls cuBLAS for it, deform_conv_cuda.cpp:625-628) done by the caller.  Per (pixel, tap) the sampling position, the four
// bilinear weights and the mask are computed once (same rule as the forward gather, dcn.cu); threads then sweep channel quads:
//   grad_x[corner][c]  += w_corner * mask * colgrad          (red.global.add.v4.f32: one 16-byte reduction per corner and quad)
//   grad_mask[pix][k]   = sum_c colgrad * bilinear(x[c])     (block-level reduction in shared memory)
//   grad_off[pix][2k+d] = sum_c colgrad * mask * d(bilinear)/d(h | w)
// The reference writes `columns` (9C floats per pixel) to HBM and re-reads it twice per image in a serial loop over the batch; here the
// whole batch is one launch.  Sums over channels run in a different order than the reference's sequential loops: equality is to fp32
// rounding (tests compare with rtol 1e-4 against the reference's own compiled extension).
#include "common.cuh"

namespace vd3d {

struct DcnBwdParams {
    const float* x; int B, H, W, C, x_cs, x_co;
    const float* off; int off_cs, off_co;
    const float* msk; int msk_cs, msk_co;           // nullptr -> DCNv1
    int KH, KW, stride, pad, dil, dg;
    int Ho, Wo;
    const float* colgrad; int cg_cs;                // [B*Ho*Wo][cg_cs], channels k*C + c
    float* grad_x; int gx_cs, gx_co;                // NHWC, accumulated into (caller zero-fills)
    float* grad_off; int go_cs, go_co;              // NHWC [pix][.. g*2K + 2k (+1)]
    float* grad_msk; int gm_cs, gm_co;              // NHWC [pix][.. g*K + k]; nullptr for DCNv1
};

constexpr int DCNB_PIX = 16;
constexpr int DCNB_THREADS = 256;
constexpr int DCNB_MAXK = 49;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(DCNB_THREADS) deform_col2im_kernel(const DcnBwdParams p) {
    __shared__ int s_base[DCNB_PIX][DCNB_MAXK][2];   // h_low * W + w_low, validity bits of the 4 neighbours (bit 4 = tap inside)
    __shared__ float s_w[DCNB_PIX][DCNB_MAXK][4];    // bilinear weights (hh*hw, hh*lw, lh*hw, lh*lw)
    __shared__ float s_f[DCNB_PIX][DCNB_MAXK][4];    // lh, lw (fractions), mask, unused
    __shared__ float s_acc[DCNB_PIX][DCNB_MAXK][3];  // sums over channels: d/dh, d/dw, mask gradient
    const int K = p.KH * p.KW;
    const long long npix = (long long)p.B * p.Ho * p.Wo;
    const long long pix0 = (long long)blockIdx.x * DCNB_PIX;
    const int cpg = p.C / p.dg;
    const int g = blockIdx.y;
    for (int i = threadIdx.x; i < DCNB_PIX * K; i += DCNB_THREADS) {
        const int pl = i / K, k = i - pl * K;
        const long long pix = pix0 + pl;
        int flags = 0, base = 0;
        float w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f, m = 1.f, lh = 0.f, lw = 0.f;
        if (pix < npix) {
            const int wo = (int)(pix % p.Wo); const long long r = pix / p.Wo; const int ho = (int)(r % p.Ho);
            const int kh = k / p.KW, kw = k - kh * p.KW;
            const float* op = p.off + pix * p.off_cs + p.off_co + g * 2 * K + 2 * k;
            const float dh = __ldg(op), dw = __ldg(op + 1);
            if (p.msk) m = __ldg(p.msk + pix * p.msk_cs + p.msk_co + g * K + k);
            const float h = (float)(ho * p.stride - p.pad + kh * p.dil) + dh;
            const float w = (float)(wo * p.stride - p.pad + kw * p.dil) + dw;
            if (h > -1.f && w > -1.f && h < (float)p.H && w < (float)p.W) {
                const int hl = (int)floorf(h), wl = (int)floorf(w);
                lh = h - (float)hl; lw = w - (float)wl;
                const float hh = 1.f - lh, hw = 1.f - lw;
                flags = 16;
                if (hl >= 0 && wl >= 0) flags |= 1;
                if (hl >= 0 && wl + 1 <= p.W - 1) flags |= 2;
                if (hl + 1 <= p.H - 1 && wl >= 0) flags |= 4;
                if (hl + 1 <= p.H - 1 && wl + 1 <= p.W - 1) flags |= 8;
                base = hl * p.W + wl;
                w1 = hh * hw; w2 = hh * lw; w3 = lh * hw; w4 = lh * lw;
            }
        }
        s_base[pl][k][0] = base; s_base[pl][k][1] = flags;
        s_w[pl][k][0] = w1; s_w[pl][k][1] = w2; s_w[pl][k][2] = w3; s_w[pl][k][3] = w4;
        s_f[pl][k][0] = lh; s_f[pl][k][1] = lw; s_f[pl][k][2] = m;
        s_acc[pl][k][0] = 0.f; s_acc[pl][k][1] = 0.f; s_acc[pl][k][2] = 0.f;
    }
    __syncthreads();
    const int cq = cpg / 4;
    const int items = DCNB_PIX * K * cq;
    for (int i0 = 0; i0 < items; i0 += DCNB_THREADS) {
        const int i = i0 + threadIdx.x;
        float a_h = 0.f, a_w = 0.f, a_m = 0.f;
        int pl = 0, k = 0;
        bool live = false;
        if (i < items) {
            const int q = i % cq; const int r = i / cq; k = r % K; pl = r / K;
            const long long pix = pix0 + pl;
            const int flags = s_base[pl][k][1];
            if (pix < npix && (flags & 16)) {
                live = true;
                const int b = (int)(pix / ((long long)p.Ho * p.Wo));
                const int c = g * cpg + 4 * q;
                const int base = s_base[pl][k][0];
                const float4 gq = ldg4(p.colgrad + pix * p.cg_cs + (long long)k * p.C + c);
                const long long xb = ((long long)b * p.H * p.W) * p.x_cs + p.x_co + c;
                const long long gb = ((long long)b * p.H * p.W) * p.gx_cs + p.gx_co + c;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 v1 = (flags & 1) ? ldg4(p.x + xb + (long long)base * p.x_cs) : z;
                const float4 v2 = (flags & 2) ? ldg4(p.x + xb + (long long)(base + 1) * p.x_cs) : z;
                const float4 v3 = (flags & 4) ? ldg4(p.x + xb + (long long)(base + p.W) * p.x_cs) : z;
                const float4 v4 = (flags & 8) ? ldg4(p.x + xb + (long long)(base + p.W + 1) * p.x_cs) : z;
                const float w1 = s_w[pl][k][0], w2 = s_w[pl][k][1], w3 = s_w[pl][k][2], w4 = s_w[pl][k][3];
                const float lh = s_f[pl][k][0], lw = s_f[pl][k][1], m = s_f[pl][k][2];
                const float hh = 1.f - lh, hw = 1.f - lw;
                // grad_input: cur_top_grad = colgrad * mask, scattered with the bilinear weights (col2im, :669-684)
                if (p.grad_x) {
                    const float t0 = gq.x * m, t1 = gq.y * m, t2 = gq.z * m, t3 = gq.w * m;
                    if (flags & 1) red_add_v4(p.grad_x + gb + (long long)base * p.gx_cs, w1 * t0, w1 * t1, w1 * t2, w1 * t3);
                    if (flags & 2) red_add_v4(p.grad_x + gb + (long long)(base + 1) * p.gx_cs, w2 * t0, w2 * t1, w2 * t2, w2 * t3);
                    if (flags & 4) red_add_v4(p.grad_x + gb + (long long)(base + p.W) * p.gx_cs, w3 * t0, w3 * t1, w3 * t2, w3 * t3);
                    if (flags & 8) red_add_v4(p.grad_x + gb + (long long)(base + p.W + 1) * p.gx_cs, w4 * t0, w4 * t1, w4 * t2, w4 * t3);
                }
                // coordinate weights (dmcn_get_coordinate_weight, :523-568): d/dh = -hw v1 - lw v2 + hw v3 + lw v4 ; d/dw = -hh v1 + hh v2 - lh v3 + lh v4
                const float dhx = -hw * v1.x - lw * v2.x + hw * v3.x + lw * v4.x, dhy = -hw * v1.y - lw * v2.y + hw * v3.y + lw * v4.y;
                const float dhz = -hw * v1.z - lw * v2.z + hw * v3.z + lw * v4.z, dhw_ = -hw * v1.w - lw * v2.w + hw * v3.w + lw * v4.w;
                const float dwx = -hh * v1.x + hh * v2.x - lh * v3.x + lh * v4.x, dwy = -hh * v1.y + hh * v2.y - lh * v3.y + lh * v4.y;
                const float dwz = -hh * v1.z + hh * v2.z - lh * v3.z + lh * v4.z, dww = -hh * v1.w + hh * v2.w - lh * v3.w + lh * v4.w;
                a_h = (gq.x * dhx + gq.y * dhy + gq.z * dhz + gq.w * dhw_) * m;
                a_w = (gq.x * dwx + gq.y * dwy + gq.z * dwz + gq.w * dww) * m;
                // mask gradient: colgrad * bilinear(x) (:748)
                a_m = gq.x * (w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x) + gq.y * (w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y) +
                      gq.z * (w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z) + gq.w * (w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w);
            }
        }
        if (live) {
            atomicAdd(&s_acc[pl][k][0], a_h);
            atomicAdd(&s_acc[pl][k][1], a_w);
            atomicAdd(&s_acc[pl][k][2], a_m);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DCNB_PIX * K; i += DCNB_THREADS) {
        const int pl = i / K, k = i - pl * K;
        const long long pix = pix0 + pl;
        if (pix >= npix) continue;
        if (p.grad_off) {
            float* o = p.grad_off + pix * p.go_cs + p.go_co + g * 2 * K + 2 * k;
            o[0] = s_acc[pl][k][0]; o[1] = s_acc[pl][k][1];
        }
        if (p.grad_msk) p.grad_msk[pix * p.gm_cs + p.gm_co + g * K + k] = s_acc[pl][k][2];
    }
}

}  // namespace vd3d

using namespace vd3d;

extern "C" int vd3d_deform_col2im_nhwc(const float* x, int B, int H, int W, int C, int x_cs, int x_co,
                                       const float* off, int off_cs, int off_co, const float* msk, int msk_cs, int msk_co,
                                       int KH, int KW, int stride, int pad, int dil, int deform_groups,
                                       const float* colgrad, int cg_cs,
                                       float* grad_x, int gx_cs, int gx_co, float* grad_off, int go_cs, int go_co,
                                       float* grad_msk, int gm_cs, int gm_co, void* stream) {
    VD3D_REQUIRE(x && off && colgrad && (grad_x || grad_off || grad_msk), "deform_col2im: null pointer");
    VD3D_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && KH * KW <= DCNB_MAXK, "deform_col2im: bad shape (<= 49 taps)");
    VD3D_REQUIRE(deform_groups >= 1 && C % deform_groups == 0 && (C / deform_groups) % 4 == 0, "deform_col2im: channels per deformable group must be a multiple of 4");
    VD3D_REQUIRE(x_cs % 4 == 0 && x_co % 4 == 0 && cg_cs % 4 == 0 && cg_cs >= KH * KW * C && (!grad_x || (gx_cs % 4 == 0 && gx_co % 4 == 0)),
                 "deform_col2im: pitches/offsets must be multiples of 4");
    VD3D_REQUIRE(!grad_msk || msk, "deform_col2im: grad_mask needs the mask");
    DcnBwdParams p;
    p.x = x; p.B = B; p.H = H; p.W = W; p.C = C; p.x_cs = x_cs; p.x_co = x_co;
    p.off = off; p.off_cs = off_cs; p.off_co = off_co; p.msk = msk; p.msk_cs = msk_cs; p.msk_co = msk_co;
    p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad; p.dil = dil; p.dg = deform_groups;
    p.Ho = (H + 2 * pad - (dil * (KH - 1) + 1)) / stride + 1;
    p.Wo = (W + 2 * pad - (dil * (KW - 1) + 1)) / stride + 1;
    VD3D_REQUIRE(p.Ho > 0 && p.Wo > 0, "deform_col2im: empty output");
    p.colgrad = colgrad; p.cg_cs = cg_cs;
    p.grad_x = grad_x; p.gx_cs = gx_cs; p.gx_co = gx_co; p.grad_off = grad_off; p.go_cs = go_cs; p.go_co = go_co;
    p.grad_msk = grad_msk; p.gm_cs = gm_cs; p.gm_co = gm_co;
    const long long npix = (long long)B * p.Ho * p.Wo;
    dim3 grid(cdiv(npix, DCNB_PIX), deform_groups);
    deform_col2im_kernel<<<grid, DCNB_THREADS, 0, (cudaStream_t)stream>>>(p);
    VD3D_CHECK_LAUNCH("deform_col2im");
    return VD3D_OK;
}
