// CenterNet-style decode for the MonoFlex head (MonoFlexHead.get_bboxes, R/heads/monoflex_head.py:114-179), batched and
// fully on the device: sigmoid + 3x3 local-maximum test (`_nms`, R/utils/rtm3d_utils.py:122-127), top-K over (class, y, x)
// (`_topk`, :201-216), gather of the 8 regression maps at the K peaks (`_gather_output`, monoflex_head.py:45-75), depth from
// exp(-d) and from three keypoint-height groups (`decode_depth_from_keypoints`, rtm3d_utils.py:141-182) merged by inverse
// uncertainty (`merge_depth`, :86-91), alpha from the two-bin rotation (`_decode_alpha`, :106-112), x4 up-scaling, ClipBoxes
// and class-agnostic torchvision-style NMS.
//
// Order equivalence: the reference keeps the K best peaks, then drops scores <= score_thr; selecting peaks > score_thr first
// and keeping the best K of those yields the same set and order (scores are sorted descending, ties by flat index).
#include "common.cuh"

namespace vd3d {

struct CnLayout {            // channel offsets inside the concatenated head-output tensor [B][H][W][cs]
    int cs, hm, bbox2d, hps, rot, dim, reg, depth, dunc, cunc;
};

__device__ __forceinline__ float sigm(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// ---- stage 1: peaks above the threshold -> unordered (key, flat index) list ------------------------------------------------
__global__ void cn_peaks_kernel(const float* __restrict__ out, int B, int H, int W, int ncls, CnLayout L, float score_thr, int cap,
                                unsigned long long* __restrict__ keys, int* __restrict__ ncand) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * ncls * H * W;
    if (idx >= total) return;
    int w = (int)(idx % W); long long r = idx / W; int h = (int)(r % H); r /= H; int c = (int)(r % ncls); int b = (int)(r / ncls);
    const float* base = out + ((long long)b * H * W) * L.cs + L.hm + c;
    float s = sigm(__ldg(base + ((long long)h * W + w) * L.cs));
    if (!(s > score_thr)) return;
    float m = s;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            int hh = h + dy, ww = w + dx;
            if ((dy | dx) == 0 || hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
            m = fmaxf(m, sigm(__ldg(base + ((long long)hh * W + ww) * L.cs)));
        }
    if (m != s) return;                                  // keep = (hmax == heat)
    int slot = atomicAdd(ncand + b, 1);
    if (slot >= cap) return;
    unsigned int flat = (unsigned int)((c * H + h) * W + w);
    keys[(long long)b * cap + slot] = ((unsigned long long)(~__float_as_uint(s)) << 32) | flat;
}

// ---- stage 2: one CTA per image: sort, keep K, decode, NMS ----------------------------------------------------------------
constexpr int CN_THREADS = 1024;
constexpr int CN_MAXK = 128;

__global__ void __launch_bounds__(CN_THREADS) cn_decode_nms_kernel(
    const float* __restrict__ out, const float* __restrict__ P2, int H, int W, int ncls, CnLayout L, int cap, int cap_pow2, int K,
    float unc_lo, float unc_hi, double iou_thr, float img_w, float img_h, int out_cap,
    const unsigned long long* __restrict__ keys, const int* __restrict__ ncand,
    float* __restrict__ o_scores, float* __restrict__ o_boxes, long long* __restrict__ o_cls, int* __restrict__ o_index,
    int* __restrict__ o_count, int* __restrict__ o_ncand) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(sm_raw);            // [cap_pow2]
    __shared__ float sbox[CN_MAXK][11];
    __shared__ float sarea[CN_MAXK];
    __shared__ unsigned char ssup[CN_MAXK];
    __shared__ int s_nkeep;
    const int b = blockIdx.x, t = threadIdx.x;
    int n = ncand[b];
    if (t == 0) o_ncand[b] = n;
    if (n > cap) { if (t == 0) o_count[b] = -1; return; }
    for (int i = t; i < cap_pow2; i += CN_THREADS) skey[i] = (i < n) ? keys[(long long)b * cap + i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= cap_pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < cap_pow2; i += CN_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    bool up = ((i & k) == 0);
                    unsigned long long a = skey[i], c = skey[ixj];
                    if ((a > c) == up) { skey[i] = c; skey[ixj] = a; }
                }
            }
            __syncthreads();
        }
    const int nk = min(n, K);
    if (t < nk) {
        unsigned int flat = (unsigned int)(skey[t] & 0xffffffffu);
        int x = flat % W; int r = flat / W; int y = r % H;
        const float* px = out + (((long long)b * H + y) * W + x) * L.cs;
        float xs = (float)x, ys = (float)y;
        float bx1 = xs - px[L.bbox2d + 0], by1 = ys - px[L.bbox2d + 1], bx2 = xs + px[L.bbox2d + 2], by2 = ys + px[L.bbox2d + 3];
        // depths
        float d0 = expf(-px[L.depth]);
        const float* kp = px + L.hps;                        // [10][2]
        float ph = px[L.dim + 1];
        float f = P2[12 * b + 0];
        float fh = f * ph;
        const float EPS = 1e-8f;
        float ch = kp[2 * 8 + 1] - kp[2 * 9 + 1];
        float h02a = kp[2 * 7 + 1] - kp[2 * 0 + 1], h02b = kp[2 * 3 + 1] - kp[2 * 4 + 1];
        float h13a = kp[2 * 2 + 1] - kp[2 * 1 + 1], h13b = kp[2 * 6 + 1] - kp[2 * 5 + 1];
        float dc = fh / (fmaxf(ch, 0.f) * 4.f + EPS);
        float d02 = (fh / (fmaxf(h02a, 0.f) * 4.f + EPS) + fh / (fmaxf(h02b, 0.f) * 4.f + EPS)) / 2.f;
        float d13 = (fh / (fmaxf(h13a, 0.f) * 4.f + EPS) + fh / (fmaxf(h13b, 0.f) * 4.f + EPS)) / 2.f;
        dc = fminf(fmaxf(dc, 0.1f), 100.f); d02 = fminf(fmaxf(d02, 0.1f), 100.f); d13 = fminf(fmaxf(d13, 0.1f), 100.f);
        float u0 = expf(fminf(fmaxf(px[L.dunc], unc_lo), unc_hi));
        float u1 = expf(fminf(fmaxf(px[L.cunc + 0], unc_lo), unc_hi));
        float u2 = expf(fminf(fmaxf(px[L.cunc + 1], unc_lo), unc_hi));
        float u3 = expf(fminf(fmaxf(px[L.cunc + 2], unc_lo), unc_hi));
        float w0 = 1.f / u0, w1 = 1.f / u1, w2 = 1.f / u2, w3 = 1.f / u3;
        float ws = ((w0 + w1) + w2) + w3;
        w0 /= ws; w1 /= ws; w2 /= ws; w3 /= ws;
        float z = ((d0 * w0 + dc * w1) + d02 * w2) + d13 * w3;
        // alpha
        const float* rot = px + L.rot;
        float a1 = atanf(rot[2] / rot[3]) + (-0.5f * 3.14159265358979323846f);
        float a2 = atanf(rot[6] / rot[7]) + (0.5f * 3.14159265358979323846f);
        float sel = (rot[1] > rot[5]) ? 1.f : 0.f;
        float alpha = a1 * sel + a2 * (1.f - sel);
        float cx = (xs + px[L.reg + 0]) * 4.f, cy = (ys + px[L.reg + 1]) * 4.f;
        bx1 *= 4.f; by1 *= 4.f; bx2 *= 4.f; by2 *= 4.f;
        bx1 = fmaxf(bx1, 0.f); by1 = fmaxf(by1, 0.f); bx2 = fminf(bx2, img_w); by2 = fminf(by2, img_h);
        float* sb = sbox[t];
        sb[0] = bx1; sb[1] = by1; sb[2] = bx2; sb[3] = by2; sb[4] = cx; sb[5] = cy; sb[6] = z;
        sb[7] = px[L.dim + 0]; sb[8] = px[L.dim + 1]; sb[9] = px[L.dim + 2]; sb[10] = alpha;
        sarea[t] = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
        ssup[t] = 0;
    }
    if (t == 0) s_nkeep = 0;
    __syncthreads();
    for (int i = 0; i < nk; ++i) {
        if (ssup[i]) continue;
        for (int j = i + 1 + t; j < nk; j += CN_THREADS) {
            if (ssup[j]) continue;
            float xx1 = fmaxf(sbox[i][0], sbox[j][0]), yy1 = fmaxf(sbox[i][1], sbox[j][1]);
            float xx2 = fminf(sbox[i][2], sbox[j][2]), yy2 = fminf(sbox[i][3], sbox[j][3]);
            float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
            float inter = __fmul_rn(w, h);
            float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(sarea[i], sarea[j]), inter));
            if ((double)ovr > iou_thr) ssup[j] = 1;
        }
        if (t == 0) {
            int k = s_nkeep++;
            unsigned long long key = skey[i];
            unsigned int flat = (unsigned int)(key & 0xffffffffu);
            o_scores[(long long)b * out_cap + k] = __uint_as_float(~(unsigned int)(key >> 32));
            o_index[(long long)b * out_cap + k] = (int)flat;
            o_cls[(long long)b * out_cap + k] = (long long)(flat / (unsigned int)(H * W));
            float* op = o_boxes + ((long long)b * out_cap + k) * 11;
#pragma unroll
            for (int q = 0; q < 11; ++q) op[q] = sbox[i][q];
        }
        __syncthreads();
    }
    __syncthreads();
    if (t == 0) o_count[b] = s_nkeep;
}

}  // namespace vd3d

using namespace vd3d;

extern "C" long long vd3d_monoflex_decode_workspace(int B, int cap) { return (long long)B * cap * 8 + (long long)B * 4 + 64; }

extern "C" int vd3d_monoflex_decode(const float* heads, int B, int H, int W, int ncls, int cs, int hm_co, int bbox2d_co, int hps_co, int rot_co,
                                    int dim_co, int reg_co, int depth_co, int dunc_co, int cunc_co, const float* P2,
                                    float score_thr, double iou_thr, int K, float unc_lo, float unc_hi, float img_w, float img_h,
                                    int cap, void* wsp, int out_cap, float* out_scores, float* out_boxes, long long* out_cls,
                                    int* out_index, int* out_count, int* out_ncand, void* stream) {
    VD3D_REQUIRE(heads && P2 && wsp && out_scores && out_boxes && out_cls && out_index && out_count && out_ncand, "monoflex_decode: null pointer");
    VD3D_REQUIRE(B > 0 && H > 0 && W > 0 && ncls > 0 && K > 0 && K <= CN_MAXK && out_cap >= K && cap >= K && cap <= 8192, "monoflex_decode: bad shape (K <= 128, cap <= 8192)");
    VD3D_REQUIRE(score_thr > 0.f, "monoflex_decode: score_thr must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    CnLayout L{cs, hm_co, bbox2d_co, hps_co, rot_co, dim_co, reg_co, depth_co, dunc_co, cunc_co};
    unsigned long long* keys = (unsigned long long*)wsp;
    int* ncand = (int*)((unsigned char*)wsp + (long long)B * cap * 8);
    VD3D_CUDA(cudaMemsetAsync(ncand, 0, sizeof(int) * B, st));
    long long total = (long long)B * ncls * H * W;
    cn_peaks_kernel<<<cdiv(total, 256), 256, 0, st>>>(heads, B, H, W, ncls, L, score_thr, cap, keys, ncand);
    VD3D_CHECK_LAUNCH("cn_peaks");
    int cp2 = 1; while (cp2 < cap) cp2 <<= 1;
    size_t smem = (size_t)cp2 * 8;
    VD3D_CUDA(cudaFuncSetAttribute(cn_decode_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cn_decode_nms_kernel<<<B, CN_THREADS, smem, st>>>(heads, P2, H, W, ncls, L, cap, cp2, K, unc_lo, unc_hi, iou_thr, img_w, img_h, out_cap,
                                                     keys, ncand, out_scores, out_boxes, out_cls, out_index, out_count, out_ncand);
    VD3D_CHECK_LAUNCH("cn_decode_nms");
    return VD3D_OK;
}
