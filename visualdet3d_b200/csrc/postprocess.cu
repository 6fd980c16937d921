// Anchor filtering, box decode, clipping and NMS for the anchor-based 3-D head, batched, fixed capacity, no host
// round trips (the reference does ~30 tiny eager kernels and >= 6 D2H syncs per image,
// R/heads/detection_3d_head.py:341-400).
//
// Bit-exactness: index sets (useful mask, score threshold, prior validity, NMS keep) are decided by fp32
// comparisons whose operands are computed with the SAME operation order as the reference's eager ops and with
// FMA contraction disabled (explicit __fmul_rn/__fadd_rn/__fdiv_rn), so they only differ where the CUDA libm
// (expf/atan2f) differs from the host libm by an ulp on a knife edge.
#include "common.cuh"

namespace vd3d {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

// ---------------------------------------------------------------------------------------------------------
// useful mask (R/heads/anchors.py:93-111)
//   x3d = (xc*z - cx*z)/fy ; y3d = (yc*z - cy*z)/fy ; mask = any_t( y3d > ymin && y3d < ymax && |x3d| < xthr )
//   xc = mean(x1, x2) computed by torch as (x1 + x2) / 2
// ---------------------------------------------------------------------------------------------------------
__global__ void anchor_mask_kernel(const float* __restrict__ anchors, const float* __restrict__ means_z, const float* __restrict__ P2,
                                   int B, int N, int T, float y_min, float y_max, float x_thr, uint8_t* __restrict__ mask) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * N) return;
    int n = (int)(idx % N), b = (int)(idx / N);
    float4 a = ldg4(anchors + 4 * (long long)n);
    float xc = __fdiv_rn(add(a.x, a.z), 2.0f);
    float yc = __fdiv_rn(add(a.y, a.w), 2.0f);
    const float* P = P2 + 12 * b;
    float fy = P[5], cy = P[6], cx = P[2];
    bool any = false;
    for (int t = 0; t < T; ++t) {
        float z = __ldg(means_z + (long long)t * N + n);
        float x3 = __fdiv_rn(sub(mul(xc, z), mul(cx, z)), fy);
        float y3 = __fdiv_rn(sub(mul(yc, z), mul(cy, z)), fy);
        any = any || ((y3 > y_min) && (y3 < y_max) && (fabsf(x3) < x_thr));
    }
    mask[idx] = any ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------
// stage 1: score + threshold + validity -> unordered candidate list (atomic compaction).  The order is fixed
// afterwards by the sort key (score desc, anchor index asc) == torchvision's stable descending sort of the
// index-ordered candidate list.
// workspace layout per image: keys u64[cap], boxes f32[cap][11], labels i32[cap]
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_ref(float x) {
    // torch CPU sigmoid: 1 / (1 + exp(-x))
    return __fdiv_rn(1.0f, add(1.0f, expf(-x)));
}

struct DecodeWs {
    unsigned long long* keys;   // [B][cap]
    float* boxes;               // [B][cap][11]
    int* labels;                // [B][cap]
    int* ncand;                 // [B]
};

__global__ void decode_candidates_kernel(const float* __restrict__ cls, const float* __restrict__ reg, const float* __restrict__ anchors,
                                         const float* __restrict__ mean_std, const uint8_t* __restrict__ mask,
                                         int B, int N, int ncls, int T, float score_thr, float img_w, float img_h, int cap, DecodeWs ws) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * N) return;
    int n = (int)(idx % N), b = (int)(idx / N);
    if (!mask[idx]) return;
    const float* cp = cls + idx * (ncls + 1);
    float best = -1.f; int label = 0;
    for (int c = 0; c < ncls; ++c) {
        float p = sigmoidf_ref(__ldg(cp + c));
        if (p > best) { best = p; label = c; }     // first maximum wins (torch.max over dim)
    }
    if (!(best > score_thr)) return;
    const float* ms = mean_std + ((long long)n * T + label) * 12;   // [6][2]
    float z_mean = __ldg(ms + 0);
    if (!(z_mean > 0.f)) return;                                    // `mask = selected_mean_std[:,0,0] > 0` (:242)
    float alpha_score = sigmoidf_ref(__ldg(cp + ncls));

    float4 a = ldg4(anchors + 4 * (long long)n);
    const float* d = reg + idx * 12;
    float widths = sub(a.z, a.x), heights = sub(a.w, a.y);
    float ctr_x = add(a.x, mul(0.5f, widths)), ctr_y = add(a.y, mul(0.5f, heights));
    float dx = mul(__ldg(d + 0), 0.1f), dy = mul(__ldg(d + 1), 0.1f);
    float dw = mul(__ldg(d + 2), 0.2f), dh = mul(__ldg(d + 3), 0.2f);
    float pcx = add(ctr_x, mul(dx, widths)), pcy = add(ctr_y, mul(dy, heights));
    float pw = mul(expf(dw), widths), ph = mul(expf(dh), heights);
    float x1 = sub(pcx, mul(0.5f, pw)), y1 = sub(pcy, mul(0.5f, ph));
    float x2 = add(pcx, mul(0.5f, pw)), y2 = add(pcy, mul(0.5f, ph));
    float cx1 = add(ctr_x, mul(mul(__ldg(d + 4), 0.1f), widths));
    float cy1 = add(ctr_y, mul(mul(__ldg(d + 5), 0.1f), heights));
    float z = add(mul(__ldg(d + 6), __ldg(ms + 1)), z_mean);
    float sn = add(mul(__ldg(d + 7), __ldg(ms + 3)), __ldg(ms + 2));
    float cs = add(mul(__ldg(d + 8), __ldg(ms + 5)), __ldg(ms + 4));
    float alpha = __fdiv_rn(atan2f(sn, cs), 2.0f);
    float w3 = add(mul(__ldg(d + 9), __ldg(ms + 7)), __ldg(ms + 6));
    float h3 = add(mul(__ldg(d + 10), __ldg(ms + 9)), __ldg(ms + 8));
    float l3 = add(mul(__ldg(d + 11), __ldg(ms + 11)), __ldg(ms + 10));
    if (alpha_score < 0.5f) alpha = add(alpha, 3.14159265358979323846f);   // `+= np.pi` on an f32 tensor
    // ClipBoxes (R/utils/utils.py:181-196)
    x1 = fmaxf(x1, 0.f); y1 = fmaxf(y1, 0.f); x2 = fminf(x2, img_w); y2 = fminf(y2, img_h);

    int slot = atomicAdd(ws.ncand + b, 1);
    if (slot >= cap) return;                       // overflow: flagged through ncand > cap
    unsigned int sb = __float_as_uint(best);       // best in (0.75, 1] -> positive float, bit pattern monotone
    ws.keys[(long long)b * cap + slot] = ((unsigned long long)(~sb) << 32) | (unsigned int)n;   // ascending key = score desc, index asc
    float* bp = ws.boxes + ((long long)b * cap + slot) * 11;
    bp[0] = x1; bp[1] = y1; bp[2] = x2; bp[3] = y2; bp[4] = cx1; bp[5] = cy1; bp[6] = z; bp[7] = w3; bp[8] = h3; bp[9] = l3; bp[10] = alpha;
    ws.labels[(long long)b * cap + slot] = label;
}

// ---------------------------------------------------------------------------------------------------------
// stage 2 (one CTA per image): bitonic sort of (key, slot) in shared memory, greedy NMS by 64-row bit-matrix blocks, ordered write-out.
// torchvision CPU nms: areas = (x2-x1)*(y2-y1); inter = max(0, xx2-xx1)*max(0, yy2-yy1);
//                      ovr = inter / (area_i + area_j - inter); suppress j if ovr > thr.
// ---------------------------------------------------------------------------------------------------------
constexpr int NMS_THREADS = 1024;

__global__ void __launch_bounds__(NMS_THREADS) sort_nms_kernel(DecodeWs ws, int cap, int cap_pow2, double iou_thr,
                                                                float* __restrict__ out_scores, float* __restrict__ out_boxes,
                                                                int64_t* __restrict__ out_cls, int32_t* __restrict__ out_anchor,
                                                                int32_t* __restrict__ out_count, int32_t* __restrict__ out_ncand) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(sm_raw);            // [cap_pow2]
    unsigned long long* smask = skey + cap_pow2;                                         // [64][cap_pow2 / 64]: suppression bits of one 64-row block
    float4* sbox = reinterpret_cast<float4*>(smask + cap_pow2);                          // [cap_pow2]
    int* sslot = reinterpret_cast<int*>(sbox + cap_pow2);                                // [cap_pow2]
    float* sarea = reinterpret_cast<float*>(sslot + cap_pow2);                           // [cap_pow2]
    int* skeep = reinterpret_cast<int*>(sarea + cap_pow2);                               // [cap_pow2] sorted positions of the kept boxes
    __shared__ int s_nkeep;
    const int b = blockIdx.x, t = threadIdx.x;
    const int n = ws.ncand[b];
    if (t == 0) out_ncand[b] = n;
    if (n > cap) { if (t == 0) out_count[b] = -1; return; }
    int np2 = 64;                                   // sort size: next power of two >= n (the padding keys ~0 sort to the end)
    while (np2 < n) np2 <<= 1;

    for (int i = t; i < np2; i += NMS_THREADS) {
        skey[i] = (i < n) ? ws.keys[(long long)b * cap + i] : ~0ull;
        sslot[i] = i;
    }
    __syncthreads();
    // bitonic sort ascending on key
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < np2; i += NMS_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    bool up = ((i & k) == 0);
                    unsigned long long a = skey[i], c = skey[ixj];
                    if ((a > c) == up) {
                        skey[i] = c; skey[ixj] = a;
                        int s0 = sslot[i]; sslot[i] = sslot[ixj]; sslot[ixj] = s0;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = t; i < n; i += NMS_THREADS) {
        const float* bp = ws.boxes + ((long long)b * cap + sslot[i]) * 11;
        float4 bx = make_float4(bp[0], bp[1], bp[2], bp[3]);
        sbox[i] = bx;
        sarea[i] = mul(sub(bx.z, bx.x), sub(bx.w, bx.y));
    }
    __syncthreads();
    // Greedy sweep (box i is kept iff no earlier KEPT box suppresses it), 64 rows at a time: all threads build the suppression bit
    // matrix of the block (bit j of word w of row r: box 64w + j comes after row box i and IoU(i, 64w + j) > thr), then warp 0 walks
    // the 64 rows in order with the removed-set held as one 64-bit word per lane.
    const int nw = (n + 63) >> 6;                   // <= 32 words (cap <= 2048) ... or 64 for cap 4096: two words per lane
    const int wpr = cap_pow2 >> 6;                  // words per mask row
    unsigned long long removed0 = 0, removed1 = 0;  // warp 0: lane l holds words l and l + 32
    int nkeep = 0;
    const int warp = t >> 5, lane = t & 31;
    for (int c = 0; c < nw; ++c) {
        const int words = nw - c;
        for (int e = t; e < 64 * words; e += NMS_THREADS) {
            const int r = e & 63, w = c + (e >> 6);
            const int i = c * 64 + r;
            unsigned long long bits = 0;
            if (i < n) {
                const float4 bi = sbox[i];
                const float ai = sarea[i];
                const int j0 = w * 64;
                const int jend = min(64, n - j0);
                for (int j = (w == c ? r + 1 : 0); j < jend; ++j) {
                    const float4 bj = sbox[j0 + j];
                    float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
                    float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
                    float ww = fmaxf(0.f, sub(xx2, xx1)), hh = fmaxf(0.f, sub(yy2, yy1));
                    float inter = mul(ww, hh);
                    float ovr = __fdiv_rn(inter, sub(add(ai, sarea[j0 + j]), inter));
                    if ((double)ovr > iou_thr) bits |= 1ull << j;    // torchvision compares the f32 IoU with a double threshold
                }
            }
            smask[r * wpr + w] = bits;
        }
        __syncthreads();
        if (warp == 0) {
            const int rows = min(64, n - c * 64);
            for (int r = 0; r < rows; ++r) {
                const unsigned long long rc = c < 32 ? __shfl_sync(0xffffffffu, removed0, c) : __shfl_sync(0xffffffffu, removed1, c - 32);
                if ((rc >> r) & 1ull) continue;                       // warp-uniform
                if (lane == 0) skeep[nkeep] = c * 64 + r;
                ++nkeep;
                if (lane >= c && lane < nw) removed0 |= smask[r * wpr + lane];
                if (lane + 32 >= c && lane + 32 < nw) removed1 |= smask[r * wpr + lane + 32];
            }
        }
        __syncthreads();
    }
    if (t == 0) s_nkeep = nkeep;
    __syncthreads();
    nkeep = s_nkeep;
    // ordered write-out of the kept boxes, all threads
    for (int k = t; k < nkeep; k += NMS_THREADS) {
        const int i = skeep[k];
        const int slot = sslot[i];
        const unsigned long long key = skey[i];
        out_scores[(long long)b * cap + k] = __uint_as_float(~(unsigned int)(key >> 32));
        out_anchor[(long long)b * cap + k] = (int)(unsigned int)(key & 0xffffffffu);
        out_cls[(long long)b * cap + k] = (int64_t)ws.labels[(long long)b * cap + slot];
        const float* bp = ws.boxes + ((long long)b * cap + slot) * 11;
        float* op = out_boxes + ((long long)b * cap + k) * 11;
#pragma unroll
        for (int q = 0; q < 11; ++q) op[q] = bp[q];
    }
    if (t == 0) out_count[b] = nkeep;
}

// fixed-capacity detection record block for the multi-GPU all-gather: rec[b] = [count | kmax x (11 box floats, score, class)]
__global__ void pack_records_kernel(const float* __restrict__ scores, const float* __restrict__ boxes, const int64_t* __restrict__ cls,
                                    const int32_t* __restrict__ count, int cap, int kmax, float* __restrict__ rec, const int* __restrict__ range_flag) {
    const int b = blockIdx.x;
    int n = count[b];
    if (range_flag && *range_flag) n = -2;                            // -2: an activation left the fp16 range of the tensor-core engine (results invalid)
    float* r = rec + (long long)b * (1 + kmax * 13);
    if (threadIdx.x == 0) r[0] = (n > kmax) ? -1.0f : (float)n;       // -1: capacity overflow (also n == -1 from the decode stage)
    const int m = n < 0 ? 0 : (n > kmax ? 0 : n);
    for (int i = threadIdx.x; i < kmax * 13; i += blockDim.x) {
        int k = i / 13, q = i - k * 13;
        float v = 0.f;
        if (k < m) v = (q < 11) ? boxes[((long long)b * cap + k) * 11 + q] : (q == 11 ? scores[(long long)b * cap + k] : (float)cls[(long long)b * cap + k]);
        r[1 + i] = v;
    }
}

}  // namespace vd3d

using namespace vd3d;

extern "C" int vd3d_pack_records(const float* scores, const float* boxes, const int64_t* cls, const int32_t* count, int B, int cap, int kmax,
                                 float* rec, void* stream) {
    VD3D_REQUIRE(scores && boxes && cls && count && rec && B > 0 && cap > 0 && kmax > 0, "pack_records: bad args");
    pack_records_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(scores, boxes, cls, count, cap, kmax, rec, fp16_range_flag());
    VD3D_CHECK_LAUNCH("pack_records");
    return VD3D_OK;
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

extern "C" int vd3d_anchor_mask(const float* anchors, const float* means_z, const float* P2, int B, int N, int T,
                                float y_min, float y_max, float x_thr, uint8_t* mask, void* stream) {
    VD3D_REQUIRE(anchors && means_z && P2 && mask && B > 0 && N > 0 && T > 0, "anchor_mask: bad args");
    anchor_mask_kernel<<<cdiv((long long)B * N, 256), 256, 0, (cudaStream_t)stream>>>(anchors, means_z, P2, B, N, T, y_min, y_max, x_thr, mask);
    VD3D_CHECK_LAUNCH("anchor_mask");
    return VD3D_OK;
}

extern "C" long long vd3d_decode_nms_workspace(int B, int cap) {
    // keys u64 + boxes 11 f32 + labels i32 per slot, + ncand i32 per image (16-byte aligned blocks)
    long long per = (long long)cap * (8 + 44 + 4);
    return (long long)B * per + 16 * ((B * 4 + 15) / 16) + 64;
}

extern "C" int vd3d_decode_nms(const float* cls, const float* reg, const float* anchors, const float* mean_std,
                               const uint8_t* mask, int B, int N, int ncls, int T, float score_thr, double iou_thr,
                               float img_w, float img_h, int cap, void* wsp,
                               float* out_scores, float* out_boxes, int64_t* out_cls, int32_t* out_anchor,
                               int32_t* out_count, int32_t* out_ncand, void* stream) {
    VD3D_REQUIRE(cls && reg && anchors && mean_std && mask && wsp && out_scores && out_boxes && out_cls && out_anchor && out_count && out_ncand,
                 "decode_nms: null pointer");
    VD3D_REQUIRE(B > 0 && N > 0 && ncls > 0 && ncls <= T && cap > 0 && cap <= 4096, "decode_nms: bad shape (cap must be <= 4096)");
    VD3D_REQUIRE(score_thr > 0.f, "decode_nms: score_thr must be positive (sort key relies on positive scores)");
    cudaStream_t st = (cudaStream_t)stream;
    DecodeWs ws;
    unsigned char* p = (unsigned char*)wsp;
    ws.keys = (unsigned long long*)p; p += (long long)B * cap * 8;
    ws.boxes = (float*)p; p += (long long)B * cap * 44;
    ws.labels = (int*)p; p += (long long)B * cap * 4;
    ws.ncand = (int*)p;
    VD3D_CUDA(cudaMemsetAsync(ws.ncand, 0, sizeof(int) * B, st));
    decode_candidates_kernel<<<cdiv((long long)B * N, 256), 256, 0, st>>>(cls, reg, anchors, mean_std, mask, B, N, ncls, T,
                                                                         score_thr, img_w, img_h, cap, ws);
    VD3D_CHECK_LAUNCH("decode_candidates");
    int cp2 = next_pow2(cap);
    if (cp2 < 64) cp2 = 64;                      // the 64-row bit-matrix blocks of the NMS sweep
    size_t smem = (size_t)cp2 * (8 + 8 + 16 + 4 + 4 + 4) + 16;
    VD3D_CUDA(cudaFuncSetAttribute(sort_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sort_nms_kernel<<<B, NMS_THREADS, smem, st>>>(ws, cap, cp2, iou_thr, out_scores, out_boxes, out_cls, out_anchor, out_count, out_ncand);
    VD3D_CHECK_LAUNCH("sort_nms");
    return VD3D_OK;
}
