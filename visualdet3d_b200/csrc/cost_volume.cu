// Stereo cost volumes (R/lib/PSM_cost_volume.py): correlation volume (PSMCosineModule) and the scale-16
// concat volume fused into its two Conv3d layers (CostVolume).  HBM-bound: every input element is read once
// from HBM (re-reads are served from shared memory / L2), every output element written once.
#include "common.cuh"
#include <cstdlib>

namespace vd3d {

// ---------------------------------------------------------------------------------------------------------
// PSMCosine, NHWC.   out[b,h,w,i] = (1/C) sum_c L[b,h,w,c] * R[b,h,w-i,c]  (w >= i), 0 otherwise.
//
// v2 kernel: CTA = (b, h, 64-pixel tile); the L tile and the R window (64 + 24 pixels) are transposed on the
// way into shared memory to [c][w] so that each thread register-tiles 4 pixels x 24 disparities and reuses every
// shared-memory word 24x / 4x.  A warp = 4 pixel groups x 8 channel splits; the 8 partial sums are combined with a
// transpose-reduce (84 shuffles for 96 accumulators).  See DESIGN.md "cost volume".
// ---------------------------------------------------------------------------------------------------------
constexpr int PSM_TW = 64;      // pixels per CTA
constexpr int PSM_D = 24;       // disparities (both PSMCosine layers of the path have D = 24)
constexpr int PSM_THREADS = 128;
constexpr int PSM_RW = PSM_TW + PSM_D;             // R window incl. halo, 88 columns: cols [w0-24, w0+64)
constexpr int PSM_LP = PSM_TW + 2;                 // row pitch of Ls, == 2 (mod 32) -> conflict-free transposed stores
constexpr int PSM_RP = PSM_RW + 10;                // 98 == 2 (mod 32)

template <int C>
__global__ void __launch_bounds__(PSM_THREADS, 4) psm_cosine_nhwc_kernel(
    const float* __restrict__ L, const float* __restrict__ R, int H, int W, int lr_cs, int lr_co,
    float* __restrict__ out, int out_cs, int out_co) {
    extern __shared__ __align__(16) float smem[];
    float* Ls = smem;                       // [C][PSM_LP]
    float* Rs = smem + C * PSM_LP;          // [C][PSM_RP]
    const int t = threadIdx.x;
    const int w0 = blockIdx.x * PSM_TW;
    const int h = blockIdx.y, b = blockIdx.z;
    const long long rowbase = ((long long)b * H + h) * W;

    // ---- load + transpose: lanes = (pixel p in 0..7) x (channel quad cq in 0..3) -> 8 x 64B segments / warp-load
    {
        // lane -> (pixel p, channel quad cq) with cq fastest: 4 adjacent lanes read 64 contiguous bytes, so a quarter-warp
        // touches 2 lines instead of 8 (round-1 ncu: 32 L1 data-pipe wavefronts per LDG.128 with p fastest).  The transposed
        // stores stay conflict-free: bank = 8*cq + p + 2*j is the same set for any lane order.
        const int cq = t & 3, p = (t >> 2) & 7, wrp = t >> 5;   // 4 warps
        constexpr int CQ = C / 4;                                  // channel quads per pixel
        // L tile: 64 pixels x CQ quads, R window: 88 pixels x CQ quads; per pass a warp covers 8 pixels x 4 quads.
        // All global loads of a batch are issued before the first shared store so NB 16-byte requests are in
        // flight per thread (the round-1 ncu capture showed 71% long-scoreboard stalls with one load in flight).
        constexpr int NWARP = PSM_THREADS / 32;
        constexpr int QG = CQ / 4;                                   // quad groups per pixel group
        constexpr int NL = (PSM_TW / 8) * QG / NWARP;                // L iterations per warp  (8 for C=64)
        constexpr int NR = ((PSM_RW / 8) * QG + NWARP - 1) / NWARP;  // R iterations per warp  (11 for C=64)
        constexpr int NT = NL + NR;                                  // 19 for C = 64, 38 for C = 128
        constexpr int NB = (NT <= 20) ? NT : (NT + 1) / 2;          // loads in flight per batch (one batch for C = 64)
        static_assert((PSM_TW / 8) * QG % NWARP == 0, "L tile must split evenly over the warps");
#pragma unroll
        for (int base = 0; base < NT; base += NB) {
            float4 v[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int i = base + u;                               // compile-time after unrolling
                if (i < NL) {
                    int it = wrp + i * NWARP;
                    int pg = it / QG, qg = it - pg * QG;
                    int w = w0 + pg * 8 + p;
                    v[u] = (w < W) ? ldg4(L + (rowbase + w) * lr_cs + lr_co + 4 * (qg * 4 + cq)) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else if (i < NT) {
                    int it = wrp + (i - NL) * NWARP;
                    int pg = it / QG, qg = it - pg * QG;
                    int w = w0 - PSM_D + pg * 8 + p;
                    bool ok = it < (PSM_RW / 8) * QG && w >= 0 && w < W;
                    v[u] = ok ? ldg4(R + (rowbase + w) * lr_cs + lr_co + 4 * (qg * 4 + cq)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int i = base + u;
                if (i < NL) {
                    int it = wrp + i * NWARP;
                    int pg = it / QG, qg = it - pg * QG;
                    float* d = Ls + (4 * (qg * 4 + cq)) * PSM_LP + pg * 8 + p;
                    d[0] = v[u].x; d[PSM_LP] = v[u].y; d[2 * PSM_LP] = v[u].z; d[3 * PSM_LP] = v[u].w;
                } else if (i < NT) {
                    int it = wrp + (i - NL) * NWARP;
                    if (it < (PSM_RW / 8) * QG) {
                        int pg = it / QG, qg = it - pg * QG;
                        float* d = Rs + (4 * (qg * 4 + cq)) * PSM_RP + pg * 8 + p;
                        d[0] = v[u].x; d[PSM_RP] = v[u].y; d[2 * PSM_RP] = v[u].z; d[3 * PSM_RP] = v[u].w;
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- compute: lane = pg*8 + s ; pixel group pg (4 pixels at 16*pg + 4*warp), channel split s (c = s + 8k)
    const int lane = t & 31, wrp = t >> 5;
    const int s = lane & 7, pg = lane >> 3;
    const int px0 = 16 * pg + 4 * wrp;            // first of this thread's 4 pixels (tile-local)
    float acc[4][PSM_D];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < PSM_D; ++j) acc[i][j] = 0.f;

#pragma unroll 1
    for (int c = s; c < C; c += 8) {
        const float* lr = Ls + c * PSM_LP + px0;
        const float* rr = Rs + c * PSM_RP + px0;   // Rs column j <-> pixel (w0 - 24 + j): pixel px0+i-d <-> column px0+i-d+24
        float l[4], r[28];
        {
            float2 a = *reinterpret_cast<const float2*>(lr), bq = *reinterpret_cast<const float2*>(lr + 2);
            l[0] = a.x; l[1] = a.y; l[2] = bq.x; l[3] = bq.y;
        }
#pragma unroll
        for (int j = 0; j < 28; j += 2) {
            float2 v = *reinterpret_cast<const float2*>(rr + j);
            r[j] = v.x; r[j + 1] = v.y;
        }
        // r[j] holds pixel (px0 - 24 + j); disparity d for pixel i needs pixel px0+i-d -> r[24 + i - d]
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int d = 0; d < PSM_D; ++d) acc[i][d] = fmaf(l[i], r[24 + i - d], acc[i][d]);
    }

    // ---- transpose-reduce across the 8 channel splits (lanes differing in bits 0..2) -------------------------
    // after step k each lane keeps half of its values; lane s ends with the 12 values v[12*... ] of its share.
    float* a = &acc[0][0];   // 96 values, index = i*24 + d
    // step 1: partner = lane ^ 4 ; lanes with (s&4)==0 keep [0,48), others keep [48,96)
#pragma unroll
    for (int j = 0; j < 48; ++j) {
        bool hi = (s & 4) != 0;
        float send = hi ? a[j] : a[j + 48];
        float recv = __shfl_xor_sync(0xffffffffu, send, 4);
        a[j] = (hi ? a[j + 48] : a[j]) + recv;
    }
#pragma unroll
    for (int j = 0; j < 24; ++j) {
        bool hi = (s & 2) != 0;
        float send = hi ? a[j] : a[j + 24];
        float recv = __shfl_xor_sync(0xffffffffu, send, 2);
        a[j] = (hi ? a[j + 24] : a[j]) + recv;
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        bool hi = (s & 1) != 0;
        float send = hi ? a[j] : a[j + 12];
        float recv = __shfl_xor_sync(0xffffffffu, send, 1);
        a[j] = (hi ? a[j + 12] : a[j]) + recv;
    }
    // lane s now owns flat indices [12*rank, 12*rank+12) with rank = (s&4 ? 4:0) + (s&2 ? 2:0) + (s&1) = s
    // -> pixel i = s/2, disparities (s&1)*12 .. +12
    {
        const int i = s >> 1, d0 = (s & 1) * 12;
        const int w = w0 + px0 + i;
        if (w < W) {
            float* op = out + (rowbase + w) * out_cs + out_co + d0;
            const float inv = 1.0f / (float)C;
            float o[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) o[j] = (w >= d0 + j) ? a[j] * inv : 0.f;
#pragma unroll
            for (int j = 0; j < 12; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// v3: same tile / shared layout as v2, but 256 threads per CTA: a thread register-tiles 4 pixels x 12 disparities
// (48 accumulators, <= 85 registers) so 3 CTAs = 24 warps are resident per SM instead of 16 (the round-1 ncu capture of v2
// showed 46% issue-slot utilisation with 16 warps: latency bound).  Warps 0..3 take disparities 0..11, warps 4..7 take 12..23.
// ---------------------------------------------------------------------------------------------------------
constexpr int PSM3_THREADS = 256;

template <int C>
__global__ void __launch_bounds__(PSM3_THREADS, 3) psm_cosine_nhwc_v3_kernel(
    const float* __restrict__ L, const float* __restrict__ R, int H, int W, int lr_cs, int lr_co,
    float* __restrict__ out, int out_cs, int out_co) {
    extern __shared__ __align__(16) float smem[];
    float* Ls = smem;                       // [C][PSM_LP]
    float* Rs = smem + C * PSM_LP;          // [C][PSM_RP]
    const int t = threadIdx.x;
    const int w0 = blockIdx.x * PSM_TW;
    const int h = blockIdx.y, b = blockIdx.z;
    const long long rowbase = ((long long)b * H + h) * W;
    {
        const int cq = t & 3, p = (t >> 2) & 7, wrp = t >> 5;      // 8 warps; lane -> (pixel p, channel quad cq), cq fastest
        constexpr int CQ = C / 4, QG = CQ / 4, NWARP = PSM3_THREADS / 32;
        constexpr int NLI = (PSM_TW / 8) * QG;                       // L work items per CTA (32 for C = 64)
        constexpr int NRI = (PSM_RW / 8) * QG;                       // R work items          (44 for C = 64)
        constexpr int NT = (NLI + NRI + NWARP - 1) / NWARP;          // items per warp        (10 for C = 64)
        constexpr int NB = (NT <= 12) ? NT : (NT + 1) / 2;
#pragma unroll
        for (int base = 0; base < NT; base += NB) {
            float4 v[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int it = wrp + (base + u) * NWARP;             // combined item index: [0, NLI) = L, [NLI, NLI+NRI) = R
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (base + u < NT && it < NLI) {
                    int pg = it / QG, qg = it - pg * QG;
                    int w = w0 + pg * 8 + p;
                    if (w < W) v[u] = ldg4(L + (rowbase + w) * lr_cs + lr_co + 4 * (qg * 4 + cq));
                } else if (base + u < NT && it < NLI + NRI) {
                    int ir = it - NLI;
                    int pg = ir / QG, qg = ir - pg * QG;
                    int w = w0 - PSM_D + pg * 8 + p;
                    if (w >= 0 && w < W) v[u] = ldg4(R + (rowbase + w) * lr_cs + lr_co + 4 * (qg * 4 + cq));
                }
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int it = wrp + (base + u) * NWARP;
                if (base + u < NT && it < NLI) {
                    int pg = it / QG, qg = it - pg * QG;
                    float* d = Ls + (4 * (qg * 4 + cq)) * PSM_LP + pg * 8 + p;
                    d[0] = v[u].x; d[PSM_LP] = v[u].y; d[2 * PSM_LP] = v[u].z; d[3 * PSM_LP] = v[u].w;
                } else if (base + u < NT && it < NLI + NRI) {
                    int ir = it - NLI;
                    int pg = ir / QG, qg = ir - pg * QG;
                    float* d = Rs + (4 * (qg * 4 + cq)) * PSM_RP + pg * 8 + p;
                    d[0] = v[u].x; d[PSM_RP] = v[u].y; d[2 * PSM_RP] = v[u].z; d[3 * PSM_RP] = v[u].w;
                }
            }
        }
    }
    __syncthreads();

    const int lane = t & 31, wrp = t >> 5;
    const int s = lane & 7, pg = lane >> 3;
    const int px0 = 16 * pg + 4 * (wrp & 3);       // first of this thread's 4 pixels (tile-local)
    const int d0 = 12 * (wrp >> 2);                // first of this thread's 12 disparities
    float acc[4][12];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 12; ++j) acc[i][j] = 0.f;
#pragma unroll 1
    for (int c = s; c < C; c += 8) {
        const float* lr = Ls + c * PSM_LP + px0;
        const float* rr = Rs + c * PSM_RP + px0 + 12 - d0;     // column of pixel (px0 - d0 - 12): even -> 8-byte aligned
        float l[4], r[16];
        {
            float2 a = *reinterpret_cast<const float2*>(lr), bq = *reinterpret_cast<const float2*>(lr + 2);
            l[0] = a.x; l[1] = a.y; l[2] = bq.x; l[3] = bq.y;
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            float2 v = *reinterpret_cast<const float2*>(rr + j);
            r[j] = v.x; r[j + 1] = v.y;
        }
        // r[j] = pixel (px0 - d0 - 12 + j); (pixel i, disparity d0 + dd) needs pixel px0 + i - d0 - dd -> r[12 + i - dd]
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int dd = 0; dd < 12; ++dd) acc[i][dd] = fmaf(l[i], r[12 + i - dd], acc[i][dd]);
    }
    float* a = &acc[0][0];   // 48 values, index = i*12 + dd
#pragma unroll
    for (int j = 0; j < 24; ++j) {
        bool hi = (s & 4) != 0;
        float send = hi ? a[j] : a[j + 24];
        float recv = __shfl_xor_sync(0xffffffffu, send, 4);
        a[j] = (hi ? a[j + 24] : a[j]) + recv;
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        bool hi = (s & 2) != 0;
        float send = hi ? a[j] : a[j + 12];
        float recv = __shfl_xor_sync(0xffffffffu, send, 2);
        a[j] = (hi ? a[j + 12] : a[j]) + recv;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        bool hi = (s & 1) != 0;
        float send = hi ? a[j] : a[j + 6];
        float recv = __shfl_xor_sync(0xffffffffu, send, 1);
        a[j] = (hi ? a[j + 6] : a[j]) + recv;
    }
    // lane s owns flat indices [6s, 6s + 6): pixel i = s / 2, disparities d0 + 6 (s & 1) .. + 6
    {
        const int i = s >> 1, dbase = d0 + 6 * (s & 1);
        const int w = w0 + px0 + i;
        if (w < W) {
            float* op = out + (rowbase + w) * out_cs + out_co + dbase;
            const float inv = 1.0f / (float)C;
            float o[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) o[j] = (w >= dbase + j) ? a[j] * inv : 0.f;
#pragma unroll
            for (int j = 0; j < 6; j += 2) *reinterpret_cast<float2*>(op + j) = make_float2(o[j], o[j + 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// v4: persistent, warp-specialised, bulk-copy fed.
//   * a tile's L pixels (64 x C floats) and R window (88 x C floats) are CONTIGUOUS in NHWC, so one elected producer thread
//     brings each in with a single cp.async.bulk (1-D TMA) into a [pixel][C] staging ring guarded by full/empty mbarriers:
//     no registers, no LSU wavefronts, no transposition, HBM latency hidden behind the previous tile's math;
//   * 12 consumer warps read the staging tile with conflict-free LDS.128 (the 8 lanes of a quarter-warp take the 8 channel
//     quads of ONE pixel = 128 contiguous bytes); a thread register-tiles 4 pixels x 8 disparities x 4 channels, the 8 lanes'
//     partial sums are combined with a 28-shuffle transpose-reduce and each lane stores one float4.
//   Out-of-image R pixels (w - i < 0) are never copied: they only feed outputs that are written as exact zeros.
// ---------------------------------------------------------------------------------------------------------
constexpr int PSM4_CONSUMERS = 384;                 // 12 warps: 4 pixel-quad offsets x 3 disparity groups of 8
constexpr int PSM4_THREADS = PSM4_CONSUMERS + 32;   // + 1 producer warp

__device__ __forceinline__ uint32_t psm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void psm_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "PSM_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PSM_DONE;\n\t"
        "bra PSM_WAIT;\n\t"
        "PSM_DONE:\n\t"
        "}\n" ::"r"(psm_smem_u32(bar)), "r"(parity) : "memory");
}

template <int C, int STAGES>
__global__ void __launch_bounds__(PSM4_THREADS, (C == 64) ? 2 : 1) psm_cosine_nhwc_v4_kernel(
    const float* __restrict__ L, const float* __restrict__ R, int B, int H, int W, int lr_cs, int lr_co,
    float* __restrict__ out, int out_cs, int out_co) {
    extern __shared__ __align__(128) uint8_t psm_smem[];
    constexpr int L_FLOATS = PSM_TW * C, R_FLOATS = PSM_RW * C, STAGE_FLOATS = L_FLOATS + R_FLOATS;
    float* stage0 = reinterpret_cast<float*>(psm_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(psm_smem + (size_t)STAGES * STAGE_FLOATS * 4);
    uint64_t* empty = full + STAGES;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int tiles_w = (W + PSM_TW - 1) / PSM_TW;
    const long long ntiles = (long long)B * H * tiles_w;
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(psm_smem_u32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(psm_smem_u32(&empty[s])), "r"(PSM4_CONSUMERS / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == PSM4_CONSUMERS / 32) {
        // ================= producer (one lane) =================
        if (lane == 0) {
            int k = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
                const int s = k % STAGES, ph = (k / STAGES) & 1;
                psm_mbar_wait(&empty[s], ph ^ 1);
                const int tw = (int)(tile % tiles_w); const long long row = tile / tiles_w;     // row = b*H + h
                const int w0 = tw * PSM_TW;
                const int nl = min(PSM_TW, W - w0);                     // valid L pixels
                const int r_lo = max(w0 - PSM_D, 0);                    // first valid R pixel
                const int nr = min(w0 + PSM_TW, W) - r_lo;              // valid R pixels
                float* ls = stage0 + (size_t)s * STAGE_FLOATS;
                float* rs = ls + L_FLOATS + (size_t)(r_lo - (w0 - PSM_D)) * C;       // slot j <-> pixel w0 - 24 + j
                const uint32_t bytes = (uint32_t)(nl + nr) * C * 4;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(psm_smem_u32(&full[s])), "r"(bytes) : "memory");
                const float* lsrc = L + (row * W + w0) * (long long)lr_cs + lr_co;
                const float* rsrc = R + (row * W + r_lo) * (long long)lr_cs + lr_co;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(psm_smem_u32(ls)), "l"(lsrc), "r"((uint32_t)nl * C * 4), "r"(psm_smem_u32(&full[s])) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(psm_smem_u32(rs)), "l"(rsrc), "r"((uint32_t)nr * C * 4), "r"(psm_smem_u32(&full[s])) : "memory");
            }
        }
        return;
    }

    // ================= consumers =================
    const int s8 = lane & 7, pg = lane >> 3;
    const int px0 = 16 * pg + 4 * (warp & 3);          // first of this thread's 4 pixels (tile-local)
    const int d0 = 8 * (warp >> 2);                    // first of this thread's 8 disparities
    int k = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
        const int s = k % STAGES, ph = (k / STAGES) & 1;
        const int tw = (int)(tile % tiles_w); const long long row = tile / tiles_w;
        const int w0 = tw * PSM_TW;
        psm_mbar_wait(&full[s], ph);
        const float* ls = stage0 + (size_t)s * STAGE_FLOATS;
        const float* rs = ls + L_FLOATS;
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 1
        for (int cb = 0; cb < C; cb += 32) {
            float4 l[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) l[i] = *reinterpret_cast<const float4*>(ls + (px0 + i) * C + cb + 4 * s8);
            // R slot of pixel p is p + 24; pixels needed: px0 - d0 - 7 + j, j = 0..10
            const float* rb = rs + (px0 - d0 - 7 + PSM_D) * C + cb + 4 * s8;
#pragma unroll
            for (int j = 0; j < 11; ++j) {
                const float4 r = *reinterpret_cast<const float4*>(rb + j * C);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int dd = i + 7 - j;                   // (pixel i, disparity d0 + dd) uses pixel px0 + i - d0 - dd
                    if (dd >= 0 && dd < 8) {
                        acc[i][dd] = fmaf(l[i].x, r.x, acc[i][dd]);
                        acc[i][dd] = fmaf(l[i].y, r.y, acc[i][dd]);
                        acc[i][dd] = fmaf(l[i].z, r.z, acc[i][dd]);
                        acc[i][dd] = fmaf(l[i].w, r.w, acc[i][dd]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(psm_smem_u32(&empty[s])) : "memory");
        // transpose-reduce over the 8 channel-quad lanes: 32 values -> 4 per lane
        float* a = &acc[0][0];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            bool hi = (s8 & 4) != 0;
            float send = hi ? a[j] : a[j + 16];
            float recv = __shfl_xor_sync(0xffffffffu, send, 4);
            a[j] = (hi ? a[j + 16] : a[j]) + recv;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            bool hi = (s8 & 2) != 0;
            float send = hi ? a[j] : a[j + 8];
            float recv = __shfl_xor_sync(0xffffffffu, send, 2);
            a[j] = (hi ? a[j + 8] : a[j]) + recv;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bool hi = (s8 & 1) != 0;
            float send = hi ? a[j] : a[j + 4];
            float recv = __shfl_xor_sync(0xffffffffu, send, 1);
            a[j] = (hi ? a[j + 4] : a[j]) + recv;
        }
        // lane s8 owns flat indices [4 s8, 4 s8 + 4): pixel i = s8 / 2, disparities d0 + 4 (s8 & 1) .. + 4
        const int i = s8 >> 1, dbase = d0 + 4 * (s8 & 1);
        const int w = w0 + px0 + i;
        if (w < W) {
            const float inv = 1.0f / (float)C;
            float4 o;
            o.x = (w >= dbase + 0) ? a[0] * inv : 0.f;
            o.y = (w >= dbase + 1) ? a[1] * inv : 0.f;
            o.z = (w >= dbase + 2) ? a[2] * inv : 0.f;
            o.w = (w >= dbase + 3) ? a[3] * inv : 0.f;
            *reinterpret_cast<float4*>(out + (row * W + w) * (long long)out_cs + out_co + dbase) = o;
        }
    }
}

// Generic fallback (any C % 4 == 0, any D <= 64): thread = (pixel, disparity), reads through L1.
__global__ void psm_cosine_nhwc_generic_kernel(const float* __restrict__ L, const float* __restrict__ R, long long npix, int W, int C,
                                               int lr_cs, int lr_co, int D, float* __restrict__ out, int out_cs, int out_co) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * D) return;
    int i = (int)(idx % D); long long pix = idx / D;
    int w = (int)(pix % W);
    float s = 0.f;
    if (w >= i) {
        const float* lp = L + pix * lr_cs + lr_co;
        const float* rp = R + (pix - i) * lr_cs + lr_co;
        for (int c = 0; c < C; c += 4) {
            float4 a = ldg4(lp + c), bq = ldg4(rp + c);
            s = fmaf(a.x, bq.x, s); s = fmaf(a.y, bq.y, s); s = fmaf(a.z, bq.z, s); s = fmaf(a.w, bq.w, s);
        }
        s = s / (float)C;
    }
    out[pix * out_cs + out_co + i] = s;
}

// NCHW op-level mirror: thread = (b, h, w); loops c with coalesced reads along w.
__global__ void psm_cosine_nchw_kernel(const float* __restrict__ L, const float* __restrict__ R, int B, int C, int H, int W, int D,
                                       float* __restrict__ out) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long HW = (long long)H * W;
    if (idx >= (long long)B * HW) return;
    int w = (int)(idx % W); long long r = idx / W; int h = (int)(r % H); int b = (int)(r / H);
    for (int d0 = 0; d0 < D; d0 += 8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int c = 0; c < C; ++c) {
            const float* lp = L + ((long long)b * C + c) * HW + (long long)h * W;
            const float* rp = R + ((long long)b * C + c) * HW + (long long)h * W;
            float l = __ldg(lp + w);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int i = d0 + j;
                if (i < D && w >= i) acc[j] = fmaf(l, __ldg(rp + w - i), acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int i = d0 + j;
            if (i < D) out[((long long)b * D + i) * HW + (long long)h * W + w] = (w >= i) ? acc[j] / (float)C : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// concat volume + Conv3d #1 (2F -> F) : thread = (b, d, h, w), F = 8 outputs.  The volume is gathered from lf / rf.
// ---------------------------------------------------------------------------------------------------------
constexpr int CV_F = 8;

__global__ void __launch_bounds__(128) concat_conv3d_1_kernel(const float* __restrict__ lf, const float* __restrict__ rf,
                                                               const float* __restrict__ w1, const float* __restrict__ b1,
                                                               float* __restrict__ mid, int B, int D, int H, int W) {
    __shared__ __align__(16) float ws[27 * 2 * CV_F * CV_F];
    for (int i = threadIdx.x; i < 27 * 2 * CV_F * CV_F; i += blockDim.x) ws[i] = w1[i];
    __syncthreads();
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * D * H * W;
    if (idx >= total) return;
    int w = (int)(idx % W); long long r = idx / W; int h = (int)(r % H); r /= H; int d = (int)(r % D); int b = (int)(r / D);
    float acc[CV_F];
#pragma unroll
    for (int f = 0; f < CV_F; ++f) acc[f] = 0.f;
    for (int kd = 0; kd < 3; ++kd) {
        int dd = d - 1 + kd;
        if (dd < 0 || dd >= D) continue;
        for (int kh = 0; kh < 3; ++kh) {
            int hh = h - 1 + kh;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                int ww = w - 1 + kw;
                if (ww < 0 || ww >= W || ww < dd) continue;          // volume is zero where w < disparity
                const float* lp = lf + (((long long)b * H + hh) * W + ww) * CV_F;
                const float* rp = rf + (((long long)b * H + hh) * W + ww - dd) * CV_F;
                float v[2 * CV_F];
                float4 x0 = ldg4(lp), x1 = ldg4(lp + 4), y0 = ldg4(rp), y1 = ldg4(rp + 4);
                v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
                v[8] = y0.x; v[9] = y0.y; v[10] = y0.z; v[11] = y0.w; v[12] = y1.x; v[13] = y1.y; v[14] = y1.z; v[15] = y1.w;
                const float* wt = ws + ((kd * 3 + kh) * 3 + kw) * (2 * CV_F * CV_F);
#pragma unroll
                for (int c = 0; c < 2 * CV_F; ++c)
#pragma unroll
                    for (int f = 0; f < CV_F; ++f) acc[f] = fmaf(v[c], wt[c * CV_F + f], acc[f]);
            }
        }
    }
    float* op = mid + idx * CV_F;
#pragma unroll
    for (int f = 0; f < CV_F; ++f) acc[f] = fmaxf(acc[f] + __ldg(b1 + f), 0.f);
    *reinterpret_cast<float4*>(op) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(op + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

__global__ void __launch_bounds__(128) conv3d_2_kernel(const float* __restrict__ mid, const float* __restrict__ w2, const float* __restrict__ b2,
                                                        float* __restrict__ out, int B, int D, int H, int W, int out_cs, int out_co) {
    __shared__ __align__(16) float ws[27 * CV_F * CV_F];
    for (int i = threadIdx.x; i < 27 * CV_F * CV_F; i += blockDim.x) ws[i] = w2[i];
    __syncthreads();
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * D * H * W;
    if (idx >= total) return;
    // thread order here: d fastest so the D outputs of one (pixel, f) are written by neighbouring threads
    int d = (int)(idx % D); long long r = idx / D; int w = (int)(r % W); r /= W; int h = (int)(r % H); int b = (int)(r / H);
    float acc[CV_F];
#pragma unroll
    for (int f = 0; f < CV_F; ++f) acc[f] = 0.f;
    for (int kd = 0; kd < 3; ++kd) {
        int dd = d - 1 + kd;
        if (dd < 0 || dd >= D) continue;
        for (int kh = 0; kh < 3; ++kh) {
            int hh = h - 1 + kh;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                int ww = w - 1 + kw;
                if (ww < 0 || ww >= W) continue;
                const float* mp = mid + ((((long long)b * D + dd) * H + hh) * W + ww) * CV_F;
                float4 x0 = ldg4(mp), x1 = ldg4(mp + 4);
                float v[CV_F] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const float* wt = ws + ((kd * 3 + kh) * 3 + kw) * (CV_F * CV_F);
#pragma unroll
                for (int c = 0; c < CV_F; ++c)
#pragma unroll
                    for (int f = 0; f < CV_F; ++f) acc[f] = fmaf(v[c], wt[c * CV_F + f], acc[f]);
            }
        }
    }
    float* op = out + (((long long)b * H + h) * W + w) * out_cs + out_co + d;
#pragma unroll
    for (int f = 0; f < CV_F; ++f) op[f * D] = fmaxf(acc[f] + __ldg(b2 + f), 0.f);
}

}  // namespace vd3d

using namespace vd3d;

extern "C" int vd3d_psm_cosine_nhwc(const float* L, const float* R, int B, int H, int W, int C, int lr_cs, int lr_co,
                                    int D, float* out, int out_cs, int out_co, void* stream) {
    VD3D_REQUIRE(L && R && out, "psm_cosine: null pointer");
    VD3D_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && D > 0 && D <= 64, "psm_cosine: bad shape");
    VD3D_REQUIRE(C % 4 == 0 && lr_cs % 4 == 0 && lr_co % 4 == 0, "psm_cosine: C, pitch, offset must be multiples of 4");
    cudaStream_t st = (cudaStream_t)stream;
    bool fast = (D == PSM_D) && (C == 64 || C == 128) && out_cs % 4 == 0 && out_co % 4 == 0 && H < 65536 && B < 65536;
    static const int variant = getenv("VD3D_PSM_VARIANT") ? atoi(getenv("VD3D_PSM_VARIANT")) : 4;
    // v4 needs 16-byte aligned, channel-dense rows (a tile's pixels must be one contiguous run of C floats each)
    const bool dense = lr_cs == C && lr_co == 0 && ((uintptr_t)L & 15) == 0 && ((uintptr_t)R & 15) == 0;
    if (fast && variant == 4 && dense) {
        const long long ntiles = (long long)B * H * cdiv(W, PSM_TW);
        if (C == 64) {
            constexpr int ST = 2;
            size_t smem = (size_t)ST * (PSM_TW + PSM_RW) * 64 * 4 + 2 * ST * 8;
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_v4_kernel<64, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_v4_kernel<64, ST>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            int grid = (int)(ntiles < 2 * kNumSMs ? ntiles : 2 * kNumSMs);
            psm_cosine_nhwc_v4_kernel<64, ST><<<grid, PSM4_THREADS, smem, st>>>(L, R, B, H, W, lr_cs, lr_co, out, out_cs, out_co);
        } else {
            constexpr int ST = 2;
            size_t smem = (size_t)ST * (PSM_TW + PSM_RW) * 128 * 4 + 2 * ST * 8;
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_v4_kernel<128, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_v4_kernel<128, ST>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
            psm_cosine_nhwc_v4_kernel<128, ST><<<grid, PSM4_THREADS, smem, st>>>(L, R, B, H, W, lr_cs, lr_co, out, out_cs, out_co);
        }
        VD3D_CHECK_LAUNCH("psm_cosine_nhwc_v4");
    } else if (fast && variant >= 3) {
        dim3 grid(cdiv(W, PSM_TW), H, B);
        size_t smem = (size_t)C * (PSM_LP + PSM_RP) * sizeof(float);
        if (C == 64) {
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_v3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_v3_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            psm_cosine_nhwc_v3_kernel<64><<<grid, PSM3_THREADS, smem, st>>>(L, R, H, W, lr_cs, lr_co, out, out_cs, out_co);
        } else {
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_v3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_v3_kernel<128>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            psm_cosine_nhwc_v3_kernel<128><<<grid, PSM3_THREADS, smem, st>>>(L, R, H, W, lr_cs, lr_co, out, out_cs, out_co);
        }
        VD3D_CHECK_LAUNCH("psm_cosine_nhwc_v3");
    } else if (fast) {
        dim3 grid(cdiv(W, PSM_TW), H, B);
        size_t smem = (size_t)C * (PSM_LP + PSM_RP) * sizeof(float);
        if (C == 64) {
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            psm_cosine_nhwc_kernel<64><<<grid, PSM_THREADS, smem, st>>>(L, R, H, W, lr_cs, lr_co, out, out_cs, out_co);
        } else {
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            VD3D_CUDA(cudaFuncSetAttribute(psm_cosine_nhwc_kernel<128>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            psm_cosine_nhwc_kernel<128><<<grid, PSM_THREADS, smem, st>>>(L, R, H, W, lr_cs, lr_co, out, out_cs, out_co);
        }
        VD3D_CHECK_LAUNCH("psm_cosine_nhwc");
    } else {
        long long npix = (long long)B * H * W;
        psm_cosine_nhwc_generic_kernel<<<cdiv(npix * D, 256), 256, 0, st>>>(L, R, npix, W, C, lr_cs, lr_co, D, out, out_cs, out_co);
        VD3D_CHECK_LAUNCH("psm_cosine_nhwc_generic");
    }
    return VD3D_OK;
}

extern "C" int vd3d_psm_cosine_nchw(const float* L, const float* R, int B, int C, int H, int W, int D, float* out, void* stream) {
    VD3D_REQUIRE(L && R && out && B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "psm_cosine_nchw: bad args");
    long long n = (long long)B * H * W;
    psm_cosine_nchw_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(L, R, B, C, H, W, D, out);
    VD3D_CHECK_LAUNCH("psm_cosine_nchw");
    return VD3D_OK;
}

extern "C" int vd3d_concat_volume_conv3d(const float* lf, const float* rf, int B, int H, int W, int F, int D,
                                         const float* w1, const float* b1, const float* w2, const float* b2,
                                         float* mid, float* out, int out_cs, int out_co, void* stream) {
    VD3D_REQUIRE(lf && rf && w1 && b1 && w2 && b2 && mid && out, "concat_volume_conv3d: null pointer");
    VD3D_REQUIRE(F == CV_F, "concat_volume_conv3d: PSM_features must be 8 (got %d)", F);
    VD3D_REQUIRE(B > 0 && H > 0 && W > 0 && D > 0, "concat_volume_conv3d: bad shape");
    long long total = (long long)B * D * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    concat_conv3d_1_kernel<<<cdiv(total, 128), 128, 0, st>>>(lf, rf, w1, b1, mid, B, D, H, W);
    VD3D_CHECK_LAUNCH("concat_conv3d_1");
    conv3d_2_kernel<<<cdiv(total, 128), 128, 0, st>>>(mid, w2, b2, out, B, D, H, W, out_cs, out_co);
    VD3D_CHECK_LAUNCH("conv3d_2");
    return VD3D_OK;
}
