// HBM-bound elementwise / pooling / depthwise kernels on NHWC activations + the library-wide error state.
// All are float4-vectorised along C (C % 4 == 0 except the layout converters) and grid-stride free:
// one thread per float4 of output, 256-thread CTAs.
#include "common.cuh"
#include <atomic>
#include <mutex>

namespace vd3d {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

static std::mutex g_flag_mu;
static int* g_range_flag[64] = {nullptr};
int* fp16_range_flag() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_flag_mu);
    if (!g_range_flag[dev]) {
        int* p = nullptr;
        if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) return nullptr;
        if (cudaMemset(p, 0, sizeof(int)) != cudaSuccess) { cudaFree(p); return nullptr; }
        g_range_flag[dev] = p;
    }
    return g_range_flag[dev];
}

// ---------------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW, int out_cs, int out_co) {
    // block: 32 pixels x 8 lanes... simple tiled transpose through smem: tile 32 (pixels) x 32 (channels)
    __shared__ float tile[32][33];
    int b = blockIdx.z;
    int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        int c = c0 + j, p = p0 + tx;
        tile[j][tx] = (c < C && p < HW) ? in[((long long)b * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int p = p0 + j, c = c0 + tx;
        if (c < C && p < HW) out[((long long)b * HW + p) * out_cs + out_co + c] = tile[tx][j];
    }
}

// few-channel form (images, C <= 4): one thread per pixel, plane reads and pixel writes both coalesced, no smem transpose
__global__ void nchw_to_nhwc_small_kernel(const float* __restrict__ in, float* __restrict__ out, int C, long long HW, long long total,
                                          int out_cs, int out_co) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long long b = idx / HW, p = idx - b * HW;
    const float* ip = in + b * C * HW + p;
    float* op = out + idx * out_cs + out_co;
    for (int c = 0; c < C; ++c) op[c] = __ldg(ip + (long long)c * HW);
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW, int in_cs, int in_co) {
    __shared__ float tile[32][33];
    int b = blockIdx.z;
    int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    int tx = threadIdx.x, ty = threadIdx.y;
    for (int j = ty; j < 32; j += 8) {
        int p = p0 + j, c = c0 + tx;
        tile[j][tx] = (c < C && p < HW) ? in[((long long)b * HW + p) * in_cs + in_co + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int c = c0 + j, p = p0 + tx;
        if (c < C && p < HW) out[((long long)b * C + c) * HW + p] = tile[tx][j];
    }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void maxpool3x3s2_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C4,
                                    int in_cs, int in_co, int Ho, int Wo, int out_cs, int out_co) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * Wo * C4;
    if (idx >= total) return;
    int c4 = (int)(idx % C4); long long r = idx / C4;
    int wo = (int)(r % Wo); r /= Wo; int ho = (int)(r % Ho); int b = (int)(r / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        int hi = ho * 2 - 1 + kh;
        if (hi < 0 || hi >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            int wi = wo * 2 - 1 + kw;
            if (wi < 0 || wi >= W) continue;
            float4 v = ldg4(in + (((long long)b * H + hi) * W + wi) * in_cs + in_co + 4 * c4);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    *reinterpret_cast<float4*>(out + (((long long)b * Ho + ho) * Wo + wo) * out_cs + out_co + 4 * c4) = m;
}

__global__ void avgpool2_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C4,
                                int in_cs, int in_co, int out_cs, int out_co) {
    int Ho = H / 2, Wo = W / 2;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * Wo * C4;
    if (idx >= total) return;
    int c4 = (int)(idx % C4); long long r = idx / C4;
    int wo = (int)(r % Wo); r /= Wo; int ho = (int)(r % Ho); int b = (int)(r / Ho);
    const float* p = in + (((long long)b * H + 2 * ho) * W + 2 * wo) * in_cs + in_co + 4 * c4;
    float4 a = ldg4(p), b4 = ldg4(p + in_cs), c = ldg4(p + (long long)W * in_cs), d = ldg4(p + (long long)(W + 1) * in_cs);
    // torch avg_pool2d sums the window in row-major order then divides by the window size
    float4 o;
    o.x = (((a.x + b4.x) + c.x) + d.x) / 4.0f;
    o.y = (((a.y + b4.y) + c.y) + d.y) / 4.0f;
    o.z = (((a.z + b4.z) + c.z) + d.z) / 4.0f;
    o.w = (((a.w + b4.w) + c.w) + d.w) / 4.0f;
    *reinterpret_cast<float4*>(out + (((long long)b * Ho + ho) * Wo + wo) * out_cs + out_co + 4 * c4) = o;
}

__global__ void dwconv3x3_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ bias,
                                 float* __restrict__ out, int B, int H, int W, int C4, int in_cs, int in_co,
                                 int out_cs, int out_co, int relu) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * H * W * C4;
    if (idx >= total) return;
    int c4 = (int)(idx % C4); long long r = idx / C4;
    int w = (int)(r % W); r /= W; int h = (int)(r % H); int b = (int)(r / H);
    int C = C4 * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        int hi = h - 1 + kh;
        if (hi < 0 || hi >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            int wi = w - 1 + kw;
            if (wi < 0 || wi >= W) continue;
            float4 v = ldg4(in + (((long long)b * H + hi) * W + wi) * in_cs + in_co + 4 * c4);
            float4 k = ldg4(wgt + (kh * 3 + kw) * C + 4 * c4);
            acc.x = fmaf(v.x, k.x, acc.x); acc.y = fmaf(v.y, k.y, acc.y); acc.z = fmaf(v.z, k.z, acc.z); acc.w = fmaf(v.w, k.w, acc.w);
        }
    }
    if (bias) { float4 bb = ldg4(bias + 4 * c4); acc.x += bb.x; acc.y += bb.y; acc.z += bb.z; acc.w += bb.w; }
    if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    *reinterpret_cast<float4*>(out + (((long long)b * H + h) * W + w) * out_cs + out_co + 4 * c4) = acc;
}

__global__ void maxpool2x2s2_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C4,
                                    int in_cs, int in_co, int out_cs, int out_co) {
    int Ho = H / 2, Wo = W / 2;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * Wo * C4;
    if (idx >= total) return;
    int c4 = (int)(idx % C4); long long r = idx / C4;
    int wo = (int)(r % Wo); r /= Wo; int ho = (int)(r % Ho); int b = (int)(r / Ho);
    const float* p = in + (((long long)b * H + 2 * ho) * W + 2 * wo) * in_cs + in_co + 4 * c4;
    float4 a = ldg4(p), b4 = ldg4(p + in_cs), c = ldg4(p + (long long)W * in_cs), d = ldg4(p + (long long)(W + 1) * in_cs);
    float4 o;
    o.x = fmaxf(fmaxf(a.x, b4.x), fmaxf(c.x, d.x)); o.y = fmaxf(fmaxf(a.y, b4.y), fmaxf(c.y, d.y));
    o.z = fmaxf(fmaxf(a.z, b4.z), fmaxf(c.z, d.z)); o.w = fmaxf(fmaxf(a.w, b4.w), fmaxf(c.w, d.w));
    *reinterpret_cast<float4*>(out + (((long long)b * Ho + ho) * Wo + wo) * out_cs + out_co + 4 * c4) = o;
}

// depthwise ConvTranspose2d(C, C, 2f, stride f, padding f/2, groups C, bias False) [+ addend]: the DLA up-sampling node
// (R/backbones/dla_utils.py:62-85).  out[y][x] = sum over the (at most) 2x2 input pixels whose kernel footprint covers (y, x).
__global__ void dw_convtranspose_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ addend,
                                        float* __restrict__ out, int B, int H, int W, int C4, int f, int in_cs, int in_co,
                                        int add_cs, int add_co, int out_cs, int out_co) {
    const int Ho = H * f, Wo = W * f, K = 2 * f, pad = f / 2, C = C4 * 4;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * Ho * Wo * C4;
    if (idx >= total) return;
    int c4 = (int)(idx % C4); long long r = idx / C4;
    int x = (int)(r % Wo); r /= Wo; int y = (int)(r % Ho); int b = (int)(r / Ho);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // torch's transposed conv accumulates input pixels in increasing (iy, ix) order for a given output pixel
    int iy_hi = (y + pad) / f, ix_hi = (x + pad) / f;
    for (int iy = iy_hi - 1; iy <= iy_hi; ++iy) {
        int ky = y + pad - iy * f;
        if (iy < 0 || iy >= H || ky < 0 || ky >= K) continue;
        for (int ix = ix_hi - 1; ix <= ix_hi; ++ix) {
            int kx = x + pad - ix * f;
            if (ix < 0 || ix >= W || kx < 0 || kx >= K) continue;
            float4 v = ldg4(in + (((long long)b * H + iy) * W + ix) * in_cs + in_co + 4 * c4);
            float4 k = ldg4(wgt + (ky * K + kx) * C + 4 * c4);
            acc.x = fmaf(v.x, k.x, acc.x); acc.y = fmaf(v.y, k.y, acc.y); acc.z = fmaf(v.z, k.z, acc.z); acc.w = fmaf(v.w, k.w, acc.w);
        }
    }
    long long opix = ((long long)b * Ho + y) * Wo + x;
    if (addend) {
        float4 a = ldg4(addend + opix * add_cs + add_co + 4 * c4);
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    *reinterpret_cast<float4*>(out + opix * out_cs + out_co + 4 * c4) = acc;
}

__global__ void copy_channels_kernel(const float* __restrict__ in, float* __restrict__ out, long long npix, int C4,
                                     int in_cs, int in_co, int out_cs, int out_co) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * C4) return;
    int c4 = (int)(idx % C4); long long pix = idx / C4;
    *reinterpret_cast<float4*>(out + pix * out_cs + out_co + 4 * c4) = ldg4(in + pix * in_cs + in_co + 4 * c4);
}

}  // namespace vd3d

using namespace vd3d;

extern "C" int vd3d_fp16_range_check(int* overflow_out, int reset, void* stream) {
    VD3D_REQUIRE(overflow_out, "fp16_range_check: null output");
    int* f = fp16_range_flag();
    VD3D_REQUIRE(f, "fp16_range_check: could not allocate the device flag");
    cudaStream_t st = (cudaStream_t)stream;
    int v = 0;
    VD3D_CUDA(cudaMemcpyAsync(&v, f, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (reset) VD3D_CUDA(cudaMemsetAsync(f, 0, sizeof(int), st));
    VD3D_CUDA(cudaStreamSynchronize(st));
    *overflow_out = v;
    return VD3D_OK;
}

extern "C" const char* vd3d_last_error(void) { return g_err; }
extern "C" int vd3d_version(void) { return 100; }
extern "C" long long vd3d_launch_count(void) { return g_launches.load(); }
extern "C" void vd3d_launch_count_reset(void) { g_launches = 0; }

extern "C" int vd3d_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, int out_cs, int out_co, void* stream) {
    VD3D_REQUIRE(in && out && B > 0 && C > 0 && H > 0 && W > 0 && out_cs >= out_co + C, "nchw_to_nhwc: bad args");
    if (C <= 4) {
        const long long total = (long long)B * H * W;
        nchw_to_nhwc_small_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, C, (long long)H * W, total, out_cs, out_co);
        VD3D_CHECK_LAUNCH("nchw_to_nhwc");
        return VD3D_OK;
    }
    dim3 grid(cdiv((long long)H * W, 32), cdiv(C, 32), B), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, out, C, H * W, out_cs, out_co);
    VD3D_CHECK_LAUNCH("nchw_to_nhwc");
    return VD3D_OK;
}

extern "C" int vd3d_nhwc_to_nchw(const float* in, float* out, int B, int C, int H, int W, int in_cs, int in_co, void* stream) {
    VD3D_REQUIRE(in && out && B > 0 && C > 0 && H > 0 && W > 0 && in_cs >= in_co + C, "nhwc_to_nchw: bad args");
    dim3 grid(cdiv((long long)H * W, 32), cdiv(C, 32), B), block(32, 8);
    nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, out, C, H * W, in_cs, in_co);
    VD3D_CHECK_LAUNCH("nhwc_to_nchw");
    return VD3D_OK;
}

#define VD3D_REQ_VEC4(name)                                                                                          \
    VD3D_REQUIRE(in && out && C > 0 && C % 4 == 0 && in_cs % 4 == 0 && in_co % 4 == 0 && out_cs % 4 == 0 && out_co % 4 == 0, \
                 name ": C, pitches and offsets must be multiples of 4")

extern "C" int vd3d_maxpool3x3s2_nhwc(const float* in, int B, int H, int W, int C, int in_cs, int in_co,
                                      float* out, int out_cs, int out_co, void* stream) {
    VD3D_REQ_VEC4("maxpool3x3s2");
    int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    long long total = (long long)B * Ho * Wo * (C / 4);
    maxpool3x3s2_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C / 4, in_cs, in_co, Ho, Wo, out_cs, out_co);
    VD3D_CHECK_LAUNCH("maxpool3x3s2");
    return VD3D_OK;
}

extern "C" int vd3d_avgpool2_nhwc(const float* in, int B, int H, int W, int C, int in_cs, int in_co,
                                  float* out, int out_cs, int out_co, void* stream) {
    VD3D_REQ_VEC4("avgpool2");
    VD3D_REQUIRE(H % 2 == 0 && W % 2 == 0, "avgpool2: H, W must be even");
    long long total = (long long)B * (H / 2) * (W / 2) * (C / 4);
    avgpool2_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C / 4, in_cs, in_co, out_cs, out_co);
    VD3D_CHECK_LAUNCH("avgpool2");
    return VD3D_OK;
}

extern "C" int vd3d_maxpool2x2s2_nhwc(const float* in, int B, int H, int W, int C, int in_cs, int in_co,
                                      float* out, int out_cs, int out_co, void* stream) {
    VD3D_REQ_VEC4("maxpool2x2s2");
    VD3D_REQUIRE(H >= 2 && W >= 2, "maxpool2x2s2: input too small");
    long long total = (long long)B * (H / 2) * (W / 2) * (C / 4);
    maxpool2x2s2_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C / 4, in_cs, in_co, out_cs, out_co);
    VD3D_CHECK_LAUNCH("maxpool2x2s2");
    return VD3D_OK;
}

extern "C" int vd3d_dw_convtranspose_nhwc(const float* in, int B, int H, int W, int C, int in_cs, int in_co, const float* wgt, int f,
                                          const float* addend, int add_cs, int add_co, float* out, int out_cs, int out_co, void* stream) {
    VD3D_REQ_VEC4("dw_convtranspose");
    VD3D_REQUIRE(wgt && f >= 2 && f % 2 == 0, "dw_convtranspose: up-sampling factor must be even (kernel 2f, stride f, padding f/2)");
    VD3D_REQUIRE(!addend || (add_cs % 4 == 0 && add_co % 4 == 0), "dw_convtranspose: addend pitch/offset must be multiples of 4");
    long long total = (long long)B * H * f * W * f * (C / 4);
    dw_convtranspose_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, wgt, addend, out, B, H, W, C / 4, f, in_cs, in_co,
                                                                             add_cs, add_co, out_cs, out_co);
    VD3D_CHECK_LAUNCH("dw_convtranspose");
    return VD3D_OK;
}

extern "C" int vd3d_dwconv3x3_nhwc(const float* in, int B, int H, int W, int C, int in_cs, int in_co,
                                   const float* wgt, const float* bias, float* out, int out_cs, int out_co, int relu, void* stream) {
    VD3D_REQ_VEC4("dwconv3x3");
    VD3D_REQUIRE(wgt, "dwconv3x3: null weights");
    long long total = (long long)B * H * W * (C / 4);
    dwconv3x3_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, wgt, bias, out, B, H, W, C / 4, in_cs, in_co, out_cs, out_co, relu);
    VD3D_CHECK_LAUNCH("dwconv3x3");
    return VD3D_OK;
}

extern "C" int vd3d_copy_channels_nhwc(const float* in, int npix, int C, int in_cs, int in_co, float* out, int out_cs, int out_co, void* stream) {
    VD3D_REQ_VEC4("copy_channels");
    long long total = (long long)npix * (C / 4);
    copy_channels_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, npix, C / 4, in_cs, in_co, out_cs, out_co);
    VD3D_CHECK_LAUNCH("copy_channels");
    return VD3D_OK;
}
