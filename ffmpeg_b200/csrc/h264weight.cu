// h264weight.cu — libavcodec H.264 explicit weighted prediction (8 bit) on sm_100a (C ABI: "h264 weighted prediction").
//
// Reference semantics reproduced bit-for-bit (checker: oracle/pel_oracle.c):
//   weight_h264_pixels{16,8,4,2}_8_c     libavcodec/h264dsp_template.c:30-62    block = clip((block*w + o') >> d)
//   biweight_h264_pixels{16,8,4,2}_8_c   libavcodec/h264dsp_template.c:63-93    dst = clip((src*ws + dst*wd + o'') >> (d+1))
// with o' = (o << d) + (d ? 1 << (d-1) : 0) and o'' = ((o + 1) | 1) << d, all in 32-bit integers.
//
// Batched kernel: one warp per block, a lane owns 4 consecutive pixels of a row (32-bit accesses when the row is 4-aligned).
#include "common.h"
#include "pel_hbd.h"
#include <cstring>

namespace {

constexpr int WARPS = 4;
__device__ __forceinline__ int clip8(int v) { return __vimin_s32_relu(v, 255); }

// params per block: p[0] = idx (width 16 >> idx) | height << 8 | log2_denom << 16, p[1] = weight (dst weight for biweight),
//                   p[2] = source weight (biweight only), p[3] = offset
template <bool BI>
__global__ void __launch_bounds__(32 * WARPS)
h264_weight_kernel(long long n, const int32_t *params, uint8_t *dst, const int64_t *dst_off, const uint8_t *src, const int64_t *src_off,
                   long long stride)
{
    const long long i = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const int4 p = __ldg(reinterpret_cast<const int4 *>(params) + i);
    const int w = 16 >> (p.x & 3), h = (p.x >> 8) & 255, d = (p.x >> 16) & 31;
    const int wd = p.y, ws = p.z;
    int off;
    if (BI) off = (int)((unsigned)((p.w + 1) | 1) << d);
    else { off = (int)((unsigned)p.w << d); if (d) off += 1 << (d - 1); }
    const int sh = BI ? d + 1 : d;
    uint8_t *dp = dst + __ldg(dst_off + i);
    const uint8_t *sp = BI ? src + __ldg(src_off + i) : nullptr;
    const int segs = w > 4 ? w >> 2 : 1, npx = w < 4 ? w : 4;          // 4-pixel segments per row
    for (int k = lane; k < h * segs; k += 32) {
        const int y = k / segs, x0 = (k - y * segs) * 4;
        uint8_t *q = dp + (long long)y * stride + x0;
        const uint8_t *r = BI ? sp + (long long)y * stride + x0 : nullptr;
        if (npx == 4 && ((reinterpret_cast<uintptr_t>(q) & 3) == 0) && (!BI || (reinterpret_cast<uintptr_t>(r) & 3) == 0)) {
            const unsigned a = *reinterpret_cast<const unsigned *>(q);
            const unsigned b = BI ? __ldg(reinterpret_cast<const unsigned *>(r)) : 0u;
            unsigned o = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int pa = (int)((a >> (8 * j)) & 255), pb = (int)((b >> (8 * j)) & 255);
                const int v = BI ? (pb * ws + pa * wd + off) >> sh : (pa * wd + off) >> sh;
                o |= (unsigned)clip8(v) << (8 * j);
            }
            *reinterpret_cast<unsigned *>(q) = o;
        } else {
            for (int j = 0; j < npx; j++) {
                const int pa = q[j], pb = BI ? (int)__ldg(r + j) : 0;
                q[j] = (uint8_t)clip8(BI ? (pb * ws + pa * wd + off) >> sh : (pa * wd + off) >> sh);
            }
        }
    }
}

void die(const char *what)
{
    fprintf(stderr, "libb200dsp: h264 weighted prediction failed: %s (%s)\n", what, b200_last_error());
    abort();
}

// drop-in: one block through the device (host pointers)
void host_op(bool bi, int idx, uint8_t *dst, const uint8_t *src, ptrdiff_t stride, int height, int log2_denom, int wd, int ws, int offset)
{
    B200Device *dev = b200_default_device();
    if (!dev) die("no device");
    if (height < 0 || height > 255 || log2_denom < 0 || log2_denom > 7) die("unsupported height / log2_denom");
    if (height == 0) return;
    if (cudaSetDevice(dev->ordinal) != cudaSuccess) die("cudaSetDevice");
    const int w = 16 >> idx;
    const size_t pitch = 16;
    B200_LOCK_DEVICE(dev);      // scratch + stream are per device: one host-pointer call at a time (released on return)
    uint8_t *scr = (uint8_t *)b200_scratch(dev, 2 * pitch * 256 + 256);
    if (!scr) die("scratch");
    uint8_t *ddst = scr, *dsrc = scr + pitch * 256, *meta = scr + 2 * pitch * 256;
    cudaStream_t st = dev->stream;
    if (b200_h2d_rows(ddst, pitch, dst, stride, w, height, st) != cudaSuccess) die("h2d dst");
    if (bi && b200_h2d_rows(dsrc, pitch, src, stride, w, height, st) != cudaSuccess) die("h2d src");
    struct { int32_t p[4]; int64_t doff, soff; } m = { { idx | (height << 8) | (log2_denom << 16), wd, ws, offset }, 0, 0 };
    if (cudaMemcpyAsync(meta, &m, sizeof(m), cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d meta");
    const int32_t *dp = (const int32_t *)meta;
    const int64_t *doff = (const int64_t *)(meta + 16), *soff = doff + 1;
    if (bi) h264_weight_kernel<true><<<1, 32 * WARPS, 0, st>>>(1, dp, ddst, doff, dsrc, soff, (long long)pitch);
    else    h264_weight_kernel<false><<<1, 32 * WARPS, 0, st>>>(1, dp, ddst, doff, nullptr, nullptr, (long long)pitch);
    B200_LAUNCHED();
    if (b200_d2h_rows(dst, stride, ddst, pitch, w, height, st) != cudaSuccess) die("d2h");
    if (cudaStreamSynchronize(st) != cudaSuccess) die("sync");
}

template <int IDX>
void weight_tab(uint8_t *block, ptrdiff_t stride, int height, int log2_denom, int weight, int offset)
{
    host_op(false, IDX, block, nullptr, stride, height, log2_denom, weight, 0, offset);
}
template <int IDX>
void biweight_tab(uint8_t *dst, uint8_t *src, ptrdiff_t stride, int height, int log2_denom, int weightd, int weights, int offset)
{
    host_op(true, IDX, dst, src, stride, height, log2_denom, weightd, weights, offset);
}

} // namespace

B200_API int b200_h264_weight_init(B200H264WeightContext *c, int bit_depth)
{
    if (!c) return B200_EINVAL;
    if (!b200_default_device()) return B200_ENODEV;
    if (bit_depth != 8) return pel_hbd_fill_weight(c, bit_depth) ? 0 : B200_ENOSYS;      // 9 / 10 / 12 / 14: uint16 samples (pel_hbd.cu)
    c->weight_pixels_tab[0] = weight_tab<0>; c->weight_pixels_tab[1] = weight_tab<1>;
    c->weight_pixels_tab[2] = weight_tab<2>; c->weight_pixels_tab[3] = weight_tab<3>;
    c->biweight_pixels_tab[0] = biweight_tab<0>; c->biweight_pixels_tab[1] = biweight_tab<1>;
    c->biweight_pixels_tab[2] = biweight_tab<2>; c->biweight_pixels_tab[3] = biweight_tab<3>;
    return 0;
}

B200_API int b200_h264_weight_batch_device(B200Device *dev, int64_t n, const int32_t *params, uint8_t *dst, const int64_t *dst_off,
                                           const uint8_t *src, const int64_t *src_off, ptrdiff_t stride)
{
    if (!dev || n < 0 || !params || !dst || !dst_off || (src && !src_off)) return B200_EINVAL;
    if ((uintptr_t)params & 15) return B200_EINVAL;
    if (n == 0) return 0;
    const long long blocks = (n + WARPS - 1) / WARPS;
    if (blocks > 0x7fffffffLL) return B200_EINVAL;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    if (src) h264_weight_kernel<true><<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, params, dst, dst_off, src, src_off, stride);
    else     h264_weight_kernel<false><<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, params, dst, dst_off, nullptr, nullptr, stride);
    B200_LAUNCHED();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}
