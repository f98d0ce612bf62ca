// pel.cu — libavcodec h264qpel (8 bit) and hpeldsp motion-compensation interpolation on sm_100a (C ABI: "h264qpel / hpeldsp").
//
// Reference semantics reproduced bit-for-bit (checker: oracle/pel_oracle.c):
//   H264_LOWPASS / H264_MC / op_put, op_avg   libavcodec/h264qpel_template.c:77-465
//   PIXOP2, rnd_avg32 / no_rnd_avg32          libavcodec/hpeldsp.c:38-333, libavcodec/rnd_avg.h:31-39
//
// Batched kernel: one warp per operation.  The warp first stages the (size+5) x (size+5) source window in shared
// memory (coalesced row reads through the read-only path), for positions that need the centre sample it builds the
// unrounded horizontal 6-tap sums once (size+5 rows of int16), and then every lane produces its pixels from shared
// memory.  The 6-tap is never recomputed per output pixel and the reference block is read from HBM exactly once.
#include "common.h"
#include <cstring>

namespace {

__device__ __forceinline__ int clip8(int v) { return __vimin_s32_relu(v, 255); }
__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

constexpr int QW = 24;                 // window pitch (bytes): 16 + 5 -> 21, padded
constexpr int WARPS = 4;

struct QpelSmem {
    uint8_t win[21 * QW];              // source rows -2 .. size+2, columns -2 .. size+2
    short hraw[21 * 16];               // unrounded horizontal sums for rows -2 .. size+2
};

__global__ void __launch_bounds__(32 * WARPS)
qpel_kernel(long long n, const uint8_t *op, uint8_t *dst, const int64_t *dst_off, const uint8_t *src,
            const int64_t *src_off, long long stride)
{
    __shared__ QpelSmem sm[WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * WARPS + warp;
    if (i >= n) return;
    QpelSmem &s = sm[warp];
    const int o = op[i];
    const int avg = o & 1, size = 16 >> ((o >> 1) & 3), qx = (o >> 3) & 3, qy = (o >> 5) & 3;
    const uint8_t *sp = src + src_off[i];
    uint8_t *dp = dst + dst_off[i];
    const int wdim = size + 5;
    for (int k = lane; k < wdim * wdim; k += 32) {
        const int r = k / wdim, c = k - r * wdim;
        s.win[r * QW + c] = __ldg(sp + (long long)(r - 2) * stride + (c - 2));
    }
    __syncwarp();
    const bool need_j = (qx == 2 && qy != 0) || (qy == 2 && qx != 0);      // positions built from the centre sample
    if (need_j) {
        for (int k = lane; k < wdim * size; k += 32) {
            const int r = k / size, c = k - r * size;
            const uint8_t *p = &s.win[r * QW + c];                             // taps at columns c-2 .. c+3 of the block
            s.hraw[r * 16 + c] = (short)tap6(p[0], p[1], p[2], p[3], p[4], p[5]);
        }
        __syncwarp();
    }
    auto F = [&](int x, int y) { return (int)s.win[(y + 2) * QW + x + 2]; };
    auto H = [&](int x, int y) { const uint8_t *p = &s.win[(y + 2) * QW + x]; return clip8((tap6(p[0], p[1], p[2], p[3], p[4], p[5]) + 16) >> 5); };
    auto V = [&](int x, int y) { const uint8_t *p = &s.win[y * QW + x + 2];
        return clip8((tap6(p[0], p[QW], p[2 * QW], p[3 * QW], p[4 * QW], p[5 * QW]) + 16) >> 5); };
    auto J = [&](int x, int y) { const short *p = &s.hraw[y * 16 + x];
        return clip8((tap6(p[0], p[16], p[32], p[48], p[64], p[80]) + 512) >> 10); };
    for (int k = lane; k < size * size; k += 32) {
        const int y = k / size, x = k - y * size;
        int v;
        if (qy == 0)            v = qx == 0 ? F(x, y) : qx == 2 ? H(x, y) : (F(x + (qx == 3), y) + H(x, y) + 1) >> 1;
        else if (qx == 0)       v = qy == 2 ? V(x, y) : (F(x, y + (qy == 3)) + V(x, y) + 1) >> 1;
        else if (qx == 2 && qy == 2) v = J(x, y);
        else if (qx == 2)       v = (H(x, y + (qy == 3)) + J(x, y) + 1) >> 1;
        else if (qy == 2)       v = (V(x + (qx == 3), y) + J(x, y) + 1) >> 1;
        else                    v = (H(x, y + (qy == 3)) + V(x + (qx == 3), y) + 1) >> 1;
        uint8_t *d = dp + (long long)y * stride + x;
        *d = (uint8_t)(avg ? (*d + v + 1) >> 1 : v);
    }
}

// hpel: one warp per operation, direct global reads (at most 4 taps per pixel, rows are contiguous)
__global__ void __launch_bounds__(32 * WARPS)
hpel_kernel(long long n, const uint8_t *op, const uint8_t *hh, uint8_t *dst, const int64_t *dst_off, const uint8_t *src,
            const int64_t *src_off, long long stride)
{
    const long long i = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const int o = op[i], h = hh[i];
    const int tab = o & 3, sidx = (o >> 2) & 3, xy = (o >> 4) & 3, w = 16 >> sidx;
    const int no_rnd = tab >= 2;
    const int avg = (tab & 1) && !(sidx == 3 && xy == 3);        // avg_pixels2_xy2 stores without averaging (hpeldsp.c:134-166)
    const uint8_t *sp = src + src_off[i];
    uint8_t *dp = dst + dst_off[i];
    for (int k = lane; k < w * h; k += 32) {
        const int y = k / w, x = k - y * w;
        const uint8_t *p = sp + (long long)y * stride + x;
        int v;
        switch (xy) {
        case 0:  v = __ldg(p); break;
        case 1:  v = (__ldg(p) + __ldg(p + 1) + 1 - no_rnd) >> 1; break;
        case 2:  v = (__ldg(p) + __ldg(p + stride) + 1 - no_rnd) >> 1; break;
        default: v = (__ldg(p) + __ldg(p + 1) + __ldg(p + stride) + __ldg(p + stride + 1) + 2 - no_rnd) >> 2; break;
        }
        uint8_t *d = dp + (long long)y * stride + x;
        *d = (uint8_t)(avg ? (*d + v + 1) >> 1 : v);
    }
}

void die(const char *what)
{
    fprintf(stderr, "libb200dsp: motion compensation failed: %s (%s)\n", what, b200_last_error());
    abort();
}

// one block through the device for the drop-in tables; `pad` = rows/cols read around the block
void host_op(bool qpel, int o, int h, int w, uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    B200Device *dev = b200_default_device();
    if (!dev) die("no device");
    if (stride < 0) die("negative stride");
    if (cudaSetDevice(dev->ordinal) != cudaSuccess) die("cudaSetDevice");
    const int before = qpel ? 2 : 0, after = qpel ? 3 : 1;
    const int sw = w + before + after, sh = h + before + after;
    const size_t pitch = 32;
    uint8_t *scr = (uint8_t *)b200_scratch(dev, pitch * (sh + h) + 256);
    if (!scr) die("scratch");
    uint8_t *dsrc = scr, *ddst = scr + pitch * sh;
    uint8_t *meta = scr + pitch * (sh + h);           // op, h, offsets
    cudaStream_t st = dev->stream;
    if (cudaMemcpy2DAsync(dsrc, pitch, src - before * stride - before, (size_t)stride, sw, sh, cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d src");
    if (cudaMemcpy2DAsync(ddst, pitch, dst, (size_t)stride, w, h, cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d dst");
    struct { int64_t doff, soff; uint8_t op, h; } m = { 0, (int64_t)(before * pitch + before), (uint8_t)o, (uint8_t)h };
    if (cudaMemcpyAsync(meta, &m, sizeof(m), cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d meta");
    const int64_t *doff = (const int64_t *)meta, *soff = doff + 1;
    const uint8_t *dop = meta + 16, *dh = meta + 17;
    // dst and src live in one buffer with the same pitch, like the reference's single stride
    if (qpel) qpel_kernel<<<1, 32 * WARPS, 0, st>>>(1, dop, ddst, doff, dsrc, soff, (long long)pitch);
    else      hpel_kernel<<<1, 32 * WARPS, 0, st>>>(1, dop, dh, ddst, doff, dsrc, soff, (long long)pitch);
    B200_LAUNCHED();
    if (cudaMemcpy2DAsync(dst, (size_t)stride, ddst, pitch, w, h, cudaMemcpyDeviceToHost, st) != cudaSuccess) die("d2h");
    if (cudaStreamSynchronize(st) != cudaSuccess) die("sync");
}

template <int AVG, int SIDX, int POS>
void qpel_tab(uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    host_op(true, AVG | (SIDX << 1) | (POS << 3), 16 >> SIDX, 16 >> SIDX, dst, src, stride);
}
template <int TAB, int SIDX, int XY>
void hpel_tab(uint8_t *block, const uint8_t *pixels, ptrdiff_t line_size, int h)
{
    host_op(false, TAB | (SIDX << 2) | (XY << 4), h, 16 >> SIDX, block, pixels, line_size);
}

template <int AVG, int SIDX>
void fill_qpel(b200_qpel_mc_func *t)
{
    t[0] = qpel_tab<AVG, SIDX, 0>;   t[1] = qpel_tab<AVG, SIDX, 1>;   t[2] = qpel_tab<AVG, SIDX, 2>;   t[3] = qpel_tab<AVG, SIDX, 3>;
    t[4] = qpel_tab<AVG, SIDX, 4>;   t[5] = qpel_tab<AVG, SIDX, 5>;   t[6] = qpel_tab<AVG, SIDX, 6>;   t[7] = qpel_tab<AVG, SIDX, 7>;
    t[8] = qpel_tab<AVG, SIDX, 8>;   t[9] = qpel_tab<AVG, SIDX, 9>;   t[10] = qpel_tab<AVG, SIDX, 10>; t[11] = qpel_tab<AVG, SIDX, 11>;
    t[12] = qpel_tab<AVG, SIDX, 12>; t[13] = qpel_tab<AVG, SIDX, 13>; t[14] = qpel_tab<AVG, SIDX, 14>; t[15] = qpel_tab<AVG, SIDX, 15>;
}
template <int TAB, int SIDX>
void fill_hpel(b200_op_pixels_func *t)
{
    t[0] = hpel_tab<TAB, SIDX, 0>; t[1] = hpel_tab<TAB, SIDX, 1>; t[2] = hpel_tab<TAB, SIDX, 2>; t[3] = hpel_tab<TAB, SIDX, 3>;
}

} // namespace

B200_API int b200_h264qpel_init(B200H264QpelContext *c, int bit_depth)
{
    if (!c) return B200_EINVAL;
    if (bit_depth != 8) return B200_ENOSYS;                        // h264qpel.c:87-103 also installs 9/10/12/14 bit tables
    if (!b200_default_device()) return B200_ENODEV;
    fill_qpel<0, 0>(c->put_h264_qpel_pixels_tab[0]); fill_qpel<0, 1>(c->put_h264_qpel_pixels_tab[1]); fill_qpel<0, 2>(c->put_h264_qpel_pixels_tab[2]);
    fill_qpel<1, 0>(c->avg_h264_qpel_pixels_tab[0]); fill_qpel<1, 1>(c->avg_h264_qpel_pixels_tab[1]); fill_qpel<1, 2>(c->avg_h264_qpel_pixels_tab[2]);
    return 0;
}

B200_API int b200_hpeldsp_init(B200HpelDSPContext *c, int flags)
{
    (void)flags;
    if (!c) return B200_EINVAL;
    if (!b200_default_device()) return B200_ENODEV;
    memset(c, 0, sizeof(*c));
    fill_hpel<0, 0>(c->put_pixels_tab[0]); fill_hpel<0, 1>(c->put_pixels_tab[1]); fill_hpel<0, 2>(c->put_pixels_tab[2]); fill_hpel<0, 3>(c->put_pixels_tab[3]);
    fill_hpel<1, 0>(c->avg_pixels_tab[0]); fill_hpel<1, 1>(c->avg_pixels_tab[1]); fill_hpel<1, 2>(c->avg_pixels_tab[2]); fill_hpel<1, 3>(c->avg_pixels_tab[3]);
    fill_hpel<2, 0>(c->put_no_rnd_pixels_tab[0]); fill_hpel<2, 1>(c->put_no_rnd_pixels_tab[1]);
    fill_hpel<3, 0>(c->avg_no_rnd_pixels_tab);
    return 0;
}

B200_API int b200_h264qpel_batch_device(B200Device *dev, int64_t n, const uint8_t *op, uint8_t *dst, const int64_t *dst_off,
                                        const uint8_t *src, const int64_t *src_off, ptrdiff_t stride)
{
    if (!dev || n < 0 || !op || !dst || !dst_off || !src || !src_off) return B200_EINVAL;
    if (n == 0) return 0;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    const long long blocks = (n + WARPS - 1) / WARPS;
    if (blocks > 0x7fffffffLL) return B200_EINVAL;
    qpel_kernel<<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, op, dst, dst_off, src, src_off, stride);
    B200_LAUNCHED();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

B200_API int b200_hpel_batch_device(B200Device *dev, int64_t n, const uint8_t *op, const uint8_t *h, uint8_t *dst,
                                    const int64_t *dst_off, const uint8_t *src, const int64_t *src_off, ptrdiff_t stride)
{
    if (!dev || n < 0 || !op || !h || !dst || !dst_off || !src || !src_off) return B200_EINVAL;
    if (n == 0) return 0;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    const long long blocks = (n + WARPS - 1) / WARPS;
    if (blocks > 0x7fffffffLL) return B200_EINVAL;
    hpel_kernel<<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, op, h, dst, dst_off, src, src_off, stride);
    B200_LAUNCHED();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}
