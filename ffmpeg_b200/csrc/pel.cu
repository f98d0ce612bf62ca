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

constexpr int QW = 32;                 // window pitch in bytes; block pixel (0,0) sits at column 8 of window row 2
constexpr int QX = 8;                  // so that every 8-pixel segment of a row starts 8-byte aligned in shared memory
constexpr int WARPS = 4;

struct __align__(16) QpelSmem {
    uint8_t win[21 * QW];              // source rows -2 .. size+2; source column x is at byte QX + x (x = -2 .. size+2)
    short hraw[21 * 16];               // unrounded horizontal 6-tap sums for the same rows, block columns 0 .. 15
};

// 16 consecutive bytes of a window row starting at 4-aligned byte offset `o`, unpacked
__device__ __forceinline__ void row16(const uint8_t *row, int o, int *b)
{
    const unsigned *w = reinterpret_cast<const unsigned *>(row + o);
    const unsigned w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        b[i] = (int)__byte_perm(w0, 0, 0x4440 | i); b[4 + i] = (int)__byte_perm(w1, 0, 0x4440 | i);
        b[8 + i] = (int)__byte_perm(w2, 0, 0x4440 | i); b[12 + i] = (int)__byte_perm(w3, 0, 0x4440 | i);
    }
}
// 8 bytes starting at byte offset QX + x0 + dx (dx = 0 or 1) of a window row
__device__ __forceinline__ void row8(const uint8_t *row, int x0, int dx, int *b)
{
    const unsigned *w = reinterpret_cast<const unsigned *>(row + QX + x0);
    unsigned w0 = w[0], w1 = w[1];
    if (dx) { const unsigned w2 = w[2]; w0 = __funnelshift_r(w0, w1, 8); w1 = __funnelshift_r(w1, w2, 8); }
#pragma unroll
    for (int i = 0; i < 4; i++) { b[i] = (int)__byte_perm(w0, 0, 0x4440 | i); b[4 + i] = (int)__byte_perm(w1, 0, 0x4440 | i); }
}
// unrounded horizontal 6-tap sums for block columns x0 .. x0+7 of window row `row`
__device__ __forceinline__ void hsum8(const uint8_t *row, int x0, int *o)
{
    int b[16];
    row16(row, QX + x0 - 4, b);                          // bytes of source columns x0-4 .. x0+11; taps need x0-2 .. x0+10
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = tap6(b[i + 2], b[i + 3], b[i + 4], b[i + 5], b[i + 6], b[i + 7]);
}

// The kernel is latency-bound if every operation waits for its descriptor, then its window, then (avg) its
// destination.  Each warp therefore walks QK consecutive operations and software-pipelines them: while operation t is
// computed, the window words of t+1 are already in flight into registers, the descriptor of t+2 is being fetched, and
// the destination row of t (avg) was requested before the shared-memory staging started.
constexpr int QK = 8;

struct QMeta { int o; long long soff, doff; };
__device__ __forceinline__ QMeta load_meta(const uint8_t *op, const int64_t *src_off, const int64_t *dst_off, long long i, long long n)
{
    QMeta m; m.o = -1; m.soff = 0; m.doff = 0;
    if (i < n) { m.o = __ldg(op + i); m.soff = __ldg(src_off + i); m.doff = __ldg(dst_off + i); }
    return m;
}
struct QWin { unsigned v[6]; unsigned sh[6]; };
// window words of one operation into registers: 7 lanes per source row, 4 rows per pass (see stage_window)
__device__ __forceinline__ void load_window(QWin &w, const uint8_t *sp, long long stride, int wdim, int lane, bool valid)
{
    const int q7 = lane / 7, j = lane - q7 * 7;
#pragma unroll
    for (int it = 0; it < 6; it++) {
        const int r = 4 * it + q7;
        const uint8_t *first = sp + (long long)(r - 2) * stride - 2;                     // column -2 of this row
        const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(first) & 3);
        const int nw = (int)((sh + wdim + 3) >> 2);
        w.sh[it] = sh;
        w.v[it] = 0;
        if (valid && lane < 28 && r < wdim && j >= 1 && j - 1 < nw) w.v[it] = __ldg(reinterpret_cast<const unsigned *>(first - sh) + (j - 1));
    }
}
// Lane j of a row holds LL[j]: LL[0] = 0, LL[1..6] = the aligned words that contain at least one needed byte (columns
// -2 .. size+2; nothing else is ever read), LL[7] = 0.  Window word jj (bytes 4jj .. 4jj+3 of the shared row, block
// column 0 at byte QX) = funnel(LL[jj-1+i0], LL[jj+i0]) with i0, shift from the row's address alignment.
__device__ __forceinline__ void stage_window(QpelSmem &s, const QWin &w, int wdim, int lane)
{
    const int q7 = lane / 7, j = lane - q7 * 7;
#pragma unroll
    for (int it = 0; it < 6; it++) {
        const int r = 4 * it + q7;
        const unsigned nxt = __shfl_down_sync(0xffffffffu, w.v[it], 1);
        const unsigned hi = j == 6 ? 0u : nxt;
        const unsigned tt = w.sh[it] + 2, i0 = tt >> 2, fs = (tt & 3) * 8;
        const int jj = j + 1 - (int)i0;
        if (lane < 28 && r < wdim && jj >= 1 && jj <= 6)
            reinterpret_cast<unsigned *>(&s.win[r * QW])[jj] = __funnelshift_r(w.v[it], hi, fs);
    }
}

// Lane l works on one 8-pixel (4 for size 4) row segment: row l>>1, half l&1 for 16x16.
__global__ void __launch_bounds__(32 * WARPS)
qpel_kernel(long long n, const uint8_t *op, uint8_t *dst, const int64_t *dst_off, const uint8_t *src,
            const int64_t *src_off, long long stride)
{
    __shared__ QpelSmem sm[WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long first_op = ((long long)blockIdx.x * WARPS + warp) * QK;
    if (first_op >= n) return;
    QpelSmem &s = sm[warp];
    QMeta m0 = load_meta(op, src_off, dst_off, first_op, n);
    QMeta m1 = load_meta(op, src_off, dst_off, first_op + 1, n);
    QWin w0, w1;
    load_window(w0, src + m0.soff, stride, (16 >> ((m0.o >> 1) & 3)) + 5, lane, m0.o >= 0);
    for (int t = 0; t < QK; t++) {
        if (m0.o < 0) break;                                                             // past the end (warp-uniform)
        const QMeta m2 = load_meta(op, src_off, dst_off, first_op + t + 2, (t + 2 < QK) ? n : 0);
        const bool next_ok = t + 1 < QK && m1.o >= 0;
        load_window(w1, src + m1.soff, stride, (16 >> ((m1.o >> 1) & 3)) + 5, lane, next_ok);

        const int o = m0.o;
        const int avg = o & 1, size = 16 >> ((o >> 1) & 3), qx = (o >> 3) & 3, qy = (o >> 5) & 3;
        uint8_t *dp = dst + m0.doff;
        const int wdim = size + 5;
        const int npx = size < 8 ? size : 8, segs = size >> 3 ? size >> 3 : 1;           // segments per row
        const bool mine = lane < size * segs;
        const int y = lane / segs, x0 = (lane - y * segs) * 8;
        uint8_t *d = dp + (long long)y * stride + x0;
        const bool vec = npx == 8 && ((reinterpret_cast<uintptr_t>(d)) & 7) == 0;
        uint2 pv = make_uint2(0, 0);
        if (mine && avg && vec) pv = *reinterpret_cast<const uint2 *>(d);                // destination row requested early

        stage_window(s, w0, wdim, lane);
        __syncwarp();
        const bool need_j = (qx == 2 && qy != 0) || (qy == 2 && qx != 0);                // positions built from the centre sample
        if (need_j) {
            for (int k = lane; k < wdim * segs; k += 32) {
                const int r = k / segs, xx = (k - r * segs) * 8;
                int h[8];
                hsum8(&s.win[r * QW], xx, h);
                uint4 pk;
                pk.x = (unsigned)(h[0] & 0xffff) | ((unsigned)h[1] << 16); pk.y = (unsigned)(h[2] & 0xffff) | ((unsigned)h[3] << 16);
                pk.z = (unsigned)(h[4] & 0xffff) | ((unsigned)h[5] << 16); pk.w = (unsigned)(h[6] & 0xffff) | ((unsigned)h[7] << 16);
                *reinterpret_cast<uint4 *>(&s.hraw[r * 16 + xx]) = pk;
            }
            __syncwarp();
        }
        if (mine) {
            // the one or two samples each quarter position averages (h264qpel_template.c:313-456): F full-pel, H horizontal
            // half, V vertical half, J centre.  acc collects them; two samples -> (a + b + 1) >> 1.
            const bool useF = (qy == 0 && qx != 2) || (qx == 0 && (qy & 1));
            const bool useH = qx != 0 && qy != 2;
            const bool useV = qy != 0 && qx != 2;
            int acc[8], tt[8], v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = 0;
            int count = 0;
            if (useF) {
                row8(&s.win[(y + 2 + (qy == 3)) * QW], x0, qx == 3, tt);
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] += tt[k];
                count++;
            }
            if (useH) {
                hsum8(&s.win[(y + 2 + (qy == 3)) * QW], x0, tt);
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] += clip8((tt[k] + 16) >> 5);
                count++;
            }
            if (useV) {
                int r0[8], r1[8], r2[8], r3[8], r4[8], r5[8];
                const int dx = qx == 3;
                row8(&s.win[(y + 0) * QW], x0, dx, r0); row8(&s.win[(y + 1) * QW], x0, dx, r1); row8(&s.win[(y + 2) * QW], x0, dx, r2);
                row8(&s.win[(y + 3) * QW], x0, dx, r3); row8(&s.win[(y + 4) * QW], x0, dx, r4); row8(&s.win[(y + 5) * QW], x0, dx, r5);
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] += clip8((tap6(r0[k], r1[k], r2[k], r3[k], r4[k], r5[k]) + 16) >> 5);
                count++;
            }
            if (need_j) {
                int j6[8];
#pragma unroll
                for (int k = 0; k < 8; k++) j6[k] = 512;
                const int coef[6] = { 1, -5, 20, 20, -5, 1 };
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    const uint4 pk = *reinterpret_cast<const uint4 *>(&s.hraw[(y + j) * 16 + x0]);
                    const unsigned ww[4] = { pk.x, pk.y, pk.z, pk.w };
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        j6[2 * k] += coef[j] * (int)(short)(ww[k] & 0xffff);
                        j6[2 * k + 1] += coef[j] * ((int)ww[k] >> 16);
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] += clip8(j6[k] >> 10);
                count++;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = count == 2 ? (acc[k] + 1) >> 1 : acc[k];
            if (vec) {
                if (avg) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        v[k] = (v[k] + (int)__byte_perm(pv.x, 0, 0x4440 | k) + 1) >> 1;
                        v[4 + k] = (v[4 + k] + (int)__byte_perm(pv.y, 0, 0x4440 | k) + 1) >> 1;
                    }
                }
                uint2 ov;
                ov.x = (unsigned)v[0] | ((unsigned)v[1] << 8) | ((unsigned)v[2] << 16) | ((unsigned)v[3] << 24);
                ov.y = (unsigned)v[4] | ((unsigned)v[5] << 8) | ((unsigned)v[6] << 16) | ((unsigned)v[7] << 24);
                *reinterpret_cast<uint2 *>(d) = ov;
            } else {
                for (int k = 0; k < npx; k++) d[k] = (uint8_t)(avg ? (d[k] + v[k] + 1) >> 1 : v[k]);
            }
        }
        __syncwarp();                                                                    // shared window is reused by the next operation
        m0 = m1; m1 = m2; w0 = w1;
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
    const long long blocks = (n + WARPS * QK - 1) / (WARPS * QK);
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
