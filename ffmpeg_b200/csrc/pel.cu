// pel.cu — libavcodec h264qpel (8 bit), h264chroma (8 bit) and hpeldsp motion-compensation interpolation on sm_100a (C ABI: "h264qpel / hpeldsp").
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
#include <algorithm>
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

// h264chroma: bilinear eighth-pel (h264chroma_template.c:27-176).  16 lanes per operation, a lane owns up to 4 consecutive
// pixels of one row; the x == 0 / y == 0 cases skip the right column / lower row like the reference's D == 0 branches.
constexpr int CH_LANES = 16;
__global__ void __launch_bounds__(32 * WARPS)
chroma_kernel(long long n, const uint8_t *op, const uint8_t *hh, const uint8_t *xy, uint8_t *dst, const int64_t *dst_off,
              const uint8_t *src, const int64_t *src_off, long long stride)
{
    const long long i = ((long long)blockIdx.x * (32 * WARPS) + threadIdx.x) / CH_LANES;
    if (i >= n) return;
    const int sub = threadIdx.x & (CH_LANES - 1);
    const int o = op[i], h = hh[i], fx = xy[i] & 7, fy = (xy[i] >> 3) & 7;
    const int avg = o & 1, w = 8 >> ((o >> 1) & 3);
    const int A = (8 - fx) * (8 - fy), B = fx * (8 - fy), C = (8 - fx) * fy, D = fx * fy;
    const int segs = w > 4 ? 2 : 1, npx = w < 4 ? w : 4;
    const uint8_t *sp = src + src_off[i];
    uint8_t *dp = dst + dst_off[i];
    for (int y = sub / segs; y < h; y += CH_LANES / segs) {
        const int x0 = (sub % segs) * 4;
        const uint8_t *p = sp + (long long)y * stride + x0;
        int t[5], b[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const bool in = k < npx || (k == npx && fx);                       // right neighbour only when B or D is non-zero
            t[k] = in ? __ldg(p + k) : 0;
            b[k] = (in && fy) ? __ldg(p + stride + k) : 0;                     // lower row only when C or D is non-zero
        }
        uint8_t *d = dp + (long long)y * stride + x0;
        uint32_t pack = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int v = (A * t[k] + B * t[k + 1] + C * b[k] + D * b[k + 1] + 32) >> 6;
            pack |= (uint32_t)v << (8 * k);
        }
        if (npx == 4 && ((uintptr_t)d & 3) == 0) {
            uint32_t *d4 = (uint32_t *)d;
            *d4 = avg ? __vavgu4(*d4, pack) : pack;
        } else {
            for (int k = 0; k < npx; k++) {
                const int v = (pack >> (8 * k)) & 255;
                d[k] = (uint8_t)(avg ? (d[k] + v + 1) >> 1 : v);
            }
        }
    }
}

// emulated_edge_mc (videodsp_template.c:24-101) as a clamped gather: one warp per window, lanes across columns.
// geom[4*i..] = block_w, block_h, src_x, src_y; origin[i] = offset of that picture's sample (0, 0) from `src`.
__global__ void __launch_bounds__(32 * WARPS)
edge_kernel(long long n, uint8_t *buf, const int64_t *buf_off, long long buf_ls, const uint8_t *src, const int64_t *origin,
            long long src_ls, const int32_t *geom, int w, int h)
{
    const long long i = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const int4 gm = __ldg((const int4 *)geom + i);
    const int bw = gm.x, bh = gm.y, sx = gm.z, sy = gm.w;
    const uint8_t *pic = src + origin[i];
    uint8_t *out = buf + buf_off[i];
    for (int x = lane; x < bw; x += 32) {
        const int px = min(max(sx + x, 0), w - 1);
        for (int y = 0; y < bh; y++) {
            const int py = min(max(sy + y, 0), h - 1);
            out[(long long)y * buf_ls + x] = __ldg(pic + (long long)py * src_ls + px);
        }
    }
}

void die(const char *what)
{
    fprintf(stderr, "libb200dsp: motion compensation failed: %s (%s)\n", what, b200_last_error());
    abort();
}

enum HostKind { HOST_QPEL, HOST_HPEL, HOST_CHROMA };

// one block through the device for the drop-in tables.  Only the source rectangle the reference function itself reads is
// copied from the caller's buffer (bx/ax columns left/right, by/ay rows above/below the block).
void host_op(HostKind kind, int o, int h, int w, uint8_t *dst, const uint8_t *src, ptrdiff_t stride,
             int bx, int ax, int by, int ay, int fxy = 0)
{
    B200Device *dev = b200_default_device();
    if (!dev) die("no device");
    if (stride < 0) die("negative stride");
    if (cudaSetDevice(dev->ordinal) != cudaSuccess) die("cudaSetDevice");
    const int before = kind == HOST_QPEL ? 2 : 0, after = kind == HOST_QPEL ? 3 : 1;   // the window the kernels may touch
    const int sh = h + before + after;
    const size_t pitch = 32;
    uint8_t *scr = (uint8_t *)b200_scratch(dev, pitch * (sh + h) + 256);
    if (!scr) die("scratch");
    uint8_t *dsrc = scr, *ddst = scr + pitch * sh;
    uint8_t *meta = scr + pitch * (sh + h);           // op, h, offsets
    cudaStream_t st = dev->stream;
    if (cudaMemcpy2DAsync(dsrc + (before - by) * pitch + (before - bx), pitch, src - by * stride - bx, (size_t)stride,
                          w + bx + ax, h + by + ay, cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d src");
    if (cudaMemcpy2DAsync(ddst, pitch, dst, (size_t)stride, w, h, cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d dst");
    struct { int64_t doff, soff; uint8_t op, h, xy; } m = { 0, (int64_t)(before * pitch + before), (uint8_t)o, (uint8_t)h, (uint8_t)fxy };
    if (cudaMemcpyAsync(meta, &m, sizeof(m), cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d meta");
    const int64_t *doff = (const int64_t *)meta, *soff = doff + 1;
    const uint8_t *dop = meta + 16, *dh = meta + 17, *dxy = meta + 18;
    // dst and src live in one buffer with the same pitch, like the reference's single stride
    if (kind == HOST_QPEL)      qpel_kernel<<<1, 32 * WARPS, 0, st>>>(1, dop, ddst, doff, dsrc, soff, (long long)pitch);
    else if (kind == HOST_HPEL) hpel_kernel<<<1, 32 * WARPS, 0, st>>>(1, dop, dh, ddst, doff, dsrc, soff, (long long)pitch);
    else                        chroma_kernel<<<1, 32 * WARPS, 0, st>>>(1, dop, dh, dxy, ddst, doff, dsrc, soff, (long long)pitch);
    B200_LAUNCHED();
    if (cudaMemcpy2DAsync(dst, (size_t)stride, ddst, pitch, w, h, cudaMemcpyDeviceToHost, st) != cudaSuccess) die("d2h");
    if (cudaStreamSynchronize(st) != cudaSuccess) die("sync");
}

template <int AVG, int SIDX, int POS>
void qpel_tab(uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    constexpr int X = POS & 3, Y = POS >> 2;           // h264qpel_template.c: mc{x}{y} filters horizontally iff x, vertically iff y
    host_op(HOST_QPEL, AVG | (SIDX << 1) | (POS << 3), 16 >> SIDX, 16 >> SIDX, dst, src, stride, X ? 2 : 0, X ? 3 : 0, Y ? 2 : 0, Y ? 3 : 0);
}
template <int TAB, int SIDX, int XY>
void hpel_tab(uint8_t *block, const uint8_t *pixels, ptrdiff_t line_size, int h)
{
    host_op(HOST_HPEL, TAB | (SIDX << 2) | (XY << 4), h, 16 >> SIDX, block, pixels, line_size, 0, XY & 1, 0, XY >> 1);
}
template <int AVG, int IDX>
void chroma_tab(uint8_t *dst, const uint8_t *src, ptrdiff_t stride, int h, int x, int y)
{
    if ((unsigned)x > 7 || (unsigned)y > 7) die("h264chroma: x, y must be in 0..7");     // av_assert2 in the reference
    host_op(HOST_CHROMA, AVG | (IDX << 1), h, 8 >> IDX, dst, src, stride, 0, x ? 1 : 0, 0, y ? 1 : 0, x | (y << 3));
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

namespace {
void edge_tab(uint8_t *buf, const uint8_t *src, ptrdiff_t buf_linesize, ptrdiff_t src_linesize, int block_w, int block_h,
              int src_x, int src_y, int w, int h)
{
    if (!w || !h) return;                                          // videodsp_template.c:33-34
    if (block_w <= 0 || block_h <= 0) return;
    B200Device *dev = b200_default_device();
    if (!dev) die("no device");
    if (buf_linesize < 0 || src_linesize < 0) die("negative linesize");
    if (cudaSetDevice(dev->ordinal) != cudaSuccess) die("cudaSetDevice");
    // the part of the picture the window can reach after clamping
    const int x0 = std::min(std::max(src_x, 0), w - 1), x1 = std::min(std::max(src_x + block_w - 1, 0), w - 1);
    const int y0 = std::min(std::max(src_y, 0), h - 1), y1 = std::min(std::max(src_y + block_h - 1, 0), h - 1);
    const size_t rw = (size_t)(x1 - x0 + 1), rh = (size_t)(y1 - y0 + 1);
    const size_t rp = (rw + 15) & ~(size_t)15, bp = ((size_t)block_w + 15) & ~(size_t)15;
    uint8_t *scr = (uint8_t *)b200_scratch(dev, rp * rh + bp * block_h + 256);
    if (!scr) die("scratch");
    uint8_t *drect = scr, *dbuf = scr + rp * rh, *meta = dbuf + bp * block_h;
    meta += (16 - ((uintptr_t)meta & 15)) & 15;
    cudaStream_t st = dev->stream;
    const uint8_t *pic = src - (ptrdiff_t)src_y * src_linesize - src_x;
    if (cudaMemcpy2DAsync(drect, rp, pic + (ptrdiff_t)y0 * src_linesize + x0, (size_t)src_linesize, rw, rh, cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d");
    struct { int32_t g[4]; int64_t boff, origin; } m = { { block_w, block_h, src_x, src_y }, 0, -((int64_t)y0 * (int64_t)rp + x0) };
    if (cudaMemcpyAsync(meta, &m, sizeof(m), cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d meta");
    edge_kernel<<<1, 32 * WARPS, 0, st>>>(1, dbuf, (const int64_t *)(meta + 16), (long long)bp, drect, (const int64_t *)(meta + 24),
                                          (long long)rp, (const int32_t *)meta, w, h);
    B200_LAUNCHED();
    if (cudaMemcpy2DAsync(buf, (size_t)buf_linesize, dbuf, bp, (size_t)block_w, (size_t)block_h, cudaMemcpyDeviceToHost, st) != cudaSuccess) die("d2h");
    if (cudaStreamSynchronize(st) != cudaSuccess) die("sync");
}
void prefetch_nop(const uint8_t *, ptrdiff_t, int) {}
} // namespace

B200_API int b200_videodsp_init(B200VideoDSPContext *c, int bpc)
{
    if (!c) return B200_EINVAL;
    if (bpc > 8) return B200_ENOSYS;                               // videodsp.c:41-45 installs the 16 bit template above 8
    if (!b200_default_device()) return B200_ENODEV;
    c->emulated_edge_mc = edge_tab;
    c->prefetch = prefetch_nop;                                    // videodsp.c:34-36: the C prefetch is an empty function
    return 0;
}

B200_API int b200_emulated_edge_mc_batch_device(B200Device *dev, int64_t n, uint8_t *buf, const int64_t *buf_off,
                                                ptrdiff_t buf_linesize, const uint8_t *src, const int64_t *origin,
                                                ptrdiff_t src_linesize, const int32_t *geom, int w, int h)
{
    if (!dev || n < 0 || !buf || !buf_off || !src || !origin || !geom || w < 0 || h < 0) return B200_EINVAL;
    if (((uintptr_t)geom & 15) != 0) return B200_EINVAL;
    if (n == 0 || !w || !h) return 0;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    const long long blocks = (n + WARPS - 1) / WARPS;
    if (blocks > 0x7fffffffLL) return B200_EINVAL;
    edge_kernel<<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, buf, buf_off, buf_linesize, src, origin, src_linesize, geom, w, h);
    B200_LAUNCHED();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

B200_API int b200_h264chroma_init(B200H264ChromaContext *c, int bit_depth)
{
    if (!c) return B200_EINVAL;
    if (bit_depth != 8) return B200_ENOSYS;                        // h264chroma.c:46-50 installs the 16 bit template above 8
    if (!b200_default_device()) return B200_ENODEV;
    memset(c, 0, sizeof(*c));                                      // entry [3] stays NULL like the reference's
    c->put_h264_chroma_pixels_tab[0] = chroma_tab<0, 0>; c->put_h264_chroma_pixels_tab[1] = chroma_tab<0, 1>; c->put_h264_chroma_pixels_tab[2] = chroma_tab<0, 2>;
    c->avg_h264_chroma_pixels_tab[0] = chroma_tab<1, 0>; c->avg_h264_chroma_pixels_tab[1] = chroma_tab<1, 1>; c->avg_h264_chroma_pixels_tab[2] = chroma_tab<1, 2>;
    return 0;
}

B200_API int b200_h264chroma_batch_device(B200Device *dev, int64_t n, const uint8_t *op, const uint8_t *h, const uint8_t *xy,
                                          uint8_t *dst, const int64_t *dst_off, const uint8_t *src, const int64_t *src_off,
                                          ptrdiff_t stride)
{
    if (!dev || n < 0 || !op || !h || !xy || !dst || !dst_off || !src || !src_off) return B200_EINVAL;
    if (n == 0) return 0;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    const int per_cta = 32 * WARPS / CH_LANES;
    const long long blocks = (n + per_cta - 1) / per_cta;
    if (blocks > 0x7fffffffLL) return B200_EINVAL;
    chroma_kernel<<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, op, h, xy, dst, dst_off, src, src_off, stride);
    B200_LAUNCHED();
    B200_CUDA_OK(cudaGetLastError());
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
