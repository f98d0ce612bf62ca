// h264lf_hbd.cu — libavcodec's H.264 in-loop deblocking filters for 9 / 10 / 12 / 14 bit samples on sm_100a: the loop-filter members of
// H264DSPContext as ff_h264dsp_init(c, depth, chroma_format_idc) installs them above 8 bits (libavcodec/h264dsp.c:109-158).
//
// Reference semantics reproduced bit for bit (checker: the 16-bit part of oracle/h264lf_oracle.c), libavcodec/h264dsp_template.c:103-340 with
// pixel = uint16_t: alpha and beta are scaled by << (depth - 8), the luma tc0 by * (1 << (depth - 8)), the chroma tc is
// ((tc0 - 1U) << (depth - 8)) + 1, and results are clipped to the sample depth.  One thread per line across an edge, like h264lf.cu.
#include "common.h"
#include "h264lf_hbd.h"
#include <cstring>

namespace {

struct LfShapeH { int intra, chroma, vert, iters; };
#define LF_SHAPE_BODY                                                                                  \
    LfShapeH s;                                                                                        \
    s.intra = (kind >= 3 && kind <= 5) || (kind >= 9 && kind <= 11) || kind >= 14;                     \
    s.chroma = kind >= 6;                                                                              \
    s.vert = kind == 0 || kind == 3 || kind == 6 || kind == 9;                                         \
    s.iters = kind < 6 ? ((kind == 2 || kind == 5) ? 2 : 4)                                            \
            : kind < 12 ? ((kind == 8 || kind == 11) ? 1 : 2)                                          \
            : ((kind == 12 || kind == 14) ? 4 : 2);                                                    \
    return s;
__device__ __forceinline__ LfShapeH lf_shape_h(int kind) { LF_SHAPE_BODY }      // same kind numbering as b200_h264_loop_filter_batch_device
inline LfShapeH lf_shape_host(int kind) { LF_SHAPE_BODY }
__device__ __forceinline__ int clip3h(int v, int lo, int hi) { return min(max(v, lo), hi); }

// edge e: kinds[e], q0 of its first line at base + off[e] (BYTES), alpha[e], beta[e] (the 8-bit table values), tc0 + 4 e; stride in BYTES
__global__ void __launch_bounds__(128)
h264_loop_filter_hbd_kernel(long long nedges, const uint8_t *kinds, uint8_t *base, const long long *off, long long stride, const uint8_t *alphas,
                            const uint8_t *betas, const int8_t *tc0s, int depth)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = t >> 4;
    const int line = (int)(t & 15);
    if (e >= nedges) return;
    const LfShapeH S = lf_shape_h(kinds[e]);
    if (line >= 4 * S.iters) return;
    const long long st = stride / 2;
    const long long xs = S.vert ? st : 1, ys = S.vert ? 1 : st;
    unsigned short *pix = reinterpret_cast<unsigned short *>(base + off[e]) + line * ys;
    const int sh = depth - 8, maxv = (1 << depth) - 1;
    const int alpha = alphas[e] << sh, beta = betas[e] << sh;
    const int p0 = pix[-1 * xs], p1 = pix[-2 * xs], q0 = pix[0], q1 = pix[1 * xs];
    if (!S.intra) {
        const int raw = tc0s[4 * e + line / S.iters];
        const int t0 = S.chroma ? (int)(((unsigned)raw - 1u) << sh) + 1 : raw * (1 << sh);
        if (S.chroma ? t0 <= 0 : t0 < 0) return;
        if (!(abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta)) return;
        int tc = t0;
        if (!S.chroma) {
            const int p2 = pix[-3 * xs], q2 = pix[2 * xs];
            if (abs(p2 - p0) < beta) {
                if (t0) pix[-2 * xs] = (unsigned short)(p1 + clip3h(((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -t0, t0));
                tc++;
            }
            if (abs(q2 - q0) < beta) {
                if (t0) pix[xs] = (unsigned short)(q1 + clip3h(((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -t0, t0));
                tc++;
            }
        }
        const int delta = clip3h((((q0 - p0) * 4) + (p1 - q1) + 4) >> 3, -tc, tc);
        pix[-xs] = (unsigned short)clip3h(p0 + delta, 0, maxv);
        pix[0] = (unsigned short)clip3h(q0 - delta, 0, maxv);
        return;
    }
    if (!(abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta)) return;
    if (S.chroma) {
        pix[-xs] = (unsigned short)((2 * p1 + p0 + q1 + 2) >> 2);
        pix[0] = (unsigned short)((2 * q1 + q0 + p1 + 2) >> 2);
        return;
    }
    const int p2 = pix[-3 * xs], q2 = pix[2 * xs];
    if (abs(p0 - q0) < ((alpha >> 2) + 2)) {
        if (abs(p2 - p0) < beta) {
            const int p3 = pix[-4 * xs];
            pix[-1 * xs] = (unsigned short)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            pix[-2 * xs] = (unsigned short)((p2 + p1 + p0 + q0 + 2) >> 2);
            pix[-3 * xs] = (unsigned short)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else
            pix[-1 * xs] = (unsigned short)((2 * p1 + p0 + q1 + 2) >> 2);
        if (abs(q2 - q0) < beta) {
            const int q3 = pix[3 * xs];
            pix[0 * xs] = (unsigned short)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            pix[1 * xs] = (unsigned short)((p0 + q0 + q1 + q2 + 2) >> 2);
            pix[2 * xs] = (unsigned short)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else
            pix[0 * xs] = (unsigned short)((2 * q1 + q0 + p1 + 2) >> 2);
    } else {
        pix[-1 * xs] = (unsigned short)((2 * p1 + p0 + q1 + 2) >> 2);
        pix[0 * xs] = (unsigned short)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}

void die(const char *what)
{
    fprintf(stderr, "libb200dsp: high-bit-depth h264 loop filter failed: %s (%s)\n", what, b200_last_error());
    abort();
}

// drop-in: one edge through the device (host pointers).  The touched window is 4 samples either side of the edge, 16 lines at most.
template <int DEPTH>
void host_op(int kind, uint8_t *pix, ptrdiff_t stride, int alpha, int beta, const int8_t *tc0)
{
    B200Device *dev = b200_default_device();
    if (!dev) die("no device");
    if (cudaSetDevice(dev->ordinal) != cudaSuccess) die("cudaSetDevice");
    const LfShapeH S = lf_shape_host(kind);
    const int lines = 4 * S.iters;
    const int w = S.vert ? lines : 8, h = S.vert ? 8 : lines;     // in samples
    uint8_t *origin = S.vert ? pix - 4 * stride : pix - 4 * 2;
    const size_t pitch = 32;
    B200_LOCK_DEVICE(dev);
    uint8_t *scr = (uint8_t *)b200_scratch(dev, pitch * 16 + 64);
    if (!scr) die("scratch");
    uint8_t *win = scr, *meta = scr + pitch * 16;
    cudaStream_t st = dev->stream;
    if (b200_h2d_rows(win, pitch, origin, stride, (size_t)w * 2, (size_t)h, st) != cudaSuccess) die("h2d");
    struct { long long off; int8_t tc[4]; uint8_t kind, alpha, beta, pad; } m;
    m.off = S.vert ? 4 * (long long)pitch : 8;
    for (int i = 0; i < 4; i++) m.tc[i] = tc0 ? tc0[i] : 0;
    m.kind = (uint8_t)kind; m.alpha = (uint8_t)alpha; m.beta = (uint8_t)beta; m.pad = 0;
    if (cudaMemcpyAsync(meta, &m, sizeof(m), cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d meta");
    h264_loop_filter_hbd_kernel<<<1, 128, 0, st>>>(1, meta + 12, win, (const long long *)meta, (long long)pitch, meta + 13, meta + 14, (const int8_t *)(meta + 8), DEPTH);
    B200_LAUNCHED();
    if (cudaGetLastError() != cudaSuccess) die("launch");
    if (b200_d2h_rows(origin, stride, win, pitch, (size_t)w * 2, (size_t)h, st) != cudaSuccess) die("d2h");
    if (cudaStreamSynchronize(st) != cudaSuccess) die("sync");
}

template <int DEPTH, int KIND> void tab_tc(uint8_t *pix, ptrdiff_t stride, int alpha, int beta, int8_t *tc0) { host_op<DEPTH>(KIND, pix, stride, alpha, beta, tc0); }
template <int DEPTH, int KIND> void tab_intra(uint8_t *pix, ptrdiff_t stride, int alpha, int beta) { host_op<DEPTH>(KIND, pix, stride, alpha, beta, nullptr); }

template <int D>
void fill(B200H264LoopFilterContext *c, bool c422)
{
    c->v_loop_filter_luma = tab_tc<D, 0>; c->h_loop_filter_luma = tab_tc<D, 1>; c->h_loop_filter_luma_mbaff = tab_tc<D, 2>;
    c->v_loop_filter_luma_intra = tab_intra<D, 3>; c->h_loop_filter_luma_intra = tab_intra<D, 4>; c->h_loop_filter_luma_mbaff_intra = tab_intra<D, 5>;
    c->v_loop_filter_chroma = tab_tc<D, 6>;
    c->h_loop_filter_chroma = c422 ? tab_tc<D, 12> : tab_tc<D, 7>;
    c->h_loop_filter_chroma_mbaff = c422 ? tab_tc<D, 13> : tab_tc<D, 8>;
    c->v_loop_filter_chroma_intra = tab_intra<D, 9>;
    c->h_loop_filter_chroma_intra = c422 ? tab_intra<D, 14> : tab_intra<D, 10>;
    c->h_loop_filter_chroma_mbaff_intra = c422 ? tab_intra<D, 15> : tab_intra<D, 11>;
}

} // namespace

bool h264lf_hbd_fill(B200H264LoopFilterContext *c, int bit_depth, int chroma_format_idc)
{
    const bool c422 = chroma_format_idc > 1;                      // h264dsp.c:116-132
    switch (bit_depth) {
    case 9:  fill<9>(c, c422);  return true;
    case 10: fill<10>(c, c422); return true;
    case 12: fill<12>(c, c422); return true;
    case 14: fill<14>(c, c422); return true;
    }
    return false;
}

B200_API int b200_h264_loop_filter_hbd_batch_device(B200Device *dev, int bit_depth, int64_t nedges, const uint8_t *kinds, uint8_t *pix,
                                                    const int64_t *pix_off, ptrdiff_t stride, const uint8_t *alpha, const uint8_t *beta, const int8_t *tc0)
{
    if (!dev) dev = b200_default_device();
    if (!dev) return B200_ENODEV;
    if (nedges < 0 || stride < 0 || (stride & 1)) return B200_EINVAL;
    if (bit_depth != 9 && bit_depth != 10 && bit_depth != 12 && bit_depth != 14) return B200_ENOSYS;
    if (nedges == 0) return 0;
    if (!kinds || !pix || !pix_off || !alpha || !beta || !tc0 || ((uintptr_t)pix & 1)) return B200_EINVAL;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    const long long threads = nedges * 16, grid = (threads + 127) / 128;
    if (grid > 0x7fffffffLL) return B200_EINVAL;
    h264_loop_filter_hbd_kernel<<<(unsigned)grid, 128, 0, dev->stream>>>(nedges, kinds, pix, (const long long *)pix_off, (long long)stride, alpha, beta, tc0, bit_depth);
    B200_LAUNCHED();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}
