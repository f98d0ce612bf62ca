// h264idct_hbd.cu — libavcodec H.264 residual adds for 9 / 10 / 12 / 14 bit samples on sm_100a: ff_h264_idct_add / idct8_add / idct_dc_add /
// idct8_dc_add of h264idct_template.c:33-181 instantiated per depth (dctcoef = int32_t, pixel = uint16_t; h264dsp.c:66-158).
//
// Reference semantics reproduced bit for bit (checker: the 16-bit part of oracle/idct_oracle.c): + 32 on the DC coefficient, the 4-point /
// 8-point butterflies on 32-bit values (sums mod 2^32 like the reference's SUINT), >> 6, clip to the sample depth, block cleared.
// Plain version, a thread per block (the 8-bit path in h264idct.cu carries the tuned kernel).
#include "common.h"
#include "h264idct_hbd.h"
#include <cstring>

namespace {

__device__ __forceinline__ void idct8_1d(const int *s, int stride, int *o)
{
    const unsigned a0 = (unsigned)s[0 * stride] + (unsigned)s[4 * stride], a2 = (unsigned)s[0 * stride] - (unsigned)s[4 * stride];
    const unsigned a4 = (unsigned)(s[2 * stride] >> 1) - (unsigned)s[6 * stride], a6 = (unsigned)(s[6 * stride] >> 1) + (unsigned)s[2 * stride];
    const unsigned b0 = a0 + a6, b2 = a2 + a4, b4 = a2 - a4, b6 = a0 - a6;
    const int a1 = (int)(-(unsigned)s[3 * stride] + (unsigned)s[5 * stride] - (unsigned)s[7 * stride] - (unsigned)(s[7 * stride] >> 1));
    const int a3 = (int)((unsigned)s[1 * stride] + (unsigned)s[7 * stride] - (unsigned)s[3 * stride] - (unsigned)(s[3 * stride] >> 1));
    const int a5 = (int)(-(unsigned)s[1 * stride] + (unsigned)s[7 * stride] + (unsigned)s[5 * stride] + (unsigned)(s[5 * stride] >> 1));
    const int a7 = (int)((unsigned)s[3 * stride] + (unsigned)s[5 * stride] + (unsigned)s[1 * stride] + (unsigned)(s[1 * stride] >> 1));
    const unsigned b1 = (unsigned)(a7 >> 2) + (unsigned)a1, b3 = (unsigned)a3 + (unsigned)(a5 >> 2);
    const unsigned b5 = (unsigned)(a3 >> 2) - (unsigned)a5, b7 = (unsigned)a7 - (unsigned)(a1 >> 2);
    o[0] = (int)(b0 + b7); o[7] = (int)(b0 - b7); o[1] = (int)(b2 + b5); o[6] = (int)(b2 - b5);
    o[2] = (int)(b4 + b3); o[5] = (int)(b4 - b3); o[3] = (int)(b6 + b1); o[4] = (int)(b6 - b1);
}

// kind 0: 4x4, 1: 8x8, 2: 4x4 DC only, 3: 8x8 DC only.  blk_off in int32 elements, dst_off / stride in bytes.
template <int KIND>
__global__ void __launch_bounds__(128)
h264_idct_hbd_kernel(long long n, int32_t *blocks, const long long *__restrict__ blk_off, uint8_t *dst, const long long *__restrict__ dst_off,
                     long long stride, int depth)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int N = (KIND & 1) ? 8 : 4;
    int32_t *b = blocks + blk_off[i];
    unsigned short *d = reinterpret_cast<unsigned short *>(dst + dst_off[i]);
    const long long st = stride / 2;
    const int maxv = (1 << depth) - 1;
    auto put = [&](int x, int y, int v) { unsigned short *q = d + y * st + x; *q = (unsigned short)min(max((int)*q + v, 0), maxv); };
    if (KIND >= 2) {
        const int dc = (b[0] + 32) >> 6;
        b[0] = 0;
        for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) put(x, y, dc);
        return;
    }
    int c[N * N];
#pragma unroll
    for (int k = 0; k < N * N; k++) c[k] = b[k];
    c[0] = (int)((unsigned)c[0] + 32u);
    if (KIND == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) {                              // columns: block[k + 4 r]
            const unsigned z0 = (unsigned)c[k] + (unsigned)c[k + 8], z1 = (unsigned)c[k] - (unsigned)c[k + 8];
            const unsigned z2 = (unsigned)(c[k + 4] >> 1) - (unsigned)c[k + 12], z3 = (unsigned)c[k + 4] + (unsigned)(c[k + 12] >> 1);
            c[k] = (int)(z0 + z3); c[k + 4] = (int)(z1 + z2); c[k + 8] = (int)(z1 - z2); c[k + 12] = (int)(z0 - z3);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {                              // rows: block[r + 4 k] -> column k of the picture block
            const unsigned z0 = (unsigned)c[4 * k] + (unsigned)c[4 * k + 2], z1 = (unsigned)c[4 * k] - (unsigned)c[4 * k + 2];
            const unsigned z2 = (unsigned)(c[4 * k + 1] >> 1) - (unsigned)c[4 * k + 3], z3 = (unsigned)c[4 * k + 1] + (unsigned)(c[4 * k + 3] >> 1);
            put(k, 0, (int)(z0 + z3) >> 6); put(k, 1, (int)(z1 + z2) >> 6); put(k, 2, (int)(z1 - z2) >> 6); put(k, 3, (int)(z0 - z3) >> 6);
        }
    } else {
        int o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {                              // first pass: elements k + 8 r
            idct8_1d(c + k, 8, o);
#pragma unroll
            for (int r = 0; r < 8; r++) c[k + 8 * r] = o[r];
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {                              // second pass: elements r + 8 k -> column k of the picture block
            idct8_1d(c + 8 * k, 1, o);
#pragma unroll
            for (int r = 0; r < 8; r++) put(k, r, o[r] >> 6);
        }
    }
#pragma unroll
    for (int k = 0; k < N * N; k++) b[k] = 0;
}

void die(const char *what)
{
    fprintf(stderr, "libb200dsp: high-bit-depth h264 idct failed: %s (%s)\n", what, b200_last_error());
    abort();
}

int launch(cudaStream_t st, int kind, int depth, long long n, int32_t *blocks, const long long *blk_off, uint8_t *dst, const long long *dst_off, long long stride)
{
    if (n <= 0) return 0;
    const long long ctas = (n + 127) / 128;
    if (ctas > 0x7fffffffLL) return B200_EINVAL;
    switch (kind) {
    case 0: h264_idct_hbd_kernel<0><<<(unsigned)ctas, 128, 0, st>>>(n, blocks, blk_off, dst, dst_off, stride, depth); break;
    case 1: h264_idct_hbd_kernel<1><<<(unsigned)ctas, 128, 0, st>>>(n, blocks, blk_off, dst, dst_off, stride, depth); break;
    case 2: h264_idct_hbd_kernel<2><<<(unsigned)ctas, 128, 0, st>>>(n, blocks, blk_off, dst, dst_off, stride, depth); break;
    default: h264_idct_hbd_kernel<3><<<(unsigned)ctas, 128, 0, st>>>(n, blocks, blk_off, dst, dst_off, stride, depth); break;
    }
    B200_LAUNCHED();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// drop-in: one block through the device (host pointers; `block` holds int32 coefficients behind the reference's int16_t * type)
template <int DEPTH, int KIND>
void host_fn(uint8_t *dst, int16_t *block16, ptrdiff_t stride)
{
    constexpr int N = (KIND & 1) ? 8 : 4, NC = N * N;
    int32_t *block = reinterpret_cast<int32_t *>(block16);
    B200Device *dev = b200_default_device();
    if (!dev) die("no device");
    if (cudaSetDevice(dev->ordinal) != cudaSuccess) die("cudaSetDevice");
    B200_LOCK_DEVICE(dev);
    uint8_t *scr = (uint8_t *)b200_scratch(dev, 1024);
    if (!scr) die("scratch");
    int32_t *dblk = (int32_t *)scr;                       // 256 B
    uint8_t *dpix = scr + 256;                            // N rows, pitch 16 bytes
    long long *meta = (long long *)(scr + 512);
    cudaStream_t st = dev->stream;
    const long long m[2] = { 0, 0 };
    if (cudaMemcpyAsync(dblk, block, (KIND >= 2 ? 1 : NC) * sizeof(int32_t), cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d block");
    if (b200_h2d_rows(dpix, 16, dst, stride, N * 2, N, st) != cudaSuccess) die("h2d dst");
    if (cudaMemcpyAsync(meta, m, sizeof(m), cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d meta");
    if (launch(st, KIND, DEPTH, 1, dblk, meta, dpix, meta + 1, 16) < 0) die("launch");
    if (b200_d2h_rows(dst, stride, dpix, 16, N * 2, N, st) != cudaSuccess) die("d2h");
    if (cudaStreamSynchronize(st) != cudaSuccess) die("sync");
    if (KIND >= 2) block[0] = 0; else memset(block, 0, NC * sizeof(int32_t));     // what the reference leaves behind
}

template <int DEPTH>
void fill(B200H264IDCTContext *c)
{
    c->idct_add = host_fn<DEPTH, 0>; c->idct8_add = host_fn<DEPTH, 1>; c->idct_dc_add = host_fn<DEPTH, 2>; c->idct8_dc_add = host_fn<DEPTH, 3>;
}

} // namespace

bool h264idct_hbd_fill(B200H264IDCTContext *c, int bit_depth)
{
    switch (bit_depth) {
    case 9:  fill<9>(c);  return true;
    case 10: fill<10>(c); return true;
    case 12: fill<12>(c); return true;
    case 14: fill<14>(c); return true;
    }
    return false;
}

B200_API int b200_h264_idct_hbd_batch_device(B200Device *dev, int bit_depth, int kind, int64_t n, int32_t *blocks, const int64_t *blk_off,
                                             uint8_t *dst, const int64_t *dst_off, ptrdiff_t stride)
{
    if (!dev || n < 0 || !blocks || !blk_off || !dst || !dst_off || kind < 0 || kind > 3) return B200_EINVAL;
    if (bit_depth != 9 && bit_depth != 10 && bit_depth != 12 && bit_depth != 14) return B200_ENOSYS;
    if (((uintptr_t)blocks & 3) || ((uintptr_t)dst & 1) || (stride & 1)) return B200_EINVAL;
    if (n == 0) return 0;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    return launch(dev->stream, kind, bit_depth, n, blocks, (const long long *)blk_off, dst, (const long long *)dst_off, (long long)stride);
}
