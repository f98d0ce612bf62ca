// mecmp_dct.cu — the transform-domain members of MECmpContext:
//   dct_sad8x8_c / dct_sad16_c         libavcodec/me_cmp.c:614-622, 952   sum |DCT(blk1 - blk2)|
//   dct_max8x8_c / dct_max16_c         libavcodec/me_cmp.c:678-693, 956   max |DCT(blk1 - blk2)| (16 wide: the maxima of the 8x8 blocks add up)
//   dct264_sad8x8_c / dct264_sad16_c   libavcodec/me_cmp.c:624-675, 954   the H.264 8x8 integer transform instead of the DCT
// The DCT is FDCTDSPContext.fdct as ff_fdctdsp_init picks it for 8-bit samples (libavcodec/fdctdsp.c:27-45): ff_jpeg_fdct_islow_8
// (jfdctint_template.c:173-340) unless dct_algo is FF_DCT_FASTINT, then ff_fdct_ifast (jfdctfst.c:140-284).  The reference's entries
// read those pointers from the encoder context they are handed; the tables here cannot (its layout is private to the encoder), so
// the choice is a library setting, b200_me_cmp_set_dct_algo(), like the nsse weight.
//
// Work split: one warp per comparison, eight lanes per 8x8 block (a 16 wide comparison has two or four blocks), a lane holds one row of
// the difference block in registers.  Row pass in the lane, 8x8 transpose across the eight lanes with three rounds of xor-shuffles,
// column pass in the lane (it now holds a column), then the |.| sum or maximum over the eight lanes and the sum over the blocks.
#include "mecmp_dct.h"
#include "fdct_dev.cuh"
#include <atomic>

namespace {

std::atomic<int> g_dct_fast{0};                                  // 1: ff_fdct_ifast

// 1-D step of the H.264 8x8 forward transform (exact integers, shifts are arithmetic)
__device__ __forceinline__ void fwd264_8(int (&v)[8])
{
    const int s07 = v[0] + v[7], s16 = v[1] + v[6], s25 = v[2] + v[5], s34 = v[3] + v[4];
    const int d07 = v[0] - v[7], d16 = v[1] - v[6], d25 = v[2] - v[5], d34 = v[3] - v[4];
    const int a0 = s07 + s34, a1 = s16 + s25, a2 = s07 - s34, a3 = s16 - s25;
    const int a4 = d16 + d25 + (d07 + (d07 >> 1)), a5 = d07 - d34 - (d25 + (d25 >> 1));
    const int a6 = d07 + d34 - (d16 + (d16 >> 1)), a7 = d16 - d25 + (d34 + (d34 >> 1));
    v[0] = a0 + a1; v[1] = a4 + (a7 >> 2); v[2] = a2 + (a3 >> 1); v[3] = a5 + (a6 >> 2);
    v[4] = a0 - a1; v[5] = a6 - (a5 >> 2); v[6] = (a2 >> 1) - a3; v[7] = (a4 >> 2) - a7;
}

// kind 0: dct_sad, 1: dct_max, 2: dct264_sad; nblk = 1 (8x8), 2 (16 wide, h = 8) or 4 (16 wide, h = 16)
__global__ void __launch_bounds__(256)
me_dct_kernel(int kind, int fast, int nblk, const uint8_t *f1, const uint8_t *f2, long long stride, const int64_t *off1, const int64_t *off2,
              long long n, int32_t *out)
{
    const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31, blk = lane >> 3, row = lane & 7;
    int v[8];
    {
        const bool on = blk < nblk;
        const long long o = (long long)((blk >> 1) * 8 + row) * stride + (blk & 1) * 8;
        const uint8_t *a = f1 + off1[i] + o, *b = f2 + off2[i] + o;
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = on ? (int)a[k] - (int)b[k] : 0;       // diff_pixels: blk1 - blk2; an idle lane group transforms zeros
    }
    if (kind == 2) {
        fwd264_8(v);
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = as_s16(v[k]);                          // the row results live in an int16 array
        transpose8(v, row);
        fwd264_8(v);
    } else if (fast) {
        fdct_fast8(v); transpose8(v, row); fdct_fast8(v);
    } else {
        fdct_slow8<false, 4, 4>(v); transpose8(v, row); fdct_slow8<true, 4, 4>(v);
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s = kind == 1 ? max(s, abs(v[k])) : s + abs(v[k]);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {                                             // over the block's eight lanes
        const int t = __shfl_xor_sync(0xffffffffu, s, o);
        s = kind == 1 ? max(s, t) : s + t;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 8);                                      // over the blocks (WRAPPER8_16_SQ adds the scores)
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    if (lane == 0) out[i] = s;
}

} // namespace

void mecmp_dct_launch(cudaStream_t st, unsigned ctas, int threads, int fn, int w, int h, const uint8_t *f1, const uint8_t *f2, long long stride,
                      const int64_t *off1, const int64_t *off2, long long n, int32_t *out)
{
    const int kind = fn == B200_MECMP_DCT_SAD ? 0 : fn == B200_MECMP_DCT_MAX ? 1 : 2;
    me_dct_kernel<<<ctas, threads, 0, st>>>(kind, g_dct_fast.load(), w == 8 ? 1 : h == 16 ? 4 : 2, f1, f2, stride, off1, off2, n, out);
}

B200_API int b200_me_cmp_set_dct_algo(int dct_algo)
{
    if (dct_algo == 6) {                                          // FF_DCT_FAAN: the floating-point AAN DCT is not built
        b200_set_error("me_cmp: FF_DCT_FAAN is not implemented");
        return B200_ENOSYS;
    }
    g_dct_fast.store(dct_algo == 1);                              // FF_DCT_FASTINT; every other value selects the islow DCT (fdctdsp.c:31-43)
    return 0;
}
