// mecmp_dct.h — transform-domain block comparisons dct_sad / dct_max / dct264_sad (mecmp_dct.cu), dispatched from mecmp.cu
#pragma once
#include "common.h"

// fn = B200_MECMP_DCT_SAD / _DCT_MAX / _DCT264_SAD, w = 16 or 8; n comparisons, one warp each (threads = 32 * warps per CTA)
void mecmp_dct_launch(cudaStream_t st, unsigned ctas, int threads, int fn, int w, int h, const uint8_t *f1, const uint8_t *f2, long long stride,
                      const int64_t *off1, const int64_t *off2, long long n, int32_t *out);
