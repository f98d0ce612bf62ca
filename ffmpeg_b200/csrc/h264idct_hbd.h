// h264idct_hbd.h — H.264 residual adds for 9 / 10 / 12 / 14 bit samples (h264idct_hbd.cu), installed by b200_h264_idct_init() in h264idct.cu
#pragma once
#include "common.h"

// fills the four members with the functions of that depth; false when the depth has none (ff_h264dsp_init knows 8, 9, 10, 12, 14)
bool h264idct_hbd_fill(B200H264IDCTContext *c, int bit_depth);
