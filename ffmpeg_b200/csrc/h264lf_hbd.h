// h264lf_hbd.h — H.264 deblocking for 9 / 10 / 12 / 14 bit samples (h264lf_hbd.cu), installed by b200_h264_loop_filter_init() in h264lf.cu
#pragma once
#include "common.h"

// fills the twelve members with the functions of that depth; false when the depth has none (ff_h264dsp_init knows 8, 9, 10, 12, 14)
bool h264lf_hbd_fill(B200H264LoopFilterContext *c, int bit_depth, int chroma_format_idc);
