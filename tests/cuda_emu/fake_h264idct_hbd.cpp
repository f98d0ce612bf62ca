// fake_h264idct_hbd.cpp — TEST INFRASTRUCTURE ONLY: h264idct_hbd.cu's functions are installed by b200_h264_idct_init() in h264idct.cu (not part
// of the emulated build); this is the same call with the member lookup spelled out.
#include "common.h"
#include "h264idct_hbd.h"
extern "C" int emu_host_h264_idct_hbd(int depth, int kind, uint8_t *dst, int32_t *block, long long stride)
{
    B200H264IDCTContext c;
    memset(&c, 0, sizeof(c));
    if (!h264idct_hbd_fill(&c, depth)) return -38;
    b200_h264_idct_fn f = kind == 0 ? c.idct_add : kind == 1 ? c.idct8_add : kind == 2 ? c.idct_dc_add : c.idct8_dc_add;
    f(dst, (int16_t *)block, (ptrdiff_t)stride);
    return 0;
}
