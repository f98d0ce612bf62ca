// fake_mecmp_dct.cpp — TEST INFRASTRUCTURE ONLY: mecmp_dct.cu's launcher is called from b200_me_cmp_batch_device() / the table entries in
// mecmp.cu (not part of the emulated build); this is the batch call with the same geometry (8 comparisons per CTA of 256 threads).
#include "common.h"
#include "mecmp_dct.h"
extern "C" int emu_host_me_dct(int fn, int w, int h, const uint8_t *f1, const uint8_t *f2, long long stride, const int64_t *off1, const int64_t *off2,
                               long long n, int32_t *out)
{
    mecmp_dct_launch(nullptr, (unsigned)((n + 7) / 8), 256, fn, w, h, f1, f2, stride, off1, off2, n, out);
    return 0;
}
