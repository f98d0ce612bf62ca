// fake_device.cpp — TEST INFRASTRUCTURE ONLY: the few services of ffmpeg_b200/csrc/device.cu (and one of idct.cu) that the
// PTX-free translation units link against, on top of the stand-in runtime.
#include "common.h"
#include <cstdarg>
#include <atomic>

std::atomic<uint64_t> g_b200_launches{0};
static char g_err[512];
void b200_set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
extern "C" const char *b200_last_error(void) { return g_err; }
extern "C" uint64_t emu_launch_count(void) { return g_b200_launches.load(); }
static B200Device g_dev;
B200Device *b200_default_device() { return &g_dev; }
extern "C" int b200_device_open(B200Device **out, int, void *) { *out = &g_dev; return 0; }        // one stand-in device
void *b200_scratch(B200Device *d, size_t bytes)
{
    if (d->scratch_bytes < bytes) { free(d->scratch); d->scratch = malloc(bytes + 64); d->scratch_bytes = bytes; }
    return d->scratch;
}
void *b200_pinned(B200Device *d, size_t bytes)
{
    if (d->pinned_bytes < bytes) { free(d->pinned); d->pinned = malloc(bytes + 64); d->pinned_bytes = bytes; }
    return d->pinned;
}
// idct_hbd.cu starts from the 8-bit table (idct.cu, full of inline PTX, is not part of this build): hand back an empty one
extern "C" int b200_idctdsp_init(B200IDCTDSPContext *c, int, int, int) { memset(c, 0, sizeof(*c)); return 0; }

// tx_pfa.cu is reached through tx.cu in the library; here directly
#include "tx_pfa.h"
extern "C" int emu_host_tx_pfa(int inv, int len, float scale, void *out, const void *in, long long stride, long long count, long long out_step, long long in_step)
{
    TxPfa *p = tx_pfa_create(inv, len, scale);
    if (!p) return -1;
    const int r = tx_pfa_launch(p, nullptr, out, in, (ptrdiff_t)stride, count, (ptrdiff_t)out_step, (ptrdiff_t)in_step);
    tx_pfa_free(p);
    return r;
}
