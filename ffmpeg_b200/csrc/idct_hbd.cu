// idct_hbd.cu — libavcodec's "simple" IDCT at 10 and 12 bit (int16 coefficients, uint16 pixels) on sm_100a.
//
// Reference semantics reproduced bit-for-bit (checker: the 10 / 12 bit part of oracle/idct_oracle.c), libavcodec/simple_idct_template.c:
//   :63-104   constants (10 bit: W3 = 19265, W4 = 16384, ROW_SHIFT 12, COL_SHIFT 19, DC_SHIFT 2; 12 bit: W1..W7 = 45451 ... 9041,
//             ROW_SHIFT 16, COL_SHIFT 17, DC_SHIFT -1)
//   :114-206  idctRowCondDC: rows without AC terms become (row[0] << 2) resp. (row[0] + 1) >> 1, & 0xffff; results back in int16
//   :209-327  IDCT_COLS, idctSparseCol{,Put,Add}; :329-368 ff_simple_idct_{put,add,}_int16_{10,12}bit
// which ff_idctdsp_init installs for bits_per_raw_sample 9, 10 (the 10-bit set) and 12 (idctdsp.c:248-266).  Sums mod 2^32.
//
// First version, correctness before speed: one thread per 8x8 block, the block held in registers (eight 16-byte loads), row
// pass and column pass fully unrolled.  HBM-bound by bytes (128 B in + 128 B out per block); the 8-bit kernel's staging through
// shared memory (idct.cu) is the model for the tuned version.
#include "common.h"
#include <cstring>

namespace {

// [device-code idct_hbd] (tests/cuda_emu runs this block on the CPU against the checker; comment markers only)
template <int DEPTH> struct K;
template <> struct K<10> { static constexpr unsigned W1 = 22725, W2 = 21407, W3 = 19265, W4 = 16384, W5 = 12873, W6 = 8867, W7 = 4520;
                           static constexpr int ROW = 12, COL = 19, DC = 2; };
template <> struct K<12> { static constexpr unsigned W1 = 45451, W2 = 42813, W3 = 38531, W4 = 32767, W5 = 25746, W6 = 17734, W7 = 9041;
                           static constexpr int ROW = 16, COL = 17, DC = -1; };
// the EXTRA_SHIFT instantiation proresdsp.c makes of the 10-bit constants (simple_idct_template.c:73-76: ROW_SHIFT 13, COL_SHIFT 18,
// DC_SHIFT 1) with the two extra bits its row pass is called with (proresdsp.c:61-62) folded in: ROW 13 + 2, DC 1 - 2
template <> struct K<110> { static constexpr unsigned W1 = 22725, W2 = 21407, W3 = 19265, W4 = 16384, W5 = 12873, W6 = 8867, W7 = 4520;
                            static constexpr int ROW = 15, COL = 18, DC = -1; };

template <int DEPTH> __device__ __forceinline__ void row_pass(int *r)     // r[0..7]: one row, values are int16 in int registers
{
    using C = K<DEPTH>;
    if (!(r[1] | r[2] | r[3] | r[4] | r[5] | r[6] | r[7])) {
        const int t = C::DC >= 0 ? r[0] * (1 << (C::DC >= 0 ? C::DC : 0)) : (r[0] + (1 << (C::DC < 0 ? -C::DC - 1 : 0))) >> (C::DC < 0 ? -C::DC : 0);
        const int dc = (int)(int16_t)(uint16_t)(t & 0xffff);
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = dc;
        return;
    }
    unsigned a0 = C::W4 * (unsigned)r[0] + (1u << (C::ROW - 1)), a1 = a0, a2 = a0, a3 = a0;
    a0 += C::W2 * (unsigned)r[2]; a1 += C::W6 * (unsigned)r[2]; a2 -= C::W6 * (unsigned)r[2]; a3 -= C::W2 * (unsigned)r[2];
    unsigned b0 = C::W1 * (unsigned)r[1] + C::W3 * (unsigned)r[3];
    unsigned b1 = C::W3 * (unsigned)r[1] - C::W7 * (unsigned)r[3];
    unsigned b2 = C::W5 * (unsigned)r[1] - C::W1 * (unsigned)r[3];
    unsigned b3 = C::W7 * (unsigned)r[1] - C::W5 * (unsigned)r[3];
    a0 += C::W4 * (unsigned)r[4] + C::W6 * (unsigned)r[6];
    a1 -= C::W4 * (unsigned)r[4] + C::W2 * (unsigned)r[6];
    a2 += C::W2 * (unsigned)r[6] - C::W4 * (unsigned)r[4];
    a3 += C::W4 * (unsigned)r[4] - C::W6 * (unsigned)r[6];
    b0 += C::W5 * (unsigned)r[5] + C::W7 * (unsigned)r[7];
    b1 -= C::W1 * (unsigned)r[5] + C::W5 * (unsigned)r[7];
    b2 += C::W7 * (unsigned)r[5] + C::W3 * (unsigned)r[7];
    b3 += C::W3 * (unsigned)r[5] - C::W1 * (unsigned)r[7];
    r[0] = (int)(int16_t)((int)(a0 + b0) >> C::ROW); r[7] = (int)(int16_t)((int)(a0 - b0) >> C::ROW);
    r[1] = (int)(int16_t)((int)(a1 + b1) >> C::ROW); r[6] = (int)(int16_t)((int)(a1 - b1) >> C::ROW);
    r[2] = (int)(int16_t)((int)(a2 + b2) >> C::ROW); r[5] = (int)(int16_t)((int)(a2 - b2) >> C::ROW);
    r[3] = (int)(int16_t)((int)(a3 + b3) >> C::ROW); r[4] = (int)(int16_t)((int)(a3 - b3) >> C::ROW);
}

template <int DEPTH> __device__ __forceinline__ void col_pass(const int *c, int *o)   // c[0..7]: one column, top to bottom
{
    using C = K<DEPTH>;
    unsigned a0 = C::W4 * (unsigned)(c[0] + (int)((1u << (C::COL - 1)) / C::W4)), a1 = a0, a2 = a0, a3 = a0;
    a0 += C::W2 * (unsigned)c[2]; a1 += C::W6 * (unsigned)c[2]; a2 -= C::W6 * (unsigned)c[2]; a3 -= C::W2 * (unsigned)c[2];
    unsigned b0 = C::W1 * (unsigned)c[1], b1 = C::W3 * (unsigned)c[1], b2 = C::W5 * (unsigned)c[1], b3 = C::W7 * (unsigned)c[1];
    b0 += C::W3 * (unsigned)c[3]; b1 -= C::W7 * (unsigned)c[3]; b2 -= C::W1 * (unsigned)c[3]; b3 -= C::W5 * (unsigned)c[3];
    a0 += C::W4 * (unsigned)c[4]; a1 -= C::W4 * (unsigned)c[4]; a2 -= C::W4 * (unsigned)c[4]; a3 += C::W4 * (unsigned)c[4];
    b0 += C::W5 * (unsigned)c[5]; b1 -= C::W1 * (unsigned)c[5]; b2 += C::W7 * (unsigned)c[5]; b3 += C::W3 * (unsigned)c[5];
    a0 += C::W6 * (unsigned)c[6]; a1 -= C::W2 * (unsigned)c[6]; a2 += C::W2 * (unsigned)c[6]; a3 -= C::W6 * (unsigned)c[6];
    b0 += C::W7 * (unsigned)c[7]; b1 -= C::W5 * (unsigned)c[7]; b2 += C::W3 * (unsigned)c[7]; b3 -= C::W1 * (unsigned)c[7];
    o[0] = (int)(a0 + b0) >> C::COL; o[1] = (int)(a1 + b1) >> C::COL; o[2] = (int)(a2 + b2) >> C::COL; o[3] = (int)(a3 + b3) >> C::COL;
    o[4] = (int)(a3 - b3) >> C::COL; o[5] = (int)(a2 - b2) >> C::COL; o[6] = (int)(a1 - b1) >> C::COL; o[7] = (int)(a0 - b0) >> C::COL;
}

// KIND 0: in place on the coefficients, 1: put, 2: add.  dest = uint16 pixels; offsets and line sizes in BYTES (even).
template <int DEPTH, int KIND>
__global__ void __launch_bounds__(128)
idct_hbd_kernel(int16_t *blocks, long long n, uint8_t *dest, const int64_t *dest_off, const int32_t *line_size, int uniform_ls)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v[64];
    const uint4 *src = reinterpret_cast<const uint4 *>(blocks + 64 * i);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint4 q = src[r];
        const unsigned w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for (int k = 0; k < 4; k++) { v[8 * r + 2 * k] = (int)(int16_t)(w[k] & 0xffffu); v[8 * r + 2 * k + 1] = (int)(int16_t)(w[k] >> 16); }
    }
#pragma unroll
    for (int r = 0; r < 8; r++) row_pass<DEPTH>(v + 8 * r);
    constexpr int MAXV = (1 << DEPTH) - 1;
    uint16_t *d = nullptr;
    long long ls = 0;
    if (KIND != 0) {
        d = reinterpret_cast<uint16_t *>(dest + __ldg(dest_off + i));
        ls = (line_size ? __ldg(line_size + i) : uniform_ls) / 2;
    }
#pragma unroll
    for (int c = 0; c < 8; c++) {
        int col[8], o[8];
#pragma unroll
        for (int j = 0; j < 8; j++) col[j] = v[8 * j + c];
        col_pass<DEPTH>(col, o);
#pragma unroll
        for (int j = 0; j < 8; j++) v[8 * j + c] = o[j];
    }
    if (KIND != 0) {
        // put / add: whole rows (8 samples = 16 bytes) when the rows are 16-byte aligned; for `add` every destination row is loaded
        // before the first store (a load - add - store per sample would make one trip to memory per sample)
        const bool vec = ((reinterpret_cast<uintptr_t>(d) | (uintptr_t)(ls * 2)) & 15) == 0;
        if (vec) {
            uint4 dv[8];
#pragma unroll
            for (int j = 0; j < 8; j++) dv[j] = KIND == 2 ? *reinterpret_cast<const uint4 *>(d + j * ls) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const unsigned w[4] = { dv[j].x, dv[j].y, dv[j].z, dv[j].w };
                unsigned o2[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int x0 = v[8 * j + 2 * k] + (int)(w[k] & 0xffffu), x1 = v[8 * j + 2 * k + 1] + (int)(w[k] >> 16);
                    o2[k] = (unsigned)min(max(x0, 0), MAXV) | ((unsigned)min(max(x1, 0), MAXV) << 16);
                }
                *reinterpret_cast<uint4 *>(d + j * ls) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) {                          // unaligned rows: a row at a time
                int dd[8];
#pragma unroll
                for (int k = 0; k < 8; k++) dd[k] = KIND == 2 ? (int)d[j * ls + k] : 0;
#pragma unroll
                for (int k = 0; k < 8; k++) d[j * ls + k] = (uint16_t)min(max(dd[k] + v[8 * j + k], 0), MAXV);
            }
        }
    }
    if (KIND == 0) {
        uint4 *out = reinterpret_cast<uint4 *>(blocks + 64 * i);
#pragma unroll
        for (int r = 0; r < 8; r++) {
            uint4 q;
            q.x = ((unsigned)v[8 * r + 0] & 0xffffu) | ((unsigned)v[8 * r + 1] << 16);
            q.y = ((unsigned)v[8 * r + 2] & 0xffffu) | ((unsigned)v[8 * r + 3] << 16);
            q.z = ((unsigned)v[8 * r + 4] & 0xffffu) | ((unsigned)v[8 * r + 5] << 16);
            q.w = ((unsigned)v[8 * r + 6] & 0xffffu) | ((unsigned)v[8 * r + 7] << 16);
            out[r] = q;
        }
    }
}

// ProresDSPContext.idct_put (libavcodec/proresdsp.c:56-82,102-167): coefficients * qmat (int16 wrap), row pass, 8192 added to the first
// row, in-place column pass (int16), pixels clipped to [4, 2^bits - 5].  VARIANT 110: 10 bit, 12: 12 bit.  One thread per block;
// the coefficient blocks are left untouched (the reference overwrites them).
template <int VARIANT>
__global__ void __launch_bounds__(128)
prores_idct_put_kernel(const int16_t *blocks, long long n, const int16_t *qmat, uint8_t *dest, const int64_t *dest_off, const int32_t *line_size,
                       int uniform_ls)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v[64];
#pragma unroll
    for (int k = 0; k < 64; k++) v[k] = (int)(int16_t)((int)blocks[64 * i + k] * (int)qmat[k]);
#pragma unroll
    for (int r = 0; r < 8; r++) row_pass<VARIANT>(v + 8 * r);
    constexpr int BITS = VARIANT == 110 ? 10 : 12, LO = 4, HI = (1 << BITS) - 4 - 1;
    uint16_t *d = reinterpret_cast<uint16_t *>(dest + dest_off[i]);
    const long long ls = (line_size ? line_size[i] : uniform_ls) >> 1;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        int col[8], o[8];
        v[c] = (int)(int16_t)(v[c] + 8192);
#pragma unroll
        for (int j = 0; j < 8; j++) col[j] = v[8 * j + c];
        col_pass<VARIANT>(col, o);
#pragma unroll
        for (int j = 0; j < 8; j++) d[j * ls + c] = (uint16_t)min(max((int)(int16_t)o[j], LO), HI);
    }
}
// [/device-code idct_hbd]

template <int DEPTH>
int launch(cudaStream_t st, int kind, int16_t *blocks, long long n, uint8_t *dest, const int64_t *off, const int32_t *ls, int uls)
{
    const long long grid = (n + 127) / 128;
    if (grid > 0x7fffffffLL) return B200_EINVAL;
    dim3 g((unsigned)grid), t(128);
    if (kind == B200_IDCT)          idct_hbd_kernel<DEPTH, 0><<<g, t, 0, st>>>(blocks, n, dest, off, ls, uls);
    else if (kind == B200_IDCT_PUT) idct_hbd_kernel<DEPTH, 1><<<g, t, 0, st>>>(blocks, n, dest, off, ls, uls);
    else                            idct_hbd_kernel<DEPTH, 2><<<g, t, 0, st>>>(blocks, n, dest, off, ls, uls);
    B200_LAUNCHED();
    return 0;
}

void die(const char *what)
{
    fprintf(stderr, "libb200dsp: high-bit-depth idct failed: %s (%s)\n", what, b200_last_error());
    abort();
}

// drop-in: one block through the device (host pointers); put / add leave the caller's coefficient block as it was (the
// reference uses it as scratch; no caller reads it afterwards)
template <int DEPTH>
void host_op(int kind, uint8_t *dest, ptrdiff_t line_size, int16_t *block)
{
    B200Device *dev = b200_default_device();
    if (!dev) die("no device");
    if (kind != B200_IDCT && ((line_size > -16 && line_size < 16) || (line_size & 1))) die("|line_size| must be even and >= 16 bytes");
    if (cudaSetDevice(dev->ordinal) != cudaSuccess) die("cudaSetDevice");
    B200_LOCK_DEVICE(dev);      // scratch + stream are per device: one host-pointer call at a time (released on return)
    uint8_t *scr = (uint8_t *)b200_scratch(dev, 128 + 8 * 16 + 16);
    if (!scr) die("scratch");
    int16_t *dblk = (int16_t *)scr;
    uint8_t *ddst = scr + 128;
    int64_t *doff = (int64_t *)(scr + 128 + 128);
    cudaStream_t st = dev->stream;
    const int64_t zero = 0;
    if (cudaMemcpyAsync(dblk, block, 128, cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d block");
    if (kind != B200_IDCT) {
        if (kind == B200_IDCT_ADD && b200_h2d_rows(ddst, 16, dest, line_size, 16, 8, st) != cudaSuccess) die("h2d dest");
        if (cudaMemcpyAsync(doff, &zero, 8, cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d off");
    }
    if (launch<DEPTH>(st, kind, dblk, 1, ddst, doff, nullptr, 16) < 0 || cudaGetLastError() != cudaSuccess) die("launch");
    if (kind == B200_IDCT) { if (cudaMemcpyAsync(block, dblk, 128, cudaMemcpyDeviceToHost, st) != cudaSuccess) die("d2h block"); }
    else if (b200_d2h_rows(dest, line_size, ddst, 16, 16, 8, st) != cudaSuccess) die("d2h dest");
    if (cudaStreamSynchronize(st) != cudaSuccess) die("sync");
}

template <int DEPTH> void tab_idct(int16_t *b) { host_op<DEPTH>(B200_IDCT, nullptr, 0, b); }
template <int DEPTH> void tab_put(uint8_t *d, ptrdiff_t ls, int16_t *b) { host_op<DEPTH>(B200_IDCT_PUT, d, ls, b); }
template <int DEPTH> void tab_add(uint8_t *d, ptrdiff_t ls, int16_t *b) { host_op<DEPTH>(B200_IDCT_ADD, d, ls, b); }

} // namespace

B200_API int b200_idctdsp_init_hbd(B200IDCTDSPContext *c, int idct_algo, int bits_per_raw_sample, int lowres)
{
    if (!c) return B200_EINVAL;
    // ff_idctdsp_init, libavcodec/idctdsp.c:248-266: 9 and 10 bit -> the int16 10-bit functions, 12 bit -> the 12-bit ones
    const int depth = bits_per_raw_sample == 9 ? 10 : bits_per_raw_sample;
    if (lowres != 0 || (depth != 10 && depth != 12) || (idct_algo != 0 && idct_algo != 2)) return B200_ENOSYS;
    int ret = b200_idctdsp_init(c, idct_algo, 8, 0);          // the three clamp helpers stay the 8-bit ones, as in the reference
    if (ret < 0) return ret;
    if (depth == 10) { c->idct = tab_idct<10>; c->idct_put = tab_put<10>; c->idct_add = tab_add<10>; }
    else             { c->idct = tab_idct<12>; c->idct_put = tab_put<12>; c->idct_add = tab_add<12>; }
    return 0;
}

B200_API int b200_idct_hbd_batch_device(B200Device *dev, int depth, int kind, int16_t *blocks, int64_t nblocks, uint8_t *dest,
                                        const int64_t *dest_off, const int32_t *line_size, int uniform_line_size)
{
    if (!dev) dev = b200_default_device();
    if (!dev) return B200_ENODEV;
    if (depth == 9) depth = 10;
    if ((depth != 10 && depth != 12) || kind < B200_IDCT || kind > B200_IDCT_ADD || nblocks < 0) return B200_EINVAL;
    if (nblocks == 0) return 0;
    if (!blocks || (reinterpret_cast<uintptr_t>(blocks) & 15)) return B200_EINVAL;
    if (kind != B200_IDCT && (!dest || !dest_off || (reinterpret_cast<uintptr_t>(dest) & 1) || (!line_size && (uniform_line_size & 1))))
        return B200_EINVAL;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    const int ret = depth == 10 ? launch<10>(dev->stream, kind, blocks, nblocks, dest, dest_off, line_size, uniform_line_size)
                                : launch<12>(dev->stream, kind, blocks, nblocks, dest, dest_off, line_size, uniform_line_size);
    if (ret < 0) return ret;
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- ProresDSPContext (libavcodec/proresdsp.h:28-35, ff_proresdsp_init proresdsp.c:169-194)
namespace {

int launch_prores(cudaStream_t st, int bits, const int16_t *blocks, long long n, const int16_t *qmat, uint8_t *dest, const int64_t *off,
                  const int32_t *ls, int uls)
{
    const long long grid = (n + 127) / 128;
    if (grid > 0x7fffffffLL) return B200_EINVAL;
    if (bits == 10) prores_idct_put_kernel<110><<<(unsigned)grid, 128, 0, st>>>(blocks, n, qmat, dest, off, ls, uls);
    else            prores_idct_put_kernel<12><<<(unsigned)grid, 128, 0, st>>>(blocks, n, qmat, dest, off, ls, uls);
    B200_LAUNCHED();
    return 0;
}

template <int BITS>
void prores_tab_put(uint16_t *out, ptrdiff_t linesize, int16_t *block, const int16_t *qmat)
{
    B200Device *dev = b200_default_device();
    if (!dev) die("no device");
    if ((linesize > -16 && linesize < 16) || (linesize & 1)) die("|line size| must be even and >= 16 bytes");
    if (cudaSetDevice(dev->ordinal) != cudaSuccess) die("cudaSetDevice");
    B200_LOCK_DEVICE(dev);      // scratch + stream are per device: one host-pointer call at a time (released on return)
    uint8_t *scr = (uint8_t *)b200_scratch(dev, 128 + 128 + 128 + 16);
    if (!scr) die("scratch");
    int16_t *dblk = (int16_t *)scr, *dq = (int16_t *)(scr + 128);
    uint8_t *ddst = scr + 256;
    int64_t *doff = (int64_t *)(scr + 384);
    cudaStream_t st = dev->stream;
    const int64_t zero = 0;
    if (cudaMemcpyAsync(dblk, block, 128, cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d block");
    if (cudaMemcpyAsync(dq, qmat, 128, cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d qmat");
    if (cudaMemcpyAsync(doff, &zero, 8, cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d off");
    if (launch_prores(st, BITS, dblk, 1, dq, ddst, doff, nullptr, 16) < 0 || cudaGetLastError() != cudaSuccess) die("launch");
    if (b200_d2h_rows(out, linesize, ddst, 16, 16, 8, st) != cudaSuccess) die("d2h");
    if (cudaStreamSynchronize(st) != cudaSuccess) die("sync");
}

} // namespace

B200_API int b200_proresdsp_init(B200ProresDSPContext *c, int bits_per_raw_sample)
{
    if (!c) return B200_EINVAL;
    if (bits_per_raw_sample != 10 && bits_per_raw_sample != 12) return B200_ENOSYS;
    if (!b200_default_device()) return B200_ENODEV;
    memset(c, 0, sizeof(*c));
    c->idct_permutation_type = 0;                                    // FF_IDCT_PERM_NONE
    for (int i = 0; i < 64; i++) c->idct_permutation[i] = (uint8_t)i;
    c->idct_put = bits_per_raw_sample == 10 ? prores_tab_put<10> : prores_tab_put<12>;
    c->idct_put_bayer = nullptr;                                     // ProRes RAW (32-bit coefficients + linearisation curve): not built
    return 0;
}

B200_API int b200_prores_idct_put_batch_device(B200Device *dev, int bits, const int16_t *blocks, int64_t nblocks, const int16_t *qmat,
                                               uint8_t *dest, const int64_t *dest_off, const int32_t *line_size, int uniform_line_size)
{
    if (!dev) dev = b200_default_device();
    if (!dev) return B200_ENODEV;
    if ((bits != 10 && bits != 12) || nblocks < 0) return B200_EINVAL;
    if (nblocks == 0) return 0;
    if (!blocks || !qmat || !dest || !dest_off || (reinterpret_cast<uintptr_t>(dest) & 1) || (!line_size && (uniform_line_size & 1))) return B200_EINVAL;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    const int ret = launch_prores(dev->stream, bits, blocks, nblocks, qmat, dest, dest_off, line_size, uniform_line_size);
    if (ret < 0) return ret;
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}
