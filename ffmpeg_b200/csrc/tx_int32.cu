// tx_int32.cu — libavutil/tx 32-bit fixed-point transforms on sm_100a: AV_TX_INT32_FFT and AV_TX_INT32_MDCT, power-of-two lengths
// (the fixed-point AAC / AC-3 decoders' transforms).
//
// Reference semantics reproduced bit for bit (checker: oracle/txi_oracle.c): the TX_INT32 instantiation of libavutil/tx_template.c
// (libavutil/tx_int32.c) with the macros of libavutil/tx_priv.h:113-155 — sums wrap modulo 2^32, products are 64-bit and rounded
// ((a*b - c*d + 2^30) >> 31), tables are clip(llrintf((float)(x * 2^31))), the forward MDCT folds its input with (a + b + 32) >> 6.
//   tables tx_template.c:65-77; butterflies / transform / combine :540-586; base cases :631-704; recursion :615-629; FFT wrapper :763-778;
//   MDCT :1223-1342; twiddles :2107-2134; permutation libavutil/tx.c:125-154.
//
// First version, correctness before speed: ONE THREAD PER TRANSFORM working in global memory (the float kernels of tx.cu, one
// transform per CTA in shared memory, are the model for the tuned version; IMAD.WIDE carries the 64-bit products).
#include "tx_int32.h"
#include <vector>
#include <cmath>
#include <cstring>

namespace {

// [device-code tx_int32] (tests/cuda_emu runs this block on the CPU against the checker; comment markers only)
struct TxiDev {
    const int *map;                 // FFT: gather permutation; MDCT: the same, doubled for the inverse (positions in floats)
    const int2 *exp;                // MDCT twiddles (inverse: pre-shuffled copy first, natural order after it)
    const int *tabs[18];            // tabs[k]: cosine table of the 2^k-point transform
    int n, log2n, len;              // n: complex points of the FFT that runs
};

__device__ __forceinline__ int mulr(long long accu) { return (int)((accu + 0x40000000LL) >> 31); }
// CMUL(dre, dim, are, aim, bre, bim)
__device__ __forceinline__ void cmul(unsigned &dre, unsigned &dim, int are, int aim, int bre, int bim)
{
    dre = (unsigned)mulr((long long)bre * are - (long long)bim * aim);
    dim = (unsigned)mulr((long long)bim * are + (long long)bre * aim);
}
__device__ __forceinline__ void butterflies(int2 &a0, int2 &a1, int2 &a2, int2 &a3, unsigned t1, unsigned t2, unsigned t5, unsigned t6)
{
    const unsigned r0 = a0.x, i0 = a0.y, r1 = a1.x, i1 = a1.y;
    const unsigned t3 = t5 - t1; t5 = t5 + t1;
    a2.x = (int)(r0 - t5); a0.x = (int)(r0 + t5);
    a3.y = (int)(i1 - t3); a1.y = (int)(i1 + t3);
    const unsigned t4 = t2 - t6; t6 = t2 + t6;
    a3.x = (int)(r1 - t4); a1.x = (int)(r1 + t4);
    a2.y = (int)(i0 - t6); a0.y = (int)(i0 + t6);
}
__device__ __forceinline__ void transform(int2 &a0, int2 &a1, int2 &a2, int2 &a3, int wre, int wim)
{
    unsigned t1, t2, t5, t6;
    cmul(t1, t2, a2.x, a2.y, wre, -wim);
    cmul(t5, t6, a3.x, a3.y, wre, wim);
    butterflies(a0, a1, a2, a3, t1, t2, t5, t6);
}
__device__ __forceinline__ void fft4(int2 *d)
{
    const unsigned s0r = d[0].x, s1r = d[1].x, s2r = d[2].x, s3r = d[3].x, s0i = d[0].y, s1i = d[1].y, s2i = d[2].y, s3i = d[3].y;
    const unsigned t3 = s0r - s1r, t1 = s0r + s1r, t8 = s3r - s2r, t6 = s3r + s2r;
    const unsigned t4 = s0i - s1i, t2 = s0i + s1i, t7 = s2i - s3i, t5 = s2i + s3i;
    d[2].x = (int)(t1 - t6); d[0].x = (int)(t1 + t6);
    d[3].y = (int)(t4 - t8); d[1].y = (int)(t4 + t8);
    d[3].x = (int)(t3 - t7); d[1].x = (int)(t3 + t7);
    d[2].y = (int)(t2 - t5); d[0].y = (int)(t2 + t5);
}

// in-place 2^K-point transform without the input permutation; the recursion is unrolled at compile time into calls
template <int K> __device__ void fft_ns(int2 *d, const TxiDev &P)
{
    constexpr int n = 1 << K, n4 = n / 4;
    fft_ns<K - 1>(d, P);
    fft_ns<K - 2>(d + 2 * n4, P);
    fft_ns<K - 2>(d + 3 * n4, P);
    constexpr int len = n4 >> 1, o1 = 2 * len, o2 = 4 * len, o3 = 6 * len;
    const int *cs = P.tabs[K], *wim = cs + o1 - 7;
    int2 *z = d;
    for (int i = 0; i < len; i += 4) {
        transform(z[0], z[o1 + 0], z[o2 + 0], z[o3 + 0], cs[0], wim[7]);
        transform(z[2], z[o1 + 2], z[o2 + 2], z[o3 + 2], cs[2], wim[5]);
        transform(z[4], z[o1 + 4], z[o2 + 4], z[o3 + 4], cs[4], wim[3]);
        transform(z[6], z[o1 + 6], z[o2 + 6], z[o3 + 6], cs[6], wim[1]);
        transform(z[1], z[o1 + 1], z[o2 + 1], z[o3 + 1], cs[1], wim[6]);
        transform(z[3], z[o1 + 3], z[o2 + 3], z[o3 + 3], cs[3], wim[4]);
        transform(z[5], z[o1 + 5], z[o2 + 5], z[o3 + 5], cs[5], wim[2]);
        transform(z[7], z[o1 + 7], z[o2 + 7], z[o3 + 7], cs[7], wim[0]);
        z += 8; cs += 8; wim -= 8;
    }
}
template <> __device__ void fft_ns<0>(int2 *, const TxiDev &) {}
template <> __device__ void fft_ns<1>(int2 *d, const TxiDev &)
{
    const unsigned re = (unsigned)d[0].x - (unsigned)d[1].x, im = (unsigned)d[0].y - (unsigned)d[1].y;
    d[0].x = (int)((unsigned)d[0].x + (unsigned)d[1].x); d[0].y = (int)((unsigned)d[0].y + (unsigned)d[1].y);
    d[1].x = (int)re; d[1].y = (int)im;
}
template <> __device__ void fft_ns<2>(int2 *d, const TxiDev &) { fft4(d); }
template <> __device__ void fft_ns<3>(int2 *d, const TxiDev &P)
{
    const int c = P.tabs[3][1];
    const int2 s4 = d[4], s5 = d[5], s6 = d[6], s7 = d[7];
    fft4(d);
    const unsigned t1 = (unsigned)s4.x - (unsigned)(-s5.x), t2 = (unsigned)s4.y - (unsigned)(-s5.y);
    const unsigned t5 = (unsigned)s6.x - (unsigned)(-s7.x), t6 = (unsigned)s6.y - (unsigned)(-s7.y);
    d[5].x = (int)((unsigned)s4.x + (unsigned)(-s5.x)); d[5].y = (int)((unsigned)s4.y + (unsigned)(-s5.y));
    d[7].x = (int)((unsigned)s6.x + (unsigned)(-s7.x)); d[7].y = (int)((unsigned)s6.y + (unsigned)(-s7.y));
    butterflies(d[0], d[2], d[4], d[6], t1, t2, t5, t6);
    transform(d[1], d[3], d[5], d[7], c, c);
}
template <> __device__ void fft_ns<4>(int2 *d, const TxiDev &P)
{
    const int *c = P.tabs[4];
    fft_ns<3>(d, P);
    fft4(d + 8);
    fft4(d + 12);
    butterflies(d[0], d[4], d[8], d[12], (unsigned)d[8].x, (unsigned)d[8].y, (unsigned)d[12].x, (unsigned)d[12].y);
    transform(d[2], d[6], d[10], d[14], c[2], c[2]);
    transform(d[1], d[5], d[9], d[13], c[1], c[3]);
    transform(d[3], d[7], d[11], d[15], c[3], c[1]);
}
__device__ void fft_ns_any(int2 *d, const TxiDev &P)
{
    switch (P.log2n) {
    case 0: break;
    case 1: fft_ns<1>(d, P); break;   case 2: fft_ns<2>(d, P); break;   case 3: fft_ns<3>(d, P); break;
    case 4: fft_ns<4>(d, P); break;   case 5: fft_ns<5>(d, P); break;   case 6: fft_ns<6>(d, P); break;
    case 7: fft_ns<7>(d, P); break;   case 8: fft_ns<8>(d, P); break;   case 9: fft_ns<9>(d, P); break;
    case 10: fft_ns<10>(d, P); break; case 11: fft_ns<11>(d, P); break; default: fft_ns<12>(d, P); break;
    }
}

__device__ __forceinline__ int fold(int x, int y) { return (int)((unsigned)x + (unsigned)y + 32u) >> 6; }

// kind 0: FFT (out-of-place gather + in-place transform), 1: inverse MDCT, 2: forward MDCT.  One thread per transform.
template <int KIND>
__global__ void __launch_bounds__(64)
tx_i32_kernel(const TxiDev P, int *out, const int *in, long long stride, long long out_step, long long in_step, long long count, int2 *scratch)
{
    const long long tr = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tr >= count) return;
    int *dst = reinterpret_cast<int *>(reinterpret_cast<char *>(out) + tr * out_step);
    const int *src = reinterpret_cast<const int *>(reinterpret_cast<const char *>(in) + tr * in_step);
    if (KIND == 0) {
        int2 *d = reinterpret_cast<int2 *>(dst);
        const int2 *s = reinterpret_cast<const int2 *>(src);
        for (int j = 0; j < P.n; j++) d[j] = s[P.map[j]];
        fft_ns_any(d, P);
        return;
    }
    const int len2 = P.len >> 1, len4 = P.len >> 2;
    const int2 *e = P.exp;
    if (KIND == 1) {
        int2 *z = reinterpret_cast<int2 *>(dst);
        const int *in1 = src, *in2 = src + (len2 * 2 - 1) * stride;
        for (int i = 0; i < len2; i++) {
            const int k = P.map[i];
            unsigned re, im;
            cmul(re, im, in2[-k * stride], in1[k * stride], e[i].x, e[i].y);
            z[i].x = (int)re; z[i].y = (int)im;
        }
        fft_ns_any(z, P);
        e += len2;
        for (int i = 0; i < len4; i++) {
            const int i0 = len4 + i, i1 = len4 - i - 1;
            const int2 s1 = make_int2(z[i1].y, z[i1].x), s0 = make_int2(z[i0].y, z[i0].x);
            unsigned a, b;
            cmul(a, b, s1.x, s1.y, e[i1].y, e[i1].x); z[i1].x = (int)a; z[i0].y = (int)b;
            cmul(a, b, s0.x, s0.y, e[i0].y, e[i0].x); z[i0].x = (int)a; z[i1].y = (int)b;
        }
    } else {
        int2 *z = scratch + tr * len2;
        const int len3 = len2 * 3;
        for (int i = 0; i < len2; i++) {
            const int k = 2 * i, idx = P.map[i];
            int re, im;
            if (k < len2) { re = fold(-src[len2 + k], src[1 * len2 - 1 - k]); im = fold(-src[len3 + k], -src[1 * len3 - 1 - k]); }
            else          { re = fold(-src[len2 + k], -src[5 * len2 - 1 - k]); im = fold(src[-len2 + k], -src[1 * len3 - 1 - k]); }
            unsigned a, b;
            cmul(a, b, re, im, e[i].x, e[i].y);
            z[idx].y = (int)a; z[idx].x = (int)b;
        }
        fft_ns_any(z, P);
        for (int i = 0; i < len4; i++) {
            const int i0 = len4 + i, i1 = len4 - i - 1;
            const int2 s1 = z[i1], s0 = z[i0];
            unsigned a, b;
            cmul(a, b, s0.x, s0.y, e[i0].y, e[i0].x); dst[(2 * i1 + 1) * stride] = (int)a; dst[2 * i0 * stride] = (int)b;
            cmul(a, b, s1.x, s1.y, e[i1].y, e[i1].x); dst[(2 * i0 + 1) * stride] = (int)a; dst[2 * i1 * stride] = (int)b;
        }
    }
}
// [/device-code tx_int32]

int sr_perm(int i, int len, int inv)
{
    len >>= 1;
    if (len <= 1) return i & 1;
    if (!(i & len)) return sr_perm(i, len, inv) * 2;
    len >>= 1;
    return sr_perm(i, len, inv) * 4 + 1 - 2 * (!(i & len) ^ inv);
}
int32_t rescale(double x)                                           // RESCALE of tx_priv.h:140 for TX_INT32
{
    const float f = (float)(x * 2147483648.0);
    long long v = llrintf(f);
    if (v < INT32_MIN) v = INT32_MIN;
    if (v > INT32_MAX) v = INT32_MAX;
    return (int32_t)v;
}

} // namespace

struct TxI32 {
    B200Device *dev = nullptr;
    int type = 0, inv = 0, len = 0;
    TxiDev d{};
    void *blob = nullptr;
    void *scratch = nullptr; size_t scratch_bytes = 0;
};

bool tx_i32_length_ok(int type, int len)
{
    if (len < 2 || (len & (len - 1))) return false;
    const int n = type == B200_TX_INT32_FFT ? len : len >> 1;
    return n >= 1 && n <= 4096 && (type == B200_TX_INT32_FFT || len >= 4);
}

// host tables, flattened: [map n][exp 2 * (inv ? 2n : n) (MDCT only)][cos tables k = 3 .. log2 n]; layout4 = word offsets + log2 n
static int i32_host_tables(std::vector<int32_t> &w, int32_t lay[4], int type, int inv, int len, float scale)
{
    const int n = type == B200_TX_INT32_FFT ? len : len >> 1;
    int k = 0;
    while ((1 << k) < n) k++;
    std::vector<int> map(n);
    const bool scatter = type == B200_TX_INT32_MDCT && !inv;
    for (int i = 0; i < n; i++) {
        const int p = n == 1 ? 0 : (-sr_perm(i, n, inv)) & (n - 1);
        if (scatter) map[p] = i; else map[i] = p;
    }
    w.clear();
    lay[0] = 0;
    lay[1] = n;
    std::vector<int32_t> exp;
    if (type == B200_TX_INT32_MDCT) {
        const int len4 = len >> 1;
        const double theta = (scale < 0 ? len4 : 0) + 1.0 / 8.0, sc = sqrt(fabs((double)scale));
        std::vector<int32_t> full(2 * (size_t)len4);
        for (int i = 0; i < len4; i++) {
            const double alpha = M_PI_2 * (i + theta) / len4;
            full[2 * i] = rescale(cos(alpha) * sc);
            full[2 * i + 1] = rescale(sin(alpha) * sc);
        }
        if (inv) {
            exp.assign(4 * (size_t)len4, 0);
            memcpy(&exp[2 * (size_t)len4], full.data(), sizeof(int32_t) * 2 * len4);
            for (int i = 0; i < len4; i++) { exp[2 * i] = full[2 * map[i]]; exp[2 * i + 1] = full[2 * map[i] + 1]; }
            for (int i = 0; i < len4; i++) map[i] <<= 1;            // "saves multiplies in loops", tx_template.c:1266-1268
        } else
            exp = full;
    }
    for (int x : map) w.push_back(x);
    for (int32_t x : exp) w.push_back(x);
    lay[2] = (int32_t)w.size();
    for (int kk = 3; kk <= k; kk++) {
        const int nn = 1 << kk;
        const double freq = 2 * M_PI / nn;
        for (int i = 0; i < nn / 4; i++) w.push_back(rescale(cos(i * freq)));
        w.push_back(0);
    }
    lay[3] = k;
    return (int)w.size();
}

B200_API int b200_tx_i32_tables(int type, int inv, int len, float scale, int32_t *words, int cap, int32_t *layout4)
{
    if ((type != B200_TX_INT32_FFT && type != B200_TX_INT32_MDCT) || !tx_i32_length_ok(type, len)) return B200_ENOSYS;
    std::vector<int32_t> w;
    int32_t lay[4];
    const int nw = i32_host_tables(w, lay, type, !!inv, len, scale);
    if (layout4) memcpy(layout4, lay, sizeof(lay));
    if (words && cap >= nw) memcpy(words, w.data(), (size_t)nw * 4);
    return nw;
}

TxI32 *tx_i32_create(B200Device *dev, int type, int inv, int len, float scale)
{
    if (!tx_i32_length_ok(type, len)) return nullptr;
    std::vector<int32_t> w;
    int32_t lay[4];
    const int nw = i32_host_tables(w, lay, type, inv, len, scale);
    TxI32 *p = new (std::nothrow) TxI32();
    if (!p) return nullptr;
    p->dev = dev; p->type = type; p->inv = inv; p->len = len;
    if (cudaMalloc(&p->blob, (size_t)nw * 4 + 16) != cudaSuccess || cudaMemcpy(p->blob, w.data(), (size_t)nw * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        b200_set_error("tx_i32_create: device tables");
        if (p->blob) cudaFree(p->blob);
        delete p;
        return nullptr;
    }
    const int32_t *b = (const int32_t *)p->blob;
    TxiDev &d = p->d;
    d.n = type == B200_TX_INT32_FFT ? len : len >> 1; d.log2n = lay[3]; d.len = len;
    d.map = b + lay[0];
    d.exp = (const int2 *)(b + lay[1]);
    const int32_t *c = b + lay[2];
    for (int k = 0; k < 18; k++) d.tabs[k] = nullptr;
    for (int k = 3; k <= d.log2n; k++) { d.tabs[k] = c; c += (1 << k) / 4 + 1; }
    return p;
}

void tx_i32_free(TxI32 *p)
{
    if (!p) return;
    if (p->blob) cudaFree(p->blob);
    if (p->scratch) cudaFree(p->scratch);
    delete p;
}

int tx_i32_launch(TxI32 *p, cudaStream_t st, void *out, const void *in, ptrdiff_t stride, int64_t count, ptrdiff_t out_step, ptrdiff_t in_step)
{
    if (count <= 0) return 0;
    // the kernels move complex values as 8-byte words: the complex side(s) of a transform must be 8-byte aligned (the reference
    // asks for 32 unless AV_TX_UNALIGNED), the real side 4
    const uintptr_t oa = reinterpret_cast<uintptr_t>(out) | (uintptr_t)out_step, ia = reinterpret_cast<uintptr_t>(in) | (uintptr_t)in_step;
    const bool cplx_out = p->type == B200_TX_INT32_FFT || p->inv, cplx_in = p->type == B200_TX_INT32_FFT;
    if ((oa & (cplx_out ? 7 : 3)) || (ia & (cplx_in ? 7 : 3))) return B200_EINVAL;
    const bool fwd_mdct = p->type == B200_TX_INT32_MDCT && !p->inv;
    const size_t l2 = (size_t)p->len >> 1;
    int64_t chunk = count;
    if (fwd_mdct) {                                                  // the forward MDCT builds its complex array in scratch
        const int64_t chunk_max = (int64_t)((size_t)(256u << 20) / (l2 * sizeof(int2)));
        if (chunk > chunk_max) chunk = chunk_max;
        const size_t need = (size_t)chunk * l2 * sizeof(int2);
        if (p->scratch_bytes < need) {
            if (p->scratch) { cudaStreamSynchronize(st); cudaFree(p->scratch); p->scratch = nullptr; p->scratch_bytes = 0; }
            B200_CUDA_OK(cudaMalloc(&p->scratch, need));
            p->scratch_bytes = need;
        }
    }
    if (chunk > 0x7fffffffLL / 2) chunk = 0x7fffffffLL / 2;
    for (int64_t c0 = 0; c0 < count; c0 += chunk) {
        const long long cnt = count - c0 < chunk ? count - c0 : chunk;
        const unsigned nb = (unsigned)((cnt + 63) / 64);
        int *o = (int *)((char *)out + c0 * out_step);
        const int *i = (const int *)((const char *)in + c0 * in_step);
        const long long sf = (long long)(stride / 4);
        if (p->type == B200_TX_INT32_FFT) tx_i32_kernel<0><<<nb, 64, 0, st>>>(p->d, o, i, sf, out_step, in_step, cnt, nullptr);
        else if (p->inv)                  tx_i32_kernel<1><<<nb, 64, 0, st>>>(p->d, o, i, sf, out_step, in_step, cnt, nullptr);
        else                              tx_i32_kernel<2><<<nb, 64, 0, st>>>(p->d, o, i, sf, out_step, in_step, cnt, (int2 *)p->scratch);
        B200_LAUNCHED();
    }
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

void tx_i32_host_fn(TxI32 *p, void *out, void *in, ptrdiff_t stride)
{
    auto fail = [](const char *what) { fprintf(stderr, "libb200dsp: av_tx_fn (int32) failed: %s (%s)\n", what, b200_last_error()); abort(); };
    B200Device *d = p->dev;
    if (cudaSetDevice(d->ordinal) != cudaSuccess) fail("cudaSetDevice");
    const size_t len = p->len;
    const bool mdct = p->type == B200_TX_INT32_MDCT;
    const size_t in_elems = !mdct ? 2 * len : p->inv ? len : 2 * len, out_elems = !mdct ? 2 * len : len;        // 32-bit words
    B200_LOCK_DEVICE(d);      // scratch + stream are per device: one host-pointer call at a time (released on return)
    int32_t *scr = (int32_t *)b200_scratch(d, (in_elems + out_elems) * 4 + 512);
    if (!scr) fail("scratch");
    int32_t *din = scr, *dout = scr + ((in_elems + 63) & ~(size_t)63);
    cudaStream_t st = d->stream;
    cudaError_t e;
    const bool strided_in = mdct && p->inv && stride != 4, strided_out = mdct && !p->inv && stride != 4;
    if (strided_in) e = cudaMemcpy2DAsync(din, 4, in, (size_t)stride, 4, in_elems, cudaMemcpyHostToDevice, st);
    else e = cudaMemcpyAsync(din, in, in_elems * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) fail("h2d");
    if (tx_i32_launch(p, st, dout, din, 4, 1, 0, 0) < 0) fail("launch");
    if (strided_out) e = cudaMemcpy2DAsync(out, (size_t)stride, dout, 4, 4, out_elems, cudaMemcpyDeviceToHost, st);
    else e = cudaMemcpyAsync(out, dout, out_elems * 4, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) fail("d2h");
}
