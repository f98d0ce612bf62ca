// pel.cu — libavcodec h264qpel (8 bit), h264chroma (8 bit) and hpeldsp motion-compensation interpolation on sm_100a (C ABI: "h264qpel / hpeldsp").
//
// Reference semantics reproduced bit-for-bit (checker: oracle/pel_oracle.c):
//   H264_LOWPASS / H264_MC / op_put, op_avg   libavcodec/h264qpel_template.c:77-465
//   PIXOP2, rnd_avg32 / no_rnd_avg32          libavcodec/hpeldsp.c:38-333, libavcodec/rnd_avg.h:31-39
//
// Batched kernel: one warp per operation.  The warp first stages the (size+5) x (size+5) source window in shared
// memory (coalesced row reads through the read-only path), for positions that need the centre sample it builds the
// unrounded horizontal 6-tap sums once (size+5 rows of int16), and then every lane produces its pixels from shared
// memory.  The 6-tap is never recomputed per output pixel and the reference block is read from HBM exactly once.
#include "common.h"
#include "pel_hbd.h"
#include <algorithm>
#include <cstring>
#include <cstdlib>

namespace {

__device__ __forceinline__ int clip8(int v) { return __vimin_s32_relu(v, 255); }

constexpr int QW = 48;                 // window pitch in bytes (12 words: the 8-byte row segments of a half-warp fall in 16 distinct
                                       // bank pairs); only bytes 0..31 are used, block pixel (0,0) sits at byte 8 of window row 2
constexpr int QX = 8;                  // so that every 8-pixel segment of a row starts 8-byte aligned in shared memory
constexpr int WARPS = 4;

struct __align__(16) QpelSmem {
    uint8_t win[24 * QW];              // source rows -2 .. size+2; source column x is at byte QX + x (x = -2 .. size+2); rows 21..23: slack
                                       // for the tensor-core path, which reads the window in blocks of 8 rows
    short hraw[21 * 16];               // unrounded horizontal 6-tap sums for the same rows, block columns 0 .. 15
    unsigned raw[12 * 32];             // scratch of stage_window (TMA-less single-window form): word k of lane l at raw[k * 32 + l]
    uint8_t ring[4][16 + 21 * 48];     // qpel_kernel: raw 16-byte chunks of QD windows in flight (cp.async), window row r at ring[slot][16 + r * 48]
                                       // (16 bytes of slack: the re-alignment may read the word in front of a row)
};

// ---- packed arithmetic helpers --------------------------------------------------------------------------------------
// The 6-tap (1,-5,20,20,-5,1) is evaluated with the integer dot-product instructions: IDP.4A for the horizontal filter
// (u8 pixels x s8 taps, 4 taps per instruction), IDP.2A for the vertical ones (two rows interleaved by PRMT, s16 tap pair
// x u8 pixels for V; s16 sums x s8 tap pair for the centre sample).  Every component of a quarter-pel position ends up as
// two packed u8x4 words, so the (a + b + 1) >> 1 combinations are byte-SIMD averages.
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c)      // c + sum_k a.u8[k] * b.s8[k]
{ int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_lo_su(int a, unsigned b, int c)   // c + a.s16[0]*b.u8[0] + a.s16[1]*b.u8[1]
{ int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_hi_su(int a, unsigned b, int c)   // c + a.s16[0]*b.u8[2] + a.s16[1]*b.u8[3]
{ int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_lo_ss(unsigned a, int b, int c)   // c + a.s16[0]*b.s8[0] + a.s16[1]*b.s8[1]
{ int d; asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

constexpr int TAP4_A = 0x1414FB01;     // s8 (1, -5, 20, 20)
constexpr int TAP4_B = 0x000001FB;     // s8 (-5, 1, 0, 0)
__host__ __device__ constexpr int tap2_16(int j) { return j == 0 ? (int)0xFFFB0001 : j == 1 ? 0x00140014 : 0x0001FFFB; }   // s16 pairs (1,-5) (20,20) (-5,1)
__host__ __device__ constexpr int tap2_8(int j)  { return j == 0 ? 0x0000FB01 : j == 1 ? 0x00001414 : 0x000001FB; }        // s8 pairs, same taps

// 8 packed bytes of a window row: block columns x0+dx .. x0+dx+7 (dx = 0 or 1)
__device__ __forceinline__ uint2 row8p(const uint8_t *row, int x0, int dx)
{
    const unsigned *w = reinterpret_cast<const unsigned *>(row + QX + x0);
    unsigned w0 = w[0], w1 = w[1];
    if (dx) { const unsigned w2 = w[2]; w0 = __funnelshift_r(w0, w1, 8); w1 = __funnelshift_r(w1, w2, 8); }
    return make_uint2(w0, w1);
}
// unrounded horizontal 6-tap sums (+ init) for block columns x0 .. x0+7 of window row `row`
__device__ __forceinline__ void hsum8(const uint8_t *row, int x0, int init, int *o)
{
    const unsigned *w = reinterpret_cast<const unsigned *>(row + QX + x0 - 4);     // A[0] = columns x0-4 .. x0-1
    const unsigned A[5] = { w[0], w[1], w[2], w[3], w[4] };
    unsigned S[12];                                                                  // S[i] = columns x0-2+i .. x0+1+i
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const int b = 2 + i;
        S[i] = (b & 3) ? __funnelshift_r(A[b >> 2], A[(b >> 2) + 1], (b & 3) * 8) : A[b >> 2];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = dp4a_us(S[i + 4], TAP4_B, dp4a_us(S[i], TAP4_A, init));
}
// one vertical tap pair over two rows of 8 packed pixels
__device__ __forceinline__ void vfold(int k2, const uint2 &r0, const uint2 &r1, int *acc)
{
    unsigned lo = __byte_perm(r0.x, r1.x, 0x5140), hi = __byte_perm(r0.x, r1.x, 0x7362);
    acc[0] = dp2a_lo_su(k2, lo, acc[0]); acc[1] = dp2a_hi_su(k2, lo, acc[1]); acc[2] = dp2a_lo_su(k2, hi, acc[2]); acc[3] = dp2a_hi_su(k2, hi, acc[3]);
    lo = __byte_perm(r0.y, r1.y, 0x5140); hi = __byte_perm(r0.y, r1.y, 0x7362);
    acc[4] = dp2a_lo_su(k2, lo, acc[4]); acc[5] = dp2a_hi_su(k2, lo, acc[5]); acc[6] = dp2a_lo_su(k2, hi, acc[6]); acc[7] = dp2a_hi_su(k2, hi, acc[7]);
}
// clip8(v[k] >> SH) for 8 values -> two packed words
template <int SH>
__device__ __forceinline__ uint2 clip_pack8(const int *v)
{
    unsigned c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = (unsigned)clip8(v[k] >> SH);
    uint2 o;
    o.x = __byte_perm(__byte_perm(c[0], c[1], 0x1140), __byte_perm(c[2], c[3], 0x1140), 0x5410);
    o.y = __byte_perm(__byte_perm(c[4], c[5], 0x1140), __byte_perm(c[6], c[7], 0x1140), 0x5410);
    return o;
}

// The kernel is latency-bound if every operation waits for its descriptor, then its window, then (avg) its
// destination.  Each warp therefore walks QK consecutive operations and software-pipelines them: while operation t is
// computed, the window words of t+1 are already in flight into registers, the descriptor of t+2 is being fetched, and
// the destination row of t (avg) was requested before the shared-memory staging started.
constexpr int QK = 32;               // operations per warp (one descriptor per lane): the start-up round trips (descriptors, first windows) are paid once per 32 blocks

struct QMeta { int o; long long soff, doff; };
__device__ __forceinline__ QMeta load_meta(const uint8_t *op, const int64_t *src_off, const int64_t *dst_off, long long i, long long n)
{
    QMeta m; m.o = -1; m.soff = 0; m.doff = 0;
    if (i < n) { m.o = __ldg(op + i); m.soff = __ldg(src_off + i); m.doff = __ldg(dst_off + i); }
    return m;
}
struct QWin { uint4 v[3]; unsigned sh; };
// Window row `lane` (source row lane-2, columns -2 .. size+2) into registers: the one to three aligned 16-byte words that
// contain at least one needed byte (nothing else is ever read), plus the byte offset of column -2 inside the first.
__device__ __forceinline__ void load_window(QWin &w, const uint8_t *sp, long long stride, int wdim, int lane, bool valid)
{
    const uintptr_t first = reinterpret_cast<uintptr_t>(sp + (long long)(lane - 2) * stride - 2);
    const uint4 *q = reinterpret_cast<const uint4 *>(first & ~(uintptr_t)15);
    w.sh = (unsigned)(first & 15);
    const int nv = (int)((w.sh + wdim + 15) >> 4);
    w.v[0] = w.v[1] = w.v[2] = make_uint4(0, 0, 0, 0);
    if (valid && lane < wdim) {
        w.v[0] = __ldg(q);
        if (nv > 1) w.v[1] = __ldg(q + 1);
        if (nv > 2) w.v[2] = __ldg(q + 2);
    }
}
// Each lane re-aligns its own row: raw words go through a private column of shared memory (a register array cannot be
// indexed by the run-time word shift), come back as 7 words at the shifted position and are funnel-shifted so that block
// column 0 sits at byte QX of the window row.  No cross-lane traffic, so no barrier between the two steps.
__device__ __forceinline__ void stage_window(QpelSmem &s, const QWin &w, int wdim, int lane)
{
    if (lane >= wdim) return;
    unsigned *raw = &s.raw[lane];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        raw[(4 * k + 0) * 32] = w.v[k].x; raw[(4 * k + 1) * 32] = w.v[k].y; raw[(4 * k + 2) * 32] = w.v[k].z; raw[(4 * k + 3) * 32] = w.v[k].w;
    }
    const int ws = (int)((w.sh + 2) >> 2) - 1;                   // window word j = raw bytes sh + 4j - 6 .. sh + 4j - 3
    const unsigned fs = ((w.sh + 2) & 3) * 8;
    unsigned x[7];
#pragma unroll
    for (int i = 0; i < 7; i++) x[i] = raw[max(ws + i, 0) * 32];   // ws + i <= 9; index -1 only feeds columns -4, -3 (unused)
    uint4 *row = reinterpret_cast<uint4 *>(&s.win[lane * QW]);
    row[0] = make_uint4(0u, __funnelshift_r(x[0], x[1], fs), __funnelshift_r(x[1], x[2], fs), __funnelshift_r(x[2], x[3], fs));
    row[1] = make_uint4(__funnelshift_r(x[3], x[4], fs), __funnelshift_r(x[4], x[5], fs), __funnelshift_r(x[5], x[6], fs), 0u);
}

// Windows in flight.  One window per warp in flight (the register-staged form above) leaves the kernel bound by memory latency:
// at 2.2 G blocks/s an SM needs ~25 windows under way and 28 resident warps provide just that — the tensor-core path computes a block
// in half the instructions and was SLOWER with it (1.6 G blocks/s at 20 warps per SM).  So the raw 16-byte chunks now go global ->
// shared with cp.async into a ring of QD windows per warp (no registers held while they fly), QD - 1 operations ahead of the one
// being computed; the lane that copied a row re-aligns it out of the ring when its turn comes.
constexpr int QD = 4;
__device__ __forceinline__ void q_cp_async16(void *smem_dst, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(g) : "memory");
}
__device__ __forceinline__ void q_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void q_cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
// request window row `lane` (source row lane - 2, columns -2 .. size + 2) of one operation: the one to three aligned 16-byte words
// that contain at least one needed byte; always closes a group, so that the group count is the operation count.
//   first = address of the row's column -2; rl = shared address of this lane's row in the ring slot
__device__ __forceinline__ void issue_window(unsigned rl, uintptr_t first, int wdim, int lane, bool valid)
{
    if (valid && lane < wdim) {
        const uint8_t *q = reinterpret_cast<const uint8_t *>(first & ~(uintptr_t)15);
        const int nv = (int)(((first & 15) + wdim + 15) >> 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(rl), "l"(q) : "memory");
        if (nv > 1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(rl + 16), "l"(q + 16) : "memory");
        if (nv > 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(rl + 32), "l"(q + 32) : "memory");
    }
    q_cp_commit();
}
// the lane's own row out of the ring, block column 0 moved to byte QX of the window row (same result as stage_window);
// sh = low four bits of the row's first address, raw = the lane's row in the ring slot
__device__ __forceinline__ void stage_from_ring(QpelSmem &s, const unsigned *raw, unsigned sh, int wdim, int lane)
{
    if (lane >= wdim) return;
    const int ws = (int)((sh + 2) >> 2) - 1;                     // window word j = raw bytes sh + 4j - 6 .. sh + 4j - 3; ws in -1 .. 3
    const unsigned fs = ((sh + 2) & 3) * 8;
    unsigned x[7];
#pragma unroll
    for (int i = 0; i < 7; i++) x[i] = raw[ws + i];              // word -1 (slack / previous row) and words never copied only feed columns nobody reads
    uint4 *row = reinterpret_cast<uint4 *>(&s.win[lane * QW]);
    row[0] = make_uint4(0u, __funnelshift_r(x[0], x[1], fs), __funnelshift_r(x[1], x[2], fs), __funnelshift_r(x[2], x[3], fs));
    row[1] = make_uint4(__funnelshift_r(x[3], x[4], fs), __funnelshift_r(x[4], x[5], fs), __funnelshift_r(x[5], x[6], fs), 0u);
}

// ---- 16 x 16 blocks on the tensor cores -------------------------------------------------------------------------------------------
// The 6-tap filters are banded matrices.  With the staged window at a fixed phase (block column 0 at byte QX of a window row) the
// horizontal filter of all 16 columns over 8 window rows is ONE mma.sync.m16n8k32:  D[x][r] = sum_k A[x][k] * B[k][r],  A[x][k] = tap[k - x - 2]
// (s8, the same for every operation: built once per lane), B[k][r] = window byte 4 + k of row r (u8: the B fragment of a lane is two
// aligned 32-bit loads of its row).  An A made of a single 1 per row instead (k = x + 4, or x + 5 for the +1 column positions) gives
// the window itself in the same layout — which is what the vertical filter wants as ITS left operand:  D2[x][y] = sum_r T[x][r] * tapV[r - y].
// A lane's accumulators of the three 8-row blocks (rows 2t, 2t+1, 8+2t, 9+2t, 16+2t, 17+2t of columns g and g + 8) are exactly six of the
// contraction slots of its A fragment, so the first pass feeds the second straight out of registers (bytes picked by PRMT; the raw
// horizontal sums of the centre position are split  h = 256 * hi + lo  into a signed and an unsigned byte plane, two MMAs); the matching
// slot -> row permutation is folded into the constant B fragments of the vertical taps.  Every plane (F, H, V, J) comes out in the
// layout  (column g / g + 8, rows 8 nb + 2t, + 1),  so rounding, clipping and the (a + b + 1) >> 1 combinations are element-wise.
// A 16 x 16 centre position costs 6 LDS + 7 MMA + ~100 ALU per warp instead of ~370 instructions on the dot-product path.
__device__ __forceinline__ void qmma_s8u8(int (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void qmma_u8s8(int (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void qmma_s8s8(int (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ int qtap(int j) { return j == 0 || j == 5 ? 1 : j == 1 || j == 4 ? -5 : j == 2 || j == 3 ? 20 : 0; }
struct QmmaConst {
    unsigned atap[4];                  // A of the horizontal filter
    unsigned aid[2][4];                // A of the identity pick, column shift 0 / 1
    unsigned bv[2][2];                 // B of the vertical filter for output rows 0..7 / 8..15 (contraction slots in accumulator order)
};
__device__ __forceinline__ void qmma_consts(QmmaConst &c, int lane)
{
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int r = 0; r < 4; r++) {                                  // a0: row g, k 4t..; a1: row g + 8; a2 / a3: the same rows, k + 16
        const int x = g + (r & 1) * 8, k0 = 4 * t + (r >> 1) * 16;
        unsigned wt = 0, w0 = 0, w1 = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int k = k0 + b;
            wt |= ((unsigned)qtap(k - x - 2) & 0xffu) << (8 * b);   // window byte 4 + k = block column k - 4; taps start at column x - 2
            w0 |= (k == x + 4 ? 1u : 0u) << (8 * b);
            w1 |= (k == x + 5 ? 1u : 0u) << (8 * b);
        }
        c.atap[r] = wt; c.aid[0][r] = w0; c.aid[1][r] = w1;
    }
#pragma unroll
    for (int nb = 0; nb < 2; nb++)
#pragma unroll
        for (int h = 0; h < 2; h++) {                              // b0: slots 4t .. 4t+3, b1: slots 16 + 4t .. ; column n = g -> output row 8 nb + g
            unsigned w = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int row = h == 0 ? (b < 2 ? 2 * t + b : 8 + 2 * t + (b - 2)) : (b < 2 ? 16 + 2 * t + b : 99);   // slot -> window row
                const int j = row - (8 * nb + g);                  // output row y reads window rows y .. y + 5
                w |= ((unsigned)(row < 24 ? qtap(j) : 0) & 0xffu) << (8 * b);
            }
            c.bv[nb][h] = w;
        }
}
// bytes `sel` (0: bits 0..7, 1: bits 8..15) of four accumulators -> one word
__device__ __forceinline__ unsigned qpick4(int r0, int r1, int r2, int r3, int sel)
{
    const unsigned s = sel ? 0x0051u : 0x0040u;
    return __byte_perm(__byte_perm((unsigned)r0, (unsigned)r1, s), __byte_perm((unsigned)r2, (unsigned)r3, s), 0x5410);
}
// A fragments of the vertical pass out of the three base-0 accumulator blocks; sel picks the low or the high byte plane
__device__ __forceinline__ void qfragV(unsigned (&a)[4], const int (&d0)[4], const int (&d1)[4], const int (&d2)[4], int sel)
{
    a[0] = qpick4(d0[0], d0[1], d1[0], d1[1], sel);                // column g:     rows 2t, 2t+1, 8+2t, 9+2t
    a[1] = qpick4(d0[2], d0[3], d1[2], d1[3], sel);                // column g + 8
    a[2] = qpick4(d2[0], d2[1], 0, 0, sel);                        // rows 16+2t, 17+2t, (no row), (no row)
    a[3] = qpick4(d2[2], d2[3], 0, 0, sel);
}

// Lane l works on one 8-pixel (4 for size 4) row segment: row l>>1, half l&1 for 16x16.
template <bool MMA>
__global__ void __launch_bounds__(32 * WARPS)
qpel_kernel(long long n, const uint8_t *op, uint8_t *dst, const int64_t *dst_off, const uint8_t *src,
            const int64_t *src_off, long long stride)
{
    __shared__ QpelSmem sm[WARPS];
    __shared__ unsigned qtab[16][32];                                                    // the per-lane MMA constants: [0..3] taps A, [4..7] / [8..11] pick A
                                                                                         // (column shift 0 / 1), [12..15] vertical taps B (rows 0..7: b0 b1, 8..15: b0 b1)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (MMA && warp == 0) {       // kept in shared memory: as loop-invariant registers ptxas rebuilt them (150 instructions) for every operation
        QmmaConst qc;
        qmma_consts(qc, lane);
#pragma unroll
        for (int r = 0; r < 4; r++) { qtab[r][lane] = qc.atap[r]; qtab[4 + r][lane] = qc.aid[0][r]; qtab[8 + r][lane] = qc.aid[1][r]; }
        qtab[12][lane] = qc.bv[0][0]; qtab[13][lane] = qc.bv[0][1]; qtab[14][lane] = qc.bv[1][0]; qtab[15][lane] = qc.bv[1][1];
    }
    if (MMA) __syncthreads();
    const long long first_op = ((long long)blockIdx.x * WARPS + warp) * QK;
    if (first_op >= n) return;
    QpelSmem &s = sm[warp];
    // lane k (< QK) owns the descriptor of the warp's operation k; everybody gets it by shuffle when the operation is issued / computed
    const QMeta mine_m = load_meta(op, src_off, dst_off, first_op + lane, lane < QK ? n : 0);
    const long long rowoff = (long long)(lane - 2) * stride - 2;                          // this lane's window row, column -2, from the block's (0, 0)
    const uintptr_t srcb = reinterpret_cast<uintptr_t>(src) + (uintptr_t)rowoff;
    const unsigned ring0 = (unsigned)__cvta_generic_to_shared(&s.ring[0][16 + lane * 48]);
    constexpr unsigned RING_SLOT = (unsigned)sizeof(s.ring[0]);
#pragma unroll
    for (int k = 0; k < QD - 1; k++) {                                                   // QD - 1 windows ahead
        const int ok = __shfl_sync(0xffffffffu, mine_m.o, k);
        const long long sk = __shfl_sync(0xffffffffu, mine_m.soff, k);
        issue_window(ring0 + k * RING_SLOT, srcb + (uintptr_t)sk, (16 >> ((ok >> 1) & 3)) + 5, lane, ok >= 0);
    }
    for (int t = 0; t < QK; t++) {
        const int o = __shfl_sync(0xffffffffu, mine_m.o, t);
        if (o < 0) break;                                                                // past the end (warp-uniform)
        const long long soff0 = __shfl_sync(0xffffffffu, mine_m.soff, t);
        const long long doff0 = __shfl_sync(0xffffffffu, mine_m.doff, t);
        {
            const int ka = t + QD - 1, kk = ka < QK ? ka : QK - 1;
            const int oa = __shfl_sync(0xffffffffu, mine_m.o, kk);
            const long long sa = __shfl_sync(0xffffffffu, mine_m.soff, kk);
            issue_window(ring0 + (ka & (QD - 1)) * RING_SLOT, srcb + (uintptr_t)sa, (16 >> ((oa >> 1) & 3)) + 5, lane, ka < QK && oa >= 0);
        }

        const int avg = o & 1, size = 16 >> ((o >> 1) & 3), qx = (o >> 3) & 3, qy = (o >> 5) & 3;
        uint8_t *dp = dst + doff0;
        const int wdim = size + 5;
        int npx = 8, segs = 2, y = lane >> 1, x0 = (lane & 1) << 3;
        bool mine = true;
        if (size != 16) {                                                                // 8 x 8 and 4 x 4: one segment per row
            npx = size < 8 ? size : 8; segs = 1;
            mine = lane < size; y = lane; x0 = 0;
        }
        uint8_t *d = dp + (long long)y * stride + x0;
        const bool vec = npx == 8 && ((reinterpret_cast<uintptr_t>(d)) & 7) == 0;
        uint2 pv = make_uint2(0, 0);
        if (mine && avg && vec) pv = *reinterpret_cast<const uint2 *>(d);                // destination row requested early

        q_cp_wait<QD - 1>();                                                             // this operation's chunks have landed (own copies)
        stage_from_ring(s, reinterpret_cast<const unsigned *>(&s.ring[t & (QD - 1)][16 + lane * 48]),
                        ((unsigned)srcb + (unsigned)soff0) & 15u, wdim, lane);
        __syncwarp();
        const bool need_j = (qx == 2 && qy != 0) || (qy == 2 && qx != 0);                // positions built from the centre sample
        if (MMA && size == 16) {
            const int g = lane >> 2, tq = lane & 3;
            const bool useF = (qy == 0 && qx != 2) || (qx == 0 && (qy & 1));
            const bool useH = qx != 0 && qy != 2;
            const bool useV = qy != 0 && qx != 2;
            const int cs = qx == 3, rbase = 2 + (qy == 3);
            unsigned aid[4] = { 0, 0, 0, 0 }, atap[4] = { 0, 0, 0, 0 }, bv[2][2] = { { 0, 0 }, { 0, 0 } };
            if (useF || useV) {
#pragma unroll
                for (int r = 0; r < 4; r++) aid[r] = qtab[4 + 4 * cs + r][lane];
            }
            if (useH || need_j) {
#pragma unroll
                for (int r = 0; r < 4; r++) atap[r] = qtab[r][lane];
            }
            if (useV || need_j) { bv[0][0] = qtab[12][lane]; bv[0][1] = qtab[13][lane]; bv[1][0] = qtab[14][lane]; bv[1][1] = qtab[15][lane]; }
            int res[2][4];
            bool have = false;                                           // warp-uniform: the branch below is not divergent
            auto add = [&](const int (&c0)[4], const int (&c1)[4]) {
                if (have) {
#pragma unroll
                    for (int i = 0; i < 4; i++) { res[0][i] = (res[0][i] + c0[i] + 1) >> 1; res[1][i] = (res[1][i] + c1[i] + 1) >> 1; }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) { res[0][i] = c0[i]; res[1][i] = c1[i]; }
                }
                have = true;
            };
            const uint8_t *wb = &s.win[g * QW + 4 + 4 * tq];            // this lane's B-fragment bytes of window row g
            if (useF || useH) {                                          // planes taken at the output rows themselves
                unsigned b0[2], b1[2];
#pragma unroll
                for (int nb = 0; nb < 2; nb++) {
                    const uint8_t *p = wb + (rbase + 8 * nb) * QW;
                    b0[nb] = *reinterpret_cast<const unsigned *>(p); b1[nb] = *reinterpret_cast<const unsigned *>(p + 16);
                }
                if (useF) {
                    int f0[4] = { 0, 0, 0, 0 }, f1[4] = { 0, 0, 0, 0 };
                    qmma_s8u8(f0, aid, b0[0], b1[0]); qmma_s8u8(f1, aid, b0[1], b1[1]);
                    add(f0, f1);
                }
                if (useH) {
                    int h0[4] = { 16, 16, 16, 16 }, h1[4] = { 16, 16, 16, 16 };
                    qmma_s8u8(h0, atap, b0[0], b1[0]); qmma_s8u8(h1, atap, b0[1], b1[1]);
#pragma unroll
                    for (int i = 0; i < 4; i++) { h0[i] = clip8(h0[i] >> 5); h1[i] = clip8(h1[i] >> 5); }
                    add(h0, h1);
                }
            }
            if (useV || need_j) {                                        // planes filtered vertically: window rows 0 .. 20 as the contraction
                unsigned b0[3], b1[3];
#pragma unroll
                for (int nb = 0; nb < 3; nb++) {
                    const uint8_t *p = wb + (8 * nb) * QW;
                    b0[nb] = *reinterpret_cast<const unsigned *>(p); b1[nb] = *reinterpret_cast<const unsigned *>(p + 16);
                }
                if (useV) {
                    int t0[4] = { 0, 0, 0, 0 }, t1[4] = { 0, 0, 0, 0 }, t2[4] = { 0, 0, 0, 0 };
                    qmma_s8u8(t0, aid, b0[0], b1[0]); qmma_s8u8(t1, aid, b0[1], b1[1]); qmma_s8u8(t2, aid, b0[2], b1[2]);
                    unsigned a[4];
                    qfragV(a, t0, t1, t2, 0);
                    int v0[4] = { 16, 16, 16, 16 }, v1[4] = { 16, 16, 16, 16 };
                    qmma_u8s8(v0, a, bv[0][0], bv[0][1]); qmma_u8s8(v1, a, bv[1][0], bv[1][1]);
#pragma unroll
                    for (int i = 0; i < 4; i++) { v0[i] = clip8(v0[i] >> 5); v1[i] = clip8(v1[i] >> 5); }
                    add(v0, v1);
                }
                if (need_j) {
                    int t0[4] = { 0, 0, 0, 0 }, t1[4] = { 0, 0, 0, 0 }, t2[4] = { 0, 0, 0, 0 };
                    qmma_s8u8(t0, atap, b0[0], b1[0]); qmma_s8u8(t1, atap, b0[1], b1[1]); qmma_s8u8(t2, atap, b0[2], b1[2]);
                    unsigned al[4], ah[4];
                    qfragV(al, t0, t1, t2, 0);                           // h = 256 * (h >> 8) + (h & 255): unsigned low plane, signed high plane
                    qfragV(ah, t0, t1, t2, 1);
                    int jl0[4] = { 512, 512, 512, 512 }, jl1[4] = { 512, 512, 512, 512 }, jh0[4] = { 0, 0, 0, 0 }, jh1[4] = { 0, 0, 0, 0 };
                    qmma_u8s8(jl0, al, bv[0][0], bv[0][1]); qmma_u8s8(jl1, al, bv[1][0], bv[1][1]);
                    qmma_s8s8(jh0, ah, bv[0][0], bv[0][1]); qmma_s8s8(jh1, ah, bv[1][0], bv[1][1]);
#pragma unroll
                    for (int i = 0; i < 4; i++) { jl0[i] = clip8((jh0[i] * 256 + jl0[i]) >> 10); jl1[i] = clip8((jh1[i] * 256 + jl1[i]) >> 10); }
                    add(jl0, jl1);
                }
            }
            // the 16 x 16 result goes through a byte tile (the hraw array is free on this path) back to the row-segment mapping of the store
            uint8_t *ot = reinterpret_cast<uint8_t *>(s.hraw);
#pragma unroll
            for (int nb = 0; nb < 2; nb++)
#pragma unroll
                for (int i = 0; i < 4; i++)
                    ot[(8 * nb + 2 * tq + (i & 1)) * 16 + g + 8 * (i >> 1)] = (uint8_t)res[nb][i];
            __syncwarp();
            uint2 r8 = *reinterpret_cast<const uint2 *>(ot + y * 16 + x0);
            if (vec) {
                if (avg) { r8.x = __vavgu4(r8.x, pv.x); r8.y = __vavgu4(r8.y, pv.y); }
                *reinterpret_cast<uint2 *>(d) = r8;
            } else {
                for (int k = 0; k < 8; k++) {
                    const int v = (int)__byte_perm(k < 4 ? r8.x : r8.y, 0, 0x4440 | (k & 3));
                    d[k] = (uint8_t)(avg ? (d[k] + v + 1) >> 1 : v);
                }
            }
            __syncwarp();                                                                // tile and window are reused by the next operation
            continue;
        }
        if (need_j) {
            for (int k = lane; k < wdim * segs; k += 32) {
                const int r = k / segs, xx = (k - r * segs) * 8;
                int h[8];
                hsum8(&s.win[r * QW], xx, 0, h);
                uint4 pk;
                pk.x = __byte_perm(h[0], h[1], 0x5410); pk.y = __byte_perm(h[2], h[3], 0x5410);
                pk.z = __byte_perm(h[4], h[5], 0x5410); pk.w = __byte_perm(h[6], h[7], 0x5410);
                *reinterpret_cast<uint4 *>(&s.hraw[r * 16 + xx]) = pk;
            }
            __syncwarp();
        }
        if (mine) {
            // the one or two samples each quarter position averages (h264qpel_template.c:313-456): F full-pel, H horizontal
            // half, V vertical half, J centre.  acc collects them; two samples -> (a + b + 1) >> 1.
            const bool useF = (qy == 0 && qx != 2) || (qx == 0 && (qy & 1));
            const bool useH = qx != 0 && qy != 2;
            const bool useV = qy != 0 && qx != 2;
            uint2 res = make_uint2(0, 0);
            bool have = false;
            auto add = [&](const uint2 &c) {
                if (have) { res.x = __vavgu4(res.x, c.x); res.y = __vavgu4(res.y, c.y); }
                else { res = c; have = true; }
            };
            if (useF) add(row8p(&s.win[(y + 2 + (qy == 3)) * QW], x0, qx == 3));
            if (useH) {
                int h[8];
                hsum8(&s.win[(y + 2 + (qy == 3)) * QW], x0, 16, h);
                add(clip_pack8<5>(h));
            }
            if (useV) {
                const int dx = qx == 3;
                int a8[8];
#pragma unroll
                for (int k = 0; k < 8; k++) a8[k] = 16;
#pragma unroll
                for (int j = 0; j < 3; j++)
                    vfold(tap2_16(j), row8p(&s.win[(y + 2 * j) * QW], x0, dx), row8p(&s.win[(y + 2 * j + 1) * QW], x0, dx), a8);
                add(clip_pack8<5>(a8));
            }
            if (need_j) {
                int j6[8];
#pragma unroll
                for (int k = 0; k < 8; k++) j6[k] = 512;
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const uint4 pa = *reinterpret_cast<const uint4 *>(&s.hraw[(y + 2 * j) * 16 + x0]);
                    const uint4 pb = *reinterpret_cast<const uint4 *>(&s.hraw[(y + 2 * j + 1) * 16 + x0]);
                    const unsigned wa[4] = { pa.x, pa.y, pa.z, pa.w }, wb[4] = { pb.x, pb.y, pb.z, pb.w };
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        j6[2 * k]     = dp2a_lo_ss(__byte_perm(wa[k], wb[k], 0x5410), tap2_8(j), j6[2 * k]);
                        j6[2 * k + 1] = dp2a_lo_ss(__byte_perm(wa[k], wb[k], 0x7632), tap2_8(j), j6[2 * k + 1]);
                    }
                }
                add(clip_pack8<10>(j6));
            }
            if (vec) {
                if (avg) { res.x = __vavgu4(res.x, pv.x); res.y = __vavgu4(res.y, pv.y); }
                *reinterpret_cast<uint2 *>(d) = res;
            } else {
                for (int k = 0; k < npx; k++) {
                    const int v = (int)__byte_perm(k < 4 ? res.x : res.y, 0, 0x4440 | (k & 3));
                    d[k] = (uint8_t)(avg ? (d[k] + v + 1) >> 1 : v);
                }
            }
        }
        __syncwarp();                                                                    // shared window is reused by the next operation
    }
    q_cp_wait<0>();                                                                      // nothing of this warp's is in flight when it leaves
}

// ------------------------------------------------------------------------------------------------ qpel on TMA-staged windows
// Same arithmetic as qpel_kernel; what changes is how a window reaches shared memory: one cp.async.bulk.tensor.2d (UTMALDG) per
// operation completes on the operation's mbarrier.  A warp issues the boxes of all its TQK operations before it computes the first
// one, so nothing of the load path runs on the LSU any more (no per-row 16-byte loads, no register -> shared staging).
// The hardware wants the box origin on a 16-byte column (measured on the B200: any other start column raises an illegal-instruction
// fault, scripts/probe/tma_probe.cu), so the box of a block at column x starts at (x - 2) & ~15 and is 48 bytes wide for 16-pixel blocks
// (window of 21 bytes at phase 0..15) and 32 bytes for 8- and 4-pixel blocks (13 / 9 bytes); the window's byte phase inside the tile
// row, ph = (x - 2) & 15, is folded into the word index and one funnel shift of the row readers.  Boxes that hang over the end of a
// line are legal (the tensor map's column extent is the line pitch: the missing bytes arrive as zeros and are never used).
// Blocks closer than 2 bytes to the start of a line or 2 rows to the start of the plane are filled by plain loads instead.
constexpr int TQK = 8;                  // operations per warp
constexpr int TILE_BYTES = 1024;        // 21 rows x 48 bytes, rounded up to the 128-byte alignment the copy engine wants
struct WinT {
    const unsigned *t;
    int pw;                             // words per tile row (12 or 8)
    int off0;                           // byte of block column 0 inside a tile row (ph + 2)
    __device__ __forceinline__ unsigned w(int row, int k) const { return t[row * pw + k]; }
};
// 8 pixels of `row` starting at block column x0 + dx
__device__ __forceinline__ uint2 t_row8p(const WinT &W, int row, int x0, int dx)
{
    const int o = W.off0 + x0 + dx, k = o >> 2;
    const unsigned sh = (unsigned)(o & 3) * 8u;
    const unsigned w0 = W.w(row, k), w1 = W.w(row, k + 1), w2 = W.w(row, k + 2);
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
// raw 6-tap horizontal sums of the 8 pixels at block columns x0 .. x0 + 7
__device__ __forceinline__ void t_hsum8(const WinT &W, int row, int x0, int init, int *o)
{
    const int s = W.off0 + x0 - 2, k = s >> 2;
    const unsigned q = (unsigned)(s & 3) * 8u;
    const unsigned R[5] = { W.w(row, k), W.w(row, k + 1), W.w(row, k + 2), W.w(row, k + 3), W.w(row, k + 4) };
    const unsigned A[4] = { __funnelshift_r(R[0], R[1], q), __funnelshift_r(R[1], R[2], q), __funnelshift_r(R[2], R[3], q), __funnelshift_r(R[3], R[4], q) };
    unsigned S[12];                     // S[i] = window bytes i .. i + 3, byte 0 = block column x0 - 2
#pragma unroll
    for (int i = 0; i < 12; i++)
        S[i] = (i & 3) ? __funnelshift_r(A[i >> 2], A[(i >> 2) + ((i >> 2) < 3 ? 1 : 0)], (i & 3) * 8) : A[i >> 2];
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = dp4a_us(S[i + 4], TAP4_B, dp4a_us(S[i], TAP4_A, init));
}

__device__ __forceinline__ uint32_t pel_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pel_mbar_wait(uint32_t a, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void pel_tma_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}

struct __align__(1024) QpelTmaSmem {
    uint8_t tile[WARPS][TQK][TILE_BYTES];
    short hraw[WARPS][21 * 16];
    unsigned long long mbar[WARPS][TQK];
};

__global__ void __launch_bounds__(32 * WARPS)
qpel_tma_kernel(const __grid_constant__ CUtensorMap tm16, const __grid_constant__ CUtensorMap tm8, const __grid_constant__ CUtensorMap tm4,
                long long n, const uint8_t *op, uint8_t *dst, const int64_t *dst_off, const uint8_t *src, const int64_t *src_off,
                long long stride, unsigned long long magic)
{
    __shared__ QpelTmaSmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long first_op = ((long long)blockIdx.x * WARPS + warp) * TQK;
    if (first_op >= n) return;
    // lanes 0 .. TQK-1 each own one operation's descriptor
    int mo = -1, inb = 0, bx = 0, by = 0, mph = 0;
    long long msoff = 0, mdoff = 0;
    if (lane < TQK) {
        const uint32_t mb = pel_smem_u32(&sm.mbar[warp][lane]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb) : "memory");
        if (first_op + lane < n) {
            mo = __ldg(op + first_op + lane); msoff = __ldg(src_off + first_op + lane); mdoff = __ldg(dst_off + first_op + lane);
        }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (mo >= 0 && msoff >= 0) {
        // row and column of the block inside the plane: soff = y * stride + x (reciprocal multiply, corrected)
        unsigned long long y = __umul64hi((unsigned long long)msoff, magic);
        long long x = msoff - (long long)y * stride;
        while (x >= stride) { x -= stride; y++; }
        if (x >= 2 && y >= 2 && y < 0x7fffff00ULL) { inb = 1; mph = (int)(x - 2) & 15; bx = (int)(x - 2) - mph; by = (int)y - 2; }
    }
    // the copies are issued by one lane (TMA operands live in uniform registers): lane 0 walks the warp's operations
    for (int k = 0; k < TQK; k++) {
        const int kin = __shfl_sync(0xffffffffu, inb, k), kx = __shfl_sync(0xffffffffu, bx, k), ky = __shfl_sync(0xffffffffu, by, k);
        const int ksize = 16 >> ((__shfl_sync(0xffffffffu, mo, k) >> 1) & 3);
        if (kin && lane == 0) {
            const CUtensorMap *tm = ksize == 16 ? &tm16 : ksize == 8 ? &tm8 : &tm4;
            pel_tma_2d(pel_smem_u32(&sm.tile[warp][k][0]), tm, kx, ky, pel_smem_u32(&sm.mbar[warp][k]), (ksize == 16 ? 48u : 32u) * (ksize + 5));
        }
    }
    short *hraw = sm.hraw[warp];
    for (int t = 0; t < TQK; t++) {
        const int o = __shfl_sync(0xffffffffu, mo, t);
        if (o < 0) break;                                                                  // past the end (warp-uniform)
        const long long doff = __shfl_sync(0xffffffffu, mdoff, t);
        const int avg = o & 1, size = 16 >> ((o >> 1) & 3), qx = (o >> 3) & 3, qy = (o >> 5) & 3;
        uint8_t *dp = dst + doff;
        const int wdim = size + 5;
        const int npx = size < 8 ? size : 8, segs = size >> 3 ? size >> 3 : 1;
        const bool mine = lane < size * segs;
        const int y = lane / segs, x0 = (lane - y * segs) * 8;
        uint8_t *d = dp + (long long)y * stride + x0;
        const bool vec = npx == 8 && ((reinterpret_cast<uintptr_t>(d)) & 7) == 0;
        uint2 pv = make_uint2(0, 0);
        if (mine && avg && vec) pv = *reinterpret_cast<const uint2 *>(d);                  // destination row requested early
        uint8_t *tile = sm.tile[warp][t];
        const int pwb = size == 16 ? 48 : 32;                                              // bytes per tile row
        const int ph = __shfl_sync(0xffffffffu, mph, t);                                   // 0 for the plain-load fill
        if (__shfl_sync(0xffffffffu, inb, t)) {
            pel_mbar_wait(pel_smem_u32(&sm.mbar[warp][t]), 0);
        } else {                                                                           // first columns / rows of the plane: plain loads
            const long long soff = __shfl_sync(0xffffffffu, msoff, t);
            if (lane < wdim) {
                const uint8_t *rp = src + soff + (long long)(lane - 2) * stride - 2;
                for (int k = 0; k < wdim; k++) tile[lane * pwb + k] = __ldg(rp + k);
            }
            __syncwarp();
        }
        const WinT W{ reinterpret_cast<const unsigned *>(tile), pwb >> 2, ph + 2 };
        const bool need_j = (qx == 2 && qy != 0) || (qy == 2 && qx != 0);
        if (need_j) {
            for (int k = lane; k < wdim * segs; k += 32) {
                const int r = k / segs, xx = (k - r * segs) * 8;
                int h[8];
                t_hsum8(W, r, xx, 0, h);
                uint4 pk;
                pk.x = __byte_perm(h[0], h[1], 0x5410); pk.y = __byte_perm(h[2], h[3], 0x5410);
                pk.z = __byte_perm(h[4], h[5], 0x5410); pk.w = __byte_perm(h[6], h[7], 0x5410);
                *reinterpret_cast<uint4 *>(&hraw[r * 16 + xx]) = pk;
            }
            __syncwarp();
        }
        if (mine) {
            const bool useF = (qy == 0 && qx != 2) || (qx == 0 && (qy & 1));
            const bool useH = qx != 0 && qy != 2;
            const bool useV = qy != 0 && qx != 2;
            uint2 res = make_uint2(0, 0);
            bool have = false;
            auto add = [&](const uint2 &c) {
                if (have) { res.x = __vavgu4(res.x, c.x); res.y = __vavgu4(res.y, c.y); }
                else { res = c; have = true; }
            };
            if (useF) add(t_row8p(W, y + 2 + (qy == 3), x0, qx == 3));
            if (useH) {
                int h[8];
                t_hsum8(W, y + 2 + (qy == 3), x0, 16, h);
                add(clip_pack8<5>(h));
            }
            if (useV) {
                const int dx = qx == 3;
                int a8[8];
#pragma unroll
                for (int k = 0; k < 8; k++) a8[k] = 16;
#pragma unroll
                for (int j = 0; j < 3; j++)
                    vfold(tap2_16(j), t_row8p(W, y + 2 * j, x0, dx), t_row8p(W, y + 2 * j + 1, x0, dx), a8);
                add(clip_pack8<5>(a8));
            }
            if (need_j) {
                int j6[8];
#pragma unroll
                for (int k = 0; k < 8; k++) j6[k] = 512;
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const uint4 pa = *reinterpret_cast<const uint4 *>(&hraw[(y + 2 * j) * 16 + x0]);
                    const uint4 pb = *reinterpret_cast<const uint4 *>(&hraw[(y + 2 * j + 1) * 16 + x0]);
                    const unsigned wa[4] = { pa.x, pa.y, pa.z, pa.w }, wb[4] = { pb.x, pb.y, pb.z, pb.w };
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        j6[2 * k]     = dp2a_lo_ss(__byte_perm(wa[k], wb[k], 0x5410), tap2_8(j), j6[2 * k]);
                        j6[2 * k + 1] = dp2a_lo_ss(__byte_perm(wa[k], wb[k], 0x7632), tap2_8(j), j6[2 * k + 1]);
                    }
                }
                add(clip_pack8<10>(j6));
            }
            if (vec) {
                if (avg) { res.x = __vavgu4(res.x, pv.x); res.y = __vavgu4(res.y, pv.y); }
                *reinterpret_cast<uint2 *>(d) = res;
            } else {
                for (int k = 0; k < npx; k++) {
                    const int v = (int)__byte_perm(k < 4 ? res.x : res.y, 0, 0x4440 | (k & 3));
                    d[k] = (uint8_t)(avg ? (d[k] + v + 1) >> 1 : v);
                }
            }
        }
        __syncwarp();                                                                      // hraw is reused by the next operation
    }
}

// hpel: one warp per operation, direct global reads (at most 4 taps per pixel, rows are contiguous)
__global__ void __launch_bounds__(32 * WARPS)
hpel_kernel(long long n, const uint8_t *op, const uint8_t *hh, uint8_t *dst, const int64_t *dst_off, const uint8_t *src,
            const int64_t *src_off, long long stride)
{
    const long long i = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const int o = op[i], h = hh[i];
    const int tab = o & 3, sidx = (o >> 2) & 3, xy = (o >> 4) & 3, w = 16 >> sidx;
    const int no_rnd = tab >= 2;
    const int avg = (tab & 1) && !(sidx == 3 && xy == 3);        // avg_pixels2_xy2 stores without averaging (hpeldsp.c:134-166)
    const uint8_t *sp = src + src_off[i];
    uint8_t *dp = dst + dst_off[i];
    for (int k = lane; k < w * h; k += 32) {
        const int y = k / w, x = k - y * w;
        const uint8_t *p = sp + (long long)y * stride + x;
        int v;
        switch (xy) {
        case 0:  v = __ldg(p); break;
        case 1:  v = (__ldg(p) + __ldg(p + 1) + 1 - no_rnd) >> 1; break;
        case 2:  v = (__ldg(p) + __ldg(p + stride) + 1 - no_rnd) >> 1; break;
        default: v = (__ldg(p) + __ldg(p + 1) + __ldg(p + stride) + __ldg(p + stride + 1) + 2 - no_rnd) >> 2; break;
        }
        uint8_t *d = dp + (long long)y * stride + x;
        *d = (uint8_t)(avg ? (*d + v + 1) >> 1 : v);
    }
}

// h264chroma: bilinear eighth-pel (h264chroma_template.c:27-176).  8 lanes per operation, lane r owns row r (and r+8 of a
// 16-row block).  A row is fetched as the one or two aligned 8-byte words that hold bytes the reference itself reads (x == 0
// never touches the column right of the block, y == 0 never the row below it) and normalised to block column 0 with funnel
// shifts; the row below comes from the next lane by shuffle.  (A*t0 + B*t1 + C*b0 + D*b1 + 32) >> 6 is evaluated with
// coefficients scaled by 4, so that the result is byte 1 of the sum: two IDP.2A per pixel and no shifts.
constexpr int CH_LANES = 8;
struct ChRow { unsigned r0, r1, r2; };
__device__ __forceinline__ ChRow chroma_row(const uint8_t *p, int nbytes, bool on)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned sh = (unsigned)(a & 7);
    const uint2 *q = reinterpret_cast<const uint2 *>(a & ~(uintptr_t)7);
    uint2 v0 = make_uint2(0, 0), v1 = make_uint2(0, 0);
    if (on) {
        v0 = __ldg(q);
        if (sh + nbytes > 8) v1 = __ldg(q + 1);
    }
    const bool up = sh & 4;
    const unsigned t0 = up ? v0.y : v0.x, t1 = up ? v1.x : v0.y, t2 = up ? v1.y : v1.x, t3 = up ? 0u : v1.y;
    const unsigned fs = (sh & 3) * 8;
    ChRow o;
    o.r0 = __funnelshift_r(t0, t1, fs); o.r1 = __funnelshift_r(t1, t2, fs); o.r2 = __funnelshift_r(t2, t3, fs);
    return o;
}
__global__ void __launch_bounds__(32 * WARPS)
chroma_kernel(long long n, const uint8_t *op, const uint8_t *hh, const uint8_t *xy, uint8_t *dst, const int64_t *dst_off,
              const uint8_t *src, const int64_t *src_off, long long stride)
{
    const long long gi = ((long long)blockIdx.x * (32 * WARPS) + threadIdx.x) / CH_LANES;
    const bool valid = gi < n;
    const long long i = valid ? gi : n - 1;                       // every lane stays in the shuffles
    const int sub = threadIdx.x & (CH_LANES - 1);
    const int o = __ldg(op + i), h = __ldg(hh + i), pq = __ldg(xy + i), fx = pq & 7, fy = (pq >> 3) & 7;
    const int avg = o & 1, w = 8 >> ((o >> 1) & 3);
    const int kab = 4 * ((8 - fx) * (8 - fy)) | (4 * (fx * (8 - fy))) << 16;      // s16 pairs (4A, 4B) and (4C, 4D)
    const int kcd = 4 * ((8 - fx) * fy) | (4 * (fx * fy)) << 16;
    const uint8_t *sp = src + __ldg(src_off + i);
    uint8_t *dp = dst + __ldg(dst_off + i);
    const int nbytes = w + (fx != 0);
    const int hmax = __reduce_max_sync(0xffffffffu, valid ? h : 0);   // the four operations of a warp may differ in height
    for (int y0 = 0; y0 < hmax; y0 += CH_LANES) {
        const int y = y0 + sub;
        const bool mine = valid && y < h;
        const uint8_t *rowp = sp + (long long)y * stride;
        const ChRow t = chroma_row(rowp, nbytes, mine);
        ChRow b;
        b.r0 = __shfl_down_sync(0xffffffffu, t.r0, 1, CH_LANES); b.r1 = __shfl_down_sync(0xffffffffu, t.r1, 1, CH_LANES);
        b.r2 = __shfl_down_sync(0xffffffffu, t.r2, 1, CH_LANES);
        const bool last = sub == CH_LANES - 1 || y == h - 1;          // the row below is not held by the next lane
        if (last) b = chroma_row(rowp + stride, nbytes, mine && fy != 0);
        if (!mine) continue;
        const unsigned ta = __funnelshift_r(t.r0, t.r1, 8), tb = __funnelshift_r(t.r1, t.r2, 8);
        const unsigned ba = __funnelshift_r(b.r0, b.r1, 8), bb = __funnelshift_r(b.r1, b.r2, 8);
        int v[8];
        v[0] = dp2a_lo_su(kcd, b.r0, dp2a_lo_su(kab, t.r0, 128)); v[1] = dp2a_lo_su(kcd, ba, dp2a_lo_su(kab, ta, 128));
        v[2] = dp2a_hi_su(kcd, b.r0, dp2a_hi_su(kab, t.r0, 128)); v[3] = dp2a_hi_su(kcd, ba, dp2a_hi_su(kab, ta, 128));
        v[4] = dp2a_lo_su(kcd, b.r1, dp2a_lo_su(kab, t.r1, 128)); v[5] = dp2a_lo_su(kcd, bb, dp2a_lo_su(kab, tb, 128));
        v[6] = dp2a_hi_su(kcd, b.r1, dp2a_hi_su(kab, t.r1, 128)); v[7] = dp2a_hi_su(kcd, bb, dp2a_hi_su(kab, tb, 128));
        unsigned w0 = __byte_perm(__byte_perm(v[0], v[1], 0x5151), __byte_perm(v[2], v[3], 0x5151), 0x5410);
        unsigned w1 = __byte_perm(__byte_perm(v[4], v[5], 0x5151), __byte_perm(v[6], v[7], 0x5151), 0x5410);
        uint8_t *d = dp + (long long)y * stride;
        const unsigned al = (unsigned)(reinterpret_cast<uintptr_t>(d) & 7);
        if (w == 8 && al == 0) {
            uint2 *d8 = reinterpret_cast<uint2 *>(d);
            if (avg) { const uint2 pv = *d8; w0 = __vavgu4(pv.x, w0); w1 = __vavgu4(pv.y, w1); }
            *d8 = make_uint2(w0, w1);
        } else if (w >= 4 && (al & 3) == 0) {
            unsigned *d4 = reinterpret_cast<unsigned *>(d);
            d4[0] = avg ? __vavgu4(d4[0], w0) : w0;
            if (w == 8) d4[1] = avg ? __vavgu4(d4[1], w1) : w1;
        } else {
            for (int k = 0; k < w; k++) {
                const int px = (int)__byte_perm(k < 4 ? w0 : w1, 0, 0x4440 | (k & 3));
                d[k] = (uint8_t)(avg ? (d[k] + px + 1) >> 1 : px);
            }
        }
    }
}

// The batched form of chroma_kernel: the same arithmetic, CK operations per 8-lane group with every global load of the CK windows
// (descriptors first, then window rows and — for avg — destination rows) issued before the first result is computed.  chroma_kernel
// has one 9-byte window per 8 lanes in flight and three dependent round trips to memory per operation (ncu: long_scoreboard 38 per
// issue, issue slots 32 % busy); here a warp keeps 4 * CK windows in flight.  Group g of a warp handles operations first + 4 * j + g.
template <int CK, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB)
chroma_kernel_k(long long n, const uint8_t *op, const uint8_t *hh, const uint8_t *xy, uint8_t *dst, const int64_t *dst_off,
                const uint8_t *src, const int64_t *src_off, long long stride)
{
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & (CH_LANES - 1);
    const long long first = ((long long)blockIdx.x * WARPS + (threadIdx.x >> 5)) * (4 * CK);
    if (first >= n) return;
    int o[CK], h[CK], pq[CK];
    long long so[CK], dof[CK];
    bool valid[CK];
#pragma unroll
    for (int j = 0; j < CK; j++) {
        const long long gi = first + 4 * j + grp;
        valid[j] = gi < n;
        const long long i = valid[j] ? gi : n - 1;
        o[j] = __ldg(op + i); h[j] = __ldg(hh + i); pq[j] = __ldg(xy + i); so[j] = __ldg(src_off + i); dof[j] = __ldg(dst_off + i);
    }
    int hm = 0;
#pragma unroll
    for (int j = 0; j < CK; j++) hm = max(hm, valid[j] ? h[j] : 0);
    const int hmax = __reduce_max_sync(0xffffffffu, hm);
    for (int y0 = 0; y0 < hmax; y0 += CH_LANES) {
        const int y = y0 + sub;
        ChRow t[CK], bl[CK];
        uint2 pv[CK];
        bool mine[CK], vec[CK];
        // every load of the CK windows
#pragma unroll
        for (int j = 0; j < CK; j++) {
            const int fx = pq[j] & 7, fy = (pq[j] >> 3) & 7, w = 8 >> ((o[j] >> 1) & 3);
            const int nbytes = w + (fx != 0);
            mine[j] = valid[j] && y < h[j];
            const uint8_t *rowp = src + so[j] + (long long)y * stride;
            t[j] = chroma_row(rowp, nbytes, mine[j]);
            const bool last = sub == CH_LANES - 1 || y == h[j] - 1;      // the row below is not held by the next lane
            bl[j] = chroma_row(rowp + stride, nbytes, mine[j] && last && fy != 0);
            const uint8_t *d = dst + dof[j] + (long long)y * stride;
            vec[j] = w == 8 && (reinterpret_cast<uintptr_t>(d) & 7) == 0;
            pv[j] = make_uint2(0, 0);
            if (mine[j] && vec[j] && (o[j] & 1)) pv[j] = *reinterpret_cast<const uint2 *>(d);
        }
#pragma unroll
        for (int j = 0; j < CK; j++) {
            const int fx = pq[j] & 7, fy = (pq[j] >> 3) & 7, avg = o[j] & 1, w = 8 >> ((o[j] >> 1) & 3);
            const int kab = 4 * ((8 - fx) * (8 - fy)) | (4 * (fx * (8 - fy))) << 16;      // s16 pairs (4A, 4B) and (4C, 4D)
            const int kcd = 4 * ((8 - fx) * fy) | (4 * (fx * fy)) << 16;
            ChRow b;
            b.r0 = __shfl_down_sync(0xffffffffu, t[j].r0, 1, CH_LANES); b.r1 = __shfl_down_sync(0xffffffffu, t[j].r1, 1, CH_LANES);
            b.r2 = __shfl_down_sync(0xffffffffu, t[j].r2, 1, CH_LANES);
            if (sub == CH_LANES - 1 || y == h[j] - 1) b = bl[j];
            if (!mine[j]) continue;
            const ChRow &tt = t[j];
            const unsigned ta = __funnelshift_r(tt.r0, tt.r1, 8), tb = __funnelshift_r(tt.r1, tt.r2, 8);
            const unsigned ba = __funnelshift_r(b.r0, b.r1, 8), bb = __funnelshift_r(b.r1, b.r2, 8);
            int v[8];
            v[0] = dp2a_lo_su(kcd, b.r0, dp2a_lo_su(kab, tt.r0, 128)); v[1] = dp2a_lo_su(kcd, ba, dp2a_lo_su(kab, ta, 128));
            v[2] = dp2a_hi_su(kcd, b.r0, dp2a_hi_su(kab, tt.r0, 128)); v[3] = dp2a_hi_su(kcd, ba, dp2a_hi_su(kab, ta, 128));
            v[4] = dp2a_lo_su(kcd, b.r1, dp2a_lo_su(kab, tt.r1, 128)); v[5] = dp2a_lo_su(kcd, bb, dp2a_lo_su(kab, tb, 128));
            v[6] = dp2a_hi_su(kcd, b.r1, dp2a_hi_su(kab, tt.r1, 128)); v[7] = dp2a_hi_su(kcd, bb, dp2a_hi_su(kab, tb, 128));
            unsigned w0 = __byte_perm(__byte_perm(v[0], v[1], 0x5151), __byte_perm(v[2], v[3], 0x5151), 0x5410);
            unsigned w1 = __byte_perm(__byte_perm(v[4], v[5], 0x5151), __byte_perm(v[6], v[7], 0x5151), 0x5410);
            uint8_t *d = dst + dof[j] + (long long)y * stride;
            if (vec[j]) {
                if (avg) { w0 = __vavgu4(pv[j].x, w0); w1 = __vavgu4(pv[j].y, w1); }
                *reinterpret_cast<uint2 *>(d) = make_uint2(w0, w1);
            } else if (w >= 4 && (reinterpret_cast<uintptr_t>(d) & 3) == 0) {
                unsigned *d4 = reinterpret_cast<unsigned *>(d);
                d4[0] = avg ? __vavgu4(d4[0], w0) : w0;
                if (w == 8) d4[1] = avg ? __vavgu4(d4[1], w1) : w1;
            } else {
                for (int k = 0; k < w; k++) {
                    const int px = (int)__byte_perm(k < 4 ? w0 : w1, 0, 0x4440 | (k & 3));
                    d[k] = (uint8_t)(avg ? (d[k] + px + 1) >> 1 : px);
                }
            }
        }
    }
}

template <int CK, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB)
chroma_kernel_p(long long n, const uint8_t *op, const uint8_t *hh, const uint8_t *xy, uint8_t *dst, const int64_t *dst_off,
                const uint8_t *src, const int64_t *src_off, long long stride)
{
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & (CH_LANES - 1);
    // persistent warps: batch b = 4 * CK consecutive operations; the descriptors of the next batch are requested before this one is
    // worked on, so only the window loads are left on the critical path of a batch
    const long long nwarps = (long long)gridDim.x * WARPS, nbatch = (n + 4 * CK - 1) / (4 * CK);
    long long b = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (b >= nbatch) return;
    int nmeta[CK];
    long long nso[CK], ndof[CK];
    auto fetch = [&](long long bb) {
#pragma unroll
        for (int j = 0; j < CK; j++) {
            const long long gi = bb * (4 * CK) + 4 * j + grp;
            const bool ok = bb < nbatch && gi < n;
            const long long i = ok ? gi : n - 1;
            nmeta[j] = (int)__ldg(op + i) | (int)__ldg(hh + i) << 8 | (int)__ldg(xy + i) << 16 | (ok ? 1 << 24 : 0);
            nso[j] = __ldg(src_off + i); ndof[j] = __ldg(dst_off + i);
        }
    };
    fetch(b);
    for (; b < nbatch; b += nwarps) {
    int o[CK], h[CK], pq[CK];
    long long so[CK], dof[CK];
    bool valid[CK];
#pragma unroll
    for (int j = 0; j < CK; j++) {
        o[j] = nmeta[j] & 255; h[j] = (nmeta[j] >> 8) & 255; pq[j] = (nmeta[j] >> 16) & 255; valid[j] = (nmeta[j] >> 24) & 1;
        so[j] = nso[j]; dof[j] = ndof[j];
    }
    fetch(b + nwarps);
    int hm = 0;
#pragma unroll
    for (int j = 0; j < CK; j++) hm = max(hm, valid[j] ? h[j] : 0);
    const int hmax = __reduce_max_sync(0xffffffffu, hm);
    for (int y0 = 0; y0 < hmax; y0 += CH_LANES) {
        const int y = y0 + sub;
        ChRow t[CK], bl[CK];
        uint2 pv[CK];
        bool mine[CK], vec[CK];
        // every load of the CK windows
#pragma unroll
        for (int j = 0; j < CK; j++) {
            const int fx = pq[j] & 7, fy = (pq[j] >> 3) & 7, w = 8 >> ((o[j] >> 1) & 3);
            const int nbytes = w + (fx != 0);
            mine[j] = valid[j] && y < h[j];
            const uint8_t *rowp = src + so[j] + (long long)y * stride;
            t[j] = chroma_row(rowp, nbytes, mine[j]);
            const bool last = sub == CH_LANES - 1 || y == h[j] - 1;      // the row below is not held by the next lane
            bl[j] = chroma_row(rowp + stride, nbytes, mine[j] && last && fy != 0);
            const uint8_t *d = dst + dof[j] + (long long)y * stride;
            vec[j] = w == 8 && (reinterpret_cast<uintptr_t>(d) & 7) == 0;
            pv[j] = make_uint2(0, 0);
            if (mine[j] && vec[j] && (o[j] & 1)) pv[j] = *reinterpret_cast<const uint2 *>(d);
        }
#pragma unroll
        for (int j = 0; j < CK; j++) {
            const int fx = pq[j] & 7, fy = (pq[j] >> 3) & 7, avg = o[j] & 1, w = 8 >> ((o[j] >> 1) & 3);
            const int kab = 4 * ((8 - fx) * (8 - fy)) | (4 * (fx * (8 - fy))) << 16;      // s16 pairs (4A, 4B) and (4C, 4D)
            const int kcd = 4 * ((8 - fx) * fy) | (4 * (fx * fy)) << 16;
            ChRow b;
            b.r0 = __shfl_down_sync(0xffffffffu, t[j].r0, 1, CH_LANES); b.r1 = __shfl_down_sync(0xffffffffu, t[j].r1, 1, CH_LANES);
            b.r2 = __shfl_down_sync(0xffffffffu, t[j].r2, 1, CH_LANES);
            if (sub == CH_LANES - 1 || y == h[j] - 1) b = bl[j];
            if (!mine[j]) continue;
            const ChRow &tt = t[j];
            const unsigned ta = __funnelshift_r(tt.r0, tt.r1, 8), tb = __funnelshift_r(tt.r1, tt.r2, 8);
            const unsigned ba = __funnelshift_r(b.r0, b.r1, 8), bb = __funnelshift_r(b.r1, b.r2, 8);
            int v[8];
            v[0] = dp2a_lo_su(kcd, b.r0, dp2a_lo_su(kab, tt.r0, 128)); v[1] = dp2a_lo_su(kcd, ba, dp2a_lo_su(kab, ta, 128));
            v[2] = dp2a_hi_su(kcd, b.r0, dp2a_hi_su(kab, tt.r0, 128)); v[3] = dp2a_hi_su(kcd, ba, dp2a_hi_su(kab, ta, 128));
            v[4] = dp2a_lo_su(kcd, b.r1, dp2a_lo_su(kab, tt.r1, 128)); v[5] = dp2a_lo_su(kcd, bb, dp2a_lo_su(kab, tb, 128));
            v[6] = dp2a_hi_su(kcd, b.r1, dp2a_hi_su(kab, tt.r1, 128)); v[7] = dp2a_hi_su(kcd, bb, dp2a_hi_su(kab, tb, 128));
            unsigned w0 = __byte_perm(__byte_perm(v[0], v[1], 0x5151), __byte_perm(v[2], v[3], 0x5151), 0x5410);
            unsigned w1 = __byte_perm(__byte_perm(v[4], v[5], 0x5151), __byte_perm(v[6], v[7], 0x5151), 0x5410);
            uint8_t *d = dst + dof[j] + (long long)y * stride;
            if (vec[j]) {
                if (avg) { w0 = __vavgu4(pv[j].x, w0); w1 = __vavgu4(pv[j].y, w1); }
                *reinterpret_cast<uint2 *>(d) = make_uint2(w0, w1);
            } else if (w >= 4 && (reinterpret_cast<uintptr_t>(d) & 3) == 0) {
                unsigned *d4 = reinterpret_cast<unsigned *>(d);
                d4[0] = avg ? __vavgu4(d4[0], w0) : w0;
                if (w == 8) d4[1] = avg ? __vavgu4(d4[1], w1) : w1;
            } else {
                for (int k = 0; k < w; k++) {
                    const int px = (int)__byte_perm(k < 4 ? w0 : w1, 0, 0x4440 | (k & 3));
                    d[k] = (uint8_t)(avg ? (d[k] + px + 1) >> 1 : px);
                }
            }
        }
    }
    }
}

// ------------------------------------------------------------------------------------------------ h264chroma on TMA-staged windows
// Same arithmetic as chroma_kernel.  Every operation's source window arrives as one cp.async.bulk.tensor.2d box of 32 bytes x
// (h + 1) rows (h rows when y == 0: the row below the block is never touched then, as in the reference) whose origin is the block's
// row and its column rounded down to 16 (the copy engine only takes 16-byte-aligned box columns, see the qpel kernel above): the 9
// window bytes sit at phase ph = x & 15 of each 32-byte tile row, and the row readers fold ph into a word index and one funnel shift.
// A warp owns CQ x 4 operations (8 lanes each, lane = row) and issues all their boxes before it computes the first: 16 windows in
// flight per warp.  Heights other than 2 / 4 / 8 / 16 take plain loads into the same tile.  The box may read up to 31 bytes around
// the reference's own window inside the line (zeros beyond the line pitch).
constexpr int CQ = 4;                   // operations per 8-lane group
constexpr int CTILE = 640;              // 17 rows x 32 bytes, rounded to the 128-byte alignment TMA wants
struct ChromaMaps { CUtensorMap m[8]; };                                // [2 * log2(h / 2) + (y != 0)]: box heights 2,3, 4,5, 8,9, 16,17
struct __align__(128) ChromaTmaSmem {
    uint8_t tile[WARPS][4 * CQ][CTILE];
    unsigned long long mbar[WARPS][4 * CQ];
};

__global__ void __launch_bounds__(32 * WARPS)
chroma_tma_kernel(const __grid_constant__ ChromaMaps maps, long long n, const uint8_t *op, const uint8_t *hh, const uint8_t *xy, uint8_t *dst,
                  const int64_t *dst_off, const uint8_t *src, const int64_t *src_off, long long stride, unsigned long long magic)
{
    __shared__ ChromaTmaSmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
    constexpr int PER_WARP = 4 * CQ;
    const long long first_op = ((long long)blockIdx.x * WARPS + warp) * PER_WARP;
    if (first_op >= n) return;
    // lanes 0 .. 15 own one operation's descriptor each (slot = lane: group lane / CQ, round lane % CQ) and issue its box
    int mo = -1, mh = 0, mpq = 0, inb = 0, bx = 0, by = 0, mph = 0;
    long long msoff = 0, mdoff = 0;
    if (lane < PER_WARP) {
        const uint32_t mb = pel_smem_u32(&sm.mbar[warp][lane]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb) : "memory");
        const long long i = first_op + lane;
        if (i < n) { mo = __ldg(op + i); mh = __ldg(hh + i); mpq = __ldg(xy + i); msoff = __ldg(src_off + i); mdoff = __ldg(dst_off + i); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (mo >= 0 && msoff >= 0 && (mh == 2 || mh == 4 || mh == 8 || mh == 16)) {
        unsigned long long y = __umul64hi((unsigned long long)msoff, magic);
        long long x = msoff - (long long)y * stride;
        while (x >= stride) { x -= stride; y++; }
        if (y < 0x7fffff00ULL) { inb = 1; mph = (int)x & 15; bx = (int)x - mph; by = (int)y; }
    }
    for (int k = 0; k < PER_WARP; k++) {                            // issued by one lane: TMA operands live in uniform registers
        const int kin = __shfl_sync(0xffffffffu, inb, k), kx = __shfl_sync(0xffffffffu, bx, k), ky = __shfl_sync(0xffffffffu, by, k);
        const int kh = __shfl_sync(0xffffffffu, mh, k), kfy = (__shfl_sync(0xffffffffu, mpq, k) >> 3) & 7;
        if (kin && lane == 0) {
            const int rows = kh + (kfy != 0);
            const int mi = 2 * (kh == 2 ? 0 : kh == 4 ? 1 : kh == 8 ? 2 : 3) + (kfy != 0);
            pel_tma_2d(pel_smem_u32(&sm.tile[warp][k][0]), &maps.m[mi], kx, ky, pel_smem_u32(&sm.mbar[warp][k]), 32u * rows);
        }
    }
    for (int r = 0; r < CQ; r++) {
        const int slot = grp * CQ + r;
        const int o = __shfl_sync(0xffffffffu, mo, slot);
        const bool valid = o >= 0;
        const int h = __shfl_sync(0xffffffffu, mh, slot), pq = __shfl_sync(0xffffffffu, mpq, slot);
        const long long doff = __shfl_sync(0xffffffffu, mdoff, slot), soff = __shfl_sync(0xffffffffu, msoff, slot);
        const int tma = __shfl_sync(0xffffffffu, inb, slot);
        const int ph = __shfl_sync(0xffffffffu, mph, slot);                               // 0 for the plain-load fill
        const int fx = pq & 7, fy = (pq >> 3) & 7;
        const int avg = o & 1, w = 8 >> ((o >> 1) & 3);
        const int kab = 4 * ((8 - fx) * (8 - fy)) | (4 * (fx * (8 - fy))) << 16;      // s16 pairs (4A, 4B) and (4C, 4D)
        const int kcd = 4 * ((8 - fx) * fy) | (4 * (fx * fy)) << 16;
        uint8_t *dp = dst + doff;
        uint8_t *tile = sm.tile[warp][slot];
        const int nbytes = w + (fx != 0);
        if (valid && tma) pel_mbar_wait(pel_smem_u32(&sm.mbar[warp][slot]), 0);
        if (!valid) continue;
        for (int y0 = 0; y0 < h; y0 += 16) {
            if (!tma) {                                                                  // plain loads: rows y0 .. y0 + 16 of the window
                __syncwarp(0xffu << (8 * grp));
                const int rows = min(h - y0, 16) + (fy != 0);
                for (int rr = sub; rr < rows; rr += 8) {
                    const uint8_t *rp = src + soff + (long long)(y0 + rr) * stride;
                    for (int k = 0; k < nbytes; k++) tile[rr * 32 + k] = __ldg(rp + k);
                }
                __syncwarp(0xffu << (8 * grp));
            }
#pragma unroll 1
            for (int yy = sub; yy < min(h - y0, 16); yy += 8) {
                uint4 tq, bq;                                                            // .x .y .z = window bytes 0..3, 4..7, 8..11
                {
                    const unsigned *tw = reinterpret_cast<const unsigned *>(tile + yy * 32) + (ph >> 2), *bw = tw + 8;
                    const unsigned sh = (unsigned)(ph & 3) * 8u;
                    const unsigned t0 = tw[0], t1 = tw[1], t2 = tw[2], t3 = tw[3], b0 = bw[0], b1 = bw[1], b2 = bw[2], b3 = bw[3];
                    tq.x = __funnelshift_r(t0, t1, sh); tq.y = __funnelshift_r(t1, t2, sh); tq.z = __funnelshift_r(t2, t3, sh); tq.w = 0;
                    bq.x = __funnelshift_r(b0, b1, sh); bq.y = __funnelshift_r(b1, b2, sh); bq.z = __funnelshift_r(b2, b3, sh); bq.w = 0;
                }
                const unsigned ta = __funnelshift_r(tq.x, tq.y, 8), tb = __funnelshift_r(tq.y, tq.z, 8);
                const unsigned ba = __funnelshift_r(bq.x, bq.y, 8), bb = __funnelshift_r(bq.y, bq.z, 8);
                int v[8];
                v[0] = dp2a_lo_su(kcd, bq.x, dp2a_lo_su(kab, tq.x, 128)); v[1] = dp2a_lo_su(kcd, ba, dp2a_lo_su(kab, ta, 128));
                v[2] = dp2a_hi_su(kcd, bq.x, dp2a_hi_su(kab, tq.x, 128)); v[3] = dp2a_hi_su(kcd, ba, dp2a_hi_su(kab, ta, 128));
                v[4] = dp2a_lo_su(kcd, bq.y, dp2a_lo_su(kab, tq.y, 128)); v[5] = dp2a_lo_su(kcd, bb, dp2a_lo_su(kab, tb, 128));
                v[6] = dp2a_hi_su(kcd, bq.y, dp2a_hi_su(kab, tq.y, 128)); v[7] = dp2a_hi_su(kcd, bb, dp2a_hi_su(kab, tb, 128));
                unsigned w0 = __byte_perm(__byte_perm(v[0], v[1], 0x5151), __byte_perm(v[2], v[3], 0x5151), 0x5410);
                unsigned w1 = __byte_perm(__byte_perm(v[4], v[5], 0x5151), __byte_perm(v[6], v[7], 0x5151), 0x5410);
                uint8_t *d = dp + (long long)(y0 + yy) * stride;
                const unsigned al = (unsigned)(reinterpret_cast<uintptr_t>(d) & 7);
                if (w == 8 && al == 0) {
                    uint2 *d8 = reinterpret_cast<uint2 *>(d);
                    if (avg) { const uint2 pv = *d8; w0 = __vavgu4(pv.x, w0); w1 = __vavgu4(pv.y, w1); }
                    *d8 = make_uint2(w0, w1);
                } else if (w >= 4 && (al & 3) == 0) {
                    unsigned *d4 = reinterpret_cast<unsigned *>(d);
                    d4[0] = avg ? __vavgu4(d4[0], w0) : w0;
                    if (w == 8) d4[1] = avg ? __vavgu4(d4[1], w1) : w1;
                } else {
                    for (int k = 0; k < w; k++) {
                        const int px = (int)__byte_perm(k < 4 ? w0 : w1, 0, 0x4440 | (k & 3));
                        d[k] = (uint8_t)(avg ? (d[k] + px + 1) >> 1 : px);
                    }
                }
            }
        }
    }
}

// emulated_edge_mc (videodsp_template.c:24-101) as a clamped gather: one warp per window, lanes across columns.
// geom[4*i..] = block_w, block_h, src_x, src_y; origin[i] = offset of that picture's sample (0, 0) from `src`.
__global__ void __launch_bounds__(32 * WARPS)
edge_kernel(long long n, uint8_t *buf, const int64_t *buf_off, long long buf_ls, const uint8_t *src, const int64_t *origin,
            long long src_ls, const int32_t *geom, int w, int h)
{
    const long long i = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const int4 gm = __ldg((const int4 *)geom + i);
    const int bw = gm.x, bh = gm.y, sx = gm.z, sy = gm.w;
    const uint8_t *pic = src + origin[i];
    uint8_t *out = buf + buf_off[i];
    for (int x = lane; x < bw; x += 32) {
        const int px = min(max(sx + x, 0), w - 1);
        for (int y = 0; y < bh; y++) {
            const int py = min(max(sy + y, 0), h - 1);
            out[(long long)y * buf_ls + x] = __ldg(pic + (long long)py * src_ls + px);
        }
    }
}

void die(const char *what)
{
    fprintf(stderr, "libb200dsp: motion compensation failed: %s (%s)\n", what, b200_last_error());
    abort();
}

enum HostKind { HOST_QPEL, HOST_HPEL, HOST_CHROMA };

// one block through the device for the drop-in tables.  Only the source rectangle the reference function itself reads is
// copied from the caller's buffer (bx/ax columns left/right, by/ay rows above/below the block).
void host_op(HostKind kind, int o, int h, int w, uint8_t *dst, const uint8_t *src, ptrdiff_t stride,
             int bx, int ax, int by, int ay, int fxy = 0)
{
    B200Device *dev = b200_default_device();
    if (!dev) die("no device");
    if (cudaSetDevice(dev->ordinal) != cudaSuccess) die("cudaSetDevice");
    const int before = kind == HOST_QPEL ? 2 : 0, after = kind == HOST_QPEL ? 3 : 1;   // the window the kernels may touch
    const int sh = h + before + after;
    const size_t pitch = 32;
    B200_LOCK_DEVICE(dev);      // scratch + stream are per device: one host-pointer call at a time (released on return)
    uint8_t *scr = (uint8_t *)b200_scratch(dev, pitch * (sh + h) + 256);
    if (!scr) die("scratch");
    uint8_t *dsrc = scr, *ddst = scr + pitch * sh;
    uint8_t *meta = scr + pitch * (sh + h);           // op, h, offsets
    cudaStream_t st = dev->stream;
    if (b200_h2d_rows(dsrc + (before - by) * pitch + (before - bx), pitch, src - by * stride - bx, stride,
                      w + bx + ax, h + by + ay, st) != cudaSuccess) die("h2d src");
    if (b200_h2d_rows(ddst, pitch, dst, stride, w, h, st) != cudaSuccess) die("h2d dst");
    struct { int64_t doff, soff; uint8_t op, h, xy; } m = { 0, (int64_t)(before * pitch + before), (uint8_t)o, (uint8_t)h, (uint8_t)fxy };
    if (cudaMemcpyAsync(meta, &m, sizeof(m), cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d meta");
    const int64_t *doff = (const int64_t *)meta, *soff = doff + 1;
    const uint8_t *dop = meta + 16, *dh = meta + 17, *dxy = meta + 18;
    // dst and src live in one buffer with the same pitch, like the reference's single stride
    if (kind == HOST_QPEL)      qpel_kernel<true><<<1, 32 * WARPS, 0, st>>>(1, dop, ddst, doff, dsrc, soff, (long long)pitch);
    else if (kind == HOST_HPEL) hpel_kernel<<<1, 32 * WARPS, 0, st>>>(1, dop, dh, ddst, doff, dsrc, soff, (long long)pitch);
    else                        chroma_kernel<<<1, 32 * WARPS, 0, st>>>(1, dop, dh, dxy, ddst, doff, dsrc, soff, (long long)pitch);
    B200_LAUNCHED();
    if (b200_d2h_rows(dst, stride, ddst, pitch, w, h, st) != cudaSuccess) die("d2h");
    if (cudaStreamSynchronize(st) != cudaSuccess) die("sync");
}

template <int AVG, int SIDX, int POS>
void qpel_tab(uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    constexpr int X = POS & 3, Y = POS >> 2;           // h264qpel_template.c: mc{x}{y} filters horizontally iff x, vertically iff y
    host_op(HOST_QPEL, AVG | (SIDX << 1) | (POS << 3), 16 >> SIDX, 16 >> SIDX, dst, src, stride, X ? 2 : 0, X ? 3 : 0, Y ? 2 : 0, Y ? 3 : 0);
}
template <int TAB, int SIDX, int XY>
void hpel_tab(uint8_t *block, const uint8_t *pixels, ptrdiff_t line_size, int h)
{
    host_op(HOST_HPEL, TAB | (SIDX << 2) | (XY << 4), h, 16 >> SIDX, block, pixels, line_size, 0, XY & 1, 0, XY >> 1);
}
template <int AVG, int IDX>
void chroma_tab(uint8_t *dst, const uint8_t *src, ptrdiff_t stride, int h, int x, int y)
{
    if ((unsigned)x > 7 || (unsigned)y > 7) die("h264chroma: x, y must be in 0..7");     // av_assert2 in the reference
    host_op(HOST_CHROMA, AVG | (IDX << 1), h, 8 >> IDX, dst, src, stride, 0, x ? 1 : 0, 0, y ? 1 : 0, x | (y << 3));
}

template <int AVG, int SIDX>
void fill_qpel(b200_qpel_mc_func *t)
{
    t[0] = qpel_tab<AVG, SIDX, 0>;   t[1] = qpel_tab<AVG, SIDX, 1>;   t[2] = qpel_tab<AVG, SIDX, 2>;   t[3] = qpel_tab<AVG, SIDX, 3>;
    t[4] = qpel_tab<AVG, SIDX, 4>;   t[5] = qpel_tab<AVG, SIDX, 5>;   t[6] = qpel_tab<AVG, SIDX, 6>;   t[7] = qpel_tab<AVG, SIDX, 7>;
    t[8] = qpel_tab<AVG, SIDX, 8>;   t[9] = qpel_tab<AVG, SIDX, 9>;   t[10] = qpel_tab<AVG, SIDX, 10>; t[11] = qpel_tab<AVG, SIDX, 11>;
    t[12] = qpel_tab<AVG, SIDX, 12>; t[13] = qpel_tab<AVG, SIDX, 13>; t[14] = qpel_tab<AVG, SIDX, 14>; t[15] = qpel_tab<AVG, SIDX, 15>;
}
template <int TAB, int SIDX>
void fill_hpel(b200_op_pixels_func *t)
{
    t[0] = hpel_tab<TAB, SIDX, 0>; t[1] = hpel_tab<TAB, SIDX, 1>; t[2] = hpel_tab<TAB, SIDX, 2>; t[3] = hpel_tab<TAB, SIDX, 3>;
}

} // namespace

B200_API int b200_h264qpel_init(B200H264QpelContext *c, int bit_depth)
{
    if (!c) return B200_EINVAL;
    if (bit_depth != 8 && bit_depth != 9 && bit_depth != 10 && bit_depth != 12 && bit_depth != 14) return B200_ENOSYS;   // h264qpel.c:87-103
    if (!b200_default_device()) return B200_ENODEV;
    if (bit_depth != 8) { pel_hbd_fill(c, bit_depth); return 0; }  // uint16 samples (pel_hbd.cu)
    fill_qpel<0, 0>(c->put_h264_qpel_pixels_tab[0]); fill_qpel<0, 1>(c->put_h264_qpel_pixels_tab[1]); fill_qpel<0, 2>(c->put_h264_qpel_pixels_tab[2]);
    fill_qpel<1, 0>(c->avg_h264_qpel_pixels_tab[0]); fill_qpel<1, 1>(c->avg_h264_qpel_pixels_tab[1]); fill_qpel<1, 2>(c->avg_h264_qpel_pixels_tab[2]);
    return 0;
}

B200_API int b200_hpeldsp_init(B200HpelDSPContext *c, int flags)
{
    (void)flags;
    if (!c) return B200_EINVAL;
    if (!b200_default_device()) return B200_ENODEV;
    memset(c, 0, sizeof(*c));
    fill_hpel<0, 0>(c->put_pixels_tab[0]); fill_hpel<0, 1>(c->put_pixels_tab[1]); fill_hpel<0, 2>(c->put_pixels_tab[2]); fill_hpel<0, 3>(c->put_pixels_tab[3]);
    fill_hpel<1, 0>(c->avg_pixels_tab[0]); fill_hpel<1, 1>(c->avg_pixels_tab[1]); fill_hpel<1, 2>(c->avg_pixels_tab[2]); fill_hpel<1, 3>(c->avg_pixels_tab[3]);
    fill_hpel<2, 0>(c->put_no_rnd_pixels_tab[0]); fill_hpel<2, 1>(c->put_no_rnd_pixels_tab[1]);
    fill_hpel<3, 0>(c->avg_no_rnd_pixels_tab);
    return 0;
}

namespace {
void edge_tab(uint8_t *buf, const uint8_t *src, ptrdiff_t buf_linesize, ptrdiff_t src_linesize, int block_w, int block_h,
              int src_x, int src_y, int w, int h)
{
    if (!w || !h) return;                                          // videodsp_template.c:33-34
    if (block_w <= 0 || block_h <= 0) return;
    B200Device *dev = b200_default_device();
    if (!dev) die("no device");
    if (cudaSetDevice(dev->ordinal) != cudaSuccess) die("cudaSetDevice");
    // the part of the picture the window can reach after clamping
    const int x0 = std::min(std::max(src_x, 0), w - 1), x1 = std::min(std::max(src_x + block_w - 1, 0), w - 1);
    const int y0 = std::min(std::max(src_y, 0), h - 1), y1 = std::min(std::max(src_y + block_h - 1, 0), h - 1);
    const size_t rw = (size_t)(x1 - x0 + 1), rh = (size_t)(y1 - y0 + 1);
    const size_t rp = (rw + 15) & ~(size_t)15, bp = ((size_t)block_w + 15) & ~(size_t)15;
    B200_LOCK_DEVICE(dev);      // scratch + stream are per device: one host-pointer call at a time (released on return)
    uint8_t *scr = (uint8_t *)b200_scratch(dev, rp * rh + bp * block_h + 256);
    if (!scr) die("scratch");
    uint8_t *drect = scr, *dbuf = scr + rp * rh, *meta = dbuf + bp * block_h;
    meta += (16 - ((uintptr_t)meta & 15)) & 15;
    cudaStream_t st = dev->stream;
    const uint8_t *pic = src - (ptrdiff_t)src_y * src_linesize - src_x;
    if (b200_h2d_rows(drect, rp, pic + (ptrdiff_t)y0 * src_linesize + x0, src_linesize, rw, rh, st) != cudaSuccess) die("h2d");
    struct { int32_t g[4]; int64_t boff, origin; } m = { { block_w, block_h, src_x, src_y }, 0, -((int64_t)y0 * (int64_t)rp + x0) };
    if (cudaMemcpyAsync(meta, &m, sizeof(m), cudaMemcpyHostToDevice, st) != cudaSuccess) die("h2d meta");
    edge_kernel<<<1, 32 * WARPS, 0, st>>>(1, dbuf, (const int64_t *)(meta + 16), (long long)bp, drect, (const int64_t *)(meta + 24),
                                          (long long)rp, (const int32_t *)meta, w, h);
    B200_LAUNCHED();
    if (b200_d2h_rows(buf, buf_linesize, dbuf, bp, (size_t)block_w, (size_t)block_h, st) != cudaSuccess) die("d2h");
    if (cudaStreamSynchronize(st) != cudaSuccess) die("sync");
}
void prefetch_nop(const uint8_t *, ptrdiff_t, int) {}
} // namespace

B200_API int b200_videodsp_init(B200VideoDSPContext *c, int bpc)
{
    if (!c) return B200_EINVAL;
    if (!b200_default_device()) return B200_ENODEV;
    c->emulated_edge_mc = edge_tab;
    if (bpc > 8) pel_hbd_fill_edge(c);                             // videodsp.c:41-45: the 16 bit template above 8
    c->prefetch = prefetch_nop;                                    // videodsp.c:34-36: the C prefetch is an empty function
    return 0;
}

B200_API int b200_emulated_edge_mc_batch_device(B200Device *dev, int64_t n, uint8_t *buf, const int64_t *buf_off,
                                                ptrdiff_t buf_linesize, const uint8_t *src, const int64_t *origin,
                                                ptrdiff_t src_linesize, const int32_t *geom, int w, int h)
{
    if (!dev || n < 0 || !buf || !buf_off || !src || !origin || !geom || w < 0 || h < 0) return B200_EINVAL;
    if (((uintptr_t)geom & 15) != 0) return B200_EINVAL;
    if (n == 0 || !w || !h) return 0;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    const long long blocks = (n + WARPS - 1) / WARPS;
    if (blocks > 0x7fffffffLL) return B200_EINVAL;
    edge_kernel<<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, buf, buf_off, buf_linesize, src, origin, src_linesize, geom, w, h);
    B200_LAUNCHED();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

B200_API int b200_h264chroma_init(B200H264ChromaContext *c, int bit_depth)
{
    if (!c) return B200_EINVAL;
    if (!b200_default_device()) return B200_ENODEV;
    if (bit_depth > 8 && bit_depth <= 16) { pel_hbd_fill_chroma(c); return 0; }       // h264chroma.c:45-50: the 16 bit template (pel_hbd.cu)
    memset(c, 0, sizeof(*c));                                      // entry [3] stays NULL like the reference's
    c->put_h264_chroma_pixels_tab[0] = chroma_tab<0, 0>; c->put_h264_chroma_pixels_tab[1] = chroma_tab<0, 1>; c->put_h264_chroma_pixels_tab[2] = chroma_tab<0, 2>;
    c->avg_h264_chroma_pixels_tab[0] = chroma_tab<1, 0>; c->avg_h264_chroma_pixels_tab[1] = chroma_tab<1, 1>; c->avg_h264_chroma_pixels_tab[2] = chroma_tab<1, 2>;
    return 0;
}

B200_API int b200_h264chroma_batch_device(B200Device *dev, int64_t n, const uint8_t *op, const uint8_t *h, const uint8_t *xy,
                                          uint8_t *dst, const int64_t *dst_off, const uint8_t *src, const int64_t *src_off,
                                          ptrdiff_t stride)
{
    if (!dev || n < 0 || !op || !h || !xy || !dst || !dst_off || !src || !src_off) return B200_EINVAL;
    if (n == 0) return 0;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    static int mode = -1;                                         // B200_CHROMA_TMA: 0 (default) batched LDG kernel, 1 TMA boxes, 2 one-window LDG kernel
    if (mode < 0) { const char *e = getenv("B200_CHROMA_TMA"); mode = e ? atoi(e) : 0; }
    ChromaMaps maps;
    bool tma = mode == 1 && stride >= 32;
    for (int k = 0; k < 8 && tma; k++)
        tma = b200_tmap_2d_u8(&maps.m[k], src, (unsigned long long)stride, 32, (2u << (k >> 1)) + (k & 1), (int)CU_TENSOR_MAP_SWIZZLE_NONE);
    if (tma) {
        const long long blocks = (n + WARPS * 4 * CQ - 1) / (WARPS * 4 * CQ);
        if (blocks > 0x7fffffffLL) return B200_EINVAL;
        chroma_tma_kernel<<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(maps, n, op, h, xy, dst, dst_off, src, src_off, stride,
                                                                            ~0ULL / (unsigned long long)stride);
    } else if (mode == 0) {
        static int var = -1;                                      // tuning knob: windows per 8-lane group / register budget
        // measured (scripts/quick_bench.py chroma, fraction of the HBM roofline): one-shot warps 0 -> 0.239, 1 -> 0.247, 2 -> 0.253, 3 -> 0.177;
        // persistent warps that request the next batch's descriptors before working on the current one: <4,4> (121 registers) 0.264,
        // <2,8> (64 registers, no spills) 0.277 = the default; every persistent shape that spills (<4,8>, <4,6>, <3,8>, <2,10>, <2,12>) loses
        if (var < 0) { const char *e = getenv("B200_CHROMA_VAR"); var = e ? atoi(e) : 7; }
        const int ck = var == 1 ? 2 : var == 3 ? 8 : 4;
        const int per_cta = WARPS * 4 * ck;
        const long long blocks = (n + per_cta - 1) / per_cta;
        if (blocks > 0x7fffffffLL) return B200_EINVAL;
        if (var == 6 || var == 7) {                                // persistent warps with the next batch's descriptors in flight
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev->ordinal);
            const int ckp = var == 7 ? 2 : 4, minb = var == 7 ? 8 : 4;
            const long long nbatch = (n + 4 * ckp - 1) / (4 * ckp), want = (nbatch + WARPS - 1) / WARPS;
            const long long cap = (long long)sms * minb;
            const unsigned g = (unsigned)(want < cap ? want : cap);
            if (var == 6) chroma_kernel_p<4, 4><<<g, 32 * WARPS, 0, dev->stream>>>(n, op, h, xy, dst, dst_off, src, src_off, stride);
            else          chroma_kernel_p<2, 8><<<g, 32 * WARPS, 0, dev->stream>>>(n, op, h, xy, dst, dst_off, src, src_off, stride);
        }
        else if (var == 1) chroma_kernel_k<2, 12><<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, op, h, xy, dst, dst_off, src, src_off, stride);
        else if (var == 2) chroma_kernel_k<4, 8><<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, op, h, xy, dst, dst_off, src, src_off, stride);
        else if (var == 3) chroma_kernel_k<8, 2><<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, op, h, xy, dst, dst_off, src, src_off, stride);
        else               chroma_kernel_k<4, 4><<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, op, h, xy, dst, dst_off, src, src_off, stride);
    } else {
        const int per_cta = 32 * WARPS / CH_LANES;
        const long long blocks = (n + per_cta - 1) / per_cta;
        if (blocks > 0x7fffffffLL) return B200_EINVAL;
        chroma_kernel<<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, op, h, xy, dst, dst_off, src, src_off, stride);
    }
    B200_LAUNCHED();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// B200_QPEL_TMA: 0 (default) = LDG-staged kernel, 1 = TMA boxes.  Measured on the B200 (profiles/r02_qpel_tma_ncu.txt): the TMA form is
// bit-exact but slower (2.12 vs 2.25 G blocks/s): one 48 x 21-byte box per operation keeps the copy engine at about one box per 150
// cycles per SM, and the smaller boxes of the chroma kernel at one per 80 — small 2-D boxes are bound by the engine's per-row requests,
// not by the LSU they bypass.
static int qpel_launch(cudaStream_t st, int64_t n, const uint8_t *op, uint8_t *dst, const int64_t *dst_off, const uint8_t *src,
                       const int64_t *src_off, ptrdiff_t stride)
{
    static int mode = -1;
    if (mode < 0) { const char *e = getenv("B200_QPEL_TMA"); mode = e ? atoi(e) : 0; }
    CUtensorMap tm16, tm8, tm4;
    const int swz = (int)CU_TENSOR_MAP_SWIZZLE_NONE;
    if (mode > 0 && stride >= 48 && b200_tmap_2d_u8(&tm16, src, (unsigned long long)stride, 48, 21, swz) &&
        b200_tmap_2d_u8(&tm8, src, (unsigned long long)stride, 32, 13, swz) && b200_tmap_2d_u8(&tm4, src, (unsigned long long)stride, 32, 9, swz)) {
        const long long blocks = (n + WARPS * TQK - 1) / (WARPS * TQK);
        if (blocks > 0x7fffffffLL) return B200_EINVAL;
        const unsigned long long magic = ~0ULL / (unsigned long long)stride;              // floor((2^64 - 1) / stride): the quotient estimate is never too large
        qpel_tma_kernel<<<(unsigned)blocks, 32 * WARPS, 0, st>>>(tm16, tm8, tm4, n, op, dst, dst_off, src, src_off, stride, magic);
    } else {
        const long long blocks = (n + WARPS * QK - 1) / (WARPS * QK);
        if (blocks > 0x7fffffffLL) return B200_EINVAL;
        // B200_QPEL_MMA=1: 16 x 16 blocks on the tensor cores.  Measured on the B200 (64 x 1080p frames of random quarter-pel blocks,
        // profiles/r02_qpel_mma_ncu.txt / r02_qpel_dp_ncu.txt): bit-exact, 2.32 G blocks/s against 2.42 G for the dot-product path — the
        // filter arithmetic is 94 of ~330 instructions per block on either path, the rest is per-block scaffolding (descriptor shuffles,
        // window re-alignment, destination addressing, stores) — so the dot-product path stays the default.
        static int mma = -1;
        if (mma < 0) { const char *e = getenv("B200_QPEL_MMA"); mma = e ? atoi(e) : 0; }
        if (mma) qpel_kernel<true><<<(unsigned)blocks, 32 * WARPS, 0, st>>>(n, op, dst, dst_off, src, src_off, stride);
        else     qpel_kernel<false><<<(unsigned)blocks, 32 * WARPS, 0, st>>>(n, op, dst, dst_off, src, src_off, stride);
    }
    B200_LAUNCHED();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

B200_API int b200_h264qpel_batch_device(B200Device *dev, int64_t n, const uint8_t *op, uint8_t *dst, const int64_t *dst_off,
                                        const uint8_t *src, const int64_t *src_off, ptrdiff_t stride)
{
    if (!dev || n < 0 || !op || !dst || !dst_off || !src || !src_off) return B200_EINVAL;
    if (n == 0) return 0;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    return qpel_launch(dev->stream, n, op, dst, dst_off, src, src_off, stride);
}

// HOST buffers: a stream of frames, each with its own reference picture, destination picture and list of motion-compensation
// operations (what a decoder holding its pictures in host memory hands over once per frame).  Frame f occupies
// [f * frame_bytes, (f + 1) * frame_bytes) of `src` and of `dst`; its operations are op_begin[f] .. op_begin[f + 1] - 1, with offsets
// counted from the start of `dst` / `src` like the device entry point, and must stay inside frame f.  Frames are cut into chunks that
// rotate over the device's three pipeline streams: H2D of the chunk's reference and destination pictures and operation lists, the
// kernel, D2H of the destination pictures.
B200_API int b200_h264qpel_frames_host(B200Device *dev, int nframes, int64_t frame_bytes, const int64_t *op_begin, const uint8_t *op,
                                       uint8_t *dst, const int64_t *dst_off, const uint8_t *src, const int64_t *src_off, ptrdiff_t stride)
{
    if (!dev || nframes < 0 || frame_bytes <= 0 || !op_begin || !op || !dst || !dst_off || !src || !src_off) return B200_EINVAL;
    if (nframes == 0) return 0;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    int64_t ops_total = op_begin[nframes] - op_begin[0];
    for (int f = 0; f < nframes; f++) if (op_begin[f + 1] < op_begin[f]) return B200_EINVAL;
    const int64_t ops_per_frame = (ops_total + nframes - 1) / nframes;
    const size_t perFrame = 2 * (size_t)frame_bytes + (size_t)ops_per_frame * 17 + 64;
    int chunk = (int)(((size_t)64 << 20) / perFrame);
    if (chunk < 1) chunk = 1;
    if (chunk > nframes) chunk = nframes;
    int64_t max_ops = 0;                                            // operations of the busiest chunk
    for (int f0 = 0; f0 < nframes; f0 += chunk) max_ops = std::max(max_ops, op_begin[std::min(nframes, f0 + chunk)] - op_begin[f0]);
    const size_t n8 = (((size_t)max_ops * 8) + 255) & ~(size_t)255;
    const size_t pics = ((2 * (size_t)frame_bytes * chunk) + 255) & ~(size_t)255;
    const size_t slotBytes = pics + 2 * n8 + (((size_t)max_ops + 255) & ~(size_t)255);
    const int K = B200Device::kPipe;
    B200_LOCK_DEVICE(dev);
    uint8_t *scr = (uint8_t *)b200_scratch(dev, slotBytes * K);
    if (!scr) return B200_ENOMEM;
    B200_CUDA_OK(cudaStreamSynchronize(dev->stream));
    int slot = 0;
    for (int f0 = 0; f0 < nframes; f0 += chunk, slot = (slot + 1) % K) {
        const int nf = nframes - f0 < chunk ? nframes - f0 : chunk;
        cudaStream_t st = dev->pipe[slot];
        uint8_t *base = scr + (size_t)slot * slotBytes;
        uint8_t *dsrc = base, *ddst = base + (size_t)frame_bytes * chunk;
        int64_t *ddo = (int64_t *)(base + pics), *dso = (int64_t *)(base + pics + n8);
        uint8_t *dop = base + pics + 2 * n8;
        const int64_t o0 = op_begin[f0], n = op_begin[f0 + nf] - o0;
        B200_CUDA_OK(cudaMemcpyAsync(dsrc, src + (size_t)f0 * frame_bytes, (size_t)nf * frame_bytes, cudaMemcpyHostToDevice, st));
        B200_CUDA_OK(cudaMemcpyAsync(ddst, dst + (size_t)f0 * frame_bytes, (size_t)nf * frame_bytes, cudaMemcpyHostToDevice, st));
        if (n > 0) {
            B200_CUDA_OK(cudaMemcpyAsync(ddo, dst_off + o0, (size_t)n * 8, cudaMemcpyHostToDevice, st));
            B200_CUDA_OK(cudaMemcpyAsync(dso, src_off + o0, (size_t)n * 8, cudaMemcpyHostToDevice, st));
            B200_CUDA_OK(cudaMemcpyAsync(dop, op + o0, (size_t)n, cudaMemcpyHostToDevice, st));
            const long long blocks = (n + WARPS * QK - 1) / (WARPS * QK);
            if (blocks > 0x7fffffffLL) return B200_EINVAL;
            // the lists keep their offsets from the start of the whole buffers: hand the kernel bases moved back by the chunk's origin
            qpel_kernel<true><<<(unsigned)blocks, 32 * WARPS, 0, st>>>(n, dop, ddst - (size_t)f0 * frame_bytes, ddo,
                                                                 dsrc - (size_t)f0 * frame_bytes, dso, stride);
            B200_LAUNCHED();
            B200_CUDA_OK(cudaGetLastError());
        }
        B200_CUDA_OK(cudaMemcpyAsync(dst + (size_t)f0 * frame_bytes, ddst, (size_t)nf * frame_bytes, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < K; i++) B200_CUDA_OK(cudaStreamSynchronize(dev->pipe[i]));
    return 0;
}

B200_API int b200_hpel_batch_device(B200Device *dev, int64_t n, const uint8_t *op, const uint8_t *h, uint8_t *dst,
                                    const int64_t *dst_off, const uint8_t *src, const int64_t *src_off, ptrdiff_t stride)
{
    if (!dev || n < 0 || !op || !h || !dst || !dst_off || !src || !src_off) return B200_EINVAL;
    if (n == 0) return 0;
    B200_CUDA_OK(cudaSetDevice(dev->ordinal));
    const long long blocks = (n + WARPS - 1) / WARPS;
    if (blocks > 0x7fffffffLL) return B200_EINVAL;
    hpel_kernel<<<(unsigned)blocks, 32 * WARPS, 0, dev->stream>>>(n, op, h, dst, dst_off, src, src_off, stride);
    B200_LAUNCHED();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}
