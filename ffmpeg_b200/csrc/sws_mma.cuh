// sws_mma.cuh — the separable scaler with its horizontal pass on the tensor cores (included by sws.cu after its writer helpers).
//
// hScale8To15_c (libswscale/swscale.c:128-142) is  dst[i] = min((sum_j src[pos[i] + j] * filter[i][j]) >> 7, 32767):  for a group of
// eight neighbouring output columns the taps reach a short run of source columns (22 bytes at 2:1 with 8 taps), so the group is ONE
// dense 16 x 32 (source lines x source columns, u8) by 32 x 8 (source columns x output columns) integer contraction:
// mma.sync.m16n8k32 with the group's banded coefficient matrix as the B operand.  The 14-bit coefficients are split
// c = 256 * hi + lo (hi signed, lo unsigned byte): one u8 x s8 and one u8 x u8 MMA per 16 lines x 8 columns, exact in int32, combined as
// (hi << 8) + lo before the reference's >> 7 and clamp.  Longer filters / larger ratios take more 32-column chunks per group.
// The B fragments are laid out on the host in register order (sws_mma_build), so a lane loads its 16 bytes per chunk once per tile.
//
// A CTA owns MT_W output columns x TR output lines of one plane (or of luma + both chroma planes for packed RGB output):
//   1. cp.async the source lines the tile's vertical taps reach, columns [c0, c0 + 16 * nseg), into shared memory (pitch = 16 x odd
//      bytes: the A-fragment loads of a warp — 8 lines x 16 bytes — fall into 32 different banks);
//   2. tensor-core horizontal pass -> 15-bit lines in shared memory, kept as 32-bit integers (the vertical pass then multiplies what it
//      loads: with packed 16-bit pairs every sample costs an extra sign-extending PRMT / SHF, and that pass is issue-bound);
//   3. vertical FIR + writer out of shared memory (same arithmetic as sws_fused_plane_kernel / sws_fused_rgb_kernel).
// Source bytes are read from HBM once (plus the vertical halo), the scaled lines never leave the SM.
constexpr int MT_W = 128;                       // output columns per tile
constexpr int MT_TPB = MT_W * 4 + 32;           // byte pitch of the scaled lines (int32) of a 128-column tile: the fragment stores of a
                                                // half warp (4 lines x 32 bytes) fall into different banks
constexpr int MT_CPB = MT_W * 2 + 32;           // same for the 64-column chroma tiles of the RGB kernel
constexpr int MT_THREADS = 256;
constexpr int MT_HDR = 16;                      // bytes in front of the staged lines (tile span words)

__device__ __forceinline__ void mma_u8s8(int (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_u8u8(int (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mt_cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mt_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Source columns the groups [G0, G0 + ng) need: first column rounded down to 16 and the number of 16-byte segments (warp 0 computes,
// everyone reads after the barrier the caller places).
__device__ __forceinline__ void mt_tile_span(const SwsMmaBank &hb, int G0, int ng, int *out /* shared: c0, nseg */)
{
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int s = 0x7fffffff, e = 0;
        if (lane < ng) {
            const int2 gi = __ldg(hb.ginfo + G0 + lane);
            s = gi.x;
            e = gi.x + 32 * (__ldg(hb.ginfo + G0 + lane + 1).y - gi.y);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { s = min(s, __shfl_xor_sync(0xffffffffu, s, o)); e = max(e, __shfl_xor_sync(0xffffffffu, e, o)); }
        if (lane == 0) { out[0] = s & ~15; out[1] = (e - (s & ~15) + 15) >> 4; }
    }
}

// lines [ra, ra + nrows) of `plane`, columns [c0, c0 + 16 * nseg) -> ssrc (pitch SP).  Columns at or beyond srcW hold zeros
// (they only ever meet zero coefficients).
__device__ __forceinline__ void mt_stage(uint8_t *ssrc, int SP, const uint8_t *plane, long long sstride, int ra, int nrows, int c0, int nseg, int srcW)
{
    // a warp per line, a lane per 16-byte segment (19 of 32 lanes busy at 2:1); whether a lane's segment is a plain 16-byte copy does
    // not depend on the line, so the line loop is a pointer bump and one LDGSTS
    const int srcW16 = srcW & ~15, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int s = lane; s < nseg; s += 32) {
        const int col = c0 + 16 * s;
        const uint8_t *g = plane + (long long)(ra + w) * sstride + col;
        uint8_t *d = ssrc + w * SP + 16 * s;
        if (col < srcW16) {
            for (int r = w; r < nrows; r += MT_THREADS / 32, g += (MT_THREADS / 32) * sstride, d += (MT_THREADS / 32) * SP) mt_cp_async16(d, g);
        } else {
            for (int r = w; r < nrows; r += MT_THREADS / 32, g += (MT_THREADS / 32) * sstride, d += (MT_THREADS / 32) * SP)
                for (int b = 0; b < 16; b++) d[b] = col + b < srcW ? g[b] : (uint8_t)0;
        }
    }
}

template <int NCH>              // chunks of 32 source columns per group: 1, 2, or 0 = any number
__device__ __forceinline__ void mt_hgroup_n(const uint8_t *arow0, int SP, const uint4 *bf, int nch, int nrb, uint8_t *tdst, int TPB)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const uint4 B0 = __ldg(bf + lane);
    uint4 B1 = make_uint4(0, 0, 0, 0);
    if (NCH != 1) B1 = __ldg(bf + 32 + lane);
    const uint8_t *arow = arow0 + g * SP + 4 * t;
    uint8_t *trow = tdst + g * TPB + 8 * t;
    const int SP8 = 8 * SP;
    for (int rb = 0; rb < nrb; rb++, arow += 2 * SP8, trow += 16 * TPB) {
        int hi[4] = { 0, 0, 0, 0 }, lo[4] = { 0, 0, 0, 0 };
        unsigned a[4];
        a[0] = *reinterpret_cast<const unsigned *>(arow);          a[1] = *reinterpret_cast<const unsigned *>(arow + SP8);
        a[2] = *reinterpret_cast<const unsigned *>(arow + 16);     a[3] = *reinterpret_cast<const unsigned *>(arow + SP8 + 16);
        mma_u8s8(hi, a, B0.x, B0.y);
        mma_u8u8(lo, a, B0.z, B0.w);
        if (NCH != 1) {
            a[0] = *reinterpret_cast<const unsigned *>(arow + 32);  a[1] = *reinterpret_cast<const unsigned *>(arow + SP8 + 32);
            a[2] = *reinterpret_cast<const unsigned *>(arow + 48);  a[3] = *reinterpret_cast<const unsigned *>(arow + SP8 + 48);
            mma_u8s8(hi, a, B1.x, B1.y);
            mma_u8u8(lo, a, B1.z, B1.w);
            if (NCH == 0)
                for (int c = 2; c < nch; c++) {
                    const uint4 Bc = __ldg(bf + 32 * c + lane);
                    a[0] = *reinterpret_cast<const unsigned *>(arow + 32 * c);       a[1] = *reinterpret_cast<const unsigned *>(arow + SP8 + 32 * c);
                    a[2] = *reinterpret_cast<const unsigned *>(arow + 32 * c + 16);  a[3] = *reinterpret_cast<const unsigned *>(arow + SP8 + 32 * c + 16);
                    mma_u8s8(hi, a, Bc.x, Bc.y);
                    mma_u8u8(lo, a, Bc.z, Bc.w);
                }
        }
        int v[4];                                                  // (256 * hi + lo) >> 7 = 2 * hi + (lo >> 7): lo >= 0
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = min(2 * hi[k] + (lo[k] >> 7), 32767);
        *reinterpret_cast<int2 *>(trow) = make_int2(v[0], v[1]);                   // line g,     columns 2t, 2t + 1
        *reinterpret_cast<int2 *>(trow + 8 * TPB) = make_int2(v[2], v[3]);         // line g + 8
    }
}
// Horizontal pass of one group of 8 output columns over nrb blocks of 16 staged lines.
//   arow0: staged line 0 at the group's window (ssrc + kstart - c0); tdst: scaled line 0 at the group's first column; TPB: its byte pitch
__device__ __forceinline__ void mt_hgroup(const uint8_t *arow0, int SP, const uint4 *bf, int nch, int nrb, uint8_t *tdst, int TPB)
{
    if (nch == 1)      mt_hgroup_n<1>(arow0, SP, bf, nch, nrb, tdst, TPB);
    else if (nch == 2) mt_hgroup_n<2>(arow0, SP, bf, nch, nrb, tdst, TPB);
    else               mt_hgroup_n<0>(arow0, SP, bf, nch, nrb, tdst, TPB);
}

// one 8-bit plane -> one 8-bit plane (yuv2planeX_8_c / yuv2plane1_8_c with the flat dither of SWS_BITEXACT, output.c:468-493)
__global__ void __launch_bounds__(MT_THREADS)
sws_mma_plane_kernel(const uint8_t *src, long long sstride, long long sfs, int srcW, int srcH, uint8_t *dst, long long ds, long long dfs,
                     int dstW, int dstH, SwsMmaBank hb, const int32_t *vcoef2, const int32_t *vpos, int vfs, int TR, int rows_cap, int SP)
{
    extern __shared__ __align__(16) uint8_t mt_smem[];
    int *span = reinterpret_cast<int *>(mt_smem);                       // first 16 bytes: the tile's source column span
    uint8_t *ssrc = mt_smem + MT_HDR, *lines = ssrc + (size_t)rows_cap * SP;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int x0 = blockIdx.x * MT_W, G0 = x0 >> 3, ng = min(MT_W / 8, hb.ngroups - G0);
    const int dy0 = blockIdx.y * TR, dy1 = min(dy0 + TR, dstH) - 1;
    const long long f = blockIdx.z;
    const int ra = min(max(max(1 - vfs, __ldg(vpos + dy0)), 0), srcH - 1);
    const int rb = min(max(max(1 - vfs, __ldg(vpos + dy1)) + vfs - 1, 0), srcH - 1);
    const int nsrc = rb - ra + 1, nrb = (nsrc + 15) >> 4;
    mt_tile_span(hb, G0, ng, span);
    __syncthreads();
    const int c0 = span[0], nseg = span[1];
    mt_stage(ssrc, SP, src + f * sfs, sstride, ra, nsrc, c0, nseg, srcW);
    mt_cp_async_wait();
    __syncthreads();
    for (int gq = warp; gq < ng; gq += MT_THREADS / 32) {
        const int2 gi = __ldg(hb.ginfo + G0 + gq);
        const int nch = __ldg(hb.ginfo + G0 + gq + 1).y - gi.y;
        mt_hgroup(ssrc + (gi.x - c0), SP, hb.bfrag + (size_t)gi.y * 32, nch, nrb, lines + gq * 32, MT_TPB);
    }
    __syncthreads();
    // vertical pass: an item = one output line x 8 columns, taken as columns 4j .. 4j+3 and 64+4j .. 64+4j+3 of the tile: the two
    // 16-byte loads of a tap are then contiguous across the 16 lanes of a line (no bank conflicts), and so are the two 4-byte stores
    const int nrows = dy1 - dy0 + 1;
    for (int it = tid; it < nrows * (MT_W / 8); it += MT_THREADS) {
        const int row = it / (MT_W / 8), j4 = it - row * (MT_W / 8), xa = x0 + j4 * 4, xb = xa + MT_W / 2;
        if (xa >= dstW) continue;
        const int dy = dy0 + row;
        const int first = max(1 - vfs, __ldg(vpos + dy));
        const int oa = j4 * 16, ob = oa + MT_W * 2;                             // byte offsets of the two column quads inside a scaled line
        int v[8];
        if (vfs == 1) {
            const uint8_t *lp = lines + (min(max(first, 0), srcH - 1) - ra) * MT_TPB;
            const int4 q0 = *reinterpret_cast<const int4 *>(lp + oa), q1 = *reinterpret_cast<const int4 *>(lp + ob);
            const int ww[8] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = (ww[j] + 64) >> 7;
        } else {
            unsigned a[8];
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = 64u << 12;
            const int32_t *k2 = vcoef2 + (long long)dy * ((vfs + 1) >> 1);
            const bool inside = first >= 0 && first + vfs <= srcH;                 // no line of this window is clamped
            for (int t = 0; t < vfs; t++) {
                const int cc = __ldg(k2 + (t >> 1));
                const int c = (t & 1) ? cc >> 16 : (int)(short)cc;
                const uint8_t *lp = lines + (inside ? first + t - ra : min(max(first + t, 0), srcH - 1) - ra) * MT_TPB;
                const int4 p0 = *reinterpret_cast<const int4 *>(lp + oa), p1 = *reinterpret_cast<const int4 *>(lp + ob);
                const int w0[8] = { p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w };
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] += (unsigned)(w0[j] * c);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = (int)a[j] >> 19;
        }
        unsigned o[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const unsigned lo = __vimin_s16x2_relu(__byte_perm((unsigned)v[4 * h], (unsigned)v[4 * h + 1], 0x5410), 0x00ff00ffu);
            const unsigned hi = __vimin_s16x2_relu(__byte_perm((unsigned)v[4 * h + 2], (unsigned)v[4 * h + 3], 0x5410), 0x00ff00ffu);
            o[h] = __byte_perm(lo, hi, 0x6420);
        }
        uint8_t *dl = dst + f * dfs + (long long)dy * ds;
        *reinterpret_cast<unsigned *>(dl + xa) = o[0];
        if (xb < dstW) *reinterpret_cast<unsigned *>(dl + xb) = o[1];
    }
}

// yuv420p -> packed RGB through the `_X` writer (yuv2rgb_X_c_template, output.c:1789-1840): luma tile MT_W x TR, MT_W / 2 columns of both
// chroma planes.  Shared memory: staged luma lines, staged U lines, staged V lines, then the 15-bit lines of the three planes.
template <int KIND>
__global__ void __launch_bounds__(MT_THREADS)
sws_mma_rgb_kernel(SwsFrameArgs a, SwsDevTables t, SwsColorConst c, SwsMmaBank hl, SwsMmaBank hc, int srcW, int chrSrcW,
                   int TR, int rowsL, int rowsC, int SPL, int SPC)
{
    constexpr int PW = OutWords<KIND>::per_pair, BPP = OutWords<KIND>::bpp, CG = MT_W / 16;     // chroma groups per tile
    extern __shared__ __align__(16) uint8_t mt_smem[];
    int *spanL = reinterpret_cast<int *>(mt_smem), *spanC = spanL + 2;  // first 16 bytes: the tile's source column spans
    uint8_t *srcL = mt_smem + MT_HDR, *srcU = srcL + (size_t)rowsL * SPL, *srcV = srcU + (size_t)rowsC * SPC;
    uint8_t *linL = srcV + (size_t)rowsC * SPC, *linU = linL + (size_t)rowsL * MT_TPB, *linV = linU + (size_t)rowsC * MT_CPB;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int x0 = blockIdx.x * MT_W, GL0 = x0 >> 3, ngl = min(MT_W / 8, hl.ngroups - GL0);
    const int GC0 = x0 >> 4, ngc = min(CG, hc.ngroups - GC0);
    const int dy0 = blockIdx.y * TR, dy1 = min(dy0 + TR, a.dstH) - 1;
    const long long f = blockIdx.z;
    const int lfs = t.vLumSize, cfs = t.vChrSize;
    const int la = min(max(max(1 - lfs, __ldg(t.vLumPos + dy0)), 0), a.srcH - 1);
    const int lb = min(max(max(1 - lfs, __ldg(t.vLumPos + dy1)) + lfs - 1, 0), a.srcH - 1);
    const int ca = min(max(max(1 - cfs, __ldg(t.vChrPos + dy0)), 0), a.chrSrcH - 1);
    const int cb_ = min(max(max(1 - cfs, __ldg(t.vChrPos + dy1)) + cfs - 1, 0), a.chrSrcH - 1);
    const int nl = lb - la + 1, nc = cb_ - ca + 1;
    mt_tile_span(hl, GL0, ngl, spanL);
    mt_tile_span(hc, GC0, ngc, spanC);
    __syncthreads();
    mt_stage(srcL, SPL, a.y + f * a.yfs, a.ys, la, nl, spanL[0], spanL[1], srcW);
    mt_stage(srcU, SPC, a.u + f * a.ufs, a.us, ca, nc, spanC[0], spanC[1], chrSrcW);
    mt_stage(srcV, SPC, a.v + f * a.vfs, a.vs, ca, nc, spanC[0], spanC[1], chrSrcW);
    mt_cp_async_wait();
    __syncthreads();
    for (int gq = warp; gq < ngl; gq += MT_THREADS / 32) {
        const int2 gi = __ldg(hl.ginfo + GL0 + gq);
        const int nch = __ldg(hl.ginfo + GL0 + gq + 1).y - gi.y;
        mt_hgroup(srcL + (gi.x - spanL[0]), SPL, hl.bfrag + (size_t)gi.y * 32, nch, (nl + 15) >> 4, linL + gq * 32, MT_TPB);
    }
    for (int gq = warp; gq < 2 * ngc; gq += MT_THREADS / 32) {
        const int pl = gq >= ngc, gg = gq - pl * ngc;
        const int2 gi = __ldg(hc.ginfo + GC0 + gg);
        const int nch = __ldg(hc.ginfo + GC0 + gg + 1).y - gi.y;
        mt_hgroup((pl ? srcV : srcU) + (gi.x - spanC[0]), SPC, hc.bfrag + (size_t)gi.y * 32, nch, (nc + 15) >> 4, (pl ? linV : linU) + gg * 32, MT_CPB);
    }
    __syncthreads();
    // vertical pass.  A lane first works on luma columns 4j .. 4j+3 and 64+4j .. 64+4j+3 (chroma 2j, 2j+1 and 32+2j, 32+2j+1) of one
    // output line, so that the loads of a tap are contiguous across the 16 lanes of the line (no bank conflicts); neighbouring lanes then
    // swap one quad each and every lane ends with 8 consecutive pixels for the writer (even lanes the left half's, odd lanes the right half's)
    const int nrows = dy1 - dy0 + 1, nitems = nrows * (MT_W / 8);
    for (int it0 = 0; it0 < nitems; it0 += MT_THREADS) {
        const int it = it0 + tid;
        const int row = min(it, nitems - 1) / (MT_W / 8), j4 = it & (MT_W / 8 - 1), odd = j4 & 1;
        const int x = x0 + (j4 >> 1) * 8 + odd * (MT_W / 2);
        const bool active = it < nitems && x < a.dstW;
        const int dy = dy0 + row;
        const int firstLum = max(1 - lfs, __ldg(t.vLumPos + dy));
        const int firstChr = max(1 - cfs, __ldg(t.vChrPos + dy));
        unsigned sY[8], sU[4], sV[4];
#pragma unroll
        for (int i = 0; i < 8; i++) sY[i] = 1u << 18;
#pragma unroll
        for (int i = 0; i < 4; i++) sU[i] = sV[i] = 1u << 18;
        {
            const int32_t *k2 = t.vLum2 + (long long)dy * ((lfs + 1) >> 1);
            const bool inside = firstLum >= 0 && firstLum + lfs <= a.srcH;
            const int oa = j4 * 16, ob = oa + MT_W * 2;
            for (int j = 0; j < lfs; j++) {
                const int cc = __ldg(k2 + (j >> 1));
                const unsigned k = (unsigned)((j & 1) ? cc >> 16 : (int)(short)cc);
                const uint8_t *lp = linL + (inside ? firstLum + j - la : min(max(firstLum + j, 0), a.srcH - 1) - la) * MT_TPB;
                const int4 p0 = *reinterpret_cast<const int4 *>(lp + oa), p1 = *reinterpret_cast<const int4 *>(lp + ob);
                const int w[8] = { p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w };
#pragma unroll
                for (int i = 0; i < 8; i++) sY[i] += (unsigned)w[i] * k;
            }
        }
        {
            const int32_t *k2 = t.vChr2 + (long long)dy * ((cfs + 1) >> 1);
            const bool inside = firstChr >= 0 && firstChr + cfs <= a.chrSrcH;
            const int oa = j4 * 8, ob = oa + MT_W;
            for (int j = 0; j < cfs; j++) {
                const int cc = __ldg(k2 + (j >> 1));
                const unsigned k = (unsigned)((j & 1) ? cc >> 16 : (int)(short)cc);
                const int lo = (inside ? firstChr + j - ca : min(max(firstChr + j, 0), a.chrSrcH - 1) - ca) * MT_CPB;
                const int2 ua = *reinterpret_cast<const int2 *>(linU + lo + oa), ub = *reinterpret_cast<const int2 *>(linU + lo + ob);
                const int2 va = *reinterpret_cast<const int2 *>(linV + lo + oa), vb = *reinterpret_cast<const int2 *>(linV + lo + ob);
                sU[0] += (unsigned)ua.x * k; sU[1] += (unsigned)ua.y * k; sU[2] += (unsigned)ub.x * k; sU[3] += (unsigned)ub.y * k;
                sV[0] += (unsigned)va.x * k; sV[1] += (unsigned)va.y * k; sV[2] += (unsigned)vb.x * k; sV[3] += (unsigned)vb.y * k;
            }
        }
        // quad swap with the neighbouring lane: the even lane keeps both left-half quads, the odd lane both right-half quads
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned give = odd ? sY[i] : sY[4 + i], got = __shfl_xor_sync(0xffffffffu, give, 1);
            if (odd) sY[i] = got; else sY[4 + i] = got;                  // even: [own a, neighbour's a]; odd: [neighbour's b, own b]
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const unsigned gu = odd ? sU[i] : sU[2 + i], gv = odd ? sV[i] : sV[2 + i];
            const unsigned ru = __shfl_xor_sync(0xffffffffu, gu, 1), rv = __shfl_xor_sync(0xffffffffu, gv, 1);
            if (odd) { sU[i] = ru; sV[i] = rv; } else { sU[2 + i] = ru; sV[2 + i] = rv; }
        }
        if (!active) continue;
        unsigned m[4 * PW];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const ChromaBase cb = chroma_base(c, (int)sU[i] >> 19, (int)sV[i] >> 19);
            pair_out<KIND>(c.cy, cb, (int)sY[2 * i] >> 19, (int)sY[2 * i + 1] >> 19, m + PW * i);
        }
        uint8_t *d = a.dst + f * a.dfs + (long long)dy * a.ds + (long long)x * BPP;
        if (KIND <= SWS_OUT_BGR24) {
            uint2 *d2 = reinterpret_cast<uint2 *>(d);
#pragma unroll
            for (int k2 = 0; k2 < 3; k2++)
                d2[k2] = make_uint2(__byte_perm(m[4 * k2], m[4 * k2 + 1], 0x6420), __byte_perm(m[4 * k2 + 2], m[4 * k2 + 3], 0x6420));
        } else {
            uint4 *d4 = reinterpret_cast<uint4 *>(d);
            d4[0] = make_uint4(m[0], m[1], m[2], m[3]); d4[1] = make_uint4(m[4], m[5], m[6], m[7]);
        }
    }
}
