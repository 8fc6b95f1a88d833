/*
 * poa_dp.cuh -- adaptive-banded convex-gap DP of one read against the graph (rows in topological
 * order), two variants:
 *
 *   dp_align32        int32 lanes, one cell per lane, band of any width in chunks of 32 cells
 *                     (whole warp; the fall-back for bands / scores the packed variant refuses)
 *   dp_band16<T,WPL>  packed int16x2 (DPX: VIADDMNMX.S16x2 / VIMNMX3.S16x2) for a TEAM of T lanes,
 *                     every lane owns 2*WPL CONSECUTIVE cells (WPL 32-bit words), band <= T*2*WPL
 *                     cells.  Scores are kept RELATIVE to match * (first column of the lane), so
 *                     they fit 16 bits for reads of any length (abPOA's int32 lane width only
 *                     changes the vector length its band is rounded to).  The insertion
 *                     recurrence F is solved with (F1,F2) packed in one word: a serial pass over
 *                     the lane's cells, ONE decayed max-scan across lanes per row, one fix-up pass.
 *
 * Both compute exactly the recurrences abPOA evaluates for `-M 5 -r 0`
 * (reference utils/SpliceDefineConsensus.py:917; restated in DESIGN.md "Algorithm"):
 *   M  = max_p H[p][j-1] + s      Ein = max_p Eout[p][j]      F[j] = max(Hhat[j-1]-oe, F[j-1]-e)
 *   H  = max(M, Ein1, Ein2, F1, F2)                          Eout = max(Ein - e, H - oe)
 * with abPOA's band (rounded to its SIMD vector length pn, clamped by the predecessors) and its
 * quirk that the diagonal is not carried into the first cell of the overlap with a predecessor.
 *
 * Every row's H/Eout1/Eout2 go (a) into the shared-memory ring for the next rows and (b) to the
 * HBM traceback area: row r at rowtb[r].x (offset in 4-byte units), three arrays of rowtb[r].y
 * elements; rowtb[r].z / .w keep the row's first predecessor and base for the traceback.
 */
#pragma once
#include "poa_graph.cuh"

namespace mpoa {

struct AlnState {
    int best_i, best_j, best_score, pn, lgpn, bits;
    unsigned long long cells, intops, full, tbbytes;
};

/* per-alignment work counters kept in 32 bits inside the row loop */
struct RowCount {
    uint32_t cells, extra, rows;
    __device__ __forceinline__ void add(int width, int npre) { cells += width; extra += (uint32_t)max(0, npre - 1) * width; ++rows; }
    __device__ __forceinline__ void flush(AlnState &R, int qlen) const {
        R.cells = cells; R.intops = 17ull * cells + 3ull * extra; R.full = (unsigned long long)rows * (qlen + 1);
    }
};

/* lane width abPOA would have used (decides the SIMD vector length the band is rounded to) */
__device__ __forceinline__ void lane_width_rule(const DevParams &P, int qlen, int N, AlnState &R) {
    const int len = max(qlen, N);
    const long long max_score = max((long long)qlen * P.match, (long long)len * P.e1 + P.o1);
    const bool is16 = max_score <= 32767 - P.mismatch - P.o1 - P.e1 - P.o2 - P.e2;
    R.pn = is16 ? P.pn16 : P.pn32;
    R.lgpn = 31 - __clz(R.pn);
    R.bits = is16 ? 16 : 32;
}

/* band of a row from its predecessors' row maxima (pull form of abPOA's max_pos_left/right) */
struct Band { int beg_sn, end_sn, dp_beg, hi_cell, width; };

__device__ __forceinline__ Band make_band(int left, int right, int minb, int maxe, int rem, int qlen, int w, int lg) {
    Band b;
    const int beg = max(0, min(left, qlen - rem) - w);
    const int end = min(qlen, max(right, qlen - rem) + w);
    b.beg_sn = max(beg >> lg, minb);
    b.end_sn = min(end >> lg, maxe + 1);
    b.dp_beg = b.beg_sn << lg;
    b.hi_cell = min(((b.end_sn + 1) << lg) - 1, qlen);
    b.width = max(0, b.hi_cell - b.dp_beg + 1);
    return b;
}

/* best end cell: the sink's predecessors in edge order, first maximum wins */
__device__ __forceinline__ void pick_best(const KernelArgs &A, const Slot &S, int N, int qlen, AlnState &R) {
    const uint32_t *in_off = in_off_p(A, S), *in_row = in_row_p(A, S);
    const int s0 = (int)in_off[N - 1], s1 = (int)in_off[N];
    int best = NEG, bi = 0, bj = 0;
    for (int e = s0; e < s1; ++e) {
        const int p = (int)in_row[e];
        const int4 pi = rowinfo_p(A, S)[p];
        const int v = rowbest_p(A, S)[p];
        if (v > best) { best = v; bi = p; bj = min(qlen, ((pi.y + 1) << R.lgpn) - 1); }
    }
    R.best_i = bi; R.best_j = bj; R.best_score = best;
}

/* ------------------------------------------------------------------------------------------ */
/* int32 lanes                                                                                 */
/* ------------------------------------------------------------------------------------------ */

__device__ __forceinline__ int dp_align32(const KernelArgs &A, const Slot &S, int N, const uint8_t *__restrict__ q,
                                          int qlen, int *ring, int4 *ring_info, int lane, AlnState &R) {
    const DevParams &P = A.P;
    const int wcap = A.wcap;
    const uint32_t *in_off = in_off_p(A, S), *in_row = in_row_p(A, S);
    int32_t *tb = reinterpret_cast<int32_t *>(tb_p(A, S));
    const uint64_t tbcap = A.L.tbcap / 4;
    lane_width_rule(P, qlen, N, R);
    const int lg = R.lgpn;
    const int w = P.wb < 0 ? qlen : P.wb + (int)__fmul_rn(P.wf, (float)qlen);
    uint32_t tb_used = 0;
    RowCount RC = {0, 0, 0};

    /* row 0: the source */
    int4 prev_info;
    {
        const int rem0 = remain_p(A, S)[0];
        const int e = min(qlen, max(0, qlen - rem0) + w);
        const int end_sn = e >> lg;
        const int hi = min(((end_sn + 1) << lg) - 1, qlen);
        const int width = hi + 1;
        if (width > wcap) return ST_RETRY_WIDE;
        const uint32_t stride = (uint32_t)(width + 3) & ~3u;
        if ((uint64_t)3 * stride > tbcap) return ST_RETRY;
        int *H = ring, *E1 = ring + wcap, *E2 = ring + 2 * wcap;
        for (int c = lane; c < width; c += 32) {
            const int h = c == 0 ? 0 : max(-(P.o1 + P.e1 * c), -(P.o2 + P.e2 * c));
            const int e1 = c == 0 ? -P.oe1 : NEG, e2 = c == 0 ? -P.oe2 : NEG;
            H[c] = h; E1[c] = e1; E2[c] = e2;
            tb[c] = h; tb[stride + c] = e1; tb[2 * stride + c] = e2;
        }
        prev_info = make_int4(0, end_sn, 0, 0);
        if (lane == 0) {
            ring_info[0] = prev_info;
            rowinfo_p(A, S)[0] = prev_info;
            rowtb_p(A, S)[0] = make_uint4(0, stride, 0, 0);
            if (meta_p(A, S)[0] & META_TOSINK) rowbest_p(A, S)[0] = hi == 0 ? 0 : max(-(P.o1 + P.e1 * hi), -(P.o2 + P.e2 * hi));   // (seeded windows only)
        }
        tb_used = 3 * stride;
        __syncwarp();
    }

    for (int w0 = 1; w0 < N - 1; w0 += 32) {
        int m_meta = 0, m_in0 = 0, m_in1 = 0, m_rem = 0, m_p0 = 0;
        {
            const int r = w0 + lane;
            if (r < N - 1) {
                m_meta = (int)meta_p(A, S)[r];
                m_in0 = (int)in_off[r];
                m_in1 = (int)in_off[r + 1];
                m_rem = remain_p(A, S)[r];
                m_p0 = m_in1 > m_in0 ? (int)in_row[m_in0] : 0;
            }
        }
        const int nrows = min(32, N - 1 - w0);
        for (int l = 0; l < nrows; ++l) {
            const int i = w0 + l;
            const int meta = __shfl_sync(FULL, m_meta, l);
            const int in0 = __shfl_sync(FULL, m_in0, l);
            const int npre = __shfl_sync(FULL, m_in1, l) - in0;
            const int rem = __shfl_sync(FULL, m_rem, l);
            const int p0 = __shfl_sync(FULL, m_p0, l);
            const int nbase = meta & META_BASE;

            /* most rows have ONE predecessor and it is the previous row: straight-line path */
            const bool simple = npre == 1 && p0 == i - 1;
            int left = N, right = 0, minb = INT_MAX, maxe = -1;
            if (simple) {
                left = min(N, prev_info.z + 1); right = max(0, prev_info.w + 1);
                minb = prev_info.x; maxe = prev_info.y;
            } else {
#pragma unroll 1
                for (int k = 0; k < npre; ++k) {
                    const int p = k == 0 ? p0 : (int)in_row[in0 + k];
                    const int4 pi = (p == i - 1) ? prev_info : ((i - p < RING) ? ring_info[p & (RING - 1)] : rowinfo_p(A, S)[p]);
                    left = min(left, pi.z + 1);
                    right = max(right, pi.w + 1);
                    minb = min(minb, pi.x);
                    maxe = max(maxe, pi.y);
                }
            }
            const Band B = make_band(left, right, minb, maxe, rem, qlen, w, lg);
            const int width = B.width, dp_beg = B.dp_beg, beg_sn = B.beg_sn, end_sn = B.end_sn;
            if (width > wcap) return ST_RETRY_WIDE;
            const uint32_t stride = (uint32_t)(width + 3) & ~3u;
            const uint32_t tbo = tb_used;
            if ((uint64_t)tb_used + 3ull * stride > tbcap) return ST_RETRY;
            tb_used += 3 * stride;
            RC.add(width, npre);

            int *Hr = ring + (i % RING) * 3 * wcap, *E1r = Hr + wcap, *E2r = Hr + 2 * wcap;
            int carry_s1 = NEG - P.oe1, carry_s2 = NEG - P.oe2;
            int rmax = NEG, lpos = -1, rpos = -1, last_h = NEG;
            const int nch = (width + 31) >> 5;
            for (int c = 0; c < nch; ++c) {
                const int col = c * 32 + lane;
                const int j = dp_beg + col;
                const bool cv = col < width;
                int mx = NEG, ei1 = NEG, ei2 = NEG;
#pragma unroll 1
                for (int k = 0; k < npre; ++k) {
                    const int p = k == 0 ? p0 : (int)in_row[in0 + k];
                    const bool near = i - p < RING;
                    const int4 pi = (p == i - 1) ? prev_info : (near ? ring_info[p % RING] : rowinfo_p(A, S)[p]);
                    const int pbeg = pi.x << lg;
                    const int lo = max(beg_sn, pi.x) << lg;
                    const int hi = min(((min(end_sn, pi.y) + 1) << lg) - 1, qlen);
                    const int *Hp;
                    int pstride;
                    if (near) { Hp = ring + (p % RING) * 3 * wcap; pstride = wcap; }
                    else { const uint4 rt = rowtb_p(A, S)[p]; Hp = tb + rt.x; pstride = (int)rt.y; }
                    const bool e_ok = cv && j >= lo && j <= hi;
                    if (e_ok) {
                        if (j > lo) mx = max(mx, Hp[j - 1 - pbeg]);
                        ei1 = max(ei1, Hp[pstride + j - pbeg]);
                        ei2 = max(ei2, Hp[2 * pstride + j - pbeg]);
                    }
                }
                int s = 0;
                if (cv && j > 0) {
                    const int qb = q[j - 1];
                    s = (nbase >= 4 || qb >= 4) ? 0 : (nbase == qb ? P.match : -P.mismatch);
                }
                const int hh = cv ? max(mx + s, max(ei1, ei2)) : NEG;
                int s1 = warp_scan_decay(hh - P.oe1, P.e1, lane);
                int s2 = warp_scan_decay(hh - P.oe2, P.e2, lane);
                s1 = max(s1, carry_s1 - P.e1 * (lane + 1));
                s2 = max(s2, carry_s2 - P.e2 * (lane + 1));
                int f1 = __shfl_up_sync(FULL, s1, 1), f2 = __shfl_up_sync(FULL, s2, 1);
                if (lane == 0) { f1 = carry_s1; f2 = carry_s2; }
                carry_s1 = __shfl_sync(FULL, s1, 31);
                carry_s2 = __shfl_sync(FULL, s2, 31);
                const int h = max(hh, max(f1, f2));
                const int e1o = max(ei1 - P.e1, h - P.oe1), e2o = max(ei2 - P.e2, h - P.oe2);
                if (cv) {
                    Hr[col] = h; E1r[col] = e1o; E2r[col] = e2o;
                    tb[tbo + col] = h; tb[tbo + stride + col] = e1o; tb[tbo + 2 * stride + col] = e2o;
                }
                const int hv = cv ? h : INT_MIN;
                const int cm = __reduce_max_sync(FULL, hv);
                if (cm >= rmax) {
                    const unsigned b = __ballot_sync(FULL, hv == cm);
                    if (cm > rmax) { rmax = cm; lpos = dp_beg + c * 32 + __ffs(b) - 1; }
                    rpos = dp_beg + c * 32 + 31 - __clz(b);
                }
                if (c == nch - 1) last_h = __shfl_sync(FULL, h, (width - 1) & 31);
            }
            prev_info = make_int4(beg_sn, end_sn, lpos, rpos);
            if (lane == 0) {
                ring_info[i % RING] = prev_info;
                rowinfo_p(A, S)[i] = prev_info;
                rowtb_p(A, S)[i] = make_uint4(tbo, stride, (uint32_t)p0, (uint32_t)nbase);
                if (meta & META_TOSINK) rowbest_p(A, S)[i] = width > 0 ? last_h : NEG;
            }
            __syncwarp();
        }
    }
    R.tbbytes = (unsigned long long)tb_used * 4;
    RC.flush(R, qlen);
    pick_best(A, S, N, qlen, R);
    return ST_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* packed int16x2 lanes                                                                        */
/* ------------------------------------------------------------------------------------------ */

__device__ __forceinline__ uint32_t pack2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return (int)w >> 16; }

/* shared-memory words of one team: RINGV rows x 3 arrays x T*WPL words (lane-stationary slots),
 * the query profile of the lanes' current columns (5 bases x T*WPL words) and the window records */
template <int WPL>
__host__ __device__ constexpr int ring16_rows() { return WPL >= 4 ? 4 : 8; }   // rows kept in shared memory
template <int T, int WPL>
__host__ __device__ constexpr int ring16_team_words() { return ring16_rows<WPL>() * 3 * T * WPL + 5 * T * WPL + T * 5; }

/* WPL words from / to global memory: 128-bit accesses when the lane's run is 16-byte aligned
 * (WPL = 4, 8), 64-bit otherwise (WPL = 2, 6) */
template <int WPL>
__device__ __forceinline__ void ld_words(const uint32_t *p, uint32_t (&v)[WPL]) {
    if constexpr (WPL % 4 == 0) {
#pragma unroll
        for (int m = 0; m < WPL; m += 4) {
            const uint4 a = *reinterpret_cast<const uint4 *>(p + m);
            v[m] = a.x; v[m + 1] = a.y; v[m + 2] = a.z; v[m + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int m = 0; m < WPL; m += 2) { const uint2 a = *reinterpret_cast<const uint2 *>(p + m); v[m] = a.x; v[m + 1] = a.y; }
    }
}
template <int WPL>
__device__ __forceinline__ void st_words(uint32_t *p, const uint32_t (&v)[WPL]) {
    if constexpr (WPL % 4 == 0) {
#pragma unroll
        for (int m = 0; m < WPL; m += 4) *reinterpret_cast<uint4 *>(p + m) = make_uint4(v[m], v[m + 1], v[m + 2], v[m + 3]);
    } else {
#pragma unroll
        for (int m = 0; m < WPL; m += 2) *reinterpret_cast<uint2 *>(p + m) = make_uint2(v[m], v[m + 1]);
    }
}

/* shared memory by 32-bit address: the row loop never converts generic pointers.  All of them are
 * volatile, so they keep their program order among themselves. */
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int WPL>
__device__ __forceinline__ void lds_words(uint32_t a, uint32_t (&v)[WPL]) {
    if constexpr (WPL % 4 == 0) {
#pragma unroll
        for (int m = 0; m < WPL; m += 4)
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[m]), "=r"(v[m + 1]), "=r"(v[m + 2]), "=r"(v[m + 3]) : "r"(a + 4 * m));
    } else {
#pragma unroll
        for (int m = 0; m < WPL; m += 2)
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v[m]), "=r"(v[m + 1]) : "r"(a + 4 * m));
    }
}
template <int WPL>
__device__ __forceinline__ void sts_words(uint32_t a, const uint32_t (&v)[WPL]) {
    if constexpr (WPL % 4 == 0) {
#pragma unroll
        for (int m = 0; m < WPL; m += 4)
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a + 4 * m), "r"(v[m]), "r"(v[m + 1]), "r"(v[m + 2]), "r"(v[m + 3]));
    } else {
#pragma unroll
        for (int m = 0; m < WPL; m += 2)
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(a + 4 * m), "r"(v[m]), "r"(v[m + 1]));
    }
}
__device__ __forceinline__ void sts_word(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v)); }
__device__ __forceinline__ uint32_t lds_word(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
/* keeps a loop-invariant value in a register (instead of re-deriving it in every row) */
__device__ __forceinline__ uint32_t pin_reg(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
/* key of a row-maximum reduction: (16-bit half of h) << 16 | t; SEL picks the half (0x1054 low, 0x3254 high) */
template <int TT>
__device__ __forceinline__ int key_of(uint32_t h, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(h), "n"(TT), "r"(sel));
    return (int)d;
}
/* the same key for the LOW half of h as one multiply-add: it runs on the FMA pipe, the packed DP
 * keeps the ALU pipe busy */
template <int TT>
__device__ __forceinline__ int key_lo(uint32_t h) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(d) : "r"(h), "n"(TT));
    return (int)d;
}
/* per-lane maxima of the keys (H << 16 | cell) and (H << 16 | 0xffff-cell) over the lane's words */
template <int M, int WPL>
struct LaneKeys {
    static __device__ __forceinline__ void run(const uint32_t (&H)[WPL], uint32_t sel, int &kr, int &kl) {
        if constexpr (M == 0) {
            kr = max(key_lo<0>(H[0]), key_of<1>(H[0], sel));
            kl = max(key_lo<0xffff>(H[0]), key_of<0xffff - 1>(H[0], sel));
        } else {
            kr = __vimax3_s32(kr, key_lo<2 * M>(H[M]), key_of<2 * M + 1>(H[M], sel));
            kl = __vimax3_s32(kl, key_lo<0xffff - 2 * M>(H[M]), key_of<0xffff - 1 - 2 * M>(H[M], sel));
        }
        if constexpr (M + 1 < WPL) LaneKeys<M + 1, WPL>::run(H, sel, kr, kl);
    }
};

/*
 * Lane-stationary packed DP of a team of T lanes.  Column c of the query always lives in team lane
 * (c / CPL) mod T, word (c mod CPL) / 2 -- whatever the band start is.  A row's band [dp_beg, hi]
 * (dp_beg a multiple of abPOA's SIMD vector length, any multiple of 4) is a ROTATED run of lanes
 * starting at the lane that holds dp_beg; it must fit T*CPL cells counted from that lane's first
 * column ("origin").  Then
 *   - a row whose only predecessor is the previous row (4 of 5 rows) takes H/E1/E2 of that row
 *     straight from the registers they were computed in: no shared-memory gather, no shifting
 *     when the band moves (a lane the band start has passed re-binds to the column T*CPL further
 *     right and resets its registers to -inf);
 *   - the shared-memory ring and the query profile are indexed by lane: every lane reads back
 *     exactly the slots it wrote (no bank conflicts, no sync between rows);
 *   - cells outside the band are kept at -inf (a per-cell mask that changes only with the band),
 *     so neighbours and later rows see what abPOA's band would have shown them; when the band
 *     start moves INSIDE a lane the cells it leaves behind are masked in the carried registers.
 * Scores are RELATIVE: a lane holds H - match*c0 (c0 = its first column), so |values| stay small
 * for reads of any length; the frame changes by match*CPL from one lane to the next (K.mhop), which
 * only the diagonal hand-over, the lane scan of the insertion recurrence and the row-maximum keys
 * see.  -inf is NEG16 in every frame; a row whose best relative score drops below LOW16 (junk
 * reads, rows without any finite cell) makes the alignment leave the packed path (ST_RETRY_32):
 * above that line no clamped value can reach a row maximum or the traceback path.
 * The insertion recurrence runs over the rotated lane order: rotate, plain max-scan of the lane
 * totals with the decay added back (T + position * decay), rotate back.
 * Per-row bookkeeping (band vectors, row maxima, traceback offsets) is parked in shared memory and
 * written to HBM T rows at a time.
 */
template <int T, int WPL>
__device__ __forceinline__ int dp_band16(const KernelArgs &A, const Slot &S, const Team<T> &tm, int N_in, const uint8_t *__restrict__ q,
                                         int qlen_in, uint32_t *ring, AlnState &R, bool on) {
    constexpr int CPL = 2 * WPL;              // cells per lane
    constexpr int WCAP = T * CPL;             // cells per row
    constexpr int RW = T * WPL;               // words of one ring / profile row
    constexpr int RINGV = ring16_rows<WPL>();
    constexpr int SEG = 4;                    // cells of one insertion fix-up chain
    static_assert(WPL == 2 || WPL == 4 || WPL == 6 || WPL == 8, "words per lane");
    const int lane = tm.tl;
    const DevParams &P = A.P;
    const Packed16 &K = A.K;
    const int N = tm.uniform(N_in), qlen = tm.uniform(qlen_in);
    const uint32_t *in_off = in_off_p(A, S), *in_row = in_row_p(A, S);
    uint32_t *tb = reinterpret_cast<uint32_t *>(tb_p(A, S));   // words = pairs of int16 cells
    const uint32_t tbcap = A.tbcap_words;   // (a kernel parameter: used straight from the constant bank)
    lane_width_rule(P, qlen, N, R);
    /* sticky status of this team: a team that fails (or is switched off) keeps walking through the
     * loops with its stores disabled -- the other team of the warp needs the lockstep */
    int err = on ? ST_OK : ST_PENDING;
    if (R.pn & 3) err = ST_RETRY_32;          // band edges must fall on word pairs
    if (N > 131071) err = ST_RETRY_32;        // remain is carried in 18 signed bits of the row metadata word
    const int hop1 = (P.e1 + P.match) * CPL, hop2 = (P.e2 + P.match) * CPL;   // what a lane hop costs the insertion scores
    if ((T - 1) * max(hop1, hop2) + P.match * CPL > 4700) err = ST_RETRY_32;   // head-room of the decay-free scan above/below NEG16
    const int lg = R.lgpn;
    const bool sublane = (R.pn % CPL) != 0;   // band edges may fall inside a lane
    const int w = P.wb < 0 ? qlen : P.wb + (int)__fmul_rn(P.wf, (float)qlen);
    uint32_t tb_used = 0;                     // words
    uint32_t cells = 0, extra = 0;

    const uint32_t NEG2 = K.neg2;
    const uint32_t ring_a = pin_reg(smem_addr(ring + lane * WPL));             // + (row & (RINGV-1)) * 3*RW*4 + array * RW*4
    const uint32_t prof_a = ring_a + RINGV * 3 * RW * 4;                        // + base * RW*4
    const uint32_t wrec_a = pin_reg(smem_addr(ring + RINGV * 3 * RW + 5 * RW)); // window records: T x {beg_sn, end_sn, lpos, rpos}, then T x tb offset
    int4 *rowinfo_g = rowinfo_p(A, S);
    uint4 *rowtb_g = rowtb_p(A, S);
    const uint32_t sel_hi = pin_reg(0x3254u);
    const uint32_t lane_up = pin_reg(pack2(hop1 * lane, hop2 * lane));   // decays added back before the scan (by position)

    /* query profile of the whole read, once: qprof[b][w] = (s(b, q[2w-1]), s(b, q[2w])), the scores
     * of cells 2w and 2w+1 against node base b.  A lane that re-binds copies its words from here. */
    uint32_t *qprof = qprof_p(A, S);
    const int qps = (int)qprof_stride(A.L.qcap);
    if (err == ST_OK) {
#pragma unroll 1
        for (int wd = lane; wd <= (qlen >> 1); wd += T) {
            const int qa = q[max(2 * wd - 1, 0)], qc = q[min(2 * wd, qlen - 1)];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int sa = qa >= 4 ? 0 : (qa == b ? P.match : -P.mismatch);
                const int sb = qc >= 4 ? 0 : (qc == b ? P.match : -P.mismatch);
                qprof[b * qps + wd] = pack2(sa, sb);
            }
        }
    }
    tm.sync();

    /* lane binding (changes only when the band does) */
    int cur_beg = -1, cur_hi = -1, cur_org = 0, cur_width = 0;   // band, first column of the lane that holds its start, cells from there
    uint32_t cur_stw = 0;
    int rl = lane, c0 = -1;                   // rotated lane index, first column of this lane
    uint32_t c_dec = 0;                       // what the scanned total of position rl-1 loses on its way to this lane's first cell
    uint32_t MK[WPL];                         // 0xffff per in-band cell
    uint32_t Hp[WPL], E1p[WPL], E2p[WPL];     // the previous row at this lane's columns
#pragma unroll
    for (int m = 0; m < WPL; ++m) { MK[m] = 0; Hp[m] = E1p[m] = E2p[m] = NEG2; sts_word(prof_a + (4 * RW + m) * 4, 0); }

    auto stw_of = [&](int width) { return (uint32_t)(((width + 1) >> 1) + WPL - 1) / WPL * WPL; };
    auto org_of = [&](int beg) { return beg / CPL * CPL; };
    auto rebind = [&](int beg, int hi) {      // no cross-lane operation in here
        const int org = org_of(beg);
        cur_width = max(0, hi - org + 1);
        if (hi < beg) cur_width = 0;
        if (cur_width > WCAP) { if (err == ST_OK) err = ST_RETRY_WIDE; cur_width = WCAP; hi = org + WCAP - 1; }
        cur_beg = beg; cur_hi = hi; cur_org = org;
        cur_stw = stw_of(cur_width);
        rl = (lane - org / CPL) & (T - 1);
        c_dec = pack2(-hop1 * (rl - 1) - P.match * CPL, -hop2 * (rl - 1) - P.match * CPL);
        const int nc0 = org + rl * CPL;
        if (nc0 != c0) {
            c0 = nc0;
#pragma unroll
            for (int m = 0; m < WPL; ++m) Hp[m] = E1p[m] = E2p[m] = NEG2;
            if (c0 <= qlen) {                 // the lane's new columns: copy their profile words
                const uint32_t *src = qprof + (c0 >> 1);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    uint32_t v[WPL];
                    ld_words<WPL>(src + b * qps, v);
                    sts_words<WPL>(prof_a + b * (RW * 4), v);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < WPL; ++m) {
            const int ca = c0 + 2 * m, cb = ca + 1;
            MK[m] = ((ca >= beg && ca <= hi) ? 0xffffu : 0u) | ((cb >= beg && cb <= hi) ? 0xffff0000u : 0u);
        }
        /* the band start moved inside this lane: the cells it left behind offer nothing any more */
        if (rl == 0 && beg != org) {
#pragma unroll
            for (int m = 0; m < WPL; ++m) {
                Hp[m] = (Hp[m] & MK[m]) | (NEG2 & ~MK[m]);
                E1p[m] = (E1p[m] & MK[m]) | (NEG2 & ~MK[m]);
                E2p[m] = (E2p[m] & MK[m]) | (NEG2 & ~MK[m]);
            }
        }
    };
    /* word of this lane that holds column qc (or -1): the first cell past a predecessor's band */
    auto word_of = [&](int qc) { const int d = qc - c0; return (d >= 0 && d < CPL) ? (d >> 1) : -1; };

    /* the previous row's band vectors and row-maximum columns (uniform in the team) */
    int p_bs = 0, p_es = 0, p_l = 0, p_r = 0;

    /* row 0: the source */
    if (err == ST_OK) {
        const int rem0 = remain_p(A, S)[0];
        const int e = min(qlen, max(0, qlen - rem0) + w);
        const int end_sn = e >> lg;
        const int hi = min(((end_sn + 1) << lg) - 1, qlen);
        if (hi + 1 > WCAP) err = ST_RETRY_WIDE;
        else {
            rebind(0, hi);
            if (3 * cur_stw > tbcap) err = ST_RETRY;
        }
        if (err == ST_OK) {
#pragma unroll
            for (int m = 0; m < WPL; ++m) {
                int hv[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int cc = c0 + 2 * m + t;
                    hv[t] = cc == 0 ? 0 : max(max(-(P.o1 + P.e1 * cc), -(P.o2 + P.e2 * cc)) - P.match * c0, NEG16);
                }
                Hp[m] = (pack2(hv[0], hv[1]) & MK[m]) | (NEG2 & ~MK[m]);
                E1p[m] = (c0 + 2 * m == 0) ? pack2(-P.oe1, NEG16) : NEG2;
                E2p[m] = (c0 + 2 * m == 0) ? pack2(-P.oe2, NEG16) : NEG2;
            }
            sts_words<WPL>(ring_a, Hp); sts_words<WPL>(ring_a + RW * 4, E1p); sts_words<WPL>(ring_a + 2 * RW * 4, E2p);
            if ((uint32_t)(rl * WPL) < cur_stw) {
                uint32_t *g = tb + rl * WPL;
                st_words<WPL>(g, Hp); st_words<WPL>(g + cur_stw, E1p); st_words<WPL>(g + 2 * cur_stw, E2p);
            }
            p_bs = 0; p_es = end_sn; p_l = 0; p_r = 0;
            if (lane == 0) {
                rowinfo_g[0] = make_int4(0, end_sn, 0, 0);
                rowtb_g[0] = make_uint4(0, 2 * cur_stw, 0, 0);
                /* the source can be a direct predecessor of the sink only in a seeded window with nothing
                 * between its anchors: the read's stretch is then one insertion */
                if (meta_p(A, S)[0] & META_TOSINK) rowbest_p(A, S)[0] = hi == 0 ? 0 : max(-(P.o1 + P.e1 * hi), -(P.o2 + P.e2 * hi));
            }
            tb_used = 3 * cur_stw;
        }
    }

    const int nwin = err == ST_OK ? (N - 2 + T - 1) / T : 0;
    const int wn = tm.wmax(nwin);
    for (int wi = 0; wi < wn; ++wi) {
        const int w0 = 1 + wi * T;
        const bool live = err == ST_OK && w0 < N - 1;
        /* row metadata of T rows at once: a = base | flags | simple<<5 | npre<<6 | remain<<14 */
        uint32_t m_a = 32u;                   // rows of a team that is switched off count as "simple"
        int m_in0 = 0, m_p0 = 0;
        {
            const int r = w0 + lane;
            if (live && r < N - 1) {
                m_in0 = (int)in_off[r];
                const int npre = (int)in_off[r + 1] - m_in0;
                m_p0 = npre > 0 ? (int)in_row[m_in0] : 0;
                const uint32_t simple = (npre == 1 && m_p0 == r - 1) ? 32u : 0u;
                m_a = (meta_p(A, S)[r] & 31u) | simple | ((uint32_t)min(npre, 255) << 6) | ((uint32_t)remain_p(A, S)[r] << 14);
            }
        }
        /* the results of the window's rows (band vectors, row-maximum columns, traceback offset) are
         * parked in shared memory by lane 0 and written to HBM by all lanes at the end of the window */
        const int nrows = live ? min(T, N - 1 - w0) : 0;
        const int nr = tm.wmax(nrows);
        for (int l = 0; l < nr; ++l) {
            const int i = w0 + l;
            const bool rowon = l < nrows && err == ST_OK;   // this team computes a row in this step
            const uint32_t ma = tm.pick(m_a, l);   // uniform in the team
            const int nbase = ma & META_BASE;
            const int rem = (int)ma >> 14;     // signed: in a seeded window a row's heaviest path may by-pass the window's end
            const bool simple = (ma & 32u) != 0 || !rowon;
            const bool all_simple = tm.wall(simple);
            int npre = 1, in0 = 0, p0 = i - 1;

            Band B;
            if (all_simple) {
                B = make_band(min(N, p_l + 1), max(0, p_r + 1), p_bs, p_es, rem, qlen, w, lg);
            } else {
                int left, right, minb, maxe;
                tm.sync();                    // rows of earlier windows / traceback rows written by other lanes
                npre = rowon ? (int)((ma >> 6) & 255) : 0;
                in0 = tm.shfl(m_in0, l);
                p0 = tm.shfl(m_p0, l);
                if (npre == 255) npre = (int)in_off[i + 1] - in0;
                left = N; right = 0; minb = INT_MAX; maxe = -1;
#pragma unroll 1
                for (int k = 0; k < npre; ++k) {          // per team; no cross-lane operation
                    const int p = k == 0 ? p0 : (int)in_row[in0 + k];
                    int bs, es, pl, pr;
                    if (p == i - 1) { bs = p_bs; es = p_es; pl = p_l; pr = p_r; }
                    else if (p >= w0) {
                        uint32_t v[4];
                        lds_words<4>(wrec_a + (p - w0) * 16, v);
                        bs = (int)v[0]; es = (int)v[1]; pl = (int)v[2]; pr = (int)v[3];
                    } else { const int4 pi = rowinfo_g[p]; bs = pi.x; es = pi.y; pl = pi.z; pr = pi.w; }
                    left = min(left, pl + 1);
                    right = max(right, pr + 1);
                    minb = min(minb, bs);
                    maxe = max(maxe, es);
                }
                if (npre == 0) { minb = p_bs; maxe = p_es; left = min(N, p_l + 1); right = max(0, p_r + 1); }
                B = make_band(left, right, minb, maxe, rem, qlen, w, lg);
            }
            if (rowon && (B.dp_beg != cur_beg || B.hi_cell != cur_hi)) rebind(B.dp_beg, B.hi_cell);
            /* match/mismatch scores of this lane's cells against the node's base (after a re-bind:
             * the lane's profile slots may just have been rewritten) */
            uint32_t S2[WPL];
            lds_words<WPL>(prof_a + nbase * (RW * 4), S2);
            const uint32_t stw = cur_stw;
            const uint32_t tbo = tb_used;
            const bool want = rowon && err == ST_OK;   // (a re-bind may just have failed)
            const bool st_ok = want && tbo + 3 * stw <= tbcap;
            tb_used = st_ok ? tbo + 3 * stw : tbo;
            cells += st_ok ? (uint32_t)B.width : 0u;
            if (want && !st_ok) err = ST_RETRY;

            /* diagonal and deletion inputs */
            uint32_t M2[WPL], EA[WPL], EB[WPL];
            if (all_simple) {
                /* the left neighbour's last cell, moved into this lane's score frame */
                const uint32_t hl = __viaddmax_s16x2(tm.shfl(Hp[WPL - 1], lane - 1), K.mhop, NEG2);
#pragma unroll
                for (int m = 0; m < WPL; ++m) {
                    M2[m] = __byte_perm(m == 0 ? hl : Hp[m - 1], Hp[m], 0x5432);
                    EA[m] = E1p[m]; EB[m] = E2p[m];
                }
                /* a predecessor offers nothing right of ITS rounded band end (abPOA only walks the
                 * overlapping vectors): the first cell past it has no diagonal either */
                if (B.end_sn > p_es) {
                    const int wq = word_of((p_es + 1) << lg);
#pragma unroll
                    for (int m = 0; m < WPL; m += 2)
                        if (wq == m) M2[m] = (M2[m] & 0xffff0000u) | (NEG2 & 0xffffu);
                }
            } else {
                if (st_ok) extra += (uint32_t)max(0, npre - 1) * B.width;
#pragma unroll
                for (int m = 0; m < WPL; ++m) { M2[m] = NEG2; EA[m] = NEG2; EB[m] = NEG2; }
                const int kmax = tm.wmax(npre);
#pragma unroll 1
                for (int k = 0; k < kmax; ++k) {
                    bool ok = k < npre;       // this team has a k-th predecessor
                    int es = 0;
                    uint32_t hw[WPL], e1w[WPL], e2w[WPL];
                    if (ok) {
                        const int p = k == 0 ? p0 : (int)in_row[in0 + k];
                        int bs;
                        uint32_t ptbo = 0;    // words: where the stored row starts (its origin column)
                        int pst = 0;          // words per array of the stored row
                        if (p >= w0) {
                            uint32_t v[4];
                            lds_words<4>(wrec_a + (p - w0) * 16, v);
                            ptbo = lds_word(wrec_a + T * 16 + (p - w0) * 4);
                            bs = (int)v[0]; es = (int)v[1];
                        } else if (p == i - 1) { bs = p_bs; es = p_es; }
                        else { const int4 pi = rowinfo_g[p]; bs = pi.x; es = pi.y; }
                        const int pbeg = bs << lg, porg = org_of(pbeg);
                        if (p >= w0) pst = (int)stw_of(min(WCAP, max(0, min(((es + 1) << lg) - 1, qlen) - porg + 1)));
                        const bool near = i - p < RINGV;
                        const int d = c0 - porg;              // my first column relative to the stored row's first column
                        ok = (unsigned)d < (unsigned)WCAP;    // same column binding as when row p was computed
                        if (near) {
                            const uint32_t a = ring_a + (p & (RINGV - 1)) * (3 * RW * 4);
                            lds_words<WPL>(a, hw); lds_words<WPL>(a + RW * 4, e1w); lds_words<WPL>(a + 2 * RW * 4, e2w);
                        } else {
                            if (p < w0) { const uint4 rt = rowtb_g[p]; ptbo = rt.x - (uint32_t)((pbeg - porg) >> 1); pst = (int)(rt.y >> 1); }
                            ok = ok && (d >> 1) + WPL <= pst;
                            const uint32_t *src = tb + ptbo + (ok ? (d >> 1) : 0);
                            if (ok) { ld_words<WPL>(src, hw); ld_words<WPL>(src + pst, e1w); ld_words<WPL>(src + 2 * pst, e2w); }
                        }
                    }
                    if (!ok) {
#pragma unroll
                        for (int m = 0; m < WPL; ++m) { hw[m] = NEG2; e1w[m] = NEG2; e2w[m] = NEG2; }
                    } else if (sublane) {
                        /* band edges inside lanes: only the cells of THIS row's band take part (the stored
                         * row is masked by its own band; whole lanes outside are never bound to it) */
#pragma unroll
                        for (int m = 0; m < WPL; ++m) {
                            hw[m] = (hw[m] & MK[m]) | (NEG2 & ~MK[m]);
                            e1w[m] = (e1w[m] & MK[m]) | (NEG2 & ~MK[m]);
                            e2w[m] = (e2w[m] & MK[m]) | (NEG2 & ~MK[m]);
                        }
                    }
                    const uint32_t hl = __viaddmax_s16x2(tm.shfl(hw[WPL - 1], lane - 1), K.mhop, NEG2);
                    const int wq = word_of((es + 1) << lg);   // nothing from p right of its rounded band end
#pragma unroll
                    for (int m = 0; m < WPL; ++m) {
                        uint32_t dg = __byte_perm(m == 0 ? hl : hw[m - 1], hw[m], 0x5432);
                        if ((m & 1) == 0 && wq == m) dg = (dg & 0xffff0000u) | (NEG2 & 0xffffu);
                        M2[m] = __vmaxs2(M2[m], dg);
                        EA[m] = __vmaxs2(EA[m], e1w[m]); EB[m] = __vmaxs2(EB[m], e2w[m]);
                    }
                }
            }
            /* first cell of the row's band: no diagonal at all (when the band starts inside the lane
             * its left neighbour was masked above / by the re-bind) */
            if (rl == 0 && cur_beg == cur_org) M2[0] = (M2[0] & 0xffff0000u) | (NEG2 & 0xffffu);

            uint32_t HH[WPL];
#pragma unroll
            for (int m = 0; m < WPL; ++m) HH[m] = __vimax3_s16x2(__vadd2(M2[m], S2[m]), EA[m], EB[m]);

            /* insertion scores: per cell a word (F1, F2); X[t] = (hh[t]-oe1, hh[t]-oe2) is what cell t
             * offers to cell t+1 */
            uint32_t X[CPL], FL[CPL];
#pragma unroll
            for (int m = 0; m < WPL; ++m) {
                const uint32_t a1 = __viaddmax_s16x2(HH[m], K.noe1, NEG2), a2 = __viaddmax_s16x2(HH[m], K.noe2, NEG2);
                X[2 * m] = __byte_perm(a1, a2, 0x5410);
                X[2 * m + 1] = __byte_perm(a1, a2, 0x7632);
            }
            /* what the lane's own cells hand to the first cell of every 4-cell segment (FL) and to the
             * next lane (TT), without a carry-in: per segment a depth-2 tree instead of a serial chain,
             *   S(X0..X3) = max(max(X3, X2-e), max(X1, X0-e) - 2e) */
            FL[0] = NEG2;
            uint32_t TT = NEG2;
#pragma unroll
            for (int g = 0; g < CPL; g += SEG) {
                const uint32_t u = __viaddmax_s16x2(X[g + 2], K.nee, X[g + 3]);
                const uint32_t v = __viaddmax_s16x2(X[g], K.nee, X[g + 1]);
                const uint32_t sg = __viaddmax_s16x2(v, K.tdec[2], u);
                TT = g == 0 ? sg : __viaddmax_s16x2(TT, K.tdec[SEG], sg);
                if (g + SEG < CPL) FL[g + SEG] = TT;
            }
            /* max-scan over the ROTATED lane order.  With position * decay added back the decayed scan
             * is a plain running maximum; C = what reaches this lane's first cell from the lanes before */
            TT = __vadd2(tm.shfl(TT, lane + lane - rl), lane_up);
#pragma unroll
            for (int dd = 1; dd < T; dd <<= 1) TT = __vmaxs2(TT, tm.shfl_up(TT, dd));
            uint32_t C = __vadd2(tm.shfl(TT, rl - 1), c_dec);
            if (rl == 0) C = NEG2;
            /* F of the lane's cells: every SEG-th cell directly, the others by the recurrence */
            uint32_t F1w[WPL], F2w[WPL];
            {
                uint32_t fa = 0, ff = 0;
#pragma unroll
                for (int t = 0; t < CPL; ++t) {
                    if (t == 0) ff = __vmaxs2(C, NEG2);
                    else if (t % SEG == 0) ff = __viaddmax_s16x2(C, K.tdec[t], FL[t]);
                    else ff = __viaddmax_s16x2(ff, K.nee, X[t - 1]);
                    if (t & 1) {
                        F1w[t >> 1] = __byte_perm(fa, ff, 0x5410);
                        F2w[t >> 1] = __byte_perm(fa, ff, 0x7632);
                    } else fa = ff;
                }
            }

            /* H; cells outside the band stay at -inf */
#pragma unroll
            for (int m = 0; m < WPL; ++m) {
                const uint32_t h = __vimax3_s16x2(HH[m], F1w[m], F2w[m]);
                Hp[m] = (h & MK[m]) | (NEG2 & ~MK[m]);
            }
            /* row maximum with its left-most and right-most column: two keyed reductions;
             * key = (H in the frame of the band's first lane) << 16 | cell (resp. reversed cell) counted
             * from that lane's first column */
            int kr, kl;
            LaneKeys<0, WPL>::run(Hp, sel_hi, kr, kl);
            /* + position of the lane and its score frame relative to the band's first lane: one multiply-add each */
            kr = tm.rmax(rl * (int)K.kc_r + kr);
            kl = tm.rmax(rl * (int)K.kc_l + kl);

            /* Eout (deletion offers to the successors), masked like H */
#pragma unroll
            for (int m = 0; m < WPL; ++m) {
                const uint32_t e1o = __viaddmax_s16x2(EA[m], K.ne1, __viaddmax_s16x2(Hp[m], K.noe1, NEG2));
                const uint32_t e2o = __viaddmax_s16x2(EB[m], K.ne2, __viaddmax_s16x2(Hp[m], K.noe2, NEG2));
                E1p[m] = (e1o & MK[m]) | (NEG2 & ~MK[m]);
                E2p[m] = (e2o & MK[m]) | (NEG2 & ~MK[m]);
            }

            /* ring (own slots) + HBM stores (band order: rotated lane rl holds words rl*WPL ...) */
            if (st_ok) {
                const uint32_t ra = ring_a + (i & (RINGV - 1)) * (3 * RW * 4);
                sts_words<WPL>(ra, Hp); sts_words<WPL>(ra + RW * 4, E1p); sts_words<WPL>(ra + 2 * RW * 4, E2p);
                if ((uint32_t)(rl * WPL) < stw) {
                    uint32_t *g = tb + tbo + rl * WPL;
                    st_words<WPL>(g, Hp); st_words<WPL>(g + stw, E1p); st_words<WPL>(g + 2 * stw, E2p);
                }
            }

            if (rowon) {
                int lpos = -1, rpos = -1;
                if (cur_width > 0) {
                    rpos = cur_org + (kr & 0xffff); lpos = cur_org + (0xffff - (kl & 0xffff));
                    /* no finite cell, or scores about to leave the 16-bit frame: the int32 kernel takes over */
                    if ((kr >> 16) < LOW16 && err == ST_OK) err = ST_RETRY_32;
                }
                p_bs = B.beg_sn; p_es = B.end_sn; p_l = lpos; p_r = rpos;
                if (lane == 0) {
                    const uint32_t v[4] = {(uint32_t)B.beg_sn, (uint32_t)B.end_sn, (uint32_t)lpos, (uint32_t)rpos};
                    sts_words<4>(wrec_a + l * 16, v);
                    sts_word(wrec_a + T * 16 + l * 4, tbo);
                }
                if (ma & META_TOSINK) {
                    /* H at the last cell of the row (absolute score): the global best is picked among these */
                    int last = NEG;
                    bool mine = B.width == 0 && lane == 0;
                    const int lc = cur_hi - cur_org;
#pragma unroll
                    for (int t = 0; t < CPL; ++t)
                        if (rl * CPL + t == lc && B.width > 0) { last = ((t & 1) ? hi16(Hp[t >> 1]) : lo16(Hp[t >> 1])) + P.match * c0; mine = true; }
                    if (mine) rowbest_p(A, S)[i] = last;
                }
            }
        }
        /* the window's bookkeeping, one row per lane */
        tm.sync();
        if (lane < nrows) {
            uint32_t v[4];
            lds_words<4>(wrec_a + lane * 16, v);
            const uint32_t d_tbo = lds_word(wrec_a + T * 16 + lane * 4);
            const int bs = (int)v[0], es = (int)v[1];
            const int pbeg = bs << lg, porg = org_of(pbeg);
            const int wd = min(WCAP, max(0, min(((es + 1) << lg) - 1, qlen) - porg + 1));
            rowinfo_g[w0 + lane] = make_int4(bs, es, (int)v[2], (int)v[3]);
            /* .x: where column (bs << lg) of the row is stored (the row itself starts at its origin column) */
            rowtb_g[w0 + lane] = make_uint4(d_tbo + (uint32_t)((pbeg - porg) >> 1), 2 * stw_of(wd), (uint32_t)m_p0, m_a & META_BASE);
        }
        tm.sync();
    }
    R.tbbytes = (unsigned long long)tb_used * 4;
    R.cells = cells; R.intops = 17ull * cells + 3ull * extra; R.full = (unsigned long long)(N - 2) * (qlen + 1);
    tm.sync();
    if (err == ST_OK) pick_best(A, S, N, qlen, R);
    return err;
}


}  // namespace mpoa
