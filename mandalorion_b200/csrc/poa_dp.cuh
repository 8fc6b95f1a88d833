/*
 * poa_dp.cuh -- adaptive-banded convex-gap DP of one read against the graph (rows in topological
 * order), two variants:
 *
 *   dp_align32      int32 lanes, one cell per lane, band of any width in chunks of 32 cells
 *   dp_align16<WPL> packed int16x2 (DPX: VIADDMNMX.S16x2 / VIMNMX3.S16x2), every lane owns 2*WPL
 *                   CONSECUTIVE cells (WPL 32-bit words), band <= 64*WPL cells.  The insertion
 *                   recurrence F is solved with (F1,F2) packed in one word: a serial pass over the
 *                   lane's cells, ONE decayed max-scan across lanes per row, one fix-up pass.
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

/* shared-memory words of one warp: RING rows x 3 arrays x (32*WPL + 2*RING_PAD), the query
 * profile of the rows' current column window (5 bases x 32*WPL words) and ring_info */
template <int WPL>
__host__ __device__ constexpr int ring16_row_words() { return 32 * WPL + 2 * RING_PAD; }
template <int WPL>
__host__ __device__ constexpr int ring16_rows() { return WPL >= 4 ? 4 : 8; }   // rows kept in shared memory
template <int WPL>
__host__ __device__ constexpr int ring16_warp_words() { return ring16_rows<WPL>() * 3 * ring16_row_words<WPL>() + 5 * 32 * WPL + RING * 4; }

template <int WPL>
__device__ __forceinline__ int dp_align16(const KernelArgs &A, const Slot &S, int N, const uint8_t *__restrict__ q,
                                          int qlen, uint32_t *ring, int4 *ring_info, int lane, AlnState &R) {
    constexpr int CPL = 2 * WPL;              // cells per lane
    constexpr int WCAP = 32 * CPL;            // cells per row
    constexpr int RW = ring16_row_words<WPL>();
    constexpr int PW = 32 * WPL;              // words of one profile row
    constexpr int RINGV = ring16_rows<WPL>();
    const DevParams &P = A.P;
    const Packed16 &K = A.K;
    const uint32_t *in_off = in_off_p(A, S), *in_row = in_row_p(A, S);
    uint32_t *tb = reinterpret_cast<uint32_t *>(tb_p(A, S));   // words = pairs of int16 cells
    const uint32_t tbcap = (uint32_t)min(A.L.tbcap / 4, (uint64_t)0xfffffff0u);
    lane_width_rule(P, qlen, N, R);
    if (R.bits != 16) return ST_RETRY_32;     // needs the int32 kernel
    const int lg = R.lgpn;
    const int w = P.wb < 0 ? qlen : P.wb + (int)__fmul_rn(P.wf, (float)qlen);
    uint32_t tb_used = 0;                     // words
    RowCount RC = {0, 0, 0};
    int err = ST_OK;                          // sticky: checked once per window of 32 rows

    const uint32_t NEG2 = K.neg2;
    const int wl0 = lane * WPL;               // first word of this lane inside a row
    const int col0 = lane * CPL;              // first cell
    uint32_t *prof = ring + RINGV * 3 * RW;    // [5][PW] match/mismatch words of the current window
    /* vector accesses of a lane's WPL words need the pred row shifted by whole lanes */
    const bool lanes_whole = ((R.pn >> 1) % WPL) == 0;

    /* row 0: the source */
    int4 prev_info;
    {
        const int rem0 = remain_p(A, S)[0];
        const int e = min(qlen, max(0, qlen - rem0) + w);
        const int end_sn = e >> lg;
        const int hi = min(((end_sn + 1) << lg) - 1, qlen);
        const int width = hi + 1;
        if (width > WCAP) return ST_RETRY_WIDE;
        const uint32_t stw = (uint32_t)(((width + 1) >> 1) + WPL - 1) / WPL * WPL;   // stored words per array
        if (3 * stw > tbcap) return ST_RETRY;
        uint32_t *Hs = ring + RING_PAD, *E1s = Hs + RW, *E2s = Hs + 2 * RW;
#pragma unroll
        for (int m = 0; m < WPL; ++m) {
            const int c = col0 + 2 * m;
            int hv[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int cc = c + t;
                hv[t] = cc == 0 ? 0 : max(max(-(P.o1 + P.e1 * cc), -(P.o2 + P.e2 * cc)), NEG16);
            }
            const uint32_t hw = pack2(hv[0], hv[1]);
            const uint32_t e1w = c == 0 ? pack2(-P.oe1, NEG16) : NEG2, e2w = c == 0 ? pack2(-P.oe2, NEG16) : NEG2;
            Hs[wl0 + m] = hw; E1s[wl0 + m] = e1w; E2s[wl0 + m] = e2w;
            if ((uint32_t)(wl0 + m) < stw) { tb[wl0 + m] = hw; tb[stw + wl0 + m] = e1w; tb[2 * stw + wl0 + m] = e2w; }
            prof[4 * PW + wl0 + m] = 0;       // a node base N scores 0 against everything
        }
        prev_info = make_int4(0, end_sn, 0, 0);
        if (lane == 0) {
            ring_info[0] = prev_info;
            rowinfo_p(A, S)[0] = prev_info;
            rowtb_p(A, S)[0] = make_uint4(0, 2 * stw, 0, 0);
        }
        tb_used = 3 * stw;
        __syncwarp();
    }

    int prof_beg = -1;                        // band start the profile was built for
    int4 *rowinfo_g = rowinfo_p(A, S);
    uint4 *rowtb_g = rowtb_p(A, S);

    for (int w0 = 1; w0 < N - 1; w0 += 32) {
        /* row metadata of 32 rows at once: a = base | flags | npre<<5 | remain<<13 */
        uint32_t m_a = 0;
        int m_in0 = 0, m_p0 = 0;
        {
            const int r = w0 + lane;
            if (r < N - 1) {
                m_in0 = (int)in_off[r];
                const int npre = (int)in_off[r + 1] - m_in0;
                m_a = (meta_p(A, S)[r] & 31u) | ((uint32_t)min(npre, 255) << 5) | ((uint32_t)remain_p(A, S)[r] << 13);
                m_p0 = npre > 0 ? (int)in_row[m_in0] : 0;
            }
        }
        const int nrows = min(32, N - 1 - w0);
        for (int l = 0; l < nrows; ++l) {
            const int i = w0 + l;
            const uint32_t ma = __shfl_sync(FULL, m_a, l);
            const int p0 = __shfl_sync(FULL, m_p0, l);
            const int nbase = ma & META_BASE;
            const int rem = (int)(ma >> 13);
            int npre = (ma >> 5) & 255;
            int in0 = 0;

            /* most rows have ONE predecessor and it is the previous row: straight-line path */
            const bool simple = npre == 1 && p0 == i - 1;
            int left, right, minb, maxe;
            if (simple) {
                left = min(N, prev_info.z + 1); right = max(0, prev_info.w + 1);
                minb = prev_info.x; maxe = prev_info.y;
            } else {
                in0 = __shfl_sync(FULL, m_in0, l);
                if (npre == 255) npre = (int)in_off[i + 1] - in0;
                left = N; right = 0; minb = INT_MAX; maxe = -1;
                for (int k = 0; k < npre; ++k) {
                    const int p = k == 0 ? p0 : (int)in_row[in0 + k];
                    const int4 pi = (p == i - 1) ? prev_info : ((i - p < RINGV) ? ring_info[p & (RINGV - 1)] : rowinfo_g[p]);
                    left = min(left, pi.z + 1);
                    right = max(right, pi.w + 1);
                    minb = min(minb, pi.x);
                    maxe = max(maxe, pi.y);
                }
            }
            const Band B = make_band(left, right, minb, maxe, rem, qlen, w, lg);
            int width = B.width;
            const int dp_beg = B.dp_beg;
            if (width > WCAP) { err = ST_RETRY_WIDE; width = 0; }
            const uint32_t stw = (uint32_t)(((width + 1) >> 1) + WPL - 1) / WPL * WPL;
            uint32_t tbo = tb_used;
            const bool ovf = tbcap - tb_used < 3 * stw;
            if (ovf) { if (err == ST_OK) err = ST_RETRY; tbo = 0; }
            else tb_used += 3 * stw;
            RC.add(width, npre);

            /* query profile of the window [dp_beg, dp_beg + WCAP): rebuilt only when the band start
             * moves (every ~pn rows); prof[b][word] = (s(b, q[j-1]), s(b, q[j])) for the word's cells */
            if (dp_beg != prof_beg) {
                prof_beg = dp_beg;
                __syncwarp();
#pragma unroll
                for (int m = 0; m < WPL; ++m) {
                    const int j0 = dp_beg + col0 + 2 * m;     // the word's first cell; it consumes q[j0-1]
                    const int qa = q[min(max(j0 - 1, 0), qlen - 1)], qb = q[min(max(j0, 0), qlen - 1)];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int sa = qa >= 4 ? 0 : (qa == b ? P.match : -P.mismatch);
                        const int sb = qb >= 4 ? 0 : (qb == b ? P.match : -P.mismatch);
                        prof[b * PW + wl0 + m] = pack2(sa, sb);
                    }
                }
                __syncwarp();
            }

            /* gather the diagonal and the deletion inputs from the predecessors.  A lane's WPL
             * words map to pred words [wp0, wp0+WPL); with the pred row shifted by whole lanes they
             * are all inside the pred's band or all outside: one vector load per array (no bank
             * conflicts), the word left of them comes from the neighbouring lane by shuffle. */
            uint32_t M2[WPL], EA[WPL], EB[WPL];
            auto gather = [&](const uint32_t *Hp, int st, int pw, int shw, bool vec, bool first) {
                const int wp0 = wl0 + shw;
                uint32_t hw[WPL], e1w[WPL], e2w[WPL];
                if (vec) {
                    const bool in = wp0 >= 0 && wp0 + WPL <= pw;
                    const uint32_t *src = Hp + (in ? wp0 : 0);
                    if constexpr (WPL == 2) {
                        const uint2 a = *reinterpret_cast<const uint2 *>(src);
                        const uint2 b = *reinterpret_cast<const uint2 *>(src + st);
                        const uint2 c = *reinterpret_cast<const uint2 *>(src + 2 * st);
                        hw[0] = a.x; hw[1] = a.y; e1w[0] = b.x; e1w[1] = b.y; e2w[0] = c.x; e2w[1] = c.y;
                    } else {
#pragma unroll
                        for (int m = 0; m < WPL; m += 4) {
                            const uint4 a = *reinterpret_cast<const uint4 *>(src + m);
                            const uint4 b = *reinterpret_cast<const uint4 *>(src + st + m);
                            const uint4 c = *reinterpret_cast<const uint4 *>(src + 2 * st + m);
                            hw[m] = a.x; hw[m + 1] = a.y; hw[m + 2] = a.z; hw[m + 3] = a.w;
                            e1w[m] = b.x; e1w[m + 1] = b.y; e1w[m + 2] = b.z; e1w[m + 3] = b.w;
                            e2w[m] = c.x; e2w[m + 1] = c.y; e2w[m + 2] = c.z; e2w[m + 3] = c.w;
                        }
                    }
                    if (!in) {
#pragma unroll
                        for (int m = 0; m < WPL; ++m) { hw[m] = NEG2; e1w[m] = NEG2; e2w[m] = NEG2; }
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < WPL; ++m) {
                        const int wp = wp0 + m;
                        const bool v = (unsigned)wp < (unsigned)pw;
                        hw[m] = v ? Hp[wp] : NEG2; e1w[m] = v ? Hp[st + wp] : NEG2; e2w[m] = v ? Hp[2 * st + wp] : NEG2;
                    }
                }
                /* word wp0-1: last word of the lane below, or a direct load at lane 0 */
                uint32_t hl = __shfl_up_sync(FULL, hw[WPL - 1], 1);
                if (lane == 0) hl = ((unsigned)(wp0 - 1) < (unsigned)pw) ? Hp[wp0 - 1] : NEG2;
                if (wp0 <= 0) hl = NEG2;
#pragma unroll
                for (int m = 0; m < WPL; ++m) {
                    /* an out-of-band word holds NEG2 in hw/e1w/e2w already; its diagonal must be NEG2 too */
                    const bool v = (unsigned)(wp0 + m) < (unsigned)pw;
                    const uint32_t dg = v ? __byte_perm(hl, hw[m], 0x5432) : NEG2;
                    if (first) { M2[m] = dg; EA[m] = e1w[m]; EB[m] = e2w[m]; }
                    else { M2[m] = __vmaxs2(M2[m], dg); EA[m] = __vmaxs2(EA[m], e1w[m]); EB[m] = __vmaxs2(EB[m], e2w[m]); }
                    hl = hw[m];
                }
            };
            if (simple) {
                const int shw = (dp_beg - (prev_info.x << lg)) >> 1;
                gather(ring + ((i - 1) & (RINGV - 1)) * 3 * RW + RING_PAD, RW,
                       ((prev_info.y - prev_info.x + 1) << lg) >> 1, shw, lanes_whole, true);
            } else {
#pragma unroll
                for (int m = 0; m < WPL; ++m) { M2[m] = NEG2; EA[m] = NEG2; EB[m] = NEG2; }
                for (int k = 0; k < npre; ++k) {
                    const int p = k == 0 ? p0 : (int)in_row[in0 + k];
                    const bool near = i - p < RINGV;
                    const int4 pi = (p == i - 1) ? prev_info : (near ? ring_info[p & (RINGV - 1)] : rowinfo_g[p]);
                    const int pw = ((pi.y - pi.x + 1) << lg) >> 1;             // words of the pred's rounded band
                    const int shw = (dp_beg - (pi.x << lg)) >> 1;              // our word 0 = pred word shw
                    if (near) gather(ring + (p & (RINGV - 1)) * 3 * RW + RING_PAD, RW, pw, shw, lanes_whole, false);
                    else {
                        const uint4 rt = rowtb_g[p];
                        const int pst = (int)(rt.y >> 1);
                        gather(tb + rt.x, pst, min(pw, pst), shw, false, false);
                    }
                }
            }
            /* first cell of the row's band: no diagonal at all */
            if (lane == 0) M2[0] = (M2[0] & 0xffff0000u) | (NEG2 & 0xffffu);

            /* match/mismatch scores from the profile row of this node's base */
            uint32_t S2[WPL];
            {
                const uint32_t *pr = prof + nbase * PW + wl0;
                if constexpr (WPL == 2) { const uint2 a = *reinterpret_cast<const uint2 *>(pr); S2[0] = a.x; S2[1] = a.y; }
                else if constexpr (WPL % 4 == 0) {
#pragma unroll
                    for (int m = 0; m < WPL; m += 4) {
                        const uint4 a = *reinterpret_cast<const uint4 *>(pr + m);
                        S2[m] = a.x; S2[m + 1] = a.y; S2[m + 2] = a.z; S2[m + 3] = a.w;
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < WPL; ++m) S2[m] = pr[m];
                }
            }

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
            FL[0] = NEG2;
#pragma unroll
            for (int t = 1; t < CPL; ++t) FL[t] = __viaddmax_s16x2(FL[t - 1], K.nee, X[t - 1]);
            uint32_t T = __viaddmax_s16x2(FL[CPL - 1], K.nee, X[CPL - 1]);
#pragma unroll
            for (int dd = 0; dd < 5; ++dd) {
                const uint32_t up = __shfl_up_sync(FULL, T, 1 << dd);
                T = __viaddmax_s16x2(up, K.dec[dd], T);
            }
            uint32_t C = __shfl_up_sync(FULL, T, 1);
            if (lane == 0) C = NEG2;
            uint32_t F1w[WPL], F2w[WPL];
            {
                uint32_t fa = 0;
#pragma unroll
                for (int t = 0; t < CPL; ++t) {
                    const uint32_t ff = __viaddmax_s16x2(C, K.tdec[t], FL[t]);
                    if (t & 1) {
                        F1w[t >> 1] = __byte_perm(fa, ff, 0x5410);
                        F2w[t >> 1] = __byte_perm(fa, ff, 0x7632);
                    } else fa = ff;
                }
            }

            uint32_t Hw[WPL], E1o[WPL], E2o[WPL];
#pragma unroll
            for (int m = 0; m < WPL; ++m) {
                Hw[m] = __vimax3_s16x2(HH[m], F1w[m], F2w[m]);
                E1o[m] = __viaddmax_s16x2(EA[m], K.ne1, __viaddmax_s16x2(Hw[m], K.noe1, NEG2));
                E2o[m] = __viaddmax_s16x2(EB[m], K.ne2, __viaddmax_s16x2(Hw[m], K.noe2, NEG2));
            }

            /* ring + HBM stores */
            {
                uint32_t *Hr = ring + (i & (RINGV - 1)) * 3 * RW + RING_PAD + wl0;
                uint32_t *g = tb + tbo + wl0;
                const bool st = !ovf && (uint32_t)wl0 < stw;
                if constexpr (WPL == 2) {
                    *reinterpret_cast<uint2 *>(Hr) = make_uint2(Hw[0], Hw[1]);
                    *reinterpret_cast<uint2 *>(Hr + RW) = make_uint2(E1o[0], E1o[1]);
                    *reinterpret_cast<uint2 *>(Hr + 2 * RW) = make_uint2(E2o[0], E2o[1]);
                    if (st) {
                        *reinterpret_cast<uint2 *>(g) = make_uint2(Hw[0], Hw[1]);
                        *reinterpret_cast<uint2 *>(g + stw) = make_uint2(E1o[0], E1o[1]);
                        *reinterpret_cast<uint2 *>(g + 2 * stw) = make_uint2(E2o[0], E2o[1]);
                    }
                } else if constexpr (WPL % 4 == 0) {
#pragma unroll
                    for (int m = 0; m < WPL; m += 4) {
                        const uint4 h4 = make_uint4(Hw[m], Hw[m + 1], Hw[m + 2], Hw[m + 3]);
                        const uint4 a4 = make_uint4(E1o[m], E1o[m + 1], E1o[m + 2], E1o[m + 3]);
                        const uint4 b4 = make_uint4(E2o[m], E2o[m + 1], E2o[m + 2], E2o[m + 3]);
                        *reinterpret_cast<uint4 *>(Hr + m) = h4;
                        *reinterpret_cast<uint4 *>(Hr + RW + m) = a4;
                        *reinterpret_cast<uint4 *>(Hr + 2 * RW + m) = b4;
                        if (st) {
                            *reinterpret_cast<uint4 *>(g + m) = h4;
                            *reinterpret_cast<uint4 *>(g + stw + m) = a4;
                            *reinterpret_cast<uint4 *>(g + 2 * stw + m) = b4;
                        }
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < WPL; ++m) {
                        Hr[m] = Hw[m]; Hr[RW + m] = E1o[m]; Hr[2 * RW + m] = E2o[m];
                        if (st) { g[m] = Hw[m]; g[stw + m] = E1o[m]; g[2 * stw + m] = E2o[m]; }
                    }
                }
            }

            /* row maximum with its left-most and right-most column: two keyed warp reductions */
            int kr = INT_MIN, kl = INT_MIN;
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int c = col0 + t;
                const int hv = (t & 1) ? (int)(Hw[t >> 1] & 0xffff0000u) : (int)(Hw[t >> 1] << 16);
                if (c < width) {
                    kr = max(kr, hv | c);
                    kl = max(kl, hv | (0xffff - c));
                }
            }
            kr = __reduce_max_sync(FULL, kr);
            kl = __reduce_max_sync(FULL, kl);
            int lpos = -1, rpos = -1;
            if (width > 0) { rpos = dp_beg + (kr & 0xffff); lpos = dp_beg + (0xffff - (kl & 0xffff)); }
            prev_info = make_int4(B.beg_sn, B.end_sn, lpos, rpos);
            if (lane == 0) {
                ring_info[i & (RINGV - 1)] = prev_info;
                rowinfo_g[i] = prev_info;
                rowtb_g[i] = make_uint4(tbo, 2 * stw, (uint32_t)p0, (uint32_t)nbase);
            }
            if (ma & META_TOSINK) {
                /* H at the last cell of the row: the global best is picked among these after the DP */
                int last = NEG;
                bool mine = width == 0 && lane == 0;
#pragma unroll
                for (int t = 0; t < CPL; ++t)
                    if (col0 + t == width - 1) { last = (t & 1) ? hi16(Hw[t >> 1]) : lo16(Hw[t >> 1]); mine = true; }
                if (mine) rowbest_p(A, S)[i] = last;
            }
            __syncwarp();
        }
        if (err != ST_OK) return err;
    }
    R.tbbytes = (unsigned long long)tb_used * 4;
    RC.flush(R, qlen);
    __syncwarp();
    pick_best(A, S, N, qlen, R);
    return ST_OK;
}


}  // namespace mpoa
