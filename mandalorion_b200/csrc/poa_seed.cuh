/*
 * poa_seed.cuh -- device side of `abpoa -S` (reference utils/SpliceDefineConsensus.py:919: groups
 * whose median read length is >= 8000): the read is not aligned to the whole graph but window by
 * window between ANCHORS, exact k-mers it shares with the previous read of the group.  The anchors
 * come from the host (csrc/seed.cpp); an anchor k-mer becomes a forced run of matches, and the
 * stretch of the read between two anchors is aligned to the SUB-GRAPH between the anchor nodes.
 *
 * extract_window() builds that sub-graph as a VIEW (Slot.sub = 1): the rows that lie on a path from
 * the begin row to the end row, renumbered in row order, with their in-edges (edge order kept),
 * base / to-sink flags and remain values relative to the end row.  The ordinary DP and traceback
 * then run on the view unchanged -- the begin row plays the source, the end row the sink.
 */
#pragma once
#include "poa_graph.cuh"

namespace mpoa {

/*
 * Rows rb .. re of the current graph (rb < re).  Scratch: the merge temporaries cnt (bit 0 reachable
 * from rb, bit 1 reaches re, bit 2 has an edge to re), addin (row -> sub-row), srcof (sub-row -> row).
 * Returns the number of sub-rows (>= 2) or 0 when re cannot be reached from rb.
 */
template <int T>
__device__ __forceinline__ int extract_window(const KernelArgs &A, const Slot &S, const Team<T> &tm, int rb, int re, bool on) {
    const int lane = tm.tl;
    const uint32_t *in_off = in_off_p(A, S), *in_row = in_row_p(A, S);
    const uint32_t *out_off = out_off_p(A, S), *out_row = out_row_p(A, S);
    int32_t *flag = cnt_p(A, S), *subidx = addin_p(A, S), *sub2row = srcof_p(A, S);
    Slot V = S;
    V.sub = 1;
    uint32_t *s_in_off = in_off_p(A, V), *s_in_row = in_row_p(A, V), *s_meta = meta_p(A, V);
    int32_t *s_remain = remain_p(A, V);
    const int32_t *remain = remain_p(A, S);
    const uint8_t *base = base_p(A, S);
    const int span = on ? re - rb : 0;                 // rows rb+1 .. re
    const int nch = tm.wmax((span + T - 1) / T);

    /* forward: reachable from rb */
    if (on && lane == 0) flag[rb] = 1;
    tm.sync();
    for (int c = 0; c < nch; ++c) {
        const int c0 = rb + 1 + c * T, r = c0 + lane;
        const bool valid = on && r <= re;
        bool m = false;
        uint32_t inmask = 0;
        if (valid) {
#pragma unroll 1
            for (uint32_t e = in_off[r]; e < in_off[r + 1]; ++e) {
                const int p = (int)in_row[e];
                if (p == rb) m = true;
                else if (p > rb) { if (p >= c0) inmask |= 1u << (p - c0); else m = m || (flag[p] & 1); }
            }
        }
#pragma unroll 1
        for (int it = 0; it < T; ++it) {               // a chain inside the chunk needs up to T-1 rounds
            const unsigned bal = tm.ballot(m);
            const bool nm = m || (inmask & bal) != 0;
            const bool changed = nm != m;
            m = nm;
            if (!tm.any_lane(changed)) break;
        }
        if (valid) flag[r] = m ? 1 : 0;
        tm.sync();
    }
    /* backward: reaches re */
    if (on && lane == 0) flag[re] |= 2;
    tm.sync();
    for (int c = 0; c < nch; ++c) {
        const int c1 = re - 1 - c * T, r = c1 - lane;
        const bool valid = on && r >= rb;
        bool m = false, tosink = false;
        uint32_t outmask = 0;
        if (valid) {
#pragma unroll 1
            for (uint32_t e = out_off[r]; e < out_off[r + 1]; ++e) {
                const int s = (int)out_row[e];
                if (s == re) { m = true; tosink = true; }
                else if (s < re) { if (s <= c1) outmask |= 1u << (c1 - s); else m = m || ((flag[s] >> 1) & 1); }
            }
        }
#pragma unroll 1
        for (int it = 0; it < T; ++it) {
            const unsigned bal = tm.ballot(m);
            const bool nm = m || (outmask & bal) != 0;
            const bool changed = nm != m;
            m = nm;
            if (!tm.any_lane(changed)) break;
        }
        if (valid) flag[r] = (flag[r] & 1) | (m ? 2 : 0) | (tosink ? 4 : 0);
        tm.sync();
    }
    const bool reach = on && (flag[re] & 1) != 0;

    /* members in row order: sub-row numbers, in-edge counts, per-row data of the view */
    const int nch2 = tm.wmax(reach ? (re - rb + T) / T : 0);   // rows rb .. re
    int carry_n = 0, carry_e = 0;
    for (int c = 0; c < nch2; ++c) {
        const int r = rb + c * T + lane;
        const bool valid = reach && r <= re;
        const bool mem = valid && (flag[r] & 3) == 3;
        int deg = 0;
        if (mem && r != rb) {
#pragma unroll 1
            for (uint32_t e = in_off[r]; e < in_off[r + 1]; ++e) {
                const int p = (int)in_row[e];
                if (p >= rb && (flag[p] & 3) == 3) ++deg;
            }
        }
        const int in_n = team_incl_sum(tm, mem ? 1 : 0), in_e = team_incl_sum(tm, deg);
        if (mem) {
            const int idx = carry_n + in_n - 1;
            subidx[r] = idx;
            sub2row[idx] = r;
            s_in_off[idx] = (uint32_t)(carry_e + in_e - deg);
            s_meta[idx] = (uint32_t)base[r] | ((flag[r] & 4) ? META_TOSINK : 0);
            s_remain[idx] = remain[r] - remain[re] - 1;
        }
        carry_n += tm.shfl(in_n, T - 1);
        carry_e += tm.shfl(in_e, T - 1);
    }
    if (reach && lane == 0) s_in_off[carry_n] = (uint32_t)carry_e;
    tm.sync();
    /* in-edges of the view: member predecessors, original edge order */
    for (int c = 0; c < nch2; ++c) {                   // no cross-lane operation
        const int r = rb + c * T + lane;
        if (reach && r <= re && r != rb && (flag[r] & 3) == 3) {
            uint32_t eo = s_in_off[subidx[r]];
#pragma unroll 1
            for (uint32_t e = in_off[r]; e < in_off[r + 1]; ++e) {
                const int p = (int)in_row[e];
                if (p >= rb && (flag[p] & 3) == 3) s_in_row[eo++] = (uint32_t)subidx[p];
            }
        }
    }
    tm.sync();
    return reach ? carry_n : 0;
}

}  // namespace mpoa
