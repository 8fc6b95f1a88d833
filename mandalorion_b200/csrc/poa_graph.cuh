/*
 * poa_graph.cuh -- the POA graph in row space: workspace accessors, warp scans, first read,
 * remain pass, merge of an aligned read, heaviest-bundle consensus.
 * Semantics follow abPOA v1.4.1 as invoked by the reference (utils/SpliceDefineConsensus.py:917):
 * add_graph_sequence / add_subgraph_alignment / BFS_set_node_remain / heaviest_bundling.
 */
#pragma once
#include <climits>
#include "poa_device.cuh"

namespace mpoa {

/* One warp's HBM workspace.  Only the base pointer is held in registers: every array address is
 * base + an offset read from the kernel parameter block (constant bank), so the ~35 array
 * pointers never occupy registers or spill to local memory.  The graph is double buffered:
 * x_p() is the CURRENT graph, n_x_p() the buffer the next merge writes. */
struct Slot {
    uint8_t *b;
    int par;
    int sub;      // 1: the sub-graph view of a seeded window (in-edges, meta, remain, qmap)
};

#define MPOA_ACC(T, name)                                                                       \
    __device__ __forceinline__ T *name##_p(const KernelArgs &A, const Slot &S) {                \
        return reinterpret_cast<T *>(S.b + A.L.name);                                           \
    }
#define MPOA_ACC2(T, name)                                                                      \
    __device__ __forceinline__ T *name##_p(const KernelArgs &A, const Slot &S) {                \
        return reinterpret_cast<T *>(S.b + A.L.name[S.par]);                                    \
    }                                                                                           \
    __device__ __forceinline__ T *n_##name##_p(const KernelArgs &A, const Slot &S) {            \
        return reinterpret_cast<T *>(S.b + A.L.name[S.par ^ 1]);                                \
    }
/* arrays that exist for the graph (S.sub = 0) and for the sub-graph view of a seeded window (S.sub = 1) */
#define MPOA_ACCV(T, name, subidx)                                                              \
    __device__ __forceinline__ T *name##_p(const KernelArgs &A, const Slot &S) {                \
        return reinterpret_cast<T *>(S.b + (S.sub ? A.L.name[subidx] : A.L.name[0]));           \
    }
MPOA_ACC2(uint8_t, base) MPOA_ACC2(uint8_t, sib) MPOA_ACC2(int32_t, creator)
MPOA_ACC2(uint32_t, out_off) MPOA_ACC2(uint32_t, out_row) MPOA_ACC2(int32_t, out_w)
__device__ __forceinline__ uint32_t *in_off_p(const KernelArgs &A, const Slot &S) { return reinterpret_cast<uint32_t *>(S.b + A.L.in_off[S.sub ? 2 : S.par]); }
__device__ __forceinline__ uint32_t *in_row_p(const KernelArgs &A, const Slot &S) { return reinterpret_cast<uint32_t *>(S.b + A.L.in_row[S.sub ? 2 : S.par]); }
__device__ __forceinline__ uint32_t *n_in_off_p(const KernelArgs &A, const Slot &S) { return reinterpret_cast<uint32_t *>(S.b + A.L.in_off[S.par ^ 1]); }
__device__ __forceinline__ uint32_t *n_in_row_p(const KernelArgs &A, const Slot &S) { return reinterpret_cast<uint32_t *>(S.b + A.L.in_row[S.par ^ 1]); }
MPOA_ACCV(int32_t, remain, 1) MPOA_ACCV(uint32_t, meta, 1) MPOA_ACCV(int32_t, qmap, 1)
MPOA_ACC(int4, rowinfo) MPOA_ACC(uint4, rowtb) MPOA_ACC(int32_t, prevrow)
MPOA_ACC(int32_t, rowbest) MPOA_ACC(uint32_t, qprof)
MPOA_ACC(int32_t, pv) MPOA_ACC(int32_t, pkey) MPOA_ACC(int32_t, pnew) MPOA_ACC(int32_t, psib)
MPOA_ACC(int32_t, nin) MPOA_ACC(int32_t, nout)
MPOA_ACC(int32_t, cnt) MPOA_ACC(int32_t, addin) MPOA_ACC(int32_t, addout) MPOA_ACC(int32_t, srcof)
MPOA_ACC(uint8_t, grow) MPOA_ACC(uint8_t, tb)

__device__ __forceinline__ Slot make_slot(const KernelArgs &A, int slot, int par) {
    Slot S;
    S.b = A.ws + (uint64_t)slot * A.L.slot_bytes;
    asm volatile("" : "+l"(S.b));   // keep the pointer in registers (it is otherwise re-derived at every access)
    __builtin_assume(__isGlobal(S.b));
    S.par = par;
    S.sub = 0;
    return S;
}

template <int T>
__device__ __forceinline__ int team_incl_sum(const Team<T> &tm, int v) {
#pragma unroll
    for (int d = 1; d < T; d <<= 1) {
        int t = tm.shfl_up(v, d);
        if (tm.tl >= d) v += t;
    }
    return v;
}

template <int T>
__device__ __forceinline__ int team_incl_max(const Team<T> &tm, int v) {
#pragma unroll
    for (int d = 1; d < T; d <<= 1) {
        int t = tm.shfl_up(v, d);
        if (tm.tl >= d) v = max(v, t);
    }
    return v;
}

/* inclusive scan of  S[l] = max_{k<=l} (x[k] - e*(l-k))  -- the insertion (F) recurrence */
__device__ __forceinline__ int warp_scan_decay(int v, int e, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v = max(v, t - d * e);
    }
    return v;
}

enum MetaBits { META_BASE = 7, META_TOSINK = 16 };

/* ------------------------------------------------------------------------------------------ */
/* graph: first read, remain pass                                                              */
/* ------------------------------------------------------------------------------------------ */

/* first read = linear chain src -> b0 -> ... -> sink, every edge weight 1 */
template <int T, bool TRACE, bool SEEDED>
__device__ __forceinline__ void init_graph(const KernelArgs &A, const Slot &S, const Team<T> &tm, const uint8_t *seq, int len, int creator0) {
    const int N = len + 2;
    for (int r = tm.tl; r <= N; r += T) {
        if (r < N) {
            const bool real = r >= 1 && r <= len;
            if constexpr (SEEDED) { if (real) prevrow_p(A, S)[r - 1] = r; }
            base_p(A, S)[r] = real ? seq[r - 1] : 0;
            sib_p(A, S)[r] = 0;
            if constexpr (TRACE) creator_p(A, S)[r] = real ? creator0 + r - 1 : -1;   // node identity: only the trace needs it
            if (r >= 1) in_row_p(A, S)[r - 1] = r - 1;
            if (r < N - 1) { out_row_p(A, S)[r] = r + 1; out_w_p(A, S)[r] = 1; }
        }
        in_off_p(A, S)[r] = r == 0 ? 0 : r - 1;
        out_off_p(A, S)[r] = r < N - 1 ? r : N - 1;
    }
    /* no cross-lane operation in here: the caller synchronises (the teams of a warp may not both call this) */
}

/*
 * remain[r] = remain[heaviest out-neighbour (first maximum)] + 1, remain[sink] = -1: what abPOA's
 * reverse BFS computes.  Rows are handled 32 at a time from the sink side; in-window chains are
 * resolved by pointer jumping.  Also emits the per-row meta word used by the DP.
 */
template <int T>
__device__ __forceinline__ void remain_pass(const KernelArgs &A, const Slot &S, const Team<T> &tm, int N, bool on) {
    const int lane = tm.tl;
    constexpr int LOGT = T == 32 ? 5 : 4;
    const uint32_t *out_off = out_off_p(A, S), *out_row = out_row_p(A, S);
    const int32_t *out_w = out_w_p(A, S);
    if (on && lane == 0) {
        remain_p(A, S)[N - 1] = -1;
        meta_p(A, S)[N - 1] = 0;
    }
    const int nwin = on ? (N - 2) / T + 1 : 0;
    const int wn = tm.wmax(nwin);
    int prev_vals = 0;  // remain of rows [w0+T, w0+2T)
    for (int it = 0; it < wn; ++it) {
        const int w0 = ((N - 2) / T - it) * T;
        const int r = w0 + lane;
        const bool active = it < nwin && r <= N - 2;
        int hs = N - 1, tosink = 0;
        if (active) {
            const uint32_t o0 = out_off[r], o1 = out_off[r + 1];
            int maxw = -1;
#pragma unroll 1
            for (uint32_t e = o0; e < o1; ++e) {
                const int t = (int)out_row[e], w = out_w[e];
                if (w > maxw) { maxw = w; hs = t; }
                tosink |= (t == N - 1);
            }
        }
        int acc, nxt = -1;
        const int src_prev = min(T - 1, max(0, hs - (w0 + T)));
        const int from_prev = tm.shfl(prev_vals, src_prev);
        if (!active) acc = 0;
        else if (hs == N - 1) acc = 0;
        else if (hs >= w0 + 2 * T) acc = remain_p(A, S)[hs] + 1;
        else if (hs >= w0 + T) acc = from_prev + 1;
        else { acc = 1; nxt = hs - w0; }
#pragma unroll
        for (int k = 0; k < LOGT; ++k) {
            const int sl = max(nxt, 0);
            const int a = tm.shfl(acc, sl);
            const int n = tm.shfl(nxt, sl);
            if (nxt >= 0) { acc += a; nxt = n; }
        }
        if (active) {
            remain_p(A, S)[r] = acc;
            meta_p(A, S)[r] = (uint32_t)base_p(A, S)[r] | (tosink ? META_TOSINK : 0);
        }
        prev_vals = acc;
        tm.sync();
    }
}

/* ------------------------------------------------------------------------------------------ */
/* graph merge                                                                                 */
/* ------------------------------------------------------------------------------------------ */

/*
 * Adds the aligned read to the graph (semantics of abPOA's add_subgraph_alignment) and re-emits
 * the graph into the other buffer with the new nodes merged into the row order:
 *   - a base aligned to a node with the same base, or to an aligned sibling with the same base,
 *     reuses that node; otherwise it becomes a new node (a new sibling if it was aligned);
 *   - new nodes are placed after the end of the sibling group of the previous path node
 *     (new siblings: after the end of the group they join), in path order;
 *   - path edges that exist get weight +1, the others are appended to the END of the edge lists
 *     of their endpoints (edge order = first-creation order).
 * Returns ST_OK or ST_RETRY (capacity).
 */
template <int T, bool SEEDED = false>
__device__ __forceinline__ int merge_read(const KernelArgs &A, const Slot &S, const Team<T> &tm, int &par, int &N_io, int &E_io,
                                          const uint8_t *__restrict__ q, int qlen_in, int creator0, int32_t *tr_aln, int32_t *tr_node, bool on) {
    const int lane = tm.tl;
    /* a team that is switched off walks through the passes with empty ranges */
    int qlen = on ? qlen_in : 0, N = on ? N_io : 0;
    int err = ST_OK;
    const uint8_t *__restrict__ base = base_p(A, S), *__restrict__ sib = sib_p(A, S);
    const uint32_t *__restrict__ out_off = out_off_p(A, S), *__restrict__ out_row = out_row_p(A, S);
    const uint32_t *__restrict__ in_off = in_off_p(A, S), *__restrict__ in_row = in_row_p(A, S);
    int32_t *out_w = out_w_p(A, S);

    /* The passes are bound by memory latency, so every pass handles MK*T elements per iteration and
     * is written LEVEL BY LEVEL: all loads of one dependency level (for all MK sub-chunks) are issued
     * before anything that depends on them, and no store sits between the levels.  The common case
     * (at most two in- and out-edges per row, a base that matches its aligned node) is covered by the
     * batched levels; the rest falls into short serial loops. */
#ifndef MPOA_MERGE_MK16
#define MPOA_MERGE_MK16 2
#endif
#ifndef MPOA_MERGE_MK32
#define MPOA_MERGE_MK32 2
#endif
    constexpr int MK = T == 16 ? MPOA_MERGE_MK16 : MPOA_MERGE_MK32;
    int32_t *cnt = cnt_p(A, S), *addin = addin_p(A, S), *addout = addout_p(A, S), *srcof = srcof_p(A, S);
    int32_t *pv = pv_p(A, S), *pkey = pkey_p(A, S), *pnew = pnew_p(A, S), *psib = psib_p(A, S);
    int32_t *nin = nin_p(A, S), *nout = nout_p(A, S);
    const int32_t *qmap = qmap_p(A, S), *creator = creator_p(A, S);
    uint8_t *grow = grow_p(A, S);

#pragma unroll 1
    for (int r = lane; r < N; r += T) { cnt[r] = 0; addin[r] = -1; addout[r] = -1; grow[r] = 0; }
    tm.sync();
    const int wq = tm.wmax(qlen);

    /* U1: resolve every query base to an existing row or a new node; order keys */
    int carry_key = 0, carry_new = 0;
    for (int t0 = 0; t0 < wq; t0 += T * MK) {
        int isnew[MK], v[MK], key[MK], sibof[MK];
        int r_[MK], b_[MK], br_[MK], sr_[MK];
#pragma unroll
        for (int u = 0; u < MK; ++u) {        // level 1: the aligned row and the query base
            const int t = t0 + u * T + lane;
            r_[u] = -1; b_[u] = 0;
            if (t < qlen) { r_[u] = qmap[t]; b_[u] = q[t]; }
        }
#pragma unroll
        for (int u = 0; u < MK; ++u) {        // level 2: that row's base and sibling byte
            br_[u] = 0; sr_[u] = 0;
            if (r_[u] >= 0) { br_[u] = base[r_[u]]; sr_[u] = sib[r_[u]]; }
        }
#pragma unroll
        for (int u = 0; u < MK; ++u) {
            const int t = t0 + u * T + lane;
            isnew[u] = 0; v[u] = -1; key[u] = -1; sibof[u] = -1;
            if (t < qlen) {
                const int r = r_[u];
                const int b = b_[u];
                if (r >= 0) {
                    if (br_[u] == b) { v[u] = r; key[u] = r + (sr_[u] & 15); }
                    else {                    // rare: look for the base among the aligned siblings
                        const int sb = sr_[u], before = sb >> 4, after = sb & 15;
#pragma unroll 1
                        for (int x = r - before; x <= r + after; ++x)
                            if (x != r && base[x] == b) v[u] = x;
                        if (v[u] < 0) { isnew[u] = 1; sibof[u] = r; key[u] = r + after; }
                        else key[u] = v[u] + (sib[v[u]] & 15);
                    }
                    if (tr_aln) tr_aln[t] = creator[r];
                } else {
                    isnew[u] = 1;
                    if (tr_aln) tr_aln[t] = -1;
                }
                if (tr_node) tr_node[t] = isnew[u] ? creator0 + t : creator[v[u]];
            }
        }
#pragma unroll
        for (int u = 0; u < MK; ++u) {
            const int t = t0 + u * T + lane;
            int ks = team_incl_max(tm, key[u]);
            ks = max(ks, carry_key);
            const int incl = team_incl_sum(tm, isnew[u]);
            const int nidx = carry_new + incl - isnew[u];
            if (t < qlen) {
                pv[t] = isnew[u] ? -1 : v[u];
                pkey[t] = ks;
                pnew[t] = nidx;
                psib[t] = sibof[u];
                if (isnew[u]) atomicAdd(&cnt[ks], 1);
                if (sibof[u] >= 0) {
                    const int sb = sib[sibof[u]];
#pragma unroll 1
                    for (int x = sibof[u] - (sb >> 4); x <= sibof[u] + (sb & 15); ++x) grow[x] = 1;
                }
            }
            carry_key = tm.shfl(ks, T - 1);
            carry_new += tm.shfl(incl, T - 1);
        }
    }
    const int n_new = carry_new;
    int N2 = N + n_new;
    if ((uint32_t)N2 > A.L.ncap) { err = ST_RETRY; qlen = 0; N = 0; N2 = 0; }
    tm.sync();

    /* U2: shift[r] = number of new nodes placed before old row r (exclusive scan of cnt) */
    {
        int carry = 0;
        const int wn = tm.wmax(N);
        for (int r0 = 0; r0 < wn; r0 += T * MK) {
            int c[MK];
#pragma unroll
            for (int u = 0; u < MK; ++u) { const int r = r0 + u * T + lane; c[u] = r < N ? cnt[r] : 0; }
#pragma unroll
            for (int u = 0; u < MK; ++u) {
                const int r = r0 + u * T + lane;
                const int incl = team_incl_sum(tm, c[u]);
                if (r < N) {
                    const int sh = carry + incl - c[u];
                    cnt[r] = sh;
                    srcof[r + sh] = r;
                }
                carry += tm.shfl(incl, T - 1);
            }
        }
        for (int t0 = 0; t0 < qlen; t0 += T * MK) {   // no cross-lane operation in this loop
            int pvv[MK], pk[MK], pn[MK];
#pragma unroll
            for (int u = 0; u < MK; ++u) {
                const int t = t0 + u * T + lane;
                pvv[u] = 0; pk[u] = 0; pn[u] = 0;
                if (t < qlen) { pvv[u] = pv[t]; pk[u] = pkey[t]; pn[u] = pnew[t]; }
            }
#pragma unroll
            for (int u = 0; u < MK; ++u) {
                const int t = t0 + u * T + lane;
                if (t < qlen && pvv[u] < 0) srcof[pk[u] + 1 + pn[u]] = -(t + 1);
            }
        }
    }
    tm.sync();

    /* U3: the path edges u[t-1] -> u[t], t = 0..qlen (u[-1] = source, u[qlen] = sink) */
    int n_new_edges = 0;
    for (int t0 = 0; t0 <= qlen && N > 0; t0 += T * MK) {   // no cross-lane operation in this loop
        int from_old[MK], to_old[MK], km[MK], nm[MK], kt[MK], nt[MK];
        int cf[MK], ct[MK], e0[MK], w0[MK];
        uint32_t o0[MK], o1[MK];
#pragma unroll
        for (int u = 0; u < MK; ++u) {        // level 1: the two path nodes of the edge
            const int t = t0 + u * T + lane;
            from_old[u] = to_old[u] = -2; km[u] = nm[u] = kt[u] = nt[u] = 0;
            if (t <= qlen) {
                from_old[u] = 0; to_old[u] = N - 1;
                if (t > 0) { from_old[u] = pv[t - 1]; km[u] = pkey[t - 1]; nm[u] = pnew[t - 1]; }
                if (t < qlen) { to_old[u] = pv[t]; kt[u] = pkey[t]; nt[u] = pnew[t]; }
            }
        }
#pragma unroll
        for (int u = 0; u < MK; ++u) {        // level 2: their shifts, the out-edge list of the source node
            cf[u] = ct[u] = 0; o0[u] = o1[u] = 0;
            if (from_old[u] >= 0) cf[u] = cnt[from_old[u]];
            if (to_old[u] >= 0) ct[u] = cnt[to_old[u]];
            if (from_old[u] >= 0 && to_old[u] >= 0) { o0[u] = out_off[from_old[u]]; o1[u] = out_off[from_old[u] + 1]; }
        }
#pragma unroll
        for (int u = 0; u < MK; ++u) {        // level 3: the first out-edge
            e0[u] = -1;
            if (o1[u] > o0[u]) e0[u] = (int)out_row[o0[u]];
        }
#pragma unroll
        for (int u = 0; u < MK; ++u) {        // level 4: its weight when it is the path edge
            w0[u] = 0;
            if (o1[u] > o0[u] && e0[u] == to_old[u]) w0[u] = out_w[o0[u]];
        }
#pragma unroll
        for (int u = 0; u < MK; ++u) {
            const int t = t0 + u * T + lane;
            if (t <= qlen) {
                const int from_new = from_old[u] >= 0 ? from_old[u] + cf[u] : km[u] + 1 + nm[u];
                const int to_new = to_old[u] >= 0 ? to_old[u] + ct[u] : kt[u] + 1 + nt[u];
                if constexpr (SEEDED) { if (t < qlen) prevrow_p(A, S)[t] = to_new; }   // where base t of this read lives in the merged graph
                bool found = false;
                if (o1[u] > o0[u]) {
                    if (e0[u] == to_old[u]) { out_w[o0[u]] = w0[u] + 1; found = true; }
                    else {
#pragma unroll 1
                        for (uint32_t e = o0[u] + 1; e < o1[u]; ++e)
                            if ((int)out_row[e] == to_old[u]) { out_w[e] += 1; found = true; break; }
                    }
                }
                if (!found) {
                    ++n_new_edges;
                    if (from_old[u] >= 0) addout[from_old[u]] = to_new; else nout[t - 1] = to_new;
                    if (to_old[u] >= 0) addin[to_old[u]] = from_new; else nin[t] = from_new;
                }
            }
        }
    }
#pragma unroll
    for (int d = T / 2; d > 0; d >>= 1) n_new_edges += tm.shfl_xor(n_new_edges, d);
    const int E2 = E_io + n_new_edges;
    if (N > 0 && (uint32_t)E2 > A.L.ecap) { err = ST_RETRY; qlen = 0; N = 0; N2 = 0; }
    tm.sync();

    /* U4: emit the merged graph */
    {
        int carry_in = 0, carry_out = 0;
        uint32_t *n_in_off = n_in_off_p(A, S), *n_out_off = n_out_off_p(A, S), *n_in_row = n_in_row_p(A, S),
                 *n_out_row = n_out_row_p(A, S);
        int32_t *n_out_w = n_out_w_p(A, S), *n_creator = n_creator_p(A, S);
        uint8_t *n_base = n_base_p(A, S), *n_sib = n_sib_p(A, S);
        const int wn2 = tm.wmax(N2);
        for (int r0 = 0; r0 < wn2; r0 += T * MK) {
            int din[MK], dout[MK], src[MK], ai[MK], ao[MK], nin_old[MK], nout_old[MK];
            uint32_t i0[MK], o0[MK];
            int xa[MK], xb[MK], ya[MK], yb[MK], wa[MK], wb[MK];      // first two in- / out-edges (+ weights)
            int nb[MK], sbv[MK], gr[MK], cre[MK];                     // base, sibling byte (+ growth), creator
            int so[MK];
#pragma unroll
            for (int u = 0; u < MK; ++u) {    // level 1: which old row (or which new node) lands here
                const int nr = r0 + u * T + lane;
                src[u] = 0;
                if (nr < N2) src[u] = srcof[nr];
            }
#pragma unroll
            for (int u = 0; u < MK; ++u) {    // level 2: its edge lists, added edges, node bytes
                const int nr = r0 + u * T + lane;
                din[u] = dout[u] = 0; ai[u] = ao[u] = -1; i0[u] = o0[u] = 0; nin_old[u] = nout_old[u] = 0;
                nb[u] = sbv[u] = gr[u] = 0; cre[u] = 0; so[u] = -1;
                xa[u] = xb[u] = ya[u] = yb[u] = 0; wa[u] = wb[u] = 0;
                if (nr < N2) {
                    const int sr = src[u];
                    if (sr >= 0) {
                        i0[u] = in_off[sr]; o0[u] = out_off[sr];
                        nin_old[u] = (int)(in_off[sr + 1] - i0[u]); nout_old[u] = (int)(out_off[sr + 1] - o0[u]);
                        ai[u] = addin[sr]; ao[u] = addout[sr];
                        nb[u] = base[sr]; sbv[u] = sib[sr]; gr[u] = grow[sr];
                        if (tr_node) cre[u] = creator[sr];
                    } else {
                        const int t = -sr - 1;
                        xa[u] = nin[t]; ya[u] = nout[t]; nb[u] = q[t]; so[u] = psib[t];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < MK; ++u) {    // level 3: the first two old edges on each side
                if (src[u] >= 0) {
                    if (nin_old[u] > 0) xa[u] = (int)in_row[i0[u]];
                    if (nin_old[u] > 1) xb[u] = (int)in_row[i0[u] + 1];
                    if (nout_old[u] > 0) { ya[u] = (int)out_row[o0[u]]; wa[u] = out_w[o0[u]]; }
                    if (nout_old[u] > 1) { yb[u] = (int)out_row[o0[u] + 1]; wb[u] = out_w[o0[u] + 1]; }
                } else if (so[u] >= 0) sbv[u] = sib[so[u]];
            }
#pragma unroll
            for (int u = 0; u < MK; ++u) {    // level 4: the shifts of their end points
                if (src[u] >= 0) {
                    if (nin_old[u] > 0) xa[u] += cnt[xa[u]];
                    if (nin_old[u] > 1) xb[u] += cnt[xb[u]];
                    if (nout_old[u] > 0) ya[u] += cnt[ya[u]];
                    if (nout_old[u] > 1) yb[u] += cnt[yb[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < MK; ++u) {
                const int nr = r0 + u * T + lane;
                if (nr < N2) {
                    if (src[u] >= 0) { din[u] = nin_old[u] + (ai[u] >= 0); dout[u] = nout_old[u] + (ao[u] >= 0); }
                    else din[u] = dout[u] = 1;
                }
                const int iin = team_incl_sum(tm, din[u]), iout = team_incl_sum(tm, dout[u]);
                if (nr < N2) {
                    uint32_t io = carry_in + iin - din[u], oo = carry_out + iout - dout[u];
                    n_in_off[nr] = io;
                    n_out_off[nr] = oo;
                    const int sr = src[u];
                    if (sr >= 0) {
                        if (nin_old[u] > 0) n_in_row[io++] = xa[u];
                        if (nin_old[u] > 1) n_in_row[io++] = xb[u];
#pragma unroll 1
                        for (int e = 2; e < nin_old[u]; ++e) {
                            const int x = (int)in_row[i0[u] + e];
                            n_in_row[io++] = x + cnt[x];
                        }
                        if (ai[u] >= 0) n_in_row[io++] = ai[u];
                        if (nout_old[u] > 0) { n_out_row[oo] = ya[u]; n_out_w[oo++] = wa[u]; }
                        if (nout_old[u] > 1) { n_out_row[oo] = yb[u]; n_out_w[oo++] = wb[u]; }
#pragma unroll 1
                        for (int e = 2; e < nout_old[u]; ++e) {
                            const int y = (int)out_row[o0[u] + e];
                            n_out_row[oo] = y + cnt[y];
                            n_out_w[oo++] = out_w[o0[u] + e];
                        }
                        if (ao[u] >= 0) { n_out_row[oo] = ao[u]; n_out_w[oo++] = 1; }
                        n_base[nr] = (uint8_t)nb[u];
                        n_sib[nr] = (uint8_t)(sbv[u] + gr[u]);
                        if (tr_node) n_creator[nr] = cre[u];
                    } else {
                        const int t = -sr - 1;
                        n_in_row[io] = xa[u];
                        n_out_row[oo] = ya[u];
                        n_out_w[oo] = 1;
                        n_base[nr] = (uint8_t)nb[u];
                        int sb = 0;
                        if (so[u] >= 0) { const int o = sbv[u]; sb = ((o >> 4) + (o & 15) + 1) << 4; }
                        n_sib[nr] = (uint8_t)sb;
                        if (tr_node) n_creator[nr] = creator0 + t;
                    }
                }
                carry_in += tm.shfl(iin, T - 1);
                carry_out += tm.shfl(iout, T - 1);
            }
        }
        if (lane == 0 && N2 > 0) { n_in_off[N2] = carry_in; n_out_off[N2] = carry_out; }
    }
    tm.sync();
    if (on && err == ST_OK) { par ^= 1; N_io = N2; E_io = E2; }
    return err;
}

/* ------------------------------------------------------------------------------------------ */
/* consensus                                                                                   */
/* ------------------------------------------------------------------------------------------ */

/*
 * Heaviest bundling (semantics of abPOA's abpoa_heaviest_bundling, one consensus): reverse sweep over the row
 * order -- per row the out-edge of maximal weight, ties to the LATER edge whose target scores at least as much
 * (the source row: ties to the earlier edge unless the later target scores strictly more) -- then the path
 * source -> sink.  By the whole team:
 *   sweep: T rows per step from the sink side.  Every lane loads its row's first two edges and the scores of
 *          the targets that lie above the step's rows (final since an earlier step) in one round of independent
 *          loads; the rows are then resolved one after the other (row hi first), a row taking the scores of
 *          targets inside the step from the team's shared memory.  Rows with more than two out-edges read the
 *          rest of their edge list when it is their turn.
 *   path:  lane l looks at row curr + l; the longest run of rows whose best edge leads to the very next row
 *          is emitted at once (in row space the consensus path mostly runs through consecutive rows).
 * The serial version (lane 0, three dependent loads per row) took 5 % of the kernel time.
 * csc: T ints of shared memory.  Returns the length or -1 (capacity).  `on`: this team has a group to finish.
 */
template <int T>
__device__ __forceinline__ int heaviest_bundle(const KernelArgs &A, const Slot &S, const Team<T> &tm, int N_in, uint8_t *cons, int cap,
                                               int *csc, bool on) {
    const uint32_t *out_off = out_off_p(A, S), *out_row = out_row_p(A, S);
    const int32_t *out_w = out_w_p(A, S);
    const uint8_t *base = base_p(A, S);
    int32_t *score = cnt_p(A, S), *maxout = addin_p(A, S);
    const int lane = tm.tl;
    const int N = on ? N_in : 2;
    if (on && lane == 0) { score[N - 1] = 0; maxout[N - 1] = -1; }
    tm.sync();
    const int nstep = on ? (N - 1 + T - 1) / T : 0;       // rows N-2 .. 0
    const int wn = tm.wmax(nstep);
#pragma unroll 1
    for (int c = 0; c < wn; ++c) {
        const int hi = N - 2 - c * T;
        const int r = hi - lane;
        const bool mine = on && c < nstep && r >= 0;
        uint32_t o0 = 0;
        int ne = 0, t0 = -1, t1 = -1, w0 = 0, w1 = 0, s0 = 0, s1 = 0;
        if (mine) {
            o0 = out_off[r];
            ne = (int)(out_off[r + 1] - o0);
            if (ne > 0) { t0 = (int)out_row[o0]; w0 = out_w[o0]; }
            if (ne > 1) { t1 = (int)out_row[o0 + 1]; w1 = out_w[o0 + 1]; }
            if (ne > 0 && t0 > hi) s0 = score[t0];
            if (ne > 1 && t1 > hi) s1 = score[t1];
        }
        int my_score = 0, my_max = -1;
#pragma unroll 1
        for (int s = 0; s < T; ++s) {
            if (mine && lane == s) {                  // one row at a time; no cross-lane operation in here
                int best_w = r == 0 ? -1 : INT_MIN, best_sc = r == 0 ? -1 : 0;
#pragma unroll 1
                for (int e = 0; e < ne; ++e) {
                    int t, w, sc;
                    if (e == 0) { t = t0; w = w0; sc = s0; }
                    else if (e == 1) { t = t1; w = w1; sc = s1; }
                    else { t = (int)out_row[o0 + e]; w = out_w[o0 + e]; sc = t > hi ? score[t] : 0; }
                    if (t <= hi) sc = csc[hi - t];    // a row of this step: resolved in an earlier turn
                    const bool take = r == 0 ? (w > best_w || (w == best_w && sc > best_sc))
                                             : (best_w < w || (best_w == w && best_sc <= sc));
                    if (take) { best_w = w; best_sc = sc; my_max = t; }
                }
                my_score = ne > 0 ? best_w + best_sc : 0;
                csc[s] = my_score;
            }
            tm.sync();
        }
        if (mine) {
            if (r != 0) score[r] = my_score;
            maxout[r] = my_max;
        }
        tm.sync();
    }
    /* the path */
    int len = 0, curr = on ? maxout[0] : -1;
    bool over = false;
#pragma unroll 1
    for (;;) {
        const bool alive = on && !over && curr != N - 1 && curr >= 0;
        if (!tm.wany(alive)) break;
        const int row = curr + lane;
        const bool valid = alive && row <= N - 2;
        int m = -1, b = 4;
        if (valid) { m = maxout[row]; b = base[row]; }
        const bool ok = valid && m == row + 1 && row + 1 <= N - 2;
        const unsigned nb = ~tm.ballot(ok);
        const int f = nb ? __ffs(nb) - 1 : T;         // leading lanes whose row hands over to the next row
        const int k = min(f + 1, T);                  // rows curr .. curr + k - 1 are on the path
        const int next = tm.shfl(m, min(f, T - 1));
        if (alive) {
            if (len + k > cap) over = true;
            else {
                if (lane < k) cons[len + lane] = "ACGTN"[b];
                len += k;
                curr = f < T ? next : curr + T;
            }
        }
    }
    return over ? -1 : len;
}

}  // namespace mpoa
