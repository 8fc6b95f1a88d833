/*
 * poa_traceback.cuh -- traceback over the stored band rows, by VALUE comparison, in the order and
 * with the open/extend state machine of abPOA's cg_backtrack (the
 * reference obtains it from `abpoa -M 5 -r 0`, utils/SpliceDefineConsensus.py:917):
 *   at (i,j):  1. diagonal  : first predecessor p (edge order) with j-1 inside p's band and
 *                             H[p][j-1] + s == H[i][j]
 *              2. deletion  : first p with j inside p's band and H[i][j] == Eout1[p][j] (then
 *                             Eout2), open vs extend decided by H[p][j] - oe == Eout[p][j]
 *              3. insertion : H[i][j] == F1[i][j] (then F2), open vs extend likewise
 * F is not stored: it is recomputed for the one row that needs it (rare: only when a step is
 * neither a diagonal nor a deletion) by all lanes of the team.
 * A team executes the traceback uniformly; its lane 0 writes qmap.
 */
#pragma once
#include "poa_dp.cuh"

namespace mpoa {

enum TbOps { OP_M = 1, OP_E1 = 2, OP_E2 = 4, OP_E = 6, OP_F1 = 8, OP_F2 = 16, OP_F = 24, OP_ALL = 31 };

/* rows held by the traceback window (descriptors + chain jump tables); a batch needs 2T of them */
template <int T> __host__ __device__ constexpr int tb_window() { return 4 * T; }
template <int T> __host__ __device__ constexpr int tb_scratch_words() { return tb_window<T>() * 8 + 5 * (tb_window<T>() + 1); }

/*
 * Stored band row.  TV = int32_t (dp_align32, absolute scores) or int16_t (dp_band16, scores
 * relative to match * first column of the lane that computed them; CPL = cells per lane of that
 * kernel).  get() always returns ABSOLUTE scores, so the logic below is abPOA's, whatever the
 * storage.
 */
template <typename TV, int CPL>
struct RowView {
    const TV *h;      // H at h[j - beg], Eout1 at h[stride + ...], Eout2 at h[2*stride + ...]
    int beg, end, hi, stride, beg_sn, end_sn, match;
    __device__ __forceinline__ int frame(int j) const {
        if constexpr (CPL == 0) return 0;
        else return match * (j / CPL * CPL);
    }
    /* cell j must be inside [beg, hi] */
    __device__ __forceinline__ int at(int arr, int j) const { return (int)h[arr * stride + (j - beg)] + frame(j); }
    __device__ __forceinline__ int get(int arr, int j) const { return (j >= beg && j <= hi) ? at(arr, j) : NEG; }
};

template <typename TV, int CPL>
__device__ __forceinline__ RowView<TV, CPL> make_view(const KernelArgs &A, const Slot &S, const int4 info, const uint4 rt, int lg, int qlen) {
    RowView<TV, CPL> v;
    v.beg_sn = info.x; v.end_sn = info.y;
    v.beg = info.x << lg;
    v.end = ((info.y + 1) << lg) - 1;
    v.hi = min(v.end, qlen);
    v.stride = (int)rt.y;
    v.match = A.P.match;
    v.h = reinterpret_cast<const TV *>(reinterpret_cast<const uint32_t *>(tb_p(A, S)) + rt.x);
    return v;
}

template <typename TV, int CPL>
__device__ __forceinline__ RowView<TV, CPL> row_view(const KernelArgs &A, const Slot &S, int r, int lg, int qlen) {
    return make_view<TV, CPL>(A, S, rowinfo_p(A, S)[r], rowtb_p(A, S)[r], lg, qlen);
}

/* row_view through the shared-memory window of the last speculative batch (slot x = row wb - x) */
template <typename TV, int CPL>
__device__ __forceinline__ RowView<TV, CPL> row_view_w(const KernelArgs &A, const Slot &S, int r, int lg, int qlen, int wb,
                                                       const int4 *winfo, const uint4 *wtb, int tbw) {
    const int x = wb - r;
    if ((unsigned)x >= (unsigned)tbw) return row_view<TV, CPL>(A, S, r, lg, qlen);
    return make_view<TV, CPL>(A, S, winfo[x], wtb[x], lg, qlen);
}

__device__ __forceinline__ int score_of(const DevParams &P, int nbase, int qb) {
    return (nbase >= 4 || qb >= 4) ? 0 : (nbase == qb ? P.match : -P.mismatch);
}

/* Hhat[i][k] = max(M + s, Ein1, Ein2): the pre-insertion score the DP used for cell (i,k) */
template <typename TV, int CPL>
__device__ __forceinline__ int hhat_cell(const KernelArgs &A, const Slot &S, const RowView<TV, CPL> &vi, int i, int k, int nbase,
                                         const uint8_t *__restrict__ q, int qlen, int lg) {
    const uint32_t *in_off = in_off_p(A, S), *in_row = in_row_p(A, S);
    const int in0 = (int)in_off[i], npre = (int)in_off[i + 1] - in0;
    int mx = NEG, e1 = NEG, e2 = NEG;
#pragma unroll 1
    for (int kk = 0; kk < npre; ++kk) {
        const int p = (int)in_row[in0 + kk];
        const RowView<TV, CPL> vp = row_view<TV, CPL>(A, S, p, lg, qlen);
        const int lo = max(vi.beg_sn, vp.beg_sn) << lg;
        const int hi = min(((min(vi.end_sn, vp.end_sn) + 1) << lg) - 1, qlen);
        if (k >= lo && k <= hi) {
            if (k > lo) mx = max(mx, vp.at(0, k - 1));
            e1 = max(e1, vp.at(1, k));
            e2 = max(e2, vp.at(2, k));
        }
    }
    const int s = k > 0 ? score_of(A.P, nbase, q[k - 1]) : 0;
    return max(mx + s, max(e1, e2));
}

/* F1/F2 of row i at columns j and j-1 (what the DP had before taking H = max(Hhat, F1, F2)).
 * Lockstep: the loop runs for the larger trip count of the warp's teams. */
template <typename TV, int CPL, int T>
__device__ __forceinline__ void f_values(const KernelArgs &A, const Slot &S, const Team<T> &tm, const RowView<TV, CPL> &vi, int i, int j,
                                         int nbase, const uint8_t *__restrict__ q, int qlen, int lg, int f[4], bool on) {
    const DevParams &P = A.P;
    int f1j = NEG, f2j = NEG, f1m = NEG, f2m = NEG;
    const int steps = tm.wmax(on ? (j - vi.beg + T - 1) / T : 0);
#pragma unroll 1
    for (int it = 0; it < steps; ++it) {
        const int k = vi.beg + tm.tl + it * T;
        if (on && k <= j - 1) {
            const int hk = hhat_cell<TV, CPL>(A, S, vi, i, k, nbase, q, qlen, lg);
            f1j = max(f1j, hk - P.oe1 - P.e1 * (j - 1 - k));
            f2j = max(f2j, hk - P.oe2 - P.e2 * (j - 1 - k));
            if (k <= j - 2) {
                f1m = max(f1m, hk - P.oe1 - P.e1 * (j - 2 - k));
                f2m = max(f2m, hk - P.oe2 - P.e2 * (j - 2 - k));
            }
        }
    }
    f[0] = tm.rmax(f1j);
    f[1] = tm.rmax(f2j);
    f[2] = tm.rmax(f1m);
    f[3] = tm.rmax(f2m);
}

/* Writes qmap[t] = row the query base t is aligned to, -1 for an inserted base.  Returns false
 * when no move is possible (abPOA dies in cg_backtrack; the reference then uses the first read).
 * Lockstep: the teams of a warp walk their own paths, but every iteration of the loop is taken by
 * both until both are done; a team takes part in the speculative batch and/or in the serial step
 * of an iteration by predicate. */
template <typename TV, int CPL, int T>
__device__ __forceinline__ bool traceback(const KernelArgs &A, const Slot &S, const Team<T> &tm, const uint8_t *__restrict__ q, int qlen,
                                          const AlnState &R, int *scratch, bool on) {
    constexpr int TBW = tb_window<T>();
    constexpr int LOGT = T == 32 ? 5 : 4;
    const int lane = tm.tl;
    const DevParams &P = A.P;
    const uint32_t *in_off = in_off_p(A, S), *in_row = in_row_p(A, S);
    int32_t *qmap = qmap_p(A, S);
    const int lg = R.lgpn;
    int i = on ? R.best_i : 0, j = on ? R.best_j : 0, cur_op = OP_ALL;
    if (on) {
#pragma unroll 1
        for (int t = j + lane; t < qlen; t += T) qmap[t] = -1;
    }
    int4 *winfo = reinterpret_cast<int4 *>(scratch);          // [TBW] rowinfo of rows i, i-1, ...
    uint4 *wtb = reinterpret_cast<uint4 *>(scratch) + TBW;    // [TBW] rowtb
    int *jump = scratch + TBW * 8;                            // [LOGT][TBW + 1] chain jump tables
    int wb = INT_MIN / 2;        // row held by window slot 0 (none yet)
    bool retry_batch = true;     // false right after a batch that stopped early: its next step is known to fail
    bool fail = false;
    for (;;) {
        const bool alive = on && !fail && i > 0 && j > 0;
        if (!tm.wany(alive)) break;
        bool took = false;
        const bool want_batch = alive && cur_op == OP_ALL && retry_batch;
        if (tm.wany(want_batch)) {
            /*
             * Speculative batch: in state ALL the next step is "diagonal to the FIRST predecessor"
             * whenever H[p0][j-1] + s == H[i][j].  The chain i -> p0(i) -> p0(p0(i)) ... is followed
             * for up to T-1 steps, lane l checks step l against the stored rows, and the longest
             * prefix of successful checks is taken at once.  Every step taken is exactly the step
             * the serial logic below would have taken; a batch of length 0 falls through to it.
             */
            const bool reload = want_batch && (unsigned)(wb - i) > (unsigned)(TBW - 2 * T);
            if (tm.wany(reload)) {
                /* (re)load the window at row i: descriptors of rows i .. i-TBW+1 and, for every slot,
                 * the slot of the row's first predecessor; then the jump tables by pointer doubling
                 * (jump[k][x] = slot 2^k chain steps after x).  Batches starting in the upper half of
                 * the window reuse all of it. */
                tm.sync();
                if (reload) {
                    for (int x = lane; x < TBW; x += T) {
                        const int row = i - x;
                        int nx = TBW;             // TBW: outside the window / no predecessor
                        if (row >= 0) {
                            const uint4 rt = rowtb_p(A, S)[row];
                            winfo[x] = rowinfo_p(A, S)[row]; wtb[x] = rt;
                            if (row > 0) nx = min(TBW, i - (int)rt.z);
                        }
                        jump[x] = nx;
                    }
                    if (lane < LOGT) jump[lane * (TBW + 1) + TBW] = TBW;
                    wb = i;
                }
                tm.sync();
#pragma unroll
                for (int k = 0; k < LOGT - 1; ++k) {
                    if (reload) {
                        const int *jk = jump + k * (TBW + 1);
                        for (int x = lane; x < TBW; x += T) jump[(k + 1) * (TBW + 1) + x] = jk[jk[x]];
                    }
                    tm.sync();
                }
            }
            /* slot reached after l chain steps from row i: lane l composes the jumps of its bits */
            int myx = -1;
            if (want_batch) {
                myx = wb - i;
#pragma unroll
                for (int k = 0; k < LOGT; ++k)
                    if ((lane >> k) & 1) myx = jump[k * (TBW + 1) + myx];
                if (myx >= TBW) myx = -1;
            }
            const int col = j - lane;
            bool inband = false;
            int hval = NEG, row = -1, s = 0;
            if (myx >= 0 && col >= 0) {
                row = wb - myx;
                const RowView<TV, CPL> v = make_view<TV, CPL>(A, S, winfo[myx], wtb[myx], lg, qlen);
                inband = col >= v.beg && col <= v.hi;
                if (inband) hval = v.at(0, col);
                if (col >= 1) s = score_of(P, (int)wtb[myx].w, q[col - 1]);
            }
            const int hnext = tm.shfl_down(hval, 1);
            const bool nok = tm.shfl_down(inband ? 1 : 0, 1) != 0;
            const bool ok = lane < T - 1 && inband && row > 0 && col >= 1 && nok && (hnext + s == hval);
            const int cnt = __ffs(~tm.ballot(ok)) - 1;
            const int ni = tm.shfl(row, cnt & (T - 1));
            if (want_batch) {
                retry_batch = cnt >= T - 1;
                if (cnt > 0) {
                    if (lane < cnt) qmap[col - 1] = row;
                    i = ni;
                    j -= cnt;
                    took = true;
                }
            }
        }
        /* serial step of the teams that did not move in a batch */
        const bool ser = alive && !took;
        if (tm.wany(ser)) {
            if (ser) retry_batch = true;
            RowView<TV, CPL> vi;
            int in0 = 0, npre = 0, nbase = 0, s = 0, hij = NEG;
            if (ser) {
                vi = row_view_w<TV, CPL>(A, S, i, lg, qlen, wb, winfo, wtb, TBW);
                in0 = (int)in_off[i]; npre = (int)in_off[i + 1] - in0;
                nbase = (int)(meta_p(A, S)[i] & META_BASE);
                s = score_of(P, nbase, q[j - 1]);
                hij = vi.get(0, j);
            }
            bool hit = false;
            /* The predecessors are probed by one lane each (T at a time); the FIRST one in edge order
             * that qualifies is taken, exactly like abPOA's loop over pre_id. */
            {
                const bool go = ser && (cur_op & OP_M);
                const int kmax = tm.wmax(go ? npre : 0);
#pragma unroll 1
                for (int k0 = 0; k0 < kmax; k0 += T) {
                    const int k = k0 + lane;
                    int p = 0;
                    bool m = false;
                    if (go && !hit && k < npre) {
                        p = (int)in_row[in0 + k];
                        const RowView<TV, CPL> vp = row_view_w<TV, CPL>(A, S, p, lg, qlen, wb, winfo, wtb, TBW);
                        if (j - 1 >= vp.beg && j - 1 <= vp.end) m = vp.get(0, j - 1) + s == hij;
                    }
                    const unsigned b = tm.ballot(m);
                    const int np = tm.shfl(p, (__ffs(b) - 1) & (T - 1));
                    if (b) {
                        if (lane == 0) qmap[j - 1] = i;
                        i = np;
                        --j; hit = true; cur_op = OP_ALL;
                    }
                }
            }
            {
                const bool go = ser && !hit && (cur_op & OP_E);
                int e1ij = NEG, e2ij = NEG;
                if (go) { e1ij = vi.get(1, j); e2ij = vi.get(2, j); }
                const int kmax = tm.wmax(go ? npre : 0);
                bool ehit = false;
#pragma unroll 1
                for (int k0 = 0; k0 < kmax; k0 += T) {
                    const int k = k0 + lane;
                    int p = 0, nop = 0;
                    if (go && !ehit && k < npre) {
                        p = (int)in_row[in0 + k];
                        const RowView<TV, CPL> vp = row_view_w<TV, CPL>(A, S, p, lg, qlen, wb, winfo, wtb, TBW);
                        if (j >= vp.beg && j <= vp.end) {
                            const int hp = vp.get(0, j);
                            if (cur_op & OP_E1) {
                                const int pe1 = vp.get(1, j);
                                const bool take = (cur_op & OP_M) ? (hij == pe1) : (e1ij == pe1 - P.e1);
                                if (take) nop = (hp - P.oe1 == pe1) ? (OP_M | OP_F) : OP_E1;
                            }
                            if (nop == 0 && (cur_op & OP_E2)) {
                                const int pe2 = vp.get(2, j);
                                const bool take = (cur_op & OP_M) ? (hij == pe2) : (e2ij == pe2 - P.e2);
                                if (take) nop = (hp - P.oe2 == pe2) ? (OP_M | OP_F) : OP_E2;
                            }
                        }
                    }
                    const unsigned b = tm.ballot(nop != 0);
                    const int sl = (__ffs(b) - 1) & (T - 1);
                    const int np = tm.shfl(p, sl), nn = tm.shfl(nop, sl);
                    if (b) { i = np; cur_op = nn; ehit = true; }
                }
                hit = hit || ehit;
            }
            {
                const bool go = ser && !hit && (cur_op & OP_F);
                if (tm.wany(go)) {
                    int f[4];
                    f_values<TV, CPL, T>(A, S, tm, vi, i, j, nbase, q, qlen, lg, f, go);
                    if (go) {
                        const int hjm = vi.get(0, j - 1);
                        if (cur_op & OP_F1) {
                            if (!(cur_op & OP_M) || hij == f[0]) {
                                if (hjm - P.oe1 == f[0]) { cur_op = OP_M | OP_E; hit = true; }
                                else if (f[2] - P.e1 == f[0]) { cur_op = OP_F1; hit = true; }
                            }
                        }
                        if (!hit && (cur_op & OP_F2)) {
                            if (!(cur_op & OP_M) || hij == f[1]) {
                                if (hjm - P.oe2 == f[1]) { cur_op = OP_M | OP_E; hit = true; }
                                else if (f[3] - P.e2 == f[1]) { cur_op = OP_F2; hit = true; }
                            }
                        }
                        if (hit) {
                            if (lane == 0) qmap[j - 1] = -1;
                            --j;
                        }
                    }
                }
            }
            if (ser && !hit) fail = true;
        }
    }
    if (on && !fail) {
#pragma unroll 1
        for (int t = lane; t < j; t += T) qmap[t] = -1;
    }
    tm.sync();
    return !fail;
}

}  // namespace mpoa
