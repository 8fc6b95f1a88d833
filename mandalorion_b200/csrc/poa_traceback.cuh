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
 * neither a diagonal nor a deletion) by all lanes of the warp.
 * The whole warp executes the traceback uniformly; lane 0 writes qmap.
 */
#pragma once
#include "poa_dp.cuh"

namespace mpoa {

constexpr int TBW = 128;  // rows held by the traceback window (descriptors + chain jump tables); a batch needs 64 of them

enum TbOps { OP_M = 1, OP_E1 = 2, OP_E2 = 4, OP_E = 6, OP_F1 = 8, OP_F2 = 16, OP_F = 24, OP_ALL = 31 };

template <typename T>
struct RowView {
    const T *h;       // H at h[j - beg], Eout1 at h[stride + ...], Eout2 at h[2*stride + ...]
    int beg, end, hi, stride, beg_sn, end_sn;
};

template <typename T>
__device__ __forceinline__ RowView<T> row_view(const KernelArgs &A, const Slot &S, int r, int lg, int qlen) {
    RowView<T> v;
    const int4 info = rowinfo_p(A, S)[r];
    const uint4 rt = rowtb_p(A, S)[r];
    v.beg_sn = info.x; v.end_sn = info.y;
    v.beg = info.x << lg;
    v.end = ((info.y + 1) << lg) - 1;
    v.hi = min(v.end, qlen);
    v.stride = (int)rt.y;
    v.h = reinterpret_cast<const T *>(reinterpret_cast<const uint32_t *>(tb_p(A, S)) + rt.x);
    return v;
}

/* row_view through the shared-memory window of the last speculative batch (slot x = row wb - x) */
template <typename T>
__device__ __forceinline__ RowView<T> row_view_w(const KernelArgs &A, const Slot &S, int r, int lg, int qlen, int wb,
                                                 const int4 *winfo, const uint4 *wtb, int tbw) {
    const int x = wb - r;
    if ((unsigned)x >= (unsigned)tbw) return row_view<T>(A, S, r, lg, qlen);
    RowView<T> v;
    const int4 info = winfo[x];
    const uint4 rt = wtb[x];
    v.beg_sn = info.x; v.end_sn = info.y;
    v.beg = info.x << lg;
    v.end = ((info.y + 1) << lg) - 1;
    v.hi = min(v.end, qlen);
    v.stride = (int)rt.y;
    v.h = reinterpret_cast<const T *>(reinterpret_cast<const uint32_t *>(tb_p(A, S)) + rt.x);
    return v;
}

template <typename T>
__device__ __forceinline__ int rv_get(const RowView<T> &v, int arr, int j) {
    return (j >= v.beg && j <= v.hi) ? (int)v.h[arr * v.stride + (j - v.beg)] : NEG;
}

__device__ __forceinline__ int score_of(const DevParams &P, int nbase, int qb) {
    return (nbase >= 4 || qb >= 4) ? 0 : (nbase == qb ? P.match : -P.mismatch);
}

/* Hhat[i][k] = max(M + s, Ein1, Ein2): the pre-insertion score the DP used for cell (i,k) */
template <typename T>
__device__ __forceinline__ int hhat_cell(const KernelArgs &A, const Slot &S, const RowView<T> &vi, int i, int k, int nbase,
                                         const uint8_t *__restrict__ q, int qlen, int lg) {
    const uint32_t *in_off = in_off_p(A, S), *in_row = in_row_p(A, S);
    const int in0 = (int)in_off[i], npre = (int)in_off[i + 1] - in0;
    int mx = NEG, e1 = NEG, e2 = NEG;
#pragma unroll 1
    for (int kk = 0; kk < npre; ++kk) {
        const int p = (int)in_row[in0 + kk];
        const RowView<T> vp = row_view<T>(A, S, p, lg, qlen);
        const int lo = max(vi.beg_sn, vp.beg_sn) << lg;
        const int hi = min(((min(vi.end_sn, vp.end_sn) + 1) << lg) - 1, qlen);
        if (k >= lo && k <= hi) {
            if (k > lo) mx = max(mx, (int)vp.h[k - 1 - vp.beg]);
            e1 = max(e1, (int)vp.h[vp.stride + k - vp.beg]);
            e2 = max(e2, (int)vp.h[2 * vp.stride + k - vp.beg]);
        }
    }
    const int s = k > 0 ? score_of(A.P, nbase, q[k - 1]) : 0;
    return max(mx + s, max(e1, e2));
}

/* F1/F2 of row i at columns j and j-1 (what the DP had before taking H = max(Hhat, F1, F2)) */
template <typename T>
__device__ __forceinline__ void f_values(const KernelArgs &A, const Slot &S, const RowView<T> &vi, int i, int j, int nbase,
                                         const uint8_t *__restrict__ q, int qlen, int lg, int lane, int f[4]) {
    const DevParams &P = A.P;
    int f1j = NEG, f2j = NEG, f1m = NEG, f2m = NEG;
#pragma unroll 1
    for (int k = vi.beg + lane; k <= j - 1; k += 32) {
        const int hk = hhat_cell<T>(A, S, vi, i, k, nbase, q, qlen, lg);
        f1j = max(f1j, hk - P.oe1 - P.e1 * (j - 1 - k));
        f2j = max(f2j, hk - P.oe2 - P.e2 * (j - 1 - k));
        if (k <= j - 2) {
            f1m = max(f1m, hk - P.oe1 - P.e1 * (j - 2 - k));
            f2m = max(f2m, hk - P.oe2 - P.e2 * (j - 2 - k));
        }
    }
    f[0] = __reduce_max_sync(FULL, f1j);
    f[1] = __reduce_max_sync(FULL, f2j);
    f[2] = __reduce_max_sync(FULL, f1m);
    f[3] = __reduce_max_sync(FULL, f2m);
}

/* Writes qmap[t] = row the query base t is aligned to, -1 for an inserted base.  Returns false
 * when no move is possible (abPOA dies in cg_backtrack; the reference then uses the first read). */
template <typename T>
__device__ __forceinline__ bool traceback(const KernelArgs &A, const Slot &S, const uint8_t *__restrict__ q, int qlen,
                                          const AlnState &R, int lane, int *scratch) {
    const DevParams &P = A.P;
    const uint32_t *in_off = in_off_p(A, S), *in_row = in_row_p(A, S);
    int32_t *qmap = qmap_p(A, S);
    const int lg = R.lgpn;
    int i = R.best_i, j = R.best_j, cur_op = OP_ALL;
#pragma unroll 1
    for (int t = j + lane; t < qlen; t += 32) qmap[t] = -1;
    int4 *winfo = reinterpret_cast<int4 *>(scratch);          // [TBW] rowinfo of rows i, i-1, ...
    uint4 *wtb = reinterpret_cast<uint4 *>(scratch) + TBW;    // [TBW] rowtb
    int *jump = scratch + TBW * 8;                            // [5][TBW + 1] chain jump tables
    int wb = INT_MIN / 2;        // row held by window slot 0 (none yet)
    bool retry_batch = true;     // false right after a batch that stopped early: its next step is known to fail
    while (i > 0 && j > 0) {
        if (cur_op == OP_ALL && retry_batch) {
            /*
             * Speculative batch: in state ALL the next step is "diagonal to the FIRST predecessor"
             * whenever H[p0][j-1] + s == H[i][j].  The chain i -> p0(i) -> p0(p0(i)) ... is followed
             * for up to 31 steps, lane l checks step l against the stored rows, and the longest
             * prefix of successful checks is taken at once.  Every step taken is exactly the step
             * the serial logic below would have taken; a batch of length 0 falls through to it.
             */
            if ((unsigned)(wb - i) > (unsigned)(TBW - 64)) {
                /* (re)load the window at row i: descriptors of rows i .. i-TBW+1 and, for every slot,
                 * the slot of the row's first predecessor; then the jump tables by pointer doubling
                 * (jump[k][x] = slot 2^k chain steps after x).  Batches starting in the upper half of
                 * the window reuse all of it. */
                __syncwarp();
                for (int x = lane; x < TBW; x += 32) {
                    const int row = i - x;
                    int nx = TBW;             // TBW: outside the window / no predecessor
                    if (row >= 0) {
                        const uint4 rt = rowtb_p(A, S)[row];
                        winfo[x] = rowinfo_p(A, S)[row]; wtb[x] = rt;
                        if (row > 0) nx = min(TBW, i - (int)rt.z);
                    }
                    jump[x] = nx;
                }
                if (lane < 5) jump[lane * (TBW + 1) + TBW] = TBW;
                wb = i;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int *jk = jump + k * (TBW + 1);
                    for (int x = lane; x < TBW; x += 32) jump[(k + 1) * (TBW + 1) + x] = jk[jk[x]];
                    __syncwarp();
                }
            }
            /* slot reached after l chain steps from row i: lane l composes the jumps of its bits */
            int myx = wb - i;
#pragma unroll
            for (int k = 0; k < 5; ++k)
                if ((lane >> k) & 1) myx = jump[k * (TBW + 1) + myx];
            if (myx >= TBW) myx = -1;
            const int col = j - lane;
            bool inband = false;
            int hval = NEG, row = -1, s = 0;
            if (myx >= 0 && col >= 0) {
                row = wb - myx;
                const int4 info = winfo[myx];
                const uint4 rt = wtb[myx];
                const int beg = info.x << lg, hi = min(((info.y + 1) << lg) - 1, qlen);
                inband = col >= beg && col <= hi;
                if (inband) hval = (int)reinterpret_cast<const T *>(reinterpret_cast<const uint32_t *>(tb_p(A, S)) + rt.x)[col - beg];
                if (col >= 1) s = score_of(P, (int)rt.w, q[col - 1]);
            }
            const int hnext = __shfl_down_sync(FULL, hval, 1);
            const bool nok = __shfl_down_sync(FULL, inband ? 1 : 0, 1) != 0;
            const bool ok = lane < 31 && inband && row > 0 && col >= 1 && nok && (hnext + s == hval);
            const int cnt = __ffs(~__ballot_sync(FULL, ok)) - 1;
            retry_batch = cnt >= 31;
            if (cnt > 0) {
                if (lane < cnt) qmap[col - 1] = row;
                i = __shfl_sync(FULL, row, cnt);
                j -= cnt;
                continue;
            }
        }
        retry_batch = true;
        const RowView<T> vi = row_view_w<T>(A, S, i, lg, qlen, wb, winfo, wtb, TBW);
        const int in0 = (int)in_off[i], npre = (int)in_off[i + 1] - in0;
        const int nbase = (int)(meta_p(A, S)[i] & META_BASE);
        const int s = score_of(P, nbase, q[j - 1]);
        const int hij = rv_get(vi, 0, j);
        bool hit = false;
        /* The predecessors are probed by one lane each (32 at a time); the FIRST one in edge order
         * that qualifies is taken, exactly like abPOA's loop over pre_id. */
        if (cur_op & OP_M) {
#pragma unroll 1
            for (int k0 = 0; k0 < npre && !hit; k0 += 32) {
                const int k = k0 + lane;
                int p = 0;
                bool m = false;
                if (k < npre) {
                    p = (int)in_row[in0 + k];
                    const RowView<T> vp = row_view_w<T>(A, S, p, lg, qlen, wb, winfo, wtb, TBW);
                    if (j - 1 >= vp.beg && j - 1 <= vp.end) m = rv_get(vp, 0, j - 1) + s == hij;
                }
                const unsigned b = __ballot_sync(FULL, m);
                if (b) {
                    if (lane == 0) qmap[j - 1] = i;
                    i = __shfl_sync(FULL, p, __ffs(b) - 1);
                    --j; hit = true; cur_op = OP_ALL;
                }
            }
        }
        if (!hit && (cur_op & OP_E)) {
            const int e1ij = rv_get(vi, 1, j), e2ij = rv_get(vi, 2, j);
#pragma unroll 1
            for (int k0 = 0; k0 < npre && !hit; k0 += 32) {
                const int k = k0 + lane;
                int p = 0, nop = 0;
                if (k < npre) {
                    p = (int)in_row[in0 + k];
                    const RowView<T> vp = row_view_w<T>(A, S, p, lg, qlen, wb, winfo, wtb, TBW);
                    if (j >= vp.beg && j <= vp.end) {
                        const int hp = rv_get(vp, 0, j);
                        if (cur_op & OP_E1) {
                            const int pe1 = rv_get(vp, 1, j);
                            const bool take = (cur_op & OP_M) ? (hij == pe1) : (e1ij == pe1 - P.e1);
                            if (take) nop = (hp - P.oe1 == pe1) ? (OP_M | OP_F) : OP_E1;
                        }
                        if (nop == 0 && (cur_op & OP_E2)) {
                            const int pe2 = rv_get(vp, 2, j);
                            const bool take = (cur_op & OP_M) ? (hij == pe2) : (e2ij == pe2 - P.e2);
                            if (take) nop = (hp - P.oe2 == pe2) ? (OP_M | OP_F) : OP_E2;
                        }
                    }
                }
                const unsigned b = __ballot_sync(FULL, nop != 0);
                if (b) {
                    const int sl = __ffs(b) - 1;
                    i = __shfl_sync(FULL, p, sl);
                    cur_op = __shfl_sync(FULL, nop, sl);
                    hit = true;
                }
            }
        }
        if (!hit && (cur_op & OP_F)) {
            int f[4];
            f_values<T>(A, S, vi, i, j, nbase, q, qlen, lg, lane, f);
            const int hjm = rv_get(vi, 0, j - 1);
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
        if (!hit) return false;
    }
#pragma unroll 1
    for (int t = lane; t < j; t += 32) qmap[t] = -1;
    __syncwarp();
    return true;
}

}  // namespace mpoa
