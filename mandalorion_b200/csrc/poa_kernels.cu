/*
 * poa_kernels.cu -- sm_100a kernels of the consensus hot path.
 *
 * Replaces, for a whole batch of isoform read groups, what the reference obtains from one
 * `abpoa -M 5 -r 0 in.fasta` process per isoform (reference utils/SpliceDefineConsensus.py:917,
 * loop at defineIsoforms.py:87-91).  Algorithmic contract (scores, band, tie-breaks): DESIGN.md
 * section "Algorithm"; data layout: poa_device.cuh.
 *
 * Kernels:
 *   encode_bases_kernel  ASCII -> nt4 codes (HBM bound, 1 B in / 1 B out per base)
 *   poa_group_kernel     persistent, one warp per read group: graph build, adaptive-banded
 *                        convex-gap DP, flag traceback, graph merge, heaviest-bundle consensus
 */
#include <climits>
#include "poa_device.cuh"

namespace mpoa {

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                               */
/* ------------------------------------------------------------------------------------------ */

struct Slot {
    uint8_t *base[2], *sib[2];
    int32_t *creator[2];
    uint32_t *in_off[2], *in_row[2], *out_off[2], *out_row[2];
    int32_t *out_w[2];
    int32_t *remain;
    uint32_t *meta;
    int4 *rowinfo;
    uint32_t *tboff;
    int32_t *rowbest;
    uint32_t *spoff;
    int32_t *qmap;
    int32_t *pv, *pkey, *pnew, *psib, *nin, *nout;
    int32_t *cnt, *addin, *addout, *srcof;
    uint8_t *grow;
    uint8_t *tb;
    int32_t *spill;
};

__device__ __forceinline__ Slot make_slot(const KernelArgs &A, int slot) {
    uint8_t *b = A.ws + (uint64_t)slot * A.L.slot_bytes;
    Slot S;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        S.base[p] = b + A.L.base[p];
        S.sib[p] = b + A.L.sib[p];
        S.creator[p] = (int32_t *)(b + A.L.creator[p]);
        S.in_off[p] = (uint32_t *)(b + A.L.in_off[p]);
        S.in_row[p] = (uint32_t *)(b + A.L.in_row[p]);
        S.out_off[p] = (uint32_t *)(b + A.L.out_off[p]);
        S.out_row[p] = (uint32_t *)(b + A.L.out_row[p]);
        S.out_w[p] = (int32_t *)(b + A.L.out_w[p]);
    }
    S.remain = (int32_t *)(b + A.L.remain);
    S.meta = (uint32_t *)(b + A.L.meta);
    S.rowinfo = (int4 *)(b + A.L.rowinfo);
    S.tboff = (uint32_t *)(b + A.L.tboff);
    S.rowbest = (int32_t *)(b + A.L.rowbest);
    S.spoff = (uint32_t *)(b + A.L.spoff);
    S.qmap = (int32_t *)(b + A.L.qmap);
    S.pv = (int32_t *)(b + A.L.pv);
    S.pkey = (int32_t *)(b + A.L.pkey);
    S.pnew = (int32_t *)(b + A.L.pnew);
    S.psib = (int32_t *)(b + A.L.psib);
    S.nin = (int32_t *)(b + A.L.nin);
    S.nout = (int32_t *)(b + A.L.nout);
    S.cnt = (int32_t *)(b + A.L.cnt);
    S.addin = (int32_t *)(b + A.L.addin);
    S.addout = (int32_t *)(b + A.L.addout);
    S.srcof = (int32_t *)(b + A.L.srcof);
    S.grow = b + A.L.grow;
    S.tb = b + A.L.tb;
    S.spill = (int32_t *)(b + A.L.spill);
    return S;
}

__device__ __forceinline__ int warp_incl_sum(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ int warp_incl_max(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v = max(v, t);
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

enum MetaBits { META_BASE = 7, META_FAR = 8, META_TOSINK = 16 };

/* ------------------------------------------------------------------------------------------ */
/* encode                                                                                      */
/* ------------------------------------------------------------------------------------------ */

__global__ void encode_bases_kernel(const uint8_t *__restrict__ ascii, uint8_t *__restrict__ codes, int64_t n) {
    /* 16 bases per thread, 128-bit loads/stores */
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    for (; i < n; i += stride) {
        if (i + 16 <= n) {
            uint4 v = *reinterpret_cast<const uint4 *>(ascii + i);
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t o = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    uint32_t c = (w[k] >> (8 * b)) & 0xdfu;  // fold case
                    uint32_t code = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
                    o |= code << (8 * b);
                }
                w[k] = o;
            }
            *reinterpret_cast<uint4 *>(codes + i) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int64_t k = i; k < n; ++k) {
                uint32_t c = ascii[k] & 0xdfu;
                codes[k] = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* graph: first read, remain pass                                                              */
/* ------------------------------------------------------------------------------------------ */

/* first read = linear chain src -> b0 -> ... -> sink, every edge weight 1 */
__device__ void init_graph(const Slot &S, int par, const uint8_t *seq, int len, int creator0, int lane) {
    const int N = len + 2;
    for (int r = lane; r <= N; r += 32) {
        if (r < N) {
            const bool real = r >= 1 && r <= len;
            S.base[par][r] = real ? seq[r - 1] : 0;
            S.sib[par][r] = 0;
            S.creator[par][r] = real ? creator0 + r - 1 : -1;
            if (r >= 1) S.in_row[par][r - 1] = r - 1;
            if (r < N - 1) { S.out_row[par][r] = r + 1; S.out_w[par][r] = 1; }
        }
        S.in_off[par][r] = r == 0 ? 0 : r - 1;
        S.out_off[par][r] = r < N - 1 ? r : N - 1;
    }
    __syncwarp();
}

/*
 * remain[r] = remain[heaviest out-neighbour (first maximum)] + 1, remain[sink] = -1: what abPOA's
 * reverse BFS computes.  Rows are handled 32 at a time from the sink side; in-window chains are
 * resolved by pointer jumping.  Also emits the per-row meta word used by the DP.
 */
__device__ void remain_pass(const Slot &S, int par, int N, int lane) {
    const uint32_t *out_off = S.out_off[par], *out_row = S.out_row[par];
    const int32_t *out_w = S.out_w[par];
    if (lane == 0) {
        S.remain[N - 1] = -1;
        S.meta[N - 1] = 0;
    }
    int prev_vals = 0;  // remain of rows [w0+32, w0+64)
    for (int w0 = ((N - 2) / 32) * 32; w0 >= 0; w0 -= 32) {
        const int r = w0 + lane;
        const bool active = r <= N - 2;
        int hs = N - 1, maxrow = 0, tosink = 0;
        if (active) {
            const uint32_t o0 = out_off[r], o1 = out_off[r + 1];
            int maxw = -1;
            for (uint32_t e = o0; e < o1; ++e) {
                const int t = (int)out_row[e], w = out_w[e];
                if (w > maxw) { maxw = w; hs = t; }
                maxrow = max(maxrow, t);
                tosink |= (t == N - 1);
            }
        }
        int acc, nxt = -1;
        const int src_prev = min(31, max(0, hs - (w0 + 32)));
        const int from_prev = __shfl_sync(FULL, prev_vals, src_prev);
        if (!active) acc = 0;
        else if (hs == N - 1) acc = 0;
        else if (hs >= w0 + 64) acc = S.remain[hs] + 1;
        else if (hs >= w0 + 32) acc = from_prev + 1;
        else { acc = 1; nxt = hs - w0; }
#pragma unroll
        for (int it = 0; it < 5; ++it) {
            const int sl = max(nxt, 0);
            const int a = __shfl_sync(FULL, acc, sl);
            const int n = __shfl_sync(FULL, nxt, sl);
            if (nxt >= 0) { acc += a; nxt = n; }
        }
        if (active) {
            S.remain[r] = acc;
            S.meta[r] = (uint32_t)S.base[par][r] | ((maxrow - r >= RING) ? META_FAR : 0) | (tosink ? META_TOSINK : 0);
        }
        prev_vals = acc;
        __syncwarp();
    }
}

/* ------------------------------------------------------------------------------------------ */
/* DP                                                                                          */
/* ------------------------------------------------------------------------------------------ */

struct AlnState {
    int best_i, best_j, best_score, pn, bits;
    unsigned long long cells, intops, full, tbbytes;
};

/* traceback byte, plane 0:
 *   bit4 e1_open  (H - oe1 == E1out)      bit5 e2_open
 *   bit6 f1_open  (H[j-1] - oe1 == F1)    bit7 f2_open
 *   rows with ONE predecessor : low nibble = 12 if the diagonal hits, else ehit*4 + fout
 *                                (ehit 0 none / 1 E1 / 2 E2)
 *   rows with SEVERAL         : low nibble = (E1hit | E2hit<<1)<<2 | fout, and three more planes:
 *                                first predecessor whose diagonal hits (255 none), first arg-max
 *                                of E1, first arg-max of E2
 *   fout: 0 none, 1 insertion opened here (-> state M|E), 2 F1 extends, 3 F2 extends
 */
__device__ int dp_align(const KernelArgs &A, const Slot &S, int par, int N, const uint8_t *__restrict__ q, int qlen,
                        int *ring, int4 *ring_info, int lane, AlnState &R) {
    const DevParams &P = A.P;
    const int wcap = A.wcap;
    const uint32_t *in_off = S.in_off[par], *in_row = S.in_row[par];
    /* lane width abPOA would have used -> SIMD vector length the band is rounded to */
    {
        const int len = max(qlen, N);
        const long long max_score = max((long long)qlen * P.match, (long long)len * P.e1 + P.o1);
        const bool is16 = max_score <= 32767 - P.mismatch - P.o1 - P.e1 - P.o2 - P.e2;
        R.pn = is16 ? P.pn16 : P.pn32;
        R.bits = is16 ? 16 : 32;
    }
    const int pn = R.pn;
    const int w = P.wb < 0 ? qlen : P.wb + (int)__fmul_rn(P.wf, (float)qlen);
    uint32_t tb_used = 0, sp_used = 0;
    R.cells = R.intops = R.full = 0;

    /* row 0: the source */
    int4 prev_info;
    {
        const int rem0 = S.remain[0];
        const int e = min(qlen, max(0, qlen - rem0) + w);
        const int end_sn = e / pn;
        const int hi = min((end_sn + 1) * pn - 1, qlen);
        const int width = hi + 1;
        if (width > wcap) return ST_RETRY;
        int *H = ring, *E1 = ring + wcap, *E2 = ring + 2 * wcap;
        for (int c = lane; c < width; c += 32) {
            H[c] = c == 0 ? 0 : max(-(P.o1 + P.e1 * c), -(P.o2 + P.e2 * c));
            E1[c] = c == 0 ? -P.oe1 : NEG;
            E2[c] = c == 0 ? -P.oe2 : NEG;
        }
        prev_info = make_int4(0, end_sn, 0, 0);
        if (lane == 0) {
            ring_info[0] = prev_info;
            S.rowinfo[0] = prev_info;
            S.tboff[0] = 0;
        }
        if (S.meta[0] & META_FAR) {
            if ((uint64_t)sp_used + 3u * width > A.L.spcap) return ST_RETRY;
            for (int c = lane; c < width; c += 32) {
                S.spill[sp_used + c] = H[c];
                S.spill[sp_used + width + c] = E1[c];
                S.spill[sp_used + 2 * width + c] = E2[c];
            }
            if (lane == 0) S.spoff[0] = sp_used;
            sp_used += 3u * width;
        }
        __syncwarp();
    }

    for (int w0 = 1; w0 < N - 1; w0 += 32) {
        /* row metadata of 32 rows at once */
        int m_meta = 0, m_in0 = 0, m_in1 = 0, m_rem = 0, m_p0 = 0;
        {
            const int r = w0 + lane;
            if (r < N - 1) {
                m_meta = (int)S.meta[r];
                m_in0 = (int)in_off[r];
                m_in1 = (int)in_off[r + 1];
                m_rem = S.remain[r];
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

            /* band from the predecessors' row maxima (pull form of abPOA's max_pos_left/right) */
            int left = N, right = 0, minb = INT_MAX, maxe = -1;
            for (int k = 0; k < npre; ++k) {
                const int p = k == 0 ? p0 : (int)in_row[in0 + k];
                const int4 pi = (p == i - 1) ? prev_info : ((i - p < RING) ? ring_info[p % RING] : S.rowinfo[p]);
                left = min(left, pi.z + 1);
                right = max(right, pi.w + 1);
                minb = min(minb, pi.x);
                maxe = max(maxe, pi.y);
            }
            const int beg = max(0, min(left, qlen - rem) - w);
            const int end = min(qlen, max(right, qlen - rem) + w);
            const int beg_sn = max(beg / pn, minb);
            const int end_sn = min(end / pn, maxe + 1);
            const int dp_beg = beg_sn * pn;
            const int hi_cell = min((end_sn + 1) * pn - 1, qlen);
            const int width = max(0, hi_cell - dp_beg + 1);
            if (width > wcap) return ST_RETRY;
            const int nplanes = npre > 1 ? 4 : 1;
            const uint32_t tbo = tb_used;
            if ((uint64_t)tb_used + (uint64_t)width * nplanes > A.L.tbcap) return ST_RETRY;
            tb_used += (uint32_t)width * nplanes;
            const bool far = (meta & META_FAR) != 0;
            uint32_t spo = 0;
            if (far) {
                if ((uint64_t)sp_used + 3u * width > A.L.spcap) return ST_RETRY;
                spo = sp_used;
                sp_used += 3u * width;
            }
            R.cells += width;
            R.intops += 17ull * width + 3ull * (unsigned)max(0, npre - 1) * width;
            R.full += qlen + 1;

            int *Hr = ring + (i % RING) * 3 * wcap, *E1r = Hr + wcap, *E2r = Hr + 2 * wcap;
            int carry_h = NEG, carry_s1 = NEG - P.oe1, carry_s2 = NEG - P.oe2;
            int rmax = NEG, lpos = -1, rpos = -1, last_h = NEG;
            const int nch = (width + 31) >> 5;
            for (int c = 0; c < nch; ++c) {
                const int col = c * 32 + lane;
                const int j = dp_beg + col;
                const bool cv = col < width;
                int mx = NEG, ei1 = NEG, ei2 = NEG, a1 = 255, a2 = 255, hraw0 = NEG;
                bool rawok0 = false;
                for (int k = 0; k < npre; ++k) {
                    const int p = k == 0 ? p0 : (int)in_row[in0 + k];
                    const bool near = i - p < RING;
                    const int4 pi = (p == i - 1) ? prev_info : (near ? ring_info[p % RING] : S.rowinfo[p]);
                    const int pbeg = pi.x * pn, pend = (pi.y + 1) * pn - 1;
                    const int lo = max(beg_sn, pi.x) * pn;
                    const int hi = min((min(end_sn, pi.y) + 1) * pn - 1, qlen);
                    const int *Hp;
                    int stride;
                    if (near) { Hp = ring + (p % RING) * 3 * wcap; stride = wcap; }
                    else { stride = max(0, min(pend, qlen) - pbeg + 1); Hp = S.spill + S.spoff[p]; }
                    const bool raw_ok = cv && (j - 1 >= pbeg) && (j - 1 <= pend);
                    const int hraw = raw_ok ? Hp[j - 1 - pbeg] : NEG;
                    const bool e_ok = cv && j >= lo && j <= hi;
                    const int mval = (e_ok && j > lo) ? hraw : NEG;
                    mx = max(mx, mval);
                    if (e_ok) {
                        const int e1v = Hp[stride + j - pbeg], e2v = Hp[2 * stride + j - pbeg];
                        if (e1v > ei1) { ei1 = e1v; a1 = k; }
                        if (e2v > ei2) { ei2 = e2v; a2 = k; }
                    }
                    if (k == 0) { hraw0 = hraw; rawok0 = raw_ok; }
                }
                int s = 0;
                if (cv && j > 0) {
                    const int qb = q[j - 1];
                    s = (nbase >= 4 || qb >= 4) ? 0 : (nbase == qb ? P.match : -P.mismatch);
                }
                const int hh = cv ? max(mx + s, max(ei1, ei2)) : NEG;
                /* insertion scores: F[j] = max(hh[j-1] - oe, F[j-1] - e) along the row */
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
                int hprev = __shfl_up_sync(FULL, h, 1);
                if (lane == 0) hprev = carry_h;
                carry_h = __shfl_sync(FULL, h, 31);
                if (cv) {
                    Hr[col] = h; E1r[col] = e1o; E2r[col] = e2o;
                    if (far) {
                        S.spill[spo + col] = h;
                        S.spill[spo + width + col] = e1o;
                        S.spill[spo + 2 * width + col] = e2o;
                    }
                    const bool e1h = ei1 == h, e2h = ei2 == h, f1h = f1 == h, f2h = f2 == h;
                    const bool f1open = hprev - P.oe1 == f1, f2open = hprev - P.oe2 == f2;
                    const int fout = f1h ? (f1open ? 1 : 2) : (f2h ? (f2open ? 1 : 3) : 0);
                    const int hin = ((h - P.oe1 == e1o) ? 16 : 0) | ((h - P.oe2 == e2o) ? 32 : 0) |
                                    (f1open ? 64 : 0) | (f2open ? 128 : 0);
                    if (npre == 1) {
                        const bool mhit = rawok0 && (hraw0 + s == h);
                        const int ehit = e1h ? 1 : (e2h ? 2 : 0);
                        S.tb[tbo + col] = (uint8_t)(hin | (mhit ? 12 : ehit * 4 + fout));
                    } else {
                        int kM = 255;
                        for (int k = 0; k < npre; ++k) {
                            const int p = k == 0 ? p0 : (int)in_row[in0 + k];
                            const bool near = i - p < RING;
                            const int4 pi = (p == i - 1) ? prev_info : (near ? ring_info[p % RING] : S.rowinfo[p]);
                            const int pbeg = pi.x * pn, pend = (pi.y + 1) * pn - 1;
                            if (j - 1 < pbeg || j - 1 > pend) continue;
                            const int *Hp = near ? ring + (p % RING) * 3 * wcap : S.spill + S.spoff[p];
                            if (Hp[j - 1 - pbeg] + s == h) { kM = k; break; }
                        }
                        S.tb[tbo + col] = (uint8_t)(hin | ((e1h ? 1 : 0) | (e2h ? 2 : 0)) << 2 | fout);
                        S.tb[tbo + width + col] = (uint8_t)kM;
                        S.tb[tbo + 2 * width + col] = (uint8_t)a1;
                        S.tb[tbo + 3 * width + col] = (uint8_t)a2;
                    }
                }
                /* row maximum, left-most and right-most position (drives the adaptive band) */
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
                S.rowinfo[i] = prev_info;
                S.tboff[i] = tbo;
                if (far) S.spoff[i] = spo;
                if (meta & META_TOSINK) S.rowbest[i] = width > 0 ? last_h : NEG;
            }
            __syncwarp();
        }
    }
    R.tbbytes = tb_used;

    /* best end cell: the sink's predecessors in edge order, first maximum wins */
    {
        const int s0 = (int)in_off[N - 1], s1 = (int)in_off[N];
        int best = NEG, bi = 0, bj = 0;
        for (int e = s0; e < s1; ++e) {
            const int p = (int)in_row[e];
            const int4 pi = S.rowinfo[p];
            const int v = S.rowbest[p];
            if (v > best) { best = v; bi = p; bj = min(qlen, (pi.y + 1) * pn - 1); }
        }
        R.best_i = bi; R.best_j = bj; R.best_score = best;
    }
    return ST_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* traceback                                                                                   */
/* ------------------------------------------------------------------------------------------ */

enum TbState { S_ALL = 0, S_MF, S_ME, S_E1, S_E2, S_F1, S_F2 };

/* Follows the flags from (best_i, best_j) to the source; writes qmap[t] = row the query base t
 * is aligned to, -1 for an inserted base.  Executed by lane 0.  Returns false when no move is
 * possible (abPOA would die in cg_backtrack; the reference then falls back to the first read). */
__device__ bool traceback(const Slot &S, int par, int qlen, const AlnState &R) {
    const uint32_t *in_off = S.in_off[par], *in_row = S.in_row[par];
    const int pn = R.pn;
    int i = R.best_i, j = R.best_j, state = S_ALL;
    for (int t = j; t < qlen; ++t) S.qmap[t] = -1;
    while (i > 0 && j > 0) {
        const int4 info = S.rowinfo[i];
        const int dp_beg = info.x * pn;
        const int width = min((info.y + 1) * pn - 1, qlen) - dp_beg + 1;
        const int col = j - dp_beg;
        if (col < 0 || col >= width) return false;
        const int in0 = (int)in_off[i], npre = (int)in_off[i + 1] - in0;
        const uint32_t tbo = S.tboff[i];
        const int fl = S.tb[tbo + col];
        int kM = 0, a1 = 0, a2 = 0, ebits, fout;
        bool mhit;
        if (npre == 1) {
            const int code = fl & 15;
            mhit = code == 12;
            const int ehit = mhit ? 0 : (code >> 2);
            ebits = ehit == 1 ? 1 : (ehit == 2 ? 2 : 0);
            fout = mhit ? 0 : (code & 3);
        } else {
            ebits = (fl >> 2) & 3;
            fout = fl & 3;
            kM = S.tb[tbo + width + col];
            mhit = kM != 255;
            a1 = S.tb[tbo + 2 * width + col];
            a2 = S.tb[tbo + 3 * width + col];
        }
        const bool hasM = state <= S_ME, hasE = state == S_ALL || state == S_ME, hasF = state == S_ALL || state == S_MF;
        if (hasM && mhit) {
            S.qmap[j - 1] = i;
            i = (int)in_row[in0 + kM];
            --j;
            state = S_ALL;
            continue;
        }
        int ek = -1, etype = 0;
        if (hasE && ebits) {
            if (ebits == 3) { if (a1 <= a2) { ek = a1; etype = 1; } else { ek = a2; etype = 2; } }
            else if (ebits == 1) { ek = a1; etype = 1; }
            else { ek = a2; etype = 2; }
        } else if (state == S_E1) { ek = a1; etype = 1; }
        else if (state == S_E2) { ek = a2; etype = 2; }
        if (etype) {
            if (ek >= npre) return false;
            const int p = (int)in_row[in0 + ek];
            const int4 pinfo = S.rowinfo[p];
            const int pbeg = pinfo.x * pn;
            const int pwidth = min((pinfo.y + 1) * pn - 1, qlen) - pbeg + 1;
            const int pcol = j - pbeg;
            if (pcol < 0 || pcol >= pwidth) return false;
            const int pfl = S.tb[S.tboff[p] + pcol];
            const bool open = etype == 1 ? (pfl & 16) : (pfl & 32);
            state = open ? S_MF : (etype == 1 ? S_E1 : S_E2);
            i = p;
            continue;
        }
        if (hasF && fout) {
            S.qmap[j - 1] = -1;
            --j;
            state = fout == 1 ? S_ME : (fout == 2 ? S_F1 : S_F2);
            continue;
        }
        if (state == S_F1 || state == S_F2) {
            const bool open = state == S_F1 ? (fl & 64) : (fl & 128);
            S.qmap[j - 1] = -1;
            --j;
            if (open) state = S_ME;
            continue;
        }
        return false;
    }
    for (int t = 0; t < j; ++t) S.qmap[t] = -1;
    return true;
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
__device__ int merge_read(const KernelArgs &A, const Slot &S, int &par, int &N, int &E, const uint8_t *__restrict__ q,
                          int qlen, int creator0, int32_t *tr_aln, int32_t *tr_node, int lane) {
    const int cur = par, nxt = par ^ 1;
    const uint8_t *base = S.base[cur], *sib = S.sib[cur];
    const uint32_t *out_off = S.out_off[cur], *out_row = S.out_row[cur], *in_off = S.in_off[cur], *in_row = S.in_row[cur];
    int32_t *out_w = S.out_w[cur];

    for (int r = lane; r < N; r += 32) { S.cnt[r] = 0; S.addin[r] = -1; S.addout[r] = -1; S.grow[r] = 0; }
    __syncwarp();

    /* U1: resolve every query base to an existing row or a new node; order keys */
    int carry_key = 0, carry_new = 0;
    for (int t0 = 0; t0 < qlen; t0 += 32) {
        const int t = t0 + lane;
        int isnew = 0, v = -1, key = -1, sibof = -1;
        if (t < qlen) {
            const int r = S.qmap[t];
            const int b = q[t];
            if (r >= 0) {
                if (base[r] == b) v = r;
                else {
                    const int sb = sib[r], before = sb >> 4, after = sb & 15;
                    for (int x = r - before; x <= r + after; ++x)
                        if (x != r && base[x] == b) v = x;
                    if (v < 0) { isnew = 1; sibof = r; key = r + after; }
                }
                if (!isnew) key = v + (sib[v] & 15);
                if (tr_aln) tr_aln[t] = S.creator[cur][r];
            } else {
                isnew = 1;
                if (tr_aln) tr_aln[t] = -1;
            }
            if (tr_node) tr_node[t] = isnew ? creator0 + t : S.creator[cur][v];
        }
        int ks = warp_incl_max(key, lane);
        ks = max(ks, carry_key);
        const int incl = warp_incl_sum(isnew, lane);
        const int nidx = carry_new + incl - isnew;
        if (t < qlen) {
            S.pv[t] = isnew ? -1 : v;
            S.pkey[t] = ks;
            S.pnew[t] = nidx;
            S.psib[t] = sibof;
            if (isnew) atomicAdd(&S.cnt[ks], 1);
            if (sibof >= 0) {
                const int sb = sib[sibof];
                for (int x = sibof - (sb >> 4); x <= sibof + (sb & 15); ++x) S.grow[x] = 1;
            }
        }
        carry_key = __shfl_sync(FULL, ks, 31);
        carry_new += __shfl_sync(FULL, incl, 31);
    }
    const int n_new = carry_new;
    const int N2 = N + n_new;
    if ((uint32_t)N2 > A.L.ncap) return ST_RETRY;
    __syncwarp();

    /* U2: shift[r] = number of new nodes placed before old row r (exclusive scan of cnt) */
    {
        int carry = 0;
        for (int r0 = 0; r0 < N; r0 += 32) {
            const int r = r0 + lane;
            const int c = r < N ? S.cnt[r] : 0;
            const int incl = warp_incl_sum(c, lane);
            if (r < N) {
                const int sh = carry + incl - c;
                S.cnt[r] = sh;
                S.srcof[r + sh] = r;
            }
            carry += __shfl_sync(FULL, incl, 31);
        }
        for (int t = lane; t < qlen; t += 32)
            if (S.pv[t] < 0) S.srcof[S.pkey[t] + 1 + S.pnew[t]] = -(t + 1);
    }
    __syncwarp();

    /* U3: the path edges u[t-1] -> u[t], t = 0..qlen (u[-1] = source, u[qlen] = sink) */
    int n_new_edges = 0;
    for (int t = lane; t <= qlen; t += 32) {
        const int from_old = t == 0 ? 0 : S.pv[t - 1];
        const int to_old = t == qlen ? N - 1 : S.pv[t];
        const int from_new = from_old >= 0 ? from_old + S.cnt[from_old] : S.pkey[t - 1] + 1 + S.pnew[t - 1];
        const int to_new = to_old >= 0 ? to_old + S.cnt[to_old] : S.pkey[t] + 1 + S.pnew[t];
        bool found = false;
        if (from_old >= 0 && to_old >= 0) {
            const uint32_t o0 = out_off[from_old], o1 = out_off[from_old + 1];
            for (uint32_t e = o0; e < o1; ++e)
                if ((int)out_row[e] == to_old) { out_w[e] += 1; found = true; break; }
        }
        if (!found) {
            ++n_new_edges;
            if (from_old >= 0) S.addout[from_old] = to_new; else S.nout[t - 1] = to_new;
            if (to_old >= 0) S.addin[to_old] = from_new; else S.nin[t] = from_new;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_new_edges += __shfl_xor_sync(FULL, n_new_edges, d);
    const int E2 = E + n_new_edges;
    if ((uint32_t)E2 > A.L.ecap) return ST_RETRY;
    __syncwarp();

    /* U4: emit the merged graph */
    {
        int carry_in = 0, carry_out = 0;
        for (int r0 = 0; r0 < N2; r0 += 32) {
            const int nr = r0 + lane;
            int din = 0, dout = 0, src = 0;
            if (nr < N2) {
                src = S.srcof[nr];
                if (src >= 0) {
                    din = (int)(in_off[src + 1] - in_off[src]) + (S.addin[src] >= 0);
                    dout = (int)(out_off[src + 1] - out_off[src]) + (S.addout[src] >= 0);
                } else din = dout = 1;
            }
            const int iin = warp_incl_sum(din, lane), iout = warp_incl_sum(dout, lane);
            if (nr < N2) {
                uint32_t io = carry_in + iin - din, oo = carry_out + iout - dout;
                S.in_off[nxt][nr] = io;
                S.out_off[nxt][nr] = oo;
                if (src >= 0) {
                    for (uint32_t e = in_off[src]; e < in_off[src + 1]; ++e) {
                        const int x = (int)in_row[e];
                        S.in_row[nxt][io++] = x + S.cnt[x];
                    }
                    if (S.addin[src] >= 0) S.in_row[nxt][io++] = S.addin[src];
                    for (uint32_t e = out_off[src]; e < out_off[src + 1]; ++e) {
                        const int y = (int)out_row[e];
                        S.out_row[nxt][oo] = y + S.cnt[y];
                        S.out_w[nxt][oo++] = out_w[e];
                    }
                    if (S.addout[src] >= 0) { S.out_row[nxt][oo] = S.addout[src]; S.out_w[nxt][oo++] = 1; }
                    S.base[nxt][nr] = base[src];
                    S.sib[nxt][nr] = (uint8_t)(sib[src] + S.grow[src]);
                    S.creator[nxt][nr] = S.creator[cur][src];
                } else {
                    const int t = -src - 1;
                    S.in_row[nxt][io] = S.nin[t];
                    S.out_row[nxt][oo] = S.nout[t];
                    S.out_w[nxt][oo] = 1;
                    S.base[nxt][nr] = q[t];
                    const int so = S.psib[t];
                    int sb = 0;
                    if (so >= 0) { const int o = sib[so]; sb = ((o >> 4) + (o & 15) + 1) << 4; }
                    S.sib[nxt][nr] = (uint8_t)sb;
                    S.creator[nxt][nr] = creator0 + t;
                }
            }
            carry_in += __shfl_sync(FULL, iin, 31);
            carry_out += __shfl_sync(FULL, iout, 31);
        }
        if (lane == 0) { S.in_off[nxt][N2] = carry_in; S.out_off[nxt][N2] = carry_out; }
    }
    __syncwarp();
    par = nxt; N = N2; E = E2;
    return ST_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* consensus                                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* heaviest bundling (semantics of abPOA's abpoa_heaviest_bundling, one consensus): reverse sweep
 * over the row order, then the path source -> sink.  Lane 0.  Returns length or -1 (capacity). */
__device__ int heaviest_bundle(const Slot &S, int par, int N, uint8_t *cons, int cap) {
    const uint32_t *out_off = S.out_off[par], *out_row = S.out_row[par];
    const int32_t *out_w = S.out_w[par];
    int32_t *score = S.cnt, *maxout = S.addin;
    score[N - 1] = 0;
    maxout[N - 1] = -1;
    for (int r = N - 2; r >= 0; --r) {
        const uint32_t o0 = out_off[r], o1 = out_off[r + 1];
        int max_id = -1;
        if (r == 0) {
            int path_score = -1, path_max_w = -1;
            for (uint32_t e = o0; e < o1; ++e) {
                const int t = (int)out_row[e], w = out_w[e];
                if (w > path_max_w || (w == path_max_w && score[t] > path_score)) {
                    max_id = t; path_score = score[t]; path_max_w = w;
                }
            }
        } else {
            int max_w = INT_MIN;
            for (uint32_t e = o0; e < o1; ++e) {
                const int t = (int)out_row[e], w = out_w[e];
                if (max_w < w) { max_w = w; max_id = t; }
                else if (max_w == w && score[max_id] <= score[t]) max_id = t;
            }
            score[r] = max_w + score[max_id];
        }
        maxout[r] = max_id;
    }
    int len = 0, curr = maxout[0];
    while (curr != N - 1 && curr >= 0) {
        if (len >= cap) return -1;
        cons[len++] = "ACGTN"[S.base[par][curr]];
        curr = maxout[curr];
    }
    return len;
}

/* ------------------------------------------------------------------------------------------ */
/* the persistent kernel                                                                       */
/* ------------------------------------------------------------------------------------------ */

__device__ int process_group(const KernelArgs &A, const Slot &S, int g, int *ring, int4 *ring_info, int lane,
                             unsigned long long *st) {
    const int64_t r0 = A.group_read_off[g], r1 = A.group_read_off[g + 1];
    if (r1 <= r0) return ST_EMPTY;
    const int64_t gbase = A.read_off[r0];
    int par = 0, N = 0, E = 0;
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t b0 = A.read_off[r], b1 = A.read_off[r + 1];
        const int len = (int)(b1 - b0);
        const uint8_t *seq = A.codes + b0;
        const int creator0 = (int)(b0 - gbase);
        int32_t *tr_aln = A.tr_aln ? A.tr_aln + b0 : nullptr;
        int32_t *tr_node = A.tr_node ? A.tr_node + b0 : nullptr;
        if (lane == 0) {
            if (A.tr_score) A.tr_score[r] = 0;
            if (A.tr_bits) A.tr_bits[r] = 0;
            if (A.tr_cells) A.tr_cells[r] = 0;
        }
        if (N == 0) {
            if (len <= 0) return ST_EMPTY;
            if ((uint32_t)(len + 2) > A.L.ncap || (uint32_t)(len + 1) > A.L.ecap) return ST_RETRY;
            init_graph(S, par, seq, len, creator0, lane);
            N = len + 2; E = len + 1;
            for (int t = lane; t < len; t += 32) {
                if (tr_aln) tr_aln[t] = -1;
                if (tr_node) tr_node[t] = creator0 + t;
            }
            continue;
        }
        if (len <= 0) continue;
        if ((uint32_t)len > A.L.qcap) return ST_RETRY;
        long long tk0 = clock64();
        remain_pass(S, par, N, lane);
        long long tk1 = clock64();
        st[SI_T_PREP] += tk1 - tk0;
        AlnState R;
        const int rc = dp_align(A, S, par, N, seq, len, ring, ring_info, lane, R);
        if (rc != ST_OK) return rc;
        tk0 = clock64();
        st[SI_T_DP] += tk0 - tk1;
        st[SI_CELLS] += R.cells; st[SI_INTOPS] += R.intops; st[SI_FULL] += R.full; st[SI_ALN] += 1;
        st[R.bits == 16 ? SI_ALN16 : SI_ALN32] += 1; st[SI_TB] += R.tbbytes;
        if (lane == 0) {
            if (A.tr_score) A.tr_score[r] = R.best_score;
            if (A.tr_bits) A.tr_bits[r] = R.bits;
            if (A.tr_cells) A.tr_cells[r] = (long long)R.cells;
        }
        __syncwarp();
        int ok = 1;
        if (lane == 0) ok = traceback(S, par, len, R) ? 1 : 0;
        ok = __shfl_sync(FULL, ok, 0);
        __syncwarp();
        if (!ok) return ST_EMPTY;
        tk1 = clock64();
        st[SI_T_TB] += tk1 - tk0;
        const int mrc = merge_read(A, S, par, N, E, seq, len, creator0, tr_aln, tr_node, lane);
        if (mrc != ST_OK) return mrc;
        st[SI_T_MERGE] += clock64() - tk1;
    }
    if (N <= 2) return ST_EMPTY;
    const long long tc0 = clock64();
    int clen = 0;
    if (lane == 0) {
        const int cap = (int)(A.cons_off[g + 1] - A.cons_off[g]);
        clen = heaviest_bundle(S, par, N, A.cons + A.cons_off[g], cap);
    }
    clen = __shfl_sync(FULL, clen, 0);
    __syncwarp();
    if (clen < 0) return ST_RETRY;
    if (lane == 0) A.cons_len[g] = clen;
    st[SI_T_CONS] += clock64() - tc0;
    return ST_OK;
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) poa_group_kernel(const KernelArgs A) {
    extern __shared__ __align__(16) int smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = RING * 3 * A.wcap + RING * 4;
    int *ring = smem + warp * per_warp;
    int4 *ring_info = reinterpret_cast<int4 *>(ring + RING * 3 * A.wcap);
    const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
    const Slot S = make_slot(A, slot);
    unsigned long long st[SI_COUNT];
#pragma unroll
    for (int k = 0; k < SI_COUNT; ++k) st[k] = 0;
    for (;;) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(A.queue_head, 1);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= A.n_queue) break;
        const int g = A.queue[qi];
        unsigned long long gst[SI_COUNT];
#pragma unroll
        for (int k = 0; k < SI_COUNT; ++k) gst[k] = 0;
        const long long tg0 = clock64();
        const int rc = process_group(A, S, g, ring, ring_info, lane, gst);
        gst[SI_T_BUSY] += clock64() - tg0;
        if (lane == 0) {
            A.status[g] = rc;
            if (rc != ST_OK) A.cons_len[g] = 0;
        }
        if (rc != ST_RETRY) {
#pragma unroll
            for (int k = 0; k < SI_COUNT; ++k) st[k] += gst[k];
        }
        __syncwarp();
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < SI_COUNT; ++k)
            if (st[k]) atomicAdd(A.stats + k, st[k]);
    }
}

/* host-callable launchers (used by poa_capi.cu) */
cudaError_t launch_encode(const uint8_t *ascii, uint8_t *codes, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (n + threads * 16 - 1) / (threads * 16);
    if (blocks > 148 * 16) blocks = 148 * 16;
    encode_bases_kernel<<<(unsigned)blocks, threads, 0, stream>>>(ascii, codes, n);
    return cudaGetLastError();
}

size_t poa_smem_bytes(int wcap, int warps_per_block) {
    return (size_t)warps_per_block * (RING * 3 * wcap + RING * 4) * sizeof(int);
}

cudaError_t launch_poa(const KernelArgs &A, int n_blocks, int warps_per_block, cudaStream_t stream) {
    const size_t smem = poa_smem_bytes(A.wcap, warps_per_block);
    cudaError_t e = cudaFuncSetAttribute(poa_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    poa_group_kernel<<<n_blocks, warps_per_block * 32, smem, stream>>>(A);
    return cudaGetLastError();
}

int poa_max_blocks_per_sm(int wcap, int warps_per_block) {
    int nb = 0;
    const size_t smem = poa_smem_bytes(wcap, warps_per_block);
    if (cudaFuncSetAttribute(poa_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, poa_group_kernel, warps_per_block * 32, smem) != cudaSuccess) return 0;
    return nb;
}

/* consensus regions -> one compact buffer; one warp per group, 1 B in / 1 B out per base */
__global__ void gather_consensus_kernel(const uint8_t *__restrict__ cons, const int64_t *__restrict__ region_off,
                                        const int32_t *__restrict__ cons_len, const int64_t *__restrict__ out_off,
                                        uint8_t *__restrict__ out, int64_t n_groups) {
    const int lane = threadIdx.x & 31;
    int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; g < n_groups; g += nw) {
        const int64_t n = out_off[g + 1] - out_off[g];
        const uint8_t *src = cons + region_off[g];
        uint8_t *dst = out + out_off[g];
        (void)cons_len;
        for (int64_t k = lane; k < n; k += 32) dst[k] = src[k];
    }
}

cudaError_t launch_gather(const uint8_t *cons, const int64_t *region_off, const int32_t *cons_len,
                          const int64_t *out_off, uint8_t *out, int64_t n_groups, cudaStream_t stream) {
    if (n_groups <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (n_groups * 32 + threads - 1) / threads;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gather_consensus_kernel<<<(unsigned)blocks, threads, 0, stream>>>(cons, region_off, cons_len, out_off, out, n_groups);
    return cudaGetLastError();
}

}  // namespace mpoa
