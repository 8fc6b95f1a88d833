/*
 * poa_kernels.cu -- sm_100a kernels of the consensus hot path.
 *
 * Replaces, for a whole batch of isoform read groups, what the reference obtains from one
 * `abpoa -M 5 -r 0 in.fasta` process per isoform (reference utils/SpliceDefineConsensus.py:917,
 * loop at defineIsoforms.py:87-91).  Algorithmic contract (scores, band, tie-breaks): DESIGN.md
 * section "Algorithm"; data layout: poa_device.cuh.
 *
 * Kernels:
 *   encode_bases_kernel      ASCII -> nt4 codes (HBM bound, 1 B in / 1 B out per base)
 *   poa_group_kernel<T,WPL>  persistent, one TEAM of T lanes per read group (T = 16: two groups per
 *                            warp): graph build, adaptive-banded convex-gap DP (WPL = 0 int32
 *                            lanes, WPL = 4/6/8 packed int16x2 DPX words per lane), value
 *                            traceback, graph merge, heaviest-bundle consensus
 *   gather_consensus_kernel  per-group consensus regions -> one compact buffer
 */
#include "poa_traceback.cuh"
#include "poa_seed.cuh"

namespace mpoa {

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
/* the persistent kernel                                                                       */
/* ------------------------------------------------------------------------------------------ */

/* per-team state of the group in flight (registers) */
struct GroupState {
    int g;               // group index, -1: none
    int64_t r, r1;       // next read, end of the group's reads
    int64_t gbase;       // first base of the group
    int N, E, par;
    long long t0;
};

/*
 * One read of the group in flight of every team of the warp (lockstep, see Team): the first read
 * becomes the chain graph, every other read is aligned (remain pass, banded DP, traceback) and
 * merged.  `act`: this team has a read to process.  Returns ST_PENDING while the team's group
 * goes on, a final status otherwise.
 * TRACE: the per-read / per-base trace outputs of the ABI (parity tests).  It is a compile-time
 * switch because the kernel's instruction footprint matters (resident warps share the
 * instruction cache).
 */
template <int T, int WPL, bool TRACE, bool SEEDED>
__device__ __forceinline__ int read_step(const KernelArgs &A, const Team<T> &tm, int slot, GroupState &G, bool act, int *ring, int4 *ring_info,
                                         unsigned long long *st /* per-team counters in shared memory, lane 0 only */) {
    const int lane = tm.tl;
    int rc = ST_PENDING;
    int64_t r = 0, b0 = 0;
    int len = 0;
    if (act) {
        r = G.r++;
        b0 = A.read_off[r];
        len = (int)(A.read_off[r + 1] - b0);
    }
    const uint8_t *seq = A.codes + b0;
    const int creator0 = (int)(b0 - G.gbase);
    Slot S = make_slot(A, slot, G.par);
    int32_t *tr_aln = nullptr, *tr_node = nullptr;
    if constexpr (TRACE) {
        tr_aln = A.tr_aln + b0; tr_node = A.tr_node + b0;
        if (act && lane == 0) { A.tr_score[r] = 0; A.tr_bits[r] = 0; A.tr_cells[r] = 0; }
    }
    bool aln = false;                         // this team aligns a read in this step
    if (act) {
        if (G.N == 0) {                       // first read of the group (no cross-lane operation in here)
            if (len <= 0) rc = ST_EMPTY;
            else if ((uint32_t)(len + 2) > A.L.ncap || (uint32_t)(len + 1) > A.L.ecap) rc = ST_RETRY;
            else {
                init_graph<T, TRACE, SEEDED>(A, S, tm, seq, len, creator0);
                G.N = len + 2; G.E = len + 1;
                if constexpr (TRACE) {
                    for (int t = lane; t < len; t += T) { tr_aln[t] = -1; tr_node[t] = creator0 + t; }
                }
            }
        } else if (len > 0) {
            if ((uint32_t)len > A.L.qcap) rc = ST_RETRY;
            else aln = true;
        }
    }
    tm.sync();
    if (!tm.wany(aln)) return rc;

    long long tk0 = clock64();
    remain_pass<T>(A, S, tm, G.N, aln);
    long long tk1 = clock64();
    if (aln && lane == 0) st[SI_T_PREP] += tk1 - tk0;
    AlnState R;
    if constexpr (!SEEDED) {
        int drc;
        if constexpr (WPL == 0) drc = dp_align32(A, S, G.N, seq, len, ring, ring_info, lane, R);
        else drc = dp_band16<T, WPL>(A, S, tm, G.N, seq, len, reinterpret_cast<uint32_t *>(ring), R, aln);
        if (aln && drc != ST_OK) { rc = drc; aln = false; }
        tk0 = clock64();
        if (aln && lane == 0) {
            st[SI_T_DP] += tk0 - tk1;
            st[SI_CELLS] += R.cells; st[SI_INTOPS] += R.intops; st[SI_FULL] += R.full; st[SI_ALN] += 1;
            st[R.bits == 16 ? SI_ALN16 : SI_ALN32] += 1; st[SI_TB] += R.tbbytes;
            if constexpr (TRACE) { A.tr_score[r] = R.best_score; A.tr_bits[r] = R.bits; A.tr_cells[r] = (long long)R.cells; }
        }
        tm.sync();
        if (!tm.wany(aln)) return rc;
        bool ok;
        if constexpr (WPL == 0) ok = traceback<int32_t, 0, T>(A, S, tm, seq, len, R, ring, aln);
        else ok = traceback<int16_t, 2 * WPL, T>(A, S, tm, seq, len, R, ring, aln);
        tm.sync();
        if (aln && !ok) { rc = ST_EMPTY; aln = false; }
        tk1 = clock64();
        if (aln && lane == 0) st[SI_T_TB] += tk1 - tk0;
    } else {
        /*
         * `abpoa -S`: the read is aligned window by window between its anchors (exact k-mers shared
         * with the previous read, csrc/seed.cpp).  A window = the stretch of the read before an anchor
         * against the sub-graph between the previous anchor's last node (or the source) and the
         * anchor's first node; the anchor k-mer itself is k forced matches onto the previous read's
         * nodes; the last window ends at the sink.  One merge per read, as without seeding.
         */
        const int k = A.seed_k;
        const int32_t *prevrow = prevrow_p(A, S);
        int32_t *qmap = qmap_p(A, S);
        const int a0 = aln ? A.anc_off[r] + 1 : 0, n_anc = aln ? __ldcg(A.anc + (a0 - 1)).x : -1;
        const int n_win = tm.wmax(n_anc + 1);
        Slot V = S;
        V.sub = 1;
        long long score = 0;
        unsigned long long w_cells = 0, w_intops = 0, w_full = 0, w_tb = 0;
        int bits = 0, rb = 0, qbeg = 0;
        for (int wi = 0; wi < n_win; ++wi) {
            const bool won = aln && wi <= n_anc;
            int re = G.N - 1, qend = len, t0 = 0;
            if (won && wi < n_anc) { const int2 an = __ldcg(A.anc + (a0 + wi)); t0 = an.x; qend = an.y; re = prevrow[t0]; }
            const int ql = qend - qbeg;
            bool run = won && ql > 0;
            if (tm.wany(run)) {
                const int Ns = extract_window<T>(A, S, tm, rb, re, run);
                if (run && Ns < 2) { rc = ST_EMPTY; aln = false; run = false; }
                const uint8_t *wq = seq + qbeg;
                int drc;
                if constexpr (WPL == 0) drc = run ? dp_align32(A, V, Ns, wq, ql, ring, ring_info, lane, R) : ST_OK;
                else drc = dp_band16<T, WPL>(A, V, tm, Ns, wq, ql, reinterpret_cast<uint32_t *>(ring), R, run);
                if (run && drc != ST_OK) { rc = drc; aln = false; run = false; }
                tm.sync();
                bool ok;
                if constexpr (WPL == 0) ok = traceback<int32_t, 0, T>(A, V, tm, wq, ql, R, ring, run);
                else ok = traceback<int16_t, 2 * WPL, T>(A, V, tm, wq, ql, R, ring, run);
                tm.sync();
                if (run && !ok) { rc = ST_EMPTY; aln = false; run = false; }
                if (run) {
                    score += R.best_score; bits = max(bits, R.bits);
                    w_cells += R.cells; w_intops += R.intops; w_full += R.full; w_tb += R.tbbytes;
                    /* rows of the view -> rows of the graph */
                    const int32_t *vq = qmap_p(A, V), *sub2row = srcof_p(A, S);
#pragma unroll 1
                    for (int t = lane; t < ql; t += T) { const int sr = vq[t]; qmap[qbeg + t] = sr < 0 ? -1 : sub2row[sr]; }
                }
                tm.sync();
            }
            if (aln && wi < n_anc) {                   // the anchor k-mer: forced matches onto the previous read's nodes
#pragma unroll 1
                for (int j = lane; j < k; j += T) qmap[qend + j] = prevrow[t0 + j];
                rb = prevrow[t0 + k - 1]; qbeg = qend + k;
                score += (long long)k * A.P.match;
            }
        }
        tm.sync();
        tk0 = clock64();
        if (aln && lane == 0) {
            st[SI_T_DP] += tk0 - tk1;
            st[SI_CELLS] += w_cells; st[SI_INTOPS] += w_intops; st[SI_FULL] += w_full; st[SI_ALN] += 1;
            st[(bits == 32) ? SI_ALN32 : SI_ALN16] += 1; st[SI_TB] += w_tb;
            if constexpr (TRACE) { A.tr_score[r] = (int)score; A.tr_bits[r] = bits ? bits : 16; A.tr_cells[r] = (long long)w_cells; }
        }
        tk1 = tk0;
    }
    if (!tm.wany(aln)) return rc;
    const int mrc = merge_read<T, SEEDED>(A, S, tm, G.par, G.N, G.E, seq, len, creator0, tr_aln, tr_node, aln);
    if (aln && mrc != ST_OK) rc = mrc;
    if (aln && lane == 0) st[SI_T_MERGE] += clock64() - tk1;
    return rc;
}

/* after the last read: heaviest-bundle consensus into the group's region (`fin`: this team finishes a group) */
template <int T>
__device__ __forceinline__ int finish_group(const KernelArgs &A, const Team<T> &tm, int slot, const GroupState &G, bool fin, unsigned long long *st,
                                            int *scratch /* the team's shared memory (free between reads) */) {
    int clen = 0;
    const long long tc0 = clock64();
    {
        const bool on = fin && G.N > 2;
        const Slot S = make_slot(A, slot, G.par);
        const int64_t c0 = on ? A.cons_off[G.g] : 0;
        const int cap = on ? (int)(A.cons_off[G.g + 1] - c0) : 0;
        tm.sync();
        clen = heaviest_bundle<T>(A, S, tm, G.N, A.cons + c0, cap, scratch, on);
    }
    tm.sync();
    if (!fin) return ST_PENDING;
    if (G.N <= 2) return ST_EMPTY;
    if (clen < 0) return ST_RETRY;
    if (tm.tl == 0) { A.cons_len[G.g] = clen; st[SI_T_CONS] += clock64() - tc0; }
    return ST_OK;
}

/* shared-memory words of one team: DP ring (also the traceback scratch) + its work counters (SI_COUNT x u64) */
template <int T, int WPL>
__host__ __device__ constexpr int team_words(int wcap) {
    int w = WPL == 0 ? RING * 3 * wcap + RING * 4 : ring16_team_words<T, (WPL == 0 ? 2 : WPL)>();
    if (w < tb_scratch_words<T>()) w = tb_scratch_words<T>();
    return ((w + 3) & ~3) + 32;
}

#ifndef MPOA_MINB_16_4
#define MPOA_MINB_16_4 4
#endif
#ifndef MPOA_MINB_16_6
#define MPOA_MINB_16_6 3
#endif
#ifndef MPOA_MINB_32_4
#define MPOA_MINB_32_4 5
#endif
#ifndef MPOA_MINB_32_2
#define MPOA_MINB_32_2 5
#endif
#ifndef MPOA_MINB_32_8
#define MPOA_MINB_32_8 3
#endif
#ifndef MPOA_MINB_32_0
#define MPOA_MINB_32_0 3
#endif
template <int T, int WPL>
__host__ __device__ constexpr int min_blocks() {
    return WPL == 0 ? MPOA_MINB_32_0 : T == 16 ? (WPL <= 4 ? MPOA_MINB_16_4 : MPOA_MINB_16_6) : (WPL <= 2 ? MPOA_MINB_32_2 : WPL <= 4 ? MPOA_MINB_32_4 : MPOA_MINB_32_8);
}

/*
 * Persistent kernel: every TEAM (T lanes) pulls read groups from the launch's queue (heaviest
 * first) and carries each one from its first read to the consensus.  The loop is flat -- one iteration = one read of every team's group -- and the
 * teams of a warp run it in lockstep (see Team): a team without work idles by predicate until the
 * other one is done as well.
 */
template <int T, int WPL, bool TRACE, bool SEEDED>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, (min_blocks<T, WPL>()))
poa_group_kernel(const __grid_constant__ KernelArgs A) {
    static_assert(WPL != 0 || T == 32, "int32 lanes use the whole warp");
    extern __shared__ __align__(16) int smem[];
    constexpr int TPW = 32 / T;                       // teams per warp
    const int warp = threadIdx.x >> 5;
    /* the lane index goes through shared memory once: ptxas otherwise re-derives it from the
     * thread-id register (S2R + LOP3) all over the row loop instead of keeping it in a register */
    int wl = threadIdx.x & 31;
    {
        volatile int *lane_box = smem + (blockDim.x >> 5) * (32 / T) * team_words<T, WPL>(A.wcap);
        lane_box[threadIdx.x] = wl;
        wl = lane_box[threadIdx.x];
    }
    const Team<T> tm(wl);
    const int team = warp * TPW + wl / T;
    const int per_team = team_words<T, WPL>(A.wcap);
    int *ring = smem + team * per_team;
    int4 *ring_info = reinterpret_cast<int4 *>(ring + RING * 3 * A.wcap);   // int32 variant only
    /* work counters of the group in flight: shared memory, touched by team lane 0 only */
    unsigned long long *gst = reinterpret_cast<unsigned long long *>(ring + per_team - 32);
    static_assert(SI_COUNT <= 16, "counter block");
    const int slot = blockIdx.x * (blockDim.x >> 5) * TPW + team;
    GroupState G;
    G.g = -1; G.r = G.r1 = 0; G.gbase = 0; G.N = G.E = G.par = 0; G.t0 = 0;
    bool dry = false;        // the queue has nothing left for this team
    bool gave_up = false;    // (seeded launches) the host stopped publishing anchors
    for (;;) {
        if (G.g < 0 && !dry) {                        // per team; team lane 0 only, no cross-lane operation
            int g = -1;
            if (tm.tl == 0) {
                for (;;) {
                    const int qi = atomicAdd(A.queue_head, 1);
                    g = qi < A.n_queue ? A.queue[qi] : -1;
                    if constexpr (SEEDED) {
                        /* the anchors of this group may still be on their way (poa_capi.cu: publish_seeds).  The wait
                         * is bounded: a group whose anchors do not arrive goes back to the host, which runs it again */
                        if (g >= 0 && A.seed_ready != nullptr) {
                            const int need = A.seed_rank[g];
                            const volatile int32_t *rdy = A.seed_ready;
                            unsigned long long t_start = 0;
                            while (!gave_up && *rdy <= need) {
                                unsigned long long now;
                                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                                if (t_start == 0) t_start = now;
                                if (now - t_start > 30000000000ull) gave_up = true;     // 30 s
                                else __nanosleep(20000);
                            }
                            if (*rdy <= need) { A.status[g] = ST_RETRY | (A.level << 8); A.cons_len[g] = 0; continue; }
                            __threadfence();
                        }
                    }
                    break;
                }
                if (g >= 0) {
#pragma unroll
                    for (int k = 0; k < SI_COUNT; ++k) gst[k] = 0;
                }
            }
            G.g = g;
        }
        G.g = tm.shfl(G.g, 0);
        if (G.g < 0) dry = true;
        else if (G.N == 0 && G.r == G.r1) {           // a group that was just fetched
            G.r = A.group_read_off[G.g]; G.r1 = A.group_read_off[G.g + 1];
            G.par = 0; G.E = 0;
            G.t0 = clock64();
            if (G.r1 > G.r) G.gbase = A.read_off[G.r];
        }
        if (!tm.wany(G.g >= 0)) break;
        const bool act = G.g >= 0 && G.r < G.r1;
        int rc = read_step<T, WPL, TRACE, SEEDED>(A, tm, slot, G, act, ring, ring_info, gst);
        if (G.g >= 0 && !act) rc = ST_EMPTY;          // a group without reads
        const bool fin = G.g >= 0 && rc == ST_PENDING && G.r == G.r1;
        if (tm.wany(fin)) {
            const int frc = finish_group<T>(A, tm, slot, G, fin, gst, ring);
            if (fin) rc = frc;
        }
        if (G.g >= 0 && rc != ST_PENDING) {
            if (tm.tl == 0) {
                const bool retry = rc == ST_RETRY || rc == ST_RETRY_WIDE || rc == ST_RETRY_32;
                A.status[G.g] = retry ? (rc | (A.level << 8)) : rc;
                if (rc != ST_OK) A.cons_len[G.g] = 0;
                if (!retry) {
                    gst[SI_T_BUSY] += clock64() - G.t0;
                    for (int k = 0; k < SI_COUNT; ++k)
                        if (gst[k]) atomicAdd(A.stats + k, gst[k]);
                }
            }
            G.g = -1; G.N = 0; G.r = G.r1 = 0;
        }
        tm.sync();
    }
}

/* consensus regions -> one compact buffer; one warp per group, 1 B in / 1 B out per base */
__global__ void gather_consensus_kernel(const uint8_t *__restrict__ cons, const int64_t *__restrict__ region_off,
                                        const int64_t *__restrict__ out_off, uint8_t *__restrict__ out, int64_t n_groups) {
    const int lane = threadIdx.x & 31;
    int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; g < n_groups; g += nw) {
        const int64_t n = out_off[g + 1] - out_off[g];
        const uint8_t *src = cons + region_off[g];
        uint8_t *dst = out + out_off[g];
        for (int64_t k = lane; k < n; k += 32) dst[k] = src[k];
    }
}

/* INT-pipe roofline probe: register-resident chains of the DPX op the DP is built on
 * (VIADDMNMX.S16x2), 8 independent chains per thread.  out[] only defeats dead-code removal. */
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *out, uint32_t seed, int iters) {
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x * 8 + k;
    const uint32_t c = seed | 0x00010001u, f = seed * 3u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = __viaddmax_s16x2(a[k], c, f + k);
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r ^= a[k];
    if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

cudaError_t launch_int_peak(uint32_t *out, int blocks, int iters, cudaStream_t stream) {
    int_peak_kernel<<<blocks, 256, 0, stream>>>(out, 12345u, iters);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------ */
/* host-callable launchers (used by poa_capi.cu)                                               */
/* ------------------------------------------------------------------------------------------ */

cudaError_t launch_encode(const uint8_t *ascii, uint8_t *codes, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (n + threads * 16 - 1) / (threads * 16);
    if (blocks > 148 * 16) blocks = 148 * 16;
    encode_bases_kernel<<<(unsigned)blocks, threads, 0, stream>>>(ascii, codes, n);
    return cudaGetLastError();
}

cudaError_t launch_gather(const uint8_t *cons, const int64_t *region_off, const int64_t *out_off, uint8_t *out,
                          int64_t n_groups, cudaStream_t stream) {
    if (n_groups <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (n_groups * 32 + threads - 1) / threads;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gather_consensus_kernel<<<(unsigned)blocks, threads, 0, stream>>>(cons, region_off, out_off, out, n_groups);
    return cudaGetLastError();
}

template <int T, int WPL>
static const void *variant_fn(bool trace, bool seeded) {
    if (seeded) return trace ? reinterpret_cast<const void *>(&poa_group_kernel<T, WPL, true, true>)
                             : reinterpret_cast<const void *>(&poa_group_kernel<T, WPL, false, true>);
    return trace ? reinterpret_cast<const void *>(&poa_group_kernel<T, WPL, true, false>)
                 : reinterpret_cast<const void *>(&poa_group_kernel<T, WPL, false, false>);
}

/* the instantiated kernel variants: (team size, words per lane) */
#define MPOA_FOR_VARIANTS(X) X(32, 2) X(32, 4) X(32, 8) X(32, 0)

static const void *kernel_of(int code, bool trace, bool seeded) {
#define X(T, W) if (code == variant_code(T, W)) return variant_fn<T, W>(trace, seeded);
    MPOA_FOR_VARIANTS(X)
#undef X
    return nullptr;
}

bool variant_exists(int code) { return kernel_of(code, false, false) != nullptr; }

size_t poa_smem_bytes(int code, int wcap, int warps_per_block) {
#define X(T, W) if (code == variant_code(T, W)) return ((size_t)warps_per_block * (32 / T) * team_words<T, W>(wcap) + warps_per_block * 32) * sizeof(int);
    MPOA_FOR_VARIANTS(X)
#undef X
    return 0;
}

cudaError_t launch_poa(int code, const KernelArgs &A, int n_blocks, int warps_per_block, cudaStream_t stream) {
    const size_t smem = poa_smem_bytes(code, A.wcap, warps_per_block);
    const void *fn = kernel_of(code, A.tr_aln != nullptr, A.anc_off != nullptr);   // all five trace arrays are set together
    if (!fn) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    void *args[] = {const_cast<KernelArgs *>(&A)};
    return cudaLaunchKernel(fn, dim3(n_blocks), dim3(warps_per_block * 32), args, smem, stream);
}

int poa_max_blocks_per_sm(int code, int wcap, int warps_per_block, bool seeded) {
    int nb = 0;
    const size_t smem = poa_smem_bytes(code, wcap, warps_per_block);
    const void *fn = kernel_of(code, false, seeded);
    if (!fn) return 0;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, warps_per_block * 32, smem) != cudaSuccess) return 0;
    return nb;
}

}  // namespace mpoa
