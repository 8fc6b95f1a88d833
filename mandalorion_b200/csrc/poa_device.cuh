/*
 * poa_device.cuh -- shared device-side declarations of the B200 consensus kernels.
 *
 * What is computed: the partial-order-alignment consensus that Mandalorion obtains from
 * `abpoa -M 5 -r 0 in.fasta` (reference utils/SpliceDefineConsensus.py:917), for many
 * independent read groups at once.  How it is laid out is ours:
 *
 *   - one WARP owns one read group from the first read to the consensus (reads of a group are
 *     sequentially dependent, groups are independent); warps pull groups from a global queue
 *     sorted by descending cost;
 *   - the POA graph lives in ROW space: nodes are stored in a topological order in which
 *     aligned-node groups are contiguous, adjacency is CSR with edges in first-creation order.
 *     After every read the graph is re-emitted (double buffered) with the new nodes merged in,
 *     so there is never a BFS re-sort; abPOA's results do not depend on which valid
 *     topological order is used (every tie-break iterates edge lists, never row numbers);
 *   - DP rows H/E1/E2 live in a per-warp shared-memory ring (int16x2 packed, DPX
 *     VIADDMNMX/VIMNMX3, for alignments abPOA itself would run in int16 lanes; int32 otherwise);
 *     every row is also streamed to HBM (the traceback matrices) with vectorised stores;
 *   - traceback compares the stored values exactly like abPOA's cg_backtrack and recomputes the
 *     insertion scores F of a row only when a step needs them.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <climits>

namespace mpoa {

constexpr unsigned FULL = 0xffffffffu;
constexpr int NEG = -(1 << 28);      // -inf surrogate in int32 arithmetic
constexpr int NEG16 = -28000;        // -inf surrogate / floor of the packed int16 path (per-lane rebased scores, see poa_dp.cuh)
constexpr int LOW16 = -20000;        // a row whose best rebased score falls below this leaves the packed path (ST_RETRY_32)
constexpr int RING = 4;              // rows kept in the shared-memory ring of the int32 variant (occupancy: 3*wcap ints per row)
constexpr int WARPS_PER_BLOCK = 4;

enum GroupStatus : int { ST_OK = 0, ST_EMPTY = 1, ST_RETRY = 2, ST_PENDING = 3, ST_RETRY_WIDE = 4, ST_RETRY_32 = 5, ST_TOO_BIG = 6 };

struct DevParams {
    int match, mismatch, o1, e1, o2, e2, oe1, oe2, wb;
    float wf;
    int pn16, pn32;
};

/* int16x2 constants of the packed DP, built once on the host (they sit in the constant bank and
 * are used directly as instruction operands).  (a,b) = a in the low half, b in the high half. */
struct Packed16 {
    uint32_t neg2;             // (NEG16, NEG16)
    uint32_t nee;              // (-e1,-e2): per-cell words holding (F1,F2)
    uint32_t noe1, noe2;       // (-oe1,-oe1), (-oe2,-oe2): per-word of two cells
    uint32_t ne1, ne2;         // (-e1,-e1), (-e2,-e2)
    uint32_t tdec[16];         // t * (-e1,-e2): decay inside a lane at cell t
    uint32_t mhop;             // (-match*CPL, -match*CPL): change of score frame between neighbouring lanes
    uint32_t kc_r, kc_l;       // per lane position: CPL cells and match*CPL of score frame in a row-maximum key (right / left)
};

/* byte offsets of the arrays inside one warp's HBM workspace ("slot") */
struct SlotLayout {
    uint32_t ncap, ecap, qcap;
    uint64_t tbcap;                                    // bytes of the traceback area
    /* [0], [1]: the double-buffered graph; in_off[2] / in_row[2], remain[1], meta[1], qmap[1]: the
     * sub-graph view of a seeded window (zero-sized unless the launch is a seeded one) */
    uint64_t base[2], sib[2], creator[2], in_off[3], in_row[3], out_off[2], out_row[2], out_w[2];
    uint64_t remain[2], meta[2], rowinfo, rowtb, rowbest, qmap[2];
    uint64_t prevrow;                                  // row of every base of the previous read (seeded launches)
    uint64_t qprof;                                    // query profile of the read being aligned: 4 bases x qprof_stride(qcap) words
    uint64_t pv, pkey, pnew, psib, nin, nout;          // per query position
    uint64_t cnt, addin, addout, grow, srcof;          // per row
    uint64_t tb;
    uint64_t slot_bytes;
};

struct KernelArgs {
    const uint8_t *codes;            // nt4 codes of all bases of the batch
    const int64_t *read_off;         // [n_reads+1]
    const int64_t *group_read_off;   // [n_groups+1]
    const int32_t *queue;            // group indices to process, heaviest first
    int n_queue;
    int *queue_head;
    uint8_t *ws;                     // n_slots * L.slot_bytes
    SlotLayout L;
    uint8_t *cons;                   // consensus bytes, region of group g at cons_off[g]
    const int64_t *cons_off;
    int32_t *cons_len;
    int32_t *status;
    unsigned long long *stats;       // see StatIdx
    int32_t *tr_score, *tr_bits;     // optional trace (NULL when off)
    long long *tr_cells;
    int32_t *tr_aln, *tr_node;
    /* `abpoa -S` launches: read r owns the slots anc[anc_off[r] ...]: (number of anchors, 0), then the anchors
     * (start in the previous read, start in this read).  The host may still be computing anchors when the launch
     * starts: group g may be started once *seed_ready > seed_rank[g] (seed_ready == NULL: everything is there);
     * the host copies a chunk of anchors and then the counter on one stream, the kernel reads both past L1. */
    const int32_t *anc_off;          // [n_reads]
    const int2 *anc;
    const int32_t *seed_rank;        // [n_groups]
    const int32_t *seed_ready;
    int seed_k;
    DevParams P;
    Packed16 K;
    uint32_t tbcap_words;            // traceback area of a slot in 32-bit words (capped to 32-bit offsets)
    int wcap;                        // cells per ring row
    int level;                       // host launch level, reported back with the retry codes
};

/* words between the base rows of the query profile (one word = the scores of two cells) */
__host__ __device__ inline uint32_t qprof_stride(uint32_t qcap) { return ((qcap >> 1) + 24u) & ~7u; }

enum StatIdx { SI_CELLS = 0, SI_INTOPS, SI_FULL, SI_ALN, SI_ALN16, SI_ALN32, SI_TB,
               SI_T_PREP, SI_T_DP, SI_T_TB, SI_T_MERGE, SI_T_CONS, SI_T_BUSY, SI_COUNT };

/*
 * A TEAM of T lanes (T = 16: two teams per warp, T = 32: the whole warp) owns one read group.
 * The two teams of a warp run in LOCKSTEP: every loop that contains a cross-lane operation runs
 * for the larger of the two teams' trip counts (wmax) and every branch around one is taken by
 * both or by neither (wany / wall); a team that has nothing to do in an iteration is switched off
 * by a predicate, never by control flow.  So the whole warp is converged at every shuffle, the
 * shuffles carry the constant full mask (a run-time member mask costs a MATCH/REDUX/VOTE check per
 * shuffle) and one instruction stream serves two groups.  Branches and loops WITHOUT cross-lane
 * operations (per-lane edge lists, re-binding a lane, rare fix-ups) may differ between the teams.
 */
template <int T>
struct Team {
    static_assert(T == 16 || T == 32, "team size");
    static constexpr int SIZE = T;
    int tl;           // lane inside the team
    int base;         // warp lane of team lane 0
    __device__ __forceinline__ explicit Team(int warp_lane) {
        /* warp_lane (0..31) comes through shared memory (poa_group_kernel), so that it stays in a register */
        tl = T == 32 ? warp_lane : (warp_lane & (T - 1));
        base = T == 32 ? 0 : (warp_lane & ~(T - 1));
    }
    template <typename V> __device__ __forceinline__ V shfl(V v, int src) const { return __shfl_sync(FULL, v, src, T); }
    template <typename V> __device__ __forceinline__ V shfl_up(V v, int d) const { return __shfl_up_sync(FULL, v, d, T); }
    template <typename V> __device__ __forceinline__ V shfl_down(V v, int d) const { return __shfl_down_sync(FULL, v, d, T); }
    template <typename V> __device__ __forceinline__ V shfl_xor(V v, int d) const { return __shfl_xor_sync(FULL, v, d, T); }
    /* bit l = predicate of team lane l */
    __device__ __forceinline__ unsigned ballot(bool p) const {
        if constexpr (T == 32) return __ballot_sync(FULL, p);
        else return (__ballot_sync(FULL, p) >> base) & 0xffffu;
    }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    /* maximum over the team */
    __device__ __forceinline__ int rmax(int v) const {
        if constexpr (T == 32) return __reduce_max_sync(FULL, v);
        else {
            const int a = __reduce_max_sync(FULL, base ? INT_MIN : v), b = __reduce_max_sync(FULL, base ? v : INT_MIN);
            return base ? b : a;
        }
    }
    /* value of team lane `src`; for a whole warp the REDUX form tells the compiler it is uniform */
    __device__ __forceinline__ uint32_t pick(uint32_t v, int src) const {
        if constexpr (T == 32) return __reduce_or_sync(FULL, tl == src ? v : 0u);
        else return __shfl_sync(FULL, v, src, T);
    }
    __device__ __forceinline__ int uniform(int v) const {
        if constexpr (T == 32) return __reduce_max_sync(FULL, v);
        else return v;
    }
    /* warp-wide agreement between the teams (for T = 32 the argument already is warp-uniform) */
    __device__ __forceinline__ int wmax(int v) const {
        if constexpr (T == 32) return v;
        else return __reduce_max_sync(FULL, v);
    }
    __device__ __forceinline__ bool wany(bool p) const {
        if constexpr (T == 32) return p;
        else return __any_sync(FULL, p);
    }
    /* any lane of the warp (the argument differs from lane to lane) */
    __device__ __forceinline__ bool any_lane(bool p) const { return __any_sync(FULL, p); }
    __device__ __forceinline__ bool wall(bool p) const {
        if constexpr (T == 32) return p;
        else return __all_sync(FULL, p);
    }
};

/* kernel variants: team size T and words per lane WPL of the packed int16x2 DP (band <= T*2*WPL
 * cells); WPL = 0: int32 lanes, one cell per lane, any band width in chunks of 32 (whole warp) */
struct Variant { int T, WPL; };
__host__ __device__ constexpr int variant_code(int T, int WPL) { return WPL == 0 ? 0 : T * 100 + WPL; }

}  // namespace mpoa
