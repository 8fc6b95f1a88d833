/*
 * mandalorion_poa.h -- C ABI of the B200-native consensus library (libmandalorion_poa.so).
 *
 * This is the drop-in boundary for Mandalorion's per-isoform consensus step.  The
 * reference has no FFI: the interface it replaces is the subprocess protocol
 *
 *     abpoa -M 5 -r 0 [-S] <in.fasta>  >  <consensus.fasta>
 *
 * issued once per isoform by determine_consensus()
 * (reference utils/SpliceDefineConsensus.py:917 and :919, called from
 * defineIsoforms.py:89).  One call of mpoa_consensus_batch() replaces MANY of those
 * subprocess calls: every group is one "in.fasta" (reads in file order), every
 * consensus one ">Consensus_sequence" record.
 *
 * Conventions: plain pointers and sizes; every buffer is caller-owned; return value
 * 0 = ok, negative = MPOA_E*; no exceptions cross the ABI; a context is bound to one
 * CUDA device and is not thread-safe (one context per host thread).  Several contexts may
 * live on one GPU: their kernels take turns (mpoa_batch_run holds a per-device lock while
 * its kernels run) and share one workspace; uploads and fetches are copies only and run
 * beside another context's kernels.
 * There is no CPU path behind this ABI: mpoa_create() fails when no CUDA device is
 * usable.
 */
#ifndef MANDALORION_POA_H
#define MANDALORION_POA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPOA_ABI_VERSION 3

/* error codes (negative return values) */
#define MPOA_OK          0
#define MPOA_EINVAL     -1   /* bad argument */
#define MPOA_ENODEV     -2   /* no usable CUDA device / CUDA error at create */
#define MPOA_ECUDA      -3   /* CUDA runtime error, see mpoa_last_error() */
#define MPOA_ENOSPC     -4   /* cons_buf too small; cons_off[n_groups] holds the need */
#define MPOA_ENOMEM     -5   /* device/host allocation failed */

/* per-group status written to group_status[] */
#define MPOA_GROUP_OK       0  /* consensus produced */
#define MPOA_GROUP_EMPTY    1  /* no consensus (abpoa would have printed nothing or died):
                                  caller falls back to the first read, reference
                                  utils/SpliceDefineConsensus.py:924-925 */
#define MPOA_GROUP_TOO_BIG  2  /* no consensus because the group outgrew every device capacity
                                  (band wider than 4096 cells, workspace larger than the free
                                  memory): NOT something abpoa would have done -- the caller
                                  decides (the Python layer falls back like for EMPTY and warns) */

/* group_flags bits */
#define MPOA_FLAG_SEED   1u    /* the reference would have passed -S (median read length
                                  >= 8000, utils/SpliceDefineConsensus.py:916-919): the group is
                                  aligned window by window between minimizer anchors, see
                                  "Seeded groups" below */

/*
 * Seeded groups (abpoa -S).  abPOA's seeding code (upstream abpoa_seed.c) is not part of the
 * reference tree; what this library implements is the structure SURVEY.md A.10 records -- (k, w)
 * minimizers of consecutive reads, colinear chaining of their hits, anchors at least MIN_W apart,
 * every anchor k-mer a forced run of matches, the stretches between anchors aligned to the
 * sub-graph between the anchor nodes -- with the rules of csrc/seed.cpp for everything the summary
 * leaves open.  The CPU oracle restates the same rules independently; neither is pinned against a
 * real `abpoa -S` run (none is reachable from the build image).
 */
#define MPOA_SEED_K      19
#define MPOA_SEED_W      10
#define MPOA_SEED_MIN_W  500

/*
 * Scoring / band parameters == the abpoa command line of the reference.
 * mpoa_default_params() fills the values of "abpoa -M 5 -r 0":
 *   match 5, mismatch 4, convex gap (4,2)/(24,1), adaptive band b=10 f=0.01.
 * simd_pn_i16 / simd_pn_i32: abPOA rounds every band to whole SIMD vectors, so the
 * band (and in rare ties the result) depends on the vector width of the abpoa BUILD
 * the reference would have run: 8/4 (SSE4.1), 16/8 (AVX2, the released binary and
 * the default here), 32/16 (AVX512BW).
 */
typedef struct mpoa_params {
    int32_t match;
    int32_t mismatch;
    int32_t gap_open1, gap_ext1;
    int32_t gap_open2, gap_ext2;
    int32_t wb;
    float   wf;
    int32_t simd_pn_i16;
    int32_t simd_pn_i32;
    int32_t debug_small_caps;  /* tests only: schedule the first attempt with deliberately too small
                                  device workspaces / band capacity so the GPU retry paths run */
    int32_t reserved[5];
} mpoa_params;

/* Work accounting of one batch call (all counters are sums over the batch). */
typedef struct mpoa_stats {
    int64_t n_groups;        /* groups processed                                         */
    int64_t n_reads;         /* reads consumed                                           */
    int64_t n_alignments;    /* read-to-graph alignments (reads 2..n of every group)     */
    int64_t band_cells;      /* DP cells inside the adaptive band (the GCUPS numerator)  */
    int64_t full_cells;      /* sum of graph_rows x (qlen+1): full-matrix equivalent     */
    int64_t int_ops;         /* 17*band_cells + 3*sum((in_degree-1)*row_cells)           */
    int64_t n_align_i16;     /* alignments abPOA would have run in int16 lanes           */
    int64_t n_align_i32;     /* ... in int32 lanes                                       */
    int64_t tb_bytes;        /* traceback bytes written to HBM                           */
    int64_t n_retry_groups;  /* groups re-run with a larger device workspace             */
    double  kernel_ms;       /* device time of the POA kernel(s), CUDA events            */
    double  h2d_ms;          /* host->device copies (0 for the *_device entry point)     */
    double  d2h_ms;          /* device->host copies                                      */
    int64_t n_kernel_launches; /* kernels of this library launched by the call           */
    int64_t phase_cycles[6];  /* SM cycles summed over warps: graph-prep, DP, traceback,
                                 merge, consensus, total busy                            */
    int64_t n_seed_groups;   /* groups flagged MPOA_FLAG_SEED (the reference would pass -S) */
    int64_t n_seed_applied;  /* ... of which were aligned with minimizer seeding; the rest
                                ran the unseeded algorithm (see mpoa_seed_policy below)     */
    int64_t n_too_big_groups;/* groups that ended as MPOA_GROUP_TOO_BIG                    */
    int64_t max_band_width;  /* widest band row of the batch, cells (oracle only; 0 here)  */
    double  host_seed_ms;    /* host time of the `-S` minimizer seeding (inside h2d_ms for mpoa_batch_upload;
                                beside the kernels of the unseeded groups for the one-call entry points) */
    double  kernel_wait_ms;  /* time mpoa_batch_run waited for another context's kernel phase on the same device
                                (contexts of one process take turns on the SMs; poa.PoaPipeline) */
    int64_t reserved[2];
} mpoa_stats;

/*
 * Optional per-read trace for parity tests (any pointer may be NULL).  Nodes are named
 * by their creator: the index of the read base (relative to the first base of the
 * group) that created the node, which is independent of internal node numbering.
 */
typedef struct mpoa_trace {
    int32_t *read_score;      /* [n_reads]  best global score, 0 for the first read       */
    int32_t *read_bits;       /* [n_reads]  16 or 32: abPOA's lane width, 0 first read    */
    int64_t *read_band_cells; /* [n_reads]                                                */
    int32_t *base_aln;        /* [n_bases]  creator of the node this base was aligned to
                                            (cigar MATCH target), -1 = insertion          */
    int32_t *base_node;       /* [n_bases]  creator of the node the base was merged into
                                            (== own index when it created a node)         */
} mpoa_trace;

typedef struct mpoa_ctx mpoa_ctx;

int  mpoa_abi_version(void);
void mpoa_default_params(mpoa_params *p);

/* device_ordinal: CUDA device index (>= 0).  Returns MPOA_ENODEV without a GPU. */
int  mpoa_create(mpoa_ctx **out, int device_ordinal, const mpoa_params *p);
void mpoa_destroy(mpoa_ctx *ctx);
const char *mpoa_last_error(mpoa_ctx *ctx);

/* Kernels are launched on this stream (a cudaStream_t).  Default: a non-blocking stream the context owns, so
 * that several contexts on one GPU (mandalorion_b200.poa.PoaPipeline: one batch uploads while another one
 * computes) never serialise on the legacy stream.  Every entry point returns with its own work finished. */
int  mpoa_set_stream(mpoa_ctx *ctx, void *cuda_stream);

/* enable != 0: the next mpoa_batch_upload() also allocates the per-read trace arrays. */
int  mpoa_set_trace(mpoa_ctx *ctx, int enable);

/* INT-pipe roofline probe: measured rate (warp-wide instructions per second, whole GPU) of the
 * DPX instruction the packed DP is built on (VIADDMNMX.S16x2). */
int  mpoa_measure_int_peak(mpoa_ctx *ctx, double *warp_instr_per_sec);

/*
 * Consensus of n_groups independent read groups, HOST buffers in, HOST buffers out.
 *   group_read_off[n_groups+1]  reads of group g are [group_read_off[g], group_read_off[g+1])
 *   read_base_off[n_reads+1]    bases of read r are bases[read_base_off[r] .. read_base_off[r+1])
 *   bases                       ASCII, any case; everything except ACGTacgt is N
 *   group_flags[n_groups]       MPOA_FLAG_* or NULL
 *   cons_off[n_groups+1]  (out) consensus of group g is cons_buf[cons_off[g] .. cons_off[g+1])
 *   cons_buf / cons_cap   (out) ASCII "ACGTN", no terminator
 *   group_status[n_groups] (out) MPOA_GROUP_*
 *   stats, trace          (out) nullable
 * Reads of a group are aligned IN THE GIVEN ORDER (progressive POA is order dependent,
 * reference order comes from utils/SpliceDefineConsensus.py:884-888).  Groups of one
 * read return that read; the reference's "<= 2 reads" bypass (:911-912) is the
 * caller's decision, the library aligns whatever it is given.
 */
int  mpoa_consensus_batch(mpoa_ctx *ctx, int64_t n_groups,
                          const int64_t *group_read_off, const int64_t *read_base_off,
                          const uint8_t *bases, const uint8_t *group_flags,
                          int64_t *cons_off, uint8_t *cons_buf, int64_t cons_cap,
                          int32_t *group_status, mpoa_stats *stats, mpoa_trace *trace);

/*
 * Same job split in three so that a caller can keep the inputs resident in HBM:
 *   mpoa_batch_upload : host -> device copy of the group arrays + scheduling
 *   mpoa_batch_run    : kernels only (may be called repeatedly on the same upload)
 *   mpoa_batch_fetch  : device -> host copy of the consensus sequences
 */
int  mpoa_batch_upload(mpoa_ctx *ctx, int64_t n_groups,
                       const int64_t *group_read_off, const int64_t *read_base_off,
                       const uint8_t *bases, const uint8_t *group_flags);
int  mpoa_batch_run(mpoa_ctx *ctx, mpoa_stats *stats);
int  mpoa_batch_fetch(mpoa_ctx *ctx, int64_t *cons_off, uint8_t *cons_buf, int64_t cons_cap,
                      int32_t *group_status, mpoa_trace *trace);

/*
 * mpoa_batch_upload() for a SUBSET of the caller's groups: sel[n_sel] are ascending indices into the
 * n_groups groups described by the offset arrays; the context then holds a batch of n_sel groups,
 * numbered in sel order (group_flags is still indexed by the caller's group numbers).  The bases of
 * the selected groups are gathered straight from the caller's buffer into the pinned staging buffers
 * of the copy -- no intermediate host copy of the shard.
 */
int  mpoa_batch_upload_subset(mpoa_ctx *ctx, int64_t n_groups,
                              const int64_t *group_read_off, const int64_t *read_base_off,
                              const uint8_t *bases, const uint8_t *group_flags,
                              int64_t n_sel, const int64_t *sel);

/*
 * ONE batch over several GPUs of one box.  The reference spreads isoform groups over a fork pool
 * (defineIsoforms.py:130-153) and they never exchange data (:88-90), so the batch is cut into
 * cost-balanced shards and every shard runs on its own GPU: no collective, no peer traffic.
 *
 * mpoa_shard_plan: owner[g] in [0, n_shards) by longest-processing-time greedy over the estimated DP
 * cost of a group (sum of read lengths x expected band width); deterministic.  p NULL = defaults.
 *
 * mpoa_consensus_batch_multi: mpoa_consensus_batch() over ctxs[0..n_ctx) (distinct contexts, normally
 * one per GPU).  One host thread per context uploads its shard (mpoa_batch_upload_subset), runs it
 * and fetches it; results come back in INPUT order.  stats: n_ctx entries or NULL (per device);
 * owner: n_groups entries or NULL (the plan that was used).  On error the message is in
 * mpoa_last_error(ctxs[0]).
 */
int  mpoa_shard_plan(int64_t n_groups, const int64_t *group_read_off, const int64_t *read_base_off,
                     const uint8_t *group_flags, const mpoa_params *p, int32_t n_shards, int32_t *owner);
int  mpoa_consensus_batch_multi(mpoa_ctx *const *ctxs, int32_t n_ctx, int64_t n_groups,
                                const int64_t *group_read_off, const int64_t *read_base_off,
                                const uint8_t *bases, const uint8_t *group_flags,
                                int64_t *cons_off, uint8_t *cons_buf, int64_t cons_cap,
                                int32_t *group_status, mpoa_stats *stats, int32_t *owner);

/*
 * Orientation of every read of a group against the group's FIRST read -- what the reference
 * obtains from mappy before it writes the abpoa input (utils/SpliceDefineConsensus.py:895,
 * :900-907: mp.Aligner(seq=first, preset='map-ont'), one output record per PRIMARY hit,
 * reverse-complemented when hit.strand == -1, reads without a primary hit dropped).
 *   hit_count[n_reads]     (out) primary hits of the read: 0 (dropped), 1, or 2 (written twice)
 *   hit_strand[2*n_reads]  (out) +1 / -1 per hit, best chain first
 * Host code (minimizer sketch k=15 w=10 + colinear chaining, the seed-chain stage of minimap2's
 * map-ont; no base-level extension), n_threads worker threads over groups; needs no context
 * and no GPU.  The Python layer uses mappy itself whenever it is importable.
 */
int  mpoa_orient_batch(int64_t n_groups, const int64_t *group_read_off, const int64_t *read_base_off,
                       const uint8_t *bases, int32_t n_threads, int8_t *hit_count, int8_t *hit_strand);

#ifdef __cplusplus
}
#endif
#endif /* MANDALORION_POA_H */
