/*
 * mpoa_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE, not product).
 *
 * The oracle is a scalar restatement of what `abpoa -M 5 -r 0 <in.fasta>` (abPOA
 * v1.4.1, pinned by reference setup.sh:17-19, invoked at reference
 * utils/SpliceDefineConsensus.py:917) computes for one FASTA file.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product library never links or calls it.
 *
 * PARITY UNPINNED: abPOA's source is not under /root/reference, is not installed in
 * the build image and cannot be fetched; the reference ships no tests or golden
 * vectors for this path.  The algorithm is restated from the published design of
 * abPOA v1.4.1 ("Oracle"), pinned only by known-answer tests we
 * authored (tests/test_oracle_kat.py) and by tests/golden/ vectors produced by this
 * oracle itself.  tests/test_real_abpoa.py diffs against a real abpoa binary whenever
 * one is reachable at run time.
 */
#ifndef MPOA_ORACLE_H
#define MPOA_ORACLE_H

#include <stdint.h>
#include "../include/mandalorion_poa.h"

#ifdef __cplusplus
extern "C" {
#endif

/* switches for the details of abPOA that were restated with less than full confidence
 * (SURVEY.md Appendix A, items tagged M/L); defaults = the believed behaviour */
typedef struct mpoa_oracle_opts {
    int32_t clamp_end_to_pred;   /* 1: end_sn <= max(pred end_sn)+1 (default 1)          */
    int32_t single_argmax;       /* 1: adaptive band follows one arg-max (old abPOA), 0:
                                    left-most and right-most arg-max (default 0)          */
    int32_t hb_tie_later_wins;   /* 1: heaviest-bundling tie -> later edge if its score
                                    is >= (default 1)                                     */
    int32_t n_threads;           /* worker threads over groups, <=1: serial               */
    /* `abpoa -S` (groups flagged MPOA_FLAG_SEED): minimizer seeding constants, 0 = default.
     * Restated with LOW confidence (upstream abpoa_seed.c is not available), see abpoa_oracle.cpp */
    int32_t seed_k;              /* minimizer k-mer (default 19)                          */
    int32_t seed_w;              /* minimizer window (default 10)                         */
    int32_t seed_min_w;          /* minimum distance between kept anchors (default 500)   */
    int32_t honour_seed_flag;    /* 1 (default): flagged groups run the seeded path; 0: the
                                    flag is ignored (the unseeded algorithm for everyone)  */
} mpoa_oracle_opts;

void mpoa_oracle_default_opts(mpoa_oracle_opts *o);

/* Same contract as mpoa_consensus_batch() (include/mandalorion_poa.h) minus the context. */
int mpoa_oracle_consensus_batch(const mpoa_params *p, const mpoa_oracle_opts *o,
                                int64_t n_groups,
                                const int64_t *group_read_off, const int64_t *read_base_off,
                                const uint8_t *bases, const uint8_t *group_flags,
                                int64_t *cons_off, uint8_t *cons_buf, int64_t cons_cap,
                                int32_t *group_status, mpoa_stats *stats, mpoa_trace *trace);

#ifdef __cplusplus
}
#endif
#endif
