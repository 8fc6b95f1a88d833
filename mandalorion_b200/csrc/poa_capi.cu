/*
 * poa_capi.cu -- host side of libmandalorion_poa.so: the C ABI declared in
 * include/mandalorion_poa.h.  Batching, length binning, device workspace, launch and retry
 * logic around the kernels of poa_kernels.cu.  There is no CPU compute path here: every
 * consensus is produced by poa_group_kernel.
 *
 * One call of mpoa_consensus_batch() stands for many `abpoa -M 5 -r 0 in.fasta` processes
 * of the reference (utils/SpliceDefineConsensus.py:917), one per group.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mandalorion_poa.h"
#include "poa_device.cuh"

namespace mpoa {
cudaError_t launch_encode(const uint8_t *ascii, uint8_t *codes, int64_t n, cudaStream_t stream);
cudaError_t launch_poa(int code, const KernelArgs &A, int n_blocks, int warps_per_block, cudaStream_t stream);
int poa_max_blocks_per_sm(int code, int wcap, int warps_per_block, bool seeded);
void seed_batch(const int64_t *gro, const int64_t *rbo, const uint8_t *bases, const int64_t *src_off,
                const int32_t *order, int64_t n_order, int64_t chunk_size, std::atomic<int64_t> *done,
                int k, int w, int min_gap, int n_threads, const int32_t *anc_off, int32_t *anc);
size_t poa_smem_bytes(int code, int wcap, int warps_per_block);
bool variant_exists(int code);
cudaError_t launch_int_peak(uint32_t *out, int blocks, int iters, cudaStream_t stream);
cudaError_t launch_gather(const uint8_t *cons, const int64_t *region_off, const int64_t *out_off, uint8_t *out,
                          int64_t n_groups, cudaStream_t stream);
}  // namespace mpoa

using namespace mpoa;

struct GroupInfo {
    int32_t n_reads, maxlen, minlen;
    int64_t sumlen;
    double cost;
    int32_t wneed;     // expected band width in cells
    int8_t lanes16;    // 1: the packed int16x2 kernels may run the group (cleared by ST_RETRY_32: junk reads, exotic parameters)
    int8_t level;      // index into kLevels (kernel variant + band capacity) of the next launch
    int8_t attempt;    // workspace-capacity escalation
};

/* grow-only device buffer: batches of similar size reuse their allocations */
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

/*
 * Host -> device copy of the bases through pinned staging buffers, filled by a few host threads.
 * The caller's buffer is ordinary (pageable) memory in every real integration (numpy arrays, Python
 * bytes): the driver's own pageable path copies at about 10 GB/s (measured: 1.66 GB in 155 ms); a few
 * threads that memcpy into pinned chunks and issue the chunks on their own streams reach 35-46 GB/s.  The same path
 * gathers a SUBSET of the groups (segments of the caller's buffer) without an intermediate copy.
 */
struct Seg { const uint8_t *src; int64_t dst, len; };   // in destination order, dst = prefix sum of len

struct Stager {
    static constexpr int K = 8;          // most copy threads (MPOA_STAGE_THREADS, default 6)
    static constexpr size_t CH = 8u << 20;
    cudaStream_t st[K] = {};
    uint8_t *pin[K][2] = {};
    cudaEvent_t ev[K][2] = {};
    bool ready = false;
    cudaError_t init() {
        if (ready) return cudaSuccess;
        for (int w = 0; w < K; ++w) {
            cudaError_t e = cudaStreamCreateWithFlags(&st[w], cudaStreamNonBlocking);
            for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
                e = cudaHostAlloc((void **)&pin[w][b], CH, cudaHostAllocDefault);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[w][b], cudaEventDisableTiming);
            }
            if (e != cudaSuccess) return e;
        }
        ready = true;
        return cudaSuccess;
    }
    void release() {
        for (int w = 0; w < K; ++w) {
            for (int b = 0; b < 2; ++b) {
                if (pin[w][b]) cudaFreeHost(pin[w][b]);
                if (ev[w][b]) cudaEventDestroy(ev[w][b]);
                pin[w][b] = nullptr; ev[w][b] = nullptr;
            }
            if (st[w]) cudaStreamDestroy(st[w]);
            st[w] = nullptr;
        }
        ready = false;
    }
};

/* expected band width of a group in cells: 2w+1 + length spread + two SIMD vectors of rounding.  Measured
 * with the oracle on every named config the widest row of a group is 2w+1 + 27..65 cells (median 43), so
 * this over-estimates by a median of 24 cells -- on purpose.  A tighter estimate (2w+1 + 48) moves half of
 * cfg2's 256-cell work into the 128-cell kernel (-2.7 % kernel time), but the one group in ~10^4 that then
 * outgrows its level is re-run ALONE in a later round, and one 50-read 4-kb group takes ~0.3 s on its single
 * warp: the step gets slower, not faster (measured: 749 ms instead of 470 ms for 16 384 groups). */
static int band_need(const mpoa_params &p, int maxlen, int minlen, bool seeded) {
    const int w = p.wb + (int)(p.wf * (float)maxlen);
    int need = 2 * w + 1 + (maxlen - minlen) + 2 * p.simd_pn_i16;
    need = std::min(need, maxlen + 1 + 2 * p.simd_pn_i16);
    if (seeded) {
        /* `abpoa -S`: the band belongs to a window between two anchors (>= MPOA_SEED_MIN_W apart, rarely
         * more than a few of those), not to the whole read; a read that shares no anchor with its
         * predecessor outgrows this and is re-run wider (ST_RETRY_WIDE) */
        const int wl = std::min<int>(maxlen, 4 * MPOA_SEED_MIN_W);
        const int ww = p.wb + (int)(p.wf * (float)wl);
        need = std::min(need, 2 * ww + 1 + 3 * p.simd_pn_i16);
    }
    return need;
}

/* launch levels: kernel variant (team size T, words per lane WPL of the packed int16x2 DP; WPL 0 =
 * int32 lanes) and the band capacity in cells.  A group escalates along this list when its band
 * or its scores outgrow the level it ran at.  The int32 ring (RING x 3 x wcap ints per warp) limits
 * the widest band to 4096 cells; wider groups are reported as MPOA_GROUP_TOO_BIG. */
struct Level { int T, WPL, wcap; };
static const Level kLevels[] = {{32, 2, 128}, {32, 4, 256}, {32, 8, 512},
                                {32, 0, 256}, {32, 0, 512}, {32, 0, 1024}, {32, 0, 2048}, {32, 0, 4096}};
static const int kNumLevels = (int)(sizeof(kLevels) / sizeof(kLevels[0]));

/* tuning aid: bit l set = packed level l may be scheduled (default: all) */
static unsigned level_mask_env() {
    if (const char *ml = getenv("MPOA_LEVELS")) return (unsigned)strtoul(ml, nullptr, 0) | ~0x7u;
    return ~0u;
}
/* narrowest level that holds a band of wneed cells */
static int pick_level(const mpoa_params &p, unsigned level_mask, bool lanes16, int wneed) {
    for (int l = 0; l < kNumLevels; ++l) {
        if (!((level_mask >> l) & 1u)) continue;
        if (kLevels[l].WPL != 0 && !lanes16) continue;
        /* a lane wider than abPOA's vector length may start up to (2*WPL - pn) cells before the band */
        const int slack = std::max(0, 2 * kLevels[l].WPL - p.simd_pn_i32);
        if (kLevels[l].wcap >= wneed + slack) return l;
    }
    return kNumLevels - 1;
}

/*
 * What the contexts of one process share per device: the kernels of two contexts never run side by side
 * (different kernel variants resident on the same SMs were measured at 1.5x the time of the same work back to
 * back -- they evict each other from the instruction cache), so the kernel phase of mpoa_batch_run holds `mu`,
 * and the workspace of the resident teams (tens of GB) exists once per device, not once per context.  The
 * copies and host passes of one context's batch run beside the kernels of another (poa.PoaPipeline).
 */
struct DevShared {
    std::mutex mu;
    std::mutex up_mu;          // uploads of one device go one after the other: the first batch's copy gets the whole link and
                               // its kernels start while the second batch is still copying (instead of both arriving late)
    uint8_t *ws = nullptr;
    size_t ws_bytes = 0;
    int users = 0;
};
static DevShared g_dev[64];
static std::mutex g_dev_users;
static std::mutex g_seed_mu;                   // host-side `-S` seeding: one batch at a time (it uses all host threads)

struct mpoa_ctx {
    int dev = 0;
    int n_sm = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;         // non-blocking: contexts of one process never serialise on the legacy stream
    mpoa_params params;
    std::string err;
    int want_trace = 0;
    /* uploaded batch */
    int64_t n_groups = 0, n_reads = 0, n_bases = 0;
    std::vector<int64_t> h_gro, h_rbo;
    std::vector<int64_t> h_src;                // subset upload: start of every read in the caller's buffer
    Stager stager;
    std::vector<uint8_t> h_flags;              // MPOA_FLAG_* per group (empty: none set)
    std::vector<GroupInfo> ginfo;
    uint8_t *d_codes = nullptr;
    int64_t *d_rbo = nullptr, *d_gro = nullptr, *d_region_off = nullptr;
    uint8_t *d_cons = nullptr;
    int32_t *d_cons_len = nullptr, *d_status = nullptr, *d_queue = nullptr;
    int *d_queue_head = nullptr;
    unsigned long long *d_stats = nullptr;
    int32_t *d_tr_score = nullptr, *d_tr_bits = nullptr, *d_tr_aln = nullptr, *d_tr_node = nullptr;
    long long *d_tr_cells = nullptr;
    DevBuf b_ascii, b_codes, b_rbo, b_gro, b_region, b_len, b_status, b_queue, b_out_off, b_out;
    DevBuf b_tr_score, b_tr_bits, b_tr_cells, b_tr_aln, b_tr_node, b_anc_off, b_anc, b_seed_rank, b_seed_ready;
    int32_t *d_anc_off = nullptr, *d_seed_rank = nullptr, *d_seed_ready = nullptr;
    int2 *d_anc = nullptr;
    int64_t n_seed_groups = 0;
    /* anchors of the `abpoa -S` groups: computed inside the upload, or -- one-call entry points, where the
     * caller's buffer stays valid -- by host threads that run beside the kernels.  The groups are seeded in the
     * order the launches ask for them (seed_order); the arena (h_anc, two ints per slot, region of read r at slot
     * h_anc_off[r]) is laid out in that order, so that a finished chunk of groups is one contiguous copy; the
     * device learns of it through the counter d_seed_ready (publish_seeds). */
    std::vector<int32_t> h_anc_off, h_anc, seed_order, seed_rank, ready_vals;
    std::vector<int64_t> chunk_slot;           // first arena slot of every chunk (+ the end)
    int64_t seed_chunk = 0, n_seed_chunks = 0;
    std::unique_ptr<std::atomic<int64_t>[]> seed_done;
    cudaStream_t pub = nullptr;                // the anchor copies: beside the running kernels
    std::thread seeder;
    bool seed_pending = false;
    double seed_ms = 0, seed_wait_ms = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t side = nullptr, side2 = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr, join2_ev = nullptr;
    double h2d_ms = 0;
    uint8_t *d_ascii = nullptr;
    bool need_encode = false;                  // the uploaded bases are still ASCII
    std::vector<int64_t> h_cons_off;           // consensus offsets of the last run (compact buffer b_out)
    bool ran = false;
    std::vector<int32_t> h_status;
    mpoa_stats last;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                         \
            return _e == cudaErrorMemoryAllocation ? MPOA_ENOMEM : MPOA_ECUDA;                     \
        }                                                                                          \
    } while (0)

static void free_batch(mpoa_ctx *ctx) {   // forgets the uploaded batch; device buffers are kept for reuse
    if (ctx->seeder.joinable()) ctx->seeder.join();
    ctx->seed_pending = false;
    ctx->d_codes = nullptr; ctx->d_rbo = ctx->d_gro = ctx->d_region_off = nullptr; ctx->d_cons = nullptr;
    ctx->d_cons_len = ctx->d_status = ctx->d_queue = nullptr;
    ctx->d_tr_score = ctx->d_tr_bits = ctx->d_tr_aln = ctx->d_tr_node = nullptr; ctx->d_tr_cells = nullptr;
    ctx->n_groups = ctx->n_reads = ctx->n_bases = 0;
    ctx->d_anc_off = nullptr; ctx->d_anc = nullptr; ctx->d_seed_rank = ctx->d_seed_ready = nullptr; ctx->n_seed_groups = 0;
    ctx->ran = false;
}

static void release_buffers(mpoa_ctx *ctx) {
    for (DevBuf *b : {&ctx->b_ascii, &ctx->b_codes, &ctx->b_rbo, &ctx->b_gro, &ctx->b_region, &ctx->b_len, &ctx->b_status,
                      &ctx->b_queue, &ctx->b_out_off, &ctx->b_out, &ctx->b_tr_score, &ctx->b_tr_bits, &ctx->b_tr_cells,
                      &ctx->b_tr_aln, &ctx->b_tr_node, &ctx->b_anc_off, &ctx->b_anc, &ctx->b_seed_rank, &ctx->b_seed_ready})
        b->release();
}

extern "C" int mpoa_abi_version(void) { return MPOA_ABI_VERSION; }

extern "C" void mpoa_default_params(mpoa_params *p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->match = 5; p->mismatch = 4;
    p->gap_open1 = 4; p->gap_ext1 = 2; p->gap_open2 = 24; p->gap_ext2 = 1;
    p->wb = 10; p->wf = 0.01f;
    p->simd_pn_i16 = 16; p->simd_pn_i32 = 8;
}

extern "C" int mpoa_create(mpoa_ctx **out, int device_ordinal, const mpoa_params *p) {
    if (!out) return MPOA_EINVAL;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return MPOA_ENODEV;
    if (device_ordinal < 0 || device_ordinal >= n_dev || device_ordinal >= 64) return MPOA_ENODEV;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) return MPOA_ENODEV;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) return MPOA_ENODEV;
    if (prop.major < 10) return MPOA_ENODEV;  // sm_100a code only
    mpoa_ctx *ctx = new mpoa_ctx();
    ctx->dev = device_ordinal;
    ctx->n_sm = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    if (p) ctx->params = *p; else mpoa_default_params(&ctx->params);
    if (ctx->params.simd_pn_i16 <= 0) ctx->params.simd_pn_i16 = 16;
    if (ctx->params.simd_pn_i32 <= 0) ctx->params.simd_pn_i32 = 8;
    if (cudaMalloc(&ctx->d_queue_head, 16 * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&ctx->d_stats, SI_COUNT * sizeof(unsigned long long)) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return MPOA_ENODEV;
    }
    ctx->stream = ctx->own_stream;
    { std::lock_guard<std::mutex> g(g_dev_users); ++g_dev[ctx->dev].users; }
    *out = ctx;
    return MPOA_OK;
}

extern "C" void mpoa_destroy(mpoa_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->dev);
    free_batch(ctx);
    release_buffers(ctx);
    {
        std::lock_guard<std::mutex> g(g_dev_users);
        DevShared &D = g_dev[ctx->dev];
        if (--D.users == 0) {                         // the last context of this device takes the workspace with it
            std::lock_guard<std::mutex> k(D.mu);
            cudaFree(D.ws);
            D.ws = nullptr; D.ws_bytes = 0;
        }
    }
    cudaFree(ctx->d_queue_head);
    cudaFree(ctx->d_stats);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
    if (ctx->join2_ev) cudaEventDestroy(ctx->join2_ev);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->side2) cudaStreamDestroy(ctx->side2);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->pub) cudaStreamDestroy(ctx->pub);
    ctx->stager.release();
    delete ctx;
}

extern "C" const char *mpoa_last_error(mpoa_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int mpoa_set_stream(mpoa_ctx *ctx, void *cuda_stream) {
    if (!ctx) return MPOA_EINVAL;
    ctx->stream = (cudaStream_t)cuda_stream;
    return MPOA_OK;
}

extern "C" int mpoa_set_trace(mpoa_ctx *ctx, int enable) {
    if (!ctx) return MPOA_EINVAL;
    ctx->want_trace = enable ? 1 : 0;
    return MPOA_OK;
}

/* Measured throughput of the DPX instruction the packed DP is built on (VIADDMNMX.S16x2), in
 * warp-wide instructions per second over the whole GPU: the INT-pipe roofline denominator. */
extern "C" int mpoa_measure_int_peak(mpoa_ctx *ctx, double *warp_instr_per_sec) {
    if (!ctx || !warp_instr_per_sec) return MPOA_EINVAL;
    CK(cudaSetDevice(ctx->dev));
    const int blocks = ctx->n_sm * 8, iters = 20000;
    uint32_t *d_out = nullptr;
    CK(cudaMalloc(&d_out, (size_t)blocks * 256 * sizeof(uint32_t)));
    CK(launch_int_peak(d_out, blocks, 1000, ctx->stream));   // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        CK(launch_int_peak(d_out, blocks, iters, ctx->stream));
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double instr = (double)blocks * 8 /* warps */ * 8 /* chains */ * (double)iters;
        best = std::max(best, instr / (ms * 1e-3));
    }
    cudaFree(d_out);
    *warp_instr_per_sec = best;
    return MPOA_OK;
}

/* one worker: bytes [lo, hi) of the destination */
static cudaError_t stage_range(Stager &sg, int w, int dev, uint8_t *d_dst, const std::vector<Seg> &segs, int64_t lo, int64_t hi) {
    cudaError_t e = cudaSetDevice(dev);
    if (e != cudaSuccess) return e;
    size_t s = (size_t)(std::upper_bound(segs.begin(), segs.end(), lo, [](int64_t v, const Seg &x) { return v < x.dst; }) - segs.begin());
    s = s > 0 ? s - 1 : 0;
    int it = 0;
    for (int64_t pos = lo; pos < hi; ++it) {
        const int b = it & 1;
        const int64_t n = std::min<int64_t>((int64_t)Stager::CH, hi - pos);
        if (it >= 2 && (e = cudaEventSynchronize(sg.ev[w][b])) != cudaSuccess) return e;
        int64_t done = 0;
        while (done < n) {
            while (s + 1 < segs.size() && segs[s + 1].dst <= pos + done) ++s;
            const Seg &x = segs[s];
            const int64_t in_seg = pos + done - x.dst;
            const int64_t take = std::min<int64_t>(n - done, x.len - in_seg);
            std::memcpy(sg.pin[w][b] + done, x.src + in_seg, (size_t)take);
            done += take;
        }
        if ((e = cudaMemcpyAsync(d_dst + pos, sg.pin[w][b], (size_t)n, cudaMemcpyHostToDevice, sg.st[w])) != cudaSuccess) return e;
        if ((e = cudaEventRecord(sg.ev[w][b], sg.st[w])) != cudaSuccess) return e;
        pos += n;
    }
    return cudaStreamSynchronize(sg.st[w]);
}

static bool is_pageable(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

/*
 * Upload of the whole batch (sel == nullptr) or of the groups sel[0..n_sel) (ascending indices into
 * the caller's arrays): the context then holds a batch of n_sel groups numbered in sel order.
 */
static int upload_impl(mpoa_ctx *ctx, int64_t n_all, const int64_t *group_read_off, const int64_t *read_base_off,
                       const uint8_t *bases, const uint8_t *group_flags, int64_t n_sel, const int64_t *sel,
                       bool defer_seed = false) {
    if (!ctx || n_all < 0 || (sel && n_sel < 0)) return MPOA_EINVAL;
    if (n_all > 0 && (!group_read_off || !read_base_off)) { ctx->err = "null offsets"; return MPOA_EINVAL; }
    CK(cudaSetDevice(ctx->dev));
    free_batch(ctx);
    const int64_t n_groups = sel ? n_sel : n_all;
    if (n_groups == 0) return MPOA_OK;
    std::lock_guard<std::mutex> one_upload_per_device(g_dev[ctx->dev].up_mu);
    if (n_all > INT_MAX / 2) { ctx->err = "too many groups"; return MPOA_EINVAL; }
    const int64_t n_reads_all = group_read_off[n_all];
    if (group_read_off[0] != 0 || n_reads_all < 0) { ctx->err = "group_read_off must start at 0"; return MPOA_EINVAL; }
    const int64_t n_bases_all = n_reads_all > 0 ? read_base_off[n_reads_all] : 0;
    if (n_reads_all > 0 && (read_base_off[0] != 0 || n_bases_all < 0 || (n_bases_all > 0 && !bases))) {
        ctx->err = "bad read_base_off / bases";
        return MPOA_EINVAL;
    }
    const auto t_start = std::chrono::steady_clock::now();

    /* the batch's own offset arrays (a subset is renumbered), validated as they are built */
    ctx->h_gro.assign(n_groups + 1, 0);
    ctx->h_src.clear();
    std::vector<Seg> segs;
    if (!sel) {
        ctx->h_gro.assign(group_read_off, group_read_off + n_groups + 1);
        ctx->h_rbo.assign(read_base_off, read_base_off + n_reads_all + 1);
        for (int64_t g = 0; g < n_groups; ++g)
            if (ctx->h_gro[g + 1] < ctx->h_gro[g] || ctx->h_gro[g + 1] > n_reads_all) { ctx->err = "group_read_off not monotone"; return MPOA_EINVAL; }
        for (int64_t r = 0; r < n_reads_all; ++r) {
            const int64_t len = ctx->h_rbo[r + 1] - ctx->h_rbo[r];
            if (len < 0 || len > (1 << 26)) { ctx->err = "bad read length"; return MPOA_EINVAL; }
        }
        if (group_flags) ctx->h_flags.assign(group_flags, group_flags + n_groups); else ctx->h_flags.clear();
        if (n_bases_all > 0) segs.push_back(Seg{bases, 0, n_bases_all});
    } else {
        ctx->h_rbo.assign(1, 0);
        if (group_flags) ctx->h_flags.resize(n_groups); else ctx->h_flags.clear();
        int64_t prev = -1;
        for (int64_t k = 0; k < n_groups; ++k) {
            const int64_t g = sel[k];
            if (g <= prev || g >= n_all) { ctx->err = "subset indices must be ascending and in range"; return MPOA_EINVAL; }
            prev = g;
            const int64_t r0 = group_read_off[g], r1 = group_read_off[g + 1];
            if (r0 < 0 || r1 < r0 || r1 > n_reads_all) { ctx->err = "group_read_off not monotone"; return MPOA_EINVAL; }
            ctx->h_gro[k + 1] = ctx->h_gro[k] + (r1 - r0);
            for (int64_t r = r0; r < r1; ++r) {
                const int64_t len = read_base_off[r + 1] - read_base_off[r];
                if (len < 0 || len > (1 << 26) || read_base_off[r + 1] > n_bases_all) { ctx->err = "bad read length"; return MPOA_EINVAL; }
                ctx->h_src.push_back(read_base_off[r]);
                ctx->h_rbo.push_back(ctx->h_rbo.back() + len);
            }
            if (group_flags) ctx->h_flags[k] = group_flags[g];
            const int64_t b0 = read_base_off[r0], b1 = read_base_off[r1];
            if (b1 > b0) {
                if (!segs.empty() && segs.back().src + segs.back().len == bases + b0) segs.back().len += b1 - b0;
                else segs.push_back(Seg{bases + b0, segs.empty() ? 0 : segs.back().dst + segs.back().len, b1 - b0});
            }
        }
    }
    const int64_t n_reads = ctx->h_gro[n_groups];
    const int64_t n_bases = ctx->h_rbo[n_reads];

    /* the one big copy (the bases) starts first: the host-side passes below run while it is in flight */
    const size_t nb = (size_t)std::max<int64_t>(n_bases, 16);
    CK(ctx->b_ascii.ensure(nb)); CK(ctx->b_codes.ensure(nb));
    uint8_t *d_ascii = (uint8_t *)ctx->b_ascii.p;
    std::vector<std::thread> copiers;
    cudaError_t copy_err[Stager::K] = {};
    const bool staged = n_bases > 0 && (sel != nullptr || (n_bases >= (int64_t)(4u << 20) && is_pageable(bases)));
    if (staged) {
        CK(ctx->stager.init());
        CK(cudaStreamSynchronize(ctx->stream));          // earlier work on this buffer (it held the last batch's consensi)
        int want = 6;
        if (const char *e = getenv("MPOA_STAGE_THREADS")) want = std::max(1, std::min((int)Stager::K, atoi(e)));
        const int nw = (int)std::min<int64_t>(want, (n_bases + (int64_t)Stager::CH - 1) / (int64_t)Stager::CH);
        for (int w = 0; w < nw; ++w) {
            const int64_t lo = n_bases * w / nw / 512 * 512, hi = w + 1 == nw ? n_bases : n_bases * (w + 1) / nw / 512 * 512;
            copiers.emplace_back([&, w, lo, hi]() { copy_err[w] = stage_range(ctx->stager, w, ctx->dev, d_ascii, segs, lo, hi); });
        }
    } else if (n_bases > 0) CK(cudaMemcpyAsync(d_ascii, bases, n_bases, cudaMemcpyHostToDevice, ctx->stream));
    auto join_copiers = [&]() { for (auto &t : copiers) t.join(); copiers.clear(); };
    auto fail = [&](int code, const std::string &msg) {   // the caller's buffer must not be read after we return
        join_copiers();
        cudaStreamSynchronize(ctx->stream);
        ctx->err = msg;
        return code;
    };
#define CKU(call)                                                                                  \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(_e == cudaErrorMemoryAllocation ? MPOA_ENOMEM : MPOA_ECUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

    ctx->ginfo.resize(n_groups);
    for (int64_t g = 0; g < n_groups; ++g) {
        GroupInfo &gi = ctx->ginfo[g];
        const int64_t r0 = ctx->h_gro[g], r1 = ctx->h_gro[g + 1];
        gi.n_reads = (int32_t)(r1 - r0);
        gi.maxlen = 0; gi.minlen = INT_MAX; gi.sumlen = 0;
        for (int64_t r = r0; r < r1; ++r) {
            const int64_t len = ctx->h_rbo[r + 1] - ctx->h_rbo[r];
            gi.maxlen = std::max<int32_t>(gi.maxlen, (int32_t)len);
            gi.minlen = std::min<int32_t>(gi.minlen, (int32_t)len);
            gi.sumlen += len;
        }
        if (gi.n_reads == 0) gi.minlen = 0;
        gi.wneed = band_need(ctx->params, gi.maxlen, gi.minlen, !ctx->h_flags.empty() && (ctx->h_flags[g] & MPOA_FLAG_SEED));
        gi.cost = (double)gi.sumlen * (double)gi.wneed;
        /* the packed kernels keep scores relative to the diagonal, so they serve abPOA's int16 AND int32
         * lane widths; a group leaves them only when the kernel itself asks for it (ST_RETRY_32) */
        gi.lanes16 = 1;
        gi.level = 0; gi.attempt = 0;
    }
    ctx->n_groups = n_groups; ctx->n_reads = n_reads; ctx->n_bases = n_bases;

    CKU(ctx->b_rbo.ensure((n_reads + 1) * sizeof(int64_t))); CKU(ctx->b_gro.ensure((n_groups + 1) * sizeof(int64_t)));
    CKU(ctx->b_region.ensure((n_groups + 1) * sizeof(int64_t)));
    CKU(ctx->b_len.ensure(n_groups * sizeof(int32_t))); CKU(ctx->b_status.ensure(n_groups * sizeof(int32_t)));
    CKU(ctx->b_queue.ensure(n_groups * sizeof(int32_t)));
    ctx->d_codes = (uint8_t *)ctx->b_codes.p;
    ctx->d_rbo = (int64_t *)ctx->b_rbo.p; ctx->d_gro = (int64_t *)ctx->b_gro.p; ctx->d_region_off = (int64_t *)ctx->b_region.p;
    /* the ASCII input is dead once it is encoded: its buffer becomes the consensus regions */
    ctx->d_cons = d_ascii;
    ctx->d_cons_len = (int32_t *)ctx->b_len.p; ctx->d_status = (int32_t *)ctx->b_status.p; ctx->d_queue = (int32_t *)ctx->b_queue.p;
    CKU(cudaMemcpyAsync(ctx->d_rbo, ctx->h_rbo.data(), (n_reads + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CKU(cudaMemcpyAsync(ctx->d_gro, ctx->h_gro.data(), (n_groups + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    /* consensus region of a group = the byte range of its reads: a consensus is a path of the
     * graph and can never hold more nodes than the group has bases */
    std::vector<int64_t> region(n_groups + 1);
    for (int64_t g = 0; g <= n_groups; ++g) region[g] = ctx->h_rbo[ctx->h_gro[g]];
    CKU(cudaMemcpyAsync(ctx->d_region_off, region.data(), (n_groups + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    /* `abpoa -S` groups: their anchors are computed on the host threads while the copies are in flight */
    ctx->n_seed_groups = 0;
    ctx->seed_ms = 0;
    for (uint8_t f : ctx->h_flags) ctx->n_seed_groups += (f & MPOA_FLAG_SEED) ? 1 : 0;
    if (ctx->n_seed_groups > 0) {
        /* seeding order = launch order: widest level first, heaviest group first (run_round, mpoa_batch_run) */
        const unsigned lmask = level_mask_env();
        std::vector<int32_t> &order = ctx->seed_order;
        order.clear();
        std::vector<int8_t> lvl(n_groups, 0);
        for (int64_t g = 0; g < n_groups; ++g)
            if (ctx->h_flags[g] & MPOA_FLAG_SEED) {
                order.push_back((int32_t)g);
                lvl[g] = (int8_t)pick_level(ctx->params, lmask, true, ctx->params.debug_small_caps ? 1 : ctx->ginfo[g].wneed);
            }
        std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            if (lvl[a] != lvl[b]) return lvl[a] > lvl[b];
            const GroupInfo &ga = ctx->ginfo[a], &gb = ctx->ginfo[b];
            return ga.cost != gb.cost ? ga.cost > gb.cost : a < b;
        });
        const int64_t n_order = (int64_t)order.size();
        ctx->seed_rank.assign(n_groups, -1);                 // unflagged groups never wait
        for (int64_t i = 0; i < n_order; ++i) ctx->seed_rank[order[i]] = (int32_t)i;
        /* anchors of a read are at least MPOA_SEED_MIN_W apart: its region can be laid out before they exist */
        ctx->seed_chunk = std::min<int64_t>(256, std::max<int64_t>(8, n_order / 64));
        ctx->n_seed_chunks = (n_order + ctx->seed_chunk - 1) / ctx->seed_chunk;
        ctx->chunk_slot.assign(ctx->n_seed_chunks + 1, 0);
        ctx->h_anc_off.assign((size_t)n_reads + 1, 0);
        int64_t slot = 0;
        for (int64_t i = 0; i < n_order; ++i) {
            if (i % ctx->seed_chunk == 0) ctx->chunk_slot[i / ctx->seed_chunk] = slot;
            const int64_t g = order[i];
            for (int64_t r = ctx->h_gro[g]; r < ctx->h_gro[g + 1]; ++r) {
                ctx->h_anc_off[r] = (int32_t)slot;
                slot += (ctx->h_rbo[r + 1] - ctx->h_rbo[r]) / MPOA_SEED_MIN_W + 2;
            }
        }
        ctx->chunk_slot[ctx->n_seed_chunks] = slot;
        if (slot > INT_MAX / 2) return fail(MPOA_EINVAL, "too many seeded bases in one batch");
        ctx->h_anc.resize((size_t)slot * 2);
        ctx->seed_done.reset(new std::atomic<int64_t>[ctx->n_seed_chunks]);
        for (int64_t c = 0; c < ctx->n_seed_chunks; ++c) ctx->seed_done[c].store(0);
        ctx->ready_vals.assign(ctx->n_seed_chunks, 0);
        CKU(ctx->b_anc_off.ensure(((size_t)n_reads + 1) * sizeof(int32_t)));
        CKU(ctx->b_anc.ensure(((size_t)slot + 1) * 2 * sizeof(int32_t)));
        CKU(ctx->b_seed_rank.ensure((size_t)n_groups * sizeof(int32_t)));
        CKU(ctx->b_seed_ready.ensure(sizeof(int32_t)));
        ctx->d_anc_off = (int32_t *)ctx->b_anc_off.p; ctx->d_anc = (int2 *)ctx->b_anc.p;
        ctx->d_seed_rank = (int32_t *)ctx->b_seed_rank.p; ctx->d_seed_ready = (int32_t *)ctx->b_seed_ready.p;
        CKU(cudaMemcpyAsync(ctx->d_anc_off, ctx->h_anc_off.data(), (size_t)n_reads * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CKU(cudaMemcpyAsync(ctx->d_seed_rank, ctx->seed_rank.data(), (size_t)n_groups * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CKU(cudaMemsetAsync(ctx->d_seed_ready, 0, sizeof(int32_t), ctx->stream));
        ctx->seed_wait_ms = 0;
        auto seed = [ctx, n_order, bases, defer_seed]() {
            /* one batch at a time: the seeding of the batch whose kernels are running is on the critical path and
             * uses every host thread; the next batch's (PoaPipeline) starts when it is done */
            std::lock_guard<std::mutex> one_at_a_time(g_seed_mu);
            const auto t0 = std::chrono::steady_clock::now();
            const int nt = (int)std::max(1u, std::thread::hardware_concurrency());
            seed_batch(ctx->h_gro.data(), ctx->h_rbo.data(), bases, ctx->h_src.empty() ? nullptr : ctx->h_src.data(),
                       ctx->seed_order.data(), n_order, ctx->seed_chunk, defer_seed ? ctx->seed_done.get() : nullptr,
                       MPOA_SEED_K, MPOA_SEED_W, MPOA_SEED_MIN_W, nt, ctx->h_anc_off.data(), ctx->h_anc.data());
            ctx->seed_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        };
        if (defer_seed) {
            ctx->seed_pending = true;
            ctx->seeder = std::thread(seed);          // its chunks are published by publish_seeds() while the kernels run
        } else {
            seed();
            if (slot > 0) CKU(cudaMemcpyAsync(ctx->d_anc, ctx->h_anc.data(), (size_t)slot * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    join_copiers();
    for (int w = 0; w < Stager::K; ++w) CKU(copy_err[w]);
    /* the bases are encoded (ASCII -> nt4) by the first mpoa_batch_run, inside its kernel phase: while the
     * persistent kernels of another context's batch hold the SMs, a kernel launched here would wait for them --
     * and the upload, which is meant to run BESIDE them, with it */
    ctx->d_ascii = d_ascii;
    ctx->need_encode = true;
    if (ctx->want_trace) {
        const size_t nr = (size_t)std::max<int64_t>(n_reads, 1);
        CKU(ctx->b_tr_score.ensure(nr * sizeof(int32_t))); CKU(ctx->b_tr_bits.ensure(nr * sizeof(int32_t)));
        CKU(ctx->b_tr_cells.ensure(nr * sizeof(long long)));
        CKU(ctx->b_tr_aln.ensure(nb * sizeof(int32_t))); CKU(ctx->b_tr_node.ensure(nb * sizeof(int32_t)));
        ctx->d_tr_score = (int32_t *)ctx->b_tr_score.p; ctx->d_tr_bits = (int32_t *)ctx->b_tr_bits.p;
        ctx->d_tr_cells = (long long *)ctx->b_tr_cells.p;
        ctx->d_tr_aln = (int32_t *)ctx->b_tr_aln.p; ctx->d_tr_node = (int32_t *)ctx->b_tr_node.p;
    }
    CKU(cudaStreamSynchronize(ctx->stream));      // region is a local; the caller's buffer is free again (unless seeding was deferred)
#undef CKU
    ctx->h2d_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    return MPOA_OK;
}

extern "C" int mpoa_batch_upload(mpoa_ctx *ctx, int64_t n_groups, const int64_t *group_read_off,
                                 const int64_t *read_base_off, const uint8_t *bases, const uint8_t *group_flags) {
    return upload_impl(ctx, n_groups, group_read_off, read_base_off, bases, group_flags, 0, nullptr);
}

extern "C" int mpoa_batch_upload_subset(mpoa_ctx *ctx, int64_t n_groups, const int64_t *group_read_off,
                                        const int64_t *read_base_off, const uint8_t *bases, const uint8_t *group_flags,
                                        int64_t n_sel, const int64_t *sel) {
    if (!sel && n_sel != 0) return MPOA_EINVAL;
    static const int64_t none = 0;
    return upload_impl(ctx, n_groups, group_read_off, read_base_off, bases, group_flags, n_sel, sel ? sel : &none);
}

/* capacities of one launch */
struct Caps {
    uint32_t ncap, ecap, qcap;
    uint64_t tbcap;
    int wcap, T, WPL;     // band capacity in cells, team size, words per lane (0: int32 lanes)
    int code() const { return variant_code(T, WPL); }
};

static uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

static SlotLayout make_layout(const Caps &c, bool seeded) {
    SlotLayout L;
    std::memset(&L, 0, sizeof(L));
    L.ncap = c.ncap; L.ecap = c.ecap; L.qcap = c.qcap; L.tbcap = c.tbcap;
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) { uint64_t o = off; off = align_up(off + bytes, 256); return o; };
    const uint64_t n = c.ncap + 2, e = c.ecap + 2, q = c.qcap + 2;
    for (int p = 0; p < 2; ++p) {
        L.base[p] = take(n); L.sib[p] = take(n); L.creator[p] = take(n * 4);
        L.in_off[p] = take(n * 4); L.in_row[p] = take(e * 4);
        L.out_off[p] = take(n * 4); L.out_row[p] = take(e * 4); L.out_w[p] = take(e * 4);
    }
    L.remain[0] = take(n * 4); L.meta[0] = take(n * 4); L.rowinfo = take(n * 16); L.rowtb = take(n * 16);
    L.rowbest = take(n * 4); L.qmap[0] = take(q * 4);
    if (seeded) {       // the sub-graph view of a window (poa_seed.cuh) and the previous read's rows
        L.in_off[2] = take(n * 4); L.in_row[2] = take(e * 4); L.remain[1] = take(n * 4); L.meta[1] = take(n * 4);
        L.qmap[1] = take(q * 4); L.prevrow = take(q * 4);
    }
    L.qprof = take(4ull * qprof_stride(c.qcap) * 4);
    L.pv = take(q * 4); L.pkey = take(q * 4); L.pnew = take(q * 4); L.psib = take(q * 4);
    L.nin = take(q * 4); L.nout = take(q * 4);
    L.cnt = take(n * 4); L.addin = take(n * 4); L.addout = take(n * 4); L.grow = take(n); L.srcof = take(n * 4);
    L.tb = take(c.tbcap + 64);
    L.slot_bytes = align_up(off, 1024);
    return L;
}

/* one kernel launch of a round: the groups of one level with one set of capacities */
struct Launch {
    int lv = 0;
    std::vector<int32_t> gs;
    Caps c;
    SlotLayout L;
    int wpb = 4, bps = 0;
    int64_t n_blocks = 0;
    uint64_t ws_off = 0;
    size_t q_off = 0;
    bool seeded = false;         // `abpoa -S` groups: the windowed kernel instantiation
    bool small = false;          // too few groups to fill the GPU: runs beside the big launches
    bool own_ws = false;         // big launch with a workspace of its own (may overlap the previous one's tail)
    uint64_t big_off = 0;
};

static void fill_args(mpoa_ctx *ctx, const Launch &ln, int k, KernelArgs &A) {
    std::memset(&A, 0, sizeof(A));
    A.codes = ctx->d_codes; A.read_off = ctx->d_rbo; A.group_read_off = ctx->d_gro;
    A.queue = ctx->d_queue + ln.q_off; A.n_queue = (int)ln.gs.size(); A.queue_head = ctx->d_queue_head + k;
    A.ws = g_dev[ctx->dev].ws + ln.ws_off; A.L = ln.L;
    A.cons = ctx->d_cons; A.cons_off = ctx->d_region_off; A.cons_len = ctx->d_cons_len; A.status = ctx->d_status;
    A.stats = ctx->d_stats;
    A.tr_score = ctx->d_tr_score; A.tr_bits = ctx->d_tr_bits; A.tr_cells = ctx->d_tr_cells;
    A.tr_aln = ctx->d_tr_aln; A.tr_node = ctx->d_tr_node;
    A.level = ln.lv;
    if (ln.seeded) {
        A.anc_off = ctx->d_anc_off; A.anc = ctx->d_anc; A.seed_k = MPOA_SEED_K;
        A.seed_rank = ctx->d_seed_rank;
        A.seed_ready = ctx->seed_pending ? ctx->d_seed_ready : nullptr;     // anchors still being computed: groups wait for theirs
    }
    A.tbcap_words = (uint32_t)std::min<uint64_t>(ln.L.tbcap / 4, 0xfffff000ull);
    const mpoa_params &p = ctx->params;
    A.P.match = std::abs(p.match); A.P.mismatch = std::abs(p.mismatch);
    A.P.o1 = p.gap_open1; A.P.e1 = p.gap_ext1; A.P.o2 = p.gap_open2; A.P.e2 = p.gap_ext2;
    A.P.oe1 = p.gap_open1 + p.gap_ext1; A.P.oe2 = p.gap_open2 + p.gap_ext2;
    A.P.wb = p.wb; A.P.wf = p.wf; A.P.pn16 = p.simd_pn_i16; A.P.pn32 = p.simd_pn_i32;
    A.wcap = ln.c.wcap;
    auto p2 = [](int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); };
    A.K.neg2 = p2(NEG16, NEG16);
    A.K.nee = p2(-A.P.e1, -A.P.e2);
    A.K.noe1 = p2(-A.P.oe1, -A.P.oe1); A.K.noe2 = p2(-A.P.oe2, -A.P.oe2);
    A.K.ne1 = p2(-A.P.e1, -A.P.e1); A.K.ne2 = p2(-A.P.e2, -A.P.e2);
    for (int t = 0; t < 16; ++t) A.K.tdec[t] = p2(-A.P.e1 * t, -A.P.e2 * t);
    A.K.mhop = p2(-A.P.match * 2 * ln.c.WPL, -A.P.match * 2 * ln.c.WPL);
    A.K.kc_r = (uint32_t)(2 * ln.c.WPL + ((A.P.match * 2 * ln.c.WPL) << 16));
    A.K.kc_l = (uint32_t)(-2 * ln.c.WPL + ((A.P.match * 2 * ln.c.WPL) << 16));
}

/*
 * Runs the launches of one round.  The BIG launches (enough groups to fill the GPU) run one after
 * the other on the context's stream, widest level first, each with as many persistent blocks as
 * the GPU holds: launches of different kernel variants that share the SMs were measured at 1.5x
 * the time of the same work run back to back (the variants evict each other from the instruction
 * cache).  SMALL launches (a few stragglers of a wide level, the int32 fall-back) occupy a handful
 * of SMs for as long as their longest group takes: they run beside the big ones on a second
 * stream, with their own workspace.  Groups that cannot be launched at all get ST_TOO_BIG.
 */
/*
 * Deferred seeding: the seeded launches are already on the device, every group waiting for its rank in the
 * seeding order (poa_group_kernel).  Chunk after chunk, as the host threads finish them: copy the chunk's slice
 * of the arena, then the new value of the counter, on a stream of their own -- the copies run beside the kernels.
 */
static int publish_seeds(mpoa_ctx *ctx) {
    if (!ctx->seed_pending) return MPOA_OK;
    if (!ctx->pub) CK(cudaStreamCreateWithFlags(&ctx->pub, cudaStreamNonBlocking));
    const int64_t n_order = (int64_t)ctx->seed_order.size();
    const auto t0 = std::chrono::steady_clock::now();
    for (int64_t c = 0; c < ctx->n_seed_chunks; ++c) {
        const int64_t want = std::min(ctx->seed_chunk, n_order - c * ctx->seed_chunk);
        while (ctx->seed_done[c].load(std::memory_order_acquire) < want) std::this_thread::sleep_for(std::chrono::microseconds(100));
        const int64_t lo = ctx->chunk_slot[c], hi = ctx->chunk_slot[c + 1];
        if (hi > lo)
            CK(cudaMemcpyAsync(ctx->d_anc + lo, ctx->h_anc.data() + 2 * lo, (size_t)(hi - lo) * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->pub));
        ctx->ready_vals[c] = (int32_t)std::min(n_order, (c + 1) * ctx->seed_chunk);
        CK(cudaMemcpyAsync(ctx->d_seed_ready, &ctx->ready_vals[c], sizeof(int32_t), cudaMemcpyHostToDevice, ctx->pub));
    }
    if (ctx->seeder.joinable()) ctx->seeder.join();
    ctx->seed_pending = false;
    CK(cudaStreamSynchronize(ctx->pub));
    ctx->seed_wait_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return MPOA_OK;
}

static int run_round(mpoa_ctx *ctx, std::vector<Launch> &launches, int64_t *n_launch) {
    std::vector<Launch *> live;
    for (Launch &ln : launches) {
        ln.wpb = 4;
        if (const char *w = getenv("MPOA_WPB")) ln.wpb = std::max(1, std::min(4, atoi(w)));   // tuning aid: warps per block
        while (ln.wpb > 1 && poa_smem_bytes(ln.c.code(), ln.c.wcap, ln.wpb) > ctx->smem_optin) ln.wpb >>= 1;
        ln.bps = poa_smem_bytes(ln.c.code(), ln.c.wcap, ln.wpb) > ctx->smem_optin
                     ? 0 : poa_max_blocks_per_sm(ln.c.code(), ln.c.wcap, ln.wpb, ln.seeded);
        if (ln.bps <= 0) {
            ctx->err = "band wider than shared memory allows";
            for (int32_t g : ln.gs) ctx->h_status[g] = ST_TOO_BIG;
            continue;
        }
        if (const char *b = getenv("MPOA_BPS")) ln.bps = std::min(ln.bps, std::max(1, atoi(b)));   // tuning aid
        ln.L = make_layout(ln.c, ln.seeded);
        live.push_back(&ln);
    }
    if (live.empty()) return MPOA_OK;
    /* widest first; while the anchors are still being computed on the host, everything unseeded goes first */
    const bool seeds_late = ctx->seed_pending;
    std::stable_sort(live.begin(), live.end(), [seeds_late](const Launch *a, const Launch *b) {
        if (seeds_late && a->seeded != b->seeded) return !a->seeded;
        return a->lv > b->lv;
    });
    const size_t min_groups = (size_t)ctx->n_sm * 16;
    size_t n_big = 0;
    for (Launch *ln : live) { ln->small = ln->gs.size() < min_groups; n_big += ln->small ? 0 : 1; }
    if (n_big == 0) live[0]->small = false;           // nothing to run beside
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    DevShared &D = g_dev[ctx->dev];                   // (the caller holds D.mu)
    const uint64_t budget = (uint64_t)(((uint64_t)free_b + D.ws_bytes) * 0.85);
    /* the big launches share one workspace (they run one after the other), every small one has its own */
    uint64_t need_big = 0, need_small = 0;
    for (Launch *ln : live) {
        const int tpb = ln->wpb * (32 / ln->c.T);     // teams (= workspace slots) per block
        int64_t nb = (int64_t)ln->bps * ctx->n_sm;
        nb = std::min<int64_t>(nb, ((int64_t)ln->gs.size() + tpb - 1) / tpb);
        nb = std::min<int64_t>(nb, (int64_t)((budget / 2) / ((uint64_t)tpb * ln->L.slot_bytes)));
        ln->n_blocks = nb;
        const uint64_t bytes = (uint64_t)nb * tpb * ln->L.slot_bytes;
        if (ln->small) { ln->ws_off = need_small; need_small += bytes; } else need_big = std::max(need_big, bytes);
    }
    /* big launches overlap their tails when every one of them can have its own workspace */
    {
        uint64_t sum_big = 0;
        for (Launch *ln : live) if (!ln->small) sum_big += (uint64_t)ln->n_blocks * ln->wpb * (32 / ln->c.T) * ln->L.slot_bytes;
        if (n_big > 1 && sum_big + need_small <= budget) {
            uint64_t off = 0;
            for (Launch *ln : live) if (!ln->small) { ln->own_ws = true; ln->big_off = off; off += (uint64_t)ln->n_blocks * ln->wpb * (32 / ln->c.T) * ln->L.slot_bytes; }
            need_big = sum_big;
        }
    }
    if (need_small > budget / 2) {                    // too much for side by side: everything back to back
        for (Launch *ln : live) { if (ln->small) need_big = std::max(need_big, (uint64_t)ln->n_blocks * ln->wpb * (32 / ln->c.T) * ln->L.slot_bytes); ln->small = false; }
        need_small = 0;
    }
    const uint64_t need = need_big + need_small;
    if (need > D.ws_bytes) {
        cudaFree(D.ws);
        D.ws = nullptr; D.ws_bytes = 0;
        CK(cudaMalloc(&D.ws, need));
        D.ws_bytes = need;
    }
    size_t q_off = 0;
    std::vector<int32_t> hq;
    bool any_small = false;
    for (Launch *ln : live) {
        ln->ws_off = ln->small ? need_big + ln->ws_off : (ln->own_ws ? ln->big_off : 0);
        ln->q_off = q_off;
        q_off += ln->gs.size();
        hq.insert(hq.end(), ln->gs.begin(), ln->gs.end());
        any_small = any_small || ln->small;
    }
    if (live.size() > 16) { ctx->err = "too many launch levels"; return MPOA_EINVAL; }
    CK(cudaMemcpyAsync(ctx->d_queue, hq.data(), hq.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_queue_head, 0, 16 * sizeof(int), ctx->stream));
    bool used_side2 = false;
    if (any_small || n_big > 1) {
        if (!ctx->side) CK(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
        if (!ctx->fork_ev) CK(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
        if (!ctx->join_ev) CK(cudaEventCreateWithFlags(&ctx->join_ev, cudaEventDisableTiming));
        CK(cudaEventRecord(ctx->fork_ev, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->side, ctx->fork_ev, 0));
    }
    const bool verbose = getenv("MPOA_VERBOSE") != nullptr;
    const bool overlap = getenv("MPOA_NO_OVERLAP") == nullptr;
    bool first_big = true;
    for (size_t k = 0; k < live.size(); ++k) {
        Launch *ln = live[k];
        if (ln->n_blocks <= 0) {
            ctx->err = "not enough device memory for one workspace slot";
            for (int32_t g : ln->gs) ctx->h_status[g] = ST_TOO_BIG;
            continue;
        }
        KernelArgs A;
        fill_args(ctx, *ln, (int)k, A);
        cudaStream_t st = ln->small ? ctx->side : ctx->stream;
        if (!ln->small && !first_big && overlap && ln->own_ws) {
            /* the next big launch waits on ANOTHER stream for SM slots: its blocks move in while the
             * previous launch drains its last groups (the persistent blocks of a launch retire one by
             * one), instead of after the last one has gone */
            if (!ctx->side2) CK(cudaStreamCreateWithFlags(&ctx->side2, cudaStreamNonBlocking));
            if (!ctx->join2_ev) CK(cudaEventCreateWithFlags(&ctx->join2_ev, cudaEventDisableTiming));
            CK(cudaStreamWaitEvent(ctx->side2, ctx->fork_ev, 0));
            st = ctx->side2;
            used_side2 = true;
        }
        if (!ln->small) first_big = false;
        CK(launch_poa(ln->c.code(), A, (int)ln->n_blocks, ln->wpb, st));
        ++*n_launch;
        if (verbose)
            fprintf(stderr, "[mpoa] level T=%d WPL=%d wcap=%d groups=%zu blocks=%lld x %d warps (bps %d) slot=%.1f MB%s%s\n",
                    ln->c.T, ln->c.WPL, ln->c.wcap, ln->gs.size(), (long long)ln->n_blocks, ln->wpb, ln->bps,
                    ln->L.slot_bytes / 1e6, ln->small ? " (side stream)" : "", ln->seeded ? " seeded" : "");
    }
    if (any_small || n_big > 1) {
        CK(cudaEventRecord(ctx->join_ev, ctx->side));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->join_ev, 0));
    }
    if (used_side2) {
        CK(cudaEventRecord(ctx->join2_ev, ctx->side2));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->join2_ev, 0));
    }
    return publish_seeds(ctx);                        // (first round only) feeds the seeded launches while they run
}

extern "C" int mpoa_batch_run(mpoa_ctx *ctx, mpoa_stats *stats) {
    if (!ctx) return MPOA_EINVAL;
    CK(cudaSetDevice(ctx->dev));
    mpoa_stats st;
    std::memset(&st, 0, sizeof(st));
    st.n_groups = ctx->n_groups; st.n_reads = ctx->n_reads;
    st.h2d_ms = ctx->h2d_ms;
    const int64_t ng = ctx->n_groups;
    ctx->h_status.assign(ng, ST_PENDING);
    if (ng == 0) { ctx->ran = true; ctx->last = st; if (stats) *stats = st; return MPOA_OK; }
    CK(cudaMemsetAsync(ctx->d_stats, 0, SI_COUNT * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_cons_len, 0, ng * sizeof(int32_t), ctx->stream));
    /* kernel phase: one context per device at a time (DevShared); taken in front of the first launch */
    std::unique_lock<std::mutex> kernel_phase(g_dev[ctx->dev].mu, std::defer_lock);

    std::vector<int32_t> pending(ng);
    std::iota(pending.begin(), pending.end(), 0);
    const unsigned level_mask = level_mask_env();
    auto pick_level = [&](const GroupInfo &gi, int wneed) { return ::pick_level(ctx->params, level_mask, gi.lanes16 != 0, wneed); };
    for (int64_t g = 0; g < ng; ++g) {
        GroupInfo &gi = ctx->ginfo[g];
        gi.lanes16 = 1;
        gi.level = (int8_t)pick_level(gi, ctx->params.debug_small_caps ? 1 : gi.wneed);
        gi.attempt = 0;
    }
    int64_t n_launch = 0;
    std::vector<int32_t> dstat(ng);
    for (int round = 0; round < 12 && !pending.empty(); ++round) {
        /* one launch per level.  A level with fewer groups than the GPU holds teams is folded into
         * the next wider level of the same lane family when that one fills the GPU (every group runs
         * correctly in any level at least as wide as its own); what stays small runs beside the big
         * launches (run_round) */
        /* bins [0, kNumLevels): unseeded groups per level; [kNumLevels, 2 kNumLevels): `abpoa -S` groups */
        auto is_seeded = [&](int32_t g) { return !ctx->h_flags.empty() && (ctx->h_flags[g] & MPOA_FLAG_SEED) != 0; };
        std::vector<std::vector<int32_t>> bins(2 * kNumLevels);
        for (int32_t g : pending) bins[ctx->ginfo[g].level + (is_seeded(g) ? kNumLevels : 0)].push_back(g);
        const size_t min_groups = (size_t)ctx->n_sm * 16;
        for (int half = 0; half < 2; ++half)
            for (int a = 0; a < kNumLevels; ++a) {
                auto &ba = bins[half * kNumLevels + a];
                if (ba.empty() || ba.size() >= min_groups) continue;
                for (int o = a + 1; o < kNumLevels; ++o) {
                    auto &bo = bins[half * kNumLevels + o];
                    if (bo.size() >= min_groups && (kLevels[o].WPL == 0) == (kLevels[a].WPL == 0)) {
                        bo.insert(bo.end(), ba.begin(), ba.end());
                        ba.clear();
                        break;
                    }
                }
            }
        std::vector<Launch> launches;
        for (int bi = 0; bi < 2 * kNumLevels; ++bi) {
            const int lv = bi % kNumLevels;
            auto &gs = bins[bi];
            if (gs.empty()) continue;
            /* heaviest first (longest-processing-time order keeps the tail of the launch short); for
             * two teams per warp: longest reads first, so that the groups the teams of one warp work on
             * at the same time have about the same number of graph rows */
            const bool by_len = kLevels[lv].T == 16;
            std::sort(gs.begin(), gs.end(), [&](int32_t a, int32_t b) {
                const GroupInfo &ga = ctx->ginfo[a], &gb = ctx->ginfo[b];
                if (by_len && ga.maxlen != gb.maxlen) return ga.maxlen > gb.maxlen;
                return ga.cost != gb.cost ? ga.cost > gb.cost : a < b;
            });
            Caps c;
            c.wcap = kLevels[lv].wcap; c.T = kLevels[lv].T; c.WPL = kLevels[lv].WPL;
            uint64_t ncap = 0, qcap = 0, sum_max = 0, tbmax = 0;
            int att_max = 0;
            const uint64_t cell_bytes = c.WPL == 0 ? 12 : 6;
            for (int32_t g : gs) {
                const GroupInfo &gi = ctx->ginfo[g];
                uint64_t est;
                if (gi.attempt == 0 && ctx->params.debug_small_caps) est = (uint64_t)gi.maxlen + 8;
                else if (gi.attempt == 0) est = (uint64_t)(gi.maxlen * (1.0 + 0.02 * gi.n_reads)) + 24ull * gi.n_reads + 256;
                else if (gi.attempt == 1) est = (uint64_t)(gi.maxlen * (1.0 + 0.15 * gi.n_reads)) + 64ull * gi.n_reads + 1024;
                else est = (uint64_t)gi.sumlen + 2;
                est = std::min<uint64_t>(est, (uint64_t)gi.sumlen + 2);
                ncap = std::max(ncap, est);
                qcap = std::max<uint64_t>(qcap, gi.maxlen);
                sum_max = std::max<uint64_t>(sum_max, gi.sumlen);
                att_max = std::max<int>(att_max, gi.attempt);
                const uint64_t wrow = std::min<uint64_t>(std::min<uint64_t>(c.wcap, gi.maxlen + 64),
                                                         gi.attempt == 0 ? (uint64_t)gi.wneed : (uint64_t)c.wcap) + 32;
                tbmax = std::max<uint64_t>(tbmax, est * wrow * cell_bytes);
            }
            c.ncap = (uint32_t)std::min<uint64_t>(ncap + 8, 0x7fffff00u);
            c.ecap = (uint32_t)std::min<uint64_t>(att_max >= 2 ? sum_max + 4096 : std::min<uint64_t>(2ull * c.ncap + 64, sum_max + 4096),
                                                  0x7fffff00u);
            c.qcap = (uint32_t)qcap + 8;
            c.tbcap = std::min<uint64_t>(tbmax + 4096, 0x3fff00000ull);
            launches.emplace_back();
            launches.back().lv = lv;
            launches.back().seeded = bi >= kNumLevels;
            launches.back().gs = gs;
            launches.back().c = c;
        }
        {
            if (!kernel_phase.owns_lock()) {
                const auto tw = std::chrono::steady_clock::now();
                kernel_phase.lock();
                st.kernel_wait_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw).count();
                CK(cudaEventRecord(ctx->ev0, ctx->stream));
                if (ctx->need_encode) {
                    CK(launch_encode(ctx->d_ascii, ctx->d_codes, ctx->n_bases, ctx->stream));
                    ctx->need_encode = false;
                    ++n_launch;
                }
            }
            const int rc = run_round(ctx, launches, &n_launch);
            if (rc < 0) return rc;
        }
        /* wait FIRST: a copy into pageable memory keeps its caller inside the driver until the stream gets to
         * it -- here: behind the kernels -- and other host threads' CUDA calls (another context uploading its
         * batch beside these kernels) were measured to wait with it */
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaMemcpyAsync(dstat.data(), ctx->d_status, ng * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::vector<int32_t> next;
        for (int32_t g : pending) {
            if (ctx->h_status[g] == ST_EMPTY || ctx->h_status[g] == ST_TOO_BIG) continue;  // not launchable
            GroupInfo &gi = ctx->ginfo[g];
            const int code = dstat[g] & 0xff;
            /* escalate from the level the group actually RAN at (it may have been folded into, or
             * stolen by, a wider launch than the one it was scheduled for) */
            const int ran = std::min(kNumLevels - 1, std::max((int)gi.level, (dstat[g] >> 8) & 0xff));
            if (code == ST_RETRY) {
                if (gi.attempt >= 3) { ctx->h_status[g] = ST_TOO_BIG; continue; }
                gi.attempt++;
                gi.level = (int8_t)ran;
                next.push_back(g);
            } else if (code == ST_RETRY_WIDE || code == ST_RETRY_32) {
                if (code == ST_RETRY_32) gi.lanes16 = 0;                       // junk reads / exotic parameters: int32 lanes
                else gi.wneed = std::max(gi.wneed, kLevels[ran].wcap + 1);     // the band outgrew the level it ran at
                const int nl = pick_level(gi, gi.wneed);
                if (kLevels[nl].wcap < gi.wneed) { ctx->h_status[g] = ST_TOO_BIG; continue; }
                gi.level = (int8_t)nl;
                next.push_back(g);
            } else ctx->h_status[g] = dstat[g];
        }
        if (round == 0) st.n_retry_groups = (int64_t)next.size();
        pending.swap(next);
    }
    for (int32_t g : pending) ctx->h_status[g] = ST_TOO_BIG;  // still too big after the last attempt
    { const int rc = publish_seeds(ctx); if (rc != MPOA_OK) return rc; }   // no launch was made at all
    if (!kernel_phase.owns_lock()) { kernel_phase.lock(); CK(cudaEventRecord(ctx->ev0, ctx->stream)); }
    /* still inside the kernel phase: the consensi are gathered into one compact buffer, so that mpoa_batch_fetch
     * is copies only (it may run while the next batch's kernels hold the SMs) */
    {
        std::vector<int32_t> len(ng);
        CK(cudaMemcpyAsync(len.data(), ctx->d_cons_len, ng * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->h_cons_off.assign(ng + 1, 0);
        for (int64_t g = 0; g < ng; ++g) ctx->h_cons_off[g + 1] = ctx->h_cons_off[g] + (ctx->h_status[g] == ST_OK ? len[g] : 0);
        const int64_t total = ctx->h_cons_off[ng];
        if (total > 0) {
            CK(ctx->b_out_off.ensure((ng + 1) * sizeof(int64_t)));
            CK(ctx->b_out.ensure(total));
            CK(cudaMemcpyAsync(ctx->b_out_off.p, ctx->h_cons_off.data(), (ng + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
            CK(launch_gather(ctx->d_cons, ctx->d_region_off, (int64_t *)ctx->b_out_off.p, (uint8_t *)ctx->b_out.p, ng, ctx->stream));
            ++n_launch;
        }
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    kernel_phase.unlock();
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    unsigned long long hs[SI_COUNT];
    CK(cudaMemcpyAsync(hs, ctx->d_stats, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    st.kernel_ms = ms;
    st.band_cells = (int64_t)hs[SI_CELLS]; st.int_ops = (int64_t)hs[SI_INTOPS]; st.full_cells = (int64_t)hs[SI_FULL];
    st.n_alignments = (int64_t)hs[SI_ALN]; st.n_align_i16 = (int64_t)hs[SI_ALN16]; st.n_align_i32 = (int64_t)hs[SI_ALN32];
    st.tb_bytes = (int64_t)hs[SI_TB];
    for (int k = 0; k < 6; ++k) st.phase_cycles[k] = (int64_t)hs[SI_T_PREP + k];
    st.n_kernel_launches = n_launch;
    /* groups the reference would have run with `abpoa -S` (median read length >= 8000) */
    st.n_seed_groups = ctx->n_seed_groups;
    st.n_seed_applied = ctx->n_seed_groups;      // every kernel variant has its windowed instantiation
    st.host_seed_ms = ctx->seed_ms;
    for (int64_t g = 0; g < ng; ++g) st.n_too_big_groups += ctx->h_status[g] == ST_TOO_BIG ? 1 : 0;
    ctx->last = st;
    ctx->ran = true;
    if (stats) *stats = st;
    return MPOA_OK;
}

extern "C" int mpoa_batch_fetch(mpoa_ctx *ctx, int64_t *cons_off, uint8_t *cons_buf, int64_t cons_cap,
                                int32_t *group_status, mpoa_trace *trace) {
    if (!ctx || !cons_off) return MPOA_EINVAL;
    if (!ctx->ran) { ctx->err = "mpoa_batch_fetch before mpoa_batch_run"; return MPOA_EINVAL; }
    CK(cudaSetDevice(ctx->dev));
    const int64_t ng = ctx->n_groups;
    cons_off[0] = 0;
    if (ng == 0) return MPOA_OK;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int64_t g = 0; g < ng; ++g) {
        const bool ok = ctx->h_status[g] == ST_OK;
        cons_off[g + 1] = ctx->h_cons_off[g + 1];
        if (group_status) group_status[g] = ok ? MPOA_GROUP_OK : ctx->h_status[g] == ST_TOO_BIG ? MPOA_GROUP_TOO_BIG : MPOA_GROUP_EMPTY;
    }
    const int64_t total = cons_off[ng];
    int rc = MPOA_OK;
    if (total > cons_cap || (total > 0 && !cons_buf)) rc = MPOA_ENOSPC;
    else if (total > 0) CK(cudaMemcpyAsync(cons_buf, ctx->b_out.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    if (trace && ctx->want_trace) {
        const int64_t nr = ctx->n_reads, nb = ctx->n_bases;
        if (trace->read_score) CK(cudaMemcpyAsync(trace->read_score, ctx->d_tr_score, nr * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (trace->read_bits) CK(cudaMemcpyAsync(trace->read_bits, ctx->d_tr_bits, nr * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (trace->read_band_cells) CK(cudaMemcpyAsync(trace->read_band_cells, ctx->d_tr_cells, nr * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        if (trace->base_aln) CK(cudaMemcpyAsync(trace->base_aln, ctx->d_tr_aln, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (trace->base_node) CK(cudaMemcpyAsync(trace->base_node, ctx->d_tr_node, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last.d2h_ms = ms;
    return rc;
}

extern "C" int mpoa_consensus_batch(mpoa_ctx *ctx, int64_t n_groups, const int64_t *group_read_off,
                                    const int64_t *read_base_off, const uint8_t *bases, const uint8_t *group_flags,
                                    int64_t *cons_off, uint8_t *cons_buf, int64_t cons_cap, int32_t *group_status,
                                    mpoa_stats *stats, mpoa_trace *trace) {
    if (!ctx) return MPOA_EINVAL;
    const int saved = ctx->want_trace;
    if (trace) ctx->want_trace = 1;
    /* the caller's buffers stay valid until this call returns: the `-S` anchors are computed beside the kernels */
    int rc = upload_impl(ctx, n_groups, group_read_off, read_base_off, bases, group_flags, 0, nullptr, true);
    if (rc == MPOA_OK) rc = mpoa_batch_run(ctx, nullptr);
    if (ctx->seeder.joinable()) ctx->seeder.join();     // an error path may have left it running on the caller's buffer
    if (rc == MPOA_OK) rc = mpoa_batch_fetch(ctx, cons_off, cons_buf, cons_cap, group_status, trace);
    ctx->want_trace = saved;
    if (stats) *stats = ctx->last;
    return rc;
}

/* ---- one batch over several GPUs of one box ---- */

extern "C" int mpoa_shard_plan(int64_t n_groups, const int64_t *group_read_off, const int64_t *read_base_off,
                               const uint8_t *group_flags, const mpoa_params *p, int32_t n_shards, int32_t *owner) {
    if (n_groups < 0 || n_shards <= 0 || (n_groups > 0 && (!group_read_off || !read_base_off || !owner))) return MPOA_EINVAL;
    mpoa_params dp;
    if (p) dp = *p; else mpoa_default_params(&dp);
    if (dp.simd_pn_i16 <= 0) dp.simd_pn_i16 = 16;
    std::vector<double> cost(n_groups);
    const int64_t n_reads = n_groups > 0 ? group_read_off[n_groups] : 0;
    for (int64_t g = 0; g < n_groups; ++g) {
        const int64_t r0 = group_read_off[g], r1 = group_read_off[g + 1];
        if (r0 < 0 || r1 < r0 || r1 > n_reads) return MPOA_EINVAL;
        int64_t sum = 0, mx = 0, mn = INT64_MAX;
        for (int64_t r = r0; r < r1; ++r) {
            const int64_t len = read_base_off[r + 1] - read_base_off[r];
            sum += len; mx = std::max(mx, len); mn = std::min(mn, len);
        }
        if (r1 == r0) mn = 0;
        /* reads 2..n are aligned; every row also pays a fixed cost (graph walk, traceback, merge) */
        cost[g] = (double)sum * (double)(band_need(dp, (int)std::min<int64_t>(mx, 1 << 26), (int)std::min<int64_t>(mn, 1 << 26),
                                                   group_flags && (group_flags[g] & MPOA_FLAG_SEED)) + 64);
    }
    /* longest-processing-time greedy: heaviest group first, each to the least loaded shard */
    std::vector<int64_t> order(n_groups);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return cost[a] != cost[b] ? cost[a] > cost[b] : a < b; });
    std::vector<double> load(n_shards, 0.0);
    for (int64_t g : order) {
        int best = 0;
        for (int s = 1; s < n_shards; ++s) if (load[s] < load[best]) best = s;
        owner[g] = best;
        load[best] += cost[g];
    }
    return MPOA_OK;
}

extern "C" int mpoa_consensus_batch_multi(mpoa_ctx *const *ctxs, int32_t n_ctx, int64_t n_groups,
                                          const int64_t *group_read_off, const int64_t *read_base_off, const uint8_t *bases,
                                          const uint8_t *group_flags, int64_t *cons_off, uint8_t *cons_buf, int64_t cons_cap,
                                          int32_t *group_status, mpoa_stats *stats, int32_t *owner_out) {
    if (!ctxs || n_ctx <= 0 || n_groups < 0 || !cons_off) return MPOA_EINVAL;
    for (int k = 0; k < n_ctx; ++k) {
        if (!ctxs[k]) return MPOA_EINVAL;
        for (int j = 0; j < k; ++j) if (ctxs[j] == ctxs[k]) { ctxs[k]->err = "the same context twice"; return MPOA_EINVAL; }
    }
    cons_off[0] = 0;
    if (stats) std::memset(stats, 0, sizeof(mpoa_stats) * (size_t)n_ctx);
    if (n_groups == 0) return MPOA_OK;
    std::vector<int32_t> owner(n_groups, 0);
    if (n_ctx > 1) {
        const int rc = mpoa_shard_plan(n_groups, group_read_off, read_base_off, group_flags, &ctxs[0]->params, n_ctx, owner.data());
        if (rc != MPOA_OK) { ctxs[0]->err = "bad group offsets"; return rc; }
    }
    if (owner_out) std::copy(owner.begin(), owner.end(), owner_out);
    struct Part {
        std::vector<int64_t> sel, off;
        std::vector<int32_t> status;
        std::unique_ptr<uint8_t[]> buf;
        int rc = MPOA_OK;
    };
    std::vector<Part> parts(n_ctx);
    for (int64_t g = 0; g < n_groups; ++g) parts[owner[g]].sel.push_back(g);
    auto run_part = [&](int k) {
        Part &pt = parts[k];
        mpoa_ctx *ctx = ctxs[k];
        const int64_t n = (int64_t)pt.sel.size();
        pt.off.assign(n + 1, 0);
        pt.status.assign(n, MPOA_GROUP_EMPTY);
        if (n == 0) { free_batch(ctx); return; }
        pt.rc = upload_impl(ctx, n_groups, group_read_off, read_base_off, bases, group_flags, n, pt.sel.data(), true);
        if (pt.rc == MPOA_OK) pt.rc = mpoa_batch_run(ctx, nullptr);
        if (ctx->seeder.joinable()) ctx->seeder.join();
        if (pt.rc != MPOA_OK) return;
        const int64_t cap = std::max<int64_t>(ctx->n_bases, 16);    // a consensus never outgrows its group's bases
        pt.buf.reset(new uint8_t[cap]);
        pt.rc = mpoa_batch_fetch(ctx, pt.off.data(), pt.buf.get(), cap, pt.status.data(), nullptr);
        if (stats) stats[k] = ctx->last;
    };
    auto for_parts = [&](auto &&fn) {
        if (n_ctx == 1) { fn(0); return; }
        std::vector<std::thread> th;
        for (int k = 0; k < n_ctx; ++k) th.emplace_back(fn, k);
        for (auto &t : th) t.join();
    };
    for_parts(run_part);
    for (int k = 0; k < n_ctx; ++k)
        if (parts[k].rc != MPOA_OK) {
            if (k != 0) ctxs[0]->err = "device " + std::to_string(ctxs[k]->dev) + ": " + ctxs[k]->err;
            return parts[k].rc;
        }
    /* results back in input order */
    for (int k = 0; k < n_ctx; ++k) {
        const Part &pt = parts[k];
        for (size_t i = 0; i < pt.sel.size(); ++i) cons_off[pt.sel[i] + 1] = pt.off[i + 1] - pt.off[i];
    }
    for (int64_t g = 0; g < n_groups; ++g) cons_off[g + 1] += cons_off[g];
    if (cons_off[n_groups] > cons_cap || (cons_off[n_groups] > 0 && !cons_buf)) {
        ctxs[0]->err = "cons_buf too small";
        return MPOA_ENOSPC;
    }
    for_parts([&](int k) {
        const Part &pt = parts[k];
        for (size_t i = 0; i < pt.sel.size(); ++i) {
            const int64_t g = pt.sel[i], n = pt.off[i + 1] - pt.off[i];
            if (n > 0) std::memcpy(cons_buf + cons_off[g], pt.buf.get() + pt.off[i], (size_t)n);
            if (group_status) group_status[g] = pt.status[i];
        }
    });
    return MPOA_OK;
}
