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
 *   poa_group_kernel<V>      persistent, one warp per read group: graph build, adaptive-banded
 *                            convex-gap DP (V=0 int32 lanes, V=2/4/8 packed int16x2 DPX), value
 *                            traceback, graph merge, heaviest-bundle consensus
 *   gather_consensus_kernel  per-group consensus regions -> one compact buffer
 */
#include "poa_traceback.cuh"

#ifndef MPOA_V2_BLOCKS
#define MPOA_V2_BLOCKS 5
#endif
#ifndef MPOA_V0_BLOCKS
#define MPOA_V0_BLOCKS 3
#endif
#ifndef MPOA_V4_BLOCKS
#define MPOA_V4_BLOCKS 5
#endif

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

/* TRACE: the per-read / per-base trace outputs of the ABI (parity tests).  It is a compile-time
 * switch because the kernel's instruction footprint matters: the production instantiation is
 * ~370 SASS instructions smaller and measurably faster (resident warps share the instruction cache). */
template <int V, bool TRACE>
__device__ __forceinline__ int process_group(const KernelArgs &A, int slot, int g, int *ring, int4 *ring_info, int lane,
                                             unsigned long long *st /* per-warp counters in shared memory, lane 0 only */) {
    const int64_t r0 = A.group_read_off[g], r1 = A.group_read_off[g + 1];
    if (r1 <= r0) return ST_EMPTY;
    const int64_t gbase = A.read_off[r0];
    int par = 0, N = 0, E = 0;
    Slot S = make_slot(A, slot, par);
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t b0 = A.read_off[r], b1 = A.read_off[r + 1];
        const int len = (int)(b1 - b0);
        const uint8_t *seq = A.codes + b0;
        const int creator0 = (int)(b0 - gbase);
        int32_t *tr_aln = nullptr, *tr_node = nullptr;
        if constexpr (TRACE) { tr_aln = A.tr_aln + b0; tr_node = A.tr_node + b0; }
        if constexpr (TRACE) {
            if (lane == 0) { A.tr_score[r] = 0; A.tr_bits[r] = 0; A.tr_cells[r] = 0; }
        }
        if (N == 0) {
            if (len <= 0) return ST_EMPTY;
            if ((uint32_t)(len + 2) > A.L.ncap || (uint32_t)(len + 1) > A.L.ecap) return ST_RETRY;
            init_graph<TRACE>(A, S, seq, len, creator0, lane);
            N = len + 2; E = len + 1;
            for (int t = lane; t < len; t += 32) {
                if constexpr (TRACE) { tr_aln[t] = -1; tr_node[t] = creator0 + t; }
            }
            continue;
        }
        if (len <= 0) continue;
        if ((uint32_t)len > A.L.qcap) return ST_RETRY;
        long long tk0 = clock64();
        remain_pass(A, S, N, lane);
        long long tk1 = clock64();
        if (lane == 0) st[SI_T_PREP] += tk1 - tk0;
        AlnState R;
        int rc;
        if constexpr (V == 0) rc = dp_align32(A, S, N, seq, len, ring, ring_info, lane, R);
        else rc = dp_align16<V>(A, S, N, seq, len, reinterpret_cast<uint32_t *>(ring), lane, R);
        if (rc != ST_OK) return rc;
        tk0 = clock64();
        if (lane == 0) {
            st[SI_T_DP] += tk0 - tk1;
            st[SI_CELLS] += R.cells; st[SI_INTOPS] += R.intops; st[SI_FULL] += R.full; st[SI_ALN] += 1;
            st[R.bits == 16 ? SI_ALN16 : SI_ALN32] += 1; st[SI_TB] += R.tbbytes;
            if constexpr (TRACE) { A.tr_score[r] = R.best_score; A.tr_bits[r] = R.bits; A.tr_cells[r] = (long long)R.cells; }
        }
        __syncwarp();
        bool ok;
        if constexpr (V == 0) ok = traceback<int32_t>(A, S, seq, len, R, lane, ring);
        else ok = traceback<int16_t>(A, S, seq, len, R, lane, ring);
        __syncwarp();
        if (!ok) return ST_EMPTY;
        tk1 = clock64();
        if (lane == 0) st[SI_T_TB] += tk1 - tk0;
        const int mrc = merge_read(A, S, par, N, E, seq, len, creator0, tr_aln, tr_node, lane);
        if (mrc != ST_OK) return mrc;
        S = make_slot(A, slot, par);
        if (lane == 0) st[SI_T_MERGE] += clock64() - tk1;
    }
    if (N <= 2) return ST_EMPTY;
    const long long tc0 = clock64();
    int clen = 0;
    if (lane == 0) {
        const int cap = (int)(A.cons_off[g + 1] - A.cons_off[g]);
        clen = heaviest_bundle(A, S, N, A.cons + A.cons_off[g], cap);
    }
    clen = __shfl_sync(FULL, clen, 0);
    __syncwarp();
    if (clen < 0) return ST_RETRY;
    if (lane == 0) { A.cons_len[g] = clen; st[SI_T_CONS] += clock64() - tc0; }
    return ST_OK;
}

template <int V>
__host__ __device__ constexpr int variant_warp_words(int wcap) {
    /* + 32 words: the warp's work counters (SI_COUNT x u64) */
    return (V == 0 ? RING * 3 * wcap + RING * 4 : ring16_warp_words<(V == 0 ? 2 : V)>()) + 32;
}

template <int V, bool TRACE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, (V == 2 ? MPOA_V2_BLOCKS : V == 4 ? MPOA_V4_BLOCKS : V == 8 ? 3 : MPOA_V0_BLOCKS))
poa_group_kernel(const __grid_constant__ KernelArgs A) {
    extern __shared__ __align__(16) int smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = variant_warp_words<V>(A.wcap);
    int *ring = smem + warp * per_warp;
    int4 *ring_info = reinterpret_cast<int4 *>(ring + per_warp - 32 - RING * 4);   // int32 variant only
    /* work counters of the group in flight: shared memory, touched by lane 0 only (they used to sit
     * in 26 registers across the DP row loop) */
    unsigned long long *gst = reinterpret_cast<unsigned long long *>(ring + per_warp - 32);
    static_assert(SI_COUNT <= 16, "counter block");
    const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
    int src = -1;   // -1: own queue, k >= 0: steal queue k
    for (;;) {
        int qi = 0, g = -1;
        for (;;) {
            int *head = src < 0 ? A.queue_head : A.steal_head[src];
            const int n = src < 0 ? A.n_queue : A.steal_n[src];
            if (lane == 0) qi = atomicAdd(head, 1);
            qi = __shfl_sync(FULL, qi, 0);
            if (qi < n) { g = (src < 0 ? A.queue : A.steal_queue[src])[qi]; break; }
            if (++src >= A.n_steal) break;
        }
        if (g < 0) break;
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < SI_COUNT; ++k) gst[k] = 0;
        }
        const long long tg0 = clock64();
        int rc;
        if constexpr (V == 0) rc = process_group<V, TRACE>(A, slot, g, ring, ring_info, lane, gst);
        else rc = process_group<V, TRACE>(A, slot, g, ring, nullptr, lane, gst);
        if (lane == 0) {
            A.status[g] = rc;
            if (rc != ST_OK) A.cons_len[g] = 0;
            if (rc != ST_RETRY && rc != ST_RETRY_WIDE && rc != ST_RETRY_32) {
                gst[SI_T_BUSY] += clock64() - tg0;
                for (int k = 0; k < SI_COUNT; ++k)
                    if (gst[k]) atomicAdd(A.stats + k, gst[k]);
            }
        }
        __syncwarp();
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

template <int V>
static const void *variant_fn(bool trace) {
    return trace ? reinterpret_cast<const void *>(&poa_group_kernel<V, true>) : reinterpret_cast<const void *>(&poa_group_kernel<V, false>);
}

static const void *kernel_of(int variant, bool trace) {
    switch (variant) {
        case 2: return variant_fn<2>(trace);
        case 4: return variant_fn<4>(trace);
        case 8: return variant_fn<8>(trace);
        default: return variant_fn<0>(trace);
    }
}

/* band capacity (cells per row) of a variant; variant 0 takes any wcap */
int variant_wcap(int variant, int wcap) { return variant == 0 ? wcap : 64 * variant; }

size_t poa_smem_bytes(int variant, int wcap, int warps_per_block) {
    int words;
    switch (variant) {
        case 2: words = variant_warp_words<2>(wcap); break;
        case 4: words = variant_warp_words<4>(wcap); break;
        case 8: words = variant_warp_words<8>(wcap); break;
        default: words = variant_warp_words<0>(wcap); break;
    }
    return (size_t)warps_per_block * words * sizeof(int);
}

cudaError_t launch_poa(int variant, const KernelArgs &A, int n_blocks, int warps_per_block, cudaStream_t stream) {
    const size_t smem = poa_smem_bytes(variant, A.wcap, warps_per_block);
    const void *fn = kernel_of(variant, A.tr_aln != nullptr);   // all five trace arrays are set together
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    void *args[] = {const_cast<KernelArgs *>(&A)};
    return cudaLaunchKernel(fn, dim3(n_blocks), dim3(warps_per_block * 32), args, smem, stream);
}

int poa_max_blocks_per_sm(int variant, int wcap, int warps_per_block) {
    int nb = 0;
    const size_t smem = poa_smem_bytes(variant, wcap, warps_per_block);
    const void *fn = kernel_of(variant, false);
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, warps_per_block * 32, smem) != cudaSuccess) return 0;
    return nb;
}

}  // namespace mpoa
