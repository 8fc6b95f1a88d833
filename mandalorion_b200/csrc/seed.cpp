/*
 * seed.cpp -- host side of `abpoa -S` (reference utils/SpliceDefineConsensus.py:916-919: median read
 * length >= 8000): the anchors of every read of a seeded group against the read before it.
 *
 * abPOA's seeding (upstream abpoa_seed.c, not available here) is summarised in SURVEY.md A.10:
 * (k = 19, w = 10) minimizers, hits between consecutive reads, colinear chaining, anchors at least
 * 500 nt apart.  The rules this library implements (include/mandalorion_poa.h, "Seeded groups"):
 *   minimizers  forward strand, minimap2 sampling rule (ties of a window all reported), N restarts
 *   hits        equal k-mer hash in read i-1 and read i; a hash that occurs more than 8 times in
 *               read i-1 is ignored; order (position in read i-1, position in read i)
 *   chain       best colinear chain over the 64 previous hits, diagonal drift <= 100, a step scores
 *               min(k, advance); first maximum wins
 *   anchors     along the chain, keep a hit whose k-mer starts >= 500 nt after the end of the last
 *               kept one (or the read start) on both reads
 * The kernels (poa_seed.cuh) turn every anchor into k forced matches and align the stretches between
 * anchors to the sub-graph between the anchor nodes.
 */
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <thread>
#include <vector>

namespace mpoa {

namespace {

struct KmerAt { uint64_t h; int32_t end; };

inline uint64_t hash_kmer(uint64_t x, uint64_t mask) {
    x = (~x + (x << 21)) & mask;
    x ^= x >> 24;
    x = (x + (x << 3) + (x << 8)) & mask;
    x ^= x >> 14;
    x = (x + (x << 2) + (x << 4)) & mask;
    x ^= x >> 28;
    x = (x + (x << 31)) & mask;
    return x;
}

struct CodeTable {
    uint8_t v[256];
    CodeTable() {
        for (int c = 0; c < 256; ++c) {
            switch (c & 0xdf) { case 'A': v[c] = 0; break; case 'C': v[c] = 1; break; case 'G': v[c] = 2; break; case 'T': v[c] = 3; break; default: v[c] = 4; }
        }
    }
};
const CodeTable kCodes;
inline int code_of(uint8_t c) { return kCodes.v[c]; }

/* forward-strand (w,k) minimizers of an ASCII read */
void forward_minimizers(const uint8_t *s, int n, int w, int k, std::vector<KmerAt> &out) {
    out.clear();
    const uint64_t mask = (1ULL << (2 * k)) - 1;
    uint64_t ring_h[256];
    int32_t ring_end[256];
    if (w > 256) w = 256;
    for (int j = 0; j < w; ++j) { ring_h[j] = UINT64_MAX; ring_end[j] = 0; }
    out.reserve((size_t)n / 4 + 16);
    uint64_t word = 0, cur_min = UINT64_MAX;
    int32_t cur_end = 0;
    int run = 0, slot = 0, min_slot = 0;
    auto put = [&](uint64_t h, int32_t e) { if (h != UINT64_MAX) out.push_back(KmerAt{h, e}); };
    auto put_equal = [&](int a, int b) { for (int j = a; j < b; ++j) if (ring_h[j] == cur_min && ring_end[j] != cur_end) put(ring_h[j], ring_end[j]); };
    for (int i = 0; i < n; ++i) {
        const int c = code_of(s[i]);
        uint64_t h = UINT64_MAX;
        if (c < 4) {
            word = ((word << 2) | (uint64_t)c) & mask;
            if (++run >= k) h = hash_kmer(word, mask);
        } else run = 0;
        ring_h[slot] = h; ring_end[slot] = i;
        if (run == w + k - 1 && cur_min != UINT64_MAX) { put_equal(slot + 1, w); put_equal(0, slot); }
        if (h <= cur_min) {
            if (run >= w + k) put(cur_min, cur_end);
            cur_min = h; cur_end = i; min_slot = slot;
        } else if (slot == min_slot) {
            if (run >= w + k - 1) put(cur_min, cur_end);
            uint64_t best = UINT64_MAX;
            int where = slot;
            for (int j = slot + 1; j < w; ++j)          // oldest to newest: slot+1 .. w-1, then 0 .. slot
                if (ring_h[j] <= best) { best = ring_h[j]; where = j; }
            for (int j = 0; j <= slot; ++j)
                if (ring_h[j] <= best) { best = ring_h[j]; where = j; }
            cur_min = best; cur_end = ring_end[where]; min_slot = where;
            if (run >= w + k - 1 && cur_min != UINT64_MAX) { put_equal(slot + 1, w); put_equal(0, slot + 1); }
        }
        slot = slot + 1 == w ? 0 : slot + 1;
    }
    put(cur_min, cur_end);
}

struct Pair { int32_t t, q; };

/* The chaining step of hit i: among hits [first, i) the one that gives the best score (colinear, diagonal
 * drift <= 100, a step scores min(k, advance)); it must beat k, the score of starting a chain at i; of
 * equally good ones the nearest.  Candidate scores first (a branch-free loop the compiler vectorises: this is
 * where the seeding spends its time), then the pick.  Returns the index or -1. */
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target_clones("avx2", "default")))
#endif
int best_link(const int32_t *ht, const int32_t *hq, const int32_t *score, int first, int i, int k) {
    int32_t cand[64];
    const int m = i - first, ti = ht[i], qi = hq[i];
    for (int x = 0; x < m; ++x) {
        const int dt = ti - ht[first + x], dq = qi - hq[first + x];
        const int drift = dt > dq ? dt - dq : dq - dt;
        const int step = dt < dq ? dt : dq;
        const bool ok = (dt > 0) & (dq > 0) & (drift <= 100);
        cand[x] = ok ? score[first + x] + (step < k ? step : k) : INT32_MIN;
    }
    int best = INT32_MIN;
    for (int x = 0; x < m; ++x) best = cand[x] > best ? cand[x] : best;
    if (best <= k) return -1;
    for (int x = m - 1; x >= 0; --x)                   // the nearest of the best: almost always one of the last few
        if (cand[x] == best) return first + x;
    return -1;
}

/* the minimizers of the previous read by hash: open addressing, up to 9 positions per hash kept in order
 * of position (9 = "more than 8": such a hash is ignored) */
struct KmerTable {
    struct Slot { uint64_t h; int32_t n; int32_t end[9]; };
    std::vector<Slot> slot;
    uint64_t mask = 0;
    void build(const std::vector<KmerAt> &ms) {
        size_t cap = 64;
        while (cap < 2 * ms.size() + 2) cap <<= 1;
        slot.resize(cap);
        mask = cap - 1;
        for (Slot &x : slot) { x.h = UINT64_MAX; x.n = 0; }
        for (const KmerAt &m : ms) {                    // ms is in order of position
            uint64_t i = (m.h * 0x9E3779B97F4A7C15ULL >> 20) & mask;
            while (slot[i].h != UINT64_MAX && slot[i].h != m.h) i = (i + 1) & mask;
            Slot &x = slot[i];
            x.h = m.h;
            if (x.n < 9) x.end[x.n] = m.end;
            if (x.n < 9) ++x.n;
        }
    }
    const Slot *find(uint64_t h) const {
        uint64_t i = (h * 0x9E3779B97F4A7C15ULL >> 20) & mask;
        while (slot[i].h != UINT64_MAX) {
            if (slot[i].h == h) return &slot[i];
            i = (i + 1) & mask;
        }
        return nullptr;
    }
};

void anchors_of(const KmerTable &prev, const std::vector<KmerAt> &cur, int k, int min_gap,
                std::vector<Pair> &hits, std::vector<int32_t> &score, std::vector<int32_t> &from, std::vector<int32_t> &ht,
                std::vector<int32_t> &hq, int32_t *out, int cap, int32_t *n_out) {
    hits.clear();
    for (const KmerAt &m : cur) {
        const KmerTable::Slot *x = prev.find(m.h);
        if (!x || x->n > 8) continue;
        for (int j = 0; j < x->n; ++j) hits.push_back(Pair{x->end[j], m.end});
    }
    std::sort(hits.begin(), hits.end(), [](const Pair &a, const Pair &b) { return a.t != b.t ? a.t < b.t : a.q < b.q; });
    const int n = (int)hits.size();
    if (n == 0) return;
    score.assign(n, k); from.assign(n, -1);
    ht.resize(n); hq.resize(n);
    for (int i = 0; i < n; ++i) { ht[i] = hits[i].t; hq[i] = hits[i].q; }
    int top = 0;
    for (int i = 0; i < n; ++i) {
        const int first = std::max(0, i - 64);
        const int j = best_link(ht.data(), hq.data(), score.data(), first, i, k);
        if (j >= 0) { from[i] = j; score[i] = score[j] + std::min(k, std::min(ht[i] - ht[j], hq[i] - hq[j])); }
        if (score[i] > score[top]) top = i;
    }
    std::vector<int32_t> path;
    for (int i = top; i >= 0; i = from[i]) path.push_back(i);
    int32_t end_t = -1, end_q = -1;
    for (auto it = path.rbegin(); it != path.rend(); ++it) {
        const int32_t t0 = hits[*it].t - k + 1, q0 = hits[*it].q - k + 1;
        if (t0 - (end_t + 1) >= min_gap && q0 - (end_q + 1) >= min_gap) {
            if (*n_out >= cap) break;                  // cannot happen: anchors are >= min_gap apart (cap = len / min_gap + 1)
            out[2 * *n_out] = t0; out[2 * *n_out + 1] = q0;
            ++*n_out;
            end_t = hits[*it].t; end_q = hits[*it].q;
        }
    }
}

}  // namespace

/*
 * Anchors of the reads of the groups order[0..n_order) (the flagged groups, in the order the device will ask
 * for them).  Read r owns the slots anc[2 * anc_off[r] ...] of the arena: slot 0 = (number of anchors, 0),
 * then up to len / min_gap + 1 (start in the previous read, start in this read) pairs -- anchors are at least
 * min_gap apart, so the regions can be laid out before the anchors exist.  n_threads host threads take the
 * groups in order; done[i / chunk_size] counts the finished groups of every chunk, so that the consumer can
 * publish chunk 0 while the later ones are still being computed.  src_off (nullable): where read r starts in
 * `bases` when the batch is a subset of the caller's arrays (default: rbo[r]).
 */
void seed_batch(const int64_t *gro, const int64_t *rbo, const uint8_t *bases, const int64_t *src_off,
                const int32_t *order, int64_t n_order, int64_t chunk_size, std::atomic<int64_t> *done,
                int k, int w, int min_gap, int n_threads, const int32_t *anc_off, int32_t *anc) {
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        std::vector<KmerAt> cur;
        KmerTable prev;
        std::vector<Pair> hits;
        std::vector<int32_t> score, from, ht, hq;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_order) break;
            const int64_t g = order[i];
            int64_t prev_r = -1;
            for (int64_t r = gro[g]; r < gro[g + 1]; ++r) {
                const int len = (int)(rbo[r + 1] - rbo[r]);
                int32_t *region = anc + 2 * (int64_t)anc_off[r];
                region[0] = 0; region[1] = 0;
                if (len <= 0) continue;                 // an empty read is skipped by the aligner too
                forward_minimizers(bases + (src_off ? src_off[r] : rbo[r]), len, w, k, cur);
                if (prev_r >= 0)
                    anchors_of(prev, cur, k, min_gap, hits, score, from, ht, hq, region + 2, len / min_gap + 1, region);
                prev.build(cur);
                prev_r = r;
            }
            if (done) done[chunk_size > 0 ? i / chunk_size : 0].fetch_add(1, std::memory_order_release);
        }
    };
    int nt = std::max(1, n_threads);
    if ((int64_t)nt > n_order) nt = (int)std::max<int64_t>(1, n_order);
    if (nt <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(worker);
        for (auto &t : th) t.join();
    }
}

}  // namespace mpoa
