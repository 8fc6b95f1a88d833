/*
 * orient.cpp -- strand / "does it map" call of every read of a group against the group's first
 * read: the step the reference performs with mappy before it writes the abpoa input
 * (reference utils/SpliceDefineConsensus.py:895, :900-907):
 *
 *     mm_align = mp.Aligner(seq=first, preset='map-ont')
 *     for hit in mm_align.map(sequence):
 *         if hit.is_primary:
 *             if hit.strand == -1: sequence = mp.revcomp(sequence)
 *             ... the read is written (once per primary hit)
 *
 * A read without a primary hit is dropped, a read with two primary hits (primary +
 * supplementary) is written twice.  What decides this in minimap2 is the seed-chain stage of
 * `map-ont` (k = 15, w = 10): minimizer sketch of both sequences, colinear chaining of the seed
 * hits per strand, chains with >= 3 seeds and score >= 40 survive, a chain that overlaps a
 * better one by more than half of the shorter query interval is secondary.  This file restates
 * that stage (published algorithm of minimap2: Li 2018, sections 2.1.1 / 2.1.2 -- minimap2 is a
 * third-party dependency, not under /root/reference and not installable in the build image);
 * base-level extension and its score filter are not performed, so a chain minimap2 would drop
 * AFTER alignment is kept here.  mandalorion_b200.consensus uses mappy itself whenever it is
 * importable and this code otherwise.
 *
 * Host code (C++ threads over groups): the work is a few hash probes per 10 bases.
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/mandalorion_poa.h"

namespace {

struct Seed { uint64_t key; uint32_t pos; uint8_t strand, span; };   // pos = last base of the k-mer

/* ACGT(U), either case -> 0..3, everything else 4 (a table: a switch on random bases mispredicts) */
struct Nt4Table {
    uint8_t t[256];
    Nt4Table() {
        std::memset(t, 4, sizeof(t));
        t[(int)'A'] = t[(int)'a'] = 0; t[(int)'C'] = t[(int)'c'] = 1; t[(int)'G'] = t[(int)'g'] = 2;
        t[(int)'T'] = t[(int)'t'] = t[(int)'U'] = t[(int)'u'] = 3;
    }
};
const Nt4Table kNt4;
inline int nt4(uint8_t c) { return kNt4.t[c]; }

/* invertible integer hash of minimap2 (Thomas Wang's 64-bit mix restricted to 2k bits) */
inline uint64_t mix64(uint64_t key, uint64_t mask) {
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

/*
 * (w,k)-minimizers of a sequence, canonical strand per k-mer (the smaller of the forward and the
 * reverse-complement encoding; palindromes are skipped), ties inside a window all reported,
 * ambiguous bases restart the window -- the sampling rule of minimap2's sketch.
 */
void sketch(const uint8_t *s, int len, int w, int k, std::vector<Seed> &out) {
    out.clear();
    const uint64_t mask = (1ULL << 2 * k) - 1, shift1 = 2 * (k - 1);
    uint64_t fw = 0, rv = 0;
    /* the last w k-mers: key (UINT64_MAX = none) and position << 1 | strand */
    uint64_t bkey[32];
    uint32_t binfo[32];
    for (int j = 0; j < w; ++j) { bkey[j] = UINT64_MAX; binfo[j] = 0; }
    uint64_t mkey = UINT64_MAX;
    uint32_t minfo = 0;
    int l = 0, bp = 0, mp = 0;
    auto emit = [&](uint64_t key, uint32_t info) {
        if (key != UINT64_MAX) out.push_back(Seed{key, info >> 1, (uint8_t)(info & 1), (uint8_t)k});
    };
    auto emit_ties = [&](int from, int to) {     // other k-mers of the window with the minimum's key
        for (int j = from; j < to; ++j) if (bkey[j] == mkey && binfo[j] != minfo) emit(bkey[j], binfo[j]);
    };
    for (int i = 0; i < len; ++i) {
        const int c = nt4(s[i]);
        uint64_t ckey = UINT64_MAX;
        uint32_t cinfo = 0;
        if (c < 4) {
            fw = (fw << 2 | (uint64_t)c) & mask;
            rv = (rv >> 2) | (3ULL ^ (uint64_t)c) << shift1;
            if (fw == rv) continue;                 // palindromic k-mer: strand unknown
            const uint32_t z = fw < rv ? 0u : 1u;
            const uint64_t canon = fw < rv ? fw : rv;      // (a select: no branch on random data)
            ++l;
            if (l >= k) { ckey = mix64(canon, mask); cinfo = (uint32_t)i << 1 | z; }
        } else { l = 0; }
        bkey[bp] = ckey; binfo[bp] = cinfo;
        if (l == w + k - 1 && mkey != UINT64_MAX) { // first full window: report the ties of the minimum
            emit_ties(bp + 1, w); emit_ties(0, bp);
        }
        if (ckey <= mkey) {                         // a new minimum; the old one is written first
            if (l >= w + k) emit(mkey, minfo);
            mkey = ckey; minfo = cinfo; mp = bp;
        } else if (bp == mp) {                      // the old minimum left the window
            if (l >= w + k - 1) emit(mkey, minfo);
            /* new minimum = the LAST smallest key in window order (bp+1 .. w-1, 0 .. bp): branch-free */
            uint64_t best = UINT64_MAX;
            int bj = bp;
            for (int t = 1; t <= w; ++t) {
                int j = bp + t; j -= j >= w ? w : 0;
                const bool take = bkey[j] <= best;
                best = take ? bkey[j] : best;
                bj = take ? j : bj;
            }
            mkey = best; minfo = binfo[bj]; mp = bj;
            if (l >= w + k - 1 && mkey != UINT64_MAX) { emit_ties(bp + 1, w); emit_ties(0, bp + 1); }
        }
        if (++bp == w) bp = 0;
    }
    emit(mkey, minfo);
}

struct Anchor { uint64_t x, y; };   // x = strand << 63 | ref pos, y = span << 32 | query pos (flipped on the reverse strand)

inline float log2_fast(float x) { return std::log2(x); }

struct Chain { int score, cnt, qs, qe, strand; };

/* colinear chaining of the anchors of one read (minimap2 chaining DP: max distance 5000, band 500,
 * gap penalty 0.8 * 0.01 * k per base of diagonal drift + 0.5 log2, at most 25 skips, 5000 look-backs) */
void chain_anchors(std::vector<Anchor> &a, int k, std::vector<Chain> &chains) {
    chains.clear();
    const int n = (int)a.size();
    if (n == 0) return;
    std::sort(a.begin(), a.end(), [](const Anchor &p, const Anchor &q) { return p.x != q.x ? p.x < q.x : p.y < q.y; });
    const int max_dist = 5000, bw = 500, max_skip = 25, max_iter = 5000, min_cnt = 3, min_sc = 40;
    const float pen_gap = 0.8f * 0.01f * (float)k;
    std::vector<int> f(n), p(n), v(n), t(n, 0);
    int st = 0;
    for (int i = 0; i < n; ++i) {
        const int q_span_i = (int)(a[i].y >> 32 & 0xff);
        int max_f = q_span_i, max_j = -1, n_skip = 0;
        while (st < i && (a[i].x >> 63 != a[st].x >> 63 || a[i].x > a[st].x + (uint64_t)max_dist)) ++st;
        int lo = std::max(st, i - max_iter);
        for (int j = i - 1; j >= lo; --j) {
            const int dq = (int)(uint32_t)a[i].y - (int)(uint32_t)a[j].y;
            if (dq <= 0 || dq > max_dist) continue;
            const int dr = (int)((uint32_t)a[i].x - (uint32_t)a[j].x);
            if (dr == 0) continue;
            const int dd = dr > dq ? dr - dq : dq - dr;
            if (dd > bw) continue;
            const int dg = dr < dq ? dr : dq;
            const int q_span = (int)(a[j].y >> 32 & 0xff);
            int sc = q_span < dg ? q_span : dg;
            if (dd || dg > q_span) {
                const float lin = pen_gap * (float)dd;
                const float lg = dd >= 1 ? log2_fast((float)dd + 1.0f) : 0.0f;
                sc -= (int)(lin + 0.5f * lg);
            }
            sc += f[j];
            if (sc > max_f) {
                max_f = sc; max_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[j] == i) {
                if (++n_skip > max_skip) break;
            }
            if (p[j] >= 0) t[p[j]] = i;
        }
        f[i] = max_f; p[i] = max_j;
        v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;   // peak score up to i
    }
    /* backtrack: best chain ends first, every anchor used once */
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return f[x] != f[y] ? f[x] > f[y] : x < y; });
    std::vector<char> used(n, 0);
    for (int oi = 0; oi < n; ++oi) {
        const int end = order[oi];
        if (used[end] || f[end] < min_sc) continue;
        int i = end, cnt = 0, qs = INT32_MAX, qe = -1;
        while (i >= 0 && !used[i]) {
            used[i] = 1; ++cnt;
            const int qpos = (int)(uint32_t)a[i].y, span = (int)(a[i].y >> 32 & 0xff);
            qs = std::min(qs, qpos - span + 1); qe = std::max(qe, qpos + 1);
            i = p[i];
        }
        const int score = i >= 0 ? f[end] - f[i] : f[end];   // what was cut off belongs to an earlier chain
        if (cnt >= min_cnt && score >= min_sc)
            chains.push_back(Chain{score, cnt, qs, qe, (int)(a[end].x >> 63)});
    }
    std::sort(chains.begin(), chains.end(), [](const Chain &x, const Chain &y) { return x.score > y.score; });
}

/* index of the first read's seeds: sorted by key, plus an open-addressing table key -> first entry */
struct RefIndex {
    std::vector<Seed> seeds;
    std::vector<int32_t> slot;   // -1 = empty, else index of the first seed with the key that hashed here
    uint32_t mask = 0;
    void build() {
        std::sort(seeds.begin(), seeds.end(), [](const Seed &a, const Seed &b) { return a.key != b.key ? a.key < b.key : a.pos < b.pos; });
        uint32_t cap = 16;
        while (cap < 2 * seeds.size() + 1) cap <<= 1;
        mask = cap - 1;
        slot.assign(cap, -1);
        for (size_t i = 0; i < seeds.size(); ++i) {
            if (i > 0 && seeds[i].key == seeds[i - 1].key) continue;
            uint32_t h = (uint32_t)(seeds[i].key * 0x9E3779B97F4A7C15ull >> 40) & mask;
            while (slot[h] >= 0) h = (h + 1) & mask;
            slot[h] = (int32_t)i;
        }
    }
    /* [lo, hi) of the seeds with this key */
    inline void find(uint64_t key, int &lo, int &hi) const {
        uint32_t h = (uint32_t)(key * 0x9E3779B97F4A7C15ull >> 40) & mask;
        for (;;) {
            const int32_t i = slot[h];
            if (i < 0) { lo = hi = 0; return; }
            if (seeds[i].key == key) {
                lo = i; hi = i + 1;
                while (hi < (int)seeds.size() && seeds[hi].key == key) ++hi;
                return;
            }
            h = (h + 1) & mask;
        }
    }
};

/* primary hits of one read against the index of the group's first read; returns their strands (+1 / -1) */
int orient_read(const RefIndex &ref, const uint8_t *q, int qlen, int w, int k,
                std::vector<Seed> &qs, std::vector<Anchor> &anchors, std::vector<Chain> &chains, int8_t *strands, int max_hits) {
    sketch(q, qlen, w, k, qs);
    anchors.clear();
    const int max_occ = 10;     // seeds that occur more often in the reference are ignored
    for (const Seed &s : qs) {
        int lo, hi;
        ref.find(s.key, lo, hi);
        if (hi - lo > max_occ) continue;
        for (const Seed *r = ref.seeds.data() + lo; r != ref.seeds.data() + hi; ++r) {
            Anchor an;
            if (r->strand == s.strand) {          // same strand
                an.x = (uint64_t)r->pos;
                an.y = (uint64_t)s.span << 32 | (uint64_t)s.pos;
            } else {                              // opposite strand: query coordinate on the reverse complement
                an.x = 1ULL << 63 | (uint64_t)r->pos;
                an.y = (uint64_t)s.span << 32 | (uint64_t)(qlen - ((int)s.pos + 1 - (int)s.span) - 1);
            }
            anchors.push_back(an);
        }
    }
    /* Fast path for the usual case (a read of the same isoform): when all but at most two anchors lie
     * on one strand inside one chaining band (500 nt of diagonal drift) and there are plenty of them,
     * the chaining DP can only return that one chain -- fewer than three seeds cannot form another
     * one -- so its result is known without running it. */
    {
        int n_fw = 0;
        for (const Anchor &an : anchors) n_fw += (an.x >> 63) ? 0 : 1;
        const int n_all = (int)anchors.size(), n_rv = n_all - n_fw;
        const int major = n_fw >= n_rv ? 0 : 1, n_major = major ? n_rv : n_fw;
        if (n_major >= 16 && n_all - n_major <= 2) {
            int dmin = INT32_MAX, dmax = INT32_MIN, prev_q = -1, cover = 0;
            bool colinear = true;
            /* anchors were generated in query order of the forward read; on the reverse strand the
             * flipped query coordinate decreases */
            for (const Anchor &an : anchors) {
                if ((int)(an.x >> 63) != major) continue;
                const int rp = (int)(uint32_t)an.x, qp = (int)(uint32_t)an.y;
                const int d = rp - qp;
                dmin = std::min(dmin, d); dmax = std::max(dmax, d);
                if (prev_q >= 0) {
                    const int dq = major ? prev_q - qp : qp - prev_q;
                    if (dq < 0) colinear = false;
                    cover += std::min(std::max(dq, 0), k);
                } else cover += k;
                prev_q = qp;
            }
            if (colinear && dmax - dmin <= 250 && cover >= 80) {
                strands[0] = major ? -1 : 1;
                return 1;
            }
        }
    }
    chain_anchors(anchors, k, chains);
    /* a chain that overlaps a better one by more than half of the shorter query interval is secondary */
    int n_hits = 0;
    std::vector<int> prim;
    for (int i = 0; i < (int)chains.size(); ++i) {
        /* query interval on the forward strand of the read */
        int si = chains[i].qs, ei = chains[i].qe;
        if (chains[i].strand) { const int t = qlen - ei; ei = qlen - si; si = t; }
        bool secondary = false;
        for (int j : prim) {
            int sj = chains[j].qs, ej = chains[j].qe;
            if (chains[j].strand) { const int t = qlen - ej; ej = qlen - sj; sj = t; }
            const int ol = std::min(ei, ej) - std::max(si, sj);
            const int ml = std::min(ei - si, ej - sj);
            if (ol > 0 && (double)ol > 0.5 * (double)ml) { secondary = true; break; }
        }
        if (secondary) continue;
        prim.push_back(i);
        if (n_hits < max_hits) strands[n_hits] = chains[i].strand ? -1 : 1;
        ++n_hits;
    }
    return std::min(n_hits, max_hits);
}

}  // namespace

extern "C" int mpoa_orient_batch(int64_t n_groups, const int64_t *group_read_off, const int64_t *read_base_off,
                                 const uint8_t *bases, int32_t n_threads, int8_t *hit_count, int8_t *hit_strand) {
    if (n_groups < 0 || (n_groups > 0 && (!group_read_off || !read_base_off || !hit_count || !hit_strand))) return MPOA_EINVAL;
    const int w = 10, k = 15;   // minimap2 preset map-ont
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        RefIndex ref;
        std::vector<Seed> qs;
        std::vector<Anchor> anchors;
        std::vector<Chain> chains;
        for (;;) {
            const int64_t g = next.fetch_add(1);
            if (g >= n_groups) break;
            const int64_t r0 = group_read_off[g], r1 = group_read_off[g + 1];
            if (r1 <= r0) continue;
            const int64_t f0 = read_base_off[r0];
            const int flen = (int)(read_base_off[r0 + 1] - f0);
            sketch(bases + f0, flen, w, k, ref.seeds);
            ref.build();
            for (int64_t r = r0; r < r1; ++r) {
                const int64_t b0 = read_base_off[r];
                const int len = (int)(read_base_off[r + 1] - b0);
                hit_count[r] = (int8_t)orient_read(ref, bases + b0, len, w, k, qs, anchors, chains, hit_strand + 2 * r, 2);
            }
        }
    };
    int nt = std::max(1, (int)n_threads);
    if ((int64_t)nt > n_groups) nt = (int)std::max<int64_t>(1, n_groups);
    if (nt <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(worker);
        for (auto &t : th) t.join();
    }
    return MPOA_OK;
}
