/*
 * abpoa_oracle.cpp -- CPU oracle: scalar restatement of `abpoa -M 5 -r 0 <in.fasta>`
 * (abPOA v1.4.1) as Mandalorion invokes it (reference utils/SpliceDefineConsensus.py:917;
 * version pinned by reference setup.sh:17-19, Mando.py:257, README.md:79).
 *
 * TEST INFRASTRUCTURE ONLY (see mpoa_oracle.h).  PARITY UNPINNED: abPOA's source is a
 * third-party dependency absent from /root/reference and from this image; every function
 * below restates the published algorithm of the upstream file named in its comment
 * (upstream paths are abPOA-v1.4.1/src/...), following SURVEY.md Appendix A.
 *
 * The restatement keeps abPOA's observable conventions, not its SIMD code:
 *   - node 0 = source, node 1 = sink, edge arrays in first-creation order (A.3, A.8);
 *   - BFS topological order with aligned-node groups enqueued together, max_remain via
 *     reverse BFS over the heaviest out-edge (A.5);
 *   - adaptive band rounded to whole SIMD vectors of `pn` lanes, clamped by the
 *     predecessors' vector ranges; the diagonal is NOT carried across the first vector
 *     of the overlap with a predecessor (an artefact of the shift-in of -inf) (A.5);
 *   - convex-gap recurrences with abPOA's finite "-inf" and saturating int16 lanes (A.6);
 *   - value-comparison traceback with the M > E1/E2 (per predecessor) > F1 > F2 priority
 *     and the open/extend state machine (A.7);
 *   - add-alignment with aligned-node cliques (A.8) and heaviest bundling (A.9).
 */
#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdint>
#include <cstring>
#include <deque>
#include <thread>
#include <vector>

#include "mpoa_oracle.h"

namespace {

enum { OP_M = 1, OP_E1 = 2, OP_E2 = 4, OP_E = 6, OP_F1 = 8, OP_F2 = 16, OP_F = 24, OP_ALL = 31 };
enum { SRC = 0, SINK = 1 };

struct Par {
    int m = 5;
    int mat[25];
    int match, mismatch;  // mismatch stored positive
    int o1, e1, o2, e2, oe1, oe2;
    int wb;
    float wf;
    int pn16, pn32;
    mpoa_oracle_opts opt;
};

/* upstream abpoa_align.c: gen_simple_mat() -- N (code 4) scores 0 against everything (A.2) */
void make_matrix(Par &P) {
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j)
            P.mat[i * 5 + j] = (i == 4 || j == 4) ? 0 : (i == j ? P.match : -P.mismatch);
}

/* upstream abpoa_seq.c: ab_nt4_table -- ACGT (either case) -> 0..3, everything else 4 */
inline uint8_t nt4(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

struct Node {
    uint8_t base = 0;
    int creator = -1;
    std::vector<int> in_id, out_id, out_w, aligned;
};

struct Graph {
    std::vector<Node> node;
    bool sorted = false;
    std::vector<int> index_to_node_id, node_id_to_index, max_remain, max_pos_left, max_pos_right;
    Graph() { node.resize(2); }
    int node_n() const { return (int)node.size(); }
};

/* upstream abpoa_graph.c: abpoa_add_graph_node() */
int add_node(Graph &g, uint8_t base, int creator) {
    g.node.emplace_back();
    g.node.back().base = base;
    g.node.back().creator = creator;
    return g.node_n() - 1;
}

/* upstream abpoa_graph.c: abpoa_add_graph_edge(); check_edge=0 appends without search */
void add_edge(Graph &g, int from, int to, int check_edge, int w) {
    Node &f = g.node[from];
    if (check_edge) {
        for (size_t i = 0; i < f.out_id.size(); ++i)
            if (f.out_id[i] == to) { f.out_w[i] += w; return; }
    }
    g.node[to].in_id.push_back(from);
    f.out_id.push_back(to);
    f.out_w.push_back(w);
}

/* upstream abpoa_graph.c: abpoa_add_graph_aligned_node() -- clique over the column (A.8) */
void add_aligned(Graph &g, int node_id, int aligned_id) {
    std::vector<int> sib = g.node[node_id].aligned;
    for (int s : sib) {
        g.node[s].aligned.push_back(aligned_id);
        g.node[aligned_id].aligned.push_back(s);
    }
    g.node[node_id].aligned.push_back(aligned_id);
    g.node[aligned_id].aligned.push_back(node_id);
}

/* upstream abpoa_graph.c: abpoa_get_aligned_id() */
int get_aligned_id(const Graph &g, int node_id, uint8_t base) {
    for (int a : g.node[node_id].aligned)
        if (g.node[a].base == base) return a;
    return -1;
}

/* upstream abpoa_graph.c: abpoa_BFS_set_node_index() (A.5) */
bool bfs_set_node_index(Graph &g) {
    int n = g.node_n();
    std::vector<int> in_degree(n);
    for (int i = 0; i < n; ++i) in_degree[i] = (int)g.node[i].in_id.size();
    std::deque<int> q;
    q.push_back(SRC);
    int index = 0;
    while (!q.empty()) {
        int cur = q.front();
        q.pop_front();
        g.index_to_node_id[index] = cur;
        g.node_id_to_index[cur] = index++;
        if (cur == SINK) return true;
        for (int out_id : g.node[cur].out_id) {
            if (--in_degree[out_id] == 0) {
                bool ready = true;
                for (int a : g.node[out_id].aligned)
                    if (in_degree[a] != 0) { ready = false; break; }
                if (!ready) continue;
                q.push_back(out_id);
                for (int a : g.node[out_id].aligned) q.push_back(a);
            }
        }
    }
    return false;  // abPOA: err_fatal("Failed to set node index")
}

/* upstream abpoa_graph.c: abpoa_BFS_set_node_remain() (A.5) */
bool bfs_set_node_remain(Graph &g) {
    int n = g.node_n();
    std::vector<int> out_degree(n);
    for (int i = 0; i < n; ++i) {
        out_degree[i] = (int)g.node[i].out_id.size();
        g.max_remain[i] = 0;
    }
    std::deque<int> q;
    q.push_back(SINK);
    g.max_remain[SINK] = -1;
    while (!q.empty()) {
        int cur = q.front();
        q.pop_front();
        if (cur != SINK) {
            int max_w = -1, max_id = -1;
            const Node &nd = g.node[cur];
            for (size_t i = 0; i < nd.out_id.size(); ++i)
                if (nd.out_w[i] > max_w) { max_w = nd.out_w[i]; max_id = nd.out_id[i]; }
            g.max_remain[cur] = g.max_remain[max_id] + 1;
        }
        if (cur == SRC) return true;
        for (int in_id : g.node[cur].in_id)
            if (--out_degree[in_id] == 0) q.push_back(in_id);
    }
    return false;
}

/* upstream abpoa_graph.c: abpoa_topological_sort() */
bool topological_sort(Graph &g) {
    int n = g.node_n();
    g.index_to_node_id.assign(n, 0);
    g.node_id_to_index.assign(n, 0);
    g.max_remain.assign(n, 0);
    if (!bfs_set_node_index(g)) return false;
    g.max_pos_right.assign(n, 0);
    g.max_pos_left.assign(n, n);
    if (!bfs_set_node_remain(g)) return false;
    g.sorted = true;
    return true;
}

struct Cig { int op; int node_id; int qidx; };  // op: 0 match, 1 ins, 2 del

struct AlnOut {
    std::vector<Cig> cigar;  // forward order
    int best_score = 0;
    int bits = 0;
    int64_t band_cells = 0, int_ops = 0, full_cells = 0;
    int max_width = 0;
    bool ok = false;
};

struct DpScratch {
    std::vector<int> dp_beg, dp_end, dp_beg_sn, dp_end_sn, cell_hi;
    std::vector<int64_t> off;
    std::vector<int32_t> H, E1, E2, F1, F2;
    std::vector<int> pre_off, pre_idx;
    std::vector<int32_t> mx, ei1, ei2;
};

inline int32_t sat(int64_t x, int32_t lo) { return x < lo ? lo : (int32_t)x; }

/*
 * upstream simd_abpoa_align.c: simd_abpoa_align_sequence_to_subgraph() with
 * abpoa_cg_global_align_sequence_to_graph_core (first row, per-row DP, best cell,
 * cg backtrack) for the whole graph in global mode, adaptive band on.
 */
void align_to_graph(Graph &g, const Par &P, const uint8_t *q, int qlen, DpScratch &S, AlnOut &R) {
    R.ok = false;
    R.cigar.clear();
    const int gn = g.node_n();
    /* lane width decision (A.4) */
    const int len = qlen > gn ? qlen : gn;
    const int64_t max_score = std::max<int64_t>((int64_t)qlen * P.match, (int64_t)len * P.e1 + P.o1);
    const bool is16 = max_score <= INT16_MAX - P.mismatch - P.o1 - P.e1 - P.o2 - P.e2;
    const int pn = is16 ? P.pn16 : P.pn32;
    const int32_t type_min = is16 ? INT16_MIN : INT32_MIN;
    const int32_t inf_min = (int32_t)(std::max({(int64_t)type_min + P.mismatch, (int64_t)type_min + P.oe1,
                                                (int64_t)type_min + P.oe2}) +
                                      31 * std::max(P.e1, P.e2));
    R.bits = is16 ? 16 : 32;
    const int w = P.wb < 0 ? qlen : P.wb + (int)(P.wf * (float)qlen);  // float product, as in C

    /* predecessor rows in in_id order */
    S.pre_off.assign(gn + 1, 0);
    S.pre_idx.clear();
    for (int i = 0; i < gn; ++i) {
        const Node &nd = g.node[g.index_to_node_id[i]];
        S.pre_off[i] = (int)S.pre_idx.size();
        for (int in_id : nd.in_id) S.pre_idx.push_back(g.node_id_to_index[in_id]);
    }
    S.pre_off[gn] = (int)S.pre_idx.size();

    S.dp_beg.assign(gn, 0); S.dp_end.assign(gn, -1); S.dp_beg_sn.assign(gn, 0); S.dp_end_sn.assign(gn, -1);
    S.cell_hi.assign(gn, -1);
    S.off.assign(gn, 0);
    S.H.clear(); S.E1.clear(); S.E2.clear(); S.F1.clear(); S.F2.clear();

    auto band_beg = [&](int node_id) {
        return std::max(0, std::min(g.max_pos_left[node_id], qlen - g.max_remain[node_id]) - w);
    };
    auto band_end = [&](int node_id) {
        return std::min(qlen, std::max(g.max_pos_right[node_id], qlen - g.max_remain[node_id]) + w);
    };

    /* first row (source) */
    {
        g.max_pos_left[SRC] = g.max_pos_right[SRC] = 0;
        for (int out_id : g.node[SRC].out_id) g.max_pos_left[out_id] = g.max_pos_right[out_id] = 1;
        int b = band_beg(SRC), e = band_end(SRC);
        S.dp_beg_sn[0] = b / pn; S.dp_end_sn[0] = e / pn;
        S.dp_beg[0] = S.dp_beg_sn[0] * pn; S.dp_end[0] = (S.dp_end_sn[0] + 1) * pn - 1;
        int hi = std::min(S.dp_end[0], qlen);
        S.cell_hi[0] = hi;
        S.off[0] = 0;
        int width = hi - S.dp_beg[0] + 1;
        S.H.assign(width, inf_min); S.E1.assign(width, inf_min); S.E2.assign(width, inf_min);
        S.F1.assign(width, inf_min); S.F2.assign(width, inf_min);
        // dp_beg[0] is 0 by construction
        S.H[0] = 0; S.E1[0] = -P.oe1; S.E2[0] = -P.oe2;
        for (int i = 1; i <= hi; ++i) {
            S.F1[i] = -(P.o1 + P.e1 * i);
            S.F2[i] = -(P.o2 + P.e2 * i);
            S.H[i] = std::max(S.F1[i], S.F2[i]);
        }
    }

    /* rows 1 .. gn-2 (the sink row is never computed) */
    for (int index_i = 1; index_i < gn - 1; ++index_i) {
        const int node_id = g.index_to_node_id[index_i];
        const int *pre = &S.pre_idx[S.pre_off[index_i]];
        const int npre = S.pre_off[index_i + 1] - S.pre_off[index_i];
        int beg = band_beg(node_id), end = band_end(node_id);
        int beg_sn = beg / pn, end_sn = end / pn;
        int min_pre_beg_sn = INT_MAX, max_pre_end_sn = -1;
        for (int k = 0; k < npre; ++k) {
            min_pre_beg_sn = std::min(min_pre_beg_sn, S.dp_beg_sn[pre[k]]);
            max_pre_end_sn = std::max(max_pre_end_sn, S.dp_end_sn[pre[k]]);
        }
        if (beg_sn < min_pre_beg_sn) beg_sn = min_pre_beg_sn;
        if (P.opt.clamp_end_to_pred && end_sn > max_pre_end_sn + 1) end_sn = max_pre_end_sn + 1;
        S.dp_beg_sn[index_i] = beg_sn; S.dp_end_sn[index_i] = end_sn;
        const int dp_beg = S.dp_beg[index_i] = beg_sn * pn;
        const int dp_end = S.dp_end[index_i] = (end_sn + 1) * pn - 1;
        const int hi_cell = std::min(dp_end, qlen);
        S.cell_hi[index_i] = hi_cell;
        const int width = std::max(0, hi_cell - dp_beg + 1);
        const int64_t off = (int64_t)S.H.size();
        S.off[index_i] = off;
        S.H.resize(off + width); S.E1.resize(off + width); S.E2.resize(off + width);
        S.F1.resize(off + width); S.F2.resize(off + width);
        R.band_cells += width;
        if (width > R.max_width) R.max_width = width;
        R.int_ops += 17LL * width + 3LL * std::max(0, npre - 1) * width;
        R.full_cells += qlen + 1;

        int max = inf_min, left_max_i = -1, right_max_i = -1;
        if (width > 0) {
            S.mx.assign(width, inf_min); S.ei1.assign(width, inf_min); S.ei2.assign(width, inf_min);
            for (int k = 0; k < npre; ++k) {
                const int p = pre[k];
                const int lo = std::max(beg_sn, S.dp_beg_sn[p]) * pn;
                const int hi = std::min((std::min(end_sn, S.dp_end_sn[p]) + 1) * pn - 1, qlen);
                const int32_t *Hp = &S.H[S.off[p]] - S.dp_beg[p];
                const int32_t *E1p = &S.E1[S.off[p]] - S.dp_beg[p];
                const int32_t *E2p = &S.E2[S.off[p]] - S.dp_beg[p];
                for (int j = lo; j <= hi; ++j) {
                    const int32_t mv = (j == lo) ? inf_min : Hp[j - 1];
                    const int c = j - dp_beg;
                    S.mx[c] = std::max(S.mx[c], mv);
                    S.ei1[c] = std::max(S.ei1[c], E1p[j]);
                    S.ei2[c] = std::max(S.ei2[c], E2p[j]);
                }
            }
            const int *mrow = &P.mat[g.node[node_id].base * 5];
            int32_t *Hr = &S.H[off], *E1r = &S.E1[off], *E2r = &S.E2[off], *F1r = &S.F1[off], *F2r = &S.F2[off];
            /* pass 1 (independent cells): Hhat = max(M + s, Ein1, Ein2), kept in Hr for now */
            for (int c = 0; c < width; ++c) {
                const int j = dp_beg + c;
                const int s = j > 0 ? mrow[q[j - 1]] : 0;
                const int32_t m = S.mx[c] < type_min - std::min(s, 0) ? type_min : S.mx[c] + s;   // saturating add
                Hr[c] = std::max(m, std::max(S.ei1[c], S.ei2[c]));
            }
            /* pass 2 (serial along the row): F[c] = max(Hhat[c-1] - oe, F[c-1] - e), saturating */
            {
                auto ssub = [type_min](int32_t x, int32_t d) { return x < type_min + d ? type_min : x - d; };
                int32_t f1 = ssub(inf_min, P.oe1), f2 = ssub(inf_min, P.oe2);
                F1r[0] = f1; F2r[0] = f2;
                for (int c = 1; c < width; ++c) {
                    f1 = std::max(ssub(Hr[c - 1], P.oe1), ssub(f1, P.e1));
                    f2 = std::max(ssub(Hr[c - 1], P.oe2), ssub(f2, P.e2));
                    F1r[c] = f1; F2r[c] = f2;
                }
            }
            /* pass 3 (independent cells): H, Eout */
            for (int c = 0; c < width; ++c) {
                const int32_t h = std::max(Hr[c], std::max(F1r[c], F2r[c]));
                Hr[c] = h;
                const int32_t a1 = S.ei1[c] < type_min + P.e1 ? type_min : S.ei1[c] - P.e1;
                const int32_t b1 = h < type_min + P.oe1 ? type_min : h - P.oe1;
                const int32_t a2 = S.ei2[c] < type_min + P.e2 ? type_min : S.ei2[c] - P.e2;
                const int32_t b2 = h < type_min + P.oe2 ? type_min : h - P.oe2;
                E1r[c] = std::max(a1, b1);
                E2r[c] = std::max(a2, b2);
            }
            /* row maximum with its left-most and right-most column */
            for (int c = 0; c < width; ++c) max = std::max(max, Hr[c]);
            if (max > inf_min) {
                for (int c = 0; c < width; ++c) if (Hr[c] == max) { left_max_i = dp_beg + c; break; }
                if (P.opt.single_argmax) right_max_i = left_max_i;
                else for (int c = width - 1; c >= 0; --c) if (Hr[c] == max) { right_max_i = dp_beg + c; break; }
            } else if (!P.opt.single_argmax) {
                /* nothing above -inf: abPOA's `else if (h == max) right = j` still fires on cells equal to it */
                for (int c = width - 1; c >= 0; --c) if (Hr[c] == max) { right_max_i = dp_beg + c; break; }
            }
        }
        for (int out_id : g.node[node_id].out_id) {
            if (right_max_i + 1 > g.max_pos_right[out_id]) g.max_pos_right[out_id] = right_max_i + 1;
            if (left_max_i + 1 < g.max_pos_left[out_id]) g.max_pos_left[out_id] = left_max_i + 1;
        }
    }

    /* best end cell: sink's in-neighbours in in_id order, strictly greater wins (A.6) */
    auto in_band = [&](int row, int j) { return j >= S.dp_beg[row] && j <= S.cell_hi[row]; };
    auto val = [&](const std::vector<int32_t> &A, int row, int j) -> int32_t {
        return in_band(row, j) ? A[S.off[row] + (j - S.dp_beg[row])] : inf_min;
    };
    int best_score = inf_min, best_i = 0, best_j = 0;
    for (int in_id : g.node[SINK].in_id) {
        const int in_index = g.node_id_to_index[in_id];
        const int end = std::min(qlen, S.dp_end[in_index]);
        const int32_t v = val(S.H, in_index, end);
        if (v > best_score) { best_score = v; best_i = in_index; best_j = end; }
    }
    R.best_score = best_score;

    /* cg backtrack (A.7) */
    std::vector<Cig> rev;
    int i = best_i, j = best_j, cur_op = OP_ALL;
    if (best_j < qlen)
        for (int t = qlen - 1; t >= best_j; --t) rev.push_back({1, -1, t});
    while (i > 0 && j > 0) {
        const int id = g.index_to_node_id[i];
        const int *pre = &S.pre_idx[S.pre_off[i]];
        const int npre = S.pre_off[i + 1] - S.pre_off[i];
        const int s = P.mat[g.node[id].base * 5 + q[j - 1]];
        const int32_t hij = val(S.H, i, j);
        bool hit = false;
        if (cur_op & OP_M) {
            for (int k = 0; k < npre; ++k) {
                const int p = pre[k];
                if (j - 1 < S.dp_beg[p] || j - 1 > S.dp_end[p]) continue;
                if ((int64_t)val(S.H, p, j - 1) + s == hij) {
                    rev.push_back({0, id, j - 1});
                    i = p; --j; hit = true; cur_op = OP_ALL;
                    break;
                }
            }
        }
        if (!hit && (cur_op & OP_E)) {
            for (int k = 0; k < npre && !hit; ++k) {
                const int p = pre[k];
                if (j < S.dp_beg[p] || j > S.dp_end[p]) continue;
                if (cur_op & OP_E1) {
                    const int32_t pe1 = val(S.E1, p, j);
                    const bool take = (cur_op & OP_M) ? (hij == pe1)
                                                      : ((int64_t)val(S.E1, i, j) == (int64_t)pe1 - P.e1);
                    if (take) {
                        cur_op = ((int64_t)val(S.H, p, j) - P.oe1 == pe1) ? (OP_M | OP_F) : OP_E1;
                        rev.push_back({2, id, j - 1});
                        i = p; hit = true;
                        break;
                    }
                }
                if (cur_op & OP_E2) {
                    const int32_t pe2 = val(S.E2, p, j);
                    const bool take = (cur_op & OP_M) ? (hij == pe2)
                                                      : ((int64_t)val(S.E2, i, j) == (int64_t)pe2 - P.e2);
                    if (take) {
                        cur_op = ((int64_t)val(S.H, p, j) - P.oe2 == pe2) ? (OP_M | OP_F) : OP_E2;
                        rev.push_back({2, id, j - 1});
                        i = p; hit = true;
                        break;
                    }
                }
            }
        }
        if (!hit && (cur_op & OP_F)) {
            if (cur_op & OP_F1) {
                const int32_t f1 = val(S.F1, i, j);
                if (!(cur_op & OP_M) || hij == f1) {
                    if ((int64_t)val(S.H, i, j - 1) - P.oe1 == f1) { cur_op = OP_M | OP_E; hit = true; }
                    else if ((int64_t)val(S.F1, i, j - 1) - P.e1 == f1) { cur_op = OP_F1; hit = true; }
                }
            }
            if (!hit && (cur_op & OP_F2)) {
                const int32_t f2 = val(S.F2, i, j);
                if (!(cur_op & OP_M) || hij == f2) {
                    if ((int64_t)val(S.H, i, j - 1) - P.oe2 == f2) { cur_op = OP_M | OP_E; hit = true; }
                    else if ((int64_t)val(S.F2, i, j - 1) - P.e2 == f2) { cur_op = OP_F2; hit = true; }
                }
            }
            if (hit) { rev.push_back({1, id, j - 1}); --j; }
        }
        if (!hit) return;  // abPOA: err_fatal("Error in cg_backtrack") -> process dies, no output
    }
    for (int t = j - 1; t >= 0; --t) rev.push_back({1, -1, t});
    R.cigar.assign(rev.rbegin(), rev.rend());
    R.ok = true;
}

/* upstream abpoa_graph.c: abpoa_add_graph_sequence() -- first read, linear chain (A.3) */
void add_sequence(Graph &g, const uint8_t *seq, int len, int creator0, int32_t *base_aln, int32_t *base_node,
                  std::vector<int> *qnode = nullptr) {
    int last = SRC;
    if (qnode) qnode->assign(len, -1);
    for (int i = 0; i < len; ++i) {
        int cur = add_node(g, seq[i], creator0 + i);
        if (qnode) (*qnode)[i] = cur;
        add_edge(g, last, cur, 0, 1);
        last = cur;
        if (base_aln) base_aln[i] = -1;
        if (base_node) base_node[i] = creator0 + i;
    }
    add_edge(g, last, SINK, 0, 1);
    g.sorted = false;
}

/* upstream abpoa_graph.c: abpoa_add_subgraph_alignment() over the whole graph (A.8) */
void add_alignment(Graph &g, const uint8_t *seq, int len, const std::vector<Cig> &cigar, int creator0,
                   int32_t *base_aln, int32_t *base_node, std::vector<int> *qnode = nullptr) {
    int last_id = SRC, last_new = 0, query_id = -1;
    if (qnode) qnode->assign(len, -1);
    for (const Cig &c : cigar) {
        if (c.op == 0) {
            ++query_id;
            const int node_id = c.node_id;
            if (base_aln) base_aln[query_id] = g.node[node_id].creator;
            if (g.node[node_id].base != seq[query_id]) {
                int aligned_id = get_aligned_id(g, node_id, seq[query_id]);
                if (aligned_id != -1) {
                    add_edge(g, last_id, aligned_id, 1 - last_new, 1);
                    last_id = aligned_id; last_new = 0;
                } else {
                    int new_id = add_node(g, seq[query_id], creator0 + query_id);
                    add_edge(g, last_id, new_id, 0, 1);
                    last_id = new_id; last_new = 1;
                    add_aligned(g, node_id, new_id);
                }
            } else {
                add_edge(g, last_id, node_id, 1 - last_new, 1);
                last_id = node_id; last_new = 0;
            }
            if (base_node) base_node[query_id] = g.node[last_id].creator;
            if (qnode) (*qnode)[query_id] = last_id;
        } else if (c.op == 1) {
            ++query_id;
            int new_id = add_node(g, seq[query_id], creator0 + query_id);
            add_edge(g, last_id, new_id, 0, 1);
            last_id = new_id; last_new = 1;
            if (base_aln) base_aln[query_id] = -1;
            if (base_node) base_node[query_id] = creator0 + query_id;
            if (qnode) (*qnode)[query_id] = new_id;
        }
    }
    (void)len;
    add_edge(g, last_id, SINK, 1 - last_new, 1);
    g.sorted = false;
}

/* upstream abpoa_output.c: abpoa_heaviest_bundling() + abpoa_set_hb_cons(), n_clu = 1 (A.9) */
bool heaviest_bundling(const Graph &g, const Par &P, std::vector<uint8_t> &cons) {
    const int n = g.node_n();
    std::vector<int> out_degree(n), score(n, 0), max_out_id(n, -1);
    for (int i = 0; i < n; ++i) out_degree[i] = (int)g.node[i].out_id.size();
    std::deque<int> q;
    q.push_back(SINK);
    bool done = false;
    while (!q.empty()) {
        int cur = q.front();
        q.pop_front();
        const Node &nd = g.node[cur];
        if (cur == SINK) {
            max_out_id[cur] = -1; score[cur] = 0;
        } else if (cur == SRC) {
            int path_score = -1, path_max_w = -1, max_id = -1;
            for (size_t i = 0; i < nd.out_id.size(); ++i) {
                int out_id = nd.out_id[i], out_w = nd.out_w[i];
                if (out_w > path_max_w || (out_w == path_max_w && score[out_id] > path_score)) {
                    max_id = out_id; path_score = score[out_id]; path_max_w = out_w;
                }
            }
            max_out_id[cur] = max_id;
        } else {
            int max_w = INT_MIN, max_id = -1;
            for (size_t i = 0; i < nd.out_id.size(); ++i) {
                int out_id = nd.out_id[i], out_w = nd.out_w[i];
                if (max_w < out_w) { max_w = out_w; max_id = out_id; }
                else if (max_w == out_w) {
                    if (P.opt.hb_tie_later_wins ? (score[max_id] <= score[out_id]) : (score[max_id] < score[out_id]))
                        max_id = out_id;
                }
            }
            score[cur] = max_w + score[max_id];
            max_out_id[cur] = max_id;
        }
        if (cur == SRC) { done = true; break; }
        for (int in_id : nd.in_id)
            if (--out_degree[in_id] == 0) q.push_back(in_id);
    }
    if (!done) return false;
    cons.clear();
    int cur = max_out_id[SRC];
    while (cur != SINK && cur >= 0) {
        cons.push_back("ACGTN"[g.node[cur].base]);
        cur = max_out_id[cur];
    }
    return cur == SINK;
}

struct GroupOut {
    std::vector<uint8_t> cons;
    int status = MPOA_GROUP_EMPTY;
    mpoa_stats st;
};

/* ------------------------------------------------------------------------------------------ */
/* `abpoa -S`: minimizer-seeded, windowed alignment (upstream abpoa_seed.c + abpoa_anchor_poa in  */
/* abpoa.c).  LOW-CONFIDENCE restatement: the upstream source is unavailable, the structure below  */
/* follows SURVEY.md A.10 -- (k,w) minimizers of consecutive reads, colinear chaining of their     */
/* hits, anchors at least min_w apart, every anchor k-mer a forced run of matches, the stretches   */
/* between anchors aligned to the sub-graph between the anchor nodes -- and every constant that   */
/* the summary does not pin (chaining band, look-back, tie rules) is OUR choice, documented here:  */
/*   minimizers  forward strand only, minimap2 sampling rule, invertible 64-bit mix of the k-mer  */
/*   hits        equal hash in read i-1 (target) and read i (query); hashes with more than        */
/*               8 occurrences in the target are ignored; sorted by (target pos, query pos)        */
/*   chain       f[i] = k + max over the 64 previous hits j with 0 < dt, 0 < dq, |dt-dq| <= 100    */
/*               of f[j] - k + min(k, dt, dq); best chain = first maximum of f                    */
/*   anchors     walk the chain left to right, keep a hit when its k-mer starts >= min_w after the */
/*               end of the last kept one (or the read start) on BOTH reads                        */
/*   sub-graph   the nodes that lie on a path from the begin node to the end node (both ends act   */
/*               as source / sink of an ordinary alignment; remain = global remain - remain[end]-1)*/
/* The product library implements the same rules independently (csrc/seed.cpp + the kernels).       */
/* ------------------------------------------------------------------------------------------ */

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

struct Mz { uint64_t key; int pos; };   // pos = last base of the k-mer

/* (w,k) minimizers of the forward strand of nt4 codes (4 = N restarts the window) */
void minimizers_fw(const uint8_t *s, int len, int w, int k, std::vector<Mz> &out) {
    out.clear();
    if (w <= 0 || w > 64 || k <= 0 || k > 28) return;
    const uint64_t mask = (1ULL << 2 * k) - 1;
    uint64_t fw = 0;
    std::vector<uint64_t> bkey(w, UINT64_MAX);
    std::vector<int> bpos(w, 0);
    uint64_t mkey = UINT64_MAX;
    int mpos = 0, l = 0, bp = 0, mp = 0;
    auto emit = [&](uint64_t key, int pos) { if (key != UINT64_MAX) out.push_back(Mz{key, pos}); };
    auto ties = [&](int from, int to) { for (int j = from; j < to; ++j) if (bkey[j] == mkey && bpos[j] != mpos) emit(bkey[j], bpos[j]); };
    for (int i = 0; i < len; ++i) {
        const int c = s[i];
        uint64_t ckey = UINT64_MAX;
        if (c < 4) {
            fw = (fw << 2 | (uint64_t)c) & mask;
            ++l;
            if (l >= k) ckey = mix64(fw, mask);
        } else l = 0;
        bkey[bp] = ckey; bpos[bp] = i;
        if (l == w + k - 1 && mkey != UINT64_MAX) { ties(bp + 1, w); ties(0, bp); }
        if (ckey <= mkey) {
            if (l >= w + k) emit(mkey, mpos);
            mkey = ckey; mpos = i; mp = bp;
        } else if (bp == mp) {
            if (l >= w + k - 1) emit(mkey, mpos);
            uint64_t best = UINT64_MAX;
            int bj = bp;
            for (int t = 1; t <= w; ++t) {
                int j = bp + t; if (j >= w) j -= w;
                if (bkey[j] <= best) { best = bkey[j]; bj = j; }
            }
            mkey = best; mpos = bpos[bj]; mp = bj;
            if (l >= w + k - 1 && mkey != UINT64_MAX) { ties(bp + 1, w); ties(0, bp + 1); }
        }
        if (++bp == w) bp = 0;
    }
    emit(mkey, mpos);
}

/* anchors (start in the previous read, start in this read) of the seeded alignment of `cur` */
void seed_anchors(const uint8_t *prev, int plen, const uint8_t *cur, int clen, int k, int w, int min_w,
                  std::vector<std::pair<int, int>> &anchors) {
    anchors.clear();
    std::vector<Mz> mt, mq;
    minimizers_fw(prev, plen, w, k, mt);
    minimizers_fw(cur, clen, w, k, mq);
    std::sort(mt.begin(), mt.end(), [](const Mz &a, const Mz &b) { return a.key != b.key ? a.key < b.key : a.pos < b.pos; });
    struct Hit { int t, q; };
    std::vector<Hit> h;
    for (const Mz &m : mq) {
        auto lo = std::lower_bound(mt.begin(), mt.end(), m.key, [](const Mz &a, uint64_t key) { return a.key < key; });
        auto hi = lo;
        while (hi != mt.end() && hi->key == m.key) ++hi;
        if (hi - lo > 8) continue;
        for (auto it = lo; it != hi; ++it) h.push_back(Hit{it->pos, m.pos});
    }
    std::sort(h.begin(), h.end(), [](const Hit &a, const Hit &b) { return a.t != b.t ? a.t < b.t : a.q < b.q; });
    const int n = (int)h.size();
    if (n == 0) return;
    std::vector<int> f(n), p(n);
    int best = 0;
    for (int i = 0; i < n; ++i) {
        f[i] = k; p[i] = -1;
        for (int j = i - 1; j >= 0 && j >= i - 64; --j) {
            const int dt = h[i].t - h[j].t, dq = h[i].q - h[j].q;
            if (dt <= 0 || dq <= 0) continue;
            const int dd = dt > dq ? dt - dq : dq - dt;
            if (dd > 100) continue;
            const int sc = f[j] + std::min(k, std::min(dt, dq));
            if (sc > f[i]) { f[i] = sc; p[i] = j; }
        }
        if (f[i] > f[best]) best = i;
    }
    std::vector<int> chain;
    for (int i = best; i >= 0; i = p[i]) chain.push_back(i);
    std::reverse(chain.begin(), chain.end());
    int last_t = -1, last_q = -1;   // last base of the last kept anchor
    for (int i : chain) {
        const int t0 = h[i].t - k + 1, q0 = h[i].q - k + 1;
        if (t0 - (last_t + 1) >= min_w && q0 - (last_q + 1) >= min_w) {
            anchors.emplace_back(t0, q0);
            last_t = h[i].t; last_q = h[i].q;
        }
    }
}

/* alignment of query[0..ql) to the sub-graph of the nodes on a path beg -> end (both exclusive in the
 * cigar: beg plays the source, end the sink).  Appends the cigar (global node ids, query index + qoff). */
bool align_to_subgraph(Graph &g, const Par &P, int beg_id, int end_id, const uint8_t *q, int ql, int qoff, DpScratch &S,
                       AlnOut &R, std::vector<Cig> &cigar, int64_t &score_sum) {
    if (ql <= 0) return true;
    const int n = g.node_n();
    std::vector<char> fw(n, 0), bw(n, 0);
    std::vector<int> stack;
    stack.push_back(beg_id); fw[beg_id] = 1;
    while (!stack.empty()) { int v = stack.back(); stack.pop_back(); for (int u : g.node[v].out_id) if (!fw[u]) { fw[u] = 1; stack.push_back(u); } }
    stack.push_back(end_id); bw[end_id] = 1;
    while (!stack.empty()) { int v = stack.back(); stack.pop_back(); for (int u : g.node[v].in_id) if (!bw[u]) { bw[u] = 1; stack.push_back(u); } }
    if (!fw[end_id]) return false;
    std::vector<int> sub_of(n, -1), glob;
    glob.push_back(beg_id); glob.push_back(end_id);
    sub_of[beg_id] = SRC; sub_of[end_id] = SINK;
    for (int v = 0; v < n; ++v)
        if (fw[v] && bw[v] && v != beg_id && v != end_id) { sub_of[v] = (int)glob.size(); glob.push_back(v); }
    Graph sg;
    sg.node.resize(glob.size());
    for (size_t s2 = 0; s2 < glob.size(); ++s2) {
        const Node &src = g.node[glob[s2]];
        Node &d = sg.node[s2];
        d.base = src.base; d.creator = src.creator;
        if ((int)s2 != SINK)
            for (size_t e = 0; e < src.out_id.size(); ++e)
                if (sub_of[src.out_id[e]] >= 0) { d.out_id.push_back(sub_of[src.out_id[e]]); d.out_w.push_back(src.out_w[e]); }
        if ((int)s2 != SRC)
            for (int u : src.in_id) if (sub_of[u] >= 0) d.in_id.push_back(sub_of[u]);
        if ((int)s2 != SRC && (int)s2 != SINK)
            for (int a : src.aligned) if (sub_of[a] >= 0) d.aligned.push_back(sub_of[a]);
    }
    if (!topological_sort(sg)) return false;
    for (size_t s2 = 0; s2 < glob.size(); ++s2)
        sg.max_remain[s2] = g.max_remain[glob[s2]] - g.max_remain[end_id] - 1;
    AlnOut W;
    align_to_graph(sg, P, q, ql, S, W);
    R.band_cells += W.band_cells; R.int_ops += W.int_ops; R.full_cells += W.full_cells;
    R.max_width = std::max(R.max_width, W.max_width);
    R.bits = std::max(R.bits, W.bits);
    if (!W.ok) return false;
    score_sum += W.best_score;
    for (const Cig &c : W.cigar) cigar.push_back(Cig{c.op, c.node_id >= 0 ? glob[c.node_id] : -1, c.qidx + qoff});
    return true;
}

/* upstream abpoa.c: abpoa_anchor_poa() for one read */
void align_seeded(Graph &g, const Par &P, const uint8_t *seq, int len, const uint8_t *prev, int plen,
                  const std::vector<int> &prev_node, DpScratch &S, AlnOut &R) {
    const int k = P.opt.seed_k > 0 ? P.opt.seed_k : 19, w = P.opt.seed_w > 0 ? P.opt.seed_w : 10;
    const int min_w = P.opt.seed_min_w > 0 ? P.opt.seed_min_w : 500;
    std::vector<std::pair<int, int>> anchors;
    seed_anchors(prev, plen, seq, len, k, w, min_w, anchors);
    R.ok = false; R.cigar.clear(); R.bits = 0;
    int64_t score = 0;
    int beg_id = SRC, beg_q = 0;
    for (const auto &a : anchors) {
        const int t0 = a.first, q0 = a.second;
        if (!align_to_subgraph(g, P, beg_id, prev_node[t0], seq + beg_q, q0 - beg_q, beg_q, S, R, R.cigar, score)) return;
        for (int j = 0; j < k; ++j) R.cigar.push_back(Cig{0, prev_node[t0 + j], q0 + j});
        score += (int64_t)k * P.match;
        beg_id = prev_node[t0 + k - 1]; beg_q = q0 + k;
    }
    if (!align_to_subgraph(g, P, beg_id, SINK, seq + beg_q, len - beg_q, beg_q, S, R, R.cigar, score)) return;
    R.best_score = (int)score;
    if (R.bits == 0) R.bits = 16;
    R.ok = true;
}

/* upstream abpoa.c: abpoa_msa1() -> abpoa_poa() -> abpoa_output(), seeding disabled */
void run_group(const Par &P, int64_t r0, int64_t r1, const int64_t *read_base_off, const uint8_t *bases,
               mpoa_trace *tr, GroupOut &out, bool seeded) {
    std::memset(&out.st, 0, sizeof(out.st));
    out.st.n_groups = 1;
    out.st.n_reads = r1 - r0;
    out.status = MPOA_GROUP_EMPTY;
    out.cons.clear();
    if (r1 <= r0) return;
    Graph g;
    DpScratch S;
    AlnOut R;
    std::vector<uint8_t> seq, prev_seq;
    std::vector<int> prev_node, cur_node;
    out.st.n_seed_groups = seeded ? 1 : 0;
    out.st.n_seed_applied = seeded ? 1 : 0;
    const int64_t gbase = read_base_off[r0];
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t b0 = read_base_off[r], b1 = read_base_off[r + 1];
        const int len = (int)(b1 - b0);
        seq.resize(len);
        for (int i = 0; i < len; ++i) seq[i] = nt4(bases[b0 + i]);
        int32_t *baln = tr && tr->base_aln ? tr->base_aln + b0 : nullptr;
        int32_t *bnode = tr && tr->base_node ? tr->base_node + b0 : nullptr;
        if (tr && tr->read_score) tr->read_score[r] = 0;
        if (tr && tr->read_bits) tr->read_bits[r] = 0;
        if (tr && tr->read_band_cells) tr->read_band_cells[r] = 0;
        if (g.node_n() == 2) {
            /* abpoa_add_graph_sequence dies on an empty first read -> no output at all */
            if (len <= 0) return;
            add_sequence(g, seq.data(), len, (int)(b0 - gbase), baln, bnode, &prev_node);
            prev_seq = seq;
            continue;
        }
        if (len <= 0) continue;  // abpoa_align_sequence_to_graph returns early, nothing is added
        if (!g.sorted && !topological_sort(g)) return;
        R.band_cells = R.int_ops = R.full_cells = 0;
        R.max_width = 0;
        if (seeded) align_seeded(g, P, seq.data(), len, prev_seq.data(), (int)prev_seq.size(), prev_node, S, R);
        else align_to_graph(g, P, seq.data(), len, S, R);
        out.st.n_alignments++;
        out.st.band_cells += R.band_cells;
        out.st.int_ops += R.int_ops;
        out.st.full_cells += R.full_cells;
        (R.bits == 16 ? out.st.n_align_i16 : out.st.n_align_i32)++;
        if (R.max_width > out.st.max_band_width) out.st.max_band_width = R.max_width;  // widest band row (scheduling calibration)
        if (tr && tr->read_score) tr->read_score[r] = R.best_score;
        if (tr && tr->read_bits) tr->read_bits[r] = R.bits;
        if (tr && tr->read_band_cells) tr->read_band_cells[r] = R.band_cells;
        if (!R.ok) return;  // abpoa exits -> empty consensus file -> reference falls back
        add_alignment(g, seq.data(), len, R.cigar, (int)(b0 - gbase), baln, bnode, &cur_node);
        prev_node.swap(cur_node);
        prev_seq = seq;
    }
    if (g.node_n() <= 2) return;
    if (heaviest_bundling(g, P, out.cons)) out.status = MPOA_GROUP_OK;
    else out.cons.clear();
}

void add_stats(mpoa_stats &a, const mpoa_stats &b) {
    a.n_groups += b.n_groups; a.n_reads += b.n_reads; a.n_alignments += b.n_alignments;
    a.band_cells += b.band_cells; a.full_cells += b.full_cells; a.int_ops += b.int_ops;
    a.n_align_i16 += b.n_align_i16; a.n_align_i32 += b.n_align_i32;
    if (b.max_band_width > a.max_band_width) a.max_band_width = b.max_band_width;
    a.n_seed_groups += b.n_seed_groups; a.n_seed_applied += b.n_seed_applied;
}

}  // namespace

extern "C" void mpoa_oracle_default_opts(mpoa_oracle_opts *o) {
    std::memset(o, 0, sizeof(*o));
    o->clamp_end_to_pred = 1;
    o->single_argmax = 0;
    o->hb_tie_later_wins = 1;
    o->n_threads = 1;
    o->seed_k = 19; o->seed_w = 10; o->seed_min_w = 500;
    o->honour_seed_flag = 1;
}

extern "C" int mpoa_oracle_consensus_batch(const mpoa_params *p, const mpoa_oracle_opts *o, int64_t n_groups,
                                           const int64_t *group_read_off, const int64_t *read_base_off,
                                           const uint8_t *bases, const uint8_t *group_flags, int64_t *cons_off,
                                           uint8_t *cons_buf, int64_t cons_cap, int32_t *group_status,
                                           mpoa_stats *stats, mpoa_trace *trace) {
    if (!p || n_groups < 0 || (n_groups > 0 && (!group_read_off || !read_base_off || !cons_off))) return MPOA_EINVAL;
    Par P;
    P.match = p->match < 0 ? -p->match : p->match;
    P.mismatch = p->mismatch < 0 ? -p->mismatch : p->mismatch;
    P.o1 = p->gap_open1; P.e1 = p->gap_ext1; P.o2 = p->gap_open2; P.e2 = p->gap_ext2;
    P.oe1 = P.o1 + P.e1; P.oe2 = P.o2 + P.e2;
    P.wb = p->wb; P.wf = p->wf;
    P.pn16 = p->simd_pn_i16 > 0 ? p->simd_pn_i16 : 16;
    P.pn32 = p->simd_pn_i32 > 0 ? p->simd_pn_i32 : 8;
    if (o) P.opt = *o; else mpoa_oracle_default_opts(&P.opt);
    if (P.opt.seed_k <= 0) P.opt.seed_k = 19;
    if (P.opt.seed_w <= 0) P.opt.seed_w = 10;
    if (P.opt.seed_min_w <= 0) P.opt.seed_min_w = 500;
    make_matrix(P);

    std::vector<GroupOut> outs((size_t)n_groups);
    int nt = std::max(1, P.opt.n_threads);
    if (nt > n_groups) nt = (int)std::max<int64_t>(1, n_groups);
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        for (;;) {
            int64_t gidx = next.fetch_add(1);
            if (gidx >= n_groups) break;
            const bool seeded = group_flags && (group_flags[gidx] & MPOA_FLAG_SEED) && P.opt.honour_seed_flag;
            run_group(P, group_read_off[gidx], group_read_off[gidx + 1], read_base_off, bases, trace, outs[gidx], seeded);
        }
    };
    if (nt <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(worker);
        for (auto &t : th) t.join();
    }
    mpoa_stats tot;
    std::memset(&tot, 0, sizeof(tot));
    int64_t pos = 0;
    bool overflow = false;
    for (int64_t gidx = 0; gidx < n_groups; ++gidx) {
        cons_off[gidx] = pos;
        const auto &c = outs[gidx].cons;
        if (!overflow && pos + (int64_t)c.size() <= cons_cap && cons_buf) std::memcpy(cons_buf + pos, c.data(), c.size());
        else if (!c.empty()) overflow = true;
        pos += (int64_t)c.size();
        if (group_status) group_status[gidx] = outs[gidx].status;
        add_stats(tot, outs[gidx].st);
    }
    cons_off[n_groups] = pos;
    if (stats) *stats = tot;
    return overflow ? MPOA_ENOSPC : MPOA_OK;
}
