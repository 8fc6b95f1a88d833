// Test harness (built and run by tests/test_seed_host.py; the oracle is the CHECKER): the product's host-side
// `-S` seeding (mandalorion_b200/csrc/seed.cpp: minimizers, chaining, anchors written into the arena regions the
// device reads, chunk counters) against the oracle's seed_anchors() on random read chains -- error-bearing copies of
// a template, low-complexity stretches, N bases, a truncated read.  Exit code 0 = identical anchors everywhere.
#include "../../oracle/abpoa_oracle.cpp"
#include <atomic>
#include <random>
namespace mpoa {
void seed_batch(const int64_t *gro, const int64_t *rbo, const uint8_t *bases, const int64_t *src_off,
                const int32_t *order, int64_t n_order, int64_t chunk_size, std::atomic<int64_t> *done,
                int k, int w, int min_gap, int n_threads, const int32_t *anc_off, int32_t *anc);
}
int main() {
    std::mt19937_64 rng(7);
    const char *AC = "ACGT";
    int bad = 0, tot = 0; long nanch = 0;
    for (int rep = 0; rep < 40; ++rep) {
        int ng = 6;
        std::vector<int64_t> gro{0}, rbo{0};
        std::string bases;
        for (int g = 0; g < ng; ++g) {
            int L = 3000 + rng() % 9000, nr = 3 + rng() % 5;
            std::string t; for (int i = 0; i < L; ++i) t += AC[rng() % 4];
            if (rep % 5 == 0) for (int i = 0; i < L / 3; ++i) t[L/3 + i] = "AC"[i % 2];   // low complexity
            for (int r = 0; r < nr; ++r) {
                std::string s;
                for (int i = 0; i < L; ++i) {
                    double u = (rng() % 100000) / 100000.0;
                    if (u < 0.004) s += AC[rng() % 4];
                    else if (u < 0.008) { s += t[i]; s += AC[rng() % 4]; }
                    else if (u < 0.012) {}
                    else if (u < 0.0125) s += 'N';
                    else s += t[i];
                }
                if (rep % 7 == 3 && r == 2) s = s.substr(0, 600);
                bases += s; rbo.push_back((int64_t)bases.size());
            }
            gro.push_back((int64_t)rbo.size() - 1);
        }
        int64_t nreads = rbo.size() - 1;
        std::vector<int32_t> order; for (int g = ng - 1; g >= 0; --g) order.push_back(g);
        std::vector<int32_t> off(nreads + 1, 0); int64_t slot = 0;
        for (int g : order) for (int64_t r = gro[g]; r < gro[g+1]; ++r) { off[r] = slot; slot += (rbo[r+1]-rbo[r]) / 500 + 2; }
        std::vector<int32_t> arena(slot * 2, -7);
        std::vector<std::atomic<int64_t>> done(3); for (auto &d : done) d = 0;
        mpoa::seed_batch(gro.data(), rbo.data(), (const uint8_t *)bases.data(), nullptr, order.data(), ng, 2, done.data(), 19, 10, 500, 3, off.data(), arena.data());
        if (done[0] != 2 || done[1] != 2 || done[2] != 2) { printf("done counters wrong\n"); ++bad; }
        // oracle: nt4 codes
        std::vector<uint8_t> codes(bases.size());
        for (size_t i = 0; i < bases.size(); ++i) { char c = bases[i]; codes[i] = c=='A'?0:c=='C'?1:c=='G'?2:c=='T'?3:4; }
        for (int g = 0; g < ng; ++g)
            for (int64_t r = gro[g] + 1; r < gro[g+1]; ++r) {
                std::vector<std::pair<int,int>> a;
                seed_anchors(codes.data() + rbo[r-1], rbo[r]-rbo[r-1], codes.data() + rbo[r], rbo[r+1]-rbo[r], 19, 10, 500, a);
                const int32_t *reg = arena.data() + 2 * off[r];
                ++tot; nanch += a.size();
                bool ok = reg[0] == (int)a.size();
                for (size_t i = 0; ok && i < a.size(); ++i) ok = reg[2+2*i] == a[i].first && reg[3+2*i] == a[i].second;
                if (!ok) { ++bad; if (bad < 5) printf("rep %d g %d r %ld: %d vs %zu\n", rep, g, (long)r, reg[0], a.size()); }
            }
    }
    printf("%d read pairs, %ld anchors, %d mismatches\n", tot, nanch, bad);
    return bad != 0;
}
