// schedule.cpp -- host construction (and a host interpreter) of the work schedule consumed by the
// TMEM-accumulator apply kernel.  See schedule.h for the model.  Integer work only; the arithmetic
// the schedule describes is MultiOperator.apply, application/egsz/multi_operator2.py:38-84.
#include "schedule.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <queue>

namespace spuq {

void schedule_column_order(int32_t L, int32_t M, const int32_t* nbr_up, const int32_t* nbr_dn, int32_t block_cols,
                           std::vector<int32_t>* perm) {
    perm->resize(L);
    std::iota(perm->begin(), perm->end(), 0);
    if (M <= 0 || !nbr_up || !nbr_dn || L <= block_cols) return;

    // multi-indices by propagation from the roots (columns without a down-neighbour)
    std::vector<int32_t> mu((size_t)L * M, 0);
    std::vector<char> done(L, 0);
    std::queue<int32_t> q;
    for (int32_t c = 0; c < L; ++c) {
        bool root = true;
        for (int32_t m = 0; m < M && root; ++m) root = nbr_dn[(size_t)m * L + c] < 0;
        if (root) { done[c] = 1; q.push(c); }
    }
    while (!q.empty()) {
        const int32_t c = q.front(); q.pop();
        for (int32_t m = 0; m < M; ++m) {
            const int32_t u = nbr_up[(size_t)m * L + c];
            if (u < 0 || done[u]) continue;
            std::memcpy(&mu[(size_t)u * M], &mu[(size_t)c * M], sizeof(int32_t) * M);
            mu[(size_t)u * M + m] += 1;
            done[u] = 1;
            q.push(u);
        }
    }
    // dimensions by number of neighbour pairs, heaviest last in the key (= varies fastest)
    std::vector<int64_t> E(M, 0);
    for (int32_t m = 0; m < M; ++m)
        for (int32_t c = 0; c < L; ++c) E[m] += nbr_up[(size_t)m * L + c] >= 0;
    std::vector<int32_t> dims(M);
    std::iota(dims.begin(), dims.end(), 0);
    std::stable_sort(dims.begin(), dims.end(), [&](int32_t a, int32_t b) { return E[a] < E[b]; });   // lightest first
    std::stable_sort(perm->begin(), perm->end(), [&](int32_t a, int32_t b) {
        const int32_t* ma = &mu[(size_t)a * M];
        const int32_t* mb = &mu[(size_t)b * M];
        for (int32_t d : dims)
            if (ma[d] != mb[d]) return ma[d] < mb[d];
        return false;
    });
}

namespace {
struct Term { int32_t slot, in_col, out; double coef; };
struct Entry { int32_t out; double coef; };
struct Job {
    int32_t in_col = -1;
    uint32_t kmask = 0;
    int32_t S = 0;
    std::vector<Entry> es[SCHED_MAX_K];
};

struct Emitter {
    Schedule* s;
    std::vector<std::vector<int32_t>> chunk_cmds;   // producer commands by chunk of the current block
    int32_t chunk0 = 0;                              // first chunk of the current block
    int32_t n_win = 0;                               // window commands of the current block so far

    void begin_block() {
        chunk0 = (int32_t)(s->recs.size() / SCHED_CHUNK_RECS);
        chunk_cmds.clear();
        n_win = 0;
    }
    int32_t pos() const { return (int32_t)(s->recs.size() % SCHED_CHUNK_RECS); }
    // make room for n contiguous units in the current chunk (else link to a fresh chunk)
    void ensure(int32_t n) {
        if (pos() + n <= SCHED_CHUNK_USABLE) return;
        s->recs.push_back(SchedRec{REC_NEXT_CHUNK, 0, 0.0});
        while (pos() != 0) s->recs.push_back(SchedRec{REC_NEXT_CHUNK, 0, 0.0});
    }
    int32_t unit(int32_t a, int32_t b, double c) {
        const int32_t chunk = (int32_t)(s->recs.size() / SCHED_CHUNK_RECS) - chunk0;
        if ((int32_t)chunk_cmds.size() <= chunk) chunk_cmds.resize(chunk + 1);
        s->recs.push_back(SchedRec{a, b, c});
        return chunk;
    }
    int32_t unit_words(int32_t w0, int32_t w1, int32_t w2, int32_t w3) {
        const uint64_t bits = (uint64_t)(uint32_t)w2 | ((uint64_t)(uint32_t)w3 << 32);
        double c;
        std::memcpy(&c, &bits, 8);
        return unit(w0, w1, c);
    }
    void cmd(int32_t chunk, int32_t c) {
        chunk_cmds[chunk].push_back(c);
        if ((c & PCMD_KIND_MASK) == PCMD_WINDOW) ++n_win;
    }
    // A window group (one W-ring stage) must not straddle an A-set or block boundary: the consumer waits
    // for the WHOLE group, and its later windows would queue behind fetches (A tiles, metadata) that
    // can only be issued after the consumer has moved on.  Complete the group with pad commands.
    void pad_window_group(int32_t chunk) {
        while (n_win % s->w_group != 0) cmd(chunk, PCMD_WINDOW | PCMD_WINDOW_PAD);
    }
    void end_block() {
        ensure(1);
        pad_window_group(unit(REC_END_ITEM, 0, 0.0));
        while (pos() != 0) s->recs.push_back(SchedRec{REC_END_ITEM, 0, 0.0});
        const int32_t nch = (int32_t)chunk_cmds.size();
        // producer stream: metadata chunks run two ahead of the fetches they describe
        for (int32_t c = 0; c < std::min(nch, 2); ++c) s->pcmd.push_back(PCMD_META | (chunk0 + c));
        for (int32_t c = 0; c < nch; ++c) {
            for (int32_t x : chunk_cmds[c]) s->pcmd.push_back(x);
            if (c + 2 < nch) s->pcmd.push_back(PCMD_META | (chunk0 + c + 2));
        }
        s->blk_pcmd.push_back((int32_t)s->pcmd.size());
    }
};
}  // namespace

void schedule_build(const TermLists& terms, const std::vector<int32_t>& order, int32_t K, int32_t n_slots,
                    int32_t block_cols, int32_t w_group, Schedule* out) {
    Schedule& s = *out;
    s = Schedule{};
    if (K > SCHED_MAX_K) K = SCHED_MAX_K;
    s.K = K;
    s.block_cols = block_cols;
    s.w_group = w_group;
    const int32_t L = terms.L;
    const int32_t C = block_cols;
    const int32_t TRASH = block_cols;
    s.nblk = (L + C - 1) / C;
    s.blk_cols.assign((size_t)s.nblk * C, -1);
    s.blk_pcmd.push_back(0);
    Emitter em{&s};

    std::vector<Term> ts;
    std::vector<int64_t> slot_count(std::max(n_slots, 1));
    for (int32_t b = 0; b < s.nblk; ++b) {
        em.begin_block();
        const int32_t ncols = std::min(C, L - b * C);
        ts.clear();
        std::fill(slot_count.begin(), slot_count.end(), 0);
        for (int32_t o = 0; o < ncols; ++o) {
            const int32_t mu = order[(size_t)b * C + o];
            s.blk_cols[(size_t)b * C + o] = mu;
            for (int32_t t = terms.ptr[mu]; t < terms.ptr[mu + 1]; ++t) {
                ts.push_back(Term{terms.slot[t], terms.col[t], o, terms.coef[t]});
                slot_count[terms.slot[t]] += 1;
            }
        }
        {   // what any grouping of this block's matrices into sets of K could reach at best
            std::map<int32_t, std::vector<int32_t>> mats_of;
            for (const Term& t : ts) mats_of[t.in_col].push_back(t.slot);
            for (auto& kv : mats_of) {
                std::sort(kv.second.begin(), kv.second.end());
                const int64_t d = std::unique(kv.second.begin(), kv.second.end()) - kv.second.begin();
                s.n_inputs += 1;
                s.n_jobs_bound += (d + K - 1) / K;
            }
        }
        // matrices of this block, busiest first; K at a time form a register set
        std::vector<int32_t> slots;
        for (int32_t sl = 0; sl < (int32_t)slot_count.size(); ++sl)
            if (slot_count[sl] > 0) slots.push_back(sl);
        std::stable_sort(slots.begin(), slots.end(), [&](int32_t a, int32_t c2) { return slot_count[a] > slot_count[c2]; });
        // ... regrouped greedily so that the matrices of a set act on the same input columns: one
        // window load then serves up to K stencils (a job is an (input column, set) incidence)
        if (K > 1 && !std::getenv("SPUQ_NO_COGROUP")) {
            std::vector<std::vector<int32_t>> cols_of(slot_count.size());
            for (const Term& t : ts) cols_of[t.slot].push_back(t.in_col);
            for (auto& v : cols_of) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
            auto common = [&](int32_t a, int32_t c2) {
                const auto &x = cols_of[a], &y = cols_of[c2];
                size_t i = 0, j = 0, n = 0;
                while (i < x.size() && j < y.size()) {
                    if (x[i] < y[j]) ++i; else if (y[j] < x[i]) ++j; else { ++n; ++i; ++j; }
                }
                return (int64_t)n;
            };
            std::vector<int32_t> rest = slots, grouped;
            if (K == 2) {
                // maximum-weight matching on the co-occurrence graph, greedy by edge weight, then improved
                // by exchanging partners between two pairs while that raises the total
                const size_t n = slots.size();
                std::vector<std::vector<int64_t>> w(n, std::vector<int64_t>(n, 0));
                struct Edge { int64_t w; size_t a, b; };
                std::vector<Edge> edges;
                for (size_t a = 0; a < n; ++a)
                    for (size_t c2 = a + 1; c2 < n; ++c2) {
                        w[a][c2] = w[c2][a] = common(slots[a], slots[c2]);
                        edges.push_back(Edge{w[a][c2], a, c2});
                    }
                std::stable_sort(edges.begin(), edges.end(), [](const Edge& x, const Edge& y) { return x.w > y.w; });
                std::vector<int64_t> mate(n, -1);
                for (const Edge& e : edges)
                    if (mate[e.a] < 0 && mate[e.b] < 0) { mate[e.a] = (int64_t)e.b; mate[e.b] = (int64_t)e.a; }
                for (bool improved = true; improved;) {
                    improved = false;
                    for (size_t a = 0; a < n; ++a) {
                        if (mate[a] < 0) continue;
                        for (size_t c2 = a + 1; c2 < n; ++c2) {
                            if (mate[c2] < 0 || (size_t)mate[a] == c2) continue;
                            const size_t b = (size_t)mate[a], d = (size_t)mate[c2];
                            const int64_t cur = w[a][b] + w[c2][d];
                            if (w[a][c2] + w[b][d] > cur) { mate[a] = (int64_t)c2; mate[c2] = (int64_t)a; mate[b] = (int64_t)d; mate[d] = (int64_t)b; improved = true; }
                            else if (w[a][d] + w[b][c2] > cur) { mate[a] = (int64_t)d; mate[d] = (int64_t)a; mate[b] = (int64_t)c2; mate[c2] = (int64_t)b; improved = true; }
                        }
                    }
                }
                std::vector<char> done(n, 0);
                for (size_t a = 0; a < n; ++a) {
                    if (done[a] || mate[a] < 0) continue;
                    grouped.push_back(slots[a]); grouped.push_back(slots[(size_t)mate[a]]);
                    done[a] = done[(size_t)mate[a]] = 1;
                }
                for (size_t a = 0; a < n; ++a) if (!done[a]) grouped.push_back(slots[a]);
                rest.clear();
                {
                    // sets in ascending order of their smallest matrix: the blocks of a tile, which run
                    // concurrently, then ask for the A tiles of a matrix at about the same time
                    std::vector<std::pair<int32_t, int32_t>> prs;
                    for (size_t i = 0; i + 1 < grouped.size(); i += 2) prs.push_back({grouped[i], grouped[i + 1]});
                    std::vector<int32_t> tail;
                    if (grouped.size() % 2) tail.push_back(grouped.back());
                    std::stable_sort(prs.begin(), prs.end(), [](const std::pair<int32_t, int32_t>& x, const std::pair<int32_t, int32_t>& y) {
                        return std::min(x.first, x.second) < std::min(y.first, y.second);
                    });
                    grouped.clear();
                    for (auto& pr : prs) { grouped.push_back(pr.first); grouped.push_back(pr.second); }
                    for (int32_t t : tail) grouped.push_back(t);
                }
            }
            while (!rest.empty()) {
                std::vector<int32_t> grp{rest.front()};
                rest.erase(rest.begin());
                while ((int32_t)grp.size() < K && !rest.empty()) {
                    size_t best = 0;
                    int64_t best_w = -1;
                    for (size_t i = 0; i < rest.size(); ++i) {
                        int64_t w = 0;
                        for (int32_t g : grp) w += common(g, rest[i]);
                        if (w > best_w) { best_w = w; best = i; }      // ties: the busier one (rest keeps that order)
                    }
                    grp.push_back(rest[best]);
                    rest.erase(rest.begin() + best);
                }
                grouped.insert(grouped.end(), grp.begin(), grp.end());
            }
            slots.swap(grouped);
        }
        std::vector<int32_t> k_of_slot(slot_count.size(), -1), set_of_slot(slot_count.size(), -1);
        for (size_t i = 0; i < slots.size(); ++i) {
            set_of_slot[slots[i]] = (int32_t)(i / K);
            k_of_slot[slots[i]] = (int32_t)(i % K);
        }
        const int32_t nsets = (int32_t)((slots.size() + K - 1) / K);

        // ---- all runs of the block first (the RUN header needs to know whether another follows)
        struct Run { int32_t set; uint32_t kmask; int32_t S; std::vector<Job> jobs; };
        std::vector<Run> runs;
        for (int32_t set = 0; set < nsets; ++set) {
            std::map<int32_t, Job> by_col;
            for (const Term& t : ts) {
                if (set_of_slot[t.slot] != set) continue;
                Job& j = by_col[t.in_col];
                j.in_col = t.in_col;
                j.es[k_of_slot[t.slot]].push_back(Entry{t.out, t.coef});
            }
            // a (column, matrix) pair with more than SCHED_MAX_S outputs is served in several rounds
            std::vector<Job> jobs;
            for (auto& kv : by_col) {
                Job& full = kv.second;
                for (size_t round = 0;; ++round) {
                    Job j;
                    j.in_col = full.in_col;
                    for (int32_t k = 0; k < K; ++k) {
                        const size_t lo = round * SCHED_MAX_S;
                        if (full.es[k].size() <= lo) continue;
                        const size_t hi = std::min(full.es[k].size(), lo + SCHED_MAX_S);
                        j.es[k].assign(full.es[k].begin() + lo, full.es[k].begin() + hi);
                        j.kmask |= 1u << k;
                        j.S = std::max<int32_t>(j.S, (int32_t)(hi - lo));
                    }
                    if (!j.kmask) break;
                    jobs.push_back(std::move(j));
                }
            }
            // one run per slot count; inside it jobs with the same matrices are adjacent (pair partners)
            std::stable_sort(jobs.begin(), jobs.end(), [](const Job& x, const Job& y) {
                if (x.S != y.S) return x.S < y.S;
                return x.kmask > y.kmask;
            });
            for (size_t i = 0; i < jobs.size();) {
                size_t j = i;
                while (j < jobs.size() && jobs[j].S == jobs[i].S) ++j;
                Run r;
                r.set = set; r.kmask = 0; r.S = jobs[i].S;
                r.jobs.assign(std::make_move_iterator(jobs.begin() + i), std::make_move_iterator(jobs.begin() + j));
                runs.push_back(std::move(r));
                i = j;
            }
        }

        // ---- pair up jobs of a run that use the same matrices and have disjoint output columns (both
        // windows then go through the stencil in one basic block)
        struct Split { std::vector<std::pair<int32_t, int32_t>> pairs; std::vector<int32_t> singles; };
        std::vector<Split> split(runs.size());
        for (size_t ri = 0; ri < runs.size(); ++ri) {
            const Run& r = runs[ri];
            std::vector<uint64_t> outs(r.jobs.size(), 0);
            for (size_t i = 0; i < r.jobs.size(); ++i)
                for (int32_t k = 0; k < K; ++k)
                    for (const Entry& e : r.jobs[i].es[k]) outs[i] |= 1ull << e.out;
            std::vector<char> used(r.jobs.size(), 0);
            for (size_t i = 0; i < r.jobs.size(); ++i) {
                if (used[i]) continue;
                used[i] = 1;
                size_t partner = r.jobs.size();
                for (size_t j = i + 1; r.S <= 2 && !std::getenv("SPUQ_NO_PAIRS") && j < r.jobs.size() && j < i + 64; ++j)
                    if (!used[j] && r.jobs[j].kmask == r.jobs[i].kmask && !(outs[i] & outs[j])) { partner = j; break; }
                if (partner < r.jobs.size()) { used[partner] = 1; split[ri].pairs.push_back({(int32_t)i, (int32_t)partner}); }
                else split[ri].singles.push_back((int32_t)i);
            }
        }

        // ---- emit, set by set: A tiles, then all pair runs (so that a pair always starts on a window
        // group boundary when groups hold two windows), then the runs of single jobs
        auto unit0 = [&](const Job& j) {
            int32_t packs[SCHED_MAX_K] = {0, 0, 0};
            for (int32_t k = 0; k < K; ++k)
                for (int32_t sl = 0; sl < SCHED_MAX_S; ++sl) {
                    const int32_t o = (sl < (int32_t)j.es[k].size()) ? j.es[k][sl].out : TRASH;
                    packs[k] |= (o * SCHED_ACC_STRIDE) << (9 * sl);      // TMEM column offset of the accumulator (< 512)
                }
            const int32_t ch = em.unit_words(REC_JOB | ((int32_t)j.kmask << REC_JOB_KMASK_SHIFT), packs[0], packs[1], packs[2]);
            em.cmd(ch, PCMD_WINDOW | j.in_col);
        };
        auto coef_units = [&](const Job& j, int32_t cu) {
            for (int32_t k = 0; k < K; ++k) {
                if (j.kmask >> k & 1u) { s.n_stencils += 1; s.n_entries += (int64_t)j.es[k].size(); }
                for (int32_t u = 0; u < cu; ++u) {
                    double c2[2] = {0.0, 0.0};
                    for (int q = 0; q < 2; ++q)
                        if (2 * u + q < (int32_t)j.es[k].size()) c2[q] = j.es[k][2 * u + q].coef;
                    SchedRec rec;
                    std::memcpy(&rec.a, &c2[0], 8);       // a,b = first coefficient, c = second
                    rec.c = c2[1];
                    s.recs.push_back(rec);
                }
            }
        };
        for (size_t r0 = 0; r0 < runs.size();) {
            size_t r1 = r0;
            while (r1 < runs.size() && runs[r1].set == runs[r0].set) ++r1;
            const int32_t cur_set = runs[r0].set;
            {
                // one SET header per register set: matrices to load and the lengths of its five runs
                // (pairs with S = 1, 2, then single jobs with S = 1, 2, 3), in the order they are emitted
                int32_t np[2] = {0, 0}, ns[3] = {0, 0, 0};
                for (size_t ri = r0; ri < r1; ++ri) {
                    const int32_t S1 = runs[ri].S - 1;
                    if (S1 < 2) np[S1] += (int32_t)split[ri].pairs.size();
                    ns[S1] += (int32_t)split[ri].singles.size();
                }
                const int32_t cnt = std::min<int32_t>(K, (int32_t)slots.size() - cur_set * K);
                em.ensure(2);
                const int32_t ch = em.unit_words(REC_ASET | (cnt << REC_CNT_SHIFT), ns[2], 0, 0);
                em.unit_words(np[0], np[1], ns[0], ns[1]);
                em.pad_window_group(ch);
                for (int32_t k = 0; k < cnt; ++k) em.cmd(ch, PCMD_ATILE | slots[(size_t)cur_set * K + k]);
                s.n_asets += cnt;
            }
            for (int pass = 0; pass < 2; ++pass)
                for (size_t ri = r0; ri < r1; ++ri) {
                    const Run& r = runs[ri];
                    const int32_t cu = (r.S + 1) / 2;               // coefficient units per register set
                    const int32_t job_units = 1 + K * cu;
                    if (pass == 0) {
                        s.n_jobs += (int64_t)r.jobs.size();
                        if (split[ri].pairs.empty()) continue;
                        for (const auto& pr : split[ri].pairs) {
                            em.ensure(2 * job_units);
                            unit0(r.jobs[pr.first]);
                            unit0(r.jobs[pr.second]);
                            coef_units(r.jobs[pr.first], cu);
                            coef_units(r.jobs[pr.second], cu);
                        }
                        s.n_pairs += (int64_t)split[ri].pairs.size();
                        for (const auto& pr : split[ri].pairs) s.hist_pairs[std::min<int>(r.S, 3) - 1][r.jobs[pr.first].kmask & 7] += 1;
                    } else {
                        if (split[ri].singles.empty()) continue;
                        for (int32_t i : split[ri].singles) {
                            s.hist_singles[std::min<int>(r.S, 3) - 1][r.jobs[i].kmask & 7] += 1;
                            em.ensure(job_units);
                            unit0(r.jobs[i]);
                            coef_units(r.jobs[i], cu);
                        }
                    }
                }
            r0 = r1;
        }
        {
            const int32_t nq = (ncols + 3) / 4;
            em.ensure(1 + nq);
            em.unit(REC_STOREALL, nq, 0.0);
            for (int32_t q = 0; q < nq; ++q) {
                int32_t c4[4];
                for (int32_t t = 0; t < 4; ++t) c4[t] = (4 * q + t < ncols) ? s.blk_cols[(size_t)b * C + 4 * q + t] : -1;
                em.unit_words(c4[0], c4[1], c4[2], c4[3]);
            }
        }
        em.end_block();
    }
}

void schedule_run_host(const Schedule& s, int32_t D, const int32_t* dx, const int32_t* dy, int32_t nx, int32_t ny,
                       const double* dia, const double* W, int64_t ldw, double* Y, int64_t ldy) {
    const int64_t N = (int64_t)nx * ny;
    const int32_t K = s.K;
    std::vector<double> acc((size_t)(s.block_cols + 1) * N), S(N);
    std::vector<int32_t> set_slot(std::max(K, 1), -1);
    auto fail = [&]() { Y[0] = 0.0 / 0.0; };
    auto words = [](const SchedRec& r, int32_t (&w)[4]) {
        uint64_t bits;
        std::memcpy(&bits, &r.c, 8);
        w[0] = r.a; w[1] = r.b; w[2] = (int32_t)(uint32_t)(bits & 0xffffffffu); w[3] = (int32_t)(uint32_t)(bits >> 32);
    };
    for (int32_t b = 0; b < s.nblk; ++b) {
        // what the producer would fetch, in order, and the chunks of the block
        int32_t chunk0 = -1;
        std::vector<int32_t> fetches, metas;
        int32_t n_win = 0;
        for (int32_t p = s.blk_pcmd[b]; p < s.blk_pcmd[b + 1]; ++p) {
            const int32_t cmd = s.pcmd[p];
            if ((cmd & PCMD_KIND_MASK) == PCMD_WINDOW) ++n_win;
            if (cmd == (PCMD_WINDOW | PCMD_WINDOW_PAD)) continue;
            // a group may only be padded, never split, at an A-set boundary
            if ((cmd & PCMD_KIND_MASK) == PCMD_ATILE && !fetches.empty() && (fetches.back() & PCMD_KIND_MASK) == PCMD_WINDOW
                && n_win % s.w_group != 0) { fail(); return; }
            if ((cmd & PCMD_KIND_MASK) == PCMD_META) { if (chunk0 < 0) chunk0 = cmd & PCMD_VAL_MASK; metas.push_back(cmd & PCMD_VAL_MASK); }
            else fetches.push_back(cmd);
        }
        if (chunk0 < 0 || n_win % s.w_group != 0) { fail(); return; }
        for (size_t i = 0; i < metas.size(); ++i) if (metas[i] != chunk0 + (int32_t)i) { fail(); return; }
        std::fill(acc.begin(), acc.end(), 0.0);
        size_t r = (size_t)chunk0 * SCHED_CHUNK_RECS, f = 0, chunks_seen = 1;
        int32_t win_seen = 0;
        auto next_chunk = [&]() { r = (r / SCHED_CHUNK_RECS + 1) * SCHED_CHUNK_RECS; ++chunks_seen; };
        bool running = true;
        while (running) {
            if ((s.recs[r].a & REC_TYPE_MASK) == REC_NEXT_CHUNK) { next_chunk(); continue; }
            const SchedRec& rec = s.recs[r++];
            switch (rec.a & REC_TYPE_MASK) {
            case REC_END_ITEM: running = false; break;
            case REC_ASET: {
                win_seen = (win_seen + s.w_group - 1) / s.w_group * s.w_group;      // the group is padded here
                const int32_t cnt = (rec.a >> REC_CNT_SHIFT) & 0xF;
                for (int32_t k = 0; k < cnt; ++k) {
                    if (f >= fetches.size() || (fetches[f] & PCMD_KIND_MASK) != PCMD_ATILE) { fail(); return; }
                    set_slot[k] = fetches[f++] & PCMD_VAL_MASK;
                }
                // run lengths: pairs S=1,2 then singles S=1,2,3 (the second header unit holds the first four)
                int32_t cw[4];
                words(s.recs[r++], cw);
                const int32_t run_n[5] = {cw[0], cw[1], cw[2], cw[3], rec.b};
                for (int32_t ri = 0; ri < 5; ++ri) {
                const bool two = ri < 2;
                uint32_t kmask = 0;
                const int32_t Sn = two ? ri + 1 : ri - 1, n = run_n[ri];
                const int32_t cu = (Sn + 1) / 2, job_units = 1 + K * cu;
                const int32_t group_units = two ? 2 * job_units : job_units, per = two ? 2 : 1;
                for (int32_t j = 0; j < n; ++j) {
                    if ((s.recs[r].a & REC_TYPE_MASK) == REC_NEXT_CHUNK) next_chunk();
                    if (r % SCHED_CHUNK_RECS + group_units > SCHED_CHUNK_USABLE) { fail(); return; }
                    if (two && s.w_group % 2 == 0 && win_seen % 2 != 0) { fail(); return; }   // pairs are group aligned
                    uint64_t seen_out = 0;
                    for (int32_t q = 0; q < per; ++q) {
                        int32_t w[4];
                        words(s.recs[r + q], w);
                        if ((w[0] & REC_TYPE_MASK) != REC_JOB) { fail(); return; }
                        const uint32_t km = (uint32_t)(w[0] >> REC_JOB_KMASK_SHIFT) & 0xFu;
                        if (q > 0 && km != kmask) { fail(); return; }       // pair partners use the same matrices
                        kmask = km;
                        if (f >= fetches.size() || (fetches[f] & PCMD_KIND_MASK) != PCMD_WINDOW) { fail(); return; }
                        const int32_t in_col = fetches[f++] & PCMD_VAL_MASK;
                        ++win_seen;
                        const double* wv = W + (int64_t)in_col * ldw;
                        const size_t cbase = r + per + (size_t)q * K * cu;
                        uint64_t my_out = 0;
                        for (int32_t k = 0; k < K; ++k) {
                            if (!(kmask >> k & 1u)) continue;
                            const double* a = dia + (size_t)set_slot[k] * D * N;
                            for (int32_t iy = 0; iy < ny; ++iy)
                                for (int32_t ix = 0; ix < nx; ++ix) {
                                    const int64_t i = (int64_t)iy * nx + ix;
                                    double v = 0.0;
                                    for (int32_t d = 0; d < D; ++d) {
                                        const int32_t jx = ix + dx[d], jy = iy + dy[d];
                                        const double x = (jx >= 0 && jx < nx && jy >= 0 && jy < ny) ? wv[(int64_t)jy * nx + jx] : 0.0;
                                        v += a[(size_t)d * N + i] * x;
                                    }
                                    S[i] = v;
                                }
                            for (int32_t sl = 0; sl < Sn; ++sl) {
                                const int32_t o = ((w[1 + k] >> (9 * sl)) & 511) / SCHED_ACC_STRIDE;
                                if (o != s.block_cols) my_out |= 1ull << o;
                                const SchedRec& cr = s.recs[cbase + (size_t)k * cu + sl / 2];
                                double coef;
                                if (sl % 2 == 0) std::memcpy(&coef, &cr.a, 8); else coef = cr.c;
                                double* ac = &acc[(size_t)o * N];
                                for (int64_t i = 0; i < N; ++i) ac[i] += coef * S[i];
                            }
                        }
                        // the two jobs of a pair run concurrently on the device: their outputs must differ
                        if (seen_out & my_out) { fail(); return; }
                        seen_out |= my_out;
                    }
                    r += group_units;
                }
                }
                break;
            }
            case REC_STOREALL: {
                for (int32_t q = 0; q < rec.b; ++q) {
                    int32_t w[4];
                    words(s.recs[r++], w);
                    for (int t = 0; t < 4; ++t)
                        if (w[t] >= 0)
                            std::memcpy(Y + (int64_t)w[t] * ldy, &acc[(size_t)(4 * q + t) * N], sizeof(double) * N);
                }
                break;
            }
            default: fail(); return;
            }
        }
        if (f != fetches.size() || chunks_seen != metas.size()) { fail(); return; }
    }
}

}  // namespace spuq
