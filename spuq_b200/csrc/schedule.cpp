// schedule.cpp -- host construction (and a host interpreter) of the work schedule consumed by the
// TMEM-accumulator apply kernel.  See schedule.h for the model.  Integer work only; the arithmetic
// the schedule describes is MultiOperator.apply, application/egsz/multi_operator2.py:38-84.
#include "schedule.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <numeric>
#include <queue>

namespace spuq {

void schedule_column_order(int32_t L, int32_t M, const int32_t* nbr_up, const int32_t* nbr_dn,
                           std::vector<int32_t>* perm) {
    perm->resize(L);
    std::iota(perm->begin(), perm->end(), 0);
    if (M <= 0 || !nbr_up || !nbr_dn || L <= SCHED_BLOCK_COLS) return;

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
struct Entry { int32_t k, out; double coef; };
struct Job { int32_t in_col; uint32_t kmask; std::vector<Entry> es; };

struct Emitter {
    Schedule* s;
    std::vector<std::vector<int32_t>> chunk_cmds;   // producer commands by chunk of the current block
    int32_t chunk0 = 0;                              // first chunk of the current block

    void begin_block() {
        chunk0 = (int32_t)(s->recs.size() / SCHED_CHUNK_RECS);
        chunk_cmds.clear();
    }
    // append a record (the last slot of every chunk is the link to the next chunk); returns the chunk
    int32_t rec(int32_t a, int32_t b, double c) {
        if (s->recs.size() % SCHED_CHUNK_RECS == SCHED_CHUNK_RECS - 1) s->recs.push_back(SchedRec{REC_NEXT_CHUNK, 0, 0.0});
        const int32_t chunk = (int32_t)(s->recs.size() / SCHED_CHUNK_RECS) - chunk0;
        if ((int32_t)chunk_cmds.size() <= chunk) chunk_cmds.resize(chunk + 1);
        s->recs.push_back(SchedRec{a, b, c});
        return chunk;
    }
    void cmd(int32_t chunk, int32_t c) { chunk_cmds[chunk].push_back(c); }
    void end_block() {
        rec(REC_END_ITEM, 0, 0.0);
        while (s->recs.size() % SCHED_CHUNK_RECS) s->recs.push_back(SchedRec{REC_END_ITEM, 0, 0.0});
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
                    Schedule* out) {
    Schedule& s = *out;
    s = Schedule{};
    s.K = K;
    const int32_t L = terms.L;
    const int32_t C = SCHED_BLOCK_COLS;
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
        // a block without any term still has to write zeros: one null term on its first column
        if (ts.empty()) { ts.push_back(Term{0, order[(size_t)b * C], 0, 0.0}); slot_count[0] = 1; }
        // matrices of this block, busiest first; K at a time form a register set
        std::vector<int32_t> slots;
        for (int32_t sl = 0; sl < (int32_t)slot_count.size(); ++sl)
            if (slot_count[sl] > 0) slots.push_back(sl);
        std::stable_sort(slots.begin(), slots.end(), [&](int32_t a, int32_t c2) { return slot_count[a] > slot_count[c2]; });
        std::vector<int32_t> k_of_slot(slot_count.size(), -1), set_of_slot(slot_count.size(), -1);
        for (size_t i = 0; i < slots.size(); ++i) {
            set_of_slot[slots[i]] = (int32_t)(i / K);
            k_of_slot[slots[i]] = (int32_t)(i % K);
        }
        const int32_t nsets = (int32_t)((slots.size() + K - 1) / K);
        std::vector<char> touched(ncols, 0);
        int32_t n_touched = 0;

        for (int32_t set = 0; set < nsets; ++set) {
            // ---- A tiles of the set
            const int32_t cnt = std::min<int32_t>(K, (int32_t)slots.size() - set * K);
            {
                const int32_t ch = em.rec(REC_ASET | (cnt << REC_CNT_SHIFT), 0, 0.0);
                for (int32_t k = 0; k < cnt; ++k) em.cmd(ch, PCMD_ATILE | slots[(size_t)set * K + k]);
                s.n_asets += cnt;
            }
            // ---- jobs of the set: one per input column, entries ordered by register set k
            std::map<int32_t, Job> by_col;
            for (const Term& t : ts) {
                if (set_of_slot[t.slot] != set) continue;
                Job& j = by_col[t.in_col];
                j.in_col = t.in_col;
                j.kmask |= 1u << k_of_slot[t.slot];
                j.es.push_back(Entry{k_of_slot[t.slot], t.out, t.coef});
            }
            std::vector<Job> jobs;
            for (auto& kv : by_col) {
                std::stable_sort(kv.second.es.begin(), kv.second.es.end(), [](const Entry& x, const Entry& y) {
                    if (x.k != y.k) return x.k < y.k;
                    return x.out < y.out;
                });
                jobs.push_back(std::move(kv.second));
            }
            // jobs that use the same matrices next to each other, so that pairs share their stencils
            std::stable_sort(jobs.begin(), jobs.end(), [](const Job& x, const Job& y) { return x.kmask > y.kmask; });
            s.n_jobs += (int64_t)jobs.size();

            auto emit_reload = [&](size_t first_job) {
                const bool two = first_job + 1 < jobs.size();
                const int32_t ch = em.rec(REC_RELOAD | REC_F_BUFA | (two ? REC_F_BUFB : 0), 0, 0.0);
                em.cmd(ch, PCMD_WINDOW | jobs[first_job].in_col);
                if (two) em.cmd(ch, PCMD_WINDOW | jobs[first_job + 1].in_col);
            };
            if (!jobs.empty()) emit_reload(0);
            for (size_t j0 = 0; j0 < jobs.size(); j0 += 2) {
                const Job& ja = jobs[j0];
                const Job* jb = (j0 + 1 < jobs.size()) ? &jobs[j0 + 1] : nullptr;
                const uint32_t kunion = ja.kmask | (jb ? jb->kmask : 0u);
                int32_t last_k = -1;
                for (int32_t k = 0; k < K; ++k) if (kunion >> k & 1u) last_k = k;
                for (int32_t k = 0; k < K; ++k) {
                    if (!(kunion >> k & 1u)) continue;
                    // accumulations of this matrix: (source buffer, entry)
                    std::vector<std::pair<int, const Entry*>> acc;
                    for (const Entry& e : ja.es) if (e.k == k) acc.push_back({0, &e});
                    if (jb) for (const Entry& e : jb->es) if (e.k == k) acc.push_back({1, &e});
                    int32_t st = 0;
                    if (ja.kmask >> k & 1u) st |= REC_F_STA;
                    if (jb && (jb->kmask >> k & 1u)) st |= REC_F_STB;
                    s.n_stencils += (st & REC_F_STA ? 1 : 0) + (st & REC_F_STB ? 1 : 0);
                    for (size_t q = 0; q < acc.size(); ++q) {
                        const Entry& e = *acc[q].second;
                        int32_t a = REC_OP | (k << REC_K_SHIFT) | (q == 0 ? st : 0);
                        if (acc[q].first) a |= REC_F_SRCB;
                        if (!touched[e.out]) { a |= REC_F_INIT; touched[e.out] = 1; ++n_touched; }
                        em.rec(a, e.out * SCHED_ACC_STRIDE, e.coef);
                        s.n_entries += 1;
                        // the windows are dead once the last stencil of the pair has been issued:
                        // fetch the next pair's while the remaining accumulations run
                        if (q == 0 && k == last_k && j0 + 2 < jobs.size()) emit_reload(j0 + 2);
                    }
                }
            }
        }
        // output columns without any term (possible for an M-shard): zero them (0 * last S_A)
        for (int32_t o = 0; o < ncols; ++o)
            if (!touched[o]) {
                em.rec(REC_OP | REC_F_INIT, o * SCHED_ACC_STRIDE, 0.0);
                touched[o] = 1; ++n_touched; s.n_entries += 1;
            }
        for (int32_t o = 0; o < ncols; o += 2) {
            const int32_t c0 = s.blk_cols[(size_t)b * C + o];
            const int32_t c1 = (o + 1 < ncols) ? s.blk_cols[(size_t)b * C + o + 1] : -1;
            const uint64_t bits = (uint64_t)(uint32_t)c0 | ((uint64_t)(uint32_t)c1 << 32);
            double c;
            std::memcpy(&c, &bits, 8);
            em.rec(REC_STORE, o * SCHED_ACC_STRIDE, c);
        }
        em.end_block();
    }
}

void schedule_run_host(const Schedule& s, int32_t D, const int32_t* dx, const int32_t* dy, int32_t nx, int32_t ny,
                       const double* dia, const double* W, int64_t ldw, double* Y, int64_t ldy) {
    const int64_t N = (int64_t)nx * ny;
    const int32_t C = SCHED_BLOCK_COLS;
    std::vector<double> acc((size_t)C * N), S[2];
    S[0].assign(N, 0.0); S[1].assign(N, 0.0);
    std::vector<int32_t> set_slot(std::max(s.K, 1), -1);
    auto fail = [&]() { Y[0] = 0.0 / 0.0; };
    for (int32_t b = 0; b < s.nblk; ++b) {
        // what the producer would fetch, in order, and the first chunk of the block
        int32_t chunk0 = -1;
        std::vector<int32_t> fetches, metas;
        for (int32_t p = s.blk_pcmd[b]; p < s.blk_pcmd[b + 1]; ++p) {
            const int32_t cmd = s.pcmd[p];
            if ((cmd & PCMD_KIND_MASK) == PCMD_META) { if (chunk0 < 0) chunk0 = cmd & PCMD_VAL_MASK; metas.push_back(cmd & PCMD_VAL_MASK); }
            else fetches.push_back(cmd);
        }
        if (chunk0 < 0) { fail(); return; }
        size_t r = (size_t)chunk0 * SCHED_CHUNK_RECS, f = 0, chunks_seen = 1;
        int32_t win_col[2] = {-1, -1};
        bool running = true;
        while (running) {
            const SchedRec& rec = s.recs[r++];
            const int32_t type = rec.a & REC_TYPE_MASK;
            switch (type) {
            case REC_END_ITEM: running = false; break;
            case REC_NEXT_CHUNK:
                if (r % SCHED_CHUNK_RECS != 0) { fail(); return; }       // must be the last slot of its chunk
                ++chunks_seen;
                break;
            case REC_ASET: {
                const int32_t cnt = (rec.a >> REC_CNT_SHIFT) & 0xF;
                for (int32_t k = 0; k < cnt; ++k) {
                    if (f >= fetches.size() || (fetches[f] & PCMD_KIND_MASK) != PCMD_ATILE) { fail(); return; }
                    set_slot[k] = fetches[f++] & PCMD_VAL_MASK;
                }
                break;
            }
            case REC_RELOAD:
                for (int q = 0; q < 2; ++q) {
                    if (!(rec.a & (q ? REC_F_BUFB : REC_F_BUFA))) continue;
                    if (f >= fetches.size() || (fetches[f] & PCMD_KIND_MASK) != PCMD_WINDOW) { fail(); return; }
                    win_col[q] = fetches[f++] & PCMD_VAL_MASK;
                }
                break;
            case REC_OP: {
                const int32_t k = (rec.a >> REC_K_SHIFT) & 3;
                for (int q = 0; q < 2; ++q) {
                    if (!(rec.a & (q ? REC_F_STB : REC_F_STA))) continue;
                    const double* a = dia + (size_t)set_slot[k] * D * N;
                    const double* w = W + (int64_t)win_col[q] * ldw;
                    for (int32_t iy = 0; iy < ny; ++iy)
                        for (int32_t ix = 0; ix < nx; ++ix) {
                            const int64_t i = (int64_t)iy * nx + ix;
                            double v = 0.0;
                            for (int32_t d = 0; d < D; ++d) {
                                const int32_t jx = ix + dx[d], jy = iy + dy[d];
                                const double wv = (jx >= 0 && jx < nx && jy >= 0 && jy < ny) ? w[(int64_t)jy * nx + jx] : 0.0;
                                v += a[(size_t)d * N + i] * wv;
                            }
                            S[q][i] = v;
                        }
                }
                if (rec.a & REC_F_NOACC) break;
                const std::vector<double>& src = S[(rec.a & REC_F_SRCB) ? 1 : 0];
                double* ac = &acc[(size_t)(rec.b / SCHED_ACC_STRIDE) * N];
                if (rec.a & REC_F_INIT) for (int64_t i = 0; i < N; ++i) ac[i] = rec.c * src[i];
                else for (int64_t i = 0; i < N; ++i) ac[i] += rec.c * src[i];
                break;
            }
            case REC_STORE: {
                uint64_t bits;
                std::memcpy(&bits, &rec.c, 8);
                const int32_t cols[2] = {(int32_t)(uint32_t)(bits & 0xffffffffu), (int32_t)(uint32_t)(bits >> 32)};
                for (int q = 0; q < 2; ++q)
                    if (cols[q] >= 0)
                        std::memcpy(Y + (int64_t)cols[q] * ldy, &acc[(size_t)(rec.b / SCHED_ACC_STRIDE + q) * N], sizeof(double) * N);
                break;
            }
            default: fail(); return;
            }
        }
        if (f != fetches.size() || chunks_seen != metas.size()) { fail(); return; }
    }
}

}  // namespace spuq
