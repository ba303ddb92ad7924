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

struct Emitter {
    Schedule* s;
    std::vector<std::vector<int32_t>> chunk_cmds;   // producer commands by chunk of the current block
    int32_t chunk0 = 0;                              // first chunk of the current block

    void begin_block() {
        chunk0 = (int32_t)(s->recs.size() / SCHED_CHUNK_RECS);
        chunk_cmds.clear();
    }
    void rec(int32_t a, int32_t b, double c, int32_t cmd = -1) {
        const int32_t chunk = (int32_t)(s->recs.size() / SCHED_CHUNK_RECS) - chunk0;
        if ((int32_t)chunk_cmds.size() <= chunk) chunk_cmds.resize(chunk + 1);
        if (cmd != -1) chunk_cmds[chunk].push_back(cmd);
        s->recs.push_back(SchedRec{a, b, c});
    }
    void end_block() {
        rec(REC_END_ITEM, 0, 0.0);
        while (s->recs.size() % SCHED_CHUNK_RECS) s->recs.push_back(SchedRec{REC_END_ITEM, 0, 0.0});
        const int32_t nch = (int32_t)chunk_cmds.size();
        // producer stream: metadata chunks run two ahead of the fetches they describe
        for (int32_t c = 0; c < std::min(nch, 2); ++c) s->pcmd.push_back(PCMD_META | (chunk0 + c));
        for (int32_t c = 0; c < nch; ++c) {
            for (int32_t cmd : chunk_cmds[c]) s->pcmd.push_back(cmd);
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
        if (ts.empty()) ts.push_back(Term{0, order[(size_t)b * C], 0, 0.0});
        // matrices of this block, busiest first; K at a time form a register set
        std::vector<int32_t> slots;
        for (int32_t sl = 0; sl < (int32_t)slot_count.size(); ++sl)
            if (slot_count[sl] > 0) slots.push_back(sl);
        if (slots.empty()) slots.push_back(ts[0].slot);
        std::stable_sort(slots.begin(), slots.end(), [&](int32_t a, int32_t c2) { return slot_count[a] > slot_count[c2]; });
        std::vector<int32_t> k_of_slot(slot_count.size(), -1), set_of_slot(slot_count.size(), -1);
        for (size_t i = 0; i < slots.size(); ++i) {
            set_of_slot[slots[i]] = (int32_t)(i / K);
            k_of_slot[slots[i]] = (int32_t)(i % K);
        }
        const int32_t nsets = (int32_t)((slots.size() + K - 1) / K);
        std::stable_sort(ts.begin(), ts.end(), [&](const Term& x, const Term& y) {
            if (set_of_slot[x.slot] != set_of_slot[y.slot]) return set_of_slot[x.slot] < set_of_slot[y.slot];
            if (x.in_col != y.in_col) return x.in_col < y.in_col;
            if (k_of_slot[x.slot] != k_of_slot[y.slot]) return k_of_slot[x.slot] < k_of_slot[y.slot];
            return x.out < y.out;
        });
        std::vector<char> touched(ncols, 0);
        int32_t n_touched = 0;
        size_t i = 0;
        for (int32_t set = 0; set < nsets; ++set) {
            for (int32_t k = 0; k < K && (size_t)(set * K + k) < slots.size(); ++k) {
                const int32_t sl = slots[(size_t)set * K + k];
                em.rec(REC_ASET | (k << REC_K_SHIFT), sl, 0.0, PCMD_ATILE | sl);
                s.n_asets += 1;
            }
            while (i < ts.size() && set_of_slot[ts[i].slot] == set) {
                const int32_t in_col = ts[i].in_col;
                size_t j = i;
                while (j < ts.size() && set_of_slot[ts[j].slot] == set && ts[j].in_col == in_col) ++j;
                const bool last_job = (j == ts.size());
                em.rec(REC_JOB | (last_job ? REC_F_LAST : 0), in_col, 0.0, PCMD_WINDOW | in_col);
                s.n_jobs += 1;
                int32_t prev_k = -1;
                for (size_t e = i; e < j; ++e) {
                    const int32_t k = k_of_slot[ts[e].slot];
                    int32_t a = REC_ENTRY | (k << REC_K_SHIFT);
                    if (k != prev_k) { a |= REC_F_FIRST; s.n_stencils += 1; }
                    prev_k = k;
                    if (!touched[ts[e].out]) { a |= REC_F_INIT; touched[ts[e].out] = 1; ++n_touched; }
                    em.rec(a, ts[e].out * SCHED_ACC_STRIDE, ts[e].coef);
                    s.n_entries += 1;
                }
                if (last_job && n_touched < ncols) {
                    // output columns without any term (possible for an M-shard): zero them with a
                    // null entry that reuses the last stencil value
                    for (int32_t o = 0; o < ncols; ++o)
                        if (!touched[o]) {
                            em.rec(REC_ENTRY | (prev_k << REC_K_SHIFT) | REC_F_INIT, o * SCHED_ACC_STRIDE, 0.0);
                            touched[o] = 1; ++n_touched; s.n_entries += 1;
                        }
                }
                i = j;
            }
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
    std::vector<double> acc((size_t)C * N), S(N);
    std::vector<int32_t> set_slot(std::max(s.K, 1), -1);
    // the producer command list must name exactly the fetches the records imply, in order
    for (int32_t b = 0; b < s.nblk; ++b) {
        size_t r = 0;
        // first chunk of the block = value of the first META command
        int32_t pc = s.blk_pcmd[b];
        const int32_t pc_end = s.blk_pcmd[b + 1];
        int32_t chunk0 = -1;
        std::vector<int32_t> fetches;
        for (int32_t p = pc; p < pc_end; ++p) {
            const int32_t cmd = s.pcmd[p];
            if ((cmd & PCMD_KIND_MASK) == PCMD_META) { if (chunk0 < 0) chunk0 = cmd & PCMD_VAL_MASK; }
            else fetches.push_back(cmd);
        }
        r = (size_t)chunk0 * SCHED_CHUNK_RECS;
        size_t f = 0;
        int32_t cur_col = -1;
        bool running = true;
        while (running) {
            const SchedRec& rec = s.recs[r++];
            const int32_t type = rec.a & REC_TYPE_MASK;
            const int32_t k = (rec.a >> REC_K_SHIFT) & 0xF;
            switch (type) {
            case REC_END_ITEM: running = false; break;
            case REC_ASET:
                set_slot[k] = rec.b;
                if (f >= fetches.size() || fetches[f++] != (PCMD_ATILE | rec.b)) { Y[0] = 0.0 / 0.0; return; }
                break;
            case REC_JOB:
                cur_col = rec.b;
                if (f >= fetches.size() || fetches[f++] != (PCMD_WINDOW | rec.b)) { Y[0] = 0.0 / 0.0; return; }
                break;
            case REC_ENTRY: {
                if (rec.a & REC_F_FIRST) {
                    const double* a = dia + (size_t)set_slot[k] * D * N;
                    const double* w = W + (int64_t)cur_col * ldw;
                    for (int32_t iy = 0; iy < ny; ++iy)
                        for (int32_t ix = 0; ix < nx; ++ix) {
                            const int64_t i = (int64_t)iy * nx + ix;
                            double v = 0.0;
                            for (int32_t d = 0; d < D; ++d) {
                                const int32_t jx = ix + dx[d], jy = iy + dy[d];
                                const double wv = (jx >= 0 && jx < nx && jy >= 0 && jy < ny) ? w[(int64_t)jy * nx + jx] : 0.0;
                                v += a[(size_t)d * N + i] * wv;
                            }
                            S[i] = v;
                        }
                }
                double* ac = &acc[(size_t)(rec.b / SCHED_ACC_STRIDE) * N];
                if (rec.a & REC_F_INIT) for (int64_t i = 0; i < N; ++i) ac[i] = rec.c * S[i];
                else for (int64_t i = 0; i < N; ++i) ac[i] += rec.c * S[i];
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
            default: break;
            }
        }
        if (f != fetches.size()) { Y[0] = 0.0 / 0.0; return; }
    }
}

}  // namespace spuq
