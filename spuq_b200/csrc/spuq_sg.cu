// spuq_sg.cu -- plan construction and launch of the SG operator application (K1) behind the C ABI
// of include/spuq_sg.h.  Reference interface replaced: MultiOperator.__init__/apply,
// application/egsz/multi_operator2.py:30-84.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>

#include "common.h"
#include "kernels_apply.cuh"
#include "kernels_tiled.cuh"
#include "kernels_tma.cuh"
#include "kernels_tmem.cuh"
#include "schedule.h"

#ifdef SPUQ_EMU
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#endif

namespace spuq {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static thread_local int g_launch_counter = 0;
int& launch_counter() { return g_launch_counter; }
static thread_local const double* g_abi_gate = nullptr;
const double* abi_gate() { return g_abi_gate; }

int sm_count(int device) {
#ifdef SPUQ_EMU
    (void)device;
    return emu_sm_count();
#else
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    return n;
#endif
}

template <typename T>
static int dev_upload(spuq_sg_plan* p, const std::vector<T>& h, T** out) {
    *out = nullptr;
    void* d = nullptr;
    SPUQ_CUDA_TRY(cudaMalloc(&d, std::max<size_t>(h.size() * sizeof(T), 16)));
    p->owned_dev.push_back(d);
    if (!h.empty()) SPUQ_CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<T*>(d);
    return SPUQ_OK;
}

static int dev_alloc(spuq_sg_plan* p, size_t bytes, void** out) {
    *out = nullptr;
    cudaError_t e = cudaMalloc(out, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return SPUQ_E_NOMEM;
    }
    p->owned_dev.push_back(*out);
    return SPUQ_OK;
}

// ---- pattern analysis: is the shared CSR pattern a 3x3-window stencil on an s x ny grid? -------
struct StencilShape {
    bool ok = false;
    int32_t s = 0, ny = 0;
    bool nine = false;
};

static bool check_grid(const std::vector<int32_t>& rowptr, const std::vector<int32_t>& colidx,
                       int64_t N, int64_t s, bool* nine) {
    if (s < 3 || N % s != 0) return false;
    *nine = false;
    for (int64_t i = 0; i < N; ++i) {
        const int64_t ix = i % s, iy = i / s;
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int64_t j = colidx[k];
            const int64_t dx = j % s - ix, dy = j / s - iy;
            if (dx < -1 || dx > 1 || dy < -1 || dy > 1) return false;
            if (dx != 0 && dy != 0) *nine = true;
        }
    }
    return true;
}

static StencilShape analyse_pattern(const std::vector<int32_t>& rowptr,
                                    const std::vector<int32_t>& colidx, int64_t n_rows,
                                    int64_t n_cols) {
    StencilShape sh;
    if (n_rows != n_cols || n_rows < 3 || n_rows > 0x7fffffffLL) return sh;
    const int64_t N = n_rows;
    int64_t big = 0;   // smallest offset magnitude > 1
    bool any_far = false;
    for (int64_t i = 0; i < N; ++i)
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            int64_t o = (int64_t)colidx[k] - i;
            if (o < 0) o = -o;
            if (o > 1 && (!any_far || o < big)) { big = o; any_far = true; }
        }
    std::vector<int64_t> cand;
    if (!any_far) cand = {N};                   // tridiagonal: one grid row
    else cand = {big, big + 1, big - 1};
    for (int64_t s : cand) {
        bool nine = false;
        if (check_grid(rowptr, colidx, N, s, &nine)) {
            sh.ok = true; sh.s = (int32_t)s; sh.ny = (int32_t)(N / s); sh.nine = nine;
            return sh;
        }
    }
    return sh;
}

// ---- term lists -----------------------------------------------------------------------------------
static void build_terms(spuq_sg_plan* p, const int32_t* nbr_up, const int32_t* nbr_dn,
                        const double* beta_up, const double* beta_dn, const double* beta_0) {
    const int32_t L = p->L;
    p->h_term_ptr.assign(L + 1, 0);
    for (int32_t mu = 0; mu < L; ++mu) {
        if (p->include_mean) {
            p->h_term_mat.push_back(0); p->h_term_col.push_back(mu); p->h_term_coef.push_back(1.0);
        }
        for (int32_t m = p->m_begin; m < p->m_end; ++m) {
            const int64_t q = (int64_t)m * L + mu;
            // same order as the reference builds cur_w (multi_operator2.py:69-79)
            if (beta_0 && beta_0[q] != 0.0) {
                p->h_term_mat.push_back(1 + m); p->h_term_col.push_back(mu);
                p->h_term_coef.push_back(-beta_0[q]);
            }
            if (nbr_up[q] >= 0) {
                p->h_term_mat.push_back(1 + m); p->h_term_col.push_back(nbr_up[q]);
                p->h_term_coef.push_back(beta_up[q]);
            }
            if (nbr_dn[q] >= 0) {
                p->h_term_mat.push_back(1 + m); p->h_term_col.push_back(nbr_dn[q]);
                p->h_term_coef.push_back(beta_dn[q]);
            }
        }
        p->h_term_ptr[mu + 1] = (int32_t)p->h_term_mat.size();
    }
    p->n_terms = (int64_t)p->h_term_mat.size();
}

// ---- column blocks of the tiled kernel: C consecutive columns (sorted-Lambda order keeps
// indices with a common leading part adjacent, which is what makes groups share a block matrix)
static int build_blocks(spuq_sg_plan* p, int32_t C) {
    const int32_t L = p->L;
    const int32_t nblk = (L + C - 1) / C;
    std::vector<int32_t> bcol((size_t)nblk * C, -1), bgrp(nblk + 1, 0), gmat, gterm, tcol, tacc;
    std::vector<double> tcoef;
    struct T { int32_t slot, col, acc; double coef; };
    std::vector<T> ts;
    for (int32_t b = 0; b < nblk; ++b) {
        ts.clear();
        for (int32_t c = 0; c < C; ++c) {
            const int32_t mu = b * C + c;
            if (mu >= L) break;
            bcol[(size_t)b * C + c] = mu;
            for (int32_t t = p->h_term_ptr[mu]; t < p->h_term_ptr[mu + 1]; ++t)
                ts.push_back({p->slot_of_mat[p->h_term_mat[t]], p->h_term_col[t], c, p->h_term_coef[t]});
        }
        std::stable_sort(ts.begin(), ts.end(), [](const T& a, const T& b2) {
            if (a.slot != b2.slot) return a.slot < b2.slot;
            return a.col < b2.col;
        });
        for (size_t i = 0; i < ts.size(); ++i) {
            if (i == 0 || ts[i].slot != ts[i - 1].slot) {
                gmat.push_back(ts[i].slot);
                gterm.push_back((int32_t)tcol.size());
            }
            tcol.push_back(ts[i].col); tacc.push_back(ts[i].acc); tcoef.push_back(ts[i].coef);
        }
        bgrp[b + 1] = (int32_t)gmat.size();
    }
    gterm.push_back((int32_t)tcol.size());
    for (void* d : p->block_dev) cudaFree(d);
    p->block_dev.clear();
    p->bv = BlockView{};
    int rc;
    int32_t *d_bcol, *d_bgrp, *d_gmat, *d_gterm, *d_tcol, *d_tacc; double* d_tcoef;
    struct TRec { int32_t col, acc; double coef; };
    static_assert(sizeof(TRec) == 16, "term record must be 16 bytes");
    std::vector<TRec> trec(tcol.size());
    for (size_t i = 0; i < tcol.size(); ++i) trec[i] = TRec{tcol[i], tacc[i], tcoef[i]};
    TRec* d_trec;
    const size_t first = p->owned_dev.size();
    if ((rc = dev_upload(p, bcol, &d_bcol))) return rc;
    if ((rc = dev_upload(p, bgrp, &d_bgrp))) return rc;
    if ((rc = dev_upload(p, gmat, &d_gmat))) return rc;
    if ((rc = dev_upload(p, gterm, &d_gterm))) return rc;
    if ((rc = dev_upload(p, tcol, &d_tcol))) return rc;
    if ((rc = dev_upload(p, tacc, &d_tacc))) return rc;
    if ((rc = dev_upload(p, tcoef, &d_tcoef))) return rc;
    if ((rc = dev_upload(p, trec, &d_trec))) return rc;
    p->block_dev.assign(p->owned_dev.begin() + first, p->owned_dev.end());
    p->owned_dev.resize(first);
    p->bv.C = C; p->bv.nblk = nblk;
    p->bv.bcol = d_bcol; p->bv.bgrp = d_bgrp; p->bv.gmat = d_gmat; p->bv.gterm = d_gterm;
    p->bv.tcol = d_tcol; p->bv.tacc = d_tacc; p->bv.tcoef = d_tcoef; p->bv.trec = d_trec;
    return SPUQ_OK;
}



// ---- schedule of the TMEM-accumulator kernel (schedule.h), rebuilt when K changes ----------------
static int build_schedule(spuq_sg_plan* p, int32_t K, int32_t block_cols, int32_t w_group) {
    std::vector<int32_t> slot(p->h_term_mat.size());
    for (size_t t = 0; t < slot.size(); ++t) slot[t] = p->slot_of_mat[p->h_term_mat[t]];
    TermLists tl;
    tl.L = p->L; tl.ptr = p->h_term_ptr.data(); tl.slot = slot.data();
    tl.col = p->h_term_col.data(); tl.coef = p->h_term_coef.data();
    std::vector<int32_t> order;
    schedule_column_order(p->L, p->M, p->h_nbr_up.empty() ? nullptr : p->h_nbr_up.data(),
                          p->h_nbr_dn.empty() ? nullptr : p->h_nbr_dn.data(), block_cols, &order);
    Schedule* sc = static_cast<Schedule*>(p->sched_host);
    if (!sc) { sc = new Schedule(); p->sched_host = sc; }
    schedule_build(tl, order, K, p->n_slots, block_cols, w_group, sc);
    if (std::getenv("SPUQ_SCHED_STATS")) {
        fprintf(stderr, "[spuq] schedule K=%d C=%d: blocks %d, jobs %lld (pairs %lld), A tiles %lld, stencils %lld, terms %lld, "
                        "units %zu, producer cmds %zu, inputs %lld, job bound %lld\n", K, block_cols, sc->nblk, (long long)sc->n_jobs, (long long)sc->n_pairs,
                (long long)sc->n_asets, (long long)sc->n_stencils, (long long)sc->n_entries, sc->recs.size(), sc->pcmd.size(),
                (long long)sc->n_inputs, (long long)sc->n_jobs_bound);
        for (int S = 0; S < 3; ++S)
            for (int km = 1; km < 8; ++km)
                if (sc->hist_pairs[S][km] || sc->hist_singles[S][km])
                    fprintf(stderr, "[spuq]   S=%d kmask=%d: pairs %lld singles %lld\n", S + 1, km,
                            (long long)sc->hist_pairs[S][km], (long long)sc->hist_singles[S][km]);
    }
    for (void* d : p->sched_dev) cudaFree(d);
    p->sched_dev.clear();
    p->d_recs = nullptr; p->d_pcmd = nullptr; p->d_blk_pcmd = nullptr; p->d_work_counter = nullptr;
    p->sched_K = 0; p->sched_C = 0;
    int rc;
    SchedRec* d_recs; int32_t *d_pcmd, *d_blk; 
    const size_t first = p->owned_dev.size();
    std::vector<unsigned long long> zero(2, 0ull);
    unsigned long long* d_cnt;
    if ((rc = dev_upload(p, sc->recs, &d_recs))) return rc;
    if ((rc = dev_upload(p, sc->pcmd, &d_pcmd))) return rc;
    if ((rc = dev_upload(p, sc->blk_pcmd, &d_blk))) return rc;
    if ((rc = dev_upload(p, zero, &d_cnt))) return rc;
    p->sched_dev.assign(p->owned_dev.begin() + first, p->owned_dev.end());
    p->owned_dev.resize(first);
    p->d_recs = d_recs; p->d_pcmd = d_pcmd; p->d_blk_pcmd = d_blk; p->d_work_counter = d_cnt;
    p->sched_K = K;
    p->sched_C = block_cols;
    return SPUQ_OK;
}

}  // namespace spuq

using namespace spuq;

extern "C" const char* spuq_sg_last_error(void) { return g_err; }
extern "C" int spuq_sg_abi_version(void) { return SPUQ_SG_ABI_VERSION; }

extern "C" int spuq_sg_plan_create(spuq_sg_plan** plan_out, int device, int64_t n_rows,
                                   int64_t n_cols, int32_t L, int32_t M,
                                   const int32_t* rowptr_dev, const int32_t* colidx_dev,
                                   const double* const* vals_host, const int32_t* nbr_up,
                                   const int32_t* nbr_dn, const double* beta_up,
                                   const double* beta_dn, const double* beta_0, int32_t m_begin,
                                   int32_t m_end, int include_mean, int32_t flags) {
    SPUQ_REQUIRE(plan_out != nullptr, "plan_create: plan_out is null");
    *plan_out = nullptr;
    SPUQ_REQUIRE(n_rows > 0 && n_cols > 0 && n_rows < 0x7fffffffLL && n_cols < 0x7fffffffLL,
                 "plan_create: bad block shape %lld x %lld", (long long)n_rows, (long long)n_cols);
    SPUQ_REQUIRE(L > 0 && M >= 0, "plan_create: need L > 0 and M >= 0 (got L=%d M=%d)", L, M);
    SPUQ_REQUIRE(M == 0 || n_rows == n_cols, "plan_create: SG operator blocks must be square");
    SPUQ_REQUIRE(rowptr_dev && vals_host, "plan_create: null CSR pointer");
    SPUQ_REQUIRE(0 <= m_begin && m_begin <= m_end && m_end <= M,
                 "plan_create: bad shard range [%d, %d) for M=%d", m_begin, m_end, M);
    SPUQ_REQUIRE(M == 0 || (nbr_up && nbr_dn && beta_up && beta_dn),
                 "plan_create: neighbour / beta tables missing");
    for (int64_t q = 0; q < (int64_t)M * L; ++q) {
        SPUQ_REQUIRE(nbr_up[q] >= -1 && nbr_up[q] < L && nbr_dn[q] >= -1 && nbr_dn[q] < L,
                     "plan_create: neighbour table entry %lld out of range", (long long)q);
    }

    SPUQ_CUDA_TRY(cudaSetDevice(device));
    spuq_sg_plan* p = new spuq_sg_plan();
    struct Guard { spuq_sg_plan* p; ~Guard() { if (p) spuq_sg_plan_destroy(p); } } guard{p};
    p->device = device; p->n_rows = n_rows; p->n_cols = n_cols; p->L = L; p->M = M;
    p->m_begin = m_begin; p->m_end = m_end; p->include_mean = include_mean ? 1 : 0;
    p->rowptr = rowptr_dev; p->colidx = colidx_dev;
    p->vals_host.assign(vals_host, vals_host + M + 1);
    p->n_sm = sm_count(device);

    // pattern to the host for analysis
    std::vector<int32_t> h_rowptr(n_rows + 1);
    SPUQ_CUDA_TRY(cudaMemcpy(h_rowptr.data(), rowptr_dev, sizeof(int32_t) * (n_rows + 1), cudaMemcpyDeviceToHost));
    SPUQ_REQUIRE(h_rowptr[0] == 0, "plan_create: rowptr[0] must be 0");
    p->nnz = h_rowptr[n_rows];
    SPUQ_REQUIRE(p->nnz == 0 || colidx_dev, "plan_create: null column index array");
    SPUQ_REQUIRE(p->nnz == 0 || !include_mean || vals_host[0], "plan_create: mean block values missing");
    for (int32_t m = m_begin; m < m_end; ++m)
        SPUQ_REQUIRE(p->nnz == 0 || vals_host[1 + m], "plan_create: values of block m=%d missing", m);
    for (int64_t i = 0; i < n_rows; ++i) {
        SPUQ_REQUIRE(h_rowptr[i + 1] >= h_rowptr[i], "plan_create: rowptr not monotone at row %lld", (long long)i);
        p->max_row_nnz = std::max(p->max_row_nnz, h_rowptr[i + 1] - h_rowptr[i]);
    }
    std::vector<int32_t> h_colidx((size_t)p->nnz);
    if (p->nnz) SPUQ_CUDA_TRY(cudaMemcpy(h_colidx.data(), colidx_dev, sizeof(int32_t) * p->nnz, cudaMemcpyDeviceToHost));
    for (int64_t k = 0; k < p->nnz; ++k)
        SPUQ_REQUIRE(h_colidx[k] >= 0 && h_colidx[k] < n_cols, "plan_create: column index out of range at %lld", (long long)k);

    // owned blocks -> compact slots
    p->slot_of_mat.assign(M + 1, -1);
    if (p->include_mean) p->slot_of_mat[0] = p->n_slots++;
    for (int32_t m = m_begin; m < m_end; ++m) p->slot_of_mat[1 + m] = p->n_slots++;

    build_terms(p, nbr_up, nbr_dn, beta_up, beta_dn, beta_0);
    if (M > 0) {
        p->h_nbr_up.assign(nbr_up, nbr_up + (size_t)M * L);
        p->h_nbr_dn.assign(nbr_dn, nbr_dn + (size_t)M * L);
    }
    int rc;
    if ((rc = dev_upload(p, p->h_term_ptr, &p->term_ptr))) return rc;
    if ((rc = dev_upload(p, p->h_term_mat, &p->term_mat))) return rc;
    if ((rc = dev_upload(p, p->h_term_col, &p->term_col))) return rc;
    if ((rc = dev_upload(p, p->h_term_coef, &p->term_coef))) return rc;
    {
        std::vector<const double*> ptrs(p->vals_host);
        const double** d = nullptr;
        if ((rc = dev_upload(p, ptrs, const_cast<const double***>(&d)))) return rc;
        p->vals_ptrs_dev = d;
    }
    int32_t* d_slot = nullptr;
    if ((rc = dev_upload(p, p->slot_of_mat, &d_slot))) return rc;

    // reporting numbers: compulsory bytes, non-trivial flops
    {
        std::set<int32_t> active;
        for (int32_t mat : p->h_term_mat) active.insert(mat);
        p->alg_bytes = 8 * (n_cols + n_rows) * (int64_t)L + 8 * p->nnz * (int64_t)active.size()
                       + 4 * p->nnz + 4 * (n_rows + 1) + (int64_t)(m_end - m_begin) * L * (2 * 4 + 3 * 8);
        p->alg_flops = 2 * p->nnz * p->n_terms;
    }

    // ---- choose the path -----------------------------------------------------------------------
    const int32_t want = flags & SPUQ_PATH_MASK;
    StencilShape sh;
    if (want == SPUQ_PATH_AUTO || want == SPUQ_PATH_STENCIL) sh = analyse_pattern(h_rowptr, h_colidx, n_rows, n_cols);
    if (want == SPUQ_PATH_STENCIL && !sh.ok) {
        set_error("plan_create: pattern is not a 3x3-window grid stencil; stencil path unavailable");
        return SPUQ_E_UNSUPPORTED;
    }
    const bool ell_ok = p->max_row_nnz <= 64 &&
                        ((int64_t)p->max_row_nnz * n_rows <= 3 * p->nnz || (int64_t)p->max_row_nnz * n_rows <= (1 << 20));
    if (want == SPUQ_PATH_ELL && !ell_ok) {
        set_error("plan_create: ELL padding too large (max row nnz %d)", p->max_row_nnz);
        return SPUQ_E_UNSUPPORTED;
    }
    if (want == SPUQ_PATH_CSR) p->path = SPUQ_PATH_CSR;
    else if (want == SPUQ_PATH_ELL) p->path = SPUQ_PATH_ELL;
    else if (sh.ok) p->path = SPUQ_PATH_STENCIL;
    else p->path = ell_ok ? SPUQ_PATH_ELL : SPUQ_PATH_CSR;

    const int threads = 128;
    const unsigned row_blocks = (unsigned)((n_rows + threads - 1) / threads);
    if (p->path == SPUQ_PATH_ELL) {
        const int32_t K = std::max(p->max_row_nnz, 1);
        p->ell_K = K;
        void* d;
        if ((rc = dev_alloc(p, sizeof(int32_t) * (size_t)K * n_rows, &d))) return rc;
        p->ell_cols = static_cast<int32_t*>(d);
        if ((rc = dev_alloc(p, sizeof(double) * (size_t)p->n_slots * K * n_rows, &d))) return rc;
        p->ell_vals = static_cast<double*>(d);
        bool cols_done = false;
        for (int32_t mat = 0; mat <= M; ++mat) {
            if (p->slot_of_mat[mat] < 0) continue;
            double* dst = p->ell_vals + (size_t)p->slot_of_mat[mat] * K * n_rows;
            SPUQ_LAUNCH(k_csr_to_ell, dim3(row_blocks), dim3(threads), 0, 0, n_rows, K, rowptr_dev, colidx_dev,
                        p->vals_host[mat], cols_done ? (int32_t*)nullptr : p->ell_cols, dst);
            cols_done = true;
        }
        if (!cols_done)
            SPUQ_LAUNCH(k_csr_to_ell, dim3(row_blocks), dim3(threads), 0, 0, n_rows, K, rowptr_dev, colidx_dev,
                        (const double*)nullptr, p->ell_cols, (double*)nullptr);
        SPUQ_CUDA_TRY(cudaGetLastError());
        SPUQ_CUDA_TRY(cudaDeviceSynchronize());
    } else if (p->path == SPUQ_PATH_STENCIL) {
        StencilView& st = p->st;
        st.s = sh.s; st.N = n_rows;
        st.D = sh.nine ? 9 : 5;
        int d = 0;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                if (!sh.nine && dx != 0 && dy != 0) continue;
                st.dx[d] = dx; st.dy[d] = dy; st.off[d] = dx + sh.s * dy;
                ++d;
            }
        void* dv;
        if ((rc = dev_alloc(p, sizeof(double) * (size_t)p->n_slots * st.D * n_rows, &dv))) return rc;
        p->dia_vals = static_cast<double*>(dv);
        st.vals = p->dia_vals;
        for (int32_t mat = 0; mat <= M; ++mat) {
            if (p->slot_of_mat[mat] < 0) continue;
            double* dst = p->dia_vals + (size_t)p->slot_of_mat[mat] * st.D * n_rows;
            SPUQ_LAUNCH(k_csr_to_dia, dim3(row_blocks), dim3(threads), 0, 0, st, rowptr_dev, colidx_dev,
                        p->vals_host[mat], dst);
        }
        SPUQ_CUDA_TRY(cudaGetLastError());
        SPUQ_CUDA_TRY(cudaDeviceSynchronize());
    }
    p->slot_dev = d_slot;
    p->ny = sh.ok ? sh.ny : 0;
    p->nine = sh.nine;
    guard.p = nullptr;
    *plan_out = p;
    return SPUQ_OK;
}

extern "C" int spuq_sg_plan_destroy(spuq_sg_plan* p) {
    if (!p) return SPUQ_OK;
    cudaSetDevice(p->device);
    for (void* d : p->block_dev) cudaFree(d);
    for (void* d : p->sched_dev) cudaFree(d);
    for (void* d : p->owned_dev) cudaFree(d);
    delete static_cast<spuq::Schedule*>(p->sched_host);
    delete p;
    return SPUQ_OK;
}

extern "C" int spuq_sg_plan_info(const spuq_sg_plan* p, int32_t* path, int64_t* n_terms,
                                 int64_t* nnz, int64_t* alg_bytes, int64_t* alg_flops) {
    SPUQ_REQUIRE(p != nullptr, "plan_info: null plan");
    if (path) *path = p->path;
    if (n_terms) *n_terms = p->n_terms;
    if (nnz) *nnz = p->nnz;
    if (alg_bytes) *alg_bytes = p->alg_bytes;
    if (alg_flops) *alg_flops = p->alg_flops;
    return SPUQ_OK;
}

extern "C" int spuq_sg_plan_tune(spuq_sg_plan* p, int32_t variant, int32_t cols_per_block,
                                 int32_t rows_per_thread, int32_t ctas_per_sm) {
    SPUQ_REQUIRE(p != nullptr, "plan_tune: null plan");
    SPUQ_REQUIRE(variant >= 0 && variant <= 4,
                 "plan_tune: variant must be 0 (auto), 1 (simple), 2 (tiled), 3 (tma) or 4 (tmem)");
    p->variant = variant; p->cols_per_block = cols_per_block;
    p->rows_per_thread = rows_per_thread; p->ctas_per_sm = ctas_per_sm;
    return SPUQ_OK;
}

extern "C" int spuq_sg_plan_last_launches(const spuq_sg_plan* p) { return p ? p->last_launches : 0; }
extern "C" const char* spuq_sg_plan_kernel_name(const spuq_sg_plan* p) { return p ? p->kernel_name.c_str() : ""; }

namespace spuq {

static int32_t pick_cols_per_cta(const spuq_sg_plan* p, int64_t row_blocks) {
    const int64_t target = 16LL * p->n_sm;
    int64_t cpc = row_blocks * p->L / target;
    if (cpc < 1) cpc = 1;
    if (cpc > 16) cpc = 16;
    return (int32_t)cpc;
}

#ifndef SPUQ_EMU
// ---- TMA descriptors ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D fp64 tensor (x fastest): dims {nx, ny, nz}, strides in elements {1, sy, sz}, box {bx, by, bz}
static int encode_3d(void* out, const void* base, int64_t nx, int64_t ny, int64_t nz, int64_t sy, int64_t sz,
                     int bx, int by, int bz) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return SPUQ_E_CUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
    cuuint64_t strides[2] = {(cuuint64_t)sy * 8, (cuuint64_t)sz * 8};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base),
                    dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return SPUQ_E_CUDA; }
    return SPUQ_OK;
}

template <int RY, int WY, int C, bool NINE>
static int launch_tma(spuq_sg_plan* p, const double* W, int64_t ldw, double* Y, int64_t ldy,
                      cudaStream_t stream, const double* gate) {
    using Cfg = TmaCfg<RY, WY, NINE>;
    int rc;
    if (p->tmap_a_rows != Cfg::TILE_Y) {
        if ((rc = encode_3d(p->tmap_a, p->dia_vals, p->st.s, p->ny, (int64_t)p->n_slots * Cfg::D, p->st.s, p->st.N,
                            Cfg::TILE_X, Cfg::TILE_Y, Cfg::D))) return rc;
        p->tmap_a_rows = Cfg::TILE_Y;
    }
    if (p->tmap_w_ptr != W || p->tmap_w_ld != ldw || p->tmap_w_rows != Cfg::WIN_ROWS) {
        if ((rc = encode_3d(p->tmap_w, W, p->st.s, p->ny, p->L, p->st.s, ldw, Cfg::BOX_W, Cfg::WIN_ROWS, 1))) return rc;
        p->tmap_w_ptr = W; p->tmap_w_ld = ldw; p->tmap_w_rows = Cfg::WIN_ROWS;
    }
    auto kern = k_apply_tma<RY, WY, C, NINE>;
    static bool attr_done = false;
    static int ctas_per_sm = 1;
    if (!attr_done) {
        SPUQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        SPUQ_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, Cfg::THREADS, Cfg::SMEM_BYTES));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        attr_done = true;
    }
    const int32_t ntx = (p->st.s + Cfg::TILE_X - 1) / Cfg::TILE_X;
    const int32_t nty = (p->ny + Cfg::TILE_Y - 1) / Cfg::TILE_Y;
    const int64_t items = (int64_t)ntx * nty * p->bv.nblk;
    const int per_sm = p->ctas_per_sm > 0 ? std::min(p->ctas_per_sm, ctas_per_sm) : ctas_per_sm;
    const int64_t grid = std::min<int64_t>(items, (int64_t)p->n_sm * per_sm);
    static const int32_t debug_flags = std::getenv("SPUQ_KDEBUG") ? std::atoi(std::getenv("SPUQ_KDEBUG")) : 0;
    CUtensorMap mw, ma;
    std::memcpy(&mw, p->tmap_w, sizeof(mw));
    std::memcpy(&ma, p->tmap_a, sizeof(ma));
    kern<<<dim3((unsigned)grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream>>>(mw, ma, p->bv, p->st.s, p->ny, ntx, nty,
                                                                                items, Y, ldy, gate);
    char name[80];
    snprintf(name, sizeof(name), "k_apply_tma<RY=%d,WY=%d,C=%d,%s>x%d", RY, WY, C, NINE ? "9pt" : "5pt", per_sm);
    p->kernel_name = name;
    return SPUQ_OK;
}

template <int K, bool NINE, int G>
static int launch_tmem(spuq_sg_plan* p, const double* W, int64_t ldw, double* Y, int64_t ldy,
                       cudaStream_t stream, const double* gate) {
    using Cfg = TmemCfg<K, NINE, G>;
    int rc;
    if (p->tmap_a_rows != Cfg::TILE_Y) {
        if ((rc = encode_3d(p->tmap_a, p->dia_vals, p->st.s, p->ny, (int64_t)p->n_slots * Cfg::D, p->st.s, p->st.N,
                            Cfg::TILE_X, Cfg::TILE_Y, Cfg::D))) return rc;
        p->tmap_a_rows = Cfg::TILE_Y;
    }
    if (p->tmap_w_ptr != W || p->tmap_w_ld != ldw || p->tmap_w_rows != Cfg::WIN_ROWS) {
        if ((rc = encode_3d(p->tmap_w, W, p->st.s, p->ny, p->L, p->st.s, ldw, Cfg::BOX_W, Cfg::WIN_ROWS, 1))) return rc;
        p->tmap_w_ptr = W; p->tmap_w_ld = ldw; p->tmap_w_rows = Cfg::WIN_ROWS;
    }
    auto kern = k_apply_tmem<K, NINE, G>;
    // the attribute is per device: set it before every launch (cheap) instead of caching a process-wide flag
    SPUQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const Schedule* sc = static_cast<const Schedule*>(p->sched_host);
    const int32_t ntx = (p->st.s + Cfg::TILE_X - 1) / Cfg::TILE_X;
    const int32_t nty_all = (p->ny + Cfg::TILE_Y - 1) / Cfg::TILE_Y;
    // tile rows of this launch (spuq_sg_apply_part): a tile row t reads the grid rows t*TILE_Y - 1 ..
    // t*TILE_Y + TILE_Y, so it touches a halo grid row (first / last row of the local grid) when t == 0 or
    // t*TILE_Y + TILE_Y >= ny - 1
    int32_t t_hi = (p->ny - 1 - Cfg::TILE_Y + Cfg::TILE_Y - 1) / Cfg::TILE_Y;      // first t with t*TY + TY >= ny - 1
    t_hi = std::max(1, std::min(t_hi, nty_all));
    int32_t off0 = 0, cnt0 = nty_all, off1 = 0, nty = nty_all;
    if (p->part == 1) { off0 = 1; cnt0 = t_hi - 1; nty = cnt0; }                              // interior
    else if (p->part == 2) { off0 = 0; cnt0 = 1; off1 = t_hi; nty = 1 + (nty_all - t_hi); }   // boundary
    if (nty_all < 3 && p->part == 1) return SPUQ_OK;                  // nothing is interior
    if (nty_all < 3 && p->part == 2) { off0 = 0; cnt0 = nty_all; nty = nty_all; }
    if (nty <= 0) return SPUQ_OK;
    const int64_t items = (int64_t)ntx * nty * sc->nblk;
    const int64_t grid = std::min<int64_t>((items + Cfg::G - 1) / Cfg::G, (int64_t)p->n_sm);
    // one counter per part: the interior and the boundary launch of one apply may be in flight together
    unsigned long long* counter = p->d_work_counter + (p->part == 2 ? 1 : 0);
    SPUQ_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
    static const int32_t debug_flags = std::getenv("SPUQ_KDEBUG") ? std::atoi(std::getenv("SPUQ_KDEBUG")) : 0;
    CUtensorMap mw, ma;
    std::memcpy(&mw, p->tmap_w, sizeof(mw));
    std::memcpy(&ma, p->tmap_a, sizeof(ma));
    kern<<<dim3((unsigned)grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream>>>(
        mw, ma, static_cast<const SchedRec*>(p->d_recs), p->d_pcmd, p->d_blk_pcmd, sc->nblk, p->st.s, p->ny, nty,
        off0, cnt0, off1, items, counter, Y, ldy, gate, debug_flags);
    char name[80];
    snprintf(name, sizeof(name), "k_apply_tmem<K=%d,G=%d,%s>", K, G, NINE ? "9pt" : "5pt");
    p->kernel_name = name;
    return SPUQ_OK;
}
#endif  // !SPUQ_EMU

template <int RY, int C, bool NINE>
static void launch_tiled(spuq_sg_plan* p, const double* W, int64_t ldw, double* Y, int64_t ldy,
                         cudaStream_t stream, const double* gate) {
    using Cfg = TiledCfg<RY, C, NINE>;
    const int32_t ntx = (p->st.s + Cfg::TILE_X - 1) / Cfg::TILE_X;
    const int32_t nty = (p->ny + Cfg::TILE_Y - 1) / Cfg::TILE_Y;
    const int64_t items = (int64_t)ntx * nty * p->bv.nblk;
    SPUQ_LAUNCH((k_apply_tiled<RY, C, NINE>), dim3((unsigned)items), dim3(Cfg::THREADS), 0, stream,
                p->st, p->bv, p->ny, ntx, nty, W, ldw, Y, ldy, gate);
    char name[64];
    snprintf(name, sizeof(name), "k_apply_tiled<RY=%d,C=%d,%s>", RY, C, NINE ? "9pt" : "5pt");
    p->kernel_name = name;
}

int apply_impl(spuq_sg_plan* p, const double* W, int64_t ldw, double* Y, int64_t ldy,
               cudaStream_t stream, const double* gate) {
    SPUQ_REQUIRE(p && W && Y, "apply: null argument");
    SPUQ_REQUIRE(ldw >= p->n_cols && ldy >= p->n_rows, "apply: leading dimension too small");
    SPUQ_REQUIRE(W != Y, "apply: input and output slab must not alias");
    p->last_launches = 0;
    const TermView tv{p->term_ptr, p->term_mat, p->term_col, p->term_coef};
    const int threads = 128;
    if (p->path == SPUQ_PATH_CSR || p->path == SPUQ_PATH_ELL) {
        const int64_t row_blocks = (p->n_rows + threads - 1) / threads;
        const int32_t cpc = p->cols_per_block > 0 ? p->cols_per_block : pick_cols_per_cta(p, row_blocks);
        const dim3 grid((unsigned)row_blocks, (unsigned)((p->L + cpc - 1) / cpc));
        SPUQ_REQUIRE(grid.y <= 65535u, "apply: too many column chunks (%u); raise cols_per_block", grid.y);
        // column-blocked form (A values of a row in registers across the terms of a matrix) whenever the
        // operator has expansion terms and the padded row fits the register budget; variant 1 keeps
        // the one-column-at-a-time kernels (cross-check)
        if (p->path == SPUQ_PATH_ELL && p->M > 0 && p->ell_K <= 16 && p->variant != 1) {
            if (p->bv.bcol == nullptr || p->bv.C != ELLBLK_C) {
                int rc = build_blocks(p, ELLBLK_C);
                if (rc) return rc;
            }
            const int64_t items = row_blocks * p->bv.nblk;
            SPUQ_REQUIRE(items < 0x7fffffffLL, "apply: too many work items (%lld)", (long long)items);
            char name[64];
            bool done = false;
#define SPUQ_ELLBLK(kw)                                                                                              \
    if (!done && p->ell_K <= kw) {                                                                                   \
        if (p->ell_K == kw)                                                                                          \
            SPUQ_LAUNCH((k_apply_ellblk<kw, true>), dim3((unsigned)items), dim3(threads), 0, stream, p->n_rows,      \
                        p->ell_K, p->ell_cols, p->ell_vals, p->bv, W, ldw, Y, ldy, gate);                            \
        else                                                                                                         \
            SPUQ_LAUNCH((k_apply_ellblk<kw, false>), dim3((unsigned)items), dim3(threads), 0, stream, p->n_rows,     \
                        p->ell_K, p->ell_cols, p->ell_vals, p->bv, W, ldw, Y, ldy, gate);                            \
        snprintf(name, sizeof(name), "k_apply_ellblk<%d>", kw);                                                      \
        done = true;                                                                                                 \
    }
            SPUQ_ELLBLK(3) SPUQ_ELLBLK(5) SPUQ_ELLBLK(7) SPUQ_ELLBLK(9) SPUQ_ELLBLK(12) SPUQ_ELLBLK(16)
#undef SPUQ_ELLBLK
            p->kernel_name = name;
        } else if (p->path == SPUQ_PATH_CSR) {
            SPUQ_LAUNCH(k_apply_csr, grid, dim3(threads), 0, stream, p->n_rows, p->L, p->rowptr, p->colidx,
                        p->vals_ptrs_dev, tv, W, ldw, Y, ldy, cpc, gate);
            p->kernel_name = "k_apply_csr";
        } else if (p->ell_K <= 8) {
            SPUQ_LAUNCH((k_apply_ell<8>), grid, dim3(threads), 0, stream, p->n_rows, p->L, p->ell_K, p->ell_cols,
                        p->ell_vals, p->slot_dev, tv, W, ldw, Y, ldy, cpc, gate);
            p->kernel_name = "k_apply_ell<8>";
        } else {
            SPUQ_LAUNCH((k_apply_ell<0>), grid, dim3(threads), 0, stream, p->n_rows, p->L, p->ell_K, p->ell_cols,
                        p->ell_vals, p->slot_dev, tv, W, ldw, Y, ldy, cpc, gate);
            p->kernel_name = "k_apply_ell<0>";
        }
    } else {
        const bool aligned = (p->st.s % 2 == 0) && (ldw % 2 == 0) && (ldy % 2 == 0) &&
                             (((uintptr_t)W | (uintptr_t)Y) & 15) == 0;
        SPUQ_REQUIRE(p->variant != 2 || aligned,
                     "apply: tiled stencil kernel needs an even grid width, even leading dimensions and 16-byte aligned slabs");
        SPUQ_REQUIRE(p->variant != 3 || aligned,
                     "apply: TMA stencil kernel needs an even grid width, even leading dimensions and 16-byte aligned slabs");
#ifdef SPUQ_EMU
        const bool can_tma = false;
#else
        const bool can_tma = aligned && ldw % 2 == 0;
#endif
        SPUQ_REQUIRE(p->variant != 4 || (aligned && p->n_slots > 0),
                     "apply: TMEM stencil kernel needs an even grid width, even leading dimensions, 16-byte aligned slabs "
                     "and at least one block matrix");
        // automatic choice: the TMEM-accumulator kernel whenever its alignment conditions hold
        const bool use_tmem = aligned && p->n_slots > 0 && (p->variant == 4 || p->variant == 0);
        if (use_tmem) {
            // rows_per_thread selects the number of consumer groups G (0 = default 2); cols_per_block = K
            const int32_t G = p->rows_per_thread > 0 ? p->rows_per_thread : 2;
            SPUQ_REQUIRE(G >= 1 && G <= 3, "apply: TMEM kernel supports 1 or 2 consumer groups, or 3 = one fused 8-warp group (got %d)", G);
            const int32_t Kmax = p->nine ? (G == 1 ? 2 : 1) : (G == 1 ? 3 : 2);
            const int32_t K = p->cols_per_block > 0 ? p->cols_per_block : Kmax;
            SPUQ_REQUIRE(K >= 1 && K <= Kmax, "apply: TMEM kernel with %d group(s) supports 1..%d register sets (got %d)",
                         G, Kmax, K);
            const int32_t block_cols = (G == 1 ? 512 : 256) / SCHED_ACC_STRIDE - 1;
            if (p->sched_K != K || p->sched_C != block_cols) {
                int rc = build_schedule(p, K, block_cols, G == 1 ? 4 : 2);      // = TmemCfg::W_GROUP
                if (rc) return rc;
            }
#ifdef SPUQ_EMU
            // no TMA / TMEM on the host: interpret the same schedule tables with plain loops
            schedule_run_host(*static_cast<const Schedule*>(p->sched_host), p->st.D, p->st.dx, p->st.dy, p->st.s, p->ny,
                              p->dia_vals, W, ldw, Y, ldy);
            {
                char name[64];
                snprintf(name, sizeof(name), "schedule_run_host<K=%d,G=%d,%s>", K, G, p->nine ? "9pt" : "5pt");
                p->kernel_name = name;
            }
            p->last_launches = 1;
            return SPUQ_OK;
#else
            int rc = SPUQ_E_UNSUPPORTED;
            if (!p->nine) {
                if (G == 1 && K == 1) rc = launch_tmem<1, false, 1>(p, W, ldw, Y, ldy, stream, gate);
                else if (G == 1 && K == 2) rc = launch_tmem<2, false, 1>(p, W, ldw, Y, ldy, stream, gate);
                else if (G == 1 && K == 3) rc = launch_tmem<3, false, 1>(p, W, ldw, Y, ldy, stream, gate);
                else if (G == 2 && K == 1) rc = launch_tmem<1, false, 2>(p, W, ldw, Y, ldy, stream, gate);
                else if (G == 2 && K == 2) rc = launch_tmem<2, false, 2>(p, W, ldw, Y, ldy, stream, gate);
                else if (G == 3 && K == 1) rc = launch_tmem<1, false, 3>(p, W, ldw, Y, ldy, stream, gate);
                else if (G == 3 && K == 2) rc = launch_tmem<2, false, 3>(p, W, ldw, Y, ldy, stream, gate);
            } else {
                if (G == 1 && K == 1) rc = launch_tmem<1, true, 1>(p, W, ldw, Y, ldy, stream, gate);
                else if (G == 1 && K == 2) rc = launch_tmem<2, true, 1>(p, W, ldw, Y, ldy, stream, gate);
                else if (G == 2 && K == 1) rc = launch_tmem<1, true, 2>(p, W, ldw, Y, ldy, stream, gate);
                else if (G == 3 && K == 1) rc = launch_tmem<1, true, 3>(p, W, ldw, Y, ldy, stream, gate);
            }
            if (rc) return rc;
            p->last_launches = 1;      // the kernel (the work counter is reset by a memset node, not a kernel)
            SPUQ_CUDA_TRY(cudaGetLastError());
            return SPUQ_OK;
#endif
        }
        // TMA pipeline on grids at least one box wide/tall, else loads through registers
        const bool use_tma = can_tma && p->variant == 3;
        const bool tiled = !use_tma && (p->variant == 2 || p->variant == 3 || (p->variant == 0 && aligned));
        if (!tiled && !use_tma) {
            const int64_t pair_blocks = ((p->n_rows + 1) / 2 + threads - 1) / threads;
            const int32_t cpc = p->cols_per_block > 0 ? p->cols_per_block : pick_cols_per_cta(p, pair_blocks);
            const dim3 grid((unsigned)pair_blocks, (unsigned)((p->L + cpc - 1) / cpc));
            SPUQ_REQUIRE(grid.y <= 65535u, "apply: too many column chunks (%u); raise cols_per_block", grid.y);
            SPUQ_LAUNCH(k_apply_stencil_simple, grid, dim3(threads), 0, stream, p->L, p->st, p->slot_dev, tv,
                        W, ldw, Y, ldy, cpc, gate);
            p->kernel_name = "k_apply_stencil_simple";
        } else {
            int32_t RY = p->rows_per_thread > 0 ? p->rows_per_thread : 2;
            int32_t C = p->cols_per_block > 0 ? p->cols_per_block : (RY == 4 ? 4 : 8);
            if (p->bv.bcol == nullptr || p->bv.C != C) {
                int rc = build_blocks(p, C);
                if (rc) return rc;
            }
            const bool nine = p->nine;
            bool launched = false;
#ifndef SPUQ_EMU
            if (use_tma) {
                int rc = SPUQ_OK;
#define SPUQ_TMA(ry, wy, c)                                                                       \
    if (!launched && RY == ry && C == c) {                                                        \
        rc = nine ? launch_tma<ry, wy, c, true>(p, W, ldw, Y, ldy, stream, gate)                  \
                  : launch_tma<ry, wy, c, false>(p, W, ldw, Y, ldy, stream, gate);                \
        launched = true;                                                                          \
    }
                SPUQ_TMA(2, 4, 4) SPUQ_TMA(2, 4, 8) SPUQ_TMA(2, 4, 12) SPUQ_TMA(2, 4, 16)
                SPUQ_TMA(4, 2, 4) SPUQ_TMA(1, 8, 8) SPUQ_TMA(1, 8, 16)
#undef SPUQ_TMA
                if (rc) return rc;
                SPUQ_REQUIRE(launched || p->variant != 3,
                             "apply: no TMA kernel instantiated for rows_per_thread=%d cols_per_block=%d", RY, C);
            }
#endif
            if (!launched) {
#define SPUQ_TILED(ry, c)                                                                  \
    if (RY == ry && C == c) {                                                              \
        if (nine) launch_tiled<ry, c, true>(p, W, ldw, Y, ldy, stream, gate);              \
        else launch_tiled<ry, c, false>(p, W, ldw, Y, ldy, stream, gate);                  \
    } else
                SPUQ_TILED(1, 8) SPUQ_TILED(1, 16) SPUQ_TILED(2, 4) SPUQ_TILED(2, 8) SPUQ_TILED(2, 16)
                SPUQ_TILED(4, 4)
                {
                    set_error("apply: no tiled kernel instantiated for rows_per_thread=%d cols_per_block=%d", RY, C);
                    return SPUQ_E_UNSUPPORTED;
                }
#undef SPUQ_TILED
            }
        }
    }
    p->last_launches = 1;
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

}  // namespace spuq

extern "C" int spuq_sg_apply_part(spuq_sg_plan* p, const double* W, int64_t ldw, double* Y, int64_t ldy,
                                  int32_t part, void* stream) {
    SPUQ_REQUIRE(p != nullptr, "apply_part: null plan");
    SPUQ_REQUIRE(part >= 0 && part <= 2, "apply_part: part must be 0 (all), 1 (interior) or 2 (boundary)");
    SPUQ_CUDA_TRY(cudaSetDevice(p->device));
    if (part != 0) {
        const bool aligned = (p->st.s % 2 == 0) && (ldw % 2 == 0) && (ldy % 2 == 0) &&
                             (((uintptr_t)W | (uintptr_t)Y) & 15) == 0;
        const bool tmem = p->path == SPUQ_PATH_STENCIL && aligned && p->n_slots > 0 && (p->variant == 4 || p->variant == 0);
#ifdef SPUQ_EMU
        const bool split = false;
#else
        const bool split = tmem;
#endif
        if (!split) {
            // kernels without a tile-row range: the whole grid in the boundary call, which is the one the
            // caller issues after the halo rows have arrived
            if (part == 1) { p->last_launches = 0; return SPUQ_OK; }
            part = 0;
        }
    }
    p->part = part;
    const int rc = apply_impl(p, W, ldw, Y, ldy, static_cast<cudaStream_t>(stream), abi_gate());
    p->part = 0;
    return rc;
}

extern "C" int spuq_sg_apply(spuq_sg_plan* p, const double* W, int64_t ldw, double* Y, int64_t ldy,
                             void* stream) {
    SPUQ_REQUIRE(p != nullptr, "apply: null plan");
    SPUQ_CUDA_TRY(cudaSetDevice(p->device));
    return apply_impl(p, W, ldw, Y, ldy, static_cast<cudaStream_t>(stream), abi_gate());
}

// ---- peer-memory halo exchange helpers (spuq_b200.sharding.PeerHaloExchange) -----------------------
#ifndef SPUQ_EMU
__global__ void k_flag_wait_geq(const volatile long long* flag, long long value) {
    while (*flag < value) __nanosleep(200);
}
#else
void k_flag_wait_geq(const volatile long long*, long long) {}
#endif

// One launch per exchange.  The kernel first tells both neighbours that THIS rank's slab is ready to receive
// exchange number `seq` -- everything enqueued before it on the stream is complete: the kernel that read the
// halo rows of the previous exchange AND whatever wrote the slab since (a vector update that sweeps over the
// halo rows, the fill of a new buffer) -- then waits for the same word from them, stores its first / last owned
// grid row of every column straight into the neighbours' halo rows (coalesced 16-byte stores over NVLink) and
// -- last CTA to finish, after a system-wide fence -- publishes the sequence number in their flag words.
//   my_flags: [0] top halo written, [1] bottom halo written, [2] rank above ready to receive, [3] rank below ready
#ifndef SPUQ_EMU
__global__ void __launch_bounds__(256)
k_halo_push(const double* __restrict__ W, int32_t L, int32_t nyl, int32_t nx,
            double* __restrict__ up_slab, int32_t up_nyl, double* __restrict__ dn_slab, int32_t dn_nyl,
            const volatile long long* my_flags, volatile long long* up_flags, volatile long long* dn_flags,
            long long seq, unsigned int* ticket) {
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (up_flags) up_flags[3] = seq;               // seen from above I am the rank below (every CTA writes the same value)
        if (dn_flags) dn_flags[2] = seq;
        if (up_slab) while (my_flags[2] < seq) __nanosleep(100);
        if (dn_slab) while (my_flags[3] < seq) __nanosleep(100);
    }
    __syncthreads();
    const int64_t row2 = nx / 2;                       // double2 per grid row (nx even)
    const int64_t total = (int64_t)L * row2;
    const double2* src = reinterpret_cast<const double2*>(W);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t col = t / row2, x2 = t - col * row2;
        if (up_slab)      // my first owned row (local row 1) -> bottom halo row of the rank above
            reinterpret_cast<double2*>(up_slab)[(col * up_nyl + (up_nyl - 1)) * row2 + x2] = src[(col * nyl + 1) * row2 + x2];
        if (dn_slab)      // my last owned row (local row nyl - 2) -> top halo row of the rank below
            reinterpret_cast<double2*>(dn_slab)[(col * dn_nyl) * row2 + x2] = src[(col * nyl + (nyl - 2)) * row2 + x2];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(ticket, 1u);
        if (done == gridDim.x - 1) {
            __threadfence_system();
            if (up_flags) up_flags[1] = seq;           // I wrote the BOTTOM halo of the rank above
            if (dn_flags) dn_flags[0] = seq;           // ... and the TOP halo of the rank below
            *ticket = 0u;
        }
    }
}
__global__ void k_halo_wait(const volatile long long* my_flags, int has_up, int has_dn, long long seq) {
    if (has_up) while (my_flags[0] < seq) __nanosleep(100);
    if (has_dn) while (my_flags[1] < seq) __nanosleep(100);
}
#endif

extern "C" int spuq_sg_halo_push(const double* W, int32_t L, int32_t nyl, int32_t nx, double* up_slab, int32_t up_nyl,
                                 double* dn_slab, int32_t dn_nyl, const int64_t* my_flags, int64_t* up_flags,
                                 int64_t* dn_flags, int64_t seq, void* ticket_dev, void* stream) {
    SPUQ_REQUIRE(W && my_flags && ticket_dev && L >= 0 && nyl >= 3 && nx >= 2 && nx % 2 == 0, "halo_push: bad argument");
#ifdef SPUQ_EMU
    set_error("halo_push: peer-memory exchange needs CUDA devices");
    return SPUQ_E_UNSUPPORTED;
#else
    const int64_t total = (int64_t)L * (nx / 2);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, 64));
    k_halo_push<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        W, L, nyl, nx, up_slab, up_nyl, dn_slab, dn_nyl, reinterpret_cast<const volatile long long*>(my_flags),
        reinterpret_cast<volatile long long*>(up_flags), reinterpret_cast<volatile long long*>(dn_flags), (long long)seq,
        static_cast<unsigned int*>(ticket_dev));
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
#endif
}
extern "C" int spuq_sg_halo_wait(const int64_t* my_flags, int has_up, int has_dn, int64_t seq, void* stream) {
    SPUQ_REQUIRE(my_flags != nullptr, "halo_wait: null flags");
#ifndef SPUQ_EMU
    k_halo_wait<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const volatile long long*>(my_flags), has_up, has_dn,
                                                               (long long)seq);
    SPUQ_CUDA_TRY(cudaGetLastError());
#endif
    return SPUQ_OK;
}
extern "C" int spuq_sg_copy2d_async(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width_bytes,
                                    int64_t height, void* stream) {
    SPUQ_REQUIRE(dst && src && width_bytes >= 0 && height >= 0 && dpitch >= width_bytes && spitch >= width_bytes,
                 "copy2d_async: bad argument");
    if (width_bytes == 0 || height == 0) return SPUQ_OK;
#ifdef SPUQ_EMU
    for (int64_t r = 0; r < height; ++r)
        std::memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, (size_t)width_bytes);
#else
    SPUQ_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width_bytes, (size_t)height,
                                    cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
#endif
    return SPUQ_OK;
}

extern "C" int spuq_sg_flag_wait_geq(const int64_t* flag_dev, int64_t value, void* stream) {
    SPUQ_REQUIRE(flag_dev != nullptr, "flag_wait_geq: null flag");
    SPUQ_LAUNCH(k_flag_wait_geq, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream),
                reinterpret_cast<const volatile long long*>(flag_dev), (long long)value);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

// Map a peer's allocation (64-byte cudaIpcMemHandle_t exported by the process that owns it) into THIS device's
// context, so that this device's kernels can store into it (a mapping made with the peer device current, as
// torch's storage sharing does, is reachable by the copy engines only).
// Export: the 64-byte IPC handle of the allocation that contains ptr, and ptr's byte offset inside it.
extern "C" int spuq_sg_ipc_export(const void* ptr, void* handle64_out, int64_t* offset_out) {
    SPUQ_REQUIRE(ptr && handle64_out && offset_out, "ipc_export: null argument");
#ifdef SPUQ_EMU
    set_error("ipc_export: needs CUDA devices");
    return SPUQ_E_UNSUPPORTED;
#else
    typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static RangeFn range_fn = nullptr;
    if (!range_fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            range_fn = reinterpret_cast<RangeFn>(f);
    }
    SPUQ_REQUIRE(range_fn != nullptr, "ipc_export: cuMemGetAddressRange entry point not available");
    CUdeviceptr base = 0;
    size_t size = 0;
    const CUresult r = range_fn(&base, &size, (CUdeviceptr)(uintptr_t)ptr);
    SPUQ_REQUIRE(r == CUDA_SUCCESS, "ipc_export: cuMemGetAddressRange failed with CUresult %d", (int)r);
    cudaIpcMemHandle_t h;
    SPUQ_CUDA_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base)));
    std::memcpy(handle64_out, &h, sizeof(h));
    *offset_out = (int64_t)((uintptr_t)ptr - (uintptr_t)base);
    return SPUQ_OK;
#endif
}

extern "C" int spuq_sg_ipc_open(int device, const void* handle64, void** ptr_out) {
    SPUQ_REQUIRE(handle64 && ptr_out, "ipc_open: null argument");
#ifdef SPUQ_EMU
    set_error("ipc_open: needs CUDA devices");
    return SPUQ_E_UNSUPPORTED;
#else
    SPUQ_CUDA_TRY(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    SPUQ_CUDA_TRY(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return SPUQ_OK;
#endif
}
extern "C" int spuq_sg_ipc_close(int device, void* ptr) {
#ifndef SPUQ_EMU
    if (ptr) { cudaSetDevice(device); cudaIpcCloseMemHandle(ptr); cudaGetLastError(); }
#endif
    return SPUQ_OK;
}

extern "C" int spuq_sg_enable_peer_access(int device, int peer_device) {
#ifndef SPUQ_EMU
    SPUQ_CUDA_TRY(cudaSetDevice(device));
    int can = 0;
    SPUQ_CUDA_TRY(cudaDeviceCanAccessPeer(&can, device, peer_device));
    SPUQ_REQUIRE(can, "enable_peer_access: device %d cannot access device %d", device, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        set_error("cudaDeviceEnablePeerAccess(%d) failed: %s", peer_device, cudaGetErrorString(e));
        return SPUQ_E_CUDA;
    }
    cudaGetLastError();
#endif
    return SPUQ_OK;
}

extern "C" int spuq_sg_set_gate(const double* gate_dev) {
    g_abi_gate = gate_dev;
    return SPUQ_OK;
}
