// precond_lu.cu -- K4 for a GENERAL mean block: S[:, mu] = A_0^{-1} R[:, mu] from a sparse LU factorisation.
//
// The reference solves with whatever matrix the assemble callback returns: ScipySolveOperator.apply is
// scipy's sparse direct solve (linalg/scipy_operator.py:45-49, SuperLU with COLAMD), repeated for every
// column and every PCG pass (application/egsz/multi_operator2.py:151-165).  Here the factorisation
// Pr A Pc = L U is computed ONCE on the host (set-up, by the Python layer, with the same SuperLU) and the
// two triangular solves run on the device for all columns of the slab:
//
//   * rows of L and of U are grouped into dependency levels (a row depends on the rows its off-diagonal
//     entries name); rows of one level are independent;
//   * one CTA owns a group of slab columns from the permutation of the right-hand side to the scatter of
//     the solution: gather -> L levels -> U levels -> scatter in ONE launch, the levels separated by
//     block barriers only (no grid-wide synchronisation, no launch per level);
//   * inside a level a thread owns (row, column of the group); the factor entries of a row are read once
//     per column group (L2-resident), the solution vector is updated in place in the workspace.
//
// Exact to rounding like any sparse direct solve; parity against scipy's splu is tested to 1e-12.
#include <algorithm>
#include <vector>

#include "common.h"
#include "precond.h"

namespace spuq {

struct TriFactor {                 // rows in level order; CSR of the strict triangle; reciprocal diagonal
    int32_t nlev = 0;
    int32_t* lev_ptr = nullptr;    // nlev + 1
    int32_t* lev_rows = nullptr;   // n
    int32_t* rowptr = nullptr;     // n + 1
    int32_t* colidx = nullptr;
    double* vals = nullptr;
    double* inv_diag = nullptr;    // n, or null for a unit diagonal
};

struct LuPlan {
    int64_t n = 0;
    TriFactor Lf, Uf;
    int32_t* perm_r = nullptr;     // y[perm_r[i]] = b[i]
    int32_t* perm_c = nullptr;     // x[i] = w[perm_c[i]]
    std::vector<void*> dev;
    int64_t nnz = 0;
};

constexpr int LU_THREADS = 512;

#ifdef SPUQ_EMU
#define LU_T0 0
#define LU_TS 1
#define LU_SYNC()
#else
#define LU_T0 ((int)threadIdx.x)
#define LU_TS ((int)blockDim.x)
#define LU_SYNC() __syncthreads()
#endif

struct TriView {
    int32_t nlev;
    const int32_t* lev_ptr;
    const int32_t* lev_rows;
    const int32_t* rowptr;
    const int32_t* colidx;
    const double* vals;
    const double* inv_diag;
};

__device__ __forceinline__ void lu_sweep(const TriView& f, double* __restrict__ X, int64_t n, int ncols) {
    for (int lev = 0; lev < f.nlev; ++lev) {
        const int r0 = f.lev_ptr[lev], nr = f.lev_ptr[lev + 1] - r0;
        for (int t = LU_T0; t < nr * ncols; t += LU_TS) {
            const int c = t / nr, r = f.lev_rows[r0 + (t - c * nr)];
            double* x = X + (int64_t)c * n;
            double s = x[r];
            for (int k = f.rowptr[r]; k < f.rowptr[r + 1]; ++k) s = fma(-f.vals[k], x[f.colidx[k]], s);
            x[r] = f.inv_diag ? s * f.inv_diag[r] : s;
        }
        LU_SYNC();
    }
}

// One CTA: columns [c0, c0 + G) of the slab.
__global__ void __launch_bounds__(LU_THREADS)
k_lu_solve(TriView Lf, TriView Uf, const int32_t* __restrict__ perm_r, const int32_t* __restrict__ perm_c,
           int64_t n, int32_t L, int32_t G, const double* __restrict__ R, double* __restrict__ S,
           double* __restrict__ work, const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
#ifdef SPUQ_EMU
    if (threadIdx.x != 0) return;
#endif
    const int64_t c0 = (int64_t)blockIdx.x * G;
    if (c0 >= L) return;
    const int ncols = (L - c0 < (int64_t)G) ? (int)(L - c0) : G;
    double* X = work + c0 * n;
    const double* Rg = R + c0 * n;
    for (int64_t t = LU_T0; t < (int64_t)ncols * n; t += LU_TS) {
        const int64_t c = t / n, i = t - c * n;
        X[c * n + perm_r[i]] = Rg[c * n + i];
    }
    LU_SYNC();
    lu_sweep(Lf, X, n, ncols);
    lu_sweep(Uf, X, n, ncols);
    double* Sg = S + c0 * n;
    for (int64_t t = LU_T0; t < (int64_t)ncols * n; t += LU_TS) {
        const int64_t c = t / n, i = t - c * n;
        Sg[c * n + i] = X[c * n + perm_c[i]];
    }
}

template <typename T>
static T* lu_upload(LuPlan* lp, const T* h, size_t count) {
    void* d = nullptr;
    if (cudaMalloc(&d, sizeof(T) * std::max<size_t>(count, 1)) != cudaSuccess) return nullptr;
    lp->dev.push_back(d);
    if (count) cudaMemcpy(d, h, sizeof(T) * count, cudaMemcpyHostToDevice);
    return static_cast<T*>(d);
}

// Level sets of a triangular CSR factor (strict part in rowptr/colidx; `lower`: dependencies have smaller
// row numbers and rows are visited upwards, else downwards).
static bool build_factor(LuPlan* lp, int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                         const double* diag, bool lower, TriFactor* out) {
    std::vector<int32_t> level(n, 0);
    int32_t nlev = 0;
    auto visit = [&](int64_t r) {
        int32_t lv = 0;
        for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k) lv = std::max(lv, level[colidx[k]] + 1);
        level[r] = lv;
        nlev = std::max(nlev, lv + 1);
    };
    if (lower) { for (int64_t r = 0; r < n; ++r) visit(r); }
    else { for (int64_t r = n - 1; r >= 0; --r) visit(r); }
    std::vector<int32_t> ptr(nlev + 1, 0), rows(n);
    for (int64_t r = 0; r < n; ++r) ptr[level[r] + 1]++;
    for (int32_t l = 0; l < nlev; ++l) ptr[l + 1] += ptr[l];
    std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t r = 0; r < n; ++r) rows[fill[level[r]]++] = (int32_t)r;
    out->nlev = nlev;
    out->lev_ptr = lu_upload(lp, ptr.data(), ptr.size());
    out->lev_rows = lu_upload(lp, rows.data(), rows.size());
    out->rowptr = lu_upload(lp, rowptr, (size_t)n + 1);
    out->colidx = lu_upload(lp, colidx, (size_t)rowptr[n]);
    out->vals = lu_upload(lp, vals, (size_t)rowptr[n]);
    bool ok = out->lev_ptr && out->lev_rows && out->rowptr && out->colidx && out->vals;
    if (diag) {
        std::vector<double> inv(n);
        for (int64_t r = 0; r < n; ++r) inv[r] = 1.0 / diag[r];
        out->inv_diag = lu_upload(lp, inv.data(), inv.size());
        ok = ok && out->inv_diag;
    }
    return ok;
}

void lu_plan_destroy(spuq_sg_precond* p) {
    LuPlan* lp = static_cast<LuPlan*>(p->lu);
    if (!lp) return;
    cudaSetDevice(p->device);
    for (void* d : lp->dev) cudaFree(d);
    delete lp;
    p->lu = nullptr;
}

int64_t lu_work_bytes(const spuq_sg_precond*, int64_t N, int32_t L) { return (int64_t)sizeof(double) * N * L; }

int lu_apply(spuq_sg_precond* p, int64_t N, int32_t L, const double* R, double* S, void* work, cudaStream_t stream,
             const double* gate) {
    LuPlan* lp = static_cast<LuPlan*>(p->lu);
    SPUQ_REQUIRE(lp && N == lp->n, "precond_sparse_lu: slab rows %lld != matrix dimension %lld", (long long)N,
                 (long long)(lp ? lp->n : -1));
    SPUQ_REQUIRE(work != nullptr, "precond_sparse_lu: workspace missing");
    if (L == 0) return SPUQ_OK;
    // column group per CTA: enough CTAs for every SM first, then wider groups (the factor entries are
    // read once per group)
    const int nsm = sm_count(p->device);
    int G = (int)std::max<int64_t>(1, std::min<int64_t>(16, L / (2 * (int64_t)nsm)));
    const unsigned grid = (unsigned)((L + G - 1) / G);
    auto view = [](const TriFactor& f) {
        return TriView{f.nlev, f.lev_ptr, f.lev_rows, f.rowptr, f.colidx, f.vals, f.inv_diag};
    };
    SPUQ_LAUNCH(k_lu_solve, dim3(grid), dim3(LU_THREADS), 0, stream, view(lp->Lf), view(lp->Uf), lp->perm_r, lp->perm_c,
                N, L, G, R, S, static_cast<double*>(work), gate);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

}  // namespace spuq

using namespace spuq;

extern "C" int spuq_sg_precond_sparse_lu(spuq_sg_precond** out, int device, int64_t n,
                                         const int32_t* l_rowptr, const int32_t* l_colidx, const double* l_vals,
                                         const int32_t* u_rowptr, const int32_t* u_colidx, const double* u_vals,
                                         const double* u_diag, const int32_t* perm_r, const int32_t* perm_c) {
    SPUQ_REQUIRE(out && n >= 1 && l_rowptr && u_rowptr && u_diag && perm_r && perm_c, "precond_sparse_lu: null argument");
    SPUQ_REQUIRE(n < 0x7fffffffLL, "precond_sparse_lu: dimension too large");
    for (int64_t r = 0; r < n; ++r) {
        SPUQ_REQUIRE(u_diag[r] != 0.0, "precond_sparse_lu: zero pivot in row %lld (matrix singular)", (long long)r);
        for (int32_t k = l_rowptr[r]; k < l_rowptr[r + 1]; ++k)
            SPUQ_REQUIRE(l_colidx[k] >= 0 && l_colidx[k] < r, "precond_sparse_lu: L is not strictly lower triangular");
        for (int32_t k = u_rowptr[r]; k < u_rowptr[r + 1]; ++k)
            SPUQ_REQUIRE(u_colidx[k] > r && u_colidx[k] < n, "precond_sparse_lu: U is not strictly upper triangular");
        SPUQ_REQUIRE(perm_r[r] >= 0 && perm_r[r] < n && perm_c[r] >= 0 && perm_c[r] < n, "precond_sparse_lu: bad permutation");
    }
    SPUQ_CUDA_TRY(cudaSetDevice(device));
    spuq_sg_precond* p = new spuq_sg_precond();
    p->kind = spuq_sg_precond::SPARSE_LU;
    p->device = device;
    LuPlan* lp = new LuPlan();
    lp->n = n;
    lp->nnz = (int64_t)l_rowptr[n] + u_rowptr[n] + n;
    p->lu = lp;
    bool ok = build_factor(lp, n, l_rowptr, l_colidx, l_vals, nullptr, true, &lp->Lf) &&
              build_factor(lp, n, u_rowptr, u_colidx, u_vals, u_diag, false, &lp->Uf);
    lp->perm_r = lu_upload(lp, perm_r, (size_t)n);
    lp->perm_c = lu_upload(lp, perm_c, (size_t)n);
    if (!ok || !lp->perm_r || !lp->perm_c) {
        lu_plan_destroy(p);
        delete p;
        set_error("precond_sparse_lu: out of device memory");
        return SPUQ_E_NOMEM;
    }
    *out = p;
    return SPUQ_OK;
}
