// pcg.cu -- K2/K3 entry points, preconditioners (K4) and the device-resident PCG loop.
// Reference interfaces replaced: application/egsz/pcg.py:15-53, MultiVector arithmetic
// (application/egsz/multi_vector.py:126-152), PreconditioningOperator.apply
// (application/egsz/multi_operator2.py:151-165).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.h"
#include "kernels_blas.cuh"
#include "precond.h"

using namespace spuq;

namespace spuq {

static inline int blas_grid(int64_t n, int n_sm, int threads) {
    int64_t want = (n + threads - 1) / threads;
    int64_t cap = (int64_t)n_sm * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

static int cur_sm_count() {
    int dev = 0;
    cudaGetDevice(&dev);
    static thread_local int cached_dev = -1, cached = 148;
    if (dev != cached_dev) { cached = sm_count(dev); cached_dev = dev; }
    return cached;
}

static int reduce_grid(int64_t n) {
    int g = blas_grid(n, cur_sm_count(), REDUCE_THREADS * 4);
    int cap = std::min(REDUCE_MAX_BLOCKS, cur_sm_count() * 4);
    return std::max(1, std::min(g, cap));
}

int dot_impl(int nout, int64_t n, const double* x, const double* y, double* out, void* scratch,
             cudaStream_t stream, const double* gate, const DotEpilogue* epi_in) {
    const DotEpilogue epi = epi_in ? *epi_in : DotEpilogue{DOT_EPI_NONE, nullptr, 0, nullptr, 0};
    SPUQ_REQUIRE(n >= 0 && x && y && out && scratch, "dot: null argument");
    double* partials = static_cast<double*>(scratch);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(partials + 2 * REDUCE_MAX_BLOCKS);
    const int grid = reduce_grid(n);
    if (nout == 1)
        SPUQ_LAUNCH((k_dot<1>), dim3(grid), dim3(REDUCE_THREADS), 0, stream, n, x, y, out, partials, ticket, gate, epi);
    else
        SPUQ_LAUNCH((k_dot<2>), dim3(grid), dim3(REDUCE_THREADS), 0, stream, n, x, y, out, partials, ticket, gate, epi);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

}  // namespace spuq

extern "C" int64_t spuq_sg_reduce_scratch_bytes(void) { return REDUCE_SCRATCH_BYTES; }

extern "C" int spuq_sg_dot(int64_t n, const double* x, const double* y, double* out, void* scratch,
                           void* stream) {
    return dot_impl(1, n, x, y, out, scratch, static_cast<cudaStream_t>(stream), abi_gate());
}

extern "C" int spuq_sg_dot2(int64_t n, const double* x, const double* y, double* out,
                            void* scratch, void* stream) {
    return dot_impl(2, n, x, y, out, scratch, static_cast<cudaStream_t>(stream), abi_gate());
}

extern "C" int spuq_sg_axpy(int64_t n, double alpha, const double* num, const double* den,
                            const double* x, double* y, void* stream) {
    SPUQ_REQUIRE(n >= 0 && x && y, "axpy: null argument");
    SPUQ_LAUNCH(k_axpy, dim3(blas_grid(n, cur_sm_count(), 256)), dim3(256), 0,
                static_cast<cudaStream_t>(stream), n, alpha, num, den, x, y, abi_gate());
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

extern "C" int spuq_sg_axpy2(int64_t n, const double* num, const double* den, const double* v,
                             double* w, const double* z, double* rho, void* stream) {
    SPUQ_REQUIRE(n >= 0 && num && den && v && w && z && rho, "axpy2: null argument");
    SPUQ_LAUNCH(k_axpy2, dim3(blas_grid(n, cur_sm_count(), 256)), dim3(256), 0,
                static_cast<cudaStream_t>(stream), n, num, den, v, w, z, rho, abi_gate());
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

extern "C" int spuq_sg_xpay(int64_t n, const double* num, const double* den, const double* s,
                            double* v, void* stream) {
    SPUQ_REQUIRE(n >= 0 && num && den && s && v, "xpay: null argument");
    SPUQ_LAUNCH(k_xpay, dim3(blas_grid(n, cur_sm_count(), 256)), dim3(256), 0,
                static_cast<cudaStream_t>(stream), n, num, den, s, v, abi_gate());
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

extern "C" int spuq_sg_combine(int64_t n, int32_t L, int32_t S, const double* W, int64_t ldw,
                               const double* coef, double* out, int64_t ldo, void* stream) {
    SPUQ_REQUIRE(n >= 0 && L >= 0 && S >= 0, "combine: negative size");
    if (n == 0 || S == 0) return SPUQ_OK;
    SPUQ_REQUIRE(W && coef && out && ldw >= n && ldo >= n, "combine: null argument or leading dimension too small");
    auto blocks = [&](int rows) {
        return dim3((unsigned)std::max<int64_t>(1, std::min<int64_t>(((n + rows - 1) / rows + 255) / 256, (int64_t)cur_sm_count() * 8)));
    };
    for (int32_t s0 = 0; s0 < S; s0 += 8) {
        if (S - s0 >= 5) {
            SPUQ_LAUNCH((k_combine<8, 2>), blocks(2), dim3(256), 0, static_cast<cudaStream_t>(stream), n, L, S, s0, W, ldw,
                        coef, out, ldo);
        } else if (S - s0 >= 3) {
            SPUQ_LAUNCH((k_combine<4, 4>), blocks(4), dim3(256), 0, static_cast<cudaStream_t>(stream), n, L, S, s0, W, ldw,
                        coef, out, ldo);
        } else {
            SPUQ_LAUNCH((k_combine<2, 4>), blocks(4), dim3(256), 0, static_cast<cudaStream_t>(stream), n, L, S, s0, W, ldw,
                        coef, out, ldo);
        }
    }
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

extern "C" int spuq_sg_neighbour_combine(int64_t n, int32_t L, const double* W, int64_t ldw, const int32_t* up,
                                         const int32_t* dn, const double* bu, const double* bd, const double* b0,
                                         double* out, int64_t ldo, void* stream) {
    SPUQ_REQUIRE(n >= 0 && L >= 0, "neighbour_combine: negative size");
    if (n == 0 || L == 0) return SPUQ_OK;
    SPUQ_REQUIRE(W && up && dn && bu && bd && b0 && out && ldw >= n && ldo >= n && W != out,
                 "neighbour_combine: null argument, leading dimension too small, or in-place call");
    SPUQ_REQUIRE(L <= 65535, "neighbour_combine: more than 65535 columns");
    const unsigned bx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 64));
    SPUQ_LAUNCH(k_neighbour_combine, dim3(bx, (unsigned)L), dim3(256), 0, static_cast<cudaStream_t>(stream), n, L, W, ldw,
                up, dn, bu, bd, b0, out, ldo);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

extern "C" int spuq_sg_scal(int64_t n, double alpha, double* x, void* stream) {
    SPUQ_REQUIRE(n >= 0 && x, "scal: null argument");
    SPUQ_LAUNCH(k_scal, dim3(blas_grid(n, cur_sm_count(), 256)), dim3(256), 0,
                static_cast<cudaStream_t>(stream), n, alpha, x);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

// ---- PCG state kernels -----------------------------------------------------------------------------
extern "C" int spuq_sg_pcg_check_zeta(double* st, int32_t pass, void* stream) {
    SPUQ_REQUIRE(st && pass >= 1, "pcg_check_zeta: bad argument");
    SPUQ_LAUNCH(k_pcg_check_zeta, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), st, pass, (double*)nullptr);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}
extern "C" int spuq_sg_pcg_check_alpha(double* st, void* stream) {
    SPUQ_REQUIRE(st, "pcg_check_alpha: null state");
    SPUQ_LAUNCH(k_pcg_check_alpha, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), st);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}
extern "C" int spuq_sg_pcg_check_zeta_hist(double* st, int32_t pass, double* hist_dev, void* stream) {
    SPUQ_REQUIRE(st && pass >= 1, "pcg_check_zeta_hist: bad argument");
    SPUQ_LAUNCH(k_pcg_check_zeta, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), st, pass, hist_dev);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}
extern "C" int spuq_sg_pcg_accept(double* st, int32_t src, int32_t dst, void* stream) {
    SPUQ_REQUIRE(st && src >= 0 && dst >= 0 && src < SPUQ_PCG_STATE_DOUBLES && dst < SPUQ_PCG_STATE_DOUBLES,
                 "pcg_accept: bad argument");
    SPUQ_LAUNCH(k_pcg_accept, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), st, src, dst);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}
extern "C" int spuq_sg_pcg_rotate(double* st, void* stream) {
    SPUQ_REQUIRE(st, "pcg_rotate: null state");
    SPUQ_LAUNCH(k_pcg_rotate, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), st);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

// ---- preconditioner handles -----------------------------------------------------------------------
extern "C" int spuq_sg_precond_identity(spuq_sg_precond** out, double scale) {
    SPUQ_REQUIRE(out, "precond_identity: null out");
    spuq_sg_precond* p = new spuq_sg_precond();
    p->kind = spuq_sg_precond::IDENTITY; p->scale = scale;
    *out = p;
    return SPUQ_OK;
}

extern "C" int spuq_sg_precond_from_plan(spuq_sg_precond** out, spuq_sg_plan* plan) {
    SPUQ_REQUIRE(out && plan, "precond_from_plan: null argument");
    SPUQ_REQUIRE(plan->n_rows == plan->n_cols, "precond_from_plan: plan must be square");
    spuq_sg_precond* p = new spuq_sg_precond();
    p->kind = spuq_sg_precond::PLAN; p->plan = plan;
    *out = p;
    return SPUQ_OK;
}

extern "C" int spuq_sg_precond_destroy(spuq_sg_precond* p) {
    if (!p) return SPUQ_OK;
    if (p->kind == spuq_sg_precond::MEAN_FD) mean_fd_destroy(p);
    if (p->kind == spuq_sg_precond::SPARSE_LU) lu_plan_destroy(p);
    delete p;
    return SPUQ_OK;
}

extern "C" int spuq_sg_precond_last_launches(const spuq_sg_precond* p) { return p ? p->last_launches : 0; }

extern "C" int64_t spuq_sg_precond_work_bytes(const spuq_sg_precond* p, int64_t N, int32_t L) {
    if (!p) return 0;
    if (p->kind == spuq_sg_precond::MEAN_FD) return mean_fd_work_bytes(p, N, L);
    if (p->kind == spuq_sg_precond::SPARSE_LU) return lu_work_bytes(p, N, L);
    return 0;
}

namespace spuq {
int precond_apply_impl(spuq_sg_precond* p, int64_t N, int32_t L, const double* R, double* S,
                       void* work, cudaStream_t stream, const double* gate, FusedDot* fd) {
    SPUQ_REQUIRE(p && R && S, "precond_apply: null argument");
    const int64_t n = N * (int64_t)L;
    struct Count {                       // kernels of this call, whatever path returns
        spuq_sg_precond* p; int c0;
        ~Count() { p->last_launches = std::max(launch_counter() - c0, p->kind == spuq_sg_precond::PLAN && p->plan ? p->plan->last_launches : 0); }
    } count{p, launch_counter()};
    switch (p->kind) {
    case spuq_sg_precond::IDENTITY: {
        // S = scale * R  (MultiplicationOperator.apply, linalg/operator.py:471-472)
        SPUQ_CUDA_TRY(cudaMemcpyAsync(S, R, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
        if (p->scale != 1.0) {
            SPUQ_LAUNCH(k_scal, dim3(blas_grid(n, cur_sm_count(), 256)), dim3(256), 0, stream, n, p->scale, S);
            SPUQ_CUDA_TRY(cudaGetLastError());
        }
        return SPUQ_OK;
    }
    case spuq_sg_precond::PLAN:
        SPUQ_REQUIRE(p->plan->n_rows == N && p->plan->L == L, "precond_apply: plan shape mismatch");
        return apply_impl(p->plan, R, N, S, N, stream, gate);
    case spuq_sg_precond::MEAN_FD:
        return mean_fd_apply(p, N, L, R, S, work, stream, gate, fd);
    case spuq_sg_precond::SPARSE_LU:
        return lu_apply(p, N, L, R, S, work, stream, gate);
    }
    return SPUQ_E_INVALID;
}
}  // namespace spuq

extern "C" int spuq_sg_precond_apply(spuq_sg_precond* p, int64_t N, int32_t L, const double* R,
                                     double* S, void* work, void* stream) {
    return precond_apply_impl(p, N, L, R, S, work, static_cast<cudaStream_t>(stream), abi_gate());
}

// ---- the solver -------------------------------------------------------------------------------------
static inline int64_t align256(int64_t b) { return (b + 255) & ~(int64_t)255; }

extern "C" int64_t spuq_sg_pcg_work_bytes(const spuq_sg_plan* A, const spuq_sg_precond* P) {
    if (!A) return 0;
    const int64_t slab = align256(sizeof(double) * A->n_rows * (int64_t)A->L);
    int64_t b = 4 * slab;                                   // rho, s, v, z
    b += align256(sizeof(double) * SPUQ_PCG_STATE_DOUBLES); // state
    b += align256(REDUCE_SCRATCH_BYTES);
    b += align256(sizeof(double) * 4096);                   // zeta history (device)
    b += align256(spuq_sg_precond_work_bytes(P, A->n_rows, A->L));
    return b;
}

extern "C" int spuq_sg_pcg(spuq_sg_plan* A, spuq_sg_precond* P, const double* F, double* Wv,
                           double eps, int32_t maxiter, int32_t poll_every, void* work,
                           double* zeta_out, int32_t* numit_out, int32_t* status_out,
                           double* zeta_hist, void* stream_) {
    SPUQ_REQUIRE(A && P && F && Wv && work, "pcg: null argument");
    SPUQ_REQUIRE(A->n_rows == A->n_cols, "pcg: operator must be square");
    SPUQ_REQUIRE(maxiter >= 1, "pcg: maxiter must be at least 1");
    // the device-side zeta history has room for 4096 passes; longer solves keep iterating and simply
    // stop recording (zeta_hist_host then holds the first 4096 entries)
    const int32_t hist_cap = 4096;
    if (poll_every < 1) poll_every = 8;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SPUQ_CUDA_TRY(cudaSetDevice(A->device));
    const int64_t N = A->n_rows;
    const int32_t L = A->L;
    const int64_t n = N * (int64_t)L;
    const int64_t slab = align256(sizeof(double) * n);
    char* base = static_cast<char*>(work);
    double* rho = reinterpret_cast<double*>(base);
    double* s = reinterpret_cast<double*>(base + slab);
    double* v = reinterpret_cast<double*>(base + 2 * slab);
    double* z = reinterpret_cast<double*>(base + 3 * slab);
    char* q = base + 4 * slab;
    double* st = reinterpret_cast<double*>(q); q += align256(sizeof(double) * SPUQ_PCG_STATE_DOUBLES);
    void* scratch = q; q += align256(REDUCE_SCRATCH_BYTES);
    double* hist = reinterpret_cast<double*>(q); q += align256(sizeof(double) * 4096);
    void* pwork = q;

    const int nsm = A->n_sm;
    int launches = 0;
    int rc;
    double h_state[SPUQ_PCG_STATE_DOUBLES] = {0};
    h_state[4] = (double)SPUQ_PCG_RUNNING;
    h_state[6] = eps * eps;
    SPUQ_CUDA_TRY(cudaMemcpyAsync(st, h_state, sizeof(h_state), cudaMemcpyHostToDevice, stream));
    SPUQ_CUDA_TRY(cudaMemsetAsync(scratch, 0, REDUCE_SCRATCH_BYTES, stream));
    const double* gate = st + 4;

    // rho = f - A w ; s = P rho ; v = s ; zeta = <rho, s>          pcg.py:26-30
    if ((rc = apply_impl(A, Wv, N, z, N, stream, nullptr))) return rc;
    SPUQ_LAUNCH(k_sub, dim3(blas_grid(n, nsm, 256)), dim3(256), 0, stream, n, F, z, rho);
    if ((rc = precond_apply_impl(P, N, L, rho, s, pwork, stream, nullptr))) return rc;
    SPUQ_CUDA_TRY(cudaMemcpyAsync(v, s, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    if ((rc = dot_impl(1, n, rho, s, st + 0, scratch, stream, nullptr))) return rc;
    launches += 4 + A->last_launches + P->last_launches;

    // A pass: apply, <z,v> (+ the singular / not-PD test and alpha), the residual update, P rho, <rho,s> (+ zeta
    // rotation, the factor of the direction update and the convergence test of the next pass), and ONE kernel for
    // the solution update and the direction update (w += alpha v is moved behind the preconditioner, where v is
    // read anyway: 50 GB instead of 60 + 30 on configuration 4).  With the strip preconditioner <rho,s> comes from
    // its final pass: seven launches.  The first convergence test stands alone.
    // (Tried on a B200 and removed: the residual update inside the strip solver's forward pass.  It saves the
    // 30 GB update kernel, but the z loads in that latency-bound kernel cost more than they save: 1.17 s against
    // 1.04 s for the configuration-4 solve.)
    SPUQ_LAUNCH(k_pcg_check_zeta, dim3(1), dim3(1), 0, stream, st, 1, hist);                 // :33-38, pass 1
    launches += 1;
    int32_t pass = 1;
    int status = SPUQ_PCG_RUNNING;
    while (pass < maxiter && status == SPUQ_PCG_RUNNING) {
        const int32_t stop = std::min(maxiter, pass + poll_every);
        for (; pass < stop; ++pass) {
            if ((rc = apply_impl(A, v, N, z, N, stream, gate))) return rc;                   // :40
            const DotEpilogue ea{DOT_EPI_ALPHA, st, pass, nullptr, 0};
            if ((rc = dot_impl(1, n, z, v, st + 8, scratch, stream, gate, &ea))) return rc;  // :41-45
            SPUQ_LAUNCH(k_axpy, dim3(blas_grid(n, nsm, 256)), dim3(256), 0, stream, n, -1.0, st + 0, st + 1,
                        z, rho, gate);                                                       // :48
            // the last pass the loop may run (maxiter - 1) has no successor to test: the reference raises
            // "did not converge" without looking at its zeta
            const DotEpilogue ez{pass + 1 < maxiter ? DOT_EPI_ZETA : DOT_EPI_NONE, st, pass, hist, hist_cap};
            FusedDot fd{&ez, st + 9, false};
            if ((rc = precond_apply_impl(P, N, L, rho, s, pwork, stream, gate, &fd))) return rc;  // :49
            if (!fd.done) {
                if ((rc = dot_impl(1, n, rho, s, st + 9, scratch, stream, gate, &ez))) return rc;  // :50, :33-38
                launches += 1;
            }
            SPUQ_LAUNCH(k_xpay_w, dim3(blas_grid(n, nsm, 256)), dim3(256), 0, stream, n, st, pass, s, v, Wv);  // :47, :51
            launches += 3 + A->last_launches + P->last_launches;
        }
        SPUQ_CUDA_TRY(cudaGetLastError());
        SPUQ_CUDA_TRY(cudaMemcpyAsync(h_state, st, sizeof(h_state), cudaMemcpyDeviceToHost, stream));
        SPUQ_CUDA_TRY(cudaStreamSynchronize(stream));
        status = (int)h_state[4];
    }
    if (status == SPUQ_PCG_RUNNING) status = SPUQ_PCG_NOT_CONVERGED;     // pcg.py:53
    const int32_t numit = (int32_t)h_state[5];
    if (zeta_hist && numit > 0)
        SPUQ_CUDA_TRY(cudaMemcpy(zeta_hist, hist, sizeof(double) * std::min(numit, hist_cap), cudaMemcpyDeviceToHost));
    // the reference formats the singular / not-positive-definite messages with alpha (pcg.py:42-45)
    if (zeta_out) *zeta_out = (status == SPUQ_PCG_SINGULAR || status == SPUQ_PCG_NOT_PD) ? h_state[1] : h_state[0];
    if (numit_out) *numit_out = numit;
    if (status_out) *status_out = status;
    A->last_launches = launches;
    return SPUQ_OK;
}
