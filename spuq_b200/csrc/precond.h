// precond.h -- preconditioner handle of libspuq_sg (K4).
#pragma once
#include "common.h"

struct spuq_sg_precond {
    enum Kind { IDENTITY = 0, PLAN = 1, MEAN_FD = 2, SPARSE_LU = 3 } kind = IDENTITY;
    double scale = 1.0;              // IDENTITY
    spuq_sg_plan* plan = nullptr;    // PLAN (borrowed)
    // MEAN_FD: A_0 = tridiag_x(ox, dxx, ox) (+) tridiag_y(oy, dyy, oy) on an nx x ny grid
    int device = 0;
    int32_t nx = 0, ny = 0;
    double d = 0.0, ox = 0.0, oy = 0.0;
    double* q_dev = nullptr;         // nx x nx sine-transform matrix (symmetric, orthogonal)
    double* lam_dev = nullptr;       // nx eigenvalues of the x part
    double* piv_dev = nullptr;       // nx x ny Thomas pivots of (T_y + lam_k I)
    // even nx: the sine transform split by the parity of the mode (two products of half the size)
    int32_t half = 0;                // nx / 2 when the split is used, else 0
    double* qsplit_dev = nullptr;    // four half x half matrices: Qe, Qo (forward), Qe^T, Qo^T (inverse)
    double* piv_split_dev = nullptr; // pivots with the modes ordered even | odd
    // two-level strip partition (precond_strip.cuh): opaque spuq::StripPlan*, null when the shape does
    // not allow it (then the full-length transform above is used)
    void* strip = nullptr;
    // SPARSE_LU: opaque spuq::LuPlan* (precond_lu.cu)
    void* lu = nullptr;
    int last_launches = 0;           // kernels of the last spuq_sg_precond_apply / precond_apply_impl on this handle
};

namespace spuq {
struct StripRowsCtx {            // the handle's grid is one band of a row-sharded grid
    int32_t world, rank;
    const int32_t* bounds;       // world + 1 global row boundaries
};
int mean_fd_destroy(spuq_sg_precond* p);
// strip path (precond_strip.cuh / precond.cu)
int strip_plan_build(spuq_sg_precond* p, const StripRowsCtx* ctx);
void strip_plan_destroy(spuq_sg_precond* p);
// general mean block through a sparse LU factorisation (precond_lu.cu)
void lu_plan_destroy(spuq_sg_precond* p);
int64_t lu_work_bytes(const spuq_sg_precond* p, int64_t N, int32_t L);
int lu_apply(spuq_sg_precond* p, int64_t N, int32_t L, const double* R, double* S, void* work, cudaStream_t stream,
             const double* gate);
int64_t mean_fd_work_bytes(const spuq_sg_precond* p, int64_t N, int32_t L);
struct DotEpilogue;
// A caller that needs <R, S> right after S = P R (pcg.py:49-50) offers the reduction to the preconditioner:
// a path that can produce it while it writes S (the strip solver's final pass) sets `done`; otherwise the
// caller launches its own dot product.
struct FusedDot {
    const DotEpilogue* epi;
    double* out;
    bool done;
};
int mean_fd_apply(spuq_sg_precond* p, int64_t N, int32_t L, const double* R, double* S, void* work,
                  cudaStream_t stream, const double* gate, FusedDot* fd = nullptr);
int precond_apply_impl(spuq_sg_precond* p, int64_t N, int32_t L, const double* R, double* S,
                       void* work, cudaStream_t stream, const double* gate, FusedDot* fd = nullptr);
int dot_impl(int nout, int64_t n, const double* x, const double* y, double* out, void* scratch,
             cudaStream_t stream, const double* gate, const DotEpilogue* epi = nullptr);
}
