// reduce_common.cuh -- what every kernel that finishes a slab reduction shares: the PCG epilogue of a dot
// product and the warp sum (included by kernels_blas.cuh and precond_strip.cuh).
#pragma once
#include "common.h"

namespace spuq {

// What the block that finishes a reduction does with the result besides storing it (PCG, pcg.py:41-51): the
// reference's branch conditions are evaluated right where the scalar is produced, so a pass needs no
// one-thread kernels between its big ones.
//   DOT_EPI_ALPHA  st[1] = <z,v>; singular / not-positive-definite test                       (:41-45)
//   DOT_EPI_ZETA   st[3] = zeta_new / zeta (the factor of the direction update, :51), zeta <- zeta_new, and
//                  the convergence test of the NEXT pass (:33-38) with its history entry
enum { DOT_EPI_NONE = 0, DOT_EPI_ALPHA = 1, DOT_EPI_ZETA = 2 };
struct DotEpilogue {
    int kind;
    double* st;        // PCG state words
    int32_t pass;      // DOT_EPI_ZETA: index of the pass that just ends (the test is that of pass + 1)
    double* hist;      // zeta history (device) or null
    int32_t hist_cap;
};

// alpha * (*num) / (*den) with device scalars (null: factor 1)
__device__ __forceinline__ double ratio_of(double alpha, const double* num, const double* den) {
    double a = alpha;
    if (num) a *= __ldg(num);
    if (den) a /= __ldg(den);
    return a;
}

__device__ __forceinline__ void dot_epilogue(const DotEpilogue& e, double a) {
    if (e.kind == DOT_EPI_ALPHA) {
        e.st[1] = a;
        e.st[7] = e.st[0] / a;         // alpha of this pass, for the solution update that follows the preconditioner
        if (a == 0.0) e.st[4] = (double)SPUQ_PCG_SINGULAR;
        else if (a < 0.0) e.st[4] = (double)SPUQ_PCG_NOT_PD;
    } else if (e.kind == DOT_EPI_ZETA) {
        e.st[2] = a;
        e.st[3] = a / e.st[0];
        e.st[0] = a;
        const int32_t next = e.pass + 1;
        if (e.hist && next <= e.hist_cap) e.hist[next - 1] = a;
        e.st[5] = (double)next;
        if (a < 0.0) e.st[4] = (double)SPUQ_PCG_PRECOND_NOT_PD;
        else if (a <= e.st[6]) e.st[4] = (double)SPUQ_PCG_CONVERGED;
    }
}

#ifndef SPUQ_EMU
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace spuq
