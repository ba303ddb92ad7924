// kernels_apply.cuh -- K1: the SG operator application kernels (fp64, CUDA cores).
//
// All kernels compute, for every output column mu and row i,
//     Y[i, mu] = sum_{t in terms(mu)} coef[t] * sum_k A_{mat[t]}[i, k] * W[col_k(i), col[t]]
// which is MultiOperator.apply of the reference (application/egsz/multi_operator2.py:38-84) with
// the two loops (over mu, over m) flattened into the per-column term lists built at plan time:
// the mean block with coefficient 1 (:54-56), then for every m the mu term with -beta[0] (:69,
// only when beta[0] != 0), the mu+e_m term with beta[1] (:72-74) and the mu-e_m term with
// beta[-1] (:77-79).  One launch writes every Y element exactly once (no atomics, no Y re-reads).
//
// `gate` (may be null) points at the PCG status word: a launch becomes a no-op once the solve has
// left the RUNNING state, which is how the PCG loop avoids a host round trip per iteration.
#pragma once
#include "common.h"

namespace spuq {

__device__ __forceinline__ bool gate_closed(const double* gate) {
    return gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING;
}

// ---------------------------------------------------------------------------------------------
// General CSR: one thread per row, lanes along rows (the slab is column-major, so a warp reads
// 32 consecutive doubles of one column); blockIdx.y walks chunks of output columns.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_apply_csr(int64_t n_rows, int32_t L, const int32_t* __restrict__ rowptr,
            const int32_t* __restrict__ colidx, const double* const* __restrict__ vals,
            TermView tv, const double* __restrict__ W, int64_t ldw, double* __restrict__ Y,
            int64_t ldy, int32_t cols_per_cta, const double* gate) {
    if (gate_closed(gate)) return;
    int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int32_t k0 = rowptr[row], k1 = rowptr[row + 1];
    int32_t mu0 = (int32_t)blockIdx.y * cols_per_cta;
    int32_t mu1 = mu0 + cols_per_cta < L ? mu0 + cols_per_cta : L;
    for (int32_t mu = mu0; mu < mu1; ++mu) {
        double acc = 0.0;
        for (int32_t t = tv.ptr[mu]; t < tv.ptr[mu + 1]; ++t) {
            const double* a = vals[tv.mat[t]];
            const double* w = W + (int64_t)tv.col[t] * ldw;
            double s = 0.0;
            for (int32_t k = k0; k < k1; ++k) s += a[k] * w[colidx[k]];
            acc += tv.coef[t] * s;
        }
        Y[row + (int64_t)mu * ldy] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// ELL: the plan keeps a padded, row-interleaved copy of the values (slot-major, then k, then row)
// so that the A loads of a warp are contiguous.  Column indices of a row are kept in registers
// across all terms when the padded width is <= KREG.
// ---------------------------------------------------------------------------------------------
template <int KREG>
__global__ void __launch_bounds__(128)
k_apply_ell(int64_t n_rows, int32_t L, int32_t K, const int32_t* __restrict__ ell_cols,
            const double* __restrict__ ell_vals, const int32_t* __restrict__ slot_of_mat,
            TermView tv, const double* __restrict__ W, int64_t ldw, double* __restrict__ Y,
            int64_t ldy, int32_t cols_per_cta, const double* gate) {
    if (gate_closed(gate)) return;
    int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    int32_t cols[KREG > 0 ? KREG : 1];
    if (KREG > 0) {
#pragma unroll
        for (int k = 0; k < KREG; ++k) cols[k] = k < K ? ell_cols[(int64_t)k * n_rows + row] : 0;
    }
    int32_t mu0 = (int32_t)blockIdx.y * cols_per_cta;
    int32_t mu1 = mu0 + cols_per_cta < L ? mu0 + cols_per_cta : L;
    for (int32_t mu = mu0; mu < mu1; ++mu) {
        double acc = 0.0;
        for (int32_t t = tv.ptr[mu]; t < tv.ptr[mu + 1]; ++t) {
            const double* a = ell_vals + ((int64_t)slot_of_mat[tv.mat[t]] * K) * n_rows + row;
            const double* w = W + (int64_t)tv.col[t] * ldw;
            double s = 0.0;
            if (KREG > 0) {
#pragma unroll
                for (int k = 0; k < KREG; ++k)
                    if (k < K) s += a[(int64_t)k * n_rows] * w[cols[k]];
            } else {
                for (int32_t k = 0; k < K; ++k)
                    s += a[(int64_t)k * n_rows] * w[ell_cols[(int64_t)k * n_rows + row]];
            }
            acc += tv.coef[t] * s;
        }
        Y[row + (int64_t)mu * ldy] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// General sparse blocks, column-blocked ("design M" of kernels_tiled.cuh without the grid): the form
// the SG operator takes on unstructured meshes (P1 FEM blocks of fem/fenics/fenics_operator.py:43-50
// behind linalg/scipy_operator.py:29-40).
//   * a CTA owns 128 rows and a block of 16 output columns whose accumulators stay in registers;
//   * the block's terms are grouped by block matrix: the row's KW values of A_m (padded ELL copy,
//     coalesced) are loaded ONCE per group and reused by every term of that matrix, the row's column
//     indices once per CTA;
//   * per term only the gather of the input column is left: KW loads of W through L1 per row.
// Work items are ordered row block major, so the 62 column blocks that read the same rows of W run
// close together in time (L2 reuse of W and of the A values).
// ---------------------------------------------------------------------------------------------
constexpr int ELLBLK_C = 16;

// acc[which] += v with a warp-uniform `which`: a jump into a table of 16 adds (a dynamically indexed
// register array would live in local memory)
__device__ __forceinline__ void ellblk_accumulate(double (&acc)[ELLBLK_C], int32_t which, double v) {
    switch (which) {
        case 0: acc[0] += v; break;   case 1: acc[1] += v; break;   case 2: acc[2] += v; break;   case 3: acc[3] += v; break;
        case 4: acc[4] += v; break;   case 5: acc[5] += v; break;   case 6: acc[6] += v; break;   case 7: acc[7] += v; break;
        case 8: acc[8] += v; break;   case 9: acc[9] += v; break;   case 10: acc[10] += v; break; case 11: acc[11] += v; break;
        case 12: acc[12] += v; break; case 13: acc[13] += v; break; case 14: acc[14] += v; break; case 15: acc[15] += v; break;
        default: break;
    }
}

// one gathered operand: W_col[cols_k] with the address formed by ONE 32 x 32 -> 64 multiply-add (the compiler's
// own code sign-extended the column index and spent a LEA pair per gather: 105 instructions per term)
__device__ __forceinline__ double ellblk_gather(const double* wcol, int32_t col_k) {
#ifdef SPUQ_EMU
    return wcol[col_k];
#else
    double v;
    asm("{\n.reg .u64 a;\nmad.wide.s32 a, %2, 8, %1;\nld.global.nc.f64 %0, [a];\n}\n" : "=d"(v) : "l"(wcol), "r"(col_k));
    return v;
#endif
}

// KW = padded row width held in registers; EXACT: every row has exactly KW stored entries slots (K == KW), so
// no slot needs a guard
template <int KW, bool EXACT>
__global__ void __launch_bounds__(128)
k_apply_ellblk(int64_t n_rows, int32_t K, const int32_t* __restrict__ ell_cols,
               const double* __restrict__ ell_vals, BlockView bv, const double* __restrict__ W, int64_t ldw,
               double* __restrict__ Y, int64_t ldy, const double* gate) {
    if (gate_closed(gate)) return;
    const int64_t item = blockIdx.x;
    const int32_t blk = (int32_t)(item % bv.nblk);
    const int64_t row = (item / bv.nblk) * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    int32_t cols[KW];
    {
        const int32_t* cp = ell_cols + row;
#pragma unroll
        for (int k = 0; k < KW; ++k) { cols[k] = (EXACT || k < K) ? *cp : 0; cp += n_rows; }
    }
    double acc[ELLBLK_C];
#pragma unroll
    for (int c = 0; c < ELLBLK_C; ++c) acc[c] = 0.0;
    const int4* trec = reinterpret_cast<const int4*>(bv.trec);        // {col, acc, coef lo, coef hi} per term
    for (int32_t g = bv.bgrp[blk]; g < bv.bgrp[blk + 1]; ++g) {
        double a[KW];
        {
            const double* ap = ell_vals + ((int64_t)bv.gmat[g] * K) * n_rows + row;
#pragma unroll
            for (int k = 0; k < KW; ++k) { a[k] = (EXACT || k < K) ? *ap : 0.0; ap += n_rows; }
        }
        // two terms per trip: all 2 x KW gathers are issued before the first product, so that two input
        // columns are in flight per warp (the gathers are the only latency on this path)
        const int32_t t1 = bv.gterm[g + 1];
        int32_t t = bv.gterm[g];
        for (; t + 1 < t1; t += 2) {
            const int4 ra = __ldg(trec + t), rb = __ldg(trec + t + 1);
            const double* wa = W + (int64_t)ra.x * ldw;
            const double* wb = W + (int64_t)rb.x * ldw;
            double ga[KW], gb[KW];
#pragma unroll
            for (int k = 0; k < KW; ++k) { ga[k] = ellblk_gather(wa, cols[k]); gb[k] = ellblk_gather(wb, cols[k]); }
            double sa = 0.0, sb = 0.0;
#pragma unroll
            for (int k = 0; k < KW; ++k) { sa = fma(a[k], ga[k], sa); sb = fma(a[k], gb[k], sb); }
            ellblk_accumulate(acc, ra.y, __hiloint2double(ra.w, ra.z) * sa);
            ellblk_accumulate(acc, rb.y, __hiloint2double(rb.w, rb.z) * sb);
        }
        if (t < t1) {
            const int4 ra = __ldg(trec + t);
            const double* wa = W + (int64_t)ra.x * ldw;
            double sa = 0.0;
#pragma unroll
            for (int k = 0; k < KW; ++k) sa = fma(a[k], ellblk_gather(wa, cols[k]), sa);
            ellblk_accumulate(acc, ra.y, __hiloint2double(ra.w, ra.z) * sa);
        }
    }
#pragma unroll
    for (int c = 0; c < ELLBLK_C; ++c) {
        const int32_t col = bv.bcol[(int64_t)blk * ELLBLK_C + c];
        if (col >= 0) Y[row + (int64_t)col * ldy] = acc[c];
    }
}

// ---------------------------------------------------------------------------------------------
// Stencil, simple form: diagonal storage, one thread per PAIR of adjacent rows (128-bit loads of
// the A diagonals and of the aligned W diagonals), one output column at a time.  This is the
// reference point for the register-tiled kernel in kernels_tiled.cuh.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_apply_stencil_simple(int32_t L, StencilView st, const int32_t* __restrict__ slot_of_mat,
                       TermView tv, const double* __restrict__ W, int64_t ldw,
                       double* __restrict__ Y, int64_t ldy, int32_t cols_per_cta,
                       const double* gate) {
    if (gate_closed(gate)) return;
    const int64_t N = st.N;
    int64_t i0 = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i0 >= N) return;
    const bool pair = i0 + 1 < N;
    int32_t mu0 = (int32_t)blockIdx.y * cols_per_cta;
    int32_t mu1 = mu0 + cols_per_cta < L ? mu0 + cols_per_cta : L;
    for (int32_t mu = mu0; mu < mu1; ++mu) {
        double acc0 = 0.0, acc1 = 0.0;
        for (int32_t t = tv.ptr[mu]; t < tv.ptr[mu + 1]; ++t) {
            const double* a = st.vals + ((int64_t)slot_of_mat[tv.mat[t]] * st.D) * N;
            const double* w = W + (int64_t)tv.col[t] * ldw;
            double s0 = 0.0, s1 = 0.0;
            for (int32_t d = 0; d < st.D; ++d) {
                const int64_t j0 = i0 + st.off[d];
                const double a0 = a[(int64_t)d * N + i0];
                const double a1 = pair ? a[(int64_t)d * N + i0 + 1] : 0.0;
                // rows outside [0, N) only ever meet a zero A entry; clamp instead of branching
                const double w0 = (j0 >= 0 && j0 < N) ? w[j0] : 0.0;
                const double w1 = (j0 + 1 >= 0 && j0 + 1 < N) ? w[j0 + 1] : 0.0;
                s0 += a0 * w0;
                s1 += a1 * w1;
            }
            const double c = tv.coef[t];
            acc0 += c * s0;
            acc1 += c * s1;
        }
        Y[i0 + (int64_t)mu * ldy] = acc0;
        if (pair) Y[i0 + 1 + (int64_t)mu * ldy] = acc1;
    }
}

// ---- plan-time layout conversion kernels ---------------------------------------------------------

// ELL columns/values from CSR (one thread per row).  Padding: column 0, value 0.
__global__ void k_csr_to_ell(int64_t n_rows, int32_t K, const int32_t* __restrict__ rowptr,
                             const int32_t* __restrict__ colidx, const double* __restrict__ vals,
                             int32_t* __restrict__ ell_cols, double* __restrict__ ell_vals_slot) {
    int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int32_t k0 = rowptr[row], cnt = rowptr[row + 1] - k0;
    for (int32_t k = 0; k < K; ++k) {
        if (ell_cols) ell_cols[(int64_t)k * n_rows + row] = k < cnt ? colidx[k0 + k] : 0;
        if (ell_vals_slot) ell_vals_slot[(int64_t)k * n_rows + row] = k < cnt ? vals[k0 + k] : 0.0;
    }
}

// Diagonal storage from CSR (one thread per row): dia[d*N + i] = A[i, i+off[d]] or 0.
__global__ void k_csr_to_dia(StencilView st, const int32_t* __restrict__ rowptr,
                             const int32_t* __restrict__ colidx, const double* __restrict__ vals,
                             double* __restrict__ dia_slot) {
    int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= st.N) return;
    for (int32_t d = 0; d < st.D; ++d) dia_slot[(int64_t)d * st.N + row] = 0.0;
    for (int32_t k = rowptr[row]; k < rowptr[row + 1]; ++k) {
        const int64_t o = (int64_t)colidx[k] - row;
        for (int32_t d = 0; d < st.D; ++d)
            if (st.off[d] == o) dia_slot[(int64_t)d * st.N + row] += vals[k];
    }
}

}  // namespace spuq
