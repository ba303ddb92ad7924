// kernels_tiled.cuh -- K1, register-tiled stencil kernel ("design M").
//
// Why this shape: every DFMA of the SG apply consumes one A value and one W
// value.  An SM can issue 64 DFMA/clk but its L1/shared pipe delivers only 16 doubles/clk, so both
// operands must be reused from registers:
//   * A reuse  -- the work of a block of C output columns is grouped by block matrix (m-major):
//                 the D x RY x 2 values of A_m for this thread's points are loaded ONCE per
//                 group and stay in registers for all terms of that m;
//   * W reuse  -- each thread owns a 2 x RY patch of grid points (x fastest, 128-bit loads) and
//                 loads a (RY+2) x 4 window of W per term, used by up to 9 diagonals;
//   * Y        -- C x RY x 2 accumulators live in registers for the whole block and are written
//                 once (the "mu-fused" form: no Y re-reads, no atomics).
// Lanes run along x, so every global access of a warp is a contiguous 512-byte (128-bit/lane) or
// 256-byte segment of one slab column.
//
// Reference semantics: identical to kernels_apply.cuh (multi_operator2.py:38-84); only the
// order in which the terms of a column are summed differs (grouped by m instead of by (m, +/-)).
#pragma once
#include "common.h"

namespace spuq {

// NINE = false: 5-point set {(0,-1), (-1,0), (0,0), (1,0), (0,1)} stored as D = 5 diagonals in
//               that order;  NINE = true: the full 3x3 window, D = 9, row-major (dy, dx).
template <int RY, int C, bool NINE>
struct TiledCfg {
    static constexpr int D = NINE ? 9 : 5;
    static constexpr int WARPS_Y = 4;                 // warps of a CTA are stacked in y
    static constexpr int THREADS = 32 * WARPS_Y;
    static constexpr int TILE_X = 64;                 // 32 lanes x 2 points
    static constexpr int TILE_Y = WARPS_Y * RY;
};

template <int RY, int C, bool NINE>
__global__ void __launch_bounds__(128)
k_apply_tiled(StencilView st, BlockView bv, int32_t ny, int32_t ntx, int32_t nty,
              const double* __restrict__ W, int64_t ldw, double* __restrict__ Y, int64_t ldy,
              const double* gate) {
    using Cfg = TiledCfg<RY, C, NINE>;
    constexpr int D = Cfg::D;
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;

    // work item: (tile, column block); tiles ordered x-strip major, then y, so that the halo rows
    // a tile shares with its y-neighbour are still in L2 when the neighbour runs
    const int64_t item = blockIdx.x;
    const int32_t blk = (int32_t)(item % bv.nblk);
    const int64_t tile = item / bv.nblk;
    const int32_t ty = (int32_t)(tile % nty);
    const int32_t tx = (int32_t)(tile / nty);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t s = st.s;
    const int64_t N = st.N;
    const int32_t x0 = tx * Cfg::TILE_X + 2 * lane;
    const int32_t y0 = ty * Cfg::TILE_Y + warp * RY;
    if (x0 >= s || y0 >= ny) return;
    const bool has_left = x0 > 0, has_right = x0 + 2 < s;

    double acc[C][RY][2];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int r = 0; r < RY; ++r) acc[c][r][0] = acc[c][r][1] = 0.0;

    const int64_t base = (int64_t)y0 * s + x0;     // first point of this thread

    for (int32_t g = bv.bgrp[blk]; g < bv.bgrp[blk + 1]; ++g) {
        // ---- A tile of this group's block matrix: registers for the whole group ---------------
        double a[D][RY][2];
        const double* ap = st.vals + ((int64_t)bv.gmat[g] * D) * N + base;
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int r = 0; r < RY; ++r) {
                if (y0 + r < ny) {
                    const double2 v = *reinterpret_cast<const double2*>(ap + (int64_t)d * N + (int64_t)r * s);
                    a[d][r][0] = v.x; a[d][r][1] = v.y;
                } else {
                    a[d][r][0] = a[d][r][1] = 0.0;
                }
            }

        for (int32_t t = bv.gterm[g]; t < bv.gterm[g + 1]; ++t) {
            const double* w = W + (int64_t)bv.tcol[t] * ldw + base;
            // ---- window: rows y0-1 .. y0+RY, x0-1 .. x0+2 ---------------------------------------
            double win[RY + 2][4];
#pragma unroll
            for (int rr = 0; rr < RY + 2; ++rr) {
                const int32_t y = y0 + rr - 1;
                const bool row_ok = y >= 0 && y < ny;
                const bool need_halo = NINE || (rr >= 1 && rr <= RY);
                const double* wr = w + (int64_t)(rr - 1) * s;
                double2 ctr = make_double2(0.0, 0.0);
                if (row_ok) ctr = *reinterpret_cast<const double2*>(wr);
                win[rr][1] = ctr.x; win[rr][2] = ctr.y;
                win[rr][0] = (need_halo && row_ok && has_left) ? wr[-1] : 0.0;
                win[rr][3] = (need_halo && row_ok && has_right) ? wr[2] : 0.0;
            }
            // ---- stencil on the patch -------------------------------------------------------
            double sres[RY][2];
#pragma unroll
            for (int r = 0; r < RY; ++r)
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    double v;
                    if (NINE) {
                        v = a[0][r][p] * win[r][p];
                        v += a[1][r][p] * win[r][p + 1];
                        v += a[2][r][p] * win[r][p + 2];
                        v += a[3][r][p] * win[r + 1][p];
                        v += a[4][r][p] * win[r + 1][p + 1];
                        v += a[5][r][p] * win[r + 1][p + 2];
                        v += a[6][r][p] * win[r + 2][p];
                        v += a[7][r][p] * win[r + 2][p + 1];
                        v += a[8][r][p] * win[r + 2][p + 2];
                    } else {
                        v = a[0][r][p] * win[r][p + 1];
                        v += a[1][r][p] * win[r + 1][p];
                        v += a[2][r][p] * win[r + 1][p + 1];
                        v += a[3][r][p] * win[r + 1][p + 2];
                        v += a[4][r][p] * win[r + 2][p + 1];
                    }
                    sres[r][p] = v;
                }
            // ---- accumulate into the (warp-uniform) target column of the block ---------------
            const double coef = bv.tcoef[t];
            // tacc is warp-uniform: a jump, not a predicated chain, picks the accumulator
#define SPUQ_ACC_CASE(k)                                                     \
    case k:                                                                  \
        if (k < C) {                                                         \
            _Pragma("unroll") for (int r = 0; r < RY; ++r) {                 \
                acc[k < C ? k : 0][r][0] += coef * sres[r][0];               \
                acc[k < C ? k : 0][r][1] += coef * sres[r][1];               \
            }                                                                \
        }                                                                    \
        break;
            switch (bv.tacc[t]) {
                SPUQ_ACC_CASE(0) SPUQ_ACC_CASE(1) SPUQ_ACC_CASE(2) SPUQ_ACC_CASE(3)
                SPUQ_ACC_CASE(4) SPUQ_ACC_CASE(5) SPUQ_ACC_CASE(6) SPUQ_ACC_CASE(7)
                SPUQ_ACC_CASE(8) SPUQ_ACC_CASE(9) SPUQ_ACC_CASE(10) SPUQ_ACC_CASE(11)
                SPUQ_ACC_CASE(12) SPUQ_ACC_CASE(13) SPUQ_ACC_CASE(14) SPUQ_ACC_CASE(15)
                default: break;
            }
#undef SPUQ_ACC_CASE
        }
    }

    // ---- write the block's columns once -------------------------------------------------------
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int32_t col = bv.bcol[(int64_t)blk * C + c];
        if (col < 0) continue;
        double* yp = Y + (int64_t)col * ldy + base;
#pragma unroll
        for (int r = 0; r < RY; ++r)
            if (y0 + r < ny)
                *reinterpret_cast<double2*>(yp + (int64_t)r * s) = make_double2(acc[c][r][0], acc[c][r][1]);
    }
}

}  // namespace spuq
