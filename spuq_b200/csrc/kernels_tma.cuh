// kernels_tma.cuh -- K1, TMA-pipelined register-tiled stencil kernel (the production apply kernel).
//
// Same arithmetic and register tiling as kernels_tiled.cuh (A tile of the group's block matrix and
// C x RY x 2 accumulators in registers, 2 x RY points per thread, lanes along x), but operands no
// longer travel through per-thread global loads.  A persistent CTA walks its (spatial tile, column
// block) work items; one producer thread runs AHEAD of four consumer warps and fetches, with TMA
// (cp.async.bulk.tensor, completion on mbarriers),
//   * the W window of every term:  box {68, TILE_Y+2, 1} of the slab viewed as the 3-D tensor
//     (x, y, column), origin (x_tile-2, y_tile-1, nu).  Out-of-range coordinates are zero-filled by
//     the TMA unit, which is exactly the homogeneous-boundary behaviour the stencil needs, so the
//     kernel has no boundary branches.  (The innermost box coordinate must be 16-byte aligned --
//     an odd fp64 start raises "illegal instruction", measured with tools/tma_probe.cu -- hence
//     x_tile-2 and not x_tile-1.)  A thread reads its centre pair with one LDS.128 and the two
//     x-halo values with LDS.64;
//   * the A tile of every group:   box {64, TILE_Y, D} of the diagonal storage viewed as
//     (x, y, slot*D + d).
// Rings of W_STAGES windows and A_STAGES tiles decouple the ~1 us L2/HBM latency from the FP64 pipe:
// ncu of the load-through-registers kernel showed 12 resident warps, 47 % L1 utilisation, 10 % FP64
// pipe and "long scoreboard" as the top stall (profiles/r1_tiled_baseline.md).
#pragma once
#include "common.h"

#ifndef SPUQ_EMU
#include <cuda.h>

namespace spuq {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1,
                                            int32_t c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

template <int RY, int WY, bool NINE>
struct TmaCfg {
    static constexpr int D = NINE ? 9 : 5;
    static constexpr int WARPS_Y = WY;
    static constexpr int CONSUMERS = 32 * WARPS_Y;
    static constexpr int THREADS = CONSUMERS + 32;          // + one producer warp
    static constexpr int TILE_X = 64, TILE_Y = WARPS_Y * RY;
    static constexpr int BOX_W = 68;                        // x_tile-2 .. x_tile+65
    static constexpr int WIN_ROWS = TILE_Y + 2;
    static constexpr int WIN_BYTES = BOX_W * WIN_ROWS * 8;
    static constexpr int WIN_STRIDE = (WIN_BYTES + 127) / 128 * 128;
    static constexpr int A_BYTES = TILE_X * TILE_Y * D * 8;
    static constexpr int W_STAGES = 6, A_STAGES = 2;
    static constexpr int SMEM_BYTES = W_STAGES * WIN_STRIDE + A_STAGES * A_BYTES + 128 + 128;   // rings + barriers + align slack
};

template <int RY, int WY, int C, bool NINE>
__global__ void __launch_bounds__((WY + 1) * 32)
k_apply_tma(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapA,
            BlockView bv, int32_t s, int32_t ny, int32_t ntx, int32_t nty, int64_t n_items,
            double* __restrict__ Y, int64_t ldy, const double* gate) {
    using Cfg = TmaCfg<RY, WY, NINE>;
    constexpr int D = Cfg::D;
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wring = smem;
    uint8_t* aring = smem + Cfg::W_STAGES * Cfg::WIN_STRIDE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(aring + Cfg::A_STAGES * Cfg::A_BYTES);
    // barrier slots: [0,W) fullW, [W,2W) emptyW, [2W,2W+A) fullA, [2W+A, 2W+2A) emptyA
    const uint32_t bar0 = smem_u32(bars);
    auto fullW = [&](int i) { return bar0 + 8u * i; };
    auto emptyW = [&](int i) { return bar0 + 8u * (Cfg::W_STAGES + i); };
    auto fullA = [&](int i) { return bar0 + 8u * (2 * Cfg::W_STAGES + i); };
    auto emptyA = [&](int i) { return bar0 + 8u * (2 * Cfg::W_STAGES + Cfg::A_STAGES + i); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::W_STAGES; ++i) { mbar_init(fullW(i), 1); mbar_init(emptyW(i), Cfg::WARPS_Y); }
        for (int i = 0; i < Cfg::A_STAGES; ++i) { mbar_init(fullA(i), 1); mbar_init(emptyA(i), Cfg::WARPS_Y); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == Cfg::WARPS_Y) {
        // ================================ producer ================================================
        if (lane != 0) return;
        int sw = 0, sa = 0;
        uint32_t pw = 0, pa = 0;                           // phase bits of the rings
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int32_t blk = (int32_t)(item % bv.nblk);
            const int64_t tile = item / bv.nblk;
            const int32_t ty = (int32_t)(tile % nty), tx = (int32_t)(tile / nty);
            const int32_t xt = tx * Cfg::TILE_X, yt = ty * Cfg::TILE_Y;
            const int32_t g0 = __ldg(bv.bgrp + blk), g1 = __ldg(bv.bgrp + blk + 1);
            for (int32_t g = g0; g < g1; ++g) {
                mbar_wait(emptyA(sa), pa ^ 1u);
                mbar_expect_tx(fullA(sa), Cfg::A_BYTES);
                tma_load_3d(smem_u32(aring + sa * Cfg::A_BYTES), &mapA, xt, yt, __ldg(bv.gmat + g) * D, fullA(sa));
                if (++sa == Cfg::A_STAGES) { sa = 0; pa ^= 1u; }
                const int32_t t0 = __ldg(bv.gterm + g), t1 = __ldg(bv.gterm + g + 1);
                for (int32_t t = t0; t < t1; ++t) {
                    mbar_wait(emptyW(sw), pw ^ 1u);
                    mbar_expect_tx(fullW(sw), Cfg::WIN_BYTES);
                    tma_load_3d(smem_u32(wring + sw * Cfg::WIN_STRIDE), &mapW, xt - 2, yt - 1, __ldg(bv.tcol + t), fullW(sw));
                    if (++sw == Cfg::W_STAGES) { sw = 0; pw ^= 1u; }
                }
            }
        }
        return;
    }

    // ==================================== consumers ================================================
    int sw = 0, sa = 0;
    uint32_t pw = 0, pa = 0;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int32_t blk = (int32_t)(item % bv.nblk);
        const int64_t tile = item / bv.nblk;
        const int32_t ty = (int32_t)(tile % nty), tx = (int32_t)(tile / nty);
        const int32_t x0 = tx * Cfg::TILE_X + 2 * lane;
        const int32_t y0 = ty * Cfg::TILE_Y + warp * RY;

        double acc[C][RY][2];
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int r = 0; r < RY; ++r) acc[c][r][0] = acc[c][r][1] = 0.0;

        const int32_t g0 = bv.bgrp[blk], g1 = bv.bgrp[blk + 1];
        for (int32_t g = g0; g < g1; ++g) {
            // ---- A tile: shared ring -> registers, slot released right away ---------------------
            double a[D][RY][2];
            mbar_wait(fullA(sa), pa);
            {
                const double* at = reinterpret_cast<const double*>(aring + sa * Cfg::A_BYTES)
                                   + (warp * RY) * Cfg::TILE_X + 2 * lane;
#pragma unroll
                for (int d = 0; d < D; ++d)
#pragma unroll
                    for (int r = 0; r < RY; ++r) {
                        const double2 v = *reinterpret_cast<const double2*>(at + (d * Cfg::TILE_Y + r) * Cfg::TILE_X);
                        a[d][r][0] = v.x; a[d][r][1] = v.y;
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(emptyA(sa));
            if (++sa == Cfg::A_STAGES) { sa = 0; pa ^= 1u; }

            const int32_t t0 = bv.gterm[g], t1 = bv.gterm[g + 1];
            for (int32_t t = t0; t < t1; ++t) {
                const double coef = bv.tcoef[t];
                const int32_t tc = bv.tacc[t];
                // ---- window rows warp*RY .. warp*RY+RY+1 of the staged box -----------------------
                double win[RY + 2][4];
                mbar_wait(fullW(sw), pw);
                {
                    const double* wt = reinterpret_cast<const double*>(wring + sw * Cfg::WIN_STRIDE)
                                       + (warp * RY) * Cfg::BOX_W + 2 * lane;
#pragma unroll
                    for (int rr = 0; rr < RY + 2; ++rr) {
                        // box element e <-> x = x_tile - 2 + e; this thread's x0 = x_tile + 2*lane
                        const double2 ctr = *reinterpret_cast<const double2*>(wt + rr * Cfg::BOX_W + 2);
                        win[rr][1] = ctr.x; win[rr][2] = ctr.y;
                        if (NINE || (rr >= 1 && rr <= RY)) {
                            win[rr][0] = wt[rr * Cfg::BOX_W + 1];
                            win[rr][3] = wt[rr * Cfg::BOX_W + 4];
                        } else {
                            win[rr][0] = win[rr][3] = 0.0;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(emptyW(sw));
                if (++sw == Cfg::W_STAGES) { sw = 0; pw ^= 1u; }

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
#define SPUQ_ACC_CASE(k)                                                     \
    case k:                                                                  \
        if (k < C) {                                                         \
            _Pragma("unroll") for (int r = 0; r < RY; ++r) {                 \
                acc[k < C ? k : 0][r][0] += coef * sres[r][0];               \
                acc[k < C ? k : 0][r][1] += coef * sres[r][1];               \
            }                                                                \
        }                                                                    \
        break;
                switch (tc) {
                    SPUQ_ACC_CASE(0) SPUQ_ACC_CASE(1) SPUQ_ACC_CASE(2) SPUQ_ACC_CASE(3)
                    SPUQ_ACC_CASE(4) SPUQ_ACC_CASE(5) SPUQ_ACC_CASE(6) SPUQ_ACC_CASE(7)
                    SPUQ_ACC_CASE(8) SPUQ_ACC_CASE(9) SPUQ_ACC_CASE(10) SPUQ_ACC_CASE(11)
                    SPUQ_ACC_CASE(12) SPUQ_ACC_CASE(13) SPUQ_ACC_CASE(14) SPUQ_ACC_CASE(15)
                    default: break;
                }
#undef SPUQ_ACC_CASE
            }
        }

        // ---- write the block's columns once --------------------------------------------------
        if (x0 < s) {
            const int64_t base = (int64_t)y0 * s + x0;
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
    }
}

}  // namespace spuq
#endif  // !SPUQ_EMU
