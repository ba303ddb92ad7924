// kernels_tmem.cuh -- K1, the production apply kernel: warp-specialised, TMA-fed, accumulators in
// tensor memory.
//
// Why this shape (rates measured with tools/microbench.cu on a B200):
//   * an SM issues 64 DFMA/clk but its shared-memory pipe moves 128 B/clk (16 doubles) and the window
//     of the 5-point stencil costs 3 doubles per grid point, so a kernel that reads one operand per
//     term from shared memory is bound by LDS at < 40 % of the FP64 rate (and this one still is the
//     shared-memory pipe's customer first: see the halo loads below);
//   * tensor memory is a second, private, 256 KB on-chip store with its own datapath (measured
//     385 B/clk/SM load, 650 B/clk/SM store) -- the tcgen05 accumulator file.  This kernel uses it
//     the way a tcgen05 GEMM does, as the accumulator home, but fills it from the FP64 CUDA cores
//     (tcgen05.ld / DFMA / tcgen05.st): 32 or 64 output columns x 4 grid points per thread.  That frees the
//     register file to hold K block-matrix tiles at once, so ONE window read from shared memory
//     serves every (matrix, output column) pair of the block that needs that input column.
//
// A CTA holds G consumer groups of 4 warps (2 x 2 grid points per thread, lanes along x: a 64 x 8 tile
// per group) and one producer warp per group.  A producer pulls work items (tile, column block) from
// a global counter and walks the block's command list (schedule.h); each of its 32 lanes owns one
// command and issues, with cp.async.bulk[.tensor] and mbarriers, whatever ring slot is free:
//     - the W window of every job   (box 68 x 10 of the slab seen as (x, y, column); zero fill
//                                    outside the grid = the homogeneous boundary),
//     - the A tile of every matrix  (box 64 x 8 x D of the diagonal storage),
//     - the 2 KB chunks of the block's record stream,
// into three shared-memory rings per group.  The consumers interpret the stream: a SET header loads
// the A tiles of a register set and gives the lengths of its five runs of regular (padded) jobs --
// window(s) from shared memory into registers, for every matrix of the job the stencil (20 DFMA per
// window) and S read-modify-writes of TMEM accumulators -- and STOREALL writes the finished columns
// of Y.  G = 3 selects ONE fused group of eight warps on a 64 x 16 tile (TmemCfg).
// Control flow is warp-uniform throughout.
//
// Reference semantics: MultiOperator.apply, application/egsz/multi_operator2.py:38-84; only the
// order in which the terms of a column are summed differs (grouped by register set and input).
#pragma once
#include "common.h"
#include "schedule.h"

#ifndef SPUQ_EMU
#include <cuda.h>
#include <type_traits>
#include "kernels_tma.cuh"   // mbarrier / TMA helpers

namespace spuq {

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t (&r)[8], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
// the wait names the registers so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld8(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t (&r)[16], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }


// ---- explicit shared-space loads (32-bit shared addresses; the ring buffers are written by TMA) ----
__device__ __forceinline__ void lds_f64x2(uint32_t addr, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double x;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(addr) : "memory");
    return x;
}
__device__ __forceinline__ int4 lds_i32x4(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// S = A_k * window on the thread's 2 x 2 patch; a[d][r*2+p], w rows y0-1..y0+2 / x0-1..x0+2
template <bool NINE>
__device__ __forceinline__ void stencil1(const double (&a)[NINE ? 9 : 5][4], const double (&w)[4][4], double (&S)[4]) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int q = r * 2 + p;
            double v;
            if (NINE) {
                v = a[0][q] * w[r][p];
                v = fma(a[1][q], w[r][p + 1], v);
                v = fma(a[2][q], w[r][p + 2], v);
                v = fma(a[3][q], w[r + 1][p], v);
                v = fma(a[4][q], w[r + 1][p + 1], v);
                v = fma(a[5][q], w[r + 1][p + 2], v);
                v = fma(a[6][q], w[r + 2][p], v);
                v = fma(a[7][q], w[r + 2][p + 1], v);
                v = fma(a[8][q], w[r + 2][p + 2], v);
            } else {
                v = a[0][q] * w[r][p + 1];
                v = fma(a[1][q], w[r + 1][p], v);
                v = fma(a[2][q], w[r + 1][p + 1], v);
                v = fma(a[3][q], w[r + 1][p + 2], v);
                v = fma(a[4][q], w[r + 2][p + 1], v);
            }
            S[q] = v;
        }
}
// wait for all outstanding tcgen05.ld; names the destination registers so that no use of them can be
// scheduled above the wait
template <int NREG> struct TmemWait;
template <> struct TmemWait<8> {
    __device__ static __forceinline__ void run(uint32_t* r) {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
    }
};
template <> struct TmemWait<16> {
    __device__ static __forceinline__ void run(uint32_t* r) {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                       "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
    }
};
template <> struct TmemWait<24> {
    __device__ static __forceinline__ void run(uint32_t* r) {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                       "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                       "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]) :: "memory");
    }
};
template <> struct TmemWait<32> {
    __device__ static __forceinline__ void run(uint32_t* r) {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                       "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                       "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                       "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]) :: "memory");
    }
};
__device__ __forceinline__ void tmem_ld8p(uint32_t* r, uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8p(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t* r, uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
// 32 TMEM columns of this warp's lanes <- 0
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                 ::"r"(taddr), "r"(z) : "memory");
}

template <int K_, bool NINE_, int G_>
struct TmemCfg {
    static constexpr int K = K_;
    static constexpr bool NINE = NINE_;
    // G consumer groups of 4 warps share the CTA: each owns 512/G TMEM columns (a block of 512/G/8 - 1
    // output columns + the scratch accumulator), its own rings and its own work items.  Two groups put
    // two warps on every SM sub-partition, which hides the barrier / LDS / tcgen05 round trips of one
    // behind the arithmetic of the other.
    // G_ = 3 is the FUSED form of two groups: ONE group of eight warps on a 64 x 16 tile (the two halves
    // of the tile use the two halves of TMEM), one job stream, one producer warp: 10 % fewer halo rows
    // to stage and half the producer work per grid point.
    static constexpr bool FUSED = (G_ == 3);
    static constexpr int G = FUSED ? 1 : G_;                       // consumer groups (own rings, own items)
    static constexpr int D = NINE ? 9 : 5;
    static constexpr int TILE_X = 64, TILE_Y = FUSED ? 16 : 8;
    static constexpr int CONSUMER_WARPS = FUSED ? 8 : 4;           // per group
    // + one producer warp; with two groups the block is padded to three warpgroups so that
    // setmaxnreg can move registers from the producer warpgroup to the consumers (the register file
    // is split per SM sub-partition: 9 warps would cap every thread at 168 registers)
    static constexpr bool WIDE = (G_ != 1);                        // 384-thread block with setmaxnreg
    static constexpr int THREADS = WIDE ? 384 : 32 * (CONSUMER_WARPS + 1);
    static constexpr int CONSUMER_REGS = 224, PRODUCER_REGS = 56;
    static constexpr int TMEM_COLS = WIDE ? 256 : 512;             // per consumer warp
    static constexpr int BLOCK_COLS = TMEM_COLS / SCHED_ACC_STRIDE - 1;
    static constexpr int BOX_W = 68, WIN_ROWS = TILE_Y + 2;
    static constexpr int WIN_BYTES = BOX_W * WIN_ROWS * 8;
    static constexpr int WIN_STRIDE = (WIN_BYTES + 127) / 128 * 128;
    static constexpr int A_BYTES = TILE_X * TILE_Y * D * 8;
    static constexpr int META_BYTES = SCHED_CHUNK_RECS * 16;
    // W ring: stages of W_GROUP windows under ONE full/empty barrier pair (a barrier round trip costs the
    // single consumer warp of a sub-partition 60-140 clk, so it is paid once per group, not per window)
    static constexpr int W_GROUP = WIDE ? 2 : 4, W_STAGES = FUSED ? 7 : (WIDE ? 6 : 4), W_SLOTS = W_GROUP * W_STAGES;
    static constexpr int A_STAGES = WIDE ? (NINE ? 1 : 2) : (NINE ? 2 : 4), M_STAGES = WIDE ? 3 : 4, I_STAGES = 4;
    // per-group shared memory region
    static constexpr int OFF_A = W_SLOTS * WIN_STRIDE;
    static constexpr int OFF_M = OFF_A + A_STAGES * A_BYTES;
    static constexpr int OFF_I = OFF_M + M_STAGES * META_BYTES;
    static constexpr int OFF_BAR = OFF_I + I_STAGES * 16;
    static constexpr int N_BARS = 2 * (W_STAGES + A_STAGES + M_STAGES + I_STAGES);
    static constexpr int GROUP_BYTES = (OFF_BAR + N_BARS * 8 + 127) / 128 * 128;
    static constexpr int OFF_TMEM = G * GROUP_BYTES;
    static constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;       // + alignment slack
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int K, bool NINE, int G_>
__global__ void __launch_bounds__((G_ == 1) ? 160 : 384, 1)
k_apply_tmem(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapA,
             const SchedRec* __restrict__ recs, const int32_t* __restrict__ pcmd,
             const int32_t* __restrict__ blk_pcmd, int32_t nblk, int32_t s, int32_t ny, int32_t nty,
             int32_t ty_off0, int32_t ty_cnt0, int32_t ty_off1,
             int64_t n_items, unsigned long long* __restrict__ work_counter,
             double* __restrict__ Y, int64_t ldy, const double* gate, int32_t debug_flags) {
    using Cfg = TmemCfg<K, NINE, G_>;
    constexpr int D = Cfg::D;
    constexpr int G = Cfg::G;
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_TMEM);
    const uint32_t smem_s = smem_u32(smem);
    // barrier slots of group g: full/empty pairs of its W, A, metadata and item rings
    constexpr int BA = 2 * Cfg::W_STAGES, BM = BA + 2 * Cfg::A_STAGES, BI = BM + 2 * Cfg::M_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int g = 0; g < G; ++g) {
            const uint32_t b = smem_s + g * Cfg::GROUP_BYTES + Cfg::OFF_BAR;
            for (int i = 0; i < Cfg::W_STAGES; ++i) { mbar_init(b + 8u * i, Cfg::W_GROUP); mbar_init(b + 8u * (Cfg::W_STAGES + i), Cfg::CONSUMER_WARPS); }
            for (int i = 0; i < Cfg::A_STAGES; ++i) { mbar_init(b + 8u * (BA + i), 1); mbar_init(b + 8u * (BA + Cfg::A_STAGES + i), Cfg::CONSUMER_WARPS); }
            for (int i = 0; i < Cfg::M_STAGES; ++i) { mbar_init(b + 8u * (BM + i), 1); mbar_init(b + 8u * (BM + Cfg::M_STAGES + i), Cfg::CONSUMER_WARPS); }
            for (int i = 0; i < Cfg::I_STAGES; ++i) { mbar_init(b + 8u * (BI + i), 1); mbar_init(b + 8u * (BI + Cfg::I_STAGES + i), Cfg::CONSUMER_WARPS); }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    if (warp >= Cfg::CONSUMER_WARPS * G) {
        if (Cfg::WIDE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::PRODUCER_REGS));
        if (warp >= Cfg::CONSUMER_WARPS * G + G) return;       // padding warps of the producer warpgroup
        const int my_group = warp - Cfg::CONSUMER_WARPS * G;   // one producer warp per consumer group
        // ================================ producer warp ============================================
        // Every lane owns one command of a batch of 32: it computes its own ring stage from a ballot
        // prefix, waits for that stage to be free and issues its own copy, so the ~500 clk round trip
        // of (barrier probe, expect_tx, TMA issue) is paid 32-fold in parallel, not per command.  Each
        // consumer group has its own producer warp.
        struct PState {
            bool active = false, done = false;
            int32_t xt = 0, yt = 0, c = 0, c1 = 0;
            uint32_t n_w = 0, n_a = 0, n_m = 0;          // fetches issued so far, per ring
            int si = 0;
            uint32_t pi = 0;
        };
        PState st[G];
        const uint32_t lt = (1u << lane) - 1u;
        bool all_done = false;
        while (!all_done) {
            all_done = true;
            bool progress = false;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                if (g != my_group) continue;
                PState& P = st[g];
                if (P.done) continue;
                all_done = false;
                const uint32_t gs = smem_s + g * Cfg::GROUP_BYTES, b = gs + Cfg::OFF_BAR;
                auto fullW = [&](uint32_t i) { return b + 8u * i; };
                auto emptyW = [&](uint32_t i) { return b + 8u * (Cfg::W_STAGES + i); };
                auto fullA = [&](uint32_t i) { return b + 8u * (BA + i); };
                auto emptyA = [&](uint32_t i) { return b + 8u * (BA + Cfg::A_STAGES + i); };
                auto fullM = [&](uint32_t i) { return b + 8u * (BM + i); };
                auto emptyM = [&](uint32_t i) { return b + 8u * (BM + Cfg::M_STAGES + i); };
                if (!P.active) {
                    unsigned long long item = 0;
                    if (lane == 0) item = atomicAdd(work_counter, 1ull);
                    item = __shfl_sync(0xffffffffu, item, 0);
                    const bool valid = item < (unsigned long long)n_items;
                    int32_t blk = 0;
                    P.xt = -1; P.yt = -1;
                    if (valid) {
                        blk = (int32_t)(item % (unsigned long long)nblk);
                        const int64_t tile = (int64_t)(item / (unsigned long long)nblk);
                        // tile rows of this launch: ty_cnt0 rows from ty_off0, the rest from ty_off1 (a launch
                        // over all rows has ty_off0 = 0, ty_cnt0 = nty); spuq_sg_apply_part uses the two ranges
                        // for the tile rows that touch the halo grid rows / for all the others
                        const int32_t tr = (int32_t)(tile % nty);
                        P.yt = (tr < ty_cnt0 ? ty_off0 + tr : ty_off1 + (tr - ty_cnt0)) * Cfg::TILE_Y;
                        P.xt = (int32_t)(tile / nty) * Cfg::TILE_X;
                    }
                    if (lane == 0) {
                        mbar_wait(b + 8u * (BI + Cfg::I_STAGES + P.si), P.pi ^ 1u);
                        *reinterpret_cast<int4*>(smem + g * Cfg::GROUP_BYTES + Cfg::OFF_I + 16 * P.si) = make_int4(P.xt, P.yt, 0, 0);
                        mbar_arrive(b + 8u * (BI + P.si));
                    }
                    if (++P.si == Cfg::I_STAGES) { P.si = 0; P.pi ^= 1u; }
                    if (!valid) { P.done = true; continue; }
                    P.c = __ldg(blk_pcmd + blk); P.c1 = __ldg(blk_pcmd + blk + 1);
                    P.active = true;
                }
                // ---- issue what can be issued NOW for this group, then turn to the next group: the
                // longest prefix (in stream order) of the next 32 commands whose ring slots are free.
                // No lane ever blocks, so a full ring of one group cannot starve the other; only the
                // first lap of each ring is eligible, which also keeps the parity probes unambiguous.
                const int32_t base = P.c;
                const bool have = base + lane < P.c1;
                const int32_t cmd = have ? __ldg(pcmd + base + lane) : 0;
                const int32_t kind = cmd & PCMD_KIND_MASK;
                int32_t val = cmd & PCMD_VAL_MASK;
                // development switch (SPUQ_KDEBUG=1): every window / A tile fetch hits the same column / matrix,
                // i.e. always L2 -- wrong results, but the time shows what the arithmetic alone costs
                if ((debug_flags & 1) && val != PCMD_WINDOW_PAD && kind != PCMD_META) val = 0;
                if ((debug_flags & 2) && kind == PCMD_ATILE) val = 0;                            // 2: A tiles only
                if ((debug_flags & 4) && kind == PCMD_WINDOW && val != PCMD_WINDOW_PAD) val = 0;   // 4: windows only
                const bool is_w = have && kind == PCMD_WINDOW, is_a = have && kind == PCMD_ATILE,
                           is_m = have && kind == PCMD_META;
                const uint32_t mw = __ballot_sync(0xffffffffu, is_w), ma = __ballot_sync(0xffffffffu, is_a),
                               mm = __ballot_sync(0xffffffffu, is_m);
                const uint32_t rw = __popc(mw & lt), ra = __popc(ma & lt), rm = __popc(mm & lt);
                bool ok = false;
                uint32_t slot = 0, sg = 0;
                if (is_w) {
                    const uint32_t n = P.n_w + rw;
                    slot = n % Cfg::W_SLOTS; sg = slot / Cfg::W_GROUP;
                    ok = rw < Cfg::W_SLOTS && mbar_test(emptyW(sg), ((n / Cfg::W_SLOTS) & 1u) ^ 1u);
                } else if (is_a) {
                    const uint32_t n = P.n_a + ra;
                    sg = n % Cfg::A_STAGES;
                    ok = ra < Cfg::A_STAGES && mbar_test(emptyA(sg), ((n / Cfg::A_STAGES) & 1u) ^ 1u);
                } else if (is_m) {
                    const uint32_t n = P.n_m + rm;
                    sg = n % Cfg::M_STAGES;
                    ok = rm < Cfg::M_STAGES && mbar_test(emptyM(sg), ((n / Cfg::M_STAGES) & 1u) ^ 1u);
                }
                const uint32_t blocked = ~__ballot_sync(0xffffffffu, ok);
                const int npre = blocked ? (__ffs(blocked) - 1) : 32;          // commands issued this visit
                if (lane < npre) {
                    if (is_w) {
                        if (val == PCMD_WINDOW_PAD) {
                            mbar_arrive(fullW(sg));                          // pad slot: arrival without bytes
                        } else {
                            mbar_expect_tx(fullW(sg), Cfg::WIN_BYTES);       // one arrival of W_GROUP + the bytes
                            tma_load_3d(gs + slot * Cfg::WIN_STRIDE, &mapW, P.xt - 2, P.yt - 1, val, fullW(sg));
                        }
                    } else if (is_a) {
                        mbar_expect_tx(fullA(sg), Cfg::A_BYTES);
                        tma_load_3d(gs + Cfg::OFF_A + sg * Cfg::A_BYTES, &mapA, P.xt, P.yt, val * D, fullA(sg));
                    } else {
                        mbar_expect_tx(fullM(sg), Cfg::META_BYTES);
                        bulk_load(gs + Cfg::OFF_M + sg * Cfg::META_BYTES, recs + (size_t)val * SCHED_CHUNK_RECS, Cfg::META_BYTES, fullM(sg));
                    }
                }
                __syncwarp();
                const uint32_t below = npre >= 32 ? 0xffffffffu : ((1u << npre) - 1u);
                P.n_w += __popc(mw & below); P.n_a += __popc(ma & below); P.n_m += __popc(mm & below);
                P.c += npre;
                if (P.c >= P.c1) P.active = false;      // (window groups end with the block: pad commands)
                if (npre > 0) progress = true;
            }
            // back-off of the idle producer: 100 ns << (bits 4-6 of SPUQ_KDEBUG); development switch for the polling cost
            if (!progress && !all_done) __nanosleep(100u << ((debug_flags >> 4) & 7));
        }
        return;
    }

    // ==================================== consumer warps ===========================================
    if (Cfg::WIDE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::CONSUMER_REGS));
    // grp: ring set / work items; lw = TMEM lane quarter (warp % 4); half: which 256 TMEM columns and, in
    // the fused form, which eight rows of the 64 x 16 tile
    const int grp = warp / Cfg::CONSUMER_WARPS, lw = warp % 4, half = warp / 4;
    const int row0 = Cfg::FUSED ? 8 * half + 2 * lw : 2 * lw;                          // first tile row of the thread
    uint32_t tmem_warp = *tmem_slot + ((uint32_t)(lw * 32) << 16) + (uint32_t)((Cfg::WIDE ? half : 0) * Cfg::TMEM_COLS);
    asm volatile("" : "+r"(tmem_warp));
    int sa = 0, sm = 0, si = 0;
    uint32_t pa = 0, pm = 0, pi = 0;
    uint32_t gs = smem_s + grp * Cfg::GROUP_BYTES;
    asm volatile("" : "+r"(gs));
    const uint32_t wring_s = gs, aring_s = gs + Cfg::OFF_A, mring_s = gs + Cfg::OFF_M, iring_s = gs + Cfg::OFF_I;
    const uint32_t bar0 = gs + Cfg::OFF_BAR;
    auto fullW = [&](uint32_t i) { return bar0 + 8u * i; };
    auto emptyW = [&](uint32_t i) { return bar0 + 8u * (Cfg::W_STAGES + i); };
    auto fullA = [&](uint32_t i) { return bar0 + 8u * (BA + i); };
    auto emptyA = [&](uint32_t i) { return bar0 + 8u * (BA + Cfg::A_STAGES + i); };
    auto fullM = [&](uint32_t i) { return bar0 + 8u * (BM + i); };
    auto emptyM = [&](uint32_t i) { return bar0 + 8u * (BM + Cfg::M_STAGES + i); };
    auto fullI = [&](uint32_t i) { return bar0 + 8u * (BI + i); };
    auto emptyI = [&](uint32_t i) { return bar0 + 8u * (BI + Cfg::I_STAGES + i); };
    // this thread's corner inside a staged window / A tile (bytes)
    uint32_t win_off = (uint32_t)((row0 * Cfg::BOX_W + 2 * lane) * 8);
    uint32_t a_off = (uint32_t)((row0 * Cfg::TILE_X + 2 * lane) * 8);
    // loop invariants the compiler would otherwise re-derive from special registers in every job
    // (it rematerialises under register pressure): make them opaque so they stay in registers
    uint32_t lane0 = (lane == 0) ? 1u : 0u;                 // this lane signals the ring barriers
    // x-halo loads (8 bytes out of every 16-byte lane slot) would hit each bank twice per half warp;
    // lanes 8-15 and 24-31 therefore fetch their RIGHT halo in the instruction in which the others
    // fetch their LEFT one (the two hit complementary banks) and swap the results afterwards:
    // 2 instead of 4 wavefronts per halo load, the largest single item of the shared-memory budget
    uint32_t halo_x = (lane & 8) ? 24u : 0u;
    asm volatile("" : "+r"(win_off), "+r"(a_off), "+r"(lane0), "+r"(halo_x));
    auto load_halos = [&](const uint32_t row_addr, double& left, double& right) {
        const double a = lds_f64(row_addr + 8 + halo_x);          // element 1 (or 4)
        const double b = lds_f64(row_addr + 32 - halo_x);         // element 4 (or 1)
        left = halo_x ? b : a;
        right = halo_x ? a : b;
    };

    double A[K][D][4];                    // [register set][diagonal][point r*2+p]
    double w0[4][4], w1[4][4];            // two window buffers: jobs alternate between them
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int d = 0; d < D; ++d) A[k][d][0] = A[k][d][1] = A[k][d][2] = A[k][d][3] = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) w0[r][0] = w0[r][1] = w0[r][2] = w0[r][3] = w1[r][0] = w1[r][1] = w1[r][2] = w1[r][3] = 0.0;

    // position in the W ring: slot of the next window and the parity of its lap (kept incrementally:
    // the ring length is not a power of two)
    uint32_t wslot = 0, wpar = 0;
    auto w_advance = [&](const uint32_t by) {
        wslot += by;
        if (wslot >= Cfg::W_SLOTS) { wslot -= Cfg::W_SLOTS; wpar ^= 1u; }
    };
    // Releasing a window group: the warp is converged here (the barrier wait above reconverges it) and
    // shared-memory reads and mbarrier arrives of a warp go through one in-order pipe, so lane 0's
    // arrive cannot overtake the other lanes' loads; a full __syncwarp() costs a divergence-check
    // branch per window, which a lone warp per scheduler cannot hide.  Compiler ordering only; the assumption is
    // exercised by tests/test_gpu_parity.py::test_tmem_window_release_stress (600 applications, bit-identical).
    auto window_reads_issued = [] { asm volatile("" ::: "memory"); };
    // next window of the W ring -> buf; a group of W_GROUP windows shares one barrier pair
    auto load_window = [&](double (&buf)[4][4]) {
        const uint32_t st = wslot / Cfg::W_GROUP, sub = wslot % Cfg::W_GROUP;
        if (sub == 0) mbar_wait(fullW(st), wpar);
        const uint32_t base = wring_s + wslot * Cfg::WIN_STRIDE + win_off;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            // box element e <-> x = x_tile - 2 + e; this thread's x0 = x_tile + 2*lane
            lds_f64x2(base + (rr * Cfg::BOX_W + 2) * 8, buf[rr][1], buf[rr][2]);
            if (NINE || rr == 1 || rr == 2) load_halos(base + rr * Cfg::BOX_W * 8, buf[rr][0], buf[rr][3]);
        }
        if (sub == Cfg::W_GROUP - 1) {
            window_reads_issued();
            if (lane0) mbar_arrive(emptyW(st));
        }
        w_advance(1);
    };
    // both windows of a pair when a window group holds exactly two windows and the pair starts on a
    // group boundary (the host emits the pair runs of an A-set first): one barrier wait, sixteen
    // loads, one release -- no per-window bookkeeping
    auto load_window_pair = [&]() {
        const uint32_t st = wslot / Cfg::W_GROUP;
        mbar_wait(fullW(st), wpar);
        const uint32_t base = wring_s + wslot * Cfg::WIN_STRIDE + win_off;
        // all of w0 first: its stencil can start while the loads of w1 are still draining
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            double (&buf)[4][4] = *(q == 0 ? &w0 : &w1);
            const uint32_t b = base + q * Cfg::WIN_STRIDE;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                lds_f64x2(b + (rr * Cfg::BOX_W + 2) * 8, buf[rr][1], buf[rr][2]);
                if (NINE || rr == 1 || rr == 2) load_halos(b + rr * Cfg::BOX_W * 8, buf[rr][0], buf[rr][3]);
            }
        }
        window_reads_issued();
        if (lane0) mbar_arrive(emptyW(st));
        w_advance(2);
    };
    // the rest of a partly used window group is padding (groups end at A-set and block boundaries)
    auto skip_window_group = [&]() {
        const uint32_t sub = wslot % Cfg::W_GROUP;
        if (sub != 0) {
            __syncwarp();
            if (lane0) mbar_arrive(emptyW(wslot / Cfg::W_GROUP));
            w_advance(Cfg::W_GROUP - sub);
        }
    };
    uint32_t rptr = 0;                    // shared address of the next unit of the record stream
    auto next_chunk = [&]() {
        __syncwarp();
        if (lane0) mbar_arrive(emptyM(sm));
        if (++sm == Cfg::M_STAGES) { sm = 0; pm ^= 1u; }
        mbar_wait(fullM(sm), pm);
        rptr = mring_s + (uint32_t)sm * Cfg::META_BYTES;
    };
    int cur = 0;                          // buffer of the next job: 0 = w0, 1 = w1

    // one RUN: n jobs with S accumulation slots per matrix; unit 0 of a job names its matrices (kmask)
    auto run_jobs = [&](auto s_tag, const int n) {
        constexpr int S = decltype(s_tag)::value;
        constexpr int CU = (S + 1) / 2, JU = 1 + K * CU;
        // one job: window in wc (already loaded); the window after it is fetched into wp
        auto job = [&](double (&wc)[4][4], double (&wp)[4][4], const bool prefetch) {
            int4 dc = lds_i32x4(rptr);
            while ((dc.x & REC_TYPE_MASK) == REC_NEXT_CHUNK) { next_chunk(); dc = lds_i32x4(rptr); }
            const uint32_t kmask = (uint32_t)(dc.x >> REC_JOB_KMASK_SHIFT) & 0xFu;
            double cf[K][2 * CU];
#pragma unroll
            for (int kk = 0; kk < K; ++kk)
#pragma unroll
                for (int u = 0; u < CU; ++u) lds_f64x2(rptr + 16 * (1 + kk * CU + u), cf[kk][2 * u], cf[kk][2 * u + 1]);
            rptr += 16 * JU;
            if (prefetch) load_window(wp);
#pragma unroll
            for (int kk = 0; kk < K; ++kk) {
                if (kmask >> kk & 1u) {
                    const uint32_t pack = (uint32_t)(kk == 0 ? dc.y : (kk == 1 ? dc.z : dc.w));
                    uint32_t ar[S * 8];
                    uint32_t ta[S];
#pragma unroll
                    for (int sl = 0; sl < S; ++sl) {
                        ta[sl] = tmem_warp + ((pack >> (9 * sl)) & 511u);
                        tmem_ld8p(ar + 8 * sl, ta[sl]);
                    }
                    double Sv[4];
                    stencil1<NINE>(A[kk], wc, Sv);
                    TmemWait<S * 8>::run(ar);
#pragma unroll
                    for (int sl = 0; sl < S; ++sl) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double v = fma(cf[kk][sl], Sv[q], __hiloint2double((int)ar[8 * sl + 2 * q + 1], (int)ar[8 * sl + 2 * q]));
                            ar[8 * sl + 2 * q] = (uint32_t)__double2loint(v); ar[8 * sl + 2 * q + 1] = (uint32_t)__double2hiint(v);
                        }
                        tmem_st8p(ta[sl], ar + 8 * sl);
                    }
                }
            }
        };
        if (n <= 0) return;
        if (cur == 0) load_window(w0); else load_window(w1);
        int j = 0;
        if (cur == 1) { job(w1, w0, n > 1); cur = 0; j = 1; }
        for (; j + 1 < n; j += 2) {
            job(w0, w1, true);
            job(w1, w0, j + 2 < n);
        }
        if (j < n) { job(w0, w1, false); cur = 1; }
    };

    // one RUN2: n pairs of jobs (same matrices, disjoint output columns, S <= 2 slots); both windows of
    // a pair go through the stencil in one basic block
    auto run_pairs = [&](auto s_tag, const int n) {
        constexpr int S = decltype(s_tag)::value;
        constexpr int CU = (S + 1) / 2, JU = 1 + K * CU;
        if (n <= 0) return;
        for (int j = 0; j < n; ++j) {
            // unit 0 of both jobs (not prefetched in the previous pass: the copies that keeps alive
            // cost more than the load latency, which the barrier wait below covers anyway)
            int4 ca = lds_i32x4(rptr), cb = lds_i32x4(rptr + 16);
            // (a loop, not an if: the compiler moves loop bodies out of line, so the common path falls through)
            while ((ca.x & REC_TYPE_MASK) == REC_NEXT_CHUNK) { next_chunk(); ca = lds_i32x4(rptr); cb = lds_i32x4(rptr + 16); }
            // (both jobs carry the same mask; reading it through both words lets unit 0 of job b come in one 16-byte load)
            const uint32_t kmask = (uint32_t)((ca.x & cb.x) >> REC_JOB_KMASK_SHIFT) & 0xFu;
            if (Cfg::W_GROUP == 2) {
                load_window_pair();
            } else {
                load_window(w0);
                load_window(w1);
            }
            const uint32_t cbase = rptr + 32;
            rptr += 16 * 2 * JU;
#pragma unroll
            for (int kk = 0; kk < K; ++kk) {
                if (kmask >> kk & 1u) {
                    double cf[2][2 * CU];
#pragma unroll
                    for (int u = 0; u < CU; ++u) {
                        lds_f64x2(cbase + 16 * (kk * CU + u), cf[0][2 * u], cf[0][2 * u + 1]);
                        lds_f64x2(cbase + 16 * (K * CU + kk * CU + u), cf[1][2 * u], cf[1][2 * u + 1]);
                    }
                    const uint32_t pa2 = (uint32_t)(kk == 0 ? ca.y : (kk == 1 ? ca.z : ca.w));
                    const uint32_t pb2 = (uint32_t)(kk == 0 ? cb.y : (kk == 1 ? cb.z : cb.w));
                    uint32_t ar[2 * S * 8];
                    uint32_t ta[2 * S];
#pragma unroll
                    for (int sl = 0; sl < S; ++sl) {
                        ta[sl] = tmem_warp + ((pa2 >> (9 * sl)) & 511u);
                        ta[S + sl] = tmem_warp + ((pb2 >> (9 * sl)) & 511u);
                    }
#pragma unroll
                    for (int sl = 0; sl < 2 * S; ++sl) tmem_ld8p(ar + 8 * sl, ta[sl]);
                    double Sv[2][4];
                    stencil1<NINE>(A[kk], w0, Sv[0]);
                    stencil1<NINE>(A[kk], w1, Sv[1]);
                    TmemWait<2 * S * 8>::run(ar);
#pragma unroll
                    for (int sl = 0; sl < 2 * S; ++sl) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double v = fma(cf[sl / S][sl % S], Sv[sl / S][q],
                                                 __hiloint2double((int)ar[8 * sl + 2 * q + 1], (int)ar[8 * sl + 2 * q]));
                            ar[8 * sl + 2 * q] = (uint32_t)__double2loint(v); ar[8 * sl + 2 * q + 1] = (uint32_t)__double2hiint(v);
                        }
                        tmem_st8p(ta[sl], ar + 8 * sl);
                    }
                }
            }
        }
    };

    for (;;) {
        // ---- next work item -----------------------------------------------------------------------
        mbar_wait(fullI(si), pi);
        const int4 it = lds_i32x4(iring_s + 16 * si);
        __syncwarp();
        if (lane0) mbar_arrive(emptyI(si));
        if (++si == Cfg::I_STAGES) { si = 0; pi ^= 1u; }
        if (it.x < 0) break;
        const int32_t x0 = it.x + 2 * lane, y0 = it.y + row0;

        // accumulators <- 0 (tcgen05 operations of a warp complete in order)
#pragma unroll
        for (int c = 0; c < Cfg::TMEM_COLS; c += 32) tmem_zero32(tmem_warp + c);

        mbar_wait(fullM(sm), pm);
        rptr = mring_s + (uint32_t)sm * Cfg::META_BYTES;
        for (;;) {
            const int4 u = lds_i32x4(rptr);
            const int type = u.x & REC_TYPE_MASK;
            if (type == REC_NEXT_CHUNK) { next_chunk(); continue; }
            rptr += 16;
            if (type == REC_ASET) {
                skip_window_group();
                const int cnt = (u.x >> REC_CNT_SHIFT) & 0xF;
                const int4 n4 = lds_i32x4(rptr);
                rptr += 16;
#pragma unroll
                for (int kk = 0; kk < K; ++kk) {
                    if (kk < cnt) {
                        mbar_wait(fullA(sa), pa);
                        const uint32_t base = aring_s + (uint32_t)sa * Cfg::A_BYTES + a_off;
#pragma unroll
                        for (int d = 0; d < D; ++d)
#pragma unroll
                            for (int r = 0; r < 2; ++r)
                                lds_f64x2(base + ((d * Cfg::TILE_Y + r) * Cfg::TILE_X) * 8, A[kk][d][2 * r], A[kk][d][2 * r + 1]);
                        __syncwarp();
                        if (lane0) mbar_arrive(emptyA(sa));
                        if (++sa == Cfg::A_STAGES) { sa = 0; pa ^= 1u; }
                    }
                }
                // the five runs of the set; their lengths sit in the header (u.y and the unit after it)
                run_pairs(std::integral_constant<int, 1>{}, n4.x);
                run_pairs(std::integral_constant<int, 2>{}, n4.y);
                run_jobs(std::integral_constant<int, 1>{}, n4.z);
                run_jobs(std::integral_constant<int, 2>{}, n4.w);
                run_jobs(std::integral_constant<int, 3>{}, u.y);
            } else if (type == REC_STOREALL) {
                // all accumulators of the block -> Y, four at a time (32 TMEM columns per load)
                tmem_wait_st();
                const bool xin = x0 < s;
                for (int q = 0; q < u.y; ++q) {
                    const int4 cols4 = lds_i32x4(rptr);
                    rptr += 16;
                    uint32_t r32[32];
                    tmem_ld32(r32, tmem_warp + (uint32_t)(q * 4 * SCHED_ACC_STRIDE));
                    TmemWait<32>::run(r32);
                    const int32_t cols[4] = {cols4.x, cols4.y, cols4.z, cols4.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (cols[t] >= 0 && xin) {
                            double* yp = Y + (int64_t)cols[t] * ldy + (int64_t)y0 * s + x0;
#pragma unroll
                            for (int r = 0; r < 2; ++r)
                                if (y0 + r < ny)
                                    // streaming store: Y is written once and must not evict the
                                    // windows that other blocks of the tile still re-read from L2
                                    __stcs(reinterpret_cast<double2*>(yp + (int64_t)r * s),
                                           make_double2(__hiloint2double((int)r32[8 * t + 4 * r + 1], (int)r32[8 * t + 4 * r]),
                                                        __hiloint2double((int)r32[8 * t + 4 * r + 3], (int)r32[8 * t + 4 * r + 2])));
                        }
                    }
                }
            } else {
                break;      // REC_END_ITEM
            }
        }
        // END_ITEM: the rest of the chunk is padding, and so is the rest of a partly used window group
        skip_window_group();
        __syncwarp();
        if (lane0) mbar_arrive(emptyM(sm));
        if (++sm == Cfg::M_STAGES) { sm = 0; pm ^= 1u; }
    }

    // all TMEM traffic of the CTA is complete (every load was waited for): release the columns
    asm volatile("tcgen05.fence::before_thread_sync;");
    asm volatile("bar.sync 1, %0;" ::"n"(32 * Cfg::CONSUMER_WARPS * G));
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(512));
}

}  // namespace spuq
#endif  // !SPUQ_EMU
