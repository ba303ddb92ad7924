// precond_strip.cuh -- K4, the fast form of the exact mean solve  S[:, mu] = A_0^{-1} R[:, mu]  for a
// constant-coefficient 5-point block  A_0 = I_y (x) T_x + T_y (x) I_x  on an nx x ny grid
// (replaces the per-column sparse direct solve of PreconditioningOperator.apply,
// application/egsz/multi_operator2.py:151-165 / linalg/scipy_operator.py:45-49).
//
// Why this shape.  Fast diagonalisation needs a sine transform of length nx (or ny) per grid line.
// For the grids of the configurations (500, 512, 1000: nx + 1 = 3*167, 27*19, 7*11*13) there is no
// FFT to speak of, and a dense transform costs nx/2 multiply-adds per dof and direction: 1.25 TFLOP
// per application on configuration 4 (54 ms in round 1, on cuBLAS).  This file removes the long
// transform from the bulk of the data by an exact two-level partition (block elimination, no
// approximation):
//
//   level 1 (grid rows): q separator rows cut the grid into strips of c <= 32 rows.  A strip of one
//       slab column is c*nx contiguous doubles (<= 110 KB) and is solved entirely in shared memory:
//         - sine transform of length c along y (one thread per grid column, c*c/2 DFMA, the matrix
//           in shared memory, mirror sums/differences halve the work);
//         - per y-mode a Toeplitz tridiagonal system along x, itself cut by x-separators into chunks
//           of ~31 points: chunk-local Thomas sweeps (one thread per (mode, chunk)), a small
//           separator system per mode, correction with the chunk's spike;
//         - transform back.
//       k_strip<false> does this for the right-hand side and keeps only the first and last row of each
//       strip's solution (all the separator equations need), k_strip<true> repeats it with the separator
//       values moved to the right-hand side and writes the solution.
//   level 2 (separator rows, q/ny ~ 3 % of the data): the Schur complement of the separator rows is
//       block tridiagonal with blocks that are functions of T_x, i.e. diagonal in the sine basis of
//       length nx: dense transform of the q rows (the in-house fp64 GEMM of precond.cu), per mode a
//       q x q tridiagonal solve, transform back.
//
// HBM traffic: R is read twice, S written once (24 N L bytes) instead of ten slab passes; flops per
// dof ~ 2.5 c + 30 instead of 2 nx.  No library call anywhere on the path.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.h"
#include "reduce_common.cuh"
#include "precond.h"

namespace spuq {

constexpr int STRIP_THREADS = 256;
constexpr int STRIP_CMAX = 32;        // strip height limit (registers of the y-transform)
constexpr int STRIP_HP = 16;          // ceil(CMAX / 2)
constexpr int STRIP_CX = 32;          // longest x-chunk (registers of the chunk sweeps)
constexpr int STRIP_QXMAX = 64;       // most x-separators

// Device view of the strip plan (passed by value).
struct StripView {
    int32_t nx, ny, P, q;             // P strips, q = P - 1 separator rows
    int32_t sp;                       // row pitch of the strip in shared memory (>= nx)
    int32_t sepv_n;                   // doubles reserved in front of the strip for the x-separator values [c][qx | 1]
    int64_t ld;                       // column stride of R and S in doubles (nx * ny, or more: a band inside a halo-padded slab)
    int32_t PX, qx;                   // x-chunks per row, x-separators (= PX - 1)
    int32_t cxmax;                    // longest chunk
    double ox, oy;
    const int32_t* ystart;            // P: first grid row of strip p   (host-side / assembly kernel use)
    const int32_t* ycls;              // P: height class of strip p
    int32_t n_tall;                   // the first n_tall strips have height ch[0] (the main strips), the rest ch[1]
    // separator values U(j, col, x) at U[j * u_js + col * u_cs + x].  One GPU: j = local separator, u_js = nx,
    // u_cs = q * nx.  Row-sharded grid (every rank holds the separators of ALL ranks, j-major): the separator
    // above the first local strip is number u_above0 (-1: none), the last local row is a separator of its own
    // when has_bottom is set; Gf has gf_ld strip slots per column (one extra for the neighbour's first strip).
    int64_t u_js, u_cs;
    int32_t u_above0, has_bottom, gf_ld, q_loc;
    const int32_t* xstart;            // PX: first x of chunk ch
    const int32_t* xlen;              // PX
    int32_t ch[2];                    // heights of the two classes
    // per class: Q (c x HP, parity folded), x-direction tables per mode j
    const double* ipivx[2];           // [c][cxmax]   reciprocal Thomas pivots of tridiag(ox, dxx + theta_j, ox)
    const double* spike[2];           // [2][c][cxmax] first spike (T_len^{-1} e_0) for the two chunk lengths
    const double* sx_ipiv[2];         // [c][qx]      x-separator system: reciprocal pivots
    const double* sx_off[2];          // [c][qx]      ... off-diagonal between separator i and i+1
    int32_t xlen_cls_len[2];          // chunk lengths of the two chunk classes
};

// The two y-transform matrices travel as a KERNEL PARAMETER (constant bank, 8 KB): every thread of a
// warp reads the same entry at the same time, which the constant cache serves without touching the
// shared-memory pipe (uniform 16-byte shared loads of the matrix were 60 % of that pipe's wavefronts in
// the first version of this kernel, profiles/r2_precond_strip_notes.md).
struct StripQ {
    double q[2][STRIP_CMAX * STRIP_HP];     // [class][j * HP + i] = Q[j][i], i < ceil(c/2), 0 beyond
};

#ifdef SPUQ_EMU
#define STRIP_T0 0
#define STRIP_TS 1
#define STRIP_SYNC()
#else
#define STRIP_T0 ((int)threadIdx.x)
#define STRIP_TS ((int)blockDim.x)
#define STRIP_SYNC() __syncthreads()
#endif

// y-transform of grid column x of the strip (row i of the column comes from `load(i)`): out[j] = sum_i Q[j][i] strip[i][x] with the mirror symmetry
// Q[j][c-1-i] = (-1)^j Q[j][i] folded: qc[j*HP + i] = Q[j][i] for i < ceil(c/2), 0 beyond.  All inputs
// are read (mirror sums p, differences m) before `emit(j, value)` is called for the first row, so the
// transform may overwrite its own column.
template <typename Load, typename Emit>
__device__ __forceinline__ void strip_ytransform(const double* __restrict__ qc, Load load, int c, Emit emit) {
    double p[STRIP_HP], m[STRIP_HP];
    const int half = c >> 1;
#pragma unroll
    for (int s = 0; s < STRIP_HP; ++s) {
        if (s < half) {
            const double a = load(s), b = load(c - 1 - s);
            p[s] = a + b; m[s] = a - b;
        } else if (s == half && (c & 1)) {
            p[s] = load(s); m[s] = 0.0;
        } else {
            p[s] = 0.0; m[s] = 0.0;
        }
    }
    // four output modes of one parity at a time: four independent DFMA chains, every mirror sum / difference
    // is used four times from registers; the matrix entries are warp-uniform constant-bank loads
    auto run = [&](const double (&f)[STRIP_HP], int j0) {
        for (int j = j0; j < c; j += 8) {
            const double* q0 = qc + j * STRIP_HP;
            const int j1 = (j + 2 < c) ? 2 : 0, j2 = (j + 4 < c) ? 4 : 0, j3 = (j + 6 < c) ? 6 : 0;
            const double* q1 = q0 + j1 * STRIP_HP;
            const double* q2 = q0 + j2 * STRIP_HP;
            const double* q3 = q0 + j3 * STRIP_HP;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int i = 0; i < STRIP_HP; ++i) {
                a0 = fma(q0[i], f[i], a0);
                a1 = fma(q1[i], f[i], a1);
                a2 = fma(q2[i], f[i], a2);
                a3 = fma(q3[i], f[i], a3);
            }
            emit(j, a0);
            if (j1) emit(j + 2, a1);
            if (j2) emit(j + 4, a2);
            if (j3) emit(j + 6, a3);
        }
    };
    run(p, 0);          // even modes: symmetric part
    run(m, 1);          // odd modes: antisymmetric part
}

// The same transform for a strip height known at compile time (the main strips of a plan): both loops are
// fully unrolled, so every matrix entry is an IMMEDIATE constant-bank operand of its DFMA -- no load
// instruction at all -- and the chains of different modes interleave freely.
template <int CT, typename Load, typename Emit>
__device__ __forceinline__ void strip_ytransform_ct(const StripQ& Q, Load load, Emit emit) {
    constexpr int half = CT / 2, hp = (CT + 1) / 2;
    double p[hp], m[hp];
#pragma unroll
    for (int s = 0; s < half; ++s) {
        const double a = load(s), b = load(CT - 1 - s);
        p[s] = a + b; m[s] = a - b;
    }
    if (CT & 1) { p[half] = load(half); m[half] = 0.0; }
#pragma unroll
    for (int j0 = 0; j0 < CT; j0 += 8) {         // groups of four modes per parity: bounded register use
        double e[4], o[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) { e[g] = 0.0; o[g] = 0.0; }
#pragma unroll
        for (int i = 0; i < hp; ++i) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (j0 + 2 * g < CT) e[g] = fma(Q.q[0][(j0 + 2 * g) * STRIP_HP + i], p[i], e[g]);
                if (j0 + 2 * g + 1 < CT && i < half) o[g] = fma(Q.q[0][(j0 + 2 * g + 1) * STRIP_HP + i], m[i], o[g]);
            }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (j0 + 2 * g < CT) emit(j0 + 2 * g, e[g]);
            if (j0 + 2 * g + 1 < CT) emit(j0 + 2 * g + 1, o[g]);
        }
    }
}

// One strip of one slab column.  FINAL = false: Gf/Gl <- first / last row of A_strip^{-1} R_strip.
// FINAL = true: S_strip <- A_strip^{-1} (R_strip - oy * separator values on the adjacent rows), and the
// separator row below the strip is copied from U.
// CT > 0: the main strips (the first n_tall of a column, class 0) have height CT and take the compile-time
// transform, the remainder strip goes through the generic code of the same kernel (block-uniform branch), so a
// pass over the grid is ONE launch; CT = 0: all heights from the plan (any <= 32).
// The launch covers strips p_off .. p_off + p_cnt - 1 of every column.
// dot.partials != null (FINAL only): the kernel also leaves <R, S> in dot.out -- every CTA one partial (the
// strip's right-hand side is read a second time, from L2 as a rule: it was loaded by this CTA a few microseconds
// earlier), the last CTA to finish adds them in index order and runs the PCG epilogue of k_dot.
struct StripDot {
    double* partials;          // one per CTA of the launch; null: no dot
    unsigned int* ticket;      // zero before the launch, reset by the last CTA
    double* out;
    DotEpilogue epi;
};

// Sum of `v` over the CTA -> dot.partials[blockIdx.x]; the last CTA of the launch adds all partials in index
// order (fixed order: the result does not depend on the scheduling), stores the sum and runs the epilogue.
__device__ __forceinline__ void strip_dot_finish(const StripDot& dot, double v) {
#ifdef SPUQ_EMU
    // one emulated thread per block, blocks in order: the running sum lives in partials[0]
    if (blockIdx.x == 0) dot.partials[0] = 0.0;
    dot.partials[0] += v;
    if (blockIdx.x == gridDim.x - 1) { dot.out[0] = dot.partials[0]; dot_epilogue(dot.epi, dot.partials[0]); }
#else
    __shared__ double sh[STRIP_THREADS / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < STRIP_THREADS / 32; ++w) a += sh[w];
        dot.partials[blockIdx.x] = a;
        __threadfence();
        is_last = atomicAdd(dot.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double a = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += STRIP_THREADS) a += __ldcg(dot.partials + i);
    a = warp_sum(a);
    __syncthreads();
    if (lane == 0) sh[wid] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < STRIP_THREADS / 32; ++w) t += sh[w];
        dot.out[0] = t;
        dot_epilogue(dot.epi, t);
        *dot.ticket = 0u;
    }
#endif
}

template <bool FINAL, int CT>
__global__ void __launch_bounds__(STRIP_THREADS, 2)
k_strip(const __grid_constant__ StripQ Q, StripView sv, int32_t L, int32_t p_off, int32_t p_cnt,
        const double* __restrict__ R, double* __restrict__ S,
        double* __restrict__ Gf, double* __restrict__ Gl, const double* __restrict__ U, const double* gate,
        StripDot dot) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
#ifdef SPUQ_EMU
    if (threadIdx.x != 0) return;
    std::vector<double> smem_store((size_t)STRIP_CMAX * sv.sp + sv.sepv_n);
    double* smem = smem_store.data();
#else
    extern __shared__ __align__(16) double smem[];
#endif
    const int nx = sv.nx, sp = sv.sp;                   // row length in the slab, row pitch in shared memory
    const int p = p_off + (int)(blockIdx.x % (unsigned)p_cnt);
    const int64_t col = blockIdx.x / (unsigned)p_cnt;
    if (col >= L) return;
    // class, height and first row from the parameters alone (no table look-up): warp-uniform values, so the
    // transform-matrix addresses below are uniform too
    const bool tall = p < sv.n_tall;
    const bool ct_path = CT > 0 && tall;                // block-uniform
    const int cls = tall ? 0 : 1;
    const int c = ct_path ? CT : sv.ch[cls];
    const int y0 = (p < sv.n_tall) ? p * (sv.ch[0] + 1) : sv.n_tall * (sv.ch[0] + 1) + (p - sv.n_tall) * (sv.ch[1] + 1);
    const double* qc = Q.q[cls];                        // [c][HP], constant bank
    const int PX = sv.PX, qx = sv.qx, cxmax = sv.cxmax;
    const int qp = qx | 1;                              // odd pitch of the x-separator values
    double* sepv = smem;                                // [c][qp]
    double* strip = sepv + sv.sepv_n;                   // [c][sp]
    const int64_t N = sv.ld;
    const double* src = R + col * N + (int64_t)y0 * nx;
    double dsep = 0.0;                                  // <R, S> on the separator row below the strip (FINAL, dot)

    // ---- phase 0: strip into shared memory -------------------------------------------------------------
#ifndef SPUQ_EMU
    // bulk copies (TMA) issued by one thread, no registers and no load instructions spent on the strip; the
    // rest of the block waits on the mbarrier.  sp == nx: the strip is c*nx contiguous doubles, ONE copy;
    // padded rows (sp > nx, bank layout of phase 2): one copy per row.
    __shared__ __align__(8) unsigned long long strip_bar;
    const uint32_t row_bytes = (uint32_t)nx * 8u;
    const bool bulk = (row_bytes & 15u) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (bulk) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&strip_bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * (uint32_t)c) : "memory");
            if (sp == nx) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"((uint32_t)__cvta_generic_to_shared(strip)), "l"(src), "r"(row_bytes * (uint32_t)c), "r"(bar) : "memory");
            } else {
                for (int i = 0; i < c; ++i)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"((uint32_t)__cvta_generic_to_shared(strip + (size_t)i * sp)), "l"(src + (size_t)i * nx),
                                   "r"(row_bytes), "r"(bar) : "memory");
            }
        }
        __syncthreads();                                     // barrier initialised before anyone waits on it
        asm volatile(
            "{\n.reg .pred P1;\nSTRIP_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra STRIP_DONE;\nbra STRIP_WAIT;\nSTRIP_DONE:\n}\n"
            ::"r"(bar) : "memory");
    } else
#endif
    for (int i = STRIP_T0; i < c * nx; i += STRIP_TS) strip[(size_t)(i / nx) * sp + i % nx] = src[i];
    if (FINAL) {
        STRIP_SYNC();
        const double oy = sv.oy;
        const int ja = sv.u_above0 + p;                               // separator above / below this strip
        const bool has_a = ja >= 0, has_b = (p < sv.P - 1) || sv.has_bottom;
        const double* ua = U + (int64_t)ja * sv.u_js + col * sv.u_cs;
        const double* ub = U + (int64_t)(ja + 1) * sv.u_js + col * sv.u_cs;
        for (int x = STRIP_T0; x < nx; x += STRIP_TS) {
            if (has_a) strip[x] -= oy * ua[x];
            if (has_b) strip[(size_t)(c - 1) * sp + x] -= oy * ub[x];
        }
        // separator row below this strip
        if (has_b) {
            double* dst = S + col * N + (int64_t)(y0 + c) * nx;
            const double* rs = R + col * N + (int64_t)(y0 + c) * nx;
            const bool with_dot = dot.partials != nullptr;
            for (int x = STRIP_T0; x < nx; x += STRIP_TS) {
                const double u = ub[x];
                if (with_dot) dsep = fma(u, rs[x], dsep);
                dst[x] = u;
            }
        }
    }
    STRIP_SYNC();

    // ---- phase 1: sine transform along y, one thread per grid column -------------------------------
    for (int x = STRIP_T0; x < nx; x += STRIP_TS) {
        auto get = [&](int i) { return strip[(size_t)i * sp + x]; };
        auto put = [&](int j, double a) { strip[(size_t)j * sp + x] = a; };
        if (ct_path) strip_ytransform_ct<(CT > 0 ? CT : 1)>(Q, get, put);
        else strip_ytransform(qc, get, c, put);
    }
    STRIP_SYNC();

    // ---- phase 2a: chunk-local Thomas sweeps along x, one thread per (mode, chunk) --------------------
    // Lanes of a warp take consecutive chunks.  With the plan's preferred layout (16 chunks per row at pitch 33,
    // rows at a pitch that is a multiple of 16 doubles) the 16 lanes of a half-warp always sit on 16 different
    // bank pairs, whether they share a row or not; other grid widths keep an odd chunk pitch (no conflicts inside
    // a row, two-way where a half-warp straddles two rows).
    const double ox = sv.ox;
    const double* ipx = sv.ipivx[cls];
    for (int task = STRIP_T0; task < c * PX; task += STRIP_TS) {
        const int j = task / PX, ch = task - j * PX;
        const int x0 = sv.xstart[ch], len = sv.xlen[ch];
        double* r = strip + (size_t)j * sp + x0;
        const double* ip = ipx + (size_t)j * cxmax;
        // The chunk lives in registers for both sweeps (static indices, guards on the chunk length); a step costs
        // its one dependent DFMA.  The entry after the chunk's last one is zero, so the first step of the
        // backward sweep needs no special case: fma(-ox pv, 0, rr pv) = rr pv.
        double rr[STRIP_CX];
#pragma unroll
        for (int t = 0; t < STRIP_CX; ++t) rr[t] = (t < len) ? r[t] : 0.0;
#pragma unroll
        for (int t = 1; t < STRIP_CX; ++t)
            if (t < len) rr[t] = fma(-ox * __ldg(ip + t - 1), rr[t - 1], rr[t]);
#pragma unroll
        for (int t = STRIP_CX - 1; t >= 0; --t) {
            if (t < len) {
                const double pv = __ldg(ip + t);
                rr[t] = fma(-ox * pv, (t + 1 < STRIP_CX) ? rr[(t + 1 < STRIP_CX) ? t + 1 : t] : 0.0, rr[t] * pv);
            }
        }
#pragma unroll
        for (int t = 0; t < STRIP_CX; ++t)
            if (t < len) r[t] = rr[t];
    }
    STRIP_SYNC();

    if (qx > 0) {
        // ---- phase 2b: x-separator system of every mode (tridiagonal, qx unknowns) ---------------------
        // right-hand sides by all threads, one per (mode, separator): lanes along the separators of a row
        for (int task = STRIP_T0; task < c * qx; task += STRIP_TS) {
            const int j = task / qx, i = task - j * qx;
            const double* row = strip + (size_t)j * sp;
            const int xs = sv.xstart[i] + sv.xlen[i];              // separator between chunk i and i+1
            sepv[(size_t)j * qp + i] = row[xs] - ox * (row[xs - 1] + row[xs + 1]);
        }
        STRIP_SYNC();
        // the recurrences, one thread per mode, on the separator values alone (odd pitch: no bank conflicts)
        for (int j = STRIP_T0; j < c; j += STRIP_TS) {
            const double* sip = sv.sx_ipiv[cls] + (size_t)j * qx;
            const double* sof = sv.sx_off[cls] + (size_t)j * qx;
            double* sv_j = sepv + (size_t)j * qp;
            double z = 0.0;
            for (int i = 0; i < qx; ++i) {
                double rhs = sv_j[i];
                if (i > 0) rhs -= __ldg(sof + i - 1) * __ldg(sip + i - 1) * z;
                z = rhs;
                sv_j[i] = z;
            }
            double u = 0.0;
            for (int i = qx - 1; i >= 0; --i) {
                const double e = (i + 1 < qx) ? __ldg(sof + i) * u : 0.0;
                u = (sv_j[i] - e) * __ldg(sip + i);
                sv_j[i] = u;
            }
        }
        STRIP_SYNC();
        // ---- phase 2c: chunks corrected with their spikes; the thread of chunk ch also stores the separator
        // value to its right ------------------------------------------------------------------------------
        for (int task = STRIP_T0; task < c * PX; task += STRIP_TS) {
            const int j = task / PX, ch = task - j * PX;
            const int x0 = sv.xstart[ch], len = sv.xlen[ch];
            const int lc = (len == sv.xlen_cls_len[0]) ? 0 : 1;
            const double* spk = sv.spike[cls] + ((size_t)lc * c + j) * cxmax;
            const double sl = (ch > 0) ? sepv[(size_t)j * qp + ch - 1] : 0.0;
            const double sr = (ch < qx) ? sepv[(size_t)j * qp + ch] : 0.0;
            double* r = strip + (size_t)j * sp + x0;
            const double a = ox * sl, b = ox * sr;
#pragma unroll 8
            for (int t = 0; t < len; ++t) r[t] -= a * __ldg(spk + t) + b * __ldg(spk + len - 1 - t);
            if (ch < qx) r[len] = sr;
        }
        STRIP_SYNC();
    }

    // ---- phase 3: transform back -------------------------------------------------------------------
    if (FINAL) {
        double* dst = S + col * N + (int64_t)y0 * nx;
        const bool with_dot = dot.partials != nullptr;
        double dsum = 0.0;
        for (int x = STRIP_T0; x < nx; x += STRIP_TS) {
            auto get = [&](int i) { return strip[(size_t)i * sp + x]; };
            auto put = [&](int j, double a) {
                if (with_dot) dsum = fma(a, src[(size_t)j * nx + x], dsum);      // before the store: R and S may alias
                dst[(size_t)j * nx + x] = a;
            };
            if (ct_path) strip_ytransform_ct<(CT > 0 ? CT : 1)>(Q, get, put);
            else strip_ytransform(qc, get, c, put);
        }
        if (with_dot) strip_dot_finish(dot, dsum + dsep);
    } else {
        double* gf = Gf + (col * sv.gf_ld + p) * nx;
        double* gl = Gl + (col * sv.gf_ld + p) * nx;
        for (int x = STRIP_T0; x < nx; x += STRIP_TS) {
            // rows 0 and c-1 of Q u: Q[0][j] and Q[c-1][j] = (-1)^j Q[0][j]
            double se = 0.0, so = 0.0;
            for (int j = 0; j < c; ++j) {
                // Q[0][j] = Q[j][0] (symmetric) = qc[j*HP + 0]
                const double t = qc[j * STRIP_HP] * strip[(size_t)j * sp + x];
                if (j & 1) so += t; else se += t;
            }
            gf[x] = se + so;
            gl[x] = se - so;
        }
    }
}

// Z(j0 + i, col, x) = R[col][ysep_i][x] - oy * (Gl[col][i][x] + Gf[col][i+1][x]) for the q_loc local separators
// (i + 1 = gf_ld - 1 names the slot that holds the first strip of the NEXT rank); Z has the layout of U.
__global__ void __launch_bounds__(256)
k_sep_assemble(StripView sv, int32_t L, const double* __restrict__ R, const double* __restrict__ Gf,
               const double* __restrict__ Gl, double* __restrict__ Z, const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    const int nx = sv.nx, q = sv.q_loc;
    const int64_t total = (int64_t)L * q * nx;
    const int64_t N = sv.ld;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(t % nx);
        const int64_t r = t / nx;
        const int i = (int)(r % q);
        const int64_t col = r / q;
        const int ys = sv.ystart[i] + sv.ch[sv.ycls[i]];
        Z[(int64_t)(sv.u_above0 + 1 + i) * sv.u_js + col * sv.u_cs + x] =
            R[col * N + (int64_t)ys * nx + x]
            - sv.oy * (Gl[(col * sv.gf_ld + i) * nx + x] + Gf[(col * sv.gf_ld + i + 1) * nx + x]);
    }
}

// Tridiagonal solves of the separator system in the transformed domain, in place on X[col][j][k]:
// one thread per (column, mode k), lanes along k.  ipiv[j*nx + k], off[j*nx + k] (coupling j <-> j+1).
__global__ void __launch_bounds__(128)
k_thomas_sep(int32_t nx, int32_t q, int32_t L, int64_t js, int64_t cs, const double* __restrict__ ipiv,
             const double* __restrict__ off, double* __restrict__ X, const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)L * nx) return;
    const int32_t k = (int32_t)(t % nx);
    const int64_t col = t / nx;
    double* x = X + col * cs + k;            // element (j, col, k) at j * js + col * cs + k
    double z = x[0];
    for (int32_t j = 1; j < q; ++j) {
        z = x[(int64_t)j * js] - off[(int64_t)(j - 1) * nx + k] * ipiv[(int64_t)(j - 1) * nx + k] * z;
        x[(int64_t)j * js] = z;
    }
    double u = z * ipiv[(int64_t)(q - 1) * nx + k];
    x[(int64_t)(q - 1) * js] = u;
    for (int32_t j = q - 2; j >= 0; --j) {
        u = (x[(int64_t)j * js] - off[(int64_t)j * nx + k] * u) * ipiv[(int64_t)j * nx + k];
        x[(int64_t)j * js] = u;
    }
}

// ---- separator system in ONE kernel --------------------------------------------------------------------------
// One CTA per slab column keeps the q separator rows of that column (q * nx doubles, 72 KB on configuration 4)
// in shared memory from the right-hand side to the solution: mirror split, transform to the sine modes of
// length nx, the q x q tridiagonal solve of every mode, transform back, mirror merge.  A thread owns CPT
// transform columns of ONE parity (even modes use the mirror sums, odd modes the differences) with their q
// accumulators in registers, so the tridiagonal solve along the separators never leaves the register file and,
// with CPT = 2, every shared-memory operand (a 16-byte broadcast load) feeds four DFMAs.  The only traffic
// besides the q rows is the transform matrix (4 * (nx/2)^2 doubles per column, streamed from L2, coalesced
// across the threads).  Replaces eight launches (assembly, split, four GEMMs, Thomas, merge) and the
// Z <-> Z2 round trips; conditions: nx even, nx <= 512, q <= 32, q rows fit shared memory.
constexpr int SCHUR_RMAX = 32;
constexpr int SCHUR_R2MAX = 24;       // two columns per thread up to this many separators (registers)
template <int CPT> struct SchurShape { static constexpr int threads = CPT == 2 ? 256 : 512; };

// acc[k * RT + r] = sum_i F[r][i] * Qm[i][c[k]] for the rows r < q and the thread's columns k < CPT
// (F: shared memory, row pitch ps, 16-byte aligned rows)
template <int RT, int CPT>
__device__ __forceinline__ void schur_half_product(const double* __restrict__ F, int ps, int h, int q,
                                                   const double* __restrict__ Qm, const int (&c)[2], double* acc) {
#pragma unroll
    for (int r = 0; r < CPT * RT; ++r) acc[r] = 0.0;
    int i = 0;
#pragma unroll 2
    for (; i + 1 < h; i += 2) {
        double q0[CPT], q1[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            q0[k] = __ldg(Qm + (size_t)i * h + c[k]);
            q1[k] = __ldg(Qm + (size_t)(i + 1) * h + c[k]);
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            if (r < q) {
                const double2 f = *reinterpret_cast<const double2*>(F + (size_t)r * ps + i);
#pragma unroll
                for (int k = 0; k < CPT; ++k) {
                    acc[k * RT + r] = fma(f.x, q0[k], acc[k * RT + r]);
                    acc[k * RT + r] = fma(f.y, q1[k], acc[k * RT + r]);
                }
            }
        }
    }
    if (i < h) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const double q0 = __ldg(Qm + (size_t)i * h + c[k]);
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if (r < q) acc[k * RT + r] = fma(F[(size_t)r * ps + i], q0, acc[k * RT + r]);
        }
    }
}

// assemble != 0: the right-hand side is formed here from R, Gf, Gl (what k_sep_assemble writes; one-GPU layout of
// the strips, u_above0 = -1); else it is read from Z.  The solution is written to Z (layout sv.u_js / sv.u_cs).
template <int RT, int CPT>
__global__ void __launch_bounds__(SchurShape<CPT>::threads, (CPT == 2 && RT <= 20) ? 2 : 1)
k_schur_fused(StripView sv, int32_t L, int32_t h, int32_t assemble, const double* __restrict__ R,
              const double* __restrict__ Gf, const double* __restrict__ Gl, double* __restrict__ Z,
              const double* __restrict__ qsplit, const double* __restrict__ ipiv, const double* __restrict__ off,
              const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    const int nx = sv.nx, q = sv.q;
    const int hp2 = (h + 1) & ~1, ps = 2 * hp2;       // sums in [0, h), differences in [hp2, hp2 + h) of a row
    const int per = CPT == 2 ? (h + 1) / 2 : h;       // threads per parity
    const int n_thr = 2 * per;                        // active threads
#ifdef SPUQ_EMU
    if (threadIdx.x != 0) return;
    std::vector<double> smem_store((size_t)q * ps), hold((size_t)(n_thr + 1) * CPT * RT);
    double* buf = smem_store.data();
    double* acc = nullptr;
#define SCHUR_FOR_T for (int t = 0; (acc = &hold[(size_t)t * CPT * RT], t < n_thr); ++t)   // acc: the "registers" of emulated thread t
#else
    extern __shared__ __align__(16) double smem[];
    double* buf = smem;
    double acc[CPT * RT];
    const int t = (int)threadIdx.x;
#define SCHUR_FOR_T if (t < n_thr)
#endif
    const int64_t col = blockIdx.x;
    if (col >= L) return;
    const int64_t N = sv.ld;
    const size_t hh = (size_t)h * h;

    // ---- right-hand side, split into mirror sums and differences ------------------------------------------
    for (int e = STRIP_T0; e < q * h; e += STRIP_TS) {
        const int j = e / h, i = e - j * h, i2 = nx - 1 - i;
        double a, b;
        if (assemble) {
            const int ys = sv.ystart[j] + sv.ch[sv.ycls[j]];
            const double* rr = R + col * N + (int64_t)ys * nx;
            const double* gl = Gl + (col * sv.gf_ld + j) * nx;
            const double* gf = Gf + (col * sv.gf_ld + j + 1) * nx;
            a = rr[i] - sv.oy * (gl[i] + gf[i]);
            b = rr[i2] - sv.oy * (gl[i2] + gf[i2]);
        } else {
            const double* z = Z + (int64_t)j * sv.u_js + col * sv.u_cs;
            a = z[i]; b = z[i2];
        }
        buf[(size_t)j * ps + i] = a + b;
        buf[(size_t)j * ps + hp2 + i] = a - b;
    }
    STRIP_SYNC();
    // the thread's columns: parity par, modes c[0] (and c[1] = c[0] + per; a thread whose second column does not
    // exist computes its first one twice and stores it once)
#define SCHUR_COLUMNS                                                            \
        const int par = t >= per ? 1 : 0;                                        \
        const int c[2] = {t - par * per, (CPT == 2 && t - par * per + per < h) ? t - par * per + per : t - par * per}; \
        const int nvalid = (CPT == 2 && c[1] != c[0]) ? 2 : 1;
    // ---- to the modes, tridiagonal solve along the separators in registers -----------------------------------
    SCHUR_FOR_T {
        SCHUR_COLUMNS
        (void)nvalid;
        schur_half_product<RT, CPT>(buf + par * hp2, ps, h, q, qsplit + par * hh, c, acc);
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int tc = par * h + c[k];                                     // position in the tables
            double* a = acc + k * RT;
            // the arithmetic of k_thomas_sep
            double z = a[0];
#pragma unroll
            for (int j = 1; j < RT; ++j)
                if (j < q) { z = a[j] - __ldg(off + (size_t)(j - 1) * nx + tc) * __ldg(ipiv + (size_t)(j - 1) * nx + tc) * z; a[j] = z; }
            // backward: the last separator starts from u = 0 (its off-diagonal entry multiplies zero); written
            // without a test for j == q - 1, which the compiler turns into a dynamically indexed -- local-memory --
            // accumulator array
            double u = 0.0;
#pragma unroll
            for (int j = RT - 1; j >= 0; --j)
                if (j < q) { u = (a[j] - __ldg(off + (size_t)j * nx + tc) * u) * __ldg(ipiv + (size_t)j * nx + tc); a[j] = u; }
        }
    }
    STRIP_SYNC();
    SCHUR_FOR_T {
        SCHUR_COLUMNS
#pragma unroll
        for (int k = 0; k < CPT; ++k)
            if (k < nvalid) {
#pragma unroll
                for (int r = 0; r < RT; ++r)
                    if (r < q) buf[(size_t)r * ps + par * hp2 + c[k]] = acc[k * RT + r];
            }
    }
    STRIP_SYNC();
    // ---- back to the grid -------------------------------------------------------------------------------------
    SCHUR_FOR_T {
        SCHUR_COLUMNS
        (void)nvalid;
        schur_half_product<RT, CPT>(buf + par * hp2, ps, h, q, qsplit + (2 + par) * hh, c, acc);
    }
    STRIP_SYNC();
    SCHUR_FOR_T {
        SCHUR_COLUMNS
#pragma unroll
        for (int k = 0; k < CPT; ++k)
            if (k < nvalid) {
#pragma unroll
                for (int r = 0; r < RT; ++r)
                    if (r < q) buf[(size_t)r * ps + par * hp2 + c[k]] = acc[k * RT + r];
            }
    }
    STRIP_SYNC();
    for (int e = STRIP_T0; e < q * h; e += STRIP_TS) {
        const int j = e / h, i = e - j * h;
        const double pp = buf[(size_t)j * ps + i], mm = buf[(size_t)j * ps + hp2 + i];
        double* z = Z + (int64_t)j * sv.u_js + col * sv.u_cs;
        z[i] = pp + mm;
        z[nx - 1 - i] = pp - mm;
    }
#undef SCHUR_FOR_T
#undef SCHUR_COLUMNS
}

}  // namespace spuq
