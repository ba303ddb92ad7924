// kernels_blas.cuh -- K2/K3: slab dot products and fused vector updates with device scalars.
//
// Replace MultiVector.__inner__ (application/egsz/multi_vector.py:147-152, per-mu np.dot summed
// in dict order) and the copy-then-in-place vector arithmetic of linalg/vector.py:103-133 that
// application/egsz/pcg.py:26-51 executes.  Scalars (zeta, alpha) never leave the device.
//
// Reductions are deterministic: a fixed grid of REDUCE_BLOCKS blocks writes one partial each;
// the last block to finish (atomic ticket) adds the partials in index order with one warp.
#pragma once
#include "common.h"
#include "reduce_common.cuh"

namespace spuq {

constexpr int REDUCE_THREADS = 256;
constexpr int REDUCE_MAX_BLOCKS = 1024;
// scratch layout: [0, 2*REDUCE_MAX_BLOCKS) doubles of partials, then one uint32 ticket
constexpr int64_t REDUCE_SCRATCH_BYTES = 2 * REDUCE_MAX_BLOCKS * 8 + 64;

#ifndef SPUQ_EMU
// NOUT = 1: out[0] = <x,y>;  NOUT = 2: out[0] = <x,y>, out[1] = <x,x>
template <int NOUT>
__global__ void __launch_bounds__(REDUCE_THREADS)
k_dot(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
      double* __restrict__ out, double* __restrict__ partials, unsigned int* ticket,
      const double* gate, DotEpilogue epi) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    double sxy = 0.0, sxx = 0.0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const bool vec = ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0;
    if (vec) {
        const int64_t n2 = n >> 1;
        const double2* x2 = reinterpret_cast<const double2*>(x);
        const double2* y2 = reinterpret_cast<const double2*>(y);
        for (int64_t i = tid; i < n2; i += nthr) {
            const double2 a = x2[i], b = y2[i];
            sxy += a.x * b.x;
            sxy += a.y * b.y;
            if (NOUT == 2) { sxx += a.x * a.x; sxx += a.y * a.y; }
        }
        if ((n & 1) && tid == 0) {
            sxy += x[n - 1] * y[n - 1];
            if (NOUT == 2) sxx += x[n - 1] * x[n - 1];
        }
    } else {
        for (int64_t i = tid; i < n; i += nthr) {
            sxy += x[i] * y[i];
            if (NOUT == 2) sxx += x[i] * x[i];
        }
    }
    __shared__ double sh[2][REDUCE_THREADS / 32];
    __shared__ bool is_last;
    sxy = warp_sum(sxy);
    if (NOUT == 2) sxx = warp_sum(sxx);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wid] = sxy; sh[1][wid] = sxx; }
    __syncthreads();
    if (wid == 0) {
        double a = lane < REDUCE_THREADS / 32 ? sh[0][lane] : 0.0;
        double b = lane < REDUCE_THREADS / 32 ? sh[1][lane] : 0.0;
        a = warp_sum(a);
        if (NOUT == 2) b = warp_sum(b);
        if (lane == 0) {
            partials[blockIdx.x] = a;
            if (NOUT == 2) partials[REDUCE_MAX_BLOCKS + blockIdx.x] = b;
            __threadfence();
            const unsigned int done = atomicAdd(ticket, 1u);
            is_last = (done == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (is_last && wid == 0) {
        __threadfence();
        double a = 0.0, b = 0.0;
        for (int i = lane; i < (int)gridDim.x; i += 32) {
            a += partials[i];
            if (NOUT == 2) b += partials[REDUCE_MAX_BLOCKS + i];
        }
        a = warp_sum(a);
        if (NOUT == 2) b = warp_sum(b);
        if (lane == 0) {
            out[0] = a;
            if (NOUT == 2) out[1] = b;
            dot_epilogue(epi, a);
            *ticket = 0u;   // ready for the next reduction on this stream
        }
    }
}
#else
template <int NOUT>
void k_dot(int64_t n, const double* x, const double* y, double* out, double*, unsigned int*,
           const double* gate, DotEpilogue epi) {
    if (gate != nullptr && *gate != (double)SPUQ_PCG_RUNNING) return;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double sxy = 0.0, sxx = 0.0;
    for (int64_t i = 0; i < n; ++i) { sxy += x[i] * y[i]; sxx += x[i] * x[i]; }
    out[0] = sxy;
    if (NOUT == 2) out[1] = sxx;
    dot_epilogue(epi, sxy);
}
#endif

// out[s][i] = sum_mu coef[s][mu] * W[i + ldw*mu], SC samples per pass (coefficients staged in shared
// memory in chunks of COMBINE_LCHUNK columns).  One thread owns two consecutive rows; a warp reads 512
// contiguous bytes of a column per step: a pure streaming kernel, 8 n L bytes per SC samples.
#ifndef SPUQ_EMU
constexpr int COMBINE_LCHUNK = 512;
// out[s][i] = sum_mu coef[s][mu] * W[i + ldw*mu], SC samples per pass.  One thread owns four consecutive
// rows (two 16-byte loads per column, a warp reads 1 KB of a column per step) and keeps 4 x SC sums in
// registers (two rows when 8 samples are summed at once, to keep three blocks per SM in flight); the
// coefficients of a chunk of columns are staged in shared memory, sample index fastest,
// so that one 16-byte shared load feeds 8 FMAs.  A pure streaming kernel: 8 n L bytes per SC samples.
template <int SC, int R>
__global__ void __launch_bounds__(256)
k_combine(int64_t n, int32_t L, int32_t S, int32_t s0, const double* __restrict__ W, int64_t ldw,
          const double* __restrict__ coef, double* __restrict__ out, int64_t ldo) {
    static_assert(SC % 2 == 0 && R % 2 == 0, "samples and rows are handled in pairs");
    __shared__ __align__(16) double cs[COMBINE_LCHUNK][SC];
    const int64_t ngrp = (n + R - 1) / R;
    const bool aligned = ((ldw & 1) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    for (int64_t qb = (int64_t)blockIdx.x * blockDim.x; qb < ngrp; qb += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = R * (qb + threadIdx.x);
        const int nin = i < n ? (int)min((int64_t)R, n - i) : 0;
        const bool vec = aligned && nin == R;
        double acc[R][SC];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < SC; ++q) acc[r][q] = 0.0;
        for (int32_t l0 = 0; l0 < L; l0 += COMBINE_LCHUNK) {
            const int32_t lc = min(COMBINE_LCHUNK, L - l0);
            __syncthreads();
            for (int t = threadIdx.x; t < SC * lc; t += blockDim.x) {
                const int q = t / lc, l = t % lc;
                cs[l][q] = (s0 + q < S) ? coef[(int64_t)(s0 + q) * L + l0 + l] : 0.0;
            }
            __syncthreads();
            if (nin == 0) continue;
            const double* w = W + i + ldw * (int64_t)l0;
#pragma unroll 4
            for (int32_t l = 0; l < lc; ++l, w += ldw) {
                double x[R];
#pragma unroll
                for (int r = 0; r < R; ++r) x[r] = 0.0;
                if (vec) {
#pragma unroll
                    for (int r = 0; r < R; r += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(w + r);
                        x[r] = v.x; x[r + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) if (r < nin) x[r] = w[r];
                }
#pragma unroll
                for (int q = 0; q < SC; q += 2) {
                    const double2 c2 = *reinterpret_cast<const double2*>(&cs[l][q]);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        acc[r][q] = fma(c2.x, x[r], acc[r][q]);
                        acc[r][q + 1] = fma(c2.y, x[r], acc[r][q + 1]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < SC; ++q)
            if (s0 + q < S) {
#pragma unroll
                for (int r = 0; r < R; ++r) if (r < nin) out[(int64_t)(s0 + q) * ldo + i + r] = acc[r][q];
            }
    }
}

#else
// host-emulated build: the same sums, evaluated by the first emulated thread
template <int SC, int R>
void k_combine(int64_t n, int32_t L, int32_t S, int32_t s0, const double* W, int64_t ldw, const double* coef,
               double* out, int64_t ldo) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int q = 0; q < SC && s0 + q < S; ++q)
        for (int64_t i = 0; i < n; ++i) {
            double a = 0.0;
            for (int32_t l = 0; l < L; ++l) a = std::fma(coef[(int64_t)(s0 + q) * L + l], W[i + ldw * (int64_t)l], a);
            out[(int64_t)(s0 + q) * ldo + i] = a;
        }
}
#endif

// out[:, mu] = -b0[mu] * W[:, mu] + bu[mu] * W[:, up[mu]] + bd[mu] * W[:, dn[mu]]  (absent neighbours: index < 0)
// The neighbour combination of one expansion term m, for every column at once: what the residual
// estimator feeds to its flux forms (residual_estimator2.py:111-131) and, with the block matrices,
// the inner sum of MultiOperator.apply.  Same operation order as the reference (no contraction), so
// the result is bit-identical to the numpy expression.
__global__ void __launch_bounds__(256)
k_neighbour_combine(int64_t n, int32_t L, const double* __restrict__ W, int64_t ldw, const int32_t* __restrict__ up,
                    const int32_t* __restrict__ dn, const double* __restrict__ bu, const double* __restrict__ bd,
                    const double* __restrict__ b0, double* __restrict__ out, int64_t ldo) {
    const int32_t mu = blockIdx.y;
    if (mu >= L) return;
    const int32_t u = up[mu], d = dn[mu];
    const double cu = bu[mu], cd = bd[mu], c0 = -b0[mu];
    const double* w = W + ldw * (int64_t)mu;
    const double* wu = u >= 0 ? W + ldw * (int64_t)u : nullptr;
    const double* wd = d >= 0 ? W + ldw * (int64_t)d : nullptr;
    double* o = out + ldo * (int64_t)mu;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double r = __dmul_rn(c0, w[i]);
        if (wu) r = __dadd_rn(r, __dmul_rn(cu, wu[i]));
        if (wd) r = __dadd_rn(r, __dmul_rn(cd, wd[i]));
        o[i] = r;
    }
}

// y += a*x
__global__ void __launch_bounds__(256)
k_axpy(int64_t n, double alpha, const double* num, const double* den,
       const double* __restrict__ x, double* __restrict__ y, const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    const double a = ratio_of(alpha, num, den);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n; i += nthr) y[i] += a * x[i];
}

// w += a*v ; rho -= a*z        (pcg.py:47-48, a = zeta/alpha)
__global__ void __launch_bounds__(256)
k_axpy2(int64_t n, const double* num, const double* den, const double* __restrict__ v,
        double* __restrict__ w, const double* __restrict__ z, double* __restrict__ rho,
        const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    const double a = ratio_of(1.0, num, den);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n; i += nthr) {
        w[i] += a * v[i];
        rho[i] -= a * z[i];
    }
}

// v = s + b*v                  (pcg.py:51, b = zeta_new/zeta_old)
__global__ void __launch_bounds__(256)
k_xpay(int64_t n, const double* num, const double* den, const double* __restrict__ s,
       double* __restrict__ v, const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    const double b = ratio_of(1.0, num, den);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n; i += nthr) v[i] = s[i] + b * v[i];
}

// The direction update of a PCG pass together with the solution update of the SAME pass (pcg.py:51 and :47; the
// latter moved behind the preconditioner, where v is read anyway):  w += a v ; v = s + b v.  a = alpha of this
// pass (st[7], stored by the <z,v> epilogue before zeta rotates), b = zeta_new / zeta (st[3]).  The solution
// update also runs in the pass whose convergence test has just succeeded (status CONVERGED with numit = pass + 1):
// the reference updates w before it tests.
__global__ void __launch_bounds__(256)
k_xpay_w(int64_t n, const double* __restrict__ st, int32_t pass, const double* __restrict__ s,
         double* __restrict__ v, double* __restrict__ w) {
    const double status = __ldg(st + 4);
    const bool running = status == (double)SPUQ_PCG_RUNNING;
    const bool last = status == (double)SPUQ_PCG_CONVERGED && __ldg(st + 5) == (double)(pass + 1);
    if (!running && !last) return;
    const double a = __ldg(st + 7), b = __ldg(st + 3);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n; i += nthr) {
        const double vi = v[i];
        w[i] += a * vi;
        if (running) v[i] = s[i] + b * vi;
    }
}

__global__ void __launch_bounds__(256)
k_scal(int64_t n, double alpha, double* __restrict__ x) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n; i += nthr) x[i] *= alpha;
}

// out = a - b   (residual rho = f - A w, pcg.py:27)
__global__ void __launch_bounds__(256)
k_sub(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
      double* __restrict__ out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n; i += nthr) out[i] = a[i] - b[i];
}

// ---- PCG branch conditions evaluated on the device (one thread) --------------------------------
// state: [0] zeta  [1] alpha  [2] zeta_new  [3] <rho,rho>  [4] status  [5] numit  [6] eps^2
__global__ void k_pcg_check_zeta(double* st, int32_t pass, double* hist) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (st[4] != (double)SPUQ_PCG_RUNNING) return;
    const double zeta = st[0];
    if (hist) hist[pass - 1] = zeta;
    st[5] = (double)pass;
    if (zeta < 0.0) st[4] = (double)SPUQ_PCG_PRECOND_NOT_PD;       // pcg.py:33-36
    else if (zeta <= st[6]) st[4] = (double)SPUQ_PCG_CONVERGED;   // pcg.py:37-38
}

__global__ void k_pcg_check_alpha(double* st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (st[4] != (double)SPUQ_PCG_RUNNING) return;
    const double alpha = st[1];
    if (alpha == 0.0) st[4] = (double)SPUQ_PCG_SINGULAR;          // pcg.py:42-43
    else if (alpha < 0.0) st[4] = (double)SPUQ_PCG_NOT_PD;        // pcg.py:44-45
}

// st[dst] = st[src] while the solve is running: takes an all-reduced scalar over from its staging slot
__global__ void k_pcg_accept(double* st, int32_t src, int32_t dst) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (st[4] != (double)SPUQ_PCG_RUNNING) return;
    st[dst] = st[src];
}

__global__ void k_pcg_rotate(double* st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (st[4] != (double)SPUQ_PCG_RUNNING) return;
    st[0] = st[2];
}

}  // namespace spuq
