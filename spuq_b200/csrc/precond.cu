// precond.cu -- K4: exact mean-based preconditioner  S[:, mu] = A_0^{-1} R[:, mu]  for every column.
//
// Replaces PreconditioningOperator.apply (application/egsz/multi_operator2.py:151-165), which calls a
// sparse direct solver once per column per PCG iteration (ScipySolveOperator.apply,
// linalg/scipy_operator.py:45-49; FEniCSSolveOperator.apply, fem/fenics/fenics_operator.py:60-63).
//
// Scope of this round: the mean block of the synthetic configurations, i.e. a 5-point stencil with
// CONSTANT coefficients on an nx x ny grid,
//     A_0 = I_y (x) T_x + T_y (x) I_x,   T_x = tridiag(ox, d/2, ox),  T_y = tridiag(oy, d/2, oy)
// (T_x = tridiag(ox, d, ox) alone when ny == 1).  It is solved exactly (to rounding) by fast
// diagonalisation: T_x = Q diag(lam) Q with the orthogonal, symmetric sine matrix
// Q[j,k] = sqrt(2/(nx+1)) sin((j+1)(k+1) pi/(nx+1)), so
//     1. Rhat = R Q           (each grid row of each column times Q: one fp64 GEMM,
//                              (L*ny) x nx  times  nx x nx)
//     2. for every mode k and column: tridiagonal solve (T_y + lam_k I) uhat = rhat along y
//                             (Thomas with precomputed reciprocal pivots; lanes along k)
//     3. S = Uhat Q           (the same GEMM)
// For even nx the transform is split by the parity of the mode (Q[nx-1-j][k] = (-1)^k Q[j][k]): the
// sums and differences of mirrored entries of a grid row go through two half-size products, which
// halves the flops of steps 1 and 3 at the price of two element-wise passes.
// The two products are plain dense fp64 GEMMs ((L*ny) x nx by nx x nx): they go to cuBLAS DGEMM when
// libcublas can be loaded at run time (dlopen, no link-time dependency); k_rowgemm below is the
// in-house register-tiled fallback (CUDA cores) and the reference for the host-emulated build.
#include <algorithm>
#include <cmath>
#include <vector>

#include "precond.h"

#ifndef SPUQ_EMU
#include <dlfcn.h>
#endif

namespace spuq {

#ifndef SPUQ_EMU
// ---- minimal run-time binding of cuBLAS (v2 API) ---------------------------------------------------
struct CublasApi {
    int (*create)(void**) = nullptr;
    int (*destroy)(void*) = nullptr;
    int (*set_stream)(void*, cudaStream_t) = nullptr;
    int (*dgemm)(void*, int, int, int, int, int, const double*, const double*, int, const double*, int,
                 const double*, double*, int) = nullptr;
    bool ok = false;
};
static CublasApi& cublas_api() {
    static CublasApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libcublas.so.12", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libcublas.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.create = reinterpret_cast<decltype(api.create)>(dlsym(h, "cublasCreate_v2"));
            api.destroy = reinterpret_cast<decltype(api.destroy)>(dlsym(h, "cublasDestroy_v2"));
            api.set_stream = reinterpret_cast<decltype(api.set_stream)>(dlsym(h, "cublasSetStream_v2"));
            api.dgemm = reinterpret_cast<decltype(api.dgemm)>(dlsym(h, "cublasDgemm_v2"));
            api.ok = api.create && api.destroy && api.set_stream && api.dgemm;
        }
    }
    return api;
}
// C[M x nc] = A[M x nk] * Q[nk x nc], all row-major (row strides lda, ldc, nc)  ==  column-major
// C^T = Q^T * A^T with Q^T = the row-major Q read column-major
static bool blas_rowgemm(spuq_sg_precond* p, int64_t Mrows, int32_t nk, int32_t nc, const double* A, int32_t lda,
                         const double* Q, double* C, int32_t ldc, cudaStream_t stream) {
    CublasApi& api = cublas_api();
    if (!api.ok || Mrows > 0x7fffffffLL) return false;
    if (!p->blas && api.create(&p->blas) != 0) { p->blas = nullptr; return false; }
    if (api.set_stream(p->blas, stream) != 0) return false;
    const double one = 1.0, zero = 0.0;
    return api.dgemm(p->blas, /*CUBLAS_OP_N*/ 0, 0, nc, (int)Mrows, nk, &one, Q, nc, A, lda, &zero, C, ldc) == 0;
}
#endif

constexpr int GM = 128, GN = 64, GK = 8;      // CTA tile; 256 threads, each 8 x 4 outputs
constexpr int G_THREADS = 256;

#ifndef SPUQ_EMU
// C[M x nc] = A[M x nk] * Q[nk x nc]; A, C row-major with row strides lda, ldc; Q row-major, stride nc.
__global__ void __launch_bounds__(G_THREADS)
k_rowgemm(int64_t Mrows, int32_t nk, int32_t nc, const double* __restrict__ A, int32_t lda,
          const double* __restrict__ Q, double* __restrict__ C, int32_t ldc, const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    __shared__ double As[2][GK][GM + 2];
    __shared__ double Bs[2][GK][GN];
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * GM;
    const int col0 = blockIdx.y * GN;
    const int tr = tid / 16, tc = tid % 16;          // 16 x 16 thread grid: rows tr*8+i, cols tc+16*j
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    // loaders: A tile GM x GK (each thread 4 elements), B tile GK x GN (each thread 2 elements)
    const int a_r = tid / 2, a_k = (tid % 2) * 4;     // 128 rows x 2 groups of 4 k's
    const int b_k = tid / 32, b_c = (tid % 32) * 2;   // 8 k's x 32 pairs of columns
    double ra[4], rb[2];
    auto load = [&](int k0) {
        const int64_t r = row0 + a_r;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + a_k + q;
            ra[q] = (r < Mrows && k < nk) ? A[r * lda + k] : 0.0;
        }
        const int k = k0 + b_k;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int c = col0 + b_c + q;
            rb[q] = (k < nk && c < nc) ? Q[(int64_t)k * nc + c] : 0.0;
        }
    };
    auto store = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 4; ++q) As[buf][a_k + q][a_r] = ra[q];
        Bs[buf][b_k][b_c] = rb[0];
        Bs[buf][b_k][b_c + 1] = rb[1];
    };
    load(0);
    store(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < nk; k0 += GK) {
        const bool more = k0 + GK < nk;
        if (more) load(k0 + GK);
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            double av[8], bv[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = As[buf][k][tr * 8 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[buf][k][tc + 16 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
        if (more) store(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = row0 + tr * 8 + i;
        if (r >= Mrows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = col0 + tc + 16 * j;
            if (c < nc) C[r * ldc + c] = acc[i][j];
        }
    }
}
#else
void k_rowgemm(int64_t Mrows, int32_t nk, int32_t nc, const double* A, int32_t lda, const double* Q, double* C,
               int32_t ldc, const double* gate) {
    if (gate != nullptr && *gate != (double)SPUQ_PCG_RUNNING) return;
    if (threadIdx.x != 0 || blockIdx.x != 0 || blockIdx.y != 0) return;
    for (int64_t r = 0; r < Mrows; ++r)
        for (int c = 0; c < nc; ++c) {
            double s = 0.0;
            for (int k = 0; k < nk; ++k) s += A[r * lda + k] * Q[(int64_t)k * nc + c];
            C[r * ldc + c] = s;
        }
}
#endif

// Mirror sums and differences of every grid row (n even, h = n/2):  out[r][j] = in[r][j] + in[r][n-1-j],
// out[r][h+j] = in[r][j] - in[r][n-1-j]
__global__ void __launch_bounds__(256)
k_mirror_split(int64_t Mrows, int32_t n, const double* __restrict__ in, double* __restrict__ out, const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    const int32_t h = n / 2;
    const int64_t total = Mrows * h;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / h;
        const int32_t j = (int32_t)(t % h);
        const double a = in[r * n + j], b = in[r * n + (n - 1 - j)];
        out[r * n + j] = a + b;
        out[r * n + h + j] = a - b;
    }
}
// Its inverse, in place: X[r] holds P (first half) and M (second half); X[r][j] = P[j] + M[j],
// X[r][n-1-j] = P[j] - M[j].  One thread owns the pair (j, h-1-j): the four positions it reads are the
// four it writes.
__global__ void __launch_bounds__(256)
k_mirror_merge(int64_t Mrows, int32_t n, double* __restrict__ X, const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    const int32_t h = n / 2, hp = (h + 1) / 2;
    const int64_t total = Mrows * hp;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / hp;
        const int32_t j = (int32_t)(t % hp), j2 = h - 1 - j;
        double* x = X + r * n;
        const double p1 = x[j], m1 = x[h + j], p2 = x[j2], m2 = x[h + j2];
        x[j] = p1 + m1;
        x[h + j2] = p1 - m1;          // position n-1-j
        if (j2 != j) {
            x[j2] = p2 + m2;
            x[h + j] = p2 - m2;       // position n-1-j2
        }
    }
}

// Thomas solves along y, in place on X (L columns of ny x nx): one thread per (column, mode k),
// lanes along k.  ipiv[y*nx + k] = reciprocal pivot of (T_y + lam_k I) at row y.
__global__ void __launch_bounds__(128)
k_thomas_y(int32_t nx, int32_t ny, int32_t L, double oy, const double* __restrict__ ipiv,
           double* __restrict__ X, const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)L * nx) return;
    const int32_t k = (int32_t)(t % nx);
    const int64_t col = t / nx;
    double* x = X + col * (int64_t)nx * ny + k;
    // forward elimination: z_y = r_y - oy * ipiv_{y-1} * z_{y-1}
    double z = x[0];
    for (int32_t y = 1; y < ny; ++y) {
        z = x[(int64_t)y * nx] - oy * ipiv[(int64_t)(y - 1) * nx + k] * z;
        x[(int64_t)y * nx] = z;
    }
    // back substitution: u_y = (z_y - oy * u_{y+1}) * ipiv_y
    double u = z * ipiv[(int64_t)(ny - 1) * nx + k];
    x[(int64_t)(ny - 1) * nx] = u;
    for (int32_t y = ny - 2; y >= 0; --y) {
        u = (x[(int64_t)y * nx] - oy * u) * ipiv[(int64_t)y * nx + k];
        x[(int64_t)y * nx] = u;
    }
}

int mean_fd_destroy(spuq_sg_precond* p) {
    cudaSetDevice(p->device);
    cudaFree(p->q_dev);
    cudaFree(p->piv_dev);
    cudaFree(p->qsplit_dev);
    cudaFree(p->piv_split_dev);
    p->q_dev = p->piv_dev = p->qsplit_dev = p->piv_split_dev = nullptr;
#ifndef SPUQ_EMU
    if (p->blas) { cublas_api().destroy(p->blas); p->blas = nullptr; }
#endif
    return SPUQ_OK;
}

int64_t mean_fd_work_bytes(const spuq_sg_precond*, int64_t N, int32_t L) {
    return (int64_t)sizeof(double) * N * L;
}

int mean_fd_apply(spuq_sg_precond* p, int64_t N, int32_t L, const double* R, double* S, void* work,
                  cudaStream_t stream, const double* gate) {
    SPUQ_REQUIRE(N == (int64_t)p->nx * p->ny, "precond_mean_fd: slab rows %lld != nx*ny", (long long)N);
    SPUQ_REQUIRE(work != nullptr, "precond_mean_fd: workspace missing");
    double* T = static_cast<double*>(work);
    const int64_t Mrows = (int64_t)L * p->ny;
    auto rowgemm = [&](const double* src, double* dst, int32_t nk, int32_t nc, const double* Q) {
#ifndef SPUQ_EMU
        if (blas_rowgemm(p, Mrows, nk, nc, src, p->nx, Q, dst, p->nx, stream)) return;
#endif
        const dim3 g((unsigned)((Mrows + GM - 1) / GM), (unsigned)((nc + GN - 1) / GN));
        SPUQ_LAUNCH(k_rowgemm, g, dim3(G_THREADS), 0, stream, Mrows, nk, nc, src, p->nx, Q, dst, p->nx, gate);
    };
    const int64_t nthreads = (int64_t)L * p->nx;
    const int32_t h = p->half;
    if (h > 0 && R != S) {
        // sums / differences of mirrored entries -> S (scratch), two half-size products each way
        const size_t hh = (size_t)h * h;
        const unsigned eb = (unsigned)std::min<int64_t>((Mrows * h + 255) / 256, 1 << 20);
        SPUQ_LAUNCH(k_mirror_split, dim3(eb), dim3(256), 0, stream, Mrows, p->nx, R, S, gate);
        rowgemm(S, T, h, h, p->qsplit_dev);
        rowgemm(S + h, T + h, h, h, p->qsplit_dev + hh);
        SPUQ_LAUNCH(k_thomas_y, dim3((unsigned)((nthreads + 127) / 128)), dim3(128), 0, stream, p->nx, p->ny, L,
                    p->oy, p->piv_split_dev, T, gate);
        rowgemm(T, S, h, h, p->qsplit_dev + 2 * hh);
        rowgemm(T + h, S + h, h, h, p->qsplit_dev + 3 * hh);
        SPUQ_LAUNCH(k_mirror_merge, dim3(eb), dim3(256), 0, stream, Mrows, p->nx, S, gate);
    } else {
        rowgemm(R, T, p->nx, p->nx, p->q_dev);
        SPUQ_LAUNCH(k_thomas_y, dim3((unsigned)((nthreads + 127) / 128)), dim3(128), 0, stream, p->nx, p->ny, L,
                    p->oy, p->piv_dev, T, gate);
        rowgemm(T, S, p->nx, p->nx, p->q_dev);
    }
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

// ---- the solve split in two, for callers that finish the y direction themselves -------------------
// forward: T = (local tridiagonal solves along y)(sine transform in x of R), in the transformed domain
// (mode of transformed position c: mean_fd_mode_order); inverse: S = sine transform back of T.
int mean_fd_forward(spuq_sg_precond* p, int64_t N, int32_t L, const double* R, double* T, void* work,
                    cudaStream_t stream) {
    SPUQ_REQUIRE(N == (int64_t)p->nx * p->ny, "precond_mean_fd: slab rows %lld != nx*ny", (long long)N);
    SPUQ_REQUIRE(R != T, "precond_mean_fd_forward: input and output must not alias");
    const int64_t Mrows = (int64_t)L * p->ny;
    const double* gate = nullptr;
    auto rowgemm = [&](const double* src, double* dst, int32_t nk, int32_t nc, const double* Q) {
#ifndef SPUQ_EMU
        if (blas_rowgemm(p, Mrows, nk, nc, src, p->nx, Q, dst, p->nx, stream)) return;
#endif
        const dim3 g((unsigned)((Mrows + GM - 1) / GM), (unsigned)((nc + GN - 1) / GN));
        SPUQ_LAUNCH(k_rowgemm, g, dim3(G_THREADS), 0, stream, Mrows, nk, nc, src, p->nx, Q, dst, p->nx, gate);
    };
    const int64_t nthreads = (int64_t)L * p->nx;
    const int32_t h = p->half;
    if (h > 0) {
        SPUQ_REQUIRE(work != nullptr, "precond_mean_fd_forward: workspace missing");
        double* U = static_cast<double*>(work);
        const size_t hh = (size_t)h * h;
        const unsigned eb = (unsigned)std::min<int64_t>((Mrows * h + 255) / 256, 1 << 20);
        SPUQ_LAUNCH(k_mirror_split, dim3(eb), dim3(256), 0, stream, Mrows, p->nx, R, U, gate);
        rowgemm(U, T, h, h, p->qsplit_dev);
        rowgemm(U + h, T + h, h, h, p->qsplit_dev + hh);
        SPUQ_LAUNCH(k_thomas_y, dim3((unsigned)((nthreads + 127) / 128)), dim3(128), 0, stream, p->nx, p->ny, L,
                    p->oy, p->piv_split_dev, T, gate);
    } else {
        rowgemm(R, T, p->nx, p->nx, p->q_dev);
        SPUQ_LAUNCH(k_thomas_y, dim3((unsigned)((nthreads + 127) / 128)), dim3(128), 0, stream, p->nx, p->ny, L,
                    p->oy, p->piv_dev, T, gate);
    }
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

int mean_fd_inverse(spuq_sg_precond* p, int64_t N, int32_t L, const double* T, double* S, cudaStream_t stream) {
    SPUQ_REQUIRE(N == (int64_t)p->nx * p->ny, "precond_mean_fd: slab rows %lld != nx*ny", (long long)N);
    SPUQ_REQUIRE(T != S, "precond_mean_fd_inverse: input and output must not alias");
    const int64_t Mrows = (int64_t)L * p->ny;
    const double* gate = nullptr;
    auto rowgemm = [&](const double* src, double* dst, int32_t nk, int32_t nc, const double* Q) {
#ifndef SPUQ_EMU
        if (blas_rowgemm(p, Mrows, nk, nc, src, p->nx, Q, dst, p->nx, stream)) return;
#endif
        const dim3 g((unsigned)((Mrows + GM - 1) / GM), (unsigned)((nc + GN - 1) / GN));
        SPUQ_LAUNCH(k_rowgemm, g, dim3(G_THREADS), 0, stream, Mrows, nk, nc, src, p->nx, Q, dst, p->nx, gate);
    };
    const int32_t h = p->half;
    if (h > 0) {
        const size_t hh = (size_t)h * h;
        const unsigned eb = (unsigned)std::min<int64_t>((Mrows * h + 255) / 256, 1 << 20);
        rowgemm(T, S, h, h, p->qsplit_dev + 2 * hh);
        rowgemm(T + h, S + h, h, h, p->qsplit_dev + 3 * hh);
        SPUQ_LAUNCH(k_mirror_merge, dim3(eb), dim3(256), 0, stream, Mrows, p->nx, S, gate);
    } else {
        rowgemm(T, S, p->nx, p->nx, p->q_dev);
    }
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

// T[col][y][c] -= zl[col][c] * sl[y][c] + zr[col][c] * sr[y][c]: the interface corrections of the
// partition method for tridiagonal systems split over ranks (lanes along the mode index c)
__global__ void __launch_bounds__(256)
k_spike_correct(int32_t L, int32_t ny, int32_t nx, double* __restrict__ T, const double* __restrict__ zl,
                const double* __restrict__ zr, const double* __restrict__ sl, const double* __restrict__ sr) {
    const int64_t total = (int64_t)L * ny * nx;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = (int32_t)(t % nx);
        const int64_t r = t / nx;
        const int32_t y = (int32_t)(r % ny);
        const int64_t col = r / ny;
        T[t] -= zl[col * nx + c] * sl[(int64_t)y * nx + c] + zr[col * nx + c] * sr[(int64_t)y * nx + c];
    }
}

}  // namespace spuq

using namespace spuq;

extern "C" int spuq_sg_precond_mean_fd_forward(spuq_sg_precond* p, int64_t N, int32_t L, const double* R, double* T,
                                               void* work, void* stream) {
    SPUQ_REQUIRE(p && p->kind == spuq_sg_precond::MEAN_FD && R && T, "precond_mean_fd_forward: needs a mean_fd handle");
    return mean_fd_forward(p, N, L, R, T, work, static_cast<cudaStream_t>(stream));
}

extern "C" int spuq_sg_precond_mean_fd_inverse(spuq_sg_precond* p, int64_t N, int32_t L, const double* T, double* S,
                                               void* stream) {
    SPUQ_REQUIRE(p && p->kind == spuq_sg_precond::MEAN_FD && T && S, "precond_mean_fd_inverse: needs a mean_fd handle");
    return mean_fd_inverse(p, N, L, T, S, static_cast<cudaStream_t>(stream));
}

extern "C" int spuq_sg_precond_mean_fd_mode_order(const spuq_sg_precond* p, int32_t* order) {
    SPUQ_REQUIRE(p && p->kind == spuq_sg_precond::MEAN_FD && order, "precond_mean_fd_mode_order: needs a mean_fd handle");
    const int32_t h = p->half;
    for (int32_t c = 0; c < p->nx; ++c) order[c] = h > 0 ? (c < h ? 2 * c : 2 * (c - h) + 1) : c;
    return SPUQ_OK;
}

extern "C" int spuq_sg_spike_correct(int32_t L, int32_t ny, int32_t nx, double* T, const double* zl, const double* zr,
                                     const double* sl, const double* sr, void* stream) {
    SPUQ_REQUIRE(L >= 0 && ny >= 1 && nx >= 1 && T && zl && zr && sl && sr, "spike_correct: bad argument");
    const int64_t total = (int64_t)L * ny * nx;
    if (total == 0) return SPUQ_OK;
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 1 << 20);
    SPUQ_LAUNCH(k_spike_correct, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), L, ny, nx, T, zl, zr, sl, sr);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}


extern "C" int spuq_sg_precond_mean_fd(spuq_sg_precond** out, int device, int32_t nx, int32_t ny,
                                       double d, double ox, double oy) {
    SPUQ_REQUIRE(out != nullptr, "precond_mean_fd: null out");
    SPUQ_REQUIRE(nx >= 1 && ny >= 1, "precond_mean_fd: bad grid %d x %d", nx, ny);
    SPUQ_CUDA_TRY(cudaSetDevice(device));
    const double dxx = ny > 1 ? 0.5 * d : d, dyy = ny > 1 ? 0.5 * d : 0.0;
    const double pi = 3.14159265358979323846;
    std::vector<double> q((size_t)nx * nx), piv((size_t)nx * ny);
    const double scale = std::sqrt(2.0 / (nx + 1.0));
    for (int32_t j = 0; j < nx; ++j)
        for (int32_t k = 0; k < nx; ++k)
            q[(size_t)j * nx + k] = scale * std::sin((double)(j + 1) * (double)(k + 1) * pi / (nx + 1.0));
    for (int32_t k = 0; k < nx; ++k) {
        const double lam = dxx + 2.0 * ox * std::cos((double)(k + 1) * pi / (nx + 1.0));
        const double b = dyy + lam;
        double pv = b;
        for (int32_t y = 0; y < ny; ++y) {
            if (y > 0) pv = b - oy * oy / pv;
            SPUQ_REQUIRE(pv != 0.0 && std::isfinite(pv), "precond_mean_fd: zero pivot (mean block singular)");
            piv[(size_t)y * nx + k] = 1.0 / pv;
        }
    }
    spuq_sg_precond* p = new spuq_sg_precond();
    p->kind = spuq_sg_precond::MEAN_FD;
    p->device = device; p->nx = nx; p->ny = ny; p->d = d; p->ox = ox; p->oy = oy;
    void* dq = nullptr; void* dp = nullptr;
    if (cudaMalloc(&dq, sizeof(double) * q.size()) != cudaSuccess ||
        cudaMalloc(&dp, sizeof(double) * piv.size()) != cudaSuccess) {
        cudaFree(dq);
        delete p;
        set_error("precond_mean_fd: out of device memory");
        return SPUQ_E_NOMEM;
    }
    p->q_dev = static_cast<double*>(dq);
    p->piv_dev = static_cast<double*>(dp);
    if (nx % 2 == 0 && nx >= 4) {
        const int32_t h = nx / 2;
        const size_t hh = (size_t)h * h;
        std::vector<double> qs(4 * hh), pvs(piv.size());
        for (int32_t j = 0; j < h; ++j)
            for (int32_t c = 0; c < h; ++c) {
                const double e = q[(size_t)j * nx + 2 * c], o = q[(size_t)j * nx + 2 * c + 1];
                qs[(size_t)j * h + c] = e;                  // Qe   [j][c]  (modes 2c)
                qs[hh + (size_t)j * h + c] = o;             // Qo   [j][c]  (modes 2c+1)
                qs[2 * hh + (size_t)c * h + j] = e;         // Qe^T [c][j]
                qs[3 * hh + (size_t)c * h + j] = o;         // Qo^T [c][j]
            }
        for (int32_t y = 0; y < ny; ++y)
            for (int32_t c = 0; c < h; ++c) {
                pvs[(size_t)y * nx + c] = piv[(size_t)y * nx + 2 * c];
                pvs[(size_t)y * nx + h + c] = piv[(size_t)y * nx + 2 * c + 1];
            }
        void* dqs = nullptr; void* dps = nullptr;
        if (cudaMalloc(&dqs, sizeof(double) * qs.size()) == cudaSuccess &&
            cudaMalloc(&dps, sizeof(double) * pvs.size()) == cudaSuccess) {
            p->qsplit_dev = static_cast<double*>(dqs);
            p->piv_split_dev = static_cast<double*>(dps);
            cudaMemcpy(p->qsplit_dev, qs.data(), sizeof(double) * qs.size(), cudaMemcpyHostToDevice);
            cudaMemcpy(p->piv_split_dev, pvs.data(), sizeof(double) * pvs.size(), cudaMemcpyHostToDevice);
            p->half = h;
        } else {
            cudaFree(dqs);          // not fatal: the full-size transform still works
        }
    }
    cudaMemcpy(p->q_dev, q.data(), sizeof(double) * q.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p->piv_dev, piv.data(), sizeof(double) * piv.size(), cudaMemcpyHostToDevice);
    *out = p;
    return SPUQ_OK;
}
