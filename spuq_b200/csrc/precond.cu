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
// The two products are plain dense fp64 GEMMs ((L*ny) x nx by nx x nx): they go to cuBLAS DGEMM when
// libcublas can be loaded at run time (dlopen, no link-time dependency); k_rowgemm below is the
// in-house register-tiled fallback (CUDA cores) and the reference for the host-emulated build.
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
// C[M x n] = A[M x n] * Q[n x n] (row-major, Q symmetric)  ==  column-major  C^T = Q * A^T
static bool blas_rowgemm(spuq_sg_precond* p, int64_t Mrows, int32_t n, const double* A, const double* Q, double* C,
                         cudaStream_t stream) {
    CublasApi& api = cublas_api();
    if (!api.ok || Mrows > 0x7fffffffLL) return false;
    if (!p->blas && api.create(&p->blas) != 0) { p->blas = nullptr; return false; }
    if (api.set_stream(p->blas, stream) != 0) return false;
    const double one = 1.0, zero = 0.0;
    return api.dgemm(p->blas, /*CUBLAS_OP_N*/ 0, 0, n, (int)Mrows, n, &one, Q, n, A, n, &zero, C, n) == 0;
}
#endif

constexpr int GM = 128, GN = 64, GK = 8;      // CTA tile; 256 threads, each 8 x 4 outputs
constexpr int G_THREADS = 256;

#ifndef SPUQ_EMU
// C[M x n] = A[M x n] * Q[n x n]; A, C row-major with leading dimension n; Q symmetric.
__global__ void __launch_bounds__(G_THREADS)
k_rowgemm(int64_t Mrows, int32_t n, const double* __restrict__ A, const double* __restrict__ Q,
          double* __restrict__ C, const double* gate) {
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
            ra[q] = (r < Mrows && k < n) ? A[r * n + k] : 0.0;
        }
        const int k = k0 + b_k;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int c = col0 + b_c + q;
            rb[q] = (k < n && c < n) ? Q[(int64_t)k * n + c] : 0.0;
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
    for (int k0 = 0; k0 < n; k0 += GK) {
        const bool more = k0 + GK < n;
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
            if (c < n) C[r * n + c] = acc[i][j];
        }
    }
}
#else
void k_rowgemm(int64_t Mrows, int32_t n, const double* A, const double* Q, double* C, const double* gate) {
    if (gate != nullptr && *gate != (double)SPUQ_PCG_RUNNING) return;
    if (threadIdx.x != 0 || blockIdx.x != 0 || blockIdx.y != 0) return;
    for (int64_t r = 0; r < Mrows; ++r)
        for (int c = 0; c < n; ++c) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += A[r * n + k] * Q[(int64_t)k * n + c];
            C[r * n + c] = s;
        }
}
#endif

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
    p->q_dev = p->piv_dev = nullptr;
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
    const dim3 grid((unsigned)((Mrows + GM - 1) / GM), (unsigned)((p->nx + GN - 1) / GN));
    auto rowgemm = [&](const double* src, double* dst) {
#ifndef SPUQ_EMU
        if (blas_rowgemm(p, Mrows, p->nx, src, p->q_dev, dst, stream)) return;
#endif
        SPUQ_LAUNCH(k_rowgemm, grid, dim3(G_THREADS), 0, stream, Mrows, p->nx, src, p->q_dev, dst, gate);
    };
    rowgemm(R, T);
    const int64_t nthreads = (int64_t)L * p->nx;
    SPUQ_LAUNCH(k_thomas_y, dim3((unsigned)((nthreads + 127) / 128)), dim3(128), 0, stream, p->nx, p->ny, L,
                p->oy, p->piv_dev, T, gate);
    rowgemm(T, S);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

}  // namespace spuq

using namespace spuq;

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
    cudaMemcpy(p->q_dev, q.data(), sizeof(double) * q.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p->piv_dev, piv.data(), sizeof(double) * piv.size(), cudaMemcpyHostToDevice);
    *out = p;
    return SPUQ_OK;
}
