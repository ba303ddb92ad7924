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
// The two products are plain dense fp64 GEMMs ((L*ny) x nx by nx x nx) done by k_rowgemm below, an
// in-house register-tiled kernel on the FP64 CUDA cores (no library on this path).
//
// ROUND 2: the full-length transform of every grid row is only the FALL-BACK now (and the engine of the
// forward/inverse halves used by the row-sharded solver).  The default path is the two-level strip
// partition of precond_strip.cuh, which needs the long transform for ~3 % of the rows only.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "precond.h"

#include "precond_strip.cuh"

namespace spuq {

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

__global__ void __launch_bounds__(256)
k_copy_gated(int64_t n, const double* __restrict__ src, double* __restrict__ dst, const double* gate) {
    if (gate != nullptr && __ldg(gate) != (double)SPUQ_PCG_RUNNING) return;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) dst[t] = src[t];
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

// ---- two-level strip partition: plan (host tables) and driver ---------------------------------------
struct StripPlan {
    StripView sv{};
    StripQ Q{};
    std::vector<void*> dev;               // device allocations
    int32_t cmax = 0;                     // tallest strip
    int32_t main_ct = 0;                  // height of the main strips when the kernel has a compile-time form for it
    size_t smem_bytes = 0;
    double* sy_ipiv = nullptr;            // [q][nx] separator system per x-mode (transform order of the GEMM path)
    double* sy_off = nullptr;
    double* partials = nullptr;           // fused-dot partials of the phase entry points (grown on demand)
    int64_t partials_cap = 0;
    unsigned int* ticket = nullptr;       // completion counter of the fused <R, S> (zero between launches)
    size_t schur_smem = 0;                // > 0: the separator system runs as k_schur_fused with this much shared memory
};

// first column of (tridiag(o, b, o) of size n)^{-1}
static void tri_first_column(int n, double b, double o, std::vector<double>* v) {
    std::vector<double> piv(n);
    v->assign(n, 0.0);
    piv[0] = b;
    for (int t = 1; t < n; ++t) piv[t] = b - o * o / piv[t - 1];
    // forward elimination of e_0: z_0 = 1, z_t = -o z_{t-1} / piv_{t-1}
    std::vector<double>& z = *v;
    z[0] = 1.0;
    for (int t = 1; t < n; ++t) z[t] = -o * z[t - 1] / piv[t - 1];
    z[n - 1] /= piv[n - 1];
    for (int t = n - 2; t >= 0; --t) z[t] = (z[t] - o * z[t + 1]) / piv[t];
}

// equal-as-possible pieces of n points separated by single separator points, each piece <= cap
static void partition_line(int n, int cap, std::vector<int32_t>* start, std::vector<int32_t>* len) {
    const int P = (n + 1 + cap) / (cap + 1);
    const int q = P - 1, tot = n - q, h = tot / P, r = tot % P;
    start->clear(); len->clear();
    int pos = 0;
    for (int p = 0; p < P; ++p) {
        const int c = h + (p < r ? 1 : 0);
        start->push_back(pos); len->push_back(c);
        pos += c + 1;
    }
}

// x-chunks of a strip row: chunks of `len` points at a pitch of len + 1 (the separator), the remainder in
// the last chunk.  The pitch is ODD (len = 30): lanes that sweep consecutive chunks of a row then hit
// shared-memory words 62 apart, i.e. 16 different bank pairs per half-warp -- a pitch of 32 doubles put
// all 32 lanes on ONE bank pair (measured: 16-way conflicts made the sweeps 4x the rest of the kernel).
static void partition_chunks(int n, int len, std::vector<int32_t>* start, std::vector<int32_t>* lens) {
    start->clear(); lens->clear();
    int pos = 0;
    while (n - pos > len + 1) {                // room for a chunk, its separator and at least one more point
        start->push_back(pos); lens->push_back(len);
        pos += len + 1;
    }
    if (n - pos == len + 1 && len + 1 > STRIP_CX) {      // one point too many for the sweeps' registers: two halves and a separator
        const int half = len / 2;
        start->push_back(pos); lens->push_back(half);
        pos += half + 1;
    }
    start->push_back(pos); lens->push_back(n - pos);     // last chunk: 1 .. min(len + 1, STRIP_CX) points
}

template <typename T>
static T* strip_upload(StripPlan* sp, const std::vector<T>& h) {
    void* d = nullptr;
    if (cudaMalloc(&d, sizeof(T) * std::max<size_t>(h.size(), 1)) != cudaSuccess) return nullptr;
    sp->dev.push_back(d);
    if (!h.empty()) cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
    return static_cast<T*>(d);
}

// instantiated separator counts: two transform columns per thread up to SCHUR_R2MAX rows, one beyond
#define SPUQ_SCHUR_FOR_EACH(X) X(4, 2) X(8, 2) X(12, 2) X(16, 2) X(18, 2) X(20, 2) X(22, 2) X(24, 2) X(28, 1) X(32, 1)
static int schur_rows_for(int q) {
    static const int rows[] = {4, 8, 12, 16, 18, 20, 22, 24, 28, 32};
    for (int r : rows) if (q <= r) return r;
    return 0;
}
#ifndef SPUQ_EMU
template <typename F>
static void strip_for_each_kernel(F f) {
#define SPUQ_STRIP_K(ct) f((const void*)k_strip<false, ct>); f((const void*)k_strip<true, ct>);
    SPUQ_STRIP_K(0) SPUQ_STRIP_K(8) SPUQ_STRIP_K(12) SPUQ_STRIP_K(16) SPUQ_STRIP_K(20) SPUQ_STRIP_K(24)
    SPUQ_STRIP_K(26) SPUQ_STRIP_K(28) SPUQ_STRIP_K(32)
#undef SPUQ_STRIP_K
}
template <typename F>
static void schur_for_each_kernel(F f) {
#define SPUQ_SCHUR_K(r, c) f((const void*)k_schur_fused<r, c>);
    SPUQ_SCHUR_FOR_EACH(SPUQ_SCHUR_K)
#undef SPUQ_SCHUR_K
}
#endif

void strip_plan_destroy(spuq_sg_precond* p) {
    StripPlan* sp = static_cast<StripPlan*>(p->strip);
    if (!sp) return;
    for (void* d : sp->dev) cudaFree(d);
    delete sp;
    p->strip = nullptr;
}

// Strips of `nrows` grid rows: main strips of ONE height from the set the kernel is instantiated for
// (compile-time transform), single separator rows between them, and a remainder strip of 1 .. cap rows at
// the end (generic kernel).  Tiny grids: equal-as-possible strips of any height.
static void choose_strips(int nrows, int cap, std::vector<int32_t>* ystart, std::vector<int32_t>* ylen) {
    static const int heights[] = {32, 28, 26, 24, 20, 16, 12, 8};
    ystart->clear(); ylen->clear();
    int cm = 0, nmain = 0, rem = 0;
    for (int h : heights) {
        if (h > cap) continue;
        const int k = (nrows - 1) / (h + 1), r = nrows - k * (h + 1);
        if (r >= 1 && r <= std::min(cap, STRIP_CMAX)) { cm = h; nmain = k; rem = r; break; }
    }
    if (cm == 0) { partition_line(nrows, cap, ystart, ylen); return; }
    int pos = 0;
    for (int i = 0; i < nmain; ++i) { ystart->push_back(pos); ylen->push_back(cm); pos += cm + 1; }
    ystart->push_back(pos); ylen->push_back(rem);
}

// ctx != null: the handle's grid is the band of rows [bounds[rank], bounds[rank+1]) of a grid whose rows are
// sharded over `world` ranks; the last row of every band but the last is a separator (it couples the band to
// the next one), and the separator system spans all bands.
int strip_plan_build(spuq_sg_precond* p, const StripRowsCtx* ctx) {
    const int32_t nx = p->nx, ny = p->ny;
    if (ny < 2 || nx < 2) return ctx ? SPUQ_E_UNSUPPORTED : SPUQ_OK;      // 1-D: the plain path
    if (const char* e = std::getenv("SPUQ_PRECOND_PLAIN")) { if (e[0] == '1') return SPUQ_OK; }   // A/B switch (development)
    const double pi = 3.14159265358979323846;
    const double dxx = 0.5 * p->d, dyy = 0.5 * p->d, ox = p->ox, oy = p->oy;
    // x-chunks: preferably chunks of 32 at a pitch of 33 when that makes 16 (or 32, ...) chunks per row: with the
    // rows at a pitch that is a multiple of 16 doubles, the 16 lanes of a half-warp that sweep consecutive chunks
    // then sit on 16 different bank pairs even when they straddle two rows (nx = 500, 512: 16 chunks).  Otherwise
    // chunks of 30 at a pitch of 31 (odd: no conflicts inside a row).
    std::vector<int32_t> ystart, ylen, xstart, xlen;
    partition_chunks(nx, 32, &xstart, &xlen);
    int32_t spitch = nx;
    if (xstart.size() % 16 == 0 && !std::getenv("SPUQ_STRIP_PITCH31")) spitch = (nx + 15) / 16 * 16;
    else partition_chunks(nx, 30, &xstart, &xlen);
    const int qx_ = (int)xstart.size() - 1;
    // strip height: two CTAs per SM when that still allows >= 12 rows, else one.  Per row of the strip: sp doubles
    // and its x-separator values (odd pitch, see the kernel)
    const size_t per_row = sizeof(double) * ((size_t)spitch + (size_t)(qx_ | 1));
    auto rows_for = [&](size_t budget) { return (int)((budget - 64) / per_row); };
    // 228 KB per SM, 1 KB reserved per CTA: two CTAs have 113 KB each
    int cap = std::min(STRIP_CMAX, rows_for(113 * 1024));
    if (cap < 12) cap = std::min(STRIP_CMAX, rows_for(224 * 1024));
    if (cap < 2) return SPUQ_OK;                          // grid too wide for shared memory: plain path
    StripPlan* sp = new StripPlan();
    StripView& sv = sp->sv;
    const int has_bottom = (ctx && ctx->rank < ctx->world - 1) ? 1 : 0;
    choose_strips(ny - has_bottom, cap, &ystart, &ylen);
    // all strips of all bands in global order (one band, this one, without a context)
    std::vector<int32_t> gheights;
    int32_t j0 = 0;
    if (ctx) {
        std::vector<int32_t> ys, yl;
        for (int r = 0; r < ctx->world; ++r) {
            const int nr = ctx->bounds[r + 1] - ctx->bounds[r] - (r < ctx->world - 1 ? 1 : 0);
            if (nr < 1) { delete sp; set_error("precond_mean_fd_rows: rank %d owns fewer than two grid rows", r); return SPUQ_E_INVALID; }
            choose_strips(nr, cap, &ys, &yl);
            if (r == ctx->rank) j0 = (int32_t)gheights.size();
            gheights.insert(gheights.end(), yl.begin(), yl.end());
        }
    } else {
        gheights = ylen;
    }
    sv.nx = nx; sv.ny = ny; sv.sp = spitch; sv.ld = (int64_t)nx * ny; sv.ox = ox; sv.oy = oy;
    sv.P = (int32_t)ystart.size(); sv.q = (int32_t)gheights.size() - 1;     // q: separators of the WHOLE grid
    sv.q_loc = sv.P - 1 + has_bottom; sv.has_bottom = has_bottom; sv.u_above0 = j0 - 1;
    sv.gf_ld = sv.P + has_bottom;
    sv.u_js = nx; sv.u_cs = (int64_t)sv.q * nx;          // one-GPU layout; the phase entry points set their own
    sv.PX = (int32_t)xstart.size(); sv.qx = sv.PX - 1;
    if (sv.qx > STRIP_QXMAX) { delete sp; return SPUQ_OK; }
    sv.cxmax = *std::max_element(xlen.begin(), xlen.end());
    if (sv.cxmax > STRIP_CX) { delete sp; return ctx ? SPUQ_E_UNSUPPORTED : SPUQ_OK; }
    sv.ch[0] = ylen.front(); sv.ch[1] = ylen.back();     // heights h+1 (first strips) and h
    sv.xlen_cls_len[0] = xlen.front(); sv.xlen_cls_len[1] = xlen.back();
    sp->cmax = std::max(sv.ch[0], sv.ch[1]);
    {
        static const int heights[] = {32, 28, 26, 24, 20, 16, 12, 8};
        for (int h : heights) if (sv.ch[0] == h && ylen.size() > 1) sp->main_ct = h;
        if (std::getenv("SPUQ_STRIP_GENERIC")) sp->main_ct = 0;        // A/B switch (development)
    }
    sv.n_tall = 0;
    for (size_t i = 0; i < ylen.size(); ++i) sv.n_tall += (ylen[i] == sv.ch[0]) ? 1 : 0;
    if (sv.ch[0] == sv.ch[1]) sv.n_tall = sv.P;
    std::vector<int32_t> ycls(sv.P);
    for (int i = 0; i < sv.P; ++i) ycls[i] = ylen[i] == sv.ch[0] ? 0 : 1;
    sp->ticket = strip_upload(sp, std::vector<unsigned int>(4, 0u));
    sv.ystart = strip_upload(sp, ystart); sv.ycls = strip_upload(sp, ycls);
    sv.xstart = strip_upload(sp, xstart); sv.xlen = strip_upload(sp, xlen);
    bool ok = sv.ystart && sv.ycls && sv.xstart && sv.xlen;
    const int cx = sv.cxmax, qx = sv.qx;
    for (int cls = 0; cls < 2 && ok; ++cls) {
        const int c = sv.ch[cls];
        const int hp = (c + 1) / 2;
        double* qc = sp->Q.q[cls];
        std::vector<double> ipx((size_t)c * cx), spk((size_t)2 * c * cx, 0.0),
            sip((size_t)c * std::max(qx, 1), 0.0), sof((size_t)c * std::max(qx, 1), 0.0);
        const double sc = std::sqrt(2.0 / (c + 1.0));
        for (int j = 0; j < c; ++j) {
            for (int i = 0; i < hp; ++i) qc[(size_t)j * STRIP_HP + i] = sc * std::sin((double)(i + 1) * (double)(j + 1) * pi / (c + 1.0));
            const double b = dxx + dyy + 2.0 * oy * std::cos((double)(j + 1) * pi / (c + 1.0));
            double pv = b;
            for (int t = 0; t < cx; ++t) {
                if (t > 0) pv = b - ox * ox / pv;
                if (!(pv != 0.0) || !std::isfinite(pv)) ok = false;
                ipx[(size_t)j * cx + t] = 1.0 / pv;
            }
            std::vector<double> v[2];
            for (int lc = 0; lc < 2; ++lc) {
                tri_first_column(sv.xlen_cls_len[lc], b, ox, &v[lc]);
                for (int t = 0; t < sv.xlen_cls_len[lc]; ++t) spk[((size_t)lc * c + j) * cx + t] = v[lc][t];
            }
            // x-separator system: D_i = b - ox^2 (G_ll(chunk i) + G_00(chunk i+1)), E_i = -ox^2 G_0l(chunk i+1)
            double prev_piv = 0.0, prev_off = 0.0;
            for (int i = 0; i < qx; ++i) {
                const int lcl = xlen[i] == sv.xlen_cls_len[0] ? 0 : 1, lcr = xlen[i + 1] == sv.xlen_cls_len[0] ? 0 : 1;
                const double D = b - ox * ox * (v[lcl][0] + v[lcr][0]);
                const double E = -ox * ox * v[lcr][xlen[i + 1] - 1];
                double pvt = D;
                if (i > 0) pvt = D - prev_off * prev_off / prev_piv;
                if (!(pvt != 0.0) || !std::isfinite(pvt)) ok = false;
                sip[(size_t)j * qx + i] = 1.0 / pvt;
                sof[(size_t)j * qx + i] = E;
                prev_piv = pvt; prev_off = E;
            }
        }
        sv.ipivx[cls] = strip_upload(sp, ipx); sv.spike[cls] = strip_upload(sp, spk);
        sv.sx_ipiv[cls] = strip_upload(sp, sip); sv.sx_off[cls] = strip_upload(sp, sof);
        ok = ok && sv.ipivx[cls] && sv.spike[cls] && sv.sx_ipiv[cls] && sv.sx_off[cls];
    }
    // separator rows: per x-mode k a q x q tridiagonal system, stored in the order the row transform
    // delivers the modes (even | odd when the transform is split by parity)
    if (ok && sv.q > 0) {
        const int q = sv.q, h = p->half;
        std::vector<double> ip((size_t)q * nx), of((size_t)q * nx);
        std::vector<int32_t> distinct(gheights);
        std::sort(distinct.begin(), distinct.end());
        distinct.erase(std::unique(distinct.begin(), distinct.end()), distinct.end());
        std::vector<std::vector<double>> v(STRIP_CMAX + 2);
        for (int32_t c0 = 0; c0 < nx && ok; ++c0) {
            const int32_t k = h > 0 ? (c0 < h ? 2 * c0 : 2 * (c0 - h) + 1) : c0;
            const double bb = dyy + dxx + 2.0 * ox * std::cos((double)(k + 1) * pi / (nx + 1.0));
            for (int32_t hh : distinct) tri_first_column(hh, bb, oy, &v[hh]);
            double prev_piv = 0.0, prev_off = 0.0;
            for (int j = 0; j < q; ++j) {
                const int ca = gheights[j], cb = gheights[j + 1];
                const double D = bb - oy * oy * (v[ca][0] + v[cb][0]);
                const double E = -oy * oy * v[cb][cb - 1];
                double pvt = D;
                if (j > 0) pvt = D - prev_off * prev_off / prev_piv;
                if (!(pvt != 0.0) || !std::isfinite(pvt)) ok = false;
                ip[(size_t)j * nx + c0] = 1.0 / pvt;
                of[(size_t)j * nx + c0] = E;
                prev_piv = pvt; prev_off = E;
            }
        }
        sp->sy_ipiv = strip_upload(sp, ip); sp->sy_off = strip_upload(sp, of);
        ok = ok && sp->sy_ipiv && sp->sy_off;
    }
    sv.sepv_n = (sp->cmax * (qx | 1) + 1) & ~1;                     // even: the strip starts 16-byte aligned
    sp->smem_bytes = sizeof(double) * ((size_t)sv.sepv_n + (size_t)sp->cmax * sv.sp);
    p->strip = sp;
    if (!ok) { strip_plan_destroy(p); return SPUQ_OK; }       // not fatal: the plain path still works
#ifndef SPUQ_EMU
    bool attr_ok = true;
    auto prep = [&](const void* fn) {
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp->smem_bytes) != cudaSuccess) attr_ok = false;
        // two CTAs of ~110 KB per SM: ask for the largest shared-memory carve-out
        cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    };
    strip_for_each_kernel(prep);
    if (!attr_ok) { cudaGetLastError(); strip_plan_destroy(p); return SPUQ_OK; }
    cudaGetLastError();
#endif
    // the separator system in one kernel when a column's separator rows fit shared memory
    {
        const int32_t h = p->half;
        const size_t need = sizeof(double) * (size_t)sv.q * 2 * (size_t)((h + 1) & ~1);
        const char* e = std::getenv("SPUQ_SCHUR_PLAIN");                   // A/B switch (development)
        if (h > 0 && nx == 2 * h && nx <= 512 && sv.q >= 1 && sv.q <= SCHUR_RMAX && need <= 200 * 1024 &&
            !(e && e[0] == '1')) {
            sp->schur_smem = need;
#ifndef SPUQ_EMU
            bool s_ok = true;
            schur_for_each_kernel([&](const void* fn) {
                if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need) != cudaSuccess) s_ok = false;
            });
            if (!s_ok) { cudaGetLastError(); sp->schur_smem = 0; }
#endif
        }
    }
    return SPUQ_OK;
}

// launch k_strip<FINAL, CT> for the strips [p_off, p_off + p_cnt) of all L columns
template <bool FINAL>
static void strip_launch(int ct, const StripPlan* sp, const StripView& view, int32_t L, int32_t p_off, int32_t p_cnt,
                         const double* R, double* S, double* Gf, double* Gl, const double* U, cudaStream_t stream,
                         const double* gate, const StripDot& dot) {
    if (p_cnt <= 0) return;
    const dim3 grid((unsigned)((int64_t)L * p_cnt)), block(STRIP_THREADS);
#define SPUQ_STRIP_L(c)                                                                                          \
    case c: SPUQ_LAUNCH((k_strip<FINAL, c>), grid, block, sp->smem_bytes, stream, sp->Q, view, L, p_off, p_cnt, \
                        R, S, Gf, Gl, U, gate, dot); break;
    switch (ct) {
        SPUQ_STRIP_L(8) SPUQ_STRIP_L(12) SPUQ_STRIP_L(16) SPUQ_STRIP_L(20) SPUQ_STRIP_L(24) SPUQ_STRIP_L(26)
        SPUQ_STRIP_L(28) SPUQ_STRIP_L(32)
        default: SPUQ_LAUNCH((k_strip<FINAL, 0>), grid, block, sp->smem_bytes, stream, sp->Q, view, L, p_off, p_cnt,
                             R, S, Gf, Gl, U, gate, dot);
    }
#undef SPUQ_STRIP_L
}

// one pass over all strips of all columns: ONE launch (the kernel of the main strips' height also takes the
// remainder strip through its generic branch)
template <bool FINAL>
static void strip_pass(const StripPlan* sp, const StripView& sv, int32_t L, const double* R, double* S, double* Gf, double* Gl,
                       const double* U, cudaStream_t stream, const double* gate, const StripDot& dot = StripDot{}) {
    strip_launch<FINAL>(sv.n_tall > 0 ? sp->main_ct : 0, sp, sv, L, 0, sv.P, R, S, Gf, Gl, U, stream, gate, dot);
}
template <bool FINAL>
static void strip_pass(const StripPlan* sp, int32_t L, const double* R, double* S, double* Gf, double* Gl, const double* U,
                       cudaStream_t stream, const double* gate, const StripDot& dot = StripDot{}) {
    strip_pass<FINAL>(sp, sp->sv, L, R, S, Gf, Gl, U, stream, gate, dot);
}

static int64_t strip_work_doubles(const StripPlan* sp, int32_t L) {
    // first / last rows of every strip, separator rows twice, one partial of the fused dot product per strip
    return (int64_t)L * sp->sv.nx * (2 * (int64_t)sp->sv.P + 2 * (int64_t)sp->sv.q) + (int64_t)L * sp->sv.P + 8;
}

// separator system in place on Z (layout of sv.u_js / u_cs, q rows of L columns); Z2: scratch of the same size.
// R != null: the right-hand side is assembled from R, Gf, Gl first (one-GPU strips).  Returns the buffer that
// holds the solution.
static const double* strip_schur(spuq_sg_precond* p, const StripPlan* sp, const StripView& sv, int32_t L, double* Z, double* Z2,
                                 cudaStream_t stream, const double* gate, const double* R = nullptr,
                                 const double* Gf = nullptr, const double* Gl = nullptr) {
    if (sp->schur_smem > 0) {
        const dim3 grid((unsigned)L);
#define SPUQ_SCHUR_L(r, c)                                                                                        \
    case r: SPUQ_LAUNCH((k_schur_fused<r, c>), grid, dim3(SchurShape<c>::threads), sp->schur_smem, stream, sv, L, p->half, \
                        R != nullptr ? 1 : 0, R, Gf, Gl, Z, p->qsplit_dev, sp->sy_ipiv, sp->sy_off, gate); break;
        switch (schur_rows_for(sv.q)) {
            SPUQ_SCHUR_FOR_EACH(SPUQ_SCHUR_L)
        }
#undef SPUQ_SCHUR_L
        return Z;
    }
    if (R != nullptr) {
        const unsigned eb = (unsigned)std::min<int64_t>(((int64_t)L * sv.q * sv.nx + 255) / 256, 1 << 20);
        SPUQ_LAUNCH(k_sep_assemble, dim3(eb), dim3(256), 0, stream, sv, L, R, Gf, Gl, Z, gate);
    }
    const int32_t nx = sv.nx, q = sv.q;
    const int64_t Mrows = (int64_t)L * q;
    auto rowgemm = [&](const double* src, double* dst, int32_t nk, int32_t nc, const double* Q) {
        const dim3 g((unsigned)((Mrows + GM - 1) / GM), (unsigned)((nc + GN - 1) / GN));
        SPUQ_LAUNCH(k_rowgemm, g, dim3(G_THREADS), 0, stream, Mrows, nk, nc, src, nx, Q, dst, nx, gate);
    };
    const int64_t nthreads = (int64_t)L * nx;
    const dim3 tg((unsigned)((nthreads + 127) / 128));
    const int32_t h = p->half;
    if (h > 0) {
        const size_t hh = (size_t)h * h;
        const unsigned mb = (unsigned)std::min<int64_t>((Mrows * h + 255) / 256, 1 << 20);
        SPUQ_LAUNCH(k_mirror_split, dim3(mb), dim3(256), 0, stream, Mrows, nx, Z, Z2, gate);
        rowgemm(Z2, Z, h, h, p->qsplit_dev);
        rowgemm(Z2 + h, Z + h, h, h, p->qsplit_dev + hh);
        SPUQ_LAUNCH(k_thomas_sep, tg, dim3(128), 0, stream, nx, q, L, sv.u_js, sv.u_cs, sp->sy_ipiv, sp->sy_off, Z, gate);
        rowgemm(Z, Z2, h, h, p->qsplit_dev + 2 * hh);
        rowgemm(Z + h, Z2 + h, h, h, p->qsplit_dev + 3 * hh);
        SPUQ_LAUNCH(k_mirror_merge, dim3(mb), dim3(256), 0, stream, Mrows, nx, Z2, gate);
        return Z2;
    }
    rowgemm(Z, Z2, nx, nx, p->q_dev);
    SPUQ_LAUNCH(k_thomas_sep, tg, dim3(128), 0, stream, nx, q, L, sv.u_js, sv.u_cs, sp->sy_ipiv, sp->sy_off, Z2, gate);
    rowgemm(Z2, Z, nx, nx, p->q_dev);
    return Z;
}

static int strip_apply(spuq_sg_precond* p, int64_t N, int32_t L, const double* R, double* S, void* work,
                       cudaStream_t stream, const double* gate, FusedDot* fd) {
    StripPlan* sp = static_cast<StripPlan*>(p->strip);
    const StripView& sv = sp->sv;
    const int32_t nx = sv.nx, q = sv.q;
    SPUQ_REQUIRE(!sv.has_bottom && sv.u_above0 == -1, "precond_apply: this handle is a band of a row-sharded grid; use the strip phase calls");
    SPUQ_REQUIRE(N == (int64_t)nx * sv.ny, "precond_mean_fd: slab rows %lld != nx*ny", (long long)N);
    SPUQ_REQUIRE(work != nullptr || q == 0, "precond_mean_fd: workspace missing");
    SPUQ_REQUIRE((int64_t)L * sv.P < 0x7fffffffLL, "precond_mean_fd: too many strips for one launch");
    double* Gf = static_cast<double*>(work);
    double* Gl = Gf + (int64_t)L * sv.P * nx;
    double* Z = Gl + (int64_t)L * sv.P * nx;
    double* Z2 = Z + (int64_t)L * q * nx;
    const double* U = nullptr;
    if (q > 0) {
        strip_pass<false>(sp, L, R, S, Gf, Gl, nullptr, stream, gate);
        U = strip_schur(p, sp, sv, L, Z, Z2, stream, gate, R, Gf, Gl);
    }
    StripDot dot{};
    if (fd != nullptr && sp->ticket != nullptr && work != nullptr) {
        dot.partials = Z2 + (int64_t)L * q * nx;
        dot.ticket = sp->ticket;
        dot.out = fd->out;
        dot.epi = *fd->epi;
        fd->done = true;
    }
    strip_pass<true>(sp, L, R, S, Gf, Gl, U, stream, gate, dot);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}

int mean_fd_destroy(spuq_sg_precond* p) {
    cudaSetDevice(p->device);
    cudaFree(p->q_dev);
    cudaFree(p->piv_dev);
    cudaFree(p->qsplit_dev);
    cudaFree(p->piv_split_dev);
    p->q_dev = p->piv_dev = p->qsplit_dev = p->piv_split_dev = nullptr;
    strip_plan_destroy(p);
    return SPUQ_OK;
}

int64_t mean_fd_work_bytes(const spuq_sg_precond* p, int64_t N, int32_t L) {
    // the strip path needs room for the first/last rows of every strip and the separator rows; callers
    // of the forward/inverse halves still need a whole slab
    int64_t need = (int64_t)sizeof(double) * N * L;
    if (p->strip) need = std::max<int64_t>(need, (int64_t)sizeof(double) * strip_work_doubles(static_cast<const StripPlan*>(p->strip), L));
    return need;
}

int mean_fd_apply(spuq_sg_precond* p, int64_t N, int32_t L, const double* R, double* S, void* work,
                  cudaStream_t stream, const double* gate, FusedDot* fd) {
    SPUQ_REQUIRE(N == (int64_t)p->nx * p->ny, "precond_mean_fd: slab rows %lld != nx*ny", (long long)N);
    if (p->strip) return strip_apply(p, N, L, R, S, work, stream, gate, fd);
    SPUQ_REQUIRE(work != nullptr, "precond_mean_fd: workspace missing");
    double* T = static_cast<double*>(work);
    const int64_t Mrows = (int64_t)L * p->ny;
    auto rowgemm = [&](const double* src, double* dst, int32_t nk, int32_t nc, const double* Q) {
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
    strip_plan_build(p, nullptr);
    *out = p;
    return SPUQ_OK;
}


// ---- the strip solver in phases, for grids whose rows are sharded over ranks (spuq_b200.sharding) ----------
// Layout of the separator arrays Z / U in these calls: [q_glob][L][nx] (separator-major), so that the block of
// one rank's separators is contiguous and can be gathered as it is.
static StripPlan* strip_of(spuq_sg_precond* p) {
    return (p && p->kind == spuq_sg_precond::MEAN_FD) ? static_cast<StripPlan*>(p->strip) : nullptr;
}
static StripView rows_view(const StripPlan* sp, int32_t L, int64_t ld = 0) {
    StripView v = sp->sv;
    v.u_js = (int64_t)L * v.nx; v.u_cs = v.nx;
    if (ld > 0) v.ld = ld;
    return v;
}

extern "C" int spuq_sg_precond_mean_fd_rows(spuq_sg_precond** out, int device, int32_t nx, double d, double ox, double oy,
                                            int32_t world, int32_t rank, const int32_t* bounds_host) {
    SPUQ_REQUIRE(out && bounds_host && world >= 1 && rank >= 0 && rank < world, "precond_mean_fd_rows: bad argument");
    const int32_t n_loc = bounds_host[rank + 1] - bounds_host[rank];
    SPUQ_REQUIRE(n_loc >= 2 && nx >= 2, "precond_mean_fd_rows: a band needs at least two grid rows and two columns");
    SPUQ_CUDA_TRY(cudaSetDevice(device));
    spuq_sg_precond* p = nullptr;
    // the band's own handle (transform matrices of length nx for the separator system); its plain tables are
    // those of an n_loc-row grid and are not used by the phase calls
    int rc = spuq_sg_precond_mean_fd(&p, device, nx, n_loc, d, ox, oy);
    if (rc) return rc;
    strip_plan_destroy(p);
    StripRowsCtx ctx{world, rank, bounds_host};
    rc = strip_plan_build(p, &ctx);
    if (rc || !p->strip) {
        spuq_sg_precond_destroy(p);
        if (!rc) set_error("precond_mean_fd_rows: the strip solver does not support this grid (nx = %d)", nx);
        return rc ? rc : SPUQ_E_UNSUPPORTED;
    }
    *out = p;
    return SPUQ_OK;
}

extern "C" int spuq_sg_precond_strip_info(const spuq_sg_precond* p, int32_t* n_strips_loc, int32_t* q_loc, int32_t* q_glob,
                                          int32_t* first_sep, int32_t* gf_ld) {
    const StripPlan* sp = strip_of(const_cast<spuq_sg_precond*>(p));
    SPUQ_REQUIRE(sp, "precond_strip_info: the handle has no strip plan");
    if (n_strips_loc) *n_strips_loc = sp->sv.P;
    if (q_loc) *q_loc = sp->sv.q_loc;
    if (q_glob) *q_glob = sp->sv.q;
    if (first_sep) *first_sep = sp->sv.u_above0 + 1;
    if (gf_ld) *gf_ld = sp->sv.gf_ld;
    return SPUQ_OK;
}

// Gf / Gl [L][gf_ld][nx] <- first / last row of the local solution of every local strip.
// ld (here and below): column stride of R / S in doubles, 0 = nx * (rows of the band); R / S point at the
// band's first row of column 0, so a band inside a halo-padded slab needs no copy.
extern "C" int spuq_sg_precond_strip_forward(spuq_sg_precond* p, int32_t L, const double* R, int64_t ld, double* Gf, double* Gl,
                                             void* stream) {
    StripPlan* sp = strip_of(p);
    SPUQ_REQUIRE(sp && R && Gf && Gl && L >= 0, "precond_strip_forward: bad argument");
    SPUQ_REQUIRE(ld == 0 || ld >= (int64_t)sp->sv.nx * sp->sv.ny, "precond_strip_forward: column stride shorter than the band");
    const StripView v = rows_view(sp, L, ld);
    strip_pass<false>(sp, v, L, R, nullptr, Gf, Gl, nullptr, static_cast<cudaStream_t>(stream), abi_gate());
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}
// rows first_sep .. first_sep + q_loc - 1 of Z <- right-hand sides of the local separators (slot gf_ld - 1 of Gf
// holds the first strip of the next rank when the band ends in a separator)
extern "C" int spuq_sg_precond_strip_assemble(spuq_sg_precond* p, int32_t L, const double* R, int64_t ld, const double* Gf,
                                              const double* Gl, double* Z, void* stream) {
    StripPlan* sp = strip_of(p);
    SPUQ_REQUIRE(sp && R && Gf && Gl && Z && L >= 0, "precond_strip_assemble: bad argument");
    SPUQ_REQUIRE(ld == 0 || ld >= (int64_t)sp->sv.nx * sp->sv.ny, "precond_strip_assemble: column stride shorter than the band");
    const StripView v = rows_view(sp, L, ld);
    if (v.q_loc == 0 || L == 0) return SPUQ_OK;
    const unsigned eb = (unsigned)std::min<int64_t>(((int64_t)L * v.q_loc * v.nx + 255) / 256, 1 << 20);
    SPUQ_LAUNCH(k_sep_assemble, dim3(eb), dim3(256), 0, static_cast<cudaStream_t>(stream), v, L, R, Gf, Gl, Z, abi_gate());
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}
// the separator system of the whole grid on Z [q_glob][L][nx]; the solution is left in Z (Z2: scratch, same size)
extern "C" int spuq_sg_precond_strip_schur(spuq_sg_precond* p, int32_t L, double* Z, double* Z2, void* stream) {
    StripPlan* sp = strip_of(p);
    SPUQ_REQUIRE(sp && Z && Z2 && L >= 0, "precond_strip_schur: bad argument");
    const StripView v = rows_view(sp, L);
    if (v.q == 0 || L == 0) return SPUQ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double* U = strip_schur(p, sp, v, L, Z, Z2, st, abi_gate());
    if (U != Z) {
        const int64_t n = (int64_t)v.q * L * v.nx;
        SPUQ_LAUNCH(k_copy_gated, dim3((unsigned)std::min<int64_t>((n + 255) / 256, 1 << 20)), dim3(256), 0, st, n, U, Z, abi_gate());
    }
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}
// S <- solution on the band's rows, given the separator values U [q_glob][L][nx] of the whole grid.
// dot_out != null: dot_out[0] <- <R, S> over the band (device scalar), produced by the same kernel.
extern "C" int spuq_sg_precond_strip_final(spuq_sg_precond* p, int32_t L, const double* R, const double* U, double* S, int64_t ld,
                                           double* dot_out, void* stream) {
    StripPlan* sp = strip_of(p);
    SPUQ_REQUIRE(sp && R && S && L >= 0 && (U || sp->sv.q == 0), "precond_strip_final: bad argument");
    SPUQ_REQUIRE(ld == 0 || ld >= (int64_t)sp->sv.nx * sp->sv.ny, "precond_strip_final: column stride shorter than the band");
    const StripView v = rows_view(sp, L, ld);
    StripDot dot{};
    if (dot_out != nullptr) {
        const int64_t need = (int64_t)L * v.P + 8;
        if (sp->partials_cap < need) {                       // grows once per column count
            SPUQ_CUDA_TRY(cudaSetDevice(p->device));
            void* d = nullptr;
            SPUQ_CUDA_TRY(cudaMalloc(&d, sizeof(double) * need));
            sp->dev.push_back(d);                            // the old buffer (if any) is freed with the plan: a launch may still read it
            sp->partials = static_cast<double*>(d);
            sp->partials_cap = need;
        }
        dot.partials = sp->partials;
        dot.ticket = sp->ticket;
        dot.out = dot_out;
        dot.epi = DotEpilogue{DOT_EPI_NONE, nullptr, 0, nullptr, 0};
    }
    strip_pass<true>(sp, v, L, R, S, nullptr, nullptr, U, static_cast<cudaStream_t>(stream), abi_gate(), dot);
    SPUQ_CUDA_TRY(cudaGetLastError());
    return SPUQ_OK;
}
