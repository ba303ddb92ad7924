// common.h -- shared declarations of libspuq_sg (error handling, launch macro, plan layout).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>

#ifdef SPUQ_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define SPUQ_LAUNCH(kernel, grid, block, smem, stream, ...) \
    do { ++spuq::launch_counter(); kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); } while (0)
#endif

#include "../../include/spuq_sg.h"

namespace spuq {

// kernels launched through SPUQ_LAUNCH by the calling thread (reporting: gpu_launches of bench.py)
int& launch_counter();

void set_error(const char* fmt, ...);

#define SPUQ_CUDA_TRY(expr)                                                               \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            spuq::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),      \
                            __FILE__, __LINE__);                                          \
            return SPUQ_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define SPUQ_REQUIRE(cond, ...)                                                           \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            spuq::set_error(__VA_ARGS__);                                                 \
            return SPUQ_E_INVALID;                                                        \
        }                                                                                 \
    } while (0)

int sm_count(int device);

// Conditional-execution scope of the C ABI (spuq_sg_set_gate): while a gate word is set for the calling
// thread, every kernel launched through an ABI entry point becomes a no-op once *gate != SPUQ_PCG_RUNNING.
// This is how a host-enqueued solver loop (spuq_b200.sharding.RowShardedPCG) stays free of host round
// trips: it enqueues whole passes and the device decides whether they still do anything.
const double* abi_gate();

// ---- device-side views handed to the kernels (plain structs, passed by value) ---------------

// Per-column term lists: column mu of Y is  sum_t coef[t] * A_{mat[t]} * W[:, col[t]]
// for t in [ptr[mu], ptr[mu+1]).  mat = 0 is the mean block, 1+m the m-th expansion block.
struct TermView {
    const int32_t* ptr;    // L+1
    const int32_t* mat;    // T
    const int32_t* col;    // T
    const double* coef;    // T
};

// Diagonal storage of a 9-point-window pattern: offsets off[d] = dx + s*dy with dx,dy in {-1,0,1};
// values of block b on diagonal d for row i at vals[(b*D + d)*N + i] (0 where the CSR pattern has
// no entry).  Block slots b index the blocks this shard owns (slot_of_mat maps mat -> slot).
struct StencilView {
    int32_t D;
    int32_t s;             // row stride of the grid (nx)
    int64_t N;
    int32_t off[9];
    int32_t dx[9], dy[9];
    const double* vals;
};

// Column blocks for the register-tiled stencil kernel: block k owns columns
// bcol[k*C + c], c < C (-1 = padding), and its work is a list of groups; group g applies ONE
// block matrix (gmat[g], already a slot index) to gcnt[g] terms starting at gterm[g]; term t
// reads column tcol[t] and accumulates coef tcoef[t] times the stencil result into accumulator
// tacc[t] (< C).
struct BlockView {
    int32_t C;
    int32_t nblk;
    const int32_t* bcol;    // nblk*C
    const int32_t* bgrp;    // nblk+1  (group range of each block)
    const int32_t* gmat;    // G
    const int32_t* gterm;   // G+1
    const int32_t* tcol;    // T
    const int32_t* tacc;    // T
    const double* tcoef;    // T
    const void* trec;       // T packed 16-byte records {int32 col, int32 acc, double coef} (general sparse kernel)
};

}  // namespace spuq

struct spuq_sg_plan;
namespace spuq {
// apply with an optional PCG gate word (see kernels_apply.cuh); the C ABI passes gate = null
int apply_impl(spuq_sg_plan* p, const double* W, int64_t ldw, double* Y, int64_t ldy,
               cudaStream_t stream, const double* gate);
}

struct spuq_sg_plan {
    int device = 0;
    int64_t n_rows = 0, n_cols = 0, nnz = 0;
    int32_t L = 0, M = 0, m_begin = 0, m_end = 0;
    int include_mean = 1;
    int32_t path = SPUQ_PATH_CSR;
    int32_t max_row_nnz = 0;
    int n_sm = 148;

    // borrowed CSR (device) + device array of the M+1 value pointers
    const int32_t* rowptr = nullptr;
    const int32_t* colidx = nullptr;
    const double** vals_ptrs_dev = nullptr;
    std::vector<const double*> vals_host;
    std::vector<int32_t> slot_of_mat;     // mat -> compact slot among owned blocks (-1 = not owned)
    int32_t* slot_dev = nullptr;          // device copy of slot_of_mat
    int32_t n_slots = 0;

    // per-column term lists
    int64_t n_terms = 0;
    int32_t* term_ptr = nullptr;
    int32_t* term_mat = nullptr;
    int32_t* term_col = nullptr;
    double* term_coef = nullptr;
    std::vector<int32_t> h_term_ptr, h_term_mat, h_term_col;
    std::vector<double> h_term_coef;

    // ELL copy
    int32_t ell_K = 0;
    int32_t* ell_cols = nullptr;      // K * n_rows
    double* ell_vals = nullptr;       // n_slots * K * n_rows

    // stencil copy
    spuq::StencilView st{};
    double* dia_vals = nullptr;       // n_slots * D * N

    // column blocks for the tiled stencil kernel (built lazily for the C in use)
    spuq::BlockView bv{};
    int32_t ny = 0;                   // grid rows of the stencil path (N = s * ny)
    bool nine = false;
    std::vector<void*> block_dev;     // device arrays of the current BlockView
    std::vector<void*> owned_dev;     // everything cudaMalloc'ed by the plan

    // TMA descriptors of the pipelined kernel (opaque 128-byte CUtensorMap blobs)
    alignas(64) unsigned char tmap_w[128];
    alignas(64) unsigned char tmap_a[128];
    const void* tmap_w_ptr = nullptr;     // slab / ld / box the W map was encoded for
    int64_t tmap_w_ld = 0;
    int32_t tmap_w_rows = 0, tmap_a_rows = 0;

    // TMEM-accumulator kernel: schedule tables (host copy + device arrays) and the work counter
    std::vector<int32_t> h_nbr_up, h_nbr_dn;          // M x L, kept for the column order
    void* sched_host = nullptr;                        // spuq::Schedule*
    int32_t sched_K = 0, sched_C = 0;
    const void* d_recs = nullptr;
    const int32_t* d_pcmd = nullptr;
    const int32_t* d_blk_pcmd = nullptr;
    unsigned long long* d_work_counter = nullptr;
    std::vector<void*> sched_dev;                      // device arrays of the current schedule

    // spuq_sg_apply_part: 0 = whole grid, 1 = tile rows that read no halo grid row, 2 = the others
    int32_t part = 0;

    // tunables
    int32_t variant = 0, cols_per_block = 0, rows_per_thread = 0, ctas_per_sm = 0;

    // reporting
    int last_launches = 0;
    std::string kernel_name;
    int64_t alg_bytes = 0, alg_flops = 0;
};
