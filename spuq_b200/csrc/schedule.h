// schedule.h -- work schedule of the TMEM-accumulator apply kernel (kernels_tmem.cuh).
//
// The SG operator  Y[:, mu] = sum_t coef_t * A_{slot_t} * W[:, col_t]  (multi_operator2.py:38-84) is
// a sparse contraction over three operands: a block matrix, an input column and an output column.
// Per SM the scarce resource is shared-memory bandwidth (128 B/clk against 64 DFMA/clk), so the
// schedule keeps BOTH stencil operands in registers and the accumulators in tensor memory:
//
//   block   = up to 64 output columns of Y whose accumulators (4 grid points per thread) fill the
//             512 TMEM columns of a CTA;
//   A-set   = up to K block matrices of that block whose tiles sit in registers at the same time;
//   job     = one input column: its window is read from shared memory ONCE and used by every
//             (matrix of the set, output column of the block) pair that needs it; jobs are processed
//             two at a time (window buffers A and B);
//   op      = one term: optional stencil evaluation S = A_k * window, then acc[out] += coef * S.
//
// The host flattens this into a stream of 16-byte records per block, cut into 512-byte chunks that
// the producer thread copies into a shared-memory ring ahead of the consumers, plus a compact
// command list for the producer itself (which window / A tile / chunk to fetch next).
#pragma once
#include <cstdint>
#include <vector>

namespace spuq {

struct SchedRec {
    int32_t a;      // type | flags (see below)
    int32_t b;      // OP/STORE: TMEM column offset of the accumulator
    double c;       // OP: coefficient; STORE: two int32 output columns (bit pattern)
};
static_assert(sizeof(SchedRec) == 16, "record must be 16 bytes");

// The consumer warps hold TWO windows (buffers A and B) and K block-matrix tiles in registers.  An
// OP record optionally evaluates the stencil of matrix k on buffer A and/or B in one basic block
// (8 independent DFMA chains: a single warp per SM sub-partition needs that much instruction-level
// parallelism to keep the FP64 pipe busy, DFMA latency being 8 clk at 2 clk issue), then adds
// coef * S_A or coef * S_B to one TMEM accumulator.
enum : int32_t {
    REC_END_ITEM = 0,     // end of the block's stream
    REC_OP = 1,           // [stencil(s)] + one accumulation
    REC_ASET = 2,         // load the next `count` A tiles of the A ring into register sets 0..count-1
    REC_RELOAD = 3,       // load the next window(s) of the W ring into buffer A and/or B
    REC_STORE = 4,        // write two accumulators to Y
    REC_NEXT_CHUNK = 5,   // last record of a 512-byte chunk: continue in the next ring stage
    REC_TYPE_MASK = 0xF,
    REC_F_STA = 1 << 4,       // OP: S_A = A_k * window A first
    REC_F_STB = 1 << 5,       // OP: S_B = A_k * window B first
    REC_K_SHIFT = 6,          // OP: k in bits 6..7
    REC_F_SRCB = 1 << 8,      // OP: accumulate S_B (else S_A)
    REC_F_INIT = 1 << 9,      // OP: first touch of the accumulator (store, do not load)
    REC_F_NOACC = 1 << 10,    // OP: stencil only
    REC_CNT_SHIFT = 4,        // ASET: count in bits 4..7
    REC_F_BUFA = 1 << 4,      // RELOAD: buffer A
    REC_F_BUFB = 1 << 5,      // RELOAD: buffer B
};

constexpr int SCHED_CHUNK_RECS = 32;             // 512 bytes
constexpr int SCHED_BLOCK_COLS = 64;             // output columns per block
constexpr int SCHED_ACC_STRIDE = 8;              // TMEM columns per accumulator (4 doubles)

enum : int32_t {                                  // producer commands
    PCMD_WINDOW = 0,        // low bits: input column
    PCMD_ATILE = 1 << 30,   // low bits: slot
    PCMD_META = 2 << 30,    // low bits: absolute chunk index
    PCMD_KIND_MASK = 3 << 30,
    PCMD_VAL_MASK = (1 << 30) - 1,
};

struct Schedule {
    int32_t K = 0;                         // register sets
    int32_t nblk = 0;
    std::vector<SchedRec> recs;            // all chunks, SCHED_CHUNK_RECS records each
    std::vector<int32_t> pcmd;             // producer commands
    std::vector<int32_t> blk_pcmd;         // nblk+1: command range of each block
    std::vector<int32_t> blk_cols;         // nblk*64 output columns (-1 = padding), for reporting
    // statistics (reported through plan_info / used by DESIGN.md numbers)
    int64_t n_jobs = 0, n_asets = 0, n_entries = 0, n_stencils = 0;
};

// Terms of the operator in per-output-column lists (slot = compact index of an owned block matrix).
struct TermLists {
    int32_t L = 0;
    const int32_t* ptr = nullptr;     // L+1
    const int32_t* slot = nullptr;    // T   (already mapped mat -> slot)
    const int32_t* col = nullptr;     // T
    const double* coef = nullptr;     // T
};

// Column order used to cut Lambda into blocks: reconstructs the multi-indices from the neighbour
// tables (every index is its down-neighbour plus e_m) and sorts them with the dimension that
// carries the most neighbour pairs varying fastest, so that a block holds indices with a common
// "light" tail.  Falls back to the identity when tables are absent.  perm[k] = column at position k.
void schedule_column_order(int32_t L, int32_t M, const int32_t* nbr_up, const int32_t* nbr_dn,
                           std::vector<int32_t>* perm);

// Build the schedule for K register sets.  `order` = column order from schedule_column_order.
void schedule_build(const TermLists& terms, const std::vector<int32_t>& order, int32_t K, int32_t n_slots,
                    Schedule* out);

// Host interpreter of a schedule: executes the record stream with plain loops on diagonal storage
// (dia[(slot*D + d)*N + i], offsets off[d]).  Used by the CPU emulation build and by the tests to
// validate the tables without a GPU; the GPU kernel consumes exactly the same records.
void schedule_run_host(const Schedule& s, int32_t D, const int32_t* dx, const int32_t* dy, int32_t nx, int32_t ny,
                       const double* dia, const double* W, int64_t ldw, double* Y, int64_t ldy);

}  // namespace spuq
