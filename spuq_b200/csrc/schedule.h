// schedule.h -- work schedule of the TMEM-accumulator apply kernel (kernels_tmem.cuh).
//
// The SG operator  Y[:, mu] = sum_t coef_t * A_{slot_t} * W[:, col_t]  (multi_operator2.py:38-84) is
// a sparse contraction over three operands: a block matrix, an input column and an output column.
// Per SM the scarce resource is shared-memory bandwidth (128 B/clk against 64 DFMA/clk), so the
// schedule keeps BOTH stencil operands in registers and the accumulators in tensor memory:
//
//   block   = 63 (one consumer group) or 31 (two groups) output columns of Y whose accumulators
//             (4 grid points per thread) fill the group's TMEM columns, plus one scratch accumulator;
//   A-set   = up to K block matrices of that block whose tiles sit in registers at the same time;
//   job     = one input column: its window is read from shared memory ONCE and used by every
//             (matrix of the set, output column of the block) pair that needs it; jobs with the same
//             matrices and disjoint outputs are processed two at a time (window buffers A and B);
//   op      = one term: optional stencil evaluation S = A_k * window, then acc[out] += coef * S.
//
// The host flattens this into a stream of 16-byte records per block, cut into 2 KB chunks that
// the producer thread copies into a shared-memory ring ahead of the consumers, plus a compact
// command list for the producer itself (which window / A tile / chunk to fetch next).
#pragma once
#include <cstdint>
#include <vector>

namespace spuq {

struct SchedRec {           // one 16-byte unit of the record stream
    int32_t a, b;
    double c;
};
static_assert(sizeof(SchedRec) == 16, "unit must be 16 bytes");

// Stream of a block (all units 16 bytes, read by the consumer warps from the metadata ring):
//
//   SET   two header units, then the jobs of one register set:
//           unit 0   a = REC_ASET | cnt << 4: load the next cnt A tiles of the A ring into register
//                    sets 0..cnt-1;  b = number of single jobs with S = 3
//           unit 1   four counts: job PAIRS with S = 1, pairs with S = 2, single jobs with S = 1,
//                    single jobs with S = 2
//         followed, in exactly that order, by the pairs (S = 1, then S = 2) and the single jobs
//         (S = 1, 2, 3).  S = accumulation slots per matrix.  A job is 1 + K*ceil(S/2) units:
//           unit 0     a = REC_JOB | kmask << 4 (matrices of the set the job uses), words 1..3: for
//                      register set kk the TMEM column offsets (accumulator number * 8) of its S
//                      slots, 9 bits each
//           unit 1+..  for every kk (used or not) ceil(S/2) units holding the slot coefficients
//                      (two doubles per unit)
//         A slot that is not needed points at the scratch accumulator (number block_cols) with
//         coefficient 0, so a job is branch-free in the data: window -> registers, then for every
//         matrix of kmask the stencil (20 DFMA) and S read-modify-writes of TMEM.  A PAIR is two jobs
//         with the same kmask and disjoint output columns, laid out as
//           [unit 0 of job a][unit 0 of job b][coef units of a][coef units of b]
//         the consumer evaluates both windows in one basic block (eight independent DFMA chains,
//         2S TMEM read-modify-writes in flight).  A job (a pair) never straddles a chunk.
//   STOREALL n                   write all accumulators to Y: n units follow, each naming the output
//                                columns of four consecutive accumulators (-1 = none)
//   NEXT_CHUNK                   continue in the next ring stage (rest of this chunk is unused); may
//                                stand wherever a unit 0 of a job or a top-level unit is expected
//   END_ITEM
//
// Each SM sub-partition runs at most two consumer warps (TMEM lane ownership, registers), which
// exposes every branch (~10 clk) and every barrier round trip; hence the regular, padded job format
// with counted runs instead of one record per term.
enum : int32_t {
    REC_END_ITEM = 0,
    REC_ASET = 2,
    REC_NEXT_CHUNK = 5,
    REC_JOB = 7,
    REC_STOREALL = 9,         // b = number of accumulator quadruples; followed by that many units of 4 output columns
    REC_JOB_KMASK_SHIFT = 4,  // JOB unit 0: matrices used by the job in bits 4..7
    REC_TYPE_MASK = 0xF,
    REC_CNT_SHIFT = 4,        // ASET: count in bits 4..7
};

constexpr int SCHED_CHUNK_RECS = 128;            // 2 KB
constexpr int SCHED_CHUNK_USABLE = 127;          // the last unit of a chunk is always NEXT_CHUNK / END
// A block holds block_cols output columns + one scratch accumulator for padded slots: 63 + 1 when one
// consumer group owns all 512 TMEM columns, 31 + 1 when two groups share them.
constexpr int SCHED_ACC_STRIDE = 8;              // TMEM columns per accumulator (4 doubles)
constexpr int SCHED_MAX_K = 3;                   // register sets (limited by the job unit layout)
constexpr int SCHED_MAX_S = 3;

enum : int32_t {                                  // producer commands
    PCMD_WINDOW = 0,        // low bits: input column
    PCMD_ATILE = 1 << 30,   // low bits: slot
    PCMD_META = 2 << 30,    // low bits: absolute chunk index
    PCMD_KIND_MASK = 3 << 30,
    PCMD_VAL_MASK = (1 << 30) - 1,
    PCMD_WINDOW_PAD = PCMD_VAL_MASK,   // WINDOW with this value: no fetch, just complete the window group
};

struct Schedule {
    int32_t K = 0;                         // register sets
    int32_t block_cols = 63;               // output columns per block; the scratch accumulator is number block_cols
    int32_t w_group = 1;                   // windows per W-ring stage: groups are padded at A-set and block ends
    int32_t nblk = 0;
    std::vector<SchedRec> recs;            // all chunks, SCHED_CHUNK_RECS records each
    std::vector<int32_t> pcmd;             // producer commands
    std::vector<int32_t> blk_pcmd;         // nblk+1: command range of each block
    std::vector<int32_t> blk_cols;         // nblk*64 output columns (-1 = padding), for reporting
    // statistics (reported through plan_info / used by DESIGN.md numbers)
    int64_t n_jobs = 0, n_pairs = 0, n_asets = 0, n_entries = 0, n_stencils = 0;
    int64_t n_inputs = 0, n_jobs_bound = 0;   // (input column, block) incidences; sum of ceil(matrices of an incidence / K)
    int64_t hist_pairs[3][8] = {}, hist_singles[3][8] = {};     // [S-1][kmask]
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
void schedule_column_order(int32_t L, int32_t M, const int32_t* nbr_up, const int32_t* nbr_dn, int32_t block_cols,
                           std::vector<int32_t>* perm);

// Build the schedule for K register sets.  `order` = column order from schedule_column_order.
void schedule_build(const TermLists& terms, const std::vector<int32_t>& order, int32_t K, int32_t n_slots,
                    int32_t block_cols, int32_t w_group, Schedule* out);

// Host interpreter of a schedule: executes the record stream with plain loops on diagonal storage
// (dia[(slot*D + d)*N + i], offsets off[d]).  Used by the CPU emulation build and by the tests to
// validate the tables without a GPU; the GPU kernel consumes exactly the same records.
void schedule_run_host(const Schedule& s, int32_t D, const int32_t* dx, const int32_t* dy, int32_t nx, int32_t ny,
                       const double* dia, const double* W, int64_t ldw, double* Y, int64_t ldy);

}  // namespace spuq
