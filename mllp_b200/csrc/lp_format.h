// lp_format.h -- device-resident sparse format of A and A' for the fused PDHG kernels.
//
// Format ("warp-tiled SELL"): the rows of a matrix are re-ordered internally and cut into
// TILES, each processed by one warp.  A tile covers 32/L consecutive (internal order) rows
// with L lanes per row (L = 1,2,4,...,32: thread-per-row SELL-32 for the short rows of A',
// warp-per-row for long rows), padded to `nsteps` steps of 2 entries per lane.  Storage is
// step-major so that one step of one warp is a single contiguous 512 B run of values
// (32 x double2, 128-bit loads) and a 256 B run of indices (32 x int2); within a step the
// entries are arranged so that ONE gather instruction covers L consecutive entries of a row
// (entry e of a step sits in lane e % L, half e / L).  Rows longer than 64*max_steps entries
// are split into chunks (one tile each, L = 32).  The chunks of all split rows form one
// sequence that is dealt to the CTAs in contiguous runs; a CTA first adds up its own chunks of
// a row in shared memory (fixed order), publishes ONE partial per (CTA, row) to a scratch
// array, and the last CTA to arrive sums the row's partials in a fixed order.  Results are
// deterministic and no fp64 atomics are used.
//
// Tiles are dealt to the CTAs of the persistent grid at build time (cost-balanced) and the
// storage is CTA-major, so a CTA's share of the matrix is one contiguous range that it can
// keep resident in shared memory across iterations.
#pragma once
#include <cstdint>
#include <vector>

namespace mllp {

struct Tile {            // 16 bytes, loaded as one int4
    uint32_t off;        // first warp-step of this tile (64 entries per warp-step)
    uint32_t row_base;   // first internal row; for a split chunk: chunk index within its row
    uint16_t nsteps;     // steps of 2 entries per lane
    uint8_t logL;        // log2(lanes per row)
    uint8_t nrows;       // valid rows in this tile (1 .. 32 >> logL)
    int32_t split;       // -1, or (chunk of a split row) the CTA-local partial slot it fills
};
static_assert(sizeof(Tile) == 16, "Tile must be 16 bytes");

struct SplitRow {        // 16 bytes
    uint32_t row;        // internal row id
    uint32_t first_slot; // first global partial slot of this row
    uint32_t nparts;     // number of CTAs that contribute a partial
    uint32_t pad;
};

struct LocalSplit {      // 16 bytes: one split row as seen by one CTA
    uint32_t split_id;   // index into the split-row table
    uint32_t gslot;      // global partial slot this CTA publishes to
    uint16_t first;      // first CTA-local slot
    uint16_t count;      // number of CTA-local slots (chunks of this row in this CTA)
    uint32_t finisher;   // 1 on the CTA that holds the row's last chunk (it finishes the row in the
                         // cooperative kernels' polled join), else 0
};

constexpr int SPLIT_SLOTS = 256;  // max split chunks per CTA (shared-memory partial slots)

struct BuildParams {
    int num_ctas = 148;      // CTAs of the persistent grid
    int pref_steps = 4;      // choose L so a row needs at most this many steps
    int max_steps = 4;       // rows longer than 64*max_steps entries are split
    bool cluster = true;     // cluster rows inside a length class by the entries they touch
    int cluster_rounds = 3;  // alternate column / row clustering this many times
    bool contiguous = false; // deal regular tiles to CTAs in contiguous runs (L1 reuse) instead of least-loaded first
    bool final_params = false; // max_steps already adjusted (effective_params is then the identity)
};

// Measured correction of the dealing of the regular tiles (the tuning rounds of mllp_lp_create: per-CTA
// phase times of a traced run feed the next build).  The summation order inside every row is untouched by
// the dealing, so results do not depend on it.
struct DealFeedback {
    std::vector<double> tile_w;    // contiguous dealing: factor on the cost of each regular tile (build order)
    std::vector<double> cta_f;     // contiguous dealing: factor on each CTA's split-chunk cost
    std::vector<double> cta_bias;  // least-loaded dealing: initial load of each CTA (load units)
};

// max_steps raised until no CTA gets more than SPLIT_SLOTS split chunks (idempotent).
BuildParams effective_params(int nrows, const int32_t* ptr, BuildParams bp);

// Host-side image of one matrix in the tiled format.
struct HostMat {
    int nrows = 0, ncols = 0;
    int64_t nnz = 0;
    int64_t nnz_emitted = 0;           // nonzeros of the rows actually emitted (row partition)
    std::vector<double> vals;          // 64 * total_steps
    std::vector<int32_t> idx;          // 64 * total_steps (internal column ids)
    std::vector<Tile> tiles;           // CTA-major
    std::vector<uint32_t> cta_begin;   // num_ctas + 1 (tile index)
    std::vector<uint32_t> cta_step_begin; // num_ctas + 1 (warp-step index)
    std::vector<SplitRow> splits;
    std::vector<LocalSplit> lsplits;       // CTA-major
    std::vector<uint32_t> cta_lsplit_begin;// num_ctas + 1
    std::vector<uint32_t> cta_nsplit;      // num_ctas: leading split-chunk tiles of each CTA
    uint32_t num_partials = 0;
    uint64_t total_steps = 0;
    int max_cta_steps = 0;             // largest per-CTA warp-step count
    int max_cta_tiles = 0;
    int max_cta_rows = 0;              // largest per-CTA count of rows in regular tiles
    // host-only bookkeeping for the tuning rounds
    std::vector<uint32_t> reg_cta;     // CTA of every regular tile, in build order
    std::vector<double> cta_load;      // modelled load of every CTA (units of the dealing that was used)
};

// Internal ordering of the rows of a CSR matrix, derived from row lengths only.
// order[k] = original row at internal position k; pos[r] = internal position of row r.
void plan_row_order(int nrows, const int32_t* ptr, const BuildParams& bp,
                    std::vector<int32_t>& order, std::vector<int32_t>& pos);

// Both orderings of an LP: rows of A (= y order) and rows of A' (= x order).  Inside a length
// class the rows are clustered lexicographically by the (internal) ids of the entries they
// touch, so that neighbouring rows gather from neighbouring addresses: one warp-wide gather
// then touches few 128 B lines (the L1 tag stage serves about one line per cycle per SM).
void plan_orders(int m, int n, const int32_t* ptr, const int32_t* ind, const int32_t* tptr, const int32_t* tind,
                 const BuildParams& bp, std::vector<int32_t>& orderY, std::vector<int32_t>& posY,
                 std::vector<int32_t>& orderX, std::vector<int32_t>& posX);

// Emit the tiled image of a CSR matrix whose rows follow `order`/`pos` and whose column
// ids are renamed through `colpos` (the other matrix's pos[]).
// `order` may list a subset of the rows (row partition over GPUs): tile k then covers the
// internal rows row_offset + k..., `ptr/ind/val` stay the full matrix.
void build_host_mat(int nrows, int ncols, const int32_t* ptr, const int32_t* ind,
                    const double* val, const std::vector<int32_t>& order,
                    const std::vector<int32_t>& colpos, const BuildParams& bp, HostMat& out,
                    uint32_t row_offset = 0, const DealFeedback* fb = nullptr);

// Row-per-lane image of a small matrix (warp-per-instance batch kernels): groups of 32 consecutive internal rows, lane = row;
// group g holds w_g = its longest row's length slots of 32 entries each (slot-major: entry s of the group's 32 rows is
// contiguous), off[g] .. off[g+1] in slots.  Shorter rows and the rows past nrows are padded with (column 0, value 0).
// Rows follow `order` (internal position -> original row), columns are renamed through `colpos`; the entries of a row keep
// their CSR order.  The internal order sorts rows by length class, so the padding stays small.
struct HostEll {
    std::vector<int32_t> idx;    // 32 * off.back()
    std::vector<double> val;     // 32 * off.back()
    std::vector<uint32_t> off;   // ngroups + 1
};
void build_host_ell(int nrows, const int32_t* ptr, const int32_t* ind, const double* val, const std::vector<int32_t>& order,
                    const std::vector<int32_t>& colpos, HostEll& out);

// Row partition over `nranks` GPUs: rows go to ranks by longest-processing-time on their
// nonzero counts; inside a rank the global (class, cluster) order is kept.  The global internal
// order is [rank 0's rows | rank 1's rows | ...], every slice padded to the same even length L
// (order_pad = -1 on padding), so one in-place all-gather exchanges the slices.
struct RowPartition {
    int L = 0;
    std::vector<int32_t> order_pad;               // nranks * L
    std::vector<int32_t> pos;                     // nrows: padded internal position of every row
    std::vector<std::vector<int32_t>> lists;      // per rank: its rows (original ids) in order
};
void partition_rows(int nrows, const int32_t* ptr, const std::vector<int32_t>& global_order, int nranks,
                    RowPartition& out);

// CSR transpose (counting sort); outputs sized by the callee.
void csr_transpose(int nrows, int ncols, const int32_t* ptr, const int32_t* ind,
                   const double* val, std::vector<int32_t>& tptr, std::vector<int32_t>& tind,
                   std::vector<double>& tval);

}  // namespace mllp
