// lp_format.cpp -- host-side builder of the warp-tiled SELL format (see lp_format.h).
// Input contract: the CSR arrays of linear_program_data.py:75-77 of the reference
// (scipy CSR: float64 data, int32 indices / indptr).
#include "lp_format.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace mllp {

namespace {

struct RowClass {
    int logL, nsteps, nchunks;
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

RowClass classify(int len, const BuildParams& bp)
{
    RowClass rc{0, 0, 1};
    const int cap = 64 * bp.max_steps;
    if (len > cap) {
        rc.nchunks = ceil_div(len, cap);
        rc.logL = 5;
        rc.nsteps = ceil_div(ceil_div(len, rc.nchunks), 64);  // steps per chunk
        return rc;
    }
    for (rc.logL = 0; rc.logL <= 5; ++rc.logL) {
        rc.nsteps = ceil_div(len, 2 << rc.logL);
        if (rc.nsteps <= bp.pref_steps) return rc;
    }
    rc.logL = 5;
    rc.nsteps = ceil_div(len, 64);
    return rc;
}

}  // namespace

void csr_transpose(int nrows, int ncols, const int32_t* ptr, const int32_t* ind,
                   const double* val, std::vector<int32_t>& tptr, std::vector<int32_t>& tind,
                   std::vector<double>& tval)
{
    const int64_t nnz = ptr[nrows];
    tptr.assign((size_t)ncols + 1, 0);
    tind.resize((size_t)nnz);
    tval.resize((size_t)nnz);
    for (int64_t k = 0; k < nnz; ++k) tptr[ind[k] + 1]++;
    for (int j = 0; j < ncols; ++j) tptr[j + 1] += tptr[j];
    std::vector<int32_t> fill(tptr.begin(), tptr.end() - 1);
    for (int i = 0; i < nrows; ++i)
        for (int32_t k = ptr[i]; k < ptr[i + 1]; ++k) {
            const int32_t q = fill[ind[k]]++;
            tind[q] = i;
            tval[q] = val[k];
        }
}

void plan_row_order(int nrows, const int32_t* ptr, const BuildParams& bp,
                    std::vector<int32_t>& order, std::vector<int32_t>& pos)
{
    const int S = std::max(bp.pref_steps, bp.max_steps) + 1;
    const int nkeys = 1 + 6 * S;
    std::vector<int32_t> key((size_t)nrows);
    std::vector<int64_t> count((size_t)nkeys + 1, 0);
    for (int r = 0; r < nrows; ++r) {
        const RowClass rc = classify(ptr[r + 1] - ptr[r], bp);
        // split rows first, then wide lanes-per-row first, then more steps first
        const int k = rc.nchunks > 1 ? 0 : 1 + (5 - rc.logL) * S + (S - 1 - rc.nsteps);
        key[r] = k;
        count[k + 1]++;
    }
    for (int k = 0; k < nkeys; ++k) count[k + 1] += count[k];
    order.resize((size_t)nrows);
    pos.resize((size_t)nrows);
    for (int r = 0; r < nrows; ++r) {  // stable: original order kept inside a class
        const int32_t p = (int32_t)count[key[r]]++;
        order[p] = r;
        pos[r] = p;
    }
}

void build_host_mat(int nrows, int ncols, const int32_t* ptr, const int32_t* ind,
                    const double* val, const std::vector<int32_t>& order,
                    const std::vector<int32_t>& colpos, const BuildParams& bp, HostMat& out)
{
    out = HostMat();
    out.nrows = nrows;
    out.ncols = ncols;
    out.nnz = ptr[nrows];

    // ---- 1. cut the internal row order into tiles --------------------------------------
    struct ProtoTile {
        uint32_t row_base;  // internal row (first of the tile, or the split row)
        uint16_t nsteps;
        uint8_t logL, nrows;
        int32_t split;      // split table index or -1
        uint32_t chunk;     // chunk index for split rows
        uint32_t chunk_len; // entries per chunk (split rows)
    };
    std::vector<ProtoTile> proto;
    int p = 0;
    while (p < nrows) {
        const int r = order[p];
        const int len = ptr[r + 1] - ptr[r];
        const RowClass rc = classify(len, bp);
        if (rc.nchunks > 1) {
            SplitRow sr{(uint32_t)p, out.num_partials, (uint32_t)rc.nchunks, 0};
            const uint32_t chunk_len = 64u * (uint32_t)rc.nsteps;
            // the last chunks may be shorter (or even empty if rounding over-covers)
            uint32_t used_chunks = (uint32_t)ceil_div(len, (int)chunk_len);
            sr.nchunks = used_chunks;
            for (uint32_t c = 0; c < used_chunks; ++c) {
                const int clen = std::min<int>((int)chunk_len, len - (int)(c * chunk_len));
                proto.push_back({(uint32_t)p, (uint16_t)ceil_div(clen, 64), 5, 1,
                                 (int32_t)out.splits.size(), c, chunk_len});
            }
            out.num_partials += used_chunks;
            out.splits.push_back(sr);
            ++p;
            continue;
        }
        const int rows_per_tile = 32 >> rc.logL;
        int q = p + 1;
        while (q < nrows && q - p < rows_per_tile) {
            const int r2 = order[q];
            const RowClass rc2 = classify(ptr[r2 + 1] - ptr[r2], bp);
            if (rc2.nchunks != 1 || rc2.logL != rc.logL || rc2.nsteps != rc.nsteps) break;
            ++q;
        }
        proto.push_back({(uint32_t)p, (uint16_t)rc.nsteps, (uint8_t)rc.logL, (uint8_t)(q - p),
                         -1, 0, 0});
        p = q;
    }

    // ---- 2. deal tiles to CTAs, heavy first, snake order --------------------------------
    const int G = std::max(1, bp.num_ctas);
    std::vector<uint32_t> by_cost(proto.size());
    std::iota(by_cost.begin(), by_cost.end(), 0u);
    std::stable_sort(by_cost.begin(), by_cost.end(), [&](uint32_t a, uint32_t b) {
        const int sa = proto[a].split >= 0, sb = proto[b].split >= 0;
        if (sa != sb) return sa > sb;  // split chunks first: their join is the critical path
        return proto[a].nsteps > proto[b].nsteps;
    });
    std::vector<std::vector<uint32_t>> per_cta((size_t)G);
    for (size_t i = 0; i < by_cost.size(); ++i) {
        const size_t pass = i / (size_t)G, k = i % (size_t)G;
        const size_t cta = (pass & 1) ? (size_t)G - 1 - k : k;
        per_cta[cta].push_back(by_cost[i]);
    }

    // ---- 3. emit CTA-major storage ---------------------------------------------------------
    uint64_t total_steps = 0;
    for (const ProtoTile& t : proto) total_steps += t.nsteps;
    out.total_steps = total_steps;
    out.vals.assign((size_t)total_steps * 64, 0.0);
    out.idx.assign((size_t)total_steps * 64, 0);
    out.tiles.reserve(proto.size());
    out.cta_begin.assign((size_t)G + 1, 0);
    out.cta_step_begin.assign((size_t)G + 1, 0);

    std::vector<std::pair<int32_t, double>> row_buf;
    std::vector<std::vector<std::pair<int32_t, double>>> split_rows(out.splits.size());
    uint64_t step_cursor = 0;
    for (int g = 0; g < G; ++g) {
        out.cta_begin[g] = (uint32_t)out.tiles.size();
        out.cta_step_begin[g] = (uint32_t)step_cursor;
        for (uint32_t ti : per_cta[g]) {
            const ProtoTile& t = proto[ti];
            const int L = 1 << t.logL;
            Tile tile;
            tile.off = (uint32_t)step_cursor;
            tile.row_base = t.split >= 0 ? t.chunk : t.row_base;
            tile.nsteps = t.nsteps;
            tile.logL = t.logL;
            tile.nrows = t.nrows;
            tile.split = t.split;
            out.tiles.push_back(tile);
            for (int rr = 0; rr < t.nrows; ++rr) {
                const int r = order[t.row_base + rr];
                const int len = ptr[r + 1] - ptr[r];
                if (t.split >= 0 && !split_rows[t.split].empty()) {
                    row_buf = split_rows[t.split];  // sorted once per split row
                } else {
                    row_buf.clear();
                    for (int32_t k = ptr[r]; k < ptr[r + 1]; ++k)
                        row_buf.emplace_back(colpos[ind[k]], val[k]);
                    std::sort(row_buf.begin(), row_buf.end(),
                              [](const std::pair<int32_t, double>& a,
                                 const std::pair<int32_t, double>& b) { return a.first < b.first; });
                    if (t.split >= 0) split_rows[t.split] = row_buf;
                }
                int e0 = 0, e1 = len;
                if (t.split >= 0) {
                    e0 = (int)(t.chunk * t.chunk_len);
                    e1 = std::min(len, e0 + (int)t.chunk_len);
                }
                const int32_t pad_idx = len > 0 ? row_buf[e0 < len ? e0 : 0].first : 0;
                const int slots = t.nsteps * 2 * L;
                for (int e = 0; e < slots; ++e) {
                    const int s = e / (2 * L), q = e % (2 * L);
                    const int half = q / L, lane = rr * L + q % L;
                    const size_t at = ((size_t)(step_cursor + s) * 32 + lane) * 2 + half;
                    if (e0 + e < e1) {
                        out.idx[at] = row_buf[e0 + e].first;
                        out.vals[at] = row_buf[e0 + e].second;
                    } else {
                        out.idx[at] = pad_idx;
                        out.vals[at] = 0.0;
                    }
                }
            }
            // lanes of unused rows in a partially filled tile keep val = 0, idx = 0
            step_cursor += t.nsteps;
        }
        const int steps_here = (int)(step_cursor - out.cta_step_begin[g]);
        out.max_cta_steps = std::max(out.max_cta_steps, steps_here);
        out.max_cta_tiles = std::max<int>(out.max_cta_tiles, (int)per_cta[g].size());
    }
    out.cta_begin[G] = (uint32_t)out.tiles.size();
    out.cta_step_begin[G] = (uint32_t)step_cursor;
}

}  // namespace mllp

// ---------------------------------------------------------------------------------------
// Host-side self check of the builder (no device needed): walks the tiles exactly as the
// kernel does (per-lane sequential sums, then the butterfly over L lanes, then the fixed
// order join of split rows) on a deterministic test vector and compares every row with
// the plain CSR dot product.  Used by the CPU test-suite to validate the format logic.
extern "C" int mllp_format_selfcheck(int32_t m, int32_t n, int64_t nnz, const int32_t* indptr,
                                     const int32_t* indices, const double* values, int32_t num_ctas,
                                     int32_t pref_steps, int32_t max_steps, double* out8)
{
    using namespace mllp;
    if (m < 0 || n < 0 || !indptr || !out8 || (int64_t)indptr[m] != nnz) return 1001;
    BuildParams bp;
    bp.num_ctas = num_ctas;
    bp.pref_steps = pref_steps;
    bp.max_steps = max_steps < pref_steps ? pref_steps : max_steps;
    std::vector<int32_t> tptr, tind;
    std::vector<double> tval;
    csr_transpose(m, n, indptr, indices, values, tptr, tind, tval);
    std::vector<int32_t> orderY, posY, orderX, posX;
    plan_row_order(m, indptr, bp, orderY, posY);
    plan_row_order(n, tptr.data(), bp, orderX, posX);
    HostMat H[2];
    build_host_mat(m, n, indptr, indices, values, orderY, posX, bp, H[0]);
    build_host_mat(n, m, tptr.data(), tind.data(), tval.data(), orderX, posY, bp, H[1]);

    double worst = 0.0;
    int64_t rows_seen = 0;
    for (int which = 0; which < 2; ++which) {
        const HostMat& M = H[which];
        const int nr = M.nrows, nc = M.ncols;
        const int32_t* ptr = which ? tptr.data() : indptr;
        const int32_t* ind = which ? tind.data() : indices;
        const double* val = which ? tval.data() : values;
        const std::vector<int32_t>& order = which ? orderX : orderY;
        const std::vector<int32_t>& colorder = which ? orderY : orderX;
        std::vector<double> v_user((size_t)nc), v_int((size_t)nc), out_int((size_t)nr, NAN), partial(M.num_partials, 0.0);
        for (int j = 0; j < nc; ++j) v_user[j] = 0.25 + (double)((j * 2654435761u) % 1000u) / 997.0;
        for (int k = 0; k < nc; ++k) v_int[k] = v_user[colorder[k]];
        std::vector<uint32_t> arrived(M.splits.size(), 0);
        if (M.cta_begin.size() != (size_t)bp.num_ctas + 1 || M.cta_begin.back() != M.tiles.size()) return 2;
        for (const Tile& t : M.tiles) {
            const int L = 1 << t.logL;
            double lane_sum[32];
            for (int lane = 0; lane < 32; ++lane) {
                double s = 0.0;
                for (int st = 0; st < t.nsteps; ++st) {
                    const size_t at = ((size_t)(t.off + st) * 32 + lane) * 2;
                    if (M.idx[at] < 0 || M.idx[at] >= nc || M.idx[at + 1] < 0 || M.idx[at + 1] >= nc) return 3;
                    s = std::fma(M.vals[at], v_int[M.idx[at]], s);
                    s = std::fma(M.vals[at + 1], v_int[M.idx[at + 1]], s);
                }
                lane_sum[lane] = s;
            }
            for (int o = L >> 1; o > 0; o >>= 1) {
                double nxt[32];
                for (int lane = 0; lane < 32; ++lane) nxt[lane] = lane_sum[lane] + lane_sum[lane ^ o];
                for (int lane = 0; lane < 32; ++lane) lane_sum[lane] = nxt[lane];
            }
            if (t.split < 0) {
                if (t.nrows < 1 || t.nrows > (32 >> t.logL)) return 4;
                for (int rr = 0; rr < t.nrows; ++rr) {
                    const uint32_t r = t.row_base + rr;
                    if (r >= (uint32_t)nr || !std::isnan(out_int[r])) return 5;  // each row exactly once
                    out_int[r] = lane_sum[rr * L];
                }
            } else {
                const SplitRow& sr = M.splits[t.split];
                if (t.row_base >= sr.nchunks) return 6;
                partial[sr.first_slot + t.row_base] = lane_sum[0];
                if (++arrived[t.split] == sr.nchunks) {
                    double ls[32] = {0};
                    for (uint32_t k = 0; k < sr.nchunks; ++k) ls[k % 32] += partial[sr.first_slot + k];
                    for (int o = 16; o > 0; o >>= 1) {
                        double nxt[32];
                        for (int lane = 0; lane < 32; ++lane) nxt[lane] = ls[lane] + ls[lane ^ o];
                        for (int lane = 0; lane < 32; ++lane) ls[lane] = nxt[lane];
                    }
                    if (sr.row >= (uint32_t)nr || !std::isnan(out_int[sr.row])) return 7;
                    out_int[sr.row] = ls[0];
                }
            }
        }
        for (int k = 0; k < nr; ++k) {
            if (std::isnan(out_int[k])) return 8;  // a row was never produced
            const int r = order[k];
            double ref = 0.0, mag = 0.0;
            for (int32_t q = ptr[r]; q < ptr[r + 1]; ++q) {
                ref += val[q] * v_user[ind[q]];
                mag += std::fabs(val[q] * v_user[ind[q]]);
            }
            const double err = std::fabs(ref - out_int[k]) / (mag > 0.0 ? mag : 1.0);
            if (err > worst) worst = err;
            ++rows_seen;
        }
    }
    out8[0] = worst;
    out8[1] = (double)H[0].tiles.size();
    out8[2] = (double)H[1].tiles.size();
    out8[3] = nnz > 0 ? (double)(H[0].total_steps * 64) / (double)nnz : 0.0;  // padding factor A
    out8[4] = nnz > 0 ? (double)(H[1].total_steps * 64) / (double)nnz : 0.0;  // padding factor A'
    out8[5] = (double)H[0].splits.size();
    out8[6] = (double)H[0].max_cta_steps;
    out8[7] = (double)H[1].max_cta_steps;
    return rows_seen == (int64_t)m + n ? 0 : 9;
}
