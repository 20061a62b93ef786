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

// Raise max_steps (the split threshold and chunk length) until no CTA gets more than
// SPLIT_SLOTS split chunks.  Deterministic in (row lengths, bp), so the row-order planner and
// the emitter agree.
BuildParams effective_params(int nrows, const int32_t* ptr, BuildParams bp)
{
    if (bp.final_params) return bp;
    const int G = std::max(1, bp.num_ctas);
    if (bp.pref_steps < 1) bp.pref_steps = 1;
    if (bp.max_steps < bp.pref_steps) bp.max_steps = bp.pref_steps;
    for (;; bp.max_steps *= 2) {
        const int64_t cap = 64LL * bp.max_steps;
        int64_t chunks = 0;
        for (int r = 0; r < nrows; ++r) {
            const int64_t len = ptr[r + 1] - ptr[r];
            if (len > cap) chunks += (len + cap - 1) / cap;
        }
        if ((chunks + G - 1) / G <= SPLIT_SLOTS || bp.max_steps >= 16384) {
            bp.final_params = true;
            return bp;
        }
    }
}


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

void plan_row_order(int nrows, const int32_t* ptr, const BuildParams& bp_in,
                    std::vector<int32_t>& order, std::vector<int32_t>& pos)
{
    const BuildParams bp = effective_params(nrows, ptr, bp_in);
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

namespace {

// Re-sort `order` inside every class: rows are compared by their sorted lists of mapped entry
// ids (otherpos).  Split rows (class key 0) keep their place.
void cluster_within_classes(int nrows, const int32_t* ptr, const int32_t* ind, const BuildParams& bp_in,
                            const std::vector<int32_t>& otherpos, std::vector<int32_t>& order,
                            std::vector<int32_t>& pos)
{
    const BuildParams bp = effective_params(nrows, ptr, bp_in);
    // mapped, sorted entry lists (flat)
    std::vector<int32_t> keys((size_t)ptr[nrows]);
    for (int r = 0; r < nrows; ++r) {
        for (int32_t k = ptr[r]; k < ptr[r + 1]; ++k) keys[k] = otherpos[ind[k]];
        std::sort(keys.begin() + ptr[r], keys.begin() + ptr[r + 1]);
    }
    auto cls = [&](int r) {
        const RowClass rc = classify(ptr[r + 1] - ptr[r], bp);
        return rc.nchunks > 1 ? -1 : rc.logL * 65536 + rc.nsteps;
    };
    size_t a = 0;
    while (a < (size_t)nrows) {
        size_t b = a + 1;
        const int ca = cls(order[a]);
        while (b < (size_t)nrows && cls(order[b]) == ca) ++b;
        if (ca >= 0)
            std::stable_sort(order.begin() + a, order.begin() + b, [&](int32_t r1, int32_t r2) {
                return std::lexicographical_compare(keys.begin() + ptr[r1], keys.begin() + ptr[r1 + 1],
                                                    keys.begin() + ptr[r2], keys.begin() + ptr[r2 + 1]);
            });
        a = b;
    }
    for (int p = 0; p < nrows; ++p) pos[order[p]] = p;
}

}  // namespace

void plan_orders(int m, int n, const int32_t* ptr, const int32_t* ind, const int32_t* tptr, const int32_t* tind,
                 const BuildParams& bp, std::vector<int32_t>& orderY, std::vector<int32_t>& posY,
                 std::vector<int32_t>& orderX, std::vector<int32_t>& posX)
{
    plan_row_order(m, ptr, bp, orderY, posY);
    plan_row_order(n, tptr, bp, orderX, posX);
    if (bp.cluster) {
        for (int round = 0; round < std::max(1, bp.cluster_rounds); ++round) {
            cluster_within_classes(n, tptr, tind, bp, posY, orderX, posX);   // columns by the rows they touch
            cluster_within_classes(m, ptr, ind, bp, posX, orderY, posY);     // rows by the columns they touch
        }
    }
}

void partition_rows(int nrows, const int32_t* ptr, const std::vector<int32_t>& global_order, int nranks,
                    RowPartition& out)
{
    nranks = std::max(1, nranks);
    std::vector<int> owner((size_t)nrows, 0);
    if (nranks > 1) {
        std::vector<int32_t> by((size_t)nrows);
        std::iota(by.begin(), by.end(), 0);
        std::stable_sort(by.begin(), by.end(),
                         [&](int32_t a, int32_t b) { return ptr[a + 1] - ptr[a] > ptr[b + 1] - ptr[b]; });
        std::vector<int64_t> load((size_t)nranks, 0), cnt((size_t)nranks, 0);
        for (int32_t r : by) {
            int best = 0;
            for (int p = 1; p < nranks; ++p)
                if (load[p] < load[best] || (load[p] == load[best] && cnt[p] < cnt[best])) best = p;
            owner[r] = best;
            load[best] += (ptr[r + 1] - ptr[r]) + 1;  // +1: every row also costs a vector update
            cnt[best]++;
        }
    }
    out.lists.assign((size_t)nranks, {});
    for (int32_t r : global_order) out.lists[owner[r]].push_back(r);  // stable: class order kept inside a rank
    size_t mx = 0;
    for (auto& l : out.lists) mx = std::max(mx, l.size());
    out.L = (int)((mx + 1) & ~(size_t)1);
    out.order_pad.assign((size_t)out.L * nranks, -1);
    out.pos.assign((size_t)nrows, 0);
    for (int p = 0; p < nranks; ++p)
        for (size_t k = 0; k < out.lists[p].size(); ++k) {
            out.order_pad[(size_t)p * out.L + k] = out.lists[p][k];
            out.pos[out.lists[p][k]] = (int32_t)((size_t)p * out.L + k);
        }
}

void build_host_mat(int nrows, int ncols, const int32_t* ptr, const int32_t* ind,
                    const double* val, const std::vector<int32_t>& order,
                    const std::vector<int32_t>& colpos, const BuildParams& bp_in, HostMat& out,
                    uint32_t row_offset, const DealFeedback* fb)
{
    const BuildParams bp = effective_params(nrows, ptr, bp_in);
    out = HostMat();
    out.nrows = nrows;
    out.ncols = ncols;
    out.nnz = 0;
    for (int p = 0; p < nrows; ++p) out.nnz += ptr[order[p] + 1] - ptr[order[p]];
    out.nnz_emitted = out.nnz;
    const int G = std::max(1, bp.num_ctas);

    struct ProtoTile {
        uint32_t row_base;  // internal row (first of the tile, or the split row)
        uint16_t nsteps;
        uint8_t logL, nrows;
        int32_t split;      // split table index or -1
        uint32_t e0, e1;    // entry range of the row (split chunks)
    };

    // ---- 1a. split rows (they lead the internal order): one chunk sequence ---------------
    int nsplit_rows = 0;
    while (nsplit_rows < nrows) {
        const int r = order[nsplit_rows];
        if (classify(ptr[r + 1] - ptr[r], bp).nchunks <= 1) break;
        ++nsplit_rows;
    }
    const int chunk_steps = bp.max_steps;  // effective_params() bounds the chunks per CTA
    std::vector<ProtoTile> chunks;
    for (int p = 0; p < nsplit_rows; ++p) {
        const int r = order[p];
        const int len = ptr[r + 1] - ptr[r];
        const int k = ceil_div(len, 64 * chunk_steps);
        const int clen = 64 * ceil_div(ceil_div(len, k), 64);  // balanced, whole warp-steps
        for (int e0 = 0; e0 < len; e0 += clen) {
            const int e1 = std::min(len, e0 + clen);
            chunks.push_back({(uint32_t)p, (uint16_t)ceil_div(e1 - e0, 64), 5, 1, p, (uint32_t)e0, (uint32_t)e1});
        }
        out.splits.push_back({(uint32_t)p + row_offset, 0u, 0u, 0u});
    }

    // ---- 1b. regular rows: tiles of 32/L consecutive rows of one class -----------------------
    std::vector<ProtoTile> regular;
    std::vector<int32_t> sector_buf;
    for (int p = nsplit_rows; p < nrows;) {
        const int r = order[p];
        const RowClass rc = classify(ptr[r + 1] - ptr[r], bp);
        const int rows_per_tile = 32 >> rc.logL;
        int q = p + 1;
        while (q < nrows && q - p < rows_per_tile) {
            const int r2 = order[q];
            const RowClass rc2 = classify(ptr[r2 + 1] - ptr[r2], bp);
            if (rc2.nchunks != 1 || rc2.logL != rc.logL || rc2.nsteps != rc.nsteps) break;
            ++q;
        }
        ProtoTile t{(uint32_t)p, (uint16_t)rc.nsteps, (uint8_t)rc.logL, (uint8_t)(q - p), -1, 0, 0};
        if (bp.contiguous) {
            // e1 doubles as the tile's cost for the contiguous dealing: steps + the distinct 32 B sectors
            // its gathers touch (what the SM has to fetch)
            sector_buf.clear();
            for (int pp = p; pp < q; ++pp) {
                const int rr = order[pp];
                for (int32_t k = ptr[rr]; k < ptr[rr + 1]; ++k) sector_buf.push_back(colpos[ind[k]] >> 2);
            }
            std::sort(sector_buf.begin(), sector_buf.end());
            t.e1 = (uint32_t)(std::unique(sector_buf.begin(), sector_buf.end()) - sector_buf.begin());
        }
        regular.push_back(t);
        p = q;
    }

    // ---- 2. deal to CTAs: chunks in contiguous runs, regular tiles to the least loaded CTA -----
    std::vector<std::vector<ProtoTile>> per_cta((size_t)G);
    std::vector<int64_t> load((size_t)G, 0);
    // chunks go to CTAs in row order; a row that fits in one CTA's quota is never cut across CTAs
    // (then the CTA finishes it alone), larger rows fill consecutive CTAs
    const size_t C = chunks.size();
    {
        // quota of the current CTA = what is left, spread over the CTAs that are left
        size_t g = 0, used = 0;
        auto quota_of = [&](size_t q_done, size_t g_now) {
            const size_t left = C - q_done, ctas = (size_t)G - g_now;
            return (left + ctas - 1) / ctas;
        };
        size_t quota = quota_of(0, 0);
        for (size_t q = 0; q < C;) {
            size_t q2 = q;
            while (q2 < C && chunks[q2].split == chunks[q].split) ++q2;
            const size_t k = q2 - q;
            // a row that would fit in a fresh CTA but not in the rest of this one starts a new CTA
            if (k <= quota && used > 0 && used + k > quota && g + 1 < (size_t)G) { ++g; used = 0; quota = quota_of(q, g); }
            for (; q < q2; ++q) {
                if (used >= quota && g + 1 < (size_t)G) { ++g; used = 0; quota = quota_of(q, g); }
                per_cta[g].push_back(chunks[q]);
                load[g] += chunks[q].nsteps + 2;
                ++used;
            }
        }
    }
    out.cta_nsplit.assign((size_t)G, 0);
    out.cta_lsplit_begin.assign((size_t)G + 1, 0);
    for (int g = 0; g < G; ++g) {
        out.cta_nsplit[g] = (uint32_t)per_cta[g].size();
        out.cta_lsplit_begin[g] = (uint32_t)out.lsplits.size();
        for (size_t k = 0; k < per_cta[g].size();) {
            size_t k2 = k;
            while (k2 < per_cta[g].size() && per_cta[g][k2].split == per_cta[g][k].split) ++k2;
            SplitRow& sr = out.splits[per_cta[g][k].split];
            out.lsplits.push_back({(uint32_t)per_cta[g][k].split, sr.nparts /* rank, fixed below */, (uint16_t)k,
                                   (uint16_t)(k2 - k), 0u});
            sr.nparts++;
            k = k2;
        }
    }
    out.cta_lsplit_begin[G] = (uint32_t)out.lsplits.size();
    for (SplitRow& sr : out.splits) {
        sr.first_slot = out.num_partials;
        out.num_partials += sr.nparts;
    }
    for (LocalSplit& ls : out.lsplits) {
        // the rank was stored in gslot: the highest rank of a row is its finisher
        ls.finisher = (ls.gslot + 1 == out.splits[ls.split_id].nparts) ? 1u : 0u;
        ls.gslot += out.splits[ls.split_id].first_slot;
    }

    out.reg_cta.assign(regular.size(), 0u);
    out.cta_load.assign((size_t)G, 0.0);
    if (bp.contiguous) {
        // Contiguous runs of the (class, cluster)-ordered tile list per CTA, cut by a cost prefix sum:
        // neighbouring rows gather from neighbouring addresses, so keeping them on ONE SM lets its L1
        // serve the sectors they share (the least-loaded dealing below scatters them over all SMs).
        // Model cost of a tile: steps + distinct gathered sectors; a tuning round scales it by what was measured.
        const bool have_w = fb && fb->tile_w.size() == regular.size();
        const bool have_f = fb && fb->cta_f.size() == (size_t)G;
        auto cost = [&](size_t k) {
            const ProtoTile& t = regular[k];
            return (double)((int64_t)t.e1 + 2 * (int64_t)t.nsteps + 4) * (have_w ? fb->tile_w[k] : 1.0);
        };
        std::vector<double> cl((size_t)G);
        for (int g = 0; g < G; ++g) cl[g] = 3.0 * (double)load[g] * (have_f ? fb->cta_f[g] : 1.0);   // split chunks: ~3 units per step
        double total = 0;
        for (int g = 0; g < G; ++g) total += cl[g];
        for (size_t k = 0; k < regular.size(); ++k) total += cost(k);
        int g = 0;
        double acc_cost = 0;   // cost handed to CTAs 0..g-1
        double mine = cl[0];
        for (size_t k = 0; k < regular.size(); ++k) {
            // move on when this CTA has reached its share of the total
            while (g + 1 < G && (acc_cost + mine) * G >= total * (double)(g + 1)) {
                acc_cost += mine;
                out.cta_load[g] = mine;
                ++g;
                mine = cl[g];
            }
            per_cta[g].push_back(regular[k]);
            out.reg_cta[k] = (uint32_t)g;
            mine += cost(k);
        }
        out.cta_load[g] = mine;
        for (int h = g + 1; h < G; ++h) out.cta_load[h] = cl[h];
    } else {
        std::vector<uint32_t> by_cost(regular.size());
        std::iota(by_cost.begin(), by_cost.end(), 0u);
        std::stable_sort(by_cost.begin(), by_cost.end(),
                         [&](uint32_t a, uint32_t b) { return regular[a].nsteps > regular[b].nsteps; });
        // min-heap of (load, cta); a tuning round biases the starting loads by what was measured
        const bool have_b = fb && fb->cta_bias.size() == (size_t)G;
        std::vector<std::pair<double, int>> heap;
        for (int g = 0; g < G; ++g) heap.emplace_back((double)load[g] + (have_b ? fb->cta_bias[g] : 0.0), g);
        auto cmp = [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a > b; };
        std::make_heap(heap.begin(), heap.end(), cmp);
        for (int g = 0; g < G; ++g) out.cta_load[g] = (double)load[g];
        for (uint32_t ti : by_cost) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            auto& top = heap.back();
            per_cta[top.second].push_back(regular[ti]);
            out.reg_cta[ti] = (uint32_t)top.second;
            top.first += regular[ti].nsteps + 2;
            out.cta_load[top.second] += regular[ti].nsteps + 2;
            std::push_heap(heap.begin(), heap.end(), cmp);
        }
    }

    // ---- 3. emit CTA-major storage ---------------------------------------------------------
    uint64_t total_steps = 0;
    size_t total_tiles = 0;
    for (int g = 0; g < G; ++g) {
        total_tiles += per_cta[g].size();
        for (const ProtoTile& t : per_cta[g]) total_steps += t.nsteps;
    }
    out.total_steps = total_steps;
    out.vals.assign((size_t)total_steps * 64, 0.0);
    out.idx.assign((size_t)total_steps * 64, 0);
    out.tiles.reserve(total_tiles);
    out.cta_begin.assign((size_t)G + 1, 0);
    out.cta_step_begin.assign((size_t)G + 1, 0);

    std::vector<std::pair<int32_t, double>> row_buf;
    int cached_split = -1;  // chunks of one row are emitted back to back: sort it once
    uint64_t step_cursor = 0;
    for (int g = 0; g < G; ++g) {
        out.cta_begin[g] = (uint32_t)out.tiles.size();
        out.cta_step_begin[g] = (uint32_t)step_cursor;
        for (size_t k = 0; k < per_cta[g].size(); ++k) {
            const ProtoTile& t = per_cta[g][k];
            const int L = 1 << t.logL;
            Tile tile;
            tile.off = (uint32_t)step_cursor;
            tile.row_base = t.row_base + row_offset;
            tile.nsteps = t.nsteps;
            tile.logL = t.logL;
            tile.nrows = t.nrows;
            tile.split = t.split >= 0 ? (int32_t)k : -1;  // split chunks lead the list: slot = position
            out.tiles.push_back(tile);
            for (int rr = 0; rr < t.nrows; ++rr) {
                const int r = order[t.row_base + rr];
                const int len = ptr[r + 1] - ptr[r];
                if (t.split < 0 || t.split != cached_split) {
                    row_buf.clear();
                    for (int32_t q = ptr[r]; q < ptr[r + 1]; ++q) row_buf.emplace_back(colpos[ind[q]], val[q]);
                    std::sort(row_buf.begin(), row_buf.end(),
                              [](const std::pair<int32_t, double>& a, const std::pair<int32_t, double>& b) {
                                  return a.first < b.first;
                              });
                    cached_split = t.split;
                }
                const int e0 = t.split >= 0 ? (int)t.e0 : 0;
                const int e1 = t.split >= 0 ? (int)t.e1 : len;
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
        int rows_here = 0;
        for (const ProtoTile& t : per_cta[g]) rows_here += t.split < 0 ? t.nrows : 0;
        out.max_cta_rows = std::max(out.max_cta_rows, rows_here);
    }
    out.cta_begin[G] = (uint32_t)out.tiles.size();
    out.cta_step_begin[G] = (uint32_t)step_cursor;
}

void build_host_ell(int nrows, const int32_t* ptr, const int32_t* ind, const double* val, const std::vector<int32_t>& order,
                    const std::vector<int32_t>& colpos, HostEll& out)
{
    const int ngroups = (nrows + 31) / 32;
    out.idx.clear(); out.val.clear();
    out.off.assign((size_t)ngroups + 1, 0u);
    for (int g = 0; g < ngroups; ++g) {
        int w = 0;
        for (int l = 0; l < 32 && g * 32 + l < nrows; ++l) {
            const int r = order[(size_t)g * 32 + l];
            w = std::max(w, (int)(ptr[r + 1] - ptr[r]));
        }
        out.off[(size_t)g + 1] = out.off[g] + (uint32_t)w;
    }
    out.idx.assign((size_t)32 * out.off.back(), 0);
    out.val.assign((size_t)32 * out.off.back(), 0.0);
    for (int g = 0; g < ngroups; ++g) {
        for (int l = 0; l < 32 && g * 32 + l < nrows; ++l) {
            const int r = order[(size_t)g * 32 + l];
            uint32_t s = out.off[g];
            for (int32_t q = ptr[r]; q < ptr[r + 1]; ++q, ++s) {
                out.idx[(size_t)s * 32 + l] = colpos[ind[q]];
                out.val[(size_t)s * 32 + l] = val[q];
            }
        }
    }
}

}  // namespace mllp

// ---------------------------------------------------------------------------------------
// Host-side self check of the builder (no device needed): walks the tiles exactly as the
// kernel does (per-lane sequential sums, then the butterfly over L lanes, then the fixed
// order join of split rows) on a deterministic test vector and compares every row with
// the plain CSR dot product.  Used by the CPU test-suite to validate the format logic.
static int format_selfcheck_impl(int32_t m, int32_t n, int64_t nnz, const int32_t* indptr,
                                 const int32_t* indices, const double* values, int32_t num_ctas,
                                 int32_t pref_steps, int32_t max_steps, bool contiguous, double* out8)
{
    using namespace mllp;
    if (m < 0 || n < 0 || !indptr || !out8 || (int64_t)indptr[m] != nnz) return 1001;
    BuildParams bp;
    bp.num_ctas = num_ctas;
    bp.pref_steps = pref_steps;
    bp.max_steps = max_steps < pref_steps ? pref_steps : max_steps;
    bp.contiguous = contiguous;
    std::vector<int32_t> tptr, tind;
    std::vector<double> tval;
    csr_transpose(m, n, indptr, indices, values, tptr, tind, tval);
    std::vector<int32_t> orderY, posY, orderX, posX;
    plan_orders(m, n, indptr, indices, tptr.data(), tind.data(), bp, orderY, posY, orderX, posX);
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
        if (M.cta_begin.size() != (size_t)bp.num_ctas + 1 || M.cta_begin.back() != M.tiles.size()) return 2;
        auto butterfly = [](double* ls, int L) {
            for (int o = L >> 1; o > 0; o >>= 1) {
                double nxt[32];
                for (int lane = 0; lane < 32; ++lane) nxt[lane] = ls[lane] + ls[lane ^ o];
                for (int lane = 0; lane < 32; ++lane) ls[lane] = nxt[lane];
            }
        };
        std::vector<uint32_t> arrived(M.splits.size(), 0);
        for (int g = 0; g < bp.num_ctas; ++g) {
            double spart[SPLIT_SLOTS];
            if (M.cta_nsplit[g] > (uint32_t)SPLIT_SLOTS) return 10;
            for (uint32_t ti = M.cta_begin[g]; ti < M.cta_begin[g + 1]; ++ti) {
                const Tile& t = M.tiles[ti];
                const uint32_t local = ti - M.cta_begin[g];
                if ((t.split >= 0) != (local < M.cta_nsplit[g])) return 11;  // split chunks lead the list
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
                butterfly(lane_sum, L);
                if (t.split < 0) {
                    if (t.nrows < 1 || t.nrows > (32 >> t.logL)) return 4;
                    for (int rr = 0; rr < t.nrows; ++rr) {
                        const uint32_t r = t.row_base + rr;
                        if (r >= (uint32_t)nr || !std::isnan(out_int[r])) return 5;  // each row exactly once
                        out_int[r] = lane_sum[rr * L];
                    }
                } else {
                    if ((uint32_t)t.split != local) return 6;
                    spart[t.split] = lane_sum[0];
                }
            }
            // publish: one partial per (CTA, split row), slots summed in order
            for (uint32_t li = M.cta_lsplit_begin[g]; li < M.cta_lsplit_begin[g + 1]; ++li) {
                const LocalSplit& ls = M.lsplits[li];
                const SplitRow& sr = M.splits[ls.split_id];
                if (ls.gslot < sr.first_slot || ls.gslot >= sr.first_slot + sr.nparts) return 12;
                double pl[32] = {0};
                for (int k = 0; k < ls.count; ++k) pl[k % 32] += spart[ls.first + k];
                butterfly(pl, 32);
                partial[ls.gslot] = pl[0];
                if (++arrived[ls.split_id] == sr.nparts) {
                    double ls32[32] = {0};   // per lane: k = k0 + 32u + lane, u ascending, k0 in steps of 256
                    for (uint32_t k = 0; k < sr.nparts; ++k) ls32[k % 32] += partial[sr.first_slot + k];
                    butterfly(ls32, 32);
                    if (sr.row >= (uint32_t)nr || !std::isnan(out_int[sr.row])) return 7;
                    out_int[sr.row] = ls32[0];
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


// Both dealings of the regular tiles (least-loaded and contiguous runs) are checked; the statistics
// returned are those of the least-loaded build.
extern "C" int mllp_format_selfcheck(int32_t m, int32_t n, int64_t nnz, const int32_t* indptr,
                                     const int32_t* indices, const double* values, int32_t num_ctas,
                                     int32_t pref_steps, int32_t max_steps, double* out8)
{
    if (!out8) return 1001;
    double tmp[8];
    const int rc = format_selfcheck_impl(m, n, nnz, indptr, indices, values, num_ctas, pref_steps, max_steps, true, tmp);
    if (rc != 0) return 100 + rc;
    if (!(tmp[0] < 1e-9)) return 199;
    return format_selfcheck_impl(m, n, nnz, indptr, indices, values, num_ctas, pref_steps, max_steps, false, out8);
}

// Host-only self check of the row-per-lane images (build_host_ell) of A and A' in the batch builder's internal orders: the
// lane walk of the warp-per-instance kernels replayed on the CPU (slot after slot, sequential sum per row) against the plain
// CSR products.  out4: [0] worst relative row error, [1] / [2] slots of A / A', [3] padding factor (stored / nonzeros).
extern "C" int mllp_ell_selfcheck(int32_t m, int32_t n, int64_t nnz, const int32_t* indptr, const int32_t* indices,
                                  const double* values, double* out4)
{
    using namespace mllp;
    if (m < 0 || n < 0 || !indptr || !out4 || (int64_t)indptr[m] != nnz) return 1001;
    BuildParams bp;
    bp.num_ctas = 1;
    bp.pref_steps = 2;
    bp.max_steps = 4;
    std::vector<int32_t> tptr, tind;
    std::vector<double> tval;
    csr_transpose(m, n, indptr, indices, values, tptr, tind, tval);
    std::vector<int32_t> orderY, posY, orderX, posX;
    plan_orders(m, n, indptr, indices, tptr.data(), tind.data(), bp, orderY, posY, orderX, posX);
    HostEll EA, ET;
    build_host_ell(m, indptr, indices, values, orderY, posX, EA);
    build_host_ell(n, tptr.data(), tind.data(), tval.data(), orderX, posY, ET);
    auto replay = [](const HostEll& E, int nrows, const std::vector<double>& vec_int, std::vector<double>& out_int) {
        const int ngroups = (nrows + 31) / 32;
        if ((int)E.off.size() != ngroups + 1 || E.idx.size() != (size_t)32 * E.off.back() || E.val.size() != E.idx.size()) return 2;
        for (int g = 0; g < ngroups; ++g)
            for (int l = 0; l < 32; ++l) {
                double dot = 0.0;
                for (uint32_t s = E.off[g]; s < E.off[(size_t)g + 1]; ++s) {
                    const int32_t j = E.idx[(size_t)s * 32 + l];
                    if (j < 0 || j >= (int32_t)vec_int.size()) return 3;
                    dot = std::fma(E.val[(size_t)s * 32 + l], vec_int[j], dot);
                }
                if (g * 32 + l < nrows) out_int[(size_t)g * 32 + l] = dot;
                else if (dot != 0.0) return 4;   // padding rows hold zeros only
            }
        return 0;
    };
    auto rel_err = [](const int32_t* ptr, const int32_t* ind, const double* val, const std::vector<double>& v, int r, double got) {
        double ref = 0.0, mag = 0.0;
        for (int32_t q = ptr[r]; q < ptr[r + 1]; ++q) { ref += val[q] * v[ind[q]]; mag += std::fabs(val[q] * v[ind[q]]); }
        return std::fabs(got - ref) / (mag > 0.0 ? mag : 1.0);
    };
    double worst = 0.0;
    {
        std::vector<double> v_user((size_t)n), v_int((size_t)n), out_int((size_t)m, NAN);
        for (int j = 0; j < n; ++j) v_user[j] = 0.25 + (double)((j * 2654435761u) % 1000u) / 997.0;
        for (int k = 0; k < n; ++k) v_int[k] = v_user[orderX[k]];
        const int rc = replay(EA, m, v_int, out_int);
        if (rc != 0) return rc;
        for (int r = 0; r < m; ++r) worst = std::max(worst, rel_err(indptr, indices, values, v_user, r, out_int[posY[r]]));
    }
    {
        std::vector<double> w_user((size_t)m), w_int((size_t)m), out_int((size_t)n, NAN);
        for (int i = 0; i < m; ++i) w_user[i] = 0.25 + (double)((i * 2654435761u) % 1000u) / 997.0;
        for (int k = 0; k < m; ++k) w_int[k] = w_user[orderY[k]];
        const int rc = replay(ET, n, w_int, out_int);
        if (rc != 0) return rc;
        for (int r = 0; r < n; ++r) worst = std::max(worst, rel_err(tptr.data(), tind.data(), tval.data(), w_user, r, out_int[posX[r]]));
    }
    out4[0] = worst;
    out4[1] = (double)EA.off.back();
    out4[2] = (double)ET.off.back();
    out4[3] = nnz > 0 ? 32.0 * ((double)EA.off.back() + (double)ET.off.back()) / (2.0 * (double)nnz) : 1.0;
    return 0;
}

// Host-only statistic of the built format: how many distinct 128 B lines the warp-wide gather
// instructions touch (the L1 tag stage serves about one line per cycle per SM, so this is the
// gather cost model).  out6: [0]/[1] total lines for A / A', [2]/[3] the largest per-CTA sum,
// [4]/[5] total gather instructions (2 per warp-step).
extern "C" int mllp_format_gather_lines(int32_t m, int32_t n, int64_t nnz, const int32_t* indptr,
                                        const int32_t* indices, const double* values, int32_t num_ctas,
                                        int32_t pref_steps, int32_t max_steps, int32_t cluster, double* out6)
{
    using namespace mllp;
    if (m < 0 || n < 0 || !indptr || !out6 || (int64_t)indptr[m] != nnz) return 1001;
    BuildParams bp;
    bp.num_ctas = num_ctas;
    bp.pref_steps = pref_steps;
    bp.max_steps = max_steps < pref_steps ? pref_steps : max_steps;
    bp.cluster = (cluster & 1) != 0;
    bp.contiguous = (cluster & 2) != 0;
    bp.cluster_rounds = ((cluster >> 4) & 7) ? 1 + ((cluster >> 4) & 7) : bp.cluster_rounds;
    std::vector<int32_t> tptr, tind;
    std::vector<double> tval;
    csr_transpose(m, n, indptr, indices, values, tptr, tind, tval);
    std::vector<int32_t> orderY, posY, orderX, posX;
    plan_orders(m, n, indptr, indices, tptr.data(), tind.data(), bp, orderY, posY, orderX, posX);
    HostMat H[2];
    build_host_mat(m, n, indptr, indices, values, orderY, posX, bp, H[0]);
    build_host_mat(n, m, tptr.data(), tind.data(), tval.data(), orderX, posY, bp, H[1]);
    for (int w = 0; w < 2; ++w) {
        const HostMat& M = H[w];
        double total = 0, worst = 0;
        for (int g = 0; g < bp.num_ctas; ++g) {
            double here = 0;
            for (uint32_t st = M.cta_step_begin[g]; st < M.cta_step_begin[g + 1]; ++st)
                for (int half = 0; half < 2; ++half) {
                    int32_t lines[32];
                    int cnt = 0;
                    for (int lane = 0; lane < 32; ++lane) {
                        const int32_t ln = M.idx[((size_t)st * 32 + lane) * 2 + half] >> 4;
                        bool seen = false;
                        for (int q = 0; q < cnt; ++q) seen |= (lines[q] == ln);
                        if (!seen) lines[cnt++] = ln;
                    }
                    here += cnt;
                }
            total += here;
            worst = std::max(worst, here);
        }
        out6[w] = total;
        out6[2 + w] = worst;
        out6[4 + w] = 2.0 * (double)M.total_steps;
        if (cluster & 4) {   // report instead: distinct 32 B sectors per CTA (what L2 serves if L1 captures all reuse)
            double tot_s = 0, worst_s = 0;
            std::vector<int32_t> secs;
            for (int g = 0; g < bp.num_ctas; ++g) {
                secs.clear();
                for (uint32_t st = M.cta_step_begin[g]; st < M.cta_step_begin[g + 1]; ++st)
                    for (int q = 0; q < 64; ++q) secs.push_back(M.idx[(size_t)st * 64 + q] >> 2);
                std::sort(secs.begin(), secs.end());
                const double d = (double)(std::unique(secs.begin(), secs.end()) - secs.begin());
                tot_s += d;
                worst_s = std::max(worst_s, d);
            }
            out6[w] = tot_s;
            out6[2 + w] = worst_s;
        }
    }
    return 0;
}

// Host-only self check of the ROW-PARTITIONED build: emulates all `nranks` ranks on the CPU.  A is partitioned by rows
// (each rank's tiles produce its slice of A v; slices are "exchanged" by writing into the shared padded vector), A' is
// built whole on every rank with its column ids renamed to the padded positions of y (replicated A' phase); both are
// compared with the plain CSR products.  out4: [0] worst relative row error, [1] padded y length, [2] x length,
// [3] largest / mean nonzeros of A per rank.
extern "C" int mllp_rowpart_selfcheck(int32_t m, int32_t n, int64_t nnz, const int32_t* indptr,
                                      const int32_t* indices, const double* values, int32_t num_ctas,
                                      int32_t nranks, double* out4)
{
    using namespace mllp;
    if (m < 0 || n < 0 || !indptr || !out4 || nranks < 1 || (int64_t)indptr[m] != nnz) return 1001;
    BuildParams bp;
    bp.num_ctas = num_ctas;
    bp.pref_steps = 4;
    bp.max_steps = 4;
    std::vector<int32_t> tptr, tind;
    std::vector<double> tval;
    csr_transpose(m, n, indptr, indices, values, tptr, tind, tval);
    std::vector<int32_t> orderY, posY, orderX, posX;
    plan_orders(m, n, indptr, indices, tptr.data(), tind.data(), bp, orderY, posY, orderX, posX);
    const BuildParams bpA = effective_params(m, indptr, bp), bpAT = effective_params(n, tptr.data(), bp);
    RowPartition PY;
    partition_rows(m, indptr, orderY, nranks, PY);
    const int mi = PY.L * nranks, ni = n;
    double worst = 0.0, max_nnz = 0.0;
    auto butterfly = [](double* ls, int L) {
        for (int o = L >> 1; o > 0; o >>= 1) {
            double nxt[32];
            for (int lane = 0; lane < 32; ++lane) nxt[lane] = ls[lane] + ls[lane ^ o];
            for (int lane = 0; lane < 32; ++lane) ls[lane] = nxt[lane];
        }
    };
    // replay of one rank's tile walk of M on v_int; rows land in out_int (must be NaN before), inside [lo, hi)
    auto replay = [&](const HostMat& M, const std::vector<double>& v_int, int nci, std::vector<double>& out_int, uint32_t lo,
                      uint32_t hi) -> int {
        std::vector<double> partial(M.num_partials, 0.0);
        std::vector<uint32_t> arrived(M.splits.size(), 0);
        for (int g = 0; g < bp.num_ctas; ++g) {
            double spart[SPLIT_SLOTS];
            for (uint32_t ti = M.cta_begin[g]; ti < M.cta_begin[g + 1]; ++ti) {
                const Tile& t = M.tiles[ti];
                const int L = 1 << t.logL;
                double lane_sum[32];
                for (int lane = 0; lane < 32; ++lane) {
                    double s = 0.0;
                    for (int st = 0; st < t.nsteps; ++st) {
                        const size_t at = ((size_t)(t.off + st) * 32 + lane) * 2;
                        if (M.idx[at] < 0 || M.idx[at] >= nci || M.idx[at + 1] < 0 || M.idx[at + 1] >= nci) return 3;
                        s = std::fma(M.vals[at], v_int[M.idx[at]], s);
                        s = std::fma(M.vals[at + 1], v_int[M.idx[at + 1]], s);
                    }
                    lane_sum[lane] = s;
                }
                butterfly(lane_sum, L);
                if (t.split < 0) {
                    for (int rr = 0; rr < t.nrows; ++rr) {
                        const uint32_t r = t.row_base + rr;
                        if (r < lo || r >= hi || !std::isnan(out_int[r])) return 5;
                        out_int[r] = lane_sum[rr * L];
                    }
                } else {
                    spart[t.split] = lane_sum[0];
                }
            }
            for (uint32_t li = M.cta_lsplit_begin[g]; li < M.cta_lsplit_begin[g + 1]; ++li) {
                const LocalSplit& ls = M.lsplits[li];
                const SplitRow& sr = M.splits[ls.split_id];
                double pl[32] = {0};
                for (int k = 0; k < ls.count; ++k) pl[k % 32] += spart[ls.first + k];
                butterfly(pl, 32);
                partial[ls.gslot] = pl[0];
                if (++arrived[ls.split_id] == sr.nparts) {
                    double ls32[32] = {0};
                    for (uint32_t k = 0; k < sr.nparts; ++k) ls32[k % 32] += partial[sr.first_slot + k];
                    butterfly(ls32, 32);
                    if (sr.row < lo || sr.row >= hi || !std::isnan(out_int[sr.row])) return 7;
                    out_int[sr.row] = ls32[0];
                }
            }
        }
        return 0;
    };
    auto row_err = [&](const int32_t* ptr, const int32_t* ind, const double* val, const std::vector<double>& v_user, int r,
                       double got) {
        double ref = 0.0, mag = 0.0;
        for (int32_t q = ptr[r]; q < ptr[r + 1]; ++q) {
            ref += val[q] * v_user[ind[q]];
            mag += std::fabs(val[q] * v_user[ind[q]]);
        }
        return std::fabs(ref - got) / (mag > 0.0 ? mag : 1.0);
    };
    {   // A v: every rank produces its slice of the padded result
        std::vector<double> v_user((size_t)n), v_int((size_t)ni), out_int((size_t)mi, NAN);
        for (int j = 0; j < n; ++j) v_user[j] = 0.25 + (double)((j * 2654435761u) % 1000u) / 997.0;
        for (int k = 0; k < ni; ++k) v_int[k] = v_user[orderX[k]];
        for (int rank = 0; rank < nranks; ++rank) {
            HostMat M;
            build_host_mat((int)PY.lists[rank].size(), ni, indptr, indices, values, PY.lists[rank], posX, bpA, M,
                           (uint32_t)(rank * PY.L));
            max_nnz = std::max(max_nnz, (double)M.nnz_emitted);
            const int rc = replay(M, v_int, ni, out_int, (uint32_t)(rank * PY.L), (uint32_t)((rank + 1) * PY.L));
            if (rc != 0) return rc;
        }
        for (int r = 0; r < m; ++r) {
            const double got = out_int[PY.pos[r]];
            if (std::isnan(got)) return 8;
            worst = std::max(worst, row_err(indptr, indices, values, v_user, r, got));
        }
        for (int k = 0; k < mi; ++k)
            if (PY.order_pad[k] < 0 && !std::isnan(out_int[k])) return 9;  // padding must stay untouched
    }
    {   // A' w: the whole product on every rank, gathering from the padded y
        std::vector<double> w_user((size_t)m), w_int((size_t)mi, 0.0), out_int((size_t)n, NAN);
        for (int j = 0; j < m; ++j) w_user[j] = 0.25 + (double)((j * 2654435761u) % 1000u) / 997.0;
        for (int k = 0; k < mi; ++k) if (PY.order_pad[k] >= 0) w_int[k] = w_user[PY.order_pad[k]];
        HostMat M;
        build_host_mat(n, mi, tptr.data(), tind.data(), tval.data(), orderX, PY.pos, bpAT, M);
        const int rc = replay(M, w_int, mi, out_int, 0u, (uint32_t)n);
        if (rc != 0) return rc;
        for (int r = 0; r < n; ++r) {
            const double got = out_int[posX[r]];
            if (std::isnan(got)) return 8;
            worst = std::max(worst, row_err(tptr.data(), tind.data(), tval.data(), w_user, r, got));
        }
    }
    out4[0] = worst;
    out4[1] = mi;
    out4[2] = ni;
    out4[3] = nnz > 0 ? max_nnz / ((double)nnz / nranks) : 1.0;
    return 0;
}
