// Host symbolic analysis: ordering, fill, level sets, schedules.
//
// The HELM recursion matrix (reference helm.py:30-60) is a 2N x 2N real matrix
// made of 2x2 blocks on the pattern of the bus adjacency: block (i,j) exists iff
// buses i and j share a branch.  Treating every block as one entry gives an
// N x N structurally symmetric pattern that is the same for every PV/PQ
// switching state (a PV row keeps a subset of the PQ row, helm.py:50-60), so one
// symbolic analysis serves every Q-limit pass and every scenario.
//
//  1. minimum-degree ordering on the elimination graph (exact external degree,
//     deterministic tie-break on the bus index);
//  2. the filled graph falls out of the elimination (neighbours at elimination
//     time = upper pattern of that pivot);
//  3. buses are renumbered by dependency level of the forward solve (a
//     topological reordering of the elimination tree leaves the fill unchanged),
//     so the rows of one L-level are contiguous;
//  4. level sets of the backward solve, slot numbering in consumption order,
//     row-wise (up-looking) numeric factorization schedule, assembly map.
#include "symbolic.h"

#include <algorithm>
#include <functional>
#include <cstdlib>
#include <queue>
#include <utility>

namespace helm {

namespace {

inline bool sorted_contains(const std::vector<int> &v, int x) {
    return std::binary_search(v.begin(), v.end(), x);
}

inline void sorted_insert(std::vector<int> &v, int x) {
    auto it = std::lower_bound(v.begin(), v.end(), x);
    if (it == v.end() || *it != x) v.insert(it, x);
}

inline void sorted_erase(std::vector<int> &v, int x) {
    auto it = std::lower_bound(v.begin(), v.end(), x);
    if (it != v.end() && *it == x) v.erase(it);
}

}  // namespace

bool build_symbolic(int N, const int32_t *adj_ptr, const int32_t *adj_idx, const uint8_t *bus_type,
                    Symbolic &S, std::string &err) {
    if (N <= 0) {
        err = "n_bus must be positive";
        return false;
    }
    S = Symbolic();
    S.N = N;
    S.nnz_adj = adj_ptr[N];

    // ---- elimination graph (off-diagonal, symmetrised) -------------------------
    std::vector<std::vector<int>> g(N);
    for (int i = 0; i < N; ++i) {
        bool has_diag = false;
        for (int p = adj_ptr[i]; p < adj_ptr[i + 1]; ++p) {
            int j = adj_idx[p];
            if (j < 0 || j >= N) {
                err = "adjacency index out of range";
                return false;
            }
            if (j == i) {
                has_diag = true;
                continue;
            }
            sorted_insert(g[i], j);
            sorted_insert(g[j], i);
        }
        if (!has_diag) {
            err = "adjacency row without its diagonal entry";
            return false;
        }
    }

    // ---- 1+2. minimum degree with explicit fill ---------------------------------
    std::vector<std::vector<int>> higher(N);  // filled-graph neighbours eliminated later
    std::vector<int> pos(N, -1), order;
    order.reserve(N);
    typedef std::pair<int, int> DI;
    std::priority_queue<DI, std::vector<DI>, std::greater<DI>> heap;
    for (int i = 0; i < N; ++i) heap.push(DI((int)g[i].size(), i));
    while (!heap.empty()) {
        DI top = heap.top();
        heap.pop();
        int v = top.second;
        if (pos[v] >= 0 || top.first != (int)g[v].size()) continue;
        pos[v] = (int)order.size();
        order.push_back(v);
        std::vector<int> nb;
        nb.swap(g[v]);
        for (int a : nb) sorted_erase(g[a], v);
        for (size_t x = 0; x < nb.size(); ++x)
            for (size_t y = x + 1; y < nb.size(); ++y) {
                int a = nb[x], b = nb[y];
                if (!sorted_contains(g[a], b)) {
                    sorted_insert(g[a], b);
                    sorted_insert(g[b], a);
                }
            }
        for (int a : nb) heap.push(DI((int)g[a].size(), a));
        higher[v].swap(nb);
    }
    if ((int)order.size() != N) {
        err = "ordering did not visit every bus";
        return false;
    }

    // ---- 3. forward-solve levels and renumbering --------------------------------
    std::vector<int> lev(N, 0), lcount(N, 0);
    for (int k = 0; k < N; ++k) {
        int v = order[k];
        for (int w : higher[v]) {
            lev[w] = std::max(lev[w], lev[v] + 1);
            lcount[w]++;
        }
    }
    // post-order of the elimination tree: rows of one subtree get neighbouring numbers, so the
    // x-gathers of neighbouring rows (lanes) in the triangular solves touch neighbouring addresses
    std::vector<int> dfs(N, 0);
    {
        std::vector<int> parent(N, -1);
        std::vector<std::vector<int>> kids(N);
        std::vector<int> roots;
        for (int v = 0; v < N; ++v) {
            int best = -1;
            for (int w : higher[v])
                if (best < 0 || pos[w] < pos[best]) best = w;
            parent[v] = best;
        }
        for (int k = 0; k < N; ++k) {            // children in elimination order
            int v = order[k];
            if (parent[v] >= 0) kids[parent[v]].push_back(v);
            else roots.push_back(v);
        }
        int counter = 0;
        std::vector<std::pair<int, size_t>> stack;
        for (int r : roots) {
            stack.push_back(std::make_pair(r, (size_t)0));
            while (!stack.empty()) {
                int v = stack.back().first;
                size_t &ci = stack.back().second;
                if (ci < kids[v].size()) {
                    int c = kids[v][ci++];
                    stack.push_back(std::make_pair(c, (size_t)0));
                } else {
                    dfs[v] = counter++;
                    stack.pop_back();
                }
            }
        }
    }
    // within a level and bus type: long rows first (balanced lanes) — measured slightly faster in the
    // solve than pure post-order (HELM_B200_SORT_DFS=1 switches to post-order for A/B runs)
    const bool sort_len = getenv("HELM_B200_SORT_DFS") == nullptr;
    std::vector<int> byl(N);
    for (int i = 0; i < N; ++i) byl[i] = i;
    std::sort(byl.begin(), byl.end(), [&](int a, int b) {
        if (lev[a] != lev[b]) return lev[a] < lev[b];
        if (bus_type && bus_type[a] != bus_type[b]) return bus_type[a] < bus_type[b];  // homogeneous warps
        if (sort_len && lcount[a] != lcount[b]) return lcount[a] > lcount[b];
        return dfs[a] < dfs[b];
    });
    S.perm.assign(N, 0);
    S.iperm.assign(N, 0);
    for (int r = 0; r < N; ++r) {
        S.perm[r] = byl[r];
        S.iperm[byl[r]] = r;
    }
    int nLlev = lev[byl[N - 1]] + 1;
    S.llev_ptr.assign(nLlev + 1, 0);
    for (int r = 0; r < N; ++r) S.llev_ptr[lev[S.perm[r]] + 1]++;
    for (int l = 0; l < nLlev; ++l) S.llev_ptr[l + 1] += S.llev_ptr[l];

    // filled pattern in the new numbering
    std::vector<std::vector<int>> Lrow(N), Urow(N);
    for (int v = 0; v < N; ++v) {
        int rv = S.iperm[v];
        for (int w : higher[v]) {
            int rw = S.iperm[w];
            if (rw <= rv) {
                err = "internal: level renumbering broke the elimination order";
                return false;
            }
            Urow[rv].push_back(rw);
            Lrow[rw].push_back(rv);
        }
    }
    S.lptr.assign(N + 1, 0);
    for (int r = 0; r < N; ++r) {
        std::sort(Lrow[r].begin(), Lrow[r].end());
        std::sort(Urow[r].begin(), Urow[r].end());
        S.lptr[r + 1] = S.lptr[r] + (int)Lrow[r].size();
    }
    S.nL = S.lptr[N];
    S.lcol.resize(S.nL);
    for (int r = 0; r < N; ++r) std::copy(Lrow[r].begin(), Lrow[r].end(), S.lcol.begin() + S.lptr[r]);

    // ---- 4a. backward-solve levels ------------------------------------------------
    std::vector<int> ulev(N, 0);
    int nUlev = 0;
    for (int r = N - 1; r >= 0; --r) {
        int l = 0;
        for (int j : Urow[r]) l = std::max(l, ulev[j] + 1);
        ulev[r] = l;
        nUlev = std::max(nUlev, l + 1);
    }
    S.urow.resize(N);
    for (int i = 0; i < N; ++i) S.urow[i] = i;
    std::sort(S.urow.begin(), S.urow.end(), [&](int a, int b) {
        if (ulev[a] != ulev[b]) return ulev[a] < ulev[b];
        if (sort_len && Urow[a].size() != Urow[b].size()) return Urow[a].size() > Urow[b].size();
        return a < b;
    });
    S.qof.assign(N, 0);
    S.ulev_ptr.assign(nUlev + 1, 0);
    S.uptr.assign(N + 1, 0);
    for (int q = 0; q < N; ++q) {
        int r = S.urow[q];
        S.qof[r] = q;
        S.ulev_ptr[ulev[r] + 1]++;
        S.uptr[q + 1] = S.uptr[q] + (int)Urow[r].size();
    }
    for (int l = 0; l < nUlev; ++l) S.ulev_ptr[l + 1] += S.ulev_ptr[l];
    S.nU = S.uptr[N];
    S.ucol.resize(S.nU);
    for (int q = 0; q < N; ++q) {
        const std::vector<int> &row = Urow[S.urow[q]];
        std::copy(row.begin(), row.end(), S.ucol.begin() + S.uptr[q]);
    }
    S.nslots = S.nL + S.nU + N;
    S.dslot.resize(N);
    for (int r = 0; r < N; ++r) S.dslot[r] = S.slotD(S.qof[r]);

    // slot lookup for block (r, c)
    auto slot_of = [&](int r, int c) -> int {
        if (c == r) return S.dslot[r];
        if (c < r) {
            auto b = S.lcol.begin() + S.lptr[r], e = S.lcol.begin() + S.lptr[r + 1];
            auto it = std::lower_bound(b, e, c);
            if (it == e || *it != c) return -1;
            return (int)(it - S.lcol.begin());
        }
        int q = S.qof[r];
        auto b = S.ucol.begin() + S.uptr[q], e = S.ucol.begin() + S.uptr[q + 1];
        auto it = std::lower_bound(b, e, c);
        if (it == e || *it != c) return -1;
        return S.slotU(q, (int)(it - b));
    };

    // ---- 4b. row-wise numeric factorization schedule -----------------------------
    S.fptr.assign(S.nL + 1, 0);
    for (int r = 0; r < N; ++r)
        for (int e = S.lptr[r]; e < S.lptr[r + 1]; ++e) {
            int k = S.lcol[e];
            const std::vector<int> &uk = Urow[k];
            int qk = S.qof[k];
            for (size_t t = 0; t < uk.size(); ++t) {
                int j = uk[t];
                int dst = slot_of(r, j);
                if (dst < 0) {
                    err = "internal: fill pattern is not closed under elimination";
                    return false;
                }
                S.fsrc.push_back(S.slotU(qk, (int)t));
                S.fdst.push_back(dst);
                int nl = S.lptr[r + 1] - S.lptr[r];
                S.fdst_local.push_back(dst < S.nL ? dst - S.lptr[r] : nl + (dst - S.dslot[r]));
            }
            S.fptr[e + 1] = (int)S.fsrc.size();
        }

    for (int r = 0; r < N; ++r) {
        int q = S.qof[r];
        S.max_row_blocks = std::max(S.max_row_blocks, (S.lptr[r + 1] - S.lptr[r]) + 1 + (S.uptr[q + 1] - S.uptr[q]));
    }

    // ---- 4b'. factorization chunks -----------------------------------------------------------
    {
        // staged 2x2 source blocks per chunk and rows per chunk (tunable for A/B runs)
        const int CAPU_LIMIT = getenv("HELM_B200_CAPU") ? atoi(getenv("HELM_B200_CAPU")) : 512;
        const int ROWS_LIMIT = getenv("HELM_B200_CHUNK_ROWS") ? atoi(getenv("HELM_B200_CHUNK_ROWS")) : 32;
        S.frow_off.assign(N, 0);
        S.fchunk_row.clear();
        S.fchunk_lpr.clear();
        S.flev_chunk.clear();
        for (int l = 0; l < nLlev; ++l) {
            S.flev_chunk.push_back((int)S.fchunk_lpr.size());     // first chunk of level l
            int r0 = S.llev_ptr[l], r1 = S.llev_ptr[l + 1];
            bool diag_only = true;
            for (int r = r0; r < r1; ++r) diag_only = diag_only && (S.lptr[r + 1] == S.lptr[r]);
            if (diag_only) {
                S.fchunk_row.push_back(r0);
                S.fchunk_lpr.push_back(0);
                continue;
            }
            int r = r0;
            while (r < r1) {
                int start = r, upd = 0, ent = 0, blk = 0;
                while (r < r1 && r - start < ROWS_LIMIT) {
                    int u = S.fptr[S.lptr[r + 1]] - S.fptr[S.lptr[r]];
                    if (r > start && upd + u > CAPU_LIMIT) break;
                    int q = S.qof[r];
                    S.frow_off[r] = blk;
                    blk += (S.lptr[r + 1] - S.lptr[r]) + 1 + (S.uptr[q + 1] - S.uptr[q]);
                    upd += u;
                    ent += S.lptr[r + 1] - S.lptr[r];
                    ++r;
                }
                int nrows = r - start, p2 = 1;
                while (p2 < nrows) p2 <<= 1;
                S.fchunk_row.push_back(start);
                S.fchunk_lpr.push_back(std::max(1, std::min(16, 32 / p2)));
                S.cap_upd = std::max(S.cap_upd, upd);
                S.cap_ent = std::max(S.cap_ent, ent);
                S.cap_rowblk = std::max(S.cap_rowblk, blk);
            }
        }
        S.fchunk_row.push_back(N);
        S.flev_chunk.push_back((int)S.fchunk_lpr.size());
    }

    // ---- 4b-flat. flat solve schedule ---------------------------------------------------------
    // The blocks of both triangular solves, in storage (= consumption) order, are cut into
    // "waves" of at most 32 consecutive slots that one warp handles with one block per lane.
    // Rows are never split unless they are longer than 32 blocks; a wave never crosses a level.
    {
        S.smeta.assign(S.nslots, 0);
        S.wave_tab.clear();
        int w_e0 = -1, w_n = 0, w_flags = 0, w_maxspan = 0;
        auto flush = [&]() {
            if (w_n > 0) {
                S.wave_tab.push_back(w_e0);
                S.wave_tab.push_back(w_n | w_flags | (w_maxspan << 9));
            }
            w_e0 = -1; w_n = 0; w_flags &= Symbolic::WAVE_ISU; w_maxspan = 0;
        };
        // one piece of a row: slots [s0, s0 + n) go to the next n lanes of the current wave
        auto put = [&](int s0, int n, int row, bool rowstart, bool rowend, bool diag_first) {
            if (w_n == 0) w_e0 = s0;
            for (int k = 0; k < n; ++k) {
                int span = n - 1 - k;
                int meta = span | (row << 9);
                if (k == 0) meta |= Symbolic::META_HEAD | (rowstart ? Symbolic::META_ROWSTART : 0) |
                                    (rowend ? Symbolic::META_ROWEND : 0);
                if (k == 0 && diag_first) meta |= Symbolic::META_DIAG;
                S.smeta[s0 + k] = meta;
                w_maxspan = std::max(w_maxspan, span);
            }
            w_n += n;
        };
        auto add_row = [&](int s0, int n, int row, bool isU, bool &first_of_level) {
            if (n <= 32 && w_n + n > 32) flush();
            if (n > 32 && w_n > 0) flush();
            int done = 0;
            while (done < n) {
                int take = std::min(32 - w_n, n - done);
                if (first_of_level && w_n == 0) { w_flags |= Symbolic::WAVE_SYNC; first_of_level = false; }
                put(s0 + done, take, row, done == 0, done + take == n, isU && done == 0);
                done += take;
                if (done < n) {                 // the continuation is ordered after this wave; a wave that
                                                // hands on a carry is ordered after the previous one, so
                                                // the single carry slot of k_solve_cta is never contended
                    w_flags |= Symbolic::WAVE_CONT | Symbolic::WAVE_SYNC;
                    flush();
                    w_flags |= Symbolic::WAVE_SYNC;
                }
            }
        };
        for (int l = 0; l < nLlev; ++l) {
            bool first = true;
            for (int r = S.llev_ptr[l]; r < S.llev_ptr[l + 1]; ++r) {
                int n = S.lptr[r + 1] - S.lptr[r];
                if (n > 0) add_row(S.lptr[r], n, r, false, first);
            }
            flush();
        }
        S.n_waves_l = (int)S.wave_tab.size() / 2;
        w_flags = Symbolic::WAVE_ISU;
        for (int l = 0; l < nUlev; ++l) {
            bool first = true;
            for (int q = S.ulev_ptr[l]; q < S.ulev_ptr[l + 1]; ++q)
                add_row(S.slotD(q), 1 + (S.uptr[q + 1] - S.uptr[q]), S.urow[q], true, first);
            flush();
        }
        // chunks of the shared-memory solve (k_solve_cta): consecutive waves covering at most
        // SOLVE_CHUNK_SLOTS slots, copied into a shared-memory ring with one bulk copy each for the
        // factor blocks and for the {column, meta} table
        S.scm.resize((size_t)S.nslots * 2 + 2, 0);
        for (int e = 0; e < S.nslots; ++e) {
            int col;
            if (e < S.nL) col = S.lcol[e];
            else col = -1;
            S.scm[2 * e] = col;
            S.scm[2 * e + 1] = S.smeta[e];
        }
        for (int q = 0; q < N; ++q) {
            S.scm[2 * S.slotD(q)] = S.urow[q];
            for (int t = 0; t < S.uptr[q + 1] - S.uptr[q]; ++t) S.scm[2 * S.slotU(q, t)] = S.ucol[S.uptr[q] + t];
        }
        {
            const int nwv = (int)S.wave_tab.size() / 2;
            S.n_waves = nwv;
            for (int k = 0; k < nwv; ++k) {      // one-wave levels: k_solve_cta runs chains of them in one warp
                bool sync_k = (S.wave_tab[2 * k + 1] & Symbolic::WAVE_SYNC) != 0;
                bool sync_next = (k + 1 >= nwv) || (S.wave_tab[2 * (k + 1) + 1] & Symbolic::WAVE_SYNC) != 0;
                if (sync_k && sync_next) S.wave_tab[2 * k + 1] |= Symbolic::WAVE_THIN;
                // one row piece: the first lane is its head and spans the whole wave
                const int e0 = S.wave_tab[2 * k], n = S.wave_tab[2 * k + 1] & 63;
                if ((S.smeta[e0] & Symbolic::META_HEAD) && (S.smeta[e0] & 31) == n - 1)
                    S.wave_tab[2 * k + 1] |= Symbolic::WAVE_ONEROW;
            }
            int wv = 0;
            while (wv < nwv) {
                int w0 = wv, s0 = S.wave_tab[2 * wv], ns = 0;
                while (wv < nwv && ns + (S.wave_tab[2 * wv + 1] & 63) <= Symbolic::SOLVE_CHUNK_SLOTS) {
                    ns += S.wave_tab[2 * wv + 1] & 63;
                    ++wv;
                }
                S.solve_chunks.push_back(w0);
                S.solve_chunks.push_back(wv - w0);
                S.solve_chunks.push_back(s0);
                S.solve_chunks.push_back(ns);
            }
            // segments of the one-CTA solve (k_solve_one): the waves between two ordering points are
            // independent of each other and are dealt to the warps of the CTA (WIDE segment, a block
            // barrier after it); a run of consecutive one-wave levels is one THIN segment that a single
            // warp walks through with warp barriers only.  2 ints per segment: {first wave, waves | thin << 30}.
            S.solve_segs.clear();
            {
                int k = 0;
                while (k < nwv) {
                    const bool thin = (S.wave_tab[2 * k + 1] & Symbolic::WAVE_THIN) != 0;
                    int e = k + 1;
                    if (thin) {
                        while (e < nwv && (S.wave_tab[2 * e + 1] & Symbolic::WAVE_THIN)) ++e;
                    } else {
                        while (e < nwv && !(S.wave_tab[2 * e + 1] & Symbolic::WAVE_SYNC)) ++e;
                    }
                    S.solve_segs.push_back(k);
                    S.solve_segs.push_back((e - k) | (thin ? (1 << 30) : 0));
                    k = e;
                }
            }
            // per-warp wave lists of k_solve_one (ONE_WARPS warps): the waves of a WIDE segment are dealt
            // round-robin, the waves of a THIN segment alternate between warps 0 and 1.  Layout: [ONE_WARPS + 1 offsets][waves in order].
            {
                const int W = Symbolic::ONE_WARPS;
                std::vector<std::vector<int>> lists(W);
                for (size_t sgi = 0; sgi + 1 < S.solve_segs.size(); sgi += 2) {
                    const int first = S.solve_segs[sgi], n = S.solve_segs[sgi + 1] & 0x3FFFFFFF;
                    const bool thin = (S.solve_segs[sgi + 1] >> 30) != 0;
                    // THIN segments: warps 0 and 1 alternate, so that everything of a wave that does not
                    // depend on x overlaps the dependent part of the wave before it
                    for (int j = 0; j < n; ++j) lists[thin ? (j & 1) : (j % W)].push_back(first + j);
                }
                S.one_lists.assign(W + 1, 0);
                for (int wq = 0; wq < W; ++wq) S.one_lists[wq + 1] = S.one_lists[wq] + (int)lists[wq].size();
                for (int wq = 0; wq < W; ++wq) S.one_lists.insert(S.one_lists.end(), lists[wq].begin(), lists[wq].end());
            }
            if (nwv & 1) { S.wave_tab.push_back(0); S.wave_tab.push_back(0); }   // 16-byte multiple
        }
        // rows without L entries get x = rhs directly: they are the rows of forward level 0 and
        // come first in the numbering
        S.n_rows_nol = 0;
        while (S.n_rows_nol < N && S.lptr[S.n_rows_nol + 1] == S.lptr[S.n_rows_nol]) S.n_rows_nol++;
        for (int r = S.n_rows_nol; r < N; ++r)
            if (S.lptr[r + 1] == S.lptr[r]) {
                err = "internal: rows without L entries are not a prefix of the numbering";
                return false;
            }
    }

    // ---- 4b-rows. row-lane solve schedule -----------------------------------------------------
    // Every row of a level gets a power-of-two group of k lanes, k ~ blocks / RL_BLOCKS, and a lane walks at
    // most RL_BLOCKS (<= 4) blocks of its row, every k-th one (U rows: the inverted pivot is the first block of the row's
    // first lane).  So the blocks of a wave are fetched with up to four 32-byte loads per lane and summed
    // ONCE per wave (segmented over the lane groups) instead of once per 32 blocks, and a level of few
    // long rows still spreads over the warp.  Rows are placed largest group first, which aligns every
    // group to its own size.
    {
        static_assert(Symbolic::RL_BLOCKS >= 1 && Symbolic::RL_BLOCKS <= 4, "the record holds four columns");
        const int RL_T = getenv("HELM_B200_ROWS_T") ? std::max(1, std::min(Symbolic::RL_BLOCKS, atoi(getenv("HELM_B200_ROWS_T")))) : Symbolic::RL_BLOCKS;
        S.rl_desc.clear();
        S.rl_nwaves = 0;
        struct Row { int slot0, n, row, k, diag; const int *cols; bool cont; };
        bool ok = N < Symbolic::RL_DIAG && S.nslots < (1 << 24);
        const int CAP = 32 * RL_T;                        // blocks of a row one wave can take
        auto emit_rows = [&](std::vector<Row> &rows, bool isU) {      // one level (or one round of pieces of long rows)
            for (Row &r : rows) {
                int want = (r.n + RL_T - 1) / RL_T, k = 1;
                while (k < want && k < 32) k <<= 1;
                r.k = k;
            }
            std::stable_sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.k > b.k; });
            size_t i = 0;
            bool first = true;
            while (i < rows.size()) {
                int lanes = 0, maxk = 1;
                const size_t i0 = i;
                while (i < rows.size() && lanes + rows[i].k <= 32) {
                    maxk = std::max(maxk, rows[i].k);
                    lanes += rows[i].k;
                    ++i;
                }
                int rounds = 0;
                while ((1 << rounds) < maxk) ++rounds;
                const int wflags = (first ? Symbolic::RL_SYNC : 0) | (isU ? Symbolic::RL_ISU : 0) | (rounds << 27);
                first = false;
                const size_t d0 = S.rl_desc.size();
                S.rl_desc.resize(d0 + 4 * 32, 0);
                int lane = 0;
                for (size_t j = i0; j < i; ++j) {
                    const Row &r = rows[j];
                    int klog = 0;
                    while ((1 << klog) < r.k) ++klog;
                    for (int sub = 0; sub < r.k; ++sub, ++lane) {
                        int c[4];
                        for (int t = 0; t < 4; ++t) {
                            const int b = sub + t * r.k;           // block of the piece: [pivot |] entries
                            if (b >= r.n) c[t] = Symbolic::RL_NONE;
                            else if (r.diag && b == 0) c[t] = Symbolic::RL_DIAG;
                            else c[t] = r.cols[b - r.diag];
                        }
                        int *d = &S.rl_desc[d0 + 4 * lane];
                        int cnt = 0;
                        while (cnt < 4 && c[cnt] != Symbolic::RL_NONE) ++cnt;
                        d[0] = c[0] | (c[1] << 16);
                        d[1] = c[2] | (c[3] << 16);
                        d[2] = (r.slot0 + sub) | (klog << 24) | (cnt << 27);
                        d[3] = (sub == 0 ? r.row : Symbolic::RL_NONE) | (klog << 16) | ((r.k - 1 - sub) << 19) |
                               (sub == 0 ? Symbolic::RL_HEAD : 0) | wflags | (r.cont ? Symbolic::RL_CONT : 0);
                    }
                }
                for (; lane < 32; ++lane) {            // idle lanes: nothing to load, nothing to store
                    int *d = &S.rl_desc[d0 + 4 * lane];
                    d[0] = d[1] = (int)0xFFFFFFFFu;
                    d[2] = 0;
                    d[3] = Symbolic::RL_NONE | wflags;
                }
            }
        };
        // rows: whole rows {first slot, blocks incl. the pivot if diag, row, -, diag, columns}.  A row longer than CAP
        // is cut: the pieces WITHOUT the pivot first (each a level of its own, the first one in the row's
        // level), the piece that starts with the pivot last.
        auto emit_level = [&](std::vector<Row> &rows, bool isU) {
            std::vector<Row> now, later;
            bool more = false;
            for (const Row &r : rows) {
                if (r.n <= CAP) { now.push_back(r); continue; }
                more = true;
                // tail piece of CAP blocks now, the rest (which keeps the pivot at its front) later
                Row tail = r;
                tail.slot0 = r.slot0 + (r.n - CAP);
                tail.n = CAP;
                tail.cols = r.cols + (r.n - CAP) - r.diag;
                tail.diag = 0;
                now.push_back(tail);
                Row rest = r;
                rest.n = r.n - CAP;
                rest.cont = true;
                later.push_back(rest);
            }
            emit_rows(now, isU);
            if (more) {
                std::function<void(std::vector<Row> &)> again = [&](std::vector<Row> &v) {
                    std::vector<Row> n2, l2;
                    for (const Row &r : v) {
                        if (r.n <= CAP) { n2.push_back(r); continue; }
                        Row tail = r;
                        tail.slot0 = r.slot0 + (r.n - CAP);
                        tail.n = CAP;
                        tail.cols = r.cols + (r.n - CAP) - r.diag;
                        tail.diag = 0;
                        n2.push_back(tail);
                        Row rest = r;
                        rest.n = r.n - CAP;
                        l2.push_back(rest);
                    }
                    emit_rows(n2, isU);
                    if (!l2.empty()) again(l2);
                };
                again(later);
            }
        };
        std::vector<Row> rows;
        for (int l = 0; l < nLlev; ++l) {
            rows.clear();
            for (int r = S.llev_ptr[l]; r < S.llev_ptr[l + 1]; ++r) {
                const int n = S.lptr[r + 1] - S.lptr[r];
                if (n > 0) rows.push_back(Row{S.lptr[r], n, r, 1, 0, S.lcol.data() + S.lptr[r], false});
            }
            emit_level(rows, false);
        }
        for (int l = 0; l < nUlev; ++l) {
            rows.clear();
            for (int q = S.ulev_ptr[l]; q < S.ulev_ptr[l + 1]; ++q)
                rows.push_back(Row{S.slotD(q), 1 + S.uptr[q + 1] - S.uptr[q], S.urow[q], 1, 1, S.ucol.data() + S.uptr[q], false});
            emit_level(rows, true);
        }
        if (ok) {
            S.rl_nwaves = (int)S.rl_desc.size() / (4 * 32);
            S.rl_desc.resize(S.rl_desc.size() + 2 * 4 * 32, 0);      // look-ahead past the last wave: two idle waves
            for (int lane = 0; lane < 64; ++lane) {
                int *d = &S.rl_desc[(size_t)S.rl_nwaves * 4 * 32 + 4 * lane];
                d[0] = d[1] = (int)0xFFFFFFFFu;
                d[3] = Symbolic::RL_NONE | Symbolic::RL_STOP;
            }
        } else {
            S.rl_desc.clear();
        }
        // per-warp programs of the CTA variant
        S.rc_lists.clear();
        S.rc_desc.clear();
        S.rc_seg.clear();
        S.rc_segend.clear();
        S.rc_nsegs = 0;
        if (ok) {
            const int W = Symbolic::RC_WARPS, nw = S.rl_nwaves;
            auto is_sync = [&](int wv) { return wv >= nw || (S.rl_desc[((size_t)wv * 32) * 4 + 3] & Symbolic::RL_SYNC) != 0; };
            std::vector<std::vector<int>> lists(W);
            int wv = 0, seg = 0;
            while (wv < nw) {
                int e = wv + 1;
                while (!is_sync(e)) ++e;                      // waves [wv, e) are one level
                if (e - wv == 1) {                            // a run of one-wave levels
                    while (e < nw && is_sync(e + 1)) ++e;
                    for (int j = wv; j < e; ++j) { lists[0].push_back(j); lists[0].push_back(seg); }
                } else {
                    for (int j = wv; j < e; ++j) { lists[(j - wv) % W].push_back(j); lists[(j - wv) % W].push_back(seg); }
                }
                // last slot of the segment: the largest slot any of its lanes touches
                int last = 0;
                for (int j = wv; j < e; ++j)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int *d = &S.rl_desc[((size_t)j * 32 + lane) * 4];
                        const int k = 1 << ((d[2] >> 24) & 7), cnt = (d[2] >> 27) & 7;
                        if (cnt > 0) last = std::max(last, (d[2] & 0xFFFFFF) + (cnt - 1) * k + 1);
                    }
                S.rc_segend.push_back(std::max(last, S.rc_segend.empty() ? 0 : S.rc_segend.back()));
                wv = e;
                ++seg;
            }
            S.rc_nsegs = seg;
            S.rc_lists.assign(W + 1, 0);
            for (int q = 0; q < W; ++q) {
                lists[q].push_back(nw);
                lists[q].push_back(seg);
                lists[q].push_back(nw + 1);
                lists[q].push_back(seg);
                S.rc_lists[q + 1] = S.rc_lists[q] + (int)lists[q].size() / 2;
            }
            for (int q = 0; q < W; ++q) S.rc_lists.insert(S.rc_lists.end(), lists[q].begin(), lists[q].end());
            S.rc_desc.clear();
            S.rc_seg.clear();
            for (int q = 0; q < W; ++q)
                for (size_t j = 0; j + 1 < lists[q].size(); j += 2) {
                    const int *src = &S.rl_desc[(size_t)lists[q][j] * 4 * 32];
                    S.rc_desc.insert(S.rc_desc.end(), src, src + 4 * 32);
                    S.rc_seg.push_back(lists[q][j + 1]);
                }
            for (int rep = 0; rep < 4; ++rep) {               // the look-ahead of the last warp reads up to eight records further
                std::vector<int> idle(S.rl_desc.begin() + (size_t)nw * 4 * 32, S.rl_desc.begin() + (size_t)(nw + 2) * 4 * 32);
                S.rc_desc.insert(S.rc_desc.end(), idle.begin(), idle.end());
                S.rc_seg.push_back(seg);
                S.rc_seg.push_back(seg);
            }
        }
    }

    // ---- 4c. adjacency in the new numbering + assembly map -------------------------
    S.adj_ptr.assign(N + 1, 0);
    for (int r = 0; r < N; ++r) {
        int o = S.perm[r];
        S.adj_ptr[r + 1] = S.adj_ptr[r] + (adj_ptr[o + 1] - adj_ptr[o]);
    }
    S.adj_idx.resize(S.nnz_adj);
    S.adj_src.resize(S.nnz_adj);
    S.adj_slot.resize(S.nnz_adj);
    S.slot_src.assign(S.nslots, -1);
    S.slot_row.assign(S.nslots, 0);
    S.slot_col.assign(S.nslots, 0);
    for (int r = 0; r < N; ++r) {
        for (int e = S.lptr[r]; e < S.lptr[r + 1]; ++e) { S.slot_row[e] = r; S.slot_col[e] = S.lcol[e]; }
        int q = S.qof[r];
        S.slot_row[S.slotD(q)] = r;
        S.slot_col[S.slotD(q)] = r;
        for (int t = 0; t < S.uptr[q + 1] - S.uptr[q]; ++t) {
            S.slot_row[S.slotU(q, t)] = r;
            S.slot_col[S.slotU(q, t)] = S.ucol[S.uptr[q] + t];
        }
    }
    for (int r = 0; r < N; ++r) {
        int o = S.perm[r];
        int a = S.adj_ptr[r];
        for (int p = adj_ptr[o]; p < adj_ptr[o + 1]; ++p, ++a) {
            int c = S.iperm[adj_idx[p]];
            S.adj_idx[a] = c;
            S.adj_src[a] = p;
            int sl = slot_of(r, c);
            if (sl < 0) {
                err = "internal: adjacency entry missing from the filled pattern";
                return false;
            }
            if (S.slot_src[sl] != -1) {
                err = "duplicate adjacency entry";
                return false;
            }
            S.adj_slot[a] = sl;
            S.slot_src[sl] = a;
        }
    }
    return true;
}

}  // namespace helm
