// Host-side symbolic analysis for the 2x2-block sparse LU used by the HELM
// recursion matrix (replaces the symbolic half of scipy `factorized`,
// reference helmpy/core/helm.py:106).  See DESIGN.md §3.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace helm {

struct Symbolic {
    int N = 0;
    int nnz_adj = 0;
    // numbering: new index r <-> original bus perm[r]; iperm[orig] = r
    std::vector<int> perm, iperm;
    // bus adjacency in the new numbering; entries keep the original row order
    std::vector<int> adj_ptr, adj_idx, adj_src;  // adj_src: index into the caller's nnz arrays
    std::vector<int> adj_slot;                   // LU slot holding block (r, adj_idx)
    // strictly-lower pattern by row, ascending columns; slot of entry e is e
    std::vector<int> lptr, lcol;
    std::vector<int> llev_ptr;                   // rows of L-level l: [llev_ptr[l], llev_ptr[l+1])
    // upper part in U-level order (roots first): position q <-> row urow[q]
    std::vector<int> urow, qof, ulev_ptr, uptr, ucol;
    int nL = 0, nU = 0, nslots = 0;
    std::vector<int> dslot;                      // row -> slot of its (inverted) diagonal block
    // numeric factorization schedule: for L entry e, updates u in [fptr[e], fptr[e+1])
    // do  LU[fdst[u]] -= L_e * LU[fsrc[u]]
    std::vector<int> fptr, fsrc, fdst;
    // fdst_local[u]: position of fdst[u] inside its row's buffer [L entries | D | U entries]
    std::vector<int> fdst_local;
    int max_row_blocks = 0;                      // max over rows of nL(r) + 1 + nU(r)
    // factorization chunks: consecutive rows of one L-level that one warp factorizes together
    // with all their operands staged in shared memory.  Chunk c covers rows
    // [fchunk_row[c], fchunk_row[c+1]); fchunk_lpr[c] = lanes per row (0: diagonal-only rows).
    std::vector<int> fchunk_row, fchunk_lpr;
    std::vector<int> flev_chunk;                 // chunks of L-level l: [flev_chunk[l], flev_chunk[l+1])
    std::vector<int> frow_off;                   // row -> block offset of its buffer inside its chunk
    int cap_upd = 0, cap_ent = 0, cap_rowblk = 0;  // staging capacities (max over chunks)
    // assembly map: slot -> adjacency entry (or -1 for fill), slot -> row
    std::vector<int> slot_src, slot_row, slot_col;

    // flat solve schedule (k_solve_flat): the slots in consumption order cut into waves of <= 32
    // consecutive slots.  wave_tab: 2 ints per wave {first slot, n | flags | maxspan << 9};
    // smeta[slot]: span (lanes after this one in its row piece) | flags | row << 9.
    std::vector<int> wave_tab, smeta;
    // k_solve_cta: scm = {column, meta} per slot; solve_chunks = 4 ints per chunk {first wave, waves, first slot, slots}
    std::vector<int> scm, solve_chunks;
    // k_solve_one: 2 ints per segment {first wave, waves | thin << 30} (see symbolic.cpp)
    std::vector<int> solve_segs;
    static constexpr int ONE_WARPS = 8;
    std::vector<int> one_lists;                  // per-warp wave lists of k_solve_one (see symbolic.cpp)
    int n_waves = 0;                             // waves (wave_tab may carry one padding entry)
    static constexpr int SOLVE_CHUNK_SLOTS = 128;
    int n_waves_l = 0;                           // waves of the forward solve (the rest is backward)
    int n_rows_nol = 0;                          // rows [0, n_rows_nol) have no L entries (forward level 0)
    static constexpr int WAVE_SYNC = 64, WAVE_CONT = 128, WAVE_ISU = 256;          // n in bits 0-5, maxspan in bits 9-13
    static constexpr int WAVE_THIN = 1 << 14;    // the wave is a whole level by itself (ordered before and after)
    static constexpr int WAVE_ONEROW = 1 << 15;  // all blocks of the wave belong to one row (one piece of it)
    static constexpr int META_HEAD = 32, META_ROWSTART = 64, META_ROWEND = 128, META_DIAG = 256;  // span in bits 0-4

    // row-lane solve schedule (k_solve_rows, symbolic.cpp 4b-rows): every row of a level gets a
    // power-of-two group of k lanes; a lane walks every k-th block of its row, at most RL_BLOCKS.
    // rl_desc: 4 ints per (wave, lane): {c0 | c1 << 16, c2 | c3 << 16, fetch word, row | flags}
    // fetch word: first slot | lanes-per-row log2 << 24 | number of blocks << 27 (all a lane needs to
    // request its factor blocks; the kernels read it two waves ahead of the rest of the record);
    // c_t: column of the lane's t-th block (RL_NONE: no block, RL_DIAG: the row's inverted pivot);
    // flags: lanes-per-row log2 in bits 16-18, span (lanes after this one in the row's group) in
    // bits 19-23, then RL_HEAD / RL_SYNC / RL_ISU, shuffle rounds of the wave in bits 27-29.
    // A row of more than 32 * RL_BLOCKS blocks is cut into pieces that run in consecutive one-wave levels,
    // each continuing from the partial result in x[row]; the piece with the pivot comes last.
    // The table ends with two idle waves (look-ahead) that say RL_STOP.
    // Empty (rl_nwaves = 0) when the grid is too large for the packed record (16-bit columns, 24-bit slots).
    std::vector<int> rl_desc;
    int rl_nwaves = 0;
    // k_solve_rows_cta (one CTA of RC_WARPS warps per scenario, x in shared memory): the waves of a level
    // with several waves are dealt to the warps (one SEGMENT, a block barrier after it); a run of
    // consecutive one-wave levels is one segment that warp 0 walks alone.  rc_lists: RC_WARPS + 1 offsets
    // (in entries), then 2 ints per entry {wave, segment}; every warp's list ends with the two idle waves.
    // rc_desc / rc_seg: the records of every warp's waves copied out in the warp's program order (so the
    // look-ahead is a plain sequential read) and the segment of each; warp q's stream starts at record
    // rc_lists[q] and ends with the two idle waves.
    static constexpr int RC_WARPS = 7;
    std::vector<int> rc_desc, rc_seg;
    std::vector<int> rc_lists;
    std::vector<int> rc_segend;                  // per segment: one past the last LU slot it consumes (slots are consumed in storage order)
    int rc_nsegs = 0;
#ifndef HELM_RL_BLOCKS
#define HELM_RL_BLOCKS 3       // measured: 3 blocks per lane fit 64 registers without spills (4 need 72)
#endif
    static constexpr int RL_BLOCKS = HELM_RL_BLOCKS;
    static constexpr int RL_NONE = 0xFFFF, RL_DIAG = 0xFFFE;
    static constexpr int RL_HEAD = 1 << 24, RL_SYNC = 1 << 25, RL_ISU = 1 << 26, RL_STOP = 1 << 30;
    static constexpr int RL_CONT = (int)0x80000000u;           // a further piece of a row too long for one wave: the head lane continues from x[row]

    inline int slotD(int q) const { return nL + uptr[q] + q; }
    inline int slotU(int q, int t) const { return nL + uptr[q] + q + 1 + t; }
};

// adj_ptr/adj_idx: symmetric bus adjacency with the diagonal included.
// Returns false and sets err on malformed input.
// bus_type (may be null): HELM_BUS_* of the base case; buses of one forward level are grouped
// by type so that warps of the per-bus kernels are homogeneous.
bool build_symbolic(int N, const int32_t *adj_ptr, const int32_t *adj_idx, const uint8_t *bus_type,
                    Symbolic &S, std::string &err);

}  // namespace helm
