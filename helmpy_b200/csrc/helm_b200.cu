// B200-native HELM + Pade hot path: CUDA kernels, per-order step machine, C ABI.
// Design notes: DESIGN.md.  Reference being replaced: HELMpy helmpy/core/helm.py
// (cited per kernel below).  sm_100a only; fp64 throughout; no tensor cores
// (nothing on this path is a dense contraction).
//
// Data layout.  Scenarios are packed in lock-step groups of G (1..32; G = 1 is the
// default and the tuned path).  Every per-scenario array is stored [group][item][G],
// so the G scenarios of a group sit in adjacent lanes and every access to a
// per-(item, scenario) value is a contiguous run of G*16 B (values are double2 =
// complex or a 2-vector).  Series histories are tiled, [group][tile of 128 buses]
// [order][128][G] (hx()).  LU block values are [group][slot][G] with 4 doubles (one
// 2x2 block) per element, slots numbered in the order the triangular solves consume
// them, so a solve streams its scenario's factors front to back.
//
// File map: assembly + numeric LU (k_assemble, k_factor*), triangular solves
// (k_solve<G>, k_solve_flat, k_solve_cta), per-order kernels (k_init, k_rhs_conv,
// k_eps, k_advance, k_qlimit, k_pass_end, k_step_begin), variants (helm_variants.cuh),
// Newton-Raphson cross-check (helm_nr.cuh), host side and the C ABI.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/helm_b200.h"
#include "symbolic.h"

namespace {

thread_local std::string g_err;

#define CUDA_TRY(expr)                                                                  \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            g_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
            return HELM_E_CUDA;                                                         \
        }                                                                               \
    } while (0)

// internal per-scenario status (superset of the public codes)
enum : int { ST_ACTIVE = 0, ST_FROZEN = 1, ST_CONVERGED = 2, ST_DIVERGED = 3, ST_SINGULAR = 4, ST_PAD = 5, ST_PASSCAP = 6 };
constexpr int PASS_CAP = 4 * HELM_MAX_PASSES;    // Q-limit passes after which a scenario is given up (HELM_ST_PASS_LIMIT)
// group flags
// PENDING: the group needs a (re)factorization; FACTORING: it is being factorized in the
// side branch of the current step and sits this order out (see the step graph below).
// NEWSCEN: the group's slots were just refilled with fresh scenarios from the queue.
enum : int { GF_PENDING = 1, GF_PASSEND = 2, GF_DONE = 4, GF_FACTORING = 8, GF_NEWSCEN = 16 };
constexpr int MAX_SIDE_DEPTH = 3;

// ---- device-side views ------------------------------------------------------------
struct PlanDev {
    int N, nL, nU, nslots, nLlev, nUlev, nadj, slack;
    const int *adj_ptr, *adj_idx, *adj_slot;
    const double2 *ytr, *yfull, *yshunt;
    const int *ph_ptr, *ph_idx;
    const double2 *ph_val;
    const double *vsp, *pg, *pd, *qd, *qmax, *qmin;
    const unsigned char *btype0;
    const int *perm;  // new index -> original bus
    const int *lptr, *lcol, *llev_ptr;
    const int *urow, *ulev_ptr, *uptr, *ucol, *dslot;
    const int *slot_src, *slot_row, *slot_col;
    double pg_imb0;   // sum(Pd) - sum(Pg) of the unscaled case (classes.py:286)
    // branch table for N-1 outage scenarios (elimination numbering)
    int n_branch;
    const int *br_f, *br_t, *br_a, *br_p;   // br_a[4b..]: adjacency entries (f,t),(t,f),(f,f),(t,t); br_p[2b..]: phase entries
    const double *br_val;                   // [10b..]: y, sh_f, sh_t, ph_ft, ph_tf (complex)
    const int *fptr, *fsrc, *fdst, *fdstl, *qof;
    const int *fchunk_row, *fchunk_lpr, *frow_off, *flev_chunk;
    int use_coop_factor;
    int max_row_blocks, nfchunks, cap_upd, cap_ent, cap_rowblk;
    const int2 *wave_tab;
    const int *smeta;
    int nwaves, n_rows_nol;
    const int *scm;
    const int4 *solve_chunks;
    int nschunks, use_cta_solve, cta_solve_max_groups;
    const int2 *solve_segs;
    const int *one_lists;
    int nsegs, use_one_solve, nonelist;
    const int4 *rl_desc;
    int rl_nwaves, use_rows_solve;
    const int *rc_lists, *rc_segend, *rc_seg;
    const int4 *rc_desc;
    int rc_nsegs, use_rows_cta, use_rows_deep;
};

struct WorkDev {
    int B, G, ngroups, K;  // B = scenarios in the queue; K = history planes = max_coef + 1
    int nbus;
    double mismatch;
    int max_coef, enforce_q;
    double *scale;                    // [Bp]
    int *s_status, *s_mconv, *s_fail, *s_viol, *s_npass, *s_nswitch, *s_ncoef;  // [Bp], ncoef [Bp*MAXP]
    int *g_n, *g_flags;               // [ngroups]
    int *active_groups;               // [1]
    int *queue_next;                  // [1] next scenario of the queue to be given a slot
    int *slot_scen;                   // [Bp] scenario currently held by a slot (-1: none)
    const double *scale_in;           // [B] scenario scales (queue)
    const int *outage_in;             // [B] removed branch per scenario (or null)
    int *o_b;                         // [Bp] removed branch of the scenario in the slot (-1: none)
    // groups being factorized in the current step, compacted by k_step_begin (double-buffered by
    // step parity) so that the side-branch kernels launch only as many CTAs as there is work
    int *fact_list;                   // [MAX_SIDE_DEPTH + 1][ngroups]
    int *fact_count;                  // [MAX_SIDE_DEPTH + 1]
    int *g_fage;                      // [ngroups] steps left until a group being factorized rejoins
    // groups that take a triangular solve in the current step, compacted by k_step_begin (same list
    // index as fact_list): the solve launches one warp per ENTRY, so slots that are idle (being
    // factorized, finished) do not occupy solve warps and more slots than solve warps can be resident
    int *solve_list;                  // [MAX_SIDE_DEPTH + 1][ngroups]
    int *solve_count;                 // [MAX_SIDE_DEPTH + 1]
    // groups whose pass ended in this step (k_advance -> k_qlimit): list in the tail of solve_list's
    // current index is not free, so it has its own array
    int *pend_list;                   // [ngroups]
    int *pend_count;                  // [1], reset by k_step_begin
    int parity;                       // list of the current step: step % (depth + 1) (host-set per launch)
    int depth;                        // steps a factorization may take (side-branch pipeline depth)
    int skip_asm;                     // k_factor_coop: the LU slots already hold the matrix (Newton-Raphson Jacobian)
    int factor_cta;                   // k_factor_coop: the warps of a CTA share one scenario (few scenarios: latency path)
    int *out_status, *out_ncoef, *out_nswitch;   // per-scenario outputs, written when a scenario ends
    // Factor cache (k_fcache): the matrix of a pass depends on the bus types only, and the scenarios of a
    // load-scaling batch run through few distinct type vectors (176-188 per pass among 4096 scenarios of
    // case2869pegase), so the buffers of `lu` are a POOL: lu_ix[slot] = buffer the slot's scenario solves
    // with, found by a hash of its type vector; only the first scenario with a new vector factorizes.
    // null: no cache, slot g owns buffer g.
    int *lu_ix;                       // [ngroups] pool buffer (-1: none / shared first-pass factors)
    int *fc_do;                       // [ngroups] 1: the slot's scenario factorizes its buffer in this side branch
    int *fc_ref, *fc_stamp, *fc_sing, *fc_rdy; // [ngroups] per buffer: scenarios using it, step of its last release, singular matrix, first step its factors may be read
    int *fc_tbuf, *fc_pend, *fc_ctr;  // hash table values [T], pending list [ngroups], counters [16]
    unsigned long long *fc_bkey, *fc_tkey, *fc_hash;   // key of a buffer [ngroups], table keys [T], keys of the step's list [ngroups]
    int fc_tmask;                     // T - 1
    unsigned char *btype;             // [group][N][G]
    double *Qg, *VVant;               // [group][N][G]
    double2 *Pcur;                    // [group][N][G]
    double2 *rhs, *xb;                // [group][N][G]
    double2 *lu;                      // [group][nslots][G][2]
    // factors of the first pass, shared by every scenario without an outage: the matrix of a pass depends
    // on the topology and the bus types only (helm.py:46-60), not on the load scale.  null: not used.
    const double2 *lu_shared;
    double2 *Vh, *Hh, *Eh;            // [group][tile][K][HT][G]
    double2 *Pb;                      // [group][tile][3][HT][G] partial sums of the blocked W recurrence
    double2 *vout;                    // [B][N] original order
    // optional log of the PV->PQ switches (helm(detailed_run_print=True) prints them, helm.py:489):
    // entry k of a slot = {original bus, pass, Q that violated the limit}; k = order of the atomicAdd
    int2 *sw_ent;                     // [Bp][sw_cap] or null
    double *sw_q;                     // [Bp][sw_cap]
    int sw_cap;
    unsigned long long *tl;           // timeline buffer (HELM_TIMELINE builds only), else null
    // variants (helm_variants.cuh): PV bus model 1 and the distributed slack
    int pv_model, dsb_method, dsb_model;
    double *ktot, *sden, *beta, *pcur;   // [slots]
    double *ploss;                       // [slots][K]  Ploss series (coefficients[2N][n])
    double2 *zb;                         // [slots][N]  A^-1 c of the bordered system
    double *Qh;                          // [slots][K][N] Q_i[n] of PV buses (PV model 1)
    double *vre, *qcur;                  // [slots][N]
};

// Optional kernel timeline (compile with -DHELM_TIMELINE): per step and kernel, the first CTA
// start and the last CTA end on the global timer; tl[0] is the step counter.
#ifdef HELM_TIMELINE
#define TL_MAXSTEPS 512
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct TlScope {
    unsigned long long *slot;
    __device__ __forceinline__ TlScope(const unsigned long long *tl_, int kid) {
        unsigned long long *tl = const_cast<unsigned long long *>(tl_);
        slot = nullptr;
        if (tl && threadIdx.x == 0) {
            unsigned long long step = tl[0];
            if (step < TL_MAXSTEPS) {
                slot = tl + 8 + (step * 16 + kid) * 2;
                atomicMin(slot, gtime());
            }
        }
    }
    __device__ __forceinline__ ~TlScope() {
        if (slot) atomicMax(slot + 1, gtime());
    }
};
#define TL_SCOPE(w, kid) TlScope _tl((w).tl, kid)
#define TL_CLK_DECL long long _c0 = clock64(), _c1
#define TL_CLK_ADD(w, idx) do { _c1 = clock64(); if ((w).tl && threadIdx.x == 0) atomicAdd((w).tl + (idx), (unsigned long long)(_c1 - _c0)); _c0 = _c1; } while (0)
#else
#define TL_SCOPE(w, kid)
#define TL_CLK_DECL
#define TL_CLK_ADD(w, idx)
#endif

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 crecip_guard(double2 d) {
    // 1/d with the zero-denominator guard of the epsilon table
    double den = d.x * d.x + d.y * d.y;
    if (!(den > 0.0) || !(den < 1.7e308)) {
        if (den == 0.0) return make_double2(1e120, 0.0);
        if (!(den == den)) return make_double2(den, den);  // NaN propagates -> check fails
        return make_double2(0.0, 0.0);                     // |d| overflowed: 1/d ~ 0
    }
    double inv = 1.0 / den;
    return make_double2(d.x * inv, -d.y * inv);
}

// ---- N-1 outage scenarios: the network of a scenario is the base network minus what one branch
// contributed (classes.py:9-113).  Kernels read the shared base arrays and correct the few
// entries the removed branch touches.
struct OutageRef { int b, f, t; };

template <int G>
__device__ __forceinline__ size_t ix(int grp, int count, int item, int s) {
    return ((size_t)grp * count + item) * G + s;
}
// Series histories are tiled: the K orders of a tile of HT consecutive buses are contiguous,
// [group][tile][order][HT buses][G], so the O(j) loops of a CTA of k_rhs_conv / k_eps (HT threads =
// one tile) walk ONE contiguous run of j * HT * 16 B per history instead of j pieces 16*N B apart
// (DRAM page locality).  HPLANE<G> = distance between consecutive orders of one bus.
constexpr int HT = 128;
template <int G>
__device__ __forceinline__ size_t hx(const WorkDev &w, int N, int grp, int k, int i, int s) {
    const int ntile = (N + HT - 1) / HT;
    return ((((size_t)grp * ntile + (i / HT)) * w.K + k) * HT + (i % HT)) * G + s;
}
template <int G>
__device__ __forceinline__ size_t hplane() { return (size_t)HT * G; }
// partial convolution sums of the blocked W recurrence: three planes per tile, same tiling
template <int G>
__device__ __forceinline__ size_t px(int N, int grp, int k, int i, int s) {
    const int ntile = (N + HT - 1) / HT;
    return ((((size_t)grp * ntile + (i / HT)) * 3 + k) * HT + (i % HT)) * G + s;
}

// G = 1: does the scenario in slot `grp` run on the shared first-pass factors?
__device__ __forceinline__ bool uses_shared_lu(const WorkDev &w, int grp) {
    return w.lu_shared != nullptr && w.s_npass[grp] == 0 && w.o_b[grp] < 0;
}
// G = 1: pool buffer holding the factors of the scenario in slot `grp`
__device__ __forceinline__ int lu_buf(const WorkDev &w, int grp) { return w.lu_ix ? w.lu_ix[grp] : grp; }
__device__ __forceinline__ const double2 *lu_of(const PlanDev &p, const WorkDev &w, int grp) {
    return uses_shared_lu(w, grp) ? w.lu_shared : w.lu + (size_t)lu_buf(w, grp) * p.nslots * 2;
}

__device__ __forceinline__ OutageRef load_outage(const PlanDev &p, const WorkDev &w, int sg) {
    OutageRef o{w.o_b[sg], -1, -1};
    if (o.b >= 0) { o.f = p.br_f[o.b]; o.t = p.br_t[o.b]; }
    return o;
}
__device__ __forceinline__ double2 br_value(const PlanDev &p, int b, int k) {   // k: 0 y, 1 sh_f, 2 sh_t, 3 ph_ft, 4 ph_tf
    return make_double2(p.br_val[(size_t)b * 10 + 2 * k], p.br_val[(size_t)b * 10 + 2 * k + 1]);
}
__device__ __forceinline__ double2 ytr_eff(const PlanDev &p, const OutageRef &o, int a) {
    double2 y = p.ytr[a];
    if (o.b >= 0) {
        const int *ai = p.br_a + (size_t)o.b * 4;
        double2 dy = br_value(p, o.b, 0);
        if (a == ai[0] || a == ai[1]) y = cadd(y, dy);           // off-diagonal was -y
        else if (a == ai[2] || a == ai[3]) y = csub(y, dy);      // diagonal was +y
    }
    return y;
}
__device__ __forceinline__ double2 ysh_eff(const PlanDev &p, const OutageRef &o, int i) {
    double2 y = p.yshunt[i];
    if (o.b >= 0) {
        if (i == o.f) y = csub(y, br_value(p, o.b, 1));
        if (i == o.t) y = csub(y, br_value(p, o.b, 2));
    }
    return y;
}
__device__ __forceinline__ double2 ph_eff(const PlanDev &p, const OutageRef &o, int a) {
    double2 v = p.ph_val[a];
    if (o.b >= 0) {
        if (a == p.br_p[2 * o.b]) v = csub(v, br_value(p, o.b, 3));
        else if (a == p.br_p[2 * o.b + 1]) v = csub(v, br_value(p, o.b, 4));
    }
    return v;
}
// full Y = Ytrans + diag(Yshunt) + phase (classes.py:168-175) of the scenario, adjacency entry a = (i, col)
__device__ __forceinline__ double2 yfull_eff(const PlanDev &p, const OutageRef &o, int a, int i, int col) {
    double2 y = p.yfull[a];
    if (o.b >= 0 && (i == o.f || i == o.t)) {
        const int *ai = p.br_a + (size_t)o.b * 4;
        double2 dy = br_value(p, o.b, 0);
        if (a == ai[0]) y = csub(cadd(y, dy), br_value(p, o.b, 3));
        else if (a == ai[1]) y = csub(cadd(y, dy), br_value(p, o.b, 4));
        else if (a == ai[2] || a == ai[3]) {
            y = csub(y, dy);
            if (col == i) {
                if (i == o.f) y = csub(y, br_value(p, o.b, 1));
                if (i == o.t) y = csub(y, br_value(p, o.b, 2));
            }
        }
    }
    return y;
}

// ---- K_asm: scatter A into the LU slots (reference helm.py:46-60) -------------------
// asm_reset: fresh per-bus state of a slot refilled from the queue (classes.py:230-297);
// asm_slot: the 2x2 block of one LU slot (fill slots get zero blocks).  Both are strided loops so
// that a CTA (k_assemble) or a single warp (k_factor_coop) can run them.
template <int G>
__device__ __forceinline__ void asm_reset(const PlanDev &p, const WorkDev &w, int grp, int tid, int nthr) {
    for (int flat = tid; flat < p.N * G; flat += nthr) {
        int i = flat / G, s = flat % G;
        if (w.s_status[grp * G + s] != ST_ACTIVE) continue;
        size_t bi = ix<G>(grp, p.N, i, s);
        w.btype[bi] = p.btype0[i];
        w.Qg[bi] = 0.0;
        w.VVant[bi] = 0.0;
        w.Pcur[bi] = make_double2(0.0, 0.0);
    }
}
template <int G>
__device__ __forceinline__ void asm_slot(const PlanDev &p, const WorkDev &w, int grp, int slot, int s) {
    int r = p.slot_row[slot];
    int a = p.slot_src[slot];
    unsigned char bt = w.btype[ix<G>(grp, p.N, r, s)];
    bool diag = (p.dslot[r] == slot);
    double2 r0 = make_double2(0.0, 0.0), r1 = make_double2(0.0, 0.0);
    if (bt == HELM_BUS_SLACK) {
        if (diag) { r0.x = 1.0; r1.y = 1.0; }                         // helm.py:47-49
    } else if (a >= 0) {
        double2 y = ytr_eff(p, load_outage(p, w, grp * G + s), a);
        r0 = make_double2(y.x, -y.y);                                 // helm.py:52-53 / 59-60
        if (bt == HELM_BUS_PQ || w.pv_model == 1) r1 = make_double2(y.y, y.x);   // helm.py:50-55
        else if (diag) r1 = make_double2(1.0, 0.0);                   // helm.py:57
    }
    if (w.pv_model == 1 && bt != HELM_BUS_SLACK) {
        // PV model 1: the column of Re V of a PV bus leaves the matrix (helm.py:72-103)
        int c = p.slot_col[slot];
        if (w.btype[ix<G>(grp, p.N, c, s)] == HELM_BUS_PV) {
            r0.x = 0.0;
            r1.x = diag ? 1.0 : 0.0;
        }
    }
    double2 *dst = w.lu + ix<G>(G == 1 ? lu_buf(w, grp) : grp, p.nslots, slot, s) * 2;
    dst[0] = r0;
    dst[1] = r1;
}

template <int G>
__global__ void __launch_bounds__(256) k_assemble(PlanDev p, WorkDev w) {
    TL_SCOPE(w, 0);
  const int nfact = w.fact_count[w.parity];
  for (int fi = blockIdx.x; fi < nfact; fi += gridDim.x) {
    const int grp = w.fact_list[(size_t)w.parity * w.ngroups + fi];
    if (w.g_flags[grp] & GF_NEWSCEN) {
        asm_reset<G>(p, w, grp, threadIdx.x, blockDim.x);
        __syncthreads();
    }
    for (int flat = threadIdx.x; flat < p.nslots * G; flat += blockDim.x) {
        int slot = flat / G, s = flat % G;
        if (w.s_status[grp * G + s] != ST_ACTIVE) continue;
        asm_slot<G>(p, w, grp, slot, s);
    }
    __syncthreads();
  }
}

struct Blk { double a, b, c, d; };  // [[a b][c d]]
__device__ __forceinline__ double bnorm2(const Blk &m) { return m.a * m.a + m.b * m.b + m.c * m.c + m.d * m.d; }
// Static pivoting has no row exchanges to fall back on, so a pivot block is declared singular when
// its smallest singular value (|det| / ||D||) has dropped to 1e-11 of the norm of the block the
// matrix started with (n0 = ||D0||_F^2): islanded buses after an outage leave Schur complements of
// ~1e-13 relative size, well-posed grids stay above ~1e-6.  NaN counts as singular.
__device__ __forceinline__ bool pivot_singular(const Blk &d, double det, double n0) {
    return !(det * det > 1e-22 * bnorm2(d) * n0);
}
__device__ __forceinline__ Blk ldblk(const double2 *lu, size_t e) {
    double2 r0 = lu[e * 2], r1 = lu[e * 2 + 1];
    return Blk{r0.x, r0.y, r1.x, r1.y};
}
__device__ __forceinline__ void stblk(double2 *lu, size_t e, Blk m) {
    lu[e * 2] = make_double2(m.a, m.b);
    lu[e * 2 + 1] = make_double2(m.c, m.d);
}
__device__ __forceinline__ Blk bmul(Blk x, Blk y) {
    return Blk{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}

// ---- K2f: numeric 2x2-block LU on the fixed pattern (replaces SuperLU gstrf, helm.py:106)
// One WARP per group (no block barriers): lane = (row slot, scenario); rows of one
// L-level are independent, levels are separated by __syncwarp().  Each lane
// eliminates whole rows up-looking: L_rk = A_rk * inv(D_k), row_r -= L_rk * U_k.
// D_r is stored inverted.  Many groups are resident per SM, which is what hides
// the dependency latency of the thin levels near the root of the elimination tree.
template <int G, int WPB>
__global__ void __launch_bounds__(WPB * 32) k_factor(PlanDev p, WorkDev w) {
    int grp = blockIdx.x * WPB + (threadIdx.x >> 5);
    if (grp >= w.ngroups) return;
    if (!(w.g_flags[grp] & GF_FACTORING)) return;
    const int R = 32 / G;
    int lane = threadIdx.x & 31;
    int s = lane % G, slot = lane / G;
    int sg = grp * G + s;
    bool active = w.s_status[sg] == ST_ACTIVE;
    double2 *lu = w.lu + (size_t)grp * p.nslots * G * 2;
    bool singular = false;
    for (int l = 0; l < p.nLlev; ++l) {
        int r1 = p.llev_ptr[l + 1];
        if (active) {
            for (int r = p.llev_ptr[l] + slot; r < r1; r += R) {
                int e0 = p.lptr[r], e1 = p.lptr[r + 1];
                const double n0 = bnorm2(ldblk(lu, (size_t)p.dslot[r] * G + s));
                for (int e = e0; e < e1; ++e) {
                    int k = p.lcol[e];
                    int u0 = p.fptr[e], u1 = p.fptr[e + 1];
                    Blk a = ldblk(lu, (size_t)e * G + s);
                    Blk dk = ldblk(lu, (size_t)p.dslot[k] * G + s);
                    Blk m = bmul(a, dk);
                    stblk(lu, (size_t)e * G + s, m);
                    int u = u0;
                    for (; u + 2 <= u1; u += 2) {      // two independent updates in flight
                        int s0 = p.fsrc[u], s1 = p.fsrc[u + 1];
                        size_t d0 = (size_t)p.fdst[u] * G + s, d1 = (size_t)p.fdst[u + 1] * G + s;
                        Blk a0 = ldblk(lu, (size_t)s0 * G + s), a1 = ldblk(lu, (size_t)s1 * G + s);
                        Blk b0 = ldblk(lu, d0), b1 = ldblk(lu, d1);
                        Blk p0 = bmul(m, a0), p1 = bmul(m, a1);
                        b0.a -= p0.a; b0.b -= p0.b; b0.c -= p0.c; b0.d -= p0.d;
                        b1.a -= p1.a; b1.b -= p1.b; b1.c -= p1.c; b1.d -= p1.d;
                        stblk(lu, d0, b0);
                        stblk(lu, d1, b1);
                    }
                    if (u < u1) {
                        Blk src = ldblk(lu, (size_t)p.fsrc[u] * G + s);
                        size_t di = (size_t)p.fdst[u] * G + s;
                        Blk dst = ldblk(lu, di);
                        Blk pr = bmul(m, src);
                        dst.a -= pr.a; dst.b -= pr.b; dst.c -= pr.c; dst.d -= pr.d;
                        stblk(lu, di, dst);
                    }
                }
                size_t dd = (size_t)p.dslot[r] * G + s;
                Blk d = ldblk(lu, dd);
                double det = d.a * d.d - d.b * d.c;
                if (pivot_singular(d, det, n0)) { singular = true; det = 1.0; }
                double inv = 1.0 / det;
                stblk(lu, dd, Blk{d.d * inv, -d.b * inv, -d.c * inv, d.a * inv});
            }
        }
        __syncwarp();
    }
    if (singular) w.s_status[sg] = ST_SINGULAR;
}

// ---- K2f (G = 1): cooperative variant.  Near the root of the elimination tree the levels hold
// one to eight long rows whose eliminations are inherently sequential per row; there a
// sub-warp of 4/8/16 lanes works on one row: the row's blocks are staged in shared memory,
// the updates of one L entry (distinct destinations) are spread over the lanes, and only
// the multiplier chain stays serial.  Wide levels keep one lane per row.
__device__ __forceinline__ Blk ldsm(const double2 *b, int i) {
    double2 r0 = b[2 * i], r1 = b[2 * i + 1];
    return Blk{r0.x, r0.y, r1.x, r1.y};
}
__device__ __forceinline__ void stsm(double2 *b, int i, Blk m) {
    b[2 * i] = make_double2(m.a, m.b);
    b[2 * i + 1] = make_double2(m.c, m.d);
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

#ifndef WPB_FACTOR_N
#define WPB_FACTOR_N 4
#endif
constexpr int WPB_FACTOR = WPB_FACTOR_N;   // warps (= scenarios) per CTA of the staged factorization

// Shared-memory plan of one warp:
// [src blocks cap_upd][Dk blocks cap_ent][row blocks cap_rowblk][dst idx cap_upd][update ptr cap_ent+1]
__host__ __device__ inline size_t factor_smem_per_warp(int cap_upd, int cap_ent, int cap_rowblk) {
    return (size_t)(cap_upd + cap_ent + cap_rowblk) * 32 +
           (size_t)((cap_upd + 3) / 4 * 4 + (cap_ent + 1 + 3) / 4 * 4) * sizeof(int);
}

// One warp factorizes one scenario chunk by chunk (csrc/symbolic.cpp §4b'): a chunk is a run of
// consecutive rows of one L-level.  Phase 1: all 32 lanes stage every operand of the chunk in
// shared memory with independent loads (source U blocks, inverted pivots, destination indices,
// the rows themselves) — this is where the HBM/L2 latency is paid, once per chunk and with full
// memory-level parallelism.  Phase 2: each row is eliminated by a sub-warp of `lpr` lanes from
// shared memory only; the updates of one L entry have distinct destinations and are spread over
// the lanes, only the multiplier chain is serial.  Phase 3: rows are written back.
// one chunk of the factorization of scenario slot `grp` by one warp (see k_factor_coop)
__device__ __forceinline__ void factor_chunk(const PlanDev &p, const WorkDev &w, double2 *lu, int grp, int c, int lane,
                                             double2 *sSrc, double2 *sDk, double2 *sRow, int *sDl, int *sUp, bool &singular) {
        const int r0 = __ldg(p.fchunk_row + c), r1 = __ldg(p.fchunk_row + c + 1);
        const int lpr = __ldg(p.fchunk_lpr + c);
        if (lpr == 0) {                                     // rows without L entries: invert the pivot
            for (int r = r0 + lane; r < r1; r += 32) {
                size_t dd = (size_t)__ldg(p.dslot + r);
                Blk d = ldblk(lu, dd);
                double det = d.a * d.d - d.b * d.c;
                if (pivot_singular(d, det, bnorm2(d))) { singular = true; det = 1.0; }
                double inv = 1.0 / det;
                stblk(lu, dd, Blk{d.d * inv, -d.b * inv, -d.c * inv, d.a * inv});
            }
            __syncwarp();
            return;
        }
        const int e_base = __ldg(p.lptr + r0), e_end = __ldg(p.lptr + r1);
        const int u_base = __ldg(p.fptr + e_base), u_end = __ldg(p.fptr + e_end);
        // ---- phase 1: stage with cp.async (LDGSTS): every copy of the chunk is in flight at once
        for (int u = u_base + lane; u < u_end; u += 128) {
            int uu[4];
            size_t src[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uu[k] = u + 32 * k;
                src[k] = (uu[k] < u_end) ? (size_t)__ldg(p.fsrc + uu[k]) : 0;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (uu[k] < u_end) {
                    cp_async16(sSrc + 2 * (uu[k] - u_base), lu + src[k] * 2);
                    cp_async16(sSrc + 2 * (uu[k] - u_base) + 1, lu + src[k] * 2 + 1);
                    cp_async4(sDl + (uu[k] - u_base), p.fdstl + uu[k]);
                }
        }
        for (int e = e_base + lane; e <= e_end; e += 32) {
            sUp[e - e_base] = __ldg(p.fptr + e) - u_base;
            if (e < e_end) {
                size_t dk = (size_t)__ldg(p.dslot + __ldg(p.lcol + e));
                cp_async16(sDk + 2 * (e - e_base), lu + dk * 2);
                cp_async16(sDk + 2 * (e - e_base) + 1, lu + dk * 2 + 1);
            }
        }
        const int rs = lane / lpr, sl = lane % lpr;
        const int r = r0 + rs;
        const bool have = r < r1;
        int e0 = 0, nl = 0, rowlen = 0, dbase = 0;
        double2 *buf = sRow;
        if (have) {
            e0 = __ldg(p.lptr + r);
            nl = __ldg(p.lptr + r + 1) - e0;
            int q = __ldg(p.qof + r);
            rowlen = nl + 1 + (__ldg(p.uptr + q + 1) - __ldg(p.uptr + q));
            dbase = __ldg(p.dslot + r);
            buf = sRow + (size_t)__ldg(p.frow_off + r) * 2;
            for (int b = sl; b < rowlen; b += lpr) {
                size_t slot = (b < nl) ? (size_t)(e0 + b) : (size_t)(dbase + (b - nl));
                cp_async16(buf + 2 * b, lu + slot * 2);
                cp_async16(buf + 2 * b + 1, lu + slot * 2 + 1);
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        // ---- phase 2: eliminate from shared memory
        if (have) {
            const unsigned mask = (lpr >= 32) ? 0xFFFFFFFFu : (((1u << lpr) - 1u) << (rs * lpr));
            const double n0 = bnorm2(ldsm(buf, nl));         // the pivot block as assembled
            for (int e = 0; e < nl; ++e) {
                Blk m = bmul(ldsm(buf, e), ldsm(sDk, e0 + e - e_base));
                int u0 = sUp[e0 + e - e_base], u1 = sUp[e0 + e + 1 - e_base];
                for (int u = u0 + sl; u < u1; u += lpr) {
                    Blk src = ldsm(sSrc, u);
                    int dl = sDl[u];
                    Blk dst = ldsm(buf, dl);
                    Blk pr = bmul(m, src);
                    dst.a -= pr.a; dst.b -= pr.b; dst.c -= pr.c; dst.d -= pr.d;
                    stsm(buf, dl, dst);
                }
                __syncwarp(mask);
                if (sl == 0) stsm(buf, e, m);
            }
            if (sl == 0) {
                Blk d = ldsm(buf, nl);
                double det = d.a * d.d - d.b * d.c;
                if (pivot_singular(d, det, n0)) { singular = true; det = 1.0; }
                double inv = 1.0 / det;
                stsm(buf, nl, Blk{d.d * inv, -d.b * inv, -d.c * inv, d.a * inv});
            }
            __syncwarp(mask);
            // ---- phase 3: write back
            for (int b = sl; b < rowlen; b += lpr) {
                size_t slot = (b < nl) ? (size_t)(e0 + b) : (size_t)(dbase + (b - nl));
                lu[slot * 2] = buf[2 * b];
                lu[slot * 2 + 1] = buf[2 * b + 1];
            }
        }
        __syncwarp();
}

template <int WPB>
__global__ void __launch_bounds__(WPB * 32) k_factor_coop(PlanDev p, WorkDev w) {
    TL_SCOPE(w, 1);
    extern __shared__ double2 smem_f[];
    const int wib = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nfact = w.fact_count[w.parity];
    char *base = (char *)smem_f + (size_t)wib * factor_smem_per_warp(p.cap_upd, p.cap_ent, p.cap_rowblk);
    double2 *sSrc = (double2 *)base;
    double2 *sDk = sSrc + (size_t)p.cap_upd * 2;
    double2 *sRow = sDk + (size_t)p.cap_ent * 2;
    int *sDl = (int *)(sRow + (size_t)p.cap_rowblk * 2);
    int *sUp = sDl + (p.cap_upd + 3) / 4 * 4;
  if (w.factor_cta) {
    // Few scenarios (the single-case path): the WPB warps of a CTA share ONE scenario.  The chunks of a
    // level are independent of each other and are dealt to the warps; a block barrier per level.
    for (int fi = blockIdx.x; fi < nfact; fi += gridDim.x) {
        const int grp = w.fact_list[(size_t)w.parity * w.ngroups + fi];
        if (w.s_status[grp] != ST_ACTIVE) continue;         // CTA-uniform
        if (!w.skip_asm) {
            if (!w.lu_ix && (w.g_flags[grp] & GF_NEWSCEN)) {   // (with the factor cache: done by k_fchash)
                asm_reset<1>(p, w, grp, threadIdx.x, WPB * 32);
                __syncthreads();
            }
            if (uses_shared_lu(w, grp)) continue;
            if (w.lu_ix && !w.fc_do[grp]) continue;         // another scenario with the same bus types factorizes (or has factorized) the buffer
            for (int slot = threadIdx.x; slot < p.nslots; slot += WPB * 32) asm_slot<1>(p, w, grp, slot, 0);
        }
        __syncthreads();
        double2 *lu = w.lu + (size_t)lu_buf(w, grp) * p.nslots * 2;
        bool singular = false;
        for (int l = 0; l < p.nLlev; ++l) {
            const int c0 = __ldg(p.flev_chunk + l), c1 = __ldg(p.flev_chunk + l + 1);
            if (c1 - c0 == 1 && __ldg(p.fchunk_lpr + c0) == 0) {
                // rows without L entries (level 0, half of the grid): pivots inverted by the whole CTA
                const int r0 = __ldg(p.fchunk_row + c0), r1 = __ldg(p.fchunk_row + c0 + 1);
                for (int r = r0 + (int)threadIdx.x; r < r1; r += WPB * 32) {
                    const size_t dd = (size_t)__ldg(p.dslot + r);
                    const Blk d = ldblk(lu, dd);
                    double det = d.a * d.d - d.b * d.c;
                    if (pivot_singular(d, det, bnorm2(d))) { singular = true; det = 1.0; }
                    const double inv = 1.0 / det;
                    stblk(lu, dd, Blk{d.d * inv, -d.b * inv, -d.c * inv, d.a * inv});
                }
            } else {
                for (int c = c0 + wib; c < c1; c += WPB) factor_chunk(p, w, lu, grp, c, lane, sSrc, sDk, sRow, sDl, sUp, singular);
            }
            __syncthreads();
        }
        if (__syncthreads_or(singular) && threadIdx.x == 0) {
            w.s_status[grp] = ST_SINGULAR;
            if (w.lu_ix) w.fc_sing[w.lu_ix[grp]] = 1;
        }
        __syncthreads();
    }
    return;
  }
  for (int fi = blockIdx.x * WPB + wib; fi < nfact; fi += gridDim.x * WPB) {
    const int grp = w.fact_list[(size_t)w.parity * w.ngroups + fi];
    if (w.s_status[grp] != ST_ACTIVE) continue;             // G = 1: warp-uniform
    // assembly of A by the same warp (no separate kernel in front: the factorization starts at the
    // beginning of the step, while the main branch's solve is being dispatched)
    if (!w.skip_asm) {
        if (!w.lu_ix && (w.g_flags[grp] & GF_NEWSCEN)) {     // (with the factor cache: done by k_fchash)
            asm_reset<1>(p, w, grp, lane, 32);
            __syncwarp();
        }
        if (uses_shared_lu(w, grp)) continue;               // first pass: the shared factors are used
        if (w.lu_ix && !w.fc_do[grp]) continue;             // factor cache: the buffer is (being) factorized by another scenario
        for (int slot = lane; slot < p.nslots; slot += 32) asm_slot<1>(p, w, grp, slot, 0);
        __syncwarp();
    }
    double2 *lu = w.lu + (size_t)lu_buf(w, grp) * p.nslots * 2;
    bool singular = false;
    for (int c = 0; c < p.nfchunks; ++c) factor_chunk(p, w, lu, grp, c, lane, sSrc, sDk, sRow, sDl, sUp, singular);
    if (__any_sync(0xFFFFFFFFu, singular) && lane == 0) {
        w.s_status[grp] = ST_SINGULAR;
        if (w.lu_ix) w.fc_sing[w.lu_ix[grp]] = 1;
    }
    __syncwarp();
  }
}

// ---- K2: level-scheduled block SpTRSV (replaces SuperLU gstrs, helm.py:602) -----------
// One WARP per group.  x = U^-1 L^-1 rhs; the forward phase writes y into xb, the
// backward phase works in place.  Rows of one level are independent: lane =
// (row slot, scenario); levels are separated by __syncwarp().  The factors are
// consumed in storage order (slots are numbered in consumption order).
__device__ __forceinline__ Blk ldblk_stream(const double2 *lu, size_t e) {
    // factors are read exactly once per solve: stream them past L1/L2 residency (ld.global.cs)
    double2 r0 = __ldcs(lu + e * 2), r1 = __ldcs(lu + e * 2 + 1);
    return Blk{r0.x, r0.y, r1.x, r1.y};
}
__device__ __forceinline__ void bfma(double2 &acc, const Blk &m, const double2 &x) {
    acc.x -= m.a * x.x + m.b * x.y;
    acc.y -= m.c * x.x + m.d * x.y;
}

// acc -= sum_t LU[slot0+t] * x[cols[t]] for one row
template <int G>
__device__ __forceinline__ void row_accumulate(const double2 *lu, const double2 *xb, const int *cols,
                                               size_t slot0, int n, int s, double2 &acc) {
    int t = 0;
    for (; t + 4 <= n; t += 4) {                       // 4 entries in flight
        int c0 = __ldg(cols + t), c1 = __ldg(cols + t + 1), c2 = __ldg(cols + t + 2), c3 = __ldg(cols + t + 3);
        Blk m0 = ldblk_stream(lu, (slot0 + t) * G + s), m1 = ldblk_stream(lu, (slot0 + t + 1) * G + s);
        Blk m2 = ldblk_stream(lu, (slot0 + t + 2) * G + s), m3 = ldblk_stream(lu, (slot0 + t + 3) * G + s);
        double2 x0 = xb[(size_t)c0 * G + s], x1 = xb[(size_t)c1 * G + s];
        double2 x2 = xb[(size_t)c2 * G + s], x3 = xb[(size_t)c3 * G + s];
        bfma(acc, m0, x0); bfma(acc, m1, x1); bfma(acc, m2, x2); bfma(acc, m3, x3);
    }
    for (; t < n; ++t) {
        int c = __ldg(cols + t);
        Blk m = ldblk_stream(lu, (slot0 + t) * G + s);
        bfma(acc, m, xb[(size_t)c * G + s]);
    }
}

// the same for two independent rows at once (wide levels: twice the loads in flight per lane)
template <int G>
__device__ __forceinline__ void row_accumulate2(const double2 *lu, const double2 *xb,
                                                const int *ca, size_t sa, int na, double2 &acc_a,
                                                const int *cb, size_t sb, int nb, double2 &acc_b, int s) {
    int t = 0;
    int nmin = na < nb ? na : nb;
    for (; t + 2 <= nmin; t += 2) {
        int a0 = __ldg(ca + t), a1 = __ldg(ca + t + 1), b0 = __ldg(cb + t), b1 = __ldg(cb + t + 1);
        Blk ma0 = ldblk_stream(lu, (sa + t) * G + s), ma1 = ldblk_stream(lu, (sa + t + 1) * G + s);
        Blk mb0 = ldblk_stream(lu, (sb + t) * G + s), mb1 = ldblk_stream(lu, (sb + t + 1) * G + s);
        double2 xa0 = xb[(size_t)a0 * G + s], xa1 = xb[(size_t)a1 * G + s];
        double2 xb0 = xb[(size_t)b0 * G + s], xb1 = xb[(size_t)b1 * G + s];
        bfma(acc_a, ma0, xa0); bfma(acc_a, ma1, xa1);
        bfma(acc_b, mb0, xb0); bfma(acc_b, mb1, xb1);
    }
    row_accumulate<G>(lu, xb, ca + t, sa + t, na - t, s, acc_a);
    row_accumulate<G>(lu, xb, cb + t, sb + t, nb - t, s, acc_b);
}

template <int G, int WPB>
__global__ void __launch_bounds__(WPB * 32) k_solve(PlanDev p, WorkDev w) {
    TL_SCOPE(w, 2);
    int grp = blockIdx.x * WPB + (threadIdx.x >> 5);
    if (grp >= w.ngroups) return;
    if (w.g_flags[grp] & (GF_DONE | GF_FACTORING)) return;
    const int R = 32 / G;
    int lane = threadIdx.x & 31;
    int s = lane % G, slot = lane / G;
    bool active = w.s_status[grp * G + s] == ST_ACTIVE;
    const double2 *lu = w.lu + (size_t)grp * p.nslots * G * 2;
    const double2 *rhs = w.rhs + (size_t)grp * p.N * G;
    double2 *xb = w.xb + (size_t)grp * p.N * G;
    for (int l = 0; l < p.nLlev; ++l) {
        int r1 = __ldg(p.llev_ptr + l + 1);
        if (active) {
            int r = __ldg(p.llev_ptr + l) + slot;
            for (; r + R < r1; r += 2 * R) {           // two rows per lane in flight
                int rb = r + R;
                int e0a = __ldg(p.lptr + r), e1a = __ldg(p.lptr + r + 1);
                int e0b = __ldg(p.lptr + rb), e1b = __ldg(p.lptr + rb + 1);
                double2 acc_a = rhs[(size_t)r * G + s], acc_b = rhs[(size_t)rb * G + s];
                row_accumulate2<G>(lu, xb, p.lcol + e0a, (size_t)e0a, e1a - e0a, acc_a,
                                   p.lcol + e0b, (size_t)e0b, e1b - e0b, acc_b, s);
                xb[(size_t)r * G + s] = acc_a;
                xb[(size_t)rb * G + s] = acc_b;
            }
            if (r < r1) {
                int e0 = __ldg(p.lptr + r), e1 = __ldg(p.lptr + r + 1);
                double2 acc = rhs[(size_t)r * G + s];
                row_accumulate<G>(lu, xb, p.lcol + e0, (size_t)e0, e1 - e0, s, acc);
                xb[(size_t)r * G + s] = acc;
            }
        }
        __syncwarp();
    }
    for (int l = 0; l < p.nUlev; ++l) {
        int q1 = __ldg(p.ulev_ptr + l + 1);
        if (active) {
            int q = __ldg(p.ulev_ptr + l) + slot;
            for (; q + R < q1; q += 2 * R) {
                int qb = q + R;
                int ra = __ldg(p.urow + q), rb = __ldg(p.urow + qb);
                int u0a = __ldg(p.uptr + q), nua = __ldg(p.uptr + q + 1) - u0a;
                int u0b = __ldg(p.uptr + qb), nub = __ldg(p.uptr + qb + 1) - u0b;
                size_t base_a = (size_t)(p.nL + u0a + q), base_b = (size_t)(p.nL + u0b + qb);
                Blk da = ldblk_stream(lu, base_a * G + s), db = ldblk_stream(lu, base_b * G + s);
                double2 acc_a = xb[(size_t)ra * G + s], acc_b = xb[(size_t)rb * G + s];
                row_accumulate2<G>(lu, xb, p.ucol + u0a, base_a + 1, nua, acc_a,
                                   p.ucol + u0b, base_b + 1, nub, acc_b, s);
                xb[(size_t)ra * G + s] = make_double2(da.a * acc_a.x + da.b * acc_a.y, da.c * acc_a.x + da.d * acc_a.y);
                xb[(size_t)rb * G + s] = make_double2(db.a * acc_b.x + db.b * acc_b.y, db.c * acc_b.x + db.d * acc_b.y);
            }
            if (q < q1) {
                int r = __ldg(p.urow + q);
                int u0 = __ldg(p.uptr + q), nu = __ldg(p.uptr + q + 1) - u0;
                size_t base = (size_t)(p.nL + u0 + q);
                Blk d = ldblk_stream(lu, base * G + s);
                double2 acc = xb[(size_t)r * G + s];
                row_accumulate<G>(lu, xb, p.ucol + u0, base + 1, nu, s, acc);
                xb[(size_t)r * G + s] = make_double2(d.a * acc.x + d.b * acc.y, d.c * acc.x + d.d * acc.y);
            }
        }
        __syncwarp();
    }
}

// ---- K2 (G = 1), flat wave solve ----------------------------------------------------------------
// One warp per scenario.  The blocks of L and of [D|U], in storage (= consumption) order, are cut on
// the host into waves of <= 32 consecutive slots (symbolic.cpp 4b-flat): one block per lane, so the
// factor loads of a wave are one contiguous run of up to 1 KB, and which row a lane's block belongs to
// comes from a per-slot table shared by all scenarios.  The row sums are segmented shuffle
// reductions towards the first lane of each row piece.  Nothing but the gathers of x depends on
// earlier results, so the factor blocks, columns and row tables of wave w+2 are already in flight
// while wave w is reduced (three register stages), and the gather of wave w+1 is issued before the
// reduction of wave w whenever both belong to the same level.
struct FlatStage { double2 m0, m1; int col, meta, info; };

// 7 warps per CTA at 64 registers: four CTAs take 57 344 of the 65 536 registers of an SM, i.e.
// 28 scenarios per SM (148 x 28 = 4144 >= the 4096 resident scenarios, one round), and the 8192
// registers left are what a CTA of the concurrent factorization (side branch) needs to start
// while the solves are running instead of after them (timeline r1_h).
constexpr int FLAT_WPB = 7;
constexpr int FLAT_NST = 2;

// opaque copy of a pointer: keeps the 64-bit base in a register pair so that every access is one
// IMAD.WIDE + load instead of a recomputation from the kernel parameters
template <typename T>
__device__ __forceinline__ T *pin_ptr(T *ptr) {
    asm volatile("" : "+l"(ptr));
    return ptr;
}

__device__ __forceinline__ void flat_prefetch(const int2 *wave_tab, const int2 *scm, int nwaves, const double2 *lu,
                                              FlatStage &S, int wp, int lane, int2 &wnext) {
    S.info = 0;
    S.meta = helm::Symbolic::META_DIAG;
    if (wp < nwaves) {
        const int2 wi = wnext;                                   // entry of wave wp, loaded one call ago
        wnext = __ldg(wave_tab + min(wp + 1, nwaves - 1));
        S.info = wi.y;
        const int n = wi.y & 63;
        const int e = wi.x + min(lane, n - 1);                   // idle lanes repeat the last block (and contribute 0)
        S.m0 = __ldcs(lu + 2 * e);
        S.m1 = __ldcs(lu + 2 * e + 1);
        const int2 cm = __ldg(scm + e);
        S.col = cm.x;
        S.meta = (lane < n) ? cm.y : helm::Symbolic::META_DIAG;
    }
}

// gather of x for the lane's block and, on the first lane of a row that starts here, the value
// the row sum is subtracted from (rhs in the forward solve, the forward result in the backward one)
__device__ __forceinline__ void flat_gather(const FlatStage &S, const double2 *rhs, const double2 *xb,
                                            double2 &xg, double2 &bs) {
    using Sy = helm::Symbolic;
    xg = make_double2(0.0, 0.0);
    if (!(S.meta & Sy::META_DIAG)) xg = xb[S.col];
    if ((S.meta & (Sy::META_HEAD | Sy::META_ROWSTART)) == (Sy::META_HEAD | Sy::META_ROWSTART)) {
        const double2 *base = (S.info & Sy::WAVE_ISU) ? xb : rhs;
        bs = base[S.meta >> 9];
    }
}

// v += t on the lanes with o <= span (predicated adds, no selects)
__device__ __forceinline__ void add_if_le(double2 &v, double tx, double ty, int o, int span) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.s32 p, %4, %5;\n\t@p add.f64 %0, %0, %2;\n\t@p add.f64 %1, %1, %3;\n\t}"
        : "+d"(v.x), "+d"(v.y) : "d"(tx), "d"(ty), "r"(o), "r"(span));
}

template <int WPB>
__global__ void __maxnreg__(64) k_solve_flat(PlanDev p, WorkDev w, int side) {
    using Sy = helm::Symbolic;
    TL_SCOPE(w, 2);
    int grp = blockIdx.x * WPB + (threadIdx.x >> 5);
    if (side) {                                              // scenarios being factorized (bordered-system / NR solves)
        if (grp >= w.ngroups) return;
        if (!(w.g_flags[grp] & GF_FACTORING)) return;
    } else {                                                 // one warp per entry of the step's solve list
        if (grp >= w.solve_count[w.parity]) return;
        grp = w.solve_list[(size_t)w.parity * w.ngroups + grp];
    }
    if (w.s_status[grp] != ST_ACTIVE) return;
    const int lane = threadIdx.x & 31;
    const double2 *lu = pin_ptr(lu_of(p, w, grp));
    const double2 *rhs = pin_ptr(w.rhs + (size_t)grp * p.N);
    double2 *xb = pin_ptr(w.xb + (size_t)grp * p.N);
    const int2 *wave_tab = pin_ptr(p.wave_tab);
    const int2 *scm = pin_ptr((const int2 *)p.scm);
    const int nw = p.nwaves;
    int2 wnext = __ldg(wave_tab);
    FlatStage st[FLAT_NST];
#pragma unroll
    for (int s = 0; s < FLAT_NST - 1; ++s) flat_prefetch(wave_tab, scm, nw, lu, st[s], s, lane, wnext);
    // rows without L entries: y = rhs
    for (int r = lane; r < p.n_rows_nol; r += 32) xb[r] = rhs[r];
    const double2 zero = make_double2(0.0, 0.0);
    double2 xg = zero, bs = zero, xgn = zero, bsn = zero;
    double2 carry = zero;
    bool early = false;
    for (int w0 = 0; w0 < nw; w0 += FLAT_NST) {
#pragma unroll
        for (int s = 0; s < FLAT_NST; ++s) {
            const int wv = w0 + s;
            if (wv < nw) {
                FlatStage &S = st[s];
                FlatStage &Sn = st[(s + 1) % FLAT_NST];
                if (!early) {
                    if (S.info & Sy::WAVE_SYNC) __syncwarp();
                    flat_gather(S, rhs, xb, xg, bs);
                }
                // stage of wave wv-1 is free: fetch wave wv + NST - 1 into it
                flat_prefetch(wave_tab, scm, nw, lu, st[(s + FLAT_NST - 1) % FLAT_NST], wv + FLAT_NST - 1, lane, wnext);
                const bool nxt_early = (wv + 1 < nw) && !(Sn.info & Sy::WAVE_SYNC);
                if (nxt_early) flat_gather(Sn, rhs, xb, xgn, bsn);
                // ---- reduce wave wv
                const int meta = S.meta, info = S.info;
                double2 v = make_double2(S.m0.x * xg.x + S.m0.y * xg.y, S.m1.x * xg.x + S.m1.y * xg.y);
                const int span = meta & 31, maxspan = (info >> 9) & 31;
                for (int o = 1; o <= maxspan; o <<= 1) {
                    double tx = __shfl_down_sync(0xFFFFFFFFu, v.x, o);
                    double ty = __shfl_down_sync(0xFFFFFFFFu, v.y, o);
                    add_if_le(v, tx, ty, o, span);
                }
                double2 acc = zero;
                const bool head = (meta & Sy::META_HEAD) != 0;
                if (head) {
                    const bool rs = (meta & Sy::META_ROWSTART) != 0;
                    const double2 b = rs ? bs : carry;
                    acc = make_double2(b.x - v.x, b.y - v.y);
                    if (meta & Sy::META_ROWEND) {
                        double2 out = acc;
                        if (info & Sy::WAVE_ISU) {
                            double2 d0 = S.m0, d1 = S.m1;
                            if (!rs) {                       // row longer than a wave: its pivot came earlier
                                const int ds = __ldg(p.dslot + (meta >> 9));
                                d0 = lu[2 * ds]; d1 = lu[2 * ds + 1];
                            }
                            out = make_double2(d0.x * acc.x + d0.y * acc.y, d1.x * acc.x + d1.y * acc.y);
                        }
                        xb[meta >> 9] = out;
                    }
                }
                if (info & Sy::WAVE_CONT) {                 // the last row continues in the next wave
                    const unsigned bal = __ballot_sync(0xFFFFFFFFu, head && !(meta & Sy::META_ROWEND));
                    const int src = __ffs(bal) - 1;
                    carry.x = __shfl_sync(0xFFFFFFFFu, acc.x, src);
                    carry.y = __shfl_sync(0xFFFFFFFFu, acc.y, src);
                }
                early = nxt_early;
                xg = xgn; bs = bsn;
            }
        }
    }
}

// ---- K2 (G = 1, large batches), row-lane solve -----------------------------------------------------
// One warp per scenario, as k_solve_flat, but the work of a wave is laid out by ROWS (symbolic.cpp
// 4b-rows): every row of a level owns a power-of-two group of k lanes and a lane walks every k-th block
// of its row, at most three of them.  A wave so covers up to 96 blocks with three 32-byte loads, three
// gathers and six DFMA pairs per lane, and pays ONE segmented sum over the lane groups (none at all
// in the wide levels, where k = 1) -- the wave solve pays a five-round segmented sum and its bookkeeping
// for every 32 blocks, which is what bounds it.  Everything a lane needs to know about
// a wave is one 16-byte record shared by all scenarios and fetched one wave ahead: four 16-bit columns,
// the first slot, and {row, lanes per row, span, flags}.  The factor blocks of wave w+1 do not depend on
// x: they are requested right after the products of wave w, before its sum and its store.
struct __align__(32) Blk256 { double a, b, c, d; };
// STREAM: evict-first (large batches read every private factor block once per solve); otherwise the
// default policy, so that the factors of a few scenarios stay in L2 from one order to the next
template <bool STREAM = true>
__device__ __forceinline__ Blk256 ld_blk_stream(const double2 *lu, int slot) {
    Blk256 r;
    if (STREAM)
        asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(lu + 2 * (size_t)slot));
    else
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(lu + 2 * (size_t)slot));
    return r;
}

#ifndef ROWS_STREAM
#define ROWS_STREAM 1      // evict-first hint on the factor loads of k_solve_rows
#endif
#ifndef ROWS_REGS
#define ROWS_REGS 64       // 28 warps per SM and room for a factorization CTA beside them, as k_solve_flat
#endif
#define ROWS_NB (helm::Symbolic::RL_BLOCKS)      // blocks per lane and wave
#ifndef ROWS_F2
#define ROWS_F2 0          // fetch word two waves ahead: measured slower (226 -> 236-251 ms per 16 384 scenarios), kept for A/B builds
#endif
#ifndef ROWS_PIN
#define ROWS_PIN 1
#endif
#ifndef ROWS_UNROLL
#define ROWS_UNROLL 2
#endif
#ifndef ROWS_WPB_N
#define ROWS_WPB_N 7
#endif
constexpr int ROWS_WPB = ROWS_WPB_N;
constexpr int ROWS_UNROLL_N = ROWS_UNROLL;

// blocks t = T0 .. of the lane, from the record's fetch word (slot | lanes-per-row log2 << 24 | blocks << 27)
template <int T0, bool STREAM = true>
__device__ __forceinline__ void rows_fetch(const double2 *lu, const int f, Blk256 (&m)[4]) {
    const int k = 1 << ((f >> 24) & 7), cnt = (f >> 27) & 7, slot = f & 0xFFFFFF;
    if (T0 == 0 && cnt > 0) m[0] = ld_blk_stream<STREAM>(lu, slot);
    if (cnt > 1) m[1] = ld_blk_stream<STREAM>(lu, slot + k);
    if (cnt > 2) m[2] = ld_blk_stream<STREAM>(lu, slot + 2 * k);
    if (ROWS_NB > 3 && cnt > 3) m[3] = ld_blk_stream<STREAM>(lu, slot + 3 * k);
}

// one wave of k_solve_rows: record d, its blocks in m; fn = fetch word of the NEXT wave, whose blocks replace m
// (m[0] last: it may be this wave's pivot)
__device__ __forceinline__ void rows_wave(const double2 *lu, const double2 *rhs, double2 *xb, const int4 d, const int fn, Blk256 (&m)[4]) {
    using Sy = helm::Symbolic;
    if (d.w & Sy::RL_SYNC) __syncwarp();
    const unsigned c0 = (unsigned)d.x & 0xFFFFu, c1 = (unsigned)d.x >> 16, c2 = (unsigned)d.y & 0xFFFFu, c3 = (unsigned)d.y >> 16;
    const bool isu = (d.w & Sy::RL_ISU) != 0, head = (d.w & Sy::RL_HEAD) != 0;
    const int row = d.w & 0xFFFF;
    double2 x0 = make_double2(0.0, 0.0), x1 = x0, x2 = x0, x3 = x0;
    // v = (first lane of a row: what the row sum is subtracted from) - sum of the lane's blocks
    double2 v = x0;
    if (c0 < Sy::RL_DIAG) x0 = xb[c0];
    if (c1 != Sy::RL_NONE) x1 = xb[c1];
    if (c2 != Sy::RL_NONE) x2 = xb[c2];
    if (ROWS_NB > 3 && c3 != Sy::RL_NONE) x3 = xb[c3];
    if (head) v = (isu || d.w < 0) ? xb[row] : __ldcs(rhs + row);   // d.w < 0: RL_CONT, a later piece of a long row
    if (c0 < Sy::RL_DIAG) { v.x -= m[0].a * x0.x + m[0].b * x0.y; v.y -= m[0].c * x0.x + m[0].d * x0.y; }
    if (c1 != Sy::RL_NONE) { v.x -= m[1].a * x1.x + m[1].b * x1.y; v.y -= m[1].c * x1.x + m[1].d * x1.y; }
    if (c2 != Sy::RL_NONE) { v.x -= m[2].a * x2.x + m[2].b * x2.y; v.y -= m[2].c * x2.x + m[2].d * x2.y; }
    if (ROWS_NB > 3 && c3 != Sy::RL_NONE) { v.x -= m[3].a * x3.x + m[3].b * x3.y; v.y -= m[3].c * x3.x + m[3].d * x3.y; }
    rows_fetch<1, ROWS_STREAM != 0>(lu, fn, m);
    const int span = (d.w >> 19) & 31, rounds = (d.w >> 27) & 7;
    for (int r = 0; r < rounds; ++r) {
        const int o = 1 << r;
        const double tx = __shfl_down_sync(0xFFFFFFFFu, v.x, o);
        const double ty = __shfl_down_sync(0xFFFFFFFFu, v.y, o);
        add_if_le(v, tx, ty, o, span);
    }
    if (head) {
        // U rows: the head lane's first block is the inverted pivot (of a long row: in its last piece only)
        if (c0 == Sy::RL_DIAG) v = make_double2(m[0].a * v.x + m[0].b * v.y, m[0].c * v.x + m[0].d * v.y);
        xb[row] = v;
    }
    if (fn >> 27) m[0] = ld_blk_stream<ROWS_STREAM != 0>(lu, fn & 0xFFFFFF);
}

template <int WPB>
__global__ void __maxnreg__(ROWS_REGS) k_solve_rows(PlanDev p, WorkDev w) {
    using Sy = helm::Symbolic;
    TL_SCOPE(w, 2);
    int grp = blockIdx.x * WPB + (threadIdx.x >> 5);
    if (grp >= w.solve_count[w.parity]) return;              // one warp per entry of the step's solve list
    grp = w.solve_list[(size_t)w.parity * w.ngroups + grp];
    if (w.s_status[grp] != ST_ACTIVE) return;
    const int lane = threadIdx.x & 31;
#if ROWS_PIN
    const double2 *lu = pin_ptr(lu_of(p, w, grp));
    const double2 *rhs = pin_ptr(w.rhs + (size_t)grp * p.N);
    double2 *xb = pin_ptr(w.xb + (size_t)grp * p.N);
    const int4 *desc = pin_ptr(p.rl_desc + lane);
    __builtin_assume(__isGlobal(lu));
    __builtin_assume(__isGlobal(rhs));
    __builtin_assume(__isGlobal(xb));
    __builtin_assume(__isGlobal(desc));
#else
    const double2 *lu = lu_of(p, w, grp);
    const double2 *rhs = w.rhs + (size_t)grp * p.N;
    double2 *xb = w.xb + (size_t)grp * p.N;
    const int4 *desc = p.rl_desc + lane;
#endif
    int4 da = __ldg(desc), db;
    Blk256 m[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) m[t] = Blk256{0.0, 0.0, 0.0, 0.0};
    rows_fetch<0, ROWS_STREAM != 0>(lu, da.z, m);
    // (L2 evict-last hints on the accesses to x: 288 -> 283 ms per 16 384 scenarios, dropped)
    for (int r = lane; r < p.n_rows_nol; r += 32) xb[r] = __ldcs(rhs + r);     // rows without L entries: y = rhs
#if ROWS_F2
    // The fetch word is read TWO waves ahead, so that the block requests of the next wave never wait for
    // a record (one L2 round trip less on every wave's chain: with the factor cache the solve is bound by
    // L2 latency, not by DRAM); two waves per trip, so that no look-ahead value is copied between
    // registers (a copy waits for the load it copies).  The table ends with two idle waves that say stop.
    int fa, fb = __ldg(&desc[32].z);
    while (true) {
        if (da.w & Sy::RL_STOP) break;
        db = __ldg(desc + 32); fa = __ldg(&desc[64].z);
        rows_wave(lu, rhs, xb, da, fb, m);
        if (db.w & Sy::RL_STOP) break;
        desc += 64;
        da = __ldg(desc); fb = __ldg(&desc[32].z);
        rows_wave(lu, rhs, xb, db, fa, m);
    }
#else
#pragma unroll ROWS_UNROLL_N
    while (!(da.w & Sy::RL_STOP)) {                                   // the table ends with idle waves that say stop
        desc += 32;
        db = __ldg(desc);
        rows_wave(lu, rhs, xb, da, db.z, m);
        da = db;
    }
#endif
}

// ---- K2 (G = 1), row-lane solve with x in shared memory ---------------------------------------------
// One CTA of RC_WARPS warps per scenario; x (16 B per bus) lives in shared memory, so the gathers, the
// row stores and the ordering between levels never leave the SM and a solve moves nothing but its
// factors, rhs in and x out (k_solve_rows re-reads 2/3 of its gathers from DRAM: the x of the resident
// scenarios, 188 MB, does not stay in L2 beside the factor stream).  Same packed records as
// k_solve_rows; the host deals the waves of a level with several waves to the warps (a SEGMENT, one
// block barrier after it) and gives a run of one-wave levels to warp 0, which walks it with warp barriers
// only (symbolic.cpp, rc_lists).  A warp fetches the record and the factor blocks of its NEXT wave before
// it waits for anything, and one thread keeps an L2 prefetch of the scenario's factor stream a fixed
// distance ahead of the segment being solved.
constexpr int RC_WARPS = helm::Symbolic::RC_WARPS;
constexpr int RC_T = RC_WARPS * 32;
#ifndef RC_PF_SLOTS
#define RC_PF_SLOTS 2048      // L2 prefetch distance of the factor stream, in 32-byte slots
#endif

__device__ __forceinline__ void bulk_prefetch_l2(const void *gsrc, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// one wave of a warp of k_solve_rows_cta: record d, blocks m (m[0] of the NEXT wave is requested at the end,
// the others right after the products), fn = fetch word of the next wave
// (an L1 prefetch of the blocks two waves ahead was measured slower: 115 vs 98 us per solve alone)
template <bool STREAM>
__device__ __forceinline__ void rc_wave(const double2 *lu, double2 *xs, const int4 d, int fn, Blk256 (&m)[4], const int abl = 0) {
    using Sy = helm::Symbolic;
    if (abl & 2) fn &= 0xFFFFFF;                                      // timing ablation: no factor loads
    if (d.w & Sy::RL_SYNC) __syncwarp();                              // a run of one-wave levels: ordered inside the warp
    const unsigned c0 = (unsigned)d.x & 0xFFFFu, c1 = (unsigned)d.x >> 16, c2 = (unsigned)d.y & 0xFFFFu, c3 = (unsigned)d.y >> 16;
    const bool head = (d.w & Sy::RL_HEAD) != 0;
    const int row = d.w & 0xFFFF;
    double2 x0 = make_double2(0.0, 0.0), x1 = x0, x2 = x0, x3 = x0;
    double2 v = x0;
    if (!(abl & 4)) {
    if (c0 < Sy::RL_DIAG) x0 = xs[c0];
    if (c1 != Sy::RL_NONE) x1 = xs[c1];
    if (c2 != Sy::RL_NONE) x2 = xs[c2];
    if (ROWS_NB > 3 && c3 != Sy::RL_NONE) x3 = xs[c3];
    }
    if (head) v = xs[row];                                            // rhs (forward) or y (backward), solved in place
    if (c0 < Sy::RL_DIAG) { v.x -= m[0].a * x0.x + m[0].b * x0.y; v.y -= m[0].c * x0.x + m[0].d * x0.y; }
    if (c1 != Sy::RL_NONE) { v.x -= m[1].a * x1.x + m[1].b * x1.y; v.y -= m[1].c * x1.x + m[1].d * x1.y; }
    if (c2 != Sy::RL_NONE) { v.x -= m[2].a * x2.x + m[2].b * x2.y; v.y -= m[2].c * x2.x + m[2].d * x2.y; }
    if (ROWS_NB > 3 && c3 != Sy::RL_NONE) { v.x -= m[3].a * x3.x + m[3].b * x3.y; v.y -= m[3].c * x3.x + m[3].d * x3.y; }
    rows_fetch<1, STREAM>(lu, fn, m);                                 // m[0] may still be this wave's pivot
    const int span = (d.w >> 19) & 31, rounds = (abl & 1) ? 0 : (d.w >> 27) & 7;
    for (int r = 0; r < rounds; ++r) {
        const int o = 1 << r;
        const double tx = __shfl_down_sync(0xFFFFFFFFu, v.x, o);
        const double ty = __shfl_down_sync(0xFFFFFFFFu, v.y, o);
        add_if_le(v, tx, ty, o, span);
    }
    if (head) {
        if (c0 == Sy::RL_DIAG) v = make_double2(m[0].a * v.x + m[0].b * v.y, m[0].c * v.x + m[0].d * v.y);
        xs[row] = v;
    }
    if (fn >> 27) m[0] = ld_blk_stream<STREAM>(lu, fn & 0xFFFFFF);
}

template <bool STREAM>
__global__ void __launch_bounds__(RC_T, 4) k_solve_rows_cta(PlanDev p, WorkDev w, int mode) {
    using Sy = helm::Symbolic;
    TL_SCOPE(w, 2);
    extern __shared__ __align__(16) double2 xs_rc[];
    const int abl = mode >> 8;                               // timing ablations (helm_debug_solve_time); 0 in production
    mode &= 255;
    int grp;
    if (mode == 2) {                                         // one CTA per entry of the step's solve list
        if ((int)blockIdx.x >= w.solve_count[w.parity]) return;
        grp = w.solve_list[(size_t)w.parity * w.ngroups + blockIdx.x];
    } else {
        grp = blockIdx.x;
        const int gfl = w.g_flags[grp];
        if (mode ? !(gfl & GF_FACTORING) : (gfl & (GF_DONE | GF_FACTORING)) != 0) return;
    }
    if (w.s_status[grp] != ST_ACTIVE) return;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int N = p.N, nsegs = p.rc_nsegs;
    double2 *xs = xs_rc;
    const double2 *lu = lu_of(p, w, grp);
    const double2 *rhs = w.rhs + (size_t)grp * N;
    double2 *xb = w.xb + (size_t)grp * N;
    const bool stream = STREAM && lu != w.lu_shared;         // the shared first-pass factors sit in L2 anyway
    int pf = 0;                                              // prefetch frontier (slot)
    if (tid == 0 && stream) {
        pf = min(RC_PF_SLOTS, p.nslots);
        bulk_prefetch_l2(lu, (unsigned)pf * 32u);
    }
    // the warp's program: its records in order, ending with two idle waves that say stop
    const int first = __ldg(p.rc_lists + wp);
    const int4 *desc = p.rc_desc + (size_t)first * 32 + lane;
    const int *sgp = p.rc_seg + first;
    int4 da = __ldg(desc), db;
    int fa, fb = __ldg(&desc[32].z);                                  // fetch word of the next wave
    int sa = __ldg(sgp), sb;
    Blk256 m[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) m[t] = Blk256{0.0, 0.0, 0.0, 0.0};
    rows_fetch<0, STREAM>(lu, da.z, m);
    for (int i = tid; i < N; i += RC_T) xs[i] = __ldcs(rhs + i);      // rows without L entries: y = rhs
    __syncthreads();
    int seg = 0;
    auto barriers = [&](int upto) {                                   // block barriers up to the wave's segment
        while (seg < upto) {
            if (!(abl & 8)) __syncthreads();
            ++seg;
            if (tid == 0 && stream) {
                const int end = min(__ldg(p.rc_segend + min(seg, nsegs - 1)) + RC_PF_SLOTS, p.nslots);
                if (end > pf) { bulk_prefetch_l2(lu + 2 * (size_t)pf, (unsigned)(end - pf) * 32u); pf = end; }
            }
        }
    };
    // two waves per trip, so that no look-ahead value is ever copied from one register to another
    // (a copy waits for the load it copies)
    while (true) {
        if (da.w & Sy::RL_STOP) break;
        db = __ldg(desc + 32); fa = __ldg(&desc[64].z); sb = __ldg(sgp + 1);
        barriers(sa);
        if (!(abl & 16)) rc_wave<STREAM>(lu, xs, da, fb, m, abl);
        if (db.w & Sy::RL_STOP) break;
        desc += 64; sgp += 2;
        da = __ldg(desc); fb = __ldg(&desc[32].z); sa = __ldg(sgp);
        barriers(sb);
        if (!(abl & 16)) rc_wave<STREAM>(lu, xs, db, fa, m, abl);
    }
    barriers(nsegs);                                                  // every warp passes every barrier
    for (int i = tid; i < N; i += RC_T) xb[i] = xs[i];
}

// Latency variant for small batches and the single-case path (one CTA per SM at most, registers are
// free): the records of a warp's next FOUR waves and the factor blocks of its next TWO waves are in
// registers, so nothing on the dependent chain of a run of one-wave levels waits for a table or a factor
// (ablation of k_solve_rows_cta alone, tools/solve_time.py: walking the records one wave ahead costs 31 of
// 98 us, the factor requests one wave ahead 16).
struct RcDeep {
    const double2 *lu;
    double2 *xs;
};
__device__ __forceinline__ void rcd_fetch_rest(const double2 *lu, const int f, Blk256 (&m)[4]) { rows_fetch<1, false>(lu, f, m); }

// one wave: record d, its blocks in m; f2 = fetch word two waves ahead, whose blocks replace m
__device__ __forceinline__ void rcd_wave(const RcDeep &c, const int4 d, const int f2, Blk256 (&m)[4]) {
    using Sy = helm::Symbolic;
    double2 *xs = c.xs;
    if (d.w & Sy::RL_SYNC) __syncwarp();
    const unsigned c0 = (unsigned)d.x & 0xFFFFu, c1 = (unsigned)d.x >> 16, c2 = (unsigned)d.y & 0xFFFFu, c3 = (unsigned)d.y >> 16;
    const bool head = (d.w & Sy::RL_HEAD) != 0;
    const int row = d.w & 0xFFFF;
    double2 x0 = make_double2(0.0, 0.0), x1 = x0, x2 = x0, x3 = x0;
    double2 v = x0;
    if (c0 < Sy::RL_DIAG) x0 = xs[c0];
    if (c1 != Sy::RL_NONE) x1 = xs[c1];
    if (c2 != Sy::RL_NONE) x2 = xs[c2];
    if (ROWS_NB > 3 && c3 != Sy::RL_NONE) x3 = xs[c3];
    if (head) v = xs[row];
    if (c0 < Sy::RL_DIAG) { v.x -= m[0].a * x0.x + m[0].b * x0.y; v.y -= m[0].c * x0.x + m[0].d * x0.y; }
    if (c1 != Sy::RL_NONE) { v.x -= m[1].a * x1.x + m[1].b * x1.y; v.y -= m[1].c * x1.x + m[1].d * x1.y; }
    if (c2 != Sy::RL_NONE) { v.x -= m[2].a * x2.x + m[2].b * x2.y; v.y -= m[2].c * x2.x + m[2].d * x2.y; }
    if (ROWS_NB > 3 && c3 != Sy::RL_NONE) { v.x -= m[3].a * x3.x + m[3].b * x3.y; v.y -= m[3].c * x3.x + m[3].d * x3.y; }
    rcd_fetch_rest(c.lu, f2, m);
    const int span = (d.w >> 19) & 31, rounds = (d.w >> 27) & 7;
    for (int r = 0; r < rounds; ++r) {
        const int o = 1 << r;
        const double tx = __shfl_down_sync(0xFFFFFFFFu, v.x, o);
        const double ty = __shfl_down_sync(0xFFFFFFFFu, v.y, o);
        add_if_le(v, tx, ty, o, span);
    }
    if (head) {
        if (c0 == Sy::RL_DIAG) v = make_double2(m[0].a * v.x + m[0].b * v.y, m[0].c * v.x + m[0].d * v.y);
        xs[row] = v;
    }
    if (f2 >> 27) m[0] = ld_blk_stream<false>(c.lu, f2 & 0xFFFFFF);
}

__global__ void __launch_bounds__(RC_T, 1) k_solve_rows_deep(PlanDev p, WorkDev w, int mode) {
    using Sy = helm::Symbolic;
    TL_SCOPE(w, 2);
    extern __shared__ __align__(16) double2 xs_rc[];
    int grp;
    if (mode == 2) {
        if ((int)blockIdx.x >= w.solve_count[w.parity]) return;
        grp = w.solve_list[(size_t)w.parity * w.ngroups + blockIdx.x];
    } else {
        grp = blockIdx.x;
        const int gfl = w.g_flags[grp];
        if (mode ? !(gfl & GF_FACTORING) : (gfl & (GF_DONE | GF_FACTORING)) != 0) return;
    }
    if (w.s_status[grp] != ST_ACTIVE) return;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int N = p.N, nsegs = p.rc_nsegs;
    RcDeep c;
    c.lu = lu_of(p, w, grp);
    c.xs = xs_rc;
    const double2 *rhs = w.rhs + (size_t)grp * N;
    double2 *xb = w.xb + (size_t)grp * N;
    const int first = __ldg(p.rc_lists + wp);
    const int4 *desc = p.rc_desc + (size_t)first * 32 + lane;
    const int *sgp = p.rc_seg + first;
    int4 r0 = __ldg(desc), r1 = __ldg(desc + 32), r2 = __ldg(desc + 64), r3 = __ldg(desc + 96);
    int s0 = __ldg(sgp), s1 = __ldg(sgp + 1), s2 = __ldg(sgp + 2), s3 = __ldg(sgp + 3);
    Blk256 ma[4], mb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) ma[t] = mb[t] = Blk256{0.0, 0.0, 0.0, 0.0};
    rows_fetch<0, false>(c.lu, r0.z, ma);
    rows_fetch<0, false>(c.lu, r1.z, mb);
    for (int i = tid; i < N; i += RC_T) c.xs[i] = rhs[i];             // rows without L entries: y = rhs
    __syncthreads();
    int seg = 0;
    auto barriers = [&](int upto) {
        while (seg < upto) { __syncthreads(); ++seg; }
    };
    // four waves per trip: a record is reloaded (four waves ahead) right after its wave, a block set
    // (two waves ahead) inside its wave
    while (true) {
        if (r0.w & Sy::RL_STOP) break;
        barriers(s0);
        rcd_wave(c, r0, r2.z, ma);
        r0 = __ldg(desc + 128); s0 = __ldg(sgp + 4);
        if (r1.w & Sy::RL_STOP) break;
        barriers(s1);
        rcd_wave(c, r1, r3.z, mb);
        r1 = __ldg(desc + 160); s1 = __ldg(sgp + 5);
        if (r2.w & Sy::RL_STOP) break;
        barriers(s2);
        rcd_wave(c, r2, r0.z, ma);
        r2 = __ldg(desc + 192); s2 = __ldg(sgp + 6);
        if (r3.w & Sy::RL_STOP) break;
        barriers(s3);
        rcd_wave(c, r3, r1.z, mb);
        r3 = __ldg(desc + 224); s3 = __ldg(sgp + 7);
        desc += 128; sgp += 4;
    }
    barriers(nsegs);                                                  // every warp passes every barrier
    for (int i = tid; i < N; i += RC_T) xb[i] = c.xs[i];
}

// ---- K2 (G = 1), shared-memory wave solve --------------------------------------------------------
// One CTA per scenario; same wave schedule as k_solve_flat, but the solution vector lives in
// shared memory, so a dependent level costs a block barrier and a shared-memory gather instead of an
// L2/HBM round trip.  The factor blocks and the {column, meta} table are consumed strictly in
// storage order, so they are streamed in chunks of <= 128 slots into a shared-memory ring with bulk
// asynchronous copies (cp.async.bulk, completion on an mbarrier) issued CS_NST chunks ahead by one
// thread: the dependency chain never waits for HBM.  Waves of one level are spread over the warps.
#ifndef CS_T_N
#define CS_T_N 128
#endif
constexpr int CS_T = CS_T_N;
constexpr int CS_NWARP = CS_T / 32;
constexpr int CS_NST = 3;
constexpr int CS_CH = helm::Symbolic::SOLVE_CHUNK_SLOTS;
constexpr int CS_STAGE_BYTES = (CS_CH * 32 + (CS_CH + 2) * 8 + 127) / 128 * 128;

struct CtaSolveLayout { size_t off_x, off_wave, off_chunk, off_ring, off_carry, total; };
__host__ __device__ inline CtaSolveLayout cta_solve_layout(int N, int nwaves, int nchunks) {
    CtaSolveLayout L;
    size_t o = 128;                                          // mbarriers
    L.off_carry = o; o += 128;                               // carry of a row continued in the next wave
    L.off_x = o; o += ((size_t)N * 16 + 127) / 128 * 128;
    L.off_wave = o; o += ((size_t)(nwaves + 1) / 2 * 16 + 127) / 128 * 128;
    L.off_chunk = o; o += ((size_t)nchunks * 16 + 127) / 128 * 128;
    L.off_ring = o; o += (size_t)CS_NST * CS_STAGE_BYTES;
    L.total = o;
    return L;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void cta_issue_chunk(const PlanDev &p, const double2 *lu, unsigned char *ring,
                                                unsigned long long *bars, int c, int4 ch) {
    const int stage = c % CS_NST;
    unsigned char *sb = ring + (size_t)stage * CS_STAGE_BYTES;
    const int s0 = ch.z, ns = ch.w;
    const int t0 = s0 & ~1, tn = ((s0 + ns + 1) & ~1) - t0;
    mbar_expect_tx(bars + stage, (unsigned)(ns * 32 + tn * 8));
    bulk_g2s(sb, lu + (size_t)s0 * 2, (unsigned)(ns * 32), bars + stage);
    bulk_g2s(sb + CS_CH * 32, p.scm + (size_t)t0 * 2, (unsigned)(tn * 8), bars + stage);
}

struct CtaOps { double2 m0, m1; int col, meta, info; };
__device__ __forceinline__ void cta_load_ops(CtaOps &o, const int2 *swave, const double2 *slu, const int2 *stab,
                                             int slot0, int tab0, int k, int lane) {
    const int2 wi = swave[k];
    o.info = wi.y;
    o.m0 = o.m1 = make_double2(0.0, 0.0);
    o.col = 0; o.meta = helm::Symbolic::META_DIAG;
    if (lane < (wi.y & 63)) {
        const int idx = wi.x - slot0 + lane;
        o.m0 = slu[2 * idx];
        o.m1 = slu[2 * idx + 1];
        const int2 t = stab[wi.x + lane - tab0];
        o.col = t.x; o.meta = t.y;
    }
}

// waves [a, e): each a level of its own, processed in order by one warp
__device__ __forceinline__ void cta_thin_run(const int2 *swave, const double2 *slu, const int2 *stab, int slot0,
                                             int tab0, double2 *xs, double2 *carry_s, int a, int e, int lane) {
    using Sy = helm::Symbolic;
    const double2 zero = make_double2(0.0, 0.0);
    CtaOps nx;
    cta_load_ops(nx, swave, slu, stab, slot0, tab0, a, lane);
    for (int k = a; k < e; ++k) {
        const CtaOps o = nx;
        if (k + 1 < e) cta_load_ops(nx, swave, slu, stab, slot0, tab0, k + 1, lane);
        if (k > a) __syncwarp();                             // x of the previous level is in shared memory
        if (o.info & Sy::WAVE_ONEROW) {
            // the whole wave is one row piece (the dense chain near the root): plain butterfly, no
            // segment logic; lane 0 is the head (and holds the pivot block in the backward solve).
            // A row longer than a wave hands its partial sum (and pivot) on through carry_s, like the
            // general path, so the two paths can follow each other on the pieces of one row.
            const int hmeta = __shfl_sync(0xFFFFFFFFu, o.meta, 0);
            const int row = hmeta >> 9;
            const bool rs = (hmeta & Sy::META_ROWSTART) != 0, re = (hmeta & Sy::META_ROWEND) != 0;
            const bool isu = (o.info & Sy::WAVE_ISU) != 0;
            double2 xg = zero;
            if (!(o.meta & Sy::META_DIAG)) xg = xs[o.col];
            const double2 b = rs ? xs[row] : carry_s[0];     // broadcast loads
            double2 v = make_double2(o.m0.x * xg.x + o.m0.y * xg.y, o.m1.x * xg.x + o.m1.y * xg.y);
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                v.x += __shfl_xor_sync(0xFFFFFFFFu, v.x, s);
                v.y += __shfl_xor_sync(0xFFFFFFFFu, v.y, s);
            }
            const double2 acc = make_double2(b.x - v.x, b.y - v.y);
            if (re) {
                double2 out = acc;
                if (isu) {
                    const double2 d0 = rs ? o.m0 : carry_s[1], d1 = rs ? o.m1 : carry_s[2];   // lane 0: the row's inverted pivot
                    out = make_double2(d0.x * acc.x + d0.y * acc.y, d1.x * acc.x + d1.y * acc.y);
                }
                if (lane == 0) xs[row] = out;
            } else if (lane == 0) {
                carry_s[0] = acc;
                if (isu && rs) { carry_s[1] = o.m0; carry_s[2] = o.m1; }
            }
            continue;
        }
        const int meta = o.meta, info = o.info;
        double2 xg = zero, b = zero;
        if (!(meta & Sy::META_DIAG)) xg = xs[o.col];
        const bool head = (meta & Sy::META_HEAD) != 0, rs = (meta & Sy::META_ROWSTART) != 0;
        if (head) b = rs ? xs[meta >> 9] : carry_s[0];
        double2 v = make_double2(o.m0.x * xg.x + o.m0.y * xg.y, o.m1.x * xg.x + o.m1.y * xg.y);
        const int span = meta & 31, maxspan = (info >> 9) & 31;
        for (int s = 1; s <= maxspan; s <<= 1) {
            double tx = __shfl_down_sync(0xFFFFFFFFu, v.x, s);
            double ty = __shfl_down_sync(0xFFFFFFFFu, v.y, s);
            if (s <= span) { v.x += tx; v.y += ty; }
        }
        if (head) {
            const double2 acc = make_double2(b.x - v.x, b.y - v.y);
            if (meta & Sy::META_ROWEND) {
                double2 out = acc;
                if (info & Sy::WAVE_ISU) {
                    const double2 d0 = rs ? o.m0 : carry_s[1];
                    const double2 d1 = rs ? o.m1 : carry_s[2];
                    out = make_double2(d0.x * acc.x + d0.y * acc.y, d1.x * acc.x + d1.y * acc.y);
                }
                xs[meta >> 9] = out;
            } else {
                carry_s[0] = acc;
                if ((info & Sy::WAVE_ISU) && rs) { carry_s[1] = o.m0; carry_s[2] = o.m1; }
            }
        }
    }
}

__global__ void __launch_bounds__(CS_T) k_solve_cta(PlanDev p, WorkDev w, int side) {
    using Sy = helm::Symbolic;
    TL_SCOPE(w, 2);
    extern __shared__ __align__(128) unsigned char smem_c[];
    const int grp = blockIdx.x;                              // G = 1: one scenario per CTA
    {
        int gfl = w.g_flags[grp];
        if (side ? !(gfl & GF_FACTORING) : (gfl & (GF_DONE | GF_FACTORING)) != 0) return;
    }
    if (w.s_status[grp] != ST_ACTIVE) return;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int N = p.N, nchunks = p.nschunks;
    const CtaSolveLayout L = cta_solve_layout(N, p.nwaves, nchunks);
    unsigned long long *bars = (unsigned long long *)smem_c;           // [0..NST-1] ring stages, [NST] x + tables
    double2 *carry_s = (double2 *)(smem_c + L.off_carry);              // [0] acc, [1..2] inverted pivot
    double2 *xs = (double2 *)(smem_c + L.off_x);
    const int2 *swave = (const int2 *)(smem_c + L.off_wave);
    const int4 *schunk = (const int4 *)(smem_c + L.off_chunk);
    unsigned char *ring = smem_c + L.off_ring;
    const double2 *lu = lu_of(p, w, grp);
    const double2 *rhs = w.rhs + (size_t)grp * N;
    double2 *xb = w.xb + (size_t)grp * N;
    if (tid == 0) {
        for (int i = 0; i <= CS_NST; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned wbytes = (unsigned)((p.nwaves + 1) / 2 * 16), cbytes = (unsigned)(nchunks * 16);
        mbar_expect_tx(bars + CS_NST, (unsigned)(N * 16) + wbytes + cbytes);
        bulk_g2s(xs, rhs, (unsigned)(N * 16), bars + CS_NST);
        bulk_g2s((void *)swave, p.wave_tab, wbytes, bars + CS_NST);
        bulk_g2s((void *)schunk, p.solve_chunks, cbytes, bars + CS_NST);
        for (int c = 0; c < CS_NST && c < nchunks; ++c) cta_issue_chunk(p, lu, ring, bars, c, __ldg(p.solve_chunks + c));
    }
    TL_CLK_DECL;
    mbar_wait(bars + CS_NST, 0);
    TL_CLK_ADD(w, 1);
    const double2 zero = make_double2(0.0, 0.0);
    for (int c = 0; c < nchunks; ++c) {
        const int stage = c % CS_NST;
        const int4 ch = schunk[c];
        mbar_wait(bars + stage, (unsigned)((c / CS_NST) & 1));
        TL_CLK_ADD(w, 2);
        const unsigned char *sb = ring + (size_t)stage * CS_STAGE_BYTES;
        const double2 *slu = (const double2 *)sb;
        const int2 *stab = (const int2 *)(sb + CS_CH * 32);
        const int tab0 = ch.z & ~1;
        int wv = ch.x;
        const int wend = ch.x + ch.y;
        bool fresh = true;                                   // a block barrier has just been passed
        while (wv < wend) {
            if (swave[wv].y & Sy::WAVE_THIN) {
                // chain of one-wave levels (the long tail near the root of the elimination tree):
                // warp 0 runs it alone — warp barriers instead of block barriers, and the operands of
                // the next wave (which do not depend on x) are fetched from the ring while this one
                // is reduced; the other warps wait at the next block barrier
                int e = wv + 1;
                while (e < wend && (swave[e].y & Sy::WAVE_THIN)) ++e;
                TL_CLK_ADD(w, 3);
                if (!fresh) __syncthreads();
                TL_CLK_ADD(w, 4);
                fresh = false;
                if (wp == 0) cta_thin_run(swave, slu, stab, ch.z, tab0, xs, carry_s, wv, e, lane);
                TL_CLK_ADD(w, 6);
                wv = e;
                continue;
            }
            int r = wv + 1;
            while (r < wend && !(swave[r].y & Sy::WAVE_SYNC)) ++r;
            TL_CLK_ADD(w, 3);
            if ((swave[wv].y & Sy::WAVE_SYNC) && !fresh) __syncthreads();
            TL_CLK_ADD(w, 4);
            fresh = false;
            for (int k = wv + wp; k < r; k += CS_NWARP) {
                const int2 wi = swave[k];
                const int info = wi.y, n = info & 63;
                double2 m0 = zero, m1 = zero;
                int col = 0, meta = Sy::META_DIAG;
                if (lane < n) {
                    const int idx = wi.x - ch.z + lane;
                    m0 = slu[2 * idx];
                    m1 = slu[2 * idx + 1];
                    const int2 t = stab[wi.x + lane - tab0];
                    col = t.x; meta = t.y;
                }
                double2 xg = zero;
                if (!(meta & Sy::META_DIAG)) xg = xs[col];
                double2 v = make_double2(m0.x * xg.x + m0.y * xg.y, m1.x * xg.x + m1.y * xg.y);
                const int span = meta & 31, maxspan = (info >> 9) & 31;
                for (int o = 1; o <= maxspan; o <<= 1) {
                    double tx = __shfl_down_sync(0xFFFFFFFFu, v.x, o);
                    double ty = __shfl_down_sync(0xFFFFFFFFu, v.y, o);
                    if (o <= span) { v.x += tx; v.y += ty; }
                }
                if (meta & Sy::META_HEAD) {
                    const int row = meta >> 9;
                    const bool rs = (meta & Sy::META_ROWSTART) != 0;
                    const double2 b = rs ? xs[row] : carry_s[0];
                    const double2 acc = make_double2(b.x - v.x, b.y - v.y);
                    if (meta & Sy::META_ROWEND) {
                        double2 out = acc;
                        if (info & Sy::WAVE_ISU) {
                            const double2 d0 = rs ? m0 : carry_s[1];
                            const double2 d1 = rs ? m1 : carry_s[2];
                            out = make_double2(d0.x * acc.x + d0.y * acc.y, d1.x * acc.x + d1.y * acc.y);
                        }
                        xs[row] = out;
                    } else {                                 // continued in the next wave (which is SYNC)
                        carry_s[0] = acc;
                        if ((info & Sy::WAVE_ISU) && rs) { carry_s[1] = m0; carry_s[2] = m1; }
                    }
                }
            }
            TL_CLK_ADD(w, 7);
            wv = r;
        }
        TL_CLK_ADD(w, 3);
        __syncthreads();                                     // the stage is free; also a level boundary
        if (tid == 0 && c + CS_NST < nchunks) cta_issue_chunk(p, lu, ring, bars, c + CS_NST, schunk[c + CS_NST]);
        TL_CLK_ADD(w, 4);
    }
    for (int i = tid; i < N; i += CS_T) xb[i] = xs[i];
    TL_CLK_ADD(w, 5);
}

// ---- K2 (G = 1), one-CTA solve on a static segment schedule --------------------------------------
// One CTA per scenario, x in shared memory, same waves and per-slot tables as the two kernels above,
// but built for the LATENCY of the dependent chain of levels (the single-case path, BASELINE config 2,
// and small batches):
//   * the waves between two ordering points form a segment (host table, symbolic.cpp): the waves of
//     a WIDE segment are dealt to the warps, one block barrier per segment; a run of one-wave levels
//     (the chain near the root of the elimination tree) is one THIN segment that warps 0 and 1 walk
//     alternately, handing over through a named barrier of the two warps, so that the x-independent
//     part of a wave overlaps the dependent part of the wave before it — no scanning of the wave
//     table, no per-chunk barriers; every warp has its list of waves from the host;
//   * everything of a wave that does not depend on x — factor blocks, columns and row metadata, one
//     contiguous KB + 256 B per wave — is copied ONE_D waves ahead from global memory / L2 into a
//     per-warp shared-memory ring with cp.async (no register is tied to a load in flight), so the
//     chain itself only holds shared-memory loads, the products, the segmented shuffle reduction
//     and the store of the row result.
constexpr int ONE_NW = helm::Symbolic::ONE_WARPS;        // warps per CTA
constexpr int ONE_T = ONE_NW * 32;
constexpr int ONE_D = 4;                                 // waves in flight per warp
struct OneStage { double2 blk[64]; int2 cm[32]; };       // 1280 B: [lane][2] factor block, {column, meta}
struct OneOps { double2 m0, m1; int col, meta, info; };

struct OneLayout { size_t off_x, off_wave, off_seg, off_list, off_ring, total; };
__host__ __device__ inline OneLayout one_layout(int N, int nwaves, int nsegs, int nlist) {
    OneLayout L;
    size_t o = 64;                                           // carry of a split row
    L.off_x = o; o += ((size_t)N * 16 + 15) / 16 * 16;
    L.off_ring = o; o += (size_t)ONE_NW * ONE_D * sizeof(OneStage);
    L.off_wave = o; o += (size_t)(nwaves + 1) * 8;
    L.off_seg = o; o += (size_t)(nsegs + 1) * 8;
    L.off_list = o; o += (size_t)(nlist + 1) * 4;
    L.total = (o + 15) / 16 * 16;
    return L;
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

// start the copies of wave k (or nothing) into a stage; one commit group either way
__device__ __forceinline__ void one_issue(OneStage *st, const int2 *swave, const double2 *lu, const int2 *scm, int k, int lane) {
    if (k >= 0) {
        const int2 wi = swave[k];
        const int e = wi.x + min(lane, (wi.y & 63) - 1);     // idle lanes repeat the last block (and contribute 0)
        cp_async16(&st->blk[2 * lane], lu + 2 * (size_t)e);
        cp_async16(&st->blk[2 * lane + 1], lu + 2 * (size_t)e + 1);
        cp_async8(&st->cm[lane], scm + e);
    }
    cp_async_commit();
}

// one wave: gather, product, segmented row sums, row results into xs (or the carry of a split row)
__device__ __forceinline__ void one_wave(const OneOps &o, double2 *xs, double2 *carry_s, int lane) {
    using Sy = helm::Symbolic;
    const double2 zero = make_double2(0.0, 0.0);
    if (o.info & Sy::WAVE_ONEROW) {
        // the whole wave is one row piece (the chain near the root): plain butterfly, no segment
        // logic; lane 0 is the head and holds the pivot block in the backward solve
        const int hmeta = __shfl_sync(0xFFFFFFFFu, o.meta, 0);
        const int row = hmeta >> 9;
        const bool rs = (hmeta & Sy::META_ROWSTART) != 0, re = (hmeta & Sy::META_ROWEND) != 0;
        const bool isu = (o.info & Sy::WAVE_ISU) != 0;
        double2 xg = zero;
        if (!(o.meta & Sy::META_DIAG)) xg = xs[o.col];
        const double2 b = rs ? xs[row] : carry_s[0];         // broadcast loads
        double2 v = make_double2(o.m0.x * xg.x + o.m0.y * xg.y, o.m1.x * xg.x + o.m1.y * xg.y);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            v.x += __shfl_xor_sync(0xFFFFFFFFu, v.x, s);
            v.y += __shfl_xor_sync(0xFFFFFFFFu, v.y, s);
        }
        if (lane == 0) {
            const double2 acc = make_double2(b.x - v.x, b.y - v.y);
            if (re) {
                double2 out = acc;
                if (isu) {
                    const double2 d0 = rs ? o.m0 : carry_s[1], d1 = rs ? o.m1 : carry_s[2];
                    out = make_double2(d0.x * acc.x + d0.y * acc.y, d1.x * acc.x + d1.y * acc.y);
                }
                xs[row] = out;
            } else {
                carry_s[0] = acc;
                if (isu && rs) { carry_s[1] = o.m0; carry_s[2] = o.m1; }
            }
        }
        return;
    }
    const int meta = o.meta, info = o.info;
    double2 xg = zero, b = zero;
    if (!(meta & Sy::META_DIAG)) xg = xs[o.col];
    const bool head = (meta & Sy::META_HEAD) != 0, rs = (meta & Sy::META_ROWSTART) != 0;
    if (head) b = rs ? xs[meta >> 9] : carry_s[0];
    double2 v = make_double2(o.m0.x * xg.x + o.m0.y * xg.y, o.m1.x * xg.x + o.m1.y * xg.y);
    const int span = meta & 31, maxspan = (info >> 9) & 31;
    for (int s = 1; s <= maxspan; s <<= 1) {                 // warp-uniform trip count
        double tx = __shfl_down_sync(0xFFFFFFFFu, v.x, s);
        double ty = __shfl_down_sync(0xFFFFFFFFu, v.y, s);
        add_if_le(v, tx, ty, s, span);
    }
    if (head) {
        const double2 acc = make_double2(b.x - v.x, b.y - v.y);
        if (meta & Sy::META_ROWEND) {
            double2 out = acc;
            if (info & Sy::WAVE_ISU) {
                const double2 d0 = rs ? o.m0 : carry_s[1];
                const double2 d1 = rs ? o.m1 : carry_s[2];
                out = make_double2(d0.x * acc.x + d0.y * acc.y, d1.x * acc.x + d1.y * acc.y);
            }
            xs[meta >> 9] = out;
        } else {
            carry_s[0] = acc;
            if ((info & Sy::WAVE_ISU) && rs) { carry_s[1] = o.m0; carry_s[2] = o.m1; }
        }
    }
}

__global__ void __launch_bounds__(ONE_T) k_solve_one(PlanDev p, WorkDev w, int side) {
    using Sy = helm::Symbolic;
    TL_SCOPE(w, 2);
    extern __shared__ __align__(16) unsigned char smem_o[];
    const int abl = side >> 8;                               // timing ablations (helm_debug_solve_time); 0 in production
    side &= 255;
    int grp;
    if (side == 2) {                                         // one CTA per entry of the step's solve list
        if ((int)blockIdx.x >= w.solve_count[w.parity]) return;
        grp = w.solve_list[(size_t)w.parity * w.ngroups + blockIdx.x];
    } else {
        grp = blockIdx.x;
        const int gfl = w.g_flags[grp];
        if (side ? !(gfl & GF_FACTORING) : (gfl & (GF_DONE | GF_FACTORING)) != 0) return;
    }
    if (w.s_status[grp] != ST_ACTIVE) return;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int N = p.N, nwaves = p.nwaves, nsegs = p.nsegs;
    const OneLayout L = one_layout(N, nwaves, nsegs, p.nonelist);
    double2 *carry_s = (double2 *)smem_o;                                           // [0] acc, [1..2] inverted pivot
    double2 *xs = (double2 *)(smem_o + L.off_x);
    OneStage *ring = (OneStage *)(smem_o + L.off_ring) + (size_t)wp * ONE_D;
    int2 *swave = (int2 *)(smem_o + L.off_wave);
    int2 *sseg = (int2 *)(smem_o + L.off_seg);
    int *slist = (int *)(smem_o + L.off_list);
    const double2 *lu = lu_of(p, w, grp);
    const double2 *rhs = w.rhs + (size_t)grp * N;
    double2 *xb = w.xb + (size_t)grp * N;
    const int2 *scm = (const int2 *)p.scm;
    for (int i = tid; i < nwaves; i += ONE_T) swave[i] = __ldg(p.wave_tab + i);
    for (int i = tid; i < nsegs; i += ONE_T) sseg[i] = __ldg(p.solve_segs + i);
    for (int i = tid; i < p.nonelist; i += ONE_T) slist[i] = __ldg(p.one_lists + i);
    for (int i = tid; i < N; i += ONE_T) xs[i] = rhs[i];
    __syncthreads();
    TL_CLK_DECL;
    TL_CLK_ADD(w, 1);
    // the warp's waves, in schedule order: slist[l0 .. l1)
    const int l0 = ONE_NW + 1 + slist[wp], l1 = ONE_NW + 1 + slist[wp + 1];
    int li = l0;                                              // next wave to issue
    int cons = 0;                                             // waves consumed so far
    if (!(abl & 64)) {
#pragma unroll
        for (int d = 0; d < ONE_D; ++d) {
            one_issue(ring + d, swave, lu, scm, li < l1 ? slist[li] : -1, lane);
            ++li;
        }
    }
    auto take = [&](int k) -> OneOps {                        // operands of wave k (the oldest group in flight)
        if (abl & 64) {                                      // ablation: no ring, direct loads
            OneOps o;
            const int2 wi = swave[k];
            o.info = wi.y;
            const int e = wi.x + min(lane, (wi.y & 63) - 1);
            o.m0 = lu[2 * (size_t)e];
            o.m1 = lu[2 * (size_t)e + 1];
            const int2 cm = __ldg(scm + e);
            o.col = cm.x;
            o.meta = (lane < (wi.y & 63)) ? cm.y : Sy::META_DIAG;
            return o;
        }
        cp_async_wait<ONE_D - 1>();
        __syncwarp();
        OneStage *st = ring + (cons % ONE_D);
        OneOps o;
        const int2 wi = swave[k];
        o.info = wi.y;
        o.m0 = st->blk[2 * lane];
        o.m1 = st->blk[2 * lane + 1];
        const int2 cm = st->cm[lane];
        o.col = cm.x;
        o.meta = (lane < (wi.y & 63)) ? cm.y : Sy::META_DIAG;
        one_issue(st, swave, lu, scm, (li < l1 && !(abl & 2)) ? slist[li] : -1, lane);
        ++li; ++cons;
        return o;
    };
    // `pre` = operands of the warp's next wave, fetched right after the previous one was processed: the
    // x-independent part of a wave (ring wait, shared-memory loads, issue of the next copies) runs
    // while the warp would otherwise wait at the barrier of the segment
    OneOps pre;
    if (l0 < l1) pre = take(slist[l0]);
    auto run = [&]() {
        OneOps o = pre;
        if (abl & 1) o.info &= ~(31 << 9);
        if (abl & 4) o.meta |= Sy::META_DIAG;
        return o;
    };
    for (int s = 0; s < nsegs; ++s) {
        const int2 sg = sseg[s];
        const int first = sg.x, n = sg.y & 0x3FFFFFFF;
        const bool thin = (sg.y >> 30) != 0;
        TL_CLK_ADD(w, 6);
        if (s > 0 && !(abl & 8)) __syncthreads();
        TL_CLK_ADD(w, 2);
        if ((abl & 16) && thin) continue;
        if ((abl & 32) && !thin) continue;
        if (thin) {
            // warps 0 and 1 alternate; wave j waits for wave j - 1 on a named barrier of the two warps
            // (id 1: an even wave is done, id 2: an odd wave is done)
            if (wp < 2) {
                for (int j = wp; j < n; j += 2) {
                    const OneOps o = run();
                    if (j > 0) asm volatile("bar.sync %0, 64;" ::"r"(wp ? 1 : 2) : "memory");
                    one_wave(o, xs, carry_s, lane);
                    if (j + 1 < n) asm volatile("bar.arrive %0, 64;" ::"r"(wp ? 2 : 1) : "memory");
                    if (l0 + cons < l1) pre = take(slist[l0 + cons]);
                }
            }
            TL_CLK_ADD(w, 3);
        } else {
            for (int k = first + wp; k < first + n; k += ONE_NW) {
                const OneOps o = run();
                one_wave(o, xs, carry_s, lane);
                if (l0 + cons < l1) pre = take(slist[l0 + cons]);
            }
            TL_CLK_ADD(w, 4);
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    TL_CLK_ADD(w, 2);
    for (int i = tid; i < N; i += ONE_T) xb[i] = xs[i];
    TL_CLK_ADD(w, 5);
}

// ---- K_init: order-0 germ and RHS of order 1 (helm.py:110-139, 551-556, n == 1 branches) ---
template <int G>
__global__ void k_init(PlanDev p, WorkDev w) {
    TL_SCOPE(w, 7);
  const int nfact = w.fact_count[w.parity];
  for (int fi = blockIdx.y; fi < nfact; fi += gridDim.y) {
    int grp = w.fact_list[(size_t)w.parity * w.ngroups + fi];
    int flat = blockIdx.x * blockDim.x + threadIdx.x;
    int i = flat / G, s = flat % G;
    if (i >= p.N) return;
    int sg = grp * G + s;
    if (w.s_status[sg] != ST_ACTIVE) continue;
    const int N = p.N;
    double sc = w.scale[sg];
    size_t bi = ix<G>(grp, N, i, s);
    unsigned char bt = w.btype[bi];
    double2 one = make_double2(1.0, 0.0);
    w.Vh[hx<G>(w, N, grp, 0, i, s)] = one;
    w.Hh[hx<G>(w, N, grp, 0, i, s)] = one;     // W[0] = 1
    w.Eh[hx<G>(w, N, grp, 0, i, s)] = one;     // partial sum S_0
    w.VVant[bi] = 0.0;
    const OutageRef og = load_outage(p, w, sg);
    double2 r;
    if (bt == HELM_BUS_SLACK) {
        r = make_double2(p.vsp[i] - 1.0, 0.0);                                   // helm.py:393-394
        w.Pcur[bi] = make_double2(p.vsp[i], 0.0);
    } else if (bt == HELM_BUS_PQ) {
        double Pi = p.pg[i] * sc - p.pd[i] * sc;
        double Qi = w.Qg[bi] - p.qd[i] * sc;
        double2 pp = make_double2(0.0, 0.0);
        for (int a = p.ph_ptr[i]; a < p.ph_ptr[i + 1]; ++a) pp = cadd(pp, ph_eff(p, og, a));
        double2 ysh = ysh_eff(p, og, i);
        r = make_double2(Pi - ysh.x - pp.x, -Qi - ysh.y - pp.y);
    } else {
        double cc = (p.pg[i] * sc - p.pd[i] * sc) - ysh_eff(p, og, i).x;         // helm.py:347-352
        for (int a = p.ph_ptr[i]; a < p.ph_ptr[i + 1]; ++a) cc -= ph_eff(p, og, a).x;
        double vs = p.vsp[i];
        r = make_double2(cc, (vs * vs - 1.0) * 0.5);                             // helm.py:354-355, 364
    }
    w.rhs[bi] = r;
  }
}

// ---- K1 / K3: V and W update and the RHS of the next order (k_rhs_conv); epsilon table and Pade
// check (k_eps).  One thread per (bus, scenario).  Replace compute_complex_voltages (helm.py:402-420),
// calculate_inverse_voltages_w_array (423-433), the per-bus evaluators (293-398) and the
// Pade/convergence test (608-641, analytic_continuation.py).  The Pade value is obtained with
// Wynn's epsilon algorithm kept incrementally (the reference states the algorithm itself in
// analytic_continuation.py:9-26).
//
// The O(j) work of one order is a loop over the history index t = 1..j:
//   PQ:  W_i[j]  = -sum_t W_i[t-1] V_i[j+1-t]                  (blocked by four orders, conv_blocked_pq)
//   PV:  PPP     =  sum_t conj(V_i[j+1-t]) CC_i[t]          (CC_i[j] = sum_k Ytrans_ik V_k[j])
//        VV      =  sum_t V_i[t] conj(V_i[j+1-t])
//   eps: eps_t^(j-t) = eps_{t-2}^(j-t+1) + 1 / (eps_{t-1}^(j-t+1) - eps_{t-1}^(j-t))
// The loops are blocked: the loads of a block (one coalesced 16 B per lane each, histories are
// tiled by 128 buses) are issued together before the dependent arithmetic, which is what gives
// the kernels their memory-level parallelism.

struct EpsState { double2 prev1, prev2, cur; };

__device__ __forceinline__ void eps_step(EpsState &st, int t, int j, double2 v, double2 eold_t, double2 *Eb,
                                         size_t plane) {
    double2 d = (t == 1) ? v : csub(st.cur, st.prev1);
    double2 e = cadd(st.prev2, crecip_guard(d));
    st.prev2 = st.prev1;
    if (t < j) st.prev1 = eold_t;
    Eb[(size_t)t * plane] = e;
    st.cur = e;
}

// Two anti-diagonals of the epsilon table per visit (even orders only).  `o` is the stored
// anti-diagonal j-2, `a` the anti-diagonal j-1 (never stored), `b` the anti-diagonal j (stored over o):
//   a_t = o_{t-2} + 1/(a_{t-1} - o_{t-1}),  b_t = a_{t-2} + 1/(b_{t-1} - a_{t-1}).
// Same operations in the same order as appending one anti-diagonal per order (eps_step), so the
// values are bit-identical; the table is read and written every second order instead of every order.
struct Eps2State { double2 o1, o2, a1, a2, b1; };

__device__ __forceinline__ void eps2_step(Eps2State &st, int t, int j, double2 vjm1, double2 v, double2 eold_t,
                                          double2 *Eb, size_t plane) {
    double2 a = st.a1;
    if (t < j) a = cadd(st.o2, crecip_guard((t == 1) ? vjm1 : csub(st.a1, st.o1)));
    double2 b = cadd(st.a2, crecip_guard((t == 1) ? v : csub(st.b1, st.a1)));
    Eb[(size_t)t * plane] = b;
    st.o2 = st.o1; st.o1 = eold_t;
    st.a2 = st.a1; st.a1 = a;
    st.b1 = b;
}

// Blocked ("relaxed") form of the same convolution C[j] = sum_{a<j} W[a] V[j-a], block size 4.
// At a block start j = T (T = 4, 8, ...) the history is walked once, and next to
// C[T] the loop also accumulates, through a three-deep window of the V just loaded, the parts of
// C[T+1..T+3] whose operands are already known: P[T+i] = sum_{a=i..T} W[a] V[T+i-a].  The three
// following orders read their partial and add the 2i-1 products with the operands that arrived
// since (V[T+1..j], W[T+1..j-1]) — one memory round trip instead of a walk of the history.
// Per block of four orders a bus reads 32 T + 384 B instead of 128 T + 320 B.
// operands of an order between two block starts (j >= 4, j % 4 = i > 0): the partial sum and the 4 i - 4
// history values that arrived since the block start.  k_rhs_conv requests them as soon as it knows j,
// together with the rest of its prologue, so that such an order costs ONE memory round trip.
struct ConvLight { double2 part, w1, vj1, wt1, va, w2, vj2, wt2, vb; };
__device__ __forceinline__ bool conv_is_light(int j) { return j >= 4 && (j & 3) != 0; }
__device__ __forceinline__ void conv_light_load(ConvLight &L, int j, const double2 *Vb, const double2 *Hb, const double2 *Pb, size_t plane) {
    const int i = j & 3, T = j & ~3;
    L.part = Pb[(size_t)(i - 1) * plane];
    if (i >= 2) {
        L.w1 = Hb[plane];                               // W[1]
        L.vj1 = Vb[(size_t)(j - 1) * plane];            // V[j-1]
        L.wt1 = Hb[(size_t)(T + 1) * plane];            // W[T+1]
        L.va = Vb[(size_t)(i - 1) * plane];             // V[i-1]
    }
    if (i == 3) {
        L.w2 = Hb[2 * plane];                           // W[2]
        L.vj2 = Vb[(size_t)(j - 2) * plane];            // V[j-2]
        L.wt2 = Hb[(size_t)(T + 2) * plane];            // W[T+2]
        L.vb = Vb[plane];                               // V[1]
    }
}
__device__ __forceinline__ double2 conv_light_sum(const ConvLight &L, int j, double2 v) {
    const int i = j & 3;
    double2 c = cadd(L.part, v);                        // a = 0: W[0] = 1 (helm.py:551)
    if (i >= 2) c = cadd(cadd(c, cmul(L.w1, L.vj1)), cmul(L.wt1, L.va));
    if (i == 3) c = cadd(cadd(c, cmul(L.w2, L.vj2)), cmul(L.wt2, L.vb));
    return c;
}

template <int KU>
__device__ __forceinline__ double2 conv_blocked_pq(int j, double2 v, const double2 *Vb, const double2 *Hb,
                                                   double2 *Pb, size_t plane) {
    const double2 zero = make_double2(0.0, 0.0);
    if (conv_is_light(j)) {                                 // (k_rhs_conv takes this path itself, with early loads)
        ConvLight L;
        L.w1 = L.vj1 = L.wt1 = L.va = L.w2 = L.vj2 = L.wt2 = L.vb = zero;
        conv_light_load(L, j, Vb, Hb, Pb, plane);
        return conv_light_sum(L, j, v);
    }
    double2 acc0 = zero, acc1 = zero, acc2 = zero, acc3 = zero;
    double2 win1 = zero, win2 = zero, win3 = zero;         // V[T-a+1], V[T-a+2], V[T-a+3] (0 while unknown)
    for (int t0 = 1; t0 <= j; t0 += KU) {
        double2 vr[KU], hh[KU];
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            int t = t0 + u;
            if (t <= j) {
                vr[u] = (t == 1) ? v : Vb[(size_t)(j + 1 - t) * plane];
                hh[u] = Hb[(size_t)(t - 1) * plane];                              // W[t-1]
            }
        }
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            int t = t0 + u;
            if (t <= j) {
                acc0 = cadd(acc0, cmul(hh[u], vr[u]));                            // helm.py:430-432
                acc1 = cadd(acc1, cmul(hh[u], win1));
                acc2 = cadd(acc2, cmul(hh[u], win2));
                acc3 = cadd(acc3, cmul(hh[u], win3));
                win3 = win2; win2 = win1; win1 = vr[u];
            }
        }
    }
    if (j >= 4) {                                          // block start: W[T] = -C[T] pairs with V[1..3]
        const double2 wT = make_double2(-acc0.x, -acc0.y);
        Pb[0] = cadd(acc1, cmul(wT, win1));
        Pb[plane] = cadd(acc2, cmul(wT, win2));
        Pb[2 * plane] = cadd(acc3, cmul(wT, win3));
    }
    return acc0;
}

// O(j) loop of a PV bus (model 2) at order j: P-row convolution and |V|^2 convolution.
template <int KU>
__device__ __forceinline__ void conv_loop_pv(int j, double2 v, double2 PP, const double2 *Vb, const double2 *Hb,
                                             size_t plane, double2 &ppp, double &vvre) {
    for (int t0 = 1; t0 <= j; t0 += KU) {
        double2 vr[KU], hh[KU], vf[KU];
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            int t = t0 + u;
            if (t <= j) {
                vr[u] = (t == 1) ? v : Vb[(size_t)(j + 1 - t) * plane];
                hh[u] = (t < j) ? Hb[(size_t)t * plane] : PP;                     // CC[t]
                vf[u] = (t < j) ? Vb[(size_t)t * plane] : v;                      // V[t]
            }
        }
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            int t = t0 + u;
            if (t <= j) {
                ppp = cadd(ppp, cmulc(hh[u], vr[u]));                             // helm.py:306-314
                vvre += vf[u].x * vr[u].x + vf[u].y * vr[u].y;                    // helm.py:356-361
            }
        }
    }
}

// 128-thread CTAs (= one history tile), 80 registers: six CTAs per SM.  The epsilon table and the
// Pade check run in their own kernel (k_eps); this one only carries the convolution state.
#ifndef CONV_EARLY
#define CONV_EARLY 1
#endif
#ifndef CONV_MINB
#define CONV_MINB 6
#define CONV_KU_PQ 4
#define CONV_KU_PV 4
#endif
template <int G>
__global__ void __launch_bounds__(128, CONV_MINB) k_rhs_conv(PlanDev p, WorkDev w) {
    TL_SCOPE(w, 3);
    const int grp = blockIdx.y;
    const int flat = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = flat / G, s = flat % G;
    if (i >= p.N) return;
    const int N = p.N;
    const int sg = grp * G + s;
    const size_t plane = hplane<G>();
    const size_t bi = ix<G>(grp, N, i, s);
    const size_t h0 = hx<G>(w, N, grp, 0, i, s);
    double2 *Vb = w.Vh + h0, *Hb = w.Hh + h0;
    const double2 *xb = w.xb + (size_t)grp * N * G;
    // Every operand of the prologue is loaded before the first branch: the addresses are valid for
    // any (bus, scenario), and issuing them together costs one memory round trip instead of six
    // dependent ones (ncu r1_c: the early-exit chain was the top stall at small orders).
    const int gf = w.g_flags[grp];
    const int status = w.s_status[sg];
    const int j = w.g_n[grp] + 1;  // order just solved
    const unsigned char bt = w.btype[bi];
#if CONV_EARLY
    // an order between two block starts of a PQ bus: its handful of operands is requested here, beside the
    // rest of the prologue (the flags, j and the bus type come from small L2-resident arrays)
    ConvLight lt;
    lt.part = lt.w1 = lt.vj1 = lt.wt1 = lt.va = lt.w2 = lt.vj2 = lt.wt2 = lt.vb = make_double2(0.0, 0.0);
    const bool light = bt == HELM_BUS_PQ && conv_is_light(j) && status == ST_ACTIVE && !(gf & (GF_DONE | GF_PASSEND | GF_FACTORING));
    if (light) conv_light_load(lt, j, Vb, Hb, w.Pb + px<G>(N, grp, 0, i, s), plane);
#endif
    const double2 v = xb[(size_t)i * G + s];
    const double sc = w.scale[sg];
    const double qg = w.Qg[bi];
    const double pgi = p.pg[i], pdi = p.pd[i], qdi = p.qd[i];
    double2 ysh = p.yshunt[i];
    const int ob = w.o_b[sg];
    const int ph0 = p.ph_ptr[i], ph1 = p.ph_ptr[i + 1];
    const int a0 = p.adj_ptr[i], a1 = p.adj_ptr[i + 1];
    if (gf & (GF_DONE | GF_PASSEND | GF_FACTORING)) return;
    if (status != ST_ACTIVE) return;
    // N-1 scenario: only the two end buses of the removed branch see different admittances
    OutageRef og{-1, -1, -1};
    if (ob >= 0) {
        OutageRef t_ = load_outage(p, w, sg);
        if (i == t_.f || i == t_.t) { og = t_; ysh = ysh_eff(p, og, i); }
    }
    Vb[(size_t)j * plane] = v;                                                    // helm.py:412-414
    if (bt == HELM_BUS_SLACK) {
        w.rhs[bi] = make_double2(0.0, 0.0);                                       // helm.py:395-397
        return;
    }
    const double2 zero = make_double2(0.0, 0.0);
    double2 out;
    if (bt == HELM_BUS_PQ) {
#if CONV_EARLY
        const double2 aux = light ? conv_light_sum(lt, j, v) : conv_blocked_pq<CONV_KU_PQ>(j, v, Vb, Hb, w.Pb + px<G>(N, grp, 0, i, s), plane);
#else
        const double2 aux = conv_blocked_pq<CONV_KU_PQ>(j, v, Vb, Hb, w.Pb + px<G>(N, grp, 0, i, s), plane);
#endif
        double2 wj = make_double2(-aux.x, -aux.y);
        Hb[(size_t)j * plane] = wj;                                               // W[j]
        // RHS of order j+1 (helm.py:367-385)
        double Pi = pgi * sc - pdi * sc;
        double Qi = qg - qdi * sc;
        double2 pp = zero;
        for (int a = ph0; a < ph1; ++a)
            pp = cadd(pp, cmul(ph_eff(p, og, a), xb[(size_t)p.ph_idx[a] * G + s]));
        double2 tt = cmul(make_double2(Pi, -Qi), make_double2(wj.x, -wj.y));
        double2 ysv = cmul(ysh, v);
        out = make_double2(tt.x - ysv.x - pp.x, tt.y - ysv.y - pp.y);
    } else {
        // PV bus, model 2 (helm.py:293-364); n = j+1 >= 2
        double2 PP = zero;
        for (int a = a0; a < a1; ++a)
            PP = cadd(PP, cmul(ytr_eff(p, og, a), xb[(size_t)p.adj_idx[a] * G + s]));   // helm.py:310-312
        Hb[(size_t)j * plane] = PP;                                               // barras_CC[i][n-1]
        double2 ppp = zero;
        double vvre = 0.0;
        conv_loop_pv<CONV_KU_PV>(j, v, PP, Vb, Hb, plane, ppp, vvre);
        double cc_ = 0.0 - ppp.x;
        if (ysh.x != 0.0) {                                                      // conduc_buses
            if (j >= 2) cc_ -= ysh.x * (w.VVant[bi] + 2.0 * v.x);                // helm.py:317-318
            else cc_ -= ysh.x * 2.0 * v.x;                                       // helm.py:336-337
        }
        if (ph1 > ph0) {                                                         // helm.py:320-327
            double2 acc = zero;
            for (int x = 0; x <= j; ++x) {
                double2 pp = zero;
                for (int a = ph0; a < ph1; ++a) {
                    int k = p.ph_idx[a];
                    double2 vk = (x == 0) ? xb[(size_t)k * G + s] : w.Vh[hx<G>(w, N, grp, j - x, k, s)];
                    pp = cadd(pp, cmul(ph_eff(p, og, a), vk));
                }
                double2 vx = (x == j) ? v : Vb[(size_t)x * plane];
                acc = cadd(acc, cmulc(pp, vx));
            }
            cc_ -= acc.x;
        }
        w.VVant[bi] = vvre;
        out = make_double2(cc_, -vvre * 0.5);
    }
    w.rhs[bi] = out;
}

// ---- K3: epsilon table + Pade check (even orders only: the orders at which the Pade value is
// needed, helm.py:611) -----------------------------------------------------------------------
// Thread per (bus, scenario).  Reads the stored anti-diagonal j-2, appends the anti-diagonals j-1 and
// j (eps2_step) and compares the new Pade value with the previous one (helm.py:608-641).  Independent
// of k_rhs_conv: both only read x and the V history.
#ifndef EPS_MINB
#define EPS_MINB 8
#define EPS_KE 6
#endif
#ifndef EPS_T
#define EPS_T 128           // threads per CTA of k_eps
#endif
constexpr int KE = EPS_KE;
template <int G>
__global__ void __launch_bounds__(EPS_T, EPS_MINB * 128 / EPS_T) k_eps(PlanDev p, WorkDev w) {
    TL_SCOPE(w, 3);
    const int grp = blockIdx.y;
    const int gf = w.g_flags[grp];
    const int j = w.g_n[grp] + 1;                          // order just solved
    if ((j & 1) || (gf & (GF_DONE | GF_PASSEND | GF_FACTORING))) return;   // uniform per CTA (G scenarios share j)
    const int flat = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = flat / G, s = flat % G;
    if (i >= p.N) return;
    const int N = p.N;
    const int sg = grp * G + s;
    const size_t plane = hplane<G>();
    const size_t bi = ix<G>(grp, N, i, s);
    const size_t h0 = hx<G>(w, N, grp, 0, i, s);
    const double2 *Vb = w.Vh + h0;
    double2 *Eb = w.Eh + h0;
    const int status = w.s_status[sg];
    const unsigned char bt = w.btype[bi];
    const double2 v = w.xb[(size_t)grp * N * G + (size_t)i * G + s];
    const double2 e_old0 = Eb[0];
    const double2 vjm1 = Vb[(size_t)(j - 1) * plane];
    const double2 pold = w.Pcur[bi];
    if (status != ST_ACTIVE || bt == HELM_BUS_SLACK) return;
    const double2 zero = make_double2(0.0, 0.0);
    Eps2State es;
    es.o1 = e_old0; es.o2 = zero;
    es.a1 = cadd(e_old0, vjm1); es.a2 = zero;
    es.b1 = cadd(es.a1, v);
    Eb[0] = es.b1;
    for (int t0 = 1; t0 <= j; t0 += KE) {
        double2 eo[KE];
#pragma unroll
        for (int u = 0; u < KE; ++u) {
            int t = t0 + u;
            eo[u] = (t < j - 1) ? Eb[(size_t)t * plane] : zero;
        }
#pragma unroll
        for (int u = 0; u < KE; ++u) {
            int t = t0 + u;
            if (t <= j) eps2_step(es, t, j, vjm1, v, eo[u], Eb, plane);
        }
    }
    const double2 cur = es.b1;                             // [j/2 / j/2] Pade at s = 1
    if (j >= 4) {                                          // helm.py:611: m > 3
        double mag1 = hypot(pold.x, pold.y), rad1 = atan2(pold.y, pold.x);
        double mag2 = hypot(cur.x, cur.y), rad2 = atan2(cur.y, cur.x);
        bool ok = (fabs(mag1 - mag2) <= w.mismatch) && (fabs(rad1 - rad2) <= w.mismatch);  // helm.py:618
        if (!ok) w.s_fail[sg] = 1;
    }
    w.Pcur[bi] = cur;
}

// ---- K_adv: per-group bookkeeping after an order (helm.py:608-657 control flow) ----------
__global__ void k_advance(WorkDev w) {
    TL_SCOPE(w, 4);
    int grp = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp >= w.ngroups) return;
    int gf = w.g_flags[grp];
    if (gf & (GF_DONE | GF_PASSEND | GF_FACTORING)) return;
    int j = w.g_n[grp] + 1, m = j + 1;
    bool checked = ((j & 1) == 0) && j >= 4;
    int nact = 0;
    for (int s = 0; s < w.G; ++s) {
        int sg = grp * w.G + s;
        if (w.s_status[sg] == ST_ACTIVE) {
            if (checked && !w.s_fail[sg]) {
                w.s_status[sg] = ST_FROZEN;
                w.s_mconv[sg] = m;
            } else if (m > w.max_coef - 1) {
                w.s_status[sg] = ST_DIVERGED;                                     // helm.py:654-657
            } else {
                nact++;
            }
        }
        w.s_fail[sg] = 0;
    }
    w.g_n[grp] = j;
    if (nact == 0) {
        gf |= GF_PASSEND;
        w.pend_list[atomicAdd(w.pend_count, 1)] = grp;      // k_qlimit works on this list
    }
    w.g_flags[grp] = gf;
}

// ---- K4: Q-limit check on the Pade voltages, result write-out (helm.py:452-490, 985) -----
template <int G>
__global__ void k_qlimit(PlanDev p, WorkDev w) {
    TL_SCOPE(w, 5);
    const int npend = *w.pend_count;
    int flat = blockIdx.x * blockDim.x + threadIdx.x;
    int i = flat / G, s = flat % G;
    if (i >= p.N) return;
  for (int pi = blockIdx.y; pi < npend; pi += gridDim.y) {
    const int grp = w.pend_list[pi];
    int sg = grp * G + s;
    if (w.s_status[sg] != ST_FROZEN) continue;
    const int N = p.N;
    size_t bi = ix<G>(grp, N, i, s);
    double2 vi = w.Pcur[bi];
    const int scen = w.slot_scen[sg];
    if (scen >= 0) w.vout[(size_t)scen * N + p.perm[i]] = vi;
    if (!w.enforce_q) continue;
    if (w.btype[bi] != HELM_BUS_PV) continue;
    const double2 *pc = w.Pcur + (size_t)grp * N * G;
    const OutageRef og = load_outage(p, w, sg);
    double q = 0.0;
    for (int a = p.adj_ptr[i]; a < p.adj_ptr[i + 1]; ++a) {                       // helm.py:461-464
        double2 y = yfull_eff(p, og, a, i, p.adj_idx[a]);
        double2 vk = pc[(size_t)p.adj_idx[a] * G + s];
        q += vi.y * (y.x * vk.x - y.y * vk.y) - vi.x * (y.x * vk.y + y.y * vk.x);
    }
    double qg = q + p.qd[i] * w.scale[sg];                                        // helm.py:481
    double lim = qg;
    if (qg > p.qmax[i] || qg < p.qmin[i]) {                                       // helm.py:483-486
        lim = (qg > p.qmax[i]) ? p.qmax[i] : p.qmin[i];
        w.btype[bi] = HELM_BUS_PQ;
        w.s_viol[sg] = 1;
        const int k = atomicAdd(&w.s_nswitch[sg], 1);
        if (w.sw_ent && k < w.sw_cap) {
            w.sw_ent[(size_t)sg * w.sw_cap + k] = make_int2(p.perm[i], w.s_npass[sg]);
            w.sw_q[(size_t)sg * w.sw_cap + k] = qg;
        }
    }
    w.Qg[bi] = lim;
  }
}

// ---- K_pend: pass bookkeeping (helm.py:936-955, 644-653) ----------------------------------
__global__ void k_pass_end(WorkDev w) {
    TL_SCOPE(w, 6);
    int grp = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp >= w.ngroups) return;
    int gf = w.g_flags[grp];
    if (!(gf & GF_PASSEND)) return;
    int any = 0;
    for (int s = 0; s < w.G; ++s) {
        int sg = grp * w.G + s;
        if (w.s_status[sg] == ST_FROZEN) {
            int np = w.s_npass[sg];
            // beyond HELM_MAX_PASSES recorded passes the last entry always holds the latest pass
            w.s_ncoef[sg * HELM_MAX_PASSES + (np < HELM_MAX_PASSES ? np : HELM_MAX_PASSES - 1)] = w.s_mconv[sg];
            w.s_npass[sg] = np + 1;
            if (w.s_viol[sg] && np + 1 < PASS_CAP) {
                w.s_status[sg] = ST_ACTIVE;
                w.s_viol[sg] = 0;
                any = 1;
            } else {
                // violations still pending at the pass cap: not a solution of the reference's loop
                w.s_status[sg] = w.s_viol[sg] ? ST_PASSCAP : ST_CONVERGED;
            }
        }
    }
    gf &= ~GF_PASSEND;
    if (any) {
        gf |= GF_PENDING;
        w.g_n[grp] = 0;
    } else {
        // every scenario of the group has ended: publish its results, then refill the slots from
        // the scenario queue (continuous batching) or retire the group
        int refilled = 0;
        for (int s = 0; s < w.G; ++s) {
            int sg = grp * w.G + s;
            int scen = w.slot_scen[sg];
            if (scen >= 0) {
                int st = w.s_status[sg];
                w.out_status[scen] = (st == ST_CONVERGED) ? HELM_ST_CONVERGED
                                   : (st == ST_DIVERGED)  ? HELM_ST_DIVERGED
                                   : (st == ST_SINGULAR)  ? HELM_ST_SINGULAR
                                   : (st == ST_PASSCAP)   ? HELM_ST_PASS_LIMIT : HELM_ST_RUNNING;
                // k_qlimit wrote the voltages of every converged pass; a scenario that ends without a
                // solution after an earlier converged pass gets its row zeroed again (rare)
                if (st != ST_CONVERGED && w.s_npass[sg] > 0 && w.vout)
                    for (int i = 0; i < w.nbus; ++i) w.vout[(size_t)scen * w.nbus + i] = make_double2(0.0, 0.0);
                if (w.out_nswitch) w.out_nswitch[scen] = w.s_nswitch[sg];
                if (w.out_ncoef)
                    for (int k = 0; k < HELM_MAX_PASSES; ++k)
                        w.out_ncoef[scen * HELM_MAX_PASSES + k] = w.s_ncoef[sg * HELM_MAX_PASSES + k];
            }
            int next = atomicAdd(w.queue_next, 1);
            if (next < w.B) {
                w.slot_scen[sg] = next;
                w.scale[sg] = w.scale_in[next];
                w.o_b[sg] = w.outage_in ? w.outage_in[next] : -1;
                w.s_status[sg] = ST_ACTIVE;
                w.s_mconv[sg] = 0; w.s_fail[sg] = 0; w.s_viol[sg] = 0; w.s_npass[sg] = 0; w.s_nswitch[sg] = 0;
                for (int k = 0; k < HELM_MAX_PASSES; ++k) w.s_ncoef[sg * HELM_MAX_PASSES + k] = 0;
                refilled = 1;
            } else {
                w.slot_scen[sg] = -1;
                w.s_status[sg] = ST_PAD;
            }
        }
        if (refilled) {
            gf |= GF_PENDING | GF_NEWSCEN;
            w.g_n[grp] = 0;
        } else {
            gf |= GF_DONE;
            atomicSub(w.active_groups, 1);
        }
    }
    w.g_flags[grp] = gf;
}

#include "helm_variants.cuh"
#include "helm_nr.cuh"

// ---- K_fcache: factor cache (see WorkDev::lu_ix).  One CTA, after k_step_begin: every scenario of the
// step's factorization list gives its buffer back and gets the buffer of its new bus-type vector -- an
// existing one (nothing to factorize) or, for a vector not seen before, the buffer that has been unused
// for the longest time, which its scenario then factorizes in the side branch.  Scenarios with a branch
// outage have their own matrix: a buffer of their own, not entered in the table.
enum : int { FC_STEP = 0, FC_HITS = 1, FC_MISS = 2, FC_USED = 3, FC_NEXT = 4 };   // fc_ctr: step counter, acquires served from the table, factorizations, non-empty table slots, first buffer never used
__device__ __forceinline__ unsigned long long fc_mix(unsigned long long x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// slot of `key` in the table, or -1
__device__ __forceinline__ int fc_find(const WorkDev &w, unsigned long long key) {
    int slot = (int)(fc_mix(key) & (unsigned)w.fc_tmask);
    for (int n = 0; n <= w.fc_tmask; ++n) {
        const unsigned long long k = w.fc_tkey[slot];
        if (k == key) return slot;
        if (k == 0) return -1;
        slot = (slot + 1) & w.fc_tmask;
    }
    return -1;
}
__device__ __forceinline__ void fc_insert(const WorkDev &w, unsigned long long key, int buf) {
    int slot = (int)(fc_mix(key) & (unsigned)w.fc_tmask);
    while (w.fc_tkey[slot] > 1) slot = (slot + 1) & w.fc_tmask;          // 0: empty, 1: deleted
    if (w.fc_tkey[slot] == 0) w.fc_ctr[FC_USED] += 1;
    w.fc_tkey[slot] = key;
    w.fc_tbuf[slot] = buf;
}

// A (many CTAs, one warp per scenario of the step's factorization list): give the old buffer back; key of
// the new bus-type vector (0: no buffer needed, 1: a buffer of its own, not shared)
__global__ void __launch_bounds__(256) k_fchash(PlanDev p, WorkDev w) {
    const int fi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (fi >= w.fact_count[w.parity]) return;
    const int grp = w.fact_list[(size_t)w.parity * w.ngroups + fi];
    const int N = p.N;
    if (lane == 0) {
        const int old = w.lu_ix[grp];
        if (old >= 0) {
            if (atomicSub(&w.fc_ref[old], 1) == 1) w.fc_stamp[old] = w.fc_ctr[FC_STEP];
            w.lu_ix[grp] = -1;
        }
        w.fc_do[grp] = 0;
        w.g_fage[grp] = 1;                                   // nothing to wait for unless k_fcache says so
    }
    // a scenario that has just taken over the slot starts from the base bus types (with the cache this
    // happens here, ahead of k_init on the main stream, instead of in the factorization kernel)
    if (w.g_flags[grp] & GF_NEWSCEN) {
        asm_reset<1>(p, w, grp, lane, 32);
        __syncwarp();
    }
    unsigned long long key;
    if (w.s_status[grp] != ST_ACTIVE || uses_shared_lu(w, grp)) key = 0;
    else if (w.o_b[grp] >= 0) key = 1;
    else {
        const unsigned char *bt = w.btype + (size_t)grp * N;
        unsigned long long h = 0;
        int i = lane;
        for (; i + 96 < N; i += 128) {                              // four loads in flight
            const unsigned b0 = bt[i], b1 = bt[i + 32], b2 = bt[i + 64], b3 = bt[i + 96];
            h += fc_mix(((unsigned long long)(i + 1) << 8) | b0) + fc_mix(((unsigned long long)(i + 33) << 8) | b1) +
                 fc_mix(((unsigned long long)(i + 65) << 8) | b2) + fc_mix(((unsigned long long)(i + 97) << 8) | b3);
        }
        for (; i < N; i += 32) h += fc_mix(((unsigned long long)(i + 1) << 8) | bt[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xFFFFFFFFu, h, o);
        key = h < 2 ? h + 2 : h;
    }
    if (lane == 0) w.fc_hash[fi] = key;
}

__global__ void __launch_bounds__(1024) k_fcache(PlanDev p, WorkDev w) {
    const int nfact = w.fact_count[w.parity];
    if (nfact == 0) return;
    const int *list = w.fact_list + (size_t)w.parity * w.ngroups;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int P = w.ngroups;
    __shared__ int s_npend, s_first, s_next;
    __shared__ unsigned long long s_red[32];
    if (tid == 0) s_npend = 0;
    const int now = w.fc_ctr[FC_STEP];                      // (k_step_begin counts the steps)
    // deleted entries pile up when buffers are recycled: rebuild the table from the buffers' keys
    if (w.fc_ctr[FC_USED] > (w.fc_tmask + 1) / 2) {
        __syncthreads();
        for (int i = tid; i <= w.fc_tmask; i += 1024) w.fc_tkey[i] = 0;
        __syncthreads();
        if (tid == 0) {
            w.fc_ctr[FC_USED] = 0;
            for (int b = 0; b < P; ++b)
                if (w.fc_bkey[b] > 1) fc_insert(w, w.fc_bkey[b], b);
        }
    }
    __syncthreads();
    // B: vectors already in the table
    for (int fi = tid; fi < nfact; fi += 1024) {
        const unsigned long long key = w.fc_hash[fi];
        if (key == 0) continue;
        const int slot = key > 1 ? fc_find(w, key) : -1;
        if (slot >= 0) {
            const int buf = w.fc_tbuf[slot];
            atomicAdd(&w.fc_ref[buf], 1);
            w.lu_ix[list[fi]] = buf;
            w.g_fage[list[fi]] = max(1, w.fc_rdy[buf] - now);   // factors in flight (previous step's side branch): wait for its join
            atomicAdd(&w.fc_ctr[FC_HITS], 1);
        } else {
            w.fc_pend[atomicAdd(&s_npend, 1)] = fi;
        }
    }
    __syncthreads();
    const int npend = s_npend;
    // C: new vectors, one per round: the first pending scenario takes a buffer and enters its key; every
    // other pending scenario with that key then finds it
    for (int round = 0; round <= npend; ++round) {
        if (tid == 0) { s_first = 0x7FFFFFFF; s_next = w.fc_ctr[FC_NEXT]; }
        __syncthreads();
        for (int i = tid; i < npend; i += 1024) {
            const int fi = w.fc_pend[i];
            if (fi < 0) continue;
            const unsigned long long key = w.fc_hash[fi];
            const int slot = (round > 0 && key > 1) ? fc_find(w, key) : -1;
            if (slot >= 0) {
                const int buf = w.fc_tbuf[slot];
                atomicAdd(&w.fc_ref[buf], 1);
                w.lu_ix[list[fi]] = buf;
                w.g_fage[list[fi]] = max(1, w.fc_rdy[buf] - now);
                atomicAdd(&w.fc_ctr[FC_HITS], 1);
                w.fc_pend[i] = -1;
            } else {
                atomicMin(&s_first, i);
            }
        }
        __syncthreads();
        const int first = s_first;
        if (first == 0x7FFFFFFF) break;
        // a buffer never used so far, else the free buffer with the oldest release stamp
        int buf = s_next;                                   // (block-uniform: the branch below has barriers)
        if (buf >= P) {
            unsigned long long best = ~0ull;
            for (int b = tid; b < P; b += 1024)
                if (w.fc_ref[b] == 0) best = min(best, ((unsigned long long)(unsigned)w.fc_stamp[b] << 32) | (unsigned)b);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
            if (lane == 0) s_red[wp] = best;
            __syncthreads();
            best = s_red[0];
            for (int k = 1; k < 32; ++k) best = min(best, s_red[k]);
            buf = (int)(best & 0xFFFFFFFFu);                // exists: every scenario holds at most one buffer, and this one holds none
        }
        if (tid == 0) {
            if (buf == s_next && buf < P) w.fc_ctr[FC_NEXT] = buf + 1;
            const int fi = w.fc_pend[first], grp = list[fi];
            const unsigned long long key = w.fc_hash[fi];
            const unsigned long long oldkey = w.fc_bkey[buf];
            if (oldkey > 1) {
                const int os = fc_find(w, oldkey);
                if (os >= 0) w.fc_tkey[os] = 1;
            }
            w.fc_bkey[buf] = key > 1 ? key : 0;
            if (key > 1) fc_insert(w, key, buf);
            w.fc_ref[buf] = 1;
            w.fc_sing[buf] = 0;
            w.fc_rdy[buf] = now + w.depth;                  // its side branch is joined at the end of step now + depth - 1
            w.lu_ix[grp] = buf;
            w.fc_do[grp] = 1;
            w.g_fage[grp] = w.depth;
            w.fc_ctr[FC_MISS] += 1;
            w.fc_pend[first] = -1;
        }
        __syncthreads();
    }
}

// ---- K_begin: start of a step.  Groups factorized during the previous step rejoin the main
// branch; groups whose pass ended in the previous step move to the factorization branch.
__global__ void k_step_begin(WorkDev w) {
#ifdef HELM_TIMELINE
    if (w.tl && blockIdx.x == 0 && threadIdx.x == 0) w.tl[0] += 1;
#endif
    int grp = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp >= w.ngroups) return;
    // the list of the next step starts empty (its previous user, depth + 1 steps ago, has been joined)
    if (grp == 0) {
        w.fact_count[(w.parity + 1) % (w.depth + 1)] = 0;
        w.solve_count[(w.parity + 1) % (w.depth + 1)] = 0;
        *w.pend_count = 0;
        if (w.fc_ctr) w.fc_ctr[FC_STEP] += 1;                // factor cache: step counter
    }
    int gf = w.g_flags[grp];
    if (gf & GF_FACTORING) {
        int age = w.g_fage[grp] - 1;
        w.g_fage[grp] = age;
        if (age <= 0) {
            gf &= ~(GF_FACTORING | GF_NEWSCEN);
            // factor cache: the scenario that factorized the buffer found the matrix singular
            if (w.lu_ix && w.s_status[grp] == ST_ACTIVE && !uses_shared_lu(w, grp)) {
                const int buf = w.lu_ix[grp];
                if (buf >= 0 && w.fc_sing[buf]) w.s_status[grp] = ST_SINGULAR;
            }
        }
    } else if (gf & GF_PENDING) {
        gf = (gf & ~GF_PENDING) | GF_FACTORING;
        w.g_fage[grp] = w.depth;
        int pos = atomicAdd(&w.fact_count[w.parity], 1);
        w.fact_list[(size_t)w.parity * w.ngroups + pos] = grp;
    }
    w.g_flags[grp] = gf;
    if (!(gf & (GF_DONE | GF_FACTORING)) && (w.G > 1 || w.s_status[grp] == ST_ACTIVE)) {
        int pos = atomicAdd(&w.solve_count[w.parity], 1);
        w.solve_list[(size_t)w.parity * w.ngroups + pos] = grp;
    }
}

// ---- shared first-pass factors: scratch control block for factorizing the base matrix once -----
// ints[0..3] = 1 (list length for any list index), ints[4] = 0 (the list: pseudo-group 0),
// ints[8] = -1 (no outage), ints[9] = status (ST_ACTIVE; ST_SINGULAR afterwards if A is singular),
// ints[10] = 0 (flags), ints[11] = 0 (passes)
__global__ void k_shared_prep(int *ints) {
    if (threadIdx.x < 16) ints[threadIdx.x] = (threadIdx.x < 4) ? 1 : (threadIdx.x == 8 ? -1 : 0);
}

// ---- K_setup: initial per-scenario state ---------------------------------------------------
template <int G>
__global__ void k_setup(PlanDev p, WorkDev w, const double *scale_in) {
    int grp = blockIdx.y;
    int flat = blockIdx.x * blockDim.x + threadIdx.x;
    int i = flat / G, s = flat % G;
    int sg = grp * G + s;
    if (i < p.N) {
        size_t bi = ix<G>(grp, p.N, i, s);
        w.btype[bi] = p.btype0[i];
        w.Qg[bi] = 0.0;                                                          // classes.py:252
        w.VVant[bi] = 0.0;
        w.Pcur[bi] = make_double2(0.0, 0.0);
    }
    if (flat < G) {
        bool real = sg < w.B;
        w.scale[sg] = real ? scale_in[sg] : 1.0;
        w.slot_scen[sg] = real ? sg : -1;
        w.o_b[sg] = (real && w.outage_in) ? w.outage_in[sg] : -1;
        w.s_status[sg] = real ? ST_ACTIVE : ST_PAD;
        w.s_mconv[sg] = 0; w.s_fail[sg] = 0; w.s_viol[sg] = 0; w.s_npass[sg] = 0; w.s_nswitch[sg] = 0;
        for (int k = 0; k < HELM_MAX_PASSES; ++k) w.s_ncoef[sg * HELM_MAX_PASSES + k] = 0;
    }
    if (flat == 0) {
        w.g_n[grp] = 0;
        w.g_flags[grp] = GF_PENDING;
        if (grp == 0) {
            *w.active_groups = w.ngroups;
            *w.queue_next = w.ngroups * G;      // slots 0..Bp-1 start with scenarios 0..Bp-1
            for (int k = 0; k <= MAX_SIDE_DEPTH; ++k) { w.fact_count[k] = 0; w.solve_count[k] = 0; }
        }
    }
}

// ---- microbenchmarks for the roofline denominators ------------------------------------------
__global__ void k_dfma_peak(double *out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// dependent-latency probe (one warp): cycles per dependent DADD, DFMA, fp64 shuffle+add round,
// shared-memory load, and L2-hit global load — the building blocks of one level of a triangular solve
__global__ void k_latency_probe(double *out, const int *chase, int n_chase) {
    __shared__ int sm_chase[1024];
    const int lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32) sm_chase[i] = (i + 33) & 1023;
    __syncwarp();
    const int R = 256;
    double a = 1.0 + lane * 1e-9, b = 1.0000001;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; ++i) a = a + b;
    long long t1 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; ++i) a = fma(a, b, b);
    long long t2 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; ++i) a += __shfl_down_sync(0xFFFFFFFFu, a, 1);
    long long t3 = clock64();
    int idx = lane;
#pragma unroll 16
    for (int i = 0; i < R; ++i) idx = sm_chase[idx];
    long long t4 = clock64();
    int g = lane;
    for (int i = 0; i < R; ++i) g = __ldcg(chase + g);          // L2 (chase array is small and was just written)
    long long t5 = clock64();
    if (lane == 0) {
        out[0] = double(t1 - t0) / R; out[1] = double(t2 - t1) / R; out[2] = double(t3 - t2) / R;
        out[3] = double(t4 - t3) / R; out[4] = double(t5 - t4) / R;
        out[5] = a + idx + g;                                   // keep the chains alive
    }
    (void)n_chase;
}
__global__ void k_copy(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) b[i] = a[i];
}

}  // namespace

// =================================================================================================
// Host side
// =================================================================================================
struct helm_plan {
    helm::Symbolic S;
    PlanDev d;
    std::vector<void *> dev_allocs;
    int n_pv = 0;
    int n_factor_updates = 0;
    // cached workspace for the host-buffer entry point
    void *ws = nullptr;
    size_t ws_bytes = 0;
    double *d_scale = nullptr, *d_vout = nullptr;
    int *d_status = nullptr, *d_ncoef = nullptr, *d_nswitch = nullptr;
    int cached_B = 0;
    int *h_pinned = nullptr;
    double *h_vout_pinned = nullptr;
    size_t h_vout_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side[MAX_SIDE_DEPTH] = {};          // side branches of consecutive steps (factorization pipeline)
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_SIDE_DEPTH] = {};
    unsigned long long *tl = nullptr;
    // the instantiated CUDA graph of `poll_every` steps is kept between calls: it is valid as long as every
    // kernel argument (buffers, sizes, parameters, stream) is the same, which `graph_key` records
    cudaGraphExec_t gexec = nullptr;
    cudaGraph_t graph = nullptr;
    std::vector<unsigned long long> graph_key;
    int *d_outage = nullptr;      // pending N-1 list for the next solve (helm_plan_set_outages)
    int outage_n = 0;
    int last_B = 0, last_G = 0, last_maxcoef = 0, last_resident = 0, last_dsb = 0, last_pv = 0, last_rec = 0;   // layout of the last host-buffer solve
};

namespace {

// The N-1 list set by helm_plan_set_outages belongs to exactly one solve call: whatever way that
// call ends (argument error, CUDA error, success), the list is gone afterwards, so that a later call
// without outages can never silently solve N-1 networks.
struct OutageGuard {
    helm_plan *pl;
    explicit OutageGuard(helm_plan *p) : pl(p) {}
    ~OutageGuard() {
        if (pl && pl->d_outage) { cudaFree(pl->d_outage); pl->d_outage = nullptr; }
        if (pl) pl->outage_n = 0;
    }
};

template <typename T>
int upload(helm_plan *pl, const std::vector<T> &v, const T **out) {
    void *d = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CUDA_TRY(cudaMalloc(&d, bytes));
    pl->dev_allocs.push_back(d);
    if (!v.empty()) CUDA_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T *)d;
    return HELM_OK;
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

int pick_group(int n_scen, int requested) {
    if (requested == 1 || requested == 2 || requested == 4 || requested == 8 || requested == 16 || requested == 32)
        return requested;
    (void)n_scen;
    return 1;   // measured fastest at every batch size (profiles/README.md): no lock-step waiting
}

struct WsLayout {
    size_t total = 0;
    size_t off_scale, off_sint, off_gint, off_btype, off_Qg, off_VVant, off_Pcur, off_rhs, off_xb, off_lu, off_Vh, off_Hh, off_Eh;
    size_t off_var_s, off_ploss, off_zb, off_Qh, off_vre, off_qcur, off_lu_shared, off_shprep, off_Pb;
    size_t off_sw_ent, off_sw_q, off_fc;
    int sw_cap, fc_T;
    int G, ngroups, Bp, K;
};

WsLayout layout(const helm_plan *pl, int B, const helm_params *prm) {
    WsLayout L;
    L.G = (prm->dsb_method || prm->pv_bus_model == 1) ? 1 : pick_group(B, prm->group);
    // resident scenarios (slots): the rest of the batch waits in the queue and takes over slots as
    // scenarios finish (continuous batching)
    int resident = prm->max_resident > 0 ? prm->max_resident : 4096;
    if (resident > B) resident = B;
    L.ngroups = (resident + L.G - 1) / L.G;
    L.Bp = L.ngroups * L.G;
    L.K = prm->max_coefficients + 1;
    size_t N = pl->S.N, per = (size_t)L.Bp * N;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes); return r; };
    L.off_scale = take(sizeof(double) * L.Bp);
    L.off_sint = take(sizeof(int) * L.Bp * (8 + HELM_MAX_PASSES));
    L.off_gint = take(sizeof(int) * ((6 + 2 * MAX_SIDE_DEPTH) * L.ngroups + 32));
    L.off_btype = take(per);
    L.off_Qg = take(per * sizeof(double));
    L.off_VVant = take(per * sizeof(double));
    L.off_Pcur = take(per * sizeof(double2));
    L.off_rhs = take(per * sizeof(double2));
    L.off_xb = take(per * sizeof(double2));
    L.off_lu = take((size_t)L.Bp * pl->S.nslots * 2 * sizeof(double2));
    L.off_lu_shared = take((size_t)pl->S.nslots * 2 * sizeof(double2));
    L.off_shprep = take(sizeof(int) * 16);
    const size_t per_h = (size_t)L.Bp * ((N + HT - 1) / HT * HT);   // histories are tiled by HT buses
    L.off_Vh = take(per_h * L.K * sizeof(double2));
    L.off_Hh = take(per_h * L.K * sizeof(double2));
    L.off_Eh = take(per_h * L.K * sizeof(double2));
    L.off_Pb = take(per_h * 3 * sizeof(double2));
    L.off_var_s = L.off_ploss = L.off_zb = L.off_Qh = L.off_vre = L.off_qcur = 0;
    if (prm->dsb_method) {
        L.off_var_s = take(sizeof(double) * 4 * L.Bp);
        L.off_ploss = take(sizeof(double) * (size_t)L.Bp * L.K);
        L.off_zb = take(per * sizeof(double2));
    }
    if (prm->pv_bus_model == 1) {
        L.off_Qh = take(per * L.K * sizeof(double));
        L.off_vre = take(per * sizeof(double));
        L.off_qcur = take(per * sizeof(double));
    }
    // factor cache: [lu_ix | fc_do | fc_ref | fc_stamp | fc_sing | fc_pend | fc_rdy][ngroups] ints, fc_tbuf[T] ints, fc_ctr[16] ints,
    // then [fc_bkey | fc_hash][ngroups] and fc_tkey[T] 64-bit words
    L.fc_T = 1024;
    while (L.fc_T < 4 * L.ngroups) L.fc_T <<= 1;
    L.off_fc = take(sizeof(int) * ((size_t)7 * L.ngroups + L.fc_T + 16) + 8 + sizeof(unsigned long long) * ((size_t)2 * L.ngroups + L.fc_T));
    L.off_sw_ent = L.off_sw_q = 0;
    L.sw_cap = 0;
    if (prm->record_switches) {
        L.sw_cap = std::max(1, pl->n_pv);                    // a bus switches at most once
        L.off_sw_ent = take(sizeof(int2) * (size_t)L.Bp * L.sw_cap);
        L.off_sw_q = take(sizeof(double) * (size_t)L.Bp * L.sw_cap);
    }
    L.total = o;
    return L;
}

WorkDev bind(const helm_plan *pl, const WsLayout &L, int B, const helm_params *prm, void *ws, double *vout) {
    WorkDev w;
    char *b = (char *)ws;
    w.B = B; w.G = L.G; w.ngroups = L.ngroups; w.K = L.K; w.nbus = pl->S.N;
    w.mismatch = prm->mismatch; w.max_coef = prm->max_coefficients; w.enforce_q = prm->enforce_q_limits;
    w.scale = (double *)(b + L.off_scale);
    int *si = (int *)(b + L.off_sint);
    w.s_status = si; w.s_mconv = si + L.Bp; w.s_fail = si + 2 * L.Bp; w.s_viol = si + 3 * L.Bp;
    w.s_npass = si + 4 * L.Bp; w.s_nswitch = si + 5 * L.Bp; w.slot_scen = si + 6 * L.Bp; w.o_b = si + 7 * L.Bp; w.s_ncoef = si + 8 * L.Bp;
    int *gi = (int *)(b + L.off_gint);
    w.g_n = gi; w.g_flags = gi + L.ngroups; w.active_groups = gi + 2 * L.ngroups; w.queue_next = gi + 2 * L.ngroups + 1;
    w.fact_count = gi + 2 * L.ngroups + 2; w.solve_count = gi + 2 * L.ngroups + 16;
    w.g_fage = gi + 2 * L.ngroups + 32; w.fact_list = gi + 3 * L.ngroups + 32;
    w.solve_list = w.fact_list + (size_t)(MAX_SIDE_DEPTH + 1) * L.ngroups;
    w.pend_list = w.solve_list + (size_t)(MAX_SIDE_DEPTH + 1) * L.ngroups; w.pend_count = gi + 2 * L.ngroups + 24; w.parity = 0;
    w.depth = 1; w.skip_asm = 0;
    w.factor_cta = (L.Bp * WPB_FACTOR <= 296 * 2 && !getenv("HELM_B200_NO_FACTOR_CTA")) ? 1 : 0;   // at most ~one CTA per SM
    w.scale_in = nullptr; w.outage_in = nullptr; w.out_status = nullptr; w.out_ncoef = nullptr; w.out_nswitch = nullptr;
    w.btype = (unsigned char *)(b + L.off_btype);
    w.Qg = (double *)(b + L.off_Qg);
    w.VVant = (double *)(b + L.off_VVant);
    w.Pcur = (double2 *)(b + L.off_Pcur);
    w.rhs = (double2 *)(b + L.off_rhs);
    w.xb = (double2 *)(b + L.off_xb);
    w.lu = (double2 *)(b + L.off_lu);
    w.lu_shared = nullptr;
    w.lu_ix = nullptr; w.fc_do = nullptr; w.fc_ref = w.fc_stamp = w.fc_sing = w.fc_rdy = w.fc_tbuf = w.fc_pend = w.fc_ctr = nullptr;
    w.fc_bkey = w.fc_tkey = w.fc_hash = nullptr; w.fc_tmask = 0;
    w.Vh = (double2 *)(b + L.off_Vh);
    w.Hh = (double2 *)(b + L.off_Hh);
    w.Eh = (double2 *)(b + L.off_Eh);
    w.Pb = (double2 *)(b + L.off_Pb);
    w.vout = (double2 *)vout;
    w.tl = pl->tl;
    w.sw_ent = nullptr; w.sw_q = nullptr; w.sw_cap = L.sw_cap;
    if (L.sw_cap) {
        w.sw_ent = (int2 *)(b + L.off_sw_ent);
        w.sw_q = (double *)(b + L.off_sw_q);
    }
    w.pv_model = prm->pv_bus_model == 1 ? 1 : 2;
    w.dsb_method = prm->dsb_method;
    w.dsb_model = prm->dsb_model;
    w.ktot = w.sden = w.beta = w.pcur = w.ploss = nullptr; w.zb = nullptr; w.Qh = w.vre = w.qcur = nullptr;
    if (prm->dsb_method) {
        double *vs = (double *)(b + L.off_var_s);
        w.ktot = vs; w.sden = vs + L.Bp; w.beta = vs + 2 * L.Bp; w.pcur = vs + 3 * L.Bp;
        w.ploss = (double *)(b + L.off_ploss);
        w.zb = (double2 *)(b + L.off_zb);
    }
    if (prm->pv_bus_model == 1) {
        w.Qh = (double *)(b + L.off_Qh);
        w.vre = (double *)(b + L.off_vre);
        w.qcur = (double *)(b + L.off_qcur);
    }
    return w;
}

constexpr int T_ELEM = 256;   // threads per CTA of the per-(item, scenario) kernels
constexpr int WPB_SOLVE = 4;  // warps (= scenario groups) per CTA of the factor / solve kernels

enum { KI_ASM = 0, KI_FACTOR, KI_SOLVE, KI_RHS, KI_ADV, KI_QLIM, KI_PEND, KI_INIT, KI_BEGIN,
       KI_SOLVE_SIDE, KI_BPREP, KI_BFIN, KI_BP, KI_VPRE, KI_PV1C_MAIN, KI_PV1C_SIDE, KI_EPS, KI_FCACHE, KI_COUNT };
// slot of helm_stats.ms_kernel a kernel is accounted under in profile mode (-1: not accounted)
inline int stat_slot(int ki) { return ki < 8 ? ki : (ki == KI_EPS ? 8 : (ki == KI_FCACHE ? 9 : -1)); }

template <int G>
void launch_kernel(int which, const PlanDev &p, const WorkDev &w, cudaStream_t st) {
    dim3 gridBus((p.N * G + T_ELEM - 1) / T_ELEM, w.ngroups);
    int gridGrp = (w.ngroups + 127) / 128;
    switch (which) {
        case KI_ASM: k_assemble<G><<<std::min(w.ngroups, 592), 256, 0, st>>>(p, w); break;
        case KI_BEGIN: k_step_begin<<<gridGrp, 128, 0, st>>>(w); break;
        case KI_FCACHE:
            k_fchash<<<(w.ngroups + 7) / 8, 256, 0, st>>>(p, w);
            k_fcache<<<1, 1024, 0, st>>>(p, w);
            break;
        case KI_FACTOR:
            if (G == 1 && p.use_coop_factor) {
                size_t smem = WPB_FACTOR * factor_smem_per_warp(p.cap_upd, p.cap_ent, p.cap_rowblk);
                const int nctas = w.factor_cta ? w.ngroups : (w.ngroups + WPB_FACTOR - 1) / WPB_FACTOR;
                k_factor_coop<WPB_FACTOR><<<std::min(nctas, 296), WPB_FACTOR * 32, smem, st>>>(p, w);
            } else {
                k_factor<G, WPB_SOLVE><<<(w.ngroups + WPB_SOLVE - 1) / WPB_SOLVE, WPB_SOLVE * 32, 0, st>>>(p, w);
            }
            break;
        case KI_INIT:
            if (w.pv_model == 1 || w.dsb_method) k_init_var<<<gridBus, T_ELEM, 0, st>>>(p, w);
            else k_init<G><<<dim3(gridBus.x, std::min(w.ngroups, 512)), T_ELEM, 0, st>>>(p, w);
            break;
        case KI_BPREP: k_border_prep<<<w.ngroups, 256, 0, st>>>(p, w); break;
        case KI_BFIN: k_border_fin<<<w.ngroups, 256, 0, st>>>(p, w); break;
        case KI_BP: k_border_p<<<gridGrp, 128, 0, st>>>(p, w); break;
        case KI_VPRE: k_var_pre<<<gridBus, T_ELEM, 0, st>>>(p, w); break;
        case KI_PV1C_MAIN: k_pv1_correct<<<gridBus, T_ELEM, 0, st>>>(p, w, 0); break;
        case KI_PV1C_SIDE: k_pv1_correct<<<gridBus, T_ELEM, 0, st>>>(p, w, 1); break;
        case KI_SOLVE_SIDE:
            if (p.use_one_solve && w.ngroups <= p.cta_solve_max_groups) k_solve_one<<<w.ngroups, ONE_T, one_layout(p.N, p.nwaves, p.nsegs, p.nonelist).total, st>>>(p, w, 1);
            else if (p.use_cta_solve && w.ngroups <= p.cta_solve_max_groups) k_solve_cta<<<w.ngroups, CS_T, cta_solve_layout(p.N, p.nwaves, p.nschunks).total, st>>>(p, w, 1);
            else k_solve_flat<FLAT_WPB><<<(w.ngroups + FLAT_WPB - 1) / FLAT_WPB, FLAT_WPB * 32, 0, st>>>(p, w, 1);
            break;
        case KI_SOLVE:
            // small batches: CTA per scenario with x in shared memory (shortest dependency chain);
            // large batches: warp per scenario, all resident scenarios in one round
            if (G == 1 && p.use_rows_deep && w.ngroups <= p.cta_solve_max_groups) {
                k_solve_rows_deep<<<w.ngroups, RC_T, (size_t)p.N * sizeof(double2), st>>>(p, w, 0);
            } else if (G == 1 && p.use_one_solve && (w.ngroups <= p.cta_solve_max_groups || p.use_one_solve == 2)) {
                if (w.ngroups <= p.cta_solve_max_groups) k_solve_one<<<w.ngroups, ONE_T, one_layout(p.N, p.nwaves, p.nsegs, p.nonelist).total, st>>>(p, w, 0);
                else k_solve_one<<<w.ngroups, ONE_T, one_layout(p.N, p.nwaves, p.nsegs, p.nonelist).total, st>>>(p, w, 2);
            } else if (G == 1 && p.use_cta_solve && w.ngroups <= p.cta_solve_max_groups) {
                k_solve_cta<<<w.ngroups, CS_T, cta_solve_layout(p.N, p.nwaves, p.nschunks).total, st>>>(p, w, 0);
            } else if (G == 1 && p.use_rows_cta) {
                k_solve_rows_cta<true><<<w.ngroups, RC_T, (size_t)p.N * sizeof(double2), st>>>(p, w, 2);
            } else if (G == 1 && p.use_rows_solve) {
                k_solve_rows<ROWS_WPB><<<(w.ngroups + ROWS_WPB - 1) / ROWS_WPB, ROWS_WPB * 32, 0, st>>>(p, w);
            } else if (G == 1) {
                k_solve_flat<FLAT_WPB><<<(w.ngroups + FLAT_WPB - 1) / FLAT_WPB, FLAT_WPB * 32, 0, st>>>(p, w, 0);
            } else {
                k_solve<G, WPB_SOLVE><<<(w.ngroups + WPB_SOLVE - 1) / WPB_SOLVE, WPB_SOLVE * 32, 0, st>>>(p, w);
            }
            break;
        case KI_RHS:
            if (w.pv_model == 1 || w.dsb_method) k_rhs_conv_var<<<gridBus, T_ELEM, 0, st>>>(p, w);
            else k_rhs_conv<G><<<dim3((p.N * G + 127) / 128, w.ngroups), 128, 0, st>>>(p, w);
            break;
        case KI_EPS: k_eps<G><<<dim3((p.N * G + EPS_T - 1) / EPS_T, w.ngroups), EPS_T, 0, st>>>(p, w); break;
        case KI_ADV: k_advance<<<gridGrp, 128, 0, st>>>(w); break;
        case KI_QLIM: k_qlimit<G><<<dim3(gridBus.x, std::min(w.ngroups, 512)), T_ELEM, 0, st>>>(p, w); break;
        case KI_PEND: k_pass_end<<<gridGrp, 128, 0, st>>>(w); break;
    }
}

void launch_any(int G, int which, const PlanDev &p, const WorkDev &w, cudaStream_t st) {
    switch (G) {
        case 1: launch_kernel<1>(which, p, w, st); break;
        case 2: launch_kernel<2>(which, p, w, st); break;
        case 4: launch_kernel<4>(which, p, w, st); break;
        case 8: launch_kernel<8>(which, p, w, st); break;
        case 16: launch_kernel<16>(which, p, w, st); break;
        default: launch_kernel<32>(which, p, w, st); break;
    }
}

void launch_setup(int G, const PlanDev &p, const WorkDev &w, const double *d_scale, cudaStream_t st) {
    dim3 gridBus((p.N * G + T_ELEM - 1) / T_ELEM, w.ngroups);
    switch (G) {
        case 1: k_setup<1><<<gridBus, T_ELEM, 0, st>>>(p, w, d_scale); break;
        case 2: k_setup<2><<<gridBus, T_ELEM, 0, st>>>(p, w, d_scale); break;
        case 4: k_setup<4><<<gridBus, T_ELEM, 0, st>>>(p, w, d_scale); break;
        case 8: k_setup<8><<<gridBus, T_ELEM, 0, st>>>(p, w, d_scale); break;
        case 16: k_setup<16><<<gridBus, T_ELEM, 0, st>>>(p, w, d_scale); break;
        default: k_setup<32><<<gridBus, T_ELEM, 0, st>>>(p, w, d_scale); break;
    }
}

// One per-order step is a fork/join:
//
//   k_step_begin ──┬── main branch : solve -> rhs_conv -> advance -> qlimit -> pass_end   (stream st)
//                  └── side branch : assemble -> factor -> init                            (stream side)
//
// Groups that start a Q-limit pass are (re)factorized in the side branch while every other
// group advances one order in the main branch, so the latency of a numeric factorization is
// hidden behind the solves of the rest of the batch; the factorized groups rejoin at the
// next k_step_begin.  Both in eager mode and under stream capture the fork/join is expressed
// with events, so the captured CUDA graph has the two branches as parallel paths.
constexpr int SEQ_MAX = 12;
struct StepSeq {
    int main_seq[SEQ_MAX], side_seq[SEQ_MAX];
    int main_len = 0, side_len = 0;
};
// default: main = solve, rhs_conv, advance, qlimit, pass_end; side = assemble, factor, init.
// Variants (helm_variants.cuh) add the bordered-system kernels and the PV-model-1 fix-ups.
StepSeq make_seq(const WorkDev &w, bool fused_asm) {
    StepSeq q;
    const bool dsb = w.dsb_method != 0, pv1 = w.pv_model == 1;
    // factor cache: most pass starts have nothing to factorize and rejoin after ONE step, so their germ
    // and first right-hand side (k_init) are computed on the main stream, not behind the factorizations
    const bool init_on_main = w.lu_ix != nullptr;
    if (!fused_asm) q.side_seq[q.side_len++] = KI_ASM;     // k_factor_coop assembles A itself
    q.side_seq[q.side_len++] = KI_FACTOR;
    if (dsb) { q.side_seq[q.side_len++] = KI_BPREP; q.side_seq[q.side_len++] = KI_SOLVE_SIDE; q.side_seq[q.side_len++] = KI_BFIN; }
    if (!init_on_main) q.side_seq[q.side_len++] = KI_INIT;
    if (pv1) q.side_seq[q.side_len++] = KI_PV1C_SIDE;
    if (init_on_main) q.main_seq[q.main_len++] = KI_INIT;
    q.main_seq[q.main_len++] = KI_SOLVE;
    if (dsb) q.main_seq[q.main_len++] = KI_BP;
    if (dsb || pv1) q.main_seq[q.main_len++] = KI_VPRE;
    if (!dsb && !pv1) q.main_seq[q.main_len++] = KI_EPS;   // epsilon table + Pade check (k_eps); the variants' kernel does both
    q.main_seq[q.main_len++] = KI_RHS;
    if (pv1) q.main_seq[q.main_len++] = KI_PV1C_MAIN;
    q.main_seq[q.main_len++] = KI_ADV;
    q.main_seq[q.main_len++] = KI_QLIM;
    q.main_seq[q.main_len++] = KI_PEND;
    return q;
}
constexpr int STEP_LEN_MAX = 1 + 2 * SEQ_MAX;

struct StepCtx {
    cudaStream_t st;
    const cudaStream_t *side;      // [depth]
    cudaEvent_t fork;
    const cudaEvent_t *join;       // [depth]
    int depth;
};

void launch_step(int G, const PlanDev &p, const WorkDev &w_in, const StepCtx &c, const StepSeq &q, long step,
                 bool last_of_batch) {
    WorkDev w = w_in;
    const int D = c.depth;
    w.parity = (int)(step % (D + 1));
    static const bool nofork = getenv("HELM_B200_NOFORK") != nullptr;   // A/B switch: serial step
    if (nofork) {
        launch_any(G, KI_BEGIN, p, w, c.st);
        if (w.lu_ix) launch_any(G, KI_FCACHE, p, w, c.st);
        for (int k = 0; k < q.side_len; ++k) launch_any(G, q.side_seq[k], p, w, c.st);
        for (int k = 0; k < q.main_len; ++k) launch_any(G, q.main_seq[k], p, w, c.st);
        return;
    }
    // The side branch of step k has `depth` steps to finish: it is joined at the end of step
    // k + depth - 1, and the groups it factorizes rejoin at the k_step_begin of step k + depth.
    const int sl = (int)(step % D);
    launch_any(G, KI_BEGIN, p, w, c.st);
    if (w.lu_ix) launch_any(G, KI_FCACHE, p, w, c.st);    // on the main stream: the cache is updated by one step at a time
    cudaEventRecord(c.fork, c.st);
    cudaStreamWaitEvent(c.side[sl], c.fork, 0);
    for (int k = 0; k < q.side_len; ++k) launch_any(G, q.side_seq[k], p, w, c.side[sl]);
    cudaEventRecord(c.join[sl], c.side[sl]);
    for (int k = 0; k < q.main_len; ++k) launch_any(G, q.main_seq[k], p, w, c.st);
    if (last_of_batch) {             // end of a graph / of a polled batch of steps: join everything
        for (long k = std::max(0L, step - (D - 1)); k <= step; ++k) cudaStreamWaitEvent(c.st, c.join[k % D], 0);
    } else if (step - (D - 1) >= 0) {
        cudaStreamWaitEvent(c.st, c.join[(step - (D - 1)) % D], 0);
    }
}

// serial variant with an event around every kernel (profile mode)
void launch_step_profiled(int G, const PlanDev &p, const WorkDev &w_in, cudaStream_t st,
                          std::vector<cudaEvent_t> &ev, helm_stats &acc, const StepSeq &q, long step) {
    WorkDev w = w_in;
    w.parity = (int)(step % (w.depth + 1));
    int seq[STEP_LEN_MAX];
    int n = 0;
    seq[n++] = KI_BEGIN;
    if (w.lu_ix) seq[n++] = KI_FCACHE;
    for (int k = 0; k < q.side_len; ++k) seq[n++] = q.side_seq[k];
    for (int k = 0; k < q.main_len; ++k) seq[n++] = q.main_seq[k];
    for (int k = 0; k < n; ++k) {
        cudaEventRecord(ev[k], st);
        launch_any(G, seq[k], p, w, st);
    }
    cudaEventRecord(ev[n], st);
    cudaEventSynchronize(ev[n]);
    for (int k = 0; k < n; ++k) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[k], ev[k + 1]);
        if (stat_slot(seq[k]) >= 0) acc.ms_kernel[stat_slot(seq[k])] += ms;
    }
}

}  // namespace

extern "C" {

const char *helm_last_error(void) { return g_err.c_str(); }

int helm_plan_create(const helm_grid_host *g, helm_plan **out) {
    if (!g || !out) { g_err = "null argument"; return HELM_E_ARG; }
    helm_plan *pl = new helm_plan();
    std::string err;
    if (!helm::build_symbolic(g->n_bus, g->adj_ptr, g->adj_idx, g->bus_type, pl->S, err)) {
        g_err = err;
        delete pl;
        return HELM_E_STRUCT;
    }
    const helm::Symbolic &S = pl->S;
    const int N = S.N;
    PlanDev &d = pl->d;
    d.N = N; d.nL = S.nL; d.nU = S.nU; d.nslots = S.nslots;
    d.nLlev = (int)S.llev_ptr.size() - 1; d.nUlev = (int)S.ulev_ptr.size() - 1;
    d.nadj = S.nnz_adj; d.slack = S.iperm[g->slack];
    {
        double spd = 0.0, spg = 0.0;
        for (int i = 0; i < N; ++i) { spd += g->pd[i]; spg += g->pg[i]; }
        d.pg_imb0 = spd - spg;
    }
    pl->n_factor_updates = (int)S.fsrc.size();
    // permute the grid data into the elimination numbering
    std::vector<double2> ytr(S.nnz_adj), yfull(S.nnz_adj), ysh(N), phv;
    std::vector<int> php(N + 1, 0), phi;
    std::vector<double> vsp(N), pg(N), pd(N), qd(N), qmax(N), qmin(N);
    std::vector<unsigned char> bt(N);
    for (int a = 0; a < S.nnz_adj; ++a) {
        int src = S.adj_src[a];
        ytr[a] = make_double2(g->ytrans[2 * src], g->ytrans[2 * src + 1]);
        yfull[a] = make_double2(g->yfull[2 * src], g->yfull[2 * src + 1]);
    }
    for (int r = 0; r < N; ++r) {
        int o = S.perm[r];
        ysh[r] = make_double2(g->yshunt[2 * o], g->yshunt[2 * o + 1]);
        vsp[r] = g->vsp[o]; pg[r] = g->pg[o]; pd[r] = g->pd[o]; qd[r] = g->qd[o];
        qmax[r] = g->qgmax[o]; qmin[r] = g->qgmin[o]; bt[r] = g->bus_type[o];
        if (bt[r] == HELM_BUS_PV) pl->n_pv++;
        for (int a = g->phase_ptr[o]; a < g->phase_ptr[o + 1]; ++a) {
            phi.push_back(S.iperm[g->phase_idx[a]]);
            phv.push_back(make_double2(g->phase_val[2 * a], g->phase_val[2 * a + 1]));
        }
        php[r + 1] = (int)phi.size();
    }
    // branch table (N-1 scenarios): bus indices and adjacency / phase entries in the new numbering
    std::vector<int> br_f, br_t, br_a, br_p;
    std::vector<double> br_val;
    d.n_branch = (g->n_branch > 0 && g->br_from && g->br_to && g->br_y) ? g->n_branch : 0;
    if (d.n_branch) {
        auto find_adj = [&](int r, int c) -> int {
            for (int a = S.adj_ptr[r]; a < S.adj_ptr[r + 1]; ++a)
                if (S.adj_idx[a] == c) return a;
            return -1;
        };
        auto find_ph = [&](int r, int c) -> int {
            for (int a = php[r]; a < php[r + 1]; ++a)
                if (phi[a] == c) return a;
            return -1;
        };
        for (int b = 0; b < d.n_branch; ++b) {
            int f = S.iperm[g->br_from[b]], t = S.iperm[g->br_to[b]];
            br_f.push_back(f);
            br_t.push_back(t);
            br_a.push_back(find_adj(f, t)); br_a.push_back(find_adj(t, f));
            br_a.push_back(find_adj(f, f)); br_a.push_back(find_adj(t, t));
            br_p.push_back(find_ph(f, t)); br_p.push_back(find_ph(t, f));
            const double *src[5] = {g->br_y, g->br_sh_from, g->br_sh_to, g->br_ph_ft, g->br_ph_tf};
            for (int k = 0; k < 5; ++k) {
                br_val.push_back(src[k] ? src[k][2 * b] : 0.0);
                br_val.push_back(src[k] ? src[k][2 * b + 1] : 0.0);
            }
        }
    }
    int rc;
#define UP(vec, field) if ((rc = upload(pl, vec, &d.field)) != HELM_OK) { helm_plan_destroy(pl); return rc; }
    UP(S.adj_ptr, adj_ptr) UP(S.adj_idx, adj_idx) UP(S.adj_slot, adj_slot)
    UP(ytr, ytr) UP(yfull, yfull) UP(ysh, yshunt)
    UP(php, ph_ptr) UP(phi, ph_idx) UP(phv, ph_val)
    UP(vsp, vsp) UP(pg, pg) UP(pd, pd) UP(qd, qd) UP(qmax, qmax) UP(qmin, qmin) UP(bt, btype0)
    UP(S.perm, perm) UP(S.lptr, lptr) UP(S.lcol, lcol) UP(S.llev_ptr, llev_ptr)
    UP(S.urow, urow) UP(S.ulev_ptr, ulev_ptr) UP(S.uptr, uptr) UP(S.ucol, ucol) UP(S.dslot, dslot)
    UP(S.slot_src, slot_src) UP(S.slot_row, slot_row) UP(S.slot_col, slot_col)
    UP(br_f, br_f) UP(br_t, br_t) UP(br_a, br_a) UP(br_p, br_p) UP(br_val, br_val)
    UP(S.fptr, fptr) UP(S.fsrc, fsrc) UP(S.fdst, fdst) UP(S.fdst_local, fdstl) UP(S.qof, qof)
    UP(S.fchunk_row, fchunk_row) UP(S.fchunk_lpr, fchunk_lpr) UP(S.frow_off, frow_off) UP(S.flev_chunk, flev_chunk)
    UP(S.smeta, smeta)
    {
        const int *wt = nullptr;
        if ((rc = upload(pl, S.wave_tab, &wt)) != HELM_OK) { helm_plan_destroy(pl); return rc; }
        d.wave_tab = (const int2 *)wt;
    }
    d.nwaves = S.n_waves; d.n_rows_nol = S.n_rows_nol;
    UP(S.scm, scm)
    {
        const int *ct = nullptr;
        if ((rc = upload(pl, S.solve_chunks, &ct)) != HELM_OK) { helm_plan_destroy(pl); return rc; }
        d.solve_chunks = (const int4 *)ct;
    }
    d.nschunks = (int)S.solve_chunks.size() / 4;
    {
        const int *rd = nullptr;
        if ((rc = upload(pl, S.rl_desc, &rd)) != HELM_OK) { helm_plan_destroy(pl); return rc; }
        d.rl_desc = (const int4 *)rd;
        d.rl_nwaves = S.rl_nwaves;
        d.use_rows_solve = S.rl_nwaves > 0 && (getenv("HELM_B200_ROWS_SOLVE") ? atoi(getenv("HELM_B200_ROWS_SOLVE")) : 1);
        UP(S.rc_lists, rc_lists) UP(S.rc_segend, rc_segend) UP(S.rc_seg, rc_seg)
        {
            const int *rq = nullptr;
            if ((rc = upload(pl, S.rc_desc, &rq)) != HELM_OK) { helm_plan_destroy(pl); return rc; }
            d.rc_desc = (const int4 *)rq;
        }
        d.rc_nsegs = S.rc_nsegs;
        const size_t xsm = (size_t)N * sizeof(double2);
        d.use_rows_cta = S.rl_nwaves > 0 && xsm <= 225 * 1024 && (getenv("HELM_B200_ROWS_CTA") ? atoi(getenv("HELM_B200_ROWS_CTA")) : 0);
        d.use_rows_deep = S.rl_nwaves > 0 && xsm <= 225 * 1024 && (getenv("HELM_B200_ROWS_DEEP") ? atoi(getenv("HELM_B200_ROWS_DEEP")) : 1);
        static size_t xsm_set = 48 * 1024;
        if (S.rl_nwaves > 0 && xsm <= 225 * 1024 && xsm > xsm_set) {
            CUDA_TRY(cudaFuncSetAttribute(k_solve_rows_cta<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsm));
            CUDA_TRY(cudaFuncSetAttribute(k_solve_rows_cta<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsm));
            CUDA_TRY(cudaFuncSetAttribute(k_solve_rows_deep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsm));
            xsm_set = xsm;
        }
    }
    {
        const int *sg = nullptr;
        if ((rc = upload(pl, S.solve_segs, &sg)) != HELM_OK) { helm_plan_destroy(pl); return rc; }
        d.solve_segs = (const int2 *)sg;
        d.nsegs = (int)S.solve_segs.size() / 2;
        if ((rc = upload(pl, S.one_lists, &d.one_lists)) != HELM_OK) { helm_plan_destroy(pl); return rc; }
        d.nonelist = (int)S.one_lists.size();
    }
    d.max_row_blocks = S.max_row_blocks;
    d.nfchunks = (int)S.fchunk_lpr.size(); d.cap_upd = S.cap_upd; d.cap_ent = S.cap_ent; d.cap_rowblk = S.cap_rowblk;
#undef UP
    {
        size_t smem = WPB_FACTOR * factor_smem_per_warp(pl->d.cap_upd, pl->d.cap_ent, pl->d.cap_rowblk);
        // staged factorization needs its chunk operands in shared memory; grids whose longest row
        // does not fit fall back to the unstaged warp-per-group kernel
        if (getenv("HELM_B200_VERBOSE"))
            fprintf(stderr, "[helm_b200] factor staging: cap_upd %d cap_ent %d cap_rowblk %d chunks %d -> %zu B per warp, %d warps per CTA\n",
                    pl->d.cap_upd, pl->d.cap_ent, pl->d.cap_rowblk, pl->d.nfchunks, smem / WPB_FACTOR, WPB_FACTOR);
        pl->d.use_coop_factor = (smem <= 200 * 1024) && getenv("HELM_B200_NO_COOP_FACTOR") == nullptr;
        if (!pl->d.use_coop_factor) smem = 0;
        // shared-memory wave solve: x, the wave/chunk tables and the ring must fit one CTA
        {
            size_t csm = cta_solve_layout(pl->d.N, pl->d.nwaves, pl->d.nschunks).total;
            static size_t csm_set = 48 * 1024;
            pl->d.use_cta_solve = (csm <= 225 * 1024) && getenv("HELM_B200_NO_CTA_SOLVE") == nullptr;
            pl->d.cta_solve_max_groups = getenv("HELM_B200_CTA_SOLVE_MAX") ? atoi(getenv("HELM_B200_CTA_SOLVE_MAX")) : 640;   // measured crossover with the warp-per-scenario solve: between 512 and 800 scenarios (tools/r2_mid.sh)
            {
                const size_t osm = one_layout(pl->d.N, pl->d.nwaves, pl->d.nsegs, pl->d.nonelist).total;
                static size_t osm_set = 48 * 1024;
                pl->d.use_one_solve = (osm <= 225 * 1024) && (getenv("HELM_B200_ONE_SOLVE") ? atoi(getenv("HELM_B200_ONE_SOLVE")) : 1);
                if (pl->d.use_one_solve && osm > osm_set) {
                    CUDA_TRY(cudaFuncSetAttribute(k_solve_one, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)osm));
                    osm_set = osm;
                }
            }
            if (pl->d.use_cta_solve && csm > csm_set) {
                CUDA_TRY(cudaFuncSetAttribute(k_solve_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
                csm_set = csm;
            }
        }
        // the attribute is per function, not per plan: only ever raise it
        static size_t smem_set = 48 * 1024;
        if (smem > smem_set) {
            CUDA_TRY(cudaFuncSetAttribute(k_factor_coop<WPB_FACTOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (getenv("HELM_B200_FACTOR_CARVEOUT")) CUDA_TRY(cudaFuncSetAttribute(k_factor_coop<WPB_FACTOR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            smem_set = smem;
        }
    }
    *out = pl;
    return HELM_OK;
}

void helm_plan_destroy(helm_plan *pl) {
    if (!pl) return;
    for (void *p : pl->dev_allocs) cudaFree(p);
    if (pl->ws) cudaFree(pl->ws);
    if (pl->d_scale) cudaFree(pl->d_scale);
    if (pl->d_vout) cudaFree(pl->d_vout);
    if (pl->d_status) cudaFree(pl->d_status);
    if (pl->d_ncoef) cudaFree(pl->d_ncoef);
    if (pl->d_nswitch) cudaFree(pl->d_nswitch);
    if (pl->h_pinned) cudaFreeHost(pl->h_pinned);
    if (pl->h_vout_pinned) cudaFreeHost(pl->h_vout_pinned);
    if (pl->d_outage) cudaFree(pl->d_outage);
    if (pl->gexec) cudaGraphExecDestroy(pl->gexec);
    if (pl->graph) cudaGraphDestroy(pl->graph);
    if (pl->stream) cudaStreamDestroy(pl->stream);
    for (int k = 0; k < MAX_SIDE_DEPTH; ++k) {
        if (pl->side[k]) cudaStreamDestroy(pl->side[k]);
        if (pl->ev_join[k]) cudaEventDestroy(pl->ev_join[k]);
    }
    if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
    delete pl;
}

int helm_plan_get_info(const helm_plan *pl, helm_plan_info *info) {
    if (!pl || !info) { g_err = "null argument"; return HELM_E_ARG; }
    info->n_bus = pl->S.N; info->nnz_adj = pl->S.nnz_adj;
    info->n_l_blocks = pl->S.nL; info->n_u_blocks = pl->S.nU; info->n_slots = pl->S.nslots;
    info->n_l_levels = pl->d.nLlev; info->n_u_levels = pl->d.nUlev;
    info->n_factor_updates = pl->n_factor_updates; info->n_pv = pl->n_pv; info->reserved = 0;
    return HELM_OK;
}

int helm_plan_get_perm(const helm_plan *pl, int32_t *perm_out) {
    if (!pl || !perm_out) { g_err = "null argument"; return HELM_E_ARG; }
    std::copy(pl->S.perm.begin(), pl->S.perm.end(), perm_out);
    return HELM_OK;
}

int helm_plan_set_outages(helm_plan *pl, int32_t n, const int32_t *h_outage) {
    if (!pl || n <= 0 || !h_outage) { g_err = "null/invalid argument"; return HELM_E_ARG; }
    if (pl->d.n_branch <= 0) { g_err = "the plan was created without a branch table"; return HELM_E_ARG; }
    for (int b = 0; b < n; ++b)
        if (h_outage[b] < -1 || h_outage[b] >= pl->d.n_branch) { g_err = "outage branch index out of range"; return HELM_E_ARG; }
    if (pl->d_outage) { cudaFree(pl->d_outage); pl->d_outage = nullptr; }
    CUDA_TRY(cudaMalloc((void **)&pl->d_outage, sizeof(int) * n));
    CUDA_TRY(cudaMemcpy(pl->d_outage, h_outage, sizeof(int) * n, cudaMemcpyHostToDevice));
    pl->outage_n = n;
    return HELM_OK;
}

size_t helm_workspace_bytes(const helm_plan *pl, int32_t n_scen, const helm_params *prm) {
    if (!pl || !prm || n_scen <= 0) return 0;
    return layout(pl, n_scen, prm).total;
}

// per-scenario input / output device buffers of the host-buffer entry points, kept in the plan
static int ensure_io_buffers(helm_plan *pl, int B) {
    if (pl->cached_B >= B) return HELM_OK;
    if (pl->d_scale) { cudaFree(pl->d_scale); cudaFree(pl->d_vout); cudaFree(pl->d_status); cudaFree(pl->d_ncoef); cudaFree(pl->d_nswitch); }
    pl->d_scale = nullptr; pl->d_vout = nullptr; pl->d_status = nullptr; pl->d_ncoef = nullptr; pl->d_nswitch = nullptr;
    pl->cached_B = 0;
    const size_t N = pl->S.N;
    CUDA_TRY(cudaMalloc((void **)&pl->d_scale, sizeof(double) * B));
    CUDA_TRY(cudaMalloc((void **)&pl->d_vout, sizeof(double) * 2 * N * B));
    CUDA_TRY(cudaMalloc((void **)&pl->d_status, sizeof(int) * B));
    CUDA_TRY(cudaMalloc((void **)&pl->d_ncoef, sizeof(int) * B * HELM_MAX_PASSES));
    CUDA_TRY(cudaMalloc((void **)&pl->d_nswitch, sizeof(int) * B));
    pl->cached_B = B;
    return HELM_OK;
}

static int check_params(const helm_params *prm) {
    if (!prm) { g_err = "null params"; return HELM_E_ARG; }
    if (prm->max_coefficients < 5) { g_err = "max_coefficients must be >= 5 (helm.py:882-884)"; return HELM_E_ARG; }
    if (!(prm->mismatch > 0)) { g_err = "mismatch must be positive"; return HELM_E_ARG; }
    if (prm->pv_bus_model != 0 && prm->pv_bus_model != 1 && prm->pv_bus_model != 2) { g_err = "pv_bus_model must be 1 or 2"; return HELM_E_ARG; }
    if (prm->dsb_method < 0 || prm->dsb_method > 2) { g_err = "dsb_method must be 0, 1 or 2"; return HELM_E_ARG; }
    if (prm->dsb_model && !prm->dsb_method) { g_err = "dsb_model needs dsb_method 1 or 2"; return HELM_E_ARG; }
    return HELM_OK;
}

int helm_solve_batch_device(helm_plan *pl, int32_t B, const double *d_scale, const helm_params *prm,
                            void *d_ws, size_t ws_bytes, double *d_vout, int32_t *d_status,
                            int32_t *d_ncoef, int32_t *d_nswitch, void *stream_, helm_stats *stats) {
    if (!pl) { g_err = "null plan"; return HELM_E_ARG; }
    OutageGuard outage_guard(pl);                            // the pending N-1 list ends with this call
    if (!d_scale || !d_ws || !d_vout || !d_status || B <= 0) { g_err = "null/invalid argument"; return HELM_E_ARG; }
    int rc = check_params(prm);
    if (rc) return rc;
    WsLayout L = layout(pl, B, prm);
    if (ws_bytes < L.total) { g_err = "workspace too small"; return HELM_E_NOMEM; }
    cudaStream_t st = (cudaStream_t)stream_;
    WorkDev w = bind(pl, L, B, prm, d_ws, d_vout);
    if (pl->d_outage) {
        if (pl->outage_n != B) { g_err = "the outage list was set for a different number of scenarios"; return HELM_E_ARG; }
        if (prm->dsb_method || prm->pv_bus_model == 1) { g_err = "outage scenarios need pv_bus_model 2 and the classic slack"; return HELM_E_ARG; }
        w.outage_in = pl->d_outage;
    }
    w.scale_in = d_scale; w.out_status = d_status; w.out_ncoef = d_ncoef; w.out_nswitch = d_nswitch;
    const PlanDev &p = pl->d;
    const bool profile = getenv("HELM_B200_PROFILE") != nullptr;
    const bool use_graph = prm->use_graph && !profile;
    // Depth of the factorization pipeline: with depth 2 a numeric factorization (2-3 ms of dependent
    // chunks for one warp) has two steps to finish and is never the critical path of a step; the
    // price is one more idle step per pass for the scenario being factorized.  Kernels that select
    // their groups by flag instead of by list (variants, generic G > 1) need depth 1.
    int depth = (L.Bp < 256) ? 1 : 2;     // a handful of scenarios: nothing to hide the factorization behind
    if (getenv("HELM_B200_SIDE_DEPTH")) depth = std::max(1, std::min(MAX_SIDE_DEPTH, atoi(getenv("HELM_B200_SIDE_DEPTH"))));
    if (!(L.G == 1 && p.use_coop_factor) || prm->dsb_method || prm->pv_bus_model == 1) depth = 1;
    w.depth = depth;
    // steps per graph launch / per poll: a multiple of depth * (depth + 1) so that list and stream
    // indices repeat from one graph launch to the next
    const int cyc = depth * (depth + 1);
    // steps per graph launch and host poll: steps after the last scenario has ended only launch kernels
    // that return at once (~10 us per step), a poll costs a stream synchronisation
    int poll_every = prm->poll_every > 0 ? prm->poll_every : 12;
    poll_every = (poll_every + cyc - 1) / cyc * cyc;
    if (!pl->h_pinned) CUDA_TRY(cudaMallocHost((void **)&pl->h_pinned, 64));

#ifdef HELM_TIMELINE
    {
        size_t n = 8 + (size_t)TL_MAXSTEPS * 16 * 2;
        if (!pl->tl) CUDA_TRY(cudaMalloc((void **)&pl->tl, n * sizeof(unsigned long long)));
        std::vector<unsigned long long> init(n, 0);
        for (size_t i = 8; i < n; i += 2) init[i] = ~0ull;
        CUDA_TRY(cudaMemcpy(pl->tl, init.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        w.tl = pl->tl;
    }
#endif
    cudaEvent_t ev0, ev1;
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    CUDA_TRY(cudaEventRecord(ev0, st));
    CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int) * B, st));           // HELM_ST_RUNNING until a scenario ends
    CUDA_TRY(cudaMemsetAsync(d_vout, 0, sizeof(double) * 2 * (size_t)pl->S.N * B, st));   // rows of scenarios without a solution stay zero
    if (d_ncoef) CUDA_TRY(cudaMemsetAsync(d_ncoef, 0, sizeof(int) * B * HELM_MAX_PASSES, st));
    if (d_nswitch) CUDA_TRY(cudaMemsetAsync(d_nswitch, 0, sizeof(int) * B, st));
    launch_setup(L.G, p, w, d_scale, st);

    // Shared first-pass factors (default variant, staged factorization): assemble and factorize the
    // base matrix once, as pseudo-group 0 of a control block that points at the base bus types.
    if (L.G == 1 && p.use_coop_factor && !prm->dsb_method && prm->pv_bus_model != 1 && L.Bp >= 64 &&
        getenv("HELM_B200_NO_SHARED_LU") == nullptr) {                 // (a handful of scenarios: not worth the extra factorization)
        int *ints = (int *)((char *)d_ws + L.off_shprep);
        WorkDev w2 = w;
        w2.lu = (double2 *)((char *)d_ws + L.off_lu_shared);
        w2.btype = const_cast<unsigned char *>(p.btype0);
        w2.fact_count = ints; w2.fact_list = ints + 4; w2.parity = 0;
        w2.o_b = ints + 8; w2.s_status = ints + 9; w2.g_flags = ints + 10; w2.s_npass = ints + 11;
        w2.ngroups = 1;
        k_shared_prep<<<1, 32, 0, st>>>(ints);
        launch_any(1, KI_FACTOR, p, w2, st);
        if (!pl->h_pinned) CUDA_TRY(cudaMallocHost((void **)&pl->h_pinned, 64));
        CUDA_TRY(cudaMemcpyAsync(pl->h_pinned, ints + 9, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (pl->h_pinned[0] == ST_ACTIVE) w.lu_shared = w2.lu;      // singular base matrix: every scenario factorizes (and fails) itself
    }
    // Factor cache for the later passes (same conditions; HELM_B200_NO_FCACHE: A/B switch)
    if (w.lu_shared && getenv("HELM_B200_NO_FCACHE") == nullptr) {
        int *fi = (int *)((char *)d_ws + L.off_fc);
        const size_t ng = (size_t)L.ngroups;
        w.lu_ix = fi; w.fc_do = fi + ng; w.fc_ref = fi + 2 * ng; w.fc_stamp = fi + 3 * ng; w.fc_sing = fi + 4 * ng; w.fc_pend = fi + 5 * ng;
        w.fc_rdy = fi + 6 * ng;
        w.fc_tbuf = fi + 7 * ng; w.fc_ctr = w.fc_tbuf + L.fc_T;
        const size_t nints = 7 * ng + L.fc_T + 16;
        unsigned long long *fq = (unsigned long long *)((char *)d_ws + L.off_fc + (nints * sizeof(int) + 7) / 8 * 8);
        w.fc_bkey = fq; w.fc_hash = fq + ng; w.fc_tkey = fq + 2 * ng;
        w.fc_tmask = L.fc_T - 1;
        CUDA_TRY(cudaMemsetAsync(fi, 0, (char *)(w.fc_tkey + L.fc_T) - (char *)fi, st));
        CUDA_TRY(cudaMemsetAsync(w.lu_ix, 0xFF, sizeof(int) * ng, st));
    }

    helm_stats local;
    memset(&local, 0, sizeof(local));
    local.kernel_launches = 1;

    if (!pl->side[0]) {
        // The side branch has few, long-running CTAs: with the higher priority they take a free slot
        // before the next CTA of the (much larger) main-branch grids does.  HELM_B200_SIDE_PRIO=0: A/B.
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const char *sp = getenv("HELM_B200_SIDE_PRIO");
        for (int k = 0; k < MAX_SIDE_DEPTH; ++k) {
            CUDA_TRY(cudaStreamCreateWithPriority(&pl->side[k], cudaStreamNonBlocking, (sp && atoi(sp) == 0) ? lo : hi));
            CUDA_TRY(cudaEventCreateWithFlags(&pl->ev_join[k], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
    }
    const StepCtx ctx{st, pl->side, pl->ev_fork, pl->ev_join, depth};
    const StepSeq seq = make_seq(w, L.G == 1 && p.use_coop_factor);
    const int STEP_LEN = 1 + seq.main_len + seq.side_len + (w.lu_ix ? 2 : 0);
    // `count` steps starting at step index `first`; every side branch is joined back into st
    auto launch_batch = [&](long first, int count) {
        for (int rep = 0; rep < count; ++rep) launch_step(L.G, p, w, ctx, seq, first + rep, rep == count - 1);
    };
    cudaGraphExec_t gexec = nullptr;
    if (use_graph) {
        auto u = [](const void *ptr) { return (unsigned long long)(uintptr_t)ptr; };
        unsigned long long mm;
        memcpy(&mm, &prm->mismatch, sizeof(mm));
        const std::vector<unsigned long long> key = {
            u(d_ws), (unsigned long long)ws_bytes, (unsigned long long)B, u(d_scale), u(d_vout), u(d_status), u(d_ncoef), u(d_nswitch),
            u(st), u(w.outage_in), u(w.lu_shared), u(w.lu_ix), u(w.tl), mm, (unsigned long long)prm->max_coefficients,
            (unsigned long long)prm->enforce_q_limits, (unsigned long long)L.G, (unsigned long long)L.ngroups,
            (unsigned long long)prm->pv_bus_model, (unsigned long long)prm->dsb_method, (unsigned long long)prm->dsb_model,
            (unsigned long long)prm->record_switches, (unsigned long long)depth, (unsigned long long)poll_every,
            (unsigned long long)L.total, (unsigned long long)p.use_one_solve, (unsigned long long)p.cta_solve_max_groups};
        if (!pl->gexec || pl->graph_key != key) {
            if (pl->gexec) { cudaGraphExecDestroy(pl->gexec); pl->gexec = nullptr; }
            if (pl->graph) { cudaGraphDestroy(pl->graph); pl->graph = nullptr; }
            CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            launch_batch(0, poll_every);
            CUDA_TRY(cudaStreamEndCapture(st, &pl->graph));
            CUDA_TRY(cudaGraphInstantiate(&pl->gexec, pl->graph, 0));
            pl->graph_key = key;
        }
        gexec = pl->gexec;
    }
    std::vector<cudaEvent_t> pev;
    if (profile) {
        pev.resize(STEP_LEN_MAX + 1);
        for (auto &e : pev) cudaEventCreate(&e);
    }
    const long max_steps = (long)(prm->max_coefficients + 2 + 2 * depth) * 4 * HELM_MAX_PASSES * ((B + L.Bp - 1) / L.Bp);
    long steps = 0;
    int active = 1;
    while (active > 0 && steps < max_steps) {
        if (use_graph) {
            CUDA_TRY(cudaGraphLaunch(gexec, st));
            local.graph_launches++;
        } else if (profile) {
            for (int rep = 0; rep < poll_every; ++rep) launch_step_profiled(L.G, p, w, st, pev, local, seq, steps + rep);
        } else {
            launch_batch(steps, poll_every);
        }
        local.kernel_launches += (int64_t)poll_every * STEP_LEN;
        steps += poll_every;
        CUDA_TRY(cudaMemcpyAsync(pl->h_pinned, w.active_groups, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        active = pl->h_pinned[0];
    }
    CUDA_TRY(cudaEventRecord(ev1, st));
    CUDA_TRY(cudaEventSynchronize(ev1));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    local.ms_total = ms;
    local.steps = steps;
    if (w.lu_ix) {                                          // factor cache: acquires served from the table / factorizations
        int ctr[4] = {0, 0, 0, 0};
        CUDA_TRY(cudaMemcpy(ctr, w.fc_ctr, sizeof(ctr), cudaMemcpyDeviceToHost));
        local.ms_kernel[10] = ctr[FC_HITS];
        local.ms_kernel[11] = ctr[FC_MISS];
    }
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    for (auto &e : pev) cudaEventDestroy(e);
    CUDA_TRY(cudaGetLastError());
    if (stats) *stats = local;
    return HELM_OK;
}

int helm_solve_batch_host(helm_plan *pl, int32_t B, const double *h_scale, const helm_params *prm,
                          double *h_vout, int32_t *h_status, int32_t *h_ncoef, int32_t *h_nswitch,
                          helm_stats *stats) {
    if (!pl) { g_err = "null plan"; return HELM_E_ARG; }
    OutageGuard outage_guard(pl);
    if (!h_scale || !h_vout || !h_status || B <= 0) { g_err = "null/invalid argument"; return HELM_E_ARG; }
    int rc = check_params(prm);
    if (rc) return rc;
    helm_params prm_local = *prm;
    if (prm_local.max_resident <= 0) {
        // auto: 4096 resident scenarios, fewer if the workspace would not fit the free device memory
        // (big grids / many coefficients: 40 MB per scenario at 13 656 buses and 60 coefficients)
        prm_local.max_resident = 4096;
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        free_b += pl->ws_bytes;                                  // the cached workspace is reused or replaced
        const size_t out_b = (size_t)B * (pl->S.N * 16 + 64);    // d_vout etc. for the whole batch
        const size_t budget = free_b > out_b ? (size_t)((free_b - out_b) * 0.9) : 0;
        while (prm_local.max_resident > 64 && layout(pl, B, &prm_local).total > budget) prm_local.max_resident /= 2;
    }
    prm = &prm_local;
    WsLayout L = layout(pl, B, prm);
    if (!pl->stream) CUDA_TRY(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    if (pl->ws_bytes < L.total) {
        if (pl->ws) cudaFree(pl->ws);
        pl->ws = nullptr; pl->ws_bytes = 0;
        CUDA_TRY(cudaMalloc(&pl->ws, L.total));
        pl->ws_bytes = L.total;
    }
    size_t N = pl->S.N;
    rc = ensure_io_buffers(pl, B);
    if (rc) return rc;
    cudaStream_t st = pl->stream;
    CUDA_TRY(cudaMemcpyAsync(pl->d_scale, h_scale, sizeof(double) * B, cudaMemcpyHostToDevice, st));
    rc = helm_solve_batch_device(pl, B, pl->d_scale, prm, pl->ws, pl->ws_bytes, pl->d_vout, pl->d_status,
                                 pl->d_ncoef, pl->d_nswitch, st, stats);
    if (rc) return rc;
    pl->last_B = B; pl->last_G = L.G; pl->last_maxcoef = prm->max_coefficients; pl->last_resident = prm->max_resident; pl->last_dsb = prm->dsb_method; pl->last_pv = prm->pv_bus_model; pl->last_rec = prm->record_switches;
    CUDA_TRY(cudaMemcpyAsync(h_vout, pl->d_vout, sizeof(double) * 2 * N * B, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h_status, pl->d_status, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    if (h_ncoef) CUDA_TRY(cudaMemcpyAsync(h_ncoef, pl->d_ncoef, sizeof(int) * B * HELM_MAX_PASSES, cudaMemcpyDeviceToHost, st));
    if (h_nswitch) CUDA_TRY(cudaMemcpyAsync(h_nswitch, pl->d_nswitch, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return HELM_OK;
}

int helm_nr_batch_host(helm_plan *pl, int32_t B, const double *h_scale, double mismatch, int32_t max_iterations,
                       int32_t enforce_q_limits, double *h_vout, int32_t *h_status, int32_t *h_iters,
                       int32_t *h_nswitch) {
    if (!pl) { g_err = "null plan"; return HELM_E_ARG; }
    OutageGuard outage_guard(pl);                            // NR has no outage scenarios: a pending list is dropped
    if (!h_scale || !h_vout || !h_status || B <= 0) { g_err = "null/invalid argument"; return HELM_E_ARG; }
    if (!(mismatch > 0) || max_iterations < 1) { g_err = "mismatch must be positive and max_iterations >= 1"; return HELM_E_ARG; }
    if (!pl->d.use_coop_factor) { g_err = "grid too large for the staged factorization used by the GPU Newton-Raphson"; return HELM_E_STRUCT; }
    const PlanDev &p = pl->d;
    const size_t N = pl->S.N;
    if (!pl->stream) CUDA_TRY(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    cudaStream_t st = pl->stream;
    if (!pl->h_pinned) CUDA_TRY(cudaMallocHost((void **)&pl->h_pinned, 64));
    const int CH = 4096;                                    // scenarios per chunk (all resident)
    helm_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.mismatch = mismatch; prm.max_coefficients = 5; prm.group = 1; prm.max_resident = CH;
    const int nmax = std::min<int>(B, CH);
    WsLayout L = layout(pl, nmax, &prm);
    if (pl->ws_bytes < L.total) {
        if (pl->ws) cudaFree(pl->ws);
        pl->ws = nullptr; pl->ws_bytes = 0;
        CUDA_TRY(cudaMalloc(&pl->ws, L.total));
        pl->ws_bytes = L.total;
    }
    pl->last_B = 0;                                         // the HELM state of the workspace is gone
    int rc = ensure_io_buffers(pl, nmax);                   // the plan's cached per-scenario buffers
    if (rc) return rc;
    double *d_scale = pl->d_scale, *d_vout = pl->d_vout;
    for (int b0 = 0; b0 < B && rc == HELM_OK; b0 += CH) {
        const int nb = std::min(CH, B - b0);
        L = layout(pl, nb, &prm);
        WorkDev w = bind(pl, L, nb, &prm, pl->ws, d_vout);
        w.scale_in = d_scale;
        w.skip_asm = 1;
        int *d_status = pl->d_status, *d_nsw = pl->d_nswitch, *d_iters = pl->d_ncoef;
        cudaMemcpyAsync(d_scale, h_scale + b0, sizeof(double) * nb, cudaMemcpyHostToDevice, st);
        dim3 gridBus((p.N + T_ELEM - 1) / T_ELEM, w.ngroups);
        const int gridGrp = (w.ngroups + 127) / 128;
        k_nr_init<<<gridBus, T_ELEM, 0, st>>>(p, w);
        const long max_rounds = (long)(max_iterations + 2) * 4 * HELM_MAX_PASSES;
        int active = 1;
        for (long round = 0; round < max_rounds && active > 0; ++round) {
            cudaMemsetAsync(w.active_groups, 0, sizeof(int), st);
            cudaMemsetAsync(w.fact_count, 0, sizeof(int), st);
            k_nr_eval<<<gridBus, T_ELEM, 0, st>>>(p, w);
            k_nr_decide<<<gridGrp, 128, 0, st>>>(w, mismatch, max_iterations);
            if (enforce_q_limits) k_nr_qlim<<<gridBus, T_ELEM, 0, st>>>(p, w);
            k_nr_passend<<<gridGrp, 128, 0, st>>>(w, enforce_q_limits);
            cudaMemcpyAsync(pl->h_pinned, w.active_groups, sizeof(int), cudaMemcpyDeviceToHost, st);
            k_nr_assemble<<<std::min(w.ngroups, 592), 256, 0, st>>>(p, w);
            launch_any(1, KI_FACTOR, p, w, st);
            launch_any(1, KI_SOLVE_SIDE, p, w, st);
            k_nr_update<<<gridBus, T_ELEM, 0, st>>>(p, w);
            if (cudaStreamSynchronize(st) != cudaSuccess) { g_err = "CUDA failure in the Newton-Raphson step"; rc = HELM_E_CUDA; break; }
            active = pl->h_pinned[0];
        }
        if (rc != HELM_OK) break;
        k_nr_out<<<gridBus, T_ELEM, 0, st>>>(p, w, (double2 *)d_vout, d_status, h_iters ? d_iters : nullptr, h_nswitch ? d_nsw : nullptr);
        cudaMemcpyAsync(h_vout + (size_t)b0 * N * 2, d_vout, sizeof(double) * 2 * N * nb, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(h_status + b0, d_status, sizeof(int) * nb, cudaMemcpyDeviceToHost, st);
        if (h_iters) cudaMemcpyAsync(h_iters + (size_t)b0 * HELM_MAX_PASSES, d_iters, sizeof(int) * nb * HELM_MAX_PASSES, cudaMemcpyDeviceToHost, st);
        if (h_nswitch) cudaMemcpyAsync(h_nswitch + b0, d_nsw, sizeof(int) * nb, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { g_err = "CUDA failure in the Newton-Raphson batch"; rc = HELM_E_CUDA; }
    }
    return rc;
}

int helm_last_state_host(helm_plan *pl, int32_t B, double *h_qg, uint8_t *h_bus_type) {
    if (!pl || B <= 0 || B != pl->last_B || !pl->ws) { g_err = "no matching host solve to read state from"; return HELM_E_ARG; }
    helm_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.mismatch = 1.0; prm.max_coefficients = pl->last_maxcoef; prm.group = pl->last_G; prm.max_resident = pl->last_resident; prm.dsb_method = pl->last_dsb; prm.pv_bus_model = pl->last_pv;
    prm.record_switches = pl->last_rec;
    WsLayout L = layout(pl, B, &prm);
    if (L.Bp < B) { g_err = "per-bus state is only kept when every scenario has its own slot (n_scen <= max_resident)"; return HELM_E_ARG; }
    const size_t N = pl->S.N, per = (size_t)L.Bp * N;
    std::vector<double> qg(per);
    std::vector<unsigned char> bt(per);
    CUDA_TRY(cudaMemcpy(qg.data(), (char *)pl->ws + L.off_Qg, per * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(bt.data(), (char *)pl->ws + L.off_btype, per, cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b) {
        int grp = b / L.G, s = b % L.G;
        for (size_t r = 0; r < N; ++r) {
            int o = pl->S.perm[r];
            size_t bi = ((size_t)grp * N + r) * L.G + s;
            if (h_qg) h_qg[(size_t)b * N + o] = qg[bi];
            if (h_bus_type) h_bus_type[(size_t)b * N + o] = bt[bi];
        }
    }
    return HELM_OK;
}

int helm_last_series_host(helm_plan *pl, int32_t B, int32_t scen, int32_t n_terms, double *h_V, double *h_H) {
    if (!pl || B <= 0 || B != pl->last_B || !pl->ws || scen < 0 || scen >= B || n_terms <= 0 || (!h_V && !h_H)) {
        g_err = "no matching host solve to read the series from";
        return HELM_E_ARG;
    }
    if (pl->last_G != 1) { g_err = "series read-back needs group = 1"; return HELM_E_ARG; }
    helm_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.mismatch = 1.0; prm.max_coefficients = pl->last_maxcoef; prm.group = 1; prm.max_resident = pl->last_resident;
    prm.dsb_method = pl->last_dsb; prm.pv_bus_model = pl->last_pv; prm.record_switches = pl->last_rec;
    WsLayout L = layout(pl, B, &prm);
    if (L.Bp < B) { g_err = "per-scenario state is only kept when every scenario has its own slot"; return HELM_E_ARG; }
    if (n_terms > L.K) { g_err = "n_terms exceeds max_coefficients + 1"; return HELM_E_ARG; }
    const size_t N = pl->S.N, ntile = (N + HT - 1) / HT, per = ntile * L.K * HT;
    std::vector<double2> buf(per);
    for (int which = 0; which < 2; ++which) {
        double *dst = which ? h_H : h_V;
        if (!dst) continue;
        const size_t off = which ? L.off_Hh : L.off_Vh;
        CUDA_TRY(cudaMemcpy(buf.data(), (char *)pl->ws + off + (size_t)scen * per * sizeof(double2), per * sizeof(double2), cudaMemcpyDeviceToHost));
        for (int k = 0; k < n_terms; ++k)
            for (size_t r = 0; r < N; ++r) {
                const double2 v = buf[((r / HT) * L.K + k) * HT + (r % HT)];
                const size_t o = (size_t)k * N + pl->S.perm[r];
                dst[2 * o] = v.x;
                dst[2 * o + 1] = v.y;
            }
    }
    return HELM_OK;
}

int helm_last_switches_host(helm_plan *pl, int32_t B, int32_t cap, int32_t *h_bus, int32_t *h_pass, double *h_q) {
    if (!pl || !h_bus || !h_pass || !h_q || B <= 0 || cap <= 0 || B != pl->last_B || !pl->ws || !pl->last_rec) {
        g_err = "no matching host solve with record_switches to read from";
        return HELM_E_ARG;
    }
    helm_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.mismatch = 1.0; prm.max_coefficients = pl->last_maxcoef; prm.group = pl->last_G; prm.max_resident = pl->last_resident;
    prm.dsb_method = pl->last_dsb; prm.pv_bus_model = pl->last_pv; prm.record_switches = 1;
    WsLayout L = layout(pl, B, &prm);
    if (L.Bp < B) { g_err = "per-scenario state is only kept when every scenario has its own slot"; return HELM_E_ARG; }
    std::vector<int2> ent((size_t)L.Bp * L.sw_cap);
    std::vector<double> q((size_t)L.Bp * L.sw_cap);
    std::vector<int> nsw(L.Bp);
    CUDA_TRY(cudaMemcpy(ent.data(), (char *)pl->ws + L.off_sw_ent, ent.size() * sizeof(int2), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(q.data(), (char *)pl->ws + L.off_sw_q, q.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(nsw.data(), (char *)pl->ws + L.off_sint + sizeof(int) * 5 * (size_t)L.Bp, sizeof(int) * L.Bp, cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < cap; ++k) {
            const bool have = k < nsw[b] && k < L.sw_cap;
            const size_t src = (size_t)b * L.sw_cap + k, dst = (size_t)b * cap + k;
            h_bus[dst] = have ? ent[src].x : -1;
            h_pass[dst] = have ? ent[src].y : -1;
            h_q[dst] = have ? q[src] : 0.0;
        }
    return HELM_OK;
}

int helm_last_ploss_host(helm_plan *pl, int32_t B, int32_t dsb_method, int32_t pv_bus_model, double *h_ploss) {
    if (!pl || !h_ploss || B <= 0 || B != pl->last_B || !pl->ws || !dsb_method) { g_err = "no matching distributed-slack host solve"; return HELM_E_ARG; }
    helm_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.mismatch = 1.0; prm.max_coefficients = pl->last_maxcoef; prm.group = 1; prm.max_resident = pl->last_resident;
    prm.dsb_method = dsb_method; prm.pv_bus_model = pv_bus_model; prm.record_switches = pl->last_rec;
    WsLayout L = layout(pl, B, &prm);
    if (L.Bp < B) { g_err = "per-scenario state is only kept when every scenario has its own slot"; return HELM_E_ARG; }
    CUDA_TRY(cudaMemcpy(h_ploss, (char *)pl->ws + L.off_ploss, sizeof(double) * (size_t)B * L.K, cudaMemcpyDeviceToHost));
    return HELM_OK;
}

int helm_debug_factor_solve(helm_plan *pl, int32_t B, int32_t group, const uint8_t *h_bt,
                            const double *h_rhs, double *h_x) {
    if (!pl || !h_bt || !h_rhs || !h_x || B <= 0) { g_err = "null/invalid argument"; return HELM_E_ARG; }
    helm_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.mismatch = 1e-8; prm.max_coefficients = 5; prm.group = group;
    WsLayout L = layout(pl, B, &prm);
    void *ws = nullptr;
    double *d_scale = nullptr;
    CUDA_TRY(cudaMalloc(&ws, L.total));
    CUDA_TRY(cudaMalloc((void **)&d_scale, sizeof(double) * B));
    std::vector<double> ones(B, 1.0);
    CUDA_TRY(cudaMemcpy(d_scale, ones.data(), sizeof(double) * B, cudaMemcpyHostToDevice));
    WorkDev w = bind(pl, L, B, &prm, ws, nullptr);
    const PlanDev &p = pl->d;
    const helm::Symbolic &S = pl->S;
    launch_setup(L.G, p, w, d_scale, 0);
    CUDA_TRY(cudaDeviceSynchronize());
    size_t N = S.N, per = (size_t)L.Bp * N;
    std::vector<unsigned char> bt(per, HELM_BUS_PQ);
    std::vector<double2> rhs(per, make_double2(0, 0));
    for (int b = 0; b < B; ++b) {
        int grp = b / L.G, s = b % L.G;
        for (size_t r = 0; r < N; ++r) {
            int o = S.perm[r];
            size_t bi = ((size_t)grp * N + r) * L.G + s;
            bt[bi] = h_bt[(size_t)b * N + o];
            rhs[bi] = make_double2(h_rhs[((size_t)b * N + o) * 2], h_rhs[((size_t)b * N + o) * 2 + 1]);
        }
    }
    CUDA_TRY(cudaMemcpy(w.btype, bt.data(), per, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(w.rhs, rhs.data(), per * sizeof(double2), cudaMemcpyHostToDevice));
    launch_any(L.G, KI_BEGIN, p, w, 0);   // PENDING -> FACTORING (list of parity 0)
    launch_any(L.G, KI_ASM, p, w, 0);
    launch_any(L.G, KI_FACTOR, p, w, 0);
    w.parity = 1;
    launch_any(L.G, KI_BEGIN, p, w, 0);   // FACTORING -> main branch
    launch_any(L.G, KI_SOLVE, p, w, 0);
    CUDA_TRY(cudaDeviceSynchronize());
    std::vector<double2> x(per);
    CUDA_TRY(cudaMemcpy(x.data(), w.xb, per * sizeof(double2), cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b) {
        int grp = b / L.G, s = b % L.G;
        for (size_t r = 0; r < N; ++r) {
            int o = S.perm[r];
            double2 v = x[((size_t)grp * N + r) * L.G + s];
            h_x[((size_t)b * N + o) * 2] = v.x;
            h_x[((size_t)b * N + o) * 2 + 1] = v.y;
        }
    }
    cudaFree(ws);
    cudaFree(d_scale);
    return HELM_OK;
}

int helm_debug_solve_time(helm_plan *pl, int32_t kernel, int32_t ablation, int32_t reps, double *ms_per_solve) {
    // one scenario, base bus types, random right-hand side: factorize once, then time `reps` launches
    // of one triangular-solve kernel (0: k_solve_one, 1: k_solve_cta, 2: k_solve_flat, 3: k_solve_rows_cta,
    // 4: k_solve_rows) with CUDA events
    if (!pl || !ms_per_solve || reps <= 0) { g_err = "null/invalid argument"; return HELM_E_ARG; }
    helm_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.mismatch = 1e-8; prm.max_coefficients = 5; prm.group = 1;
    WsLayout L = layout(pl, 1, &prm);
    void *ws = nullptr;
    double *d_scale = nullptr;
    CUDA_TRY(cudaMalloc(&ws, L.total));
    CUDA_TRY(cudaMalloc((void **)&d_scale, sizeof(double)));
    const double one = 1.0;
    CUDA_TRY(cudaMemcpy(d_scale, &one, sizeof(double), cudaMemcpyHostToDevice));
    WorkDev w = bind(pl, L, 1, &prm, ws, nullptr);
    const PlanDev &p = pl->d;
    launch_setup(1, p, w, d_scale, 0);
    std::vector<double2> rhs(p.N);
    for (int i = 0; i < p.N; ++i) rhs[i] = make_double2(sin(0.37 * i), cos(0.11 * i));
    CUDA_TRY(cudaMemcpy(w.rhs, rhs.data(), rhs.size() * sizeof(double2), cudaMemcpyHostToDevice));
    launch_any(1, KI_BEGIN, p, w, 0);
    launch_any(1, KI_ASM, p, w, 0);
    launch_any(1, KI_FACTOR, p, w, 0);
    w.parity = 1;
    launch_any(1, KI_BEGIN, p, w, 0);
    auto launch = [&]() {
        if (kernel == 0) k_solve_one<<<1, ONE_T, one_layout(p.N, p.nwaves, p.nsegs, p.nonelist).total>>>(p, w, ablation << 8);
        else if (kernel == 1) k_solve_cta<<<1, CS_T, cta_solve_layout(p.N, p.nwaves, p.nschunks).total>>>(p, w, 0);
        else if (kernel == 3 && p.rl_nwaves > 0) {
            k_solve_rows_cta<false><<<1, RC_T, (size_t)p.N * sizeof(double2)>>>(p, w, ablation << 8);
        }
        else if (kernel == 4 && p.rl_nwaves > 0) k_solve_rows<ROWS_WPB><<<1, ROWS_WPB * 32>>>(p, w);
        else if (kernel == 5 && p.rl_nwaves > 0) k_solve_rows_deep<<<1, RC_T, (size_t)p.N * sizeof(double2)>>>(p, w, 0);
        else k_solve_flat<FLAT_WPB><<<1, FLAT_WPB * 32>>>(p, w, 0);
    };
    for (int i = 0; i < 3; ++i) launch();
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_per_solve = ms / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(ws); cudaFree(d_scale);
    CUDA_TRY(cudaGetLastError());
    return HELM_OK;
}

int helm_debug_timeline(helm_plan *pl, uint64_t *out, int32_t max_steps) {
#ifdef HELM_TIMELINE
    if (!pl || !pl->tl || !out) { g_err = "no timeline"; return HELM_E_ARG; }
    int n = std::min<int>(max_steps, TL_MAXSTEPS);
    CUDA_TRY(cudaMemcpy(out, pl->tl + 8, (size_t)n * 16 * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    {   // cycle counters tl[1..7] are returned in the (unused) kernel slots 8..14 of step 0
        unsigned long long c[8];
        CUDA_TRY(cudaMemcpy(c, pl->tl, sizeof(c), cudaMemcpyDeviceToHost));
        for (int i = 1; i < 8; ++i) out[(0 * 16 + 7 + i) * 2] = c[i];
    }
    return HELM_OK;
#else
    (void)pl; (void)out; (void)max_steps;
    g_err = "library built without -DHELM_TIMELINE";
    return HELM_E_ARG;
#endif
}

int helm_measure_latencies(double *out5) {
    if (!out5) { g_err = "null argument"; return HELM_E_ARG; }
    double *d_out = nullptr;
    int *d_chase = nullptr;
    const int n = 1 << 14;
    std::vector<int> h(n);
    for (int i = 0; i < n; ++i) h[i] = (i + 4099) % n;        // stride of ~16 KB: different lines, L2-resident
    CUDA_TRY(cudaMalloc((void **)&d_out, 8 * sizeof(double)));
    CUDA_TRY(cudaMalloc((void **)&d_chase, n * sizeof(int)));
    CUDA_TRY(cudaMemcpy(d_chase, h.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    for (int rep = 0; rep < 2; ++rep) k_latency_probe<<<1, 32>>>(d_out, d_chase, n);
    CUDA_TRY(cudaDeviceSynchronize());
    double r[6];
    CUDA_TRY(cudaMemcpy(r, d_out, sizeof(r), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 5; ++i) out5[i] = r[i];
    cudaFree(d_out); cudaFree(d_chase);
    return HELM_OK;
}

int helm_measure_peaks(double *dfma_tflops, double *copy_gbs) {
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float ms = 0;
    if (dfma_tflops) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int blocks = sms * 8, threads = 256, iters = 1 << 16;
        double *out = nullptr;
        CUDA_TRY(cudaMalloc((void **)&out, sizeof(double) * blocks * threads));
        k_dfma_peak<<<blocks, threads>>>(out, 1024);
        double best = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k_dfma_peak<<<blocks, threads>>>(out, iters);
            cudaEventRecord(e1);
            CUDA_TRY(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * 8 * (double)iters * blocks * threads;
            best = std::max(best, fl / (ms * 1e-3) / 1e12);
        }
        *dfma_tflops = best;
        cudaFree(out);
    }
    if (copy_gbs) {
        size_t n = (size_t)1 << 26;  // 64 Mi double2 = 1 GiB each way
        double2 *a = nullptr, *b = nullptr;
        CUDA_TRY(cudaMalloc((void **)&a, n * sizeof(double2)));
        CUDA_TRY(cudaMalloc((void **)&b, n * sizeof(double2)));
        cudaMemset(a, 0, n * sizeof(double2));
        double best = 0;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            k_copy<<<148 * 16, 512>>>(a, b, n);
            cudaEventRecord(e1);
            CUDA_TRY(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
            best = std::max(best, 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9);
        }
        *copy_gbs = best;
        cudaFree(a);
        cudaFree(b);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return HELM_OK;
}

}  // extern "C"
