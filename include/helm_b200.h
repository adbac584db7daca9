/*
 * helm_b200.h — C ABI of the B200-native HELM + Pade hot path.
 *
 * The reference (HELMpy, pure Python) has no FFI; these entry points are what a
 * binding for its hot path would call.  Each one names the reference code it
 * replaces (paths relative to the reference repository):
 *
 *   helm_plan_create        helmpy/core/helm.py:30-107  modif_Ytrans  (matrix
 *                           structure + scipy `factorized` symbolic part) and
 *                           helmpy/core/classes.py:228-297 RunVariables (state)
 *   helm_solve_batch_*      helmpy/core/helm.py:896-985 helm(): Q-limit pass
 *                           loop (936-955), coefficient loop
 *                           computing_voltages_mismatch (528-659), RHS
 *                           evaluators (293-398), solve (602), V/W update
 *                           (402-433), Pade + convergence (608-641,
 *                           analytic_continuation.py:29-50), Q limits (452-490)
 *
 * Conventions: plain pointers and sizes only.  All functions return 0 on
 * success or a negative HELM_E_* code; helm_last_error() gives the message.
 * Complex numbers are interleaved (re, im) doubles.  Everything is fp64.
 */
#ifndef HELM_B200_H
#define HELM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HELM_OK 0
#define HELM_E_ARG (-1)
#define HELM_E_CUDA (-2)
#define HELM_E_NOMEM (-3)
#define HELM_E_STRUCT (-4)

/* per-scenario status written by the solver */
#define HELM_ST_RUNNING 0
#define HELM_ST_CONVERGED 2   /* helm() would return the voltage vector        */
#define HELM_ST_DIVERGED 3    /* helm() would return None (helm.py:654-657)    */
#define HELM_ST_SINGULAR 4    /* (numerically) singular 2x2 pivot in the numeric factorization */
#define HELM_ST_PASS_LIMIT 5  /* Q-limit violations still pending after 4*HELM_MAX_PASSES passes */

#define HELM_MAX_PASSES 16    /* Q-limit passes recorded in ncoef[]; with more passes the last
                                 entry holds the most recent pass                 */

/* bus type codes (reference strings 'PQ', 'PV'/'PVLIM', 'Slack') */
#define HELM_BUS_PQ 0
#define HELM_BUS_PV 1
#define HELM_BUS_SLACK 2

/* Host description of one grid: the fields of the reference's CaseData that
 * helm() consumes (classes.py:180-225), network in CSR over the sorted bus
 * adjacency `branches_buses` (diagonal included). */
typedef struct helm_grid_host {
    int32_t n_bus;
    int32_t slack;              /* bus index of the slack                       */
    const int32_t *adj_ptr;     /* [n_bus+1]                                    */
    const int32_t *adj_idx;     /* [nnz] ascending per row                      */
    const double *ytrans;       /* [nnz] complex: series-admittance Laplacian   */
    const double *yfull;        /* [nnz] complex: Ytrans + diag(Yshunt) + phase */
    const double *yshunt;       /* [n_bus] complex                              */
    const int32_t *phase_ptr;   /* [n_bus+1] phase-shifter corrections per bus  */
    const int32_t *phase_idx;   /* [nph] neighbour bus                          */
    const double *phase_val;    /* [nph] complex                                */
    const double *vsp;          /* [n_bus] specified |V| (generator buses)      */
    const double *pg;           /* [n_bus] p.u.                                 */
    const double *pd;           /* [n_bus]                                      */
    const double *qd;           /* [n_bus]                                      */
    const double *qgmax;        /* [n_bus]                                      */
    const double *qgmin;        /* [n_bus]                                      */
    const uint8_t *bus_type;    /* [n_bus] HELM_BUS_*                           */
    /* optional: what every branch contributed to the network (classes.py:9-113), needed only for
     * N-1 branch-outage scenarios; n_branch = 0 disables them */
    int32_t n_branch;
    const int32_t *br_from;     /* [n_branch] bus index                         */
    const int32_t *br_to;       /* [n_branch]                                   */
    const double *br_y;         /* [n_branch] complex: series admittance stamped into Ytrans */
    const double *br_sh_from;   /* [n_branch] complex: added to Yshunt[from]    */
    const double *br_sh_to;     /* [n_branch] complex: added to Yshunt[to]      */
    const double *br_ph_ft;     /* [n_branch] complex: phase term of (from, to) */
    const double *br_ph_tf;     /* [n_branch] complex: phase term of (to, from) */
} helm_grid_host;

typedef struct helm_params {
    double mismatch;            /* helm(mismatch=)                              */
    int32_t max_coefficients;   /* helm(max_coefficients=)                      */
    int32_t enforce_q_limits;   /* helm(enforce_Q_limits=)                      */
    int32_t group;              /* scenarios per lock-step group: 1,2,4,8,16,32; 0 = auto */
    int32_t use_graph;          /* replay the per-order step as a CUDA graph    */
    int32_t poll_every;         /* steps between host polls of the done counter; 0 = auto */
    int32_t max_resident;       /* scenarios resident on the GPU at once (slots); the rest of the
                                   batch queues and takes over slots as scenarios end; 0 = auto
                                   (4096, fewer when the workspace would not fit the free memory;
                                   the device-pointer call sizes its workspace for 4096) */
    int32_t pv_bus_model;       /* helm(pv_bus_model=): 1 or 2; 0 = 2 (the reference default)   */
    int32_t dsb_method;         /* helm(DSB_model_method=): 0 = classic slack, 1 or 2            */
    int32_t dsb_model;          /* helm(DSB_model=): participation factors (needs dsb_method)    */
    int32_t record_switches;    /* keep a log of the PV->PQ switches (bus, pass, violating Q) for
                                   helm_last_switches_host: helm(detailed_run_print=True), helm.py:489 */
} helm_params;

typedef struct helm_plan_info {
    int32_t n_bus;
    int32_t nnz_adj;
    int32_t n_l_blocks;         /* strictly-lower 2x2 blocks of L               */
    int32_t n_u_blocks;         /* strictly-upper 2x2 blocks of U               */
    int32_t n_slots;            /* n_l + n_u + n_bus                            */
    int32_t n_l_levels;
    int32_t n_u_levels;
    int32_t n_factor_updates;   /* 2x2 block updates per numeric factorization  */
    int32_t n_pv;               /* generator (PV) buses in the base case        */
    int32_t reserved;
} helm_plan_info;

typedef struct helm_plan helm_plan;

/* per-solve timing/launch statistics (filled by helm_solve_batch_*) */
typedef struct helm_stats {
    int64_t steps;              /* per-order steps executed                     */
    int64_t kernel_launches;    /* kernels launched (graph nodes count)         */
    int64_t graph_launches;
    double ms_total;            /* device time of the whole solve (CUDA events) */
    double ms_kernel[12];       /* per-kernel accumulated ms when profile=1: assemble, factor, solve, rhs_conv, advance, qlimit, pass_end, init, eps, fcache;
                                   [10], [11]: factor cache counters of the call (always filled): pass starts that found their
                                   factors in the cache / factorizations performed */
} helm_stats;

const char *helm_last_error(void);

/* Host symbolic analysis: minimum-degree ordering of the bus graph, symbolic
 * 2x2-block LU fill on the all-PQ superset pattern, level sets, factorization
 * schedule; uploads the shared index/grid arrays to the current CUDA device. */
int helm_plan_create(const helm_grid_host *grid, helm_plan **out);
void helm_plan_destroy(helm_plan *plan);
int helm_plan_get_info(const helm_plan *plan, helm_plan_info *info);
/* elimination order: perm[new index] = original bus index; buffer of n_bus */
int helm_plan_get_perm(const helm_plan *plan, int32_t *perm_out);

/* N-1 scenarios: branch removed in scenario b of the NEXT helm_solve_batch_* call with the same
 * n_scen (h_outage[b] = branch index into the grid's branch table, or -1 for none).  The list is
 * consumed by that call.  Only the default variant (pv_bus_model 2, classic slack). */
int helm_plan_set_outages(helm_plan *plan, int32_t n_scen, const int32_t *h_outage);

/* bytes of device workspace needed for n_scen scenarios */
size_t helm_workspace_bytes(const helm_plan *plan, int32_t n_scen, const helm_params *params);

/* Solve n_scen load-scaling scenarios of the plan's grid.  Scenario b is
 * helm(case, scale=scale[b], ...) (classes.py:215-219: Pd, Qd, Pg *= scale).
 * Device-pointer variant: every pointer is device memory of the current
 * device; work is enqueued on `stream` (a cudaStream_t) and the call returns
 * after the last poll, with results complete on the stream.
 *   d_scale    [n_scen] double
 *   d_vout     [n_scen * n_bus] complex, scenario-major, original bus order; zeroed by the call,
 *              rows of scenarios that end without a solution stay zero
 *   d_status   [n_scen] int32 HELM_ST_*
 *   d_ncoef    [n_scen * HELM_MAX_PASSES] int32: coefficients per pass (run.list_coef), 0-padded
 *   d_nswitch  [n_scen] int32: PV->PQ switches
 */
int helm_solve_batch_device(helm_plan *plan, int32_t n_scen, const double *d_scale,
                            const helm_params *params, void *d_workspace, size_t workspace_bytes,
                            double *d_vout, int32_t *d_status, int32_t *d_ncoef,
                            int32_t *d_nswitch, void *stream, helm_stats *stats);

/* Host-buffer variant (the drop-in call): copies scale in, results out; owns
 * and caches its device workspace inside the plan. */
int helm_solve_batch_host(helm_plan *plan, int32_t n_scen, const double *h_scale,
                          const helm_params *params, double *h_vout, int32_t *h_status,
                          int32_t *h_ncoef, int32_t *h_nswitch, helm_stats *stats);

/* Newton-Raphson cross-check, batched (reference helmpy/core/nr.py:506-598 `nr`, classic slack; host
 * restatement helmpy_b200/core/nr.py): polar NR from the flat start for n_scen load scales on the
 * plan's LU pattern and kernels; converged when max|dP, dQ| <= mismatch; PVLIM buses that violate a
 * Q limit after a converged pass become PQ at the limit and the iteration continues (nr.py:409-436).
 * h_vout [n_scen*n_bus] complex (zeros unless converged), h_status HELM_ST_*, h_iters
 * [n_scen*HELM_MAX_PASSES] Newton steps per pass PLUS ONE, 0 = pass not run (may be NULL), h_nswitch
 * [n_scen] (may be NULL). */
int helm_nr_batch_host(helm_plan *plan, int32_t n_scen, const double *h_scale, double mismatch,
                       int32_t max_iterations, int32_t enforce_q_limits, double *h_vout,
                       int32_t *h_status, int32_t *h_iters, int32_t *h_nswitch);

/* Per-bus state at the end of the last helm_solve_batch_host call with the same n_scen
 * (reference run.Qg and run.Buses_type, classes.py:239, 252; needed by the result report,
 * helm.py:670-759).  h_qg [n_scen*n_bus] p.u., h_bus_type [n_scen*n_bus] HELM_BUS_*;
 * original bus order; either pointer may be NULL. */
int helm_last_state_host(helm_plan *plan, int32_t n_scen, double *h_qg, uint8_t *h_bus_type);

/* Series coefficients of the LAST pass of scenario `scen` of the last helm_solve_batch_host call
 * (group 1, every scenario resident): h_V[k*n_bus + i] = V_i[k] (reference run.V_complex[i][k],
 * classes.py:262), h_H[k*n_bus + i] = W_i[k] for buses that ended as PQ (run.W, classes.py:263) and
 * barras_CC_i[k] for buses that ended as PV (helm.py:310-313), k < n_terms <= max_coefficients + 1;
 * complex, original bus order; either pointer may be NULL.  Per-order parity tests. */
int helm_last_series_host(helm_plan *plan, int32_t n_scen, int32_t scen, int32_t n_terms,
                          double *h_V, double *h_H);

/* PV->PQ switches of the last helm_solve_batch_host call made with params.record_switches
 * (reference check_PVLIM_violation, helm.py:468-490, whose detailed print is helm.py:489): for
 * scenario b, entries k < min(nswitch[b], cap): h_bus[b*cap+k] = original bus index, h_pass = Q-limit
 * pass (0-based) in which it switched, h_q = the reactive generation that violated the limit (p.u.);
 * unused entries are -1 / -1 / 0.  Entries of one pass are in no particular order. */
int helm_last_switches_host(helm_plan *plan, int32_t n_scen, int32_t cap, int32_t *h_bus,
                            int32_t *h_pass, double *h_q);

/* Ploss series of the last distributed-slack helm_solve_batch_host call (reference
 * run.coefficients[2N][:], used by the report, helm.py:963-965): h_ploss [n_scen*(max_coefficients+1)]. */
int helm_last_ploss_host(helm_plan *plan, int32_t n_scen, int32_t dsb_method, int32_t pv_bus_model,
                         double *h_ploss);

/* Debug/testing hooks: run one kernel stage in isolation on device buffers.
 * Used by the parity tests to compare individual stages with the oracle. */
int helm_debug_factor_solve(helm_plan *plan, int32_t n_scen, int32_t group,
                            const uint8_t *h_bus_type /* [n_scen*n_bus] original order */,
                            const double *h_rhs /* [n_scen*2*n_bus] original order */,
                            double *h_x /* [n_scen*2*n_bus] */);

/* Microbenchmark: device time (ms) of one triangular solve of one scenario of the plan's grid with
 * one kernel alone on the GPU: kernel 0 = k_solve_one, 1 = k_solve_cta, 2 = k_solve_flat, 3 = k_solve_rows_cta,
 * 4 = k_solve_rows, 5 = k_solve_rows_deep.  `ablation`
 * (k_solve_one and k_solve_rows_cta only; 0 = the real kernel) switches parts of the kernel off for timing: 1 no shuffle
 * reduction, 2 no operand copies after the first, 4 no gather of x, 8 no block barriers, 16 no thin
 * segments, 32 no wide segments — results are then wrong, only the time is meaningful. */
int helm_debug_solve_time(helm_plan *plan, int32_t kernel, int32_t ablation, int32_t reps, double *ms_per_solve);

/* Kernel timeline of the last solve (only in builds with -DHELM_TIMELINE): for step s and kernel
 * k (index as in helm_stats.ms_kernel), out[(s*16+k)*2+{0,1}] = first-CTA start / last-CTA end
 * on the GPU global timer (ns); untouched entries are {~0, 0}. */
int helm_debug_timeline(helm_plan *plan, uint64_t *out, int32_t max_steps);

/* Host-only view of the symbolic analysis (no CUDA): ordering, fill, level
 * sets and schedules as int32 arrays, by name (see csrc/symbolic.h). */
typedef struct helm_symbolic helm_symbolic;
int helm_symbolic_create(int32_t n_bus, const int32_t *adj_ptr, const int32_t *adj_idx,
                         const uint8_t *bus_type /* may be NULL */, helm_symbolic **out);
void helm_symbolic_destroy(helm_symbolic *sym);
int helm_symbolic_array(const helm_symbolic *sym, const char *name, const int32_t **data,
                        int32_t *len);

/* measured fp64 FMA throughput microbenchmark (TFLOP/s) and copy bandwidth (GB/s) */
int helm_measure_peaks(double *dfma_tflops, double *copy_gbs);
/* Dependent-operation latencies in cycles, one warp alone: out5 = {DADD, DFMA, fp64 shuffle + add,
 * shared-memory load, L2-hit global load} — what one level of a triangular solve is made of. */
int helm_measure_latencies(double *out5);

#ifdef __cplusplus
}
#endif
#endif /* HELM_B200_H */
