/*
 * revs_admm.h -- C ABI of the B200-native REVS distributed EV-charging ADMM path.
 *
 * Drop-in boundary: these are the entry points a maintainer of the reference
 * (rounak-meyur/revs-admm, pure Python + Gurobi) would bind with ctypes to replace the
 * hot path of lpsolver.py.  Plain pointers and sizes only; every array is dense,
 * row-major, float64 unless stated.  All functions return REVS_OK (0) or an error
 * code; revs_last_error() gives the text.  There is NO CPU fallback: every entry point
 * fails with REVS_ERR_CUDA when no sm_100 device is usable.
 *
 * Reference interface replaced by each entry point (file:line in /root/reference):
 *
 *   revs_set_sensitivity / revs_set_feeder_tree(s) lpsolver.py:17-26  compute_Rmat()
 *                                                 lpsolver.py:183-194 Utility.network()
 *   revs_set_homes / revs_set_tariff              lpsolver.py:45-58   Home.__init__ inputs
 *                                                 extract.py:91-132   get_homes_ev_param()
 *   revs_solve_admm                               lpsolver.py:242-290 solve_ADMM()
 *   revs_admm_begin / revs_admm_step              lpsolver.py:254-287 one while-iteration
 *   revs_home_step                                lpsolver.py:44-160  Home(...).solve()
 *   revs_utility_step                             lpsolver.py:163-238 Utility(...).solve()
 *   revs_solve_individual                         lpsolver.py:430-460 solve_residence()
 *   revs_reliability                              drawing.py:29-78    compute_flows(),
 *                                                                     compute_voltage()
 *   revs_get_results / revs_get_schedule(_ld)     lpsolver.py:289-290 return diff,P_sch,S,C
 *   revs_set_option / revs_get_stats / revs_version / revs_last_error / revs_device_count /
 *   revs_reliability_sharded / revs_gather_export / revs_gather_attach
 *                                                 drawing.py:29-78 for one feeder over several GPUs (rows partitioned)
 *   revs_comm_export / revs_comm_attach / revs_comm_detach / revs_zone_arrays
 *                                                 no reference counterpart (library plumbing; the reference is one process)
 */
#ifndef REVS_ADMM_H
#define REVS_ADMM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REVS_OK 0
#define REVS_ERR_ARG 1          /* bad argument / call order                          */
#define REVS_ERR_CUDA 2         /* CUDA error or no usable device                     */
#define REVS_ERR_INFEASIBLE 3   /* a home sub-problem has no solution (lpsolver.py:154)*/
#define REVS_ERR_NOCONV 4       /* utility QP hit its iteration / working-set limit    */

#define REVS_REL_VOLTAGE 0      /* out = sqrt(vset^2 - S_v P)   (drawing.py:75)        */
#define REVS_REL_FLOW 1         /* out = scale[row] * (S_f P)   (drawing.py:56-58)     */
#define REVS_REL_DROP 2         /* out = S_v P                  (R@P of lpsolver.py:192)*/

typedef struct revs_solver revs_solver;

/* counters of the last revs_solve_* call (what ran on the device) */
typedef struct revs_stats {
    int64_t kernel_launches;      /* kernels of this library launched                  */
    int64_t gemm_launches;        /* ... of which sensitivity contractions             */
    int64_t gemm_full_launches;   /* ... of which over all columns (first round of an iteration) */
    int64_t qp_outer_iterations;  /* utility working-set rounds, summed over ADMM iters */
    int64_t qp_newton_iterations; /* restricted Newton steps, summed over columns       */
    int32_t admm_iterations;
    int32_t max_working_set;      /* largest per-(feeder,hour) working set seen         */
    double primal_residual;       /* ||P_est-P_sch||_F / sqrt(H T), last iteration      */
    double dual_residual;         /* kappa ||P_sch-P_sch_prev||_F / sqrt(H T)           */
    double qp_flops;              /* algorithmic FP64 flops of the utility QP kernels   */
    float gemm_ms;                /* device time in sensitivity contractions (events)   */
    float gemm_full_ms;           /* ... of which the launches over ALL (feeder,hour) columns */
    float home_ms;                /* device time in the batched home solve              */
    float dual_ms;                /* device time in the fused dual/residual kernel      */
    float qp_ms;                  /* device time in the per-column QP kernels (both)    */
    float qp_big_ms;              /* ... of which the |W|>32 instantiation              */
    float total_ms;               /* device time of the whole solve                     */
    float qp_warp_ms;             /* ... of qp_ms: warp-per-column kernels (zones <= 256) */
    float qp_init_ms;             /* ... of qp_ms: start-of-solve kernel                 */
    int64_t qp_columns;           /* (zone,hour) columns that entered a QP kernel, summed over rounds */
    int64_t qp_warp_rounds;       /* working-set rounds in which the warp-per-column kernels ran */
} revs_stats;

const char* revs_last_error(void);
int revs_version(void);
int revs_device_count(int* count);

/* A solver owns the device state of a batch of feeders that live on ONE GPU.
 * feeder_off[n_feeders+1]: residences of feeder f are homes feeder_off[f]..feeder_off[f+1]-1
 * of every [H, *] array below (H = feeder_off[n_feeders]).  T = horizon length. */
int revs_create(revs_solver** out, int device, int n_feeders, const int64_t* feeder_off, int T);
int revs_destroy(revs_solver* s);

/* Residence-by-residence voltage sensitivity block R_res (n_f x n_f, symmetric, >=0)
 * of one feeder, from host memory. */
int revs_set_sensitivity(revs_solver* s, int feeder, const double* R_res);

/* Same block built ON THE DEVICE from the radial feeder itself: n_nodes non-substation
 * nodes in topological order (parent[i] < i, -1 = substation), r[i] = resistance of the
 * edge above node i, res_node[n_f] = node index of each residence.  Also enables
 * revs_reliability() on arbitrary nodes / edges of this feeder. */
int revs_set_feeder_tree(revs_solver* s, int feeder, int n_nodes, const int32_t* parent,
                         const double* r, const int32_t* res_node);

/* All feeders of the solver at once (one upload, one kernel): node_off[n_feeders+1] offsets
 * into the concatenated parent / r arrays (parent indices are local to their feeder),
 * res_node [H] in the home order of revs_create (node indices local to the feeder). */
int revs_set_feeder_trees(revs_solver* s, const int64_t* node_off, const int32_t* parent,
                          const double* r, const int32_t* res_node);

/* Per-home inputs (host).  load [H,T] kW; has_ev [H]; rating kW, capacity kWh, initial
 * SOC, start/end = plug-in window [start,end) in steps.  EV arrays are ignored where
 * has_ev==0. */
int revs_set_homes(revs_solver* s, const double* load, const uint8_t* has_ev,
                   const double* rating, const double* capacity, const double* initial,
                   const int32_t* start, const int32_t* end);
int revs_set_tariff(revs_solver* s, const double* cost /* [T] */);

/* Full ADMM run == solve_ADMM(): iter_max iterations from P_est=P_sch=Gamma=0.
 * tol<=0 reproduces the reference (always iter_max iterations); tol>0 stops early when
 * both residuals of revs_stats fall below tol.  iters_done may be NULL. */
int revs_solve_admm(revs_solver* s, double kappa, int iter_max, double vset, double vlow,
                    double vhigh, double tol, int* iters_done);

/* The same loop one iteration at a time (multi-GPU drivers all-reduce the residual
 * sums between steps).  sums[3] = {sum (P_est-P_sch)^2, sum (P_sch-P_sch_prev)^2, H*T}
 * over this solver's homes. */
int revs_admm_begin(revs_solver* s, double kappa, int iter_max, double vset, double vlow,
                    double vhigh);
int revs_admm_step(revs_solver* s, double sums[3]);

/* The two sub-problems on their own, host in/out, all homes of the solver at once.
 * revs_home_step  == Home(cost, homedata, p_est, p_sch, gamma, kappa).solve() per home:
 *   P_sch_new = g_opt [H,T], P_ev = p_opt [H,T] (either may be NULL).
 * revs_utility_step == Utility(graph, P_est, P_sch, Gamma, kappa, vset, vlow, vhigh).solve():
 *   P_est_new = g_opt [H,T]; lam0 (may be NULL) warm-starts the voltage-row multipliers,
 *   lam_out (may be NULL) returns them, both [H,T]. */
int revs_home_step(revs_solver* s, double kappa, const double* p_est, const double* p_sch,
                   const double* gamma, double* P_sch_new, double* P_ev);
int revs_utility_step(revs_solver* s, double kappa, double vset, double vlow, double vhigh,
                      const double* p_est, const double* p_sch, const double* gamma,
                      const double* lam0, double* P_est_new, double* lam_out);

/* Results of the last ADMM run, to host.  Any pointer may be NULL.
 * P_sch [H,T], P_ev [H,T], SOC [H,T+1], diff [diff_rows,H] (lpsolver.py:284): one row per iteration
 * that ran; REVS_ERR_ARG when diff_rows is smaller than that (nothing is written past the buffer). */
int revs_get_results(const revs_solver* s, double* P_sch, double* P_ev, double* SOC,
                     double* diff, int diff_rows);
/* The same results in compact form: the schedule P_sch [H,T] and the charging decisions as bit masks,
 * hour_mask [H, mask_words] with mask_words = ceil(T/64), bit (t % 64) of word t/64 set when the charger of the
 * home runs in step t.  S = rating * bit and the SOC recursion C (lpsolver.py:105-108) follow from the mask and
 * the per-home inputs the caller already holds, so a third of the bytes of revs_get_results cross PCIe. */
int revs_get_schedule(const revs_solver* s, double* P_sch, uint64_t* hour_mask, int mask_words,
                      double* diff, int diff_rows);
/* revs_get_schedule with a row stride for diff: row k of the convergence values goes to diff + k * diff_ld
 * (diff_ld >= H, in doubles), so that several solvers sharing one GPU write their column blocks of one
 * [iterations, all homes] array (lpsolver.py:284 keeps one such array) straight from the device, without a host copy. */
int revs_get_schedule_ld(const revs_solver* s, double* P_sch, uint64_t* hour_mask, int mask_words,
                         double* diff, int diff_rows, int64_t diff_ld);

/* Utility-side iterates of the last run: P_est [H,T], Gamma [H,T]. */
int revs_get_estimate(const revs_solver* s, double* P_est, double* Gamma);

/* Individual optimum of every home == solve_residence() per home. */
int revs_solve_individual(revs_solver* s, double* P_res, double* P_ev, double* SOC);

/* LinDistFlow reliability check of a schedule on one feeder (needs set_feeder_tree):
 * rows = node indices (voltage/drop) or edge indices (= child-node index of the edge,
 * flow).  P [n_f,T] host schedule of the feeder's residences, or NULL to use the last
 * ADMM result.  out [n_rows,T] host.  scale [n_rows] only for REVS_REL_FLOW (signed
 * 1/rating), may be NULL (=1). */
int revs_reliability(revs_solver* s, int feeder, int kind, int n_rows, const int32_t* rows,
                     const double* scale, double vset, const double* P, double* out);

/* The same check with the ROWS PARTITIONED over the GPUs of one box (one process per GPU, every process holds the
 * feeder's tree; BASELINE north_star: "the sensitivity contraction is row-partitioned"): every rank builds and
 * contracts only its block of the requested rows, and the epilogue of the contraction kernel stores each output
 * element straight into the gather buffer of EVERY rank through peer-mapped (NVLink) memory -- the all-gather is
 * fused into the FP64 tensor-core kernel, there is no NCCL call and no extra copy; arrival flags (system-scope
 * release / acquire) close the exchange.  All ranks call it with the same arguments; P (the feeder's whole
 * schedule, [n_f,T]) is required; out [n_rows,T] is complete on every rank.
 * Protocol: revs_gather_export (allocates this rank's buffer for up to capacity_doubles = n_rows*T outputs, returns
 * its 64-byte CUDA IPC handle), exchange the handles on the host, revs_gather_attach with all `world` handles in
 * rank order (<= 16 ranks), a host barrier, then any number of revs_reliability_sharded calls (collective).
 * Replaces drawing.py:29-78 for a feeder too large for one GPU's share of the time; the reference is one process. */
int revs_gather_export(revs_solver* s, int64_t capacity_doubles, void* handle64);
int revs_gather_attach(revs_solver* s, int world, int rank, const void* handles);
int revs_reliability_sharded(revs_solver* s, int feeder, int kind, int n_rows, const int32_t* rows,
                             const double* scale, double vset, const double* P, double* out);

/* Plain sensitivity contraction C[M,T] = A[M,K] @ B[K,T] on the tensor cores (FP64
 * DMMA), host in/out -- exposed so that the GEMM kernel can be tested on its own. */
int revs_contract(int device, int M, int K, int T, const double* A, const double* B, double* C);

/* The in-loop screening contraction on its own: C ~ A @ B with BF16 operands and FP32
 * accumulation (relative error <= 0.5 % for non-negative data).  impl 0 = mma.sync kernel,
 * impl 1 = tcgen05 / TMEM / TMA kernel (T <= 96). */
int revs_screen_contract(int device, int M, int K, int T, const double* A, const double* B, double* C,
                         int impl);

/* Host-only helper (no GPU needed): the static per-zone arrays the tree-structured operator kernel derives from a
 * radial zone (csrc/tree_qp.cu) -- with the residences in depth-first order perm[], R[i][j] = min(c[i..j-1]) and
 * R g is three prefix sums over the Cartesian tree of c (nodes as lo | hi << 16, sorted by lo and by hi, weights w,
 * cnt = #nodes with lo <= p | (#nodes with hi < p) << 16).  Every output has n_res entries.  Exposed so that the
 * decomposition can be checked against compute_Rmat (lpsolver.py:17-26) without a device. */
int revs_zone_arrays(int n_nodes, const int32_t* parent, const double* r, int n_res, const int32_t* res_node, int32_t* perm,
                     double* c, double* d, double* e, int32_t* node_lo, double* w_lo, int32_t* node_hi, double* w_hi,
                     int32_t* cnt);

/* Global stopping rule over the GPUs of one box (one process per GPU, each with its own revs_solver over its
 * share of the feeders): the residual sums of revs_stats / revs_admm_step / the tol test of revs_solve_admm then
 * run over ALL ranks.  The all-reduce happens inside the fused dual-update kernel, through mailboxes in peer
 * (NVLink) memory -- no NCCL call, no extra launch, nothing returns to the host.  Protocol: every rank calls
 * revs_comm_export (a 64-byte CUDA IPC handle of its mailbox), the host exchanges the handles (e.g.
 * torch.distributed.all_gather), every rank calls revs_comm_attach with all `world` handles in rank order
 * (<= 16 ranks).  The mailbox is zeroed by attach: synchronise the ranks (a barrier) between attach and the
 * first revs_admm_begin.  All ranks must then make the same sequence of revs_admm_begin / step / solve calls.
 * No reference counterpart (the reference is a single process). */
int revs_comm_export(revs_solver* s, void* handle64);
int revs_comm_attach(revs_solver* s, int world, int rank, const void* handles);
int revs_comm_detach(revs_solver* s);

/* Options: "tree" (default 0; set it BEFORE revs_set_feeder_tree(s)) = zones given as trees (up to 320 residences) are
 * solved by the tree-structured kernel, which needs no sensitivity matrix (rows and products of R from O(n) static
 * arrays); 0 = dense kernels (BF16 tensor-core screening + FP64 rows of R), the faster of the two on B200.  "graph" (default 1) = revs_solve_admm runs the whole loop from one captured CUDA graph whose loops
 * (ADMM iterations, working-set rounds) are decided on the device; 0 = host-driven loop with CUDA-event spans per
 * kernel family in revs_stats (profiling).  "screen" (default 1) = BF16 tensor-core screening of the voltage rows with exact
 * FP64 recheck of the candidates inside the loop; 0 = FP64 DMMA contraction of every row.
 * "screen_impl" = 0 mma.sync screening kernel, 1 tcgen05/TMEM/TMA screening kernel. */
int revs_set_option(revs_solver* s, const char* name, double value);

int revs_get_stats(const revs_solver* s, revs_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* REVS_ADMM_H */
