/*
 * nsx_b200.h - C ABI of the B200-native network-simplex engine (libnsx_b200.so).
 *
 * The reference (jeffreyhorn/network_flow_solver) is pure Python and defines no FFI;
 * its boundary for this path is the Python call
 *     solve_min_cost_flow(problem, options, ...) -> FlowResult      (src/network_solver/solver.py:13-104)
 * which constructs NetworkSimplex(problem, options) and calls .solve()
 *                                                                   (src/network_solver/simplex.py:99-265, 1446-1765).
 * The entry points below are what a ctypes binding placed inside that function would call:
 * the canonical structure-of-arrays problem in the reference's internal index space goes in,
 * raw flows / potentials / tree flags / the entering-arc trace come out, and the Python layer does
 * the dict building and rounding of simplex.py:1703-1765.  See INTEGRATION.md for the stub.
 *
 * Conventions: plain pointers and sizes, caller-owned HOST buffers (the *_dev variants take
 * caller-owned DEVICE buffers), no pointer is kept after return, return value 0 = call completed
 * (see result->status for the solver outcome), negative = error (nsx_last_error() has the text).
 * Thread-safe: every call builds its own device state; nsx_last_error() is thread-local.
 */
#ifndef NSX_B200_H
#define NSX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSX_ABI_VERSION 4 /* 4: nsx_result.star_* / blk_rebuilds
                             2: nsx_options.node_mask, NSX_SPECIAL_*, NSX_PRICING_DEVEX_LOOP, nsx_solve_warm, NSX_ARC_STALE
                             3: nsx_options.spin_timeout_ms, nsx_options.flags reserved (NSX_FLAG_FAST_POTENTIALS removed),
                                nsx_mailbox_abort, nsx_result.fault */

/* pricing rules: SolverOptions.pricing_strategy (src/network_solver/data.py:459-488) */
#define NSX_PRICING_DANTZIG 0 /* DantzigPricing.select_entering_arc, simplex_pricing.py:97-137 */
#define NSX_PRICING_DEVEX 1   /* DevexPricing._select_entering_arc_vectorized, simplex_pricing.py:310-357 + simplex.py:528-617 */
#define NSX_PRICING_CANDIDATE_LIST 2 /* CandidateListPricing (simplex_pricing.py:375-542), which is also what the
                                        reference's default "adaptive" strategy (simplex_pricing.py:545-639) runs in
                                        practice: candidate scan in the pivot CTA, top-100 refresh as a grid-wide sweep */

#define NSX_PRICING_DEVEX_LOOP 3 /* DevexPricing.select_entering_arc with use_vectorized_pricing=False (simplex_pricing.py:
                                    205-269): same blocks, weights and reset cadence as NSX_PRICING_DEVEX, but a sequential
                                    scan - tree cost of the current phase, a later arc wins only when its merit is larger by
                                    more than the tolerance, first zero-reduced-cost arc of the block, the block pointer
                                    advances after a zero pick, no exclusion of the last degenerate arc */

/* structure-specific entering rules tried before `pricing` on every iteration; when one finds nothing the configured
 * rule runs (NetworkSimplex._find_entering_arc, simplex.py:1058-1075; select_pivot_strategy,
 * specialized_pivots.py:452-527).  Values of nsx_options.row_scan_first. */
#define NSX_SPECIAL_NONE 0
#define NSX_SPECIAL_ROW_SCAN 1      /* transportation: Dantzig-rule row scan, specialized_pivots.py:80-120 */
#define NSX_SPECIAL_ASSIGNMENT 2    /* forward arcs only, a candidate replaces the incumbent when its reduced cost is
                                       lower by MORE than the tolerance (sequential scan), specialized_pivots.py:179-209 */
#define NSX_SPECIAL_MAX_FLOW 3      /* largest residual * |reduced cost|, first arc on ties, specialized_pivots.py:294-343 */
#define NSX_SPECIAL_SHORTEST_PATH 4 /* the assignment scan restricted to forward arcs whose tail is reachable from the
                                       source (nsx_options.node_mask) plus backward arcs, specialized_pivots.py:368-424 */

/* solver outcome (FlowResult.status, data.py:269-322; UnboundedProblemError, simplex.py:1231-1246) */
#define NSX_STATUS_OPTIMAL 0
#define NSX_STATUS_INFEASIBLE 1          /* artificial flow left after Phase 1, simplex.py:1600-1624 */
#define NSX_STATUS_ITERATION_LIMIT 2     /* limit hit in Phase 2 (flows are feasible), simplex.py:1678-1699 */
#define NSX_STATUS_UNBOUNDED 3           /* theta = +inf in the ratio test */
#define NSX_STATUS_ITERATION_LIMIT_P1 4  /* limit hit in Phase 1 with artificial flow left, simplex.py:1602-1613 */

/* error codes (negative return values) */
#define NSX_ERR_INVALID_ARGUMENT (-1)
#define NSX_ERR_CUDA (-2)
#define NSX_ERR_NO_DEVICE (-3)
#define NSX_ERR_INTERNAL (-4)

/* per-arc state byte returned in nsx_result.state */
#define NSX_ARC_IN_TREE 1u
#define NSX_ARC_CAN_FWD 2u  /* upper - flow > tol */
#define NSX_ARC_CAN_BWD 4u  /* flow > tol */
#define NSX_ARC_STALE 16u  /* warm start only: the arc was put into the initial tree and has not been on a pivot
                               cycle since; the reference's ratio test still sees its cold-start flow there
                               (forward/backward_residuals are refreshed per cycle arc only, simplex.py:1276-1283, and
                               never after _apply_warm_start_basis) and the engine reproduces that */
#define NSX_ARC_TOUCHED 8u  /* flow was written by a pivot with theta > 0 (drives NumPy-vs-Python rounding
                               of the result, simplex.py:1703-1721) */

/*
 * Canonical problem, reference index space (simplex.py:149-163, 392-432, 1431-1440):
 *   node 0 = artificial root, nodes 1..n_nodes-1 = problem nodes in sorted-id order;
 *   real arcs 0..n_arcs-1 in stable (tail id, head id) order, lower bounds already shifted out;
 *   artificial arc n_arcs + (v-1) joins node v and the root (simplex.py:645-698) - built by the engine.
 */
typedef struct nsx_problem {
    int32_t n_nodes;         /* including the root */
    int64_t n_arcs;          /* real arcs M */
    const int32_t* tail;     /* [M] */
    const int32_t* head;     /* [M] */
    const double* pert_cost; /* [M] perturbed Phase-2 cost  c_i + 1e-10 * 1.00001^i */
    const double* upper;     /* [M] capacity minus lower bound, +inf = uncapacitated */
    const double* supply;    /* [n_nodes] after the lower-bound shift; supply[0] ignored */
    double penalty;          /* artificial-arc cost  max|c| * (n_nodes + 1) */
} nsx_problem;

typedef struct nsx_options {
    int32_t pricing;          /* NSX_PRICING_* */
    int32_t row_scan_first;   /* NSX_SPECIAL_*: 1 = transportation row-scan rule tried first (specialized_pivots.py:80-120,
                                 simplex.py:1060-1064), 2..4 the other structure rules; all fall through to `pricing`
                                 when they find nothing */
    int64_t block_size;       /* initial Devex block size (simplex_adaptive.py:70-96 when auto) */
    int32_t auto_block;       /* 1: x1.5 / x0.75 adaptation every 50 pivots (simplex_adaptive.py:98-151) */
    int32_t ft_update_limit;  /* Devex weights and pricing block reset on every (limit+1)-th tree change
                                 (simplex.py:1373-1400) */
    int64_t max_iterations;   /* > 0 */
    double tolerance;         /* SolverOptions.tolerance */
    int64_t trace_capacity;   /* entries available in result->entering_trace (0 = no trace) */
    int32_t device;           /* CUDA device ordinal */
    uint32_t flags;           /* reserved, must be 0 (the engine has one arithmetic mode: the reference's float64 operation
                                 order with the exact top-down potential recompute, SURVEY.md 8/a0 + a8) */
    const uint8_t* node_mask; /* [n_nodes] host buffer, NSX_SPECIAL_SHORTEST_PATH only: 1 = node reachable from the
                                 source over the real arcs (the nodes that carry a distance label in
                                 ShortestPathPivotStrategy, specialized_pivots.py:426-450); NULL otherwise */
    int32_t spin_timeout_ms;  /* deadline of every device-side wait (a sweep worker's answer, a peer GPU's candidate, the
                                 next command): when it passes, all CTAs leave the resident kernel and the call returns
                                 NSX_ERR_INTERNAL instead of hanging the GPU.  0 = default (30 000 ms) */
} nsx_options;

typedef struct nsx_result {
    /* caller-allocated outputs */
    double* flow;            /* [M + n_nodes - 1] real then artificial arcs */
    double* potential;       /* [n_nodes] reference sign convention: rc = c + pi[tail] - pi[head] */
    uint8_t* state;          /* [M + n_nodes - 1] NSX_ARC_* bits */
    int32_t* entering_trace; /* [trace_capacity] arc * 2 + (direction < 0) per pivot, or NULL */
    /* scalars written by the call */
    int64_t trace_len;
    int64_t iterations;
    int64_t phase1_iterations;
    int64_t degenerate_pivots;    /* theta <= tol, simplex.py:1249-1251 */
    int64_t artificial_with_flow; /* after Phase 1 */
    int64_t tree_updates;         /* pivots that changed the tree */
    int64_t weight_resets;        /* Devex reset cadence hits */
    int64_t final_block_size;
    int64_t arcs_priced;          /* arcs examined by all pricing sweeps */
    int64_t sweeps;               /* pricing sweeps (grid-wide command rounds) */
    int64_t unbounded_arc;        /* entering arc when status == NSX_STATUS_UNBOUNDED */
    double unbounded_rc;
    int32_t status;               /* NSX_STATUS_* */
    int32_t grid_ctas;            /* CTAs of the resident kernel (1 pivot CTA + sweep workers) */
    int32_t bytes_per_arc;        /* bytes a sweep streams per arc in the layout chosen for this instance
                                     (2 * node id + cost + 1 state byte; Devex adds 4) */
    int32_t ring_stages;          /* depth of the shared-memory tile ring of a sweeping CTA */
    int32_t resident_mode;        /* node state held in the pivot CTA's shared memory: 0 none, 1 records +
                                     potentials, 2 everything */
    int32_t store_layout;         /* encoding of the pricing store: node ids (0 int32, 1 uint16) | cost (0 float64,
                                     1 int32, 2 int16) << 8 */
    /* device-side timing (milliseconds unless stated) */
    double solve_ms;              /* CUDA-event time of the resident pivot loop */
    double h2d_ms, d2h_ms;        /* host<->device copies (host-buffer entry points only) */
    double pricing_ms;            /* accumulated device clock in pricing sweeps (max over CTAs' leader) */
    double pivot_ms;              /* accumulated device clock in ratio test + tree/potential update */
    double sync_ms;               /* accumulated device clock in grid-wide handshakes */
    double exchange_ms;           /* accumulated device clock in the cross-GPU candidate exchange (sharded solves) */
    int64_t sum_cycle_len;        /* pivot statistics for DESIGN.md / profiles */
    int64_t sum_subtree;
    int64_t max_subtree;
    int64_t sum_rounds;           /* wavefront rounds of the exact potential recompute */
    int64_t sum_window;           /* preorder-array entries moved by all tree updates */
    int64_t phase_cycles[12];     /* SM-clock cycles spent per pivot phase (walk, residuals, ratio test, flow
                                     update, bookkeeping, stem snapshot, window permutation, copy-back + stem,
                                     potentials, reset cadence, spare, spare) */
    int64_t handshake_ns[8];      /* grid handshake timeline, ns after the command is published, summed over all
                                     sweeps: worker 1 saw the command, entered the sweep, potentials staged,
                                     tiles done, reduced, arrived; pivot CTA saw all arrivals, merged */
    int32_t fault;                /* 0, or why the resident kernel gave up (the call then returns NSX_ERR_INTERNAL):
                                     1 a sweep worker did not answer, 2 a peer GPU did not deliver its candidate,
                                     3 a peer GPU raised its abort word, 4 a worker saw no command, 5 bad node id,
                                     6 a sweep worker did not reach the barrier of a star-pricing update */
    int32_t star_pricing;         /* 1: the Dantzig rule was priced from the row cache (star pricing) */
    int64_t star_updates;         /* pricing steps that brought the row cache up to date after a pivot */
    int64_t star_builds;          /* pricing steps that priced every row afresh (start, phase switch, large subtrees) */
    int64_t star_rescans;         /* rows priced afresh because their cached arc got worse */
    int64_t blk_rebuilds;         /* blocked preorder array: fresh layouts (trees in HBM) */
} nsx_result;

/* Solve one instance on one GPU; all nsx_problem / nsx_result pointers are HOST memory. */
int nsx_solve(const nsx_problem* problem, const nsx_options* options, nsx_result* result);

/* Same, but tail/head/pert_cost/upper are DEVICE pointers already resident in HBM (supply stays host,
 * it is n_nodes doubles).  Outputs are still host buffers. Used for the kernel-only throughput figure. */
int nsx_solve_resident(const nsx_problem* problem_dev, const nsx_options* options, nsx_result* result);

/* Measurement aid: `sweeps` full pricing sweeps over the initial state of the instance (device arrays as in
 * nsx_solve_resident) through the command / arrival protocol of a solve, without pivoting.  result->solve_ms,
 * arcs_priced and pricing_ms describe the sweeps; flows / potentials are the initial ones. */
int nsx_sweep_probe(const nsx_problem* problem_dev, const nsx_options* options, int32_t sweeps,
                    nsx_result* result);

/*
 * Arc-sharded pricing over several GPUs of one node (one process per GPU).  Every rank passes the SAME full
 * instance; a pricing sweep is split over the sweep CTAs of all ranks, the per-rank candidates are exchanged
 * through peer-mapped mailboxes over NVLink (one 64-byte record per rank and sweep) and every rank applies the
 * identical pivot, so all ranks return identical results.  No reference counterpart (the reference is single
 * process); the pricing rule, tie-break and results are those of nsx_solve.
 *   mailboxes[r] = mailbox of rank r as a device pointer valid on THIS rank's GPU: the rank's own mailbox from
 *   nsx_mailbox_create, the others from nsx_mailbox_open on the 64-byte handles exchanged by the host layer.
 *   Every rank must nsx_mailbox_reset its own mailbox and pass a host barrier before each sharded call.
 */
typedef struct nsx_shard {
    int32_t rank, world;      /* world <= 8 */
    void* const* mailboxes;   /* [world] */
} nsx_shard;
int nsx_solve_sharded(const nsx_problem* problem, const nsx_options* options, nsx_result* result,
                      const nsx_shard* shard);
/* Same with tail / head / pert_cost / upper already resident in THIS rank's HBM (device pointers, as in
 * nsx_solve_resident): the ranks then start their kernels within microseconds of the host barrier instead of after
 * differently long uploads. */
int nsx_solve_sharded_resident(const nsx_problem* problem_dev, const nsx_options* options, nsx_result* result,
                               const nsx_shard* shard);
int nsx_sweep_probe_sharded(const nsx_problem* problem_dev, const nsx_options* options, int32_t sweeps,
                            nsx_result* result, const nsx_shard* shard);
int64_t nsx_mailbox_bytes(void);
int nsx_mailbox_create(int32_t device, void** mailbox, unsigned char handle[64]);
int nsx_mailbox_open(int32_t device, const unsigned char handle[64], void** mailbox);
int nsx_mailbox_reset(int32_t device, void* mailbox);
int nsx_mailbox_close(int32_t device, void* mailbox, int32_t is_local);
/* Raise the abort word of a mailbox (own or peer-mapped): the kernel polling that mailbox leaves with fault 3.  The host
 * layer calls it on every peer when its own rank fails between the barrier and the launch, so the others do not wait out
 * the full deadline. */
int nsx_mailbox_abort(int32_t device, void* mailbox);

/*
 * Warm start (NetworkSimplex._apply_warm_start_basis / _recompute_tree_flows, simplex.py:740-1021; phase choice
 * simplex.py:1494-1530).  The host layer turns the reference's Basis(tree_arcs, arc_flows) into the initial state
 * below - which arcs (real and artificial) form the spanning tree, every arc's flow, and whether Phase 1 is
 * skipped - and the engine starts its resident pivot loop from that tree instead of the all-artificial star.
 * in_tree must mark exactly n_nodes - 1 arcs that span all nodes (artificial arc M + v - 1 joins node v and the
 * root); anything else is NSX_ERR_INVALID_ARGUMENT.  Host buffers, nsx_solve's conventions otherwise.
 */
typedef struct nsx_warm_start {
    const uint8_t* in_tree; /* [M + n_nodes - 1] 1 = arc belongs to the initial tree */
    const double* flow;     /* [M + n_nodes - 1] initial flows (0 on arcs outside the tree) */
    int32_t start_phase;    /* 1 = Phase 1 first; 2 = Phase 1 skipped, accepted only when no artificial arc is marked
                               (simplex.py:1504-1513). A tree that spans the root always contains an artificial arc - only
                               those touch the root - so 2 is rejected for every valid tree, as the reference's own
                               skip never triggers; the host layer always passes 1 */
} nsx_warm_start;
int nsx_solve_warm(const nsx_problem* problem, const nsx_options* options, const nsx_warm_start* warm,
                   nsx_result* result);

/* Solve `count` independent instances on one GPU, one CTA per instance (batched config).
 * problems[i] / results[i] as in nsx_solve; options are shared. */
int nsx_solve_batch(int64_t count, const nsx_problem* problems, const nsx_options* options,
                    nsx_result* results);

const char* nsx_last_error(void);
void nsx_version(int32_t* abi, int32_t* sm_arch);
int nsx_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NSX_B200_H */
