// nsx_core.cuh - pivot logic of the B200 network-simplex engine (ratio test, flow update,
// spanning-tree re-hang on a preorder array, exact node-potential recompute).
//
// The code is written as CTA-wide phases: NSX_PAR_FOR loops whose iterations are independent,
// separated by NSX_SYNC().  On the device a phase is executed by all threads of one CTA; with
// NSX_HOST_EMU defined the same source compiles with g++ as a serial program (one "thread"),
// which tests/ use to check the device logic on machines without a GPU.  The emulation is test
// infrastructure: the product library is built by nvcc only and has no host execution path.
//
// Reference behaviour restated here (paths relative to the reference root):
//   ratio test / flow update / bound flip ........ src/network_solver/simplex.py:1176-1334
//   cycle orientation ............................ src/network_solver/basis.py:178-241
//   potentials = fold of +-cost along root path .. src/network_solver/basis.py:82-125
//   Devex weight = tree-path length .............. simplex_pricing.py:271-292
//   reset cadence (weights := 1, block := 0) ..... simplex.py:1373-1425
//   block-size adaptation ........................ simplex_adaptive.py:98-151
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/nsx_b200.h"

#if defined(__CUDACC__)
#define NSX_HD __host__ __device__
#else
#define NSX_HD
#endif

#if defined(__CUDACC__) && !defined(NSX_HOST_EMU)
#define NSX_ON_DEVICE 1
#define NSX_FN __device__ __forceinline__
#define NSX_FN_COLD __device__ __noinline__  /* kept out of the hot paths' register allocation */
#define NSX_PAR_FOR(i, lo, hi) \
    for (int64_t i = (int64_t)(lo) + (int64_t)threadIdx.x; i < (int64_t)(hi); i += (int64_t)blockDim.x)
// __syncwarp() first: bar.sync is an *aligned* barrier and nvcc may thread a preceding
// thread-0-only block into code that reaches the barrier un-converged (observed on sm_100a:
// warp 0 arrived twice and the CTA slipped one barrier phase).
#define NSX_SYNC()        \
    do {                  \
        __syncwarp();     \
        __syncthreads();  \
    } while (0)
#define NSX_SINGLE if (threadIdx.x == 0)
#define NSX_TID ((int)threadIdx.x)
#define NSX_NTHREADS ((int)blockDim.x)
// IEEE float64 without fused multiply-add: the reference's operation order
#define NSX_ADD(a, b) __dadd_rn((a), (b))
#define NSX_SUB(a, b) __dsub_rn((a), (b))
#define NSX_MUL(a, b) __dmul_rn((a), (b))
#define NSX_DIV(a, b) __ddiv_rn((a), (b))
#define NSX_INF __longlong_as_double(0x7ff0000000000000LL)
#define NSX_ATOMIC_ADD_I32(p, v) atomicAdd((p), (v))
#define NSX_ATOMIC_ADD_F64(p, v) atomicAdd((p), (v))
#define NSX_CLOCK() clock64()
#define NSX_ATOMIC_MIN_I32(p, v) atomicMin((p), (v))
#define NSX_ATOMIC_MAX_I32(p, v) atomicMax((p), (v))
#elif defined(NSX_HOST_MT)
// Test-only: the CTA emulated by real host threads (tests/emu, NSX_HOST_MT).  NSX_SYNC is a thread barrier, so a barrier
// that some "thread" skips deadlocks and a missing barrier shows up as a data race (ThreadSanitizer) - the two hazards
// the serial emulation cannot see.
extern thread_local int nsx_mt_tid;
extern int nsx_mt_nthreads;
void nsx_mt_barrier();
#define NSX_ON_DEVICE 0
#define NSX_FN static inline
#define NSX_FN_COLD static inline
#define NSX_PAR_FOR(i, lo, hi) for (int64_t i = (int64_t)(lo) + nsx_mt_tid; i < (int64_t)(hi); i += nsx_mt_nthreads)
#define NSX_SYNC() nsx_mt_barrier()
#define NSX_SINGLE if (nsx_mt_tid == 0)
#define NSX_TID nsx_mt_tid
#define NSX_NTHREADS nsx_mt_nthreads
#define NSX_HOST_SERIAL if (nsx_mt_tid == 0)  /* host-only stand-ins for warp-level device code run on one thread */
#else
#define NSX_ON_DEVICE 0
#define NSX_FN static inline
#define NSX_FN_COLD static inline
#define NSX_PAR_FOR(i, lo, hi) for (int64_t i = (int64_t)(lo); i < (int64_t)(hi); ++i)
#define NSX_SYNC() ((void)0)
#define NSX_SINGLE
#define NSX_TID 0
#define NSX_NTHREADS 1
#define NSX_HOST_SERIAL
#endif
// several threads may raise the same flag at once (same value, nobody reads it before the next barrier)
#if defined(NSX_HOST_MT) && !(defined(__CUDACC__) && !defined(NSX_HOST_EMU))
#define NSX_RAISE_TO(flag, v) __atomic_store_n(&(flag), (v), __ATOMIC_RELAXED)
#else
#define NSX_RAISE_TO(flag, v) ((flag) = (v))
#endif
#if !NSX_ON_DEVICE
#define NSX_ADD(a, b) ((a) + (b)) /* compiled with -ffp-contract=off */
#define NSX_SUB(a, b) ((a) - (b))
#define NSX_MUL(a, b) ((a) * (b))
#define NSX_DIV(a, b) ((a) / (b))
#define NSX_INF INFINITY
#define NSX_CLOCK() 0ll
#if defined(NSX_HOST_MT)
#define NSX_ATOMIC_ADD_I32(p, v) __atomic_fetch_add((p), (v), __ATOMIC_RELAXED)
static inline void nsx_mt_addf(double* p, double v) {
    uint64_t o = __atomic_load_n((uint64_t*)p, __ATOMIC_RELAXED), n;
    double t;
    do { memcpy(&t, &o, 8); t += v; memcpy(&n, &t, 8); } while (!__atomic_compare_exchange_n((uint64_t*)p, &o, n, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}
#define NSX_ATOMIC_ADD_F64(p, v) nsx_mt_addf((p), (v))
static inline int32_t nsx_mt_min(int32_t* p, int32_t v) { int32_t o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v < o && !__atomic_compare_exchange_n(p, &o, v, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
static inline int32_t nsx_mt_max(int32_t* p, int32_t v) { int32_t o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v > o && !__atomic_compare_exchange_n(p, &o, v, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
#define NSX_ATOMIC_MIN_I32(p, v) nsx_mt_min((p), (v))
#define NSX_ATOMIC_MAX_I32(p, v) nsx_mt_max((p), (v))
#else
#define NSX_ATOMIC_ADD_I32(p, v) ((*(p) += (v)) - (v))
#define NSX_ATOMIC_ADD_F64(p, v) (*(p) += (v))
#define NSX_ATOMIC_MIN_I32(p, v) (*(p) = (v) < *(p) ? (v) : *(p))
#define NSX_ATOMIC_MAX_I32(p, v) (*(p) = (v) > *(p) ? (v) : *(p))
#endif
#endif

// ------------------------------------------------------------------------------------------
// Device-resident state
// ------------------------------------------------------------------------------------------
// One 16-byte record per node so that a hop of the cycle walk is a single 128-bit load.
struct alignas(16) NsxNode {
    int32_t parent;  // parent node (root: itself)
    int32_t pred2;   // (arc joining the node to its parent) * 2 + (arc points child->parent), -1 for root
    int32_t pos;     // position in the preorder array
    int32_t size;    // subtree size (node included): subtree = order[pos .. pos+size)
};

// Star pricing (Dantzig rule without full sweeps).  The arcs are sorted by tail, so the out-arcs of a node are one
// index range ("row").  rc[v] caches the best improving arc of row v - smallest key, lowest arc*2+dir on ties, exactly
// the order of nsx_dantzig_improving.  A pivot changes the potentials of the re-hung subtree S only, so only arcs with
// an endpoint in S can change their reduced cost: the rows of S are priced afresh, the in-arcs of S (CSC copy of the
// arcs, grouped by head) propose themselves to the rows of their tails, a row whose cached arc got worse is priced
// afresh too, and the entering arc is the minimum over the row cache.  Same arc as a full sweep, ~|S| * degree arcs
// examined instead of m.
struct alignas(16) NsxRC { double key; int32_t arc2; int32_t pad; };  // arc2 < 0: the row has no improving arc; pad = star round in
                                                                     // which the pivot last emptied the row (it is priced afresh then)

// One changed potential, as the sweep workers patch it into their shared-memory copy (NsxDev::pi_delta)
struct alignas(16) NsxPiDelta { int32_t node, pad; double pi; };
#define NSX_PI_DELTA_CAP 2048

struct NsxDev {
    int32_t n;   // nodes incl. root
    int64_t m;   // real arcs
    int64_t ma;  // m + n - 1 (real + artificial)
    // arcs, structure-of-arrays in HBM
    const int32_t* tail;  // [m]   real arcs (caller's arrays, used in place)
    const int32_t* head;  // [m]
    const double* pert;   // [m]   perturbed Phase-2 cost
    const double* upper;  // [m]
    int32_t* atail;       // [n-1] artificial arcs m + v - 1 (engine-owned)
    int32_t* ahead;       // [n-1]
    double* aupper;       // [n-1]
    double* flow;         // [ma]
    uint8_t* state;       // [ma] NSX_ARC_* bits
    uint32_t* wgt;        // [m]  Devex weight: (epoch << 24) | path length; stale epoch => weight 1
    // nodes
    NsxNode* node;     // [n]
    int32_t* depth;    // [n]
    double* pi;        // [n]
    double* pi_mirror; // optional second copy of pi kept in step (global copy read by the sweep
                       // CTAs when the pivot CTA keeps its own node state in shared memory)
    int32_t* order;    // [n] preorder array
    int32_t* tmp;      // [n] scratch for the re-hang permutation
    int32_t* gpath_h;  // [n] cycle path spill (node ids), head side
    int32_t* gpath_t;  // [n] tail side
    int32_t* garc2;    // [2n+1] spill of cycle arcs in scan order
    double* gres;      // [2n+1] spill of residuals in scan order
    double penalty;
    double tol;
    int32_t scan_walk;  // node records are in shared memory: find the cycle by a parallel ancestor scan
    uint16_t* par16;    // optional shared-memory mirror of the parent pointers (parent - 1), for the cycle walk
    uint32_t* root_bits;  // bit v set: the parent of v is the root (the one value parent - 1 cannot encode)
    const uint8_t* node_mask;  // [n] NSX_SPECIAL_SHORTEST_PATH: node reachable from the source (HBM), else null
    double* imbalance;         // [n] warm starts only (else null): flow that clamping to a bound added at / removed from
                               // each node - the reference checks conservation after Phase 1 (simplex.py:1575-1598)
    // star pricing (null / unused when it is off)
    NsxRC* rc;             // [n] row cache; Devex: [2n], forward candidates first (key = -merit), backward candidates from n on
    uint32_t* csc_wgt;     // [m] Devex weights in CSC order (written together with wgt), star pricing under Devex only
    int32_t* dlist;        // [n] nodes whose potential the last pivot changed (its re-hung subtree)
    int32_t* dstamp;       // [n] star round in which the node was last in dlist
    int32_t* dinfo;        // [8 (n + 1)] per listed node {node, first out-arc, out-degree, first CSC entry, in-degree, -, -, -}: what
                           // the sweep workers turn into work items (the row of the entering arc is appended with in-degree 0)
    const int32_t* row_begin;  // [n+1] out-arcs of v = arcs row_begin[v] .. row_begin[v+1]
    const int32_t* col_begin;  // [n+1] in-arcs of v = entries col_begin[v] .. col_begin[v+1] of the CSC copy
    int32_t* csc_pos;      // [m] arc -> its entry in the CSC copy (the state byte is written in both places)
    uint8_t* csc_state;    // [m] state bytes in CSC order
    struct NsxBlk* blk; // trees that live in HBM: the preorder array is kept in blocks with slack (see NsxBlk); `order` is
                        // then the block arena (NSX_BLK_MAX << blk->lg entries) and node.pos a physical index into it
    int32_t* sidx;      // [n] blocked mode: index of a node inside the sequence nsx_recompute_potentials runs over
    // [NSX_PI_DELTA_CAP] grid kernel whose workers keep the potentials in shared memory, else null: the potentials the last
    // pivot changed (its re-hung subtree).  A sweep command carries their number (NsxPivotScratch::pi_delta_n) and the
    // workers patch their copy instead of copying all n potentials again.
    NsxPiDelta* pi_delta = nullptr;
};

#define NSX_CL_SIZE 100    // candidate-list length (simplex.py:232)
#define NSX_CL_REFRESH 10  // major iterations between full refreshes (simplex.py:233)
#define NSX_CL_MINOR 3     // candidate scans per major iteration (simplex_pricing.py:400)
// Solver scalars; lives in global memory, written by the pivot CTA only.
struct NsxCtl {
    int32_t phase;        // 1 or 2
    int32_t status;       // NSX_STATUS_*; -1 while running
    int64_t it;           // pivots in the current phase
    int64_t total;        // pivots in finished phases
    int64_t maxit;
    int64_t bs, pb;       // Devex block size / current block
    int32_t last_deg;     // arc excluded from the next Devex search, -1 none
    int32_t ftc;          // tree changes since the last weight reset
    int32_t ft_limit;
    int32_t auto_block;
    int32_t pricing;
    int32_t row_scan_first;
    int64_t tuner_total, tuner_deg, tuner_last;
    int64_t art_with_flow;
    // candidate-list pricing (simplex_pricing.py:375-542): arcs of the list, refresh / minor-iteration counters
    int32_t cl_count, cl_since, cl_minor;
    int32_t cl_list[NSX_CL_SIZE];
    uint32_t wepoch;      // current Devex weight epoch (8 bits used)
    int32_t need_wfill;   // epoch wrapped: weights must be physically refilled
    int32_t warm;         // started from a caller-supplied tree (nsx_solve_warm): NSX_ARC_STALE bits may be set
    int32_t unbalanced;   // warm start: some node's balance is off by more than tol after Phase 1 (simplex.py:1575-1598)
    // star pricing: on for this solve (1 Dantzig rule, 2 Devex with a single block) / row cache consistent with the state before the pending pivot / pivots since the
    // cache was last brought up to date / current round (stamp) / nodes in dlist / row of the entering arc
    int32_t star_on, star_valid, star_pending, star_round, star_nd, star_extra;
    int32_t star_ne;             // entries of dinfo for the next update (star_nd, + 1 when the row of the entering arc is not among them)
    int32_t star_excl_prev;      // Devex: arc the last star command left out of the cache (DevexPricing.last_degenerate_arc), -1 none
    int64_t star_evaluated;   // arcs examined by the last star command (reported by the sweep workers)
    int64_t star_updates, star_builds, star_rescans;
    int64_t blk_rebuilds;     // blocked preorder array: fresh layouts
    // statistics
    int64_t degenerate, tree_updates, resets, arcs_priced, sweeps;
    int64_t avg_cycle;    // running mean of the cycle length * 16 (chooses the cycle-walk variant)
    int64_t sum_cycle, sum_subtree, max_subtree, sum_rounds, sum_window;
    int64_t phase1_iterations, art_after_p1;
    int64_t trace_len, trace_cap;
    int64_t unbounded_arc;
    double unbounded_rc;
    int64_t clk_pricing, clk_pivot, clk_sync, clk_xchg;
    int64_t ph[12];       // SM-clock cycles per pivot phase (thread 0): walk, residuals, ratio, flow,
                          // bookkeeping, snapshot+sizes, window, copy+stem, potentials, cadence, driver
    int32_t fault;        // a device-side wait ran past its deadline or met an abort word (nsx_result.fault); the loop ends
    int32_t n_special;    // real arcs outside the tree whose residual bits are anything but "forward only" (nsx_special): while
                          // there is none - e.g. for the whole solve of an uncapacitated instance - the Dantzig sweep needs no
                          // state bytes (nsx_price_tile_dz).  Counted once by nsx_count_special, kept up to date by nsx_pivot.
};
// bar.sync blocks lazily (at the next access to barrier-protected state): touch shared memory first
// so that the clock is read after the barrier has really been passed
#define NSX_PH(c, k, t0) NSX_SINGLE { long long t1__ = NSX_CLOCK() + (*(volatile int32_t*)&(c).phase & 0); (c).ph[k] += t1__ - (t0); (t0) = t1__; }

#define NSX_PATH_CAP 512  // cycle entries per side kept in shared memory (longer cycles spill to HBM)

// Scratch of the pivot CTA (shared memory on the device).
struct NsxPivotScratch {
    int32_t nh, nt;            // path lengths (nodes below the join on each side)
    int32_t join;
    int32_t leave_k;           // scan position of the leaving arc
    int32_t leave_arc;
    int32_t spill;             // 1: paths longer than NSX_PATH_CAP, use the global spill arrays
    int32_t art_delta;
    int32_t rounds;
    int32_t pending;
    int32_t p_pos;             // preorder position of the new parent p (tree update)
    int32_t jkey;              // scan walk: min over common ancestors of (size << 16 | node)
    // Tree bookkeeping the next pricing step does not need is deferred (nsx_pivot_flush): it runs while the sweep workers
    // price, or at the latest before the next reader of the preorder array.  def_kind 1: dense array - shift the entries
    // between the old and the new place of S into tmp and copy the window back; 2: blocked array - take S out, put it back.
    int32_t def_kind, def_p, def_sz, def_pad;
    int32_t spec_e0;           // nsx_special() of the entering arc before the pivot touched it
    int32_t pi_delta_n;        // entries of NsxDev::pi_delta that describe ALL potential changes since the last sweep command
                               // (0: none), or -1: the workers must copy all potentials (start, phase switch, two tree
                               // updates without a command in between, a subtree larger than the buffer)
    int64_t def_lo, def_hi, def_S0, def_S1, def_xshift;
    int64_t def_moved;
    int32_t sp_any;            // rule scan: number of the last round in which some thread saw an arc that beats the incumbent
    int32_t sp_best, sp_zero;  // rule scan: incumbent arc*2 + (dir<0) / first zero-reduced-cost candidate, -1 none
    double theta;
    double sp_key;             // rule scan: key of the incumbent
    int32_t path_h[NSX_PATH_CAP];
    int32_t path_t[NSX_PATH_CAP];
    int32_t arc2[2 * NSX_PATH_CAP + 1];  // scan order: arc*2 + (sign<0)
    double res[2 * NSX_PATH_CAP + 1];
    // stem (old pos / size / depth of s_0..s_k), aliases res[] storage after the ratio test
};

NSX_FN double nsx_arc_cost(const NsxDev& d, int32_t phase, int64_t a) {
    // simplex.py:1162-1168: Phase 1 real arc = pert - 1.0 - 1e-6*idx ; artificial = penalty
    if (a >= d.m) return d.penalty;
    double c = d.pert[a];
    if (phase == 1) c = NSX_SUB(NSX_SUB(c, 1.0), NSX_MUL(1e-6, (double)a));
    return c;
}

NSX_FN bool nsx_isinf(double x) { return x == NSX_INF; }
NSX_FN int32_t nsx_tail(const NsxDev& d, int64_t a) { return a < d.m ? d.tail[a] : d.atail[a - d.m]; }
NSX_FN int32_t nsx_head(const NsxDev& d, int64_t a) { return a < d.m ? d.head[a] : d.ahead[a - d.m]; }
NSX_FN double nsx_upper(const NsxDev& d, int64_t a) { return a < d.m ? d.upper[a] : d.aupper[a - d.m]; }

NSX_FN int32_t nsx_parent(const NsxDev& d, int32_t v) {
    if (d.par16) return ((d.root_bits[v >> 5] >> (v & 31)) & 1u) ? 0 : (int32_t)d.par16[v] + 1;
    return d.node[v].parent;
}
NSX_FN void nsx_set_parent_mirror(const NsxDev& d, int32_t v, int32_t parent) {
    // (called by one thread per node; root_bits words are shared between nodes => atomic bit ops)
    if (!d.par16) return;
#if NSX_ON_DEVICE
    if (parent == 0) atomicOr(&d.root_bits[v >> 5], 1u << (v & 31));
    else { atomicAnd(&d.root_bits[v >> 5], ~(1u << (v & 31))); d.par16[v] = (uint16_t)(parent - 1); }
#else
    if (parent == 0) d.root_bits[v >> 5] |= 1u << (v & 31);
    else { d.root_bits[v >> 5] &= ~(1u << (v & 31)); d.par16[v] = (uint16_t)(parent - 1); }
#endif
}

// ------------------------------------------------------------------------------------------
// Blocked preorder array (trees that live in HBM / L2: n beyond what the pivot CTA can hold on-chip).
// A dense preorder array makes re-hanging a subtree S cost a shift of every entry between its old and its new
// place - ~n/3 entries per pivot, 92 us at n = 2^20 even when S is a single node.  Here the array is cut into
// NSX_BLK_MAX blocks of CAP = 1 << lg slots that are filled to CAP/2 on average; `dir` lists the blocks in
// preorder, a node's stored position is PHYSICAL (block * CAP + offset) and its preorder rank is
// prefix[dirpos[block]] + offset, with `prefix` (entries before each directory slot) recomputed by one scan of the
// directory per tree update.  Taking S out closes the gap inside at most two blocks and drops the blocks it covered
// entirely; putting S back behind its new parent shifts the tail of one block or spills into fresh blocks: at most
// 2 CAP entries move per pivot, plus |S|.  The directory, its inverse, the fill counts, the free list and the prefix
// sums live in the pivot CTA's shared memory (48 KB).  When the free list runs short the whole array is laid out
// afresh (O(n), rare).
// ------------------------------------------------------------------------------------------
#define NSX_BLK_MAX 4096
struct NsxBlk {
    uint16_t dir[NSX_BLK_MAX];      // block ids in preorder, ndir entries
    uint16_t dirpos[NSX_BLK_MAX];   // block id -> its slot in dir (blocks in use)
    uint16_t cnt[NSX_BLK_MAX];      // entries in use, by block id
    uint16_t free_[NSX_BLK_MAX];    // stack of unused block ids, nfree entries
    int32_t prefix[NSX_BLK_MAX + 1];  // prefix[k] = entries in dir[0 .. k)
    int32_t wsum[32];               // scan scratch (warp totals)
    int32_t ndir, nfree, rebuilds;
    int32_t lg, nb;                 // set before nsx_blk_init: log2 of the block capacity, blocks in the arena (<= NSX_BLK_MAX)
};
// smallest power-of-two block capacity for which a half-filled layout of n entries uses at most 5/8 of the blocks
static inline NSX_HD int32_t nsx_blk_lg(int64_t n) {
    int32_t lg = 5;
    while ((n + (1ll << (lg - 1)) - 1) >> (lg - 1) > NSX_BLK_MAX * 5 / 8) ++lg;
    return lg;
}
// preorder rank of a stored position (a no-op for dense arrays)
NSX_FN int32_t nsx_lpos(const NsxDev& d, int32_t pos) {
    const NsxBlk* B = d.blk;
    if (!B) return pos;
    return B->prefix[B->dirpos[pos >> B->lg]] + (pos & ((1 << B->lg) - 1));
}
// directory slot that holds preorder rank x (largest k with prefix[k] <= x)
NSX_FN int32_t nsx_blk_find(const NsxBlk& B, int32_t x) {
    int32_t lo = 0, hi = B.ndir - 1;
    while (lo < hi) {
        const int32_t mid = (lo + hi + 1) >> 1;
        if (B.prefix[mid] <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}
// physical index of preorder rank x
NSX_FN int64_t nsx_blk_phys(const NsxBlk& B, int32_t x) {
    const int32_t k = nsx_blk_find(B, x);
    return ((int64_t)B.dir[k] << B.lg) + (x - B.prefix[k]);
}

// 1 for an arc outside the tree that the Dantzig rule cannot treat as "candidate iff rc < -tol" (NsxCtl::n_special)
NSX_FN int32_t nsx_special(uint8_t st) {
    return (!(st & NSX_ARC_IN_TREE) && (st & (NSX_ARC_CAN_FWD | NSX_ARC_CAN_BWD)) != NSX_ARC_CAN_FWD) ? 1 : 0;
}
NSX_FN uint8_t nsx_bounds_bits(double f, double up, double tol) {
    uint8_t b = 0;
    if (nsx_isinf(up) || NSX_SUB(up, f) > tol) b |= NSX_ARC_CAN_FWD;
    if (f > tol) b |= NSX_ARC_CAN_BWD;
    return b;
}

// ------------------------------------------------------------------------------------------
// Exact potential recompute over the preorder range [lo, hi): pi[v] = pi[parent] +- cost(pred)
// parent-before-child (basis.py:111-117).  Chunks of the preorder array are processed in order;
// inside a chunk a node waits (in rounds) until its parent, which precedes it in preorder, is
// final.  `flags` is a per-thread-slot byte array of NSX_CHUNK entries.
// ------------------------------------------------------------------------------------------
#if NSX_ON_DEVICE
#ifndef NSX_THREADS
#define NSX_THREADS 512
#endif
#define NSX_CHUNK NSX_THREADS  // one preorder entry per thread
#else
#define NSX_CHUNK 256
#endif

// Where the parent of the entry at index x of the sequence `arr[lo .. hi)` sits in that sequence; anything below the
// current chunk start means "already final".  Dense preorder arrays (sidx == null): the parent's stored position.
// Blocked mode: the sequence is a flattened copy (d.tmp) and `sidx` maps a node to its index in it; the first entry
// hangs below a node outside the sequence, and so does every child of the root (whole-tree recompute).
NSX_FN int32_t nsx_seq_parent(const NsxDev& d, const int32_t* sidx, int64_t x, int64_t lo, int32_t parent) {
    if (x == lo) return -1;  // (dense arrays too: the stored position of that node may still await the deferred window shift)
    if (!sidx) return d.node[parent].pos;
    return parent == 0 ? -1 : sidx[parent];
}

struct NsxPotScratch {
    double val[NSX_CHUNK];
    double cst[NSX_CHUNK];         // signed cost to add
    int32_t par_local[NSX_CHUNK];  // parent index inside the chunk, -1 = parent already final
    int32_t dep[NSX_CHUNK];        // tree depth of a node that still waits for its parent, -1 = final
    int32_t dmin, dmax;            // depth range of the waiting nodes of the chunk
    int32_t rounds;
};

#if NSX_ON_DEVICE
// Device version: one preorder entry per thread.  The value of an entry is the left-to-right fold of
// the signed costs along its path from the nearest ancestor whose value is already known - the same
// additions, in the same order, as the reference's top-down walk.  Each thread climbs up to NSX_HOP
// ancestors inside the chunk (parent index + signed cost per entry live in shared memory), keeping
// the costs in registers, until it meets an entry that already holds a value (slots start as a NaN
// bit pattern that no potential can take; 8-byte shared-memory stores are single transactions);
// then it folds back down and publishes its own value.  If NSX_HOP hops were not enough it polls
// the ancestor it stopped at.  k tree levels therefore cost about k / NSX_HOP dependent waits
// instead of k, and no CTA-wide barrier.  Parents precede children in preorder, so waits terminate.
#define NSX_POT_EMPTY 0x7ff8dead0badc0deLL
#define NSX_HOP 8
NSX_FN void nsx_recompute_potentials(const NsxDev& d, int32_t phase, const int32_t* arr, const int32_t* sidx,
                                     int64_t lo, int64_t hi, NsxPotScratch& s, int64_t* rounds_out, NsxPiDelta* delta = nullptr) {
    int32_t rounds = 0;
    volatile long long* vbits = reinterpret_cast<volatile long long*>(s.val);
    for (int64_t c0 = lo; c0 < hi; c0 += NSX_CHUNK) {
        const int64_t c1 = c0 + NSX_CHUNK < hi ? c0 + NSX_CHUNK : hi;
        NSX_SYNC();  // scratch is free; potentials written for earlier chunks are visible
        const int32_t j = NSX_TID;
        const int64_t x = c0 + j;
        const bool active = j < NSX_CHUNK && x < c1;
        bool done = true;
        int32_t v = 0;
        double val = 0.0;
        if (active) {
            v = arr[x];
            const NsxNode r = d.node[v];
            double cst = nsx_arc_cost(d, phase, r.pred2 >> 1);
            // pred2 low bit set: arc points child->parent => pi[child] = pi[parent] - cost
            cst = (r.pred2 & 1) ? -cst : cst;
            const int32_t ppos = nsx_seq_parent(d, sidx, x, lo, r.parent);
            s.cst[j] = cst;
            if (ppos >= c0) { s.par_local[j] = (int32_t)(ppos - c0); done = false; vbits[j] = NSX_POT_EMPTY; }
            else { s.par_local[j] = -1; val = NSX_ADD(d.pi[r.parent], cst); vbits[j] = __double_as_longlong(val); }  // x - c == x + (-c) exactly
        }
        NSX_SYNC();
        // climb: c[0] = own cost, c[h] = cost of the h-th ancestor; `a` = first entry not yet folded in
        double c[NSX_HOP];
        int32_t len = 0, a = 0;
        long long base = 0;
        bool pending = !done;
        if (pending) {
            len = 1; a = s.par_local[j];
            c[0] = s.cst[j];
            base = vbits[a];
#pragma unroll
            for (int h = 1; h < NSX_HOP; ++h) {
                if (base == NSX_POT_EMPTY) {
                    c[h] = s.cst[a];
                    a = s.par_local[a];  // >= 0: an entry without a value has its parent inside the chunk
                    base = vbits[a];
                    len = h + 1;
                } else {
                    c[h] = 0.0;
                }
            }
        }
        // fold + publish inside a warp-uniform loop (a lane may be waiting for another lane of its own
        // warp: the publishing store must not sit behind a reconvergence point the waiter never reaches)
        __syncwarp();
        while (__any_sync(0xffffffffu, pending)) {
            if (pending) {
                if (base == NSX_POT_EMPTY) base = vbits[a];  // deeper than NSX_HOP: wait for that ancestor
                if (base != NSX_POT_EMPTY) {
                    val = __longlong_as_double(base);
#pragma unroll
                    for (int h = NSX_HOP - 1; h >= 0; --h)
                        if (h < len) val = NSX_ADD(val, c[h]);
                    vbits[j] = __double_as_longlong(val);
                    pending = false;
                }
            }
            ++rounds;
            __syncwarp();
        }
        __syncwarp();
        if (active) {
            d.pi[v] = val;
            if (d.pi_mirror) d.pi_mirror[v] = val;
            if (delta) { NsxPiDelta ent; ent.node = v; ent.pad = 0; ent.pi = val; delta[x - lo] = ent; }  // (for the sweep workers, NsxDev::pi_delta)
        }
    }
    NSX_SYNC();
    // statistics only: polls of warp 0 (it holds the first entries of every chunk), summed over chunks - kept free of
    // barriers: this runs once per pivot on the critical path
    if (rounds_out) { NSX_SINGLE { *rounds_out += rounds; } }
}
#else
// A chunk is finished level by level: a waiting node of depth L has its parent at depth L-1, which
// is either outside the chunk (final), or final from the chunk set-up, or was computed in the
// previous level step.  One barrier per tree level present in the chunk.
NSX_FN void nsx_recompute_potentials(const NsxDev& d, int32_t phase, const int32_t* arr, const int32_t* sidx,
                                     int64_t lo, int64_t hi, NsxPotScratch& s, int64_t* rounds_out, NsxPiDelta* delta = nullptr) {
    NSX_SINGLE { s.rounds = 0; }
    for (int64_t c0 = lo; c0 < hi; c0 += NSX_CHUNK) {
        int64_t c1 = c0 + NSX_CHUNK < hi ? c0 + NSX_CHUNK : hi;
        NSX_SYNC();
        NSX_SINGLE { s.dmin = 0x7fffffff; s.dmax = -1; }
        NSX_SYNC();
        NSX_PAR_FOR(x, c0, c1) {
            int32_t j = (int32_t)(x - c0);
            int32_t v = arr[x];
            NsxNode r = d.node[v];
            int32_t a = r.pred2 >> 1;
            double cst = nsx_arc_cost(d, phase, a);
            // pred2 low bit set: arc points child->parent => pi[child] = pi[parent] - cost
            cst = (r.pred2 & 1) ? -cst : cst;
            s.cst[j] = cst;
            int32_t ppos = nsx_seq_parent(d, sidx, x, lo, r.parent);
            if (ppos >= c0) {
                int32_t dep = d.depth[v];
                s.par_local[j] = (int32_t)(ppos - c0);
                s.dep[j] = dep;
                NSX_ATOMIC_MIN_I32(&s.dmin, dep);
                NSX_ATOMIC_MAX_I32(&s.dmax, dep);
            } else {
                s.par_local[j] = -1;
                s.dep[j] = -1;
                s.val[j] = NSX_ADD(d.pi[r.parent], cst);  // x - c == x + (-c) exactly
            }
        }
        NSX_SYNC();
        const int32_t dmin = s.dmin, dmax = s.dmax;
        for (int32_t lev = dmin; lev <= dmax; ++lev) {
            NSX_PAR_FOR(x, c0, c1) {
                int32_t j = (int32_t)(x - c0);
                if (s.dep[j] == lev) s.val[j] = NSX_ADD(s.val[s.par_local[j]], s.cst[j]);
            }
            NSX_SYNC();
        }
        NSX_SINGLE { if (dmax >= dmin) s.rounds += dmax - dmin + 1; }
        NSX_PAR_FOR(x, c0, c1) {
            int32_t j = (int32_t)(x - c0);
            int32_t v = arr[x];
            d.pi[v] = s.val[j];
            if (d.pi_mirror) d.pi_mirror[v] = s.val[j];
        }
    }
    NSX_SYNC();
    NSX_SINGLE { if (rounds_out) *rounds_out += s.rounds; }
}

#endif

// ---- blocked preorder array: maintenance (all routines are called by every thread of the pivot CTA) ----
// prefix[k] = entries in directory slots [0, k)
NSX_FN_COLD void nsx_blk_scan(NsxBlk& B) {
    NSX_SYNC();
#if NSX_ON_DEVICE
    const int T = NSX_NTHREADS, tid = NSX_TID, nd = B.ndir;
    const int per = (nd + T - 1) / T;
    const int k0 = tid * per < nd ? tid * per : nd, k1 = k0 + per < nd ? k0 + per : nd;
    int32_t sum = 0;
    for (int k = k0; k < k1; ++k) sum += B.cnt[B.dir[k]];
    int32_t incl = sum;
    __syncwarp();
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int32_t o = __shfl_up_sync(0xffffffffu, incl, off);
        if ((tid & 31) >= off) incl += o;
    }
    if ((tid & 31) == 31) B.wsum[tid >> 5] = incl;
    NSX_SYNC();
    if (tid < 32) {
        const int32_t w = tid < (T >> 5) ? B.wsum[tid] : 0;
        int32_t wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int32_t o = __shfl_up_sync(0xffffffffu, wi, off);
            if (tid >= off) wi += o;
        }
        B.wsum[tid] = wi - w;  // exclusive
    }
    NSX_SYNC();
    int32_t run = B.wsum[tid >> 5] + incl - sum;
    for (int k = k0; k < k1; ++k) { B.prefix[k] = run; run += B.cnt[B.dir[k]]; }
    if (k1 == nd && (k0 < nd || tid == 0)) B.prefix[nd] = run;
#else
    NSX_SINGLE {
        int32_t run = 0;
        for (int k = 0; k < B.ndir; ++k) { B.prefix[k] = run; run += B.cnt[B.dir[k]]; }
        B.prefix[B.ndir] = run;
    }
#endif
    NSX_SYNC();
}
// dst[x - skip] = node of preorder rank x, x in [skip, n); with `sidx` also the inverse map (prefix must be valid)
NSX_FN_COLD void nsx_blk_flatten(const NsxDev& d, const NsxBlk& B, int32_t* dst, int32_t skip, int32_t* sidx) {
    NSX_PAR_FOR(x, skip, d.n) {
        const int32_t v = d.order[nsx_blk_phys(B, (int32_t)x)];
        dst[x - skip] = v;
        if (sidx) sidx[v] = (int32_t)(x - skip);
    }
    NSX_SYNC();
}
// Lay the dense preorder sequence tmp[0 .. n) out in half-filled blocks: arena, stored positions, directory, free list.
NSX_FN_COLD void nsx_blk_layout(const NsxDev& d, NsxBlk& B) {
    const int32_t lg = B.lg, F = 1 << (lg - 1);
    const int32_t nblk = (d.n + F - 1) / F;
    NSX_SYNC();
    NSX_PAR_FOR(x, 0, d.n) {
        const int32_t v = d.tmp[x];
        const int32_t dst = (int32_t)(((x / F) << lg) + (x % F));
        d.order[dst] = v;
        d.node[v].pos = dst;
    }
    NSX_PAR_FOR(k, 0, B.nb) {
        if (k < nblk) {
            B.dir[k] = (uint16_t)k; B.dirpos[k] = (uint16_t)k;
            B.cnt[k] = (uint16_t)(d.n - k * F < F ? d.n - k * F : F);
        } else {
            B.free_[k - nblk] = (uint16_t)k; B.cnt[k] = 0;
        }
    }
    NSX_SINGLE { B.ndir = nblk; B.nfree = B.nb - nblk; }
    NSX_SYNC();
}
// first use: `order` still holds the dense preorder array the init kernel (or the warm-start layout) wrote
NSX_FN_COLD void nsx_blk_init(const NsxDev& d, NsxBlk& B) {
    NSX_SINGLE { B.rebuilds = 0; }
    NSX_PAR_FOR(x, 0, d.n) { d.tmp[x] = d.order[x]; }
    NSX_SYNC();
    nsx_blk_layout(d, B);
}
// the free list ran short: flatten and lay out afresh
NSX_FN_COLD void nsx_blk_rebuild(const NsxDev& d, NsxBlk& B) {
    nsx_blk_scan(B);
    nsx_blk_flatten(d, B, d.tmp, 0, (int32_t*)0);
    nsx_blk_layout(d, B);
    NSX_SINGLE { B.rebuilds++; }
}
// order[src .. src + len) moves by `delta` slots (the ranges may overlap; same block): rounds of 4 entries per thread,
// loaded before any is stored, taken from the end the data moves towards.  Stored positions follow.
NSX_FN void nsx_blk_move(const NsxDev& d, int64_t src, int32_t len, int32_t delta) {
    const int64_t T = NSX_NTHREADS, R = 4 * T;
    const int64_t nr = (len + R - 1) / R;
    for (int64_t b = 0; b < nr; ++b) {
        const int64_t b_lo = delta > 0 ? src + len - (b + 1) * R : src + b * R;
        int32_t v[4];
        NSX_SYNC();  // the previous round has been written
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t x = b_lo + u * T + NSX_TID;
            v[u] = (x >= src && x < src + len) ? d.order[x] : -1;
        }
        NSX_SYNC();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t x = b_lo + u * T + NSX_TID;
            if (v[u] >= 0) { d.order[x + delta] = v[u]; d.node[v[u]].pos = (int32_t)(x + delta); }
        }
    }
}
// directory slots [from, from + count) move by `delta` (same scheme), dirpos follows
NSX_FN void nsx_blk_dirshift(NsxBlk& B, int32_t from, int32_t count, int32_t delta) {
    const int32_t T = NSX_NTHREADS, R = 4 * T;
    const int32_t nr = (count + R - 1) / R;
    for (int32_t b = 0; b < nr; ++b) {
        const int32_t b_lo = delta > 0 ? from + count - (b + 1) * R : from + b * R;
        int32_t v[4];
        NSX_SYNC();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int32_t k = b_lo + u * T + NSX_TID;
            v[u] = (k >= from && k < from + count) ? (int32_t)B.dir[k] : -1;
        }
        NSX_SYNC();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int32_t k = b_lo + u * T + NSX_TID;
            if (v[u] >= 0) { B.dir[k + delta] = (uint16_t)v[u]; B.dirpos[v[u]] = (uint16_t)(k + delta); }
        }
    }
}
// Take the preorder ranks [S0, S0 + sz) out (their nodes have been copied elsewhere).  `prefix` must be valid on
// entry and is stale afterwards.  Returns the number of entries that moved.
NSX_FN_COLD int32_t nsx_blk_remove(const NsxDev& d, NsxBlk& B, int32_t S0, int32_t sz) {
    const int32_t lg = B.lg;
    const int32_t k0 = nsx_blk_find(B, S0), k1 = nsx_blk_find(B, S0 + sz - 1);
    const int32_t b0 = B.dir[k0], b1 = B.dir[k1];
    const int32_t o0 = S0 - B.prefix[k0], o1 = S0 + sz - B.prefix[k1];  // first offset taken in b0 / first kept in b1
    const int32_t c1 = B.cnt[b1], nd = B.ndir, nf = B.nfree;
    NSX_SYNC();  // every thread holds the values above
    int32_t newc0, newc1;
    if (k0 == k1) { nsx_blk_move(d, ((int64_t)b0 << lg) + o1, c1 - o1, -sz); newc0 = newc1 = c1 - sz; }
    else { nsx_blk_move(d, ((int64_t)b1 << lg) + o1, c1 - o1, -o1); newc0 = o0; newc1 = c1 - o1; }
    const bool keepA = newc0 > 0, keepB = k1 != k0 && newc1 > 0;
    const int32_t w = k0 + (keepA ? 1 : 0) + (keepB ? 1 : 0);  // where the directory tail lands
    const int32_t shift = (k1 + 1) - w;
    const int32_t mid = k1 > k0 ? k1 - k0 - 1 : 0;
    NSX_PAR_FOR(k, k0 + 1, k1) { B.free_[nf + (k - k0 - 1)] = B.dir[k]; }  // blocks that S covered entirely
    NSX_SYNC();
    NSX_SINGLE {
        int32_t f = nf + mid, ww = k0;
        B.cnt[b0] = (uint16_t)newc0;
        if (k1 != k0) B.cnt[b1] = (uint16_t)newc1;
        if (keepA) { B.dir[ww] = (uint16_t)b0; B.dirpos[b0] = (uint16_t)ww; ++ww; } else B.free_[f++] = (uint16_t)b0;
        if (k1 != k0) { if (keepB) { B.dir[ww] = (uint16_t)b1; B.dirpos[b1] = (uint16_t)ww; ++ww; } else B.free_[f++] = (uint16_t)b1; }
        B.nfree = f;
        B.ndir = nd - shift;
    }
    if (shift > 0) nsx_blk_dirshift(B, k1 + 1, nd - (k1 + 1), -shift);
    NSX_SYNC();
    return c1 - o1;
}
// Put the sz nodes tmp[0 .. sz) right behind node p in preorder.  Needs nfree >= (sz + CAP) / CAP + 1.
NSX_FN_COLD int32_t nsx_blk_insert(const NsxDev& d, NsxBlk& B, int32_t p, int32_t sz) {
    const int32_t lg = B.lg, CAP = 1 << lg;
    const int32_t pp = d.node[p].pos;  // (p may have moved when S was taken out)
    const int32_t bp = pp >> lg, op = pp & (CAP - 1), kp = B.dirpos[bp], cp = B.cnt[bp];
    const int32_t t = cp - (op + 1);   // entries behind p in its block
    const int64_t base = ((int64_t)bp << lg) + op + 1;
    const int32_t nd = B.ndir, nf = B.nfree;
    NSX_SYNC();
    if (cp + sz <= CAP) {
        nsx_blk_move(d, base, t, sz);
        NSX_PAR_FOR(j, 0, sz) {
            const int32_t v = d.tmp[j];
            d.order[base + j] = v;
            d.node[v].pos = (int32_t)(base + j);
        }
        NSX_SINGLE { B.cnt[bp] = (uint16_t)(cp + sz); }
    } else {
        // S and the tail of p's block go to fresh blocks (half filled while the free list allows)
        const int32_t total = sz + t;
        int32_t F = CAP >> 1, q = (total + F - 1) / F;
        if (q > nf) { F = CAP; q = (total + F - 1) / F; }
        NSX_PAR_FOR(j, 0, total) {
            const int32_t v = j < sz ? d.tmp[j] : d.order[base + (j - sz)];
            const int32_t nb = B.free_[nf - 1 - (int32_t)(j / F)];
            const int64_t dst = ((int64_t)nb << lg) + (j % F);
            d.order[dst] = v;
            d.node[v].pos = (int32_t)dst;
        }
        nsx_blk_dirshift(B, kp + 1, nd - (kp + 1), q);
        NSX_PAR_FOR(i, 0, q) {
            const int32_t nb = B.free_[nf - 1 - (int32_t)i];
            B.dir[kp + 1 + i] = (uint16_t)nb;
            B.dirpos[nb] = (uint16_t)(kp + 1 + i);
            B.cnt[nb] = (uint16_t)(total - (int32_t)i * F < F ? total - (int32_t)i * F : F);
        }
        NSX_SINGLE { B.cnt[bp] = (uint16_t)(op + 1); B.ndir = nd + q; B.nfree = nf - q; }
    }
    NSX_SYNC();
    return t;
}
// Potentials of the whole tree (start of a phase): dense arrays walk `order`, the blocked array is flattened first.
// BLK (here and below): the blocked-array code is compiled in; kernels for trees that fit the pivot CTA's shared memory are
// instantiated without it (smaller register footprint of the latency-bound pivot code).
template <bool BLK>
NSX_FN void nsx_recompute_all_potentials(const NsxDev& d, int32_t phase, NsxPotScratch& ps) {
    if (BLK && d.blk) {
        nsx_blk_scan(*d.blk);
        nsx_blk_flatten(d, *d.blk, d.tmp, 1, d.sidx);
        nsx_recompute_potentials(d, phase, d.tmp, d.sidx, 0, (int64_t)d.n - 1, ps, (int64_t*)0);
    } else {
        nsx_recompute_potentials(d, phase, d.order, (const int32_t*)0, 1, d.n, ps, (int64_t*)0);
    }
}

// ------------------------------------------------------------------------------------------
// Cycle walk: tree paths from both ends of the entering arc up to their join, found with the
// preorder ancestor test  pos[u] <= pos[x] < pos[u] + size[u]  (no depth comparison needed).
// Executed by two lanes on the device (one per side), serially in the emulation.
// ------------------------------------------------------------------------------------------
NSX_FN void nsx_walk_side(const NsxDev& d, int32_t from, int32_t other_pos, int32_t* spath,
                          int32_t* gpath, int32_t* len_out, int32_t* join_out) {
    int32_t u = from;
    int32_t len = 0;
    NsxNode r = d.node[u];
    while (!(r.pos <= other_pos && other_pos < r.pos + r.size)) {
        if (len < NSX_PATH_CAP) spath[len] = u; else gpath[len] = u;
        ++len;
        u = r.parent;
        r = d.node[u];
    }
    *len_out = len;
    *join_out = u;
}

// ------------------------------------------------------------------------------------------
// Ratio-test reduction record: m1 = smallest residual, a1 / k1 = lowest arc index (and its scan
// position) among residuals equal to m1, m2 = smallest residual strictly above m1.
// ------------------------------------------------------------------------------------------
struct NsxRatio { double m1, m2; int32_t a1, k1; };
NSX_FN void nsx_ratio_init(NsxRatio& r) { r.m1 = NSX_INF; r.m2 = NSX_INF; r.a1 = 0x7fffffff; r.k1 = -1; }
NSX_FN void nsx_ratio_add(NsxRatio& r, double v, int32_t a, int32_t k) {
    if (v < r.m1) { r.m2 = r.m1; r.m1 = v; r.a1 = a; r.k1 = k; }
    else if (v == r.m1) { if (a < r.a1) { r.a1 = a; r.k1 = k; } }
    else if (v < r.m2) r.m2 = v;
}
NSX_FN void nsx_ratio_merge(NsxRatio& x, const NsxRatio& y) {
    if (y.m1 < x.m1) {
        x.m2 = x.m1 < y.m2 ? x.m1 : y.m2;
        x.m1 = y.m1; x.a1 = y.a1; x.k1 = y.k1;
    } else if (y.m1 == x.m1) {
        if (y.m2 < x.m2) x.m2 = y.m2;
        if (y.a1 < x.a1) { x.a1 = y.a1; x.k1 = y.k1; }
    } else if (y.m1 < x.m2) {
        x.m2 = y.m1;
    }
}
// Decide from the reduced record, or replay the reference's sequential scan (single thread).
NSX_FN void nsx_ratio_finish(const NsxRatio& rr, const double* res, const int32_t* arc2, int32_t ncyc,
                             int32_t e, double tol, NsxPivotScratch& s) {
    double theta; int32_t leave, leave_k;
    const bool safe = nsx_isinf(rr.m1) || NSX_SUB(rr.m2, rr.m1) > NSX_MUL(4.0, tol);
    if (safe) {
        theta = rr.m1;
        leave = nsx_isinf(rr.m1) ? e : rr.a1;
        leave_k = nsx_isinf(rr.m1) ? ncyc - 1 : rr.k1;
    } else {
        double best = -NSX_INF;
        theta = NSX_INF; leave = e; leave_k = ncyc - 1;
        for (int32_t k = 0; k < ncyc; ++k) {
            double r = res[k];
            int32_t a = arc2[k] >> 1;
            if (r < NSX_SUB(theta, tol)) {
                theta = r; leave = a; best = r; leave_k = k;
            } else if (fabs(NSX_SUB(r, theta)) <= tol) {
                if (r > NSX_ADD(best, tol) || (fabs(NSX_SUB(r, best)) <= tol && a < leave)) {
                    leave = a; best = r; leave_k = k;
                }
            }
        }
    }
    s.theta = theta;
    s.leave_arc = leave;
    s.leave_k = leave_k;
    s.art_delta = 0;
}

// ------------------------------------------------------------------------------------------
// One pivot on entering arc e (direction dir = +1 forward / -1 backward).
// Returns (block-uniform) 0 = ok, 3 = unbounded.
// ------------------------------------------------------------------------------------------
// Deferred tree bookkeeping of the last pivot (all threads of the pivot CTA; a no-op when nothing is pending).
template <bool BLK>
NSX_FN void nsx_pivot_flush(const NsxDev& d, NsxCtl& c, NsxPivotScratch& s) {
    const int32_t kind = s.def_kind;
    if (!kind) return;
    int64_t moved;
    if (BLK && kind == 2) {
        moved = s.def_sz + nsx_blk_remove(d, *d.blk, (int32_t)s.def_S0, s.def_sz);
        moved += nsx_blk_insert(d, *d.blk, s.def_p, s.def_sz);
    } else {
        const int64_t lo = s.def_lo, hi = s.def_hi, S0 = s.def_S0, S1 = s.def_S1, xshift = s.def_xshift;
        const int64_t T = NSX_NTHREADS;
        NSX_SYNC();
        // entries between the old and the new place of S move by +-|S| (into tmp, next to the permuted S) ...
        for (int64_t x0 = lo + NSX_TID; x0 < hi; x0 += 4 * T) {
            int32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { int64_t x = x0 + u * T; v[u] = (x < hi && !(x >= S0 && x < S1)) ? d.order[x] : -1; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int64_t x = x0 + u * T;
                if (v[u] >= 0) { d.tmp[x + xshift] = v[u]; d.node[v[u]].pos = (int32_t)(x + xshift); }
            }
        }
        NSX_SYNC();
        // ... and the window goes back to the preorder array
        for (int64_t x0 = lo + NSX_TID; x0 < hi; x0 += 4 * T) {
            int32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { int64_t x = x0 + u * T; v[u] = x < hi ? d.tmp[x] : -1; }
#pragma unroll
            for (int u = 0; u < 4; ++u) { int64_t x = x0 + u * T; if (x < hi) d.order[x] = v[u]; }
        }
        moved = hi - lo;
    }
    NSX_SYNC();
    NSX_SINGLE { s.def_kind = 0; c.sum_window += moved; }
    NSX_SYNC();
}

// star pricing, end of a pivot (one thread): the row of the entering arc joins the work list unless its node is in it already
NSX_FN void nsx_star_close(const NsxDev& d, NsxCtl& c) {
    if (!c.star_on || !c.star_valid) return;
    const int32_t x = c.star_extra;
    c.star_ne = c.star_nd;
    if (x >= 0 && d.dstamp[x] != c.star_round) {
        int32_t* info = d.dinfo + 8 * c.star_nd;
        const int32_t rb = d.row_begin[x];
        info[0] = x; info[1] = rb; info[2] = d.row_begin[x + 1] - rb; info[3] = 0; info[4] = 0;
        c.star_ne = c.star_nd + 1;
    }
}

template <bool BLK>
NSX_FN int nsx_pivot(const NsxDev& d, NsxCtl& c, NsxPivotScratch& s, NsxPotScratch& ps, int32_t e,
                     int32_t dir, int32_t want_weight) {
    const double tol = d.tol;
    const int32_t t = dir == 1 ? d.tail[e] : d.head[e];
    const int32_t h = dir == 1 ? d.head[e] : d.tail[e];

    nsx_pivot_flush<BLK>(d, c, s);  // (usually done already, while the sweep workers priced)
    long long tph = NSX_CLOCK();
    // star pricing: this pivot opens a new round; its re-hung subtree (recorded after the potential recompute) and the
    // row of the entering arc are what the next pricing step has to look at
    NSX_SINGLE {
        s.spec_e0 = -1;  // (nsx_special of the entering arc before this pivot: set by the flow update when it runs)
        if (c.star_on) {
            if (c.star_pending >= 1) c.star_valid = 0;  // two pivots without a pricing step in between
            c.star_pending++;
            c.star_round++;
            c.star_nd = 0;
            c.star_extra = d.tail[e];
            if (c.star_valid) {
                NsxRC none; none.key = 0.0; none.arc2 = -1; none.pad = c.star_round;
                d.rc[d.tail[e]] = none;
                if (c.star_on == 2) d.rc[d.n + d.tail[e]] = none;
            }
        }
    }
    // ---- 1. walk both sides up to the join ------------------------------------------------
#if NSX_ON_DEVICE
    if (d.scan_walk && c.avg_cycle > 40 * 16) {
        // (long cycles expected: recent average, in 1/16 arcs, above 40)
        // Every node tests in parallel whether it is an ancestor of h / of t (preorder interval
        // test).  Ancestors of exactly one endpoint are the cycle; the slot of such a node in its
        // side's path is its depth distance to the endpoint.  Hits are remembered in registers and
        // their depths are fetched after the scan, all at once (depth[] may live in L2), so the scan
        // itself has no dependent load and no atomics.  The join is the common ancestor of smallest
        // size (warp-reduced, one shared-memory atomic per warp).
        NSX_SINGLE { s.jkey = 0x7fffffff; s.nh = 0; s.nt = 0; }
        const int32_t ph = d.node[h].pos, pt = d.node[t].pos;
        const int32_t dh = d.depth[h], dt = d.depth[t];
        NSX_SYNC();
        int32_t jk = 0x7fffffff;
        int32_t hit_w[2] = {-1, -1};
        bool hit_h[2] = {false, false};
        int32_t cnt_h = 0, cnt_t = 0;  // path lengths = number of ancestors of exactly one endpoint (no depth[join] round trip)
        NSX_PAR_FOR(w, 0, d.n) {
            const NsxNode rec = d.node[w];
            const bool in_h = rec.pos <= ph && ph < rec.pos + rec.size;
            const bool in_t = rec.pos <= pt && pt < rec.pos + rec.size;
            if (in_h && in_t) {
                const int32_t key = (rec.size << 16) | (int32_t)w;
                jk = key < jk ? key : jk;
            } else if (in_h || in_t) {
                if (in_h) ++cnt_h; else ++cnt_t;
                if (hit_w[0] < 0) { hit_w[0] = (int32_t)w; hit_h[0] = in_h; }
                else if (hit_w[1] < 0) { hit_w[1] = (int32_t)w; hit_h[1] = in_h; }
                else {  // a third hit in one thread (rare): place it right away
                    const int32_t k = (in_h ? dh : dt) - d.depth[w];
                    int32_t* sp = in_h ? s.path_h : s.path_t;
                    int32_t* gp = in_h ? d.gpath_h : d.gpath_t;
                    if (k < NSX_PATH_CAP) sp[k] = (int32_t)w; else gp[k] = (int32_t)w;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (hit_w[i] >= 0) {
                const int32_t k = (hit_h[i] ? dh : dt) - d.depth[hit_w[i]];
                int32_t* sp = hit_h[i] ? s.path_h : s.path_t;
                int32_t* gp = hit_h[i] ? d.gpath_h : d.gpath_t;
                if (k < NSX_PATH_CAP) sp[k] = hit_w[i]; else gp[k] = hit_w[i];
            }
        }
        __syncwarp();
        jk = __reduce_min_sync(0xffffffffu, jk);
        cnt_h = __reduce_add_sync(0xffffffffu, cnt_h);
        cnt_t = __reduce_add_sync(0xffffffffu, cnt_t);
        if ((threadIdx.x & 31) == 0) {
            if (jk != 0x7fffffff) NSX_ATOMIC_MIN_I32(&s.jkey, jk);
            if (cnt_h) NSX_ATOMIC_ADD_I32(&s.nh, cnt_h);
            if (cnt_t) NSX_ATOMIC_ADD_I32(&s.nt, cnt_t);
        }
        NSX_SYNC();
        NSX_SINGLE { s.join = s.jkey & 0xffff; }
    } else if (BLK && d.blk) {
        // Depth-synchronised climb (no positions needed): lane 0 holds the head-side node, lane 1
        // the tail-side node; the deeper one climbs, both climb when level, until they meet.
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            int32_t du = d.depth[h], dv = d.depth[t];
            int32_t me = lane == 0 ? h : t;
            int32_t len = 0;
            for (;;) {
                __syncwarp();
                const int32_t u = __shfl_sync(0xffffffffu, me, 0), v = __shfl_sync(0xffffffffu, me, 1);
                if (u == v) break;
                const bool climb_u = du >= dv, climb_v = dv >= du;
                if ((lane == 0 && climb_u) || (lane == 1 && climb_v)) {
                    if (len < NSX_PATH_CAP) (lane == 0 ? s.path_h : s.path_t)[len] = me;
                    else (lane == 0 ? d.gpath_h : d.gpath_t)[len] = me;
                    ++len;
                    me = nsx_parent(d, me);
                }
                du -= climb_u ? 1 : 0;
                dv -= climb_v ? 1 : 0;
            }
            if (lane == 0) { s.nh = len; s.join = me; }
            if (lane == 1) s.nt = len;
        }
    } else if (threadIdx.x < 2) {
        int32_t from = threadIdx.x == 0 ? h : t;
        int32_t opos = d.node[threadIdx.x == 0 ? t : h].pos;
        int32_t len, join;
        nsx_walk_side(d, from, opos, threadIdx.x == 0 ? s.path_h : s.path_t,
                      threadIdx.x == 0 ? d.gpath_h : d.gpath_t, &len, &join);
        if (threadIdx.x == 0) { s.nh = len; s.join = join; } else { s.nt = len; }
    }
#else
    NSX_HOST_SERIAL {
    if (BLK && d.blk) {
        int32_t u = h, v = t, du = d.depth[h], dv = d.depth[t], nh_ = 0, nt_ = 0;
        while (u != v) {
            const bool climb_u = du >= dv, climb_v = dv >= du;
            if (climb_u) { if (nh_ < NSX_PATH_CAP) s.path_h[nh_] = u; else d.gpath_h[nh_] = u; ++nh_; u = nsx_parent(d, u); --du; }
            if (climb_v) { if (nt_ < NSX_PATH_CAP) s.path_t[nt_] = v; else d.gpath_t[nt_] = v; ++nt_; v = nsx_parent(d, v); --dv; }
        }
        s.nh = nh_; s.nt = nt_; s.join = u;
    } else {
        int32_t len, join;
        nsx_walk_side(d, h, d.node[t].pos, s.path_h, d.gpath_h, &len, &join);
        s.nh = len; s.join = join;
        nsx_walk_side(d, t, d.node[h].pos, s.path_t, d.gpath_t, &len, &join);
        s.nt = len;
    }
    }
#endif
    NSX_SYNC();
    NSX_PH(c, 0, tph);
    const int32_t nh = s.nh, nt = s.nt;
    const int32_t ncyc = nh + nt + 1;
    const bool spill = nh > NSX_PATH_CAP || nt > NSX_PATH_CAP;
    if (spill) {  // long cycle: the shared-memory prefixes join the rest of the paths in HBM
        NSX_PAR_FOR(i, 0, nh < NSX_PATH_CAP ? nh : NSX_PATH_CAP) { d.gpath_h[i] = s.path_h[i]; }
        NSX_PAR_FOR(i, 0, nt < NSX_PATH_CAP ? nt : NSX_PATH_CAP) { d.gpath_t[i] = s.path_t[i]; }
        NSX_SYNC();
    }
    const int32_t* path_h = spill ? d.gpath_h : s.path_h;
    const int32_t* path_t = spill ? d.gpath_t : s.path_t;
    int32_t* arc2 = spill ? d.garc2 : s.arc2;
    double* res = spill ? d.gres : s.res;

    // ---- 2. residuals in the reference's scan order (simplex.py:1198-1229) ----------------
    NSX_PAR_FOR(k, 0, ncyc) {
        int32_t a, sign;
        if (k < nh) {
            int32_t p2 = d.node[path_h[k]].pred2;
            a = p2 >> 1;
            sign = (p2 & 1) ? 1 : -1;  // walking child->parent: arc pointing child->parent is traversed forward
        } else if (k < nh + nt) {
            int32_t p2 = d.node[path_t[nt - 1 - (k - nh)]].pred2;
            a = p2 >> 1;
            sign = (p2 & 1) ? -1 : 1;  // walking parent->child
        } else {
            a = e;
            sign = dir;
        }
        double f = d.flow[a];
        if (c.warm && (d.state[a] & NSX_ARC_STALE)) {  // the reference's residual mirrors still hold the cold-start flow
            double up0 = nsx_upper(d, a);               // (see NSX_ARC_STALE): 0 on a real arc, |supply| on an artificial one
            f = (a < d.m || nsx_isinf(up0)) ? 0.0 : up0;
        }
        double r;
        if (sign == 1) {
            double up = nsx_upper(d, a);
            r = nsx_isinf(up) ? NSX_INF : NSX_SUB(up, f);
        } else {
            r = f;  // flow - lower with lower = 0
        }
        arc2[k] = a * 2 + (sign < 0 ? 1 : 0);
        res[k] = r;
    }
    NSX_SYNC();
    NSX_PH(c, 1, tph);

    // ---- 3. ratio test -------------------------------------------------------------------
    // The reference scans the cycle sequentially with tolerance comparisons (simplex.py:1205-1229).
    // When no residual lies within (theta*, theta* + 4 tol] of the minimum theta*, that scan provably
    // returns theta* and the lowest arc index among the residuals equal to theta*, which one warp
    // finds with a shuffle reduction; otherwise thread 0 replays the sequential scan.
#if NSX_ON_DEVICE
    if (threadIdx.x < 32) {
        NsxRatio rr; nsx_ratio_init(rr);
        for (int32_t k = threadIdx.x; k < ncyc; k += 32) nsx_ratio_add(rr, res[k], arc2[k] >> 1, k);
        __syncwarp();
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            NsxRatio o;
            o.m1 = __shfl_down_sync(0xffffffffu, rr.m1, off);
            o.m2 = __shfl_down_sync(0xffffffffu, rr.m2, off);
            o.a1 = __shfl_down_sync(0xffffffffu, rr.a1, off);
            o.k1 = __shfl_down_sync(0xffffffffu, rr.k1, off);
            nsx_ratio_merge(rr, o);
        }
        if (threadIdx.x == 0) nsx_ratio_finish(rr, res, arc2, ncyc, e, tol, s);
    }
#else
    NSX_HOST_SERIAL {
        NsxRatio lanes[4];
        for (int l = 0; l < 4; ++l) nsx_ratio_init(lanes[l]);
        for (int32_t k = 0; k < ncyc; ++k) nsx_ratio_add(lanes[k & 3], res[k], arc2[k] >> 1, k);
        for (int l = 1; l < 4; ++l) nsx_ratio_merge(lanes[0], lanes[l]);
        nsx_ratio_finish(lanes[0], res, arc2, ncyc, e, tol, s);
    }
#endif
    NSX_SYNC();
    NSX_PH(c, 2, tph);
    if (nsx_isinf(s.theta)) {
        NSX_SINGLE {
            c.unbounded_arc = e;
            double rc = NSX_SUB(NSX_ADD(nsx_arc_cost(d, c.phase, e), d.pi[d.tail[e]]), d.pi[d.head[e]]);
            c.unbounded_rc = dir == 1 ? rc : -rc;
        }
        return 3;
    }
    const double theta = s.theta > 0.0 ? s.theta : 0.0;
    const int32_t leave = s.leave_arc;
    const int32_t leave_k = s.leave_k;

    // ---- 4. flow update (simplex.py:1255-1283); nothing changes when theta == 0 - except after a warm start,
    // where the pass also drops the NSX_ARC_STALE marks of the cycle arcs (their residual mirrors get refreshed)
    if (theta > 0.0 || c.warm) {
        NSX_PAR_FOR(k, 0, ncyc) {
            int32_t a = arc2[k] >> 1;
            int32_t sign = (arc2[k] & 1) ? -1 : 1;
            double old = d.flow[a];
            double up = nsx_upper(d, a);
            double f = sign > 0 ? NSX_ADD(old, theta) : NSX_SUB(old, theta);
            uint8_t st = d.state[a];
            if (k == ncyc - 1) s.spec_e0 = a < d.m ? nsx_special(st) : 0;  // (the entering arc is the last of the scan order)
            if (theta > 0.0) st |= NSX_ARC_TOUCHED;
            const double pushed = f;
            if (f < NSX_SUB(0.0, tol)) { f = 0.0; st &= (uint8_t)~NSX_ARC_TOUCHED; }
            if (!nsx_isinf(up) && f > NSX_ADD(up, tol)) { f = up; st &= (uint8_t)~NSX_ARC_TOUCHED; }
            if (d.imbalance && f != pushed) {
                // A clamp takes flow out of (or puts it into) the arc without a matching change elsewhere. It cannot
                // happen from a consistent state; after a warm start the stale residuals (NSX_ARC_STALE) make the ratio
                // test over-push, the reference clamps silently (simplex.py:1263-1266) and only its conservation check
                // after Phase 1 notices.  Every other flow change is a cycle push, so the node balances are exactly the
                // sums of these corrections.
                const double fix = NSX_SUB(f, pushed);
                NSX_ATOMIC_ADD_F64(&d.imbalance[nsx_tail(d, a)], -fix);
                NSX_ATOMIC_ADD_F64(&d.imbalance[nsx_head(d, a)], fix);
            }
            d.flow[a] = f;
            st = (uint8_t)((st & (NSX_ARC_IN_TREE | NSX_ARC_TOUCHED)) | nsx_bounds_bits(f, up, tol));
            d.state[a] = st;
            if (d.csc_state && a < d.m) d.csc_state[d.csc_pos[a]] = st;
            if (a >= d.m) {
                int had = old > tol, has = f > tol;
                if (had != has) NSX_ATOMIC_ADD_I32(&s.art_delta, has - had);
            }
        }
    }
    NSX_SYNC();
    NSX_PH(c, 3, tph);
    NSX_SINGLE {
        c.art_with_flow += s.art_delta;
        if (theta <= tol) c.degenerate++;
        c.sum_cycle += ncyc;
        c.avg_cycle += ncyc - (c.avg_cycle >> 4);  // running mean of the cycle length, scaled by 16
        // Devex weight := number of tree arcs on the tail-head path, written by pricing before
        // the pivot in the reference (simplex_pricing.py:350-352)
        if (want_weight && e < d.m) {
            d.wgt[e] = (c.wepoch << 24) | (uint32_t)(nh + nt);
            if (d.csc_wgt) d.csc_wgt[d.csc_pos[e]] = d.wgt[e];
        }
        int is_deg = (leave == e) || (fabs(theta) < tol);  // simplex.py:1317
        c.tuner_total++;
        if (is_deg) c.tuner_deg++;
        if (leave == e) {  // bound flip (simplex.py:1320-1334)
            c.last_deg = e;
            if (e < d.m && s.spec_e0 >= 0) c.n_special += nsx_special(d.state[e]) - s.spec_e0;  // (no flow update: nothing changed)
        }
    }
    NSX_SYNC();
    NSX_PH(c, 4, tph);
    if (leave == e) {  // tree unchanged
        NSX_SINGLE { nsx_star_close(d, c); }
        return 0;
    }

    // ---- 5. tree update: re-hang the subtree below the leaving arc under the entering arc --
    // stem s_0 = q (entering endpoint inside the cut subtree) ... s_k = r (its pred arc leaves)
    const bool on_h = leave_k < nh;
    const int32_t kk = on_h ? leave_k : nt - 1 - (leave_k - nh);  // index of r in its side path
    const int32_t* spath = on_h ? path_h : path_t;                // side containing the stem
    const int32_t* opath = on_h ? path_t : path_h;                // side of p
    const int32_t slen = on_h ? nh : nt;
    const int32_t olen = on_h ? nt : nh;
    const int32_t p = on_h ? t : h;
    const int32_t r = spath[kk];
    const NsxNode rec_r = d.node[r];
    const NsxNode rec_p = d.node[p];
    const int32_t sz = rec_r.size;
    const int32_t depth_q_new = d.depth[p] + 1;
    if (BLK && d.blk) {
        // (uniform: every thread reads the same shared-memory words) enough free blocks for the worst case of this update?
        if (d.blk->nfree < (sz >> d.blk->lg) + 3) { nsx_blk_rebuild(d, *d.blk); NSX_SINGLE { c.blk_rebuilds++; } }
        nsx_blk_scan(*d.blk);  // preorder ranks of stored positions are valid from here until the array is edited
    }
    // stem snapshot (old pos / size / depth / pred2): shared scratch reusing res[] / arc2[], or the
    // global spill arrays when the cycle did not fit (garc2: 2n+1 ints, gres: 2n+1 doubles)
    int32_t* st_pos = spill ? d.garc2 : (int32_t*)s.res;
    int32_t* st_size = spill ? d.garc2 + d.n : ((int32_t*)s.res) + NSX_PATH_CAP;
    int32_t* st_depth = spill ? (int32_t*)d.gres : ((int32_t*)s.res) + 2 * NSX_PATH_CAP;
    int32_t* st_pred2 = spill ? ((int32_t*)d.gres) + d.n : s.arc2;
    NSX_SYNC();  // everyone has read arc2/res for the flow update before they are reused
    NSX_PAR_FOR(i, 0, kk + 2) {
        if (i == kk + 1) { s.p_pos = BLK ? nsx_lpos(d, d.node[p].pos) : d.node[p].pos; continue; }  // preorder rank of the new parent
        NsxNode x = d.node[spath[i]];
        st_pos[i] = BLK ? nsx_lpos(d, x.pos) : x.pos; st_size[i] = x.size; st_pred2[i] = x.pred2;
        st_depth[i] = d.depth[spath[i]];
    }
    // subtree-size bookkeeping of the untouched ancestors on both sides of the cycle
    NSX_PAR_FOR(i, kk + 1, slen) { d.node[spath[i]].size -= sz; }
    NSX_PAR_FOR(i, 0, olen) { d.node[opath[i]].size += sz; }
    NSX_SYNC();
    NSX_PH(c, 5, tph);

    // The block S = [a0, a0+sz) of the preorder sequence is re-rooted at q and moved under p.  Dense array: either
    // right behind p or to the end of p's old subtree, whichever shifts fewer entries (both are valid preorders);
    // `ins` is the insertion point in old coordinates.  Blocked array: always right behind p.
    const int32_t a0 = st_pos[kk], P = s.p_pos;  // r = s_k is the root of the cut subtree
    const int64_t S0 = a0, S1 = (int64_t)a0 + sz;
    const int64_t insA = (int64_t)P + 1, insB = (int64_t)P + rec_p.size;
    const int64_t dA = insA <= S0 ? S0 - insA : insA - S1;
    const int64_t dB = insB <= S0 ? S0 - insB : insB - S1;
    const int64_t ins = dB < dA ? insB : insA;
    int64_t lo, hi, s_base, xshift;
    if (ins <= S0) { lo = ins; hi = S1; s_base = ins; xshift = sz; }        // [ins, S0) moves right
    else           { lo = S0; hi = ins; s_base = ins - sz; xshift = -(int64_t)sz; }  // [S1, ins) moves left
    const int32_t k_stem = kk;
    // Blocked mode: only S is permuted (into tmp[0 .. sz), with the inverse map in sidx); the dense array also shifts
    // the entries between the old and the new place of S in the same pass.
    const bool blocked = BLK && d.blk != nullptr;
    const int64_t w_lo = S0, w_hi = S1;  // (the entries between the old and the new place of S: nsx_pivot_flush)
    {
        // four entries per thread and step: the loads of all four are issued before any store
        const int64_t T = NSX_NTHREADS;
        for (int64_t x0 = w_lo + NSX_TID; x0 < w_hi; x0 += 4 * T) {
            int32_t v[4], dv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int64_t x = x0 + u * T;
                v[u] = x < w_hi ? d.order[blocked ? nsx_blk_phys(*d.blk, (int32_t)x) : x] : -1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int64_t x = x0 + u * T;
                dv[u] = (v[u] >= 0 && x >= S0 && x < S1) ? d.depth[v[u]] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t x = x0 + u * T;
                if (v[u] < 0) continue;
                int64_t fx;
                if (x >= S0 && x < S1) {
                    // smallest i with x inside old subtree(s_i): ranges are nested, growing with i
                    int32_t lo_i = 0, hi_i = k_stem;
                    while (lo_i < hi_i) {
                        int32_t mid = (lo_i + hi_i) >> 1;
                        if (x >= st_pos[mid] && x < (int64_t)st_pos[mid] + st_size[mid]) hi_i = mid; else lo_i = mid + 1;
                    }
                    const int32_t i = lo_i;
                    int64_t rel;
                    if (i == 0) rel = x - st_pos[0];
                    else if (x < st_pos[i - 1]) rel = (int64_t)st_size[i - 1] + (x - st_pos[i]);
                    else rel = (int64_t)st_size[i - 1] + (st_pos[i - 1] - st_pos[i]) +
                               (x - ((int64_t)st_pos[i - 1] + st_size[i - 1]));
                    fx = s_base + rel;
                    d.depth[v[u]] = dv[u] - st_depth[i] + depth_q_new + i;
                    if (blocked) { d.tmp[rel] = v[u]; d.sidx[v[u]] = (int32_t)rel; continue; }
                } else {
                    fx = x + xshift;
                }
                d.tmp[fx] = v[u];
                d.node[v[u]].pos = (int32_t)fx;
            }
        }
    }
    NSX_SINGLE {
        s.def_kind = blocked ? 2 : 1; s.def_p = p; s.def_sz = sz;
        s.def_lo = lo; s.def_hi = hi; s.def_S0 = S0; s.def_S1 = S1; s.def_xshift = xshift;
    }
    const int64_t moved = 0;  // (counted by nsx_pivot_flush)
    NSX_SYNC();
    NSX_PH(c, 6, tph);
    // stem: reverse parent pointers, new subtree sizes
    NSX_PAR_FOR(i, 0, k_stem + 1) {
        int32_t v = spath[i];
        nsx_set_parent_mirror(d, v, i == 0 ? p : spath[i - 1]);
        if (i == 0) {
            d.node[v].parent = p;
            d.node[v].pred2 = e * 2 + (d.tail[e] == p ? 0 : 1);
            d.node[v].size = sz;
        } else {
            d.node[v].parent = spath[i - 1];
            d.node[v].pred2 = st_pred2[i - 1] ^ 1;  // same arc, now pointing the other way
            d.node[v].size = sz - st_size[i - 1];
        }
    }
    NSX_SINGLE {
        const uint8_t st_e = d.state[e];
        if (s.spec_e0 < 0) s.spec_e0 = e < d.m ? nsx_special(st_e) : 0;  // no flow update ran: this IS the state before the pivot
        d.state[e] = (uint8_t)(st_e | NSX_ARC_IN_TREE);
        d.state[leave] &= (uint8_t)~NSX_ARC_IN_TREE;
        if (d.csc_state) {
            d.csc_state[d.csc_pos[e]] = d.state[e];
            if (leave < d.m) d.csc_state[d.csc_pos[leave]] = d.state[leave];
        }
        // (the entering arc is in the tree now, the leaving arc outside with the residual bits the flow update gave it)
        c.n_special += (leave < d.m ? nsx_special(d.state[leave]) : 0) - s.spec_e0;
        c.tree_updates++;
        c.sum_subtree += sz;
        if (sz > c.max_subtree) c.max_subtree = sz;
        c.sum_window += moved;
    }
    NSX_SYNC();
    NSX_PH(c, 7, tph);

    // ---- 6. potentials of the re-hung subtree, parent before child ------------------------
    // (the dense array has S in tmp at its final place; `order` catches up in nsx_pivot_flush)
    // (pi_delta_n is block-uniform here: last written before several barriers; the recompute ends with a barrier, so
    // every thread has read it before thread 0 writes it)
    const bool pd_fits = d.pi_delta && s.pi_delta_n == 0 && sz <= NSX_PI_DELTA_CAP;
    if (blocked) nsx_recompute_potentials(d, c.phase, d.tmp, d.sidx, 0, sz, ps, &c.sum_rounds, pd_fits ? d.pi_delta : nullptr);
    else nsx_recompute_potentials(d, c.phase, d.tmp, (const int32_t*)0, s_base, s_base + sz, ps, &c.sum_rounds, pd_fits ? d.pi_delta : nullptr);
    if (d.pi_delta) { NSX_SINGLE { s.pi_delta_n = pd_fits ? sz : -1; } }
    const bool star_record = c.star_on && c.star_valid;  // (block-uniform: written by thread 0 before several barriers)
    if (star_record) {
        NSX_SYNC();  // every thread has read the flags before thread 0 may change them
        if ((int64_t)sz * 4 > d.n) {
            NSX_SINGLE { c.star_valid = 0; }  // a large part of the tree moved: pricing every row afresh is cheaper
        } else {
            const int32_t* seq = blocked ? d.tmp : d.tmp + s_base;
            const int32_t round = c.star_round;
            NSX_PAR_FOR(j, 0, sz) {
                const int32_t v = seq[j];
                NsxRC none; none.key = 0.0; none.arc2 = -1; none.pad = round;
                d.dlist[j] = v; d.dstamp[v] = round; d.rc[v] = none;
                if (c.star_on == 2) d.rc[d.n + v] = none;
                int32_t* info = d.dinfo + 8 * j;
                const int32_t rb = d.row_begin[v], cb = d.col_begin[v];
                info[0] = v; info[1] = rb; info[2] = d.row_begin[v + 1] - rb; info[3] = cb; info[4] = d.col_begin[v + 1] - cb;
            }
            NSX_SINGLE { c.star_nd = sz; }
        }
        NSX_SYNC();
        NSX_SINGLE { nsx_star_close(d, c); }
    }
    NSX_PH(c, 8, tph);

    // ---- 7. reset cadence (simplex.py:1373-1425) -------------------------------------------
    NSX_SINGLE {
        if (c.ftc >= c.ft_limit) {
            c.ftc = 0;
            c.pb = 0;
            c.cl_count = 0; c.cl_since = 0; c.cl_minor = 0;  // pricing_strategy.reset(), simplex.py:1771-1776
            c.resets++;
            c.wepoch = (c.wepoch + 1) & 0xffu;
            if (c.wepoch == 0) c.need_wfill = 1;  // epoch tags wrapped: refill physically
            if (c.star_on == 2) c.star_valid = 0;  // every weight is 1 again: every cached merit is stale
        } else {
            c.ftc++;
        }
    }
    NSX_SYNC();
    NSX_PH(c, 9, tph);
    return 0;
}

// Candidate scan (CandidateListPricing._scan_candidates, simplex_pricing.py:460-505): the <= 100 listed arcs are
// evaluated in parallel, then one thread folds the results in list order exactly like the reference's loop
// (improving arcs: smallest key, first in list order on ties; zero-reduced-cost arcs only while nothing is chosen).
NSX_FN void nsx_cl_scan(const NsxDev& d, NsxCtl& c, int32_t* out_arc2, NsxPivotScratch& s, int allow_zero) {
    const double tol = d.tol;
    NSX_SYNC();
    NSX_PAR_FOR(k, 0, c.cl_count) {
        const int32_t i = c.cl_list[k];
        const uint8_t st = d.state[i];
        int32_t code = 0;
        double key = 0.0;
        if (!(st & NSX_ARC_IN_TREE)) {
            const double rc = NSX_SUB(NSX_ADD(nsx_arc_cost(d, c.phase, i), d.pi[d.tail[i]]), d.pi[d.head[i]]);
            if ((st & NSX_ARC_CAN_FWD) && rc < -tol) { code = 1; key = rc; }
            else if ((st & NSX_ARC_CAN_BWD) && rc > tol) { code = 2; key = -rc; }
            else if (fabs(rc) <= tol) code = (st & NSX_ARC_CAN_FWD) ? 3 : ((st & NSX_ARC_CAN_BWD) ? 4 : 0);
        }
        s.arc2[k] = code;
        s.res[k] = key;
    }
    NSX_SYNC();
    NSX_SINGLE {
        int32_t best = -1;
        double best_rc = 0.0;
        for (int32_t k = 0; k < c.cl_count; ++k) {
            const int32_t code = s.arc2[k], i = c.cl_list[k];
            if (code == 1 || code == 2) {
                if (best < 0 || s.res[k] < best_rc) { best = i * 2 + (code == 2 ? 1 : 0); best_rc = s.res[k]; }
            } else if (allow_zero && best < 0 && (code == 3 || code == 4)) {
                best = i * 2 + (code == 4 ? 1 : 0);
            }
        }
        *out_arc2 = best;
    }
    NSX_SYNC();
}

// Sequential entering rules evaluated by the pivot CTA: the structure-specific rules (specialized_pivots.py:150-450;
// c.row_scan_first = NSX_SPECIAL_ASSIGNMENT / MAX_FLOW / SHORTEST_PATH) and the loop-based Devex block scan
// (simplex_pricing.py:205-269).  Three of the four are NOT reductions - a later arc replaces the incumbent only if its key
// beats it by MORE than the tolerance - so the result depends on the scan order:
//   assignment    (:179-209)  forward arcs:  rc < best - tol                      -> best = rc
//   shortest path (:368-424)  forward arcs with a labelled tail (node_mask): same; backward arcs: -rc < best - tol
//   max flow      (:294-343)  merit = residual * |rc| in either direction, strictly larger wins (first arc on ties)
//   loop Devex    (:205-269)  merit = rc^2 / w  > best + tol (_is_better_candidate in ascending index order), tree cost of
//                             the phase, forward tested before backward, first zero-reduced-cost arc of the block
// The incumbent only ever moves one way, so an arc that the sequential scan would take also beats the incumbent the
// round STARTED with.  Each round all threads evaluate NSX_SP_PER_THREAD arcs apiece against that start value (registers
// only); only if some thread saw such an arc is the round folded in index order - chunk-parallel evaluation into shared
// memory, one thread replaying the reference's loop.  Rounds without a hit (almost all after the first pivots) cost two
// barriers and one coalesced read of the arcs.
#define NSX_SP_CHUNK 1024
#define NSX_SP_PER_THREAD 16
#define NSX_RULE_DEVEX_LOOP 5  // internal rule id next to NSX_SPECIAL_ASSIGNMENT (2) / MAX_FLOW (3) / SHORTEST_PATH (4)

// arc i under `rule`: 0 = not a candidate, 1 / 2 = improving forward / backward with *key, 3 / 4 = zero-reduced-cost
// forward / backward (loop Devex in Phase 1 only)
NSX_FN int32_t nsx_rule_eval(const NsxDev& d, const NsxCtl& c, int32_t rule, int allow_zero, int64_t i, double tol, double* key) {
    const uint8_t st = d.state[i];
    if (st & NSX_ARC_IN_TREE) return 0;
    const int32_t tl = d.tail[i];
    const double rc = NSX_SUB(NSX_ADD(nsx_arc_cost(d, c.phase, i), d.pi[tl]), d.pi[d.head[i]]);
    const bool fv = (st & NSX_ARC_CAN_FWD) && rc < -tol;
    const bool bv = !fv && (st & NSX_ARC_CAN_BWD) && rc > tol;
    if (rule == NSX_RULE_DEVEX_LOOP) {
        if (fv || bv) {
            const uint32_t wraw = d.wgt[i];
            const double w = (wraw >> 24) == c.wepoch ? (double)(wraw & 0xffffffu) : 1.0;
            *key = NSX_DIV(NSX_MUL(rc, rc), w);
            return fv ? 1 : 2;
        }
        if (allow_zero && fabs(rc) <= tol) return (st & NSX_ARC_CAN_FWD) ? 3 : ((st & NSX_ARC_CAN_BWD) ? 4 : 0);
        return 0;
    }
    if (rule == NSX_SPECIAL_MAX_FLOW) {
        if (fv) {
            const double up = d.upper[i];
            *key = NSX_MUL(nsx_isinf(up) ? NSX_INF : NSX_SUB(up, d.flow[i]), fabs(rc));
            return 1;
        }
        if (bv) { *key = NSX_MUL(d.flow[i], fabs(rc)); return 2; }
        return 0;
    }
    if (fv) {
        if (rule == NSX_SPECIAL_ASSIGNMENT || d.node_mask[tl]) { *key = rc; return 1; }
        return 0;
    }
    if (bv && rule == NSX_SPECIAL_SHORTEST_PATH) { *key = -rc; return 2; }
    return 0;
}
NSX_FN bool nsx_rule_beats(int32_t rule, double key, double incumbent, double tol) {
    if (rule == NSX_SPECIAL_MAX_FLOW) return key > incumbent;
    if (rule == NSX_RULE_DEVEX_LOOP) return key > NSX_ADD(incumbent, tol);
    return key < NSX_SUB(incumbent, tol);
}
// in-order fold of [lo, hi) into the incumbent kept in s.sp_key / s.sp_best / s.sp_zero
NSX_FN void nsx_rule_fold(const NsxDev& d, const NsxCtl& c, int32_t rule, int64_t lo, int64_t hi, int allow_zero,
                          NsxPivotScratch& s) {
    const double tol = d.tol;
    for (int64_t base = lo; base < hi; base += NSX_SP_CHUNK) {
        const int64_t cnt = hi - base < NSX_SP_CHUNK ? hi - base : NSX_SP_CHUNK;
        NSX_PAR_FOR(k, 0, cnt) {
            double key = 0.0;
            s.arc2[k] = nsx_rule_eval(d, c, rule, allow_zero, base + k, tol, &key);
            s.res[k] = key;
        }
        NSX_SYNC();
        NSX_SINGLE {
            double best_key = s.sp_key;
            int32_t best = s.sp_best, zero = s.sp_zero;
            for (int64_t k = 0; k < cnt; ++k) {
                const int32_t code = s.arc2[k];
                if (code == 1 || code == 2) {
                    if (nsx_rule_beats(rule, s.res[k], best_key, tol)) { best_key = s.res[k]; best = (int32_t)(base + k) * 2 + (code == 2 ? 1 : 0); }
                } else if (code && zero < 0) {
                    zero = (int32_t)(base + k) * 2 + (code == 4 ? 1 : 0);
                }
            }
            s.sp_key = best_key; s.sp_best = best; s.sp_zero = zero;
        }
        NSX_SYNC();
    }
}
// *out_arc2 = the rule's pick over [lo, hi), *out_zero2 = first zero candidate (both arc*2 + (dir<0), -1 none)
NSX_FN void nsx_rule_scan(const NsxDev& d, NsxCtl& c, int32_t rule, int64_t lo, int64_t hi, int allow_zero,
                          int32_t* out_arc2, int32_t* out_zero2, NsxPivotScratch& s) {
    const double tol = d.tol;
    const bool larger_wins = rule == NSX_SPECIAL_MAX_FLOW || rule == NSX_RULE_DEVEX_LOOP;
    NSX_SYNC();
    NSX_SINGLE { s.sp_key = larger_wins ? -NSX_INF : 0.0; s.sp_best = -1; s.sp_zero = -1; s.sp_any = 0; }
    NSX_SYNC();
    const int64_t span = (int64_t)NSX_NTHREADS * NSX_SP_PER_THREAD;
    int32_t round = 0;
    for (int64_t r0 = lo; r0 < hi; r0 += span) {
        const int64_t r1 = hi - r0 < span ? hi : r0 + span;
        ++round;
        const double incumbent = s.sp_key;  // written before the last barrier
        const bool want_zero = allow_zero && s.sp_zero < 0;
        bool hit = false;
        NSX_PAR_FOR(i, r0, r1) {
            double key = 0.0;
            const int32_t code = nsx_rule_eval(d, c, rule, allow_zero, i, tol, &key);
            if (code == 1 || code == 2) { if (nsx_rule_beats(rule, key, incumbent, tol)) hit = true; }
            else if (code && want_zero) hit = true;
        }
        if (hit) NSX_RAISE_TO(s.sp_any, round);  // several threads may store the same round number
        NSX_SYNC();
        const bool fold = s.sp_any == round;
        NSX_SYNC();  // everybody has read the flag before the next round can raise it again
        if (fold) nsx_rule_fold(d, c, rule, r0, r1, allow_zero, s);
    }
    NSX_SINGLE { *out_arc2 = s.sp_best; if (out_zero2) *out_zero2 = s.sp_zero; c.arcs_priced += hi - lo; }
    NSX_SYNC();
}
NSX_FN void nsx_special_scan(const NsxDev& d, NsxCtl& c, int32_t* out_arc2, NsxPivotScratch& s) {
    nsx_rule_scan(d, c, c.row_scan_first, 0, d.m, 0, out_arc2, (int32_t*)0, s);
}
NSX_FN void nsx_devex_loop_scan(const NsxDev& d, NsxCtl& c, int64_t lo, int64_t hi, int allow_zero,
                                int32_t* out_arc2, int32_t* out_zero2, NsxPivotScratch& s) {
    nsx_rule_scan(d, c, NSX_RULE_DEVEX_LOOP, lo, hi, allow_zero, out_arc2, out_zero2, s);
}

// Block-size adaptation after each pivot (simplex_adaptive.py:98-151). Single thread.
NSX_FN void nsx_adapt_block(NsxCtl& c, int64_t m, int64_t iteration) {
    if (!c.auto_block) return;
    if (iteration - c.tuner_last < 50) return;
    if (c.tuner_total < 10) return;
    double ratio = (double)c.tuner_deg / (double)c.tuner_total;
    if (ratio > 0.30) {
        int64_t nb = (int64_t)NSX_MUL((double)c.bs, 1.5);
        c.bs = nb < m ? nb : m;
    } else if (ratio < 0.10) {
        int64_t nb = (int64_t)NSX_MUL((double)c.bs, 0.75);
        c.bs = nb > 10 ? nb : 10;
    }
    c.tuner_deg = 0;
    c.tuner_total = 0;
    c.tuner_last = iteration;
}

// ------------------------------------------------------------------------------------------
// Per-arc pricing predicates shared by the sweep kernels and the emulation.
// rc = (cost + pi[tail]) - pi[head]  (simplex.py:508-512)
// ------------------------------------------------------------------------------------------
struct NsxCand {      // Dantzig / row-scan: minimum key, lowest index
    double key;       // signed reduced cost (rc forward, -rc backward)
    int32_t arc2;     // arc*2 + (dir<0); -1 = none
    int32_t zero2;    // first zero-reduced-cost candidate arc*2+(dir<0); INT32_MAX = none
};
struct NsxDevexCand { // Devex block: first arg-max of rc^2/w per direction, zero candidates
    double fm, bm;
    int32_t fi, bi;   // -1 = none
    int32_t fz, bz;   // INT32_MAX = none
};

NSX_FN void nsx_cand_init(NsxCand& k) { k.key = 0.0; k.arc2 = -1; k.zero2 = 0x7fffffff; }
NSX_FN void nsx_devex_init(NsxDevexCand& k) {
    k.fm = -NSX_INF; k.bm = -NSX_INF; k.fi = -1; k.bi = -1; k.fz = 0x7fffffff; k.bz = 0x7fffffff;
}
// merge b (covering higher arc indices or an unordered partition) into a: lowest index wins ties
NSX_FN void nsx_cand_merge(NsxCand& a, const NsxCand& b) {
    if (b.arc2 >= 0 && (a.arc2 < 0 || b.key < a.key || (b.key == a.key && b.arc2 < a.arc2))) {
        a.key = b.key; a.arc2 = b.arc2;
    }
    if (b.zero2 < a.zero2) a.zero2 = b.zero2;
}
NSX_FN void nsx_devex_merge(NsxDevexCand& a, const NsxDevexCand& b) {
    if (b.fi >= 0 && (a.fi < 0 || b.fm > a.fm || (b.fm == a.fm && b.fi < a.fi))) { a.fm = b.fm; a.fi = b.fi; }
    if (b.bi >= 0 && (a.bi < 0 || b.bm > a.bm || (b.bm == a.bm && b.bi < a.bi))) { a.bm = b.bm; a.bi = b.bi; }
    if (b.fz < a.fz) a.fz = b.fz;
    if (b.bz < a.bz) a.bz = b.bz;
}

// DantzigPricing.select_entering_arc body for one arc (simplex_pricing.py:110-135)
// ---- star pricing: one arc's candidate, and the order of the row cache (same as nsx_dantzig_improving / nsx_cand_merge) ----
// returns arc*2 + (backward) with its key, or -1 when the arc is not an improving candidate
NSX_FN int32_t nsx_star_candidate(int64_t a, uint32_t st, double rc, double tol, double* key) {
    if (st & NSX_ARC_IN_TREE) return -1;
    if ((st & NSX_ARC_CAN_FWD) && rc < -tol) { *key = rc; return (int32_t)(a * 2); }
    if ((st & NSX_ARC_CAN_BWD) && rc > tol) { *key = -rc; return (int32_t)(a * 2 + 1); }
    return -1;
}
NSX_FN bool nsx_rc_better(double key, int32_t arc2, const NsxRC& cur) {
    return cur.arc2 < 0 || key < cur.key || (key == cur.key && arc2 < cur.arc2);
}
// pricing cost of a real arc from its perturbed Phase-2 cost (nsx_arc_cost without the array access)
NSX_FN double nsx_phase_cost(int32_t phase, double pert, int64_t a) {
    return phase == 1 ? NSX_SUB(NSX_SUB(pert, 1.0), NSX_MUL(1e-6, (double)a)) : pert;
}

NSX_FN void nsx_price_dantzig(NsxCand& k, int32_t i, uint8_t st, double rc, double tol) {
    if (st & NSX_ARC_IN_TREE) return;
    if ((st & NSX_ARC_CAN_FWD) && rc < -tol) {
        if (k.arc2 < 0 || rc < k.key || (rc == k.key && i * 2 < k.arc2)) { k.key = rc; k.arc2 = i * 2; }
    } else if ((st & NSX_ARC_CAN_BWD) && rc > tol) {
        double nk = -rc;
        if (k.arc2 < 0 || nk < k.key || (nk == k.key && i * 2 + 1 < k.arc2)) { k.key = nk; k.arc2 = i * 2 + 1; }
    } else if (fabs(rc) <= tol) {
        if (st & NSX_ARC_CAN_FWD) { if (i * 2 < k.zero2) k.zero2 = i * 2; }
        else if (st & NSX_ARC_CAN_BWD) { if (i * 2 + 1 < k.zero2) k.zero2 = i * 2 + 1; }
    }
}

// NetworkSimplex._select_entering_arc_vectorized body for one arc (simplex.py:571-615)
NSX_FN void nsx_price_devex(NsxDevexCand& k, int32_t i, uint8_t st, double rc, uint32_t wraw,
                            uint32_t wepoch, double tol) {
    if (st & NSX_ARC_IN_TREE) return;
    bool fv = (st & NSX_ARC_CAN_FWD) && rc < -tol;
    bool bv = (st & NSX_ARC_CAN_BWD) && rc > tol;
    if (fv || bv) {
        double w = (wraw >> 24) == wepoch ? (double)(wraw & 0xffffffu) : 1.0;
        double merit = NSX_DIV(NSX_MUL(rc, rc), w);
        if (fv) { if (k.fi < 0 || merit > k.fm || (merit == k.fm && i < k.fi)) { k.fm = merit; k.fi = i; } }
        else    { if (k.bi < 0 || merit > k.bm || (merit == k.bm && i < k.bi)) { k.bm = merit; k.bi = i; } }
    } else if (fabs(rc) <= tol) {
        if ((st & NSX_ARC_CAN_FWD) && i < k.fz) k.fz = i;
        if ((st & NSX_ARC_CAN_BWD) && i < k.bz) k.bz = i;
    }
}

// Final decision of one Devex block from its merged candidate (simplex.py:589-617).
// Returns arc2 (arc*2 + (dir<0)) or -1; *merit_pos = 1 when the weight must be refreshed.
NSX_FN int32_t nsx_devex_decide(const NsxDevexCand& k, int allow_zero, int32_t* merit_pos) {
    *merit_pos = 0;
    double fm = k.fi >= 0 ? k.fm : -NSX_INF, bm = k.bi >= 0 ? k.bm : -NSX_INF;
    if (fm > bm) {
        if (k.fi >= 0) { *merit_pos = fm > 0.0; return k.fi * 2; }
    } else {
        if (k.bi >= 0) { *merit_pos = bm > 0.0; return k.bi * 2 + 1; }
    }
    if (allow_zero) {
        if (k.fz != 0x7fffffff) return k.fz * 2;
        if (k.bz != 0x7fffffff) return k.bz * 2 + 1;
    }
    return -1;
}

// ------------------------------------------------------------------------------------------
// Initial state (simplex.py:619-728): flows 0, every node hangs off the root by one artificial
// arc whose direction / flow follow the supply sign.  Element-wise, any launch shape.
// ------------------------------------------------------------------------------------------
NSX_FN void nsx_init_real_arc(const NsxDev& d, int64_t i) {
    d.flow[i] = 0.0;
    d.state[i] = nsx_bounds_bits(0.0, d.upper[i], d.tol);
    if (d.wgt) d.wgt[i] = 1u;  // epoch 0, weight 1
}
// Writes node v >= 1 and its artificial arc m + v - 1.
NSX_FN void nsx_init_node(const NsxDev& d, int32_t v, double supply) {
    if (v == 0) {
        NsxNode r; r.parent = 0; r.pred2 = -1; r.pos = 0; r.size = d.n;
        d.node[0] = r; d.depth[0] = 0; d.pi[0] = 0.0; d.order[0] = 0;
        return;
    }
    int64_t a = d.m + (v - 1);
    double f, up; int32_t tl, hd;
    if (fabs(supply) <= d.tol) { tl = 0; hd = v; up = NSX_INF; f = 0.0; }
    else if (supply > 0) { tl = v; hd = 0; up = supply; f = supply; }
    else { tl = 0; hd = v; up = -supply; f = -supply; }
    d.atail[v - 1] = tl; d.ahead[v - 1] = hd; d.aupper[v - 1] = up;
    d.flow[a] = f;
    d.state[a] = (uint8_t)(NSX_ARC_IN_TREE | nsx_bounds_bits(f, up, d.tol));
    NsxNode r; r.parent = 0; r.pred2 = (int32_t)(a * 2 + (tl == 0 ? 0 : 1)); r.pos = v; r.size = 1;
    d.node[v] = r; d.depth[v] = 1; d.order[v] = v;
}

// Warm start (nsx_solve_warm): arc a of [0, ma) from the caller's tree flags and flows.  Artificial arcs keep the
// geometry of the cold start (direction / capacity by the sign of the supply, simplex.py:645-698); node records, depth and
// the preorder array are laid out on the host (csrc/nsx_warm.h) and copied in.  Returns 1 for an artificial arc with flow.
NSX_FN int nsx_init_arc_warm(const NsxDev& d, int64_t a, const double* supply, const uint8_t* in_tree) {
    const uint8_t tree = in_tree[a] ? (uint8_t)(NSX_ARC_IN_TREE | NSX_ARC_STALE) : (uint8_t)0;
    const double f = d.flow[a];
    if (a < d.m) {
        d.state[a] = (uint8_t)(tree | nsx_bounds_bits(f, d.upper[a], d.tol));
        if (d.wgt) d.wgt[a] = 1u;
        return 0;
    }
    const int32_t v = (int32_t)(a - d.m) + 1;
    const double sp = supply[v];
    double up; int32_t tl, hd;
    if (fabs(sp) <= d.tol) { tl = 0; hd = v; up = NSX_INF; }
    else if (sp > 0) { tl = v; hd = 0; up = sp; }
    else { tl = 0; hd = v; up = -sp; }
    d.atail[v - 1] = tl; d.ahead[v - 1] = hd; d.aupper[v - 1] = up;
    d.state[a] = (uint8_t)(tree | nsx_bounds_bits(f, up, d.tol));
    return f > d.tol ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// Driver state machine (single thread of the pivot CTA).  Restates the control flow of
// NetworkSimplex.solve / _run_simplex_iterations / _find_entering_arc
// (simplex.py:1058-1075, 1109-1160, 1534-1701) and DevexPricing's block loop
// (simplex_pricing.py:325-357).
// ------------------------------------------------------------------------------------------
// *_ZERO commands look for zero-reduced-cost candidates only (Phase 1, after the improving sweep of
// the same range found nothing): the hot sweeps then carry no zero bookkeeping.
// NSX_CMD_TOPK refreshes the candidate list: the NSX_CL_SIZE improving arcs of largest |rc| (ties: larger index).
enum { NSX_CMD_EXIT = 0, NSX_CMD_DANTZIG = 1, NSX_CMD_DEVEX = 2, NSX_CMD_DANTZIG_ZERO = 3, NSX_CMD_DEVEX_ZERO = 4,
       NSX_CMD_TOPK = 5,
       // star pricing: NSX_CMD_STAR_BUILD prices every row afresh (cache invalid), NSX_CMD_STAR brings the cache up to
       // date after one pivot: lo = nodes in dlist, hi = round stamp, excluded = row of the entering arc (-1 none)
       NSX_CMD_STAR = 6, NSX_CMD_STAR_BUILD = 7 };
enum { NSX_ST_ROWSCAN = 1, NSX_ST_DANTZIG = 2, NSX_ST_DEVEX = 3, NSX_ST_DANTZIG_ZERO = 4, NSX_ST_DEVEX_ZERO = 5,
       // candidate list (CandidateListPricing.select_entering_arc, simplex_pricing.py:418-458): quick scan of the
       // list, scan after the (optional) periodic refresh, forced refresh, scan after the forced refresh
       NSX_ST_CL_QUICK = 6, NSX_ST_CL_REFRESH = 7, NSX_ST_CL_MAIN = 8, NSX_ST_CL_FORCED = 9, NSX_ST_CL_LAST = 10,
       NSX_ST_SPECIAL = 11,    // structure-specific rule (assignment / max flow / shortest path) before the configured one
       NSX_ST_DEVEX_LOOP = 12 };  // loop-based Devex: one block scanned by the pivot CTA per step
enum { NSX_ACT_SWEEP = 0, NSX_ACT_PIVOT = 1, NSX_ACT_PHASE_END = 2, NSX_ACT_EXIT = 3, NSX_ACT_RECOMPUTE = 4,
       NSX_ACT_CL_SCAN = 5,    // CL_SCAN: the pivot CTA evaluates the <= 100 listed arcs itself, no sweep
       NSX_ACT_SPECIAL_SCAN = 6,    // the pivot CTA runs nsx_special_scan over all arcs
       NSX_ACT_BLOCK_SCAN = 7 };    // the pivot CTA runs nsx_devex_loop_scan over the block [cmd.lo, cmd.hi)

struct NsxCmd {       // what every CTA does next
    int32_t kind;
    int32_t phase;    // selects the pricing cost for NSX_CMD_DANTZIG
    int64_t lo, hi;   // arc range
    int32_t excluded; // Devex: arc skipped (last bound-flip arc), -1 none
    uint32_t wepoch;
    int32_t reverse;  // tiles visited in descending order (alternates per sweep: the tail of the
                      // previous sweep is still in L2 / shared memory when the next one starts)
    int32_t pad[3];
};
struct NsxAction { int32_t kind, arc, dir, want_weight; };
struct NsxDrv {       // driver scalars (kept beside NsxCtl)
    int32_t stage, final_check;
    int64_t bc, blocks_left, budget;
};
struct NsxLoopShared {
    NsxCmd cmd;
    NsxAction act;
    NsxCand dz;
    NsxDevexCand dx;
    NsxDrv drv;
    int32_t rc;
    int32_t cl_arc2;  // result of the last candidate scan (arc*2 + (dir<0)), -1 none
    int32_t zero2;    // loop-based Devex: first zero-reduced-cost candidate of the scanned block, -1 none
};

NSX_FN void nsx_drv_devex_cmd(NsxCtl& c, int64_t m, NsxCmd& cmd) {
    int64_t st = c.pb * c.bs;
    if (st >= m) { c.pb = 0; st = 0; }
    int64_t en = st + c.bs < m ? st + c.bs : m;
    cmd.kind = NSX_CMD_DEVEX; cmd.phase = c.phase; cmd.lo = st; cmd.hi = en;
    cmd.excluded = c.last_deg; cmd.wepoch = c.wepoch; cmd.reverse ^= 1;
    cmd.pad[0] = 0;
    if (c.star_on == 2 && c.bs >= m) {
        // one block = all arcs: the search is an arg-max over everything, kept in the row cache (star pricing).  pad[0] = 1
        // marks the Devex flavour, pad[1] = arc left out (last_degenerate_arc), pad[2] = arc to put back in (left out by
        // the previous command, not any more)
        const int32_t back = (c.star_excl_prev >= 0 && c.star_excl_prev != c.last_deg) ? c.star_excl_prev : -1;
        cmd.pad[0] = 1; cmd.pad[1] = c.last_deg; cmd.pad[2] = back;
        if (c.star_valid && c.star_pending <= 1) {
            cmd.kind = NSX_CMD_STAR;
            cmd.lo = c.star_pending ? c.star_ne : 0; cmd.hi = c.star_round;
            cmd.excluded = c.star_pending ? c.star_extra : -1;
        } else {
            cmd.kind = NSX_CMD_STAR_BUILD;
        }
    }
}
NSX_FN void nsx_drv_devex_begin(NsxCtl& c, NsxDrv& v, int64_t m, NsxCmd& cmd) {
    v.stage = NSX_ST_DEVEX;
    v.bc = (m + c.bs - 1) / c.bs;
    if (v.bc < 1) v.bc = 1;
    v.blocks_left = v.bc;
    nsx_drv_devex_cmd(c, m, cmd);
}
// ---- candidate list (CandidateListPricing.select_entering_arc, simplex_pricing.py:418-458) ----
NSX_FN void nsx_drv_cl_refresh_cmd(NsxCtl& c, int64_t m, NsxCmd& cmd) {
    cmd.kind = NSX_CMD_TOPK; cmd.phase = c.phase; cmd.lo = 0; cmd.hi = m;
    cmd.excluded = -1; cmd.wepoch = c.wepoch; cmd.reverse ^= 1;
}
// "time for a refresh" (:434-441): count a major iteration, refresh when due, then scan
NSX_FN void nsx_drv_cl_major(NsxCtl& c, NsxDrv& v, int64_t m, NsxCmd& cmd, NsxAction& act) {
    c.cl_since++;
    c.cl_minor = 0;
    if (c.cl_since >= NSX_CL_REFRESH || c.cl_count == 0) {
        v.stage = NSX_ST_CL_REFRESH; act.kind = NSX_ACT_SWEEP;
        nsx_drv_cl_refresh_cmd(c, m, cmd);
    } else {
        v.stage = NSX_ST_CL_MAIN; act.kind = NSX_ACT_CL_SCAN;
    }
}
NSX_FN void nsx_drv_cl_begin(NsxCtl& c, NsxDrv& v, int64_t m, NsxCmd& cmd, NsxAction& act) {
    if (c.cl_count > 0 && c.cl_minor < NSX_CL_MINOR) { v.stage = NSX_ST_CL_QUICK; act.kind = NSX_ACT_CL_SCAN; return; }
    nsx_drv_cl_major(c, v, m, cmd, act);
}

// ---- loop-based Devex (simplex_pricing.py:205-269): the block loop of select_entering_arc, one block per step ----
NSX_FN void nsx_drv_devex_loop_block(NsxCtl& c, int64_t m, NsxCmd& cmd, NsxAction& act) {
    int64_t st = c.pb * c.bs;
    if (st >= m) { c.pb = 0; st = 0; }
    cmd.lo = st; cmd.hi = st + c.bs < m ? st + c.bs : m; cmd.phase = c.phase;
    act.kind = NSX_ACT_BLOCK_SCAN;
}
NSX_FN void nsx_drv_devex_loop_begin(NsxCtl& c, NsxDrv& v, int64_t m, NsxCmd& cmd, NsxAction& act) {
    v.stage = NSX_ST_DEVEX_LOOP;
    v.bc = (m + c.bs - 1) / c.bs;
    if (v.bc < 1) v.bc = 1;
    v.blocks_left = v.bc;
    nsx_drv_devex_loop_block(c, m, cmd, act);
}

// full-range Dantzig pricing: a sweep of all arcs, or - star pricing - an update / rebuild of the row cache
NSX_FN void nsx_drv_dantzig_cmd(NsxCtl& c, int64_t m, NsxCmd& cmd) {
    cmd.kind = NSX_CMD_DANTZIG; cmd.phase = c.phase; cmd.lo = 0; cmd.hi = m;
    cmd.excluded = -1; cmd.wepoch = c.wepoch; cmd.reverse ^= 1;
    cmd.pad[0] = 0;
    cmd.pad[1] = c.n_special == 0;  // the sweep may skip the state bytes (nsx_price_tile_dz)
    if (c.star_on != 1) return;
    if (c.star_valid && c.star_pending <= 1) {
        cmd.kind = NSX_CMD_STAR;
        cmd.lo = c.star_pending ? c.star_ne : 0; cmd.hi = c.star_round;
        cmd.excluded = c.star_pending ? c.star_extra : -1;
    } else {
        cmd.kind = NSX_CMD_STAR_BUILD;
    }
}
// top of an iteration: first sweep command, or phase end when the budget is used up
NSX_FN void nsx_drv_begin(NsxCtl& c, NsxDrv& v, int64_t m, NsxCmd& cmd, NsxAction& act) {
    if (!v.final_check && c.it >= v.budget) { act.kind = NSX_ACT_PHASE_END; return; }
    if (c.row_scan_first >= NSX_SPECIAL_ASSIGNMENT) { v.stage = NSX_ST_SPECIAL; act.kind = NSX_ACT_SPECIAL_SCAN; return; }
    act.kind = NSX_ACT_SWEEP;
    if (c.row_scan_first || c.pricing == NSX_PRICING_DANTZIG) {
        v.stage = c.row_scan_first ? NSX_ST_ROWSCAN : NSX_ST_DANTZIG;
        nsx_drv_dantzig_cmd(c, m, cmd);
    } else if (c.pricing == NSX_PRICING_CANDIDATE_LIST) {
        nsx_drv_cl_begin(c, v, m, cmd, act);
    } else if (c.pricing == NSX_PRICING_DEVEX_LOOP) {
        nsx_drv_devex_loop_begin(c, v, m, cmd, act);
    } else {
        nsx_drv_devex_begin(c, v, m, cmd);
    }
}
NSX_FN void nsx_drv_choose(NsxCtl& c, NsxDrv& v, NsxAction& act, int32_t arc2, int32_t want_weight,
                           int32_t* trace) {
    if (v.final_check) {  // simplex.py:1678-1699: an entering arc still exists at the limit
        c.status = NSX_STATUS_ITERATION_LIMIT;
        act.kind = NSX_ACT_EXIT;
        return;
    }
    act.kind = NSX_ACT_PIVOT; act.arc = arc2 >> 1; act.dir = (arc2 & 1) ? -1 : 1;
    act.want_weight = want_weight;
    if (trace && c.trace_len < c.trace_cap) trace[c.trace_len] = arc2;
    c.trace_len++;
}
NSX_FN void nsx_drv_none(NsxCtl& c, NsxDrv& v, NsxAction& act) {
    if (v.final_check) { c.status = NSX_STATUS_OPTIMAL; act.kind = NSX_ACT_EXIT; }
    else act.kind = NSX_ACT_PHASE_END;
}
// Devex: the current block yielded nothing; move to the next block or give up
NSX_FN void nsx_drv_devex_next(NsxCtl& c, NsxDrv& v, int64_t m, NsxCmd& cmd, NsxAction& act) {
    c.pb = (c.pb + 1) % v.bc;
    v.blocks_left--;
    if (v.blocks_left == 0) { nsx_drv_none(c, v, act); return; }
    v.stage = NSX_ST_DEVEX;
    act.kind = NSX_ACT_SWEEP;
    nsx_drv_devex_cmd(c, m, cmd);
}
// result of a candidate scan (NSX_ACT_CL_SCAN)
NSX_FN void nsx_drv_on_scan(NsxCtl& c, NsxDrv& v, int64_t m, int32_t arc2, NsxCmd& cmd, NsxAction& act, int32_t* trace) {
    if (v.stage == NSX_ST_CL_QUICK) {
        if (arc2 >= 0) { c.cl_minor++; nsx_drv_choose(c, v, act, arc2, 0, trace); return; }
        nsx_drv_cl_major(c, v, m, cmd, act);
    } else if (v.stage == NSX_ST_CL_MAIN) {
        if (arc2 >= 0) { nsx_drv_choose(c, v, act, arc2, 0, trace); return; }
        if (c.cl_since > 0) {  // "if still no improving arc, force full refresh" (:448-452)
            v.stage = NSX_ST_CL_FORCED; act.kind = NSX_ACT_SWEEP;
            nsx_drv_cl_refresh_cmd(c, m, cmd);
            return;
        }
        nsx_drv_none(c, v, act);
    } else {  // NSX_ST_CL_LAST
        if (arc2 >= 0) { nsx_drv_choose(c, v, act, arc2, 0, trace); return; }
        nsx_drv_none(c, v, act);
    }
}
// after nsx_special_scan: pivot on its arc, or hand over to the configured strategy (simplex.py:1061-1075)
NSX_FN void nsx_drv_on_special(NsxCtl& c, NsxDrv& v, int64_t m, int32_t arc2, NsxCmd& cmd, NsxAction& act, int32_t* trace) {
    if (arc2 >= 0) { nsx_drv_choose(c, v, act, arc2, 0, trace); return; }
    act.kind = NSX_ACT_SWEEP;
    if (c.pricing == NSX_PRICING_DANTZIG) {
        v.stage = NSX_ST_DANTZIG;
        nsx_drv_dantzig_cmd(c, m, cmd);
    } else if (c.pricing == NSX_PRICING_CANDIDATE_LIST) {
        nsx_drv_cl_begin(c, v, m, cmd, act);
    } else if (c.pricing == NSX_PRICING_DEVEX_LOOP) {
        nsx_drv_devex_loop_begin(c, v, m, cmd, act);
    } else {
        nsx_drv_devex_begin(c, v, m, cmd);
    }
}
// after nsx_devex_loop_scan of one block (simplex_pricing.py:256-266)
NSX_FN void nsx_drv_on_block(NsxCtl& c, NsxDrv& v, int64_t m, int32_t arc2, int32_t zero2, NsxCmd& cmd, NsxAction& act,
                             int32_t* trace) {
    if (arc2 >= 0) { nsx_drv_choose(c, v, act, arc2, 1, trace); return; }  // weight updated for the selected arc only
    c.pb = (c.pb + 1) % v.bc;
    if (zero2 >= 0) { nsx_drv_choose(c, v, act, zero2, 0, trace); return; }
    v.blocks_left--;
    if (v.blocks_left == 0) { nsx_drv_none(c, v, act); return; }
    nsx_drv_devex_loop_block(c, m, cmd, act);
}
NSX_FN void nsx_drv_on_result(NsxCtl& c, NsxDrv& v, int64_t m, const NsxCand& dz,
                              const NsxDevexCand& dx, NsxCmd& cmd, NsxAction& act, int32_t* trace) {
    const int allow_zero = (c.phase == 1) && !v.final_check;
    if (cmd.kind == NSX_CMD_STAR || cmd.kind == NSX_CMD_STAR_BUILD) {  // the row cache is up to date again
        c.arcs_priced += c.star_evaluated;
        if (cmd.kind == NSX_CMD_STAR) c.star_updates++; else c.star_builds++;
        c.star_valid = 1; c.star_pending = 0;
        cmd.lo = 0; cmd.hi = m;  // (a zero-candidate pass may follow: it sweeps the arc range)
        cmd.excluded = cmd.pad[0] ? cmd.pad[1] : -1;
        if (cmd.pad[0]) c.star_excl_prev = cmd.pad[1];
    } else {
        c.arcs_priced += cmd.hi - cmd.lo;
    }
    c.sweeps++;
    if (v.stage == NSX_ST_CL_REFRESH || v.stage == NSX_ST_CL_FORCED) {  // the list has just been refreshed
        c.cl_since = 0;
        v.stage = v.stage == NSX_ST_CL_REFRESH ? NSX_ST_CL_MAIN : NSX_ST_CL_LAST;
        act.kind = NSX_ACT_CL_SCAN;
        return;
    }
    if (v.stage == NSX_ST_ROWSCAN || v.stage == NSX_ST_DANTZIG) {
        if (dz.arc2 >= 0) { nsx_drv_choose(c, v, act, dz.arc2, 0, trace); return; }
        if (v.stage == NSX_ST_ROWSCAN && c.pricing == NSX_PRICING_CANDIDATE_LIST) {
            nsx_drv_cl_begin(c, v, m, cmd, act);  // fall through to the configured strategy (simplex.py:1066-1075)
            return;
        }
        if (v.stage == NSX_ST_ROWSCAN && c.pricing == NSX_PRICING_DEVEX_LOOP) {
            nsx_drv_devex_loop_begin(c, v, m, cmd, act);  // fall through to the configured strategy (simplex.py:1066-1075)
            return;
        }
        if (v.stage == NSX_ST_ROWSCAN && c.pricing == NSX_PRICING_DEVEX) {
            act.kind = NSX_ACT_SWEEP;  // fall through to the configured strategy (simplex.py:1066-1075)
            nsx_drv_devex_begin(c, v, m, cmd);
            return;
        }
        if (allow_zero) {  // simplex_pricing.py:132-135: zero-reduced-cost candidates, second pass
            v.stage = NSX_ST_DANTZIG_ZERO;
            act.kind = NSX_ACT_SWEEP;
            cmd.kind = NSX_CMD_DANTZIG_ZERO; cmd.reverse ^= 1;
            return;
        }
        nsx_drv_none(c, v, act);
    } else if (v.stage == NSX_ST_DANTZIG_ZERO) {
        if (dz.zero2 != 0x7fffffff) { nsx_drv_choose(c, v, act, dz.zero2, 0, trace); return; }
        nsx_drv_none(c, v, act);
    } else if (v.stage == NSX_ST_DEVEX) {
        int32_t mp = 0;
        int32_t a2 = nsx_devex_decide(dx, 0, &mp);
        if (a2 >= 0) { c.last_deg = -1; nsx_drv_choose(c, v, act, a2, mp, trace); return; }
        if (allow_zero) {  // simplex.py:603-615: zero candidates of the same block, second pass
            v.stage = NSX_ST_DEVEX_ZERO;
            act.kind = NSX_ACT_SWEEP;
            cmd.kind = NSX_CMD_DEVEX_ZERO; cmd.reverse ^= 1;
            return;
        }
        nsx_drv_devex_next(c, v, m, cmd, act);
    } else {  // NSX_ST_DEVEX_ZERO
        int32_t a2 = dx.fz != 0x7fffffff ? dx.fz * 2 : (dx.bz != 0x7fffffff ? dx.bz * 2 + 1 : -1);
        if (a2 >= 0) { c.last_deg = -1; nsx_drv_choose(c, v, act, a2, 0, trace); return; }
        nsx_drv_devex_next(c, v, m, cmd, act);
    }
}
NSX_FN void nsx_drv_after_pivot(NsxCtl& c, NsxDrv& v, int64_t m, int32_t rc, NsxCmd& cmd, NsxAction& act) {
    if (rc == 3) {
        c.status = NSX_STATUS_UNBOUNDED;
        if (c.phase == 2) c.total += c.it;
        act.kind = NSX_ACT_EXIT;
        return;
    }
    c.it++;
    nsx_adapt_block(c, m, (c.phase == 1 ? 0 : c.total) + c.it);
    if (c.phase == 1 && c.art_with_flow == 0) { act.kind = NSX_ACT_PHASE_END; return; }  // simplex.py:1157
    nsx_drv_begin(c, v, m, cmd, act);
}
NSX_FN void nsx_drv_phase_end(NsxCtl& c, NsxDrv& v, NsxAction& act) {
    if (c.phase == 1) {
        c.total = c.it;
        c.phase1_iterations = c.it;
        c.art_after_p1 = c.art_with_flow;
        if (c.art_with_flow > 0 || c.unbalanced) {  // simplex.py:1571-1624
            c.status = c.total >= c.maxit ? NSX_STATUS_ITERATION_LIMIT_P1 : NSX_STATUS_INFEASIBLE;
            act.kind = NSX_ACT_EXIT;
            return;
        }
        c.phase = 2;
        v.budget = c.maxit - c.total > 0 ? c.maxit - c.total : 0;
        c.it = 0;
        act.kind = NSX_ACT_RECOMPUTE;  // simplex.py:1632-1633: Phase-2 costs, potentials rebuilt
    } else {
        c.total += c.it;
        c.it = 0;
        if (c.total >= c.maxit) { v.final_check = 1; act.kind = NSX_ACT_RECOMPUTE; }
        else { c.status = NSX_STATUS_OPTIMAL; act.kind = NSX_ACT_EXIT; }
    }
}

// Conservation check after Phase 1 of a warm solve (simplex.py:1575-1598): any node other than the root whose balance is
// off by more than tol makes the run "infeasible" (see the clamp note in nsx_pivot).
NSX_FN void nsx_check_conservation(const NsxDev& d, NsxCtl& c) {
    const double tol = d.tol;
    NSX_SYNC();
    NSX_PAR_FOR(v, 1, d.n) { if (fabs(d.imbalance[v]) > tol) NSX_RAISE_TO(c.unbalanced, 1); }
    NSX_SYNC();
}

// The resident loop of the pivot CTA.  `Sweep::run(cmd, dz, dx)` prices the arc range of `cmd`
// (grid-wide on the device, serially in the emulation) and leaves the merged candidates in
// dz / dx, visible to all threads of this CTA on return.  `Sweep::finish()` releases workers.
// NsxCtl::n_special of the initial state (cold or warm), once per solve
NSX_FN void nsx_count_special(const NsxDev& d, NsxCtl& c, NsxPivotScratch& s) {
    NSX_SINGLE { s.spec_e0 = 0; }
    NSX_SYNC();
    int32_t cnt = 0;
#if NSX_ON_DEVICE
    // sixteen state bytes per load while the pointer allows it
    const int64_t m16 = ((reinterpret_cast<uintptr_t>(d.state) & 15) == 0) ? (d.m >> 4) : 0;
    const uint4* s16 = reinterpret_cast<const uint4*>(d.state);
    for (int64_t i = threadIdx.x; i < m16; i += blockDim.x) {
        const uint4 v = s16[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int b = 0; b < 4; ++b) cnt += nsx_special((uint8_t)(w[k] >> (8 * b)));
        }
    }
    NSX_PAR_FOR(i, m16 << 4, d.m) { cnt += nsx_special(d.state[i]); }
#else
    NSX_PAR_FOR(i, 0, d.m) { cnt += nsx_special(d.state[i]); }
#endif
    if (cnt) NSX_ATOMIC_ADD_I32(&s.spec_e0, cnt);
    NSX_SYNC();
    NSX_SINGLE { c.n_special = s.spec_e0; }
    NSX_SYNC();
}

template <bool BLK, class Sweep>
NSX_FN void nsx_solve_loop(const NsxDev& d, NsxCtl& c, NsxLoopShared& L, NsxPivotScratch& s,
                           NsxPotScratch& ps, int32_t* trace, Sweep& sweep) {
    NSX_SINGLE { s.def_kind = 0; s.pi_delta_n = -1; }
    if (BLK && d.blk) nsx_blk_init(d, *d.blk);
    // (block-uniform; thread 0 writes the count only behind the first barrier inside, i.e. after every thread has tested it)
    if (c.n_special < 0) nsx_count_special(d, c, s);
    NSX_SYNC();
    // Phase-1 costs on the initial star; a warm start may begin in Phase 2 (no artificial arc in its tree)
    nsx_recompute_all_potentials<BLK>(d, c.phase, ps);
    NSX_SINGLE {
        L.drv.stage = 0; L.drv.final_check = 0; L.drv.bc = 1; L.drv.blocks_left = 0;
        L.drv.budget = c.maxit;
        L.cmd.reverse = 0;
        nsx_drv_begin(c, L.drv, d.m, L.cmd, L.act);
    }
    for (;;) {
        NSX_SYNC();
        const int32_t kind = L.act.kind;
        sweep.alive();
        NSX_SYNC();
        if (kind == NSX_ACT_SWEEP) {
            sweep.run(L.cmd, L.dz, L.dx, c, [&]() { nsx_pivot_flush<BLK>(d, c, s); });
            NSX_SYNC();
            if (c.fault) {  // (block-uniform: written before the barrier) a worker / peer GPU never answered
                sweep.finish();
                break;
            }
            NSX_SINGLE { nsx_drv_on_result(c, L.drv, d.m, L.dz, L.dx, L.cmd, L.act, trace); }
        } else if (kind == NSX_ACT_CL_SCAN) {
            nsx_cl_scan(d, c, &L.cl_arc2, s, (c.phase == 1) && !L.drv.final_check);
            NSX_SINGLE { nsx_drv_on_scan(c, L.drv, d.m, L.cl_arc2, L.cmd, L.act, trace); }
        } else if (kind == NSX_ACT_BLOCK_SCAN) {
            nsx_devex_loop_scan(d, c, L.cmd.lo, L.cmd.hi, (c.phase == 1) && !L.drv.final_check, &L.cl_arc2, &L.zero2, s);
            NSX_SINGLE { nsx_drv_on_block(c, L.drv, d.m, L.cl_arc2, L.zero2, L.cmd, L.act, trace); }
        } else if (kind == NSX_ACT_SPECIAL_SCAN) {
            nsx_special_scan(d, c, &L.cl_arc2, s);
            NSX_SINGLE { nsx_drv_on_special(c, L.drv, d.m, L.cl_arc2, L.cmd, L.act, trace); }
        } else if (kind == NSX_ACT_PIVOT) {
            int32_t rc = nsx_pivot<BLK>(d, c, s, ps, L.act.arc, L.act.dir, L.act.want_weight);
            if (c.need_wfill) {  // Devex epoch tags wrapped: physically reset the weights
                NSX_SYNC();
                NSX_PAR_FOR(i, 0, d.m) { if (d.wgt) d.wgt[i] = 1u; if (d.csc_wgt) d.csc_wgt[i] = 1u; }
                NSX_SYNC();
                NSX_SINGLE { c.need_wfill = 0; }
            }
            NSX_SYNC();
            NSX_SINGLE { nsx_drv_after_pivot(c, L.drv, d.m, rc, L.cmd, L.act); }
        } else if (kind == NSX_ACT_PHASE_END) {
            // (d.imbalance first: without it nobody but thread 0 may touch c.phase here, thread 0 is about to change it)
            if (d.imbalance && c.phase == 1) nsx_check_conservation(d, c);
            NSX_SINGLE { nsx_drv_phase_end(c, L.drv, L.act); }
        } else if (kind == NSX_ACT_RECOMPUTE) {
            nsx_pivot_flush<BLK>(d, c, s);
            if (!L.drv.final_check) {
                nsx_recompute_all_potentials<BLK>(d, 2, ps);
                NSX_SINGLE { c.star_valid = 0; s.pi_delta_n = -1; }  // Phase-2 costs: every reduced cost (and potential) changed
            }
            NSX_SYNC();
            NSX_SINGLE { nsx_drv_begin(c, L.drv, d.m, L.cmd, L.act); }
        } else {  // NSX_ACT_EXIT
            sweep.finish();
            break;
        }
    }
}
