/*
 * nsx_oracle.c - CPU restatement of the reference's network-simplex pivot loop.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA engine is compared against; it is
 * never linked into, imported by or called from the product path (network_flow_solver_b200/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_golden.py checks this restatement against
 * (i) the reference's own golden fixtures and (ii) entering-arc sequences / flows / potentials
 * recorded by running the unmodified reference in the build container
 * (tests/golden/make_golden.py -> tests/golden/*.json).
 *
 * Each function cites the reference lines it restates (paths relative to the reference root).
 * All arithmetic is IEEE float64 in the reference's operation order; compile with
 * -ffp-contract=off so that no multiply-add is fused.
 *
 * The tree is kept as parent pointers plus child lists; after a tree-changing pivot only the
 * re-hung subtree is re-walked (the reference re-walks the whole tree from the root every pivot,
 * basis.py:82-125 - the potentials it produces are a sequential fold of +-cost along the root
 * path, which a top-down walk of the changed subtree reproduces bit for bit).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/nsx_b200.h"

typedef struct {
    int32_t arc; /* -1 = none */
    int32_t dir;
    double key;
} cand_t;

typedef struct {
    /* sizes */
    int32_t n;  /* nodes incl. root */
    int64_t m;  /* real arcs */
    int64_t ma; /* all arcs = m + n - 1 */
    double tol;
    double penalty;
    /* arcs */
    const int32_t* tail_r;
    const int32_t* head_r;
    const double* pert;
    const double* upper_r;
    double* tcost;  /* [ma] tree cost of the current phase */
    double* flow;   /* [ma] */
    uint8_t* intree;
    uint8_t* touched;
    uint8_t* stale; /* warm start only: the reference's residual mirrors still hold the cold-start flow of this arc */
    double* weight; /* [m] Devex weights as seen by pricing */
    /* nodes */
    int32_t* parent;
    int32_t* pred;
    int8_t* pdir;
    int32_t* depth;
    double* pi;
    int32_t* fchild;
    int32_t* nsib;
    int32_t* psib;
    /* scratch */
    int32_t* stack;
    int32_t* path_h;
    int32_t* path_t;
    /* devex / tuner state */
    int64_t bs, pb;
    int32_t last_deg;
    int32_t ftc, ft_limit;
    int auto_block;
    int64_t tuner_total, tuner_deg, tuner_last;
    int64_t art_with_flow;
    int phase;
    int nthreads;
    /* candidate-list pricing state (simplex_pricing.py:375-542) */
    int32_t cl_list[128];
    int32_t cl_count, cl_since_refresh, cl_minor;
} oracle_t;

/* --- arc accessors over real + artificial ranges --- */
typedef struct {
    oracle_t o;
    int32_t* atail; /* [n-1] */
    int32_t* ahead;
    double* aupper;
} ctx_t;

static inline int32_t TAIL(const ctx_t* c, int64_t a) {
    return a < c->o.m ? c->o.tail_r[a] : c->atail[a - c->o.m];
}
static inline int32_t HEAD(const ctx_t* c, int64_t a) {
    return a < c->o.m ? c->o.head_r[a] : c->ahead[a - c->o.m];
}
static inline double UPPER(const ctx_t* c, int64_t a) {
    return a < c->o.m ? c->o.upper_r[a] : c->aupper[a - c->o.m];
}

/* Phase costs (simplex.py:1162-1168): Phase 1 real arc = pert - 1.0 - 1e-6*idx, artificial = penalty. */
static void apply_phase_costs(ctx_t* c, int phase) {
    oracle_t* o = &c->o;
    if (phase == 1) {
        for (int64_t i = 0; i < o->m; ++i) {
            double a = o->pert[i] - 1.0;
            double b = 1e-6 * (double)i;
            o->tcost[i] = a - b;
        }
    } else {
        for (int64_t i = 0; i < o->m; ++i) o->tcost[i] = o->pert[i];
    }
    for (int64_t i = o->m; i < o->ma; ++i) o->tcost[i] = o->penalty;
    o->phase = phase;
}

/* Child-list maintenance */
static inline void link_child(oracle_t* o, int32_t p, int32_t v) {
    int32_t f = o->fchild[p];
    o->nsib[v] = f;
    o->psib[v] = -1;
    if (f >= 0) o->psib[f] = v;
    o->fchild[p] = v;
}
static inline void unlink_child(oracle_t* o, int32_t p, int32_t v) {
    int32_t a = o->psib[v], b = o->nsib[v];
    if (a >= 0) o->nsib[a] = b; else o->fchild[p] = b;
    if (b >= 0) o->psib[b] = a;
    o->nsib[v] = o->psib[v] = -1;
}

/* Potential of child v from its parent (basis.py:111-117). */
static inline void set_pi(ctx_t* c, int32_t v) {
    oracle_t* o = &c->o;
    int32_t p = o->parent[v];
    double cst = o->tcost[o->pred[v]];
    o->pi[v] = (o->pdir[v] > 0) ? (o->pi[p] + cst) : (o->pi[p] - cst);
    o->depth[v] = o->depth[p] + 1;
}

/* Top-down recompute of the subtree rooted at r (r itself included). Returns subtree size. */
static int64_t recompute_subtree(ctx_t* c, int32_t r, int32_t* out_height) {
    oracle_t* o = &c->o;
    int64_t sp = 0, cnt = 0;
    int32_t base = o->depth[o->parent[r]] + 1, maxd = base;
    o->stack[sp++] = r;
    while (sp) {
        int32_t v = o->stack[--sp];
        set_pi(c, v);
        if (o->depth[v] > maxd) maxd = o->depth[v];
        ++cnt;
        for (int32_t ch = o->fchild[v]; ch >= 0; ch = o->nsib[ch]) o->stack[sp++] = ch;
    }
    if (out_height) *out_height = maxd - base + 1;
    return cnt;
}

/* Full rebuild of potentials (TreeBasis.rebuild, basis.py:82-125) - used at the phase switch. */
static void recompute_all(ctx_t* c) {
    oracle_t* o = &c->o;
    o->pi[0] = 0.0;
    o->depth[0] = 0;
    for (int32_t ch = o->fchild[0]; ch >= 0; ch = o->nsib[ch]) recompute_subtree(c, ch, NULL);
}

/* Initial tree: one artificial arc per node (simplex.py:619-728). */
static void init_tree(ctx_t* c, const double* supply) {
    oracle_t* o = &c->o;
    o->art_with_flow = 0;
    for (int32_t v = 0; v < o->n; ++v) {
        o->fchild[v] = o->nsib[v] = o->psib[v] = -1;
    }
    o->parent[0] = 0; o->pred[0] = -1; o->pdir[0] = 0; o->depth[0] = 0; o->pi[0] = 0.0;
    for (int64_t i = 0; i < o->m; ++i) { o->flow[i] = 0.0; o->intree[i] = 0; o->touched[i] = 0; }
    for (int32_t v = o->n - 1; v >= 1; --v) { /* reverse so that child lists end up in node order */
        int64_t a = o->m + (v - 1);
        double s = supply[v];
        if (fabs(s) <= o->tol) {
            c->atail[v - 1] = 0; c->ahead[v - 1] = v; c->aupper[v - 1] = INFINITY; o->flow[a] = 0.0;
        } else if (s > 0) {
            c->atail[v - 1] = v; c->ahead[v - 1] = 0; c->aupper[v - 1] = s; o->flow[a] = s;
            o->art_with_flow++;
        } else {
            c->atail[v - 1] = 0; c->ahead[v - 1] = v; c->aupper[v - 1] = -s; o->flow[a] = -s;
            o->art_with_flow++;
        }
        o->intree[a] = 1; o->touched[a] = 0;
        o->parent[v] = 0; o->pred[v] = (int32_t)a;
        o->pdir[v] = (c->atail[v - 1] == 0) ? 1 : -1;
        link_child(o, 0, v);
    }
}

/* Warm start (simplex.py:740-1021): the caller supplies which arcs form the tree and every arc's flow; parent
 * pointers come from a walk from the root over the tree arcs (TreeBasis.rebuild, basis.py:82-125). Returns 0, or
 * -1 when the marked arcs do not span all nodes. Artificial arcs keep the geometry init_tree gave them. */
static int init_tree_warm(ctx_t* c, const double* supply, const nsx_warm_start* w) {
    oracle_t* o = &c->o;
    init_tree(c, supply);
    int64_t marked = 0;
    for (int64_t a = 0; a < o->ma; ++a) { o->intree[a] = w->in_tree[a] ? 1 : 0; o->flow[a] = w->flow[a]; marked += o->intree[a]; }
    if (marked != o->n - 1) return -1;
    o->stale = malloc((size_t)o->ma + 1);
    memcpy(o->stale, o->intree, (size_t)o->ma);
    o->art_with_flow = 0;
    for (int64_t a = o->m; a < o->ma; ++a) if (o->flow[a] > o->tol) o->art_with_flow++;
    /* adjacency of the tree arcs (CSR) */
    int64_t* start = calloc((size_t)o->n + 1, sizeof(int64_t));
    int64_t* adj = malloc(sizeof(int64_t) * 2 * (size_t)(o->n > 1 ? o->n - 1 : 1));
    for (int64_t a = 0; a < o->ma; ++a) if (o->intree[a]) { start[TAIL(c, a) + 1]++; start[HEAD(c, a) + 1]++; }
    for (int32_t v = 0; v < o->n; ++v) start[v + 1] += start[v];
    int64_t* fill = malloc(sizeof(int64_t) * (size_t)o->n);
    for (int32_t v = 0; v < o->n; ++v) fill[v] = start[v];
    for (int64_t a = 0; a < o->ma; ++a) if (o->intree[a]) { adj[fill[TAIL(c, a)]++] = a; adj[fill[HEAD(c, a)]++] = a; }
    for (int32_t v = 0; v < o->n; ++v) { o->fchild[v] = o->nsib[v] = o->psib[v] = -1; o->parent[v] = -1; }
    o->parent[0] = 0; o->pred[0] = -1; o->pdir[0] = 0;
    int64_t sp = 0, seen = 1;
    o->stack[sp++] = 0;
    while (sp) {
        int32_t u = o->stack[--sp];
        for (int64_t k = start[u]; k < start[u + 1]; ++k) {
            int64_t a = adj[k];
            int32_t v = TAIL(c, a) == u ? HEAD(c, a) : TAIL(c, a);
            if (v == u || o->parent[v] >= 0) continue;
            o->parent[v] = u; o->pred[v] = (int32_t)a; o->pdir[v] = TAIL(c, a) == u ? 1 : -1;
            link_child(o, u, v);
            o->stack[sp++] = v; ++seen;
        }
    }
    free(start); free(adj); free(fill);
    return seen == o->n ? 0 : -1;
}

/* ---------------- pricing ---------------- */

static inline double rc_of(const ctx_t* c, int64_t i, double cost) {
    const oracle_t* o = &c->o;
    return (cost + o->pi[o->tail_r[i]]) - o->pi[o->head_r[i]]; /* simplex.py:508-512 */
}
static inline double fwd_res(const ctx_t* c, int64_t i) {
    double u = c->o.upper_r[i];
    return isinf(u) ? INFINITY : u - c->o.flow[i]; /* simplex.py:453-455 */
}

/* Dantzig over [a,b): best improving (min key, lowest index) and first zero candidate
 * (simplex_pricing.py:107-137).  Transportation row-scan (specialized_pivots.py:89-120) is the
 * same selection without zero candidates. */
static void dantzig_range(const ctx_t* c, int64_t a, int64_t b, cand_t* best, cand_t* zero) {
    const oracle_t* o = &c->o;
    const double tol = o->tol;
    best->arc = -1; best->dir = 0; best->key = 0.0;
    zero->arc = -1; zero->dir = 0; zero->key = 0.0;
    for (int64_t i = a; i < b; ++i) {
        if (o->intree[i]) continue;
        double rc = rc_of(c, i, o->tcost[i]);
        double fr = fwd_res(c, i), br = o->flow[i];
        if (fr > tol && rc < -tol) {
            if (best->arc < 0 || rc < best->key) { best->arc = (int32_t)i; best->dir = 1; best->key = rc; }
        } else if (br > tol && rc > tol) {
            if (best->arc < 0 || -rc < best->key) { best->arc = (int32_t)i; best->dir = -1; best->key = -rc; }
        } else if (zero->arc < 0 && fabs(rc) <= tol) {
            if (fr > tol) { zero->arc = (int32_t)i; zero->dir = 1; }
            else if (br > tol) { zero->arc = (int32_t)i; zero->dir = -1; }
        }
    }
}

/* Structure-specific entering rules (specialized_pivots.py). All scan the real arcs in index order with the tree cost
 * of the current phase and return 0 when nothing qualifies (the configured strategy then runs, simplex.py:1066-1075). */
/* AssignmentPivotStrategy.find_entering_arc_min_cost (specialized_pivots.py:179-209): forward direction only; a later
 * arc replaces the incumbent only when it is better by more than the tolerance. */
static int assignment_select(const ctx_t* c, int32_t* arc, int32_t* dir) {
    const oracle_t* o = &c->o;
    const double tol = o->tol;
    double best_rc = 0.0; int found = 0;
    for (int64_t i = 0; i < o->m; ++i) {
        if (o->intree[i]) continue;
        double rc = rc_of(c, i, o->tcost[i]);
        if (fwd_res(c, i) > tol && rc < best_rc - tol) { best_rc = rc; *arc = (int32_t)i; *dir = 1; found = 1; }
    }
    return found;
}
/* MaxFlowPivotStrategy.find_entering_arc (specialized_pivots.py:294-343): merit = residual * |rc|, strictly larger wins. */
static int maxflow_select(const ctx_t* c, int32_t* arc, int32_t* dir) {
    const oracle_t* o = &c->o;
    const double tol = o->tol;
    double best = -INFINITY; int found = 0;
    for (int64_t i = 0; i < o->m; ++i) {
        if (o->intree[i]) continue;
        double rc = rc_of(c, i, o->tcost[i]);
        double fr = fwd_res(c, i), br = o->flow[i];
        if (fr > tol && rc < -tol) { double merit = fr * fabs(rc); if (merit > best) { best = merit; *arc = (int32_t)i; *dir = 1; found = 1; } }
        if (br > tol && rc > tol) { double merit = br * fabs(rc); if (merit > best) { best = merit; *arc = (int32_t)i; *dir = -1; found = 1; } }
    }
    return found;
}
/* ShortestPathPivotStrategy.find_entering_arc (specialized_pivots.py:368-424). The distance labels only ever decide
 * "does the tail have a label"; labels are seeded by a BFS from the source over all real arcs (:426-450) and a head gets
 * one when its tail has one, so the labelled set is the set reachable from the source for the whole solve: mask[]. */
static int shortest_path_select(const ctx_t* c, const uint8_t* mask, int32_t* arc, int32_t* dir) {
    const oracle_t* o = &c->o;
    const double tol = o->tol;
    double best_rc = 0.0; int found = 0;
    for (int64_t i = 0; i < o->m; ++i) {
        if (o->intree[i]) continue;
        double rc = rc_of(c, i, o->tcost[i]);
        double fr = fwd_res(c, i), br = o->flow[i];
        if (fr > tol && rc < -tol) {
            if (mask[o->tail_r[i]] && rc < best_rc - tol) { best_rc = rc; *arc = (int32_t)i; *dir = 1; found = 1; }
        }
        if (br > tol && rc > tol && -rc < best_rc - tol) { best_rc = -rc; *arc = (int32_t)i; *dir = -1; found = 1; }
    }
    return found;
}

static int dantzig_select(const ctx_t* c, int allow_zero, int32_t* arc, int32_t* dir) {
    const oracle_t* o = &c->o;
    int T = o->nthreads > 1 && o->m >= 65536 ? o->nthreads : 1;
    cand_t bests[64], zeros[64];
    if (T > 64) T = 64;
    if (T == 1) {
        dantzig_range(c, 0, o->m, &bests[0], &zeros[0]);
    } else {
#ifdef _OPENMP
#pragma omp parallel for num_threads(T) schedule(static, 1)
#endif
        for (int t = 0; t < T; ++t) {
            int64_t a = o->m * t / T, b = o->m * (t + 1) / T;
            dantzig_range(c, a, b, &bests[t], &zeros[t]);
        }
    }
    cand_t best = bests[0], zero = zeros[0];
    for (int t = 1; t < T; ++t) { /* chunks are in index order: strict < keeps the lowest index */
        if (bests[t].arc >= 0 && (best.arc < 0 || bests[t].key < best.key)) best = bests[t];
        if (zero.arc < 0 && zeros[t].arc >= 0) zero = zeros[t];
    }
    if (best.arc >= 0) { *arc = best.arc; *dir = best.dir; return 1; }
    if (allow_zero && zero.arc >= 0) { *arc = zero.arc; *dir = zero.dir; return 1; }
    return 0;
}

/* One Devex block [st,en) (simplex.py:551-617): separate forward / backward first-argmax of rc^2/w,
 * forward wins only if strictly greater; zero candidates when allowed.  Uses the perturbed Phase-2
 * costs in BOTH phases (arc_costs is never refreshed: simplex.py:440,1162-1168). */
typedef struct { int32_t fi, bi, fz, bz; double fm, bm; } devex_part_t;

static void devex_range(const ctx_t* c, int64_t a, int64_t b, devex_part_t* r) {
    const oracle_t* o = &c->o;
    const double tol = o->tol;
    r->fi = r->bi = r->fz = r->bz = -1; r->fm = r->bm = -INFINITY;
    for (int64_t i = a; i < b; ++i) {
        if (o->intree[i] || i == o->last_deg) continue;
        double rc = rc_of(c, i, o->pert[i]);
        double fr = fwd_res(c, i), br = o->flow[i];
        double merit = (rc * rc) / o->weight[i];
        if (fr > tol && rc < -tol) { if (merit > r->fm) { r->fm = merit; r->fi = (int32_t)i; } }
        if (br > tol && rc > tol) { if (merit > r->bm) { r->bm = merit; r->bi = (int32_t)i; } }
        if (fabs(rc) <= tol) {
            if (r->fz < 0 && fr > tol) r->fz = (int32_t)i;
            if (r->bz < 0 && br > tol) r->bz = (int32_t)i;
        }
    }
}

static int devex_block(const ctx_t* c, int64_t st, int64_t en, int allow_zero, int32_t* arc,
                       int32_t* dir, double* merit) {
    const oracle_t* o = &c->o;
    int T = o->nthreads > 1 && (en - st) >= 65536 ? o->nthreads : 1;
    devex_part_t parts[64];
    if (T > 64) T = 64;
    if (T == 1) {
        devex_range(c, st, en, &parts[0]);
    } else {
#ifdef _OPENMP
#pragma omp parallel for num_threads(T) schedule(static, 1)
#endif
        for (int t = 0; t < T; ++t) {
            int64_t a = st + (en - st) * t / T, b = st + (en - st) * (t + 1) / T;
            devex_range(c, a, b, &parts[t]);
        }
    }
    devex_part_t r = parts[0];
    for (int t = 1; t < T; ++t) {
        if (parts[t].fm > r.fm) { r.fm = parts[t].fm; r.fi = parts[t].fi; }
        if (parts[t].bm > r.bm) { r.bm = parts[t].bm; r.bi = parts[t].bi; }
        if (r.fz < 0) r.fz = parts[t].fz;
        if (r.bz < 0) r.bz = parts[t].bz;
    }
    if (r.fm > r.bm) {
        if (r.fm > -INFINITY) { *arc = r.fi; *dir = 1; *merit = r.fm; return 1; }
    } else {
        if (r.bm > -INFINITY) { *arc = r.bi; *dir = -1; *merit = r.bm; return 1; }
    }
    if (allow_zero) {
        if (r.fz >= 0) { *arc = r.fz; *dir = 1; *merit = 0.0; return 1; }
        if (r.bz >= 0) { *arc = r.bz; *dir = -1; *merit = 0.0; return 1; }
    }
    return 0;
}

/* Block search (simplex_pricing.py:325-357). *want_weight = 1 when the selected arc's Devex weight
 * must be set to its tree-path length (merit > 0). */
static int devex_select(ctx_t* c, int allow_zero, int32_t* arc, int32_t* dir, int* want_weight,
                        int64_t* priced) {
    oracle_t* o = &c->o;
    int64_t m = o->m;
    int64_t bc = (m + o->bs - 1) / o->bs;
    if (bc < 1) bc = 1;
    for (int64_t k = 0; k < bc; ++k) {
        int64_t st = o->pb * o->bs;
        if (st >= m) { o->pb = 0; st = 0; }
        int64_t en = st + o->bs < m ? st + o->bs : m;
        double merit = 0.0;
        *priced += en - st;
        if (en > st && devex_block(c, st, en, allow_zero, arc, dir, &merit)) {
            o->last_deg = -1;
            *want_weight = merit > 0.0;
            return 1;
        }
        o->pb = (o->pb + 1) % bc;
    }
    return 0;
}

/* Loop-based Devex (DevexPricing.select_entering_arc, use_vectorized_pricing=False; simplex_pricing.py:205-269).
 * _is_better_candidate (:294-308) reduces to "merit > best + tol" because the scan runs in ascending index order. */
static int devex_loop_select(ctx_t* c, int allow_zero, int32_t* arc, int32_t* dir, int* want_weight, int64_t* priced) {
    oracle_t* o = &c->o;
    const double tol = o->tol;
    int64_t m = o->m;
    int64_t bc = (m + o->bs - 1) / o->bs;
    if (bc < 1) bc = 1;
    for (int64_t k = 0; k < bc; ++k) {
        int64_t st = o->pb * o->bs;
        if (st >= m) { o->pb = 0; st = 0; }
        int64_t en = st + o->bs < m ? st + o->bs : m;
        *priced += en - st;
        double best_merit = -INFINITY;
        int32_t best = -1, bdir = 0, zero = -1, zdir = 0;
        for (int64_t i = st; i < en; ++i) {
            if (o->intree[i]) continue;
            double rc = rc_of(c, i, o->tcost[i]);
            double fr = fwd_res(c, i), br = o->flow[i];
            int d = 0;
            if (fr > tol && rc < -tol) d = 1;
            else if (br > tol && rc > tol) d = -1;
            if (d) {
                double w = o->weight[i] > 1e-12 ? o->weight[i] : 1e-12;
                double merit = (rc * rc) / w;
                if (merit > best_merit + tol) { best_merit = merit; best = (int32_t)i; bdir = d; }
                continue;
            }
            if (allow_zero && zero < 0 && fabs(rc) <= tol) {
                if (fr > tol) { zero = (int32_t)i; zdir = 1; }
                else if (br > tol) { zero = (int32_t)i; zdir = -1; }
            }
        }
        if (best >= 0) { *arc = best; *dir = bdir; *want_weight = 1; return 1; }
        o->pb = (o->pb + 1) % bc;
        if (zero >= 0) { *arc = zero; *dir = zdir; *want_weight = 0; return 1; }
    }
    return 0;
}

/* ---------------- candidate-list pricing (SolverOptions.pricing_strategy "candidate_list"; also what
 * "adaptive" - the reference's default - amounts to: AdaptivePricing (simplex_pricing.py:545-639) only
 * leaves the candidate list after 5 CONSECUTIVE searches that return None, and a search that returns
 * None ends the phase (simplex.py:1113-1116), so the counter never gets past 2) ---------------- */
#define CL_SIZE 100   /* simplex.py:232 */
#define CL_REFRESH 10 /* simplex.py:233 */
#define CL_MINOR 3    /* simplex_pricing.py:400 */

/* _scan_candidates (simplex_pricing.py:460-505): Dantzig rule over the list, list order, strict < */
static int cl_scan(const ctx_t* c, int allow_zero, int32_t* arc, int32_t* dir) {
    const oracle_t* o = &c->o;
    const double tol = o->tol;
    int32_t best = -1, bdir = 0;
    double best_rc = 0.0;
    for (int32_t k = 0; k < o->cl_count; ++k) {
        int32_t i = o->cl_list[k];
        if (o->intree[i]) continue;
        double rc = rc_of(c, i, o->tcost[i]);
        double fr = fwd_res(c, i), br = o->flow[i];
        if (fr > tol && rc < -tol) {
            if (best < 0 || rc < best_rc) { best = i; bdir = 1; best_rc = rc; }
        } else if (br > tol && rc > tol) {
            if (best < 0 || -rc < best_rc) { best = i; bdir = -1; best_rc = -rc; }
        } else if (allow_zero && fr > tol && fabs(rc) <= tol && best < 0) {
            best = i; bdir = 1;
        } else if (allow_zero && br > tol && fabs(rc) <= tol && best < 0) {
            best = i; bdir = -1;
        }
    }
    if (best < 0) return 0;
    *arc = best; *dir = bdir;
    return 1;
}

/* _refresh_candidate_list (simplex_pricing.py:507-536): the CL_SIZE arcs of largest |rc| among the improving
 * ones; Python sorts the (merit, idx) tuples in descending order, so ties go to the LARGER index. */
static void cl_refresh(ctx_t* c, int64_t* priced) {
    oracle_t* o = &c->o;
    const double tol = o->tol;
    double hm[CL_SIZE]; int32_t hi[CL_SIZE]; /* min-heap on (merit, idx) */
    int32_t cnt = 0;
    *priced += o->m;
    for (int64_t i = 0; i < o->m; ++i) {
        if (o->intree[i]) continue;
        double rc = rc_of(c, i, o->tcost[i]);
        double fr = fwd_res(c, i), br = o->flow[i];
        double merit = 0.0;
        if ((fr > tol && rc < -tol) || (br > tol && rc > tol)) merit = fabs(rc);
        if (!(merit > tol)) continue;
#define CL_LESS(m1, i1, m2, i2) ((m1) < (m2) || ((m1) == (m2) && (i1) < (i2)))
        if (cnt < CL_SIZE) {
            int32_t k = cnt++;
            hm[k] = merit; hi[k] = (int32_t)i;
            while (k > 0) { /* sift up */
                int32_t p = (k - 1) / 2;
                if (!CL_LESS(hm[k], hi[k], hm[p], hi[p])) break;
                double tm = hm[k]; hm[k] = hm[p]; hm[p] = tm;
                int32_t ti = hi[k]; hi[k] = hi[p]; hi[p] = ti;
                k = p;
            }
        } else if (CL_LESS(hm[0], hi[0], merit, (int32_t)i)) {
            hm[0] = merit; hi[0] = (int32_t)i;
            int32_t k = 0;
            for (;;) { /* sift down */
                int32_t l = 2 * k + 1, r = l + 1, sm = k;
                if (l < cnt && CL_LESS(hm[l], hi[l], hm[sm], hi[sm])) sm = l;
                if (r < cnt && CL_LESS(hm[r], hi[r], hm[sm], hi[sm])) sm = r;
                if (sm == k) break;
                double tm = hm[k]; hm[k] = hm[sm]; hm[sm] = tm;
                int32_t ti = hi[k]; hi[k] = hi[sm]; hi[sm] = ti;
                k = sm;
            }
        }
    }
    /* descending (merit, idx): selection sort of <= 100 entries */
    for (int32_t a = 0; a < cnt; ++a) {
        int32_t mx = a;
        for (int32_t b = a + 1; b < cnt; ++b)
            if (CL_LESS(hm[mx], hi[mx], hm[b], hi[b])) mx = b;
        double tm = hm[a]; hm[a] = hm[mx]; hm[mx] = tm;
        int32_t ti = hi[a]; hi[a] = hi[mx]; hi[mx] = ti;
        o->cl_list[a] = hi[a];
    }
    o->cl_count = cnt;
#undef CL_LESS
}

/* CandidateListPricing.select_entering_arc (simplex_pricing.py:418-458) */
static int cl_select(ctx_t* c, int allow_zero, int32_t* arc, int32_t* dir, int64_t* priced) {
    oracle_t* o = &c->o;
    if (o->cl_count > 0 && o->cl_minor < CL_MINOR) {
        if (cl_scan(c, allow_zero, arc, dir)) { o->cl_minor++; return 1; }
    }
    o->cl_since_refresh++;
    o->cl_minor = 0;
    if (o->cl_since_refresh >= CL_REFRESH || o->cl_count == 0) { cl_refresh(c, priced); o->cl_since_refresh = 0; }
    if (cl_scan(c, allow_zero, arc, dir)) return 1;
    if (o->cl_since_refresh > 0) {
        cl_refresh(c, priced);
        o->cl_since_refresh = 0;
        return cl_scan(c, allow_zero, arc, dir);
    }
    return 0;
}

/* ---------------- pivot ---------------- */

typedef struct {
    int64_t sum_cycle, sum_subtree, max_subtree, sum_height, tree_updates, resets, degenerate;
} ostats_t;

/* Returns 0 ok, 3 unbounded. (simplex.py:1176-1425) */
static int pivot(ctx_t* c, int32_t e, int32_t dir, int want_weight, ostats_t* st) {
    oracle_t* o = &c->o;
    const double tol = o->tol;
    int32_t t = dir == 1 ? TAIL(c, e) : HEAD(c, e);
    int32_t h = dir == 1 ? HEAD(c, e) : TAIL(c, e);

    /* tree path h -> join -> t (basis.py:178-241 returns it listed from the head side) */
    int64_t nh = 0, nt = 0;
    int32_t u = h, v = t;
    while (u != v) {
        if (o->depth[u] >= o->depth[v]) { o->path_h[nh++] = u; u = o->parent[u]; }
        else { o->path_t[nt++] = v; v = o->parent[v]; }
    }
    /* ratio test in the reference's scan order (simplex.py:1201-1229) */
    double theta = INFINITY, best = -INFINITY;
    int32_t leave = e;
    int leave_side = 0; /* 0 entering, 1 h-side, 2 t-side */
    int64_t leave_pos = -1;
    for (int64_t k = 0; k < nh + nt + 1; ++k) {
        int32_t a; int sign; int side; int64_t pos;
        if (k < nh) { int32_t x = o->path_h[k]; a = o->pred[x]; sign = o->pdir[x] < 0 ? 1 : -1; side = 1; pos = k; }
        else if (k < nh + nt) { pos = nt - 1 - (k - nh); int32_t x = o->path_t[pos]; a = o->pred[x]; sign = o->pdir[x] > 0 ? 1 : -1; side = 2; }
        else { a = e; sign = dir; side = 0; pos = -1; }
        double r, fl = o->flow[a];
        if (o->stale && o->stale[a]) { /* forward/backward_residuals are only refreshed for arcs of a pivot cycle
                                          (simplex.py:1276-1283) and never after a warm start (simplex.py:728 is the only
                                          full sync): the ratio test still sees the cold-start flow - 0 on a real arc,
                                          |supply| on the artificial arc of a supply / demand node */
            double up0 = UPPER(c, a);
            fl = (a < o->m || isinf(up0)) ? 0.0 : up0;
        }
        if (sign == 1) { double up = UPPER(c, a); r = isinf(up) ? INFINITY : up - fl; }
        else r = fl;
        if (r < theta - tol) { theta = r; leave = a; best = r; leave_side = side; leave_pos = pos; }
        else if (fabs(r - theta) <= tol) {
            if (r > best + tol || (fabs(r - best) <= tol && a < leave)) { leave = a; best = r; leave_side = side; leave_pos = pos; }
        }
    }
    if (isinf(theta)) return 3;
    if (!(theta > 0.0)) theta = 0.0;
    if (theta <= tol) st->degenerate++;
    st->sum_cycle += nh + nt + 1;

    /* flow update (simplex.py:1255-1283) */
    for (int64_t k = 0; k < nh + nt + 1; ++k) {
        int32_t a; int sign;
        if (k < nh) { int32_t x = o->path_h[k]; a = o->pred[x]; sign = o->pdir[x] < 0 ? 1 : -1; }
        else if (k < nh + nt) { int32_t x = o->path_t[k - nh]; a = o->pred[x]; sign = o->pdir[x] > 0 ? 1 : -1; }
        else { a = e; sign = dir; }
        double old = o->flow[a], up = UPPER(c, a);
        int art = a >= o->m;
        int had = art && old > tol;
        double f = old + (double)sign * theta;
        if (theta > 0.0) o->touched[a] = 1;
        if (f < 0.0 - tol) { f = 0.0; o->touched[a] = 0; }
        if (!isinf(up) && f > up + tol) { f = up; o->touched[a] = 0; }
        o->flow[a] = f;
        if (o->stale) o->stale[a] = 0;
        if (art) { int has = f > tol; if (had && !has) o->art_with_flow--; else if (!had && has) o->art_with_flow++; }
    }
    /* Devex weight := tree-path length, set by pricing before the pivot (simplex_pricing.py:350-352;
     * ||B^-1 a||^2 of a tree basis = number of tree arcs on the path) */
    if (want_weight && e < o->m) {
        double w = (double)(nh + nt);
        if (w <= 1e-12) w = 1e-12; else if (w > 1e12) w = 1e12;
        o->weight[e] = w;
    }
    /* tuner bookkeeping (simplex.py:1317-1318) */
    int is_deg = (leave == e) || (fabs(theta) < tol);
    o->tuner_total++;
    if (is_deg) o->tuner_deg++;

    if (leave == e) { /* bound flip (simplex.py:1320-1334) */
        o->last_deg = e; /* only consulted by Devex pricing */
        return 0;
    }
    /* tree update: the subtree below the leaving arc is re-hung under the entering arc */
    o->intree[e] = 1; o->intree[leave] = 0;
    int32_t q, p; /* q = entering endpoint inside the cut subtree, p = the other endpoint */
    int32_t* stem; int64_t stem_len;
    if (leave_side == 1) { q = h; p = t; stem = o->path_h; stem_len = leave_pos + 1; }
    else { q = t; p = h; stem = o->path_t; stem_len = leave_pos + 1; }
    /* stem[0] = q ... stem[stem_len-1] = r whose pred arc is the leaving arc */
    int32_t r = stem[stem_len - 1];
    unlink_child(o, o->parent[r], r);
    /* reverse parent pointers along the stem */
    int32_t carry_arc = e;
    int8_t carry_dir = (TAIL(c, e) == p) ? 1 : -1; /* arc points parent(p) -> child(q)? */
    int32_t newpar = p;
    for (int64_t k = 0; k < stem_len; ++k) {
        int32_t x = stem[k];
        int32_t old_arc = o->pred[x]; int8_t old_dir = o->pdir[x]; int32_t old_par = o->parent[x];
        if (k + 1 < stem_len) unlink_child(o, old_par, x); /* x leaves its old parent (next stem node) */
        o->parent[x] = newpar; o->pred[x] = carry_arc; o->pdir[x] = carry_dir;
        link_child(o, newpar, x);
        carry_arc = old_arc; carry_dir = (int8_t)-old_dir; newpar = x;
    }
    int32_t height = 0;
    int64_t sz = recompute_subtree(c, q, &height);
    st->sum_subtree += sz; if (sz > st->max_subtree) st->max_subtree = sz;
    st->sum_height += height; st->tree_updates++;

    /* reset cadence (simplex.py:1373-1425): weights := 1 and pricing_block := 0 */
    if (o->ftc >= o->ft_limit) {
        o->ftc = 0;
        for (int64_t i = 0; i < o->m; ++i) o->weight[i] = 1.0;
        o->pb = 0;
        o->cl_count = 0; o->cl_since_refresh = 0; o->cl_minor = 0; /* pricing_strategy.reset(), simplex.py:1771-1776 */
        st->resets++;
    } else o->ftc++;
    return 0;
}

/* Block-size adaptation (simplex_adaptive.py:98-151) */
static void adapt_block(oracle_t* o, int64_t iteration) {
    if (!o->auto_block) return;
    if (iteration - o->tuner_last < 50) return;
    if (o->tuner_total < 10) return;
    double ratio = (double)o->tuner_deg / (double)o->tuner_total;
    if (ratio > 0.30) { int64_t nb = (int64_t)((double)o->bs * 1.5); o->bs = nb < o->m ? nb : o->m; }
    else if (ratio < 0.10) { int64_t nb = (int64_t)((double)o->bs * 0.75); o->bs = nb > 10 ? nb : 10; }
    o->tuner_deg = 0; o->tuner_total = 0; o->tuner_last = iteration;
}

static int find_entering(ctx_t* c, const nsx_options* opt, int allow_zero, int32_t* arc, int32_t* dir,
                         int* want_weight, int64_t* priced) {
    oracle_t* o = &c->o;
    *want_weight = 0;
    if (opt->row_scan_first) { /* simplex.py:1060-1064 */
        *priced += o->m;
        int hit;
        if (opt->row_scan_first == NSX_SPECIAL_ASSIGNMENT) hit = assignment_select(c, arc, dir);
        else if (opt->row_scan_first == NSX_SPECIAL_MAX_FLOW) hit = maxflow_select(c, arc, dir);
        else if (opt->row_scan_first == NSX_SPECIAL_SHORTEST_PATH) hit = opt->node_mask ? shortest_path_select(c, opt->node_mask, arc, dir) : 0;
        else hit = dantzig_select(c, 0, arc, dir);
        if (hit) return 1;
    }
    if (opt->pricing == NSX_PRICING_DANTZIG) { *priced += o->m; return dantzig_select(c, allow_zero, arc, dir); }
    if (opt->pricing == NSX_PRICING_CANDIDATE_LIST) return cl_select(c, allow_zero, arc, dir, priced);
    if (opt->pricing == NSX_PRICING_DEVEX_LOOP) return devex_loop_select(c, allow_zero, arc, dir, want_weight, priced);
    return devex_select(c, allow_zero, arc, dir, want_weight, priced);
}

int nsx_oracle_solve_warm(const nsx_problem* pb, const nsx_options* opt, const nsx_warm_start* warm, nsx_result* res, int nthreads);
int nsx_oracle_solve(const nsx_problem* pb, const nsx_options* opt, nsx_result* res, int nthreads) {
    return nsx_oracle_solve_warm(pb, opt, NULL, res, nthreads);
}
int nsx_oracle_solve_warm(const nsx_problem* pb, const nsx_options* opt, const nsx_warm_start* warm, nsx_result* res, int nthreads) {
    ctx_t cx; memset(&cx, 0, sizeof cx);
    oracle_t* o = &cx.o;
    o->n = pb->n_nodes; o->m = pb->n_arcs; o->ma = o->m + o->n - 1;
    o->tol = opt->tolerance; o->penalty = pb->penalty;
    o->tail_r = pb->tail; o->head_r = pb->head; o->pert = pb->pert_cost; o->upper_r = pb->upper;
    o->nthreads = nthreads > 0 ? nthreads : 1;
    size_t n = (size_t)o->n, ma = (size_t)o->ma, m = (size_t)o->m;
    o->tcost = malloc(ma * 8 + 8); o->flow = malloc(ma * 8 + 8);
    o->intree = calloc(ma + 1, 1); o->touched = calloc(ma + 1, 1);
    o->weight = malloc(m * 8 + 8);
    o->parent = malloc(n * 4); o->pred = malloc(n * 4); o->pdir = malloc(n); o->depth = malloc(n * 4);
    o->pi = malloc(n * 8); o->fchild = malloc(n * 4); o->nsib = malloc(n * 4); o->psib = malloc(n * 4);
    o->stack = malloc(n * 4); o->path_h = malloc(n * 4); o->path_t = malloc(n * 4);
    cx.atail = malloc(n * 4); cx.ahead = malloc(n * 4); cx.aupper = malloc(n * 8);
    for (size_t i = 0; i < m; ++i) o->weight[i] = 1.0;
    o->bs = opt->block_size > 0 ? opt->block_size : 1; o->pb = 0; o->last_deg = -1;
    o->ftc = 0; o->ft_limit = opt->ft_update_limit; o->auto_block = opt->auto_block;
    o->tuner_total = o->tuner_deg = o->tuner_last = 0;

    int bad_warm = 0;
    if (warm) bad_warm = init_tree_warm(&cx, pb->supply, warm); else init_tree(&cx, pb->supply);
    ostats_t st; memset(&st, 0, sizeof st);
    int64_t total = 0, priced = 0, trace_len = 0;
    int status = NSX_STATUS_OPTIMAL;
    int64_t maxit = opt->max_iterations;
    int64_t it = 0;
    int32_t arc = -1, dir = 0; int ww = 0;
    if (bad_warm) { status = -1; goto done; }
    if (warm && warm->start_phase == 2) { /* simplex.py:1504-1513: no artificial arc in the tree, Phase 1 skipped */
        res->phase1_iterations = 0;
        res->artificial_with_flow = o->art_with_flow;
        goto phase2;
    }

    /* Phase 1 (simplex.py:1541-1554) */
    apply_phase_costs(&cx, 1);
    recompute_all(&cx);
    while (it < maxit) {
        if (!find_entering(&cx, opt, 1, &arc, &dir, &ww, &priced)) break;
        if (res->entering_trace && trace_len < opt->trace_capacity) res->entering_trace[trace_len] = arc * 2 + (dir < 0);
        trace_len++;
        int rc = pivot(&cx, arc, dir, ww, &st);
        if (rc == 3) { status = NSX_STATUS_UNBOUNDED; res->unbounded_arc = arc; { double urc = rc_of(&cx, arc, o->tcost[arc]); res->unbounded_rc = dir == 1 ? urc : -urc; } /* simplex.py:1233-1239 */ goto done; }
        it++;
        adapt_block(o, it);
        if (o->art_with_flow == 0) break;
    }
    total = it;
    res->phase1_iterations = it;
    res->artificial_with_flow = o->art_with_flow;
    int unbalanced = 0;
    if (warm) { /* conservation check after Phase 1 (simplex.py:1575-1598), every arc in index order per node. It can only
                   fail after a warm start: the stale residuals let the ratio test over-push an arc and the flow update
                   clamps it back to its bound (simplex.py:1263-1266) */
        double* net = malloc(n * 8);
        for (size_t v = 0; v < n; ++v) net[v] = pb->supply[v];
        for (int64_t a = 0; a < o->ma; ++a) { net[TAIL(&cx, a)] -= o->flow[a]; net[HEAD(&cx, a)] += o->flow[a]; }
        for (size_t v = 1; v < n; ++v) if (fabs(net[v]) > o->tol) unbalanced = 1;
        free(net);
    }
    if (o->art_with_flow > 0 || unbalanced) {
        status = total >= maxit ? NSX_STATUS_ITERATION_LIMIT_P1 : NSX_STATUS_INFEASIBLE;
        goto done;
    }
    /* Phase 2 (simplex.py:1626-1644) */
phase2:
    {
        int64_t remaining = maxit - total; if (remaining < 0) remaining = 0;
        apply_phase_costs(&cx, 2);
        recompute_all(&cx);
        it = 0;
        while (it < remaining) {
            if (!find_entering(&cx, opt, 0, &arc, &dir, &ww, &priced)) break;
            if (res->entering_trace && trace_len < opt->trace_capacity) res->entering_trace[trace_len] = arc * 2 + (dir < 0);
            trace_len++;
            int rc = pivot(&cx, arc, dir, ww, &st);
            if (rc == 3) { status = NSX_STATUS_UNBOUNDED; res->unbounded_arc = arc; { double urc = rc_of(&cx, arc, o->tcost[arc]); res->unbounded_rc = dir == 1 ? urc : -urc; } /* simplex.py:1233-1239 */ total += it; goto done; }
            it++;
            adapt_block(o, total + it);
        }
        total += it;
        if (total >= maxit) { /* simplex.py:1678-1699 */
            status = find_entering(&cx, opt, 0, &arc, &dir, &ww, &priced) ? NSX_STATUS_ITERATION_LIMIT : NSX_STATUS_OPTIMAL;
        }
    }
done:
    res->status = status;
    res->iterations = total;
    res->trace_len = trace_len;
    res->degenerate_pivots = st.degenerate;
    res->tree_updates = st.tree_updates;
    res->weight_resets = st.resets;
    res->final_block_size = o->bs;
    res->arcs_priced = priced;
    res->sum_cycle_len = st.sum_cycle; res->sum_subtree = st.sum_subtree; res->max_subtree = st.max_subtree;
    res->sum_rounds = st.sum_height;
    if (res->flow) memcpy(res->flow, o->flow, ma * 8);
    if (res->potential) memcpy(res->potential, o->pi, n * 8);
    if (res->state) {
        for (size_t i = 0; i < ma; ++i) {
            double up = i < m ? o->upper_r[i] : cx.aupper[i - m];
            double fr = isinf(up) ? INFINITY : up - o->flow[i];
            res->state[i] = (uint8_t)((o->intree[i] ? NSX_ARC_IN_TREE : 0) | (fr > o->tol ? NSX_ARC_CAN_FWD : 0) |
                                      (o->flow[i] > o->tol ? NSX_ARC_CAN_BWD : 0) | (o->touched[i] ? NSX_ARC_TOUCHED : 0) |
                                      (o->stale && o->stale[i] ? NSX_ARC_STALE : 0));
        }
    }
    free(o->tcost); free(o->flow); free(o->intree); free(o->touched); free(o->weight); free(o->stale);
    free(o->parent); free(o->pred); free(o->pdir); free(o->depth); free(o->pi);
    free(o->fchild); free(o->nsib); free(o->psib); free(o->stack); free(o->path_h); free(o->path_t);
    free(cx.atail); free(cx.ahead); free(cx.aupper);
    return bad_warm ? -1 : 0;
}

int nsx_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
