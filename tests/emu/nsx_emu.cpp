// nsx_emu.cpp - serial HOST EMULATION of the device pivot code (nsx_core.cuh with NSX_HOST_EMU).
//
// Test infrastructure only: lets the CPU test-suite exercise the exact source the CUDA kernels
// are built from (tree re-hang on the preorder array, wavefront potential recompute, driver state
// machine) on machines without a GPU.  Not linked into libnsx_b200.so, not importable from the
// package; the product has no host execution path.
#define NSX_HOST_EMU 1
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#ifdef NSX_HOST_MT
// Multi-threaded variant (libnsx_emu_mt.so): the pivot CTA is NSX_EMU_THREADS real threads, NSX_SYNC a barrier.
#include <pthread.h>
#include <thread>
thread_local int nsx_mt_tid = 0;
int nsx_mt_nthreads = 1;
static pthread_barrier_t nsx_mt_bar;
void nsx_mt_barrier() { pthread_barrier_wait(&nsx_mt_bar); }
#endif

#include "../../network_flow_solver_b200/csrc/nsx_core.cuh"
#include "../../network_flow_solver_b200/csrc/nsx_warm.h"

namespace {

// Design statistics (NSX_EMU_INCSTATS=1, Dantzig full sweeps): how much of a sweep an incremental scheme would have to
// redo - arcs incident to nodes whose potential changed since the previous sweep, and tail groups ("rows") whose cached
// best arc points into such a node.
struct IncStats {
    bool on = false;
    std::vector<double> prev_pi;
    std::vector<int32_t> row_best;   // per tail node: arc of the row's best candidate at the previous sweep, -1 none
    std::vector<int64_t> row_begin;  // CSR by tail (arcs are sorted by tail)
    std::vector<int32_t> indeg;
    long long sweeps = 0, dirty_nodes = 0, star_arcs = 0, rescan_rows = 0, rescan_arcs = 0, total_arcs = 0;
};
static IncStats g_inc;

struct SerialSweep {
    const NsxDev& d;
    const int32_t* csc_arc = nullptr;  // star pricing: arc ids grouped by head (the emulation's CSC copy)
    SerialSweep(const NsxDev& dev) : d(dev) {}
    // Star pricing, serial restatement of what the sweep workers do (nsx_engine.cu, NSX_CMD_STAR / NSX_CMD_STAR_BUILD):
    // same row cache(s), same marking rule, same order of the phases.  Dantzig rule: one cache (key = reduced-cost key);
    // Devex with a single block (cmd.pad[0]): a forward and a backward cache (key = -merit), arc cmd.pad[1] left out,
    // arc cmd.pad[2] put back in.
    struct Ev { int cache; double key; int32_t arc2; };  // cache < 0: not a candidate
    Ev eval(const NsxCmd& cmd, int64_t a, uint32_t st, int32_t tl, int32_t hd, uint32_t wraw) {
        Ev ev; ev.cache = -1; ev.key = 0.0; ev.arc2 = -1;
        if (cmd.pad[0]) {
            if (a == cmd.pad[1] || (st & NSX_ARC_IN_TREE)) return ev;
            const double rc = NSX_SUB(NSX_ADD(d.pert[a], d.pi[tl]), d.pi[hd]);
            const bool fv = (st & NSX_ARC_CAN_FWD) && rc < -d.tol, bv = (st & NSX_ARC_CAN_BWD) && rc > d.tol;
            if (!(fv || bv)) return ev;
            const double w = (wraw >> 24) == cmd.wepoch ? (double)(wraw & 0xffffffu) : 1.0;
            ev.key = -NSX_DIV(NSX_MUL(rc, rc), w);
            ev.cache = fv ? 0 : 1; ev.arc2 = (int32_t)(a * 2 + (fv ? 0 : 1));
        } else {
            const double rc = NSX_SUB(NSX_ADD(nsx_phase_cost(cmd.phase, d.pert[a], a), d.pi[tl]), d.pi[hd]);
            ev.arc2 = nsx_star_candidate(a, st, rc, d.tol, &ev.key);
            ev.cache = ev.arc2 >= 0 ? 0 : -1;
        }
        return ev;
    }
    void propose(int32_t row, const Ev& ev) {
        NsxRC& r = d.rc[(size_t)ev.cache * d.n + row];
        if (nsx_rc_better(ev.key, ev.arc2, r)) { r.key = ev.key; r.arc2 = ev.arc2; }
    }
    void price_row(const NsxCmd& cmd, int32_t v, int64_t& evaluated) {
        for (int64_t a = d.row_begin[v]; a < d.row_begin[v + 1]; ++a) {
            const Ev ev = eval(cmd, a, d.state[a], v, d.head[a], d.wgt ? d.wgt[a] : 1u);
            ++evaluated;
            if (ev.cache >= 0) propose(v, ev);
        }
    }
    void run_star(const NsxCmd& cmd, NsxCand& dz, NsxDevexCand& dx, NsxCtl& c) {
        int64_t evaluated = 0;
        const int ncache = cmd.pad[0] ? 2 : 1;
        if (cmd.kind == NSX_CMD_STAR_BUILD) {
            for (int32_t v = 1; v < d.n; ++v) {
                for (int x = 0; x < ncache; ++x) { NsxRC& r = d.rc[(size_t)x * d.n + v]; r.key = 0.0; r.arc2 = -1; r.pad = 0; }
                price_row(cmd, v, evaluated);
            }
        } else {
            const int32_t ne = (int32_t)cmd.lo, round = (int32_t)cmd.hi, extra = cmd.excluded;
            const int32_t nd = (extra >= 0 && d.dstamp[extra] != round) ? ne - 1 : ne;  // (the last entry is the row of the entering arc)
            for (int32_t k = 0; k < ne; ++k) {  // the work list the pivot wrote for the sweep workers
                const int32_t* info = d.dinfo + 8 * k;
                const int32_t v = k < nd ? d.dlist[k] : extra;
                if (info[0] != v || info[1] != d.row_begin[v] || info[2] != d.row_begin[v + 1] - d.row_begin[v] ||
                    info[4] != (k < nd ? d.col_begin[v + 1] - d.col_begin[v] : 0) || (k < nd && info[3] != d.col_begin[v])) abort();
            }
            for (int32_t k = 0; k < nd; ++k) price_row(cmd, d.dlist[k], evaluated);  // (emptied by the pivot)
            if (nd < ne) price_row(cmd, extra, evaluated);
            std::vector<int32_t> rq;
            for (int32_t k = 0; k < nd; ++k) {
                const int32_t v = d.dlist[k];
                for (int32_t e = d.col_begin[v]; e < d.col_begin[v + 1]; ++e) {
                    const int64_t a = csc_arc[e];
                    const int32_t i = d.tail[a];
                    if (d.rc[i].pad == round) continue;  // the pivot emptied that row (its node is listed, or it is the row of the entering arc): priced afresh anyway
                    if (d.dstamp[i] == round || i == extra) abort();  // (the two ways of saying it agree)
                    const Ev ev = eval(cmd, a, d.csc_state[e], i, v, d.csc_wgt ? d.csc_wgt[e] : 1u);
                    ++evaluated;
                    bool queued = false;
                    for (int x = 0; x < ncache; ++x) {
                        NsxRC& cur = d.rc[(size_t)x * d.n + i];
                        if (cur.arc2 >= 0 && (cur.arc2 >> 1) == a) {  // the cached arc of the row changed its reduced cost
                            if (ev.cache == x && ev.key <= cur.key) { cur.key = ev.key; cur.arc2 = ev.arc2; }
                            else { cur.key = 0.0; cur.arc2 = -1; if (!queued) rq.push_back(i); queued = true; }
                        } else if (ev.cache == x) {
                            propose(i, ev);
                        }
                    }
                }
            }
            if (cmd.pad[0] && cmd.pad[2] >= 0) {  // Devex: the arc left out by the previous command is a candidate again
                const int64_t a = cmd.pad[2];
                const Ev ev = eval(cmd, a, d.state[a], d.tail[a], d.head[a], d.wgt[a]);
                ++evaluated;
                if (ev.cache >= 0) propose(d.tail[a], ev);
            }
            for (int32_t i : rq) price_row(cmd, i, evaluated);
            c.star_rescans += (int64_t)rq.size();
        }
        nsx_cand_init(dz);
        nsx_devex_init(dx);
        for (int32_t v = 1; v < d.n; ++v) {
            const NsxRC r = d.rc[v];
            if (cmd.pad[0]) {
                if (r.arc2 >= 0 && (dx.fi < 0 || -r.key > dx.fm || (-r.key == dx.fm && (r.arc2 >> 1) < dx.fi))) { dx.fm = -r.key; dx.fi = r.arc2 >> 1; }
                const NsxRC b = d.rc[(size_t)d.n + v];
                if (b.arc2 >= 0 && (dx.bi < 0 || -b.key > dx.bm || (-b.key == dx.bm && (b.arc2 >> 1) < dx.bi))) { dx.bm = -b.key; dx.bi = b.arc2 >> 1; }
            } else if (r.arc2 >= 0 && (dz.arc2 < 0 || r.key < dz.key || (r.key == dz.key && r.arc2 < dz.arc2))) {
                dz.key = r.key; dz.arc2 = r.arc2;
            }
        }
        c.star_evaluated = evaluated;
    }
    void inc_stats(const NsxCmd& cmd) {
        IncStats& I = g_inc;
        const int32_t n = d.n; const int64_t m = d.m;
        if (I.prev_pi.empty()) {
            I.prev_pi.assign(d.pi, d.pi + n); I.row_best.assign(n, -1); I.row_begin.assign(n + 1, 0); I.indeg.assign(n, 0);
            for (int64_t i = 0; i < m; ++i) { I.row_begin[d.tail[i] + 1]++; I.indeg[d.head[i]]++; }
            for (int32_t v = 0; v < n; ++v) I.row_begin[v + 1] += I.row_begin[v];
        } else {
            std::vector<uint8_t> dirty(n, 0);
            for (int32_t v = 0; v < n; ++v) if (d.pi[v] != I.prev_pi[v]) { dirty[v] = 1; I.dirty_nodes++; I.star_arcs += (I.row_begin[v + 1] - I.row_begin[v]) + I.indeg[v]; }
            for (int32_t v = 1; v < n; ++v) {
                if (dirty[v]) continue;
                const int32_t b = I.row_best[v];
                if (b >= 0 && (dirty[d.head[b]] || (d.state[b] & NSX_ARC_IN_TREE))) { I.rescan_rows++; I.rescan_arcs += I.row_begin[v + 1] - I.row_begin[v]; }
            }
            I.sweeps++; I.total_arcs += m;
            memcpy(I.prev_pi.data(), d.pi, (size_t)n * 8);
        }
        for (int32_t v = 1; v < n; ++v) {  // row bests of this sweep
            NsxCand k; nsx_cand_init(k);
            for (int64_t i = I.row_begin[v]; i < I.row_begin[v + 1]; ++i) {
                double rc = NSX_SUB(NSX_ADD(nsx_arc_cost(d, cmd.phase, i), d.pi[d.tail[i]]), d.pi[d.head[i]]);
                nsx_price_dantzig(k, (int32_t)i, d.state[i], rc, d.tol);
            }
            I.row_best[v] = k.arc2 >= 0 ? (k.arc2 >> 1) : -1;
        }
    }
    template <class Deferred>
    void run(const NsxCmd& cmd, NsxCand& dz, NsxDevexCand& dx, NsxCtl& c, Deferred deferred) {
        // NSX_EMU_DEFER=late leaves the deferred tree bookkeeping to the next reader of the preorder array (the engine's
        // pivot CTA does it while the workers price: either order must give the same pivots)
        const bool late = getenv("NSX_EMU_DEFER") && !strcmp(getenv("NSX_EMU_DEFER"), "late");
        if (!late) deferred();
        NSX_SYNC();
        NSX_SINGLE { run_serial(cmd, dz, dx, c); }  // the grid sweep is not what this emulation is about
        NSX_SYNC();
    }
    void run_serial(const NsxCmd& cmd, NsxCand& dz, NsxDevexCand& dx, NsxCtl& c) {
        nsx_cand_init(dz);
        nsx_devex_init(dx);
        if (cmd.kind == NSX_CMD_STAR || cmd.kind == NSX_CMD_STAR_BUILD) { run_star(cmd, dz, dx, c); return; }
        if (cmd.kind == NSX_CMD_TOPK) {  // candidate-list refresh (simplex_pricing.py:507-536)
            std::vector<std::pair<double, int32_t>> cands;
            for (int64_t i = cmd.lo; i < cmd.hi; ++i) {
                const uint8_t st = d.state[i];
                if (st & NSX_ARC_IN_TREE) continue;
                double rc = NSX_SUB(NSX_ADD(nsx_arc_cost(d, cmd.phase, i), d.pi[d.tail[i]]), d.pi[d.head[i]]);
                double merit = 0.0;
                if (((st & NSX_ARC_CAN_FWD) && rc < -d.tol) || ((st & NSX_ARC_CAN_BWD) && rc > d.tol)) merit = fabs(rc);
                if (merit > d.tol) cands.push_back({merit, (int32_t)i});
            }
            std::sort(cands.begin(), cands.end(), [](const std::pair<double, int32_t>& a, const std::pair<double, int32_t>& b) {
                return a.first > b.first || (a.first == b.first && a.second > b.second); });
            c.cl_count = (int32_t)(cands.size() < NSX_CL_SIZE ? cands.size() : NSX_CL_SIZE);
            for (int32_t k = 0; k < c.cl_count; ++k) c.cl_list[k] = cands[k].second;
            return;
        }
        if (g_inc.on && cmd.kind == NSX_CMD_DANTZIG && cmd.lo == 0 && cmd.hi == d.m) inc_stats(cmd);
        if (cmd.kind == NSX_CMD_DANTZIG) {
            // the count that lets the CUDA sweep skip the state bytes (NsxCtl::n_special) must match a recount at every sweep
            int32_t cnt = 0;
            for (int64_t i = 0; i < d.m; ++i) cnt += nsx_special(d.state[i]);
            if (getenv("NSX_EMU_DEBUG") && cnt) fprintf(stderr, "nsx_emu: sweep %lld n_special %d\n", (long long)c.sweeps, cnt);
            if (cnt != c.n_special || cmd.pad[1] != (cnt == 0 ? 1 : 0)) {
                fprintf(stderr, "nsx_emu: n_special %d, recount %d, command flag %d\n", c.n_special, cnt, cmd.pad[1]);
                abort();
            }
        }
        if (cmd.kind == NSX_CMD_DANTZIG || cmd.kind == NSX_CMD_DANTZIG_ZERO) {
            for (int64_t i = cmd.lo; i < cmd.hi; ++i) {
                double rc = NSX_SUB(NSX_ADD(nsx_arc_cost(d, cmd.phase, i), d.pi[d.tail[i]]), d.pi[d.head[i]]);
                nsx_price_dantzig(dz, (int32_t)i, d.state[i], rc, d.tol);
            }
        } else {
            for (int64_t i = cmd.lo; i < cmd.hi; ++i) {
                if ((int32_t)i == cmd.excluded) continue;
                double rc = NSX_SUB(NSX_ADD(d.pert[i], d.pi[d.tail[i]]), d.pi[d.head[i]]);
                nsx_price_devex(dx, (int32_t)i, d.state[i], rc, d.wgt[i], cmd.wepoch, d.tol);
            }
        }
    }
    void alive() {}
    void finish() {}
};

}  // namespace

static int nsx_emu_solve_impl(const nsx_problem* pb, const nsx_options* opt, const nsx_warm_start* warm, nsx_result* res) {
    const int32_t n = pb->n_nodes;
    const int64_t m = pb->n_arcs, ma = m + n - 1;
    // NSX_EMU_BLOCKED=1: blocked preorder array (what the engine uses for trees that live in HBM); NSX_EMU_BLK_LG / NSX_EMU_BLK_NB
    // shrink the blocks / the arena so that small instances split, merge and rebuild all the time
    const char* bk = getenv("NSX_EMU_BLOCKED");
    const bool blocked = bk && *bk && atoi(bk) != 0;
    int32_t blk_lg = nsx_blk_lg(n), blk_nb = NSX_BLK_MAX;
    if (blocked) {
        const char* e1 = getenv("NSX_EMU_BLK_LG"); if (e1 && *e1) blk_lg = atoi(e1);
        if (blk_lg < 1) blk_lg = 1;
        while (((n + (1 << (blk_lg - 1)) - 1) >> (blk_lg - 1)) + 8 > NSX_BLK_MAX) ++blk_lg;  // (the directory has NSX_BLK_MAX slots)
        const int32_t half = 1 << (blk_lg - 1), need = (n + half - 1) / half;
        const char* e2 = getenv("NSX_EMU_BLK_NB"); if (e2 && *e2) blk_nb = need + atoi(e2);  // (value = spare blocks beyond a half-filled layout)
        if (blk_nb > NSX_BLK_MAX) blk_nb = NSX_BLK_MAX;
        if (blk_nb < need + 4) return -7;  // knobs leave no room
    }
    std::vector<int32_t> atail(n), ahead(n), depth(n), order(blocked ? ((size_t)blk_nb << blk_lg) + n : (size_t)n), tmp(n), sidx(n), gph(n), gpt(n), garc2(2 * (size_t)n + 1);
    std::vector<double> aupper(n), flow(ma), pi(n), gres(2 * (size_t)n + 1);
    std::vector<uint8_t> state(ma);
    std::vector<uint32_t> wgt(m > 0 ? m : 1);
    std::vector<NsxNode> node(n);
    NsxDev d;
    d.n = n; d.m = m; d.ma = ma;
    d.tail = pb->tail; d.head = pb->head; d.pert = pb->pert_cost; d.upper = pb->upper;
    d.atail = atail.data(); d.ahead = ahead.data(); d.aupper = aupper.data();
    d.flow = flow.data(); d.state = state.data(); d.wgt = wgt.data();
    d.node = node.data(); d.depth = depth.data(); d.pi = pi.data(); d.pi_mirror = nullptr; d.order = order.data();
    d.tmp = tmp.data(); d.gpath_h = gph.data(); d.gpath_t = gpt.data(); d.garc2 = garc2.data();
    d.node_mask = opt->node_mask; d.imbalance = nullptr; d.gres = gres.data(); d.penalty = pb->penalty; d.tol = opt->tolerance; d.scan_walk = 0; d.par16 = nullptr; d.root_bits = nullptr;
    d.sidx = sidx.data();
    NsxBlk* blk = blocked ? new NsxBlk : nullptr;
    if (blk) { memset(blk, 0, sizeof *blk); blk->lg = blk_lg; blk->nb = blk_nb; }
    d.blk = blk;

    NsxCtl c;
    memset(&c, 0, sizeof c);
    c.n_special = -1;  // counted by nsx_solve_loop
    c.phase = 1; c.status = -1; c.maxit = opt->max_iterations;
    c.bs = opt->block_size > 0 ? opt->block_size : 1; c.pb = 0; c.last_deg = -1;
    c.ft_limit = opt->ft_update_limit; c.auto_block = opt->auto_block;
    c.pricing = opt->pricing; c.row_scan_first = opt->row_scan_first;
    c.trace_cap = res->entering_trace ? opt->trace_capacity : 0;
    c.unbounded_arc = -1;
    // NSX_EMU_STAR=1: star pricing (row cache + CSC copy), as the engine runs Dantzig pricing on multi-CTA grids
    const char* sp = getenv("NSX_EMU_STAR");
    const bool star = sp && *sp && atoi(sp) != 0 && !warm;
    std::vector<NsxRC> rcache(star ? 2 * (size_t)n : 0);
    std::vector<uint32_t> csc_wgt(star && opt->pricing == NSX_PRICING_DEVEX ? (size_t)m : 0, 1u);
    std::vector<int32_t> dinfo(star ? 8 * ((size_t)n + 1) : 0);
    std::vector<int32_t> dlist(star ? n : 0), dstamp(star ? n : 0, 0), row_begin(star ? n + 1 : 0, 0), col_begin(star ? n + 1 : 0, 0),
        csc_pos(star ? (size_t)m : 0), csc_arc(star ? (size_t)m : 0);
    std::vector<uint8_t> csc_state(star ? (size_t)m : 0);
    d.csc_wgt = nullptr; d.dinfo = nullptr; d.rc = nullptr; d.dlist = nullptr; d.dstamp = nullptr; d.row_begin = nullptr; d.col_begin = nullptr; d.csc_pos = nullptr; d.csc_state = nullptr;
    if (star) {
        for (int64_t a = 0; a + 1 < m; ++a) if (pb->tail[a] > pb->tail[a + 1]) return -8;  // rows need arcs sorted by tail
        for (int64_t a = 0; a < m; ++a) { row_begin[pb->tail[a] + 1]++; col_begin[pb->head[a] + 1]++; }
        for (int32_t v = 0; v < n; ++v) { row_begin[v + 1] += row_begin[v]; col_begin[v + 1] += col_begin[v]; }
        std::vector<int32_t> cur(col_begin.begin(), col_begin.end() - 1);
        for (int64_t a = 0; a < m; ++a) { const int32_t e = cur[pb->head[a]]++; csc_arc[e] = (int32_t)a; csc_pos[a] = e; }
        d.dinfo = dinfo.data();
        d.rc = rcache.data(); d.dlist = dlist.data(); d.dstamp = dstamp.data(); d.row_begin = row_begin.data();
        d.col_begin = col_begin.data(); d.csc_pos = csc_pos.data(); d.csc_state = csc_state.data();
        d.csc_wgt = csc_wgt.empty() ? nullptr : csc_wgt.data();
        c.star_on = opt->pricing == NSX_PRICING_DEVEX && !opt->row_scan_first ? 2 : 1;
        c.star_excl_prev = -1;
    }

    int64_t art = 0;
    std::vector<double> imbalance(warm ? (size_t)n : 0, 0.0);
    if (warm) d.imbalance = imbalance.data();
    if (warm) {  // same steps as nsx_solve_warm: host layout, copies, element-wise arc init
        int bad = nsx_warm_layout(n, m, pb->tail, pb->head, pb->supply, d.tol, warm->in_tree, node, depth, order);
        if (bad) return bad;
        memcpy(flow.data(), warm->flow, (size_t)ma * 8);
        for (int64_t a = 0; a < ma; ++a) art += nsx_init_arc_warm(d, a, pb->supply, warm->in_tree);
        pi[0] = 0.0;
        c.warm = 1;
        c.phase = warm->start_phase == 2 ? 2 : 1;
    } else {
        for (int64_t i = 0; i < m; ++i) nsx_init_real_arc(d, i);
        for (int32_t v = 0; v < n; ++v) {
            nsx_init_node(d, v, pb->supply[v]);
            if (v > 0 && flow[m + v - 1] > d.tol) art++;
        }
    }
    c.art_with_flow = art;

    { const char* e = getenv("NSX_EMU_INCSTATS"); g_inc = IncStats(); g_inc.on = e && *e && atoi(e) != 0; }
    NsxLoopShared* L = new NsxLoopShared;
    NsxPivotScratch* s = new NsxPivotScratch;
    NsxPotScratch* ps = new NsxPotScratch;
    if (star) for (int64_t a = 0; a < m; ++a) csc_state[csc_pos[a]] = state[a];  // (the engine fills it when it builds the CSC copy)
    SerialSweep sweep(d);
    sweep.csc_arc = star ? csc_arc.data() : nullptr;
#ifdef NSX_HOST_MT
    {
        const char* nt = getenv("NSX_EMU_THREADS");
        nsx_mt_nthreads = nt && *nt ? atoi(nt) : 8;
        if (nsx_mt_nthreads < 1) nsx_mt_nthreads = 1;
        pthread_barrier_init(&nsx_mt_bar, nullptr, (unsigned)nsx_mt_nthreads);
        std::vector<std::thread> team;
        for (int t = 0; t < nsx_mt_nthreads; ++t)
            team.emplace_back([&, t] { nsx_mt_tid = t; nsx_solve_loop<true>(d, c, *L, *s, *ps, res->entering_trace, sweep); });
        for (auto& th : team) th.join();
        pthread_barrier_destroy(&nsx_mt_bar);
    }
#else
    if (blk) nsx_solve_loop<true>(d, c, *L, *s, *ps, res->entering_trace, sweep);
    else nsx_solve_loop<false>(d, c, *L, *s, *ps, res->entering_trace, sweep);
#endif
    delete L; delete s; delete ps;
    if (g_inc.on && g_inc.sweeps > 0)
        fprintf(stderr, "incstats: sweeps %lld  dirty nodes/sweep %.1f  star arcs/sweep %.0f (%.2f%% of m)  rescan rows/sweep %.1f  rescan arcs/sweep %.0f (%.2f%% of m)\n",
                g_inc.sweeps, (double)g_inc.dirty_nodes / g_inc.sweeps, (double)g_inc.star_arcs / g_inc.sweeps,
                100.0 * g_inc.star_arcs / g_inc.total_arcs, (double)g_inc.rescan_rows / g_inc.sweeps,
                (double)g_inc.rescan_arcs / g_inc.sweeps, 100.0 * g_inc.rescan_arcs / g_inc.total_arcs);
    const int32_t rebuilds = blk ? blk->rebuilds : 0;
    delete blk;

    res->status = c.status;
    res->iterations = c.total;
    res->phase1_iterations = c.phase1_iterations;
    res->trace_len = c.trace_len;
    res->degenerate_pivots = c.degenerate;
    res->artificial_with_flow = c.art_after_p1;
    res->tree_updates = c.tree_updates;
    res->weight_resets = c.resets;
    res->final_block_size = c.bs;
    res->arcs_priced = c.arcs_priced;
    res->unbounded_arc = c.unbounded_arc;
    res->unbounded_rc = c.unbounded_rc;
    res->sum_cycle_len = c.sum_cycle; res->sum_subtree = c.sum_subtree; res->max_subtree = c.max_subtree;
    res->sum_rounds = c.sum_rounds; res->sum_window = c.sum_window;
    res->pricing_ms = (double)c.sum_window;  // emulation only: moved preorder entries, for design stats
    res->pivot_ms = (double)rebuilds;        // emulation only: re-layouts of the blocked preorder array
    res->sync_ms = (double)c.star_updates;   // emulation only: star-pricing updates / rebuilds / rows priced afresh
    res->exchange_ms = (double)c.star_builds;
    res->h2d_ms = (double)c.star_rescans;
    if (res->flow) memcpy(res->flow, flow.data(), ma * 8);
    if (res->potential) memcpy(res->potential, pi.data(), (size_t)n * 8);
    if (res->state) memcpy(res->state, state.data(), ma);
    return 0;
}

extern "C" int nsx_emu_solve(const nsx_problem* pb, const nsx_options* opt, nsx_result* res) {
    return nsx_emu_solve_impl(pb, opt, nullptr, res);
}
extern "C" int nsx_emu_solve_warm(const nsx_problem* pb, const nsx_options* opt, const nsx_warm_start* warm, nsx_result* res) {
    return nsx_emu_solve_impl(pb, opt, warm, res);
}

// Consistency check of the preorder representation (used by tests after a solve is not possible
// from outside; exposed separately): returns 0 when parent/pos/size/order/depth agree.
extern "C" int nsx_emu_selfcheck_enabled(void) { return 1; }
