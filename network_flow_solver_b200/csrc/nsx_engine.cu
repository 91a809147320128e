// nsx_engine.cu - B200 (sm_100a) network-simplex engine: resident pivot loop + C ABI.
//
// One persistent cooperative kernel holds the whole solve on the device:
//   * every CTA prices its share of the arc range named by the current command (streaming
//     128-bit loads of tail / head / cost / state, node potentials gathered from shared memory
//     or L2, warp-shuffle + shared-memory arg-min with lowest-index tie-break);
//   * CTA 0 merges the per-CTA candidates, runs the pivot (nsx_core.cuh: two-lane cycle walk,
//     ratio test, flow update, preorder-array tree re-hang, exact potential recompute) and
//     publishes the next command with a release store that the other CTAs acquire-poll.
// No host round trip happens between launch and the final status.
//
// Reference behaviour: see nsx_core.cuh for the pivot, and for pricing
//   DantzigPricing.select_entering_arc ............. simplex_pricing.py:97-137
//   TransportationPivotStrategy row scan ........... specialized_pivots.py:80-120
//   NetworkSimplex._select_entering_arc_vectorized . simplex.py:528-617
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <string>
#include <thread>
#include <vector>

#ifndef NSX_THREADS
#define NSX_THREADS 512
#endif
#include "nsx_core.cuh"
#include "nsx_warm.h"

#define NSX_PI_SMEM_MAX_NODES 12288  // node potentials staged in shared memory up to this many nodes

// ------------------------------------------------------------------------------------------
// Grid-wide command / arrival handshake
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long nsx_globaltimer();
// One candidate slot per sweep worker: payload, then the command sequence number it answers
// (release store).  The pivot CTA polls one slot per thread - no shared arrival counter.
struct alignas(64) NsxSlot {
    int4 v[2];      // NsxCand (16 B) or NsxDevexCand (32 B)
    int32_t seq;
    int32_t pad[7];
};
struct NsxGridCtl {
    int32_t seq;  // command sequence number, release-published by CTA 0
    int32_t abort;  // raised by a sweep worker that saw no command before its deadline (fault 4)
    unsigned long long arrived;  // heartbeat of the pivot CTA: bumped once per step of its loop, so that workers waiting for a
                                 // command can tell "the pivot CTA is busy" (rule scans, long pivots) from "it is gone"
    NsxCmd cmd;
    unsigned long long t_pub;    // globaltimer at the publication of the current command
    unsigned long long tl[8];    // handshake timeline of worker 1, ns after t_pub, accumulated over sweeps
};
struct NsxCand;
#define NSX_TL(grid, k) do { if (blockIdx.x == 1 && threadIdx.x == 0) (grid)->tl[k] += nsx_globaltimer() - (grid)->t_pub; } while (0)

__device__ __forceinline__ int32_t nsx_ld_acquire(const int32_t* p) {
    int32_t v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long nsx_ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void nsx_red_release_add(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void nsx_st_release(int32_t* p, int32_t v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 1-D bulk copies global -> shared through the TMA unit (cp.async.bulk), completion on an mbarrier.
// Candidate of a Dantzig-rule sweep, worker -> pivot CTA, as four self-validating 64-bit words {data : 32, command
// number : 32} (the "LL" idea of NCCL): the reader needs ONE round trip to L2 - no sequence word to acquire first and
// no second load for the payload - and accepts the record when all four words carry the number it waits for.  64-bit
// elements of a vector access are single-copy atomic; nothing else the reader looks at is published with this record.
__device__ __forceinline__ void nsx_ll_store(NsxSlot* sl, double key, int32_t arc2, int32_t zero2, int32_t seq) {
    const unsigned long long s = (unsigned long long)(uint32_t)seq << 32;
    const unsigned long long kb = (unsigned long long)__double_as_longlong(key);
    const unsigned long long w0 = s | (kb & 0xffffffffull), w1 = s | (kb >> 32);
    const unsigned long long w2 = s | (unsigned long long)(uint32_t)arc2, w3 = s | (unsigned long long)(uint32_t)zero2;
    asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(&sl->v[0]), "l"(w0), "l"(w1) : "memory");
    asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(&sl->v[1]), "l"(w2), "l"(w3) : "memory");
}
__device__ __forceinline__ bool nsx_ll_load(const NsxSlot* sl, int32_t seq, double& key, int32_t& arc2, int32_t& zero2) {
    unsigned long long w0, w1, w2, w3;
    asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(&sl->v[0]) : "memory");
    asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(&sl->v[1]) : "memory");
    const uint32_t s = (uint32_t)seq;
    if ((uint32_t)(w0 >> 32) != s || (uint32_t)(w1 >> 32) != s || (uint32_t)(w2 >> 32) != s || (uint32_t)(w3 >> 32) != s) return false;
    key = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
    arc2 = (int32_t)(uint32_t)w2; zero2 = (int32_t)(uint32_t)w3;
    return true;
}

__device__ __forceinline__ uint32_t nsx_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void nsx_mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nsx_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void nsx_mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// copy only: the caller has already posted the expected byte count on `bar`
__device__ __forceinline__ void nsx_bulk_copy(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(nsx_smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(nsx_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void nsx_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nsx_smem_addr(bar)), "r"(bytes) : "memory");
    nsx_bulk_copy(dst_smem, src_gmem, bytes, bar);
}
__device__ __forceinline__ void nsx_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(nsx_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool nsx_mbar_try(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(nsx_smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void nsx_mbar_wait(unsigned long long* bar, uint32_t parity) {
    while (!nsx_mbar_try(bar, parity)) { }
}
// orders earlier generic-proxy accesses (the acquired writes of other CTAs) before later
// async-proxy (TMA) reads issued by this thread
__device__ __forceinline__ void nsx_fence_proxy_async() {
    asm volatile("fence.proxy.async;" ::: "memory");
}

__device__ __forceinline__ unsigned long long nsx_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ------------------------------------------------------------------------------------------
// Candidate reductions (warp shuffle, then one shared-memory hop)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void nsx_warp_reduce(NsxCand& k) {
    __syncwarp();  // a shuffle reached by a diverged warp takes a slow path
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        NsxCand o;
        o.key = __shfl_down_sync(0xffffffffu, k.key, off);
        o.arc2 = __shfl_down_sync(0xffffffffu, k.arc2, off);
        o.zero2 = __shfl_down_sync(0xffffffffu, k.zero2, off);
        nsx_cand_merge(k, o);
    }
}
__device__ __forceinline__ void nsx_warp_reduce(NsxDevexCand& k) {
    __syncwarp();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        NsxDevexCand o;
        o.fm = __shfl_down_sync(0xffffffffu, k.fm, off);
        o.bm = __shfl_down_sync(0xffffffffu, k.bm, off);
        o.fi = __shfl_down_sync(0xffffffffu, k.fi, off);
        o.bi = __shfl_down_sync(0xffffffffu, k.bi, off);
        o.fz = __shfl_down_sync(0xffffffffu, k.fz, off);
        o.bz = __shfl_down_sync(0xffffffffu, k.bz, off);
        nsx_devex_merge(k, o);
    }
}
__device__ __forceinline__ void nsx_init(NsxCand& k) { nsx_cand_init(k); }
__device__ __forceinline__ void nsx_init(NsxDevexCand& k) { nsx_devex_init(k); }
__device__ __forceinline__ void nsx_merge(NsxCand& a, const NsxCand& b) { nsx_cand_merge(a, b); }
__device__ __forceinline__ void nsx_merge(NsxDevexCand& a, const NsxDevexCand& b) { nsx_devex_merge(a, b); }

// Block-wide reduction; the result is valid in thread 0. `buf` holds 32 entries of T in smem.
template <class T>
__device__ __forceinline__ void nsx_block_reduce(T& k, T* buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    nsx_warp_reduce(k);
    NSX_SYNC();  // buf may still be read from a previous reduction
    if (lane == 0) buf[warp] = k;
    NSX_SYNC();
    if (warp == 0) {
        if (lane < nwarp) k = buf[lane]; else nsx_init(k);
        nsx_warp_reduce(k);
    }
}

// ------------------------------------------------------------------------------------------
// Pricing store: the arrays a sweep streams, held tile-padded in HBM in the narrowest exact
// encoding the instance allows (chosen on the host side of the C ABI, see nsx_choose_layout):
//   node ids   int32, or uint16 holding id-1 when n-1 <= 65536 (real arcs never touch the root);
//   cost       float64 perturbed cost, or int32 / int16 when every cost is an integer of that
//              range (perturbation off) - converted to float64 exactly before any arithmetic;
//   state      one byte per arc (NSX_ARC_* bits), shared with the pivot code;
//   weight     uint32 epoch-tagged Devex weight (Devex solves only).
// Algorithmic bytes per arc = 2*node + cost + 1 (+4 Devex): 17 / 13 / 7 for the three layouts
// used by the BASELINE configs.
// ------------------------------------------------------------------------------------------
// The last warp of a CTA is the TMA producer of the tile ring during sweeps; the others consume.
#define NSX_CONSUMERS (NSX_THREADS - 32)
#define NSX_TILE (4 * NSX_CONSUMERS)  // arcs per tile: four arcs per consumer thread
#define NSX_MAX_STAGES 8

// NSX_NODE_U16X8 / NSX_COST_I16M1 (always together; n - 1 <= 8192 and int16 costs - BASELINE config 3): the uint16 holds
// (id - 1) * 8, i.e. the byte offset of the node's potential behind pi[1], and the int16 holds cost - 1 (the Phase-1 cost is
// (cost - 1) - 1e-6 idx, simplex.py:1162-1168) - two address multiplies and one subtraction less per arc in the hot loop.
enum { NSX_NODE_I32 = 0, NSX_NODE_U16 = 1, NSX_NODE_U16X8 = 2 };
enum { NSX_COST_F64 = 0, NSX_COST_I32 = 1, NSX_COST_I16 = 2, NSX_COST_I16M1 = 3 };

struct NsxStore {
    const unsigned char* base;  // [tiles][tail column | head column | cost column], tile_bytes each
    uint32_t tile_bytes;        // 2 * b_node + b_cost
    int32_t node_kind, cost_kind;
    uint32_t b_node, b_cost;    // bytes per tile of a node-id array / of the cost array
    uint32_t off_head, off_cost, off_state, off_wgt, stage_bytes;  // layout of one ring stage
    int32_t has_wgt;
};

static inline __host__ __device__ uint32_t nsx_node_bytes(int kind) { return kind == NSX_NODE_I32 ? 4u : 2u; }
static inline __host__ __device__ uint32_t nsx_cost_bytes(int kind) {
    return kind == NSX_COST_F64 ? 8u : kind == NSX_COST_I32 ? 4u : 2u;
}
static inline __host__ __device__ void nsx_store_layout(NsxStore& st, int node_kind, int cost_kind, int has_wgt) {
    st.node_kind = node_kind; st.cost_kind = cost_kind; st.has_wgt = has_wgt;
    st.b_node = NSX_TILE * nsx_node_bytes(node_kind);
    st.b_cost = NSX_TILE * nsx_cost_bytes(cost_kind);
    st.off_head = st.b_node;
    st.off_cost = 2 * st.b_node;
    st.tile_bytes = 2 * st.b_node + st.b_cost;
    st.off_state = st.off_cost + st.b_cost;
    st.off_wgt = st.off_state + NSX_TILE;
    st.stage_bytes = st.off_wgt + (has_wgt ? 4u * NSX_TILE : 0u);
}

// How much node state the pivot CTA keeps in shared memory: the cycle walk is a chain of
// dependent 16-byte record loads, so holding the records on-chip turns ~L2 latency per hop into
// shared-memory latency.
enum { NSX_RES_NONE = 0,   // tree in HBM/L2
       NSX_RES_NODES = 1,  // node records + potentials resident (24 B / node)
       NSX_RES_ALL = 2 };  // + depth, preorder array, permutation scratch (36 B / node)

// Dynamic shared memory of one CTA: [potentials | resident node state][ring of tile stages].
struct NsxSmemPlan {
    int32_t mode;       // NSX_RES_* (pivot CTA only)
    int32_t stage_pi;   // potentials are copied into shared memory before each sweep (TMA bulk copy)
    uint32_t ring_off;  // byte offset of the tile ring inside the dynamic part
    int32_t stages;     // ring depth (0: this CTA never sweeps)
    int32_t par16;      // pivot CTA keeps a uint16 mirror of the parent pointers (+ root bitmap) in shared memory
    int32_t blk_off;    // byte offset of the NsxBlk directory (trees in HBM: blocked preorder array), -1 = dense array
};

static inline __host__ __device__ size_t nsx_align16(size_t x) { return (x + 15) & ~(size_t)15; }
static inline __host__ __device__ size_t nsx_resident_bytes(int mode, size_t n) {
    if (mode == NSX_RES_ALL) return nsx_align16(nsx_align16(8 * n) + 16 * n + 12 * n) + 16;
    if (mode == NSX_RES_NODES) return nsx_align16(8 * n) + 16 * n;
    return 0;
}
// Plan for a CTA that pivots (and, when `sweeps`, also prices: single-CTA and batch modes).
static inline __host__ __device__ NsxSmemPlan nsx_plan_pivot(size_t n, size_t limit, uint32_t stage_bytes,
                                                             bool sweeps, int want_mode, int want_stage) {
    NsxSmemPlan p; p.mode = NSX_RES_NONE; p.stage_pi = 0; p.ring_off = 0; p.stages = 0; p.par16 = 0; p.blk_off = -1;
    const size_t ring_min = sweeps ? 2 * (size_t)stage_bytes : 0;
    const size_t blk_bytes = nsx_align16(sizeof(NsxBlk));
    // a CTA that also sweeps wants a ring of >= 4 stages more than it wants depth / preorder on-chip
    for (int want_ring = sweeps ? 4 : 0; p.mode == NSX_RES_NONE && want_ring >= (sweeps ? 2 : 0); want_ring -= 2) {
        for (int mode = want_mode; mode >= NSX_RES_NODES; --mode) {
            if (nsx_resident_bytes(mode, n) + (size_t)want_ring * stage_bytes <= limit) { p.mode = mode; break; }
        }
        if (!sweeps) break;
    }
    size_t used = nsx_resident_bytes(p.mode, n);
    if (p.mode == NSX_RES_NONE && sweeps && want_stage && nsx_align16(8 * n) + blk_bytes + ring_min <= limit) {
        p.stage_pi = 1; used = nsx_align16(8 * n);
    }
    // trees that stay in HBM: the parent pointers alone (2 B / node + 1 bit) make the cycle walk a
    // chain of shared-memory loads instead of L2 round trips
    if (p.mode == NSX_RES_NONE && !sweeps && n <= 65537 && nsx_align16(2 * n) + nsx_align16(n / 8 + 8) + blk_bytes <= limit) {
        p.par16 = 1; used = nsx_align16(2 * n) + nsx_align16(n / 8 + 8);
    }
    // ... and their preorder array is kept in blocks; the directory lives here (nsx_core.cuh, NsxBlk)
    if (p.mode == NSX_RES_NONE) { p.blk_off = (int32_t)used; used += blk_bytes; }
    p.ring_off = (uint32_t)used;
    if (sweeps) {
        size_t s = (limit - used) / stage_bytes;
        p.stages = (int32_t)(s > NSX_MAX_STAGES ? NSX_MAX_STAGES : s);
    }
    return p;
}
// Plan for a sweep-only worker CTA.
static inline __host__ __device__ NsxSmemPlan nsx_plan_worker(size_t n, size_t limit, uint32_t stage_bytes, int want_stage) {
    NsxSmemPlan p; p.mode = NSX_RES_NONE; p.stage_pi = 0; p.ring_off = 0; p.stages = 0; p.par16 = 0; p.blk_off = -1;
    size_t used = 0;
    if (want_stage && nsx_align16(8 * n) + 3 * (size_t)stage_bytes <= limit) { p.stage_pi = 1; used = nsx_align16(8 * n); }
    p.ring_off = (uint32_t)used;
    size_t s = (limit - used) / stage_bytes;
    p.stages = (int32_t)(s > NSX_MAX_STAGES ? NSX_MAX_STAGES : s);
    return p;
}
static inline __host__ __device__ size_t nsx_plan_bytes(const NsxSmemPlan& p, uint32_t stage_bytes) {
    return (size_t)p.ring_off + (size_t)p.stages * stage_bytes;
}

// Shared-memory layout of one CTA: fixed part, then (dynamic) node state and the tile ring.
struct NsxCtaShared {
    NsxLoopShared L;
    NsxCtl ctl;  // solver scalars live here during the solve (copied in / out of HBM once)
    NsxCmd cmd;  // worker copy of the command
    NsxCand dz_buf[32];
    NsxDevexCand dx_buf[32];
    unsigned long long gate_bits;               // Dantzig sweep: raw bits of the best (most negative) key any
                                                // thread of this CTA has found so far in the current sweep
    int32_t gate_arc2;                          // ... and, after the sweep, the lowest arc*2+dir among the threads that hold it
    // candidate-list refresh (NSX_CMD_TOPK): entries in the top-k buffer and the (merit bits, arc) an arc has to
    // exceed to enter it; the buffer itself aliases piv.res (keys) and piv.arc2 (arcs), idle during sweeps
    int32_t tk_cnt;
    int32_t tk_thr_idx;
    unsigned long long tk_thr_key;
    int4 x_mine[2];                             // cross-GPU exchange: this rank's record ...
    int4 x_recs[8][2];                          // ... and the records of all ranks (filled by threads 0 .. world-1)
    int32_t x_fault;
    NsxGridCtl* tl_grid;                        // handshake timeline sink (worker CTAs of the grid kernel), or null
    unsigned long long mbar;                    // completion barrier of the potentials bulk copy
    unsigned long long full[NSX_MAX_STAGES];    // tile landed in the stage (TMA complete_tx)
    unsigned long long empty[NSX_MAX_STAGES];   // every warp is done with the stage
    NsxPivotScratch piv;
    NsxPotScratch pot;
};

// Redirect the node arrays of `d` into shared memory according to the plan and fill them.
__device__ __forceinline__ NsxDev nsx_make_resident(const NsxDev& d, const NsxSmemPlan plan,
                                                    unsigned char* dyn, double** pis_out) {
    NsxDev dl = d;
    double* pis = reinterpret_cast<double*>(dyn);
    *pis_out = pis;
    dl.blk = plan.blk_off >= 0 ? reinterpret_cast<NsxBlk*>(dyn + plan.blk_off) : nullptr;
    if (dl.blk && threadIdx.x == 0) { dl.blk->lg = nsx_blk_lg(d.n); dl.blk->nb = NSX_BLK_MAX; }  // (barriers follow on every path)
    if (plan.mode == NSX_RES_NONE && plan.par16) {
        uint16_t* par = reinterpret_cast<uint16_t*>(dyn);
        uint32_t* bits = reinterpret_cast<uint32_t*>(dyn + nsx_align16(2 * (size_t)d.n));
        for (int32_t w = threadIdx.x; w < (d.n + 31) / 32; w += blockDim.x) bits[w] = 0u;
        NSX_SYNC();
        dl.par16 = par; dl.root_bits = bits;
        for (int32_t v = threadIdx.x; v < d.n; v += blockDim.x) nsx_set_parent_mirror(dl, v, d.node[v].parent);
        NSX_SYNC();
        return dl;
    }
    if (plan.mode == NSX_RES_NONE) return dl;
    size_t off = nsx_align16((size_t)d.n * 8);
    NsxNode* node_s = reinterpret_cast<NsxNode*>(dyn + off);
    off += (size_t)d.n * sizeof(NsxNode);
    for (int32_t v = threadIdx.x; v < d.n; v += blockDim.x) { node_s[v] = d.node[v]; pis[v] = d.pi[v]; }
    dl.node = node_s; dl.pi = pis; dl.pi_mirror = d.pi;
    if (plan.mode == NSX_RES_ALL) {
        int32_t* depth_s = reinterpret_cast<int32_t*>(dyn + off); off += (size_t)d.n * 4;
        int32_t* order_s = reinterpret_cast<int32_t*>(dyn + off); off += (size_t)d.n * 4;
        int32_t* tmp_s = reinterpret_cast<int32_t*>(dyn + off);
        for (int32_t v = threadIdx.x; v < d.n; v += blockDim.x) { depth_s[v] = d.depth[v]; order_s[v] = d.order[v]; }
        dl.depth = depth_s; dl.order = order_s; dl.tmp = tmp_s;
    }
    dl.scan_walk = d.n <= 32767 ? 1 : 0;  // (shared-memory trees keep a dense preorder array: cheap scatter, needed by the scan walk)
    NSX_SYNC();
    return dl;
}

// ------------------------------------------------------------------------------------------
// Tile ring: thread 0 keeps `stages` tiles in flight with TMA bulk copies (one mbarrier per
// stage counts the bytes of the 4-5 column copies of a tile); all warps consume a landed tile
// straight from shared memory and release the stage through a second mbarrier.  No registers
// hold data in flight, so the depth of the memory pipeline is set by shared memory alone.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void nsx_ring_issue(const NsxDev& d, const NsxStore& st, unsigned char* ring,
                                               NsxCtaShared& sh, uint32_t stage, int32_t tile, bool with_wgt) {
    unsigned char* dst = ring + (size_t)stage * st.stage_bytes;
    unsigned long long* bar = &sh.full[stage];
    const uint32_t total = st.tile_bytes + NSX_TILE + (with_wgt ? 4u * NSX_TILE : 0u);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nsx_smem_addr(bar)), "r"(total) : "memory");
    nsx_bulk_copy(dst, st.base + (size_t)tile * st.tile_bytes, st.tile_bytes, bar);  // tail | head | cost
    nsx_bulk_copy(dst + st.off_state, d.state + (size_t)tile * NSX_TILE, NSX_TILE, bar);
    if (with_wgt) nsx_bulk_copy(dst + st.off_wgt, d.wgt + (size_t)tile * NSX_TILE, 4u * NSX_TILE, bar);
}

// improving candidates of DantzigPricing.select_entering_arc (simplex_pricing.py:110-131)
__device__ __forceinline__ void nsx_dantzig_improving(NsxCand& k, int32_t i, uint32_t st, double rc, double tol) {
    if ((st & NSX_ARC_CAN_FWD) && rc < -tol) {
        if (k.arc2 < 0 || rc < k.key || (rc == k.key && i * 2 < k.arc2)) { k.key = rc; k.arc2 = i * 2; }
    } else if ((st & NSX_ARC_CAN_BWD) && rc > tol) {
        const double nk = -rc;
        if (k.arc2 < 0 || nk < k.key || (nk == k.key && i * 2 + 1 < k.arc2)) { k.key = nk; k.arc2 = i * 2 + 1; }
    }
}

enum { NSX_MODE_DANTZIG = 0, NSX_MODE_DEVEX = 1, NSX_MODE_DANTZIG_ZERO = 2, NSX_MODE_DEVEX_ZERO = 3, NSX_MODE_TOPK = 4 };

// One landed tile.  Thread t prices arcs t, t + T, t + 2T, t + 3T of the tile (T = block size):
// consecutive lanes take consecutive arcs, so the column reads from the stage and - on instances
// whose arcs are sorted by endpoint, such as the dense transportation case - the potential
// gathers are free of shared-memory bank conflicts.  `gate` caches the key a Dantzig candidate
// has to reach (-tol while there is none): a single compare rejects almost every arc.
// PISMEM: potentials are gathered from shared memory (`pis`), else from L2.
// Q landed tiles are priced together (Q = 2 in the steady state): 4Q independent dependency chains
// per thread hide the latency of the potential gathers and of the float64 pipe.
template <int MODE, bool PHASE1, bool PISMEM, int Q>
__device__ __forceinline__ void nsx_price_tile(const NsxDev& d, const NsxStore& st, const NsxCmd& cmd,
                                               const double* pis, const unsigned char* const (&spq)[Q],
                                               const int32_t (&tbq)[Q], int32_t lo, int32_t hi, NsxCand& dz,
                                               NsxDevexCand& dx, NsxCtaShared& sh) {
    const int tid = threadIdx.x;
    constexpr int NA = 4 * Q;  // arcs per thread
#define SPU(u) (spq[(u) >> 2])
#define OFFU(u) (((u) & 3) * NSX_CONSUMERS + tid)
#define IDXU(u) (tbq[(u) >> 2] + ((u) & 3) * NSX_CONSUMERS + tid)
    uint32_t sb[NA];
#pragma unroll
    for (int u = 0; u < NA; ++u) sb[u] = SPU(u)[st.off_state + OFFU(u)];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if (tbq[q] < lo || tbq[q] + NSX_TILE > hi) {  // ragged first / last tile of the range
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
                const int32_t i = tbq[q] + uu * NSX_CONSUMERS + tid;
                if (i < lo || i >= hi) sb[4 * q + uu] = 0;
            }
        }
    }
    // eligibility bits: residual forward / backward and not in the tree
    uint32_t any = 0;
#pragma unroll
    for (int u = 0; u < NA; ++u) {
        sb[u] = sb[u] & ((sb[u] & NSX_ARC_IN_TREE) ? 0u : (NSX_ARC_CAN_FWD | NSX_ARC_CAN_BWD));
        any |= sb[u];
    }
    if (!any) return;
    // node ids are stored zero-based from node 1 (uint16 layout) - `pi1` points at pi[1]
    int32_t tl[NA], hd[NA];
    double c[NA];
    if (st.node_kind != NSX_NODE_I32) {
        const int sh8 = st.node_kind == NSX_NODE_U16X8 ? 3 : 0;
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            tl[u] = reinterpret_cast<const uint16_t*>(SPU(u))[OFFU(u)] >> sh8;
            hd[u] = reinterpret_cast<const uint16_t*>(SPU(u) + st.off_head)[OFFU(u)] >> sh8;
        }
    } else {
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            tl[u] = reinterpret_cast<const int32_t*>(SPU(u))[OFFU(u)];
            hd[u] = reinterpret_cast<const int32_t*>(SPU(u) + st.off_head)[OFFU(u)];
        }
    }
    if (st.cost_kind == NSX_COST_F64) {
#pragma unroll
        for (int u = 0; u < NA; ++u) c[u] = reinterpret_cast<const double*>(SPU(u) + st.off_cost)[OFFU(u)];
    } else if (st.cost_kind == NSX_COST_I32) {
#pragma unroll
        for (int u = 0; u < NA; ++u) c[u] = (double)reinterpret_cast<const int32_t*>(SPU(u) + st.off_cost)[OFFU(u)];
    } else {
        const int32_t bias = st.cost_kind == NSX_COST_I16M1 ? 1 : 0;
#pragma unroll
        for (int u = 0; u < NA; ++u) c[u] = (double)((int32_t)reinterpret_cast<const int16_t*>(SPU(u) + st.off_cost)[OFFU(u)] + bias);
    }
    const double tol = d.tol;
    const double* pi1 = (PISMEM ? pis : d.pi) + (st.node_kind != NSX_NODE_I32 ? 1 : 0);
    // all four reduced costs first (independent dependency chains), decisions after
    double rc[NA];
    double i0[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) i0[q] = (double)(tbq[q] + tid);
#pragma unroll
    for (int u = 0; u < NA; ++u) {
        double cost = c[u];
        // Phase-1 tree cost  pert - 1 - 1e-6*idx  (simplex.py:1162-1168); Devex prices with the
        // perturbed Phase-2 cost in both phases.  idx as a double:
        // i0 + u*T is exact
        if (PHASE1) cost = NSX_SUB(NSX_SUB(cost, 1.0), NSX_MUL(1e-6, NSX_ADD(i0[u >> 2], (double)((u & 3) * NSX_CONSUMERS))));
        const double pt = PISMEM ? pi1[tl[u]] : __ldcg(pi1 + tl[u]);
        const double ph = PISMEM ? pi1[hd[u]] : __ldcg(pi1 + hd[u]);
        rc[u] = NSX_SUB(NSX_ADD(cost, pt), ph);
    }
    if (MODE == NSX_MODE_DANTZIG) {
        // Candidate keys are negative doubles (key <= -tol), so "key a <= key b" is "raw bits of a >=
        // raw bits of b" as unsigned integers, and a non-negative rc never passes.  The gate is the
        // best key found so far by ANY thread of the CTA (shared memory, atomicMax on the raw
        // bits): it only filters - the exact rule runs in nsx_dantzig_improving on the few arcs that
        // reach it, so a stale gate costs time, never correctness.
        const unsigned long long g = *reinterpret_cast<volatile unsigned long long*>(&sh.gate_bits);
        uint32_t hit = 0;
#pragma unroll
        for (int u = 0; u < NA; ++u)
            hit |= ((unsigned long long)__double_as_longlong(rc[u]) >= g) ? (sb[u] & NSX_ARC_CAN_FWD) : 0u;
        if (any & NSX_ARC_CAN_BWD) {  // arcs with flow to push back are rare
#pragma unroll
            for (int u = 0; u < NA; ++u)
                hit |= (((unsigned long long)__double_as_longlong(rc[u]) ^ 0x8000000000000000ull) >= g) ? (sb[u] & NSX_ARC_CAN_BWD) : 0u;
        }
        if (hit) {
            const int32_t before = dz.arc2;
            const double kbefore = dz.key;
#pragma unroll
            for (int u = 0; u < NA; ++u) nsx_dantzig_improving(dz, IDXU(u), sb[u], rc[u], tol);
            if (dz.arc2 >= 0 && (before < 0 || dz.key < kbefore))
                atomicMax(&sh.gate_bits, (unsigned long long)__double_as_longlong(dz.key));
        }
    } else if (MODE == NSX_MODE_DANTZIG_ZERO) {
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            if (sb[u] && fabs(rc[u]) <= tol) {  // simplex_pricing.py:132-135
                const int32_t cand = (IDXU(u)) * 2 + ((sb[u] & NSX_ARC_CAN_FWD) ? 0 : 1);
                if (cand < dz.zero2) dz.zero2 = cand;
            }
        }
    } else {
        uint32_t wv[NA];
#pragma unroll
        for (int u = 0; u < NA; ++u) wv[u] = 1u;
        if (MODE == NSX_MODE_DEVEX) {
#pragma unroll
            for (int u = 0; u < NA; ++u) wv[u] = reinterpret_cast<const uint32_t*>(SPU(u) + st.off_wgt)[OFFU(u)];
        }
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            const int32_t i = IDXU(u);
            const uint32_t s = sb[u];
            if (i == cmd.excluded) continue;
            const bool fv = (s & NSX_ARC_CAN_FWD) && rc[u] < -tol;
            const bool bv = (s & NSX_ARC_CAN_BWD) && rc[u] > tol;
            if (MODE == NSX_MODE_DEVEX) {
                if (fv || bv) {
                    const double wd = (wv[u] >> 24) == cmd.wepoch ? (double)(wv[u] & 0xffffffu) : 1.0;
                    const double merit = NSX_DIV(NSX_MUL(rc[u], rc[u]), wd);
                    if (fv) { if (dx.fi < 0 || merit > dx.fm || (merit == dx.fm && i < dx.fi)) { dx.fm = merit; dx.fi = i; } }
                    else    { if (dx.bi < 0 || merit > dx.bm || (merit == dx.bm && i < dx.bi)) { dx.bm = merit; dx.bi = i; } }
                }
            } else {  // NSX_MODE_DEVEX_ZERO, simplex.py:603-615
                if (!(fv || bv) && fabs(rc[u]) <= tol) {
                    if ((s & NSX_ARC_CAN_FWD) && i < dx.fz) dx.fz = i;
                    if ((s & NSX_ARC_CAN_BWD) && i < dx.bz) dx.bz = i;
                }
            }
        }
    }
}
#undef SPU
#undef OFFU
#undef IDXU

// The improving sweep of the Dantzig rule / transportation row scan - the hot loop of BASELINE config 3 - in a form
// that costs fewer issue slots per arc than the general routine above (same arithmetic, same candidates):
//  * potentials in shared memory are read through a 32-bit shared address computed once per sweep (`pi_s`; no
//    generic-pointer select and no per-arc index adjustment);
//  * eligibility is three integer operations per arc; the gate test is one float64 compare per arc (keys are negative
//    doubles: "raw bits of rc >= raw bits of gate" is "rc <= gate", and a non-negative rc never passes) accumulating the
//    eligibility bits themselves;
//  * the Phase-1 cost `(c - 1) - 1e-6 idx` (simplex.py:1162-1168) takes its `c - 1` as an integer subtraction when the
//    store holds integer costs (exact either way: |c| < 2^31);
//  * `ragged` = false: the tiles lie inside [lo, hi) (every tile of a worker but its first and last).
__device__ __forceinline__ double nsx_lds_f64(uint32_t addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
// "some rc[u] <= g": one predicate-accumulating float64 compare per value (the C++ `||` chain compiles to a min chain
// with NaN handling, eight instructions per value)
template <int NA>
__device__ __forceinline__ bool nsx_any_le(const double (&rc)[NA], double g) {
    static_assert(NA == 4 || NA == 8, "four arcs per tile and thread");
    uint32_t h;
    // pairs accumulate into separate predicates (dependent chains of length two, not NA), combined at the end
    if (NA == 4) {
        asm("{\n\t.reg .pred p, q;\n\t"
            "setp.le.f64 p, %1, %5;\n\tsetp.le.f64 q, %3, %5;\n\t"
            "setp.le.or.f64 p, %2, %5, p;\n\tsetp.le.or.f64 q, %4, %5, q;\n\t"
            "or.pred p, p, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(h) : "d"(rc[0]), "d"(rc[1]), "d"(rc[2]), "d"(rc[3]), "d"(g));
    } else {
        asm("{\n\t.reg .pred p, q, r, s;\n\t"
            "setp.le.f64 p, %1, %9;\n\tsetp.le.f64 q, %3, %9;\n\tsetp.le.f64 r, %5, %9;\n\tsetp.le.f64 s, %7, %9;\n\t"
            "setp.le.or.f64 p, %2, %9, p;\n\tsetp.le.or.f64 q, %4, %9, q;\n\tsetp.le.or.f64 r, %6, %9, r;\n\tsetp.le.or.f64 s, %8, %9, s;\n\t"
            "or.pred p, p, q;\n\tor.pred r, r, s;\n\tor.pred p, p, r;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(h) : "d"(rc[0]), "d"(rc[1]), "d"(rc[2]), "d"(rc[3]), "d"(rc[NA - 4]), "d"(rc[NA - 3]), "d"(rc[NA - 2]), "d"(rc[NA - 1]), "d"(g));
    }
    return h != 0;
}
//  * LAYOUT = 1: the store is NSX_NODE_U16X8 + NSX_COST_I16M1 (known at compile time: no layout branches; the node
//    columns ARE the byte offsets of the potentials, the cost column IS the Phase-1 `c - 1`); 0: read st.node_kind / cost_kind.
template <bool PHASE1, bool PISMEM, int Q, int LAYOUT>
__device__ __forceinline__ void nsx_price_tile_dz(const NsxDev& d, const NsxStore& st, const double* pi1, uint32_t pi_s,
                                                  const unsigned char* const (&spq)[Q], const int32_t (&tbq)[Q], bool ragged,
                                                  bool plain, int32_t lo, int32_t hi, NsxCand& dz, NsxCtaShared& sh) {
    const int tid = threadIdx.x;
    constexpr int NA = 4 * Q;
#define SPU(u) (spq[(u) >> 2])
#define OFFU(u) (((u) & 3) * NSX_CONSUMERS + tid)
#define IDXU(u) (tbq[(u) >> 2] + ((u) & 3) * NSX_CONSUMERS + tid)
    // `plain` (block-uniform): the tiles lie inside [lo, hi) and NO arc outside the tree has anything but forward
    // residual (NsxCtl::n_special == 0 - e.g. every pivot of an uncapacitated instance): the candidates are exactly the
    // arcs with rc < -tol - tree arcs have |rc| ~ 0 - so the state bytes are only looked at when an arc reaches the gate.
    uint32_t sb[NA];
    uint32_t any = 0;
    if (!plain) {
#pragma unroll
        for (int u = 0; u < NA; ++u) sb[u] = SPU(u)[st.off_state + OFFU(u)];
        if (ragged) {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                if (tbq[q] < lo || tbq[q] + NSX_TILE > hi) {
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) {
                        const int32_t i = tbq[q] + uu * NSX_CONSUMERS + tid;
                        if (i < lo || i >= hi) sb[4 * q + uu] = 0;
                    }
                }
            }
        }
        // residual forward / backward and not in the tree:  bits & 6 & ((bits & 1) - 1)
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            sb[u] &= ((sb[u] & NSX_ARC_IN_TREE) - 1u) & (NSX_ARC_CAN_FWD | NSX_ARC_CAN_BWD);
            any |= sb[u];
        }
        if (!any) return;
    }
    int32_t tl[NA], hd[NA];  // node index behind pi1 - LAYOUT 1: byte offset behind pi1
    double c[NA];
    if (LAYOUT == 1 || st.node_kind != NSX_NODE_I32) {
        const int sh8 = (LAYOUT != 1 && st.node_kind == NSX_NODE_U16X8) ? 3 : 0;
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            tl[u] = reinterpret_cast<const uint16_t*>(SPU(u))[OFFU(u)] >> sh8;
            hd[u] = reinterpret_cast<const uint16_t*>(SPU(u) + st.off_head)[OFFU(u)] >> sh8;
        }
    } else {
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            tl[u] = reinterpret_cast<const int32_t*>(SPU(u))[OFFU(u)];
            hd[u] = reinterpret_cast<const int32_t*>(SPU(u) + st.off_head)[OFFU(u)];
        }
    }
    // c = cost (Phase 2) or cost - 1 (Phase 1)
    if (LAYOUT == 1) {
#pragma unroll
        for (int u = 0; u < NA; ++u) c[u] = (double)((int32_t)reinterpret_cast<const int16_t*>(SPU(u) + st.off_cost)[OFFU(u)] + (PHASE1 ? 0 : 1));
    } else if (st.cost_kind == NSX_COST_F64) {
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            c[u] = reinterpret_cast<const double*>(SPU(u) + st.off_cost)[OFFU(u)];
            if (PHASE1) c[u] = NSX_SUB(c[u], 1.0);
        }
    } else if (st.cost_kind == NSX_COST_I32) {
#pragma unroll
        for (int u = 0; u < NA; ++u) c[u] = (double)(reinterpret_cast<const int32_t*>(SPU(u) + st.off_cost)[OFFU(u)] - (PHASE1 ? 1 : 0));
    } else {
        const int32_t adj = (st.cost_kind == NSX_COST_I16M1 ? 1 : 0) - (PHASE1 ? 1 : 0);
#pragma unroll
        for (int u = 0; u < NA; ++u) c[u] = (double)((int32_t)reinterpret_cast<const int16_t*>(SPU(u) + st.off_cost)[OFFU(u)] + adj);
    }
    double rc[NA];
    double i0[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) i0[q] = (double)(tbq[q] + tid);
#pragma unroll
    for (int u = 0; u < NA; ++u) {
        double cost = c[u];
        // idx as a double: i0 + k T is exact (and i0 + 0.0 == i0: no -0.0 here)
        if (PHASE1) cost = NSX_SUB(cost, NSX_MUL(1e-6, (u & 3) ? NSX_ADD(i0[u >> 2], (double)((u & 3) * NSX_CONSUMERS)) : i0[u >> 2]));
        double pt, ph;
        if (LAYOUT == 1) {
            pt = PISMEM ? nsx_lds_f64(pi_s + (uint32_t)tl[u]) : __ldcg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(pi1) + tl[u]));
            ph = PISMEM ? nsx_lds_f64(pi_s + (uint32_t)hd[u]) : __ldcg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(pi1) + hd[u]));
        } else {
            pt = PISMEM ? nsx_lds_f64(pi_s + 8u * (uint32_t)tl[u]) : __ldcg(pi1 + tl[u]);
            ph = PISMEM ? nsx_lds_f64(pi_s + 8u * (uint32_t)hd[u]) : __ldcg(pi1 + hd[u]);
        }
        rc[u] = NSX_SUB(NSX_ADD(cost, pt), ph);
    }
    // gate: the best key found so far by ANY thread of the CTA (shared memory, atomicMax on the raw bits of negative
    // doubles).  It only filters - the exact rule runs in nsx_dantzig_improving on the few arcs that reach it.
    const double gd = __longlong_as_double((long long)*reinterpret_cast<volatile unsigned long long*>(&sh.gate_bits));
    uint32_t hit = 0, hitb = 0;
    if (plain) {
        if (!nsx_any_le(rc, gd)) return;
        hit = NSX_ARC_CAN_FWD;
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            sb[u] = SPU(u)[st.off_state + OFFU(u)];
            sb[u] &= ((sb[u] & NSX_ARC_IN_TREE) - 1u) & (NSX_ARC_CAN_FWD | NSX_ARC_CAN_BWD);
        }
    } else {
#pragma unroll
        for (int u = 0; u < NA; ++u) hit |= rc[u] <= gd ? sb[u] : 0u;
        if (any & NSX_ARC_CAN_BWD) {  // arcs with flow to push back are rare
#pragma unroll
            for (int u = 0; u < NA; ++u) hitb |= rc[u] >= -gd ? sb[u] : 0u;
        }
    }
    if ((hit & NSX_ARC_CAN_FWD) | (hitb & NSX_ARC_CAN_BWD)) {
        const double tol = d.tol;
        const int32_t before = dz.arc2;
        const double kbefore = dz.key;
#pragma unroll
        for (int u = 0; u < NA; ++u) nsx_dantzig_improving(dz, IDXU(u), sb[u], rc[u], tol);
        if (dz.arc2 >= 0 && (before < 0 || dz.key < kbefore))
            atomicMax(&sh.gate_bits, (unsigned long long)__double_as_longlong(dz.key));
    }
}
#undef SPU
#undef OFFU
#undef IDXU

// ------------------------------------------------------------------------------------------
// Candidate-list refresh (CandidateListPricing._refresh_candidate_list, simplex_pricing.py:507-536): the
// NSX_CL_SIZE improving arcs of largest merit |rc|, ties to the LARGER arc index (Python sorts (merit, idx)
// tuples in descending order).  Each CTA keeps a buffer of NSX_TK_CAP (merit bits, arc) pairs in shared memory;
// an arc is appended when it beats the current threshold (the 100th best seen so far by this CTA); when fewer
// than one sub-step's worth of slots is left the buffer is bitonic-sorted, cut to 100 and the threshold raised.
// The per-CTA lists go to HBM and the pivot CTA merges them with the same routine.
// ------------------------------------------------------------------------------------------
#define NSX_TK_CAP 1024
struct NsxTopkOut {  // one per CTA, in HBM
    int32_t count;
    int32_t pad[3];
    unsigned long long key[NSX_CL_SIZE];
    int32_t idx[NSX_CL_SIZE];
};
struct NsxBarConsumers { __device__ __forceinline__ void operator()() const { __syncwarp(); asm volatile("bar.sync 1, %0;" ::"n"(NSX_CONSUMERS) : "memory"); } };
struct NsxBarAll { __device__ __forceinline__ void operator()() const { NSX_SYNC(); } };
// Barrier that also ORs a predicate over the participants.  Used for "is the buffer nearly full?": every thread
// reads the count after its own append, so the thread that appended last contributes the exact count and the
// OR is the same, exact answer for everybody - one barrier instead of two.
__device__ __forceinline__ bool nsx_bar_or_consumers(bool pred) {
    uint32_t out;
    __syncwarp();
    asm volatile("{\n.reg .pred p, q;\nsetp.ne.u32 q, %1, 0;\nbar.red.or.pred p, 1, %2, q;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(out) : "r"((uint32_t)pred), "n"(NSX_CONSUMERS) : "memory");
    return out != 0;
}

__device__ __forceinline__ bool nsx_tk_less(unsigned long long ka, int32_t ia, unsigned long long kb, int32_t ib) {
    return ka < kb || (ka == kb && ia < ib);
}
__device__ __forceinline__ void nsx_tk_reset(NsxCtaShared& sh, double tol) {
    sh.tk_cnt = 0;
    sh.tk_thr_key = (unsigned long long)__double_as_longlong(tol);  // merit must exceed the tolerance
    sh.tk_thr_idx = 0x7fffffff;
}
// Sort the buffer in descending (merit, arc) order, keep the best NSX_CL_SIZE, raise the threshold.
// Called by P threads (t = 0..P-1) that all passed a barrier after the last append.
template <class Bar>
__device__ __forceinline__ void nsx_tk_compact(NsxCtaShared& sh, int t, int P, Bar bar) {
    unsigned long long* key = reinterpret_cast<unsigned long long*>(sh.piv.res);
    int32_t* idx = sh.piv.arc2;
    const int32_t cnt = sh.tk_cnt;
    for (int i = t; i < NSX_TK_CAP; i += P) if (i >= cnt) { key[i] = 0ull; idx[i] = -1; }
    bar();
    for (int k = 2; k <= NSX_TK_CAP; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < NSX_TK_CAP; i += P) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long ka = key[i], kb = key[l];
                    const int32_t ia = idx[i], ib = idx[l];
                    const bool desc = (i & k) == 0;  // descending run
                    const bool swap = desc ? nsx_tk_less(ka, ia, kb, ib) : nsx_tk_less(kb, ib, ka, ia);
                    if (swap) { key[i] = kb; idx[i] = ib; key[l] = ka; idx[l] = ia; }
                }
            }
            bar();
        }
    }
    if (t == 0) {
        const int32_t keep = cnt < NSX_CL_SIZE ? cnt : NSX_CL_SIZE;
        sh.tk_cnt = keep;
        if (keep == NSX_CL_SIZE) { sh.tk_thr_key = key[NSX_CL_SIZE - 1]; sh.tk_thr_idx = idx[NSX_CL_SIZE - 1]; }
    }
    bar();
}
// Append this thread's candidate (if it beats the threshold); every lane of the warp must call it.
__device__ __forceinline__ void nsx_tk_append(NsxCtaShared& sh, bool has, unsigned long long k, int32_t i) {
    const bool pass = has && nsx_tk_less(sh.tk_thr_key, sh.tk_thr_idx, k, i);
    const unsigned mask = __ballot_sync(0xffffffffu, pass);
    if (mask) {
        const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
        int32_t base = 0;
        if (lane == leader) base = atomicAdd(&sh.tk_cnt, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pass) {
            const int32_t pos = base + __popc(mask & ((1u << lane) - 1u));
            reinterpret_cast<unsigned long long*>(sh.piv.res)[pos] = k;
            sh.piv.arc2[pos] = i;
        }
    }
}
// One landed tile in candidate-refresh mode: four sub-steps of one arc per consumer thread.
template <bool PHASE1, bool PISMEM>
__device__ __forceinline__ void nsx_topk_tile(const NsxDev& d, const NsxStore& st, const NsxCmd& cmd, const double* pis,
                                              const unsigned char* sp, int32_t tile_base, NsxCtaShared& sh) {
    const int tid = threadIdx.x;
    const double tol = d.tol;
    const double* pi1 = (PISMEM ? pis : d.pi) + (st.node_kind != NSX_NODE_I32 ? 1 : 0);
    NsxBarConsumers bar;
#pragma unroll 1
    for (int u = 0; u < 4; ++u) {
        const int off = u * NSX_CONSUMERS + tid;
        const int32_t i = tile_base + off;
        uint32_t sbits = sp[st.off_state + off];
        if ((int64_t)i < cmd.lo || (int64_t)i >= cmd.hi || (sbits & NSX_ARC_IN_TREE)) sbits = 0;
        bool has = false;
        unsigned long long key = 0ull;
        if (sbits & (NSX_ARC_CAN_FWD | NSX_ARC_CAN_BWD)) {
            int32_t tl, hd;
            double cost;
            if (st.node_kind != NSX_NODE_I32) {
                const int sh8 = st.node_kind == NSX_NODE_U16X8 ? 3 : 0;
                tl = reinterpret_cast<const uint16_t*>(sp)[off] >> sh8; hd = reinterpret_cast<const uint16_t*>(sp + st.off_head)[off] >> sh8;
            } else {
                tl = reinterpret_cast<const int32_t*>(sp)[off]; hd = reinterpret_cast<const int32_t*>(sp + st.off_head)[off];
            }
            if (st.cost_kind == NSX_COST_F64) cost = reinterpret_cast<const double*>(sp + st.off_cost)[off];
            else if (st.cost_kind == NSX_COST_I32) cost = (double)reinterpret_cast<const int32_t*>(sp + st.off_cost)[off];
            else cost = (double)((int32_t)reinterpret_cast<const int16_t*>(sp + st.off_cost)[off] + (st.cost_kind == NSX_COST_I16M1 ? 1 : 0));
            if (PHASE1) cost = NSX_SUB(NSX_SUB(cost, 1.0), NSX_MUL(1e-6, (double)i));
            const double pt = PISMEM ? pi1[tl] : __ldcg(pi1 + tl);
            const double ph = PISMEM ? pi1[hd] : __ldcg(pi1 + hd);
            const double rc = NSX_SUB(NSX_ADD(cost, pt), ph);
            if (((sbits & NSX_ARC_CAN_FWD) && rc < -tol) || ((sbits & NSX_ARC_CAN_BWD) && rc > tol)) {
                const double merit = fabs(rc);
                if (merit > tol) { has = true; key = (unsigned long long)__double_as_longlong(merit); }
            }
        }
        // (also orders the appends of the previous sub-step before this one's)
        if (nsx_bar_or_consumers(sh.tk_cnt > NSX_TK_CAP - NSX_CONSUMERS)) nsx_tk_compact(sh, tid, NSX_CONSUMERS, bar);
        nsx_tk_append(sh, has, key, i);
    }
}

// Sweep of [cmd.lo, cmd.hi) by sweeper `worker` of `nworkers`: tiles worker, worker + nworkers, ...
// of the range, ascending or (cmd.reverse) descending.  `pos` is the ring position of this CTA
// (register copy, identical in all threads): bits 0-15 = stage of the next tile, bit 16 = mbarrier
// phase parity of that stage.
template <int MODE, bool PHASE1, bool PISMEM, int LAYOUT = 0>
__device__ __forceinline__ void nsx_sweep_ring(const NsxDev& d, const NsxStore& st, const NsxCmd& cmd,
                                               double* pis, unsigned char* ring, int stages,
                                               NsxCtaShared& sh, uint32_t& pos, int worker, int nworkers,
                                               bool stage_pi, int32_t pi_n, uint32_t& stage_count, NsxCand& dz,
                                               NsxDevexCand& dx) {
    const int32_t t0 = (int32_t)(cmd.lo / NSX_TILE), t1 = (int32_t)((cmd.hi + NSX_TILE - 1) / NSX_TILE);
    const int32_t ntiles = t1 - t0;
    const int32_t my_n = ntiles > worker ? (ntiles - worker + nworkers - 1) / nworkers : 0;
    const bool with_wgt = MODE == NSX_MODE_DEVEX;
    const int lane = threadIdx.x & 31;
    const int32_t step = cmd.reverse ? -nworkers : nworkers;
    const int32_t first = t0 + worker + (cmd.reverse ? (my_n - 1) * nworkers : 0);
    uint32_t stage = pos & 0xffffu, parity = pos >> 16;
    // pi_n < 0: copy all potentials into shared memory; >= 0: the copy of the previous command is patched with pi_n entries
    const bool full_pi = stage_pi && pi_n < 0;
    if (threadIdx.x >= NSX_CONSUMERS) {
        // ---- producer warp: one lane keeps the ring full ----
        if (lane == 0) {
            // writes of the pivot CTA (state bytes, weights, potentials) were acquired through the
            // generic proxy; the bulk copies below read them through the async proxy
            nsx_fence_proxy_async();
            if (full_pi) nsx_bulk_load(pis, d.pi, (uint32_t)(((size_t)d.n * 8 + 15) & ~(size_t)15), &sh.mbar);
            int32_t tile = first;
            uint32_t s = stage, par = parity;
            for (int32_t j = 0; j < my_n; ++j) {
                if (j >= stages) nsx_mbar_wait(&sh.empty[s], par ^ 1u);  // previous tenant of the stage was read by every warp
                nsx_ring_issue(d, st, ring, sh, s, tile, with_wgt);
                tile += step;
                if (++s == (uint32_t)stages) { s = 0; par ^= 1u; }
            }
        }
    } else {
        // ---- consumer warps ----
        if (full_pi) {
            nsx_mbar_wait(&sh.mbar, stage_count & 1u);  // potentials of this sweep have landed
        } else if (stage_pi && pi_n > 0) {
            // only the potentials the last pivot changed (its re-hung subtree, NsxDev::pi_delta): patch the copy
            for (int32_t k = threadIdx.x; k < pi_n; k += NSX_CONSUMERS) {
                const int4 ent = __ldcg(reinterpret_cast<const int4*>(d.pi_delta + k));
                pis[ent.x] = __hiloint2double(ent.w, ent.z);
            }
            NsxBarConsumers bar;
            bar();
        }
        if (sh.tl_grid) NSX_TL(sh.tl_grid, 2);
        int32_t tile = first;
        const int32_t lo = (int32_t)cmd.lo, hi = (int32_t)cmd.hi;
        uint32_t s = stage, par = parity;
        int32_t j = 0;
        if (MODE == NSX_MODE_TOPK) {
            for (; j < my_n; ++j) {
                nsx_mbar_wait(&sh.full[s], par);
                nsx_topk_tile<PHASE1, PISMEM>(d, st, cmd, pis, ring + s * st.stage_bytes, tile * NSX_TILE, sh);
                __syncwarp();
                if (lane == 0) nsx_mbar_arrive(&sh.empty[s]);
                tile += step;
                if (++s == (uint32_t)stages) { s = 0; par ^= 1u; }
            }
            NsxBarConsumers bar;
            bar();
            nsx_tk_compact(sh, (int)threadIdx.x, NSX_CONSUMERS, bar);  // sorted list of this CTA
        }
        // (Dantzig improving sweep: nsx_price_tile_dz) potentials base: node ids are stored zero-based from node 1 in the uint16 layout
        const double* pi1 = (PISMEM ? pis : d.pi) + (st.node_kind != NSX_NODE_I32 ? 1 : 0);
        const uint32_t pi_s = PISMEM ? nsx_smem_addr(pi1) : 0u;
        if (MODE != NSX_MODE_TOPK && stages >= 4) {
            for (; j + 1 < my_n; j += 2) {  // two tiles per step
                uint32_t s2 = s + 1, par2 = par;
                if (s2 == (uint32_t)stages) { s2 = 0; par2 ^= 1u; }
                nsx_mbar_wait(&sh.full[s], par);
                nsx_mbar_wait(&sh.full[s2], par2);
                const unsigned char* const sp[2] = {ring + s * st.stage_bytes, ring + s2 * st.stage_bytes};
                const int32_t tb[2] = {tile * NSX_TILE, (tile + step) * NSX_TILE};
                if (MODE == NSX_MODE_DANTZIG) {
                    const bool inner = j > 0 && j + 2 < my_n;  // neither the first nor the last tile of this worker
                    nsx_price_tile_dz<PHASE1, PISMEM, 2, LAYOUT>(d, st, pi1, pi_s, sp, tb, !inner, inner && cmd.pad[1] != 0, lo, hi, dz, sh);
                }
                else nsx_price_tile<MODE, PHASE1, PISMEM, 2>(d, st, cmd, pis, sp, tb, lo, hi, dz, dx, sh);
                __syncwarp();
                if (lane == 0) { nsx_mbar_arrive(&sh.empty[s]); nsx_mbar_arrive(&sh.empty[s2]); }
                tile += 2 * step;
                s = s2 + 1; par = par2;
                if (s == (uint32_t)stages) { s = 0; par ^= 1u; }
            }
        }
        for (; j < my_n; ++j) {
            nsx_mbar_wait(&sh.full[s], par);
            const unsigned char* const sp[1] = {ring + s * st.stage_bytes};
            const int32_t tb[1] = {tile * NSX_TILE};
            if (MODE == NSX_MODE_DANTZIG) nsx_price_tile_dz<PHASE1, PISMEM, 1, LAYOUT>(d, st, pi1, pi_s, sp, tb, true, false, lo, hi, dz, sh);
            else nsx_price_tile<MODE, PHASE1, PISMEM, 1>(d, st, cmd, pis, sp, tb, lo, hi, dz, dx, sh);
            __syncwarp();
            if (lane == 0) nsx_mbar_arrive(&sh.empty[s]);
            tile += step;
            if (++s == (uint32_t)stages) { s = 0; par ^= 1u; }
        }
        if (sh.tl_grid) NSX_TL(sh.tl_grid, 3);
    }
    if (full_pi) stage_count++;
    const uint32_t adv = stage + (uint32_t)my_n;
    pos = (adv % (uint32_t)stages) | ((parity ^ ((adv / (uint32_t)stages) & 1u)) << 16);
}

template <int MODE, bool PHASE1>
__device__ __forceinline__ void nsx_sweep_ring_pi(const NsxDev& d, const NsxStore& st, const NsxCmd& cmd,
                                                  double* pis, unsigned char* ring, int stages,
                                                  NsxCtaShared& sh, uint32_t& pos, int worker, int nworkers,
                                                  bool wait_pi, int32_t pi_n, uint32_t& stage_count, NsxCand& dz,
                                                  NsxDevexCand& dx) {
    // (the improving Dantzig sweep over the pre-scaled store - config 3 - has its own instantiation without layout branches)
    if (MODE == NSX_MODE_DANTZIG && pis && st.node_kind == NSX_NODE_U16X8 && st.cost_kind == NSX_COST_I16M1)
        nsx_sweep_ring<MODE, PHASE1, true, 1>(d, st, cmd, pis, ring, stages, sh, pos, worker, nworkers, wait_pi, pi_n, stage_count, dz, dx);
    else if (pis) nsx_sweep_ring<MODE, PHASE1, true>(d, st, cmd, pis, ring, stages, sh, pos, worker, nworkers, wait_pi, pi_n, stage_count, dz, dx);
    else nsx_sweep_ring<MODE, PHASE1, false>(d, st, cmd, pis, ring, stages, sh, pos, worker, nworkers, wait_pi, pi_n, stage_count, dz, dx);
}

// One sweep of this CTA: refresh the staged potentials (optional), price, block-reduce into
// thread 0.  `stage_count` is the per-thread register copy of the number of potential copies this
// CTA has issued (its low bit is the mbarrier phase to wait for).
__device__ __forceinline__ void nsx_cta_sweep(const NsxDev& d, const NsxStore& st, const NsxCmd& cmd,
                                              double* pis, bool stage, int32_t pi_n, uint32_t& stage_count,
                                              unsigned char* ring, int stages, uint32_t& pos, int worker,
                                              int nworkers, NsxCtaShared& sh, NsxCand& dz, NsxDevexCand& dx) {
    nsx_cand_init(dz);
    nsx_devex_init(dx);
    // first candidate must satisfy rc < -tol: start the gate at the largest double below -tol
    if (threadIdx.x == 0) { sh.gate_bits = (unsigned long long)__double_as_longlong(-d.tol) + 1ull; sh.gate_arc2 = 0x7fffffff; nsx_tk_reset(sh, d.tol); }
    NSX_SYNC();  // the command is visible; reduction buffers / staged potentials are free again
    if (sh.tl_grid) NSX_TL(sh.tl_grid, 1);
    if (cmd.kind == NSX_CMD_TOPK) {
        // (tk_cnt / threshold were reset by thread 0 before the barrier above)
        if (cmd.phase == 1) nsx_sweep_ring_pi<NSX_MODE_TOPK, true>(d, st, cmd, pis, ring, stages, sh, pos, worker, nworkers, stage, pi_n, stage_count, dz, dx);
        else nsx_sweep_ring_pi<NSX_MODE_TOPK, false>(d, st, cmd, pis, ring, stages, sh, pos, worker, nworkers, stage, pi_n, stage_count, dz, dx);
        NSX_SYNC();  // the sorted top list of this CTA is in piv.res / piv.arc2, its length in tk_cnt
        return;
    }
    if (cmd.kind == NSX_CMD_DANTZIG) {
        if (cmd.phase == 1) nsx_sweep_ring_pi<NSX_MODE_DANTZIG, true>(d, st, cmd, pis, ring, stages, sh, pos, worker, nworkers, stage, pi_n, stage_count, dz, dx);
        else nsx_sweep_ring_pi<NSX_MODE_DANTZIG, false>(d, st, cmd, pis, ring, stages, sh, pos, worker, nworkers, stage, pi_n, stage_count, dz, dx);
        // The CTA's best key IS the gate (every thread raised it to its own best key; it still holds its start value when
        // nobody found a candidate): the threads that hold that key agree on the lowest arc with one shared-memory atomic -
        // two barriers instead of a shuffle tree over 16 warps.
        NSX_SYNC();
        const unsigned long long gbest = sh.gate_bits;
        if (dz.arc2 >= 0 && (unsigned long long)__double_as_longlong(dz.key) == gbest) atomicMin(&sh.gate_arc2, dz.arc2);
        NSX_SYNC();
        if (threadIdx.x == 0) {
            nsx_cand_init(dz);
            if (sh.gate_arc2 != 0x7fffffff) { dz.key = __longlong_as_double((long long)gbest); dz.arc2 = sh.gate_arc2; }
        }
    } else if (cmd.kind == NSX_CMD_DEVEX) {
        nsx_sweep_ring_pi<NSX_MODE_DEVEX, false>(d, st, cmd, pis, ring, stages, sh, pos, worker, nworkers, stage, pi_n, stage_count, dz, dx);
        nsx_block_reduce(dx, sh.dx_buf);
    } else if (cmd.kind == NSX_CMD_DANTZIG_ZERO) {
        if (cmd.phase == 1) nsx_sweep_ring_pi<NSX_MODE_DANTZIG_ZERO, true>(d, st, cmd, pis, ring, stages, sh, pos, worker, nworkers, stage, pi_n, stage_count, dz, dx);
        else nsx_sweep_ring_pi<NSX_MODE_DANTZIG_ZERO, false>(d, st, cmd, pis, ring, stages, sh, pos, worker, nworkers, stage, pi_n, stage_count, dz, dx);
        nsx_block_reduce(dz, sh.dz_buf);
    } else {
        nsx_sweep_ring_pi<NSX_MODE_DEVEX_ZERO, false>(d, st, cmd, pis, ring, stages, sh, pos, worker, nworkers, stage, pi_n, stage_count, dz, dx);
        nsx_block_reduce(dx, sh.dx_buf);
    }
}

// ------------------------------------------------------------------------------------------
// Grid-resident kernel: CTA 0 pivots, the other CTAs price.
// ------------------------------------------------------------------------------------------
struct NsxSweepCtx {       // what a CTA needs to run sweeps
    const NsxStore* st;
    double* pis;           // potentials in shared memory (resident master copy or staging buffer), or null
    bool stage;            // refresh `pis` from HBM before each sweep
    unsigned char* ring;
    int stages;
};

// ------------------------------------------------------------------------------------------
// Arc-sharded pricing across GPUs (one process per GPU).  Every GPU holds the full tree and runs
// the identical pivot; a sweep is split over the sweepers of ALL GPUs (tile t belongs to sweeper
// t mod (world * W)), so each GPU streams 1/world of the arcs.  After its local merge a pivot CTA
// stores its candidate straight into every peer's mailbox over NVLink (payload, then a
// system-scope release of the sequence word) and polls its own mailbox for the peers' candidates;
// all GPUs then merge the same `world` records with the same rule and pick the same entering arc.
// ------------------------------------------------------------------------------------------
#define NSX_MAX_WORLD 8
struct NsxMailSlot {        // 64 bytes
    int4 payload[2];        // NsxCand (16 B) or NsxDevexCand (32 B)
    unsigned long long seq; // exchange number the payload belongs to (written last, release.sys)
    unsigned long long pad[3];
};
struct NsxMailbox {
    NsxMailSlot slot[2][NSX_MAX_WORLD];  // [exchange parity][sender rank]
    unsigned long long abort;            // raised by a peer (kernel or host) that gave up: leave with fault 3
    unsigned long long pad[7];
    NsxTopkOut topk[2][NSX_MAX_WORLD];   // candidate-list refresh: the sender's sorted top list (payload of the slot)
};
struct NsxShard {
    int32_t rank, world;
    NsxMailbox* box[NSX_MAX_WORLD];  // box[r]: mailbox of rank r as mapped into this GPU's address space
};

__device__ __forceinline__ void nsx_st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long nsx_ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long nsx_ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Thread r < world of the pivot CTA handles peer r: it stores `mine` into that peer's mailbox (payload, then a
// system-scope release of the sequence word) and polls its own mailbox for the record of rank r - every peer in
// parallel, so the exchange costs one NVLink round trip plus the wait for the slowest rank, not `world` of them.
// `recs[r]` (shared memory) receives the payload of rank r, own included.  Returns 0, or the fault code when the
// deadline passed (2) or somebody raised this rank's abort word (3).
__device__ __forceinline__ int nsx_exchange_peer(const NsxShard& shd, int r, unsigned long long xseq, const int4* mine,
                                                 int nvec, int4 (*recs)[2], unsigned long long limit_ns) {
    const int par = (int)(xseq & 1ull);
    if (r == shd.rank) { recs[r][0] = mine[0]; recs[r][1] = nvec > 1 ? mine[1] : make_int4(0, 0, 0, 0); return 0; }
    NsxMailSlot* dst = &shd.box[r]->slot[par][shd.rank];
    dst->payload[0] = mine[0];
    if (nvec > 1) dst->payload[1] = mine[1];
    nsx_st_release_sys_u64(&dst->seq, xseq);
    const NsxMailbox* own = shd.box[shd.rank];
    const NsxMailSlot* src = &own->slot[par][r];
    const unsigned long long t0 = nsx_globaltimer();
    uint32_t spins = 0;
    while (nsx_ld_acquire_sys_u64(&src->seq) != xseq) {
        if ((++spins & 255u) == 0) {
            if (nsx_ld_relaxed_sys_u64(&own->abort) != 0ull) return 3;
            if (nsx_globaltimer() - t0 > limit_ns) return 2;
        }
    }
    recs[r][0] = __ldcg(&src->payload[0]);
    recs[r][1] = nvec > 1 ? __ldcg(&src->payload[1]) : make_int4(0, 0, 0, 0);
    return 0;
}
// a rank that gives up tells its peers, so they leave at once instead of waiting out their own deadline
__device__ __forceinline__ void nsx_raise_peer_aborts(const NsxShard& shd) {
    for (int r = 0; r < shd.world; ++r)
        if (r != shd.rank && shd.box[r]) nsx_st_release_sys_u64(&shd.box[r]->abort, 1ull);
}

// The sorted top list of this CTA (piv.res / piv.arc2, tk_cnt entries) becomes the candidate list.
__device__ __forceinline__ void nsx_tk_publish_list(NsxCtaShared& sh, NsxCtl& c) {
    NSX_SYNC();
    const int32_t cnt = sh.tk_cnt;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) c.cl_list[k] = sh.piv.arc2[k];
    if (threadIdx.x == 0) c.cl_count = cnt;
    NSX_SYNC();
}

// ------------------------------------------------------------------------------------------
// Star pricing on the sweep workers (see NsxRC in nsx_core.cuh for the scheme; the serial restatement the CPU tests
// check against the oracle is SerialSweep::run_star in tests/emu/nsx_emu.cpp).
//   NSX_CMD_STAR_BUILD  every worker prices the rows of its slice of the node range afresh.
//   NSX_CMD_STAR        phase A: the rows of the nodes in dlist (and the row of the entering arc) are priced afresh,
//                       the in-arcs of those nodes propose themselves to the rows of their tails, a row whose cached
//                       arc got worse is emptied and queued;  barrier over the workers;  phase B: every worker prices
//                       the queued rows of its slice afresh.
//   Both end with the minimum over the worker's slice of the row cache, delivered like a sweep candidate.
// Rows are updated with 128-bit compare-and-swap (key + arc): a proposal only ever lowers a row in the order of
// nsx_rc_better, so the result does not depend on the order in which proposals arrive.
// ------------------------------------------------------------------------------------------
struct NsxStar {
    int32_t on;
    int32_t cost_i32;           // csc_cost holds int32 (every cost an exact integer) instead of float64
    const int32_t* csc_arc;     // [m] arc ids grouped by head
    const int32_t* csc_tail;    // [m] tail of that arc
    const void* csc_cost;       // [m] its perturbed Phase-2 cost
    int32_t* rq;                // [n] rows whose cached arc got worse
    int32_t* rq_n;              // entries in rq (reset by the pivot CTA before each NSX_CMD_STAR)
    unsigned int* bar;          // arrivals at the workers' barrier, cumulative
};
#define NSX_STAR_CH 256         // arcs per work item of phase A (one warp)
#define NSX_STAR_BATCH 2048     // nodes of dlist whose work items are indexed at a time (prefix sums in shared memory)

__device__ __forceinline__ NsxRC nsx_rc_load(const NsxRC* p) {
    const int4 v = __ldcg(reinterpret_cast<const int4*>(p));
    NsxRC r;
    r.key = __longlong_as_double(((long long)(uint32_t)v.y << 32) | (uint32_t)v.x);
    r.arc2 = v.z; r.pad = v.w;
    return r;
}
__device__ __forceinline__ bool nsx_rc_cas(NsxRC* addr, const NsxRC& expect, const NsxRC& desired, NsxRC& old) {
    const unsigned long long e0 = (unsigned long long)__double_as_longlong(expect.key);
    const unsigned long long e1 = ((unsigned long long)(uint32_t)expect.pad << 32) | (uint32_t)expect.arc2;
    const unsigned long long d0 = (unsigned long long)__double_as_longlong(desired.key);
    const unsigned long long d1 = ((unsigned long long)(uint32_t)desired.pad << 32) | (uint32_t)desired.arc2;
    unsigned long long o0, o1;
    asm volatile("{\n\t.reg .b128 e, d, o;\n\tmov.b128 e, {%2, %3};\n\tmov.b128 d, {%4, %5};\n\t"
                 "atom.global.relaxed.gpu.cas.b128 o, [%6], e, d;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(o0), "=l"(o1) : "l"(e0), "l"(e1), "l"(d0), "l"(d1), "l"(addr) : "memory");
    old.key = __longlong_as_double((long long)o0); old.arc2 = (int32_t)(uint32_t)o1; old.pad = (int32_t)(o1 >> 32);
    return o0 == e0 && o1 == e1;
}
__device__ __forceinline__ NsxRC nsx_rc_none(int32_t pad) { NsxRC r; r.key = 0.0; r.arc2 = -1; r.pad = pad; return r; }
// lower the row to (key, arc2) unless it already holds something at least as good
__device__ __forceinline__ void nsx_rc_propose(NsxRC* row, double key, int32_t arc2, NsxRC cur) {
    for (;;) {
        if (!nsx_rc_better(key, arc2, cur)) return;
        NsxRC want; want.key = key; want.arc2 = arc2; want.pad = cur.pad;
        NsxRC old;
        if (nsx_rc_cas(row, cur, want, old)) return;
        cur = old;
    }
}
// One arc under the rule of the command: which cache it is a candidate for (0: the Dantzig cache / Devex forward, 1: Devex
// backward, -1: none), with its key (Dantzig: the reduced-cost key; Devex: -merit) and arc*2 + (backward).
struct NsxStarRule {
    int32_t devex, phase, excluded; uint32_t wepoch; double tol;
};
__device__ __forceinline__ int nsx_star_eval(const NsxStarRule& R, int64_t a, uint32_t st, double pert, double pt, double ph,
                                             uint32_t wraw, double& key, int32_t& arc2) {
    if (R.devex) {
        if ((st & NSX_ARC_IN_TREE) || (int32_t)a == R.excluded) return -1;
        const double rc = NSX_SUB(NSX_ADD(pert, pt), ph);  // Devex prices with the Phase-2 cost in both phases
        const bool fv = (st & NSX_ARC_CAN_FWD) && rc < -R.tol, bv = (st & NSX_ARC_CAN_BWD) && rc > R.tol;
        if (!(fv || bv)) return -1;
        const double w = (wraw >> 24) == R.wepoch ? (double)(wraw & 0xffffffu) : 1.0;
        key = -NSX_DIV(NSX_MUL(rc, rc), w);
        arc2 = (int32_t)(a * 2 + (fv ? 0 : 1));
        return fv ? 0 : 1;
    }
    const double rc = NSX_SUB(NSX_ADD(nsx_phase_cost(R.phase, pert, a), pt), ph);
    arc2 = nsx_star_candidate(a, st, rc, R.tol, &key);
    return arc2 >= 0 ? 0 : -1;
}
__device__ __forceinline__ void nsx_star_take(double k2, int32_t c2, double& key, int32_t& arc2) {
    if (c2 >= 0 && (arc2 < 0 || k2 < key || (k2 == key && c2 < arc2))) { key = k2; arc2 = c2; }
}
// Out-arcs [lo, hi) of node v, spread over `nth` threads (this thread is `t`): the thread's best candidate per cache.
__device__ __forceinline__ void nsx_star_price_row_part(const NsxDev& d, const NsxStarRule& R, double pv, int64_t lo, int64_t hi,
                                                        int t, int nth, double (&key)[2], int32_t (&arc2)[2]) {
    for (int64_t a0 = lo + t; a0 < hi; a0 += 4 * (int64_t)nth) {
        int32_t hd[4]; double ct[4]; uint32_t st[4], wr[4]; double ph[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t a = a0 + (int64_t)u * nth;
            const bool in = a < hi;
            hd[u] = in ? __ldcs(d.head + a) : 0;
            ct[u] = in ? __ldcs(d.pert + a) : 0.0;
            st[u] = in ? (uint32_t)__ldcg(d.state + a) : (uint32_t)NSX_ARC_IN_TREE;
            wr[u] = (in && R.devex) ? __ldcg(d.wgt + a) : 1u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) ph[u] = __ldcg(d.pi + hd[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t a = a0 + (int64_t)u * nth;
            double k2 = 0.0; int32_t c2 = -1;
            const int x = nsx_star_eval(R, a, st[u], ct[u], pv, ph[u], wr[u], k2, c2);
            if (x == 0) nsx_star_take(k2, c2, key[0], arc2[0]);
            else if (x == 1) nsx_star_take(k2, c2, key[1], arc2[1]);
        }
    }
}
__device__ __forceinline__ void nsx_star_warp_min(double& key, int32_t& arc2) {
    __syncwarp();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double k2 = __shfl_down_sync(0xffffffffu, key, off);
        const int32_t c2 = __shfl_down_sync(0xffffffffu, arc2, off);
        nsx_star_take(k2, c2, key, arc2);
    }
}
// Work entries of one batch in shared memory: entry q is a row (out-arcs of ent_v[q]: arc indices) or a column (in-arcs:
// indices into the CSC copy), ent_lo / ent_n its index range; pfx[q] = work items (chunks of NSX_STAR_CH) before entry q.
struct NsxStarTab {
    int32_t* v; int32_t* lo; int32_t* n; int32_t* pfx;  // [NSX_STAR_ENT], pfx [NSX_STAR_ENT + 1]; column entries carry v | 0x80000000
};
#define NSX_STAR_ENT (2 * NSX_STAR_BATCH)
// pfx[1 .. E] holds the chunk counts on entry: turn them into inclusive sums (pfx[0] = 0).  All threads of the CTA.
__device__ __forceinline__ void nsx_star_scan(int32_t* pfx, int32_t E, NsxCtaShared& sh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int32_t per_t = (E + (int32_t)blockDim.x - 1) / (int32_t)blockDim.x;
    const int32_t q0 = 1 + tid * per_t, q1 = q0 + per_t < E + 1 ? q0 + per_t : E + 1;
    int32_t sum = 0;
    for (int32_t q = q0; q < q1; ++q) sum += pfx[q];
    int32_t incl = sum;
    __syncwarp();
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int32_t o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
    int32_t* wsum = reinterpret_cast<int32_t*>(sh.dz_buf);  // (32 ints of scratch)
    if (lane == 31) wsum[warp] = incl;
    NSX_SYNC();
    if (warp == 0) {
        const int32_t w = lane < nwarp ? wsum[lane] : 0;
        int32_t wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int32_t o = __shfl_up_sync(0xffffffffu, wi, off); if (lane >= off) wi += o; }
        wsum[lane] = wi - w;
    }
    NSX_SYNC();
    int32_t run = wsum[warp] + incl - sum;
    for (int32_t q = q0; q < q1; ++q) { run += pfx[q]; pfx[q] = run; }
    NSX_SYNC();
}
// The work items of the E entries in `tab`, dealt to the warps gw, gw + GW, ... (phase A / B: all warps of all workers;
// build: the warps of this CTA).  Rows propose their chunk's best to the row cache; in-arcs follow the marking rule.
__device__ __forceinline__ void nsx_star_items(const NsxDev& d, const NsxStar& sp, const NsxStarTab& tab, int32_t E, const NsxStarRule& R,
                                               int32_t round, int gw, int GW, int64_t& evaluated) {
    const int lane = threadIdx.x & 31;
    const int ncache = R.devex ? 2 : 1;
    const int32_t total = tab.pfx[E];
    for (int32_t item = gw; item < total; item += GW) {
        int32_t lo_q = 0, hi_q = E - 1;  // entry q with pfx[q] <= item < pfx[q + 1]
        while (lo_q < hi_q) { const int32_t mid = (lo_q + hi_q + 1) >> 1; if (tab.pfx[mid] <= item) lo_q = mid; else hi_q = mid - 1; }
        const int32_t q = lo_q, c = item - tab.pfx[q];
        const int32_t vraw = tab.v[q], v = vraw & 0x7fffffff;
        const int32_t lo = tab.lo[q] + c * NSX_STAR_CH, end = tab.lo[q] + tab.n[q], hi = lo + NSX_STAR_CH < end ? lo + NSX_STAR_CH : end;
        const double pv = __ldcg(d.pi + v);
        if (vraw >= 0) {  // a chunk of the row of v
            double key[2] = {0.0, 0.0}; int32_t arc2[2] = {-1, -1};
            nsx_star_price_row_part(d, R, pv, lo, hi, lane, 32, key, arc2);
            nsx_star_warp_min(key[0], arc2[0]);
            if (R.devex) nsx_star_warp_min(key[1], arc2[1]);
            if (lane == 0) {
                if (arc2[0] >= 0) nsx_rc_propose(d.rc + v, key[0], arc2[0], nsx_rc_load(d.rc + v));
                if (R.devex && arc2[1] >= 0) nsx_rc_propose(d.rc + d.n + v, key[1], arc2[1], nsx_rc_load(d.rc + d.n + v));
                evaluated += hi - lo;
            }
        } else {          // a chunk of the in-arcs of v
            for (int32_t e0 = lo + lane; e0 < hi; e0 += 4 * 32) {
                int32_t a[4], i[4]; double ct[4]; uint32_t st[4], wr[4]; double pt[4]; NsxRC cur[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int32_t e = e0 + u * 32;
                    const bool in = e < hi;
                    a[u] = in ? __ldcs(sp.csc_arc + e) : -1;
                    i[u] = in ? __ldcs(sp.csc_tail + e) : 0;
                    ct[u] = !in ? 0.0 : sp.cost_i32 ? (double)__ldcs(reinterpret_cast<const int32_t*>(sp.csc_cost) + e)
                                                    : __ldcs(reinterpret_cast<const double*>(sp.csc_cost) + e);
                    st[u] = in ? (uint32_t)__ldcg(d.csc_state + e) : (uint32_t)NSX_ARC_IN_TREE;
                    wr[u] = (in && R.devex) ? __ldcg(d.csc_wgt + e) : 1u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) { pt[u] = __ldcg(d.pi + i[u]); cur[u] = nsx_rc_load(d.rc + i[u]); }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (a[u] < 0 || cur[u].pad == round) continue;  // (the pivot emptied that row this round: it is priced afresh anyway)
                    double key = 0.0; int32_t arc2 = -1;
                    const int x = nsx_star_eval(R, a[u], st[u], ct[u], pt[u], pv, wr[u], key, arc2);
                    bool queued = false;
                    for (int k = 0; k < ncache; ++k) {
                        NsxRC* row = d.rc + (size_t)k * d.n + i[u];
                        const NsxRC have = k == 0 ? cur[u] : nsx_rc_load(row);
                        if (have.arc2 >= 0 && (have.arc2 >> 1) == a[u]) {
                            // the cached arc of the row: still the best when it did not get worse, else the row starts over
                            NsxRC old;
                            if (x == k && key <= have.key) {
                                NsxRC want; want.key = key; want.arc2 = arc2; want.pad = have.pad;
                                if (!nsx_rc_cas(row, have, want, old)) nsx_rc_propose(row, key, arc2, old);
                            } else if (nsx_rc_cas(row, have, nsx_rc_none(have.pad), old)) {
                                if (!queued) sp.rq[atomicAdd(sp.rq_n, 1)] = i[u];
                                queued = true;
                            } else if (x == k) {
                                nsx_rc_propose(row, key, arc2, old);  // somebody lowered the row in between: it needs no fresh start
                            }
                        } else if (x == k) {
                            nsx_rc_propose(row, key, arc2, have);
                        }
                    }
                }
                if (lane == 0) evaluated += (hi - e0 < 128 ? hi - e0 : 128);
            }
        }
    }
}
// barrier over the worker CTAs (thread 0 arrives with a release and acquires the count); false when the deadline passed
__device__ __forceinline__ bool nsx_star_barrier(const NsxStar& sp, NsxCtaShared& sh, uint32_t& bar_rounds, int nworkers,
                                                 unsigned long long spin_ns) {
    NSX_SYNC();
    ++bar_rounds;
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(sp.bar) : "memory");
        const unsigned int target = bar_rounds * (unsigned int)nworkers;
        const unsigned long long t0 = nsx_globaltimer();
        uint32_t spins = 0;
        while ((unsigned int)nsx_ld_acquire(reinterpret_cast<const int32_t*>(sp.bar)) < target) {
            if ((++spins & 1023u) == 0 && nsx_globaltimer() - t0 > spin_ns) { sh.x_fault = 6; break; }
        }
        sh.tk_cnt = __ldcg(sp.rq_n);  // (tk_cnt: idle during star commands) rows queued so far, the same number on every worker
    }
    NSX_SYNC();
    return sh.x_fault != 6;
}

// One star command on a worker CTA.  Returns the worker's candidate in thread 0's dz; `evaluated` = arcs examined by
// this thread's warp (lane 0) - summed by the caller.  `fault` is set when a barrier ran past its deadline.
__device__ __forceinline__ void nsx_cta_star(const NsxDev& d, const NsxStar& sp, const NsxCmd& cmd, int worker, int nworkers,
                                             unsigned char* dyn, NsxCtaShared& sh, uint32_t& bar_rounds,
                                             unsigned long long spin_ns, NsxCand& dz, NsxDevexCand& dx, int64_t& evaluated,
                                             int32_t& fault) {
    const int tid = threadIdx.x, warp = tid >> 5, nwarp = blockDim.x >> 5;
    NsxStarRule R;
    R.devex = cmd.pad[0]; R.phase = cmd.phase; R.excluded = cmd.pad[0] ? cmd.pad[1] : -1; R.wepoch = cmd.wepoch; R.tol = d.tol;
    const int ncache = R.devex ? 2 : 1;
    if (tid == 0) sh.x_fault = 0;
    NsxStarTab tab;
    tab.v = reinterpret_cast<int32_t*>(dyn); tab.lo = tab.v + NSX_STAR_ENT; tab.n = tab.lo + NSX_STAR_ENT; tab.pfx = tab.n + NSX_STAR_ENT;
    // this worker's slice of the rows
    const int32_t per = (d.n + nworkers - 1) / nworkers;
    const int32_t r0 = worker * per < d.n ? worker * per : d.n;
    const int32_t r1 = r0 + per < d.n ? r0 + per : d.n;
    const int gw = worker * nwarp + warp, GW = nworkers * nwarp;
    NSX_SYNC();
    if (cmd.kind == NSX_CMD_STAR_BUILD) {
        // Every row of the slice (nobody else touches these rows during a build).  Short rows: an 8-lane group prices a
        // whole row - four rows in flight per warp - and stores the result directly.  Wide rows: emptied, then their chunks
        // are dealt to the warps of this CTA, which propose.
        const int32_t wide = 512;
        {
            const int lane = tid & 31, grp = lane >> 3, gl = lane & 7;
            for (int32_t v0 = r0 + warp * 4; v0 < r1; v0 += nwarp * 4) {
                const int32_t v = v0 + grp;
                int64_t lo = 0, hi = 0;
                if (v < r1) { lo = d.row_begin[v]; hi = d.row_begin[v + 1]; }
                const bool mine = v < r1 && hi - lo < wide;
                double key[2] = {0.0, 0.0}; int32_t arc2[2] = {-1, -1};
                if (mine && hi > lo) nsx_star_price_row_part(d, R, __ldcg(d.pi + v), lo, hi, gl, 8, key, arc2);
                __syncwarp();
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) {
                    const double k0 = __shfl_down_sync(0xffffffffu, key[0], off, 8), k1 = __shfl_down_sync(0xffffffffu, key[1], off, 8);
                    const int32_t c0 = __shfl_down_sync(0xffffffffu, arc2[0], off, 8), c1 = __shfl_down_sync(0xffffffffu, arc2[1], off, 8);
                    nsx_star_take(k0, c0, key[0], arc2[0]);
                    nsx_star_take(k1, c1, key[1], arc2[1]);
                }
                if (mine && gl == 0) {
                    NsxRC f; f.key = arc2[0] >= 0 ? key[0] : 0.0; f.arc2 = arc2[0]; f.pad = 0;
                    d.rc[v] = f;
                    if (R.devex) { NsxRC b; b.key = arc2[1] >= 0 ? key[1] : 0.0; b.arc2 = arc2[1]; b.pad = 0; d.rc[d.n + v] = b; }
                    evaluated += hi - lo;
                }
            }
        }
        for (int32_t v = r0 + tid; v < r1; v += blockDim.x)
            if (d.row_begin[v + 1] - d.row_begin[v] >= wide) { d.rc[v] = nsx_rc_none(0); if (R.devex) d.rc[d.n + v] = nsx_rc_none(0); }
        __threadfence();  // (rare command) the plain stores are in L2 before any proposal - an L2 atomic - of another thread
        for (int32_t b0 = r0; b0 < r1; b0 += NSX_STAR_ENT) {
            const int32_t E = r1 - b0 < NSX_STAR_ENT ? r1 - b0 : NSX_STAR_ENT;
            NSX_SYNC();
            for (int32_t q = tid; q < E; q += blockDim.x) {
                const int32_t rb = d.row_begin[b0 + q], re = d.row_begin[b0 + q + 1];
                tab.v[q] = b0 + q; tab.lo[q] = rb; tab.n[q] = re - rb;
                tab.pfx[q + 1] = re - rb >= wide ? (re - rb + NSX_STAR_CH - 1) / NSX_STAR_CH : 0;
            }
            if (tid == 0) tab.pfx[0] = 0;
            NSX_SYNC();
            nsx_star_scan(tab.pfx, E, sh);
            nsx_star_items(d, sp, tab, E, R, 0, warp, nwarp, evaluated);
        }
    } else {
        const int32_t ne = (int32_t)cmd.lo, round = (int32_t)cmd.hi;
        const int4* info = reinterpret_cast<const int4*>(d.dinfo);
        // ---- phase A: rows and in-arcs of the listed nodes (the pivot CTA wrote their index ranges to dinfo) ----
        for (int32_t b0 = 0; b0 < ne; b0 += NSX_STAR_BATCH) {
            const int32_t bn = ne - b0 < NSX_STAR_BATCH ? ne - b0 : NSX_STAR_BATCH;
            NSX_SYNC();
            for (int32_t k = tid; k < bn; k += blockDim.x) {
                const int4 x = __ldcg(info + 2 * (b0 + k)), y = __ldcg(info + 2 * (b0 + k) + 1);  // {v, row lo, row n, col lo} {col n, ...}
                tab.v[2 * k] = x.x; tab.lo[2 * k] = x.y; tab.n[2 * k] = x.z; tab.pfx[2 * k + 1] = (x.z + NSX_STAR_CH - 1) / NSX_STAR_CH;
                tab.v[2 * k + 1] = x.x | (int32_t)0x80000000; tab.lo[2 * k + 1] = x.w; tab.n[2 * k + 1] = y.x;
                tab.pfx[2 * k + 2] = (y.x + NSX_STAR_CH - 1) / NSX_STAR_CH;
            }
            if (tid == 0) tab.pfx[0] = 0;
            NSX_SYNC();
            nsx_star_scan(tab.pfx, 2 * bn, sh);
            nsx_star_items(d, sp, tab, 2 * bn, R, round, gw, GW, evaluated);
        }
        if (R.devex && cmd.pad[2] >= 0 && gw == 0 && (tid & 31) == 0) {
            // Devex: the arc the previous command left out (last degenerate arc) is a candidate again
            const int64_t a = cmd.pad[2];
            const int32_t tl = d.tail[a], hd = d.head[a];
            double key = 0.0; int32_t arc2 = -1;
            const int x = nsx_star_eval(R, a, (uint32_t)__ldcg(d.state + a), d.pert[a], __ldcg(d.pi + tl), __ldcg(d.pi + hd), __ldcg(d.wgt + a), key, arc2);
            if (x >= 0) { NsxRC* row = d.rc + (size_t)x * d.n + tl; nsx_rc_propose(row, key, arc2, nsx_rc_load(row)); }
            evaluated += 1;
        }
        if (sh.tl_grid) NSX_TL(sh.tl_grid, 2);
        // ---- every proposal of phase A is in the row cache (L2 atomics; the CTA barrier plus thread 0's release publish
        // them to whoever acquires the counter - a per-thread fence here costs tens of microseconds) ----
        if (!nsx_star_barrier(sp, sh, bar_rounds, nworkers, spin_ns)) fault = 6;
        if (sh.tl_grid) NSX_TL(sh.tl_grid, 3);
        // ---- phase B: the queued rows (emptied when they were queued) are priced afresh by all workers ----
        const int32_t nq = sh.tk_cnt;
        if (nq > 0 && !fault) {
            for (int32_t b0 = 0; b0 < nq; b0 += NSX_STAR_ENT) {
                const int32_t E = nq - b0 < NSX_STAR_ENT ? nq - b0 : NSX_STAR_ENT;
                NSX_SYNC();
                for (int32_t q = tid; q < E; q += blockDim.x) {
                    const int32_t v = __ldcg(sp.rq + b0 + q);
                    const int32_t rb = d.row_begin[v], re = d.row_begin[v + 1];
                    tab.v[q] = v; tab.lo[q] = rb; tab.n[q] = re - rb; tab.pfx[q + 1] = (re - rb + NSX_STAR_CH - 1) / NSX_STAR_CH;
                }
                if (tid == 0) tab.pfx[0] = 0;
                NSX_SYNC();
                nsx_star_scan(tab.pfx, E, sh);
                nsx_star_items(d, sp, tab, E, R, round, gw, GW, evaluated);
            }
            if (!nsx_star_barrier(sp, sh, bar_rounds, nworkers, spin_ns)) fault = 6;
        }
    }
    // ---- minimum over this worker's slice of the row cache ----
    NSX_SYNC();
    if (sh.tl_grid) NSX_TL(sh.tl_grid, 4);
    nsx_cand_init(dz);
    nsx_devex_init(dx);
    for (int32_t v = r0 + tid; v < r1; v += blockDim.x) {
        const NsxRC r = nsx_rc_load(d.rc + v);
        if (!R.devex) {
            if (r.arc2 >= 0 && (dz.arc2 < 0 || r.key < dz.key || (r.key == dz.key && r.arc2 < dz.arc2))) { dz.key = r.key; dz.arc2 = r.arc2; }
        } else {
            if (r.arc2 >= 0 && (dx.fi < 0 || -r.key > dx.fm || (-r.key == dx.fm && (r.arc2 >> 1) < dx.fi))) { dx.fm = -r.key; dx.fi = r.arc2 >> 1; }
            const NsxRC b = nsx_rc_load(d.rc + d.n + v);
            if (b.arc2 >= 0 && (dx.bi < 0 || -b.key > dx.bm || (-b.key == dx.bm && (b.arc2 >> 1) < dx.bi))) { dx.bm = -b.key; dx.bi = b.arc2 >> 1; }
        }
    }
    if (R.devex) nsx_block_reduce(dx, sh.dx_buf); else nsx_block_reduce(dz, sh.dz_buf);
    (void)ncache;
}

// Sweep functor of CTA 0.
struct GridSweep {
    const NsxDev& d;      // global view (state bytes, weights, global potentials)
    NsxGridCtl* g;
    NsxSlot* slots;
    NsxTopkOut* topk;
    NsxCtaShared& sh;
    NsxSweepCtx cx;
    uint32_t& stage_count;
    uint32_t& q0;
    int32_t seq;
    unsigned long long target;
    unsigned long long t_price, t_sync;
    const NsxShard& shd;
    unsigned long long xseq, t_xchg;
    unsigned long long spin_ns;  // deadline of every wait in this functor
    const NsxStar& star;
    int timeline;   // accumulate the handshake timeline (two global read-modify-writes per sweep on the critical path)
    unsigned long long beat;

    // Candidates of the other GPUs.  Called by ALL threads of the pivot CTA; thread 0 holds the local best in kz / kx
    // and receives the merged best.  A deadline or a raised abort word ends in c.fault (and tells the peers).
    __device__ __forceinline__ void exchange_all(bool devex, NsxCand& kz, NsxDevexCand& kx, NsxCtl& c) {
        ++xseq;  // every thread keeps its own copy of the exchange number
        if (threadIdx.x == 0) {
            if (devex) { union { NsxDevexCand c; int4 v[2]; } m; m.c = kx; sh.x_mine[0] = m.v[0]; sh.x_mine[1] = m.v[1]; }
            else { union { NsxCand c; int4 v; } m; m.c = kz; sh.x_mine[0] = m.v; sh.x_mine[1] = make_int4(0, 0, 0, 0); }
            sh.x_fault = 0;
        }
        NSX_SYNC();
        unsigned long long t1 = 0;
        if (threadIdx.x == 0) t1 = nsx_globaltimer();
        if ((int)threadIdx.x < shd.world) {
            const int f = nsx_exchange_peer(shd, (int)threadIdx.x, xseq, sh.x_mine, devex ? 2 : 1, sh.x_recs, spin_ns);
            if (f) atomicMax(&sh.x_fault, f);
        }
        NSX_SYNC();
        if (threadIdx.x == 0) {
            t_xchg += nsx_globaltimer() - t1;
            if (sh.x_fault) { c.fault = sh.x_fault; nsx_raise_peer_aborts(shd); }
            else if (devex) {
                nsx_devex_init(kx);
                for (int r = 0; r < shd.world; ++r) { union { NsxDevexCand c; int4 v[2]; } o; o.v[0] = sh.x_recs[r][0]; o.v[1] = sh.x_recs[r][1]; nsx_devex_merge(kx, o.c); }
            } else {
                nsx_cand_init(kz);
                for (int r = 0; r < shd.world; ++r) { union { NsxCand c; int4 v; } o; o.v = sh.x_recs[r][0]; nsx_cand_merge(kz, o.c); }
            }
        }
    }

    __device__ __forceinline__ void publish(const NsxCmd& cmd, int32_t pi_n = -1) {
        if (threadIdx.x == 0) {
            union { NsxCmd c; int4 v[3]; } tmp;
            tmp.c = cmd;
            tmp.c.pad[2] = pi_n;
            int4* dst = reinterpret_cast<int4*>(&g->cmd);
            dst[0] = tmp.v[0]; dst[1] = tmp.v[1]; dst[2] = tmp.v[2];
            g->t_pub = nsx_globaltimer();
            // release is cumulative: it orders the command words above and every pivot write the other
            // threads of this CTA made before the preceding barrier
            nsx_st_release(&g->seq, seq + 1);
        }
        ++seq;  // every thread keeps its own copy of the sequence number (the slot polls compare against it)
    }
    // Candidate-list refresh across GPUs.  On entry the sorted top list of THIS rank is in piv.res / piv.arc2 (tk_cnt
    // entries, all threads past a barrier).  Every rank stores its list into every peer's mailbox (parity-buffered like
    // the candidate slots), the slot sequence words signal completion, and all ranks merge the same `world` lists with
    // the same total order (merit, arc) - so they end up with the same candidate list.
    __device__ __forceinline__ void exchange_topk(NsxCtl& c) {
        ++xseq;
        const int par = (int)(xseq & 1ull);
        const int32_t cnt = sh.tk_cnt;
        const unsigned long long* key = reinterpret_cast<const unsigned long long*>(sh.piv.res);
        const int32_t* idx = sh.piv.arc2;
        for (int r = 0; r < shd.world; ++r) {
            NsxTopkOut* dst = &shd.box[r]->topk[par][shd.rank];
            for (int k = threadIdx.x; k < cnt; k += blockDim.x) { dst->key[k] = key[k]; dst->idx[k] = idx[k]; }
            if (threadIdx.x == 0) dst->count = cnt;
        }
        __threadfence_system();  // every storing thread: its list entries are visible system-wide before the barrier
        if (threadIdx.x == 0) { sh.x_mine[0] = make_int4(cnt, 0, 0, 0); sh.x_mine[1] = make_int4(0, 0, 0, 0); sh.x_fault = 0; }
        NSX_SYNC();
        unsigned long long t1 = 0;
        if (threadIdx.x == 0) t1 = nsx_globaltimer();
        if ((int)threadIdx.x < shd.world) {
            const int f = nsx_exchange_peer(shd, (int)threadIdx.x, xseq, sh.x_mine, 1, sh.x_recs, spin_ns);
            if (f) atomicMax(&sh.x_fault, f);
        }
        NSX_SYNC();
        if (threadIdx.x == 0) {
            t_xchg += nsx_globaltimer() - t1;
            if (sh.x_fault) { c.fault = sh.x_fault; nsx_raise_peer_aborts(shd); }
            nsx_tk_reset(sh, d.tol);
        }
        NSX_SYNC();
        if (c.fault) return;
        NsxBarAll bar;
        const NsxMailbox* own = shd.box[shd.rank];
        for (int r = 0; r < shd.world; ++r) {  // <= 8 lists of <= 100 entries: the buffer (1024) holds them all
            const NsxTopkOut* src = &own->topk[par][r];
            const int32_t n = __ldcg(&src->count);
            for (int k0 = 0; k0 < NSX_CL_SIZE; k0 += (int)blockDim.x) {
                const int k = k0 + (int)threadIdx.x;
                const bool has = k < n;
                nsx_tk_append(sh, has, has ? __ldcg(&src->key[k]) : 0ull, has ? __ldcg(&src->idx[k]) : -1);
            }
        }
        bar();
        nsx_tk_compact(sh, (int)threadIdx.x, (int)blockDim.x, bar);
    }
    // candidate-list refresh: merge the sorted per-CTA lists of the workers (HBM) into the final list
    __device__ __forceinline__ void merge_topk(NsxCtl& c) {
        if (threadIdx.x == 0) nsx_tk_reset(sh, d.tol);
        NsxBarAll bar;
        const int P = (int)blockDim.x, t = (int)threadIdx.x;
        for (int b0 = 1; b0 < (int)gridDim.x; b0 += 4) {  // four lists (<= 400 entries) per step
            __syncwarp();
            if (__syncthreads_or(sh.tk_cnt > NSX_TK_CAP - 4 * NSX_CL_SIZE)) nsx_tk_compact(sh, t, P, bar);
            bool has = false; unsigned long long k = 0ull; int32_t i = -1;
            if (t < 4 * NSX_CL_SIZE) {
                const int b = b0 + t / NSX_CL_SIZE, e = t % NSX_CL_SIZE;
                if (b < (int)gridDim.x && e < __ldcg(&topk[b].count)) {
                    has = true; k = __ldcg(&topk[b].key[e]); i = __ldcg(&topk[b].idx[e]);
                }
            }
            nsx_tk_append(sh, has, k, i);
        }
        bar();
        nsx_tk_compact(sh, t, P, bar);
        if (shd.world > 1) { exchange_topk(c); if (c.fault) return; }
        nsx_tk_publish_list(sh, c);
    }
    // `deferred`: tree bookkeeping of the last pivot that pricing does not need (nsx_pivot_flush) - run by this CTA while
    // the workers price
    template <class Deferred>
    __device__ __forceinline__ void run(const NsxCmd& cmd_in, NsxCand& out_dz, NsxDevexCand& out_dx, NsxCtl& c, Deferred deferred) {
        unsigned long long t0 = 0;
        if (threadIdx.x == 0) t0 = nsx_globaltimer();
        if (gridDim.x == 1) deferred();
        NSX_SYNC();  // pivot writes of all threads precede thread 0's fence + release
        const NsxCmd cmd = cmd_in;
        // potentials: how many entries of pi_delta bring a worker's shared-memory copy up to date (-1: copy everything);
        // travels in pad[2] of the published command
        const int32_t pi_n = (gridDim.x > 1 && d.pi_delta) ? sh.piv.pi_delta_n : -1;
        const bool starcmd = cmd.kind == NSX_CMD_STAR || cmd.kind == NSX_CMD_STAR_BUILD;
        const bool devex = cmd.kind == NSX_CMD_DEVEX || cmd.kind == NSX_CMD_DEVEX_ZERO || (starcmd && cmd.pad[0]);
        if (gridDim.x == 1) {  // alone: this CTA prices everything itself
            NsxCand dz; NsxDevexCand dx;
            if (threadIdx.x == 0) __threadfence();
            nsx_cta_sweep(d, *cx.st, cmd, cx.pis, cx.stage, -1, stage_count, cx.ring, cx.stages, q0, shd.rank, shd.world, sh, dz, dx);
            if (cmd.kind == NSX_CMD_TOPK) {
                if (shd.world > 1) { exchange_topk(c); if (c.fault) return; }
                nsx_tk_publish_list(sh, c);
                if (threadIdx.x == 0) t_price += nsx_globaltimer() - t0;
                return;
            }
            if (shd.world > 1) exchange_all(devex, dz, dx, c);
            if (threadIdx.x == 0) {
                if (devex) out_dx = dx; else out_dz = dz;
                t_price += nsx_globaltimer() - t0;
            }
            NSX_SYNC();
            return;
        }
        if (starcmd && threadIdx.x == 0) { *star.rq_n = 0; *reinterpret_cast<unsigned long long*>(&sh.x_recs[0][0]) = 0ull; }  // (x_recs: idle on one GPU, holds the evaluated-arc count)
        if (starcmd) NSX_SYNC();
        publish(cmd, pi_n);
        if (threadIdx.x == 0) sh.piv.pi_delta_n = 0;  // (every thread read it before the barrier in front of publish)
        deferred();
        unsigned long long t1 = 0;
        if (threadIdx.x == 0) t1 = nsx_globaltimer();
        // every worker's candidate: thread b polls slot b until it carries this command's number
        NsxDevexCand kx; nsx_devex_init(kx);
        NsxCand kz; nsx_cand_init(kz);
        unsigned long long ev = 0ull;
        int32_t nq_seen = 0;  // star update: rows that were queued, as reported by the worker whose slot this thread polls
        const bool ll = cmd.kind == NSX_CMD_DANTZIG || cmd.kind == NSX_CMD_DANTZIG_ZERO;  // (nsx_ll_store on the worker side)
        for (int b = 1 + threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
            const NsxSlot* sl = slots + b;
            uint32_t spins = 0;
            bool lost = false;
            const unsigned long long t_wait = nsx_globaltimer();
            if (ll) {
                NsxCand got; nsx_cand_init(got);
                while (!nsx_ll_load(sl, seq, got.key, got.arc2, got.zero2)) {
                    if ((++spins & 1023u) == 0 && (nsx_ld_acquire(&g->abort) != 0 || nsx_globaltimer() - t_wait > spin_ns)) { lost = true; break; }
                }
                if (lost) { c.fault = 1; break; }
                nsx_cand_merge(kz, got);
                continue;
            }
            while (nsx_ld_acquire(&sl->seq) != seq) {
                if ((++spins & 1023u) == 0 && (nsx_ld_acquire(&g->abort) != 0 || nsx_globaltimer() - t_wait > spin_ns)) { lost = true; break; }
            }
            if (lost) { c.fault = 1; break; }  // (same value from every thread that gives up)
            if (devex) {
                union { NsxDevexCand c; int4 v[2]; } tmp;
                tmp.v[0] = __ldcg(&sl->v[0]); tmp.v[1] = __ldcg(&sl->v[1]);
                nsx_devex_merge(kx, tmp.c);
                if (starcmd) { ev += ((unsigned long long)(uint32_t)__ldcg(&sl->pad[1]) << 32) | (uint32_t)__ldcg(&sl->pad[0]); if (__ldcg(&sl->pad[2])) c.fault = __ldcg(&sl->pad[2]); nq_seen = __ldcg(&sl->pad[3]); }
            } else {
                union { NsxCand c; int4 v; } tmp;
                tmp.v = __ldcg(&sl->v[0]);
                nsx_cand_merge(kz, tmp.c);
                if (starcmd) { ev += ((unsigned long long)(uint32_t)__ldcg(&sl->pad[1]) << 32) | (uint32_t)__ldcg(&sl->pad[0]); if (__ldcg(&sl->pad[2])) c.fault = __ldcg(&sl->pad[2]); nq_seen = __ldcg(&sl->pad[3]); }
            }
        }
        if (starcmd && ev) atomicAdd(reinterpret_cast<unsigned long long*>(&sh.x_recs[0][0]), ev);
        NSX_SYNC();
        if (c.fault) {  // a worker never answered (block-uniform after the barrier): tell the peers, the loop ends
            if (threadIdx.x == 0 && shd.world > 1) nsx_raise_peer_aborts(shd);
            return;
        }
        if (threadIdx.x == 0) { t_sync += nsx_globaltimer() - t1; if (timeline) g->tl[6] += nsx_globaltimer() - g->t_pub; }
        if (starcmd && threadIdx.x == 0) {
            c.star_evaluated = (int64_t)*reinterpret_cast<unsigned long long*>(&sh.x_recs[0][0]);
            if (cmd.kind == NSX_CMD_STAR) c.star_rescans += nq_seen;  // (thread 0 polled worker 1's slot; the count is the same on every worker)
        }
        if (cmd.kind == NSX_CMD_TOPK) {
            merge_topk(c);
            if (threadIdx.x == 0) t_price += nsx_globaltimer() - t0;
            return;
        }
        if (devex) nsx_block_reduce(kx, sh.dx_buf); else nsx_block_reduce(kz, sh.dz_buf);
        if (shd.world > 1) exchange_all(devex, kz, kx, c);
        if (threadIdx.x == 0) { if (devex) out_dx = kx; else out_dz = kz; }
        if (threadIdx.x == 0) { t_price += nsx_globaltimer() - t0; if (timeline) g->tl[7] += nsx_globaltimer() - g->t_pub; }
        NSX_SYNC();
    }
    // once per step of the pivot loop: tells the waiting workers that this CTA is alive
    __device__ __forceinline__ void alive() {
        if (threadIdx.x == 0 && gridDim.x > 1) *reinterpret_cast<volatile unsigned long long*>(&g->arrived) = ++beat;
    }
    __device__ __forceinline__ void finish() {
        NsxCmd cmd;
        cmd.kind = NSX_CMD_EXIT; cmd.phase = 0; cmd.lo = cmd.hi = 0; cmd.excluded = -1; cmd.wepoch = 0;
        cmd.reverse = 0; cmd.pad[0] = cmd.pad[1] = cmd.pad[2] = 0;
        NSX_SYNC();
        if (gridDim.x > 1) publish(cmd);
    }
};

static_assert(sizeof(NsxCand) == 16, "NsxCand is moved as one int4");
static_assert(sizeof(NsxCmd) == 48, "NsxCmd is moved as three int4");
static_assert(offsetof(NsxGridCtl, cmd) % 16 == 0, "command block must be 16-byte aligned");
static_assert(sizeof(NsxDevexCand) == 32, "NsxDevexCand is moved as two int4");

struct NsxKernelArgs {
    NsxDev d;
    NsxStore st;
    NsxCtl* ctl;
    NsxGridCtl* grid;
    NsxSlot* slots;
    NsxTopkOut* topk;
    int32_t* trace;
    NsxSmemPlan plan;    // CTA 0
    NsxSmemPlan wplan;   // sweep workers
    int32_t probe_sweeps;  // > 0: measurement aid, run this many sweeps of the initial state and stop
    NsxShard shard;        // world == 1: single GPU
    NsxStar star;          // star pricing (on == 0: full sweeps)
    int32_t timeline;      // measurement aid: record the handshake timeline (nsx_result.handshake_ns); sweep probes and NSX_TIMELINE=1
    unsigned long long spin_ns;  // deadline of device-side waits (nsx_options.spin_timeout_ms)
};

__device__ __forceinline__ void nsx_copy_ctl(NsxCtl* dst, const NsxCtl* src) {
    const int32_t* s = reinterpret_cast<const int32_t*>(src);
    int32_t* t = reinterpret_cast<int32_t*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(NsxCtl) / 4); i += blockDim.x) t[i] = s[i];
}

__device__ __forceinline__ void nsx_init_barriers(NsxCtaShared& sh) {
    if (threadIdx.x == 0) {
        sh.tl_grid = nullptr;
        nsx_mbar_init(&sh.mbar, 1);
        for (int s = 0; s < NSX_MAX_STAGES; ++s) {
            nsx_mbar_init(&sh.full[s], 1);
            nsx_mbar_init(&sh.empty[s], NSX_CONSUMERS / 32);
        }
        nsx_mbar_init_fence();
    }
}

// Measurement aid (nsx_sweep_probe): `count` sweeps of the initial state through exactly the
// command / arrival protocol of a solve, no pivots.
template <bool BLK, class Sweep>
__device__ __forceinline__ void nsx_probe_loop(const NsxDev& d, NsxCtl& c, NsxLoopShared& L, NsxPivotScratch& pv,
                                               NsxPotScratch& ps, Sweep& sweep, int32_t count) {
    if (BLK && d.blk) nsx_blk_init(d, *d.blk);
    if (c.n_special < 0) nsx_count_special(d, c, pv);
    NSX_SYNC();
    nsx_recompute_all_potentials<BLK>(d, 1, ps);
    for (int32_t k = 0; k < count; ++k) {
        NSX_SYNC();
        if (threadIdx.x == 0) {
            NsxCmd& cmd = L.cmd;
            if (c.pricing == NSX_PRICING_DEVEX && !c.row_scan_first) {
                int64_t st = (k % ((d.m + c.bs - 1) / c.bs)) * c.bs;
                cmd.kind = NSX_CMD_DEVEX; cmd.lo = st; cmd.hi = st + c.bs < d.m ? st + c.bs : d.m;
            } else {
                cmd.kind = NSX_CMD_DANTZIG; cmd.lo = 0; cmd.hi = d.m;
                cmd.pad[1] = c.n_special == 0;
            }
            pv.pi_delta_n = -1;  // (a probe sweep copies all potentials, like the first sweep after a phase switch)
            cmd.phase = 1; cmd.excluded = -1; cmd.wepoch = 0; cmd.reverse = k & 1;
            c.arcs_priced += cmd.hi - cmd.lo;
            c.sweeps++;
        }
        NSX_SYNC();
        sweep.run(L.cmd, L.dz, L.dx, c, []() {});
    }
    NSX_SYNC();
    if (threadIdx.x == 0) { c.status = NSX_STATUS_OPTIMAL; c.total = 0; }
    sweep.finish();
}

// BLK = false: the tree fits the pivot CTA's shared memory (dense preorder array); true: it lives in HBM (blocked array).
template <bool BLK>
__device__ __forceinline__ void nsx_resident_body(const NsxKernelArgs& a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NsxCtaShared& sh = *reinterpret_cast<NsxCtaShared*>(smem_raw);
    unsigned char* dyn = smem_raw + nsx_align16(sizeof(NsxCtaShared));
    const NsxDev& d = a.d;
    nsx_init_barriers(sh);
    uint32_t stage_count = 0, q0 = 0;
    NSX_SYNC();

    if (blockIdx.x == 0) {
        unsigned long long t_begin = 0;
        if (threadIdx.x == 0) t_begin = nsx_globaltimer();
        nsx_copy_ctl(&sh.ctl, a.ctl);
        double* pis = nullptr;
        const NsxDev dl = nsx_make_resident(d, a.plan, dyn, &pis);
        NSX_SYNC();
        const bool resident = a.plan.mode != NSX_RES_NONE;
        NsxSweepCtx cx{&a.st, (resident || a.plan.stage_pi) ? pis : nullptr, !resident && a.plan.stage_pi != 0,
                       dyn + a.plan.ring_off, a.plan.stages};
        GridSweep sweep{d, a.grid, a.slots, a.topk, sh, cx, stage_count, q0, 0, 0ull, 0ull, 0ull, a.shard, 0ull, 0ull, a.spin_ns, a.star, a.timeline, 0ull};
        if (a.probe_sweeps > 0) nsx_probe_loop<BLK>(dl, sh.ctl, sh.L, sh.piv, sh.pot, sweep, a.probe_sweeps);
        else nsx_solve_loop<BLK>(dl, sh.ctl, sh.L, sh.piv, sh.pot, a.trace, sweep);
        NSX_SYNC();
        if (threadIdx.x == 0) {
            unsigned long long total = nsx_globaltimer() - t_begin;
            sh.ctl.clk_pricing = (int64_t)sweep.t_price;
            sh.ctl.clk_sync = (int64_t)sweep.t_sync;
            sh.ctl.clk_xchg = (int64_t)sweep.t_xchg;
            sh.ctl.clk_pivot = (int64_t)(total - sweep.t_price);
        }
        NSX_SYNC();
        nsx_copy_ctl(a.ctl, &sh.ctl);
        return;
    }
    // worker CTAs: wait for a command, price, deliver, repeat
    if (threadIdx.x == 0) sh.tl_grid = a.timeline ? a.grid : nullptr;
    double* pis = a.wplan.stage_pi ? reinterpret_cast<double*>(dyn) : nullptr;
    unsigned char* ring = dyn + a.wplan.ring_off;
    int32_t seen = 0;
    uint32_t bar_rounds = 0;
    bool pi_valid = false;  // the potentials in `pis` are those of the previous command (a star command overwrites them)
    for (;;) {
        if (threadIdx.x == 0) {
            int32_t s;
            uint32_t spins = 0;
            bool lost = false;
            unsigned long long t_wait = nsx_globaltimer(), beat = 0ull;
            // the deadline runs from the last sign of life of the pivot CTA (it may be waiting for a peer GPU - up to
            // spin_ns - or scanning arcs itself for a long time between two commands)
            while ((s = nsx_ld_acquire(&a.grid->seq)) == seen) {
                __nanosleep(20);
                if ((++spins & 4095u) == 0) {
                    const unsigned long long b = *reinterpret_cast<volatile unsigned long long*>(&a.grid->arrived);
                    if (b != beat) { beat = b; t_wait = nsx_globaltimer(); }
                    else if (nsx_globaltimer() - t_wait > 4ull * a.spin_ns) { lost = true; break; }
                }
            }
            if (lost) {
                atomicExch(&a.grid->abort, 4);
                sh.cmd.kind = NSX_CMD_EXIT;
            } else {
                seen = s;
                if (a.timeline) NSX_TL(a.grid, 0);
                union { NsxCmd c; int4 v[3]; } tmp;
                const int4* src = reinterpret_cast<const int4*>(&a.grid->cmd);
                tmp.v[0] = __ldcg(src); tmp.v[1] = __ldcg(src + 1); tmp.v[2] = __ldcg(src + 2);
                sh.cmd = tmp.c;
            }
        }
        NSX_SYNC();
        const NsxCmd cmd = sh.cmd;
        if (cmd.kind == NSX_CMD_EXIT) return;
        NsxCand dz; NsxDevexCand dx;
        const bool starcmd = cmd.kind == NSX_CMD_STAR || cmd.kind == NSX_CMD_STAR_BUILD;
        if (starcmd) {
            pi_valid = false;
            int64_t evaluated = 0; int32_t fault = 0;
            nsx_cta_star(d, a.star, cmd, (int)blockIdx.x - 1, (int)gridDim.x - 1, dyn, sh, bar_rounds, a.spin_ns, dz, dx, evaluated, fault);
            // arcs examined by this CTA (lane 0 of every warp counted its warp's): summed through shared memory
            if (threadIdx.x == 0) *reinterpret_cast<unsigned long long*>(&sh.x_recs[0][0]) = 0ull;
            NSX_SYNC();
            if (evaluated) atomicAdd(reinterpret_cast<unsigned long long*>(&sh.x_recs[0][0]), (unsigned long long)evaluated);
            NSX_SYNC();
            if (threadIdx.x == 0) {
                const unsigned long long ev = *reinterpret_cast<unsigned long long*>(&sh.x_recs[0][0]);
                NsxSlot* sl = a.slots + blockIdx.x;
                if (cmd.pad[0]) { union { NsxDevexCand c; int4 v[2]; } tmp; tmp.c = dx; sl->v[0] = tmp.v[0]; sl->v[1] = tmp.v[1]; }
                else { union { NsxCand c; int4 v; } tmp; tmp.c = dz; sl->v[0] = tmp.v; }
                sl->pad[0] = (int32_t)(uint32_t)ev; sl->pad[1] = (int32_t)(uint32_t)(ev >> 32); sl->pad[2] = fault;
                sl->pad[3] = cmd.kind == NSX_CMD_STAR ? sh.tk_cnt : 0;  // rows that were queued (the same count on every worker)
                nsx_st_release(&sl->seq, seen);
                if (a.timeline) NSX_TL(a.grid, 5);
            }
            continue;
        }
        nsx_cta_sweep(d, a.st, cmd, pis, pis != nullptr, pi_valid ? cmd.pad[2] : -1, stage_count, ring, a.wplan.stages, q0,
                      a.shard.rank * ((int)gridDim.x - 1) + (int)blockIdx.x - 1, a.shard.world * ((int)gridDim.x - 1), sh, dz, dx);
        pi_valid = pis != nullptr;
        if (cmd.kind == NSX_CMD_TOPK) {  // this CTA's sorted list -> HBM (the slot release below orders it)
            NsxTopkOut* out = a.topk + blockIdx.x;
            const int32_t cnt = sh.tk_cnt;
            for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
                out->key[k] = reinterpret_cast<unsigned long long*>(sh.piv.res)[k];
                out->idx[k] = sh.piv.arc2[k];
            }
            if (threadIdx.x == 0) out->count = cnt;
            __threadfence();
            NSX_SYNC();
        }
        if (threadIdx.x == 0) {
            if (a.timeline) NSX_TL(a.grid, 4);
            NsxSlot* sl = a.slots + blockIdx.x;
            if (cmd.kind == NSX_CMD_DANTZIG || cmd.kind == NSX_CMD_DANTZIG_ZERO) {
                nsx_ll_store(sl, dz.key, dz.arc2, dz.zero2, seen);  // self-validating words: no sequence word, no fence
                if (dz.arc2 >= 0) {
                    // If this candidate wins, the pivot CTA starts with dependent reads of its arc record (endpoints, then
                    // flow / capacity / cost): pull those lines into L2 now (random HBM access -> L2 hit, ~0.4 us each)
                    const int64_t arc = dz.arc2 >> 1;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(d.tail + arc));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(d.head + arc));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(d.flow + arc));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(d.upper + arc));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(d.pert + arc));
                }
            } else {
                if (cmd.kind == NSX_CMD_DEVEX || cmd.kind == NSX_CMD_DEVEX_ZERO) {
                    union { NsxDevexCand c; int4 v[2]; } tmp; tmp.c = dx;
                    sl->v[0] = tmp.v[0]; sl->v[1] = tmp.v[1];
                } else {
                    union { NsxCand c; int4 v; } tmp; tmp.c = dz;
                    sl->v[0] = tmp.v;
                }
                nsx_st_release(&sl->seq, seen);  // the payload above is ordered before the sequence number
            }
            if (a.timeline) NSX_TL(a.grid, 5);
        }
    }
}

#ifdef NSX_DEV_SWEEP_ONLY
// Development aid (never part of the library): the worker side of a sweep alone, so that `nvcc -cubin -DNSX_DEV_SWEEP_ONLY`
// plus `cuobjdump -sass` shows the pricing loop within seconds instead of a full build of the four resident kernels.
extern "C" __global__ void __launch_bounds__(NSX_THREADS, 1) nsx_dev_sweep_kernel(const NsxKernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NsxCtaShared& sh = *reinterpret_cast<NsxCtaShared*>(smem_raw);
    unsigned char* dyn = smem_raw + nsx_align16(sizeof(NsxCtaShared));
    nsx_init_barriers(sh);
    uint32_t stage_count = 0, q0 = 0;
    NSX_SYNC();
    double* pis = a.wplan.stage_pi ? reinterpret_cast<double*>(dyn) : nullptr;
    NsxCand dz; NsxDevexCand dx;
    const NsxCmd cmd = a.grid->cmd;
    nsx_cta_sweep(a.d, a.st, cmd, pis, pis != nullptr, -1, stage_count, dyn + a.wplan.ring_off, a.wplan.stages, q0, (int)blockIdx.x, (int)gridDim.x, sh, dz, dx);
    if (threadIdx.x == 0) { union { NsxCand c; int4 v; } tmp; tmp.c = dz; a.slots[blockIdx.x].v[0] = tmp.v; }
}
extern "C" __global__ void nsx_resident_kernel(const NsxKernelArgs a);      // (declared for the host code below, not built)
extern "C" __global__ void nsx_resident_kernel_hbm(const NsxKernelArgs a);
struct NsxBatchItem;
extern "C" __global__ void nsx_batch_kernel(const NsxBatchItem* items, int64_t count, unsigned long long* next, size_t limit_bytes, int want_mode, int want_stage);
extern "C" __global__ void nsx_batch_kernel_hbm(const NsxBatchItem* items, int64_t count, unsigned long long* next, size_t limit_bytes, int want_mode, int want_stage);
#else
extern "C" __global__ void __launch_bounds__(NSX_THREADS, 1) nsx_resident_kernel(const NsxKernelArgs a) { nsx_resident_body<false>(a); }
extern "C" __global__ void __launch_bounds__(NSX_THREADS, 1) nsx_resident_kernel_hbm(const NsxKernelArgs a) { nsx_resident_body<true>(a); }
#endif

// Initial state: real arcs, nodes + artificial arcs, artificial-flow count.
extern "C" __global__ void nsx_init_kernel(const NsxDev d, const double* supply, NsxCtl* ctl) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    int32_t special = 0;  // NsxCtl::n_special of the initial state (the host set it to 0: the pivot CTA need not count)
    for (int64_t i = g; i < d.m; i += T) { nsx_init_real_arc(d, i); special += nsx_special(d.state[i]); }
    if (special) atomicAdd(&ctl->n_special, special);
    unsigned long long art = 0;
    for (int64_t v = g; v < d.n; v += T) {
        nsx_init_node(d, (int32_t)v, supply[v]);
        if (v > 0 && d.flow[d.m + v - 1] > d.tol) art++;
    }
    if (art) atomicAdd((unsigned long long*)&ctl->art_with_flow, art);
}

// Warm start: arc state from the caller's tree flags; flows, node records, depth and the preorder array were copied in.
extern "C" __global__ void nsx_init_warm_kernel(const NsxDev d, const double* supply, const uint8_t* in_tree, NsxCtl* ctl) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    unsigned long long art = 0;
    int32_t special = 0;
    for (int64_t a = g; a < d.ma; a += T) {
        art += (unsigned long long)nsx_init_arc_warm(d, a, supply, in_tree);
        if (a < d.m) special += nsx_special(d.state[a]);
    }
    if (special) atomicAdd(&ctl->n_special, special);
    if (g == 0) d.pi[0] = 0.0;
    if (art) atomicAdd((unsigned long long*)&ctl->art_with_flow, art);
}

// Which compact cost encodings are exact for this instance: bit 0 set = some cost is not an
// int32-valued integer, bit 1 set = some cost is not an int16-valued integer.
// Bit 2 set = some arc endpoint lies outside 1 .. n-1 (real arcs never touch the root): the call is refused.
extern "C" __global__ void nsx_classify_costs_kernel(const double* pert, const int32_t* tail, const int32_t* head,
                                                     int32_t n, int64_t m, unsigned int* flags) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    unsigned int f = 0;
    for (int64_t i = g; i < m; i += T) {
        const double c = pert[i];
        if (!(c == rint(c)) || fabs(c) > 2147483000.0) f |= 3u;
        else if (fabs(c) > 32767.0) f |= 2u;
        const int32_t tl = tail[i], hd = head[i];
        if (tl < 1 || tl >= n || hd < 1 || hd >= n) f |= 4u;
        if (i + 1 < m && tail[i + 1] < tl) f |= 8u;  // not grouped by ascending tail: no rows, no star pricing
    }
    if (f) atomicOr(flags, f);
}

// ---- star pricing: CSR offsets of the rows and the CSC copy of the arcs (built once per solve) ----
extern "C" __global__ void nsx_star_count_kernel(const int32_t* tail, const int32_t* head, int64_t m, int32_t* row_cnt, int32_t* col_cnt) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, T = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = g; i < m; i += T) { atomicAdd(row_cnt + tail[i] + 1, 1); atomicAdd(col_cnt + head[i] + 1, 1); }
}
// in-place inclusive scan of two arrays of n + 1 counters (element 0 is 0): one CTA, chunk per thread
extern "C" __global__ void nsx_star_scan_kernel(int32_t* a, int32_t* b, int32_t* cursor, int32_t n1) {
    __shared__ int32_t part[2][1024];
    const int T = blockDim.x, t = threadIdx.x;
    const int32_t per = (n1 + T - 1) / T, lo = t * per < n1 ? t * per : n1, hi = lo + per < n1 ? lo + per : n1;
    int32_t sa = 0, sb = 0;
    for (int32_t i = lo; i < hi; ++i) { sa += a[i]; sb += b[i]; }
    part[0][t] = sa; part[1][t] = sb;
    __syncthreads();
    if (t == 0) { int32_t ra = 0, rb = 0; for (int k = 0; k < T; ++k) { int32_t x = part[0][k]; part[0][k] = ra; ra += x; x = part[1][k]; part[1][k] = rb; rb += x; } }
    __syncthreads();
    int32_t ra = part[0][t], rb = part[1][t];
    for (int32_t i = lo; i < hi; ++i) { ra += a[i]; a[i] = ra; rb += b[i]; b[i] = rb; }
    __syncthreads();
    for (int32_t i = lo; i < hi; ++i) if (i + 1 < n1) cursor[i] = b[i];  // fill cursor of column i = col_begin[i]
}
extern "C" __global__ void nsx_star_fill_kernel(const int32_t* tail, const int32_t* head, const double* pert, const uint8_t* state,
                                                int64_t m, int32_t* cursor, int32_t* csc_arc, int32_t* csc_tail, void* csc_cost,
                                                int32_t cost_i32, int32_t* csc_pos, uint8_t* csc_state, uint32_t* csc_wgt) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, T = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = g; i < m; i += T) {
        const int32_t e = atomicAdd(cursor + head[i], 1);  // (the order inside a column is free: proposals commute)
        csc_arc[e] = (int32_t)i; csc_tail[e] = tail[i]; csc_pos[i] = e; csc_state[e] = state[i];
        if (csc_wgt) csc_wgt[e] = 1u;  // epoch 0, weight 1 (nsx_init_real_arc)
        if (cost_i32) reinterpret_cast<int32_t*>(csc_cost)[e] = (int32_t)pert[i];
        else reinterpret_cast<double*>(csc_cost)[e] = pert[i];
    }
}

// Canonical arrays (int32 tail / head, float64 perturbed cost) -> tile-padded pricing store.
// Element i of tile i / NSX_TILE; any launch shape (g = first element, T = stride).
__device__ __forceinline__ void nsx_pack_range(const int32_t* tail, const int32_t* head, const double* pert,
                                               int64_t m, int64_t mpad, const NsxStore& st, int64_t g, int64_t T) {
    unsigned char* base = const_cast<unsigned char*>(st.base);
    for (int64_t i = g; i < mpad; i += T) {
        const bool in = i < m;
        const int32_t tl = in ? tail[i] : 1, hd = in ? head[i] : 1;
        const double c = in ? pert[i] : 0.0;
        const int64_t tile = i / NSX_TILE;
        const int32_t r = (int32_t)(i - tile * NSX_TILE);
        unsigned char* tb = base + (size_t)tile * st.tile_bytes;
        if (st.node_kind != NSX_NODE_I32) {
            const int sh8 = st.node_kind == NSX_NODE_U16X8 ? 3 : 0;
            reinterpret_cast<uint16_t*>(tb)[r] = (uint16_t)((tl - 1) << sh8);
            reinterpret_cast<uint16_t*>(tb + st.off_head)[r] = (uint16_t)((hd - 1) << sh8);
        } else {
            reinterpret_cast<int32_t*>(tb)[r] = tl;
            reinterpret_cast<int32_t*>(tb + st.off_head)[r] = hd;
        }
        if (st.cost_kind == NSX_COST_F64) reinterpret_cast<double*>(tb + st.off_cost)[r] = c;
        else if (st.cost_kind == NSX_COST_I32) reinterpret_cast<int32_t*>(tb + st.off_cost)[r] = (int32_t)c;
        else reinterpret_cast<int16_t*>(tb + st.off_cost)[r] = (int16_t)((int32_t)c - (st.cost_kind == NSX_COST_I16M1 ? 1 : 0));
    }
}
extern "C" __global__ void nsx_pack_kernel(const int32_t* tail, const int32_t* head, const double* pert,
                                           int64_t m, int64_t mpad, NsxStore st) {
    nsx_pack_range(tail, head, pert, m, mpad, st, (int64_t)blockIdx.x * blockDim.x + threadIdx.x,
                   (int64_t)gridDim.x * blockDim.x);
}

// ------------------------------------------------------------------------------------------
// Batched variant: one CTA solves one independent instance end to end (config 4).
// ------------------------------------------------------------------------------------------
struct NsxBatchItem {
    NsxDev d;
    NsxStore st;
    int64_t mpad;
    NsxCtl* ctl;
    int32_t* trace;
    const double* supply;
};

struct LocalSweep {
    const NsxDev& d;
    NsxCtaShared& sh;
    NsxSweepCtx cx;
    uint32_t& stage_count;
    uint32_t& q0;
    template <class Deferred>
    __device__ void run(const NsxCmd& cmd_in, NsxCand& out_dz, NsxDevexCand& out_dx, NsxCtl& c, Deferred deferred) {
        const NsxCmd cmd = cmd_in;
        NsxCand dz; NsxDevexCand dx;
        deferred();  // (one CTA does everything: nothing to overlap with)
        NSX_SYNC();
        if (threadIdx.x == 0) __threadfence();  // this CTA's own state / potential writes reach L2 before the bulk reads
        nsx_cta_sweep(d, *cx.st, cmd, cx.pis, cx.stage, -1, stage_count, cx.ring, cx.stages, q0, 0, 1, sh, dz, dx);  // (potentials staged in full, if at all)
        if (cmd.kind == NSX_CMD_TOPK) { nsx_tk_publish_list(sh, c); return; }
        if (threadIdx.x == 0) {
            if (cmd.kind == NSX_CMD_DEVEX || cmd.kind == NSX_CMD_DEVEX_ZERO) out_dx = dx; else out_dz = dz;
        }
        NSX_SYNC();
    }
    __device__ void alive() {}
    __device__ void finish() {}
};

// The shared-memory plan of an item is made in the kernel from its node count; `limit_bytes` =
// dynamic bytes available after the fixed part (sized by the host for the largest instance).
template <bool BLK>
__device__ __forceinline__ void nsx_batch_body(const NsxBatchItem* items, int64_t count, unsigned long long* next, size_t limit_bytes,
                 int want_mode, int want_stage) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NsxCtaShared& sh = *reinterpret_cast<NsxCtaShared*>(smem_raw);
    unsigned char* dyn = smem_raw + nsx_align16(sizeof(NsxCtaShared));
    __shared__ unsigned long long my_item;
    nsx_init_barriers(sh);
    uint32_t stage_count = 0, q0 = 0;
    for (;;) {
        NSX_SYNC();
        if (threadIdx.x == 0) my_item = atomicAdd(next, 1ull);
        NSX_SYNC();
        const unsigned long long it = my_item;
        if (it >= (unsigned long long)count) return;
        const NsxBatchItem& item = items[it];
        const NsxDev& d = item.d;
        nsx_copy_ctl(&sh.ctl, item.ctl);
        NSX_SYNC();
        for (int64_t i = threadIdx.x; i < d.m; i += blockDim.x) {
            nsx_init_real_arc(d, i);
            const int32_t tl = d.tail[i], hd = d.head[i];
            if (tl < 1 || tl >= d.n || hd < 1 || hd >= d.n) sh.ctl.fault = 5;  // (same value from every thread)
        }
        NSX_SYNC();
        if (sh.ctl.fault) {  // refused: the instance keeps status -1 and the call returns an error
            nsx_copy_ctl(item.ctl, &sh.ctl);
            continue;
        }
        for (int64_t v = threadIdx.x; v < d.n; v += blockDim.x) nsx_init_node(d, (int32_t)v, item.supply[v]);
        nsx_pack_range(d.tail, d.head, d.pert, d.m, item.mpad, item.st, threadIdx.x, blockDim.x);
        NSX_SYNC();
        if (threadIdx.x == 0) {
            int64_t art = 0;
            for (int32_t v = 1; v < d.n; ++v) art += d.flow[d.m + v - 1] > d.tol;
            sh.ctl.art_with_flow = art;
        }
        NSX_SYNC();
        const NsxSmemPlan plan = nsx_plan_pivot((size_t)d.n, limit_bytes, item.st.stage_bytes, true, want_mode, want_stage);
        double* pis = nullptr;
        const NsxDev dl = nsx_make_resident(d, plan, dyn, &pis);
        NSX_SYNC();
        const bool resident = plan.mode != NSX_RES_NONE;
        NsxSweepCtx cx{&item.st, (resident || plan.stage_pi) ? pis : nullptr, !resident && plan.stage_pi != 0,
                       dyn + plan.ring_off, plan.stages};
        LocalSweep sweep{d, sh, cx, stage_count, q0};
        nsx_solve_loop<BLK>(dl, sh.ctl, sh.L, sh.piv, sh.pot, item.trace, sweep);
        NSX_SYNC();
        nsx_copy_ctl(item.ctl, &sh.ctl);
    }
}
#ifndef NSX_DEV_SWEEP_ONLY
// nsx_batch_kernel: every instance of the batch fits its CTA's shared memory (what the host checked); _hbm: some do not
extern "C" __global__ void __launch_bounds__(NSX_THREADS, 1)
nsx_batch_kernel(const NsxBatchItem* items, int64_t count, unsigned long long* next, size_t limit_bytes, int want_mode, int want_stage) {
    nsx_batch_body<false>(items, count, next, limit_bytes, want_mode, want_stage);
}
extern "C" __global__ void __launch_bounds__(NSX_THREADS, 1)
nsx_batch_kernel_hbm(const NsxBatchItem* items, int64_t count, unsigned long long* next, size_t limit_bytes, int want_mode, int want_stage) {
    nsx_batch_body<true>(items, count, next, limit_bytes, want_mode, want_stage);
}
#endif

// ------------------------------------------------------------------------------------------
// Host side: C ABI
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int nsx_fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define NSX_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t err__ = (call);                                                             \
        if (err__ != cudaSuccess) {                                                             \
            arena.release(); inputs.release();                                                  \
            return nsx_fail(NSX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__)); \
        }                                                                                       \
    } while (0)

// One cudaMalloc per arena; sub-allocations are 256-byte aligned.
struct Arena {
    unsigned char* base = nullptr;
    size_t size = 0, used = 0;
    size_t plan(size_t bytes) { size_t off = size; size += (bytes + 255) & ~(size_t)255; return off; }
    cudaError_t commit() { return cudaMalloc((void**)&base, size ? size : 256); }
    template <class T> T* at(size_t off) { return reinterpret_cast<T*>(base + off); }
    void release() { if (base) { cudaFree(base); base = nullptr; } }
    ~Arena() { release(); }
};
// Stream + timing events of one call; destroyed on every exit path.
struct CallResources {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaError_t create() {
        cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
        for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ev[i]);
        return e;
    }
    ~CallResources() {
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

static int nsx_validate(const nsx_problem* p, const nsx_options* o, const nsx_result* r) {
    if (!p || !o || !r) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null argument");
    if (p->n_nodes < 1 || p->n_arcs < 0) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "bad sizes");
    if (p->n_arcs + p->n_nodes >= (1ll << 30)) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "arc count exceeds 2^30");
    if (p->n_arcs > 0 && (!p->tail || !p->head || !p->pert_cost || !p->upper))
        return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null arc array");
    if (!p->supply) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null supply");
    if (o->pricing != NSX_PRICING_DANTZIG && o->pricing != NSX_PRICING_DEVEX && o->pricing != NSX_PRICING_CANDIDATE_LIST &&
        o->pricing != NSX_PRICING_DEVEX_LOOP)
        return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "unknown pricing rule");
    if (o->max_iterations < 0 || !(o->tolerance > 0) || o->ft_update_limit <= 0 || o->spin_timeout_ms < 0)
        return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "bad option value");
    if (o->flags != 0) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "nsx_options.flags is reserved and must be 0");
    if (o->row_scan_first < NSX_SPECIAL_NONE || o->row_scan_first > NSX_SPECIAL_SHORTEST_PATH)
        return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "unknown structure-specific rule (nsx_options.row_scan_first)");
    if (o->row_scan_first == NSX_SPECIAL_SHORTEST_PATH && !o->node_mask)
        return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "the shortest-path rule needs nsx_options.node_mask");
    return 0;
}

static void nsx_fill_ctl(NsxCtl& c, const nsx_options* o, bool trace) {
    memset(&c, 0, sizeof c);
    c.phase = 1; c.status = -1; c.maxit = o->max_iterations;
    c.bs = o->block_size > 0 ? o->block_size : 1; c.pb = 0; c.last_deg = -1;
    c.ft_limit = o->ft_update_limit; c.auto_block = o->auto_block;
    c.pricing = o->pricing; c.row_scan_first = o->row_scan_first;
    c.trace_cap = trace ? o->trace_capacity : 0;
    c.unbounded_arc = -1;
    c.n_special = -1;  // "not counted yet": nsx_solve_loop counts it (batch kernel); nsx_solve_impl lets its init kernels count
}

static int nsx_env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}
static unsigned long long nsx_spin_ns(const nsx_options* o) {
    long long ms = o->spin_timeout_ms > 0 ? o->spin_timeout_ms : nsx_env_int("NSX_SPIN_TIMEOUT_MS", 30000);
    if (ms < 1) ms = 1;
    return (unsigned long long)ms * 1000000ull;
}
static const char* nsx_fault_text(int fault) {
    switch (fault) {
        case 1: return "a sweep worker CTA did not deliver its candidate before the deadline";
        case 2: return "a peer GPU did not deliver its candidate before the deadline";
        case 3: return "a peer GPU (or its host process) raised the abort word";
        case 4: return "a sweep worker CTA saw no command before the deadline";
        case 5: return "arc endpoint outside 1 .. n_nodes-1";
        default: return "resident kernel ended without a status";
    }
}

static void nsx_harvest(const NsxCtl& c, nsx_result* res) {
    if (getenv("NSX_DEBUG")) fprintf(stderr, "[nsx] n_special at exit: %d\n", c.n_special);
    res->fault = c.fault;
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
    res->sweeps = c.sweeps;
    res->unbounded_arc = c.unbounded_arc;
    res->unbounded_rc = c.unbounded_rc;
    res->sum_cycle_len = c.sum_cycle;
    res->sum_subtree = c.sum_subtree;
    res->max_subtree = c.max_subtree;
    res->sum_rounds = c.sum_rounds;
    res->sum_window = c.sum_window;
    res->pricing_ms = (double)c.clk_pricing * 1e-6;
    res->pivot_ms = (double)c.clk_pivot * 1e-6;
    res->sync_ms = (double)c.clk_sync * 1e-6;
    res->exchange_ms = (double)c.clk_xchg * 1e-6;
    for (int i = 0; i < 12; ++i) res->phase_cycles[i] = c.ph[i];
    res->star_pricing = c.star_on;
    res->star_updates = c.star_updates; res->star_builds = c.star_builds; res->star_rescans = c.star_rescans;
    res->blk_rebuilds = c.blk_rebuilds;
}

struct DeviceInfo { int sms = 0; int coop = 0; size_t smem_optin = 0; bool ok = false; };
// cudaGetDeviceProperties costs milliseconds: query each device once per process
static int nsx_device_info(int dev, DeviceInfo& info) {
    static std::mutex mu;
    static DeviceInfo cache[64];
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && cache[dev].ok) { info = cache[dev]; return 0; }
    int sms = 0, coop = 0, smem = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return nsx_fail(NSX_ERR_NO_DEVICE, std::string("cudaDeviceGetAttribute: ") + cudaGetErrorString(e));
    info.sms = sms; info.coop = coop; info.smem_optin = (size_t)smem; info.ok = true;
    if (dev >= 0 && dev < 64) cache[dev] = info;
    return 0;
}

static size_t nsx_smem_fixed() { return (sizeof(NsxCtaShared) + 15) & ~(size_t)15; }
// entries of the preorder array: the block arena of the blocked layout (which also covers the dense array of n entries)
static size_t nsx_order_len(size_t n) {
    const size_t arena = (size_t)NSX_BLK_MAX << nsx_blk_lg((int64_t)n);
    return arena > n ? arena : n;
}
static int64_t nsx_pad_tiles(int64_t m) {
    int64_t t = (m + NSX_TILE - 1) / NSX_TILE;
    return (t < 1 ? 1 : t) * (int64_t)NSX_TILE;
}

// Narrowest exact encoding of the pricing store (see the NsxStore comment).  `cost_flags` is the
// result of nsx_classify_costs_kernel.  NSX_LAYOUT=wide forces int32 ids + float64 costs.
static void nsx_choose_layout(int32_t n, unsigned int cost_flags, bool devex, NsxStore& st) {
    int node_kind = (n - 1 <= 65536) ? NSX_NODE_U16 : NSX_NODE_I32;
    int cost_kind = !(cost_flags & 2u) ? NSX_COST_I16 : !(cost_flags & 1u) ? NSX_COST_I32 : NSX_COST_F64;
    const char* force = getenv("NSX_LAYOUT");
    if (force && !strcmp(force, "wide")) { node_kind = NSX_NODE_I32; cost_kind = NSX_COST_F64; }
    if (force && !strcmp(force, "i32")) { node_kind = NSX_NODE_I32; if (cost_kind == NSX_COST_I16) cost_kind = NSX_COST_I32; }
    // pre-scaled ids + cost - 1 (see NSX_NODE_U16X8): (id - 1) * 8 fits a uint16 up to n - 1 = 8192; NSX_LAYOUT=plain16 keeps the plain columns
    if (node_kind == NSX_NODE_U16 && cost_kind == NSX_COST_I16 && n - 1 <= 8192 && !(force && !strcmp(force, "plain16"))) {
        node_kind = NSX_NODE_U16X8; cost_kind = NSX_COST_I16M1;
    }
    nsx_store_layout(st, node_kind, cost_kind, devex ? 1 : 0);
}

// Common implementation; `resident` = arc arrays are device pointers; probe_sweeps > 0 = sweep probe.
static int nsx_solve_impl(const nsx_problem* pb, const nsx_options* opt, nsx_result* res, bool resident,
                          int32_t probe_sweeps, const nsx_shard* shard = nullptr, const nsx_warm_start* warm = nullptr) {
    Arena arena, inputs;
    int rc = nsx_validate(pb, opt, res);
    if (rc) return rc;
    // warm start: lay the caller's tree out as preorder array + node records on the host (csrc/nsx_warm.h)
    std::vector<NsxNode> w_node;
    std::vector<int32_t> w_depth, w_order;
    if (warm) {
        if (resident || shard || probe_sweeps) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "warm start is a host-buffer, single-GPU entry point");
        if (!warm->in_tree || !warm->flow || (warm->start_phase != 1 && warm->start_phase != 2))
            return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "bad warm start description");
        for (int64_t a = 0; a < pb->n_arcs; ++a)  // nsx_warm_layout indexes host arrays with these
            if (pb->tail[a] < 1 || pb->tail[a] >= pb->n_nodes || pb->head[a] < 1 || pb->head[a] >= pb->n_nodes)
                return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "arc endpoint outside 1 .. n_nodes-1 (node 0 is the artificial root)");
        int bad = nsx_warm_layout(pb->n_nodes, pb->n_arcs, pb->tail, pb->head, pb->supply, opt->tolerance, warm->in_tree,
                                  w_node, w_depth, w_order);
        if (bad) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, bad == -1 ? "warm start: in_tree must mark exactly n_nodes - 1 arcs"
                                                                     : "warm start: the marked arcs do not span all nodes");
        if (warm->start_phase == 2)
            for (int64_t a = pb->n_arcs; a < pb->n_arcs + pb->n_nodes - 1; ++a)
                if (warm->in_tree[a]) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "warm start: Phase 1 can only be skipped when no artificial arc is in the tree");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return nsx_fail(NSX_ERR_NO_DEVICE, "no CUDA device visible (the engine has no CPU fallback)");
    if (opt->device < 0 || opt->device >= ndev) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    NSX_CUDA(cudaSetDevice(opt->device));
    DeviceInfo info;
    if ((rc = nsx_device_info(opt->device, info))) return rc;
    if (!info.coop) return nsx_fail(NSX_ERR_NO_DEVICE, "device lacks cooperative launch");

    const int32_t n = pb->n_nodes;
    const int64_t m = pb->n_arcs, ma = m + n - 1, mpad = nsx_pad_tiles(m);
    const bool devex = opt->pricing == NSX_PRICING_DEVEX || opt->pricing == NSX_PRICING_DEVEX_LOOP;
    const bool want_trace = res->entering_trace && opt->trace_capacity > 0;

    CallResources rs;
    NSX_CUDA(rs.create());
    cudaStream_t stream = rs.stream;
    cudaEvent_t* ev = rs.ev;

    // ---- inputs: canonical arrays in HBM (uploaded, or the caller's resident copies) ----
    size_t i_tail = 0, i_head = 0, i_pert = 0, i_upper = 0;
    if (!resident) {
        i_tail = inputs.plan((size_t)(m + 4) * 4); i_head = inputs.plan((size_t)(m + 4) * 4);
        i_pert = inputs.plan((size_t)(m + 4) * 8); i_upper = inputs.plan((size_t)(m + 4) * 8);
    }
    size_t i_supply = inputs.plan((size_t)n * 8), i_flags = inputs.plan(16);
    size_t i_wtree = warm ? inputs.plan((size_t)ma + 16) : 0;
    const bool want_mask = opt->row_scan_first == NSX_SPECIAL_SHORTEST_PATH;
    size_t i_mask = want_mask ? inputs.plan((size_t)n + 16) : 0;
    NSX_CUDA(inputs.commit());
    NsxKernelArgs ka;
    NsxDev& d = ka.d;
    d.n = n; d.m = m; d.ma = ma;
    NSX_CUDA(cudaEventRecord(ev[0], stream));
    if (resident) {
        d.tail = pb->tail; d.head = pb->head; d.pert = pb->pert_cost; d.upper = pb->upper;
    } else {
        d.tail = inputs.at<int32_t>(i_tail); d.head = inputs.at<int32_t>(i_head);
        d.pert = inputs.at<double>(i_pert); d.upper = inputs.at<double>(i_upper);
        if (m > 0) {
            NSX_CUDA(cudaMemcpyAsync((void*)d.tail, pb->tail, (size_t)m * 4, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync((void*)d.head, pb->head, (size_t)m * 4, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync((void*)d.pert, pb->pert_cost, (size_t)m * 8, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync((void*)d.upper, pb->upper, (size_t)m * 8, cudaMemcpyHostToDevice, stream));
        }
    }
    double* d_supply = inputs.at<double>(i_supply);
    unsigned int* d_flags = inputs.at<unsigned int>(i_flags);
    NSX_CUDA(cudaMemcpyAsync(d_supply, pb->supply, (size_t)n * 8, cudaMemcpyHostToDevice, stream));
    NSX_CUDA(cudaMemsetAsync(d_flags, 0, 16, stream));
    d.node_mask = nullptr;
    if (want_mask) {
        uint8_t* d_mask = inputs.at<uint8_t>(i_mask);
        NSX_CUDA(cudaMemcpyAsync(d_mask, opt->node_mask, (size_t)n, cudaMemcpyHostToDevice, stream));
        d.node_mask = d_mask;
    }
    const int util_blocks = info.sms * 8;
    if (m > 0) {
        nsx_classify_costs_kernel<<<util_blocks, 256, 0, stream>>>(d.pert, d.tail, d.head, n, m, d_flags);
        NSX_CUDA(cudaGetLastError());
    }
    unsigned int cost_flags = 3u;
    NSX_CUDA(cudaMemcpyAsync(&cost_flags, d_flags, 4, cudaMemcpyDeviceToHost, stream));
    NSX_CUDA(cudaStreamSynchronize(stream));
    if (m > 0 && (cost_flags & 4u)) {
        arena.release(); inputs.release();
        return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "arc endpoint outside 1 .. n_nodes-1 (node 0 is the artificial root)");
    }
    NsxStore& st = ka.st;
    nsx_choose_layout(n, cost_flags & 3u, devex, st);

    // ---- launch shape: one pivot CTA + sweep workers; small instances are priced by the pivot CTA alone ----
    int64_t arcs_per_cta = nsx_env_int("NSX_ARCS_PER_CTA", 8192);
    int64_t workers = (m + arcs_per_cta - 1) / arcs_per_cta;
    if (workers > info.sms - 1) workers = info.sms - 1;
    int grid = m < nsx_env_int("NSX_SINGLE_CTA_ARCS", 65536) || workers < 2 ? 1 : (int)workers + 1;
    int forced = nsx_env_int("NSX_GRID", 0);
    if (forced > 0) grid = forced < info.sms ? forced : info.sms;
    // Star pricing (row cache instead of full Dantzig sweeps): whenever the driver prices with the Dantzig rule over the
    // whole arc range - Dantzig pricing, or the transportation row scan in front of any rule - on a multi-CTA grid of one
    // GPU, cold start, arcs grouped by tail.  NSX_STAR=0 keeps the full sweeps (what the sweep roofline is quoted on).
    // Default: on when the average degree is moderate (m / n <= 512) - on dense instances (config 3: 2048 arcs per node, ~470
    // re-hung nodes per pivot) an update touches a tenth of the arcs at several times the bytes per arc and the TMA sweep
    // of the packed store is as fast; NSX_STAR=1 / 0 forces it on / off.
    const int star_env = nsx_env_int("NSX_STAR", -1);
    // ... and only where a full sweep is expensive (m >= 4 M arcs): on the 1 M-arc config 2 a sweep costs ~20 us, the row
    // cache ~22 us per step plus ~4 us of bookkeeping in the pivot (18.6 K against 19.4 K pivots/s over the whole solve)
    const bool star_wanted = star_env >= 0 ? star_env != 0 : (m / (n > 0 ? n : 1) <= 512 && m >= (4ll << 20));
    const bool star = star_wanted && grid > 1 && !shard && !warm && probe_sweeps == 0 && m > 0 && m < (1ll << 30) &&
                      !(cost_flags & 8u) && (opt->pricing == NSX_PRICING_DANTZIG || opt->row_scan_first == NSX_SPECIAL_ROW_SCAN ||
                                             (opt->pricing == NSX_PRICING_DEVEX && opt->row_scan_first == 0));
    // Devex (vectorised, no structure rule in front): the row cache holds arg-max candidates whenever one block covers all arcs
    const bool star_devex = star && opt->pricing == NSX_PRICING_DEVEX && opt->row_scan_first == 0;
    const bool star_i32 = !(cost_flags & 1u);  // every cost an exact int32

    // ---- engine-owned device memory ----
    const size_t state_len = (size_t)(ma > mpad ? ma : mpad) + 16;
    size_t o_store = arena.plan((size_t)(mpad / NSX_TILE) * st.tile_bytes);
    size_t o_atail = arena.plan((size_t)n * 4), o_ahead = arena.plan((size_t)n * 4), o_aupper = arena.plan((size_t)n * 8);
    size_t o_flow = arena.plan((size_t)(ma + 4) * 8), o_state = arena.plan(state_len);
    size_t o_wgt = devex ? arena.plan((size_t)mpad * 4) : 0;
    size_t o_node = arena.plan((size_t)n * sizeof(NsxNode)), o_depth = arena.plan((size_t)n * 4);
    const size_t order_len = nsx_order_len((size_t)n);
    size_t o_pi = arena.plan((size_t)n * 8 + 16), o_order = arena.plan(order_len * 4), o_tmp = arena.plan((size_t)n * 4);
    size_t o_sidx = arena.plan((size_t)n * 4);
    size_t o_gph = arena.plan((size_t)n * 4), o_gpt = arena.plan((size_t)n * 4);
    size_t o_garc2 = arena.plan(((size_t)2 * n + 1) * 4), o_gres = arena.plan(((size_t)2 * n + 1) * 8);
    size_t o_ctl = arena.plan(sizeof(NsxCtl)), o_grid = arena.plan(sizeof(NsxGridCtl));
    size_t o_slots = arena.plan(sizeof(NsxSlot) * 1024);
    size_t o_topk = arena.plan(sizeof(NsxTopkOut) * 160);
    size_t o_trace = want_trace ? arena.plan((size_t)opt->trace_capacity * 4) : 0;
    size_t o_imb = warm ? arena.plan((size_t)n * 8) : 0;
    size_t o_pidelta = arena.plan((size_t)NSX_PI_DELTA_CAP * sizeof(NsxPiDelta));
    size_t o_rc = 0, o_dlist = 0, o_dstamp = 0, o_rowb = 0, o_colb = 0, o_cursor = 0, o_cpos = 0, o_cstate = 0, o_carc = 0, o_ctail = 0,
           o_ccost = 0, o_rq = 0, o_rqn = 0, o_dinfo = 0, o_cwgt = 0;
    if (star) {
        o_rc = arena.plan((size_t)n * sizeof(NsxRC) * 2); o_cwgt = star_devex ? arena.plan((size_t)m * 4) : 0; o_dlist = arena.plan((size_t)n * 4); o_dstamp = arena.plan((size_t)n * 4);
        o_rowb = arena.plan(((size_t)n + 2) * 4); o_colb = arena.plan(((size_t)n + 2) * 4); o_cursor = arena.plan(((size_t)n + 2) * 4);
        o_cpos = arena.plan((size_t)m * 4); o_cstate = arena.plan((size_t)m + 16); o_carc = arena.plan((size_t)m * 4);
        o_ctail = arena.plan((size_t)m * 4); o_ccost = arena.plan((size_t)m * (star_i32 ? 4 : 8));
        o_rq = arena.plan((size_t)n * 4); o_rqn = arena.plan(256); o_dinfo = arena.plan(((size_t)n + 1) * 32);
    }
    NSX_CUDA(arena.commit());

    st.base = arena.at<unsigned char>(o_store);
    d.atail = arena.at<int32_t>(o_atail); d.ahead = arena.at<int32_t>(o_ahead); d.aupper = arena.at<double>(o_aupper);
    d.flow = arena.at<double>(o_flow); d.state = arena.at<uint8_t>(o_state);
    d.wgt = devex ? arena.at<uint32_t>(o_wgt) : nullptr;
    d.node = arena.at<NsxNode>(o_node); d.depth = arena.at<int32_t>(o_depth); d.pi = arena.at<double>(o_pi); d.pi_mirror = nullptr;
    d.order = arena.at<int32_t>(o_order); d.tmp = arena.at<int32_t>(o_tmp);
    d.sidx = arena.at<int32_t>(o_sidx); d.blk = nullptr;  // (the pivot CTA points blk at its shared memory)
    d.csc_wgt = nullptr; d.dinfo = nullptr; d.rc = nullptr; d.dlist = nullptr; d.dstamp = nullptr; d.row_begin = nullptr; d.col_begin = nullptr; d.csc_pos = nullptr; d.csc_state = nullptr;
    memset(&ka.star, 0, sizeof ka.star);
    if (star) {
        d.dinfo = arena.at<int32_t>(o_dinfo);
        d.csc_wgt = star_devex ? arena.at<uint32_t>(o_cwgt) : nullptr;
        d.rc = arena.at<NsxRC>(o_rc); d.dlist = arena.at<int32_t>(o_dlist); d.dstamp = arena.at<int32_t>(o_dstamp);
        d.row_begin = arena.at<int32_t>(o_rowb); d.col_begin = arena.at<int32_t>(o_colb);
        d.csc_pos = arena.at<int32_t>(o_cpos); d.csc_state = arena.at<uint8_t>(o_cstate);
        ka.star.on = 1; ka.star.cost_i32 = star_i32 ? 1 : 0;
        ka.star.csc_arc = arena.at<int32_t>(o_carc); ka.star.csc_tail = arena.at<int32_t>(o_ctail);
        ka.star.csc_cost = arena.at<unsigned char>(o_ccost);
        ka.star.rq = arena.at<int32_t>(o_rq); ka.star.rq_n = arena.at<int32_t>(o_rqn);
        ka.star.bar = reinterpret_cast<unsigned int*>(arena.at<int32_t>(o_rqn) + 16);
    }
    d.gpath_h = arena.at<int32_t>(o_gph); d.gpath_t = arena.at<int32_t>(o_gpt);
    d.garc2 = arena.at<int32_t>(o_garc2); d.gres = arena.at<double>(o_gres);
    d.penalty = pb->penalty; d.tol = opt->tolerance; d.scan_walk = 0; d.par16 = nullptr; d.root_bits = nullptr;
    d.imbalance = warm ? arena.at<double>(o_imb) : nullptr;
    d.pi_delta = arena.at<NsxPiDelta>(o_pidelta);  // (used by grids whose workers stage the potentials; set to null below otherwise)
    ka.ctl = arena.at<NsxCtl>(o_ctl); ka.grid = arena.at<NsxGridCtl>(o_grid);
    ka.slots = arena.at<NsxSlot>(o_slots);
    ka.topk = arena.at<NsxTopkOut>(o_topk);
    ka.trace = want_trace ? arena.at<int32_t>(o_trace) : nullptr;
    ka.probe_sweeps = probe_sweeps;
    ka.timeline = (probe_sweeps > 0 || nsx_env_int("NSX_TIMELINE", 0)) ? 1 : 0;
    ka.spin_ns = nsx_spin_ns(opt);
    memset(&ka.shard, 0, sizeof ka.shard);
    ka.shard.rank = 0; ka.shard.world = 1;
    if (shard) {
        if (shard->world < 1 || shard->world > NSX_MAX_WORLD || shard->rank < 0 || shard->rank >= shard->world || !shard->mailboxes) {
            arena.release(); inputs.release();
            return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "bad shard description");
        }
        ka.shard.rank = shard->rank; ka.shard.world = shard->world;
        for (int r = 0; r < shard->world; ++r) ka.shard.box[r] = reinterpret_cast<NsxMailbox*>(shard->mailboxes[r]);
    }

    NsxCtl hctl;
    nsx_fill_ctl(hctl, opt, want_trace);
    if (warm) { hctl.warm = 1; hctl.phase = warm->start_phase; }
    hctl.star_on = star ? (star_devex ? 2 : 1) : 0;
    hctl.star_excl_prev = -1;
    hctl.n_special = 0;  // accumulated by nsx_init_kernel / nsx_init_warm_kernel
    NSX_CUDA(cudaMemcpyAsync(ka.ctl, &hctl, sizeof hctl, cudaMemcpyHostToDevice, stream));
    NSX_CUDA(cudaMemsetAsync(ka.grid, 0, sizeof(NsxGridCtl), stream));
    NSX_CUDA(cudaMemsetAsync(ka.slots, 0, sizeof(NsxSlot) * 1024, stream));
    NSX_CUDA(cudaMemsetAsync(d.state, 0, state_len, stream));
    if (devex) NSX_CUDA(cudaMemsetAsync(d.wgt, 0, (size_t)mpad * 4, stream));
    nsx_pack_kernel<<<util_blocks, 256, 0, stream>>>(d.tail, d.head, d.pert, m, mpad, st);
    NSX_CUDA(cudaGetLastError());
    NSX_CUDA(cudaEventRecord(ev[1], stream));

    // ---- launch shape (grid: chosen above) ----
    const size_t fixed = nsx_smem_fixed();
    if (fixed + 2 * (size_t)st.stage_bytes > info.smem_optin) { arena.release(); inputs.release(); return nsx_fail(NSX_ERR_INTERNAL, "tile ring does not fit in shared memory"); }
    const size_t limit = info.smem_optin - fixed;
    const int want_mode = nsx_env_int("NSX_RESIDENT", 2), want_stage = nsx_env_int("NSX_STAGE_PI", 1);
    ka.plan = nsx_plan_pivot((size_t)n, limit, st.stage_bytes, grid == 1, want_mode, want_stage);
    if (!nsx_env_int("NSX_PAR16", 1) && ka.plan.par16) {  // (test knob: cycle walk on the node records in L2)
        ka.plan.par16 = 0; ka.plan.blk_off = 0; ka.plan.ring_off = (uint32_t)nsx_align16(sizeof(NsxBlk));
    }
    ka.wplan = nsx_plan_worker((size_t)n, limit, st.stage_bytes, want_stage);
    int max_stages = nsx_env_int("NSX_STAGES", NSX_MAX_STAGES);
    if (max_stages < 2) max_stages = 2;
    if (ka.plan.stages > max_stages) ka.plan.stages = max_stages;
    if (ka.wplan.stages > max_stages) ka.wplan.stages = max_stages;
    size_t dyn = nsx_plan_bytes(ka.plan, st.stage_bytes);
    if (star && dyn < (size_t)(4 * NSX_STAR_ENT + 1) * 4 + 64) dyn = (size_t)(4 * NSX_STAR_ENT + 1) * 4 + 64;  // star work tables of a worker
    if (grid > 1 && nsx_plan_bytes(ka.wplan, st.stage_bytes) > dyn) dyn = nsx_plan_bytes(ka.wplan, st.stage_bytes);
    if (grid == 1 || !ka.wplan.stage_pi || nsx_env_int("NSX_PI_DELTA", 1) == 0) d.pi_delta = nullptr;
    const size_t smem = fixed + dyn;
    if (smem > info.smem_optin || (grid == 1 ? ka.plan.stages : ka.wplan.stages) < 2) {
        arena.release(); inputs.release();
        return nsx_fail(NSX_ERR_INTERNAL, "shared memory plan exceeds the device limit");
    }
    auto* resident_kernel = ka.plan.mode == NSX_RES_NONE ? nsx_resident_kernel_hbm : nsx_resident_kernel;
    NSX_CUDA(cudaFuncSetAttribute(resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    NSX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, resident_kernel, NSX_THREADS, smem));
    if (per_sm < 1) { arena.release(); inputs.release(); return nsx_fail(NSX_ERR_INTERNAL, "resident kernel does not fit on an SM"); }

    {
        int ib = (int)((m + n + 1023) / 1024);
        if (ib < 1) ib = 1;
        if (ib > info.sms * 8) ib = info.sms * 8;
        if (warm) {
            NSX_CUDA(cudaMemsetAsync(d.imbalance, 0, (size_t)n * 8, stream));
            uint8_t* d_wtree = inputs.at<uint8_t>(i_wtree);
            NSX_CUDA(cudaMemcpyAsync(d_wtree, warm->in_tree, (size_t)ma, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync(d.flow, warm->flow, (size_t)ma * 8, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync(d.node, w_node.data(), (size_t)n * sizeof(NsxNode), cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync(d.depth, w_depth.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync(d.order, w_order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream));
            nsx_init_warm_kernel<<<ib, 1024, 0, stream>>>(d, d_supply, d_wtree, ka.ctl);
        } else {
            nsx_init_kernel<<<ib, 1024, 0, stream>>>(d, d_supply, ka.ctl);
        }
        NSX_CUDA(cudaGetLastError());
        if (star) {  // rows / CSC copy / empty stamps (after the init kernel: the CSC copy takes the initial state bytes)
            NSX_CUDA(cudaMemsetAsync((void*)d.row_begin, 0, ((size_t)n + 2) * 4, stream));
            NSX_CUDA(cudaMemsetAsync((void*)d.col_begin, 0, ((size_t)n + 2) * 4, stream));
            NSX_CUDA(cudaMemsetAsync(d.dstamp, 0, (size_t)n * 4, stream));
            NSX_CUDA(cudaMemsetAsync(ka.star.rq_n, 0, 256, stream));
            nsx_star_count_kernel<<<util_blocks, 256, 0, stream>>>(d.tail, d.head, m, (int32_t*)d.row_begin, (int32_t*)d.col_begin);
            nsx_star_scan_kernel<<<1, 1024, 0, stream>>>((int32_t*)d.row_begin, (int32_t*)d.col_begin, arena.at<int32_t>(o_cursor), n + 1);
            nsx_star_fill_kernel<<<util_blocks, 256, 0, stream>>>(d.tail, d.head, d.pert, d.state, m, arena.at<int32_t>(o_cursor),
                                                                 (int32_t*)ka.star.csc_arc, (int32_t*)ka.star.csc_tail, (void*)ka.star.csc_cost,
                                                                 ka.star.cost_i32, d.csc_pos, d.csc_state, d.csc_wgt);
            NSX_CUDA(cudaGetLastError());
        }
    }
    void* kargs[] = {(void*)&ka};
    // The kernel synchronises its CTAs through its own flags; the cooperative launch is only there for the guarantee
    // that all of them are resident at once.  NSX_LAUNCH_PLAIN=1 (tests that run several ranks of a sharded solve as
    // host threads on ONE device) uses an ordinary launch instead: the driver does not run two cooperative grids side by
    // side, and grid <= SM count with one CTA per SM is co-resident on an otherwise idle device anyway.
    if (nsx_env_int("NSX_LAUNCH_PLAIN", 0)) {
        // Ranks as host threads on one device: a device allocation or memset issued by one thread while another thread's
        // resident kernel is already spinning would wait for that kernel (implicit synchronisation) - which waits for this
        // rank.  So every rank finishes its set-up, then all meet here (NSX_LAUNCH_PLAIN = number of ranks) and launch.
        NSX_CUDA(cudaStreamSynchronize(stream));
        if (shard) {
            static std::mutex mu;
            static std::condition_variable cv;
            static int waiting = 0;
            static unsigned generation = 0;
            const int ranks = nsx_env_int("NSX_LAUNCH_PLAIN", 1);
            std::unique_lock<std::mutex> lock(mu);
            const unsigned mine = generation;
            if (++waiting >= ranks) { waiting = 0; ++generation; cv.notify_all(); }
            else cv.wait_for(lock, std::chrono::seconds(120), [&] { return generation != mine; });
        }
        resident_kernel<<<dim3(grid), dim3(NSX_THREADS), smem, stream>>>(ka);
        NSX_CUDA(cudaGetLastError());
    } else {
        NSX_CUDA(cudaLaunchCooperativeKernel((void*)resident_kernel, dim3(grid), dim3(NSX_THREADS), kargs, smem, stream));
    }
    NSX_CUDA(cudaEventRecord(ev[2], stream));

    // ---- results ----
    if (res->flow) NSX_CUDA(cudaMemcpyAsync(res->flow, d.flow, (size_t)ma * 8, cudaMemcpyDeviceToHost, stream));
    if (res->potential) NSX_CUDA(cudaMemcpyAsync(res->potential, d.pi, (size_t)n * 8, cudaMemcpyDeviceToHost, stream));
    if (res->state) NSX_CUDA(cudaMemcpyAsync(res->state, d.state, (size_t)ma, cudaMemcpyDeviceToHost, stream));
    NSX_CUDA(cudaMemcpyAsync(&hctl, ka.ctl, sizeof hctl, cudaMemcpyDeviceToHost, stream));
    NsxGridCtl hgrid;
    NSX_CUDA(cudaMemcpyAsync(&hgrid, ka.grid, sizeof hgrid, cudaMemcpyDeviceToHost, stream));
    NSX_CUDA(cudaEventRecord(ev[3], stream));
    NSX_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < 8; ++i) res->handshake_ns[i] = (int64_t)hgrid.tl[i];
    if (want_trace) {
        int64_t cnt = hctl.trace_len < opt->trace_capacity ? hctl.trace_len : opt->trace_capacity;
        if (cnt > 0) NSX_CUDA(cudaMemcpy(res->entering_trace, ka.trace, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
    }
    float ms = 0;
    nsx_harvest(hctl, res);
    cudaEventElapsedTime(&ms, ev[0], ev[1]); res->h2d_ms = ms;
    cudaEventElapsedTime(&ms, ev[1], ev[2]); res->solve_ms = ms;
    cudaEventElapsedTime(&ms, ev[2], ev[3]); res->d2h_ms = ms;
    res->grid_ctas = grid;
    res->bytes_per_arc = (int32_t)(2 * nsx_node_bytes(st.node_kind) + nsx_cost_bytes(st.cost_kind) + 1);
    res->ring_stages = grid == 1 ? ka.plan.stages : ka.wplan.stages;
    res->store_layout = st.node_kind | (st.cost_kind << 8);
    res->resident_mode = ka.plan.mode;
    arena.release();
    inputs.release();
    if (hgrid.abort && !res->fault) res->fault = hgrid.abort;
    if (res->fault || hctl.status < 0) return nsx_fail(NSX_ERR_INTERNAL, std::string("resident kernel gave up: ") + nsx_fault_text(res->fault));
    return 0;
}

extern "C" int nsx_solve(const nsx_problem* problem, const nsx_options* options, nsx_result* result) {
    return nsx_solve_impl(problem, options, result, false, 0);
}
extern "C" int nsx_solve_warm(const nsx_problem* problem, const nsx_options* options, const nsx_warm_start* warm,
                              nsx_result* result) {
    if (!warm) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null warm start");
    return nsx_solve_impl(problem, options, result, false, 0, nullptr, warm);
}
extern "C" int nsx_solve_resident(const nsx_problem* problem_dev, const nsx_options* options, nsx_result* result) {
    return nsx_solve_impl(problem_dev, options, result, true, 0);
}
extern "C" int nsx_sweep_probe(const nsx_problem* problem_dev, const nsx_options* options, int32_t sweeps,
                               nsx_result* result) {
    if (sweeps < 1) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "sweeps must be positive");
    return nsx_solve_impl(problem_dev, options, result, true, sweeps);
}

extern "C" int nsx_solve_sharded(const nsx_problem* problem, const nsx_options* options, nsx_result* result,
                                 const nsx_shard* shard) {
    if (!shard) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null shard");
    return nsx_solve_impl(problem, options, result, false, 0, shard);
}
extern "C" int nsx_solve_sharded_resident(const nsx_problem* problem_dev, const nsx_options* options, nsx_result* result,
                                          const nsx_shard* shard) {
    if (!shard) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null shard");
    return nsx_solve_impl(problem_dev, options, result, true, 0, shard);
}
extern "C" int nsx_sweep_probe_sharded(const nsx_problem* problem_dev, const nsx_options* options, int32_t sweeps,
                                       nsx_result* result, const nsx_shard* shard) {
    if (!shard || sweeps < 1) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "bad argument");
    return nsx_solve_impl(problem_dev, options, result, true, sweeps, shard);
}
extern "C" int64_t nsx_mailbox_bytes(void) { return (int64_t)sizeof(NsxMailbox); }
// the pinned word and the stream nsx_mailbox_abort copies from / on: made when a mailbox is created or opened, because a
// page-locked allocation issued while a resident kernel spins would wait for that kernel (implicit synchronisation)
static std::mutex g_abort_mu;
static unsigned long long* g_abort_word = nullptr;
static cudaStream_t g_abort_stream = nullptr;  // (of the device current at the first call: one device per process, one process per GPU)
static cudaError_t nsx_abort_prepare() {
    std::lock_guard<std::mutex> lock(g_abort_mu);
    if (g_abort_word) return cudaSuccess;
    cudaError_t e = cudaMallocHost((void**)&g_abort_word, 8);
    if (e == cudaSuccess) *g_abort_word = 1ull;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g_abort_stream, cudaStreamNonBlocking);
    return e;
}
extern "C" int nsx_mailbox_create(int32_t device, void** mailbox, unsigned char handle[64]) {
    if (!mailbox || !handle) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = nsx_abort_prepare();
    if (e == cudaSuccess) e = cudaMalloc(mailbox, sizeof(NsxMailbox));
    if (e == cudaSuccess) e = cudaMemset(*mailbox, 0, sizeof(NsxMailbox));
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *mailbox);
    if (e != cudaSuccess) return nsx_fail(NSX_ERR_CUDA, std::string("nsx_mailbox_create: ") + cudaGetErrorString(e));
    memcpy(handle, &h, 64);
    return 0;
}
extern "C" int nsx_mailbox_open(int32_t device, const unsigned char handle[64], void** mailbox) {
    if (!mailbox || !handle) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaIpcOpenMemHandle(mailbox, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return nsx_fail(NSX_ERR_CUDA, std::string("nsx_mailbox_open: ") + cudaGetErrorString(e));
    return 0;
}
extern "C" int nsx_mailbox_reset(int32_t device, void* mailbox) {
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMemset(mailbox, 0, sizeof(NsxMailbox));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return nsx_fail(NSX_ERR_CUDA, std::string("nsx_mailbox_reset: ") + cudaGetErrorString(e));
    return 0;
}
extern "C" int nsx_mailbox_abort(int32_t device, void* mailbox) {
    if (!mailbox) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null argument");
    cudaError_t e = cudaSetDevice(device);
    // A resident kernel is spinning on this device: the word has to get there by DMA alone - from PINNED memory (a pageable
    // source is staged through the driver and was seen to wait for the kernel), on a private non-blocking stream (the legacy
    // stream would wait as well).
    if (e == cudaSuccess) e = nsx_abort_prepare();  // (normally done by nsx_mailbox_create / _open)
    if (e == cudaSuccess) e = cudaMemcpyAsync(reinterpret_cast<unsigned char*>(mailbox) + offsetof(NsxMailbox, abort), g_abort_word, 8,
                                              cudaMemcpyHostToDevice, g_abort_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_abort_stream);
    if (e != cudaSuccess) return nsx_fail(NSX_ERR_CUDA, std::string("nsx_mailbox_abort: ") + cudaGetErrorString(e));
    return 0;
}
extern "C" int nsx_mailbox_close(int32_t device, void* mailbox, int32_t is_local) {
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = is_local ? cudaFree(mailbox) : cudaIpcCloseMemHandle(mailbox);
    if (e != cudaSuccess) return nsx_fail(NSX_ERR_CUDA, std::string("nsx_mailbox_close: ") + cudaGetErrorString(e));
    return 0;
}

extern "C" int nsx_solve_batch(int64_t count, const nsx_problem* problems, const nsx_options* opt, nsx_result* results) {
    Arena arena, inputs;
    if (count < 0 || (count > 0 && (!problems || !results)) || !opt) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null argument");
    if (count == 0) return 0;
    if (opt->row_scan_first == NSX_SPECIAL_SHORTEST_PATH)
        return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "the shortest-path rule needs a per-instance node mask; batches share one option record");
    for (int64_t i = 0; i < count; ++i) {
        int rc = nsx_validate(&problems[i], opt, &results[i]);
        if (rc) return rc;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return nsx_fail(NSX_ERR_NO_DEVICE, "no CUDA device visible (the engine has no CPU fallback)");
    if (opt->device < 0 || opt->device >= ndev) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    NSX_CUDA(cudaSetDevice(opt->device));
    DeviceInfo info;
    int rc = nsx_device_info(opt->device, info);
    if (rc) return rc;
    const bool devex = opt->pricing == NSX_PRICING_DEVEX || opt->pricing == NSX_PRICING_DEVEX_LOOP;

    // batch instances keep float64 costs; node ids are narrowed when every instance allows it.
    // Each CTA packs the store of its instance itself before the first sweep.
    bool narrow = !(getenv("NSX_LAYOUT") && !strcmp(getenv("NSX_LAYOUT"), "wide"));
    for (int64_t i = 0; i < count; ++i) if (problems[i].n_nodes - 1 > 65536) narrow = false;
    NsxStore layout;
    nsx_store_layout(layout, narrow ? NSX_NODE_U16 : NSX_NODE_I32, NSX_COST_F64, devex ? 1 : 0);
    struct Off { size_t store; size_t tail, head, pert, upper, atail, ahead, aupper, flow, state, wgt, node, depth, pi, order, tmp, sidx, gph, gpt, garc2, gres, supply, ctl, trace; };
    std::vector<Off> off(count);
    int32_t max_n = 1;
    const size_t fixed = nsx_smem_fixed();
    const size_t limit_all = info.smem_optin - fixed;
    const int want_mode = nsx_env_int("NSX_RESIDENT", 2), want_stage = nsx_env_int("NSX_STAGE_PI", 1);
    for (int64_t i = 0; i < count; ++i) if (problems[i].n_nodes > max_n) max_n = problems[i].n_nodes;
    // shared memory of every CTA: sized for the largest instance (each CTA re-plans per instance within this size)
    NsxSmemPlan plan = nsx_plan_pivot((size_t)max_n, limit_all, layout.stage_bytes, true, want_mode, want_stage);
    if (plan.stages < 2) return nsx_fail(NSX_ERR_INTERNAL, "shared memory plan exceeds the device limit");
    int max_stages = nsx_env_int("NSX_BATCH_STAGES", 4);
    if (max_stages < 2) max_stages = 2;
    if (plan.stages > max_stages) plan.stages = max_stages;
    const size_t dyn = nsx_plan_bytes(plan, layout.stage_bytes);
    const size_t smem = fixed + dyn;
    bool any_hbm = false;
    for (int64_t i = 0; i < count; ++i) {
        const nsx_problem& p = problems[i];
        const size_t n = p.n_nodes, m = p.n_arcs, ma = m + n - 1, mpad = (size_t)nsx_pad_tiles((int64_t)m);
        Off& o = off[i];
        o.tail = arena.plan((m + 4) * 4); o.head = arena.plan((m + 4) * 4);
        o.pert = arena.plan((m + 4) * 8); o.upper = arena.plan((m + 4) * 8);
        o.store = arena.plan((mpad / NSX_TILE) * layout.tile_bytes);
        o.atail = arena.plan(n * 4); o.ahead = arena.plan(n * 4); o.aupper = arena.plan(n * 8);
        o.flow = arena.plan((ma + 4) * 8); o.state = arena.plan((ma > mpad ? ma : mpad) + 16);
        o.wgt = devex ? arena.plan(mpad * 4) : 0;
        o.node = arena.plan(n * sizeof(NsxNode)); o.depth = arena.plan(n * 4); o.pi = arena.plan(n * 8 + 16);
        // (a tree too large for the CTA's shared memory keeps its preorder array in blocks: same plan as the kernel makes)
        const bool in_hbm = nsx_plan_pivot(n, dyn, layout.stage_bytes, true, want_mode, want_stage).mode == NSX_RES_NONE;
        any_hbm = any_hbm || in_hbm;
        o.order = arena.plan((in_hbm ? nsx_order_len(n) : n) * 4); o.tmp = arena.plan(n * 4); o.sidx = arena.plan(n * 4);
        o.gph = arena.plan(n * 4); o.gpt = arena.plan(n * 4);
        o.garc2 = arena.plan((2 * n + 1) * 4); o.gres = arena.plan((2 * n + 1) * 8);
        o.supply = arena.plan(n * 8); o.ctl = arena.plan(sizeof(NsxCtl));
        o.trace = (results[i].entering_trace && opt->trace_capacity > 0) ? arena.plan((size_t)opt->trace_capacity * 4) : 0;
    }
    size_t o_items = arena.plan(sizeof(NsxBatchItem) * (size_t)count);
    size_t o_next = arena.plan(8);
    NSX_CUDA(arena.commit());

    CallResources rs;
    NSX_CUDA(rs.create());
    cudaStream_t stream = rs.stream;
    cudaEvent_t* ev = rs.ev;
    std::vector<NsxBatchItem> items(count);
    std::vector<NsxCtl> ctls(count);
    NSX_CUDA(cudaEventRecord(ev[0], stream));
    NSX_CUDA(cudaMemsetAsync(arena.base, 0, arena.size, stream));  // tile padding reads as "no arc"
    for (int64_t i = 0; i < count; ++i) {
        const nsx_problem& p = problems[i];
        const size_t n = p.n_nodes, m = p.n_arcs;
        const Off& o = off[i];
        NsxDev& d = items[i].d;
        const bool tr = o.trace != 0;
        d.n = (int32_t)n; d.m = (int64_t)m; d.ma = (int64_t)(m + n - 1);
        d.tail = arena.at<int32_t>(o.tail); d.head = arena.at<int32_t>(o.head);
        d.pert = arena.at<double>(o.pert); d.upper = arena.at<double>(o.upper);
        d.atail = arena.at<int32_t>(o.atail); d.ahead = arena.at<int32_t>(o.ahead); d.aupper = arena.at<double>(o.aupper);
        d.flow = arena.at<double>(o.flow); d.state = arena.at<uint8_t>(o.state);
        d.wgt = devex ? arena.at<uint32_t>(o.wgt) : nullptr;
        d.node = arena.at<NsxNode>(o.node); d.depth = arena.at<int32_t>(o.depth); d.pi = arena.at<double>(o.pi); d.pi_mirror = nullptr;
        d.order = arena.at<int32_t>(o.order); d.tmp = arena.at<int32_t>(o.tmp);
        d.gpath_h = arena.at<int32_t>(o.gph); d.gpath_t = arena.at<int32_t>(o.gpt);
        d.garc2 = arena.at<int32_t>(o.garc2); d.gres = arena.at<double>(o.gres);
        d.penalty = p.penalty; d.tol = opt->tolerance; d.scan_walk = 0; d.par16 = nullptr; d.root_bits = nullptr;
        d.node_mask = nullptr; d.imbalance = nullptr;
        d.blk = nullptr; d.sidx = arena.at<int32_t>(o.sidx);
        d.dinfo = nullptr; d.rc = nullptr; d.dlist = nullptr; d.dstamp = nullptr; d.row_begin = nullptr; d.col_begin = nullptr;
        d.csc_pos = nullptr; d.csc_state = nullptr; d.csc_wgt = nullptr;
        items[i].st = layout;
        items[i].st.base = arena.at<unsigned char>(o.store);
        items[i].mpad = nsx_pad_tiles((int64_t)m);
        items[i].ctl = arena.at<NsxCtl>(o.ctl);
        items[i].trace = tr ? arena.at<int32_t>(o.trace) : nullptr;
        items[i].supply = arena.at<double>(o.supply);
        if (m > 0) {
            NSX_CUDA(cudaMemcpyAsync((void*)d.tail, p.tail, m * 4, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync((void*)d.head, p.head, m * 4, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync((void*)d.pert, p.pert_cost, m * 8, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync((void*)d.upper, p.upper, m * 8, cudaMemcpyHostToDevice, stream));
        }
        NSX_CUDA(cudaMemcpyAsync((void*)items[i].supply, p.supply, n * 8, cudaMemcpyHostToDevice, stream));
        nsx_fill_ctl(ctls[i], opt, tr);
        NSX_CUDA(cudaMemcpyAsync(items[i].ctl, &ctls[i], sizeof(NsxCtl), cudaMemcpyHostToDevice, stream));
    }
    NsxBatchItem* d_items = arena.at<NsxBatchItem>(o_items);
    unsigned long long* d_next = arena.at<unsigned long long>(o_next);
    NSX_CUDA(cudaMemcpyAsync(d_items, items.data(), sizeof(NsxBatchItem) * (size_t)count, cudaMemcpyHostToDevice, stream));
    NSX_CUDA(cudaMemsetAsync(d_next, 0, 8, stream));
    NSX_CUDA(cudaEventRecord(ev[1], stream));

    auto* batch_kernel = any_hbm ? nsx_batch_kernel_hbm : nsx_batch_kernel;
    NSX_CUDA(cudaFuncSetAttribute(batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    NSX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, batch_kernel, NSX_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)info.sms * per_sm;
    if (grid > count) grid = count;
    batch_kernel<<<(int)grid, NSX_THREADS, smem, stream>>>(d_items, count, d_next, dyn, want_mode, want_stage);
    NSX_CUDA(cudaGetLastError());
    NSX_CUDA(cudaEventRecord(ev[2], stream));
    for (int64_t i = 0; i < count; ++i) {
        const NsxDev& d = items[i].d;
        nsx_result& r = results[i];
        if (r.flow) NSX_CUDA(cudaMemcpyAsync(r.flow, d.flow, (size_t)d.ma * 8, cudaMemcpyDeviceToHost, stream));
        if (r.potential) NSX_CUDA(cudaMemcpyAsync(r.potential, d.pi, (size_t)d.n * 8, cudaMemcpyDeviceToHost, stream));
        if (r.state) NSX_CUDA(cudaMemcpyAsync(r.state, d.state, (size_t)d.ma, cudaMemcpyDeviceToHost, stream));
        NSX_CUDA(cudaMemcpyAsync(&ctls[i], items[i].ctl, sizeof(NsxCtl), cudaMemcpyDeviceToHost, stream));
    }
    NSX_CUDA(cudaEventRecord(ev[3], stream));
    NSX_CUDA(cudaStreamSynchronize(stream));
    float h2d = 0, solve = 0, d2h = 0;
    cudaEventElapsedTime(&h2d, ev[0], ev[1]);
    cudaEventElapsedTime(&solve, ev[1], ev[2]);
    cudaEventElapsedTime(&d2h, ev[2], ev[3]);
    int bad = 0, bad_ids = 0;
    for (int64_t i = 0; i < count; ++i) {
        nsx_result& r = results[i];
        nsx_harvest(ctls[i], &r);
        r.h2d_ms = h2d; r.solve_ms = solve; r.d2h_ms = d2h; r.grid_ctas = (int32_t)grid;
        r.bytes_per_arc = (int32_t)(2 * nsx_node_bytes(layout.node_kind) + 9); r.ring_stages = plan.stages; r.resident_mode = plan.mode; r.store_layout = layout.node_kind | (layout.cost_kind << 8);
        if (items[i].trace) {
            int64_t cnt = ctls[i].trace_len < opt->trace_capacity ? ctls[i].trace_len : opt->trace_capacity;
            if (cnt > 0) NSX_CUDA(cudaMemcpy(r.entering_trace, items[i].trace, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
        }
        if (ctls[i].status < 0) { bad++; if (ctls[i].fault == 5) bad_ids++; }
    }
    arena.release();
    if (bad_ids) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "arc endpoint outside 1 .. n_nodes-1 in a batch instance (node 0 is the artificial root)");
    if (bad) return nsx_fail(NSX_ERR_INTERNAL, "batch kernel left instances without a status");
    return 0;
}

extern "C" const char* nsx_last_error(void) { return g_last_error.c_str(); }
extern "C" void nsx_version(int32_t* abi, int32_t* sm_arch) {
    if (abi) *abi = NSX_ABI_VERSION;
    if (sm_arch) *sm_arch = 100;
}
extern "C" int nsx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
