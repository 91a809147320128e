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

#include <string>
#include <vector>

#include "nsx_core.cuh"

#ifndef NSX_THREADS
#define NSX_THREADS 512
#endif
#define NSX_PI_SMEM_MAX_NODES 12288  // node potentials staged in shared memory up to this many nodes

// ------------------------------------------------------------------------------------------
// Grid-wide command / arrival handshake
// ------------------------------------------------------------------------------------------
struct NsxGridCtl {
    int32_t seq;  // command sequence number, release-published by CTA 0
    int32_t pad;
    unsigned long long arrived;  // CTAs (other than 0) that delivered their candidate, cumulative
    NsxCmd cmd;
};

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
__device__ __forceinline__ void nsx_st_release(int32_t* p, int32_t v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit (cp.async.bulk), completion on an mbarrier.
__device__ __forceinline__ uint32_t nsx_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void nsx_mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nsx_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void nsx_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nsx_smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(nsx_smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(nsx_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void nsx_mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NSX_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NSX_DONE;\n"
        "bra NSX_WAIT;\n"
        "NSX_DONE:\n"
        "}\n" ::"r"(nsx_smem_addr(bar)), "r"(parity) : "memory");
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
// Pricing sweep over [cmd.lo, cmd.hi): thread `g` of `T` takes quads (4 consecutive arcs)
// g, g+T, ... so that a warp streams 128 consecutive arcs with 128-bit loads.
// Algorithmic bytes per arc: tail 4 + head 4 + cost 8 + state 1 = 17 (Dantzig / row scan),
// + 4 Devex weight = 21.  Potentials are gathered from `pis` (shared memory) when staged,
// otherwise from L2 (ld.global.cg: they are rewritten between sweeps by the pivot CTA).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double nsx_pi_at(const NsxDev& d, const double* pis, int32_t v) {
    return pis ? pis[v] : __ldcg(d.pi + v);
}

template <bool DEVEX>
__device__ __forceinline__ void nsx_price_arc(const NsxDev& d, const NsxCmd& cmd, const double* pis,
                                              int32_t i, int32_t tl, int32_t hd, double pert,
                                              uint32_t st, uint32_t wraw, NsxCand& dz, NsxDevexCand& dx) {
    if ((st & NSX_ARC_IN_TREE) || !(st & (NSX_ARC_CAN_FWD | NSX_ARC_CAN_BWD))) return;
    double cost = pert;
    if (!DEVEX && cmd.phase == 1) cost = NSX_SUB(NSX_SUB(pert, 1.0), NSX_MUL(1e-6, (double)i));
    double rc = NSX_SUB(NSX_ADD(cost, nsx_pi_at(d, pis, tl)), nsx_pi_at(d, pis, hd));
    if (DEVEX) {
        if (i == cmd.excluded) return;
        nsx_price_devex(dx, i, (uint8_t)st, rc, wraw, cmd.wepoch, d.tol);
    } else {
        nsx_price_dantzig(dz, i, (uint8_t)st, rc, d.tol);
    }
}

// One quad (4 consecutive arcs, 16-byte aligned columns) held in registers.
struct NsxQuad {
    int4 t, h;
    double2 c0, c1;
    uint32_t st;
    uint4 w;
};

template <bool DEVEX>
__device__ __forceinline__ void nsx_load_quad(const NsxDev& d, int64_t base, NsxQuad& q) {
    q.t = __ldg((const int4*)(d.tail + base));
    q.h = __ldg((const int4*)(d.head + base));
    q.c0 = __ldg((const double2*)(d.pert + base));
    q.c1 = __ldg((const double2*)(d.pert + base + 2));
    q.st = __ldcg((const unsigned int*)(d.state + base));
    if (DEVEX) q.w = __ldcg((const uint4*)(d.wgt + base)); else q.w = make_uint4(1u, 1u, 1u, 1u);
}

template <bool DEVEX>
__device__ __forceinline__ void nsx_price_quad(const NsxDev& d, const NsxCmd& cmd, const double* pis,
                                               int64_t base, const NsxQuad& q, NsxCand& dz,
                                               NsxDevexCand& dx) {
    // fast reject: nothing to do when every arc of the quad is in the tree or has no residual
    const uint32_t st = q.st;
    const uint32_t live = ((st >> 1) | (st >> 2)) & ~st & 0x01010101u;
    if (!live) return;
    nsx_price_arc<DEVEX>(d, cmd, pis, (int32_t)base + 0, q.t.x, q.h.x, q.c0.x, st & 0xffu, q.w.x, dz, dx);
    nsx_price_arc<DEVEX>(d, cmd, pis, (int32_t)base + 1, q.t.y, q.h.y, q.c0.y, (st >> 8) & 0xffu, q.w.y, dz, dx);
    nsx_price_arc<DEVEX>(d, cmd, pis, (int32_t)base + 2, q.t.z, q.h.z, q.c1.x, (st >> 16) & 0xffu, q.w.z, dz, dx);
    nsx_price_arc<DEVEX>(d, cmd, pis, (int32_t)base + 3, q.t.w, q.h.w, q.c1.y, (st >> 24) & 0xffu, q.w.w, dz, dx);
}

template <bool DEVEX>
__device__ __forceinline__ void nsx_price_scalar(const NsxDev& d, const NsxCmd& cmd, const double* pis,
                                                 int64_t lo, int64_t hi, NsxCand& dz, NsxDevexCand& dx) {
    for (int64_t i = lo; i < hi; ++i)
        nsx_price_arc<DEVEX>(d, cmd, pis, (int32_t)i, __ldg(d.tail + i), __ldg(d.head + i),
                             __ldg(d.pert + i), __ldcg(d.state + i), DEVEX ? __ldcg(d.wgt + i) : 1u, dz, dx);
}

#ifndef NSX_UNROLL
#define NSX_UNROLL 2
#endif

// Thread g of T handles quads g, g+T, g+2T, ... of the aligned interior; NSX_UNROLL quads are
// loaded before the first one is priced so that several 128-bit requests are in flight per thread.
template <bool DEVEX>
__device__ __forceinline__ void nsx_sweep(const NsxDev& d, const NsxCmd& cmd, const double* pis,
                                          int64_t g, int64_t T, NsxCand& dz, NsxDevexCand& dx) {
    const int64_t lo = cmd.lo, hi = cmd.hi;
    const int64_t a0 = (lo + 3) & ~(int64_t)3;          // first aligned arc
    const int64_t a1 = hi & ~(int64_t)3;                // end of the aligned interior
    if (a1 <= a0) {                                     // tiny range: scalar only
        if (g == 0) nsx_price_scalar<DEVEX>(d, cmd, pis, lo, hi, dz, dx);
        return;
    }
    if (g == 0) {                                       // ragged head / tail (at most 3 arcs each)
        nsx_price_scalar<DEVEX>(d, cmd, pis, lo, a0, dz, dx);
        nsx_price_scalar<DEVEX>(d, cmd, pis, a1, hi, dz, dx);
    }
    const int64_t nq = (a1 - a0) >> 2;
    int64_t q = g;
    for (; q + (NSX_UNROLL - 1) * T < nq; q += NSX_UNROLL * T) {
        NsxQuad quad[NSX_UNROLL];
#pragma unroll
        for (int u = 0; u < NSX_UNROLL; ++u) nsx_load_quad<DEVEX>(d, a0 + ((q + u * T) << 2), quad[u]);
#pragma unroll
        for (int u = 0; u < NSX_UNROLL; ++u)
            nsx_price_quad<DEVEX>(d, cmd, pis, a0 + ((q + u * T) << 2), quad[u], dz, dx);
    }
    for (; q < nq; q += T) {
        NsxQuad quad;
        nsx_load_quad<DEVEX>(d, a0 + (q << 2), quad);
        nsx_price_quad<DEVEX>(d, cmd, pis, a0 + (q << 2), quad, dz, dx);
    }
}

// Shared-memory layout of one CTA: fixed part, then (dynamic) the staged / resident node state.
struct NsxCtaShared {
    NsxLoopShared L;
    NsxCtl ctl;  // solver scalars live here during the solve (copied in / out of HBM once)
    NsxCmd cmd;  // worker copy of the command
    NsxCand dz_buf[32];
    NsxDevexCand dx_buf[32];
    unsigned long long mbar;  // completion barrier of the potentials bulk copy
    NsxPivotScratch piv;
    NsxPotScratch pot;
};

// How much node state the pivot CTA keeps in shared memory (chosen by the host from n and the
// opt-in shared-memory limit): the cycle walk is a chain of dependent 16-byte record loads, so
// holding the records on-chip turns ~L2 latency per hop into shared-memory latency.
enum { NSX_RES_NONE = 0,   // tree in HBM/L2; potentials optionally staged per sweep
       NSX_RES_NODES = 1,  // node records + potentials resident (24 B / node)
       NSX_RES_ALL = 2 };  // + depth, preorder array, permutation scratch (36 B / node)

struct NsxSmemPlan {
    int32_t mode;      // NSX_RES_*
    int32_t stage_pi;  // workers (and NSX_RES_NONE pivot CTAs) copy pi into shared memory per sweep
};

__device__ __forceinline__ size_t nsx_align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Redirect the node arrays of `d` into shared memory according to the plan and fill them.
__device__ __forceinline__ NsxDev nsx_make_resident(const NsxDev& d, const NsxSmemPlan plan,
                                                    unsigned char* dyn, double** pis_out) {
    NsxDev dl = d;
    double* pis = reinterpret_cast<double*>(dyn);
    *pis_out = pis;
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
    NSX_SYNC();
    return dl;
}

// One sweep of this CTA: stage potentials (optional), price, block-reduce into thread 0.
// `stage_count` is a per-thread register copy of the number of bulk copies issued so far by this
// CTA (all threads count alike); its low bit is the mbarrier phase to wait for.
__device__ __forceinline__ void nsx_cta_sweep(const NsxDev& d, const NsxCmd& cmd, const double* pi_src,
                                              double* pis, bool stage, uint32_t& stage_count, int64_t g,
                                              int64_t T, NsxCtaShared& sh, NsxCand& dz, NsxDevexCand& dx) {
    nsx_cand_init(dz);
    nsx_devex_init(dx);
    if (stage) {
        // refresh the shared-memory copy of the node potentials: one TMA bulk copy, every thread
        // waits on the mbarrier phase (no register staging, no per-thread loads)
        NSX_SYNC();
        if (threadIdx.x == 0)
            nsx_bulk_load(pis, pi_src, (uint32_t)(((size_t)d.n * 8 + 15) & ~(size_t)15), &sh.mbar);
        nsx_mbar_wait(&sh.mbar, stage_count & 1u);
        stage_count++;
    }
    if (cmd.kind == NSX_CMD_DEVEX) {
        nsx_sweep<true>(d, cmd, pis, g, T, dz, dx);
        nsx_block_reduce(dx, sh.dx_buf);
    } else {
        nsx_sweep<false>(d, cmd, pis, g, T, dz, dx);
        nsx_block_reduce(dz, sh.dz_buf);
    }
}

// Sweep functor of CTA 0 in the grid-resident kernel.
struct GridSweep {
    const NsxDev& d;      // global view (arc arrays, global pi mirror)
    NsxGridCtl* g;
    NsxCand* dzc;
    NsxDevexCand* dxc;
    NsxCtaShared& sh;
    double* pis;          // potentials in shared memory (resident master copy, or staging buffer), or null
    bool stage;           // this CTA must refresh `pis` from HBM before each sweep
    uint32_t& stage_count;
    int32_t seq;
    unsigned long long target;
    unsigned long long t_price, t_sync;

    __device__ void publish(const NsxCmd& cmd) {
        if (threadIdx.x == 0) {
            g->cmd = cmd;
            __threadfence();
            nsx_st_release(&g->seq, ++seq);
        }
    }
    __device__ void run(const NsxCmd& cmd_in, NsxCand& out_dz, NsxDevexCand& out_dx) {
        unsigned long long t0 = 0;
        if (threadIdx.x == 0) t0 = nsx_globaltimer();
        NSX_SYNC();  // pivot writes of all threads precede thread 0's fence + release
        if (gridDim.x > 1) publish(cmd_in);
        const NsxCmd cmd = cmd_in;
        NsxCand dz; NsxDevexCand dx;
        nsx_cand_init(dz); nsx_devex_init(dx);
        // with workers present this CTA only pivots and merges; alone it prices everything itself
        if (gridDim.x == 1)
            nsx_cta_sweep(d, cmd, d.pi, pis, stage, stage_count, (int64_t)threadIdx.x, (int64_t)blockDim.x, sh, dz, dx);
        if (gridDim.x == 1) {
            if (threadIdx.x == 0) {
                if (cmd.kind == NSX_CMD_DEVEX) out_dx = dx; else out_dz = dz;
                t_price += nsx_globaltimer() - t0;
            }
            NSX_SYNC();
            return;
        }
        if (threadIdx.x == 0) {
            unsigned long long t1 = nsx_globaltimer();
            target += gridDim.x - 1;
            while (nsx_ld_acquire_u64(&g->arrived) < target) { }
            __threadfence();
            t_sync += nsx_globaltimer() - t1;
        }
        NSX_SYNC();
        // merge the candidates of the other CTAs (one per thread), then reduce across the block
        if (cmd.kind == NSX_CMD_DEVEX) {
            NsxDevexCand k; nsx_devex_init(k);
            if (threadIdx.x == 0) k = dx;
            for (int b = 1 + threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
                const int4* src = (const int4*)(dxc + b);
                union { NsxDevexCand c; int4 v[2]; } tmp;
                tmp.v[0] = __ldcg(src); tmp.v[1] = __ldcg(src + 1);
                nsx_devex_merge(k, tmp.c);
            }
            nsx_block_reduce(k, sh.dx_buf);
            if (threadIdx.x == 0) out_dx = k;
        } else {
            NsxCand k; nsx_cand_init(k);
            if (threadIdx.x == 0) k = dz;
            for (int b = 1 + threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
                union { NsxCand c; int4 v; } tmp;
                tmp.v = __ldcg((const int4*)(dzc + b));
                nsx_cand_merge(k, tmp.c);
            }
            nsx_block_reduce(k, sh.dz_buf);
            if (threadIdx.x == 0) out_dz = k;
        }
        if (threadIdx.x == 0) t_price += nsx_globaltimer() - t0;
        NSX_SYNC();
    }
    __device__ void finish() {
        NsxCmd cmd;
        cmd.kind = NSX_CMD_EXIT; cmd.phase = 0; cmd.lo = cmd.hi = 0; cmd.excluded = -1; cmd.wepoch = 0;
        NSX_SYNC();
        if (gridDim.x > 1) publish(cmd);
    }
};

static_assert(sizeof(NsxCand) == 16, "NsxCand is moved as one int4");
static_assert(sizeof(NsxCmd) == 32, "NsxCmd is moved as two int4");
static_assert(offsetof(NsxGridCtl, cmd) % 16 == 0, "command block must be 16-byte aligned");
static_assert(sizeof(NsxDevexCand) == 32, "NsxDevexCand is moved as two int4");

struct NsxKernelArgs {
    NsxDev d;
    NsxCtl* ctl;
    NsxGridCtl* grid;
    NsxCand* dzc;
    NsxDevexCand* dxc;
    int32_t* trace;
    NsxSmemPlan plan;
};

__device__ __forceinline__ void nsx_copy_ctl(NsxCtl* dst, const NsxCtl* src) {
    const int32_t* s = reinterpret_cast<const int32_t*>(src);
    int32_t* t = reinterpret_cast<int32_t*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(NsxCtl) / 4); i += blockDim.x) t[i] = s[i];
}

extern "C" __global__ void __launch_bounds__(NSX_THREADS, 1)
nsx_resident_kernel(const NsxKernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NsxCtaShared& sh = *reinterpret_cast<NsxCtaShared*>(smem_raw);
    unsigned char* dyn = smem_raw + nsx_align16(sizeof(NsxCtaShared));
    const NsxDev& d = a.d;
    if (threadIdx.x == 0) nsx_mbar_init(&sh.mbar, 1);
    uint32_t stage_count = 0;
    NSX_SYNC();

    if (blockIdx.x == 0) {
        unsigned long long t_begin = 0;
        if (threadIdx.x == 0) t_begin = nsx_globaltimer();
        nsx_copy_ctl(&sh.ctl, a.ctl);
        double* pis = nullptr;
        const NsxDev dl = nsx_make_resident(d, a.plan, dyn, &pis);
        NSX_SYNC();
        const bool resident = a.plan.mode != NSX_RES_NONE;
        GridSweep sweep{d, a.grid, a.dzc, a.dxc, sh, (resident || a.plan.stage_pi) ? pis : nullptr,
                        !resident && a.plan.stage_pi != 0, stage_count, 0, 0ull, 0ull, 0ull};
        nsx_solve_loop(dl, sh.ctl, sh.L, sh.piv, sh.pot, a.trace, sweep);
        NSX_SYNC();
        if (threadIdx.x == 0) {
            unsigned long long total = nsx_globaltimer() - t_begin;
            sh.ctl.clk_pricing = (int64_t)sweep.t_price;
            sh.ctl.clk_sync = (int64_t)sweep.t_sync;
            sh.ctl.clk_pivot = (int64_t)(total - sweep.t_price);
        }
        NSX_SYNC();
        nsx_copy_ctl(a.ctl, &sh.ctl);
        return;
    }
    // worker CTAs: wait for a command, price, deliver, repeat
    double* pis = a.plan.stage_pi ? reinterpret_cast<double*>(dyn) : nullptr;
    int32_t seen = 0;
    for (;;) {
        if (threadIdx.x == 0) {
            int32_t s;
            while ((s = nsx_ld_acquire(&a.grid->seq)) == seen) { __nanosleep(32); }
            seen = s;
            __threadfence();
            union { NsxCmd c; int4 v[2]; } tmp;
            tmp.v[0] = __ldcg((const int4*)&a.grid->cmd);
            tmp.v[1] = __ldcg(((const int4*)&a.grid->cmd) + 1);
            sh.cmd = tmp.c;
        }
        NSX_SYNC();
        const NsxCmd cmd = sh.cmd;
        if (cmd.kind == NSX_CMD_EXIT) return;
        NsxCand dz; NsxDevexCand dx;
        nsx_cta_sweep(d, cmd, d.pi, pis, pis != nullptr, stage_count, (int64_t)(blockIdx.x - 1) * blockDim.x + threadIdx.x,
                      (int64_t)(gridDim.x - 1) * blockDim.x, sh, dz, dx);
        if (threadIdx.x == 0) {
            if (cmd.kind == NSX_CMD_DEVEX) a.dxc[blockIdx.x] = dx; else a.dzc[blockIdx.x] = dz;
            __threadfence();
            atomicAdd(&a.grid->arrived, 1ull);
        }
        NSX_SYNC();
    }
}

// Initial state: real arcs, nodes + artificial arcs, artificial-flow count.
extern "C" __global__ void nsx_init_kernel(const NsxDev d, const double* supply, NsxCtl* ctl) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = g; i < d.m; i += T) nsx_init_real_arc(d, i);
    unsigned long long art = 0;
    for (int64_t v = g; v < d.n; v += T) {
        nsx_init_node(d, (int32_t)v, supply[v]);
        if (v > 0 && d.flow[d.m + v - 1] > d.tol) art++;
    }
    if (art) atomicAdd((unsigned long long*)&ctl->art_with_flow, art);
}

// ------------------------------------------------------------------------------------------
// Batched variant: one CTA solves one independent instance end to end (config 4).
// ------------------------------------------------------------------------------------------
struct NsxBatchItem {
    NsxDev d;
    NsxCtl* ctl;
    int32_t* trace;
    const double* supply;
};

struct LocalSweep {
    const NsxDev& d;
    NsxCtaShared& sh;
    double* pis;
    bool stage;
    uint32_t& stage_count;
    __device__ void run(const NsxCmd& cmd_in, NsxCand& out_dz, NsxDevexCand& out_dx) {
        const NsxCmd cmd = cmd_in;
        NsxCand dz; NsxDevexCand dx;
        nsx_cta_sweep(d, cmd, d.pi, pis, stage, stage_count, (int64_t)threadIdx.x, (int64_t)blockDim.x, sh, dz, dx);
        if (threadIdx.x == 0) { if (cmd.kind == NSX_CMD_DEVEX) out_dx = dx; else out_dz = dz; }
        NSX_SYNC();
    }
    __device__ void finish() {}
};

// per-item smem need is decided on the host from the largest instance; smaller instances simply
// use less of it.  `limit_bytes` = dynamic bytes available after the fixed part.
extern "C" __global__ void __launch_bounds__(NSX_THREADS, 1)
nsx_batch_kernel(const NsxBatchItem* items, int64_t count, unsigned long long* next, size_t limit_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NsxCtaShared& sh = *reinterpret_cast<NsxCtaShared*>(smem_raw);
    unsigned char* dyn = smem_raw + nsx_align16(sizeof(NsxCtaShared));
    __shared__ unsigned long long my_item;
    if (threadIdx.x == 0) nsx_mbar_init(&sh.mbar, 1);
    uint32_t stage_count = 0;
    for (;;) {
        NSX_SYNC();
        if (threadIdx.x == 0) my_item = atomicAdd(next, 1ull);
        NSX_SYNC();
        const unsigned long long it = my_item;
        if (it >= (unsigned long long)count) return;
        const NsxBatchItem& item = items[it];
        const NsxDev& d = item.d;
        for (int64_t i = threadIdx.x; i < d.m; i += blockDim.x) nsx_init_real_arc(d, i);
        for (int64_t v = threadIdx.x; v < d.n; v += blockDim.x) nsx_init_node(d, (int32_t)v, item.supply[v]);
        nsx_copy_ctl(&sh.ctl, item.ctl);
        NSX_SYNC();
        if (threadIdx.x == 0) {
            int64_t art = 0;
            for (int32_t v = 1; v < d.n; ++v) art += d.flow[d.m + v - 1] > d.tol;
            sh.ctl.art_with_flow = art;
        }
        NSX_SYNC();
        NsxSmemPlan plan;
        const size_t n = (size_t)d.n;
        plan.mode = (36 * n + 64 <= limit_bytes) ? NSX_RES_ALL : (24 * n + 64 <= limit_bytes) ? NSX_RES_NODES : NSX_RES_NONE;
        plan.stage_pi = (plan.mode == NSX_RES_NONE && 8 * n + 64 <= limit_bytes) ? 1 : 0;
        double* pis = nullptr;
        const NsxDev dl = nsx_make_resident(d, plan, dyn, &pis);
        NSX_SYNC();
        const bool resident = plan.mode != NSX_RES_NONE;
        LocalSweep sweep{d, sh, (resident || plan.stage_pi) ? pis : nullptr, !resident && plan.stage_pi != 0, stage_count};
        nsx_solve_loop(dl, sh.ctl, sh.L, sh.piv, sh.pot, item.trace, sweep);
        NSX_SYNC();
        nsx_copy_ctl(item.ctl, &sh.ctl);
    }
}

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
            arena.release();                                                                    \
            return nsx_fail(NSX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__)); \
        }                                                                                       \
    } while (0)

// One cudaMalloc per call; sub-allocations are 256-byte aligned.
struct Arena {
    unsigned char* base = nullptr;
    size_t size = 0, used = 0;
    size_t plan(size_t bytes) { size_t off = size; size += (bytes + 255) & ~(size_t)255; return off; }
    cudaError_t commit() { return cudaMalloc((void**)&base, size ? size : 256); }
    template <class T> T* at(size_t off) { return reinterpret_cast<T*>(base + off); }
    void release() { if (base) { cudaFree(base); base = nullptr; } }
};

static int nsx_validate(const nsx_problem* p, const nsx_options* o, const nsx_result* r) {
    if (!p || !o || !r) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null argument");
    if (p->n_nodes < 1 || p->n_arcs < 0) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "bad sizes");
    if (p->n_arcs + p->n_nodes >= (1ll << 30)) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "arc count exceeds 2^30");
    if (p->n_arcs > 0 && (!p->tail || !p->head || !p->pert_cost || !p->upper))
        return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null arc array");
    if (!p->supply) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null supply");
    if (o->pricing != NSX_PRICING_DANTZIG && o->pricing != NSX_PRICING_DEVEX)
        return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "unknown pricing rule");
    if (o->max_iterations < 0 || !(o->tolerance > 0) || o->ft_update_limit <= 0)
        return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "bad option value");
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
}

static void nsx_harvest(const NsxCtl& c, nsx_result* res) {
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
    res->sum_cycle_len = c.sum_cycle;
    res->sum_subtree = c.sum_subtree;
    res->max_subtree = c.max_subtree;
    res->sum_rounds = c.sum_rounds;
    res->pricing_ms = (double)c.clk_pricing * 1e-6;
    res->pivot_ms = (double)c.clk_pivot * 1e-6;
    res->sync_ms = (double)c.clk_sync * 1e-6;
    for (int i = 0; i < 12; ++i) res->phase_cycles[i] = c.ph[i];
}

struct DeviceInfo { int sms = 0; int coop = 0; size_t smem_optin = 0; bool ok = false; };
static int nsx_device_info(int dev, DeviceInfo& info) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return nsx_fail(NSX_ERR_NO_DEVICE, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    info.sms = prop.multiProcessorCount;
    info.coop = prop.cooperativeLaunch;
    info.smem_optin = prop.sharedMemPerBlockOptin;
    info.ok = true;
    return 0;
}

static int nsx_env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

static size_t nsx_smem_fixed() { return (sizeof(NsxCtaShared) + 15) & ~(size_t)15; }
// Choose how much node state lives in shared memory for an n-node instance under `limit` bytes.
static NsxSmemPlan nsx_plan_smem(int32_t n, size_t limit, size_t* bytes) {
    NsxSmemPlan plan; plan.mode = NSX_RES_NONE; plan.stage_pi = 0;
    const size_t fixed = nsx_smem_fixed(), nn = (size_t)n;
    const int want = nsx_env_int("NSX_RESIDENT", 2);
    size_t dyn = 0;
    if (want >= 2 && fixed + 36 * nn + 64 <= limit) { plan.mode = NSX_RES_ALL; plan.stage_pi = 1; dyn = 36 * nn + 64; }
    else if (want >= 1 && fixed + 24 * nn + 64 <= limit) { plan.mode = NSX_RES_NODES; plan.stage_pi = 1; dyn = 24 * nn + 64; }
    else if (nsx_env_int("NSX_STAGE_PI", 1) != 0 && fixed + 8 * nn + 64 <= limit && n <= NSX_PI_SMEM_MAX_NODES) {
        plan.stage_pi = 1; dyn = 8 * nn + 64;
    }
    *bytes = fixed + dyn;
    return plan;
}

// Common implementation; `resident` = arc arrays are device pointers.
static int nsx_solve_impl(const nsx_problem* pb, const nsx_options* opt, nsx_result* res, bool resident) {
    Arena arena;
    int rc = nsx_validate(pb, opt, res);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return nsx_fail(NSX_ERR_NO_DEVICE, "no CUDA device visible (the engine has no CPU fallback)");
    if (opt->device < 0 || opt->device >= ndev) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    NSX_CUDA(cudaSetDevice(opt->device));
    DeviceInfo info;
    if ((rc = nsx_device_info(opt->device, info))) return rc;
    if (!info.coop) return nsx_fail(NSX_ERR_NO_DEVICE, "device lacks cooperative launch");

    const int32_t n = pb->n_nodes;
    const int64_t m = pb->n_arcs, ma = m + n - 1;
    const bool devex = opt->pricing == NSX_PRICING_DEVEX;
    const bool want_trace = res->entering_trace && opt->trace_capacity > 0;

    // ---- device memory plan ----
    size_t o_tail = 0, o_head = 0, o_pert = 0, o_upper = 0;
    if (!resident) {
        o_tail = arena.plan((size_t)(m + 4) * 4); o_head = arena.plan((size_t)(m + 4) * 4);
        o_pert = arena.plan((size_t)(m + 4) * 8); o_upper = arena.plan((size_t)(m + 4) * 8);
    }
    size_t o_atail = arena.plan((size_t)n * 4), o_ahead = arena.plan((size_t)n * 4), o_aupper = arena.plan((size_t)n * 8);
    size_t o_flow = arena.plan((size_t)(ma + 4) * 8), o_state = arena.plan((size_t)ma + 16);
    size_t o_wgt = devex ? arena.plan((size_t)(m + 4) * 4) : 0;
    size_t o_node = arena.plan((size_t)n * sizeof(NsxNode)), o_depth = arena.plan((size_t)n * 4);
    size_t o_pi = arena.plan((size_t)n * 8 + 16), o_order = arena.plan((size_t)n * 4), o_tmp = arena.plan((size_t)n * 4);
    size_t o_gph = arena.plan((size_t)n * 4), o_gpt = arena.plan((size_t)n * 4);
    size_t o_garc2 = arena.plan(((size_t)2 * n + 1) * 4), o_gres = arena.plan(((size_t)2 * n + 1) * 8);
    size_t o_supply = arena.plan((size_t)n * 8);
    size_t o_ctl = arena.plan(sizeof(NsxCtl)), o_grid = arena.plan(sizeof(NsxGridCtl));
    size_t o_dzc = arena.plan(sizeof(NsxCand) * 1024), o_dxc = arena.plan(sizeof(NsxDevexCand) * 1024);
    size_t o_trace = want_trace ? arena.plan((size_t)opt->trace_capacity * 4) : 0;
    NSX_CUDA(arena.commit());

    cudaStream_t stream;
    NSX_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cudaEvent_t ev[4];
    for (auto& e : ev) NSX_CUDA(cudaEventCreate(&e));

    NsxKernelArgs ka;
    NsxDev& d = ka.d;
    d.n = n; d.m = m; d.ma = ma;
    NSX_CUDA(cudaEventRecord(ev[0], stream));
    if (resident) {
        d.tail = pb->tail; d.head = pb->head; d.pert = pb->pert_cost; d.upper = pb->upper;
        if (((uintptr_t)d.tail | (uintptr_t)d.head | (uintptr_t)d.pert | (uintptr_t)d.upper) & 15) {
            arena.release();
            return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "resident arc arrays must be 16-byte aligned");
        }
    } else {
        d.tail = arena.at<int32_t>(o_tail); d.head = arena.at<int32_t>(o_head);
        d.pert = arena.at<double>(o_pert); d.upper = arena.at<double>(o_upper);
        if (m > 0) {
            NSX_CUDA(cudaMemcpyAsync((void*)d.tail, pb->tail, (size_t)m * 4, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync((void*)d.head, pb->head, (size_t)m * 4, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync((void*)d.pert, pb->pert_cost, (size_t)m * 8, cudaMemcpyHostToDevice, stream));
            NSX_CUDA(cudaMemcpyAsync((void*)d.upper, pb->upper, (size_t)m * 8, cudaMemcpyHostToDevice, stream));
        }
    }
    double* d_supply = arena.at<double>(o_supply);
    NSX_CUDA(cudaMemcpyAsync(d_supply, pb->supply, (size_t)n * 8, cudaMemcpyHostToDevice, stream));
    d.atail = arena.at<int32_t>(o_atail); d.ahead = arena.at<int32_t>(o_ahead); d.aupper = arena.at<double>(o_aupper);
    d.flow = arena.at<double>(o_flow); d.state = arena.at<uint8_t>(o_state);
    d.wgt = devex ? arena.at<uint32_t>(o_wgt) : nullptr;
    d.node = arena.at<NsxNode>(o_node); d.depth = arena.at<int32_t>(o_depth); d.pi = arena.at<double>(o_pi); d.pi_mirror = nullptr;
    d.order = arena.at<int32_t>(o_order); d.tmp = arena.at<int32_t>(o_tmp);
    d.gpath_h = arena.at<int32_t>(o_gph); d.gpath_t = arena.at<int32_t>(o_gpt);
    d.garc2 = arena.at<int32_t>(o_garc2); d.gres = arena.at<double>(o_gres);
    d.penalty = pb->penalty; d.tol = opt->tolerance;
    ka.ctl = arena.at<NsxCtl>(o_ctl); ka.grid = arena.at<NsxGridCtl>(o_grid);
    ka.dzc = arena.at<NsxCand>(o_dzc); ka.dxc = arena.at<NsxDevexCand>(o_dxc);
    ka.trace = want_trace ? arena.at<int32_t>(o_trace) : nullptr;

    NsxCtl hctl;
    nsx_fill_ctl(hctl, opt, want_trace);
    NSX_CUDA(cudaMemcpyAsync(ka.ctl, &hctl, sizeof hctl, cudaMemcpyHostToDevice, stream));
    NSX_CUDA(cudaMemsetAsync(ka.grid, 0, sizeof(NsxGridCtl), stream));
    NSX_CUDA(cudaEventRecord(ev[1], stream));

    // ---- launch shape ----
    size_t smem = 0;
    ka.plan = nsx_plan_smem(n, info.smem_optin, &smem);
    if (smem > info.smem_optin) { arena.release(); return nsx_fail(NSX_ERR_INTERNAL, "shared memory plan exceeds the device limit"); }
    NSX_CUDA(cudaFuncSetAttribute(nsx_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    NSX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nsx_resident_kernel, NSX_THREADS, smem));
    if (per_sm < 1) { arena.release(); return nsx_fail(NSX_ERR_INTERNAL, "resident kernel does not fit on an SM"); }
    // one pivot CTA + sweep workers; small instances are priced by the pivot CTA alone
    int64_t arcs_per_cta = nsx_env_int("NSX_ARCS_PER_CTA", 8192);
    int64_t workers = (m + arcs_per_cta - 1) / arcs_per_cta;
    if (workers > info.sms - 1) workers = info.sms - 1;
    int grid = m < nsx_env_int("NSX_SINGLE_CTA_ARCS", 65536) || workers < 2 ? 1 : (int)workers + 1;
    int forced = nsx_env_int("NSX_GRID", 0);
    if (forced > 0) grid = forced < info.sms * per_sm ? forced : info.sms * per_sm;
    if (grid > 1024) grid = 1024;

    {
        int ib = (int)((m + n + 1023) / 1024);
        if (ib < 1) ib = 1;
        if (ib > info.sms * 8) ib = info.sms * 8;
        nsx_init_kernel<<<ib, 1024, 0, stream>>>(d, d_supply, ka.ctl);
        NSX_CUDA(cudaGetLastError());
    }
    void* kargs[] = {(void*)&ka};
    NSX_CUDA(cudaLaunchCooperativeKernel((void*)nsx_resident_kernel, dim3(grid), dim3(NSX_THREADS), kargs, smem, stream));
    NSX_CUDA(cudaEventRecord(ev[2], stream));

    // ---- results ----
    if (res->flow) NSX_CUDA(cudaMemcpyAsync(res->flow, d.flow, (size_t)ma * 8, cudaMemcpyDeviceToHost, stream));
    if (res->potential) NSX_CUDA(cudaMemcpyAsync(res->potential, d.pi, (size_t)n * 8, cudaMemcpyDeviceToHost, stream));
    if (res->state) NSX_CUDA(cudaMemcpyAsync(res->state, d.state, (size_t)ma, cudaMemcpyDeviceToHost, stream));
    NSX_CUDA(cudaMemcpyAsync(&hctl, ka.ctl, sizeof hctl, cudaMemcpyDeviceToHost, stream));
    NSX_CUDA(cudaEventRecord(ev[3], stream));
    NSX_CUDA(cudaStreamSynchronize(stream));
    if (want_trace) {
        int64_t cnt = hctl.trace_len < opt->trace_capacity ? hctl.trace_len : opt->trace_capacity;
        if (cnt > 0) NSX_CUDA(cudaMemcpy(res->entering_trace, ka.trace, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
    }
    float ms = 0;
    nsx_harvest(hctl, res);
    cudaEventElapsedTime(&ms, ev[0], ev[1]); res->h2d_ms = ms;
    cudaEventElapsedTime(&ms, ev[1], ev[2]); res->solve_ms = ms;
    cudaEventElapsedTime(&ms, ev[2], ev[3]); res->d2h_ms = ms;
    res->reserved = grid;
    for (auto& e : ev) cudaEventDestroy(e);
    cudaStreamDestroy(stream);
    arena.release();
    if (hctl.status < 0) return nsx_fail(NSX_ERR_INTERNAL, "resident kernel ended without a status");
    return 0;
}

extern "C" int nsx_solve(const nsx_problem* problem, const nsx_options* options, nsx_result* result) {
    return nsx_solve_impl(problem, options, result, false);
}
extern "C" int nsx_solve_resident(const nsx_problem* problem_dev, const nsx_options* options, nsx_result* result) {
    return nsx_solve_impl(problem_dev, options, result, true);
}

extern "C" int nsx_solve_batch(int64_t count, const nsx_problem* problems, const nsx_options* opt, nsx_result* results) {
    Arena arena;
    if (count < 0 || (count > 0 && (!problems || !results)) || !opt) return nsx_fail(NSX_ERR_INVALID_ARGUMENT, "null argument");
    if (count == 0) return 0;
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
    const bool devex = opt->pricing == NSX_PRICING_DEVEX;

    struct Off { size_t tail, head, pert, upper, atail, ahead, aupper, flow, state, wgt, node, depth, pi, order, tmp, gph, gpt, garc2, gres, supply, ctl, trace; };
    std::vector<Off> off(count);
    int32_t max_n = 1;
    for (int64_t i = 0; i < count; ++i) {
        const nsx_problem& p = problems[i];
        const size_t n = p.n_nodes, m = p.n_arcs, ma = m + n - 1;
        if ((int32_t)n > max_n) max_n = (int32_t)n;
        Off& o = off[i];
        o.tail = arena.plan((m + 4) * 4); o.head = arena.plan((m + 4) * 4);
        o.pert = arena.plan((m + 4) * 8); o.upper = arena.plan((m + 4) * 8);
        o.atail = arena.plan(n * 4); o.ahead = arena.plan(n * 4); o.aupper = arena.plan(n * 8);
        o.flow = arena.plan((ma + 4) * 8); o.state = arena.plan(ma + 16);
        o.wgt = devex ? arena.plan((m + 4) * 4) : 0;
        o.node = arena.plan(n * sizeof(NsxNode)); o.depth = arena.plan(n * 4); o.pi = arena.plan(n * 8 + 16);
        o.order = arena.plan(n * 4); o.tmp = arena.plan(n * 4); o.gph = arena.plan(n * 4); o.gpt = arena.plan(n * 4);
        o.garc2 = arena.plan((2 * n + 1) * 4); o.gres = arena.plan((2 * n + 1) * 8);
        o.supply = arena.plan(n * 8); o.ctl = arena.plan(sizeof(NsxCtl));
        o.trace = (results[i].entering_trace && opt->trace_capacity > 0) ? arena.plan((size_t)opt->trace_capacity * 4) : 0;
    }
    size_t o_items = arena.plan(sizeof(NsxBatchItem) * (size_t)count);
    size_t o_next = arena.plan(8);
    NSX_CUDA(arena.commit());

    cudaStream_t stream;
    NSX_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cudaEvent_t ev[4];
    for (auto& e : ev) NSX_CUDA(cudaEventCreate(&e));
    std::vector<NsxBatchItem> items(count);
    std::vector<NsxCtl> ctls(count);
    NSX_CUDA(cudaEventRecord(ev[0], stream));
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
        d.penalty = p.penalty; d.tol = opt->tolerance;
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

    size_t smem = 0;
    (void)nsx_plan_smem(max_n, info.smem_optin, &smem);
    if (smem > info.smem_optin) { arena.release(); return nsx_fail(NSX_ERR_INTERNAL, "shared memory plan exceeds the device limit"); }
    NSX_CUDA(cudaFuncSetAttribute(nsx_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    NSX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nsx_batch_kernel, NSX_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)info.sms * per_sm;
    if (grid > count) grid = count;
    nsx_batch_kernel<<<(int)grid, NSX_THREADS, smem, stream>>>(d_items, count, d_next, smem - nsx_smem_fixed());
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
    int bad = 0;
    for (int64_t i = 0; i < count; ++i) {
        nsx_result& r = results[i];
        nsx_harvest(ctls[i], &r);
        r.h2d_ms = h2d; r.solve_ms = solve; r.d2h_ms = d2h; r.reserved = (int32_t)grid;
        if (items[i].trace) {
            int64_t cnt = ctls[i].trace_len < opt->trace_capacity ? ctls[i].trace_len : opt->trace_capacity;
            if (cnt > 0) NSX_CUDA(cudaMemcpy(r.entering_trace, items[i].trace, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
        }
        if (ctls[i].status < 0) bad++;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    cudaStreamDestroy(stream);
    arena.release();
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
