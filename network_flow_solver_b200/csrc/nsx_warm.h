// nsx_warm.h - host-side layout of a caller-supplied initial spanning tree (nsx_solve_warm).
//
// The reference rebuilds its parent pointers from the tree-arc flags by a walk from the root
// (TreeBasis.rebuild, basis.py:82-125) after _apply_warm_start_basis (simplex.py:740-911).  The engine keeps the
// tree as a preorder array plus {parent, pred arc, position, subtree size} records and a depth array
// (nsx_core.cuh), so the same walk is done here, depth first, and the arrays are copied to the device before the
// resident kernel starts.  The order of siblings is free: every preorder of the same tree gives the same pivots.
// Plain C++ (no CUDA): shared by the engine and by the host emulation of the device core (tests/emu).
#ifndef NSX_WARM_H
#define NSX_WARM_H

#include <math.h>
#include <stdint.h>

#include <vector>

// Returns 0, -1 when in_tree does not mark exactly n - 1 arcs, -2 when they do not reach every node.
// NsxNode comes from nsx_core.cuh (included first by both users).
static inline int nsx_warm_layout(int32_t n, int64_t m, const int32_t* tail, const int32_t* head, const double* supply,
                                  double tol, const uint8_t* in_tree, std::vector<NsxNode>& node,
                                  std::vector<int32_t>& depth, std::vector<int32_t>& order) {
    const int64_t ma = m + n - 1;
    auto a_tail = [&](int64_t a) -> int32_t {
        if (a < m) return tail[a];
        const int32_t v = (int32_t)(a - m) + 1;  // artificial arc: v -> root for a supply node, root -> v otherwise
        return (fabs(supply[v]) > tol && supply[v] > 0) ? v : 0;
    };
    auto a_head = [&](int64_t a) -> int32_t {
        if (a < m) return head[a];
        const int32_t v = (int32_t)(a - m) + 1;
        return (fabs(supply[v]) > tol && supply[v] > 0) ? 0 : v;
    };
    int64_t marked = 0;
    std::vector<int64_t> start((size_t)n + 1, 0);
    for (int64_t a = 0; a < ma; ++a)
        if (in_tree[a]) { ++marked; ++start[(size_t)a_tail(a) + 1]; ++start[(size_t)a_head(a) + 1]; }
    if (marked != (int64_t)n - 1) return -1;
    for (int32_t v = 0; v < n; ++v) start[(size_t)v + 1] += start[v];
    std::vector<int64_t> adj((size_t)(2 * marked > 0 ? 2 * marked : 1)), fill(start.begin(), start.end() - 1);
    for (int64_t a = 0; a < ma; ++a)
        if (in_tree[a]) { adj[(size_t)fill[a_tail(a)]++] = a; adj[(size_t)fill[a_head(a)]++] = a; }

    node.assign((size_t)n, NsxNode{-1, -1, 0, 1});
    depth.assign((size_t)n, 0);
    order.assign((size_t)n, 0);
    std::vector<int64_t> next(start.begin(), start.end() - 1);  // per node: next adjacency entry to look at
    std::vector<int32_t> stack;
    stack.reserve(64);
    node[0].parent = 0; node[0].pred2 = -1; node[0].pos = 0;
    int32_t count = 1;
    stack.push_back(0);
    while (!stack.empty()) {
        const int32_t u = stack.back();
        if (next[u] == start[(size_t)u + 1]) { stack.pop_back(); continue; }
        const int64_t a = adj[(size_t)next[u]++];
        const int32_t t = a_tail(a), h = a_head(a);
        const int32_t v = t == u ? h : t;
        if (v == u || node[v].parent >= 0) continue;  // the arc to u's own parent (or a self-loop)
        node[v].parent = u;
        node[v].pred2 = (int32_t)(a * 2 + (t == v ? 1 : 0));  // bit 0: the arc points child -> parent
        node[v].pos = count;
        depth[v] = depth[u] + 1;
        order[(size_t)count++] = v;
        stack.push_back(v);
    }
    if (count != n) return -2;
    for (int32_t k = n - 1; k >= 1; --k) {  // children come after their parent in a preorder
        const int32_t v = order[(size_t)k];
        node[node[v].parent].size += node[v].size;
    }
    return 0;
}

#endif  // NSX_WARM_H
