// K9 core -- complete-linkage clustering of a handful of points, cut at a distance threshold, with
// exactly the flat-cluster LABELS that scipy produces, because the reference emits clusters in
// label order (SVIM_COMBINE.py:134-138,155-159; SVIM_inter.py:47-51).
//
// The reference calls scipy.cluster.hierarchy.linkage(method="complete") and
// fcluster(Z, t, criterion="distance") [ext: scipy is a third-party dependency of the reference,
// setup.py:38].  scipy's published algorithm is: nearest-neighbour chain over the condensed
// distance vector; rows stably sorted by height; relabelling through a union-find whose merged
// roots are numbered n, n+1, ...; flat clusters by a left-first depth-first walk from the root,
// opening a cluster at the first node whose maximum height is <= t.
// Pinned against the installed scipy in tests/test_linkage.py.
//
// One thread runs one problem; n <= LINK_MAXN.  The same source compiles for the host
// (tests/hostcheck) so the logic is exercised without a GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SVB_HD __host__ __device__ __forceinline__
#else
#define SVB_HD inline
#endif

constexpr int LINK_MAXN = 32;

SVB_HD int link_cidx(int n, int i, int j) {      // condensed index, any order of i != j
    if (i > j) { const int t = i; i = j; j = t; }
    return n * i - i * (i + 1) / 2 + (j - i - 1);
}

// The algorithm only COMPARES distances (nearest neighbour: `<`; complete linkage: max; stable sort by height: `>`; the
// cut: `<= t`), so it is written once over an `Ops` policy: plain doubles for the clustering itself, and intervals
// "all that is known about a distance" for the question pair.cu asks before it decides which far pairs need the exact
// edit-distance kernel: did any decision depend on where inside their intervals the unknown values lie?
struct LinkDoubleOps {
    typedef double T;
    SVB_HD bool lt(double a, double b) { return a < b; }
    SVB_HD bool le_threshold(double a, double t) { return a <= t; }
    SVB_HD double infinity() { return 1.0 / 0.0; }
};

// lo == hi: the value is known.  lo < hi: it lies somewhere in [lo, hi]; `id` names the quantity (two intervals with the
// same id are the same unknown number).  A comparison whose outcome is not the same for every admissible pair of values sets
// `ambiguous`; the result of the run is then meaningless and the caller computes the values exactly.
struct LinkInterval {
    double lo, hi;
    uint32_t id;
};
struct LinkIntervalOps {
    typedef LinkInterval T;
    bool ambiguous;
    SVB_HD bool lt(const LinkInterval& a, const LinkInterval& b) {
        if (a.lo < a.hi && b.lo < b.hi && a.id == b.id) return false;      // the same unknown: equal
        if (a.hi < b.lo) return true;
        if (a.lo >= b.hi) return false;
        ambiguous = true;
        return a.lo < b.lo;
    }
    SVB_HD bool le_threshold(const LinkInterval& a, double t) {
        if (a.hi <= t) return true;
        if (a.lo > t) return false;
        ambiguous = true;
        return false;
    }
    SVB_HD LinkInterval infinity() {
        LinkInterval v;
        v.lo = v.hi = 1.0 / 0.0;
        v.id = 0xFFFFFFFFu;
        return v;
    }
};

// D: condensed distances, n*(n-1)/2 entries, overwritten.  labels: n entries, 1-based.
// Returns the number of flat clusters.
template <class Ops>
SVB_HD int link_complete_fcluster_ops(int n, typename Ops::T* D, double threshold, int* labels, Ops& ops) {
    typedef typename Ops::T T;
    if (n <= 0) return 0;
    if (n == 1) { labels[0] = 1; return 1; }
    int size[LINK_MAXN];
    int chain[LINK_MAXN];
    int zx[LINK_MAXN], zy[LINK_MAXN];
    T zd[LINK_MAXN];
    for (int i = 0; i < n; ++i) size[i] = 1;
    int chain_len = 0;
    for (int k = 0; k < n - 1; ++k) {
        if (chain_len == 0) {
            for (int i = 0; i < n; ++i)
                if (size[i] > 0) { chain[0] = i; break; }
            chain_len = 1;
        }
        int x, y = -1;
        T cur;
        while (true) {
            x = chain[chain_len - 1];
            if (chain_len > 1) {
                y = chain[chain_len - 2];
                cur = D[link_cidx(n, x, y)];
            } else {
                cur = ops.infinity();
            }
            for (int i = 0; i < n; ++i) {
                if (size[i] == 0 || i == x) continue;
                const T d = D[link_cidx(n, x, i)];
                if (ops.lt(d, cur)) { cur = d; y = i; }
            }
            if (chain_len > 1 && y == chain[chain_len - 2]) break;
            chain[chain_len++] = y;
        }
        chain_len -= 2;
        if (x > y) { const int t = x; x = y; y = t; }
        zx[k] = x; zy[k] = y; zd[k] = cur;
        const int nx = size[x], ny = size[y];
        size[x] = 0;
        size[y] = nx + ny;
        for (int i = 0; i < n; ++i) {
            if (size[i] == 0 || i == y) continue;
            const T a = D[link_cidx(n, i, x)], b = D[link_cidx(n, i, y)];
            D[link_cidx(n, i, y)] = ops.lt(b, a) ? a : b;                // complete linkage
        }
    }
    // stable sort of the n-1 merges by height (insertion sort)
    int order[LINK_MAXN];
    for (int k = 0; k < n - 1; ++k) order[k] = k;
    for (int k = 1; k < n - 1; ++k) {
        const int o = order[k];
        int m = k - 1;
        while (m >= 0 && ops.lt(zd[o], zd[order[m]])) { order[m + 1] = order[m]; --m; }
        order[m + 1] = o;
    }
    // relabel through a union-find; merged root of sorted row k is node n + k
    int parent[2 * LINK_MAXN];
    for (int i = 0; i < 2 * n - 1; ++i) parent[i] = i;
    int left[LINK_MAXN], right[LINK_MAXN];
    T md[LINK_MAXN];
    for (int k = 0; k < n - 1; ++k) {
        int rx = zx[order[k]], ry = zy[order[k]];
        while (parent[rx] != rx) rx = parent[rx];
        while (parent[ry] != ry) ry = parent[ry];
        left[k] = rx < ry ? rx : ry;
        right[k] = rx < ry ? ry : rx;
        parent[rx] = n + k;
        parent[ry] = n + k;
        T m = zd[order[k]];                                        // max height inside the subtree
        if (left[k] >= n && ops.lt(m, md[left[k] - n])) m = md[left[k] - n];
        if (right[k] >= n && ops.lt(m, md[right[k] - n])) m = md[right[k] - n];
        md[k] = m;
    }
    // flat clusters: left-first DFS from the root, a cluster opens at the first node with md <= t
    int stack[LINK_MAXN];
    bool visited[LINK_MAXN];
    for (int k = 0; k < n - 1; ++k) visited[k] = false;
    int n_cluster = 0, leader = -1, sp = 0;
    stack[0] = n - 2;
    while (sp >= 0) {
        const int root = stack[sp];
        const int lc = left[root], rc = right[root];
        if (leader == -1 && ops.le_threshold(md[root], threshold)) { leader = root; ++n_cluster; }
        if (lc >= n && !visited[lc - n]) { visited[lc - n] = true; stack[++sp] = lc - n; continue; }
        if (rc >= n && !visited[rc - n]) { visited[rc - n] = true; stack[++sp] = rc - n; continue; }
        if (lc < n) { if (leader == -1) ++n_cluster; labels[lc] = n_cluster; }
        if (rc < n) { if (leader == -1) ++n_cluster; labels[rc] = n_cluster; }
        if (leader == root) leader = -1;
        --sp;
    }
    return n_cluster;
}

SVB_HD int link_complete_fcluster(int n, double* D, double threshold, int* labels) {
    LinkDoubleOps ops;
    return link_complete_fcluster_ops(n, D, threshold, labels, ops);
}

// True when the flat clustering of `n` points is the same for EVERY choice of values inside the given intervals (then
// `labels` holds it).  D is overwritten.
SVB_HD bool link_labels_determined(int n, LinkInterval* D, double threshold, int* labels) {
    LinkIntervalOps ops;
    ops.ambiguous = false;
    link_complete_fcluster_ops(n, D, threshold, labels, ops);
    return !ops.ambiguous;
}
