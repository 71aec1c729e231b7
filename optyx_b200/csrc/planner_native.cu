// Host-side (CPU, no kernels) parts of the contraction planner that are too slow in
// Python for networks of thousands of tensors: the randomised greedy pair-order search
// and the cost analysis of a pair order.  The reference delegates the path search to
// cotengra (a HyperOptimizer handed to QuimbBackend, optyx/core/backends.py:436-455),
// whose own greedy/partitioning code is compiled as well; this is the planner of
// optyx_b200/planner.py (hyper-index aware: an index may sit on any number of tensors
// and on the output), moved below the C ABI.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <queue>
#include <unordered_set>
#include <vector>
#include "b2_common.cuh"

namespace {

struct Rng {                                 // xorshift64*: reproducible across platforms
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull) {
        if (!s) s = 1;
    }
    double uniform() {
        s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
        return (double)((s * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0);
    }
};

struct Cand {
    double score;
    int64_t tick;
    int a, b;
    bool operator<(const Cand& o) const {    // min-heap on (score, tick)
        if (score != o.score) return score > o.score;
        return tick > o.tick;
    }
};

}  // namespace

// Greedy SSA pair order: repeatedly contract the pair minimising
//   sign * log2(1 + |size(out) - costmod * (size(a) + size(b))|)  (+ Gumbel noise)
// Tensors 0..n-1 are the inputs (index lists in CSR form, ids 0..n_inds-1); step k
// creates tensor n + k.  path_out receives 2 * (n - 1) ids.
extern "C" int b2_plan_greedy(int32_t n, const int32_t* ind_ptr, const int32_t* inds,
                              int32_t n_inds, const int32_t* dims, const uint8_t* is_output,
                              uint64_t seed, double temperature, double costmod,
                              int32_t* path_out) {
    if (n < 1 || !ind_ptr || !inds || !dims || !path_out) {
        b2::set_error("b2_plan_greedy: bad arguments");
        return 1;
    }
    const int total = 2 * n - 1;
    std::vector<std::vector<int>> tens(total);
    std::vector<char> alive(total, 0);
    std::vector<double> size(total, 1.0);
    std::vector<int> count(n_inds, 0);
    std::vector<std::vector<int>> where(n_inds);        // lazy: may hold dead tensors
    for (int k = 0; k < n; ++k) {
        std::vector<int>& t = tens[k];
        t.assign(inds + ind_ptr[k], inds + ind_ptr[k + 1]);
        std::sort(t.begin(), t.end());
        t.erase(std::unique(t.begin(), t.end()), t.end());
        alive[k] = 1;
        double s = 1.0;
        for (int q : t) { ++count[q]; where[q].push_back(k); s *= dims[q]; }
        size[k] = s;
    }
    for (int q = 0; q < n_inds; ++q) if (is_output && is_output[q]) ++count[q];
    Rng rng(seed);
    std::priority_queue<Cand> heap;
    int64_t tick = 0;
    std::vector<int> out_buf;
    auto out_inds = [&](int a, int b, std::vector<int>& res) {
        res.clear();
        const std::vector<int>&A = tens[a], &B = tens[b];
        size_t i = 0, j = 0;
        while (i < A.size() || j < B.size()) {
            int q, here;
            if (j >= B.size() || (i < A.size() && A[i] < B[j])) { q = A[i++]; here = 1; }
            else if (i >= A.size() || B[j] < A[i]) { q = B[j++]; here = 1; }
            else { q = A[i]; ++i; ++j; here = 2; }
            if (count[q] > here) res.push_back(q);
        }
    };
    auto push = [&](int a, int b) {
        out_inds(a, b, out_buf);
        double so = 1.0;
        for (int q : out_buf) so *= dims[q];
        double sc = so - costmod * (size[a] + size[b]);
        sc = std::copysign(std::log2(1.0 + std::fabs(sc)), sc);
        if (temperature > 0.0) {
            double u = rng.uniform();
            if (u < 1e-300) u = 1e-300;
            sc -= temperature * -std::log(-std::log(u));
        }
        heap.push(Cand{sc, tick++, a, b});
    };
    std::vector<int> nb, tmp;
    auto neighbours = [&](int a, std::vector<int>& res) {
        res.clear();
        for (int q : tens[a]) {
            std::vector<int>& w = where[q];
            // compact the lazy list
            w.erase(std::remove_if(w.begin(), w.end(), [&](int k) { return !alive[k]; }), w.end());
            if (w.size() > 24) {               // huge hyper-edge: a few smallest partners do
                tmp = w;
                std::partial_sort(tmp.begin(), tmp.begin() + 24, tmp.end(),
                                  [&](int x, int y) { return size[x] < size[y] || (size[x] == size[y] && x < y); });
                res.insert(res.end(), tmp.begin(), tmp.begin() + 24);
            } else res.insert(res.end(), w.begin(), w.end());
        }
        std::sort(res.begin(), res.end());
        res.erase(std::unique(res.begin(), res.end()), res.end());
        res.erase(std::remove(res.begin(), res.end(), a), res.end());
    };
    for (int a = 0; a < n; ++a) {
        neighbours(a, nb);
        for (int b : nb) if (a < b) push(a, b);
    }
    int live = n, nxt = n, written = 0;
    std::vector<int> merged;
    while (live > 1) {
        int a = -1, b = -1;
        while (!heap.empty()) {
            const Cand c = heap.top();
            heap.pop();
            if (alive[c.a] && alive[c.b]) { a = c.a; b = c.b; break; }
        }
        if (a < 0) {                           // disconnected components: outer products
            int best[2] = {-1, -1};
            for (int k = 0; k < nxt; ++k) {
                if (!alive[k]) continue;
                if (best[0] < 0 || size[k] < size[best[0]]) { best[1] = best[0]; best[0] = k; }
                else if (best[1] < 0 || size[k] < size[best[1]]) best[1] = k;
            }
            a = best[0]; b = best[1];
        }
        out_inds(a, b, merged);
        for (int k : {a, b}) {
            for (int q : tens[k]) --count[q];
            alive[k] = 0;
        }
        tens[nxt] = merged;
        alive[nxt] = 1;
        double s = 1.0;
        for (int q : merged) { ++count[q]; where[q].push_back(nxt); s *= dims[q]; }
        size[nxt] = s;
        path_out[written++] = a;
        path_out[written++] = b;
        --live;
        neighbours(nxt, nb);
        for (int c : nb) push(nxt, c);
        ++nxt;
    }
    return 0;
}

// Cost of a pair order with some indices sliced (fixed): per-slice multiply-adds, the
// largest intermediate, and the roofline time estimate t_inv + nslices * t_dep of
// optyx_b200/planner.py::analyse_path (a step costs max(bytes / HBM, flops / FP64 pipe)
// plus a launch floor; steps that touch no sliced index run once).
// out: [macs, peak, nslices, time_est]; step_sizes_out (optional, n - 1): log2 of each
// step's result size after slicing.
extern "C" int b2_plan_cost(int32_t n, const int32_t* ind_ptr, const int32_t* inds,
                            int32_t n_inds, const int32_t* dims, const uint8_t* is_output,
                            const int32_t* path, const uint8_t* is_sliced, double* out,
                            double* step_sizes_out) {
    if (n < 1 || !ind_ptr || !inds || !dims || !path || !out) {
        b2::set_error("b2_plan_cost: bad arguments");
        return 1;
    }
    const double BW = 6.5e12, PIPE = 37e12, ES = 16.0;
    const int total = 2 * n - 1;
    std::vector<std::vector<int>> tens(total);
    std::vector<char> dep(total, 0);
    std::vector<int> count(n_inds, 0);
    double nslices = 1.0;
    for (int q = 0; q < n_inds; ++q) {
        if (is_output && is_output[q]) ++count[q];
        if (is_sliced && is_sliced[q]) nslices *= dims[q];
    }
    auto fsize = [&](const std::vector<int>& t) {
        double s = 1.0;
        for (int q : t) if (!(is_sliced && is_sliced[q])) s *= dims[q];
        return s;
    };
    for (int k = 0; k < n; ++k) {
        std::vector<int>& t = tens[k];
        t.assign(inds + ind_ptr[k], inds + ind_ptr[k + 1]);
        std::sort(t.begin(), t.end());
        t.erase(std::unique(t.begin(), t.end()), t.end());
        for (int q : t) { ++count[q]; if (is_sliced && is_sliced[q]) dep[k] = 1; }
    }
    double macs = 0.0, peak = 0.0, t_inv = 0.0, t_dep = 0.0;
    int nxt = n;
    std::vector<int> uni, res;
    for (int s = 0; s < n - 1; ++s) {
        const int a = path[2 * s], b = path[2 * s + 1];
        if (a < 0 || b < 0 || a >= nxt || b >= nxt || a == b) {
            b2::set_error("b2_plan_cost: bad path entry %d", s);
            return 1;
        }
        const std::vector<int>&A = tens[a], &B = tens[b];
        uni.clear(); res.clear();
        size_t i = 0, j = 0;
        while (i < A.size() || j < B.size()) {
            int q, here;
            if (j >= B.size() || (i < A.size() && A[i] < B[j])) { q = A[i++]; here = 1; }
            else if (i >= A.size() || B[j] < A[i]) { q = B[j++]; here = 1; }
            else { q = A[i]; ++i; ++j; here = 2; }
            uni.push_back(q);
            if (count[q] > here) res.push_back(q);
        }
        for (int q : A) --count[q];
        for (int q : B) --count[q];
        for (int q : res) ++count[q];
        const double m = fsize(uni), sc = fsize(res);
        const double nbytes = ES * (fsize(A) + fsize(B) + sc);
        const double t = std::max(nbytes / BW, 8.0 * m / PIPE) + 2e-6;
        const char d = dep[a] | dep[b];
        dep[nxt] = d;
        if (d) t_dep += t; else t_inv += t;
        macs += m;
        if (sc > peak) peak = sc;
        if (step_sizes_out) step_sizes_out[s] = std::log2(sc);
        tens[nxt] = res;
        std::vector<int>().swap(tens[a]);
        std::vector<int>().swap(tens[b]);
        ++nxt;
    }
    out[0] = macs; out[1] = peak; out[2] = nslices; out[3] = t_inv + nslices * t_dep;
    return 0;
}

// ---------------------------------------------------------------------------------------
// Subtree reconfiguration: walk the contraction tree, and for every subtree of at most
// `max_leaves` frontier nodes replace its internal pair order by the OPTIMAL one (dynamic
// programming over subsets, cost = multiply-adds with the sliced indices fixed).  The
// classic local search of tensor-network path optimisers (cotengra's
// `subtree_reconfigure`), here hyper-index aware: an index survives a partial result as
// long as some tensor outside it -- inside or outside the subtree -- or the output still
// carries it.
// ---------------------------------------------------------------------------------------
namespace {

struct Node {
    int left = -1, right = -1;          // children (-1: an input tensor)
    std::vector<int> inds;              // sorted result indices
};

static void merge_inds(const std::vector<int>& A, const std::vector<int>& B,
                       const std::vector<int>& outside, std::vector<int>& res,
                       std::vector<int>* uni) {
    // outside[q]: how many tensors OUTSIDE A and B (incl. the output) carry q
    res.clear();
    if (uni) uni->clear();
    size_t i = 0, j = 0;
    while (i < A.size() || j < B.size()) {
        int q;
        if (j >= B.size() || (i < A.size() && A[i] < B[j])) q = A[i++];
        else if (i >= A.size() || B[j] < A[i]) q = B[j++];
        else { q = A[i]; ++i; ++j; }
        if (uni) uni->push_back(q);
        if (outside[q] > 0) res.push_back(q);
    }
}

}  // namespace

// path: SSA pair order (2 * (n - 1) ids), rewritten in place.  Returns the number of
// subtrees that were improved through *improved_out.
extern "C" int b2_plan_reconfigure(int32_t n, const int32_t* ind_ptr, const int32_t* inds,
                                   int32_t n_inds, const int32_t* dims, const uint8_t* is_output,
                                   const uint8_t* is_sliced, int32_t max_leaves, int32_t* path,
                                   int32_t* improved_out) {
    if (n < 2 || !ind_ptr || !inds || !dims || !path) { b2::set_error("b2_plan_reconfigure: bad arguments"); return 1; }
    if (max_leaves < 3) max_leaves = 3;
    if (max_leaves > 10) max_leaves = 10;
    const int total = 2 * n - 1;
    std::vector<double> ldim(n_inds);
    for (int q = 0; q < n_inds; ++q) ldim[q] = (is_sliced && is_sliced[q]) ? 0.0 : std::log2((double)dims[q]);
    // ---- build the tree from the SSA path
    std::vector<Node> node(total);
    std::vector<int> carriers(n_inds, 0);            // live tensors + output carrying q
    for (int k = 0; k < n; ++k) {
        std::vector<int>& t = node[k].inds;
        t.assign(inds + ind_ptr[k], inds + ind_ptr[k + 1]);
        std::sort(t.begin(), t.end());
        t.erase(std::unique(t.begin(), t.end()), t.end());
        for (int q : t) ++carriers[q];
    }
    for (int q = 0; q < n_inds; ++q) if (is_output && is_output[q]) ++carriers[q];
    {
        std::vector<int> cnt = carriers, res;
        for (int s = 0; s < n - 1; ++s) {
            const int a = path[2 * s], b = path[2 * s + 1], c = n + s;
            if (a < 0 || b < 0 || a >= c || b >= c || a == b) { b2::set_error("b2_plan_reconfigure: bad path"); return 1; }
            for (int q : node[a].inds) --cnt[q];
            for (int q : node[b].inds) --cnt[q];
            merge_inds(node[a].inds, node[b].inds, cnt, res, nullptr);
            for (int q : res) ++cnt[q];
            node[c].left = a; node[c].right = b; node[c].inds = res;
        }
    }
    const int root = total - 1;
    int improved = 0;
    // ---- local search: every internal node is the root of one candidate subtree
    std::vector<int> order;                          // internal nodes, root first (BFS)
    {
        std::vector<int> q{root};
        for (size_t h = 0; h < q.size(); ++h) {
            const int v = q[h];
            if (node[v].left < 0) continue;
            order.push_back(v);
            q.push_back(node[v].left); q.push_back(node[v].right);
        }
    }
    std::vector<int> frontier, internal, local_of(n_inds, -1), touched;
    for (int v : order) {
        if (node[v].left < 0) continue;
        // frontier: expand the largest expandable node until max_leaves
        frontier.assign({node[v].left, node[v].right});
        internal.assign({v});
        while ((int)frontier.size() < max_leaves) {
            int pick = -1; double best = -1.0;
            for (size_t i = 0; i < frontier.size(); ++i) {
                const int f = frontier[i];
                if (node[f].left < 0) continue;
                double s = 0.0;
                for (int q : node[f].inds) s += ldim[q];
                if (s > best) { best = s; pick = (int)i; }
            }
            if (pick < 0) break;
            const int f = frontier[pick];
            internal.push_back(f);
            frontier[pick] = node[f].left;
            frontier.push_back(node[f].right);
        }
        const int L = (int)frontier.size();
        if (L < 3) continue;
        // local index table and per-leaf bitsets (up to 64 * W local indices)
        touched.clear();
        for (int f : frontier) for (int q : node[f].inds) if (local_of[q] < 0) { local_of[q] = (int)touched.size(); touched.push_back(q); }
        const int nl = (int)touched.size(), W = (nl + 63) / 64;
        std::vector<uint64_t> leafbits((size_t)L * W, 0), extbits(W, 0);
        std::vector<int> inside(nl, 0);
        for (int i = 0; i < L; ++i) for (int q : node[frontier[i]].inds) {
            const int l = local_of[q];
            leafbits[(size_t)i * W + l / 64] |= 1ull << (l % 64);
            ++inside[l];
        }
        // external: carried by a tensor outside the subtree or by the output
        //   = appears in the subtree root's result
        for (int q : node[v].inds) { const int l = local_of[q]; if (l >= 0) extbits[l / 64] |= 1ull << (l % 64); }
        std::vector<double> lw(nl);
        for (int l = 0; l < nl; ++l) lw[l] = ldim[touched[l]];
        const int NS = 1 << L;
        std::vector<uint64_t> ib((size_t)NS * W, 0);          // inds of a subset (union of leaves)
        for (int S = 1; S < NS; ++S) {
            const int low = __builtin_ctz(S), rest = S & (S - 1);
            for (int w = 0; w < W; ++w) ib[(size_t)S * W + w] = ib[(size_t)rest * W + w] | leafbits[(size_t)low * W + w];
        }
        // kept(S) = inds(S) & (ext | inds(~S))
        auto kept_size = [&](int S1, int S2, std::vector<uint64_t>& tmp) {   // log2 MACs of S1 x S2
            const int comp1 = (NS - 1) & ~S1, comp2 = (NS - 1) & ~S2;
            double s = 0.0;
            for (int w = 0; w < W; ++w) {
                const uint64_t k1 = ib[(size_t)S1 * W + w] & (extbits[w] | ib[(size_t)comp1 * W + w]);
                const uint64_t k2 = ib[(size_t)S2 * W + w] & (extbits[w] | ib[(size_t)comp2 * W + w]);
                uint64_t u = k1 | k2;
                while (u) { const int b = __builtin_ctzll(u); u &= u - 1; s += lw[w * 64 + b]; }
            }
            (void)tmp;
            return s;
        };
        std::vector<double> bestc(NS, 1e300);
        std::vector<int> split(NS, 0);
        std::vector<uint64_t> tmp(W);
        for (int i = 0; i < L; ++i) bestc[1 << i] = 0.0;
        for (int S = 1; S < NS; ++S) {
            if ((S & (S - 1)) == 0) continue;
            // enumerate splits S1 < S2, S1 | S2 = S
            for (int S1 = (S - 1) & S; S1 > 0; S1 = (S1 - 1) & S) {
                const int S2 = S ^ S1;
                if (S1 < S2) continue;
                if (bestc[S1] >= 1e299 || bestc[S2] >= 1e299) continue;
                const double c = std::exp2(kept_size(S1, S2, tmp));
                const double tot = bestc[S1] + bestc[S2] + c;
                if (tot < bestc[S]) { bestc[S] = tot; split[S] = S1; }
            }
        }
        // current cost of the subtree's internal order
        double cur = 0.0;
        {
            std::vector<int> uni, res;
            for (int u : internal) {
                const std::vector<int>&A = node[node[u].left].inds, &B = node[node[u].right].inds;
                double s = 0.0;
                size_t i = 0, j = 0;
                while (i < A.size() || j < B.size()) {
                    int q;
                    if (j >= B.size() || (i < A.size() && A[i] < B[j])) q = A[i++];
                    else if (i >= A.size() || B[j] < A[i]) q = B[j++];
                    else { q = A[i]; ++i; ++j; }
                    s += ldim[q];
                }
                cur += std::exp2(s);
            }
        }
        if (bestc[NS - 1] < cur * (1.0 - 1e-9)) {
            // rebuild: reuse the internal node slots (v stays the subtree root)
            ++improved;
            std::vector<int> slots(internal.begin() + 1, internal.end());
            std::vector<int> made(NS, -1);
            for (int i = 0; i < L; ++i) made[1 << i] = frontier[i];
            // post-order construction
            std::vector<int> stack{NS - 1};
            std::vector<int> post;
            while (!stack.empty()) {
                const int S = stack.back(); stack.pop_back();
                if ((S & (S - 1)) == 0) continue;
                post.push_back(S);
                stack.push_back(split[S]); stack.push_back(S ^ split[S]);
            }
            for (auto it = post.rbegin(); it != post.rend(); ++it) {
                const int S = *it, S1 = split[S], S2 = S ^ S1;
                int slot;
                if (S == NS - 1) slot = v;
                else { slot = slots.back(); slots.pop_back(); }
                node[slot].left = made[S1]; node[slot].right = made[S2];
                // result indices: kept(S)
                std::vector<int>& r = node[slot].inds;
                if (S != NS - 1) {
                    r.clear();
                    const int comp = (NS - 1) & ~S;
                    for (int w = 0; w < W; ++w) {
                        uint64_t k = ib[(size_t)S * W + w] & (extbits[w] | ib[(size_t)comp * W + w]);
                        while (k) { const int b = __builtin_ctzll(k); k &= k - 1; r.push_back(touched[w * 64 + b]); }
                    }
                    std::sort(r.begin(), r.end());
                }
                made[S] = slot;
            }
        }
        for (int q : touched) local_of[q] = -1;
    }
    // ---- back to an SSA path (post-order; children before parents)
    {
        std::vector<int> ssa_id(total, -1);
        for (int k = 0; k < n; ++k) ssa_id[k] = k;
        int nxt = n, written = 0;
        std::vector<std::pair<int, int>> st;          // (node, state)
        st.push_back({root, 0});
        while (!st.empty()) {
            auto [u, state] = st.back();
            st.pop_back();
            if (node[u].left < 0) continue;
            if (state == 0) {
                st.push_back({u, 1});
                st.push_back({node[u].right, 0});
                st.push_back({node[u].left, 0});
            } else {
                path[written++] = ssa_id[node[u].left];
                path[written++] = ssa_id[node[u].right];
                ssa_id[u] = nxt++;
            }
        }
        if (written != 2 * (n - 1)) { b2::set_error("b2_plan_reconfigure: tree corrupted"); return 1; }
    }
    if (improved_out) *improved_out = improved;
    return 0;
}
