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
