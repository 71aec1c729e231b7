// Piece (1) of the hot path: build every dense truncated-Fock / ZX leaf array of
// an eval (or a batch of evals) directly in HBM with ONE launch.
//
// Reference: Box.truncation (optyx/core/diagram.py:536-590) loops in Python over
// every input basis state and calls a generator per state; here each thread
// owns one arena element, decodes its multi-index and evaluates the closed-form
// predicate of the box.  Write-only, HBM-write-bound: 16 B (complex128) per
// element, no reads except the tiny descriptor table (L1/L2 resident) and the
// host-computed params of VEC/DENSE leaves.
#include <stdlib.h>
#include "b2_common.cuh"

namespace b2 {

// exact binomial in 64-bit integers (n stays < 64 for any truncation the
// reference can produce: wire dims are capped by photon count + 1)
__device__ __forceinline__ unsigned long long binom(int n, int k) {
    if (k < 0 || k > n) return 0ull;
    if (k > n - k) k = n - k;
    unsigned long long c = 1ull;
    for (int j = 1; j <= k; ++j) c = c * (unsigned long long)(n - k + j) / (unsigned long long)j;
    return c;
}

__device__ __forceinline__ double leaf_value(const b2_leaf_desc& d, const int* idx,
                                             bool& from_params) {
    from_params = false;
    switch (d.kind) {
    case B2_K_ONEHOT: return idx[0] == d.ipar ? 1.0 : 0.0;
    case B2_K_ONES: return 1.0;
    case B2_K_VEC:
    case B2_K_DENSE: from_params = true; return 0.0;
    case B2_K_SUM: {   // W: sqrt(n! / prod n_i!) iff sum n_i == n   (zw.py:231-244)
        int s = 0;
        for (int a = 1; a < d.ndim; ++a) s += idx[a];
        if (s != idx[0]) return 0.0;
        double m = 1.0;
        int rem = s;
        for (int a = 1; a < d.ndim; ++a) { m *= (double)binom(rem, idx[a]); rem -= idx[a]; }
        return sqrt(m);
    }
    case B2_K_ADD: {   // zw.py:676-683
        int s = 0;
        for (int a = 1; a < d.ndim; ++a) s += idx[a];
        return s == idx[0] ? 1.0 : 0.0;
    }
    case B2_K_MUL: {   // zw.py:720-727
        long long p = 1;
        for (int a = 1; a < d.ndim; ++a) p *= idx[a];
        return p == idx[0] ? 1.0 : 0.0;
    }
    case B2_K_DIV:     // zw.py:763-772   [q, a, b]
        return (idx[2] != 0 && idx[1] == idx[0] * idx[2]) ? 1.0 : 0.0;
    case B2_K_MOD2:    // zw.py:818-825   [b, n]
        return (idx[1] % 2 == idx[0]) ? 1.0 : 0.0;
    case B2_K_DUALRAIL:  // diagram.py:904-913   [b, m0, m1]
        return (idx[1] == 1 - idx[0] && idx[2] == idx[0]) ? 1.0 : 0.0;
    case B2_K_PTD:     // diagram.py:975-983   [b, n]
        return (idx[0] == (idx[1] < 1 ? idx[1] : 1)) ? 1.0 : 0.0;
    case B2_K_EMBED:   // diagram.py:1005-1027
        return idx[0] == idx[1] ? 1.0 : 0.0;
    default: return 0.0;
    }
}

// One thread owns one arena ELEMENT and writes it for a whole group of evals: the leaf
// lookup (a bounded binary search between the first and last leaf of the thread's
// 256-element block, `block_leaf`, tabulated at plan time), the 32-bit multi-index decode
// and the closed-form value are computed once and amortised over the evals of the
// group; consecutive threads write consecutive addresses of one eval (coalesced 16-byte
// stores).  Host-valued leaves (VEC / DENSE) read their per-eval value instead.
template <typename T>
__global__ void __launch_bounds__(256)
leaf_build_kernel(const b2_leaf_desc* __restrict__ descs, int n_leaves,
                  const int* __restrict__ block_leaf, long long arena_elems,
                  const double2* __restrict__ params, long long params_stride,
                  const double2* __restrict__ shared_params,
                  typename cplx<T>::type* __restrict__ arena, long long arena_stride,
                  int batch, int ev_per) {
    using C = typename cplx<T>::type;
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= arena_elems) return;
    int lo, hi;
    if (block_leaf) { lo = block_leaf[blockIdx.x]; hi = block_leaf[blockIdx.x + 1]; }
    else { lo = 0; hi = n_leaves - 1; }
    while (lo < hi) {                    // largest leaf with offset <= e
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&descs[mid].offset) <= e) lo = mid; else hi = mid - 1;
    }
    const b2_leaf_desc& dref = descs[lo];
    b2_leaf_desc d;
    d.kind = __ldg(&dref.kind);
    d.ndim = __ldg(&dref.ndim);
    d.ipar = __ldg(&dref.ipar);
    const unsigned local = (unsigned)(e - __ldg(&dref.offset));
    int idx[B2_MAX_LEAF_RANK];
    bool from_params = d.kind == B2_K_VEC || d.kind == B2_K_DENSE;
    double v = 0.0;
    if (!from_params) {
        unsigned rem = local;
        #pragma unroll
        for (int a = B2_MAX_LEAF_RANK - 1; a >= 0; --a) {
            if (a < d.ndim) {
                const unsigned dim = (unsigned)__ldg(&dref.dims[a]);
                d.dims[a] = (int)dim;
                const unsigned q = rem / dim;
                idx[a] = (int)(rem - q * dim);
                rem = q;
            } else idx[a] = 0;
        }
        v = leaf_value(d, idx, from_params);
    }
    const long long poff = from_params ? __ldg(&dref.poff) + (long long)local : 0;
    // host values that are the same for every eval of the batch live in one shared vector
    const bool shared_vals = from_params && __ldg(&dref.shared) != 0;
    for (int g = blockIdx.y; (long long)g * ev_per < batch; g += gridDim.y) {
        const int ev0 = g * ev_per, ev1 = min(batch, ev0 + ev_per);
        C* out = arena + (long long)ev0 * arena_stride + e;
        if (shared_vals) {
            const double2 pv = shared_params[poff];
            C val; val.x = (T)pv.x; val.y = (T)pv.y;
            for (int ev = ev0; ev < ev1; ++ev) { *out = val; out += arena_stride; }
        } else if (from_params) {
            const double2* par = params + (long long)ev0 * params_stride + poff;
            for (int ev = ev0; ev < ev1; ++ev) {
                const double2 pv = *par;
                C val; val.x = (T)pv.x; val.y = (T)pv.y;
                *out = val;
                out += arena_stride; par += params_stride;
            }
        } else {
            C val; val.x = (T)v; val.y = (T)0;
            for (int ev = ev0; ev < ev1; ++ev) { *out = val; out += arena_stride; }
        }
    }
}

}  // namespace b2

// block_leaf_dev: for every 256-element block of the arena the index of the leaf that
// holds its first element, plus one trailing entry (= index of the leaf holding the last
// element); may be null (then every thread searches the whole table).
extern "C" int b2_build_leaves_indexed(const b2_leaf_desc* desc_dev, int32_t n_leaves,
                                       const int32_t* block_leaf_dev, int64_t arena_elems,
                                       const void* params_dev, int64_t params_stride,
                                       const void* shared_params_dev, void* arena_dev,
                                       int64_t arena_stride, int32_t batch, int32_t dtype,
                                       void* stream) {
    if (n_leaves <= 0 || arena_elems <= 0 || batch <= 0) return 0;
    if (!desc_dev || !arena_dev) { b2::set_error("b2_build_leaves: null pointer"); return 1; }
    int sms = b2::num_sms();
    if (sms <= 0) return 1;
    const long long blocks = (arena_elems + 255) / 256;
    if (blocks > 2147483647ll) { b2::set_error("b2_build_leaves: arena too large"); return 1; }
    // evals per thread: enough CTAs to fill the GPU a few times over, long enough store
    // runs to amortise the decode
    static const int ctas_per_sm = getenv("B2_LEAF_CTAS_PER_SM") ? atoi(getenv("B2_LEAF_CTAS_PER_SM")) : 16;
    long long ev_per = (long long)batch * blocks / ((long long)sms * ctas_per_sm);
    if (ev_per < 1) ev_per = 1;
    if (ev_per > 64) ev_per = 64;
    long long gy = (batch + ev_per - 1) / ev_per;
    if (gy > 65535) gy = 65535;
    dim3 grid((unsigned)blocks, (unsigned)gy);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == B2_C128)
        b2::leaf_build_kernel<double><<<grid, 256, 0, s>>>(
            desc_dev, n_leaves, block_leaf_dev, arena_elems, (const double2*)params_dev,
            params_stride, (const double2*)shared_params_dev, (double2*)arena_dev, arena_stride,
            batch, (int)ev_per);
    else if (dtype == B2_C64)
        b2::leaf_build_kernel<float><<<grid, 256, 0, s>>>(
            desc_dev, n_leaves, block_leaf_dev, arena_elems, (const double2*)params_dev,
            params_stride, (const double2*)shared_params_dev, (float2*)arena_dev, arena_stride,
            batch, (int)ev_per);
    else { b2::set_error("b2_build_leaves: bad dtype %d", dtype); return 1; }
    return b2::check_launch("b2_build_leaves");
}

extern "C" int b2_build_leaves(const b2_leaf_desc* desc_dev, int32_t n_leaves,
                               int64_t arena_elems, const void* params_dev,
                               int64_t params_stride, void* arena_dev,
                               int64_t arena_stride, int32_t batch, int32_t dtype,
                               void* stream) {
    return b2_build_leaves_indexed(desc_dev, n_leaves, nullptr, arena_elems, params_dev,
                                   params_stride, nullptr, arena_dev, arena_stride, batch, dtype,
                                   stream);
}
