// Piece (1) of the hot path: build every dense truncated-Fock / ZX leaf array of
// an eval (or a batch of evals) directly in HBM with ONE launch.
//
// Reference: Box.truncation (optyx/core/diagram.py:536-590) loops in Python over
// every input basis state and calls a generator per state; here each thread
// owns one arena element, decodes its multi-index and evaluates the closed-form
// predicate of the box.  Write-only, HBM-write-bound: 16 B (complex128) per
// element, no reads except the tiny descriptor table (L1/L2 resident) and the
// host-computed params of VEC/DENSE leaves.
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

template <typename T>
__global__ void __launch_bounds__(256)
leaf_build_kernel(const b2_leaf_desc* __restrict__ descs, int n_leaves, long long arena_elems,
                  const double2* __restrict__ params, long long params_stride,
                  typename cplx<T>::type* __restrict__ arena, long long arena_stride) {
    using C = typename cplx<T>::type;
    const int ev = blockIdx.y;
    C* out = arena + (long long)ev * arena_stride;
    const double2* par = params ? params + (long long)ev * params_stride : nullptr;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < arena_elems;
         e += (long long)gridDim.x * blockDim.x) {
        // largest leaf with offset <= e
        int lo = 0, hi = n_leaves - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (descs[mid].offset <= e) lo = mid; else hi = mid - 1;
        }
        const b2_leaf_desc d = descs[lo];
        long long local = e - d.offset;
        int idx[B2_MAX_LEAF_RANK];
        long long rem = local;
        #pragma unroll
        for (int a = B2_MAX_LEAF_RANK - 1; a >= 0; --a) {
            if (a < d.ndim) { idx[a] = (int)(rem % d.dims[a]); rem /= d.dims[a]; }
            else idx[a] = 0;
        }
        bool from_params;
        double v = leaf_value(d, idx, from_params);
        C val;
        if (from_params) {
            double2 p = par[d.poff + local];
            val.x = (T)p.x; val.y = (T)p.y;
        } else {
            val.x = (T)v; val.y = (T)0;
        }
        out[e] = val;
    }
}

}  // namespace b2

extern "C" int b2_build_leaves(const b2_leaf_desc* desc_dev, int32_t n_leaves,
                               int64_t arena_elems, const void* params_dev,
                               int64_t params_stride, void* arena_dev,
                               int64_t arena_stride, int32_t batch, int32_t dtype,
                               void* stream) {
    if (n_leaves <= 0 || arena_elems <= 0 || batch <= 0) return 0;
    if (!desc_dev || !arena_dev) { b2::set_error("b2_build_leaves: null pointer"); return 1; }
    const int threads = 256;
    long long blocks = (arena_elems + threads - 1) / threads;
    int sms = b2::num_sms();
    if (sms <= 0) return 1;
    long long cap = (long long)sms * 8;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)batch);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == B2_C128)
        b2::leaf_build_kernel<double><<<grid, threads, 0, s>>>(
            desc_dev, n_leaves, arena_elems, (const double2*)params_dev, params_stride,
            (double2*)arena_dev, arena_stride);
    else if (dtype == B2_C64)
        b2::leaf_build_kernel<float><<<grid, threads, 0, s>>>(
            desc_dev, n_leaves, arena_elems, (const double2*)params_dev, params_stride,
            (float2*)arena_dev, arena_stride);
    else { b2::set_error("b2_build_leaves: bad dtype %d", dtype); return 1; }
    return b2::check_launch("b2_build_leaves");
}
