// Result hand-over: a device-side scan for the non-zero entries of a result
// tensor, so that a sparse result (a GHZ / Clifford density tensor, a
// photon-number-conserving amplitude tensor) crosses PCIe as a short
// (flat index, value) list instead of a mostly-zero multi-GB array.  The host
// side (B200Backend._to_host) scatters the list into a zero-filled array and
// falls back to the dense copy when the list overflows `cap`.
//
// One pass over the tensor at HBM read speed (16 B / element, no writes to
// speak of for a sparse tensor).  "Non-zero" is numpy's test (np.flatnonzero, the
// one EvalResult itself applies, backends.py:274-289): re != 0 or im != 0, so a
// -0.0 arrives as +0.0 and a NaN is kept.
#include "b2_common.cuh"

namespace b2 {

__device__ __forceinline__ bool any_bit(const double2 v) { return v.x != 0.0 || v.y != 0.0; }
__device__ __forceinline__ bool any_bit(const float2 v) { return v.x != 0.0f || v.y != 0.0f; }

constexpr int CP_THREADS = 256, CP_UNROLL = 4;

template <typename C>
__global__ void __launch_bounds__(CP_THREADS)
compact_nonzero_kernel(long long n, const C* __restrict__ t, unsigned long long cap,
                       long long* __restrict__ idx, C* __restrict__ val,
                       unsigned long long* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * CP_THREADS;
    // warp-uniform loop: every lane of a warp walks the same iterations
    for (long long base = (long long)blockIdx.x * CP_THREADS + (threadIdx.x & ~31); base < n;
         base += stride * CP_UNROLL) {
        C v[CP_UNROLL];
        bool nz[CP_UNROLL];
        #pragma unroll
        for (int u = 0; u < CP_UNROLL; ++u) {          // all loads first: CP_UNROLL in flight
            const long long i = base + u * stride + lane;
            nz[u] = false;
            if (i < n) { v[u] = __ldcs(t + i); nz[u] = any_bit(v[u]); }
        }
        #pragma unroll
        for (int u = 0; u < CP_UNROLL; ++u) {
            const unsigned m = __ballot_sync(0xffffffffu, nz[u]);
            if (m == 0) continue;
            unsigned long long pos0 = 0;
            if (lane == 0) pos0 = atomicAdd(count, (unsigned long long)__popc(m));
            pos0 = __shfl_sync(0xffffffffu, pos0, 0);
            if (pos0 >= cap) return;                   // list overflowed: the caller copies densely
            if (nz[u]) {
                const unsigned long long p = pos0 + __popc(m & ((1u << lane) - 1u));
                if (p < cap) { idx[p] = base + u * stride + lane; val[p] = v[u]; }
            }
        }
    }
}

}  // namespace b2

extern "C" int b2_compact_nonzero(int64_t n, const void* t_dev, int32_t dtype, int64_t cap,
                                  int64_t* idx_dev, void* val_dev, uint64_t* count_dev,
                                  void* stream) {
    if (n < 0 || cap < 0) { b2::set_error("b2_compact_nonzero: negative size"); return 1; }
    const int sms = b2::num_sms();
    if (sms <= 0) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(count_dev, 0, sizeof(uint64_t), s) != cudaSuccess) {
        b2::set_error("b2_compact_nonzero: memset failed");
        return 1;
    }
    if (n == 0) return 0;
    long long blocks = (n + b2::CP_THREADS - 1) / b2::CP_THREADS;
    const long long full = (long long)sms * 8;           // one resident wave
    if (blocks > full) blocks = full;
    if (dtype == B2_C128)
        b2::compact_nonzero_kernel<double2><<<(unsigned)blocks, b2::CP_THREADS, 0, s>>>(
            n, (const double2*)t_dev, (unsigned long long)cap, (long long*)idx_dev,
            (double2*)val_dev, (unsigned long long*)count_dev);
    else if (dtype == B2_C64)
        b2::compact_nonzero_kernel<float2><<<(unsigned)blocks, b2::CP_THREADS, 0, s>>>(
            n, (const float2*)t_dev, (unsigned long long)cap, (long long*)idx_dev,
            (float2*)val_dev, (unsigned long long*)count_dev);
    else { b2::set_error("b2_compact_nonzero: bad dtype"); return 1; }
    return b2::check_launch("b2_compact_nonzero");
}
