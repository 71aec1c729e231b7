// Piece (4): result tensor -> probabilities, restating EvalResult._prob_dist_pure
// (optyx/core/backends.py:291-304) and _prob_dist_mixed (306-356) on the device.
// Both kernels stream the result tensor once (HBM-read-bound, 16 B/element).
#include "b2_common.cuh"

namespace b2 {

// AMP: p[i] = |t[i]|^2 ; nz[i] = exact non-zero test of np.flatnonzero (274-289)
template <typename T>
__global__ void __launch_bounds__(256)
probs_amp_kernel(long long n, const typename cplx<T>::type* __restrict__ t,
                 double* __restrict__ p, uint8_t* __restrict__ nz) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const auto v = t[i];
        const double re = (double)v.x, im = (double)v.y;
        // abs(v) ** 2 in the reference is hypot(re, im) squared
        const double a = hypot(re, im);
        p[i] = a * a;
        nz[i] = (v.x != (T)0 || v.y != (T)0) ? 1 : 0;
    }
}

struct dm_layout {
    int ndim;
    int dims[B2_MAX_MODES];
    int kind[B2_MAX_MODES];   // 1 measured, 2 ket, 3 bra
};

// pass 1 (streaming, coalesced): seen[meas] = 1 for every non-zero entry
// (backends.py:339-341: every measured key that occurs is kept, 353-354).
// The axes are split into a SLOW group (decoded once per tile with div/mod) and a FAST
// group of `fast` elements (contiguous in memory) whose contribution to the measured
// key is tabulated once per CTA in shared memory: per element one 16-byte load, one
// table lookup, one compare.  A key is stored only while it still reads 0 (benign race:
// every writer stores 1), so a dense tensor does not hammer a handful of addresses.
constexpr int DM_FAST_MAX = 4096;

template <typename T>
__global__ void __launch_bounds__(256)
probs_dm_seen_kernel(const __grid_constant__ dm_layout L, int n_slow_axes, int fast,
                     long long n_tiles, long long fast_key_mul,
                     const typename cplx<T>::type* __restrict__ t,
                     uint8_t* __restrict__ seen) {
    __shared__ unsigned tab[DM_FAST_MAX];
    // key contribution of every position inside the fast block (axes n_slow_axes..ndim)
    for (int f = threadIdx.x; f < fast; f += blockDim.x) {
        int rem = f;
        unsigned key = 0, mul = 1;
        for (int a = L.ndim - 1; a >= n_slow_axes; --a) {
            const int d = L.dims[a];
            const int i = rem % d;
            rem /= d;
            if (L.kind[a] == 1) { key += (unsigned)i * mul; mul *= (unsigned)d; }
        }
        tab[f] = key;
    }
    __syncthreads();
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // measured key of the slow axes of this tile
        long long rem = tile, key_slow = 0, mul = fast_key_mul;
        for (int a = n_slow_axes - 1; a >= 0; --a) {
            const int d = L.dims[a];
            const long long i = rem % d;
            rem /= d;
            if (L.kind[a] == 1) { key_slow += i * mul; mul *= d; }
        }
        const typename cplx<T>::type* row = t + tile * (long long)fast;
        for (int f = threadIdx.x; f < fast; f += blockDim.x) {
            const auto v = row[f];
            if (v.x != (T)0 || v.y != (T)0) {
                uint8_t* s = seen + key_slow + tab[f];
                if (*s == 0) *s = 1;
            }
        }
    }
}

// pass 2: the diagonal of the unmeasured (ket, bra) pairs of every measured key.
// `lanes` threads (a power of two <= 1024: a whole CTA, a sub-warp group, or one
// thread) share one key: lane l sums diagonal entries l, l + lanes, ... in increasing
// order and the partial sums are combined by a fixed butterfly, so the result is
// deterministic.  The loads are strided by construction (the diagonal of a C-ordered
// tensor); what matters is that all of them are in flight at once.
template <typename T>
__global__ void __launch_bounds__(256)
probs_dm_sum_kernel(const __grid_constant__ dm_layout L, long long n_meas, long long n_diag,
                    int lanes, const typename cplx<T>::type* __restrict__ t,
                    double* __restrict__ p, double* __restrict__ max_imag) {
    __shared__ double red[8], redm[8];
    long long stride[B2_MAX_MODES];
    {
        long long s = 1;
        for (int a = L.ndim - 1; a >= 0; --a) { stride[a] = s; s *= L.dims[a]; }
    }
    const int keys_per_cta = lanes >= 256 ? 1 : 256 / lanes;
    const int lane = lanes >= 256 ? threadIdx.x : threadIdx.x % lanes;
    const int step = lanes >= 256 ? 256 : lanes;
    double cta_max = 0.0;
    for (long long k0 = (long long)blockIdx.x * keys_per_cta; k0 < n_meas;
         k0 += (long long)gridDim.x * keys_per_cta) {
        const long long k = k0 + (lanes >= 256 ? 0 : threadIdx.x / lanes);
        const bool valid = k < n_meas;
        long long base = 0, rem = valid ? k : 0;
        for (int a = L.ndim - 1; a >= 0; --a)
            if (L.kind[a] == 1) { base += (rem % L.dims[a]) * stride[a]; rem /= L.dims[a]; }
        double acc = 0.0, mx = 0.0;
        if (valid)
            for (long long u = lane; u < n_diag; u += step) {
                long long off = base, r = u;
                for (int a = L.ndim - 1; a >= 0; --a)
                    if (L.kind[a] == 3) {   // bra half; its ket twin is axis a-1
                        const long long i = r % L.dims[a];
                        r /= L.dims[a];
                        off += i * (stride[a] + stride[a - 1]);
                    }
                const auto v = t[off];
                if (v.x == (T)0 && v.y == (T)0) continue;   // not a dict entry
                double val = (double)v.x;
                const double im = fabs((double)v.y);
                if (im > mx) mx = im;
                if (val < 0 && fabs(val) < 1e-12) val = 0.0;   // backends.py:349-350
                acc += val;
            }
        // combine the lanes of one key
        const int width = lanes >= 32 ? 32 : lanes;
        for (int m = width >> 1; m > 0; m >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, m);
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        }
        if (lanes >= 256) {                    // one key per CTA: 8 warps
            __syncthreads();
            if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = acc; redm[threadIdx.x >> 5] = mx; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double a8 = 0.0, m8 = 0.0;
                for (int w = 0; w < 8; ++w) { a8 += red[w]; m8 = fmax(m8, redm[w]); }
                p[k] = a8;
                if (m8 > cta_max) cta_max = m8;
            }
        } else if (lanes > 32) {               // 64 or 128 lanes: 2 or 4 warps per key
            __syncthreads();
            if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = acc; redm[threadIdx.x >> 5] = mx; }
            __syncthreads();
            const int wpk = lanes / 32;
            if (valid && lane == 0) {
                const int w0 = (threadIdx.x >> 5);
                double a8 = 0.0, m8 = 0.0;
                for (int w = 0; w < wpk; ++w) { a8 += red[w0 + w]; m8 = fmax(m8, redm[w0 + w]); }
                p[k] = a8;
                if (m8 > cta_max) cta_max = m8;
            }
        } else if (valid && lane == 0) {
            p[k] = acc;
            if (mx > cta_max) cta_max = mx;
        }
    }
    // max |Im| (np.real_if_close guard, backends.py:348): non-negative doubles
    // order like their bit patterns
    if (cta_max > 0.0)
        atomicMax((unsigned long long*)max_imag, (unsigned long long)__double_as_longlong(cta_max));
}

template <typename T>
__global__ void __launch_bounds__(256)
axpy_kernel(long long n, T ar, T ai, const typename cplx<T>::type* __restrict__ x,
            typename cplx<T>::type* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const auto a = x[i];
        auto b = y[i];
        b.x += ar * a.x - ai * a.y;
        b.y += ar * a.y + ai * a.x;
        y[i] = b;
    }
}

static unsigned grid_for(long long n, int threads, int per_sm) {
    long long blocks = (n + threads - 1) / threads;
    long long cap = (long long)num_sms() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace b2

extern "C" int b2_probs_amp(int64_t n, const void* t_dev, double* p_dev, uint8_t* nz_dev,
                            int32_t dtype, void* stream) {
    if (n <= 0) return 0;
    if (b2::num_sms() <= 0) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = b2::grid_for(n, 256, 8);
    if (dtype == B2_C128)
        b2::probs_amp_kernel<double><<<g, 256, 0, s>>>(n, (const double2*)t_dev, p_dev, nz_dev);
    else if (dtype == B2_C64)
        b2::probs_amp_kernel<float><<<g, 256, 0, s>>>(n, (const float2*)t_dev, p_dev, nz_dev);
    else { b2::set_error("b2_probs_amp: bad dtype"); return 1; }
    return b2::check_launch("b2_probs_amp");
}

extern "C" int b2_probs_dm(int32_t ndim, const int32_t* dims, const int32_t* axis_kind,
                           const void* t_dev, double* p_dev, uint8_t* seen_dev,
                           double* max_imag_dev, int32_t dtype, void* stream) {
    if (ndim < 0 || ndim > B2_MAX_MODES) { b2::set_error("b2_probs_dm: bad ndim"); return 1; }
    if (b2::num_sms() <= 0) return 1;
    b2::dm_layout L;
    L.ndim = ndim;
    long long n = 1, n_meas = 1, n_diag = 1;
    for (int a = 0; a < ndim; ++a) {
        L.dims[a] = dims[a];
        L.kind[a] = axis_kind[a];
        n *= dims[a];
        if (axis_kind[a] == 1) n_meas *= dims[a];
        else if (axis_kind[a] == 3) {
            if (a == 0 || axis_kind[a - 1] != 2 || dims[a - 1] != dims[a]) {
                b2::set_error("b2_probs_dm: bra axis %d has no ket twin", a);
                return 1;
            }
            n_diag *= dims[a];
        } else if (axis_kind[a] != 2) { b2::set_error("b2_probs_dm: bad axis kind"); return 1; }
    }
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(seen_dev, 0, (size_t)n_meas, s);
    cudaMemsetAsync(max_imag_dev, 0, sizeof(double), s);
    // pass 1 split: trailing axes whose extents multiply to <= DM_FAST_MAX form the
    // fast block (at least one axis, unless it alone is larger: then no fast axes)
    int n_slow = ndim;
    long long fast = 1, fast_key_mul = 1;
    while (n_slow > 0 && fast * dims[n_slow - 1] <= b2::DM_FAST_MAX) {
        --n_slow;
        fast *= dims[n_slow];
        if (axis_kind[n_slow] == 1) fast_key_mul *= dims[n_slow];
    }
    const long long n_tiles = n / fast;
    // pass 2 parallelism: as many lanes per key as its diagonal has entries
    int lanes = 1;
    while (lanes < 256 && lanes < n_diag) lanes <<= 1;
    const long long keys_per_cta = lanes >= 256 ? 1 : 256 / lanes;
    long long g2l = (n_meas + keys_per_cta - 1) / keys_per_cta;
    const long long cap2 = (long long)b2::num_sms() * 8;
    if (g2l > cap2) g2l = cap2;
    if (g2l < 1) g2l = 1;
    long long g1l = n_tiles;
    const long long cap1 = (long long)b2::num_sms() * 8;
    if (g1l > cap1) g1l = cap1;
    if (g1l < 1) g1l = 1;
    const unsigned g1 = (unsigned)g1l, g2 = (unsigned)g2l;
    if (dtype == B2_C128) {
        b2::probs_dm_seen_kernel<double><<<g1, 256, 0, s>>>(L, n_slow, (int)fast, n_tiles, fast_key_mul,
                                                          (const double2*)t_dev, seen_dev);
        b2::probs_dm_sum_kernel<double><<<g2, 256, 0, s>>>(L, n_meas, n_diag, lanes,
                                                         (const double2*)t_dev, p_dev, max_imag_dev);
    } else if (dtype == B2_C64) {
        b2::probs_dm_seen_kernel<float><<<g1, 256, 0, s>>>(L, n_slow, (int)fast, n_tiles, fast_key_mul,
                                                         (const float2*)t_dev, seen_dev);
        b2::probs_dm_sum_kernel<float><<<g2, 256, 0, s>>>(L, n_meas, n_diag, lanes,
                                                        (const float2*)t_dev, p_dev, max_imag_dev);
    } else { b2::set_error("b2_probs_dm: bad dtype"); return 1; }
    return b2::check_launch("b2_probs_dm");
}

extern "C" int b2_axpy(int64_t n, double alpha_re, double alpha_im, const void* x_dev,
                       void* y_dev, int32_t dtype, void* stream) {
    if (n <= 0) return 0;
    if (b2::num_sms() <= 0) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = b2::grid_for(n, 256, 8);
    if (dtype == B2_C128)
        b2::axpy_kernel<double><<<g, 256, 0, s>>>(n, alpha_re, alpha_im, (const double2*)x_dev,
                                                 (double2*)y_dev);
    else if (dtype == B2_C64)
        b2::axpy_kernel<float><<<g, 256, 0, s>>>(n, (float)alpha_re, (float)alpha_im,
                                                (const float2*)x_dev, (float2*)y_dev);
    else { b2::set_error("b2_axpy: bad dtype"); return 1; }
    return b2::check_launch("b2_axpy");
}

extern "C" int b2_fill_zero(int64_t n, void* y_dev, int32_t dtype, void* stream) {
    if (n <= 0) return 0;
    size_t bytes = (size_t)n * (dtype == B2_C128 ? 16 : 8);
    cudaError_t e = cudaMemsetAsync(y_dev, 0, bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) { b2::set_error("b2_fill_zero: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}
