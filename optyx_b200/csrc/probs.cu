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

// pass 1 (elementwise, coalesced): seen[meas] = 1 for every non-zero entry
// (backends.py:339-341: every measured key that occurs is kept, 353-354)
template <typename T>
__global__ void __launch_bounds__(256)
probs_dm_seen_kernel(const __grid_constant__ dm_layout L, long long n,
                     const typename cplx<T>::type* __restrict__ t,
                     uint8_t* __restrict__ seen) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        const auto v = t[e];
        if (v.x == (T)0 && v.y == (T)0) continue;
        long long rem = e, meas = 0, mul = 1;
        for (int a = L.ndim - 1; a >= 0; --a) {
            const int d = L.dims[a];
            const int i = (int)(rem % d);
            rem /= d;
            if (L.kind[a] == 1) { meas += i * mul; mul *= d; }
        }
        seen[meas] = 1;
    }
}

// pass 2: one thread per measured key walks the unmeasured diagonal in C order
// (the order in which the reference's dict loop adds them, 339-351), so the
// floating-point sum is deterministic and ordered like the reference's.
template <typename T>
__global__ void __launch_bounds__(128)
probs_dm_sum_kernel(const __grid_constant__ dm_layout L, long long n_meas, long long n_diag,
                    const typename cplx<T>::type* __restrict__ t, double* __restrict__ p,
                    double* __restrict__ max_imag) {
    double local_max = 0.0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n_meas;
         k += (long long)gridDim.x * blockDim.x) {
        // strides of every axis
        long long stride[B2_MAX_MODES];
        long long s = 1;
        for (int a = L.ndim - 1; a >= 0; --a) { stride[a] = s; s *= L.dims[a]; }
        // base offset from the measured key (C order over measured axes)
        long long base = 0, rem = k;
        for (int a = L.ndim - 1; a >= 0; --a)
            if (L.kind[a] == 1) { base += (rem % L.dims[a]) * stride[a]; rem /= L.dims[a]; }
        double acc = 0.0;
        for (long long u = 0; u < n_diag; ++u) {
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
            if (im > local_max) local_max = im;
            if (val < 0 && fabs(val) < 1e-12) val = 0.0;   // backends.py:349-350
            acc += val;
        }
        p[k] = acc;
    }
    // max |Im| (np.real_if_close guard, backends.py:348): non-negative doubles
    // order like their bit patterns
    if (local_max > 0.0)
        atomicMax((unsigned long long*)max_imag, (unsigned long long)__double_as_longlong(local_max));
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
    unsigned g1 = b2::grid_for(n, 256, 8), g2 = b2::grid_for(n_meas, 128, 8);
    if (dtype == B2_C128) {
        b2::probs_dm_seen_kernel<double><<<g1, 256, 0, s>>>(L, n, (const double2*)t_dev, seen_dev);
        b2::probs_dm_sum_kernel<double><<<g2, 128, 0, s>>>(L, n_meas, n_diag,
                                                         (const double2*)t_dev, p_dev, max_imag_dev);
    } else if (dtype == B2_C64) {
        b2::probs_dm_seen_kernel<float><<<g1, 256, 0, s>>>(L, n, (const float2*)t_dev, seen_dev);
        b2::probs_dm_sum_kernel<float><<<g2, 128, 0, s>>>(L, n_meas, n_diag,
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
