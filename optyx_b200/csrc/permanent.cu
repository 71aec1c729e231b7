// GPU permanents for pure linear-optical circuits -- the other
// evaluation algorithm the reference offers for this path (PermanentBackend,
// optyx/core/backends.py:1019-1098; amplitude formula Matrix.eval path.py:345-384;
// permanent = Glynn's formula in Gray-code order, npperm path.py:141-163).
//
//   amp[c] = norm[c] * perm(A_c),   A_c[i][j] = U[rows[i]][cols_c[j]]   (n x n)
//   perm(A) = 2^(1-n) * sum_{delta in {+-1}^n, delta_0 = +1} (prod_k delta_k) prod_j (sum_i delta_i A[i][j])
//
// One warp per output configuration.  The 2^(n-1) sign vectors are split into 32
// contiguous Gray-code ranges, one per lane: a lane builds its column sums for the
// first code of its range (n^2 adds), then walks the range with one row update
// (n complex adds) and one n-term complex product per code.  U lives in shared
// memory (m <= 64: 64 KB); the warp's n x n submatrix is gathered into registers-
// adjacent shared memory once.  FP64 throughout (compute bound: ~8 n flops per code).
#include "b2_common.cuh"

namespace b2 {

constexpr int PM_WARPS = 4;          // warps (configurations) per CTA

// NMAX: compile-time bound of n -- the column sums v[] stay in registers (fully
// unrolled, predicated loops) for NMAX <= 16
template <int NMAX>
__global__ void __launch_bounds__(PM_WARPS * 32)
permanent_kernel(const double2* __restrict__ U, int m, int n, const int* __restrict__ rows_g,
                 const unsigned char* __restrict__ cols, const double* __restrict__ norm,
                 long long n_configs, double2* __restrict__ amp) {
    extern __shared__ __align__(16) unsigned char pm_smem[];
    double2* Us = reinterpret_cast<double2*>(pm_smem);                 // [m][m]
    double2* As = Us + m * m;                                          // [PM_WARPS][n][n]
    __shared__ int rows[B2_PERM_MAX_N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < m * m; i += blockDim.x) Us[i] = U[i];
    if (threadIdx.x < n) rows[threadIdx.x] = rows_g[threadIdx.x];
    __syncthreads();
    double2* A = As + warp * n * n;
    const long long total = n > 0 ? (1ll << (n - 1)) : 1;
    for (long long c = (long long)blockIdx.x * PM_WARPS + warp; c < n_configs;
         c += (long long)gridDim.x * PM_WARPS) {
        const unsigned char* cc = cols + c * n;
        __syncwarp();
        for (int e = lane; e < n * n; e += 32) {
            const int i = e / n, j = e - i * n;
            A[e] = Us[rows[i] * m + cc[j]];
        }
        __syncwarp();
        double sx = 0.0, sy = 0.0;
        if (n == 0) {
            if (lane == 0) sx = 1.0;
        } else {
            // this lane's range of Gray-code ranks [g0, g1)
            const long long per = (total + 31) / 32;
            const long long g0 = lane * per, g1 = g0 + per < total ? g0 + per : total;
            if (g0 < g1) {
                // sign vector of rank g: bits of gray(g) = g ^ (g >> 1) say which of rows
                // 1..n-1 are negated (row 0 keeps delta = +1)
                double2 v[NMAX];
                long long gray = g0 ^ (g0 >> 1);
                int sign = 1;
                #pragma unroll
                for (int j = 0; j < NMAX; ++j) {
                    if (j < n) {
                        double2 s = A[j];                              // row 0
                        for (int i = 1; i < n; ++i) {
                            const double2 a = A[i * n + j];
                            if ((gray >> (i - 1)) & 1) { s.x -= a.x; s.y -= a.y; }
                            else { s.x += a.x; s.y += a.y; }
                        }
                        v[j] = s;
                    }
                }
                sign = (__popcll(gray) & 1) ? -1 : 1;
                for (long long gi = g0; gi < g1; ++gi) {
                    if (gi != g0) {
                        // rank gi differs from gi - 1 in bit ctz(gi): flip that row
                        const int b = __ffsll(gi) - 1;                 // row b + 1
                        const bool now_neg = ((gi ^ (gi >> 1)) >> b) & 1;
                        const double f = now_neg ? -2.0 : 2.0;
                        const double2* ar = A + (b + 1) * n;
                        #pragma unroll
                        for (int j = 0; j < NMAX; ++j) {
                            if (j < n) {
                                v[j].x = fma(f, ar[j].x, v[j].x);
                                v[j].y = fma(f, ar[j].y, v[j].y);
                            }
                        }
                        sign = -sign;
                    }
                    double px = v[0].x, py = v[0].y;
                    #pragma unroll
                    for (int j = 1; j < NMAX; ++j) {
                        if (j < n) {
                            const double tx = px * v[j].x - py * v[j].y;
                            py = px * v[j].y + py * v[j].x;
                            px = tx;
                        }
                    }
                    sx += sign * px; sy += sign * py;
                }
            }
        }
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, off);
            sy += __shfl_xor_sync(0xffffffffu, sy, off);
        }
        if (lane == 0) {
            const double scale = norm[c] * (n > 0 ? ldexp(1.0, 1 - n) : 1.0);
            amp[c] = make_double2(sx * scale, sy * scale);
        }
    }
}

// ---- histogram variant: lossy / partially distinguishable interferometers ----------
// Every warp owns a contiguous range of configuration ranks: it unranks the first
// one (combinatorial number system over non-decreasing mode sequences), then steps
// to the next multiset in place.  Per configuration: gather the n x n submatrix of V,
// Glynn permanent as above, |perm|^2 / prod(out!) added to the bin of the detected
// pattern (shared-memory histogram per CTA, flushed once with FP64 atomics).
constexpr int PH_WARPS = 8;
constexpr int PH_SMEM_BINS = 4096;

template <int NMAX>
__global__ void __launch_bounds__(PH_WARPS * 32)
permanent_hist_kernel(const double2* __restrict__ V, int n, int m,
                      const long long* __restrict__ binom, const int* __restrict__ mode_stride,
                      double scale, long long n_configs, int n_bins, double* __restrict__ hist) {
    extern __shared__ __align__(16) unsigned char ph_smem[];
    double2* Vs = reinterpret_cast<double2*>(ph_smem);                 // [n][m]
    double2* As = Vs + n * m;                                          // [PH_WARPS][n][n]
    double* hs = reinterpret_cast<double*>(As + PH_WARPS * n * n);     // [min(n_bins, PH_SMEM_BINS)]
    int* strides = reinterpret_cast<int*>(hs + (n_bins <= PH_SMEM_BINS ? n_bins : 0));   // [m]
    const bool smem_hist = n_bins <= PH_SMEM_BINS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n * m; i += blockDim.x) Vs[i] = V[i];
    for (int i = threadIdx.x; i < m; i += blockDim.x) strides[i] = mode_stride[i];
    if (smem_hist) for (int i = threadIdx.x; i < n_bins; i += blockDim.x) hs[i] = 0.0;
    __syncthreads();
    double2* A = As + warp * n * n;
    const long long n_warps = (long long)gridDim.x * PH_WARPS;
    const long long wid = (long long)blockIdx.x * PH_WARPS + warp;
    const long long per_w = (n_configs + n_warps - 1) / n_warps;
    const long long r0 = wid * per_w, r1 = r0 + per_w < n_configs ? r0 + per_w : n_configs;
    const long long total = 1ll << (n - 1);
    const long long per = (total + 31) / 32;
    const long long g0 = lane * per, g1 = g0 + per < total ? g0 + per : total;
    int cols[NMAX];
    if (r0 < r1) {
        // unrank r0: number of non-decreasing sequences of length L over [v, m) is C(m-v+L-1, L)
        long long r = r0;
        int prev = 0;
        #pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            if (j < n) {
                const int L = n - j - 1;
                int v = prev;
                for (; v < m - 1; ++v) {
                    const long long cnt = binom[(long long)(m - v + L - 1) * (n + 1) + L];
                    if (r < cnt) break;
                    r -= cnt;
                }
                cols[j] = v; prev = v;
            }
        }
    }
    for (long long rk = r0; rk < r1; ++rk) {
        __syncwarp();
        for (int e = lane; e < n * n; e += 32) {
            const int i = e / n, j = e - i * n;
            int cj = 0;
            #pragma unroll
            for (int q = 0; q < NMAX; ++q) if (q == j) cj = cols[q];
            A[e] = Vs[i * m + cj];
        }
        __syncwarp();
        double sx = 0.0, sy = 0.0;
        if (g0 < g1) {
            double2 v[NMAX];
            const long long gray = g0 ^ (g0 >> 1);
            #pragma unroll
            for (int j = 0; j < NMAX; ++j) {
                if (j < n) {
                    double2 s = A[j];
                    for (int i = 1; i < n; ++i) {
                        const double2 a = A[i * n + j];
                        if ((gray >> (i - 1)) & 1) { s.x -= a.x; s.y -= a.y; }
                        else { s.x += a.x; s.y += a.y; }
                    }
                    v[j] = s;
                }
            }
            int sign = (__popcll(gray) & 1) ? -1 : 1;
            for (long long gi = g0; gi < g1; ++gi) {
                if (gi != g0) {
                    const int b = __ffsll(gi) - 1;
                    const bool now_neg = ((gi ^ (gi >> 1)) >> b) & 1;
                    const double f = now_neg ? -2.0 : 2.0;
                    const double2* ar = A + (b + 1) * n;
                    #pragma unroll
                    for (int j = 0; j < NMAX; ++j) {
                        if (j < n) {
                            v[j].x = fma(f, ar[j].x, v[j].x);
                            v[j].y = fma(f, ar[j].y, v[j].y);
                        }
                    }
                    sign = -sign;
                }
                double px = v[0].x, py = v[0].y;
                #pragma unroll
                for (int j = 1; j < NMAX; ++j) {
                    if (j < n) {
                        const double tx = px * v[j].x - py * v[j].y;
                        py = px * v[j].y + py * v[j].x;
                        px = tx;
                    }
                }
                sx += sign * px; sy += sign * py;
            }
        }
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, off);
            sy += __shfl_xor_sync(0xffffffffu, sy, off);
        }
        // bin and 1 / prod(out!) from the run lengths of the sorted mode sequence
        int bin = 0, run = 1;
        double inv = 1.0;
        #pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            if (j < n) {
                bin += strides[cols[j]];
                if (j > 0) {
                    bool same = false;
                    #pragma unroll
                    for (int q = 0; q < NMAX; ++q) if (q == j - 1) same = cols[q] == cols[j];
                    run = same ? run + 1 : 1;
                    inv /= (double)run;
                }
            }
        }
        if (lane == 0) {
            const double w = ldexp(1.0, 2 - 2 * n) * (sx * sx + sy * sy) * inv * scale;
            if (smem_hist) atomicAdd(&hs[bin], w); else atomicAdd(&hist[bin], w);
        }
        // next multiset: bump the right-most entry below m - 1, copy it to the right
        int jb = -1;
        #pragma unroll
        for (int j = 0; j < NMAX; ++j) if (j < n && cols[j] < m - 1) jb = j;
        if (jb >= 0) {
            int nv = 0;
            #pragma unroll
            for (int j = 0; j < NMAX; ++j) if (j == jb) nv = cols[j] + 1;
            #pragma unroll
            for (int j = 0; j < NMAX; ++j) if (j < n && j >= jb) cols[j] = nv;
        }
    }
    __syncthreads();
    if (smem_hist)
        for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
            if (hs[i] != 0.0) atomicAdd(&hist[i], hs[i]);
}

// Thread-per-configuration variant for few photons (2^(n-1) sign vectors are too few
// to feed 32 lanes: the per-lane set-up would dominate).  Every THREAD owns a
// contiguous range of ranks, keeps its column sums in registers and reads the
// photon rows straight from shared memory (no submatrix gather).
constexpr int PT_THREADS = 128;

template <int NMAX>
__global__ void __launch_bounds__(PT_THREADS)
permanent_hist_thread_kernel(const double2* __restrict__ V, int n, int m,
                             const long long* __restrict__ binom,
                             const int* __restrict__ mode_stride, double scale,
                             long long n_configs, int n_bins, double* __restrict__ hist) {
    extern __shared__ __align__(16) unsigned char pt_smem[];
    double2* Vs = reinterpret_cast<double2*>(pt_smem);                 // [n][m]
    double* hs = reinterpret_cast<double*>(Vs + n * m);
    int* strides = reinterpret_cast<int*>(hs + (n_bins <= PH_SMEM_BINS ? n_bins : 0));
    const bool smem_hist = n_bins <= PH_SMEM_BINS;
    for (int i = threadIdx.x; i < n * m; i += blockDim.x) Vs[i] = V[i];
    for (int i = threadIdx.x; i < m; i += blockDim.x) strides[i] = mode_stride[i];
    if (smem_hist) for (int i = threadIdx.x; i < n_bins; i += blockDim.x) hs[i] = 0.0;
    __syncthreads();
    const long long n_thr = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_t = (n_configs + n_thr - 1) / n_thr;
    const long long r0 = tid * per_t, r1 = r0 + per_t < n_configs ? r0 + per_t : n_configs;
    const long long total = 1ll << (n - 1);
    const double pre = ldexp(1.0, 2 - 2 * n) * scale;
    int cols[NMAX];
    if (r0 < r1) {
        long long r = r0;
        int prev = 0;
        #pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            if (j < n) {
                const int L = n - j - 1;
                int v = prev;
                for (; v < m - 1; ++v) {
                    const long long cnt = binom[(long long)(m - v + L - 1) * (n + 1) + L];
                    if (r < cnt) break;
                    r -= cnt;
                }
                cols[j] = v; prev = v;
            }
        }
    }
    for (long long rk = r0; rk < r1; ++rk) {
        // column sums for the all-plus sign vector, then the Gray-code walk
        double2 v[NMAX];
        #pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            if (j < n) {
                double2 s = make_double2(0.0, 0.0);
                for (int i = 0; i < n; ++i) { const double2 a = Vs[i * m + cols[j]]; s.x += a.x; s.y += a.y; }
                v[j] = s;
            }
        }
        double sx = 0.0, sy = 0.0;
        int sign = 1;
        for (long long gi = 0; gi < total; ++gi) {
            if (gi != 0) {
                const int b = __ffsll(gi) - 1;
                const bool now_neg = ((gi ^ (gi >> 1)) >> b) & 1;
                const double f = now_neg ? -2.0 : 2.0;
                const double2* vr = Vs + (b + 1) * m;
                #pragma unroll
                for (int j = 0; j < NMAX; ++j) {
                    if (j < n) {
                        const double2 a = vr[cols[j]];
                        v[j].x = fma(f, a.x, v[j].x);
                        v[j].y = fma(f, a.y, v[j].y);
                    }
                }
                sign = -sign;
            }
            double px = v[0].x, py = v[0].y;
            #pragma unroll
            for (int j = 1; j < NMAX; ++j) {
                if (j < n) {
                    const double tx = px * v[j].x - py * v[j].y;
                    py = px * v[j].y + py * v[j].x;
                    px = tx;
                }
            }
            sx += sign * px; sy += sign * py;
        }
        int bin = 0, run = 1;
        double inv = 1.0;
        #pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            if (j < n) {
                bin += strides[cols[j]];
                if (j > 0) {
                    run = cols[j - 1] == cols[j] ? run + 1 : 1;
                    inv /= (double)run;
                }
            }
        }
        const double w = pre * (sx * sx + sy * sy) * inv;
        if (w != 0.0) { if (smem_hist) atomicAdd(&hs[bin], w); else atomicAdd(&hist[bin], w); }
        int jb = -1;
        #pragma unroll
        for (int j = 0; j < NMAX; ++j) if (j < n && cols[j] < m - 1) jb = j;
        if (jb >= 0) {
            int nv = 0;
            #pragma unroll
            for (int j = 0; j < NMAX; ++j) if (j == jb) nv = cols[j] + 1;
            #pragma unroll
            for (int j = 0; j < NMAX; ++j) if (j < n && j >= jb) cols[j] = nv;
        }
    }
    __syncthreads();
    if (smem_hist)
        for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
            if (hs[i] != 0.0) atomicAdd(&hist[i], hs[i]);
}

}  // namespace b2

extern "C" int b2_permanent_histogram(const void* v_dev, int32_t n, int32_t m,
                                      const int64_t* binom_dev, const int32_t* mode_stride_dev,
                                      double scale, int64_t n_configs, int32_t n_bins,
                                      double* hist_dev, void* stream) {
    using namespace b2;
    if (n_configs <= 0) return 0;
    if (!v_dev || !binom_dev || !mode_stride_dev || !hist_dev) {
        set_error("b2_permanent_histogram: null pointer");
        return 1;
    }
    if (n < 1 || n > 16 || m < 1 || m > B2_PERM_HIST_MAX_M || n_bins < 1) {
        set_error("b2_permanent_histogram: n = %d (1..16), m = %d (1..%d), bins = %d", n, m,
                  B2_PERM_HIST_MAX_M, n_bins);
        return 1;
    }
    int sms = num_sms();
    if (sms <= 0) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)(n * m + PH_WARPS * n * n) * sizeof(double2) +
                        (size_t)(n_bins <= PH_SMEM_BINS ? n_bins : 0) * sizeof(double) +
                        (size_t)m * sizeof(int);
    long long blocks = (n_configs + PH_WARPS - 1) / PH_WARPS;
    const long long cap = (long long)sms * 4;
    if (blocks > cap) blocks = cap;
#define B2_PH_LAUNCH(NM)                                                                       \
    do {                                                                                       \
        static b2::once_flags attr_done;                                                         \
        if (!b2::is_done(attr_done)) {                                                                      \
            cudaFuncSetAttribute(permanent_hist_kernel<NM>,                                    \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);     \
            b2::mark_done(attr_done);                                                                  \
        }                                                                                      \
        permanent_hist_kernel<NM><<<(unsigned)blocks, PH_WARPS * 32, smem, s>>>(               \
            (const double2*)v_dev, n, m, (const long long*)binom_dev, mode_stride_dev, scale,  \
            n_configs, n_bins, hist_dev);                                                      \
    } while (0)
    if (n <= 8 && n_configs >= 4096) {
        // few photons, many configurations: one thread per configuration
        static b2::once_flags attr_done;
        if (!b2::is_done(attr_done)) {
            cudaFuncSetAttribute(permanent_hist_thread_kernel<8>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            b2::mark_done(attr_done);
        }
        const size_t smem_t = (size_t)(n * m) * sizeof(double2) +
                              (size_t)(n_bins <= PH_SMEM_BINS ? n_bins : 0) * sizeof(double) +
                              (size_t)m * sizeof(int);
        // one resident wave, every thread an equal share of the ranks
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, permanent_hist_thread_kernel<8>,
                                                          PT_THREADS, smem_t) != cudaSuccess || occ < 1)
            occ = 1;
        long long tb = (n_configs + PT_THREADS - 1) / PT_THREADS;
        const long long tcap = (long long)sms * occ;
        if (tb > tcap) tb = tcap;
        if (tb < 1) tb = 1;
        permanent_hist_thread_kernel<8><<<(unsigned)tb, PT_THREADS, smem_t, s>>>(
            (const double2*)v_dev, n, m, (const long long*)binom_dev, mode_stride_dev, scale,
            n_configs, n_bins, hist_dev);
        set_last_kernel("permanent_hist_thread");
        return check_launch("b2_permanent_histogram");
    }
    if (n <= 8) B2_PH_LAUNCH(8); else B2_PH_LAUNCH(16);
#undef B2_PH_LAUNCH
    set_last_kernel("permanent_hist");
    return check_launch("b2_permanent_histogram");
}

extern "C" int b2_permanent_amplitudes(const void* u_dev, int32_t m, const int32_t* rows_dev,
                                       int32_t n, const uint8_t* cols_dev,
                                       const double* norm_dev, int64_t n_configs,
                                       void* amp_dev, void* stream) {
    using namespace b2;
    if (n_configs <= 0) return 0;
    if (!u_dev || !cols_dev || !norm_dev || !amp_dev || !rows_dev) {
        set_error("b2_permanent_amplitudes: null pointer");
        return 1;
    }
    if (m < 1 || m > B2_PERM_MAX_M || n < 0 || n > B2_PERM_MAX_N) {
        set_error("b2_permanent_amplitudes: m = %d (max %d), n = %d (max %d)", m, B2_PERM_MAX_M, n,
                  B2_PERM_MAX_N);
        return 1;
    }
    int sms = num_sms();
    if (sms <= 0) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)(m * m + PM_WARPS * n * n) * sizeof(double2);
    long long blocks = (n_configs + PM_WARPS - 1) / PM_WARPS;
    const long long cap = (long long)sms * 8;
    if (blocks > cap) blocks = cap;
#define B2_PERM_LAUNCH(NM)                                                                     \
    do {                                                                                       \
        static b2::once_flags attr_done;                                                         \
        if (!b2::is_done(attr_done)) {                                                                      \
            cudaFuncSetAttribute(permanent_kernel<NM>,                                         \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);     \
            b2::mark_done(attr_done);                                                                  \
        }                                                                                      \
        permanent_kernel<NM><<<(unsigned)blocks, PM_WARPS * 32, smem, s>>>(                    \
            (const double2*)u_dev, m, n, rows_dev, cols_dev, norm_dev, n_configs,              \
            (double2*)amp_dev);                                                                \
    } while (0)
    if (n <= 8) B2_PERM_LAUNCH(8);
    else if (n <= 16) B2_PERM_LAUNCH(16);
    else B2_PERM_LAUNCH(B2_PERM_MAX_N);
#undef B2_PERM_LAUNCH
    set_last_kernel("permanent");
    return check_launch("b2_permanent_amplitudes");
}
