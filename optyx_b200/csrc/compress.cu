// the device side of compressed (approximate) contraction --
// what quimb's contract_compressed does for the reference when QuimbBackend is given a
// HyperCompressedOptimizer (optyx/core/backends.py:516-545): after a pairwise
// contraction every bond between the new tensor and a neighbour is truncated to at most
// max_bond singular values.
//
// The big reductions (the Gram matrices G_A = A^H A and G_B = B B^H over everything but
// the bond, and the application of the projectors) are ordinary pairwise contractions and
// run through b2_contract_direct / b2_contract_gemm.  This file holds the small dense
// part, ONE CTA per bond: with A = Q_A R_A, B = R_B' Q_B^H (R_A = L_A^(1/2) V_A^H and
// R_B' = V_B L_B^(1/2) from the eigen-decompositions of the Gram matrices) the bond
// matrix is M = R_A R_B' = U S W^H, and the oblique projectors
//     P_A = R_B' W_k S_k^(-1/2)          P_B = S_k^(-1/2) U_k^H R_A
// give A P_A P_B B = Q_A U_k S_k W_k^H Q_B^H, the rank-k truncation of A B.  Only the KEPT
// singular values are inverted.  All three factorisations are one-sided (Hestenes)
// Jacobi sweeps with the round-robin pair order, one warp per column pair.
#include "b2_common.cuh"

namespace b2 {

constexpr int BC_THREADS = 512;
constexpr int BC_MAX_BOND = 256;
constexpr int BC_SWEEPS = 40;

__device__ __forceinline__ double2 cmulc(const double2 a, const double2 b) {   // conj(a) * b
    return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// One-sided Jacobi on the n x n column-major matrix X (columns are rotated until they are
// mutually orthogonal); the same rotations are applied to V (starts as the identity):
// on return X = X0 V, column j of X = u_j * s_j.  Called by every thread of the CTA.
__device__ void jacobi_onesided(double2* X, double2* V, int n, int* s_rot) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int i = tid; i < n * n; i += blockDim.x) {
        const int r = i % n, c = i / n;
        V[i] = make_double2(r == c ? 1.0 : 0.0, 0.0);
    }
    __syncthreads();
    const int ne = n + (n & 1), m = ne - 1;               // round robin over an even count
    for (int sweep = 0; sweep < BC_SWEEPS; ++sweep) {
        if (tid == 0) *s_rot = 0;
        __syncthreads();
        for (int r = 0; r < m; ++r) {
            for (int k = warp; k < ne / 2; k += nwarps) {
                int p = (r + k) % m, q = k == 0 ? m : (r - k + m) % m;
                if (p > q) { const int t = p; p = q; q = t; }
                if (q >= n) continue;                      // the padding column
                double2* a = X + (size_t)p * n;
                double2* b = X + (size_t)q * n;
                double alpha = 0, beta = 0, gr = 0, gi = 0;
                for (int i = lane; i < n; i += 32) {
                    const double2 x = a[i], y = b[i];
                    alpha += x.x * x.x + x.y * x.y;
                    beta += y.x * y.x + y.y * y.y;
                    const double2 g = cmulc(x, y);
                    gr += g.x; gi += g.y;
                }
                alpha = warp_sum(alpha); beta = warp_sum(beta);
                gr = warp_sum(gr); gi = warp_sum(gi);
                const double g = sqrt(gr * gr + gi * gi);
                if (g == 0.0 || g <= 1e-15 * sqrt(alpha * beta)) continue;
                if (lane == 0) atomicAdd(s_rot, 1);
                const double phr = gr / g, phi = gi / g;   // e^{i phi} of a^H b
                const double zeta = (beta - alpha) / (2.0 * g);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                // b~ = conj(e^{i phi}) b ;  a' = c a - s b~ ;  b' = s a + c b~
                for (int pass = 0; pass < 2; ++pass) {
                    double2* xa = (pass ? V : X) + (size_t)p * n;
                    double2* xb = (pass ? V : X) + (size_t)q * n;
                    for (int i = lane; i < n; i += 32) {
                        const double2 u = xa[i], w = xb[i];
                        const double2 wt = make_double2(w.x * phr + w.y * phi, w.y * phr - w.x * phi);
                        xa[i] = make_double2(c * u.x - s * wt.x, c * u.y - s * wt.y);
                        xb[i] = make_double2(s * u.x + c * wt.x, s * u.y + c * wt.y);
                    }
                }
            }
            __syncthreads();
        }
        if (*s_rot == 0) break;
        __syncthreads();
    }
    __syncthreads();
}

// scratch: 7 n x n complex128 matrices + 3 n doubles.
__global__ void __launch_bounds__(BC_THREADS)
bond_compress_kernel(const double2* __restrict__ GA, const double2* __restrict__ GB, int n,
                     int max_bond, double cutoff, double2* __restrict__ PA,
                     double2* __restrict__ PB, int* __restrict__ k_out,
                     double* __restrict__ sv_out, double2* __restrict__ scratch) {
    __shared__ int s_rot;
    __shared__ double s_max;
    const int tid = threadIdx.x, nn = n * n;
    double2* XA = scratch;            // G_A -> U_A L_A
    double2* VA = XA + nn;
    double2* XB = VA + nn;
    double2* VB = XB + nn;
    double2* RA = VB + nn;            // L_A^(1/2) V_A^H          (column-major, like all)
    double2* M = RA + nn;             // R_A R_B' -> U S
    double2* W = M + nn;
    double* lamA = reinterpret_cast<double*>(W + nn);
    double* lamB = lamA + n;
    double* sv = lamB + n;
    // inputs are row-major Hermitian: element (i, j) at i * n + j
    for (int i = tid; i < nn; i += blockDim.x) {
        const int r = i % n, c = i / n;                    // column-major position (r, c)
        XA[i] = GA[(size_t)r * n + c];
        XB[i] = GB[(size_t)r * n + c];
    }
    __syncthreads();
    jacobi_onesided(XA, VA, n, &s_rot);
    jacobi_onesided(XB, VB, n, &s_rot);
    for (int j = tid; j < n; j += blockDim.x) {            // eigenvalues = column norms
        double a = 0, b = 0;
        for (int i = 0; i < n; ++i) {
            const double2 x = XA[(size_t)j * n + i], y = XB[(size_t)j * n + i];
            a += x.x * x.x + x.y * x.y;
            b += y.x * y.x + y.y * y.y;
        }
        lamA[j] = sqrt(a); lamB[j] = sqrt(b);
    }
    __syncthreads();
    // R_A(i, j) = sqrt(lamA_i) conj(V_A(j, i));  R_B'(i, j) = V_B(i, j) sqrt(lamB_j) (reuse XB)
    for (int i = tid; i < nn; i += blockDim.x) {
        const int r = i % n, c = i / n;
        const double2 v = VA[(size_t)r * n + c];           // V_A(c, r)
        const double sa = sqrt(lamA[r]);
        RA[i] = make_double2(sa * v.x, -sa * v.y);
    }
    __syncthreads();
    for (int i = tid; i < nn; i += blockDim.x) {
        const int c = i / n;
        const double sb = sqrt(lamB[c]);
        const double2 v = VB[i];
        XB[i] = make_double2(sb * v.x, sb * v.y);          // R_B'
    }
    __syncthreads();
    for (int i = tid; i < nn; i += blockDim.x) {           // M = R_A R_B'
        const int r = i % n, c = i / n;
        double2 acc = make_double2(0.0, 0.0);
        for (int l = 0; l < n; ++l) cfma(acc, RA[(size_t)l * n + r], XB[(size_t)c * n + l]);
        M[i] = acc;
    }
    __syncthreads();
    jacobi_onesided(M, W, n, &s_rot);
    if (tid == 0) s_max = 0.0;
    __syncthreads();
    for (int j = tid; j < n; j += blockDim.x) {
        double a = 0;
        for (int i = 0; i < n; ++i) { const double2 x = M[(size_t)j * n + i]; a += x.x * x.x + x.y * x.y; }
        sv[j] = sqrt(a);
    }
    __syncthreads();
    if (tid == 0) {
        double mx = 0.0;
        for (int j = 0; j < n; ++j) mx = fmax(mx, sv[j]);
        s_max = mx;
        *k_out = 0;
    }
    __syncthreads();
    const double thr = cutoff * s_max;
    const int cap = (max_bond > 0 && max_bond < n) ? max_bond : n;
    // rank of every singular value (descending, ties by index); the kept ones are the
    // first `k` ranks.  lamA is reused as the rank table.
    for (int j = tid; j < n; j += blockDim.x) {
        int rank = 0;
        for (int i = 0; i < n; ++i) rank += (sv[i] > sv[j]) || (sv[i] == sv[j] && i < j);
        const bool keep = rank < cap && sv[j] > thr && sv[j] > 0.0;
        lamB[j] = keep || rank == 0 ? (double)rank : -1.0;  // always keep the largest
        if (keep || rank == 0) { atomicAdd(k_out, 1); sv_out[rank] = sv[j]; }
    }
    __syncthreads();
    // P_A(i, rank) = sum_l R_B'(i, l) W(l, j) / sqrt(s_j)          row-major [n][n]
    // P_B(rank, i) = sum_l conj(US(l, j)) R_A(l, i) / (s_j sqrt(s_j))   row-major [n][n]
    for (int idx = tid; idx < nn; idx += blockDim.x) {
        const int i = idx % n, j = idx / n;
        if (lamB[j] < 0.0) continue;
        const int rank = (int)lamB[j];
        const double s = sv[j];
        if (!(s > 0.0)) {                                  // a zero bond matrix: zero projectors
            PA[(size_t)i * n + rank] = make_double2(0.0, 0.0);
            PB[(size_t)rank * n + i] = make_double2(0.0, 0.0);
            continue;
        }
        const double rs = 1.0 / sqrt(s);
        double2 pa = make_double2(0.0, 0.0), pb = make_double2(0.0, 0.0);
        for (int l = 0; l < n; ++l) {
            cfma(pa, XB[(size_t)l * n + i], W[(size_t)j * n + l]);
            const double2 g = cmulc(M[(size_t)j * n + l], RA[(size_t)i * n + l]);
            pb.x += g.x; pb.y += g.y;
        }
        PA[(size_t)i * n + rank] = make_double2(pa.x * rs, pa.y * rs);
        PB[(size_t)rank * n + i] = make_double2(pb.x * rs / s, pb.y * rs / s);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
conj_kernel(long long n, const typename cplx<T>::type* __restrict__ x,
            typename cplx<T>::type* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        auto v = x[i];
        v.y = -v.y;
        y[i] = v;
    }
}

}  // namespace b2

extern "C" int b2_conj(int64_t n, const void* x_dev, void* y_dev, int32_t dtype, void* stream) {
    if (n <= 0) return 0;
    const int sms = b2::num_sms();
    if (sms <= 0) return 1;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == B2_C128)
        b2::conj_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(n, (const double2*)x_dev, (double2*)y_dev);
    else if (dtype == B2_C64)
        b2::conj_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(n, (const float2*)x_dev, (float2*)y_dev);
    else { b2::set_error("b2_conj: bad dtype"); return 1; }
    return b2::check_launch("b2_conj");
}

extern "C" int64_t b2_bond_compress_scratch_bytes(int32_t bond) {
    return (int64_t)7 * bond * bond * 16 + (int64_t)3 * bond * 8;
}

extern "C" int b2_bond_compress(const void* ga_dev, const void* gb_dev, int32_t bond,
                                int32_t max_bond, double cutoff, void* pa_dev, void* pb_dev,
                                int32_t* k_dev, double* sv_dev, void* scratch_dev, void* stream) {
    using namespace b2;
    if (bond <= 0 || bond > BC_MAX_BOND) {
        set_error("b2_bond_compress: bond dimension %d (supported: 1..%d)", bond, BC_MAX_BOND);
        return 1;
    }
    if (!ga_dev || !gb_dev || !pa_dev || !pb_dev || !k_dev || !sv_dev || !scratch_dev) {
        set_error("b2_bond_compress: null pointer");
        return 1;
    }
    if (num_sms() <= 0) return 1;
    bond_compress_kernel<<<1, BC_THREADS, 0, (cudaStream_t)stream>>>(
        (const double2*)ga_dev, (const double2*)gb_dev, bond, max_bond, cutoff, (double2*)pa_dev,
        (double2*)pb_dev, k_dev, sv_dev, (double2*)scratch_dev);
    return check_launch("b2_bond_compress");
}
