// Piece (2a): thin / bandwidth-bound pairwise contraction.
//
//   C[out] (+)= alpha * sum_red A[out|red] * B[out|red]
//
// One thread per output element (grid-stride, grid sized in multiples of the
// SM count).  The index permutation the reference performs as a separate
// transpose pass before BLAS (numpy.tensordot under quimb, reached from
// optyx/core/backends.py:514 / 539-545) is fused into the address computation:
// every mode carries its own stride into A, B and C; stride 0 broadcasts.  The
// host merges stride-compatible modes first, so a typical step decodes 2-6
// modes per element.  C is written in its memory order (coalesced 16 B stores);
// A/B reads are coalesced whenever the fastest output mode is also fast in the
// operand, otherwise they are served by L1/L2.
#include "b2_common.cuh"

namespace b2 {

constexpr int kMaxRed = 16;

template <typename T, typename I>
__global__ void __launch_bounds__(256)
contract_direct_kernel(const __grid_constant__ b2_modes md,
                       const typename cplx<T>::type* __restrict__ A,
                       const typename cplx<T>::type* __restrict__ B,
                       typename cplx<T>::type* __restrict__ Cout, I n_out_elems,
                       long long n_red_elems, T alpha_re, T alpha_im, int accumulate) {
    using C = typename cplx<T>::type;
    const int n_out = md.n_out, n_red = md.n_red;
    for (I e = (I)blockIdx.x * blockDim.x + threadIdx.x; e < n_out_elems;
         e += (I)gridDim.x * blockDim.x) {
        long long oa = 0, ob = 0, oc = 0;
        I rem = e;
        for (int m = n_out - 1; m >= 0; --m) {
            const I d = (I)md.dim[m];
            const I q = rem / d;
            const long long i = (long long)(rem - q * d);
            rem = q;
            oa += i * md.sa[m];
            ob += i * md.sb[m];
            oc += i * md.sc[m];
        }
        C acc = make_c<T>((T)0, (T)0);
        if (n_red == 0) {
            cfma(acc, A[oa], B[ob]);
        } else if (n_red == 1) {
            const long long sa = md.sa[n_out], sb = md.sb[n_out];
            const int d = md.dim[n_out];
            #pragma unroll 4
            for (int r = 0; r < d; ++r) cfma(acc, A[oa + r * sa], B[ob + r * sb]);
        } else {
            int cnt[kMaxRed];
            #pragma unroll
            for (int m = 0; m < kMaxRed; ++m) cnt[m] = 0;
            long long ra = 0, rb = 0;
            for (long long r = 0; r < n_red_elems; ++r) {
                cfma(acc, A[oa + ra], B[ob + rb]);
                #pragma unroll
                for (int m = kMaxRed - 1; m >= 0; --m) {
                    if (m < n_red) {
                        const int g = n_out + m;
                        ra += md.sa[g]; rb += md.sb[g];
                        if (++cnt[m] < md.dim[g]) break;
                        ra -= (long long)md.dim[g] * md.sa[g];
                        rb -= (long long)md.dim[g] * md.sb[g];
                        cnt[m] = 0;
                    }
                }
            }
        }
        C res;
        res.x = alpha_re * acc.x - alpha_im * acc.y;
        res.y = alpha_re * acc.y + alpha_im * acc.x;
        if (accumulate) { C old = Cout[oc]; res.x += old.x; res.y += old.y; }
        Cout[oc] = res;
    }
}

template <typename T>
int launch_direct(const b2_modes* md, const void* a, const void* b, void* c, double ar,
                  double ai, int accumulate, cudaStream_t s) {
    using C = typename cplx<T>::type;
    long long n_out_elems = 1, n_red_elems = 1;
    for (int m = 0; m < md->n_out; ++m) n_out_elems *= md->dim[m];
    for (int m = 0; m < md->n_red; ++m) n_red_elems *= md->dim[md->n_out + m];
    if (n_out_elems == 0) return 0;
    const int threads = 256;
    long long blocks = (n_out_elems + threads - 1) / threads;
    int sms = num_sms();
    if (sms <= 0) return 1;
    long long cap = (long long)sms * 16;
    if (blocks > cap) blocks = cap;
    if (n_out_elems < (1ll << 31))
        contract_direct_kernel<T, unsigned><<<(unsigned)blocks, threads, 0, s>>>(
            *md, (const C*)a, (const C*)b, (C*)c, (unsigned)n_out_elems, n_red_elems, (T)ar,
            (T)ai, accumulate);
    else
        contract_direct_kernel<T, unsigned long long><<<(unsigned)blocks, threads, 0, s>>>(
            *md, (const C*)a, (const C*)b, (C*)c, (unsigned long long)n_out_elems,
            n_red_elems, (T)ar, (T)ai, accumulate);
    return check_launch("b2_contract_direct");
}

}  // namespace b2

extern "C" int b2_contract_direct(const b2_modes* modes, const void* a_dev, const void* b_dev,
                                  void* c_dev, double alpha_re, double alpha_im,
                                  int32_t accumulate, int32_t dtype, void* stream) {
    if (!modes || !a_dev || !b_dev || !c_dev) {
        b2::set_error("b2_contract_direct: null pointer");
        return 1;
    }
    if (modes->n_out < 0 || modes->n_red < 0 || modes->n_out + modes->n_red > B2_MAX_MODES ||
        modes->n_red > b2::kMaxRed) {
        b2::set_error("b2_contract_direct: bad mode counts out=%d red=%d", modes->n_out,
                      modes->n_red);
        return 1;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == B2_C128)
        return b2::launch_direct<double>(modes, a_dev, b_dev, c_dev, alpha_re, alpha_im,
                                         accumulate, s);
    if (dtype == B2_C64)
        return b2::launch_direct<float>(modes, a_dev, b_dev, c_dev, alpha_re, alpha_im,
                                        accumulate, s);
    b2::set_error("b2_contract_direct: bad dtype %d", dtype);
    return 1;
}
