// Piece (2a): thin / bandwidth-bound pairwise contraction.
//
//   C[out] (+)= alpha * sum_red A[out|red] * B[out|red]
//
// The index permutation the reference performs as a separate transpose pass
// before BLAS (numpy.tensordot under quimb, reached from
// optyx/core/backends.py:514 / 539-545) is fused into the address computation:
// every mode carries its own stride into A, B and C; stride 0 broadcasts.
//
// Two kernels:
//  * tiled (default): the fastest output modes form a tile of up to 2048
//    elements that is CONTIGUOUS in C.  Their A/B offsets depend only on the
//    position inside the tile, so they are decoded once per CTA into shared
//    memory, as are the offsets of the reduction modes; each persistent CTA
//    then walks tiles (grid = SMs x resident CTAs), decoding only the outer
//    modes per tile.  Per element: 2 LDS + R x (2 LDS + 2 LDG.128 + 4 DFMA) +
//    one coalesced STG.128.
//  * generic (fallback): one thread per output element with a full decode, for
//    outputs that are not contiguous (generalised diagonals) or reductions too
//    long for the table.
#include "b2_common.cuh"

namespace b2 {

int launch_fma(const b2_modes& md, const void* a, const void* b, void* c, double ar, double ai,
               int accumulate, int dtype, cudaStream_t s, bool& handled);

constexpr int kMaxRed = 16;
constexpr int kTileMax = 2048;
constexpr int kRedTableMax = 512;
constexpr int kTiledThreads = 256;

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

// ---- tiled kernel -------------------------------------------------------------
// modes [0, n_outer) are decoded per tile; modes [n_outer, n_out) span the tile
// (contiguous in C, fastest last); modes [n_out, n_out + n_red) are summed.
template <typename T>
__global__ void __launch_bounds__(kTiledThreads)
contract_direct_tiled_kernel(const __grid_constant__ b2_modes md, int n_outer, int tile,
                             int n_red_elems, long long n_tiles,
                             const typename cplx<T>::type* __restrict__ A,
                             const typename cplx<T>::type* __restrict__ B,
                             typename cplx<T>::type* __restrict__ Cout, T alpha_re, T alpha_im,
                             int accumulate) {
    using C = typename cplx<T>::type;
    extern __shared__ long long tab[];
    long long* ia = tab;                 // [tile]
    long long* ib = tab + tile;          // [tile]
    long long* ra = tab + 2 * tile;      // [n_red_elems]
    long long* rb = ra + n_red_elems;    // [n_red_elems]
    const int n_out = md.n_out, n_red = md.n_red;
    for (int e = threadIdx.x; e < tile; e += kTiledThreads) {
        long long oa = 0, ob = 0;
        int rem = e;
        for (int m = n_out - 1; m >= n_outer; --m) {
            const int d = md.dim[m];
            const int q = rem / d;
            const long long i = rem - q * d;
            rem = q;
            oa += i * md.sa[m];
            ob += i * md.sb[m];
        }
        ia[e] = oa; ib[e] = ob;
    }
    for (int r = threadIdx.x; r < n_red_elems; r += kTiledThreads) {
        long long oa = 0, ob = 0;
        int rem = r;
        for (int m = n_out + n_red - 1; m >= n_out; --m) {
            const int d = md.dim[m];
            const int q = rem / d;
            const long long i = rem - q * d;
            rem = q;
            oa += i * md.sa[m];
            ob += i * md.sb[m];
        }
        ra[r] = oa; rb[r] = ob;
    }
    __syncthreads();
    // each CTA owns a CONTIGUOUS range of tiles: one full decode of the outer
    // modes, then an odometer increment per tile (no divisions in the loop)
    const long long per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per;
    const long long t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
    if (t0 >= t1) return;
    int cnt[B2_MAX_MODES];
    long long oa = 0, ob = 0, oc = 0;
    {
        long long rem = t0;
        for (int m = n_outer - 1; m >= 0; --m) {
            const long long d = md.dim[m];
            const long long q = rem / d;
            const int i = (int)(rem - q * d);
            rem = q;
            cnt[m] = i;
            oa += i * md.sa[m];
            ob += i * md.sb[m];
            oc += i * md.sc[m];
        }
    }
    for (long long t = t0; t < t1; ++t) {
        if (t != t0) {
            for (int m = n_outer - 1; m >= 0; --m) {
                oa += md.sa[m]; ob += md.sb[m]; oc += md.sc[m];
                if (++cnt[m] < md.dim[m]) break;
                oa -= (long long)md.dim[m] * md.sa[m];
                ob -= (long long)md.dim[m] * md.sb[m];
                oc -= (long long)md.dim[m] * md.sc[m];
                cnt[m] = 0;
            }
        }
        const C* __restrict__ At = A + oa;
        const C* __restrict__ Bt = B + ob;
        C* __restrict__ Ct = Cout + oc;
        #pragma unroll 2
        for (int e = threadIdx.x; e < tile; e += kTiledThreads) {
            const C* a = At + ia[e];
            const C* b = Bt + ib[e];
            C acc = make_c<T>((T)0, (T)0);
            if (n_red_elems == 1) {
                cfma(acc, a[0], b[0]);
            } else {
                #pragma unroll 4
                for (int r = 0; r < n_red_elems; ++r) cfma(acc, a[ra[r]], b[rb[r]]);
            }
            C res;
            res.x = alpha_re * acc.x - alpha_im * acc.y;
            res.y = alpha_re * acc.y + alpha_im * acc.x;
            if (accumulate) { const C old = Ct[e]; res.x += old.x; res.y += old.y; }
            Ct[e] = res;
        }
    }
}

// Split the output modes of `md` into outer/inner for the tiled kernel.  The
// inner modes are the longest C-contiguous tail whose extent fits kTileMax; one
// mode may be split (dim = q * f) so that the tile is as large as possible.
// Returns false when the output is not contiguous in its fastest mode.
static bool make_tiled(const b2_modes& md, long long n_out_elems, int sms, b2_modes& out,
                       int& n_outer, int& tile) {
    const int no = md.n_out;
    if (no == 0) return false;
    if (md.sc[no - 1] != 1) return false;
    // keep tiles small enough to give every SM a few of them
    long long cap = kTileMax;
    while (cap > 256 && n_out_elems / cap < (long long)sms * 4) cap >>= 1;
    long long ext = 1;
    int first_inner = no;             // modes [first_inner, no) fully inside the tile
    int split_mode = -1, split_f = 1;
    for (int m = no - 1; m >= 0; --m) {
        if (md.sc[m] != ext) break;   // not nested-contiguous in C any more
        const long long d = md.dim[m];
        if (ext * d <= cap) { ext *= d; first_inner = m; continue; }
        // partial: largest divisor f of d with ext * f <= cap
        int best = 1;
        for (int f = 2; (long long)f * ext <= cap && f <= d; ++f)
            if (d % f == 0) best = f;
        if (best > 1) { split_mode = m; split_f = best; ext *= best; }
        break;
    }
    if (ext < 32) return false;       // tile too small to be worth the tables
    if (split_mode >= 0 && md.n_out + md.n_red + 1 > B2_MAX_MODES) return false;
    out = md;
    if (split_mode >= 0) {
        // shift modes after split_mode by one; insert the inner part
        for (int m = md.n_out + md.n_red; m > split_mode + 1; --m) {
            out.dim[m] = md.dim[m - 1]; out.sa[m] = md.sa[m - 1];
            out.sb[m] = md.sb[m - 1]; out.sc[m] = md.sc[m - 1];
        }
        const int d = md.dim[split_mode];
        out.dim[split_mode] = d / split_f;
        out.sa[split_mode] = md.sa[split_mode] * split_f;
        out.sb[split_mode] = md.sb[split_mode] * split_f;
        out.sc[split_mode] = md.sc[split_mode] * split_f;
        out.dim[split_mode + 1] = split_f;
        out.sa[split_mode + 1] = md.sa[split_mode];
        out.sb[split_mode + 1] = md.sb[split_mode];
        out.sc[split_mode + 1] = md.sc[split_mode];
        out.n_out = md.n_out + 1;
        n_outer = split_mode + 1;
    } else {
        n_outer = first_inner;
    }
    tile = (int)ext;
    return true;
}


// ---- split-K reduction kernel ---------------------------------------------------
// Few outputs, huge reduction (e.g. the last pair of a photonic network: a 7 x 7
// result summed over 4 million index values).  Each CTA owns a chunk of the
// reduction range, decodes its A/B offsets once into shared memory, then every
// warp sums its outputs over the chunk (lanes stride the chunk, shuffle
// reduce) and adds the partial into C with two FP64 atomics.
constexpr int kRedChunk = 1024;

template <typename T>
__global__ void __launch_bounds__(256)
contract_reduce_kernel(const __grid_constant__ b2_modes md, int n_out_elems, long long n_red_elems,
                       const typename cplx<T>::type* __restrict__ A,
                       const typename cplx<T>::type* __restrict__ B,
                       typename cplx<T>::type* __restrict__ Cout, T alpha_re, T alpha_im) {
    using C = typename cplx<T>::type;
    __shared__ long long ra[kRedChunk], rb[kRedChunk];
    const int n_out = md.n_out, n_red = md.n_red;
    const long long r0 = (long long)blockIdx.x * kRedChunk;
    const int nr = (int)((n_red_elems - r0) < kRedChunk ? (n_red_elems - r0) : kRedChunk);
    for (int i = threadIdx.x; i < nr; i += blockDim.x) {
        long long rem = r0 + i, oa = 0, ob = 0;
        for (int m = n_out + n_red - 1; m >= n_out; --m) {
            const long long d = md.dim[m];
            const long long q = rem / d, x = rem - q * d;
            rem = q;
            oa += x * md.sa[m];
            ob += x * md.sb[m];
        }
        ra[i] = oa; rb[i] = ob;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int o = warp; o < n_out_elems; o += nwarps) {
        long long oa = 0, ob = 0, oc = 0;
        int rem = o;
        for (int m = n_out - 1; m >= 0; --m) {
            const int d = md.dim[m];
            const int q = rem / d, x = rem - q * d;
            rem = q;
            oa += (long long)x * md.sa[m];
            ob += (long long)x * md.sb[m];
            oc += (long long)x * md.sc[m];
        }
        C acc = make_c<T>((T)0, (T)0);
        #pragma unroll 4
        for (int i = lane; i < nr; i += 32) cfma(acc, A[oa + ra[i]], B[ob + rb[i]]);
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        if (lane == 0) {
            T* dst = reinterpret_cast<T*>(Cout + oc);
            atomicAdd(dst, alpha_re * acc.x - alpha_im * acc.y);
            atomicAdd(dst + 1, alpha_re * acc.y + alpha_im * acc.x);
        }
    }
}

template <typename T>
__global__ void zero_strided_kernel(const __grid_constant__ b2_modes md, int n_out_elems,
                                    typename cplx<T>::type* __restrict__ Cout) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_out_elems;
         o += gridDim.x * blockDim.x) {
        long long oc = 0;
        int rem = o;
        for (int m = md.n_out - 1; m >= 0; --m) {
            const int d = md.dim[m];
            const int q = rem / d, x = rem - q * d;
            rem = q;
            oc += (long long)x * md.sc[m];
        }
        Cout[oc] = make_c<T>((T)0, (T)0);
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
    int sms = num_sms();
    if (sms <= 0) return 1;

    if (n_out_elems <= 4096 && n_red_elems >= 4096 && n_out_elems * 8 <= n_red_elems) {
        // reduction-dominated pair: parallelise over the reduction range
        const long long chunks = (n_red_elems + kRedChunk - 1) / kRedChunk;
        if (chunks <= 2147483647ll) {
            if (!accumulate)
                zero_strided_kernel<T><<<(unsigned)((n_out_elems + 255) / 256), 256, 0, s>>>(
                    *md, (int)n_out_elems, (C*)c);
            contract_reduce_kernel<T><<<(unsigned)chunks, 256, 0, s>>>(
                *md, (int)n_out_elems, n_red_elems, (const C*)a, (const C*)b, (C*)c, (T)ar, (T)ai);
            set_last_kernel("reduce");
            return check_launch("b2_contract_direct(reduce)");
        }
    }
    // element-wise pair (every output mode lives in BOTH operands, nothing is reduced: a
    // product over shared hyper-indices): there is no small operand to stage, the fma kernel's
    // B block would be as large as the tile -- measured 502 GB/s on cfg5a's (6.7 M, 1, 1, 1)
    // step.  The tiled streaming kernel below reads A, B and writes C once, coalesced.
    bool elementwise = n_red_elems == 1;
    for (int m = 0; m < md->n_out && elementwise; ++m)
        if (md->dim[m] > 1 && (md->sa[m] == 0 || md->sb[m] == 0)) elementwise = false;
    if (!elementwise) {   // primary path: C-tiled, B block in shared memory, A rows in registers
        bool handled = false;
        const int dt = sizeof(T) == 8 ? B2_C128 : B2_C64;
        int st = launch_fma(*md, a, b, c, ar, ai, accumulate, dt, s, handled);
        if (st != 0 || handled) return st;
    }
    b2_modes tiled;
    int n_outer = 0, tile = 0;
    if (n_red_elems <= kRedTableMax && make_tiled(*md, n_out_elems, sms, tiled, n_outer, tile)) {
        const long long n_tiles = n_out_elems / tile;
        const size_t smem = (size_t)(2 * tile + 2 * n_red_elems) * sizeof(long long);
        long long blocks = n_tiles;
        const long long cap = (long long)sms * 6;
        if (blocks > cap) blocks = cap;
        contract_direct_tiled_kernel<T><<<(unsigned)blocks, kTiledThreads, smem, s>>>(
            tiled, n_outer, tile, (int)n_red_elems, n_tiles, (const C*)a, (const C*)b, (C*)c,
            (T)ar, (T)ai, accumulate);
        set_last_kernel("tiled");
        return check_launch("b2_contract_direct(tiled)");
    }

    set_last_kernel("generic");
    const int threads = 256;
    long long blocks = (n_out_elems + threads - 1) / threads;
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
