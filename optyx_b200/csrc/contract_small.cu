// Piece (2c): the small-step interpreter.  Networks of the reference's everyday
// sizes (cfg1 / cfg2 / cfg5: hundreds of leaves of <= 343 elements, reference
// Box.truncation arrays diagram.py:536-590) are launch-bound: most tree nodes are
// pairs of a few dozen elements.  Instead of one launch per node (quimb/cotengra:
// one numpy call per node, backends.py:539-545) every SUBTREE of tiny nodes runs
// inside one CTA, all subtrees in one launch.  A CTA copies its group's leaves
// into shared memory once, then walks the steps in order -- one thread per output
// element, operands and intermediates in shared memory, a block barrier between
// steps -- so a dependent step costs a shared-memory round trip instead of a
// kernel launch plus an L2 round trip.  Only the root of a subtree is written to
// global memory.
#include "b2_common.cuh"

namespace b2 {

constexpr int SM_THREADS = 256, SM_CHUNK = 16, SM_RED_TABLE = 512;
constexpr int SM_SCRATCH_MAX = 160 * 1024;

template <typename T>
__global__ void __launch_bounds__(SM_THREADS)
contract_small_groups_kernel(const b2_small_step* __restrict__ steps,
                             const int* __restrict__ group_start,
                             const b2_small_preload* __restrict__ preload,
                             const int* __restrict__ pre_start) {
    using C = typename cplx<T>::type;
    extern __shared__ __align__(16) unsigned char scratch[];
    __shared__ b2_small_step sts[SM_CHUNK];
    __shared__ int red_a[SM_RED_TABLE], red_b[SM_RED_TABLE];
    const int g = blockIdx.x;
    // ---- leaves of the group -> scratch (16-byte words; leaves are 16 B aligned
    //      for complex128, 8 B for complex64: copy in units of C)
    for (int p = pre_start[g]; p < pre_start[g + 1]; ++p) {
        const b2_small_preload e = preload[p];
        const C* src = reinterpret_cast<const C*>(e.src);
        C* dst = reinterpret_cast<C*>(scratch + e.dst_off);
        const int n = e.nbytes / (int)sizeof(C);
        for (int i = threadIdx.x; i < n; i += SM_THREADS) dst[i] = src[i];
    }
    const int s0 = group_start[g], s1 = group_start[g + 1];
    for (int base = s0; base < s1; base += SM_CHUNK) {
        const int cnt = s1 - base < SM_CHUNK ? s1 - base : SM_CHUNK;
        __syncthreads();                   // previous chunk's descriptors no longer needed
        {
            const int* src = reinterpret_cast<const int*>(steps + base);
            int* dst = reinterpret_cast<int*>(sts);
            const int words = cnt * (int)(sizeof(b2_small_step) / 4);
            for (int i = threadIdx.x; i < words; i += SM_THREADS) dst[i] = src[i];
        }
        __syncthreads();                   // descriptors (and, first time, the leaves) visible
        for (int s = 0; s < cnt; ++s) {
            const b2_small_step& st = sts[s];
            const C* A = reinterpret_cast<const C*>(
                (st.a_smem ? scratch : reinterpret_cast<const unsigned char*>(st.a_ptr)) + st.a_off);
            const C* B = reinterpret_cast<const C*>(
                (st.b_smem ? scratch : reinterpret_cast<const unsigned char*>(st.b_ptr)) + st.b_off);
            C* Cc = reinterpret_cast<C*>(
                (st.c_smem ? scratch : reinterpret_cast<unsigned char*>(st.c_ptr)) + st.c_off);
            const int n_out = st.n_out, n_red = st.n_red;
            // reduction offsets once per step (tiny steps: a few dozen index values)
            const bool red_table = st.n_red_elems <= SM_RED_TABLE;
            if (red_table) {
                for (int r = threadIdx.x; r < st.n_red_elems; r += SM_THREADS) {
                    int ra = 0, rb = 0;
                    unsigned rr = (unsigned)r;
                    for (int m = n_out + n_red - 1; m >= n_out; --m) {
                        const unsigned q = __umulhi(rr, st.magic[m]), x = rr - q * (unsigned)st.dim[m];
                        rr = q;
                        ra += (int)x * (int)st.sa[m];
                        rb += (int)x * (int)st.sb[m];
                    }
                    red_a[r] = ra; red_b[r] = rb;
                }
                __syncthreads();
            }
            for (int o = threadIdx.x; o < st.n_out_elems; o += SM_THREADS) {
                int oa = 0, ob = 0, oc = 0;
                unsigned rem = (unsigned)o;
                for (int m = n_out - 1; m >= 0; --m) {
                    // exact for rem < 2^16 (at most SMALL_OUT outputs): magic = ceil(2^32 / dim)
                    const unsigned q = __umulhi(rem, st.magic[m]), x = rem - q * (unsigned)st.dim[m];
                    rem = q;
                    oa += (int)x * (int)st.sa[m];
                    ob += (int)x * (int)st.sb[m];
                    oc += (int)x * (int)st.sc[m];
                }
                C acc = make_c<T>((T)0, (T)0);
                if (red_table) {
                    for (int r = 0; r < st.n_red_elems; ++r) cfma(acc, A[oa + red_a[r]], B[ob + red_b[r]]);
                } else {
                    for (int r = 0; r < st.n_red_elems; ++r) {
                        int ra = oa, rb = ob;
                        unsigned rr = (unsigned)r;
                        for (int m = n_out + n_red - 1; m >= n_out; --m) {
                            const unsigned q = __umulhi(rr, st.magic[m]), x = rr - q * (unsigned)st.dim[m];
                            rr = q;
                            ra += (int)x * (int)st.sa[m];
                            rb += (int)x * (int)st.sb[m];
                        }
                        cfma(acc, A[ra], B[rb]);
                    }
                }
                Cc[oc] = acc;
            }
            __syncthreads();               // the next step may read what this one wrote
        }
    }
}

}  // namespace b2

extern "C" int b2_contract_small_groups(const b2_small_step* steps_dev,
                                        const int32_t* group_start_dev,
                                        const b2_small_preload* preload_dev,
                                        const int32_t* pre_start_dev, int32_t n_groups,
                                        int32_t scratch_bytes, int32_t dtype, void* stream) {
    using namespace b2;
    if (n_groups <= 0) return 0;
    if (!steps_dev || !group_start_dev || !preload_dev || !pre_start_dev) {
        set_error("b2_contract_small_groups: null pointer");
        return 1;
    }
    if (scratch_bytes < 0 || scratch_bytes > SM_SCRATCH_MAX) {
        set_error("b2_contract_small_groups: scratch of %d bytes (max %d)", scratch_bytes, SM_SCRATCH_MAX);
        return 1;
    }
    if (num_sms() <= 0) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    static b2::once_flags attr_done;
    if (!b2::is_done(attr_done)) {
        cudaError_t e1 = cudaFuncSetAttribute(contract_small_groups_kernel<double>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SCRATCH_MAX);
        cudaError_t e2 = cudaFuncSetAttribute(contract_small_groups_kernel<float>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SCRATCH_MAX);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_error("b2_contract_small_groups: smem attr: %s",
                      cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            return 1;
        }
        b2::mark_done(attr_done);
    }
    if (dtype == B2_C128)
        contract_small_groups_kernel<double><<<(unsigned)n_groups, SM_THREADS, scratch_bytes, s>>>(
            steps_dev, group_start_dev, preload_dev, pre_start_dev);
    else if (dtype == B2_C64)
        contract_small_groups_kernel<float><<<(unsigned)n_groups, SM_THREADS, scratch_bytes, s>>>(
            steps_dev, group_start_dev, preload_dev, pre_start_dev);
    else { set_error("b2_contract_small_groups: bad dtype %d", dtype); return 1; }
    set_last_kernel("small_groups");
    return check_launch("b2_contract_small_groups");
}
