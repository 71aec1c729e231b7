// Piece (2c): the persistent small-step interpreter.  Networks of the reference's
// everyday sizes (cfg1 / cfg2 / cfg5, and the slice-invariant part of cfg4: hundreds to
// thousands of leaves of <= 343 elements, reference Box.truncation arrays
// diagram.py:536-590) are launch-bound: almost every tree node is a pair of a few dozen
// elements.  Instead of one launch per node (quimb/cotengra: one numpy call per node,
// backends.py:539-545) ALL early tiny nodes of an eval -- or of a batch of evals -- run
// in ONE launch:
//
//  * the host cuts them into tasks (chains of steps; optyx_b200/planner.py
//    _build_small_program) and tabulates, per step, the operand offsets of every
//    output index and every reduction index; everything a task needs to know is one
//    blob of words;
//  * persistent CTAs take (eval, task) work items from a ticket counter in
//    topological order, copy the task's blob into shared memory (one coalesced pass:
//    records and, for small tasks, the offset tables -- a step never starts with a
//    global-memory round trip), spin (acquire) on the done-flags of the tasks they
//    depend on, import leaves / hand-overs warp-parallel, run the chain with one block
//    barrier per step, export what leaves the task and publish their flag (release);
//  * a dependency always holds a smaller ticket, so its holder is already running:
//    the scheme cannot deadlock and needs no co-residency of the grid;
//  * flags hold the launch epoch: CUDA-graph replays need no reset (the last CTA
//    to finish bumps the epoch and rewinds the ticket counter).
#include "b2_common.cuh"

namespace b2 {

constexpr int SP_MAX_THREADS = 1024;               // blockDim is chosen per launch (256 .. 1024)
constexpr int SP_SMEM_MAX_BYTES = 226 * 1024;
constexpr int SP_HDR = 8, SP_COPY_WORDS = 8, SP_STEP_WORDS = 8;

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename C>
__device__ __forceinline__ C shfl_xor_c(C v, int m) {
    v.x = __shfl_xor_sync(0xffffffffu, v.x, m);
    v.y = __shfl_xor_sync(0xffffffffu, v.y, m);
    return v;
}

// one step; OT / RT point at its tables (shared memory when staged, else global)
template <typename T, typename C>
__device__ __forceinline__ void sp_step(const uint32_t* __restrict__ st, C* __restrict__ S,
                                        const uint2* __restrict__ ot,
                                        const uint32_t* __restrict__ rt, int tid) {
    const int SP_THREADS = (int)blockDim.x;
    const int n_out = (int)st[0], n_red = (int)st[1], rs = (int)st[5];
    const C* A = S + st[2];
    const C* B = S + st[3];
    C* Cc = S + st[4];
    if (rs == 1) {
        // two outputs per thread in flight: independent FMA chains
        int o = tid;
        for (; o + SP_THREADS < n_out; o += 2 * SP_THREADS) {
            const uint2 t0 = ot[o], t1 = ot[o + SP_THREADS];
            const C* a0 = A + (t0.x & 0xffffu); const C* b0 = B + (t0.x >> 16);
            const C* a1 = A + (t1.x & 0xffffu); const C* b1 = B + (t1.x >> 16);
            C acc0 = make_c<T>((T)0, (T)0), acc1 = acc0;
            #pragma unroll 2
            for (int r = 0; r < n_red; ++r) {
                const uint32_t q = rt[r];
                const int ra = q & 0xffffu, rb = q >> 16;
                cfma(acc0, a0[ra], b0[rb]);
                cfma(acc1, a1[ra], b1[rb]);
            }
            Cc[t0.y] = acc0;
            Cc[t1.y] = acc1;
        }
        if (o < n_out) {
            const uint2 tt = ot[o];
            const C* a = A + (tt.x & 0xffffu);
            const C* b = B + (tt.x >> 16);
            C acc0 = make_c<T>((T)0, (T)0), acc1 = acc0;
            int r = 0;
            for (; r + 1 < n_red; r += 2) {        // two partial sums: shorter dependent chain
                const uint32_t q0 = rt[r], q1 = rt[r + 1];
                cfma(acc0, a[q0 & 0xffffu], b[q0 >> 16]);
                cfma(acc1, a[q1 & 0xffffu], b[q1 >> 16]);
            }
            if (r < n_red) { const uint32_t q = rt[r]; cfma(acc0, a[q & 0xffffu], b[q >> 16]); }
            acc0.x += acc1.x; acc0.y += acc1.y;
            Cc[tt.y] = acc0;
        }
    } else {
        // few outputs, longer reduction: rs lanes (a power of two, inside one warp)
        // split the reduction of one output and butterfly-add
        const int part = tid & (rs - 1), per_round = SP_THREADS / rs;
        for (int base = 0; base < n_out; base += per_round) {
            const int o = base + tid / rs;
            const bool valid = o < n_out;
            const uint2 tt = valid ? ot[o] : make_uint2(0u, 0u);
            const C* a = A + (tt.x & 0xffffu);
            const C* b = B + (tt.x >> 16);
            C acc = make_c<T>((T)0, (T)0);
            if (valid)
                for (int r = part; r < n_red; r += rs) {
                    const uint32_t q = rt[r];
                    cfma(acc, a[q & 0xffffu], b[q >> 16]);
                }
            for (int m = rs >> 1; m > 0; m >>= 1) {
                const C other = shfl_xor_c(acc, m);
                acc.x += other.x; acc.y += other.y;
            }
            if (valid && part == 0) Cc[tt.y] = acc;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(SP_MAX_THREADS)
small_program_kernel(const __grid_constant__ b2_small_program P) {
    using C = typename cplx<T>::type;
    extern __shared__ __align__(16) unsigned char sp_raw[];
    uint32_t* H = reinterpret_cast<uint32_t*>(sp_raw);          // the task's blob
    __shared__ int s_item;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int SP_THREADS = (int)blockDim.x, SP_WARPS = SP_THREADS >> 5;
    int* sync = P.sync_dev;
    int* flags = sync + 4;
    const int n_tasks = P.n_tasks;
    const long long n_items = (long long)n_tasks * P.batch;
    // every CTA reads the epoch before any CTA can be the last one to finish
    const int epoch = ld_acquire(sync + 2) + 1;

    for (;;) {
        if (tid == 0) s_item = atomicAdd(sync + 0, 1);
        __syncthreads();
        const long long item = s_item;
        if (item >= n_items) break;
        const int ev = (int)(item / n_tasks), t = (int)(item - (long long)ev * n_tasks);
        const int b0 = __ldg(P.blob_off_dev + t), words = __ldg(P.blob_off_dev + t + 1) - b0;
        {   // the blob: 16-byte vectors (blob offsets are multiples of 4 words by construction
            // of the host encoder only when aligned; fall back to words otherwise)
            const uint32_t* src = P.blob_dev + b0;
            for (int i = tid; i < words; i += SP_THREADS) H[i] = __ldg(src + i);
        }
        __syncthreads();                    // blob visible (and everyone has read s_item)
        const int n_dep = (int)H[0], n_imp = (int)H[1], n_exp = (int)H[2], n_steps = (int)H[3];
        const bool staged = H[4] != 0;
        const uint32_t* deps = H + SP_HDR;
        const uint32_t* copies = deps + n_dep + (n_dep & 1);
        const uint32_t* steps = copies + SP_COPY_WORDS * (n_imp + n_exp);
        C* S = reinterpret_cast<C*>(sp_raw + (((size_t)words * 4 + 15) & ~(size_t)15));
        // ---- wait for the tasks this one imports from (same eval)
        for (int d = tid; d < n_dep; d += SP_THREADS) {
            const int* f = flags + (long long)ev * n_tasks + (int)deps[d];
            while (ld_acquire(f) != epoch) __nanosleep(32);
        }
        if (n_dep) __syncthreads();
        // ---- imports, one warp per record: leaves from the arena, hand-overs from the
        //      workspace (written by another SM during this launch: L2, .cg)
        for (int c = warp; c < n_imp; c += SP_WARPS) {
            const uint32_t* cp = copies + SP_COPY_WORDS * c;
            const long long off = (long long)(((unsigned long long)cp[5] << 32) | cp[4]) +
                                  (long long)ev * (long long)(((unsigned long long)cp[7] << 32) | cp[6]);
            const C* src = reinterpret_cast<const C*>(P.base_dev[cp[0]]) + off;
            C* dst = S + cp[1];
            const int n = (int)cp[2];
            for (int i = lane; i < n; i += 32) dst[i] = __ldcg(src + i);
        }
        __syncthreads();
        // ---- the chain: one block barrier per step
        if (staged) {
            for (int s = 0; s < n_steps; ++s) {
                const uint32_t* st = steps + SP_STEP_WORDS * s;
                sp_step<T, C>(st, S, reinterpret_cast<const uint2*>(H + st[6]), H + st[7], tid);
                __syncthreads();
            }
        } else {
            for (int s = 0; s < n_steps; ++s) {
                const uint32_t* st = steps + SP_STEP_WORDS * s;
                sp_step<T, C>(st, S, reinterpret_cast<const uint2*>(P.out_tab_dev) + st[6],
                              P.red_tab_dev + st[7], tid);
                __syncthreads();
            }
        }
        // ---- exports (warp per record), then publish
        for (int c = n_imp + warp; c < n_imp + n_exp; c += SP_WARPS) {
            const uint32_t* cp = copies + SP_COPY_WORDS * c;
            const long long off = (long long)(((unsigned long long)cp[5] << 32) | cp[4]) +
                                  (long long)ev * (long long)(((unsigned long long)cp[7] << 32) | cp[6]);
            C* dst = reinterpret_cast<C*>(P.base_dev[cp[0]]) + off;
            const C* src = S + cp[1];
            const int n = (int)cp[2];
            for (int i = lane; i < n; i += 32) dst[i] = src[i];
        }
        __threadfence();                    // this thread's exports are visible device-wide ...
        __syncthreads();                    // ... all of them, before the flag goes up
        if (tid == 0) st_release(flags + (long long)ev * n_tasks + t, epoch);
    }
    // ---- the last CTA to leave rewinds the ticket counter and bumps the epoch
    if (tid == 0) {
        __threadfence();
        const int done = atomicAdd(sync + 1, 1);
        if (done == (int)gridDim.x - 1) {
            sync[0] = 0;
            sync[1] = 0;
            __threadfence();
            st_release(sync + 2, epoch);
        }
    }
}

}  // namespace b2

extern "C" int b2_run_small_program(const b2_small_program* prog, int32_t dtype, void* stream) {
    using namespace b2;
    if (!prog) { set_error("b2_run_small_program: null program"); return 1; }
    const b2_small_program& P = *prog;
    if (P.n_tasks <= 0 || P.batch <= 0) return 0;
    if (!P.blob_dev || !P.blob_off_dev || !P.out_tab_dev || !P.red_tab_dev || !P.sync_dev) {
        set_error("b2_run_small_program: null table pointer");
        return 1;
    }
    if (dtype != B2_C128 && dtype != B2_C64) { set_error("b2_run_small_program: bad dtype %d", dtype); return 1; }
    const size_t smem = (size_t)P.smem_bytes;
    if (P.smem_bytes < 0 || smem > (size_t)SP_SMEM_MAX_BYTES) {
        set_error("b2_run_small_program: %zu bytes of shared memory (max %d)", smem, SP_SMEM_MAX_BYTES);
        return 1;
    }
    const int sms = num_sms();
    if (sms <= 0) return 1;
    static once_flags attr_done;
    if (!is_done(attr_done)) {
        cudaError_t e1 = cudaFuncSetAttribute(small_program_kernel<double>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              SP_SMEM_MAX_BYTES);
        cudaError_t e2 = cudaFuncSetAttribute(small_program_kernel<float>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              SP_SMEM_MAX_BYTES);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_error("b2_run_small_program: smem attr: %s",
                      cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            return 1;
        }
        mark_done(attr_done);
    }
    // threads per CTA: shared memory decides how many CTAs fit an SM; with few CTAs
    // per SM the CTA itself must bring the warps that hide the latencies (>= 32 per SM)
    int threads = 256, occ = 0;
    for (;;) {
        cudaError_t e = dtype == B2_C128
            ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, small_program_kernel<double>, threads, smem)
            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, small_program_kernel<float>, threads, smem);
        if (e != cudaSuccess || occ < 1) occ = 1;
        if (occ * threads >= 1024 || threads >= SP_MAX_THREADS) break;
        threads *= 2;
    }
    if (occ > 8) occ = 8;
    long long grid = (long long)P.n_tasks * P.batch;
    if (grid > (long long)sms * occ) grid = (long long)sms * occ;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == B2_C128)
        small_program_kernel<double><<<(unsigned)grid, threads, smem, s>>>(P);
    else
        small_program_kernel<float><<<(unsigned)grid, threads, smem, s>>>(P);
    set_last_kernel("small_program");
    return check_launch("b2_run_small_program");
}
