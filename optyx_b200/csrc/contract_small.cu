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
//    output index and every reduction index;
//  * persistent CTAs take (eval, task) work items from a ticket counter in
//    topological order, spin (acquire) on the done-flags of the tasks they depend
//    on, import leaves / hand-overs into shared memory, run the chain with one block
//    barrier per step, export what leaves the task and publish their flag (release);
//  * a dependency always holds a smaller ticket, so its holder is already running:
//    the scheme cannot deadlock and needs no co-residency of the grid;
//  * flags hold the launch epoch: CUDA-graph replays need no reset (the last CTA
//    to finish bumps the epoch and rewinds the ticket counter).
#include "b2_common.cuh"

namespace b2 {

constexpr int SP_THREADS = 256;
constexpr int SP_CHUNK = 32;                       // step descriptors staged in shared memory
constexpr int SP_SCRATCH_MAX_BYTES = 220 * 1024;

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

template <typename T>
__global__ void __launch_bounds__(SP_THREADS)
small_program_kernel(const __grid_constant__ b2_small_program P) {
    using C = typename cplx<T>::type;
    extern __shared__ __align__(16) unsigned char sp_raw[];
    C* S = reinterpret_cast<C*>(sp_raw);
    __shared__ int s_item;
    __shared__ b2_sp_step s_steps[SP_CHUNK];
    const int tid = threadIdx.x;
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
        const b2_sp_task task = P.tasks_dev[t];
        // ---- wait for the tasks this one imports from (same eval)
        for (int d = task.dep_begin + tid; d < task.dep_end; d += SP_THREADS) {
            const int* f = flags + (long long)ev * n_tasks + P.deps_dev[d];
            while (ld_acquire(f) != epoch) __nanosleep(40);
        }
        __syncthreads();                    // (also: everyone has read s_item)
        // ---- imports: leaves from the arena, hand-overs from the workspace (L2: .cg)
        for (int c = task.imp_begin; c < task.imp_end; ++c) {
            const b2_sp_copy cp = P.copies_dev[c];
            const C* src = reinterpret_cast<const C*>(P.base_dev[cp.space]) + cp.glob_off +
                           (long long)ev * cp.eval_stride;
            C* dst = S + cp.smem_off;
            for (int i = tid; i < cp.n; i += SP_THREADS) dst[i] = __ldcg(src + i);
        }
        __syncthreads();
        // ---- the chain (descriptors staged SP_CHUNK at a time: a step must not start
        //      with a global-memory round trip)
        for (int s = task.step_begin; s < task.step_end; ++s) {
            const int slot = (s - task.step_begin) % SP_CHUNK;
            if (slot == 0) {
                const int cnt = min(SP_CHUNK, task.step_end - s);
                const int* src = reinterpret_cast<const int*>(P.steps_dev + s);
                int* dst = reinterpret_cast<int*>(s_steps);
                for (int i = tid; i < cnt * (int)(sizeof(b2_sp_step) / 4); i += SP_THREADS)
                    dst[i] = __ldg(src + i);
                __syncthreads();
            }
            const b2_sp_step st = s_steps[slot];
            const C* A = S + st.a_off;
            const C* B = S + st.b_off;
            C* Cc = S + st.c_off;
            const uint2* ot = reinterpret_cast<const uint2*>(P.out_tab_dev) + st.out_tab;
            const uint32_t* rt = P.red_tab_dev + st.red_tab;
            const int n_out = st.n_out, n_red = st.n_red, rs = st.rsplit;
            if (rs == 1) {
                for (int o = tid; o < n_out; o += SP_THREADS) {
                    const uint2 tt = __ldg(ot + o);
                    const C* a = A + (tt.x & 0xffffu);
                    const C* b = B + (tt.x >> 16);
                    C acc = make_c<T>((T)0, (T)0);
                    #pragma unroll 4
                    for (int r = 0; r < n_red; ++r) {
                        const uint32_t q = __ldg(rt + r);
                        cfma(acc, a[q & 0xffffu], b[q >> 16]);
                    }
                    Cc[tt.y] = acc;
                }
            } else {
                // few outputs, longer reduction: rs lanes (a power of two, inside one
                // warp) split the reduction of one output and butterfly-add
                const int part = tid & (rs - 1), per_round = SP_THREADS / rs;
                for (int base = 0; base < n_out; base += per_round) {
                    const int o = base + tid / rs;
                    const bool valid = o < n_out;
                    const uint2 tt = valid ? __ldg(ot + o) : make_uint2(0u, 0u);
                    const C* a = A + (tt.x & 0xffffu);
                    const C* b = B + (tt.x >> 16);
                    C acc = make_c<T>((T)0, (T)0);
                    if (valid)
                        for (int r = part; r < n_red; r += rs) {
                            const uint32_t q = __ldg(rt + r);
                            cfma(acc, a[q & 0xffffu], b[q >> 16]);
                        }
                    for (int m = rs >> 1; m > 0; m >>= 1) {
                        const C other = shfl_xor_c(acc, m);
                        acc.x += other.x; acc.y += other.y;
                    }
                    if (valid && part == 0) Cc[tt.y] = acc;
                }
            }
            __syncthreads();                // the next step may read what this one wrote
        }
        // ---- exports, then publish
        for (int c = task.exp_begin; c < task.exp_end; ++c) {
            const b2_sp_copy cp = P.copies_dev[c];
            C* dst = reinterpret_cast<C*>(P.base_dev[cp.space]) + cp.glob_off +
                     (long long)ev * cp.eval_stride;
            const C* src = S + cp.smem_off;
            for (int i = tid; i < cp.n; i += SP_THREADS) dst[i] = src[i];
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
    if (!P.tasks_dev || !P.steps_dev || !P.copies_dev || !P.out_tab_dev || !P.red_tab_dev ||
        !P.sync_dev) {
        set_error("b2_run_small_program: null table pointer");
        return 1;
    }
    const size_t esize = dtype == B2_C128 ? 16 : 8;
    if (dtype != B2_C128 && dtype != B2_C64) { set_error("b2_run_small_program: bad dtype %d", dtype); return 1; }
    const size_t smem = (size_t)P.scratch_elems * esize;
    if (P.scratch_elems < 0 || smem > (size_t)SP_SCRATCH_MAX_BYTES) {
        set_error("b2_run_small_program: scratch of %zu bytes (max %d)", smem, SP_SCRATCH_MAX_BYTES);
        return 1;
    }
    const int sms = num_sms();
    if (sms <= 0) return 1;
    static once_flags attr_done;
    if (!is_done(attr_done)) {
        cudaError_t e1 = cudaFuncSetAttribute(small_program_kernel<double>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              SP_SCRATCH_MAX_BYTES);
        cudaError_t e2 = cudaFuncSetAttribute(small_program_kernel<float>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              SP_SCRATCH_MAX_BYTES);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_error("b2_run_small_program: smem attr: %s",
                      cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            return 1;
        }
        mark_done(attr_done);
    }
    int occ = 0;
    cudaError_t e = dtype == B2_C128
        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, small_program_kernel<double>, SP_THREADS, smem)
        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, small_program_kernel<float>, SP_THREADS, smem);
    if (e != cudaSuccess || occ < 1) occ = 1;
    if (occ > 4) occ = 4;
    long long grid = (long long)P.n_tasks * P.batch;
    if (grid > (long long)sms * occ) grid = (long long)sms * occ;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == B2_C128)
        small_program_kernel<double><<<(unsigned)grid, SP_THREADS, smem, s>>>(P);
    else
        small_program_kernel<float><<<(unsigned)grid, SP_THREADS, smem, s>>>(P);
    set_last_kernel("small_program");
    return check_launch("b2_run_small_program");
}
