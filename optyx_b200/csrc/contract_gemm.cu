// Piece (2b): large dense pairwise contraction as a fused index-permute +
// complex GEMM on the FP64 tensor pipe.
//
//   C[b, m, n] (+)= alpha * sum_k A[b, m, k] * B[b, k, n]
//
// b, m, n, k are *flattened groups of tensor modes*; each mode carries its own
// element stride into A, B and C, so the transposes the reference performs as
// separate numpy passes before BLAS zgemm (tensordot under quimb/cotengra,
// reached from optyx/core/backends.py:514 / 539-545) never touch HBM: the
// gather happens on the way into shared memory (cp.async, 16 B per complex128
// element, 4-stage ring) and the scatter on the way out of the accumulators.
//
// FP64 has no tcgen05 kind on sm_100a; the FP64 tensor path is the warp-level
// DMMA `mma.sync.aligned.m8n8k4.f64`.  A complex product is 4 real DMMAs on the
// (re, im) halves of interleaved fragments loaded with one LDS.128 each.
//
// CTA tile 64 x 64 x 8 (complex), 4 warps as 2 x 2, warp tile 32 x 32:
// per k4 step a warp issues 64 DMMAs for 16 LDS.128 -> the math pipe is the
// only busy unit.  Shared pitch 66 (x16 B) makes fragment loads conflict-free.
#include <stdlib.h>
#include "b2_common.cuh"

namespace b2 {

constexpr int BM = 64, BN = 64, PITCH = 66, GEMM_THREADS = 128;
// K chunk and ring depth are template parameters of the kernel: (8, 4) for short
// reductions, (16, 3) for long ones (half the block barriers per multiply-add: the ncu
// capture of the (8, 4) kernel shows 20 % of the stall samples on the barrier)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// flat group index -> three offsets; mode 0 of the group varies fastest
__device__ __forceinline__ void decode3(long long flat, const b2_gemm_modes& g, int first, int count,
                                        long long& oa, long long& ob, long long& oc) {
    oa = ob = oc = 0;
    for (int m = 0; m < count; ++m) {
        const int d = g.dim[first + m];
        const long long q = flat / d;
        const long long i = flat - q * d;
        flat = q;
        oa += i * g.sa[first + m];
        ob += i * g.sb[first + m];
        oc += i * g.sc[first + m];
    }
}

// One K chunk of one operand, global -> shared ([k][pitch] layout), 16 B per cp.async.
// `row_off` / `k_off` are BYTE offsets (tile rows / chunk reduction values through the
// per-mode strides).  Rows beyond the matrix edge point at row 0 (valid memory, their
// results are never stored); only the K tail needs zero fill (`kvalid` < BK in the last
// chunk).  The thread -> element map keeps the per-copy work at one table load + one
// 64-bit add, keeps global reads sector-exact and shared-memory writes conflict-free:
//   K contiguous in the operand: a quarter warp = 4 consecutive k x 2 rows (2 full
//     sectors per row; bank = 8*k + 4*row (mod 32) -- with the naive "k fastest" map the
//     66-element pitch makes every write 2-way conflicted), a warp = BK k x 32/BK... rows
//   rows contiguous: consecutive threads = consecutive rows of one k.
template <int ROWS, int BK, int THREADS, int PITCH_>
__device__ __forceinline__ void gather_chunk(double2* dst, const double2* __restrict__ base,
                                             const long long* row_off, const long long* k_off,
                                             int tid, bool kfast, int kvalid) {
    constexpr int P = ROWS * BK / THREADS;
    static_assert(P >= 1 && (ROWS * BK) % THREADS == 0, "tile / thread count mismatch");
    const char* b = reinterpret_cast<const char*>(base);
    if (kfast) {
        constexpr int LOGBK = BK == 64 ? 6 : BK == 32 ? 5 : BK == 16 ? 4 : 3;
        static_assert(BK == 8 || BK == 16 || BK == 32 || BK == 64, "BK");
        const int j = (tid & 3) | ((tid >> 1) & (BK - 4));
        const int i0 = ((tid >> 2) & 1) | ((tid >> (LOGBK + 1)) << 1);
        const char* src = b + k_off[j];
        const int bytes = j < kvalid ? 16 : 0;
        double2* d = dst + j * PITCH_ + i0;
        #pragma unroll
        for (int p = 0; p < P; ++p)
            cp_async16(d + p * (THREADS / BK), src + row_off[i0 + p * (THREADS / BK)], bytes);
    } else {
        const int i = tid % ROWS, j0 = tid / ROWS;
        const char* src = b + row_off[i];
        double2* d = dst + j0 * PITCH_ + i;
        #pragma unroll
        for (int p = 0; p < P; ++p) {
            const int j = j0 + p * (THREADS / ROWS);
            cp_async16(d + p * (THREADS / ROWS) * PITCH_, src + k_off[j], j < kvalid ? 16 : 0);
        }
    }
}

// Tile order: groups of RASTER_GM row tiles x all column tiles, rows fastest inside a
// group.  The ~2 x 148 CTAs resident at one time then form a near-square block of the
// tile grid, so every operand tile is fetched from HBM about once and re-used out of L2
// (ncu, cfg4's (2,16384,512,2048) pair with rows fastest over the WHOLE grid: 11.4 GB
// read for 1.1 GB of operands -- A streamed once per column tile).
constexpr int RASTER_GM = 16;
__device__ __forceinline__ void raster_tile(long long tile, int tiles_m, int tiles_n, int& tm, int& tn) {
    const long long per_group = (long long)RASTER_GM * tiles_n;
    const int group = (int)(tile / per_group);
    const int first_m = group * RASTER_GM;
    const int gm = tiles_m - first_m < RASTER_GM ? tiles_m - first_m : RASTER_GM;
    const int r = (int)(tile - (long long)group * per_group);
    tm = first_m + r % gm;
    tn = r / gm;
}

template <int BK, int STAGES>
struct gemm_smem {
    double2 a[STAGES][BK][PITCH];
    double2 b[STAGES][BK][PITCH];
    long long row_a[BM], row_c[BM], col_b[BN], col_c[BN];
    long long k_a[STAGES][BK], k_b[STAGES][BK];
};

template <bool SPLITK, int BK, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
contract_gemm_c128_kernel(const __grid_constant__ b2_gemm_modes g, const double2* __restrict__ A,
                          const double2* __restrict__ B, double2* __restrict__ Cout,
                          long long M, long long N, long long K, int tiles_m, long long tiles_mn,
                          int a_kfast, int b_kfast, double alpha_re, double alpha_im,
                          int accumulate, int ksplit) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using smem_t = gemm_smem<BK, STAGES>;
    smem_t& sm = *reinterpret_cast<smem_t*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    // grid.x = batch x tiles (the batch extent -- eval batch times every hyper-index
    // batch mode -- can exceed the 65535 limit of grid.y)
    const long long tile = blockIdx.x % tiles_mn, batch_id = blockIdx.x / tiles_mn;
    int tile_m, tile_n;
    raster_tile(tile, tiles_m, (int)(tiles_mn / tiles_m), tile_m, tile_n);
    const long long m0 = (long long)tile_m * BM, n0 = (long long)tile_n * BN;
    const int f_b = 0, f_m = g.n_batch, f_n = f_m + g.n_m, f_k = f_n + g.n_n;

    long long ba, bb, bc;
    decode3(batch_id, g, f_b, g.n_batch, ba, bb, bc);
    A += ba; B += bb; Cout += bc;

    // per-tile row / column offsets (index permutation, decoded once per CTA)
    if (tid < BM) {
        long long oa, ob, oc;
        const long long m = m0 + tid;
        decode3(m < M ? m : 0, g, f_m, g.n_m, oa, ob, oc);
        sm.row_a[tid] = oa * 16; sm.row_c[tid] = oc;          // operand tables in BYTES
    } else {
        const int t = tid - BM;
        long long oa, ob, oc;
        const long long n = n0 + t;
        decode3(n < N ? n : 0, g, f_n, g.n_n, oa, ob, oc);
        sm.col_b[t] = ob * 16; sm.col_c[t] = oc;
    }
    // split-K (grid z): this CTA reduces chunks [c_begin, c_begin + nchunks) and adds
    // its partial tile into the zero-filled C with FP64 atomics
    // (the unsplit instantiation is the plain kernel: c_begin folds to 0)
    const int all_chunks = (int)((K + BK - 1) / BK);
    const int per_split = SPLITK ? (all_chunks + ksplit - 1) / ksplit : all_chunks;
    const int c_begin = SPLITK ? (int)blockIdx.z * per_split : 0;
    const int nchunks = !SPLITK ? all_chunks
                        : (all_chunks - c_begin < per_split
                               ? (all_chunks - c_begin > 0 ? all_chunks - c_begin : 0) : per_split);

    // reduction offsets of one chunk: 32-bit mixed-radix decode (K < 2^31) by the
    // last warp, so the 64-bit divisions of a many-mode K never sit on the critical
    // path of the warps that issue the copies first
    auto decode_k = [&](int chunk) {
        const int t = tid - (GEMM_THREADS - 32);
        if (t >= 0 && t < BK) {
            unsigned k = (unsigned)(c_begin + chunk) * BK + (unsigned)t;
            if ((long long)k >= K) k = 0;
            long long oa = 0, ob = 0;
            for (int m = 0; m < g.n_k; ++m) {
                const unsigned d = (unsigned)g.dim[f_k + m];
                const unsigned q = k / d, i = k - q * d;
                k = q;
                oa += (long long)i * g.sa[f_k + m];
                ob += (long long)i * g.sb[f_k + m];
            }
            sm.k_a[chunk % STAGES][t] = oa * 16;
            sm.k_b[chunk % STAGES][t] = ob * 16;
        }
    };
    auto issue = [&](int chunk) {
        const int st = chunk % STAGES;
        const long long kleft = K - (long long)(c_begin + chunk) * BK;
        const int kvalid = kleft < BK ? (int)kleft : BK;
        gather_chunk<BM, BK, GEMM_THREADS, PITCH>(&sm.a[st][0][0], A, sm.row_a, sm.k_a[st], tid,
                                                   a_kfast != 0, kvalid);
        gather_chunk<BN, BK, GEMM_THREADS, PITCH>(&sm.b[st][0][0], B, sm.col_b, sm.k_b[st], tid,
                                                   b_kfast != 0, kvalid);
    };

    for (int c = 0; c < STAGES - 1; ++c) if (c < nchunks) decode_k(c);
    __syncthreads();
    for (int c = 0; c < STAGES - 1; ++c) {
        if (c < nchunks) issue(c);
        cp_async_commit();
    }

    double cre[4][4][2], cim[4][4][2];
    #pragma unroll
    for (int r = 0; r < 4; ++r)
        #pragma unroll
        for (int c = 0; c < 4; ++c) { cre[r][c][0] = cre[r][c][1] = 0; cim[r][c][0] = cim[r][c][1] = 0; }

    const int frow = lane >> 2, fk = lane & 3;
    for (int c = 0; c < nchunks; ++c) {
        const int nxt = c + STAGES - 1;
        if (nxt < nchunks) decode_k(nxt);
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (nxt < nchunks) issue(nxt);
        cp_async_commit();
        const int st = c % STAGES;
        #pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double2 af[4], bf[4];
            #pragma unroll
            for (int r = 0; r < 4; ++r) af[r] = sm.a[st][k4 + fk][wm * 32 + r * 8 + frow];
            #pragma unroll
            for (int q = 0; q < 4; ++q) bf[q] = sm.b[st][k4 + fk][wn * 32 + q * 8 + frow];
            // the asm statements are volatile (kept in source order): issue the 8
            // independent accumulators of a row before coming back to any of them,
            // so that a dependent DMMA is 8 instructions behind its producer
            #pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double nai = -af[r].y;
                #pragma unroll
                for (int q = 0; q < 4; ++q) dmma884(cre[r][q][0], cre[r][q][1], af[r].x, bf[q].x);
                #pragma unroll
                for (int q = 0; q < 4; ++q) dmma884(cim[r][q][0], cim[r][q][1], af[r].x, bf[q].y);
                #pragma unroll
                for (int q = 0; q < 4; ++q) dmma884(cre[r][q][0], cre[r][q][1], nai, bf[q].y);
                #pragma unroll
                for (int q = 0; q < 4; ++q) dmma884(cim[r][q][0], cim[r][q][1], af[r].y, bf[q].x);
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: scale, scatter through the C permutation
    #pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = wm * 32 + r * 8 + frow;
        if (m0 + i >= M) continue;
        const long long ro = sm.row_c[i];
        #pragma unroll
        for (int q = 0; q < 4; ++q) {
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = wn * 32 + q * 8 + fk * 2 + h;
                if (n0 + j >= N) continue;
                double2 v;
                v.x = alpha_re * cre[r][q][h] - alpha_im * cim[r][q][h];
                v.y = alpha_re * cim[r][q][h] + alpha_im * cre[r][q][h];
                double2* dst = Cout + ro + sm.col_c[j];
                if (SPLITK) {
                    atomicAdd(&dst->x, v.x);
                    atomicAdd(&dst->y, v.y);
                } else {
                    if (accumulate) { const double2 o = *dst; v.x += o.x; v.y += o.y; }
                    *dst = v;
                }
            }
        }
    }
}

// ---- 3M complex multiplication for long reductions.  A complex product needs only
// three real products:   P1 = sum ar*br,  P2 = sum ai*bi,  P3 = sum (ar+ai)*(br+bi)
//                        C_re = P1 - P2,  C_im = P3 - P1 - P2
// i.e. 3 DMMAs per complex tile product instead of 4: the FP64 tensor pipe -- the unit
// that bounds these pairs -- does 25 % less work per contraction (cuBLAS exposes the
// same trade as zgemm3m).  The price is a third accumulator set, so the warp tile is
// 32 x 16 (48 accumulator doubles per thread) and the two operand sums are formed in
// registers from the fragments (6 DADD per 24 DMMA).  The imaginary part is a
// difference of sums: its error is bounded relative to |A||B| (a few ulp of the larger
// of the two parts of an output), far inside the 1e-10 bar of this path.
// CTA tile 64 x BN3 (BN3 = 64: 8 warps, one CTA per SM; BN3 = 32: 4 warps, two CTAs
// per SM), K chunk 16, 3-stage cp.async gather ring as above.
template <int BN3>
struct gemm3m_smem {
    double2 a[3][16][PITCH];
    double2 b[3][16][BN3 + 2];
    long long row_a[BM], row_c[BM], col_b[BN3], col_c[BN3];
    long long k_a[3][16], k_b[3][16];
};

template <bool SPLITK, int BN3>
__global__ void __launch_bounds__(BN3 * 4, BN3 == 64 ? 1 : 2)
contract_gemm3m_c128_kernel(const __grid_constant__ b2_gemm_modes g, const double2* __restrict__ A,
                            const double2* __restrict__ B, double2* __restrict__ Cout,
                            long long M, long long N, long long K, int tiles_m, long long tiles_mn,
                            int a_kfast, int b_kfast, double alpha_re, double alpha_im,
                            int accumulate, int ksplit) {
    constexpr int BK = 16, STAGES = 3, THREADS = BN3 * 4, WN = BN3 / 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using smem_t = gemm3m_smem<BN3>;
    smem_t& sm = *reinterpret_cast<smem_t*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const long long tile = blockIdx.x % tiles_mn, batch_id = blockIdx.x / tiles_mn;
    int tile_m, tile_n;
    raster_tile(tile, tiles_m, (int)(tiles_mn / tiles_m), tile_m, tile_n);
    const long long m0 = (long long)tile_m * BM, n0 = (long long)tile_n * BN3;
    const int f_m = g.n_batch, f_n = f_m + g.n_m, f_k = f_n + g.n_n;

    long long ba, bb, bc;
    decode3(batch_id, g, 0, g.n_batch, ba, bb, bc);
    A += ba; B += bb; Cout += bc;
    if (tid < BM) {
        long long oa, ob, oc;
        const long long m = m0 + tid;
        decode3(m < M ? m : 0, g, f_m, g.n_m, oa, ob, oc);
        sm.row_a[tid] = oa * 16; sm.row_c[tid] = oc;
    } else if (tid < BM + BN3) {
        const int t = tid - BM;
        long long oa, ob, oc;
        const long long n = n0 + t;
        decode3(n < N ? n : 0, g, f_n, g.n_n, oa, ob, oc);
        sm.col_b[t] = ob * 16; sm.col_c[t] = oc;
    }
    const int all_chunks = (int)((K + BK - 1) / BK);
    const int per_split = SPLITK ? (all_chunks + ksplit - 1) / ksplit : all_chunks;
    const int c_begin = SPLITK ? (int)blockIdx.z * per_split : 0;
    const int nchunks = !SPLITK ? all_chunks
                        : (all_chunks - c_begin < per_split
                               ? (all_chunks - c_begin > 0 ? all_chunks - c_begin : 0) : per_split);
    auto decode_k = [&](int chunk) {
        const int t = tid - (THREADS - 32);
        if (t >= 0 && t < BK) {
            unsigned k = (unsigned)(c_begin + chunk) * BK + (unsigned)t;
            if ((long long)k >= K) k = 0;
            long long oa = 0, ob = 0;
            for (int m = 0; m < g.n_k; ++m) {
                const unsigned d = (unsigned)g.dim[f_k + m];
                const unsigned q = k / d, i = k - q * d;
                k = q;
                oa += (long long)i * g.sa[f_k + m];
                ob += (long long)i * g.sb[f_k + m];
            }
            sm.k_a[chunk % STAGES][t] = oa * 16;
            sm.k_b[chunk % STAGES][t] = ob * 16;
        }
    };
    auto issue = [&](int chunk) {
        const int st = chunk % STAGES;
        const long long kleft = K - (long long)(c_begin + chunk) * BK;
        const int kvalid = kleft < BK ? (int)kleft : BK;
        gather_chunk<BM, BK, THREADS, PITCH>(&sm.a[st][0][0], A, sm.row_a, sm.k_a[st], tid,
                                              a_kfast != 0, kvalid);
        gather_chunk<BN3, BK, THREADS, BN3 + 2>(&sm.b[st][0][0], B, sm.col_b, sm.k_b[st], tid,
                                                 b_kfast != 0, kvalid);
    };

    for (int c = 0; c < STAGES - 1; ++c) if (c < nchunks) decode_k(c);
    __syncthreads();
    for (int c = 0; c < STAGES - 1; ++c) {
        if (c < nchunks) issue(c);
        cp_async_commit();
    }

    double p1[4][2][2], p2[4][2][2], p3[4][2][2];
    #pragma unroll
    for (int r = 0; r < 4; ++r)
        #pragma unroll
        for (int q = 0; q < 2; ++q)
            #pragma unroll
            for (int h = 0; h < 2; ++h) p1[r][q][h] = p2[r][q][h] = p3[r][q][h] = 0.0;

    const int frow = lane >> 2, fk = lane & 3;
    for (int c = 0; c < nchunks; ++c) {
        const int nxt = c + STAGES - 1;
        if (nxt < nchunks) decode_k(nxt);
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (nxt < nchunks) issue(nxt);
        cp_async_commit();
        const int st = c % STAGES;
        #pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double2 af[4], bf[2];
            double as[4], bs[2];
            #pragma unroll
            for (int r = 0; r < 4; ++r) af[r] = sm.a[st][k4 + fk][wm * 32 + r * 8 + frow];
            #pragma unroll
            for (int q = 0; q < 2; ++q) bf[q] = sm.b[st][k4 + fk][wn * 16 + q * 8 + frow];
            #pragma unroll
            for (int r = 0; r < 4; ++r) as[r] = af[r].x + af[r].y;
            #pragma unroll
            for (int q = 0; q < 2; ++q) bs[q] = bf[q].x + bf[q].y;
            // 24 independent accumulators: a dependent DMMA is 24 instructions behind
            #pragma unroll
            for (int r = 0; r < 4; ++r)
                #pragma unroll
                for (int q = 0; q < 2; ++q) dmma884(p1[r][q][0], p1[r][q][1], af[r].x, bf[q].x);
            #pragma unroll
            for (int r = 0; r < 4; ++r)
                #pragma unroll
                for (int q = 0; q < 2; ++q) dmma884(p2[r][q][0], p2[r][q][1], af[r].y, bf[q].y);
            #pragma unroll
            for (int r = 0; r < 4; ++r)
                #pragma unroll
                for (int q = 0; q < 2; ++q) dmma884(p3[r][q][0], p3[r][q][1], as[r], bs[q]);
        }
    }
    cp_async_wait<0>();

    #pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = wm * 32 + r * 8 + frow;
        if (m0 + i >= M) continue;
        const long long ro = sm.row_c[i];
        #pragma unroll
        for (int q = 0; q < 2; ++q) {
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = wn * 16 + q * 8 + fk * 2 + h;
                if (n0 + j >= N) continue;
                const double re = p1[r][q][h] - p2[r][q][h];
                const double im = p3[r][q][h] - p1[r][q][h] - p2[r][q][h];
                double2 v;
                v.x = alpha_re * re - alpha_im * im;
                v.y = alpha_re * im + alpha_im * re;
                double2* dst = Cout + ro + sm.col_c[j];
                if (SPLITK) {
                    atomicAdd(&dst->x, v.x);
                    atomicAdd(&dst->y, v.y);
                } else {
                    if (accumulate) { const double2 o = *dst; v.x += o.x; v.y += o.y; }
                    *dst = v;
                }
            }
        }
    }
}

// ---- thin pairs: the whole M x N extent fits one small tile (<= 32 x 32) and the
// reduction is long (the closing pairs of a sliced amplitude network: two big
// intermediates that share almost all of their modes, e.g. (batch 2, 16 x 16,
// K = 2^19)).  A 64 x 64 tile would be 1/16 full -- 16x wasted gathers and DMMAs
// (cfg4, r01 kernel: 1959 us for 82 us of HBM traffic).  Here the tile is (8T) x (8T),
// T = 1 | 2 | 4, K chunks of 1024 / (16 T) reduction values, the reduction range is split
// over enough CTAs to fill the GPU, and inside a CTA the 4 warps split the k4 steps of
// every chunk (each warp accumulates the full tile with the 3M scheme; one shared-
// memory reduction at the end, then one atomic add per output and CTA).  HBM-bound.
template <int T> struct thin_cfg {
    static constexpr int ROWS = 8 * T, BK = T == 4 ? 32 : 64, PITCH_T = ROWS + 2;
    static constexpr int STAGES = 3, THREADS = 128;
};
template <int T>
struct thin_smem {
    using cfg = thin_cfg<T>;
    double2 a[cfg::STAGES][cfg::BK][cfg::PITCH_T];
    double2 b[cfg::STAGES][cfg::BK][cfg::PITCH_T];
    long long row_a[cfg::ROWS], row_c[cfg::ROWS], col_b[cfg::ROWS], col_c[cfg::ROWS];
    long long k_a[cfg::STAGES][cfg::BK], k_b[cfg::STAGES][cfg::BK];
};

template <int T>
__global__ void __launch_bounds__(128, 2)
contract_gemm_thin_kernel(const __grid_constant__ b2_gemm_modes g, const double2* __restrict__ A,
                          const double2* __restrict__ B, double2* __restrict__ Cout,
                          long long M, long long N, long long K, int a_kfast, int b_kfast,
                          double alpha_re, double alpha_im, int ksplit) {
    using cfg = thin_cfg<T>;
    constexpr int BK = cfg::BK, STAGES = cfg::STAGES, THREADS = cfg::THREADS, ROWS = cfg::ROWS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    thin_smem<T>& sm = *reinterpret_cast<thin_smem<T>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f_m = g.n_batch, f_n = f_m + g.n_m, f_k = f_n + g.n_n;
    const long long batch_id = blockIdx.x / ksplit;
    const int split = (int)(blockIdx.x % ksplit);
    long long ba, bb, bc;
    decode3(batch_id, g, 0, g.n_batch, ba, bb, bc);
    A += ba; B += bb; Cout += bc;
    if (tid < ROWS) {
        long long oa, ob, oc;
        decode3(tid < M ? tid : 0, g, f_m, g.n_m, oa, ob, oc);
        sm.row_a[tid] = oa * 16; sm.row_c[tid] = oc;
    } else if (tid < 2 * ROWS) {
        const int t = tid - ROWS;
        long long oa, ob, oc;
        decode3(t < N ? t : 0, g, f_n, g.n_n, oa, ob, oc);
        sm.col_b[t] = ob * 16; sm.col_c[t] = oc;
    }
    const int all_chunks = (int)((K + BK - 1) / BK);
    const int per_split = (all_chunks + ksplit - 1) / ksplit;
    const int c_begin = split * per_split;
    const int nchunks = all_chunks - c_begin < per_split
                            ? (all_chunks - c_begin > 0 ? all_chunks - c_begin : 0) : per_split;
    auto decode_k = [&](int chunk) {
        if (tid < BK) {
            unsigned k = (unsigned)(c_begin + chunk) * BK + (unsigned)tid;
            if ((long long)k >= K) k = 0;
            long long oa = 0, ob = 0;
            for (int m = 0; m < g.n_k; ++m) {
                const unsigned d = (unsigned)g.dim[f_k + m];
                const unsigned q = k / d, i = k - q * d;
                k = q;
                oa += (long long)i * g.sa[f_k + m];
                ob += (long long)i * g.sb[f_k + m];
            }
            sm.k_a[chunk % STAGES][tid] = oa * 16;
            sm.k_b[chunk % STAGES][tid] = ob * 16;
        }
    };
    auto issue = [&](int chunk) {
        const int st = chunk % STAGES;
        const long long kleft = K - (long long)(c_begin + chunk) * BK;
        const int kvalid = kleft < BK ? (int)kleft : BK;
        gather_chunk<ROWS, BK, THREADS, cfg::PITCH_T>(&sm.a[st][0][0], A, sm.row_a, sm.k_a[st], tid,
                                                       a_kfast != 0, kvalid);
        gather_chunk<ROWS, BK, THREADS, cfg::PITCH_T>(&sm.b[st][0][0], B, sm.col_b, sm.k_b[st], tid,
                                                       b_kfast != 0, kvalid);
    };
    for (int c = 0; c < STAGES - 1; ++c) if (c < nchunks) decode_k(c);
    __syncthreads();
    for (int c = 0; c < STAGES - 1; ++c) {
        if (c < nchunks) issue(c);
        cp_async_commit();
    }
    double p1[T][T][2], p2[T][T][2], p3[T][T][2];
    #pragma unroll
    for (int r = 0; r < T; ++r)
        #pragma unroll
        for (int q = 0; q < T; ++q)
            #pragma unroll
            for (int h = 0; h < 2; ++h) p1[r][q][h] = p2[r][q][h] = p3[r][q][h] = 0.0;
    const int frow = lane >> 2, fk = lane & 3;
    for (int c = 0; c < nchunks; ++c) {
        const int nxt = c + STAGES - 1;
        if (nxt < nchunks) decode_k(nxt);
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (nxt < nchunks) issue(nxt);
        cp_async_commit();
        const int st = c % STAGES;
        #pragma unroll
        for (int s4 = 0; s4 < BK / 16; ++s4) {              // this warp's k4 steps of the chunk
            const int k4 = (s4 * 4 + warp) * 4;
            double2 af[T], bf[T];
            double as[T], bs[T];
            #pragma unroll
            for (int r = 0; r < T; ++r) { af[r] = sm.a[st][k4 + fk][r * 8 + frow]; as[r] = af[r].x + af[r].y; }
            #pragma unroll
            for (int q = 0; q < T; ++q) { bf[q] = sm.b[st][k4 + fk][q * 8 + frow]; bs[q] = bf[q].x + bf[q].y; }
            #pragma unroll
            for (int r = 0; r < T; ++r)
                #pragma unroll
                for (int q = 0; q < T; ++q) dmma884(p1[r][q][0], p1[r][q][1], af[r].x, bf[q].x);
            #pragma unroll
            for (int r = 0; r < T; ++r)
                #pragma unroll
                for (int q = 0; q < T; ++q) dmma884(p2[r][q][0], p2[r][q][1], af[r].y, bf[q].y);
            #pragma unroll
            for (int r = 0; r < T; ++r)
                #pragma unroll
                for (int q = 0; q < T; ++q) dmma884(p3[r][q][0], p3[r][q][1], as[r], bs[q]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();                                        // the ring is free: reuse it
    // cross-warp reduction: red[warp][tile element][re, im]
    double2* red = reinterpret_cast<double2*>(&sm.a[0][0][0]);
    #pragma unroll
    for (int r = 0; r < T; ++r)
        #pragma unroll
        for (int q = 0; q < T; ++q)
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = r * 8 + frow, j = q * 8 + fk * 2 + h;
                double2 v;
                v.x = p1[r][q][h] - p2[r][q][h];
                v.y = p3[r][q][h] - p1[r][q][h] - p2[r][q][h];
                red[warp * ROWS * ROWS + i * ROWS + j] = v;
            }
    __syncthreads();
    for (int e = tid; e < ROWS * ROWS; e += THREADS) {
        const int i = e / ROWS, j = e % ROWS;
        if (i >= M || j >= N) continue;
        double re = 0.0, im = 0.0;
        #pragma unroll
        for (int w = 0; w < 4; ++w) { const double2 v = red[w * ROWS * ROWS + e]; re += v.x; im += v.y; }
        double2* dst = Cout + sm.row_c[i] + sm.col_c[j];
        atomicAdd(&dst->x, alpha_re * re - alpha_im * im);
        atomicAdd(&dst->y, alpha_re * im + alpha_im * re);
    }
}

template <int T>
static int launch_thin_t(const b2_gemm_modes& g, const void* a, const void* b, void* c, long long Bt,
                         long long M, long long N, long long K, int a_kfast, int b_kfast, double ar,
                         double ai, cudaStream_t s) {
    using cfg = thin_cfg<T>;
    const int smem = (int)sizeof(thin_smem<T>);
    static b2::once_flags attr_set;
    if (!b2::is_done(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(contract_gemm_thin_kernel<T>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("b2_contract_gemm(thin): smem attr: %s", cudaGetErrorString(e)); return 1; }
        b2::mark_done(attr_set);
    }
    const long long chunks = (K + cfg::BK - 1) / cfg::BK;
    long long ksplit = (4ll * num_sms() + Bt - 1) / Bt;      // ~2 waves of 2 CTAs per SM
    if (ksplit > chunks / 8) ksplit = chunks / 8;            // >= 8 chunks per CTA
    if (ksplit < 1) ksplit = 1;
    if (Bt * ksplit > 2147483647ll) { set_error("b2_contract_gemm(thin): grid too large"); return 1; }
    contract_gemm_thin_kernel<T><<<(unsigned)(Bt * ksplit), cfg::THREADS, smem, s>>>(
        g, (const double2*)a, (const double2*)b, (double2*)c, M, N, K, a_kfast, b_kfast, ar, ai,
        (int)ksplit);
    return check_launch("b2_contract_gemm(thin)");
}

// ---- few outputs, very long reduction with many modes (the last pair of a sliced
// amplitude network: <psi_left | psi_right> over 2^20 index values spread over 20
// binary modes).  A tile of the DMMA kernel would be 1/4096 full; here every CTA takes a
// contiguous range of the flat reduction index, decodes it once (mixed radix) and then
// walks it with an odometer, one complex FMA per element; block-reduce, one atomic per
// output and CTA.  C must hold zeros (or the running sum when accumulating).
constexpr int DOT_THREADS = 256;
constexpr int DOT_MAX_OUT = 16;

template <typename T>
__global__ void __launch_bounds__(DOT_THREADS)
contract_dot_kernel(const __grid_constant__ b2_gemm_modes g, long long n_out, long long K,
                    const typename cplx<T>::type* __restrict__ A,
                    const typename cplx<T>::type* __restrict__ B,
                    typename cplx<T>::type* __restrict__ Cout, T alpha_re, T alpha_im) {
    using C = typename cplx<T>::type;
    __shared__ double red_re[DOT_THREADS / 32], red_im[DOT_THREADS / 32];
    const int f_k = g.n_batch + g.n_m + g.n_n, n_free = f_k;
    const long long per_cta = (K + gridDim.x - 1) / gridDim.x;
    const long long k0 = (long long)blockIdx.x * per_cta;
    const long long k1 = k0 + per_cta < K ? k0 + per_cta : K;
    for (long long o = 0; o < n_out; ++o) {
        long long oa = 0, ob = 0, oc = 0, rem = o;
        for (int m = 0; m < n_free; ++m) {               // mode 0 of every group is fastest
            const int d = g.dim[m];
            const long long i = rem % d;
            rem /= d;
            oa += i * g.sa[m]; ob += i * g.sb[m]; oc += i * g.sc[m];
        }
        double acc_re = 0.0, acc_im = 0.0;
        // thread t takes k0 + t, k0 + t + 256, ...: decode, then step the odometer by 256
        long long k = k0 + threadIdx.x;
        if (k < k1) {
            int idx[B2_MAX_MODES];
            long long ra = 0, rb = 0, r = k;
            for (int m = 0; m < g.n_k; ++m) {
                const int d = g.dim[f_k + m];
                idx[m] = (int)(r % d);
                r /= d;
                ra += idx[m] * g.sa[f_k + m]; rb += idx[m] * g.sb[f_k + m];
            }
            for (; k < k1; k += DOT_THREADS) {
                const C x = A[oa + ra], y = B[ob + rb];
                acc_re += (double)x.x * (double)y.x - (double)x.y * (double)y.y;
                acc_im += (double)x.x * (double)y.y + (double)x.y * (double)y.x;
                // advance the mixed-radix counter by DOT_THREADS
                long long carry = DOT_THREADS;
                for (int m = 0; m < g.n_k && carry; ++m) {
                    const int d = g.dim[f_k + m];
                    const long long v = idx[m] + carry;
                    const int nv = (int)(v % d);
                    carry = v / d;
                    ra += (long long)(nv - idx[m]) * g.sa[f_k + m];
                    rb += (long long)(nv - idx[m]) * g.sb[f_k + m];
                    idx[m] = nv;
                }
            }
        }
        for (int m = 16; m > 0; m >>= 1) {
            acc_re += __shfl_xor_sync(0xffffffffu, acc_re, m);
            acc_im += __shfl_xor_sync(0xffffffffu, acc_im, m);
        }
        if ((threadIdx.x & 31) == 0) { red_re[threadIdx.x >> 5] = acc_re; red_im[threadIdx.x >> 5] = acc_im; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double sr = 0.0, si = 0.0;
            for (int w = 0; w < DOT_THREADS / 32; ++w) { sr += red_re[w]; si += red_im[w]; }
            C* dst = Cout + oc;
            atomicAdd(&dst->x, (T)(alpha_re * sr - alpha_im * si));
            atomicAdd(&dst->y, (T)(alpha_re * si + alpha_im * sr));
        }
        __syncthreads();
    }
}

// returns handled = true when the pair was a tiny-output / long-reduction one
static int launch_dot(const b2_gemm_modes& g, const void* a, const void* b, void* c, double ar,
                      double ai, int accumulate, int dtype, cudaStream_t s, bool& handled) {
    handled = false;
    const int f_k = g.n_batch + g.n_m + g.n_n;
    long long n_out = 1, K = 1, span = 0;
    for (int i = 0; i < f_k; ++i) { n_out *= g.dim[i]; span += (long long)(g.dim[i] - 1) * g.sc[i]; }
    for (int i = 0; i < g.n_k; ++i) K *= g.dim[f_k + i];
    if (n_out > DOT_MAX_OUT || K < 4096) return 0;
    const int sms = num_sms();
    if (sms <= 0) return 1;
    const size_t esize = dtype == B2_C128 ? 16 : 8;
    if (!accumulate) {                                   // the CTAs add into C
        if (span + 1 != n_out) return 0;                 // C is not one dense block
        cudaError_t e = cudaMemsetAsync(c, 0, (size_t)n_out * esize, s);
        if (e != cudaSuccess) { set_error("b2_contract_gemm(dot): memset: %s", cudaGetErrorString(e)); return 1; }
    }
    long long blocks = (K + 4 * DOT_THREADS - 1) / (4 * DOT_THREADS);
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    if (dtype == B2_C128)
        contract_dot_kernel<double><<<(unsigned)blocks, DOT_THREADS, 0, s>>>(
            g, n_out, K, (const double2*)a, (const double2*)b, (double2*)c, ar, ai);
    else
        contract_dot_kernel<float><<<(unsigned)blocks, DOT_THREADS, 0, s>>>(
            g, n_out, K, (const float2*)a, (const float2*)b, (float2*)c, (float)ar, (float)ai);
    handled = true;
    set_last_kernel("dot");
    return check_launch("b2_contract_gemm(dot)");
}

int launch_ctile(const b2_gemm_modes& g, const void* a, const void* b, void* c, double ar,
                 double ai, int accumulate, cudaStream_t s, bool& handled);
int launch_gemm_c64(const b2_gemm_modes& g, const void* a_dev, const void* b_dev, void* c_dev,
                    double ar, double ai, int accumulate, cudaStream_t s);
int launch_gemm_ws(const b2_gemm_modes& g, const void* a, const void* b, void* c, long long Bt,
                   long long M, long long N, long long K, int a_kfast, int b_kfast, double ar,
                   double ai, int accumulate, cudaStream_t s, bool& handled);

}  // namespace b2

extern "C" int b2_contract_gemm(const b2_gemm_modes* modes, const void* a_dev, const void* b_dev,
                                void* c_dev, double alpha_re, double alpha_im,
                                int32_t accumulate, int32_t dtype, void* stream) {
    using namespace b2;
    if (!modes || !a_dev || !b_dev || !c_dev) { set_error("b2_contract_gemm: null pointer"); return 1; }
    if (dtype != B2_C128 && dtype != B2_C64) {
        set_error("b2_contract_gemm: bad dtype %d", dtype);
        return 1;
    }
    const b2_gemm_modes& g = *modes;
    const int total = g.n_batch + g.n_m + g.n_n + g.n_k;
    if (g.n_batch < 0 || g.n_m < 0 || g.n_n < 0 || g.n_k < 0 || total > B2_MAX_MODES) {
        set_error("b2_contract_gemm: bad mode counts");
        return 1;
    }
    if (num_sms() <= 0) return 1;
    long long k_total = 1, n_total = 1;
    for (int i = 0; i < g.n_k; ++i) k_total *= g.dim[g.n_batch + g.n_m + g.n_n + i];
    for (int i = 0; i < g.n_n; ++i) n_total *= g.dim[g.n_batch + g.n_m + i];
    if (k_total >= (1ll << 31)) {
        set_error("b2_contract_gemm: reduction extent %lld too large (slice the network)", k_total);
        return 1;
    }
    {   // a handful of outputs over a very long reduction: neither tensor-core tiling fits
        bool handled = false;
        int st = launch_dot(g, a_dev, b_dev, c_dev, alpha_re, alpha_im, accumulate, dtype,
                            (cudaStream_t)stream, handled);
        if (st != 0 || handled) return st;
    }
    if (dtype == B2_C64)      // tcgen05 kind::tf32, 3-product split precision (contract_gemm_c64.cu)
        return launch_gemm_c64(g, a_dev, b_dev, c_dev, alpha_re, alpha_im, accumulate,
                               (cudaStream_t)stream);
    long long Bt = 1, M = 1, N = 1, K = 1;
    {
        int p = 0;
        for (int i = 0; i < g.n_batch; ++i) Bt *= g.dim[p++];
        for (int i = 0; i < g.n_m; ++i) M *= g.dim[p++];
        for (int i = 0; i < g.n_n; ++i) N *= g.dim[p++];
        for (int i = 0; i < g.n_k; ++i) K *= g.dim[p++];
    }
    if (Bt == 0 || M == 0 || N == 0) return 0;
    const long long m_total = M, b_total = Bt;
    // load direction: walk the group that is contiguous in the operand
    auto min_stride = [&](const int64_t* s, int first, int count) {
        long long best = (1ll << 62);
        for (int i = 0; i < count; ++i)
            if (g.dim[first + i] > 1 && s[first + i] > 0 && s[first + i] < best) best = s[first + i];
        return best;
    };
    const int f_m = g.n_batch, f_n = f_m + g.n_m, f_k = f_n + g.n_n;
    const int a_kfast = min_stride(g.sa, f_k, g.n_k) < min_stride(g.sa, f_m, g.n_m);
    const int b_kfast = min_stride(g.sb, f_k, g.n_k) < min_stride(g.sb, f_n, g.n_n);
    // the persistent warp-specialised kernel (contract_gemm_ws.cu) takes every pair with at
    // least one full row panel, unless its tiles would leave most SMs idle while the reduction
    // is long (split-K below).  (B2_GEMM_WS = 0: off -- experiment knob)
    {
        static const bool use_ws = !(getenv("B2_GEMM_WS") && atoi(getenv("B2_GEMM_WS")) == 0);
        if (use_ws && M >= 64 && N >= 8 && K >= 16) {
            bool handled = false;
            int st = launch_gemm_ws(g, a_dev, b_dev, c_dev, Bt, M, N, K, a_kfast, b_kfast, alpha_re,
                                    alpha_im, accumulate, (cudaStream_t)stream, handled);
            if (st != 0 || handled) return st;
        }
    }
    // thin pairs (M, N <= 32, long K): split-K over the whole GPU with a tile of their size
    if (M <= 32 && N <= 32 && K >= 256 && !getenv("B2_NO_THIN")) {
        long long span = 0;
        for (int i = 0; i < f_k; ++i) span += (long long)(g.dim[i] - 1) * g.sc[i];
        if (span + 1 == Bt * M * N || accumulate) {          // C is one dense block (zero fill)
            if (!accumulate) {
                cudaError_t e = cudaMemsetAsync(c_dev, 0, (size_t)(Bt * M * N) * sizeof(double2),
                                                (cudaStream_t)stream);
                if (e != cudaSuccess) { set_error("b2_contract_gemm: memset: %s", cudaGetErrorString(e)); return 1; }
            }
            set_last_kernel("gemm_thin");
            const long long mx = M > N ? M : N;
            if (mx <= 8)
                return launch_thin_t<1>(g, a_dev, b_dev, c_dev, Bt, M, N, K, a_kfast, b_kfast, alpha_re, alpha_im, (cudaStream_t)stream);
            if (mx <= 16)
                return launch_thin_t<2>(g, a_dev, b_dev, c_dev, Bt, M, N, K, a_kfast, b_kfast, alpha_re, alpha_im, (cudaStream_t)stream);
            return launch_thin_t<4>(g, a_dev, b_dev, c_dev, Bt, M, N, K, a_kfast, b_kfast, alpha_re, alpha_im, (cudaStream_t)stream);
        }
    }
    const bool underfilled = ((m_total + BM - 1) / BM) * ((n_total + BN - 1) / BN) * b_total * 2 <=
                                 2ll * num_sms() && k_total >= 256;
    // reductions up to ctile_kmax take the ctile kernel (B2_CTILE_KMAX: experiment knob).
    // Measured (profiles/r02_shortk.md): from K = 16 on the 64x64 tiling is faster on
    // plain AND permuted layouts -- (2,16384,1024,32) of cfg4: 1336 -> 489 us
    static const long long ctile_kmax = getenv("B2_CTILE_KMAX") ? atoll(getenv("B2_CTILE_KMAX")) : 8;
    if ((k_total <= ctile_kmax || n_total < 64) && !underfilled) {
        // short reductions / skinny pairs: tile from the output side (coalesced
        // staged stores, one gather per tile).  Long-K square-ish pairs keep the
        // classic 64x64 tiling below (4x4 register blocking per warp).
        bool handled = false;
        int st = launch_ctile(g, a_dev, b_dev, c_dev, alpha_re, alpha_im, accumulate,
                              (cudaStream_t)stream, handled);
        if (st != 0 || handled) return st;
    }
    // long reductions: 3M complex multiplication (3 DMMAs per complex tile product, K chunk
    // 16); short ones: the 4-product kernel with chunk 8 / 4 stages.
    // (B2_GEMM_3M = 0 | 32 | 64: off / N tile of the 3M kernel; B2_GEMM_BK: experiment knobs.)
    static const int n3m = getenv("B2_GEMM_3M") ? atoi(getenv("B2_GEMM_3M")) : 32;
    static const int force_bk = getenv("B2_GEMM_BK") ? atoi(getenv("B2_GEMM_BK")) : 0;
    const bool three_m = (n3m == 32 || n3m == 64) && K >= 128;
    const int bn = three_m ? n3m : BN;
    const long long tiles_m = (M + BM - 1) / BM, tiles_n = (N + bn - 1) / bn;
    if (tiles_m * tiles_n * Bt > 2147483647ll) {
        set_error("b2_contract_gemm: grid too large (tiles=%lld batch=%lld)", tiles_m * tiles_n, Bt);
        return 1;
    }
    const int BK = three_m ? 16 : (force_bk ? force_bk : (K >= 128 ? 16 : 8));
    const int smem = three_m ? (bn == 64 ? (int)sizeof(gemm3m_smem<64>) : (int)sizeof(gemm3m_smem<32>))
                             : BK == 16 ? (int)sizeof(gemm_smem<16, 3>) : (int)sizeof(gemm_smem<8, 4>);
    static b2::once_flags attr_set;
    if (!b2::is_done(attr_set)) {
        cudaError_t e = cudaSuccess;
        auto set = [&](const void* fn, int bytes) {
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        };
        set((const void*)contract_gemm_c128_kernel<false, 8, 4>, (int)sizeof(gemm_smem<8, 4>));
        set((const void*)contract_gemm_c128_kernel<true, 8, 4>, (int)sizeof(gemm_smem<8, 4>));
        set((const void*)contract_gemm_c128_kernel<false, 16, 3>, (int)sizeof(gemm_smem<16, 3>));
        set((const void*)contract_gemm_c128_kernel<true, 16, 3>, (int)sizeof(gemm_smem<16, 3>));
        set((const void*)contract_gemm3m_c128_kernel<false, 64>, (int)sizeof(gemm3m_smem<64>));
        set((const void*)contract_gemm3m_c128_kernel<true, 64>, (int)sizeof(gemm3m_smem<64>));
        set((const void*)contract_gemm3m_c128_kernel<false, 32>, (int)sizeof(gemm3m_smem<32>));
        set((const void*)contract_gemm3m_c128_kernel<true, 32>, (int)sizeof(gemm3m_smem<32>));
        if (e != cudaSuccess) { set_error("b2_contract_gemm: smem attr: %s", cudaGetErrorString(e)); return 1; }
        b2::mark_done(attr_set);
    }
    // split-K when the tiles alone leave most SMs idle (small M x N, long K: the
    // reduction-heavy pairs of photonic networks).  C must be one dense block so
    // that it can be zero-filled first (the partial tiles are added atomically).
    int ksplit = 1;
    {
        const long long tiles = tiles_m * tiles_n * Bt;
        const long long slots = (three_m && bn == 64 ? 1ll : 2ll) * num_sms();    // resident CTAs
        const long long chunks = (K + BK - 1) / BK;
        long long span = 0;
        for (int i = 0; i < f_k; ++i) span += (long long)(g.dim[i] - 1) * g.sc[i];
        const bool dense_c = span + 1 == Bt * M * N;
        if (tiles * 2 <= slots && chunks >= 32 && (dense_c || accumulate)) {
            long long want = slots / tiles;              // one resident wave
            if (want > chunks / 8) want = chunks / 8;
            if (want > 64) want = 64;
            if (want > 1) ksplit = (int)want;
        } else if (tiles < 6 * slots && chunks >= 128 && (dense_c || accumulate)) {
            // a few waves of long-K tiles: the last, partly filled wave idles much of the
            // GPU (256 tiles on 296 slots: 86 %).  Splitting K multiplies the CTA count;
            // take the smallest split whose last wave is at least 97 % full.
            auto fill = [&](long long ks) {
                const long long ctas = tiles * ks, waves = (ctas + slots - 1) / slots;
                return (double)ctas / (double)(waves * slots);
            };
            const long long max_split = chunks / 64 < 32 ? chunks / 64 : 32;
            double best = fill(1);
            for (long long ks = 2; ks <= max_split && best < 0.97; ++ks)
                if (fill(ks) > best + 0.02) { best = fill(ks); ksplit = (int)ks; }
        }
        if (ksplit > 1 && !accumulate) {
            cudaError_t e = cudaMemsetAsync(c_dev, 0, (size_t)(Bt * M * N) * sizeof(double2),
                                            (cudaStream_t)stream);
            if (e != cudaSuccess) { set_error("b2_contract_gemm: memset: %s", cudaGetErrorString(e)); return 1; }
        }
    }
    set_last_kernel(three_m ? (ksplit > 1 ? "gemm3m_splitk" : "gemm3m") : (ksplit > 1 ? "gemm_splitk" : "gemm"));
    dim3 grid((unsigned)(tiles_m * tiles_n * Bt), 1u, (unsigned)ksplit);
#define B2_GEMM_ARGS                                                                              \
    g, (const double2*)a_dev, (const double2*)b_dev, (double2*)c_dev, M, N, K, (int)tiles_m,      \
        tiles_m * tiles_n, a_kfast, b_kfast, alpha_re, alpha_im, accumulate, ksplit
#define B2_GEMM_LAUNCH(SPLIT, BKV, STV)                                                           \
    contract_gemm_c128_kernel<SPLIT, BKV, STV><<<grid, GEMM_THREADS, smem, (cudaStream_t)stream>>>(B2_GEMM_ARGS)
#define B2_GEMM3M_LAUNCH(SPLIT, BNV)                                                              \
    contract_gemm3m_c128_kernel<SPLIT, BNV><<<grid, BNV * 4, smem, (cudaStream_t)stream>>>(B2_GEMM_ARGS)
    if (three_m) {
        if (bn == 64) { if (ksplit > 1) B2_GEMM3M_LAUNCH(true, 64); else B2_GEMM3M_LAUNCH(false, 64); }
        else          { if (ksplit > 1) B2_GEMM3M_LAUNCH(true, 32); else B2_GEMM3M_LAUNCH(false, 32); }
    } else if (BK == 16) {
        if (ksplit > 1) B2_GEMM_LAUNCH(true, 16, 3); else B2_GEMM_LAUNCH(false, 16, 3);
    } else {
        if (ksplit > 1) B2_GEMM_LAUNCH(true, 8, 4); else B2_GEMM_LAUNCH(false, 8, 4);
    }
#undef B2_GEMM_LAUNCH
#undef B2_GEMM3M_LAUNCH
#undef B2_GEMM_ARGS
    return check_launch("b2_contract_gemm");
}
