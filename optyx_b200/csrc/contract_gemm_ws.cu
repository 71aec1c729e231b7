// Piece (2b), complex128, the many-tile case: persistent, warp-specialised fused
// index-permute + complex GEMM on the FP64 tensor pipe.
//
//   C[b, m, n] (+)= alpha * sum_k A[b, m, k] * B[b, k, n]
//
// (reference: the transpose + tensordot/zgemm quimb/cotengra run per tree node, call
// sites optyx/core/backends.py:514, 539-545.)
//
// Why another kernel: in the barrier-per-chunk kernels of contract_gemm.cu every warp both
// gathers and multiplies, in lock step.  ncu on cfg4's pairs (profiles/r02_*): the DMMA
// pipe idles while all warps compute gather addresses, wait on the block barrier, decode
// reduction offsets or scatter a tile; a short reduction (K <= 128) pays the cold start of
// the cp.async ring and the row / column decode once per 64 x 64 tile (~10 us next to
// 8 - 17 us of math).  Here:
//   * ONE CTA per SM, persistent: it walks work items blockIdx.x, blockIdx.x + gridDim.x, ...
//     (an item = a tile, or a tile x a part of the reduction range when the tiles alone would
//     leave most SMs idle; grouped raster order, batch index fastest, so that the items in
//     flight share operand panels in L2);
//   * two PRODUCER warps, one per operand, own everything that is not math: per-tile row /
//     column offset tables, reduction offsets (a two-level table: K = low x high, the low
//     table lives in shared memory, the high offset is recomputed only when the low index
//     wraps), and all cp.async gathers into a 5-stage ring; stages are handed over with
//     mbarriers (cp.async.mbarrier.arrive), the producers run ahead ACROSS tile boundaries,
//     so the ring never drains between tiles;
//   * 8 MMA warps (warp tile 8 RT x 16) only wait for a stage, issue LDS + DMMA (3M complex
//     multiplication: P1 = ar*br, P2 = ai*bi, P3 = (ar+ai)(br+bi)), release the stage, and
//     scatter their part of the tile at its end (32-byte stores where two columns are
//     adjacent in C) -- each warp on its own clock, so one warp's epilogue overlaps the
//     others' math.
#include <stdlib.h>
#include "b2_common.cuh"

namespace b2 {

namespace ws {

constexpr int BK = 16, STAGES = 5;       // tile: BM x BN = (MW / RT) * 8 RT  x  16 RT  (templates)
constexpr int TAB_GEN = 4;                     // generations of the per-tile C tables
constexpr int KLOW_MAX = 512;                  // entries of the low reduction-offset table

template <int BM, int BN>
struct smem_t {
    double2 a[STAGES][BK][BM + 2];
    double2 b[STAGES][BK][BN + 2];
    long long row_c[TAB_GEN][BM], col_c[TAB_GEN][BN];      // read by the MMA warps
    long long row_a[2][BM], col_b[2][BN];                  // producer only (byte offsets)
    long long klow_a[KLOW_MAX], klow_b[KLOW_MAX];          // byte offsets of k % klow
    uint64_t full[STAGES], empty[STAGES], tab_free[TAB_GEN];
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(s32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s32(bar)) : "memory");
}
// the barrier receives one arrival once every cp.async this thread issued so far has landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void cp16(uint32_t smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void decode_u32(unsigned flat, const b2_gemm_modes& g, int first, int count,
                                           long long& oa, long long& ob, long long& oc) {
    oa = ob = oc = 0;
    for (int m = 0; m < count; ++m) {
        const unsigned d = (unsigned)g.dim[first + m];
        const unsigned q = flat / d, i = flat - q * d;
        flat = q;
        oa += (long long)i * g.sa[first + m];
        ob += (long long)i * g.sb[first + m];
        oc += (long long)i * g.sc[first + m];
    }
}

constexpr int RASTER_GM = 16;
__device__ __forceinline__ void raster(unsigned tile, int tiles_m, int tiles_n, int& tm, int& tn) {
    const unsigned per_group = (unsigned)RASTER_GM * (unsigned)tiles_n;
    const unsigned group = tile / per_group;
    const int first_m = (int)group * RASTER_GM;
    const int gm = tiles_m - first_m < RASTER_GM ? tiles_m - first_m : RASTER_GM;
    const unsigned r = tile - group * per_group;
    tn = (int)(r / (unsigned)gm);
    tm = first_m + (int)(r - (unsigned)tn * (unsigned)gm);
}

// One operand tile chunk (ROWS rows x 16 k, shared layout [k][ROWS + 2]) by the 32 producer
// lanes: ROWS / 2 copies per lane.
//   rows contiguous in the operand: consecutive lanes = consecutive rows of one k
//   K contiguous: a quarter warp = 4 consecutive k x 2 rows (see contract_gemm.cu)
template <int ROWS>
__device__ __forceinline__ void gather_rows(uint32_t dst, const char* base, const long long* row_off,
                                            const long long* klow, int klo, long long khi, int lane,
                                            bool kfast, int kvalid) {
    constexpr int P_ = ROWS + 2;
    if (kfast) {
        const int j = (lane & 3) | ((lane >> 1) & 12);
        const int i0 = (lane >> 2) & 1;
        const char* src = base + klow[klo + j] + khi;
        const int bytes = j < kvalid ? 16 : 0;
        const uint32_t d = dst + (uint32_t)(j * P_ + i0) * 16u;
        #pragma unroll 8
        for (int p = 0; p < ROWS / 2; ++p) cp16(d + (uint32_t)p * 32u, src + row_off[i0 + 2 * p], bytes);
    } else if (ROWS >= 32) {
        constexpr int R = ROWS >= 32 ? ROWS / 32 : 1;   // rows per lane
        const char* sr[R];
        #pragma unroll
        for (int r = 0; r < R; ++r) sr[r] = base + row_off[lane + 32 * r] + khi;
        const uint32_t d = dst + (uint32_t)lane * 16u;
        #pragma unroll 8
        for (int j = 0; j < BK; ++j) {
            const long long ko = klow[klo + j];
            const int bytes = j < kvalid ? 16 : 0;
            #pragma unroll
            for (int r = 0; r < R; ++r) cp16(d + (uint32_t)(j * P_ + 32 * r) * 16u, sr[r] + ko, bytes);
        }
    } else {                                            // 16 rows: the half warps split the k values
        const char* s0 = base + row_off[lane & 15] + khi;
        const uint32_t d = dst + (uint32_t)(lane & 15) * 16u;
        #pragma unroll
        for (int jj = 0; jj < BK / 2; ++jj) {
            const int j = 2 * jj + (lane >> 4);
            cp16(d + (uint32_t)(j * P_) * 16u, s0 + klow[klo + j], j < kvalid ? 16 : 0);
        }
    }
}

template <int RT, int MW>
__device__ __forceinline__ void
ws_body(const b2_gemm_modes& g, const double2* __restrict__ A, const double2* __restrict__ B,
        double2* __restrict__ Cout, int M, int N, int K, int tiles_m, int tiles_n,
        long long total_tiles, int klow, int n_klow_modes, int a_kfast, int b_kfast, double alpha_re,
        double alpha_im, int accumulate, int ksplit) {
    // MW MMA warps as (MW / WN) x WN, warp tile 8 RT x 16; warp MW is the producer
    constexpr int BN = 16 * RT, WN = RT, BM = (MW / WN) * 8 * RT, PITCH = BM + 2, PITCH_B = BN + 2;
    constexpr int MMA_WARPS = MW, THREADS = (MW + 2) * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    smem_t<BM, BN>& sm = *reinterpret_cast<smem_t<BM, BN>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f_m = g.n_batch, f_n = f_m + g.n_m, f_k = f_n + g.n_n;
    const long long nbatch = total_tiles / ((long long)tiles_m * tiles_n);
    const int all_chunks = (K + BK - 1) / BK;
    // work item = (tile, part of the reduction range): ksplit > 1 only when the tiles alone would
    // leave most SMs idle; the partial tiles are then added into the zero-filled C atomically
    const int per_split = (all_chunks + ksplit - 1) / ksplit;
    const long long total_items = total_tiles * ksplit;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 64); mbar_init(&sm.empty[s], MMA_WARPS); }
        for (int t = 0; t < TAB_GEN; ++t) mbar_init(&sm.tab_free[t], MMA_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // low reduction-offset table: k_lo in [0, klow) over the first n_klow_modes K modes
    // (entries up to the next multiple of the chunk are addressed by the zero-filled K tail)
    const int klow_pad = (klow + BK - 1) / BK * BK < KLOW_MAX ? (klow + BK - 1) / BK * BK : KLOW_MAX;
    for (int k = tid; k < klow_pad; k += THREADS) {
        unsigned r = k < klow ? (unsigned)k : 0u;
        long long oa = 0, ob = 0;
        for (int m = 0; m < n_klow_modes; ++m) {
            const unsigned d = (unsigned)g.dim[f_k + m];
            const unsigned q = r / d, i = r - q * d;
            r = q;
            oa += (long long)i * g.sa[f_k + m];
            ob += (long long)i * g.sb[f_k + m];
        }
        sm.klow_a[k] = oa * 16; sm.klow_b[k] = ob * 16;
    }
    __syncthreads();

    if (warp >= MMA_WARPS) {
        // ======================= producer warps: A (warp MW), B (warp MW + 1) =======================
        // Each owns one operand: its offset table per tile, its reduction offsets, its gathers.
        // (One producer warp was the limit of short reductions and narrow column tiles: ~1100
        // instructions of decode + gather issue per 64 x 64 x 16 chunk against 3072 clk of math.)
        const bool is_a = warp == MMA_WARPS;
        const char* base = reinterpret_cast<const char*>(is_a ? (const void*)A : (const void*)B);
        const long long* klow_tab = is_a ? sm.klow_a : sm.klow_b;
        const bool kfast = is_a ? a_kfast != 0 : b_kfast != 0;
        long long chunk_seq = 0;                     // chunks issued so far (ring position)
        int tcount = 0;
        for (unsigned item = blockIdx.x; item < (unsigned)total_items; item += gridDim.x, ++tcount) {
            // (32-bit index arithmetic: the host only launches item counts below 2^31)
            const unsigned tile = ksplit > 1 ? item / (unsigned)ksplit : item;
            const int c0 = ksplit > 1 ? (int)(item - tile * (unsigned)ksplit) * per_split : 0;
            const int nchunks = all_chunks - c0 < per_split ? (all_chunks - c0 > 0 ? all_chunks - c0 : 0)
                                                            : per_split;
            // batch index fastest: batch modes often sit between the row modes in the operand's
            // memory order, so the tiles of all batch values share cache lines -- run them together
            const unsigned rest = tile / (unsigned)nbatch;
            const long long batch = tile - rest * (unsigned)nbatch;
            int tm, tn;
            raster(rest, tiles_m, tiles_n, tm, tn);
            const int gen = tcount % TAB_GEN, pg = tcount & 1;
            // the C tables of generation `gen` are free once the MMA warps finished the
            // tile that used them (TAB_GEN tiles ago)
            if (tcount >= TAB_GEN) mbar_wait(&sm.tab_free[gen], (uint32_t)((tcount / TAB_GEN - 1) & 1));
            long long ba, bb, bc;
            decode_u32((unsigned)batch, g, 0, g.n_batch, ba, bb, bc);
            if (is_a) {
                #pragma unroll
                for (int h = 0; h < BM / 32; ++h) {
                    const int t = lane + 32 * h;
                    long long oa, ob, oc;
                    const int m = tm * BM + t;
                    decode_u32(m < M ? (unsigned)m : 0u, g, f_m, g.n_m, oa, ob, oc);
                    sm.row_a[pg][t] = (ba + oa) * 16; sm.row_c[gen][t] = bc + oc;
                }
            } else {
                #pragma unroll
                for (int h = 0; h < (BN + 31) / 32; ++h) {
                    const int t = lane + 32 * h;
                    if (t < BN) {
                        long long oa, ob, oc;
                        const int n = tn * BN + t;
                        decode_u32(n < N ? (unsigned)n : 0u, g, f_n, g.n_n, oa, ob, oc);
                        sm.col_b[pg][t] = (bb + ob) * 16; sm.col_c[gen][t] = oc;
                    }
                }
            }
            __threadfence_block();
            __syncwarp();
            int klo = (int)(((long long)c0 * BK) % klow);                  // (chunk * 16) % klow
            unsigned khi_idx = (unsigned)(((long long)c0 * BK) / klow);    // (chunk * 16) / klow
            long long khi = 0;
            {
                unsigned r = khi_idx;
                for (int m = n_klow_modes; m < g.n_k; ++m) {
                    const unsigned d = (unsigned)g.dim[f_k + m];
                    const unsigned q = r / d, i = r - q * d;
                    r = q;
                    khi += (long long)i * (is_a ? g.sa[f_k + m] : g.sb[f_k + m]);
                }
                khi *= 16;
            }
            for (int c = c0; c < c0 + nchunks; ++c, ++chunk_seq) {
                const int st = (int)(chunk_seq % STAGES);
                const uint32_t ph = (uint32_t)((chunk_seq / STAGES) & 1);
                mbar_wait(&sm.empty[st], ph ^ 1u);   // passes at once on the first lap
                const int kleft = K - c * BK;
                const int kvalid = kleft < BK ? kleft : BK;
                if (is_a)
                    gather_rows<BM>(s32(&sm.a[st][0][0]), base, sm.row_a[pg], klow_tab, klo, khi, lane,
                                    kfast, kvalid);
                else
                    gather_rows<BN>(s32(&sm.b[st][0][0]), base, sm.col_b[pg], klow_tab, klo, khi, lane,
                                    kfast, kvalid);
                cp_async_arrive(&sm.full[st]);
                klo += BK;
                if (klo >= klow) {                    // klow is a multiple of 16 (or >= K)
                    klo = 0;
                    ++khi_idx;
                    unsigned r = khi_idx;
                    khi = 0;
                    for (int m = n_klow_modes; m < g.n_k; ++m) {
                        const unsigned d = (unsigned)g.dim[f_k + m];
                        const unsigned q = r / d, i = r - q * d;
                        r = q;
                        khi += (long long)i * (is_a ? g.sa[f_k + m] : g.sb[f_k + m]);
                    }
                    khi *= 16;
                }
            }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
    } else {
        // ============================== MMA warps ==============================
        const int wm = warp / WN, wn = warp % WN;
        const int frow = lane >> 2, fk = lane & 3;
        long long chunk_seq = 0;
        int tcount = 0;
        for (unsigned item = blockIdx.x; item < (unsigned)total_items; item += gridDim.x, ++tcount) {
            // (32-bit index arithmetic: the host only launches item counts below 2^31)
            const unsigned tile = ksplit > 1 ? item / (unsigned)ksplit : item;
            const int c0 = ksplit > 1 ? (int)(item - tile * (unsigned)ksplit) * per_split : 0;
            const int nchunks = all_chunks - c0 < per_split ? (all_chunks - c0 > 0 ? all_chunks - c0 : 0)
                                                            : per_split;
            int tm, tn;
            raster(tile / (unsigned)nbatch, tiles_m, tiles_n, tm, tn);
            double p1[RT][2][2], p2[RT][2][2], p3[RT][2][2];
            #pragma unroll
            for (int r = 0; r < RT; ++r)
                #pragma unroll
                for (int q = 0; q < 2; ++q)
                    #pragma unroll
                    for (int h = 0; h < 2; ++h) p1[r][q][h] = p2[r][q][h] = p3[r][q][h] = 0.0;
            // (Loading the fragments one k4 step ahead across the stage hand-over by hand was
            // tried: 24 more live registers push the kernel over the 168 it may use with 9 warps
            // and the spills cost more than the hidden latency -- 6561 -> 6831 us on cfg4's
            // dominant pair.)
            for (int c = 0; c < nchunks; ++c, ++chunk_seq) {
                const int st = (int)(chunk_seq % STAGES);
                mbar_wait(&sm.full[st], (uint32_t)((chunk_seq / STAGES) & 1));
                const double2* ap = &sm.a[st][fk][wm * (8 * RT) + frow];
                const double2* bp = &sm.b[st][fk][wn * 16 + frow];
                #pragma unroll
                for (int k4 = 0; k4 < BK; k4 += 4) {
                    double2 af[RT], bf[2];
                    double as[RT], bs[2];
                    #pragma unroll
                    for (int r = 0; r < RT; ++r) af[r] = ap[k4 * PITCH + r * 8];
                    #pragma unroll
                    for (int q = 0; q < 2; ++q) bf[q] = bp[k4 * PITCH_B + q * 8];
                    #pragma unroll
                    for (int r = 0; r < RT; ++r) as[r] = af[r].x + af[r].y;
                    #pragma unroll
                    for (int q = 0; q < 2; ++q) bs[q] = bf[q].x + bf[q].y;
                    #pragma unroll
                    for (int r = 0; r < RT; ++r)
                        #pragma unroll
                        for (int q = 0; q < 2; ++q) dmma(p1[r][q][0], p1[r][q][1], af[r].x, bf[q].x);
                    #pragma unroll
                    for (int r = 0; r < RT; ++r)
                        #pragma unroll
                        for (int q = 0; q < 2; ++q) dmma(p2[r][q][0], p2[r][q][1], af[r].y, bf[q].y);
                    #pragma unroll
                    for (int r = 0; r < RT; ++r)
                        #pragma unroll
                        for (int q = 0; q < 2; ++q) dmma(p3[r][q][0], p3[r][q][1], as[r], bs[q]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[st]);
            }
            // ---- scatter this warp's (8 RT) x 16 part of the tile through the C permutation
            const int gen = tcount % TAB_GEN;
            const int m0 = tm * BM, n0 = tn * BN;
            #pragma unroll
            for (int r = 0; r < RT; ++r) {
                const int i = wm * (8 * RT) + r * 8 + frow;
                if (m0 + i >= M) continue;
                const long long ro = sm.row_c[gen][i];
                #pragma unroll
                for (int q = 0; q < 2; ++q) {
                    // a lane holds two ADJACENT columns (2 fk, 2 fk + 1) of its row: when they are
                    // adjacent in C too (the usual case: the column modes are C's fastest), one
                    // 32-byte store (STG.256) writes a whole sector instead of two half sectors
                    const int j = wn * 16 + q * 8 + fk * 2;
                    if (n0 + j >= N) continue;
                    double2 v[2];
                    #pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double re = p1[r][q][h] - p2[r][q][h];
                        const double im = p3[r][q][h] - p1[r][q][h] - p2[r][q][h];
                        v[h].x = alpha_re * re - alpha_im * im;
                        v[h].y = alpha_re * im + alpha_im * re;
                    }
                    const long long c0off = sm.col_c[gen][j], c1off = sm.col_c[gen][j + 1];
                    double2* dst = Cout + ro + c0off;
                    const bool second = n0 + j + 1 < N;
                    if (ksplit > 1) {
                        atomicAdd(&dst->x, v[0].x);
                        atomicAdd(&dst->y, v[0].y);
                        if (second) {
                            double2* d1 = Cout + ro + c1off;
                            atomicAdd(&d1->x, v[1].x);
                            atomicAdd(&d1->y, v[1].y);
                        }
                    } else if (second && c1off == c0off + 1 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
                        if (accumulate) {
                            double o0, o1, o2, o3;
                            asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                                         : "=d"(o0), "=d"(o1), "=d"(o2), "=d"(o3) : "l"(dst));
                            v[0].x += o0; v[0].y += o1; v[1].x += o2; v[1].y += o3;
                        }
                        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "d"(v[0].x),
                                     "d"(v[0].y), "d"(v[1].x), "d"(v[1].y) : "memory");
                    } else {
                        if (accumulate) { const double2 o = *dst; v[0].x += o.x; v[0].y += o.y; }
                        *dst = v[0];
                        if (second) {
                            double2* d1 = Cout + ro + c1off;
                            if (accumulate) { const double2 o = *d1; v[1].x += o.x; v[1].y += o.y; }
                            *d1 = v[1];
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tab_free[gen]);
        }
    }
}

#define B2_WS_PARAMS                                                                              \
    const __grid_constant__ b2_gemm_modes g, const double2* __restrict__ A,                       \
        const double2* __restrict__ B, double2* __restrict__ Cout, int M, int N, int K, int tiles_m, \
        int tiles_n, long long total_tiles, int klow, int n_klow_modes, int a_kfast, int b_kfast,  \
        double alpha_re, double alpha_im, int accumulate, int ksplit
#define B2_WS_ARGS                                                                                \
    g, A, B, Cout, M, N, K, tiles_m, tiles_n, total_tiles, klow, n_klow_modes, a_kfast, b_kfast,   \
        alpha_re, alpha_im, accumulate, ksplit

// 8 MMA warps + the two producers = 320 threads.  (A 12 + 1 warp, 96-row variant was tried: with
// 13 warps one scheduler hosts 4 of them, which caps the kernel at 128 registers per thread
// -- the 3M accumulators then spill.)
template <int RT>
__global__ void __launch_bounds__(320, 1) contract_gemm_ws_c128_kernel(B2_WS_PARAMS) {
    ws_body<RT, 8>(B2_WS_ARGS);
}

}  // namespace ws

template <int RT, int MW>
static int launch_ws_rt(const b2_gemm_modes& g, const void* a, const void* b, void* c, long long Bt,
                        long long M, long long N, long long K, long long klow, int n_low, int a_kfast,
                        int b_kfast, double ar, double ai, int accumulate, int ksplit, cudaStream_t s) {
    using namespace ws;
    constexpr int BN = 16 * RT, BM = (MW / RT) * 8 * RT;
    const int sms = num_sms();
    if (sms <= 0) return 1;
    const long long tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
    const long long total = tiles_m * tiles_n * Bt;
    const int smem = (int)sizeof(smem_t<BM, BN>);
    static b2::once_flags attr_set;
    if (!b2::is_done(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(contract_gemm_ws_c128_kernel<RT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("b2_contract_gemm(ws): smem attr: %s", cudaGetErrorString(e)); return 1; }
        b2::mark_done(attr_set);
    }
    const long long items = total * ksplit;
    if (items >= (1ll << 31) || tiles_n * RASTER_GM >= (1ll << 31)) {
        set_error("b2_contract_gemm(ws): too many tiles (%lld)", items);
        return 1;
    }
    const long long grid = items < sms ? items : sms;
    if (ksplit > 1 && !accumulate) {
        cudaError_t e = cudaMemsetAsync(c, 0, (size_t)(Bt * M * N) * sizeof(double2), s);
        if (e != cudaSuccess) { set_error("b2_contract_gemm(ws): memset: %s", cudaGetErrorString(e)); return 1; }
    }
    contract_gemm_ws_c128_kernel<RT><<<(unsigned)grid, (MW + 2) * 32, smem, s>>>(
        g, (const double2*)a, (const double2*)b, (double2*)c, (int)M, (int)N, (int)K, (int)tiles_m,
        (int)tiles_n, total, (int)klow, n_low, a_kfast, b_kfast, ar, ai, accumulate, ksplit);
    return check_launch("b2_contract_gemm(ws)");
}

// returns handled = false when the pair does not fit this kernel (the caller falls
// back to the barrier-per-chunk kernels of contract_gemm.cu).
int launch_gemm_ws(const b2_gemm_modes& g_in, const void* a, const void* b, void* c, long long Bt,
                   long long M, long long N, long long K, int a_kfast, int b_kfast, double ar,
                   double ai, int accumulate, cudaStream_t s, bool& handled) {
    using namespace ws;
    handled = false;
    // The planner merges stride-compatible reduction modes; the two-level offset table needs
    // a leading group of at most KLOW_MAX values, so a long merged mode is split back into
    // (low, high) factors here: index = lo + f * hi, strides (s, f * s) -- the same addresses.
    b2_gemm_modes g = g_in;
    {
        const int f_k0 = g.n_batch + g.n_m + g.n_n;
        long long low = 1;
        for (int i = 0; i < g.n_k; ++i) {
            const long long d = g.dim[f_k0 + i];
            if (low * d <= KLOW_MAX) { low *= d; continue; }
            // largest factor of d that keeps the low group within the table and, if it
            // can, makes its extent a multiple of the K chunk
            long long best = 0;
            for (long long f = KLOW_MAX / low; f >= 2; --f)
                if (d % f == 0 && ((low * f) % BK == 0 || best == 0)) {
                    if ((low * f) % BK == 0) { best = f; break; }
                    best = f;
                }
            const int total = f_k0 + g.n_k;
            if (best >= 2 && total < B2_MAX_MODES) {
                for (int j = total; j > f_k0 + i + 1; --j) {         // make room for the high factor
                    g.dim[j] = g.dim[j - 1]; g.sa[j] = g.sa[j - 1]; g.sb[j] = g.sb[j - 1]; g.sc[j] = g.sc[j - 1];
                }
                g.dim[f_k0 + i] = (int)best;
                g.dim[f_k0 + i + 1] = (int)(d / best);
                g.sa[f_k0 + i + 1] = g.sa[f_k0 + i] * best;
                g.sb[f_k0 + i + 1] = g.sb[f_k0 + i] * best;
                g.sc[f_k0 + i + 1] = 0;
                ++g.n_k;
            }
            break;
        }
    }
    if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31) || Bt >= (1ll << 31)) return 0;
    const int bn = N >= 48 ? 64 : N > 16 ? 32 : 16;
    const int bm = 64;
    const long long tiles = ((M + bm - 1) / bm) * ((N + bn - 1) / bn) * Bt;
    // few tiles over a long reduction: split the reduction range so that every SM gets ~2 items
    // (C must be one dense block: it is zero-filled and the partial tiles are added atomically)
    int ksplit = 1;
    if (tiles * 2 <= num_sms() && K >= 256) {
        const int f_free = g.n_batch + g.n_m + g.n_n;
        long long span = 0;
        for (int i = 0; i < f_free; ++i) span += (long long)(g.dim[i] - 1) * g.sc[i];
        if (span + 1 != Bt * M * N && !accumulate) return 0;
        const long long chunks = (K + BK - 1) / BK;
        long long want = 2ll * num_sms() / tiles;             // items just under two full rounds
        if (want > chunks / 4) want = chunks / 4;            // >= 4 chunks per item
        if (want > 1) ksplit = (int)want;
    }
    // two-level reduction offsets: the low table covers the leading K modes, up to
    // KLOW_MAX entries; with more than one level its extent must be a multiple of the
    // K chunk so that a chunk never straddles two high indices
    const int f_k = g.n_batch + g.n_m + g.n_n;
    long long klow = 1;
    int n_low = 0;
    while (n_low < g.n_k && klow * g.dim[f_k + n_low] <= KLOW_MAX) klow *= g.dim[f_k + n_low++];
    if (n_low < g.n_k && klow % BK != 0) {
        // try a shorter prefix whose extent is a multiple of 16
        while (n_low > 0 && klow % BK != 0) klow /= g.dim[f_k + --n_low];
        if (n_low == 0 || klow % BK != 0) return 0;
    }
    handled = true;
    set_last_kernel(ksplit > 1 ? "gemm_ws_splitk" : bn == 64 ? "gemm_ws" : bn == 32 ? "gemm_ws32" : "gemm_ws16");
    if (bn == 64) return launch_ws_rt<4, 8>(g, a, b, c, Bt, M, N, K, klow, n_low, a_kfast, b_kfast, ar, ai, accumulate, ksplit, s);
    if (bn == 32) return launch_ws_rt<2, 8>(g, a, b, c, Bt, M, N, K, klow, n_low, a_kfast, b_kfast, ar, ai, accumulate, ksplit, s);
    return launch_ws_rt<1, 8>(g, a, b, c, Bt, M, N, K, klow, n_low, a_kfast, b_kfast, ar, ai, accumulate, ksplit, s);
}

}  // namespace b2
