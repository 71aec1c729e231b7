// Piece (2b), complex64: large dense pairwise contraction as a fused
// index-permute + complex GEMM on the 5th-generation tensor cores (tcgen05),
// FP32 accuracy by split-precision TF32 emulation.
//
//   C[b, m, n] (+)= alpha * sum_k A[b, m, k] * B[b, k, n]        (complex64)
//
// (reference: the transpose + tensordot/zgemm quimb/cotengra run per tree node,
// call sites optyx/core/backends.py:514, 539-545; the complex64 mode is the
// north star's reduced-precision variant, tolerance 1e-5.)
//
// Complex -> real embedding.  With interleaved (re, im) storage a complex GEMM
// M x N x K is ONE real GEMM  M x 2N x 2K:
//     A' [m][2k], A'[m][2k+1]      =  ar, ai                (A as it is stored)
//     B' [2n  ][2k], B'[2n  ][2k+1] =  br, -bi              (row of C_re)
//     B' [2n+1][2k], B'[2n+1][2k+1] =  bi,  br              (row of C_im)
//     D  [m][2n], D[m][2n+1]        =  C_re, C_im           (C as it is stored)
// Split precision: x = hi + lo with hi = x truncated to TF32 (the tensor core
// reads only the top 19 bits of each 32-bit operand) and lo = x - hi (exact in
// FP32).  a*b ~= ahi*bhi + ahi*blo + alo*bhi, accumulated in FP32 in tensor
// memory: 3 tcgen05.mma.kind::tf32 per k-step, relative error ~2^-21.
//
// Data flow per CTA (tile 128 x NT complex, K chunk 16 complex = 32 real), warp
// specialised, no block-wide barrier in the loop:
//   8 converter warps : gather the chunk's A rows / B columns through the per-mode
//                 strides (the index permutation; no transposes in HBM) with 8-byte
//                 cp.async, 3 stages deep; split hi/lo in registers;
//                 A -> TENSOR MEMORY with tcgen05.st.32x32b (lane = row, 16 columns
//                 = the thread's 8 complex k values), B -> shared memory in the
//                 canonical K-major no-swizzle UMMA layout [k-group][row][16 B];
//                 fence, arrive on the stage's `full` mbarrier
//   MMA warp    : one thread issues 12 tcgen05.mma.kind::tf32 per chunk (4 k-steps x
//                 3 split products; A operand from TMEM, B from a shared-memory
//                 descriptor), D in TMEM (128 lanes x 2NT fp32 columns, two sets),
//                 tcgen05.commit -> the stage's `empty` mbarrier
//   converters  : every 4 chunks the finished TMEM set is drained with
//                 tcgen05.ld.32x32b into FP32 register accumulators (two-level
//                 accumulation, see below); epilogue = alpha scaling + scatter
//                 through the C permutation.
// Tensor memory map (512 or 256 columns): [D set 0 | D set 1 | A hi/lo stage 0 | A
// hi/lo stage 1].
#include "b2_common.cuh"

namespace b2 {

constexpr int G4_CONV = 256;                      // converter threads (8 warps)
constexpr int G4_THREADS = G4_CONV + 32;         // + the MMA warp
constexpr int G4_BM = 128, G4_KC = 16;
constexpr int G4_RAW = 3;                        // cp.async stages of raw complex64 operands
constexpr int G4_STAGES = 2;                     // UMMA operand stages (hi/lo split tiles); 4 stages measured
                                                 // slower (r02: 124 -> 119 TFLOP/s): the converter warps, not the
                                                 // hand-over latency, bound the kernel (shared-memory pipe 65 % busy)
constexpr int G4_KG = G4_KC / 2;                 // 16-byte k-groups (2 complex) per chunk
constexpr int G4_SEG = 4;                        // chunks per TMEM accumulation segment

__device__ __forceinline__ uint32_t g4_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void g4_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(g4_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void g4_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(g4_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void g4_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                 ::"r"(g4_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void g4_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand in tensor memory (128 lanes x 8 columns of 32-bit), B from shared memory
__device__ __forceinline__ void g4_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void g4_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
          "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
          "r"(v[14]), "r"(v[15]) : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor: 8 x 16 B core matrices,
// `lbo` bytes between core matrices along K, `sbo` bytes along M/N.
__device__ __forceinline__ uint64_t g4_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);          // version = 1 (Blackwell)
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void g4_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
// 8-byte asynchronous copy global -> shared, zero-filled when !ok
__device__ __forceinline__ void g4_cp8(void* smem, const void* gmem, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(g4_smem_u32(smem)),
                 "l"(gmem), "r"(ok ? 8 : 0));
}

__device__ __forceinline__ void g4_decode3(long long flat, const b2_gemm_modes& g, int first,
                                           int count, long long& oa, long long& ob, long long& oc) {
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

__device__ __forceinline__ float g4_hi(float x) {
    return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}

// B rows per k-group are padded by one 16 B slot (the descriptor's leading byte
// offset is then 16 B off a multiple of 128 B: conflict-free tile stores)

template <int NT>
struct g4_smem {
    static constexpr int ROWS_B = 2 * NT + 1;
    static constexpr int LOADS = NT * G4_KC / G4_CONV;               // 8 B copies of B per thread, chunk
    // UMMA B operand tiles: [stage][hi/lo][k-group][row][4 floats]  (A lives in TMEM)
    float4 b[G4_STAGES][2][G4_KG][ROWS_B];
    // cp.async landing zone of A: one 128 x 16 complex tile per raw stage, [row][k]
    // (pitch 17) when K is contiguous in the operand, else [k][row]
    float2 raw_a[G4_RAW][G4_BM * (G4_KC + 1)];
    // cp.async landing slots of B, private per thread: [raw stage][copy][thread]
    float2 raw[G4_RAW][LOADS][G4_CONV];
    long long row_a[G4_BM], row_c[G4_BM], col_b[NT], col_c[NT];
    long long k_a[4][G4_KC], k_b[4][G4_KC];
    uint64_t full[G4_STAGES], empty[G4_STAGES];
    uint32_t tmem_base;
};

template <int NT>
__global__ void __launch_bounds__(G4_THREADS, 1)
contract_gemm_c64_kernel(const __grid_constant__ b2_gemm_modes g, const float2* __restrict__ A,
                         const float2* __restrict__ B, float2* __restrict__ Cout, long long M,
                         long long N, long long K, int tiles_m, long long tiles_mn, int a_kfast,
                         int b_kfast, float alpha_re, float alpha_im, int accumulate) {
    extern __shared__ __align__(1024) unsigned char g4_raw[];
    g4_smem<NT>& sm = *reinterpret_cast<g4_smem<NT>*>(g4_raw);
    constexpr int NP = 2 * NT;                               // real columns of D
    // tensor memory columns: [D set 0 | D set 1 | A stage 0 hi, lo | A stage 1 hi, lo]
    constexpr int KR = 2 * G4_KC;                            // real K columns of a chunk (32)
    constexpr int TM_A0 = 2 * NP;
    constexpr int TM_COLS = (2 * NP + 2 * G4_STAGES * KR) <= 256 ? 256 : 512;
    constexpr int A_LOADS = G4_BM * G4_KC / G4_CONV;         // complex A elements per thread, chunk (8)
    constexpr int B_ITEMS = NT * G4_KG / G4_CONV;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // grid.x = batch x tiles (a batch extent can exceed the 65535 limit of grid.y)
    const long long tile = blockIdx.x % tiles_mn, batch_id = blockIdx.x / tiles_mn;
    const long long m0 = (tile % tiles_m) * G4_BM, n0 = (tile / tiles_m) * NT;
    const int f_m = g.n_batch, f_n = f_m + g.n_m, f_k = f_n + g.n_n;

    {
        long long ba, bb, bc;
        g4_decode3(batch_id, g, 0, g.n_batch, ba, bb, bc);
        A += ba; B += bb; Cout += bc;
    }
    if (tid < G4_BM) {
        long long oa, ob, oc;
        const long long m = m0 + tid;
        g4_decode3(m < M ? m : 0, g, f_m, g.n_m, oa, ob, oc);
        sm.row_a[tid] = oa; sm.row_c[tid] = oc;
    } else if (tid - G4_BM < NT) {
        const int t = tid - G4_BM;
        long long oa, ob, oc;
        const long long n = n0 + t;
        g4_decode3(n < N ? n : 0, g, f_n, g.n_n, oa, ob, oc);
        sm.col_b[t] = ob; sm.col_c[t] = oc;
    }
    const int nchunks = (int)((K + G4_KC - 1) / G4_KC);
    // reduction offsets of one chunk (32-bit mixed-radix decode; K < 2^31) by 16
    // lanes of converter warp `w`
    auto decode_k = [&](int chunk, int w) {
        const int t = tid - w * 32;
        if (t >= 0 && t < G4_KC && chunk < nchunks) {
            unsigned k = (unsigned)chunk * G4_KC + (unsigned)t;
            if ((long long)k >= K) k = 0;
            long long oa = 0, ob = 0;
            for (int m = 0; m < g.n_k; ++m) {
                const unsigned d = (unsigned)g.dim[f_k + m];
                const unsigned qq = k / d, i = k - qq * d;
                k = qq;
                oa += (long long)i * g.sa[f_k + m];
                ob += (long long)i * g.sb[f_k + m];
            }
            sm.k_a[chunk & 3][t] = oa;
            sm.k_b[chunk & 3][t] = ob;
        }
    };
    decode_k(0, 4);
    decode_k(1, 5);
    decode_k(2, 6);
    __syncwarp();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     ::"r"(g4_smem_u32(&sm.tmem_base)), "r"(TM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 32) {
        for (int s = 0; s < G4_STAGES; ++s) {
            g4_mbar_init(&sm.full[s], G4_CONV);              // every converter thread arrives
            g4_mbar_init(&sm.empty[s], 1);                   // one tcgen05.commit
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = sm.tmem_base;

    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 2NT, M = 128
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) |
                               ((uint32_t)(G4_BM >> 4) << 24);
    constexpr uint32_t b_lbo = g4_smem<NT>::ROWS_B * 16, sbo = 128;
    constexpr uint32_t B_HALF_BYTES = G4_KG * g4_smem<NT>::ROWS_B * 16, B_STAGE_BYTES = 2 * B_HALF_BYTES;

    if (warp == G4_CONV / 32) {
        // =================== MMA warp: one thread issues ===================
        if (lane == 0) {
            // B descriptor of stage 0 (hi tile); the lo tile, the other stage and the
            // k-steps are constant offsets of the 14-bit (address >> 4) field.  A is
            // read from tensor memory: 128 lanes x 8 columns per MMA.
            const uint64_t desc_b_hi = g4_desc(g4_smem_u32(&sm.b[0][0][0][0]), b_lbo, sbo);
            #pragma unroll 1
            for (int c = 0; c < nchunks; ++c) {
                const int st = c % G4_STAGES;
                g4_mbar_wait(&sm.full[st], (uint32_t)((c / G4_STAGES) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t d = tmem_d + (uint32_t)(((c / G4_SEG) & 1) * NP);
                const bool first = (c % G4_SEG) == 0;
                const uint32_t a_hi = tmem_d + (uint32_t)(TM_A0 + st * 2 * KR), a_lo = a_hi + KR;
                const uint64_t sb = (uint64_t)((uint32_t)st * (B_STAGE_BYTES >> 4));
                #pragma unroll
                for (int ks = 0; ks < G4_KG / 2; ++ks) {      // UMMA K = 8 tf32 = 2 k-groups
                    const uint64_t dbh = desc_b_hi + sb + (uint64_t)(ks * 2 * b_lbo >> 4);
                    const uint64_t dbl = dbh + (B_HALF_BYTES >> 4);
                    g4_mma_tf32_ts(d, a_hi + 8 * ks, dbh, idesc, (first && ks == 0) ? 0u : 1u);
                    g4_mma_tf32_ts(d, a_hi + 8 * ks, dbl, idesc, 1u);
                    g4_mma_tf32_ts(d, a_lo + 8 * ks, dbh, idesc, 1u);
                }
                g4_commit(&sm.empty[st]);     // frees the stage; also "chunk c is in TMEM"
            }
        }
    } else {
        // =================== converter warps (256 threads) ===================
        // Two-level accumulation.  The tensor core truncates its FP32 accumulator
        // on every add (measured: relative error grows ~4e-8 per accumulation), so
        // the TMEM accumulator only ever sums one SEGMENT of G4_SEG chunks (48
        // MMAs); the segments are added in FP32 registers (round-to-nearest).  Two
        // TMEM sets alternate: segment s+1 accumulates while segment s is drained.
        const int q = warp & 3, half = warp >> 2;           // TMEM lane quarter, column half
        constexpr int COLS_PER_WARP = NP / 2;
        float acc[COLS_PER_WARP];
        #pragma unroll
        for (int j = 0; j < COLS_PER_WARP; ++j) acc[j] = 0.f;
        int next_drain = 0;                                 // first segment not yet drained
        auto drain = [&]() {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t base = tmem_d + (uint32_t)((next_drain & 1) * NP) +
                                  ((uint32_t)(q * 32) << 16) + (uint32_t)(half * COLS_PER_WARP);
            #pragma unroll
            for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) {
                uint32_t v[16];
                g4_ld16(base + (uint32_t)c0, v);
                #pragma unroll
                for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            ++next_drain;
        };

        // Work items of this thread, the same for every chunk.  Each thread converts
        // exactly the elements it copied itself, so the copies need no block-wide
        // barrier.  A: thread = (row q*32 + lane, K half): a warp may only write its
        // own quarter of the TMEM lanes, so the row is fixed by (warp % 4, lane) and
        // the two warps of a quarter split the chunk's 16 k values.  B: K contiguous
        // in the operand -> 8 lanes cover the 8 k-pairs of a row, else lanes = rows.
        const int a_row = q * 32 + lane, a_k0 = half * (G4_KC / 2);
        // A is copied with a coalesced mapping (16 lanes cover a row's 16 k values,
        // or lanes = consecutive rows) into the shared raw tile and read back by the
        // row's owner after the per-chunk converter barrier
        const int a_pr = a_kfast ? G4_KC + 1 : 1, a_pk = a_kfast ? 1 : G4_BM;
        const int c_r0 = a_kfast ? tid / G4_KC : tid % G4_BM, c_rstep = a_kfast ? G4_CONV / G4_KC : 0;
        const int c_k0 = a_kfast ? tid % G4_KC : tid / G4_BM, c_kstep = a_kfast ? 0 : G4_CONV / G4_BM;
        const uint32_t raw_a0 = g4_smem_u32(&sm.raw_a[0][0]) + 8u * (uint32_t)(c_r0 * a_pr + c_k0 * a_pk);
        const uint32_t a_dstep = 8u * (uint32_t)(c_rstep * a_pr + c_kstep * a_pk);
        constexpr uint32_t RAW_A_STAGE = G4_BM * (G4_KC + 1) * 8;
        const int b_r0 = b_kfast ? tid / G4_KG : tid % NT, b_rstep = b_kfast ? G4_CONV / G4_KG : 0;
        const int b_p0 = b_kfast ? tid % G4_KG : (tid / NT) * B_ITEMS, b_pstep = b_kfast ? 0 : 1;
        const uint32_t raw0 = g4_smem_u32(&sm.raw[0][0][tid]);
        constexpr uint32_t RAW_SLOT = G4_CONV * 8, RAW_STAGE = g4_smem<NT>::LOADS * RAW_SLOT;

        // asynchronous gather of chunk c through the per-mode strides.  Per copy: one
        // table load + one 64-bit add (r01: ~28 instructions of index arithmetic and
        // predicates per copy made the converter warps issue-bound -- tensor pipe 28 %
        // active).  Rows / columns beyond the edge read row 0 (valid memory, never
        // stored); only the K tail is zero-filled.
        auto issue = [&](int c) {
            if (c < nchunks) {
                const int kt = c & 3;
                const uint32_t dst = raw0 + (uint32_t)(c % G4_RAW) * RAW_STAGE;
                const uint32_t dst_a = raw_a0 + (uint32_t)(c % G4_RAW) * RAW_A_STAGE;
                const long long kleft = K - (long long)c * G4_KC;
                const int krem = (int)(kleft < G4_KC ? kleft : G4_KC);
                if (a_kfast) {
                    const float2* src = A + sm.k_a[kt][c_k0];
                    const int bytes = c_k0 < krem ? 8 : 0;
                    #pragma unroll
                    for (int p = 0; p < A_LOADS; ++p)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(
                                         dst_a + (uint32_t)p * a_dstep),
                                     "l"(src + sm.row_a[c_r0 + p * (G4_CONV / G4_KC)]), "r"(bytes));
                } else {
                    const float2* src = A + sm.row_a[c_r0];
                    #pragma unroll
                    for (int p = 0; p < A_LOADS; ++p) {
                        const int k = c_k0 + p * (G4_CONV / G4_BM);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(
                                         dst_a + (uint32_t)p * a_dstep),
                                     "l"(src + sm.k_a[kt][k]), "r"(k < krem ? 8 : 0));
                    }
                }
                if (b_kfast) {
                    const float2* s0 = B + sm.k_b[kt][2 * b_p0];
                    const float2* s1 = B + sm.k_b[kt][2 * b_p0 + 1];
                    const int y0 = 2 * b_p0 < krem ? 8 : 0, y1 = 2 * b_p0 + 1 < krem ? 8 : 0;
                    #pragma unroll
                    for (int p = 0; p < B_ITEMS; ++p) {
                        const long long co = sm.col_b[b_r0 + p * (G4_CONV / G4_KG)];
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(
                                         dst + (uint32_t)(2 * p) * RAW_SLOT), "l"(s0 + co), "r"(y0));
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(
                                         dst + (uint32_t)(2 * p + 1) * RAW_SLOT), "l"(s1 + co), "r"(y1));
                    }
                } else {
                    const float2* src = B + sm.col_b[b_r0];
                    #pragma unroll
                    for (int p = 0; p < B_ITEMS; ++p) {
                        const int kp = b_p0 + p;
                        #pragma unroll
                        for (int h = 0; h < 2; ++h)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(
                                             dst + (uint32_t)(2 * p + h) * RAW_SLOT),
                                         "l"(src + sm.k_b[kt][2 * kp + h]), "r"(2 * kp + h < krem ? 8 : 0));
                    }
                }
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        };
        issue(0);
        issue(1);

        #pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
            const int st = c % G4_STAGES, rs = c % G4_RAW;
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");     // own copies of chunk c landed
            // converter barrier: everyone's copies of chunk c are visible (the A tile is
            // read by the rows' owners), the reduction offsets of chunk c + 2 are
            // written, and the raw stage of chunk c - 1 is free for chunk c + 2
            asm volatile("bar.sync 1, %0;\n" ::"n"(G4_CONV) : "memory");
            issue(c + 2);
            // the UMMA stage is free once the MMAs of chunk c - 2 have completed; if
            // that chunk closed a segment, its TMEM set is complete: drain it
            if (c >= G4_STAGES) {
                g4_mbar_wait(&sm.empty[st], (uint32_t)((c / G4_STAGES - 1) & 1));
                if ((c - G4_STAGES) % G4_SEG == G4_SEG - 1) drain();
            }
            // ---- A: split hi/lo in registers and store straight into tensor memory
            //      (lane = row, 16 columns = this thread's 8 complex k values)
            {
                uint32_t hi[16], lo[16];
                #pragma unroll
                for (int j = 0; j < A_LOADS; ++j) {
                    const float2 v = sm.raw_a[rs][a_row * a_pr + (a_k0 + j) * a_pk];
                    const float hx = g4_hi(v.x), hy = g4_hi(v.y);
                    hi[2 * j] = __float_as_uint(hx); hi[2 * j + 1] = __float_as_uint(hy);
                    lo[2 * j] = __float_as_uint(v.x - hx); lo[2 * j + 1] = __float_as_uint(v.y - hy);
                }
                const uint32_t ta = tmem_d + ((uint32_t)(q * 32) << 16) +
                                    (uint32_t)(TM_A0 + st * 2 * KR + half * (KR / 2));
                g4_st16(ta, hi);
                g4_st16(ta + KR, lo);
            }
            // ---- B: split hi/lo and store in the UMMA K-major shared-memory layout
            #pragma unroll
            for (int p = 0; p < B_ITEMS; ++p) {
                const int n = b_r0 + p * b_rstep, kp = b_p0 + p * b_pstep;
                const float2 v0 = sm.raw[rs][2 * p][tid];
                const float2 v1 = sm.raw[rs][2 * p + 1][tid];
                const float4 xr = make_float4(v0.x, -v0.y, v1.x, -v1.y);      // row of C_re
                const float4 xi = make_float4(v0.y, v0.x, v1.y, v1.x);        // row of C_im
                const float4 hr = make_float4(g4_hi(xr.x), g4_hi(xr.y), g4_hi(xr.z), g4_hi(xr.w));
                const float4 hi = make_float4(g4_hi(xi.x), g4_hi(xi.y), g4_hi(xi.z), g4_hi(xi.w));
                sm.b[st][0][kp][2 * n] = hr;
                sm.b[st][0][kp][2 * n + 1] = hi;
                sm.b[st][1][kp][2 * n] = make_float4(xr.x - hr.x, xr.y - hr.y, xr.z - hr.z, xr.w - hr.w);
                sm.b[st][1][kp][2 * n + 1] = make_float4(xi.x - hi.x, xi.y - hi.y, xi.z - hi.z, xi.w - hi.w);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(g4_smem_u32(&sm.full[st]))
                         : "memory");
            // reduction offsets of chunk c + 3, visible to everyone after the next
            // converter barrier (the MMA warp never touches the tables)
            decode_k(c + 3, c & 7);
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        // ---- all MMAs done (a commit tracks every earlier MMA of the issuing
        //      thread): drain the segments still in TMEM (at most two)
        {
            const int last = nchunks - 1;
            g4_mbar_wait(&sm.empty[last % G4_STAGES], (uint32_t)((last / G4_STAGES) & 1));
            const int nseg = (nchunks + G4_SEG - 1) / G4_SEG;
            while (next_drain < nseg) drain();
        }
        // ---- epilogue: thread = row q*32 + lane, columns = (re, im) pairs of its half
        {
            const int row = q * 32 + lane;
            const bool okm = m0 + row < M;
            float2* crow = Cout + sm.row_c[row];
            #pragma unroll
            for (int j = 0; j < COLS_PER_WARP / 2; ++j) {
                const int n = half * (COLS_PER_WARP / 2) + j;
                if (okm && n0 + n < N) {
                    const float re = acc[2 * j], im = acc[2 * j + 1];
                    float2 r;
                    r.x = alpha_re * re - alpha_im * im;
                    r.y = alpha_re * im + alpha_im * re;
                    float2* dst = crow + sm.col_c[n];
                    if (accumulate) { const float2 o = *dst; r.x += o.x; r.y += o.y; }
                    *dst = r;
                }
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(TM_COLS));
    }
}

template <int NT>
static int launch_c64_nt(const b2_gemm_modes& g, const float2* a, const float2* b, float2* c,
                         long long Bt, long long M, long long N, long long K, int a_kfast,
                         int b_kfast, double ar, double ai, int accumulate, cudaStream_t s) {
    const long long tiles_m = (M + G4_BM - 1) / G4_BM, tiles_n = (N + NT - 1) / NT;
    if (tiles_m * tiles_n * Bt > 2147483647ll) {
        set_error("b2_contract_gemm(c64): grid too large (tiles=%lld batch=%lld)", tiles_m * tiles_n, Bt);
        return 1;
    }
    const int smem = (int)sizeof(g4_smem<NT>) + 1024;
    static b2::once_flags attr_set;
    if (!b2::is_done(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(contract_gemm_c64_kernel<NT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("b2_contract_gemm(c64): smem attr: %s", cudaGetErrorString(e)); return 1; }
        b2::mark_done(attr_set);
    }
    dim3 grid((unsigned)(tiles_m * tiles_n * Bt));
    contract_gemm_c64_kernel<NT><<<grid, G4_THREADS, smem, s>>>(
        g, a, b, c, M, N, K, (int)tiles_m, tiles_m * tiles_n, a_kfast, b_kfast, (float)ar,
        (float)ai, accumulate);
    return check_launch("b2_contract_gemm(c64 tcgen05)");
}

int launch_gemm_c64(const b2_gemm_modes& g, const void* a_dev, const void* b_dev, void* c_dev,
                    double ar, double ai, int accumulate, cudaStream_t s) {
    long long Bt = 1, M = 1, N = 1, K = 1;
    int p = 0;
    for (int i = 0; i < g.n_batch; ++i) Bt *= g.dim[p++];
    for (int i = 0; i < g.n_m; ++i) M *= g.dim[p++];
    for (int i = 0; i < g.n_n; ++i) N *= g.dim[p++];
    for (int i = 0; i < g.n_k; ++i) K *= g.dim[p++];
    if (Bt == 0 || M == 0 || N == 0) return 0;
    auto min_stride = [&](const int64_t* st, int first, int count) {
        long long best = (1ll << 62);
        for (int i = 0; i < count; ++i)
            if (g.dim[first + i] > 1 && st[first + i] > 0 && st[first + i] < best) best = st[first + i];
        return best;
    };
    const int f_m = g.n_batch, f_n = f_m + g.n_m, f_k = f_n + g.n_n;
    const int a_kfast = min_stride(g.sa, f_k, g.n_k) < min_stride(g.sa, f_m, g.n_m);
    const int b_kfast = min_stride(g.sb, f_k, g.n_k) < min_stride(g.sb, f_n, g.n_n);
    const float2* A = (const float2*)a_dev; const float2* B = (const float2*)b_dev;
    float2* C = (float2*)c_dev;
    set_last_kernel("gemm_c64_tcgen05");
    if (N > 32)
        return launch_c64_nt<64>(g, A, B, C, Bt, M, N, K, a_kfast, b_kfast, ar, ai, accumulate, s);
    return launch_c64_nt<32>(g, A, B, C, Bt, M, N, K, a_kfast, b_kfast, ar, ai, accumulate, s);
}

}  // namespace b2
