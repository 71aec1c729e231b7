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
// Data flow per CTA (tile 128 x NT complex, K chunk 16 complex = 32 real):
//   all 8 warps : gather the chunk's A rows / B columns through the per-mode
//                 strides (the index permutation; no transposes in HBM), split
//                 hi/lo in registers, st.shared into the canonical K-major
//                 no-swizzle UMMA layout [k-group][row][16 B], fence.proxy.async
//   one thread  : 12 tcgen05.mma (4 k-steps x 3 split products) per chunk, D in
//                 TMEM (128 lanes x 2NT fp32 columns, two sets), tcgen05.commit
//                 -> mbarrier frees the stage; the next gather overlaps the MMAs
//   all 8 warps : every 4 chunks the finished TMEM set is drained with
//                 tcgen05.ld 32x32b (lane = row, columns = (re, im) pairs) into
//                 FP32 register accumulators (two-level accumulation, see below)
//   epilogue    : alpha scaling, scatter through the C permutation.
#include "b2_common.cuh"

namespace b2 {

constexpr int G4_THREADS = 256, G4_BM = 128, G4_KC = 16, G4_STAGES = 2;
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

template <int NT>
struct g4_smem {
    // [stage][hi/lo][k-group][row][4 floats]
    float4 a[G4_STAGES][2][G4_KG][G4_BM];
    float4 b[G4_STAGES][2][G4_KG][2 * NT];
    long long row_a[G4_BM], row_c[G4_BM], col_b[NT], col_c[NT];
    long long k_a[2][G4_KC], k_b[2][G4_KC];
    uint64_t bar[G4_STAGES];
    uint32_t tmem_base;
};

template <int NT>
__global__ void __launch_bounds__(G4_THREADS, 1)
contract_gemm_c64_kernel(const __grid_constant__ b2_gemm_modes g, const float2* __restrict__ A,
                         const float2* __restrict__ B, float2* __restrict__ Cout, long long M,
                         long long N, long long K, int tiles_m, int a_kfast, int b_kfast,
                         float alpha_re, float alpha_im, int accumulate) {
    extern __shared__ __align__(1024) unsigned char g4_raw[];
    g4_smem<NT>& sm = *reinterpret_cast<g4_smem<NT>*>(g4_raw);
    constexpr int NP = 2 * NT;                       // real columns of D
    constexpr int A_ITEMS = G4_BM * G4_KG / G4_THREADS;      // (row, k-pair) items per thread
    constexpr int B_ITEMS = (NT * G4_KG + G4_THREADS - 1) / G4_THREADS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tile = blockIdx.x;
    const long long m0 = (tile % tiles_m) * G4_BM, n0 = (tile / tiles_m) * NT;
    const int f_m = g.n_batch, f_n = f_m + g.n_m, f_k = f_n + g.n_n;

    {
        long long ba, bb, bc;
        g4_decode3(blockIdx.y, g, 0, g.n_batch, ba, bb, bc);
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
    auto decode_k = [&](int chunk) {
        if (tid < G4_KC) {
            long long oa, ob, oc;
            const long long k = (long long)chunk * G4_KC + tid;
            g4_decode3(k < K ? k : 0, g, f_k, g.n_k, oa, ob, oc);
            sm.k_a[chunk & 1][tid] = oa;
            sm.k_b[chunk & 1][tid] = ob;
        }
    };
    decode_k(0);
    __syncwarp();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     ::"r"(g4_smem_u32(&sm.tmem_base)), "r"(2 * NP));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 32) {
        for (int s = 0; s < G4_STAGES; ++s) g4_mbar_init(&sm.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = sm.tmem_base;

    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 2NT, M = 128
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) |
                               ((uint32_t)(G4_BM >> 4) << 24);
    const int nchunks = (int)((K + G4_KC - 1) / G4_KC);
    const float2 zero2 = make_float2(0.f, 0.f);

    // Two-level accumulation.  The tensor core truncates its FP32 accumulator on
    // every add (measured: relative error grows ~4e-8 per accumulation), so the
    // TMEM accumulator only ever sums one SEGMENT of G4_SEG chunks (48 MMAs); the
    // segments are added in FP32 registers (round-to-nearest).  Two TMEM sets
    // alternate: segment s+1 accumulates while segment s is drained.
    const int q = warp & 3, half = warp >> 2;               // TMEM lane quarter, column half
    constexpr int COLS_PER_WARP = NP / 2;
    float acc[COLS_PER_WARP];
    #pragma unroll
    for (int j = 0; j < COLS_PER_WARP; ++j) acc[j] = 0.f;
    auto drain = [&](int seg) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t base = tmem_d + (uint32_t)((seg & 1) * NP) + ((uint32_t)(q * 32) << 16) +
                              (uint32_t)(half * COLS_PER_WARP);
        #pragma unroll
        for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) {
            uint32_t v[16];
            g4_ld16(base + (uint32_t)c0, v);
            #pragma unroll
            for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(v[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    };
    int next_drain = 0;                                     // first segment not yet drained

    for (int c = 0; c < nchunks; ++c) {
        const int st = c & 1;
        const long long kbase = (long long)c * G4_KC;
        // ---- gather chunk c: global -> registers (overlaps the MMAs in flight)
        float2 av[A_ITEMS][2], bv[B_ITEMS][2];
        int a_r[A_ITEMS], a_kp[A_ITEMS], b_n[B_ITEMS], b_kp[B_ITEMS];
        #pragma unroll
        for (int p = 0; p < A_ITEMS; ++p) {
            const int e = tid + p * G4_THREADS;
            int r, kp;
            if (a_kfast) { kp = e % G4_KG; r = e / G4_KG; } else { r = e % G4_BM; kp = e / G4_BM; }
            a_r[p] = r; a_kp[p] = kp;
            const bool okr = m0 + r < M;
            const float2* base = A + sm.row_a[r];
            av[p][0] = (okr && kbase + 2 * kp < K) ? base[sm.k_a[st][2 * kp]] : zero2;
            av[p][1] = (okr && kbase + 2 * kp + 1 < K) ? base[sm.k_a[st][2 * kp + 1]] : zero2;
        }
        #pragma unroll
        for (int p = 0; p < B_ITEMS; ++p) {
            const int e = tid + p * G4_THREADS;
            int n, kp;
            if (b_kfast) { kp = e % G4_KG; n = e / G4_KG; } else { n = e % NT; kp = e / NT; }
            b_n[p] = n; b_kp[p] = kp;
            const bool okn = n0 + n < N;
            const float2* base = B + sm.col_b[n];
            bv[p][0] = (okn && kbase + 2 * kp < K) ? base[sm.k_b[st][2 * kp]] : zero2;
            bv[p][1] = (okn && kbase + 2 * kp + 1 < K) ? base[sm.k_b[st][2 * kp + 1]] : zero2;
        }
        if (c + 1 < nchunks) decode_k(c + 1);
        // ---- the stage is free once the MMAs of chunk c - 2 have completed; if
        //      that chunk closed a segment, its TMEM set is complete: drain it
        if (c >= G4_STAGES) {
            g4_mbar_wait(&sm.bar[st], (uint32_t)(((c >> 1) - 1) & 1));
            if ((c - 2) % G4_SEG == G4_SEG - 1) { drain(next_drain); ++next_drain; }
        }
        // ---- split and store in the UMMA K-major layout
        #pragma unroll
        for (int p = 0; p < A_ITEMS; ++p) {
            const float4 x = make_float4(av[p][0].x, av[p][0].y, av[p][1].x, av[p][1].y);
            const float4 h = make_float4(g4_hi(x.x), g4_hi(x.y), g4_hi(x.z), g4_hi(x.w));
            const float4 l = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
            sm.a[st][0][a_kp[p]][a_r[p]] = h;
            sm.a[st][1][a_kp[p]][a_r[p]] = l;
        }
        #pragma unroll
        for (int p = 0; p < B_ITEMS; ++p) {
            const float br0 = bv[p][0].x, bi0 = bv[p][0].y, br1 = bv[p][1].x, bi1 = bv[p][1].y;
            const float4 xr = make_float4(br0, -bi0, br1, -bi1);      // row of C_re
            const float4 xi = make_float4(bi0, br0, bi1, br1);        // row of C_im
            const float4 hr = make_float4(g4_hi(xr.x), g4_hi(xr.y), g4_hi(xr.z), g4_hi(xr.w));
            const float4 hi = make_float4(g4_hi(xi.x), g4_hi(xi.y), g4_hi(xi.z), g4_hi(xi.w));
            const int n = b_n[p], kp = b_kp[p];
            sm.b[st][0][kp][2 * n] = hr;
            sm.b[st][0][kp][2 * n + 1] = hi;
            sm.b[st][1][kp][2 * n] = make_float4(xr.x - hr.x, xr.y - hr.y, xr.z - hr.z, xr.w - hr.w);
            sm.b[st][1][kp][2 * n + 1] = make_float4(xi.x - hi.x, xi.y - hi.y, xi.z - hi.z, xi.w - hi.w);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            constexpr uint32_t a_lbo = G4_BM * 16, b_lbo = NP * 16, sbo = 128;
            const uint32_t a_hi = g4_smem_u32(&sm.a[st][0][0][0]), a_lo = g4_smem_u32(&sm.a[st][1][0][0]);
            const uint32_t b_hi = g4_smem_u32(&sm.b[st][0][0][0]), b_lo = g4_smem_u32(&sm.b[st][1][0][0]);
            const int seg = c / G4_SEG;
            const uint32_t d = tmem_d + (uint32_t)((seg & 1) * NP);
            const bool first = (c % G4_SEG) == 0;
            #pragma unroll
            for (int ks = 0; ks < G4_KG / 2; ++ks) {          // UMMA K = 8 tf32 = 2 k-groups
                const uint64_t dah = g4_desc(a_hi + ks * 2 * a_lbo, a_lbo, sbo);
                const uint64_t dal = g4_desc(a_lo + ks * 2 * a_lbo, a_lbo, sbo);
                const uint64_t dbh = g4_desc(b_hi + ks * 2 * b_lbo, b_lbo, sbo);
                const uint64_t dbl = g4_desc(b_lo + ks * 2 * b_lbo, b_lbo, sbo);
                g4_mma_tf32(d, dah, dbh, idesc, (first && ks == 0) ? 0u : 1u);
                g4_mma_tf32(d, dah, dbl, idesc, 1u);
                g4_mma_tf32(d, dal, dbh, idesc, 1u);
            }
            g4_commit(&sm.bar[st]);
        }
    }
    // ---- all MMAs done (a commit tracks every earlier MMA of the issuing thread):
    //      drain the segments still in TMEM (at most two)
    {
        const int last = nchunks - 1;
        g4_mbar_wait(&sm.bar[last & 1], (uint32_t)((last >> 1) & 1));
        const int nseg = (nchunks + G4_SEG - 1) / G4_SEG;
        for (; next_drain < nseg; ++next_drain) drain(next_drain);
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
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(2 * NP));
    }
}

template <int NT>
static int launch_c64_nt(const b2_gemm_modes& g, const float2* a, const float2* b, float2* c,
                         long long Bt, long long M, long long N, long long K, int a_kfast,
                         int b_kfast, double ar, double ai, int accumulate, cudaStream_t s) {
    const long long tiles_m = (M + G4_BM - 1) / G4_BM, tiles_n = (N + NT - 1) / NT;
    if (tiles_m * tiles_n > 2147483647ll || Bt > 65535) {
        set_error("b2_contract_gemm(c64): grid too large (tiles=%lld batch=%lld)", tiles_m * tiles_n, Bt);
        return 1;
    }
    const int smem = (int)sizeof(g4_smem<NT>) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(contract_gemm_c64_kernel<NT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("b2_contract_gemm(c64): smem attr: %s", cudaGetErrorString(e)); return 1; }
        attr_set = true;
    }
    dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)Bt);
    contract_gemm_c64_kernel<NT><<<grid, G4_THREADS, smem, s>>>(
        g, a, b, c, M, N, K, (int)tiles_m, a_kfast, b_kfast, (float)ar, (float)ai, accumulate);
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
