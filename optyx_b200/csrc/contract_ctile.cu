// Piece (2b), primary kernel: fused index-permute + complex GEMM, tiled FROM THE
// OUTPUT SIDE.
//
//   C[b, m, n] (+)= alpha * sum_k A[b, m, k] * B[b, k, n]
//
// The CTA tile is built from the FASTEST C modes (up to 4096 elements): each of
// them is a free mode of A (an "M" mode) or of B (an "N" mode), so the tile is a
// TM x TN block of the GEMM whatever the output permutation is (64 x 64 for
// interleaved ket/bra axes of a density tensor -- one contiguous 64 KB region;
// 64 contiguous elements x 64 segments for an outer product; 256 x 16 when a
// small gate block is applied to a big state, ...).  Consequences:
//   * stores are perfectly coalesced: accumulators are staged in shared memory
//     at their C position and streamed out with 16 B/lane consecutive stores;
//   * the A rows / B columns a tile needs are gathered ONCE into shared memory
//     (cp.async 16 B per complex128 element, zero-filled K tail), whatever the
//     operand layouts are -- no transposes in HBM, no reliance on L1 hit rates;
//   * per-tile offsets are tile-invariant tables decoded once per persistent
//     CTA; moving to the next tile is an odometer increment over the outer
//     modes (no divisions in the loop).
// Math: FP64 DMMA mma.sync.m8n8k4 (no tcgen05 kind exists for f64), a complex
// product = 4 real DMMAs on interleaved (re, im) fragments.  The tile is a grid
// of up to 64 m8n8 blocks; each of the 8 warps owns 8 of them.
#include "b2_common.cuh"

namespace b2 {

constexpr int CT_THREADS = 256, CT_BK = 8, CT_TMAX = 4096, CT_MAXDIM = 256;
constexpr int CT_NB = 8;      // m8n8 blocks per warp (8 warps x 8 = 64 blocks per tile)
constexpr size_t CT_SMEM_MAX = 110 * 1024;

struct ct_desc {
    int n_outer, n_im, n_in, n_k;       // mode counts, stored in this order
    int TM, TN, TMp, TNp;               // tile extents and their padding to 8
    int stages, pitch_a, pitch_b;
    long long K;
    int dim[B2_MAX_MODES];
    int rk[B2_MAX_MODES];               // tile modes: rank multiplier (C-address order)
    long long sa[B2_MAX_MODES], sb[B2_MAX_MODES], sc[B2_MAX_MODES];
};

__device__ __forceinline__ void ct_cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem),
                 "r"(src_bytes));
}
__device__ __forceinline__ void ct_commit() { asm volatile("cp.async.commit_group;\n"); }
__device__ __forceinline__ void ct_wait_all() { asm volatile("cp.async.wait_group 0;\n"); }
__device__ __forceinline__ void ct_wait_1() { asm volatile("cp.async.wait_group 1;\n"); }
__device__ __forceinline__ void ct_wait_2() { asm volatile("cp.async.wait_group 2;\n"); }

__device__ __forceinline__ void ct_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(CT_THREADS, 2)
contract_ctile_c128_kernel(const __grid_constant__ ct_desc g, long long n_tiles,
                           const double2* __restrict__ A, const double2* __restrict__ B,
                           double2* __restrict__ Cout, double alpha_re, double alpha_im,
                           int accumulate) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // ---- carve shared memory
    const int T = g.TM * g.TN;
    const int stage_elems = CT_BK * (g.pitch_a + g.pitch_b);
    double2* ring = reinterpret_cast<double2*>(smem_raw);              // stages x stage_elems
    size_t ring_bytes = (size_t)g.stages * stage_elems * sizeof(double2);
    size_t cs_bytes = (size_t)T * sizeof(double2);
    size_t big = ring_bytes > cs_bytes ? ring_bytes : cs_bytes;        // Cs aliases the ring
    double2* Cs = ring;
    long long* row_a = reinterpret_cast<long long*>(smem_raw + big);   // [TMp]
    long long* col_b = row_a + g.TMp;                                  // [TNp]
    long long* k_a = col_b + g.TNp;                                    // [stages][BK]
    long long* k_b = k_a + g.stages * CT_BK;
    long long* caddr = k_b + g.stages * CT_BK;                         // [T] C offset by rank
    int* pos_m = reinterpret_cast<int*>(caddr + T);                    // [TMp] rank part of row
    int* pos_n = pos_m + g.TMp;                                        // [TNp] rank part of col

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f_im = g.n_outer, f_in = f_im + g.n_im, f_k = f_in + g.n_in;

    // ---- tile-invariant tables (mode 0 of each group is the fastest)
    // rank = position of a tile element in C-address order (mixed radix over
    // the tile's modes sorted by C stride; g.rk[] holds each mode's multiplier)
    for (int i = tid; i < g.TMp; i += CT_THREADS) {
        long long oa = 0; int pc = 0; int rem = i;
        if (i < g.TM) {
            for (int m = 0; m < g.n_im; ++m) {
                const int d = g.dim[f_im + m];
                const int q = rem / d, x = rem - q * d;
                rem = q;
                oa += (long long)x * g.sa[f_im + m];
                pc += x * g.rk[f_im + m];
            }
        }
        row_a[i] = oa; pos_m[i] = pc;
    }
    for (int j = tid; j < g.TNp; j += CT_THREADS) {
        long long ob = 0; int pc = 0; int rem = j;
        if (j < g.TN) {
            for (int m = 0; m < g.n_in; ++m) {
                const int d = g.dim[f_in + m];
                const int q = rem / d, x = rem - q * d;
                rem = q;
                ob += (long long)x * g.sb[f_in + m];
                pc += x * g.rk[f_in + m];
            }
        }
        col_b[j] = ob; pos_n[j] = pc;
    }
    for (int e = tid; e < T; e += CT_THREADS) {
        // e = i + TM * j  ->  rank and C offset
        const int i = e % g.TM, j = e / g.TM;
        long long off = 0; int r = 0; int rem = i;
        for (int m = 0; m < g.n_im; ++m) {
            const int d = g.dim[f_im + m];
            const int q = rem / d, x = rem - q * d;
            rem = q;
            off += (long long)x * g.sc[f_im + m];
            r += x * g.rk[f_im + m];
        }
        rem = j;
        for (int m = 0; m < g.n_in; ++m) {
            const int d = g.dim[f_in + m];
            const int q = rem / d, x = rem - q * d;
            rem = q;
            off += (long long)x * g.sc[f_in + m];
            r += x * g.rk[f_in + m];
        }
        caddr[r] = off;
    }

    // ---- contiguous range of outer tiles for this CTA + odometer
    const long long per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per;
    const long long t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
    int cnt[B2_MAX_MODES];
    long long oa = 0, ob = 0, oc = 0;
    {
        long long rem = t0 < n_tiles ? t0 : 0;
        for (int m = 0; m < g.n_outer; ++m) {          // outer mode 0 fastest
            const long long d = g.dim[m];
            const long long q = rem / d;
            const int x = (int)(rem - q * d);
            rem = q;
            cnt[m] = x;
            oa += x * g.sa[m]; ob += x * g.sb[m]; oc += x * g.sc[m];
        }
    }
    const int nchunks = (int)((g.K + CT_BK - 1) / CT_BK);
    const int n_blocks_n = g.TNp >> 3;
    const int n_blocks = (g.TMp >> 3) * n_blocks_n;
    const int frow = lane >> 2, fk = lane & 3;
    const int S = g.stages;

    auto decode_k = [&](int chunk) {
        if (tid < CT_BK) {
            long long xa = 0, xb = 0;
            unsigned k = (unsigned)chunk * CT_BK + (unsigned)tid;     // K < 2^31: 32-bit divisions
            if ((long long)k >= g.K) k = 0;
            for (int m = 0; m < g.n_k; ++m) {
                const unsigned d = (unsigned)g.dim[f_k + m];
                const unsigned q = k / d, x = k - q * d;
                k = q;
                xa += (long long)x * g.sa[f_k + m];
                xb += (long long)x * g.sb[f_k + m];
            }
            k_a[(chunk % S) * CT_BK + tid] = xa;
            k_b[(chunk % S) * CT_BK + tid] = xb;
        }
    };
    __syncthreads();

    // ---- per-thread constants of the MMA blocks this warp owns
    // (fragment offsets and staging positions: no index math inside the loops)
    int a_off[CT_NB], b_off[CT_NB], spos[CT_NB][2];
    int my_blocks = n_blocks - warp * CT_NB;
    my_blocks = my_blocks < 0 ? 0 : (my_blocks > CT_NB ? CT_NB : my_blocks);
    #pragma unroll
    for (int q = 0; q < CT_NB; ++q) {
        const int blk = warp * CT_NB + q;
        const int bi = blk / n_blocks_n, bj = blk - bi * n_blocks_n;
        a_off[q] = bi * 8 + frow;
        b_off[q] = bj * 8 + frow;
        const int i = bi * 8 + frow;
        #pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = bj * 8 + fk * 2 + h;
            spos[q][h] = (blk < n_blocks && i < g.TM && j < g.TN) ? pos_m[i] + pos_n[j] : -1;
        }
    }
    const bool unit_alpha = (alpha_re == 1.0 && alpha_im == 0.0);

    for (long long t = t0; t < t1; ++t) {
        if (t != t0) {
            for (int m = 0; m < g.n_outer; ++m) {
                oa += g.sa[m]; ob += g.sb[m]; oc += g.sc[m];
                if (++cnt[m] < g.dim[m]) break;
                oa -= (long long)g.dim[m] * g.sa[m];
                ob -= (long long)g.dim[m] * g.sb[m];
                oc -= (long long)g.dim[m] * g.sc[m];
                cnt[m] = 0;
            }
        }
        const double2* At = A + oa;
        const double2* Bt = B + ob;
        auto issue = [&](int chunk) {
            double2* as = ring + (size_t)(chunk % S) * stage_elems;
            double2* bs = as + CT_BK * g.pitch_a;
            const long long kbase = (long long)chunk * CT_BK;
            const long long* ka = k_a + (chunk % S) * CT_BK;
            const long long* kb = k_b + (chunk % S) * CT_BK;
            for (int e = tid; e < g.TMp * CT_BK; e += CT_THREADS) {
                const int j = e & (CT_BK - 1), i = e >> 3;       // k fastest
                const bool ok = (i < g.TM) && (kbase + j < g.K);
                ct_cp_async16(&as[j * g.pitch_a + i], ok ? At + row_a[i] + ka[j] : At, ok ? 16 : 0);
            }
            for (int e = tid; e < g.TNp * CT_BK; e += CT_THREADS) {
                const int j = e & (CT_BK - 1), i = e >> 3;
                const bool ok = (i < g.TN) && (kbase + j < g.K);
                ct_cp_async16(&bs[j * g.pitch_b + i], ok ? Bt + col_b[i] + kb[j] : Bt, ok ? 16 : 0);
            }
        };

        for (int c = 0; c < S - 1; ++c) if (c < nchunks) decode_k(c);
        __syncthreads();          // k tables visible; previous tile's Cs reads done
        for (int c = 0; c < S - 1; ++c) {
            if (c < nchunks) issue(c);
            ct_commit();
        }
        double cre[CT_NB][2], cim[CT_NB][2];
        #pragma unroll
        for (int q = 0; q < CT_NB; ++q) { cre[q][0] = cre[q][1] = 0; cim[q][0] = cim[q][1] = 0; }

        for (int c = 0; c < nchunks; ++c) {
            const int nxt = c + S - 1;
            if (nxt < nchunks) decode_k(nxt);
            if (S == 2) ct_wait_all(); else if (S == 3) ct_wait_1(); else ct_wait_2();
            __syncthreads();
            if (nxt < nchunks) issue(nxt);
            ct_commit();
            const double2* as = ring + (size_t)(c % S) * stage_elems;
            const double2* bs = as + CT_BK * g.pitch_a;
            const int klim = (g.K - (long long)c * CT_BK) > 4 ? CT_BK : 4;   // skip an all-zero k4
            for (int k4 = 0; k4 < klim; k4 += 4) {
                const double2* ak = as + (k4 + fk) * g.pitch_a;
                const double2* bk = bs + (k4 + fk) * g.pitch_b;
                #pragma unroll
                for (int q = 0; q < CT_NB; ++q) {
                    if (q < my_blocks) {                 // warp-uniform
                        const double2 af = ak[a_off[q]];
                        const double2 bf = bk[b_off[q]];
                        ct_dmma(cre[q][0], cre[q][1], af.x, bf.x);
                        ct_dmma(cre[q][0], cre[q][1], -af.y, bf.y);
                        ct_dmma(cim[q][0], cim[q][1], af.x, bf.y);
                        ct_dmma(cim[q][0], cim[q][1], af.y, bf.x);
                    }
                }
            }
        }
        ct_wait_all();
        __syncthreads();          // every warp done reading the ring -> reuse as Cs
        #pragma unroll
        for (int q = 0; q < CT_NB; ++q) {
            if (q < my_blocks) {
                #pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int sp = spos[q][h];
                    if (sp >= 0) {
                        double2 v;
                        if (unit_alpha) { v.x = cre[q][h]; v.y = cim[q][h]; }
                        else {
                            v.x = alpha_re * cre[q][h] - alpha_im * cim[q][h];
                            v.y = alpha_re * cim[q][h] + alpha_im * cre[q][h];
                        }
                        Cs[sp] = v;
                    }
                }
            }
        }
        __syncthreads();
        double2* Ct = Cout + oc;
        if (accumulate) {
            for (int e = tid; e < T; e += CT_THREADS) {
                double2 v = Cs[e];
                double2* dst = Ct + caddr[e];
                const double2 o = *dst;
                v.x += o.x; v.y += o.y;
                *dst = v;
            }
        } else {
            #pragma unroll 4
            for (int e = tid; e < T; e += CT_THREADS) Ct[caddr[e]] = Cs[e];
        }
        // the __syncthreads at the top of the next tile protects Cs
    }
}

// Build the C-tiled descriptor from the grouped modes.  The tile takes the
// fastest C modes: first whatever is fastest until the contiguous run is >= 16
// elements (256 B), then M and N modes alternately so the tile stays as square
// as the shapes allow (operand reuse), up to 4096 elements and 256 per side.
// The planner merges stride-compatible modes; a tile needs finer grain, so the
// free modes are split back into factors of at most 16 first.
static void refine_gemm_modes(const b2_gemm_modes& in, b2_gemm_modes& out) {
    const int counts[4] = {in.n_batch, in.n_m, in.n_n, in.n_k};
    int newc[4] = {0, 0, 0, 0};
    int p = 0, src = 0;
    const int total = in.n_batch + in.n_m + in.n_n + in.n_k;
    for (int grp = 0; grp < 4; ++grp) {
        for (int i = 0; i < counts[grp]; ++i, ++src) {
            long long dd = in.dim[src], sa = in.sa[src], sb = in.sb[src], sc = in.sc[src];
            while (grp < 3 && dd > 16 && p + (total - src) < B2_MAX_MODES - 1) {
                int f = 1;
                for (int c = 16; c >= 2; --c) if (dd % c == 0) { f = c; break; }
                if (f == 1) break;
                out.dim[p] = f; out.sa[p] = sa; out.sb[p] = sb; out.sc[p] = sc; ++p; ++newc[grp];
                sa *= f; sb *= f; sc *= f; dd /= f;
            }
            out.dim[p] = (int)dd; out.sa[p] = sa; out.sb[p] = sb; out.sc[p] = sc; ++p; ++newc[grp];
        }
    }
    out.n_batch = newc[0]; out.n_m = newc[1]; out.n_n = newc[2]; out.n_k = newc[3];
}

static bool make_ctile(const b2_gemm_modes& g_in, ct_desc& d, long long& n_tiles, size_t& smem) {
    b2_gemm_modes g;
    refine_gemm_modes(g_in, g);
    const int f_m = g.n_batch, f_n = f_m + g.n_m, f_k = f_n + g.n_n;
    const int n_free = g.n_batch + g.n_m + g.n_n;
    int order[B2_MAX_MODES];
    int cnt = 0;
    for (int i = 0; i < n_free; ++i) if (g.dim[i] > 1) order[cnt++] = i;
    for (int i = 1; i < cnt; ++i) {                     // sort by C stride ascending
        int x = order[i], j = i - 1;
        while (j >= 0 && g.sc[order[j]] > g.sc[x]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = x;
    }
    if (cnt == 0 || g.sc[order[0]] != 1) return false;  // output not unit-stride
    bool inner[B2_MAX_MODES] = {false};
    long long T = 1;
    int TM = 1, TN = 1;
    auto is_m = [&](int i) { return i >= f_m && i < f_n; };
    auto is_n = [&](int i) { return i >= f_n && i < f_k; };
    auto fits = [&](int i) {
        const long long dd = g.dim[i];
        if (T * dd > CT_TMAX) return false;
        if (is_m(i)) return (long long)TM * dd <= CT_MAXDIM;
        if (is_n(i)) return (long long)TN * dd <= CT_MAXDIM;
        return false;
    };
    auto take = [&](int i) {
        if (is_m(i)) TM *= g.dim[i]; else TN *= g.dim[i];
        T *= g.dim[i];
        inner[i] = true;
    };
    int o = 0;
    for (; o < cnt && T < 16; ++o) {                    // phase 1: the fastest run
        if (!fits(order[o])) break;
        take(order[o]);
    }
    bool m_open = true, n_open = true;                  // phase 2: balance
    while (m_open || n_open) {
        const bool want_m = m_open && (!n_open || TM <= TN);
        int pick = -1;
        for (int q = o; q < cnt; ++q) {
            const int i = order[q];
            if (inner[i]) continue;
            if (want_m ? is_m(i) : is_n(i)) { pick = i; break; }
        }
        if (pick < 0 || !fits(pick)) { if (want_m) m_open = false; else n_open = false; continue; }
        take(pick);
    }
    const int TMp = (TM + 7) & ~7, TNp = (TN + 7) & ~7;
    if (TM < 4 || TN < 4 || (TMp >> 3) * (TNp >> 3) > 64) return false;
    int p = 0;
    auto put = [&](int i) {
        d.dim[p] = g.dim[i]; d.sa[p] = g.sa[i]; d.sb[p] = g.sb[i]; d.sc[p] = g.sc[i];
        d.rk[p] = 0; ++p;
    };
    n_tiles = 1;
    d.n_outer = 0;
    for (int q = 0; q < cnt; ++q) {
        const int i = order[q];
        if (!inner[i]) { put(i); ++d.n_outer; n_tiles *= g.dim[i]; }
    }
    // rank multipliers: mixed radix over ALL tile modes in C-stride order
    int rank_mult[B2_MAX_MODES];
    {
        int mult = 1;
        for (int q = 0; q < cnt; ++q) {
            const int i = order[q];
            if (inner[i]) { rank_mult[i] = mult; mult *= g.dim[i]; }
        }
    }
    d.n_im = 0;
    for (int q = 0; q < cnt; ++q) {
        const int i = order[q];
        if (inner[i] && is_m(i)) { put(i); d.rk[p - 1] = rank_mult[i]; ++d.n_im; }
    }
    d.n_in = 0;
    for (int q = 0; q < cnt; ++q) {
        const int i = order[q];
        if (inner[i] && is_n(i)) { put(i); d.rk[p - 1] = rank_mult[i]; ++d.n_in; }
    }
    d.n_k = 0;
    d.K = 1;
    for (int i = f_k; i < f_k + g.n_k; ++i) if (g.dim[i] > 1) { put(i); ++d.n_k; d.K *= g.dim[i]; }
    d.TM = TM; d.TN = TN; d.TMp = TMp; d.TNp = TNp;
    d.pitch_a = TMp + 2; d.pitch_b = TNp + 2;             // pitch % 8 == 2: conflict-free frags
    const size_t stage_bytes = (size_t)CT_BK * (d.pitch_a + d.pitch_b) * 16;
    const int nchunks = (int)((d.K + CT_BK - 1) / CT_BK);
    int stages = (int)((72 * 1024) / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages > nchunks + 1) stages = nchunks + 1;
    if (stages < 2) stages = 2;
    d.stages = stages;
    const size_t ring = stage_bytes * stages, cs = (size_t)TM * TN * 16;
    smem = (ring > cs ? ring : cs) + (size_t)(TMp + TNp) * 8 + (size_t)stages * CT_BK * 16 +
           (size_t)TM * TN * 8 + (size_t)(TMp + TNp) * 4 + 64;
    return smem <= CT_SMEM_MAX;
}

int launch_ctile(const b2_gemm_modes& g, const void* a, const void* b, void* c, double ar,
                 double ai, int accumulate, cudaStream_t s, bool& handled) {
    handled = false;
    ct_desc d;
    long long n_tiles = 0;
    size_t smem = 0;
    if (!make_ctile(g, d, n_tiles, smem)) return 0;
    int sms = num_sms();
    if (sms <= 0) return 1;
    static b2::once_flags attr_set;                    // kernel attributes are per device
    if (!b2::is_done(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(contract_ctile_c128_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)CT_SMEM_MAX);
        if (e != cudaSuccess) { set_error("ctile smem attr: %s", cudaGetErrorString(e)); return 1; }
        b2::mark_done(attr_set);
    }
    long long blocks = n_tiles;
    const long long cap = (long long)sms * 2;
    if (blocks > cap) blocks = cap;
    contract_ctile_c128_kernel<<<(unsigned)blocks, CT_THREADS, smem, s>>>(
        d, n_tiles, (const double2*)a, (const double2*)b, (double2*)c, ar, ai, accumulate);
    handled = true;
    set_last_kernel("ctile");
    return check_launch("b2_contract_gemm(ctile)");
}

}  // namespace b2
