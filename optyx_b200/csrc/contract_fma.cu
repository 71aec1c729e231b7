// Piece (2a), primary kernel for bandwidth-bound pairs with a short reduction
// (K <= 16): outer products, "apply a small tensor to a huge one", thin pairs,
// element-wise products over shared (hyper) indices.
//
//   C[b, m, n] (+)= alpha * sum_k A[b, m, k] * B[b, k, n]
//
// Tiled from the output side like contract_ctile.cu: the CTA tile is TM x TN
// elements made of the FASTEST C modes, oriented so that the fastest C mode is
// a ROW mode (a mode of A: a free M mode or a batch mode shared with B).  A work
// item is (row i, group of U columns): the thread pulls the K values A[i, :]
// into registers once (16 B loads, lanes = consecutive rows -> consecutive C
// addresses), reads B[:, j] from shared memory (the K x TB x TN block of B is
// staged once per tile, or once per CTA when it does not change; without batch
// modes in the tile the reads are pure broadcasts), does 4*K FMAs per output
// and stores each output straight from registers: lanes of a warp write
// consecutive rows = consecutive addresses.  No accumulator staging, at most
// one barrier pair per tile, every load independent of every other.  alpha is
// folded into the staged B block.  Wide tiles without batch modes inside run RM
// rows per thread (B reuse, see the RM branch); stores are streaming
// (st.global.cs) and CTAs take tiles round-robin (FM_DEFAULT_FLAGS).
#include "b2_common.cuh"
#include <stdlib.h>

namespace b2 {

constexpr int FM_THREADS = 256, FM_TMAX = 65536, FM_MAXM = 1024, FM_MAXN = 64, FM_U = 16;
constexpr int FM_STREAM_STORES = 1, FM_ROUND_ROBIN = 2, FM_DEFAULT_FLAGS = FM_STREAM_STORES | FM_ROUND_ROBIN,
              FM_DEFAULT_RM = 4;
constexpr int FM_KMAX = 16, FM_MAXTB = 64, FM_BTILE_BYTES = 32 * 1024, FM_SMEM_MAX = 64 * 1024;

struct fm_desc {
    int n_outer, n_ir, n_in, n_k;       // mode counts, stored in this order
    int TM, TN, TB, K;                  // rows, columns, batch combos inside the rows, reduction
    int b_static;                       // B's tile does not depend on the outer modes
    int dim[B2_MAX_MODES];
    long long sa[B2_MAX_MODES], sb[B2_MAX_MODES], sc[B2_MAX_MODES];
};

template <typename T, int KP, int RMAX, int RM = 1>
__global__ void __launch_bounds__(FM_THREADS)
contract_fma_kernel(const __grid_constant__ fm_desc g, long long n_tiles,
                    const typename cplx<T>::type* __restrict__ A,
                    const typename cplx<T>::type* __restrict__ B,
                    typename cplx<T>::type* __restrict__ Cout, T alpha_re, T alpha_im,
                    int accumulate, int flags) {
    using C = typename cplx<T>::type;
    extern __shared__ __align__(16) unsigned char fm_smem[];
    long long* row_a = reinterpret_cast<long long*>(fm_smem);      // [TM] A offset of row i
    long long* pos_m = row_a + g.TM;                                // [TM] C offset of row i
    long long* col_b = pos_m + g.TM;                                // [TN] B offset of col j
    long long* pos_n = col_b + g.TN;                                // [TN] C offset of col j
    long long* bt_off = pos_n + g.TN;                               // [TB] B offset of batch combo
    long long* k_a = bt_off + ((g.TB + 1) & ~1);                    // [KP] (keeps Bs 16 B aligned)
    long long* k_b = k_a + KP;                                      // [KP]
    C* Bs = reinterpret_cast<C*>(k_b + KP);                         // [KP][TB][TN]
    int* row_bs = reinterpret_cast<int*>(Bs + (size_t)KP * g.TB * g.TN);   // [TM] bt(i) * TN

    const int tid = threadIdx.x;
    const int f_ir = g.n_outer, f_in = f_ir + g.n_ir, f_k = f_in + g.n_in;

    for (int i = tid; i < g.TM; i += FM_THREADS) {
        long long oa = 0, oc = 0; int rem = i, bt = 0, btm = 1;
        for (int m = 0; m < g.n_ir; ++m) {
            const int d = g.dim[f_ir + m];
            const int q = rem / d, x = rem - q * d;
            rem = q;
            oa += (long long)x * g.sa[f_ir + m];
            oc += (long long)x * g.sc[f_ir + m];
            if (g.sb[f_ir + m] != 0) { bt += x * btm; btm *= d; }
        }
        row_a[i] = oa; pos_m[i] = oc; row_bs[i] = bt * g.TN;
    }
    for (int b = tid; b < g.TB; b += FM_THREADS) {
        long long ob = 0; int rem = b;
        for (int m = 0; m < g.n_ir; ++m) {
            if (g.sb[f_ir + m] == 0) continue;
            const int d = g.dim[f_ir + m];
            const int q = rem / d, x = rem - q * d;
            rem = q;
            ob += (long long)x * g.sb[f_ir + m];
        }
        bt_off[b] = ob;
    }
    for (int j = tid; j < g.TN; j += FM_THREADS) {
        long long ob = 0, oc = 0; int rem = j;
        for (int m = 0; m < g.n_in; ++m) {
            const int d = g.dim[f_in + m];
            const int q = rem / d, x = rem - q * d;
            rem = q;
            ob += (long long)x * g.sb[f_in + m];
            oc += (long long)x * g.sc[f_in + m];
        }
        col_b[j] = ob; pos_n[j] = oc;
    }
    if (tid < KP) {
        long long xa = 0, xb = 0; int rem = tid < g.K ? tid : 0;
        for (int m = 0; m < g.n_k; ++m) {
            const int d = g.dim[f_k + m];
            const int q = rem / d, x = rem - q * d;
            rem = q;
            xa += (long long)x * g.sa[f_k + m];
            xb += (long long)x * g.sb[f_k + m];
        }
        k_a[tid] = xa; k_b[tid] = xb;
    }

    // contiguous range of outer tiles + odometer (outer mode 0 = fastest in C)
    // (FM_ROUND_ROBIN: tile t = blockIdx + i * grid instead -- the CTAs resident at
    // any moment then write NEIGHBOURING tiles, i.e. a few long contiguous runs of C
    // instead of grid x TN short ones scattered over the whole output)
    const bool rr = RMAX == 1 && (flags & FM_ROUND_ROBIN);
    const long long per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = rr ? (long long)blockIdx.x : (long long)blockIdx.x * per;
    const long long t1 = rr ? n_tiles : (t0 + per < n_tiles ? t0 + per : n_tiles);
    const long long t_step = rr ? (long long)gridDim.x : (long long)RMAX;
    if (t0 >= t1) return;
    int cnt[B2_MAX_MODES];
    long long oa = 0, ob = 0, oc = 0;
    auto decode = [&](long long rem) {
        oa = ob = oc = 0;
        for (int m = 0; m < g.n_outer; ++m) {
            const long long d = g.dim[m];
            const long long q = rem / d;
            const int x = (int)(rem - q * d);
            rem = q;
            cnt[m] = x;
            oa += x * g.sa[m]; ob += x * g.sb[m]; oc += x * g.sc[m];
        }
    };
    decode(t0);
    __syncthreads();

    const int n_jg = (g.TN + FM_U - 1) / FM_U;
    const int n_items = g.TM * n_jg;
    const int bplane = g.TB * g.TN;
    const bool unit_alpha = (alpha_re == (T)1 && alpha_im == (T)0);
    const C alpha = make_c<T>(alpha_re, alpha_im);
    const C zero = make_c<T>((T)0, (T)0);

    // R consecutive tiles are processed together (same work item, R tile offsets):
    // with K = 1 or 2 one item is one or two 16 B loads per thread, far too little
    // in flight to cover the HBM latency; R tiles at once give R times as much and
    // amortise the per-item table look-ups.  Needs a tile-invariant B block.
    // (RMAX > 1 is instantiated for narrow tiles only: TN <= 8 and a static B)
    const int R = RMAX;
    auto advance = [&]() {
        for (int m = 0; m < g.n_outer; ++m) {
            oa += g.sa[m]; ob += g.sb[m]; oc += g.sc[m];
            if (++cnt[m] < g.dim[m]) break;
            oa -= (long long)g.dim[m] * g.sa[m];
            ob -= (long long)g.dim[m] * g.sb[m];
            oc -= (long long)g.dim[m] * g.sc[m];
            cnt[m] = 0;
        }
    };
    for (long long t = t0; t < t1; t += t_step) {
        const int nt = rr ? 1 : (int)(t1 - t < R ? t1 - t : R);
        long long oa_r[RMAX], oc_r[RMAX];
        if (rr && t != t0) decode(t);
        #pragma unroll
        for (int r = 0; r < RMAX; ++r) {
            if (r < nt) {
                if (!rr && t + r != t0) advance();
                oa_r[r] = oa; oc_r[r] = oc;
            } else { oa_r[r] = oa; oc_r[r] = oc; }
        }
        if (t == t0 || !g.b_static) {
            if (t != t0) __syncthreads();              // everyone done with the old Bs
            for (int e = tid; e < KP * bplane; e += FM_THREADS) {
                const int k = e / bplane, r = e - k * bplane;
                const int bt = r / g.TN, j = r - bt * g.TN;
                // alpha is folded into the staged B block (K x TB x TN multiplies per tile
                // instead of one complex multiply per output)
                C bv = zero;
                if (k < g.K) {
                    bv = B[ob + bt_off[bt] + col_b[j] + k_b[k]];
                    if (!unit_alpha) bv = cmul(alpha, bv);
                }
                Bs[e] = bv;
            }
            __syncthreads();
        }
        if constexpr (RM > 1) {
            // RM rows per thread (rows i, i + S, ..., S = TM / RM; lanes = consecutive
            // rows): every B value read from shared memory feeds RM outputs.  The LSU
            // data pipe -- broadcast LDS of B plus the STG of C -- is this kernel's
            // busiest unit (ncu: 67 % with RM = 1, 28 % of it the B reads), not HBM and
            // not the FP64 pipe.  Host guarantees TB == 1 and TM % RM == 0.
            const int S = g.TM / RM;
            const int n_slots = S * n_jg;
            int prev_i = -1;
            C a[RM][KP];
            long long pm[RM];
            C* cp = Cout + oc_r[0];
            for (int w = tid; w < n_slots; w += FM_THREADS) {
                const int jg = w / S, i = w - jg * S;
                if (i != prev_i) {                       // S == FM_THREADS: once per tile
                    #pragma unroll
                    for (int q = 0; q < RM; ++q) {
                        const C* ap = A + oa_r[0] + row_a[i + q * S];
                        pm[q] = pos_m[i + q * S];
                        #pragma unroll
                        for (int k = 0; k < KP; ++k) a[q][k] = (k < g.K) ? ap[k_a[k]] : zero;
                    }
                    prev_i = i;
                }
                const int j0 = jg * FM_U;
                const int j1 = j0 + FM_U < g.TN ? j0 + FM_U : g.TN;
                #pragma unroll 2
                for (int j = j0; j < j1; ++j) {
                    C b[KP];
                    #pragma unroll
                    for (int k = 0; k < KP; ++k) b[k] = Bs[k * bplane + j];
                    const long long pn = pos_n[j];
                    #pragma unroll
                    for (int q = 0; q < RM; ++q) {
                        C acc = zero;
                        if constexpr (KP >= 4) {
                            C acc2 = zero;
                            #pragma unroll
                            for (int k = 0; k < KP; k += 2) {
                                cfma(acc, a[q][k], b[k]);
                                cfma(acc2, a[q][k + 1], b[k + 1]);
                            }
                            acc.x += acc2.x; acc.y += acc2.y;
                        } else {
                            #pragma unroll
                            for (int k = 0; k < KP; ++k) cfma(acc, a[q][k], b[k]);
                        }
                        C res = acc;
                        C* dst = cp + pm[q] + pn;
                        if (accumulate) { const C o = *dst; res.x += o.x; res.y += o.y; }
                        if (flags & FM_STREAM_STORES) __stcs(dst, res);
                        else *dst = res;
                    }
                }
            }
        } else
        for (int w = tid; w < n_items; w += FM_THREADS) {
            const int jg = w / g.TM, i = w - jg * g.TM;       // lanes = consecutive rows
            const long long ra = row_a[i], pm = pos_m[i];
            const C* bp = Bs + row_bs[i];
            C a[RMAX][KP];
            #pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                const C* ap = A + oa_r[r] + ra;
                #pragma unroll
                for (int k = 0; k < KP; ++k) a[r][k] = (r < nt && k < g.K) ? ap[k_a[k]] : zero;
            }
            const int j0 = jg * FM_U;
            const int j1 = j0 + FM_U < g.TN ? j0 + FM_U : g.TN;
            #pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                if (r >= nt) break;
                C* cp = Cout + oc_r[r] + pm;
                #pragma unroll 4
                for (int j = j0; j < j1; ++j) {
                    C acc = zero;
                    if constexpr (KP >= 4) {
                        // two independent accumulation chains halve the dependent
                        // FMA latency per output
                        C acc2 = zero;
                        #pragma unroll
                        for (int k = 0; k < KP; k += 2) {
                            cfma(acc, a[r][k], bp[k * bplane + j]);
                            cfma(acc2, a[r][k + 1], bp[(k + 1) * bplane + j]);
                        }
                        acc.x += acc2.x; acc.y += acc2.y;
                    } else {
                        #pragma unroll
                        for (int k = 0; k < KP; ++k) cfma(acc, a[r][k], bp[k * bplane + j]);
                    }
                    C res = acc;
                    C* dst = cp + pos_n[j];
                    if (accumulate) { const C o = *dst; res.x += o.x; res.y += o.y; }
                    if (flags & FM_STREAM_STORES) __stcs(dst, res);      // written once, never re-read
                    else *dst = res;
                }
            }
        }
    }
}

// Tile limits and kernel flags.  The defaults are the measured best; the
// environment overrides exist for tools/tune_fma.py (B2_FMA_TUNE=1 re-reads them
// on every launch so that one process can sweep configurations).
struct fm_tune { int tmax, maxm, maxn, flags, grid_mult, rm; };
static fm_tune tune_defaults() { return fm_tune{FM_TMAX, FM_MAXM, FM_MAXN, FM_DEFAULT_FLAGS, 0, FM_DEFAULT_RM}; }
static fm_tune get_tune() {
    static int mode = -1;            // 0: defaults, 1: re-read every launch
    static fm_tune cached = tune_defaults();
    if (mode == 0) return cached;
    if (mode < 0) mode = getenv("B2_FMA_TUNE") ? 1 : 0;
    if (mode == 0) return cached;
    fm_tune t = tune_defaults();
    auto rd = [](const char* k, int dflt) { const char* v = getenv(k); return v ? atoi(v) : dflt; };
    t.tmax = rd("B2_FMA_TMAX", t.tmax);
    t.maxm = rd("B2_FMA_MAXM", t.maxm);
    t.maxn = rd("B2_FMA_MAXN", t.maxn);
    t.flags = rd("B2_FMA_FLAGS", t.flags);
    t.grid_mult = rd("B2_FMA_GRID", t.grid_mult);
    t.rm = rd("B2_FMA_RM", t.rm);
    if (t.rm != 2 && t.rm != 4) t.rm = 1;
    if (t.maxm > 1024) t.maxm = 1024;
    if (t.maxn > 64) t.maxn = 64;
    if (t.tmax > 65536) t.tmax = 65536;
    return t;
}

// Classify the modes of the direct-kernel descriptor and build the tile.
// Returns false when this kernel does not apply (caller falls back).
// The planner merges stride-compatible modes (a contiguous 2^21 run arrives as
// ONE mode); a tile needs finer grain, so output modes are first split back into
// factors of at most 4 (the slow part keeps stride * factor).
static void refine_modes(const b2_modes& in, b2_modes& out) {
    int p = 0;
    const int budget = B2_MAX_MODES - in.n_red;
    for (int i = 0; i < in.n_out; ++i) {
        long long d = in.dim[i];
        long long sa = in.sa[i], sb = in.sb[i], sc = in.sc[i];
        // pieces are emitted slow -> fast, as the surrounding list is
        long long piece[32], psa[32], psb[32], psc[32];
        int np = 0;
        while (d > 4 && np < 30 && p + np + (in.n_out - i) < budget - 1) {
            int f = 1;
            for (int c = 4; c >= 2; --c) if (d % c == 0) { f = c; break; }
            if (f == 1) break;                           // odd prime extent (3, 5, 7): keep whole
            piece[np] = f; psa[np] = sa; psb[np] = sb; psc[np] = sc; ++np;   // fast part
            sa *= f; sb *= f; sc *= f; d /= f;
        }
        piece[np] = d; psa[np] = sa; psb[np] = sb; psc[np] = sc; ++np;       // slowest part
        for (int q = np - 1; q >= 0; --q) {
            out.dim[p] = (int)piece[q]; out.sa[p] = psa[q]; out.sb[p] = psb[q]; out.sc[p] = psc[q];
            ++p;
        }
    }
    out.n_out = p;
    out.n_red = in.n_red;
    for (int m = 0; m < in.n_red; ++m) {
        out.dim[p + m] = in.dim[in.n_out + m]; out.sa[p + m] = in.sa[in.n_out + m];
        out.sb[p + m] = in.sb[in.n_out + m]; out.sc[p + m] = in.sc[in.n_out + m];
    }
}

static bool make_fma(const b2_modes& md, fm_desc& d, long long& n_tiles, bool swap_ab,
                     size_t elem_bytes, const fm_tune& tn_) {
    const int FM_TMAX = tn_.tmax, FM_MAXM = tn_.maxm, FM_MAXN = tn_.maxn;
    const int no = md.n_out;
    long long K = 1;
    for (int m = 0; m < md.n_red; ++m) K *= md.dim[no + m];
    if (no == 0 || K > FM_KMAX) return false;
    int KP = 1;
    while (KP < K) KP <<= 1;
    int order[B2_MAX_MODES];
    int cnt = 0;
    for (int i = 0; i < no; ++i) if (md.dim[i] > 1) order[cnt++] = i;
    for (int i = 1; i < cnt; ++i) {                     // sort by C stride ascending
        int x = order[i], j = i - 1;
        while (j >= 0 && md.sc[order[j]] > md.sc[x]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = x;
    }
    if (cnt == 0 || md.sc[order[0]] != 1) return false;
    auto SA = [&](int i) { return swap_ab ? md.sb[i] : md.sa[i]; };
    auto SB = [&](int i) { return swap_ab ? md.sa[i] : md.sb[i]; };
    auto is_row = [&](int i) { return SA(i) != 0; };                 // M or batch mode
    auto is_bat = [&](int i) { return SA(i) != 0 && SB(i) != 0; };
    auto is_n = [&](int i) { return SA(i) == 0 && SB(i) != 0; };
    bool inner[B2_MAX_MODES] = {false};
    long long T = 1;
    int TM = 1, TN = 1, TB = 1;
    auto fits = [&](int i) {
        const long long dd = md.dim[i];
        if (T * dd > FM_TMAX) return false;
        long long tb = TB, tn = TN;
        if (is_row(i)) {
            if ((long long)TM * dd > FM_MAXM) return false;
            if (is_bat(i)) tb *= dd;
        } else if (is_n(i)) {
            tn *= dd;
            if (tn > FM_MAXN) return false;
        } else return false;
        if (tb > FM_MAXTB) return false;
        return (long long)KP * tb * tn * (long long)elem_bytes <= FM_BTILE_BYTES;
    };
    auto take = [&](int i) {
        if (is_row(i)) { TM *= md.dim[i]; if (is_bat(i)) TB *= md.dim[i]; } else TN *= md.dim[i];
        T *= md.dim[i];
        inner[i] = true;
    };
    // rows first: row modes in C order until a warp's lanes are all rows
    bool m_open = true, n_open = true;
    while (m_open && TM < 64) {
        int pick = -1;
        for (int q = 0; q < cnt; ++q) { const int i = order[q]; if (!inner[i] && is_row(i)) { pick = i; break; } }
        if (pick < 0 || !fits(pick)) { m_open = false; break; }
        take(pick);
    }
    // shared (batch) modes next, wherever they sit in C: with all of them inside
    // the tile the B block is the same for every tile (loaded once per CTA)
    for (int q = 0; q < cnt; ++q) {
        const int i = order[q];
        if (!inner[i] && is_bat(i) && fits(i)) take(i);
    }
    if (TM >= FM_MAXM) m_open = false;
    // then columns (reuse of the A rows held in registers), then more rows
    while (m_open || n_open) {
        const bool want_n = n_open && (TN < FM_U || !m_open || TN < TM);
        int pick = -1;
        for (int q = 0; q < cnt; ++q) {
            const int i = order[q];
            if (inner[i]) continue;
            if (want_n ? is_n(i) : is_row(i)) { pick = i; break; }
        }
        if (pick < 0 || !fits(pick)) { if (want_n) n_open = false; else m_open = false; continue; }
        take(pick);
    }
    if (TM < 16) return false;
    int p = 0;
    auto put = [&](int i) {
        d.dim[p] = md.dim[i]; d.sa[p] = SA(i); d.sb[p] = SB(i); d.sc[p] = md.sc[i]; ++p;
    };
    n_tiles = 1;
    d.n_outer = 0;
    d.b_static = 1;
    for (int q = 0; q < cnt; ++q) {
        const int i = order[q];
        if (!inner[i]) { put(i); ++d.n_outer; n_tiles *= md.dim[i]; if (SB(i) != 0) d.b_static = 0; }
    }
    d.n_ir = 0;
    for (int q = 0; q < cnt; ++q) { const int i = order[q]; if (inner[i] && is_row(i)) { put(i); ++d.n_ir; } }
    d.n_in = 0;
    for (int q = 0; q < cnt; ++q) { const int i = order[q]; if (inner[i] && is_n(i)) { put(i); ++d.n_in; } }
    d.n_k = 0;
    for (int i = no; i < no + md.n_red; ++i) if (md.dim[i] > 1) { put(i); ++d.n_k; }
    // (any fixed enumeration of the reduction modes works: A and B offsets are
    // decoded from the same flat k)
    d.TM = TM; d.TN = TN; d.TB = TB; d.K = (int)K;
    return true;
}

template <typename T>
int launch_fma_t(const fm_desc& d, long long n_tiles, const void* a, const void* b, void* c,
                 double ar, double ai, int accumulate, cudaStream_t s, const fm_tune& tune) {
    using C = typename cplx<T>::type;
    int sms = num_sms();
    if (sms <= 0) return 1;
    int KP = 1;
    while (KP < d.K) KP <<= 1;
    const size_t smem = (size_t)(2 * d.TM + 2 * d.TN + ((d.TB + 1) & ~1) + 2 * KP) * 8 +
                        (size_t)KP * d.TB * d.TN * sizeof(C) + (size_t)d.TM * 4 + 16;
    if (smem > (size_t)FM_SMEM_MAX) { set_error("fma: tile tables need %zu B of shared memory", smem); return 1; }
    const C* A = (const C*)a; const C* B = (const C*)b; C* Cc = (C*)c;
    // tables (<= 21 KB) + the B block (<= 32 KB) can exceed the 48 KB default
#define B2_FMA_LAUNCH_R(KPV, RV) B2_FMA_LAUNCH_RM(KPV, RV, 1)
#define B2_FMA_LAUNCH_RM(KPV, RV, RMV)                                                       \
    do {                                                                                     \
        static b2::once_flags attr_done;                                                       \
        if (!b2::is_done(attr_done)) {                                                                    \
            cudaError_t e = cudaFuncSetAttribute(contract_fma_kernel<T, KPV, RV, RMV>,       \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 FM_SMEM_MAX);                               \
            if (e != cudaSuccess) {                                                          \
                set_error("fma smem attr: %s", cudaGetErrorString(e));                       \
                return 1;                                                                    \
            }                                                                                \
            b2::mark_done(attr_done);                                                                \
        }                                                                                    \
        /* persistent grid = exactly one resident wave (every CTA gets an equal share   \
           of the tiles; a partial second wave would idle most of the GPU) */           \
        int occ = 0;                                                                         \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                   \
                &occ, contract_fma_kernel<T, KPV, RV, RMV>, FM_THREADS, smem) != cudaSuccess ||\
            occ < 1)                                                                         \
            occ = 1;                                                                         \
        long long blocks = n_tiles;                                                          \
        if (tune.grid_mult > 0 && tune.grid_mult < occ) occ = tune.grid_mult;                \
        const long long cap = (long long)sms * occ;                                          \
        if (blocks > cap) blocks = cap;                                                      \
        contract_fma_kernel<T, KPV, RV, RMV><<<(unsigned)blocks, FM_THREADS, smem, s>>>(     \
            d, n_tiles, A, B, Cc, (T)ar, (T)ai, accumulate, tune.flags);                     \
    } while (0)
    // several tiles in flight per thread only where one item is little work
    const bool narrow = d.b_static && d.TN <= 8;
    // several rows per thread (B reuse) for wide tiles without batch modes inside
    int rm = 1;
    if (!narrow && d.TB == 1 && d.TN >= 8) {
        rm = tune.rm;
        while (rm > 1 && (d.TM % (32 * rm) != 0)) rm >>= 1;
    }
#define B2_FMA_LAUNCH(KPV, RV)                                                               \
    do {                                                                                     \
        if (narrow) B2_FMA_LAUNCH_R(KPV, RV);                                                \
        else if (rm == 4) B2_FMA_LAUNCH_RM(KPV, 1, 4);                                       \
        else if (rm == 2) B2_FMA_LAUNCH_RM(KPV, 1, 2);                                       \
        else B2_FMA_LAUNCH_R(KPV, 1);                                                        \
    } while (0)
    switch (KP) {
    case 1: B2_FMA_LAUNCH(1, 4); break;
    case 2: B2_FMA_LAUNCH(2, 4); break;
    case 4: B2_FMA_LAUNCH(4, 2); break;
    case 8: B2_FMA_LAUNCH_R(8, 1); break;
    default: B2_FMA_LAUNCH_R(16, 1); break;
    }
#undef B2_FMA_LAUNCH_R
#undef B2_FMA_LAUNCH_RM
#undef B2_FMA_LAUNCH
    return check_launch("b2_contract_direct(fma)");
}

int launch_fma(const b2_modes& md, const void* a, const void* b, void* c, double ar, double ai,
               int accumulate, int dtype, cudaStream_t s, bool& handled) {
    handled = false;
    if (md.n_out == 0) return 0;
    b2_modes fine;
    refine_modes(md, fine);
    // orientation: rows should own the fastest C mode; if that leaves too few
    // rows (the fastest mode belongs to a tiny operand), try the other way round
    int fastest = -1;
    for (int i = 0; i < fine.n_out; ++i)
        if (fine.dim[i] > 1 && fine.sc[i] == 1) fastest = i;
    if (fastest < 0) return 0;
    const size_t eb = dtype == B2_C128 ? 16 : 8;
    fm_desc d, d2;
    long long n_tiles = 0, n_tiles2 = 0;
    bool swap_ab = fine.sa[fastest] == 0;
    fm_tune tune = get_tune();
    if (!getenv("B2_FMA_TUNE")) {
        // the 1024 x 64 tiles, the streaming stores and 4 rows per thread pay on outputs
        // far larger than L2 (the multi-GB final pair of a density tensor).  A mid-size
        // output (a batched plan's tens-of-MB intermediates, measured on cfg5a) needs
        // more, smaller tiles to fill the GPU, and its consumer finds it in L2 only if
        // it is stored with the default policy.
        long long total = 1;
        for (int i = 0; i < fine.n_out; ++i) total *= fine.dim[i];
        if (total < (long long)FM_TMAX * 1184) { tune.tmax = 16384; tune.rm = 2; }
        if (total * (long long)eb < (256ll << 20))
            tune.flags &= ~(FM_STREAM_STORES | FM_ROUND_ROBIN);   // contiguous tiles: B / A reuse in L1
    }
    bool ok = make_fma(fine, d, n_tiles, swap_ab, eb, tune);
    if (!ok || d.TM < 64) {
        const bool ok2 = make_fma(fine, d2, n_tiles2, !swap_ab, eb, tune);
        if (ok2 && (!ok || d2.TM > d.TM)) { d = d2; n_tiles = n_tiles2; swap_ab = !swap_ab; ok = true; }
    }
    if (!ok) return 0;
    const void* pa = swap_ab ? b : a;
    const void* pb = swap_ab ? a : b;
    handled = true;
    {
        static thread_local char info[96];
        snprintf(info, sizeof(info), "fma TM=%d TN=%d TB=%d K=%d tiles=%lld bstatic=%d swap=%d rm=%d",
                 d.TM, d.TN, d.TB, d.K, n_tiles, d.b_static, (int)swap_ab, tune.rm);
        set_last_kernel(info);
    }
    if (dtype == B2_C128)
        return launch_fma_t<double>(d, n_tiles, pa, pb, c, ar, ai, accumulate, s, tune);
    return launch_fma_t<float>(d, n_tiles, pa, pb, c, ar, ai, accumulate, s, tune);
}

}  // namespace b2
