/*
 * optyx_b200 -- C ABI of the B200-native evaluation backend for Optyx diagrams.
 *
 * This is the drop-in boundary for the ONE hot path of the reference:
 *   diagram.eval(backend)  ->  dense leaf arrays  ->  pairwise contractions
 *   ->  (sliced sum)  ->  prob_dist
 * (reference: optyx/core/backends.py:457-550 QuimbBackend.eval/_process_term,
 *  optyx/core/diagram.py:536-590 Box.truncation, backends.py:196-356
 *  EvalResult.prob_dist).  The reference reaches its arithmetic through
 * quimb/cotengra/numpy; a maintainer binds these entry points with ctypes
 * (see INTEGRATION.md).
 *
 * Conventions
 *  - every pointer named *_dev is DEVICE memory owned by the caller (torch
 *    CUDA tensors -> data_ptr()); the library never allocates or frees it;
 *  - every entry point is stream-ordered on `stream` (a cudaStream_t passed as
 *    void*), launches kernels only, and performs no host synchronisation;
 *  - return value 0 = ok, non-zero = error; b2_last_error() gives the text;
 *  - complex numbers are interleaved (re, im); dtype B2_C128 = 2 x f64,
 *    B2_C64 = 2 x f32;  tensors are C-ordered unless strides are given;
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with an error.
 */
#ifndef OPTYX_B200_H
#define OPTYX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_ABI_VERSION 1

#define B2_C128 0
#define B2_C64 1

/* ---- leaf kinds: what Box.truncation builds (diagram.py:536-590) -------- */
#define B2_K_ONEHOT 1    /* Create/Select           zw.py:450-460, 567-573   */
#define B2_K_ONES 2      /* free end of a spider    diagram.py:705-707       */
#define B2_K_VEC 3       /* Z-box diagonal (host values) zw.py:320-329, 641-646 */
#define B2_K_DENSE 4     /* array boxes (host values)    diagram.py:542-548  */
#define B2_K_SUM 5       /* W: sqrt(multinomial)    zw.py:231-244            */
#define B2_K_ADD 6       /* Add                     zw.py:676-683            */
#define B2_K_MUL 7       /* Multiply                zw.py:720-727            */
#define B2_K_DIV 8       /* Divide                  zw.py:763-772            */
#define B2_K_MOD2 9      /* Mod2                    zw.py:818-825            */
#define B2_K_DUALRAIL 10 /* DualRail                diagram.py:904-913       */
#define B2_K_PTD 11      /* PhotonThresholdDetector diagram.py:975-983       */
#define B2_K_EMBED 12    /* EmbeddingTensor         diagram.py:1005-1027     */

#define B2_MAX_LEAF_RANK 8

/* One leaf of the arena.  `offset`/`nelem` are in complex elements inside ONE
 * arena (one eval); `poff` indexes the per-eval params buffer (complex128). */
typedef struct {
    int32_t kind;
    int32_t ndim;
    int32_t dims[B2_MAX_LEAF_RANK];
    int32_t ipar;
    int32_t shared;     /* VEC / DENSE: 1 = poff indexes the SHARED values (same for every
                           eval of a batch), 0 = the per-eval params buffer               */
    int64_t offset;
    int64_t nelem;
    int64_t poff;
} b2_leaf_desc;

/* ---- pairwise contraction ------------------------------------------------
 * C[out modes] (+)= alpha * sum_{red modes} A[...] * B[...]
 * Modes 0..n_out-1 index C (C order: last fastest), modes n_out..n_out+n_red-1
 * are summed.  A mode absent from an operand has stride 0 (broadcast).  This
 * one form covers what quimb/cotengra do per tree node (transpose + tensordot
 * / einsum with batch indices; reference call sites backends.py:514, 539-545),
 * hyper-indices (spiders), diagonal views (an index wired twice to one leaf:
 * strides add) and index slicing (base pointer offsets).                     */
#define B2_MAX_MODES 64

typedef struct {
    int32_t n_out;
    int32_t n_red;
    int32_t dim[B2_MAX_MODES];
    int64_t sa[B2_MAX_MODES];   /* element strides into A (0 = absent)        */
    int64_t sb[B2_MAX_MODES];   /* element strides into B (0 = absent)        */
    int64_t sc[B2_MAX_MODES];   /* element strides into C (out modes only)    */
} b2_modes;

/* GEMM-shaped view of the same contraction for the tiled tensor-core kernel:
 * modes are grouped [batch | M | N | K]; within a group mode 0 is the fastest
 * varying in the flattened group index.                                      */
typedef struct {
    int32_t n_batch, n_m, n_n, n_k;
    int32_t dim[B2_MAX_MODES];
    int64_t sa[B2_MAX_MODES];
    int64_t sb[B2_MAX_MODES];
    int64_t sc[B2_MAX_MODES];
} b2_gemm_modes;

/* version / errors */
int b2_abi_version(void);
const char* b2_last_error(void);
/* which kernel variant the last contraction call of this thread launched
 * ("fma ...", "tiled", "reduce", "generic", "ctile", "dot", "gemm_ws" / "gemm_ws32" /
 * "gemm_ws16" / "gemm_ws_splitk", "gemm_thin", "gemm3m[_splitk]", "gemm[_splitk]",
 * "gemm_c64_tcgen05", "small_program"): profiling aid */
const char* b2_last_kernel(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int b2_device_sms(void);

/* (1) array build: fill `batch` arenas of `arena_elems` complex elements each
 * from `n_leaves` descriptors (device table, sorted by offset).
 * params_dev: complex128 host-computed values, `params_stride` elements per
 * eval (0 = shared by all evals of the batch).                               */
int b2_build_leaves(const b2_leaf_desc* desc_dev, int32_t n_leaves,
                    int64_t arena_elems, const void* params_dev,
                    int64_t params_stride, void* arena_dev,
                    int64_t arena_stride, int32_t batch, int32_t dtype,
                    void* stream);

/* the same with a plan-time index: block_leaf_dev[b] = leaf holding arena element
 * 256 * b (one trailing entry: the leaf holding the last element), so that a thread
 * searches only the leaves of its own 256-element block; shared_params_dev holds the
 * host values of the leaves flagged `shared` (one copy for the whole batch)           */
int b2_build_leaves_indexed(const b2_leaf_desc* desc_dev, int32_t n_leaves,
                            const int32_t* block_leaf_dev, int64_t arena_elems,
                            const void* params_dev, int64_t params_stride,
                            const void* shared_params_dev, void* arena_dev,
                            int64_t arena_stride, int32_t batch, int32_t dtype, void* stream);

/* (2a) thin / bandwidth-bound pairs: one thread per output element.
 * accumulate != 0: C += alpha * (...)                                        */
int b2_contract_direct(const b2_modes* modes, const void* a_dev,
                       const void* b_dev, void* c_dev, double alpha_re,
                       double alpha_im, int32_t accumulate, int32_t dtype,
                       void* stream);

/* (2b) large dense pairs: complex GEMM with fused index permutation on load
 * and store.  complex128: gather -> shared memory -> FP64 DMMA (mma.sync, no
 * tcgen05 kind exists for f64).  complex64: gather + hi/lo split -> shared
 * memory (UMMA layout) -> tcgen05.mma kind::tf32, 3 split products per k-step,
 * FP32 accumulators in tensor memory.                                        */
int b2_contract_gemm(const b2_gemm_modes* modes, const void* a_dev,
                     const void* b_dev, void* c_dev, double alpha_re,
                     double alpha_im, int32_t accumulate, int32_t dtype,
                     void* stream);

/* (2c) launch-bound small networks (hundreds of leaves of <= 343 elements; the
 * reference pays one numpy call per tree node, backends.py:539-545): EVERY early tiny
 * pairwise contraction of an eval -- or of a whole batch of evals -- in ONE persistent
 * launch, intermediates in SHARED MEMORY, no launch per step.
 *
 * The plan's tiny steps are cut into TASKS (chains of steps one CTA executes back to
 * back: leaves and hand-overs are imported from global memory into the CTA's scratch,
 * every step reads and writes the scratch, a block barrier between steps, results that
 * leave the task are exported to global memory).  Work items are (eval, task) pairs in
 * topological order; persistent CTAs take them from a ticket counter, wait on the
 * done-flags of the tasks they depend on (a dependency always holds a smaller ticket,
 * so whoever holds it is already running: no deadlock, no co-residency requirement),
 * and publish their own flag with release semantics.  Flags carry the launch EPOCH, so
 * a CUDA-graph replay needs no reset: the last CTA to finish bumps the epoch and
 * rewinds the ticket counter.
 *
 * A step is   C[oc(o)] = sum_r A[oa(o) + ra(r)] * B[ob(o) + rb(r)]   with the operand
 * offsets of every output index o and every reduction index r TABULATED at plan time
 * (element offsets inside the scratch, < 2^16): no index arithmetic on the device.
 *
 * Per task ONE blob of 32-bit words, copied into shared memory in a single coalesced
 * pass when the task starts (a step never begins with a global-memory round trip):
 *   header   [n_dep, n_imp, n_exp, n_steps, staged, 0, 0, 0]
 *   deps     n_dep task ids (padded to an even count)
 *   copies   (n_imp + n_exp) x 8 words: space (0 leaf arena, 1 workspace, 2 output),
 *            scratch element offset, elements, 0, global element offset (lo, hi),
 *            eval stride (lo, hi): address = base[space] + offset + eval * stride
 *   steps    n_steps x 8 words: n_out, n_red, a_off, b_off, c_off (scratch elements),
 *            rsplit (lanes sharing one output's reduction, 2^k <= 32), out_tab, red_tab
 *   tables   when staged: out_tab / red_tab are WORD offsets into this blob (out rows
 *            are 2 words: oa | ob << 16, oc; red rows 1 word: ra | rb << 16); otherwise
 *            they are ROW offsets into the global tables below.
 * Shared memory of a CTA: [blob, 16-byte aligned][scratch].                           */
typedef struct {
    const uint32_t* blob_dev;        /* all task blobs, back to back                    */
    const int32_t* blob_off_dev;     /* [n_tasks + 1] word offsets                      */
    const uint32_t* out_tab_dev;     /* tables of the tasks that do not stage theirs    */
    const uint32_t* red_tab_dev;
    void* base_dev[3];               /* leaf arena, workspace, output                   */
    int32_t* sync_dev;               /* [0] ticket [1] finished CTAs [2] epoch [3] pad,
                                        then batch * n_tasks done-flags; zero-filled ONCE */
    int32_t n_tasks;
    int32_t batch;
    int32_t smem_bytes;              /* dynamic shared memory of the launch (largest task) */
    int32_t reserved;
} b2_small_program;

int b2_run_small_program(const b2_small_program* prog, int32_t dtype, void* stream);

/* (2d) a whole compiled plan in ONE call: the host-side loop of the reference
 * (cotengra's contract_tree / gather_slices walking the tree node by node and slice by
 * slice, reached from backends.py:539-545) as a C loop over a step table.  Order:
 * optional zero-fill of the output, the small-step program, the slice-invariant steps,
 * then for every slice id in [slice_begin, slice_end) the per-slice steps (operand base
 * pointers offset by slice value x stride; the final step accumulates into the output).
 * The table lives in HOST memory and holds no pointers: the three buffers are bound per
 * call, so one table serves any number of buffer sets.                               */
#define B2_STEP_DIRECT 0
#define B2_STEP_GEMM 1
#define B2_MAX_SLICED 16

typedef struct {
    int32_t kind;                       /* B2_STEP_DIRECT / B2_STEP_GEMM               */
    int32_t per_slice;                  /* 1: runs once per slice id                   */
    int32_t final;                      /* 1: scaled by alpha (and accumulated when sliced) */
    int32_t a_space, b_space, c_space;  /* 0 leaf arena, 1 workspace, 2 output         */
    int32_t n_a_slice, n_b_slice;
    int64_t a_off, b_off, c_off;        /* element offsets inside the space            */
    int64_t c_elems;                    /* extent of the result (all evals of a batch) */
    int32_t a_slice_pos[B2_MAX_SLICED]; /* which sliced index ...                      */
    int32_t b_slice_pos[B2_MAX_SLICED];
    int64_t a_slice_stride[B2_MAX_SLICED]; /* ... times its element stride             */
    int64_t b_slice_stride[B2_MAX_SLICED];
    union { b2_modes direct; b2_gemm_modes gemm; } desc;
} b2_plan_step;

/* what to run: bit 0 zero-fill the output, bit 1 small-step program + invariant steps,
 * bit 2 per-slice steps                                                              */
#define B2_RUN_ZERO 1
#define B2_RUN_INVARIANT 2
#define B2_RUN_SLICES 4

int b2_run_plan(const b2_plan_step* steps, int32_t n_steps, const int32_t* sliced_dims,
                int32_t n_sliced, const b2_small_program* small, void* arena_dev,
                void* ws_dev, void* out_dev, int64_t out_elems, int32_t what,
                int64_t slice_begin, int64_t slice_end, double alpha_re, double alpha_im,
                int32_t dtype, void* stream, int32_t* launches_out);
/* workspace the table needs, in bytes (max over workspace results of offset + extent) */
int64_t b2_workspace_bytes(const b2_plan_step* steps, int32_t n_steps, int32_t dtype);

/* (3) slice / Sum-term accumulation: y += alpha * x over n complex elements  */
int b2_axpy(int64_t n, double alpha_re, double alpha_im, const void* x_dev,
            void* y_dev, int32_t dtype, void* stream);
int b2_fill_zero(int64_t n, void* y_dev, int32_t dtype, void* stream);

/* (4) prob_dist (backends.py:291-356).
 * AMP: p[i] = |t[i]|^2, nz[i] = (t[i] != 0)            (_prob_dist_pure)
 * DM : axis kinds per result axis: 1 = measured (bit/mode), 2 = ket half,
 *      3 = bra half of an unmeasured quantum wire (always adjacent pairs).
 *      p[meas] = sum over unmeasured diagonal of Re t, with the reference's
 *      tiny-negative clamp; seen[meas] = any non-zero entry with that key;
 *      max_imag_dev[0] = max |Im| over the entries that were summed.
 * p_dev is float64 for both dtypes.                                          */
int b2_probs_amp(int64_t n, const void* t_dev, double* p_dev,
                 uint8_t* nz_dev, int32_t dtype, void* stream);
int b2_probs_dm(int32_t ndim, const int32_t* dims, const int32_t* axis_kind,
                const void* t_dev, double* p_dev, uint8_t* seen_dev,
                double* max_imag_dev, int32_t dtype, void* stream);

/* (4b) result hand-over for sparse results (the array EvalResult wraps,
 * backends.py:480-494, is dense on the host; most of a GHZ / Clifford density
 * tensor or a photon-number-conserving amplitude tensor is exact zeros): one
 * read pass appends (flat C-order index, value) of every entry with re != 0 or
 * im != 0 (numpy's flatnonzero test) to idx_dev / val_dev (cap entries each, unordered) and counts them in
 * count_dev[0] (zeroed by this call).  count > cap means the list overflowed
 * and its contents are incomplete: the caller copies the tensor densely.      */
int b2_compact_nonzero(int64_t n, const void* t_dev, int32_t dtype, int64_t cap,
                       int64_t* idx_dev, void* val_dev, uint64_t* count_dev,
                       void* stream);

/* (5) permanents: amplitudes of a pure linear-optical circuit (the reference's
 * PermanentBackend, backends.py:1019-1098; Matrix.eval path.py:345-384; npperm
 * path.py:141-163).  u_dev: m x m single-particle matrix (complex128, row
 * major); rows_dev[n]: input mode of each created photon (device, at least one
 * element); cols_dev[n_configs][n]:
 * output mode of each detected photon, per configuration; norm_dev[n_configs]:
 * 1 / sqrt(prod in! out!).  amp_dev[c] = norm[c] * perm(U[rows, cols_c])
 * (Glynn's formula in Gray-code order, one warp per configuration, the 2^(n-1)
 * sign vectors split over the lanes).  n <= B2_PERM_MAX_N, m <= B2_PERM_MAX_M. */
#define B2_PERM_MAX_N 30
#define B2_PERM_MAX_M 64
int b2_permanent_amplitudes(const void* u_dev, int32_t m, const int32_t* rows_dev,
                            int32_t n, const uint8_t* cols_dev,
                            const double* norm_dev, int64_t n_configs,
                            void* amp_dev, void* stream);

/* (5b) the same for LOSSY / partially distinguishable interferometers measured on
 * some modes: probabilities of the detected photon counts, marginalised on the
 * device over every unmeasured mode (loss environments, discarded modes, internal
 * states).  v_dev: n x M complex128, row i = single-particle output amplitudes of
 * photon i over ALL M output modes; configurations (multisets of n output modes)
 * are enumerated on the device in lexicographic order from their rank
 * (binom_dev[(a) * (n + 1) + b] = C(a, b), a <= M + n); every configuration adds
 * scale * |perm|^2 / prod(out!) to hist_dev[sum_j mode_stride_dev[mode_j]]
 * (hist_dev zero-filled by the caller, n_bins entries).                       */
#define B2_PERM_HIST_MAX_M 128
int b2_permanent_histogram(const void* v_dev, int32_t n, int32_t m,
                           const int64_t* binom_dev,
                           const int32_t* mode_stride_dev, double scale,
                           int64_t n_configs, int32_t n_bins, double* hist_dev,
                           void* stream);

/* (7) compressed (approximate) contraction, SURVEY.md 8(f) row f4: what quimb's
 * contract_compressed does for the reference when QuimbBackend holds a
 * HyperCompressedOptimizer (backends.py:516-545).  The reductions over everything but the
 * bond (Gram matrices G_A = A^H A, G_B = B B^H) and the application of the projectors are
 * ordinary pairwise contractions (entry points (2a) / (2b)); b2_conj provides conj(A);
 * b2_bond_compress is the small dense part, one CTA per bond: from the two bond x bond
 * Gram matrices (row-major complex128) it writes the oblique projectors P_A [bond][bond]
 * (columns 0..k-1 valid) and P_B [bond][bond] (rows 0..k-1 valid) with
 * A P_A P_B B = the rank-k truncation of A B, k = number of singular values of the bond
 * matrix that exceed cutoff * largest, at most max_bond (<= 0: no cap), at least 1;
 * *k_dev receives k, sv_dev[0..k) the kept singular values in descending order.
 * Factorisations: one-sided Jacobi (round-robin pair order, one warp per column pair). */
int b2_conj(int64_t n, const void* x_dev, void* y_dev, int32_t dtype, void* stream);
int64_t b2_bond_compress_scratch_bytes(int32_t bond);
int b2_bond_compress(const void* ga_dev, const void* gb_dev, int32_t bond, int32_t max_bond,
                     double cutoff, void* pa_dev, void* pb_dev, int32_t* k_dev, double* sv_dev,
                     void* scratch_dev, void* stream);

/* (6) host side of the planner (CPU only, no kernels): the randomised greedy pair-order
 * search and the cost of a pair order under slicing, for networks of thousands of
 * tensors (the reference hands this to cotengra, backends.py:436-455).  Tensors are index
 * lists in CSR form over index ids 0..n_inds-1; hyper-indices and open indices allowed.
 * b2_plan_greedy writes the SSA pair order (2 * (n - 1) ids, step k creates id n + k);
 * b2_plan_cost writes [multiply-adds per slice, largest intermediate (elements),
 * number of slices, roofline time estimate in seconds] and, optionally, log2 of every
 * step's result size.                                                                */
int b2_plan_greedy(int32_t n, const int32_t* ind_ptr, const int32_t* inds, int32_t n_inds,
                   const int32_t* dims, const uint8_t* is_output, uint64_t seed,
                   double temperature, double costmod, int32_t* path_out);
int b2_plan_cost(int32_t n, const int32_t* ind_ptr, const int32_t* inds, int32_t n_inds,
                 const int32_t* dims, const uint8_t* is_output, const int32_t* path,
                 const uint8_t* is_sliced, double* out, double* step_sizes_out);
/* subtree reconfiguration: every subtree of at most max_leaves (3..10) frontier nodes gets
 * the optimal internal pair order (dynamic programming over subsets; multiply-adds with
 * the sliced indices fixed); the SSA path is rewritten in place                          */
int b2_plan_reconfigure(int32_t n, const int32_t* ind_ptr, const int32_t* inds, int32_t n_inds,
                        const int32_t* dims, const uint8_t* is_output, const uint8_t* is_sliced,
                        int32_t max_leaves, int32_t* path, int32_t* improved_out);

/* micro-benchmarks used by bench.py to measure the FP64 pipes of this GPU:
 * kind 0 = DFMA, 1 = DMMA m8n8k4, 2 = DMMA m16n8k16 (falls back to 1 if the
 * shape is unavailable).  Launches `iters` dependent-chain iterations on a
 * full grid; returns the number of FLOPs one launch performs in *flops_out.  */
int b2_fp64_peak_kernel(int32_t kind, int32_t iters, double* sink_dev,
                        double* flops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OPTYX_B200_H */
