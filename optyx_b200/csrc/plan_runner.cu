// Piece (2d): the plan interpreter of the C ABI.  One call walks a compiled plan --
// zero-fill, small-step program, slice-invariant steps, then the per-slice steps of
// every slice id of this rank -- and launches every kernel on the caller's stream.
// The reference's counterpart is a Python loop per tree node and per slice inside
// cotengra (reached from optyx/core/backends.py:539-545).
#include "b2_common.cuh"

extern "C" int64_t b2_workspace_bytes(const b2_plan_step* steps, int32_t n_steps, int32_t dtype) {
    if (!steps || n_steps <= 0) return 0;
    int64_t top = 0;
    for (int32_t k = 0; k < n_steps; ++k)
        if (steps[k].c_space == 1 && steps[k].c_off + steps[k].c_elems > top)
            top = steps[k].c_off + steps[k].c_elems;
    return top * (dtype == B2_C128 ? 16 : 8);
}

extern "C" int b2_run_plan(const b2_plan_step* steps, int32_t n_steps, const int32_t* sliced_dims,
                           int32_t n_sliced, const b2_small_program* small, void* arena_dev,
                           void* ws_dev, void* out_dev, int64_t out_elems, int32_t what,
                           int64_t slice_begin, int64_t slice_end, double alpha_re,
                           double alpha_im, int32_t dtype, void* stream, int32_t* launches_out) {
    using namespace b2;
    if (dtype != B2_C128 && dtype != B2_C64) { set_error("b2_run_plan: bad dtype %d", dtype); return 1; }
    if (n_steps < 0 || n_sliced < 0 || n_sliced > B2_MAX_SLICED || (n_steps > 0 && !steps)) {
        set_error("b2_run_plan: bad table (n_steps=%d n_sliced=%d)", n_steps, n_sliced);
        return 1;
    }
    const int64_t esize = dtype == B2_C128 ? 16 : 8;
    char* base[3] = {(char*)arena_dev, (char*)ws_dev, (char*)out_dev};
    int launches = 0;
    if ((what & B2_RUN_ZERO) && out_elems > 0) {
        if (b2_fill_zero(out_elems, out_dev, dtype, stream)) return 1;
    }
    int64_t vals[B2_MAX_SLICED] = {0};
    auto launch = [&](const b2_plan_step& st, int accumulate) -> int {
        for (int sp = 0; sp < 3; ++sp) {
            const bool used = st.a_space == sp || st.b_space == sp || st.c_space == sp;
            if (used && !base[sp]) { set_error("b2_run_plan: buffer %d is null", sp); return 1; }
        }
        int64_t a = st.a_off, b = st.b_off;
        for (int i = 0; i < st.n_a_slice; ++i) a += vals[st.a_slice_pos[i]] * st.a_slice_stride[i];
        for (int i = 0; i < st.n_b_slice; ++i) b += vals[st.b_slice_pos[i]] * st.b_slice_stride[i];
        const void* ap = base[st.a_space] + a * esize;
        const void* bp = base[st.b_space] + b * esize;
        void* cp = base[st.c_space] + st.c_off * esize;
        const double ar = st.final ? alpha_re : 1.0, ai = st.final ? alpha_im : 0.0;
        ++launches;
        if (st.kind == B2_STEP_GEMM)
            return b2_contract_gemm(&st.desc.gemm, ap, bp, cp, ar, ai, accumulate, dtype, stream);
        if (st.kind == B2_STEP_DIRECT)
            return b2_contract_direct(&st.desc.direct, ap, bp, cp, ar, ai, accumulate, dtype, stream);
        set_error("b2_run_plan: bad step kind %d", st.kind);
        return 1;
    };
    if (what & B2_RUN_INVARIANT) {
        if (small && small->n_tasks > 0) {
            b2_small_program sp = *small;          // buffers are bound per call
            sp.base_dev[0] = arena_dev; sp.base_dev[1] = ws_dev; sp.base_dev[2] = out_dev;
            if (b2_run_small_program(&sp, dtype, stream)) return 1;
            ++launches;
        }
        for (int32_t k = 0; k < n_steps; ++k)
            if (!steps[k].per_slice && launch(steps[k], 0)) return 1;
    }
    if ((what & B2_RUN_SLICES) && n_sliced > 0) {
        if (!sliced_dims) { set_error("b2_run_plan: sliced_dims is null"); return 1; }
        for (int64_t sid = slice_begin; sid < slice_end; ++sid) {
            int64_t rem = sid;
            for (int i = n_sliced - 1; i >= 0; --i) { vals[i] = rem % sliced_dims[i]; rem /= sliced_dims[i]; }
            for (int32_t k = 0; k < n_steps; ++k)
                if (steps[k].per_slice && launch(steps[k], steps[k].final ? 1 : 0)) return 1;
        }
    }
    if (launches_out) *launches_out = launches;
    return 0;
}
