"""Per-step CUDA-event timing of one slice of the cfg4 sliced plan: where does the
time go, and how far is each step from its roofline?"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from optyx_b200 import workloads as cases  # noqa: E402
from optyx_b200 import executor  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] in ("cfg3", "cfg1"):
    net = cases.compile_channel(cases.cfg3_interferometer(width=6, n_photons=3, measured=2)
                                if sys.argv[1] == "cfg3" else
                                cases.cfg1_interferometer(6, (1, 1, 1, 0, 0, 0)))
    import numpy as np
    dtype = np.complex64 if os.environ.get("B2_DTYPE") == "c64" else np.complex128
    plan = executor.get_plan(net, dtype=dtype)
    for g in ([plan.small] if plan.small is not None else []):                      # time every step on its own
        for k in g.step_ids:
            plan.steps[k].grouped = False
    plan.small = None
    for st in plan.steps:
        st.per_slice = True          # time every step
else:
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 41
    tgt = int(sys.argv[2]) if len(sys.argv) > 2 else 26
    net = cases.compile_channel(cases.random_clifford_t(30, depth, seed=0))
    plan = executor.get_plan(net, target_size=float(2 ** tgt), trials=16)
sess = executor.Session(plan)
sess.set_leaves([net])
sess.build_leaves()
sess.contract(net.scalar, slice_range=(0, 1)) if plan.sliced else sess.run([net])
torch.cuda.synchronize()
stream = torch.cuda.current_stream().cuda_stream
vals = [0] * len(plan.sliced)
rows = []
prof = os.environ.get("B2_PROFILE_MNK")          # e.g. "2,16777216,1,1": ncu --profile-from-start off
prof = tuple(int(x) for x in prof.split(",")) if prof else None
for st in plan.steps:
    if not st.per_slice:
        continue
    if prof is not None and tuple(st.mnk) == prof:
        d = st.desc
        counts = (d.n_out, d.n_red) if st.kernel == "direct" else (d.n_batch, d.n_m, d.n_n, d.n_k)
        n = sum(counts)
        print("profile step", st.kernel, st.mnk, "mode counts", counts, "dims", list(d.dim[:n]),
              "sa", list(d.sa[:n]), "sb", list(d.sb[:n]), "sc", list(d.sc[:n]), flush=True)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        sess._launch(st, vals, 1.0 + 0j, 0, stream)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        prof = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sess._launch(st, vals, 1.0 + 0j, 0, stream)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    ideal = max(st.bytes / 6.556e12, 8 * st.macs / (37e12 if plan.dtype.itemsize == 16 else 367e12))
    rows.append((t, ideal, st, sess.lib.b2_last_kernel().decode()))
tot = sum(r[0] for r in rows)
print(f"nslices {plan.nslices} per-slice steps {len(rows)} measured {tot*1e3:.1f} ms "
      f"ideal {sum(r[1] for r in rows)*1e3:.1f} ms")
by = {}
for t, ideal, st, kn in rows:
    k = kn.split()[0]
    by.setdefault(k, [0, 0.0, 0.0])
    by[k][0] += 1; by[k][1] += t; by[k][2] += ideal
for k, v in by.items():
    print(f"  {k}: n={v[0]} measured {v[1]*1e3:.1f} ms ideal {v[2]*1e3:.1f} ms")
print("top 30 by measured time:")
for t, ideal, st, kn in sorted(rows, key=lambda r: -r[0])[:30]:
    d = st.desc
    nm = (d.n_out, d.n_red) if st.kernel == "direct" else (d.n_batch, d.n_m, d.n_n, d.n_k)
    print(f"  {kn.split()[0]:7s} mnk={st.mnk} t={t*1e3:7.3f} ms ideal={ideal*1e3:6.3f} eff={ideal/t:5.2f} "
          f"GB/s={st.bytes/t/1e9:7.0f} modes={nm} {' '.join(kn.split()[1:])}")
small = [r for r in rows if r[1] < 5e-6]
print(f"steps with ideal < 5us: {len(small)}, measured {sum(r[0] for r in small)*1e3:.2f} ms")
