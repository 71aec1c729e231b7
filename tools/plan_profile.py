"""Where does the device time of a bench workload go?  Leaf build, the small-step
program, and every remaining launch with its kernel variant, shape and roofline.
    python tools/plan_profile.py cfg1|cfg2|cfg3|cfg4|cfg5a|cfg5b
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from optyx_b200 import executor  # noqa: E402
from optyx_b200.core import network as nw  # noqa: E402
from optyx_b200.core.backends import B200Backend  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = bench.workload(name)
backend = B200Backend(target_size=w["target"], trials=w["trials"], min_slices=w.get("min_slices", 1))
if w["kind"] == "batch":
    nets, _ = backend.compile_batch(w["make"], w["batch"])
    for net in nets:
        net.leaves.append(nw.Leaf(nw.K_DENSE, (), (), 0, np.array([net.scalar], dtype=complex), "scalar"))
        net.scalar = 1.0 + 0j
    batch = w["batch"]
else:
    nets = [backend._get_network(w["factory"](), nested_eval=backend._nested_eval)]
    batch = 1
thr = tuple(float(x) for x in os.environ["B2_GEMM_THRESHOLD"].split(",")) if "B2_GEMM_THRESHOLD" in os.environ \
    else (32, 16, 8, 4.0)       # experiment knob: planner's (M, N, K, flop/B) bar for the DMMA kernels
plan = executor.get_plan(nets[0], batch=batch, target_size=w["target"], trials=w["trials"],
                         min_slices=w.get("min_slices", 1), gemm_threshold=thr)
sess = executor.Session(plan)
sess.set_leaves(nets)
sess.build_leaves()
sess.contract(nets[0].scalar, slice_range=(0, 1) if plan.sliced else None)
torch.cuda.synchronize()
stream = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


t_leaf = timed(sess.build_leaves)
print(f"{name}: leaves {len(nets[0].leaves)} steps {len(plan.steps)} batch {batch} nslices {plan.nslices}")
print(f"leaf build           {t_leaf*1e6:9.1f} us   ({plan.arena_elems*16*batch/t_leaf/1e9:.0f} GB/s)")
if plan.small is not None:
    sp = plan.small
    t_small = timed(lambda: sess._launch_small_groups(stream))
    macs = sum(plan.steps[k].macs for k in sp.step_ids)
    nbytes = sum(plan.steps[k].bytes for k in sp.step_ids)
    longest = max(int(sp.blob[sp.blob_off[t] + 3]) for t in range(sp.n_tasks))
    staged = sum(int(sp.blob[sp.blob_off[t] + 4]) for t in range(sp.n_tasks))
    print(f"small-step program   {t_small*1e6:9.1f} us   {len(sp.step_ids)} steps in {sp.n_tasks} tasks "
          f"(longest {longest}, {staged} with staged tables), smem {sp.smem_bytes(16)/1024:.0f} KB, {macs:.3g} MACs, "
          f"{nbytes/t_small/1e9:.0f} GB/s algorithmic")
vals = [0] * len(plan.sliced)
rows = []
for k, st in enumerate(plan.steps):
    if st.grouped:
        continue
    t = timed(lambda: sess._launch(st, vals, 1.0 + 0j, 0, stream), 10)
    ideal = max(st.bytes / 6.556e12, 8 * st.macs / 37e12)
    rows.append((t, ideal, st, sess.lib.b2_last_kernel().decode(), k))
tot = sum(r[0] * (plan.nslices if r[2].per_slice else 1) for r in rows)
print(f"other launches: {len(rows)}, summed {tot*1e3:.3f} ms (x nslices for per-slice steps)")
for t, ideal, st, kn, k in sorted(rows, key=lambda r: -r[0])[:40]:
    print(f"  #{k:4d} {kn.split()[0]:12s} {'S' if st.per_slice else ' '} mnk={str(st.mnk):34s} t={t*1e6:9.1f} us "
          f"ideal={ideal*1e6:8.2f} eff={ideal/t:5.2f} GB/s={st.bytes/t/1e9:7.0f} TF={8*st.macs/t/1e12:6.2f}")
