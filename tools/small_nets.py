"""Device time of the launch-bound configs (graph replay), for the small-subtree launch."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from optyx_b200 import workloads as cases
from optyx_b200 import executor, planner
def timed(fn, reps=50):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, d in (("cfg2_k10", cases.ghz(20, open_qubits=10, noise=0.95)), ("cfg2_measured", cases.ghz(20, noise=0.95, measure=True)),
                ("cfg1", cases.cfg1_interferometer(6, (1, 1, 1, 0, 0, 0))), ("hom", cases.lossy_hom(0.8))):
    net = cases.compile_channel(d)
    for grouping, n_streams in ((True, 8), (False, 8), (True, 16), (True, 32)):
        executor.BRANCH_STREAMS = n_streams
        plan = planner.plan_network(net)
        if not grouping:
            for g in ([plan.small] if plan.small is not None else []):
                for k in g.step_ids: plan.steps[k].grouped = False
            plan.small = None
        sess = executor.Session(plan)
        sess.set_leaves([net]); sess.run([net])
        sess.capture(net.scalar)
        print(name, "streams", n_streams, "grouping", grouping, "groups", (len(plan.small.tasks) if plan.small is not None else 0), "steps", len(plan.steps), "ms %.4f" % timed(sess.replay), flush=True)
