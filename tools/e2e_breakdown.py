"""Where does the end-to-end time of diagram.eval(B200Backend()) go (cfg2)?"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from optyx_b200 import workloads as cases  # noqa: E402
from optyx_b200 import executor  # noqa: E402
from optyx_b200.core.backends import B200Backend  # noqa: E402

backend = B200Backend()
keep = []
for it in range(5):
    t = [time.perf_counter()]
    d = cases.ghz(20, open_qubits=14, noise=0.95); t.append(time.perf_counter())
    net = backend._get_network(d, nested_eval=backend._nested_eval); t.append(time.perf_counter())
    plan = executor.get_plan(net); t.append(time.perf_counter())
    sess = executor.Session(plan); torch.cuda.synchronize(); t.append(time.perf_counter())
    sess.set_leaves([net]); torch.cuda.synchronize(); t.append(time.perf_counter())
    sess.build_leaves(); sess.contract(net.scalar); torch.cuda.synchronize(); t.append(time.perf_counter())
    arr = backend._to_host(sess.out[0]); t.append(time.perf_counter())
    names = ["diagram", "compile", "plan(cache)", "session alloc", "set_leaves+h2d", "launch+run", "d2h pinned"]
    print(it, " ".join(f"{n}={1e3*(b-a):.1f}ms" for n, a, b in zip(names, t, t[1:])), f"total={1e3*(t[-1]-t[0]):.1f}ms")
    keep = [arr]          # drop the previous result: its pinned block returns to the pool
