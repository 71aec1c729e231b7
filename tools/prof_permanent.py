"""One evaluation of the full-size cfg3 through PermanentBackend.prob_dist_measured (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from optyx_b200 import workloads as cases
from optyx_b200.core.backends import PermanentBackend
pb = PermanentBackend()
p = pb.prob_dist_measured(cases.cfg3_interferometer(width=12, n_photons=6, measured=4, inflate=False))
print(pb.last_configs, pb.last_kernel_ms, sum(p.values()))
