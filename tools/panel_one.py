"""One panel-kernel microbenchmark configuration (for ncu): python tools/panel_one.py WHAT NT BATCH"""
import sys
sys.path.insert(0, ".")
from gp_b200 import capi
h = capi.Handle(0)
what, nt, batch = (int(a) for a in sys.argv[1:4])
print("ms", h.debug_bench_panel(what, nt, batch, 3))
