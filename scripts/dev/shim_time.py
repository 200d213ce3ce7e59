"""Developer aid: wall time of optimizer.optimize(10) through the g2o shim (oracle/shim_harness.cpp) at cfg3;
SSBA_TIMING=1 adds the shim's and the library's own section timers on stderr."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import bindings
from ssvio_b200 import synth
g = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
shim = bindings.ShimHarness()
shim.set_write_back_every_iteration(False)
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 8):
    r = shim.optimize(g)["report"]
    print("optimize %.2f ms  initializeOptimization %.2f ms chi2 %.6f" % (1e3 * r.seconds_total, 1e3 * r.seconds_setup, r.chi2_robust), flush=True)
