import sys, time
sys.path.insert(0, '/root/repo')
from oracle import bindings
from ssvio_b200 import synth
g = synth.make_config("cfg3")
shim = bindings.ShimHarness()
shim.set_write_back_every_iteration(False)
import os
for i in range(4):
    r = shim.optimize(g)["report"]
    print("optimize %.2f ms  initializeOptimization %.2f ms chi2 %.6f" % (1e3*r.seconds_total, 1e3*r.seconds_setup, r.chi2_robust))
os.environ["SSBA_TIMING"]="1"
