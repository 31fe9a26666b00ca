import sys, numpy as np, shutil
sys.path.insert(0, '.')
which, node = sys.argv[1], int(sys.argv[2])
if which == "orc":
    shutil.copy("oracle/_build/liborc.so", "/tmp/liborc_save.so"); shutil.copy("oracle/_build/liborc_trace.so", "oracle/_build/liborc.so")
from galacticus_b200 import abi, synthetic
from tests import cases
P = abi.P
p = cases.standard_params(); p.resolveInterruptsOnDevice = 0
props, flags, tend = synthetic.standard_nodes(p, 4000, seed=5)
q = props[node:node+1].copy(); f = flags[node:node+1].copy()
if which == "gpu":
    from galacticus_b200.evolver import Evolver
    ev = Evolver(0); synthetic.install(ev, p)
    print(ev.evolve_batch(q, f, tend[node:node+1]))
else:
    from oracle import orc
    o = orc.Oracle(); synthetic.install(o, p)
    print(o.evolve_batch(q, f, tend[node:node+1]))
    shutil.copy("/tmp/liborc_save.so", "oracle/_build/liborc.so")
