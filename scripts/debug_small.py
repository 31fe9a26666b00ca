import sys, numpy as np
sys.path.insert(0, '.')
from galacticus_b200 import abi, synthetic
from galacticus_b200.evolver import Evolver
from tests import cases
n = int(sys.argv[1]); seed = int(sys.argv[2]) if len(sys.argv) > 2 else 100 + n
p = cases.standard_params()
props, flags, tend = synthetic.standard_nodes(p, n, seed=seed)
ev = Evolver(0); synthetic.install(ev, p)
s, i, c = ev.evolve_batch(props, flags, tend)
print("n", n, c, "status ok", (s == 0).all(), flush=True)
