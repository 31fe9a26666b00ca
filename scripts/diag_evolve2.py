import sys, numpy as np
sys.path.insert(0, '.')
from galacticus_b200 import abi, synthetic
from galacticus_b200.evolver import Evolver
from oracle import orc
from tests import cases
P = abi.P
names = {v: k for k, v in P.items()}
p = cases.standard_params(); p.resolveInterruptsOnDevice = 0
ev = Evolver(0); synthetic.install(ev, p)
o = orc.Oracle(); synthetic.install(o, p)
props, flags, tend = synthetic.standard_nodes(p, 4000, seed=5)
pr = cases.standard_params()
ev2 = Evolver(0); synthetic.install(ev2, pr)
o2 = orc.Oracle(); synthetic.install(o2, pr)
pg, fg = props.copy(), flags.copy(); po, fo_ = props.copy(), flags.copy()
ev2.evolve_batch(pg, fg, tend); o2.evolve_batch(po, fo_, tend, n_threads=8)
bad = np.where((pg != po).any(axis=1))[0]
print("bad nodes", bad)
for i in bad[:4]:
    q = props[i:i+1].copy(); f = flags[i:i+1].copy(); qo = q.copy(); fo = f.copy()
    print("node", i)
    for seg in range(5):
        s1, i1, c1 = ev.evolve_batch(q, f, tend[i:i+1]); s2, i2, c2 = o.evolve_batch(qo, fo, tend[i:i+1])
        d = np.where(q[0] != qo[0])[0]
        print("  seg", seg, "int", i1[0], i2[0], "t", repr(q[0, P['TIME']]), repr(qo[0, P['TIME']]), c1['steps_accepted'], c1['steps_rejected'], c1['rhs_evaluations'], "|", c2['steps_accepted'], c2['steps_rejected'], c2['rhs_evaluations'], "diff cols", [names[j] for j in d[:6]])
        if i1[0] == 0 and i2[0] == 0: break
        for code, bit in ((abi.GLC_INT_HOTHALO_CREATE, abi.GLC_F_HAS_HOTHALO), (abi.GLC_INT_DISK_CREATE, abi.GLC_F_HAS_DISK), (abi.GLC_INT_SPHEROID_CREATE, abi.GLC_F_HAS_SPHEROID)):
            if i1[0] == code: f[0] |= bit
            if i2[0] == code: fo[0] |= bit
