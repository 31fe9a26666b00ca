import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
from galacticus_b200 import abi, synthetic
from galacticus_b200.evolver import Evolver
from oracle import orc
from tests import cases
P = abi.P
p = cases.standard_params()
ev = Evolver(0); synthetic.install(ev, p)
o = orc.Oracle(); synthetic.install(o, p)
n = 3000
props, flags, tend = synthetic.standard_nodes(p, n, seed=5)
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
ev.L.glc_debug_probe.argtypes = [C.c_void_p, C.c_int64, dp, ip, dp]
og = np.zeros((n, 16))
rc = ev.L.glc_debug_probe(ev.h, n, props, flags, og); assert rc == 0, rc
oo = np.zeros((n, 16))
o.L.orc_probe_node.argtypes = [C.POINTER(abi.glc_params), C.c_void_p, dp, C.c_int, dp]
for i in range(n):
    row = props[i].copy(); out = np.zeros(16)
    o.L.orc_probe_node(C.byref(o.params), o.T, row, int(flags[i]), out); oo[i] = out
labels = ["rvir","vvir","tvir","hhRho0","nfwM(r0)","orbitalMean","vc2bary","Mdm(r0)","bessel(.37)","Mhh(r0)","fastexp","r_from_j","sfr_disk","rcool","vtot","log(v/r)"]
for k in range(16):
    ne = og[:, k] != oo[:, k]
    rel = np.abs(og[:, k] - oo[:, k]) / (np.abs(oo[:, k]) + 1e-300)
    print("%-12s non-identical %5d  max rel %.3e" % (labels[k], ne.sum(), rel.max()))
