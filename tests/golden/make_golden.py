"""Regenerates the committed golden vectors: seeded node batches and the records the CPU checker (oracle/, parity
build: gcc -O2 -ffp-contract=off) evolves them to.  The reference itself cannot be run in this image (no Fortran
toolchain), so these vectors pin OUR restatement against accidental change; the reference's own golden values
(closedBox / leakyBox, testSuite/test-reproducibility.py:46-67) are asserted in tests/test_oracle_golden.py.

usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from galacticus_b200 import abi, synthetic  # noqa: E402
from oracle import orc  # noqa: E402
from tests import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run(name, p, props, flags, t_end, tables):
    o = orc.Oracle()
    if tables:
        synthetic.install(o, p)
    else:
        o.set_params(p)
    po, fo = props.copy(), flags.copy()
    status, interrupt, c = o.evolve_batch(po, fo, t_end, n_threads=1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), props_in=props, flags_in=flags, t_end=t_end, props_out=po,
                        flags_out=fo, status=status, interrupt=interrupt,
                        counters=np.array([c[k] for k in sorted(c)], dtype=np.int64), counter_names=np.array(sorted(c)))
    print(name, props.shape, c)


def main():
    orc.build()
    p = cases.standard_params()
    props, flags, t_end = synthetic.standard_nodes(p, 96, seed=4242)
    run("standard_96", p, props, flags, t_end, True)
    # full quickTest operator list (black-hole seed / accretion / winds, jet-power heating of the CGM)
    p = cases.standard_params(with_black_holes=True)
    props, flags, t_end = cases.standard_bh_nodes(p, 96, seed=4244)
    run("standard_bh_96", p, props, flags, t_end, True)
    from galacticus_b200.evolver import params_default

    pb = params_default(abi.GLC_MODEL_BOX)
    pb.box_timescaleStarFormation = 0.5
    pb.box_fractionOutflow = 1.0
    props, flags, t_end = cases.box_nodes(96, seed=4243, leaky=True)
    run("box_leaky_96", pb, props, flags, t_end, False)
    forest_golden()


def forest_golden():
    """Tree level: a small seeded forest and what the CPU checker's walk (oracle/orc_tree.c) makes of it."""
    p = cases.standard_params(with_black_holes=True)
    f = synthetic.binary_split_forest(p, 5, 3.0e11, 2.5e10, seed=4245, mass_root_max=1.5e12)
    o = orc.Oracle()
    synthetic.install(o, p)
    rec, flags, state, fc, c = o.forest_evolve(f, n_threads=1)
    np.savez_compressed(os.path.join(HERE, "forest_5.npz"), records=rec, flags=flags, state=state,
                        forest_counters=np.array([fc[k] for k in sorted(fc) if k != "rounds"], dtype=np.int64),
                        forest_counter_names=np.array([k for k in sorted(fc) if k != "rounds"]),
                        counters=np.array([c[k] for k in sorted(c)], dtype=np.int64), counter_names=np.array(sorted(c)),
                        **{"in_" + k: v for k, v in f.items()})
    print("forest_5", rec.shape, fc, c)


if __name__ == "__main__":
    main()
