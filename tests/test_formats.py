"""Tree-input format: the XML documents of mergerTreeConstructorFullySpecified
(source/merger_trees/construct/fully_specified.F90) read into the flat forest arrays and node records of the C-ABI
(galacticus_b200/formats.py).  When the reference tree is present its own test trees
(testSuite/parameters/reproducibility/*Tree.xml) are parsed and pushed through the checker to the reference's goldens; a
round trip through the writer covers the format without it."""
import os

import numpy as np
import pytest

from galacticus_b200 import abi, formats, synthetic
from tests import cases
from tests.test_oracle_golden import CLOSED, CLOSED_TOL, LEAKY, LEAKY_TOL

P = abi.P
REF = "/root/reference/testSuite/parameters/reproducibility"
needs_reference = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")

LEAKY_DOC = """<?xml version="1.0"?>
<tree>
  <node><index>1</index><parent>2</parent><firstChild>-1</firstChild><sibling>-1</sibling>
    <basic><time>12.47</time><mass>1.0e12</mass></basic>
    <spin><angularMomentum>1.635e12</angularMomentum></spin>
    <hotHalo><mass>0.0</mass><abundances><metals>0.0</metals></abundances></hotHalo>
    <disk><massStellar>0.0</massStellar><massGas>1.0e11</massGas>
      <abundancesStellar><metals>0.0</metals></abundancesStellar><abundancesGas><metals>0.0</metals></abundancesGas></disk>
  </node>
  <node><index>2</index><parent>-1</parent><firstChild>1</firstChild><sibling>-1</sibling>
    <basic><time>13.47</time><mass>1.0e12</mass></basic>
    <spin><angularMomentum>1.635e12</angularMomentum></spin>
  </node>
</tree>
"""


def _evolve_box(orc, doc, leaky):
    p = orc.params_default(abi.GLC_MODEL_BOX)
    if leaky:
        p.box_timescaleStarFormation = 0.5
        p.box_fractionOutflow = 1.0
    o = orc.Oracle()
    o.set_params(p)
    tips = np.nonzero(doc["forest"]["parent"] >= 0)[0]
    props = np.ascontiguousarray(doc["records"][tips])
    flags = np.ascontiguousarray(doc["flags"][tips])
    status, interrupt, _ = o.evolve_batch(props, flags, np.ascontiguousarray(doc["time_end"][tips]))
    assert status[0] == 0 and interrupt[0] == 0
    return props[0]


def test_document_to_records(oracle_lib):
    """A document in the reference's format (same content as leakyBoxTree.xml) gives the hand-built record of
    tests/cases.py, and the checker evolves it to the reference's leakyBox golden."""
    doc = formats.read_fully_specified(LEAKY_DOC)
    assert doc["forest"]["parent"].tolist() == [1, -1] and doc["index"].tolist() == [1, 2]
    assert doc["flags"].tolist() == [abi.GLC_F_HAS_DISK | abi.GLC_F_HAS_HOTHALO, 0]
    ref_props, ref_flags, ref_tend = cases.reproducibility_box(leaky=True)
    assert doc["time_end"][0] == ref_tend[0] == 13.47
    for col in ("TIME", "TIME_STEP", "DISK_MASS_GAS", "BASIC_MASS", "MASS_TARGET", "TIME_TARGET", "SPIN_TARGET"):
        assert doc["records"][0, P[col]] == ref_props[0, P[col]], col
    assert doc["unknown"] == []
    row = _evolve_box(oracle_lib, doc, leaky=True)
    for k, v in LEAKY.items():
        assert abs(row[P[k]] - v) <= LEAKY_TOL[k] * v, k


@needs_reference
def test_reference_reproducibility_trees(oracle_lib):
    closed = formats.read_fully_specified(os.path.join(REF, "closedBoxTree.xml"))
    row = _evolve_box(oracle_lib, closed, leaky=False)
    for k, v in CLOSED.items():
        assert abs(row[P[k]] - v) <= CLOSED_TOL[k] * v, k
    leaky = formats.read_fully_specified(os.path.join(REF, "leakyBoxTree.xml"))
    row = _evolve_box(oracle_lib, leaky, leaky=True)
    for k, v in LEAKY.items():
        assert abs(row[P[k]] - v) <= LEAKY_TOL[k] * v, k
    ac = formats.read_fully_specified(os.path.join(REF, "adiabaticContractionTree.xml"))
    assert ac["forest"]["parent"].tolist() == [1, -1]
    r = ac["records"][0]
    assert (r[P["TIME"]], ac["time_end"][0]) == (13.46, 13.48)
    assert (r[P["BASIC_MASS"]], r[P["DMSCALE"]], r[P["SPH_MASS_STELLAR"]], r[P["SPH_ANGMOM"]]) == (1.0e12, 0.03, 1.0e10, 1.0e10)
    assert ac["flags"][0] == abi.GLC_F_HAS_SPHEROID


def test_round_trip_forest():
    """forest -> XML -> forest: same trees, and Forest::init's progenitor order (descending mass) is what <firstChild> /
    <sibling> say."""
    p = cases.standard_params()
    f = synthetic.binary_split_forest(p, 3, 1.0e12, 2.0e10, seed=11)
    text = formats.write_fully_specified(f)
    doc = formats.read_fully_specified(text)
    order = doc["index"] - 1  # position in the original arrays
    assert sorted(order.tolist()) == list(range(f["parent"].shape[0]))
    back = {k: np.empty_like(np.asarray(f[k])) for k in ("mass", "time", "scale_radius", "angular_momentum")}
    for k in back:
        back[k][order] = doc["forest"][k]
        np.testing.assert_array_equal(back[k], f[k])
    parent_back = np.full(order.shape[0], -2, dtype=np.int64)
    parent_back[order] = np.where(doc["forest"]["parent"] >= 0, order[np.maximum(doc["forest"]["parent"], 0)], -1)
    np.testing.assert_array_equal(parent_back, f["parent"])
    assert len(set(doc["tree"].tolist())) == 3


def test_round_trip_records():
    p = cases.standard_params(with_black_holes=True)
    props, flags, _ = cases.standard_bh_nodes(p, 24, seed=2)
    n = props.shape[0]
    forest = {"parent": np.full(n, -1, dtype=np.int32), "mass": props[:, P["BASIC_MASS"]].copy(), "time": props[:, P["TIME"]].copy(),
              "scale_radius": props[:, P["DMSCALE"]].copy(), "angular_momentum": props[:, P["SPIN"]].copy()}
    doc = formats.read_fully_specified(formats.write_fully_specified(forest, props, flags))
    comp = abi.GLC_F_HAS_DISK | abi.GLC_F_HAS_SPHEROID | abi.GLC_F_HAS_HOTHALO | abi.GLC_F_HAS_BH
    np.testing.assert_array_equal(doc["flags"], flags & comp)
    for (c, prop), col in formats.PROPERTY_COLUMNS.items():
        if c in formats.COMPONENT_FLAGS and not col.endswith(("RADIUS", "VELOCITY")) or col == "HH_OUTER_RADIUS":
            has = (flags & formats.COMPONENT_FLAGS[c]) != 0
            np.testing.assert_array_equal(doc["records"][has, P[col]], props[has, P[col]], err_msg=col)


def test_malformed_documents_are_rejected():
    bad_root = LEAKY_DOC.replace("<parent>2</parent>", "<parent>-1</parent>")
    with pytest.raises(formats.TreeFormatError, match="multiple root"):
        formats.read_fully_specified(bad_root)
    with pytest.raises(formats.TreeFormatError, match="required index"):
        formats.read_fully_specified(LEAKY_DOC.replace("<sibling>-1</sibling>", "", 1))
    with pytest.raises(formats.TreeFormatError, match="not a node"):
        formats.read_fully_specified(LEAKY_DOC.replace("<parent>2</parent>", "<parent>7</parent>"))
    with pytest.raises(formats.TreeFormatError, match="no root"):
        formats.read_fully_specified(LEAKY_DOC.replace("<parent>-1</parent>", "<parent>1</parent>"))


def test_document_forest_through_the_tree_walk(oracle_lib):
    """A forest written in the reference's tree format and read back evolves, through the checker's tree walk, to exactly
    what the original arrays evolve to (the reader re-indexes the nodes; the trees are the same)."""
    p = cases.standard_params(with_black_holes=True)
    f = synthetic.binary_split_forest(p, 3, 1.0e12, 5.0e10, seed=5)
    doc = formats.read_fully_specified(formats.write_fully_specified(f))
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    ra, fa, sa, fca, ca = o.forest_evolve(f, n_threads=2)
    rb, fb, sb, fcb, cb = o.forest_evolve(doc["forest"], n_threads=2)
    order = doc["index"] - 1
    assert fca == fcb and ca == cb
    np.testing.assert_array_equal(sb, sa[order])
    np.testing.assert_array_equal(fb, fa[order])
    assert np.array_equal(rb, ra[order])


@pytest.mark.gpu
def test_document_forest_on_the_device(oracle_lib):
    """The same document through glc_forest_evolve: the device evolves the trees read from XML to the checker's records."""
    from galacticus_b200.evolver import Evolver

    p = cases.standard_params(with_black_holes=True)
    f = synthetic.binary_split_forest(p, 6, 1.0e12, 2.0e10, seed=9)
    doc = formats.read_fully_specified(formats.write_fully_specified(f))
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    ro, fo, so, fco, co = o.forest_evolve(doc["forest"], n_threads=8)
    ev = Evolver(0)
    synthetic.install(ev, p)
    rg, fg, sg, fcg, cg = ev.forest_evolve(doc["forest"])
    np.testing.assert_array_equal(sg, so)
    np.testing.assert_array_equal(fg, fo)
    assert cg == co
    alive = so != abi.GLC_FOREST_NODE_PROMOTED
    assert np.array_equal(rg[alive], ro[alive])
    ev.close()
