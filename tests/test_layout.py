"""The node record layout (``enum glc_prop`` of include/glc_b200.h) against the serialization order that the reference's OWN
component generators produce.  tests/golden/serialization_order.json is written by
tests/golden/make_layout.py, which runs python/Galacticus/Build/Components/generate_output over every <component>
directive of the reference tree and lists, in the class order of treeNodeSerializeValuesToArray
(TreeNodes/ODESolver.py:95-138), the evolvable properties of the implementations parameters/quickTest.xml selects."""
import json
import os

from galacticus_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))

# reference property -> GLC_P_ name; abundances objects are one scalar ("metals") at configs[0]
NAMES = {
    ("blackHole", "mass"): "BH_MASS", ("blackHole", "spin"): "BH_SPIN",
    ("disk", "massStellar"): "DISK_MASS_STELLAR", ("disk", "abundancesStellar"): "DISK_ABUND_STELLAR",
    ("disk", "massGas"): "DISK_MASS_GAS", ("disk", "abundancesGas"): "DISK_ABUND_GAS", ("disk", "angularMomentum"): "DISK_ANGMOM",
    ("hotHalo", "mass"): "HH_MASS", ("hotHalo", "abundances"): "HH_ABUND", ("hotHalo", "angularMomentum"): "HH_ANGMOM",
    ("hotHalo", "outflowedMass"): "HH_OUTFLOWED_MASS", ("hotHalo", "outflowedAngularMomentum"): "HH_OUTFLOWED_ANGMOM",
    ("hotHalo", "outflowedAbundances"): "HH_OUTFLOWED_ABUND", ("hotHalo", "unaccretedMass"): "HH_UNACCRETED_MASS",
    ("hotHalo", "unaccretedAbundances"): "HH_UNACCRETED_ABUND", ("hotHalo", "outerRadius"): "HH_OUTER_RADIUS",
    ("hotHalo", "strippedMass"): "HH_STRIPPED_MASS", ("hotHalo", "strippedAbundances"): "HH_STRIPPED_ABUND",
    ("satellite", "boundMass"): "SAT_BOUND_MASS",
    ("spheroid", "massStellar"): "SPH_MASS_STELLAR", ("spheroid", "abundancesStellar"): "SPH_ABUND_STELLAR",
    ("spheroid", "massGas"): "SPH_MASS_GAS", ("spheroid", "abundancesGas"): "SPH_ABUND_GAS", ("spheroid", "angularMomentum"): "SPH_ANGMOM",
}
# zero-length at configs (no chemicals, no luminosity filters, no histories: SURVEY 8a)
EMPTY_TYPES = {"chemicalAbundances", "stellarLuminosities", "history"}
# marked analytic by the quickTest node operators (removed from the ODE state by pack(.not.nodeAnalytics),
# TreeNodes/ODESolver.py:262-275): cosmicTime, DMOInterpolate, darkMatterProfileScaleInterpolate,
# haloAngularMomentumInterpolate, starFormationDisks/Spheroids (massStellarFormed), barInstability (fractionMassRetained)
ANALYTIC = {("basic", "mass"), ("basic", "time"), ("spin", "angularMomentum"), ("darkMatterProfile", "scale"),
            ("disk", "massStellarFormed"), ("disk", "fractionMassRetained"), ("spheroid", "massStellarFormed")}


def test_state_vector_order_matches_the_generators():
    ref = json.load(open(os.path.join(HERE, "golden", "serialization_order.json")))
    expected = []
    for cls in ref["class_order"]:
        for prop in ref["evolvable_properties"].get(cls, []):
            key = (cls, prop["name"])
            if prop["type"] in EMPTY_TYPES or key in ANALYTIC:
                continue
            assert prop["rank"] == 0 and prop["type"] in ("double", "abundances"), key
            assert key in NAMES, f"reference property {key} has no place in the node record"
            expected.append(NAMES[key])
    ours = sorted((v, k) for k, v in abi.P.items() if v < abi.NY)
    assert [k for _, k in ours] == expected
    assert len(expected) == abi.NY == 24


def test_analytic_properties_have_record_words():
    # the analytically solved properties travel in the second part of the record
    for name in ("TIME", "BASIC_MASS", "DMSCALE", "SPIN"):
        assert abi.P[name] >= abi.NY
