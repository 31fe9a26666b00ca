"""On-disk formats either side of the path, tree-input side: the XML documents of
``mergerTreeConstructorFullySpecified`` (source/merger_trees/construct/fully_specified.F90:204-370).

A document holds one or more ``<tree>`` elements (or a single ``<initialConditions>`` root); every ``<node>`` carries
``<index>``, ``<parent>``, ``<firstChild>``, ``<sibling>`` (-1 = none; ``<firstSatellite>`` optional) and one element per
component whose children are property values (``componentBuilder``: abundances are an element with one child per element,
here ``<metals>``).  ``read_fully_specified`` turns such a document into the flat arrays the C-ABI takes:

* ``forest``  -- parent / mass / time / scale_radius / angular_momentum (``glc_forest_evolve``), nodes of all trees
  concatenated, parents re-indexed; the reference's own test trees (testSuite/parameters/reproducibility/*Tree.xml) are
  what the parity tests feed through this reader;
* ``records`` / ``flags`` -- node records (``enum glc_prop``) with every component property the document sets, for
  ``glc_evolve_batch`` (a branch tip evolved towards its parent: ``time_end`` = the parent's time).

``write_fully_specified`` is the inverse (records -> document), so that a forest of this repository can be handed to the
reference (``<mergerTreeConstructor value="fullySpecified">``).  The HDF5 output side (merger_trees/outputter/standard.F90)
is not restated: there is no HDF5 library in this image.
"""
import xml.etree.ElementTree as ET

import numpy as np

from . import abi

P = abi.P

# component / property of the reference -> record column (tests/test_layout.py checks the same names against the generators)
PROPERTY_COLUMNS = {
    ("basic", "time"): "TIME", ("basic", "mass"): "BASIC_MASS", ("basic", "timeLastIsolated"): "TIME_LAST_ISOLATED",
    ("darkMatterProfile", "scale"): "DMSCALE", ("spin", "angularMomentum"): "SPIN",
    ("satellite", "boundMass"): "SAT_BOUND_MASS",
    ("blackHole", "mass"): "BH_MASS", ("blackHole", "spin"): "BH_SPIN",
    ("disk", "massStellar"): "DISK_MASS_STELLAR", ("disk", "massGas"): "DISK_MASS_GAS", ("disk", "angularMomentum"): "DISK_ANGMOM",
    ("disk", "abundancesStellar"): "DISK_ABUND_STELLAR", ("disk", "abundancesGas"): "DISK_ABUND_GAS",
    ("disk", "radius"): "DISK_RADIUS", ("disk", "velocity"): "DISK_VELOCITY",
    ("spheroid", "massStellar"): "SPH_MASS_STELLAR", ("spheroid", "massGas"): "SPH_MASS_GAS",
    ("spheroid", "angularMomentum"): "SPH_ANGMOM", ("spheroid", "abundancesStellar"): "SPH_ABUND_STELLAR",
    ("spheroid", "abundancesGas"): "SPH_ABUND_GAS", ("spheroid", "radius"): "SPH_RADIUS", ("spheroid", "velocity"): "SPH_VELOCITY",
    ("hotHalo", "mass"): "HH_MASS", ("hotHalo", "abundances"): "HH_ABUND", ("hotHalo", "angularMomentum"): "HH_ANGMOM",
    ("hotHalo", "outflowedMass"): "HH_OUTFLOWED_MASS", ("hotHalo", "outflowedAngularMomentum"): "HH_OUTFLOWED_ANGMOM",
    ("hotHalo", "outflowedAbundances"): "HH_OUTFLOWED_ABUND", ("hotHalo", "unaccretedMass"): "HH_UNACCRETED_MASS",
    ("hotHalo", "unaccretedAbundances"): "HH_UNACCRETED_ABUND", ("hotHalo", "outerRadius"): "HH_OUTER_RADIUS",
    ("hotHalo", "strippedMass"): "HH_STRIPPED_MASS", ("hotHalo", "strippedAbundances"): "HH_STRIPPED_ABUND",
}
COMPONENT_FLAGS = {"disk": abi.GLC_F_HAS_DISK, "spheroid": abi.GLC_F_HAS_SPHEROID, "hotHalo": abi.GLC_F_HAS_HOTHALO,
                   "blackHole": abi.GLC_F_HAS_BH}
_INDEX_TAGS = ("index", "parent", "firstChild", "sibling", "firstSatellite")


class TreeFormatError(ValueError):
    """The document breaks a rule fully_specified.F90 enforces (missing index, several roots, unknown parent ...)."""


def _index(node, tag, required=True):
    found = node.findall(tag)
    if len(found) > 1:
        raise TreeFormatError(f"multiple <{tag}> indices specified")  # fully_specified.F90:406
    if not found:
        if required:
            raise TreeFormatError(f"required index <{tag}> not specified")  # :409
        return -1
    return int(found[0].text.strip())


def _value(element):
    """A property value: plain number, or an abundances element (one child per tracked element: <metals> at configs[0])."""
    children = list(element)
    if not children:
        return float(element.text.strip())
    metals = element.find("metals")
    if metals is None:
        raise TreeFormatError(f"<{element.tag}>: only <metals> is tracked by this component set")
    return float(metals.text.strip())


def read_fully_specified(path_or_text):
    """Parse a fullySpecified document.  Returns a dict with ``forest`` (flat arrays for glc_forest_evolve), ``records``
    [n][NPROP], ``flags`` [n], ``time_end`` [n] (the parent's time; a root's own time), ``index`` [n] (the document's node
    indices), ``tree`` [n] (0-based tree number) and ``unknown`` (component properties this record layout has no column for)."""
    text = path_or_text
    if "<" not in str(path_or_text):
        with open(path_or_text) as f:
            text = f.read()
    root = ET.fromstring(text)
    trees = [root] if root.tag in ("tree", "initialConditions") and root.findall("node") else root.findall(".//tree")
    if not trees:
        raise TreeFormatError("no <tree> element found")
    parent, mass, time, scale, angmom, index, tree_of = [], [], [], [], [], [], []
    rows, flags, unknown = [], [], []
    for t, tree in enumerate(trees):
        nodes = tree.findall("node")
        if not nodes:
            raise TreeFormatError("no nodes were specified")  # :305
        base = len(index)
        local = {}
        for k, nd in enumerate(nodes):
            i = _index(nd, "index")
            if i in local:
                raise TreeFormatError(f"node index {i} appears twice")
            local[i] = base + k
        roots = 0
        for nd in nodes:
            pi = _index(nd, "parent")
            _index(nd, "firstChild")
            _index(nd, "sibling")
            _index(nd, "firstSatellite", required=False)
            if pi >= 0 and pi not in local:
                raise TreeFormatError(f"parent {pi} is not a node of the tree")
            roots += pi < 0
            row = np.zeros(abi.NPROP)
            row[P["TIME_STEP"]] = -1.0
            fl = 0
            for comp in nd:
                if comp.tag in _INDEX_TAGS:
                    continue
                fl |= COMPONENT_FLAGS.get(comp.tag, 0)
                for prop in comp:
                    col = PROPERTY_COLUMNS.get((comp.tag, prop.tag))
                    if col is None:
                        unknown.append((comp.tag, prop.tag))
                        continue
                    row[P[col]] = _value(prop)
            if nd.find("basic") is None or nd.find("basic/time") is None or nd.find("basic/mass") is None:
                raise TreeFormatError("every node needs <basic><time> and <basic><mass>")
            parent.append(local[pi] if pi >= 0 else -1)
            mass.append(row[P["BASIC_MASS"]])
            time.append(row[P["TIME"]])
            scale.append(row[P["DMSCALE"]])
            angmom.append(row[P["SPIN"]])
            index.append(_index(nd, "index"))
            tree_of.append(t)
            rows.append(row)
            flags.append(fl)
        if roots > 1:
            raise TreeFormatError("multiple root nodes found in the tree")  # :357
        if roots == 0:
            raise TreeFormatError("no root node was found")  # :370
    n = len(index)
    records = np.array(rows).reshape(n, abi.NPROP)
    parent = np.array(parent, dtype=np.int32)
    time = np.array(time)
    # what the interpolating node operators set at initialisation for a node evolved towards its parent (no mass growth is
    # implied by the document: targets = the node's own values)
    records[:, P["MASS_TARGET"]] = records[:, P["BASIC_MASS"]]
    records[:, P["DMSCALE_TARGET"]] = records[:, P["DMSCALE"]]
    records[:, P["SPIN_TARGET"]] = records[:, P["SPIN"]]
    records[:, P["SAT_BOUND_MASS"]] = np.where(records[:, P["SAT_BOUND_MASS"]] > 0.0, records[:, P["SAT_BOUND_MASS"]], records[:, P["BASIC_MASS"]])
    time_end = np.where(parent >= 0, time[np.maximum(parent, 0)], time)
    records[:, P["TIME_TARGET"]] = time_end
    forest = {"parent": parent, "mass": np.array(mass), "time": time, "scale_radius": np.array(scale),
              "angular_momentum": np.array(angmom)}
    return {"forest": forest, "records": records, "flags": np.array(flags, dtype=np.int32), "time_end": time_end,
            "index": np.array(index, dtype=np.int64), "tree": np.array(tree_of, dtype=np.int32), "unknown": sorted(set(unknown))}


def write_fully_specified(forest, records=None, flags=None, tree=None):
    """The inverse of read_fully_specified: an XML document (str) with one <tree> per root of ``forest``.  Node indices are
    1-based positions in the flat arrays; children are ordered by descending mass (the first one is the primary progenitor,
    as in Forest::init).  With ``records`` / ``flags`` the components a node has are written out too."""
    parent = np.asarray(forest["parent"])
    n = parent.shape[0]
    children = [[] for _ in range(n)]
    for i in range(n):
        if parent[i] >= 0:
            children[parent[i]].append(i)
    for c in children:
        c.sort(key=lambda k: (-forest["mass"][k], k))
    root_of = np.full(n, -1, dtype=np.int64)
    for i in range(n):
        r = i
        while parent[r] >= 0:
            r = parent[r]
        root_of[i] = r
    by_component = {}
    for (comp, prop), col in PROPERTY_COLUMNS.items():
        by_component.setdefault(comp, []).append((prop, col))
    doc = ET.Element("trees")
    for r in [i for i in range(n) if parent[i] < 0]:
        te = ET.SubElement(doc, "tree")
        for i in np.nonzero(root_of == r)[0]:
            nd = ET.SubElement(te, "node")
            sib = -1
            if parent[i] >= 0:
                sibs = children[parent[i]]
                k = sibs.index(i)
                sib = sibs[k + 1] + 1 if k + 1 < len(sibs) else -1
            for tag, v in (("index", i + 1), ("parent", parent[i] + 1 if parent[i] >= 0 else -1),
                           ("firstChild", children[i][0] + 1 if children[i] else -1), ("sibling", sib)):
                ET.SubElement(nd, tag).text = str(int(v))
            b = ET.SubElement(nd, "basic")
            ET.SubElement(b, "time").text = repr(float(forest["time"][i]))
            ET.SubElement(b, "mass").text = repr(float(forest["mass"][i]))
            ET.SubElement(ET.SubElement(nd, "darkMatterProfile"), "scale").text = repr(float(forest["scale_radius"][i]))
            ET.SubElement(ET.SubElement(nd, "spin"), "angularMomentum").text = repr(float(forest["angular_momentum"][i]))
            if records is None:
                continue
            for comp, bit in COMPONENT_FLAGS.items():
                if not (int(flags[i]) & bit):
                    continue
                ce = ET.SubElement(nd, comp)
                for prop, col in by_component[comp]:
                    if col in ("DISK_RADIUS", "DISK_VELOCITY", "SPH_RADIUS", "SPH_VELOCITY"):
                        continue  # structure-solver outputs, not initial conditions
                    pe = ET.SubElement(ce, prop)
                    if "bundances" in prop:
                        ET.SubElement(pe, "metals").text = repr(float(records[i, P[col]]))
                    else:
                        pe.text = repr(float(records[i, P[col]]))
    ET.indent(doc)
    return '<?xml version="1.0" encoding="UTF-8"?>\n' + ET.tostring(doc, encoding="unicode") + "\n"
