"""Dumps the ODE serialization order of the reference's node components by RUNNING the reference's own component generators
(python/Galacticus/Build/Components, driven as Galacticus.Build.SourceTree.Process.ComponentBuilder drives them) over every
<component> directive of source/objects/nodes/components, and writes tests/golden/serialization_order.json.

The reference tree is read here, in the build container, only; the committed JSON travels.  tests/test_layout.py checks
``enum glc_prop`` of include/glc_b200.h against it.

usage: python tests/golden/make_layout.py [/root/reference]"""
import json
import os
import subprocess
import sys
import tempfile

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "python"))
os.environ["GALACTICUS_EXEC_PATH"] = REF
os.environ["BUILDPATH"] = tempfile.mkdtemp(prefix="glc_layout_")

import Galacticus.Build.Components as Components  # noqa: E402
from Galacticus.Build.Components.Implementations import Utils as ImplUtils  # noqa: E402
from Galacticus.Build.Components.TreeNodes import ODESolver as TreeODE  # noqa: E402
from Galacticus.Build.Directives import extract_directives  # noqa: E402

files = subprocess.run(["grep", "-rlE", "<component( |>)", os.path.join(REF, "source", "objects", "nodes", "components")],
                       capture_output=True, text=True, check=True).stdout.split()
build = {}
for f in sorted(files):
    for doc in extract_directives(f, "component", force_array={"data", "property", "binding"}, include_raw_xml=True):
        doc.pop("rawXML", None)
        build["currentDocument"] = doc
        Components.parse_directive(build)
Components.generate_output(build)

# class order of treeNodeSerializeValuesToArray (TreeNodes/ODESolver.py:95-138)
classes = [c["name"] for c in TreeODE._active_classes(build)]
# the implementations selected by parameters/quickTest.xml (component<Class> value="...")
selected = {"basic": "standard", "blackHole": "standard", "darkMatterProfile": "scale", "disk": "standard", "hotHalo": "standard",
            "satellite": "standard", "spheroid": "standard", "spin": "scalar"}
out = {"reference": "galacticusorg/galacticus python/Galacticus/Build/Components (generate_output)", "class_order": classes,
       "selected_implementations": selected, "evolvable_properties": {}}
for cls in classes:
    impl = selected.get(cls)
    if impl is None:
        continue
    member = build["components"][cls[:1].upper() + cls[1:] + impl[:1].upper() + impl[1:]]
    props = []
    for p in ImplUtils.list_real_evolvers(member):
        data = p.get("data") or {}
        props.append({"name": p["name"], "type": data.get("type"), "rank": int(data.get("rank") or 0)})
    out["evolvable_properties"][cls] = props
json.dump(out, open(os.path.join(HERE, "serialization_order.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
