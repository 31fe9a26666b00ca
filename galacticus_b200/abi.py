"""ctypes mirror of ``include/glc_b200.h``.

The structs and enums are parsed from the header itself so the Python view of the
C-ABI can never drift from what the shared library was compiled against.
"""
from __future__ import annotations

import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "glc_b200.h")

_CTYPES = {
    "int32_t": C.c_int32,
    "uint32_t": C.c_uint32,
    "int64_t": C.c_int64,
    "uint64_t": C.c_uint64,
    "double": C.c_double,
    "float": C.c_float,
}


def _strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def _parse_enums(src: str) -> dict[str, int]:
    out: dict[str, int] = {}
    for body in re.findall(r"enum\s+\w+\s*\{(.*?)\}", src, flags=re.S):
        val = -1
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, expr = [s.strip() for s in item.split("=", 1)]
                expr = re.sub(r"(\d+)u\b", r"\1", expr)
                expr = re.sub(r"0x([0-9a-fA-F]+)u\b", r"0x\1", expr)
                val = int(eval(expr, {}, out))  # expressions over earlier enumerators only
            else:
                name = item
                val += 1
            out[name] = val
    return out


def _parse_struct(src: str, name: str):
    m = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), src, flags=re.S)
    if not m:
        raise RuntimeError(f"struct {name} not found in {HEADER}")
    fields = []
    for decl in m.group(1).split(";"):
        decl = decl.strip()
        if not decl:
            continue
        parts = decl.split(None, 1)
        ctype = _CTYPES[parts[0]]
        for var in parts[1].split(","):
            var = var.strip()
            m2 = re.match(r"(\w+)\[(\w+)\]$", var)
            if m2:  # fixed-size array member: the bound is a #define or an enumerator
                bound = m2.group(2)
                nelem = int(bound) if bound.isdigit() else (ENUMS.get(bound) or _DEFINES[bound])
                fields.append((m2.group(1), ctype * nelem))
            else:
                fields.append((var, ctype))
    return type(name, (C.Structure,), {"_fields_": fields})


_SRC = _strip_comments(open(HEADER).read())
_DEFINES = {k: int(v) for k, v in re.findall(r"#define\s+(\w+)\s+(\d+)\s*$", _SRC, flags=re.M)}
ENUMS = _parse_enums(_SRC)
globals().update(ENUMS)

glc_params = _parse_struct(_SRC, "glc_params")
glc_counters = _parse_struct(_SRC, "glc_counters")
glc_forest_counters = _parse_struct(_SRC, "glc_forest_counters")
glc_profile = _parse_struct(_SRC, "glc_profile")
glc_error_report = _parse_struct(_SRC, "glc_error_report")

GLC_ABI_VERSION = int(re.search(r"#define\s+GLC_ABI_VERSION\s+(\d+)", _SRC).group(1))
NPROP = ENUMS["GLC_NPROP"]
NY = ENUMS["GLC_NY"]

# property-name -> column index of the node record, e.g. P["DISK_MASS_GAS"]
P = {k[len("GLC_P_"):]: v for k, v in ENUMS.items() if k.startswith("GLC_P_")}
F = {k[len("GLC_F_"):]: v for k, v in ENUMS.items() if k.startswith("GLC_F_")}

# every entry point the header declares (used by the symbol-export test)
DECLARED_FUNCTIONS = sorted(set(re.findall(r"\b(glc_\w+)\s*\(", _SRC)))


def counters_dict(c: "glc_counters") -> dict[str, int]:
    return {name: int(getattr(c, name)) for name, _ in c._fields_}


def profile_dict(pr: "glc_profile") -> dict:
    import numpy as np

    n = int(pr.n_bins)
    out = {"n_bins": n, "time_step_smallest": float(pr.time_step_smallest), "property_hits_unknown": int(pr.property_hits_unknown)}
    for name in ("time_step", "time_step_count", "evaluation_count", "time_step_count_interrupted", "evaluation_count_interrupted"):
        out[name] = np.array(list(getattr(pr, name))[:n])
    out["property_hits"] = np.array(list(pr.property_hits))
    return out
