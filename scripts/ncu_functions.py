"""Executed warp instructions of a profiled kernel by source FUNCTION (innermost inlined frame of -lineinfo).
usage: GLC_PROFILE_LIB=lib.so GLC_PROFILE_SRC=/path/to/tree python scripts/ncu_functions.py REP KERNEL_SUBSTR [TOPN]
(the library must be the build the report was captured with: the script checks the instruction count)"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, kname = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
libp = os.path.abspath(os.environ.get("GLC_PROFILE_LIB", "galacticus_b200/libglcb200.so"))
srcroot = os.environ.get("GLC_PROFILE_SRC")
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s >/dev/null 2>&1" % (tmp, libp), shell=True)
cub = [c for c in glob.glob(tmp + "/*.cubin") if "params" not in c][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and kname in l)
fdef = re.compile(r'^\s*(?:template\s*<[^>]*>\s*)?(?:GLC_[A-Z_]+|static|inline|__device__|__forceinline__|__global__)[\w\s\*&:<>,]*?\b(\w+)\s*\(')
funcs = {}


def enclosing(path, line):
    if srcroot and '/galacticus_b200/' in path:
        path = os.path.join(srcroot, 'galacticus_b200', path.split('/galacticus_b200/')[1])
    if path not in funcs:
        tab = []
        try:
            for n, l in enumerate(open(path), 1):
                m = fdef.match(l)
                if m and not l.strip().endswith(';'):
                    tab.append((n, m.group(1)))
        except OSError:
            pass
        funcs[path] = tab
    name = '?'
    for n, f in funcs[path]:
        if n > line:
            break
        name = f
    return name


linemap, cur = {}, None
pat = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/")
for l in dis[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m2 = pat.search(l)
    if m2 and cur:
        linemap[int(m2.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
ia, ii, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = int(data[0][ia], 16)
agg, aggt, tot, nrow = collections.Counter(), collections.Counter(), 0, 0
for r in data:
    if len(r) <= it:
        continue
    nrow += 1
    a, w, t = int(r[ia], 16) - base, int(r[ii]), int(r[it])
    tot += w
    k = linemap.get(a)
    name = (k[0].split('/')[-1] + ':' + enclosing(*k)) if k else '?'
    agg[name] += w
    aggt[name] += t
print('# %s: %d SASS instructions in the report, %d with line info in %s; %d warp instructions executed' % (kname, nrow, len(linemap), os.path.basename(libp), tot))
for k, v in agg.most_common(topn):
    print("%6.2f%%  avg thr %4.1f  %s" % (100 * v / tot, aggt[k] / max(v, 1), k))
