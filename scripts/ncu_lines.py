"""Attribute warp instructions of a kernel to source lines.  usage: ncu_lines.py REP KERNEL_SUBSTR [TOPN]"""
import csv, subprocess, sys, collections, re, os, glob, tempfile
rep, kname = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 45
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s >/dev/null 2>&1" % (tmp, os.path.abspath(os.environ.get("GLC_PROFILE_LIB", "galacticus_b200/libglcb200.so"))), shell=True)
cub = [c for c in glob.glob(tmp + "/*.cubin") if "params" not in c][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.split("\n")
start = None
for i, l in enumerate(dis):
    if l.startswith("\t.section\t.text.") and kname in l:
        start = i; break
linemap = {}; cur = None
pat = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/")
for l in dis[start + 1:]:
    if l.startswith("\t.section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m2 = pat.search(l)
    if m2 and cur: linemap[int(m2.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]; data = rows[2:]
ia = hdr.index("Address"); ii = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed")
base = int(data[0][ia], 16)
agg = collections.Counter(); aggt = collections.Counter(); tot = 0
for r in data:
    if len(r) <= it: continue
    a = int(r[ia], 16) - base; w = int(r[ii]); t = int(r[it]); tot += w
    k = linemap.get(a, ("?", 0)); agg[k] += w; aggt[k] += t
byfile = collections.Counter()
for k, v in agg.items(): byfile[k[0]] += v
print({k: "%.1f%%" % (100 * v / tot) for k, v in byfile.most_common()})
for k, v in agg.most_common(topn): print("%-28s line %4d  %5.2f%%  avg thr %.1f" % (k[0], k[1], 100 * v / tot, aggt[k] / max(v, 1)))
