"""Attribute the warp-state samples of one stall reason to source lines.
usage: ncu_stall_lines.py REP KERNEL_SUBSTR [STALL_COLUMN=stall_long_sb] [TOPN]   (GLC_PROFILE_LIB = the library the capture ran)"""
import csv, subprocess, sys, collections, re, os, glob, tempfile
rep, kname = sys.argv[1], sys.argv[2]
colname = sys.argv[3] if len(sys.argv) > 3 else "stall_long_sb"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s >/dev/null 2>&1" % (tmp, os.path.abspath(os.environ.get("GLC_PROFILE_LIB", "galacticus_b200/libglcb200.so"))), shell=True)
cub = [c for c in glob.glob(tmp + "/*.cubin") if "params" not in c][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.split("\n")
start = None
for i, l in enumerate(dis):
    if l.startswith("\t.section\t.text.") and kname in l:
        start = i; break
linemap = {}; sass = {}; cur = None
pat = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
for l in dis[start + 1:]:
    if l.startswith("\t.section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m2 = pat.search(l)
    if m2 and cur:
        linemap[int(m2.group(1), 16)] = cur; sass[int(m2.group(1), 16)] = m2.group(2)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]; data = rows[2:]
col = {h: i for i, h in enumerate(hdr)}
ia = col["Address"]; ic = col[colname]; isamp = col["# Samples"]
base = int(data[0][ia], 16)
agg = collections.Counter(); tot = 0; totall = 0; byaddr = collections.Counter()
for r in data:
    if len(r) < len(hdr): continue
    a = int(r[ia], 16) - base; v = int(r[ic] or 0); tot += v; totall += int(r[isamp] or 0)
    agg[linemap.get(a, ("?", 0))] += v; byaddr[a] += v
print("%s: %d samples of %d (%.1f%%)" % (colname, tot, totall, 100.0 * tot / max(totall, 1)))
for k, v in agg.most_common(topn): print("%-28s line %4d  %5.2f%%" % (k[0], k[1], 100 * v / max(tot, 1)))
print("-- top instructions")
for a, v in byaddr.most_common(25): print("%6x %5.2f%%  %-26s:%4d  %s" % (a, 100 * v / max(tot, 1), linemap.get(a, ("?", 0))[0], linemap.get(a, ("?", 0))[1], sass.get(a, "")[:90]))
