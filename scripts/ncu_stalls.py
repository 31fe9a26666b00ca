"""Stall samples by function region of a kernel.  usage: ncu_stalls.py REP"""
import csv, subprocess, sys, collections, re, os, glob, tempfile
rep = sys.argv[1]
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s >/dev/null 2>&1" % (tmp, os.path.abspath(os.environ.get("GLC_PROFILE_LIB", "galacticus_b200/libglcb200.so"))), shell=True)
cub = [c for c in glob.glob(tmp + "/*.cubin") if "params" not in c][0]
sym = subprocess.run(["readelf", "-sW", cub], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); kname = rows[0][1]; hdr = rows[1]; data = rows[2:]
col = {h: i for i, h in enumerate(hdr)}
# function symbols inside the kernel's text section: find the kernel symbol to get its section index
funcs = []
ksec = None
for l in sym.splitlines():
    f = l.split()
    if len(f) >= 8 and f[3] == "FUNC" and f[6].isdigit():
        funcs.append((int(f[1], 16), int(f[2], 0), f[6], f[7]))
base = int(data[0][col["Address"]], 16)
# the kernel entry symbol has value 0 in its section; choose section of the symbol whose name contains 'machine_kernel' or 'evolve_kernel'
target = "machine_kernel" if "machine" in kname else ("drain_kernel" if "drain" in kname else "evolve_kernel")
for v, sz, sec, name in funcs:
    if target in name and ("ModelStandard" in name or "machine" in name): ksec = sec
regions = sorted([(v, sz, name) for v, sz, sec, name in funcs if sec == ksec])
def region(off):
    best = "kernel body"
    for v, sz, name in regions:
        if v <= off < v + sz and v != 0: best = name
    best = best.split("$")[-1]
    return re.sub(r"^_ZN3glc\d*|^_Z\d*|E?NS_7SlotRef.*|__cuda_sm20_", "", best)[:40]
agg = collections.defaultdict(lambda: collections.Counter())
keys = ["# Samples", "stall_no_inst", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_barrier", "stall_lg", "stall_math", "stall_selected", "Instructions Executed", "Thread Instructions Executed"]
tot = collections.Counter()
for r in data:
    if len(r) < len(hdr): continue
    off = int(r[col["Address"]], 16) - base
    g = region(off)
    for k in keys:
        v = int(r[col[k]] or 0); agg[g][k] += v; tot[k] += v
print("%-42s %7s %7s %7s %7s %7s %7s %9s %5s" % ("region", "samp%", "noinst", "longsb", "wait", "barr", "sel", "winstr%", "thr"))
for g, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"]):
    s = max(c["# Samples"], 1)
    print("%-42s %6.1f%% %6.1f%% %6.1f%% %6.1f%% %6.1f%% %6.1f%% %8.1f%% %5.1f" % (g, 100 * c["# Samples"] / tot["# Samples"], 100 * c["stall_no_inst"] / s, 100 * c["stall_long_sb"] / s, 100 * c["stall_wait"] / s, 100 * c["stall_barrier"] / s, 100 * c["stall_selected"] / s, 100 * c["Instructions Executed"] / tot["Instructions Executed"], c["Thread Instructions Executed"] / max(c["Instructions Executed"], 1)))
