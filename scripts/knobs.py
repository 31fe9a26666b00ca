"""Execution-knob experiment: one warm-up pass and one logged pass of the bench workload under the GLC_* environment
knobs given on the command line (KEY=VALUE ...), optionally with another build of the library (LIB=path).
usage: python scripts/knobs.py N [LIB=path] [GLC_DRAIN_BELOW=30000] ..."""
import os
import sys

sys.path.insert(0, '.')
n = int(sys.argv[1])
lib = None
for kv in sys.argv[2:]:
    k, v = kv.split('=', 1)
    if k == 'LIB':
        lib = v
    else:
        os.environ[k] = v
import bench  # noqa: E402
from galacticus_b200 import evolver, synthetic  # noqa: E402

if lib:
    evolver.LIB_PATH = os.path.abspath(lib)
p, props, flags, tend = bench.workload(n, 219)
ev = evolver.Evolver(0)
synthetic.install(ev, p)
ev.arena_upload(props, flags, tend)
ev.arena_snapshot(n)
ev.evolve_arena(n)
ev.arena_restore(n)
os.environ.get('GLC_SLICE_LOG') and sys.stderr.write('--- logged pass\n')
c, ms = ev.evolve_arena(n)
print('KNOBS', ' '.join(sys.argv[2:]) or 'default', 'n', n, 'ms %.1f' % ms, 'rhs/s %.3e' % (c['rhs_evaluations'] / ms * 1e3),
      'steps/s %.3e' % (c['steps_accepted'] / ms * 1e3), flush=True)
