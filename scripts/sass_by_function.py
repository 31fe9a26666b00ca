"""Static SASS size of one kernel by source function: every instruction of `nvdisasm -g` output is attributed to the
innermost source line (-lineinfo), and lines are mapped to the enclosing function definition of the header they are in.
usage: cuobjdump -xelf all libglcb200.so; nvdisasm -g glc_api.sm_100a.cubin > all.sass; python scripts/sass_by_function.py all.sass KERNEL_SUBSTRING"""
import collections
import re
import sys

sass, key = sys.argv[1], sys.argv[2]
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and key in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('.text.')), len(lines))
fdef = re.compile(r'^\s*(?:template\s*<[^>]*>\s*)?(?:GLC_[A-Z_]+|static|inline|__device__|__forceinline__|__global__)[\w\s\*&:<>,]*?\b(\w+)\s*\(')
funcs = {}


def enclosing(path, line):
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


cur = ('?', 0)
count = collections.Counter()
marker = re.compile(r'//## File "([^"]+)", line (\d+)')
instr = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/')
for l in lines[start:end]:
    m = marker.search(l)
    if m:
        cur = (m.group(1), int(m.group(2)))
    elif instr.match(l):
        count[(cur[0].split('/')[-1], enclosing(*cur))] += 1
total = sum(count.values())
print('%s: %d instructions (%.0f KB)' % (key, total, total * 16 / 1024))
for (f, fn), c in count.most_common(45):
    print('%6d %5.1f%%  %s:%s' % (c, 100.0 * c / total, f, fn))
