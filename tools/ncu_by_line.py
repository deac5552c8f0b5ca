"""Join an .ncu-rep SASS page with nvdisasm line info of the SAME build: stall samples / executed instructions per source line.
usage: ncu_by_line.py <rep> <cubin> <kernel-substring> [top]"""
import csv, subprocess, sys, collections, re
rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
dis = subprocess.run(['nvdisasm', '-g', cubin], capture_output=True, text=True).stdout.splitlines()
# locate the .text section of the kernel
start = None
for i, l in enumerate(dis):
    if l.startswith('.text.') and kname in l and l.rstrip().endswith(':'):
        start = i; break
assert start is not None, 'kernel not found in cubin'
instrs = []   # (file, line)
cur = ('?', 0)
for l in dis[start + 1:]:
    if l.startswith('//---') or (l.startswith('.text.') and l.rstrip().endswith(':')): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+\S', l): instrs.append(cur)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout.splitlines()
k = 0; rows = None
while k < len(out):
    if out[k].startswith('"Kernel Name"') and kname_ok(out[k]) if False else out[k].startswith('"Kernel Name"'):
        name = out[k]
        hdr = next(csv.reader([out[k + 1]])); k += 2; rr = []
        while k < len(out) and not out[k].startswith('"Kernel Name"'):
            r = next(csv.reader([out[k]]))
            if len(r) == len(hdr): rr.append(dict(zip(hdr, r)))
            k += 1
        if rows is None and re.sub(r'[^A-Za-z0-9_]', '', kname.split('ILi')[0])[:20] in re.sub(r'[^A-Za-z0-9_]', '', name): rows = rr
    else: k += 1
print('sass rows', len(rows), 'disasm instrs', len(instrs))
n = min(len(rows), len(instrs))
samp = collections.Counter(); ex = collections.Counter()
for r, loc in zip(rows[:n], instrs[:n]):
    samp[loc] += int(r['# Samples']); ex[loc] += int(r['Instructions Executed'])
ts, te = sum(samp.values()), sum(ex.values())
src = {}
def getline(f, ln):
    import os
    for d in ('evennicer_slam_b200/csrc/', ''):
        p = d + f
        if os.path.exists(p):
            if p not in src: src[p] = open(p).read().splitlines()
            return src[p][ln - 1].strip()[:110] if 0 < ln <= len(src[p]) else ''
    return ''
order = ex.most_common(top) if len(sys.argv) > 5 else samp.most_common(top)
for loc, v in order:
    print(f"{100*samp[loc]/ts:5.1f}% smp {100*ex[loc]/te:5.1f}% ex  {loc[0]}:{loc[1]:<4d} {getline(*loc)}")
