"""Aggregate warp-stall samples of an .ncu-rep by SASS opcode class and by hot address ranges."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
lines = out.splitlines()
# find header rows
k = 0
while k < len(lines):
    if lines[k].startswith('"Kernel Name"'):
        kn = lines[k]
        hdr = next(csv.reader([lines[k + 1]]))
        rows = []
        k += 2
        while k < len(lines) and not lines[k].startswith('"Kernel Name"'):
            r = next(csv.reader([lines[k]]))
            if len(r) == len(hdr): rows.append(dict(zip(hdr, r)))
            k += 1
        print(kn[:120])
        tot = sum(int(r['# Samples']) for r in rows)
        byop = collections.Counter(); cnt = collections.Counter(); execd = collections.Counter()
        for r in rows:
            op = r['Source'].split()[0] if not r['Source'].strip().startswith('@') else r['Source'].split()[1]
            op = op.split('.')[0]
            byop[op] += int(r['# Samples']); cnt[op] += 1; execd[op] += int(r['Instructions Executed'])
        print('total samples', tot, 'static instrs', len(rows), 'executed warp-instrs', sum(execd.values()))
        for op, v in byop.most_common(22):
            print(f"  {op:12s} samples {v:7d} ({100*v/tot:5.1f}%)  static {cnt[op]:6d}  executed {execd[op]:10d} ({100*execd[op]/sum(execd.values()):4.1f}%)")
        # stall reason columns
        reasons = [h for h in hdr if h.startswith('stall_')]
        agg = collections.Counter()
        for r in rows:
            for h in reasons:
                try: agg[h] += int(r[h])
                except: pass
        print('  stall reasons:', ', '.join(f"{h[6:]}={v}" for h, v in agg.most_common(10)))
    else:
        k += 1
