"""stall samples of an ncu report aggregated by SASS opcode:  python scripts/ncu_stalls.py report.ncu-rep"""
import csv, collections, subprocess, sys, io
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
S, E, src = ix['# Samples'], ix['Instructions Executed'], ix['Source']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = 0; byop = collections.Counter(); byst = collections.Counter(); opst = collections.defaultdict(collections.Counter); execd = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    s = int(r[S] or 0); tot += s
    parts = r[src].split(); op = parts[1] if parts[0].startswith('@') else parts[0]; op = op.split('.')[0]
    byop[op] += s; execd[op] += int(r[E] or 0)
    for st in stalls:
        v = int(r[ix[st]] or 0); byst[st] += v; opst[op][st] += v
print('total samples', tot, ' warp-instructions', sum(execd.values()))
for op, v in byop.most_common(16):
    top = ', '.join(f"{k[6:]}:{c}" for k, c in opst[op].most_common(4))
    print(f"{op:8s} {v:7d} {100*v/tot:5.1f}%  exec={execd[op]:9d}  {top}")
print({k[6:]: v for k, v in byst.most_common(12)})
