"""Per-instruction warp-stall samples of an `ncu --set full --import-source on` capture (read here, on the CPU):
   ncu -i X.ncu-rep --page source --csv --print-source sass > X.csv ; python tools/ncu_stalls.py X.csv [kernel index] [top N]"""
import csv, io, sys
lines = open(sys.argv[1]).read().split('\n')
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
blocks, cur = [], None
for l in lines:
    if l.startswith('"Kernel Name"'):
        cur = [l]; blocks.append(cur); continue
    if cur is not None and l.strip(): cur.append(l)
b = blocks[which]
print(b[0][:160])
rd = list(csv.DictReader(io.StringIO('\n'.join(b[1:]))))
tot = sum(int(r['# Samples']) for r in rd)
stalls = [k for k in rd[0].keys() if k.startswith('stall_') and 'Not Issued' not in k]
agg = {k: sum(int(r[k]) for r in rd) for k in stalls}
print('instructions', len(rd), 'samples', tot)
print('stall totals:', [(k, v, round(100 * v / tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]])
for r in sorted(rd, key=lambda r: -int(r['# Samples']))[:topn]:
    st = sorted(((k, int(r[k])) for k in stalls), key=lambda x: -x[1])[:2]
    print(rd.index(r), r['# Samples'], round(100 * int(r['# Samples']) / tot, 1), r['Instructions Executed'], r['Source'][:80], st)
