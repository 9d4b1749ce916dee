"""Executed-instruction histogram by opcode (per unit of work) and the hottest SASS lines from an ncu source-page CSV.
python tools/ncu_ops.py src.csv units"""
import csv, collections, sys
src = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = src[1]; ix = {c: i for i, c in enumerate(h)}
ci, cs = ix['Instructions Executed'], ix['Warp Stall Sampling (All Samples)']
tot = 0; by = collections.Counter(); samp = []; stot = 0
for r in src[2:]:
    if len(r) <= ci: continue
    try: v = float(r[ci].replace(',', ''))
    except ValueError: continue
    toks = r[ix['Source']].split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
    by[op.split('.')[0]] += v; tot += v
    sv = float(r[cs].replace(',', '') or 0); stot += sv
    samp.append((sv, v, r[ix['Source']].strip()))
print('total warp-inst %.0f  per unit %.1f' % (tot, tot / units))
for op, v in by.most_common(30):
    print('  %-10s %8.2f' % (op, v / units))
print('hottest SASS lines by stall samples:')
for sv, v, s in sorted(samp, reverse=True)[:30]:
    print('  %5.2f%%  exec/unit %6.2f  %s' % (100 * sv / stot, v / units, s[:100]))
