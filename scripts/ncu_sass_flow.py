"""Per-SASS-instruction listing of one launch of an ncu source page, in address order: share of warp instructions, lanes active,
stall samples, innermost source line.  usage: ncu -i rep --page source --csv --print-source sass,cuda > src.csv; ncu_sass_flow.py src.csv [launch]"""
import csv, sys, io, collections
path = sys.argv[1]; launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
sections = []; cur = None
for r in rows:
    if len(r) >= 1 and r[0] == 'File Path': cur = {'file': r[1].split('/')[-1], 'rows': []}; sections.append(cur)
    elif cur is not None and len(r) > 5 and r[0] == 'Line No': cur['hdr'] = r
    elif cur is not None and len(r) > 5: cur['rows'].append(r)
launches = [[]]; seen = set()
for s in sections:
    if s['file'] in seen: launches.append([]); seen = set()
    seen.add(s['file']); launches[-1].append(s)
secs = launches[launch]
by_addr = {}
for s in secs:
    h = s['hdr']; iI = h.index('Instructions Executed'); iT = h.index('Thread Instructions Executed'); iS = h.index('# Samples'); iP = h.index('Predicated-On Thread Instructions Executed')
    line = None
    for r in s['rows']:
        if r[0].strip().isdigit(): line = int(r[0]); continue
        if not r[2].startswith('0x'): continue
        a = int(r[2], 16)
        e = by_addr.setdefault(a, {'sass': r[3].strip(), 'ie': int(r[iI] or 0), 'te': int(r[iT] or 0), 'pe': int(r[iP] or 0), 'sm': int(r[iS] or 0), 'src': []})
        e['src'].append(f"{s['file']}:{line}")
tot = sum(e['ie'] for e in by_addr.values())
print(f"# launch {launch}: {tot} warp inst, {len(by_addr)} sass")
base = min(by_addr)
for a in sorted(by_addr):
    e = by_addr[a]
    print(f"{(a-base)//16:5d} {100*e['ie']/tot:5.2f}% ie={e['ie']:9d} thr={e['te']/max(1,e['ie']):5.1f} pred={e['pe']/max(1,e['ie']):5.1f} smp={e['sm']:5d} | {e['sass'][:60]:60s} | {' < '.join(e['src'])}")
