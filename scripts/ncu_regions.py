"""Region breakdown of an ncu source page: share of warp instructions, lanes active and stall samples per code region.
usage: ncu_regions.py report.ncu-rep [launch_index]
SASS rows from inlined helpers (rt_platform.h, rt_math.h, CUDA headers) take the region of the nearest classified
instruction in address order."""
import csv, collections, subprocess, sys, io, bisect
rep = sys.argv[1]; launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
sections = []; cur = None
for r in rows:
    if len(r) >= 1 and r[0] == 'File Path': cur = {'file': r[1].split('/')[-1], 'rows': []}; sections.append(cur)
    elif cur is not None and len(r) > 5 and r[0] == 'Line No': cur['hdr'] = r
    elif cur is not None and len(r) > 5: cur['rows'].append(r)
# split sections into launches: a file name repeating starts a new launch
launches = [[]]; seen = set()
for s in sections:
    if s['file'] in seen: launches.append([]); seen = set()
    seen.add(s['file']); launches[-1].append(s)
secs = launches[launch]
REGIONS = [  # (file, first line, last line, region) — line numbers of the committed sources
    ('rt_traverse.h', 13, 29, 'ray setup (shear/init)'), ('rt_traverse.h', 30, 58, 'triangle test'), ('rt_traverse.h', 59, 69, 'instance xform'),
    ('rt_traverse.h', 70, 124, 'node test'), ('rt_traverse.h', 138, 174, 'ray setup (shear/init)'), ('rt_traverse.h', 175, 201, 'node step (fetch, stack push)'),
    ('rt_traverse.h', 202, 232, 'instance enter'), ('rt_traverse.h', 233, 250, 'commit'),
    ('rt_kernels.h', 235, 249, 'publish ray'), ('rt_kernels.h', 250, 261, 'coop round: setup/load'), ('rt_kernels.h', 262, 271, 'triangle test'),
    ('rt_kernels.h', 272, 303, 'coop round: resolve'), ('rt_kernels.h', 307, 336, 'fetch rays'), ('rt_kernels.h', 337, 353, 'acquire/pop/retire'),
    ('rt_kernels.h', 354, 375, 'leaf handling'), ('rt_kernels.h', 376, 409, 'queue append'), ('rt_kernels.h', 410, 418, 'flush control'),
    ('rt_kernels.h', 419, 441, 'retire/drain/tail'), ('rt_kernels.h', 442, 470, 'load/store ray+hit'),
]
def region_of(f, l):
    for rf, a, b, name in REGIONS:
        if f == rf and a <= l <= b: return name
    return None
insts = []  # (addr, region|None, inst, thr, samples, file, line)
for s in secs:
    h = s['hdr']; iI = h.index('Instructions Executed'); iT = h.index('Thread Instructions Executed'); iS = h.index('# Samples')
    line = None
    for r in s['rows']:
        if r[0].strip().isdigit(): line = int(r[0]); continue
        if not r[2].startswith('0x'): continue
        try: ie = int(r[iI] or 0); te = int(r[iT] or 0); sm = int(r[iS] or 0)
        except ValueError: continue
        insts.append([int(r[2], 16), region_of(s['file'], line), ie, te, sm, s['file'], line, r[3].strip()])
# an inlined instruction is listed once per level of its inline stack: keep one row per address, classified by the
# innermost (most specific) region
PRIORITY = ['triangle test', 'node test', 'instance xform', 'ray setup (shear/init)', 'commit', 'node step (fetch, stack push)', 'instance enter',
            'publish ray', 'coop round: setup/load', 'coop round: resolve', 'load/store ray+hit', 'fetch rays', 'acquire/pop/retire', 'leaf handling',
            'queue append', 'flush control', 'retire/drain/tail']
by_addr = {}
for x in insts:
    o = by_addr.get(x[0])
    if o is None: by_addr[x[0]] = x; continue
    if x[1] and (not o[1] or PRIORITY.index(x[1]) < PRIORITY.index(o[1])): o[1] = x[1]
insts = sorted(by_addr.values(), key=lambda x: x[0])
known = [i for i, x in enumerate(insts) if x[1]]
for i, x in enumerate(insts):
    if x[1]: continue
    k = bisect.bisect_left(known, i)
    cands = [known[j] for j in (k - 1, k) if 0 <= j < len(known)]
    x[1] = insts[min(cands, key=lambda j: abs(insts[j][0] - x[0]))][1] if cands else 'other'
agg = collections.defaultdict(lambda: [0, 0, 0]); pipes = collections.defaultdict(lambda: collections.Counter())
for a, reg, ie, te, sm, f, l, sass in insts:
    agg[reg][0] += ie; agg[reg][1] += te; agg[reg][2] += sm
    op = sass.split()[1] if sass.startswith('@') else sass.split()[0] if sass else '?'
    pipes[reg][op.split('.')[0]] += ie
tot = sum(v[0] for v in agg.values()) or 1; tots = sum(v[2] for v in agg.values()) or 1
print(f'launch {launch}: {tot} warp instructions, {tots} samples')
for reg, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    top = ', '.join(f'{o} {100 * c / max(1, v[0]):.0f}%' for o, c in pipes[reg].most_common(6))
    print(f'{reg:32s} inst {100 * v[0] / tot:5.1f}%  lanes {v[1] / max(1, v[0]):5.1f}  samples {100 * v[2] / tots:5.1f}%   [{top}]')
