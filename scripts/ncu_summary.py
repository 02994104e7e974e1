"""Summarise an ncu report: headline metrics per launch + instruction share / active threads per source region."""
import csv, collections, subprocess, sys, io, re
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
det = subprocess.run(['ncu', '-i', rep, '--page', 'details', '--csv'], capture_output=True, text=True).stdout
keep = ['Duration', 'Registers Per Thread', 'Achieved Occupancy', 'Theoretical Occupancy', 'L1/TEX Hit Rate', 'L2 Hit Rate', 'Compute (SM) Throughput', 'DRAM Throughput',
        'Executed Ipc Active', 'Issue Slots Busy', 'Avg. Active Threads Per Warp', 'Avg. Not Predicated Off Threads Per Warp', 'No Eligible', 'Branch Efficiency', 'Memory Throughput']
for row in csv.DictReader(io.StringIO(det)):
    if int(row['ID']) < nl and row['Metric Name'] in keep:
        print(row['ID'], row['Metric Name'].ljust(42), row['Metric Value'], row['Metric Unit'])
raw = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout)))
hdr = raw[0]
for i, h in enumerate(hdr):
    if h in ('smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
             'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum'):
        print(h, raw[1][i], [row[i] for row in raw[2:2 + nl]])
stalls = {h: [float(row[i] or 0) for row in raw[2:2 + nl]] for i, h in enumerate(hdr) if h.startswith('smsp__pcsamp_warps_issue_stalled') and 'not_issued' not in h}
tot = [sum(v[k] for v in stalls.values()) for k in range(nl)]
print('stalls(%):', {h.replace('smsp__pcsamp_warps_issue_stalled_', ''): [round(100 * v[k] / max(1, tot[k])) for k in range(nl)] for h, v in stalls.items() if max(v) / max(1, max(tot)) > 0.03})
srccsv = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(srccsv)))
sections = []; cur = None
for r in rows:
    if len(r) >= 1 and r[0] == 'File Path': cur = {'file': r[1], 'rows': []}; sections.append(cur)
    elif cur is not None and len(r) > 5 and r[0] == 'Line No': cur['hdr'] = r
    elif cur is not None and len(r) > 5: cur['rows'].append(r)
seen = set(); agg = collections.defaultdict(lambda: [0, 0, 0])
for s in sections:
    if s['file'] in seen: break
    seen.add(s['file'])
    h = s['hdr']; iI = h.index('Instructions Executed'); iT = h.index('Thread Instructions Executed'); iS = h.index('# Samples')
    for r in s['rows']:
        if not r[0].strip().isdigit(): continue
        try: ie = int(r[iI] or 0); te = int(r[iT] or 0); sm = int(r[iS] or 0)
        except ValueError: continue
        k = (s['file'].split('/')[-1], int(r[0])); agg[k][0] += ie; agg[k][1] += te; agg[k][2] += sm
tot = sum(v[0] for v in agg.values()) or 1
print('total warp inst (launch 0)', tot)
files = {}
def srcline(f, l):
    import os
    p = '/root/repo/rustracer_b200/csrc/' + f
    if f not in files: files[f] = open(p).read().split('\n') if os.path.exists(p) else []
    return files[f][l - 1].strip()[:100] if 0 < l <= len(files[f]) else ''
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{f}:{l:4d} inst {100 * v[0] / tot:5.2f}% thr {v[1] / max(1, v[0]):5.1f} smp {v[2]:5d} | {srcline(f, l)}")
