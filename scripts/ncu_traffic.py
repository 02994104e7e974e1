"""dram bytes per launch + issue-slot figures from an `ncu --set full` report -> profiles/rNN_extend_traffic.json (read by bench.py
for roofline.traffic / roofline.issue_slots).  usage: ncu_traffic.py report.ncu-rep out.json "kernel label" "source command" """
import csv, io, json, subprocess, sys
rep, out, label, src = sys.argv[1:5]
raw = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout)))
hdr, units, rows = raw[0], raw[1], raw[2:]
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}
def col(name):
    i = hdr.index(name)
    return [float(r[i].replace(',', '')) * SCALE.get(units[i], 1.0) for r in rows]
rd, wr, dur = col('dram__bytes_read.sum'), col('dram__bytes_write.sum'), col('gpu__time_duration.sum')
tot = [a + b for a, b in zip(rd, wr)]
inst = col('smsp__inst_executed.sum'); thr = col('smsp__thread_inst_executed.sum') if 'smsp__thread_inst_executed.sum' in hdr else None
def first(*names):
    for n in names:
        if n in hdr: return col(n)
    return None
issue = first('sm__inst_issued.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct'); ipc = first('sm__inst_executed.avg.per_cycle_active', 'smsp__inst_executed.avg.per_cycle_active')
lanes = first('smsp__thread_inst_executed_per_inst_executed.ratio'); alu = first('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'); fma = first('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active')
w = [d / sum(dur) for d in dur]      # duration-weighted means over the launches of one frame
def wmean(v): return None if v is None else sum(a * b for a, b in zip(v, w))
json.dump({"kernel": label, "source": src, "dram_bytes_per_launch": tot, "mean_dram_bytes_per_launch": sum(tot) / len(tot), "duration_us": dur,
           "issue_slots": {"what": "ncu, duration-weighted over the launches of one frame (bounces 0..7); the kernel's real roof is instruction issue, not HBM",
                           "issue_slots_busy_pct": wmean(issue), "ipc_per_sm": wmean(ipc), "lanes_active_per_warp_inst": wmean(lanes),
                           "alu_pipe_pct": wmean(alu), "fma_pipe_pct": wmean(fma), "warp_inst_per_launch": inst,
                           "per_launch": {"issue_slots_busy_pct": issue, "lanes_active": lanes, "alu_pipe_pct": alu}}},
          open(out, 'w'), indent=1)
print(out, 'mean dram bytes/launch', sum(tot) / len(tot), 'launches', len(tot), 'issue', wmean(issue), 'lanes', wmean(lanes))
