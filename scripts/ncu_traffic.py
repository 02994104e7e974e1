"""dram__bytes_read.sum + dram__bytes_write.sum per launch from an `ncu --set full` report -> profiles/rNN_extend_traffic.json
usage: ncu_traffic.py report.ncu-rep out.json "kernel label" "source command" """
import csv, io, json, subprocess, sys
rep, out, label, src = sys.argv[1:5]
raw = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout)))
hdr, units, rows = raw[0], raw[1], raw[2:]
def col(name):
    i = hdr.index(name); u = units[i]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}[u]
    return [float(r[i].replace(',', '')) * scale for r in rows]
rd, wr, dur = col('dram__bytes_read.sum'), col('dram__bytes_write.sum'), col('gpu__time_duration.sum')
tot = [a + b for a, b in zip(rd, wr)]
json.dump({"kernel": label, "source": src, "dram_bytes_per_launch": tot, "mean_dram_bytes_per_launch": sum(tot) / len(tot), "duration_us": dur},
          open(out, 'w'), indent=1)
print(out, 'mean bytes/launch', sum(tot) / len(tot), 'launches', len(tot))
