"""Compact CSV of the metrics the roofline discussion needs from an .ncu-rep (runs where ncu is installed):
    python tools/ncu_extract.py prof.ncu-rep > profiles/rNN_prof_<kernel>_summary.csv"""
import csv, io, re, subprocess, sys

KEEP = re.compile(r"^(ID|Kernel Name|Block Size|Grid Size)$|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|"
                  r"gpu__dram_throughput\.avg\.pct|dram__throughput\.avg\.pct|pipe_tensor.*pct_of_peak_sustained_active|"
                  r"sm__throughput\.avg\.pct|sm__warps_active\.avg\.pct|launch__registers_per_thread$|"
                  r"launch__occupancy_limit|sm__inst_executed_pipe_(xu|fma|alu)\.avg\.pct|smsp__issue_active\.avg\.pct|"
                  r"lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct|l1tex__throughput\.avg\.pct|"
                  r"lts__t_bytes\.sum$|smsp__average_warps_issue_stalled_(long_scoreboard|wait|barrier)_per_issue_active|"
                  r"sm__cycles_active\.avg$|launch__grid_size$")
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
cols = [i for i, h in enumerate(hdr) if KEEP.search(h)]
w = csv.writer(sys.stdout)
for r in rows:
    w.writerow([re.sub(r"\(CUtensorMap.*$|\(const .*$", "", r[i]) if hdr[i] == "Kernel Name" else r[i] for i in cols])
