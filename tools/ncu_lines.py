"""Aggregate an ncu report's source page per CUDA source line:
   python tools/ncu_lines.py report.ncu-rep [top_n]   (needs -lineinfo + --import-source on)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
data, tot_i, tot_s = [], 0, 0
cur_file, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or r[0] == "" or r[0] == "Function Name":
        continue
    try:
        ln, ins, smp = int(r[0]), int(r[ie] or 0), int(r[isamp] or 0)
    except ValueError:
        continue
    data.append((cur_file, ln, r[1], ins, smp))
    tot_i += ins
    tot_s += smp
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
data.sort(key=lambda d: -d[3])
for f, ln, src, ins, smp in data[:top]:
    print("%s:%4d %5.1f%% inst %5.1f%% samples | %s" % (f, ln, 100.0 * ins / max(tot_i, 1),
                                                        100.0 * smp / max(tot_s, 1), src.strip()[:110]))
