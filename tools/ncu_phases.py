"""Per-phase instruction/sample split of a narrow beam kernel ncu report (phases = '// ---- Px' markers).
   python tools/ncu_phases.py report.ncu-rep frames_per_launch"""
import csv
import re
import subprocess
import sys

rep, frames = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 128000.0
src = open(__file__.rsplit("/tools/", 1)[0] + "/ctc-beam-search-op_b200/csrc/" + (sys.argv[3] if len(sys.argv) > 3 else "ctcx_beam_v4.cuh")).read().splitlines()
bounds = [(1, "init")]
for i, ln in enumerate(src, 1):
    m = re.search(r"// ---- (P[A-G])", ln)
    if m:
        bounds.append((i, m.group(1)))
    elif "auto load_row" in ln:
        bounds.append((i, "S:load_row"))
    elif "auto prepare1" in ln:
        bounds.append((i, "S:part1"))
    elif "auto prepare2" in ln:
        bounds.append((i, "S:part2"))
    elif "---- initial state" in ln:
        bounds.append((i, "init"))
    elif "// ---- S" in ln:
        bounds.append((i, "S:call"))
    elif "---- final beam" in ln:
        bounds.append((i, "final"))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
items, curfile, curline, hdr = [], None, None, None
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        curfile = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if r[0] == "Function Name":
        continue
    if r[0] != "":
        try:
            curline = int(r[0])
        except ValueError:
            curline = None
        continue
    if hdr and len(r) > ie and r[2].startswith("0x"):
        try:
            items.append((int(r[2], 16), curfile, curline, int(r[ie] or 0), int(r[isamp] or 0)))
        except ValueError:
            pass
items.sort()


def phase(line):
    p = "init"
    for b, nm in bounds:
        if line >= b:
            p = nm
    return p


agg, cur, tot, tots = {}, "init", 0, 0
for a, f, l, i, s in items:
    if f == (sys.argv[3] if len(sys.argv) > 3 else "ctcx_beam_v4.cuh") and l:
        cur = phase(l)
    agg.setdefault(cur, [0, 0])
    agg[cur][0] += i
    agg[cur][1] += s
    tot += i
    tots += s
for k, (i, s) in agg.items():
    print("%-12s %5.1f%% inst (%7.0f/frame)  %5.1f%% samples" % (k, 100.0 * i / tot, i / frames, 100.0 * s / tots))
print("total warp-instructions/frame %.0f" % (tot / frames))
