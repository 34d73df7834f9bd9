"""SASS instruction census of one kernel per source region. Usage:
  cuobjdump -xelf all lib/obj/k_narrow.o && nvdisasm -g k_narrow.sm_100a.cubin > all.sass
  python tools/sass_lines.py all.sass <mangled kernel name substring> <source file> [marker ...]
Markers are substrings of source lines; instructions are attributed to the region that starts at the
last marker line at or before their source line; instructions of inlined helpers from other files
(math, intrinsics) are charged to the region of the last instruction of the kernel's own file. Also counts BAR / ATOMS / LDS /
STS / SHFL / fp64 instructions per region."""
import collections
import re
import sys


def main():
    sass, kern, srcfile = sys.argv[1], sys.argv[2], sys.argv[3]
    markers = sys.argv[4:]
    src = open(srcfile).read().split("\n")
    base = srcfile.split("/")[-1]
    starts = []
    for m in markers:
        for i, l in enumerate(src):
            if m in l:
                starts.append((i + 1, m))
                break
    starts.sort()
    lines = open(sass).read().split("\n")
    inside = False
    cur = None
    per = collections.defaultdict(collections.Counter)
    other = collections.Counter()
    last_region = "(before first marker)"
    for ln in lines:
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m or cur is None:
            continue
        opc = m.group(1)
        # attribute to the v4 source line: for inlined code use the "inlined at" chain when present
        f, l, rest = cur
        if f != base:
            mm = re.findall(r'inlined at "([^"]+)", line (\d+)', rest)
            hit = [(a.split("/")[-1], int(b)) for a, b in mm if a.split("/")[-1] == base]
            if hit:
                f, l = hit[-1]
        if f != base:  # inlined helper (math, intrinsics): charge the region of the last line of our file
            other[f] += 1
            region = last_region
        else:
            region = "(before first marker)"
            for s, name in starts:
                if l >= s:
                    region = name
            last_region = region
        c = per[region]
        c["total"] += 1
        for key, pat in (("BAR", r"^BAR"), ("ATOMS", r"^ATOMS"), ("LDS", r"^LDS"), ("STS", r"^STS"),
                         ("SHFL", r"^SHFL"), ("FP64", r"^(DADD|DMUL|DFMA|DSETP|F2F\.F64|F2F\.F32\.F64)"),
                         ("LDG/STG", r"^(LDG|STG)")):
            if re.match(pat, opc):
                c[key] += 1
    tot = sum(c["total"] for c in per.values())
    print("%-46s %6s %5s %5s %5s %5s %5s %5s %5s" % ("region", "instr", "BAR", "ATOMS", "LDS", "STS", "SHFL", "FP64", "LD/ST"))
    for s, name in [(0, "(before first marker)")] + starts:
        c = per.get(name)
        if c:
            print("%-46s %6d %5d %5d %5d %5d %5d %5d %5d" % (name[:46], c["total"], c["BAR"], c["ATOMS"], c["LDS"], c["STS"],
                                                      c["SHFL"], c["FP64"], c["LDG/STG"]))
    print("total in %s: %d; other files: %s" % (base, tot, dict(other)))


if __name__ == "__main__":
    main()
