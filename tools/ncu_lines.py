"""
Attribute ncu stall samples to CUDA source lines.

  python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top]

ncu's CSV source page is per SASS instruction; nvdisasm --print-line-info on the cubin extracted
from the shared library gives the source line of every instruction of the same function.  The
two listings are matched by instruction order.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fastbox_b200", "libfastbox_b200.so")


def sass_lines_from_cubins(func_regex):
    out = {}
    bdir = os.path.join(ROOT, "fastbox_b200", "csrc", "build")
    cubins = []
    for o in sorted(os.listdir(bdir)):
        if not o.endswith(".o"):
            continue
        tmp = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(bdir, o)], cwd=tmp, capture_output=True)
        cubins += [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    for f in cubins:
        txt = subprocess.run(["nvdisasm", "--print-line-info", f], capture_output=True, text=True).stdout
        cur, line, items = None, None, []
        for ln in txt.split("\n"):
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                if cur:
                    out[cur] = items
                cur, items, line = m.group(1), [], None
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur:
                items.append((line, m.group(2).strip()))
        if cur:
            out[cur] = items
    return {k: v for k, v in out.items() if re.search(func_regex, k)}


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    parts = re.split(r'(?m)^"Kernel Name",', txt)
    cub = sass_lines_from_cubins(".")
    for part in parts[1:]:
        lines = part.split("\n")
        name = lines[0].strip().strip('",')
        if not re.search(rx, name):
            continue
        rdr = csv.reader(io.StringIO("\n".join(lines[1:])))
        hdr = next(rdr)
        ci = {h: i for i, h in enumerate(hdr)}
        rows = [r for r in rdr if len(r) == len(hdr)]
        # find the cubin function with the same instruction count
        cand = [(k, v) for k, v in cub.items() if len(v) == len(rows)]
        # disambiguate by demangled-ish name pieces (template ints)
        ints = re.findall(r"\(int\)(-?\d+)", name)
        best = None
        for k, v in cand:
            if all(("li%se" % i.replace("-", "n")) in k.lower() for i in ints):
                best = (k, v)
        if best is None and cand:
            best = cand[0]
        print("=====", name, "sass", len(rows), "->", best[0] if best else None)
        if not best:
            continue
        per = defaultdict(lambda: [0, 0])
        tot = 0
        for r, (line, ins) in zip(rows, best[1]):
            s = int(r[ci["# Samples"]])
            e = int(r[ci["Instructions Executed"]])
            per[line][0] += s
            per[line][1] += e
            tot += s
        tote = sum(v[1] for v in per.values())
        srccache = {}
        for line, (s, e) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
            text = ""
            if line:
                fn = os.path.join(ROOT, "fastbox_b200", "csrc", line[0])
                if os.path.isfile(fn):
                    if fn not in srccache:
                        srccache[fn] = open(fn).read().split("\n")
                    if line[1] - 1 < len(srccache[fn]):
                        text = srccache[fn][line[1] - 1].strip()[:90]
            print("%5.1f%% samples %5.1f%% inst  %s:%s  %s" % (100. * s / max(tot, 1), 100. * e / max(tote, 1),
                                                           line[0] if line else "?", line[1] if line else "?", text))


if __name__ == "__main__":
    main()
