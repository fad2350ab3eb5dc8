"""SASS evidence of the built library (no GPU needed): which kernels use the bulk-tensor copy engine (UTMALDG /
UTMASTG), L2 bulk prefetch (UBLKPF), packed 2 x fp32 arithmetic (FFMA2 / FADD2 / FMUL2) and the FP64 tensor path (DMMA).

    python tools/sass_evidence.py > profiles/r2_sass_tma_packed.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "fastbox_b200", "libfastbox_b200.so")
OPS = ["UTMALDG", "UTMASTG", "UBLKPF", "UBLKCP", "SYNCS", "DMMA", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "DFMA",
       "LDG", "STG", "LDS", "STS", "ATOMS", "RED", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    per, cur, it = collections.OrderedDict(), None, iter(names)
    for ln in sass.splitlines():
        if "Function : " in ln:
            cur = next(it)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", ln)
        if m and cur is not None:
            per[cur][m.group(1)] += 1
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("SASS evidence, libfastbox_b200.so built for sm_100a (cuobjdump -sass), final round-2 build; tools/sass_evidence.py")
    print("opcode totals over the library: " + ", ".join("%s %d" % (o, tot[o]) for o in OPS))

    def show(title, pred, ops):
        print("\n" + title)
        for k, c in per.items():
            if pred(k, c):
                print("  %-100s %s" % (k[:100], " ".join("%s %d" % (o, c[o]) for o in ops)))
    show("kernels that move their tiles with the bulk-tensor copy engine (UTMALDG / UTMASTG) or prefetch into L2 with bulk "
         "requests (UBLKPF), per kernel instruction counts:",
         lambda k, c: c["UTMALDG"] or c["UTMASTG"] or c["UBLKPF"] or c["UBLKCP"],
         ["UTMALDG", "UTMASTG", "UBLKPF", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "LDG", "STG"])
    show("packed 2 x fp32 arithmetic (FFMA2 / FADD2 / FMUL2) in the 1024^3 pass kernels:",
         lambda k, c: c["FFMA2"] and "1024" in k and ("k_rows" in k or "k_cols" in k or "k_x_" in k or "k_beam_x" in k),
         ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "LDG", "STG", "LDS", "STS"])
    show("FP64 tensor path (DMMA) and float64 FMA in the PCA kernels:", lambda k, c: "k_pca" in k or "k_fp64" in k,
         ["DMMA", "DFMA", "LDS", "STS", "LDG"])


if __name__ == "__main__":
    main()
