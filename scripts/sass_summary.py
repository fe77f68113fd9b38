"""Per-kernel counts of the SASS mnemonics that show what the shipped library runs on
(`cuobjdump -sass` of libirr_b200.so): tcgen05 MMAs (UTCHMMA, .2CTA = cta_group::2), TMA loads
(UTMALDG, UBLKCP = 1-D bulk), tcgen05.commit (UTCBAR, .MULTICAST), TMEM loads (LDTM), packed fp32
math (FFMA2 / FMUL2), the NaN-propagating 3-input maximum of the epilogue (FMNMX3.NAN) ...

    python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "imageretrievalresearch_b200", "libirr_b200.so")
COLS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG.2D.2CTA", "UTMALDG.2D", "UTCBAR.2CTA.MULTICAST", "UTCBAR",
        "LDTM", "UBLKCP", "FFMA2", "FMUL2", "FMNMX3.NAN", "REDUX", "HMMA", "FFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    names = [f.split("\n", 1)[0].strip() for f in funcs]
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    rows, total = [], collections.Counter()
    for f, d in zip(funcs, dem):
        ops = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", f)
        c = collections.Counter()
        for op in ops:
            for col in COLS:
                if op == col or op.startswith(col + "."):
                    # count an instruction under its most specific column only
                    best = max((x for x in COLS if op == x or op.startswith(x + ".")), key=len)
                    c[best] += 1
                    break
        short = re.sub(r"irr::\(anonymous namespace\)::", "", d)
        short = re.sub(r"\(.*$", "", short).replace("void ", "")
        rows.append((short, len(ops), c))
        total.update(c)
    print(f"SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a): instructions per kernel")
    print("columns: " + ", ".join(COLS))
    print()
    w = max(len(r[0]) for r in rows)
    print(f"{'kernel':{w}s} {'insts':>6s} " + " ".join(f"{c[-9:]:>9s}" for c in COLS))
    for short, n, c in sorted(rows):
        print(f"{short:{w}s} {n:6d} " + " ".join(f"{c[col]:9d}" for col in COLS))
    print()
    print(f"{'total':{w}s} {sum(r[1] for r in rows):6d} " + " ".join(f"{total[col]:9d}" for col in COLS))
    print("\nno CUTLASS / CuTe / cuBLAS symbols: " +
          str(not re.search(r"cutlass|cute::|cublas", "\n".join(dem))))


if __name__ == "__main__":
    sys.exit(main())
