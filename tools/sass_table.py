"""Per-kernel table of the Blackwell-specific SASS instructions in libcavit_sm100a.so (cuobjdump -sass, no GPU needed):
tcgen05 MMAs (UTCHMMA, .2CTA = cta_group::2), TMA loads / prefetches / stores (UTMALDG, UTMAPF, UTMASTG), TMEM loads /
stores (LDTM, STTM), tcgen05 commit barriers (UTCBAR), packed fp32x2 math (FFMA2 / FMUL2 / FADD2), MUFU.

    python tools/sass_table.py [--md profiles/sass_table_r02.md]
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cross-attention-vit_b200", "cavit", "libcavit_sm100a.so")
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMAPF", "UTMASTG", "LDTM", "STTM", "UTCBAR", "FFMA2", "MUFU", "total"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.strip().splitlines()
    return [re.sub(r"\(.*", "", o).replace("void ", "").replace("cavit::", "") for o in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
        if not m:
            continue
        op = m.group(1)
        cur["total"] += 1
        base = op.split(".")[0]
        if base == "UTCHMMA":
            cur["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        elif base in ("UTMALDG", "UTMAPF", "UTMASTG", "LDTM", "STTM", "UTCBAR", "MUFU"):
            cur[base] += 1
        elif base in ("FFMA2", "FMUL2", "FADD2"):
            cur["FFMA2"] += 1
    names = list(funcs)
    pretty = demangle(names)
    rows = []
    for n, pn in zip(names, pretty):
        c = funcs[n]
        if any(c[k] for k in COLS[:7]):
            rows.append((pn, [c[k] for k in COLS]))
    rows.sort(key=lambda r: r[0])
    tot = Counter()
    for c in funcs.values():
        tot.update(c)
    lines = ["| kernel | " + " | ".join(COLS) + " |", "|---|" + "---:|" * len(COLS)]
    for pn, vals in rows:
        lines.append(f"| `{pn}` | " + " | ".join(str(v) for v in vals) + " |")
    lines.append("| **whole library** (%d kernels) | " % len(funcs) + " | ".join(str(tot[k]) for k in COLS) + " |")
    text = "\n".join(lines)
    print(text)
    if "--md" in sys.argv:
        path = sys.argv[sys.argv.index("--md") + 1]
        with open(path, "w") as f:
            f.write("# SASS instruction table of `libcavit_sm100a.so` (sm_100a only)\n\n"
                    "`python tools/sass_table.py --md " + os.path.relpath(path, ROOT) + "` — `cuobjdump -sass`, counted per kernel.\n"
                    "FFMA2 column = FFMA2 + FMUL2 + FADD2 (packed fp32x2). Kernels without any tcgen05 / TMA / TMEM instruction are "
                    "summed into the last row only.\n\n" + text + "\n")


if __name__ == "__main__":
    main()
