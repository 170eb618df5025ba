#!/usr/bin/env python3
"""Per-kernel SASS evidence for profiles/: `cuobjdump -sass` of libplvi_cuda.so, one mnemonic histogram per kernel
(instruction counts in the binary, not executed counts) plus the first lines of the hot loop markers the notes cite
(UBLKCP = cp.async.bulk, IDP.4A = dp4a, POPC, DADD / DMUL / DFMA, LDG.E.128, ATOMS, SHFL, VOTE, REDUX).

  python tools/sass_summary.py profiles/r02_sass.md
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "pl_vi_orbslam3_b200" / "libplvi_cuda.so"
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "LDG.E.128", "LDG.E.64", "LDG.E.U8", "LDG.E", "LDS", "STS", "STG", "ATOMS", "ATOMG", "RED", "IDP.4A", "IDP.2A",
         "POPC", "VABSDIFF", "VIMNMX", "LOP3", "SHF", "PRMT", "IMAD", "IADD3", "ISETP", "FFMA", "FMUL", "FADD", "MUFU", "DFMA", "DMUL", "DADD",
         "DSETP", "I2F", "F2I", "F2F", "SHFL", "VOTE", "MATCH", "REDUX", "BAR", "BRA", "HMMA", "UTCHMMA"]


def main():
    out = Path(sys.argv[1])
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    kernels, cur, arch = collections.OrderedDict(), None, set()
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            name = name.replace("void ", "").replace("plvi::", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
                    break
    with open(out, "w") as f:
        f.write("# SASS summary of libplvi_cuda.so\n\n`cuobjdump -sass pl_vi_orbslam3_b200/libplvi_cuda.so` (" + ", ".join(sorted(arch)) +
                "), static instruction counts per kernel.\nTensor pipes are unused by design (no `HMMA` / `UTC*MMA`): no stage is a contraction.\n\n")
        cols = [w for w in WATCH if any(k[w] for k in kernels.values())]
        f.write("| kernel | total | " + " | ".join(cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
        for name, c in kernels.items():
            f.write(f"| `{name[:60]}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |\n")
    print(out, len(kernels), "kernels")


if __name__ == "__main__":
    main()
