"""Per-kernel SASS opcode histogram of the built library plus the main loop of the dominant kernel.

usage: python tools/sass_report.py > profiles/sass_rNN.txt      (cuobjdump only, no GPU)
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gp_b200", "lib", "libgpb200.so")
KEYS = ["DMMA", "DFMA", "DMUL", "DADD", "MUFU.RSQ64H", "MUFU.RCP64H", "MUFU.EX2", "LDGSTS", "LDS", "STS",
        "LDG", "STG", "BAR", "SHFL"]
DOMINANT = "_ZN3gpb16gemm_tile_kernelINS_7GemmCfgILi4ELi2ELi64ELi3ELi2ELi128EEELb0ELb0ELi0EEEvNS_10GemmParamsE"


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)
    print("# `cuobjdump -sass gp_b200/lib/libgpb200.so` (sm_100a): per-kernel count of the FP64 and data-movement")
    print("# mnemonics, then the main loop of the dominant kernel (the NT half-tile GEMM of the Cholesky update).")
    print("# DMMA.8x8x4 is what mma.sync.m8n8k4.f64 compiles to: the only FP64 tensor-core instruction sm_100a has")
    print("# (tcgen05.mma has no f64 kind). LDGSTS = cp.async (16-byte, L2-only). 64-bit fragments cannot use LDSM.")
    print()
    rows = []
    bodies = {}
    for p in parts[1:]:
        name = p.split("\n", 1)[0].strip()
        ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Za-z0-9_.]+)", p, flags=re.M)
        c = collections.Counter()
        for o in ops:
            for k in KEYS:
                if o.startswith(k):
                    c[k] += 1
                    break
        rows.append((name, len(ops), c))
        bodies[name] = p
    names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True,
                           text=True).stdout.splitlines()
    print("%-104s %6s " % ("kernel", "insts") + " ".join("%6s" % k.replace("MUFU.", "")[:6] for k in KEYS))
    for (name, n, c), d in sorted(zip(rows, names), key=lambda r: -r[0][2]["DMMA"]):
        print("%-104s %6d " % (d[:104], n) + " ".join("%6d" % c[k] for k in KEYS))
    body = bodies.get(DOMINANT)
    if body:
        lines = body.splitlines()
        idx = [i for i, l in enumerate(lines) if "DMMA" in l]
        # the steady-state loop is the longest run of DMMAs closed by a backward branch
        print("\n# ---- main loop of", DOMINANT, "(instructions only) ----")
        lo, hi = idx[0], idx[-1]
        for l in lines[max(0, lo - 40):hi + 12]:
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
            if m:
                print(f"  /*{m.group(1)}*/ {m.group(2).rstrip()}")


if __name__ == "__main__":
    main()
