#!/usr/bin/env python3
"""Instruction census of scl_list_kernel from the built library (cuobjdump -sass): per out-of-line routine (split at
RET) the count of FP64-pipe instructions (DFMA/DADD/DMUL/DSETP/DMNMX) against everything else, by opcode.
Backs the "18 FP64 instructions per phi" figure used in bench.py's issue bound.  Usage: tools/sass_census.py [lib.so]"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "echoseal_b200", "libechoseal_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(txt) if "Function :" in l]
    for si in starts:
        name = txt[si].split("Function :")[1].strip()
        if "scl_list_kernel" not in name:
            continue
        end = min([j for j in starts if j > si] + [len(txt)])
        ins = []
        for l in txt[si:end]:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                ins.append(re.sub(r"^@!?U?P\d+\s+", "", m.group(2)))
        print(f"== {name}: {len(ins)} SASS instructions")
        seg, k = [], 0
        for t in ins:
            seg.append(t)
            if t.startswith("RET") or t.startswith("EXIT") and len(seg) > 200:
                c = collections.Counter(x.split()[0].split(".")[0] for x in seg)
                f = sum(v for op, v in c.items() if op in FP64)
                calls = sum(1 for x in seg if x.startswith("CALL"))
                role = ""
                if 70 <= f <= 80 and len(seg) < 200: role = "  <- psi4: four evaluations of psi = |x|/2 + phi(|x|) in lock-step (4 x 19 FP64)"
                elif 34 <= f <= 38 and len(seg) < 110: role = "  <- phi2: two phi evaluations"
                elif 17 <= f <= 19 and len(seg) < 60: role = "  <- phi1: one phi evaluation"
                print(f"  routine {k}: {len(seg):4d} instr, FP64 pipe {f:3d} ({100 * f / len(seg):4.1f}%), calls {calls}{role}")
                print("      " + " ".join(f"{op}:{v}" for op, v in c.most_common(14)))
                seg, k = [], k + 1
        print()


if __name__ == "__main__":
    main()
