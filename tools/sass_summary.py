"""Static evidence for a kernel from the in-tree objects (no GPU needed): ptxas resource lines from the build log and an
opcode histogram of its SASS (cuobjdump). Usage:
    python tools/sass_summary.py attention 'attention_kernelILb0ELi2' attention_pair 'attention_pair_kernelILb0ELi2' > profiles/attention_sass_r1.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

BUILD = Path(__file__).resolve().parents[1] / "l4p_b200" / "csrc" / "build"
TENSOR = ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCCP", "SYNCS", "UCGABAR", "MUFU", "FFMA2", "FADD2",
          "FMNMX3", "FMNMX", "F2FP", "STS", "LDS", "STG", "LDG", "STL", "LDL")


def summarize(obj: str, pattern: str) -> None:
    log = (BUILD / f"{obj}.log").read_text().splitlines()
    sass = subprocess.run(["cuobjdump", "-sass", str(BUILD / f"{obj}.o")], capture_output=True, text=True).stdout.splitlines()
    name, body = None, []
    for line in sass:
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name is not None:
                break
            if pattern in m.group(1):
                name = m.group(1)
            continue
        if name is not None:
            body.append(line)
    if name is None:
        print(f"{obj}: no function matching {pattern}")
        return
    print(f"== {name}  ({obj}.cu)")
    for i, line in enumerate(log):
        if name in line and "Function properties" in line:
            print("   " + log[i + 1].strip())
            print("   " + log[i + 2].strip().replace("ptxas info    : ", ""))
    ops = collections.Counter()
    full = collections.Counter()
    for line in body:
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            ops[m.group(1)] += 1
            full[m.group(1) + m.group(2)] += 1
    total = sum(ops.values())
    print(f"   {total} SASS instructions; tensor / async / softmax-relevant opcodes:")
    print("   " + ", ".join(f"{k} {ops[k]}" for k in TENSOR if ops[k]))
    mma = {k: v for k, v in full.items() if k.startswith(("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "STTM"))}
    print("   " + ", ".join(f"{k} {v}" for k, v in sorted(mma.items())))
    print("   top opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))


if __name__ == "__main__":
    args = sys.argv[1:]
    for obj, pat in zip(args[0::2], args[1::2]):
        summarize(obj, pat)
