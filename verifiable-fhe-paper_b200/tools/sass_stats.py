#!/usr/bin/env python3
"""Per-kernel SASS statistics for a built .so/.cubin: instruction histogram and the instruction
count of every loop body (backward branch), to estimate issue slots per permutation/butterfly
before spending GPU time.  Usage: sass_stats.py <file> [kernel-substring]"""
import collections
import re
import subprocess
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)[1:]
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if want not in name:
            continue
        ins = []  # (addr, opcode, full)
        for line in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                text = m.group(2).strip()
                op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0]
                ins.append((int(m.group(1), 16), op, text))
        hist = collections.Counter(op.split(".")[0] for _, op, _ in ins)
        wide = sum(1 for _, op, _ in ins if op.startswith("IMAD.WIDE"))
        print(f"== {name}: {len(ins)} instructions, IMAD.WIDE {wide}")
        print("   " + ", ".join(f"{k}:{v}" for k, v in hist.most_common(14)))
        addrs = [a for a, _, _ in ins]
        for a, op, text in ins:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", text)
                if m:
                    tgt = int(m.group(1), 16)
                    if tgt < a:
                        body = [i for i in ins if tgt <= i[0] <= a]
                        h = collections.Counter(o.split(".")[0] for _, o, _ in body)
                        w = sum(1 for _, o, _ in body if o.startswith("IMAD.WIDE"))
                        print(f"   loop {tgt:#x}..{a:#x}: {len(body)} instr (IMAD.WIDE {w}; "
                              + ", ".join(f"{k}:{v}" for k, v in h.most_common(6)) + ")")


if __name__ == "__main__":
    main()
