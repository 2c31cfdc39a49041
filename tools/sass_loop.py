#!/usr/bin/env python
"""Histogram the SASS opcodes of the innermost loops of one kernel (CPU-side check before GPU time).

    python tools/sass_loop.py <object-or-so> <demangled-name-substring>
"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(f"cuobjdump -sass {obj} | c++filt", shell=True, capture_output=True, text=True).stdout
    blocks = out.split("Function : ")
    for blk in blocks[1:]:
        name = blk.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = []
        for line in blk.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        print(f"== {name}: {len(ins)} SASS instructions")
        addr_index = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, text) in enumerate(ins):
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_index:
                    loops.append((addr_index[tgt], i))
        for lo, hi in loops:
            body = ins[lo : hi + 1]
            hist = collections.Counter()
            for _, text in body:
                t = text.split()
                op = t[1] if t[0].startswith("@") else t[0]
                hist[op.split(".")[0]] += 1
            print(f"  loop [{ins[lo][0]:#x}..{ins[hi][0]:#x}] {len(body)} instr: {hist.most_common()}")


if __name__ == "__main__":
    main()
