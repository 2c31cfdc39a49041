#!/usr/bin/env python
"""Print the SASS of one kernel between two addresses (control-flow view for checking what sits on the hot path).

    python tools/sass_dump.py <so> <demangled-name-substring> [lo hi] [--cf]   # --cf: control flow / memory ops only
"""
import re
import subprocess
import sys


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    cf = "--cf" in sys.argv
    obj, pat = args[0], args[1]
    lo = int(args[2], 16) if len(args) > 2 else 0
    hi = int(args[3], 16) if len(args) > 3 else 1 << 62
    out = subprocess.run(f"cuobjdump -sass {obj} | c++filt", shell=True, capture_output=True, text=True).stdout
    for blk in out.split("Function : ")[1:]:
        name = blk.split("\n", 1)[0]
        if pat not in name:
            continue
        print("==", name)
        for line in blk.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if not m:
                continue
            a = int(m.group(1), 16)
            if not (lo <= a <= hi):
                continue
            text = m.group(2).strip()
            if cf and not re.search(r"BRA|CALL|BSSY|BSYNC|FMNMX3|VOTE|ATOM|STL|LDL|SYNCS|UBLKCP|WARPSYNC|STG|LDS|RET|EXIT|BAR|RED", text):
                continue
            print(f"{a:#06x}  {text}")
        return


if __name__ == "__main__":
    main()
