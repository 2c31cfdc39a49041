"""Summarise `nvcc -Xptxas -v` output: registers / spills per kernel (demangled, shortened).
    python tools/ptxas_report.py build.log [other.log]   # two logs -> side by side diff"""
import re, subprocess, sys

def parse(path):
    out, name = {}, None
    for line in open(path):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name:
            out.setdefault(name, {})["spill"] = (int(m.group(2)), int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            out.setdefault(name, {})["regs"] = int(m.group(1))
    return out

def demangle(names):
    res = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    short = []
    for r in res:
        r = r.replace("ds::", "").replace("(ds::LossParams)", "").replace("void ", "")
        r = r.replace("__nv_bfloat16", "bf16").replace("__half", "f16").replace("unsigned char", "u8").replace("long", "i64")
        r = r.replace("(bool)1", "1").replace("(bool)0", "0").replace("(int)", "")
        short.append(r)
    return dict(zip(names, short))

a = parse(sys.argv[1])
b = parse(sys.argv[2]) if len(sys.argv) > 2 else None
names = demangle(list(a))
for k, v in a.items():
    line = f"{names[k]:60s} regs {v.get('regs')} spill {v.get('spill')}"
    if b is not None:
        w = b.get(k, {})
        if w.get("regs") == v.get("regs") and w.get("spill") == v.get("spill"):
            continue
        line += f"   | other: regs {w.get('regs')} spill {w.get('spill')}"
    print(line)
