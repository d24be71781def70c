#!/usr/bin/env python
"""Opcode mix of the innermost backward-branch loop(s) of a kernel: python tools/sass_loop.py 'ILi2ELi2'"""
import collections, re, subprocess, sys
pat = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else "landhydrology.jl_b200/csrc/liblh_soil.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name = None
ins = []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); continue
    if name and pat in name:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for addr, text in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?(\w+)\)?", text)
    if "BRA" in text:
        t = re.search(r"0x([0-9a-f]+)", text)
        if t and int(t.group(1), 16) < addr:
            loops.append((int(t.group(1), 16), addr))
print("backward branches:", [(hex(a), hex(b), (b - a) // 16) for a, b in loops])
if loops:
    a, b = max(loops, key=lambda ab: ab[1] - ab[0])
    mix = collections.Counter()
    for addr, text in ins:
        if a <= addr <= b:
            t = text.split()
            op = t[1] if t[0].startswith("@") else t[0]
            mix[op.split(".")[0]] += 1
    print("largest loop:", hex(a), hex(b), "instructions:", sum(mix.values()))
    for op, n in mix.most_common(40):
        print(f"  {op:10s} {n}")
