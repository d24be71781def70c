#!/usr/bin/env python
"""Static SASS opcode mix of a kernel in the built library: python tools/sass_mix.py 'ILi2ELi2' [so]"""
import collections, re, subprocess, sys
pat = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else "landhydrology.jl_b200/csrc/liblh_soil.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name = None
mix = collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); continue
    if name and pat in name:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            mix[m.group(2)] += 1
tot = sum(mix.values())
print("total", tot)
for op, n in mix.most_common(30):
    print(f"{op:12s} {n}")
