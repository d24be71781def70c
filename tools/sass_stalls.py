#!/usr/bin/env python
"""Static issue model of a kernel's hot loop from its SASS control codes (no GPU needed).

    python tools/sass_stalls.py <cubin|so> <mangled-name pattern> [cells per loop trip]

For the largest backward-branch loop: instruction count, fp64-pipe instruction count, and the sum of
the per-instruction stall counts (bits 41..44 of the control word) = the cycles ONE warp needs for one
trip when nothing else delays it.  With w warps per scheduler the loop cannot run faster than
max(stall_sum / w, 2 * fp64 instructions, instructions) cycles per trip.
"""
import collections
import re
import subprocess
import sys


def parse(path, pat):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout.splitlines()
    name, ins, i = None, [], 0
    while i < len(out):
        m = re.search(r"Function : (\S+)", out[i])
        if m:
            name = m.group(1)
        elif name and pat in name:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", out[i])
            if m and i + 1 < len(out):
                m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", out[i + 1])
                if m2:
                    w = int(m2.group(1), 16)
                    ins.append((int(m.group(1), 16), m.group(2).strip(), (w >> 41) & 0xF))
                    i += 1
        i += 1
    return ins


def main():
    path, pat = sys.argv[1], sys.argv[2]
    cells = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ins = parse(path, pat)
    loops = []
    for addr, text, _ in ins:
        if "BRA" in text:
            t = re.search(r"0x([0-9a-f]+)", text)
            if t and int(t.group(1), 16) < addr:
                loops.append((int(t.group(1), 16), addr))
    if not loops:
        print("no loop found")
        return
    a, b = max(loops, key=lambda ab: ab[1] - ab[0])
    body = [x for x in ins if a <= x[0] <= b]
    mix = collections.Counter()
    for _, text, _ in body:
        t = text.split()
        op = t[1] if t[0].startswith("@") else t[0]
        mix[op.split(".")[0]] += 1
    fp64 = sum(mix[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    stall = sum(x[2] for x in body)
    n = len(body)
    print(f"loop {hex(a)}..{hex(b)}: {n} instr ({n / cells:.1f}/cell), fp64 {fp64} ({fp64 / cells:.1f}/cell), "
          f"stall sum {stall} ({stall / cells:.0f}/cell), spills LDL/STL {mix['LDL']}/{mix['STL']}")
    print("  " + "  ".join(f"{k}:{v}" for k, v in mix.most_common(14)))


if __name__ == "__main__":
    main()
