#!/usr/bin/env python
"""Per-kernel SASS fingerprints of csrc/liblh_soil.so.

    python tools/sass_identity.py --record profiles/<name>.json     # fingerprint the current build
    python tools/sass_identity.py --check  profiles/<name>.json     # which kernels differ from the recorded build?

A host-side refactor (launch spelling, header moves, the hooks of the CPU-only test build) must leave every
kernel's machine code untouched: `--check` against the fingerprints of the last build whose `-m gpu` suite ran green on a B200
proves it without a GPU.  A kernel change shows up as exactly the set of variants it was meant to touch."""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "landhydrology.jl_b200", "csrc", "liblh_soil.so")


def fingerprints(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    fps, name, h, n = {}, None, None, 0
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                fps[name] = {"sha1": h.hexdigest(), "instructions": n}
            name, h, n = m.group(1), hashlib.sha1(), 0
            continue
        if name and "/*" in line:
            h.update(line.strip().encode())
            if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
                n += 1
    if name:
        fps[name] = {"sha1": h.hexdigest(), "instructions": n}
    return fps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--record")
    ap.add_argument("--check")
    ap.add_argument("--lib", default=LIB)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    fps = fingerprints(a.lib)
    if a.record:
        json.dump({"note": a.note, "kernels": fps}, open(a.record, "w"), indent=0, sort_keys=True)
        print(f"{len(fps)} kernels recorded in {a.record}")
        return 0
    ref = json.load(open(a.check))["kernels"]
    changed = sorted(k for k in fps if k in ref and ref[k]["sha1"] != fps[k]["sha1"])
    added = sorted(k for k in fps if k not in ref)
    removed = sorted(k for k in ref if k not in fps)
    print(f"{len(fps)} kernels: {len(fps) - len(changed) - len(added)} identical, {len(changed)} changed, {len(added)} new, {len(removed)} gone")
    for k in changed:
        d = subprocess.run(["cu++filt", k], capture_output=True, text=True).stdout.strip() or k
        print(f"  changed: {d[:110]}  ({ref[k]['instructions']} -> {fps[k]['instructions']} instructions)")
    for k in added:
        print("  new:", k[:120])
    for k in removed:
        print("  gone:", k[:120])
    return 1 if (changed or added or removed) else 0


if __name__ == "__main__":
    sys.exit(main())
