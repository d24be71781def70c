#!/usr/bin/env python
"""Static issue-slot model of every stage-kernel variant in the built library.

For each lh_soil_stage_kernel<MODEL, STAGE, FLAGS> the innermost layer loop (the largest backward-branch loop, two cells
per trip) is disassembled with cuobjdump -sass and its opcodes counted: fp64 (DFMA/DMUL/DADD/DSETP: each holds the issue
port for two cycles on B200) and everything else.  cycles/warp-cell = 2 F + O is the issue-slot floor of one cell-stage;
with 4 schedulers per SM it gives the variant's instruction roofline (bench.py: roofline.issue_model).  Also records, per
kernel, registers-free facts that matter for hygiene: local-memory instructions (LDL/STL = spills) in the whole kernel.

    python tools/sass_loop_mix.py [library.so] > profiles/r02_sass_loop_mix.json
Divergent variants (ICE: per-lane `icy` branches) are counted statically, i.e. both sides of every branch.
"""
import collections, json, re, subprocess, sys, os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "landhydrology.jl_b200", "csrc", "liblh_soil.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs, name = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); funcs[name] = []; continue
    if name:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            funcs[name].append((int(m.group(1), 16), m.group(2).strip()))
F64 = ("DFMA", "DMUL", "DADD", "DSETP")
per = collections.defaultdict(dict)
spills = {}
for name, ins in funcs.items():
    m = re.search(r"(lh_soil_stage_kernel|lh_soil_ssprk33_persistent_kernel)ILi(\d)ELi(\d+)E(?:Li(\d+)E)?", name)
    if not m:
        continue
    kind, M = m.group(1), int(m.group(2))
    if kind == "lh_soil_stage_kernel":
        S, FL = int(m.group(3)), int(m.group(4))
    else:
        S, FL = 6, int(m.group(3))
    nspill = sum(1 for _, t in ins if re.search(r"\b(LDL|STL)\b", t))
    spills[f"{kind}<{M},{S},{FL}>"] = {"LDL_STL": nspill, "instructions": len(ins)}
    loops = []
    for addr, text in ins:
        if "BRA" in text:
            t = re.search(r"0x([0-9a-f]+)", text)
            if t and int(t.group(1), 16) < addr:
                loops.append((int(t.group(1), 16), addr))
    if not loops:
        continue
    a, b = max(loops, key=lambda ab: ab[1] - ab[0])
    mix = collections.Counter()
    for addr, text in ins:
        if a <= addr <= b:
            t = text.split()
            op = t[1] if t[0].startswith("@") else t[0]
            mix[op.split(".")[0]] += 1
    tot = sum(mix.values()); f = sum(mix[k] for k in F64)
    per[f"{M}_{FL}"][S] = {"fp64": f / 2, "other": (tot - f) / 2, "cycles": (tot + f) / 2, "lds": mix["LDS"] / 2, "mufu": mix["MUFU"] / 2,
                           "branches": mix["BRA"] / 2, "top": dict(mix.most_common(12))}
res = {}
for key, st in per.items():
    if all(s in st for s in (1, 2, 3)):
        res[key] = {
            "fp64_mean": sum(st[s]["fp64"] for s in (1, 2, 3)) / 3, "other_mean": sum(st[s]["other"] for s in (1, 2, 3)) / 3,
            "cycles_per_warp_cell_stage_mean": sum(st[s]["cycles"] for s in (1, 2, 3)) / 3,
            "stages": {str(s): st[s] for s in sorted(st)},
        }
res["_what"] = ("key = MODEL_FLAGS (MODEL 0 Richards, 1 heat, 2 coupled; FLAGS bit 1 ICE, 2 GEN, 4 VG2, 8 HET); per cell and stage, "
                "static counts of the layer loop; cycles = 2 x fp64 + other")
res["_spills"] = {k: v for k, v in sorted(spills.items()) if v["LDL_STL"]}
res["_kernels"] = len(spills)
print(json.dumps(res, indent=1))
