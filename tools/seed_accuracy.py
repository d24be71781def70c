#!/usr/bin/env python
"""Relative error of the MUFU reciprocal / rsqrt seeds on the GPU (run under gpurun)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import workloads as w
lh = w.lh
ctx = lh.SoilContext(lh.cuda_library(), w.coupled_workload(ncol=32, nlayer=4, seed=1).config())
rng = np.random.default_rng(0)
x = np.concatenate([rng.uniform(1.0, 4.0, 2_000_000), np.exp(rng.uniform(-30, 30, 200_000))])
r = ctx.eval_math(7, x); print("rcp seed   max rel err = 2^%.2f" % np.log2(np.max(np.abs(r * x - 1.0))))
r = ctx.eval_math(8, x); print("rsqrt seed max rel err = 2^%.2f" % np.log2(np.max(np.abs(r * r * x - 1.0)) / 2))
