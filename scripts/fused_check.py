#!/usr/bin/env python
"""Runs on the GPU box: fused persistent kernel vs the multi-kernel path (bit-level), a timing of both at the cfg3 shard, and the
per-phase cycle profile of the fused kernel."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from momentum_b200 import solver as ms
from momentum_b200.problems import humanoid_problem, chain_problem, chain22_problem
from tests import parity

def compare(name, ch, efs, theta0, mode=ms.FUSED_PERSISTENT, **kw):
    B = theta0.shape[0]
    fn = parity.build_function(ch, efs, B)
    outs = []
    for fm in (ms.FUSED_OFF, mode):
        opts = ms.GaussNewtonSolverOptions(regularization=0.05, fused_mode=fm, store_error_history=True, **kw)
        s = ms.GaussNewtonSolver(opts, fn)
        o = s.solve(theta0)
        o["hist"] = s.get_error_history()
        o["fused"] = s.get_fused_profile()["fused"]
        outs.append(o)
    a, b = outs
    print(name, "fused flags", a["fused"], b["fused"], "params identical", np.array_equal(a["params"], b["params"]), "max diff", np.abs(a["params"] - b["params"]).max(),
          "iters identical", np.array_equal(a["iterations"], b["iterations"]), "errors rel diff", np.max(np.abs(a["errors"] - b["errors"]) / np.maximum(np.abs(a["errors"]), 1e-30)),
          "status", a["status"].sum(), b["status"].sum())

ch, efs, theta0, _ = humanoid_problem(64, orientation=True)
compare("humanoid cfg3 fixed 10", ch, efs, theta0, min_iterations=10, max_iterations=10)
compare("humanoid cfg3 converge", ch, efs, theta0, min_iterations=1, max_iterations=30)
compare("humanoid cfg3 fixed 10 gram+chol", ch, efs, theta0, mode=ms.FUSED_GRAM_CHOLESKY, min_iterations=10, max_iterations=10)
ch, efs, theta0, _ = humanoid_problem(7, orientation=False)
compare("humanoid cfg2 B=7", ch, efs, theta0, min_iterations=1, max_iterations=8)
ch, efs, theta0, ts = chain_problem(J=20, B=3, seed=33, families=("position", "orientation", "state", "limit", "plane", "halfplane", "model_parameters"))
compare("chain20 all families (tile path)", ch, efs, ts + 0.05 * theta0, min_iterations=1, max_iterations=5, cholesky_mode=ms.CHOLESKY_TILES_SPARSE)

# timing at the cfg3 shard (and a small rig: chain22 x 8192, where launches dominate the multi-kernel path)
B = 8192
def tile(a, B):
    a = np.asarray(a); return np.tile(a, (B // a.shape[0] + 1,) + (1,) * (a.ndim - 1))[:B]
c1 = chain22_problem()
import copy
efs1 = []
for e in c1[1]:
    e = copy.copy(e); e.targets = tile(e.targets, B); efs1.append(e)
probs = [("chain22", c1[0], efs1, tile(c1[2], B)), ("cfg3", *humanoid_problem(B, orientation=True)[:3])]
for pname, ch, efs, theta0 in probs:
  print("==", pname)
  fn = ms.SkeletonSolverFunction(ch, B, efs)
  fn.upload_targets()
  st = torch.cuda.Stream(); torch.cuda.set_stream(st)
  t0d = torch.from_numpy(theta0.astype(np.float32)).cuda(); td = torch.empty_like(t0d)
  for fm, nm in ((ms.FUSED_OFF, "three kernels"), (ms.FUSED_GRAM_CHOLESKY, "gram+cholesky"), (ms.FUSED_ON, "persistent")):
      s = ms.GaussNewtonSolver(ms.GaussNewtonSolverOptions(min_iterations=10, max_iterations=10, regularization=0.05, fused_mode=fm, cholesky_mode=ms.CHOLESKY_TILES_SPARSE), fn)
      for _ in range(3):
          td.copy_(t0d); s.solve_device(td.data_ptr(), st.cuda_stream)
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      tot = 0.0
      for _ in range(5):
          td.copy_(t0d); torch.cuda.synchronize(); e0.record(); s.solve_device(td.data_ptr(), st.cuda_stream); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
      r = s.get_results()
      print(f"{nm}: {tot / 5:.3f} ms per solve (8192 x 10 it) -> {8192 * 10 / (tot / 5) * 1e3 / 1e6:.2f} M it/s; sum err {r['errors'].sum():.6f}")
      if fm != ms.FUSED_OFF:
          s.set_profiling(True); td.copy_(t0d); s.solve_device(td.data_ptr(), st.cuda_stream); torch.cuda.synchronize(); s.get_results()
          p = s.get_fused_profile(); tot_c = sum(p["phase_cycles"].values())
          print("fused profile: groups", p["groups"], "kernel ms", p["kernel_ms"], "plan", s.get_plan_stats())
          for k, v in p["phase_cycles"].items():
              print(f"   {k:22s} {v:12d} cycles {100.0 * v / max(tot_c, 1):5.1f} %")
