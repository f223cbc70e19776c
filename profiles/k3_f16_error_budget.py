"""Error budget of `ex2.approx.ftz.f16x2` in the logistic epilogue, evaluated on the CPU with numpy.

The tcgen05 epilogue computes, per logit s' (log2 units), t = 2^-|s'| and accumulates log2(1 + t).  A half-precision ex2
returns t rounded to 11 significant bits (relative error <= 2^-11, flushed to zero below 2^-14).  This script applies
exactly that rounding to the exact t of the bench problem (N = 10^6 rows, d = 32, walkers in a ball of radius 1e-3) and
reports the error of logp and of the DIFFERENCES of logp between neighbouring walkers -- the quantity the accept test
consumes (src/samplers.jl:260) -- next to the measured error of the shipped FP32 epilogue.
    python profiles/k3_f16_error_budget.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

wl = dict(bench.WORKLOADS["logistic32d"])
params, x0 = bench.make_inputs(wl, 1)
X = np.asarray(wl["_data"][:wl["ndata"] * wl["d"]], dtype=np.float64).reshape(wl["ndata"], wl["d"])
pts = x0[:16]
S = (X @ pts.T) * np.log2(np.e)                      # logits in log2 units, [N, 16]
t = np.exp2(-np.abs(S))
exact = np.log2(1.0 + t).sum(0) * np.log(2.0)
t16 = t.astype(np.float16).astype(np.float64)         # round to nearest, 11 significant bits
t16[t < 2.0 ** -14] = 0.0                             # .ftz: subnormal halves are flushed
half = np.log2(1.0 + t16).sum(0) * np.log(2.0)
err = half - exact
print("N = %d, d = %d, %d neighbouring walkers (ball radius 1e-3)" % (X.shape[0], X.shape[1], len(pts)))
print("f16 ex2:  logp error mean %+.3e  max|.| %.3e ;  neighbour differences: rms %.3e  max|.| %.3e"
      % (err.mean(), np.abs(err).max(), np.diff(err).std(), np.abs(np.diff(err)).max()))
print("shipped FP32 epilogue (profiles/r2_k3_f32x2.log, 512 points): neighbour differences rms 6.1e-04 max 2.0e-03")
print("ratio of the rms errors of the accept test's input: %.0fx" % (np.diff(err).std() / 6.1e-4))
