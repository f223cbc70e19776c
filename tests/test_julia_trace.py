"""Replay of a GENUINE Julia reference trace (oracle/julia_reference_trace.jl) if someone with
Julia has generated one under tests/golden/julia/.  Skipped while the files are absent (Julia
is not in the build image).  CPU part: oracle vs Julia chains; GPU part: CUDA path vs Julia."""
from pathlib import Path

import numpy as np
import pytest

J = Path(__file__).resolve().parent / "golden" / "julia"
pytestmark = pytest.mark.skipif(not (J / "meta.txt").exists(), reason="no Julia-generated reference trace present")


def _load():
    name, nw, nitw, nbw, nthin, a = (J / "meta.txt").read_text().split()
    nw, nitw, nbw, nthin, a = int(nw), int(nitw), int(nbw), int(nthin), float(a)
    x0 = np.fromfile(J / "theta0s.f64").reshape(nw, -1)
    d = x0.shape[1]
    ns = (nitw - nbw) // nthin
    return dict(nw=nw, nitw=nitw, nbw=nbw, nthin=nthin, a=a, x0=x0, d=d,
                replay=(np.fromfile(J / "partner.i64", dtype=np.int64), np.fromfile(J / "z.f64"), np.fromfile(J / "u.f64")),
                chain=np.fromfile(J / "chain.f64").reshape(nw, ns, d), logp=np.fromfile(J / "logp.f64").reshape(nw, ns),
                ar=np.fromfile(J / "accept_ratio.f64"))


def test_oracle_matches_julia(orc):
    t = _load()
    r = orc.emcee(orc.Density("rosenbrock", 2, [1.0, 100.0, 20.0]), t["x0"], t["nitw"], t["nbw"], t["nthin"], t["a"],
                  replay=t["replay"])
    assert np.array_equal(r["accept_ratio"], t["ar"])                 # decisions bit-exact
    np.testing.assert_allclose(r["chain_x"], t["chain"], rtol=1e-12)   # FP64 states within 1e-12
    np.testing.assert_allclose(r["chain_lp"], t["logp"], rtol=1e-12)


@pytest.mark.gpu
def test_cuda_matches_julia(km):
    t = _load()
    s = km.Sampler(km.rosenbrock(), t["x0"], t["nitw"], t["nbw"], t["nthin"], t["a"], 0, km.MODE_REPLAY)
    s.set_replay(*t["replay"])
    s.run(-1)
    th, lp, ar = s.results()
    s.close()
    assert np.array_equal(ar, t["ar"])
    np.testing.assert_allclose(th, t["chain"], rtol=1e-12)
    np.testing.assert_allclose(lp, t["logp"], rtol=1e-12)
