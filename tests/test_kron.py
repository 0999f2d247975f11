"""Kronecker-aware multi-output solve (gumbi_b200/kron.py): P independent n x n problems must give the numbers of
the dense stacked (nP x nP) system the reference factorises (gumbi/regression/pymc/GP.py:724-727, :560-569, :580, :845-847).

CPU: the host arithmetic (rotation, recombination, gradient chain rule, distribution of blocks over ranks) around oracle block
engines against the dense oracle.  GPU: CUDA block engines against the dense CUDA path and the committed golden vectors."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden
from gumbi_b200 import kron
from gumbi_b200.backend import ArrayRegressor, B200Backend
from gumbi_b200.map import named_gradient
from test_backend_host import HostGP, OracleEngine


class HostKronGP(B200Backend, ArrayRegressor):
    """Kronecker path with oracle block engines (CPU)."""

    def __init__(self, *a, multioutput="kron", **k):
        ArrayRegressor.__init__(self, *a, **k)
        self._init_backend(multioutput=multioutput)

    def _make_block_engine(self):
        return OracleEngine()

    def _make_dense_engine(self):
        return OracleEngine()

    def _wants_kron_for_test(self):
        return self.build_model().engine


def from_golden(g, cls, **kw):
    m = g["meta"]
    gp = cls(g["X"], g["y"], m["continuous_dims"], linear_dims=m["linear_dims"], categorical_dims=m["categorical_dims"],
             categorical_levels=dict(m["categorical_levels"]), out_col=m["out_col"], outputs=m["outputs"], additive=m["additive"], **kw)
    gp.build_model(continuous_kernel=m["continuous_kernel"], ARD=m["ARD"])
    return gp


def synthetic(n=37, P=3, d=2, extra_cat=0, seed=0, linear=True):
    """Aligned stacked data: [continuous..., (other categorical), output coordinate]."""
    rng = np.random.default_rng(seed)
    Xc = rng.standard_normal((n, d))
    cols = [Xc]
    cat_dims, levels = [], {}
    if extra_cat:
        cols.append(rng.integers(0, extra_cat, size=(n, 1)).astype(float))
        cat_dims.append("Cat")
        levels["Cat"] = [f"l{i}" for i in range(extra_cat)]
    Xb = np.hstack(cols)
    X = np.vstack([np.hstack([Xb, np.full((n, 1), float(p))]) for p in range(P)])
    y = np.concatenate([np.sin(Xc @ rng.standard_normal(d) + p) + 0.1 * rng.standard_normal(n) for p in range(P)])
    dims = [f"x{j}" for j in range(d)]
    kw = dict(continuous_dims=dims, linear_dims=dims[:1] if linear else None, categorical_dims=cat_dims + ["Variable"],
              categorical_levels={**levels, "Variable": [f"o{p}" for p in range(P)]}, out_col="Variable",
              outputs=[f"o{p}" for p in range(P)])
    return X, y, kw


def random_point(gp, seed):
    rng = np.random.default_rng(seed)
    pt = {}
    for name, shape in gp.param_shapes().items():
        base = name.split("_")[0]
        if base in ("ls", "η", "τ", "κ", "σ"):
            pt[name] = rng.uniform(0.3, 1.5, size=shape)
        else:
            pt[name] = rng.standard_normal(size=shape)
    return pt


def test_alignment_detection():
    X, y, kw = synthetic(n=5, P=3)
    Xb, rows = kron.aligned_blocks(X, X.shape[1] - 1, 3)
    assert Xb.shape == (5, 2) and rows.shape == (3, 5)
    assert np.array_equal(X[rows[2], :-1], Xb) and np.all(X[rows[2], -1] == 2)
    with pytest.raises(kron.NotAligned):                       # one observation missing (NaN row dropped, base.py:469-471)
        kron.aligned_blocks(X[:-1], X.shape[1] - 1, 3)
    X2 = X.copy()
    X2[7, 0] += 1e-9
    with pytest.raises(kron.NotAligned):                       # outputs observed at different inputs
        kron.aligned_blocks(X2, X.shape[1] - 1, 3)
    X3 = X.copy()
    X3[0, -1] = 3
    with pytest.raises(kron.NotAligned):
        kron.aligned_blocks(X3, X.shape[1] - 1, 3)
    # interleaved order (output coordinate varying fastest) is still aligned
    perm = np.arange(15).reshape(3, 5).T.reshape(-1)
    Xb4, rows4 = kron.aligned_blocks(X[perm], X.shape[1] - 1, 3)
    assert np.array_equal(Xb4, Xb) and np.array_equal(perm[rows4], rows)


def test_backend_selects_the_solver():
    X, y, kw = synthetic(n=6, P=2)
    assert isinstance(HostKronGP(X, y, **kw).build_model().engine, kron.KronEngine)
    assert isinstance(HostKronGP(X, y, multioutput="auto", **kw).build_model().engine, kron.KronEngine)
    with pytest.raises(ValueError, match="not applicable"):
        HostKronGP(X[:-1], y[:-1], **kw).build_model()
    Xa, ya, kwa = synthetic(n=6, P=2, extra_cat=2)
    with pytest.raises(ValueError, match="not applicable"):
        HostKronGP(Xa, ya, additive=True, **kwa).build_model()
    assert not isinstance(HostKronGP(Xa, ya, additive=True, multioutput="auto", **kwa)._wants_kron_for_test(), kron.KronEngine)
    with pytest.raises(ValueError, match="single output"):
        HostKronGP(X[:, :2], y, continuous_dims=["x0", "x1"]).build_model()
    with pytest.raises(ValueError, match="multioutput"):
        HostKronGP(X, y, multioutput="eig", **kw)


@pytest.mark.parametrize("hetero", [True, False])
@pytest.mark.parametrize("extra_cat", [0, 3])
def test_kron_equals_dense_oracle(hetero, extra_cat):
    X, y, kw = synthetic(n=41, P=4, d=2, extra_cat=extra_cat, seed=3)
    dense = HostGP(X, y, **kw)
    dense.build_model(continuous_kernel="Matern52", heteroskedastic_outputs=hetero)
    kr = HostKronGP(X, y, **kw)
    kr.build_model(continuous_kernel="Matern52", heteroskedastic_outputs=hetero)
    pt = random_point(dense, 11)
    dense.find_MAP(point=pt)
    kr.find_MAP(point=pt)
    rng = np.random.default_rng(5)
    pts = X[rng.integers(0, len(X), 60)].copy()
    pts[:, :2] += 0.3 * rng.standard_normal((60, 2))
    pts[:, -1] = rng.integers(0, 4, 60)                       # ragged: not every output at every point
    for noise in (True, False):
        mu_d, var_d = dense.predict(pts, with_noise=noise)
        mu_k, var_k = kr.predict(pts, with_noise=noise)
        np.testing.assert_allclose(mu_k, mu_d, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(var_k, var_d, rtol=1e-8, atol=1e-12)
    assert kr.marginal_log_likelihood() == pytest.approx(dense.marginal_log_likelihood(), rel=1e-11)
    m_d, c_d = dense.conditional(pts[:25], pred_noise=True)
    m_k, c_k = kr.conditional(pts[:25], pred_noise=True)
    np.testing.assert_allclose(m_k, m_d, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(c_k, c_d, rtol=1e-8, atol=1e-11)
    # gradient: same named entries as the dense engine's
    spec = dense.spec_from_point(dense.MAP)
    v_d, g_d = dense.engine.mll_grad(spec)
    v_k, g_k = kr.engine.mll_grad(spec)
    assert v_k == pytest.approx(v_d, rel=1e-11)
    n_d, n_k = named_gradient(dense, g_d), named_gradient(kr, g_k)
    assert set(n_d) == set(n_k)
    for name in n_d:
        np.testing.assert_allclose(n_k[name], n_d[name], rtol=1e-7, atol=1e-8, err_msg=name)
    # one solve per distinct input row and block, whatever the number of outputs predicted there
    tiled = np.vstack([np.hstack([pts[:10, :-1], np.full((10, 1), float(p))]) for p in range(4)])
    seen = []
    for e in kr.engine.blocks.values():
        e.predict = (lambda f: (lambda Xs, pred_noise=True: (seen.append(len(Xs)), f(Xs, pred_noise))[1]))(e.predict)
    kr.predict(tiled)
    assert seen == [10, 10, 10, 10]


def test_kron_on_the_reference_shaped_golden():
    g = load_golden("multioutput_regression")
    gp = from_golden(g, HostKronGP)
    gp.find_MAP(point=g["meta"]["point"])
    mu, var = gp.predict(g["points"], with_noise=True)
    np.testing.assert_allclose(mu, g["mean"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, g["var"], rtol=1e-8, atol=1e-11)


def test_find_map_through_the_kron_objective_matches_dense():
    X, y, kw = synthetic(n=30, P=3, d=1, seed=8)
    maps = []
    for cls in (HostGP, HostKronGP):
        gp = cls(X, y, **kw)
        gp.build_model()
        maps.append(gp.find_MAP(options={"maxiter": 25}))
    for name in maps[0]:
        np.testing.assert_allclose(maps[1][name], maps[0][name], rtol=1e-5, atol=1e-7, err_msg=name)


def test_error_behaviour():
    X, y, kw = synthetic(n=8, P=2)
    gp = HostKronGP(X, y, **kw)
    gp.build_model()
    with pytest.raises(RuntimeError):
        gp.predict(X[:3])
    gp.find_MAP(point=random_point(gp, 1))
    bad = X[:3].copy()
    bad[0, -1] = 5
    with pytest.raises(ValueError):
        gp.predict(bad)
    with pytest.raises(ValueError):
        gp.predict(X[:3, :2])
    mu, var = gp.predict(X[:0])
    assert mu.shape == (0,) and var.shape == (0,)
    pt = random_point(gp, 1)
    pt["κ_Variable"] = np.array([1e-200, 1e-200])             # B ~ 0: K is the noise alone, still positive definite
    pt["W_Variable"] = np.zeros((2, 2))
    gp.find_MAP(point=pt)
    dense = HostGP(X, y, **kw)
    dense.build_model()
    dense.find_MAP(point=pt)
    np.testing.assert_allclose(gp.predict(X[:3])[1], dense.predict(X[:3])[1], rtol=1e-10)
    with pytest.raises(np.linalg.LinAlgError):
        kron.rotation(np.array([[1.0, 2.0], [2.0, 1.0]]), np.ones(2))


KRON_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from test_kron import HostKronGP, synthetic, random_point
from test_backend_host import HostGP
from gumbi_b200.map import named_gradient
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
X, y, kw = synthetic(n=33, P=3, d=2, seed=4)
gp = HostKronGP(X, y, **kw)
gp.distributed = True
gp.build_model()
assert gp.engine.kron_world == world and gp.engine.mine == [q for q in range(3) if q % world == rank]
dense = HostGP(X, y, **kw); dense.build_model()
pt = random_point(dense, 2)
gp.find_MAP(point=pt); dense.find_MAP(point=pt)
mu, var = gp.predict(X[::3]); mu_d, var_d = dense.predict(X[::3])
assert sorted(gp.engine.blocks) == gp.engine.mine             # only this rank's blocks were factorised here
np.testing.assert_allclose(mu, mu_d, rtol=1e-9, atol=1e-11); np.testing.assert_allclose(var, var_d, rtol=1e-8, atol=1e-12)
spec = dense.spec_from_point(dense.MAP)
v, g = gp.engine.mll_grad(spec); v_d, g_d = dense.engine.mll_grad(spec)
assert abs(v - v_d) < 1e-9 * abs(v_d)
a, b = named_gradient(gp, g), named_gradient(dense, g_d)
for name in b:
    np.testing.assert_allclose(a[name], b[name], rtol=1e-7, atol=1e-8, err_msg=name)
dist.barrier(); dist.destroy_process_group()
print("WORKER_OK", rank)
"""


def test_blocks_dealt_over_two_ranks_gloo(tmp_path):
    script = tmp_path / "kron_worker.py"
    script.write_text(KRON_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29537", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"WORKER_OK {r}" in o, o


def test_blocks_driven_from_several_host_threads_give_identical_results():
    X, y, kw = synthetic(n=35, P=4, d=2, seed=9)
    a = HostKronGP(X, y, **kw)
    a.build_model()
    b = HostKronGP(X, y, **kw)
    b.build_model()
    b.engine.threads = 3
    pt = random_point(a, 6)
    a.find_MAP(point=pt)
    b.find_MAP(point=pt)
    ra, rb = a.predict(X[::2]), b.predict(X[::2])
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
    assert a.marginal_log_likelihood() == b.marginal_log_likelihood()
    spec = a.spec_from_point(a.MAP)
    ga, gb = a.engine.mll_grad(spec), b.engine.mll_grad(spec)
    assert ga[0] == gb[0]
    np.testing.assert_array_equal(ga[1]["terms"][0]["coreg"][0]["W"], gb[1]["terms"][0]["coreg"][0]["W"])


def test_kron_with_a_periodic_kernel_shifts_the_warped_columns():
    """Periodic lowering appends warped columns AFTER the output column; the block problems drop the output column, so every
    column index above it moves down by one (kron.split_spec)."""
    X, y, kw = synthetic(n=29, P=3, d=2, seed=12, linear=True)
    period = {"x0": 1.9, "x1": 2.4}
    dense = HostGP(X, y, **kw)
    dense.build_model(continuous_kernel="Periodic", period=period)
    kr = HostKronGP(X, y, **kw)
    kr.build_model(continuous_kernel="Periodic", period=period)
    pt = random_point(dense, 3)
    dense.find_MAP(point=pt)
    kr.find_MAP(point=pt)
    pts = X[::3].copy()
    pts[:, 0] += 0.1
    mu_d, var_d = dense.predict(pts)
    mu_k, var_k = kr.predict(pts)
    np.testing.assert_allclose(mu_k, mu_d, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var_k, var_d, rtol=1e-8, atol=1e-12)
    spec = dense.spec_from_point(dense.MAP)
    n_d, n_k = named_gradient(dense, dense.engine.mll_grad(spec)[1]), named_gradient(kr, kr.engine.mll_grad(spec)[1])
    for name in n_d:
        np.testing.assert_allclose(n_k[name], n_d[name], rtol=1e-7, atol=1e-8, err_msg=name)


def test_fused_cold_predict_through_the_blocks():
    """predict_cold(fused=True) on the Kronecker solver: every block uses its engine's one-pass entry point."""
    class Fused(OracleEngine):
        n_fused = 0

        def factorize_predict(self, Xs, pred_noise=True):
            Fused.n_fused += 1
            self.factorize()
            return self.predict(Xs, pred_noise)

    class G(HostKronGP):
        def _make_block_engine(self):
            return Fused()

    X, y, kw = synthetic(n=31, P=3, d=2, seed=2)
    gp = G(X, y, **kw)
    gp.build_model()
    ref = HostKronGP(X, y, **kw)
    ref.build_model()
    pt = random_point(gp, 4)
    gp.find_MAP(point=pt)
    ref.find_MAP(point=pt)
    pts = X[::2]
    for noise in (True, False):
        a = gp.predict_cold(pts, with_noise=noise, fused=True)
        b = ref.predict(pts, with_noise=noise)
        np.testing.assert_allclose(a[0], b[0], rtol=1e-12)
        np.testing.assert_allclose(a[1], b[1], rtol=1e-12)
    assert Fused.n_fused == 6
    np.testing.assert_allclose(gp.predict(pts)[0], ref.predict(pts)[0], rtol=1e-12)   # the factors stay resident
