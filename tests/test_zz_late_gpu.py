"""Device tests of what was built after round 1's GPU minutes were spent: the Kronecker-aware multi-output solve
(gumbi_b200/kron.py), ``gb2_get_alpha``, the ``solve_streams`` option, the fused cold predict (``gb2_factorize_predict``) and the
periodic-kernel lowering -- CUDA engines through the C ABI against the dense CUDA path, the oracle and the committed golden vectors.
(The Kronecker blocks and the periodic kernels only use device code paths the parity tests above already cover.)"""
import numpy as np
import pytest

from conftest import load_golden
from test_kron import from_golden, random_point, synthetic

pytestmark = [pytest.mark.gpu]


def pair(X, y, kw, **build):
    from gumbi_b200 import ArrayGP

    dense = ArrayGP(X, y, **kw)
    dense.build_model(**build)
    kr = ArrayGP(X, y, multioutput="kron", **kw)
    kr.build_model(**build)
    return dense, kr


@pytest.mark.parametrize("kernel,extra_cat,hetero", [("ExpQuad", 0, True), ("Matern52", 3, True), ("Matern32", 0, False)])
def test_kron_matches_dense_cuda_path(lib_built, kernel, extra_cat, hetero):
    from gumbi_b200 import kron
    from gumbi_b200.map import named_gradient

    X, y, kw = synthetic(n=301, P=3, d=2, extra_cat=extra_cat, seed=7)
    dense, kr = pair(X, y, kw, continuous_kernel=kernel, heteroskedastic_outputs=hetero)
    assert isinstance(kr.engine, kron.KronEngine)
    pt = random_point(dense, 13)
    dense.find_MAP(point=pt)
    kr.find_MAP(point=pt)
    rng = np.random.default_rng(1)
    pts = X[rng.integers(0, len(X), 500)].copy()
    pts[:, :2] += 0.2 * rng.standard_normal((500, 2))
    pts[:, -1] = rng.integers(0, 3, 500)
    for noise in (True, False):
        mu_d, var_d = dense.predict(pts, with_noise=noise)
        mu_k, var_k = kr.predict(pts, with_noise=noise)
        np.testing.assert_allclose(mu_k, mu_d, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(var_k, var_d, rtol=1e-6, atol=1e-9)
    assert kr.marginal_log_likelihood() == pytest.approx(dense.marginal_log_likelihood(), rel=1e-9)
    m_d, c_d = dense.conditional(pts[:40], pred_noise=True)
    m_k, c_k = kr.conditional(pts[:40], pred_noise=True)
    np.testing.assert_allclose(m_k, m_d, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(c_k, c_d, rtol=1e-6, atol=1e-9)
    spec = dense.spec_from_point(dense.MAP)
    v_d, g_d = dense.engine.mll_grad(spec)
    v_k, g_k = kr.engine.mll_grad(spec)
    assert v_k == pytest.approx(v_d, rel=1e-9)
    n_d, n_k = named_gradient(dense, g_d), named_gradient(kr, g_k)
    for name in n_d:
        np.testing.assert_allclose(n_k[name], n_d[name], rtol=1e-5, atol=1e-6, err_msg=name)
    assert kr.predict(pts[:0])[0].shape == (0,)
    seq = kr.predict_cold(pts)
    kr.engine.threads = 3                                         # all blocks in flight at once: same arithmetic per block
    par = kr.predict_cold(pts)
    assert np.array_equal(seq[0], par[0]) and np.array_equal(seq[1], par[1])
    dense.engine.close()
    kr.engine.close()


def test_kron_on_the_reference_shaped_golden(lib_built):
    from gumbi_b200 import ArrayGP

    g = load_golden("multioutput_regression")
    gp = from_golden(g, ArrayGP, multioutput="kron")
    gp.find_MAP(point=g["meta"]["point"])
    mu, var = gp.predict(g["points"], with_noise=True)
    np.testing.assert_allclose(mu, g["mean"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(var, g["var"], rtol=1e-6, atol=1e-9)
    gp.engine.close()


def test_find_map_on_device_blocks(lib_built):
    X, y, kw = synthetic(n=120, P=3, d=1, seed=8)
    dense, kr = pair(X, y, kw)
    a = dense.find_MAP(options={"maxiter": 15})
    b = kr.find_MAP(options={"maxiter": 15})
    for name in a:
        np.testing.assert_allclose(b[name], a[name], rtol=1e-4, atol=1e-6, err_msg=name)
    dense.engine.close()
    kr.engine.close()


def test_get_alpha_needs_the_gradient_of_the_current_factor(lib_built):
    from gumbi_b200 import GPEngine
    from oracle import gp_oracle as orc

    rng = np.random.default_rng(0)
    X, y = rng.standard_normal((200, 2)), rng.standard_normal(200)
    spec = {"terms": [{"kind": "ExpQuad", "cont_idx": [0, 1], "ls": [1.0, 1.3], "eta": 1.1, "lin_idx": [], "c": [], "tau": 0.0, "coreg": []}],
            "sigma": 0.2, "noise_coreg": None, "jitter": 1e-6}
    e = GPEngine()
    e.set_train(X, y)
    e.set_kernel(spec)
    e.factorize()
    with pytest.raises(ValueError):
        e.get_alpha()
    e.mll_grad(spec)
    L, v = orc.factorize(spec, X, y)
    from scipy.linalg import solve_triangular

    np.testing.assert_allclose(e.get_alpha(), solve_triangular(L, v, lower=True, trans="T"), rtol=1e-7, atol=1e-9)
    e.factorize()
    with pytest.raises(ValueError):
        e.get_alpha()
    e.close()


@pytest.mark.parametrize("streams", [2, 4])
def test_solve_streams_gives_identical_posterior(lib_built, streams):
    """"solve_streams": the prediction rows are solved as independent slabs on concurrent streams -- same arithmetic per row, so
    the posterior is bit-identical to the single-stream solve (also with fewer row tiles than streams)."""
    from gumbi_b200 import GPEngine
    from oracle import gp_oracle as orc

    spec, X, y, Xs = orc.synthetic_problem(900, 3, M_res=25)      # M = 625 -> 5 row tiles
    e = GPEngine()
    e.set_train(X, y)
    e.set_kernel(spec)
    e.factorize()
    ref = e.predict(Xs)
    small = e.predict(Xs[:100])
    e.set_option("solve_streams", streams)
    got = e.predict(Xs)
    got_small = e.predict(Xs[:100])                               # one row tile: falls back to a single stream
    e.set_option("solve_streams", 1)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    assert np.array_equal(got_small[0], small[0]) and np.array_equal(got_small[1], small[1])
    with pytest.raises(ValueError):
        e.set_option("solve_streams", 9)
    e.close()


@pytest.mark.parametrize("n,d,P,kind", [(900, 3, 1, "ExpQuad"), (1000, 2, 1, "Matern52"), (127, 1, 1, "ExpQuad"), (300, 2, 2, "ExpQuad")])
def test_fused_cold_predict_matches_factorize_then_predict(lib_built, n, d, P, kind):
    """gb2_factorize_predict: prediction points carried through the factorisation as extra rows of the factor -- same posterior as
    gb2_factorize + gb2_predict (different summation order only), same factor left behind, same oracle numbers."""
    from gumbi_b200 import GPEngine
    from oracle import gp_oracle as orc

    spec, X, y, Xs = orc.synthetic_problem(n, d, P=P, M_res=21 if d >= 2 else 77, kind=kind)
    e = GPEngine()
    e.set_train(X, y)
    e.set_kernel(spec)
    e.factorize()
    L_ref = e.get_L()
    ref = {noise: e.predict(Xs, noise) for noise in (True, False)}
    for noise in (True, False):
        mu, var = e.factorize_predict(Xs, noise)
        np.testing.assert_allclose(mu, ref[noise][0], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(var, ref[noise][1], rtol=1e-7, atol=1e-10)
    for group in (1, 2, 8):                                       # column blocks per bulk update of the prediction rows (default 4)
        e.set_option("fused_group", group)
        mu, var = e.factorize_predict(Xs, True)
        np.testing.assert_allclose(mu, ref[True][0], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(var, ref[True][1], rtol=1e-7, atol=1e-10)
    e.set_option("fused_group", 4)
    assert np.array_equal(e.get_L(), L_ref)                       # the factorisation itself is untouched by the extra rows
    mu2, var2 = e.predict(Xs[:50], True)                          # and the handle is left factorised
    np.testing.assert_allclose(mu2, ref[True][0][:50], rtol=1e-12, atol=1e-13)
    L, v = orc.factorize(spec, X, y)
    mu_o, var_o = orc.conditional(spec, X, L, v, Xs, True)
    mu, var = e.factorize_predict(Xs, True)
    np.testing.assert_allclose(mu, mu_o, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(var, var_o, rtol=1e-5, atol=1e-8)
    import torch

    dXs = torch.from_numpy(np.ascontiguousarray(Xs)).cuda()        # device-pointer twin: no host copies
    dout = torch.zeros(2 * len(Xs), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    e.factorize_predict_device(dXs.data_ptr(), len(Xs), True, dout.data_ptr(), dout.data_ptr() + 8 * len(Xs))
    assert np.array_equal(dout[:len(Xs)].cpu().numpy(), mu) and np.array_equal(dout[len(Xs):].cpu().numpy(), var)
    mu1, var1 = e.factorize_predict(Xs[:1], True)                 # a single point
    np.testing.assert_allclose(mu1, mu_o[:1], rtol=1e-6, atol=1e-8)
    # Indefinite K: a Linear term with tau < 0 makes K_ii = eta^2 - 4 |x_i - c|^2 negative for every point with |x_i| > 1/2.  (An
    # exactly duplicated row with sigma = jitter = 0 only leaves a rounding residue of either sign in the pivot -- LAPACK would
    # not flag it reliably either; the first device run of this test showed exactly that for the Matern52 case.)
    bad = dict(spec, terms=[dict(spec["terms"][0], lin_idx=[0], c=[0.0], tau=-4.0)] + list(spec["terms"][1:]))
    e.set_kernel(bad)
    with pytest.raises(np.linalg.LinAlgError, match="not positive definite"):
        e.factorize_predict(Xs, True)
    with pytest.raises(np.linalg.LinAlgError, match="not positive definite"):
        e.factorize()
    e.close()


@pytest.mark.parametrize("kernel,d", [("Periodic", 2), ("Matern52+Periodic", 1)])
def test_periodic_kernels_on_the_device(lib_built, kernel, d):
    """Periodic kernels are lowered host-side onto the stationary CUDA kernels over warped coordinates (backend._periodic_warp);
    posterior and likelihood against the reference formulation restated in the oracle (pm.gp.cov.Periodic / WarpedInput)."""
    from gumbi_b200 import ArrayGP
    from oracle import gp_oracle as orc
    from test_backend_host import reference_form

    rng = np.random.default_rng(4)
    X = rng.standard_normal((400, d))
    y = np.sin(3 * X[:, 0]) + 0.1 * rng.standard_normal(400)
    dims = [f"x{j}" for j in range(d)]
    zp = [1.7 + 0.4 * j for j in range(d)]
    gp = ArrayGP(X, y, dims, linear_dims=dims[:1])
    gp.build_model(continuous_kernel=kernel, period=dict(zip(dims, zp)))
    gp.find_MAP(point={"ls_total": rng.uniform(0.5, 1.5, size=d), "η_total": 1.3, "σ": 0.2, "c_total": [0.1], "τ_total": 0.3})
    ref = reference_form(gp, gp.spec_from_point(gp.MAP), "Periodic" if kernel == "Periodic" else "warped", zp)
    Xs = rng.standard_normal((300, d))
    mu, var = gp.predict(Xs, with_noise=True)
    mu0, var0 = orc.predict(ref, X, y, Xs, True)
    np.testing.assert_allclose(mu, mu0, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(var, var0, rtol=1e-5, atol=1e-8)
    assert gp.marginal_log_likelihood() == pytest.approx(orc.mll(ref, X, y), rel=1e-8)
    gp.engine.close()


def test_timeline_trace_is_ordered(lib_built):
    """set_option("trace", 1) / gb2_get_trace: six %globaltimer stamps per block step, ordered along each stream; the factor is the
    same with and without the stamps."""
    from gumbi_b200 import GPEngine
    from oracle import gp_oracle as orc

    spec, X, y, _ = orc.synthetic_problem(700, 3)
    e = GPEngine()
    e.set_train(X, y)
    e.set_kernel(spec)
    e.factorize()
    L0 = e.get_L()
    with pytest.raises(ValueError):
        e.get_trace()
    e.set_option("trace", 1)
    e.factorize()
    t = e.get_trace().astype(np.int64)
    assert t.shape == ((700 + 1 + 127) // 128, 6)
    assert np.array_equal(e.get_L(), L0)
    steps = t.shape[0]
    assert np.all(t[:, 0] > 0) and np.all(t[:, 1] >= t[:, 0])                    # diagonal kernel: eligible <= done
    assert np.all(t[:-1, 2] >= t[:-1, 1]) and np.all(t[:-1, 5] >= t[:-1, 4])     # panel solve after it; bulk update: eligible <= done
    assert np.all(np.diff(t[:, 0]) > 0)                                          # the panel stream runs the steps in order
    e.set_option("trace", 0)
    e.factorize()
    assert np.array_equal(e.get_L(), L0)
    e.close()


@pytest.mark.parametrize("pw", [2, 4])
def test_two_level_blocking_of_the_fp64_factorisation(lib_built, pw):
    """set_option("fp64_panel", pw): panels of pw column blocks, one deep update per panel (split: next panel's columns / the rest on a
    second bulk stream).  Different summation grouping only: factor, v, log-likelihood and posterior agree to rounding; fused path too."""
    from gumbi_b200 import GPEngine
    from oracle import gp_oracle as orc

    for n, d in ((1500, 3), (700, 2), (300, 2)):
        spec, X, y, Xs = orc.synthetic_problem(n, d, M_res=15)
        e = GPEngine()
        e.set_train(X, y)
        e.set_kernel(spec)
        e.factorize()
        L0, v0, m0, p0 = e.get_L(), e.get_v(), e.mll(), e.predict(Xs)
        e.set_option("fp64_panel", pw)
        e.factorize()
        np.testing.assert_allclose(e.get_L(), L0, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(e.get_v(), v0, rtol=1e-8, atol=1e-10)
        assert e.mll() == pytest.approx(m0, rel=1e-10)
        p1 = e.predict(Xs)
        np.testing.assert_allclose(p1[0], p0[0], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(p1[1], p0[1], rtol=1e-6, atol=1e-10)
        mu, var = e.factorize_predict(Xs, True)
        np.testing.assert_allclose(mu, p0[0], rtol=1e-8, atol=1e-10)
        e.close()
