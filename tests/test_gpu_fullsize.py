"""Full-size GPU parity for BASELINE configs 3 and 4 (stacked N = 32768): the CUDA path through the C ABI against the LAPACK oracle
on a 256-point subsample of the prediction grid, fp64 (north-star gate rtol 1e-5) and split-TF32 (gate 1e-2), plus the
size-independent properties (variance bounds, L v = y, |L L^T z - K z| by random probes through the engine's own hooks).

The oracle side is the plain numpy/scipy restatement (oracle/gp_oracle.py) with K assembled by row blocks so that the host peak
stays at one N x N matrix (8.6 GB); one K-build + dpotrf per configuration (about a minute on the GPU box's host cores), shared by
the fp64 and the tf32 test of that configuration.
"""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import gp_oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

CONFIGS = {
    # name: (n, d, P, kind)  -- BASELINE.json configs[2], configs[3]
    "c3": (16384, 4, 2, "ExpQuad"),
    "c4": (32768, 8, 1, "Matern52"),
}
RTOL_GATE_FP64 = 1e-5   # north_star
RTOL_GATE_TF32 = 1e-2   # north_star


def oracle_subsample(spec, X, y, Xsel, block=4096):
    """orc.factorize + orc.conditional with K built block-row by block-row (same formulas: orc.cov_full / orc.noise_diag)."""
    N = len(y)
    K = np.empty((N, N))
    for i0 in range(0, N, block):
        K[i0:i0 + block] = orc.cov_full(spec, X[i0:i0 + block], X)
    K[np.diag_indices_from(K)] += spec.get("jitter", orc.JITTER_DEFAULT) + orc.noise_diag(spec, X)
    L = sla.cholesky(K, lower=True, check_finite=False, overwrite_a=True)
    del K
    v = sla.solve_triangular(L, y, lower=True, check_finite=False)
    mu, var = orc.conditional(spec, X, L, v, Xsel, True)
    mu_nf, var_nf = orc.conditional(spec, X, L, v, Xsel, False)
    mll = -0.5 * N * np.log(2.0 * np.pi) - np.sum(np.log(np.diag(L))) - 0.5 * float(v @ v)
    return {"mu": mu, "var": var, "mu_nf": mu_nf, "var_nf": var_nf, "mll": mll, "v_head": v[:4096].copy()}


@pytest.fixture(scope="module", params=sorted(CONFIGS))
def full_case(request):
    n, d, P, kind = CONFIGS[request.param]
    spec, X, y, Xs = orc.synthetic_problem(n, d, P=P, M_res=100, kind=kind)
    sel = np.random.default_rng(0).choice(len(Xs), 256, replace=False)
    ref = oracle_subsample(spec, X, y, Xs[sel])
    return {"name": request.param, "spec": spec, "X": X, "y": y, "Xs": Xs, "sel": sel, "ref": ref}


def test_full_size_fp64_against_oracle(lib_built, full_case):
    from gumbi_b200 import GPEngine

    c, ref = full_case, full_case["ref"]
    N, M = len(c["y"]), len(c["Xs"])
    e = GPEngine(0, "fp64")
    try:
        e.set_train(c["X"], c["y"])
        e.set_kernel(c["spec"])
        e.factorize()
        mu, var = e.predict(c["Xs"], True)
        assert mu.shape == (M,) and np.all(np.isfinite(mu)) and np.all(np.isfinite(var))
        sel = c["sel"]
        # the gate of the north star, and what the fp64 path actually delivers at this size (cond(K) ~ 1e6..1e8)
        np.testing.assert_allclose(mu[sel], ref["mu"], rtol=RTOL_GATE_FP64, atol=1e-9)
        np.testing.assert_allclose(var[sel], ref["var"], rtol=RTOL_GATE_FP64, atol=1e-9)
        np.testing.assert_allclose(mu[sel], ref["mu"], rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(var[sel], ref["var"], rtol=1e-7, atol=1e-9)
        mu_nf, var_nf = e.predict(c["Xs"][sel], False)
        np.testing.assert_allclose(mu_nf, ref["mu_nf"], rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(var_nf, ref["var_nf"], rtol=RTOL_GATE_FP64, atol=1e-8)   # noise-free variance: a cancellation residue near the data
        assert e.mll() == pytest.approx(ref["mll"], rel=1e-10)
        np.testing.assert_allclose(e.get_v()[:4096], ref["v_head"], rtol=1e-8, atol=1e-10)
        # properties that do not need the oracle: sigma^2 B_noise <= var <= prior + noise
        s2 = c["spec"]["sigma"] ** 2
        prior = orc.cov_diag(c["spec"], c["Xs"]) + orc.noise_diag(c["spec"], c["Xs"])
        assert np.all(var >= s2 * (1 - 1e-6)) and np.all(var <= prior * (1 + 1e-9))
    finally:
        e.close()


def test_full_size_tf32_against_oracle(lib_built, full_case):
    from gumbi_b200 import GPEngine

    c, ref = full_case, full_case["ref"]
    e = GPEngine(0, "tf32")
    try:
        e.set_train(c["X"], c["y"])
        e.set_kernel(c["spec"])
        e.factorize()
        sel = c["sel"]
        mu, var = e.predict(c["Xs"][sel], True)
        if c["name"] == "c4":
            # BASELINE config 4 is THE tf32 configuration of the north star: its gate (rtol 1e-2) must hold
            np.testing.assert_allclose(mu, ref["mu"], rtol=RTOL_GATE_TF32, atol=1e-3)
            np.testing.assert_allclose(var, ref["var"], rtol=RTOL_GATE_TF32, atol=1e-6)
            assert e.mll() == pytest.approx(ref["mll"], rel=1e-4)
        else:
            # C3 (2-output ICM, d = 4) is an fp64 configuration; split-TF32 is measured on it, not promised: its backward error
            # (~1e-6, fp32 accumulation over 1024 columns) is amplified by |K| / (sigma^2 + jitter), which is ~50x larger here than at
            # C4 (B = W W^T + diag(kappa) scales the prior variance up to ~5, and 16384 points in 4 dimensions sit much closer than
            # 32768 in 8).  Measured on the device (profiles/r02n_pytest_gpu.log): max |d mean| = 0.036 = 1.7e-2 of max |mean|.
            err_mu = np.max(np.abs(mu - ref["mu"])) / np.max(np.abs(ref["mu"]))
            err_var = np.max(np.abs(var - ref["var"])) / np.max(np.abs(ref["var"]))
            print(f"C3 split-TF32: max |d mean| / max |mean| = {err_mu:.2e}, max |d var| / max var = {err_var:.2e}")
            assert err_mu <= 5e-2 and err_var <= 5e-2
            assert e.mll() == pytest.approx(ref["mll"], rel=1e-2)
    finally:
        e.close()
