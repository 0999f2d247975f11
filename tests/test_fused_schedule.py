"""CPU emulation of the block schedule behind ``gb2_factorize_predict`` (gumbi_b200/csrc/cholesky.cuh, factor_steps, the ``ext`` rows):
the prediction rows E = K(X*, X) ride through the right-looking blocked Cholesky as extra rows of the factor, with the column
blocks grouped ``fused_group`` at a time (narrow left-looking update inside a group, one bulk update when it closes).  The
emulation follows the launch sequence of the CUDA driver loop statement by statement with numpy blocks and checks that E ends as
K(X*, X) L^-T for every group width, block count and ragged last block -- the index arithmetic of the schedule, independent of
the GPU (the device test is tests/test_zz_late_gpu.py)."""
import numpy as np
import pytest


def emulate(K, Ks, T, ncols, w):
    A, E = K.copy(), Ks.copy()
    Np = A.shape[0]
    nb = Np // T
    for k in range(nb):
        g0 = k * T
        Lkk = np.linalg.cholesky(A[g0:g0 + T, g0:g0 + T])                     # potrf_diag_kernel: factor + inverse
        A[g0:g0 + T, g0:g0 + T] = Lkk
        Dk = np.linalg.inv(Lkk)
        below = Np - g0 - T
        if below > 0:
            A[g0 + T:, g0:g0 + T] = A[g0 + T:, g0:g0 + T] @ Dk.T                # panel solve (GM_SET with inv(L_kk)^T)
            P = A[g0 + T:, g0:g0 + T]
            A[g0 + T:, g0 + T:] -= P @ P.T                                     # next-column + bulk trailing update
        if k < ncols:
            gw = (k // w) * w
            p = k - gw
            if p > 0:                                                          # narrow left-looking update inside the group
                E[:, g0:g0 + T] -= E[:, gw * T:g0] @ A[g0:g0 + T, gw * T:g0].T
            E[:, g0:g0 + T] = E[:, g0:g0 + T] @ Dk.T                            # the rows' own panel solve
            if (p == w - 1 or k == ncols - 1) and k + 1 < ncols:               # the group closes: one bulk update to its right
                E[:, g0 + T:ncols * T] -= E[:, gw * T:g0 + T] @ A[g0 + T:ncols * T, gw * T:g0 + T].T
    return np.tril(A), E


@pytest.mark.parametrize("w", [1, 2, 4, 8])
@pytest.mark.parametrize("N", [3, 4, 5, 16, 21, 36])
def test_prediction_rows_end_as_the_solved_panel(N, w):
    T = 4                                                                      # the schedule does not depend on the tile size
    rng = np.random.default_rng(N * 10 + w)
    Np = (N + 1 + T - 1) // T * T                                              # augmented row N (y^T) + identity padding, as dA
    X = rng.standard_normal((N, 2))
    Xs = rng.standard_normal((7, 2))
    k = lambda a, b: np.exp(-0.5 * ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1))
    K = np.eye(Np)
    K[:N, :N] = k(X, X) + 0.1 * np.eye(N)
    y = rng.standard_normal(N)
    K[N, :N] = y                                                               # lower triangle only matters
    K[:N, N] = y
    K[N, N] = 1.0 + y @ np.linalg.solve(K[:N, :N], y)                          # keeps the emulated augmented pivot at 1 (the kernel forces it)
    Ks = np.zeros((7, Np))
    Ks[:, :N] = k(Xs, X)
    ncols = (N + T - 1) // T
    L, E = emulate(K, Ks, T, ncols, w)
    Lnn = np.linalg.cholesky(K[:N, :N])
    np.testing.assert_allclose(L[:N, :N], Lnn, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(E[:, :N], np.linalg.solve(Lnn, Ks[:, :N].T).T, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(L[N, :N], np.linalg.solve(Lnn, y), rtol=1e-9, atol=1e-11)   # the augmented row became v^T


def emulate_two_level(K, T, pw):
    """cholesky_enqueue with the "fp64_panel" option: panels of pw column blocks factored by factor_steps(c0, c1, col_limit=c1)
    (every update restricted to the panel's own columns), then one deep update of the trailing matrix, split into the next panel's
    columns (1) and the rest (2)."""
    A = K.copy()
    Np = A.shape[0]
    nb = Np // T
    blk = lambda b0, b1: slice(b0 * T, b1 * T)
    for c0 in range(0, nb, pw):
        c1 = min(c0 + pw, nb)
        for k in range(c0, c1):                                               # factor_steps(c0, c1, c1)
            Lkk = np.linalg.cholesky(A[blk(k, k + 1), blk(k, k + 1)])
            A[blk(k, k + 1), blk(k, k + 1)] = Lkk
            if k + 1 < nb:
                A[blk(k + 1, nb), blk(k, k + 1)] = A[blk(k + 1, nb), blk(k, k + 1)] @ np.linalg.inv(Lkk).T   # panel solve: ALL rows below
                if k + 1 < c1:                                               # updates: columns (k, c1) only
                    P = A[blk(k + 1, nb), blk(k, k + 1)]
                    A[blk(k + 1, nb), blk(k + 1, c1)] -= P @ A[blk(k + 1, c1), blk(k, k + 1)].T
        if c1 >= nb:
            break
        c2 = min(c1 + pw, nb)
        A[blk(c1, nb), blk(c1, c2)] -= A[blk(c1, nb), blk(c0, c1)] @ A[blk(c1, c2), blk(c0, c1)].T          # (1) next panel's columns
        if c2 < nb:
            A[blk(c2, nb), blk(c2, nb)] -= A[blk(c2, nb), blk(c0, c1)] @ A[blk(c2, nb), blk(c0, c1)].T      # (2) the rest
    return np.tril(A)


@pytest.mark.parametrize("pw", [2, 3, 4])
@pytest.mark.parametrize("nb", [3, 5, 8, 9])
def test_two_level_blocking_gives_the_cholesky_factor(nb, pw):
    T = 3
    rng = np.random.default_rng(nb * 10 + pw)
    B = rng.standard_normal((nb * T, nb * T))
    K = B @ B.T + nb * T * np.eye(nb * T)
    np.testing.assert_allclose(emulate_two_level(K, T, pw), np.linalg.cholesky(K), rtol=1e-10, atol=1e-12)
