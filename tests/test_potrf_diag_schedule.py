"""Host-side emulation of the SCHEDULE of the diagonal-panel kernel (gumbi_b200/csrc/cholesky.cuh::potrf_diag_kernel): which block product of
the inverse assembly runs in which sub-phase of the 128x128 factorisation.  The kernel issues every product as soon as its operands are
final, on the warps that idle beside warp 0's pivot chains / substitutions; sub-phases are separated by __syncthreads, so a job may only
read what an EARLIER sub-phase wrote.  The emulation enforces exactly that: every job of a sub-phase reads a snapshot taken at the start
of the sub-phase and writes the live state -- a job that depended on a same-phase write would read stale data and the final inverse would
be wrong.  (The arithmetic itself is verified on the device: tools/micro_potrf, tests/test_gpu_parity.py.)"""
import copy

import numpy as np
import scipy.linalg as sla

PB = 32


def _spd(seed):
    rng = np.random.default_rng(seed)
    i = np.arange(4 * PB)
    A = np.exp(-0.5 * (i[:, None] - i[None, :]) ** 2 / 400.0) + 0.01 * np.eye(4 * PB)
    return A + 1e-3 * np.diag(rng.random(4 * PB))


def _blk(M, i, j):
    return M[i * PB:(i + 1) * PB, j * PB:(j + 1) * PB]


class State:
    """L: the 128x128 block being factored in place (lower sub-blocks); X[(i, j)]: T_ij, then X_ij (the kernel keeps them transposed in
    one buffer -- irrelevant here); Xd[p]: the diagonal inverses."""

    def __init__(self, A):
        self.L = np.tril(A).copy()
        self.X = {}
        self.Xd = {}


# side jobs, named as in the kernel
def job_A(j):            # T_ij = L_ij X_jj, i > j
    def run(old, new):
        for i in range(j + 1, 4):
            new.X[(i, j)] = _blk(old.L, i, j) @ old.Xd[j]
    return run


def job_B(j):            # X_{j+1,j} = -X_{j+1,j+1} T_{j+1,j}
    def run(old, new):
        new.X[(j + 1, j)] = -old.Xd[j + 1] @ old.X[(j + 1, j)]
    return run


def job_C1(j):           # T_{j+2,j} += L_{j+2,j+1} X_{j+1,j}
    def run(old, new):
        new.X[(j + 2, j)] = old.X[(j + 2, j)] + _blk(old.L, j + 2, j + 1) @ old.X[(j + 1, j)]
    return run


def job_C2(j):           # X_{j+2,j} = -X_{j+2,j+2} T_{j+2,j}
    def run(old, new):
        new.X[(j + 2, j)] = -old.Xd[j + 2] @ old.X[(j + 2, j)]
    return run


def job_D1(old, new):    # T_30 += L_31 X_10 + L_32 X_20
    new.X[(3, 0)] = old.X[(3, 0)] + _blk(old.L, 3, 1) @ old.X[(1, 0)] + _blk(old.L, 3, 2) @ old.X[(2, 0)]


def job_last(old, new):  # X_3j = -X_33 T_3j
    for j in range(3):
        new.X[(3, j)] = -old.Xd[3] @ old.X[(3, j)]


# main work of the three sub-phases of step p
def main_potrf(p):
    def run(old, new):
        Lpp = _blk(old.L, p, p)
        _blk(new.L, p, p)[:] = np.linalg.cholesky(Lpp + np.tril(Lpp, -1).T)          # the state holds the lower triangle only
    return run


def main_subst(p):       # inv32 (warp 0) || trsm32 (warps 1 .. 3-p)
    def run(old, new):
        Lpp = _blk(old.L, p, p)
        new.Xd[p] = sla.solve_triangular(Lpp, np.eye(PB), lower=True)
        for i in range(p + 1, 4):
            _blk(new.L, i, p)[:] = sla.solve_triangular(Lpp, _blk(old.L, i, p).T, lower=True).T
    return run


def main_update(p):      # C_ij -= L_ip L_jp^T, p < j <= i
    def run(old, new):
        for i in range(p + 1, 4):
            for j in range(p + 1, i + 1):
                _blk(new.L, i, j)[:] = _blk(old.L, i, j) - _blk(old.L, i, p) @ _blk(old.L, j, p).T
    return run


# the schedule of the kernel: (step, sub-phase) -> side jobs beside the main work
SIDE = {
    (1, "potrf"): [job_A(0)],
    (2, "potrf"): [job_A(1), job_B(0)],
    (2, "subst"): [job_C1(0)],
    (2, "update"): [job_C2(0), job_B(1)],
    (3, "potrf"): [job_C1(1), job_D1, job_A(2)],
}


def run_schedule(A, side=SIDE):
    st = State(A)
    for p in range(4):
        for name, main in (("potrf", main_potrf(p)), ("subst", main_subst(p)), ("update", main_update(p))):
            if name == "update" and p == 3:
                continue
            old = copy.deepcopy(st)                    # what a job may read: the state at the last __syncthreads
            main(old, st)
            for job in side.get((p, name), []):
                job(old, st)
    old = copy.deepcopy(st)
    job_last(old, st)
    X = np.zeros_like(A)
    for p in range(4):
        _blk(X, p, p)[:] = st.Xd[p]
    for (i, j), v in st.X.items():
        _blk(X, i, j)[:] = v
    return np.tril(st.L), X


def test_schedule_respects_the_sub_phase_dependencies():
    A = _spd(0)
    L, X = run_schedule(A)
    np.testing.assert_allclose(L @ L.T, A, atol=1e-13)
    np.testing.assert_allclose(X @ L, np.eye(4 * PB), atol=1e-11)
    assert np.array_equal(X, np.tril(X))


def test_a_job_moved_one_sub_phase_early_is_caught():
    """The emulation has teeth: C2(0) needs T_20 complete (C1(0), step 2 substitutions) and X_22 (the same sub-phase's diagonal inverse) --
    running it beside the substitutions instead of behind them must give a wrong inverse (or fail outright)."""
    bad = {k: list(v) for k, v in SIDE.items()}
    bad[(2, "update")] = [job_B(1)]
    bad[(2, "subst")] = [job_C1(0), job_C2(0)]
    A = _spd(1)
    try:
        L, X = run_schedule(A, bad)
    except KeyError:
        return                                          # X_22 does not exist yet: caught
    assert np.max(np.abs(X @ L - np.eye(4 * PB))) > 1e-6
