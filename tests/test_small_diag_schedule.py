"""CPU emulation of the shared-memory slot plan of ``potrf_diag_small_kernel`` (gumbi_b200/csrc/cholesky.cuh, SMALL branch of
potrf_diag_body): the inverse of the 4 x 4-block lower-triangular factor is assembled block column by block column inside the
slots of the factor's diagonal sub-blocks (plus the slot of L_10), because the small-footprint kernel has no room for the ten
transposed inverse blocks of the full-size kernel.  The emulation performs the kernel's pd_unit calls in order on aliased numpy
buffers -- a slot overwritten before its last use, or a wrong operand, gives a wrong inverse."""
import numpy as np
import pytest


def unit(op, C, A, Bt, A2=None, Bt2=None):
    """pd_unit<OP, TR=true>: P = A Bt^T (+ A2 Bt2^T); C holds its block TRANSPOSED.  OP 1: C^T = P, 2: C^T += P, 3: C^T = -P."""
    P = A @ Bt.T
    if A2 is not None:
        P = P + A2 @ Bt2.T
    if op == 1:
        C[...] = P.T
    elif op == 2:
        C[...] = C + P.T
    elif op == 3:
        C[...] = -P.T
    return C.T.copy()                                            # what the G argument writes to global memory (row-major block)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_slot_plan_assembles_the_inverse(seed):
    b = 3
    rng = np.random.default_rng(seed)
    Lfull = np.tril(rng.standard_normal((4 * b, 4 * b))) + 4 * np.eye(4 * b)
    blk = lambda M, i, j: M[i * b:(i + 1) * b, j * b:(j + 1) * b]
    L = {(i, j): blk(Lfull, i, j).copy() for i in range(4) for j in range(i + 1)}      # Lb: ten sub-block slots
    Xd = [np.linalg.inv(L[(p, p)]) for p in range(4)]                                  # inv32 results (not transposed)
    G = np.zeros_like(Lfull)                                                           # Dinv in global memory
    for p in range(4):
        blk(G, p, p)[...] = Xd[p]
    S0, S1, S2, S3, S4 = L[(0, 0)], L[(1, 1)], L[(2, 2)], L[(3, 3)], L[(1, 0)]         # aliases, exactly as in the kernel
    # --- block column 0
    S0[...] = Xd[0].T
    for dst, i in ((S1, 1), (S2, 2), (S3, 3)):
        unit(1, dst, L[(i, 0)], S0)                                                    # T_i0 = L_i0 X_00
    blk(G, 1, 0)[...] = unit(3, S1, Xd[1], S1)                                         # X_10
    unit(2, S2, L[(2, 1)], S1)                                                         # T_20 += L_21 X_10
    blk(G, 2, 0)[...] = unit(3, S2, Xd[2], S2)                                         # X_20
    unit(2, S3, L[(3, 1)], S1, L[(3, 2)], S2)                                          # T_30 += L_31 X_10 + L_32 X_20
    blk(G, 3, 0)[...] = unit(3, S3, Xd[3], S3)                                         # X_30
    # --- block columns 1 and 2
    S0[...] = Xd[1].T
    S1[...] = Xd[2].T
    unit(1, S2, L[(2, 1)], S0)                                                         # T_21 = L_21 X_11
    unit(1, S3, L[(3, 1)], S0)                                                         # T_31 = L_31 X_11
    unit(1, S4, L[(3, 2)], S1)                                                         # T_32 = L_32 X_22
    blk(G, 2, 1)[...] = unit(3, S2, Xd[2], S2)                                         # X_21
    blk(G, 3, 2)[...] = unit(3, S4, Xd[3], S4)                                         # X_32
    unit(2, S3, L[(3, 2)], S2)                                                         # T_31 += L_32 X_21
    blk(G, 3, 1)[...] = unit(3, S3, Xd[3], S3)                                         # X_31
    np.testing.assert_allclose(G, np.linalg.inv(Lfull), rtol=1e-10, atol=1e-12)
