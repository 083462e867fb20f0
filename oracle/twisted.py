"""
twisted.py -- TEST INFRASTRUCTURE. numpy statement of the twisted ("burn at both ends") block-Thomas elimination and
solve that gb_btddod_full_invert_twisted_batch / gb_btddod_full_solve_inv_batch implement on the device
(include/griffon_b200.h; spitfire_b200/csrc/gb_btinv.cu, gb_btddod.cu). An extension of the reference's block-Thomas path
(btddod_matrix_kernels.cpp:19-119: same recurrences D'_i = D_i - L_i diag(sup_{i-1}), L_i = diag(sub_{i-1}) D'_{i-1}^{-1},
run from both ends of the system and joined in the middle block), so it has no reference counterpart to be pinned on:
it is pinned on the dense LAPACK solve of the assembled matrix, the reference's own check of its block solver
(tests/griffon/test_btddod.py). The product (spitfire_b200/) never imports this module.

Storage (BTDDOD, flamelet_kernels.cpp:31-90): nb column-major bs x bs diagonal blocks, then (nb-1)*bs sub-diagonal
scalars, then (nb-1)*bs super-diagonal scalars. sub_i couples block row i+1 to x_i, sup_i block row i to x_{i+1}.
"""
import numpy as np

MAGIC = 2.718281828459045e-300  # BT_TWIST_MAGIC (gb_kernels.cuh)


def _views(A, nb, bs):
    D = [A[i * bs * bs:(i + 1) * bs * bs].reshape(bs, bs).T for i in range(nb)]  # column-major blocks as matrices
    off = nb * bs * bs
    sub = A[off:off + (nb - 1) * bs].reshape(nb - 1, bs)
    sup = A[off + (nb - 1) * bs:off + 2 * (nb - 1) * bs].reshape(nb - 1, bs)
    return D, sub, sup


def twisted_invert(A, nb, bs):
    """(l_values, dinv, m) in the layout of the device kernel: l_values[i] = L_i for 1 <= i <= m, U_{i-1} for i > m, block 0
    holds the tag {m, MAGIC}; dinv[i] = inverse of D'_i (i < m), D*_m, D''_i (i > m); blocks column-major"""
    D, sub, sup = _views(np.asarray(A, dtype=np.float64), nb, bs)
    m = (nb - 1) // 2
    L = [np.zeros((bs, bs)) for _ in range(nb)]
    Di = [None] * nb
    # downwards from block 0 (btddod_matrix_kernels.cpp:48-75)
    Di[0] = np.linalg.inv(D[0])
    for i in range(1, m):
        L[i] = sub[i - 1][:, None] * Di[i - 1]
        Di[i] = np.linalg.inv(D[i] - L[i] * sup[i - 1][None, :])
    # upwards from block nb-1, the roles of the off-diagonals exchanged
    Di[nb - 1] = np.linalg.inv(D[nb - 1])
    for i in range(nb - 2, m, -1):
        U = sup[i][:, None] * Di[i + 1]
        L[i + 1] = U
        Di[i] = np.linalg.inv(D[i] - U * sub[i][None, :])
    # the meeting block
    Dm = D[m].copy()
    if m > 0:
        L[m] = sub[m - 1][:, None] * Di[m - 1]
        Dm = Dm - L[m] * sup[m - 1][None, :]
    if m < nb - 1:
        L[m + 1] = sup[m][:, None] * Di[m + 1]
        Dm = Dm - L[m + 1] * sub[m][None, :]
    Di[m] = np.linalg.inv(Dm)
    L[0][0, 0], L[0][1, 0] = float(m), MAGIC
    pack = lambda blocks: np.concatenate([b.T.ravel() for b in blocks])
    return pack(L), pack(Di), m


def twisted_solve(A, l_values, dinv, rhs, nb, bs):
    """the two-sided sweep of k_btddod_solve_inv for factors of twisted_invert"""
    _, sub, sup = _views(np.asarray(A, dtype=np.float64), nb, bs)
    blk = lambda a, i: a[i * bs * bs:(i + 1) * bs * bs].reshape(bs, bs).T
    assert l_values[1] == MAGIC
    m = int(l_values[0])
    b = np.asarray(rhs, dtype=np.float64).reshape(nb, bs)
    y, z, x = b.copy(), b.copy(), np.zeros((nb, bs))
    for i in range(1, m + 1):                       # y_i = b_i - L_i y_{i-1}
        y[i] = b[i] - blk(l_values, i) @ y[i - 1]
    for i in range(nb - 2, m, -1):                  # z_i = b_i - U_i z_{i+1}
        z[i] = b[i] - blk(l_values, i + 1) @ z[i + 1]
    t = blk(l_values, m + 1) @ z[m + 1] if m < nb - 1 else np.zeros(bs)
    x[m] = blk(dinv, m) @ (y[m] - t)
    for i in range(m - 1, -1, -1):                  # x_i = D'_i^{-1} (y_i - sup_i o x_{i+1})
        x[i] = blk(dinv, i) @ (y[i] - sup[i] * x[i + 1])
    for i in range(m + 1, nb):                      # x_i = D''_i^{-1} (z_i - sub_{i-1} o x_{i-1})
        x[i] = blk(dinv, i) @ (z[i] - sub[i - 1] * x[i - 1])
    return x.ravel()
