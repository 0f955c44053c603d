"""Pin the oracle's small dense algebra against independent numpy / scipy implementations.

The reference delegates these to Eigen / MRPT (not in its tree): LDLT solve (FrontEnd.cpp:642,
SegmentationBackground.cpp:168), SelfAdjointEigenSolver (:719), matrix exp / log (:766-769).
"""
import numpy as np
import pytest
import scipy.linalg


def hat(xi):
    M = np.zeros((4, 4))
    M[0, 1], M[1, 0] = -xi[5], xi[5]
    M[0, 2], M[2, 0] = xi[4], -xi[4]
    M[1, 2], M[2, 1] = -xi[3], xi[3]
    M[:3, 3] = xi[:3]
    return M  # FrontEnd.cpp:759-763


@pytest.mark.parametrize("scale", [1e-6, 1e-3, 0.02, 0.049, 0.051, 0.3, 1.5])
def test_se3_exp_matches_expm(oracle_mod, scale):
    rng = np.random.default_rng(int(scale * 1e6) + 1)
    for _ in range(20):
        xi = rng.standard_normal(6) * scale
        T = oracle_mod.se3_exp(xi)
        assert np.allclose(T, scipy.linalg.expm(hat(xi)), rtol=0, atol=1e-13)


@pytest.mark.parametrize("scale", [1e-6, 1e-3, 0.02, 0.049, 0.051, 0.3, 1.5])
def test_se3_log_inverts_exp(oracle_mod, scale):
    rng = np.random.default_rng(int(scale * 1e6) + 2)
    for _ in range(20):
        xi = rng.standard_normal(6) * scale
        if np.linalg.norm(xi[3:]) > 2.8:  # the principal logarithm needs |omega| < pi
            xi[3:] *= 2.8 / np.linalg.norm(xi[3:])
        T = scipy.linalg.expm(hat(xi))
        got = oracle_mod.se3_log(T)
        assert np.allclose(got, xi, rtol=0, atol=1e-12)
        L = scipy.linalg.logm(T).real  # twist extraction as FrontEnd.cpp:769-771
        ref = np.array([L[0, 3], L[1, 3], L[2, 3], -L[1, 2], L[0, 2], -L[0, 1]])
        assert np.allclose(got, ref, rtol=0, atol=1e-10)


def test_se3_exp_identity(oracle_mod):
    assert np.array_equal(oracle_mod.se3_exp(np.zeros(6)), np.eye(4))
    assert np.array_equal(oracle_mod.se3_log(np.eye(4)), np.zeros(6))


@pytest.mark.parametrize("n", [6, 24])
def test_ldlt_solve_matches_numpy(oracle_mod, n):
    rng = np.random.default_rng(n)
    for _ in range(10):
        M = rng.standard_normal((n + 5, n))
        A = M.T @ M + 1e-3 * np.eye(n)
        b = rng.standard_normal(n)
        x, nz = oracle_mod.ldlt_solve(A, b)
        assert nz == 0
        assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)


def test_ldlt_zero_pivot_gives_zero_component(oracle_mod):
    """an empty, unconnected cluster is a zero row/column -> that b_l is 0."""
    rng = np.random.default_rng(5)
    M = rng.standard_normal((30, 24))
    A = M.T @ M
    for dead in (0, 7, 23):
        A2 = A.copy()
        A2[dead, :] = 0
        A2[:, dead] = 0
        b = rng.standard_normal(24)
        b[dead] = 0
        x, nz = oracle_mod.ldlt_solve(A2, b)
        assert nz == 1 and x[dead] == 0.0
        keep = [i for i in range(24) if i != dead]
        assert np.allclose(x[keep], np.linalg.solve(A2[np.ix_(keep, keep)], b[keep]), rtol=1e-9, atol=1e-12)


def test_jacobi_matches_eigh(oracle_mod):
    rng = np.random.default_rng(9)
    for k in range(10):
        M = rng.standard_normal((6, 6)) * (10.0 ** rng.uniform(-6, 0))
        A = M @ M.T
        ev, V = oracle_mod.jacobi_eig6(A)
        assert np.allclose(np.sort(ev), np.linalg.eigvalsh(A), rtol=1e-10, atol=1e-14 * np.abs(A).max())
        assert np.allclose(V @ np.diag(ev) @ V.T, A, rtol=0, atol=1e-12 * np.abs(A).max())
        assert np.allclose(V.T @ V, np.eye(6), atol=1e-13)
