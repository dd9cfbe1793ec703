"""Unit tests of the kernels' spatial-algebra primitives (mecano_b200/csrc/spatial.cuh, compiled for the host by tests/emu)
against dense 6 x 6 numpy, the way the reference tests its own unrolled helpers against EJML:

  * ArticulatedBodyInertiaTest.testApplyTransform (T/algorithms/ArticulatedBodyInertiaTest.java:21-64, 1e-12): the congruence
    transform of an articulated-body inertia against the dense product -- here abi_to_parent<0 / 1 / 2>;
  * SpatialInertiaBasicsTest / MecanoToolsTest.testTranslateMomentOfInertia (1e-12): rigid-body inertia change of frame --
    rbi_to_parent (row a7);
  * MecanoToolsTest.testComputeDynamicWrench (T/tools/MecanoToolsTest.java:463-560): the Newton-Euler wrench computed from the
    centre-of-mass quantities equals I a + v x* (I v) in any frame -- newton_euler (row a5);
  * ForwardDynamicsCalculatorTest.testAddEquals / testMult / testMultTransA / testMultAdd (:491-582): the unrolled 6 x 6
    helpers vs dense -- mul, abi_add_rbi, abi_downdate, abi_solve (row a9's building blocks);
  * the motion / force transforms and cross products (row a3) against the Pluecker matrices.

No GPU: the product library is not involved, only the header the kernels inline."""
import ctypes

import numpy as np
import pytest

import emu_lib as el

EPS = 1.0e-12  # ArticulatedBodyInertiaTest / MecanoToolsTest


def spatial(op, *parts, n_out):
    lib = el.lib()
    lib.emu_spatial.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    buf = np.ascontiguousarray(np.concatenate([np.ravel(p) for p in parts]), dtype=np.float64)
    out = np.full(64, np.nan)
    got = lib.emu_spatial(op, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    assert got == n_out, (op, got)
    return out[:n_out].copy()


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def random_rotation(rng):
    q = rng.standard_normal(4)
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def motion_matrix(R, p):
    """Pluecker matrix taking motion vectors (angular, linear) from the parent frame to the child frame whose pose in the parent
    is (R, p)."""
    X = np.zeros((6, 6))
    X[:3, :3] = R.T
    X[3:, 3:] = R.T
    X[3:, :3] = -R.T @ skew(p)
    return X


def sym6(s):
    return np.array([[s[0], s[1], s[2]], [s[1], s[3], s[4]], [s[2], s[4], s[5]]])


def to_s6(S):
    return np.array([S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]])


def rbi_dense(flat):
    """[[I, h~], [h~^T, m 1]] from (I, h = m c, m) about the frame origin."""
    I, h, m = sym6(flat[:6]), flat[6:9], flat[9]
    D = np.zeros((6, 6))
    D[:3, :3] = I
    D[:3, 3:] = skew(h)
    D[3:, :3] = skew(h).T
    D[3:, 3:] = m * np.eye(3)
    return D


def rbi_flat(rng):
    """A physical rigid-body inertia about the frame origin: random principal moments, orientation, centre of mass."""
    m = rng.uniform(0.5, 5.0)
    c = rng.uniform(-1, 1, 3)
    Rc = random_rotation(rng)
    J = Rc @ np.diag(rng.uniform(0.1, 2.0, 3)) @ Rc.T
    Io = J + m * skew(c) @ skew(c).T
    return np.concatenate([to_s6(Io), m * c, [m]]), (J, c, m)


def abi_dense(flat):
    A, C, L = sym6(flat[:6]), flat[6:15].reshape(3, 3), sym6(flat[15:21])
    D = np.zeros((6, 6))
    D[:3, :3] = A
    D[:3, 3:] = C
    D[3:, :3] = C.T
    D[3:, 3:] = L
    return D


def abi_flat_from_dense(D):
    return np.concatenate([to_s6(D[:3, :3]), D[:3, 3:].ravel(), to_s6(D[3:, 3:])])


def random_abi(rng):
    """Symmetric positive definite 6 x 6: a rigid-body inertia plus a few transformed ones (what pass two accumulates)."""
    D = rbi_dense(rbi_flat(rng)[0])
    for _ in range(3):
        X = motion_matrix(random_rotation(rng), rng.uniform(-1, 1, 3))
        D = D + X.T @ rbi_dense(rbi_flat(rng)[0]) @ X
    return D


def close(a, b, eps=EPS):
    return np.max(np.abs(a - b)) <= eps * max(1.0, np.max(np.abs(b)))


@pytest.mark.parametrize("seed", range(20))
def test_motion_and_force_transforms_and_cross_products(seed):
    rng = np.random.default_rng(seed)
    R, p = random_rotation(rng), rng.uniform(-2, 2, 3)
    X = motion_matrix(R, p)
    m, f, v = rng.standard_normal(6), rng.standard_normal(6), rng.standard_normal(6)
    mc = spatial(0, R, p, m, n_out=6)
    fp = spatial(1, R, p, f, n_out=6)
    assert close(mc, X @ m)
    assert close(fp, X.T @ f)  # the dual transform, in the opposite direction
    # power is frame-invariant: (f in the child) . (m brought to the child) == (f brought to the parent) . (m in the parent)
    assert abs(f @ mc - fp @ m) < EPS * max(1.0, abs(fp @ m))
    crm = np.zeros((6, 6))
    crm[:3, :3] = skew(v[:3])
    crm[3:, 3:] = skew(v[:3])
    crm[3:, :3] = skew(v[3:])
    assert close(spatial(8, v, m, n_out=6), crm @ m)
    assert close(spatial(9, v, f, n_out=6), -crm.T @ f)
    add = rng.standard_normal(6)
    assert close(spatial(13, v, f, add, n_out=6), add - crm.T @ f)


@pytest.mark.parametrize("seed", range(20))
def test_rigid_body_inertia_change_of_frame_and_newton_euler_wrench(seed):
    rng = np.random.default_rng(100 + seed)
    flat, (J, c, m) = rbi_flat(rng)
    I = rbi_dense(flat)
    R, p = random_rotation(rng), rng.uniform(-2, 2, 3)
    X = motion_matrix(R, p)
    got = spatial(2, R, p, flat, n_out=10)
    assert close(rbi_dense(got), X.T @ I @ X)
    assert got[9] == flat[9]
    mv = rng.standard_normal(6)
    assert close(spatial(7, flat, mv, n_out=6), I @ mv)
    # Newton-Euler from the centre-of-mass quantities (J about the CoM, c, m) == I a + v x* (I v) about the origin
    v, a = rng.standard_normal(6), rng.standard_normal(6)
    crf = np.zeros((6, 6))
    crf[:3, :3] = skew(v[:3])
    crf[3:, 3:] = skew(v[:3])
    crf[:3, 3:] = skew(v[3:])
    want = I @ a + crf @ (I @ v)
    assert close(spatial(4, to_s6(J), c, [m], v, a, n_out=6), want)
    # ... and the wrench is the same physical quantity in another frame (MecanoToolsTest.testComputeDynamicWrench): evaluate it
    # in the parent frame from the transformed inertia / twist / acceleration and bring the child-frame result up
    Xi = np.linalg.inv(X)
    Ip = X.T @ I @ X
    vp, ap = Xi @ v, Xi @ a
    crfp = np.zeros((6, 6))
    crfp[:3, :3] = skew(vp[:3])
    crfp[3:, 3:] = skew(vp[:3])
    crfp[:3, 3:] = skew(vp[3:])
    assert close(spatial(1, R, p, want, n_out=6), Ip @ ap + crfp @ (Ip @ vp), 1e-11)


@pytest.mark.parametrize("seed", range(20))
def test_articulated_inertia_congruence_downdate_and_solve(seed):
    rng = np.random.default_rng(200 + seed)
    IA = random_abi(rng)
    flat = abi_flat_from_dense(IA)
    R, p = random_rotation(rng), rng.uniform(-2, 2, 3)
    X = motion_matrix(R, p)
    # ArticulatedBodyInertia.applyTransform vs the dense congruence
    assert close(abi_dense(spatial(3, R, p, flat, n_out=21)), X.T @ IA @ X)
    mv = rng.standard_normal(6)
    assert close(spatial(6, flat, mv, n_out=6), IA @ mv)
    # 6 x 6 solve of the floating joint (EJML symmPosDef in the reference)
    b = rng.standard_normal(6)
    x = spatial(5, flat, b, n_out=6)
    assert np.max(np.abs(IA @ x - b)) < 1e-10 * max(1.0, np.max(np.abs(b)))
    assert close(x, np.linalg.solve(IA, b), 1e-9)
    # I^A + rigid-body inertia
    rflat, _ = rbi_flat(rng)
    assert close(abi_dense(spatial(14, flat, rflat, n_out=21)), IA + rbi_dense(rflat))
    # rank-one downdate along a joint axis, general and zero-structured forms; then the congruence that skips the zero row / column
    for axis, op in ((2, 11), (5, 12)):  # revolute about z (angular z), prismatic along z (linear z)
        U = IA[:, axis].copy()
        g = U / U[axis]
        want = IA - np.outer(U, U) / U[axis]
        gen = abi_dense(spatial(10, flat, U, g, n_out=21))
        assert close(gen, want)
        both = spatial(op, R, p, flat, U, g, n_out=42)
        down, up = abi_dense(both[:21]), abi_dense(both[21:])
        assert close(down, want)
        assert not down[axis, :].any() and not down[:, axis].any(), "the joint's row / column is written as exact zeros"
        assert close(up, X.T @ want @ X)


def test_branch_free_sincos_accuracy():
    """The kernels' own sin/cos (jointmath.cuh: Cody-Waite reduction + fdlibm polynomials, no branch) against libm over the fast
    range |x| < 1e5, at quadrant boundaries, and -- through the large-angle safeguards of the ops -- far beyond it.  Mecano calls
    Math.sin / Math.cos (1 ulp); the bound here is an absolute 3e-16, i.e. about one ulp of a value near one."""
    rng = np.random.default_rng(0)
    k = np.arange(-2000, 2001, dtype=np.float64)
    x = np.concatenate([
        rng.uniform(-np.pi, np.pi, 200000), rng.uniform(-1.0e5, 1.0e5, 200000), rng.uniform(-1.0e-3, 1.0e-3, 20000),
        k * (np.pi / 2), np.nextafter(k * (np.pi / 2), np.inf), np.nextafter(k * (np.pi / 4), -np.inf),
        [0.0, -0.0, 1.0e5 - 1.0e-9, -(1.0e5 - 1.0e-9), 99999.99999, 7.0e-310],
        rng.uniform(-1.0e9, 1.0e9, 20000), [1.0e5, -1.0e5, 1.0e15, -3.0e18, 1.0e300]])
    s, c = np.empty_like(x), np.empty_like(x)
    lib = el.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    lib.emu_sincos.argtypes = [ctypes.c_long, dp, dp, dp]
    lib.emu_sincos.restype = None
    lib.emu_sincos(len(x), x.ctypes.data_as(dp), s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    assert not (np.isnan(s).any() or np.isnan(c).any()), "the two large-angle forms disagree somewhere"
    assert np.max(np.abs(s - np.sin(x))) < 3.0e-16 and np.max(np.abs(c - np.cos(x))) < 3.0e-16
    assert np.max(np.abs(s * s + c * c - 1.0)) < 5.0e-16
    # relative accuracy where sin is small (the linear term must survive): one-DoF joints near their zero position
    small = np.abs(x) < 1.0e-3
    assert np.max(np.abs(s[small] - np.sin(x[small])) / np.maximum(np.abs(x[small]), 1e-300)) < 3.0e-16
