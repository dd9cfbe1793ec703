"""CPU tests of the oracle (oracle/mecano_oracle.c).  The oracle's parity with the JVM is UNPINNED (no JDK here, no
golden vectors in the reference); it is pinned instead against (i) the reference's own randomized invariants at the
reference's tolerances (ForwardDynamicsCalculatorTest.java:38-41, 767-1003), (ii) an independent textbook
formulation (featherstone_np.py), (iii) oracle-generated regression fixtures (tests/golden)."""
import os

import numpy as np
import pytest

import featherstone_np as fs
import oracle_lib as ol
import treedesc as td

# ForwardDynamicsCalculatorTest.java:38-41
ONE_DOF_JOINT_EPSILON = 8.0e-12
FLOATING_JOINT_EPSILON = 4.0e-11


def err(actual, expected):
    """ForwardDynamicsCalculatorTest.java:1099-1107: |delta| <= eps * max(1, ||expected||)"""
    return float(np.max(np.abs(actual - expected)) / max(1.0, np.linalg.norm(expected)))


def cases(rng):
    return [
        ("revolute chain 1", td.chain(rng, 1)),
        ("revolute chain 7 (A7)", td.chain(rng, 7)),
        ("prismatic chain 5", td.chain(rng, 5, prismatic_fraction=1.0)),
        ("one-dof tree 25", td.random_tree(rng, 25, prismatic_fraction=0.5)),
        ("floating only", td.chain(rng, 0, floating=True)),
        ("floating + chain 12", td.chain(rng, 12, floating=True)),
        ("floating + tree 40, rotated inertia poses", td.random_tree(rng, 40, floating=True, com_rotation=True, prismatic_fraction=0.2)),
        ("axis-aligned tree", td.random_tree(rng, 10, floating=True, axis_aligned=True)),
        ("H37", td.humanoid(rng, 2)),
        ("H36", td.humanoid(rng, 1)),
        # SphericalJoint / PlanarJoint (ForwardDynamicsCalculatorTest.testJointChain, :253-281: chains of every joint type)
        ("all joint types chain", td.mixed_chain(rng, [td.REVOLUTE, td.SPHERICAL, td.PRISMATIC, td.PLANAR, td.REVOLUTE, td.SPHERICAL, td.PLANAR, td.SIXDOF, td.REVOLUTE])),
        ("spherical chain 6", td.mixed_chain(rng, [td.SPHERICAL] * 6)),
        ("floating + mixed tree 30", td.mixed_tree(rng, 30, floating=True, com_rotation=True)),
        ("planar root + mixed tree 12", td.mixed_tree(rng, 12, weights=(1, 1, 0, 1, 2))),
    ]


N_CASES = 14


@pytest.mark.parametrize("idx", range(N_CASES))
def test_oracle_matches_featherstone(idx):
    rng = np.random.default_rng(100 + idx)
    name, t = cases(rng)[idx]
    g = (0.0, 0.0, -rng.uniform(1, 10))
    o, m = ol.Oracle(t, gravity=g), fs.Model(t)
    q, qd, qdd, tau = td.random_states(rng, t, 3)
    for s in range(3):
        fext = rng.uniform(-1, 1, size=(t.nb, 6))
        assert err(o.rnea(q[:, s], qd[:, s], qdd[:, s], fext), m.rnea(q[:, s], qd[:, s], qdd[:, s], g, fext)) < 1e-12, name
        assert err(o.aba(q[:, s], qd[:, s], tau[:, s], fext), m.aba(q[:, s], qd[:, s], tau[:, s], g, fext)) < 1e-11, name
        assert err(o.crba(q[:, s]), m.crba(q[:, s])) < 1e-12, name


@pytest.mark.parametrize("idx", range(N_CASES))
def test_reference_invariants(idx):
    """FD(ID(qdd)) == qdd and M qdd + ID(qdd = 0) == ID(qdd), with random external wrenches and gravity in [-10, -1]
    (ForwardDynamicsCalculatorTest.java:767-901, 904-1003), at the reference's tolerances."""
    rng = np.random.default_rng(200 + idx)
    name, t = cases(rng)[idx]
    eps = FLOATING_JOINT_EPSILON if (t.jtype >= td.SIXDOF).any() else 2.0 * ONE_DOF_JOINT_EPSILON
    o = ol.Oracle(t, gravity=(0.0, 0.0, -rng.uniform(1, 10)))
    n = 50
    q, qd, qdd, _ = td.random_states(rng, t, n)
    for fext in (None, np.ascontiguousarray(rng.uniform(-1, 1, size=(6 * t.nb, n)))):
        tau = o.rnea_batch(q, qd, qdd, fext)
        back = o.aba_batch(q, qd, tau, fext)
        bias = o.rnea_batch(q, qd, np.zeros_like(qdd), fext)
        M = o.crba_batch(q)
        for s in range(n):
            assert err(back[:, s], qdd[:, s]) < eps, name
            assert err(M[:, :, s] @ qdd[:, s] + bias[:, s], tau[:, s]) < eps, name
            assert np.allclose(M[:, :, s], M[:, :, s].T, rtol=0, atol=1e-12)
            assert np.linalg.eigvalsh(M[:, :, s]).min() > 0


@pytest.mark.parametrize("idx", range(N_CASES))
def test_reference_invariant_mixed_source_modes(idx):
    """ForwardDynamicsCalculatorTest.testJointMixedSourceModeGeneral (:389-488) and
    testJointAccelerationSourceWithZeroVelocityAcceleration (:283-386): with a random subset of joints switched to
    ACCELERATION_SOURCE (their efforts zeroed so that they cannot be used), forward dynamics returns the accelerations AND
    the efforts of inverse dynamics, to 1e-12 * max(1, max element).  The reference tests revolute chains; here every tree
    of the suite, floating joints included, with and without external wrenches."""
    rng = np.random.default_rng(900 + idx)
    name, t = cases(rng)[idx]
    o = ol.Oracle(t, gravity=(0.0, 0.0, -9.81))
    n = 20
    q, qd, qdd, _ = td.random_states(rng, t, n)
    for s in range(n):
        locked = np.zeros(t.nb, np.int32)
        locked[rng.permutation(t.nb)[: 1 if t.nb == 1 else rng.integers(1, t.nb)]] = 1
        qd_s, qdd_s = qd[:, s].copy(), qdd[:, s].copy()
        if s % 2:  # the zero-velocity / zero-acceleration variant
            for b in np.nonzero(locked)[0]:
                nd = td.NDOF[int(t.jtype[b])]
                qd_s[t.dof_off[b]:t.dof_off[b] + nd] = 0.0
                qdd_s[t.dof_off[b]:t.dof_off[b] + nd] = 0.0
        fext = None if s % 3 else rng.uniform(-1, 1, size=(t.nb, 6))
        tau = o.rnea(q[:, s], qd_s, qdd_s, fext)
        tau_in = tau.copy()
        for b in np.nonzero(locked)[0]:
            nd = td.NDOF[int(t.jtype[b])]
            tau_in[t.dof_off[b]:t.dof_off[b] + nd] = 0.0
        qdd_out, tau_out = o.aba_sources(q[:, s], qd_s, tau_in, qdd_s, locked, fext)
        scale = 4.0 if (t.jtype >= td.SIXDOF).any() else 1.0  # the reference's floating-joint tolerances are 5x its one-DoF ones
        assert np.max(np.abs(qdd_out - qdd_s)) <= scale * 1.0e-12 * max(1.0, np.max(np.abs(qdd_s))), name
        assert np.max(np.abs(tau_out - tau)) <= scale * 1.0e-12 * max(1.0, np.max(np.abs(tau))), name
        # nothing locked: the plain algorithm
        free_qdd, free_tau = o.aba_sources(q[:, s], qd_s, tau, qdd_s, np.zeros(t.nb, np.int32), fext)
        assert np.array_equal(free_qdd, o.aba(q[:, s], qd_s, tau, fext)) and np.array_equal(free_tau, tau)


@pytest.mark.parametrize("idx", range(N_CASES))
def test_centroidal_momentum_matrix_and_convective_term(idx):
    """CompositeRigidBodyMassMatrixCalculatorTest.java:62-82 compares getCentroidalMomentumMatrix() / getCentroidalConvectiveTerm()
    with CentroidalMomentumRateCalculator at 1e-10 (EPSILON); the same two facts against plain sums over the bodies
    (featherstone_np.Model.centroidal): A qd = momentum, A qdd + convective term = momentum rate, in the root frame and in the
    centre-of-mass frame."""
    rng = np.random.default_rng(1100 + idx)
    name, t = cases(rng)[idx]
    o, m = ol.Oracle(t), fs.Model(t)
    q, qd, qdd, _ = td.random_states(rng, t, 4)
    for s in range(4):
        h, hdot, com, mass = m.centroidal(q[:, s], qd[:, s], qdd[:, s])
        M, A, c, mtot = o.crba_centroidal(q[:, s], 0)
        assert np.array_equal(M, o.crba(q[:, s]))
        assert abs(mtot - mass) < 1e-12 * mass and np.max(np.abs(c - com)) < 1e-12, name
        b = o.centroidal_convective_term(q[:, s], qd[:, s], 0)
        assert err(A @ qd[:, s], h) < 1e-10, name
        assert err(A @ qdd[:, s] + b, hdot) < 1e-10, name
        # centre-of-mass frame: same axes, moments taken about the CoM
        shift = lambda w: np.concatenate([w[:3] - np.cross(com, w[3:]), w[3:]])
        _, Ac, _, _ = o.crba_centroidal(q[:, s], 1)
        bc = o.centroidal_convective_term(q[:, s], qd[:, s], 1)
        assert err(Ac @ qd[:, s], shift(h)) < 1e-10, name
        assert err(Ac @ qdd[:, s] + bc, shift(hdot)) < 1e-10, name
        # linear momentum = mass * CoM velocity: the linear rows of A in the CoM frame do not depend on the frame origin
        assert np.array_equal(Ac[3:], A[3:])


@pytest.mark.parametrize("idx", range(N_CASES))
def test_coriolis_matrix(idx):
    """CompositeRigidBodyMassMatrixCalculatorTest.testCoriolisMatrix (:85-138): C(q, qd) qd equals the joint efforts of inverse
    dynamics without joint accelerations (zero gravity, the calculators' default), at 1e-11.  Beyond the reference's test:
    dM/dt - 2 C is skew-symmetric, i.e. dM/dt = C + C^T (finite difference along qd; one-DoF trees, where q' = qd)."""
    rng = np.random.default_rng(1300 + idx)
    name, t = cases(rng)[idx]
    o = ol.Oracle(t, gravity=(0.0, 0.0, 0.0))
    q, qd, _, _ = td.random_states(rng, t, 4)
    for s in range(4):
        M, C = o.coriolis(q[:, s], qd[:, s])
        assert np.array_equal(M, o.crba(q[:, s])), name
        want = o.rnea(q[:, s], qd[:, s], np.zeros(t.nv), flags=2)
        assert np.max(np.abs(C @ qd[:, s] - want)) < 1e-11 * max(1.0, np.max(np.abs(want))), name
        if not (t.jtype >= td.SIXDOF).any():  # q + h qd is the configuration after h only when every joint is one-DoF
            h = 1e-6
            Mp, Mm = o.crba(q[:, s] + h * qd[:, s]), o.crba(q[:, s] - h * qd[:, s])
            assert np.max(np.abs((Mp - Mm) / (2 * h) - (C + C.T))) < 1e-6 * max(1.0, np.max(np.abs(C))), name


def test_flags_match_zeroed_inputs():
    """setConsiderCoriolisAndCentrifugalForces(false) == zero velocities; setConsiderJointAccelerations(false) == zero
    accelerations (InverseDynamicsCalculator.java:882-915)."""
    rng = np.random.default_rng(5)
    t = td.random_tree(rng, 15, floating=True)
    o = ol.Oracle(t)
    q, qd, qdd, _ = td.random_states(rng, t, 4)
    z = np.zeros_like(qd)
    assert np.allclose(o.rnea_batch(q, qd, qdd, flags=1), o.rnea_batch(q, z, qdd), atol=1e-12)
    assert np.allclose(o.rnea_batch(q, qd, qdd, flags=2), o.rnea_batch(q, qd, z), atol=1e-12)


def test_body_accelerations_consistent():
    """Body accelerations of RNEA pass one (RigidBodyAccelerationProvider) reproduce gravity for a body at rest."""
    rng = np.random.default_rng(6)
    t = td.chain(rng, 4)
    o = ol.Oracle(t, gravity=(0, 0, -9.81))
    q = rng.uniform(-1, 1, t.nq)
    acc = o.body_accelerations(q, np.zeros(t.nv), np.zeros(t.nv))
    assert np.allclose(np.linalg.norm(acc[:, 3:], axis=1), 9.81, atol=1e-12)
    assert np.allclose(acc[:, :3], 0, atol=1e-14)


@pytest.mark.parametrize("idx", [1, 2, 3])
def test_body_accelerations_against_numerical_differentiation_of_the_twists(idx):
    """The Coriolis-aware change of frame of a spatial acceleration (SpatialAccelerationBasics.changeFrame, the a4 row of SURVEY 8)
    the way the reference tests it (SpatialAccelerationTest.testChangeFrameUsingNumericalDifferentiationVersusAnalytical,
    T/spatial/SpatialAccelerationTest.java:28-46): analytical against numerically differentiated twists.  The twist of every body,
    expressed in its own centre-of-mass frame, comes from the same propagation run at rest (at zero velocity the acceleration of
    a body is S qdd carried up the tree, so `qdd := qd` and no gravity gives the twist); its coordinates differentiated along
    q(t) = q + t qd + t^2 qdd / 2 are the spatial acceleration (v x v = 0).  One-DoF trees, where q(t) is exact."""
    rng = np.random.default_rng(700 + idx)
    name, t = cases(rng)[idx]
    o = ol.Oracle(t, gravity=(0.0, 0.0, 0.0))
    q, qd, qdd, _ = td.random_states(rng, t, 3)
    zero = np.zeros(t.nv)
    h = 1.0e-5
    for s in range(3):
        twist = lambda dt: o.body_accelerations(q[:, s] + dt * qd[:, s] + 0.5 * dt * dt * qdd[:, s], zero, qd[:, s] + dt * qdd[:, s])  # noqa: E731
        numerical = (twist(h) - twist(-h)) / (2.0 * h)
        analytical = o.body_accelerations(q[:, s], qd[:, s], qdd[:, s])
        assert np.max(np.abs(numerical - analytical)) < 2.0e-8 * max(1.0, np.max(np.abs(analytical))), name
        # and the flags: without the velocity terms the acceleration is the twist propagation itself
        assert np.allclose(o.body_accelerations(q[:, s], qd[:, s], qdd[:, s], flags=1), o.body_accelerations(q[:, s], zero, qdd[:, s]), atol=1e-13)


@pytest.mark.parametrize("idx", [5, 6, 8, 10, 12])
def test_body_accelerations_against_differentiated_twists_along_the_integrator(idx):
    """The same for trees with SixDoF / spherical / planar joints, where q(t) is not a polynomial: the configurations at t +- h come
    from the oracle's state integrator (MultiBodySystemStateIntegrator.doubleIntegrateFromAcceleration, second-order accurate), so
    the test also ties the integrator's multi-DoF updates to the acceleration convention of the dynamics (the joint acceleration of
    a floating joint is a spatial acceleration in the frame after the joint)."""
    rng = np.random.default_rng(720 + idx)
    name, t = cases(rng)[idx]
    o = ol.Oracle(t, gravity=(0.0, 0.0, 0.0))
    q, qd, qdd, _ = td.random_states(rng, t, 2)
    zero = np.zeros(t.nv)
    h = 1.0e-5

    def twist(s, dt):
        q1, qd1, _ = o.integrate(dt, q[:, s], qd[:, s], qdd[:, s])
        return o.body_accelerations(q1, zero, qd1)

    for s in range(2):
        numerical = (twist(s, h) - twist(s, -h)) / (2.0 * h)
        analytical = o.body_accelerations(q[:, s], qd[:, s], qdd[:, s])
        assert np.max(np.abs(numerical - analytical)) < 5.0e-8 * max(1.0, np.max(np.abs(analytical))), name


@pytest.mark.parametrize("idx", [1, 3, 5, 6, 8, 10, 12, 13])
def test_power_balance_along_the_integrator(idx):
    """A law of mechanics that no recursion enters: along the motion, d/dt (qd^T M(q) qd / 2) = qd^T (tau - g(q)).  Left: the mass
    matrix (CRBA) at the configurations the state integrator reaches at t +- h under the accelerations of forward dynamics (ABA),
    differentiated numerically; right: the applied efforts minus the gravity efforts of inverse dynamics at rest (RNEA).  Holds
    only if the three calculators, the integrator and the conventions for multi-DoF joints (velocities of a floating joint = the
    twist in the frame after the joint, efforts = the wrench there) agree with one another and with physics."""
    rng = np.random.default_rng(740 + idx)
    name, t = cases(rng)[idx]
    o = ol.Oracle(t, gravity=(0.3, -0.2, -9.81))
    q, qd, _, tau = td.random_states(rng, t, 2)
    zero = np.zeros(t.nv)
    h = 1.0e-5
    for s in range(2):
        qdd = o.aba(q[:, s], qd[:, s], tau[:, s])

        def kinetic(dt):
            q1, qd1, _ = o.integrate(dt, q[:, s], qd[:, s], qdd)
            return 0.5 * qd1 @ o.crba(q1) @ qd1

        numerical = (kinetic(h) - kinetic(-h)) / (2.0 * h)
        power = qd[:, s] @ (tau[:, s] - o.rnea(q[:, s], zero, zero))
        assert abs(numerical - power) < 2.0e-8 * max(1.0, abs(power)), (name, numerical, power)


def test_golden_fixtures():
    """Regression vectors generated by tests/golden/make_golden.py from this oracle (NOT from the JVM)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz")
    data = np.load(path)
    for name in sorted({k.split("/")[0] for k in data.files}):
        t = td.TreeDesc(**{f: data["%s/%s" % (name, f)] for f in ("parent", "jtype", "axis", "off_R", "off_p", "com_R", "com_p", "J", "mass", "dof_off", "cfg_off")},
                        nb=int(data[name + "/dims"][0]), nv=int(data[name + "/dims"][1]), nq=int(data[name + "/dims"][2])).contiguous()
        o = ol.Oracle(t, gravity=tuple(data[name + "/gravity"]))
        q, qd, qdd, tau, fext = (data["%s/%s" % (name, f)] for f in ("q", "qd", "qdd", "tau", "fext"))
        assert np.allclose(o.rnea_batch(q, qd, qdd, fext), data[name + "/rnea"], rtol=0, atol=1e-11)
        assert np.allclose(o.aba_batch(q, qd, tau, fext), data[name + "/aba"], rtol=0, atol=1e-10)
        assert np.allclose(o.crba_batch(q), data[name + "/crba"], rtol=0, atol=1e-11)


def _planar_two_link(l1, lc1, lc2, m1, m2, i1, i2):
    """Textbook planar two-link arm (Spong, Hutchinson, Vidyasagar, Robot Modeling and Control, eq. 7.80-7.88): both joints
    revolute about +z, link 1 along its x axis with length l1, CoMs at lc1 / lc2 along x, gravity along -y."""
    eye = np.eye(3)
    return td.TreeDesc(
        nb=2, nv=2, nq=2, parent=np.array([-1, 0], np.int32), jtype=np.array([td.REVOLUTE, td.REVOLUTE], np.int32),
        axis=np.array([[0.0, 0.0, 1.0], [0.0, 0.0, 1.0]]), off_R=np.stack([eye, eye]), off_p=np.array([[0.0, 0.0, 0.0], [l1, 0.0, 0.0]]),
        com_R=np.stack([eye, eye]), com_p=np.array([[lc1, 0.0, 0.0], [lc2, 0.0, 0.0]]),
        J=np.stack([np.diag([0.01, 0.02, i1]), np.diag([0.03, 0.04, i2])]), mass=np.array([m1, m2]),
        dof_off=np.array([0, 1], np.int32), cfg_off=np.array([0, 1], np.int32)).contiguous()


def test_closed_form_planar_two_link_arm():
    """Physics pin that depends on neither the reference nor a second recursive formulation: the closed-form equations of motion
    of the planar two-link arm, M(q) qdd + C(q, qd) qd + g(q) = tau, against the oracle's RNEA, CRBA, Coriolis matrix and ABA."""
    rng = np.random.default_rng(77)
    l1, lc1, lc2, m1, m2, i1, i2, g = 0.9, 0.4, 0.35, 2.0, 1.5, 0.11, 0.07, 9.81
    t = _planar_two_link(l1, lc1, lc2, m1, m2, i1, i2)
    o = ol.Oracle(t, gravity=(0.0, -g, 0.0))
    o0 = ol.Oracle(t, gravity=(0.0, 0.0, 0.0))
    for _ in range(20):
        q, qd, qdd = rng.uniform(-np.pi, np.pi, 2), rng.uniform(-2, 2, 2), rng.uniform(-3, 3, 2)
        c2, s2 = np.cos(q[1]), np.sin(q[1])
        d11 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * c2) + i1 + i2
        d12 = m2 * (lc2 ** 2 + l1 * lc2 * c2) + i2
        d22 = m2 * lc2 ** 2 + i2
        M = np.array([[d11, d12], [d12, d22]])
        h = -m2 * l1 * lc2 * s2
        C = np.array([[h * qd[1], h * (qd[0] + qd[1])], [-h * qd[0], 0.0]])
        grav = np.array([(m1 * lc1 + m2 * l1) * g * np.cos(q[0]) + m2 * lc2 * g * np.cos(q[0] + q[1]), m2 * lc2 * g * np.cos(q[0] + q[1])])
        tau = M @ qdd + C @ qd + grav
        assert np.allclose(o.crba(q), M, rtol=0, atol=1e-13)
        assert np.allclose(o.rnea(q, qd, qdd), tau, rtol=0, atol=1e-12)
        assert np.allclose(o.aba(q, qd, tau), qdd, rtol=0, atol=1e-11)
        # the Coriolis matrix is not unique, its product with qd is (and Mecano's factorization also satisfies dM/dt = C + C^T)
        _, Co = o0.coriolis(q, qd)
        assert np.allclose(Co @ qd, C @ qd, rtol=0, atol=1e-12)
        Mdot = np.array([[2 * h * qd[1], h * qd[1]], [h * qd[1], 0.0]])
        assert np.allclose(Co + Co.T, Mdot, rtol=0, atol=1e-12)
        # momentum of the system about the base origin, z component = row 2 of the centroidal momentum matrix in the world frame
        _, A, com, mass = o0.crba_centroidal(q, 0)
        assert abs(mass - (m1 + m2)) < 1e-14
        p1 = lc1 * np.array([np.cos(q[0]), np.sin(q[0])])
        p2 = l1 * np.array([np.cos(q[0]), np.sin(q[0])]) + lc2 * np.array([np.cos(q[0] + q[1]), np.sin(q[0] + q[1])])
        assert np.allclose(com[:2], (m1 * p1 + m2 * p2) / (m1 + m2), atol=1e-13)
        v1 = lc1 * qd[0] * np.array([-np.sin(q[0]), np.cos(q[0])])
        v2 = l1 * qd[0] * np.array([-np.sin(q[0]), np.cos(q[0])]) + lc2 * (qd[0] + qd[1]) * np.array([-np.sin(q[0] + q[1]), np.cos(q[0] + q[1])])
        Lz = i1 * qd[0] + i2 * (qd[0] + qd[1]) + m1 * (p1[0] * v1[1] - p1[1] * v1[0]) + m2 * (p2[0] * v2[1] - p2[1] * v2[0])
        hq = A @ qd
        assert abs(hq[2] - Lz) < 1e-12 and np.allclose(hq[3:5], m1 * v1 + m2 * v2, atol=1e-12)


def test_golden_fixtures_of_the_next_rows():
    """Regression vectors generated by tests/golden/make_golden_next.py from this oracle (NOT from the JVM): RNEA by-products,
    joint source modes, centroidal quantities, Coriolis matrix, on the trees and states of oracle_golden.npz."""
    import importlib.util

    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_next", os.path.join(here, "make_golden_next.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    data = np.load(os.path.join(here, "oracle_golden_next.npz"))
    for name, t, g, s in gen.load_cases():
        got = gen.evaluate(t, g, s, data[name + "/accel_source"])
        for k, v in got.items():
            want = data["%s/%s" % (name, k)]
            assert v.shape == want.shape and np.allclose(v, want, rtol=0, atol=1e-10 * max(1.0, np.max(np.abs(want)))), (name, k)
