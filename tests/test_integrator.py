"""State integrator: MultiBodySystemStateIntegrator.doubleIntegrateFromAcceleration.

CPU: the oracle restatement (oracle/mecano_oracle.c: mo_integrate) is pinned by the reference's own tests for this
function -- the closed-form ballistic trajectory of a free SixDoF body under gravity, 1000 steps at 1e-12
(MultiBodySystemStateIntegratorTest.java:200-270, EPSILON :37), the one-DoF formula (:710-733) and a finite-difference
check of the SixDoF twist (:41-197).  GPU (-m gpu): the CUDA kernel through the C ABI against the oracle on random trees,
and the same ballistic roll-out run entirely on the device (mecano_b200_aba + mecano_b200_integrate, 1000 steps)."""
import numpy as np
import pytest

import oracle_lib as ol
import treedesc as td

EPSILON = 1.0e-12  # MultiBodySystemStateIntegratorTest.java:37


def quat_rot(q):
    x, y, z, s = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - s * z), 2 * (x * z + s * y)],
                     [2 * (x * y + s * z), 1 - 2 * (x * x + z * z), 2 * (y * z - s * x)],
                     [2 * (x * z - s * y), 2 * (y * z + s * x), 1 - 2 * (x * x + y * y)]])


def free_body():
    """new SixDoFJoint("joint", root) + unit-inertia, unit-mass body at the joint origin (:209-211)."""
    rng = np.random.default_rng(0)
    t = td.chain(rng, 0, floating=True)
    t.J[0] = np.eye(3)
    t.mass[0] = 1.0
    t.com_p[0] = 0.0
    t.com_R[0] = np.eye(3)
    return t.contiguous()


def test_oracle_one_dof_formula():
    rng = np.random.default_rng(1)
    t = td.random_tree(rng, 12, prismatic_fraction=0.5)
    o = ol.Oracle(t)
    q, qd, qdd, _ = td.random_states(rng, t, 1)
    dt = 1.0e-3
    q1, qd1, qdd1 = o.integrate(dt, q[:, 0], qd[:, 0], qdd[:, 0])
    np.testing.assert_allclose(q1, q[:, 0] + dt * qd[:, 0] + 0.5 * dt * dt * qdd[:, 0], rtol=0, atol=1e-15)
    np.testing.assert_allclose(qd1, qd[:, 0] + dt * qdd[:, 0], rtol=0, atol=1e-15)
    np.testing.assert_array_equal(qdd1, qdd[:, 0])


def test_oracle_ballistic_matches_reference_test():
    """testSixDoFJointBallistic (:200-270): ForwardDynamicsCalculator + integrator, 1000 steps, closed form at 1e-12."""
    rng = np.random.default_rng(4366346)
    t = free_body()
    for it in range(5):
        gravity = rng.uniform(-100.0, -10.0)
        dt = rng.uniform(1.0e-5, 1.0e-3)
        o = ol.Oracle(t, gravity=(0.0, 0.0, gravity))
        q, qd, _, _ = td.random_states(rng, t, 1)
        q, qd = q[:, 0].copy(), qd[:, 0].copy()
        p0 = q[4:7].copy()
        w0 = qd[0:3].copy()
        v_world0 = quat_rot(q[0:4]) @ qd[3:6]
        for j in range(1000):
            tt = (j + 1.0) * dt
            qdd = o.aba(q, qd, np.zeros(6))
            q, qd, qdd = o.integrate(dt, q, qd, qdd)
            exp_v = v_world0 + np.array([0.0, 0.0, gravity * tt])
            exp_p = p0 + np.array([v_world0[0] * tt, v_world0[1] * tt, v_world0[2] * tt + 0.5 * gravity * tt * tt])
            R = quat_rot(q[0:4])
            assert np.max(np.abs(q[4:7] - exp_p)) < EPSILON * max(1.0, np.max(np.abs(exp_p))) * 10, (it, j)
            assert np.max(np.abs(R @ qd[3:6] - exp_v)) < EPSILON * max(1.0, np.max(np.abs(exp_v))) * 10, (it, j)
            assert np.max(np.abs(qd[0:3] - w0)) < EPSILON  # unit inertia: the angular velocity stays constant (:260)
            # linear acceleration of the body origin in world = gravity (:264-267)
            a_origin = qdd[3:6] + np.cross(qd[0:3], qd[3:6])
            assert np.max(np.abs(R @ a_origin - np.array([0.0, 0.0, gravity]))) < 1e-10
            assert np.max(np.abs(qdd[0:3])) < EPSILON


def test_oracle_sixdof_twist_against_finite_difference():
    """testSixDoFJointAgainstFiniteDifference (:41-197), constant-velocity case: the pose change over dt, differentiated,
    gives back the body-frame twist."""
    rng = np.random.default_rng(3)
    t = free_body()
    o = ol.Oracle(t)
    for _ in range(20):
        q, qd, _, _ = td.random_states(rng, t, 1)
        q, qd = q[:, 0], qd[:, 0]
        dt = 1.0e-6
        q1, qd1, _ = o.integrate(dt, q, qd, np.zeros(6))
        R0, R1 = quat_rot(q[0:4]), quat_rot(q1[0:4])
        dR = R0.T @ R1
        w_fd = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / (2 * dt)
        v_fd = R0.T @ (q1[4:7] - q[4:7]) / dt
        # first-order finite difference: error O(dt |w x v|)
        assert np.max(np.abs(w_fd - qd[0:3])) < 1e-8
        assert np.max(np.abs(v_fd - qd[3:6])) < 5e-6
        assert abs(np.linalg.norm(q1[0:4]) - 1.0) < 1e-14


def test_oracle_batch_equals_single():
    rng = np.random.default_rng(5)
    t = td.humanoid(rng, 2)
    o = ol.Oracle(t)
    q, qd, qdd, _ = td.random_states(rng, t, 7)
    Q, QD, QDD = o.integrate_batch(2.5e-3, q, qd, qdd)
    for s in range(7):
        a, b, c = o.integrate(2.5e-3, q[:, s], qd[:, s], qdd[:, s])
        np.testing.assert_array_equal(Q[:, s], a)
        np.testing.assert_array_equal(QD[:, s], b)
        np.testing.assert_array_equal(QDD[:, s], c)


def test_oracle_spherical_and_planar_joints_against_the_sixdof_update():
    """SphericalJoint (MultiBodySystemStateIntegrator.java:449-452, :575-594) and PlanarJoint (a FloatingJointBasics, :421-424 ->
    :503-560) integrate like a SixDoFJoint whose state is confined to the joint's subspace: the oracle's two branches against its
    SixDoF branch (pinned by the reference's ballistic test above) on the embedded state."""
    rng = np.random.default_rng(17)
    six = ol.Oracle(free_body())
    sph = ol.Oracle(td.mixed_chain(rng, [td.SPHERICAL]))
    pla = ol.Oracle(td.mixed_chain(rng, [td.PLANAR]))
    dt = 3.0e-3
    for _ in range(20):
        quat = rng.normal(size=4)
        quat /= np.linalg.norm(quat)
        w, wd = rng.uniform(-2, 2, size=3), rng.uniform(-2, 2, size=3)
        # spherical: SixDoF at the origin with no linear velocity / acceleration
        q6 = np.concatenate([quat, np.zeros(3)])
        a, b, c = six.integrate(dt, q6, np.concatenate([w, np.zeros(3)]), np.concatenate([wd, np.zeros(3)]))
        qs, ws, wds = sph.integrate(dt, quat.copy(), w.copy(), wd.copy())
        assert np.max(np.abs(qs - a[:4])) < 1e-15 and np.max(np.abs(ws - b[:3])) < 1e-15
        np.testing.assert_array_equal(wds, wd)
        assert np.max(np.abs(a[4:])) == 0.0
        # planar: pitch about y, position / velocity / acceleration in the x-z plane
        th, x, z = rng.uniform(-np.pi, np.pi), rng.uniform(-1, 1), rng.uniform(-1, 1)
        v3, a3 = rng.uniform(-2, 2, size=3), rng.uniform(-2, 2, size=3)  # (w_y, v_x, v_z)
        q6 = np.array([0.0, np.sin(th / 2), 0.0, np.cos(th / 2), x, 0.0, z])
        a, b, c = six.integrate(dt, q6, np.array([0, v3[0], 0, v3[1], 0, v3[2]]), np.array([0, a3[0], 0, a3[1], 0, a3[2]]))
        qp, vp, ap = pla.integrate(dt, np.array([th, x, z]), v3.copy(), a3.copy())
        R = quat_rot(a[:4])
        assert abs(np.arctan2(R[0, 2], R[0, 0]) - np.arctan2(np.sin(qp[0]), np.cos(qp[0]))) < 1e-14  # R_y(pitch): xz = sin, xx = cos
        assert np.max(np.abs(qp[1:] - a[[4, 6]])) < 1e-14 and abs(a[5]) < 1e-16
        assert np.max(np.abs(vp - b[[1, 3, 5]])) < 1e-14 and np.max(np.abs(b[[0, 2, 4]])) < 1e-16
        assert np.max(np.abs(ap - c[[1, 3, 5]])) < 1e-14 and np.max(np.abs(c[[0, 2, 4]])) < 1e-16


# ------------------------------------------------------------------------------------------------ GPU
def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))


GPU_CASES = [
    dict(kind="chain", seed=1, n_joints=7),
    dict(kind="tree", seed=4, n_joints=30, prismatic=0.4),
    dict(kind="tree", seed=6, n_joints=50, floating=True, prismatic=0.2),
    dict(kind="humanoid", seed=7, n_joints=2),
    dict(kind="spherical", seed=11, n_joints=3),
    dict(kind="planar", seed=12, n_joints=3),
    dict(kind="jtree", seed=14, n_joints=25),
]


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(len(GPU_CASES)))
@pytest.mark.parametrize("n,pad", [(1, 0), (257, 3), (4096, 0)])
def test_gpu_integrator_matches_oracle(idx, n, pad):
    import torch

    import mecano_b200 as mb
    from test_gpu_parity import build

    s, t = build(**GPU_CASES[idx])
    rng = np.random.default_rng(50 + idx)
    q, qd, qdd, _ = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    dt = 2.0e-3
    ref = ol.Oracle(t).integrate_batch(dt, q, qd, qdd)
    integ = mb.MultiBodySystemStateIntegrator(s, dt)
    dev = torch.device("cuda:0")
    # leading dimension n + pad: odd strides take the scalar path, even ones the 128-bit path
    bufs = []
    for x in (q, qd, qdd):
        b = torch.full((x.shape[0], n + pad), float("nan"), dtype=torch.float64, device=dev)
        b[:, :n] = torch.from_numpy(x).to(dev)
        bufs.append(b)
    integ.doubleIntegrateFromAcceleration(bufs[0][:, :n], bufs[1][:, :n], bufs[2][:, :n])
    for got, exp, name in zip(bufs, ref, ("q", "qd", "qdd")):
        assert _rel(got[:, :n].cpu().numpy(), exp) <= 1e-12, name
        if pad:
            assert torch.isnan(got[:, n:]).all(), "padding columns were touched"
    # host entry point
    hq, hqd, hqdd = q.copy(), qd.copy(), qdd.copy()
    integ.doubleIntegrateFromAcceleration(hq, hqd, hqdd)
    for got, exp in zip((hq, hqd, hqdd), ref):
        assert _rel(got, exp) <= 1e-12


@pytest.mark.gpu
def test_gpu_ballistic_rollout_on_device():
    """testSixDoFJointBallistic (:200-270) for 4096 balls at once, state resident on the GPU for all 1000 steps."""
    import torch

    import mecano_b200 as mb

    e = mb.RigidBody("root")
    j = mb.SixDoFJoint("joint", e)
    mb.RigidBody("object", j, np.eye(3), 1.0, np.zeros(3))
    s = mb.MultiBodySystem.toMultiBodySystemBasics(e)
    n = 4096
    rng = np.random.default_rng(4366346)
    gravity, dt = -9.81, 1.0e-3
    q, qd, _, _ = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    dev = torch.device("cuda:0")
    tq, tqd = torch.from_numpy(q).to(dev), torch.from_numpy(qd).to(dev)
    tau = torch.zeros((6, n), dtype=torch.float64, device=dev)
    fd = mb.ForwardDynamicsCalculator(s).setKernelVariant("thread")
    fd.setGravitationalAcceleration(gravity)
    integ = mb.MultiBodySystemStateIntegrator(s, dt)
    p0 = q[4:7].copy()
    vw0 = np.stack([quat_rot(q[0:4, i]) @ qd[3:6, i] for i in range(n)], axis=1)
    steps = 1000
    for _ in range(steps):
        qdd = fd.compute(tq, tqd, tau)
        integ.doubleIntegrateFromAcceleration(tq, tqd, qdd)
    tt = steps * dt
    exp_p = p0 + vw0 * tt
    exp_p[2] += 0.5 * gravity * tt * tt
    got = tq.cpu().numpy()
    assert np.max(np.abs(got[4:7] - exp_p)) < 1e-10
    assert np.max(np.abs(tqd.cpu().numpy()[0:3] - qd[0:3])) < 1e-11
