"""Fixed joints and jointsToIgnore: host-side welding at flatten time.

A FixedJoint's successor, and every ignored subtree at its stored configuration, is folded into the nearest moving ancestor
(csrc/host/multibody.hpp: FlatTables::weld; InverseDynamicsCalculator.java:832-860, MultiBodySystemTools.java:32-64,
FixedJoint.java:40-62).  The check is physical and independent of that code: the oracle evaluates the FULL tree, with the
welded joints modelled as ordinary joints held still (q = q0, qd = qdd = 0); for the remaining joints
   RNEA:  tau_welded = tau_full[moving rows]
   CRBA:  M_welded   = M_full[moving, moving]
   ABA:   qdd_welded = M_sub^-1 (tau - bias_full[moving rows])       (the locked joints' constraint forces do no work)
CPU: the kernel source compiled for the host (tests/emu) runs on the welded tables; GPU (-m gpu): the CUDA kernels."""
import ctypes

import numpy as np
import pytest

import emu_lib as el
import oracle_lib as ol
import treedesc as td

G = (0.3, -0.2, -9.81)


def _rand_transform(rng):
    import mecano_b200 as mb

    return mb.RigidBodyTransform(td.random_rotation(rng), rng.uniform(-1, 1, size=3))


def _rand_body(mb, rng, name, joint):
    pose = mb.RigidBodyTransform(td.random_rotation(rng), rng.uniform(-0.5, 0.5, size=3))
    return mb.RigidBody(name, joint, td.random_spd_inertia(rng), 0.1 + rng.uniform(), pose)


def build_pair(seed, n_joints, weld_fraction, floating, mode):
    """The same random tree twice: `welded` (some joints Fixed, or ignored with a stored configuration) and `full` (the same
    joints as plain revolute / prismatic joints).  Returns (welded system, full system, {name: q0} of the held joints)."""
    import mecano_b200 as mb

    out = []
    held = {}
    for welded in (True, False):
        rng = np.random.default_rng(seed)
        root = mb.RigidBody("elevator")
        bodies = [root]
        if floating:
            fj = mb.SixDoFJoint("floating", root)
            bodies = [_rand_body(mb, rng, "pelvis", fj)]
        ignore = []
        for k in range(n_joints):
            pred = bodies[rng.integers(len(bodies))]
            T = _rand_transform(rng)
            axis = rng.normal(size=3)
            prismatic = rng.uniform() < 0.3
            hold = rng.uniform() < weld_fraction and not pred.isRootBody()
            q0 = rng.uniform(-1, 1)
            name = "j%d" % k
            if hold and welded and mode == "fixed":
                j = mb.FixedJoint(name, pred, T)
                held[name] = 0.0
            else:
                j = (mb.PrismaticJoint if prismatic else mb.RevoluteJoint)(name, pred, T, axis)
                if hold and mode == "ignore":
                    j.setQ(q0)
                    held[name] = q0
                    if welded:
                        ignore.append(j)
                elif hold:
                    held[name] = 0.0
            bodies.append(_rand_body(mb, rng, "b%d" % k, j))
        out.append(mb.MultiBodySystem.toMultiBodySystemBasics(root, ignore if (welded and mode == "ignore") else None))
    return out[0], out[1], held


def held_closure(full, held):
    """In `ignore` mode every descendant of an ignored joint is ignored too (at its own stored configuration, default 0)."""
    names = dict(held)
    for j in full.getJointsToConsider():
        p = j.getPredecessor().getParentJoint()
        if p is not None and p.getName() in names and j.getName() not in names:
            names[j.getName()] = 0.0
    return names


def run_emu(system, algo, q, qd, x, gravity=G):
    n = q.shape[1]
    nv = system.getNumberOfDoFs()
    out = np.full((nv * nv if algo == 2 else nv, n), np.nan)
    err = ctypes.create_string_buffer(256)
    g = np.ascontiguousarray(gravity, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    f = lambda a: None if a is None else np.ascontiguousarray(a).ctypes.data_as(dp)  # noqa: E731
    q, qd, x = (None if a is None else np.ascontiguousarray(a) for a in (q, qd, x))
    rc = el.lib().emu_run(algo, 0, ctypes.cast(system.tables(), ctypes.c_void_p), f(g), ctypes.c_long(n), ctypes.c_long(n), f(q), f(qd), f(x), None,
                          out.ctypes.data_as(dp), ctypes.c_uint(0), err, 256)
    assert rc == 0, err.value.decode()
    return out


def run_gpu(system, algo, q, qd, x, gravity=G):
    import torch

    import mecano_b200 as mb

    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    if algo == 0:
        c = mb.InverseDynamicsCalculator(system).setKernelVariant("thread")
        c.setGravitationalAcceleration(*gravity)
        return c.compute(t(q), t(qd), t(x)).cpu().numpy()
    if algo == 1:
        c = mb.ForwardDynamicsCalculator(system).setKernelVariant("thread")
        c.setGravitationalAcceleration(*gravity)
        return c.compute(t(q), t(qd), t(x)).cpu().numpy()
    c = mb.CompositeRigidBodyMassMatrixCalculator(system).setKernelVariant("thread")
    return c.getMassMatrix(t(q)).cpu().numpy()


def check_pair(welded, full, held, run, n=5, seed=0):
    import mecano_b200 as mb

    rng = np.random.default_rng(seed)
    t_full = td.TreeDesc(**full.describe()).contiguous()
    oracle = ol.Oracle(t_full, gravity=G)
    qf, qdf, qddf, tauf = mb.MultiBodySystemRandomTools.nextState(rng, full, n)
    pw, pf = welded.getJointMatrixIndexProvider(), full.getJointMatrixIndexProvider()
    by_name = {j.getName(): j for j in full.getJointsToConsider()}
    # hold the welded joints still in the full model
    for name, q0 in held.items():
        r = pf.getJointDoFIndices(by_name[name])[0]
        c = pf.getJointConfigurationIndices(by_name[name])[0]
        qf[c], qdf[r], qddf[r] = q0, 0.0, 0.0
    # rows of the moving joints in both systems
    rows_w, rows_f, cfg_w, cfg_f = [], [], [], []
    for j in welded.getJointsToConsider():
        if j.getDegreesOfFreedom() == 0:
            continue
        rows_w += pw.getJointDoFIndices(j)
        rows_f += pf.getJointDoFIndices(by_name[j.getName()])
        cfg_w += pw.getJointConfigurationIndices(j)
        cfg_f += pf.getJointConfigurationIndices(by_name[j.getName()])
    nvw, nqw = welded.getNumberOfDoFs(), welded.getConfigurationMatrixSize()
    assert sorted(rows_w) == list(range(nvw)) and len(rows_f) == nvw and nvw == full.getNumberOfDoFs() - len(held)
    qw, qdw, qddw, tauw = np.zeros((nqw, n)), np.zeros((nvw, n)), np.zeros((nvw, n)), np.zeros((nvw, n))
    qw[cfg_w], qdw[rows_w], qddw[rows_w], tauw[rows_w] = qf[cfg_f], qdf[rows_f], qddf[rows_f], tauf[rows_f]

    def err(a, b):
        return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))

    # RNEA
    tau_ref = oracle.rnea_batch(qf, qdf, qddf)
    tau = run(welded, 0, qw, qdw, qddw)
    assert err(tau[rows_w], tau_ref[rows_f]) < 1e-10
    # CRBA
    M_ref = oracle.crba_batch(qf)
    M = run(welded, 2, qw, None, None).reshape(nvw, nvw, n)
    assert err(M[np.ix_(rows_w, rows_w)], M_ref[np.ix_(rows_f, rows_f)]) < 1e-10
    # ABA
    bias = oracle.rnea_batch(qf, qdf, np.zeros_like(qddf))
    qdd = run(welded, 1, qw, qdw, tauw)
    for s in range(n):
        ref = np.linalg.solve(M_ref[np.ix_(rows_f, rows_f)][:, :, s], tauf[rows_f, s] - bias[rows_f, s])
        assert err(qdd[rows_w, s], ref) < 1e-8


CASES = [
    dict(seed=11, n_joints=8, weld_fraction=0.4, floating=False, mode="fixed"),
    dict(seed=12, n_joints=20, weld_fraction=0.3, floating=True, mode="fixed"),
    dict(seed=13, n_joints=10, weld_fraction=0.3, floating=False, mode="ignore"),
    dict(seed=14, n_joints=24, weld_fraction=0.2, floating=True, mode="ignore"),
]


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_welded_systems_match_full_tree_with_held_joints_emulated(idx):
    welded, full, held = build_pair(**CASES[idx])
    if CASES[idx]["mode"] == "ignore":
        held = held_closure(full, held)
    assert held, "case does not weld anything"
    check_pair(welded, full, held, run_emu, seed=idx)


def test_index_provider_and_errors():
    import mecano_b200 as mb

    welded, full, held = build_pair(**CASES[0])
    names = [j.getName() for j in welded.getJointsToConsider()]
    assert names == [j.getName() for j in full.getJointsToConsider()]  # fixed joints keep their place in the joint list
    fixed = [j for j in welded.getJointsToConsider() if isinstance(j, mb.FixedJoint)]
    assert fixed and all(j.getDegreesOfFreedom() == 0 and welded.getJointMatrixIndexProvider().getJointDoFIndices(j) == [] for j in fixed)
    assert welded.tables().contents.n_bodies == len(names) - len(fixed)
    with pytest.raises(NotImplementedError):
        welded.describe()
    # ignored joints disappear from the joint list together with their descendants
    welded, full, held = build_pair(**CASES[2])
    closure = held_closure(full, held)
    assert {j.getName() for j in welded.getJointsToIgnore()} == set(closure)
    assert [j.getName() for j in welded.getJointsToConsider()] == [j.getName() for j in full.getJointsToConsider() if j.getName() not in closure]
    # a tree whose every joint is fixed has nothing to compute
    e = mb.RigidBody("elevator")
    mb.RigidBody("b", mb.FixedJoint("f", e), np.eye(3), 1.0, np.zeros(3))
    with pytest.raises(mb.ScrewTheoryException):
        mb.MultiBodySystem.toMultiBodySystemBasics(e)


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(len(CASES)))
def test_welded_systems_match_full_tree_with_held_joints_gpu(idx):
    welded, full, held = build_pair(**CASES[idx])
    if CASES[idx]["mode"] == "ignore":
        held = held_closure(full, held)
    check_pair(welded, full, held, run_gpu, n=64, seed=idx)


# ---------------------------------------------------------------------------------------------------------------------------
# External wrenches and per-body results on systems with fixed / ignored joints (InverseDynamicsCalculator.java:469-472, :578-602
# with :832-860): the calculators run them on the expanded tables (MultiBodySystem.expanded(): nothing welded, the fixed /
# ignored joints held).  Checked against the oracle on the FULL tree with the held joints kept still, as above.
def _extended(m, extra, fill=None):
    out = np.zeros((m.shape[0] + extra, m.shape[1]))
    out[:m.shape[0]] = m
    if extra and fill is not None:
        out[m.shape[0]:] = np.asarray(fill)[:, None]
    return out


def wrenches_emu(welded, q, qd, qdd, tau, fext):
    """What calculators.py does for such a call, by hand, on the kernel source compiled for the host."""
    x = welded.expanded()
    n = q.shape[1]
    nv, nj = welded.getNumberOfDoFs(), welded.getNumberOfJoints()
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(dp)  # noqa: E731
    g = np.ascontiguousarray(G, dtype=np.float64)
    err = ctypes.create_string_buffer(256)
    q2, qd2, qdd2, tau2 = _extended(q, x.n_extra_cfg, x.q_extra), _extended(qd, x.n_extra_dof), _extended(qdd, x.n_extra_dof), _extended(tau, x.n_extra_dof)
    f2 = _extended(fext, 6 * x.n_extra_wrench_blocks)
    rows2 = 6 * (nj + x.n_extra_wrench_blocks)
    out, acc, wr = np.full((nv + x.n_extra_dof, n), np.nan), np.full((rows2, n), np.nan), np.full((rows2, n), np.nan)
    tb = ctypes.cast(x.tables(), ctypes.c_void_p)
    rc = el.lib().emu_rnea_full(tb, P(g), ctypes.c_long(n), ctypes.c_long(n), P(q2), P(qd2), P(qdd2), P(f2), P(out), P(acc), P(wr), ctypes.c_uint(0), err, 256)
    assert rc == 0, err.value.decode()
    qdd_fd = np.full((nv + x.n_extra_dof, n), np.nan)
    locked = np.ascontiguousarray(x.locked, dtype=np.int32)
    rc = el.lib().emu_aba_sources(tb, P(g), ctypes.c_long(n), ctypes.c_long(n), P(q2), P(qd2), P(tau2), P(np.zeros_like(qd2)), P(f2), locked.ctypes.data_as(ip),
                                  P(qdd_fd), err, 256)
    assert rc == 0, err.value.decode()
    return out[:nv], acc[:6 * nj], wr[:6 * nj], qdd_fd[:nv]


def wrenches_gpu(welded, q, qd, qdd, tau, fext):
    import torch

    import mecano_b200 as mb

    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    ident = mb.InverseDynamicsCalculator(welded).setComputeByProducts()
    ident.setGravitationalAcceleration(*G)
    ident.setExternalWrenches(t(fext))
    tau_d = ident.compute(t(q), t(qd), t(qdd)).cpu().numpy()
    acc_d, wr_d = ident.getBodyAccelerationMatrix().cpu().numpy(), ident.getComputedJointWrenchMatrix().cpu().numpy()
    # host matrices through the same calculator: the same kernels
    ident.setExternalWrenches(fext)
    tau_h = ident.compute(q, qd, qdd)
    assert np.array_equal(tau_h, tau_d) and np.array_equal(ident.getBodyAccelerationMatrix(), acc_d) and np.array_equal(ident.getComputedJointWrenchMatrix(), wr_d)
    # getters of the reference API
    j = welded.getJointsToConsider()[-1]
    k = welded.getAllJoints().index(j)
    assert np.array_equal(ident.getComputedJointWrench(j), wr_d[6 * k:6 * k + 6]) and np.array_equal(ident.getBodyAcceleration(j.getSuccessor()), acc_d[6 * k:6 * k + 6])
    fdyn = mb.ForwardDynamicsCalculator(welded)
    fdyn.setGravitationalAcceleration(*G)
    fdyn.setExternalWrenches(t(fext))
    qdd_fd = fdyn.compute(t(q), t(qd), t(tau)).cpu().numpy()
    return tau_d, acc_d, wr_d, qdd_fd


def check_pair_wrenches(welded, full, held, run, n=4, seed=0):
    import mecano_b200 as mb

    rng = np.random.default_rng(100 + seed)
    t_full = td.TreeDesc(**full.describe()).contiguous()
    oracle = ol.Oracle(t_full, gravity=G)
    qf, qdf, qddf, tauf = mb.MultiBodySystemRandomTools.nextState(rng, full, n)
    pw, pf = welded.getJointMatrixIndexProvider(), full.getJointMatrixIndexProvider()
    by_name = {j.getName(): j for j in full.getJointsToConsider()}
    full_index = {j.getName(): i for i, j in enumerate(full.getJointsToConsider())}
    for name, q0 in held.items():
        r, c = pf.getJointDoFIndices(by_name[name])[0], pf.getJointConfigurationIndices(by_name[name])[0]
        qf[c], qdf[r], qddf[r] = q0, 0.0, 0.0
    rows_w, rows_f, cfg_w, cfg_f = [], [], [], []
    for j in welded.getJointsToConsider():
        rows_w += pw.getJointDoFIndices(j)
        cfg_w += pw.getJointConfigurationIndices(j)
        if j.getDegreesOfFreedom() > 0:
            rows_f += pf.getJointDoFIndices(by_name[j.getName()])
            cfg_f += pf.getJointConfigurationIndices(by_name[j.getName()])
    nvw, nqw, njw = welded.getNumberOfDoFs(), welded.getConfigurationMatrixSize(), welded.getNumberOfJoints()
    qw, qdw, qddw, tauw = np.zeros((nqw, n)), np.zeros((nvw, n)), np.zeros((nvw, n)), np.zeros((nvw, n))
    qw[cfg_w], qdw[rows_w], qddw[rows_w], tauw[rows_w] = qf[cfg_f], qdf[rows_f], qddf[rows_f], tauf[rows_f]
    # a wrench on every body the welded system knows (the successors of FixedJoints included), none on ignored bodies
    fext_w = rng.uniform(-1, 1, size=(6 * njw, n))
    fext_f = np.zeros((6 * t_full.nb, n))
    blocks_f = []
    for k, j in enumerate(welded.getJointsToConsider()):
        i = full_index[j.getName()]
        fext_f[6 * i:6 * i + 6] = fext_w[6 * k:6 * k + 6]
        blocks_f += list(range(6 * i, 6 * i + 6))
    tau, acc, wr, qdd_fd = run(welded, qw, qdw, qddw, tauw, fext_w)

    def err(a, b):
        return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))

    assert not (np.isnan(tau).any() or np.isnan(acc).any() or np.isnan(wr).any() or np.isnan(qdd_fd).any())
    M_ref = oracle.crba_batch(qf)
    bias = oracle.rnea_batch(qf, qdf, np.zeros_like(qddf), fext_f)
    for s in range(n):
        tau_o, acc_o, wr_o = oracle.rnea_full(qf[:, s], qdf[:, s], qddf[:, s], np.ascontiguousarray(fext_f[:, s].reshape(t_full.nb, 6)))
        assert err(tau[rows_w, s], tau_o[rows_f]) < 1e-10
        assert err(acc[:, s], acc_o.reshape(-1)[blocks_f]) < 1e-10  # every considered body, in its own CoM frame
        assert err(wr[:, s], wr_o.reshape(-1)[blocks_f]) < 1e-10    # every considered joint, FixedJoints included
        ref = np.linalg.solve(M_ref[np.ix_(rows_f, rows_f)][:, :, s], tauf[rows_f, s] - bias[rows_f, s])
        assert err(qdd_fd[rows_w, s], ref) < 1e-8


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_wrenches_and_byproducts_on_welded_systems_emulated(idx):
    welded, full, held = build_pair(**CASES[idx])
    if CASES[idx]["mode"] == "ignore":
        held = held_closure(full, held)
    x = welded.expanded()
    assert x.n_bodies == full.getNumberOfJoints() and int(x.locked.sum()) == len(held)
    assert x.n_extra_dof == len(held) and x.n_extra_cfg == len(held)
    assert x.n_extra_wrench_blocks == (len(held) if CASES[idx]["mode"] == "ignore" else 0)
    check_pair_wrenches(welded, full, held, wrenches_emu, seed=idx)


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(len(CASES)))
def test_wrenches_and_byproducts_on_welded_systems_gpu(idx):
    welded, full, held = build_pair(**CASES[idx])
    if CASES[idx]["mode"] == "ignore":
        held = held_closure(full, held)
    check_pair_wrenches(welded, full, held, wrenches_gpu, n=33, seed=idx)
