"""GPU tests (run with -m gpu on a B200) of the host-facing half of the boundary added in round 2: the packed mass-matrix
layout, the fused host step (mecano_b200_step_host), and the multi-device engine (mecano_b200_multi_*), whose slices must
concatenate to the bit-identical single-device result (test matrix "multi-GPU").

The multi-device tests run on whatever the box has: with one GPU the device list repeats device 0 (every entry of the list owns
its own handle, staging buffers and streams, so the slicing, the pointer offsets and the round-robin issue are exercised just
the same); with 2 / 4 / 8 GPUs the real devices are used as well."""
import numpy as np
import pytest

import oracle_lib as ol
import treedesc as td
from test_gpu_parity import build, rel, torch_dev  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

TOL = 1e-9


def pinned(shape):
    import torch

    return torch.empty(shape, dtype=torch.float64).pin_memory().numpy()


def states(mb, s, rng, n, ld=None):
    ld = ld or n
    q, qd, qdd, tau = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    out = []
    for a in (q, qd, qdd, tau):
        b = pinned((a.shape[0], ld))
        b[:] = np.nan
        b[:, :n] = a
        out.append(b[:, :n])
    return out


@pytest.mark.parametrize("case", ["H37", "tree", "two_floating"])
def test_packed_mass_matrix_device_and_host(torch_dev, case):
    """MECANO_B200_CRBA_PACKED on the device and on the host entry point: packed -> dense through the exported index map equals the
    oracle (1e-9) and the dense kernel bit for bit; entries the map does not list are structurally zero."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    if case == "H37":
        s, t = build(kind="humanoid", seed=7, n_joints=2)
    elif case == "tree":
        s, t = build(kind="tree", seed=31, n_joints=45, floating=True, prismatic=0.3)
    else:
        e = mb.RigidBody("elevator")
        b1 = mb.MultiBodySystemRandomTools.nextFloatingBase(5, e).getSuccessor()
        b2 = mb.MultiBodySystemRandomTools.nextFloatingBase(6, b1).getSuccessor()
        mb.MultiBodySystemRandomTools.nextOneDoFJointTree(7, b2, 6, 0.5)
        s = mb.MultiBodySystem.toMultiBodySystemBasics(e)
        t = td.TreeDesc(**s.describe()).contiguous()
    o = ol.Oracle(t)
    rng = np.random.default_rng(4)
    n, nv = 3001, t.nv
    q = states(mb, s, rng, n)[0]
    tq = torch.from_numpy(q).to(dev)
    crba = mb.CompositeRigidBodyMassMatrixCalculator(s).setKernelVariant("thread")
    row, col = crba.getMassMatrixPackedIndex()
    assert np.all(row <= col) and len(set(zip(row.tolist(), col.tolist()))) == len(row)
    if case == "H37":
        assert len(row) == 362
    P = crba.getMassMatrix(tq, torch.full((len(row), n), float("nan"), dtype=torch.float64, device=dev), packed=True)
    assert not torch.isnan(P).any(), "every packed row must be written"
    P = P.cpu().numpy()
    M = np.zeros((nv, nv, n))
    M[row, col] = P
    M[col, row] = P
    Mo = o.crba_batch(q)
    assert rel(M, Mo) < TOL
    dense = crba.getMassMatrix(tq).cpu().numpy().reshape(nv, nv, n)
    assert np.array_equal(M, dense), "packed and dense kernels must agree bit for bit"
    covered = np.zeros((nv, nv), bool)
    covered[row, col] = covered[col, row] = True
    assert np.all(Mo[~covered] == 0.0)
    # host entry point, ld > n
    Ph = pinned((len(row), n + 39))[:, :n]
    qh = pinned((t.nq, n + 39))[:, :n]
    qh[:] = q
    Ph[:] = np.nan
    crba.getMassMatrix(qh, Ph, packed=True)
    assert np.array_equal(Ph, P)
    # the layouts do not combine
    with pytest.raises(mb.MecanoB200Error):
        crba._engine.crba_host(qh, Ph, mb._capi.CRBA_PACKED | mb._capi.CRBA_ZEROS_PRESENT)
    with pytest.raises(mb.MecanoB200Error):
        crba.setPrecision("fp32").getMassMatrix(tq, packed=True)


def test_fused_host_step_is_bit_identical_to_separate_calls(torch_dev):
    """mecano_b200_step_host = InverseDynamicsCalculator.compute + ForwardDynamicsCalculator.compute + getMassMatrix on the same
    states, q / qd uploaded once: bit-identical to the three host calls, every layout, and each part can be left out."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="humanoid", seed=7, n_joints=2)
    rng = np.random.default_rng(9)
    n, ld, nv = 20011, 20096, t.nv
    q, qd, qdd, tau = states(mb, s, rng, n, ld)
    g = (0.3, -0.2, -9.81)
    ident = mb.InverseDynamicsCalculator(s)
    fdyn = mb.ForwardDynamicsCalculator(s)
    crba = mb.CompositeRigidBodyMassMatrixCalculator(s)
    step = mb.MultiBodyDynamicsStep(s)
    for c in (ident, fdyn, step):
        c.setGravitationalAcceleration(g)
    new = lambda rows: pinned((rows, ld))[:, :n]  # noqa: E731
    tau_ref, qdd_ref = ident.compute(q, qd, qdd, new(nv)), fdyn.compute(q, qd, tau, new(nv))
    o = ol.Oracle(t, gravity=g)
    assert rel(tau_ref, o.rnea_batch(q, qd, qdd)) < TOL and rel(qdd_ref, o.aba_batch(q, qd, tau)) < TOL
    for layout in ("dense", "packed", "stateMajor"):
        if layout == "stateMajor":
            M_ref = crba.getMassMatrix(q, pinned((n, nv * nv)), stateMajor=True)
            M = pinned((n, nv * nv))
        else:
            rows = step.getMassMatrixRows(packed=layout == "packed")
            M_ref = crba.getMassMatrix(q, new(rows), packed=layout == "packed")
            M = new(rows)
        M[:] = np.nan
        tau_out, qdd_out = new(nv), new(nv)
        tau_out[:] = np.nan
        qdd_out[:] = np.nan
        step.compute(q, qd, qdd=qdd, tau=tau, tauOut=tau_out, qddOut=qdd_out, massMatrix=M, packed=layout == "packed", stateMajor=layout == "stateMajor")
        assert np.array_equal(tau_out, tau_ref) and np.array_equal(qdd_out, qdd_ref) and np.array_equal(M, M_ref), layout
    # parts left out
    qdd_only = new(nv)
    step.compute(q, qd, tau=tau, qddOut=qdd_only)
    assert np.array_equal(qdd_only, qdd_ref)
    M_only = new(nv * nv)
    step.compute(q, None, massMatrix=M_only)
    assert np.array_equal(M_only, crba.getMassMatrix(q, new(nv * nv)))
    with pytest.raises(mb.MecanoB200Error):
        step.compute(q, qd, qdd=qdd)  # an input without its output
    with pytest.raises(mb.MecanoB200Error):
        step.compute(q, None)  # nothing to do


def device_lists():
    import torch

    ndev = torch.cuda.device_count()
    lists = [[0], [0, 0], [0, 0, 0], [0] * 8]
    for k in (2, 4, 8):
        if ndev >= k:
            lists.append(list(range(k)))
    return lists


@pytest.mark.parametrize("n", [100003, 1500, 200])
def test_multi_device_slices_are_bit_identical(torch_dev, n):
    """"slices 1/2/4/8 concatenate to the bit-identical single-GPU result".  One call on a MultiDeviceEngine
    over 1, 2, 3, 8 list entries (and over the real GPUs of the box when there are 2 / 4 / 8) against the single-device engine,
    for every host entry point and layout, with a ragged batch and ld > n.  n = 1500 and 200 sit below the AUTO crossover: the
    warp-per-state kernels must then be chosen for every slice, by the size of the whole batch."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="humanoid", seed=7, n_joints=2)
    rng = np.random.default_rng(n)
    ld, nv = n + 61, t.nv
    q, qd, qdd, tau = states(mb, s, rng, n, ld)
    fext = pinned((6 * t.nb, ld))[:, :n]
    fext[:] = rng.uniform(-1, 1, size=fext.shape)
    g = (0.0, 0.0, -9.81)
    new = lambda rows: pinned((rows, ld))[:, :n]  # noqa: E731
    single = {}
    for devices in device_lists():
        dev_arg = devices[0] if len(devices) == 1 else devices
        ident = mb.InverseDynamicsCalculator(s, device=dev_arg)
        fdyn = mb.ForwardDynamicsCalculator(s, device=dev_arg)
        crba = mb.CompositeRigidBodyMassMatrixCalculator(s, device=dev_arg)
        step = mb.MultiBodyDynamicsStep(s, device=dev_arg)
        for c in (ident, fdyn, step):
            c.setGravitationalAcceleration(g)
        res = {}
        res["tau"] = ident.compute(q, qd, qdd, new(nv)).copy()
        ident.setExternalWrenches(fext)
        res["tau_fext"] = ident.compute(q, qd, qdd, new(nv)).copy()
        ident.setExternalWrenchesToZero()
        ident.setConsiderCoriolisAndCentrifugalForces(False)
        res["tau_nocor"] = ident.compute(q, qd, qdd, new(nv)).copy()
        res["qdd"] = fdyn.compute(q, qd, tau, new(nv)).copy()
        res["M"] = crba.getMassMatrix(q, new(nv * nv)).copy()
        res["M_sm"] = crba.getMassMatrix(q, pinned((n, nv * nv)), stateMajor=True).copy()
        res["M_pk"] = crba.getMassMatrix(q, new(step.getMassMatrixRows(packed=True)), packed=True).copy()
        t_o, q_o, M_o = new(nv), new(nv), new(step.getMassMatrixRows(packed=True))
        step.compute(q, qd, qdd=qdd, tau=tau, tauOut=t_o, qddOut=q_o, massMatrix=M_o, packed=True)
        res["step_tau"], res["step_qdd"], res["step_M"] = t_o.copy(), q_o.copy(), M_o.copy()
        if len(devices) == 1:
            single = res
            o = ol.Oracle(t, gravity=g)
            assert rel(res["tau"], o.rnea_batch(q, qd, qdd)) < TOL and rel(res["qdd"], o.aba_batch(q, qd, tau)) < TOL
            assert rel(res["M"].reshape(nv, nv, n), o.crba_batch(q)) < TOL
            assert np.array_equal(res["step_tau"], res["tau"]) and np.array_equal(res["step_M"], res["M_pk"])
            continue
        eng = ident._engine
        sl = eng.slices(n)
        assert sum(c for _, c in sl) == n and all(sl[i][0] + sl[i][1] == sl[i + 1][0] for i in range(len(sl) - 1)), "slices must tile the batch"
        for key, val in res.items():
            assert not np.isnan(val).any(), (devices, key)
            assert np.array_equal(val, single[key]), "devices %s: %s differs from the single-device result" % (devices, key)
        with pytest.raises(TypeError):
            ident.compute(torch.from_numpy(q).to(dev), torch.from_numpy(qd).to(dev), torch.from_numpy(qdd).to(dev))


def test_multi_device_errors_and_empty_batch(torch_dev):
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="chain", seed=1, n_joints=7)
    with pytest.raises(mb.MecanoB200Error):
        mb.InverseDynamicsCalculator(s, device=[0, 99])
    with pytest.raises(mb.MecanoB200Error):
        mb.InverseDynamicsCalculator(s, device=[])
    ident = mb.InverseDynamicsCalculator(s, device=[0, 0])
    z = np.zeros((7, 0))
    assert ident.compute(z, z, z).shape == (7, 0)
    one = np.zeros((7, 1))
    assert ident.compute(one, one, one).shape == (7, 1)  # the second entry's slice is empty


def test_host_state_major_owned_matrix_survives_other_calls(torch_dev):
    """ADVICE (round 1): getMassMatrix(q, stateMajor=True) on host matrices, calculator-owned result (ZEROS_PRESENT from the second
    call on), with other host calls on the same handle in between: the structurally zero entries must still be zero."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="humanoid", seed=7, n_joints=2)
    o = ol.Oracle(t)
    rng = np.random.default_rng(12)
    n, nv = 6000, t.nv
    crba = mb.CompositeRigidBodyMassMatrixCalculator(s)
    for k in range(3):
        q = states(mb, s, rng, n)[0]
        M = crba.getMassMatrix(q, stateMajor=True)
        assert rel(M.reshape(n, nv, nv).transpose(1, 2, 0), o.crba_batch(q)) < TOL, k
        # same handle, same staging buffers: a caller-supplied entry-major matrix full of NaN-free garbage in between
        junk = states(mb, s, rng, n)[0]
        big = crba.getMassMatrix(junk, pinned((nv * nv, n)))
        assert np.isfinite(big).all()


def test_fused_host_step_owned_mass_matrix(torch_dev):
    """MultiBodyDynamicsStep.compute(ownedMassMatrix=True): the step's own dense matrix, structurally zero entries written by the
    first call and neither recomputed nor transferred from the second on (mecano_b200_step_host with ENTRY_MAJOR | ZEROS_PRESENT).
    Every call must return the full dense matrix bit for bit, with other host calls dirtying the staging buffers in between; the
    state-major layout ignores the flag (the kernel rewrites the zeros) instead of returning stale entries."""
    import mecano_b200 as mb
    from mecano_b200 import _capi

    torch, dev = torch_dev
    s, t = build(kind="humanoid", seed=7, n_joints=2)
    rng = np.random.default_rng(21)
    n, ld, nv = 9001, 9216, t.nv
    crba = mb.CompositeRigidBodyMassMatrixCalculator(s)
    ident = mb.InverseDynamicsCalculator(s)
    step = mb.MultiBodyDynamicsStep(s)
    new = lambda rows: pinned((rows, ld))[:, :n]  # noqa: E731
    zero_rows = None
    for k in range(3):
        q, qd, qdd, tau = states(mb, s, rng, n, ld)
        M_ref = crba.getMassMatrix(q, new(nv * nv)).copy()
        tau_out, qdd_out = new(nv), new(nv)
        _, _, M = step.compute(q, qd, qdd=qdd, tau=tau, tauOut=tau_out, qddOut=qdd_out, ownedMassMatrix=True)
        assert np.array_equal(M, M_ref), k
        assert np.array_equal(tau_out, ident.compute(q, qd, qdd, new(nv))), k
        if zero_rows is None:
            zero_rows = np.flatnonzero(~M_ref.any(axis=1))
            assert len(zero_rows) > 0  # a humanoid has unrelated branches
        # the step hands out the same buffer every time; dirty the staging buffers of the handle with a full dense call
        assert step.compute(q, None, ownedMassMatrix=True)[2] is M
        junk = new(nv * nv)
        step.compute(states(mb, s, rng, n, ld)[0], None, massMatrix=junk)
        assert np.isfinite(junk).all()
    # the same through the multi-device engine (three slices on one GPU: mecano_b200_multi_step_host)
    multi = mb.MultiBodyDynamicsStep(s, device=[0, 0, 0])
    for k in range(2):
        q = states(mb, s, rng, n, ld)[0]
        assert np.array_equal(multi.compute(q, None, ownedMassMatrix=True)[2], crba.getMassMatrix(q, new(nv * nv))), k
    # raw C ABI, state-major with the flag set: the zeros come back as zeros although the host buffer held garbage
    q = states(mb, s, rng, n, ld)[0]
    Ms = pinned((n, nv * nv))
    Ms[:] = 7.0
    step._engine.step_host(q, None, M=Ms, layout=_capi.CRBA_STATE_MAJOR | _capi.CRBA_ZEROS_PRESENT)
    assert np.array_equal(Ms, crba.getMassMatrix(q, pinned((n, nv * nv)), stateMajor=True))
    with pytest.raises(ValueError):
        step.compute(q, None, massMatrix=Ms, ownedMassMatrix=True)
