"""GPU parity tests (run with -m gpu on a B200): every kernel, through the calculator API and the C ABI underneath,
against the CPU oracle on the same seeded inputs, against the committed golden fixtures, and -- at BASELINE.json's
full sizes -- through the size-independent invariants of the reference's own tests (FD o ID = id,
M qdd + ID(qdd = 0) = ID(qdd)).  Tolerance: 1e-9 relative in fp64 (north_star), error idiom of
ForwardDynamicsCalculatorTest.java:1099-1107."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import treedesc as td

pytestmark = pytest.mark.gpu

TOL = 1e-9


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))


@pytest.fixture(scope="module")
def torch_dev():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device; the product has no CPU fallback"
    return torch, torch.device("cuda:0")


def build(kind, seed, n_joints=0, prismatic=0.0, floating=False):
    import mecano_b200 as mb

    e = mb.RigidBody("elevator")
    base = e
    if kind == "humanoid":
        mb.MultiBodySystemRandomTools.nextHumanoid(seed, e, n_joints)
    else:
        if floating:
            base = mb.MultiBodySystemRandomTools.nextFloatingBase(seed + 1000, e).getSuccessor()
        if kind == "chain":
            mb.MultiBodySystemRandomTools.nextOneDoFJointChain(seed, base, n_joints, prismatic)
        elif kind == "jchain":  # joints of all five moving types (MultiBodySystemRandomTools.nextJointChain)
            mb.MultiBodySystemRandomTools.nextJointChain(seed, base, n_joints)
        elif kind == "jtree":
            mb.MultiBodySystemRandomTools.nextJointTree(seed, base, n_joints)
        elif kind in ("spherical", "planar"):
            cls = mb.SphericalJoint if kind == "spherical" else mb.PlanarJoint
            rng = np.random.default_rng(seed)
            for k in range(n_joints):
                T = mb.RigidBodyTransform(td.random_rotation(rng), rng.uniform(-1, 1, size=3))
                j = cls("j%d" % k, base, T)
                base = mb.RigidBody("b%d" % k, j, td.random_spd_inertia(rng), 0.1 + rng.uniform(), rng.uniform(-1, 1, size=3))
        else:
            mb.MultiBodySystemRandomTools.nextOneDoFJointTree(seed, base, n_joints, prismatic)
    s = mb.MultiBodySystem.toMultiBodySystemBasics(e)
    return s, td.TreeDesc(**s.describe()).contiguous()


CASES = [
    ("A7 revolute chain", dict(kind="chain", seed=1, n_joints=7)),
    ("single joint", dict(kind="chain", seed=2, n_joints=1)),
    ("prismatic chain", dict(kind="chain", seed=3, n_joints=6, prismatic=1.0)),
    ("one-dof tree 30", dict(kind="tree", seed=4, n_joints=30, prismatic=0.4)),
    ("floating + chain 20", dict(kind="chain", seed=5, n_joints=20, floating=True)),
    ("floating + tree 50", dict(kind="tree", seed=6, n_joints=50, floating=True, prismatic=0.2)),
    ("H37", dict(kind="humanoid", seed=7, n_joints=2)),
    ("H36", dict(kind="humanoid", seed=8, n_joints=1)),
    ("tree 100", dict(kind="tree", seed=9, n_joints=100, floating=True)),
    ("deep chain 60", dict(kind="chain", seed=10, n_joints=60)),
    # SphericalJoint / PlanarJoint (ForwardDynamicsCalculatorTest.testJointChain / testJointTree: all joint types)
    ("spherical chain 5", dict(kind="spherical", seed=11, n_joints=5)),
    ("planar chain 4", dict(kind="planar", seed=12, n_joints=4)),
    ("joint chain 12", dict(kind="jchain", seed=13, n_joints=12)),
    ("joint tree 25", dict(kind="jtree", seed=14, n_joints=25)),
    ("floating + joint tree 40", dict(kind="jtree", seed=15, n_joints=40, floating=True)),
]


def has_3dof(t):
    return bool(np.any((np.asarray(t.jtype) == td.SPHERICAL) | (np.asarray(t.jtype) == td.PLANAR)))


@pytest.mark.parametrize("variant", ["thread", "warp"])
@pytest.mark.parametrize("idx", range(len(CASES)))
def test_kernels_match_oracle(torch_dev, idx, variant):
    import mecano_b200 as mb

    torch, dev = torch_dev
    name, kw = CASES[idx]
    s, t = build(**kw)
    if variant == "warp" and has_3dof(t):
        # one lane per body (a warp per state up to 32 bodies, a team of two to four warps up to 128), one-DoF and SixDoF joints:
        # other trees are refused, never silently rerouted
        with pytest.raises(mb.MecanoB200Error):
            mb.InverseDynamicsCalculator(s).setKernelVariant("warp")
        return
    rng = np.random.default_rng(1000 + idx)
    g = (rng.uniform(-1, 1), rng.uniform(-1, 1), -rng.uniform(1, 10))
    o = ol.Oracle(t, gravity=g)
    n = 777  # ragged: not a multiple of the block size
    q, qd, qdd, tau = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    fext = np.ascontiguousarray(rng.uniform(-1, 1, size=(6 * t.nb, n)))
    tq, tqd, tqdd, ttau, tf = (torch.from_numpy(x).to(dev) for x in (q, qd, qdd, tau, fext))
    nv = t.nv

    ident = mb.InverseDynamicsCalculator(s).setKernelVariant(variant)
    ident.setGravitationalAcceleration(g)
    assert ident.kernelInfo(n)["variant"] == {"thread": 1, "warp": 2}[variant]
    assert rel(ident.compute(tq, tqd, tqdd).cpu().numpy(), o.rnea_batch(q, qd, qdd)) < TOL, name
    ident.setExternalWrenches(tf)
    assert rel(ident.compute(tq, tqd, tqdd).cpu().numpy(), o.rnea_batch(q, qd, qdd, fext)) < TOL, name
    ident.setExternalWrenchesToZero()
    ident.setConsiderCoriolisAndCentrifugalForces(False)
    assert rel(ident.compute(tq, tqd, tqdd).cpu().numpy(), o.rnea_batch(q, qd, qdd, flags=1)) < TOL, name
    ident.setConsiderJointAccelerations(False)
    assert rel(ident.compute(tq, tqd, tqdd).cpu().numpy(), o.rnea_batch(q, qd, qdd, flags=3)) < TOL, name

    fdyn = mb.ForwardDynamicsCalculator(s).setKernelVariant(variant)
    fdyn.setGravitationalAcceleration(*g)
    assert rel(fdyn.compute(tq, tqd, ttau).cpu().numpy(), o.aba_batch(q, qd, tau)) < TOL, name
    fdyn.setExternalWrenches(tf)
    assert rel(fdyn.compute(tq, tqd, ttau).cpu().numpy(), o.aba_batch(q, qd, tau, fext)) < TOL, name

    crba = mb.CompositeRigidBodyMassMatrixCalculator(s).setKernelVariant(variant)
    Mo = o.crba_batch(q)
    M = crba.getMassMatrix(tq, torch.full((nv * nv, n), float("nan"), dtype=torch.float64, device=dev))
    assert not torch.isnan(M).any(), "every entry of the dense matrix must be written"
    assert rel(M.cpu().numpy().reshape(nv, nv, n), Mo) < TOL, name
    Ms = crba.getMassMatrix(tq, stateMajor=True)  # Mecano's per-state dense layout
    assert rel(Ms.cpu().numpy().reshape(n, nv, nv).transpose(1, 2, 0), Mo) < TOL, name


@pytest.mark.parametrize("variant", ["thread", "warp"])
@pytest.mark.parametrize("idx", [0, 3, 8])
def test_angles_beyond_the_fast_sincos_range(torch_dev, idx, variant):
    """Joint angles of any magnitude (Math.sin / cos reduce exactly): the fast sin/cos of the kernels covers |q| < 1e5, larger
    angles are redone by the library routine (jointmath.cuh: mb_angle_large); prismatic displacements of that size pass through."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    name, kw = CASES[idx]
    s, t = build(**kw)
    rng = np.random.default_rng(5200 + idx)
    o = ol.Oracle(t, gravity=(0.0, 0.0, -9.81))
    n = 1500
    q, qd, qdd, tau = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    rev = [int(t.cfg_off[i]) for i in range(t.nb) if t.jtype[i] == td.REVOLUTE]
    pris = [int(t.cfg_off[i]) for i in range(t.nb) if t.jtype[i] == td.PRISMATIC]
    for st in range(0, n, 3):
        for r in rng.permutation(rev)[: 1 + st % 4]:
            q[r, st] = rng.choice([-1.0, 1.0]) * 10.0 ** rng.uniform(5.0, 9.0)
    q[rev[0], 1] = 1.0e5
    ident = mb.InverseDynamicsCalculator(s).setKernelVariant(variant)
    ident.setGravitationalAcceleration(-9.81)
    fdyn = mb.ForwardDynamicsCalculator(s).setKernelVariant(variant)
    fdyn.setGravitationalAcceleration(-9.81)
    crba = mb.CompositeRigidBodyMassMatrixCalculator(s).setKernelVariant(variant)
    tq, tqd, tqdd, ttau = (torch.from_numpy(x).to(dev) for x in (q, qd, qdd, tau))
    assert rel(ident.compute(tq, tqd, tqdd).cpu().numpy(), o.rnea_batch(q, qd, qdd)) < TOL, name
    assert rel(fdyn.compute(tq, tqd, ttau).cpu().numpy(), o.aba_batch(q, qd, tau)) < TOL, name
    assert rel(crba.getMassMatrix(tq).cpu().numpy().reshape(t.nv, t.nv, n), o.crba_batch(q)) < TOL, name
    if pris:
        # a displacement of that size is not an angle (forward dynamics is too ill-conditioned there to compare at 1e-9)
        for st in range(0, n, 2):
            q[rng.choice(pris), st] = rng.choice([-1.0, 1.0]) * 10.0 ** rng.uniform(5.0, 6.0)
        tq = torch.from_numpy(q).to(dev)
        assert rel(ident.compute(tq, tqd, tqdd).cpu().numpy(), o.rnea_batch(q, qd, qdd)) < TOL, name
        assert rel(crba.getMassMatrix(tq).cpu().numpy().reshape(t.nv, t.nv, n), o.crba_batch(q)) < TOL, name


def test_work_distribution_of_the_persistent_grids_is_result_neutral(torch_dev):
    """The warps of a persistent grid (RNEA / ABA with the stack in tensor memory) draw their states from a counter; MECANO_B200_DRAW=0
    falls back to tiles taken round-robin.  Which warp evaluates a state must not show in the result: both settings, twice each (the
    counter re-arms itself between launches), bit for bit; and a batch that is not a multiple of the warp size."""
    import subprocess
    import sys

    code = (
        "import hashlib, numpy as np, torch, mecano_b200 as mb\n"
        "e = mb.RigidBody('elevator'); mb.MultiBodySystemRandomTools.nextHumanoid(11, e, 2)\n"
        "s = mb.MultiBodySystem.toMultiBodySystemBasics(e); n = 300007\n"
        "rng = np.random.default_rng(5); q, qd, qdd, tau = mb.MultiBodySystemRandomTools.nextState(rng, s, n)\n"
        "d = torch.device('cuda:0'); tq, tqd, tqdd, ttau = (torch.from_numpy(x).to(d) for x in (q, qd, qdd, tau))\n"
        "i = mb.InverseDynamicsCalculator(s); f = mb.ForwardDynamicsCalculator(s)\n"
        "assert i.kernelInfo(n)['tmem_stack_slots'] > 0 and f.kernelInfo(n)['tmem_stack_slots'] > 0\n"
        "for k in range(2):\n"
        "    a = i.compute(tq, tqd, tqdd).cpu().numpy(); b = f.compute(tq, tqd, ttau).cpu().numpy()\n"
        "    print(hashlib.sha256(a.tobytes() + b.tobytes()).hexdigest())\n")
    out = []
    for draw in ("1", "0"):
        env = dict(os.environ, MECANO_B200_DRAW=draw, PYTHONPATH=os.pathsep.join([os.path.dirname(os.path.dirname(os.path.abspath(__file__)))] + sys.path))
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out += r.stdout.split()
    assert len(out) == 4 and len(set(out)) == 1, out


@pytest.mark.parametrize("idx", [8, 9])
def test_persistent_grids_on_ragged_batches_of_deep_trees(torch_dev, idx):
    """Trees too deep for the tensor-memory stack run the shared-memory kernels; forward dynamics is a persistent grid there too, and
    its warps draw their states from the counter: a batch that is neither a multiple of the block size nor of the warp size, large
    enough for the grid to be persistent, first and last states against the oracle, the rest against a second launch."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    name, kw = CASES[idx]
    s, t = build(**kw)
    rng = np.random.default_rng(6100 + idx)
    o = ol.Oracle(t, gravity=(0.0, 0.0, -9.81))
    n = 148 * 384 * 2 + 77
    q, qd, qdd, tau = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    tq, tqd, tqdd, ttau = (torch.from_numpy(x).to(dev) for x in (q, qd, qdd, tau))
    ident = mb.InverseDynamicsCalculator(s)
    ident.setGravitationalAcceleration(-9.81)
    fdyn = mb.ForwardDynamicsCalculator(s)
    fdyn.setGravitationalAcceleration(-9.81)
    got_tau = ident.compute(tq, tqd, tqdd).cpu().numpy()
    got_qdd = fdyn.compute(tq, tqd, ttau).cpu().numpy()
    assert not (np.isnan(got_tau).any() or np.isnan(got_qdd).any())
    for sl in (slice(0, 200), slice(n - 200, n)):
        assert rel(got_tau[:, sl], o.rnea_batch(np.ascontiguousarray(q[:, sl]), np.ascontiguousarray(qd[:, sl]), np.ascontiguousarray(qdd[:, sl]))) < TOL, name
        assert rel(got_qdd[:, sl], o.aba_batch(np.ascontiguousarray(q[:, sl]), np.ascontiguousarray(qd[:, sl]), np.ascontiguousarray(tau[:, sl]))) < TOL, name
    assert np.array_equal(fdyn.compute(tq, tqd, ttau).cpu().numpy(), got_qdd)
    assert np.array_equal(ident.compute(tq, tqd, tqdd).cpu().numpy(), got_tau)


def test_host_entry_points_and_leading_dimension(torch_dev):
    """*_host entry points (numpy in, numpy out), with ld > n, plus empty and single-state batches."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="tree", seed=21, n_joints=25, floating=True, prismatic=0.3)
    o = ol.Oracle(t, gravity=(0, 0, -9.81))
    rng = np.random.default_rng(21)
    n, ld = 5000, 5120
    q, qd, qdd, tau = mb.MultiBodySystemRandomTools.nextState(rng, s, ld)
    ident = mb.InverseDynamicsCalculator(s)
    ident.setGravitationalAcceleration(-9.81)
    fdyn = mb.ForwardDynamicsCalculator(s)
    fdyn.setGravitationalAcceleration(-9.81)
    crba = mb.CompositeRigidBodyMassMatrixCalculator(s)
    out = np.full((t.nv, ld), np.nan)
    ident.compute(q[:, :n], qd[:, :n], qdd[:, :n], out[:, :n])
    assert np.isnan(out[:, n:]).all(), "columns beyond n must not be touched"
    assert rel(out[:, :n], o.rnea_batch(q[:, :n], qd[:, :n], qdd[:, :n])) < TOL
    assert rel(fdyn.compute(q[:, :n], qd[:, :n], tau[:, :n]), o.aba_batch(q[:, :n], qd[:, :n], tau[:, :n])) < TOL
    ns = 300
    assert rel(crba.getMassMatrix(q[:, :ns]).reshape(t.nv, t.nv, ns), o.crba_batch(q[:, :ns])) < TOL
    assert rel(crba.getMassMatrix(np.ascontiguousarray(q[:, :ns]), stateMajor=True).reshape(ns, t.nv, t.nv).transpose(1, 2, 0), o.crba_batch(q[:, :ns])) < TOL
    # device path with ld > n
    tq, tqd, tqdd = (torch.from_numpy(x).to(dev) for x in (q, qd, qdd))
    tout = torch.full((t.nv, ld), float("nan"), dtype=torch.float64, device=dev)
    ident.compute(tq[:, :n], tqd[:, :n], tqdd[:, :n], tout[:, :n])
    assert torch.isnan(tout[:, n:]).all()
    assert rel(tout[:, :n].cpu().numpy(), out[:, :n]) == 0.0, "host and device entry points run the same kernel"
    # empty and single-state batches
    assert ident.compute(tq[:, :0], tqd[:, :0], tqdd[:, :0]).shape == (t.nv, 0)
    one = ident.compute(tq[:, :1], tqd[:, :1], tqdd[:, :1]).cpu().numpy()
    assert rel(one, out[:, :1]) < 1e-13  # AUTO runs a single state on the warp-per-state kernels: same mathematics, other rounding order


def test_error_behaviour(torch_dev):
    """Shape errors raise like EJML's MatrixDimensionException (ForwardDynamicsCalculator.java:522-533)."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="chain", seed=30, n_joints=4)
    ident = mb.InverseDynamicsCalculator(s)
    good = torch.zeros((4, 10), dtype=torch.float64, device=dev)
    with pytest.raises(mb.MatrixDimensionException):
        ident.compute(good, good, torch.zeros((5, 10), dtype=torch.float64, device=dev))
    with pytest.raises(mb.MatrixDimensionException):
        ident.compute(good, good[:, :9], good)
    with pytest.raises(TypeError):
        ident.compute(good, good, good.float())
    ident.setExternalWrenches(torch.zeros((6 * 3, 10), dtype=torch.float64, device=dev))
    with pytest.raises(mb.MatrixDimensionException):
        ident.compute(good, good, good)


def test_golden_fixtures_on_gpu(torch_dev):
    """The committed regression vectors (tests/golden, generated by the oracle) through the raw C ABI."""
    import mecano_b200 as mb
    from mecano_b200 import _capi

    import emu_lib as el

    torch, dev = torch_dev
    data = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz"))
    for name in sorted({k.split("/")[0] for k in data.files}):
        t = td.TreeDesc(**{f: data["%s/%s" % (name, f)] for f in ("parent", "jtype", "axis", "off_R", "off_p", "com_R", "com_p", "J", "mass", "dof_off", "cfg_off")},
                        nb=int(data[name + "/dims"][0]), nv=int(data[name + "/dims"][1]), nq=int(data[name + "/dims"][2])).contiguous()
        d, keep, order = el.tree_desc_c(t)  # level-ordered tables, wrench rows in table order
        e = mb.Engine(_capi.TreeDesc.from_buffer_copy(bytes(d)), 0, keepalive=keep)
        e.set_gravity(*data[name + "/gravity"])
        q, qd, qdd, tau, fext = (data["%s/%s" % (name, f)] for f in ("q", "qd", "qdd", "tau", "fext"))
        n = q.shape[1]
        fe = np.ascontiguousarray(fext.reshape(t.nb, 6, n)[order].reshape(6 * t.nb, n))
        tq, tqd, tqdd, ttau, tf = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (q, qd, qdd, tau, fe))
        out = torch.empty_like(tqd)
        assert rel(e.rnea(tq, tqd, tqdd, out, fext=tf).cpu().numpy(), data[name + "/rnea"]) < TOL
        assert rel(e.aba(tq, tqd, ttau, out, fext=tf).cpu().numpy(), data[name + "/aba"]) < TOL
        M = torch.empty((t.nv * t.nv, n), dtype=torch.float64, device=dev)
        assert rel(e.crba(tq, M).cpu().numpy().reshape(t.nv, t.nv, n), data[name + "/crba"]) < TOL
        e.close()


def test_baseline_config_2_a7_65536_states_per_state(torch_dev):
    """BASELINE.json configs[1]: batched RNEA, 7-DoF revolute arm, 65,536 random states, verified per state."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="chain", seed=65536, n_joints=7)
    n = 65536
    rng = np.random.default_rng(65536)
    q, qd, qdd, _ = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    ident = mb.InverseDynamicsCalculator(s)
    ident.setGravitationalAcceleration(-9.81)
    tau = ident.compute(*(torch.from_numpy(x).to(dev) for x in (q, qd, qdd))).cpu().numpy()
    ref = ol.Oracle(t, gravity=(0, 0, -9.81)).rnea_batch(q, qd, qdd)
    per_state = np.max(np.abs(tau - ref), axis=0) / np.maximum(1.0, np.max(np.abs(ref), axis=0))
    assert per_state.max() < TOL


@pytest.mark.parametrize("neck", [2, 1])
def test_baseline_configs_3_4_humanoid_1m_states_invariants(torch_dev, neck):
    """BASELINE.json configs[2..3]: ABA and CRBA on the humanoid at 1M states.  Checked on device through the reference's
    invariants (size-independent), and against the oracle on a strided sample of 512 states."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="humanoid", seed=99, n_joints=neck)
    nv, nq, n = t.nv, t.nq, 1 << 20
    gen = torch.Generator(device=dev).manual_seed(5)
    q = (torch.rand((nq, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * np.pi
    quat = torch.randn((4, n), dtype=torch.float64, device=dev, generator=gen)
    q[0:4] = quat / quat.norm(dim=0, keepdim=True)
    q[4:7] = torch.rand((3, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    qd = torch.rand((nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    qdd = torch.rand((nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    ident = mb.InverseDynamicsCalculator(s)
    fdyn = mb.ForwardDynamicsCalculator(s)
    crba = mb.CompositeRigidBodyMassMatrixCalculator(s)
    ident.setGravitationalAcceleration(-9.81)
    fdyn.setGravitationalAcceleration(-9.81)
    tau = ident.compute(q, qd, qdd)
    back = fdyn.compute(q, qd, tau)
    # FD(ID(qdd)) == qdd  (ForwardDynamicsCalculatorTest.java:223-251, tolerance 4e-11 scaled by max(1, ||expected||))
    scale = torch.clamp(qdd.norm(dim=0), min=1.0)
    assert float(((back - qdd).abs().max(dim=0).values / scale).max()) < 1e-9
    # M qdd + ID(qdd = 0) == ID(qdd)  (:904-1003), in chunks to bound memory
    ident.setConsiderJointAccelerations(False)
    bias = ident.compute(q, qd, qdd)
    chunk = 1 << 17
    worst = 0.0
    for a in range(0, n, chunk):
        M = crba.getMassMatrix(q[:, a:a + chunk].contiguous()).reshape(nv, nv, -1)
        lhs = torch.einsum("ijs,js->is", M, qdd[:, a:a + chunk]) + bias[:, a:a + chunk]
        sc = torch.clamp(tau[:, a:a + chunk].norm(dim=0), min=1.0)
        worst = max(worst, float(((lhs - tau[:, a:a + chunk]).abs().max(dim=0).values / sc).max()))
        assert float((M - M.transpose(0, 1)).abs().max()) == 0.0, "symmetric entries are written from one value"
    assert worst < 1e-9
    # strided sample against the oracle
    idx = torch.arange(0, n, n // 512, device=dev)
    hq, hqd, hqdd, htau, hback = (x[:, idx].cpu().numpy().copy() for x in (q, qd, qdd, tau, back))
    o = ol.Oracle(t, gravity=(0, 0, -9.81))
    assert rel(htau, o.rnea_batch(hq, hqd, hqdd)) < TOL
    assert rel(hback, o.aba_batch(hq, hqd, htau)) < TOL
    Ms = crba.getMassMatrix(q[:, idx].contiguous()).cpu().numpy().reshape(nv, nv, -1)
    assert rel(Ms, o.crba_batch(hq)) < TOL


def test_power_balance_on_device_1m_states(torch_dev):
    """A size-independent property that involves every kernel of the path and the integrator, at BASELINE's full batch: along the
    motion d/dt (qd^T M(q) qd / 2) = qd^T (tau - g(q)).  Forward dynamics (ABA) gives the accelerations, the state integrator the
    states at t +- h, the mass matrix (CRBA) their kinetic energy, inverse dynamics at rest (RNEA) the gravity efforts -- all on
    the device; the same check runs on the oracle in tests/test_oracle.py::test_power_balance_along_the_integrator."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="humanoid", seed=99, n_joints=2)
    nv, nq, n = t.nv, t.nq, 1 << 20
    gen = torch.Generator(device=dev).manual_seed(8)
    q = (torch.rand((nq, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * np.pi
    quat = torch.randn((4, n), dtype=torch.float64, device=dev, generator=gen)
    q[0:4] = quat / quat.norm(dim=0, keepdim=True)
    q[4:7] = torch.rand((3, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    qd = torch.rand((nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    tau = (torch.rand((nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * 10.0
    g = (0.3, -0.2, -9.81)
    ident = mb.InverseDynamicsCalculator(s)
    fdyn = mb.ForwardDynamicsCalculator(s)
    crba = mb.CompositeRigidBodyMassMatrixCalculator(s)
    ident.setGravitationalAcceleration(g)
    fdyn.setGravitationalAcceleration(g)
    qdd = fdyn.compute(q, qd, tau)
    zero = torch.zeros_like(qd)
    gravity_efforts = ident.compute(q, zero, zero)
    power = (qd * (tau - gravity_efforts)).sum(dim=0)
    # central differences with h = 1e-5: truncation (h^2 times the third derivative) and rounding (eps |T| / h) meet there; on the
    # oracle the worst of 1e5 such states is off by 5e-7 .. 9e-7 of max(1, |power|), the median by 1e-9 (h = 3e-5: 3e-6, 1e-6: 1.5e-6)
    h = 1.0e-5
    chunk = 1 << 17
    worst = 0.0
    medians = []
    integrators = {dt: mb.MultiBodySystemStateIntegrator(s, dt) for dt in (h, -h)}
    for a in range(0, n, chunk):
        kinetic = []
        for dt in (h, -h):
            q1, qd1, qdd1 = (x[:, a:a + chunk].clone(memory_format=torch.contiguous_format) for x in (q, qd, qdd))
            integrators[dt].doubleIntegrateFromAcceleration(q1, qd1, qdd1)
            M = crba.getMassMatrix(q1).reshape(nv, nv, -1)
            kinetic.append(0.5 * torch.einsum("is,ijs,js->s", qd1, M, qd1))
        numerical = (kinetic[0] - kinetic[1]) / (2.0 * h)
        p = power[a:a + chunk]
        e = (numerical - p).abs() / torch.clamp(p.abs(), min=1.0)
        worst = max(worst, float(e.max()))
        medians.append(float(e.median()))
    assert worst < 5.0e-5 and max(medians) < 1.0e-7, (worst, medians)


SPEC_CASES = [
    ("A7 revolute chain", dict(kind="chain", seed=1, n_joints=7)),
    ("prismatic chain", dict(kind="chain", seed=3, n_joints=6, prismatic=1.0)),
    ("floating + mixed chain 6", dict(kind="chain", seed=31, n_joints=6, floating=True, prismatic=0.3)),
    ("one-dof tree 15", dict(kind="tree", seed=32, n_joints=15, prismatic=0.2)),
]


@pytest.mark.parametrize("idx", range(len(SPEC_CASES)))
def test_tree_specialised_kernels_match_oracle(torch_dev, idx, tmp_path, monkeypatch):
    """mecano_b200_specialize(): kernels unrolled for one tree and compiled by NVRTC at run time, against the oracle and
    against the generic kernels, ragged batch; then again through the cubin cache."""
    import mecano_b200 as mb

    monkeypatch.setenv("MECANO_B200_CACHE", str(tmp_path))
    torch, dev = torch_dev
    name, kw = SPEC_CASES[idx]
    s, t = build(**kw)
    rng = np.random.default_rng(2000 + idx)
    g = (rng.uniform(-1, 1), rng.uniform(-1, 1), -rng.uniform(1, 10))
    o = ol.Oracle(t, gravity=g)
    n = 1000 + 13
    q, qd, qdd, tau = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    tq, tqd, tqdd, ttau = (torch.from_numpy(x).to(dev) for x in (q, qd, qdd, tau))
    for attempt in range(2):  # 0: compiled, 1: loaded from the cache
        ident = mb.InverseDynamicsCalculator(s)
        ident.setGravitationalAcceleration(g)
        generic = ident.compute(tq, tqd, tqdd).cpu().numpy()
        ident.specialize(force=True)
        info = ident.kernelInfo()
        assert info["specialized"] == 1 + attempt, (name, info)
        out = torch.full((t.nv, n + 19), float("nan"), dtype=torch.float64, device=dev)  # ld > n: the tail must stay NaN
        lead = [torch.empty((x.shape[0], n + 19), dtype=torch.float64, device=dev) for x in (tq, tqd, tqdd)]
        for dst, src in zip(lead, (tq, tqd, tqdd)):
            dst[:, :n] = src
        res = ident.compute(lead[0][:, :n], lead[1][:, :n], lead[2][:, :n], out[:, :n]).cpu().numpy()
        assert torch.isnan(out[:, n:]).all(), "specialised kernel wrote beyond n_states"
        assert rel(res, o.rnea_batch(q, qd, qdd)) < TOL, name
        assert rel(res, generic) < 1e-12, name
        # calls the specialised kernel does not cover fall back to the generic kernels of the same handle
        ident.setConsiderCoriolisAndCentrifugalForces(False)
        assert rel(ident.compute(tq, tqd, tqdd).cpu().numpy(), o.rnea_batch(q, qd, qdd, flags=1)) < TOL, name

        fdyn = mb.ForwardDynamicsCalculator(s)
        fdyn.setGravitationalAcceleration(*g)
        generic = fdyn.compute(tq, tqd, ttau).cpu().numpy()
        fdyn.specialize(force=True)
        assert fdyn.kernelInfo()["specialized"] == 1 + attempt, name
        res = fdyn.compute(tq, tqd, ttau).cpu().numpy()
        assert rel(res, o.aba_batch(q, qd, tau)) < TOL, name
        assert rel(res, generic) < 1e-10, name


def test_specialize_policy(torch_dev, tmp_path, monkeypatch):
    """Large trees keep the generic kernel (unrolled code would not fit the instruction caches) unless forced; CRBA is never
    specialised; small trees are."""
    import mecano_b200 as mb

    monkeypatch.setenv("MECANO_B200_CACHE", str(tmp_path))
    s, _ = build(kind="humanoid", seed=7, n_joints=2)
    ident = mb.InverseDynamicsCalculator(s).specialize()
    assert ident.kernelInfo()["specialized"] == 0
    crba = mb.CompositeRigidBodyMassMatrixCalculator(s).specialize(force=True)
    assert crba.kernelInfo()["specialized"] == 0
    s7, _ = build(kind="chain", seed=1, n_joints=7)
    assert mb.InverseDynamicsCalculator(s7).specialize().kernelInfo()["specialized"] >= 1
    assert mb.ForwardDynamicsCalculator(s7).specialize().kernelInfo()["specialized"] >= 1


@pytest.mark.parametrize("variant", ["thread", "warp"])
def test_calculator_owned_mass_matrix_skips_structural_zeros_only(torch_dev, variant):
    """getMassMatrix(q) without an output argument returns the calculator's own buffer (Mecano hands out a reference to its
    internal matrix, CompositeRigidBodyMassMatrixCalculator.java:344-348); from the second call on the structurally zero
    entries are neither rewritten nor re-transferred (MECANO_B200_CRBA_ZEROS_PRESENT).  Every call must still equal the
    oracle in full, on the device path and on the host path."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build("humanoid", 7, 2)
    o = ol.Oracle(t)
    nv = s.getNumberOfDoFs()
    n = 300
    rng = np.random.default_rng(77)
    crba = mb.CompositeRigidBodyMassMatrixCalculator(s).setKernelVariant(variant)
    first = None
    for it in range(3):
        q = mb.MultiBodySystemRandomTools.nextState(rng, s, n)[0]
        ref = o.crba_batch(q)
        M = crba.getMassMatrix(torch.from_numpy(q).to(dev))
        if first is None:
            first = M
        assert M.data_ptr() == first.data_ptr(), "the calculator-owned buffer is reused"
        assert rel(M.cpu().numpy().reshape(nv, nv, n), ref) <= TOL, (variant, it, "device")
    host = mb.CompositeRigidBodyMassMatrixCalculator(s).setKernelVariant(variant)
    for it in range(3):
        q = mb.MultiBodySystemRandomTools.nextState(rng, s, n)[0]
        ref = o.crba_batch(q)
        Mh = host.getMassMatrix(q)
        assert rel(Mh.reshape(nv, nv, n), ref) <= TOL, (variant, it, "host")
        if it == 1:
            Mh[:] = np.where(ref.reshape(nv * nv, n) == 0.0, Mh, np.nan)  # poison everything that must be refreshed
    # a caller-supplied matrix is always written in full
    mine = torch.full((nv * nv, n), float("nan"), dtype=torch.float64, device=dev)
    q = mb.MultiBodySystemRandomTools.nextState(rng, s, n)[0]
    crba.getMassMatrix(torch.from_numpy(q).to(dev), mine)
    assert rel(mine.cpu().numpy().reshape(nv, nv, n), o.crba_batch(q)) <= TOL


@pytest.mark.parametrize("idx", [0, 3, 5, 6, 8, 10, 11, 13, 14])
def test_rnea_byproducts_match_oracle(torch_dev, idx):
    """getBodyAcceleration(body) / getComputedJointWrench(joint) for N states (InverseDynamicsCalculator.java:578-602):
    mecano_b200_rnea_full on device buffers and mecano_b200_rnea_full_host on host buffers, with and without external
    wrenches, against the oracle state by state; the joint efforts of the same call must equal the plain kernel's."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    name, kw = CASES[idx]
    s, t = build(**kw)
    rng = np.random.default_rng(4000 + idx)
    g = (rng.uniform(-1, 1), rng.uniform(-1, 1), -rng.uniform(1, 10))
    o = ol.Oracle(t, gravity=g)
    n = 333
    q, qd, qdd, _ = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    fext = np.ascontiguousarray(rng.uniform(-1, 1, size=(6 * t.nb, n)))
    tq, tqd, tqdd, tf = (torch.from_numpy(x).to(dev) for x in (q, qd, qdd, fext))
    plain = mb.InverseDynamicsCalculator(s)
    plain.setGravitationalAcceleration(g)
    ident = mb.InverseDynamicsCalculator(s).setComputeByProducts()
    ident.setGravitationalAcceleration(g)
    for f_host, f_dev in ((None, None), (fext, tf)):
        plain.setExternalWrenches(f_dev)
        ident.setExternalWrenches(f_dev)
        tau = ident.compute(tq, tqd, tqdd).cpu().numpy()
        assert rel(tau, plain.compute(tq, tqd, tqdd).cpu().numpy()) < 1e-13, name
        acc = ident.getBodyAccelerationMatrix().cpu().numpy().reshape(t.nb, 6, n)
        wr = ident.getComputedJointWrenchMatrix().cpu().numpy().reshape(t.nb, 6, n)
        for k in range(0, n, 11):
            fo = None if f_host is None else np.ascontiguousarray(f_host[:, k].reshape(t.nb, 6))
            tau_o, acc_o, wr_o = o.rnea_full(q[:, k], qd[:, k], qdd[:, k], fo)
            assert rel(tau[:, k], tau_o) < TOL, name
            assert rel(acc[:, :, k], acc_o) < TOL, name
            assert rel(wr[:, :, k], wr_o) < TOL, name
        # the per-object getters are views of the same matrices
        j = s.getAllJoints()[t.nb // 2]
        assert torch.equal(ident.getComputedJointWrench(j), ident.getComputedJointWrenchMatrix()[6 * (t.nb // 2):6 * (t.nb // 2) + 6])
        assert torch.equal(ident.getBodyAcceleration(j.getSuccessor()), ident.getBodyAccelerationMatrix()[6 * (t.nb // 2):6 * (t.nb // 2) + 6])
        assert torch.equal(ident.getComputedJointTau(j), ident.getJointTauMatrix()[t.dof_off[t.nb // 2]:t.dof_off[t.nb // 2] + j.getDegreesOfFreedom()])
        # host path: same kernel behind pinned staging
        ident.setExternalWrenches(f_host)
        tau_h = ident.compute(q, qd, qdd)
        assert rel(tau_h, tau) == 0.0 and rel(ident.getBodyAccelerationMatrix().reshape(t.nb, 6, n), acc) == 0.0
        assert rel(ident.getComputedJointWrenchMatrix().reshape(t.nb, 6, n), wr) == 0.0
    # only one of the two by-products
    only = mb.InverseDynamicsCalculator(s).setComputeByProducts(bodyAccelerations=False)
    only.setGravitationalAcceleration(g)
    only.setExternalWrenches(tf)
    only.compute(tq, tqd, tqdd)
    assert only.getBodyAccelerationMatrix() is None
    assert rel(only.getComputedJointWrenchMatrix().cpu().numpy().reshape(t.nb, 6, n), wr) == 0.0


@pytest.mark.parametrize("idx", [0, 2, 3, 5, 6, 8, 10, 11, 13, 14])
def test_forward_dynamics_joint_source_modes(torch_dev, idx):
    """ForwardDynamicsCalculator with joints in JointSourceMode.ACCELERATION_SOURCE (ForwardDynamicsCalculator.java:400-444,
    :1237-1253, :1286-1298, pass four :1315-1363): against the oracle state by state, and through the reference's own invariant
    (ForwardDynamicsCalculatorTest.testJointMixedSourceModeGeneral :389-488: FD with a random subset of joints locked returns the
    accelerations and the efforts of ID) on the whole batch."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    name, kw = CASES[idx]
    s, t = build(**kw)
    rng = np.random.default_rng(5000 + idx)
    g = (0.0, 0.0, -9.81)
    o = ol.Oracle(t, gravity=g)
    n = 1500
    q, qd, qdd, _ = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    fext = np.ascontiguousarray(rng.uniform(-1, 1, size=(6 * t.nb, n)))
    tq, tqd, tqdd, tf = (torch.from_numpy(x).to(dev) for x in (q, qd, qdd, fext))
    ident = mb.InverseDynamicsCalculator(s)
    ident.setGravitationalAcceleration(g)
    fdyn = mb.ForwardDynamicsCalculator(s)
    fdyn.setGravitationalAcceleration(g)
    joints = s.getAllJoints()
    for trial, f_host, f_dev in ((0, None, None), (1, fext, tf), (2, None, None)):
        locked = np.zeros(t.nb, np.int32)
        locked[rng.permutation(t.nb)[: 1 if t.nb == 1 else rng.integers(1, t.nb)]] = 1
        if trial == 2:
            locked[:] = 1  # every joint given: forward dynamics degenerates to inverse dynamics
        fdyn.resetJointSourceModes()
        fdyn.setJointSourceModes(lambda j: mb.JointSourceMode.ACCELERATION_SOURCE if locked[joints.index(j)] else None)
        assert [fdyn.getJointSourceMode(j) == mb.JointSourceMode.ACCELERATION_SOURCE for j in joints] == [bool(v) for v in locked]
        ident.setExternalWrenches(f_dev)
        fdyn.setExternalWrenches(f_dev)
        tau = ident.compute(tq, tqd, tqdd)
        tau_in = tau.clone()
        for b in np.nonzero(locked)[0]:
            nd = 6 if t.jtype[b] == td.SIXDOF else 1
            tau_in[t.dof_off[b]:t.dof_off[b] + nd] = float("nan")  # the efforts of locked joints must not be read
        with pytest.raises(mb.MatrixDimensionException):
            fdyn.compute(tq, tqd, tau_in)  # locked joints need their accelerations
        got_qdd = fdyn.compute(tq, tqd, tau_in, jointAccelerationInput=tqdd).cpu().numpy()
        got_tau = fdyn.getJointTauMatrix().cpu().numpy()
        assert not (np.isnan(got_qdd).any() or np.isnan(got_tau).any()), name
        jl = joints[-1]
        rl = slice(t.dof_off[t.nb - 1], t.dof_off[t.nb - 1] + jl.getDegreesOfFreedom())
        assert torch.equal(fdyn.getComputedJointAcceleration(jl), fdyn.getJointAccelerationMatrix()[rl])
        assert torch.equal(fdyn.getJointTau(jl), fdyn.getJointTauMatrix()[rl])
        tol = 4e-11 if (t.jtype == td.SIXDOF).any() else 1.6e-11  # ForwardDynamicsCalculatorTest.java:38-41
        assert rel(got_qdd, qdd) < tol * max(1, t.nb / 10), name
        assert rel(got_tau, tau.cpu().numpy()) < tol * max(1, t.nb / 10), name
        tau_np = tau_in.cpu().numpy()
        for k in range(0, n, 97):
            fo = None if f_host is None else np.ascontiguousarray(f_host[:, k].reshape(t.nb, 6))
            tin = np.nan_to_num(tau_np[:, k])
            want_qdd, want_tau = o.aba_sources(q[:, k], qd[:, k], tin, qdd[:, k], locked, fo)
            assert rel(got_qdd[:, k], want_qdd) < TOL, name
            assert rel(got_tau[:, k], want_tau) < TOL, name
        # host path
        fdyn.setExternalWrenches(f_host)
        h_qdd = fdyn.compute(q, qd, np.nan_to_num(tau_np), jointAccelerationInput=qdd)
        assert rel(h_qdd, got_qdd) < 1e-13 and rel(fdyn.getJointTauMatrix(), got_tau) < 1e-13, name
    # back to the default: the plain kernel, bit for bit
    fdyn.resetJointSourceModes()
    fdyn.setExternalWrenchesToZero()
    ident.setExternalWrenchesToZero()
    tau = ident.compute(tq, tqd, tqdd)
    plain = mb.ForwardDynamicsCalculator(s)
    plain.setGravitationalAcceleration(g)
    assert torch.equal(fdyn.compute(tq, tqd, tau), plain.compute(tq, tqd, tau))
    assert fdyn.getJointTauMatrix() is tau


@pytest.mark.parametrize("idx", [0, 3, 5, 6, 8, 9, 10, 11, 13, 14])
def test_centroidal_momentum_matrix_and_convective_term(torch_dev, idx):
    """getCentroidalMomentumMatrix() / getCentroidalConvectiveTerm() of the mass-matrix calculator
    (CompositeRigidBodyMassMatrixCalculator.java:380-440, :801-839) for N states, in the world frame and in the centre-of-mass
    frame: against the oracle state by state, and on the whole batch through the facts the reference tests
    (CompositeRigidBodyMassMatrixCalculatorTest.java:62-82): A qd = momentum, A qdd + convective term = momentum rate = the
    wrench at the root of inverse dynamics without gravity."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    name, kw = CASES[idx]
    s, t = build(**kw)
    rng = np.random.default_rng(6000 + idx)
    o = ol.Oracle(t)
    n = 1100
    nv = t.nv
    q, qd, qdd, _ = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    tq, tqd = torch.from_numpy(q).to(dev), torch.from_numpy(qd).to(dev)
    plain = mb.CompositeRigidBodyMassMatrixCalculator(s).setKernelVariant("thread")  # the by-products run thread-per-state
    M_plain = plain.getMassMatrix(tq, torch.empty((nv * nv, n), dtype=torch.float64, device=dev)).cpu().numpy()
    for frame_name, frame in (("worldFrame", 0), ("centerOfMassFrame", 1)):
        calc = mb.CompositeRigidBodyMassMatrixCalculator(s, frame_name)
        assert calc.getCentroidalMomentumFrame() == frame_name
        A = calc.getCentroidalMomentumMatrix(tq).cpu().numpy().reshape(6, nv, n)
        assert np.array_equal(calc.getMassMatrix().cpu().numpy(), M_plain), "the by-products must not change the mass matrix"
        com = calc.getCenterOfMass().cpu().numpy()
        b = calc.getCentroidalConvectiveTermMatrix(tq, tqd).cpu().numpy()
        assert not (np.isnan(A).any() or np.isnan(com).any() or np.isnan(b).any()), name
        # the convective term took its centre of mass from the centre-of-mass-only launch (mecano_b200_center_of_mass: the
        # by-product kernel without a matrix): the same rows, bit for bit, as the full launch left
        assert calc.getCenterOfMass() is not None and np.array_equal(calc.getCenterOfMass().cpu().numpy(), com), name
        assert np.array_equal(mb.CompositeRigidBodyMassMatrixCalculator(s, frame_name).getCenterOfMass(tq).cpu().numpy(), com), name
        for k in range(0, n, 53):
            _, Ao, co, mo = o.crba_centroidal(q[:, k], frame)
            assert rel(A[:, :, k], Ao) < TOL, name
            assert rel(com[:3, k], co) < TOL and abs(com[3, k] - mo) < TOL * mo, name
            assert rel(b[:, k], o.centroidal_convective_term(q[:, k], qd[:, k], frame)) < TOL, name
        # host path: same kernels behind plain staging
        hc = mb.CompositeRigidBodyMassMatrixCalculator(s, frame_name)
        Ah = hc.getCentroidalMomentumMatrix(q).reshape(6, nv, n)
        bh = hc.getCentroidalConvectiveTermMatrix(q, qd)
        assert rel(Ah, A) == 0.0 and rel(bh, b) == 0.0 and rel(hc.getMassMatrix(), M_plain) == 0.0, name
        assert np.array_equal(hc.getCenterOfMass(q), com), name
    # batch-wide invariant in the world frame: A qdd + b = sum over the root's children of their joint wrench, in the world frame,
    # which for a single floating root joint is the pelvis wrench rotated / shifted by the pelvis pose; checked through linear
    # momentum only (frame-origin independent): total force = mass * CoM acceleration = rows 3..5
    calc = mb.CompositeRigidBodyMassMatrixCalculator(s)
    A = calc.getCentroidalMomentumMatrix(tq).cpu().numpy().reshape(6, nv, n)
    b = calc.getCentroidalConvectiveTermMatrix(tq, tqd).cpu().numpy()
    hdot = np.einsum("rjs,js->rs", A, qdd) + b
    ident = mb.InverseDynamicsCalculator(s).setComputeByProducts(bodyAccelerations=False)
    ident.compute(tq, tqd, torch.from_numpy(qdd).to(dev))  # zero gravity (the calculator's default)
    wr = ident.getComputedJointWrenchMatrix().cpu().numpy().reshape(t.nb, 6, n)
    for k in range(0, n, 211):
        total = np.zeros(3)
        for body in range(t.nb):
            if t.parent[body] < 0:
                # rotation frameAfterJoint -> world of a child of the root body
                qi = q[t.cfg_off[body]:, k]
                if t.jtype[body] in (td.SIXDOF, td.SPHERICAL):
                    x, y, z, w = qi[:4] / np.linalg.norm(qi[:4])
                    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
                elif t.jtype[body] == td.REVOLUTE:
                    u, a = t.axis[body] / np.linalg.norm(t.axis[body]), qi[0]
                    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
                    R = np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K
                elif t.jtype[body] == td.PLANAR:  # rotation about y by the pitch
                    R = np.array([[np.cos(qi[0]), 0, np.sin(qi[0])], [0, 1, 0], [-np.sin(qi[0]), 0, np.cos(qi[0])]])
                else:
                    R = np.eye(3)
                total += t.off_R[body] @ R @ wr[body, 3:, k]
        assert np.max(np.abs(hdot[3:, k] - total)) < 1e-9 * max(1.0, np.max(np.abs(total))), name


@pytest.mark.parametrize("idx", [0, 2, 3, 5, 6, 8, 9, 10, 11, 13, 14])
def test_coriolis_matrix(torch_dev, idx):
    """getCoriolisMatrix() (CompositeRigidBodyMassMatrixCalculator.java:278-281, :358-366, :588-799) for N states: against the
    oracle state by state, and on the whole batch through the reference's own test
    (CompositeRigidBodyMassMatrixCalculatorTest.testCoriolisMatrix :85-138): C qd = inverse dynamics without joint accelerations,
    at its 1e-11."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    name, kw = CASES[idx]
    s, t = build(**kw)
    rng = np.random.default_rng(7000 + idx)
    o = ol.Oracle(t)
    n = 900
    nv = t.nv
    q, qd, _, _ = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    tq, tqd = torch.from_numpy(q).to(dev), torch.from_numpy(qd).to(dev)
    calc = mb.CompositeRigidBodyMassMatrixCalculator(s)
    with pytest.raises(RuntimeError):
        calc.getCoriolisMatrix(tq, tqd)  # disabled by default (UnsupportedOperationException in the reference)
    calc.setEnableCoriolisMatrixCalculation(True)
    C = calc.getCoriolisMatrix(tq, tqd).cpu().numpy().reshape(nv, nv, n)
    M = calc.getMassMatrix().cpu().numpy().reshape(nv, nv, n)
    assert not (np.isnan(C).any() or np.isnan(M).any()), "every entry of both dense matrices must be written"
    plain = mb.CompositeRigidBodyMassMatrixCalculator(s).setKernelVariant("thread")
    assert rel(M, plain.getMassMatrix(tq).cpu().numpy().reshape(nv, nv, n)) < 1e-13, name
    for k in range(0, n, 41):
        Mo, Co = o.coriolis(q[:, k], qd[:, k])
        assert rel(C[:, :, k], Co) < TOL and rel(M[:, :, k], Mo) < TOL, name
    ident = mb.InverseDynamicsCalculator(s)  # zero gravity, the calculators' default
    ident.setConsiderJointAccelerations(False)
    want = ident.compute(tq, tqd, tqd).cpu().numpy()
    got = np.einsum("ijs,js->is", C, qd)
    assert np.max(np.abs(got - want)) < 1e-11 * max(1.0, np.max(np.abs(want))) * max(1, t.nb / 10), name
    # host path
    hc = mb.CompositeRigidBodyMassMatrixCalculator(s)
    hc.setEnableCoriolisMatrixCalculation(True)
    assert rel(hc.getCoriolisMatrix(q, qd).reshape(nv, nv, n), C) == 0.0 and rel(hc.getMassMatrix().reshape(nv, nv, n), M) == 0.0, name


def test_fp32_variant(torch_dev):
    """The optional fp32 variant (mecano_b200_set_precision): reported separately with its own tolerance (north star), never
    chosen implicitly, and refusing what it does not cover instead of running in fp64."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="humanoid", seed=7, n_joints=2)
    rng = np.random.default_rng(8000)
    o = ol.Oracle(t, gravity=(0.0, 0.0, -9.81))
    n = 3000
    nv = t.nv
    q, qd, qdd, tau = mb.MultiBodySystemRandomTools.nextState(rng, s, n)
    tq, tqd, tqdd, ttau = (torch.from_numpy(x).to(dev) for x in (q, qd, qdd, tau))
    ident = mb.InverseDynamicsCalculator(s).setKernelVariant("thread").setPrecision("fp32")
    ident.setGravitationalAcceleration(-9.81)
    fdyn = mb.ForwardDynamicsCalculator(s).setKernelVariant("thread").setPrecision("fp32")
    fdyn.setGravitationalAcceleration(-9.81)
    crba = mb.CompositeRigidBodyMassMatrixCalculator(s).setKernelVariant("thread").setPrecision("fp32")
    e_rnea = rel(ident.compute(tq, tqd, tqdd).cpu().numpy(), o.rnea_batch(q, qd, qdd))
    e_aba = rel(fdyn.compute(tq, tqd, ttau).cpu().numpy(), o.aba_batch(q, qd, tau))
    M = crba.getMassMatrix(tq, torch.full((nv * nv, n), float("nan"), dtype=torch.float64, device=dev))
    assert not torch.isnan(M).any()
    e_crba = rel(M.cpu().numpy().reshape(nv, nv, n), o.crba_batch(q))
    print("fp32 variant, H37, relative to the oracle: rnea %.2e aba %.2e crba %.2e" % (e_rnea, e_aba, e_crba))
    # tolerances a decade above the worst state of 2^20 (bench.py extras.fp32_variant: 1.7e-6 / 2.7e-6 / 5.7e-7)
    assert 1e-9 < e_rnea < 1e-5, "fp32 arithmetic must actually run (and stay within its tolerance)"
    assert 1e-9 < e_crba < 1e-5
    assert 1e-9 < e_aba < 1e-4
    # host entry points run the same kernels
    assert rel(ident.compute(q, qd, qdd), o.rnea_batch(q, qd, qdd)) < 1e-5
    # not covered: external wrenches, flags, by-products, trees outside the compiled configuration -> an error, not fp64
    ident.setExternalWrenches(torch.zeros((6 * t.nb, n), dtype=torch.float64, device=dev))
    with pytest.raises(mb.MecanoB200Error):
        ident.compute(tq, tqd, tqdd)
    ident.setExternalWrenchesToZero()
    ident.setConsiderJointAccelerations(False)
    with pytest.raises(mb.MecanoB200Error):
        ident.compute(tq, tqd, tqdd)
    big, _ = build(kind="tree", seed=9, n_joints=100, floating=True)
    qb, qdb, qddb, _ = mb.MultiBodySystemRandomTools.nextState(rng, big, 64)
    with pytest.raises(mb.MecanoB200Error):
        mb.InverseDynamicsCalculator(big).setKernelVariant("thread").setPrecision("fp32").compute(*(torch.from_numpy(x).to(dev) for x in (qb, qdb, qddb)))
    # back to fp64: bit-identical to a calculator that never left it
    ident.setConsiderJointAccelerations(True)
    ident.setPrecision("fp64")
    ref = mb.InverseDynamicsCalculator(s).setKernelVariant("thread")
    ref.setGravitationalAcceleration(-9.81)
    assert torch.equal(ident.compute(tq, tqd, tqdd), ref.compute(tq, tqd, tqdd))


def test_new_entry_points_with_leading_dimension(torch_dev):
    """The by-product / source-mode / centroidal / Coriolis entry points on matrices with ld > n (views of wider buffers) and a
    ragged n: same results as on compact matrices."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="tree", seed=31, n_joints=18, floating=True, prismatic=0.3)
    rng = np.random.default_rng(31)
    n, ld = 1237, 1536
    q, qd, qdd, tau = (torch.from_numpy(x).to(dev) for x in mb.MultiBodySystemRandomTools.nextState(rng, s, ld))
    wide = lambda x: x[:, :n]  # noqa: E731  (stride ld)
    compact = lambda x: x[:, :n].contiguous()  # noqa: E731

    def run(view):
        out = []
        ident = mb.InverseDynamicsCalculator(s).setComputeByProducts()
        ident.setGravitationalAcceleration(-9.81)
        out.append(ident.compute(view(q), view(qd), view(qdd)))
        out.append(ident.getBodyAccelerationMatrix())
        out.append(ident.getComputedJointWrenchMatrix())
        fdyn = mb.ForwardDynamicsCalculator(s)
        fdyn.setGravitationalAcceleration(-9.81)
        joints = s.getAllJoints()
        fdyn.setJointSourceModes(lambda j: mb.JointSourceMode.ACCELERATION_SOURCE if joints.index(j) % 3 == 0 else None)
        out.append(fdyn.compute(view(q), view(qd), view(tau), jointAccelerationInput=view(qdd)))
        out.append(fdyn.getJointTauMatrix())
        crba = mb.CompositeRigidBodyMassMatrixCalculator(s, "centerOfMassFrame")
        out.append(crba.getCentroidalMomentumMatrix(view(q)))
        out.append(crba.getCenterOfMass())
        out.append(crba.getCentroidalConvectiveTermMatrix(view(q), view(qd)))
        crba.setEnableCoriolisMatrixCalculation(True)
        out.append(crba.getCoriolisMatrix(view(q), view(qd)))
        out.append(crba.getMassMatrix())
        return out

    a, b = run(wide), run(compact)
    for x, y in zip(a, b):
        assert x.shape == y.shape and x.shape[1] == n
        assert torch.equal(x, y)
    assert a[0].stride(0) == ld and b[0].stride(0) == n


def test_humanoid_1m_states_invariants_of_the_next_rows(torch_dev):
    """The SURVEY 8f rows at BASELINE.json's full size (H37, 2^20 states), through size-independent properties checked on the
    device: forward dynamics with half of the joints locked returns the accelerations and efforts of inverse dynamics
    (ForwardDynamicsCalculatorTest.java:389-488); C(q, qd) qd = inverse dynamics without joint accelerations
    (CompositeRigidBodyMassMatrixCalculatorTest.java:85-138); the linear rows of A qdd + Adot qd = the force the floating joint
    transmits, rotated into the world (Newton's law for the whole system, :62-82); the joint wrench of the floating joint
    projects onto its six efforts."""
    import mecano_b200 as mb

    torch, dev = torch_dev
    s, t = build(kind="humanoid", seed=98, n_joints=2)
    nv, nq, n = t.nv, t.nq, 1 << 20
    gen = torch.Generator(device=dev).manual_seed(6)
    q = (torch.rand((nq, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * np.pi
    quat = torch.randn((4, n), dtype=torch.float64, device=dev, generator=gen)
    q[0:4] = quat / quat.norm(dim=0, keepdim=True)
    q[4:7] = torch.rand((3, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    qd = torch.rand((nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    qdd = torch.rand((nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1

    def worst(actual, expected):
        sc = torch.clamp(expected.abs().amax(dim=0), min=1.0)
        return float(((actual - expected).abs().amax(dim=0) / sc).max())

    # ---- mixed joint source modes
    ident = mb.InverseDynamicsCalculator(s).setComputeByProducts(bodyAccelerations=False)
    ident.setGravitationalAcceleration(-9.81)
    tau = ident.compute(q, qd, qdd)
    wr = ident.getComputedJointWrenchMatrix()
    assert worst(wr[0:6], tau[0:6]) < 1e-12, "SixDoF joint: S = identity, the efforts are the joint wrench"
    joints = s.getAllJoints()
    locked_rows = [i for i, j in enumerate(joints) if j.getDegreesOfFreedom() == 1 and i % 2 == 0]
    fdyn = mb.ForwardDynamicsCalculator(s)
    fdyn.setGravitationalAcceleration(-9.81)
    fdyn.setJointSourceModes(lambda j: mb.JointSourceMode.ACCELERATION_SOURCE if joints.index(j) in locked_rows else None)
    tau_in = tau.clone()
    for i in locked_rows:
        tau_in[t.dof_off[i]] = float("nan")
    back = fdyn.compute(q, qd, tau_in, jointAccelerationInput=qdd)
    assert worst(back, qdd) < 1e-9
    assert worst(fdyn.getJointTauMatrix(), tau) < 1e-9
    del back, tau_in, wr
    # ---- centroidal quantities, world frame: total force = R_pelvis * (force rows of the floating joint's efforts), gravity off
    ident0 = mb.InverseDynamicsCalculator(s)
    tau0 = ident0.compute(q, qd, qdd)
    cen = mb.CompositeRigidBodyMassMatrixCalculator(s)
    A = cen.getCentroidalMomentumMatrix(q).reshape(6, nv, n)
    b = cen.getCentroidalConvectiveTermMatrix(q, qd)
    hdot_lin = torch.einsum("rjs,js->rs", A[3:], qdd) + b[3:]
    x, y, z, w = q[0], q[1], q[2], q[3]
    R = torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)]),
                     torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)]),
                     torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)])])  # [3, 3, n]
    force_world = torch.einsum("ijs,js->is", R, tau0[3:6])
    assert worst(hdot_lin, force_world) < 1e-9
    mass = cen.getCenterOfMass()[3]
    assert float((mass - mass[0]).abs().max()) < 1e-12 * float(mass[0]), "the total mass does not depend on the state"
    del A, b, hdot_lin, R, force_world
    # ---- Coriolis matrix, in chunks to bound memory
    ident0.setConsiderJointAccelerations(False)
    want = ident0.compute(q, qd, qdd)
    cen.setEnableCoriolisMatrixCalculation(True)
    chunk = 1 << 17
    for a in range(0, n, chunk):
        C = cen.getCoriolisMatrix(q[:, a:a + chunk].contiguous(), qd[:, a:a + chunk].contiguous()).reshape(nv, nv, -1)
        got = torch.einsum("ijs,js->is", C, qd[:, a:a + chunk])
        assert worst(got, want[:, a:a + chunk]) < 1e-9


def test_golden_fixtures_of_the_next_rows_on_gpu(torch_dev):
    """tests/golden/oracle_golden_next.npz (RNEA by-products, joint source modes, centroidal quantities, Coriolis matrix) through
    the raw C ABI; body rows in the order of the level-ordered tables (wrench_index = NULL)."""
    import importlib.util

    import mecano_b200 as mb
    from mecano_b200 import _capi

    import emu_lib as el

    torch, dev = torch_dev
    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_next", os.path.join(here, "make_golden_next.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    data = np.load(os.path.join(here, "oracle_golden_next.npz"))
    for name, t, g, s in gen.load_cases():
        d, keep, order = el.tree_desc_c(t)
        e = mb.Engine(_capi.TreeDesc.from_buffer_copy(bytes(d)), 0, keepalive=keep)
        n, nv, nb = s["q"].shape[1], t.nv, t.nb
        back = np.empty(nb, dtype=np.int64)
        back[order] = np.arange(nb)
        fe = np.ascontiguousarray(s["fext"].reshape(nb, 6, n)[order].reshape(6 * nb, n))
        tq, tqd, tqdd, ttau, tf = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (s["q"], s["qd"], s["qdd"], s["tau"], fe))
        new = lambda rows: torch.full((rows, n), float("nan"), dtype=torch.float64, device=dev)  # noqa: E731
        # RNEA by-products
        e.set_gravity(*g)
        acc, wr = new(6 * nb), new(6 * nb)
        e.rnea(tq, tqd, tqdd, new(nv), fext=tf, body_acc=acc, joint_wrench=wr)
        assert rel(acc.cpu().numpy().reshape(nb, 6, n)[back], data[name + "/acc"]) < TOL, name
        assert rel(wr.cpu().numpy().reshape(nb, 6, n)[back], data[name + "/wr"]) < TOL, name
        # joint source modes
        e.set_joint_source_modes(np.asarray(data[name + "/accel_source"])[order])
        qdd_out, tau_out = new(nv), new(nv)
        e.aba_sources(tq, tqd, ttau, tqdd, qdd_out, tau_out, fext=tf)
        assert rel(qdd_out.cpu().numpy(), data[name + "/qdd_src"]) < TOL, name
        assert rel(tau_out.cpu().numpy(), data[name + "/tau_src"]) < TOL, name
        e.set_joint_source_modes(None)
        # centroidal quantities and the Coriolis matrix (no gravity involved)
        for frame, key in ((0, "world"), (1, "com")):
            M, cmm, com, conv = new(nv * nv), new(6 * nv), new(4), new(6)
            e.crba_centroidal(tq, M, cmm, com, frame)
            e.centroidal_convective_term(tq, tqd, com, conv, frame)
            assert rel(cmm.cpu().numpy().reshape(6, nv, n), data["%s/cmm_%s" % (name, key)]) < TOL, name
            assert rel(com.cpu().numpy(), data[name + "/com"]) < TOL, name
            assert rel(conv.cpu().numpy(), data["%s/conv_%s" % (name, key)]) < TOL, name
        M, C = new(nv * nv), new(nv * nv)
        e.coriolis(tq, tqd, M, C)
        assert rel(C.cpu().numpy().reshape(nv, nv, n), data[name + "/coriolis"]) < TOL, name
        e.close()
