"""CPU tests of the kernel mathematics: the per-state algorithm templates that the sm_100a kernels inline
(mecano_b200/csrc/algorithms.cuh) are compiled for the host by tests/emu and compared with the oracle.  This is test
infrastructure: the product library has no host compute path."""
import numpy as np
import pytest

import emu_lib as el
import oracle_lib as ol
import treedesc as td

TOL = 1e-9  # north_star: <= 1e-9 relative error in fp64


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))


def trees(rng):
    return [
        td.chain(rng, 1), td.chain(rng, 7), td.chain(rng, 6, prismatic_fraction=1.0), td.random_tree(rng, 30, prismatic_fraction=0.4),
        td.chain(rng, 0, floating=True), td.chain(rng, 15, floating=True), td.random_tree(rng, 45, floating=True, com_rotation=True, prismatic_fraction=0.2),
        td.random_tree(rng, 12, floating=True, axis_aligned=True), td.humanoid(rng, 2), td.humanoid(rng, 1), td.random_tree(rng, 100, floating=True),
        td.chain(rng, 60),
        # SphericalJoint / PlanarJoint (multidof.cuh): chains of one type, mixed trees, a floating base above them
        td.mixed_chain(rng, [td.SPHERICAL] * 5), td.mixed_chain(rng, [td.PLANAR, td.REVOLUTE, td.PLANAR, td.PRISMATIC, td.SPHERICAL]),
        td.mixed_tree(rng, 14, weights=(2, 1, 1, 2, 2), com_rotation=True), td.mixed_tree(rng, 30, floating=True),
    ]


N_TREES = 16


@pytest.mark.parametrize("idx", range(N_TREES))
def test_emulated_kernels_match_oracle(idx):
    rng = np.random.default_rng(300 + idx)
    t = trees(rng)[idx]
    g = (rng.uniform(-1, 1), rng.uniform(-1, 1), -rng.uniform(1, 10))
    o, e = ol.Oracle(t, gravity=g), el.Emu(t, gravity=g)
    n = 6
    q, qd, qdd, tau = td.random_states(rng, t, n)
    fext = rng.uniform(-1, 1, size=(t.nb, 6, n))
    fo = np.ascontiguousarray(fext.reshape(6 * t.nb, n))
    assert rel(e.rnea(q, qd, qdd), o.rnea_batch(q, qd, qdd)) < TOL
    assert rel(e.rnea(q, qd, qdd, fext), o.rnea_batch(q, qd, qdd, fo)) < TOL
    assert rel(e.rnea(q, qd, qdd, flags=1), o.rnea_batch(q, qd, qdd, flags=1)) < TOL
    assert rel(e.rnea(q, qd, qdd, flags=2), o.rnea_batch(q, qd, qdd, flags=2)) < TOL
    assert rel(e.rnea(q, qd, qdd, flags=3), o.rnea_batch(q, qd, qdd, flags=3)) < TOL
    assert rel(e.aba(q, qd, tau), o.aba_batch(q, qd, tau)) < TOL
    assert rel(e.aba(q, qd, tau, fext), o.aba_batch(q, qd, tau, fo)) < TOL
    M = e.crba(q)
    assert not np.isnan(M).any(), "every mass-matrix entry must be written (zeros included)"
    assert rel(M, o.crba_batch(q)) < TOL


@pytest.mark.parametrize("idx", [1, 3, 8, 13])
def test_emulated_kernels_with_angles_beyond_the_fast_sincos_range(idx):
    """Joint angles of any magnitude are legal in Mecano (Math.sin / cos reduce exactly).  The kernels evaluate a fast sin/cos for
    |q| < 1e5 and redo it with the library routine otherwise (jointmath.cuh: mb_angle_large); prismatic displacements of that
    size must pass through untouched."""
    rng = np.random.default_rng(4100 + idx)
    t = trees(rng)[idx]
    o, e = ol.Oracle(t), el.Emu(t)
    n = 8
    q, qd, qdd, tau = td.random_states(rng, t, n)
    rev = [int(t.cfg_off[i]) for i in range(t.nb) if t.jtype[i] == td.REVOLUTE]
    pris = [int(t.cfg_off[i]) for i in range(t.nb) if t.jtype[i] == td.PRISMATIC]
    for s in range(n):
        for r in rng.permutation(rev)[: 1 + s % 3]:
            q[r, s] = rng.choice([-1.0, 1.0]) * 10.0 ** rng.uniform(5.0, 9.0)
    q[rev[0], 0] = 1.0e5  # the boundary itself
    assert rel(e.rnea(q, qd, qdd), o.rnea_batch(q, qd, qdd)) < TOL
    assert rel(e.aba(q, qd, tau), o.aba_batch(q, qd, tau)) < TOL
    assert rel(e.crba(q), o.crba_batch(q)) < TOL
    if pris:
        # a displacement of that size is not an angle (forward dynamics is too ill-conditioned there to compare at 1e-9)
        for s in range(n):
            q[rng.choice(pris), s] = rng.choice([-1.0, 1.0]) * 10.0 ** rng.uniform(5.0, 6.0)
        assert rel(e.rnea(q, qd, qdd), o.rnea_batch(q, qd, qdd)) < TOL
        assert rel(e.crba(q), o.crba_batch(q)) < TOL


@pytest.mark.parametrize("idx", range(N_TREES))
def test_emulated_rnea_byproducts_match_oracle(idx):
    """getBodyAcceleration / getComputedJointWrench (InverseDynamicsCalculator.java:578-602): the kernel routines leave them
    in the frames the reference returns them in (CoM frame, frameAfterJoint)."""
    rng = np.random.default_rng(700 + idx)
    t = trees(rng)[idx]
    g = (rng.uniform(-1, 1), rng.uniform(-1, 1), -rng.uniform(1, 10))
    o, e = ol.Oracle(t, gravity=g), el.Emu(t, gravity=g)
    n = 3
    q, qd, qdd, _ = td.random_states(rng, t, n)
    fext = rng.uniform(-1, 1, size=(t.nb, 6, n))
    for f in (None, fext):
        tau, acc, wr = e.rnea_full(q, qd, qdd, f)
        assert not (np.isnan(tau).any() or np.isnan(acc).any() or np.isnan(wr).any())
        for s in range(n):
            fo = None if f is None else np.ascontiguousarray(f[:, :, s])
            tau_o, acc_o, wr_o = o.rnea_full(q[:, s], qd[:, s], qdd[:, s], fo)
            assert rel(tau[:, s], tau_o) < TOL
            assert rel(acc[:, :, s], acc_o) < TOL
            assert rel(wr[:, :, s], wr_o) < TOL


@pytest.mark.parametrize("idx", range(N_TREES))
def test_emulated_aba_source_modes_match_oracle(idx):
    """ForwardDynamicsCalculator with joints in ACCELERATION_SOURCE mode (ForwardDynamicsCalculator.java:1237-1253, :1286-1298)."""
    rng = np.random.default_rng(800 + idx)
    t = trees(rng)[idx]
    g = (rng.uniform(-1, 1), rng.uniform(-1, 1), -rng.uniform(1, 10))
    o, e = ol.Oracle(t, gravity=g), el.Emu(t, gravity=g)
    n = 4
    q, qd, qdd_in, tau = td.random_states(rng, t, n)
    fext = rng.uniform(-1, 1, size=(t.nb, 6, n))
    for trial in range(3):
        locked = np.zeros(t.nb, np.int32)
        locked[rng.permutation(t.nb)[: rng.integers(1, t.nb + 1)]] = 1
        if trial == 2:
            locked[:] = 1
        for f in (None, fext):
            got = e.aba_sources(q, qd, tau, qdd_in, locked, f)
            assert not np.isnan(got).any()
            for s in range(n):
                fo = None if f is None else np.ascontiguousarray(f[:, :, s])
                want, _ = o.aba_sources(q[:, s], qd[:, s], tau[:, s], qdd_in[:, s], locked, fo)
                assert rel(got[:, s], want) < TOL
    # no joint locked: the plain algorithm, bit for bit
    assert np.array_equal(e.aba_sources(q, qd, tau, qdd_in, np.zeros(t.nb, np.int32)), e.aba(q, qd, tau))


@pytest.mark.parametrize("idx", range(N_TREES))
def test_emulated_centroidal_byproducts_match_oracle(idx):
    """getCentroidalMomentumMatrix() / getCentroidalConvectiveTerm() (CompositeRigidBodyMassMatrixCalculator.java:801-839) in the
    root frame, as the kernels leave them before the centre-of-mass shift."""
    rng = np.random.default_rng(1200 + idx)
    t = trees(rng)[idx]
    o, e = ol.Oracle(t), el.Emu(t)
    n = 3
    q, qd, _, _ = td.random_states(rng, t, n)
    M, cmm, com = e.crba_centroidal(q)
    rw = e.rnea_root_wrench(q, qd)
    assert not (np.isnan(M).any() or np.isnan(cmm).any() or np.isnan(com).any() or np.isnan(rw).any())
    assert np.array_equal(M, e.crba(q)), "the by-products must not change the mass matrix"
    # the centre-of-mass-only launch (no matrix buffers: a store to either would fault in the emulator) gives the same rows
    assert np.array_equal(e.center_of_mass(q), com)
    for s in range(n):
        Mo, Ao, co, mo = o.crba_centroidal(q[:, s], 0)
        assert rel(cmm[:, :, s], Ao) < TOL
        assert rel(com[:3, s] / com[3, s], co) < TOL and abs(com[3, s] - mo) < TOL * mo
        assert rel(rw[:, s], o.centroidal_convective_term(q[:, s], qd[:, s], 0)) < TOL


@pytest.mark.parametrize("idx", range(N_TREES))
def test_emulated_coriolis_matrix_matches_oracle(idx):
    """getCoriolisMatrix() (CompositeRigidBodyMassMatrixCalculator.java:358-366, :588-799), with the mass matrix of the same
    recursion; every entry of both dense matrices written."""
    rng = np.random.default_rng(1400 + idx)
    t = trees(rng)[idx]
    o, e = ol.Oracle(t), el.Emu(t)
    n = 3
    q, qd, _, _ = td.random_states(rng, t, n)
    M, C = e.coriolis(q, qd)
    assert not (np.isnan(M).any() or np.isnan(C).any())
    assert rel(M, e.crba(q)) < 1e-13
    for s in range(n):
        Mo, Co = o.coriolis(q[:, s], qd[:, s])
        assert rel(M[:, :, s], Mo) < TOL
        assert rel(C[:, :, s], Co) < TOL


def test_table_order_does_not_matter():
    """The C-ABI accepts any topological listing of the bodies (level order from the Java host, or DFS)."""
    rng = np.random.default_rng(9)
    t = td.random_tree(rng, 20, floating=True, prismatic_fraction=0.3)
    q, qd, qdd, tau = td.random_states(rng, t, 4)
    a = el.Emu(t, level_ordered=True)
    b = el.Emu(t, level_ordered=False)
    assert np.array_equal(a.rnea(q, qd, qdd), b.rnea(q, qd, qdd))
    assert np.array_equal(a.aba(q, qd, tau), b.aba(q, qd, tau))
    assert np.array_equal(a.crba(q), b.crba(q))


def test_fp32_variant_tolerance():
    """The optional fp32 instantiation is reported separately with its own tolerance: measured 1e-7..5e-7 relative on H37 (worst
    of 2^20 states on the GPU: 1.7e-6 / 2.7e-6 / 5.7e-7), bounds a decade above that: RNEA 1e-5, ABA 1e-4, CRBA 1e-5."""
    rng = np.random.default_rng(10)
    t = td.humanoid(rng, 2)
    o, e = ol.Oracle(t), el.Emu(t, fp32=True)
    q, qd, qdd, tau = td.random_states(rng, t, 8)
    assert rel(e.rnea(q, qd, qdd), o.rnea_batch(q, qd, qdd)) < 1e-5
    assert rel(e.crba(q), o.crba_batch(q)) < 1e-5
    assert rel(e.aba(q, qd, tau), o.aba_batch(q, qd, tau)) < 1e-4


def test_stack_sizes_follow_depth_not_body_count():
    """The interleaved traversal keeps per-state data only for the current root-to-leaf path."""
    rng = np.random.default_rng(11)
    h37 = el.Emu(td.humanoid(rng, 2))
    big = el.Emu(td.random_tree(rng, 100, floating=True))
    # per level (doubles): RNEA wrench + sin/cos, ABA twist + sin/cos, CRBA sin/cos; leaves keep theirs in registers
    for algo, per_level in ((0, 8), (1, 8), (2, 2)):
        info = h37.program_info(algo)
        assert info["nops"] == 2 * info["nb"]
        assert info["stack"] <= per_level * (info["max_depth"] - 1) + 16
        assert big.program_info(algo)["stack"] <= per_level * big.program_info(algo)["max_depth"] + 16
    assert h37.program_info(0)["stack"] == 6 + 8 * 9  # pelvis wrench, then 3 spine + 6 non-leaf arm levels
    assert h37.program_info(1)["rec"] == 8 * 32  # pass-three records: four double2 per body (program.h: MB_ABA_REC)


def test_emulated_kernels_match_the_golden_fixtures_of_the_next_rows():
    """The kernel routines (compiled for the host) against tests/golden/oracle_golden_next.npz."""
    import importlib.util
    import os

    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_next", os.path.join(here, "make_golden_next.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    data = np.load(os.path.join(here, "oracle_golden_next.npz"))
    for name, t, g, s in gen.load_cases():
        e, e0 = el.Emu(t, gravity=g), el.Emu(t, gravity=(0.0, 0.0, 0.0))
        q, qd, qdd, tau = (np.ascontiguousarray(s[k]) for k in ("q", "qd", "qdd", "tau"))
        n = q.shape[1]
        fext = np.ascontiguousarray(s["fext"].reshape(t.nb, 6, n))
        _, acc, wr = e.rnea_full(q, qd, qdd, fext)
        assert rel(acc, data[name + "/acc"]) < TOL and rel(wr, data[name + "/wr"]) < TOL, name
        assert rel(e.aba_sources(q, qd, tau, qdd, data[name + "/accel_source"], fext), data[name + "/qdd_src"]) < TOL, name
        _, cmm, com = e0.crba_centroidal(q)
        assert rel(cmm, data[name + "/cmm_world"]) < TOL, name
        assert rel(com[:3] / com[3], data[name + "/com"][:3]) < TOL, name
        assert rel(e0.rnea_root_wrench(q, qd), data[name + "/conv_world"]) < TOL, name
        _, C = e0.coriolis(q, qd)
        assert rel(C, data[name + "/coriolis"]) < TOL, name


@pytest.mark.parametrize("kind", ["chain7", "humanoid", "tree40", "floating_chain", "two_floating"])
def test_packed_mass_matrix_layout(kind):
    """MECANO_B200_CRBA_PACKED: the packed instantiation of the CRBA routine writes every unique, structurally non-zero entry
    exactly once; scattered through the exported index map (both triangles) it is the dense matrix of the oracle, and what the
    map does not cover is structurally zero (CompositeRigidBodyMassMatrixCalculator.java:296, :700-707, :772-797)."""
    rng = np.random.default_rng(77)
    if kind == "chain7":
        t = td.chain(rng, 7)
    elif kind == "humanoid":
        t = td.humanoid(rng, 2)
    elif kind == "tree40":
        t = td.random_tree(rng, 40, floating=True, prismatic_fraction=0.3)
    elif kind == "floating_chain":
        t = td.chain(rng, 5, floating=True)
    else:  # two SixDoF joints in series: a 6 x 6 diagonal block whose column has SixDoF ancestors
        t = td.make_tree(rng, [-1, 0, 1, 1], [td.SIXDOF, td.SIXDOF, td.REVOLUTE, td.PRISMATIC])
    e, o = el.Emu(t), ol.Oracle(t)
    q = td.random_states(rng, t, 5)[0]
    row, col = e.packed_index()
    nv = t.nv
    assert len(set(zip(row.tolist(), col.tolist()))) == len(row), "an entry is listed twice"
    assert np.all(row <= col), "Mecano's depth-first DoF order puts ancestors first: the packed entries are the upper triangle"
    P = e.crba_packed(q)
    assert not np.isnan(P).any(), "a packed row was not written"
    M = np.zeros((nv, nv, 5))
    M[row, col] = P
    M[col, row] = P
    ref = o.crba_batch(q)
    assert rel(M, ref) < 1e-12
    dense = e.crba(q)
    assert np.array_equal(M, dense), "packed and dense instantiations must agree bit for bit"
    covered = np.zeros((nv, nv), bool)
    covered[row, col] = covered[col, row] = True
    assert np.all(ref[~covered] == 0.0)
    if kind == "humanoid":
        assert len(row) == 362  # 37-DoF humanoid: 362 of 1,369 entries


def test_fuzz_of_random_trees():
    """A fixed slice of scripts/fuzz_emulation.py (random trees of every joint type and shape, all emulated entry points against the
    oracle; on ill-conditioned trees the equation-of-motion residual decides): 150 seeds, a few seconds."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "fuzz_emulation.py")
    spec = importlib.util.spec_from_file_location("fuzz_emulation", path)
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    for seed in range(400000, 400150):
        rng = np.random.default_rng(seed)
        name, t = fz.random_case(rng)
        try:
            ok, _ = fz.check(rng, "%s seed %d" % (name, seed), t)
        except RuntimeError as ex:  # a tree the flattener refuses
            assert "rc=" in str(ex)
            continue
        assert ok, (name, seed)
