"""CPU tests of the host side: the model mirror (Mecano API names), the flattener's tables, index ordering, the
C-ABI library (loads, exports every declared symbol, validates arguments without a GPU) and the multi-GPU slicing."""
import ctypes
import os
import re

import numpy as np
import pytest

import mecano_b200 as mb
from mecano_b200 import _capi, multibody, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def small_system():
    elevator = mb.RigidBody("elevator")
    root = mb.SixDoFJoint("root", elevator)
    pelvis = mb.RigidBody("pelvis", root, np.diag([1.0, 2.0, 3.0]), 5.0, [0.1, 0.0, 0.0])
    hipL = mb.RevoluteJoint("hipL", pelvis, [0.0, 0.1, -0.2], [0.0, 1.0, 0.0])
    thighL = mb.RigidBody("thighL", hipL, (0.1, 0.2, 0.3), 2.0, [0.0, 0.0, -0.2])
    kneeL = mb.PrismaticJoint("kneeL", thighL, [0.0, 0.0, -0.4], [0.0, 0.0, 1.0])
    mb.RigidBody("shinL", kneeL, (0.1, 0.1, 0.05), 1.0, [0.0, 0.0, -0.2])
    hipR = mb.RevoluteJoint("hipR", pelvis, [0.0, -0.1, -0.2], [1.0, 1.0, 0.0])
    mb.RigidBody("thighR", hipR, (0.1, 0.2, 0.3), 2.0, [0.0, 0.0, -0.2])
    return elevator, mb.MultiBodySystem.toMultiBodySystemBasics(elevator)


def test_joint_order_is_depth_first_preorder():
    """JointIterator.java:153-162: depth-first, children in insertion order; JointMatrixIndexProvider.java:77-101."""
    _, s = small_system()
    assert [j.getName() for j in s.getJointsToConsider()] == ["root", "hipL", "kneeL", "hipR"]
    prov = s.getJointMatrixIndexProvider()
    j = {x.getName(): x for x in s.getJointsToConsider()}
    assert prov.getJointDoFIndices(j["root"]) == [0, 1, 2, 3, 4, 5]
    assert prov.getJointConfigurationIndices(j["root"]) == [0, 1, 2, 3, 4, 5, 6]  # SixDoFJointReadOnly.java:21-26
    assert prov.getJointDoFIndices(j["hipL"]) == [6] and prov.getJointConfigurationIndices(j["hipL"]) == [7]
    assert prov.getJointDoFIndices(j["hipR"]) == [8] and prov.getJointConfigurationIndices(j["hipR"]) == [9]
    assert s.getNumberOfDoFs() == 9 and s.getConfigurationMatrixSize() == 10


def test_level_ordered_tables():
    _, s = small_system()
    t = s.tables().contents
    assert t.struct_size == ctypes.sizeof(_capi.TreeDesc)
    assert (t.n_bodies, t.n_dofs, t.n_cfg, t.n_levels) == (4, 9, 10, 3)
    assert [t.level_start[i] for i in range(4)] == [0, 1, 3, 4]
    assert [t.parent[i] for i in range(4)] == [-1, 0, 0, 1]            # root | hipL hipR | kneeL
    assert [t.joint_type[i] for i in range(4)] == [2, 0, 0, 1]
    assert [t.dof_offset[i] for i in range(4)] == [0, 6, 8, 7]         # Mecano rows survive the re-ordering
    assert [t.wrench_index[i] for i in range(4)] == [0, 1, 3, 2]
    axis = np.array([t.axis[3 * 2 + k] for k in range(3)])
    assert np.allclose(axis, np.array([1.0, 1.0, 0.0]) / np.sqrt(2))   # axes are normalised like Mecano does
    assert np.allclose([t.offset_pos[3 * 3 + k] for k in range(3)], [0.0, 0.0, -0.4])


def test_generators_and_describe_roundtrip():
    e = mb.RigidBody("elevator")
    joints = mb.MultiBodySystemRandomTools.nextHumanoid(3, e, 2)
    s = mb.MultiBodySystem.toMultiBodySystemBasics(e)
    assert len(joints) == 32 and s.getNumberOfDoFs() == 37 and s.getConfigurationMatrixSize() == 38
    d = s.describe()
    assert d["jtype"][0] == 2 and (d["jtype"][1:] == 0).all()
    assert (np.diff(d["dof_off"]) > 0).all()
    depth = np.zeros(32, int)
    for i in range(32):
        depth[i] = 0 if d["parent"][i] < 0 else depth[d["parent"][i]] + 1
    assert depth.max() + 1 == 11                                       # max depth 11
    assert np.allclose(np.linalg.norm(d["axis"][1:], axis=1), 1.0)
    for J in d["J"]:
        assert np.allclose(J, J.T) and np.linalg.eigvalsh(J).min() > 0  # MecanoRandomTools.java:623-647
    assert (d["mass"] >= 0.1).all() and (d["mass"] <= 1.1).all()
    e36 = mb.RigidBody("elevator")
    mb.MultiBodySystemRandomTools.nextHumanoid(3, e36, 1)
    assert mb.MultiBodySystem.toMultiBodySystemBasics(e36).getNumberOfDoFs() == 36


def test_model_errors():
    e = mb.RigidBody("elevator")
    j = mb.RevoluteJoint("j", e, None, [0, 0, 1])
    with pytest.raises(mb.ScrewTheoryException):
        mb.MultiBodySystem.toMultiBodySystemBasics(e)                  # joint without successor
    mb.RigidBody("b", j, (1, 1, 1), 1.0, [0, 0, 0])
    with pytest.raises(mb.ScrewTheoryException):
        mb.RigidBody("b2", j, (1, 1, 1), 1.0, [0, 0, 0])               # joint already has a successor
    with pytest.raises(mb.ScrewTheoryException):
        mb.RevoluteJoint("bad", e, None, [0, 0, 0])                    # zero axis


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mecano_(?:b200|model)_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = _declared("mecano_b200.h") + _declared("mecano_b200_model.h")
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_capi.EXPORTS) == _declared("mecano_b200.h")
    assert sorted(multibody.MODEL_EXPORTS) == _declared("mecano_b200_model.h")
    assert lib.mecano_b200_version() == 200


def test_create_validates_without_a_gpu():
    """Topology / shape errors are reported by mecano_b200_create before any device work (status < 0, message set)."""
    _, s = small_system()
    good = s.tables().contents

    def attempt(mutate):
        d = _capi.TreeDesc.from_buffer_copy(bytes(good))
        keep = mutate(d)
        h = ctypes.c_void_p()
        rc = _capi.lib.mecano_b200_create(ctypes.byref(d), 0, ctypes.byref(h))
        msg = _capi.lib.mecano_b200_last_error(None).decode()
        assert not h.value or rc == 0
        if h.value:
            _capi.lib.mecano_b200_destroy(h)
        return rc, msg, keep

    rc, msg, _ = attempt(lambda d: setattr(d, "struct_size", 8))
    assert rc == -1 and "struct_size" in msg

    def bad_type(d):
        a = (ctypes.c_int32 * 4)(2, 0, 7, 1)
        d.joint_type = ctypes.cast(a, ctypes.POINTER(ctypes.c_int32))
        return a
    rc, msg, _ = attempt(bad_type)
    assert rc == -2 and "unsupported joint type" in msg                # e.g. SphericalJoint: no CPU fallback, error at create

    def loop(d):
        a = (ctypes.c_int32 * 4)(-1, 0, 3, 1)
        d.parent = ctypes.cast(a, ctypes.POINTER(ctypes.c_int32))
        return a
    rc, msg, _ = attempt(loop)
    assert rc == -2 and "parent" in msg

    rc, msg, _ = attempt(lambda d: setattr(d, "n_dofs", 8))
    assert rc == -3
    rc, msg, _ = attempt(lambda d: setattr(d, "n_bodies", 1000))
    assert rc == -5

    # a valid description fails only because there is no device here (or succeeds on a GPU box)
    rc, msg, _ = attempt(lambda d: None)
    assert rc in (0, -4)


def test_slices_cover_the_batch():
    for n in (0, 1, 7, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 4, 8):
            sl = [sharding.slice_for_rank(n, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.slice_for_rank(10, 2, 2)


def test_tree_specialised_source_generates_and_compiles_without_a_gpu():
    """mecano_b200_generate_source / mecano_b200_jit_check: the run-time compiled path (specialize.cpp + NVRTC) for a
    7-joint arm with a floating base; no device involved."""
    e = mb.RigidBody("elevator")
    base = mb.MultiBodySystemRandomTools.nextFloatingBase(5, e).getSuccessor()
    mb.MultiBodySystemRandomTools.nextOneDoFJointChain(6, base, 5, 0.4)
    s = mb.MultiBodySystem.toMultiBodySystemBasics(e)
    need = ctypes.c_int64()
    assert _capi.lib.mecano_b200_generate_source(s.tables(), 0, 256, 0, None, 0, ctypes.byref(need)) == 0
    buf = ctypes.create_string_buffer(need.value)
    assert _capi.lib.mecano_b200_generate_source(s.tables(), 0, 256, 0, buf, need.value, None) == 0
    src = buf.value.decode()
    assert "mb_spec_kernel" in src and src.count("rnea_op<") == 2 * s.getNumberOfJoints()  # one DESCEND + one ASCEND per joint
    for algo, block, tm in ((0, 512, 32), (1, 256, 0)):
        n = ctypes.c_int64()
        rc = _capi.lib.mecano_b200_jit_check(s.tables(), algo, block, tm, ctypes.byref(n))
        if rc != 0 and b"libnvrtc not found" in _capi.lib.mecano_b200_last_error(None):
            pytest.skip("NVRTC is not installed on this machine")
        assert rc == 0, _capi.lib.mecano_b200_last_error(None).decode()
        assert n.value > 10000
    # unknown algorithm
    assert _capi.lib.mecano_b200_jit_check(s.tables(), 7, 256, 0, None) == _capi_err("INVALID_ARGUMENT")


def _capi_err(name):
    return {"INVALID_ARGUMENT": -1, "UNSUPPORTED_TOPOLOGY": -2, "SHAPE": -3, "NO_DEVICE": -4, "TOO_LARGE": -5, "JIT": -6}[name]


def test_cpp_host_mirror_compiles_and_links(tmp_path):
    """The C++ face of the calculator API (mecano_b200/csrc/host/calculators.hpp) is header-only: compile a translation unit
    that names every member and link it against the C-ABI library (no GPU needed)."""
    import subprocess

    exe = str(tmp_path / "host_mirror_check")
    libdir = os.path.dirname(_capi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "cpp", "host_mirror_check.cpp"), "-o", exe,
                           "-L", libdir, "-lmecano_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe]).decode()
    assert "host mirror links: 1" in out


def test_jvm_exchange_file_round_trip(tmp_path, capsys):
    """scripts/java_exchange.py: the file handed to baseline/java/MecanoHarness.java reads back to the same tables and states, and
    the comparison passes when the results are the oracle's own (standing in for the JVM output, which cannot be produced here)."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import java_exchange as jx
    import oracle_lib as ol

    x, r = str(tmp_path / "exchange.bin"), str(tmp_path / "results.bin")
    jx.write("tree12", 9, x, seed=5)
    t, g, (q, qd, qdd, tau) = jx.read(x)
    assert t.nb == 13 and t.nv == 18 and q.shape == (t.nq, 9) and g == jx.GRAVITY
    ref = jx.build("tree12", 5).describe()
    assert np.array_equal(t.parent, ref["parent"]) and np.array_equal(t.J, ref["J"]) and np.array_equal(t.dof_off, ref["dof_off"])
    o = ol.Oracle(t, gravity=g)
    np.concatenate([o.rnea_batch(q, qd, qdd).ravel(), o.aba_batch(q, qd, tau).ravel(), o.crba_batch(q).ravel()]).astype("<f8").tofile(r)
    capsys.readouterr()
    jx.compare(x, r)
    assert '"pass": true' in capsys.readouterr().out


def test_new_entry_points_reject_a_null_handle():
    """Every entry point added for the SURVEY 8f rows returns MECANO_B200_ERR_INVALID_ARGUMENT for a NULL handle instead of
    crashing (no GPU needed)."""
    lib = _capi.lib
    n, ld = ctypes.c_int64(4), ctypes.c_int64(4)
    null = ctypes.c_void_p(None)
    bad = -1  # MECANO_B200_ERR_INVALID_ARGUMENT
    assert lib.mecano_b200_rnea_full(null, n, ld, null, null, null, null, null, null, null, 0, null) == bad
    assert lib.mecano_b200_rnea_full_host(null, n, ld, null, null, null, null, null, null, null, 0) == bad
    assert lib.mecano_b200_set_joint_source_modes(null, null) == bad
    assert lib.mecano_b200_aba_sources(null, n, ld, null, null, null, null, null, null, null, null) == bad
    assert lib.mecano_b200_aba_sources_host(null, n, ld, null, null, null, null, null, null, null) == bad
    assert lib.mecano_b200_crba_centroidal(null, n, ld, null, null, null, null, 0, null) == bad
    assert lib.mecano_b200_crba_centroidal_host(null, n, ld, null, null, null, null, 0) == bad
    assert lib.mecano_b200_centroidal_convective_term(null, n, ld, null, null, null, null, 0, null) == bad
    assert lib.mecano_b200_centroidal_convective_term_host(null, n, ld, null, null, null, null, 0) == bad
    assert lib.mecano_b200_center_of_mass(null, n, ld, null, null, null) == bad
    assert lib.mecano_b200_center_of_mass_host(null, n, ld, null, null) == bad
    assert lib.mecano_b200_coriolis(null, n, ld, null, null, null, null, null) == bad
    assert lib.mecano_b200_coriolis_host(null, n, ld, null, null, null, null) == bad
    assert lib.mecano_b200_set_precision(null, 1) == bad
    assert lib.mecano_b200_set_grid_limit(null, 0, 10) == bad
