"""ctypes binding to the kernel-source emulation harness (tests/emu).  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_EMU = os.path.join(_HERE, "emu")
_CSRC = os.path.join(os.path.dirname(_HERE), "mecano_b200", "csrc")
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


class TreeDescC(ctypes.Structure):
    """mirror of mecano_b200_tree_desc (include/mecano_b200.h)"""
    _fields_ = [
        ("struct_size", ctypes.c_int32), ("n_bodies", ctypes.c_int32), ("n_dofs", ctypes.c_int32), ("n_cfg", ctypes.c_int32),
        ("n_levels", ctypes.c_int32), ("level_start", _ip), ("parent", _ip), ("joint_type", _ip), ("axis", _dp),
        ("offset_rot", _dp), ("offset_pos", _dp), ("com_rot", _dp), ("com_pos", _dp), ("inertia", _dp), ("mass", _dp),
        ("dof_offset", _ip), ("cfg_offset", _ip), ("wrench_index", _ip),
    ]


def level_order(tree):
    """Re-list the bodies of a TreeDesc level by level (what the Java host hands to the C-ABI)."""
    nb = tree.nb
    depth = np.zeros(nb, dtype=np.int64)
    for i in range(nb):
        depth[i] = 0 if tree.parent[i] < 0 else depth[tree.parent[i]] + 1
    order = np.argsort(depth, kind="stable")
    new_of = np.empty(nb, dtype=np.int64)
    new_of[order] = np.arange(nb)
    parent = np.array([-1 if tree.parent[o] < 0 else new_of[tree.parent[o]] for o in order], dtype=np.int32)
    nlev = int(depth.max()) + 1
    level_start = np.zeros(nlev + 1, dtype=np.int32)
    for dpt in depth:
        level_start[dpt + 1] += 1
    level_start = np.cumsum(level_start).astype(np.int32)
    return order, parent, level_start


def tree_desc_c(tree, level_ordered=True):
    """Build the C struct (and keep the arrays alive) from a tests.treedesc.TreeDesc."""
    if level_ordered:
        order, parent, level_start = level_order(tree)
    else:
        order, parent, level_start = np.arange(tree.nb), tree.parent.copy(), np.zeros(1, np.int32)
    keep = {
        "level_start": np.ascontiguousarray(level_start, np.int32),
        "parent": np.ascontiguousarray(parent, np.int32),
        "joint_type": np.ascontiguousarray(tree.jtype[order], np.int32),
        "axis": np.ascontiguousarray(tree.axis[order]),
        "offset_rot": np.ascontiguousarray(tree.off_R[order]),
        "offset_pos": np.ascontiguousarray(tree.off_p[order]),
        "com_rot": np.ascontiguousarray(tree.com_R[order]),
        "com_pos": np.ascontiguousarray(tree.com_p[order]),
        "inertia": np.ascontiguousarray(tree.J[order]),
        "mass": np.ascontiguousarray(tree.mass[order]),
        "dof_offset": np.ascontiguousarray(tree.dof_off[order], np.int32),
        "cfg_offset": np.ascontiguousarray(tree.cfg_off[order], np.int32),
    }
    d = TreeDescC()
    d.struct_size = ctypes.sizeof(TreeDescC)
    d.n_bodies, d.n_dofs, d.n_cfg = tree.nb, tree.nv, tree.nq
    d.n_levels = len(level_start) - 1 if level_ordered else 0
    d.level_start = keep["level_start"].ctypes.data_as(_ip) if level_ordered else None
    for name in ("parent", "joint_type", "dof_offset", "cfg_offset"):
        setattr(d, name, keep[name].ctypes.data_as(_ip))
    for name in ("axis", "offset_rot", "offset_pos", "com_rot", "com_pos", "inertia", "mass"):
        setattr(d, name, keep[name].ctypes.data_as(_dp))
    return d, keep, order


def build():
    src = [os.path.join(_EMU, "emu.cpp"), os.path.join(_CSRC, "flatten.cpp")]
    out = os.path.join(_EMU, "libmecano_emu.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++"] + src + ["-o", out])


def lib():
    global _LIB
    if _LIB is None:
        out = os.path.join(_EMU, "libmecano_emu.so")
        deps = [os.path.join(_EMU, "emu.cpp")] + [os.path.join(_CSRC, f) for f in ("flatten.cpp", "flatten.h", "program.h", "spatial.cuh", "algorithms.cuh", "jointmath.cuh", "multidof.cuh", "multidof_aba.cuh", "rnea.cuh", "aba.cuh", "crba.cuh", "coriolis.cuh")]
        if not os.path.exists(out) or any(os.path.getmtime(out) < os.path.getmtime(f) for f in deps):
            build()
        _LIB = ctypes.CDLL(out)
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


class Emu:
    RNEA, ABA, CRBA = 0, 1, 2

    def __init__(self, tree, gravity=(0.0, 0.0, -9.81), fp32=False, level_ordered=True):
        self.tree = tree
        self.desc, self._keep, self.order = tree_desc_c(tree, level_ordered)
        self.g = np.ascontiguousarray(gravity, dtype=np.float64)
        self.fp32 = int(fp32)
        self.lib = lib()

    def _fext_rows(self, fext, n):
        """fext [nb_tree_order, 6, n] -> rows in the order of the C description."""
        if fext is None:
            return None
        return np.ascontiguousarray(fext[self.order].reshape(6 * self.tree.nb, n))

    def _run(self, algo, q, qd, x, fext, out, flags=0):
        n = q.shape[1]
        err = ctypes.create_string_buffer(256)
        rc = self.lib.emu_run(algo, self.fp32, ctypes.byref(self.desc), _d(self.g), ctypes.c_long(n), ctypes.c_long(n), _d(q), _d(qd), _d(x),
                              _d(fext), _d(out), ctypes.c_uint(flags), err, 256)
        if rc != 0:
            raise RuntimeError("emu_run rc=%d: %s" % (rc, err.value.decode()))
        return out

    def rnea(self, q, qd, qdd, fext=None, flags=0):
        n = q.shape[1]
        return self._run(0, q, qd, qdd, self._fext_rows(fext, n), np.full((self.tree.nv, n), np.nan), flags)

    def rnea_full(self, q, qd, qdd, fext=None, flags=0):
        """(tau [nv, n], body accelerations [nb, 6, n] in CoM frames, joint wrenches [nb, 6, n] in frameAfterJoint), bodies in
        the order of the tree description."""
        n = q.shape[1]
        nb = self.tree.nb
        tau, acc, wr = np.full((self.tree.nv, n), np.nan), np.full((6 * nb, n), np.nan), np.full((6 * nb, n), np.nan)
        err = ctypes.create_string_buffer(256)
        rc = self.lib.emu_rnea_full(ctypes.byref(self.desc), _d(self.g), ctypes.c_long(n), ctypes.c_long(n), _d(q), _d(qd), _d(qdd),
                                    _d(self._fext_rows(fext, n)), _d(tau), _d(acc), _d(wr), ctypes.c_uint(flags), err, 256)
        if rc != 0:
            raise RuntimeError("emu_rnea_full rc=%d: %s" % (rc, err.value.decode()))
        back = np.empty(nb, dtype=np.int64)
        back[self.order] = np.arange(nb)
        return tau, acc.reshape(nb, 6, n)[back], wr.reshape(nb, 6, n)[back]

    def aba(self, q, qd, tau, fext=None):
        n = q.shape[1]
        return self._run(1, q, qd, tau, self._fext_rows(fext, n), np.full((self.tree.nv, n), np.nan))

    def aba_sources(self, q, qd, tau, qdd_in, accel_source, fext=None):
        """Passes one to three with joints in ACCELERATION_SOURCE mode; accel_source [nb] in the order of the tree."""
        n = q.shape[1]
        qdd = np.full((self.tree.nv, n), np.nan)
        src = np.ascontiguousarray(np.asarray(accel_source)[self.order], dtype=np.int32)
        err = ctypes.create_string_buffer(256)
        rc = self.lib.emu_aba_sources(ctypes.byref(self.desc), _d(self.g), ctypes.c_long(n), ctypes.c_long(n), _d(q), _d(qd), _d(tau), _d(qdd_in),
                                      _d(self._fext_rows(fext, n)), src.ctypes.data_as(_ip), _d(qdd), err, 256)
        if rc != 0:
            raise RuntimeError("emu_aba_sources rc=%d: %s" % (rc, err.value.decode()))
        return qdd

    def crba(self, q):
        n = q.shape[1]
        nv = self.tree.nv
        return self._run(2, q, None, None, None, np.full((nv * nv, n), np.nan)).reshape(nv, nv, n)

    def packed_index(self):
        """(row, col) int32 arrays of the packed mass-matrix layout (MECANO_B200_CRBA_PACKED): packed row p = M[row[p], col[p]]."""
        n = self.lib.emu_packed_index(ctypes.byref(self.desc), None, None)
        assert n > 0
        row, col = np.zeros(n, np.int32), np.zeros(n, np.int32)
        assert self.lib.emu_packed_index(ctypes.byref(self.desc), row.ctypes.data_as(_ip), col.ctypes.data_as(_ip)) == n
        return row, col

    def crba_packed(self, q):
        """[packed rows, n]: every unique structurally non-zero entry once (the kernel's packed instantiation)."""
        n = q.shape[1]
        row, _ = self.packed_index()
        return self._run(2, q, None, None, None, np.full((len(row), n), np.nan), flags=4)

    def crba_centroidal(self, q):
        """(M [nv, nv, n], centroidal momentum matrix [6, nv, n] in the root frame, (mass * CoM, mass) [4, n])."""
        n = q.shape[1]
        nv = self.tree.nv
        M, cmm, com = np.full((nv * nv, n), np.nan), np.full((6 * nv, n), np.nan), np.full((4, n), np.nan)
        err = ctypes.create_string_buffer(256)
        rc = self.lib.emu_crba_centroidal(ctypes.byref(self.desc), ctypes.c_long(n), ctypes.c_long(n), _d(q), _d(M), _d(cmm), _d(com), err, 256)
        if rc != 0:
            raise RuntimeError("emu_crba_centroidal rc=%d: %s" % (rc, err.value.decode()))
        return M.reshape(nv, nv, n), cmm.reshape(6, nv, n), com

    def center_of_mass(self, q):
        """(mass * CoM, mass) [4, n] from the centre-of-mass-only launch of the by-product CRBA routine (no matrix buffers)."""
        n = q.shape[1]
        com = np.full((4, n), np.nan)
        err = ctypes.create_string_buffer(256)
        rc = self.lib.emu_center_of_mass(ctypes.byref(self.desc), ctypes.c_long(n), ctypes.c_long(n), _d(q), _d(com), err, 256)
        if rc != 0:
            raise RuntimeError("emu_center_of_mass rc=%d: %s" % (rc, err.value.decode()))
        return com

    def rnea_root_wrench(self, q, qd):
        """Wrench at the root, in the root frame, of inverse dynamics with zero joint accelerations and no gravity: [6, n]."""
        n = q.shape[1]
        tau, rw = np.full((self.tree.nv, n), np.nan), np.full((6, n), np.nan)
        err = ctypes.create_string_buffer(256)
        rc = self.lib.emu_rnea_root_wrench(ctypes.byref(self.desc), ctypes.c_long(n), ctypes.c_long(n), _d(q), _d(qd), _d(tau), _d(rw), err, 256)
        if rc != 0:
            raise RuntimeError("emu_rnea_root_wrench rc=%d: %s" % (rc, err.value.decode()))
        return rw

    def coriolis(self, q, qd):
        """(M [nv, nv, n], Coriolis and centrifugal matrix C [nv, nv, n])."""
        n = q.shape[1]
        nv = self.tree.nv
        M, C = np.full((nv * nv, n), np.nan), np.full((nv * nv, n), np.nan)
        err = ctypes.create_string_buffer(256)
        rc = self.lib.emu_coriolis(ctypes.byref(self.desc), ctypes.c_long(n), ctypes.c_long(n), _d(q), _d(qd), _d(M), _d(C), err, 256)
        if rc != 0:
            raise RuntimeError("emu_coriolis rc=%d: %s" % (rc, err.value.decode()))
        return M.reshape(nv, nv, n), C.reshape(nv, nv, n)

    def count_flops(self, algo, q, qd, x):
        """Algorithmic operation counts of one state (counting-scalar instantiation of the kernel routines)."""
        out = (ctypes.c_long * 5)()
        q1, qd1, x1 = (np.ascontiguousarray(a[:, :1]) for a in (q, qd, x))
        rc = self.lib.emu_count_flops(algo, ctypes.byref(self.desc), _d(q1), _d(qd1), _d(x1), out)
        assert rc == 0
        return dict(zip(("add", "mul", "div", "sincos", "flops"), list(out)))

    def program_info(self, algo):
        out = (ctypes.c_int * 8)()
        rc = self.lib.emu_program_info(ctypes.byref(self.desc), algo, out)
        assert rc == 0
        return dict(zip(("nb", "nops", "stack", "aux", "rec", "max_depth", "nv", "nq"), list(out)))
