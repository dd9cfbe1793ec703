"""ctypes binding to the CPU oracle (oracle/libmecano_oracle.so).  Test infrastructure only: the
product never imports this module (see oracle/mecano_oracle.h)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


class _MoTree(ctypes.Structure):
    _fields_ = [
        ("nb", ctypes.c_int), ("nv", ctypes.c_int), ("nq", ctypes.c_int),
        ("parent", _ip), ("jtype", _ip), ("axis", _dp), ("off_R", _dp), ("off_p", _dp),
        ("com_R", _dp), ("com_p", _dp), ("J", _dp), ("mass", _dp), ("dof_off", _ip), ("cfg_off", _ip),
    ]


def build():
    subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ORACLE_DIR, "libmecano_oracle.so")
        src = os.path.join(_ORACLE_DIR, "mecano_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.mo_max_threads.restype = ctypes.c_int
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class Oracle:
    def __init__(self, tree, gravity=(0.0, 0.0, -9.81)):
        self.t = tree.contiguous()
        t = self.t
        self.c = _MoTree(t.nb, t.nv, t.nq, _i(t.parent), _i(t.jtype), _d(t.axis), _d(t.off_R), _d(t.off_p),
                         _d(t.com_R), _d(t.com_p), _d(t.J), _d(t.mass), _i(t.dof_off), _i(t.cfg_off))
        self.g = np.ascontiguousarray(gravity, dtype=np.float64)
        self.lib = lib()

    @staticmethod
    def _f64(a):
        return None if a is None else np.ascontiguousarray(a, dtype=np.float64)

    # ---- single state
    def rnea(self, q, qd, qdd, fext=None, flags=0):
        q, qd, qdd, fext = map(self._f64, (q, qd, qdd, fext))
        tau = np.zeros(self.t.nv)
        self.lib.mo_rnea(ctypes.byref(self.c), _d(self.g), _d(q), _d(qd), _d(qdd), _d(fext), ctypes.c_int(flags), _d(tau))
        return tau

    def rnea_full(self, q, qd, qdd, fext=None, flags=0):
        """(tau [nv], body accelerations [nb, 6] in CoM frames, joint wrenches [nb, 6] in frameAfterJoint) of one state."""
        q, qd, qdd, fext = map(self._f64, (q, qd, qdd, fext))
        tau, acc, wr = np.zeros(self.t.nv), np.zeros((self.t.nb, 6)), np.zeros((self.t.nb, 6))
        self.lib.mo_rnea_full(ctypes.byref(self.c), _d(self.g), _d(q), _d(qd), _d(qdd), _d(fext), ctypes.c_int(flags), _d(tau), _d(acc), _d(wr))
        return tau, acc, wr

    def body_accelerations(self, q, qd, qdd, flags=0):
        q, qd, qdd = map(self._f64, (q, qd, qdd))
        acc = np.zeros((self.t.nb, 6))
        self.lib.mo_rnea_body_accelerations(ctypes.byref(self.c), _d(self.g), _d(q), _d(qd), _d(qdd), ctypes.c_int(flags), _d(acc))
        return acc

    def aba(self, q, qd, tau, fext=None):
        q, qd, tau, fext = map(self._f64, (q, qd, tau, fext))
        qdd = np.zeros(self.t.nv)
        self.lib.mo_aba(ctypes.byref(self.c), _d(self.g), _d(q), _d(qd), _d(tau), _d(fext), _d(qdd))
        return qdd

    def aba_sources(self, q, qd, tau, qdd_in, accel_source, fext=None):
        """ForwardDynamicsCalculator with per-joint source modes: accel_source [nb] (non-zero = ACCELERATION_SOURCE).
        Returns (getJointAccelerationMatrix(), getJointTauMatrix()) of one state."""
        q, qd, tau, qdd_in, fext = map(self._f64, (q, qd, tau, qdd_in, fext))
        src = np.ascontiguousarray(accel_source, dtype=np.int32)
        qdd, tau_out = np.zeros(self.t.nv), np.zeros(self.t.nv)
        self.lib.mo_aba_sources(ctypes.byref(self.c), _d(self.g), _d(q), _d(qd), _d(tau), _d(qdd_in), _d(fext), _i(src), _d(qdd), _d(tau_out))
        return qdd, tau_out

    def crba(self, q):
        q = self._f64(q)
        M = np.zeros((self.t.nv, self.t.nv))
        self.lib.mo_crba(ctypes.byref(self.c), _d(q), _d(M))
        return M

    def crba_centroidal(self, q, frame=0):
        """(M [nv, nv], centroidal momentum matrix [6, nv], CoM [3], total mass); frame 0 = root frame, 1 = CoM frame."""
        q = self._f64(q)
        nv = self.t.nv
        M, cmm, com4 = np.zeros((nv, nv)), np.zeros((6, nv)), np.zeros(4)
        self.lib.mo_crba_centroidal(ctypes.byref(self.c), _d(q), ctypes.c_int(frame), _d(M), _d(cmm), _d(com4))
        return M, cmm, com4[:3].copy(), float(com4[3])

    def centroidal_convective_term(self, q, qd, frame=0):
        q, qd = map(self._f64, (q, qd))
        out = np.zeros(6)
        self.lib.mo_centroidal_convective_term(ctypes.byref(self.c), _d(q), _d(qd), ctypes.c_int(frame), _d(out))
        return out

    def coriolis(self, q, qd):
        """(M [nv, nv], Coriolis and centrifugal matrix C [nv, nv]) of one state."""
        q, qd = map(self._f64, (q, qd))
        nv = self.t.nv
        M, C = np.zeros((nv, nv)), np.zeros((nv, nv))
        self.lib.mo_coriolis(ctypes.byref(self.c), _d(q), _d(qd), _d(M), _d(C))
        return M, C

    def integrate(self, dt, q, qd, qdd):
        """doubleIntegrateFromAcceleration on one state; returns the updated (q, qd, qdd) copies."""
        q, qd, qdd = (np.array(x, dtype=np.float64, copy=True) for x in (q, qd, qdd))
        self.lib.mo_integrate(ctypes.byref(self.c), ctypes.c_double(dt), _d(q), _d(qd), _d(qdd))
        return q, qd, qdd

    def integrate_batch(self, dt, q, qd, qdd):
        q, qd, qdd = (np.array(x, dtype=np.float64, copy=True, order="C") for x in (q, qd, qdd))
        n = q.shape[1]
        self.lib.mo_integrate_batch(ctypes.byref(self.c), ctypes.c_double(dt), ctypes.c_long(n), ctypes.c_long(n), _d(q), _d(qd), _d(qdd))
        return q, qd, qdd

    # ---- batched, [k, s] buffers
    def rnea_batch(self, q, qd, qdd, fext=None, flags=0, nthreads=0):
        q, qd, qdd, fext = map(self._f64, (q, qd, qdd, fext))
        n = q.shape[1]
        tau = np.zeros((self.t.nv, n))
        self.lib.mo_rnea_batch(ctypes.byref(self.c), _d(self.g), ctypes.c_long(n), ctypes.c_long(n), _d(q), _d(qd), _d(qdd),
                               _d(fext), ctypes.c_int(flags), _d(tau), ctypes.c_int(nthreads))
        return tau

    def aba_batch(self, q, qd, tau, fext=None, nthreads=0):
        q, qd, tau, fext = map(self._f64, (q, qd, tau, fext))
        n = q.shape[1]
        qdd = np.zeros((self.t.nv, n))
        self.lib.mo_aba_batch(ctypes.byref(self.c), _d(self.g), ctypes.c_long(n), ctypes.c_long(n), _d(q), _d(qd), _d(tau),
                              _d(fext), _d(qdd), ctypes.c_int(nthreads))
        return qdd

    def crba_batch(self, q, nthreads=0):
        q = self._f64(q)
        n = q.shape[1]
        M = np.zeros((self.t.nv * self.t.nv, n))
        self.lib.mo_crba_batch(ctypes.byref(self.c), ctypes.c_long(n), ctypes.c_long(n), _d(q), _d(M), ctypes.c_int(nthreads))
        return M.reshape(self.t.nv, self.t.nv, n)
