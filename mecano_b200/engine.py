"""Low-level engine: one handle = one flattened tree on one device.  Thin, allocation-free wrapper
over the C ABI for torch CUDA tensors (device entry points) and numpy / pinned arrays (host entry
points).  Buffers are [rows, n_states] float64 with unit stride along states (ld = stride of rows)."""
import ctypes

import numpy as np

from . import _capi
from ._capi import check, lib


def _dev_ptr_ld(t, rows, n, device=None):
    """(pointer, ld) of a torch CUDA tensor [rows, n] float64 with contiguous states (on `device`, if given)."""
    import torch

    if t is None:
        return None, None
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64):
        raise TypeError("expected a float64 CUDA tensor")
    if device is not None and t.device.index != device:
        raise ValueError("tensor lives on cuda:%s, the engine on cuda:%d" % (t.device.index, device))
    if t.dim() != 2 or t.shape[0] != rows or t.shape[1] != n:
        raise ValueError("expected shape [%d, %d], got %s" % (rows, n, tuple(t.shape)))
    if n > 1 and t.stride(1) != 1:
        raise ValueError("states must be contiguous (stride 1 along dim 1)")
    return t.data_ptr(), (t.stride(0) if rows > 1 else max(n, t.stride(0)))


def _host_ptr_ld(a, rows, n):
    if a is None:
        return None, None
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64):
        raise TypeError("expected a float64 numpy array")
    if a.ndim != 2 or a.shape[0] != rows or a.shape[1] != n:
        raise ValueError("expected shape [%d, %d], got %s" % (rows, n, a.shape))
    if n > 1 and a.strides[1] != 8:
        raise ValueError("states must be contiguous")
    return a.ctypes.data, (a.strides[0] // 8 if rows > 1 else max(n, a.strides[0] // 8))


def _same_ld(lds):
    lds = [l for l in lds if l is not None]
    if any(l != lds[0] for l in lds):
        raise ValueError("all buffers of one call must share the same leading dimension")
    return lds[0]


class Engine:
    def __init__(self, desc, device=0, keepalive=None):
        """desc: _capi.TreeDesc (level-ordered tables)."""
        self._h = ctypes.c_void_p()
        self._keep = keepalive
        rc = lib.mecano_b200_create(ctypes.byref(desc), int(device), ctypes.byref(self._h))
        if rc != 0:
            msg = lib.mecano_b200_last_error(None)
            raise _capi.MecanoB200Error(rc, msg.decode() if msg else "")
        self.device = int(device)
        self.nv = lib.mecano_b200_n_dofs(self._h)
        self.nq = lib.mecano_b200_n_cfg(self._h)
        self.nb = lib.mecano_b200_n_bodies(self._h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.mecano_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_gravity(self, gx, gy, gz):
        check(lib.mecano_b200_set_gravity(self._h, float(gx), float(gy), float(gz)), self._h)

    def set_variant(self, variant):
        """0 = auto (by batch size), 1 = one thread per state, 2 = one warp per state (trees of up to 32 bodies)."""
        check(lib.mecano_b200_set_variant(self._h, int(variant)), self._h)

    def set_precision(self, precision):
        """"fp64" (default) or "fp32": the optional single-precision variant of the plain RNEA / ABA / CRBA calls."""
        check(lib.mecano_b200_set_precision(self._h, {"fp64": 0, "fp32": 1}[precision]), self._h)

    def set_grid_limit(self, algo, max_blocks):
        """Cap the persistent grid of one algorithm (0 = whole device): leaves SMs to kernels running concurrently on other streams."""
        check(lib.mecano_b200_set_grid_limit(self._h, int(algo), int(max_blocks)), self._h)

    def specialize(self, algos=("rnea", "aba", "crba"), force=False):
        """Compile tree-specialised kernels (mecano_b200_specialize): seconds per algorithm, cached on disk.  Algorithms whose
        unrolled code would not fit the instruction caches keep the generic kernel unless force=True."""
        mask = 0x100 if force else 0
        for a in algos:
            mask |= 1 << {"rnea": 0, "aba": 1, "crba": 2}[a] if isinstance(a, str) else 1 << int(a)
        check(lib.mecano_b200_specialize(self._h, mask), self._h)

    def kernel_info(self, algo, n_states=0):
        info = _capi.KernelInfo()
        check(lib.mecano_b200_kernel_info_get(self._h, int(algo), int(n_states), ctypes.byref(info)), self._h)
        return info.as_dict()

    def _dp(self, t, rows, n):
        return _dev_ptr_ld(t, rows, n, self.device)

    def _stream(self):
        """torch's current stream ON THE ENGINE'S DEVICE (not on torch's current device)."""
        import torch

        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- device entry points (asynchronous on torch's current stream)
    def rnea(self, q, qd, qdd, tau, fext=None, flags=0, body_acc=None, joint_wrench=None):
        """body_acc / joint_wrench ([6 * nb, n], optional): the by-products of mecano_b200_rnea_full."""
        n = q.shape[1]
        pq, l0 = self._dp(q, self.nq, n)
        pqd, l1 = self._dp(qd, self.nv, n)
        pqdd, l2 = self._dp(qdd, self.nv, n)
        pt, l3 = self._dp(tau, self.nv, n)
        pf, l4 = self._dp(fext, 6 * self.nb, n)
        if body_acc is None and joint_wrench is None:
            check(lib.mecano_b200_rnea(self._h, n, _same_ld([l0, l1, l2, l3, l4]), pq, pqd, pqdd, pf, pt, flags, self._stream()), self._h)
            return tau
        pa, l5 = self._dp(body_acc, 6 * self.nb, n)
        pw, l6 = self._dp(joint_wrench, 6 * self.nb, n)
        check(lib.mecano_b200_rnea_full(self._h, n, _same_ld([l0, l1, l2, l3, l4, l5, l6]), pq, pqd, pqdd, pf, pt, pa, pw, flags, self._stream()), self._h)
        return tau

    def aba(self, q, qd, tau, qdd, fext=None, flags=0):
        n = q.shape[1]
        pq, l0 = self._dp(q, self.nq, n)
        pqd, l1 = self._dp(qd, self.nv, n)
        pt, l2 = self._dp(tau, self.nv, n)
        pqdd, l3 = self._dp(qdd, self.nv, n)
        pf, l4 = self._dp(fext, 6 * self.nb, n)
        check(lib.mecano_b200_aba(self._h, n, _same_ld([l0, l1, l2, l3, l4]), pq, pqd, pt, pf, pqdd, flags, self._stream()), self._h)
        return qdd

    def set_joint_source_modes(self, accel_source):
        """accel_source: [n_bodies] ints in the order of the tree description (non-zero = ACCELERATION_SOURCE), or None to reset."""
        if accel_source is None:
            check(lib.mecano_b200_set_joint_source_modes(self._h, None), self._h)
            return
        src = np.ascontiguousarray(accel_source, dtype=np.int32)
        if src.shape != (self.nb,):
            raise ValueError("expected %d source modes" % self.nb)
        check(lib.mecano_b200_set_joint_source_modes(self._h, src.ctypes.data), self._h)

    def aba_sources(self, q, qd, tau, qdd_in, qdd, tau_out=None, fext=None):
        n = q.shape[1]
        pq, l0 = self._dp(q, self.nq, n)
        pqd, l1 = self._dp(qd, self.nv, n)
        pt, l2 = self._dp(tau, self.nv, n)
        pi, l3 = self._dp(qdd_in, self.nv, n)
        pqdd, l4 = self._dp(qdd, self.nv, n)
        pto, l5 = self._dp(tau_out, self.nv, n)
        pf, l6 = self._dp(fext, 6 * self.nb, n)
        check(lib.mecano_b200_aba_sources(self._h, n, _same_ld([l0, l1, l2, l3, l4, l5, l6]), pq, pqd, pt, pi, pf, pqdd, pto, self._stream()), self._h)
        return qdd

    def aba_sources_host(self, q, qd, tau, qdd_in, qdd, tau_out=None, fext=None):
        n = q.shape[1]
        pq, l0 = _host_ptr_ld(q, self.nq, n)
        pqd, l1 = _host_ptr_ld(qd, self.nv, n)
        pt, l2 = _host_ptr_ld(tau, self.nv, n)
        pi, l3 = _host_ptr_ld(qdd_in, self.nv, n)
        pqdd, l4 = _host_ptr_ld(qdd, self.nv, n)
        pto, l5 = _host_ptr_ld(tau_out, self.nv, n)
        pf, l6 = _host_ptr_ld(fext, 6 * self.nb, n)
        check(lib.mecano_b200_aba_sources_host(self._h, n, _same_ld([l0, l1, l2, l3, l4, l5, l6]), pq, pqd, pt, pi, pf, pqdd, pto), self._h)
        return qdd

    def crba(self, q, M, layout=_capi.CRBA_ENTRY_MAJOR):
        """M: [nv*nv, n] (entry-major) or [n, nv*nv] (state-major) float64 CUDA tensor."""
        n = q.shape[1]
        pq, ld = self._dp(q, self.nq, n)
        if not (layout & _capi.CRBA_STATE_MAJOR):
            pm, lm = self._dp(M, self.nv * self.nv, n)
            ld = _same_ld([ld, lm])
        else:
            if tuple(M.shape) != (n, self.nv * self.nv) or not M.is_contiguous():
                raise ValueError("state-major mass matrix must be a contiguous [n, nv*nv] tensor")
            pm = M.data_ptr()
        check(lib.mecano_b200_crba(self._h, n, ld, pq, pm, layout, self._stream()), self._h)
        return M

    def coriolis(self, q, qd, M, C):
        """M, C [nv*nv, n] entry-major; torch CUDA tensors or numpy arrays (host path)."""
        n = q.shape[1]
        host = isinstance(q, np.ndarray)
        f = _host_ptr_ld if host else self._dp
        pq, l0 = f(q, self.nq, n)
        pqd, l1 = f(qd, self.nv, n)
        pm, l2 = f(M, self.nv * self.nv, n)
        pc, l3 = f(C, self.nv * self.nv, n)
        ld = _same_ld([l0, l1, l2, l3])
        if host:
            check(lib.mecano_b200_coriolis_host(self._h, n, ld, pq, pqd, pm, pc), self._h)
        else:
            check(lib.mecano_b200_coriolis(self._h, n, ld, pq, pqd, pm, pc, self._stream()), self._h)

    def crba_centroidal(self, q, M, cmm, com, frame=_capi.FRAME_WORLD):
        """M [nv*nv, n] entry-major, cmm [6*nv, n], com [4, n]; torch CUDA tensors or numpy arrays (host path)."""
        n = q.shape[1]
        host = isinstance(q, np.ndarray)
        f = _host_ptr_ld if host else self._dp
        pq, l0 = f(q, self.nq, n)
        pm, l1 = f(M, self.nv * self.nv, n)
        pa, l2 = f(cmm, 6 * self.nv, n)
        pc, l3 = f(com, 4, n)
        ld = _same_ld([l0, l1, l2, l3])
        if host:
            check(lib.mecano_b200_crba_centroidal_host(self._h, n, ld, pq, pm, pa, pc, int(frame)), self._h)
        else:
            check(lib.mecano_b200_crba_centroidal(self._h, n, ld, pq, pm, pa, pc, int(frame), self._stream()), self._h)

    def centroidal_convective_term(self, q, qd, com, out, frame=_capi.FRAME_WORLD):
        """out [6, n]; com [4, n] as written by crba_centroidal (None allowed for the world frame)."""
        n = q.shape[1]
        host = isinstance(q, np.ndarray)
        f = _host_ptr_ld if host else self._dp
        pq, l0 = f(q, self.nq, n)
        pqd, l1 = f(qd, self.nv, n)
        pc, l2 = f(com, 4, n)
        po, l3 = f(out, 6, n)
        ld = _same_ld([l0, l1, l2, l3])
        if host:
            check(lib.mecano_b200_centroidal_convective_term_host(self._h, n, ld, pq, pqd, pc, po, int(frame)), self._h)
        else:
            check(lib.mecano_b200_centroidal_convective_term(self._h, n, ld, pq, pqd, pc, po, int(frame), self._stream()), self._h)

    def integrate(self, dt, q, qd, qdd):
        """doubleIntegrateFromAcceleration on device matrices, in place."""
        n = q.shape[1]
        pq, l0 = self._dp(q, self.nq, n)
        pqd, l1 = self._dp(qd, self.nv, n)
        pqdd, l2 = self._dp(qdd, self.nv, n)
        check(lib.mecano_b200_integrate(self._h, n, _same_ld([l0, l1, l2]), float(dt), pq, pqd, pqdd, self._stream()), self._h)

    def integrate_host(self, dt, q, qd, qdd):
        n = q.shape[1]
        pq, l0 = _host_ptr_ld(q, self.nq, n)
        pqd, l1 = _host_ptr_ld(qd, self.nv, n)
        pqdd, l2 = _host_ptr_ld(qdd, self.nv, n)
        check(lib.mecano_b200_integrate_host(self._h, n, _same_ld([l0, l1, l2]), float(dt), pq, pqd, pqdd), self._h)

    # ---- host entry points (synchronous; inputs/outputs in host memory, pinned for full speed)
    def rnea_host(self, q, qd, qdd, tau, fext=None, flags=0, body_acc=None, joint_wrench=None):
        n = q.shape[1]
        pq, l0 = _host_ptr_ld(q, self.nq, n)
        pqd, l1 = _host_ptr_ld(qd, self.nv, n)
        pqdd, l2 = _host_ptr_ld(qdd, self.nv, n)
        pt, l3 = _host_ptr_ld(tau, self.nv, n)
        pf, l4 = _host_ptr_ld(fext, 6 * self.nb, n)
        if body_acc is None and joint_wrench is None:
            check(lib.mecano_b200_rnea_host(self._h, n, _same_ld([l0, l1, l2, l3, l4]), pq, pqd, pqdd, pf, pt, flags), self._h)
            return tau
        pa, l5 = _host_ptr_ld(body_acc, 6 * self.nb, n)
        pw, l6 = _host_ptr_ld(joint_wrench, 6 * self.nb, n)
        check(lib.mecano_b200_rnea_full_host(self._h, n, _same_ld([l0, l1, l2, l3, l4, l5, l6]), pq, pqd, pqdd, pf, pt, pa, pw, flags), self._h)
        return tau

    def aba_host(self, q, qd, tau, qdd, fext=None, flags=0):
        n = q.shape[1]
        pq, l0 = _host_ptr_ld(q, self.nq, n)
        pqd, l1 = _host_ptr_ld(qd, self.nv, n)
        pt, l2 = _host_ptr_ld(tau, self.nv, n)
        pqdd, l3 = _host_ptr_ld(qdd, self.nv, n)
        pf, l4 = _host_ptr_ld(fext, 6 * self.nb, n)
        check(lib.mecano_b200_aba_host(self._h, n, _same_ld([l0, l1, l2, l3, l4]), pq, pqd, pt, pf, pqdd, flags), self._h)
        return qdd

    def crba_host(self, q, M, layout=_capi.CRBA_ENTRY_MAJOR):
        n = q.shape[1]
        pq, ld = _host_ptr_ld(q, self.nq, n)
        if not (layout & _capi.CRBA_STATE_MAJOR):
            pm, lm = _host_ptr_ld(M, self.nv * self.nv, n)
            ld = _same_ld([ld, lm])
        else:
            if M.shape != (n, self.nv * self.nv) or not M.flags.c_contiguous:
                raise ValueError("state-major mass matrix must be a contiguous [n, nv*nv] array")
            pm = M.ctypes.data
        check(lib.mecano_b200_crba_host(self._h, n, ld, pq, pm, layout), self._h)
        return M


def measure_fp64_peak(device=0):
    v = ctypes.c_double()
    check(lib.mecano_b200_measure_fp64_peak(device, ctypes.byref(v)))
    return v.value


def measure_hbm_peak(device=0):
    v = ctypes.c_double()
    check(lib.mecano_b200_measure_hbm_peak(device, ctypes.byref(v)))
    return v.value
