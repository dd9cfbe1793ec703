"""Independent textbook formulation of RNEA / CRBA / ABA (Featherstone, "Rigid Body Dynamics
Algorithms", tables 5.1, 6.2, 7.1) with dense 6x6 Pluecker matrices in numpy.

This is a *second opinion* for the C oracle (oracle/mecano_oracle.c): it shares no code and no
algebraic shortcuts with it (body-frame recursion with local 6x6 transforms, versus Mecano's
frame-tree / through-the-root formulation).  Pure-Python loops: small cases only.

Spatial vectors are angular-first [w; v], like Mecano and like Featherstone.
"""
import numpy as np

REVOLUTE, PRISMATIC, SIXDOF, SPHERICAL, PLANAR = 0, 1, 2, 3, 4


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def rot_axis_angle(u, q):
    u = np.asarray(u, dtype=float)
    u = u / np.linalg.norm(u)
    K = skew(u)
    return np.eye(3) + np.sin(q) * K + (1.0 - np.cos(q)) * (K @ K)


def rot_quat(qx, qy, qz, qs):
    n = np.sqrt(qx * qx + qy * qy + qz * qz + qs * qs)
    qx, qy, qz, qs = qx / n, qy / n, qz / n, qs / n
    return np.array(
        [
            [1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qs * qz), 2 * (qx * qz + qs * qy)],
            [2 * (qx * qy + qs * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qs * qx)],
            [2 * (qx * qz - qs * qy), 2 * (qy * qz + qs * qx), 1 - 2 * (qx * qx + qy * qy)],
        ]
    )


def motion_xform(R, p):
    """6x6 motion transform from parent coordinates to child coordinates, the child frame being
    located at (R, p) in the parent (p_parent = R p_child + p)."""
    E = R.T
    X = np.zeros((6, 6))
    X[:3, :3] = E
    X[3:, 3:] = E
    X[3:, :3] = -E @ skew(p)
    return X


def crm(v):
    M = np.zeros((6, 6))
    M[:3, :3] = skew(v[:3])
    M[3:, 3:] = skew(v[:3])
    M[3:, :3] = skew(v[3:])
    return M


def crf(v):
    return -crm(v).T


class Model:
    def __init__(self, tree):
        """tree: tests.treedesc.TreeDesc (bodies in any topological order)."""
        self.t = tree
        nb = tree.nb
        self.S = []
        self.I = []
        self.Xcom = []
        for i in range(nb):
            jt = tree.jtype[i]
            if jt == REVOLUTE:
                S = np.zeros((6, 1))
                S[:3, 0] = tree.axis[i]
            elif jt == PRISMATIC:
                S = np.zeros((6, 1))
                S[3:, 0] = tree.axis[i]
            elif jt == SPHERICAL:  # angular velocity in the body frame (SphericalJointReadOnly.java:31-60)
                S = np.zeros((6, 3))
                S[0, 0] = S[1, 1] = S[2, 2] = 1.0
            elif jt == PLANAR:  # (w_y, v_x, v_z) in the body frame (MecanoTools.java:920-952)
                S = np.zeros((6, 3))
                S[1, 0] = S[3, 1] = S[5, 2] = 1.0
            else:
                S = np.eye(6)
            self.S.append(S)
            Rc, c = tree.com_R[i], tree.com_p[i]
            Ic = Rc @ tree.J[i] @ Rc.T
            m = tree.mass[i]
            C = skew(c)
            I6 = np.zeros((6, 6))
            I6[:3, :3] = Ic + m * C @ C.T
            I6[:3, 3:] = m * C
            I6[3:, :3] = m * C.T
            I6[3:, 3:] = m * np.eye(3)
            self.I.append(I6)
            self.Xcom.append(motion_xform(Rc, c))  # afterJoint -> CoM frame

    def joint_X(self, i, q):
        t = self.t
        qi = q[t.cfg_off[i]:]
        jt = t.jtype[i]
        if jt == REVOLUTE:
            RJ, tJ = rot_axis_angle(t.axis[i], qi[0]), np.zeros(3)
        elif jt == PRISMATIC:
            RJ, tJ = np.eye(3), qi[0] * np.asarray(t.axis[i])
        elif jt == SPHERICAL:
            RJ, tJ = rot_quat(*qi[:4]), np.zeros(3)
        elif jt == PLANAR:  # rotation about y by the pitch, translation in the x-z plane
            RJ, tJ = rot_axis_angle(np.array([0.0, 1.0, 0.0]), qi[0]), np.array([qi[1], 0.0, qi[2]])
        else:
            RJ, tJ = rot_quat(*qi[:4]), np.asarray(qi[4:7])
        R = t.off_R[i] @ RJ
        p = t.off_R[i] @ tJ + t.off_p[i]
        return motion_xform(R, p)

    def _nd(self, i):
        return self.S[i].shape[1]

    def _fext_after(self, i, fext):
        # wrench given in the CoM frame -> afterJoint coordinates
        return self.Xcom[i].T @ fext[i]

    def rnea(self, q, qd, qdd, gravity, fext=None):
        t = self.t
        nb = t.nb
        a0 = np.concatenate([np.zeros(3), -np.asarray(gravity, dtype=float)])
        X, v, a, f = [None] * nb, [None] * nb, [None] * nb, [None] * nb
        for i in range(nb):
            p = t.parent[i]
            d = t.dof_off[i]
            nd = self._nd(i)
            X[i] = self.joint_X(i, q)
            vJ = self.S[i] @ qd[d:d + nd]
            vp = v[p] if p >= 0 else np.zeros(6)
            ap = a[p] if p >= 0 else a0
            v[i] = X[i] @ vp + vJ
            a[i] = X[i] @ ap + self.S[i] @ qdd[d:d + nd] + crm(v[i]) @ vJ
            f[i] = self.I[i] @ a[i] + crf(v[i]) @ self.I[i] @ v[i]
            if fext is not None:
                f[i] = f[i] - self._fext_after(i, fext)
        tau = np.zeros(t.nv)
        for i in range(nb - 1, -1, -1):
            d = t.dof_off[i]
            nd = self._nd(i)
            tau[d:d + nd] = self.S[i].T @ f[i]
            p = t.parent[i]
            if p >= 0:
                f[p] = f[p] + X[i].T @ f[i]
        return tau

    def centroidal(self, q, qd, qdd):
        """Momentum of the whole system and its rate (no gravity) about the origin of the root frame, in root coordinates, plus
        the centre of mass and total mass: (h [6], hdot [6], com [3], mass), angular parts first.  Independent of the
        calculators: plain sums over the bodies."""
        t = self.t
        nb = t.nb
        X0, v, a = [None] * nb, [None] * nb, [None] * nb
        h, hdot, mc, mass = np.zeros(6), np.zeros(6), np.zeros(3), 0.0
        for i in range(nb):
            p = t.parent[i]
            d = t.dof_off[i]
            nd = self._nd(i)
            X = self.joint_X(i, q)
            vJ = self.S[i] @ qd[d:d + nd]
            v[i] = X @ (v[p] if p >= 0 else np.zeros(6)) + vJ
            a[i] = X @ (a[p] if p >= 0 else np.zeros(6)) + self.S[i] @ qdd[d:d + nd] + crm(v[i]) @ vJ
            X0[i] = X @ (X0[p] if p >= 0 else np.eye(6))  # root coordinates -> body coordinates (motion)
            h += X0[i].T @ (self.I[i] @ v[i])
            hdot += X0[i].T @ (self.I[i] @ a[i] + crf(v[i]) @ self.I[i] @ v[i])
            # position of the body's CoM in the root frame: E = X0[:3, :3] (root -> body), -E r~ = X0[3:, :3]
            E = X0[i][:3, :3]
            rx = -E.T @ X0[i][3:, :3]
            r = np.array([rx[2, 1], rx[0, 2], rx[1, 0]])
            mc += t.mass[i] * (r + E.T @ np.asarray(t.com_p[i]))
            mass += t.mass[i]
        return h, hdot, mc / mass, mass

    def crba(self, q):
        t = self.t
        nb = t.nb
        X = [self.joint_X(i, q) for i in range(nb)]
        Ic = [I.copy() for I in self.I]
        for i in range(nb - 1, -1, -1):
            p = t.parent[i]
            if p >= 0:
                Ic[p] = Ic[p] + X[i].T @ Ic[i] @ X[i]
        H = np.zeros((t.nv, t.nv))
        for i in range(nb):
            di, ni = t.dof_off[i], self._nd(i)
            F = Ic[i] @ self.S[i]
            H[di:di + ni, di:di + ni] = self.S[i].T @ F
            j = i
            while t.parent[j] >= 0:
                F = X[j].T @ F
                j = t.parent[j]
                dj, nj = t.dof_off[j], self._nd(j)
                H[di:di + ni, dj:dj + nj] = F.T @ self.S[j]
                H[dj:dj + nj, di:di + ni] = (F.T @ self.S[j]).T
        return H

    def aba(self, q, qd, tau, gravity, fext=None):
        t = self.t
        nb = t.nb
        a0 = np.concatenate([np.zeros(3), -np.asarray(gravity, dtype=float)])
        X, v, c, IA, pA = [None] * nb, [None] * nb, [None] * nb, [None] * nb, [None] * nb
        for i in range(nb):
            p = t.parent[i]
            d, nd = t.dof_off[i], self._nd(i)
            X[i] = self.joint_X(i, q)
            vJ = self.S[i] @ qd[d:d + nd]
            vp = v[p] if p >= 0 else np.zeros(6)
            v[i] = X[i] @ vp + vJ
            c[i] = crm(v[i]) @ vJ
            IA[i] = self.I[i].copy()
            pA[i] = crf(v[i]) @ self.I[i] @ v[i]
            if fext is not None:
                pA[i] = pA[i] - self._fext_after(i, fext)
        U, Dinv, u = [None] * nb, [None] * nb, [None] * nb
        for i in range(nb - 1, -1, -1):
            d, nd = t.dof_off[i], self._nd(i)
            U[i] = IA[i] @ self.S[i]
            Dinv[i] = np.linalg.inv(self.S[i].T @ U[i])
            u[i] = tau[d:d + nd] - self.S[i].T @ pA[i]
            p = t.parent[i]
            if p >= 0:
                Ia = IA[i] - U[i] @ Dinv[i] @ U[i].T
                pa = pA[i] + Ia @ c[i] + U[i] @ Dinv[i] @ u[i]
                IA[p] = IA[p] + X[i].T @ Ia @ X[i]
                pA[p] = pA[p] + X[i].T @ pa
        qdd = np.zeros(t.nv)
        a = [None] * nb
        for i in range(nb):
            p = t.parent[i]
            d, nd = t.dof_off[i], self._nd(i)
            ap = a[p] if p >= 0 else a0
            a1 = X[i] @ ap + c[i]
            qdd[d:d + nd] = Dinv[i] @ (u[i] - U[i].T @ a1)
            a[i] = a1 + self.S[i] @ qdd[d:d + nd]
        return qdd
