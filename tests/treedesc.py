"""Neutral description of a multi-body tree (numpy arrays, Mecano depth-first order) shared by the
oracle binding, the numpy Featherstone second opinion and the product tests, plus small numpy
generators that mimic Mecano's random tools (MultiBodySystemRandomTools.java:483-496, 908-923,
1211-1231, 1365-1371; MecanoRandomTools.java:623-647).  Test-side only.
"""
from dataclasses import dataclass

import numpy as np

REVOLUTE, PRISMATIC, SIXDOF, SPHERICAL, PLANAR = 0, 1, 2, 3, 4
NDOF = {REVOLUTE: 1, PRISMATIC: 1, SIXDOF: 6, SPHERICAL: 3, PLANAR: 3}
NCFG = {REVOLUTE: 1, PRISMATIC: 1, SIXDOF: 7, SPHERICAL: 4, PLANAR: 3}


@dataclass
class TreeDesc:
    nb: int
    nv: int
    nq: int
    parent: np.ndarray  # int32 [nb]
    jtype: np.ndarray  # int32 [nb]
    axis: np.ndarray  # [nb,3]
    off_R: np.ndarray  # [nb,3,3]
    off_p: np.ndarray  # [nb,3]
    com_R: np.ndarray  # [nb,3,3]
    com_p: np.ndarray  # [nb,3]
    J: np.ndarray  # [nb,3,3]
    mass: np.ndarray  # [nb]
    dof_off: np.ndarray  # int32 [nb]
    cfg_off: np.ndarray  # int32 [nb]

    def contiguous(self):
        for name in ("parent", "jtype", "dof_off", "cfg_off"):
            setattr(self, name, np.ascontiguousarray(getattr(self, name), dtype=np.int32))
        for name in ("axis", "off_R", "off_p", "com_R", "com_p", "J", "mass"):
            setattr(self, name, np.ascontiguousarray(getattr(self, name), dtype=np.float64))
        return self


def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    x, y, z, s = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - s * z), 2 * (x * z + s * y)],
            [2 * (x * y + s * z), 1 - 2 * (x * x + z * z), 2 * (y * z - s * x)],
            [2 * (x * z - s * y), 2 * (y * z + s * x), 1 - 2 * (x * x + y * y)],
        ]
    )


def random_spd_inertia(rng):
    L = np.zeros((3, 3))
    L[0, 0], L[1, 1], L[2, 2] = rng.uniform(1e-4, 2.0, size=3)
    L[1, 0], L[2, 0], L[2, 1] = rng.uniform(-0.5, 0.5, size=3)
    return L @ L.T


def dfs_order(parent):
    """Mecano joint order: depth-first pre-order, children in insertion (index) order."""
    nb = len(parent)
    children = [[] for _ in range(nb + 1)]
    for i, p in enumerate(parent):
        children[p + 1].append(i)
    order = []
    stack = list(reversed(children[0]))
    while stack:
        i = stack.pop()
        order.append(i)
        stack.extend(reversed(children[i + 1]))
    return order


def make_tree(rng, parent, jtype, com_rotation=False, axis_aligned=False):
    """Random physical parameters for the given topology; bodies are re-listed in Mecano DFS order."""
    parent = list(parent)
    order = dfs_order(parent)
    new_index = {old: new for new, old in enumerate(order)}
    nb = len(parent)
    t = TreeDesc(
        nb=nb, nv=0, nq=0,
        parent=np.zeros(nb, np.int32), jtype=np.zeros(nb, np.int32), axis=np.zeros((nb, 3)),
        off_R=np.zeros((nb, 3, 3)), off_p=np.zeros((nb, 3)), com_R=np.zeros((nb, 3, 3)), com_p=np.zeros((nb, 3)),
        J=np.zeros((nb, 3, 3)), mass=np.zeros(nb), dof_off=np.zeros(nb, np.int32), cfg_off=np.zeros(nb, np.int32),
    )
    nv = nq = 0
    for new, old in enumerate(order):
        p = parent[old]
        t.parent[new] = -1 if p < 0 else new_index[p]
        t.jtype[new] = jtype[old]
        if axis_aligned:
            a = np.zeros(3)
            a[rng.integers(3)] = 1.0
        else:
            a = rng.normal(size=3)
            a /= np.linalg.norm(a)
        t.axis[new] = a
        if p < 0:  # joints attached to the root body get a null offset (MultiBodySystemRandomTools.java:1229)
            t.off_R[new] = np.eye(3)
            t.off_p[new] = 0.0
        else:
            t.off_R[new] = random_rotation(rng)
            t.off_p[new] = rng.uniform(-1, 1, size=3)
        t.com_R[new] = random_rotation(rng) if com_rotation else np.eye(3)
        t.com_p[new] = rng.uniform(-1, 1, size=3)
        t.J[new] = random_spd_inertia(rng)
        t.mass[new] = 0.1 + rng.uniform()
        t.dof_off[new] = nv
        t.cfg_off[new] = nq
        nv += NDOF[jtype[old]]
        nq += NCFG[jtype[old]]
    t.nv, t.nq = nv, nq
    return t.contiguous()


def chain(rng, n, floating=False, prismatic_fraction=0.0, **kw):
    parent, jtype = [], []
    if floating:
        parent.append(-1)
        jtype.append(SIXDOF)
    for _ in range(n):
        parent.append(len(parent) - 1)
        jtype.append(PRISMATIC if rng.uniform() < prismatic_fraction else REVOLUTE)
    return make_tree(rng, parent, jtype, **kw)


def random_tree(rng, n, floating=False, prismatic_fraction=0.0, **kw):
    """nextRevoluteJointTree-style: each new joint hangs off a uniformly chosen existing body."""
    parent, jtype = [], []
    if floating:
        parent.append(-1)
        jtype.append(SIXDOF)
    first = len(parent)
    pred = len(parent) - 1
    for _ in range(n):
        parent.append(pred)
        jtype.append(PRISMATIC if rng.uniform() < prismatic_fraction else REVOLUTE)
        pred = int(rng.integers(first, len(parent)))
    return make_tree(rng, parent, jtype, **kw)


def mixed_chain(rng, types, floating=False, **kw):
    """A serial chain with the given joint types (ForwardDynamicsCalculatorTest.testJointChain: all joint types)."""
    parent, jtype = [], []
    if floating:
        parent.append(-1)
        jtype.append(SIXDOF)
    for jt in types:
        parent.append(len(parent) - 1)
        jtype.append(jt)
    return make_tree(rng, parent, jtype, **kw)


def mixed_tree(rng, n, floating=False, weights=(0.4, 0.2, 0.0, 0.2, 0.2), **kw):
    """Random tree whose joints are drawn from (revolute, prismatic, sixdof, spherical, planar) with the given weights."""
    parent, jtype = [], []
    if floating:
        parent.append(-1)
        jtype.append(SIXDOF)
    first = len(parent)
    pred = len(parent) - 1
    w = np.asarray(weights, dtype=float) / np.sum(weights)
    for _ in range(n):
        parent.append(pred)
        jtype.append(int(rng.choice(5, p=w)))
        pred = int(rng.integers(first, len(parent)))
    return make_tree(rng, parent, jtype, **kw)


def humanoid(rng, neck=2, **kw):
    """SixDoF pelvis + 2 legs x 6 + spine 3 + 2 arms x 7 + neck (2 -> H37, 1 -> H36); SURVEY.md 8(d)."""
    parent, jtype = [-1], [SIXDOF]

    def limb(root, n):
        prev = root
        for _ in range(n):
            parent.append(prev)
            jtype.append(REVOLUTE)
            prev = len(parent) - 1
        return prev

    limb(0, 6)
    limb(0, 6)
    chest = limb(0, 3)
    limb(chest, 7)
    limb(chest, 7)
    limb(chest, neck)
    return make_tree(rng, parent, jtype, **kw)


def random_states(rng, t, n):
    """DoF-major / state-minor buffers [k, s]; distributions per SURVEY.md 8(d)."""
    q = rng.uniform(-np.pi, np.pi, size=(t.nq, n))
    for i in range(t.nb):
        if t.jtype[i] in (SIXDOF, SPHERICAL):
            c = t.cfg_off[i]
            quat = rng.normal(size=(4, n))
            quat /= np.linalg.norm(quat, axis=0)
            q[c:c + 4] = quat
            if t.jtype[i] == SIXDOF:
                q[c + 4:c + 7] = rng.uniform(-1, 1, size=(3, n))
        elif t.jtype[i] == PLANAR:  # (pitch, x, z)
            c = t.cfg_off[i]
            q[c + 1:c + 3] = rng.uniform(-1, 1, size=(2, n))
    qd = rng.uniform(-1, 1, size=(t.nv, n))
    qdd = rng.uniform(-1, 1, size=(t.nv, n))
    tau = rng.uniform(-1, 1, size=(t.nv, n))
    return tuple(np.ascontiguousarray(x) for x in (q, qd, qdd, tau))
