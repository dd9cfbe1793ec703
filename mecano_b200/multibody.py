"""Host-side mirror of the Mecano model API needed by the hot path (Python face of the C++ mirror in
csrc/host/multibody.hpp, reached through include/mecano_b200_model.h).

    RigidBody / RevoluteJoint / PrismaticJoint / SixDoFJoint     M/multiBodySystem/*.java
    MultiBodySystemBasics.toMultiBodySystemBasics(rootBody)       M/multiBodySystem/interfaces/MultiBodySystemBasics.java:76-142
    JointMatrixIndexProvider                                      M/multiBodySystem/interfaces/JointMatrixIndexProvider.java:71-123
    MultiBodySystemRandomTools                                    M/tools/MultiBodySystemRandomTools.java

Method names follow the reference (camelCase) so that tests read like the reference's own tests.
Joints carry no state here: q, qd, qdd, tau are [rows, N] matrices handed to the calculators.
"""
import ctypes

import numpy as np

from . import _capi
from ._capi import lib

_vp = ctypes.c_void_p
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)

lib.mecano_model_create.argtypes = [ctypes.c_char_p]
lib.mecano_model_create.restype = _vp
lib.mecano_model_destroy.argtypes = [_vp]
lib.mecano_model_destroy.restype = None
lib.mecano_model_last_error.argtypes = [_vp]
lib.mecano_model_last_error.restype = ctypes.c_char_p
lib.mecano_model_add_revolute_joint.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int, _dp, _dp]
lib.mecano_model_add_prismatic_joint.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int, _dp, _dp]
lib.mecano_model_add_sixdof_joint.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int, _dp]
lib.mecano_model_add_fixed_joint.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int, _dp]
lib.mecano_model_add_spherical_joint.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int, _dp]
lib.mecano_model_add_planar_joint.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int, _dp]
lib.mecano_model_next_joint_chain.argtypes = [_vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
lib.mecano_model_next_joint_tree.argtypes = [_vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
lib.mecano_model_set_joint_configuration.argtypes = [_vp, ctypes.c_int, _dp, ctypes.c_int]
lib.mecano_model_finalize_ignoring.argtypes = [_vp, _ip, ctypes.c_int]
lib.mecano_model_add_rigid_body.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int, _dp, ctypes.c_double, _dp]
lib.mecano_model_next_one_dof_joint_chain.argtypes = [_vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_double]
lib.mecano_model_next_one_dof_joint_tree.argtypes = [_vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_double]
lib.mecano_model_next_floating_base.argtypes = [_vp, ctypes.c_uint64, ctypes.c_int]
lib.mecano_model_next_humanoid.argtypes = [_vp, ctypes.c_uint64, ctypes.c_int]
lib.mecano_model_finalize.argtypes = [_vp]
for _f in ("mecano_model_n_joints", "mecano_model_n_dofs", "mecano_model_n_cfg"):
    getattr(lib, _f).argtypes = [_vp]
lib.mecano_model_joint_order.argtypes = [_vp, _ip, _ip, _ip]
lib.mecano_model_joint_info.argtypes = [_vp, ctypes.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp]
lib.mecano_model_joint_name.argtypes = [_vp, ctypes.c_int]
lib.mecano_model_joint_name.restype = ctypes.c_char_p
lib.mecano_model_body_name.argtypes = [_vp, ctypes.c_int]
lib.mecano_model_body_name.restype = ctypes.c_char_p
lib.mecano_model_tables.argtypes = [_vp]
lib.mecano_model_tables.restype = ctypes.POINTER(_capi.TreeDesc)
lib.mecano_model_table_row.argtypes = [_vp, ctypes.c_int]
lib.mecano_model_expanded_tables.argtypes = [_vp]
lib.mecano_model_expanded_tables.restype = ctypes.POINTER(_capi.TreeDesc)
lib.mecano_model_expanded_info.argtypes = [_vp, _ip, _ip, _ip, _ip]
lib.mecano_model_expanded_fill.argtypes = [_vp, _dp, _ip, _ip]

MODEL_EXPORTS = [
    "mecano_model_create", "mecano_model_destroy", "mecano_model_last_error", "mecano_model_add_revolute_joint",
    "mecano_model_add_prismatic_joint", "mecano_model_add_sixdof_joint", "mecano_model_add_fixed_joint", "mecano_model_add_rigid_body",
    "mecano_model_add_spherical_joint", "mecano_model_add_planar_joint", "mecano_model_next_joint_chain", "mecano_model_next_joint_tree",
    "mecano_model_set_joint_configuration", "mecano_model_finalize_ignoring",
    "mecano_model_next_one_dof_joint_chain", "mecano_model_next_one_dof_joint_tree", "mecano_model_next_floating_base",
    "mecano_model_next_humanoid", "mecano_model_finalize", "mecano_model_n_joints", "mecano_model_n_dofs", "mecano_model_n_cfg",
    "mecano_model_joint_order", "mecano_model_joint_info", "mecano_model_joint_name", "mecano_model_body_name",
    "mecano_model_tables", "mecano_model_table_row",
    "mecano_model_expanded_tables", "mecano_model_expanded_info", "mecano_model_expanded_fill",
]


class ScrewTheoryException(RuntimeError):
    """M/exceptions/ScrewTheoryException.java"""


class RigidBodyTransform:
    """Euclid RigidBodyTransform: rotation (3x3) + translation (3)."""

    def __init__(self, rotation=None, translation=None):
        self.rotation = np.eye(3) if rotation is None else np.asarray(rotation, dtype=np.float64).reshape(3, 3)
        self.translation = np.zeros(3) if translation is None else np.asarray(translation, dtype=np.float64).reshape(3)

    def _flat12(self):
        return np.ascontiguousarray(np.concatenate([self.rotation.reshape(9), self.translation]), dtype=np.float64)


def _as_transform(t):
    if t is None:
        return None
    if isinstance(t, RigidBodyTransform):
        return t
    t = np.asarray(t, dtype=np.float64)
    if t.shape == (3,):  # translation-only constructors (RevoluteJoint.java:56-59)
        return RigidBodyTransform(translation=t)
    if t.shape == (4, 4):
        return RigidBodyTransform(t[:3, :3], t[:3, 3])
    raise ValueError("transform must be a RigidBodyTransform, a 3-vector offset or a 4x4 matrix")


class _Model:
    """Owns the C++ model (mecano_model)."""

    def __init__(self, root_name):
        self.h = _vp(lib.mecano_model_create(root_name.encode()))
        self.joints = []  # id -> Joint
        self.bodies = []  # id -> RigidBody
        self.finalized = False

    def __del__(self):
        try:
            if self.h:
                lib.mecano_model_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def check(self, rc):
        if rc < 0:
            raise ScrewTheoryException(lib.mecano_model_last_error(self.h).decode())
        return rc

    def adopt_generated(self):
        """Create Python proxies for joints/bodies appended by a C++ generator."""
        n = lib.mecano_model_n_joints(self.h)
        for j in range(len(self.joints), n):
            jt, pred = ctypes.c_int32(), ctypes.c_int32()
            lib.mecano_model_joint_info(self.h, j, ctypes.byref(jt), ctypes.byref(pred), None, None, None, None, None)
            cls = {0: RevoluteJoint, 1: PrismaticJoint, 2: SixDoFJoint, SPHERICAL: SphericalJoint, PLANAR: PlanarJoint, FIXED: FixedJoint}[jt.value]
            joint = cls.__new__(cls)
            joint._adopt(self, j, self.bodies[pred.value])
            body = RigidBody.__new__(RigidBody)
            body._adopt(self, j + 1, joint)


class RigidBody:
    """RigidBody(name) creates the root body; RigidBody(name, parentJoint, momentOfInertia, mass, centerOfMassOffset |
    inertiaPose) creates the successor of a joint (RigidBody.java:79-182).  momentOfInertia may be a 3x3 matrix or
    (Ixx, Iyy, Izz)."""

    def __init__(self, name, parentJoint=None, momentOfInertia=None, mass=None, inertiaPose=None):
        self._name = name
        self._children = []
        if parentJoint is None:
            self._model = _Model(name)
            self._id = 0
            self._parent = None
            self._model.bodies.append(self)
            return
        if momentOfInertia is None or mass is None:
            raise ValueError("momentOfInertia and mass are required for a non-root body")
        inertia = np.asarray(momentOfInertia, dtype=np.float64)
        if inertia.shape == (3,):
            inertia = np.diag(inertia)
        inertia = np.ascontiguousarray(inertia.reshape(3, 3))
        pose = _as_transform(inertiaPose) or RigidBodyTransform()
        p12 = pose._flat12()
        m = parentJoint._model
        bid = m.check(lib.mecano_model_add_rigid_body(m.h, name.encode(), parentJoint._id, inertia.ctypes.data_as(_dp), float(mass), p12.ctypes.data_as(_dp)))
        self._model, self._id, self._parent = m, bid, parentJoint
        parentJoint._successor = self
        while len(m.bodies) <= bid:
            m.bodies.append(None)
        m.bodies[bid] = self

    def _adopt(self, model, bid, parent_joint):
        self._model, self._id, self._parent = model, bid, parent_joint
        self._name = lib.mecano_model_body_name(model.h, bid).decode()
        self._children = []
        parent_joint._successor = self
        while len(model.bodies) <= bid:
            model.bodies.append(None)
        model.bodies[bid] = self

    def getName(self):
        return self._name

    def isRootBody(self):
        return self._parent is None

    def getParentJoint(self):
        return self._parent

    def getChildrenJoints(self):
        return list(self._children)


SPHERICAL, PLANAR = 3, 4  # MECANO_B200_SPHERICAL / MECANO_B200_PLANAR (include/mecano_b200.h)
FIXED = 5  # MECANO_MODEL_FIXED (include/mecano_b200_model.h)


class Joint:
    _type = None

    def _create(self, name, predecessor, transformToParent, jointAxis):
        if predecessor is None:
            raise ValueError("predecessor can not be null")
        m = predecessor._model
        t = _as_transform(transformToParent)
        t12 = None if t is None else t._flat12()
        tp = None if t12 is None else t12.ctypes.data_as(_dp)
        if self._type == _capi.SIXDOF:
            jid = lib.mecano_model_add_sixdof_joint(m.h, name.encode(), predecessor._id, tp)
        elif self._type == SPHERICAL:
            jid = lib.mecano_model_add_spherical_joint(m.h, name.encode(), predecessor._id, tp)
        elif self._type == PLANAR:
            jid = lib.mecano_model_add_planar_joint(m.h, name.encode(), predecessor._id, tp)
        elif self._type == FIXED:
            jid = lib.mecano_model_add_fixed_joint(m.h, name.encode(), predecessor._id, tp)
        else:
            ax = np.ascontiguousarray(jointAxis, dtype=np.float64).reshape(3)
            fn = lib.mecano_model_add_revolute_joint if self._type == _capi.REVOLUTE else lib.mecano_model_add_prismatic_joint
            jid = fn(m.h, name.encode(), predecessor._id, tp, ax.ctypes.data_as(_dp))
        m.check(jid)
        self._model, self._id, self._name, self._predecessor, self._successor = m, jid, name, predecessor, None
        predecessor._children.append(self)
        m.joints.append(self)

    def _adopt(self, model, jid, predecessor):
        self._model, self._id, self._predecessor, self._successor = model, jid, predecessor, None
        self._name = lib.mecano_model_joint_name(model.h, jid).decode()
        predecessor._children.append(self)
        model.joints.append(self)

    def getName(self):
        return self._name

    def getPredecessor(self):
        return self._predecessor

    def getSuccessor(self):
        return self._successor

    def getDegreesOfFreedom(self):
        return {_capi.SIXDOF: 6, SPHERICAL: 3, PLANAR: 3, FIXED: 0}.get(self._type, 1)

    def getConfigurationMatrixSize(self):
        return {_capi.SIXDOF: 7, SPHERICAL: 4, PLANAR: 3, FIXED: 0}.get(self._type, 1)

    def setJointConfiguration(self, q):
        """The configuration this joint keeps if it ends up in jointsToIgnore: Mecano lumps an ignored subtree into its parent
        body at the configuration the joints have when the calculator is built (InverseDynamicsCalculator.java:236,
        :832-860).  One-DoF: [q] (also setQ); SixDoF: [qx qy qz qs x y z]; Spherical: [qx qy qz qs]; Planar: [pitch x z].  Must be
        called before toMultiBodySystemBasics."""
        q = np.ascontiguousarray(np.atleast_1d(q), dtype=np.float64)
        self._model.check(lib.mecano_model_set_joint_configuration(self._model.h, self._id, q.ctypes.data_as(_dp), int(q.size)))

    def setQ(self, q):
        self.setJointConfiguration([float(q)])

    def _info(self):
        jt, pred = ctypes.c_int32(), ctypes.c_int32()
        axis, t12, I9, p12 = np.zeros(3), np.zeros(12), np.zeros(9), np.zeros(12)
        mass = ctypes.c_double()
        lib.mecano_model_joint_info(self._model.h, self._id, ctypes.byref(jt), ctypes.byref(pred), axis.ctypes.data_as(_dp), t12.ctypes.data_as(_dp),
                                    I9.ctypes.data_as(_dp), ctypes.byref(mass), p12.ctypes.data_as(_dp))
        return jt.value, pred.value, axis, t12, I9, mass.value, p12

    def getJointAxis(self):
        return self._info()[2]


class RevoluteJoint(Joint):
    """RevoluteJoint(name, predecessor, transformToParent | jointOffset | None, jointAxis)  (RevoluteJoint.java:42-74)"""
    _type = _capi.REVOLUTE

    def __init__(self, name, predecessor, transformToParent=None, jointAxis=(0.0, 0.0, 1.0)):
        self._create(name, predecessor, transformToParent, jointAxis)


class PrismaticJoint(Joint):
    """PrismaticJoint(name, predecessor, transformToParent | jointOffset, jointAxis)  (PrismaticJoint.java:34-51)"""
    _type = _capi.PRISMATIC

    def __init__(self, name, predecessor, transformToParent=None, jointAxis=(0.0, 0.0, 1.0)):
        self._create(name, predecessor, transformToParent, jointAxis)


class SixDoFJoint(Joint):
    """SixDoFJoint(name, predecessor[, transformToParent])  (SixDoFJoint.java:52-70)"""
    _type = _capi.SIXDOF

    def __init__(self, name, predecessor, transformToParent=None):
        self._create(name, predecessor, transformToParent, None)


class SphericalJoint(Joint):
    """SphericalJoint(name, predecessor[, transformToParent | jointOffset])  (SphericalJoint.java:43-69): 3 DoF; configuration rows
    [qx qy qz qs], velocity / acceleration / effort rows = the angular part in frameAfterJoint (SphericalJointReadOnly.java:31-71)."""
    _type = SPHERICAL

    def __init__(self, name, predecessor, transformToParent=None):
        self._create(name, predecessor, transformToParent, None)


class PlanarJoint(Joint):
    """PlanarJoint(name, predecessor[, transformToParent])  (PlanarJoint.java:37-61): 3 DoF in the x-z plane of frameBeforeJoint;
    configuration rows [pitch x z], velocity / acceleration / effort rows [w_y v_x v_z] in frameAfterJoint
    (PlanarJointReadOnly.java:20-58)."""
    _type = PLANAR

    def __init__(self, name, predecessor, transformToParent=None):
        self._create(name, predecessor, transformToParent, None)


class FixedJoint(Joint):
    """FixedJoint(name, predecessor[, transformToParent])  (M/multiBodySystem/FixedJoint.java:40-62): 0 DoF.  Host-only: the
    flattener welds its successor (inertia, children, offsets) into the nearest moving ancestor, so the GPU tables hold moving
    joints only and the calculators need no special case."""
    _type = FIXED

    def __init__(self, name, predecessor, transformToParent=None):
        self._create(name, predecessor, transformToParent, None)


class JointMatrixIndexProvider:
    def __init__(self, joints, dof, cfg):
        self._joints, self._dof, self._cfg = joints, dof, cfg

    def getIndexedJointsInOrder(self):
        return list(self._joints)

    def getJointDoFIndices(self, joint):
        i = self._joints.index(joint)
        return list(range(self._dof[i], self._dof[i] + joint.getDegreesOfFreedom()))

    def getJointConfigurationIndices(self, joint):
        i = self._joints.index(joint)
        return list(range(self._cfg[i], self._cfg[i] + joint.getConfigurationMatrixSize()))


class MultiBodySystem:
    """MultiBodySystemBasics: the tree below a root body with its joint order fixed."""

    def __init__(self, rootBody, jointsToIgnore=None):
        if not rootBody.isRootBody():
            raise ScrewTheoryException("toMultiBodySystemBasics expects the root body")
        m = rootBody._model
        ignore = np.ascontiguousarray([j._id for j in (jointsToIgnore or [])], dtype=np.int32)
        if m.finalized and (ignore.size or getattr(m, "ignored_ids", ())):
            raise ScrewTheoryException("this model was already turned into a system; build the tree again to choose other jointsToIgnore")
        m.check(lib.mecano_model_finalize_ignoring(m.h, ignore.ctypes.data_as(_ip), int(ignore.size)))
        m.finalized = True
        m.ignored_ids = tuple(int(i) for i in ignore)
        self._model, self._root = m, rootBody
        n_all = lib.mecano_model_n_joints(m.h)
        ids, dof, cfg = (np.zeros(n_all, np.int32) for _ in range(3))
        n = lib.mecano_model_joint_order(m.h, ids.ctypes.data_as(_ip), dof.ctypes.data_as(_ip), cfg.ctypes.data_as(_ip))
        ids, dof, cfg = ids[:n], dof[:n], cfg[:n]
        self._joints = [m.joints[i] for i in ids]
        self._ignored = [j for j in m.joints if j not in self._joints]
        self._provider = JointMatrixIndexProvider(self._joints, dof.tolist(), cfg.tolist())
        self._ndofs = lib.mecano_model_n_dofs(m.h)
        self._ncfg = lib.mecano_model_n_cfg(m.h)

    @staticmethod
    def toMultiBodySystemBasics(rootBody, jointsToIgnore=None):
        """MultiBodySystemBasics.toMultiBodySystemBasics(rootBody[, jointsToIgnore]) (MultiBodySystemBasics.java:76-142): a joint
        is ignored if it is listed or descends from a listed joint; the inertia of every ignored subtree is lumped into its
        parent body at the joints' stored configuration (Joint.setJointConfiguration)."""
        return MultiBodySystem(rootBody, jointsToIgnore)

    def hasWeldedBodies(self):
        """True if fixed joints or ignored subtrees were folded into other bodies at flatten time."""
        return bool(self._ignored) or any(j._type == FIXED for j in self._joints)

    def getRootBody(self):
        return self._root

    def getAllJoints(self):
        return list(self._joints)

    def getJointsToConsider(self):
        return list(self._joints)

    def getJointsToIgnore(self):
        return list(self._ignored)

    def getJointMatrixIndexProvider(self):
        return self._provider

    def getNumberOfDoFs(self):
        return self._ndofs

    def getConfigurationMatrixSize(self):
        return self._ncfg

    def getNumberOfJoints(self):
        return len(self._joints)

    # ---- what the engine consumes
    def tables(self):
        """ctypes pointer to the level-ordered mecano_b200_tree_desc (owned by the model)."""
        return lib.mecano_model_tables(self._model.h)

    def expanded(self):
        """The same system with nothing welded (mecano_model_expanded_tables): every joint of the tree is a body, the fixed /
        ignored joints are held (q = stored configuration, qd = qdd = 0, locked in forward dynamics).  The calculators run
        external wrenches and per-body results of systems with fixed / ignored joints on it.  Returns an object with
        tables() (as this system's), n_bodies, n_extra_dof, n_extra_cfg, n_extra_wrench_blocks, q_extra, locked,
        row_of_considered."""
        if getattr(self, "_expanded", None) is None:
            self._expanded = _ExpandedSystem(self)
        return self._expanded

    def tableRow(self, joint):
        """Row of `joint` in the tree description handed to the engine (-1: fixed, ignored or foreign joint)."""
        return int(lib.mecano_model_table_row(self._model.h, joint._id))

    def describe(self):
        """Plain numpy description in JointMatrixIndexProvider (depth-first) order, read back through the model
        getters (not through the flattener): used by the tests to feed the oracle."""
        if self.hasWeldedBodies():
            raise NotImplementedError("describe() lists one body per joint; this system has fixed / ignored joints folded into other bodies")
        nb = len(self._joints)
        out = {
            "nb": nb, "nv": self._ndofs, "nq": self._ncfg,
            "parent": np.zeros(nb, np.int32), "jtype": np.zeros(nb, np.int32), "axis": np.zeros((nb, 3)),
            "off_R": np.zeros((nb, 3, 3)), "off_p": np.zeros((nb, 3)), "com_R": np.zeros((nb, 3, 3)), "com_p": np.zeros((nb, 3)),
            "J": np.zeros((nb, 3, 3)), "mass": np.zeros(nb), "dof_off": np.zeros(nb, np.int32), "cfg_off": np.zeros(nb, np.int32),
        }
        for i, j in enumerate(self._joints):
            jt, pred, axis, t12, I9, mass, p12 = j._info()
            pb = self._model.bodies[pred]
            out["parent"][i] = -1 if pb.isRootBody() else self._joints.index(pb.getParentJoint())
            out["jtype"][i] = jt
            out["axis"][i] = axis
            out["off_R"][i] = t12[:9].reshape(3, 3)
            out["off_p"][i] = t12[9:]
            out["com_R"][i] = p12[:9].reshape(3, 3)
            out["com_p"][i] = p12[9:]
            out["J"][i] = I9.reshape(3, 3)
            out["mass"][i] = mass
            out["dof_off"][i] = self._provider._dof[i]
            out["cfg_off"][i] = self._provider._cfg[i]
        return out


class _ExpandedSystem:
    def __init__(self, system):
        m = system._model
        self._system = system  # keeps the model (owner of the tables) alive
        self._desc = lib.mecano_model_expanded_tables(m.h)
        if not self._desc:
            raise ScrewTheoryException(lib.mecano_model_last_error(m.h).decode())
        v = [ctypes.c_int32() for _ in range(4)]
        m.check(lib.mecano_model_expanded_info(m.h, *(ctypes.byref(x) for x in v)))
        self.n_bodies, self.n_extra_dof, self.n_extra_cfg, self.n_extra_wrench_blocks = (int(x.value) for x in v)
        self.q_extra = np.zeros(max(1, self.n_extra_cfg))
        self.locked = np.zeros(self.n_bodies, np.int32)
        self.row_of_considered = np.zeros(max(1, system.getNumberOfJoints()), np.int32)
        m.check(lib.mecano_model_expanded_fill(m.h, self.q_extra.ctypes.data_as(_dp), self.locked.ctypes.data_as(_ip), self.row_of_considered.ctypes.data_as(_ip)))
        self.q_extra = self.q_extra[:self.n_extra_cfg]
        self.row_of_considered = self.row_of_considered[:system.getNumberOfJoints()]

    def tables(self):
        return self._desc


class MultiBodySystemRandomTools:
    """Generators with Mecano's distributions (MultiBodySystemRandomTools.java:483-496, 908-923, 1211-1231, 1365-1371,
    1380-1486).  `seed` replaces java.util.Random (whose stream, through Euclid's random tools, is not reproducible
    here)."""

    @staticmethod
    def nextRevoluteJointChain(seed, rootBody, numberOfJoints):
        return MultiBodySystemRandomTools.nextOneDoFJointChain(seed, rootBody, numberOfJoints, 0.0)

    @staticmethod
    def nextOneDoFJointChain(seed, rootBody, numberOfJoints, prismaticFraction=0.5):
        m = rootBody._model
        m.check(lib.mecano_model_next_one_dof_joint_chain(m.h, int(seed), rootBody._id, int(numberOfJoints), float(prismaticFraction)))
        before = len(m.joints)
        m.adopt_generated()
        return m.joints[before:]

    @staticmethod
    def nextRevoluteJointTree(seed, rootBody, numberOfJoints):
        return MultiBodySystemRandomTools.nextOneDoFJointTree(seed, rootBody, numberOfJoints, 0.0)

    @staticmethod
    def nextOneDoFJointTree(seed, rootBody, numberOfJoints, prismaticFraction=0.5):
        m = rootBody._model
        m.check(lib.mecano_model_next_one_dof_joint_tree(m.h, int(seed), rootBody._id, int(numberOfJoints), float(prismaticFraction)))
        before = len(m.joints)
        m.adopt_generated()
        return m.joints[before:]

    @staticmethod
    def nextJointChain(seed, rootBody, numberOfJoints):
        """nextJointChain (MultiBodySystemRandomTools.java:424-440): joints of random types, all five moving joint types."""
        m = rootBody._model
        m.check(lib.mecano_model_next_joint_chain(m.h, int(seed), rootBody._id, int(numberOfJoints)))
        before = len(m.joints)
        m.adopt_generated()
        return m.joints[before:]

    @staticmethod
    def nextJointTree(seed, rootBody, numberOfJoints):
        """nextJointTree (:844-860)."""
        m = rootBody._model
        m.check(lib.mecano_model_next_joint_tree(m.h, int(seed), rootBody._id, int(numberOfJoints)))
        before = len(m.joints)
        m.adopt_generated()
        return m.joints[before:]

    @staticmethod
    def nextFloatingBase(seed, rootBody):
        m = rootBody._model
        m.check(lib.mecano_model_next_floating_base(m.h, int(seed), rootBody._id))
        m.adopt_generated()
        return m.joints[-1]

    @staticmethod
    def nextHumanoid(seed, rootBody, neckJoints=2):
        """SixDoF pelvis + 2 legs x 6 + spine 3 + 2 arms x 7 + neck: 37 DoF (neckJoints=2) or 36 DoF (1)."""
        m = rootBody._model
        m.check(lib.mecano_model_next_humanoid(m.h, int(seed), int(neckJoints)))
        before = len(m.joints)
        m.adopt_generated()
        return m.joints[before:]

    @staticmethod
    def nextState(rng, system, n):
        """Random joint states for n samples, as [rows, n] float64 numpy arrays (q, qd, qdd, tau), following
        MultiBodySystemRandomTools.nextState (:45-68): revolute q in [-pi, pi], SixDoF unit quaternion + position in
        [-1, 1]^3, velocity-like entries in [-1, 1]."""
        nv, nq = system.getNumberOfDoFs(), system.getConfigurationMatrixSize()
        q = rng.uniform(-np.pi, np.pi, size=(nq, n))
        prov = system.getJointMatrixIndexProvider()
        for j in system.getJointsToConsider():
            if isinstance(j, (SixDoFJoint, SphericalJoint)):
                c = prov.getJointConfigurationIndices(j)[0]
                quat = rng.normal(size=(4, n))
                quat /= np.linalg.norm(quat, axis=0)
                q[c:c + 4] = quat
                if isinstance(j, SixDoFJoint):
                    q[c + 4:c + 7] = rng.uniform(-1, 1, size=(3, n))
            elif isinstance(j, PlanarJoint):  # (pitch, x, z)
                c = prov.getJointConfigurationIndices(j)[0]
                q[c + 1:c + 3] = rng.uniform(-1, 1, size=(2, n))
        return tuple(np.ascontiguousarray(rng.uniform(-1, 1, size=(nv, n))) if k else np.ascontiguousarray(q) for k in range(4))
