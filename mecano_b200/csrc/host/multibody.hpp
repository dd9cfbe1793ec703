// multibody.hpp -- host-side mirror (C++) of the part of Mecano's model API that the hot path needs:
// RigidBody, RevoluteJoint, PrismaticJoint, SixDoFJoint, SphericalJoint, PlanarJoint, FixedJoint, MultiBodySystem + JointMatrixIndexProvider, the
// flattener producing the level-ordered tables of include/mecano_b200.h, and the synthetic generators of
// MultiBodySystemRandomTools.  ("M/" = /root/reference/src/main/java/us/ihmc/mecano/)
//
//   RigidBody          M/multiBodySystem/RigidBody.java:79-182
//   RevoluteJoint      M/multiBodySystem/RevoluteJoint.java:42-74
//   PrismaticJoint     M/multiBodySystem/PrismaticJoint.java:34-51
//   SixDoFJoint        M/multiBodySystem/SixDoFJoint.java:52-70
//   SphericalJoint     M/multiBodySystem/SphericalJoint.java:43-69, interfaces/SphericalJointReadOnly.java:31-71
//   PlanarJoint        M/multiBodySystem/PlanarJoint.java:37-61, interfaces/PlanarJointReadOnly.java:20-58
//   MultiBodySystem    M/multiBodySystem/interfaces/MultiBodySystemBasics.java:76-142, MultiBodySystemReadOnly.java:167-205
//   index provider     M/multiBodySystem/interfaces/JointMatrixIndexProvider.java:71-123
//   joint order        M/multiBodySystem/iterators/JointIterator.java:130-177 (depth-first pre-order, children in insertion order)
//
// Differences that follow from batching: joints hold no state (q, qd, qdd, tau are N-column matrices handed to
// the calculators) and there is no ReferenceFrame tree (the kernels rebuild frames from q for every state).
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/mecano_b200.h"

namespace mecano
{
struct Vector3D
{
   double x = 0, y = 0, z = 0;
};

struct Matrix3D
{
   double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; // row-major
   static Matrix3D diagonal(double a, double b, double c)
   {
      Matrix3D r;
      r.m[0] = a; r.m[4] = b; r.m[8] = c;
      return r;
   }
};

struct RigidBodyTransform
{
   Matrix3D rotation;
   Vector3D translation;
   RigidBodyTransform() = default;
   RigidBodyTransform(const Matrix3D &R, const Vector3D &t) : rotation(R), translation(t) {}
   explicit RigidBodyTransform(const Vector3D &t) : translation(t) {}
};

class ScrewTheoryException : public std::runtime_error // M/exceptions/ScrewTheoryException.java
{
 public:
   using std::runtime_error::runtime_error;
};

class RigidBody;

// Fixed joints (M/multiBodySystem/FixedJoint.java) exist on the host only: the flattener welds their successor into the
// nearest moving ancestor, so the kernels never see them.
enum class JointType
{
   Revolute = MECANO_B200_REVOLUTE, Prismatic = MECANO_B200_PRISMATIC, SixDoF = MECANO_B200_SIXDOF, Spherical = MECANO_B200_SPHERICAL,
   Planar = MECANO_B200_PLANAR, Fixed = 5
};

// ---- small rigid-transform / inertia algebra used when bodies are welded together at flatten time
inline Matrix3D matmul(const Matrix3D &A, const Matrix3D &B)
{
   Matrix3D C;
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
         C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
   return C;
}
inline Matrix3D transposed(const Matrix3D &A)
{
   Matrix3D C;
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
         C.m[3 * i + j] = A.m[3 * j + i];
   return C;
}
inline Vector3D matvec(const Matrix3D &A, const Vector3D &v)
{
   return Vector3D{A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z, A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
// a then b, both "child in parent": x_parent = A (B x + tb) + ta
inline RigidBodyTransform compose(const RigidBodyTransform &a, const RigidBodyTransform &b)
{
   RigidBodyTransform c;
   c.rotation = matmul(a.rotation, b.rotation);
   const Vector3D t = matvec(a.rotation, b.translation);
   c.translation = Vector3D{t.x + a.translation.x, t.y + a.translation.y, t.z + a.translation.z};
   return c;
}

class Joint
{
 public:
   virtual ~Joint() = default;
   const std::string &getName() const { return name_; }
   RigidBody *getPredecessor() const { return predecessor_; }
   RigidBody *getSuccessor() const { return successor_; }
   JointType getType() const { return type_; }
   virtual int getDegreesOfFreedom() const = 0;
   virtual int getConfigurationMatrixSize() const = 0;
   const Vector3D &getJointAxis() const { return axis_; }
   // transform from frameBeforeJoint to the predecessor's frameAfterJoint (identity when constructed with null)
   const RigidBodyTransform &getTransformToParent() const { return transformToParent_; }
   void setSuccessor(RigidBody *successor) { successor_ = successor; }
   // The configuration a joint has when it is ignored (MultiBodySystemBasics.toMultiBodySystemBasics(root, jointsToIgnore)):
   // Mecano lumps the inertia of an ignored subtree into its parent body at the configuration the joints have when the
   // calculator is built (InverseDynamicsCalculator.java:236, :832-860).  One-DoF: q; SixDoF: qx qy qz qs x y z; Spherical: qx qy qz qs;
   // Planar: pitch x z.  Default: zero / identity.
   void setJointConfiguration(const double *q, int n) { q_.assign(q, q + n); }
   const std::vector<double> &getJointConfiguration() const { return q_; }
   // frameAfterJoint in frameBeforeJoint at the stored configuration (MecanoFactories.java:231-260,
   // PrismaticJointReadOnly.java:18-22, FloatingJointReadOnly.java:34-37)
   RigidBodyTransform getJointTransform() const
   {
      RigidBodyTransform X;
      if (type_ == JointType::Revolute)
      {
         const double q = q_.empty() ? 0.0 : q_[0], c = std::cos(q), s = std::sin(q), t = 1.0 - c;
         const double x = axis_.x, y = axis_.y, z = axis_.z;
         const double R[9] = {t * x * x + c, t * x * y - s * z, t * x * z + s * y, t * x * y + s * z, t * y * y + c, t * y * z - s * x,
                              t * x * z - s * y, t * y * z + s * x, t * z * z + c};
         for (int i = 0; i < 9; i++) X.rotation.m[i] = R[i];
      }
      else if (type_ == JointType::Prismatic)
      {
         const double q = q_.empty() ? 0.0 : q_[0];
         X.translation = Vector3D{q * axis_.x, q * axis_.y, q * axis_.z};
      }
      else if (type_ == JointType::Planar && q_.size() == 3)
      {
         // PlanarJointReadOnly.java:40-48: rotation about y by the pitch, translation in the x-z plane
         const double c = std::cos(q_[0]), s = std::sin(q_[0]);
         const double R[9] = {c, 0, s, 0, 1, 0, -s, 0, c};
         for (int i = 0; i < 9; i++) X.rotation.m[i] = R[i];
         X.translation = Vector3D{q_[1], 0.0, q_[2]};
      }
      else if ((type_ == JointType::SixDoF && q_.size() == 7) || (type_ == JointType::Spherical && q_.size() == 4))
      {
         double qx = q_[0], qy = q_[1], qz = q_[2], qs = q_[3];
         const double n = std::sqrt(qx * qx + qy * qy + qz * qz + qs * qs);
         if (n > 1e-14)
         {
            qx /= n; qy /= n; qz /= n; qs /= n;
            const double R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qs * qz), 2 * (qx * qz + qs * qy), 2 * (qx * qy + qs * qz),
                                 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qs * qx), 2 * (qx * qz - qs * qy), 2 * (qy * qz + qs * qx),
                                 1 - 2 * (qx * qx + qy * qy)};
            for (int i = 0; i < 9; i++) X.rotation.m[i] = R[i];
         }
         if (type_ == JointType::SixDoF)
            X.translation = Vector3D{q_[4], q_[5], q_[6]};
      }
      return X;
   }

 protected:
   Joint(std::string name, RigidBody *predecessor, const RigidBodyTransform *transformToParent, JointType type);
   std::string name_;
   RigidBody *predecessor_;
   RigidBody *successor_ = nullptr;
   RigidBodyTransform transformToParent_;
   JointType type_;
   Vector3D axis_{0, 0, 1};
   std::vector<double> q_;
};

class RigidBody
{
 public:
   // root body ("elevator"), RigidBody.java:79-82
   explicit RigidBody(std::string name) : name_(std::move(name)) {}
   // RigidBody(name, parentJoint, Ixx, Iyy, Izz, mass, centerOfMassOffset), RigidBody.java:123-128
   RigidBody(std::string name, Joint *parentJoint, double Ixx, double Iyy, double Izz, double mass, const Vector3D &centerOfMassOffset)
       : RigidBody(std::move(name), parentJoint, Matrix3D::diagonal(Ixx, Iyy, Izz), mass, RigidBodyTransform(centerOfMassOffset))
   {
   }
   // RigidBody(name, parentJoint, momentOfInertia, mass, centerOfMassOffset), RigidBody.java:141-146
   RigidBody(std::string name, Joint *parentJoint, const Matrix3D &momentOfInertia, double mass, const Vector3D &centerOfMassOffset)
       : RigidBody(std::move(name), parentJoint, momentOfInertia, mass, RigidBodyTransform(centerOfMassOffset))
   {
   }
   // RigidBody(name, parentJoint, momentOfInertia, mass, inertiaPose), RigidBody.java:163-168
   RigidBody(std::string name, Joint *parentJoint, const Matrix3D &momentOfInertia, double mass, const RigidBodyTransform &inertiaPose)
       : name_(std::move(name)), parentJoint_(parentJoint), momentOfInertia_(momentOfInertia), mass_(mass), inertiaPose_(inertiaPose)
   {
      if (!parentJoint)
         throw std::invalid_argument("parentJoint can not be null");
      parentJoint->setSuccessor(this);
   }
   const std::string &getName() const { return name_; }
   bool isRootBody() const { return parentJoint_ == nullptr; }
   Joint *getParentJoint() const { return parentJoint_; }
   const std::vector<Joint *> &getChildrenJoints() const { return children_; }
   void addChildJoint(Joint *j) { children_.push_back(j); }
   const Matrix3D &getMomentOfInertia() const { return momentOfInertia_; }
   double getMass() const { return mass_; }
   const RigidBodyTransform &getInertiaPose() const { return inertiaPose_; }

 private:
   std::string name_;
   Joint *parentJoint_ = nullptr;
   std::vector<Joint *> children_;
   Matrix3D momentOfInertia_;
   double mass_ = 0;
   RigidBodyTransform inertiaPose_;
};

inline Joint::Joint(std::string name, RigidBody *predecessor, const RigidBodyTransform *transformToParent, JointType type)
    : name_(std::move(name)), predecessor_(predecessor), type_(type)
{
   if (!predecessor)
      throw std::invalid_argument("predecessor can not be null");
   if (transformToParent)
      transformToParent_ = *transformToParent;
   predecessor->addChildJoint(this);
}

class OneDoFJoint : public Joint
{
 public:
   int getDegreesOfFreedom() const override { return 1; }
   int getConfigurationMatrixSize() const override { return 1; }

 protected:
   OneDoFJoint(std::string name, RigidBody *predecessor, const RigidBodyTransform *transformToParent, const Vector3D &axis, JointType type)
       : Joint(std::move(name), predecessor, transformToParent, type)
   {
      const double n = std::sqrt(axis.x * axis.x + axis.y * axis.y + axis.z * axis.z);
      if (!(n > 0))
         throw std::invalid_argument("joint axis can not be zero");
      axis_ = Vector3D{axis.x / n, axis.y / n, axis.z / n};
   }
};

class RevoluteJoint : public OneDoFJoint
{
 public:
   RevoluteJoint(std::string name, RigidBody *predecessor, const Vector3D &jointAxis)
       : OneDoFJoint(std::move(name), predecessor, nullptr, jointAxis, JointType::Revolute) {}
   RevoluteJoint(std::string name, RigidBody *predecessor, const Vector3D &jointOffset, const Vector3D &jointAxis)
       : RevoluteJoint(std::move(name), predecessor, RigidBodyTransform(jointOffset), jointAxis) {}
   RevoluteJoint(std::string name, RigidBody *predecessor, const RigidBodyTransform &transformToParent, const Vector3D &jointAxis)
       : OneDoFJoint(std::move(name), predecessor, &transformToParent, jointAxis, JointType::Revolute) {}
};

class PrismaticJoint : public OneDoFJoint
{
 public:
   PrismaticJoint(std::string name, RigidBody *predecessor, const Vector3D &jointOffset, const Vector3D &jointAxis)
       : PrismaticJoint(std::move(name), predecessor, RigidBodyTransform(jointOffset), jointAxis) {}
   PrismaticJoint(std::string name, RigidBody *predecessor, const RigidBodyTransform &transformToParent, const Vector3D &jointAxis)
       : OneDoFJoint(std::move(name), predecessor, &transformToParent, jointAxis, JointType::Prismatic) {}
};

class SixDoFJoint : public Joint
{
 public:
   SixDoFJoint(std::string name, RigidBody *predecessor) : Joint(std::move(name), predecessor, nullptr, JointType::SixDoF) {}
   SixDoFJoint(std::string name, RigidBody *predecessor, const RigidBodyTransform &transformToParent)
       : Joint(std::move(name), predecessor, &transformToParent, JointType::SixDoF) {}
   int getDegreesOfFreedom() const override { return 6; }
   int getConfigurationMatrixSize() const override { return 7; }
};

// SphericalJoint (M/multiBodySystem/SphericalJoint.java:43-69): 3 DoF, configuration = orientation quaternion (qx qy qz qs),
// velocity-like rows = angular part in frameAfterJoint (SphericalJointReadOnly.java:31-71)
class SphericalJoint : public Joint
{
 public:
   SphericalJoint(std::string name, RigidBody *predecessor) : Joint(std::move(name), predecessor, nullptr, JointType::Spherical) {}
   SphericalJoint(std::string name, RigidBody *predecessor, const Vector3D &jointOffset) : SphericalJoint(std::move(name), predecessor, RigidBodyTransform(jointOffset)) {}
   SphericalJoint(std::string name, RigidBody *predecessor, const RigidBodyTransform &transformToParent)
       : Joint(std::move(name), predecessor, &transformToParent, JointType::Spherical) {}
   int getDegreesOfFreedom() const override { return 3; }
   int getConfigurationMatrixSize() const override { return 4; }
};

// PlanarJoint (M/multiBodySystem/PlanarJoint.java:37-61): 3 DoF in the x-z plane of frameBeforeJoint, configuration (pitch, x, z),
// velocity-like rows (w_y, v_x, v_z) in frameAfterJoint (PlanarJointReadOnly.java:20-58)
class PlanarJoint : public Joint
{
 public:
   PlanarJoint(std::string name, RigidBody *predecessor) : Joint(std::move(name), predecessor, nullptr, JointType::Planar) {}
   PlanarJoint(std::string name, RigidBody *predecessor, const RigidBodyTransform &transformToParent)
       : Joint(std::move(name), predecessor, &transformToParent, JointType::Planar) {}
   int getDegreesOfFreedom() const override { return 3; }
   int getConfigurationMatrixSize() const override { return 3; }
};

// FixedJoint (M/multiBodySystem/FixedJoint.java:40-62): 0 DoF, welds its successor to its predecessor
class FixedJoint : public Joint
{
 public:
   FixedJoint(std::string name, RigidBody *predecessor) : Joint(std::move(name), predecessor, nullptr, JointType::Fixed) {}
   FixedJoint(std::string name, RigidBody *predecessor, const RigidBodyTransform &transformToParent)
       : Joint(std::move(name), predecessor, &transformToParent, JointType::Fixed) {}
   int getDegreesOfFreedom() const override { return 0; }
   int getConfigurationMatrixSize() const override { return 0; }
};

// MultiBodySystemBasics.toMultiBodySystemBasics(rootBody[, jointsToIgnore]) + JointMatrixIndexProvider
class MultiBodySystem
{
 public:
   // A joint is ignored if it is in jointsToIgnore or a descendant of one (MultiBodySystemReadOnly.java:284-300)
   static MultiBodySystem toMultiBodySystemBasics(RigidBody *rootBody, const std::vector<Joint *> &jointsToIgnore = {})
   {
      if (!rootBody)
         throw std::invalid_argument("rootBody can not be null");
      MultiBodySystem s;
      s.root_ = rootBody;
      auto ignored = [&](const Joint *j) {
         for (const Joint *k : jointsToIgnore)
            if (k == j)
               return true;
         return false;
      };
      // SubtreeStreams.fromChildren: depth-first pre-order (JointIterator.java:153-162)
      std::vector<Joint *> stack(rootBody->getChildrenJoints().rbegin(), rootBody->getChildrenJoints().rend());
      while (!stack.empty())
      {
         Joint *j = stack.back();
         stack.pop_back();
         if (!j->getSuccessor())
            throw ScrewTheoryException("joint " + j->getName() + " has no successor");
         if (ignored(j))
         {
            s.ignored_.push_back(j); // the roots of the ignored subtrees
            continue;
         }
         s.dofIndex_.push_back(s.nDoFs_);
         s.cfgIndex_.push_back(s.nCfg_);
         s.nDoFs_ += j->getDegreesOfFreedom();
         s.nCfg_ += j->getConfigurationMatrixSize();
         s.joints_.push_back(j);
         const auto &ch = j->getSuccessor()->getChildrenJoints();
         for (auto it = ch.rbegin(); it != ch.rend(); ++it)
            stack.push_back(*it);
      }
      return s;
   }
   bool isIgnoredSubtreeRoot(const Joint *j) const
   {
      for (const Joint *k : ignored_)
         if (k == j)
            return true;
      return false;
   }
   const std::vector<Joint *> &getIgnoredSubtreeRoots() const { return ignored_; }
   RigidBody *getRootBody() const { return root_; }
   const std::vector<Joint *> &getJointsToConsider() const { return joints_; } // == getIndexedJointsInOrder()
   int getNumberOfDoFs() const { return nDoFs_; }
   int getConfigurationMatrixSize() const { return nCfg_; }
   int indexOf(const Joint *j) const
   {
      for (size_t i = 0; i < joints_.size(); i++)
         if (joints_[i] == j)
            return (int)i;
      return -1;
   }
   // JointMatrixIndexProvider.getJointDoFIndices / getJointConfigurationIndices: first row of the joint
   int getJointDoFIndex(const Joint *j) const { return dofIndex_.at((size_t)indexOf(j)); }
   int getJointConfigurationIndex(const Joint *j) const { return cfgIndex_.at((size_t)indexOf(j)); }
   int dofIndexAt(size_t i) const { return dofIndex_[i]; }
   int cfgIndexAt(size_t i) const { return cfgIndex_[i]; }

 private:
   RigidBody *root_ = nullptr;
   std::vector<Joint *> joints_, ignored_;
   std::vector<int> dofIndex_, cfgIndex_;
   int nDoFs_ = 0, nCfg_ = 0;
};

// Level-ordered parent / joint-type / axis / inertia / transform tables: the argument of mecano_b200_create.
struct FlatTables
{
   std::vector<int32_t> level_start, parent, joint_type, dof_offset, cfg_offset, wrench_index;
   std::vector<double> axis, offset_rot, offset_pos, com_rot, com_pos, inertia, mass;
   std::vector<int> body_of_joint; // DFS joint index -> row in the tables
   mecano_b200_tree_desc desc{};

   // mass, first moment and second moment (about the frame origin, frame axes) of a set of bodies welded together
   struct Lump
   {
      double m = 0, h[3] = {0, 0, 0}, I[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      // body with inertia J about its CoM (CoM frame axes), CoM pose P in the body frame, body frame at X in the lump frame
      void add(const RigidBody &b, const RigidBodyTransform &X)
      {
         const RigidBodyTransform C = compose(X, b.getInertiaPose()); // CoM frame in the lump frame
         const Matrix3D J = matmul(matmul(C.rotation, b.getMomentOfInertia()), transposed(C.rotation));
         const double mass = b.getMass(), c[3] = {C.translation.x, C.translation.y, C.translation.z};
         const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
         m += mass;
         for (int i = 0; i < 3; i++)
         {
            h[i] += mass * c[i];
            for (int j = 0; j < 3; j++)
               I[3 * i + j] += J.m[3 * i + j] + mass * ((i == j ? cc : 0.0) - c[i] * c[j]);
         }
      }
   };

   // Everything rigidly attached to `body` (whose frame is at X in the lump frame): the body itself, the successors of its
   // FixedJoint children and -- at their stored configuration -- whole ignored subtrees (computeSubtreeInertia,
   // M/tools/MultiBodySystemTools.java:32-64; InverseDynamicsCalculator.java:832-860).
   static void weld(const MultiBodySystem &sys, const RigidBody &body, const RigidBodyTransform &X, bool whole_subtree, Lump &lump)
   {
      lump.add(body, X);
      for (const Joint *c : body.getChildrenJoints())
      {
         const bool ign = whole_subtree || sys.isIgnoredSubtreeRoot(c);
         if (!ign && c->getType() != JointType::Fixed)
            continue; // a moving, considered joint: its successor is a body of its own
         if (!c->getSuccessor())
            continue;
         RigidBodyTransform Xc = compose(X, c->getTransformToParent());
         if (ign)
            Xc = compose(Xc, c->getJointTransform());
         weld(sys, *c->getSuccessor(), Xc, ign, lump);
      }
   }

   static FlatTables flatten(const MultiBodySystem &sys)
   {
      FlatTables f;
      const auto &joints = sys.getJointsToConsider();
      const int nj = (int)joints.size();
      // moving joints become the bodies of the tables; fixed joints are folded into the offset of the moving joints below them
      std::vector<int> depth(nj, 0), parent_dfs(nj, -1);
      std::vector<RigidBodyTransform> offset((size_t)nj);
      int nlev = 0, nb = 0;
      for (int i = 0; i < nj; i++)
      {
         if (joints[i]->getType() == JointType::Fixed)
            continue;
         nb++;
         RigidBodyTransform X = joints[i]->getTransformToParent();
         const RigidBody *pred = joints[i]->getPredecessor();
         while (!pred->isRootBody() && pred->getParentJoint()->getType() == JointType::Fixed)
         {
            X = compose(pred->getParentJoint()->getTransformToParent(), X);
            pred = pred->getParentJoint()->getPredecessor();
         }
         offset[(size_t)i] = X;
         parent_dfs[i] = pred->isRootBody() ? -1 : sys.indexOf(pred->getParentJoint());
         depth[i] = parent_dfs[i] < 0 ? 0 : depth[parent_dfs[i]] + 1;
         nlev = std::max(nlev, depth[i] + 1);
      }
      std::vector<int> order; // table row -> DFS joint index, level by level
      f.level_start.assign(nlev + 1, 0);
      for (int l = 0; l < nlev; l++)
      {
         f.level_start[l] = (int)order.size();
         for (int i = 0; i < nj; i++)
            if (joints[i]->getType() != JointType::Fixed && depth[i] == l)
               order.push_back(i);
      }
      f.level_start[nlev] = nb;
      f.body_of_joint.assign(nj, -1);
      for (int r = 0; r < nb; r++)
         f.body_of_joint[order[r]] = r;
      for (int r = 0; r < nb; r++)
      {
         const int i = order[r];
         const Joint *j = joints[i];
         const RigidBody *b = j->getSuccessor();
         f.parent.push_back(parent_dfs[i] < 0 ? -1 : f.body_of_joint[parent_dfs[i]]);
         f.joint_type.push_back((int)j->getType());
         f.dof_offset.push_back(sys.dofIndexAt(i));
         f.cfg_offset.push_back(sys.cfgIndexAt(i));
         f.wrench_index.push_back(i); // external wrenches are handed over in joint (index-provider) order
         f.axis.insert(f.axis.end(), {j->getJointAxis().x, j->getJointAxis().y, j->getJointAxis().z});
         const RigidBodyTransform &T = offset[(size_t)i];
         f.offset_rot.insert(f.offset_rot.end(), T.rotation.m, T.rotation.m + 9);
         f.offset_pos.insert(f.offset_pos.end(), {T.translation.x, T.translation.y, T.translation.z});
         bool welded = false;
         for (const Joint *c : b->getChildrenJoints())
            welded = welded || c->getType() == JointType::Fixed || sys.isIgnoredSubtreeRoot(c);
         if (!welded)
         {
            const RigidBodyTransform &P = b->getInertiaPose();
            f.com_rot.insert(f.com_rot.end(), P.rotation.m, P.rotation.m + 9);
            f.com_pos.insert(f.com_pos.end(), {P.translation.x, P.translation.y, P.translation.z});
            f.inertia.insert(f.inertia.end(), b->getMomentOfInertia().m, b->getMomentOfInertia().m + 9);
            f.mass.push_back(b->getMass());
            continue;
         }
         // the body plus what is welded to it, as one rigid body: CoM frame = joint frame axes at the common CoM
         Lump lump;
         weld(sys, *b, RigidBodyTransform(), false, lump);
         double c[3] = {0, 0, 0};
         if (lump.m > 0)
            for (int k = 0; k < 3; k++)
               c[k] = lump.h[k] / lump.m;
         const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
         Matrix3D Jc, I3;
         for (int a = 0; a < 3; a++)
            for (int bq = 0; bq < 3; bq++)
               Jc.m[3 * a + bq] = lump.I[3 * a + bq] - lump.m * ((a == bq ? cc : 0.0) - c[a] * c[bq]);
         f.com_rot.insert(f.com_rot.end(), I3.m, I3.m + 9);
         f.com_pos.insert(f.com_pos.end(), {c[0], c[1], c[2]});
         f.inertia.insert(f.inertia.end(), Jc.m, Jc.m + 9);
         f.mass.push_back(lump.m);
      }
      f.bind(sys.getNumberOfDoFs(), sys.getConfigurationMatrixSize());
      return f;
   }

   // ---- the same system with nothing welded: every joint of the tree is a body of the tables, and the joints that flatten()
   // folds away are HELD instead -- a FixedJoint becomes a revolute joint kept at q = 0, an ignored joint keeps its own type at
   // its stored configuration, all with qd = qdd = 0 (the calculators lock them: ACCELERATION_SOURCE with zero acceleration).
   // Their configuration / DoF rows are appended BEHIND the system's own rows, their wrench rows behind the considered joints',
   // so the caller's matrices are the leading rows of the expanded ones.  This is what external wrenches and per-body results
   // on systems with fixed / ignored joints run on (InverseDynamicsCalculator.java:469-472 with :832-860): the held bodies exist,
   // so a wrench on a FixedJoint's successor acts where it is applied and every body reports its acceleration in its own frame.
   struct Expanded
   {
      int n_extra_dof = 0, n_extra_cfg = 0, n_extra_bodies = 0; // rows / wrench blocks appended for held joints
      std::vector<double> q_extra;                              // [n_extra_cfg] configuration of the held joints
      std::vector<int32_t> locked;                              // [n_bodies] 1 = held joint
      std::vector<int32_t> row_of_considered;                   // [considered joints] -> row in these tables
   };
   static FlatTables flattenExpanded(const MultiBodySystem &sys, Expanded &x)
   {
      FlatTables f;
      const int nv = sys.getNumberOfDoFs(), nq = sys.getConfigurationMatrixSize(), nj = (int)sys.getJointsToConsider().size();
      x = Expanded();
      x.row_of_considered.assign((size_t)nj, -1);
      struct Item
      {
         const Joint *j;
         int parent_row;
         bool ignored;
      };
      std::vector<Item> stack;
      const auto &rc = sys.getRootBody()->getChildrenJoints();
      for (auto it = rc.rbegin(); it != rc.rend(); ++it)
         stack.push_back({*it, -1, sys.isIgnoredSubtreeRoot(*it)});
      while (!stack.empty())
      {
         const Item it = stack.back();
         stack.pop_back();
         const Joint *j = it.j;
         const RigidBody *b = j->getSuccessor();
         if (!b)
            throw ScrewTheoryException("joint " + j->getName() + " has no successor");
         const int row = (int)f.parent.size();
         const bool fixed = j->getType() == JointType::Fixed;
         const bool held = fixed || it.ignored;
         f.parent.push_back(it.parent_row);
         f.joint_type.push_back(fixed ? (int)JointType::Revolute : (int)j->getType());
         const int nd = fixed ? 1 : j->getDegreesOfFreedom(), nc = fixed ? 1 : j->getConfigurationMatrixSize();
         if (held)
         {
            f.dof_offset.push_back(nv + x.n_extra_dof);
            f.cfg_offset.push_back(nq + x.n_extra_cfg);
            x.n_extra_dof += nd;
            x.n_extra_cfg += nc;
            // stored configuration (default: zero / identity orientation)
            std::vector<double> q0 = fixed ? std::vector<double>{0.0} : j->getJointConfiguration();
            if ((int)q0.size() != nc)
            {
               q0.assign((size_t)nc, 0.0);
               if (j->getType() == JointType::SixDoF || j->getType() == JointType::Spherical)
                  q0[3] = 1.0;
            }
            x.q_extra.insert(x.q_extra.end(), q0.begin(), q0.end());
         }
         else
         {
            const int i = sys.indexOf(j);
            f.dof_offset.push_back(sys.dofIndexAt((size_t)i));
            f.cfg_offset.push_back(sys.cfgIndexAt((size_t)i));
         }
         if (it.ignored)
            f.wrench_index.push_back(nj + x.n_extra_bodies++); // a zero block behind the caller's wrench rows
         else
         {
            const int i = sys.indexOf(j);
            f.wrench_index.push_back(i);
            x.row_of_considered[(size_t)i] = row;
         }
         x.locked.push_back(held ? 1 : 0);
         const Vector3D ax = fixed ? Vector3D{0, 0, 1} : j->getJointAxis();
         f.axis.insert(f.axis.end(), {ax.x, ax.y, ax.z});
         const RigidBodyTransform &T = j->getTransformToParent();
         f.offset_rot.insert(f.offset_rot.end(), T.rotation.m, T.rotation.m + 9);
         f.offset_pos.insert(f.offset_pos.end(), {T.translation.x, T.translation.y, T.translation.z});
         const RigidBodyTransform &P = b->getInertiaPose();
         f.com_rot.insert(f.com_rot.end(), P.rotation.m, P.rotation.m + 9);
         f.com_pos.insert(f.com_pos.end(), {P.translation.x, P.translation.y, P.translation.z});
         f.inertia.insert(f.inertia.end(), b->getMomentOfInertia().m, b->getMomentOfInertia().m + 9);
         f.mass.push_back(b->getMass());
         const auto &ch = b->getChildrenJoints();
         for (auto c = ch.rbegin(); c != ch.rend(); ++c)
            stack.push_back({*c, row, it.ignored || sys.isIgnoredSubtreeRoot(*c)});
      }
      f.bind(nv + x.n_extra_dof, nq + x.n_extra_cfg); // depth-first listing, no level table (the C ABI accepts any topological order)
      return f;
   }

   void bind(int ndofs, int ncfg)
   {
      desc.struct_size = (int32_t)sizeof(mecano_b200_tree_desc);
      desc.n_bodies = (int32_t)parent.size();
      desc.n_dofs = ndofs;
      desc.n_cfg = ncfg;
      desc.n_levels = level_start.empty() ? 0 : (int32_t)level_start.size() - 1;
      desc.level_start = level_start.empty() ? nullptr : level_start.data();
      desc.parent = parent.data();
      desc.joint_type = joint_type.data();
      desc.axis = axis.data();
      desc.offset_rot = offset_rot.data();
      desc.offset_pos = offset_pos.data();
      desc.com_rot = com_rot.data();
      desc.com_pos = com_pos.data();
      desc.inertia = inertia.data();
      desc.mass = mass.data();
      desc.dof_offset = dof_offset.data();
      desc.cfg_offset = cfg_offset.data();
      desc.wrench_index = wrench_index.empty() ? nullptr : wrench_index.data();
   }
};

// ---------------------------------------------------------------------------------------------------------
// Synthetic generators with Mecano's distributions (MultiBodySystemRandomTools.java:483-496, 908-923,
// 1211-1231, 1365-1371; MecanoRandomTools.java:623-647).  java.util.Random / Euclid's random tools are not
// reproducible here, so a fixed 64-bit generator (splitmix64) is used and the seed is reported by the bench.
class Random
{
 public:
   explicit Random(uint64_t seed) : s_(seed) {}
   uint64_t nextLong()
   {
      uint64_t z = (s_ += 0x9e3779b97f4a7c15ull);
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
      return z ^ (z >> 31);
   }
   double nextDouble() { return (double)(nextLong() >> 11) * (1.0 / 9007199254740992.0); }
   double nextDouble(double lo, double hi) { return lo + (hi - lo) * nextDouble(); }
   int nextInt(int bound) { return (int)(nextLong() % (uint64_t)bound); }
   double nextGaussian()
   {
      const double u1 = 1.0 - nextDouble(), u2 = nextDouble();
      return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
   }

 private:
   uint64_t s_;
};

// Owns the joints and bodies of one generated (or hand-built) system.
class MultiBodyArena
{
 public:
   RigidBody *newRootBody(const std::string &name = "elevator")
   {
      bodies_.push_back(std::make_unique<RigidBody>(name));
      return bodies_.back().get();
   }
   template <class J, class... A> J *newJoint(A &&...a)
   {
      auto p = std::make_unique<J>(std::forward<A>(a)...);
      J *raw = p.get();
      joints_.push_back(std::move(p));
      return raw;
   }
   template <class... A> RigidBody *newRigidBody(A &&...a)
   {
      bodies_.push_back(std::make_unique<RigidBody>(std::forward<A>(a)...));
      return bodies_.back().get();
   }

 private:
   std::vector<std::unique_ptr<RigidBody>> bodies_;
   std::vector<std::unique_ptr<Joint>> joints_;
};

namespace MultiBodySystemRandomTools
{
inline Vector3D nextVector3D(Random &r) { return Vector3D{r.nextDouble(-1, 1), r.nextDouble(-1, 1), r.nextDouble(-1, 1)}; }
inline Vector3D nextUnitVector3D(Random &r)
{
   double x, y, z, n;
   do
   {
      x = r.nextGaussian(); y = r.nextGaussian(); z = r.nextGaussian();
      n = std::sqrt(x * x + y * y + z * z);
   } while (n < 1e-6);
   return Vector3D{x / n, y / n, z / n};
}
inline Matrix3D nextRotationMatrix(Random &r)
{
   double q[4], n = 0;
   for (double &v : q) { v = r.nextGaussian(); n += v * v; }
   n = std::sqrt(n);
   const double x = q[0] / n, y = q[1] / n, z = q[2] / n, s = q[3] / n;
   Matrix3D R;
   R.m[0] = 1 - 2 * (y * y + z * z); R.m[1] = 2 * (x * y - s * z); R.m[2] = 2 * (x * z + s * y);
   R.m[3] = 2 * (x * y + s * z); R.m[4] = 1 - 2 * (x * x + z * z); R.m[5] = 2 * (y * z - s * x);
   R.m[6] = 2 * (x * z - s * y); R.m[7] = 2 * (y * z + s * x); R.m[8] = 1 - 2 * (x * x + y * y);
   return R;
}
inline RigidBodyTransform nextRigidBodyTransform(Random &r) { return RigidBodyTransform(nextRotationMatrix(r), nextVector3D(r)); }
// MecanoRandomTools.nextSymmetricPositiveDefiniteMatrix3D(random, 1e-4, 2.0, 0.5): L L^T
inline Matrix3D nextSymmetricPositiveDefiniteMatrix3D(Random &r, double minDiag = 1e-4, double maxDiag = 2.0, double offDiag = 0.5)
{
   double L[9] = {0};
   L[0] = r.nextDouble(minDiag, maxDiag);
   L[3] = r.nextDouble(-offDiag, offDiag);
   L[4] = r.nextDouble(minDiag, maxDiag);
   L[6] = r.nextDouble(-offDiag, offDiag);
   L[7] = r.nextDouble(-offDiag, offDiag);
   L[8] = r.nextDouble(minDiag, maxDiag);
   Matrix3D M;
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
         M.m[3 * i + j] = L[3 * i] * L[3 * j] + L[3 * i + 1] * L[3 * j + 1] + L[3 * i + 2] * L[3 * j + 2];
   return M;
}
// nextRigidBody, MultiBodySystemRandomTools.java:1365-1371
inline RigidBody *nextRigidBody(Random &r, MultiBodyArena &a, const std::string &name, Joint *parentJoint)
{
   const Matrix3D I = nextSymmetricPositiveDefiniteMatrix3D(r);
   const double mass = 0.1 + r.nextDouble();
   const Vector3D com = nextVector3D(r);
   return a.newRigidBody(name, parentJoint, I, mass, com);
}
// nextRevoluteJoint / nextPrismaticJoint, :1211-1231 (null offset for joints attached to the root body)
inline Joint *nextOneDoFJoint(Random &r, MultiBodyArena &a, const std::string &name, RigidBody *predecessor, bool prismatic)
{
   const Vector3D axis = nextUnitVector3D(r);
   if (predecessor->isRootBody())
   {
      if (prismatic) return a.newJoint<PrismaticJoint>(name, predecessor, RigidBodyTransform(), axis);
      return a.newJoint<RevoluteJoint>(name, predecessor, axis);
   }
   const RigidBodyTransform T = nextRigidBodyTransform(r);
   if (prismatic) return a.newJoint<PrismaticJoint>(name, predecessor, T, axis);
   return a.newJoint<RevoluteJoint>(name, predecessor, T, axis);
}
// nextJoint (:1116-1136 style): a joint of a random type among all the moving joint types
inline Joint *nextJoint(Random &r, MultiBodyArena &a, const std::string &name, RigidBody *predecessor)
{
   const int kind = r.nextInt(5);
   if (kind < 2)
      return nextOneDoFJoint(r, a, name, predecessor, kind == 1);
   const RigidBodyTransform T = predecessor->isRootBody() ? RigidBodyTransform() : nextRigidBodyTransform(r);
   if (kind == 2) return a.newJoint<SixDoFJoint>(name, predecessor, T);
   if (kind == 3) return a.newJoint<SphericalJoint>(name, predecessor, T);
   return a.newJoint<PlanarJoint>(name, predecessor, T);
}
// nextJointChain (:424-440): a chain of joints of random types
inline RigidBody *nextJointChain(Random &r, MultiBodyArena &a, const std::string &prefix, RigidBody *root, int n)
{
   RigidBody *pred = root;
   for (int i = 0; i < n; i++)
   {
      Joint *j = nextJoint(r, a, prefix + "Joint" + std::to_string(i), pred);
      pred = nextRigidBody(r, a, prefix + "Body" + std::to_string(i), j);
   }
   return pred;
}
// nextJointTree (:844-860): a tree of joints of random types
inline void nextJointTree(Random &r, MultiBodyArena &a, const std::string &prefix, RigidBody *root, int n)
{
   std::vector<RigidBody *> successors;
   RigidBody *pred = root;
   for (int i = 0; i < n; i++)
   {
      Joint *j = nextJoint(r, a, prefix + "Joint" + std::to_string(i), pred);
      successors.push_back(nextRigidBody(r, a, prefix + "Body" + std::to_string(i), j));
      pred = successors[(size_t)r.nextInt((int)successors.size())];
   }
}
// nextRevoluteJointChain / nextOneDoFJointChain, :483-496
inline RigidBody *nextOneDoFJointChain(Random &r, MultiBodyArena &a, const std::string &prefix, RigidBody *root, int n, double prismaticFraction = 0.0)
{
   RigidBody *pred = root;
   for (int i = 0; i < n; i++)
   {
      Joint *j = nextOneDoFJoint(r, a, prefix + "Joint" + std::to_string(i), pred, r.nextDouble() < prismaticFraction);
      pred = nextRigidBody(r, a, prefix + "Body" + std::to_string(i), j);
   }
   return pred;
}
// nextRevoluteJointTree / nextOneDoFJointTree, :908-923
inline void nextOneDoFJointTree(Random &r, MultiBodyArena &a, const std::string &prefix, RigidBody *root, int n, double prismaticFraction = 0.0)
{
   std::vector<RigidBody *> successors;
   RigidBody *pred = root;
   for (int i = 0; i < n; i++)
   {
      Joint *j = nextOneDoFJoint(r, a, prefix + "Joint" + std::to_string(i), pred, r.nextDouble() < prismaticFraction);
      successors.push_back(nextRigidBody(r, a, prefix + "Body" + std::to_string(i), j));
      pred = successors[(size_t)r.nextInt((int)successors.size())];
   }
}
// RandomFloatingRevoluteJointChain, :1380-1486
inline RigidBody *nextFloatingBase(Random &r, MultiBodyArena &a, RigidBody *elevator, const std::string &name = "root")
{
   SixDoFJoint *j = a.newJoint<SixDoFJoint>(name + "Joint", elevator);
   return nextRigidBody(r, a, name + "Body", j);
}
// Humanoid of SURVEY.md 8(d): SixDoF pelvis, 2 legs x 6, spine 3, 2 arms x 7 on the last spine body, neck
// (2 revolute -> 37 DoF "H37", 1 -> 36 DoF "H36").
inline void nextHumanoid(Random &r, MultiBodyArena &a, RigidBody *elevator, int neckJoints = 2)
{
   RigidBody *pelvis = nextFloatingBase(r, a, elevator, "pelvis");
   nextOneDoFJointChain(r, a, "leftLeg", pelvis, 6);
   nextOneDoFJointChain(r, a, "rightLeg", pelvis, 6);
   RigidBody *chest = nextOneDoFJointChain(r, a, "spine", pelvis, 3);
   nextOneDoFJointChain(r, a, "leftArm", chest, 7);
   nextOneDoFJointChain(r, a, "rightArm", chest, 7);
   nextOneDoFJointChain(r, a, "neck", chest, neckJoints);
}
} // namespace MultiBodySystemRandomTools
} // namespace mecano
