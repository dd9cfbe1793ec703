// model_capi.cpp -- C view of multibody.hpp (include/mecano_b200_model.h).  Host only.
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/mecano_b200_model.h"
#include "multibody.hpp"

using namespace mecano;

struct mecano_model
{
   MultiBodyArena arena;
   std::vector<RigidBody *> bodies; // id -> body (0 = root)
   std::vector<Joint *> joints;     // id -> joint
   bool finalized = false;
   MultiBodySystem system;
   FlatTables tables;
   FlatTables expanded; // nothing welded, fixed / ignored joints held (lazy)
   FlatTables::Expanded expanded_info;
   bool has_expanded = false;
   std::string error;
};

namespace
{
RigidBodyTransform to_transform(const double *t12)
{
   RigidBodyTransform T;
   if (t12)
   {
      std::memcpy(T.rotation.m, t12, 9 * sizeof(double));
      T.translation = Vector3D{t12[9], t12[10], t12[11]};
   }
   return T;
}

int fail(mecano_model *m, const std::string &msg)
{
   if (m) m->error = msg;
   return -1;
}

// joints / bodies created by the generators are registered by walking what was appended under `pred`
void register_subtree(mecano_model *m, RigidBody *body)
{
   for (Joint *j : body->getChildrenJoints())
   {
      bool known = false;
      for (Joint *k : m->joints)
         if (k == j) { known = true; break; }
      if (!known)
      {
         m->joints.push_back(j);
         m->bodies.push_back(j->getSuccessor());
      }
      if (j->getSuccessor()) register_subtree(m, j->getSuccessor());
   }
}

template <class F> int guarded(mecano_model *m, F f)
{
   if (!m) return -1;
   if (m->finalized) return fail(m, "model is finalized");
   try
   {
      return f();
   }
   catch (const std::exception &e)
   {
      return fail(m, e.what());
   }
}
} // namespace

extern "C" {
#pragma GCC visibility push(default)

mecano_model *mecano_model_create(const char *root_body_name)
{
   mecano_model *m = new mecano_model();
   m->bodies.push_back(m->arena.newRootBody(root_body_name ? root_body_name : "elevator"));
   return m;
}

void mecano_model_destroy(mecano_model *m) { delete m; }

const char *mecano_model_last_error(const mecano_model *m) { return m ? m->error.c_str() : "model is NULL"; }

static int add_joint(mecano_model *m, int type, const char *name, int pred, const double *t12, const double *axis3)
{
   return guarded(m, [&]() -> int {
      if (pred < 0 || pred >= (int)m->bodies.size() || !m->bodies[pred]) return fail(m, "predecessor body id out of range");
      const RigidBodyTransform T = to_transform(t12);
      const std::string n = name ? name : "joint" + std::to_string(m->joints.size());
      Joint *j = nullptr;
      if (type == MECANO_MODEL_FIXED)
         j = m->arena.newJoint<FixedJoint>(n, m->bodies[pred], T);
      else if (type == MECANO_B200_SIXDOF)
         j = m->arena.newJoint<SixDoFJoint>(n, m->bodies[pred], T);
      else if (type == MECANO_B200_SPHERICAL)
         j = m->arena.newJoint<SphericalJoint>(n, m->bodies[pred], T);
      else if (type == MECANO_B200_PLANAR)
         j = m->arena.newJoint<PlanarJoint>(n, m->bodies[pred], T);
      else
      {
         if (!axis3) return fail(m, "axis is NULL");
         const Vector3D a{axis3[0], axis3[1], axis3[2]};
         if (type == MECANO_B200_REVOLUTE)
            j = m->arena.newJoint<RevoluteJoint>(n, m->bodies[pred], T, a);
         else
            j = m->arena.newJoint<PrismaticJoint>(n, m->bodies[pred], T, a);
      }
      m->joints.push_back(j);
      m->bodies.push_back(nullptr); // successor not created yet
      return (int)m->joints.size() - 1;
   });
}

int mecano_model_add_revolute_joint(mecano_model *m, const char *name, int pred, const double *t12, const double *axis3)
{
   return add_joint(m, MECANO_B200_REVOLUTE, name, pred, t12, axis3);
}
int mecano_model_add_prismatic_joint(mecano_model *m, const char *name, int pred, const double *t12, const double *axis3)
{
   return add_joint(m, MECANO_B200_PRISMATIC, name, pred, t12, axis3);
}
int mecano_model_add_sixdof_joint(mecano_model *m, const char *name, int pred, const double *t12)
{
   return add_joint(m, MECANO_B200_SIXDOF, name, pred, t12, nullptr);
}

int mecano_model_add_spherical_joint(mecano_model *m, const char *name, int pred, const double *t12)
{
   return add_joint(m, MECANO_B200_SPHERICAL, name, pred, t12, nullptr);
}
int mecano_model_add_planar_joint(mecano_model *m, const char *name, int pred, const double *t12)
{
   return add_joint(m, MECANO_B200_PLANAR, name, pred, t12, nullptr);
}

int mecano_model_add_fixed_joint(mecano_model *m, const char *name, int pred, const double *t12)
{
   return add_joint(m, MECANO_MODEL_FIXED, name, pred, t12, nullptr);
}

int mecano_model_set_joint_configuration(mecano_model *m, int joint, const double *q, int n)
{
   return guarded(m, [&]() -> int {
      if (joint < 0 || joint >= (int)m->joints.size()) return fail(m, "joint id out of range");
      if (!q || n != m->joints[joint]->getConfigurationMatrixSize()) return fail(m, "configuration size does not match the joint");
      m->joints[joint]->setJointConfiguration(q, n);
      return 0;
   });
}

int mecano_model_add_rigid_body(mecano_model *m, const char *name, int joint, const double *inertia9, double mass, const double *pose12)
{
   return guarded(m, [&]() -> int {
      if (joint < 0 || joint >= (int)m->joints.size()) return fail(m, "parent joint id out of range");
      if (m->bodies[joint + 1]) return fail(m, "joint already has a successor");
      if (!inertia9) return fail(m, "inertia is NULL");
      Matrix3D I;
      std::memcpy(I.m, inertia9, 9 * sizeof(double));
      const std::string n = name ? name : "body" + std::to_string(joint);
      m->bodies[joint + 1] = m->arena.newRigidBody(n, m->joints[joint], I, mass, to_transform(pose12));
      return joint + 1;
   });
}

int mecano_model_next_one_dof_joint_chain(mecano_model *m, uint64_t seed, int pred, int n, double prismatic_fraction)
{
   return guarded(m, [&]() -> int {
      if (pred < 0 || pred >= (int)m->bodies.size() || !m->bodies[pred]) return fail(m, "predecessor body id out of range");
      Random r(seed);
      MultiBodySystemRandomTools::nextOneDoFJointChain(r, m->arena, "chain" + std::to_string(m->joints.size()), m->bodies[pred], n, prismatic_fraction);
      register_subtree(m, m->bodies[0]);
      return (int)m->bodies.size() - 1;
   });
}

int mecano_model_next_one_dof_joint_tree(mecano_model *m, uint64_t seed, int pred, int n, double prismatic_fraction)
{
   return guarded(m, [&]() -> int {
      if (pred < 0 || pred >= (int)m->bodies.size() || !m->bodies[pred]) return fail(m, "predecessor body id out of range");
      Random r(seed);
      MultiBodySystemRandomTools::nextOneDoFJointTree(r, m->arena, "tree" + std::to_string(m->joints.size()), m->bodies[pred], n, prismatic_fraction);
      register_subtree(m, m->bodies[0]);
      return (int)m->bodies.size() - 1;
   });
}

int mecano_model_next_joint_chain(mecano_model *m, uint64_t seed, int pred, int n)
{
   return guarded(m, [&]() -> int {
      if (pred < 0 || pred >= (int)m->bodies.size() || !m->bodies[pred]) return fail(m, "predecessor body id out of range");
      Random r(seed);
      MultiBodySystemRandomTools::nextJointChain(r, m->arena, "chain" + std::to_string(m->joints.size()), m->bodies[pred], n);
      register_subtree(m, m->bodies[0]);
      return (int)m->bodies.size() - 1;
   });
}

int mecano_model_next_joint_tree(mecano_model *m, uint64_t seed, int pred, int n)
{
   return guarded(m, [&]() -> int {
      if (pred < 0 || pred >= (int)m->bodies.size() || !m->bodies[pred]) return fail(m, "predecessor body id out of range");
      Random r(seed);
      MultiBodySystemRandomTools::nextJointTree(r, m->arena, "tree" + std::to_string(m->joints.size()), m->bodies[pred], n);
      register_subtree(m, m->bodies[0]);
      return (int)m->bodies.size() - 1;
   });
}

int mecano_model_next_floating_base(mecano_model *m, uint64_t seed, int pred)
{
   return guarded(m, [&]() -> int {
      if (pred < 0 || pred >= (int)m->bodies.size() || !m->bodies[pred]) return fail(m, "predecessor body id out of range");
      Random r(seed);
      MultiBodySystemRandomTools::nextFloatingBase(r, m->arena, m->bodies[pred], "floating" + std::to_string(m->joints.size()));
      register_subtree(m, m->bodies[0]);
      return (int)m->bodies.size() - 1;
   });
}

int mecano_model_next_humanoid(mecano_model *m, uint64_t seed, int neck_joints)
{
   return guarded(m, [&]() -> int {
      Random r(seed);
      MultiBodySystemRandomTools::nextHumanoid(r, m->arena, m->bodies[0], neck_joints);
      register_subtree(m, m->bodies[0]);
      return (int)m->bodies.size() - 1;
   });
}

int mecano_model_finalize(mecano_model *m) { return mecano_model_finalize_ignoring(m, nullptr, 0); }

int mecano_model_finalize_ignoring(mecano_model *m, const int32_t *joints_to_ignore, int n_ignore)
{
   if (!m) return -1;
   if (m->finalized) return 0;
   try
   {
      for (size_t j = 0; j < m->joints.size(); j++)
         if (!m->bodies[j + 1]) return fail(m, "joint " + m->joints[j]->getName() + " has no successor");
      std::vector<Joint *> ignore;
      for (int k = 0; k < n_ignore; k++)
      {
         if (!joints_to_ignore || joints_to_ignore[k] < 0 || joints_to_ignore[k] >= (int)m->joints.size()) return fail(m, "joint id to ignore out of range");
         ignore.push_back(m->joints[joints_to_ignore[k]]);
      }
      m->system = MultiBodySystem::toMultiBodySystemBasics(m->bodies[0], ignore);
      bool moving = false;
      for (const Joint *j : m->system.getJointsToConsider())
         moving = moving || j->getDegreesOfFreedom() > 0;
      if (!moving) return fail(m, "the system has no moving joints");
      m->tables = FlatTables::flatten(m->system);
      m->tables.bind(m->system.getNumberOfDoFs(), m->system.getConfigurationMatrixSize());
      m->finalized = true;
      return 0;
   }
   catch (const std::exception &e)
   {
      return fail(m, e.what());
   }
}

int mecano_model_n_joints(const mecano_model *m) { return m ? (int)m->joints.size() : -1; }
int mecano_model_n_dofs(const mecano_model *m) { return m && m->finalized ? m->system.getNumberOfDoFs() : -1; }
int mecano_model_n_cfg(const mecano_model *m) { return m && m->finalized ? m->system.getConfigurationMatrixSize() : -1; }

int mecano_model_joint_order(const mecano_model *m, int32_t *joint_ids, int32_t *dof_index, int32_t *cfg_index)
{
   if (!m || !m->finalized) return -1;
   const auto &js = m->system.getJointsToConsider();
   for (size_t i = 0; i < js.size(); i++)
   {
      int id = -1;
      for (size_t k = 0; k < m->joints.size(); k++)
         if (m->joints[k] == js[i]) { id = (int)k; break; }
      if (joint_ids) joint_ids[i] = id;
      if (dof_index) dof_index[i] = m->system.dofIndexAt(i);
      if (cfg_index) cfg_index[i] = m->system.cfgIndexAt(i);
   }
   return (int)js.size();
}

int mecano_model_joint_info(const mecano_model *m, int joint, int32_t *type, int32_t *pred_body, double *axis3, double *t12, double *inertia9,
                            double *mass, double *pose12)
{
   if (!m || joint < 0 || joint >= (int)m->joints.size()) return -1;
   const Joint *j = m->joints[joint];
   if (type) *type = (int32_t)j->getType();
   if (pred_body)
   {
      *pred_body = -1;
      for (size_t b = 0; b < m->bodies.size(); b++)
         if (m->bodies[b] == j->getPredecessor()) { *pred_body = (int32_t)b; break; }
   }
   if (axis3) { axis3[0] = j->getJointAxis().x; axis3[1] = j->getJointAxis().y; axis3[2] = j->getJointAxis().z; }
   if (t12)
   {
      std::memcpy(t12, j->getTransformToParent().rotation.m, 9 * sizeof(double));
      t12[9] = j->getTransformToParent().translation.x; t12[10] = j->getTransformToParent().translation.y; t12[11] = j->getTransformToParent().translation.z;
   }
   const RigidBody *b = j->getSuccessor();
   if (b)
   {
      if (inertia9) std::memcpy(inertia9, b->getMomentOfInertia().m, 9 * sizeof(double));
      if (mass) *mass = b->getMass();
      if (pose12)
      {
         std::memcpy(pose12, b->getInertiaPose().rotation.m, 9 * sizeof(double));
         pose12[9] = b->getInertiaPose().translation.x; pose12[10] = b->getInertiaPose().translation.y; pose12[11] = b->getInertiaPose().translation.z;
      }
   }
   return 0;
}

const char *mecano_model_joint_name(const mecano_model *m, int joint)
{
   if (!m || joint < 0 || joint >= (int)m->joints.size()) return "";
   return m->joints[joint]->getName().c_str();
}

const char *mecano_model_body_name(const mecano_model *m, int body)
{
   if (!m || body < 0 || body >= (int)m->bodies.size() || !m->bodies[body]) return "";
   return m->bodies[body]->getName().c_str();
}

const mecano_b200_tree_desc *mecano_model_tables(const mecano_model *m) { return m && m->finalized ? &m->tables.desc : nullptr; }

// ---- the expanded tables (FlatTables::flattenExpanded): nothing welded, the fixed / ignored joints held
static bool ensure_expanded(mecano_model *m)
{
   if (!m || !m->finalized) return false;
   if (!m->has_expanded)
   {
      m->expanded = FlatTables::flattenExpanded(m->system, m->expanded_info);
      m->expanded.bind(m->system.getNumberOfDoFs() + m->expanded_info.n_extra_dof, m->system.getConfigurationMatrixSize() + m->expanded_info.n_extra_cfg);
      m->has_expanded = true;
   }
   return true;
}

const mecano_b200_tree_desc *mecano_model_expanded_tables(mecano_model *m)
{
   try
   {
      return ensure_expanded(m) ? &m->expanded.desc : nullptr;
   }
   catch (const std::exception &e)
   {
      fail(m, e.what());
      return nullptr;
   }
}

int mecano_model_expanded_info(mecano_model *m, int32_t *n_bodies, int32_t *n_extra_dof, int32_t *n_extra_cfg, int32_t *n_extra_wrench_blocks)
{
   try
   {
      if (!ensure_expanded(m)) return -1;
   }
   catch (const std::exception &e)
   {
      return fail(m, e.what());
   }
   if (n_bodies) *n_bodies = (int32_t)m->expanded.parent.size();
   if (n_extra_dof) *n_extra_dof = m->expanded_info.n_extra_dof;
   if (n_extra_cfg) *n_extra_cfg = m->expanded_info.n_extra_cfg;
   if (n_extra_wrench_blocks) *n_extra_wrench_blocks = m->expanded_info.n_extra_bodies;
   return 0;
}

int mecano_model_expanded_fill(mecano_model *m, double *q_extra, int32_t *locked, int32_t *row_of_considered)
{
   try
   {
      if (!ensure_expanded(m)) return -1;
   }
   catch (const std::exception &e)
   {
      return fail(m, e.what());
   }
   const auto &x = m->expanded_info;
   if (q_extra) std::copy(x.q_extra.begin(), x.q_extra.end(), q_extra);
   if (locked) std::copy(x.locked.begin(), x.locked.end(), locked);
   if (row_of_considered) std::copy(x.row_of_considered.begin(), x.row_of_considered.end(), row_of_considered);
   return 0;
}

int mecano_model_table_row(const mecano_model *m, int joint)
{
   if (!m || !m->finalized || joint < 0 || joint >= (int)m->joints.size()) return -1;
   const int i = m->system.indexOf(m->joints[joint]);
   return i < 0 ? -1 : m->tables.body_of_joint[(size_t)i];
}

#pragma GCC visibility pop
} // extern "C"
