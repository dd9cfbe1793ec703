/*
 * mecano_b200_model.h -- C view of the C++ host-side model mirror (mecano_b200/csrc/host/multibody.hpp),
 * so that non-C++ hosts (the Python package, tests) can build a MultiBodySystem the way Mecano code does
 * and obtain the level-ordered tables for mecano_b200_create().  A Java host does not need this header:
 * it flattens its own MultiBodySystemReadOnly.
 *
 * Mirrors: RigidBody / RevoluteJoint / PrismaticJoint / SixDoFJoint / SphericalJoint / PlanarJoint / FixedJoint constructors
 * (M/multiBodySystem/*.java), MultiBodySystemBasics.toMultiBodySystemBasics
 * (M/multiBodySystem/interfaces/MultiBodySystemBasics.java:76-142), and the generators of
 * M/tools/MultiBodySystemRandomTools.java.
 *
 * Bodies are referred to by integer ids: 0 is the root body (elevator); the successor of the k-th
 * created joint has id k + 1.  All functions return >= 0 on success and < 0 on error.
 */
#ifndef MECANO_B200_MODEL_H
#define MECANO_B200_MODEL_H

#include <stdint.h>

#include "mecano_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mecano_model mecano_model;

mecano_model *mecano_model_create(const char *root_body_name);
void mecano_model_destroy(mecano_model *m);
const char *mecano_model_last_error(const mecano_model *m);

/* Joints.  transform12 = row-major rotation (9) + translation (3) of frameBeforeJoint in the predecessor's
 * frameAfterJoint, or NULL for identity.  Returns the joint id (>= 0). */
int mecano_model_add_revolute_joint(mecano_model *m, const char *name, int predecessor_body, const double *transform12, const double *axis3);
int mecano_model_add_prismatic_joint(mecano_model *m, const char *name, int predecessor_body, const double *transform12, const double *axis3);
int mecano_model_add_sixdof_joint(mecano_model *m, const char *name, int predecessor_body, const double *transform12);
/* SphericalJoint (M/multiBodySystem/SphericalJoint.java:43-69) and PlanarJoint (PlanarJoint.java:37-61): three DoFs each */
int mecano_model_add_spherical_joint(mecano_model *m, const char *name, int predecessor_body, const double *transform12);
int mecano_model_add_planar_joint(mecano_model *m, const char *name, int predecessor_body, const double *transform12);
/* FixedJoint (M/multiBodySystem/FixedJoint.java:40-62): 0 DoF.  Host-only: at finalize its successor is welded into the nearest
 * moving ancestor (inertia, children and offsets), so the GPU tables contain moving joints only. */
#define MECANO_MODEL_FIXED 5
int mecano_model_add_fixed_joint(mecano_model *m, const char *name, int predecessor_body, const double *transform12);
/* Configuration a joint keeps when it is ignored (one-DoF: q; SixDoF: qx qy qz qs x y z; Spherical: qx qy qz qs; Planar: pitch x z);
 * default zero / identity. */
int mecano_model_set_joint_configuration(mecano_model *m, int joint, const double *q, int n);
/* RigidBody(name, parentJoint, momentOfInertia, mass, inertiaPose).  Returns the body id (= joint id + 1). */
int mecano_model_add_rigid_body(mecano_model *m, const char *name, int parent_joint, const double *inertia9, double mass, const double *inertia_pose12);

/* MultiBodySystemRandomTools generators (seeded splitmix64; Mecano's distributions).  Each appends to the model. */
int mecano_model_next_one_dof_joint_chain(mecano_model *m, uint64_t seed, int predecessor_body, int n_joints, double prismatic_fraction);
int mecano_model_next_one_dof_joint_tree(mecano_model *m, uint64_t seed, int predecessor_body, int n_joints, double prismatic_fraction);
int mecano_model_next_floating_base(mecano_model *m, uint64_t seed, int predecessor_body); /* returns the new body id */
/* nextJointChain / nextJointTree (MultiBodySystemRandomTools.java:424-440, :844-860): joints of random types, all five moving types */
int mecano_model_next_joint_chain(mecano_model *m, uint64_t seed, int predecessor_body, int n_joints);
int mecano_model_next_joint_tree(mecano_model *m, uint64_t seed, int predecessor_body, int n_joints);
int mecano_model_next_humanoid(mecano_model *m, uint64_t seed, int neck_joints);

/* toMultiBodySystemBasics(rootBody): fixes the joint order / index provider and builds the tables. */
int mecano_model_finalize(mecano_model *m);
/* MultiBodySystemBasics.toMultiBodySystemBasics(rootBody, jointsToIgnore) (MultiBodySystemBasics.java:90-142): the listed joints
 * and their descendants are left out of the index provider; the inertia of each ignored subtree is lumped into its parent body
 * at the joints' stored configuration (InverseDynamicsCalculator.java:236, :832-860; MultiBodySystemTools.java:32-64). */
int mecano_model_finalize_ignoring(mecano_model *m, const int32_t *joints_to_ignore, int n_ignore);
int mecano_model_n_joints(const mecano_model *m);
int mecano_model_n_dofs(const mecano_model *m);
int mecano_model_n_cfg(const mecano_model *m);
/* joint ids in JointMatrixIndexProvider order, and their first DoF / configuration rows */
int mecano_model_joint_order(const mecano_model *m, int32_t *joint_ids, int32_t *dof_index, int32_t *cfg_index);
/* per joint (by id): type, predecessor body id, axis, transform, and its successor's inertia parameters */
int mecano_model_joint_info(const mecano_model *m, int joint, int32_t *type, int32_t *predecessor_body, double *axis3, double *transform12,
                            double *inertia9, double *mass, double *inertia_pose12);
const char *mecano_model_joint_name(const mecano_model *m, int joint);
const char *mecano_model_body_name(const mecano_model *m, int body);
/* level-ordered tables (valid until the model is destroyed) */
const mecano_b200_tree_desc *mecano_model_tables(const mecano_model *m);
/* row of joint `joint` in the level-ordered tables */
int mecano_model_table_row(const mecano_model *m, int joint);

/* The same system with nothing welded, for external wrenches and per-body results on systems with fixed / ignored joints
 * (InverseDynamicsCalculator.java:469-472 with :832-860): every joint of the tree is a body of these tables; a FixedJoint is a
 * revolute joint held at q = 0, an ignored joint keeps its type at its stored configuration.  The held joints' configuration /
 * DoF rows come BEHIND the system's own (n_extra_cfg / n_extra_dof more rows), the wrench blocks of ignored bodies behind the
 * considered joints' (n_extra_wrench_blocks more blocks of six rows): the caller's matrices are the leading rows of the expanded
 * ones.  expanded_fill: q_extra[n_extra_cfg] = what to put into the extra configuration rows (velocities and accelerations of
 * held joints are zero); locked[n_bodies] = 1 for held joints (forward dynamics: ACCELERATION_SOURCE with zero acceleration,
 * mecano_b200_set_joint_source_modes); row_of_considered[considered joints] = table row of each considered joint. */
const mecano_b200_tree_desc *mecano_model_expanded_tables(mecano_model *m);
int mecano_model_expanded_info(mecano_model *m, int32_t *n_bodies, int32_t *n_extra_dof, int32_t *n_extra_cfg, int32_t *n_extra_wrench_blocks);
int mecano_model_expanded_fill(mecano_model *m, double *q_extra, int32_t *locked, int32_t *row_of_considered);

#ifdef __cplusplus
}
#endif
#endif
