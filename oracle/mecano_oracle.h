/*
 * mecano_oracle.h -- CPU restatement of Mecano's RNEA / ABA / CRBA calculators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (ihmcrobotics/mecano, pure Java) cannot be run in
 * this environment (no JDK) and its tests hold no golden vectors for this path
 * (SURVEY.md section 8c).  This oracle is pinned instead against
 *   (i)  the reference's own randomized invariants at the reference's tolerances
 *        (ForwardDynamicsCalculatorTest.java:38-41, 767-1003), and
 *   (ii) an independently written textbook (Featherstone, dense 6x6) formulation
 *        in tests/featherstone_np.py.
 *
 * Conventions (all from the reference):
 *   - spatial vectors are angular-first [wx wy wz | vx vy vz]
 *     (spatial/interfaces/SpatialVectorReadOnly.java:268-272)
 *   - a transform (R,t) attached to a child frame maps child coords to parent coords
 *   - SixDoF configuration is [qx qy qz qs | x y z], velocity-like vectors are the
 *     body-frame twist (multiBodySystem/interfaces/SixDoFJointReadOnly.java:21-26)
 *   - bodies are listed in Mecano's depth-first pre-order, children in insertion
 *     order (multiBodySystem/iterators/JointIterator.java:153-162)
 */
#ifndef MECANO_ORACLE_H
#define MECANO_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* SphericalJoint: configuration [qx qy qz qs], velocity-like rows = angular part in frameAfterJoint
 * (multiBodySystem/interfaces/SphericalJointReadOnly.java:31-71).  PlanarJoint: configuration [pitch x z], velocity-like rows
 * [w_y v_x v_z] in frameAfterJoint (PlanarJointReadOnly.java:20-58, tools/MecanoTools.java:920-952). */
enum { MO_REVOLUTE = 0, MO_PRISMATIC = 1, MO_SIXDOF = 2, MO_SPHERICAL = 3, MO_PLANAR = 4 };

/* flags for mo_rnea (InverseDynamicsCalculator.java:291-306) */
enum { MO_NO_CORIOLIS = 1, MO_NO_ACCELERATIONS = 2 };

typedef struct mo_tree
{
   int nb;              /* number of non-root bodies (= number of joints) */
   int nv;              /* degrees of freedom */
   int nq;              /* configuration entries */
   const int *parent;   /* [nb] parent body, -1 = root body (elevator) */
   const int *jtype;    /* [nb] MO_REVOLUTE / MO_PRISMATIC / MO_SIXDOF / MO_SPHERICAL / MO_PLANAR */
   const double *axis;  /* [nb][3] unit joint axis (ignored for SixDoF) */
   const double *off_R; /* [nb][9] row-major rotation of frameBeforeJoint in parent's frameAfterJoint */
   const double *off_p; /* [nb][3] translation of the same */
   const double *com_R; /* [nb][9] inertia pose: bodyFixedFrame (CoM frame) in frameAfterJoint */
   const double *com_p; /* [nb][3] */
   const double *J;     /* [nb][9] moment of inertia about the CoM, in bodyFixedFrame */
   const double *mass;  /* [nb] */
   const int *dof_off;  /* [nb] first row of this joint in the nv-vectors */
   const int *cfg_off;  /* [nb] first row of this joint in the nq-vector */
} mo_tree;

/* Single-state calculators.  fext may be NULL; otherwise [nb][6], the external wrench on each body
 * expressed in that body's CoM frame (InverseDynamicsCalculator.java:819, :946). */
void mo_rnea(const mo_tree *t, const double *gravity3, const double *q, const double *qd, const double *qdd, const double *fext,
             int flags, double *tau);
void mo_aba(const mo_tree *t, const double *gravity3, const double *q, const double *qd, const double *tau, const double *fext,
            double *qdd);
void mo_crba(const mo_tree *t, const double *q, double *M /* [nv][nv] row-major */);

/* By-products used by the tests to mirror ForwardDynamicsCalculatorTest: per-body spatial accelerations
 * (expressed in each body's CoM frame) from RNEA pass one, [nb][6]. */
void mo_rnea_body_accelerations(const mo_tree *t, const double *gravity3, const double *q, const double *qd, const double *qdd,
                                int flags, double *acc);

/* RNEA with its by-products: body_acc [nb][6] = getBodyAcceleration(body), the spatial acceleration of each body expressed in its
 * CoM frame (InverseDynamicsCalculator.java:578-591, pass one :880-910); joint_wrench [nb][6] = getComputedJointWrench(joint), the
 * wrench transmitted by each joint expressed in its frameAfterJoint (:593-602, pass two :943-950).  Either may be NULL. */
void mo_rnea_full(const mo_tree *t, const double *gravity3, const double *q, const double *qd, const double *qdd, const double *fext,
                  int flags, double *tau, double *body_acc, double *joint_wrench);

/* CRBA by-products (CompositeRigidBodyMassMatrixCalculator.java:380-440, :801-839).  frame: the centroidal momentum frame.
 * M [nv][nv]; cmm (nullable) [6][nv] row-major, angular rows first: getCentroidalMomentumMatrix(); com4 (nullable): centre of mass of
 * the system in the root frame and its total mass. */
#define MO_FRAME_WORLD 0 /* the inertial (root) frame */
#define MO_FRAME_COM 1   /* axes of the inertial frame, origin at the centre of mass (CenterOfMassReferenceFrame) */
void mo_crba_centroidal(const mo_tree *t, const double *q, int frame, double *M, double *cmm, double *com4);
/* getCentroidalConvectiveTerm(): d/dt(cmm) qd, [6] angular first, in the centroidal frame */
void mo_centroidal_convective_term(const mo_tree *t, const double *q, const double *qd, int frame, double *out6);

/* Coriolis and centrifugal matrix C(q, qd) [nv][nv] row-major (getCoriolisMatrix(), CompositeRigidBodyMassMatrixCalculator.java:358-366
 * after setEnableCoriolisMatrixCalculation(true)); C qd = the joint efforts of inverse dynamics with zero joint accelerations and no
 * gravity.  M (nullable): the mass matrix computed by the same recursion. */
void mo_coriolis(const mo_tree *t, const double *q, const double *qd, double *M, double *C);

/* ABA with per-joint source modes (ForwardDynamicsCalculator.java:45-57, :400-444, :508-520): accel_source [nb], non-zero = the
 * joint's acceleration is an input (qdd_in [nv], its rows only) and its effort an output (pass four :1315-1363).  qdd [nv] holds all
 * joint accelerations, tau_out (nullable, [nv]) all joint efforts as getJointTauMatrix() :566-590 returns them. */
void mo_aba_sources(const mo_tree *t, const double *gravity3, const double *q, const double *qd, const double *tau, const double *qdd_in,
                    const double *fext, const int *accel_source, double *qdd, double *tau_out);

/* Batched drivers, DoF-major / state-minor buffers x[k*ld + s] (same layout as the C-ABI).
 * fext (nullable) is [(6*nb)][ld].  M is entry-major [(i*nv+j)*ld + s].  nthreads<=0 -> all cores. */
void mo_rnea_batch(const mo_tree *t, const double *gravity3, long n, long ld, const double *q, const double *qd, const double *qdd,
                   const double *fext, int flags, double *tau, int nthreads);
void mo_aba_batch(const mo_tree *t, const double *gravity3, long n, long ld, const double *q, const double *qd, const double *tau,
                  const double *fext, double *qdd, int nthreads);
void mo_crba_batch(const mo_tree *t, long n, long ld, const double *q, double *M, int nthreads);

/* MultiBodySystemStateIntegrator.doubleIntegrateFromAcceleration (tools/MultiBodySystemStateIntegrator.java:365-470, :503-560,
 * :710-733): q, qd and the SixDoF rows of qdd are updated in place.  Batched form: same buffer layout as above. */
void mo_integrate(const mo_tree *t, double dt, double *q, double *qd, double *qdd);
void mo_integrate_batch(const mo_tree *t, double dt, long n, long ld, double *q, double *qd, double *qdd);

int mo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
