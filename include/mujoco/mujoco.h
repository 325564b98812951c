/* mujoco.h — source-compatibility shim for the subset of the MuJoCo 2.3.7 C API
 * that HoangGiang93/mujoco_sim touches on its per-tick hot path.
 *
 * This is NOT MuJoCo.  It declares, with MuJoCo's names, the types, fields and
 * functions the reference's L2/L3 sources use, so that
 * `#include <mujoco/mujoco.h>` (reference include/mujoco_sim/mj_model.h:23)
 * resolves to the B200 batched engine in libb2sim.so.  Struct layout is our own:
 * compatibility is at source level (field names / meanings / units), not ABI level.
 *
 * Call sites replaced (reference file:line):
 *   mj_step1      src/mj_main.cpp:83            mj_step2     src/mj_main.cpp:108
 *   mjcb_control  src/mj_main.cpp:196           mj_inverse   src/mujoco_sim/mj_hw_interface.cpp:61
 *   mj_mulM       src/mujoco_sim/mj_sim.cpp:1057
 *   mj_forward    src/mujoco_sim/mj_ros.cpp:608,1421
 *   mj_makeData   src/mujoco_sim/mj_sim.cpp:816,835 ; src/mujoco_sim/mj_ros.cpp:571
 *   mj_loadXML    include/mujoco_sim/mj_util.h:190
 *   mj_name2id / mj_id2name   (34 + 23 sites, e.g. mj_hw_interface.cpp:64,79; mj_sim.cpp:473)
 *   mj_deleteData / mj_deleteModel   src/mj_main.cpp:232-233
 */
#ifndef B2_MUJOCO_SHIM_H_
#define B2_MUJOCO_SHIM_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef double mjtNum;
typedef unsigned char mjtByte;

#define mjMINVAL 1E-15
#define mjPI 3.14159265358979323846
#define mjMAXVAL 1E+10
#define mjMINMU 1E-5
#define mjMINIMP 0.0001
#define mjMAXIMP 0.9999
#define mjNEQDATA 11
#define mjNIMP 5
#define mjNREF 2
#define mjMAXCONPAIR 8 /* upper bound on contacts one geom pair can emit */

typedef enum mjtObj_ {
  mjOBJ_UNKNOWN = 0, mjOBJ_BODY, mjOBJ_XBODY, mjOBJ_JOINT, mjOBJ_DOF, mjOBJ_GEOM, mjOBJ_SITE,
  mjOBJ_CAMERA, mjOBJ_LIGHT, mjOBJ_MESH, mjOBJ_SKIN, mjOBJ_HFIELD, mjOBJ_TEXTURE, mjOBJ_MATERIAL,
  mjOBJ_PAIR, mjOBJ_EXCLUDE, mjOBJ_EQUALITY, mjOBJ_TENDON, mjOBJ_ACTUATOR, mjOBJ_SENSOR
} mjtObj;

typedef enum mjtJoint_ { mjJNT_FREE = 0, mjJNT_BALL, mjJNT_SLIDE, mjJNT_HINGE } mjtJoint;

typedef enum mjtGeom_ {
  mjGEOM_PLANE = 0, mjGEOM_HFIELD, mjGEOM_SPHERE, mjGEOM_CAPSULE, mjGEOM_ELLIPSOID,
  mjGEOM_CYLINDER, mjGEOM_BOX, mjGEOM_MESH, mjNGEOMTYPES
} mjtGeom;

typedef enum mjtEq_ { mjEQ_CONNECT = 0, mjEQ_WELD, mjEQ_JOINT, mjEQ_TENDON, mjEQ_DISTANCE } mjtEq;

typedef enum mjtSensor_ { mjSENS_TOUCH = 0, mjSENS_ACCELEROMETER, mjSENS_VELOCIMETER, mjSENS_GYRO,
                          mjSENS_FORCE, mjSENS_TORQUE } mjtSensor;

typedef enum mjtConstraint_ {
  mjCNSTR_EQUALITY = 0, mjCNSTR_FRICTION_DOF, mjCNSTR_FRICTION_TENDON, mjCNSTR_LIMIT_JOINT,
  mjCNSTR_LIMIT_TENDON, mjCNSTR_CONTACT_FRICTIONLESS, mjCNSTR_CONTACT_PYRAMIDAL, mjCNSTR_CONTACT_ELLIPTIC
} mjtConstraint;

typedef enum mjtIntegrator_ { mjINT_EULER = 0, mjINT_RK4, mjINT_IMPLICIT } mjtIntegrator;
typedef enum mjtSolver_ { mjSOL_PGS = 0, mjSOL_CG, mjSOL_NEWTON } mjtSolver;
typedef enum mjtCone_ { mjCONE_PYRAMIDAL = 0, mjCONE_ELLIPTIC } mjtCone;

typedef enum mjtDisableBit_ {
  mjDSBL_CONSTRAINT = 1 << 0, mjDSBL_EQUALITY = 1 << 1, mjDSBL_FRICTIONLOSS = 1 << 2, mjDSBL_LIMIT = 1 << 3,
  mjDSBL_CONTACT = 1 << 4, mjDSBL_PASSIVE = 1 << 5, mjDSBL_GRAVITY = 1 << 6, mjDSBL_CLAMPCTRL = 1 << 7,
  mjDSBL_WARMSTART = 1 << 8, mjDSBL_FILTERPARENT = 1 << 9, mjDSBL_ACTUATION = 1 << 10, mjDSBL_REFSAFE = 1 << 11,
  mjDSBL_SENSOR = 1 << 12, mjDSBL_MIDPHASE = 1 << 13, mjDSBL_EULERDAMP = 1 << 14
} mjtDisableBit;
typedef enum mjtEnableBit_ { mjENBL_OVERRIDE = 1 << 0, mjENBL_ENERGY = 1 << 1 } mjtEnableBit;

typedef struct mjOption_ {
  mjtNum timestep;       /* read AND written by the reference every tick (src/mj_main.cpp:150-163) */
  mjtNum impratio;
  mjtNum tolerance;
  mjtNum noslip_tolerance;
  mjtNum gravity[3];
  int integrator;        /* RK4 in the shipped worlds; the split step API only knows Euler */
  int cone;
  int solver;            /* this engine always runs PGS (north_star); field kept for printing */
  int iterations;
  int noslip_iterations;
  int disableflags;
  int enableflags;
} mjOption;

typedef struct mjStatistic_ {
  mjtNum meaninertia;
  mjtNum meanmass;
  mjtNum extent;
} mjStatistic;

typedef struct mjContact_ {
  mjtNum dist;           /* distance between nearest points; <0: penetration */
  mjtNum pos[3];         /* midpoint between the two surfaces */
  mjtNum frame[9];       /* rows: normal (geom1 -> geom2), tangent1, tangent2 */
  mjtNum includemargin;  /* margin - gap */
  mjtNum friction[5];    /* tangent1, tangent2, spin, roll1, roll2 */
  mjtNum solref[mjNREF];
  mjtNum solimp[mjNIMP];
  mjtNum mu;
  int dim;               /* condim: 1, 3, 4 or 6 */
  int geom1, geom2;      /* bit-exact parity fields */
  int exclude;
  int efc_address;       /* first constraint row of this contact, -1 if none */
  int pair;              /* index into the static candidate pair list (extension) */
} mjContact;

typedef struct mjVFS_ mjVFS; /* opaque; callers pass NULL (include/mujoco_sim/mj_util.h:190) */

typedef struct mjModel_ {
  /* sizes */
  int nq, nv, nu, na, nbody, njnt, ngeom, nmesh, nmeshvert, neq, nexclude, nM, nmocap;
  int nsensor, nsensordata, nnames;
  int npair;             /* extension: static candidate geom pairs after all compile-time filters */
  int nconmax, njmax;    /* per-environment caps for contacts and constraint rows */

  mjOption opt;
  mjStatistic stat;

  mjtNum* qpos0;         /* nq */
  mjtNum* qpos_spring;   /* nq */

  /* bodies */
  int* body_parentid; int* body_rootid; int* body_weldid; int* body_mocapid;
  int* body_jntnum; int* body_jntadr; int* body_dofnum; int* body_dofadr;
  int* body_geomnum; int* body_geomadr;
  mjtNum* body_pos;      /* nbody x 3 */
  mjtNum* body_quat;     /* nbody x 4 */
  mjtNum* body_ipos;     /* nbody x 3 */
  mjtNum* body_iquat;    /* nbody x 4 */
  mjtNum* body_mass;     /* nbody */
  mjtNum* body_subtreemass;
  mjtNum* body_inertia;  /* nbody x 3 */
  mjtNum* body_invweight0; /* nbody x 2 */
  mjtNum* body_gravcomp; /* nbody */

  /* joints */
  int* jnt_type; int* jnt_qposadr; int* jnt_dofadr; int* jnt_bodyid;
  mjtByte* jnt_limited;
  mjtNum* jnt_solref;    /* njnt x 2 */
  mjtNum* jnt_solimp;    /* njnt x 5 */
  mjtNum* jnt_pos;       /* njnt x 3 */
  mjtNum* jnt_axis;      /* njnt x 3 */
  mjtNum* jnt_stiffness; /* njnt */
  mjtNum* jnt_range;     /* njnt x 2 */
  mjtNum* jnt_margin;    /* njnt */

  /* dofs */
  int* dof_bodyid; int* dof_jntid; int* dof_parentid; int* dof_Madr;
  mjtNum* dof_solref;    /* nv x 2 */
  mjtNum* dof_solimp;    /* nv x 5 */
  mjtNum* dof_frictionloss; mjtNum* dof_armature; mjtNum* dof_damping; mjtNum* dof_invweight0;

  /* geoms */
  int* geom_type; int* geom_contype; int* geom_conaffinity; int* geom_condim; int* geom_bodyid;
  int* geom_dataid; int* geom_priority;
  mjtNum* geom_size;     /* ngeom x 3 */
  mjtNum* geom_rbound;   /* ngeom */
  mjtNum* geom_pos;      /* ngeom x 3 */
  mjtNum* geom_quat;     /* ngeom x 4 */
  mjtNum* geom_friction; /* ngeom x 3 */
  mjtNum* geom_solmix; mjtNum* geom_solref; mjtNum* geom_solimp; mjtNum* geom_margin; mjtNum* geom_gap;
  float* geom_rgba;      /* ngeom x 4 */

  /* meshes (convex hull vertices used for collision) */
  int* mesh_vertadr; int* mesh_vertnum;
  mjtNum* mesh_vert;     /* nmeshvert x 3 (MuJoCo stores float; kept double on the host here) */

  /* equality constraints */
  int* eq_type; int* eq_obj1id; int* eq_obj2id; mjtByte* eq_active;
  mjtNum* eq_solref; mjtNum* eq_solimp; mjtNum* eq_data; /* neq x mjNEQDATA */

  /* static candidate geom pairs (extension; canonical contact order = pair order) */
  int* pair_geom1; int* pair_geom2;

  /* sensors (none in the shipped models; fields exist because mj_sim.cpp:973-1014 reads them) */
  int* sensor_type; int* sensor_objid; int* sensor_adr;

  /* names */
  int* name_bodyadr; int* name_jntadr; int* name_geomadr; int* name_meshadr;
  char* names;

  void* owner_;          /* private: C++ storage backing every pointer above */
} mjModel;

typedef struct mjData_ {
  int ncon, nefc, ne, nf;
  mjtNum time;
  mjtNum energy[2];

  mjtNum* qpos; mjtNum* qvel; mjtNum* qacc; mjtNum* qacc_warmstart;
  mjtNum* qfrc_applied;  /* nv */
  mjtNum* xfrc_applied;  /* nbody x 6: force then torque, world frame, at body CoM */
  mjtNum* mocap_pos; mjtNum* mocap_quat;
  mjtNum* sensordata;

  /* position stage */
  mjtNum* xpos; mjtNum* xquat; mjtNum* xmat; mjtNum* xipos; mjtNum* ximat;
  mjtNum* xanchor; mjtNum* xaxis; mjtNum* geom_xpos; mjtNum* geom_xmat;
  mjtNum* subtree_com; mjtNum* cinert; mjtNum* crb; mjtNum* cdof;
  mjtNum* qM; mjtNum* qLD; mjtNum* qLDiagInv;

  /* velocity stage */
  mjtNum* cvel; mjtNum* cdof_dot; mjtNum* qfrc_bias; mjtNum* qfrc_passive;

  /* acceleration stage */
  mjtNum* cacc; mjtNum* cfrc_int;
  mjtNum* qfrc_smooth; mjtNum* qacc_smooth; mjtNum* qfrc_constraint; mjtNum* qfrc_inverse;

  /* constraints */
  mjContact* contact;    /* nconmax */
  int* efc_type; int* efc_id;
  mjtNum* efc_J;         /* njmax x nv, dense row-major */
  mjtNum* efc_pos; mjtNum* efc_margin; mjtNum* efc_frictionloss; mjtNum* efc_diagApprox;
  mjtNum* efc_KBIP;      /* njmax x 4 */
  mjtNum* efc_D; mjtNum* efc_R; mjtNum* efc_vel; mjtNum* efc_aref; mjtNum* efc_b; mjtNum* efc_force;
  mjtNum* efc_AR;        /* njmax x njmax */
  int solver_iter;

  void* owner_;          /* private */
} mjData;

typedef void (*mjfGeneric)(const mjModel* m, mjData* d);
extern mjfGeneric mjcb_control;

/* ---- model / data lifecycle ---- */
mjModel* mj_loadXML(const char* filename, const mjVFS* vfs, char* error, int error_sz);
void mj_saveModel(const mjModel* m, const char* filename, void* buffer, int buffer_sz);  /* binary model image (.mjb) */
mjModel* mj_loadModel(const char* filename, const mjVFS* vfs);
mjModel* mj_loadXMLString(const char* xml, const char* basedir, char* error, int error_sz); /* extension */
int mj_saveLastXML(const char* filename, const mjModel* m, char* error, int error_sz);
mjData* mj_makeData(const mjModel* m);
void mj_deleteData(mjData* d);
void mj_deleteModel(mjModel* m);
void mj_resetData(const mjModel* m, mjData* d);

/* ---- stepping (GPU-backed: env 0 of the batch bound to (m,d); see b2_batch.h) ---- */
void mj_step(const mjModel* m, mjData* d);
void mj_step1(const mjModel* m, mjData* d);
void mj_step2(const mjModel* m, mjData* d);
void mj_forward(const mjModel* m, mjData* d);
void mj_inverse(const mjModel* m, mjData* d);
void mj_mulM(const mjModel* m, const mjData* d, mjtNum* res, const mjtNum* vec);

/* ---- names ---- */
int mj_name2id(const mjModel* m, int type, const char* name);
const char* mj_id2name(const mjModel* m, int type, int id);

/* ---- printing ---- */
void mj_printModel(const mjModel* m, const char* filename);
void mj_printData(const mjModel* m, mjData* d, const char* filename);

/* ---- utilities used by the reference ---- */
void* mju_malloc(unsigned long size);
void mju_free(void* ptr);
void mju_zero(mjtNum* res, int n);
void mju_copy(mjtNum* res, const mjtNum* data, int n);
void mju_addTo3(mjtNum res[3], const mjtNum vec[3]);
void mju_mulQuat(mjtNum res[4], const mjtNum a[4], const mjtNum b[4]);
void mju_rotVecQuat(mjtNum res[3], const mjtNum vec[3], const mjtNum quat[4]);
void mju_mat2Quat(mjtNum quat[4], const mjtNum mat[9]);
void mju_quat2Mat(mjtNum mat[9], const mjtNum quat[4]);
void mju_warning(const char* msg, ...);
void mju_error(const char* msg, ...);
mjtNum mju_abs(mjtNum x);
mjtNum mju_sin(mjtNum x);
mjtNum mju_cos(mjtNum x);
mjtNum mju_sqrt(mjtNum x);
mjtNum mju_ceil(mjtNum x);

#ifdef __cplusplus
}
#endif
#endif /* B2_MUJOCO_SHIM_H_ */
