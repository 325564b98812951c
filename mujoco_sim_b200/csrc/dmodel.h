// dmodel.h — device-side model blob: every per-model constant the kernels need, packed as 32-bit words behind a
// header of word offsets.  The blob is built once per batch on the host (batch.cu: pack_model), lives in HBM and is
// staged into shared memory by one TMA bulk copy per CTA (k_common.cuh: stage_model).  Field names are MuJoCo's.
#pragma once
#include <cstdint>

namespace b2 {

// name, element kind (I = int32, F = real), count expression over the size fields
#define B2_MODEL_ARRAYS(X)                                                                                      \
  X(body_parentid, I, nbody) X(body_rootid, I, nbody) X(body_weldid, I, nbody) X(body_mocapid, I, nbody)       \
  X(body_jntnum, I, nbody) X(body_jntadr, I, nbody) X(body_dofnum, I, nbody) X(body_dofadr, I, nbody)          \
  X(body_lastdof, I, nbody)                                                                                     \
  X(body_pos, F, 3 * nbody) X(body_quat, F, 4 * nbody) X(body_ipos, F, 3 * nbody) X(body_iquat, F, 4 * nbody)  \
  X(body_mass, F, nbody) X(body_subtreemass, F, nbody) X(body_inertia, F, 3 * nbody)                           \
  X(body_invweight0, F, 2 * nbody) X(body_gravcomp, F, nbody)                                                  \
  X(jnt_type, I, njnt) X(jnt_qposadr, I, njnt) X(jnt_dofadr, I, njnt) X(jnt_bodyid, I, njnt)                   \
  X(jnt_limited, I, njnt) X(jnt_pos, F, 3 * njnt) X(jnt_axis, F, 3 * njnt) X(jnt_stiffness, F, njnt)           \
  X(jnt_range, F, 2 * njnt) X(jnt_margin, F, njnt) X(jnt_solref, F, 2 * njnt) X(jnt_solimp, F, 5 * njnt)       \
  X(qpos0, F, nq) X(qpos_spring, F, nq)                                                                        \
  X(dof_bodyid, I, nv) X(dof_jntid, I, nv) X(dof_parentid, I, nv) X(dof_Madr, I, nv) X(dof_controlled, I, nv)  \
  X(dof_Mcnt, I, nv) X(dof_anc, I, nM) /* row i of M: dof_Mcnt[i] entries, entry a belongs to dof_anc[dof_Madr[i] + a] */ \
  X(dof_armature, F, nv) X(dof_damping, F, nv) X(dof_frictionloss, F, nv) X(dof_invweight0, F, nv)             \
  X(dof_solref, F, 2 * nv) X(dof_solimp, F, 5 * nv)                                                            \
  X(geom_type, I, ngeom) X(geom_bodyid, I, ngeom) X(geom_condim, I, ngeom) X(geom_priority, I, ngeom)          \
  X(geom_size, F, 3 * ngeom) X(geom_rbound, F, ngeom) X(geom_pos, F, 3 * ngeom) X(geom_quat, F, 4 * ngeom)     \
  X(geom_friction, F, 3 * ngeom) X(geom_solmix, F, ngeom) X(geom_solref, F, 2 * ngeom)                         \
  X(geom_solimp, F, 5 * ngeom) X(geom_margin, F, ngeom) X(geom_gap, F, ngeom)                                  \
  X(geom_vertadr, I, ngeom) X(geom_vertnum, I, ngeom) /* mesh geoms: vertex range in the mesh_vert tail */       \
  X(eq_type, I, neq) X(eq_obj1id, I, neq) X(eq_obj2id, I, neq) X(eq_active, I, neq)                            \
  X(eq_solref, F, 2 * neq) X(eq_solimp, F, 5 * neq) X(eq_data, F, 11 * neq)                                    \
  X(pair_geom1, I, npair) X(pair_geom2, I, npair)                                                              \
  X(odom_dof, I, 6 * nodom) X(odom_qpos, I, 3 * nodom)                                                          \
  X(body_treeid, I, nbody) X(dof_treeid, I, nv) X(tree_dofadr, I, ntree) X(tree_dofnum, I, ntree)              \
  /* tree-parallel kernels: per-lane item lists (kind 0 body, 1 dof, 2 joint, 3 geom; lane_off[9 * kind + lane] .. [+ 1]) */ \
  X(lane_off, I, 36) X(lane_body, I, nbody) X(lane_dof, I, nv) X(lane_jnt, I, njnt) X(lane_geom, I, ngeom)     \
  X(opt_real, F, 8) /* gravity[3], tolerance, meaninertia, impratio, 0, 0 in batch precision */

// Workspace arrays (per environment, strided by the workspace stride): name, count expression
#define B2_WS_ARRAYS(X)                                                                                        \
  X(xpos, 3 * nbody) X(xquat, 4 * nbody) X(xmat, 9 * nbody) X(xipos, 3 * nbody) X(ximat, 9 * nbody)            \
  X(xanchor, 3 * njnt) X(xaxis, 3 * njnt) X(subtree_com, 3 * nbody) X(cinert, 10 * nbody) X(crb, 10 * nbody)   \
  X(cdof, 6 * nv) X(cvel, 6 * nbody) X(cdof_dot, 6 * nv) X(cacc, 6 * nbody) X(cfrc, 6 * nbody)                 \
  X(qM, nM) X(qLD, nM) X(qLDiagInv, nv) X(qfrc_passive, nv) X(qfrc_smooth, nv) X(qacc_smooth, nv) X(tmpv, nv)

struct DModel {
  // sizes
  int nq, nv, nbody, njnt, ngeom, nM, neq, npair, nconmax, njmax, nmocap, nodom;
  int ntree, wmax;  // kinematic trees with dofs; widest compact constraint row (dofs of the two largest trees)
  int o_mesh_vert;  // word offset of the mesh vertices (kind F, 3 per vertex): BEHIND nwords, read from HBM, never staged
  int nmeshvert;
  int disableflags, enableflags, iterations, nwords;  // nwords: blob size in 32-bit words (multiple of 4)
  int has_damping, has_gravcomp, has_stiffness, has_limits, has_frictionloss, has_controlled, has_xfrc, pad0;
  float gravity[3];
  float tolerance, meaninertia, impratio;
  int real_bytes;  // 4: arrays of kind F are float, 8: double (two words per element)
  int pad1;
#define X(name, kind, count) int o_##name;
  B2_MODEL_ARRAYS(X)
#undef X
#define X(name, count) int w_##name;  // workspace slot offsets (in elements)
  B2_WS_ARRAYS(X)
#undef X
  int ws_slots;  // total workspace elements per environment
  int pad2[3];
};

}  // namespace b2
