// model_store.h — C++ storage behind the mjModel / mjData views declared in include/mujoco/mujoco.h.
#pragma once
#include <string>
#include <vector>

#include "mujoco/mujoco.h"

namespace b2 {

struct ModelStore {
  mjModel view{};  // must stay first: mjModel* <-> ModelStore* by owner_
  std::string source_xml;  // canonical MJCF text re-emitted by mj_saveLastXML
  std::string source_dir;

  std::vector<double> qpos0, qpos_spring;
  std::vector<int> body_parentid, body_rootid, body_weldid, body_mocapid, body_jntnum, body_jntadr, body_dofnum,
      body_dofadr, body_geomnum, body_geomadr;
  std::vector<double> body_pos, body_quat, body_ipos, body_iquat, body_mass, body_subtreemass, body_inertia,
      body_invweight0, body_gravcomp;
  std::vector<int> jnt_type, jnt_qposadr, jnt_dofadr, jnt_bodyid;
  std::vector<mjtByte> jnt_limited;
  std::vector<double> jnt_solref, jnt_solimp, jnt_pos, jnt_axis, jnt_stiffness, jnt_range, jnt_margin;
  std::vector<int> dof_bodyid, dof_jntid, dof_parentid, dof_Madr;
  std::vector<double> dof_solref, dof_solimp, dof_frictionloss, dof_armature, dof_damping, dof_invweight0;
  std::vector<int> geom_type, geom_contype, geom_conaffinity, geom_condim, geom_bodyid, geom_dataid, geom_priority;
  std::vector<double> geom_size, geom_rbound, geom_pos, geom_quat, geom_friction, geom_solmix, geom_solref,
      geom_solimp, geom_margin, geom_gap;
  std::vector<float> geom_rgba;
  std::vector<int> mesh_vertadr, mesh_vertnum;
  std::vector<double> mesh_vert;
  std::vector<int> eq_type, eq_obj1id, eq_obj2id;
  std::vector<mjtByte> eq_active;
  std::vector<double> eq_solref, eq_solimp, eq_data;
  std::vector<int> pair_geom1, pair_geom2;
  std::vector<int> sensor_type, sensor_objid, sensor_adr;
  std::vector<int> name_bodyadr, name_jntadr, name_geomadr, name_meshadr;
  std::vector<char> names;
  std::vector<int> exclude_signature;  // (body1<<16)|body2 with body1<body2

  // Point every mjModel pointer at the vectors above. A spare element is kept in front of the
  // jnt_qposadr / jnt_dofadr arrays because the reference indexes them with -1 when an odom joint
  // is absent (src/mujoco_sim/mj_sim.cpp:1083-1091).
  void finalize();
  // binary image of a compiled model (mj_saveModel / mj_loadModel): header + every array above + the source text
  void save(const std::string& path) const;
  static ModelStore* load(const std::string& path);
  std::vector<int> jnt_qposadr_padded, jnt_dofadr_padded;
};

struct DataStore {
  mjData view{};
  std::vector<double> buf;       // one arena for every mjtNum array
  std::vector<int> ibuf;
  std::vector<mjContact> contacts;
};

mjData* make_data(const mjModel* m);
void reset_data(const mjModel* m, mjData* d);

// String-keyed access for the Python/ctypes host mirror and for tests. Returns element count, -1 if unknown.
int model_int(const mjModel* m, const char* name, int* out);
int model_array(const mjModel* m, const char* name, const void** ptr, int* is_int);
int data_array(const mjModel* m, mjData* d, const char* name, void** ptr, int* is_int);

}  // namespace b2
