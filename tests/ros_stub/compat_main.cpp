// compat_main.cpp — drives the reference's OWN hot-path sources (compiled in place from /root/reference by
// tests/test_source_compat.py: src/mujoco_sim/mj_hw_interface.cpp whole, MjSim::controller / MjSim::set_odom_vels
// extracted from src/mujoco_sim/mj_sim.cpp at test time) against libb2sim.so, in the order of the reference's loop
// body (src/mj_main.cpp:82-112).  What is defined HERE is only what the rest of the ROS node would provide: the
// globals of mj_model.cpp, the statics of MjSim, a stand-in for the controller manager, and main().
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>

#include "mj_hw_interface.h"

mjModel* m = NULL;
mjData* d = NULL;
std::mutex mtx;
double rtf = 0.0;

double MjSim::max_time_step;
std::map<std::string, std::vector<std::string>> MjSim::joint_names;
std::map<std::string, mjtNum> MjSim::odom_vels;
std::set<std::string> MjSim::robot_link_names;
mjtNum* MjSim::dq = NULL;
mjtNum* MjSim::ddq = NULL;
mjtNum* MjSim::tau = NULL;
mjtNum MjSim::sim_start;
std::map<std::string, std::map<std::string, bool>> MjSim::add_odom_joints;
std::set<std::string> MjSim::controlled_joints;
std::set<std::string> MjSim::robot_names;
std::map<size_t, std::string> MjSim::sensors;
std::map<std::string, std::vector<float>> MjSim::pose_inits;
bool MjSim::reload_mesh = true;
std::set<std::string> MjSim::spawned_object_body_names;
std::map<int, std::vector<mjtNum>> MjSim::geom_pose;
bool MjSim::disable_gravity = true;
MjSim::~MjSim() {}

static MjSim& mj_sim = MjSim::get_instance();
static void controller(const mjModel*, mjData*) { mj_sim.controller(); }  // as in src/mj_main.cpp:49-52

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: compat_tick model.xml nticks\n"); return 2; }
  char err[1000] = "";
  m = mj_loadXML(argv[1], NULL, err, 1000);
  if (!m) { std::fprintf(stderr, "load: %s\n", err); return 1; }
  d = mj_makeData(m);
  const int nticks = std::atoi(argv[2]);
  // what MjSim::init / MjRos would have set up
  const std::string robot = "robot";
  MjSim::robot_names.insert(robot);
  MjSim::tau = (mjtNum*)mju_malloc(m->nv * sizeof(mjtNum));
  MjSim::ddq = (mjtNum*)mju_malloc(m->nv * sizeof(mjtNum));
  MjSim::dq = (mjtNum*)mju_malloc(m->nv * sizeof(mjtNum));
  mju_zero(MjSim::tau, m->nv); mju_zero(MjSim::ddq, m->nv); mju_zero(MjSim::dq, m->nv);
  for (int j = 0; j < m->njnt; j++) {
    const char* name = mj_id2name(m, mjOBJ_JOINT, j);
    if (!name) continue;
    const std::string s = name;
    if (s.find("_odom_") != std::string::npos) continue;  // set_joint_names() skips odom joints (mj_sim.cpp:70)
    MjSim::joint_names[robot].push_back(s);
    MjSim::controlled_joints.insert(s);
  }
  for (const char* k : {"lin_odom_x_joint", "lin_odom_y_joint", "lin_odom_z_joint", "ang_odom_x_joint", "ang_odom_y_joint", "ang_odom_z_joint"}) {
    MjSim::add_odom_joints[robot][k] = true;
    MjSim::odom_vels[robot + "_" + k] = 0;
  }
  MjSim::odom_vels[robot + "_lin_odom_x_joint"] = 0.4;
  MjSim::odom_vels[robot + "_lin_odom_y_joint"] = -0.1;
  MjSim::odom_vels[robot + "_ang_odom_z_joint"] = 0.3;
  mjcb_control = controller;  // src/mj_main.cpp:196
  MjHWInterface hw(robot);
  auto* eff = hw.get<hardware_interface::EffortJointInterface>();
  auto* vel = hw.get<hardware_interface::VelocityJointInterface>();
  const std::vector<std::string>& names = MjSim::joint_names[robot];
  for (int t = 0; t < nticks; t++) {
    mtx.lock();
    mj_step1(m, d);                      // :83
    hw.read();                           // :91-94  (mj_inverse + gathers)
    // stand-in for ControllerManager::update(): a fixed computed-torque command plus one velocity command
    for (size_t i = 0; i < names.size(); i++) {
      const double q = eff->getHandle(names[i]).getPosition(), qd = eff->getHandle(names[i]).getVelocity();
      eff->getHandle(names[i]).setCommand(20.0 * (0.3 * (i + 1) - q) - 4.0 * qd);
      vel->getHandle(names[i]).setCommand((i == 1 && t % 10 == 3) ? 0.2 : 0.0);
    }
    hw.write();                          // :103-106
    mj_step2(m, d);                      // :108
    mj_sim.set_odom_vels();              // :110
    mtx.unlock();
  }
  std::printf("{\"time\": %.17g, \"qpos\": [", d->time);
  for (int i = 0; i < m->nq; i++) std::printf("%s%.17g", i ? ", " : "", d->qpos[i]);
  std::printf("], \"qvel\": [");
  for (int i = 0; i < m->nv; i++) std::printf("%s%.17g", i ? ", " : "", d->qvel[i]);
  std::printf("], \"effort\": [");
  for (size_t i = 0; i < names.size(); i++) std::printf("%s%.17g", i ? ", " : "", eff->getHandle(names[i]).getEffort());
  std::printf("]}\n");
  mj_deleteData(d);
  mj_deleteModel(m);
  return 0;
}
