/* oracle.h — CPU restatement (fp64, scalar, one environment) of the per-tick hot path of
 * HoangGiang93/mujoco_sim.  TEST INFRASTRUCTURE ONLY: nothing in mujoco_sim_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in MuJoCo 2.3.7 (closed binary dependency fetched by the
 * reference's Makefile:3-13, linked at CMakeLists.txt:65,80,97,115), which is absent from /root/reference and from
 * this image; the reference ships no golden vectors.  Each function restates the
 * published MuJoCo algorithm ("Computation" chapter) and is anchored on the reference's call sites:
 *   src/mj_main.cpp:82-112 (tick order), src/mujoco_sim/mj_sim.cpp:1055-1077 (controller),
 *   src/mujoco_sim/mj_hw_interface.cpp:59-91 (read/mj_inverse, write), src/mujoco_sim/mj_sim.cpp:1079-1153 (odom).
 * It is pinned instead by analytic known-answer tests (tests/test_oracle_known_answers.py).
 */
#ifndef B2_ORACLE_H_
#define B2_ORACLE_H_
#include "mujoco/mujoco.h"

#ifdef __cplusplus
extern "C" {
#endif

/* position stage */
void omj_kinematics(const mjModel* m, mjData* d);
void omj_comPos(const mjModel* m, mjData* d);
void omj_crb(const mjModel* m, mjData* d);
void omj_factorM(const mjModel* m, mjData* d);
void omj_collision(const mjModel* m, mjData* d);
void omj_makeConstraint(const mjModel* m, mjData* d);
void omj_projectConstraint(const mjModel* m, mjData* d);
void omj_fwdPosition(const mjModel* m, mjData* d);
/* velocity stage */
void omj_comVel(const mjModel* m, mjData* d);
void omj_passive(const mjModel* m, mjData* d);
void omj_referenceConstraint(const mjModel* m, mjData* d);
void omj_rne(const mjModel* m, mjData* d, int flg_acc, mjtNum* result);
void omj_fwdVelocity(const mjModel* m, mjData* d);
/* acceleration stage */
void omj_fwdAcceleration(const mjModel* m, mjData* d);
void omj_fwdConstraint(const mjModel* m, mjData* d);
void omj_Euler(const mjModel* m, mjData* d);
void omj_energy(const mjModel* m, mjData* d);
void omj_rnePostConstraint(const mjModel* m, mjData* d);  /* cacc, cfrc_int with applied wrenches and contact forces */
/* top level (mirror mj_step1 / mj_step2 / mj_forward / mj_inverse / mj_mulM) */
void omj_step1(const mjModel* m, mjData* d);
void omj_step2(const mjModel* m, mjData* d);
void omj_step(const mjModel* m, mjData* d);
void omj_forward(const mjModel* m, mjData* d);
void omj_inverse(const mjModel* m, mjData* d);
void omj_mulM(const mjModel* m, const mjData* d, mjtNum* res, const mjtNum* vec);
void omj_solveM(const mjModel* m, const mjData* d, mjtNum* x, int n);
void omj_fullM(const mjModel* m, const mjData* d, mjtNum* dst);
void omj_jac(const mjModel* m, const mjData* d, mjtNum* jacp, mjtNum* jacr, const mjtNum point[3], int body);

/* in-tree hot functions of the reference */
/* MjSim::controller (src/mujoco_sim/mj_sim.cpp:1055-1077): tau = M ddq (+ bias on controlled dofs); qfrc_applied = tau;
 * qvel override by dq; zero the commands. `controlled` is a per-dof 0/1 mask. */
void omj_controller(const mjModel* m, mjData* d, mjtNum* ddq, mjtNum* dq, const mjtByte* controlled);
/* MjSim::set_odom_vels (src/mujoco_sim/mj_sim.cpp:1079-1153). dof/qpos indices of the six odom joints
 * (lin x,y,z then ang x,y,z), -1 when absent; vels = commanded twist in the odom frame. */
void omj_set_odom_vels(const mjModel* m, mjData* d, const int lin_dof[3], const int ang_dof[3], const int ang_qpos[3],
                       const mjtNum vels[6]);
/* one full tick in the order of src/mj_main.cpp:82-112:
 * step1 -> controller -> (read: mj_inverse) -> step2.  do_inverse mirrors "at least one robot". */
void omj_tick(const mjModel* m, mjData* d, mjtNum* ddq, mjtNum* dq, const mjtByte* controlled, int do_inverse);

/* CPU baseline driver: advance nenv environments nsteps ticks each, one std::thread per pool entry.
 * state arrays are [nenv][n] row-major doubles; pool holds one mjData per thread. Returns threads used. */
int omj_tick_batch(const mjModel* m, mjData** pool, int npool, int nenv, int nsteps, mjtNum* qpos, mjtNum* qvel,
                   mjtNum* qacc_warmstart, const mjtNum* qfrc_applied, const mjtNum* ddq, const mjtNum* dq,
                   const mjtByte* controlled, int do_inverse, mjtNum* qfrc_inverse_out);
int omj_tick_batch_pd(const mjModel* m, mjData** pool, int npool, int nenv, int nsteps, mjtNum* qpos, mjtNum* qvel,
                   mjtNum* qacc_warmstart, const mjtNum* qfrc_applied, const mjtNum* ddq, const mjtNum* dq,
                   const mjtByte* controlled, int do_inverse, mjtNum* qfrc_inverse_out, const mjtNum* pd_kp, const mjtNum* pd_kd);

#ifdef __cplusplus
}
#endif
#endif
