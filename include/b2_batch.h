/* b2_batch.h — thin extern "C" ABI of the B200 batched rigid-body step.
 *
 * This is the drop-in boundary for the per-tick hot path of HoangGiang93/mujoco_sim.  The reference has no plugin
 * interface: its hot path is the body of simulate() (src/mj_main.cpp:82-112), which calls the MuJoCo C API on ONE
 * environment.  The MuJoCo-named entry points that replace those calls are declared in include/mujoco/mujoco.h;
 * they are implemented on top of the batched entry points below (environment 0 of a batch), and a C++ host that
 * wants N environments calls these directly.  Plain pointers and sizes only; every function returns 0 on success and
 * a negative code on failure (b2_last_error() has the text), and never throws.  There is no CPU fallback: without a
 * usable CUDA device b2_create fails.
 *
 * Reference interfaces replaced (file:line):
 *   b2_tick          the loop body                      src/mj_main.cpp:82-112
 *     B2_TICK_CONTROLLER   MjSim::controller            src/mujoco_sim/mj_sim.cpp:1055-1077 (mjcb_control, mj_main.cpp:49-52,196)
 *     B2_TICK_INVERSE      MjHWInterface::read/mj_inverse  src/mujoco_sim/mj_hw_interface.cpp:59-71
 *     B2_TICK_ODOM         MjSim::set_odom_vels         src/mujoco_sim/mj_sim.cpp:1079-1153
 *   b2_write_commands  MjHWInterface::write            src/mujoco_sim/mj_hw_interface.cpp:73-91
 *   b2_read_joints     MjHWInterface::read (gather)    src/mujoco_sim/mj_hw_interface.cpp:62-70
 *   b2_forward         mj_forward                       src/mujoco_sim/mj_ros.cpp:608,1421
 *   b2_set_controlled  MjSim::controlled_joints         src/mujoco_sim/mj_ros.cpp:634-668 (producer), mj_sim.cpp:1058-1063
 *   b2_set_timestep    m->opt.timestep mutation         src/mj_main.cpp:150-163
 */
#ifndef B2_BATCH_H_
#define B2_BATCH_H_
#include "mujoco/mujoco.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2_batch b2_batch;

enum {
  B2_TICK_CONTROLLER = 1 << 0,
  B2_TICK_INVERSE = 1 << 1,
  B2_TICK_INTEGRATE = 1 << 2,
  B2_TICK_ODOM = 1 << 3,
  B2_TICK_NOSOLVE = 1 << 9,  /* stop after constraint assembly (mj_step1) */
  B2_TICK_READ_POST = 1 << 10 /* b2_tick_host / b2_tick_resident: return the joint positions / velocities AFTER the integration
                                 of this tick (what read() of the next tick would see) instead of the reference's order, where
                                 MjHWInterface::read runs between mj_step1 and mj_step2 (src/mj_main.cpp:91-108) */
};
enum { B2_F32 = 4, B2_F64 = 8, B2_EXPORT_STAGES = 0x100 /* OR into precision: keep stage arrays (qM, xmat, geom poses ...) readable even for contact-free models */ };
enum { B2_ENV_MAJOR = 0, /* host buffer is [env][n] */ B2_NATIVE = 1 /* host buffer is [n][env] */ };

const char* b2_last_error(void);
int b2_device_count(void);

/* nenv environments of model m on CUDA device `device`, state initialised to qpos0.  precision: B2_F32 (product
 * path) or B2_F64 (validation of the algorithm against the fp64 oracle without rounding noise). */
b2_batch* b2_create(const mjModel* m, int nenv, int device, int precision);
void b2_destroy(b2_batch* b);
const char* b2_path_name(const b2_batch* b); /* kernels a tick launches, e.g. "k_chain<7>" */
int b2_nenv(const b2_batch* b);
int b2_nenv_padded(const b2_batch* b);
int b2_precision(const b2_batch* b);

/* per-dof "controlled" mask (length nv) consumed by the controller: tau[dof] += qfrc_bias[dof] */
int b2_set_controlled(b2_batch* b, const unsigned char* mask);
/* odom joints of nrobot robots: dof[6*r+k] = dof index of lin x,y,z / ang x,y,z odom joint (-1: absent/disabled),
 * qposadr[3*r+k] = qpos address of the angular x,y,z odom joints (-1: absent -> angle 0) */
int b2_set_odom(b2_batch* b, int nrobot, const int* dof, const int* qposadr);
int b2_set_timestep(b2_batch* b, double h);
/* iterations, tolerance, disableflags (mjOption); scheduling only, results do not depend on them: subbatches (1..8,
 * default 1: windows of the batch that run the tick's kernels side by side on their own streams) and subbatch_min
 * (smallest window in environments, default 2048; a batch too small for two such windows is one window) */
int b2_set_option(b2_batch* b, const char* name, double value);

/* field access by MuJoCo name: qpos qvel qacc qacc_warmstart qfrc_applied xfrc_applied mocap_pos mocap_quat ddq dq
 * odom_vels time qfrc_bias qfrc_inverse xpos xquat xmat geom_xpos geom_xmat subtree_com cdof qM qLD qLDiagInv
 * qfrc_passive qfrc_smooth qacc_smooth qfrc_constraint efc_J efc_pos efc_margin efc_frictionloss efc_diagApprox
 * efc_R efc_D efc_vel efc_aref efc_b efc_force efc_AR contact (26 numbers per contact)
 * and the int fields ncon nefc efc_type efc_id contact_int (5 per contact) solver_iter status.
 * Host buffers hold environments [env_lo, env_hi). Returns the per-environment element count. */
int b2_field_size(const b2_batch* b, const char* field);
int b2_set_field_f32(b2_batch* b, const char* field, const float* host, int env_lo, int env_hi, int layout);
int b2_set_field_f64(b2_batch* b, const char* field, const double* host, int env_lo, int env_hi, int layout);
int b2_get_field_f32(b2_batch* b, const char* field, float* host, int env_lo, int env_hi, int layout);
int b2_get_field_f64(b2_batch* b, const char* field, double* host, int env_lo, int env_hi, int layout);
int b2_get_field_i32(b2_batch* b, const char* field, int* host, int env_lo, int env_hi, int layout);
/* raw device pointer of a field (SoA [element][nenv_padded], element type = batch precision or int32) */
void* b2_device_ptr(b2_batch* b, const char* field);

/* reset environments [env_lo, env_hi) to qpos0 / zero velocity */
int b2_reset(b2_batch* b, int env_lo, int env_hi);
/* one tick of every environment (asynchronous on the batch's stream) */
int b2_tick(b2_batch* b, int flags);
int b2_step(b2_batch* b, int nsteps);   /* nsteps x tick(INTEGRATE [+ CONTROLLER/INVERSE/ODOM as configured by b2_set_tick_flags]) */
int b2_set_tick_flags(b2_batch* b, int flags);
int b2_forward(b2_batch* b);            /* mj_forward: everything but the integration */
int b2_sync(b2_batch* b);
void* b2_stream(b2_batch* b);           /* cudaStream_t the batch launches on */
long long b2_launch_count(const b2_batch* b); /* kernels launched so far */

/* hardware-interface exchange for `njoint` scalar joints (jnt ids), all environments, native layout [joint][env]:
 * write: per controlled joint, |vel_cmd| > mjMINVAL -> dq[dof] = vel_cmd else ddq[dof] = effort_cmd
 * read : position, velocity, effort(qfrc_inverse) gathers */
int b2_set_hw_joints(b2_batch* b, int njoint, const int* jnt_ids);
int b2_write_commands(b2_batch* b, const float* vel_cmd_host, const float* effort_cmd_host);
int b2_read_joints(b2_batch* b, float* pos_host, float* vel_host, float* effort_host);
/* device-side PD stage in front of write(): with gains set (kp, kd per hardware joint; NULL, NULL switches it off) the
 * effort-command buffer carries position TARGETS q* and the command consumed by write() becomes
 * kp (q* - q) - kd qdot, evaluated per environment on the GPU.  This is what the reference's ros_control PID
 * controllers compute on the host for its single environment (gains: model/ontology/box/box.yaml:5-13; the result is
 * read as a desired acceleration, src/mujoco_sim/mj_hw_interface.cpp:73-91). */
int b2_set_pd(b2_batch* b, const float* kp, const float* kd);
/* end-to-end tick through host buffers: H2D commands, tick, D2H joint states, synchronised.  Order of the reference
 * (src/mj_main.cpp:82-112, mj_hw_interface.cpp:59-71): the joint states are those read() sees between mj_step1 and
 * mj_step2: qpos of the tick's start, qvel after the controller's velocity override, qfrc_inverse of this tick
 * (B2_TICK_READ_POST in b2_set_tick_flags switches to the post-integration positions / velocities). */
int b2_tick_host(b2_batch* b, const float* vel_cmd_host, const float* effort_cmd_host, float* pos_host, float* vel_host,
                 float* effort_host);

/* Zero-copy exchange: buffers of b2_tick_host that are device memory or host memory the caller pinned (cudaHostAlloc,
 * cudaHostRegister or b2_register_host below) are read / written in place by the tick kernels; pageable buffers are
 * staged through HBM with two + three copies.  b2_register_host pins a caller-owned range for that purpose; the caller
 * keeps it alive and calls b2_unregister_host before freeing it (b2_destroy unregisters what is left). */
int b2_register_host(b2_batch* b, void* host, long long bytes);
int b2_unregister_host(b2_batch* b, void* host);

/* the same control tick with the command buffers of the last upload re-issued from HBM and the joint states left in
 * HBM: no host<->device traffic, asynchronous (throughput with resident inputs) */
int b2_tick_resident(b2_batch* b);

/* Runtime spawn / destroy as SLOT ACTIVATION (reference: the spawn_objects / destroy_objects services re-author the
 * world XML and reload the whole model with the simulation thread blocked, src/mujoco_sim/mj_ros.cpp:906-1428,
 * mj_sim.cpp:465-558,573-710).  Here the model is compiled once with `nslot` free bodies that serve as object slots;
 * every environment carries an active flag per slot.  An inactive slot is parked far above the scene at rest (it is
 * held there after every tick, collides with nothing and produces no constraint rows), spawning writes a pose and a
 * twist into the slot's free joint of one environment and activates it, destroying parks it again.
 *   b2_set_slots     body ids of the slot bodies (each must carry exactly one free joint); all slots start ACTIVE
 *   b2_spawn         n requests: env[i], slot[i], pose7[i] = x y z qw qx qy qz, twist6[i] = v(3) w(3) (NULL: at rest)
 *   b2_destroy_slots n requests: env[i], slot[i]
 *   b2_slot_active   read back the flags of environments [env_lo, env_hi) as [env][slot] bytes
 * Host arrays; the requests are applied in order on the batch's stream before the next tick. */
int b2_set_slots(b2_batch* b, int nslot, const int* body_ids);
int b2_spawn(b2_batch* b, int n, const int* env, const int* slot, const float* pose7, const float* twist6);
int b2_destroy_slots(b2_batch* b, int n, const int* env, const int* slot);
int b2_slot_active(b2_batch* b, unsigned char* active, int env_lo, int env_hi);

/* Full-recompile fallback for topology changes the slots cannot express (MESH / whole-robot spawn, mj_ros.cpp:941-1325):
 * the caller compiles the new world (mj_loadXML), creates a batch for it and carries the old state over as the
 * reference's add_old_state does (src/mujoco_sim/mj_sim.cpp:465-558): for every body NAME present in both models with
 * the same joint / dof counts, qpos, qvel, qacc, qacc_warmstart, qfrc_applied of its joints are copied for every
 * environment (device to device), and the simulation time; bodies only in the new model keep their qpos0.  Both batches
 * must have the same environment count and precision.  Returns the number of bodies carried over, < 0 on error. */
int b2_transfer_state(b2_batch* src, b2_batch* dst);

/* observation exchange ("at most one all-gather of observations per control tick"): pack
 * [qpos | qvel] of every environment of this shard as fp32, native layout [nq + nv][nenv], into the DEVICE buffer
 * obs_dev ((nq + nv) * nenv floats) on the batch's stream.  The collective itself is the caller's: one process per GPU,
 * ncclAllGather / torch.distributed.all_gather_into_tensor of that buffer on the same stream (bench.py does exactly that). */
int b2_pack_obs(b2_batch* b, float* obs_dev);

/* The same exchange fused into the tick (no pack kernel, no collective call): every GPU owns a buffer
 * [world][nq + nv][nenv] fp32 and each tick's integrate epilogue stores the new state of its environments into slice
 * `rank` of EVERY GPU's buffer through NVLink peer mappings.
 *   b2_obs_create   allocate this GPU's buffer for `world` shards of nenv environments; returns its device pointer
 *   b2_obs_handle   64-byte CUDA IPC handle of the buffer (one process per GPU: all-gather the handles once, at setup)
 *   b2_obs_attach   open the peers: `handles` = world x 64 bytes (IPC, other processes) or `ptrs` = world device
 *                   pointers (one process driving several devices, peer access is enabled here); switches the exchange on
 *   b2_obs_enable   switch the exchange on / off for the following ticks
 * A reader of the buffer synchronises with the writers' streams first (a barrier after the tick). */
float* b2_obs_create(b2_batch* b, int world, int rank);
int b2_obs_handle(b2_batch* b, void* handle64);
int b2_obs_attach(b2_batch* b, const void* handles, float* const* ptrs);
int b2_obs_enable(b2_batch* b, int on);
int b2_obs_read(b2_batch* b, float* host);   /* copy this GPU's whole buffer to the host (after synchronising its stream) */

/* One host process driving several devices (b2_create(m, nenv, devices, ndev)): contiguous shards of
 * nenv environments, one batch and stream per device; b2_multi_tick launches the tick on every device before it returns
 * (the devices run concurrently), b2_multi_sync waits for all of them.  With obs != 0 the fused observation exchange is
 * set up between the shards (peer access).  b2_multi_shard gives access to shard i for state I/O. */
typedef struct b2_multi b2_multi;
b2_multi* b2_create_multi(const mjModel* m, int nenv, const int* devices, int ndev, int precision, int obs);
void b2_multi_destroy(b2_multi* mb);
int b2_multi_count(const b2_multi* mb);
b2_batch* b2_multi_shard(b2_multi* mb, int i);
int b2_multi_tick(b2_multi* mb, int flags);
int b2_multi_sync(b2_multi* mb);

/* benchmarking aid: write `bytes` of scratch on the batch's stream so that the state leaves the L2 between timed steps */
int b2_l2_flush(b2_batch* b, long long bytes);

/* per-kernel device timing with CUDA events on the batch's stream: profile up to max_ticks ticks, then read the summed
 * milliseconds per kernel slot {hw_write, smooth, collide, make_constraint, project, pgs, integrate, hw_read} */
int b2_profile_begin(b2_batch* b, int max_ticks);
int b2_profile_end(b2_batch* b, double* ms_per_slot, int nslot);

/* copy environment `env` into the legacy single-environment mjData view (fields of SURVEY.md Appendix C) */
int b2_mirror_env(b2_batch* b, int env, mjData* d);
/* load environment `env` from an mjData (qpos qvel qacc qacc_warmstart qfrc_applied xfrc_applied mocap time) */
int b2_load_env(b2_batch* b, int env, const mjData* d);

#ifdef __cplusplus
}
#endif
#endif
