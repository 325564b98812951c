// k_args.h — kernel argument block shared by host (batch.cu) and device code.
#pragma once
#include <cstdint>

#include "dmodel.h"

namespace b2 {

enum TickFlags : int {
  B2F_CONTROLLER = 1 << 0,   // run MjSim::controller semantics on ddq/dq (reference mj_sim.cpp:1055-1077)
  B2F_INVERSE = 1 << 1,      // compute qfrc_inverse (MjHWInterface::read -> mj_inverse, mj_hw_interface.cpp:61)
  B2F_INTEGRATE = 1 << 2,    // advance the state (mj_step2's Euler); off for mj_forward
  B2F_ODOM = 1 << 3,         // MjSim::set_odom_vels after integration (mj_sim.cpp:1079-1153)
  B2F_FUSED = 1 << 4,        // model has no constraint source at all: the smooth kernel also integrates
  B2F_XFRC = 1 << 5,         // xfrc_applied is non-zero somewhere
  B2F_WS_GLOBAL = 1 << 6,    // workspace in HBM instead of shared memory
  B2F_EXPORT = 1 << 7,
  B2F_NOSOLVE = 1 << 8,      // stop after constraint assembly (mj_step1)
  B2F_FUSABLE = 1 << 9,      // joint limits are the only constraint source: environments without an active limit are
                             // integrated by the smooth kernel (status bit 8) and skipped by the constraint pipeline
  B2F_LD_SMEM = 1 << 11,     // k_smooth (workspace in HBM) factorises M in a shared-memory scratch column per thread
  B2F_READ_POST = 1 << 12,   // hardware read returns post-integration qpos / qvel (default: the reference's pre-integration order)
  B2F_OBS = 1 << 13,         // publish [qpos | qvel] of every environment into the observation buffer of every GPU (peer stores)
  B2F_HWIO = 1 << 10,        // k_chain also does MjHWInterface::write / read (hardware joint j == dof j): commands are read
                             // from, and joint states written to, the hw_* buffers (HBM or mapped host memory)
};

// contact record, SoA: field f of contact c of env e at con[(f * nconmax + c) * nenvp + e]
enum ConField : int {
  CF_DIST = 0, CF_POS = 1, CF_FRAME = 4, CF_INCLUDEMARGIN = 13, CF_FRICTION = 14, CF_SOLREF = 19, CF_SOLIMP = 21,
  CF_NFLOAT = 26
};
enum ConIField : int { CI_GEOM1 = 0, CI_GEOM2 = 1, CI_DIM = 2, CI_PAIR = 3, CI_EFC = 4, CI_NINT = 5 };

template <typename T>
struct KArgs {
  const uint32_t* model;  // packed DModel blob in HBM
  int model_words;        // its size in 32-bit words (so that the TMA staging copy needs no dependent header load)
  int nenv, nenvp;        // environments, padded to a multiple of the CTA size (nenvp is the env stride of every SoA array)
  int ncount;             // environments THIS launch covers (a multiple of 128): nenvp, or a sub-batch window of it whose
                          // pointers were advanced to its first environment (batch.cu: window_args)
  int env_base;           // first environment of the window in the whole batch (0 without windows): observation slices
  int flags;
  int ws_block;           // CTA size the shared workspace was sized for
  T h;                    // timestep of this call (the reference mutates m->opt.timestep every tick, mj_main.cpp:150-163)
  const T* hp;            // the same value in device memory: a captured CUDA graph keeps following b2_set_timestep
#ifdef __CUDACC__
  __device__ __forceinline__ T dt() const { return hp ? *hp : h; }
#endif

  // persistent state, SoA [element][nenvp]
  T *qpos, *qvel, *qacc, *qacc_warmstart, *qfrc_applied, *xfrc_applied, *mocap_pos, *mocap_quat;
  T *ddq, *dq, *odom_vels, *time;
  // per-tick outputs
  T *qfrc_bias, *qfrc_inverse, *xpos, *xquat;
  // stage results exported for the constraint pipeline and the legacy mjData mirror
  T *xmat, *geom_xpos, *geom_xmat, *subtree_com, *cdof, *qM, *qLD, *qLDiagInv, *qfrc_passive, *qfrc_smooth,
      *qacc_smooth, *qfrc_constraint;
  T* ws;                  // global workspace [ws_slots][nenvp] when B2F_WS_GLOBAL

  // contacts and constraint rows
  T* con;                 // [CF_NFLOAT][nconmax][nenvp]
  int* coni;              // [CI_NINT][nconmax][nenvp]
  int* ncon;              // [nenvp]
  int* nefc;              // [nenvp]
  int* efc_type;          // [njmax][nenvp]
  int* efc_id;            // [njmax][nenvp]
  int* efc_tree;          // [njmax][2][nenvp] kinematic trees of the row (-1: none)
  T *efc_J;               // [njmax][wmax][nenvp] compact rows (k_constraint.cuh); a contact's first condim rows hold its base directions
  T *efc_pos, *efc_margin, *efc_frictionloss, *efc_diagApprox, *efc_R, *efc_D, *efc_KBI, *efc_vel, *efc_aref,
      *efc_b, *efc_force, *efc_finv; // [njmax][nenvp] (KBI: [3][njmax][nenvp]); efc_finv: primal force of mj_inverse per row
  T* efc_ARdiag;          // [njmax][nenvp] diagonal of J M^-1 J^T + R
  T* efc_B;               // [njmax][wmax][nenvp] M^-1 J^T of every row, layout of efc_J; only for wide trees (k_solve_rows), else null
  // tensor-core projection (k_project_tc.cuh; one-tree models, fp32): environment-major copies, rows of 64 floats
  T* efc_Jem;             // [nenvp][em_rows][64] J of every row (written next to efc_J by k_make_rows)
  T* efc_Bem;             // [nenvp][em_rows][64] B = J M^-1 from the tcgen05 kernel (read by k_make_blocks instead of efc_B)
  T* minv_em;             // [nenvp][64][64] dense M^-1 (k_dense_minv)
  int em_rows;            // rows per environment in the two arrays above (a multiple of 128 >= njmax); 0: path off
  T* efc_blocks;          // [nenvp][block_capw] environment-major block records streamed by the solver (k_constraint.cuh)
  int* efc_nwords;        // [nenvp] words of efc_blocks in use
  int* env_order;         // [nenvp] visit order of the solver (k_order_envs)
  int *blk_row0, *blk_off; // [njmax][nenvp] block table: first row / word offset of block i (k_make_rows -> k_make_blocks)
  int* nblk;              // [nenvp] blocks of the environment
  // constraint islands (k_make_rows -> k_pgs_island): connected components of the graph "kinematic trees joined by the
  // blocks that touch two of them".  The blocks of an island are contiguous in the slab (original order kept inside it).
  int *isl_off, *isl_end; // [isl_cap][nenvp] word range of island i in the environment's slab
  int* nisl;              // [nenvp] islands of the environment
  int isl_cap;            // island slots allocated per environment (0: no island ordering, the slab is in row order)
  int* maxblk;            // [1] largest block count of this tick (cleared by a memset node in front of k_make_rows)
  int block_capw;         // words of efc_blocks per environment
  int block_npar;         // header + parameter words of the largest block (k_make_blocks' shared-memory column layout)
  int stage_cap;          // words of an environment's records the solver keeps in shared memory
  int wp;                 // (unused) padded compact row width
  int ld_extra;           // k_smooth: the shared-memory factor scratch has a fourth vector (B2F_LD_SMEM)
  int row_tab;            // k_make_rows: a shared-memory row table (one word per row and environment) sits behind its other columns
  int row_nb;             // k_make_rows: base rows its per-thread shared-memory column holds (0: rows are accumulated in HBM)
  int* solver_iter;       // [nenvp]
  int* status;            // [nenvp] bit 0: contact cap hit, bit 1: row cap hit, bit 2: state reset (bad value),
                          //         bit 3: integrated by the smooth kernel this tick (B2F_FUSABLE)
  // hardware-interface exchange, native layout [joint][nenv] fp32; HBM staging or mapped (zero-copy) host memory
  const float *hw_vel, *hw_eff;
  float *hw_pos, *hw_velo, *hw_effo;
  const float *hw_kp, *hw_kd;   // [nhw] PD gains (b2_set_pd): hw_eff holds position targets when non-null
  // observation exchange: obs_peers[p] is GPU p's buffer [world][nq + nv][obs_nenv] fp32 (own memory for
  // p == obs_rank, NVLink peer mappings otherwise); this GPU fills slice obs_rank of every one of them
  float* const* obs_peers;
  int obs_world, obs_rank, obs_nenv;
  int* pending;           // [1] environments that need the constraint pipeline this tick (B2F_FUSABLE); cleared by a
                          //     memset node in front of the smooth kernel
};

}  // namespace b2
