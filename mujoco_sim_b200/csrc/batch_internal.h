// batch_internal.h — host-side state of a batch, shared by the translation units that launch kernels.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "b2_batch.h"
#include "dmodel.h"
#include "k_args.h"

namespace b2 {
struct Field {
  void* ptr = nullptr;
  int count = 0;   // elements per environment
  int kind = 0;    // 0 real (batch precision), 1 int32
};
int set_error(const std::string& msg);  // stores the text for b2_last_error(), returns -1
}  // namespace b2

struct b2_batch {
  const mjModel* m = nullptr;
  int nenv = 0, nenvp = 0, device = 0, prec = 4, nsm = 148;
  double h = 0.002;
  cudaStream_t stream = nullptr;
  b2::DModel hdr{};
  std::vector<uint32_t> blob;
  uint32_t* blob_dev = nullptr;
  std::vector<unsigned char> controlled;
  std::vector<int> odom_dof, odom_qpos;
  std::map<std::string, b2::Field> fields;
  std::vector<void*> allocs;
  int tick_flags = 0;
  bool fused = false, ws_global = false, export_stages = false;
  int chain_n = 0;       // > 0: the model is a serial chain of chain_n scalar joints (ChainP<chain_n> kernels)
  int chain_variant = 0; // register budget variant of the chain kernel (B2_CHAIN_VARIANT)
  bool chain_team = false;    // ... with an 8-lane team per environment (small batches: latency-bound otherwise)
  bool chain_single = false;  // limit-only chain: the whole tick runs in k_chain (unless xfrc_applied is in use)
  bool fusable = false;  // joint limits are the only constraint source
  bool use_graph = true;  // replay the tick's kernel sequence as a CUDA graph (B2_NO_GRAPH=1 disables)
  struct GraphEntry { cudaGraphExec_t exec = nullptr; int kernels = 0; };
  std::map<std::pair<unsigned long long, unsigned long long>, GraphEntry> graphs;  // keyed by (tick flags, timestep; exchange buffers)
  int wp = 16, epl = 2;  // (legacy)
  // block records of the constraint pipeline (k_constraint.cuh)
  int block_capw = 0;    // words of efc_blocks per environment
  int rec_max = 0;       // words per thread of k_make_blocks' shared-memory column: parameters + one base direction (J | B)
  int block_npar = 0;    // ... of which header + parameters
  int make_block = 128;  // CTA size of k_make_constraint
  int pgs_lanes = 8;     // lanes per environment in k_pgs_block
  // observation exchange: this GPU's buffer, the peers' mappings (device array of world pointers), slice geometry
  float* obs_buf = nullptr;
  float** obs_peers_dev = nullptr;
  std::vector<float*> obs_peers_host;
  std::vector<bool> obs_peer_ipc;
  int obs_world = 0, obs_rank = 0;
  bool obs_on = false;
  void* h_dev = nullptr; // {double h, float h}: the timestep in device memory (KArgs::hp)
  // sub-batches: windows of the batch that run the pipeline side by side on their own streams (batch.cu: run_tick)
  int nsub = 1;            // windows asked for (b2_set_option "subbatches", B2_SUBBATCH); halved until they are whole tiles.
                           // Default 1: worth 2-3 % before the solver visited environments by last tick's iterations, nothing since
                           // (C3 1.420 / 1.412 / 1.439 / 1.457 ms per tick with 1 / 2 / 4 / 8 windows, C5 3.22 / 3.34 with 1 / 4)
  int sub_min_envs = 2048; // ... of at least this many environments (b2_set_option "subbatch_min")
  std::vector<cudaStream_t> sub_stream;
  std::vector<cudaEvent_t> sub_join;
  cudaEvent_t sub_fork = nullptr;
  int tree_lanes = 1;    // lanes per environment in k_smooth / k_integrate (tree-parallel form; 1: thread per environment)
  int tc_rows = 0;       // tensor-core projection (k_project_tc): rows per environment of the environment-major arrays; 0: off
  int tc_passes = 3;     // 3: 3xTF32 (fp32-level accuracy), 1: plain TF32
  int isl_cap = 0;       // island slots per environment (k_pgs_island: models made of several small trees); 0: k_pgs_block
  int pgs_isl = 8;       // lanes (= islands relaxed side by side) per environment in k_pgs_island
  int isl_stage = 0;     // words of records k_pgs_island stages per environment
  size_t isl_smem = 0;   // k_make_rows: island label columns
  int ld_extra = 0;      // ... with a fourth vector (second product of k_smooth's fused M x pass)
  size_t ld_smem = 0;    // bytes of the shared-memory factor scratch of k_smooth / k_integrate (workspace in HBM), 0: off
  int row_nb = 0;        // k_make_rows' shared-memory row column: base rows it holds (0: off)
  size_t row_smem = 0;
  int solve_rows = 0;    // row-warps per CTA of k_solve_rows (0: rows are solved inside k_make_blocks)
  size_t solve_smem = 0;
  int stage_cap = 0;     // words of records per environment staged in shared memory by the solver
  int pgs_ctas_per_sm = 4;
  int smooth_block = 32;
  size_t smooth_smem = 0, blob_smem = 0;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  void* stage_dev = nullptr;
  size_t stage_bytes = 0;
  long long launches = 0;
  int opt_iterations = 100, opt_disableflags = 0;
  double opt_tolerance = 1e-8;
  // hardware-interface joints
  int nhw = 0;
  int *hw_qadr = nullptr, *hw_dadr = nullptr, *hw_ctl = nullptr;
  float* hw_buf = nullptr;  // [5][nhw][nenv] fp32 staging: vel_cmd, effort_cmd, pos, vel, effort
  float *hw_kp = nullptr, *hw_kd = nullptr;  // [nhw] device-side PD gains (b2_set_pd)
  bool hw_identity = false; // hardware joint j is dof j for every dof (lets k_chain do the exchange itself)
  // buffers the exchange of the tick being launched reads / writes: the hw_buf slots, or device-accessible aliases of the
  // caller's host buffers (pinned / registered memory is read and written in place over PCIe, no staging copies)
  const float* io_in[2] = {nullptr, nullptr};
  float* io_out[3] = {nullptr, nullptr, nullptr};
  std::map<const void*, size_t> registered;  // host ranges pinned through b2_register_host (the caller owns their lifetime)
  // device aliases of the caller buffers seen by the last b2_tick_host (the control loop passes the same ones every tick)
  const void* alias_host[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float* alias_dev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t alias_bytes[5] = {0, 0, 0, 0, 0};   // the cache key is (pointer, exchange bytes)
  // kernel argument block, rebuilt only after a field (re)allocation
  b2::KArgs<float> args_f{};
  b2::KArgs<double> args_d{};
  bool args_valid = false;
  // object slots (b2_set_slots): runtime spawn / destroy as slot activation
  int nslot = 0;
  int *slot_qadr = nullptr, *slot_dadr = nullptr;  // [nslot] device: qpos / dof address of each slot's free joint
  unsigned char* slot_active = nullptr;            // [nslot][nenvp] device
  // per-kernel CUDA-event profiling (b2_profile_begin / b2_profile_end)
  std::vector<cudaEvent_t> prof_ev;  // [max_ticks][B2_NSLOT + 1]
  std::vector<unsigned> prof_mask;   // which boundary events of each tick were recorded
  int prof_max = 0, prof_n = 0;
  bool prof_on = false;
  int prof_tick_open = -1;
};

namespace b2 {
// serial-chain kernels live in their own translation units (chain_f32.cu / chain_f64.cu): they are the slowest to compile
int launch_chain_f32(b2_batch* b, const KArgs<float>& a, int grid);
int launch_chain_f64(b2_batch* b, const KArgs<double>& a, int grid);
bool have_chain_kernel(int n, int precision);
// the whole tick of a limit-only serial chain in one kernel (k_chain.cuh, chain1_f32.cu / chain1_f64.cu)
int launch_chain1_f32(b2_batch* b, const KArgs<float>& a, int grid);
int launch_chain1_f64(b2_batch* b, const KArgs<double>& a, int grid);
// the same tick with an 8-lane team per environment (k_chain_team.cuh): small batches
int launch_chain_team_f32(b2_batch* b, const KArgs<float>& a);
int launch_chain_team_f64(b2_batch* b, const KArgs<double>& a);
}  // namespace b2
