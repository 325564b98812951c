// chain1_f64.cu — fp64 instantiations of the single-kernel serial-chain tick (k_chain.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "batch_internal.h"
#include "k_chain.cuh"
#include "k_chain_team.cuh"

namespace b2 {
namespace {
template <typename T, int N, int BLOCK, int MINB>
int launch1(b2_batch* b, const KArgs<T>& a, int grid) {
  k_chain<T, N, BLOCK, MINB><<<grid, BLOCK, b->blob_smem, b->stream>>>(a);
  b->launches++;
  return 0;
}
template <typename T, int N, int BLOCK>
int launch_team_b(b2_batch* b, const KArgs<T>& a) {
  constexpr int EPB = BLOCK / 8;
  const int ntiles = b->nenvp / EPB;
  const int grid = std::max(1, std::min(ntiles, b->nsm * (1024 / BLOCK)));
  k_chain_team<T, N, BLOCK><<<grid, BLOCK, b->blob_smem + (size_t)EPB * (64 * sizeof(T) + 40 * sizeof(float)), b->stream>>>(a);
  b->launches++;
  return 0;
}
template <typename T, int N>
int launch_team(b2_batch* b, const KArgs<T>& a) {
  // 32 environments per CTA: the hardware-interface exchange moves 128-byte runs per joint (PCIe when zero-copy)
  static const int blk = getenv("B2_TEAM_BLOCK") ? atoi(getenv("B2_TEAM_BLOCK")) : 256;
  return blk == 128 ? launch_team_b<T, N, 128>(b, a) : launch_team_b<T, N, 256>(b, a);
}
}  // namespace

int launch_chain1_f64(b2_batch* b, const KArgs<double>& a, int grid) {
  if (b->chain_n == 7) return launch1<double, 7, 32, 1>(b, a, grid);
  return set_error("no fp64 single-kernel chain tick for this chain length");
}

// small batches: an 8-lane team per environment (k_chain_team.cuh)
int launch_chain_team_f64(b2_batch* b, const KArgs<double>& a) {
  if (b->chain_n == 7) return launch_team<double, 7>(b, a);
  if (b->chain_n == 6) return launch_team<double, 6>(b, a);
  return set_error("no team chain kernel for this chain length");
}

}  // namespace b2
