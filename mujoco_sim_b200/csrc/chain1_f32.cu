// chain1_f32.cu — fp32 instantiations of the single-kernel serial-chain tick (k_chain.cuh).
#include <cuda_runtime.h>

#include "batch_internal.h"
#include "k_chain.cuh"

namespace b2 {
namespace {
template <typename T, int N, int BLOCK, int MINB>
int launch1(b2_batch* b, const KArgs<T>& a, int grid) {
  k_chain<T, N, BLOCK, MINB><<<grid, BLOCK, b->blob_smem, b->stream>>>(a);
  b->launches++;
  return 0;
}
}  // namespace

int launch_chain1_f32(b2_batch* b, const KArgs<float>& a, int grid) {
  const int v = b->chain_variant;
  if (b->chain_n == 7) {
    if (b->smooth_block == 32) return launch1<float, 7, 32, 1>(b, a, grid);
    return v == 1 ? launch1<float, 7, 128, 2>(b, a, grid) : launch1<float, 7, 128, 4>(b, a, grid);
  }
  if (b->chain_n == 6) return b->smooth_block == 32 ? launch1<float, 6, 32, 1>(b, a, grid) : launch1<float, 6, 128, 4>(b, a, grid);
  return set_error("no fp32 single-kernel chain tick for this chain length");
}

}  // namespace b2
