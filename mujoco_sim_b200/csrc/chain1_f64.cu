// chain1_f64.cu — fp64 instantiations of the single-kernel serial-chain tick (k_chain.cuh).
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

int launch_chain1_f64(b2_batch* b, const KArgs<double>& a, int grid) {
  if (b->chain_n == 7) return launch1<double, 7, 32, 1>(b, a, grid);
  return set_error("no fp64 single-kernel chain tick for this chain length");
}

}  // namespace b2
