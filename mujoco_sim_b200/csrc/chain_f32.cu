// chain_f32.cu — fp32 instantiations of the register-resident serial-chain smooth kernel (k_smooth.cuh, ChainP<N>).
#include <cuda_runtime.h>

#include "batch_internal.h"
#include "k_smooth.cuh"

namespace b2 {
namespace {
template <typename T, int BLOCK, typename P, int MINB = 1>
int launch(b2_batch* b, const KArgs<T>& a, int grid) {
  static bool attr_set[8] = {false};
  const int dev = b->device & 7;
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(k_smooth<T, BLOCK, P, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
      return set_error("cudaFuncSetAttribute(k_smooth chain) failed");
    attr_set[dev] = true;
  }
  k_smooth<T, BLOCK, P, MINB><<<grid, BLOCK, b->smooth_smem, b->stream>>>(a);
  b->launches++;
  return 0;
}
}  // namespace

int launch_chain_f32(b2_batch* b, const KArgs<float>& a, int grid) {
  // CTA of 32 threads / 230 registers for batches that cannot fill the SMs, 128 threads / 128 registers otherwise
  // (profiles/r01_chain_variants.txt)
  if (b->chain_n == 7) return b->smooth_block == 32 ? launch<float, 32, ChainP<7>>(b, a, grid) : launch<float, 128, ChainP<7>, 4>(b, a, grid);
  if (b->chain_n == 6) return b->smooth_block == 32 ? launch<float, 32, ChainP<6>>(b, a, grid) : launch<float, 128, ChainP<6>, 4>(b, a, grid);
  return set_error("no fp32 chain kernel for this chain length");
}
bool have_chain_kernel(int n, int precision) { return precision == 4 ? (n == 6 || n == 7) : n == 7; }

}  // namespace b2
