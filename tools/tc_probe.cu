// Standalone check of k_project_tc against a double-precision product: tools/tc_probe [nenv]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "k_project_tc.cuh"
using namespace b2;
int main(int argc, char** argv) {
  const int nenv = argc > 1 ? atoi(argv[1]) : 64, arows = 128, mt = 1, nv = 49;
  std::vector<float> A((size_t)nenv * arows * 64, 0.f), B((size_t)nenv * 64 * 64, 0.f), D((size_t)nenv * arows * 64, 0.f);
  srand(1);
  for (int e = 0; e < nenv; e++) {
    for (int r = 0; r < 116; r++) for (int k = 0; k < nv; k++) A[((size_t)e * arows + r) * 64 + k] = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
    for (int n = 0; n < nv; n++) for (int k = 0; k <= n; k++) { float v = (rand() / (float)RAND_MAX - 0.5f); B[((size_t)e * 64 + n) * 64 + k] = v; B[((size_t)e * 64 + k) * 64 + n] = v; }
  }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 2 * tc::A_BYTES + 2 * tc::B_BYTES;
  cudaFuncSetAttribute(k_project_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int passes : {1, 3}) {
    cudaMemset(dD, 0, D.size() * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_project_tc<<<std::min(nenv, 296), 128, smem>>>(dA, dB, dD, nullptr, nenv, mt, arows, passes);
    cudaEventRecord(e0);
    k_project_tc<<<std::min(nenv, 296), 128, smem>>>(dA, dB, dD, nullptr, nenv, mt, arows, passes);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int e = 0; e < nenv; e++) for (int r = 0; r < 116; r++) for (int n = 0; n < nv; n++) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += (double)A[((size_t)e * arows + r) * 64 + k] * B[((size_t)e * 64 + n) * 64 + k];
      maxerr = fmax(maxerr, fabs(s - D[((size_t)e * arows + r) * 64 + n])); maxref = fmax(maxref, fabs(s));
    }
    printf("passes %d: %s, %.3f ms for %d envs, max |err| %.3e (max |ref| %.2f, rel %.2e)\n", passes, cudaGetErrorString(err), ms, nenv, maxerr, maxref, maxerr / maxref);
  }
  return 0;
}
