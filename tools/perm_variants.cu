// Throughput of the width-16 Poseidon2 permutation (one per thread, the Merkle kernels' form) under different pipe
// placements of the plain additions (field.cuh, LSK_ALU_ADD).  A measurement tool, not product code.
// Build one binary per variant:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I lurk_b200/csrc [-DLSK_ALU_ADD=k] -o perm_variant_k tools/perm_variants.cu
#include <cstdio>
#include <cuda_runtime.h>
#define LSK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#include "poseidon2.cuh"
using namespace lsk;
#ifndef LSK_ALU_ADD
#define LSK_ALU_ADD 0
#endif

// the leaf-sponge shape: absorb 8 fresh words, permute, repeat (loads are synthetic: a counter)
__global__ void __launch_bounds__(128) sponge_chain(uint32_t* out, int n) {
  uint32_t s[16];
  for (int i = 0; i < 16; i++) s[i] = (threadIdx.x * 977u + blockIdx.x * 31u + i) % P;
  for (int k = 0; k < n; k++) {
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = (s[i] ^ (uint32_t)k) & 0x3fffffffu;   // stays below P
    p2_permute<16>(s);
  }
  uint32_t acc = 0;
  for (int i = 0; i < 8; i++) acc ^= s[i];
  out[blockIdx.x * 128 + threadIdx.x] = acc;
}

int main() {
  LSK_CUDA(p2_upload_constants());
  cudaDeviceProp prop;
  LSK_CUDA(cudaGetDeviceProperties(&prop, 0));
  const int ctas = prop.multiProcessorCount * 16, n = 64;
  uint32_t* d_out;
  LSK_CUDA(cudaMalloc(&d_out, (size_t)ctas * 128 * 4));
  sponge_chain<<<ctas, 128>>>(d_out, 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e9f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(a);
    sponge_chain<<<ctas, 128>>>(d_out, n);
    cudaEventRecord(b);
    LSK_CUDA(cudaEventSynchronize(b));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  uint32_t h[4];
  LSK_CUDA(cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost));
  const double perms = (double)ctas * 128 * n;
  printf("{\"probe\": \"poseidon2_w16_sponge_chain\", \"alu_add_mask\": %d, \"ms\": %.4f, \"G_perm_per_s\": %.4f, \"checksum\": %u}\n",
         LSK_ALU_ADD, best, perms / best / 1e6, h[0] ^ h[1] ^ h[2] ^ h[3]);
  return 0;
}
