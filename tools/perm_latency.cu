// Latency of one width-16 Poseidon2 permutation on sm_100a (a measurement tool, not product code):
// a single warp runs a dependent chain of permutations and reads the SM clock.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I lurk_b200/csrc -o tools/perm_latency tools/perm_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define LSK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#include "poseidon2.cuh"
using namespace lsk;

__global__ void coop_chain(uint32_t* out, long long* cycles, int n) {
  unsigned lane = threadIdx.x & 31;
  Coop16 k = coop16_load(lane);
  uint32_t s = lane * 7919u + 1;
  s = coop16_permute(s, k, lane);  // warm the instruction cache
  long long t0 = clock64();
  for (int i = 0; i < n; i++) s = coop16_permute(s, k, lane);
  long long t1 = clock64();
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[0] = (t1 - t0) / n;
}
__global__ void thread_chain(uint32_t* out, long long* cycles, int n) {
  uint32_t s[16];
  for (int i = 0; i < 16; i++) s[i] = threadIdx.x * 977u + i;
  p2_permute<16>(s);
  long long t0 = clock64();
  for (int i = 0; i < n; i++) p2_permute<16>(s);
  long long t1 = clock64();
  out[threadIdx.x] = s[0];
  if (threadIdx.x == 0) cycles[1] = (t1 - t0) / n;
}
__global__ void cold_thread(uint32_t* out, long long* cycles) {  // first execution: instruction fetch included
  uint32_t s[16];
  for (int i = 0; i < 16; i++) s[i] = threadIdx.x * 977u + i;
  long long t0 = clock64();
  p2_permute<16>(s);
  long long t1 = clock64();
  out[threadIdx.x] = s[0];
  if (threadIdx.x == 0) cycles[2] = t1 - t0;
}
__global__ void cold_coop(uint32_t* out, long long* cycles) {
  unsigned lane = threadIdx.x & 31;
  Coop16 k = coop16_load(lane);
  uint32_t s = lane * 7919u + 1;
  long long t0 = clock64();
  s = coop16_permute(s, k, lane);
  long long t1 = clock64();
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[3] = t1 - t0;
}
__global__ void shfl_chain(uint32_t* out, long long* cycles, int n) {
  uint32_t v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1;
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cycles[4] = (t1 - t0) * 100 / n;
}
__global__ void mul_chain(uint32_t* out, long long* cycles, int n) {
  uint32_t v = threadIdx.x + 3;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) v = fmul(v, v);
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cycles[5] = (t1 - t0) * 100 / n;
}
__global__ void add_chain(uint32_t* out, long long* cycles, int n) {
  uint32_t v = threadIdx.x + 3, w = out[0] % P;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) v = fadd(v, w);
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cycles[6] = (t1 - t0) * 100 / n;
}

int main() {
  LSK_CUDA(p2_upload_constants());
  LSK_CUDA(coop16_upload_constants());
  uint32_t* out;
  long long *cyc, h[8] = {0};
  LSK_CUDA(cudaMalloc(&out, 4096));
  LSK_CUDA(cudaMemset(out, 0, 4096));
  LSK_CUDA(cudaMalloc(&cyc, 64));
  cold_thread<<<1, 32>>>(out, cyc);
  cold_coop<<<1, 32>>>(out, cyc);
  coop_chain<<<1, 32>>>(out, cyc, 200);
  thread_chain<<<1, 32>>>(out, cyc, 50);
  shfl_chain<<<1, 32>>>(out, cyc, 1000);
  mul_chain<<<1, 32>>>(out, cyc, 1000);
  add_chain<<<1, 32>>>(out, cyc, 1000);
  LSK_CUDA(cudaDeviceSynchronize());
  LSK_CUDA(cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost));
  printf("cooperative (16 lanes) permutation: %lld cycles warm, %lld cold\n", h[0], h[3]);
  printf("one-thread permutation:             %lld cycles warm, %lld cold\n", h[1], h[2]);
  printf("dependent shfl+add: %.2f cycles; dependent fmul: %.2f cycles; dependent fadd: %.2f cycles\n", h[4] / 100.0,
         h[5] / 100.0, h[6] / 100.0);
  return 0;
}
