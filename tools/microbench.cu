// Integer-pipe microbenchmarks for BabyBear arithmetic on sm_100a (a measurement tool, not product
// code): per-SM throughput of the multiply instructions and of candidate modmul formulations.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr uint32_t P = 0x78000001u, PINV = 0x88000001u;

#define DEV __device__ __forceinline__

DEV uint32_t mont_wide(uint32_t a, uint32_t b) {  // current: mul.wide + mul.lo + mul.hi, canonical result
  uint64_t x = (uint64_t)a * b;
  uint32_t t = (uint32_t)x * PINV;
  uint32_t hu = __umulhi(t, P);
  uint32_t hx = (uint32_t)(x >> 32);
  uint32_t r = hx - hu;
  return hx < hu ? r + P : r;
}
DEV uint32_t mont_nowide(uint32_t a, uint32_t b) {
  uint32_t lo = a * b, hx = __umulhi(a, b);
  uint32_t t = lo * PINV;
  uint32_t hu = __umulhi(t, P);
  uint32_t r = hx - hu;
  return hx < hu ? r + P : r;
}
DEV uint32_t mont_shift(uint32_t a, uint32_t b) {  // t = lo * (2^31 + 2^27 + 1) by shifts
  uint64_t x = (uint64_t)a * b;
  uint32_t lo = (uint32_t)x;
  uint32_t t = lo + (lo << 27) + (lo << 31);
  uint32_t hu = __umulhi(t, P);
  uint32_t hx = (uint32_t)(x >> 32);
  uint32_t r = hx - hu;
  return hx < hu ? r + P : r;
}
DEV uint32_t mont_min(uint32_t a, uint32_t b) {  // branch-free min() correction
  uint64_t x = (uint64_t)a * b;
  uint32_t t = (uint32_t)x * PINV;
  uint32_t hu = __umulhi(t, P);
  uint32_t r = (uint32_t)(x >> 32) - hu;
  return min(r, r + P);
}
DEV uint32_t mont_lazy(uint32_t a, uint32_t b) {  // result in (0, 2P): no correction (needs a*b < P*2^32)
  uint64_t x = (uint64_t)a * b;
  uint32_t t = (uint32_t)x * PINV;
  uint32_t hu = __umulhi(t, P);
  return (uint32_t)(x >> 32) - hu + P;
}
DEV uint32_t shoup(uint32_t a, uint32_t c, uint32_t cp) {  // cp = floor(c * 2^32 / P); result canonical
  uint32_t q = __umulhi(a, cp);
  uint32_t r = a * c - q * P;
  return min(r, r - P);
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(uint32_t* out, int iters, uint32_t seed) {
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = (seed + threadIdx.x * 977u + i * 131071u + blockIdx.x) % P;
  uint32_t c = seed % P | 1u, cp = (uint32_t)(((uint64_t)c << 32) / P);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) v[i] = mont_wide(v[i], v[(i + 1) & 7]);
      if (MODE == 1) v[i] = mont_nowide(v[i], v[(i + 1) & 7]);
      if (MODE == 2) v[i] = mont_shift(v[i], v[(i + 1) & 7]);
      if (MODE == 3) v[i] = mont_min(v[i], v[(i + 1) & 7]);
      if (MODE == 4) v[i] = mont_lazy(v[i] >> 1, v[(i + 1) & 7] >> 1);
      if (MODE == 5) v[i] = shoup(v[i], c, cp);
      if (MODE == 6) v[i] = v[i] * v[(i + 1) & 7] + c;                                  // IMAD
      if (MODE == 7) v[i] = __umulhi(v[i], v[(i + 1) & 7]) + c;                          // IMAD.HI
      if (MODE == 8) {                                                                    // IMAD.WIDE
        uint64_t x = (uint64_t)v[i] * v[(i + 1) & 7] + c;
        v[i] = (uint32_t)x ^ (uint32_t)(x >> 32);
      }
      if (MODE == 9) { uint32_t s = v[i] + v[(i + 1) & 7]; v[i] = s >= P ? s - P : s; }   // fadd, compare form
      if (MODE == 10) { uint32_t s = v[i] + v[(i + 1) & 7]; v[i] = min(s, s - P); }       // fadd, min form
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
static void run(const char* name, uint32_t* d_out) {
  const int iters = 4096, grid = 148 * 8;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  bench<MODE><<<grid, 256>>>(d_out, 64, 12345);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  bench<MODE><<<grid, 256>>>(d_out, iters, 12345);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  double ops = (double)grid * 256 * iters * 8;
  printf("%-28s %8.3f ms  %8.2f Gop/s  (%.2f ops/clk/SM at 1.9 GHz)\n", name, ms, ops / ms / 1e6,
         ops / ms / 1e6 / 148 / 1.9);
}

int main() {
  uint32_t* d_out;
  cudaMalloc(&d_out, 148 * 8 * 256 * 4);
  run<0>("modmul mont wide (current)", d_out);
  run<1>("modmul mont no-wide", d_out);
  run<2>("modmul mont shift-PINV", d_out);
  run<3>("modmul mont min-correct", d_out);
  run<4>("modmul mont lazy (0,2P)", d_out);
  run<5>("modmul shoup const", d_out);
  run<6>("IMAD lo", d_out);
  run<7>("IMAD.HI", d_out);
  run<8>("IMAD.WIDE", d_out);
  run<9>("fadd compare", d_out);
  run<10>("fadd min", d_out);
  return 0;
}
