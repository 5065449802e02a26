// Does the FP64 pipe of sm_100a (B200: full-rate DFMA, unlike sm_103a) give BabyBear modmuls that run beside the
// integer multiply pipe?  (a measurement tool, not product code)
//   1. exactness of a 6-instruction FP64 modmul (signed residues in (-p, p)) against 64-bit integer arithmetic
//   2. per-SM throughput: DFMA alone, integer Montgomery product alone, FP64 modmul alone, and both kinds of
//      warp resident on the same SM (even warps integer, odd warps FP64)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr uint32_t P = 0x78000001u, PINV = 0x88000001u;
#define DEV __device__ __forceinline__

DEV uint32_t mont_min(uint32_t a, uint32_t b) {
  uint64_t x = (uint64_t)a * b;
  uint32_t t = (uint32_t)x * PINV;
  uint32_t hu = __umulhi(t, P);
  uint32_t r = (uint32_t)(x >> 32) - hu;
  return min(r, r + P);
}

// a * b mod p for |a|, |b| < 2^32 (integers held in doubles); result an integer in [-p/2 - 1, p/2 + 1] + small slack
DEV double dmodmul(double a, double b) {
  const double dp = (double)P, pinv = 1.0 / (double)P, magic = 6755399441055744.0;  // 1.5 * 2^52
  double h = a * b;
  double l = __fma_rn(a, b, -h);            // exact low part
  double q = __fma_rn(h, pinv, magic) - magic;  // rint(h / p)
  double r = __fma_rn(-q, dp, h);           // exact: |h - q p| < 2^33
  return r + l;
}

__global__ void exact_kernel(const uint32_t* a, const uint32_t* b, int n, unsigned long long* bad, double* maxabs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // signed inputs as the lazy pipeline would have them: a - p*(i&1), b - p*((i>>1)&1)
  double da = (double)a[i] - ((i & 1) ? (double)P : 0.0), db = (double)b[i] - ((i & 2) ? (double)P : 0.0);
  double r = dmodmul(da, db);
  long long ri = (long long)r;
  if ((double)ri != r) atomicAdd(bad, 1ull);
  long long want = (long long)(((unsigned long long)a[i] * b[i]) % P);
  long long got = ((ri % (long long)P) + P) % P;
  if (got != want) atomicAdd(bad, 1ull);
  double m = fabs(r);
  // chain: r^2 * r (values up to |r| ~ p) stays exact too
  double r3 = dmodmul(dmodmul(r, r), r);
  long long w3 = (long long)(((unsigned long long)want * want) % P);
  w3 = (long long)(((unsigned long long)w3 * want) % P);
  long long g3 = (((long long)r3 % (long long)P) + P) % P;
  if (g3 != w3) atomicAdd(bad, 1ull);
  m = fmax(m, fabs(r3));
  if (m > *maxabs) *maxabs = m;  // racy max: indicative only
}

template <int MODE>  // 0 DFMA, 1 integer Montgomery, 2 FP64 modmul, 3 mixed (even warps 1, odd warps 2)
__global__ void __launch_bounds__(256) bench(uint32_t* out, int iters, uint32_t seed) {
  int warp = threadIdx.x >> 5;
  bool fp = MODE == 0 || MODE == 2 || (MODE == 3 && (warp & 1));
  uint32_t acc = 0;
  if (!fp) {
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (seed + threadIdx.x * 977u + i * 131071u + blockIdx.x) % P;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = mont_min(v[i], v[(i + 1) & 7]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= v[i];
  } else {
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (double)((seed + threadIdx.x * 977u + i * 131071u + blockIdx.x) % P);
    if (MODE == 0) {
      double c = 1.0000001, d = 1e-9;
      for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __fma_rn(v[i], c, d);
      }
    } else {
      for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = dmodmul(v[i], v[(i + 1) & 7]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= (uint32_t)(long long)v[i];
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
static double run(const char* name, int ops_per_iter_thread, uint32_t* d_out, int sms, double clock_ghz) {
  int iters = 2000, ctas = sms * 4;
  bench<MODE><<<ctas, 256>>>(d_out, 10, 1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<MODE><<<ctas, 256>>>(d_out, iters, 7);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)ctas * 256 * iters * ops_per_iter_thread;
  double per_clk_sm = ops / (ms * 1e-3) / (clock_ghz * 1e9) / sms;
  printf("{\"probe\": \"%s\", \"ms\": %.3f, \"Gops\": %.1f, \"ops_per_clk_per_sm\": %.2f}\n", name, ms, ops / ms * 1e-6, per_clk_sm);
  return ops / ms * 1e-6;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz * 1e-6;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", prop.name, sms, ghz);
  const int n = 1 << 22;
  uint32_t *a, *b, *d_out;
  unsigned long long* bad;
  double* mx;
  cudaMallocManaged(&a, n * 4);
  cudaMallocManaged(&b, n * 4);
  cudaMallocManaged(&bad, 8);
  cudaMallocManaged(&mx, 8);
  cudaMalloc(&d_out, 64);
  uint64_t s = 88172645463325252ull;
  for (int i = 0; i < n; i++) {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    a[i] = (uint32_t)(s % P);
    b[i] = (uint32_t)((s >> 32) % P);
  }
  a[0] = P - 1, b[0] = P - 1, a[1] = P - 1, b[1] = 1, a[2] = 0, b[2] = P - 1, a[3] = P - 1, b[3] = P - 1;
  *bad = 0, *mx = 0;
  exact_kernel<<<n / 256, 256>>>(a, b, n, bad, mx);
  cudaDeviceSynchronize();
  printf("{\"probe\": \"fp64_modmul_exact\", \"cases\": %d, \"mismatches\": %llu, \"max_abs_over_p\": %.3f}\n", n, *bad,
         *mx / (double)P);
  run<0>("dfma", 8, d_out, sms, ghz);
  double gi = run<1>("mont_int", 8, d_out, sms, ghz);
  double gf = run<2>("modmul_fp64", 8, d_out, sms, ghz);
  double gm = run<3>("mixed_int_fp64", 8, d_out, sms, ghz);
  printf("{\"probe\": \"summary\", \"mixed_over_int\": %.3f, \"mixed_over_sum_of_halves\": %.3f}\n", gm / gi, gm / (0.5 * gi + 0.5 * gf));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("cuda error %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
