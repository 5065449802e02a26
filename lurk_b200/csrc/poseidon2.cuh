// Poseidon2 over BabyBear, register-resident, one permutation per thread.
//
// Replaces (reference file:line):
//   src/core/poseidon.rs:29-39,61-63     PoseidonChipset::hash / execute_simple (W = 24/32/40)
//   src/poseidon/wide/trace.rs:12-82     Poseidon2Cols::populate (wide witness, optional `Wit` sink)
//   src/poseidon/config.rs:104-119       internal layer x_i*diag_i + sum
//   third-party p3-poseidon2             round order and external layer (M4 = circ(2,3,1,1))
//   third-party p3-symmetric             PaddingFreeSponge<16,8,8>, TruncatedPermutation<2,8,16>
// The kernel is INT32-ALU bound: ~13 modmul per input byte at W = 32, so the state
// lives in registers, round constants in __constant__ memory (uniform broadcast reads) and global
// traffic is staged through padded shared memory only to keep it coalesced.
#pragma once
#include "field.cuh"
#include "poseidon2_constants.h"

namespace lsk {

template <int T> struct P2Rounds;
template <> struct P2Rounds<16> { static constexpr int RP = 13; };
template <> struct P2Rounds<24> { static constexpr int RP = 21; };
template <> struct P2Rounds<32> { static constexpr int RP = 30; };
template <> struct P2Rounds<40> { static constexpr int RP = 38; };

template <int T>
struct P2Const {
  uint32_t diag[T];        // plain residues: the internal layer multiplies by them Shoup-style
  uint32_t diag_shoup[T];
  uint32_t full_rc[8 * T];
  uint32_t part_rc[P2Rounds<T>::RP];
};

#if defined(__CUDACC__)
// Montgomery-form constants.  `static`: every translation unit that hashes owns a copy (4.4 KB of
// the 64 KB constant bank) and uploads it with p2_upload_constants() from its init hook, which
// keeps the library free of relocatable device code.
static __constant__ P2Const<16> g_p2c16;
static __constant__ P2Const<24> g_p2c24;
static __constant__ P2Const<32> g_p2c32;
static __constant__ P2Const<40> g_p2c40;

template <int T> LSK_D const P2Const<T>& p2c();
template <> LSK_D const P2Const<16>& p2c<16>() { return g_p2c16; }
template <> LSK_D const P2Const<24>& p2c<24>() { return g_p2c24; }
template <> LSK_D const P2Const<32>& p2c<32>() { return g_p2c32; }
template <> LSK_D const P2Const<40>& p2c<40>() { return g_p2c40; }

struct NoWitness {
  LSK_D void ext_state(int, int, uint32_t) {}
  LSK_D void ext_sbox(int, int, uint32_t) {}
  LSK_D void int_init(int, uint32_t) {}
  LSK_D void int_state0(int, uint32_t) {}
  LSK_D void int_sbox(int, uint32_t) {}
};

// x < 2^38 -> x mod P, using 2^31 == 2^27 - 1 (mod P) twice and one conditional subtraction.
LSK_D uint32_t reduce_small64(uint64_t x) {
  uint64_t v = (x >> 31) * ((1u << 27) - 1u) + (x & 0x7fffffffu);   // < 2^34 + 2^31
  uint32_t w = (uint32_t)(v >> 31) * ((1u << 27) - 1u) + ((uint32_t)v & 0x7fffffffu);  // < 2^30 + 2^31 < 2P
  return w >= P ? w - P : w;
}

template <int T>
LSK_D void p2_external_linear(uint32_t (&s)[T]) {
#pragma unroll
  for (int i = 0; i < T; i += 4) {
    uint32_t x0 = s[i], x1 = s[i + 1], x2 = s[i + 2], x3 = s[i + 3];
    uint32_t t01 = fadd(x0, x1), t23 = fadd(x2, x3);
    uint32_t t0123 = fadd(t01, t23);
    uint32_t t01123 = fadd(t0123, x1), t01233 = fadd(t0123, x3);
    s[i + 3] = fadd(t01233, fdbl(x0));
    s[i + 1] = fadd(t01123, fdbl(x2));
    s[i] = fadd(t01123, t01);
    s[i + 2] = fadd(t01233, t23);
  }
  uint32_t sums[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < T; i++) sums[i & 3] = fadd(sums[i & 3], s[i]);
#pragma unroll
  for (int i = 0; i < T; i++) s[i] = fadd(s[i], sums[i & 3]);
}

// One copy of each round body (external rounds 0-3, the partial rounds, external rounds 4-7 run
// through the same code): ~17 KB of SASS per instantiation, so a kernel with a single call site
// stays inside the 32 KB instruction cache level even when its warps are in different phases.
template <int T, class Wit>
LSK_D void p2_permute(uint32_t (&s)[T], Wit& wit) {
  const P2Const<T>& c = p2c<T>();
  constexpr int RP = P2Rounds<T>::RP;
  p2_external_linear<T>(s);
#pragma unroll 1
  for (int r = 0; r < 8; r++) {
    if (r == 4) {
#pragma unroll 1
      for (int q = 0; q < RP; q++) {
        if (q == 0) {
#pragma unroll
          for (int i = 0; i < T; i++) wit.int_init(i, s[i]);
        } else {
          wit.int_state0(q - 1, s[0]);
        }
        uint32_t x = fadd(s[0], c.part_rc[q]);
        uint32_t x3 = fmul(fsqr(x), x);
        wit.int_sbox(q, x3);
        s[0] = fmul(x, fsqr(x3));
        // sum of the state: each term < P, T <= 40 terms < 2^37 -> one 64-bit accumulate + reduce
        uint64_t acc = 0;
#pragma unroll
        for (int i = 0; i < T; i += 2) acc += (uint32_t)(s[i] + s[i + 1]);  // each term < 2^32
        uint32_t sum = reduce_small64(acc);
#pragma unroll
        for (int i = 0; i < T; i++) s[i] = fadd(fmul_shoup(s[i], c.diag[i], c.diag_shoup[i]), sum);
      }
    }
#pragma unroll
    for (int i = 0; i < T; i++) {
      wit.ext_state(r, i, s[i]);
      uint32_t x = fadd(s[i], c.full_rc[r * T + i]);
      uint32_t x3 = fmul(fsqr(x), x);
      wit.ext_sbox(r, i, x3);
      s[i] = fmul(x, fsqr(x3));
    }
    p2_external_linear<T>(s);
  }
}

template <int T>
LSK_D void p2_permute(uint32_t (&s)[T]) {
  NoWitness w;
  p2_permute<T, NoWitness>(s, w);
}

// TruncatedPermutation<Perm,2,8,16>::compress on two Montgomery-form digests
LSK_D void p2_compress(const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t (&out)[8]) {
  uint32_t s[16];
#pragma unroll
  for (int i = 0; i < 8; i++) s[i] = l[i], s[8 + i] = r[i];
  p2_permute<16>(s);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = s[i];
}

template <int T>
static inline void p2_fill_const(P2Const<T>& h, const uint32_t* diag, const uint32_t* full, const uint32_t* part) {
  for (int i = 0; i < T; i++) h.diag[i] = diag[i] % P, h.diag_shoup[i] = shoup_precompute(diag[i] % P);
  for (int i = 0; i < 8 * T; i++) h.full_rc[i] = to_mont(full[i]);
  for (int i = 0; i < P2Rounds<T>::RP; i++) h.part_rc[i] = to_mont(part[i]);
}
// call once per translation unit (per device) before launching any hashing kernel of that unit
static inline cudaError_t p2_upload_constants() {
  P2Const<16> h16; P2Const<24> h24; P2Const<32> h32; P2Const<40> h40;
  p2_fill_const<16>(h16, P2_DIAG_16, P2_FULL_RC_16, P2_PART_RC_16);
  p2_fill_const<24>(h24, P2_DIAG_24, P2_FULL_RC_24, P2_PART_RC_24);
  p2_fill_const<32>(h32, P2_DIAG_32, P2_FULL_RC_32, P2_PART_RC_32);
  p2_fill_const<40>(h40, P2_DIAG_40, P2_FULL_RC_40, P2_PART_RC_40);
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(g_p2c16, &h16, sizeof h16)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(g_p2c24, &h24, sizeof h24)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(g_p2c32, &h32, sizeof h32)) != cudaSuccess) return e;
  return cudaMemcpyToSymbol(g_p2c40, &h40, sizeof h40);
}
// ---- 16 lanes per permutation (width 16) -----------------------------------------------------------------
// A permutation computed by one thread takes ~10 us end to end (measured: every extra level of a
// single-CTA tree tail costs that much), so the last levels of every tree -- too few nodes to fill
// the machine -- are latency-bound.  There a half-warp shares one permutation, one state word per
// lane: an S-box round is 4 dependent products instead of 64 interleaved ones, the external layer
// is 5 shuffles (circ(2,3,1,1) inside each 4-lane group, then the column sums over the 4 groups),
// and in the partial rounds every lane recomputes the S-box of word 0 (no broadcast) while the sum
// of the other 15 words is carried incrementally (rest' = sum_i d_i s_i + 15 * sum), its butterfly
// running under the next round's S-box chain.  ~5x lower latency for ~1.4x the multiplies.
static __device__ uint32_t g_coop16[10 * 16];  // [r][lane]: full-round constants r = 0..7, diag (plain), diag (Shoup)
struct Coop16 {
  uint32_t rc[8], d, dsh;
};
LSK_D Coop16 coop16_load(unsigned lane) {
  Coop16 k;
#pragma unroll
  for (int r = 0; r < 8; r++) k.rc[r] = g_coop16[r * 16 + (lane & 15)];
  k.d = g_coop16[8 * 16 + (lane & 15)];
  k.dsh = g_coop16[9 * 16 + (lane & 15)];
  return k;
}
LSK_D uint32_t coop16_sum(uint32_t v) {  // sum over the 16 lanes of this half-warp, in every lane
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) v = fadd(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
LSK_D uint32_t coop16_linear(uint32_t x, unsigned lane) {
  // two rounds of three independent shuffles (the other three lanes of the group / the other three
  // groups) rather than four dependent butterfly steps: the layer sits on the latency-critical path
  const unsigned base = lane & ~3u;
  uint32_t a = __shfl_sync(0xffffffffu, x, base | ((lane + 1) & 3u));
  uint32_t b = __shfl_sync(0xffffffffu, x, base | ((lane + 2) & 3u));
  uint32_t c = __shfl_sync(0xffffffffu, x, base | ((lane + 3) & 3u));
  // row j of circ(2,3,1,1): 2 x_j + 3 x_{j+1} + x_{j+2} + x_{j+3}
  uint32_t y = fadd(fadd(fdbl(x), fadd(fdbl(a), a)), fadd(b, c));
  uint32_t y4 = __shfl_xor_sync(0xffffffffu, y, 4);
  uint32_t y8 = __shfl_xor_sync(0xffffffffu, y, 8);
  uint32_t y12 = __shfl_xor_sync(0xffffffffu, y, 12);
  return fadd(fadd(fdbl(y), y4), fadd(y8, y12));
}
LSK_D uint32_t sbox7(uint32_t x) {
  uint32_t x3 = fmul(fsqr(x), x);
  return fmul(x, fsqr(x3));
}
LSK_D uint32_t coop16_permute(uint32_t s, const Coop16& k, unsigned lane) {
  const P2Const<16>& c = p2c<16>();
  const bool first = (lane & 15) == 0;
  s = coop16_linear(s, lane);
#pragma unroll
  for (int r = 0; r < 4; r++) s = coop16_linear(sbox7(fadd(s, k.rc[r])), lane);
  uint32_t s0 = __shfl_sync(0xffffffffu, s, lane & 16u);
  uint32_t rest = coop16_sum(first ? 0u : s);
#pragma unroll 1
  for (int q = 0; q < P2Rounds<16>::RP; q++) {
    uint32_t m = first ? 0u : fmul_shoup(s, k.d, k.dsh);
    uint32_t dsum = coop16_sum(m);
    uint32_t s0n = sbox7(fadd(s0, c.part_rc[q]));
    uint32_t sum = fadd(s0n, rest);
    s0 = fadd(fmul_shoup(s0n, c.diag[0], c.diag_shoup[0]), sum);
    s = first ? s0 : fadd(m, sum);
    uint32_t s2 = fdbl(sum), s4 = fdbl(s2), s8 = fdbl(s4);
    rest = fadd(dsum, fsub(fdbl(s8), sum));
  }
#pragma unroll
  for (int r = 4; r < 8; r++) s = coop16_linear(sbox7(fadd(s, k.rc[r])), lane);
  return s;
}
static inline cudaError_t coop16_upload_constants() {
  uint32_t h[10 * 16];
  for (int i = 0; i < 16; i++) {
    for (int r = 0; r < 8; r++) h[r * 16 + i] = to_mont(P2_FULL_RC_16[r * 16 + i]);
    h[8 * 16 + i] = P2_DIAG_16[i] % P;
    h[9 * 16 + i] = shoup_precompute(P2_DIAG_16[i] % P);
  }
  return cudaMemcpyToSymbol(g_coop16, h, sizeof h);
}
#endif  // __CUDACC__

}  // namespace lsk
