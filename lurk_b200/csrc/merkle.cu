// Merkle commitment kernels (see merkle.cuh for the contract and the reference map).
//
// One thread owns one sponge / compression: the 16-word Poseidon2 state stays in registers and a
// leaf absorbs 8 columns at a time.  Matrices are column-major, so for a fixed column consecutive
// threads (rows) read consecutive words: leaf hashing streams every LDE cell exactly once
// (8*N*w bytes for a blow-up-2 commit) and writes 32 B per row.  The stage is bound by the
// integer pipes (one width-16 permutation ~ 750 modmuls per 32 B absorbed), not by HBM.
#include <algorithm>
#include <cstdlib>

#include "fri.cuh"
#include "merkle.cuh"

#include "comm.cuh"
#include "poseidon2.cuh"

namespace lsk {


LSK_D void load_digest(const uint32_t* p, uint32_t* d) {
  uint4 a = reinterpret_cast<const uint4*>(p)[0], b = reinterpret_cast<const uint4*>(p)[1];
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
  d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
LSK_D void store_digest(uint32_t* p, const uint32_t* d) {
  reinterpret_cast<uint4*>(p)[0] = make_uint4(d[0], d[1], d[2], d[3]);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(d[4], d[5], d[6], d[7]);
}

// The one hashing routine of the tree, written around a SINGLE permutation call site:
//   sponge over row r of a height group (ncols columns given as base pointers), or over the 8
//   words at `packed_row` (FRI layers), giving the row digest;
//   if `children`: node = compress(children[0], children[1]); with a row digest present the result
//   is compress(node, row digest) (mixed-height injection).
// Result in st[0..8).
//   `rowdig` (tail levels, see tail_rows_kernel): the row digest was computed beforehand -- no sponge steps here.
LSK_D void merkle_digest(const uint32_t* const* __restrict__ cols, uint32_t ncols, size_t r,
                         const uint32_t* __restrict__ packed_row, const uint32_t* __restrict__ children,
                         uint32_t (&st)[16], const uint32_t* __restrict__ rowdig = nullptr) {
  if (rowdig) {
    if (!children) {
      load_digest(rowdig, st);
      return;
    }
    load_digest(children, st);
    load_digest(children + 8, st + 8);
    p2_permute<16>(st);
    load_digest(rowdig, st + 8);
    p2_permute<16>(st);
    return;
  }
  const uint32_t nabs = packed_row ? 1u : (ncols + 7) / 8;
  const uint32_t total = nabs + (children ? (nabs ? 2u : 1u) : 0u);
  uint32_t rd[8];
#pragma unroll
  for (int i = 0; i < 16; i++) st[i] = 0;
#pragma unroll 1
  for (uint32_t step = 0; step < total; step++) {
    if (step < nabs) {
      if (packed_row) {
        load_digest(packed_row, st);
      } else {
        const uint32_t k = step * 8;
#pragma unroll
        for (int c = 0; c < 8; c++)
          if (k + c < ncols) st[c] = cols[k + c][r];
      }
    } else if (step == nabs) {
#pragma unroll
      for (int i = 0; i < 8; i++) rd[i] = st[i];
      load_digest(children, st);
      load_digest(children + 8, st + 8);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) st[8 + i] = rd[i];
    }
    p2_permute<16>(st);
  }
}

// merkle_digest with the state spread over a half-warp: the lane's state word after the last
// permutation (lanes 0..7 of the half hold the digest).  All 32 lanes of the warp must call it with
// the same ncols / packed / children-ness.
LSK_D uint32_t merkle_digest_coop(const uint32_t* const* __restrict__ cols, uint32_t ncols, size_t r,
                                  const uint32_t* __restrict__ packed_row, const uint32_t* __restrict__ children,
                                  const Coop16& k, unsigned lane, const uint32_t* __restrict__ rowdig = nullptr) {
  const unsigned g = lane & 15;
  if (rowdig) {   // the row digest was computed beforehand (tail_rows_kernel)
    if (!children) return g < 8 ? rowdig[g] : 0u;
    uint32_t s = coop16_permute(children[g], k, lane);
    if (g >= 8) s = rowdig[g - 8];
    return coop16_permute(s, k, lane);
  }
  const uint32_t nabs = packed_row ? 1u : (ncols + 7) / 8;
  const uint32_t total = nabs + (children ? (nabs ? 2u : 1u) : 0u);
  uint32_t s = 0, rd = 0;
#pragma unroll 1
  for (uint32_t step = 0; step < total; step++) {
    if (step < nabs) {
      if (packed_row) {
        if (g < 8) s = packed_row[g];
      } else if (g < 8 && step * 8 + g < ncols) {
        s = cols[step * 8 + g][r];
      }
    } else if (step == nabs) {
      rd = s;
      s = children[g];
    } else {
      uint32_t t = __shfl_sync(0xffffffffu, rd, (lane & 16u) | (g & 7u));
      if (g >= 8) s = t;
    }
    s = coop16_permute(s, k, lane);
  }
  return s;
}

// One launch builds up to ST_LEVELS + 1 layers: a CTA takes 2^ST_LEVELS consecutive nodes of layer l0 (in
// the leaf modes it first hashes them from the matrices) and reduces them to one node of layer
// l0 + ST_LEVELS; every level's digests go to the tree in global memory (the query openings need them)
// and to the next level through shared memory.  A 2^21-leaf commitment is 7 flat launches + 2 of
// these, and the sequential tail of a tree costs one launch.
constexpr int ST_THREADS = 512;
constexpr unsigned ST_LEVELS = 9;  // a CTA reduces 2^ST_LEVELS inputs to one node
struct SubtreeArgs {
  uint32_t* nodes;
  const uint32_t* const* cols;
  const uint32_t* fri_vals;
  uint32_t l0, levels, log_in;           // input layer, layers produced, log2(inputs per CTA)
  uint32_t off[ST_LEVELS + 1];           // digest offsets of layers l0 .. l0 + levels
  uint32_t inj_first[ST_LEVELS + 1];     // column-list range of the height group of layer l0 + k
  uint32_t inj_ncols[ST_LEVELS + 1];     // (k = 0: the leaf group in LEAF mode)
  const uint32_t* rowdig;                // digests of the injected rows, laid out like `nodes` (tail_rows_kernel), or null
  FriChallenge fc;                       // set on the launch that produces a FRI layer's root
};
enum { ST_NODES = 0, ST_LEAVES = 1, ST_FRI_LEAVES = 2 };

// Levels (MODE == ST_NODES ? 1 : 0) .. levels of the subtree whose 2^log_in inputs start at node in0 of
// the input layer.  Measured on B200 (tools/perm_latency.cu): one thread computes a permutation in
// 18.7 k cycles, a half-warp in 5.2 k -- but at ~11x the instructions, so on the single SM a CTA owns
// the 16-lane form only wins while a level has <= COOP_MAX nodes; wider levels use one thread per
// node.  A level's digests go to the tree and, through shared memory (`sdig`, 2 x 512 digests), to
// the next level.
constexpr uint32_t COOP_MAX = 64;
template <int MODE>
LSK_D void subtree_build(uint32_t* __restrict__ nodes, const uint32_t* const* __restrict__ cols,
                         const uint32_t* __restrict__ fri_vals, uint32_t levels, uint32_t log_in,
                         const uint32_t* __restrict__ off, const uint32_t* __restrict__ inj_first,
                         const uint32_t* __restrict__ inj_ncols, size_t in0, const Coop16& ck, uint32_t* sdig,
                         const uint32_t* __restrict__ rowdig = nullptr) {
  const uint32_t nin = 1u << log_in;
  const unsigned lane = threadIdx.x & 31;
  const uint32_t k0 = MODE == ST_NODES ? 1u : 0u;
#pragma unroll 1
  for (uint32_t k = k0; k <= levels; k++) {
    const uint32_t cnt = nin >> k;
    uint32_t* layer = nodes + (size_t)off[k] * 8;
    const uint32_t* prev = k ? nodes + (size_t)off[k - 1] * 8 : nullptr;
    const bool packed = MODE == ST_FRI_LEAVES && k == 0;
    const uint32_t ncols = packed || !inj_ncols ? 0u : inj_ncols[k];
    const uint32_t* const* gcols = inj_first ? cols + inj_first[k] : cols;
    const uint32_t* rd_layer = rowdig && ncols ? rowdig + (size_t)off[k] * 8 : nullptr;
    uint32_t* scur = sdig + (k & 1) * (512 * 8);
    const uint32_t* sprev = sdig + ((k - 1) & 1) * (512 * 8);
    if (cnt <= COOP_MAX) {
      // whole warps loop together (the permutation shuffles): slots past cnt recompute the last node
#pragma unroll 1
      for (uint32_t t0 = (threadIdx.x >> 5) * 2; t0 < cnt; t0 += ST_THREADS / 16) {
        const uint32_t t = t0 + (lane >> 4), tc = t < cnt ? t : cnt - 1;
        const size_t i = (in0 >> k) + tc;
        const uint32_t* children = !k ? nullptr : (k > k0 ? sprev + 16 * tc : prev + 16 * i);
        uint32_t s = merkle_digest_coop(gcols, ncols, i, packed ? fri_vals + i * 8 : nullptr, children, ck, lane,
                                        rd_layer ? rd_layer + i * 8 : nullptr);
        if (t < cnt && (lane & 15) < 8) {
          layer[i * 8 + (lane & 15)] = s;
          scur[t * 8 + (lane & 15)] = s;
        }
      }
    } else {
#pragma unroll 1
      for (uint32_t t = threadIdx.x; t < cnt; t += ST_THREADS) {
        const size_t i = (in0 >> k) + t;
        const uint32_t* children = !k ? nullptr : (k > k0 ? sprev + 16 * t : prev + 16 * i);
        uint32_t st[16];
        merkle_digest(gcols, ncols, i, packed ? fri_vals + i * 8 : nullptr, children, st, rd_layer ? rd_layer + i * 8 : nullptr);
        store_digest(layer + i * 8, st);
        store_digest(scur + t * 8, st);
      }
    }
    __syncthreads();
  }
}

// The FRI commit phase's transcript step, on the device (one warp): observe the layer's root, sample
// the folding challenge.  DuplexChallenger<.., 16, 8> with an empty input buffer (it always is here:
// the previous operation was a sample): state[0..8) = root, permute, beta = the last four of the
// first eight output words, popped from the back (challenger.h::sample_ext).
LSK_D void fri_challenge(const FriChallenge& fc, const uint32_t* __restrict__ root, const Coop16& ck) {
  const unsigned lane = threadIdx.x & 31, g = lane & 15;
  uint32_t s = g < 8 ? root[g] : fc.chal[g];
  s = coop16_permute(s, ck, lane);
  if (lane < 16) {
    fc.chal[g] = s;
    if (g >= 4 && g < 8) fc.beta_out->c[7 - g] = s;
    if (g < 8) fc.root_out[g] = root[g];
  }
}

template <int MODE>
__global__ void __launch_bounds__(ST_THREADS) merkle_subtree_kernel(SubtreeArgs a) {
  const Coop16 ck = coop16_load(threadIdx.x & 31);
  __shared__ __align__(16) uint32_t sdig[2 * 512 * 8];
  subtree_build<MODE>(a.nodes, a.cols, a.fri_vals, a.levels, a.log_in, a.off, a.inj_first, a.inj_ncols,
                      (size_t)blockIdx.x << a.log_in, ck, sdig, a.rowdig);
  if (a.fc.chal && blockIdx.x == 0 && threadIdx.x < 32) fri_challenge(a.fc, a.nodes + (size_t)a.off[a.levels] * 8, ck);
}

// The small end of the FRI commit phase in ONE launch of one CTA: for every remaining layer, hash
// the leaves (row i = values 2i, 2i+1), build the tree, run the transcript step, fold into the next
// layer's vector.  Replaces ~2 launches + a device->host round trip per layer.
struct FriTailArgs {
  uint32_t nlayers, first_lfh;            // layers of 2^first_lfh, 2^(first_lfh-1), ..., 2 rows
  Ext* vals[FRI_TAIL_LOG + 2];            // vals[j]: input of layer j; vals[nlayers]: the final 2 values
  uint32_t* nodes[FRI_TAIL_LOG + 1];
  const Ext* add[FRI_TAIL_LOG + 1];       // reduced openings joining the folded vector of layer j (or null)
  uint32_t ginv[FRI_TAIL_LOG + 1];        // 1 / w_{2^(lfh+1)}
  uint32_t one_half;
  FriChallenge fc;                        // beta_out / root_out: slot of the first tail layer
};
__global__ void __launch_bounds__(ST_THREADS) fri_tail_kernel(FriTailArgs a) {
  const Coop16 ck = coop16_load(threadIdx.x & 31);
  __shared__ uint32_t off[ST_LEVELS + 1];
  __shared__ __align__(16) uint32_t sdig[2 * 512 * 8];
#pragma unroll 1
  for (uint32_t j = 0; j < a.nlayers; j++) {
    const uint32_t lfh = a.first_lfh - j;
    if (threadIdx.x <= lfh) off[threadIdx.x] = (2u << lfh) - (2u << (lfh - threadIdx.x));
    __syncthreads();
    subtree_build<ST_FRI_LEAVES>(a.nodes[j], nullptr, (const uint32_t*)a.vals[j], lfh, lfh, off, nullptr, nullptr, 0, ck,
                                 sdig);
    FriChallenge fc{a.fc.chal, a.fc.beta_out + j, a.fc.root_out + 8 * j};
    if (threadIdx.x < 32) fri_challenge(fc, a.nodes[j] + (size_t)off[lfh] * 8, ck);
    __syncthreads();
    const Ext beta = *fc.beta_out;
    const Ext half_beta = emul_base(beta, a.one_half);
    for (uint32_t i = threadIdx.x; i < (1u << lfh); i += ST_THREADS) {
      Ext v = fri_fold_pair(a.vals[j][2 * i], a.vals[j][2 * i + 1], half_beta, a.one_half, a.ginv[j], i, lfh);
      if (a.add[j]) v = eadd(v, a.add[j][i]);
      a.vals[j + 1][i] = v;
    }
    __syncthreads();
  }
}

// Row digests of everything injected into the TAIL of a tree, in one launch.  A Lurk shard has wide chips with a handful
// of rows (hash3/4/5: 493-815 columns): inside the tail such a row is a chain of ~100 dependent permutations, and the
// levels were walked one after another (1.4 ms for the main commitment of a Lurk fib shard, 1.0 of it one 128-node
// level on a single SM).  Here every injected row of every tail level is a half-warp job (16 lanes per permutation)
// spread over the machine, so the chains of different levels and rows run side by side; the tail then only compresses.
struct TailRowJob { uint32_t first_col, ncols, nrows, off; };   // column-list range, rows of the level, digest offset
constexpr int TAIL_ROWS_THREADS = 128;   // 8 half-warps per CTA: many CTAs, few chains per SM
__global__ void __launch_bounds__(TAIL_ROWS_THREADS) tail_rows_kernel(const uint32_t* const* __restrict__ cols,
                                                                      const TailRowJob* __restrict__ jobs, uint32_t njobs,
                                                                      uint32_t* __restrict__ rowdig) {
  const unsigned lane = threadIdx.x & 31;
  const Coop16 ck = coop16_load(lane);
  // global warp index -> (job, pair of rows): both half-warps of a warp work on the SAME job (the permutation's shuffles
  // are warp-wide, so the two halves must run the same number of permutations); jobs are few, a linear walk is cheap.
  // A half-warp past the end of its job recomputes the job's last row and does not store.
  uint32_t w = (blockIdx.x * TAIL_ROWS_THREADS + threadIdx.x) >> 5;
  uint32_t j = 0;
  while (j + 1 < njobs && w >= (jobs[j].nrows + 1) / 2) w -= (jobs[j].nrows + 1) / 2, j++;
  const TailRowJob job = jobs[j];
  const uint32_t hw = 2 * w + (lane >> 4);
  const bool live = hw < job.nrows;
  const uint32_t row = live ? hw : job.nrows - 1;
  uint32_t s = merkle_digest_coop(cols + job.first_col, job.ncols, row, nullptr, nullptr, ck, lane);
  if (live && (lane & 15) < 8) rowdig[((size_t)job.off + row) * 8 + (lane & 15)] = s;
}

// Throughput-bound part of a tree: one thread per leaf / per node, one launch per layer.
constexpr int MK_BLOCK = 128;
constexpr unsigned TAIL_LOG = 14;  // layers with <= 2^TAIL_LOG nodes are latency-bound: finished by subtree launches
template <int MODE>
__global__ void __launch_bounds__(MK_BLOCK) merkle_flat_kernel(const uint32_t* const* __restrict__ cols, uint32_t ncols,
                                                                const uint32_t* __restrict__ fri_vals,
                                                                const uint32_t* __restrict__ prev, size_t node0,
                                                                size_t n, uint32_t* __restrict__ out) {
  // nodes [node0, n) of the layer (node0 != 0: this rank's range of a tree built by a device group)
  size_t i = node0 + (size_t)blockIdx.x * MK_BLOCK + threadIdx.x;
  if (i >= n) return;
  uint32_t st[16];
  merkle_digest(cols, MODE == ST_FRI_LEAVES ? 0u : ncols, i, MODE == ST_FRI_LEAVES ? fri_vals + i * 8 : nullptr,
                MODE == ST_NODES ? prev + 16 * i : nullptr, st);
  store_digest(out + i * 8, st);
}

void init_merkle_kernels() {
  LSK_CUDA(p2_upload_constants());
  LSK_CUDA(coop16_upload_constants());
}

static void layout_layers(lsk_ctx* ctx, MerkleTreeDev& tree) {
  tree.layer_off.resize(tree.log_max + 1);
  size_t off = 0;
  for (unsigned i = 0; i <= tree.log_max; i++) {
    tree.layer_off[i] = off;
    off += (size_t)1 << (tree.log_max - i);
  }
  if (off > 0xffffffffull) fail(LSK_ERR_ARG, "Merkle tree too large");
  tree.nodes.alloc(ctx, off * 8);
}

// inputs per CTA of a tail launch (log2).  ST_LEVELS = 9 (512 nodes per CTA) puts the 256- and 128-node levels of every
// subtree on ONE SM in thread mode (9.5 us each); a smaller split spreads those levels over more SMs and keeps every
// level cooperative.  LSK_TAIL_SPLIT overrides (tuning).
static unsigned tail_log() {   // layers with <= 2^tail_log nodes are finished by subtree launches (LSK_TAIL_LOG: tuning)
  static const unsigned v = [] {
    const char* e = getenv("LSK_TAIL_LOG");
    long x = e ? atol(e) : (long)TAIL_LOG;
    return (unsigned)std::min<long>(std::max<long>(x, 10), 18);
  }();
  return v;
}
static unsigned tail_split() {
  static const unsigned v = [] {
    const char* e = getenv("LSK_TAIL_SPLIT");
    long x = e ? atol(e) : 8;   // measured (fib 2^20, all tails of one prove): 9 -> 1.885 ms, 8 -> 1.810, 7 -> 1.920, 6 -> 1.973, 5 -> 2.234
    return (unsigned)std::min<long>(std::max<long>(x, 3), ST_LEVELS);
  }();
  return v;
}

// groups: per layer l (0..log_max) the column-list range injected there (layer 0 = leaves)
static void build_tree(lsk_ctx* ctx, MerkleTreeDev& tree, const std::vector<std::pair<uint32_t, uint32_t>>& groups,
                       int leaf_mode, const uint32_t* fri_vals, FriChallenge fc = {}) {
  cudaStream_t st = ctx->stream;
  unsigned l0 = 0;
  bool first = true;
  // width-16 permutations of a layer: leaves absorb 8 words per permutation, a node compresses once and absorbs
  // (+ compresses) what is injected at its layer
  auto layer_perms = [&](unsigned l) {
    const double n = (double)((size_t)1 << (tree.log_max - l));
    const double inj = groups[l].second ? (double)((groups[l].second + 7) / 8) + (l ? 1.0 : 0.0) : 0.0;
    if (l == 0) return n * (leaf_mode == ST_FRI_LEAVES ? 1.0 : inj);
    return n * (1.0 + inj);
  };
  auto layer_bytes = [&](unsigned l) {
    return (double)((size_t)1 << (tree.log_max - l)) * (32.0 + 4.0 * groups[l].second + (l ? 64.0 : 0.0) +
                                                        (l == 0 && leaf_mode == ST_FRI_LEAVES ? 32.0 : 0.0));
  };
  // big layers: one thread per digest, one launch per layer.  In a device group every rank builds the layers
  // over its own range of nodes (the children of a range are the same range one layer down, so nothing is
  // exchanged on the way up); the last of these layers is all-gathered for the tail, the ones below stay
  // distributed (tree.dist_layers) and are read through their owner when queries are answered.
  const bool split = comm_splits(ctx, tree.log_max);
  while (tree.log_max - l0 > tail_log() || (first && tree.log_max > ST_LEVELS)) {
    size_t n = (size_t)1 << (tree.log_max - l0), b = 0, e = n;
    if (split) comm_range(ctx, n, &b, &e);
    uint32_t* out = tree.nodes.p + tree.layer_off[l0] * 8;
    const uint32_t* const* cols = tree.colptrs.p + groups[l0].first;
    unsigned grid = div_up(e - b, MK_BLOCK);
    const double share = (double)(e - b) / (double)n;
    if (first) {
      ProfScope ps(ctx, leaf_mode == ST_LEAVES ? PT_MERKLE_LEAF : PT_FRI_LEAF, 1, layer_bytes(0) * share,
                   layer_perms(0) * share);
      if (leaf_mode == ST_LEAVES)
        merkle_flat_kernel<ST_LEAVES><<<grid, MK_BLOCK, 0, st>>>(cols, groups[0].second, nullptr, nullptr, b, e, out);
      else
        merkle_flat_kernel<ST_FRI_LEAVES><<<grid, MK_BLOCK, 0, st>>>(cols, 0, fri_vals, nullptr, b, e, out);
      first = false;
    } else {
      ProfScope ps(ctx, PT_MERKLE_LAYER, 1, layer_bytes(l0) * share, layer_perms(l0) * share);
      merkle_flat_kernel<ST_NODES><<<grid, MK_BLOCK, 0, st>>>(cols, groups[l0].second, nullptr,
                                                             tree.nodes.p + tree.layer_off[l0 - 1] * 8, b, e, out);
    }
    l0++;
  }
  if (!first) l0--;  // l0 = last layer already built
  tree.dist_layers = 0;
  if (split && !first) {
    LSK_CUDA(cudaGetLastError());
    size_t n = (size_t)1 << (tree.log_max - l0);
    comm_allgather(ctx, tree.nodes.p + tree.layer_off[l0] * 8, n / ctx->comm->size * 32);
    tree.dist_layers = l0;
  }
  // row digests of everything injected into the tail, all levels at once (tail_rows_kernel)
  DevBuf<uint32_t> rowdig;
  uint32_t* rowdig_biased = nullptr;
  DevBuf<TailRowJob> d_jobs;
  if (leaf_mode == ST_LEAVES) {
    std::vector<TailRowJob> jobs;
    uint32_t nwarps = 0, longest = 0;
    const unsigned lfirst = first ? 0 : l0 + 1;
    // the buffer covers the tail layers only; the pointer handed to the kernels is biased so that layer offsets of the
    // whole tree index it
    const size_t base = lfirst <= tree.log_max ? tree.layer_off[lfirst] : 0;
    for (unsigned l = lfirst; l <= tree.log_max; l++) {
      if (!groups[l].second) continue;
      const uint32_t nrows = 1u << (tree.log_max - l);
      jobs.push_back(TailRowJob{groups[l].first, groups[l].second, nrows, (uint32_t)tree.layer_off[l]});
      nwarps += (nrows + 1) / 2;
      longest = std::max(longest, groups[l].second);
    }
    if (!jobs.empty() && longest > 16) {   // worth a launch once some row is a chain of more than two permutations
      rowdig.alloc(ctx, tree.nodes.n - base * 8);
      rowdig_biased = rowdig.p - base * 8;
      d_jobs.alloc(ctx, jobs.size());
      d_jobs.upload(jobs.data(), jobs.size());
      double perms = 0, bytes = 0;
      for (auto& j : jobs) perms += (double)j.nrows * ((j.ncols + 7) / 8), bytes += (double)j.nrows * (4.0 * j.ncols + 32.0);
      ProfScope ps(ctx, PT_MERKLE_TOP, 1, bytes);
      tail_rows_kernel<<<div_up((size_t)nwarps * 32, TAIL_ROWS_THREADS), TAIL_ROWS_THREADS, 0, st>>>(tree.colptrs.p, d_jobs.p,
                                                                                                   (uint32_t)jobs.size(), rowdig_biased);
    }
  }
  // latency-bound tail (or a whole small tree): subtree launches
  while (first || l0 < tree.log_max) {
    unsigned log_nodes = tree.log_max - l0;  // log2(nodes of layer l0)
    unsigned log_in = std::min(log_nodes, tail_split());
    SubtreeArgs a{};
    a.nodes = tree.nodes.p;
    a.cols = tree.colptrs.p;
    a.fri_vals = fri_vals;
    a.rowdig = rowdig_biased;
    a.l0 = l0;
    a.levels = log_in;
    a.log_in = log_in;
    double bytes = 0;
    for (unsigned k = 0; k <= a.levels; k++) {
      a.off[k] = (uint32_t)tree.layer_off[l0 + k];
      a.inj_first[k] = groups[l0 + k].first;
      a.inj_ncols[k] = groups[l0 + k].second;
      if (k > 0 || first) bytes += layer_bytes(l0 + k);
    }
    unsigned grid = 1u << (log_nodes - log_in);
    int mode = first ? leaf_mode : ST_NODES;
    if (l0 + a.levels == tree.log_max) a.fc = fc;  // this launch produces the root
    ProfScope ps(ctx, PT_MERKLE_TOP, 1, bytes);
    if (mode == ST_LEAVES) merkle_subtree_kernel<ST_LEAVES><<<grid, ST_THREADS, 0, st>>>(a);
    else if (mode == ST_FRI_LEAVES) merkle_subtree_kernel<ST_FRI_LEAVES><<<grid, ST_THREADS, 0, st>>>(a);
    else merkle_subtree_kernel<ST_NODES><<<grid, ST_THREADS, 0, st>>>(a);
    l0 += a.levels;
    first = false;
    if (a.levels == 0) break;
  }
  LSK_CUDA(cudaGetLastError());
}

void merkle_commit_dev(lsk_ctx* ctx, const std::vector<MatRef>& mats, MerkleTreeDev& tree) {
  if (mats.empty()) fail(LSK_ERR_ARG, "merkle_commit: no matrices");
  tree.mats = mats;
  tree.log_max = 0;
  for (auto& m : mats) tree.log_max = std::max<unsigned>(tree.log_max, m.log_h);
  // column pointer list grouped by height, tallest first, stable within a height (p3 sorts by
  // height descending, stable)
  std::vector<const uint32_t*> ptrs;
  std::vector<std::pair<uint32_t, uint32_t>> groups(tree.log_max + 1, {0u, 0u});
  for (unsigned l = 0; l <= tree.log_max; l++) {
    unsigned lh = tree.log_max - l;
    groups[l].first = (uint32_t)ptrs.size();
    for (auto& m : mats)
      if (m.log_h == lh)
        for (uint32_t c = 0; c < m.w; c++) ptrs.push_back(m.p + (size_t)c * m.stride);
    groups[l].second = (uint32_t)ptrs.size() - groups[l].first;
  }
  ptrs.push_back(nullptr);  // never empty
  tree.colptrs.alloc(ctx, ptrs.size());
  tree.colptrs.upload(ptrs.data(), ptrs.size());
  layout_layers(ctx, tree);
  build_tree(ctx, tree, groups, ST_LEAVES, nullptr);
}

void merkle_commit_fri_layer_dev(lsk_ctx* ctx, const uint32_t* vals, unsigned log_rows, MerkleTreeDev& tree,
                                 FriChallenge fc) {
  tree.mats.clear();
  tree.mats.push_back(MatRef{vals, 0, log_rows, 8});
  tree.log_max = log_rows;
  const uint32_t* none = nullptr;
  tree.colptrs.alloc(ctx, 1);
  tree.colptrs.upload(&none, 1);
  layout_layers(ctx, tree);
  std::vector<std::pair<uint32_t, uint32_t>> groups(log_rows + 1, {0u, 0u});
  build_tree(ctx, tree, groups, ST_FRI_LEAVES, vals, fc);
}

void fri_commit_tail_dev(lsk_ctx* ctx, unsigned first_lfh, Ext* const* vals, MerkleTreeDev* trees, const Ext* const* add,
                         FriChallenge fc) {
  if (first_lfh < 1 || first_lfh > FRI_TAIL_LOG) fail(LSK_ERR_ARG, "bad FRI tail height");
  FriTailArgs a{};
  a.nlayers = first_lfh;
  a.first_lfh = first_lfh;
  a.one_half = finv(to_mont(2));
  a.fc = fc;
  double bytes = 0;
  for (unsigned j = 0; j < first_lfh; j++) {
    unsigned lfh = first_lfh - j;
    MerkleTreeDev& t = trees[j];
    t.mats.clear();
    t.mats.push_back(MatRef{(const uint32_t*)vals[j], 0, lfh, 8});
    t.log_max = lfh;
    t.colptrs.release();
    layout_layers(ctx, t);
    a.vals[j] = vals[j];
    a.nodes[j] = t.nodes.p;
    a.add[j] = add[j];
    a.ginv[j] = finv(two_adic_generator(lfh + 1));
    bytes += (double)((size_t)1 << lfh) * (32.0 + 32.0 + 64.0);
  }
  a.vals[first_lfh] = vals[first_lfh];
  ProfScope ps(ctx, PT_MERKLE_TOP, 1, bytes);
  fri_tail_kernel<<<1, ST_THREADS, 0, ctx->stream>>>(a);
  LSK_CUDA(cudaGetLastError());
}

void merkle_root_host(lsk_ctx* ctx, const MerkleTreeDev& tree, uint32_t out8[8]) {
  LSK_CUDA(cudaMemcpyAsync(out8, tree.root_dev(), 32, cudaMemcpyDeviceToHost, ctx->stream));
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 8; i++) out8[i] = from_mont(out8[i]);
}

}  // namespace lsk
