// Merkle commitment kernels (see merkle.cuh for the contract and the reference map).
//
// One thread owns one sponge / compression: the 16-word Poseidon2 state stays in registers and a
// leaf absorbs 8 columns at a time.  Matrices are column-major, so for a fixed column consecutive
// threads (rows) read consecutive words: leaf hashing streams every LDE cell exactly once
// (8*N*w bytes for a blow-up-2 commit) and writes 32 B per row.  The stage is bound by the
// integer pipes (one width-16 permutation ~ 750 modmuls per 32 B absorbed), not by HBM.
#include <algorithm>

#include "merkle.cuh"
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
LSK_D void merkle_digest(const uint32_t* const* __restrict__ cols, uint32_t ncols, size_t r,
                         const uint32_t* __restrict__ packed_row, const uint32_t* __restrict__ children,
                         uint32_t (&st)[16]) {
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

// One launch builds up to ST_LEVELS + 1 layers: a CTA takes 2 * ST_THREADS consecutive nodes of
// layer l0 (in the leaf modes it first hashes them from the matrices) and reduces them to one node
// of layer l0 + ST_LEVELS; intermediate digests go straight to their layers in global memory (they
// are needed for the query openings) and are re-read by the same CTA through L1/L2.  A 2^21-leaf
// commitment is 3 launches instead of 22, and the sequential tail of a tree costs one launch.
constexpr int ST_THREADS = 256;
constexpr unsigned ST_LEVELS = 9;  // log2(2 * ST_THREADS)
struct SubtreeArgs {
  uint32_t* nodes;
  const uint32_t* const* cols;
  const uint32_t* fri_vals;
  uint32_t l0, levels, log_in;           // input layer, layers produced, log2(inputs per CTA)
  uint32_t off[ST_LEVELS + 1];           // digest offsets of layers l0 .. l0 + levels
  uint32_t inj_first[ST_LEVELS + 1];     // column-list range of the height group of layer l0 + k
  uint32_t inj_ncols[ST_LEVELS + 1];     // (k = 0: the leaf group in LEAF mode)
};
enum { ST_NODES = 0, ST_LEAVES = 1, ST_FRI_LEAVES = 2 };
template <int MODE>
__global__ void __launch_bounds__(ST_THREADS) merkle_subtree_kernel(SubtreeArgs a) {
  const size_t in0 = (size_t)blockIdx.x << a.log_in;
  const uint32_t nin = 1u << a.log_in;
#pragma unroll 1
  for (uint32_t k = (MODE == ST_NODES ? 1u : 0u); k <= a.levels; k++) {
    const uint32_t cnt = nin >> k;
    uint32_t* layer = a.nodes + (size_t)a.off[k] * 8;
    const uint32_t* prev = k ? a.nodes + (size_t)a.off[k - 1] * 8 : nullptr;
    const bool packed = MODE == ST_FRI_LEAVES && k == 0;
#pragma unroll 1
    for (uint32_t t = threadIdx.x; t < cnt; t += ST_THREADS) {
      const size_t i = (in0 >> k) + t;
      uint32_t st[16];
      merkle_digest(a.cols + a.inj_first[k], packed ? 0u : a.inj_ncols[k], i, packed ? a.fri_vals + i * 8 : nullptr,
                    k ? prev + 16 * i : nullptr, st);
      store_digest(layer + i * 8, st);
    }
    __syncthreads();
  }
}

// Throughput-bound part of a tree: one thread per leaf / per node, one launch per layer.
constexpr int MK_BLOCK = 128;
constexpr unsigned TAIL_LOG = 14;  // layers with <= 2^TAIL_LOG nodes are latency-bound: finished by subtree launches
template <int MODE>
__global__ void __launch_bounds__(MK_BLOCK) merkle_flat_kernel(const uint32_t* const* __restrict__ cols, uint32_t ncols,
                                                                const uint32_t* __restrict__ fri_vals,
                                                                const uint32_t* __restrict__ prev, size_t n,
                                                                uint32_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * MK_BLOCK + threadIdx.x;
  if (i >= n) return;
  uint32_t st[16];
  merkle_digest(cols, MODE == ST_FRI_LEAVES ? 0u : ncols, i, MODE == ST_FRI_LEAVES ? fri_vals + i * 8 : nullptr,
                MODE == ST_NODES ? prev + 16 * i : nullptr, st);
  store_digest(out + i * 8, st);
}

void init_merkle_kernels() { LSK_CUDA(p2_upload_constants()); }

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

// groups: per layer l (0..log_max) the column-list range injected there (layer 0 = leaves)
static void build_tree(lsk_ctx* ctx, MerkleTreeDev& tree, const std::vector<std::pair<uint32_t, uint32_t>>& groups,
                       int leaf_mode, const uint32_t* fri_vals) {
  cudaStream_t st = ctx->stream;
  unsigned l0 = 0;
  bool first = true;
  auto layer_bytes = [&](unsigned l) {
    return (double)((size_t)1 << (tree.log_max - l)) * (32.0 + 4.0 * groups[l].second + (l ? 64.0 : 0.0) +
                                                        (l == 0 && leaf_mode == ST_FRI_LEAVES ? 32.0 : 0.0));
  };
  // big layers: one thread per digest, one launch per layer
  while (tree.log_max - l0 > TAIL_LOG || (first && tree.log_max > ST_LEVELS)) {
    size_t n = (size_t)1 << (tree.log_max - l0);
    uint32_t* out = tree.nodes.p + tree.layer_off[l0] * 8;
    const uint32_t* const* cols = tree.colptrs.p + groups[l0].first;
    unsigned grid = div_up(n, MK_BLOCK);
    if (first) {
      ProfScope ps(ctx, leaf_mode == ST_LEAVES ? PT_MERKLE_LEAF : PT_FRI_LEAF, 1, layer_bytes(0));
      if (leaf_mode == ST_LEAVES)
        merkle_flat_kernel<ST_LEAVES><<<grid, MK_BLOCK, 0, st>>>(cols, groups[0].second, nullptr, nullptr, n, out);
      else
        merkle_flat_kernel<ST_FRI_LEAVES><<<grid, MK_BLOCK, 0, st>>>(cols, 0, fri_vals, nullptr, n, out);
      first = false;
    } else {
      ProfScope ps(ctx, PT_MERKLE_LAYER, 1, layer_bytes(l0));
      merkle_flat_kernel<ST_NODES><<<grid, MK_BLOCK, 0, st>>>(cols, groups[l0].second, nullptr,
                                                             tree.nodes.p + tree.layer_off[l0 - 1] * 8, n, out);
    }
    l0++;
  }
  if (!first) l0--;  // l0 = last layer already built
  // latency-bound tail (or a whole small tree): subtree launches
  while (first || l0 < tree.log_max) {
    unsigned log_nodes = tree.log_max - l0;  // log2(nodes of layer l0)
    unsigned log_in = std::min(log_nodes, ST_LEVELS);
    SubtreeArgs a{};
    a.nodes = tree.nodes.p;
    a.cols = tree.colptrs.p;
    a.fri_vals = fri_vals;
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

void merkle_commit_fri_layer_dev(lsk_ctx* ctx, const uint32_t* vals, unsigned log_rows, MerkleTreeDev& tree) {
  tree.mats.clear();
  tree.mats.push_back(MatRef{vals, 0, log_rows, 8});
  tree.log_max = log_rows;
  const uint32_t* none = nullptr;
  tree.colptrs.alloc(ctx, 1);
  tree.colptrs.upload(&none, 1);
  layout_layers(ctx, tree);
  std::vector<std::pair<uint32_t, uint32_t>> groups(log_rows + 1, {0u, 0u});
  build_tree(ctx, tree, groups, ST_FRI_LEAVES, vals);
}

void merkle_root_host(lsk_ctx* ctx, const MerkleTreeDev& tree, uint32_t out8[8]) {
  LSK_CUDA(cudaMemcpyAsync(out8, tree.root_dev(), 32, cudaMemcpyDeviceToHost, ctx->stream));
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 8; i++) out8[i] = from_mont(out8[i]);
}

}  // namespace lsk
