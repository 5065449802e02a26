// Merkle commitment kernels (see merkle.cuh for the contract and the reference map).
//
// One thread owns one sponge / compression: the 16-word Poseidon2 state stays in registers and a
// leaf absorbs 8 columns at a time.  Matrices are column-major, so for a fixed column consecutive
// threads (rows) read consecutive words: leaf hashing streams every LDE cell exactly once
// (8*N*w bytes for a blow-up-2 commit) and writes 32 B per row.  The stage is bound by the
// integer pipes (one width-16 permutation ~ 750 modmuls per 32 B absorbed), not by HBM.
// Layers with <= 1024 nodes are finished by a single CTA in one launch.
#include <algorithm>

#include "merkle.cuh"
#include "poseidon2.cuh"

namespace lsk {

constexpr int MK_BLOCK = 128;
constexpr unsigned TOP_LOG = 10;  // layers of <= 2^TOP_LOG nodes are built by one CTA

// sponge over the concatenated row `r` of a height group given as a list of column base pointers
LSK_D void hash_row(const uint32_t* const* __restrict__ cols, uint32_t ncols, size_t r, uint32_t (&st)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++) st[i] = 0;
  for (uint32_t k = 0; k < ncols; k += 8) {
#pragma unroll
    for (int c = 0; c < 8; c++)
      if (k + c < ncols) st[c] = cols[k + c][r];
    p2_permute<16>(st);
  }
}
LSK_D void load_digest(const uint32_t* p, uint32_t* d) {
  uint4 a = reinterpret_cast<const uint4*>(p)[0], b = reinterpret_cast<const uint4*>(p)[1];
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
  d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
LSK_D void store_digest(uint32_t* p, const uint32_t* d) {
  reinterpret_cast<uint4*>(p)[0] = make_uint4(d[0], d[1], d[2], d[3]);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(d[4], d[5], d[6], d[7]);
}

__global__ void __launch_bounds__(MK_BLOCK) merkle_leaf_kernel(const uint32_t* const* __restrict__ cols, uint32_t ncols,
                                                                size_t rows, uint32_t* __restrict__ digests) {
  size_t r = (size_t)blockIdx.x * MK_BLOCK + threadIdx.x;
  if (r >= rows) return;
  uint32_t st[16];
  hash_row(cols, ncols, r, st);
  store_digest(digests + r * 8, st);
}

// FRI layer leaves: row r = 8 consecutive words
__global__ void __launch_bounds__(MK_BLOCK) merkle_fri_leaf_kernel(const uint32_t* __restrict__ vals, size_t rows,
                                                                    uint32_t* __restrict__ digests) {
  size_t r = (size_t)blockIdx.x * MK_BLOCK + threadIdx.x;
  if (r >= rows) return;
  uint32_t st[16];
  load_digest(vals + r * 8, st);
#pragma unroll
  for (int i = 8; i < 16; i++) st[i] = 0;
  p2_permute<16>(st);
  store_digest(digests + r * 8, st);
}

// next[i] = compress(prev[2i], prev[2i+1]); with an injected height group: compress(that, hash(rows_i))
LSK_D void node_step(const uint32_t* __restrict__ prev, uint32_t* __restrict__ next, size_t i,
                     const uint32_t* const* __restrict__ inj_cols, uint32_t inj_ncols) {
  uint32_t st[16];
  load_digest(prev + 16 * i, st);
  load_digest(prev + 16 * i + 8, st + 8);
  p2_permute<16>(st);
  if (inj_ncols) {
    uint32_t rd[16];
    hash_row(inj_cols, inj_ncols, i, rd);
#pragma unroll
    for (int k = 0; k < 8; k++) st[8 + k] = rd[k];
    p2_permute<16>(st);
  }
  store_digest(next + 8 * i, st);
}
__global__ void __launch_bounds__(MK_BLOCK) merkle_layer_kernel(const uint32_t* __restrict__ prev,
                                                                 uint32_t* __restrict__ next, size_t n,
                                                                 const uint32_t* const* __restrict__ inj_cols,
                                                                 uint32_t inj_ncols) {
  size_t i = (size_t)blockIdx.x * MK_BLOCK + threadIdx.x;
  if (i < n) node_step(prev, next, i, inj_cols, inj_ncols);
}

struct TopLayers {  // layer l (l = first .. log_max): nodes at base + off[l] * 8
  uint32_t first, log_max;
  uint32_t off[TOP_LOG + 2];        // digest offsets of layers first-1 .. log_max
  uint32_t inj_first[TOP_LOG + 1];  // column-list offset of the group injected at layer first + k
  uint32_t inj_ncols[TOP_LOG + 1];
};
__global__ void __launch_bounds__(1 << TOP_LOG) merkle_top_kernel(uint32_t* __restrict__ nodes, TopLayers tl,
                                                                   const uint32_t* const* __restrict__ cols) {
  for (uint32_t l = tl.first; l <= tl.log_max; l++) {
    uint32_t k = l - tl.first;
    size_t n = (size_t)1 << (tl.log_max - l);
    if (threadIdx.x < n)
      node_step(nodes + (size_t)tl.off[k] * 8, nodes + (size_t)tl.off[k + 1] * 8, threadIdx.x, cols + tl.inj_first[k],
                tl.inj_ncols[k]);
    __syncthreads();
  }
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
static void build_upper(lsk_ctx* ctx, MerkleTreeDev& tree, const std::vector<std::pair<uint32_t, uint32_t>>& groups) {
  cudaStream_t st = ctx->stream;
  unsigned l = 1;
  for (; l <= tree.log_max && tree.log_max - l > TOP_LOG; l++) {
    size_t n = (size_t)1 << (tree.log_max - l);
    ProfScope ps(ctx, PT_MERKLE_LAYER, 1, (double)n * (96.0 + 4.0 * groups[l].second));
    merkle_layer_kernel<<<div_up(n, MK_BLOCK), MK_BLOCK, 0, st>>>(
        tree.nodes.p + tree.layer_off[l - 1] * 8, tree.nodes.p + tree.layer_off[l] * 8, n,
        tree.colptrs.p + groups[l].first, groups[l].second);
  }
  if (l <= tree.log_max) {
    TopLayers tl{};
    tl.first = l;
    tl.log_max = tree.log_max;
    for (unsigned k = 0; l - 1 + k <= tree.log_max; k++) tl.off[k] = (uint32_t)tree.layer_off[l - 1 + k];
    for (unsigned k = 0; l + k <= tree.log_max; k++) {
      tl.inj_first[k] = groups[l + k].first;
      tl.inj_ncols[k] = groups[l + k].second;
    }
    unsigned threads = 1u << (tree.log_max - l);
    threads = std::max(32u, threads);
    ProfScope ps(ctx, PT_MERKLE_TOP, 1, 96.0 * 2 * threads);
    merkle_top_kernel<<<1, threads, 0, st>>>(tree.nodes.p, tl, tree.colptrs.p);
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
  size_t rows = (size_t)1 << tree.log_max;
  {
    ProfScope ps(ctx, PT_MERKLE_LEAF, 1, (double)rows * (4.0 * groups[0].second + 32.0));
    merkle_leaf_kernel<<<div_up(rows, MK_BLOCK), MK_BLOCK, 0, ctx->stream>>>(tree.colptrs.p + groups[0].first,
                                                                            groups[0].second, rows, tree.nodes.p);
  }
  LSK_CUDA(cudaGetLastError());
  build_upper(ctx, tree, groups);
}

void merkle_commit_fri_layer_dev(lsk_ctx* ctx, const uint32_t* vals, unsigned log_rows, MerkleTreeDev& tree) {
  tree.mats.clear();
  tree.mats.push_back(MatRef{vals, 0, log_rows, 8});
  tree.log_max = log_rows;
  const uint32_t* none = nullptr;
  tree.colptrs.alloc(ctx, 1);
  tree.colptrs.upload(&none, 1);
  layout_layers(ctx, tree);
  size_t rows = (size_t)1 << log_rows;
  {
    ProfScope ps(ctx, PT_FRI_LEAF, 1, 64.0 * rows);
    merkle_fri_leaf_kernel<<<div_up(rows, MK_BLOCK), MK_BLOCK, 0, ctx->stream>>>(vals, rows, tree.nodes.p);
  }
  LSK_CUDA(cudaGetLastError());
  std::vector<std::pair<uint32_t, uint32_t>> groups(log_rows + 1, {0u, 0u});
  build_upper(ctx, tree, groups);
}

void merkle_root_host(lsk_ctx* ctx, const MerkleTreeDev& tree, uint32_t out8[8]) {
  LSK_CUDA(cudaMemcpyAsync(out8, tree.root_dev(), 32, cudaMemcpyDeviceToHost, ctx->stream));
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 8; i++) out8[i] = from_mont(out8[i]);
}

}  // namespace lsk
