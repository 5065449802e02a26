// A9 (second half): Merkle commitment of column-major LDE matrices of mixed heights.
//
// Replaces the third-party `FieldMerkleTreeMmcs::commit` (PaddingFreeSponge<Perm16,16,8,8> leaves,
// TruncatedPermutation<Perm16,2,8,16> nodes, shorter matrices injected at their layer as
// compress(node, hash(rows))) that sphinx's prover reaches from `machine.prove`
// (src/core/cli/repl.rs:196); restated by the oracle in oracle/stark.h (merkle_commit).
#pragma once
#include <vector>

#include "device.cuh"
#include "field.cuh"

namespace lsk {

struct MatRef {             // a committed matrix: column-major, Montgomery form, rows in commitment order
  const uint32_t* p = nullptr;
  size_t stride = 0;        // words between columns
  uint32_t log_h = 0, w = 0;
};

struct MerkleTreeDev {
  std::vector<MatRef> mats;            // original (caller) order
  unsigned log_max = 0;
  DevBuf<uint32_t> nodes;              // all layers, 8 Montgomery words per digest; layer i at layer_off[i]
  std::vector<size_t> layer_off;       // in digests; layer 0 = leaves, layer log_max = root
  DevBuf<const uint32_t*> colptrs;     // per-column base pointers, grouped by height (tallest first)
  unsigned dist_layers = 0;            // built by a device group (comm.cuh): layers [0, dist_layers) hold only this
                                       // rank's range of nodes, the layers above are complete on every rank
  const uint32_t* layer(unsigned i) const { return nodes.p + layer_off[i] * 8; }
  const uint32_t* root_dev() const { return layer(log_max); }
};

// Device-side transcript step of the FRI commit phase: `chal` = the DuplexChallenger's 16-word state
// (Montgomery form, input buffer empty); after a layer's root is known the device observes it, samples
// the folding challenge into *beta_out and copies the root to root_out[8] -- no host round trip.
struct FriChallenge {
  uint32_t* chal = nullptr;
  Ext* beta_out = nullptr;
  uint32_t* root_out = nullptr;
};
constexpr unsigned FRI_TAIL_LOG = 9;   // layers with <= 2^9 rows are committed and folded by one fused launch

void init_merkle_kernels();
// Commits `mats` (heights are powers of two; at least one matrix).  Stream-ordered; nothing is
// copied back.
void merkle_commit_dev(lsk_ctx* ctx, const std::vector<MatRef>& mats, MerkleTreeDev& tree);
// FRI layer: `vals` is an array of n_ext quartic-extension elements (16 B each, Montgomery); the
// committed matrix is (n_ext / 2) x 8, row i = (vals[2i], vals[2i+1]).
void merkle_commit_fri_layer_dev(lsk_ctx* ctx, const uint32_t* vals, unsigned log_rows, MerkleTreeDev& tree,
                                 FriChallenge fc = {});
// Layers of 2^first_lfh .. 2 rows in one launch: vals[j] (j = 0 .. first_lfh) are the layer inputs and
// the final 2-element vector (all pre-allocated), trees[j] receives layer j's tree, add[j] (nullable)
// is added to the vector folded from layer j; fc.beta_out / fc.root_out point at layer 0's slots.
void fri_commit_tail_dev(lsk_ctx* ctx, unsigned first_lfh, Ext* const* vals, MerkleTreeDev* trees, const Ext* const* add,
                         FriChallenge fc);
// canonical root -> host (synchronises the stream)
void merkle_root_host(lsk_ctx* ctx, const MerkleTreeDev& tree, uint32_t out8[8]);

}  // namespace lsk
