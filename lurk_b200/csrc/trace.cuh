// A4/A5 device-side declarations: the uploaded (device-resident) flattened record and the trace
// generation launchers shared by the boundary API (row-major canonical output) and the prover
// pipeline (column-major, bit-reversed rows, Montgomery form).
#pragma once
#include <vector>

#include "device.cuh"
#include "record.h"

namespace lsk {

struct DevFuncTable {
  const uint32_t *args, *outs, *provide, *req_off, *reqs, *hint_off, *hints;
  const uint32_t *depth, *dreq_off, *dreqs;   // partial funcs only (else null)
  uint32_t in, out, rows;
};
struct DevMemTable {
  const uint32_t *vals, *provide;
  uint32_t len, rows;
};

// Owns the device copies of a record's tables (uploaded once per prove call; this is the
// host->device traffic of the end-to-end path).
struct DeviceRecord {
  std::vector<DevBuf<uint32_t>> bufs;
  DevBuf<DevFuncTable> funcs;
  DevBuf<DevMemTable> mems;
  DevBuf<uint32_t> program;
  DevBuf<uint32_t> bytes;  // [65536][12] or empty
  std::vector<DevFuncTable> h_funcs;
  std::vector<DevMemTable> h_mems;
  size_t h2d_bytes = 0;
  // Outside a device group the tables travel on the context's side stream, in two waves: first what every trace kernel
  // may gather from (outputs, depths, arguments of invertible funcs, memory values, descriptors), then func by func the
  // per-row tables only that func's own trace reads.  A trace launch waits for `common_ready` and its func's event, so
  // trace generation and the LDE of the first chips run while the tables of the later ones are still crossing PCIe.
  cudaEvent_t common_ready = nullptr, all_ready = nullptr;
  std::vector<cudaEvent_t> func_ready;
  void wait_common(lsk_ctx* ctx) const { if (common_ready) cudaStreamWaitEvent(ctx->stream, common_ready, 0); }
  void wait_func(lsk_ctx* ctx, uint32_t f) const {
    wait_common(ctx);
    if (f < func_ready.size() && func_ready[f]) cudaStreamWaitEvent(ctx->stream, func_ready[f], 0);
  }
  void wait_all(lsk_ctx* ctx) const { if (all_ready) cudaStreamWaitEvent(ctx->stream, all_ready, 0); }
  DeviceRecord() = default;
  DeviceRecord(const DeviceRecord&) = delete;
  DeviceRecord& operator=(const DeviceRecord&) = delete;
  ~DeviceRecord() {
    if (all_ready) cudaEventSynchronize(all_ready);   // no copy may still be writing the buffers about to be released
    for (cudaEvent_t e : func_ready) if (e) cudaEventDestroy(e);
    if (common_ready) cudaEventDestroy(common_ready);
    if (all_ready) cudaEventDestroy(all_ready);
  }
  // shard_size == 0: the whole record.  Otherwise, if the record is row-ordered (lsk_record_finalize), only
  // what proving `shard` touches: the per-row tables of its row range (device pointers are biased so that
  // kernels keep indexing by absolute row) plus the tables other rows gather from (outputs, depths,
  // arguments of invertible funcs, memory values).
  void upload(lsk_ctx* ctx, const lsk_program& prog, const lsk_record& rec, uint32_t shard = 0, uint32_t shard_size = 0);
};

// Output layouts
enum class TraceLayout { RowMajorCanonical, ColMajorBitrevMont };

// d_out must be zero-filled by the callee: the launchers do it.
void func_trace_dev(lsk_ctx* ctx, const lsk_program& prog, const DeviceRecord& drec, uint32_t func, uint32_t row_start,
                    uint32_t real_rows, uint32_t height, TraceLayout layout, uint32_t* d_out);
void mem_trace_dev(lsk_ctx* ctx, const DeviceRecord& drec, uint32_t mem_idx, uint32_t height, TraceLayout layout,
                   uint32_t* d_out);
void bytes_trace_dev(lsk_ctx* ctx, const DeviceRecord& drec, bool real, TraceLayout layout, uint32_t* d_out);
void bytes_preprocessed_dev(lsk_ctx* ctx, TraceLayout layout, uint32_t* d_out);
// single-row Entrypoint trace = public values (src/lair/lair_chip.rs:112-118)
void init_trace_kernels();

}  // namespace lsk
