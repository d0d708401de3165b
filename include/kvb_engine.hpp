// kvb_engine.hpp — C++17 mirror of the reference's compiled engine class over the C ABI (kvb.h).
//
// Same constructor arguments and methods as storage_offload.StorageOffloadEngine
// (kv_connectors/llmd_fs_backend/csrc/storage/storage_offload.hpp:50-109, pybind signature
// storage_offload_bindings.cpp:25-94), with raw device pointers where the reference takes torch::Tensor
// (a (num_blocks, page_bytes) canonical KV tensor is just base pointer + num_blocks + page_bytes).
// Errors: booleans out, nothing thrown after construction (storage_offload.cpp:338-347).  Header-only; link -lkvb.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "kvb.h"

namespace kvb {

struct KVTensor {  // canonical KV-cache tensor: (num_blocks, page_bytes) bytes, contiguous rows
  const void* data;
  int64_t num_blocks;
  int64_t page_bytes;
};

class StorageOffloadEngine {
 public:
  StorageOffloadEngine(int io_threads, int gpu_blocks_per_file, const std::vector<KVTensor>& tensors,
                       int read_preferring_workers, const std::string& gds_mode = "disabled",
                       float max_write_queued_seconds = 10.0f, int device = 0, int tier = KVB_TIER_FILE,
                       int64_t host_arena_bytes = 0) {
    if (tensors.empty()) throw std::invalid_argument("TensorCopier: tensors is empty");            // tensor_copier.cu:34
    if (gpu_blocks_per_file <= 0) throw std::invalid_argument("TensorCopier: gpu_blocks_per_file must be > 0");
    std::vector<const void*> ptrs;
    for (const auto& t : tensors) {
      if (t.num_blocks != tensors[0].num_blocks || t.page_bytes != tensors[0].page_bytes)
        throw std::invalid_argument("all KV tensors must share num_blocks and page bytes");
      ptrs.push_back(t.data);
    }
    if (kvb_pool_create(device, ptrs.data(), (int32_t)ptrs.size(), tensors[0].num_blocks, tensors[0].page_bytes,
                        tensors[0].page_bytes, &pool_) != KVB_OK)
      throw std::runtime_error(kvb_last_error());
    kvb_engine_opts_t o;
    kvb_engine_default_opts(&o);
    o.io_threads = io_threads;
    o.gpu_blocks_per_file = gpu_blocks_per_file;
    o.read_preferring_workers = read_preferring_workers;
    o.max_write_queued_seconds = max_write_queued_seconds;
    o.tier = tier;
    o.host_arena_bytes = host_arena_bytes;
    // parse_gds_mode (gds_file_io.cpp:425-446): anything else means "disabled"
    if (gds_mode == "read_only") o.gds_mode = KVB_GDS_READ;
    else if (gds_mode == "write_only") o.gds_mode = KVB_GDS_WRITE;
    else if (gds_mode == "read_write") o.gds_mode = KVB_GDS_READ | KVB_GDS_WRITE;
    else if (gds_mode == "bb_read_only") o.gds_mode = KVB_GDS_READ | KVB_GDS_BOUNCE;
    else if (gds_mode == "bb_write_only") o.gds_mode = KVB_GDS_WRITE | KVB_GDS_BOUNCE;
    else if (gds_mode == "bb_read_write") o.gds_mode = KVB_GDS_READ | KVB_GDS_WRITE | KVB_GDS_BOUNCE;
    if (kvb_engine_create(pool_, &o, &eng_) != KVB_OK) {
      std::string msg = kvb_last_error();
      kvb_pool_destroy(pool_);
      throw std::runtime_error(msg);
    }
  }
  ~StorageOffloadEngine() {
    kvb_engine_destroy(eng_);
    kvb_pool_destroy(pool_);
  }
  StorageOffloadEngine(const StorageOffloadEngine&) = delete;
  StorageOffloadEngine& operator=(const StorageOffloadEngine&) = delete;

  // Async GPU -> Storage transfer (PUT); submit only (storage_offload.cpp:249)
  bool async_store_gpu_blocks(int job_id, const std::vector<std::string>& dst_files,
                              const std::vector<std::vector<int64_t>>& all_block_ids, void* stream = nullptr) {
    return submit(kvb_engine_store, job_id, dst_files, all_block_ids, stream);
  }
  // Async Storage -> GPU transfer (GET) (storage_offload.cpp:362)
  bool async_load_gpu_blocks(int job_id, const std::vector<std::string>& src_files,
                             const std::vector<std::vector<int64_t>>& all_block_ids, void* stream = nullptr) {
    return submit(kvb_engine_load, job_id, src_files, all_block_ids, stream);
  }
  // Return finished jobs and their success status; a job is reported once (storage_offload.cpp:185)
  std::vector<std::pair<int, bool>> get_finished() {
    std::vector<std::pair<int, bool>> out;
    int64_t ids[256];
    int32_t ok[256];
    for (;;) {
      const int n = kvb_engine_poll(eng_, ids, ok, 256);
      for (int i = 0; i < n; ++i) out.emplace_back((int)ids[i], ok[i] != 0);
      if (n < 256) return out;
    }
  }
  // Wait for all tasks of the job; cancels the ones still queued (storage_offload.cpp:214)
  void wait_job(int job_id) { kvb_engine_wait(eng_, job_id); }
  // SharedStorageOffloadingManager.lookup (llmd_fs_backend/manager.py:43-53) in one call: consecutive hits from the start.
  // `hashes` = the low 64 bits of the block hashes; the names <base>/<hhh>/<hh>/<016x>.bin are built inside the library.
  int lookup(const std::string& base_path, const std::vector<uint64_t>& hashes) {
    int32_t hits = 0;
    if (kvb_engine_lookup_prefix_hashes(eng_, base_path.c_str(), hashes.data(), (int32_t)hashes.size(), &hits) != KVB_OK) return 0;
    return hits;
  }

 private:
  template <typename Fn>
  bool submit(Fn fn, int job_id, const std::vector<std::string>& files, const std::vector<std::vector<int64_t>>& ids,
              void* stream) {
    if (files.size() != ids.size()) return false;
    std::vector<const char*> paths;
    std::vector<int64_t> flat, off{0};
    for (size_t i = 0; i < files.size(); ++i) {
      paths.push_back(files[i].c_str());
      flat.insert(flat.end(), ids[i].begin(), ids[i].end());
      off.push_back((int64_t)flat.size());
    }
    return fn(eng_, job_id, (int32_t)files.size(), paths.data(), flat.data(), off.data(), stream) == KVB_OK;
  }
  kvb_pool_t* pool_ = nullptr;
  kvb_engine_t* eng_ = nullptr;
};

}  // namespace kvb
