// index.cu — kvblock index (request key -> pod entries) and longest-prefix scorer.
//
// Replaces InMemoryIndex (pkg/kvcache/kvblock/in_memory.go:57-304) and
// LongestPrefixScorer.Score (pkg/kvcache/kvblock_scorer.go:91-154).
//
// Split of work
//   * MUTATIONS (Add / Evict / GetRequestKey; in_memory.go:154-304) arrive serially from the KV-event
//     stream and carry exact LRU semantics (10 pods per key, `Size` keys, engine->request map).  They are
//     applied to a host-side authoritative structure and recorded as dirty keys.
//   * READS — Lookup (in_memory.go:107-148) and Score — run on the GPU against a device mirror:
//     an open-addressing table of 64 B buckets {key, state|count, 13 x (pod:16 | tier:8 | spec:8)} in HBM.
//     Dirty keys are pushed as 64 B bucket images by one upsert kernel before the next read.
//   * Batched scoring: one warp per prompt.  32 keys are probed in parallel (random 64 B reads), their
//     buckets staged in shared memory, then the serial prefix walk adds float64 weights in key order,
//     which keeps the sums bit-identical to the Go loop (kvblock_scorer.go:132-150).
// Random-access bound (2 sectors per probe), not bandwidth bound.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <list>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "kvb_internal.h"

namespace kvb {

constexpr int kMaxEnt = KVB_INDEX_MAX_PODS_PER_KEY;  // 13
constexpr uint32_t kEmpty = 0, kFull = 1, kTomb = 2, kBusy = 3;

struct __align__(64) Bucket {
  uint64_t key;
  uint32_t meta;  // bits 0-1 state, bits 8-15 entry count
  uint32_t ent[kMaxEnt];
};
static_assert(sizeof(Bucket) == 64, "bucket must be one 64 B line");

struct __align__(64) Op {  // one mutation shipped to the device
  uint64_t key;
  uint32_t count;  // 0xffffffff = delete
  uint32_t ent[kMaxEnt];
};
static_assert(sizeof(Op) == 64, "op must be 64 B");

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t k) {  // murmur3 fmix64
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}
__host__ __device__ __forceinline__ uint32_t pack_entry(uint16_t pod, uint8_t tier, uint8_t spec) {
  return (uint32_t)pod | ((uint32_t)tier << 16) | ((uint32_t)(spec ? 1 : 0) << 24);
}

// ----------------------------------------------------------------------------------------- kernels
__global__ void index_apply_kernel(Bucket* __restrict__ table, uint64_t mask, const Op* __restrict__ ops, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Op op = ops[i];
  const bool del = op.count == 0xffffffffu;
  uint64_t slot = mix64(op.key) & mask;
  int64_t found = -1, reuse = -1;
  for (uint64_t probes = 0; probes <= mask; ++probes, slot = (slot + 1) & mask) {
    const uint32_t m = *reinterpret_cast<volatile uint32_t*>(&table[slot].meta);
    const uint32_t st = m & 3u;
    if (st == kEmpty) {
      if (reuse < 0) reuse = (int64_t)slot;
      break;
    }
    if (st == kTomb) {
      if (reuse < 0) reuse = (int64_t)slot;
    } else if (st == kFull && *reinterpret_cast<volatile uint64_t*>(&table[slot].key) == op.key) {
      found = (int64_t)slot;
      break;
    }
  }
  if (found >= 0) {
    Bucket& b = table[found];
    if (del) {
      b.meta = kTomb;
    } else {
#pragma unroll
      for (int e = 0; e < kMaxEnt; ++e) b.ent[e] = op.ent[e];
      b.meta = kFull | (op.count << 8);
    }
    return;
  }
  if (del) return;
  // claim a free slot; ops in one batch carry distinct keys, so a lost race just moves on
  slot = reuse >= 0 ? (uint64_t)reuse : (mix64(op.key) & mask);
  for (uint64_t probes = 0; probes <= mask; ++probes, slot = (slot + 1) & mask) {
    const uint32_t m = *reinterpret_cast<volatile uint32_t*>(&table[slot].meta);
    const uint32_t st = m & 3u;
    if (st != kEmpty && st != kTomb) continue;
    if (atomicCAS(&table[slot].meta, m, kBusy) != m) continue;
    Bucket& b = table[slot];
    b.key = op.key;
#pragma unroll
    for (int e = 0; e < kMaxEnt; ++e) b.ent[e] = op.ent[e];
    __threadfence();
    b.meta = kFull | (op.count << 8);
    return;
  }
}

__device__ __forceinline__ int64_t probe(const Bucket* __restrict__ table, uint64_t mask, uint64_t key) {
  uint64_t slot = mix64(key) & mask;
  for (uint64_t probes = 0; probes <= mask; ++probes, slot = (slot + 1) & mask) {
    const uint32_t st = table[slot].meta & 3u;
    if (st == kEmpty) return -1;
    if (st == kFull && table[slot].key == key) return (int64_t)slot;
  }
  return -1;
}

__device__ __forceinline__ bool pod_allowed(const uint32_t* __restrict__ filter_bits, uint32_t pod) {
  return filter_bits == nullptr || ((filter_bits[pod >> 5] >> (pod & 31)) & 1u);
}

// Lookup: one thread per key.  counts: -1 absent, -2 present but empty, else #entries after the pod filter.
__global__ void index_lookup_kernel(const Bucket* __restrict__ table, uint64_t mask, const uint64_t* __restrict__ keys,
                                    int64_t n, const uint32_t* __restrict__ filter_bits, int32_t* __restrict__ counts,
                                    uint32_t* __restrict__ out_ent) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t s = probe(table, mask, keys[i]);
  if (s < 0) {
    counts[i] = -1;
    return;
  }
  const Bucket& b = table[s];
  const int cnt = (int)((b.meta >> 8) & 0xff);
  if (cnt == 0) {
    counts[i] = -2;
    return;
  }
  int k = 0;
  for (int e = 0; e < cnt; ++e) {
    const uint32_t v = b.ent[e];
    if (pod_allowed(filter_bits, v & 0xffffu)) out_ent[i * kMaxEnt + k++] = v;
  }
  counts[i] = k;
}

// Score: one warp per prompt (4 warps per CTA).
constexpr int kScoreWarps = 4;

__global__ void __launch_bounds__(kScoreWarps * 32)
    index_score_kernel(const Bucket* __restrict__ table, uint64_t mask, const uint64_t* __restrict__ keys,
                       const int64_t* __restrict__ key_off, int32_t n_prompts,
                       const uint32_t* __restrict__ filter_bits, const double* __restrict__ tier_w,
                       int32_t* __restrict__ out_n, uint16_t* __restrict__ out_pods, double* __restrict__ out_scores,
                       uint8_t* __restrict__ found_flags /* nullable: one byte per key */) {
  __shared__ Bucket tile[kScoreWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * kScoreWarps + warp;
  if (p >= n_prompts) return;
  const int64_t k0 = key_off[p];
  const int64_t nk = key_off[p + 1] - k0;
  const unsigned FULL = 0xffffffffu;

  bool active = false;  // this lane owns a pod that is still on the consecutive prefix
  bool owner = false;   // this lane owns a pod that appeared at key 0 (it is reported)
  uint32_t my_pod = 0xffffffffu;
  double score = 0.0;
  bool chain_alive = true;

  for (int64_t base = 0; base < nk; base += 32) {
    if (!chain_alive && found_flags == nullptr) break;
    // ---- phase A: 32 independent probes, buckets staged in shared memory
    const int64_t ki = base + lane;
    int64_t slot = -1;
    if (ki < nk) slot = probe(table, mask, keys[k0 + ki]);
    if (found_flags != nullptr && ki < nk) found_flags[k0 + ki] = slot >= 0 ? 1 : 0;
    if (slot >= 0) {
      const uint4* src = reinterpret_cast<const uint4*>(&table[slot]);
      uint4* dst = reinterpret_cast<uint4*>(&tile[warp][lane]);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = src[q];
    }
    __syncwarp();
    if (!chain_alive) continue;
    const int in_tile = (int)min((int64_t)32, nk - base);
    // ---- phase B: serial walk over the tile's keys
    for (int j = 0; j < in_tile; ++j) {
      const int64_t sj = __shfl_sync(FULL, slot, j);
      // bucket entries sit on lanes 16..28, the reported pods (owners) on lanes 0..12 in the order they appeared at
      // key 0: ONE match.any per key pairs every owner with the entries that carry its pod (13 x 3 shuffles before)
      bool valid = false;
      uint32_t pod = 0x10000u + lane;  // unique sentinel for lanes without a valid entry
      double w = 0.0;
      if (sj >= 0) {
        const Bucket& b = tile[warp][j];
        const int cnt = (int)((b.meta >> 8) & 0xff);
        const int e = lane - 16;
        if (e >= 0 && e < cnt && e < kMaxEnt) {
          const uint32_t v = b.ent[e];
          if (pod_allowed(filter_bits, v & 0xffffu)) {
            valid = true;
            pod = v & 0xffffu;
            w = tier_w[(v >> 16) & 0xffu];
          }
        }
      }
      if (base + j == 0) {
        // key 0: active set = distinct pods, weight = max over that pod's tiers (fillMaxWeights)
        const unsigned group = __match_any_sync(FULL, pod) & 0x1fff0000u;  // entry lanes with this lane's pod
        const bool leader = valid && (group & ((1u << lane) - 1u)) == 0u;   // first entry of its pod
        double wmax = w;
        unsigned rest = valid ? (group & ~(1u << lane)) : 0u;
        while (__any_sync(FULL, rest != 0u)) {  // usually zero or one round: a pod on two tiers
          const int src = rest ? __ffs((int)rest) - 1 : lane;
          const double we = __shfl_sync(FULL, w, src);
          if (rest) {
            if (we > wmax) wmax = we;
            rest &= rest - 1u;
          }
        }
        const bool l2 = __shfl_down_sync(FULL, leader ? 1 : 0, 16) != 0;
        const uint32_t p2 = __shfl_down_sync(FULL, pod, 16);
        const double w2 = __shfl_down_sync(FULL, wmax, 16);
        owner = active = lane < kMaxEnt && l2;
        my_pod = owner ? p2 : 0xffffffffu;
        score = owner ? w2 : 0.0;
      } else {
        const uint32_t val = lane < 16 ? (active ? my_pod : 0x20000u + lane) : pod;
        unsigned em = __match_any_sync(FULL, val) >> 16;  // entry lanes (as bits 0..12) that carry my_pod
        if (!(lane < 16 && active)) em = 0u;
        const bool hit = em != 0u;
        double wm = 0.0;
        bool first = true;
        while (__any_sync(FULL, em != 0u)) {
          const int src = em ? 16 + __ffs((int)em) - 1 : lane;
          const double we = __shfl_sync(FULL, w, src);
          if (em) {
            if (first || we > wm) wm = we;
            first = false;
            em &= em - 1u;
          }
        }
        if (active) {
          if (hit) score += wm;  // float64, key order: same sum as the Go loop
          else active = false;
        }
      }
      if (!__any_sync(FULL, active)) {
        chain_alive = false;
        break;
      }
    }
    __syncwarp();
  }
  // compact (pod, score) pairs of the owner lanes
  const unsigned om = __ballot_sync(FULL, owner);
  if (owner) {
    const int pos = __popc(om & ((1u << lane) - 1u));
    out_pods[(int64_t)p * kMaxEnt + pos] = (uint16_t)my_pod;
    out_scores[(int64_t)p * kMaxEnt + pos] = score;
  }
  if (lane == 0) out_n[p] = __popc(om);
}

// ----------------------------------------------------------------------------------------- host side
struct KeyNode {
  uint8_t count = 0;
  kvb_pod_entry_t e[kMaxEnt];  // oldest -> newest (golang-lru Keys() order)
  std::list<uint64_t>::iterator lru;
};
struct EngNode {
  std::vector<uint64_t> rks;
  std::list<uint64_t>::iterator lru;
};

static inline bool same_entry(const kvb_pod_entry_t& a, const kvb_pod_entry_t& b) {
  return a.pod == b.pod && a.tier == b.tier && (a.speculative != 0) == (b.speculative != 0);
}

}  // namespace kvb

using namespace kvb;

struct kvb_index {
  int device = 0;
  int64_t max_keys = 0;
  int pods_per_key = 10;
  std::mutex mu;

  // host authoritative state
  std::unordered_map<uint64_t, KeyNode> data;
  std::list<uint64_t> data_lru;  // front = oldest
  std::unordered_map<uint64_t, EngNode> eng;
  std::list<uint64_t> eng_lru;
  std::vector<uint64_t> dirty;

  // device mirror
  Bucket* table = nullptr;
  uint64_t slots = 0;
  int64_t dev_live = 0, dev_tomb = 0;
  bool stale_mirror = false;  // the device table misses host entries (a rebuild failed after the swap)
  double tier_w_host[256];
  double* tier_w = nullptr;
  cudaStream_t stream = nullptr;
  // scratch
  uint8_t* d_scratch = nullptr;
  size_t d_scratch_cap = 0;
  uint8_t* h_scratch = nullptr;
  size_t h_scratch_cap = 0;
  uint32_t* d_filter = nullptr;  // 65536 bits

  int ensure_scratch(size_t dev_bytes, size_t host_bytes) {
    if (dev_bytes > d_scratch_cap) {
      if (d_scratch) cudaFree(d_scratch);
      d_scratch = nullptr;
      d_scratch_cap = 0;  // a failed allocation below must not leave a stale capacity behind
      size_t cap = std::max<size_t>(dev_bytes, 1 << 20);
      cap = (cap * 3 / 2 + 255) & ~size_t(255);
      KVB_CUDA_TRY(cudaMalloc(&d_scratch, cap));
      d_scratch_cap = cap;
    }
    if (host_bytes > h_scratch_cap) {
      if (h_scratch) cudaFreeHost(h_scratch);
      h_scratch = nullptr;
      h_scratch_cap = 0;
      size_t cap = std::max<size_t>(host_bytes, 1 << 20);
      cap = (cap * 3 / 2 + 255) & ~size_t(255);
      KVB_CUDA_TRY(host_alloc_near(device, reinterpret_cast<void**>(&h_scratch), cap, cudaHostAllocDefault));
      h_scratch_cap = cap;
    }
    return KVB_OK;
  }

  void touch(std::unordered_map<uint64_t, KeyNode>::iterator it) {
    data_lru.erase(it->second.lru);
    data_lru.push_back(it->first);
    it->second.lru = std::prev(data_lru.end());
  }
  void touch_eng(std::unordered_map<uint64_t, EngNode>::iterator it) {
    eng_lru.erase(it->second.lru);
    eng_lru.push_back(it->first);
    it->second.lru = std::prev(eng_lru.end());
  }
  void eng_add(uint64_t ek, std::vector<uint64_t>&& rks) {  // lru.Add: update + move to front, evict oldest
    auto it = eng.find(ek);
    if (it != eng.end()) {
      it->second.rks = std::move(rks);
      touch_eng(it);
      return;
    }
    EngNode n;
    n.rks = std::move(rks);
    eng_lru.push_back(ek);
    n.lru = std::prev(eng_lru.end());
    eng.emplace(ek, std::move(n));
    if ((int64_t)eng.size() > max_keys) {
      uint64_t old = eng_lru.front();
      eng_lru.pop_front();
      eng.erase(old);
    }
  }
  void erase_key(std::unordered_map<uint64_t, KeyNode>::iterator it) {
    dirty.push_back(it->first);
    data_lru.erase(it->second.lru);
    data.erase(it);
  }

  int set_filter(const uint16_t* pods, int32_t n, const uint32_t** out) {
    *out = nullptr;
    if (n <= 0) return KVB_OK;
    std::vector<uint32_t> bits(2048, 0u);
    for (int32_t i = 0; i < n; ++i) bits[pods[i] >> 5] |= 1u << (pods[i] & 31);
    if (!d_filter) KVB_CUDA_TRY(cudaMalloc(&d_filter, 2048 * sizeof(uint32_t)));
    KVB_CUDA_TRY(cudaMemcpyAsync(d_filter, bits.data(), 2048 * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    KVB_CUDA_TRY(cudaStreamSynchronize(stream));  // bits is a stack temporary
    *out = d_filter;
    return KVB_OK;
  }

  int rebuild(uint64_t new_slots);
  int flush_locked();
};

static void node_image(uint64_t key, const KeyNode& n, Op* op) {
  op->key = key;
  op->count = n.count;
  for (int e = 0; e < kMaxEnt; ++e)
    op->ent[e] = e < n.count ? pack_entry(n.e[e].pod, n.e[e].tier, n.e[e].speculative) : 0u;
}

static int apply_ops(kvb_index* idx, const std::vector<Op>& ops) {
  if (ops.empty()) return KVB_OK;
  const size_t bytes = ops.size() * sizeof(Op);
  int rc = idx->ensure_scratch(bytes, 0);
  if (rc) return rc;
  // ops is pageable host memory: the copy is staged by the runtime before returning
  KVB_CUDA_TRY(cudaMemcpyAsync(idx->d_scratch, ops.data(), bytes, cudaMemcpyHostToDevice, idx->stream));
  const int threads = 128;
  const int64_t grid = ((int64_t)ops.size() + threads - 1) / threads;
  index_apply_kernel<<<(unsigned)grid, threads, 0, idx->stream>>>(idx->table, idx->slots - 1,
                                                                 reinterpret_cast<const Op*>(idx->d_scratch),
                                                                 (int64_t)ops.size());
  KVB_CUDA_TRY(cudaGetLastError());
  count_launch();
  KVB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
  return KVB_OK;
}

int kvb_index::rebuild(uint64_t new_slots) {
  // the new table is allocated BEFORE the old one goes: a failed allocation leaves the mirror as it was (stale but
  // consistent; dirty keys stay queued), a failure while refilling leaves stale_mirror set so the next flush retries
  Bucket* fresh = nullptr;
  KVB_CUDA_TRY(cudaMalloc(&fresh, new_slots * sizeof(Bucket)));
  cudaError_t e = cudaMemsetAsync(fresh, 0, new_slots * sizeof(Bucket), stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) {
    cudaFree(fresh);
    set_error("index rebuild: %s", cudaGetErrorString(e));
    return KVB_ERR_CUDA;
  }
  if (table) cudaFree(table);
  table = fresh;
  slots = new_slots;
  dev_live = 0;
  dev_tomb = 0;
  stale_mirror = true;
  std::vector<Op> ops;
  const size_t batch = 1 << 20;
  ops.reserve(std::min(batch, data.size()));
  for (auto& kv : data) {
    Op op;
    node_image(kv.first, kv.second, &op);
    ops.push_back(op);
    if (ops.size() == batch) {
      int rc = apply_ops(this, ops);
      if (rc) return rc;
      ops.clear();
    }
  }
  int rc = apply_ops(this, ops);
  if (rc) return rc;
  dev_live = (int64_t)data.size();
  dirty.clear();
  stale_mirror = false;
  return KVB_OK;
}

int kvb_index::flush_locked() {
  if (stale_mirror) return rebuild(slots);  // an earlier rebuild stopped half-way
  if (dirty.empty()) return KVB_OK;
  std::sort(dirty.begin(), dirty.end());
  dirty.erase(std::unique(dirty.begin(), dirty.end()), dirty.end());
  // grow before the table gets crowded: live + tombstones + incoming <= 0.6 * slots
  const uint64_t need = (uint64_t)(data.size() + dev_tomb + dirty.size());
  if (need * 10 > slots * 6) {
    uint64_t ns = slots;
    while ((uint64_t)data.size() * 10 > ns * 3) ns <<= 1;  // target load <= 0.3 after rebuild
    if (ns == slots && (uint64_t)(data.size() + dirty.size()) * 10 > slots * 6) ns <<= 1;
    return rebuild(ns);  // rebuild drops every tombstone and clears dirty
  }
  std::vector<Op> ops;
  ops.reserve(dirty.size());
  int64_t dels = 0;
  for (uint64_t k : dirty) {
    Op op;
    auto it = data.find(k);
    if (it == data.end()) {
      op.key = k;
      op.count = 0xffffffffu;
      std::memset(op.ent, 0, sizeof(op.ent));
      ++dels;
    } else {
      node_image(k, it->second, &op);
    }
    ops.push_back(op);
  }
  int rc = apply_ops(this, ops);
  if (rc) return rc;
  dev_tomb += dels;  // upper bound (a delete of a never-flushed key leaves no tombstone)
  dirty.clear();
  return KVB_OK;
}

extern "C" {

int kvb_index_create(int device, int64_t max_keys, int32_t pods_per_key, int64_t expected_keys, kvb_index_t** out) {
  return kvb::guarded([&]() -> int {
    KVB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    KVB_REQUIRE(max_keys > 0, "must provide a positive size");  // golang-lru New()
    KVB_REQUIRE(pods_per_key > 0, "must provide a positive size");
    if (pods_per_key > kMaxEnt) {
      set_error("podCacheSize %d exceeds the %d entries of a device bucket", pods_per_key, kMaxEnt);
      return KVB_ERR_UNSUPPORTED;
    }
    DeviceGuard g(device);
    if (!g.ok) {
      set_error("cannot select CUDA device %d", device);
      return KVB_ERR_CUDA;
    }
    // kvb_index_destroy releases whatever CUDA resources exist if a later step fails
    std::unique_ptr<kvb_index, void (*)(kvb_index*)> idx(new kvb_index(), kvb_index_destroy);
    idx->device = device;
    idx->max_keys = max_keys;
    idx->pods_per_key = pods_per_key;
    for (int i = 0; i < 256; ++i) idx->tier_w_host[i] = 1.0;  // unknown tier -> 1.0 (kvblock_scorer.go:93-98)
    KVB_CUDA_TRY(cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking));
    KVB_CUDA_TRY(cudaMalloc(&idx->tier_w, 256 * sizeof(double)));
    KVB_CUDA_TRY(cudaMemcpy(idx->tier_w, idx->tier_w_host, 256 * sizeof(double), cudaMemcpyHostToDevice));
    int64_t exp_keys = std::max<int64_t>(expected_keys, 1024);
    exp_keys = std::min<int64_t>(exp_keys, max_keys);
    uint64_t slots = 2048;
    while (slots * 3 < (uint64_t)exp_keys * 10) slots <<= 1;  // load <= 0.3 at expected size
    idx->slots = slots;
    KVB_CUDA_TRY(cudaMalloc(&idx->table, slots * sizeof(Bucket)));
    KVB_CUDA_TRY(cudaMemset(idx->table, 0, slots * sizeof(Bucket)));
    if (expected_keys > 0) idx->data.reserve((size_t)std::min<int64_t>(expected_keys, max_keys));
    *out = idx.release();
    return KVB_OK;
  });
}

void kvb_index_destroy(kvb_index_t* idx) {
  if (!idx) return;
  DeviceGuard g(idx->device);
  if (idx->stream) cudaStreamSynchronize(idx->stream);
  if (idx->table) cudaFree(idx->table);
  if (idx->tier_w) cudaFree(idx->tier_w);
  if (idx->d_scratch) cudaFree(idx->d_scratch);
  if (idx->h_scratch) cudaFreeHost(idx->h_scratch);
  if (idx->d_filter) cudaFree(idx->d_filter);
  if (idx->stream) cudaStreamDestroy(idx->stream);
  delete idx;
}

int kvb_index_set_tier_weight(kvb_index_t* idx, uint8_t tier, double weight, int known) {
  return kvb::guarded([&]() -> int {
    KVB_REQUIRE(idx != nullptr, "index is NULL");
    std::lock_guard<std::mutex> lk(idx->mu);
    DeviceGuard g(idx->device);
    idx->tier_w_host[tier] = known ? weight : 1.0;
    KVB_CUDA_TRY(cudaMemcpy(idx->tier_w + tier, &idx->tier_w_host[tier], sizeof(double), cudaMemcpyHostToDevice));
    return KVB_OK;
  });
}

int kvb_index_add(kvb_index_t* idx, const uint64_t* engine_keys, int64_t n_engine, int has_engine_keys,
                  const uint64_t* request_keys, int64_t n_request, const kvb_pod_entry_t* entries, int32_t n_entries) {
  return kvb::guarded([&]() -> int {
    KVB_REQUIRE(idx != nullptr, "index is NULL");
    // in_memory.go:155-157
    KVB_REQUIRE(n_request > 0 && n_entries > 0 && request_keys && entries,
                "no keys or entries provided for adding to index");
    KVB_REQUIRE(!has_engine_keys || n_engine > 0, "engineKeys is non-nil but empty");
    std::lock_guard<std::mutex> lk(idx->mu);
    if (has_engine_keys) {  // in_memory.go:166-177
      const int64_t n = std::max(n_engine, n_request);
      std::vector<uint64_t> order;
      std::unordered_map<uint64_t, std::vector<uint64_t>> m;
      for (int64_t i = 0; i < n; ++i) {
        const uint64_t ek = engine_keys[i * n_engine / n];
        const uint64_t rk = request_keys[i * n_request / n];
        auto it = m.find(ek);
        if (it == m.end()) {
          order.push_back(ek);
          m[ek].push_back(rk);
        } else {
          it->second.push_back(rk);
        }
      }
      for (uint64_t ek : order) idx->eng_add(ek, std::move(m[ek]));
    }
    for (int64_t i = 0; i < n_request; ++i) {  // in_memory.go:180-221
      const uint64_t rk = request_keys[i];
      auto it = idx->data.find(rk);
      if (it != idx->data.end()) {
        idx->touch(it);
      } else {
        KeyNode n;
        idx->data_lru.push_back(rk);
        n.lru = std::prev(idx->data_lru.end());
        it = idx->data.emplace(rk, n).first;
        if ((int64_t)idx->data.size() > idx->max_keys) {  // outer LRU evicts the oldest key
          const uint64_t old = idx->data_lru.front();
          auto oit = idx->data.find(old);
          if (oit != idx->data.end()) idx->erase_key(oit);
        }
      }
      KeyNode& node = it->second;
      for (int32_t e = 0; e < n_entries; ++e) {  // inner lru.Add
        int at = -1;
        for (int k = 0; k < node.count; ++k)
          if (same_entry(node.e[k], entries[e])) {
            at = k;
            break;
          }
        kvb_pod_entry_t v = entries[e];
        v.speculative = v.speculative ? 1 : 0;
        if (at >= 0) {  // move to newest
          for (int k = at; k + 1 < node.count; ++k) node.e[k] = node.e[k + 1];
          node.e[node.count - 1] = v;
        } else {
          if (node.count == idx->pods_per_key) {  // evict oldest pod entry
            for (int k = 0; k + 1 < node.count; ++k) node.e[k] = node.e[k + 1];
            node.count--;
          }
          node.e[node.count++] = v;
        }
      }
      idx->dirty.push_back(rk);
    }
    return KVB_OK;
  });
}

static void evict_from_request_key(kvb_index* idx, uint64_t rk, const kvb_pod_entry_t* entries, int32_t n) {
  auto it = idx->data.find(rk);  // in_memory.go:260-264 (Get refreshes recency)
  if (it == idx->data.end()) return;
  idx->touch(it);
  KeyNode& node = it->second;
  for (int32_t e = 0; e < n; ++e) {
    for (int k = 0; k < node.count; ++k)
      if (same_entry(node.e[k], entries[e])) {
        for (int q = k; q + 1 < node.count; ++q) node.e[q] = node.e[q + 1];
        node.count--;
        break;
      }
  }
  if (node.count == 0) {
    idx->erase_key(it);  // in_memory.go:279-292
  } else {
    idx->dirty.push_back(rk);
  }
}

int kvb_index_evict(kvb_index_t* idx, uint64_t key, int key_type, const kvb_pod_entry_t* entries, int32_t n_entries) {
  return kvb::guarded([&]() -> int {
    KVB_REQUIRE(idx != nullptr, "index is NULL");
    KVB_REQUIRE(n_entries > 0 && entries, "no entries provided for eviction from index");  // in_memory.go:230-232
    std::lock_guard<std::mutex> lk(idx->mu);
    if (key_type == KVB_KEY_ENGINE) {
      auto it = idx->eng.find(key);
      if (it == idx->eng.end()) return KVB_OK;  // nothing to evict (in_memory.go:238-242)
      std::vector<uint64_t> rks = it->second.rks;
      for (uint64_t rk : rks) evict_from_request_key(idx, rk, entries, n_entries);
      it = idx->eng.find(key);
      if (it != idx->eng.end()) {
        idx->eng_lru.erase(it->second.lru);
        idx->eng.erase(it);
      }
      return KVB_OK;
    }
    if (key_type == KVB_KEY_REQUEST) {
      evict_from_request_key(idx, key, entries, n_entries);
      return KVB_OK;
    }
    set_error("unknown key type: %d", key_type);  // in_memory.go:252-254
    return KVB_ERR_INVALID;
  });
}

int kvb_index_get_request_key(kvb_index_t* idx, uint64_t engine_key, uint64_t* out) {
  return kvb::guarded([&]() -> int {
    KVB_REQUIRE(idx && out, "NULL argument");
    std::lock_guard<std::mutex> lk(idx->mu);
    auto it = idx->eng.find(engine_key);
    if (it == idx->eng.end() || it->second.rks.empty()) {  // in_memory.go:299-302
      set_error("engine key not found: %llu", (unsigned long long)engine_key);
      *out = 0;
      return KVB_ERR_NOTFOUND;
    }
    idx->touch_eng(it);
    *out = it->second.rks.back();
    return KVB_OK;
  });
}

int64_t kvb_index_num_keys(kvb_index_t* idx) {
  if (!idx) return 0;
  std::lock_guard<std::mutex> lk(idx->mu);
  return (int64_t)idx->data.size();
}

int kvb_index_host_peek(kvb_index_t* idx, uint64_t request_key, kvb_pod_entry_t* out_entries, int32_t cap) {
  return kvb::guarded([&]() -> int {
    if (!idx) return -1;
    std::lock_guard<std::mutex> lk(idx->mu);
    auto it = idx->data.find(request_key);
    if (it == idx->data.end()) return -1;
    const int n = std::min<int>(it->second.count, cap);
    for (int i = 0; i < n; ++i) out_entries[i] = it->second.e[i];
    return it->second.count;
  });
}

int kvb_index_flush(kvb_index_t* idx, void* stream) {
  return kvb::guarded([&]() -> int {
    (void)stream;
    KVB_REQUIRE(idx != nullptr, "index is NULL");
    std::lock_guard<std::mutex> lk(idx->mu);
    DeviceGuard g(idx->device);
    return idx->flush_locked();
  });
}

int kvb_index_lookup(kvb_index_t* idx, const uint64_t* keys, int64_t n, const uint16_t* pod_filter, int32_t n_filter,
                     int32_t* out_counts, kvb_pod_entry_t* out_entries, int64_t* out_cut) {
  return kvb::guarded([&]() -> int {
    KVB_REQUIRE(idx != nullptr, "index is NULL");
    KVB_REQUIRE(n > 0 && keys, "no requestKeys provided for lookup");  // in_memory.go:110-112
    KVB_REQUIRE(out_counts && out_entries && out_cut, "NULL output");
    std::lock_guard<std::mutex> lk(idx->mu);
    DeviceGuard g(idx->device);
    int rc = idx->flush_locked();
    if (rc) return rc;
    const uint32_t* filt = nullptr;
    rc = idx->set_filter(pod_filter, n_filter, &filt);
    if (rc) return rc;
    const size_t o_keys = 0, o_cnt = (size_t)n * 8, o_ent = o_cnt + (((size_t)n * 4 + 255) & ~size_t(255));
    const size_t total = o_ent + (size_t)n * kMaxEnt * 4;
    rc = idx->ensure_scratch(total, total);
    if (rc) return rc;
    std::memcpy(idx->h_scratch, keys, (size_t)n * 8);
    cudaStream_t s = idx->stream;
    KVB_CUDA_TRY(cudaMemcpyAsync(idx->d_scratch + o_keys, idx->h_scratch, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    const int threads = 128;
    index_lookup_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(
        idx->table, idx->slots - 1, reinterpret_cast<const uint64_t*>(idx->d_scratch + o_keys), n, filt,
        reinterpret_cast<int32_t*>(idx->d_scratch + o_cnt), reinterpret_cast<uint32_t*>(idx->d_scratch + o_ent));
    KVB_CUDA_TRY(cudaGetLastError());
    count_launch();
    KVB_CUDA_TRY(cudaMemcpyAsync(idx->h_scratch + o_cnt, idx->d_scratch + o_cnt, total - o_cnt, cudaMemcpyDeviceToHost, s));
    KVB_CUDA_TRY(cudaStreamSynchronize(s));
    const int32_t* cnt = reinterpret_cast<const int32_t*>(idx->h_scratch + o_cnt);
    const uint32_t* ent = reinterpret_cast<const uint32_t*>(idx->h_scratch + o_ent);
    int64_t cut = n;
    for (int64_t i = 0; i < n; ++i) {
      if (cut < n) {  // after the cut nothing is looked at (in_memory.go:121-124 returns)
        out_counts[i] = -1;
        continue;
      }
      if (cnt[i] == -1) {
        out_counts[i] = -1;
        continue;
      }
      // data.Get refreshes the outer LRU for every key that is found (in_memory.go:120)
      auto it = idx->data.find(keys[i]);
      if (it != idx->data.end()) idx->touch(it);
      if (cnt[i] == -2) {
        cut = i;
        out_counts[i] = 0;
        continue;
      }
      out_counts[i] = cnt[i];
      for (int e = 0; e < cnt[i]; ++e) {
        const uint32_t v = ent[i * kMaxEnt + e];
        out_entries[i * kMaxEnt + e].pod = (uint16_t)(v & 0xffffu);
        out_entries[i * kMaxEnt + e].tier = (uint8_t)((v >> 16) & 0xffu);
        out_entries[i * kMaxEnt + e].speculative = (uint8_t)((v >> 24) & 1u);
      }
    }
    *out_cut = cut;
    return KVB_OK;
  });
}

// keys already on the device at d_keys / d_koff (inside idx->d_scratch or elsewhere)
static int score_device(kvb_index* idx, const uint64_t* d_keys, const int64_t* d_koff, int32_t n_prompts,
                        const uint32_t* filt, uint8_t* d_found, int32_t* d_n, uint16_t* d_pods, double* d_scores) {
  const int64_t grid = ((int64_t)n_prompts + kScoreWarps - 1) / kScoreWarps;
  index_score_kernel<<<(unsigned)grid, kScoreWarps * 32, 0, idx->stream>>>(
      idx->table, idx->slots - 1, d_keys, d_koff, n_prompts, filt, idx->tier_w, d_n, d_pods, d_scores, d_found);
  KVB_CUDA_TRY(cudaGetLastError());
  count_launch();
  return KVB_OK;
}

static int score_common(kvb_index* idx, const uint64_t* keys_host, const int64_t* key_off, int32_t n_prompts,
                        const uint32_t* tokens, const int64_t* prompt_off, const uint64_t* parents,
                        int32_t block_size, const uint8_t* extra, const int64_t* extra_off,
                        const uint16_t* pod_filter, int32_t n_filter, int32_t flags, int32_t* out_n,
                        uint16_t* out_pods, double* out_scores) {
  std::lock_guard<std::mutex> lk(idx->mu);
  DeviceGuard g(idx->device);
  int rc = idx->flush_locked();
  if (rc) return rc;
  const uint32_t* filt = nullptr;
  rc = idx->set_filter(pod_filter, n_filter, &filt);
  if (rc) return rc;
  const bool from_tokens = tokens != nullptr;
  std::vector<int64_t> koff_local;
  if (from_tokens) {
    koff_local.resize((size_t)n_prompts + 1);
    koff_local[0] = 0;
    for (int32_t p = 0; p < n_prompts; ++p) {
      KVB_REQUIRE(prompt_off[p + 1] >= prompt_off[p], "prompt_off not monotonic at %d", p);
      koff_local[p + 1] = koff_local[p] + (prompt_off[p + 1] - prompt_off[p]) / block_size;
    }
    key_off = koff_local.data();
  }
  const int64_t total_keys = key_off[n_prompts] - key_off[0];
  const int64_t total_tok = from_tokens ? prompt_off[n_prompts] - prompt_off[0] : 0;
  const int64_t extra_bytes = (from_tokens && extra_off) ? extra_off[total_keys] : 0;
  const bool touch = (flags & KVB_SCORE_TOUCH_LRU) != 0;
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  // device/host scratch layout (same offsets on both sides)
  size_t o = 0;
  const size_t o_keys = o;   o += al((size_t)total_keys * 8);
  const size_t o_koff = o;   o += al(((size_t)n_prompts + 1) * 8);
  const size_t o_n = o;      o += al((size_t)n_prompts * 4);
  const size_t o_pods = o;   o += al((size_t)n_prompts * kMaxEnt * 2);
  const size_t o_sc = o;     o += al((size_t)n_prompts * kMaxEnt * 8);
  const size_t o_found = o;  o += touch ? al((size_t)total_keys) : 0;
  const size_t o_tok = o;    o += from_tokens ? al((size_t)total_tok * 4) : 0;
  const size_t o_poff = o;   o += from_tokens ? al(((size_t)n_prompts + 1) * 8) : 0;
  const size_t o_par = o;    o += from_tokens ? al((size_t)n_prompts * 8) : 0;
  const size_t o_eoff = o;   o += (from_tokens && extra_off) ? al(((size_t)total_keys + 1) * 8) : 0;
  const size_t o_ext = o;    o += (from_tokens && extra_off) ? al((size_t)extra_bytes) : 0;
  rc = idx->ensure_scratch(o, o);
  if (rc) return rc;
  uint8_t* H = idx->h_scratch;
  uint8_t* D = idx->d_scratch;
  cudaStream_t s = idx->stream;
  // stage inputs in pinned memory, one H2D per array
  int64_t* h_koff = reinterpret_cast<int64_t*>(H + o_koff);
  for (int32_t p = 0; p <= n_prompts; ++p) h_koff[p] = key_off[p] - key_off[0];
  KVB_CUDA_TRY(cudaMemcpyAsync(D + o_koff, H + o_koff, ((size_t)n_prompts + 1) * 8, cudaMemcpyHostToDevice, s));
  if (from_tokens) {
    int64_t* h_poff = reinterpret_cast<int64_t*>(H + o_poff);
    for (int32_t p = 0; p <= n_prompts; ++p) h_poff[p] = prompt_off[p] - prompt_off[0];
    std::memcpy(H + o_par, parents, (size_t)n_prompts * 8);
    // tokens go straight from the caller's buffer: pinned memory (kvb_host_alloc / cudaHostAlloc / registered) is read
    // by the copy engine in place, pageable memory is staged by the driver (measured faster than staging it here:
    // 0.32 vs 0.36-0.42 ms per 4 MB batch, tools/ab_stage.py); prompt_off and parents are adjacent in the scratch
    KVB_CUDA_TRY(cudaMemcpyAsync(D + o_tok, tokens + prompt_off[0], (size_t)total_tok * 4, cudaMemcpyHostToDevice, s));
    KVB_CUDA_TRY(cudaMemcpyAsync(D + o_poff, H + o_poff, (o_par + (size_t)n_prompts * 8) - o_poff,
                                 cudaMemcpyHostToDevice, s));
    if (extra_off) {
      std::memcpy(H + o_eoff, extra_off, ((size_t)total_keys + 1) * 8);
      if (extra_bytes) std::memcpy(H + o_ext, extra, (size_t)extra_bytes);
      KVB_CUDA_TRY(cudaMemcpyAsync(D + o_eoff, H + o_eoff, (o_ext + (size_t)extra_bytes) - o_eoff,
                                   cudaMemcpyHostToDevice, s));
    }
    if (total_keys > 0) {
      rc = launch_hash_blocks(reinterpret_cast<uint32_t*>(D + o_tok), reinterpret_cast<int64_t*>(D + o_poff),
                              reinterpret_cast<uint64_t*>(D + o_par), n_prompts, block_size,
                              extra_off ? D + o_ext : nullptr,
                              extra_off ? reinterpret_cast<int64_t*>(D + o_eoff) : nullptr,
                              reinterpret_cast<uint64_t*>(D + o_keys), reinterpret_cast<int64_t*>(D + o_koff), s);
      if (rc) return rc;
    }
  } else if (total_keys > 0) {
    std::memcpy(H + o_keys, keys_host + key_off[0], (size_t)total_keys * 8);
    KVB_CUDA_TRY(cudaMemcpyAsync(D + o_keys, H + o_keys, (size_t)total_keys * 8, cudaMemcpyHostToDevice, s));
  }
  rc = score_device(idx, reinterpret_cast<uint64_t*>(D + o_keys), reinterpret_cast<int64_t*>(D + o_koff), n_prompts,
                    filt, touch ? D + o_found : nullptr, reinterpret_cast<int32_t*>(D + o_n),
                    reinterpret_cast<uint16_t*>(D + o_pods), reinterpret_cast<double*>(D + o_sc));
  if (rc) return rc;
  // results: n | pods | scores (| found) are adjacent: one D2H
  const size_t back_end = touch ? o_found + (size_t)total_keys : o_sc + (size_t)n_prompts * kMaxEnt * 8;
  KVB_CUDA_TRY(cudaMemcpyAsync(H + o_n, D + o_n, back_end - o_n, cudaMemcpyDeviceToHost, s));
  if (touch && from_tokens && total_keys > 0)
    KVB_CUDA_TRY(cudaMemcpyAsync(H + o_keys, D + o_keys, (size_t)total_keys * 8, cudaMemcpyDeviceToHost, s));
  KVB_CUDA_TRY(cudaStreamSynchronize(s));
  std::memcpy(out_n, H + o_n, (size_t)n_prompts * 4);
  std::memcpy(out_pods, H + o_pods, (size_t)n_prompts * kMaxEnt * 2);
  std::memcpy(out_scores, H + o_sc, (size_t)n_prompts * kMaxEnt * 8);
  if (touch) {
    const uint64_t* hk = reinterpret_cast<const uint64_t*>(H + o_keys);
    const uint8_t* hf = H + o_found;
    for (int64_t i = 0; i < total_keys; ++i) {
      if (!hf[i]) continue;
      auto it = idx->data.find(hk[i]);
      if (it != idx->data.end()) idx->touch(it);
    }
  }
  return KVB_OK;
}

int kvb_index_score_batch(kvb_index_t* idx, const uint64_t* keys, const int64_t* key_off, int32_t n_prompts,
                          const uint16_t* pod_filter, int32_t n_filter, int32_t flags, int32_t* out_n,
                          uint16_t* out_pods, double* out_scores) {
  return kvb::guarded([&]() -> int {
    KVB_REQUIRE(idx != nullptr, "index is NULL");
    KVB_REQUIRE(n_prompts >= 0, "negative prompt count");
    if (n_prompts == 0) return KVB_OK;
    KVB_REQUIRE(key_off && out_n && out_pods && out_scores, "NULL argument");
    KVB_REQUIRE(keys || key_off[n_prompts] == key_off[0], "keys is NULL");
    return score_common(idx, keys, key_off, n_prompts, nullptr, nullptr, nullptr, 0, nullptr, nullptr, pod_filter,
                        n_filter, flags, out_n, out_pods, out_scores);
  });
}

int kvb_index_score_tokens_batch(kvb_index_t* idx, const uint32_t* tokens, const int64_t* prompt_off,
                                 const uint64_t* parents, int32_t n_prompts, int32_t block_size, const uint8_t* extra,
                                 const int64_t* extra_off, const uint16_t* pod_filter, int32_t n_filter,
                                 int32_t flags, int32_t* out_n, uint16_t* out_pods, double* out_scores) {
  return kvb::guarded([&]() -> int {
    KVB_REQUIRE(idx != nullptr, "index is NULL");
    KVB_REQUIRE(block_size > 0, "blockSize must be greater than 0, got %d", block_size);
    KVB_REQUIRE(n_prompts >= 0, "negative prompt count");
    if (n_prompts == 0) return KVB_OK;
    KVB_REQUIRE(tokens && prompt_off && parents && out_n && out_pods && out_scores, "NULL argument");
    return score_common(idx, nullptr, nullptr, n_prompts, tokens, prompt_off, parents, block_size, extra, extra_off,
                        pod_filter, n_filter, flags, out_n, out_pods, out_scores);
  });
}

}  // extern "C"
