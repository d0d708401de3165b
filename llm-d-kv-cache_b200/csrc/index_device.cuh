// index_device.cuh — device-side definitions of the block index shared by index.cu (table kernels, scoring) and
// hash_kernels.cu (the fused tokens -> scores kernel): bucket layout, probe, and the longest-prefix walk of one warp.
#pragma once

namespace kvb {

constexpr int kMaxEnt = KVB_INDEX_MAX_PODS_PER_KEY;  // 13
constexpr uint32_t kEmpty = 0, kFull = 1, kTomb = 2, kBusy = 3;
constexpr uint32_t kNoSlot = 0xffffffffu;

struct __align__(64) Bucket {
  uint64_t key;
  uint32_t meta;  // bits 0-1 state, bits 8-15 entry count
  uint32_t ent[kMaxEnt];
};
static_assert(sizeof(Bucket) == 64, "bucket must be one 64 B line");

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

__device__ __host__ __forceinline__ int64_t probe(const Bucket* __restrict__ table, uint64_t mask, uint64_t key) {
  uint64_t slot = mix64(key) & mask;
  for (uint64_t probes = 0; probes <= mask; ++probes, slot = (slot + 1) & mask) {
    const uint32_t st = table[slot].meta & 3u;
    if (st == kEmpty) return -1;
    if (st == kFull && table[slot].key == key) return (int64_t)slot;
  }
  return -1;
}

__device__ __host__ __forceinline__ bool pod_allowed(const uint32_t* __restrict__ filter_bits, uint32_t pod) {
  return filter_bits == nullptr || ((filter_bits[pod >> 5] >> (pod & 31)) & 1u);
}


#ifndef KVB_HOST_SIM
// LongestPrefixScorer.Score (kvblock_scorer.go:106-154) for ONE prompt by ONE warp, fed 32 keys at a time.
//   tile(): phase A — 32 independent probes (random 64 B reads), found buckets staged in shared memory and re-stamped
//           (Lookup's data.Get refreshes every key it finds, in_memory.go:119-120);
//           phase B — the serial walk over the tile's keys: bucket entries sit on lanes 16..28, the reported pods
//           ("owners", in the order they appeared at key 0) on lanes 0..12, ONE match.any per key pairs every owner with
//           the entries that carry its pod; scores are float64 sums in key order, bit-identical to the Go loop.
struct ScoreWalker {
  bool active = false;  // this lane owns a pod that is still on the consecutive prefix
  bool owner = false;   // this lane owns a pod that appeared at key 0 (it is reported)
  uint32_t my_pod = 0xffffffffu;
  double score = 0.0;
  bool chain_alive = true;

  __device__ __forceinline__ void tile(const Bucket* __restrict__ table, uint64_t mask, Bucket* __restrict__ tile_smem,
                                       uint64_t key, bool have_key, int64_t base, int in_tile,
                                       const uint32_t* __restrict__ filter_bits, const double* __restrict__ tier_w,
                                       unsigned long long* __restrict__ ts, unsigned long long stamp) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int64_t slot = -1;
    if (have_key) slot = probe(table, mask, key);
    if (ts != nullptr && slot >= 0) atomicMax(&ts[slot], stamp);
    if (slot >= 0) {
      const uint4* src = reinterpret_cast<const uint4*>(&table[slot]);
      uint4* dst = reinterpret_cast<uint4*>(&tile_smem[lane]);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = src[q];
    }
    __syncwarp();
    if (!chain_alive) return;
    for (int j = 0; j < in_tile; ++j) {
      const int64_t sj = __shfl_sync(FULL, slot, j);
      bool valid = false;
      uint32_t pod = 0x10000u + lane;  // unique sentinel for lanes without a valid entry
      double w = 0.0;
      if (sj >= 0) {
        const Bucket& b = tile_smem[j];
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
  __device__ __forceinline__ void finish(int64_t p, int32_t* __restrict__ out_n, uint16_t* __restrict__ out_pods,
                                         double* __restrict__ out_scores) const {
    const int lane = threadIdx.x & 31;
    const unsigned om = __ballot_sync(0xffffffffu, owner);
    if (owner) {
      const int pos = __popc(om & ((1u << lane) - 1u));
      out_pods[p * kMaxEnt + pos] = (uint16_t)my_pod;
      out_scores[p * kMaxEnt + pos] = score;
    }
    if (lane == 0) out_n[p] = __popc(om);
  }
};

// arguments of the chain kernel of hash_kernels.cu (hash only, or fused tokens -> keys -> lookup -> scores)
struct ChainArgs {
  const uint32_t* tokens;     // device memory, or pinned host memory through its device alias (zero copy)
  const uint32_t* tokens_lo;  // bounds of the caller's token array: 16 B granules are only read inside them
  const uint32_t* tokens_hi;
  int32_t single;             // one prompt: its offsets and parent travel in the arguments below
  int64_t single_tokens;
  uint64_t single_parent;
  unsigned* done_counter;     // nullable: completion word (fused form), see chain_signal_done
  unsigned done_target;
  unsigned long long* done_flag_host;
  unsigned long long done_value;
  const int64_t* prompt_off;
  const uint64_t* parents;
  const uint8_t* extra;
  const int64_t* extra_off;
  uint64_t* out_keys;         // nullable in the fused form
  const int64_t* key_off;
  // fused scoring (SCORE)
  const Bucket* table;
  uint64_t mask;
  const uint32_t* filter_bits;
  const double* tier_w;
  int32_t* out_n;
  uint16_t* out_pods;
  double* out_scores;
  int32_t score_min_batch;    // keys the scorer waits for before it probes (1..32)
  unsigned long long* ts;     // nullable: stamp found keys
  unsigned long long stamp_base;
};

// one launch: tokens -> keys -> lookup -> scores.  false if the block size has no chain kernel (hash and score separately)
bool launch_chain_score(const ChainArgs& a, int32_t n_prompts, int32_t block_size, cudaStream_t s);
#endif  // !KVB_HOST_SIM

}  // namespace kvb
