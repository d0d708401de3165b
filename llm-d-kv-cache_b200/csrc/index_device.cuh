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
//   tile(): one key per lane.  Each lane probes its key (a random 64 B read), re-stamps the bucket it finds (Lookup's
//   data.Get refreshes every key it finds, in_memory.go:119-120) and keeps the bucket in registers.  Key 0 names the
//   reported pods ("owners", distinct pods in the order their entries appear, lanes 0..n_own-1).  Then, per owner, every
//   lane looks for that pod among ITS key's entries (max weight over the pod's tiers) and one ballot finds the first key
//   of the tile without it; the owner's lane adds the weights of the keys before that one in key order — float64,
//   the same sum as the Go loop — out of a [8 owners][32 keys] matrix in shared memory (the caller's 2 KiB tile buffer).
//   The walk is parallel over keys; only the additions are serial (round 1/2's walk went key by key: ~1000 cycles per key).
struct ScoreWalker {
  bool active = false;  // this lane owns a pod that is still on the consecutive prefix
  bool owner = false;   // this lane owns a pod that appeared at key 0 (it is reported)
  uint32_t my_pod = 0xffffffffu;
  double score = 0.0;
  bool chain_alive = true;
  int n_own = 0;

  __device__ __forceinline__ void tile(const Bucket* __restrict__ table, uint64_t mask, Bucket* __restrict__ tile_smem,
                                       uint64_t key, bool have_key, int64_t base, int in_tile,
                                       const uint32_t* __restrict__ filter_bits, const double* __restrict__ tier_w,
                                       unsigned long long* __restrict__ ts, unsigned long long stamp) {
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(kMaxEnt == 13, "bucket words are unpacked by hand below");
    const int lane = threadIdx.x & 31;
    int64_t slot = -1;
    if (have_key) slot = probe(table, mask, key);
    if (ts != nullptr && slot >= 0) atomicMax(&ts[slot], stamp);
    if (!chain_alive) return;
    uint32_t ent[kMaxEnt];
    int cnt = 0;
    if (slot >= 0) {
      const uint4* src = reinterpret_cast<const uint4*>(&table[slot]);
      const uint4 q0 = src[0], q1 = src[1], q2 = src[2], q3 = src[3];  // key | meta ent0 | ent1-4 | ent5-8 | ent9-12
      cnt = min((int)((q0.z >> 8) & 0xffu), kMaxEnt);
      ent[0] = q0.w;
      ent[1] = q1.x, ent[2] = q1.y, ent[3] = q1.z, ent[4] = q1.w;
      ent[5] = q2.x, ent[6] = q2.y, ent[7] = q2.z, ent[8] = q2.w;
      ent[9] = q3.x, ent[10] = q3.y, ent[11] = q3.z, ent[12] = q3.w;
    } else {
#pragma unroll
      for (int e = 0; e < kMaxEnt; ++e) ent[e] = 0u;
    }
    if (base == 0) {
      // owners: distinct pods of key 0's entries that pass the pod filter, in entry order (fillMaxWeights' key set)
      const int cnt0 = __shfl_sync(FULL, cnt, 0);
      uint32_t v = 0u;
#pragma unroll
      for (int e = 0; e < kMaxEnt; ++e) {
        const uint32_t x = __shfl_sync(FULL, ent[e], 0);
        if (lane == e) v = x;
      }
      const bool valid = lane < cnt0 && pod_allowed(filter_bits, v & 0xffffu);
      const uint32_t pod = valid ? (v & 0xffffu) : (0x10000u + lane);  // unique sentinel for lanes without an entry
      const unsigned group = __match_any_sync(FULL, pod);
      const bool leader = valid && (group & ((1u << lane) - 1u)) == 0u;  // first entry of its pod
      const unsigned lm = __ballot_sync(FULL, leader);
      n_own = __popc(lm);
      const unsigned src = __fns(lm, 0, lane + 1);  // lane of the (lane + 1)-th leader
      const uint32_t p2 = __shfl_sync(FULL, pod, (int)(src & 31u));
      owner = active = lane < n_own;
      my_pod = owner ? p2 : 0xffffffffu;
      score = 0.0;
    }
    double* wmat = reinterpret_cast<double*>(tile_smem);  // [8][32], sizeof(Bucket[32]) == 2048
    const unsigned in_mask = in_tile >= 32 ? FULL : ((1u << in_tile) - 1u);
    for (int o0 = 0; o0 < n_own; o0 += 8) {
      const int oc = min(8, n_own - o0);
      int my_f = in_tile;
      for (int oo = 0; oo < oc; ++oo) {
        const uint32_t po = __shfl_sync(FULL, my_pod, o0 + oo);
        if (__shfl_sync(FULL, active ? 1 : 0, o0 + oo) == 0) continue;  // this pod already left the prefix (uniform)
        bool hit = false;
        double best = 0.0;
#pragma unroll
        for (int e = 0; e < kMaxEnt; ++e) {
          if (e < cnt && (ent[e] & 0xffffu) == po) {
            const double w = tier_w[(ent[e] >> 16) & 0xffu];
            if (!hit || w > best) best = w;  // a pod on two tiers counts with the larger weight
            hit = true;
          }
        }
        const unsigned absent = __ballot_sync(FULL, !hit) & in_mask;
        wmat[oo * 32 + lane] = best;
        if (lane == o0 + oo) my_f = absent ? __ffs((int)absent) - 1 : in_tile;
      }
      __syncwarp();
      if (active && lane >= o0 && lane < o0 + oc) {
        const double* row = wmat + (lane - o0) * 32;
        for (int j = 0; j < my_f; ++j) score += row[j];  // float64, key order: same sum as the Go loop
        if (my_f < in_tile) active = false;
      }
      __syncwarp();
    }
    if (!__any_sync(FULL, active)) chain_alive = false;
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
// the same for small batches by the table ("spec") kernel: any block size, total_keys = key_off[n_prompts];
// false = not applicable (batch too large or switched off), true = launched or failed (*rc_out)
// h_*: host-readable copies of prompt_off / key_off / parents
struct SpecScratch;
bool launch_spec_score(const ChainArgs& a, int32_t n_prompts, int32_t block_size, int64_t total_keys, cudaStream_t s,
                       int* rc_out, const int64_t* h_prompt_off, const int64_t* h_key_off, const uint64_t* h_parents,
                       SpecScratch* scratch);
#endif  // !KVB_HOST_SIM

}  // namespace kvb
