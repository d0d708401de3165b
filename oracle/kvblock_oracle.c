/* kvblock_oracle.c — plain-C restatement of the reference's kvblock read path, for TIMING the CPU baseline
 * (the Python oracle in kvblock_oracle.py is the parity checker; this file is cross-checked against it in
 * tests/test_oracle_c.py).  TEST INFRASTRUCTURE ONLY — never linked into libkvb.so.
 *
 * Restates (paths relative to /root/reference):
 *   pkg/kvcache/kvblock/token_processor.go:123-135  hash(): CBOR-canonical marshal of [parent, tokens, extra], FNV-64a
 *   pkg/kvcache/kvblock/token_processor.go:139-205  prefixHashes / chunkTokens / TokensToKVBlockKeys
 *   pkg/kvcache/kvblock/in_memory.go:107-148        Lookup (hash-map probe per key)
 *   pkg/kvcache/kvblock_scorer.go:91-154            LongestPrefixScorer.Score (float64 adds in key order)
 * The Go code allocates a []interface{} and a CBOR buffer per block and takes two mutex-guarded LRU gets per key;
 * this C version does neither, so it is a FASTER-than-reference baseline (conservative for any speed-up claim).
 * Third-party pieces restated: fxamacker/cbor v2.7.0 canonical heads (RFC 8949 4.2.1), Go hash/fnv New64a.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FNV_OFFSET 0xcbf29ce484222325ull
#define FNV_PRIME 0x100000001b3ull

uint64_t kvo_fnv64a(const uint8_t* p, size_t n) {
  uint64_t h = FNV_OFFSET;
  for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * FNV_PRIME;
  return h;
}

static size_t put_head(uint8_t* o, uint8_t major, uint64_t v) {
  if (v < 24) { o[0] = major | (uint8_t)v; return 1; }
  if (v < 0x100ull) { o[0] = major | 24; o[1] = (uint8_t)v; return 2; }
  if (v < 0x10000ull) { o[0] = major | 25; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)v; return 3; }
  if (v < 0x100000000ull) { o[0] = major | 26; for (int i = 0; i < 4; ++i) o[1 + i] = (uint8_t)(v >> (24 - 8 * i)); return 5; }
  o[0] = major | 27; for (int i = 0; i < 8; ++i) o[1 + i] = (uint8_t)(v >> (56 - 8 * i)); return 9;
}

/* one block: marshal then hash, like the reference (payload buffer must hold 11 + 5*bs + extra_len bytes) */
static uint64_t block_hash(uint8_t* buf, uint64_t parent, const uint32_t* tok, int bs, const uint8_t* extra, int64_t extra_len) {
  size_t n = 0;
  buf[n++] = 0x83;
  n += put_head(buf + n, 0x00, parent);
  n += put_head(buf + n, 0x80, (uint64_t)bs);
  for (int j = 0; j < bs; ++j) n += put_head(buf + n, 0x00, tok[j]);
  if (extra_len > 0) { memcpy(buf + n, extra, (size_t)extra_len); n += (size_t)extra_len; }
  else buf[n++] = 0xf6;
  return kvo_fnv64a(buf, n);
}

uint64_t kvo_init_hash(uint64_t seed_hash, const uint8_t* model, size_t len) {
  uint8_t* buf = (uint8_t*)malloc(32 + len);
  size_t n = 0;
  buf[n++] = 0x83;
  n += put_head(buf + n, 0x00, seed_hash);
  buf[n++] = 0xf6;
  n += put_head(buf + n, 0x60, len);
  memcpy(buf + n, model, len);
  n += len;
  uint64_t h = kvo_fnv64a(buf, n);
  free(buf);
  return h;
}

/* batched TokensToKVBlockKeys; key_off precomputed by the caller; threads <= 0 -> all cores */
void kvo_hash_batch(const uint32_t* tokens, const int64_t* prompt_off, const uint64_t* parents, int32_t n_prompts,
                    int32_t bs, const uint8_t* extra, const int64_t* extra_off, uint64_t* out_keys,
                    const int64_t* key_off, int threads) {
#ifdef _OPENMP
  omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
#endif
  int64_t max_extra = 0;
  if (extra_off) for (int64_t k = 0; k < key_off[n_prompts]; ++k) {
    int64_t l = extra_off[k + 1] - extra_off[k];
    if (l > max_extra) max_extra = l;
  }
#pragma omp parallel
  {
    uint8_t* buf = (uint8_t*)malloc((size_t)(16 + 5 * (int64_t)bs + max_extra));
#pragma omp for schedule(dynamic, 4)
    for (int32_t p = 0; p < n_prompts; ++p) {
      const uint32_t* tk = tokens + prompt_off[p];
      int64_t nblk = (prompt_off[p + 1] - prompt_off[p]) / bs, k0 = key_off[p];
      uint64_t parent = parents[p];
      for (int64_t i = 0; i < nblk; ++i) {
        int64_t e0 = extra_off ? extra_off[k0 + i] : 0, e1 = extra_off ? extra_off[k0 + i + 1] : 0;
        parent = block_hash(buf, parent, tk + i * bs, bs, extra ? extra + e0 : NULL, e1 - e0);
        out_keys[k0 + i] = parent;
      }
    }
    free(buf);
  }
}

/* ---------------- index (read-only restatement for timing): open addressing, <= 13 entries per key -------- */
typedef struct { uint64_t key; uint8_t used, count; uint16_t pod[13]; uint8_t tier[13]; } kvo_slot;
typedef struct { kvo_slot* s; uint64_t mask; } kvo_index;

static uint64_t mix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k; }

kvo_index* kvo_index_new(uint64_t slots_pow2) {
  kvo_index* ix = (kvo_index*)malloc(sizeof(kvo_index));
  ix->s = (kvo_slot*)calloc(slots_pow2, sizeof(kvo_slot));
  ix->mask = slots_pow2 - 1;
  return ix;
}
void kvo_index_free(kvo_index* ix) { if (ix) { free(ix->s); free(ix); } }

/* Add(nil, keys, [entry]) for distinct or repeated keys (entries appended, duplicates ignored, cap 13) */
void kvo_index_add(kvo_index* ix, const uint64_t* keys, int64_t n, uint16_t pod, uint8_t tier) {
  for (int64_t i = 0; i < n; ++i) {
    uint64_t s = mix(keys[i]) & ix->mask;
    while (ix->s[s].used && ix->s[s].key != keys[i]) s = (s + 1) & ix->mask;
    kvo_slot* sl = &ix->s[s];
    sl->used = 1; sl->key = keys[i];
    int dup = 0;
    for (int e = 0; e < sl->count; ++e) if (sl->pod[e] == pod && sl->tier[e] == tier) dup = 1;
    if (!dup && sl->count < 13) { sl->pod[sl->count] = pod; sl->tier[sl->count] = tier; sl->count++; }
  }
}

/* Lookup + Score for a batch of prompts (no pod filter): per prompt up to 13 (pod, score) pairs */
void kvo_score_batch(const kvo_index* ix, const uint64_t* keys, const int64_t* key_off, int32_t n_prompts,
                     const double* tier_w, int32_t* out_n, uint16_t* out_pods, double* out_scores, int threads) {
#ifdef _OPENMP
  omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
#endif
#pragma omp parallel for schedule(dynamic, 8)
  for (int32_t p = 0; p < n_prompts; ++p) {
    uint16_t pods[13]; double sc[13]; int act[13]; int np_ = 0;
    int64_t k0 = key_off[p], nk = key_off[p + 1] - k0;
    for (int64_t i = 0; i < nk; ++i) {
      uint64_t s = mix(keys[k0 + i]) & ix->mask;
      const kvo_slot* sl = NULL;
      while (ix->s[s].used) { if (ix->s[s].key == keys[k0 + i]) { sl = &ix->s[s]; break; } s = (s + 1) & ix->mask; }
      if (i == 0) {
        if (!sl) break;
        for (int e = 0; e < sl->count; ++e) {
          double w = tier_w[sl->tier[e]]; int f = -1;
          for (int q = 0; q < np_; ++q) if (pods[q] == sl->pod[e]) f = q;
          if (f < 0) { pods[np_] = sl->pod[e]; sc[np_] = w; act[np_] = 1; np_++; }
          else if (w > sc[f]) sc[f] = w;
        }
        continue;
      }
      int any = 0;
      for (int q = 0; q < np_; ++q) {
        if (!act[q]) continue;
        int hit = 0; double wm = 0.0;
        if (sl) for (int e = 0; e < sl->count; ++e) if (sl->pod[e] == pods[q]) { double w = tier_w[sl->tier[e]]; if (!hit || w > wm) wm = w; hit = 1; }
        if (hit) { sc[q] += wm; any = 1; } else act[q] = 0;
      }
      if (!any) break;
    }
    out_n[p] = np_;
    for (int q = 0; q < np_; ++q) { out_pods[p * 13 + q] = pods[q]; out_scores[p * 13 + q] = sc[q]; }
  }
}
