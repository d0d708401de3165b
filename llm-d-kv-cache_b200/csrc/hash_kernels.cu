// hash_kernels.cu — chained kvblock prefix hash on the device.
//
// Replaces chunkedTokenDatabase.{hash,prefixHashes,TokensToKVBlockKeys}
// (pkg/kvcache/kvblock/token_processor.go:123-205):
//     key_i = FNV64a( 0x83 | U(parent) | ARR(chunk_i) | X(extra_i) ),  parent = key_{i-1}
// with U/ARR the RFC 8949 shortest-form heads the reference gets from fxamacker/cbor v2.7.0
// CanonicalEncOptions (token_processor.go:97).  The payload never reaches global memory: token bytes
// are staged in a per-lane strip of shared memory and folded from registers.
//
// Parallelism: FNV-1a is a serial byte chain and keys chain across blocks, so one prompt is one serial
// chain: one LANE per prompt (32 chains per warp), parallel only across prompts.
//   hash_chain_kernel_2w<BS<=16>  two warps per 32 chains: warp 0 encodes block i+1, warp 1 folds block i
//   hash_chain_kernel<BS>         one warp does both (any block size; A/B switch KVB_HASH_ONE_WARP)
// Bound: issue/latency of a lone warp (xor -> wide multiply, >= 10-12 cycles per payload byte on B200,
// tools/micro/hash_micro.cu), NOT HBM: the only memory traffic is 4 B/token in and 8 B/key out.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kvb_internal.h"

namespace kvb {

constexpr uint64_t kFnvOffset = 0xcbf29ce484222325ull;
constexpr uint64_t kFnvPrime = 0x100000001b3ull;

// FNV-1a state kept as two 32-bit halves.  h * 0x100000001b3 = h * 0x1b3 + (h << 40), so per byte:
//   x = lo ^ b;  lo' = x * 0x1b3 (low half);  hi' = hi * 0x1b3 + mulhi(x, 0x1b3) + (x << 8)
// The serial dependency is xor -> mul.lo on the low half (2 ops per byte); the high half trails it with one
// multiply-add per byte.  (A plain 64-bit multiply compiles to ~5 dependent IMADs + selects per byte.)
struct Fnv {
  uint32_t lo, hi;
};
__device__ __forceinline__ void fold(Fnv& h, uint32_t byte) {
  const uint32_t x = h.lo ^ byte;
  // one IMAD.WIDE gives both halves of x * 0x1b3 with a ZERO addend: written as PTX so the compiler cannot fold the
  // high-half accumulation into the wide multiply's addend, which would chain lo' behind the hi update
  // (LOP3 -> SHL -> IMAD -> IMAD.WIDE per byte instead of LOP3 -> IMAD.WIDE; measured 33 vs ~11 cycles per byte)
  uint32_t lo2, carry;
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, 435;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo2), "=r"(carry) : "r"(x));
  h.hi = h.hi * 0x1b3u + (carry + (x << 8));
  h.lo = lo2;
}
// Four bytes (one little-endian word of the byte stream).  The low half is the serial chain (LOP3 -> IMAD.WIDE per
// byte).  The high half is  hi' = hi*M^4 + c0*M^3 + c1*M^2 + c2*M + c3  with c_j = carry_j + (x_j << 8): evaluated
// as (hi*M^2 + (c0*M + c1))*M^2 + (c2*M + c3), i.e. only TWO multiply-adds depend on the previous hi instead of four
// — an in-order lone warp otherwise stalls ~20 cycles per word on that tail (tools/micro/hash_phase_profile.cu).
__device__ __forceinline__ void fold4(Fnv& h, uint32_t word) {
  constexpr uint32_t M = 0x1b3u, M2 = M * M;
  uint32_t lo = h.lo, c[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t x = lo ^ ((word >> (8 * j)) & 0xffu);
    uint32_t carry;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, 435;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(carry) : "r"(x));
    c[j] = carry + (x << 8);
  }
  h.hi = (h.hi * M2 + (c[0] * M + c[1])) * M2 + (c[2] * M + c[3]);
  h.lo = lo;
}
__device__ __forceinline__ Fnv fnv_init();
__device__ __forceinline__ void fold_head64(Fnv& h, uint32_t major, uint64_t n);
__device__ __forceinline__ void fold(Fnv& h, uint32_t byte);

// Block prefix  0x83 | U(parent) | array-head(bs)  for the common shape (parent >= 2^32 so U(parent) is the 9-byte
// form, bs < 24 so the array head is one byte): 11 bytes packed into words and folded with fold4.
__device__ __forceinline__ void fold_prefix(Fnv& h, uint64_t parent, uint32_t bs) {
  if (parent >= 0x100000000ull && bs < 24u) {
    const uint32_t ph = (uint32_t)(parent >> 32), pl = (uint32_t)parent;
    // stream: 83 1b P7 P6 | P5 P4 P3 P2 | P1 P0 (80+bs)          (P7 = most significant byte of parent)
    fold4(h, 0x83u | (0x1bu << 8) | ((ph >> 24) << 16) | (((ph >> 16) & 0xffu) << 24));
    fold4(h, ((ph >> 8) & 0xffu) | ((ph & 0xffu) << 8) | ((pl >> 24) << 16) | (((pl >> 16) & 0xffu) << 24));
    fold(h, (pl >> 8) & 0xffu);
    fold(h, pl & 0xffu);
    fold(h, 0x80u | bs);
  } else {
    fold(h, 0x83u);
    fold_head64(h, 0x00u, parent);
    fold_head64(h, 0x80u, (uint64_t)bs);
  }
}
__device__ __forceinline__ Fnv fnv_init() { return Fnv{(uint32_t)kFnvOffset, (uint32_t)(kFnvOffset >> 32)}; }
__device__ __forceinline__ uint64_t fnv_value(const Fnv& h) { return ((uint64_t)h.hi << 32) | h.lo; }

// CBOR head for major type `major` (already shifted <<5) with argument n, shortest form.
__device__ __forceinline__ void fold_head64(Fnv& h, uint32_t major, uint64_t n) {
  if (n < 24) {
    fold(h, major | (uint32_t)n);
  } else if (n < 0x100ull) {
    fold(h, major | 24);
    fold(h, (uint32_t)n);
  } else if (n < 0x10000ull) {
    fold(h, major | 25);
    fold(h, (uint32_t)(n >> 8));
    fold(h, (uint32_t)n & 0xff);
  } else if (n < 0x100000000ull) {
    fold(h, major | 26);
#pragma unroll
    for (int s = 24; s >= 0; s -= 8) fold(h, (uint32_t)(n >> s) & 0xff);
  } else {
    fold(h, major | 27);
#pragma unroll
    for (int s = 56; s >= 0; s -= 8) fold(h, (uint32_t)(n >> s) & 0xff);
  }
}

// Token bytes are produced in two passes per block so that a lone warp (one chain per lane, nothing else on its
// scheduler to hide latency) issues as few instructions as possible:
//   pass 1 (independent across tokens, full ILP): every token's CBOR unsigned-int encoding (1, 2, 3 or 5 bytes) is
//           written to the lane's private strip of shared memory with five unconditional byte stores — bytes past
//           the encoding's length are overwritten by the next token, so there is no branch on the token width;
//   pass 2 (the serial chain): the strip is read back a word at a time and folded byte by byte.
#ifdef KVB_HASH_PROFILE
__device__ long long g_hash_prof[16];
#endif
constexpr int kStageTokens = 16;  // tokens staged per pass
constexpr int kStageWords = 28;   // 112 B strip (16 x 5 B + slack), 16 B aligned so it is read back with 128-bit loads

__device__ __forceinline__ int stage_token(uint8_t* buf, int n, uint32_t t) {
  const bool ge24 = t >= 24u, ge256 = t >= 0x100u, ge64k = t >= 0x10000u;
  const uint32_t head = ge64k ? 0x1au : (ge256 ? 0x19u : (ge24 ? 0x18u : t));
  // payload, left-aligned big-endian: 4 bytes (>= 65536), 2 bytes (>= 256) or 1 byte (>= 24)
  const uint32_t pay = ge64k ? t : (ge256 ? (t << 16) : (t << 24));
  buf[n] = (uint8_t)head;
  buf[n + 1] = (uint8_t)(pay >> 24);
  buf[n + 2] = (uint8_t)(pay >> 16);
  buf[n + 3] = (uint8_t)(pay >> 8);
  buf[n + 4] = (uint8_t)pay;
  return n + (ge64k ? 5 : (ge256 ? 3 : (ge24 ? 2 : 1)));
}

template <bool BRANCH_FREE>
__device__ __forceinline__ void fold_staged(Fnv& h, const uint8_t* buf, int n) {
  // Pull the whole strip into registers first (five 128-bit LDS, one shared-memory latency in total), then fold
  // from registers: folding out of shared memory word by word exposes the LDS latency on every word to the in-order
  // issue of a lone warp (29 vs 14 cycles/byte, tools/micro/hash_micro.cu).  asm volatile pins the loads up front.
  constexpr int kWords = (kStageTokens * 5 + 3) / 4;  // 20
  uint32_t v[kWords];
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(buf);
#pragma unroll
  for (int q = 0; q < kWords / 4; ++q)
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[4 * q]), "=r"(v[4 * q + 1]), "=r"(v[4 * q + 2]), "=r"(v[4 * q + 3])
                 : "r"(saddr + 16 * q));  // words past n hold stale bytes and are never folded
  // Whole words, branch-free: every lane runs the same straight-line code (a branch per word costs a lone warp
  // ~4 cycles/byte, tools/micro/hash_micro.cu) and keeps the folded state only for the words it really has.
  // The first 12 words (48 B = 16 three-byte tokens, the minimum for real vocabularies) are usually all present.
  // Large batches have many warps per scheduler: there the issue slots, not the latency, are scarce, and skipping
  // absent words (a branch per word) wins — BRANCH_FREE = false.
  const int nwords = n >> 2;
#pragma unroll
  for (int i = 0; i < kWords; ++i) {
    if (BRANCH_FREE) {
      Fnv t = h;
      fold4(t, v[i]);
      const bool take = i < nwords;
      h.lo = take ? t.lo : h.lo;
      h.hi = take ? t.hi : h.hi;
    } else {
      if (i >= nwords) break;
      fold4(h, v[i]);
    }
  }
  const int tail = n & 3;  // 0..3 trailing bytes live in word n/4 (dynamic index: re-read it from the strip)
  if (tail) {
    const uint32_t tw = reinterpret_cast<const uint32_t*>(buf)[n >> 2];
    fold(h, tw & 0xffu);
    if (tail > 1) fold(h, (tw >> 8) & 0xffu);
    if (tail > 2) fold(h, (tw >> 16) & 0xffu);
  }
}

template <int BS>
__global__ void __launch_bounds__(128) hash_chain_kernel(const uint32_t* __restrict__ tokens,
                                                         const int64_t* __restrict__ prompt_off,
                                                         const uint64_t* __restrict__ parents, int32_t n_prompts,
                                                         int32_t block_size_rt, const uint8_t* __restrict__ extra,
                                                         const int64_t* __restrict__ extra_off,
                                                         uint64_t* __restrict__ out_keys,
                                                         const int64_t* __restrict__ key_off) {
  __shared__ __align__(16) uint32_t strips[128 * kStageWords];
  uint8_t* buf = reinterpret_cast<uint8_t*>(strips + threadIdx.x * kStageWords);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_prompts) return;
  const int bs = BS > 0 ? BS : block_size_rt;
  const int64_t t0 = prompt_off[p];
  const int64_t nblk = (prompt_off[p + 1] - t0) / bs;  // tail tokens dropped (token_processor.go:166-168)
  const int64_t k0 = key_off[p];
  const uint32_t* tk = tokens + t0;
  uint64_t parent = parents[p];

  uint32_t cur[BS > 0 ? BS : 1];
  if (BS > 0 && nblk > 0) {
#pragma unroll
    for (int j = 0; j < BS; ++j) cur[j] = __ldg(tk + j);
  }
  for (int64_t i = 0; i < nblk; ++i) {
    uint32_t nxt[BS > 0 ? BS : 1];
    if (BS > 0 && i + 1 < nblk) {  // prefetch the next block's tokens under this block's fold chain
#pragma unroll
      for (int j = 0; j < BS; ++j) nxt[j] = __ldg(tk + (i + 1) * BS + j);
    }
    Fnv h = fnv_init();
    fold_prefix(h, parent, (uint32_t)bs);           // array(3) | parent | array(bs)
    if (BS > 0 && BS <= kStageTokens) {
      int n = 0;
#pragma unroll
      for (int j = 0; j < BS; ++j) n = stage_token(buf, n, cur[j]);
      fold_staged<true>(h, buf, n);
    } else {
      for (int j0 = 0; j0 < bs; j0 += kStageTokens) {
        const int m = min(kStageTokens, bs - j0);
        int n = 0;
        for (int j = 0; j < m; ++j) n = stage_token(buf, n, __ldg(tk + i * bs + j0 + j));
        fold_staged<true>(h, buf, n);
      }
    }
    bool text = true;
    if (extra_off != nullptr) {                     // pre-encoded X(extra_i), host-built (extra_keys.go)
      const int64_t e0 = extra_off[k0 + i], e1 = extra_off[k0 + i + 1];
      if (e1 > e0) {
        text = false;
        for (int64_t e = e0; e < e1; ++e) fold(h, extra[e]);
      }
    }
    if (text) fold(h, 0xf6u);                       // nil extra -> CBOR null
    parent = fnv_value(h);
    out_keys[k0 + i] = parent;
    if (BS > 0) {
#pragma unroll
      for (int j = 0; j < BS; ++j) cur[j] = nxt[j];
    }
  }
}

// Two-warp variant for block sizes <= 16 (the vLLM default is 16): 32 chains per CTA, warp 0 STAGES block i+1 of
// every chain (pass 1) while warp 1 FOLDS block i (pass 2), double-buffered strips, one CTA barrier per block.
// The two warps sit on different SM sub-partitions, so the encode work leaves the serial chain's issue stream:
// measured on B200 the one-warp kernel is issue-bound (1048 instructions per block at 3.1 cycles per instruction
// for a lone warp, profiles/r01_ncu_hash_*.txt).
template <int BS, bool BRANCH_FREE>
__global__ void __launch_bounds__(64) hash_chain_kernel_2w(const uint32_t* __restrict__ tokens,
                                                           const int64_t* __restrict__ prompt_off,
                                                           const uint64_t* __restrict__ parents, int32_t n_prompts,
                                                           const uint8_t* __restrict__ extra,
                                                           const int64_t* __restrict__ extra_off,
                                                           uint64_t* __restrict__ out_keys,
                                                           const int64_t* __restrict__ key_off) {
  static_assert(BS > 0 && BS <= kStageTokens, "two-warp kernel stages one whole block per strip");
  __shared__ __align__(16) uint32_t strips[2][32 * kStageWords];
  __shared__ int nbytes[2][32];
  const int lane = threadIdx.x & 31;
  const bool stager = threadIdx.x < 32;
  const int p = blockIdx.x * 32 + lane;
  const bool live = p < n_prompts;
  int64_t t0 = 0, nblk = 0, k0 = 0;
  if (live) {
    t0 = prompt_off[p];
    nblk = (prompt_off[p + 1] - t0) / BS;
    k0 = key_off[p];
  }
  // both warps must run the same number of barriers: longest chain in this CTA
  int64_t nmax = nblk;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
  const uint32_t* tk = tokens + t0;
  uint64_t parent = (!stager && live) ? parents[p] : 0;

  uint32_t cur[BS];  // stager: tokens of the next block to stage, loaded one barrier interval ahead
  auto load_block = [&](int64_t i) {
    if (i < nblk) {
#pragma unroll
      for (int j = 0; j < BS; ++j) cur[j] = __ldg(tk + i * BS + j);
    }
  };
  auto stage_block = [&](int64_t i) {  // stages `cur` as block i, then starts loading block i+1
    uint8_t* buf = reinterpret_cast<uint8_t*>(strips[i & 1] + lane * kStageWords);
    int n = 0;
#pragma unroll
    for (int j = 0; j < BS; ++j) n = stage_token(buf, n, cur[j]);
    nbytes[i & 1][lane] = n;
    load_block(i + 1);
  };
  if (stager) {
    load_block(0);
    if (nblk > 0) stage_block(0);
  }
  __syncthreads();
#ifdef KVB_HASH_PROFILE
  long long pc[5] = {0, 0, 0, 0, 0};  // stager work, folder prefix, folder staged, folder rest, barrier wait
#define PROF(slot, t_begin) pc[slot] += clock64() - (t_begin)
#else
#define PROF(slot, t_begin)
#endif
  for (int64_t i = 0; i < nmax; ++i) {
#ifdef KVB_HASH_PROFILE
    long long tb = clock64();
#endif
    if (stager) {
      if (i + 1 < nblk) stage_block(i + 1);
      PROF(0, tb);
    } else if (i < nblk) {
      Fnv h = fnv_init();
      fold_prefix(h, parent, (uint32_t)BS);
      PROF(1, tb);
#ifdef KVB_HASH_PROFILE
      long long ts = clock64();
#endif
      fold_staged<BRANCH_FREE>(h, reinterpret_cast<const uint8_t*>(strips[i & 1] + lane * kStageWords), nbytes[i & 1][lane]);
      PROF(2, ts);
#ifdef KVB_HASH_PROFILE
      long long tr = clock64();
#endif
      bool text = true;
      if (extra_off != nullptr) {
        const int64_t e0 = extra_off[k0 + i], e1 = extra_off[k0 + i + 1];
        if (e1 > e0) {
          text = false;
          for (int64_t e = e0; e < e1; ++e) fold(h, extra[e]);
        }
      }
      if (text) fold(h, 0xf6u);
      parent = fnv_value(h);
      out_keys[k0 + i] = parent;
      PROF(3, tr);
    }
#ifdef KVB_HASH_PROFILE
    long long tw = clock64();
#endif
    __syncthreads();
    PROF(4, tw);
  }
#ifdef KVB_HASH_PROFILE
  if (blockIdx.x == 0 && lane == 0) {
    long long* dst = g_hash_prof + (stager ? 0 : 5);
    for (int q = 0; q < 5; ++q) dst[q] = pc[q];
    g_hash_prof[10] = nmax;
  }
#endif
}

// getInitHash: H(seed_hash, nil, model_name) = FNV64a(83 | U(seed) | f6 | text(model))
__global__ void init_hash_kernel(uint64_t seed_hash, const uint8_t* __restrict__ name, uint32_t len,
                                 uint64_t* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Fnv h = fnv_init();
  fold(h, 0x83u);
  fold_head64(h, 0x00u, seed_hash);
  fold(h, 0xf6u);
  fold_head64(h, 0x60u, len);
  for (uint32_t i = 0; i < len; ++i) fold(h, name[i]);
  *out = fnv_value(h);
}

int launch_hash_blocks(const uint32_t* tokens, const int64_t* prompt_off, const uint64_t* parents, int32_t n_prompts,
                       int32_t block_size, const uint8_t* extra, const int64_t* extra_off, uint64_t* out_keys,
                       const int64_t* key_off, cudaStream_t s) {
  if (n_prompts <= 0) return KVB_OK;
  // small batches: 32-thread CTAs so the chains spread over the SMs; large: 128
  const int threads = n_prompts >= 148 * 128 ? 128 : 32;
  const int grid = (n_prompts + threads - 1) / threads;
  static const bool one_warp = std::getenv("KVB_HASH_ONE_WARP") != nullptr;  // A/B switch for profiling
  const int grid2 = (n_prompts + 31) / 32;
  // fewer CTAs than ~2 per SM sub-partition: every warp is alone on its scheduler -> optimise latency, else issue slots
  int dev = 0;
  cudaGetDevice(&dev);
  const bool latency_bound = grid2 <= 4 * sm_count(dev);
  switch (one_warp ? -1 : block_size) {
#define KVB_LAUNCH_2W(BSV)                                                                                          \
  if (latency_bound)                                                                                                \
    hash_chain_kernel_2w<BSV, true><<<grid2, 64, 0, s>>>(tokens, prompt_off, parents, n_prompts, extra, extra_off,  \
                                                         out_keys, key_off);                                        \
  else                                                                                                              \
    hash_chain_kernel_2w<BSV, false><<<grid2, 64, 0, s>>>(tokens, prompt_off, parents, n_prompts, extra, extra_off, \
                                                          out_keys, key_off)
    case 16:
      KVB_LAUNCH_2W(16);
      break;
    case 8:
      KVB_LAUNCH_2W(8);
      break;
    case 4:
      KVB_LAUNCH_2W(4);
      break;
#undef KVB_LAUNCH_2W
    case -1:
      if (block_size == 16) {
        hash_chain_kernel<16><<<grid, threads, 0, s>>>(tokens, prompt_off, parents, n_prompts, block_size, extra,
                                                       extra_off, out_keys, key_off);
        break;
      }
      [[fallthrough]];
    default:
      hash_chain_kernel<0><<<grid, threads, 0, s>>>(tokens, prompt_off, parents, n_prompts, block_size, extra,
                                                    extra_off, out_keys, key_off);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("hash kernel launch failed: %s", cudaGetErrorString(e));
    return KVB_ERR_CUDA;
  }
  count_launch();
  return KVB_OK;
}

}  // namespace kvb

using namespace kvb;

extern "C" {

uint64_t kvb_fnv64a(const void* data, size_t len) {
  // seed hash only (token_processor.go:90-95): once per processor, a handful of bytes
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint64_t h = kFnvOffset;
  for (size_t i = 0; i < len; ++i) h = (h ^ p[i]) * kFnvPrime;
  return h;
}

int kvb_init_hash(int device, uint64_t seed_hash, const char* model_name, size_t model_len, uint64_t* out) {
  KVB_REQUIRE(out != nullptr, "out is NULL");
  KVB_REQUIRE(model_name != nullptr || model_len == 0, "model_name is NULL");
  KVB_REQUIRE(model_len < (1u << 31), "model name too long");
  DeviceGuard g(device);
  if (!g.ok) {
    set_error("cannot select CUDA device %d", device);
    return KVB_ERR_CUDA;
  }
  uint8_t* d = nullptr;
  KVB_CUDA_TRY(cudaMalloc(&d, model_len + 8 + 8));
  uint64_t* d_out = reinterpret_cast<uint64_t*>(d);
  uint8_t* d_name = d + 8;
  cudaError_t e = cudaSuccess;
  if (model_len) e = cudaMemcpy(d_name, model_name, model_len, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    init_hash_kernel<<<1, 32>>>(seed_hash, d_name, (uint32_t)model_len, d_out);
    count_launch();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  KVB_CUDA_TRY(e);
  return KVB_OK;
}

int kvb_hash_token_blocks_dev(int device, const uint32_t* tokens, const int64_t* prompt_off, const uint64_t* parents,
                              int32_t n_prompts, int32_t block_size, const uint8_t* extra, const int64_t* extra_off,
                              uint64_t* out_keys, const int64_t* key_off, void* stream) {
  KVB_REQUIRE(block_size > 0, "blockSize must be greater than 0, got %d", block_size);  // token_processor.go:86-88
  KVB_REQUIRE(n_prompts >= 0, "negative prompt count");
  if (n_prompts == 0) return KVB_OK;
  KVB_REQUIRE(tokens && prompt_off && parents && out_keys && key_off, "NULL argument");
  DeviceGuard g(device);
  return launch_hash_blocks(tokens, prompt_off, parents, n_prompts, block_size, extra, extra_off, out_keys, key_off,
                            static_cast<cudaStream_t>(stream));
}

int kvb_hash_token_blocks(int device, const uint32_t* tokens, const int64_t* prompt_off, const uint64_t* parents,
                          int32_t n_prompts, int32_t block_size, const uint8_t* extra, const int64_t* extra_off,
                          uint64_t* out_keys, int64_t* out_key_off, void* stream) {
  KVB_REQUIRE(block_size > 0, "blockSize must be greater than 0, got %d", block_size);
  KVB_REQUIRE(n_prompts >= 0, "negative prompt count");
  KVB_REQUIRE(out_key_off != nullptr, "out_key_off is NULL");
  out_key_off[0] = 0;
  if (n_prompts == 0) return KVB_OK;
  KVB_REQUIRE(prompt_off && parents, "NULL argument");
  for (int32_t p = 0; p < n_prompts; ++p) {
    KVB_REQUIRE(prompt_off[p + 1] >= prompt_off[p], "prompt_off not monotonic at %d", p);
    out_key_off[p + 1] = out_key_off[p] + (prompt_off[p + 1] - prompt_off[p]) / block_size;
  }
  const int64_t total_keys = out_key_off[n_prompts];
  const int64_t total_tok = prompt_off[n_prompts] - prompt_off[0];
  if (total_keys == 0) return KVB_OK;
  KVB_REQUIRE(tokens && out_keys, "NULL argument");
  DeviceGuard g(device);
  if (!g.ok) {
    set_error("cannot select CUDA device %d", device);
    return KVB_ERR_CUDA;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t extra_bytes = extra_off ? extra_off[total_keys] : 0;
  // one device scratch: [tokens | prompt_off | key_off | parents | keys | extra_off | extra]
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t o_tok = 0;
  const size_t o_poff = o_tok + al(total_tok * 4);
  const size_t o_koff = o_poff + al((n_prompts + 1) * 8);
  const size_t o_par = o_koff + al((n_prompts + 1) * 8);
  const size_t o_keys = o_par + al(n_prompts * 8);
  const size_t o_eoff = o_keys + al(total_keys * 8);
  const size_t o_ext = o_eoff + (extra_off ? al((total_keys + 1) * 8) : 0);
  const size_t total = o_ext + (extra_off ? al(extra_bytes) : 0) + 256;
  uint8_t* d = nullptr;
  KVB_CUDA_TRY(cudaMallocAsync(&d, total, s));
  std::vector<int64_t> poff_rel(n_prompts + 1);
  for (int32_t p = 0; p <= n_prompts; ++p) poff_rel[p] = prompt_off[p] - prompt_off[0];
  cudaError_t e = cudaMemcpyAsync(d + o_tok, tokens + prompt_off[0], total_tok * 4, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d + o_poff, poff_rel.data(), (n_prompts + 1) * 8, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + o_koff, out_key_off, (n_prompts + 1) * 8, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + o_par, parents, n_prompts * 8, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && extra_off) {
    e = cudaMemcpyAsync(d + o_eoff, extra_off, (total_keys + 1) * 8, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && extra_bytes)
      e = cudaMemcpyAsync(d + o_ext, extra, extra_bytes, cudaMemcpyHostToDevice, s);
  }
  int rc = KVB_OK;
  if (e == cudaSuccess) {
    rc = launch_hash_blocks(reinterpret_cast<uint32_t*>(d + o_tok), reinterpret_cast<int64_t*>(d + o_poff),
                            reinterpret_cast<uint64_t*>(d + o_par), n_prompts, block_size,
                            extra_off ? d + o_ext : nullptr,
                            extra_off ? reinterpret_cast<int64_t*>(d + o_eoff) : nullptr,
                            reinterpret_cast<uint64_t*>(d + o_keys), reinterpret_cast<int64_t*>(d + o_koff), s);
    if (rc == KVB_OK) e = cudaMemcpyAsync(out_keys, d + o_keys, total_keys * 8, cudaMemcpyDeviceToHost, s);
  }
  cudaError_t e2 = cudaStreamSynchronize(s);  // poff_rel and the caller's buffers must outlive the copies
  cudaFreeAsync(d, s);
  if (rc != KVB_OK) return rc;
  KVB_CUDA_TRY(e);
  KVB_CUDA_TRY(e2);
  return KVB_OK;
}

}  // extern "C"
