// C++ round-trip through kvb::StorageOffloadEngine (include/kvb_engine.hpp), the mirror of the reference's compiled
// engine class: store 8 blocks of 3 tensors in two files, zero the device memory, load them back, compare bytes;
// both tiers.  Device memory comes from kvb-independent cudaMalloc (the engine only borrows pointers).
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>

#include "kvb_engine.hpp"

#define CHECK(cond)                                                                        \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      std::fprintf(stderr, "CHECK failed at %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, kvb_last_error()); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

static bool drain(kvb::StorageOffloadEngine& e, int job, bool* ok) {
  for (int spin = 0; spin < 20000; ++spin) {
    for (auto& jr : e.get_finished())
      if (jr.first == job) {
        *ok = jr.second;
        return true;
      }
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  return false;
}

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "/tmp/kvb-cpp-engine";
  const int T = 3, N = 16;
  const int64_t page = 8192;
  std::vector<std::vector<uint8_t>> host(T, std::vector<uint8_t>(N * page));
  std::vector<void*> dev(T);
  std::vector<kvb::KVTensor> tensors;
  for (int t = 0; t < T; ++t) {
    for (size_t i = 0; i < host[t].size(); ++i) host[t][i] = (uint8_t)((i * 2654435761u + t * 97) >> 13);
    CHECK(cudaMalloc(&dev[t], N * page) == cudaSuccess);
    CHECK(cudaMemcpy(dev[t], host[t].data(), N * page, cudaMemcpyHostToDevice) == cudaSuccess);
    tensors.push_back({dev[t], N, page});
  }
  const std::vector<std::vector<int64_t>> groups{{11, 3}, {0, 15, 7, 8}};  // first file partial, like the handlers produce
  for (int tier : {KVB_TIER_FILE, KVB_TIER_HOST_ARENA}) {
    kvb::StorageOffloadEngine eng(2, 4, tensors, 1, "disabled", 0.0f, 0, tier, 1 << 22);
    const std::vector<std::string> files{dir + "/t" + std::to_string(tier) + "/a.bin", dir + "/t" + std::to_string(tier) + "/b.bin"};
    bool ok = false;
    CHECK(eng.async_store_gpu_blocks(1, files, groups));
    CHECK(drain(eng, 1, &ok) && ok);
    for (int t = 0; t < T; ++t) CHECK(cudaMemset(dev[t], 0, N * page) == cudaSuccess);
    CHECK(eng.async_load_gpu_blocks(2, files, groups));
    CHECK(drain(eng, 2, &ok) && ok);
    {  // manager lookup in one call: files named as FileMapper names them, <base>/<hhh>/<hh>/<016x>.bin
      const std::string base = dir + "/lk" + std::to_string(tier);
      auto name = [&](uint64_t h) {
        char hex[17];
        std::snprintf(hex, sizeof(hex), "%016llx", (unsigned long long)h);
        const std::string x(hex);
        return base + "/" + x.substr(0, 3) + "/" + x.substr(3, 2) + "/" + x + ".bin";
      };
      const std::vector<uint64_t> hs{0x1111222233334444ull, 0xaaaabbbbccccddddull, 0x0123456789abcdefull, 0xfeedfacecafebeefull};
      CHECK(eng.async_store_gpu_blocks(7, {name(hs[0]), name(hs[1]), name(hs[3])}, {{1}, {2}, {4}}));
      CHECK(drain(eng, 7, &ok) && ok);
      CHECK(eng.lookup(base, hs) == 2);                      // the hole at hs[2] ends the prefix
      CHECK(eng.lookup(base, {hs[3], hs[0]}) == 2 && eng.lookup(base, {hs[2]}) == 0 && eng.lookup(base, {}) == 0);
    }
    eng.wait_job(12345);  // unknown job: returns
    CHECK(!eng.async_store_gpu_blocks(3, files, {{0}}));          // files / id lists differ in length -> false
    CHECK(!eng.async_store_gpu_blocks(4, {"x"}, {{99}}));        // block id out of range -> false, nothing thrown
    std::vector<uint8_t> back(N * page);
    for (int t = 0; t < T; ++t) {
      CHECK(cudaMemcpy(back.data(), dev[t], N * page, cudaMemcpyDeviceToHost) == cudaSuccess);
      for (int b = 0; b < N; ++b) {
        bool listed = false;
        for (auto& g : groups)
          for (auto id : g) listed |= id == b;
        const uint8_t* got = back.data() + b * page;
        if (listed) {
          CHECK(std::memcmp(got, host[t].data() + b * page, page) == 0);  // restored bit-exact
        } else {
          for (int64_t i = 0; i < page; ++i) CHECK(got[i] == 0);          // untouched (still zero)
        }
      }
      CHECK(cudaMemcpy(dev[t], host[t].data(), N * page, cudaMemcpyHostToDevice) == cudaSuccess);
    }
  }
  std::printf("OK engine round trips\n");
  return 0;
}
