// CPU test of the copy pool behind the pinned staging ring (csrc/host_stage.h): many rounds of
// segment lists of ragged sizes copied by 1..6 threads must equal a plain memcpy.  No GPU, no CUDA
// calls (the ring and the uploader are exercised by the -m gpu tests with pageable inputs).
#include <cstdio>
#include <random>

#include "../../verifiable-fhe-paper_b200/csrc/host_stage.h"

int main() {
  std::mt19937_64 rng(1);
  for (unsigned threads = 1; threads <= 6; threads++) {
    hoststage::CopyPool pool(threads);
    if (pool.threads() != threads) return 2;
    for (int round = 0; round < 40; round++) {
      const int nseg = 1 + (int)(rng() % 7);
      std::vector<std::vector<unsigned char>> src(nseg);
      std::vector<hoststage::Segment> segs;
      size_t total = 0;
      for (auto& s : src) {
        s.resize(round == 0 ? 1 : (size_t)(rng() % (3u << 20)) + 1);
        for (auto& b : s) b = (unsigned char)rng();
        total += s.size();
      }
      std::vector<unsigned char> dst(total, 0), want(total);
      size_t off = 0;
      for (auto& s : src) {
        segs.push_back({reinterpret_cast<char*>(dst.data()) + off, reinterpret_cast<const char*>(s.data()), s.size()});
        memcpy(want.data() + off, s.data(), s.size());
        off += s.size();
      }
      pool.copy(segs);
      if (dst != want) {
        printf("mismatch threads=%u round=%d\n", threads, round);
        return 1;
      }
    }
  }
  printf("host stage copy pool ok\n");
  return 0;
}
