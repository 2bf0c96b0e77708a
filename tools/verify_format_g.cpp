// verify_format_g.cpp -- compares the frame writer's number formatting (clsph_host_format_g, exported by
// libclsph_host.so) with snprintf("%g") over float bit patterns 0, stride, 2*stride, ... (stride 1 = all
// 2^32, about four minutes on 8 cores).
//   g++ -O2 -std=c++17 tools/verify_format_g.cpp -o /tmp/verify_format_g -pthread -Llibclsph_b200 -lclsph_host \
//       -Wl,-rpath,$PWD/libclsph_b200 && /tmp/verify_format_g 1
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

extern "C" int clsph_host_format_g(float v, char* out);

int main(int argc, char** argv) {
  const uint64_t stride = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 997;
  const unsigned threads = std::max(1u, std::thread::hardware_concurrency());
  std::atomic<uint64_t> bad{0}, seen{0};
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      char a[64], b[64];
      uint64_t n = 0;
      for (uint64_t u = t * stride; u < (1ull << 32); u += threads * stride) {
        const uint32_t bits = static_cast<uint32_t>(u);
        float v;
        std::memcpy(&v, &bits, 4);
        if (((bits >> 23) & 0xffu) == 0xffu) continue;  // inf / nan take the C library's path by construction
        clsph_host_format_g(v, a);
        std::snprintf(b, sizeof(b), "%g", static_cast<double>(v));
        ++n;
        if (std::strcmp(a, b) && bad++ < 10) std::fprintf(stderr, "MISMATCH bits %08x: writer '%s' printf '%s'\n", bits, a, b);
      }
      seen += n;
    });
  for (std::thread& th : pool) th.join();
  std::printf("stride %llu: %llu values compared, %llu mismatches\n", (unsigned long long)stride, (unsigned long long)seen.load(),
              (unsigned long long)bad.load());
  return bad ? 1 : 0;
}
