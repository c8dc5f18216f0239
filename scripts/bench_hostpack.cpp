/* Micro-benchmark of fp_pack_bases_row (host 2-bit packer): rows/s and GB/s of bases for a given thread count.
   g++ -O2 -std=c++17 -I include -I fastp_b200/csrc scripts/bench_hostpack.cpp fastp_b200/csrc/fp_hostpack.o -o /tmp/bench_hostpack -lpthread */
#include "fp_hostpack.h"
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
#include <cstdlib>
int main(int argc, char** argv) {
    const int NT = argc > 1 ? atoi(argv[1]) : 8;
    const long n = 2000000; const int HP = 150, PB = 40;
    std::vector<uint8_t> s((size_t)n * HP), d((size_t)n * PB);
    for (size_t i = 0; i < s.size(); i++) s[i] = "ACGT"[(i * 2654435761u >> 13) & 3];
    for (long r = 0; r < n; r += 37) s[(size_t)r * HP + (r % HP)] = 'N';
    for (int rep = 0; rep < 3; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < NT; t++) th.emplace_back([&, t]() {
            std::vector<fp_npos> nl;
            for (long r = n * t / NT; r < n * (t + 1) / NT; r++) fp_pack_bases_row(&s[(size_t)r * HP], HP, &d[(size_t)r * PB], (uint32_t)r, 0, nl);
        });
        for (auto& x : th) x.join();
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("threads %d: %.1f M rows/s  %.2f GB/s\n", NT, n / dt / 1e6, n * (double)HP / dt / 1e9);
    }
    unsigned x = 0; for (auto v : d) x += v; printf("%u\n", x);
}
