/*
 * dup_emulation.cpp -- runs the per-thread bodies of fastp_b200/csrc/fp_dup.h ON THE HOST, one "thread" at a time in a shuffled
 * order per pass (a stand-in for the device's arbitrary scheduling), over the batches of a file written by
 * tests/test_duplicate_oracle.py, and prints one 0/1 flag per unit.  The device kernels (fp_dup.cuh) wrap the same bodies.
 * file: int64 nbatches; per batch: int64 n, int32 stride, int32 paired, seq1[n*stride], len1[n] (u16), [seq2, len2]
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>
#include <random>
#include <vector>
#include "fp_dup.h"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: dup_emulation <file> <accuracy> [seed]\n"); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    const int accuracy = atoi(argv[2]);
    std::mt19937_64 rng(argc > 3 ? atoll(argv[3]) : 1);
    uint64_t buf_bytes; int buf_num;
    fp_dup_sizes(accuracy, &buf_bytes, &buf_num);
    std::vector<uint64_t> primes((size_t)buf_num * FP_DUP_PRIME_LEN);
    fp_dup_primes(primes.data(), (int)primes.size());
    fp_dup_state S;
    S.buf_num = buf_num; S.buf_bits = buf_bytes << 3; S.offset_mask = (uint64_t)FP_DUP_PRIME_LEN * buf_num - 1;
    S.bits = (uint32_t*)calloc((size_t)buf_num * (buf_bytes / 4), 4);
    S.primes = primes.data();
    int64_t nb = 0; f.read((char*)&nb, 8);
    for (int64_t bi = 0; bi < nb; bi++) {
        int64_t n; int32_t stride, paired;
        f.read((char*)&n, 8); f.read((char*)&stride, 4); f.read((char*)&paired, 4);
        std::vector<uint8_t> s1((size_t)n * stride), s2; std::vector<uint16_t> l1(n), l2;
        f.read((char*)s1.data(), s1.size()); f.read((char*)l1.data(), n * 2);
        if (paired) { s2.resize((size_t)n * stride); l2.resize(n); f.read((char*)s2.data(), s2.size()); f.read((char*)l2.data(), n * 2); }
        if (!f) { fprintf(stderr, "short file\n"); return 1; }
        std::vector<uint64_t> pos((size_t)n * buf_num);
        uint64_t cap = 1; while (cap < (uint64_t)(2 * n * buf_num + 16)) cap <<= 1;
        std::vector<uint64_t> keys(cap, FP_DUP_EMPTY); std::vector<uint32_t> vals(cap, 0xFFFFFFFFu);
        S.pos = pos.data(); S.keys = keys.data(); S.vals = vals.data(); S.table_mask = cap - 1;
        std::vector<long long> order(n), order2((size_t)n * buf_num);
        std::iota(order.begin(), order.end(), 0); std::iota(order2.begin(), order2.end(), 0);
        std::shuffle(order.begin(), order.end(), rng);
        for (long long u : order) fp_dup_hash_unit(S, u, s1.data(), l1.data(), paired ? s2.data() : nullptr, paired ? l2.data() : nullptr, stride, paired);
        std::shuffle(order2.begin(), order2.end(), rng);
        for (long long t : order2) fp_dup_first(S, t);
        std::vector<uint8_t> flag(n);
        std::shuffle(order.begin(), order.end(), rng);
        for (long long u : order) flag[u] = (uint8_t)fp_dup_decide(S, u);
        std::shuffle(order2.begin(), order2.end(), rng);
        for (long long t : order2) fp_dup_commit(S, t);
        for (int64_t i = 0; i < n; i++) putchar('0' + flag[i]);
        putchar('\n');
    }
    free(S.bits);
    return 0;
}
