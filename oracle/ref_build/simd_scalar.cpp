// Scalar stand-in for the reference's src/simd.cpp (Google Highway 1.3.0 is not vendored / installed
// here).  Each function follows the reference's OWN scalar restatement of its SIMD kernel
// (src/simd.cpp:281-324), which the reference's testSimd (src/simd.cpp:326-564) pins the SIMD code to.
// countMismatchesBounded has no scalar twin in the reference; its contract (src/simd.h:28-31) is "the
// mismatch count if <= limit, or a value > limit": the full count satisfies it.
#include "simd.h"
#include <cstdint>
#include <cstring>
#include <string>
namespace fastp_simd {
void countQualityMetrics(const char* qualstr, const char* seqstr, int len, char qualThreshold,
                         int& lowQualNum, int& nBaseNum, int& totalQual) {
    lowQualNum = 0; nBaseNum = 0; totalQual = 0;
    for (int i = 0; i < len; i++) {
        uint8_t q = static_cast<uint8_t>(qualstr[i]);
        totalQual += q - 33;
        if (q < static_cast<uint8_t>(qualThreshold)) lowQualNum++;
        if (seqstr[i] == 'N') nBaseNum++;
    }
}
void reverseComplement(const char* src, char* dst, int len) {
    for (int i = 0; i < len; i++) {
        char c;
        switch (src[i]) {
            case 'A': case 'a': c = 'T'; break;
            case 'T': case 't': c = 'A'; break;
            case 'C': case 'c': c = 'G'; break;
            case 'G': case 'g': c = 'C'; break;
            default: c = 'N'; break;
        }
        dst[len - 1 - i] = c;
    }
}
int countAdjacentDiffs(const char* data, int len) {
    int diff = 0;
    for (int i = 0; i < len - 1; i++) if (data[i] != data[i + 1]) diff++;
    return diff;
}
int countMismatches(const char* a, const char* b, int len) {
    int diff = 0;
    for (int i = 0; i < len; i++) if (a[i] != b[i]) diff++;
    return diff;
}
int countMismatchesBounded(const char* a, const char* b, int len, int limit) {
    (void)limit;
    return countMismatches(a, b, len);
}
// Known-answer checks in the spirit of src/simd.cpp:326-564 (empty / len-1 / long / mixed inputs).
bool testSimd() {
    bool pass = true;
    { char d[8]; reverseComplement("ACGTN", d, 5); pass &= (memcmp(d, "NACGT", 5) == 0); }
    { char d[1]; reverseComplement("", d, 0); }
    { std::string s(68, 'A'); std::string t(68, 'A'); t[3] = 'C'; t[67] = 'G';
      pass &= countMismatches(s.data(), t.data(), 68) == 2;
      pass &= countMismatchesBounded(s.data(), t.data(), 68, 5) == 2;
      pass &= countAdjacentDiffs(t.data(), 68) == 3; }
    { int lq, nb, tq; countQualityMetrics("II#5", "ACNN", 4, '0', lq, nb, tq);
      pass &= (lq == 1 && nb == 2 && tq == 40 + 40 + 2 + 20); }
    pass &= countAdjacentDiffs("A", 1) == 0 && countAdjacentDiffs("", 0) == 0;
    return pass;
}
}
