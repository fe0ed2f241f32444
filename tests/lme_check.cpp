// host check of csrc/libm_exact.h against the C library's sinf / cosf (tests/test_libm_exact.py)
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include "libm_exact.h"
int main(int argc, char** argv) {
    const unsigned stride = argc > 1 ? (unsigned)atoi(argv[1]) : 509u;
    bool fma = __builtin_cpu_supports("fma") && __builtin_cpu_supports("avx2");   // glibc's ifunc rule for s_sinf / s_cosf
    if (argc > 2) fma = atoi(argv[2]) != 0;   // forced variant (run with GLIBC_TUNABLES=glibc.cpu.hwcaps=-AVX2,-FMA to make the C library pick its SSE2 build)
    float top = 8.0f; uint32_t topBits; memcpy(&topBits, &top, 4);
    unsigned long long n = 0, bad = 0;
    for (uint32_t b = 0; b <= topBits; b += stride) {
        float x; memcpy(&x, &b, 4);
        for (int sgn = 0; sgn < 2; sgn++) {
            const float v = sgn ? -x : x;
            volatile float in = v;   // keep the compiler from folding the libm calls
            const float s = sinf(in), c = cosf(in);
            const float s2 = libm_exact::sinf_glibc(fma, v), c2 = libm_exact::cosf_glibc(fma, v);
            n += 2;
            if (memcmp(&s, &s2, 4) != 0) { if (bad < 5) printf("sin mismatch x=%a libm=%a mine=%a\n", v, s, s2); bad++; }
            if (memcmp(&c, &c2, 4) != 0) { if (bad < 5) printf("cos mismatch x=%a libm=%a mine=%a\n", v, c, c2); bad++; }
        }
    }
    printf("fma_variant %d checked %llu mismatches %llu\n", (int)fma, n, bad);
    return bad ? 1 : 0;
}
