// Checks the oracle's restatement of the host libm's float functions (oracle/oracle_common.h, namespace glibcm)
// against the libm of the machine it runs on: sinf / cosf on every float with |x| < 120, atanf on all 2^32 floats,
// atan2f on 6e8 pairs from three distributions.  About two minutes on one core.
//   g++ -O2 -std=c++17 -ffp-contract=off -fno-builtin -mfma tools/scan_libm.cpp -o /tmp/scan_libm && /tmp/scan_libm
// Result in the build image (glibc 2.39, x86-64 with FMA3): 0 mismatches everywhere.
#include <cstdio>
#include "../oracle/oracle_common.h"
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
int main() {
  long nc = 0, ns = 0, n = 0;
  for (int neg = 0; neg < 2; neg++)
    for (uint32_t u = 0; u <= f2u(119.9f); u++) {
      const float x = u2f(u | (neg ? 0x80000000u : 0));
      nc += f2u(::cosf(x)) != f2u(plvio::glibcm::cosf(x));
      ns += f2u(::sinf(x)) != f2u(plvio::glibcm::sinf(x));
      n++;
    }
  printf("sinf/cosf: %ld floats, cos mismatches %ld, sin mismatches %ld\n", n, nc, ns);
  long nt = 0;
  for (uint64_t u = 0; u <= 0xffffffffull; u++) {
    const float x = u2f((uint32_t)u);
    if (x != x) continue;
    nt += f2u(::atanf(x)) != f2u(plvio::glibcm::atanf(x));
  }
  printf("atanf: all floats, mismatches %ld\n", nt);
  long na = 0, nn = 0;
  uint64_t st = 88172645463325252ull;
  for (long i = 0; i < 600000000; i++) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    float y, x;
    if (i % 3 == 0) { y = (float)((int)(st & 0xffff) - 32768) * 0.125f; x = (float)((int)((st >> 16) & 0xffff) - 32768) * 0.125f; }
    else if (i % 3 == 1) { y = u2f((uint32_t)(st >> 32)); x = u2f((uint32_t)st); if (!(std::fabs(y) < 1e30f) || !(std::fabs(x) < 1e30f)) continue; }
    else { y = (float)((double)(st & 0xffffff) / 16777216.0 * 1500.0 - 750.0); x = (float)((double)((st >> 24) & 0xffffff) / 16777216.0 * 1500.0 - 750.0); }
    na += f2u(::atan2f(y, x)) != f2u(plvio::glibcm::atan2f(y, x));
    nn++;
  }
  printf("atan2f: %ld pairs, mismatches %ld\n", nn, na);
  return (nc | ns | nt | na) != 0;
}
