/* Exhaustive-ish check of the division used by the input stage (ray3d_b200/csrc/r3d_stage_kernels.cu, div_by):
 *   q0 = a * rb;  r0 = fma(-q0, b, a);  q1 = fma(r0, rb, q0);  r1 = fma(-q1, b, a);  q2 = fma(r1, rb, q1)
 * with rb = RN(1 / b) must equal the IEEE quotient RN(a / b) (Markstein: a faithful q1 and a correctly rounded
 * reciprocal make the last step exact).  Operands follow the path: a = (double)float32 pixel - (double)float32 centre,
 * b = (double)float32 focal length; plus uniformly random doubles as a stress test.
 *   gcc -O2 -mfma -o /tmp/check_fma_division scripts/check_fma_division.c -lm && /tmp/check_fma_division */
#include <math.h>
#include <stdint.h>
#include <stdio.h>

static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static uint64_t rnd(void) {   /* xorshift128+ */
  uint64_t x = s[0], y = s[1];
  s[0] = y; x ^= x << 23; s[1] = x ^ y ^ (x >> 17) ^ (y >> 26);
  return s[1] + y;
}
static double u01(void) { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }

static double div_by(double a, double b, double rb) {
  double q = a * rb;
  double r = fma(-q, b, a);
  q = fma(r, rb, q);
  r = fma(-q, b, a);
  return fma(r, rb, q);
}

int main(void) {
  long bad = 0, n = 0;
  for (long i = 0; i < 400000000L; ++i) {
    double a, b;
    if (i & 1) {   /* path-like operands */
      float px = (float)(u01() * 4096.0 - 1024.0), c = (float)(u01() * 2048.0), f = (float)(200.0 + u01() * 4000.0);
      a = (double)px - (double)c; b = (double)f;
    } else {       /* random mantissas, moderate exponents */
      a = ldexp(1.0 + u01(), (int)(rnd() % 40) - 20) * ((rnd() & 1) ? 1 : -1);
      b = ldexp(1.0 + u01(), (int)(rnd() % 40) - 20);
    }
    const double rb = 1.0 / b;
    if (div_by(a, b, rb) != a / b) { if (bad < 5) printf("MISMATCH a=%a b=%a got=%a want=%a\n", a, b, div_by(a, b, rb), a / b); ++bad; }
    ++n;
  }
  printf("%ld cases, %ld mismatches\n", n, bad);
  return bad != 0;
}
