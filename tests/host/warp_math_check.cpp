// Host-side check of csrc/warp_math.h: the Markstein division must equal IEEE division bit for bit, and the full
// coordinate replay must equal the straightforward "mul, div, sub, add, div, mul" sequence.
// Build: g++ -O2 -ffp-contract=off -o warp_math_check warp_math_check.cpp ; exit code 0 = pass.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "../../video-frame-interpolation_b200/csrc/warp_math.h"

static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static float rnd_unit() { return (float)((rnd() >> 40) / 16777216.0); }

static float ref_coord(int pix, float disp, long long size) {
  volatile float v = (float)pix + disp;
  volatile float denom = (float)(size - 1 > 1 ? size - 1 : 1);
  volatile float g = 2.0f * v;
  g = g / denom;
  g = g - 1.0f;
  volatile float t = g + 1.0f;
  t = t / 2.0f;
  volatile float i = t * (float)(size - 1);
  float hi = (float)size + 4.0f;
  float r = i; if (!(r >= -4.0f)) r = -4.0f; if (r > hi) r = hi; return r;
}

int main(int argc, char** argv) {
  long long n = argc > 1 ? atoll(argv[1]) : 4000000;
  const long long sizes[] = {1, 2, 3, 7, 24, 32, 64, 131, 255, 256, 257, 480, 640, 1080, 1920, 2160, 3840, 4096, 8191, 16384};
  long long bad = 0, total = 0;
  for (long long size : sizes) {
    WarpAxis ax = make_warp_axis(size);
    for (long long it = 0; it < n / 20; ++it) {
      int pix = (int)(rnd() % (unsigned long long)size);
      float mag = (it % 4 == 0) ? 0.05f : (it % 4 == 1) ? 3.0f : (it % 4 == 2) ? 64.0f : (float)size;
      float disp = (rnd_unit() * 2.0f - 1.0f) * mag;
      if (it % 97 == 0) disp = (float)((int)(rnd() % 9) - 4);  // exact integers
      // division alone
      float a = 2.0f * ((float)pix + disp);
      volatile float qref = a / ax.denom;
      float q = vfi_div_exact(a, ax);
      if (bits(q) != bits(qref)) { if (bad < 10) printf("div mismatch a=%a d=%g\n", a, ax.denom); ++bad; }
      float c = vfi_warp_coord(pix, disp, ax), cr = ref_coord(pix, disp, size);
      if (bits(c) != bits(cr)) { if (bad < 10) printf("coord mismatch size=%lld pix=%d disp=%a\n", size, pix, disp); ++bad; }
      ++total;
    }
  }
  // raw random bit patterns for the numerator (normal range)
  for (long long it = 0; it < n; ++it) {
    uint32_t u = (uint32_t)rnd(); float a; memcpy(&a, &u, 4);
    if (!(a == a) || a - a != 0.0f) continue;
    float fa = a < 0 ? -a : a;
    if (fa < 1e-30f || fa > 1e30f) continue;
    long long size = sizes[it % 20];
    WarpAxis ax = make_warp_axis(size);
    volatile float qref = a / ax.denom;
    if (bits(vfi_div_exact(a, ax)) != bits(qref)) { if (bad < 10) printf("div mismatch a=%a d=%g\n", a, ax.denom); ++bad; }
    ++total;
  }
  printf("checked %lld cases, %lld mismatches\n", total, bad);
  return bad ? 1 : 0;
}
