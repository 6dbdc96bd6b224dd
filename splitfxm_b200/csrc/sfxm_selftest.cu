// sfxm_selftest.cu - device self-test of div_known (sfxm_device.cuh): compares the division by a known reciprocal
// (one multiplication + two FMAs) with the IEEE quotient a / b, bit for bit, on pseudo-random, adversarial and
// nearest-to-midpoint operands.
#include <cuda_runtime.h>
#include <cstdint>
#include "sfxm_device.cuh"
#include "sfxm_host.h"

namespace sfxm {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// mode 0: random mantissas, exponents in 2^+-60;  mode 1: divisor mantissa all ones / all zeros (+ few low
// bits), numerator a product of the divisor and a random value nudged by -1..1 ulp (hard rounding cases);
// mode 2: quotients as close to a rounding boundary as a quotient of two doubles can be.  For an odd 53-bit divisor
// Bi and odd k, the 54-bit odd M = -k Bi^-1 mod 2^54 (a midpoint of the double grid, as an integer) satisfies
// M Bi + k = A 2^54, so with b = Bi 2^-52, a = A 2^-51:  a / b = M 2^-53 + k / (2^105 b) - 2^-53 ulp from the midpoint,
// the case the error analysis of div_known turns on.  Divisors: random, or within 512 ulp below 2 / above 1 (where
// that analysis has the least slack).  Half the cases use the modulus 2^53 instead (numerators below 2).
__global__ void div_selftest_kernel(int64_t n, uint64_t seed, int mode, unsigned long long *mismatch) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t h1 = mix64(seed + 2 * (uint64_t)t), h2 = mix64(seed + 2 * (uint64_t)t + 1);
  const uint64_t mant = 0x000fffffffffffffull;
  double a, b;
  if (mode == 0) {
    const uint64_t ea = 1023 - 60 + (h1 >> 52) % 121, eb = 1023 - 60 + (h2 >> 52) % 121;
    a = __longlong_as_double((long long)(((h1 >> 63) << 63) | (ea << 52) | (h1 & mant)));
    b = __longlong_as_double((long long)(((h2 >> 63) << 63) | (eb << 52) | (h2 & mant)));
  } else if (mode == 2) {
    uint64_t Bi = (h2 & mant) | (1ull << 52) | 1ull;
    const unsigned sel = (unsigned)(h2 >> 60);
    if (sel < 4) Bi = (1ull << 53) - 1ull - 2ull * ((h2 >> 44) & 0xff);
    else if (sel < 7) Bi = (1ull << 52) + 1ull + 2ull * ((h2 >> 44) & 0xff);
    uint64_t inv = Bi;                                   // Newton iteration for the inverse modulo 2^64
#pragma unroll
    for (int it = 0; it < 6; ++it) inv *= 2ull - Bi * inv;
    const long long k = (long long)(2 * ((h1 >> 58) & 3) + 1) * (((h1 >> 57) & 1) ? -1 : 1);   // +-1, 3, 5, 7
    const int bits = ((h1 >> 56) & 1) ? 54 : 53;
    const uint64_t modmask = (1ull << bits) - 1ull;
    uint64_t M = ((uint64_t)(-k) * inv) & modmask;
    if (bits == 53) M += 1ull << 53;
    a = 0.0;
    b = (double)Bi * 0x1p-52;
    if (M >> 53) {                                       // a 54-bit odd integer: a midpoint in [1, 2)
      uint64_t lo = M * Bi, hi = __umul64hi(M, Bi);
      const uint64_t lo2 = lo + (uint64_t)k;
      if (k > 0) hi += (lo2 < lo); else hi -= (lo2 > lo);
      lo = lo2;
      if ((lo & modmask) == 0 && (hi >> (bits == 54 ? 43 : 42)) == 0) {   // A = P >> bits fits in 53 bits
        const uint64_t A = (hi << (64 - bits)) | (lo >> bits);
        a = (double)A * (bits == 54 ? 0x1p-51 : 0x1p-52);
      }
    }
    const int ea = (int)((h1 >> 40) % 81) - 40, eb = (int)((h2 >> 32) % 81) - 40;
    a = ldexp(a, ea) * (((h1 >> 39) & 1) ? -1.0 : 1.0);
    b = ldexp(b, eb) * (((h2 >> 31) & 1) ? -1.0 : 1.0);
  } else {
    const uint64_t eb = 1023 - 20 + (h2 >> 52) % 41;
    const uint64_t low = h2 & 0x7;
    const uint64_t mb = (h2 & 8) ? (mant ^ low) : low;
    b = __longlong_as_double((long long)((eb << 52) | mb));
    const double q = __longlong_as_double((long long)((1023ull << 52) | (h1 & mant)));
    const double p = q * b;
    const long long nudge = (long long)((h1 >> 52) % 3) - 1;
    a = __longlong_as_double(__double_as_longlong(p) + nudge);
    if ((h1 & 0xff0000ull) == 0) a = 0.0;   // exact-zero numerators are common in flat regions
  }
  const double rb = 1.0 / b;
  const double fast = div_known<true>(a, b, rb);
  const double ieee = a / b;
  const bool same = __double_as_longlong(fast) == __double_as_longlong(ieee) || (fast == 0.0 && ieee == 0.0);
  if (!same) atomicAdd(mismatch, 1ull);   // the sign of an exact zero is not tracked (0.0 == -0.0)
}

}  // namespace sfxm

extern "C" int sfxm_selftest_div(int64_t n, uint64_t seed, int32_t mode, uint64_t *mismatch_dev, void *stream) {
  if (!mismatch_dev || n < 0) return sfxm::set_error("sfxm_selftest_div: bad argument");
  if (n == 0) return 0;
  sfxm::div_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n, seed, mode, (unsigned long long *)mismatch_dev);
  return sfxm::check_launch("sfxm_selftest_div");
}
