// Range arithmetic of the entropy-coded .ecdc stream, shared by the host coder and the device decoder.
//
// Restates ArithmeticCoder / ArithmeticDecoder of the reference (encodec/quantization/ac.py:56-260) on 64-bit integers:
// the reference keeps (low, high[, current]) as Python integers below 2^(max_bit + 1) with max_bit <= 61 asserted
// (ac.py:156), so uint64 holds every value it can reach. The two places where the reference leaves integers are
//   effective_low  = ceil (range_low  * (delta / 2^total_range_bits))     (ac.py:145, 240)
//   effective_high = floor(range_high * (delta / 2^total_range_bits))     (ac.py:146, 241)
// in Python floats: `delta / 2^bits` is the correctly rounded double of delta scaled by a power of two, the product is ONE
// IEEE double multiplication. For the coder's default 24 range bits (the only setting compress.py uses) that product is exact
// and both expressions are evaluated in 64-bit integers (make_scale below); otherwise they are reproduced operation for
// operation in doubles (explicitly rounded intrinsics on the device, so that no contraction can change them). Bits travel
// through BitPacker(bits=1) (binary.py:55-89): bit i of the stream is bit (i % 8) of byte i / 8.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ECB_AC_HD __host__ __device__ __forceinline__
#else
#define ECB_AC_HD inline
#endif

namespace ecb {
namespace ac {

enum Status { AC_OK = 0, AC_EOF = 1, AC_SEARCH_FAILED = 2, AC_RANGE_OVERFLOW = 3, AC_BAD_CDF = 4 };

// Exact integer form for bits <= 26: delta < 2^(bits + 1) whenever a symbol is coded (the refill loop stops at the first
// delta >= 2^bits) and every cdf value is <= 2^bits + card, so range * delta < 2^53: the double product above is exact and
// ceil / floor of it are the integer expressions below. compress.py only ever uses bits = 24. Wider settings keep the
// reference's double arithmetic operation for operation.
struct Scale {
  uint64_t delta;
  double ratio;
  int bits;
  bool exact_int;
};
ECB_AC_HD Scale make_scale(uint64_t delta, int bits) {
  Scale s;
  s.delta = delta;
  s.bits = bits;
  s.exact_int = bits <= 26 && delta < (1ull << (bits + 1));
  // delta / 2**bits (ac.py:145): conversion rounds to nearest-even like Python's int / int, the scaling is exact
#if defined(__CUDA_ARCH__)
  s.ratio = s.exact_int ? 0.0 : __dmul_rn(__ull2double_rn((unsigned long long)delta), 1.0 / (double)(1ull << bits));
#else
  s.ratio = (double)delta * (1.0 / (double)(1ull << bits));
#endif
  return s;
}
ECB_AC_HD uint64_t eff_low(int64_t range_low, const Scale& sc) {
  if (sc.exact_int) return ((uint64_t)range_low * sc.delta + ((1ull << sc.bits) - 1ull)) >> sc.bits;
#if defined(__CUDA_ARCH__)
  return (uint64_t)ceil(__dmul_rn((double)range_low, sc.ratio));
#else
  volatile double p = (double)range_low * sc.ratio;   // volatile: one rounded product, never fused or kept in extended precision
  return (uint64_t)ceil(p);
#endif
}
ECB_AC_HD uint64_t eff_high(int64_t range_high, const Scale& sc) {
  if (sc.exact_int) return ((uint64_t)range_high * sc.delta) >> sc.bits;
#if defined(__CUDA_ARCH__)
  return (uint64_t)floor(__dmul_rn((double)range_high, sc.ratio));
#else
  volatile double p = (double)range_high * sc.ratio;
  return (uint64_t)floor(p);
#endif
}

// ---- decoder (host and device) -------------------------------------------------------------------------------------
struct Decoder {
  uint64_t low, high, current;
  int64_t bit_pos;      // next stream bit
  int32_t max_bit;
  int32_t status;       // sticky Status
};

ECB_AC_HD void decoder_init(Decoder& d, int64_t first_bit) {
  d.low = 0;
  d.high = 0;
  d.current = 0;
  d.bit_pos = first_bit;
  d.max_bit = -1;
  d.status = AC_OK;
}

// The two bit loops of the reference in closed form (same result, O(1) per symbol):
//  * refill (ac.py:224-231): `while delta < 2^bits: low *= 2; high = high * 2 + 1; current = current * 2 + bit; max_bit += 1`
//    runs n = max(0, bits + 1 - bit_length(delta)) times; the n stream bits enter `current` first-read = most significant;
//  * flush (ac.py:195-212): strips the c leading bits (of the max_bit + 1 wide representation) that low and high share --
//    current lies between them, so it shares them too.
ECB_AC_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}
ECB_AC_HD uint64_t brev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __brevll(x);
#else
  x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
  return __builtin_bswap64(x);
#endif
}
// returns false with d.status set when the stream ends first (the reference returns None: "ended sooner than expected")
ECB_AC_HD bool refill(Decoder& d, const uint8_t* data, int64_t n_bits, int bits) {
  const uint64_t delta = d.high - d.low + 1;
  const int len = 64 - clz64(delta);
  const int n = bits + 1 - len;
  if (n <= 0) return true;
  if (d.bit_pos + n > n_bits) {
    d.status = AC_EOF;
    return false;
  }
  if (d.max_bit + n > 61) {
    d.status = AC_RANGE_OVERFLOW;
    return false;
  }
  const int64_t byte0 = d.bit_pos >> 3, n_bytes = (n_bits + 7) >> 3;
  uint64_t w = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i)   // n <= 31 bits + 7 bits of offset
    if (byte0 + i < n_bytes) w |= (uint64_t)data[byte0 + i] << (8 * i);
  const uint64_t chunk = (w >> (d.bit_pos & 7)) & ((1ull << n) - 1ull);   // stream order: first bit = least significant
  d.bit_pos += n;
  d.low <<= n;
  d.high = ((d.high + 1) << n) - 1;
  d.current = (d.current << n) | (brev64(chunk) >> (64 - n));
  d.max_bit += n;
  return true;
}
ECB_AC_HD void flush_prefix(Decoder& d) {
  if (d.max_bit < 0) return;
  const uint64_t x = d.low ^ d.high;
  const int width = d.max_bit + 1;
  const int c = x ? clz64(x) - (64 - width) : width;   // common leading bits of the width-bit representations
  if (c <= 0) return;
  const uint64_t mask = (width - c) >= 64 ? ~0ull : ((1ull << (width - c)) - 1ull);
  d.low &= mask;
  d.high &= mask;
  d.current &= mask;
  d.max_bit -= c;
}

// ArithmeticDecoder.pull (ac.py:214-260). cdf: int32 [card], the quantised cdf of this symbol (cdf[i] = upper bound,
// exclusive, of symbol i's range). Returns the symbol, or -1 with d.status set (AC_EOF: "the stream ended sooner than
// expected", compress.py:143-144).
ECB_AC_HD int pull(Decoder& d, const uint8_t* data, int64_t n_bits, const int32_t* cdf, int card, int bits) {
  if (d.status != AC_OK) return -1;
  if (!refill(d, data, n_bits, bits)) return -1;                          // ac.py:224-231
  const Scale ratio = make_scale(d.high - d.low + 1, bits);
  int lo_idx = 0, hi_idx = card - 1, mid = 0;
  uint64_t low = 0, high = 0;
  for (;;) {                                                             // bin_search, ac.py:233-251
    if (hi_idx < lo_idx) {
      d.status = AC_SEARCH_FAILED;
      return -1;
    }
    mid = (lo_idx + hi_idx) / 2;
    const int64_t range_low = mid > 0 ? cdf[mid - 1] : 0;
    const int64_t range_high = (int64_t)cdf[mid] - 1;
    if (range_high < range_low) {                                        // an empty range cannot be produced by the cdf builder
      d.status = AC_BAD_CDF;
      return -1;
    }
    low = eff_low(range_low, ratio) + d.low;
    high = eff_high(range_high, ratio) + d.low;
    if (d.current >= low) {
      if (d.current <= high) break;
      lo_idx = mid + 1;
    } else {
      hi_idx = mid - 1;
    }
  }
  d.low = low;
  d.high = high;
  flush_prefix(d);                                                       // _flush_common_prefix, ac.py:195-212
  return mid;
}

ECB_AC_HD int64_t bytes_consumed(const Decoder& d) { return (d.bit_pos + 7) >> 3; }   // BitUnpacker reads whole bytes

// ---- coder (host) --------------------------------------------------------------------------------------------------
struct Encoder {
  uint64_t low = 0, high = 0;
  int32_t max_bit = -1;
  int32_t status = AC_OK;
  uint8_t* out = nullptr;
  int64_t cap = 0, n_bytes = 0;
  uint32_t cur = 0;
  int32_t n_cur = 0;
  bool overflow = false;

  void push_bit(uint64_t b) {                                            // BitPacker.push with bits = 1 (binary.py:69-77)
    cur |= (uint32_t)(b & 1u) << n_cur;
    if (++n_cur == 8) {
      if (n_bytes < cap) out[n_bytes] = (uint8_t)cur; else overflow = true;
      ++n_bytes;
      cur = 0;
      n_cur = 0;
    }
  }
  // ArithmeticCoder.push (ac.py:127-157): the symbol's range is [range_low, range_high_excl) of the quantised cdf
  bool push(int64_t range_low, int64_t range_high_excl, int bits) {
    if (status != AC_OK) return false;
    const uint64_t full = 1ull << bits;
    while (high - low + 1 < full) {
      if (max_bit >= 61) { status = AC_RANGE_OVERFLOW; return false; }   // the reference asserts max_bit <= 61 (ac.py:156)
      low *= 2;
      high = high * 2 + 1;
      ++max_bit;
    }
    if (range_high_excl - 1 < range_low || range_low < 0) { status = AC_BAD_CDF; return false; }
    const Scale ratio = make_scale(high - low + 1, bits);
    const uint64_t el = eff_low(range_low, ratio), eh = eff_high(range_high_excl - 1, ratio);
    high = low + eh;
    low = low + el;
    if (low > high) { status = AC_BAD_CDF; return false; }
    while (max_bit >= 0) {                                               // _flush_common_prefix, ac.py:109-125
      const uint64_t b1 = low >> max_bit, b2 = high >> max_bit;
      if (b1 != b2) break;
      low -= b1 << max_bit;
      high -= b1 << max_bit;
      --max_bit;
      push_bit(b1);
    }
    return true;
  }
  void flush() {                                                         // ac.py:159-166 + BitPacker.flush (binary.py:79-86)
    while (max_bit >= 0) {
      push_bit((low >> max_bit) & 1u);
      --max_bit;
    }
    if (n_cur) {
      if (n_bytes < cap) out[n_bytes] = (uint8_t)cur; else overflow = true;
      ++n_bytes;
      cur = 0;
      n_cur = 0;
    }
  }
};

}  // namespace ac
}  // namespace ecb
