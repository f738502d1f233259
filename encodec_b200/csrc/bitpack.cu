// n-bit packing of RVQ codes for the .ecdc stream without entropy coding.
//
// Replaces the Python loops of compress_to_file / decompress_from_file (reference compress.py:66-89,130-155) around
// binary.BitPacker / BitUnpacker (binary.py:55-122): for one frame, values are pushed time-major
// (for t: for k: push(codes[k][t])), each `bits` wide, least-significant bit first into a little-endian bit stream;
// the last partial byte is zero-padded (BitPacker.flush). Byte j therefore holds stream bits [8j, 8j+8), value v
// occupies stream bits [v*bits, (v+1)*bits). One thread per output byte (pack) / per value (unpack): pure byte work,
// HBM-bound and tiny next to the codec itself.
#include "common.cuh"

namespace ecb {
namespace {

__global__ void pack_codes_kernel(const long long* __restrict__ codes, long long k_stride, long long t_stride, int K, long long T,
                                  int bits, unsigned char* __restrict__ out, long long n_bytes) {
  const long long n_vals = (long long)K * T;
  const unsigned long long mask = (1ull << bits) - 1ull;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n_bytes; j += (long long)gridDim.x * blockDim.x) {
    const long long bit0 = j * 8;
    long long v = bit0 / bits;
    unsigned int byte = 0;
    for (; v < n_vals && v * bits < bit0 + 8; ++v) {
      const long long t = v / K;
      const int k = (int)(v - t * K);
      const unsigned long long val = (unsigned long long)codes[k * k_stride + t * t_stride] & mask;
      const long long shift = v * bits - bit0;   // position of the value's bit 0 relative to this byte
      byte |= (unsigned int)((shift >= 0 ? (val << shift) : (val >> (-shift))) & 0xffull);
    }
    out[j] = (unsigned char)byte;
  }
}

__global__ void unpack_codes_kernel(const unsigned char* __restrict__ in, long long n_bytes, int K, long long T, int bits,
                                    long long* __restrict__ codes, long long k_stride, long long t_stride) {
  const long long n_vals = (long long)K * T;
  const unsigned long long mask = (1ull << bits) - 1ull;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n_vals; v += (long long)gridDim.x * blockDim.x) {
    const long long bit0 = v * bits;
    const long long b0 = bit0 >> 3;
    unsigned long long acc = 0;
    for (int i = 0; i < 4 && b0 + i < n_bytes; ++i) acc |= (unsigned long long)in[b0 + i] << (8 * i);   // bits <= 24 fit 4 bytes
    const long long t = v / K;
    const int k = (int)(v - t * K);
    codes[k * k_stride + t * t_stride] = (long long)((acc >> (bit0 & 7)) & mask);
  }
}

}  // namespace

int launch_pack_codes(const long long* codes, long long k_stride, long long t_stride, int K, long long T, int bits,
                      unsigned char* out, cudaStream_t s) {
  ECB_REQUIRE(K > 0 && T > 0 && bits >= 1 && bits <= 24, "pack_codes: bad K=%d T=%lld bits=%d", K, T, bits);
  const long long n_bytes = ((long long)K * T * bits + 7) / 8;
  ProfScope prof(PROF_MISC, s, 0.0, 8.0 * K * T + (double)n_bytes);
  const long long blocks = cdiv(n_bytes, 256);
  pack_codes_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, s>>>(codes, k_stride, t_stride, K, T, bits, out, n_bytes);
  ECB_LAUNCHED();
  return 0;
}

int launch_unpack_codes(const unsigned char* in, long long n_bytes, int K, long long T, int bits, long long* codes,
                        long long k_stride, long long t_stride, cudaStream_t s) {
  ECB_REQUIRE(K > 0 && T > 0 && bits >= 1 && bits <= 24, "unpack_codes: bad K=%d T=%lld bits=%d", K, T, bits);
  ECB_REQUIRE(n_bytes * 8 >= (long long)K * T * bits, "unpack_codes: the stream ended sooner than expected");
  ProfScope prof(PROF_MISC, s, 0.0, 8.0 * K * T + (double)n_bytes);
  const long long blocks = cdiv((long long)K * T, 256);
  unpack_codes_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, s>>>(in, n_bytes, K, T, bits, codes, k_stride, t_stride);
  ECB_LAUNCHED();
  return 0;
}

}  // namespace ecb
