// Edge convolutions of the SEANet stacks: the two HBM-bound layers that touch raw audio.
//
//  conv_in : encoder.model.0  SConv1d(C -> 32, k7)  reference modules/seanet.py:112-115, conv.py:202-221
//            reads the reference's channels-first audio in place (also the 48 kHz model's overlapping
//            1 s segments, model.py:168-170), divides by the segment loudness scale (model.py:184)
//            on the fly, reflect-pads by index mirroring, writes channels-last [item][T][32].
//  conv_out: decoder.model.15 SConv1d(32 -> C, k7, norm='none')  reference modules/seanet.py:225-228
//            reads channels-last [item][T][32], writes channels-first audio, times the segment scale
//            (model.py:244-245).
//
// Both move ~132 bytes per sample and do < 500 FLOP per sample: they are judged on GB/s.
#include "common.cuh"

namespace ecb {
namespace {

constexpr int EDGE_TILE = 128;  // samples per CTA == threads per CTA
constexpr int NF = 32;          // n_filters
constexpr int MAX_TAPS = 16;    // K * C_in <= 16 (K = 7, C_in <= 2)

__global__ void __launch_bounds__(EDGE_TILE)
conv_in_kernel(const ConvInParams p) {
  __shared__ float xs[2][EDGE_TILE + 16];
  __shared__ __align__(16) float ws[MAX_TAPS * NF];
  __shared__ __align__(16) float os[EDGE_TILE * NF];
  __shared__ float red[2][EDGE_TILE / 32];

  const int tid = threadIdx.x;
  const int item = blockIdx.y;
  const int t0 = blockIdx.x * EDGE_TILE;
  const int ntap = p.K * p.C_in;
  const float* __restrict__ xb = p.x + (long long)(item / p.n_seg) * p.batch_stride + (long long)(item % p.n_seg) * p.seg_stride;
  const float sc = p.scale ? p.scale[item] : 1.f;

  for (int i = tid; i < ntap * NF; i += EDGE_TILE) ws[i] = p.w[i];
  const int span = EDGE_TILE + p.K - 1;
  for (int c = 0; c < p.C_in; ++c) {
    for (int i = tid; i < span; i += EDGE_TILE) {
      int r = t0 + i - p.pad_left;
      float v = 0.f;
      if (r - (p.T_ref - 1) < p.T_ref) {  // beyond that the row is out of this tile's valid outputs anyway
        r = reflect_index(r, p.T_ref);
        if (r < p.T) {
          v = __ldg(xb + (long long)c * p.chan_stride + r);
          if (p.scale) v = v / sc;  // true division, as model.py:184
        }
      }
      xs[c][i] = v;
    }
  }
  __syncthreads();

  float acc[NF];
#pragma unroll
  for (int i = 0; i < NF; ++i) acc[i] = p.bias[i];
  // weight row index = j*C_in + c  (packed [K][C_in][32])
  for (int j = 0; j < p.K; ++j) {
    for (int c = 0; c < p.C_in; ++c) {
      const float xv = xs[c][tid + j];
      const float4* wr = reinterpret_cast<const float4*>(&ws[(j * p.C_in + c) * NF]);
#pragma unroll
      for (int q = 0; q < NF / 4; ++q) {
        const float4 wv = wr[q];
        acc[q * 4 + 0] = fmaf(xv, wv.x, acc[q * 4 + 0]);
        acc[q * 4 + 1] = fmaf(xv, wv.y, acc[q * 4 + 1]);
        acc[q * 4 + 2] = fmaf(xv, wv.z, acc[q * 4 + 2]);
        acc[q * 4 + 3] = fmaf(xv, wv.w, acc[q * 4 + 3]);
      }
    }
  }
  const bool valid = (t0 + tid) < p.T;
  if (p.stats) {
    float s = 0.f, sq = 0.f;
    if (valid) {
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        s += acc[i];
        sq += acc[i] * acc[i];
      }
    }
    s = warp_sum(s);
    sq = warp_sum(sq);
    if ((tid & 31) == 0) {
      red[0][tid >> 5] = s;
      red[1][tid >> 5] = sq;
    }
  }
  // stage through shared memory (XOR-swizzled 16-byte chunks) so that the global stores are coalesced
#pragma unroll
  for (int q = 0; q < NF / 4; ++q) {
    float4 v = make_float4(acc[q * 4 + 0], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
    *reinterpret_cast<float4*>(&os[tid * NF + ((q ^ (tid & 7)) << 2)]) = v;
  }
  __syncthreads();
  if (p.stats && tid == 0) {
    double s = 0.0, sq = 0.0;
    for (int i = 0; i < EDGE_TILE / 32; ++i) {
      s += (double)red[0][i];
      sq += (double)red[1][i];
    }
    const long long slot = ((long long)item * gridDim.x + blockIdx.x) * 2;
    p.stats[slot] = s;
    p.stats[slot + 1] = sq;
  }
  const long long ostride = p.out_item_stride ? p.out_item_stride : (long long)p.T * NF;
  const int rows = min(EDGE_TILE, p.T - t0);
  for (int f = tid; f < rows * (NF / 4); f += EDGE_TILE) {
    const int t = f >> 3;
    const int q = f & 7;
    const float4 v = *reinterpret_cast<const float4*>(&os[t * NF + ((q ^ (t & 7)) << 2)]);
    float4 e = v;
    if (p.out_elu) e = make_float4(elu1(v.x), elu1(v.y), elu1(v.z), elu1(v.w));   // the fused residual block reads X only
    const int g = t0 + t;   // row inside the item
    int mir[2] = {0x7fffffff, 0x7fffffff};
    if (p.halo > 0) {
      if (g >= 1 && g <= p.halo) mir[0] = -g;
      if (g <= p.T - 2 && g >= p.T - 1 - p.halo) mir[1] = 2 * (p.T - 1) - g;
    }
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      float* base = o == 0 ? p.out : p.out_elu;
      if (!base) continue;
      base += (long long)item * ostride + q * 4;
      const float4 w = o == 0 ? v : e;
      *reinterpret_cast<float4*>(base + (long long)g * NF) = w;
      if (mir[0] != 0x7fffffff) *reinterpret_cast<float4*>(base + (long long)mir[0] * NF) = w;
      if (mir[1] != 0x7fffffff) *reinterpret_cast<float4*>(base + (long long)mir[1] * NF) = w;
    }
  }
}

constexpr int OUT_LD = 36;  // padded row (floats): 16-byte aligned and conflict-free for per-row float4 reads

__global__ void __launch_bounds__(EDGE_TILE)
conv_out_kernel(const ConvOutParams p) {
  __shared__ __align__(16) float xs[(EDGE_TILE + 8) * OUT_LD];
  __shared__ __align__(16) float ws[8 * NF * 2];

  const int tid = threadIdx.x;
  const int item = blockIdx.y;
  const int t0 = blockIdx.x * EDGE_TILE;
  const float* __restrict__ ib = p.in + (long long)item * (p.in_item_stride ? p.in_item_stride : (long long)p.T * NF);
  const int span = EDGE_TILE + p.K - 1;

  for (int i = tid; i < p.K * NF * p.C_out; i += EDGE_TILE) ws[i] = p.w[i];
  for (int f = tid; f < span * (NF / 4); f += EDGE_TILE) {
    const int i = f >> 3;
    const int q = f & 7;
    int r = t0 + i - p.pad_left;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r - (p.T_ref - 1) < p.T_ref) {
      r = reflect_index(r, p.T_ref);
      if (r < p.T) v = __ldg(reinterpret_cast<const float4*>(ib + (long long)r * NF + q * 4));
    }
    *reinterpret_cast<float4*>(&xs[i * OUT_LD + q * 4]) = v;
  }
  __syncthreads();

  float acc0 = p.bias[0];
  float acc1 = p.C_out > 1 ? p.bias[1] : 0.f;
  if (p.C_out == 1) {
    for (int j = 0; j < p.K; ++j) {
#pragma unroll
      for (int q = 0; q < NF / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(&xs[(tid + j) * OUT_LD + q * 4]);
        const float4 wv = *reinterpret_cast<const float4*>(&ws[j * NF + q * 4]);
        acc0 = fmaf(v.x, wv.x, acc0);
        acc0 = fmaf(v.y, wv.y, acc0);
        acc0 = fmaf(v.z, wv.z, acc0);
        acc0 = fmaf(v.w, wv.w, acc0);
      }
    }
  } else {
    for (int j = 0; j < p.K; ++j) {
#pragma unroll
      for (int q = 0; q < NF / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(&xs[(tid + j) * OUT_LD + q * 4]);
        // weights packed [K][32][2]
        const float4 w01 = *reinterpret_cast<const float4*>(&ws[(j * NF + q * 4) * 2]);
        const float4 w23 = *reinterpret_cast<const float4*>(&ws[(j * NF + q * 4) * 2 + 4]);
        acc0 = fmaf(v.x, w01.x, acc0); acc1 = fmaf(v.x, w01.y, acc1);
        acc0 = fmaf(v.y, w01.z, acc0); acc1 = fmaf(v.y, w01.w, acc1);
        acc0 = fmaf(v.z, w23.x, acc0); acc1 = fmaf(v.z, w23.y, acc1);
        acc0 = fmaf(v.w, w23.z, acc0); acc1 = fmaf(v.w, w23.w, acc1);
      }
    }
  }
  const int t = t0 + tid;
  if (t < p.T) {
    const float sc = p.scale ? p.scale[item] : 1.f;
    float* __restrict__ ob = p.out + (long long)item * p.C_out * p.T;
    ob[t] = p.scale ? acc0 * sc : acc0;
    if (p.C_out > 1) ob[(long long)p.T + t] = p.scale ? acc1 * sc : acc1;
  }
}

}  // namespace

int conv_in_stat_slots(const ConvInParams& p) { return (int)cdiv(p.T, EDGE_TILE); }

int launch_conv_in(const ConvInParams& p, cudaStream_t stream) {
  ECB_REQUIRE(p.C_in >= 1 && p.C_in <= 2 && p.K * p.C_in <= MAX_TAPS && p.K <= 16,
              "conv_in: unsupported C_in=%d K=%d", p.C_in, p.K);
  ECB_REQUIRE(p.T >= 1 && p.T_ref >= p.T && p.T_ref > p.K - 1, "conv_in: bad reflection length %d for T=%d", p.T_ref, p.T);
  ECB_REQUIRE(p.n_items > 0 && p.n_items <= 65535, "conv_in: bad item count %d", p.n_items);
  ECB_REQUIRE(p.out || p.out_elu, "conv_in: no output");
  ECB_REQUIRE(p.halo == 0 || p.T > p.halo, "conv_in: %d samples are too few for a %d-row halo", p.T, p.halo);
  dim3 grid((unsigned)cdiv(p.T, EDGE_TILE), (unsigned)p.n_items);
  const double rows_in = (double)p.T * p.n_items;
  ProfScope prof(PROF_CONV_IN, stream, 2.0 * rows_in * NF * p.K * p.C_in, 4.0 * rows_in * (p.C_in + NF * ((p.out ? 1 : 0) + (p.out_elu ? 1 : 0))));
  conv_in_kernel<<<grid, EDGE_TILE, 0, stream>>>(p);
  ECB_LAUNCHED();
  return 0;
}

int launch_conv_out(const ConvOutParams& p, cudaStream_t stream) {
  ECB_REQUIRE(p.C_out >= 1 && p.C_out <= 2 && p.K <= 8, "conv_out: unsupported C_out=%d K=%d", p.C_out, p.K);
  ECB_REQUIRE(p.T >= 1 && p.T_ref >= p.T && p.T_ref > p.K - 1, "conv_out: bad reflection length %d for T=%d", p.T_ref, p.T);
  ECB_REQUIRE(p.n_items > 0 && p.n_items <= 65535, "conv_out: bad item count %d", p.n_items);
  dim3 grid((unsigned)cdiv(p.T, EDGE_TILE), (unsigned)p.n_items);
  const double rows_out = (double)p.T * p.n_items;
  ProfScope prof(PROF_CONV_OUT, stream, 2.0 * rows_out * NF * p.K * p.C_out, 4.0 * rows_out * (p.C_out + NF));
  conv_out_kernel<<<grid, EDGE_TILE, 0, stream>>>(p);
  ECB_LAUNCHED();
  return 0;
}

}  // namespace ecb
