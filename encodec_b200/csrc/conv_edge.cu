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


// Register-blocked form for the standard kernel size (K = 7): thread (cg, ts) owns 8 output channels (cg = tid & 3) of 4
// consecutive samples (ts = tid >> 2); its 7 x C_in x 8 weights live in registers, the input window comes from shared memory
// with three vector loads, and the 4 lanes of a sample write its 128-byte channels-last row directly (no staging). ~75
// instructions per sample instead of ~1200: the kernel is then bound by its 132 bytes per sample. Same summation order as
// the generic kernel (bias, then taps in order), so both give bit-identical results.
template <int C_IN>
__global__ void __launch_bounds__(EDGE_TILE)
conv_in7_kernel(const ConvInParams p) {
  constexpr int K = 7, S = 4, CG = 8;
  __shared__ __align__(16) float xs[C_IN][EDGE_TILE + 16];
  __shared__ float red[2][EDGE_TILE / 32];
  const int tid = threadIdx.x;
  const int item = blockIdx.y;
  const int t0 = blockIdx.x * EDGE_TILE;
  const float* __restrict__ xb = p.x + (long long)(item / p.n_seg) * p.batch_stride + (long long)(item % p.n_seg) * p.seg_stride;
  const float sc = p.scale ? p.scale[item] : 1.f;
  const int cg = tid & 3, ts = tid >> 2;
  float w[K * C_IN][CG];
#pragma unroll
  for (int j = 0; j < K * C_IN; ++j) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.w + j * NF + cg * CG));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.w + j * NF + cg * CG + 4));
    w[j][0] = a.x; w[j][1] = a.y; w[j][2] = a.z; w[j][3] = a.w;
    w[j][4] = b.x; w[j][5] = b.y; w[j][6] = b.z; w[j][7] = b.w;
  }
  const int span = EDGE_TILE + K - 1;
#pragma unroll
  for (int c = 0; c < C_IN; ++c) {
    for (int i = tid; i < EDGE_TILE + 16; i += EDGE_TILE) {
      float v = 0.f;
      int r = t0 + i - p.pad_left;
      if (i < span && r - (p.T_ref - 1) < p.T_ref) {  // beyond that the row is out of this tile's valid outputs anyway
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
  float acc[S][CG];
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.bias + cg * CG));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + cg * CG + 4));
#pragma unroll
    for (int s = 0; s < S; ++s) {
      acc[s][0] = a.x; acc[s][1] = a.y; acc[s][2] = a.z; acc[s][3] = a.w;
      acc[s][4] = b.x; acc[s][5] = b.y; acc[s][6] = b.z; acc[s][7] = b.w;
    }
  }
  float xw[C_IN][S + K - 1 + 2];   // 12 floats: three 16-byte loads
#pragma unroll
  for (int c = 0; c < C_IN; ++c) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(&xs[c][ts * S + q * 4]);
      xw[c][q * 4 + 0] = v.x; xw[c][q * 4 + 1] = v.y; xw[c][q * 4 + 2] = v.z; xw[c][q * 4 + 3] = v.w;
    }
  }
  // weight row index = j*C_in + c  (packed [K][C_in][32])
#pragma unroll
  for (int j = 0; j < K; ++j)
#pragma unroll
    for (int c = 0; c < C_IN; ++c)
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int i = 0; i < CG; ++i) acc[s][i] = fmaf(xw[c][s + j], w[j * C_IN + c][i], acc[s][i]);
  if (p.stats) {
    float sm = 0.f, sq = 0.f;
#pragma unroll
    for (int s = 0; s < S; ++s)
      if (t0 + ts * S + s < p.T) {
#pragma unroll
        for (int i = 0; i < CG; ++i) {
          sm += acc[s][i];
          sq += acc[s][i] * acc[s][i];
        }
      }
    sm = warp_sum(sm);
    sq = warp_sum(sq);
    if ((tid & 31) == 0) {
      red[0][tid >> 5] = sm;
      red[1][tid >> 5] = sq;
    }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0, b = 0.0;
      for (int i = 0; i < EDGE_TILE / 32; ++i) {
        a += (double)red[0][i];
        b += (double)red[1][i];
      }
      const long long slot = ((long long)item * gridDim.x + blockIdx.x) * 2;
      p.stats[slot] = a;
      p.stats[slot + 1] = b;
    }
  }
  const long long ostride = p.out_item_stride ? p.out_item_stride : (long long)p.T * NF;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int g = t0 + ts * S + s;   // row inside the item
    if (g >= p.T) break;
    long long mir[2] = {0, 0};
    bool has[2] = {false, false};
    if (p.halo > 0) {
      if (g >= 1 && g <= p.halo) { mir[0] = -g; has[0] = true; }
      if (g <= p.T - 2 && g >= p.T - 1 - p.halo) { mir[1] = 2LL * (p.T - 1) - g; has[1] = true; }
    }
    const float4 v0 = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
    const float4 v1 = make_float4(acc[s][4], acc[s][5], acc[s][6], acc[s][7]);
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      float* base = o == 0 ? p.out : p.out_elu;
      if (!base) continue;
      base += (long long)item * ostride + cg * CG;
      float4 a = v0, b = v1;
      if (o == 1) {
        a = make_float4(elu1(v0.x), elu1(v0.y), elu1(v0.z), elu1(v0.w));
        b = make_float4(elu1(v1.x), elu1(v1.y), elu1(v1.z), elu1(v1.w));
      }
      *reinterpret_cast<float4*>(base + (long long)g * NF) = a;
      *reinterpret_cast<float4*>(base + (long long)g * NF + 4) = b;
#pragma unroll
      for (int m = 0; m < 2; ++m)
        if (has[m]) {
          *reinterpret_cast<float4*>(base + mir[m] * NF) = a;
          *reinterpret_cast<float4*>(base + mir[m] * NF + 4) = b;
        }
    }
  }
}

constexpr int OUT_LD = 36;  // padded row (floats): 16-byte aligned and conflict-free for per-row float4 reads

__global__ void __launch_bounds__(EDGE_TILE)
conv_out_kernel(const ConvOutParams p) {
  __shared__ __align__(16) float xs[(EDGE_TILE + 16) * OUT_LD];
  __shared__ __align__(16) float ws[8 * NF * 2];

  const int tid = threadIdx.x;
  const int item = blockIdx.y;
  const int t0 = blockIdx.x * EDGE_TILE;
  const float* __restrict__ ib = p.in + (long long)item * (p.in_item_stride ? p.in_item_stride : (long long)p.T * NF);
  const int span = EDGE_TILE + p.K - 1;

  for (int i = tid; i < p.K * NF * p.C_out; i += EDGE_TILE) ws[i] = p.w[i];
  // stage the input tile: all of a thread's 16-byte loads are issued before the first is stored (the kernel lives on memory
  // level parallelism: one load at a time per thread left it at a third of the HBM rate)
  {
    constexpr int NLD = ((EDGE_TILE + 15) * (NF / 4) + EDGE_TILE - 1) / EDGE_TILE;   // K <= 16
    float4 v[NLD];
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int f = tid + k * EDGE_TILE;
      const int i = f >> 3;
      const int q = f & 7;
      int r = t0 + i - p.pad_left;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < span * (NF / 4) && r - (p.T_ref - 1) < p.T_ref) {
        r = reflect_index(r, p.T_ref);
        if (r < p.T) v[k] = __ldg(reinterpret_cast<const float4*>(ib + (long long)r * NF + q * 4));
      }
    }
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int f = tid + k * EDGE_TILE;
      if (f < span * (NF / 4)) *reinterpret_cast<float4*>(&xs[(f >> 3) * OUT_LD + (f & 7) * 4]) = v[k];
    }
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


// Register-blocked form for K = 7: thread (cg, ts) takes 8 input channels (cg = tid & 3) of the 4 consecutive output samples
// 4 ts .. 4 ts + 3; its 7 x 8 x C_out weights live in registers, the 10-row window is read once from shared memory (two
// 16-byte loads per row, their order alternating with ts so that a warp's loads cover all banks), the four channel groups
// are added with two shuffles, and lane cg stores sample 4 ts + cg: a warp writes 32 consecutive samples.
template <int C_OUT>
__global__ void __launch_bounds__(EDGE_TILE)
conv_out7_kernel(const ConvOutParams p) {
  constexpr int K = 7, S = 4, CG = 8;
  __shared__ __align__(16) float xs[(EDGE_TILE + 8) * OUT_LD];
  const int tid = threadIdx.x;
  const int item = blockIdx.y;
  const int t0 = blockIdx.x * EDGE_TILE;
  const float* __restrict__ ib = p.in + (long long)item * (p.in_item_stride ? p.in_item_stride : (long long)p.T * NF);
  const int span = EDGE_TILE + K - 1;
  const int cg = tid & 3, ts = tid >> 2;
  float w[K][CG][C_OUT];   // weights packed [K][32][C_out]
#pragma unroll
  for (int j = 0; j < K; ++j)
#pragma unroll
    for (int i = 0; i < CG; ++i)
#pragma unroll
      for (int co = 0; co < C_OUT; ++co) w[j][i][co] = __ldg(p.w + (j * NF + cg * CG + i) * C_OUT + co);
  // stage the input tile: all of a thread's 16-byte loads are issued before the first is stored (the kernel lives on memory
  // level parallelism: one load at a time per thread left it at a third of the HBM rate)
  {
    constexpr int NLD = ((EDGE_TILE + 15) * (NF / 4) + EDGE_TILE - 1) / EDGE_TILE;   // K <= 16
    float4 v[NLD];
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int f = tid + k * EDGE_TILE;
      const int i = f >> 3;
      const int q = f & 7;
      int r = t0 + i - p.pad_left;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < span * (NF / 4) && r - (p.T_ref - 1) < p.T_ref) {
        r = reflect_index(r, p.T_ref);
        if (r < p.T) v[k] = __ldg(reinterpret_cast<const float4*>(ib + (long long)r * NF + q * 4));
      }
    }
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int f = tid + k * EDGE_TILE;
      if (f < span * (NF / 4)) *reinterpret_cast<float4*>(&xs[(f >> 3) * OUT_LD + (f & 7) * 4]) = v[k];
    }
  }
  __syncthreads();
  float acc[S][C_OUT];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int co = 0; co < C_OUT; ++co) acc[s][co] = 0.f;
  const int first = ts & 1;   // which 16-byte half of the 8 channels is loaded first
#pragma unroll
  for (int rr = 0; rr < S + K - 1; ++rr) {
    const float* row = &xs[(ts * S + rr) * OUT_LD + cg * CG];
    const float4 va = *reinterpret_cast<const float4*>(row + first * 4);
    const float4 vb = *reinterpret_cast<const float4*>(row + (1 - first) * 4);
    const float4 lo = first ? vb : va, hi = first ? va : vb;
    const float x8[CG] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int j = rr - s;   // tap of sample s that reads this row
      if (j < 0 || j >= K) continue;
#pragma unroll
      for (int i = 0; i < CG; ++i)
#pragma unroll
        for (int co = 0; co < C_OUT; ++co) acc[s][co] = fmaf(x8[i], w[j][i][co], acc[s][co]);
    }
  }
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int co = 0; co < C_OUT; ++co) {
      acc[s][co] += __shfl_xor_sync(0xffffffffu, acc[s][co], 1);
      acc[s][co] += __shfl_xor_sync(0xffffffffu, acc[s][co], 2);
    }
  const int t = t0 + ts * S + cg;
  if (t < p.T) {
    const float sc = p.scale ? p.scale[item] : 1.f;
    float* __restrict__ ob = p.out + (long long)item * C_OUT * p.T;
#pragma unroll
    for (int co = 0; co < C_OUT; ++co) {
      float v = acc[0][co];
      if (cg == 1) v = acc[1][co];
      if (cg == 2) v = acc[2][co];
      if (cg == 3) v = acc[3][co];
      v += __ldg(p.bias + co);
      ob[(long long)co * p.T + t] = p.scale ? v * sc : v;
    }
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
  // mono: the register-blocked kernel (4.1 TB/s against 2.7); stereo keeps the generic one (twice the weights per thread
  // cost the blocked form its occupancy: measured 1.85 against 2.03 TB/s)
  if (p.K == 7 && p.C_in == 1) conv_in7_kernel<1><<<grid, EDGE_TILE, 0, stream>>>(p);
  else conv_in_kernel<<<grid, EDGE_TILE, 0, stream>>>(p);
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
  if (p.K == 7 && p.C_out == 1) conv_out7_kernel<1><<<grid, EDGE_TILE, 0, stream>>>(p);
  else if (p.K == 7 && p.C_out == 2) conv_out7_kernel<2><<<grid, EDGE_TILE, 0, stream>>>(p);   // stereo: 2.82 -> 2.69 ms per config-3 step
  else conv_out_kernel<<<grid, EDGE_TILE, 0, stream>>>(p);
  ECB_LAUNCHED();
  return 0;
}

}  // namespace ecb
