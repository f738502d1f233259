// Language model + arithmetic coder of the entropy-coded .ecdc stream.
//
// Replaces, for compress_to_file / decompress_from_file with use_lm=True (reference encodec/compress.py:63-87,125-152):
//   LMModel.forward (model.py:65-83): sum of per-codebook embeddings -> StreamingTransformerEncoder
//   (modules/transformer.py:62-119: LayerNorm, sinusoidal positions, post-norm nn.TransformerEncoderLayer x num_layers with
//   a causal window of past_context rows) -> one Linear(dim, card) + softmax per codebook;
//   build_stable_quantized_cdf (quantization/ac.py:18-53) and ArithmeticCoder / ArithmeticDecoder (ac.py:56-260).
//
// Design. The reference evaluates the LM one time step at a time in both directions. Here
//   * COMPRESSION knows every code, so all steps of all frames are rows of ONE batched pass (28 launches per call whatever
//     the length); the logits kernel forms softmax -> quantised cdf on chip and emits only the two cdf values the coder
//     needs per symbol (8 bytes instead of a 4 KB distribution); the coder itself is the sequential integer recurrence of
//     ac_core.h on the host.
//   * DECOMPRESSION is sequential by construction (step t + 1 is conditioned on the symbols decoded at step t), so the whole
//     loop stays on the device: the keys / values of earlier steps in a cache, the step's transformer as ONE 8-CTA cluster
//     kernel (one head per CTA, the row exchanged through distributed shared memory behind cluster barriers; other shapes:
//     the batched kernels with one row), and a one-block kernel in which one warp runs ArithmeticDecoder.pull for the K
//     codebooks of the step and writes the codes the next step's embedding reads. One captured step (4 launches) is replayed
//     per latent frame; no host round trip per step (the reference makes K .item() calls per step).
//   * Both directions MUST see bit-identical probabilities or the stream decodes to garbage. Every kernel therefore computes
//     an output element with the same instruction sequence whatever the number of rows in the launch: one warp per
//     (row, output) dot product with a fixed lane partition of K and a fixed shuffle tree, one warp per (row, head)
//     attention over keys in window order, per-row softmax reductions; the arithmetic lives in device functions of explicitly
//     rounded intrinsics shared by all kernels. tests/test_zz_lm_gpu.py asserts the identity.
// The state of a stream is its key / value cache: row 0 of an item holds the projections of the all-zero row the reference
// seeds its state with (transformer.py:103-104; it is attended like a real position until the window drops it), row p + 1
// position p. These are CUDA-core kernels: the work is a few MFLOP per step and latency-bound (decode) or a few GFLOP per
// call (compress); what matters is the launch count and that nothing returns to the host inside the loop.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/encodec_b200.h"
#include "ac_core.h"
#include "common.cuh"

namespace ecb {
namespace {

constexpr int kWarps = 8;        // warps per block of the row kernels
constexpr int kRowTile = 8;      // rows that share one weight fetch in lm_linear
constexpr int kColsPerWarp = 4;  // output columns a warp walks in lm_linear
constexpr int kHeadRows = 1024;  // rows whose logits are materialised at a time ([rows][K * card] floats of workspace)
constexpr int kAcThreads = 512;  // block of the decoder-pull kernel (stages the cdfs; warp 0 decodes)
constexpr int kAcSmem = 160 * 1024;
constexpr int kMaxDimPerLane = 8;  // dim <= 256

struct CdfParams {
  float roundoff;   // ac.py:19 (1e-8 as float32: the reference's pdf is a float32 tensor)
  float scale;      // (1 - alpha) * total_range as float32, alpha = min_range * card / total_range (ac.py:40-44)
  int min_range;
};

struct Tokens {
  const long long* p;   // element (item, k, t) at p[item * item_stride + k * k_stride + t * t_stride]
  long long item_stride, k_stride, t_stride;
  int are_codes;        // 1: p holds the codes of the frame, step t reads 1 + code[t - 1] (0 at t = 0; compress.py:69,78)
                        // 0: p holds the LM's input indices of steps t0 .. t0 + n_t - 1 (LMModel.forward's `indices`)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ long long token_index(const Tokens& tk, long long item, int k, long long t, long long t0, int card) {
  long long v;
  if (tk.are_codes) v = t == 0 ? 0 : 1 + tk.p[item * tk.item_stride + k * tk.k_stride + (t - 1) * tk.t_stride];
  else v = tk.p[item * tk.item_stride + k * tk.k_stride + (t - t0) * tk.t_stride];
  return v < 0 ? 0 : (v > card ? card : v);
}

// ---- arithmetic shared by the batched kernels and the one-row cluster kernel ------------------------------------------
// Every floating-point operation below is an explicitly rounded intrinsic (or an explicit fmaf): the compiler may not
// contract or reassociate them differently in different kernels, so a value computed by lm_step_cluster_kernel and by the
// batched kernels is the same bit pattern -- the property the arithmetic coder needs (see the header of this file).

// LayerNorm of one row held as dim / 32 values per lane (nn.LayerNorm, eps inside the square root, biased variance)
__device__ __forceinline__ void warp_layer_norm(float (&v)[kMaxDimPerLane], int dim, int lane, const float* __restrict__ w,
                                                const float* __restrict__ b, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxDimPerLane; ++i) s = __fadd_rn(s, (lane + 32 * i < dim) ? v[i] : 0.f);
  const float mean = __fdiv_rn(warp_sum(s), (float)dim);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxDimPerLane; ++i) {
    const float d = (lane + 32 * i < dim) ? __fsub_rn(v[i], mean) : 0.f;
    q = fmaf(d, d, q);
  }
  const float rstd = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(__fdiv_rn(warp_sum(q), (float)dim), eps)));
#pragma unroll
  for (int i = 0; i < kMaxDimPerLane; ++i) {
    const int d = lane + 32 * i;
    if (d < dim) v[i] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[i], mean), rstd), w[d]), b[d]);
  }
}

// model.py:79 (sum of embeddings) + transformer.py:106-111 (norm_in, + sinusoidal position embedding) for one row by one warp
__device__ __forceinline__ void embed_row(const Tokens& tk, const float* __restrict__ emb, const float* __restrict__ nw,
                                          const float* __restrict__ nb, const float* __restrict__ pos_div, long long item,
                                          long long t, long long t0, int K, int card, int dim, float eps, int lane,
                                          float (&v)[kMaxDimPerLane]) {
#pragma unroll
  for (int i = 0; i < kMaxDimPerLane; ++i) v[i] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 32) {   // the indices first (one lane per codebook), so that the row loads do not wait on them
    const long long mine = (k0 + lane < K) ? token_index(tk, item, k0 + lane, t, t0, card) : 0;
    const int g = (K - k0) < 32 ? (K - k0) : 32;
    for (int k1 = 0; k1 < g; k1 += 8) {    // eight embedding rows in flight, then added in codebook order (model.py:79)
      float e[8][kMaxDimPerLane];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long idx = __shfl_sync(0xffffffffu, mine, (k1 + u) & 31);
        const float* row_e = emb + ((size_t)(k0 + k1 + u) * (card + 1) + (size_t)idx) * dim;
#pragma unroll
        for (int i = 0; i < kMaxDimPerLane; ++i) {
          const int d = lane + 32 * i;
          e[u][i] = (k1 + u < g && d < dim) ? row_e[d] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (k1 + u < g) {
#pragma unroll
          for (int i = 0; i < kMaxDimPerLane; ++i) v[i] = __fadd_rn(v[i], e[u][i]);
        }
      }
    }
  }
  warp_layer_norm(v, dim, lane, nw, nb, eps);
  const int half = dim >> 1;
#pragma unroll
  for (int i = 0; i < kMaxDimPerLane; ++i) {
    const int d = lane + 32 * i;
    if (d < dim) {
      // create_sin_embedding (transformer.py:16-27): float32 phase = position / max_period^(j / (half - 1)), cat(cos, sin)
      const float phase = __fdiv_rn((float)t, pos_div[d < half ? d : d - half]);
      v[i] = __fadd_rn(v[i], d < half ? cosf(phase) : sinf(phase));
    }
  }
}

// NC weight rows of a warp in registers: k = lane + 32 i (zero beyond K or for columns n < 0)
template <int KPL, int NC>
__device__ __forceinline__ void load_cols(const float* __restrict__ W, const float* __restrict__ b, int K, const int (&n)[NC],
                                          int lane, float (&w)[NC][KPL], float (&bias)[NC]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int k = lane + 32 * i;
      w[c][i] = (n[c] >= 0 && k < K) ? W[(size_t)n[c] * K + k] : 0.f;
    }
    bias[c] = n[c] >= 0 ? b[n[c]] : 0.f;
  }
}
// b[n] + sum_k W[n][k] x[k] for the warp's NC columns: lane-local fmaf chain in i order, xor tree, + bias (all lanes hold it)
template <int KPL, int NC>
__device__ __forceinline__ void dot_cols(const float (&w)[NC][KPL], const float (&bias)[NC], const float* xs, int K, int lane,
                                         float (&out)[NC]) {
  float acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.f;
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const int k = lane + 32 * i;
    if (k < K) {
      const float xv = xs[k];
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] = fmaf(w[c][i], xv, acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) out[c] = __fadd_rn(warp_sum(acc[c]), bias[c]);
}
__device__ __forceinline__ float gelu_exact(float v) {   // F.gelu, erf form
  return __fmul_rn(__fmul_rn(0.5f, v), __fadd_rn(1.f, erff(__fmul_rn(v, 0.70710678118654752440f))));
}

constexpr int kAttnThreads = 512;   // 16 warps per (row, head): at past_context = 262 a warp owns <= 17 keys
// _sa_block (transformer.py:42-59) for one (row, head) by one block of kAttnThreads: keys / values = nk cache rows starting
// at kv (already offset to the head's slice). Warp w owns keys w, w + 16, ...; a key's head slice (hd <= 32 contiguous
// floats) is read by the lanes of the warp (one coalesced request per key), the score is the xor-tree sum of the lane
// products; max and sum are block reductions in warp order; the weighted values are one accumulator per lane (= head
// dimension) and warp, summed across the warps in warp order. Returns the output of dimension tid (valid for tid < hd).
__device__ __forceinline__ float attn_head(float qv, const float* kv, int nk, int dim, int hd, float* sc,
                                           float* red, float (*part)[32], int tid) {
  constexpr int NW = kAttnThreads / 32;
  const int lane = tid & 31, warp = tid >> 5;
  const float scale = __fdiv_rn(1.f, __fsqrt_rn((float)hd));
  float m = -INFINITY;
#pragma unroll 8
  for (int i = warp; i < nk; i += NW) {
    const float kval = lane < hd ? kv[(size_t)i * 2 * dim + lane] : 0.f;
    const float s = __fmul_rn(warp_sum(__fmul_rn(qv, kval)), scale);
    if (lane == 0) sc[i] = s;
    m = fmaxf(m, s);
  }
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w2 = 1; w2 < NW; ++w2) m = fmaxf(m, red[w2]);
  __syncthreads();
  float l = 0.f, acc = 0.f;
#pragma unroll 8
  for (int i = warp; i < nk; i += NW) {
    const float e = expf(__fsub_rn(sc[i], m));
    const float vval = lane < hd ? kv[(size_t)i * 2 * dim + dim + lane] : 0.f;
    l = __fadd_rn(l, e);
    acc = fmaf(e, vval, acc);
  }
  if (lane == 0) red[warp] = l;
  part[warp][lane] = acc;
  __syncthreads();
  float o = 0.f;
  if (tid < hd) {
    float L = red[0], O = part[0][tid];
#pragma unroll
    for (int w2 = 1; w2 < NW; ++w2) {
      L = __fadd_rn(L, red[w2]);
      O = __fadd_rn(O, part[w2][tid]);
    }
    o = __fdiv_rn(O, L);
  }
  __syncthreads();   // sc / red / part may be reused by the caller
  return o;
}

// ---- batched kernels (any number of rows) ----------------------------------------------------------------------------
__global__ void lm_embed_kernel(Tokens tk, const float* __restrict__ emb, const float* __restrict__ nw,
                                const float* __restrict__ nb, const float* __restrict__ pos_div, float* __restrict__ x,
                                long long n_rows, int n_t, long long t0_arg, const long long* __restrict__ t_ptr, int K,
                                int card, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const long long t0 = t_ptr ? *t_ptr : t0_arg;   // the decoding loop replays one captured step: t lives on the device
  float v[kMaxDimPerLane];
  embed_row(tk, emb, nw, nb, pos_div, row / n_t, t0 + row % n_t, t0, K, card, dim, eps, lane, v);
#pragma unroll
  for (int i = 0; i < kMaxDimPerLane; ++i) {
    const int d = lane + 32 * i;
    if (d < dim) x[row * dim + d] = v[i];
  }
}

enum { EPI_QKV = 0, EPI_GELU = 1, EPI_RESID = 2, EPI_PLAIN = 3 };
struct LinArgs {
  const float* x;       // [n_rows][K]
  const float* W;       // [N][K] (nn.Linear layout)
  const float* b;       // [N]
  int K, N;
  long long n_rows;
  float* out;           // GELU / RESID / PLAIN: [n_rows][N]; QKV: q [n_rows][dim]
  const float* resid;   // RESID: out = resid + (W x + b)
  float* cache;         // QKV: this layer's cache, [n_items][capacity + 1][2 * dim]
  long long capacity;
  int n_t;
  long long t0;
  const long long* t_ptr;
  int dim;
  // LayerNorm applied on load (the post-norm layer's norm1 / norm2, transformer.py:38-39, never exist as a pass of their own):
  const float *ln_w, *ln_b;     // non-null: x rows are LayerNorm(x) * ln_w + ln_b (needs K <= 256)
  const float *rln_w, *rln_b;   // non-null (RESID): the residual rows likewise (needs N <= 256)
  float eps;
};

// y[r][n] = epilogue(b[n] + sum_k W[n][k] x[r][k]). A warp owns kColsPerWarp output columns: their weight rows sit in
// registers (all loads of the warp in flight together, issued before the activation tile is awaited), the rows of the tile
// come from shared memory; per (row, column) the fixed sequence of dot_cols -- the same for every row, whatever the launch shape.
template <int KPL, int EPI>
__global__ void __launch_bounds__(kWarps * 32) lm_linear_kernel(LinArgs a) {
  extern __shared__ float xs[];   // [kRowTile][K]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * kRowTile;
  const int nr = (int)((a.n_rows - r0) < kRowTile ? (a.n_rows - r0) : kRowTile);
  float* rs = xs + kRowTile * a.K;   // [kRowTile][N], only with rln_w
  const int n0 = (blockIdx.x * kWarps + warp) * kColsPerWarp;
  int ncol[kColsPerWarp];
#pragma unroll
  for (int c = 0; c < kColsPerWarp; ++c) ncol[c] = (n0 + c < a.N) ? n0 + c : -1;
  float w[kColsPerWarp][KPL], bias[kColsPerWarp];
  load_cols<KPL, kColsPerWarp>(a.W, a.b, a.K, ncol, lane, w, bias);
  for (int i = threadIdx.x; i < nr * a.K; i += blockDim.x) xs[i] = a.x[r0 * a.K + i];
  if (EPI == EPI_RESID && a.rln_w)
    for (int i = threadIdx.x; i < nr * a.N; i += blockDim.x) rs[i] = a.resid[r0 * a.N + i];
  __syncthreads();
  if (a.ln_w || (EPI == EPI_RESID && a.rln_w)) {   // kRowTile == kWarps: warp r normalises row r, the same function for every launch shape
    if (warp < nr) {
      float v[kMaxDimPerLane];
      if (a.ln_w) {
#pragma unroll
        for (int i = 0; i < kMaxDimPerLane; ++i) {
          const int d = lane + 32 * i;
          v[i] = d < a.K ? xs[warp * a.K + d] : 0.f;
        }
        warp_layer_norm(v, a.K, lane, a.ln_w, a.ln_b, a.eps);
#pragma unroll
        for (int i = 0; i < kMaxDimPerLane; ++i) {
          const int d = lane + 32 * i;
          if (d < a.K) xs[warp * a.K + d] = v[i];
        }
      }
      if (EPI == EPI_RESID && a.rln_w) {
#pragma unroll
        for (int i = 0; i < kMaxDimPerLane; ++i) {
          const int d = lane + 32 * i;
          v[i] = d < a.N ? rs[warp * a.N + d] : 0.f;
        }
        warp_layer_norm(v, a.N, lane, a.rln_w, a.rln_b, a.eps);
#pragma unroll
        for (int i = 0; i < kMaxDimPerLane; ++i) {
          const int d = lane + 32 * i;
          if (d < a.N) rs[warp * a.N + d] = v[i];
        }
      }
    }
    __syncthreads();
  }
  if (n0 >= a.N) return;   // warp-uniform, no block-wide barrier below
  const long long t0 = (EPI == EPI_QKV && a.t_ptr) ? *a.t_ptr : a.t0;
  for (int r = 0; r < nr; ++r) {
    float res[kColsPerWarp];
    dot_cols<KPL, kColsPerWarp>(w, bias, xs + r * a.K, a.K, lane, res);
    float v = 0.f;   // lane c finishes column n0 + c
#pragma unroll
    for (int c = 0; c < kColsPerWarp; ++c)
      if (lane == c) v = res[c];
    const int n = n0 + lane;
    if (lane < kColsPerWarp && n < a.N) {
      const long long row = r0 + r;
      if (EPI == EPI_GELU) {
        a.out[row * a.N + n] = gelu_exact(v);
      } else if (EPI == EPI_RESID) {
        a.out[row * a.N + n] = __fadd_rn(a.rln_w ? rs[r * a.N + n] : a.resid[row * a.N + n], v);
      } else if (EPI == EPI_PLAIN) {
        a.out[row * a.N + n] = v;
      } else {
        if (n < a.dim) {
          a.out[row * a.dim + n] = v;
        } else {
          const long long item = row / a.n_t, t = t0 + row % a.n_t;
          float* base = a.cache + (size_t)item * (a.capacity + 1) * 2 * a.dim;
          base[(size_t)(t + 1) * 2 * a.dim + (n - a.dim)] = v;
          if (t == 0) base[n - a.dim] = a.b[n];   // projections of the reference's all-zero seed row: W 0 + b
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kAttnThreads) lm_attn_kernel(const float* __restrict__ q, const float* __restrict__ cache,
                                                                float* __restrict__ out, int n_t, long long t0_arg,
                                                                const long long* __restrict__ t_ptr, long long capacity, int dim,
                                                                int heads, int past_context) {
  extern __shared__ float sc[];   // [past_context + 1] scores
  __shared__ float red[kAttnThreads / 32];
  __shared__ float part[kAttnThreads / 32][32];
  const int tid = threadIdx.x, lane = tid & 31;
  const long long w = blockIdx.x;
  const long long t0 = t_ptr ? *t_ptr : t0_arg;
  const long long row = w / heads;
  const int h = (int)(w % heads);
  const long long item = row / n_t, t = t0 + row % n_t;
  const int hd = dim / heads;
  // the rows the streaming state still holds (transformer.py:116-117): cache rows [lo, t + 1], lo = t + 1 - min(t + 1, past_context)
  const long long n_past = (t + 1) < past_context ? (t + 1) : past_context;
  const long long lo = t + 1 - n_past;
  const int nk = (int)n_past + 1;
  const float* kv = cache + ((size_t)item * (capacity + 1) + (size_t)lo) * 2 * dim + h * hd;
  const float qv = lane < hd ? q[row * dim + h * hd + lane] : 0.f;
  const float o = attn_head(qv, kv, nk, dim, hd, sc, red, part, tid);
  if (tid < hd) out[row * dim + h * hd + tid] = o;
}

// ---- the decoding step's transformer as ONE cluster kernel (one row) --------------------------------------------------
// Eight CTAs = eight heads. CTA h computes the q / k / v columns of ITS head (so attention needs no exchange), then its eighth
// of out_proj, linear1 and linear2; after each of the four phases of a layer the finished values are written into every
// CTA's copy of the row (distributed shared memory) behind one hardware cluster barrier. 20 barriers replace the 25 launches
// of lm_trunk on a single row; the arithmetic is the shared functions above, bit for bit what the batched kernels compute.
constexpr int kClusterCtas = 8;
constexpr int kMaxLayers = 8;
struct ClusterArgs {
  Tokens tk;
  const float *emb, *nin_w, *nin_b, *pos_div;
  const float* lw[kMaxLayers][12];   // in_w, in_b, out_w, out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b
  int n_layers;
  float* cache;                      // [layers][1][capacity + 1][2 dim]
  long long capacity;
  const long long* t_ptr;
  long long t0;
  int K, card, dim, heads, hidden, past_context;
  float eps;
  float* x_out;                      // [dim]: raw output of the last layer (norm2 pending, applied by the heads on load)
};

template <int KPL, int NC, class MapFn, class EpiFn>
__device__ __forceinline__ void cluster_cols(const float* __restrict__ W, const float* __restrict__ b, int K, int n_local,
                                             MapFn map, const float* xs, int lane, int warp, int n_warps, EpiFn epi) {
  for (int g = warp; g * NC < n_local; g += n_warps) {
    int ncol[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) ncol[c] = (g * NC + c < n_local) ? map(g * NC + c) : -1;
    float w[NC][KPL], bias[NC], res[NC];
    load_cols<KPL, NC>(W, b, K, ncol, lane, w, bias);
    dot_cols<KPL, NC>(w, bias, xs, K, lane, res);
    float v = 0.f;
    int n = -1;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (lane == c) {
        v = res[c];
        n = ncol[c];
      }
    if (lane < NC && n >= 0) epi(n, v);
  }
}

template <int KPL_D, int KPL_H>
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kAttnThreads) lm_step_cluster_kernel(ClusterArgs a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sc[];                  // [past_context + 1] attention scores
  __shared__ float X[256], XN[256], ATT[256], Y[256], YN[256], HID[1024], QL[32];
  __shared__ float red[kAttnThreads / 32];
  __shared__ float part[kAttnThreads / 32][32];
  constexpr int NWARPS = kAttnThreads / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = (int)cluster.block_rank();       // this CTA's head
  const int dim = a.dim, hd = dim / a.heads, hidden = a.hidden;
  const long long t = a.t_ptr ? *a.t_ptr : a.t0;
  const size_t layer_cache = (size_t)(a.capacity + 1) * 2 * dim;
  if (warp == 0) {   // every CTA forms the input row itself (one warp per row, as lm_embed_kernel)
    float v[kMaxDimPerLane];
    embed_row(a.tk, a.emb, a.nin_w, a.nin_b, a.pos_div, 0, t, t, a.K, a.card, dim, a.eps, lane, v);
#pragma unroll
    for (int i = 0; i < kMaxDimPerLane; ++i) {
      const int d = lane + 32 * i;
      if (d < dim) X[d] = v[i];
    }
  }
  cluster.sync();    // also: no CTA writes into a peer before that peer has started
  const long long n_past = (t + 1) < a.past_context ? (t + 1) : a.past_context;
  const long long lo = t + 1 - n_past;
  const int nk = (int)n_past + 1;
  for (int l = 0; l < a.n_layers; ++l) {
    const float* const* lw = a.lw[l];
    float* lc = a.cache + (size_t)l * layer_cache;
    if (warp == 0) {   // XN = the layer's input: the embedding (layer 0) or norm2 of the previous layer applied to its raw output
      float v[kMaxDimPerLane];
#pragma unroll
      for (int i = 0; i < kMaxDimPerLane; ++i) {
        const int d = lane + 32 * i;
        v[i] = d < dim ? X[d] : 0.f;
      }
      if (l) warp_layer_norm(v, dim, lane, a.lw[l - 1][10], a.lw[l - 1][11], a.eps);
#pragma unroll
      for (int i = 0; i < kMaxDimPerLane; ++i) {
        const int d = lane + 32 * i;
        if (d < dim) XN[d] = v[i];
      }
    }
    __syncthreads();
    // q / k / v of head h: local column j -> in_proj row (j / hd) * dim + h * hd + j % hd
    cluster_cols<KPL_D, 5>(lw[0], lw[1], dim, 3 * hd, [&](int j) { return (j / hd) * dim + h * hd + j % hd; }, XN, lane, warp, NWARPS,
                           [&](int n, float v) {
                             if (n < dim) {
                               QL[n - h * hd] = v;
                             } else {
                               lc[(size_t)(t + 1) * 2 * dim + (n - dim)] = v;
                               if (t == 0) lc[n - dim] = lw[1][n];
                             }
                           });
    __syncthreads();   // QL and this block's own cache writes are visible to the block
    {
      const float qv = lane < hd ? QL[lane] : 0.f;
      const float o = attn_head(qv, lc + (size_t)lo * 2 * dim + h * hd, nk, dim, hd, sc, red, part, tid);
      if (tid < hd) {
#pragma unroll
        for (int r = 0; r < kClusterCtas; ++r) cluster.map_shared_rank(ATT, r)[h * hd + tid] = o;
      }
    }
    cluster.sync();
    // out_proj + residual: this CTA's dim / 8 columns
    cluster_cols<KPL_D, 2>(lw[2], lw[3], dim, dim / kClusterCtas, [&](int j) { return h * (dim / kClusterCtas) + j; }, ATT, lane, warp,
                           NWARPS, [&](int n, float v) {
                             const float y = __fadd_rn(XN[n], v);
#pragma unroll
                             for (int r = 0; r < kClusterCtas; ++r) cluster.map_shared_rank(Y, r)[n] = y;
                           });
    cluster.sync();
    if (warp == 0) {   // YN = norm1(y)
      float v[kMaxDimPerLane];
#pragma unroll
      for (int i = 0; i < kMaxDimPerLane; ++i) {
        const int d = lane + 32 * i;
        v[i] = d < dim ? Y[d] : 0.f;
      }
      warp_layer_norm(v, dim, lane, lw[8], lw[9], a.eps);
#pragma unroll
      for (int i = 0; i < kMaxDimPerLane; ++i) {
        const int d = lane + 32 * i;
        if (d < dim) YN[d] = v[i];
      }
    }
    __syncthreads();
    cluster_cols<KPL_D, 7>(lw[4], lw[5], dim, hidden / kClusterCtas, [&](int j) { return h * (hidden / kClusterCtas) + j; }, YN, lane,
                           warp, NWARPS, [&](int n, float v) {
                             const float g = gelu_exact(v);
#pragma unroll
                             for (int r = 0; r < kClusterCtas; ++r) cluster.map_shared_rank(HID, r)[n] = g;
                           });
    cluster.sync();
    cluster_cols<KPL_H, 2>(lw[6], lw[7], hidden, dim / kClusterCtas, [&](int j) { return h * (dim / kClusterCtas) + j; }, HID, lane, warp,
                           NWARPS, [&](int n, float v) {
                             const float y2 = __fadd_rn(YN[n], v);
#pragma unroll
                             for (int r = 0; r < kClusterCtas; ++r) cluster.map_shared_rank(X, r)[n] = y2;
                           });
    cluster.sync();
  }
  if (h == 0 && tid < dim) a.x_out[tid] = X[tid];
}

// pdf -> range width of build_stable_quantized_cdf (ac.py:36-45) in the float32 arithmetic of the torch CPU kernels
__device__ __forceinline__ int pdf_to_range(float p, const CdfParams& cp) {
  const float fl = __fmul_rn(floorf(__fdiv_rn(p, cp.roundoff)), cp.roundoff);
  return (int)floorf(__fmul_rn(cp.scale, fl)) + cp.min_range;
}

// Inclusive scan of the `card` range widths in vals[] (one contiguous chunk per thread) -> cdf in vals[] (ac.py:46).
__device__ __forceinline__ void block_cumsum(int* vals, int card, int* wsum) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int per = (card + nthr - 1) / nthr;
  const int c0 = min(tid * per, card), c1 = min(c0 + per, card);
  int local = 0;
  for (int n = c0; n < c1; ++n) local += vals[n];
  int inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  int run = inc - local;
  for (int w2 = 0; w2 < warp; ++w2) run += wsum[w2];
  for (int n = c0; n < c1; ++n) {
    run += vals[n];
    vals[n] = run;
  }
  __syncthreads();
}

// softmax over the codebook (model.py:83) + ac.py:18-53 for one (row, codebook) per block, from the logits the per-codebook
// Linear layers (model.py:81-82; one lm_linear launch over the concatenated [K * card][dim] weights) left in `logits`.
// Outputs (each optional): probabilities, the quantised cdf, and for compression the coder's pair (cdf[s - 1], cdf[s]) of
// the symbol s actually coded at (item, k, t).
__global__ void __launch_bounds__(kWarps * 32) lm_softmax_cdf_kernel(const float* __restrict__ logits, long long row0, int card,
                                                                      int K, Tokens tk, int n_t, long long t0_arg,
                                                                      const long long* __restrict__ t_ptr, CdfParams cp,
                                                                      float* __restrict__ probas, int* __restrict__ cdf,
                                                                      int* __restrict__ sym_ranges) {
  extern __shared__ float row[];   // [card]
  __shared__ float red[kWarps];
  __shared__ int wsum[kWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.y;
  const long long grow = row0 + blockIdx.x;
  const float* src = logits + ((size_t)blockIdx.x * K + k) * card;
  float m = -INFINITY;
  for (int n = tid; n < card; n += blockDim.x) {
    const float v = src[n];
    row[n] = v;
    m = fmaxf(m, v);
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w2 = 1; w2 < kWarps; ++w2) m = fmaxf(m, red[w2]);
  __syncthreads();
  float s = 0.f;
  for (int n = tid; n < card; n += blockDim.x) {
    const float e = expf(row[n] - m);
    row[n] = e;
    s += e;
  }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float total = red[0];
#pragma unroll
  for (int w2 = 1; w2 < kWarps; ++w2) total += red[w2];
  int* irow = reinterpret_cast<int*>(row);
  for (int n = tid; n < card; n += blockDim.x) {
    const float p = __fdiv_rn(row[n], total);
    if (probas) probas[((size_t)grow * K + k) * card + n] = p;
    irow[n] = pdf_to_range(p, cp);
  }
  __syncthreads();
  block_cumsum(irow, card, wsum);
  if (cdf)
    for (int n = tid; n < card; n += blockDim.x) cdf[((size_t)grow * K + k) * card + n] = irow[n];
  if (sym_ranges && tid == 0) {
    const long long t0 = t_ptr ? *t_ptr : t0_arg;
    const long long item = grow / n_t, t = t0 + grow % n_t;
    long long sv = tk.p[item * tk.item_stride + k * tk.k_stride + t * tk.t_stride];
    sv = sv < 0 ? 0 : (sv >= card ? card - 1 : sv);
    sym_ranges[((size_t)grow * K + k) * 2 + 0] = sv ? irow[sv - 1] : 0;
    sym_ranges[((size_t)grow * K + k) * 2 + 1] = irow[sv];
  }
}

// build_stable_quantized_cdf alone: pdf [n_rows][card] float32 -> cdf [n_rows][card] int32. One block per row.
__global__ void __launch_bounds__(kWarps * 32) lm_cdf_kernel(const float* __restrict__ pdf, int* __restrict__ cdf, int card,
                                                              CdfParams cp) {
  extern __shared__ int iv[];
  __shared__ int wsum[kWarps];
  const size_t row = blockIdx.x;
  for (int n = threadIdx.x; n < card; n += blockDim.x) iv[n] = pdf_to_range(pdf[row * card + n], cp);
  __syncthreads();
  block_cumsum(iv, card, wsum);
  for (int n = threadIdx.x; n < card; n += blockDim.x) cdf[row * card + n] = iv[n];
}

// ArithmeticDecoder.pull (ac.py:214-260) by one warp: every lane carries the same decoder state; the reference's binary search
// for the symbol whose range holds `current` becomes a 32-ary search (lane j tests candidate base + j * step with one
// rounded double product each, a ballot picks the last candidate whose lower bound is <= current). The ranges of
// consecutive symbols are disjoint and ordered, so the last symbol whose lower bound is <= current is the only one that can
// contain it: the symbol, the new (low, high) and the failure condition are the reference's (ac_core.h is the scalar form
// the host runs; tests decode the reference's own streams through both).
__device__ __forceinline__ int ac_pull_warp(ac::Decoder& d, const unsigned char* __restrict__ data, long long n_bits,
                                            const int* cdf, int card, int bits, int lane) {
  if (d.status != ac::AC_OK) return -1;
  if (!ac::refill(d, data, n_bits, bits)) return -1;
  const ac::Scale ratio = ac::make_scale(d.high - d.low + 1, bits);
  int base = 0, span = card;
  while (span > 1) {
    const int step = (span + 31) >> 5;
    const int s = base + lane * step;
    bool ok = false;
    if (s < base + span) {
      const int64_t range_low = s > 0 ? cdf[s - 1] : 0;
      ok = d.current >= ac::eff_low(range_low, ratio) + d.low;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (mask == 0) {   // cannot happen: the first candidate's lower bound is <= current by induction
      d.status = ac::AC_SEARCH_FAILED;
      return -1;
    }
    const int last = 31 - __clz(mask);
    const int nb = base + last * step;
    span = min(step, base + span - nb);
    base = nb;
  }
  const int64_t range_low = base > 0 ? cdf[base - 1] : 0;
  const int64_t range_high = (int64_t)cdf[base] - 1;
  if (range_high < range_low) {
    d.status = ac::AC_BAD_CDF;
    return -1;
  }
  const uint64_t low = ac::eff_low(range_low, ratio) + d.low;
  const uint64_t high = ac::eff_high(range_high, ratio) + d.low;
  if (d.current < low || d.current > high) {   // "Binary search failed" (ac.py:236): current fell between two ranges
    d.status = ac::AC_SEARCH_FAILED;
    return -1;
  }
  d.low = low;
  d.high = high;
  ac::flush_prefix(d);
  return base;
}

// The K symbols of step t (compress.py:137-148) -- or any K consecutive symbols with their cdf rows: the cdfs are staged in
// shared memory `group` rows at a time by the whole block, warp 0 decodes them in order. The stream, the decoder state and
// the step counter stay on the device; codes[k * k_stride + t] is what the next step's embedding kernel reads.
__global__ void lm_ac_init_kernel(ac::Decoder* st, long long first_bit, long long* t_ptr) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    ac::decoder_init(*st, first_bit);
    if (t_ptr) *t_ptr = 0;
  }
}
__global__ void __launch_bounds__(kAcThreads) lm_ac_pull_kernel(ac::Decoder* st, const unsigned char* __restrict__ data, long long n_bits,
                                                           const int* __restrict__ cdf, int K, int card, int bits, int group,
                                                           long long* __restrict__ codes, long long k_stride, long long t_arg,
                                                           long long* t_ptr) {
  extern __shared__ __align__(16) int scdf[];   // [group][card]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long t = t_ptr ? *t_ptr : t_arg;
  ac::Decoder d;
  if (warp == 0) d = *st;
  for (int k0 = 0; k0 < K; k0 += group) {
    const int g = min(group, K - k0);
    const int total = g * card;
    const int* src = cdf + (size_t)k0 * card;
    if (((size_t)src & 15) == 0 && (total & 3) == 0) {
      const int4* s4 = reinterpret_cast<const int4*>(src);
      int4* d4 = reinterpret_cast<int4*>(scdf);
#pragma unroll 8
      for (int i = tid; i < (total >> 2); i += kAcThreads) d4[i] = s4[i];
    } else {
#pragma unroll 8
      for (int i = tid; i < total; i += kAcThreads) scdf[i] = src[i];
    }
    __syncthreads();
    if (warp == 0) {
      for (int k = 0; k < g; ++k) {
        const int s = ac_pull_warp(d, data, n_bits, scdf + k * card, card, bits, lane);
        if (lane == 0) codes[(k0 + k) * k_stride + t] = s < 0 ? 0 : s;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    *st = d;
    if (t_ptr) *t_ptr = t + 1;
  }
}
__global__ void lm_ac_result_kernel(const ac::Decoder* st, long long* result) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    result[0] = st->status;
    result[1] = ac::bytes_consumed(*st);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
struct LmSpec {
  int n_q, card, dim, n_layers, n_heads, hidden, past_context;
  float max_period;
};

struct Lm {
  LmSpec spec;
  int device = -1;
  float* weights = nullptr;   // one allocation, slices below
  size_t n_weights = 0;
  std::map<std::string, std::pair<size_t, size_t>> slots;   // key -> (offset, numel)
  std::map<std::string, bool> loaded;
  bool finalized = false;
  const float* w(const std::string& key) const { return weights + slots.at(key).first; }
  size_t emb_off = 0, lin_w_off = 0, lin_b_off = 0, pos_off = 0;
  struct LayerW {
    const float *in_w, *in_b, *out_w, *out_b, *l1_w, *l1_b, *l2_w, *l2_b, *n1_w, *n1_b, *n2_w, *n2_b;
  };
  std::vector<LayerW> layers;   // resolved at finalize: no map lookups inside the decoding loop
  const float *nin_w = nullptr, *nin_b = nullptr;
};

void add_slot(Lm& lm, const std::string& key, size_t numel, size_t& cursor) {
  lm.slots[key] = {cursor, numel};
  lm.loaded[key] = false;
  cursor += (numel + 3) & ~(size_t)3;
}

template <int EPI>
int launch_linear(const LinArgs& a, cudaStream_t s) {
  const int kpl = (a.K + 31) / 32;
  dim3 grid((unsigned)cdiv((long long)a.N, (long long)kWarps * kColsPerWarp), (unsigned)cdiv(a.n_rows, (long long)kRowTile));
  static_assert(kRowTile == kWarps, "LayerNorm on load: one warp per row of the tile");
  const size_t smem = (size_t)kRowTile * (a.K + (a.rln_w ? a.N : 0)) * sizeof(float);
  ECB_REQUIRE(kpl <= 32 && smem <= 48 * 1024, "lm: linear layer with K = %d is not supported (K <= 1024)", a.K);
  ECB_REQUIRE((!a.ln_w || a.K <= 32 * kMaxDimPerLane) && (!a.rln_w || a.N <= 32 * kMaxDimPerLane), "lm: LayerNorm on load needs <= 256 channels");
  ProfScope prof(PROF_LM_LINEAR, s, 2.0 * a.n_rows * a.N * a.K, 4.0 * ((double)a.N * a.K + (double)a.n_rows * (a.K + a.N)));
  if (kpl <= 2) lm_linear_kernel<2, EPI><<<grid, kWarps * 32, smem, s>>>(a);
  else if (kpl <= 4) lm_linear_kernel<4, EPI><<<grid, kWarps * 32, smem, s>>>(a);
  else if (kpl <= 7) lm_linear_kernel<7, EPI><<<grid, kWarps * 32, smem, s>>>(a);
  else if (kpl <= 8) lm_linear_kernel<8, EPI><<<grid, kWarps * 32, smem, s>>>(a);
  else if (kpl <= 16) lm_linear_kernel<16, EPI><<<grid, kWarps * 32, smem, s>>>(a);
  else if (kpl <= 25) lm_linear_kernel<25, EPI><<<grid, kWarps * 32, smem, s>>>(a);
  else lm_linear_kernel<32, EPI><<<grid, kWarps * 32, smem, s>>>(a);
  ECB_LAUNCHED();
  return 0;
}

CdfParams cdf_params(int card, int bits) {
  // ac.py:39-44: total_range = 2^bits, alpha = min_range * card / total_range, scale = (1 - alpha) * total_range; the
  // Python double becomes a float32 when it multiplies the float32 pdf
  const double total = (double)(1ull << bits);
  const double alpha = 2.0 * card / total;
  CdfParams cp;
  cp.roundoff = (float)1e-8;
  cp.scale = (float)((1.0 - alpha) * total);
  cp.min_range = 2;
  return cp;
}

struct Workspace {
  float *x, *y, *q, *att, *hid, *logits;
  int* cdf;              // decode: [K][card]
  ac::Decoder* dec;
  long long* step;       // decode: the step counter the captured step reads and the pull kernel advances
};

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
// rows whose logits are materialised at a time; ECB_LM_HEAD_ROWS (diagnostic, tests) lowers it to exercise the row chunks
long long head_rows_limit() {
  const char* e = getenv("ECB_LM_HEAD_ROWS");
  const long long v = e ? atoll(e) : 0;
  return v > 0 ? v : (long long)kHeadRows;
}

size_t plan_workspace(const LmSpec& sp, long long n_rows, long long K, Workspace* ws, char* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  const long long head_rows = n_rows < head_rows_limit() ? n_rows : head_rows_limit();
  float* x = (float*)take((size_t)n_rows * sp.dim * 4);
  float* y = (float*)take((size_t)n_rows * sp.dim * 4);
  float* q = (float*)take((size_t)n_rows * sp.dim * 4);
  float* att = (float*)take((size_t)n_rows * sp.dim * 4);
  float* hid = (float*)take((size_t)n_rows * sp.hidden * 4);
  float* logits = (float*)take((size_t)head_rows * K * sp.card * 4);
  int* cdf = (int*)take((size_t)K * sp.card * 4);
  ac::Decoder* dec = (ac::Decoder*)take(sizeof(ac::Decoder));
  long long* step = (long long*)take(sizeof(long long));
  if (ws) *ws = Workspace{x, y, q, att, hid, logits, cdf, dec, step};
  return off;
}

// One pass of the LM over rows (item, t0 .. t0 + n_t - 1): everything up to the transformer output in ws.x.
// t_ptr != null: t0 is read from the device (the captured decoding step).
int lm_trunk(const Lm& lm, const Tokens& tk, long long n_items, int K, long long t0, const long long* t_ptr, long long n_t,
             float* cache, long long capacity, const Workspace& ws, cudaStream_t s) {
  const LmSpec& sp = lm.spec;
  const long long n_rows = n_items * n_t;
  const float eps = 1e-5f;
  const unsigned row_blocks = (unsigned)cdiv(n_rows, (long long)kWarps);
  {
    ProfScope prof(PROF_LM_MISC, s, 0.0, 0.0);
    lm_embed_kernel<<<row_blocks, kWarps * 32, 0, s>>>(tk, lm.weights + lm.emb_off, lm.nin_w, lm.nin_b, lm.weights + lm.pos_off,
                                                      ws.x, n_rows, (int)n_t, t0, t_ptr, K, sp.card, sp.dim, eps);
    ECB_LAUNCHED();
  }
  // Post-norm layer (transformer.py:34-40) with both LayerNorms applied by their readers (lm_linear's on-load form):
  //   y  = x_in + out_proj(attn(x_in))          x_in = the embedding (layer 0) or LayerNorm2 of the previous layer's y2
  //   y2 = LayerNorm1(y) + linear2(gelu(linear1(LayerNorm1(y))))
  // ws.x holds x_in of layer 0, afterwards the raw y2 of the last finished layer; ws.y the raw y. The heads read
  // LayerNorm2(y2) of the last layer the same way.
  const size_t layer_cache = (size_t)n_items * (capacity + 1) * 2 * sp.dim;
  for (int l = 0; l < sp.n_layers; ++l) {
    const Lm::LayerW& lw = lm.layers[l];
    const float* pn_w = l ? lm.layers[l - 1].n2_w : nullptr;   // norm2 of the previous layer, pending on ws.x
    const float* pn_b = l ? lm.layers[l - 1].n2_b : nullptr;
    float* lc = cache + (size_t)l * layer_cache;
    LinArgs a{};
    a.x = ws.x; a.W = lw.in_w; a.b = lw.in_b; a.ln_w = pn_w; a.ln_b = pn_b; a.eps = eps;
    a.K = sp.dim; a.N = 3 * sp.dim; a.n_rows = n_rows; a.out = ws.q; a.cache = lc; a.capacity = capacity;
    a.n_t = (int)n_t; a.t0 = t0; a.t_ptr = t_ptr; a.dim = sp.dim;
    if (launch_linear<EPI_QKV>(a, s)) return 1;
    {
      ProfScope prof_attn(PROF_LM_ATTN, s, 0.0, 0.0);
      lm_attn_kernel<<<(unsigned)(n_rows * sp.n_heads), kAttnThreads, (size_t)(sp.past_context + 1) * sizeof(float), s>>>(
          ws.q, lc, ws.att, (int)n_t, t0, t_ptr, capacity, sp.dim, sp.n_heads, sp.past_context);
      ECB_LAUNCHED();
    }
    LinArgs o{};
    o.x = ws.att; o.W = lw.out_w; o.b = lw.out_b; o.eps = eps;
    o.K = sp.dim; o.N = sp.dim; o.n_rows = n_rows; o.out = ws.y; o.resid = ws.x; o.rln_w = pn_w; o.rln_b = pn_b;
    if (launch_linear<EPI_RESID>(o, s)) return 1;
    LinArgs f1{};
    f1.x = ws.y; f1.W = lw.l1_w; f1.b = lw.l1_b; f1.ln_w = lw.n1_w; f1.ln_b = lw.n1_b; f1.eps = eps;
    f1.K = sp.dim; f1.N = sp.hidden; f1.n_rows = n_rows; f1.out = ws.hid;
    if (launch_linear<EPI_GELU>(f1, s)) return 1;
    LinArgs f2{};
    f2.x = ws.hid; f2.W = lw.l2_w; f2.b = lw.l2_b; f2.eps = eps;
    f2.K = sp.hidden; f2.N = sp.dim; f2.n_rows = n_rows; f2.out = ws.x; f2.resid = ws.y; f2.rln_w = lw.n1_w; f2.rln_b = lw.n1_b;
    if (launch_linear<EPI_RESID>(f2, s)) return 1;
  }
  return 0;
}

// The decoding step's transformer (one stream, one row) as one cluster launch where the shape allows it (the reference's LM:
// 8 heads, dim 200, hidden 800); ECB_LM_CLUSTER=0 keeps the per-phase launches of lm_trunk.
bool cluster_step_supported(const LmSpec& sp) {
  const char* e = getenv("ECB_LM_CLUSTER");
  if (e && atoi(e) == 0) return false;
  return sp.n_heads == kClusterCtas && sp.dim % kClusterCtas == 0 && sp.hidden % kClusterCtas == 0 && sp.dim <= 32 * 7 &&
         sp.hidden <= 32 * 25 && sp.n_layers <= kMaxLayers;
}
int lm_trunk_step(const Lm& lm, const Tokens& tk, int K, long long t0, const long long* t_ptr, float* cache, long long capacity,
                  const Workspace& ws, cudaStream_t s) {
  const LmSpec& sp = lm.spec;
  if (!cluster_step_supported(sp)) return lm_trunk(lm, tk, 1, K, t0, t_ptr, 1, cache, capacity, ws, s);
  ClusterArgs a{};
  a.tk = tk;
  a.emb = lm.weights + lm.emb_off; a.nin_w = lm.nin_w; a.nin_b = lm.nin_b; a.pos_div = lm.weights + lm.pos_off;
  for (int l = 0; l < sp.n_layers; ++l) {
    const Lm::LayerW& w = lm.layers[l];
    const float* p[12] = {w.in_w, w.in_b, w.out_w, w.out_b, w.l1_w, w.l1_b, w.l2_w, w.l2_b, w.n1_w, w.n1_b, w.n2_w, w.n2_b};
    for (int i = 0; i < 12; ++i) a.lw[l][i] = p[i];
  }
  a.n_layers = sp.n_layers; a.cache = cache; a.capacity = capacity; a.t_ptr = t_ptr; a.t0 = t0;
  a.K = K; a.card = sp.card; a.dim = sp.dim; a.heads = sp.n_heads; a.hidden = sp.hidden; a.past_context = sp.past_context;
  a.eps = 1e-5f; a.x_out = ws.x;
  ProfScope prof(PROF_LM_LINEAR, s, 0.0, 0.0);
  lm_step_cluster_kernel<7, 25><<<kClusterCtas, kAttnThreads, (size_t)(sp.past_context + 1) * sizeof(float), s>>>(a);
  ECB_LAUNCHED();
  return 0;
}

// model.py:81-83 + ac.py:18-53 on the transformer output in ws.x: the K Linear(dim, card) layers as ONE lm_linear launch over
// their concatenated weights (N = K * card columns: 1024 blocks at K = 32 also when there is a single row), then softmax ->
// quantised cdf per (row, codebook); kHeadRows rows at a time so that the logits stay a bounded scratch buffer.
int lm_heads(const Lm& lm, const Tokens& tk, long long n_items, int K, long long t0, const long long* t_ptr, long long n_t,
             const Workspace& ws, float* probas, int* cdf, int* sym_ranges, cudaStream_t s) {
  const LmSpec& sp = lm.spec;
  const long long n_rows = n_items * n_t;
  const CdfParams cp = cdf_params(sp.card, 24);
  const long long chunk = head_rows_limit();
  for (long long r0 = 0; r0 < n_rows; r0 += chunk) {
    const long long nr = (n_rows - r0) < chunk ? (n_rows - r0) : chunk;
    LinArgs a{};
    a.x = ws.x + (size_t)r0 * sp.dim; a.W = lm.weights + lm.lin_w_off; a.b = lm.weights + lm.lin_b_off;
    a.ln_w = lm.layers.back().n2_w; a.ln_b = lm.layers.back().n2_b; a.eps = 1e-5f;   // norm2 of the last layer, pending on ws.x
    a.K = sp.dim; a.N = K * sp.card; a.n_rows = nr; a.out = ws.logits;
    if (launch_linear<EPI_PLAIN>(a, s)) return 1;
    dim3 grid((unsigned)nr, (unsigned)K);
    {
      ProfScope prof(PROF_LM_MISC, s, 0.0, 0.0);
      lm_softmax_cdf_kernel<<<grid, kWarps * 32, (size_t)sp.card * sizeof(float), s>>>(ws.logits, r0, sp.card, K, tk, (int)n_t, t0,
                                                                                    t_ptr, cp, probas, cdf, sym_ranges);
      ECB_LAUNCHED();
    }
  }
  return 0;
}

int launch_ac_pull(ac::Decoder* st, const unsigned char* data, long long n_bytes, const int* cdf, long long K, int card, int bits,
                   long long* codes, long long k_stride, long long t, long long* t_ptr, cudaStream_t s) {
  int group = kAcSmem / (card * (int)sizeof(int));
  ECB_REQUIRE(group >= 1, "ac: card = %d does not fit the decoder kernel's shared memory", card);
  if (group > K) group = (int)K;
  ECB_CUDA(cudaFuncSetAttribute(lm_ac_pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAcSmem));   // per device, cheap
  ProfScope prof(PROF_AC_PULL, s, 0.0, 4.0 * (double)K * card);
  lm_ac_pull_kernel<<<1, kAcThreads, (size_t)group * card * sizeof(int), s>>>(st, data, n_bytes * 8, cdf, (int)K, card, bits, group,
                                                                            codes, k_stride, t, t_ptr);
  ECB_LAUNCHED();
  return 0;
}

// One captured decoding step, replayed n_steps times (ecb_lm_decode_frame); at most one is kept per LM handle.
struct StepGraphKey {
  const void *data, *codes, *cache, *ws;
  long long n_bytes, K, n_steps, capacity;
  int cluster;
  bool operator==(const StepGraphKey& o) const {
    return data == o.data && codes == o.codes && cache == o.cache && ws == o.ws && n_bytes == o.n_bytes && K == o.K &&
           n_steps == o.n_steps && capacity == o.capacity && cluster == o.cluster;
  }
};
struct StepGraph {
  StepGraphKey key{};
  cudaGraphExec_t exec = nullptr;
};
std::map<const Lm*, StepGraph> g_step_graphs;
cudaStream_t g_lm_capture_stream = nullptr;

int check_call(const Lm* lm, long long n_items, long long K, long long t0, long long n_t, long long capacity) {
  ECB_REQUIRE(lm && lm->finalized, "lm: handle is null or not finalized");
  ECB_REQUIRE(n_items >= 1 && n_t >= 1 && t0 >= 0, "lm: bad n_items=%lld t0=%lld n_t=%lld", n_items, t0, n_t);
  ECB_REQUIRE(K >= 1 && K <= lm->spec.n_q, "lm: %lld codebooks, the model has %d", K, lm->spec.n_q);
  ECB_REQUIRE(t0 + n_t <= capacity, "lm: steps %lld..%lld exceed the cache capacity %lld", t0, t0 + n_t, capacity);
  ECB_REQUIRE(n_items * n_t <= 65535ll * kRowTile, "lm: %lld rows in one call (at most %lld): split the frames", n_items * n_t,
              65535ll * kRowTile);
  return 0;
}

}  // namespace
}  // namespace ecb

using namespace ecb;

extern "C" {

int ecb_lm_create(const ecb_lm_spec* spec, ecb_lm** out) {
  ECB_REQUIRE(spec && out, "lm_create: null argument");
  ECB_REQUIRE(spec->dim >= 32 && spec->dim <= 32 * kMaxDimPerLane && spec->dim % 2 == 0, "lm_create: dim %d (even, 32..256)", spec->dim);
  ECB_REQUIRE(spec->n_heads >= 1 && spec->dim % spec->n_heads == 0 && spec->dim / spec->n_heads <= 32,
              "lm_create: head dimension must divide dim and be <= 32 (dim %d, heads %d)", spec->dim, spec->n_heads);
  ECB_REQUIRE(spec->hidden >= 32 && spec->hidden <= 1024, "lm_create: hidden %d (32..1024)", spec->hidden);
  ECB_REQUIRE(spec->card >= 2 && spec->card <= 2048, "lm_create: card %d (2..2048)", spec->card);
  ECB_REQUIRE(spec->n_q >= 1 && spec->n_layers >= 1 && spec->past_context >= 1 && spec->past_context <= 8192,
              "lm_create: n_q %d, layers %d, past_context %d (1..8192)", spec->n_q, spec->n_layers, spec->past_context);
  ECB_REQUIRE(2.0 * spec->card <= (double)(1 << 24), "lm_create: card too large for 24 range bits");
  Lm* lm = new Lm();
  lm->spec = LmSpec{spec->n_q, spec->card, spec->dim, spec->n_layers, spec->n_heads, spec->hidden, spec->past_context, spec->max_period};
  const size_t d = spec->dim, h = spec->hidden;
  size_t cur = 0;
  add_slot(*lm, "transformer.norm_in.weight", d, cur);
  add_slot(*lm, "transformer.norm_in.bias", d, cur);
  for (int l = 0; l < spec->n_layers; ++l) {
    const std::string p = "transformer.layers." + std::to_string(l);
    add_slot(*lm, p + ".self_attn.in_proj_weight", 3 * d * d, cur);
    add_slot(*lm, p + ".self_attn.in_proj_bias", 3 * d, cur);
    add_slot(*lm, p + ".self_attn.out_proj.weight", d * d, cur);
    add_slot(*lm, p + ".self_attn.out_proj.bias", d, cur);
    add_slot(*lm, p + ".linear1.weight", h * d, cur);
    add_slot(*lm, p + ".linear1.bias", h, cur);
    add_slot(*lm, p + ".linear2.weight", d * h, cur);
    add_slot(*lm, p + ".linear2.bias", d, cur);
    add_slot(*lm, p + ".norm1.weight", d, cur);
    add_slot(*lm, p + ".norm1.bias", d, cur);
    add_slot(*lm, p + ".norm2.weight", d, cur);
    add_slot(*lm, p + ".norm2.bias", d, cur);
  }
  // contiguous per-codebook tables: emb [n_q][card + 1][dim], linears weight [n_q][card][dim], bias [n_q][card]
  lm->emb_off = cur;
  for (int k = 0; k < spec->n_q; ++k) {
    lm->slots["emb." + std::to_string(k) + ".weight"] = {cur, (size_t)(spec->card + 1) * d};
    lm->loaded["emb." + std::to_string(k) + ".weight"] = false;
    cur += (size_t)(spec->card + 1) * d;
  }
  cur = (cur + 3) & ~(size_t)3;
  lm->lin_w_off = cur;
  for (int k = 0; k < spec->n_q; ++k) {
    lm->slots["linears." + std::to_string(k) + ".weight"] = {cur, (size_t)spec->card * d};
    lm->loaded["linears." + std::to_string(k) + ".weight"] = false;
    cur += (size_t)spec->card * d;
  }
  lm->lin_b_off = cur;
  for (int k = 0; k < spec->n_q; ++k) {
    lm->slots["linears." + std::to_string(k) + ".bias"] = {cur, (size_t)spec->card};
    lm->loaded["linears." + std::to_string(k) + ".bias"] = false;
    cur += (size_t)spec->card;
  }
  cur = (cur + 3) & ~(size_t)3;
  lm->pos_off = cur;
  cur += d / 2;
  lm->n_weights = cur;   // allocated by the first ecb_lm_load_tensor (create works without a device)
  *out = reinterpret_cast<ecb_lm*>(lm);
  return 0;
}

void ecb_lm_destroy(ecb_lm* h) {
  Lm* lm = reinterpret_cast<Lm*>(h);
  if (!lm) return;
  auto it = g_step_graphs.find(lm);
  if (it != g_step_graphs.end()) {
    if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
    g_step_graphs.erase(it);
  }
  if (lm->weights) cudaFree(lm->weights);
  delete lm;
}

/* key: a key of the reference's LMModel.state_dict(); data: DEVICE float32, numel values, copied. */
int ecb_lm_load_tensor(ecb_lm* h, const char* key, const float* data, int64_t numel, void* stream) {
  Lm* lm = reinterpret_cast<Lm*>(h);
  ECB_REQUIRE(lm && key && data, "lm_load_tensor: null argument");
  if (!lm->weights) {
    ECB_CUDA(cudaGetDevice(&lm->device));
    ECB_CUDA(cudaMalloc(&lm->weights, lm->n_weights * sizeof(float)));
  }
  auto it = lm->slots.find(key);
  ECB_REQUIRE(it != lm->slots.end(), "lm_load_tensor: unexpected key '%s'", key);
  ECB_REQUIRE((size_t)numel == it->second.second, "lm_load_tensor: '%s' has %lld values, expected %zu", key, (long long)numel,
              it->second.second);
  ECB_CUDA(cudaMemcpyAsync(lm->weights + it->second.first, data, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice,
                           reinterpret_cast<cudaStream_t>(stream)));
  lm->loaded[key] = true;
  lm->finalized = false;
  return 0;
}

/* pos_divisor: optional HOST float32 [dim / 2] = max_period ** (j / (dim / 2 - 1)) as the caller's framework rounds it
 * (transformer.py:23: a float32 torch pow); null: computed here in double precision and rounded once. */
int ecb_lm_finalize(ecb_lm* h, const float* pos_divisor, void* stream) {
  Lm* lm = reinterpret_cast<Lm*>(h);
  ECB_REQUIRE(lm, "lm_finalize: null handle");
  for (auto& kv : lm->loaded) ECB_REQUIRE(kv.second, "lm_finalize: tensor '%s' was not loaded", kv.first.c_str());
  const int half = lm->spec.dim / 2;
  std::vector<float> div(half);
  for (int j = 0; j < half; ++j)
    div[j] = pos_divisor ? pos_divisor[j] : (float)pow((double)lm->spec.max_period, (double)((float)j / (float)(half - 1)));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ECB_CUDA(cudaMemcpyAsync(lm->weights + lm->pos_off, div.data(), half * sizeof(float), cudaMemcpyHostToDevice, s));
  ECB_CUDA(cudaStreamSynchronize(s));   // div dies with this frame
  lm->nin_w = lm->w("transformer.norm_in.weight");
  lm->nin_b = lm->w("transformer.norm_in.bias");
  lm->layers.clear();
  for (int l = 0; l < lm->spec.n_layers; ++l) {
    const std::string p = "transformer.layers." + std::to_string(l);
    lm->layers.push_back(Lm::LayerW{lm->w(p + ".self_attn.in_proj_weight"), lm->w(p + ".self_attn.in_proj_bias"),
                                    lm->w(p + ".self_attn.out_proj.weight"), lm->w(p + ".self_attn.out_proj.bias"),
                                    lm->w(p + ".linear1.weight"), lm->w(p + ".linear1.bias"), lm->w(p + ".linear2.weight"),
                                    lm->w(p + ".linear2.bias"), lm->w(p + ".norm1.weight"), lm->w(p + ".norm1.bias"),
                                    lm->w(p + ".norm2.weight"), lm->w(p + ".norm2.bias")});
  }
  lm->finalized = true;
  return 0;
}

/* K/V cache of n_items independent streams of at most `capacity` steps: float32 [n_layers][n_items][capacity + 1][2 dim]. */
size_t ecb_lm_cache_bytes(ecb_lm* h, int64_t n_items, int64_t capacity) {
  Lm* lm = reinterpret_cast<Lm*>(h);
  if (!lm || n_items < 1 || capacity < 1) return 0;
  return (size_t)lm->spec.n_layers * n_items * (capacity + 1) * 2 * lm->spec.dim * sizeof(float);
}

size_t ecb_lm_workspace_bytes(ecb_lm* h, int64_t n_rows, int64_t n_codebooks) {
  Lm* lm = reinterpret_cast<Lm*>(h);
  if (!lm || n_rows < 1 || n_codebooks < 1) return 0;
  return plan_workspace(lm->spec, n_rows, n_codebooks, nullptr, nullptr);
}

/* LMModel.forward (model.py:65-83) for steps t0 .. t0 + n_t - 1 of n_items streams whose earlier steps are in `cache`.
 * tokens (DEVICE int64): element (item, k, t) at tokens[item * item_stride + k * k_stride + t * t_stride];
 *   tokens_are_codes = 0: the LM's input indices of the n_t steps (1 + previous code, 0 = none), t counted from t0;
 *   tokens_are_codes = 1: the codes of the whole frame, t absolute (step t is fed 1 + code[t - 1], 0 at t = 0).
 * Outputs, DEVICE, each optional: probas float32 [n_items][n_t][K][card] (the reference returns this permuted to
 * [B, card, K, T]); cdf int32, same shape: build_stable_quantized_cdf(probas[...], 24, check=False) (ac.py:18-53);
 * sym_ranges int32 [n_items][n_t][K][2] (needs tokens_are_codes = 1): (cdf[s - 1] or 0, cdf[s]) of the code s at (item, k, t),
 * the two numbers ArithmeticCoder.push reads (ac.py:143-144). */
int ecb_lm_forward(ecb_lm* h, const int64_t* tokens, int64_t item_stride, int64_t k_stride, int64_t t_stride,
                   int32_t tokens_are_codes, int64_t n_items, int64_t n_codebooks, int64_t t0, int64_t n_t, float* cache,
                   int64_t capacity, float* probas, int32_t* cdf, int32_t* sym_ranges, void* workspace,
                   size_t workspace_bytes, void* stream) {
  Lm* lm = reinterpret_cast<Lm*>(h);
  if (check_call(lm, n_items, n_codebooks, t0, n_t, capacity)) return 1;
  ECB_REQUIRE(tokens && cache && workspace, "lm_forward: null argument");
  ECB_REQUIRE(!sym_ranges || tokens_are_codes, "lm_forward: sym_ranges needs tokens_are_codes = 1");
  Workspace ws;
  const size_t need = plan_workspace(lm->spec, n_items * n_t, n_codebooks, &ws, static_cast<char*>(workspace));
  ECB_REQUIRE(workspace_bytes >= need, "lm_forward: workspace of %zu bytes, %zu needed", workspace_bytes, need);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  Tokens tk{reinterpret_cast<const long long*>(tokens), item_stride, k_stride, t_stride, tokens_are_codes ? 1 : 0};
  if (lm_trunk(*lm, tk, n_items, (int)n_codebooks, t0, nullptr, n_t, cache, capacity, ws, s)) return 1;
  return lm_heads(*lm, tk, n_items, (int)n_codebooks, t0, nullptr, n_t, ws, probas, cdf, sym_ranges, s);
}

/* The decoding loop of decompress_from_file for ONE frame (compress.py:125-152) without leaving the device: for t in
 * 0 .. n_steps - 1: LM step -> quantised cdfs -> ArithmeticDecoder.pull for the K codebooks -> codes[k][t].
 * data: DEVICE bytes of the stream, the frame's coder starts at byte `first_byte`; codes: DEVICE int64 [K][n_steps];
 * result: DEVICE int64 [2] = (status: 0 ok, 1 the stream ended sooner than expected, 2 search failed, 3 range overflow,
 * 4 bad cdf; bytes of `data` consumed up to the end of this frame -- where the next frame's scale / coder starts). */
int ecb_lm_decode_frame(ecb_lm* h, const uint8_t* data, int64_t n_bytes, int64_t first_byte, int64_t n_codebooks,
                        int64_t n_steps, int64_t* codes, float* cache, int64_t capacity, int64_t* result, void* workspace,
                        size_t workspace_bytes, void* stream) {
  Lm* lm = reinterpret_cast<Lm*>(h);
  if (check_call(lm, 1, n_codebooks, 0, n_steps, capacity)) return 1;
  ECB_REQUIRE(data && codes && cache && result && workspace, "lm_decode_frame: null argument");
  ECB_REQUIRE(first_byte >= 0 && first_byte <= n_bytes, "lm_decode_frame: first_byte %lld of %lld", (long long)first_byte,
              (long long)n_bytes);
  Workspace ws;
  const size_t need = plan_workspace(lm->spec, 1, n_codebooks, &ws, static_cast<char*>(workspace));
  ECB_REQUIRE(workspace_bytes >= need, "lm_decode_frame: workspace of %zu bytes, %zu needed", workspace_bytes, need);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int K = (int)n_codebooks;
  Tokens tk{reinterpret_cast<const long long*>(codes), 0, n_steps, 1, 1};
  long long* lcodes = reinterpret_cast<long long*>(codes);
  lm_ac_init_kernel<<<1, 32, 0, s>>>(ws.dec, first_byte * 8, ws.step);
  ECB_LAUNCHED();
  const char* genv = getenv("ECB_LM_GRAPH");   // diagnostic: 0 = launch every step's kernels from the host
  const bool use_graph = (genv ? atoi(genv) != 0 : true) && n_steps >= 8 && !prof_enabled();   // the profiler times launches
  if (!use_graph) {
    for (long long t = 0; t < n_steps; ++t) {
      if (lm_trunk_step(*lm, tk, K, t, nullptr, cache, capacity, ws, s)) return 1;
      if (lm_heads(*lm, tk, 1, K, t, nullptr, 1, ws, nullptr, ws.cdf, nullptr, s)) return 1;
      if (launch_ac_pull(ws.dec, data, n_bytes, ws.cdf, K, lm->spec.card, 24, lcodes, n_steps, t, nullptr, s)) return 1;
    }
  } else {
    // the step's launches (29 for 5 layers) as one CUDA graph: every kernel reads the step index from ws.step, the pull kernel advances it
    StepGraph& sg = g_step_graphs[lm];
    const StepGraphKey key{data, codes, cache, workspace, n_bytes, K, n_steps, capacity, cluster_step_supported(lm->spec) ? 1 : 0};
    if (!sg.exec || !(sg.key == key)) {
      if (sg.exec) {
        cudaGraphExecDestroy(sg.exec);
        sg.exec = nullptr;
      }
      if (!g_lm_capture_stream) ECB_CUDA(cudaStreamCreateWithFlags(&g_lm_capture_stream, cudaStreamNonBlocking));
      ECB_CUDA(cudaStreamBeginCapture(g_lm_capture_stream, cudaStreamCaptureModeRelaxed));
      int rc = lm_trunk_step(*lm, tk, K, 0, ws.step, cache, capacity, ws, g_lm_capture_stream);
      if (!rc) rc = lm_heads(*lm, tk, 1, K, 0, ws.step, 1, ws, nullptr, ws.cdf, nullptr, g_lm_capture_stream);
      if (!rc) rc = launch_ac_pull(ws.dec, data, n_bytes, ws.cdf, K, lm->spec.card, 24, lcodes, n_steps, 0, ws.step, g_lm_capture_stream);
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(g_lm_capture_stream, &graph);
      if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return 1;
      }
      ECB_CUDA(ce);
      const cudaError_t ie = cudaGraphInstantiate(&sg.exec, graph, 0);
      cudaGraphDestroy(graph);
      ECB_CUDA(ie);
      sg.key = key;
    }
    for (long long t = 0; t < n_steps; ++t) {
      ECB_CUDA(cudaGraphLaunch(sg.exec, s));
      g_launches.fetch_add((cluster_step_supported(lm->spec) ? 1 : 1 + 5 * lm->spec.n_layers) + 3, std::memory_order_relaxed);
    }
  }
  lm_ac_result_kernel<<<1, 32, 0, s>>>(ws.dec, reinterpret_cast<long long*>(result));
  ECB_LAUNCHED();
  return 0;
}

/* ArithmeticDecoder.pull on the DEVICE alone -- the warp-parallel decoder of the loop above against given cdfs: symbol i is
 * decoded against cdfs[i * card .. (i + 1) * card) (DEVICE int32). symbols: DEVICE int64 [n]; result: DEVICE int64 [8]
 * (result[0] = status as in ecb_lm_decode_frame, result[1] = bytes consumed; the rest is scratch). */
int ecb_ac_decode_device(const uint8_t* data, int64_t n_bytes, const int32_t* cdfs, int64_t n, int32_t card,
                         int32_t total_range_bits, int64_t* symbols, int64_t* result, void* stream) {
  ECB_REQUIRE(data && cdfs && symbols && result && n >= 1 && card >= 1, "ac_decode_device: bad arguments");
  ECB_REQUIRE(total_range_bits >= 8 && total_range_bits <= 30, "ac_decode_device: total_range_bits %d", total_range_bits);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ac::Decoder* st = reinterpret_cast<ac::Decoder*>(result + 2);
  static_assert(sizeof(ac::Decoder) <= 6 * sizeof(int64_t), "decoder state must fit result[2..8)");
  lm_ac_init_kernel<<<1, 32, 0, s>>>(st, 0, nullptr);
  ECB_LAUNCHED();
  if (launch_ac_pull(st, data, n_bytes, cdfs, n, card, total_range_bits, reinterpret_cast<long long*>(symbols), 1, 0, nullptr, s))
    return 1;
  lm_ac_result_kernel<<<1, 32, 0, s>>>(st, reinterpret_cast<long long*>(result));
  ECB_LAUNCHED();
  return 0;
}

/* build_stable_quantized_cdf (ac.py:18-53, check=False) alone: pdf DEVICE float32 [n_rows][card] -> cdf DEVICE int32. */
int ecb_quantized_cdf(const float* pdf, int64_t n_rows, int32_t card, int32_t total_range_bits, int32_t* cdf, void* stream) {
  ECB_REQUIRE(pdf && cdf && n_rows >= 1 && card >= 1 && card <= 8192, "quantized_cdf: bad arguments (card <= 8192)");
  ECB_REQUIRE(total_range_bits >= 8 && total_range_bits <= 30 && 2.0 * card <= (double)(1ull << total_range_bits),
              "quantized_cdf: total_range_bits %d", total_range_bits);
  lm_cdf_kernel<<<(unsigned)n_rows, kWarps * 32, (size_t)card * sizeof(int), reinterpret_cast<cudaStream_t>(stream)>>>(
      pdf, cdf, card, cdf_params(card, total_range_bits));
  ECB_LAUNCHED();
  return 0;
}

/* ArithmeticCoder (ac.py:56-166) on the HOST: n symbols given as their cdf ranges [low, high_exclusive) (sym_ranges: HOST
 * int32 [n][2], e.g. copied back from ecb_lm_forward), pushed in order, then flush(). out: HOST buffer; *out_len = bytes
 * the reference would have written. Fails when `capacity` is too small (4 * n + 16 always suffices). */
int ecb_ac_encode(const int32_t* sym_ranges, int64_t n, int32_t total_range_bits, uint8_t* out, int64_t capacity,
                  int64_t* out_len) {
  ECB_REQUIRE(sym_ranges && out && out_len && n >= 0, "ac_encode: null argument");
  ECB_REQUIRE(total_range_bits >= 8 && total_range_bits <= 30, "ac_encode: total_range_bits %d (ac.py:100: <= 30)", total_range_bits);
  ac::Encoder enc;
  enc.out = out;
  enc.cap = capacity;
  for (int64_t i = 0; i < n; ++i)
    if (!enc.push(sym_ranges[2 * i], sym_ranges[2 * i + 1], total_range_bits)) {
      set_error("ac_encode: symbol %lld: %s", (long long)i,
                enc.status == ac::AC_RANGE_OVERFLOW ? "range representation exceeds 62 bits" : "empty or inverted cdf range");
      return 1;
    }
  enc.flush();
  ECB_REQUIRE(!enc.overflow, "ac_encode: output buffer of %lld bytes is too small (%lld needed)", (long long)capacity,
              (long long)enc.n_bytes);
  *out_len = enc.n_bytes;
  return 0;
}

/* ArithmeticDecoder (ac.py:169-260) on the HOST -- the very function the device decoder runs: n symbols, symbol i decoded
 * against cdfs[i * card .. (i + 1) * card) (HOST int32). symbols: HOST int32 [n]; *bytes_consumed: bytes read from data.
 * Returns 0, or 1 with the reference's condition in ecb_last_error(). */
int ecb_ac_decode(const uint8_t* data, int64_t n_bytes, const int32_t* cdfs, int64_t n, int32_t card,
                  int32_t total_range_bits, int32_t* symbols, int64_t* bytes_consumed) {
  ECB_REQUIRE(data && cdfs && symbols && n >= 0 && card >= 1, "ac_decode: null argument");
  ECB_REQUIRE(total_range_bits >= 8 && total_range_bits <= 30, "ac_decode: total_range_bits %d", total_range_bits);
  ac::Decoder d;
  ac::decoder_init(d, 0);
  for (int64_t i = 0; i < n; ++i) {
    const int s = ac::pull(d, data, n_bytes * 8, cdfs + (size_t)i * card, card, total_range_bits);
    if (s < 0) {
      set_error("ac_decode: symbol %lld: %s", (long long)i,
                d.status == ac::AC_EOF ? "The stream ended sooner than expected." : d.status == ac::AC_SEARCH_FAILED ? "Binary search failed" : "invalid coder state");
      return 1;
    }
    symbols[i] = s;
  }
  if (bytes_consumed) *bytes_consumed = ac::bytes_consumed(d);
  return 0;
}

}  // extern "C"
