// Weight preparation and the small HBM-bound helpers of the codec path.
#include "common.cuh"

namespace ecb {
namespace {

// scale[d0] = g[d0] / ||v[d0, :]||  -- torch.nn.utils.weight_norm(dim=0), reference modules/conv.py:28-29
__global__ void weight_scale_kernel(const float* __restrict__ g, const float* __restrict__ v, float* __restrict__ scale,
                                    int inner) {
  const int d0 = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < inner; i += blockDim.x) {
    const double x = (double)v[(long long)d0 * inner + i];
    s += x * x;
  }
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    const float nrm = (float)sqrt(t);
    scale[d0] = g[d0] / nrm;
  }
}

// w [Co][Ci][K] -> out [K][Ci][Co], w scaled per Co (w = v * (g/||v||))
__global__ void pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ scale, float* __restrict__ out,
                                 int Co, int Ci, int K) {
  const long long n = (long long)Co * Ci * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Co);
    const long long r = i / Co;
    const int ci = (int)(r % Ci);
    const int k = (int)(r / Ci);
    float v = w[((long long)co * Ci + ci) * K + k];
    if (scale) v = v * scale[co];
    out[i] = v;
  }
}

// w [Ci][Co][2s] -> out [2][Ci][s*Co] with out[j][ci][r*Co + co] = w[ci][co][r + (1-j)*s], scaled per Ci.
// (y[q*s + r] = x[q] W[:, :, r] + x[q-1] W[:, :, r+s]; tap 0 of the 2-tap GEMM reads frame q-1.)
__global__ void pack_convtr_kernel(const float* __restrict__ w, const float* __restrict__ scale, float* __restrict__ out,
                                   int Ci, int Co, int s) {
  const long long n = 2LL * Ci * s * Co;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Co);
    long long r1 = i / Co;
    const int r = (int)(r1 % s);
    r1 /= s;
    const int ci = (int)(r1 % Ci);
    const int j = (int)(r1 / Ci);
    float v = w[((long long)ci * Co + co) * (2 * s) + r + (1 - j) * s];
    if (scale) v = v * scale[ci];
    out[i] = v;
  }
}

__global__ void expand_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int Co, int reps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Co * reps) out[i] = b[i % Co];
}

__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// [b][rows][cols] -> [b][cols][rows], 32x32 tiles through shared memory
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const long long base = (long long)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[base + (long long)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[base + (long long)c * rows + r] = tile[threadIdx.x][i];
  }
}

// W_hh [4H][H] -> out [H][4H'], column n' = (u / 4) * 16 + g * 4 + u % 4 for gate g of unit u (TcCell, common.cuh)
__global__ void gate_interleave_kernel(const float* __restrict__ whh, float* __restrict__ out, int H) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= 4LL * H * H) return;
  const int k = (int)(idx / (4 * H)), np = (int)(idx % (4 * H));
  const int chunk = np >> 4, g = (np >> 2) & 3, j = np & 3;
  const int u = chunk * 4 + j;
  out[idx] = whh[((long long)g * H + u) * H + k];
}

// Halo-padded activation layout of the tensor-core path: (item i, row r) at row0 + i*item_stride + r*C with
// r in [-halo, T + halo); rows outside [0, T) hold the reflected samples of pad1d (reference modules/conv.py:80-97):
// row -j = row j, row T-1+j = row T-1-j. With src != nullptr the interior is first copied from a plain
// [item][T][C] tensor; with src == nullptr only the halo rows are (re)written from the interior in place.
__global__ void halo_fill_kernel(const float* __restrict__ src, float* __restrict__ row0, long long item_stride, int T,
                                 int C, int halo, int apply_elu) {
  const int item = blockIdx.y;
  const int c4n = C >> 2;
  const int rows = src ? T + 2 * halo : 2 * halo;
  float* ob = row0 + (long long)item * item_stride;
  for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < (long long)rows * c4n;
       f += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(f % c4n);
    int r = (int)(f / c4n);
    if (src) r -= halo;                               // r in [-halo, T + halo)
    else r = r < halo ? r - halo : T + (r - halo);    // halo rows only
    int sr = r < 0 ? -r : (r >= T ? 2 * (T - 1) - r : r);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sr >= 0 && sr < T) {
      v = src ? __ldg(reinterpret_cast<const float4*>(src + ((long long)item * T + sr) * C) + c4)
              : *(reinterpret_cast<const float4*>(ob + (long long)sr * C) + c4);
      if (apply_elu) {
        v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w);
      }
    }
    *(reinterpret_cast<float4*>(ob + (long long)r * C) + c4) = v;
  }
}

// scale = 1e-8 + sqrt(mean_t(mean_c(x)^2))  -- EncodecModel._encode_frame, reference model.py:180-185
__global__ void segment_scale_kernel(const float* __restrict__ x, long long batch_stride, long long seg_stride,
                                     long long chan_stride, int n_seg, int T, int C, float* __restrict__ scale) {
  const int item = blockIdx.x;
  const float* xb = x + (long long)(item / n_seg) * batch_stride + (long long)(item % n_seg) * seg_stride;
  double s = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float m = 0.f;
    for (int c = 0; c < C; ++c) m += xb[(long long)c * chan_stride + t];
    m = m / (float)C;
    s += (double)(m * m);
  }
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    const float volume = sqrtf((float)(t / (double)T));
    scale[item] = 1e-8f + volume;
  }
}

// nn.GroupNorm(1, C) applied after the fact (reference modules/conv.py:50,125,162): the conv kernels emit per-CTA partial
// (sum, sumsq) in fp64; gn_finalize reduces an item's partials ONCE (one CTA per item and source) and leaves (mean, rstd)
// in the first slot, which the streaming apply kernel then reads -- two doubles per CTA instead of the whole partial list.
__global__ void __launch_bounds__(256) gn_finalize_kernel(GnSrc a, GnSrc b, float eps, float* mr_out = nullptr) {
  __shared__ double red[16];
  const GnSrc& s = blockIdx.y ? b : a;
  const int item = blockIdx.x;
  double* pp = const_cast<double*>(s.partial) + (long long)item * s.slots * 2;
  double u = 0.0, v = 0.0;
  for (int i = threadIdx.x; i < s.slots; i += blockDim.x) {
    u += pp[2 * i];
    v += pp[2 * i + 1];
  }
  for (int o = 16; o > 0; o >>= 1) {
    u += __shfl_xor_sync(0xffffffffu, u, o);
    v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[(threadIdx.x >> 5) * 2] = u;
    red[(threadIdx.x >> 5) * 2 + 1] = v;
  }
  __syncthreads();   // also: every partial of this item has been read before slot 0 is overwritten
  if (threadIdx.x == 0) {
    u = 0.0;
    v = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      u += red[2 * i];
      v += red[2 * i + 1];
    }
    const double mean = u / s.count;
    double var = v / s.count - mean * mean;
    if (var < 0.0) var = 0.0;
    pp[0] = mean;
    pp[1] = 1.0 / sqrt(var + (double)eps);
    if (mr_out && blockIdx.y == 0) {   // the float pair a consumer that normalises on load reads (tc_conv: TcConvParams::norm_mr)
      mr_out[2 * item] = (float)pp[0];
      mr_out[2 * item + 1] = (float)pp[1];
    }
  }
}

// round-to-nearest TF32 (operands of a single-pass TF32 consumer are rounded by their producer, tc_conv.cu)
__device__ __forceinline__ float gn_rn_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ float4 gn_affine(const float4& v, float mean, float rstd, const float4& g, const float4& be) {
  return make_float4((v.x - mean) * rstd * g.x + be.x, (v.y - mean) * rstd * g.y + be.y, (v.z - mean) * rstd * g.z + be.z,
                     (v.w - mean) * rstd * g.w + be.w);
}

template <bool HAS_B, bool RAW, bool ELU>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const GnSrc a, const GnSrc b, float* __restrict__ out_raw, float* __restrict__ out_elu,
                long long out_item_stride, long long rows, int C, int round_out) {
  const int item = blockIdx.y;
  const float mean_a = (float)a.partial[(long long)item * a.slots * 2], rstd_a = (float)a.partial[(long long)item * a.slots * 2 + 1];
  const float mean_b = HAS_B ? (float)b.partial[(long long)item * b.slots * 2] : 0.f;
  const float rstd_b = HAS_B ? (float)b.partial[(long long)item * b.slots * 2 + 1] : 0.f;
  const long long n4 = rows * C / 4;
  const float4* xa = reinterpret_cast<const float4*>(a.x + (long long)item * (a.item_stride ? a.item_stride : rows * C));
  const float4* xb = HAS_B ? reinterpret_cast<const float4*>(b.x + (long long)item * (b.item_stride ? b.item_stride : rows * C)) : nullptr;
  const long long ostride = out_item_stride ? out_item_stride : rows * C;
  float4* o_raw = RAW ? reinterpret_cast<float4*>(out_raw + (long long)item * ostride) : nullptr;
  float4* o_elu = ELU ? reinterpret_cast<float4*>(out_elu + (long long)item * ostride) : nullptr;
  const int c4n = C / 4;
  constexpr int U = 4;   // independent 16-byte loads in flight per thread and source
  for (long long base = (long long)blockIdx.x * (256 * U); base < n4; base += (long long)gridDim.x * (256 * U)) {
    float4 v[U], u[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long i = base + threadIdx.x + k * 256;
      if (i < n4) {
        v[k] = __ldcs(xa + i);
        if (HAS_B) u[k] = __ldcs(xb + i);
      }
    }
    float e[4 * U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long i = base + threadIdx.x + k * 256;
      if (i < n4) {
        const int c = (int)(i % c4n) * 4;
        v[k] = gn_affine(v[k], mean_a, rstd_a, __ldg(reinterpret_cast<const float4*>(a.gamma + c)),
                         __ldg(reinterpret_cast<const float4*>(a.beta + c)));
        if (HAS_B) {
          const float4 w = gn_affine(u[k], mean_b, rstd_b, __ldg(reinterpret_cast<const float4*>(b.gamma + c)),
                                     __ldg(reinterpret_cast<const float4*>(b.beta + c)));
          v[k].x += w.x; v[k].y += w.y; v[k].z += w.z; v[k].w += w.w;
        }
        if (RAW) o_raw[i] = round_out ? make_float4(gn_rn_tf32(v[k].x), gn_rn_tf32(v[k].y), gn_rn_tf32(v[k].z), gn_rn_tf32(v[k].w)) : v[k];
      }
      e[4 * k + 0] = v[k].x; e[4 * k + 1] = v[k].y; e[4 * k + 2] = v[k].z; e[4 * k + 3] = v[k].w;
    }
    if (ELU) {
      elu_vec<4 * U>(e);
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long i = base + threadIdx.x + k * 256;
        if (i < n4)
          o_elu[i] = round_out ? make_float4(gn_rn_tf32(e[4 * k + 0]), gn_rn_tf32(e[4 * k + 1]), gn_rn_tf32(e[4 * k + 2]), gn_rn_tf32(e[4 * k + 3]))
                               : make_float4(e[4 * k + 0], e[4 * k + 1], e[4 * k + 2], e[4 * k + 3]);
      }
    }
  }
}

// ConvLayerNorm = nn.LayerNorm over channels, per time step (reference modules/norm.py:16-30): one row of the channels-last
// activation is one normalisation group. LPR lanes share a row (16 bytes per lane and step, V steps), so a warp covers
// 32 / LPR rows at once; mean, then the centred second moment (two passes over registers), biased variance.
template <int C>
__global__ void __launch_bounds__(256)
ln_apply_kernel(const GnSrc a, const GnSrc b, int has_b, float* __restrict__ out_raw, float* __restrict__ out_elu,
                long long out_item_stride, long long rows, int c_real, float eps) {
  // c_real <= C: channels that exist in the model; [c_real, C) is zero padding of the stored row (the 16-channel hidden
  // layer of the 32-channel residual block is stored 32 wide) and takes no part in the statistics
  constexpr int LPR = C / 4 < 32 ? C / 4 : 32;   // lanes per row
  constexpr int V = C / (4 * LPR);               // float4 per lane
  constexpr int RPW = 32 / LPR;                  // rows per warp step
  const int item = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, li = lane % LPR;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  const float* xa = a.x + (long long)item * (a.item_stride ? a.item_stride : rows * C);
  const float* xb = has_b ? b.x + (long long)item * (b.item_stride ? b.item_stride : rows * C) : nullptr;
  const long long ostride = out_item_stride ? out_item_stride : rows * C;
  float* o_raw = out_raw ? out_raw + (long long)item * ostride : nullptr;
  float* o_elu = out_elu ? out_elu + (long long)item * ostride : nullptr;
  float4 ga[V], ba[V], gb[V], bb[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int c = (li + v * LPR) * 4;
    ga[v] = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
    ba[v] = __ldg(reinterpret_cast<const float4*>(a.beta + c));
    if (has_b) {
      gb[v] = __ldg(reinterpret_cast<const float4*>(b.gamma + c));
      bb[v] = __ldg(reinterpret_cast<const float4*>(b.beta + c));
    }
  }
  auto normalise = [&](float4 (&x)[V], const float4 (&g)[V], const float4 (&be)[V]) {
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) s += (x[v].x + x[v].y) + (x[v].z + x[v].w);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv_c = 1.f / (float)c_real;
    const float mean = s * inv_c;
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int c = (li + v * LPR) * 4;
      x[v].x = c + 0 < c_real ? x[v].x - mean : 0.f;
      x[v].y = c + 1 < c_real ? x[v].y - mean : 0.f;
      x[v].z = c + 2 < c_real ? x[v].z - mean : 0.f;
      x[v].w = c + 3 < c_real ? x[v].w - mean : 0.f;
      q += (x[v].x * x[v].x + x[v].y * x[v].y) + (x[v].z * x[v].z + x[v].w * x[v].w);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.f / sqrtf(q * inv_c + eps);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      x[v].x = x[v].x * rstd * g[v].x + be[v].x;
      x[v].y = x[v].y * rstd * g[v].y + be[v].y;
      x[v].z = x[v].z * rstd * g[v].z + be[v].z;
      x[v].w = x[v].w * rstd * g[v].w + be[v].w;
    }
  };
  // U independent row groups per warp step (more for narrow rows): their loads are issued together
  constexpr int U = V >= 4 ? 1 : (V == 2 ? 2 : 4);
  for (long long r0 = warp * (RPW * U); r0 < rows; r0 += n_warps * (RPW * U)) {
    float4 x[U][V], y[U][V];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * RPW + sub;
      live[u] = r < rows;   // whole sub-groups go idle together, the shuffles below stay inside a sub-group
#pragma unroll
      for (int v = 0; v < V; ++v) {
        x[u][v] = live[u] ? __ldcs(reinterpret_cast<const float4*>(xa + r * C) + li + v * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_b)
          y[u][v] = live[u] ? __ldcs(reinterpret_cast<const float4*>(xb + r * C) + li + v * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      normalise(x[u], ga, ba);
      if (has_b) {
        normalise(y[u], gb, bb);
#pragma unroll
        for (int v = 0; v < V; ++v) { x[u][v].x += y[u][v].x; x[u][v].y += y[u][v].y; x[u][v].z += y[u][v].z; x[u][v].w += y[u][v].w; }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!live[u]) continue;
      const long long r = r0 + u * RPW + sub;
      float e[4 * V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        if (o_raw) *(reinterpret_cast<float4*>(o_raw + r * C) + li + v * LPR) = x[u][v];
        e[4 * v + 0] = x[u][v].x; e[4 * v + 1] = x[u][v].y; e[4 * v + 2] = x[u][v].z; e[4 * v + 3] = x[u][v].w;
      }
      if (o_elu) {
        elu_vec<4 * V>(e);
#pragma unroll
        for (int v = 0; v < V; ++v)
          *(reinterpret_cast<float4*>(o_elu + r * C) + li + v * LPR) = make_float4(e[4 * v + 0], e[4 * v + 1], e[4 * v + 2], e[4 * v + 3]);
      }
    }
  }
}

// utils._linear_overlap_add, reference utils.py:17-56. weight(i) = 0.5 - |t_i - 0.5| with
// t = linspace(0, 1, seg_len + 2)[1:-1]; a shorter last frame uses the head of the same triangle.
__device__ __forceinline__ float ola_weight(int i, int seg_len) {
  const float t = (float)(i + 1) / (float)(seg_len + 1);
  return 0.5f - fabsf(t - 0.5f);
}

__global__ void overlap_add_kernel(const float* __restrict__ frames, const int* __restrict__ seg_lens, int channels,
                                   int n_seg, int seg_len, int stride, float* __restrict__ out, long long total) {
  const int bc = blockIdx.y;  // b * channels + c
  const int b = bc / channels, c = bc % channels;
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < total; n += (long long)gridDim.x * blockDim.x) {
    const int f_hi = (int)min((long long)(n_seg - 1), n / stride);
    const int f_lo = max(0, f_hi - 1);  // stride * 2 >= seg_len: at most two frames cover a sample
    float acc = 0.f, wsum = 0.f;
    // frames are accumulated in ascending order, exactly like the reference loop (utils.py:50-54)
    for (int f = f_lo; f <= f_hi; ++f) {
      const long long i = n - (long long)f * stride;
      if (i < 0 || i >= __ldg(seg_lens + f)) continue;
      const float wv = ola_weight((int)i, seg_len);
      const float xv = frames[(((long long)b * n_seg + f) * channels + c) * seg_len + i];
      acc = __fadd_rn(acc, __fmul_rn(wv, xv));
      wsum = __fadd_rn(wsum, wv);
    }
    out[(long long)bc * total + n] = acc / wsum;
  }
}

}  // namespace

int launch_weight_scale(const float* g, const float* v, float* scale, int dim0, int inner, cudaStream_t s) {
  weight_scale_kernel<<<dim0, 256, 0, s>>>(g, v, scale, inner);
  ECB_LAUNCHED();
  return 0;
}
int launch_pack_conv(const float* w, const float* scale, float* out, int Co, int Ci, int K, cudaStream_t s) {
  const long long n = (long long)Co * Ci * K;
  pack_conv_kernel<<<(unsigned)min(cdiv(n, 256), 4096LL), 256, 0, s>>>(w, scale, out, Co, Ci, K);
  ECB_LAUNCHED();
  return 0;
}
int launch_pack_convtr(const float* w, const float* scale, float* out, int Ci, int Co, int s, cudaStream_t st) {
  const long long n = 2LL * Ci * s * Co;
  pack_convtr_kernel<<<(unsigned)min(cdiv(n, 256), 4096LL), 256, 0, st>>>(w, scale, out, Ci, Co, s);
  ECB_LAUNCHED();
  return 0;
}
int launch_expand_bias(const float* b, float* out, int Co, int reps, cudaStream_t s) {
  expand_bias_kernel<<<(unsigned)cdiv((long long)Co * reps, 256), 256, 0, s>>>(b, out, Co, reps);
  ECB_LAUNCHED();
  return 0;
}
int launch_add_vec(const float* a, const float* b, float* out, int n, cudaStream_t s) {
  add_vec_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(a, b, out, n);
  ECB_LAUNCHED();
  return 0;
}
int launch_transpose(const float* in, float* out, long long batch, int rows, int cols, cudaStream_t s) {
  ECB_REQUIRE(batch > 0 && batch <= 65535, "transpose: bad batch %lld", batch);
  dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 32), (unsigned)batch);
  ECB_REQUIRE(grid.y <= 65535, "transpose: too many rows");
  ProfScope prof(PROF_MISC, s, 0.0, 8.0 * (double)batch * rows * cols);
  transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(in, out, rows, cols);
  ECB_LAUNCHED();
  return 0;
}
int launch_halo_fill(const float* src, float* row0, long long item_stride, long long T, int C, int n_items, int halo,
                     int apply_elu, cudaStream_t s) {
  ECB_REQUIRE(C % 4 == 0 && T >= 2 && n_items > 0 && n_items <= 65535, "halo_fill: bad shape T=%lld C=%d", T, C);
  const long long rows = src ? T + 2 * halo : 2 * halo;
  const long long work = rows * (C / 4);
  dim3 grid((unsigned)(cdiv(work, 256) < 1024 ? cdiv(work, 256) : 1024), (unsigned)n_items);
  ProfScope prof(PROF_MISC, s, 0.0, 8.0 * rows * C * n_items);
  halo_fill_kernel<<<grid, 256, 0, s>>>(src, row0, item_stride, (int)T, C, halo, apply_elu);
  ECB_LAUNCHED();
  return 0;
}

int launch_segment_scale(const float* x, long long batch_stride, long long seg_stride, long long chan_stride,
                         int n_seg, int n_items, int T, int C, float* scale, cudaStream_t s) {
  ProfScope prof(PROF_MISC, s, 0.0, 4.0 * (double)n_items * T * C);
  segment_scale_kernel<<<n_items, 512, 0, s>>>(x, batch_stride, seg_stride, chan_stride, n_seg, T, C, scale);
  ECB_LAUNCHED();
  return 0;
}
int launch_gn_apply2(const GnSrc& a, const GnSrc* b, float* out_raw, float* out_elu, long long out_item_stride, int n_items,
                     long long rows, int C, float eps, cudaStream_t s, int round_out, int finalized) {
  ECB_REQUIRE(C % 4 == 0 && (out_raw || out_elu), "gn_apply: C=%d", C);
  const long long n4 = rows * C / 4;
  dim3 grid((unsigned)min(cdiv(n4, 256 * 4), 4096LL), (unsigned)n_items);
  ProfScope prof(PROF_GN_APPLY, s, 0.0,
                 4.0 * (double)rows * C * n_items * ((b ? 2 : 1) + (out_raw ? 1 : 0) + (out_elu ? 1 : 0)));
  if (!finalized) gn_finalize_kernel<<<dim3((unsigned)n_items, b ? 2 : 1), 256, 0, s>>>(a, b ? *b : a, eps);
  const GnSrc bb = b ? *b : a;
#define ECB_GN_LAUNCH(HB, R, E) gn_apply_kernel<HB, R, E><<<grid, 256, 0, s>>>(a, bb, out_raw, out_elu, out_item_stride, rows, C, round_out)
  if (b) {
    if (out_raw && out_elu) ECB_GN_LAUNCH(true, true, true);
    else if (out_raw) ECB_GN_LAUNCH(true, true, false);
    else ECB_GN_LAUNCH(true, false, true);
  } else {
    if (out_raw && out_elu) ECB_GN_LAUNCH(false, true, true);
    else if (out_raw) ECB_GN_LAUNCH(false, true, false);
    else ECB_GN_LAUNCH(false, false, true);
  }
#undef ECB_GN_LAUNCH
  ECB_LAUNCHED();
  return 0;
}
// statistics only: (mean, rstd) of every item -> mr_out [n_items][2] floats (and slot 0 of the partial list, as gn_apply expects)
int launch_gn_finalize(const GnSrc& a, float* mr_out, int n_items, float eps, cudaStream_t s) {
  gn_finalize_kernel<<<dim3((unsigned)n_items, 1), 256, 0, s>>>(a, a, eps, mr_out);
  ECB_LAUNCHED();
  return 0;
}
int launch_gn_apply(const GnSrc& a, const GnSrc* b, float* out, int n_items, long long rows, int C, int out_elu,
                    float eps, cudaStream_t s) {
  return launch_gn_apply2(a, b, out_elu ? nullptr : out, out_elu ? out : nullptr, 0, n_items, rows, C, eps, s);
}
int launch_ln_apply2(const GnSrc& a, const GnSrc* b, float* out_raw, float* out_elu, long long out_item_stride, int n_items,
                     long long rows, int C, int c_real, float eps, cudaStream_t s) {
  ECB_REQUIRE((out_raw || out_elu) && c_real >= 1 && c_real <= C, "ln_apply: bad argument (C=%d, real %d)", C, c_real);
  const int lpr = C / 4 < 32 ? C / 4 : 32;
  const int vv = C / (4 * lpr);
  const int uu = vv >= 4 ? 1 : (vv == 2 ? 2 : 4);
  const long long warps = cdiv(rows, (32 / lpr) * uu);
  dim3 grid((unsigned)min(cdiv(warps, 8), 4096LL), (unsigned)n_items);
  ProfScope prof(PROF_GN_APPLY, s, 0.0,
                 4.0 * (double)rows * C * n_items * ((b ? 2 : 1) + (out_raw ? 1 : 0) + (out_elu ? 1 : 0)));
  const GnSrc bb = b ? *b : a;
#define ECB_LN_CASE(CC) \
  case CC: ln_apply_kernel<CC><<<grid, 256, 0, s>>>(a, bb, b ? 1 : 0, out_raw, out_elu, out_item_stride, rows, c_real, eps); break;
  switch (C) {
    ECB_LN_CASE(32)
    ECB_LN_CASE(64)
    ECB_LN_CASE(128)
    ECB_LN_CASE(256)
    ECB_LN_CASE(512)
    ECB_LN_CASE(1024)
    default:
      ECB_REQUIRE(false, "ln_apply: C=%d unsupported (32, 64, ..., 1024)", C);
  }
#undef ECB_LN_CASE
  ECB_LAUNCHED();
  return 0;
}
int launch_lstm_gate_interleave(const float* whh, float* out, int H, cudaStream_t s) {
  gate_interleave_kernel<<<(unsigned)cdiv(4LL * H * H, 256), 256, 0, s>>>(whh, out, H);
  ECB_LAUNCHED();
  return 0;
}
int launch_overlap_add(const float* frames, const int* seg_lens, long long batch, int channels, int n_seg,
                       int seg_len, int stride, float* out, long long total, cudaStream_t s) {
  ECB_REQUIRE(batch * channels <= 65535, "overlap_add: batch*channels too large");
  dim3 grid((unsigned)min(cdiv(total, 256), 8192LL), (unsigned)(batch * channels));
  ProfScope prof(PROF_MISC, s, 0.0, 8.0 * (double)batch * channels * total);
  overlap_add_kernel<<<grid, 256, 0, s>>>(frames, seg_lens, channels, n_seg, seg_len, stride, out, total);
  ECB_LAUNCHED();
  return 0;
}

}  // namespace ecb
