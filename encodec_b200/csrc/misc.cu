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

// nn.GroupNorm(1, C) applied after the fact (reference modules/conv.py:50,125,162): the conv kernels emit
// per-CTA partial (sum, sumsq); every CTA here re-reduces its item's partials (a few hundred doubles).
__device__ __forceinline__ void gn_coeffs(const GnSrc& s, int item, float eps, float* mean_out, float* rstd_out,
                                          double* red) {
  double a = 0.0, b = 0.0;
  const double* pp = s.partial + (long long)item * s.slots * 2;
  for (int i = threadIdx.x; i < s.slots; i += blockDim.x) {
    a += pp[2 * i];
    b += pp[2 * i + 1];
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    red[(threadIdx.x >> 5) * 2] = a;
    red[(threadIdx.x >> 5) * 2 + 1] = b;
  }
  __syncthreads();
  a = 0.0;
  b = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
    a += red[2 * i];
    b += red[2 * i + 1];
  }
  const double mean = a / s.count;
  double var = b / s.count - mean * mean;
  if (var < 0.0) var = 0.0;
  *mean_out = (float)mean;
  *rstd_out = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void __launch_bounds__(256)
gn_apply_kernel(const GnSrc a, const GnSrc b, int has_b, float* __restrict__ out_raw, float* __restrict__ out_elu,
                long long out_item_stride, long long rows, int C, float eps) {
  __shared__ double red[16];
  const int item = blockIdx.y;
  float mean_a, rstd_a, mean_b = 0.f, rstd_b = 0.f;
  gn_coeffs(a, item, eps, &mean_a, &rstd_a, red);
  if (has_b) gn_coeffs(b, item, eps, &mean_b, &rstd_b, red);
  const long long n4 = rows * C / 4;
  const float4* xa = reinterpret_cast<const float4*>(a.x + (long long)item * (a.item_stride ? a.item_stride : rows * C));
  const float4* xb = has_b ? reinterpret_cast<const float4*>(b.x + (long long)item * (b.item_stride ? b.item_stride : rows * C)) : nullptr;
  const long long ostride = out_item_stride ? out_item_stride : rows * C;
  float4* o_raw = out_raw ? reinterpret_cast<float4*>(out_raw + (long long)item * ostride) : nullptr;
  float4* o_elu = out_elu ? reinterpret_cast<float4*>(out_elu + (long long)item * ostride) : nullptr;
  const int c4n = C / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    float4 v = xa[i];
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(a.beta + c));
    v.x = (v.x - mean_a) * rstd_a * g.x + be.x;
    v.y = (v.y - mean_a) * rstd_a * g.y + be.y;
    v.z = (v.z - mean_a) * rstd_a * g.z + be.z;
    v.w = (v.w - mean_a) * rstd_a * g.w + be.w;
    if (has_b) {
      float4 u = xb[i];
      const float4 g2 = __ldg(reinterpret_cast<const float4*>(b.gamma + c));
      const float4 b2 = __ldg(reinterpret_cast<const float4*>(b.beta + c));
      v.x += (u.x - mean_b) * rstd_b * g2.x + b2.x;
      v.y += (u.y - mean_b) * rstd_b * g2.y + b2.y;
      v.z += (u.z - mean_b) * rstd_b * g2.z + b2.z;
      v.w += (u.w - mean_b) * rstd_b * g2.w + b2.w;
    }
    if (o_raw) o_raw[i] = v;
    if (o_elu) o_elu[i] = make_float4(elu1(v.x), elu1(v.y), elu1(v.z), elu1(v.w));
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
                     long long rows, int C, float eps, cudaStream_t s) {
  ECB_REQUIRE(C % 4 == 0 && (out_raw || out_elu), "gn_apply: C=%d", C);
  const long long n4 = rows * C / 4;
  dim3 grid((unsigned)min(cdiv(n4, 256 * 4), 4096LL), (unsigned)n_items);
  ProfScope prof(PROF_GN_APPLY, s, 0.0,
                 4.0 * (double)rows * C * n_items * ((b ? 2 : 1) + (out_raw ? 1 : 0) + (out_elu ? 1 : 0)));
  gn_apply_kernel<<<grid, 256, 0, s>>>(a, b ? *b : a, b ? 1 : 0, out_raw, out_elu, out_item_stride, rows, C, eps);
  ECB_LAUNCHED();
  return 0;
}
int launch_gn_apply(const GnSrc& a, const GnSrc* b, float* out, int n_items, long long rows, int C, int out_elu,
                    float eps, cudaStream_t s) {
  return launch_gn_apply2(a, b, out_elu ? nullptr : out, out_elu ? out : nullptr, 0, n_items, rows, C, eps, s);
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
