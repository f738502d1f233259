// Fused residual vector quantiser (fp32, CUDA cores).
//
// Replaces the Python loop of ResidualVectorQuantization.forward/encode (reference
// quantization/core_vq.py:385-432) and, per layer, EuclideanCodebook.quantize + dequantize
// (core_vq.py:178-202): dist = (|x|^2 - 2 x.E^T) + |E|^2 in that association order, arg-min with the
// lowest index on ties (== the reference's first arg-max of -dist), gather, residual -= q, out += q.
//
// One CTA carries a tile of 64 frames through ALL n_q layers: the residual tile lives in shared memory,
// the running quantised sum in registers, and each 512 KB codebook streams from L2 through a
// double-buffered cp.async pipeline in [128 entries x 32 dims] slabs. Nothing but the codes (and the final
// quantised frames) goes back to HBM.
#include "common.cuh"

namespace ecb {
namespace {

constexpr int R_FT = 64;       // frames per CTA
constexpr int R_EC = 128;      // codebook entries per slab
constexpr int R_DC = 32;       // dims per slab
constexpr int E_LD = R_DC + 4; // slab row stride (floats)
constexpr int R_THREADS = 256;
constexpr int SLAB_FLOATS = R_EC * E_LD;

struct RvqParams {
  const float* frames;     // [n][D]
  const float* codebooks;  // [n_q][bins][D]
  const float* e2;         // [n_q][bins]
  long long* codes;        // [n_q][n]
  float* quantized;        // [n][D] or nullptr
  float* stack;            // [n_q][n][D] or nullptr
  long long n;
  int n_q, bins;
};

__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  const unsigned int sa = (unsigned int)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
}

// RD = frame dimension: 128 (EnCodec 24 / 48 kHz) or 256 (the fork's 10 Hz models). A warp owns whole frames
// (RD / 128 float4 per lane), 8 frames per warp.
template <int RD>
__global__ void __launch_bounds__(R_THREADS, RD == 128 ? 2 : 1)
rvq_encode_kernel(const RvqParams p) {
  constexpr int R_LD = RD + 4;               // residual row stride (floats)
  constexpr int V = RD / 128;                // float4 per lane and frame
  constexpr int DCH = RD / R_DC;             // dim chunks per entry slab row
  extern __shared__ __align__(16) float smem[];
  float* Rs = smem;                          // [64][RD + 4]
  float* Es = Rs + R_FT * R_LD;              // [2][128][36]
  float* x2s = Es + 2 * SLAB_FLOATS;         // [64]
  int* code_s = reinterpret_cast<int*>(x2s + R_FT);  // [64]

  const int tid = threadIdx.x;
  const int tx = tid & 15;   // entries tx + 16 j
  const int ty = tid >> 4;   // frames  ty + 16 i
  const long long n0 = (long long)blockIdx.x * R_FT;
  const int chunks_per_layer = (p.bins / R_EC) * DCH;
  const int wid = tid >> 5, lane = tid & 31;
  const int n_stage = p.n_q * chunks_per_layer;

  auto issue_slab = [&](int s, int buf) {
    const int layer = s / chunks_per_layer;
    const int r = s - layer * chunks_per_layer;
    const int ec = r / DCH;
    const int dc = r % DCH;
    const float* src = p.codebooks + ((long long)layer * p.bins + ec * R_EC) * RD + dc * R_DC;
    float* dst = Es + buf * SLAB_FLOATS;
#pragma unroll
    for (int k = 0; k < (R_EC * R_DC / 4) / R_THREADS; ++k) {
      const int f = tid + k * R_THREADS;
      const int e = f >> 3;
      const int c = f & 7;
      cp_async16(dst + e * E_LD + c * 4, src + (long long)e * RD + c * 4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  issue_slab(0, 0);

  // residual tile + |x|^2; warp w owns frames w + 8 k, lane l the float4 l + 32 v of each
  float4 qacc[8][V];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int f = wid + 8 * k;
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int d4 = lane + 32 * v;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + f < p.n) x = __ldg(reinterpret_cast<const float4*>(p.frames + (n0 + f) * RD + d4 * 4));
      *reinterpret_cast<float4*>(Rs + f * R_LD + d4 * 4) = x;
      qacc[k][v] = make_float4(0.f, 0.f, 0.f, 0.f);
      s += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
    }
    s = warp_sum(s);
    if (lane == 0) x2s[f] = s;
  }

  float acc[4][8];
  float best_d[4];
  int best_i[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    best_d[i] = INFINITY;
    best_i[i] = 0;
  }

  for (int s = 0; s < n_stage; ++s) {
    const int buf = s & 1;
    const int layer = s / chunks_per_layer;
    const int r = s - layer * chunks_per_layer;
    const int ec = r / DCH;
    const int dc = r % DCH;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // slab s landed for everyone; everyone is done with slab s-1 (and with Rs updates)
    if (s + 1 < n_stage) issue_slab(s + 1, buf ^ 1);

    if (dc == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    const float* Eb = Es + buf * SLAB_FLOATS;
#pragma unroll
    for (int d4 = 0; d4 < R_DC / 4; ++d4) {
      float4 rv[4], ev[8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        rv[i] = *reinterpret_cast<const float4*>(Rs + (ty + 16 * i) * R_LD + dc * R_DC + d4 * 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) ev[j] = *reinterpret_cast<const float4*>(Eb + (tx + 16 * j) * E_LD + d4 * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] = fmaf(rv[i].x, ev[j].x, acc[i][j]);
          acc[i][j] = fmaf(rv[i].y, ev[j].y, acc[i][j]);
          acc[i][j] = fmaf(rv[i].z, ev[j].z, acc[i][j]);
          acc[i][j] = fmaf(rv[i].w, ev[j].w, acc[i][j]);
        }
    }
    if (dc == DCH - 1) {
      // distances of this slab of entries, running arg-min (strict < keeps the lowest index)
      const float* e2 = p.e2 + (long long)layer * p.bins + ec * R_EC;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = tx + 16 * j;
        const float e2v = __ldg(e2 + e);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float t = __fsub_rn(x2s[ty + 16 * i], 2.f * acc[i][j]);
          const float d = __fadd_rn(t, e2v);
          if (d < best_d[i]) {
            best_d[i] = d;
            best_i[i] = ec * R_EC + e;
          }
        }
      }
      if (r == chunks_per_layer - 1) {
        // ---- end of layer: arg-min across the 16 lanes that share the frames, gather, residual update
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, best_d[i], o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i[i], o);
            if (od < best_d[i] || (od == best_d[i] && oi < best_i[i])) {
              best_d[i] = od;
              best_i[i] = oi;
            }
          }
          if (tx == 0) {
            const int f = ty + 16 * i;
            code_s[f] = best_i[i];
            if (n0 + f < p.n) p.codes[(long long)layer * p.n + n0 + f] = (long long)best_i[i];
          }
          best_d[i] = INFINITY;
          best_i[i] = 0;
        }
        __syncthreads();
        const float* cb = p.codebooks + (long long)layer * p.bins * RD;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int f = wid + 8 * k;
          float sq = 0.f;
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const int d4 = lane + 32 * v;
            const float4 q = __ldg(reinterpret_cast<const float4*>(cb + (long long)code_s[f] * RD + d4 * 4));
            float4 x = *reinterpret_cast<float4*>(Rs + f * R_LD + d4 * 4);
            x.x -= q.x; x.y -= q.y; x.z -= q.z; x.w -= q.w;           // core_vq.py:402
            *reinterpret_cast<float4*>(Rs + f * R_LD + d4 * 4) = x;
            qacc[k][v].x += q.x; qacc[k][v].y += q.y; qacc[k][v].z += q.z; qacc[k][v].w += q.w;  // core_vq.py:404
            if (p.stack && n0 + f < p.n)
              *reinterpret_cast<float4*>(p.stack + ((long long)layer * p.n + n0 + f) * RD + d4 * 4) = q;
            sq += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
          }
          sq = warp_sum(sq);
          if (lane == 0) x2s[f] = sq;
        }
        // the __syncthreads at the top of the next stage orders these writes before the next reads
      }
    }
  }
  if (p.quantized) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int f = wid + 8 * k;
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (n0 + f < p.n) *reinterpret_cast<float4*>(p.quantized + (n0 + f) * RD + (lane + 32 * v) * 4) = qacc[k][v];
    }
  }
}

// |E|^2 per entry, one warp per entry
__global__ void rvq_e2_kernel(const float* __restrict__ cb, float* __restrict__ e2, long long rows, int dim) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float s = 0.f;
  for (int d4 = threadIdx.x & 31; d4 < dim / 4; d4 += 32) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(cb + row * dim) + d4);
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) e2[row] = s;
}

// ResidualVectorQuantization.decode (core_vq.py:434-445): out = ((E_0[c0] + E_1[c1]) + ...)
__global__ void rvq_decode_kernel(const long long* __restrict__ codes, long long n, const float* __restrict__ cb,
                                  int n_q, int bins, int RD, float* __restrict__ out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int D4 = RD / 4;
  const long long f = idx / D4;
  const int d4 = (int)(idx % D4);
  if (f >= n) return;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 0; l < n_q; ++l) {
    long long c = codes[(long long)l * n + f];
    c = c < 0 ? 0 : (c >= bins ? bins - 1 : c);
    const float4 q = __ldg(reinterpret_cast<const float4*>(cb + ((long long)l * bins + c) * RD + d4 * 4));
    a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
  }
  *reinterpret_cast<float4*>(out + f * RD + d4 * 4) = a;
}

constexpr size_t rvq_smem(int rd) { return sizeof(float) * (R_FT * (rd + 4) + 2 * SLAB_FLOATS + R_FT) + sizeof(int) * R_FT; }

}  // namespace

int launch_rvq_prepare(const float* codebooks, long long n_q, long long bins, int dim, float* e2, cudaStream_t s) {
  ECB_REQUIRE(dim == 128 || dim == 256, "rvq: dimension %d unsupported (128 or 256)", dim);
  const long long rows = n_q * bins;
  rvq_e2_kernel<<<(unsigned)cdiv(rows * 32, 256), 256, 0, s>>>(codebooks, e2, rows, dim);
  ECB_LAUNCHED();
  return 0;
}

int launch_rvq_encode(const float* frames, long long n, const float* codebooks, const float* e2, int n_q, int bins, int dim,
                      long long* codes, float* quantized, float* stack, cudaStream_t s) {
  ECB_REQUIRE(dim == 128 || dim == 256, "rvq: dimension %d unsupported (128 or 256)", dim);
  const int RD = dim;
  ECB_REQUIRE(n > 0 && n_q > 0, "rvq: empty input (n=%lld, n_q=%d)", n, n_q);
  ECB_REQUIRE(bins % R_EC == 0, "rvq: bins=%d must be a multiple of %d", bins, R_EC);
  static DeviceOnce attr_set;
  if (!attr_set.done()) {
    ECB_CUDA(cudaFuncSetAttribute(rvq_encode_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rvq_smem(128)));
    ECB_CUDA(cudaFuncSetAttribute(rvq_encode_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rvq_smem(256)));
    attr_set.mark();
  }
  RvqParams p;
  p.frames = frames;
  p.codebooks = codebooks;
  p.e2 = e2;
  p.codes = codes;
  p.quantized = quantized;
  p.stack = stack;
  p.n = n;
  p.n_q = n_q;
  p.bins = bins;
  ProfScope prof(PROF_RVQ, s, 2.0 * (double)n * n_q * bins * RD,
                 4.0 * ((double)n * RD * (quantized ? 2 : 1) + (double)n_q * bins * RD) + 8.0 * (double)n * n_q);
  if (dim == 128) rvq_encode_kernel<128><<<(unsigned)cdiv(n, R_FT), R_THREADS, rvq_smem(128), s>>>(p);
  else rvq_encode_kernel<256><<<(unsigned)cdiv(n, R_FT), R_THREADS, rvq_smem(256), s>>>(p);
  ECB_LAUNCHED();
  return 0;
}

int launch_rvq_decode(const long long* codes, long long n, const float* codebooks, int n_q, int bins, int dim,
                      float* quantized, cudaStream_t s) {
  ECB_REQUIRE(n > 0 && n_q > 0 && dim % 4 == 0, "rvq decode: empty input");
  rvq_decode_kernel<<<(unsigned)cdiv(n * (dim / 4), 256), 256, 0, s>>>(codes, n, codebooks, n_q, bins, dim, quantized);
  ECB_LAUNCHED();
  return 0;
}

}  // namespace ecb
