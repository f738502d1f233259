// Shared helpers for the encodec_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

namespace ecb {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define ECB_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess) {                                                                       \
      ::ecb::set_error("%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);       \
      return 1;                                                                                     \
    }                                                                                               \
  } while (0)

// Optional per-kernel-class timing (bench.py's roofline leg): CUDA events recorded on the launching stream
// around every launch while enabled; off by default (no events, no overhead).
enum ProfCat { PROF_CONV_GEMM = 0, PROF_CONV_IN, PROF_CONV_OUT, PROF_LSTM_REC, PROF_RVQ, PROF_GN_APPLY, PROF_MISC, PROF_TC_CONV_NARROW, PROF_TC_CONV_WIDE, PROF_TC_RES, PROF_LM_LINEAR, PROF_LM_ATTN, PROF_LM_MISC, PROF_AC_PULL, PROF_NCAT };
bool prof_enabled();
void prof_begin(int cat, cudaStream_t st, double flops, double bytes);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(int cat, cudaStream_t s, double flops, double bytes) : st(s), on(prof_enabled()) {
    if (on) prof_begin(cat, st, flops, bytes);
  }
  ~ProfScope() {
    if (on) prof_end(st);
  }
};

#define ECB_LAUNCHED()                                                                              \
  do {                                                                                              \
    ::ecb::g_launches.fetch_add(1, std::memory_order_relaxed);                                      \
    cudaError_t e__ = cudaGetLastError();                                                           \
    if (e__ != cudaSuccess) {                                                                       \
      ::ecb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 1;                                                                                     \
    }                                                                                               \
  } while (0)

#define ECB_REQUIRE(cond, ...)                                                                      \
  do {                                                                                              \
    if (!(cond)) {                                                                                  \
      ::ecb::set_error(__VA_ARGS__);                                                                \
      return 1;                                                                                     \
    }                                                                                               \
  } while (0)

// ELU(alpha = 1) (reference modules/seanet.py:43: nn.ELU): v > 0 ? v : expm1(v). expm1 by range reduction
// v = n ln2 + r, |r| <= ln2/2: expm1(v) = 2^n (expm1(r) + 1) - 1 with a degree-7 polynomial for expm1(r);
// about 1 ulp, no cancellation near 0, ~16 instructions (the epilogues apply it to every stored element).
__device__ __forceinline__ float elu1(float v) {
  const float x = fmaxf(v, -20.f);                       // exp(-20) - 1 == -1 in fp32
  const float t = fmaf(x, 1.4426950408889634f, 12582912.f);
  const float n = t - 12582912.f;                        // rint(x / ln2)
  float r = fmaf(n, -0.693145751953125f, x);             // ln2 high part: exact product
  r = fmaf(n, -1.428606765330187e-06f, r);               // ln2 low part
  float q = 1.9841270e-4f;
  q = fmaf(q, r, 1.3888889e-3f);
  q = fmaf(q, r, 8.3333333e-3f);
  q = fmaf(q, r, 4.1666667e-2f);
  q = fmaf(q, r, 1.6666667e-1f);
  q = fmaf(q, r, 0.5f);
  q = fmaf(q * r, r, r);                                 // expm1(r)
  const float s = __int_as_float((__float_as_int(t) << 23) + 0x3f800000);  // 2^n
  const float e = fmaf(q, s, s - 1.f);
  return v > 0.f ? v : e;
}

// ELU over N (even) independent values on packed fp32 pairs (fma.rn.f32x2 / add / mul .f32x2: two IEEE-rounded operations
// per issue slot, results identical to elu1) and stage by stage across all pairs, so the dependent operations of one
// expm1 overlap N/2-fold. The epilogues are instruction-issue bound by this function: ~11 issue slots per element
// instead of 20. elu(v) = max(v, 0) + expm1(clamp(v, -20, 0)).
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
template <int N>
__device__ __forceinline__ void elu_vec(float (&v)[N]) {
  static_assert(N % 2 == 0, "elu_vec works on pairs");
  constexpr int P = N / 2;
#define ECB_F2C(x) f2_pack((x), (x))
  unsigned long long x[P], t[P], n[P], r[P], q[P];
#pragma unroll
  for (int i = 0; i < P; ++i)
    x[i] = f2_pack(fminf(fmaxf(v[2 * i], -20.f), 0.f), fminf(fmaxf(v[2 * i + 1], -20.f), 0.f));
#pragma unroll
  for (int i = 0; i < P; ++i) t[i] = f2_fma(x[i], ECB_F2C(1.4426950408889634f), ECB_F2C(12582912.f));
#pragma unroll
  for (int i = 0; i < P; ++i) n[i] = f2_add(t[i], ECB_F2C(-12582912.f));          // rint(x / ln2)
#pragma unroll
  for (int i = 0; i < P; ++i) r[i] = f2_fma(n[i], ECB_F2C(-0.693145751953125f), x[i]);
#pragma unroll
  for (int i = 0; i < P; ++i) r[i] = f2_fma(n[i], ECB_F2C(-1.428606765330187e-06f), r[i]);
#pragma unroll
  for (int i = 0; i < P; ++i) q[i] = f2_fma(ECB_F2C(1.9841270e-4f), r[i], ECB_F2C(1.3888889e-3f));
#pragma unroll
  for (int i = 0; i < P; ++i) q[i] = f2_fma(q[i], r[i], ECB_F2C(8.3333333e-3f));
#pragma unroll
  for (int i = 0; i < P; ++i) q[i] = f2_fma(q[i], r[i], ECB_F2C(4.1666667e-2f));
#pragma unroll
  for (int i = 0; i < P; ++i) q[i] = f2_fma(q[i], r[i], ECB_F2C(1.6666667e-1f));
#pragma unroll
  for (int i = 0; i < P; ++i) q[i] = f2_fma(q[i], r[i], ECB_F2C(0.5f));
#pragma unroll
  for (int i = 0; i < P; ++i) q[i] = f2_fma(f2_mul(q[i], r[i]), r[i], r[i]);       // expm1(r)
#pragma unroll
  for (int i = 0; i < P; ++i) {
    float t0, t1;
    f2_unpack(t[i], t0, t1);
    const unsigned long long s = f2_pack(__int_as_float((__float_as_int(t0) << 23) + 0x3f800000),
                                         __int_as_float((__float_as_int(t1) << 23) + 0x3f800000));   // 2^n
    const unsigned long long e = f2_fma(q[i], s, f2_add(s, ECB_F2C(-1.f)));         // 2^n expm1(r) + (2^n - 1)
    float e0, e1;
    f2_unpack(e, e0, e1);
    v[2 * i] = fmaxf(v[2 * i], 0.f) + e0;
    v[2 * i + 1] = fmaxf(v[2 * i + 1], 0.f) + e1;
  }
#undef ECB_F2C
}

// ELU for values that are rounded to TF32 when they are stored (the single-pass TF32 decoder: operands carry 11 significand
// bits): exp through the special-function unit (ex2.approx, relative error ~2^-22: three orders below the rounding that
// follows) -- 6 instructions per element instead of ~12. Not for the fp32-accurate paths (exp(x) - 1 cancels near 0).
template <int N>
__device__ __forceinline__ void elu_vec_fast(float (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v[i], 0.f) * 1.4426950408889634f));
    v[i] = fmaxf(v[i], 0.f) + (e - 1.f);
  }
}
// elu_vec for fp32-accurate values, elu_vec_fast when the result is TF32-rounded anyway
template <bool FAST, int N>
__device__ __forceinline__ void elu_any(float (&v)[N]) {
  if (FAST) elu_vec_fast<N>(v);
  else elu_vec<N>(v);
}

// index of a reflect-padded signal of length T (valid while the pad is < T); conv.py:80-97
__device__ __forceinline__ int reflect_index(int r, int T) {
  if (r < 0) r = -r;
  if (r >= T) r = 2 * (T - 1) - r;
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

int sm_count();   // SMs of the current device (tc_conv.cu)

static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// Per-device once-flags: function attributes (dynamic shared memory opt-in) belong to a device, and a process may drive
// several. The flag is set after the attribute call, so a concurrent first use merely sets the attribute twice.
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  static unsigned long long bit() {
    int dev = 0;
    cudaGetDevice(&dev);
    return 1ull << (dev & 63);
  }
  bool done() const { return (mask.load(std::memory_order_acquire) & bit()) != 0; }
  void mark() { mask.fetch_or(bit(), std::memory_order_release); }
};

// Reflection length for a signal of T samples padded by (left, right): pad1d, reference conv.py:80-97.
static inline int reflect_length(long long T, long long left, long long right) {
  const long long max_pad = left > right ? left : right;
  return (int)(T > max_pad ? T : max_pad + 1);
}

// ------------------------------------------------------------------------------------------------
// Generic channels-last implicit-GEMM convolution (conv_gemm.cu)
// ------------------------------------------------------------------------------------------------
struct ConvSrc {
  const float* ptr;       // channels-last [item][T][C]
  long long item_stride;  // floats between items
  int C;                  // channels, multiple of 16
  int taps;               // kernel taps (0: source unused)
  int T;                  // valid rows per item
  int T_ref;              // reflection length: T, or max_pad + 1 when T <= max_pad (the reference then
                          // zero-extends before reflecting, conv.py:88-95); rows in [T, T_ref) read as 0
  int elu;                // apply ELU while loading
};

struct ConvParams {
  ConvSrc s0;             // main source: strided window, reflect / zero padding
  ConvSrc s1;             // optional second source (1 tap, stride 1, no padding): fused 1x1 shortcut
  const float* w;         // [taps0*C0 + taps1*C1][N], N contiguous
  const float* bias;      // [N] or nullptr
  float* out;             // item i: elements [out_lo, out_hi) of the virtual row-major [M][N] matrix,
                          // stored at out + i*out_item_stride + (m*N + n - out_lo)
  long long out_item_stride;
  long long out_lo, out_hi;
  int N;                  // output columns (multiple of 16)
  int M;                  // output rows computed per item
  int n_items;
  int stride;             // s0 window start for row m: m*stride - pad_left
  int pad_left;
  int pad_zero;           // 0: reflect padding, 1: zero padding
  int out_elu;            // apply ELU to the output
  double* stats;          // nullptr, or [item][gridDim.x*gridDim.y][2] partial (sum, sum of squares)
};

int launch_conv_gemm(const ConvParams& p, cudaStream_t stream);
int conv_gemm_stat_slots(const ConvParams& p);  // gridDim.x*gridDim.y the launch will use

// ------------------------------------------------------------------------------------------------
// Tensor-core implicit-GEMM convolution (tc_conv.cu): tcgen05 / TMEM / TMA, halo-padded activations
// ------------------------------------------------------------------------------------------------
constexpr int ACT_HALO = 16;  // reflected rows kept before and after every item of a halo-padded activation

// LSTM cell epilogue of the step-wise recurrence (tc_conv as rec = h_{t-1} W_hh^T with gate-interleaved weight columns:
// column (u / 4) * 16 + g * 4 + u % 4 holds gate g (i, f, g, o) of unit u, so every 16-column chunk carries the four gates
// of four units). The epilogue then finishes the step in registers instead of storing rec: gates = pre_t + rec,
// c = f c + i g, h = o tanh(c), out_t = h (+ skip_t) (ELU) -- reference modules/lstm.py:22-28 / nn.LSTM.
struct TcCell {
  const float* pre;        // pre-gates of this step, (item b) at pre + b * pre_stride, [4H] in reference gate order
  long long pre_stride;
  float* c;                // [B][H] cell state, updated in place
  float* h_out;            // [B][H] h_t (the next step's A operand)
  float* h_lo_out;         // [B][H] its TF32 remainder rn_tf32(h - trunc_tf32(h)) (the next step's a0_lo)
  const float* skip;       // skip input of this step (item b at skip + b * skip_stride, [H]) or nullptr
  long long skip_stride;
  float* out;              // layer output of this step, item b at out + b * out_stride, [H]
  long long out_stride;
  int H;
  int out_elu;
};

struct TcConvParams {
  // source 0: channels-last rows; a0 points at (item 0, sample a0_first, channel 0) and a0_rows samples are
  // addressable from there (reads outside are zero). Output row m reads samples m*stride - pad_left ... + taps - 1.
  const float* a0;
  long long a0_item_stride;   // floats
  int C0;                     // multiple of 32
  long long a0_first, a0_rows;
  int taps, stride, pad_left;
  // optional source 1 (1 tap, stride 1, row m <-> output row m): fused 1x1 shortcut; nullptr if unused
  const float* a1;
  long long a1_item_stride;
  int C1;
  long long a1_rows;
  const float* w_hi;          // [N][Ktot] K-major, Ktot = taps*C0 + C1 (from launch_split_weights)
  const float* w_lo;          // remainder (split == 3)
  const float* bias;          // [N] or nullptr
  float* out_raw;             // (item 0, row 0) of the raw output [M][N], or nullptr
  float* out_elu;             // same for ELU(output), or nullptr
  long long out_item_stride;  // floats (both outputs)
  int N;                      // multiple of 32
  long long M;                // output rows per item
  int n_items;
  int halo;                   // also write the reflected rows -1..-halo and M..M+halo-1 of each output (0: none)
  int round_out;              // round stored values to TF32 (for split == 1 consumers)
  int split;                  // 3: fp32-accurate split TF32 operands; 1: single TF32 pass; 2: fp32-accurate fp16 PAIR operands
                              // (w_hi / w_lo then point at the [N][Ktot] __half arrays of launch_split_weights_f16; activations
                              // beyond +-65504 saturate, so callers use it for tensors whose range they know)
  const float* a0_lo = nullptr;  // optional: the TF32 remainder of source 0, rn_tf32(a - trunc_tf32(a)), same layout as a0,
                              // written by the producer of a0 (split == 3, no second source): the tile then comes in
                              // by TMA as well and the in-kernel operand transform is skipped
  const TcCell* cell = nullptr;  // LSTM cell epilogue instead of storing the output (out_raw / out_elu / bias unused)
  int bn_max = 0;             // 0: widest N tile that divides N; else cap (32 | 64 | 128): more, shorter tiles for
                              // latency-bound launches with few rows (the per-step GEMM of the step-wise LSTM)
  double* stats;              // nullptr, or [item][tc_stat_slots(p)][2] partial (sum, sum of squares) of the raw output
                              // (GroupNorm statistics; requires out_elu == nullptr, halo == 0, round_out == 0)
  // GroupNorm of source 0 applied ON LOAD (split == 3, no a0_lo): a0 holds the RAW output of the producing conv and the operand
  // the tensor core sees is act(((a - mean) * rstd) * gamma[c] + beta[c]), computed by the transform warps on the staged tile
  // (the separate normalise pass and its tensor disappear). norm_mr = [item][2] (mean, rstd) from launch_gn_finalize.
  const float* norm_mr = nullptr;
  const float* norm_gamma = nullptr;   // [C0]
  const float* norm_beta = nullptr;    // [C0]
  int norm_elu = 0;                    // act = ELU (1) or identity (0)
};
int tc_stat_slots(const TcConvParams& p);   // partial-statistics slots per item a launch writes
int launch_tc_conv(const TcConvParams& p, cudaStream_t stream);
int launch_split_weights(const float* w, float* hi, float* lo, int K, int N, int K_pad, int N_pad, cudaStream_t s);
int launch_split_weights_f16(const float* w, void* h1, void* h2, int K, int N, int K_pad, int N_pad, cudaStream_t s);
long long tc_f16_saturation_count(int reset);
int tc_pick_bn(int N, int split, int bn_max = 0);

// Fused SEANetResnetBlock at 32 channels (tc_res.cu): Y = ELU(shortcut(X) + block3(ELU(block1(ELU(X))))), X read once.
struct TcResParams {
  const float* x;             // (item 0, sample x_first, channel 0) of the halo-padded raw input [T][32]
  long long x_item_stride;
  long long x_first, x_rows;  // first addressable sample (<= -pad_left) and number of addressable samples
  int pad_left;               // left padding of the k3 conv: 2 causal, 1 otherwise
  const float* w1_hi;         // block.1 weights [32 (hidden, zero-padded)][96] K-major split (launch_split_weights)
  const float* w1_lo;
  const float* wc_hi;         // [block.3 ; shortcut] weights [32][32 + 32] K-major split
  const float* wc_lo;
  const float* b1;            // [32] hidden bias, zero-padded
  const float* bcat;          // [32] b3 + bs
  float* out;                 // (item 0, row 0) of Y [M][32]
  long long out_item_stride;
  long long M;
  int n_items;
  int halo;                   // reflected rows to write around Y
  int split = 3;              // 3: fp32-accurate split operands; 1: one TF32 pass (x must arrive TF32-rounded)
  int round_out = 0;          // store TF32-rounded Y
};
int launch_tc_res32(const TcResParams& p, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// Edge convolutions (conv_edge.cu): audio [B,C,T] channels-first <-> 32-channel channels-last
// ------------------------------------------------------------------------------------------------
struct ConvInParams {
  const float* x;
  long long batch_stride, seg_stride, chan_stride;
  int n_seg, n_items, T, T_ref, C_in, K, pad_left;
  const float* scale;     // per item divisor or nullptr
  const float* w;         // [K*C_in][32]
  const float* bias;      // [32]
  float* out;             // raw output: (item i, row t) at out + i*out_item_stride + t*32, or nullptr
  float* out_elu;         // same layout, ELU(output), or nullptr
  long long out_item_stride;  // floats; 0 means the dense T*32
  int halo;               // also write the reflected rows -1..-halo and T..T+halo-1 of every item
  double* stats;          // nullptr or [item][gridDim.x][2]
};
int launch_conv_in(const ConvInParams& p, cudaStream_t stream);
int conv_in_stat_slots(const ConvInParams& p);

struct ConvOutParams {
  const float* in;        // (item i, row t) at in + i*in_item_stride + t*32
  long long in_item_stride;   // floats; 0 means the dense T*32
  int n_items, T, T_ref, C_out, K, pad_left;
  const float* w;         // [K][32][C_out]
  const float* bias;      // [C_out]
  const float* scale;     // per item multiplier or nullptr
  float* out;             // [item][C_out][T]
};
int launch_conv_out(const ConvOutParams& p, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// misc.cu
// ------------------------------------------------------------------------------------------------
int launch_weight_scale(const float* g, const float* v, float* scale, int dim0, int inner, cudaStream_t s);
// conv weight [Co][Ci][K] (optionally scaled per Co) -> packed [K][Ci][Co]
int launch_pack_conv(const float* w, const float* scale, float* out, int Co, int Ci, int K, cudaStream_t s);
// convtr weight [Ci][Co][K=2s] (optionally scaled per Ci) -> packed [2][Ci][s*Co]; tap 0 pairs with frame q-1
int launch_pack_convtr(const float* w, const float* scale, float* out, int Ci, int Co, int s, cudaStream_t st);
int launch_expand_bias(const float* b, float* out, int Co, int reps, cudaStream_t s);
int launch_add_vec(const float* a, const float* b, float* out, int n, cudaStream_t s);
int launch_transpose(const float* in, float* out, long long batch, int rows, int cols, cudaStream_t s);  // [b][rows][cols]->[b][cols][rows]
int launch_halo_fill(const float* src, float* row0, long long item_stride, long long T, int C, int n_items, int halo,
                     int apply_elu, cudaStream_t s);
int launch_segment_scale(const float* x, long long batch_stride, long long seg_stride, long long chan_stride,
                         int n_seg, int n_items, int T, int C, float* scale, cudaStream_t s);
struct GnSrc {
  const float* x;         // [item][rows][C] raw conv output (stored region)
  long long item_stride;  // floats between items; 0 means the dense rows*C
  const double* partial;  // [item][slots][2]
  int slots;
  double count;           // elements the statistics cover (untrimmed length * C)
  const float* gamma;     // [C]
  const float* beta;      // [C]
};
// out = act(GN(a) [+ GN(b)]), elementwise over [n_items][rows][C]
int launch_gn_apply(const GnSrc& a, const GnSrc* b, float* out, int n_items, long long rows, int C, int out_elu,
                    float eps, cudaStream_t s);
// general form: strided items, raw and/or ELU output (either may alias a.x / b->x element for element)
// round_out: store TF32-rounded values (the consumer is a single-pass TF32 conv)
int launch_gn_apply2(const GnSrc& a, const GnSrc* b, float* out_raw, float* out_elu, long long out_item_stride, int n_items,
                     long long rows, int C, float eps, cudaStream_t s, int round_out = 0, int finalized = 0);
int launch_gn_finalize(const GnSrc& a, float* mr_out, int n_items, float eps, cudaStream_t s);
// ConvLayerNorm (reference modules/norm.py:16-30, conv.py:44-46): LayerNorm over the C channels of every time step.
// out = act(LN(a) [+ LN(b)]) over [n_items][rows][C]; GnSrc::partial / slots / count are unused. C in {32,...,1024} is
// the stored row width, c_real <= C the channels that exist (the rest is zero padding, kept zero).
int launch_ln_apply2(const GnSrc& a, const GnSrc* b, float* out_raw, float* out_elu, long long out_item_stride, int n_items,
                     long long rows, int C, int c_real, float eps, cudaStream_t s);
// W_hh [4H][H] (reference layout) -> [H][4H] with gate-interleaved columns (see TcCell)
int launch_lstm_gate_interleave(const float* whh, float* out, int H, cudaStream_t s);
int launch_overlap_add(const float* frames, const int* seg_lens, long long batch, int channels, int n_seg,
                       int seg_len, int stride, float* out, long long total, cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// bitpack.cu: .ecdc code stream without entropy coding (binary.py:55-122)
// ------------------------------------------------------------------------------------------------
int launch_pack_codes(const long long* codes, long long k_stride, long long t_stride, int K, long long T, int bits,
                      unsigned char* out, cudaStream_t s);
int launch_unpack_codes(const unsigned char* in, long long n_bytes, int K, long long T, int bits, long long* codes,
                        long long k_stride, long long t_stride, cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// lstm.cu
// ------------------------------------------------------------------------------------------------
// Pre-gates pre[b][t][4H] (input projection + both biases already added) -> h sequence. w_hh is the reference's
// [4H][H] tensor. If skip != nullptr the written output is act(h + skip) (SLSTM skip connection, lstm.py:25-26) and
// the raw h stays in the recurrent state only.
int lstm_recurrent_workspace_floats(int batch, int H);
// one step of the step-wise (large-batch) recurrence: gates = pre_t + rec -> c, h_out, out_t = h (+ skip_t) (ELU)
int launch_lstm_cell(const float* pre_t, long long pre_item_stride, const float* rec, float* c, float* h_out, float* h_lo_out,
                     const float* skip_t,
                     long long skip_item_stride, float* out_t, long long out_item_stride, int batch, int H, int first, int out_elu,
                     cudaStream_t s);
// skip / out rows of item b start at + b*skip_item_stride / + b*out_item_stride floats (0 means the dense T*H).
int launch_lstm_recurrent(const float* pre, const float* w_hh_packed, const float* skip, long long skip_item_stride,
                          float* out, long long out_item_stride, int batch, int T, int H, int out_elu,
                          float* workspace, cudaStream_t s);

// lstm_tc.cu: the recurrence on the tensor cores (H = 512, up to 1024 items per launch). w_packed comes from
// launch_lstm_tc_pack (fp16 split slices of W_hh, 4H * H * 2 halves); same pre / skip / out conventions as above.
extern long long* g_lstm_tc_trace;
bool lstm_tc_supported(int batch, int H);
int lstm_tc_workspace_floats(int batch);
int launch_lstm_tc_pack(const float* whh, void* packed, int H, cudaStream_t s);
int launch_lstm_tc(const float* pre, long long pre_item_stride, const void* w_packed, const float* skip, long long skip_item_stride,
                   float* out, long long out_item_stride, int batch, int T, int out_elu, float* workspace, cudaStream_t s);
// second form of the kernel (lstm_tcw_kernel): both SLSTM layers as one wavefront kernel (up to 128 items; packings of 8 units per
// CTA from launch_lstm_tc_pack_upc: W_hh of layer 1, W_ih and W_hh of layer 2; bias2 = b_ih + b_hh of layer 2), and one layer
// with 16 units per CTA for large launches (packing of 16 units per CTA)
int launch_lstm_tc_pack_upc(const float* w, void* packed, int H, int upc, cudaStream_t s);
bool lstm_tc2_supported(int batch, int H);
int launch_lstm_tc2(const float* pre, long long pre_item_stride, const void* w1h, const void* w2x, const void* w2h, const float* bias2,
                    const float* skip, long long skip_item_stride, float* out, long long out_item_stride, int batch, int T, int out_elu,
                    float* workspace, cudaStream_t s);
int launch_lstm_tc16(const float* pre, long long pre_item_stride, const void* w_packed, const float* skip, long long skip_item_stride,
                     float* out, long long out_item_stride, int batch, int T, int out_elu, float* workspace, cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// rvq.cu
// ------------------------------------------------------------------------------------------------
int launch_rvq_prepare(const float* codebooks, long long n_q, long long bins, int dim, float* e2, cudaStream_t s);
int launch_rvq_encode(const float* frames, long long n, const float* codebooks, const float* e2, int n_q, int bins, int dim,
                      long long* codes, float* quantized, float* stack, cudaStream_t s);
// tensor-core variant (rvq_tc.cu): cb_hi / cb_lo = split codebooks from launch_rvq_split, [n_q_total][bins][128]
int launch_rvq_split(const float* codebooks, float* hi, float* lo, long long numel, cudaStream_t s);
int launch_rvq_split_f16(const float* codebooks, void* e1, void* e2, long long numel, cudaStream_t s);
long long rvq_f16_saturation_count(int reset);
// f16_pair != 0: cb_hi / cb_lo are the half arrays of launch_rvq_split_f16 (fp16 pair operands, as tc_conv's split == 2)
int launch_rvq_encode_tc(const float* frames, long long n, const float* codebooks, const float* cb_hi, const float* cb_lo,
                         const float* e2, int n_q_total, int n_q, int bins, long long* codes, float* quantized, float* stack,
                         cudaStream_t s, int f16_pair = 0);
int launch_rvq_decode(const long long* codes, long long n, const float* codebooks, int n_q, int bins, int dim,
                      float* quantized, cudaStream_t s);

}  // namespace ecb
