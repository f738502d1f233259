// SLSTM recurrence on the tensor cores (reference modules/lstm.py:12-28 -> nn.LSTM(512, 512, 2), gates i,f,g,o, zero
// state), H = 512. One persistent kernel per LSTM layer; the input projection of all steps is a tc_conv GEMM (pre).
//
// Geometry: 128 CTAs, CTA j owns hidden units 4j .. 4j+3 = 16 gate rows of W_hh for ALL items. Its [16 x 512] slice sits in
// shared memory for the whole sequence (un-replicated: 128 slices = W_hh exactly once on the chip), and the recurrent
// product of a group of 64 items is ONE tensor-core contraction per step,
//     rec[64 items][16 gate columns] = h_{t-1}[64][512] . W_slice^T        (tcgen05.mma, M = 64, N = 16 | 32, K = 512)
// with the accumulator in TMEM. What moves per step is h_{t-1} (all 128 CTAs need all of it): every CTA publishes its 4
// units of h_t to an L2-resident buffer, bumps an arrival counter, and pulls the complete h_t back with TMA.
//
// fp32 accuracy from fp16 tensor-core operands. fp16 has the 11-bit significand of TF32 at half the bytes and twice the
// MMA rate, and h in (-1, 1) / LSTM weights fit its range. Operands are split the way tc_conv splits TF32:
//     h = h1 + 2^-11 h2,  h1 = fp16(h), h2 = fp16((h - h1) 2^11)        (written by the producer of h)
//     w = w1 + 2^-11 w2                                                 (split once at load, lstm_tc_pack)
//     rec = sum h1 w1  +  2^-11 (sum h1 w2 + sum h2 w1)                 (h2 w2 2^-22 dropped, as in the 3xTF32 scheme)
// fp16 x fp16 products are exact in the fp32 accumulator. Per K step of 16: one MMA with N = 32, D[main | corr] (+)=
// h1 . [w1 | w2], and one with N = 16, D[corr] += h2 . w1. The tensor core truncates when it adds into its accumulator
// (tc_conv.cu), so each K half gets its own accumulator pair (16 accumulating MMAs each) and the epilogue adds the halves.
//
// Roles (12 warps): warp 0 = loader (polls the 8 per-K-tile arrival counters of the group with one coalesced acquire load,
// then TMA-loads h1 / h2 tiles [64 items x 64 units], SWIZZLE_128B, into an 8-stage ring), warp 1 = MMA issuer, warps 4-11 =
// cell epilogue: tcgen05.ld the 16 gate columns, + pre-gates, cell update (c stays in shared memory), h_t -> (h1, h2) fp16 to
// L2, layer output y = h (+ skip) (ELU) to HBM; one CTA barrier, one fence, one release-increment of the counter.
// Batches above 64 items run as independent groups of 64 through the same ring (two TMEM accumulator buffers), so the
// exchange latency of one group hides behind the loads and MMAs of the others.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ecb {
namespace {
using namespace tc;

constexpr int LT_H = 512;
constexpr int LT_CTAS = 128;        // unit blocks
constexpr int LT_UNITS = 4;         // hidden units per CTA
constexpr int LT_NCOL = 16;         // gate columns per CTA: n = unit * 4 + gate
constexpr int LT_ITEMS = 64;        // items per group = MMA M
constexpr int LT_KT = 64;           // K elements per shared-memory tile (128 bytes of fp16)
constexpr int LT_NKT = LT_H / LT_KT;   // 8
constexpr int LT_STAGES = 8;
constexpr int LT_THREADS = 384;
constexpr int LT_A_TILE = LT_ITEMS * 128;         // 8 KB
constexpr int LT_STAGE_BYTES = 2 * LT_A_TILE;     // h1 tile + h2 tile
constexpr int LT_W_TILE = 2 * LT_NCOL * 128;      // [w1 (16 rows) | w2 (16 rows)] x 64 k = 4 KB
constexpr int LT_ACC_COLS = 64;                   // per accumulator buffer: 2 K halves x [main 16 | corr 16]
constexpr int LT_MAX_GROUPS = 16;
constexpr float LT_LO_SCALE = 2048.f;             // 2^11

struct LstmTcParams {
  const float* pre;      // item b at pre + b * pre_stride, [T][4H] in reference gate order (i, f, g, o)
  long long pre_stride;
  const float* skip;     // item b at skip + b * skip_stride, [T][H], or nullptr
  long long skip_stride;
  float* out;            // item b at out + b * out_stride, [T][H]
  long long out_stride;
  __half* h1g;           // [2][G * 64][512] exchange buffers (L2-resident)
  __half* h2g;
  unsigned int* cnt;     // [G][8] arrival counters (one per group and K tile), zeroed by the host
  int B, T, G, out_elu;
  long long* trace;      // diagnostic: [3 roles][LT_TR_STEPS][16] clock64 stamps of CTA 0, or nullptr
};
constexpr int LT_TR_STEPS = 8, LT_TR_T0 = 20;
#define LT_TRACE(role, t, ev)                                                                                     \
  if (p.trace && cta == 0 && (t) >= LT_TR_T0 && (t) < LT_TR_T0 + LT_TR_STEPS && lane == 0)                           \
    p.trace[((role) * LT_TR_STEPS + ((t) - LT_TR_T0)) * 16 + (ev)] = clock64();

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// kind::f16 instruction descriptor: D fp32 (bits [4,6) = 1), A / B fp16 (format 0), both K-major, N >> 3 in [17,23), M >> 4 in [24,29)
constexpr uint32_t idesc_f16(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap map_h1, const __grid_constant__ CUtensorMap map_h2,
               const __grid_constant__ CUtensorMap map_w, const LstmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;             // SWIZZLE_128B tiles: 1024-byte aligned
  uint8_t* smem_gen = smem_raw + (base - raw_addr);
  const uint32_t a_ring = base;                                   // [LT_STAGES][h1 tile | h2 tile]
  const uint32_t w_smem = base + LT_STAGES * LT_STAGE_BYTES;      // [LT_NKT][LT_W_TILE]
  constexpr int CS_OFF = LT_STAGES * LT_STAGE_BYTES + LT_NKT * LT_W_TILE;
  float* cs = reinterpret_cast<float*>(smem_gen + CS_OFF);        // [G][64 items][4 units] cell state
  const int cs_bytes = p.G * LT_ITEMS * LT_UNITS * 4;
  const uint32_t bar_base = base + CS_OFF + cs_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (LT_STAGES + s); };
  auto accf_bar = [&](int b) { return bar_base + 8u * (2 * LT_STAGES + b); };
  auto acce_bar = [&](int b) { return bar_base + 8u * (2 * LT_STAGES + 2 + b); };
  const uint32_t w_bar = bar_base + 8u * (2 * LT_STAGES + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + CS_OFF + cs_bytes + 8 * (2 * LT_STAGES + 5));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < LT_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf_bar(b), 1);
      mbar_init(acce_bar(b), 8);   // one arrive per epilogue warp
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.G * LT_ITEMS * LT_UNITS; i += LT_THREADS) cs[i] = 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * LT_ACC_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int rows_per_buf = p.G * LT_ITEMS;

  if (warp == 0) {
    // ================================ loader ================================
    if (elect_one()) {   // this CTA's weight slice, once: 8 tiles [32 rows x 64 k]
      mbar_expect_tx(w_bar, LT_NKT * LT_W_TILE);
      for (int j = 0; j < LT_NKT; ++j) tma_load_2d(w_smem + j * LT_W_TILE, &map_w, w_bar, j * LT_KT, cta * 2 * LT_NCOL);
    }
    __syncwarp();
    uint32_t it = 0;
    for (int t = 0; t < p.T; ++t) {
      const unsigned int target = 16u * (unsigned int)t;   // 16 CTAs publish each K tile of 64 units
      const int row0 = (t & 1) * rows_per_buf;
      for (int g = 0; g < p.G; ++g) {
        int next = 0;
        unsigned int spins = 0;
        if (g == 0) LT_TRACE(0, t, 0)
        while (next < LT_NKT) {
          unsigned int v = 0xffffffffu;
          if (t > 0 && lane < LT_NKT)
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.cnt + g * LT_NKT + lane) : "memory");
          const unsigned int ready = __ballot_sync(0xffffffffu, v >= target);
          if (!((ready >> next) & 1u)) {
            if (++spins > (1u << 22)) __trap();   // a lost arrival must not hang the device
            continue;
          }
          while (next < LT_NKT && ((ready >> next) & 1u)) {
            const int st = (int)(it % LT_STAGES);
            mbar_wait(empty_bar(st), ((it / LT_STAGES) & 1u) ^ 1u);
            if (elect_one()) {
              asm volatile("fence.proxy.async;" ::: "memory");   // h was written through the generic proxy (other SMs), read by TMA
              mbar_expect_tx(full_bar(st), LT_STAGE_BYTES);
              tma_load_2d(a_ring + st * LT_STAGE_BYTES, &map_h1, full_bar(st), next * LT_KT, row0 + g * LT_ITEMS);
              tma_load_2d(a_ring + st * LT_STAGE_BYTES + LT_A_TILE, &map_h2, full_bar(st), next * LT_KT, row0 + g * LT_ITEMS);
            }
            __syncwarp();
            if (g == 0) LT_TRACE(0, t, 1 + next)
            ++next;
            ++it;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc32 = idesc_f16(LT_ITEMS, 2 * LT_NCOL);
    constexpr uint32_t idesc16 = idesc_f16(LT_ITEMS, LT_NCOL);
    constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
    auto mk_desc = [](uint32_t addr) { return ((uint64_t)DESC_HI << 32) | (uint64_t)(((addr & 0x3FFFFu) >> 4) | (1u << 16)); };
    mbar_wait(w_bar, 0);
    uint32_t it = 0, n = 0;
    for (int t = 0; t < p.T; ++t) {
      for (int g = 0; g < p.G; ++g, ++n) {
        const int acc = (int)(n & 1u);
        mbar_wait(acce_bar(acc), ((n >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator buffer
        for (int j = 0; j < LT_NKT; ++j, ++it) {
          const int st = (int)(it % LT_STAGES);
          mbar_wait(full_bar(st), (it / LT_STAGES) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (g == 0) LT_TRACE(1, t, j)
          if (elect_one()) {
            const uint32_t d = tmem_base + (uint32_t)(acc * LT_ACC_COLS + (j / (LT_NKT / 2)) * 32);
            const uint64_t da1 = mk_desc(a_ring + st * LT_STAGE_BYTES);
            const uint64_t da2 = mk_desc(a_ring + st * LT_STAGE_BYTES + LT_A_TILE);
            const uint64_t db = mk_desc(w_smem + j * LT_W_TILE);
#pragma unroll
            for (int k = 0; k < LT_KT / 16; ++k) {
              mma_f16(d, da1 + 2u * k, db + 2u * k, idesc32, (j % (LT_NKT / 2) != 0 || k != 0) ? 1u : 0u);   // [main | corr] (+)= h1 [w1 | w2]
              mma_f16(d + LT_NCOL, da2 + 2u * k, db + 2u * k, idesc16, 1u);                                    // corr += h2 w1
            }
            tcgen05_commit(empty_bar(st));
            if (j + 1 == LT_NKT) tcgen05_commit(accf_bar(acc));
          }
          __syncwarp();
          if (g == 0 && j + 1 == LT_NKT) LT_TRACE(1, t, 8)
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ cell epilogue ================================
    // M = 64 accumulator layout: item row r of the group sits in TMEM lane 32 (r / 16) + r % 16. Warp w reads lane quadrant
    // w % 4 (hardware rule), lanes 0-15 hold items; warps 4-7 take units 0-1 (columns 0-7), warps 8-11 units 2-3. The values
    // of the second unit move to lanes 16-31, so every lane finishes one (item, unit) pair.
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = lane & 15;
    const int u = 2 * half + (lane >> 4);
    const int unit = cta * LT_UNITS + u;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(8 * half);
    uint32_t n = 0;
    for (int t = 0; t < p.T; ++t) {
      __half* h1n = p.h1g + (long long)((t + 1) & 1) * rows_per_buf * LT_H;
      __half* h2n = p.h2g + (long long)((t + 1) & 1) * rows_per_buf * LT_H;
      for (int g = 0; g < p.G; ++g, ++n) {
        const int acc = (int)(n & 1u);
        const int il = quad * 16 + r;               // item within the group
        const int item = g * LT_ITEMS + il;
        const bool valid = item < p.B;
        // pre-gates / skip input of this step: independent of the recurrence, in flight while we wait for the MMAs
        float pg[4] = {0.f, 0.f, 0.f, 0.f}, skipv = 0.f;
        if (valid) {
          const float* pr = p.pre + (long long)item * p.pre_stride + (long long)t * (4 * LT_H) + unit;
#pragma unroll
          for (int k = 0; k < 4; ++k) pg[k] = __ldg(pr + k * LT_H);
          if (p.skip) skipv = __ldg(p.skip + (long long)item * p.skip_stride + (long long)t * LT_H + unit);
        }
        if (g == 0 && warp == 4) LT_TRACE(2, t, 0)
        mbar_wait(accf_bar(acc), (n >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (g == 0 && warp == 4) LT_TRACE(2, t, 1)
        float ma[8], mb[8], ca[8], cb[8];
        const uint32_t a0 = lane_base + (uint32_t)(acc * LT_ACC_COLS);
        tmem_ld8(a0, ma);
        tmem_ld8(a0 + LT_NCOL, ca);
        tmem_ld8(a0 + 32, mb);
        tmem_ld8(a0 + 32 + LT_NCOL, cb);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(acce_bar(acc));
        if (g == 0 && warp == 4) LT_TRACE(2, t, 2)
        float rec[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float lo = (ma[k] + mb[k]) + (ca[k] + cb[k]) * (1.f / LT_LO_SCALE);
          const float hi = (ma[4 + k] + mb[4 + k]) + (ca[4 + k] + cb[4 + k]) * (1.f / LT_LO_SCALE);
          const float other = __shfl_sync(0xffffffffu, hi, r);   // second unit of the item, computed by lane r
          rec[k] = lane < 16 ? lo : other;
        }
        const float gi = sigmoid_acc(pg[0] + rec[0]);
        const float gf = sigmoid_acc(pg[1] + rec[1]);
        const float gg = tanhf(pg[2] + rec[2]);
        const float go = sigmoid_acc(pg[3] + rec[3]);
        float* cptr = cs + (g * LT_ITEMS + il) * LT_UNITS + u;
        const float c_new = gf * (*cptr) + gi * gg;
        *cptr = c_new;
        const float h_new = go * tanhf(c_new);
        const __half q1 = __float2half_rn(h_new);
        const __half q2 = __float2half_rn((h_new - __half2float(q1)) * LT_LO_SCALE);
        const unsigned int own = (unsigned int)__half_as_ushort(q1) | ((unsigned int)__half_as_ushort(q2) << 16);
        const unsigned int oth = __shfl_down_sync(0xffffffffu, own, 16);
        if (lane < 16 && t + 1 < p.T) {   // units (2 half, 2 half + 1) of this item: one 4-byte store per split part
          const long long off = (long long)(g * LT_ITEMS + il) * LT_H + cta * LT_UNITS + 2 * half;
          const unsigned int v1 = (own & 0xffffu) | (oth << 16);
          const unsigned int v2 = (own >> 16) | (oth & 0xffff0000u);
          asm volatile("st.global.cg.b32 [%0], %1;" ::"l"(h1n + off), "r"(v1) : "memory");
          asm volatile("st.global.cg.b32 [%0], %1;" ::"l"(h2n + off), "r"(v2) : "memory");
        }
        if (valid) {
          float y = h_new + skipv;
          if (p.out_elu) y = elu1(y);
          p.out[(long long)item * p.out_stride + (long long)t * LT_H + unit] = y;
        }
        if (g == 0 && warp == 4) LT_TRACE(2, t, 3)
        if (t + 1 < p.T) {
          asm volatile("bar.sync 1, 256;" ::: "memory");   // all 8 epilogue warps have stored their part of h_t
          if (g == 0 && warp == 4) LT_TRACE(2, t, 4)
          if (warp == 4 && lane == 0) {
            __threadfence();
            if (g == 0) LT_TRACE(2, t, 5)
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.cnt + g * LT_NKT + cta / 16) : "memory");
            if (g == 0) LT_TRACE(2, t, 6)
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * LT_ACC_COLS) : "memory");
  }
}

// W_hh [4H][H] (reference layout) -> fp16 split slices [128 CTAs][w1: 16 rows | w2: 16 rows][512], row n = unit * 4 + gate
__global__ void lstm_tc_pack_kernel(const float* __restrict__ whh, __half* __restrict__ wp) {
  const int row = blockIdx.x;                 // 0 .. 128 * 32 - 1
  const int cta = row / (2 * LT_NCOL);
  const int part = (row % (2 * LT_NCOL)) / LT_NCOL;
  const int n = row % LT_NCOL;
  const int src = (n % 4) * LT_H + cta * LT_UNITS + n / 4;
  for (int k = threadIdx.x; k < LT_H; k += blockDim.x) {
    const float w = whh[(long long)src * LT_H + k];
    const __half w1 = __float2half_rn(w);
    wp[(long long)row * LT_H + k] = part == 0 ? w1 : __float2half_rn((w - __half2float(w1)) * LT_LO_SCALE);
  }
}

size_t lt_smem_bytes(int G) {
  return 1024 + LT_STAGES * LT_STAGE_BYTES + LT_NKT * LT_W_TILE + (size_t)G * LT_ITEMS * LT_UNITS * 4 + 8 * (2 * LT_STAGES + 5) + 16;
}

}  // namespace

int launch_lstm_tc_pack(const float* whh, void* packed, int H, cudaStream_t s) {
  ECB_REQUIRE(H == LT_H, "lstm_tc: hidden size %d unsupported", H);
  lstm_tc_pack_kernel<<<LT_CTAS * 2 * LT_NCOL, 128, 0, s>>>(whh, reinterpret_cast<__half*>(packed));
  ECB_LAUNCHED();
  return 0;
}

long long* g_lstm_tc_trace = nullptr;   // diagnostic (ecb_debug_lstm_trace): device buffer of 3 * 8 * 16 stamps

bool lstm_tc_supported(int batch, int H) { return H == LT_H && batch >= 1 && batch <= LT_MAX_GROUPS * LT_ITEMS; }

// floats of workspace: h1g + h2g ([2][G * 64][512] fp16 each) + counters
int lstm_tc_workspace_floats(int batch) {
  const int G = (batch + LT_ITEMS - 1) / LT_ITEMS;
  return G * LT_ITEMS * LT_H * 2 + G * LT_NKT + 64;
}

int launch_lstm_tc(const float* pre, long long pre_item_stride, const void* w_packed, const float* skip, long long skip_item_stride,
                   float* out, long long out_item_stride, int batch, int T, int out_elu, float* workspace, cudaStream_t s) {
  ECB_REQUIRE(lstm_tc_supported(batch, LT_H) && T > 0, "lstm_tc: bad batch %d / T %d", batch, T);
  const int G = (batch + LT_ITEMS - 1) / LT_ITEMS;
  const size_t smem = lt_smem_bytes(G);
  static DeviceOnce attr_set;
  if (!attr_set.done()) {
    ECB_CUDA(cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lt_smem_bytes(LT_MAX_GROUPS)));
    attr_set.mark();
  }
  LstmTcParams p;
  p.pre = pre;
  p.pre_stride = pre_item_stride;
  p.skip = skip;
  p.skip_stride = skip_item_stride;
  p.out = out;
  p.out_stride = out_item_stride;
  const long long hbuf = 2LL * G * LT_ITEMS * LT_H;   // halves per split part
  p.h1g = reinterpret_cast<__half*>(workspace);
  p.h2g = p.h1g + hbuf;
  p.cnt = reinterpret_cast<unsigned int*>(p.h2g + hbuf);
  p.B = batch;
  p.T = T;
  p.G = G;
  p.out_elu = out_elu;
  p.trace = g_lstm_tc_trace;
  // h_{-1} = 0 (both buffers: padding rows of the last group stay finite) and the arrival counters
  ECB_CUDA(cudaMemsetAsync(workspace, 0, sizeof(__half) * 2 * hbuf + sizeof(unsigned int) * (size_t)(G * LT_NKT), s));
  CUtensorMap maps[3];
  {
    const cuuint64_t dims[2] = {(cuuint64_t)LT_H, (cuuint64_t)(2 * G * LT_ITEMS)};
    const cuuint64_t strides[1] = {(cuuint64_t)LT_H * 2};
    const cuuint32_t box[2] = {LT_KT, LT_ITEMS};
    if (make_tensor_map_f16(&maps[0], p.h1g, 2, dims, strides, box)) return 1;
    if (make_tensor_map_f16(&maps[1], p.h2g, 2, dims, strides, box)) return 1;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)LT_H, (cuuint64_t)(LT_CTAS * 2 * LT_NCOL)};
    const cuuint64_t strides[1] = {(cuuint64_t)LT_H * 2};
    const cuuint32_t box[2] = {LT_KT, 2 * LT_NCOL};
    if (make_tensor_map_f16(&maps[2], w_packed, 2, dims, strides, box)) return 1;
  }
  const double bt = (double)batch * T;
  ProfScope prof(PROF_LSTM_REC, s, 2.0 * bt * 4 * LT_H * LT_H, 4.0 * (bt * 4 * LT_H + bt * LT_H * (skip ? 2 : 1) + 4.0 * LT_H * LT_H));
  void* args[] = {(void*)&maps[0], (void*)&maps[1], (void*)&maps[2], (void*)&p};
  // all 128 CTAs spin on each other's arrivals: they must be co-resident
  ECB_CUDA(cudaLaunchCooperativeKernel((void*)lstm_tc_kernel, dim3(LT_CTAS), dim3(LT_THREADS), args, smem, s));
  ECB_LAUNCHED();
  return 0;
}

}  // namespace ecb
