// SLSTM recurrence on the tensor cores (reference modules/lstm.py:12-28 -> nn.LSTM(512, 512, 2), gates i,f,g,o, zero
// state), H = 512. Two kernels share the step described here: lstm_tc_kernel (one persistent kernel per LSTM layer; the
// input projection of all steps is a tc_conv GEMM, `pre`) and, further down, lstm_tcw_kernel (both layers as one wavefront
// kernel for launches of up to 128 items, or one layer with 16 units per CTA for large launches).
//
// Geometry: 128 CTAs = NU unit blocks x NB batch parts. A unit block owns UPC hidden units (4 UPC gate rows of W_hh); its
// [4 UPC x 512] slice sits in shared memory for the whole sequence (un-replicated inside a batch part), and the recurrent
// product of a GROUP of items is a tensor-core contraction per step,
//     rec[items][4 UPC gate columns] = h_{t-1}[items][512] . W_slice^T        (tcgen05.mma, accumulator in TMEM).
// What moves per step is h_{t-1}: every CTA publishes its UPC units of h_t to an L2-resident buffer, bumps an arrival
// counter, and pulls the complete h_t of its groups back with bulk copies. UPC = 4: 128 unit blocks, one batch part (the
// shortest chain: launches below 192 items). UPC = 8: 64 unit blocks x 2 batch parts -- an MMA of this size costs little
// more with twice the N, so the groups of a large batch are split over two sets of SMs that step independently (15-18 %
// more throughput at 512-960 items).
//
// fp32 accuracy from fp16 tensor-core operands. fp16 has the 11-bit significand of TF32 at half the bytes and twice the
// MMA rate, and h in (-1, 1) / LSTM weights fit its range. Operands are split the way tc_conv splits TF32:
//     h = h1 + 2^-11 h2,  h1 = fp16(h), h2 = fp16((h - h1) 2^11)        (written by the producer of h)
//     w = w1 + 2^-11 w2                                                 (split once at load, lstm_tc_pack)
//     rec = sum h1 w1  +  2^-11 (sum h1 w2 + sum h2 w1)                 (h2 w2 2^-22 dropped, as in the 3xTF32 scheme)
// fp16 x fp16 products are exact in the fp32 accumulator. All three products come from ONE MMA per K step of 16: the A
// tile stacks the h1 rows and the h2 rows of the items (rows 32 q + r = h1 of item 16 q + r, rows 32 q + 16 + r = its h2, so
// both land in the same TMEM lane quadrant), B = [w1 | w2]; D rows of h1 give [h1 w1 | h1 w2], D rows of h2 give
// [h2 w1 | (h2 w2, unused)], and the epilogue adds the pieces with one warp shuffle.
//
// A step is a chain of ~1000-cycle hops (publish, poll, copy, commit) around the MMAs, so the K loop is shortened by STACKING
// K tiles along M: a group of IG items has R = 2 IG operand rows, KP = 128 / R consecutive K tiles of it form one 128-row A
// tile, and B holds the KP matching weight tiles side by side. One MMA then covers KP x 16 of K; its diagonal blocks
// D[rows of tile i][columns of tile i] are the useful ones, the epilogue warps of the KP row blocks add their parts through
// shared memory. A batch of <= 64 items runs as groups of 32 (KP = 2, 16 MMAs per group-step), larger ones as groups of 64.
//
// The exchange buffers hold the SWIZZLE_128B shared-memory image of every K tile (the epilogue stores into swizzled
// positions), so the KP tiles of a stage are ONE contiguous 16 KB bulk copy: a tensor-map box of 128-byte rows at a 1 KB
// pitch moved only ~35 B per cycle into the SM and was the longest stretch of the step.
//
// Roles (12 warps): warp 0 = loader (polls the 8 per-K-tile arrival counters of a group with one coalesced acquire load,
// one cross-proxy fence, then the bulk copies into an 8-stage ring), warp 1 = MMA issuer, warp 2 = publisher (one
// release-increment per group-step once the epilogue warps have stored h_t), warps 4-11 = cell epilogue: tcgen05.ld the gate
// columns, + pre-gates, cell update (c stays in shared memory), h_t -> (h1, h2) fp16 to L2, layer output y = h (+ skip)
// (ELU) to HBM. Groups run through the same ring with two TMEM accumulator buffers, so the exchange latency of one group
// hides behind the copies and MMAs of the others.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ecb {
namespace {
using namespace tc;

constexpr int LT_H = 512;
constexpr int LT_CTAS = 128;
constexpr int LT_KT = 64;           // K elements per shared-memory tile (128 bytes of fp16)
constexpr int LT_NKT = LT_H / LT_KT;   // 8
constexpr int LT_STAGES = 8;
constexpr int LT_THREADS = 384;
constexpr int LT_STAGE_BYTES = 128 * 128;         // ring slot: one [128 rows x 64 k] A tile = KP K tiles of a group, 16 KB
constexpr int LT_TMEM_COLS = 256;                 // two accumulator buffers of up to 128 columns
constexpr int LT_MAX_GROUPS = 16;
constexpr float LT_LO_SCALE = 2048.f;             // 2^11

struct LstmTcParams {
  const float* pre;      // item b at pre + b * pre_stride, [T][4H] in reference gate order (i, f, g, o)
  long long pre_stride;
  const float* skip;     // item b at skip + b * skip_stride, [T][H], or nullptr
  long long skip_stride;
  float* out;            // item b at out + b * out_stride, [T][H]
  long long out_stride;
  __half* hg;            // [2][G][8 K tiles][2 IG rows][64] exchange buffers (L2-resident), each K tile stored as its SWIZZLE_128B
                         // shared-memory image (16-byte chunk c of row r at chunk c ^ (r & 7)). Row 32 q + 16 part + r of a group
                         // = split part `part` (0: h1, 1: h2) of item 16 q + r
  unsigned int* cnt;     // [G][8] arrival counters (one per group and K tile), zeroed by the host
  int B, T, G, out_elu;
  long long* trace;      // diagnostic: [3 roles][LT_TR_STEPS][16] clock64 stamps of CTA 0, or nullptr
};
constexpr int LT_TR_STEPS = 8, LT_TR_T0 = 20;
#define LT_TRACE(role, t, ev)                                                                                     \
  if (p.trace && cta == 0 && (t) >= LT_TR_T0 && (t) < LT_TR_T0 + LT_TR_STEPS && lane == 0)                           \
    p.trace[((role) * LT_TR_STEPS + ((t) - LT_TR_T0)) * 16 + (ev)] = clock64();

// sigmoid of two values on packed fp32 pairs (fma.rn.f32x2: two IEEE FMAs per issue slot): e^-x by range reduction
// -x = n ln2 + r and the degree-7 expm1 polynomial of elu1 (about 1 ulp), then 1 / (1 + 2^n (1 + expm1(r))) with a
// correctly rounded reciprocal. Absolute error about 1e-7; tanh(x) = 2 sigmoid(2 x) - 1 inherits it (absolute, which is
// what the cell update needs: gates and states are O(1)).
__device__ __forceinline__ void sigmoid_pair(float (&v)[2]) {
#define LT_F2C(x) f2_pack((x), (x))
  const unsigned long long y = f2_pack(fminf(fmaxf(-v[0], -87.f), 87.f), fminf(fmaxf(-v[1], -87.f), 87.f));
  const unsigned long long t = f2_fma(y, LT_F2C(1.4426950408889634f), LT_F2C(12582912.f));
  const unsigned long long n = f2_add(t, LT_F2C(-12582912.f));
  unsigned long long r = f2_fma(n, LT_F2C(-0.693145751953125f), y);
  r = f2_fma(n, LT_F2C(-1.428606765330187e-06f), r);
  unsigned long long q = f2_fma(LT_F2C(1.9841270e-4f), r, LT_F2C(1.3888889e-3f));
  q = f2_fma(q, r, LT_F2C(8.3333333e-3f));
  q = f2_fma(q, r, LT_F2C(4.1666667e-2f));
  q = f2_fma(q, r, LT_F2C(1.6666667e-1f));
  q = f2_fma(q, r, LT_F2C(0.5f));
  q = f2_fma(f2_mul(q, r), r, r);   // expm1(r)
  float t0, t1;
  f2_unpack(t, t0, t1);
  const unsigned long long s = f2_pack(__int_as_float((__float_as_int(t0) << 23) + 0x3f800000),
                                       __int_as_float((__float_as_int(t1) << 23) + 0x3f800000));   // 2^n
  const unsigned long long d = f2_fma(q, s, f2_add(s, LT_F2C(1.f)));   // 1 + e^-x
  float d0, d1;
  f2_unpack(d, d0, d1);
  v[0] = __frcp_rn(d0);
  v[1] = __frcp_rn(d1);
#undef LT_F2C
}

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int W>
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, float (&v)[W]);
template <>
__device__ __forceinline__ void tmem_ldw<8>(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ldw<16>(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  tcgen05_ld16(taddr, r);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 1-D bulk copies global -> shared (the exchange buffers are ready-made shared-memory images)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// kind::f16 instruction descriptor: D fp32 (bits [4,6) = 1), A / B fp16 (format 0), both K-major, N >> 3 in [17,23), M >> 4 in [24,29)
constexpr uint32_t idesc_f16(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

// CL = cluster size (1, 2, 4): the CTAs of a cluster need the same h tiles, so each copies 1 / CL of the rows of every tile and
// multicasts them to all. IG = items per group (64 | 32). UPC = hidden units per CTA (4 | 8).
template <int CL, int IG, int UPC>
__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap map_w, const LstmTcParams p) {
  constexpr int NCOL = 4 * UPC;                 // gate columns per CTA: n = unit * 4 + gate
  constexpr int W_TILE = 2 * NCOL * 128;        // [w1 (NCOL rows) | w2 (NCOL rows)] x 64 k
  constexpr int NU = LT_H / UPC;                // unit blocks
  constexpr int NB = LT_CTAS / NU;              // batch parts: groups g = part, part + NB, ... belong to a part
  constexpr int ROWS = 2 * IG;                  // operand rows per group: h1 and h2 row of every item
  constexpr int KP = 128 / ROWS;                // K tiles stacked along M in one MMA (1 | 2)
  constexpr int NACC = KP == 1 ? 2 : 1;         // accumulators per buffer (<= 16 accumulating MMAs each: the tensor core truncates)
  constexpr int NMMA = 2 * NCOL * KP;           // N of one MMA
  constexpr int ACC_COLS = NACC * NMMA;         // TMEM columns per accumulator buffer
  constexpr int QPG = IG / 16;                  // TMEM lane quadrants per K part
  constexpr int UW = UPC / 2;                   // units per epilogue warp (2 | 4)
  constexpr int PP = UW / 2;                    // (item, unit) pairs per epilogue lane (1 | 2)
  constexpr int CW = 4 * UW;                    // gate columns per epilogue warp (8 | 16)
  static_assert(2 * ACC_COLS <= LT_TMEM_COLS, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;             // SWIZZLE_128B tiles: 1024-byte aligned
  uint8_t* smem_gen = smem_raw + (base - raw_addr);
  const uint32_t a_ring = base;                                   // [LT_STAGES][128 rows x 128 B]
  const uint32_t w_smem = base + LT_STAGES * LT_STAGE_BYTES;      // [LT_NKT][W_TILE]
  constexpr int CS_OFF = LT_STAGES * LT_STAGE_BYTES + LT_NKT * W_TILE;
  const int cta = blockIdx.x;
  const int ub = cta % NU;                       // unit block
  const int bp = cta / NU;                       // batch part
  const int n_local = p.G > bp ? (p.G - bp + NB - 1) / NB : 0;   // groups of this batch part
  float* cs = reinterpret_cast<float*>(smem_gen + CS_OFF);        // [n_local][IG items][UPC units] cell state
  const int cs_bytes = ((p.G + NB - 1) / NB) * IG * UPC * 4;
  constexpr int XS_BYTES = 4 * 16 * CW * 4;
  float* xs = reinterpret_cast<float*>(smem_gen + CS_OFF + cs_bytes);   // [teams][16][CW] K-part exchange between epilogue warps (KP = 2)
  const uint32_t bar_base = base + CS_OFF + cs_bytes + XS_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (LT_STAGES + s); };
  auto accf_bar = [&](int b) { return bar_base + 8u * (2 * LT_STAGES + b); };
  auto acce_bar = [&](int b) { return bar_base + 8u * (2 * LT_STAGES + 2 + b); };
  const uint32_t w_bar = bar_base + 8u * (2 * LT_STAGES + 4);
  auto hst_bar = [&](int b) { return bar_base + 8u * (2 * LT_STAGES + 5 + b); };   // the epilogue warps have stored h_t
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + CS_OFF + cs_bytes + XS_BYTES + 8 * (2 * LT_STAGES + 7));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t crank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1u);
  constexpr int ROWS_PER_CTA = ROWS / CL;       // rows of every h tile this CTA fetches for the cluster

  if (threadIdx.x == 0) {
    for (int s = 0; s < LT_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CL);   // every CTA of the cluster writes into this stage: all their MMA warps release it
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf_bar(b), 1);
      mbar_init(acce_bar(b), 8);        // one arrive per epilogue warp
      mbar_init(hst_bar(b), 2 * QPG);   // the warps that store h
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < n_local * IG * UPC; i += LT_THREADS) cs[i] = 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(LT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // barriers of every CTA are initialised before any remote arrive / multicast write
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const long long buf_bytes = (long long)p.G * LT_NKT * ROWS * 128;   // one exchange buffer
  // When all of this CTA's groups' stages fit the ring at once, local group gl owns stages gl * 8 / KP ..: a stage is then
  // known to be free when its counters say that h_t is complete (this CTA's own epilogue of step t - 1 came after its
  // MMAs), and the empty barriers -- ~150 cycles per stage on the step's critical path -- are not used at all.
  const bool private_ring = n_local * (LT_NKT / KP) <= LT_STAGES;

  if (warp == 0) {
    // ================================ loader ================================
    if (elect_one()) {   // this CTA's weight slice, once: 8 tiles [2 NCOL rows x 64 k]
      mbar_expect_tx(w_bar, LT_NKT * W_TILE);
      for (int j = 0; j < LT_NKT; ++j) tma_load_2d(w_smem + j * W_TILE, &map_w, w_bar, j * LT_KT, ub * 2 * NCOL);
    }
    __syncwarp();
    constexpr unsigned int STAGE_MASK = (1u << KP) - 1u;
    uint32_t it = 0;
    for (int t = 0; t < p.T; ++t) {
      const unsigned int target = (unsigned int)(LT_KT / UPC) * (unsigned int)t;   // the CTAs that publish one K tile of 64 units
      for (int gl = 0; gl < n_local; ++gl) {
        const int g = bp + gl * NB;
        const uint8_t* src_g = reinterpret_cast<const uint8_t*>(p.hg) + (long long)(t & 1) * buf_bytes + (long long)g * (LT_NKT * ROWS * 128);
        int next = 0;   // next stage (KP K tiles) of this group-step
        unsigned int spins = 0;
        if (gl == 0) LT_TRACE(0, t, 0)
        while (next < LT_NKT / KP) {
          unsigned int v = 0xffffffffu;
          if (t > 0 && lane < LT_NKT)
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.cnt + g * LT_NKT + lane) : "memory");
          const unsigned int ready = __ballot_sync(0xffffffffu, v >= target);
          if (((ready >> (next * KP)) & STAGE_MASK) != STAGE_MASK) {
            if (++spins > (1u << 22)) __trap();   // a lost arrival must not hang the device
            continue;
          }
          // h was written through the generic proxy (by other SMs) and is read through the async proxy: one fence per poll
          if (gl == 0 && next == 0) LT_TRACE(0, t, 9)
          if (t > 0) asm volatile("fence.proxy.async.global;" ::: "memory");
          if (gl == 0 && next == 0) LT_TRACE(0, t, 10)
          while (next < LT_NKT / KP && ((ready >> (next * KP)) & STAGE_MASK) == STAGE_MASK) {
            const int st = private_ring ? gl * (LT_NKT / KP) + next : (int)(it % LT_STAGES);
            if (!private_ring) mbar_wait(empty_bar(st), ((it / LT_STAGES) & 1u) ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(full_bar(st), LT_STAGE_BYTES);
              const uint32_t dst = a_ring + st * LT_STAGE_BYTES;
              const uint8_t* src = src_g + (long long)next * LT_STAGE_BYTES;   // the KP tiles of a stage are contiguous
              if (CL == 1) {
                bulk_load(dst, src, LT_STAGE_BYTES, full_bar(st));
              } else {   // this CTA's row slice of every tile, delivered to the same offset in every CTA of the cluster
#pragma unroll
                for (int sub = 0; sub < KP; ++sub) {
                  const uint32_t o = (uint32_t)(sub * ROWS + (int)crank * ROWS_PER_CTA) * 128u;
                  bulk_load_mc(dst + o, src + o, ROWS_PER_CTA * 128, full_bar(st), MC_MASK);
                }
              }
            }
            __syncwarp();
            if (gl == 0) LT_TRACE(0, t, 1 + next)
            ++next;
            ++it;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = idesc_f16(128, NMMA);
    constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
    auto mk_desc = [](uint32_t addr) { return ((uint64_t)DESC_HI << 32) | (uint64_t)(((addr & 0x3FFFFu) >> 4) | (1u << 16)); };
    mbar_wait(w_bar, 0);
    uint32_t it = 0, n = 0;
    for (int t = 0; t < p.T; ++t) {
      for (int gl = 0; gl < n_local; ++gl, ++n) {
        const int acc = (int)(n & 1u);
        mbar_wait(acce_bar(acc), ((n >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator buffer
        for (int j = 0; j < LT_NKT / KP; ++j, ++it) {
          const int st = private_ring ? gl * (LT_NKT / KP) + j : (int)(it % LT_STAGES);
          mbar_wait(full_bar(st), private_ring ? ((uint32_t)t & 1u) : ((it / LT_STAGES) & 1u));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (gl == 0) LT_TRACE(1, t, j)
          if (elect_one()) {
            const uint32_t d = tmem_base + (uint32_t)(acc * ACC_COLS);
            const uint64_t da = mk_desc(a_ring + st * LT_STAGE_BYTES);
            const uint64_t db = mk_desc(w_smem + j * KP * W_TILE);   // KP weight tiles side by side: N = 2 NCOL KP rows
#pragma unroll
            for (int k = 0; k < LT_KT / 16; ++k) {   // rows of h1: [main | corr] (+)= h1 [w1 | w2]; rows of h2: [corr | -] (+)= h2 [w1 | w2]
              const int mm = j * (LT_KT / 16) + k;
              mma_f16(d + (uint32_t)((mm % NACC) * NMMA), da + 2u * k, db + 2u * k, idesc, mm >= NACC ? 1u : 0u);
            }
            if (!private_ring) {
              if (CL == 1) tcgen05_commit(empty_bar(st));
              else commit_mc(empty_bar(st), MC_MASK);
            }
            if (j + 1 == LT_NKT / KP) tcgen05_commit(accf_bar(acc));
          }
          __syncwarp();
          if (gl == 0 && j + 1 == LT_NKT / KP) LT_TRACE(1, t, 8)
        }
      }
    }
  } else if (warp == 2) {
    // ================================ publisher ================================
    // Once the storing epilogue warps have written their part of h_t (mbarrier: release.cta arrive / acquire.cta wait), ONE
    // release-increment at gpu scope publishes it: the release is cumulative over the stores observed through the barrier.
    // A separate warp, so that the epilogue warps go on to the layer-output stores and the next group instead of sitting
    // in the ~1000-cycle fence.
    uint32_t n = 0;
    for (int t = 0; t < p.T; ++t) {
      for (int gl = 0; gl < n_local; ++gl, ++n) {
        const int g = bp + gl * NB;
        mbar_wait(hst_bar((int)(n & 1u)), (n >> 1) & 1u);
        if (t + 1 < p.T && lane == 0)
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.cnt + g * LT_NKT + (ub * UPC) / LT_KT) : "memory");
        __syncwarp();
        if (gl == 0) LT_TRACE(2, t, 6)
      }
    }
  } else if (warp >= 4) {
    // ================================ cell epilogue ================================
    // Operand row block b (16 rows) of a group = split part b % 2 (0: h1, 1: h2) of items 16 (b / 2) .. + 15, and K tile i of a
    // stage sits in MMA rows i R .. (accumulator row m = TMEM lane m). A lane quadrant therefore holds, for one K part and 16
    // items, the h1 rows in lanes 0-15 and the h2 rows in lanes 16-31 (one shuffle adds them); the KP quadrants of the same
    // items add their K parts through shared memory. Warp w reads quadrant w % 4 (hardware rule); warps 4-7 take the first
    // UW units of the CTA, warps 8-11 the other UW. In the end lane r of a storing warp finishes its first PP units of
    // item r and lane 16 + r the other PP.
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = lane & 15;
    const int lg = lane >> 4;
    const int kpart = quad / QPG;                 // which K tile of a stage this quadrant's rows belong to
    const int iq = quad % QPG;                    // 16-item block of the group
    const bool storing = kpart == 0;              // this warp finishes items; the others hand their K part over
    const int il = iq * 16 + r;                   // item within the group
    const int team = half * QPG + iq;             // the KP warps that share (units, items)
    const int u0 = half * UW + lg * PP;           // this lane's first unit within the CTA
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(2 * NCOL * kpart + CW * half);
    uint32_t n = 0;
    for (int t = 0; t < p.T; ++t) {
      uint8_t* hn = reinterpret_cast<uint8_t*>(p.hg) + (long long)((t + 1) & 1) * buf_bytes;
      for (int gl = 0; gl < n_local; ++gl, ++n) {
        const int g = bp + gl * NB;
        const int acc = (int)(n & 1u);
        const int item = g * IG + il;
        const bool valid = storing && item < p.B;
        // pre-gates / skip input of this step: independent of the recurrence, in flight while we wait for the MMAs
        float pg[PP][4], skipv[PP];
#pragma unroll
        for (int e = 0; e < PP; ++e) {
          skipv[e] = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) pg[e][k] = 0.f;
          if (valid) {
            const int unit = ub * UPC + u0 + e;
            const float* pr = p.pre + (long long)item * p.pre_stride + (long long)t * (4 * LT_H) + unit;
#pragma unroll
            for (int k = 0; k < 4; ++k) pg[e][k] = __ldg(pr + k * LT_H);
            if (p.skip) skipv[e] = __ldg(p.skip + (long long)item * p.skip_stride + (long long)t * LT_H + unit);
          }
        }
        if (gl == 0 && warp == 4) LT_TRACE(2, t, 0)
        mbar_wait(accf_bar(acc), (n >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (gl == 0 && warp == 4) LT_TRACE(2, t, 1)
        float mm[NACC][CW], cc[NACC][CW];
        const uint32_t a0 = lane_base + (uint32_t)(acc * ACC_COLS);
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          tmem_ldw<CW>(a0 + NMMA * a, mm[a]);          // main columns of this K part (h1 rows: h1 w1, h2 rows: h2 w1)
          tmem_ldw<CW>(a0 + NMMA * a + NCOL, cc[a]);   // correction columns (h1 rows: h1 w2)
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(acce_bar(acc));
        if (gl == 0 && warp == 4) LT_TRACE(2, t, 2)
        // h1 rows (lanes 0-15): main + corr / 2^11 with corr = h1 w2; h2 rows (lanes 16-31): their "main" columns are h2 w1
        float part[CW];
#pragma unroll
        for (int k = 0; k < CW; ++k) {
          float m = mm[0][k], c = cc[0][k];
          if (NACC == 2) {
            m += mm[NACC - 1][k];
            c += cc[NACC - 1][k];
          }
          part[k] = lane < 16 ? m + c * (1.f / LT_LO_SCALE) : m * (1.f / LT_LO_SCALE);
        }
#pragma unroll
        for (int k = 0; k < CW; ++k) part[k] += __shfl_xor_sync(0xffffffffu, part[k], 16);
        if (KP > 1) {   // add the K parts of the other quadrants (same units, same items)
          float* x = xs + (team * 16 + r) * CW;
          if (!storing && lane < 16) {
#pragma unroll
            for (int k = 0; k < CW; k += 4) *reinterpret_cast<float4*>(x + k) = make_float4(part[k], part[k + 1], part[k + 2], part[k + 3]);
          }
          asm volatile("bar.sync %0, %1;" ::"r"(2 + team), "r"(32 * KP) : "memory");   // the team's K parts are in shared memory
          if (storing) {
#pragma unroll
            for (int k = 0; k < CW; k += 4) {
              const float4 xv = *reinterpret_cast<const float4*>(x + k);
              part[k] += xv.x; part[k + 1] += xv.y; part[k + 2] += xv.z; part[k + 3] += xv.w;
            }
          }
          asm volatile("bar.sync %0, %1;" ::"r"(2 + team), "r"(32 * KP) : "memory");   // ... and read: the slot may be rewritten
          if (!storing) continue;
        }
        if (gl == 0 && warp == 4) LT_TRACE(2, t, 4)
        unsigned int own[PP];
        float h_new[PP];
#pragma unroll
        for (int e = 0; e < PP; ++e) {
          float rec[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // columns of unit lg PP + e of this warp (compile-time indices: both candidates, then select)
            const float lo = part[e * 4 + k], hi = part[(PP + e) * 4 + k];
            rec[k] = lane < 16 ? lo : hi;
          }
          // gates i, f, o = sigmoid, g = tanh = 2 sigmoid(2 x) - 1: four sigmoids on two packed pairs
          float s_if[2] = {pg[e][0] + rec[0], pg[e][1] + rec[1]};
          float s_go[2] = {2.f * (pg[e][2] + rec[2]), pg[e][3] + rec[3]};
          sigmoid_pair(s_if);
          sigmoid_pair(s_go);
          float* cptr = cs + (gl * IG + il) * UPC + u0 + e;
          const float c_new = s_if[1] * (*cptr) + s_if[0] * (2.f * s_go[0] - 1.f);
          *cptr = c_new;
          float s_c[2] = {2.f * c_new, 0.f};
          sigmoid_pair(s_c);
          h_new[e] = s_go[1] * (2.f * s_c[0] - 1.f);
          const __half q1 = __float2half_rn(h_new[e]);
          const __half q2 = __float2half_rn((h_new[e] - __half2float(q1)) * LT_LO_SCALE);
          own[e] = (unsigned int)__half_as_ushort(q1) | ((unsigned int)__half_as_ushort(q2) << 16);
        }
        if (gl == 0 && warp == 4) LT_TRACE(2, t, 5)
        unsigned int oth[PP];
#pragma unroll
        for (int e = 0; e < PP; ++e) oth[e] = __shfl_down_sync(0xffffffffu, own[e], 16);
        if (lane < 16 && t + 1 < p.T) {   // the UW units of this item: one store per split part
          // shared-memory image of K tile (UPC ub) / 64: row (g, iq, r), 16-byte chunk (unit % 64) / 8 swizzled by r & 7
          const int kk = (ub * UPC + half * UW) % LT_KT;
          const long long row = ((long long)g * LT_NKT + (ub * UPC) / LT_KT) * ROWS + iq * 32 + r;
          uint8_t* dst = hn + row * 128 + (((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2;
          if (PP == 1) {
            const unsigned int v1 = (own[0] & 0xffffu) | (oth[0] << 16);
            const unsigned int v2 = (own[0] >> 16) | (oth[0] & 0xffff0000u);
            asm volatile("st.global.cg.b32 [%0], %1;" ::"l"(dst), "r"(v1) : "memory");                // h1 row
            asm volatile("st.global.cg.b32 [%0], %1;" ::"l"(dst + 16 * 128), "r"(v2) : "memory");     // h2 row (same r & 7)
          } else {
            const unsigned int a1 = (own[0] & 0xffffu) | (own[PP - 1] << 16), b1 = (oth[0] & 0xffffu) | (oth[PP - 1] << 16);
            const unsigned int a2 = (own[0] >> 16) | (own[PP - 1] & 0xffff0000u), b2 = (oth[0] >> 16) | (oth[PP - 1] & 0xffff0000u);
            asm volatile("st.global.cg.v2.b32 [%0], {%1, %2};" ::"l"(dst), "r"(a1), "r"(b1) : "memory");
            asm volatile("st.global.cg.v2.b32 [%0], {%1, %2};" ::"l"(dst + 16 * 128), "r"(a2), "r"(b2) : "memory");
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(hst_bar(acc));   // -> publisher warp
        if (gl == 0 && warp == 4) LT_TRACE(2, t, 3)
        // the layer output is nobody's business inside this kernel: stored after h_t is on its way
        if (valid) {
#pragma unroll
          for (int e = 0; e < PP; ++e) {
            float y = h_new[e] + skipv[e];
            if (p.out_elu) y = elu1(y);
            p.out[(long long)item * p.out_stride + (long long)t * LT_H + ub * UPC + u0 + e] = y;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // nobody leaves while a neighbour may still multicast into it / arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(LT_TMEM_COLS) : "memory");
  }
}


// ====================================================================================================
// Second form: `lstm_tcw_kernel<CL, IG, UPC, TWO>` -- the same step (exchange through L2, one MMA per K step, fp16 split
// operands), with the shared-memory budget spent differently:
//
// TWO = true, UPC = 8 (launches of up to 64 / 128 items): BOTH layers of the SLSTM in one kernel, as a wavefront. CTAs 0-63
// run layer 1, CTAs 64-127 layer 2; layer 2 is one step behind. What bounds the single-layer kernel is the chain publish ->
// poll -> copy -> MMA -> cell of one step (~6500 cycles at 64 items), so running step t of layer 2 beside step t + 1 of layer 1
// halves the number of sequential hops of the block (4 x T -> 2 x (T + 1) for encoder + decoder). A layer-2 CTA holds its
// slices of W_ih AND W_hh (128 KB): the input projection of layer 2 is not hoisted any more, it is the "x part" of the step,
//     gates_2(t) = b + y1_t W_ih^T (x part, accumulator X) + h2_{t-1} W_hh^T (h part, accumulator H),
// and y1_t = h1_t is read from the very exchange buffer layer 1 publishes for its own next step -- the layer-1 output never
// goes to HBM, nor do the layer-2 pre-gates (2 x 393 MB written and read per config-2 layer before). The x part only needs
// h1_t, which exists a whole step before h2_{t-1}: its copies and MMAs run while the layer-2 chain waits for the exchange.
// h1 lives in a ring of 4 slots so that layer 1 may run ahead; it re-uses a slot only when layer 2's arrival counters show
// that every layer-2 CTA has consumed it (flow control: one more coalesced acquire load in the poll).
//
// TWO = false, UPC = 16 (large launches): 32 unit blocks x 4 batch parts. Every CTA of a unit block has to pull the complete
// h_t of its items into its SM (measured: 32-57 B / cycle / SM, the bound of the UPC = 8 geometry from 256 items: 512 KB per
// CTA and step at 512 items); 16 units per CTA halve that volume per step, and the MMAs get N = 128.
constexpr int LW_CS_OFF = 5 * LT_STAGE_BYTES + 16 * 8192;   // 208 KB: ring + weights of the largest layout
constexpr int LW_RING1 = 4;                                  // exchange slots of h1 (TWO)
constexpr int LW_TMEM_COLS = 512;

struct LstmTcwParams {
  const float* pre;      // (first layer) item b at pre + b * pre_stride, [T][4H] in reference gate order
  long long pre_stride;
  const float* bias2;    // TWO: [4H] b_ih + b_hh of the second layer (reference gate order)
  const float* skip;     // item b at skip + b * skip_stride, [T][H], or nullptr
  long long skip_stride;
  float* out;            // output of the last layer
  long long out_stride;
  __half* h1;            // [TWO ? LW_RING1 : 2][G][8 K tiles][2 IG rows][64]: swizzled K-tile images as in LstmTcParams::hg
  __half* h2;            // TWO: [2][G][8][2 IG][64]
  unsigned int* cnt1;    // [G][8] arrival counters of h1
  unsigned int* cnt2;    // TWO: [G][8] arrival counters of h2
  int B, T, G, out_elu;
  long long* trace;
  int trace_cta;         // diagnostic: the CTA whose stamps are recorded (ECB_LSTM_TRACE_CTA; default: first CTA of the last layer)
};
#define LW_TRACE(role, t, ev)                                                                                     \
  if (p.trace && cta == trace_cta && (t) >= LT_TR_T0 && (t) < LT_TR_T0 + LT_TR_STEPS && lane == 0)                 \
    p.trace[((role) * LT_TR_STEPS + ((t) - LT_TR_T0)) * 16 + (ev)] = clock64();

// mbarrier wait that cannot hang the device: a lost arrival traps after ~2^31 cycles
__device__ __forceinline__ void mbar_wait_g(uint32_t bar, uint32_t parity) {
  if (mbar_test(bar, parity)) return;
  const long long t0 = clock64();
  unsigned int n = 0;
  while (!mbar_test(bar, parity)) {
    if ((++n & 0xfffu) == 0 && clock64() - t0 > (1LL << 31)) __trap();
  }
}

template <int W>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {   // W = 8 | 16 | 32 columns of this warp's lanes
  if (W == 8) {
    float t[8];
    tmem_ldw<8>(taddr, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = t[i];
  } else {
#pragma unroll
    for (int c = 0; c < W; c += 16) {
      float t[16];
      tmem_ldw<16>(taddr + (uint32_t)c, t);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[c + i] = t[i];
    }
  }
}

template <int CL, int IG, int UPC, bool TWO>
__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_tcw_kernel(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2x,
                const __grid_constant__ CUtensorMap map_w2h, const LstmTcwParams p) {
  constexpr int NCOL = 4 * UPC;                 // gate columns per CTA: n = unit * 4 + gate
  constexpr int W_TILE = 2 * NCOL * 128;        // [w1 (NCOL rows) | w2 (NCOL rows)] x 64 k
  constexpr int NU = LT_H / UPC;                // unit blocks (= CTAs of a layer when TWO)
  constexpr int NB = TWO ? 1 : LT_CTAS / NU;    // batch parts
  constexpr int ROWS = 2 * IG;
  constexpr int KP = 128 / ROWS;                // K tiles stacked along M in one MMA (1 | 2)
  constexpr int NST = LT_NKT / KP;              // ring stages per group-step and operand part
  constexpr int NACC = KP == 1 ? 2 : 1;
  constexpr int NMMA = 2 * NCOL * KP;
  constexpr int ACC_COLS = NACC * NMMA;
  constexpr int BUF_COLS = TWO ? 2 * ACC_COLS : ACC_COLS;   // accumulator buffer: [H part | X part]
  constexpr int QPG = IG / 16;
  constexpr int UW = UPC / 2;                   // units per epilogue warp
  constexpr int PP = UW / 2;                    // (item, unit) pairs per epilogue lane
  constexpr int CW = 4 * UW;                    // gate columns per epilogue warp
  static_assert(2 * BUF_COLS <= LW_TMEM_COLS, "TMEM budget");
  static_assert(!TWO || UPC == 8, "the two-layer form is built for 64 + 64 CTAs");
  static_assert(NU % CL == 0, "clusters must not straddle batch parts");
  static_assert(PP == 2 || PP == 4, "8 or 16 units per CTA");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (base - raw_addr);
  const int cta = blockIdx.x;
  const bool second = TWO && cta >= NU;          // this CTA runs layer 2
  const int ub = TWO ? cta % NU : cta % NU;
  const int bp = TWO ? 0 : cta / NU;
  const int n_local = p.G > bp ? (p.G - bp + NB - 1) / NB : 0;
  const int n_wt = second ? 2 * LT_NKT : LT_NKT;                       // weight tiles
  const int n_stages = (n_wt * W_TILE + 8 * LT_STAGE_BYTES <= LW_CS_OFF) ? 8 : 5;
  const uint32_t a_ring = base;
  const uint32_t w_smem = base + (uint32_t)(n_stages * LT_STAGE_BYTES);
  float* cs = reinterpret_cast<float*>(smem_gen + LW_CS_OFF);          // [n_local][IG][UPC] cell state
  const int cs_bytes = ((p.G + NB - 1) / NB) * IG * UPC * 4;
  constexpr int XS_BYTES = KP > 1 ? 4 * 2 * 16 * (CW / 2) * 4 : 0;   // [teams][K part][16 items][columns of the partner's units]
  float* xs = reinterpret_cast<float*>(smem_gen + LW_CS_OFF + cs_bytes);
  const uint32_t bar_base = base + LW_CS_OFF + cs_bytes + XS_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (LT_STAGES + s); };
  auto accf_bar = [&](int b) { return bar_base + 8u * (2 * LT_STAGES + b); };
  auto acce_bar = [&](int b) { return bar_base + 8u * (2 * LT_STAGES + 2 + b); };
  const uint32_t w_bar = bar_base + 8u * (2 * LT_STAGES + 4);
  auto hst_bar = [&](int b) { return bar_base + 8u * (2 * LT_STAGES + 5 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + LW_CS_OFF + cs_bytes + XS_BYTES + 8 * (2 * LT_STAGES + 7));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int trace_cta = p.trace_cta;
  uint32_t crank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1u);
  constexpr int ROWS_PER_CTA = ROWS / CL;

  if (threadIdx.x == 0) {
    for (int s = 0; s < LT_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CL);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf_bar(b), 1);
      mbar_init(acce_bar(b), 8);
      mbar_init(hst_bar(b), 8);          // every epilogue warp stores a part of h
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < n_local * IG * UPC; i += LT_THREADS) cs[i] = 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(LW_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const long long grp_bytes = (long long)LT_NKT * ROWS * 128;        // one group's h: 8 K-tile images
  const long long buf_bytes = (long long)p.G * grp_bytes;            // one exchange slot
  // operand parts of a step, in issue order: layer 2 first takes its x part (h1_t, ready a step early), then the h part
  const int part0 = second ? 0 : 1;   // 0: x part, 1: h part
  // private ring (no empty barriers): every (group, stage) of a step has its own slot. Only without an x part.
  const bool private_ring = !second && n_local * NST <= n_stages;
  // exchange slot that holds h_{t-1} of THIS layer for step t / that h1_t is written to
  auto slot_h = [&](int t) { return second ? (t & 1) : (TWO ? t % LW_RING1 : (t & 1)); };

  if (warp == 0) {
    // ================================ loader ================================
    if (elect_one()) {
      mbar_expect_tx(w_bar, (uint32_t)(n_wt * W_TILE));
      if (!second) {
        for (int j = 0; j < LT_NKT; ++j) tma_load_2d(w_smem + j * W_TILE, &map_w1, w_bar, j * LT_KT, ub * 2 * NCOL);
      } else {   // tiles 0-7: W_ih of layer 2 (x part), tiles 8-15: its W_hh
        for (int j = 0; j < LT_NKT; ++j) tma_load_2d(w_smem + j * W_TILE, &map_w2x, w_bar, j * LT_KT, ub * 2 * NCOL);
        for (int j = 0; j < LT_NKT; ++j) tma_load_2d(w_smem + (LT_NKT + j) * W_TILE, &map_w2h, w_bar, j * LT_KT, ub * 2 * NCOL);
      }
    }
    __syncwarp();
    constexpr unsigned int STAGE_MASK = (1u << KP) - 1u;
    constexpr unsigned int PER_TILE = LT_KT / UPC;   // CTAs that publish one K tile
    int ring_st = 0;            // shared ring: next stage and its phase, stepped without divisions (they cost ~100 cycles each)
    uint32_t ring_ph = 0;
    for (int t = 0; t < p.T; ++t) {
      for (int part = part0; part < 2; ++part) {
        // x part: h1_t (slot of layer-1 step t + 1), complete when cnt1 = PER_TILE (t + 1); h part: h_{t-1} of this layer
        const __half* src_buf = part == 0 ? p.h1 : (second ? p.h2 : p.h1);
        const int slot = part == 0 ? (t + 1) % LW_RING1 : slot_h(t);
        const unsigned int* cnt = (part == 0 || !second) ? p.cnt1 : p.cnt2;
        const unsigned int target = PER_TILE * (unsigned int)(part == 0 ? t + 1 : t);
        const bool need_poll = part == 0 || t > 0;
        // flow control (layer 1 of TWO): h1_t goes to the slot that held h1_{t - RING}: every layer-2 CTA must have published
        // h2_{t - RING} (its MMAs of that step have read the slot)
        const bool flow = TWO && !second && t >= LW_RING1;
        const unsigned int flow_target = PER_TILE * (unsigned int)(t - LW_RING1 + 1);
        for (int gl = 0; gl < n_local; ++gl) {
          const int g = bp + gl * NB;
          const uint8_t* src_g = reinterpret_cast<const uint8_t*>(src_buf) + (long long)slot * buf_bytes + (long long)g * grp_bytes;
          int next = 0;
          bool flow_ok = !flow;
          long long t_spin = 0;
          unsigned int spins = 0;
          if (gl == 0 && part == 1) LW_TRACE(0, t, 0)
          while (next < NST) {
            unsigned int v = 0xffffffffu, tgt = 0;
            if (need_poll && lane < LT_NKT) {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt + g * LT_NKT + lane) : "memory");
              tgt = target;
            } else if (!flow_ok && lane >= 8 && lane < 8 + LT_NKT) {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.cnt2 + g * LT_NKT + (lane - 8)) : "memory");
              tgt = flow_target;
            }
            const unsigned int ready = __ballot_sync(0xffffffffu, v >= tgt);
            if (!flow_ok && ((ready >> 8) & 0xffu) == 0xffu) flow_ok = true;
            if (!flow_ok || ((ready >> (next * KP)) & STAGE_MASK) != STAGE_MASK) {
              if ((++spins & 0x3ffu) == 0) {   // a lost arrival must not hang the device
                if (t_spin == 0) t_spin = clock64();
                else if (clock64() - t_spin > (1LL << 31)) __trap();
              }
              continue;
            }
            if (gl == 0 && next == 0 && part == 1) LW_TRACE(0, t, 9)
            if (need_poll) asm volatile("fence.proxy.async.global;" ::: "memory");
            if (gl == 0 && next == 0 && part == 1) LW_TRACE(0, t, 10)
            while (next < NST && ((ready >> (next * KP)) & STAGE_MASK) == STAGE_MASK) {
              const int st = private_ring ? gl * NST + next : ring_st;
              if (!private_ring) mbar_wait_g(empty_bar(st), ring_ph ^ 1u);
              if (elect_one()) {
                mbar_expect_tx(full_bar(st), LT_STAGE_BYTES);
                const uint32_t dst = a_ring + st * LT_STAGE_BYTES;
                const uint8_t* src = src_g + (long long)next * LT_STAGE_BYTES;
                if (CL == 1) {
                  bulk_load(dst, src, LT_STAGE_BYTES, full_bar(st));
                } else {
#pragma unroll
                  for (int sub = 0; sub < KP; ++sub) {
                    const uint32_t o = (uint32_t)(sub * ROWS + (int)crank * ROWS_PER_CTA) * 128u;
                    bulk_load_mc(dst + o, src + o, ROWS_PER_CTA * 128, full_bar(st), MC_MASK);
                  }
                }
              }
              __syncwarp();
              if (gl == 0 && part == 1) LW_TRACE(0, t, 1 + next)
              ++next;
              if (++ring_st == n_stages) {
                ring_st = 0;
                ring_ph ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = idesc_f16(128, NMMA);
    constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);
    auto mk_desc = [](uint32_t addr) { return ((uint64_t)DESC_HI << 32) | (uint64_t)(((addr & 0x3FFFFu) >> 4) | (1u << 16)); };
    mbar_wait_g(w_bar, 0);
    int ring_st = 0;
    uint32_t ring_ph = 0;
    for (int t = 0; t < p.T; ++t) {
      for (int part = part0; part < 2; ++part) {
        for (int gl = 0; gl < n_local; ++gl) {
          const uint32_t n = (uint32_t)t * (uint32_t)n_local + (uint32_t)gl;
          const int acc = (int)(n & 1u);
          if (part == part0) mbar_wait_g(acce_bar(acc), ((n >> 1) & 1u) ^ 1u);   // the epilogue has drained this buffer (both parts)
          const uint32_t d = tmem_base + (uint32_t)(acc * BUF_COLS + (part == 0 ? ACC_COLS : 0));
          const int wt0 = (second && part == 1) ? LT_NKT : 0;
          for (int j = 0; j < NST; ++j) {
            const int st = private_ring ? gl * NST + j : ring_st;
            mbar_wait_g(full_bar(st), private_ring ? ((uint32_t)t & 1u) : ring_ph);
            if (++ring_st == n_stages) {
              ring_st = 0;
              ring_ph ^= 1u;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (gl == 0 && part == 1) LW_TRACE(1, t, j)
            if (elect_one()) {
              const uint64_t da = mk_desc(a_ring + st * LT_STAGE_BYTES);
              const uint64_t db = mk_desc(w_smem + (uint32_t)((wt0 + j * KP) * W_TILE));
#pragma unroll
              for (int k = 0; k < LT_KT / 16; ++k) {
                const int mm = j * (LT_KT / 16) + k;
                mma_f16(d + (uint32_t)((mm % NACC) * NMMA), da + 2u * k, db + 2u * k, idesc, mm >= NACC ? 1u : 0u);
              }
              if (!private_ring) {
                if (CL == 1) tcgen05_commit(empty_bar(st));
                else commit_mc(empty_bar(st), MC_MASK);
              }
              if (part == 1 && j + 1 == NST) tcgen05_commit(accf_bar(acc));
            }
            __syncwarp();
            if (gl == 0 && part == 1 && j + 1 == NST) LW_TRACE(1, t, 8)
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================ publisher ================================
    // layer 1 of TWO publishes every step (layer 2 still needs h1_{T-1}); otherwise the last h is nobody's operand
    unsigned int* cnt = second ? p.cnt2 : p.cnt1;
    uint32_t n = 0;
    for (int t = 0; t < p.T; ++t) {
      for (int gl = 0; gl < n_local; ++gl, ++n) {
        const int g = bp + gl * NB;
        mbar_wait_g(hst_bar((int)(n & 1u)), (n >> 1) & 1u);
        if ((t + 1 < p.T || (TWO && !second)) && lane == 0)
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt + g * LT_NKT + (ub * UPC) / LT_KT) : "memory");
        __syncwarp();
        if (gl == 0) LW_TRACE(2, t, 6)
      }
    }
  } else if (warp >= 4) {
    // ================================ cell epilogue ================================
    // Accumulator layout as in lstm_tc_kernel: a lane quadrant holds, for one K part and 16 items, the h1 rows in lanes 0-15 and
    // the h2 rows in lanes 16-31; warp w reads quadrant w % 4, warps 4-7 the first UW units of the CTA, warps 8-11 the others.
    // KP == 2: the two warps of a team (same units and items, one K part each) hand each other the partial sums of HALF of
    // the team's units through shared memory and both go on to the cell update -- every lane finishes ONE (item, unit) pair
    // (PPL = PP / 2) instead of the K-part-0 warp finishing two while its partner idles: the cell math (three dependent
    // sigmoid evaluations) is the longest stretch of the epilogue.
    constexpr int PPL = KP == 2 ? PP / 2 : PP;    // (item, unit) pairs per lane
    constexpr int UWW = KP == 2 ? UW / 2 : UW;    // units finished by one warp
    static_assert(KP == 1 || PP == 2, "the shared epilogue is written for 8 units per CTA");
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = lane & 15;
    const int lg = lane >> 4;
    const int kpart = quad / QPG;
    const int iq = quad % QPG;
    const int il = iq * 16 + r;
    const int team = half * QPG + iq;
    const int uw0 = half * UW + (KP == 2 ? kpart * UWW : 0);   // first unit (within the CTA) this warp finishes
    const int u0 = uw0 + lg * PPL;                             // ... and this lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(2 * NCOL * kpart + CW * half);
    const bool last_layer = !TWO || second;
    float bias_r[PPL][4];
#pragma unroll
    for (int e = 0; e < PPL; ++e)
#pragma unroll
      for (int k = 0; k < 4; ++k) bias_r[e][k] = second ? __ldg(p.bias2 + k * LT_H + ub * UPC + u0 + e) : 0.f;
    uint32_t n = 0;
    for (int t = 0; t < p.T; ++t) {
      const bool store_h = t + 1 < p.T || (TWO && !second);
      uint8_t* hn = reinterpret_cast<uint8_t*>(second ? p.h2 : p.h1) + (long long)slot_h(t + 1) * buf_bytes;
      for (int gl = 0; gl < n_local; ++gl, ++n) {
        const int g = bp + gl * NB;
        const int acc = (int)(n & 1u);
        const int item = g * IG + il;
        const bool valid = item < p.B;
        float pg[PPL][4], skipv[PPL];
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
          skipv[e] = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) pg[e][k] = bias_r[e][k];
          if (valid) {
            const int unit = ub * UPC + u0 + e;
            if (!second) {
              const float* pr = p.pre + (long long)item * p.pre_stride + (long long)t * (4 * LT_H) + unit;
#pragma unroll
              for (int k = 0; k < 4; ++k) pg[e][k] = __ldg(pr + k * LT_H);
            }
            if (last_layer && p.skip) skipv[e] = __ldg(p.skip + (long long)item * p.skip_stride + (long long)t * LT_H + unit);
          }
        }
        if (gl == 0 && warp == 4) LW_TRACE(2, t, 0)
        mbar_wait_g(accf_bar(acc), (n >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (gl == 0 && warp == 4) LW_TRACE(2, t, 1)
        // h1 rows (lanes 0-15): main + corr / 2^11 with corr = h1 w2; h2 rows (lanes 16-31): their "main" columns are h2 w1.
        // Sources: the accumulators of the h part, then (layer 2) those of the x part.
        float part[CW];
#pragma unroll
        for (int k = 0; k < CW; ++k) part[k] = 0.f;
        const uint32_t a0 = lane_base + (uint32_t)(acc * BUF_COLS);
#pragma unroll
        for (int srcp = 0; srcp < (TWO ? 2 : 1); ++srcp) {
          if (srcp == 1 && !second) break;
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
            float mt[CW], ct[CW];
            tmem_ld_cols<CW>(a0 + (uint32_t)(srcp * ACC_COLS + NMMA * a), mt);
            tmem_ld_cols<CW>(a0 + (uint32_t)(srcp * ACC_COLS + NMMA * a + NCOL), ct);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int k = 0; k < CW; ++k)
              part[k] += lane < 16 ? fmaf(ct[k], 1.f / LT_LO_SCALE, mt[k]) : mt[k] * (1.f / LT_LO_SCALE);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(acce_bar(acc));
        if (gl == 0 && warp == 4) LW_TRACE(2, t, 2)
#pragma unroll
        for (int k = 0; k < CW; ++k) part[k] += __shfl_xor_sync(0xffffffffu, part[k], 16);
        // rec[e][k]: recurrent (+ x part) sum of gate k of this lane's pair e
        float rec[PPL][4];
        if (KP == 2) {
          // columns [4 UWW kp', 4 UWW (kp' + 1)) of `part` belong to the units warp kp' of the team finishes: hand the partner's
          // columns over, add the partner's contribution to the own ones
          constexpr int HC = 4 * UWW;                                     // 8 columns per warp
          float* xw = xs + ((team * 2 + (1 - kpart)) * 16 + r) * HC;      // slot read by the partner
          const float* xr = xs + ((team * 2 + kpart) * 16 + r) * HC;      // slot written by the partner
          if (lane < 16) {
#pragma unroll
            for (int k = 0; k < HC; k += 4) {
              float4 v;
              if (kpart == 0) v = make_float4(part[HC + k], part[HC + k + 1], part[HC + k + 2], part[HC + k + 3]);
              else v = make_float4(part[k], part[k + 1], part[k + 2], part[k + 3]);
              *reinterpret_cast<float4*>(xw + k) = v;
            }
          }
          asm volatile("bar.sync %0, %1;" ::"r"(2 + team), "r"(64) : "memory");   // both K parts are in shared memory
          const float4 xv = *reinterpret_cast<const float4*>(xr + lg * 4);        // the partner's part of this lane's unit
          asm volatile("bar.sync %0, %1;" ::"r"(2 + team), "r"(64) : "memory");   // ... and read: the slots may be rewritten
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float o0 = part[k], o1 = part[4 + k], o2 = part[HC + k], o3 = part[HC + 4 + k];   // own sums of the team's 4 units
            const float own = kpart == 0 ? (lane < 16 ? o0 : o1) : (lane < 16 ? o2 : o3);
            rec[0][k] = own + (k == 0 ? xv.x : k == 1 ? xv.y : k == 2 ? xv.z : xv.w);
          }
        } else {
#pragma unroll
          for (int e = 0; e < PPL; ++e)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float lo = part[e * 4 + k], hi = part[(PPL + e) * 4 + k];
              rec[e][k] = lane < 16 ? lo : hi;
            }
        }
        if (gl == 0 && warp == 4) LW_TRACE(2, t, 4)
        unsigned int own[PPL];
        float h_new[PPL];
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
          // gates i, f, o = sigmoid, g = tanh = 2 sigmoid(2 x) - 1: four sigmoids on two packed pairs
          float s_if[2] = {pg[e][0] + rec[e][0], pg[e][1] + rec[e][1]};
          float s_go[2] = {2.f * (pg[e][2] + rec[e][2]), pg[e][3] + rec[e][3]};
          sigmoid_pair(s_if);
          sigmoid_pair(s_go);
          float* cptr = cs + (gl * IG + il) * UPC + u0 + e;
          const float c_new = s_if[1] * (*cptr) + s_if[0] * (2.f * s_go[0] - 1.f);
          *cptr = c_new;
          float s_c[2] = {2.f * c_new, 0.f};
          sigmoid_pair(s_c);
          h_new[e] = s_go[1] * (2.f * s_c[0] - 1.f);
          const __half q1 = __float2half_rn(h_new[e]);
          const __half q2 = __float2half_rn((h_new[e] - __half2float(q1)) * LT_LO_SCALE);
          own[e] = (unsigned int)__half_as_ushort(q1) | ((unsigned int)__half_as_ushort(q2) << 16);
        }
        if (gl == 0 && warp == 4) LW_TRACE(2, t, 5)
        unsigned int oth[PPL];
#pragma unroll
        for (int e = 0; e < PPL; ++e) oth[e] = __shfl_down_sync(0xffffffffu, own[e], 16);
        if (lane < 16 && store_h) {   // the UWW units of this item this warp finished: one store per split part
          // shared-memory image of K tile (UPC ub) / 64: row (g, iq, r), 16-byte chunk (unit % 64) / 8 swizzled by r & 7
          const int kk = (ub * UPC + uw0) % LT_KT;
          const long long row = ((long long)g * LT_NKT + (ub * UPC) / LT_KT) * ROWS + iq * 32 + r;
          uint8_t* dst = hn + row * 128 + (((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2;
          if (PPL == 1) {
            const unsigned int v1 = (own[0] & 0xffffu) | (oth[0] << 16);
            const unsigned int v2 = (own[0] >> 16) | (oth[0] & 0xffff0000u);
            asm volatile("st.global.cg.b32 [%0], %1;" ::"l"(dst), "r"(v1) : "memory");                // h1 row
            asm volatile("st.global.cg.b32 [%0], %1;" ::"l"(dst + 16 * 128), "r"(v2) : "memory");     // h2 row (same r & 7)
          } else {
            // words of the h1 row: this lane's PPL units, then the PPL units of lane + 16; same for the h2 row
            constexpr int PW = PPL >= 2 ? PPL : 2;
            unsigned int w1[PW], w2[PW];
#pragma unroll
            for (int i = 0; i < PPL / 2; ++i) {
              w1[i] = (own[2 * i] & 0xffffu) | (own[2 * i + 1] << 16);
              w1[PPL / 2 + i] = (oth[2 * i] & 0xffffu) | (oth[2 * i + 1] << 16);
              w2[i] = (own[2 * i] >> 16) | (own[2 * i + 1] & 0xffff0000u);
              w2[PPL / 2 + i] = (oth[2 * i] >> 16) | (oth[2 * i + 1] & 0xffff0000u);
            }
            if (PPL == 2) {
              asm volatile("st.global.cg.v2.b32 [%0], {%1, %2};" ::"l"(dst), "r"(w1[0]), "r"(w1[1]) : "memory");
              asm volatile("st.global.cg.v2.b32 [%0], {%1, %2};" ::"l"(dst + 16 * 128), "r"(w2[0]), "r"(w2[1]) : "memory");
            } else {
              asm volatile("st.global.cg.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(w1[0]), "r"(w1[1]), "r"(w1[PPL == 4 ? 2 : 0]),
                           "r"(w1[PPL == 4 ? 3 : 0])
                           : "memory");
              asm volatile("st.global.cg.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 16 * 128), "r"(w2[0]), "r"(w2[1]),
                           "r"(w2[PPL == 4 ? 2 : 0]), "r"(w2[PPL == 4 ? 3 : 0])
                           : "memory");
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(hst_bar(acc));   // -> publisher warp
        if (gl == 0 && warp == 4) LW_TRACE(2, t, 3)
        // the layer output is nobody's business inside this kernel: stored after h_t is on its way
        if (valid && last_layer) {
#pragma unroll
          for (int e = 0; e < PPL; ++e) {
            float y = h_new[e] + skipv[e];
            if (p.out_elu) y = elu1(y);
            p.out[(long long)item * p.out_stride + (long long)t * LT_H + ub * UPC + u0 + e] = y;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(LW_TMEM_COLS) : "memory");
  }
}

// W_hh [4H][H] (reference layout) -> fp16 split slices [unit blocks][w1: 4 upc rows | w2: 4 upc rows][512], row n = unit * 4 + gate
__global__ void lstm_tc_pack_kernel(const float* __restrict__ whh, __half* __restrict__ wp, int upc) {
  const int ncol = 4 * upc;
  const int row = blockIdx.x;                 // 0 .. 2 * 4H - 1
  const int ub = row / (2 * ncol);
  const int part = (row % (2 * ncol)) / ncol;
  const int n = row % ncol;
  const int src = (n % 4) * LT_H + ub * upc + n / 4;
  for (int k = threadIdx.x; k < LT_H; k += blockDim.x) {
    const float w = whh[(long long)src * LT_H + k];
    const __half w1 = __float2half_rn(w);
    wp[(long long)row * LT_H + k] = part == 0 ? w1 : __float2half_rn((w - __half2float(w1)) * LT_LO_SCALE);
  }
}

size_t lt_smem_bytes(int G, int ig, int upc) {
  const int nb = upc == 8 ? 2 : 1;
  return 1024 + LT_STAGES * LT_STAGE_BYTES + LT_NKT * (2 * 4 * upc * 128) + (size_t)((G + nb - 1) / nb) * ig * upc * 4 + 4 * 16 * (2 * upc) * 4 +
         8 * (2 * LT_STAGES + 7) + 16;
}

}  // namespace

// packed: 2 x [8H][H] halves: the UPC = 4 slices followed by the UPC = 8 slices
int launch_lstm_tc_pack(const float* whh, void* packed, int H, cudaStream_t s) {
  ECB_REQUIRE(H == LT_H, "lstm_tc: hidden size %d unsupported", H);
  __half* wp = reinterpret_cast<__half*>(packed);
  lstm_tc_pack_kernel<<<8 * LT_H, 128, 0, s>>>(whh, wp, 4);
  lstm_tc_pack_kernel<<<8 * LT_H, 128, 0, s>>>(whh, wp + 8LL * LT_H * LT_H, 8);
  ECB_LAUNCHED();
  return 0;
}

// one [4H][H] matrix (W_hh or W_ih) -> [8H][H] halves in the slice order of `upc` units per CTA
int launch_lstm_tc_pack_upc(const float* w, void* packed, int H, int upc, cudaStream_t s) {
  ECB_REQUIRE(H == LT_H && (upc == 4 || upc == 8 || upc == 16), "lstm_tc: pack H %d / upc %d unsupported", H, upc);
  lstm_tc_pack_kernel<<<8 * LT_H, 128, 0, s>>>(w, reinterpret_cast<__half*>(packed), upc);
  ECB_LAUNCHED();
  return 0;
}

long long* g_lstm_tc_trace = nullptr;   // diagnostic (ecb_debug_lstm_trace): device buffer of 3 * 8 * 16 stamps

bool lstm_tc_supported(int batch, int H) { return H == LT_H && batch >= 1 && batch <= LT_MAX_GROUPS * 64; }

// floats of workspace: exchange buffers ([2][G][8][2 IG][64] fp16) + counters
int lstm_tc_workspace_floats(int batch) {
  // exchange buffers of 2 (single layer) or 4 + 2 (two-layer wavefront, up to 128 items) slots of [items rounded up to 64][2 parts][512]
  // halves, + arrival counters
  const int rb = (batch + 63) / 64 * 64;
  return rb * (batch <= 128 ? 3072 : 1024) + 1024;
}

int launch_lstm_tc(const float* pre, long long pre_item_stride, const void* w_packed, const float* skip, long long skip_item_stride,
                   float* out, long long out_item_stride, int batch, int T, int out_elu, float* workspace, cudaStream_t s) {
  ECB_REQUIRE(lstm_tc_supported(batch, LT_H) && T > 0, "lstm_tc: bad batch %d / T %d", batch, T);
  // group size: batches of up to 64 items run as groups of 32, larger ones as groups of 64; units per CTA: 8 (64 unit
  // blocks x 2 batch parts) from 192 items -- measured (us per layer step, UPC 4 / 8): 64 items 4.5 / 5.1, 128: 6.8 / 7.7,
  // 256: 12.5 / 11.2, 512: 24.8 / 20.2, 960: 46.3 / 39.3 (the wider MMA and the doubled cell work per lane are not free)
  int ig = batch <= 64 ? 32 : 64;
  int upc = batch < 192 ? 4 : 8;
  if (const char* e = getenv("ECB_LSTM_GROUP")) {   // diagnostic overrides
    const int v = atoi(e);
    if (v == 32 || v == 64) ig = v;
  }
  if (const char* e = getenv("ECB_LSTM_UPC")) {
    const int v = atoi(e);
    if (v == 4 || v == 8) upc = v;
  }
  if ((batch + ig - 1) / ig > LT_MAX_GROUPS) ig = 64;
  const int G = (batch + ig - 1) / ig;
  const int rows = 2 * ig;
  const int nb = upc == 8 ? 2 : 1;
  const size_t smem = lt_smem_bytes(G, ig, upc);
  ECB_REQUIRE(smem <= 227 * 1024, "lstm_tc: %zu bytes of shared memory for %d groups", smem, G);
  int cl = (G + nb - 1) / nb >= 4 ? 4 : 1;   // cluster size: CTAs that share their h tiles through multicast (pays once L2 -> SM traffic matters)
  if (const char* e = getenv("ECB_LSTM_CLUSTER")) cl = atoi(e);
  ECB_REQUIRE(cl == 1 || cl == 2 || cl == 4, "lstm_tc: ECB_LSTM_CLUSTER=%d (1, 2 or 4)", cl);
  LstmTcParams p;
  p.pre = pre;
  p.pre_stride = pre_item_stride;
  p.skip = skip;
  p.skip_stride = skip_item_stride;
  p.out = out;
  p.out_stride = out_item_stride;
  const long long hbuf = 2LL * G * rows * LT_H;   // halves
  p.hg = reinterpret_cast<__half*>(workspace);
  p.cnt = reinterpret_cast<unsigned int*>(p.hg + hbuf);
  p.B = batch;
  p.T = T;
  p.G = G;
  p.out_elu = out_elu;
  p.trace = g_lstm_tc_trace;
  // h_{-1} = 0 (both buffers: padding rows of the last group stay finite) and the arrival counters
  ECB_CUDA(cudaMemsetAsync(workspace, 0, sizeof(__half) * hbuf + sizeof(unsigned int) * (size_t)(G * LT_NKT), s));
  CUtensorMap map_w;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)LT_H, (cuuint64_t)(8 * LT_H)};
    const cuuint64_t strides[1] = {(cuuint64_t)LT_H * 2};
    const cuuint32_t box[2] = {LT_KT, (cuuint32_t)(2 * 4 * upc)};
    const __half* wp = reinterpret_cast<const __half*>(w_packed) + (upc == 8 ? 8LL * LT_H * LT_H : 0LL);
    if (make_tensor_map_f16(&map_w, wp, 2, dims, strides, box)) return 1;
  }
  const double bt = (double)batch * T;
  ProfScope prof(PROF_LSTM_REC, s, 2.0 * bt * 4 * LT_H * LT_H, 4.0 * (bt * 4 * LT_H + bt * LT_H * (skip ? 2 : 1) + 4.0 * LT_H * LT_H));
  // all 128 CTAs spin on each other's arrivals: they must be co-resident (cooperative launch), in clusters of `cl`. Where a
  // clustered cooperative launch is refused (seen under ncu), the same kernel runs unclustered.
  void* args[] = {(void*)&map_w, (void*)&p};
#define LT_FN(CL_, IG_) {(const void*)lstm_tc_kernel<CL_, IG_, 4>, (const void*)lstm_tc_kernel<CL_, IG_, 8>}
  const void* fns[3][2][2] = {{LT_FN(1, 32), LT_FN(1, 64)}, {LT_FN(2, 32), LT_FN(2, 64)}, {LT_FN(4, 32), LT_FN(4, 64)}};
#undef LT_FN
  static DeviceOnce attr_set[3][2][2];
  for (;;) {
    const int ci = cl == 4 ? 2 : cl == 2 ? 1 : 0, gi = ig == 64 ? 1 : 0, ui = upc == 8 ? 1 : 0;
    const void* fn = fns[ci][gi][ui];
    DeviceOnce& once = attr_set[ci][gi][ui];
    if (!once.done()) {
      ECB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      once.mark();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(LT_CTAS);
    cfg.blockDim = dim3(LT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    attrs[na].id = cudaLaunchAttributeCooperative;
    attrs[na].val.cooperative = 1;
    ++na;
    if (cl > 1) {
      attrs[na].id = cudaLaunchAttributeClusterDimension;
      attrs[na].val.clusterDim.x = (unsigned)cl;
      attrs[na].val.clusterDim.y = 1;
      attrs[na].val.clusterDim.z = 1;
      ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = (unsigned)na;
    const cudaError_t le = cudaLaunchKernelExC(&cfg, fn, args);
    if (le != cudaSuccess && cl > 1) {
      cudaGetLastError();   // not sticky: retry without clusters
      cl = 1;
      continue;
    }
    ECB_CUDA(le);
    break;
  }
  ECB_LAUNCHED();
  return 0;
}


namespace {
template <int CL, int IG, int UPC, bool TWO>
int launch_tcw_one(const CUtensorMap* maps, const LstmTcwParams& p, int cl_req, size_t smem, cudaStream_t s) {
  const void* fn = (const void*)lstm_tcw_kernel<CL, IG, UPC, TWO>;
  static DeviceOnce once;
  if (!once.done()) {
    ECB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    once.mark();
  }
  (void)cl_req;
  void* args[] = {(void*)&maps[0], (void*)&maps[1], (void*)&maps[2], (void*)&p};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(LT_CTAS);
  cfg.blockDim = dim3(LT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  attrs[na].id = cudaLaunchAttributeCooperative;   // all CTAs spin on each other's arrivals: they must be co-resident
  attrs[na].val.cooperative = 1;
  ++na;
  if (CL > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = (unsigned)CL;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = (unsigned)na;
  const cudaError_t le = cudaLaunchKernelExC(&cfg, fn, args);
  if (le != cudaSuccess) {
    cudaGetLastError();
    return CL > 1 ? -1 : (set_error("lstm_tcw launch failed: %s", cudaGetErrorString(le)), 1);   // -1: retry without clusters
  }
  return 0;
}

int make_w_map(CUtensorMap* map, const void* packed, int upc) {
  const cuuint64_t dims[2] = {(cuuint64_t)LT_H, (cuuint64_t)(8 * LT_H)};
  const cuuint64_t strides[1] = {(cuuint64_t)LT_H * 2};
  const cuuint32_t box[2] = {LT_KT, (cuuint32_t)(2 * 4 * upc)};
  return make_tensor_map_f16(map, packed, 2, dims, strides, box);
}
}  // namespace

bool lstm_tc2_supported(int batch, int H) { return H == LT_H && batch >= 1 && batch <= 128; }

// Both layers of the SLSTM as one wavefront kernel (lstm_tcw_kernel<.., TWO = true>). pre: first layer's pre-gates; w1h / w2x / w2h:
// UPC = 8 packings (launch_lstm_tc_pack_upc) of W_hh(layer 1), W_ih(layer 2), W_hh(layer 2); bias2 = b_ih + b_hh of layer 2.
int launch_lstm_tc2(const float* pre, long long pre_item_stride, const void* w1h, const void* w2x, const void* w2h, const float* bias2,
                    const float* skip, long long skip_item_stride, float* out, long long out_item_stride, int batch, int T, int out_elu,
                    float* workspace, cudaStream_t s) {
  ECB_REQUIRE(lstm_tc2_supported(batch, LT_H) && T > 0, "lstm_tc2: bad batch %d / T %d", batch, T);
  int ig = batch <= 64 ? 32 : 64;
  if (const char* e = getenv("ECB_LSTM_GROUP")) {
    const int v = atoi(e);
    if (v == 64 || (v == 32 && batch <= 64)) ig = v;
  }
  const int G = (batch + ig - 1) / ig;
  const int rows = 2 * ig;
  LstmTcwParams p;
  p.pre = pre;
  p.pre_stride = pre_item_stride;
  p.bias2 = bias2;
  p.skip = skip;
  p.skip_stride = skip_item_stride;
  p.out = out;
  p.out_stride = out_item_stride;
  const long long slot = (long long)G * rows * LT_H;   // halves per exchange slot
  p.h1 = reinterpret_cast<__half*>(workspace);
  p.h2 = p.h1 + LW_RING1 * slot;
  p.cnt1 = reinterpret_cast<unsigned int*>(p.h2 + 2 * slot);
  p.cnt2 = p.cnt1 + G * LT_NKT;
  p.B = batch;
  p.T = T;
  p.G = G;
  p.out_elu = out_elu;
  p.trace = g_lstm_tc_trace;
  p.trace_cta = getenv("ECB_LSTM_TRACE_CTA") ? atoi(getenv("ECB_LSTM_TRACE_CTA")) : 64;
  ECB_CUDA(cudaMemsetAsync(workspace, 0, sizeof(__half) * (LW_RING1 + 2) * slot + sizeof(unsigned int) * (size_t)(2 * G * LT_NKT), s));
  CUtensorMap maps[3];
  if (make_w_map(&maps[0], w1h, 8) || make_w_map(&maps[1], w2x, 8) || make_w_map(&maps[2], w2h, 8)) return 1;
  const size_t smem = 1024 + LW_CS_OFF + (size_t)G * ig * 8 * 4 + 4 * 16 * 16 * 4 + 8 * (2 * LT_STAGES + 7) + 16;
  ECB_REQUIRE(smem <= 227 * 1024, "lstm_tc2: %zu bytes of shared memory", smem);
  const double bt = (double)batch * T;
  ProfScope prof(PROF_LSTM_REC, s, 2.0 * bt * 4 * LT_H * LT_H * 3, 4.0 * (bt * 4 * LT_H + bt * LT_H * 2 + 12.0 * LT_H * LT_H));
  const int rc = ig == 32 ? launch_tcw_one<1, 32, 8, true>(maps, p, 1, smem, s) : launch_tcw_one<1, 64, 8, true>(maps, p, 1, smem, s);
  if (rc) return 1;
  ECB_LAUNCHED();
  return 0;
}

// One layer, 16 units per CTA (32 unit blocks x 4 batch parts): large launches. w_packed = UPC 16 packing of W_hh.
int launch_lstm_tc16(const float* pre, long long pre_item_stride, const void* w_packed, const float* skip, long long skip_item_stride,
                     float* out, long long out_item_stride, int batch, int T, int out_elu, float* workspace, cudaStream_t s) {
  ECB_REQUIRE(lstm_tc_supported(batch, LT_H) && T > 0, "lstm_tc16: bad batch %d / T %d", batch, T);
  constexpr int ig = 64, nb = 4;
  const int G = (batch + ig - 1) / ig;
  const int rows = 2 * ig;
  LstmTcwParams p;
  p.pre = pre;
  p.pre_stride = pre_item_stride;
  p.bias2 = nullptr;
  p.skip = skip;
  p.skip_stride = skip_item_stride;
  p.out = out;
  p.out_stride = out_item_stride;
  const long long slot = (long long)G * rows * LT_H;
  p.h1 = reinterpret_cast<__half*>(workspace);
  p.h2 = nullptr;
  p.cnt1 = reinterpret_cast<unsigned int*>(p.h1 + 2 * slot);
  p.cnt2 = nullptr;
  p.B = batch;
  p.T = T;
  p.G = G;
  p.out_elu = out_elu;
  p.trace = g_lstm_tc_trace;
  p.trace_cta = getenv("ECB_LSTM_TRACE_CTA") ? atoi(getenv("ECB_LSTM_TRACE_CTA")) : 0;
  ECB_CUDA(cudaMemsetAsync(workspace, 0, sizeof(__half) * 2 * slot + sizeof(unsigned int) * (size_t)(G * LT_NKT), s));
  CUtensorMap maps[3];
  if (make_w_map(&maps[0], w_packed, 16)) return 1;
  maps[1] = maps[0];
  maps[2] = maps[0];
  const size_t smem = 1024 + LW_CS_OFF + (size_t)((G + nb - 1) / nb) * ig * 16 * 4 + 8 * (2 * LT_STAGES + 7) + 16;
  ECB_REQUIRE(smem <= 227 * 1024, "lstm_tc16: %zu bytes of shared memory for %d groups", smem, G);
  int cl = (G + nb - 1) / nb >= 2 ? 4 : 1;
  if (const char* e = getenv("ECB_LSTM_CLUSTER")) cl = atoi(e);
  ECB_REQUIRE(cl == 1 || cl == 2 || cl == 4, "lstm_tc16: ECB_LSTM_CLUSTER=%d (1, 2 or 4)", cl);
  const double bt = (double)batch * T;
  ProfScope prof(PROF_LSTM_REC, s, 2.0 * bt * 4 * LT_H * LT_H, 4.0 * (bt * 4 * LT_H + bt * LT_H * (skip ? 2 : 1) + 4.0 * LT_H * LT_H));
  int rc = -1;
  if (cl == 4) rc = launch_tcw_one<4, 64, 16, false>(maps, p, cl, smem, s);
  else if (cl == 2) rc = launch_tcw_one<2, 64, 16, false>(maps, p, cl, smem, s);
  if (rc < 0) rc = launch_tcw_one<1, 64, 16, false>(maps, p, 1, smem, s);   // also the retry where a clustered cooperative launch is refused
  if (rc) return 1;
  ECB_LAUNCHED();
  return 0;
}

}  // namespace ecb
