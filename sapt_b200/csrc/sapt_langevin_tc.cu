// tcgen05 / TMEM / TMA batch-Langevin gradient (sm_100a only): the fixed-theta restatement of
// Network.langevin_gradient (LG:138-156, BackwardPass LG:96-106)
//     theta_gd = theta + lr * scale * sum_rows g_row(theta)
// with ALL five contractions on the tensor cores. Per chain and per 128-row tile:
//
//   MMA-1  D1[rows, HP]   = X . W1'                      smem (K-major) x smem -> TMEM   forward layer 1
//   P1     h = sigmoid(D1) -> bf16 pieces [h1 | h2] -> shared memory Hs (row-major, SWIZZLE_128B)
//   MMA-2  D2[rows, 3 ON] = Hs . [w1|w2|w3], += h2 . [w1|w2]   Hs read K-major               forward layer 2
//   P2     out = sigmoid(D2 + b2); do = (1[y] - out) out (1 - out) -> pieces [do_hi | do_lo] -> Ds
//   G2     G2[2 HP, 2 ON] += Hs^T . Ds                    both operands read MN-major, K = rows   grad W2
//   MMA-3  D3[rows, HP]   = Ds . W2^T                     Ds read K-major                         back-propagation
//   P3     dh = D3 * h (1 - h) -> pieces [dh_hi | dh_lo] -> Hs (overwrites h)
//   G1     G1[x columns, 2 HP] += X^T . Hs                X tile and Hs read MN-major, K = rows   grad W1 (+ B1)
//
// The row contractions (G1, G2) accumulate in TMEM over all tiles of the CTA's row chunk; at the end of a chain
// the epilogue warps fold the bf16-piece partial products into one fp32 gradient vector in shared memory and write
// it to part[chunk][chain][P]; langevin_batch_finalize (sapt_langevin.cu) sums the chunks and applies the step.
// The same bytes of Hs / Ds / the X tile serve as K-major operands (forward / back-propagation: K = features)
// and as MN-major operands (gradients: K = rows): only the descriptor changes.
// B1 gradient: a column of ones in the X image (the bias chunk, or for D > 32 the first spare column of the x1
// piece) makes -sum(dh) a row of G1. B2 gradient: per-thread sums of do, reduced at the end of the chain.
// Operands are bf16 pieces as in round 1 of sapt_loglik_tc2.cu: x = x1 + x2 (+ x3 dropped: 2^-17), w = w1 + w2 (+ w3 in the
// output layer), h, do, dh = two pieces each; the cross terms kept are listed at each MMA (gradient accuracy: 2e-6 max|g|).
//
// Work item = (chain, row chunk); one CTA per SM takes items round-robin, one after the other (the accumulators are per item).
// Inside an item TWO tile streams run independently: epilogue warps 0-7 take the even tiles, warps 8-15 the odd ones; a stream owns one
// Hs buffer, one TMEM scratch block, its G1 / G2 accumulators and its own MMA-issuer warp (17 / 18; warp 17 also allocates TMEM).
// Warp 16: TMA producer (ring of three X tiles shared by the streams; per-chain weight panels). Epilogue warp = (stream, quarter q =
// TMEM lanes 32q.., half = HP/2 hidden units in P1 / P3 and half the outputs in P2). Single-thread instructions are issued by an
// elected lane of a converged warp (elect_one).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "sapt_device.cuh"
#include "sapt_internal.h"
#include "sapt_tc_ptx.cuh"

namespace sapt {

int launch_tc2_pack(const float* theta, int64_t ld, const int32_t* rows, const int32_t* n_rows, int R, int D, int H, int O, int HP, int ON, int CH,
                    int DKC, int KROW, int biasx, __nv_bfloat16* b1img, __nv_bfloat16* w2img, float* bias2, int n_blocks, cudaStream_t st);   // sapt_loglik_tc2.cu
int tc_make_map_2d(CUtensorMap* map, void* base, uint64_t inner_elems, uint64_t rows, uint32_t box_rows);                       // sapt_tc_host.cu
int launch_langevin_batch_finalize(const float* part, int n_chunks, int R, int P, const float* theta, float* theta_gd, int64_t ld,
                                   const int32_t* rows, const int32_t* n_rows, double lr_scale, cudaStream_t st);                // sapt_langevin.cu

namespace lt {
constexpr int kTileM = 128;
constexpr int kEpiWarps = 16;
constexpr int kThreads = 32 * (3 + kEpiWarps);
constexpr int kTmaWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;      // MMA issuers: warp kMmaWarp (tile stream 0) and kMmaWarp + 1 (stream 1)
constexpr int ON = 16;                    // outputs padded (O <= 16)
constexpr int N2 = 3 * ON;
constexpr uint32_t kBlk = 128 * 128;      // one 64-column block of a 128-row tile

// DK packs DKC (DK & 7) and "O <= 8" (DK >= 8): the output phase then keeps 8 instead of 16 accumulator columns per
// partial sum in registers. With ~210 KB of shared memory in use the L1 behind local memory is tiny, so spills are
// served from L2 (see the same finding for the likelihood kernel, profiles/r1_c4_loglik_summary.json).
//
// Shared memory is what limits the overlap, so the X tile is kept to the blocks the products need: x1 and x2 (x3 is not used,
// see MMA-1) — and for DKC >= 3 the bias rides in the spare x1 column (a 1.0 written by tc_prepare_dataset; the bias pieces sit
// in that column of the w1 / w2 pieces of B1), so the tile is two 64-column blocks whatever D is: THREE X stages fit beside
// two Hs buffers, and the tile a group needs next is already in shared memory when its previous tile retires.
template <int HP, int DK>
struct Cfg {
  static constexpr int DKC = DK & 7;
  static constexpr int NRL = (DK >= 8) ? 8 : 16;              // output columns read / accumulated per partial sum
  static constexpr int NQ = NRL / 2;                          // ... per epilogue warp (the two halves of a group split the outputs)
  static constexpr int DW = 16 * DKC;                         // features per x piece (padded)
  static constexpr bool BIASX = DKC >= 3;                     // bias in the spare x1 column (needs D < 16 DKC: langevin_tc_supported)
  static constexpr int KBA = (3 * DKC + 1 + 3) / 4;           // 128-byte K blocks of an image row (host side: row length)
  static constexpr int NSLOT = BIASX ? 2 : KBA;               // X / B1 blocks kept in shared memory (1 or 2)
  static constexpr int BIAS_CHUNK = 3 * DKC;                  // bias chunk of the image row (used when !BIASX)
  static constexpr int NSTAGE = 3;                            // X tile ring
  static constexpr int NBH = (2 * HP + 63) / 64;              // blocks of Hs: [piece 1 (HP) | piece 2 (HP)]
  static constexpr int HW = HP / 2;                           // hidden units per epilogue warp
  static constexpr int NG1 = NBH * 64;                        // G1 columns
  static constexpr uint32_t kAStage = NSLOT * kBlk;
  static constexpr uint32_t kB1Bytes = HP * NSLOT * 128;
  static constexpr uint32_t kW2Bytes = N2 * 128;
  static constexpr uint32_t kW2TBytes = HP * 128;
  static constexpr uint32_t kHsBytes = NBH * kBlk;
  __host__ __device__ static constexpr uint32_t smem_bytes() {   // X ring, one Hs buffer per group, one shared Ds block
    return 1024u + NSTAGE * kAStage + 2 * kHsBytes + kBlk + ((kB1Bytes + 1023u) & ~1023u) + ((kW2TBytes + 1023u) & ~1023u) +
           ((kW2Bytes + 1023u) & ~1023u) + 512u;
  }
  // TMEM columns: scratch D1/D2/D3 per group | G1 per group | G2 per group
  static constexpr int kDS = 0, kG1 = 128, kG2 = 128 + 2 * NG1, kCols = 512;
  static_assert(128 + 2 * NG1 + 128 <= kCols && HP <= 64 && N2 <= 64, "TMEM budget");
  // column of ones in the X image that turns -sum(dh) into a row of G1 (see tc_prepare_dataset)
  __host__ __device__ static constexpr int ones_col(int D) { return DKC <= 2 ? 3 * DW : D; }
};

struct Params {
  const int32_t* y;
  const float* bias2;                 // [pos][ON] log2(e) * B2 (used when the bias is not folded into the W2 image)
  const int32_t* n_rows; int R;
  int D, H, O, fold_bias;
  int tiles_per_chunk; int64_t n_tiles;
  int n_chunks;                       // work item = (chain, row chunk); item i runs on CTA i % gridDim.x
  float* part;                        // [chunk][R][P]
  int P;
  long long* trace;                   // development: clock64 stamps of CTA 0 (SAPT_LT_TRACE=1), else null
};
// development trace: event e of tile-visit n (first 64) of a role; roles: 0 = epilogue group 0 warp 0, 1 = group 1 warp 8, 2 / 3 = MMA warp (group 0 / 1)
#define LT_TRACE2(role, n, e) do { if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[((role) * 64 + (n)) * 8 + (e)] = clock64(); } while (0)
#define LT_TRACE(role, n, e) do { if (p.trace && blockIdx.x == 0 && lane == 0 && c == 0 && (n) < 64) p.trace[((role) * 64 + (n)) * 8 + (e)] = clock64(); } while (0)

__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {   // MN-major, SWIZZLE_128B
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // stride between 64-element blocks along M / N
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride between groups of 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_bf16(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
template <int N>
__device__ __forceinline__ void tmem_ldN(uint32_t taddr, uint32_t (&r)[N]) {
  if constexpr (N == 16) tmem_ld16(taddr, r);
  else if constexpr (N == 8) tmem_ld8(taddr, r);
  else tmem_ld4(taddr, r);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); }

// writes n bf16 values (n = 4, 8 or 16; packed pairs, N/2 words) of one data row into a swizzled [128 rows][64 bf16] block at column col0
template <int N>
__device__ __forceinline__ void store_row_packed(uint8_t* block, int row, int col0, const uint32_t* w) {
  const uint32_t rowp = smem_u32(block) + (uint32_t)row * 128;
  if constexpr (N == 4) {
    const int chunk = col0 >> 3;
    st_shared_v2(rowp + (uint32_t)(((chunk ^ (row & 7)) << 4) | ((col0 & 7) * 2)), w[0], w[1]);
  } else {
#pragma unroll
    for (int c = 0; c < N / 8; ++c) {
      const int chunk = (col0 >> 3) + c;
      st_shared_v4(rowp + (uint32_t)((chunk ^ (row & 7)) << 4), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
    }
  }
}

// W2^T image for MMA-3: row h = [w1 (16 outputs) | w2 (16 outputs) | 0 ..] of +W2[h][:] (unscaled)
__global__ void lt_pack_w2t_kernel(const float* theta, int64_t ld, const int32_t* rows, const int32_t* n_rows, int R, int D, int H, int O, int HP,
                                   __nv_bfloat16* img) {
  const int nact = n_rows ? *n_rows : R;
  const int pos = blockIdx.x;
  __nv_bfloat16* dst = img + (int64_t)pos * HP * 64;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  if (pos >= nact) { for (int i = threadIdx.x; i < HP * 64; i += blockDim.x) dst[i] = zero; return; }
  const int row = rows ? rows[pos] : pos;
  const float* src = theta + (int64_t)row * ld;
  for (int i = threadIdx.x; i < HP * 64; i += blockDim.x) {
    const int h = i >> 6, c = i & 63, o = c & 15, piece = c >> 4;
    __nv_bfloat16 v = zero;
    if (h < H && o < O && piece < 2) {
      const float w = src[D * H + h * O + o];
      const __nv_bfloat16 p1 = __float2bfloat16_rn(w);
      v = (piece == 0) ? p1 : __float2bfloat16_rn(w - __bfloat162float(p1));
    }
    dst[i] = v;
  }
}

// Two tile streams per CTA: epilogue warps 0-7 (group 0) take the even tiles of the item's row chunk, warps 8-15 (group 1) the odd
// ones; each group owns one Hs buffer, one TMEM scratch block and one G2 accumulator and walks its tiles through
// P1 -> P2 -> P3 at its own pace. The MMA warp serves whichever group is ready (a four-state machine per group, polled), so the
// tensor pipe works on one tile while the other tile's epilogue computes — round 1 / early round 2 ran the two tiles in lockstep
// (every stage issued for both before the next one was waited for) and reloaded an X buffer only after its tile had retired:
// 5 400 cycles per tile against ~2 000 of tensor-pipe and ~2 000 of issue work.
template <int HP, int DK>
__global__ void __launch_bounds__(kThreads, 1) langevin_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b1,
                                                                  const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_w2t,
                                                                  Params p) {
  using C = Cfg<HP, DK>;
  constexpr int DKC = C::DKC, NQ = C::NQ, NSTAGE = C::NSTAGE;
  constexpr int DW = C::DW, NSLOT = C::NSLOT, NBH = C::NBH, HW = C::HW, NG1 = C::NG1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                          // [NSTAGE][NSLOT blocks]
  uint8_t* sHs = sA + NSTAGE * C::kAStage;                     // [2 groups][NBH blocks]
  uint8_t* sDs = sHs + 2 * C::kHsBytes;                        // ONE block: group g owns columns [32 g, 32 g + 32)
  uint8_t* sB1 = sDs + kBlk;
  uint8_t* sW2T = sB1 + ((C::kB1Bytes + 1023u) & ~1023u);
  uint8_t* sW2 = sW2T + ((C::kW2TBytes + 1023u) & ~1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW2 + ((C::kW2Bytes + 1023u) & ~1023u));
  uint64_t* full_a = bars;                  // [NSTAGE]
  uint64_t* empty_a = full_a + NSTAGE;      // [NSTAGE]
  // per group g
  uint64_t* d1_full = empty_a + NSTAGE;     // [2]
  uint64_t* hs_full = d1_full + 2;          // [2] 8 warps
  uint64_t* d2_full = hs_full + 2;          // [2]
  uint64_t* ds_full = d2_full + 2;          // [2] 8 warps
  uint64_t* d3_full = ds_full + 2;          // [2]
  uint64_t* dh_full = d3_full + 2;          // [2] 8 warps
  uint64_t* w_full = dh_full + 2;           // weight panels of the current chain landed
  uint64_t* w_free = w_full + 1;            // all MMAs of the chain done: the panels may be overwritten
  uint64_t* g_full = w_free + 1;            // G1 / G2 of the chain complete
  uint64_t* g_free = g_full + 1;            // 16 warps: accumulators read out
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(g_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nact = p.n_rows ? *p.n_rows : p.R;
  // work items (chain, row chunk), chain fastest so that CTAs running side by side stream the same rows through L2;
  // the number of active chains is only known on the device, so the items are dealt round-robin here
  const int n_items = nact * p.n_chunks;
  if ((int)blockIdx.x >= n_items) return;
  struct Item { int pos, chunk, n_tiles; int64_t t_begin; };
  auto get_item = [&](int item) {
    Item it;
    it.pos = item % nact; it.chunk = item / nact;
    it.t_begin = (int64_t)it.chunk * p.tiles_per_chunk;
    it.n_tiles = (int)max((int64_t)0, min(it.t_begin + (int64_t)p.tiles_per_chunk, p.n_tiles) - it.t_begin);
    return it;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&d1_full[g], 1); mbar_init(&hs_full[g], kEpiWarps / 2); mbar_init(&d2_full[g], 1); mbar_init(&ds_full[g], kEpiWarps / 2);
      mbar_init(&d3_full[g], 1); mbar_init(&dh_full[g], kEpiWarps / 2);
    }
    mbar_init(w_full, 1); mbar_init(w_free, 2); mbar_init(g_full, 2); mbar_init(g_free, kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Ds columns that no output phase writes (outputs >= NRL, and the columns past [do_hi | do_lo]) are operands of MMA-3 / G2: zero once
  for (int i = threadIdx.x; i < (int)(kBlk / 16); i += kThreads) st_shared_v4(smem_u32(sDs) + 16u * i, 0u, 0u, 0u, 0u);
  fence_async_smem();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(C::kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == kTmaWarp) {
    // ================= TMA producer: the warp stays converged, one elected lane issues (see elect_one) =================
    uint32_t n_fill = 0;                      // tiles loaded so far (over all items): stage = n_fill % NSTAGE
    for (int item = blockIdx.x, c = 0; item < n_items; item += gridDim.x, ++c) {
      const Item w = get_item(item);
      if (c > 0) mbar_wait<500>(w_free, (c - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(w_full, C::kB1Bytes + C::kW2Bytes + C::kW2TBytes);
#pragma unroll
        for (int kb = 0; kb < NSLOT; ++kb) tma_load_2d(&map_b1, w_full, sB1 + kb * HP * 128, kb * 64, w.pos * HP);
        tma_load_2d(&map_w2, w_full, sW2, 0, w.pos * N2);
        tma_load_2d(&map_w2t, w_full, sW2T, 0, w.pos * HP);
      }
      __syncwarp();
      for (int t = 0; t < w.n_tiles; ++t, ++n_fill) {
        const uint32_t s = n_fill % NSTAGE;
        mbar_wait<200>(&empty_a[s], ((n_fill / NSTAGE) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_a[s], C::kAStage);
#pragma unroll
          for (int j = 0; j < NSLOT; ++j) tma_load_2d(&map_a, &full_a[s], sA + s * C::kAStage + j * kBlk, j * 64, (int)((w.t_begin + t) * kTileM));
        }
        __syncwarp();
      }
    }
  } else if (warp >= kMmaWarp) {
    // ================= MMA issuers: one warp per tile stream (g = warp - kMmaWarp), each blocking on its own stream's barriers; the
    // warp stays converged (waits, descriptor arithmetic on the uniform datapath) and an elected lane issues tcgen05.mma / commit —
    // under `if (lane == 0)` every MMA cost ~100 cycles of ELECT / R2UR.BROADCAST / BRA.U.ANY (see elect_one). A single issuer
    // serving both streams by polling was measured slower: tcgen05.mma issue blocks while the pipe is busy, so each stream's next
    // batch waited for the other stream's batch to be ISSUED, and a failed timed poll costs ~450 cycles. ====
    const int g = warp - kMmaWarp;
    const uint32_t id1 = idesc_bf16(kTileM, HP, false, false);
    const uint32_t id2a = idesc_bf16(kTileM, N2, false, false), id2b = idesc_bf16(kTileM, 2 * ON, false, false);
    const uint32_t id3 = idesc_bf16(kTileM, HP, false, false);
    const uint32_t idg2 = idesc_bf16(128, 64, true, true), idg1 = idesc_bf16(128, NG1, true, true);
    // Operand descriptors: the address field is the low 14 bits (address >> 4) and shared memory ends below 256 KB, so the descriptor
    // of base + offset is descriptor(base) + (offset >> 4). The bases are built once; per tile one multiply-add selects the X stage,
    // per MMA one add of a compile-time constant.
    const uint64_t dXk0 = umma_desc_sw128(smem_u32(sA)), dXm0 = desc_mn(smem_u32(sA), NSLOT > 1 ? kBlk : 0);
    const uint64_t hk = umma_desc_sw128(smem_u32(sHs + g * C::kHsBytes)), hm = desc_mn(smem_u32(sHs + g * C::kHsBytes), NBH > 1 ? kBlk : 0);
    const uint64_t dB1 = umma_desc_sw128(smem_u32(sB1)), dW2 = umma_desc_sw128(smem_u32(sW2)), dW2T = umma_desc_sw128(smem_u32(sW2T));
    const uint64_t da = umma_desc_sw128(smem_u32(sDs) + g * 64), dDsm = desc_mn(smem_u32(sDs), 0);   // da: the stream's 32 Ds columns
    const uint32_t tD = tmem_base + C::kDS + g * 64, tG1 = tmem_base + C::kG1 + g * NG1, tG2 = tmem_base + C::kG2 + g * 64;
    // shared-memory chunk (16 bf16 = 32 B) -> descriptor offset inside a tile / panel made of 64-column blocks
    auto a_off = [](int chunk) { return (uint64_t)(((uint32_t)(chunk >> 2) * kBlk + (uint32_t)(chunk & 3) * 32) >> 4); };
    auto b1_off = [](int chunk) { return (uint64_t)(((uint32_t)(chunk >> 2) * (HP * 128) + (uint32_t)(chunk & 3) * 32) >> 4); };
    uint32_t n_base = 0;                      // tiles of earlier items (the X ring runs across items)
    uint32_t ph = 0;                          // parity of the stream's barriers: they complete once per tile of the stream
    for (int item = blockIdx.x, c = 0; item < n_items; item += gridDim.x, ++c) {
      const Item w = get_item(item);
      const int n_tiles = w.n_tiles;
      mbar_wait(w_full, c & 1);
      if (c == 1) LT_TRACE2(2 + g, 62, 0);
      if (c > 0) mbar_wait(g_free, (c - 1) & 1);     // accumulators of the previous item have been read out
      if (c == 1) LT_TRACE2(2 + g, 62, 1);
      tc_fence_after();
      for (int t = g; t < n_tiles; t += 2) {
        const uint32_t n = n_base + (uint32_t)t;
        const uint32_t s = n % NSTAGE;
        const uint64_t xk = dXk0 + (uint64_t)(s * (C::kAStage >> 4)), xm = dXm0 + (uint64_t)(s * (C::kAStage >> 4));
        const uint32_t first = (t < 2) ? 0u : 1u;
        // ---- MMA-1: x1w1 x1w2 x2w1 (+ bias chunk). x chunks: x1 = [0, DKC), x2 = [DKC, 2 DKC) ----
        // Three piece products: the dropped x1w3 / x2w2 terms are 2^-16 |x w|, i.e. ~1e-4 absolute on a pre-activation, zero-mean
        // over rows — far inside the 1e-4 (relative to max |g|) the gradient needs. (The likelihood kernel keeps fp32-grade
        // products: an MH decision hangs on a difference of sums over all rows.) The scratch block is free (P3 of the stream's
        // previous tile has read D3: dh_full was waited for) and Hs is not written before this batch — issued after that tile's
        // G1 — has completed.
        mbar_wait(&full_a[s], (n / NSTAGE) & 1);
        LT_TRACE(2 + g, t >> 1, 0);
        if (c == 1 && t < 2) LT_TRACE2(2 + g, 62, 2);
        tc_fence_after();
        constexpr int pa[3] = {0, 0, 1};
        constexpr int pb[3] = {0, 1, 0};
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < DKC; ++j)
              tc_mma_bf16(tD, xk + a_off(pa[k] * DKC + j), dB1 + b1_off(pb[k] * DKC + j), id1, (k | j) ? 1u : 0u);
          if constexpr (!C::BIASX) tc_mma_bf16(tD, xk + a_off(C::BIAS_CHUNK), dB1 + b1_off(C::BIAS_CHUNK), id1, 1u);
          tc_commit(&d1_full[g]);
        }
        __syncwarp();
        // ---- MMA-2: D2 = h1.[w1|w2|w3] + h2.[w1|w2], A = Hs read K-major ----
        mbar_wait(&hs_full[g], ph);
        LT_TRACE(2 + g, t >> 1, 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HP / 16; ++k) tc_mma_bf16(tD, hk + a_off(k), dW2 + (uint64_t)(k * 2), id2a, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < HP / 16; ++k) tc_mma_bf16(tD, hk + a_off(HP / 16 + k), dW2 + (uint64_t)(k * 2), id2b, 1u);
          tc_commit(&d2_full[g]);
        }
        __syncwarp();
        // ---- G2[g] += Hs^T . Ds (K = rows; all 64 Ds columns are multiplied, the 32 of stream g are meaningful and the only ones
        //      read out), then MMA-3: D3 = do_hi.w1 + do_hi.w2 + do_lo.w1 ----
        mbar_wait(&ds_full[g], ph);
        LT_TRACE(2 + g, t >> 1, 2);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kTileM / 16; ++k) tc_mma_bf16(tG2, hm + (uint64_t)(k * 128), dDsm + (uint64_t)(k * 128), idg2, (k ? 1u : first));
          tc_mma_bf16(tD, da, dW2T, id3, 0u);
          tc_mma_bf16(tD, da, dW2T + 2, id3, 1u);
          tc_mma_bf16(tD, da + 2, dW2T, id3, 1u);
          tc_commit(&d3_full[g]);
        }
        __syncwarp();
        // ---- G1[g] += X^T . Hs(dh) (K = rows); the X stage is free afterwards ----
        mbar_wait(&dh_full[g], ph);
        LT_TRACE(2 + g, t >> 1, 3);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kTileM / 16; ++k) tc_mma_bf16(tG1, xm + (uint64_t)(k * 128), hm + (uint64_t)(k * 128), idg1, (k ? 1u : first));
          tc_commit(&empty_a[s]);
        }
        __syncwarp();
        ph ^= 1;
      }
      if (elect_one()) {      // each issuer's commits cover its own MMAs: the barriers count both
        tc_commit(g_full);
        tc_commit(w_free);
      }
      __syncwarp();
      n_base += (uint32_t)n_tiles;
    }
  } else {
    // ================= epilogue warps: group g = warp / 8 (tile stream), quarter = TMEM lanes 32 q.., half hf = which half of
    // the hidden units (P1, P3) and of the outputs (P2) =================
    const int grp = warp >> 3, quarter = warp & 3, hf = (warp >> 2) & 1;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int O = p.O, D = p.D, H = p.H;
    float* sGrad = reinterpret_cast<float*>(sHs);        // [P] fp32 at the end of a chain (Hs is idle then)
    uint8_t* hs = sHs + grp * C::kHsBytes;
    const uint32_t tScr = tmem_base + lane_base + (uint32_t)(C::kDS + grp * 64);
    uint32_t ph = 0;
    for (int item = blockIdx.x, c = 0; item < n_items; item += gridDim.x, ++c) {
      const Item w = get_item(item);
      const int pos = w.pos, n_tiles = w.n_tiles;
      float accB2[NQ];
#pragma unroll
      for (int o = 0; o < NQ; ++o) accB2[o] = 0.f;
      for (int t = grp; t < n_tiles; t += 2) {
        float h[HW];
        const int y = __ldg(p.y + (w.t_begin + t) * kTileM + row_in_tile);
        // ---- P1: h = sigmoid(D1), pieces -> Hs. The epilogue is SFU bound: one shared reciprocal per four sigmoids and an integer
        //      bf16 split instead of conversion instructions (see split_pack_bf16x2) ----
        const bool tr = (quarter == 0 && hf == 0);
        if (tr) LT_TRACE(grp, t >> 1, 6);
        mbar_wait(&d1_full[grp], ph);
        if (tr) LT_TRACE(grp, t >> 1, 0);
        tc_fence_after();
        constexpr int HB = (HW > 16) ? 16 : HW;          // hidden units per register batch
#pragma unroll
        for (int j0 = 0; j0 < HW; j0 += HB) {
          uint32_t zr[HB];
          tmem_ldN<HB>(tScr + (uint32_t)(hf * HW + j0), zr);
          tmem_ld_wait();
          uint32_t h1[HB / 2], h2[HB / 2];
#pragma unroll
          for (int j = 0; j < HB; j += 4) {
            sigmoid4_from_neg_log2e(__uint_as_float(zr[j]), __uint_as_float(zr[j + 1]), __uint_as_float(zr[j + 2]), __uint_as_float(zr[j + 3]),
                                    h[j0 + j], h[j0 + j + 1], h[j0 + j + 2], h[j0 + j + 3]);
            split_pack_bf16x2(h[j0 + j], h[j0 + j + 1], h1[j / 2], h2[j / 2]);
            split_pack_bf16x2(h[j0 + j + 2], h[j0 + j + 3], h1[j / 2 + 1], h2[j / 2 + 1]);
          }
          const int col = hf * HW + j0;
          store_row_packed<HB>(hs + (col >> 6) * kBlk, row_in_tile, col & 63, h1);
          store_row_packed<HB>(hs + ((HP + col) >> 6) * kBlk, row_in_tile, (HP + col) & 63, h2);
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hs_full[grp]);
        if (tr) LT_TRACE(grp, t >> 1, 1);
        // ---- P2: output layer and its delta -> Ds; the two halves of the group split the outputs ----
        {
          mbar_wait(&d2_full[grp], ph);
          if (tr) LT_TRACE(grp, t >> 1, 2);
          tc_fence_after();
          uint32_t d0[NQ], d1[NQ], d2[NQ];
          tmem_ldN<NQ>(tScr + (uint32_t)(hf * NQ), d0); tmem_ldN<NQ>(tScr + (uint32_t)(ON + hf * NQ), d1); tmem_ldN<NQ>(tScr + (uint32_t)(2 * ON + hf * NQ), d2);
          tmem_ld_wait();
          uint32_t dhi[NQ / 2], dlo[NQ / 2];
          const bool live = (y >= 0);
#pragma unroll
          for (int o4 = 0; o4 < NQ; o4 += 4) {
            float tt[4], out[4], dv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int o = hf * NQ + o4 + j;
              tt[j] = (__uint_as_float(d0[o4 + j]) + __uint_as_float(d1[o4 + j])) + __uint_as_float(d2[o4 + j]);
              if (!p.fold_bias) tt[j] += __ldg(p.bias2 + (int64_t)pos * ON + o);
            }
            sigmoid4_from_neg_log2e(tt[0], tt[1], tt[2], tt[3], out[0], out[1], out[2], out[3]);   // columns >= O hold t = 0: finite, masked below
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int o = hf * NQ + o4 + j;
              dv[j] = (o < O && live) ? ((o == y ? 1.0f : 0.0f) - out[j]) * (out[j] * (1.0f - out[j])) : 0.0f;     // LG:100
              accB2[o4 + j] += dv[j];
            }
            split_pack_bf16x2(dv[0], dv[1], dhi[o4 / 2], dlo[o4 / 2]);
            split_pack_bf16x2(dv[2], dv[3], dhi[o4 / 2 + 1], dlo[o4 / 2 + 1]);
          }
          store_row_packed<NQ>(sDs, row_in_tile, 32 * grp + hf * NQ, dhi);
          store_row_packed<NQ>(sDs, row_in_tile, 32 * grp + ON + hf * NQ, dlo);
          fence_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ds_full[grp]);
          if (tr) LT_TRACE(grp, t >> 1, 3);
        }
        // ---- P3: dh = (do . W2^T) h (1 - h), pieces -> Hs ----
        mbar_wait(&d3_full[grp], ph);
        if (tr) LT_TRACE(grp, t >> 1, 4);
        ph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int j0 = 0; j0 < HW; j0 += HB) {
          uint32_t zr[HB];
          tmem_ldN<HB>(tScr + (uint32_t)(hf * HW + j0), zr);
          tmem_ld_wait();
          uint32_t g1[HB / 2], g2[HB / 2];
#pragma unroll
          for (int j = 0; j < HB; j += 2)
            split_pack_bf16x2(__uint_as_float(zr[j]) * (h[j0 + j] * (1.0f - h[j0 + j])),                 // LG:101
                              __uint_as_float(zr[j + 1]) * (h[j0 + j + 1] * (1.0f - h[j0 + j + 1])), g1[j / 2], g2[j / 2]);
          const int col = hf * HW + j0;
          store_row_packed<HB>(hs + (col >> 6) * kBlk, row_in_tile, col & 63, g1);
          store_row_packed<HB>(hs + ((HP + col) >> 6) * kBlk, row_in_tile, (HP + col) & 63, g2);
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dh_full[grp]);
        if (tr) LT_TRACE(grp, t >> 1, 5);
      }
      // ---- end of the chain: fold the accumulators into one fp32 gradient vector (all 16 warps) ----
      if (warp == 0 && c == 0) LT_TRACE2(0, 62, 0);
      mbar_wait(g_full, c & 1);
      if (warp == 0 && c == 0) LT_TRACE2(0, 62, 1);
      tc_fence_after();
      const int P = p.P;
      if (warp == 0 && c == 0) LT_TRACE2(0, 62, 2);
      // Fold without atomics (shared-memory float atomics are compare-and-swap loops: the first version spent ~43 000 cycles per item
      // here, 10 % of the kernel). Every (X piece, column d, hidden unit) / (h piece, hidden unit, output) has exactly ONE thread that
      // holds all its accumulator columns (both tile streams, the hi and lo delta pieces), so each slot is written once with a plain
      // store; the copy-out adds the pieces. Row strides are odd ([D][H|1], [H][O|1]): at a given instruction the lanes of a warp are
      // different X columns / hidden units, and with the natural strides H / O (64, 16, 2 ...) their stores would share one bank.
      constexpr int NPW = (DKC <= 2) ? 3 : 2;                 // X pieces among the 128 image columns of G1
      const int HS = H | 1, OS = O | 1;
      const int oW2 = NPW * D * HS, oB1 = oW2 + 2 * H * OS, oB2 = oB1 + H;      // then [16 warps][NQ] partial B2 sums
      {
        const int slice = warp >> 2;
        const int m = row_in_tile;
        // G1: TMEM lane m = column m of the X image tile; this warp sums, for its HP/4 hidden units, the dh_hi and dh_lo columns of both streams
        const int piece = m / DW, d = m - piece * DW;
        const bool is_w = (piece < NPW) && (d < D);          // DKC = 1: lanes >= 64 repeat the block and fall out here
        const bool is_b = (m == C::ones_col(D));
        constexpr int CQ = HP / 4;
        float acc[CQ];
#pragma unroll
        for (int j = 0; j < CQ; ++j) acc[j] = 0.f;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          if (b == 0 || n_tiles > 1) {
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
              uint32_t r[CQ];
              tmem_ldN<CQ>(tmem_base + lane_base + (uint32_t)(C::kG1 + b * NG1 + pc * HP + slice * CQ), r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < CQ; ++j) acc[j] += __uint_as_float(r[j]);
            }
          }
        }
        if (is_w || is_b) {
#pragma unroll
          for (int j = 0; j < CQ; ++j) {
            const int hh = slice * CQ + j;
            if (hh < H) {
              if (is_w) sGrad[(piece * D + d) * HS + hh] = acc[j];                 // LG:104: W1 += x^T dh
              else sGrad[oB1 + hh] = -acc[j];                                       // LG:105: B1 += -dh
            }
          }
        }
        // G2[g]: TMEM lane m = [h1 (HP) | h2 (HP)] hidden index; columns 32 g + [do_hi (16) | do_lo (16)]; this warp takes outputs 4 slice ..
        const int hp2 = m < HP ? 0 : 1, hh2 = m - hp2 * HP;
        float acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          if (b == 0 || n_tiles > 1) {
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
              uint32_t r2[4];
              tmem_ld4(tmem_base + lane_base + (uint32_t)(C::kG2 + b * 64 + 32 * b + pc * ON + slice * 4), r2);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 4; ++j) acc2[j] += __uint_as_float(r2[j]);
            }
          }
        }
        if (m < 2 * HP && hh2 < H) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (slice * 4 + j < O) sGrad[oW2 + (hp2 * H + hh2) * OS + slice * 4 + j] = acc2[j];   // LG:102: W2 += h^T do
        }
#pragma unroll
        for (int o = 0; o < NQ; ++o) {
          const float sres = warp_sum(accB2[o]);
          if (lane == 0) sGrad[oB2 + warp * NQ + o] = sres;                         // LG:103: B2 += -do (summed over the warps below)
        }
      }
      tc_fence_before();
      epi_sync();
      if (warp == 0 && c == 0) LT_TRACE2(0, 62, 3);
      float* dst = p.part + ((int64_t)w.chunk * p.R + pos) * P;
      for (int i = threadIdx.x; i < P; i += 32 * kEpiWarps) {
        int k = i;
        float v = 0.f;
        if (k < D * H) {
          const int src = (k / H) * HS + (k % H);
#pragma unroll
          for (int pc = 0; pc < NPW; ++pc) v += sGrad[pc * D * HS + src];
        } else if ((k -= D * H) < H * O) {
          const int src = oW2 + (k / O) * OS + (k % O);
          v = sGrad[src] + sGrad[src + H * OS];
        } else if ((k -= H * O) < H) {
          v = sGrad[oB1 + k];
        } else {
          k -= H;                                  // output k: warps with hf == k / NQ hold its partial sums
          const int hfo = k / NQ, j = k - hfo * NQ;
          for (int wq = 0; wq < kEpiWarps; ++wq)
            if (((wq >> 2) & 1) == hfo) v -= sGrad[oB2 + wq * NQ + j];
        }
        dst[i] = v;
      }
      epi_sync();
      if (warp == 0 && c == 0) LT_TRACE2(0, 62, 4);
      fence_async_smem();     // Hs goes back to being an MMA operand
      if (lane == 0) mbar_arrive(g_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kCols));
  }
}

template <int HP, int DKC>
static int launch_variant(const CUtensorMap& ma, const CUtensorMap& mb1, const CUtensorMap& mw2, const CUtensorMap& mw2t, const Params& a, dim3 grid,
                          cudaStream_t st) {
  using C = Cfg<HP, DKC>;
  constexpr size_t smem = C::smem_bytes();
  if constexpr (smem > 227 * 1024) {
    set_error("tcgen05 Langevin: variant HP=%d DKC=%d does not fit in shared memory", HP, DKC);
    return SAPT_ERR_UNSUPPORTED;
  } else {
    auto k = langevin_tc_kernel<HP, DKC>;
    SAPT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kThreads, smem, st>>>(ma, mb1, mw2, mw2t, a);
    SAPT_CHECK_LAUNCH();
    return SAPT_OK;
  }
}
}  // namespace lt

// the tensor-core batch-Langevin path covers O <= 16, H <= 64, D <= 64; for D > 32 it needs a spare column in the
// x1 piece of the operand image for the column of ones (D not a multiple of 16)
bool langevin_tc_supported(const sapt_problem* p) {
  if (!p->tc_ok || p->O > 16 || p->H > 64 || p->D > 64) return false;
  const int dkc = p->D <= 16 ? 1 : (p->D + 15) / 16;
  if (dkc >= 3 && p->D == 16 * dkc) return false;
  return true;
}

int launch_langevin_tc(sapt_problem* p, const float* theta, float* theta_gd, int64_t ld, int R, const int32_t* rows, const int32_t* n_rows,
                       double lr_scale, cudaStream_t st) {
  using namespace lt;
  const DataSet& ds = p->set[0];
  const int HP = p->H <= 16 ? 16 : (p->H <= 32 ? 32 : 64);
  const int DKC = p->D <= 16 ? 1 : (p->D + 15) / 16;
  const int KROW = ((48 * DKC + 16 + 63) / 64) * 64;
  const int P = p->P;
  const int64_t n_tiles = ds.n_pad / kTileM;
  // Work items = (chain, row chunk), dealt round-robin to one CTA per SM inside the kernel (the number of active
  // chains is a device-side value). ~16 items per CTA at the host-side bound R keep the load balanced to a few
  // percent even when only half the chains take a Langevin step; chunks stay >= 32 tiles so that the per-item
  // cost (weight panels, folding the accumulators) is amortised.
  int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(1, n_tiles / 32), ((int64_t)16 * p->num_sms + R - 1) / R));
  const int tiles_per_chunk = (int)((n_tiles + n_chunks - 1) / n_chunks);
  n_chunks = (n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;

  auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  const size_t b1_bytes = (size_t)R * HP * KROW * 2, w2_bytes = (size_t)R * N2 * 64 * 2, w2t_bytes = (size_t)R * HP * 64 * 2;
  const size_t bias_bytes = (size_t)R * ON * sizeof(float), part_bytes = (size_t)n_chunks * R * P * sizeof(float);
  int rc = ensure_workspace(p, up(part_bytes) + up(b1_bytes) + up(w2_bytes) + up(w2t_bytes) + up(bias_bytes));
  if (rc) return rc;
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->ws);
  float* part = reinterpret_cast<float*>(ws);
  __nv_bfloat16* b1img = reinterpret_cast<__nv_bfloat16*>(ws + up(part_bytes));
  __nv_bfloat16* w2img = reinterpret_cast<__nv_bfloat16*>(ws + up(part_bytes) + up(b1_bytes));
  __nv_bfloat16* w2timg = reinterpret_cast<__nv_bfloat16*>(ws + up(part_bytes) + up(b1_bytes) + up(w2_bytes));
  float* bias2 = reinterpret_cast<float*>(ws + up(part_bytes) + up(b1_bytes) + up(w2_bytes) + up(w2t_bytes));

  if ((rc = launch_tc2_pack(theta, ld, rows, n_rows, R, p->D, p->H, p->O, HP, ON, 1, DKC, KROW, DKC >= 3 ? 1 : 0, b1img, w2img, bias2, R, st))) return rc;
  lt_pack_w2t_kernel<<<R, 256, 0, st>>>(theta, ld, rows, n_rows, R, p->D, p->H, p->O, HP, w2timg);
  SAPT_CHECK_LAUNCH();
  CUtensorMap mb1, mw2, mw2t;
  if ((rc = tc_make_map_2d(&mb1, b1img, (uint64_t)KROW, (uint64_t)R * HP, (uint32_t)HP))) return rc;
  if ((rc = tc_make_map_2d(&mw2, w2img, 64, (uint64_t)R * N2, (uint32_t)N2))) return rc;
  if ((rc = tc_make_map_2d(&mw2t, w2timg, 64, (uint64_t)R * HP, (uint32_t)HP))) return rc;
  Params a{};
  a.y = ds.y; a.bias2 = bias2; a.n_rows = n_rows; a.R = R; a.D = p->D; a.H = p->H; a.O = p->O; a.fold_bias = (p->H < HP) ? 1 : 0;
  a.tiles_per_chunk = tiles_per_chunk; a.n_tiles = n_tiles; a.n_chunks = (int)n_chunks; a.part = part; a.P = P;
  dim3 grid((unsigned)std::min<int64_t>(p->num_sms, (int64_t)R * n_chunks));
  static long long* d_trace = nullptr;
  const bool want_trace = getenv("SAPT_LT_TRACE") != nullptr;
  if (want_trace && !d_trace) { cudaMalloc(&d_trace, 4 * 64 * 8 * sizeof(long long)); }
  if (want_trace) { cudaMemsetAsync(d_trace, 0, 4 * 64 * 8 * sizeof(long long), st); a.trace = d_trace; }
  const CUtensorMap& ma = *reinterpret_cast<const CUtensorMap*>(ds.tmap);
  const int DKX = (HP == 64 && p->O <= 8) ? DKC + 8 : DKC;   // narrow output phase, see Cfg
#define SAPT_LT(H_, K_) if (HP == H_ && DKX == K_) rc = launch_variant<H_, K_>(ma, mb1, mw2, mw2t, a, grid, st); else
  SAPT_LT(16, 1) SAPT_LT(16, 2) SAPT_LT(16, 3) SAPT_LT(16, 4) SAPT_LT(32, 1) SAPT_LT(32, 2) SAPT_LT(32, 3) SAPT_LT(32, 4)
  SAPT_LT(64, 1) SAPT_LT(64, 2) SAPT_LT(64, 3) SAPT_LT(64, 4) SAPT_LT(64, 9) SAPT_LT(64, 10) SAPT_LT(64, 11) SAPT_LT(64, 12) { set_error("tcgen05 Langevin: no variant for HP=%d DKC=%d", HP, DKC); rc = SAPT_ERR_UNSUPPORTED; }
#undef SAPT_LT
  if (rc) return rc;
  if (want_trace) {
    std::vector<long long> h(4 * 64 * 8);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long t0 = h[(2 * 64 + 0) * 8 + 0];
    const char* roles[4] = {"epi0", "epi1", "mma0", "mma1"};
    for (int r = 0; r < 4; ++r) {
      printf("trace %s e0 of visits 0..63:", roles[r]);
      for (int n = 0; n < 64; ++n) { long long v = h[(r * 64 + n) * 8 + 0]; printf(" %lld", v ? v - t0 : -1); }
      printf("\n");
    }
    for (int n = 6; n < 64; n += 8)
      for (int r = 0; r < 4; ++r) {
        printf("trace %s visit %2d:", roles[r], n);
        for (int e = 0; e < 8; ++e) { long long v = h[(r * 64 + n) * 8 + e]; printf(" %8lld", v ? v - t0 : -1); }
        printf("\n");
      }
  }
  return launch_langevin_batch_finalize(part, (int)n_chunks, R, P, theta, theta_gd, ld, rows, n_rows, lr_scale, st);
}

}  // namespace sapt
