// tcgen05 / TMEM / TMA likelihood kernel: BOTH layers of every chain on the tensor cores (sm_100a only).
//
// Replaces, for all chains a GPU hosts: ptReplica.likelihood_func (_revamp.py:391-402) ->
// Network.evaluate_proposal (:155-174) -> ForwardPass (:70-76) + softmax (:129-131).
//
//   MMA-1  D1[128 rows, UC chains x HP] = X . W1'     smem x smem -> TMEM     (kind::f16, bf16 in, fp32 acc)
//   phase 1 (thread = data row):  tcgen05.ld D1 -> h = sigmoid -> h = h1 + h2 (bf16 pieces) ->
//           tcgen05.st HV[chain] = [h1 | h2]          (the A operand of MMA-2, never leaves TMEM)
//   MMA-2  D2[128 rows, N2] = H . W2'                 TMEM x smem -> TMEM, per chain, into its own D2 columns
//   phase 2: tcgen05.ld D2 -> sigmoid -> softmax / argmax over exactly O outputs -> fp64 log-likelihood sums
//
// fp32-grade accuracy from bf16 tensor cores: every fp32 operand is split into bf16 pieces that sum to it
// (x = x1+x2+x3, w = w1+w2+w3 exactly; h = h1+h2 to 2^-16, unbiased) and the significant cross terms are
// accumulated by pointing successive MMA K-steps at different (A chunk, B chunk) pairs — each piece is
// stored once. MMA-1 (K = (6 DKC + 1) x 16, DKC = 16-feature chunks per piece): (x1,w1) (x1,w2) (x2,w1) (x1,w3)
// (x2,w2) (x3,w1) (1,b): dropped terms O(2^-24). MMA-2: h1.[w1|w2|w3] and h2.[w1|w2] with the weight pieces side by
// side along N (partial sums per output, added by the epilogue), so each activation piece is read from TMEM once.
// Weights are pre-scaled by -log2(e) (biases +log2(e)), so an accumulator is t = -log2(e) z and
// sigmoid(z) = 1 / (1 + 2^t). When H < HP the output bias rides in MMA-2: the spare hidden slot's activation is
// sigmoid(0) = 1/2 and its W2 row holds twice the bias.
//
// The SIMT side is bound by the SFU (MUFU) pipe and by instruction issue, so reciprocals are batched four at a
// time (1/a = (1/(abcd)) * bcd: one MUFU.RCP + 9 FMUL per four sigmoids), the bf16 split uses integer ops
// (truncate / add-half + PRMT) instead of F2FP conversions, which also issue on the SFU pipe, and the sigmoid
// groups of phase 1 are kept in order (see phase 1) so that SFU and FP32/integer work alternate finely.
//
// Geometry = TcCfg<HP, ON, DKC> (HP = 16/32/64 padded hidden width, ON = 16/32 padded outputs, DKC = ceil(D/16)):
// CTA = (row chunk, CH chains), CH and the A-ring depth from the 224 KB shared-memory budget; a unit = UC chains
// (4, or 2 for HP = 64). Warp 16 lane 0: TMA producer (A ring, KBA boxes per tile; B1, W2 panels once). Warp 17:
// TMEM alloc + single-thread MMA issue. Warps 0-15: epilogue; warp (quarter q, index e) owns TMEM lanes 32q.. and
// one chain of every unit (for HP = 64: half the hidden units of a chain; the two halves alternate on phase 2).
// TMEM (512 columns): D1 (UC x HP) | HV (UC x HP) | D2 (UC x N2, doubled when it fits). The epilogue is software
// pipelined (phase 1 of unit u+1 before phase 2 of unit u) so that every MMA has a full unit of slack before its
// result is awaited.
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <type_traits>
#include <vector>

#include <cuda_fp16.h>

#include "sapt_device.cuh"
#include "sapt_internal.h"
#include "sapt_tc_ptx.cuh"

// First-layer operands as TWO fp16 pieces (x = x_hi + x_lo, w = w_hi + w_lo: 22 significant bits each) instead of three bf16
// pieces: MMA-1 then needs the products (x_hi,w_hi) (x_hi,w_lo) (x_lo,w_hi) — 3 instead of 6 K passes — for a dropped term of
// 2^-22 |x w|, and an operand row is 1/3 shorter (D = 54: 2 instead of 3 K blocks per row and stage). fp16's range is no
// obstacle: features are z-scored, weights pre-scaled by log2(e) saturate at +-65504, and subnormals keep the absolute error
// of a small operand below 2^-25. 0: the three-piece bf16 operands of round 1 (kept for A/B measurements).
#ifndef SAPT_X_FP16
#define SAPT_X_FP16 1
#endif

namespace sapt {

namespace v2 {
constexpr int kTileM = 128;
constexpr int kEpiWarps = 16;           // four per TMEM lane quarter
#ifndef SAPT_MAXNREG
#define SAPT_MAXNREG 0     // > 0: registers per epilogue thread after setmaxnreg (the producer warp group gives its share away)
#endif
// With SAPT_MAXNREG the two single-lane roles sit in a full warp group of their own (two idle warps added) so that setmaxnreg, a
// warp-group instruction, can move registers from it to the four epilogue warp groups.
constexpr int kThreads = 32 * ((SAPT_MAXNREG ? 4 : 2) + kEpiWarps);
// The SMSP arbiter favours the highest warp id: the two single-thread roles that feed everyone else
// (TMA producer, MMA issuer) sit above the epilogue warps so they are never starved by them. (Folding the
// producer into the MMA thread was measured 2 % slower and frees no registers: warps are allocated in fours.)
constexpr int kTmaWarp = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + 1;
constexpr int kSmemLimit = 224 * 1024;
}  // namespace v2

// Compile-time geometry of one kernel variant. HP = padded hidden width (16/32/64), ON = padded output width
// (16/32), DKC = 16-feature chunks per x piece (D <= 16 DKC).
// The third parameter DK packs two things: DK & 7 = DKC, and DK >= 8 = "bias in x1": when D is not a multiple of 16 the
// first spare column of the x1 piece of the operand image holds 1.0 (tc_prepare_dataset) and the bias pieces b1, b2, b3
// sit in that column of the w1, w2, w3 pieces of B1, so MMA-1 needs no bias K-step and no bias chunk: one MMA less per
// unit, and for D > 48 one 128-byte K block less per row (A stage 64 -> 48 KB, B1 panel 3/4), which doubles the
// chains a CTA can keep at HP = 64.
// The second parameter ONX is the padded output width ON (16 / 32), or 8 / 12 = "ON = 16 geometry, O <= 8 / 12": phase 2 then
// reads 8 / 12 instead of 16 accumulator columns per partial sum (24 / 36 instead of 48 registers). Instantiated for HP = 64, where the
// spills of the wide variant cost most: with 225 KB of shared memory in use the L1 that backs local memory is a few KB
// (ncu: local loads hit 4 %).
template <int HP, int ONX, int DK>
struct TcCfg {
  static constexpr int ON = (ONX == 8 || ONX == 12) ? 16 : ONX;
  static constexpr int NR = (ONX == 32) ? 20 : ONX;      // accumulator columns phase 2 reads per partial sum
  static constexpr bool BIASX = DK >= 8;
  static constexpr int DKC = DK & 7;
  static constexpr int HALVES = (HP == 64) ? 2 : 1;       // epilogue warps sharing one chain's hidden units
  static constexpr int UC = 4 / HALVES;                   // chains per unit (one MMA-1)
  static constexpr int HW = HP / HALVES;                  // hidden units per epilogue warp
  static constexpr int N1 = UC * HP;                      // MMA-1 N (<= 128)
  static constexpr bool XF16 = (SAPT_X_FP16 != 0);        // first-layer operands: two fp16 pieces, else three bf16 pieces
  static constexpr int XP = XF16 ? 2 : 3;                 // pieces per operand
  static constexpr int NPROD = XF16 ? 3 : 6;              // piece products accumulated by MMA-1
  static constexpr int KA = 16 * XP * DKC + (BIASX ? 0 : 16);  // used elements per A / B1 row: [pieces | bias chunk]
  static constexpr int KBA = (KA + 63) / 64;              // 128-byte K blocks per row
  static constexpr int N2 = (ON == 16) ? 48 : 64;         // D2 columns per chain
  static constexpr int PARTS = N2 / ON;
  static constexpr int W2ROWS = 3 * ON;                   // W2 image rows per chain (pieces w1, w2, w3), K = HP <= 64
  static constexpr uint32_t kAStage = KBA * 128 * 128;    // bytes per A stage
  static constexpr int STAGES = (KBA == 1) ? 4 : ((KBA == 2) ? 3 : 2);
  static constexpr uint32_t smem_bytes(int upt) {
    const int ch = UC * upt;
    return 1024u + (uint32_t)ch * HP * KBA * 128 + STAGES * kAStage + (uint32_t)ch * W2ROWS * 128 + (uint32_t)ch * ON * 4 + 2048u +
           (uint32_t)ch * 4 * HALVES * 3 * 8;
  }
  static constexpr int pick_upt() {
    for (int upt = 4; upt >= 1; upt >>= 1)
      if (UC * upt * HP <= 256 && smem_bytes(upt) <= (uint32_t)v2::kSmemLimit) return upt;
    return 0;
  }
  static constexpr int UPT = pick_upt();                  // units per row tile (0: variant does not fit)
  static constexpr int CH = UC * (UPT > 0 ? UPT : 1);     // chains per CTA
  static constexpr int kW2BoxMax = (ON == 16) ? 4 : 2;    // chains per W2 TMA box: box rows <= 256, divides CH
  static constexpr int kW2Box = kW2BoxMax < CH ? kW2BoxMax : CH;
  static constexpr int kD2Bufs = (2 * N1 + 2 * UC * N2 <= 512) ? 2 : 1;
  static_assert(2 * N1 + kD2Bufs * UC * N2 <= 512, "TMEM budget");
  static_assert(ON <= HP, "D2 partial sums reuse the chain's hidden width");
  static_assert(CH % kW2Box == 0 && kW2Box * W2ROWS <= 256, "W2 TMA box");
};

struct Tc2Params {
  const int32_t* y; int64_t n_pad;
  const float* bias2;            // [pos][ON] log2(e) * B2
  const int32_t* n_rows; int R;
  int O; int fold_bias;          // fold_bias: H < HP, the output bias is row H of the W2 image
  uint32_t zero;                 // 0 at run time, unknown to the compiler: glues ordering-only dependencies (see phase 1)
  int tiles_per_chunk; int64_t n_tiles;
  double* partL; int32_t* partNc; int32_t* partSq;
  long long* trace;              // development (-DSAPT_TC2_TRACE): clock64 stamps of CTA (0, 0)
};
// development trace: event e of unit u (first 64) of a role; roles 0..3 = epilogue warps 0, 4, 8, 12 (quarter 0, e = 0..3), 4 = MMA warp
#ifdef SAPT_TC2_TRACE
#define TC2_TRACE(role, u, e) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (u) < 64) p.trace[((role) * 64 + (u)) * 8 + (e)] = clock64(); } while (0)
#else
#define TC2_TRACE(role, u, e) do { } while (0)
#endif

// ---- per-step operand packing ------------------------------------------------------------------------
// B1: [group][CH*HP rows][KBA*64] : row (chain, h) = [w1 (16 DKC) | w2 | w3 | b1 b2 b3 0..], w = -log2(e) W1[:, h]
// W2: [group][CH][3 ON rows][64]  : row (p, o) = piece p of -log2(e) W2[:, o] over the HP hidden units (rest 0).
//     MMA-2 reads rows [0, N2) as one operand against h1 and rows [0, 2 ON) again against h2.
__global__ void tc2_pack_kernel(const float* theta, int64_t ld, const int32_t* rows, const int32_t* n_rows, int R, int D, int H, int O,
                                int HP, int ON, int CH, int DKC, int KROW, int biasx, int xf16, __nv_bfloat16* b1img, __nv_bfloat16* w2img, float* bias2) {
  constexpr int KW = 64;
  const int nact = n_rows ? *n_rows : R;
  const int pos = blockIdx.x;
  const int n_pos_pad = ((nact + CH - 1) / CH) * CH;
  if (pos >= n_pos_pad) return;
  const int grp = pos / CH, within = pos - grp * CH;
  __nv_bfloat16* brow0 = b1img + ((int64_t)grp * CH * HP + (int64_t)within * HP) * KROW;
  __nv_bfloat16* wrow0 = w2img + ((int64_t)grp * CH + within) * (3 * ON) * KW;
  float* bb = bias2 + (int64_t)pos * ON;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  if (pos >= nact) {
    for (int i = threadIdx.x; i < HP * KROW; i += blockDim.x) brow0[i] = zero;
    for (int i = threadIdx.x; i < 3 * ON * KW; i += blockDim.x) wrow0[i] = zero;
    for (int i = threadIdx.x; i < ON; i += blockDim.x) bb[i] = 0.f;
    return;
  }
  const int row = rows ? rows[pos] : pos;
  const float* src = theta + (int64_t)row * ld;
  const int DW = 16 * DKC;                      // features per piece (padded)
  if (xf16) {
    // two fp16 pieces per weight: row (chain, h) = [w_hi (DW) | w_lo (DW) | bias chunk | 0..]; the bias pieces sit in column D of
    // the two weight pieces (they meet the 1.0 in column D of the x_hi piece) or in the first two slots of the bias chunk
    __half* hrow0 = reinterpret_cast<__half*>(brow0);
    const __half hz = __float2half_rn(0.f);
    for (int i = threadIdx.x; i < HP * KROW; i += blockDim.x) hrow0[i] = hz;
    __syncthreads();
    auto split_h = [](float v, __half& hi, __half& lo) {
      v = fminf(fmaxf(v, -65504.0f), 65504.0f);
      hi = __float2half_rn(v);
      lo = __float2half_rn(v - __half2float(hi));
    };
    for (int i = threadIdx.x; i < H * (D + 1); i += blockDim.x) {
      const int h = i / (D + 1), d = i - h * (D + 1);
      __half hi, lo;
      __half* dst = hrow0 + (int64_t)h * KROW;
      if (d < D) {
        split_h(-kLog2e * src[d * H + h], hi, lo);
        dst[d] = hi; dst[DW + d] = lo;
      } else {
        split_h(kLog2e * src[D * H + H * O + h], hi, lo);
        if (biasx) { dst[D] = hi; dst[DW + D] = lo; }
        else { dst[2 * DW] = hi; dst[2 * DW + 1] = lo; }
      }
    }
  } else {
  for (int i = threadIdx.x; i < HP * DW; i += blockDim.x) {
    const int h = i / DW, d = i - h * DW;
    __nv_bfloat16 p[3] = {zero, zero, zero};
    if (h < H && d < D) split3(-kLog2e * src[d * H + h], p[0], p[1], p[2]);
    __nv_bfloat16* dst = brow0 + (int64_t)h * KROW;
    dst[d] = p[0]; dst[DW + d] = p[1]; dst[2 * DW + d] = p[2];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < HP * (KROW - 3 * DW); i += blockDim.x) {   // bias chunk + tail padding
    const int h = i / (KROW - 3 * DW), q = i - h * (KROW - 3 * DW);
    __nv_bfloat16 v = zero;
    if (!biasx && h < H && q < 3) { __nv_bfloat16 b[3]; split3(kLog2e * src[D * H + H * O + h], b[0], b[1], b[2]); v = b[q]; }
    brow0[(int64_t)h * KROW + 3 * DW + q] = v;
  }
  if (biasx) {   // bias pieces in column D of the w1 / w2 / w3 pieces: they meet the 1.0 in column D of the x1 piece
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
      __nv_bfloat16 b[3];
      split3(kLog2e * src[D * H + H * O + h], b[0], b[1], b[2]);
      __nv_bfloat16* dst = brow0 + (int64_t)h * KROW;
      dst[D] = b[0]; dst[DW + D] = b[1]; dst[2 * DW + D] = b[2];
    }
  }
  }
  for (int i = threadIdx.x; i < ON * KW; i += blockDim.x) {
    const int o = i / KW, h = i - o * KW;
    __nv_bfloat16 p[3] = {zero, zero, zero};
    if (o < O && h < H) split3(-kLog2e * src[D * H + h * O + o], p[0], p[1], p[2]);
    else if (o < O && h == H && H < HP) split3(2.0f * kLog2e * src[D * H + H * O + H + o], p[0], p[1], p[2]);   // bias x 1/sigmoid(0)
    wrow0[(int64_t)o * KW + h] = p[0];
    wrow0[(int64_t)(ON + o) * KW + h] = p[1];
    wrow0[(int64_t)(2 * ON + o) * KW + h] = p[2];
  }
  for (int i = threadIdx.x; i < ON; i += blockDim.x) bb[i] = (i < O) ? kLog2e * src[D * H + H * O + H + i] : 0.f;   // used when H == HP (no spare hidden slot)
}

int launch_tc2_pack(const float* theta, int64_t ld, const int32_t* rows, const int32_t* n_rows, int R, int D, int H, int O, int HP, int ON, int CH,
                    int DKC, int KROW, int biasx, __nv_bfloat16* b1img, __nv_bfloat16* w2img, float* bias2, int n_blocks, cudaStream_t st) {
  tc2_pack_kernel<<<n_blocks, 256, 0, st>>>(theta, ld, rows, n_rows, R, D, H, O, HP, ON, CH, DKC, KROW, biasx, 0, b1img, w2img, bias2);   // three bf16 pieces
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

// four sigmoids with one reciprocal:  s_i = 1 / a_i,  a_i = 1 + 2^min(t_i, 30)  (clamp: |error| < 1e-9,
// keeps the product of four below 2^120)
__device__ __forceinline__ void sigmoid4(float t0, float t1, float t2, float t3, float& s0, float& s1, float& s2, float& s3) {
  const float a0 = 1.0f + ex2_approx(fminf(t0, 30.0f));
  const float a1 = 1.0f + ex2_approx(fminf(t1, 30.0f));
  const float a2 = 1.0f + ex2_approx(fminf(t2, 30.0f));
  const float a3 = 1.0f + ex2_approx(fminf(t3, 30.0f));
  const float ab = a0 * a1, cd = a2 * a3;
  const float r = rcp_approx(ab * cd);
  const float rab = r * cd, rcd = r * ab;
  s0 = rab * a1; s1 = rab * a0; s2 = rcd * a3; s3 = rcd * a2;
}

// same with the clamp limit in a register (see phase 1)
__device__ __forceinline__ void sigmoid4_lim(float t0, float t1, float t2, float t3, float lim, float& s0, float& s1, float& s2,
                                             float& s3) {
  const float a0 = 1.0f + ex2_approx(fminf(t0, lim));
  const float a1 = 1.0f + ex2_approx(fminf(t1, lim));
  const float a2 = 1.0f + ex2_approx(fminf(t2, lim));
  const float a3 = 1.0f + ex2_approx(fminf(t3, lim));
  const float ab = a0 * a1, cd = a2 * a3;
  const float r = rcp_approx(ab * cd);
  const float rab = r * cd, rcd = r * ab;
  s0 = rab * a1; s1 = rab * a0; s2 = rcd * a3; s3 = rcd * a2;
}
#ifndef SAPT_F32X2
#define SAPT_F32X2 1      // packed fp32 pairs (FADD2 / FMUL2) in the epilogue arithmetic
#endif
#ifndef SAPT_CONST_O
#define SAPT_CONST_O 1    // compile-time output count for the BASELINE shapes (see loglik_tc2_kernel)
#endif
#ifndef SAPT_ELECT
#define SAPT_ELECT 1      // single-thread instructions issued by an elected lane of a converged warp (0: `if (lane == 0)`, round 1)
#endif
#ifndef SAPT_SPLIT_D1_ANY
#define SAPT_SPLIT_D1_ANY 1   // D1 in two halves also for the ordered one-quad groups (HP = 64: 5.39 -> 5.36 ms)
#endif
#ifndef SAPT_C3_HC
#define SAPT_C3_HC 30   // compile-time hidden count of the C3 / C5 specialisations (0: compute all padded slots), for A/B builds
#endif
#ifndef SAPT_C5_HC
#define SAPT_C5_HC 0
#endif
#ifndef SAPT_P1_DEP
#define SAPT_P1_DEP -1    // phase-1 ordering (see phase 1): 0 = groups ordered on a finished sigmoid, 1 = on the group's first 2^t, 2 = software
#endif                    // pipeline over the quads; -1 = per shape (measured: 2 for HP <= 32, 1 with one-quad groups for HP = 64)
#ifndef SAPT_P1_CHAIN
#define SAPT_P1_CHAIN -1  // quads per ordered group in phase 1 (0: let ptxas schedule freely); -1 = per shape
#endif

// same, times a common factor: s_i = scale / a_i (the factor rides on the shared reciprocal)
__device__ __forceinline__ void sigmoid4_scaled(float t0, float t1, float t2, float t3, float scale, float& s0, float& s1, float& s2,
                                                float& s3) {
  const float a0 = 1.0f + ex2_approx(fminf(t0, 30.0f));
  const float a1 = 1.0f + ex2_approx(fminf(t1, 30.0f));
  const float a2 = 1.0f + ex2_approx(fminf(t2, 30.0f));
  const float a3 = 1.0f + ex2_approx(fminf(t3, 30.0f));
  const float ab = a0 * a1, cd = a2 * a3;
  const float r = rcp_approx(ab * cd) * scale;
  const float rab = r * cd, rcd = r * ab;
  s0 = rab * a1; s1 = rab * a0; s2 = rcd * a3; s3 = rcd * a2;
}

// OC: the number of outputs as a compile-time constant (0 = read it from the parameters). With a constant O the output phase has
// no branches and ptxas schedules all its sigmoid groups together; instantiated for the topologies of the BASELINE configs.
// HC: the number of hidden units as a compile-time constant (0 = all HP slots are computed): the padded slots HC.. of the hidden
// layer cost no SFU work — slot HC is the bias slot (activation 1/2, see the header), the others meet zero rows of the W2 image.
// DYN: the CTA walks only the units of a tile that hold a chain (run-time count). For launches whose single chain group is far from
// full — a handful of chains on a large data set, the reference's usual 8-10 replicas, the X-streaming regime — the empty chain
// slots then cost nothing (6 chains at 16-30-10: two of four units). A run-time loop bound costs the full-group kernels 4-6 %, so
// this is a separate set of instantiations that the launcher picks only then.
template <int HP, int ONX, int DK, int OC = 0, int HC = 0, bool DYN = false>
__global__ void __launch_bounds__(v2::kThreads, 1) loglik_tc2_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                     const __grid_constant__ CUtensorMap map_b1,
                                                                     const __grid_constant__ CUtensorMap map_w2, Tc2Params p) {
  using namespace v2;
  using C = TcCfg<HP, ONX, DK>;
  constexpr int DKC = C::DKC, ON = C::ON;
  constexpr int HALVES = C::HALVES, UC = C::UC, HW = C::HW, N1 = C::N1, KBA = C::KBA, N2 = C::N2, PARTS = C::PARTS;
  constexpr int UPT = C::UPT, CH = C::CH, STAGES = C::STAGES, kD2Bufs = C::kD2Bufs;
  static_assert(UPT >= 1, "variant does not fit in shared memory");
  constexpr uint32_t kAStage = C::kAStage;
  constexpr uint32_t kABlock = 128 * 128;                     // one K block of an A tile: 128 rows x 128 B
  constexpr uint32_t kB1Block = CH * HP * 128;                // one K block of the B1 panel
  constexpr uint32_t kB1Bytes = KBA * kB1Block;
  constexpr uint32_t kW2ChainBytes = C::W2ROWS * 128;
  constexpr uint32_t kW2Bytes = CH * kW2ChainBytes;
  // TMEM columns: [D1 | HV (A operand of MMA-2) | D2 buffers]. D2 of a chain is N2 wide: the partial sums h.w1,
  // h.w2 (, h.w3) land side by side (each A piece is read from TMEM once — an MMA with A in TMEM costs ~64 cycles
  // whatever N is) and are added by the epilogue.
  constexpr int kHvBase = N1;
  constexpr int kD2Base = 2 * N1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB1 = smem;
  uint8_t* sA = sB1 + kB1Bytes;
  uint8_t* sW2 = sA + STAGES * kAStage;
  float* sBias = reinterpret_cast<float*>(sW2 + kW2Bytes);          // [CH][ON]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + CH * ON);
  uint64_t* full_a = bars;                  // [STAGES]
  uint64_t* empty_a = full_a + STAGES;      // [STAGES]
  uint64_t* b_full = empty_a + STAGES;      // [1]
  uint64_t* d1_full = b_full + 1;           // [1]
  uint64_t* d1_empty = d1_full + 1;         // [1]
  // hv_full / d2_full are per (unit parity, chain): with one barrier per chain a fast warp could complete
  // phase u+1 before the waiter has looked at phase u, and a parity wait cannot see a barrier two phases ahead
  uint64_t* hv_full = d1_empty + 1;         // [2][UC]
  uint64_t* d2_full = hv_full + 2 * UC;     // [2][UC]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d2_full + 2 * UC);
  double* red = reinterpret_cast<double*>(tmem_ptr + 2);  // [CH][4 quarters * HALVES][3]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nact = p.n_rows ? *p.n_rows : p.R;
  const int g0 = blockIdx.y * CH;
  if (g0 >= nact) return;
  const int64_t t_begin = (int64_t)blockIdx.x * p.tiles_per_chunk;
  const int64_t t_end = min(t_begin + (int64_t)p.tiles_per_chunk, p.n_tiles);
  const int n_my_tiles = (int)max((int64_t)0, t_end - t_begin);
  const int upt = DYN ? (min(CH, nact - g0) + UC - 1) / UC : UPT;   // units per tile that are walked (the same in every role)
  const int n_units = n_my_tiles * upt;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
    mbar_init(b_full, 1);
    mbar_init(d1_full, 1);
    mbar_init(d1_empty, kEpiWarps);
    for (int c = 0; c < 2 * UC; ++c) { mbar_init(&hv_full[c], 4 * HALVES); mbar_init(&d2_full[c], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    const float* src = p.bias2 + (int64_t)g0 * ON;
    const int ng = min(CH, nact - g0);
    for (int i = threadIdx.x; i < CH * ON; i += kThreads) sBias[i] = (i < ng * ON) ? src[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_hv = tmem_base + kHvBase;

#if SAPT_MAXNREG
  if (warp >= kEpiWarps) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SAPT_MAXNREG));
#endif
  if (warp == kTmaWarp) {
    // ================= TMA producer (the warp stays converged, one elected lane issues: see elect_one) =================
    if (SAPT_ELECT ? elect_one() : (lane == 0)) {
      mbar_expect_tx(b_full, kB1Bytes + kW2Bytes);
#pragma unroll
      for (int kb = 0; kb < KBA; ++kb) tma_load_2d(&map_b1, b_full, sB1 + kb * kB1Block, kb * 64, blockIdx.y * CH * HP);
#pragma unroll
      for (int q = 0; q < CH / C::kW2Box; ++q)  // boxes of kW2Box chains x 3 ON rows
        tma_load_2d(&map_w2, b_full, sW2 + q * C::kW2Box * kW2ChainBytes, 0, (blockIdx.y * CH + q * C::kW2Box) * C::W2ROWS);
    }
    __syncwarp();
    for (int it = 0; it < n_my_tiles; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      if (!SAPT_ELECT && lane != 0) break;
      mbar_wait<1000>(&empty_a[s], ph ^ 1);
      if (SAPT_ELECT ? elect_one() : true) {
        mbar_expect_tx(&full_a[s], kAStage);
#pragma unroll
        for (int kb = 0; kb < KBA; ++kb)
          tma_load_2d(&map_a, &full_a[s], sA + s * kAStage + kb * kABlock, kb * 64, (int)((t_begin + it) * kTileM));
      }
      if (SAPT_ELECT) __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer: the warp stays converged (waits, descriptor arithmetic on the uniform datapath), one elected
    // lane issues tcgen05.mma / commit (see elect_one) =================
    if (SAPT_ELECT || lane == 0) {
      // instruction descriptor (kind::f16): D fp32 (bit 4), A / B format in bits 7-9 / 10-12 (0 = fp16, 1 = bf16), N >> 3, M >> 4
      auto idesc = [](int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24); };
      const uint32_t idesc1 = C::XF16 ? ((1u << 4) | ((uint32_t)(N1 >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24)) : idesc(N1);
      const uint32_t idesc2a = idesc(N2), idesc2b = idesc(2 * ON), idesc2c = idesc(ON);
      const uint64_t dA0 = umma_desc_sw128(smem_u32(sA)), dB10 = umma_desc_sw128(smem_u32(sB1)), dW20 = umma_desc_sw128(smem_u32(sW2));
      mbar_wait(b_full, 0);
      auto issue_mma1 = [&](int u) {
        const int it = u / upt, uu = u - it * upt;
        const int s = it % STAGES;
        TC2_TRACE(4, u, 0);
        if (u > 0) mbar_wait(d1_empty, (u - 1) & 1);  // every warp holds D1(u-1) in registers
        if (uu == 0) mbar_wait(&full_a[s], (it / STAGES) & 1);
        TC2_TRACE(4, u, 1);
        tc_fence_after();
        // split product over (piece of x, piece of w): fp16 pieces x_hi w_hi, x_hi w_lo, x_lo w_hi; bf16 pieces x1w1 x1w2 x2w1 x1w3
        // x2w2 x3w1. DKC chunks of 16 features each, then the bias chunk. chunk c = 16 elements = 32 bytes inside K block c/4.
        // Descriptors: the address field is the low 14 bits (address >> 4) and shared memory ends below 256 KB, so the descriptor of
        // base + offset is descriptor(base) + (offset >> 4): one multiply-add per batch, one add of a constant per operand.
        constexpr int pa[6] = {0, 0, 1, 0, 1, 2};
        constexpr int pb[6] = {0, 1, 0, 2, 1, 0};
        const uint64_t a_base = dA0 + (uint64_t)((uint32_t)s * (kAStage >> 4)), b_base = dB10 + (uint64_t)((uint32_t)uu * ((N1 * 128) >> 4));
        auto a_at = [&](int chunk) { return a_base + (uint64_t)(((uint32_t)(chunk >> 2) * kABlock + (uint32_t)(chunk & 3) * 32) >> 4); };
        auto b_at = [&](int chunk) { return b_base + (uint64_t)(((uint32_t)(chunk >> 2) * kB1Block + (uint32_t)(chunk & 3) * 32) >> 4); };
        if (SAPT_ELECT ? elect_one() : true) {
#pragma unroll
          for (int q = 0; q < C::NPROD; ++q)
#pragma unroll
            for (int j = 0; j < DKC; ++j)
              tc_mma_bf16(tmem_base, a_at(pa[q] * DKC + j), b_at(pb[q] * DKC + j), idesc1, (q | j) ? 1u : 0u);
          if constexpr (!C::BIASX) tc_mma_bf16(tmem_base, a_at(C::XP * DKC), b_at(C::XP * DKC), idesc1, 1u);
          if (uu == upt - 1) tc_commit(&empty_a[s]);
          tc_commit(d1_full);
        }
        if (SAPT_ELECT) __syncwarp();
      };
      if (n_units > 0) issue_mma1(0);
      for (int u = 0; u < n_units; ++u) {
        const int uu = u % upt;
        // D1(u+1) first: its trigger (all warps loaded D1(u)) precedes the HV(u) hand-over in every warp
        if (u + 1 < n_units) issue_mma1(u + 1);
#pragma unroll
        for (int c = 0; c < UC; ++c) {
          mbar_wait(&hv_full[(u & 1) * UC + c], (u >> 1) & 1);
          if (c == 0) TC2_TRACE(4, u, 2);
          if (c == UC - 1) TC2_TRACE(4, u, 3);
          tc_fence_after();
          const uint32_t d2 = tmem_base + (uint32_t)(kD2Base + (u % kD2Bufs) * (UC * N2) + c * N2);
          const uint32_t hv = tmem_hv + (uint32_t)(c * HP);
          const uint64_t wb = dW20 + (uint64_t)((uint32_t)(uu * UC + c) * (kW2ChainBytes >> 4));
          // ON=16: D2[:, 0:3ON] = h1.[w1|w2|w3], D2[:, 0:2ON] += h2.[w1|w2]
          // ON=32: D2[:, 0:2ON] = h1.[w1|w2],    D2[:, 0:ON] += h1.w3,  D2[:, 0:2ON] += h2.[w1|w2]   (h piece = HP/2 columns)
          if (SAPT_ELECT ? elect_one() : true) {
#pragma unroll
            for (int half = 0; half < HP / 16; ++half)
              tc_mma_bf16_ts(d2, hv + (uint32_t)(half * 8), wb + (uint64_t)(half * 2), idesc2a, half ? 1u : 0u);
            if constexpr (ON == 32) {
#pragma unroll
              for (int half = 0; half < HP / 16; ++half)
                tc_mma_bf16_ts(d2, hv + (uint32_t)(half * 8), wb + (uint64_t)((2 * ON * 128) >> 4) + (uint64_t)(half * 2), idesc2c, 1u);
            }
#pragma unroll
            for (int half = 0; half < HP / 16; ++half)
              tc_mma_bf16_ts(d2, hv + (uint32_t)(HP / 2 + half * 8), wb + (uint64_t)(half * 2), idesc2b, 1u);
            tc_commit(&d2_full[(u & 1) * UC + c]);
          }
          if (SAPT_ELECT) __syncwarp();
        }
      }
    }
  } else if (warp < kEpiWarps) {
    // ================= epilogue warps =================
    // warp (quarter q, index e): TMEM lanes 32q.., chain c = e / HALVES of every unit, hidden units
    // [hf*HW, hf*HW + HW) with hf = e % HALVES. With HALVES = 2 the two warps of a chain alternate on phase 2.
    const int quarter = warp & 3;
    const int e = warp >> 2;
    const int c = e / HALVES, hf = e % HALVES;
    const int O = OC ? OC : p.O;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    double accL[UPT];
    int accNc[UPT], accSq[UPT];
#pragma unroll
    for (int a = 0; a < UPT; ++a) { accL[a] = 0.0; accNc[a] = 0; accSq[a] = 0; }
    // Software pipeline over units: phase 1 of unit u+1 (sigmoid + split, SIMT) runs while the tensor pipe
    // computes MMA-2 of unit u, so the warp never waits for D2 right after handing over its activations.
    uint32_t hv[HW];  // [h1 pairs (HW/2) | h2 pairs (HW/2)] of the unit whose phase 1 ran last
    const bool trw = (quarter == 0) && (hf == 0);
    auto phase1 = [&](int u) {
      if (trw) TC2_TRACE(c, u, 0);
      mbar_wait(d1_full, u & 1);
      if (trw) TC2_TRACE(c, u, 1);
      tc_fence_after();
      uint32_t zr[HW];
      const uint32_t taddr = tmem_base + lane_base + (uint32_t)(c * HP + hf * HW);
      // (D1 in two halves when the quads are software pipelined: the second load's latency hides behind the first quads)
      constexpr int kDep = (SAPT_P1_DEP >= 0) ? SAPT_P1_DEP : ((HP == 64) ? 1 : 2);
      constexpr int kChain = (SAPT_P1_CHAIN >= 0) ? SAPT_P1_CHAIN : ((kDep == 1) ? 1 : 2);
      constexpr bool kPipe = (kDep == 2) && (SAPT_F32X2 != 0);
      constexpr bool kSplitD1 = (kPipe || (SAPT_SPLIT_D1_ANY && kChain == 1 && SAPT_F32X2)) && (HW == 32);
      auto d1_loaded = [&]() {
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d1_empty);  // D1 is in registers: the next MMA-1 may overwrite it
      };
      if constexpr (kSplitD1) {
        tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&zr[0]));
        tmem_ld_wait();
        tmem_ld16(taddr + 16, *reinterpret_cast<uint32_t(*)[16]>(&zr[HW / 2]));
      } else {
        if constexpr (HW == 32) tmem_ld32(taddr, zr); else tmem_ld16(taddr, zr);
        d1_loaded();
      }
      // ptxas would hoist all HW MUFU.EX2 to the front: a warp then monopolises the SFU queue for ~40 x 8 cycles
      // and leaves it idle for the ~200 instructions that follow, and so do its three SMSP neighbours. Chaining
      // the clamp limit of quad group k+1 to a result of group k (times zero: one FFMA per group, not foldable
      // under IEEE rules) keeps the groups in order, so SFU work and FP32/integer work alternate finely.
      if constexpr (kPipe) {
      // Software pipeline over the quads, in program order  A(0) A(1) B(0) A(2) B(1) A(3) ... B(7):  A(j) = clamp + four MUFU.EX2,
      // B(j) = everything else of quad j (sum, shared reciprocal, products, bf16 split: ~27 instructions, one MUFU.RCP). The clamp
      // limit of A(j+2) hangs on the first 2^t of A(j+1) (times zero: keeps the exponential groups in order) AND on the last word
      // of B(j)'s split (and-ed with a run-time zero, one LOP3): so the SFU pipe gets five operations per ~35 instructions all the
      // way through, instead of 40 in the first 140 instructions and none in the ~110 of the split that ptxas sinks to the end —
      // the trace (SAPT_TC2_TRACE) showed the four warps of a scheduler in this phase together: 1 846 cycles for 1 280 of SFU work.
      {
        float ex[HW];
        constexpr int HREAL = (HC > 0 && HALVES == 1) ? HC : HP;       // hidden slots that hold a real unit
        auto stage_a = [&](int j, float lim) {
#pragma unroll
          for (int i = 0; i < 4; ++i) ex[4 * j + i] = (4 * j + i < HREAL) ? ex2_approx(fminf(__uint_as_float(zr[4 * j + i]), lim)) : 0.0f;
        };
        auto stage_b = [&](int j) {
          const f32x2 one = pack2(1.0f, 1.0f);
          const f32x2 a01 = add2(pack2(ex[4 * j], ex[4 * j + 1]), one), a23 = add2(pack2(ex[4 * j + 2], ex[4 * j + 3]), one);
          float a0, a1, a2, a3;
          unpack2(a01, a0, a1); unpack2(a23, a2, a3);
          const float ab = a0 * a1, cd = a2 * a3;
          const float r = rcp_approx(ab * cd);
          float rcd, rab;
          unpack2(mul2(pack2(cd, ab), pack2(r, r)), rcd, rab);
          const f32x2 s10 = mul2(a01, pack2(rcd, rcd));   // (a0 r cd, a1 r cd) = (s1, s0)
          const f32x2 s32 = mul2(a23, pack2(rab, rab));   // (a2 r ab, a3 r ab) = (s3, s2)
          float h0, h1, h2, h3;
          unpack2(s10, h1, h0); unpack2(s32, h3, h2);
          // padded slots: 1/2 in the bias slot, 0 elsewhere (compile-time; the products above fold away with their a = 1)
          auto pad = [&](int i, float& h) { if (4 * j + i >= HREAL) h = (4 * j + i == HREAL) ? 0.5f : 0.0f; };
          pad(0, h0); pad(1, h1); pad(2, h2); pad(3, h3);
          const uint32_t u0 = __float_as_uint(h0), u1 = __float_as_uint(h1), u2 = __float_as_uint(h2), u3 = __float_as_uint(h3);
          hv[2 * j] = __byte_perm(u0, u1, 0x7632);
          hv[2 * j + 1] = __byte_perm(u2, u3, 0x7632);
          float r0, r1, r2, r3;
          r0 = h0 - __uint_as_float(u0 & 0xffff0000u); r1 = h1 - __uint_as_float(u1 & 0xffff0000u);
          r2 = h2 - __uint_as_float(u2 & 0xffff0000u); r3 = h3 - __uint_as_float(u3 & 0xffff0000u);
          if (4 * j + 3 < HREAL) {
            unpack2(sub2(s10, pack2(__uint_as_float(u1 & 0xffff0000u), __uint_as_float(u0 & 0xffff0000u))), r1, r0);
            unpack2(sub2(s32, pack2(__uint_as_float(u3 & 0xffff0000u), __uint_as_float(u2 & 0xffff0000u))), r3, r2);
          }
          hv[HW / 2 + 2 * j] = __byte_perm(__float_as_uint(r0) + 0x8000u, __float_as_uint(r1) + 0x8000u, 0x7632);
          hv[HW / 2 + 2 * j + 1] = __byte_perm(__float_as_uint(r2) + 0x8000u, __float_as_uint(r3) + 0x8000u, 0x7632);
        };
        constexpr int NQD = HW / 4;
        stage_a(0, 30.0f);
        if constexpr (NQD > 1) stage_a(1, fmaf(ex[0], 0.0f, 30.0f));
#pragma unroll
        for (int j = 0; j < NQD; ++j) {
          stage_b(j);
          if (kSplitD1 && j + 2 == NQD / 2) d1_loaded();      // A(NQD / 2) is the first stage that reads the second half
          if (j + 2 < NQD)
            stage_a(j + 2, __uint_as_float(__float_as_uint(fmaf(ex[4 * (j + 1)], 0.0f, 30.0f)) | ((hv[HW / 2 + 2 * j] ^ hv[HW / 2 + 2 * j + 1]) & p.zero)));
        }
      }
      } else {
      float lim = 30.0f;
#pragma unroll
      for (int j = 0; j < HW / 4; ++j) {
#if SAPT_F32X2
        if (kSplitD1 && j == HW / 8) d1_loaded();      // the second half of D1 (ordered groups of one quad: quad HW/8 is the first to read it)
        // packed fp32 pairs (FADD2 / FMUL2): 31 instead of 38 instructions per four hidden units
        f32x2 s10, s32;
        float e0;
        sigmoid4_packed_e<false>(__uint_as_float(zr[4 * j]), __uint_as_float(zr[4 * j + 1]), __uint_as_float(zr[4 * j + 2]),
                                 __uint_as_float(zr[4 * j + 3]), lim, 1.0f, s10, s32, e0);
        float h0, h1, h2, h3;
        unpack2(s10, h1, h0); unpack2(s32, h3, h2);
        if constexpr (kChain > 0) { if (j % kChain == kChain - 1) lim = fmaf(kDep ? e0 : h0, 0.0f, 30.0f); }
        const uint32_t u0 = __float_as_uint(h0), u1 = __float_as_uint(h1), u2 = __float_as_uint(h2), u3 = __float_as_uint(h3);
        hv[2 * j] = __byte_perm(u0, u1, 0x7632);
        hv[2 * j + 1] = __byte_perm(u2, u3, 0x7632);
        float r0, r1, r2, r3;
        unpack2(sub2(s10, pack2(__uint_as_float(u1 & 0xffff0000u), __uint_as_float(u0 & 0xffff0000u))), r1, r0);
        unpack2(sub2(s32, pack2(__uint_as_float(u3 & 0xffff0000u), __uint_as_float(u2 & 0xffff0000u))), r3, r2);
        hv[HW / 2 + 2 * j] = __byte_perm(__float_as_uint(r0) + 0x8000u, __float_as_uint(r1) + 0x8000u, 0x7632);
        hv[HW / 2 + 2 * j + 1] = __byte_perm(__float_as_uint(r2) + 0x8000u, __float_as_uint(r3) + 0x8000u, 0x7632);
#else
        float h0, h1, h2, h3;
        sigmoid4_lim(__uint_as_float(zr[4 * j]), __uint_as_float(zr[4 * j + 1]), __uint_as_float(zr[4 * j + 2]),
                     __uint_as_float(zr[4 * j + 3]), lim, h0, h1, h2, h3);
        if constexpr (kChain > 0) { if (j % kChain == kChain - 1) lim = fmaf(h0, 0.0f, 30.0f); }
        // h1 piece = h truncated to bf16 (one PRMT packs two), h2 piece = the exact, non-negative residual
        // rounded half-up to bf16. The residual is not zero-mean, so MMA-2 also accumulates h2.w2.
        const uint32_t u0 = __float_as_uint(h0), u1 = __float_as_uint(h1), u2 = __float_as_uint(h2), u3 = __float_as_uint(h3);
        hv[2 * j] = __byte_perm(u0, u1, 0x7632);
        hv[2 * j + 1] = __byte_perm(u2, u3, 0x7632);
        const uint32_t r0 = __float_as_uint(h0 - __uint_as_float(u0 & 0xffff0000u)) + 0x8000u;
        const uint32_t r1 = __float_as_uint(h1 - __uint_as_float(u1 & 0xffff0000u)) + 0x8000u;
        const uint32_t r2 = __float_as_uint(h2 - __uint_as_float(u2 & 0xffff0000u)) + 0x8000u;
        const uint32_t r3 = __float_as_uint(h3 - __uint_as_float(u3 & 0xffff0000u)) + 0x8000u;
        hv[HW / 2 + 2 * j] = __byte_perm(r0, r1, 0x7632);
        hv[HW / 2 + 2 * j + 1] = __byte_perm(r2, r3, 0x7632);
#endif
      }
      }
    };
    // activations of unit u -> TMEM (A operand of MMA-2): the stores are issued as soon as the buffer is free; the hand-over to the
    // MMA warp (wait::st, fence, arrive) follows after phase 2 — MMA-2(u) has a whole unit of slack, and the store latency is no
    // longer a bubble that all four warps of a scheduler sit through together
    auto publish_issue = [&]() {
      // chain HV region = [h1 pairs of all HP hidden units (HP/2 cols) | h2 pairs (HP/2 cols)]
      const uint32_t hva = tmem_hv + lane_base + (uint32_t)(c * HP);
      if constexpr (HALVES == 1) {
        if constexpr (HW == 32) tmem_st32(hva, hv); else tmem_st16(hva, hv);
      } else {
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { lo[j] = hv[j]; hi[j] = hv[16 + j]; }
        tmem_st16(hva + (uint32_t)(hf * 16), lo);
        tmem_st16(hva + (uint32_t)(HP / 2 + hf * 16), hi);
      }
    };
    auto publish_commit = [&](int u) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hv_full[(u & 1) * UC + c]);
    };
    if (n_units > 0) { phase1(0); publish_issue(); publish_commit(0); }
    for (int it = 0; it < n_my_tiles; ++it) {
      const int64_t row = (t_begin + it) * kTileM + quarter * 32 + lane;
      const int y = __ldg(p.y + row);
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) {
        if (DYN && uu >= upt) break;
        const int u = it * upt + uu;
        const bool has_next = (u + 1 < n_units);
        const bool mine = (HALVES == 1) || ((u & 1) == hf);   // which of the chain's warps runs phase 2 of this unit
        if (has_next) phase1(u + 1);
        if (trw) TC2_TRACE(c, u, 2);
        // ---- D2 of unit u: wait for MMA-2 (this also frees the HV buffer), pull it into registers ----
        // the partial sums of D2 (dr2 unused when PARTS == 2). ON = 32 serves O = 17..20: only 20 columns are read
        constexpr int NR = C::NR;
        uint32_t dr0[NR], dr1[NR], dr2[ON == 32 ? 1 : NR];
        mbar_wait(&d2_full[(u & 1) * UC + c], (u >> 1) & 1);
        if (trw) TC2_TRACE(c, u, 3);
        tc_fence_after();
        if (mine) {
          const uint32_t taddr = tmem_base + lane_base + (uint32_t)(kD2Base + (u % kD2Bufs) * (UC * N2) + c * N2);
          if constexpr (ON == 32) {
            tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&dr0[0])); tmem_ld4(taddr + 16, *reinterpret_cast<uint32_t(*)[4]>(&dr0[16]));
            tmem_ld16(taddr + ON, *reinterpret_cast<uint32_t(*)[16]>(&dr1[0])); tmem_ld4(taddr + ON + 16, *reinterpret_cast<uint32_t(*)[4]>(&dr1[16]));
          }
          else if constexpr (NR == 16) { tmem_ld16(taddr, dr0); tmem_ld16(taddr + ON, dr1); tmem_ld16(taddr + 2 * ON, dr2); }
          else if constexpr (NR == 8) { tmem_ld8(taddr, dr0); tmem_ld8(taddr + ON, dr1); tmem_ld8(taddr + 2 * ON, dr2); }
          else {
            auto ld12 = [&](uint32_t ta, uint32_t* d) { tmem_ld8(ta, *reinterpret_cast<uint32_t(*)[8]>(d)); tmem_ld4(ta + 8, *reinterpret_cast<uint32_t(*)[4]>(d + 8)); };
            ld12(taddr, dr0); ld12(taddr + ON, dr1); ld12(taddr + 2 * ON, dr2);
          }
        }
        if (has_next) publish_issue();      // the HV buffer is free: MMA-2(u) has completed
        if (mine) tmem_ld_wait();
        if (trw) TC2_TRACE(c, u, 4);
        if (!mine) { if (has_next) publish_commit(u + 1); continue; }
        // ---- phase 2: output layer (exactly O outputs: quads, then a pair and / or a single one) ----
        // s is kept as log2(e) * sigmoid. With FOLD the output bias rides in the MMA: hidden slot H is padding whose
        // activation is sigmoid(0) = 1/2, and its W2 row holds twice the bias (tc2_pack_kernel).
        float tmin = 3.0e38f; int pm = 0; float sum = 0.f, oy = 0.f;
        auto phase2 = [&](auto fold_tag) {
          constexpr bool FOLD = decltype(fold_tag)::value;
          const float* bias = sBias + (uu * UC + c) * ON;
          auto tval = [&](int o) {  // o is a compile-time constant after unrolling
            float v;
            if constexpr (PARTS == 3) v = (__uint_as_float(dr0[o]) + __uint_as_float(dr1[o])) + __uint_as_float(dr2[o]);
            else v = __uint_as_float(dr0[o]) + __uint_as_float(dr1[o]);
            if constexpr (!FOLD) v += bias[o];
            return v;
          };
          // two neighbouring outputs at once: the partial sums are added as packed pairs (o is even; the registers of a
          // tcgen05.ld are consecutive, so the pairs need no moves)
          auto tpair = [&](int o, float& ta, float& tb) {
            f32x2 v = add2(pack2(__uint_as_float(dr0[o]), __uint_as_float(dr0[o + 1])), pack2(__uint_as_float(dr1[o]), __uint_as_float(dr1[o + 1])));
            if constexpr (PARTS == 3) v = add2(v, pack2(__uint_as_float(dr2[o]), __uint_as_float(dr2[o + 1])));
            if constexpr (!FOLD) v = add2(v, pack2(bias[o], bias[o + 1]));
            unpack2(v, ta, tb);
          };
          auto emit = [&](int o, float t, float sv) {
            if (t < tmin) { tmin = t; pm = o; }
            sum += ex2_approx(sv);
            if (o == y) oy = sv;
          };
          auto single = [&](int o) {
            const float t = tval(o);
            emit(o, t, kLog2e * rcp_approx(1.0f + ex2_approx(fminf(t, 30.0f))));
          };
#pragma unroll
          for (int g = 0; g < NR / 4; ++g) {
            const int rem = O - 4 * g;   // warp-uniform
            if (rem >= 4) {
#if SAPT_F32X2
              float t0, t1, t2, t3;
              tpair(4 * g, t0, t1); tpair(4 * g + 2, t2, t3);
              f32x2 s10, s32;
              sigmoid4_packed<true>(t0, t1, t2, t3, 30.0f, kLog2e, s10, s32);
              float s0, s1, s2, s3;
              unpack2(s10, s1, s0); unpack2(s32, s3, s2);
#else
              const float t0 = tval(4 * g), t1 = tval(4 * g + 1), t2 = tval(4 * g + 2), t3 = tval(4 * g + 3);
              float s0, s1, s2, s3;
              sigmoid4_scaled(t0, t1, t2, t3, kLog2e, s0, s1, s2, s3);
#endif
              emit(4 * g, t0, s0); emit(4 * g + 1, t1, s1); emit(4 * g + 2, t2, s2); emit(4 * g + 3, t3, s3);
            } else if (rem > 0) {
              if (rem & 2) {
#if SAPT_F32X2
                float t0, t1;
                tpair(4 * g, t0, t1);
#else
                const float t0 = tval(4 * g), t1 = tval(4 * g + 1);
#endif
                const float a0 = 1.0f + ex2_approx(fminf(t0, 30.0f)), a1 = 1.0f + ex2_approx(fminf(t1, 30.0f));
                const float r = rcp_approx(a0 * a1) * kLog2e;
                emit(4 * g, t0, r * a1); emit(4 * g + 1, t1, r * a0);
                if (rem & 1) single(4 * g + 2);
              } else {
                single(4 * g);
              }
            }
          }
        };
        if (p.fold_bias) phase2(std::true_type{}); else phase2(std::false_type{});
        if (has_next) publish_commit(u + 1);
        if (trw) TC2_TRACE(c, u, 5);
        if (y >= 0) {
          accL[uu] += (double)(kLn2 * (oy - lg2_approx(sum)));
          accNc[uu] += (pm == y);
          const int d = pm - y;
          accSq[uu] += d * d;
        }
      }
    }
#pragma unroll
    for (int uu = 0; uu < UPT; ++uu) {
      const int chain = uu * UC + c;
      const double l = warp_sum(accL[uu]);
      const int a = warp_sum(accNc[uu]);
      const int b = warp_sum(accSq[uu]);
      if (lane == 0) {
        double* r = red + ((chain * 4 + quarter) * HALVES + hf) * 3;
        r[0] = l; r[1] = (double)a; r[2] = (double)b;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
  if (threadIdx.x < CH && g0 + (int)threadIdx.x < nact) {
    const double* r = red + threadIdx.x * (4 * HALVES * 3);
    double l = 0.0, a = 0.0, b = 0.0;
#pragma unroll
    for (int q = 0; q < 4 * HALVES; ++q) { l += r[3 * q]; a += r[3 * q + 1]; b += r[3 * q + 2]; }
    const int64_t o = (int64_t)blockIdx.x * p.R + (g0 + threadIdx.x);
    p.partL[o] = l;
    p.partNc[o] = (int32_t)a;
    p.partSq[o] = (int32_t)b;
  }
}

// ---- host ------------------------------------------------------------------------------------------
int tc_make_map_2d(CUtensorMap* map, void* base, uint64_t inner_elems, uint64_t rows, uint32_t box_rows);  // sapt_tc_host.cu

struct Tc2Geom { int CH, KROW, W2BOX, UPT; size_t smem; };

template <int HP, int ON, int DKC>
static Tc2Geom geom_of() {
  using C = TcCfg<HP, ON, DKC>;
  return Tc2Geom{C::CH, C::KBA * 64, C::kW2Box, C::UPT, (size_t)C::smem_bytes(C::UPT > 0 ? C::UPT : 1)};
}

template <int HP, int ON, int DKC, int OC = 0, int HC = 0, bool DYN = false>
static int launch_tc2(const CUtensorMap& ma, const CUtensorMap& mb1, const CUtensorMap& mw2, const Tc2Params& a, dim3 grid, cudaStream_t st) {
  using C = TcCfg<HP, ON, DKC>;
  if constexpr (C::UPT < 1) {
    set_error("tcgen05 path: variant HP=%d ON=%d DKC=%d does not fit in shared memory", HP, ON, DKC);
    return SAPT_ERR_UNSUPPORTED;
  } else {
    const size_t smem = C::smem_bytes(C::UPT);
    auto k = loglik_tc2_kernel<HP, ON, DKC, OC, HC, (DYN && (C::UPT > 1))>;
    SAPT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, v2::kThreads, smem, st>>>(ma, mb1, mw2, a);
    SAPT_CHECK_LAUNCH();
    return SAPT_OK;
  }
}

// dispatch over the compiled variants; fn is a generic lambda called with three integral_constants
template <typename F>
static int tc2_dispatch(int HP, int ON, int DKC, F&& fn) {
#define SAPT_V(H, O, K) if (HP == H && ON == O && DKC == K) return fn(std::integral_constant<int, H>{}, std::integral_constant<int, O>{}, std::integral_constant<int, K>{});
#ifdef SAPT_TC2_FEW   // development builds: only the BASELINE shapes (C3, C4, C5)
  SAPT_V(32, 16, 1) SAPT_V(64, 8, 12) SAPT_V(64, 16, 12) SAPT_V(32, 32, 9)
#else
  SAPT_V(16, 16, 1) SAPT_V(16, 16, 2) SAPT_V(16, 16, 3) SAPT_V(16, 16, 4)
  SAPT_V(32, 16, 1) SAPT_V(32, 16, 2) SAPT_V(32, 16, 3) SAPT_V(32, 16, 4)
  SAPT_V(32, 32, 1) SAPT_V(32, 32, 2) SAPT_V(32, 32, 3) SAPT_V(32, 32, 4)
  SAPT_V(64, 16, 1) SAPT_V(64, 16, 2) SAPT_V(64, 16, 3) SAPT_V(64, 16, 4)
  SAPT_V(64, 32, 1) SAPT_V(64, 32, 2) SAPT_V(64, 32, 3) SAPT_V(64, 32, 4)
  // bias in the spare x1 column (D not a multiple of 16)
  SAPT_V(16, 16, 9) SAPT_V(16, 16, 10) SAPT_V(16, 16, 11) SAPT_V(16, 16, 12)
  SAPT_V(32, 16, 9) SAPT_V(32, 16, 10) SAPT_V(32, 16, 11) SAPT_V(32, 16, 12)
  SAPT_V(32, 32, 9) SAPT_V(32, 32, 10) SAPT_V(32, 32, 11) SAPT_V(32, 32, 12)
  SAPT_V(64, 16, 9) SAPT_V(64, 16, 10) SAPT_V(64, 16, 11) SAPT_V(64, 16, 12)
  SAPT_V(64, 32, 9) SAPT_V(64, 32, 10) SAPT_V(64, 32, 11) SAPT_V(64, 32, 12)
  // O <= 8 at HP = 64 (see TcCfg)
  SAPT_V(64, 8, 1) SAPT_V(64, 8, 2) SAPT_V(64, 8, 3) SAPT_V(64, 8, 4) SAPT_V(64, 8, 9) SAPT_V(64, 8, 10) SAPT_V(64, 8, 11) SAPT_V(64, 8, 12)
  SAPT_V(64, 12, 1) SAPT_V(64, 12, 2) SAPT_V(64, 12, 3) SAPT_V(64, 12, 4) SAPT_V(64, 12, 9) SAPT_V(64, 12, 10) SAPT_V(64, 12, 11) SAPT_V(64, 12, 12)
#endif
#undef SAPT_V
  set_error("tcgen05 path: no variant for HP=%d ON=%d DKC=%d", HP, ON, DKC);
  return SAPT_ERR_UNSUPPORTED;
}

static int tc2_dkc(int D) { return D <= 16 ? 1 : (D + 15) / 16; }   // 16-feature chunks per x piece

static void tc2_shape(const sapt_problem* p, int& HP, int& ON, int& DKC) {
  HP = p->H <= 16 ? 16 : (p->H <= 32 ? 32 : 64);
  ON = p->O <= 16 ? 16 : 32;
  DKC = tc2_dkc(p->D);
  if (p->D < 16 * DKC) DKC += 8;   // see TcCfg: bias in the spare x1 column
}

bool tc2_supported(int D, int H, int O) {
  if (D > 64 || H > 64 || O > 20) return false;
  sapt_problem q{}; q.D = D; q.H = H; q.O = O;
  int HP, ON, DKC; tc2_shape(&q, HP, ON, DKC);
  if (ON > HP) return false;
  int upt = 0;
  tc2_dispatch(HP, ON, DKC, [&](auto h, auto o, auto k) { upt = TcCfg<h(), o(), k()>::UPT; return 0; });
  return upt >= 1;
}

// What ONE replica x row evaluation executes on this path (bench.py's governing roofline is computed from this, not from a
// model of what an ideal kernel would do): MUFU operations and tensor-pipe FLOPs including padding and the split products.
//   phase 1: HP sigmoids = HP ex2 + HP/4 shared reciprocals (sigmoid4);  phase 2: O ex2 + (quads, pair, single) reciprocals for
//   the output sigmoids, O ex2 for the softmax, one lg2.
//   The variants with a compile-time hidden count (tc2_const_hidden) skip the padded slots: H ex2 + ceil(H / 4) reciprocals.
static int tc2_const_hidden(int HP, int ONX, int DK, int H, int O) {   // HC of the kernel launch_loglik_tc2 picks for this topology, 0 = generic
#if SAPT_CONST_O
  if (HP == 32 && ONX == 16 && DK == 1 && O == 10 && H == 30) return SAPT_C3_HC;
  if (HP == 32 && ONX == 32 && DK == 9 && O == 18 && H == 25) return SAPT_C5_HC;   // 0: at 6-25-18 the skipped slots measured 6 % SLOWER (3.86 vs 3.62 ms)
#endif
  return 0;
}
void tc2_work_model(int D, int H, int O, double* mufu, double* mma_flops) {
  sapt_problem q{}; q.D = D; q.H = H; q.O = O;
  int HP, ON, DK; tc2_shape(&q, HP, ON, DK);
  const int DKC = DK & 7;
  const int rcp2 = O / 4 + ((O & 2) ? 1 : 0) + ((O & 1) ? 1 : 0);
  const int hc = tc2_const_hidden(HP, ON, DK, H, O);
  const int p1 = hc ? hc + (hc + 3) / 4 : HP + HP / 4;
  *mufu = (double)p1 + O + rcp2 + O + 1;
  const int k1 = 16 * ((SAPT_X_FP16 ? 3 : 6) * DKC + (DK >= 8 ? 0 : 1));   // MMA-1: piece products (+ bias chunk) of 16 DKC features
  const int n2 = (ON == 16) ? (48 + 32) : (64 + 32 + 64);       // MMA-2: output columns summed over the passes h1.[w1|w2(|w3)], h2.[w1|w2]
  *mma_flops = 2.0 * k1 * HP + 2.0 * HP * n2;                   // every pass has K = HP
}

int tc2_a_row_elems(int D) { return ((48 * tc2_dkc(D) + 16 + 63) / 64) * 64; }          // three-piece bf16 image (batch-Langevin kernel)
int tc2_a16_row_elems(int D) {                                                            // two-piece fp16 image (this kernel)
  const int dkc = tc2_dkc(D);
  return ((32 * dkc + (D < 16 * dkc ? 0 : 16) + 63) / 64) * 64;
}
int tc2_piece_width(int D) { return 16 * tc2_dkc(D); }

int launch_loglik_tc2(sapt_problem* p, int which, const float* theta, int64_t ld, int R, const int32_t* rows, const int32_t* n_rows,
                      double* L, int32_t* nc, double* sq, cudaStream_t st) {
  const DataSet& ds = p->set[which];
  if (R <= 0 || ds.n <= 0) return SAPT_OK;
  if (SAPT_X_FP16 && !ds.f16_range_ok) {
    set_error("tcgen05 path: this data set has features beyond the fp16 operand range (|x| > 6e4); use SAPT_PATH_F32 / SAPT_PATH_AUTO");
    return SAPT_ERR_UNSUPPORTED;
  }
  int HP, ON, DKC; tc2_shape(p, HP, ON, DKC);
  if (ON > HP) { set_error("tcgen05 path: O=%d needs H > 16", p->O); return SAPT_ERR_UNSUPPORTED; }
  Tc2Geom g{};
  const int ONX = (HP == 64 && ON == 16 && p->O <= 12) ? (p->O <= 8 ? 8 : 12) : ON;   // narrow accumulator read-out, see TcCfg
  int rc = tc2_dispatch(HP, ONX, DKC, [&](auto h, auto o, auto k) { g = geom_of<h(), o(), k()>(); return 0; });
  if (rc) return rc;
  if (g.UPT < 1) { set_error("tcgen05 path: topology %d-%d-%d does not fit", p->D, p->H, p->O); return SAPT_ERR_UNSUPPORTED; }
  const int CH = g.CH;
  const int n_groups = (R + CH - 1) / CH;
  const int64_t n_tiles = ds.n_pad / v2::kTileM;
  // Row chunks: one CTA per (chunk, chain group), one CTA per SM at a time. The chunk count that minimises the makespan
  //   waves(n) x (tiles per chunk + ~1 tile of per-CTA start-up), waves(n) = ceil(groups x n / SMs)
  // is taken: with a fixed "about 8 CTAs per SM" rule C3 (128 groups) ran 1280 CTAs = 8.65 waves and C4 (64 groups) 1216 = 8.2 waves,
  // i.e. the last wave left a third / four fifths of the GPU idle (4 % / 9 % of the kernel).
  int64_t n_chunks = 1;
  {
    const int64_t max_chunks = std::max<int64_t>(1, std::min<int64_t>(n_tiles / 16, 2 * (int64_t)p->num_sms));
    // at least ~8 CTAs per SM when the rows allow it: CTAs do not all take the same time, and many small ones even that out
    // (C5, 512 groups x 132 tiles: 3 chunks measured 5 % faster than the 2 a pure wave count prefers)
    // (with so few chain groups that 8 waves are out of reach anyway — the X-streaming regime — every count is a candidate)
    const int64_t want = ((int64_t)8 * p->num_sms + n_groups - 1) / n_groups;
    const int64_t min_chunks = (want <= max_chunks) ? want : 1;
    double best = 1e300;
    for (int64_t n = min_chunks; n <= max_chunks; ++n) {
      const int64_t waves = ((int64_t)n_groups * n + p->num_sms - 1) / p->num_sms;
      const double cost = (double)waves * ((double)((n_tiles + n - 1) / n) + 1.0);
      if (cost < best * 0.999) { best = cost; n_chunks = n; }
    }
  }
  const int tiles_per_chunk = (int)((n_tiles + n_chunks - 1) / n_chunks);
  n_chunks = (n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;

  auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  const size_t b1_bytes = (size_t)n_groups * CH * HP * g.KROW * 2;
  const size_t w2_bytes = (size_t)n_groups * CH * 3 * ON * 64 * 2;
  const size_t bias_bytes = (size_t)n_groups * CH * ON * sizeof(float);
  const size_t part_bytes = (size_t)n_chunks * R * (sizeof(double) + 2 * sizeof(int32_t));
  rc = ensure_workspace(p, up(b1_bytes) + up(w2_bytes) + up(bias_bytes) + up(part_bytes));
  if (rc) return rc;
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->ws);
  __nv_bfloat16* b1img = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* w2img = reinterpret_cast<__nv_bfloat16*>(ws + up(b1_bytes));
  float* bias2 = reinterpret_cast<float*>(ws + up(b1_bytes) + up(w2_bytes));
  double* partL = reinterpret_cast<double*>(ws + up(b1_bytes) + up(w2_bytes) + up(bias_bytes));
  int32_t* partNc = reinterpret_cast<int32_t*>(partL + (size_t)n_chunks * R);
  int32_t* partSq = partNc + (size_t)n_chunks * R;

  tc2_pack_kernel<<<n_groups * CH, 256, 0, st>>>(theta, ld, rows, n_rows, R, p->D, p->H, p->O, HP, ON, CH, DKC & 7, g.KROW, DKC >= 8 ? 1 : 0, SAPT_X_FP16,
                                                 b1img, w2img, bias2);
  SAPT_CHECK_LAUNCH();
  CUtensorMap mb1, mw2;
  rc = tc_make_map_2d(&mb1, b1img, (uint64_t)g.KROW, (uint64_t)n_groups * CH * HP, (uint32_t)(CH * HP));
  if (rc) return rc;
  rc = tc_make_map_2d(&mw2, w2img, 64, (uint64_t)n_groups * CH * 3 * ON, (uint32_t)(g.W2BOX * 3 * ON));
  if (rc) return rc;
  Tc2Params a{};
  a.y = ds.y; a.n_pad = ds.n_pad; a.bias2 = bias2; a.n_rows = n_rows; a.R = R; a.O = p->O; a.fold_bias = (p->H < HP) ? 1 : 0; a.zero = 0u;
  a.tiles_per_chunk = tiles_per_chunk; a.n_tiles = n_tiles; a.partL = partL; a.partNc = partNc; a.partSq = partSq;
  dim3 grid((unsigned)n_chunks, (unsigned)n_groups);
#ifdef SAPT_TC2_TRACE
  static long long* d_trace = nullptr;
  if (!d_trace) cudaMalloc(&d_trace, 5 * 64 * 8 * sizeof(long long));
  cudaMemsetAsync(d_trace, 0, 5 * 64 * 8 * sizeof(long long), st);
  a.trace = d_trace;
#endif
  const CUtensorMap& ma = *reinterpret_cast<const CUtensorMap*>(SAPT_X_FP16 ? ds.tmap_f16 : ds.tmap);
  // empty chain slots in the last (or only) chain group: when at least ~8 % of the launch's units would be padding, the
  // instantiations that walk only the occupied units (see DYN: they cost a full group 4-6 %). SAPT_TC2_UNITS=static / dynamic
  // overrides the choice (tests run both).
  const int UC = CH / g.UPT;
  const int units_all = n_groups * g.UPT, units_used = (R + UC - 1) / UC;
  bool use_dyn = (g.UPT > 1) && ((units_all - units_used) * 12 >= units_all);
  if (const char* e = getenv("SAPT_TC2_UNITS")) {
    if (e[0] == 's') use_dyn = false;
    else if (e[0] == 'd') use_dyn = (g.UPT > 1);
  }
  if (use_dyn) rc = tc2_dispatch(HP, ONX, DKC, [&](auto h, auto o, auto k) { return launch_tc2<h(), o(), k(), 0, 0, true>(ma, mb1, mw2, a, grid, st); });
  else
#if SAPT_CONST_O
  // output count (and hidden count, see HC) known at compile time for the shapes of the BASELINE configs: C3 16-30-10, C5 6-25-18,
  // C4 54-64-7 (5.69 -> 5.39 ms; while that kernel was still bound by MMA issue the specialisation had measured 3 % slower)
  if (HP == 32 && ONX == 16 && DKC == 1 && p->O == 10 && p->H == 30) rc = launch_tc2<32, 16, 1, 10, SAPT_C3_HC>(ma, mb1, mw2, a, grid, st);
  else if (HP == 32 && ONX == 32 && DKC == 9 && p->O == 18 && p->H == 25) rc = launch_tc2<32, 32, 9, 18, SAPT_C5_HC>(ma, mb1, mw2, a, grid, st);
  else if (HP == 64 && ONX == 8 && DKC == 12 && p->O == 7) rc = launch_tc2<64, 8, 12, 7, 0>(ma, mb1, mw2, a, grid, st);
  else
#endif
  rc = tc2_dispatch(HP, ONX, DKC, [&](auto h, auto o, auto k) { return launch_tc2<h(), o(), k()>(ma, mb1, mw2, a, grid, st); });
  if (rc) return rc;
#ifdef SAPT_TC2_TRACE
  {
    std::vector<long long> h(5 * 64 * 8);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    const long long t0 = h[(4 * 64 + 0) * 8 + 1];
    const char* roles[5] = {"epi0", "epi1", "epi2", "epi3", "mma "};
    for (int u = 16; u < 28; ++u)
      for (int r = 0; r < 5; ++r) {
        printf("trace %s unit %2d:", roles[r], u);
        for (int e = 0; e < 6; ++e) { long long v = h[(r * 64 + u) * 8 + e]; printf(" %8lld", v ? v - t0 : -1); }
        printf("\n");
      }
  }
#endif
  return launch_loglik_finalize(partL, partNc, partSq, (int)n_chunks, R, rows, n_rows, L, nc, sq, st);
}

}  // namespace sapt
