// Host side of the tcgen05 / TMEM / TMA likelihood path (sm_100a only): TMA tensor maps and the
// per-data-set A operand image. The kernel itself is in sapt_loglik_tc2.cu.
//
// A image: bf16 [n_pad][64] = [ x1 | x2 | x3 | 1 1 1 0.. ] (four 16-element chunks = one 128-byte
// swizzle row), x = x1 + x2 + x3 exactly (three bf16 pieces of the fp32 feature). The cross terms
// of the split product are formed by pointing successive MMA K-steps at different (A chunk, B chunk)
// pairs, so no piece is stored twice.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "sapt_device.cuh"
#include "sapt_internal.h"
#include "sapt_tc_ptx.cuh"

namespace sapt {

constexpr int kTileM = 128;
constexpr int kKA = 64;  // bf16 per A / B1 row

bool tc2_supported(int D, int H, int O);   // sapt_loglik_tc2.cu
int tc2_a_row_elems(int D);
int tc2_piece_width(int D);
int tc2_a16_row_elems(int D);
bool tc_topology_supported(int D, int H, int O) { return tc2_supported(D, H, O); }

// ---- per-step operand packing -----------------------------------------------------------------
// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// bf16 [rows][inner_elems] row-major, box = 64 elements (one 128-byte swizzle row) x box_rows
int tc_make_map_2d(CUtensorMap* map, void* base, uint64_t inner_elems, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return SAPT_ERR_CUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)inner_elems, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)inner_elems * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return SAPT_ERR_CUDA; }
  return SAPT_OK;
}

static inline uint16_t f2bf(float v) {  // round-to-nearest-even, host side
  uint32_t u; memcpy(&u, &v, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float v; memcpy(&v, &u, 4); return v; }

int tc_prepare_dataset(sapt_problem* p, int which, const float* x_host, int64_t n) {
  DataSet& ds = p->set[which];
  const int krow = tc2_a_row_elems(p->D);      // bf16 per row: [x1 | x2 | x3 | 1 1 1 0..] padded to 128-byte blocks
  const int dw = tc2_piece_width(p->D);        // features per piece (padded to 16/32/64)
  std::vector<uint16_t> img((size_t)ds.n_pad * krow, 0);
  const uint16_t one = f2bf(1.0f);
  for (int64_t i = 0; i < ds.n_pad; ++i) {
    uint16_t* row = &img[(size_t)i * krow];
    if (i < n) {
      for (int d = 0; d < p->D; ++d) {
        const float v = x_host[(size_t)i * p->DP + d];
        uint16_t pc[3];
        pc[0] = f2bf(v); const float r1 = v - bf2f(pc[0]);
        pc[1] = f2bf(r1); const float r2 = r1 - bf2f(pc[1]);
        pc[2] = f2bf(r2);
        row[d] = pc[0]; row[dw + d] = pc[1]; row[2 * dw + d] = pc[2];
      }
    }
    row[3 * dw] = one; row[3 * dw + 1] = one; row[3 * dw + 2] = one;
    // when D is not a multiple of 16 the first spare column of the x1 piece carries a 1.0 as well: the likelihood kernel
    // then takes its bias from there (TcCfg BIASX, sapt_loglik_tc2.cu), and for D > 32, where the bias chunk lies outside
    // the 128 image columns of the batch-Langevin contraction X^T.dh, so does that kernel (sapt_langevin_tc.cu)
    if (p->D < dw) row[p->D] = one;
  }
  SAPT_CUDA(cudaMalloc(&ds.a_bf16, img.size() * 2));
  SAPT_CUDA(cudaMemcpy(ds.a_bf16, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  ds.tmap = aligned_alloc(64, sizeof(CUtensorMap));
  int rc = tc_make_map_2d(reinterpret_cast<CUtensorMap*>(ds.tmap), ds.a_bf16, (uint64_t)krow, (uint64_t)ds.n_pad, kTileM);
  if (rc) return rc;
  // the likelihood kernel's image: two fp16 pieces x = x_hi + x_lo (22 significant bits; fp16 subnormals keep the absolute error
  // of a feature below 2^-25), [x_hi | x_lo | 1 0..]: half the MMA-1 products of the three-piece bf16 image
  const int krow16 = tc2_a16_row_elems(p->D);
  float amax = 0.f;
  for (int64_t i = 0; i < n; ++i)
    for (int d = 0; d < p->D; ++d) amax = std::max(amax, fabsf(x_host[(size_t)i * p->DP + d]));
  ds.f16_range_ok = (amax <= 6.0e4f);     // un-normalised features beyond fp16's range: the fp32 tiles serve this data set
  if (!ds.f16_range_ok) return SAPT_OK;
  std::vector<uint16_t> img16((size_t)ds.n_pad * krow16, 0);
  const __half h_one = __float2half_rn(1.0f);
  uint16_t one16; memcpy(&one16, &h_one, 2);
  for (int64_t i = 0; i < ds.n_pad; ++i) {
    uint16_t* row = &img16[(size_t)i * krow16];
    if (i < n) {
      for (int d = 0; d < p->D; ++d) {
        float v = x_host[(size_t)i * p->DP + d];
        v = std::min(65504.0f, std::max(-65504.0f, v));
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        memcpy(&row[d], &hi, 2); memcpy(&row[dw + d], &lo, 2);
      }
    }
    if (p->D < dw) row[p->D] = one16;          // bias rides in the spare x_hi column (TcCfg BIASX)
    else { row[2 * dw] = one16; row[2 * dw + 1] = one16; }   // else a bias chunk [1 1 0 ..] after the two pieces: it meets [b_hi b_lo 0 ..]
  }
  SAPT_CUDA(cudaMalloc(&ds.a_f16, img16.size() * 2));
  SAPT_CUDA(cudaMemcpy(ds.a_f16, img16.data(), img16.size() * 2, cudaMemcpyHostToDevice));
  ds.tmap_f16 = aligned_alloc(64, sizeof(CUtensorMap));
  return tc_make_map_2d(reinterpret_cast<CUtensorMap*>(ds.tmap_f16), ds.a_f16, (uint64_t)krow16, (uint64_t)ds.n_pad, kTileM);
}

int launch_loglik_tc2(sapt_problem* p, int which, const float* theta, int64_t ld, int R, const int32_t* rows, const int32_t* n_rows,
                      double* L, int32_t* nc, double* sq, cudaStream_t st);  // sapt_loglik_tc2.cu

int launch_loglik_tc(sapt_problem* p, int which, const float* theta, int64_t ld, int R, const int32_t* rows, const int32_t* n_rows,
                     double* L, int32_t* nc, double* sq, cudaStream_t st) {
  return launch_loglik_tc2(p, which, theta, ld, R, rows, n_rows, L, nc, sq, st);
}

}  // namespace sapt
