// Per-chain MCMC kernels: proposal (+prior, +surrogate coin), Metropolis-Hastings accept with
// the reference's bookkeeping, the LG mid-step, and the sequential replica-exchange sweep.
// Reference lines are cited per kernel (RV=_revamp.py, RW/LG = the two misc_codeapril2020 forks).
#include <math.h>

#include <map>
#include <mutex>
#include <utility>

#include "sapt_device.cuh"
#include "sapt_internal.h"
#include "sapt_mh.cuh"

namespace sapt {

// -------------------------------------------------------------------------------------------
// prior: -(P/2) ln s2 - sum(w^2)/(2 s2)                                        RV:404-410
// -------------------------------------------------------------------------------------------
__global__ void prior_kernel(const float* theta, int64_t ld, int P, double s2, double* prior) {
  __shared__ double scratch[32];
  const float* w = theta + (int64_t)blockIdx.x * ld;
  double acc = 0.0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) { const double v = (double)w[i]; acc += v * v; }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) prior[blockIdx.x] = prior_from_sumsq(acc, P, s2);
}

// -------------------------------------------------------------------------------------------
// proposal                                                       RV:521,525 / LG:551-560,651
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) propose_kernel(sapt_propose_args a) {
  __shared__ double scratch[32];
  __shared__ int s_sur;
  const int nrows = a.n_rows ? *a.n_rows : a.R;
  if ((int)blockIdx.x >= nrows) return;
  const int row = a.rows ? a.rows[blockIdx.x] : (int)blockIdx.x;
  const int slot = a.slot_of_row ? a.slot_of_row[row] : (row + a.slot_offset);
  const int step_i = a.step + (a.step_dev ? *a.step_dev : 0);
  if (threadIdx.x == 0) {
    int sur = 0;
    if (a.decide) {
      const double ku = a.ku_inject ? a.ku_inject[slot] : (double)philox_uniform(a.seed, SAPT_STREAM_KU, step_i, (uint32_t)slot, 0);
      sur = (a.surrogate_ready && ku < a.surrogate_prob && step_i >= a.gate + 1) ? 1 : 0;
      a.is_sur[row] = sur;
      if (sur) { const int k = atomicAdd(a.n_sur, 1); a.sur_rows[k] = row; }
      else     { const int k = atomicAdd(a.n_true, 1); a.true_rows[k] = row; }
    }
    s_sur = sur;
  }
  __syncthreads();
  const bool second = a.decide && a.redraw && s_sur;
  const float* zin = second ? a.z2_inject : a.z_inject;
  const int kind = second ? SAPT_STREAM_Z2 : a.kind;
  const float* __restrict__ b = a.base + (int64_t)row * a.ld;
  float* __restrict__ o = a.proposal + (int64_t)row * a.ld;
  const float step = (float)a.step_w;
  double acc = 0.0;
  if (zin) {
    const float* z = zin + (int64_t)slot * a.P;
    for (int i = threadIdx.x; i < a.P; i += blockDim.x) {
      const float v = fmaf(step, z[i], b[i]);
      o[i] = v;
      acc += (double)v * (double)v;
    }
  } else {
    const int nblk = (a.P + 3) >> 2;
    for (int q = threadIdx.x; q < nblk; q += blockDim.x) {
      float z[4];
      philox_normal4(a.seed, kind, step_i, (uint32_t)slot, (uint32_t)q, z);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = 4 * q + j;
        if (i < a.P) {
          const float v = fmaf(step, z[j], b[i]);
          o[i] = v;
          acc += (double)v * (double)v;
        }
      }
    }
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) a.prior_prop[row] = prior_from_sumsq(acc, a.P, a.sigma_squared);
}

// -------------------------------------------------------------------------------------------
// LG: bookkeeping after the FIRST true evaluation + Langevin coin               LG:622-649
// -------------------------------------------------------------------------------------------
__global__ void lg_first_kernel(sapt_lg_first_args a) {
  const int n = a.n_true ? *a.n_true : a.R;
  if ((int)blockIdx.x >= n) return;
  const int row = a.true_rows ? a.true_rows[blockIdx.x] : (int)blockIdx.x;
  const int slot = a.st.slot_of_row[row];
  const int step_i = a.step + (a.step_dev ? *a.step_dev : 0);
  __shared__ int s_cnt;
  if (threadIdx.x == 0) {
    const double L = a.L_train[row];
    a.surg_first[row] = L / a.st.temperature[slot];
    int cnt = -1;
    if (a.train_count) {
      cnt = a.train_count[slot];
      if (cnt < a.train_cap) { a.train_L[(int64_t)slot * a.train_cap + cnt] = L; a.train_count[slot] = cnt + 1; }
      else cnt = -1;
    }
    s_cnt = cnt;
    if (a.use_langevin) {
      const double lx = a.lx2_inject ? a.lx2_inject[slot] : (double)philox_uniform(a.seed, SAPT_STREAM_LX2, step_i, (uint32_t)slot, 0);
      if (lx < a.l_prob) { const int k = atomicAdd(a.n_lang, 1); a.lang_rows[k] = row; }
    }
  }
  __syncthreads();
  if (s_cnt >= 0 && a.train_theta) {
    const float* src = a.proposal + (int64_t)row * a.ld;
    float* dst = a.train_theta + ((int64_t)slot * a.train_cap + s_cnt) * a.P;
    for (int i = threadIdx.x; i < a.P; i += blockDim.x) dst[i] = src[i];
  }
}

// -------------------------------------------------------------------------------------------
// Metropolis-Hastings accept + bookkeeping                                      RV:525-642
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mh_accept_kernel(sapt_accept_args a) {
  const int row = blockIdx.x;
  const int slot = a.st.slot_of_row[row];
  const int S = a.hist.n_slots;
  const int step_i = a.step + (a.step_dev ? *a.step_dev : 0);
  __shared__ int s_accept, s_cnt;
  const int64_t hi = (int64_t)step_i * S + slot;        // history row i
  const int64_t ho = (int64_t)(step_i + 1) * S + slot;  // history row i+1
  if (threadIdx.x == 0) {
    MhInputs in;
    in.sur = a.is_sur && a.is_sur[row];
    in.sur_s = a.sur_s ? a.sur_s[row] : 0.f;
    in.L_train = a.L_train[row]; in.nc_train = a.nc_train[row]; in.sq_train = a.sq_train[row];
    in.L_test = a.n_test > 0 ? a.L_test[row] : 0.0; in.nc_test = a.n_test > 0 ? a.nc_test[row] : 0; in.sq_test = a.n_test > 0 ? a.sq_test[row] : 0.0;
    in.prior_prop = a.prior_prop[row];
    in.has_surg_first = a.surg_first != nullptr; in.surg_first = a.surg_first ? a.surg_first[row] : 0.0;
    in.ls_pred = a.ls_pred ? a.ls_pred[row] : 0.0; in.ls_avg = a.ls_avg ? a.ls_avg[row] : 0.0;
    int cnt;
    s_accept = mh_decide(a, slot, step_i, in, &cnt) ? 1 : 0;
    s_cnt = cnt;
  }
  __syncthreads();
  const bool acc = s_accept != 0;
  const float* __restrict__ prop = a.proposal + (int64_t)row * a.ld;
  float* __restrict__ th = a.theta + (int64_t)row * a.ld;
  float* tr = (s_cnt >= 0 && a.train_theta) ? a.train_theta + ((int64_t)slot * a.train_cap + s_cnt) * a.P : nullptr;
  float* pw_out = ((a.flags & SAPT_F_RECORD_POS) && a.hist.pos_w) ? a.hist.pos_w + ho * a.P : nullptr;
  const float* pw_in = pw_out ? a.hist.pos_w + hi * a.P : nullptr;
  for (int i = threadIdx.x; i < a.P; i += blockDim.x) {
    const float v = prop[i];
    if (tr) tr[i] = v;
    if (acc) th[i] = v;
    if (pw_out) pw_out[i] = acc ? v : pw_in[i];  // RV:590 / RV:608: reject copies HISTORY row i
  }
}

// -------------------------------------------------------------------------------------------
// replica exchange: sequential sweep over all chains                 RW:869-898, RW:1095-1099
// -------------------------------------------------------------------------------------------
__global__ void advance_kernel(int32_t* counter, int delta) { *counter += delta; }

// The sweep is a recurrence on ONE carried message: at pair k the carried message (what position k holds after
// the earlier pairs) competes with the message chain k+1 posted; the loser of the comparison stays at k, the
// other is carried on. Chains are staged through shared memory in chunks so any R fits; thread 0 walks a chunk,
// the CTA loads / scatters around it.
// The walk itself must stay sequential, so everything that does not depend on the carried message is done by the
// whole CTA beforehand: u < min(1, 0.5 exp(min(709, p - c)))  <=>  u < 1 and c < p - ln(2u), i.e. the exponential
// of the reference's rule becomes one logarithm per pair computed in parallel (t = p - ln(2u)) and the serial
// step is a compare + select (~25 cycles instead of ~280; the sweep over 65 536 chains went from 9.3 ms to <1 ms).
// The two forms can differ only when u is within an ulp of the acceptance probability (a stated near-threshold
// tie); u <= 0 keeps the literal formula, NaNs behave like Python's min(1, nan) == 1. The compare runs on
// order-preserving int64 keys of the doubles (integer compare + selects; fp64 compares sit on the slow pipe).
constexpr int kSweepChunk = 2048;

// monotone map double -> int64 (non-NaN); the callers give NaN the key they need
__device__ __forceinline__ long long order_key(double x) {
  const long long b = __double_as_longlong(x);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}

__global__ void swap_sweep_kernel(int R, double* lik_post, const double* temperature, int32_t* row_of_slot, uint32_t flags,
                                  uint32_t seed, int sweep_base, const int32_t* sweep_dev, const double* u_inject,
                                  uint8_t* swapped, int32_t* num_swap, int32_t* src_slot, int row0, int R_local,
                                  int32_t* slot_of_row_local, int32_t* scratch_src /*[R]*/, double* scratch_post /*[R]*/) {
  __shared__ long long s_kp[kSweepChunk];  // key of the likelihood chain k+1 posted; NaN -> smallest key (always swaps onward)
  __shared__ long long s_kt[kSweepChunk];  // key of t = p - ln(2u); NaN -> largest key (swap); literal pairs: bits of u
  __shared__ int32_t s_src[kSweepChunk];   // output: which chain's message stays at position k
  __shared__ uint8_t s_f[kSweepChunk];     // bit 0 = (u < 1), bit 1 = literal formula
  __shared__ uint8_t s_sw[kSweepChunk];    // output: swapped
  __shared__ long long c_key;              // carried message: key and chain
  __shared__ int c_src, c_ns;
  const int sweep = sweep_base + (sweep_dev ? *sweep_dev : 0);
  auto posted = [&](int k) {  // RW:623 posts likelihood*adapttemp ; LG:805 posts the tempered value
    return (flags & SAPT_F_POST_TEMPERED) ? lik_post[k] : lik_post[k] * temperature[k];
  };
  auto pkey = [&](double v) { return (v != v) ? (long long)0x8000000000000000ULL : order_key(v); };
  if (threadIdx.x == 0) { c_key = pkey(posted(0)); c_src = 0; c_ns = 0; }
  for (int k0 = 0; k0 < R - 1; k0 += kSweepChunk) {
    const int n = min(kSweepChunk, R - 1 - k0);
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const int k = k0 + j;
      const double pk = posted(k + 1);
      const double u = u_inject ? u_inject[k] : (double)philox_uniform(seed, SAPT_STREAM_SWAP, sweep, 0u, (uint32_t)k);
      const bool literal = !(u > 0.0);
      const double t = pk - log(2.0 * u);
      s_kp[j] = pkey(pk);
      s_kt[j] = literal ? __double_as_longlong(u) : ((t != t) ? 0x7fffffffffffffffLL : order_key(t));
      s_f[j] = (uint8_t)((u < 1.0 ? 1 : 0) | (literal ? 2 : 0));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long ck = c_key; int cs = c_src; int ns = c_ns;
      constexpr int kB = 8;                    // operands of kB pairs are fetched before the dependent chain starts
      for (int j0 = 0; j0 < n; j0 += kB) {
        long long kp[kB], kt[kB]; int f[kB]; int src[kB]; uint8_t swv[kB];
#pragma unroll
        for (int e = 0; e < kB; ++e) {
          const int j = min(j0 + e, n - 1);
          kp[e] = s_kp[j]; kt[e] = s_kt[j]; f[e] = s_f[j];
        }
#pragma unroll
        for (int e = 0; e < kB; ++e) {
          const int k1 = k0 + j0 + e + 1;
          bool sw;
          if (f[e] & 2) {   // RW:886-888 literally (u <= 0)
            const double cp = posted(cs), np_ = posted(min(k1, R - 1));
            const double d = fmin(709.0, np_ - cp);
            double pr = 0.5 * exp(d);
            if (d != d) pr = 1.0;  // min(1, nan) == 1 in Python
            pr = fmin(1.0, pr);
            sw = __longlong_as_double(kt[e]) < pr;
          } else {
            sw = (f[e] & 1) && (ck < kt[e]);   // c < t; a NaN on either side swaps, like min(1, nan) == 1
          }
          if (j0 + e >= n) sw = true;          // padding of the last batch: leaves the carried message alone
          src[e] = sw ? k1 : cs;               // swap: chain k+1's message drops to position k; else the carried one stays
          ns += (sw && j0 + e < n) ? 1 : 0;
          ck = sw ? ck : kp[e];                // no swap: k+1's message is carried on
          cs = sw ? cs : k1;
          swv[e] = sw ? 1 : 0;
        }
#pragma unroll
        for (int e = 0; e < kB; ++e)
          if (j0 + e < n) { s_src[j0 + e] = src[e]; s_sw[j0 + e] = swv[e]; }
      }
      c_key = ck; c_src = cs; c_ns = ns;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const int k = k0 + j;
      scratch_src[k] = s_src[j];
      scratch_post[k] = posted(s_src[j]);      // lik_post is not modified before the re-binding loop below
      if (swapped) swapped[k] = s_sw[j];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    scratch_src[R - 1] = c_src; scratch_post[R - 1] = posted(c_src);
    if (num_swap) *num_swap += c_ns;
  }
  __syncthreads();
  // new binding: chain k now holds the vector chain src[k] posted. row_of_slot is permuted through a copy.
  int32_t* old_rows = scratch_src + R;  // [R]
  for (int k = threadIdx.x; k < R; k += blockDim.x) old_rows[k] = row_of_slot[k];
  __syncthreads();
  for (int k = threadIdx.x; k < R; k += blockDim.x) {
    const int newrow = old_rows[scratch_src[k]];
    row_of_slot[k] = newrow;
    if (src_slot) src_slot[k] = scratch_src[k];
    if (flags & SAPT_F_TAKE_POST_LIK) lik_post[k] = scratch_post[k];  // LG:813-815
    if (slot_of_row_local && newrow >= row0 && newrow < row0 + R_local) slot_of_row_local[newrow - row0] = k;
  }
}

// -------------------------------------------------------------------------------------------
// The same sweep without a serial walk (large ladders: 65 536 chains took 3.4 ms in the walk above).
// After a pair does NOT swap, the carried message is simply chain k+1's: the state "reset at m" (about to
// process pair m carrying chain m's own message) does not depend on the past. From such a state every pair swaps
// until the first j >= m whose threshold is not above chain m's likelihood, i.e. the first j >= m with
// kt[j] <= kp[m] — a next-smaller-or-equal query, answered for ALL m at once with a range-minimum sparse table.
// nxt[m] = that j + 1 is the next reset state, so the states the real sweep visits are the list 0 -> nxt[0] -> ...;
// it is marked by pointer doubling (S <- S u J(S), J <- J o J). Then: pair j did not swap  <=>  state j+1 is on
// the list; the message left at position k is chain k+1's if pair k swapped, else the message carried at that
// time = the last list state <= k (a prefix maximum). Same decision rule as the serial walk (integer keys), so
// the two kernels agree bit for bit; pairs that need the literal formula (u <= 0) make the kernel fall back to it.
// One CTA of 1024 threads; all phases are O(R log R / 1024) with R-sized global scratch.
constexpr int kParThreads = 1024;
constexpr int kSweepGridMin = 16384;      // from here on the grid-wide phase kernels (below) replace the one-CTA kernel
constexpr long long kKeyMin = (long long)0x8000000000000000ULL, kKeyMax = 0x7fffffffffffffffLL;

__global__ void __launch_bounds__(kParThreads) swap_sweep_parallel_kernel(
    int R, double* lik_post, const double* temperature, int32_t* row_of_slot, uint32_t flags, uint32_t seed, int sweep_base,
    const int32_t* sweep_dev, const double* u_inject, uint8_t* swapped, int32_t* num_swap, int32_t* src_slot, int row0, int R_local,
    int32_t* slot_of_row_local, int32_t* scratch_src /*[2R]*/, double* scratch_post /*[R]*/, long long* st /*[levels][R]*/,
    long long* kp /*[R]*/, int32_t* jmp /*[2][R+1]*/, uint8_t* mark /*[2][R+1]*/, int32_t* literal_flag) {
  const int sweep = sweep_base + (sweep_dev ? *sweep_dev : 0);
  const int tid = threadIdx.x;
  const int NP = R - 1;   // pairs
  auto posted = [&](int k) { return (flags & SAPT_F_POST_TEMPERED) ? lik_post[k] : lik_post[k] * temperature[k]; };
  __shared__ int s_any_literal, s_ns;
  __shared__ int s_scan[kParThreads];
  if (tid == 0) { s_any_literal = 0; s_ns = 0; }
  __syncthreads();
  // ---- keys: kp[k] of every chain; level 0 of the table = threshold keys (kKeyMin forces "no swap" when u >= 1) ----
  for (int k = tid; k < R; k += kParThreads) {
    const double pk = posted(k);
    kp[k] = (pk != pk) ? kKeyMin : order_key(pk);
  }
  for (int j = tid; j < NP; j += kParThreads) {
    const double pk = posted(j + 1);
    const double u = u_inject ? u_inject[j] : (double)philox_uniform(seed, SAPT_STREAM_SWAP, sweep, 0u, (uint32_t)j);
    if (!(u > 0.0)) s_any_literal = 1;
    const double t = pk - log(2.0 * u);
    long long kt = (t != t) ? kKeyMax : order_key(t);
    if (!(u < 1.0)) kt = kKeyMin;
    st[j] = kt;
  }
  __syncthreads();
  if (s_any_literal) { if (tid == 0) *literal_flag = 1; return; }   // the host re-runs the serial kernel (never with Philox: u > 0)
  int levels = 1;
  while ((1 << levels) <= NP) ++levels;
  for (int l = 1; l < levels; ++l) {
    const long long* a = st + (int64_t)(l - 1) * R;
    long long* b = st + (int64_t)l * R;
    const int half = 1 << (l - 1), span = 1 << l;
    for (int j = tid; j + span <= NP; j += kParThreads) b[j] = min(a[j], a[j + half]);
    __syncthreads();
  }
  // ---- nxt[m] = 1 + first j >= m with kt[j] <= kp[m]  (a swap needs kp[m] < kt[j]); R = no such pair ----
  int32_t* J0 = jmp; int32_t* J1 = jmp + (R + 1);
  uint8_t* M0 = mark; uint8_t* M1 = mark + (R + 1);
  for (int m = tid; m <= R; m += kParThreads) {
    int nx = R;
    if (m < NP) {
      const long long c = kp[m];
      int j = m;
      for (int l = levels - 1; l >= 0; --l)
        if (j + (1 << l) <= NP && st[(int64_t)l * R + j] > c) j += (1 << l);
      nx = (j < NP) ? j + 1 : R;
    }
    J0[m] = nx;
    M0[m] = (m == 0) ? 1 : 0;
  }
  __syncthreads();
  // ---- states on the list from 0: pointer doubling ----
  for (int l = 0; l < levels + 1; ++l) {
    for (int m = tid; m <= R; m += kParThreads) M1[m] = M0[m];
    __syncthreads();
    for (int m = tid; m < R; m += kParThreads) {
      const int t = J0[m];
      if (M0[m] && t < R) M1[t] = 1;
      J1[m] = (t < R) ? J0[t] : R;
    }
    if (tid == 0) J1[R] = R;
    __syncthreads();
    int32_t* tj = J0; J0 = J1; J1 = tj;
    uint8_t* tm = M0; M0 = M1; M1 = tm;
  }
  // ---- decisions, and the carried chain at every position: prefix maximum of the list states ----
  // scratch_src[k] temporarily holds (state k on the list ? k : -1); chunked inclusive max-scan
  int carry = -1;
  for (int base = 0; base < R; base += kParThreads) {
    const int k = base + tid;
    int v = (k < R && M0[k]) ? k : -1;
    s_scan[tid] = v;
    __syncthreads();
    for (int off = 1; off < kParThreads; off <<= 1) {
      const int o = (tid >= off) ? s_scan[tid - off] : -1;
      __syncthreads();
      s_scan[tid] = max(s_scan[tid], o);
      __syncthreads();
    }
    const int last = max(carry, s_scan[tid]);
    const int chunk_max = s_scan[kParThreads - 1];
    if (k < R) {
      int src;
      if (k < NP) {
        const bool sw = !M0[k + 1];                    // pair k swapped <=> no reset at k+1
        src = sw ? k + 1 : last;
        if (swapped) swapped[k] = sw ? 1 : 0;
        if (sw) atomicAdd(&s_ns, 1);
      } else {
        src = last;                                    // the message carried out of the last pair
      }
      scratch_src[k] = src;
      scratch_post[k] = posted(src);
    }
    carry = max(carry, chunk_max);
    __syncthreads();
  }
  if (tid == 0 && num_swap) *num_swap += s_ns;
  // ---- new binding (as in the serial kernel) ----
  int32_t* old_rows = scratch_src + R;
  for (int k = tid; k < R; k += kParThreads) old_rows[k] = row_of_slot[k];
  __syncthreads();
  for (int k = tid; k < R; k += kParThreads) {
    const int newrow = old_rows[scratch_src[k]];
    row_of_slot[k] = newrow;
    if (src_slot) src_slot[k] = scratch_src[k];
    if (flags & SAPT_F_TAKE_POST_LIK) lik_post[k] = scratch_post[k];  // LG:813-815
    if (slot_of_row_local && newrow >= row0 && newrow < row0 + R_local) slot_of_row_local[newrow - row0] = k;
  }
}

// -------------------------------------------------------------------------------------------
// The same algorithm over the whole GPU for very long ladders (multi-GPU runs: every rank replays the sweep of ALL chains, 65 536 at
// 8 x 8192). The one-CTA kernel above is bound by global-memory latency (~25 cycles per chain and phase: 1.6 ms at 65 536 chains);
// here every phase is its own grid-wide launch — ~40 short kernels, stream ordered, capturable. Same integer keys, same decisions.
struct SweepGrid {
  int R; double* lik_post; const double* temperature; int32_t* row_of_slot; uint32_t flags, seed; int sweep_base; const int32_t* sweep_dev;
  const double* u_inject; uint8_t* swapped; int32_t* num_swap; int32_t* src_slot; int row0, R_local; int32_t* slot_of_row_local;
  int32_t* scratch_src; double* scratch_post; long long* st; long long* kp; int32_t* jmp; uint8_t* mark; int32_t* literal_flag; int32_t* blk;
};
__device__ __forceinline__ double sweep_posted(const SweepGrid& a, int k) {
  return (a.flags & SAPT_F_POST_TEMPERED) ? a.lik_post[k] : a.lik_post[k] * a.temperature[k];
}
constexpr int kScanBlock = 1024;

__global__ void sweep_grid_keys_kernel(SweepGrid a) {
  const int sweep = a.sweep_base + (a.sweep_dev ? *a.sweep_dev : 0);
  const int NP = a.R - 1;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= a.R; k += gridDim.x * blockDim.x) {
    if (k < a.R) {
      const double pk = sweep_posted(a, k);
      a.kp[k] = (pk != pk) ? kKeyMin : order_key(pk);
    }
    if (k < NP) {
      const double pk = sweep_posted(a, k + 1);
      const double u = a.u_inject ? a.u_inject[k] : (double)philox_uniform(a.seed, SAPT_STREAM_SWAP, sweep, 0u, (uint32_t)k);
      if (!(u > 0.0)) *a.literal_flag = 1;
      const double t = pk - log(2.0 * u);
      long long kt = (t != t) ? kKeyMax : order_key(t);
      if (!(u < 1.0)) kt = kKeyMin;
      a.st[k] = kt;
    }
  }
}
__global__ void sweep_grid_level_kernel(SweepGrid a, int l) {
  if (*a.literal_flag) return;
  const int NP = a.R - 1;
  const long long* lo = a.st + (int64_t)(l - 1) * a.R;
  long long* hi = a.st + (int64_t)l * a.R;
  const int half = 1 << (l - 1), span = 1 << l;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j + span <= NP; j += gridDim.x * blockDim.x) hi[j] = min(lo[j], lo[j + half]);
}
__global__ void sweep_grid_next_kernel(SweepGrid a, int levels) {
  if (*a.literal_flag) return;
  const int R = a.R, NP = R - 1;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m <= R; m += gridDim.x * blockDim.x) {
    int nx = R;
    if (m < NP) {
      const long long c = a.kp[m];
      int j = m;
      for (int l = levels - 1; l >= 0; --l)
        if (j + (1 << l) <= NP && a.st[(int64_t)l * R + j] > c) j += (1 << l);
      nx = (j < NP) ? j + 1 : R;
    }
    a.jmp[m] = nx;
    a.mark[m] = (m == 0) ? 1 : 0;
  }
}
// one round of pointer doubling. The marks are monotone (a state, once known to be on the list, stays), so they are updated in place:
// a mark read before or after a concurrent set both give a subset of the true list and a superset of the synchronous round.
__global__ void sweep_grid_double_kernel(SweepGrid a, int32_t* j0, int32_t* j1) {
  if (*a.literal_flag) return;
  const int R = a.R;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m <= R; m += gridDim.x * blockDim.x) {
    if (m == R) { j1[R] = R; continue; }
    const int t = j0[m];
    if (t < R && a.mark[m]) a.mark[t] = 1;
    j1[m] = (t < R) ? j0[t] : R;
  }
}
__global__ void __launch_bounds__(kScanBlock) sweep_grid_blockmax_kernel(SweepGrid a) {   // last list state inside each block of kScanBlock chains
  if (*a.literal_flag) return;
  __shared__ int s_w[32];
  const int k = blockIdx.x * kScanBlock + threadIdx.x;
  int v = (k < a.R && a.mark[k]) ? k : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = s_w[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (threadIdx.x == 0) a.blk[blockIdx.x] = w;
  }
}
__global__ void sweep_grid_blockscan_kernel(SweepGrid a, int n_blk) {   // exclusive prefix maximum over the blocks (a few hundred at most)
  if (*a.literal_flag) return;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int carry = -1;
    for (int b = 0; b < n_blk; ++b) { const int v = a.blk[b]; a.blk[b] = carry; carry = max(carry, v); }
  }
}
__global__ void __launch_bounds__(kScanBlock) sweep_grid_decide_kernel(SweepGrid a) {
  if (*a.literal_flag) return;
  __shared__ int s_w[32];
  __shared__ int s_ns;
  const int R = a.R, NP = R - 1;
  const int k = blockIdx.x * kScanBlock + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_ns = 0;
  int v = (k < R && a.mark[k]) ? k : -1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = max(v, t); }
  if (lane == 31) s_w[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w = max(w, t); }
    s_w[lane] = w;
  }
  __syncthreads();
  const int last = max(max(a.blk[blockIdx.x], warp > 0 ? s_w[warp - 1] : -1), v);    // last list state <= k
  if (k < R) {
    int src;
    if (k < NP) {
      const bool sw = !a.mark[k + 1];                  // pair k swapped <=> no reset at k+1
      src = sw ? k + 1 : last;
      if (a.swapped) a.swapped[k] = sw ? 1 : 0;
      if (sw) atomicAdd(&s_ns, 1);
    } else {
      src = last;                                      // the message carried out of the last pair
    }
    a.scratch_src[k] = src;
    a.scratch_post[k] = sweep_posted(a, src);
    a.scratch_src[R + k] = a.row_of_slot[k];           // old binding
  }
  __syncthreads();
  if (threadIdx.x == 0 && a.num_swap && s_ns) atomicAdd(a.num_swap, s_ns);
}
__global__ void sweep_grid_bind_kernel(SweepGrid a) {
  if (*a.literal_flag) return;
  const int R = a.R;
  const int32_t* old_rows = a.scratch_src + R;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < R; k += gridDim.x * blockDim.x) {
    const int newrow = old_rows[a.scratch_src[k]];
    a.row_of_slot[k] = newrow;
    if (a.src_slot) a.src_slot[k] = a.scratch_src[k];
    if (a.flags & SAPT_F_TAKE_POST_LIK) a.lik_post[k] = a.scratch_post[k];  // LG:813-815
    if (a.slot_of_row_local && newrow >= a.row0 && newrow < a.row0 + a.R_local) a.slot_of_row_local[newrow - a.row0] = k;
  }
}

__global__ void philox_fill_kernel(uint32_t seed, int kind, int step, int chain, int n, int normal, float* out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (normal) {
    if (4 * q >= n) return;
    float z[4];
    philox_normal4(seed, kind, step, (uint32_t)chain, (uint32_t)q, z);
    for (int j = 0; j < 4; ++j) if (4 * q + j < n) out[4 * q + j] = z[j];
  } else {
    if (q < n) out[q] = philox_uniform(seed, kind, step, (uint32_t)chain, (uint32_t)q);
  }
}

}  // namespace sapt

using namespace sapt;

extern "C" int sapt_prior(const float* theta, int64_t ld_theta, int R, int P, double sigma_squared, double* prior, void* stream) {
  SAPT_REQUIRE(theta && prior && R >= 0 && P > 0 && ld_theta >= P, "sapt_prior");
  if (R == 0) return SAPT_OK;
  prior_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(theta, ld_theta, P, sigma_squared, prior);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_propose(const sapt_propose_args* a, void* stream) {
  SAPT_REQUIRE(a && a->base && a->proposal && a->prior_prop && a->R >= 0 && a->P > 0 && a->ld >= a->P, "sapt_propose");
  if (a->decide) SAPT_REQUIRE(a->is_sur && a->true_rows && a->n_true && a->sur_rows && a->n_sur, "sapt_propose: decision outputs");
  if (a->R == 0) return SAPT_OK;
  propose_kernel<<<a->R, 256, 0, (cudaStream_t)stream>>>(*a);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_lg_first_eval(const sapt_lg_first_args* a, void* stream) {
  SAPT_REQUIRE(a && a->proposal && a->L_train && a->surg_first && a->st.slot_of_row && a->st.temperature, "sapt_lg_first_eval");
  if (a->use_langevin) SAPT_REQUIRE(a->lang_rows && a->n_lang, "sapt_lg_first_eval: langevin outputs");
  if (a->R == 0) return SAPT_OK;
  lg_first_kernel<<<a->R, 128, 0, (cudaStream_t)stream>>>(*a);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_mh_accept(const sapt_accept_args* a, void* stream) {
  SAPT_REQUIRE(a && a->theta && a->proposal && a->prior_prop && a->L_train, "sapt_mh_accept");
  SAPT_REQUIRE(a->st.slot_of_row && a->st.temperature && a->st.lik_cur && a->st.lik_post && a->st.prior_cur && a->st.mov &&
               a->st.likl_wo_temp && a->st.naccept && a->st.last_rmse_train && a->st.last_rmse_test &&
               a->st.last_acc_train && a->st.last_acc_test, "sapt_mh_accept: chain state");
  SAPT_REQUIRE(a->nc_train && a->sq_train && a->n_train > 0, "sapt_mh_accept: train stats");
  if (a->n_test > 0) SAPT_REQUIRE(a->L_test && a->nc_test && a->sq_test, "sapt_mh_accept: test stats");
  if (a->R == 0) return SAPT_OK;
  mh_accept_kernel<<<a->R, 128, 0, (cudaStream_t)stream>>>(*a);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_advance(int32_t* counter, int delta, void* stream) {
  SAPT_REQUIRE(counter, "sapt_advance");
  advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_swap_sweep(int R_global, double* lik_post, const double* temperature, int32_t* row_of_slot, uint32_t flags,
                               uint32_t seed, int sweep, const double* u_inject, uint8_t* swapped, int32_t* num_swap,
                               int32_t* src_slot, int row0, int R_local, int32_t* slot_of_row_local, void* stream) {
  return sapt_swap_sweep_dev(R_global, lik_post, temperature, row_of_slot, flags, seed, sweep, nullptr, u_inject, swapped, num_swap,
                             src_slot, row0, R_local, slot_of_row_local, stream);
}

extern "C" int sapt_swap_sweep_dev(int R_global, double* lik_post, const double* temperature, int32_t* row_of_slot, uint32_t flags,
                                   uint32_t seed, int sweep, const int32_t* sweep_dev, const double* u_inject, uint8_t* swapped,
                                   int32_t* num_swap, int32_t* src_slot, int row0, int R_local, int32_t* slot_of_row_local,
                                   void* stream) {
  SAPT_REQUIRE(R_global >= 1 && lik_post && temperature && row_of_slot, "sapt_swap_sweep");
  // Scratch: src[R] + old_rows[R] (int32), post[R] (double) and, for the parallel kernel, the range-minimum table, jump lists
  // and marks. One set per (device, stream) of this process, found under a mutex: sweeps enqueued on one stream are ordered by
  // the stream, sweeps on different streams (two engines in one process) never share a buffer. A set grows on demand: the call
  // that grows it synchronises ITS stream and reallocates, so the first sweep of a given size on a stream is not capturable
  // into a CUDA graph (later ones are) — see sapt.h.
  struct SweepScratch { int32_t* sc_i = nullptr; double* sc_d = nullptr; int sc_cap = 0;
                        long long* sp_st = nullptr; int32_t* sp_i = nullptr; uint8_t* sp_m = nullptr; int sp_cap = 0; };
  static std::mutex mu;
  static std::map<std::pair<int, void*>, SweepScratch> table;
  int dev = 0;
  SAPT_CUDA(cudaGetDevice(&dev));
  std::unique_lock<std::mutex> lock(mu);
  SweepScratch& sc = table[std::make_pair(dev, stream)];
  lock.unlock();     // the entry is only ever touched by callers of this (device, stream); std::map never moves it
  int32_t*& sc_i = sc.sc_i; double*& sc_d = sc.sc_d; int& sc_cap = sc.sc_cap;
  if (R_global > sc_cap) {
    SAPT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (sc_i) { SAPT_CUDA(cudaFree(sc_i)); SAPT_CUDA(cudaFree(sc_d)); }
    const int cap = R_global + R_global / 4 + 64;
    SAPT_CUDA(cudaMalloc(&sc_i, sizeof(int32_t) * 2 * cap));
    SAPT_CUDA(cudaMalloc(&sc_d, sizeof(double) * cap));
    sc_cap = cap;
  }
  // NOTE: old_rows lives at sc_i + R_global (the kernel derives it from scratch_src + R)
  const bool parallel = (flags & SAPT_F_SWEEP_PARALLEL) || R_global >= 4096;
  if (parallel && R_global >= 3) {
    int levels = 1;
    while ((1 << levels) <= R_global - 1) ++levels;
    if (R_global > sc.sp_cap) {
      SAPT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
      if (sc.sp_st) { SAPT_CUDA(cudaFree(sc.sp_st)); SAPT_CUDA(cudaFree(sc.sp_i)); SAPT_CUDA(cudaFree(sc.sp_m)); }
      const int cap = R_global + R_global / 4 + 64;
      int lv = 1;
      while ((1 << lv) <= cap) ++lv;
      SAPT_CUDA(cudaMalloc(&sc.sp_st, sizeof(long long) * (size_t)(lv + 1) * cap));       // table levels + kp
      SAPT_CUDA(cudaMalloc(&sc.sp_i, sizeof(int32_t) * (2 * (size_t)(cap + 1) + 4 + cap / kScanBlock + 8)));      // the literal flag, jump arrays, block maxima
      SAPT_CUDA(cudaMalloc(&sc.sp_m, 2 * (size_t)(cap + 1)));
      SAPT_CUDA(cudaMemsetAsync(sc.sp_i, 0, sizeof(int32_t) * (2 * (size_t)(cap + 1) + 4 + cap / kScanBlock + 8), (cudaStream_t)stream));
      sc.sp_cap = cap;
    }
    long long* kp = sc.sp_st + (size_t)levels * R_global;
    int32_t* lit = sc.sp_i;                    // fixed slot; the jump arrays follow it
    if (R_global >= kSweepGridMin || (flags & SAPT_F_SWEEP_GRID)) {
      cudaStream_t cs = (cudaStream_t)stream;
      SweepGrid a{R_global, lik_post, temperature, row_of_slot, flags, seed, sweep, sweep_dev, u_inject, swapped, num_swap, src_slot, row0, R_local,
                  slot_of_row_local, sc_i, sc_d, sc.sp_st, kp, sc.sp_i + 4, sc.sp_m, lit, sc.sp_i + 4 + 2 * ((size_t)sc.sp_cap + 1)};
      const int tpb = 256;
      const int grid = std::max(1, std::min((R_global + tpb) / tpb, 4 * 148));
      const int n_blk = (R_global + kScanBlock - 1) / kScanBlock;
      sweep_grid_keys_kernel<<<grid, tpb, 0, cs>>>(a);
      for (int l = 1; l < levels; ++l) sweep_grid_level_kernel<<<grid, tpb, 0, cs>>>(a, l);
      sweep_grid_next_kernel<<<grid, tpb, 0, cs>>>(a, levels);
      int32_t* j0 = a.jmp; int32_t* j1 = a.jmp + (R_global + 1);
      for (int l = 0; l < levels + 1; ++l) { sweep_grid_double_kernel<<<grid, tpb, 0, cs>>>(a, j0, j1); std::swap(j0, j1); }
      sweep_grid_blockmax_kernel<<<n_blk, kScanBlock, 0, cs>>>(a);
      sweep_grid_blockscan_kernel<<<1, 32, 0, cs>>>(a, n_blk);
      sweep_grid_decide_kernel<<<n_blk, kScanBlock, 0, cs>>>(a);
      sweep_grid_bind_kernel<<<grid, tpb, 0, cs>>>(a);
      SAPT_CHECK_LAUNCH();
    } else
    swap_sweep_parallel_kernel<<<1, kParThreads, 0, (cudaStream_t)stream>>>(R_global, lik_post, temperature, row_of_slot, flags, seed, sweep,
                                                                            sweep_dev, u_inject, swapped, num_swap, src_slot, row0, R_local,
                                                                            slot_of_row_local, sc_i, sc_d, sc.sp_st, kp, sc.sp_i + 4, sc.sp_m, lit);
    SAPT_CHECK_LAUNCH();
    if (!u_inject) return SAPT_OK;              // Philox uniforms lie in (0, 1]: the literal formula is never needed
    int32_t h_lit = 0;                          // injected streams (parity tests): u <= 0 is possible, check and redo serially
    SAPT_CUDA(cudaMemcpyAsync(&h_lit, lit, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    SAPT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (!h_lit) return SAPT_OK;
    SAPT_CUDA(cudaMemsetAsync(lit, 0, sizeof(int32_t), (cudaStream_t)stream));
  }
  swap_sweep_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(R_global, lik_post, temperature, row_of_slot, flags, seed, sweep, sweep_dev,
                                                         u_inject, swapped, num_swap, src_slot, row0, R_local, slot_of_row_local,
                                                         sc_i, sc_d);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_philox_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr, key[0], key[1], out);
  return SAPT_OK;
}

extern "C" int sapt_philox_fill(uint32_t seed, int kind, int step, int chain, int n, int normal, float* out, void* stream) {
  SAPT_REQUIRE(out && n >= 0, "sapt_philox_fill");
  if (n == 0) return SAPT_OK;
  const int work = normal ? (n + 3) / 4 : n;
  philox_fill_kernel<<<(work + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, kind, step, chain, n, normal, out);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}
