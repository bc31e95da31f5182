// stan4bart_b200/csrc/glmm.cu
// Device evaluation of the GLMM log-density data terms for stan4bart's continuous.stan and
// the O(d) host-side chain rule around them.
//
//   reference density   /root/reference/src/stan_files/continuous.stan:344-366 (eta, normal_lpdf),
//                       continuous.hpp:2168-2638 (log_prob_impl), :528-806 (make_theta_L, make_b),
//                       :823-916 (decov_lp), :2640-2938 (write_array), :3626-3768 (offset / response /
//                       parametric mean)
//   reference gradient  reverse-mode AD, src/include/stan/model/gradient.hpp:21-35
//
// Per evaluation one N-length pass produces S = sum e^2, X'e and Z'e (e = y - offset - X beta - Z b);
// everything else is O(K + q) and stays on the host next to the NUTS control logic.
// Z is stored as ELL (one int32 column index + one fp64 value per non-zero slot, slot-major so every
// stream is coalesced); the per-column sums are segmented reductions into lane-private shared-memory
// bins (no atomics, fixed order => deterministic).
#include "glmm.hpp"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <stdexcept>

namespace s4b {

constexpr int kGBlock = 256;
constexpr int kGBulkWarps = 16;      // consumer warps of the bulk-copy data pass (its consumers are latency bound: 8 warps left the issue slots 70 % idle)
constexpr int kGBulk = kGBulkWarps * 32;
constexpr int kMaxNc = 16;         // coefficients per ranef block
constexpr size_t kThetaSmemMax = 40 * 1024;   // coefficient vector staged in (default-limit) shared memory up to this size
constexpr int kGFast = 4;          // fast path of the data pass: K and non-zeros per row of Z up to this

__device__ __forceinline__ double g_warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Shared tail of the binned data passes: the (nb + 1) x 32 lane-private bins of every consumer warp -> one partial row per CTA -> the
// last CTA to finish sums the rows of all CTAs in CTA order.  `bins0` = the bins of warp 0, `wb` = this warp's (ignored when the
// warp holds none: `has_bins` false, e.g. a producer warp); every thread of the block calls this.
template <int WARPS>
__device__ __forceinline__ void data_terms_epilogue(const GlmmDev& g, double* bins0, double* wb, bool has_bins)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = g.K + g.q;
  if (has_bins) {
    __syncwarp();
    // warp reduction, transposed: lane l sums row j0 + l over its 32 copies, starting at copy l (conflict-free: the 32 lanes touch 32
    // different banks), in a fixed order
    for (int j0 = 0; j0 <= nb; j0 += 32) {
      const int j = j0 + lane;
      double v = 0.0;
      if (j <= nb) {
#pragma unroll 8
        for (int c = 0; c < 32; ++c) v += wb[j * 32 + ((c + lane) & 31)];
      }
      __syncwarp();
      if (j <= nb) wb[j * 32] = v;
    }
  }
  __syncthreads();
  const int G = gridDim.x;
  for (int j = tid; j <= nb; j += blockDim.x) {
    double acc = 0.0;
    for (int w = 0; w < WARPS; ++w) acc += bins0[(size_t) w * (nb + 1) * 32 + j * 32];
    // value order in partials / result: S, X'e, Z'e
    int out_j = j == nb ? 0 : j + 1;
    g.partials[(long long) out_j * G + blockIdx.x] = acc;
  }
  __shared__ unsigned int s_ticket;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(g.ticket, 1u);
  __syncthreads();
  if (s_ticket != (unsigned int) G - 1) return;
  __threadfence();
  // one warp per row of partials, four rows at a time with every load issued before the first sum: the rows were written by other SMs a
  // moment ago, so each round of loads is one L2 round trip -- serialising them per row costs microseconds at the tail of the pass
  const int W = blockDim.x / 32;
  for (int v0 = warp; v0 <= nb; v0 += 4 * W) {
    double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (int b = lane; b < G; b += 32) {
#pragma unroll
      for (int r = 0; r < 4; ++r) { const int v = v0 + r * W; if (v <= nb) acc[r] += __ldcg(g.partials + (long long) v * G + b); }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int v = v0 + r * W;
      const double a = g_warp_sum(acc[r]);
      if (v <= nb && lane == 0) g.result[v] = a;
    }
  }
  if (tid == 0) *g.ticket = 0u;
}

// algorithmic bytes per observation: r 8 + X 8K + Z (4 + 8 [0 if the slot is an indicator]) per slot
// KT, ST: compile-time bounds of the fast path (K <= KT dense columns, <= ST non-zeros per row of Z), so that the register arrays of
// the loads in flight are as small as the model allows; <4, 4> also carries the general path
template <int KT, int ST>
__global__ void __launch_bounds__(kGBlock, 2) k_glmm_data_terms(GlmmDev g)
{
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = g.K + g.q;
  double* sth = smem;                                   // [nb] beta, b
  double* wb = smem + nb + (size_t) warp * (nb + 1) * 32; // lane-private bins of this warp
  for (int j = tid; j < nb; j += kGBlock) sth[j] = g.theta[j];
  for (int j = lane; j < (nb + 1) * 32; j += 32) wb[j] = 0.0;
  __syncthreads();

  const long long N = g.N, npad = g.npad;
  const int K = g.K, slots = g.slots;
  double S = 0.0;
  double gx[KT];                                     // fast path: X'e in registers (the dense columns are the same for every row)
#pragma unroll
  for (int k = 0; k < KT; ++k) gx[k] = 0.0;
  if (K <= KT && slots <= ST) {
    // fast path (K, non-zeros per row <= 4: every BASELINE config): two observations per thread and load, two such pairs
    // in flight per iteration, every global load of the iteration issued before the first use -- the pass is bound by
    // memory latency otherwise (ncu: long-scoreboard stalls, 1.1 TB/s)
    const long long stride = (long long) gridDim.x * kGBlock * 2;
    for (long long i0 = ((long long) blockIdx.x * kGBlock + tid) * 2; i0 < N; i0 += 2 * stride) {
      double2 r2[2], w2[2], x2[2][KT], v2[2][ST]; int2 c2[2][ST];
      bool live[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long i = i0 + u * stride;
        live[u] = i < N;
        if (live[u]) {
          r2[u] = __ldg(reinterpret_cast<const double2*>(g.r + i));
          w2[u] = g.wt != nullptr ? __ldg(reinterpret_cast<const double2*>(g.wt + i)) : make_double2(1.0, 1.0);
#pragma unroll
          for (int k = 0; k < KT; ++k) if (k < K) x2[u][k] = __ldg(reinterpret_cast<const double2*>(g.X + (long long) k * npad + i));
#pragma unroll
          for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) {
            c2[u][s_] = __ldg(reinterpret_cast<const int2*>(g.zidx + (long long) s_ * npad + i));
            v2[u][s_] = ((g.ones_mask >> s_) & 1u) ? make_double2(1.0, 1.0) : __ldg(reinterpret_cast<const double2*>(g.zval + (long long) s_ * npad + i));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!live[u]) continue;
        const long long i = i0 + u * stride;
        const bool second = i + 1 < N;            // the padded tail holds zeros, but indicator slots are forced to 1
        double eta0 = 0.0, eta1 = 0.0;
#pragma unroll
        for (int k = 0; k < KT; ++k) if (k < K) { eta0 += x2[u][k].x * sth[k]; eta1 += x2[u][k].y * sth[k]; }
#pragma unroll
        for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) { eta0 += v2[u][s_].x * sth[K + c2[u][s_].x]; eta1 += v2[u][s_].y * sth[K + c2[u][s_].y]; }
        double e0 = r2[u].x - eta0, e1 = second ? r2[u].y - eta1 : 0.0;
        if (g.wt != nullptr) {       // continuous.stan:365: sum w e^2; its gradient carries w e
          const double we0 = w2[u].x * e0, we1 = w2[u].y * e1;
          S += we0 * e0; S += we1 * e1;
          e0 = we0; e1 = we1;
        } else { S += e0 * e0; S += e1 * e1; }
#pragma unroll
        for (int k = 0; k < KT; ++k) if (k < K) { gx[k] = fma(x2[u][k].x, e0, gx[k]); gx[k] = fma(x2[u][k].y, e1, gx[k]); }
        // the non-zeros of one row of Z sit in distinct columns (compressed sparse rows), so a row's bin updates are independent of each
        // other: all loads, then all stores -- one shared-memory round trip per row instead of one per non-zero; the second row of the
        // pair may share columns with the first and follows it
        if (g.row_distinct) {
          double b0[ST];
#pragma unroll
          for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) b0[s_] = wb[(K + c2[u][s_].x) * 32 + lane];
#pragma unroll
          for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) wb[(K + c2[u][s_].x) * 32 + lane] = fma(v2[u][s_].x, e0, b0[s_]);
#pragma unroll
          for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) b0[s_] = wb[(K + c2[u][s_].y) * 32 + lane];
#pragma unroll
          for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) wb[(K + c2[u][s_].y) * 32 + lane] = fma(v2[u][s_].y, e1, b0[s_]);
        } else {
#pragma unroll
          for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) {
            wb[(K + c2[u][s_].x) * 32 + lane] += v2[u][s_].x * e0;
            wb[(K + c2[u][s_].y) * 32 + lane] += v2[u][s_].y * e1;
          }
        }
      }
    }
  } else if (KT == kGFast && ST == kGFast)
  for (long long i = (long long) blockIdx.x * kGBlock + tid; i < N; i += (long long) gridDim.x * kGBlock) {
    double eta = 0.0;
    for (int k = 0; k < K; ++k) eta += __ldg(g.X + (long long) k * npad + i) * sth[k];
    for (int s = 0; s < slots; ++s) {
      int c = __ldg(g.zidx + (long long) s * npad + i);
      double v = ((g.ones_mask >> s) & 1u) ? 1.0 : __ldg(g.zval + (long long) s * npad + i);
      eta += v * sth[K + c];
    }
    double e = __ldg(g.r + i) - eta;
    if (g.wt != nullptr) { const double we = __ldg(g.wt + i) * e; S += we * e; e = we; } else S += e * e;
    for (int k = 0; k < K; ++k) wb[k * 32 + lane] += __ldg(g.X + (long long) k * npad + i) * e;
    for (int s = 0; s < slots; ++s) {
      int c = __ldg(g.zidx + (long long) s * npad + i);
      double v = ((g.ones_mask >> s) & 1u) ? 1.0 : __ldg(g.zval + (long long) s * npad + i);
      wb[(K + c) * 32 + lane] += v * e;
    }
  }
  if (K <= KT && slots <= ST) {
#pragma unroll
    for (int k = 0; k < KT; ++k) if (k < K) wb[k * 32 + lane] = gx[k];
  }
  wb[nb * 32 + lane] = S;
  data_terms_epilogue<kGBlock / 32>(g, smem + nb, wb, true);
}

// ---------------------------------------------------------------------------------------
// The same pass with the operand streams staged through shared memory by the bulk-copy engine (TMA, cp.async.bulk + mbarrier):
// one persistent CTA per SM, 16 consumer warps + 1 producer warp, a ring of `stages` tiles of `tile` observations.  The producer
// keeps `stages` tiles (~45 KB each for the Friedman model) in flight per SM regardless of what the consumers are doing, so the
// memory system never waits for the arithmetic and the consumers need no registers for loads in flight -- the register version
// above alternates between a burst of loads and a burst of shared-memory bin updates (ncu: long scoreboard, 2.1 TB/s).
// Tile layout in shared memory, every stream contiguous: r | weights (if any) | X[0..K) | zidx[0..slots) | zval of the slots whose
// values are not all 1.  Same bins, same fixed-order reductions, same results as k_glmm_data_terms up to the order of the sums.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kGMaxStages = 4;
struct BulkBars { unsigned long long full[kGMaxStages], empty[kGMaxStages]; };

template <int KT, int ST>
__global__ void __launch_bounds__(kGBulk + 32, 1) k_glmm_data_terms_bulk(GlmmDev g, int tile, int stages, int tile_bytes)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BulkBars& bars = *reinterpret_cast<BulkBars*>(smem_raw);
  double* sth = reinterpret_cast<double*>(smem_raw + sizeof(BulkBars));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = g.K + g.q, K = g.K, slots = g.slots;
  double* bins0 = sth + nb;
  double* wb = bins0 + (size_t) (warp < kGBulkWarps ? warp : 0) * (nb + 1) * 32;
  unsigned char* ring = smem_raw + (((sizeof(BulkBars) + sizeof(double) * ((size_t) nb + (size_t) (kGBulkWarps) * (nb + 1) * 32)) + 127) / 128) * 128;
  const bool weighted = g.wt != nullptr;
  if (tid == 0) {
    for (int s_ = 0; s_ < stages; ++s_) { mbar_init(&bars.full[s_], 1u); mbar_init(&bars.empty[s_], (unsigned) (kGBulkWarps)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();                                      // the barriers exist: the producer starts streaming at once
  if (warp < kGBulkWarps) {
    // consumers: coefficients and zeroed bins while the first tiles are in flight
    for (int j = tid; j < nb; j += kGBulk) sth[j] = g.theta[j];
    for (int j = lane; j < (nb + 1) * 32; j += 32) wb[j] = 0.0;
    asm volatile("bar.sync 1, %0;" ::"n"(kGBulk) : "memory");
  }
  const long long N = g.N, npad = g.npad;
  const long long ntiles = (N + tile - 1) / tile;
  // stream offsets inside a tile
  const int off_w = tile * 8, off_x = off_w + (weighted ? tile * 8 : 0), off_i = off_x + K * tile * 8, off_v = off_i + slots * tile * 4;
  if (warp == kGBulkWarps) {
    // ------------------------------------------------------------- producer warp: one lane issues the bulk copies
    if (lane == 0) {
      int k = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++k) {
        const int st = k % stages;
        if (k >= stages) mbar_wait(&bars.empty[st], (unsigned) ((k / stages - 1) & 1));
        const long long base = t * tile;
        const unsigned rows = (unsigned) (npad - base < tile ? npad - base : tile);          // npad and tile are multiples of 16 rows
        unsigned char* dst = ring + (size_t) st * tile_bytes;
        int nval = 0;
        for (int s_ = 0; s_ < slots; ++s_) nval += ((g.ones_mask >> s_) & 1u) ? 0 : 1;
        mbar_expect_tx(&bars.full[st], rows * (8u * (1u + (weighted ? 1u : 0u) + (unsigned) K + (unsigned) nval) + 4u * (unsigned) slots));
        bulk_g2s(dst, g.r + base, rows * 8u, &bars.full[st]);
        if (weighted) bulk_g2s(dst + off_w, g.wt + base, rows * 8u, &bars.full[st]);
        for (int c = 0; c < K; ++c) bulk_g2s(dst + off_x + c * tile * 8, g.X + (long long) c * npad + base, rows * 8u, &bars.full[st]);
        for (int s_ = 0; s_ < slots; ++s_) bulk_g2s(dst + off_i + s_ * tile * 4, g.zidx + (long long) s_ * npad + base, rows * 4u, &bars.full[st]);
        int v = 0;
        for (int s_ = 0; s_ < slots; ++s_) if (!((g.ones_mask >> s_) & 1u)) { bulk_g2s(dst + off_v + v * tile * 8, g.zval + (long long) s_ * npad + base, rows * 8u, &bars.full[st]); ++v; }
      }
    }
    __syncwarp();
  } else {
  // --------------------------------------------------------------- consumer warps
  double S = 0.0;
  double gx[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) gx[k] = 0.0;
  // position of slot s_'s values in the tile (slots whose values are all 1 have none)
  int vpos[ST];
  { int v = 0;
#pragma unroll
    for (int s_ = 0; s_ < ST; ++s_) { vpos[s_] = ((g.ones_mask >> s_) & 1u) ? -1 : v; if (s_ < slots && vpos[s_] >= 0) ++v; } }
  // byte offsets of this thread's pair of rows inside every stream of a tile, formed once: the loop below then addresses shared memory
  // with one add per access (the pass is bound by the consumers' instruction latency: 8 warps per SM, ~2 fixed-latency stall cycles per
  // instruction in ncu, and address arithmetic was most of the instruction stream)
  int offx[KT], offi[ST], offv[ST];
#pragma unroll
  for (int c = 0; c < KT; ++c) offx[c] = off_x + c * tile * 8;
#pragma unroll
  for (int s_ = 0; s_ < ST; ++s_) { offi[s_] = off_i + s_ * tile * 4; offv[s_] = vpos[s_] < 0 ? -1 : off_v + vpos[s_] * tile * 8; }
  double* const wbl = wb + lane + K * 32;               // bin of column K + c: wbl[c * 32]
  const double* const sthK = sth + K;
  int k = 0;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++k) {
    const int st = k % stages;
    mbar_wait(&bars.full[st], (unsigned) ((k / stages) & 1));
    const unsigned char* src = ring + (unsigned) st * (unsigned) tile_bytes;
    const long long base = t * tile;
    const int rows_here = (int) (N - base < tile ? N - base : tile);       // rows of this tile that exist
    for (int o = 2 * tid; o < rows_here; o += 2 * kGBulk) {
      const bool second = o + 1 < rows_here;
      const unsigned char* p8 = src + o * 8;
      const unsigned char* p4 = src + o * 4;
      const double2 r2 = *reinterpret_cast<const double2*>(p8);
      double2 x2[KT], v2[ST]; int2 c2[ST];
#pragma unroll
      for (int c = 0; c < KT; ++c) if (c < K) x2[c] = *reinterpret_cast<const double2*>(p8 + offx[c]);
#pragma unroll
      for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) {
        c2[s_] = *reinterpret_cast<const int2*>(p4 + offi[s_]);
        v2[s_] = offv[s_] < 0 ? make_double2(1.0, 1.0) : *reinterpret_cast<const double2*>(p8 + offv[s_]);
      }
      double eta0 = 0.0, eta1 = 0.0;
#pragma unroll
      for (int c = 0; c < KT; ++c) if (c < K) { eta0 += x2[c].x * sth[c]; eta1 += x2[c].y * sth[c]; }
#pragma unroll
      for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) { eta0 += v2[s_].x * sthK[c2[s_].x]; eta1 += v2[s_].y * sthK[c2[s_].y]; }
      double e0 = r2.x - eta0, e1 = second ? r2.y - eta1 : 0.0;
      if (weighted) {
        const double2 w2 = *reinterpret_cast<const double2*>(p8 + off_w);
        const double we0 = w2.x * e0, we1 = w2.y * e1;
        S += we0 * e0; S += we1 * e1;
        e0 = we0; e1 = we1;
      } else { S += e0 * e0; S += e1 * e1; }
#pragma unroll
      for (int c = 0; c < KT; ++c) if (c < K) { gx[c] = fma(x2[c].x, e0, gx[c]); gx[c] = fma(x2[c].y, e1, gx[c]); }
      if (g.row_distinct) {
        double b0[ST];
#pragma unroll
        for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) b0[s_] = wbl[c2[s_].x * 32];
#pragma unroll
        for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) wbl[c2[s_].x * 32] = fma(v2[s_].x, e0, b0[s_]);
#pragma unroll
        for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) b0[s_] = wbl[c2[s_].y * 32];
#pragma unroll
        for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) wbl[c2[s_].y * 32] = fma(v2[s_].y, e1, b0[s_]);
      } else {
#pragma unroll
        for (int s_ = 0; s_ < ST; ++s_) if (s_ < slots) {
          wbl[c2[s_].x * 32] += v2[s_].x * e0;
          wbl[c2[s_].y * 32] += v2[s_].y * e1;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars.empty[st]);              // this warp has read everything it needs from the stage
  }
#pragma unroll
  for (int c = 0; c < KT; ++c) if (c < K) wb[c * 32 + lane] = gx[c];
  wb[nb * 32 + lane] = S;
  }
  // one call site for the whole block: the epilogue's block-wide barriers must be reached by producer and consumers alike
  data_terms_epilogue<kGBulkWarps>(g, bins0, wb, warp < kGBulkWarps);
}

using DataTermsBulkKernel = void (*)(GlmmDev, int, int, int);
static DataTermsBulkKernel data_terms_bulk_kernel(int K, int slots)
{
  if (K <= 2) return slots <= 2 ? k_glmm_data_terms_bulk<2, 2> : k_glmm_data_terms_bulk<2, 4>;
  return slots <= 2 ? k_glmm_data_terms_bulk<4, 2> : k_glmm_data_terms_bulk<4, 4>;
}

using DataTermsKernel = void (*)(GlmmDev);
static DataTermsKernel data_terms_kernel(int K, int slots)
{
  if (K > kGFast || slots > kGFast) return k_glmm_data_terms<4, 4>;
  if (K <= 2) return slots <= 2 ? k_glmm_data_terms<2, 2> : k_glmm_data_terms<2, 4>;
  return slots <= 2 ? k_glmm_data_terms<4, 2> : k_glmm_data_terms<4, 4>;
}

// ---------------------------------------------------------------------------------------
// Column path of the data pass: models whose K + q exceeds what the lane-private shared-memory bins hold (~110 columns)
// -- many grouping levels.  Pass 1 writes the weighted residual w e and block partials of S = sum w e^2; the dense columns
// X'(w e) are chunked dot products; every column of Z (compressed sparse columns, entries in observation order, built once)
// is summed by one warp in a fixed order.  No atomics: results are deterministic.
// algorithmic bytes per observation: pass 1 as the binned pass + 8 (w e written), pass 2: 16 K + (4 + 8 + 8) per non-zero of Z
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double g_block_sum(double v, double* red /* [kGBlock / 32] */)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = g_warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double acc = 0.0;
  for (int w = 0; w < kGBlock / 32; ++w) acc += red[w];
  return acc;
}

__global__ void __launch_bounds__(kGBlock) k_glmm_resid_we(GlmmDev g, double* __restrict__ we_out, double* __restrict__ partials)
{
  __shared__ double red[kGBlock / 32];
  const long long N = g.N, npad = g.npad;
  const int K = g.K, slots = g.slots;
  double S = 0.0;
  for (long long i = (long long) blockIdx.x * kGBlock + threadIdx.x; i < N; i += (long long) gridDim.x * kGBlock) {
    double eta = 0.0;
    for (int k = 0; k < K; ++k) eta += __ldg(g.X + (long long) k * npad + i) * __ldg(g.theta + k);
    for (int s = 0; s < slots; ++s) {
      const int c = __ldg(g.zidx + (long long) s * npad + i);
      const double v = ((g.ones_mask >> s) & 1u) ? 1.0 : __ldg(g.zval + (long long) s * npad + i);
      eta += v * __ldg(g.theta + K + c);
    }
    const double e = __ldg(g.r + i) - eta;
    const double we = g.wt != nullptr ? __ldg(g.wt + i) * e : e;
    we_out[i] = we;
    S += we * e;
  }
  S = g_block_sum(S, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = S;
}

// block (c, j): partial dot product of dense column j with w e over chunk c
__global__ void __launch_bounds__(kGBlock) k_glmm_dense_cols(GlmmDev g, const double* __restrict__ we, double* __restrict__ partials /* [(1 + K)][G] */)
{
  __shared__ double red[kGBlock / 32];
  const int j = blockIdx.y, G = gridDim.x;
  const double* __restrict__ col = g.X + (long long) j * g.npad;
  double acc = 0.0;
  for (long long i = (long long) blockIdx.x * kGBlock + threadIdx.x; i < g.N; i += (long long) G * kGBlock) acc += __ldg(col + i) * __ldg(we + i);
  acc = g_block_sum(acc, red);
  if (threadIdx.x == 0) partials[(long long) (1 + j) * G + blockIdx.x] = acc;
}

// rows 0 .. K of the partials (S and the dense columns): one warp per row, lanes stride over the blocks
__global__ void __launch_bounds__(kGBlock) k_glmm_finish_dense(int rows, int G, const double* __restrict__ partials, double* __restrict__ result)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int v = warp; v < rows; v += kGBlock / 32) {
    double acc = 0.0;
    for (int b = lane; b < G; b += 32) acc += partials[(long long) v * G + b];
    acc = g_warp_sum(acc);
    if (lane == 0) result[v] = acc;
  }
}

// one warp per column of Z: sum of value * (w e)[observation] over the column's entries
__global__ void __launch_bounds__(kGBlock) k_glmm_z_cols(int q, const long long* __restrict__ col_ptr, const int* __restrict__ obs, const double* __restrict__ val,
                                                        const double* __restrict__ we, double* __restrict__ result_z)
{
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long) blockIdx.x * kGBlock + threadIdx.x) >> 5, nwarps = ((long long) gridDim.x * kGBlock) >> 5;
  for (long long c = warp0; c < q; c += nwarps) {
    const long long lo = col_ptr[c], hi = col_ptr[c + 1];
    double a0 = 0.0, a1 = 0.0;
    long long k = lo + lane;
    for (; k + 32 < hi; k += 64) { a0 += __ldg(val + k) * __ldg(we + __ldg(obs + k)); a1 += __ldg(val + k + 32) * __ldg(we + __ldg(obs + k + 32)); }
    if (k < hi) a0 += __ldg(val + k) * __ldg(we + __ldg(obs + k));
    const double acc = g_warp_sum(a0 + a1);
    if (lane == 0) result_z[c] = acc;
  }
}

__global__ void __launch_bounds__(kGBlock) k_glmm_linear_predictor(GlmmDev g, double* __restrict__ out, int include_fixed, int include_random)
{
  extern __shared__ double smem[];
  const int nb = g.K + g.q;
  const bool staged = (size_t) nb * sizeof(double) <= kThetaSmemMax;      // otherwise the coefficients are read through L1 / L2
  if (staged) { for (int j = threadIdx.x; j < nb; j += kGBlock) smem[j] = g.theta[j]; __syncthreads(); }
  const double* th = staged ? smem : g.theta;
  if (g.K <= kGFast && g.slots <= kGFast) {
    const bool out_aligned = (reinterpret_cast<unsigned long long>(out) & 15ull) == 0ull;
    // small models: two rows per thread and load (16-byte loads; the arrays are padded to a multiple of 16 rows), every load of the
    // pair issued before the first use -- the pass is latency bound otherwise (ncu: long scoreboard)
    for (long long i = ((long long) blockIdx.x * kGBlock + threadIdx.x) * 2; i < g.N; i += (long long) gridDim.x * kGBlock * 2) {
      double2 x2[kGFast], v2[kGFast]; int2 c2[kGFast];
#pragma unroll
      for (int k = 0; k < kGFast; ++k) if (include_fixed && k < g.K) x2[k] = __ldg(reinterpret_cast<const double2*>(g.X + (long long) k * g.npad + i));
#pragma unroll
      for (int s = 0; s < kGFast; ++s) if (include_random && s < g.slots) {
        c2[s] = __ldg(reinterpret_cast<const int2*>(g.zidx + (long long) s * g.npad + i));
        v2[s] = ((g.ones_mask >> s) & 1u) ? make_double2(1.0, 1.0) : __ldg(reinterpret_cast<const double2*>(g.zval + (long long) s * g.npad + i));
      }
      double eta0 = 0.0, eta1 = 0.0;
#pragma unroll
      for (int k = 0; k < kGFast; ++k) if (include_fixed && k < g.K) { eta0 += x2[k].x * th[k]; eta1 += x2[k].y * th[k]; }
#pragma unroll
      for (int s = 0; s < kGFast; ++s) if (include_random && s < g.slots) { eta0 += v2[s].x * th[g.K + c2[s].x]; eta1 += v2[s].y * th[g.K + c2[s].y]; }
      if (i + 1 < g.N && out_aligned) *reinterpret_cast<double2*>(out + i) = make_double2(eta0, eta1);
      else { out[i] = eta0; if (i + 1 < g.N) out[i + 1] = eta1; }
    }
    return;
  }
  for (long long i = (long long) blockIdx.x * kGBlock + threadIdx.x; i < g.N; i += (long long) gridDim.x * kGBlock) {
    double eta = 0.0;
    if (include_fixed) for (int k = 0; k < g.K; ++k) eta += __ldg(g.X + (long long) k * g.npad + i) * th[k];
    if (include_random) for (int s = 0; s < g.slots; ++s) {
      int c = __ldg(g.zidx + (long long) s * g.npad + i);
      double v = ((g.ones_mask >> s) & 1u) ? 1.0 : __ldg(g.zval + (long long) s * g.npad + i);
      eta += v * th[g.K + c];
    }
    out[i] = eta;
  }
}

// out = G d for the dense symmetric Gram matrix G = [X Z]' W [X Z] kept on the device (models whose Z'WZ fills in: crossed grouping
// factors with many levels).  One warp per row, lanes stride over the columns with two accumulators, shuffle tree: a fixed order,
// so the product is reproducible.  nb^2 x 8 bytes per product (18 MB at nb = 1 500: L2 resident).
__global__ void __launch_bounds__(kGBlock) k_gram_matvec(int nb, const double* __restrict__ G, const double* __restrict__ d, double* __restrict__ out)
{
  const int lane = threadIdx.x & 31;
  const int row = (int) (((long long) blockIdx.x * kGBlock + threadIdx.x) >> 5);
  if (row >= nb) return;
  const double* __restrict__ g = G + (size_t) row * nb;
  double a0 = 0.0, a1 = 0.0;
  int c = lane;
  for (; c + 32 < nb; c += 64) { a0 = fma(__ldg(g + c), __ldg(d + c), a0); a1 = fma(__ldg(g + c + 32), __ldg(d + c + 32), a1); }
  if (c < nb) a0 = fma(__ldg(g + c), __ldg(d + c), a0);
  const double acc = g_warp_sum(a0 + a1);
  if (lane == 0) out[row] = acc;
}

__global__ void k_glmm_residual(long long N, const double* __restrict__ y, const double* __restrict__ offset, double* __restrict__ r)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long) gridDim.x * blockDim.x) r[i] = y[i] - offset[i];
}

// new offset and / or response in one pass: copies what changed into the model's own buffers and refreshes r = y - offset
__global__ void k_glmm_set_inputs(long long N, const double* __restrict__ new_offset, const double* __restrict__ new_y, double* __restrict__ offset,
                                  double* __restrict__ y, double* __restrict__ r)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long) gridDim.x * blockDim.x) {
    double o, v;
    if (new_offset != nullptr) { o = new_offset[i]; offset[i] = o; } else o = offset[i];
    if (new_y != nullptr) { v = new_y[i]; y[i] = v; } else v = y[i];
    r[i] = v - o;
  }
}

// ---------------------------------------------------------------------------------------
struct GlmmModel::Params {
  const double *z_beta, *extra_u, *z_b, *z_T, *rho_u, *zeta_u, *tau_u;
  double aux_u = 0, aux_unscaled = 0, aux = 1, disp = 1;
  std::vector<double> rho, zeta, tau, beta, b, theta_L, extra;
  std::vector<std::complex<double>> cin;        // complex-step scratch of the coefficient-prior map
};

static const double kHalfLog2Pi = 0.91893853320467274178;
static const double kLog2 = 0.693147180559945286;

GlmmModel::GlmmModel(const s4b_glmm_data& d, cudaStream_t stream, ShardContext* shard) : stream_(stream), shard_(shard)
{
  if (shard_ != nullptr && !shard_->attached()) throw std::invalid_argument("sharded glmm: attach the peer mailboxes first");
  if (d.prior_dist < 0 || d.prior_dist > 7) throw std::invalid_argument("glmm: prior_dist out of range (0 none, 1 normal, 2 student_t, 3 hs, 4 hs_plus, 5 laplace, 6 lasso, 7 product_normal)");
  if ((d.prior_dist == 3 || d.prior_dist == 4) && d.is_binary)
    throw std::invalid_argument("glmm: the horseshoe priors scale with the error sd aux[1] (continuous.stan:300), which a binary response does not have");
  if ((d.prior_dist == 2 || d.prior_dist == 3 || d.prior_dist == 4 || d.prior_dist == 6) && d.K > 0 && d.prior_df == nullptr) throw std::invalid_argument("glmm: this prior_dist needs prior_df");
  if (d.prior_dist == 7 && d.K > 0 && d.num_normals == nullptr) throw std::invalid_argument("glmm: product_normal needs num_normals");
  if (d.prior_dist_for_aux < 0 || d.prior_dist_for_aux > 3) throw std::invalid_argument("glmm: prior_dist_for_aux out of range");
  for (int i = 0; i < d.t; ++i) {
    if (d.p[i] < 1 || d.l[i] < 1) throw std::invalid_argument("glmm: p[i] / l[i] must be >= 1");
    if (d.p[i] > kMaxNc) throw std::invalid_argument("glmm: ranef blocks with more than 16 coefficients are not supported");
  }
  N_ = d.N; N_total_ = sharded() ? shard_->total_obs() : d.N; K_ = d.K; q_ = d.q; t_ = d.t; len_theta_L_ = d.len_theta_L; len_conc_ = d.len_concentration;
  is_binary_ = d.is_binary; prior_dist_ = d.prior_dist; prior_dist_for_aux_ = d.prior_dist_for_aux;
  prior_scale_for_aux_ = d.prior_scale_for_aux; prior_mean_for_aux_ = d.prior_mean_for_aux; prior_df_for_aux_ = d.prior_df_for_aux;
  prior_scale_.assign(d.prior_scale, d.prior_scale + K_); prior_mean_.assign(d.prior_mean, d.prior_mean + K_);
  p_.assign(d.p, d.p + t_); l_.assign(d.l, d.l + t_);
  shape_.assign(d.shape, d.shape + t_); scale_.assign(d.scale, d.scale + t_);
  regularization_.assign(d.regularization, d.regularization + d.len_regularization);
  // transformed data (src/stan_sampler.cpp:142-182): delta restarts at concentration[0] for every term
  delta_.assign((size_t) len_conc_, 0.0);
  int pos = 0, sum_p = 0, qq = 0;
  for (int i = 0; i < t_; ++i) {
    if (p_[i] > 1) for (int j = 0; j < p_[i]; ++j) delta_[(size_t) pos++] = d.concentration[j];
    sum_p += p_[i]; qq += p_[i] * l_[i];
  }
  if (qq != q_) throw std::invalid_argument("glmm: q != sum(p * l)");
  len_rho_ = sum_p - t_;
  len_z_T_ = 0;
  for (int i = 0; i < t_; ++i) if (p_[i] > 2) len_z_T_ += (p_[i] - 2) * (p_[i] - 1);      // continuous.stan:258
  has_aux_ = is_binary_ ? 0 : 1;
  // coefficient priors beyond normal (continuous.stan:262-270): z_beta may be longer than K (product_normal) and the
  // shrinkage priors add positive parameters global[hs], local[hs][K], caux[hs > 0], mix[K], one_over_lambda
  if (d.prior_df != nullptr) prior_df_.assign(d.prior_df, d.prior_df + K_);
  if (d.num_normals != nullptr) num_normals_.assign(d.num_normals, d.num_normals + K_);
  global_prior_df_ = d.global_prior_df; global_prior_scale_ = d.global_prior_scale; slab_df_ = d.slab_df; slab_scale_ = d.slab_scale;
  hs_ = prior_dist_ == 3 ? 2 : (prior_dist_ == 4 ? 4 : 0);
  len_zbeta_ = K_;
  if (prior_dist_ == 7) { len_zbeta_ = 0; for (int k = 0; k < K_; ++k) { if (num_normals_[(size_t) k] < 2) throw std::invalid_argument("glmm: num_normals must be >= 2"); len_zbeta_ += num_normals_[(size_t) k]; } }
  len_extra_ = hs_ + hs_ * K_ + (hs_ > 0 ? 1 : 0) + ((prior_dist_ == 5 || prior_dist_ == 6) ? K_ : 0) + (prior_dist_ == 6 ? 1 : 0);
  num_params_ = len_zbeta_ + len_extra_ + q_ + len_z_T_ + len_rho_ + len_conc_ + t_ + has_aux_;

  // ---- CSR (w, v, u) -> slot-major ELL ----
  npad_ = (N_ + 15) / 16 * 16;
  int max_nnz = 0;
  for (long long i = 0; i < N_; ++i) max_nnz = std::max(max_nnz, d.u[i + 1] - d.u[i]);
  slots_ = max_nnz;
  if (slots_ > 32) throw std::invalid_argument("glmm: more than 32 non-zeros per row of Z");
  std::vector<int> zidx((size_t) std::max(1, slots_) * npad_, 0);
  std::vector<double> zval((size_t) std::max(1, slots_) * npad_, 0.0);
  ones_mask_ = 0;
  for (int s = 0; s < slots_; ++s) {
    bool all_one = true;
    for (long long i = 0; i < N_; ++i) {
      int k = d.u[i] + s;
      if (k < d.u[i + 1]) {
        if (d.v[k] < 0 || d.v[k] >= q_) throw std::invalid_argument("glmm: column index out of range in v");
        zidx[(size_t) s * npad_ + i] = d.v[k]; zval[(size_t) s * npad_ + i] = d.w[k];
        if (d.w[k] != 1.0) all_one = false;
      } else all_one = false;
    }
    if (all_one) ones_mask_ |= 1u << s;
  }
  // padding of rows with fewer non-zeros: columns the row does not use (value 0), so that the column indices of a row stay pairwise
  // distinct and its bin updates in the data pass are independent of each other
  row_distinct_ = 1;
  {
    std::vector<int> used;
    for (long long i = 0; i < N_; ++i) {
      const int nnz = d.u[i + 1] - d.u[i];
      bool dup = false;
      for (int a = 0; a < nnz && !dup; ++a) for (int b = a + 1; b < nnz; ++b) if (d.v[d.u[i] + a] == d.v[d.u[i] + b]) { dup = true; break; }
      if (dup) row_distinct_ = 0;
      if (nnz == slots_) continue;
      if (q_ < slots_) { row_distinct_ = 0; continue; }
      used.assign(d.v + d.u[i], d.v + d.u[i + 1]);
      int c = 0;
      for (int s = nnz; s < slots_; ++s) {
        while (std::find(used.begin(), used.end(), c) != used.end()) ++c;
        zidx[(size_t) s * npad_ + i] = c; used.push_back(c);
      }
    }
  }
  auto dalloc = [&](double** p, size_t count) { S4B_CUDA(cudaMalloc(p, sizeof(double) * std::max<size_t>(count, 1))); zero_device_sync(*p, sizeof(double) * std::max<size_t>(count, 1), stream_); };
  dalloc(&d_X_, (size_t) std::max(1, K_) * npad_);
  for (int k = 0; k < K_; ++k) S4B_CUDA(cudaMemcpy(d_X_ + (size_t) k * npad_, d.X + (size_t) k * N_, sizeof(double) * (size_t) N_, cudaMemcpyHostToDevice));
  dalloc(&d_y_, (size_t) npad_); dalloc(&d_offset_, (size_t) npad_); dalloc(&d_r_, (size_t) npad_); dalloc(&d_tmp_, (size_t) npad_);
  S4B_CUDA(cudaMemcpy(d_y_, d.y, sizeof(double) * (size_t) N_, cudaMemcpyHostToDevice));
  if (d.weights != nullptr) {
    for (long long i = 0; i < N_; ++i) if (!(d.weights[i] >= 0.0) || !std::isfinite(d.weights[i])) throw std::invalid_argument("glmm: weights must be finite and non-negative");
    dalloc(&d_wt_, (size_t) npad_);
    S4B_CUDA(cudaMemcpy(d_wt_, d.weights, sizeof(double) * (size_t) N_, cudaMemcpyHostToDevice));
  }
  dalloc(&d_zval_, zval.size());
  S4B_CUDA(cudaMemcpy(d_zval_, zval.data(), sizeof(double) * zval.size(), cudaMemcpyHostToDevice));
  S4B_CUDA(cudaMalloc(&d_zidx_, sizeof(int) * zidx.size()));
  S4B_CUDA(cudaMemcpy(d_zidx_, zidx.data(), sizeof(int) * zidx.size(), cudaMemcpyHostToDevice));

  int dev = 0, sms = 0; S4B_CUDA(cudaGetDevice(&dev));
  S4B_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int nb = K_ + q_;
  smem_bytes_ = sizeof(double) * ((size_t) nb + (size_t) (kGBlock / 32) * (nb + 1) * 32);
  int max_smem = 0; S4B_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // more columns than the lane-private shared-memory bins hold (or forced, for the tests): the column path
  columns_ = smem_bytes_ > (size_t) max_smem || (getenv("S4B_GLMM_COLUMNS") != nullptr && atoi(getenv("S4B_GLMM_COLUMNS")) != 0);
  int per_sm = 1;      // resident blocks per SM (registers and shared memory): the grid is one full wave
  if (columns_) {
    // compressed sparse columns of Z, entries of a column in observation order
    std::vector<long long> col_ptr((size_t) q_ + 1, 0);
    for (long long k = 0; k < d.num_non_zero; ++k) ++col_ptr[(size_t) d.v[k] + 1];
    for (int c = 0; c < q_; ++c) col_ptr[(size_t) c + 1] += col_ptr[(size_t) c];
    std::vector<long long> fill(col_ptr.begin(), col_ptr.end() - 1);
    std::vector<int> obs((size_t) std::max<long long>(1, d.num_non_zero)); std::vector<double> val(obs.size(), 0.0);
    for (long long i = 0; i < N_; ++i) for (int k = d.u[i]; k < d.u[i + 1]; ++k) { const long long at = fill[(size_t) d.v[k]]++; obs[(size_t) at] = (int) i; val[(size_t) at] = d.w[k]; }
    S4B_CUDA(cudaMalloc(&d_col_ptr_, sizeof(long long) * col_ptr.size()));
    S4B_CUDA(cudaMemcpy(d_col_ptr_, col_ptr.data(), sizeof(long long) * col_ptr.size(), cudaMemcpyHostToDevice));
    S4B_CUDA(cudaMalloc(&d_col_obs_, sizeof(int) * obs.size()));
    S4B_CUDA(cudaMemcpy(d_col_obs_, obs.data(), sizeof(int) * obs.size(), cudaMemcpyHostToDevice));
    dalloc(&d_col_val_, val.size());
    S4B_CUDA(cudaMemcpy(d_col_val_, val.data(), sizeof(double) * val.size(), cudaMemcpyHostToDevice));
    dalloc(&d_we_, (size_t) npad_);
    per_sm = 4;
  } else {
    S4B_CUDA(s4b_allow_max_dynamic_smem((const void*) data_terms_kernel(K_, slots_)));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, data_terms_kernel(K_, slots_), kGBlock, smem_bytes_) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
  }
  num_sms_ = sms;
  long long want = (N_ + 2 * kGBlock - 1) / (2 * kGBlock);
  grid_ = (int) std::max<long long>(1, std::min<long long>(want, (long long) sms * per_sm));
  // the bulk-copy (TMA) version of the binned pass: small models (K, non-zeros per row <= 4) whose bins leave room for a ring of tiles;
  // from ~64 k rows on (below that the pass is a few microseconds of latency either way); S4B_GLMM_BULK = 0 / 1 overrides
  if (!columns_ && K_ <= kGFast && slots_ <= kGFast) {
    int nval = 0;
    for (int s = 0; s < slots_; ++s) nval += ((ones_mask_ >> s) & 1u) ? 0 : 1;
    const size_t per_obs = 8u * (size_t) (1 + (d.weights != nullptr ? 1 : 0) + K_ + nval) + 4u * (size_t) slots_;
    const size_t head = ((sizeof(BulkBars) + sizeof(double) * ((size_t) nb + (size_t) kGBulkWarps * (nb + 1) * 32) + 127) / 128) * 128;
    for (int tile : { 1024, 512, 256 }) {
      const size_t need = head + 3 * (size_t) tile * per_obs;
      if (need <= (size_t) max_smem) { bulk_tile_ = tile; bulk_stages_ = 3; bulk_tile_bytes_ = (int) ((size_t) tile * per_obs); bulk_smem_ = need; break; }
    }
    const char* ev = getenv("S4B_GLMM_BULK");
    bulk_ = bulk_tile_ > 0 && (ev != nullptr ? atoi(ev) != 0 : N_ >= 65536);
    if (bulk_) {
      if (s4b_allow_max_dynamic_smem((const void*) data_terms_bulk_kernel(K_, slots_)) != cudaSuccess) { cudaGetLastError(); bulk_ = false; }
      const long long ntiles = (N_ + bulk_tile_ - 1) / bulk_tile_;
      bulk_grid_ = (int) std::max<long long>(1, std::min<long long>(std::min<long long>(ntiles, sms), grid_));
    }
  }
  dalloc(&d_theta_, (size_t) nb + 1); dalloc(&d_partials_, (size_t) ((columns_ ? K_ : nb) + 1) * grid_); dalloc(&d_result_, (size_t) nb + 1);
  S4B_CUDA(cudaMalloc(&d_ticket_, sizeof(unsigned int))); zero_device_sync(d_ticket_, sizeof(unsigned int), stream_);
  S4B_CUDA(cudaMallocHost(&h_pinned_, sizeof(double) * 2 * ((size_t) nb + 1)));
  // Gram matrix G = [X Z]' W [X Z] (host, once): the model matrices never change during sampling
  if (nb > 0 && nb <= 512) {
    gram_.assign((size_t) nb * nb, 0.0);
    std::vector<int> cols((size_t) K_ + slots_);
    std::vector<double> vals((size_t) K_ + slots_);
    for (long long i = 0; i < N_; ++i) {
      int m = 0;
      for (int k = 0; k < K_; ++k) { cols[(size_t) m] = k; vals[(size_t) m] = d.X[(size_t) k * N_ + i]; ++m; }
      for (int z = d.u[i]; z < d.u[i + 1]; ++z) { cols[(size_t) m] = K_ + d.v[z]; vals[(size_t) m] = d.w[z]; ++m; }
      const double wi = d.weights != nullptr ? d.weights[i] : 1.0;
      for (int a = 0; a < m; ++a) for (int bb = 0; bb < m; ++bb) gram_[(size_t) cols[(size_t) a] * nb + cols[(size_t) bb]] += wi * vals[(size_t) a] * vals[(size_t) bb];
    }
    if (sharded()) shard_->allreduce_host(gram_.data(), (long long) gram_.size(), kOpSum, stream_);
    theta0_.assign((size_t) nb, 0.0); g0_.assign((size_t) nb, 0.0);
    mode_ = getenv("S4B_GLMM_MODE") ? atoi(getenv("S4B_GLMM_MODE")) : 1;
  } else if (nb > 512 && !sharded()) {
    // many grouping levels: the Gram matrix is kept in pieces -- X'WX and X'WZ dense (K is small), Z'WZ sparse (an
    // observation touches one level per grouping term, so a row of Z'WZ has a handful of entries per term it shares rows
    // with).  Triples (row, column, w z_a z_b) in observation order, sorted by position and merged: a fixed summation order.
    const size_t K = (size_t) K_, q = (size_t) q_;
    gxx_.assign(K * K, 0.0); gxz_.assign(K * q, 0.0);
    struct Trip { unsigned long long key; double val; };
    std::vector<Trip> trips;
    trips.reserve((size_t) d.num_non_zero * (size_t) std::max(1, slots_));
    for (long long i = 0; i < N_; ++i) {
      const double wi = d.weights != nullptr ? d.weights[i] : 1.0;
      for (size_t a = 0; a < K; ++a) {
        const double xa = wi * d.X[a * (size_t) N_ + (size_t) i];
        for (size_t bb = 0; bb < K; ++bb) gxx_[a * K + bb] += xa * d.X[bb * (size_t) N_ + (size_t) i];
        for (int z = d.u[i]; z < d.u[i + 1]; ++z) gxz_[a * q + (size_t) d.v[z]] += xa * d.w[z];
      }
      for (int za = d.u[i]; za < d.u[i + 1]; ++za) for (int zb = d.u[i]; zb < d.u[i + 1]; ++zb)
        trips.push_back({ (unsigned long long) d.v[za] * q + (unsigned long long) d.v[zb], wi * d.w[za] * d.w[zb] });
    }
    std::stable_sort(trips.begin(), trips.end(), [](const Trip& a, const Trip& b) { return a.key < b.key; });
    gz_ptr_.assign(q + 1, 0);
    for (size_t k = 0; k < trips.size();) {
      size_t e = k; double acc = 0.0;
      while (e < trips.size() && trips[e].key == trips[k].key) acc += trips[e++].val;
      gz_col_.push_back((int) (trips[k].key % q)); gz_val_.push_back(acc); ++gz_ptr_[(size_t) (trips[k].key / q) + 1];
      k = e;
    }
    for (size_t c = 0; c < q; ++c) gz_ptr_[c + 1] += gz_ptr_[c];
    sparse_gram_ = true;
    theta0_.assign((size_t) nb, 0.0); g0_.assign((size_t) nb, 0.0);
    // crossed grouping factors with many levels fill Z'WZ in (500 x 500 levels => 10^6 entries): the host-side expansion then costs
    // more per evaluation (~1 ns per entry) than one device pass (tens of microseconds), so the per-evaluation device path is the
    // default there
    // ... unless the Gram matrix fits the device densely (nb <= 4096: 134 MB): then G d is one small kernel (k_gram_matvec, L2 resident)
    // and the sweep-level expansion stays the default
    const bool expansion_pays = gz_val_.size() <= 60000;
    const bool device_gram = !expansion_pays && nb <= 4096 && !(getenv("S4B_GLMM_DEVICE_GRAM") && atoi(getenv("S4B_GLMM_DEVICE_GRAM")) == 0);
    if (device_gram) {
      std::vector<double> G((size_t) nb * nb, 0.0);
      const size_t K = (size_t) K_, q = (size_t) q_;
      for (size_t a = 0; a < K; ++a) {
        for (size_t bb = 0; bb < K; ++bb) G[a * nb + bb] = gxx_[a * K + bb];
        for (size_t c = 0; c < q; ++c) { G[a * nb + K + c] = gxz_[a * q + c]; G[(K + c) * nb + a] = gxz_[a * q + c]; }
      }
      for (size_t r = 0; r < q; ++r) for (long long k = gz_ptr_[r]; k < gz_ptr_[r + 1]; ++k) G[(K + r) * nb + K + (size_t) gz_col_[(size_t) k]] = gz_val_[(size_t) k];
      S4B_CUDA(cudaMalloc(&d_gram_, sizeof(double) * G.size()));
      S4B_CUDA(cudaMemcpy(d_gram_, G.data(), sizeof(double) * G.size(), cudaMemcpyHostToDevice));
      S4B_CUDA(cudaMalloc(&d_dl_, sizeof(double) * 2 * ((size_t) nb + 1)));
      S4B_CUDA(cudaMallocHost(&h_dl_, sizeof(double) * 2 * ((size_t) nb + 1)));
    }
    mode_ = getenv("S4B_GLMM_MODE") ? atoi(getenv("S4B_GLMM_MODE")) : ((expansion_pays || device_gram) ? 1 : 0);
  }
  dl_.assign((size_t) nb + 1, 0.0); Gd_.assign((size_t) nb + 1, 0.0);
  refresh_r();
  S4B_CUDA(cudaStreamSynchronize(stream_));
}

GlmmModel::~GlmmModel()
{
  cudaFree(d_X_); cudaFree(d_y_); cudaFree(d_offset_); cudaFree(d_r_); cudaFree(d_wt_); cudaFree(d_we_); cudaFree(d_col_ptr_); cudaFree(d_col_obs_); cudaFree(d_col_val_); cudaFree(d_tmp_); cudaFree(d_zval_); cudaFree(d_zidx_);
  cudaFree(d_theta_); cudaFree(d_partials_); cudaFree(d_result_); cudaFree(d_ticket_); cudaFreeHost(h_pinned_);
  cudaFree(d_gram_); cudaFree(d_dl_); cudaFreeHost(h_dl_);
  delete scratch_;
}

void GlmmModel::set_mode(int mode)
{
  if (mode != 0 && mode != 1) throw std::invalid_argument("glmm mode must be 0 (pass per evaluation) or 1 (pass per sweep)");
  if (mode == 1 && gram_.empty() && !sparse_gram_) throw std::invalid_argument("glmm: sweep-level expansion unavailable (K + q > 512 on a sharded chain)");
  mode_ = mode; expansion_valid_ = false;
}

void GlmmModel::data_terms_auto(const double* beta, const double* b, double* S, double* gbeta, double* gb)
{
  if (mode_ == 0) { data_terms(beta, b, S, gbeta, gb); return; }
  const int nb = K_ + q_;
  if (!expansion_valid_) {
    // first evaluation after the residual changed: one device pass anchors the expansion at this point
    data_terms(beta, b, &S0_, g0_.data(), g0_.data() + K_);
    for (int k = 0; k < K_; ++k) theta0_[(size_t) k] = beta[k];
    for (int k = 0; k < q_; ++k) theta0_[(size_t) (K_ + k)] = b[k];
    expansion_valid_ = true;
    *S = S0_;
    for (int k = 0; k < K_; ++k) gbeta[k] = g0_[(size_t) k];
    for (int k = 0; k < q_; ++k) gb[k] = g0_[(size_t) (K_ + k)];
    return;
  }
  if (sparse_gram_) { expand_sparse(beta, b, S, gbeta, gb); return; }
  double dl[512], Gd[512];
  for (int k = 0; k < K_; ++k) dl[k] = beta[k] - theta0_[(size_t) k];
  for (int k = 0; k < q_; ++k) dl[K_ + k] = b[k] - theta0_[(size_t) (K_ + k)];
  // G d accumulated column by column (G is symmetric: column c is row c): the inner loop has no loop-carried dependency
  // and vectorises, unlike a row-wise dot product whose additions form one latency-bound chain per row
  for (int a = 0; a < nb; ++a) Gd[a] = 0.0;
  {
    const double* __restrict__ G = gram_.data();
    double* __restrict__ out = Gd;
    for (int c = 0; c < nb; ++c) {
      const double dc = dl[c];
      const double* __restrict__ col = G + (size_t) c * nb;
      for (int a = 0; a < nb; ++a) out[a] += col[a] * dc;
    }
  }
  double quad = 0.0, lin = 0.0;
  for (int a = 0; a < nb; ++a) { quad += dl[a] * Gd[a]; lin += g0_[(size_t) a] * dl[a]; }
  *S = S0_ - 2.0 * lin + quad;
  for (int k = 0; k < K_; ++k) gbeta[k] = g0_[(size_t) k] - Gd[k];
  for (int k = 0; k < q_; ++k) gb[k] = g0_[(size_t) (K_ + k)] - Gd[K_ + k];
}

// the same expansion with the Gram matrix in pieces (K + q > 512): X'WX and X'WZ dense, Z'WZ compressed sparse rows
void GlmmModel::expand_sparse(const double* beta, const double* b, double* S, double* gbeta, double* gb)
{
  const int nb = K_ + q_;
  const size_t K = (size_t) K_, q = (size_t) q_;
  double* dl = dl_.data(); double* Gd = Gd_.data();
  for (int k = 0; k < K_; ++k) dl[k] = beta[k] - theta0_[(size_t) k];
  for (int k = 0; k < q_; ++k) dl[K_ + k] = b[k] - theta0_[(size_t) (K_ + k)];
  if (d_gram_ != nullptr) {
    // G d on the device: 12 KB up, one kernel over the L2-resident matrix, 12 KB down
    ++num_gram_products_;
    std::memcpy(h_dl_, dl, sizeof(double) * (size_t) nb);
    S4B_CUDA(cudaMemcpyAsync(d_dl_, h_dl_, sizeof(double) * (size_t) nb, cudaMemcpyHostToDevice, stream_));
    const int blocks = (int) (((long long) nb * 32 + kGBlock - 1) / kGBlock);
    k_gram_matvec<<<blocks, kGBlock, 0, stream_>>>(nb, d_gram_, d_dl_, d_dl_ + nb + 1);
    S4B_CUDA(cudaMemcpyAsync(h_dl_ + nb + 1, d_dl_ + nb + 1, sizeof(double) * (size_t) nb, cudaMemcpyDeviceToHost, stream_));
    S4B_CUDA(cudaStreamSynchronize(stream_));
    std::memcpy(Gd, h_dl_ + nb + 1, sizeof(double) * (size_t) nb);
  } else {
  for (int a = 0; a < nb; ++a) Gd[a] = 0.0;
  const double* db = dl + K;
  for (size_t a = 0; a < K; ++a) {
    double acc = 0.0;
    for (size_t bb = 0; bb < K; ++bb) acc += gxx_[a * K + bb] * dl[bb];
    const double* __restrict__ row = gxz_.data() + a * q;
    double acc2 = 0.0;
    for (size_t c = 0; c < q; ++c) acc2 += row[c] * db[c];
    Gd[a] = acc + acc2;
    const double da = dl[a];
    double* __restrict__ out = Gd + K;
    for (size_t c = 0; c < q; ++c) out[c] += row[c] * da;           // (X'WZ)' d_beta
  }
  for (size_t r = 0; r < q; ++r) {
    double acc = 0.0;
    for (long long k = gz_ptr_[r]; k < gz_ptr_[r + 1]; ++k) acc += gz_val_[(size_t) k] * db[gz_col_[(size_t) k]];
    Gd[K + r] += acc;
  }
  }
  double quad = 0.0, lin = 0.0;
  for (int a = 0; a < nb; ++a) { quad += dl[a] * Gd[a]; lin += g0_[(size_t) a] * dl[a]; }
  *S = S0_ - 2.0 * lin + quad;
  for (int k = 0; k < K_; ++k) gbeta[k] = g0_[(size_t) k] - Gd[k];
  for (int k = 0; k < q_; ++k) gb[k] = g0_[(size_t) (K_ + k)] - Gd[K_ + k];
}

std::vector<std::string> GlmmModel::param_names() const
{
  std::vector<std::string> out;
  auto add = [&](const char* base, int count) { for (int i = 0; i < count; ++i) out.push_back(std::string(base) + "." + std::to_string(i + 1)); };
  add("z_beta", len_zbeta_);
  add("global", hs_);
  for (int j = 0; j < hs_; ++j) for (int k = 0; k < K_; ++k) out.push_back("local." + std::to_string(j + 1) + "." + std::to_string(k + 1));
  if (hs_ > 0) out.push_back("caux.1");
  if (prior_dist_ == 5 || prior_dist_ == 6) for (int k = 0; k < K_; ++k) out.push_back("mix.1." + std::to_string(k + 1));
  if (prior_dist_ == 6) out.push_back("one_over_lambda.1");
  add("z_b", q_); add("z_T", len_z_T_); add("rho", len_rho_); add("zeta", len_conc_); add("tau", t_);
  if (has_aux_) { out.push_back("aux_unscaled.1"); out.push_back("aux.1"); }
  add("beta", K_); add("b", q_); add("theta_L", len_theta_L_);
  return out;
}

void GlmmModel::refresh_r()
{
  expansion_valid_ = false;
  int grid = elementwise_grid(N_, 256, num_sms_);
  k_glmm_residual<<<grid, 256, 0, stream_>>>(N_, d_y_, d_offset_, d_r_);
  S4B_CUDA(cudaGetLastError());
}

void GlmmModel::set_offset_host(const double* offset) { S4B_CUDA(cudaMemcpyAsync(d_offset_, offset, sizeof(double) * (size_t) N_, cudaMemcpyHostToDevice, stream_)); refresh_r(); S4B_CUDA(cudaStreamSynchronize(stream_)); }
void GlmmModel::set_response_host(const double* y) { S4B_CUDA(cudaMemcpyAsync(d_y_, y, sizeof(double) * (size_t) N_, cudaMemcpyHostToDevice, stream_)); refresh_r(); S4B_CUDA(cudaStreamSynchronize(stream_)); }
void GlmmModel::set_inputs_device(const double* d_offset, const double* d_y)
{
  expansion_valid_ = false;
  int grid = elementwise_grid(N_, 256, num_sms_);
  k_glmm_set_inputs<<<grid, 256, 0, stream_>>>(N_, d_offset, d_y, d_offset_, d_y_, d_r_);
  S4B_CUDA(cudaGetLastError());
}
void GlmmModel::set_offset_device(const double* d_offset) { set_inputs_device(d_offset, nullptr); }
void GlmmModel::set_response_device(const double* d_y) { set_inputs_device(nullptr, d_y); }

// device time of the data pass alone (the kernels data_terms() launches, coefficients as last uploaded): CUDA events around `reps`
// launches; flush_l2 != 0 writes a 256 MB scratch buffer before every launch (not timed) so that each pass starts from HBM
// read pass over a buffer (timing only): after the write that flushes the L2, it replaces the dirty lines of that write by clean ones,
// so that their write-back is not charged to the launch that is timed next
__global__ void k_glmm_read_scrub(const uint4* __restrict__ src, size_t count, unsigned int* __restrict__ sink)
{
  unsigned int acc = 0u;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) { const uint4 v = src[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x9E3779B9u) *sink = acc;                 // (keeps the loads alive; practically never taken)
}

double GlmmModel::time_data_pass(int reps, int flush_l2)
{
  const int nb = K_ + q_;
  std::vector<double> th((size_t) nb + 1, 0.01);
  S4B_CUDA(cudaMemcpyAsync(d_theta_, th.data(), sizeof(double) * (size_t) nb, cudaMemcpyHostToDevice, stream_));
  void* scratch = nullptr;
  const size_t flush_bytes = (size_t) 256 << 20;
  if (flush_l2) { S4B_CUDA(cudaMalloc(&scratch, 2 * flush_bytes)); S4B_CUDA(cudaMemsetAsync(scratch, 1, 2 * flush_bytes, stream_)); }
  cudaEvent_t a, b; S4B_CUDA(cudaEventCreate(&a)); S4B_CUDA(cudaEventCreate(&b));
  launch_data_pass();
  S4B_CUDA(cudaStreamSynchronize(stream_));
  double total = 0.0;
  if (!flush_l2) {
    S4B_CUDA(cudaEventRecord(a, stream_));
    for (int r = 0; r < reps; ++r) launch_data_pass();
    S4B_CUDA(cudaEventRecord(b, stream_));
    S4B_CUDA(cudaEventSynchronize(b));
    float t = 0.f; S4B_CUDA(cudaEventElapsedTime(&t, a, b)); total = t;
  } else {
    for (int r = 0; r < reps; ++r) {
      // flush: write a buffer larger than the L2, then read another one, so that the L2 holds clean lines of neither the operands nor the write
      S4B_CUDA(cudaMemsetAsync(scratch, r & 0xFF, flush_bytes, stream_));
      k_glmm_read_scrub<<<num_sms_ * 8, 256, 0, stream_>>>(reinterpret_cast<const uint4*>(static_cast<unsigned char*>(scratch) + flush_bytes), flush_bytes / sizeof(uint4), reinterpret_cast<unsigned int*>(scratch));
      S4B_CUDA(cudaEventRecord(a, stream_));
      launch_data_pass();
      S4B_CUDA(cudaEventRecord(b, stream_));
      S4B_CUDA(cudaEventSynchronize(b));
      float t = 0.f; S4B_CUDA(cudaEventElapsedTime(&t, a, b)); total += t;
    }
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(scratch);
  expansion_valid_ = false;
  return total / reps;
}

void GlmmModel::launch_data_pass()
{
  const int nb = K_ + q_;
  GlmmDev g;
  g.N = N_; g.npad = npad_; g.K = K_; g.q = q_; g.slots = slots_; g.bins = nb; g.X = d_X_; g.r = d_r_; g.wt = d_wt_; g.zidx = d_zidx_; g.zval = d_zval_;
  g.ones_mask = ones_mask_; g.row_distinct = row_distinct_; g.theta = d_theta_; g.partials = d_partials_; g.result = d_result_; g.ticket = d_ticket_;
  if (!columns_ && bulk_) {
    int tile = bulk_tile_, stages = bulk_stages_, tile_bytes = bulk_tile_bytes_;
    void* args[] = { &g, &tile, &stages, &tile_bytes };
    S4B_CUDA(cudaLaunchKernel((const void*) data_terms_bulk_kernel(K_, slots_), dim3(bulk_grid_), dim3(kGBulk + 32), args, bulk_smem_, stream_));
  } else if (!columns_) {
    void* args[] = { &g };
    S4B_CUDA(cudaLaunchKernel((const void*) data_terms_kernel(K_, slots_), dim3(grid_), dim3(kGBlock), args, smem_bytes_, stream_));
  }
  else {
    k_glmm_resid_we<<<grid_, kGBlock, 0, stream_>>>(g, d_we_, d_partials_);
    if (K_ > 0) k_glmm_dense_cols<<<dim3((unsigned) grid_, (unsigned) K_), kGBlock, 0, stream_>>>(g, d_we_, d_partials_);
    k_glmm_finish_dense<<<1, kGBlock, 0, stream_>>>(K_ + 1, grid_, d_partials_, d_result_);
    if (q_ > 0) {
      const int zgrid = elementwise_grid((long long) q_ * 32, kGBlock, num_sms_);
      k_glmm_z_cols<<<zgrid, kGBlock, 0, stream_>>>(q_, d_col_ptr_, d_col_obs_, d_col_val_, d_we_, d_result_ + 1 + K_);
    }
  }
  S4B_CUDA(cudaGetLastError());
}

void GlmmModel::data_terms(const double* beta, const double* b, double* S, double* gbeta, double* gb)
{
  ++num_passes_;
  const int nb = K_ + q_;
  double* h_theta = h_pinned_;
  double* h_res = h_pinned_ + nb + 1;
  for (int k = 0; k < K_; ++k) h_theta[k] = beta[k];
  for (int k = 0; k < q_; ++k) h_theta[K_ + k] = b[k];
  if (nb) S4B_CUDA(cudaMemcpyAsync(d_theta_, h_theta, sizeof(double) * (size_t) nb, cudaMemcpyHostToDevice, stream_));
  launch_data_pass();
  if (sharded()) for (int off = 0; off < nb + 1; off += kMailVec) shard_->allreduce(d_result_ + off, std::min(kMailVec, nb + 1 - off), kOpSum, stream_);
  S4B_CUDA(cudaMemcpyAsync(h_res, d_result_, sizeof(double) * (size_t) (nb + 1), cudaMemcpyDeviceToHost, stream_));
  S4B_CUDA(cudaStreamSynchronize(stream_));
  *S = h_res[0];
  for (int k = 0; k < K_; ++k) gbeta[k] = h_res[1 + k];
  for (int k = 0; k < q_; ++k) gb[k] = h_res[1 + K_ + k];
}

void GlmmModel::parametric_mean_device(const double* beta, const double* b, double* d_out, bool include_fixed, bool include_random)
{
  const int nb = K_ + q_;
  double* h_theta = h_pinned_;
  for (int k = 0; k < K_; ++k) h_theta[k] = beta[k];
  for (int k = 0; k < q_; ++k) h_theta[K_ + k] = b[k];
  if (nb) S4B_CUDA(cudaMemcpyAsync(d_theta_, h_theta, sizeof(double) * (size_t) nb, cudaMemcpyHostToDevice, stream_));
  GlmmDev g;
  g.N = N_; g.npad = npad_; g.K = K_; g.q = q_; g.slots = slots_; g.bins = nb; g.X = d_X_; g.r = d_r_; g.wt = d_wt_; g.zidx = d_zidx_; g.zval = d_zval_;
  g.ones_mask = ones_mask_; g.row_distinct = row_distinct_; g.theta = d_theta_; g.partials = d_partials_; g.result = d_result_; g.ticket = d_ticket_;
  int grid = elementwise_grid(N_, kGBlock, num_sms_);
  k_glmm_linear_predictor<<<grid, kGBlock, (size_t) nb * sizeof(double) <= kThetaSmemMax ? sizeof(double) * (size_t) std::max(1, nb) : 8, stream_>>>(g, d_out, include_fixed ? 1 : 0, include_random ? 1 : 0);
  S4B_CUDA(cudaGetLastError());
  // h_pinned_ is reused by the next call: make sure the H2D copy has been consumed
  S4B_CUDA(cudaStreamSynchronize(stream_));
}

void GlmmModel::parametric_mean_host(const double* constrained, double* out, bool include_fixed, bool include_random)
{
  const double* beta = constrained + num_params_ + has_aux_;
  const double* b = beta + K_;
  parametric_mean_device(beta, b, d_tmp_, include_fixed, include_random);
  S4B_CUDA(cudaMemcpyAsync(out, d_tmp_, sizeof(double) * (size_t) N_, cudaMemcpyDeviceToHost, stream_));
  S4B_CUDA(cudaStreamSynchronize(stream_));
}

// Lower-triangular factor T (row major, nc x nc) of one ranef block with nc >= 2 coefficients (continuous.stan:20-50, the
// scaled onion rows for nc > 2; the off-diagonal entries of row r + 1 are scaled with the standard deviation of row r, as
// the Stan program does).  Complex arguments: the imaginary part carries the complex-step derivative used for the adjoint.
using cplx = std::complex<double>;
static void block_T(int nc, cplx c, const cplx* zeta, const cplx* rho, const cplx* zT, cplx* T)
{
  const cplx trace = c * c * (double) nc;
  cplx zs = 0.0;
  for (int k = 0; k < nc; ++k) zs += zeta[k];
  for (int k = 0; k < nc * nc; ++k) T[k] = 0.0;
  cplx sd = std::sqrt(zeta[0] / zs * trace);
  T[0] = sd;
  sd = std::sqrt(zeta[1] / zs * trace);
  const cplx r21 = 2.0 * rho[0] - 1.0;
  T[nc + 1] = sd * std::sqrt(1.0 - r21 * r21);
  T[nc] = sd * r21;
  int zmark = 0;
  for (int r = 2; r < nc; ++r) {
    cplx dot = 0.0;
    for (int k = 0; k < r; ++k) dot += zT[zmark + k] * zT[zmark + k];
    const cplx sf = std::sqrt(rho[r - 1] / dot) * sd;
    sd = std::sqrt(zeta[r] / zs * trace);
    for (int k = 0; k < r; ++k) T[r * nc + k] = zT[zmark + k] * sf;
    T[r * nc + r] = std::sqrt(1.0 - rho[r - 1]) * sd;
    zmark += r;
  }
}

// beta as a function of z_beta, the constrained extra parameters and aux: continuous.stan:293-322 with hs_prior / hsplus_prior
// (:123-143) and CFt (:146-158).  Complex arguments carry the complex-step derivative of the host chain rule.
void GlmmModel::coef_beta(const cplx* z, const cplx* ex, cplx aux, cplx* beta) const
{
  const int K = K_, hs = hs_;
  switch (prior_dist_) {
    case 0: for (int k = 0; k < K; ++k) beta[k] = z[k]; break;
    case 1: for (int k = 0; k < K; ++k) beta[k] = z[k] * prior_scale_[(size_t) k] + prior_mean_[(size_t) k]; break;
    case 2:
      for (int k = 0; k < K; ++k) {
        const cplx z1 = z[k], z2 = z1 * z1, z3 = z2 * z1, z5 = z2 * z3, z7 = z2 * z5, z9 = z2 * z7;
        const double df = prior_df_[(size_t) k], df2 = df * df, df3 = df2 * df, df4 = df2 * df2;
        const cplx cft = z1 + (z3 + z1) / (4.0 * df) + (5.0 * z5 + 16.0 * z3 + 3.0 * z1) / (96.0 * df2)
                         + (3.0 * z7 + 19.0 * z5 + 17.0 * z3 - 15.0 * z1) / (384.0 * df3)
                         + (79.0 * z9 + 776.0 * z7 + 1482.0 * z5 - 1920.0 * z3 - 945.0 * z1) / (92160.0 * df4);
        beta[k] = cft * prior_scale_[(size_t) k] + prior_mean_[(size_t) k];
      }
      break;
    case 3: case 4: {
      const cplx* global = ex;
      const cplx* local = ex + hs;                 // local[j][k] = local[j * K + k]
      const cplx c2 = slab_scale_ * slab_scale_ * ex[hs + hs * K];
      const cplx tau = global[0] * std::sqrt(global[1]) * global_prior_scale_ * aux;
      for (int k = 0; k < K; ++k) {
        cplx lambda = local[k] * std::sqrt(local[K + k]);
        if (hs == 4) lambda = lambda * (local[2 * K + k] * std::sqrt(local[3 * K + k]));
        const cplx l2 = lambda * lambda;
        beta[k] = z[k] * std::sqrt(c2 * l2 / (c2 + tau * tau * l2)) * tau;
      }
      break;
    }
    case 5: for (int k = 0; k < K; ++k) beta[k] = prior_mean_[(size_t) k] + prior_scale_[(size_t) k] * std::sqrt(2.0 * ex[k]) * z[k]; break;
    case 6: for (int k = 0; k < K; ++k) beta[k] = prior_mean_[(size_t) k] + ex[K] * prior_scale_[(size_t) k] * std::sqrt(2.0 * ex[k]) * z[k]; break;
    default: {
      int zp = 0;
      for (int k = 0; k < K; ++k) {
        cplx v = z[zp++];
        for (int n = 2; n <= num_normals_[(size_t) k]; ++n) v = v * z[zp++];
        beta[k] = v * std::pow(prior_scale_[(size_t) k], (double) num_normals_[(size_t) k]) + prior_mean_[(size_t) k];
      }
    }
  }
}

static inline double inv_logit(double x) { return x >= 0.0 ? 1.0 / (1.0 + std::exp(-x)) : std::exp(x) / (1.0 + std::exp(x)); }

void GlmmModel::transform(const double* q, Params& P) const
{
  int pos = 0;
  P.z_beta = q + pos; pos += len_zbeta_;
  P.extra_u = q + pos; pos += len_extra_;
  P.extra.resize((size_t) len_extra_);
  for (int i = 0; i < len_extra_; ++i) P.extra[(size_t) i] = std::exp(P.extra_u[i]);
  P.z_b = q + pos; pos += q_;
  P.z_T = q + pos; pos += len_z_T_;
  P.rho_u = q + pos; pos += len_rho_;
  P.zeta_u = q + pos; pos += len_conc_;
  P.tau_u = q + pos; pos += t_;
  P.aux_u = has_aux_ ? q[pos] : 0.0;
  P.rho.resize((size_t) len_rho_); P.zeta.resize((size_t) len_conc_); P.tau.resize((size_t) t_);
  P.beta.resize((size_t) K_); P.b.resize((size_t) q_); P.theta_L.resize((size_t) len_theta_L_);
  for (int i = 0; i < len_rho_; ++i) P.rho[(size_t) i] = inv_logit(P.rho_u[i]);
  for (int i = 0; i < len_conc_; ++i) P.zeta[(size_t) i] = std::exp(P.zeta_u[i]);
  for (int i = 0; i < t_; ++i) P.tau[(size_t) i] = std::exp(P.tau_u[i]);
  if (has_aux_) {
    P.aux_unscaled = std::exp(P.aux_u);
    if (prior_dist_for_aux_ == 0) P.aux = P.aux_unscaled;
    else { P.aux = prior_scale_for_aux_ * P.aux_unscaled; if (prior_dist_for_aux_ <= 2) P.aux += prior_mean_for_aux_; }
    P.disp = P.aux;
  } else { P.aux_unscaled = 0.0; P.aux = 1.0; P.disp = 1.0; }
  if (prior_dist_ <= 1) {
    for (int k = 0; k < K_; ++k) P.beta[(size_t) k] = prior_dist_ == 0 ? P.z_beta[k] : P.z_beta[k] * prior_scale_[(size_t) k] + prior_mean_[(size_t) k];
  } else {
    const int nin = len_zbeta_ + len_extra_;
    P.cin.resize((size_t) (nin + K_ + 1));
    for (int i = 0; i < len_zbeta_; ++i) P.cin[(size_t) i] = P.z_beta[i];
    for (int i = 0; i < len_extra_; ++i) P.cin[(size_t) (len_zbeta_ + i)] = P.extra[(size_t) i];
    coef_beta(P.cin.data(), P.cin.data() + len_zbeta_, P.aux, P.cin.data() + nin);
    for (int k = 0; k < K_; ++k) P.beta[(size_t) k] = P.cin[(size_t) (nin + k)].real();
  }
  int zeta_mark = 0, rho_mark = 0, th = 0, b_mark = 0, zT_mark = 0;
  for (int i = 0; i < t_; ++i) {
    const double c = P.tau[(size_t) i] * scale_[(size_t) i] * P.disp;
    if (p_[(size_t) i] == 1) {
      P.theta_L[(size_t) th++] = c;
      for (int s = 0; s < l_[(size_t) i]; ++s) P.b[(size_t) (b_mark + s)] = c * P.z_b[b_mark + s];
      b_mark += l_[(size_t) i];
    } else if (p_[(size_t) i] > 2) {
      const int nc = p_[(size_t) i];
      const int nzT = (nc - 2) * (nc + 1) / 2;          // 2 + 3 + ... + (nc - 1) elements of z_T, one running mark
      cplx T[kMaxNc * kMaxNc], zc[kMaxNc], rc[kMaxNc], tc[kMaxNc * kMaxNc];
      for (int k = 0; k < nc; ++k) zc[k] = P.zeta[(size_t) (zeta_mark + k)];
      for (int k = 0; k < nc - 1; ++k) rc[k] = P.rho[(size_t) (rho_mark + k)];
      for (int k = 0; k < nzT; ++k) tc[k] = P.z_T[zT_mark + k];
      block_T(nc, c, zc, rc, tc, T);
      for (int c2 = 0; c2 < nc; ++c2) for (int r = c2; r < nc; ++r) P.theta_L[(size_t) th++] = T[r * nc + c2].real();     // vech
      for (int j = 0; j < l_[(size_t) i]; ++j) {
        for (int r = 0; r < nc; ++r) { double acc = 0.0; for (int k = 0; k <= r; ++k) acc += T[r * nc + k].real() * P.z_b[b_mark + k]; P.b[(size_t) (b_mark + r)] = acc; }
        b_mark += nc;
      }
      zeta_mark += nc; rho_mark += nc - 1; zT_mark += nzT;
    } else {
      const double trace = c * c * 2.0;
      const double zs = P.zeta[(size_t) zeta_mark] + P.zeta[(size_t) zeta_mark + 1];
      const double pi1 = P.zeta[(size_t) zeta_mark] / zs, pi2 = P.zeta[(size_t) zeta_mark + 1] / zs;
      zeta_mark += 2;
      const double sd1 = std::sqrt(pi1 * trace), sd2 = std::sqrt(pi2 * trace);
      const double r = 2.0 * P.rho[(size_t) rho_mark++] - 1.0;
      const double T11 = sd1, T21 = sd2 * r, T22 = sd2 * std::sqrt(1.0 - r * r);
      P.theta_L[(size_t) th++] = T11; P.theta_L[(size_t) th++] = T21; P.theta_L[(size_t) th++] = T22;
      for (int j = 0; j < l_[(size_t) i]; ++j) {
        const double z0 = P.z_b[b_mark], z1 = P.z_b[b_mark + 1];
        P.b[(size_t) b_mark] = T11 * z0; P.b[(size_t) b_mark + 1] = T21 * z0 + T22 * z1;
        b_mark += 2;
      }
    }
  }
}

int GlmmModel::log_prob_grad(const double* q, double* lp_out, double* grad)
{
  if (scratch_ == nullptr) {
    // log-gamma values of the hyper-parameters never change: look them up instead of ~6 lgamma calls per evaluation
    lg_delta_.resize(delta_.size()); for (size_t i = 0; i < delta_.size(); ++i) lg_delta_[i] = std::lgamma(delta_[i]);
    lg_shape_.resize(shape_.size()); for (size_t i = 0; i < shape_.size(); ++i) lg_shape_[i] = std::lgamma(shape_[i]);
    lg_reg_.clear(); lg_reg2_.clear();
    { int pr = 0; for (int i = 0; i < t_; ++i) if (p_[(size_t) i] > 1) { const double nu = regularization_[(size_t) pr++] + 0.5 * (p_[(size_t) i] - 2); lg_reg_.push_back(std::lgamma(nu)); lg_reg2_.push_back(std::lgamma(2.0 * nu)); } }
  }
  if (scratch_ == nullptr) { scratch_ = new Params; gbeta_.assign((size_t) K_ + 1, 0.0); gb_.assign((size_t) q_ + 1, 0.0); }
  Params& P = *scratch_; transform(q, P);
  ++num_grad_;
  const double N = (double) N_total_;
  double lp = 0.0;
  for (int i = 0; i < len_rho_; ++i) { double x = std::fabs(P.rho_u[i]); lp += -x - 2.0 * std::log1p(std::exp(-x)); }
  for (int i = 0; i < len_conc_; ++i) lp += P.zeta_u[i];
  for (int i = 0; i < t_; ++i) lp += P.tau_u[i];
  if (has_aux_) lp += P.aux_u;

  std::vector<double>& gbeta = gbeta_; std::vector<double>& gb = gb_;
  double S = 0.0;
  data_terms_auto(P.beta.data(), P.b.data(), &S, gbeta.data(), gb.data());
  const double sigma = has_aux_ ? P.aux : 1.0;
  lp += -0.5 * S / (sigma * sigma) - N * std::log(sigma) - N * kHalfLog2Pi;

  double d_au_prior = 0.0;
  if (has_aux_ && prior_dist_for_aux_ > 0 && prior_scale_for_aux_ > 0.0) {
    const double au = P.aux_unscaled;
    if (prior_dist_for_aux_ == 1) { lp += -0.5 * au * au - kHalfLog2Pi + kLog2; d_au_prior = -au; }
    else if (prior_dist_for_aux_ == 2) {
      const double nu = prior_df_for_aux_;
      lp += std::lgamma(0.5 * (nu + 1.0)) - std::lgamma(0.5 * nu) - 0.5 * std::log(nu * 3.14159265358979323846) - 0.5 * (nu + 1.0) * std::log1p(au * au / nu) + kLog2;
      d_au_prior = -(nu + 1.0) * au / (nu + au * au);
    } else { lp += -au; d_au_prior = -1.0; }
  }
  if (prior_dist_ >= 1) { for (int k = 0; k < len_zbeta_; ++k) lp += -0.5 * P.z_beta[k] * P.z_beta[k]; lp -= len_zbeta_ * kHalfLog2Pi; }
  // the extra parameters of the shrinkage priors: lb_constrain Jacobian + their priors (continuous.stan:381-408);
  // d_extra[i] = d prior / d (constrained value)
  if (len_extra_ > 0) d_extra_.assign((size_t) len_extra_ + 1, 0.0);
  for (int i = 0; i < len_extra_; ++i) lp += P.extra_u[i];
  if (hs_ > 0) {
    auto inv_gamma = [&](double x, double a, double& dx) { dx = -(a + 1.0) / x + a / (x * x); return a * std::log(a) - std::lgamma(a) - (a + 1.0) * std::log(x) - a / x; };
    const double* local = P.extra.data() + hs_;
    double* d_local = d_extra_.data() + hs_;
    for (int j = 0; j < hs_; ++j) for (int k = 0; k < K_; ++k) {
      const double x = local[j * K_ + k];
      if ((j & 1) == 0) { lp += -0.5 * x * x - kHalfLog2Pi; d_local[j * K_ + k] = -x; }      // normal_lpdf(local[j] | 0, 1); - log_half once per statement
      else lp += inv_gamma(x, 0.5 * (j == 1 ? prior_df_[(size_t) k] : prior_scale_[(size_t) k]), d_local[j * K_ + k]);   // hs_plus: prior_scale as a second df
    }
    lp += kLog2 * (hs_ / 2);
    { const double x = P.extra[0]; lp += -0.5 * x * x - kHalfLog2Pi + kLog2; d_extra_[0] = -x; }
    lp += inv_gamma(P.extra[1], 0.5 * global_prior_df_, d_extra_[1]);
    lp += inv_gamma(P.extra[(size_t) (hs_ + hs_ * K_)], 0.5 * slab_df_, d_extra_[(size_t) (hs_ + hs_ * K_)]);
  } else if (prior_dist_ == 5 || prior_dist_ == 6) {
    for (int k = 0; k < K_; ++k) { lp += -P.extra[(size_t) k]; d_extra_[(size_t) k] = -1.0; }      // exponential_lpdf(mix | 1)
    if (prior_dist_ == 6) {                                                                        // chi_square_lpdf(one_over_lambda | prior_df[1])
      const double nu = prior_df_[0], x = P.extra[(size_t) K_];
      lp += -(0.5 * nu) * kLog2 - std::lgamma(0.5 * nu) + (0.5 * nu - 1.0) * std::log(x) - 0.5 * x;
      d_extra_[(size_t) K_] = (0.5 * nu - 1.0) / x - 0.5;
    }
  }
  for (int k = 0; k < q_; ++k) lp += -0.5 * P.z_b[k] * P.z_b[k];
  lp -= q_ * kHalfLog2Pi;
  for (int k = 0; k < len_z_T_; ++k) lp += -0.5 * P.z_T[k] * P.z_T[k];
  lp -= len_z_T_ * kHalfLog2Pi;
  {
    int pos_reg = 0, pos_rho = 0;
    for (int i = 0; i < t_; ++i) if (p_[(size_t) i] > 1) {
      const double nu = regularization_[(size_t) pos_reg++] + 0.5 * (p_[(size_t) i] - 2);
      const double r = P.rho[(size_t) pos_rho++];
      {   // continuous.stan:111-115: the further correlations of a block with more than two coefficients
        double nuj = nu;
        for (int j = 2; j < p_[(size_t) i]; ++j) {
          nuj -= 0.5;
          const double s1 = 0.5 * j, s2 = nuj, rj = P.rho[(size_t) pos_rho++];
          lp += (s1 - 1.0) * std::log(rj) + (s2 - 1.0) * std::log1p(-rj) + std::lgamma(s1 + s2) - std::lgamma(s1) - std::lgamma(s2);
        }
      }
      // a zero coefficient contributes exactly 0 (Stan's multiply_log(0, .) convention): the default decov(1, 1, 1, 1) then
      // needs none of these logarithms
      const double kl = nu - 1.0;
      lp += (kl != 0.0 ? kl * std::log(r) + kl * std::log1p(-r) : 0.0) + lg_reg2_[(size_t) (pos_reg - 1)] - 2.0 * lg_reg_[(size_t) (pos_reg - 1)];
    }
  }
  for (int i = 0; i < len_conc_; ++i) { const double kl = delta_[(size_t) i] - 1.0; lp += (kl != 0.0 ? kl * std::log(P.zeta[(size_t) i]) : 0.0) - P.zeta[(size_t) i] - lg_delta_[(size_t) i]; }
  for (int i = 0; i < t_; ++i) { const double kl = shape_[(size_t) i] - 1.0; lp += (kl != 0.0 ? kl * std::log(P.tau[(size_t) i]) : 0.0) - P.tau[(size_t) i] - lg_shape_[(size_t) i]; }

  // ---- adjoints ----
  int pos = 0;
  double* g_zbeta = grad + pos; pos += len_zbeta_;
  double* g_extra = grad + pos; pos += len_extra_;
  double* g_zb = grad + pos; pos += q_;
  double* g_zT = grad + pos; pos += len_z_T_;
  double* g_rho = grad + pos; pos += len_rho_;
  double* g_zeta = grad + pos; pos += len_conc_;
  double* g_tau = grad + pos; pos += t_;
  const double inv_s2 = 1.0 / (sigma * sigma);
  double adj_disp = 0.0;
  if (prior_dist_ <= 1) {
    for (int k = 0; k < K_; ++k) {
      const double dbeta = gbeta[(size_t) k] * inv_s2;
      g_zbeta[k] = prior_dist_ == 0 ? dbeta : dbeta * prior_scale_[(size_t) k] - P.z_beta[k];
    }
  } else {
    // d beta / d (z_beta, extras, aux) by the complex-step method: the map is a handful of elementary functions of K numbers
    const double h = 1e-20;
    const int nin = len_zbeta_ + len_extra_;
    cplx* in = P.cin.data();
    for (int ip = 0; ip < nin + (hs_ > 0 ? 1 : 0); ++ip) {
      for (int i = 0; i < len_zbeta_; ++i) in[i] = P.z_beta[i];
      for (int i = 0; i < len_extra_; ++i) in[len_zbeta_ + i] = P.extra[(size_t) i];
      cplx aux = P.aux;
      if (ip < nin) in[ip] += cplx(0.0, h); else aux += cplx(0.0, h);
      coef_beta(in, in + len_zbeta_, aux, in + nin);
      double acc = 0.0;
      for (int k = 0; k < K_; ++k) acc += gbeta[(size_t) k] * inv_s2 * (in[nin + k].imag() / h);
      if (ip < len_zbeta_) g_zbeta[ip] = acc - P.z_beta[ip];
      else if (ip < nin) { const int e = ip - len_zbeta_; g_extra[e] = (acc + d_extra_[(size_t) e]) * P.extra[(size_t) e] + 1.0; }
      else adj_disp += acc;                           // hs_prior's error_scale is aux
    }
  }
  // the Stan program declares (p - 2)(p - 1) elements of z_T per block but its onion rows consume 2 + ... + (p - 1) of them
  // through one running mark: the surplus elements only see their normal prior
  for (int k = 0; k < len_z_T_; ++k) g_zT[k] = -P.z_T[k];
  int zeta_mark = 0, rho_mark = 0, th = 0, b_mark = 0, pos_reg = 0, zT_mark = 0;
  for (int i = 0; i < t_; ++i) {
    const double tau = P.tau[(size_t) i], sc = scale_[(size_t) i];
    double adj_c = 0.0;     // adjoint of c = tau * scale * dispersion
    if (p_[(size_t) i] > 2) {
      const int nc = p_[(size_t) i];
      const int nzT = (nc - 2) * (nc + 1) / 2;
      double Tr[kMaxNc * kMaxNc], A[kMaxNc * kMaxNc];
      for (int k = 0; k < nc * nc; ++k) { Tr[k] = 0.0; A[k] = 0.0; }
      for (int c2 = 0; c2 < nc; ++c2) for (int r = c2; r < nc; ++r) Tr[r * nc + c2] = P.theta_L[(size_t) th++];
      for (int j = 0; j < l_[(size_t) i]; ++j) {
        for (int k = 0; k < nc; ++k) {              // g_zb = T' db - z ;  A[r][k] += db[r] z[k]
          double acc = 0.0;
          for (int r = k; r < nc; ++r) acc += Tr[r * nc + k] * (gb[(size_t) (b_mark + r)] * inv_s2);
          g_zb[b_mark + k] = acc - P.z_b[b_mark + k];
        }
        for (int r = 0; r < nc; ++r) for (int k = 0; k <= r; ++k) A[r * nc + k] += (gb[(size_t) (b_mark + r)] * inv_s2) * P.z_b[b_mark + k];
        b_mark += nc;
      }
      // dL / d(c, zeta, rho, z_T) = sum A .* dT/dparam, with dT/dparam by the complex-step method (exact to rounding)
      const double h = 1e-20;
      cplx Tc[kMaxNc * kMaxNc], zc[kMaxNc], rc[kMaxNc], tc[kMaxNc * kMaxNc];
      const double cval = tau * sc * P.disp;
      const int npar = 1 + nc + (nc - 1) + nzT;
      double dpar[1 + kMaxNc + kMaxNc + kMaxNc * kMaxNc];
      for (int ip = 0; ip < npar; ++ip) {
        cplx cc = cval;
        for (int k = 0; k < nc; ++k) zc[k] = P.zeta[(size_t) (zeta_mark + k)];
        for (int k = 0; k < nc - 1; ++k) rc[k] = P.rho[(size_t) (rho_mark + k)];
        for (int k = 0; k < nzT; ++k) tc[k] = P.z_T[zT_mark + k];
        if (ip == 0) cc += cplx(0.0, h);
        else if (ip < 1 + nc) zc[ip - 1] += cplx(0.0, h);
        else if (ip < 1 + nc + nc - 1) rc[ip - 1 - nc] += cplx(0.0, h);
        else tc[ip - 1 - nc - (nc - 1)] += cplx(0.0, h);
        block_T(nc, cc, zc, rc, tc, Tc);
        double acc = 0.0;
        for (int r = 0; r < nc; ++r) for (int k = 0; k <= r; ++k) acc += A[r * nc + k] * (Tc[r * nc + k].imag() / h);
        dpar[ip] = acc;
      }
      adj_c = dpar[0];
      for (int k = 0; k < nc; ++k) {
        const double zv = P.zeta[(size_t) (zeta_mark + k)];
        const double d_z = dpar[1 + k] + (delta_[(size_t) (zeta_mark + k)] - 1.0) / zv - 1.0;
        g_zeta[zeta_mark + k] = d_z * zv + 1.0;
      }
      {
        double nu = regularization_[(size_t) pos_reg++] + 0.5 * (nc - 2);
        for (int k = 0; k < nc - 1; ++k) {
          double s1, s2;
          if (k == 0) { s1 = nu; s2 = nu; } else { nu -= 0.5; s1 = 0.5 * (k + 1); s2 = nu; }
          const double rho = P.rho[(size_t) (rho_mark + k)];
          const double d_rho = dpar[1 + nc + k] + (s1 - 1.0) / rho - (s2 - 1.0) / (1.0 - rho);
          g_rho[rho_mark + k] = d_rho * rho * (1.0 - rho) + (1.0 - 2.0 * rho);
        }
      }
      for (int k = 0; k < nzT; ++k) g_zT[zT_mark + k] += dpar[1 + nc + (nc - 1) + k];
      zeta_mark += nc; rho_mark += nc - 1; zT_mark += nzT;
    } else if (p_[(size_t) i] == 1) {
      const double theta = P.theta_L[(size_t) th++];
      for (int s = 0; s < l_[(size_t) i]; ++s) {
        const double db = gb[(size_t) (b_mark + s)] * inv_s2;
        g_zb[b_mark + s] = theta * db - P.z_b[b_mark + s];
        adj_c += db * P.z_b[b_mark + s];
      }
      b_mark += l_[(size_t) i];
    } else {
      const double T11 = P.theta_L[(size_t) th], T21 = P.theta_L[(size_t) th + 1], T22 = P.theta_L[(size_t) th + 2]; th += 3;
      double a11 = 0.0, a21 = 0.0, a22 = 0.0;
      for (int j = 0; j < l_[(size_t) i]; ++j) {
        const double db0 = gb[(size_t) b_mark] * inv_s2, db1 = gb[(size_t) b_mark + 1] * inv_s2;
        const double z0 = P.z_b[b_mark], z1 = P.z_b[b_mark + 1];
        g_zb[b_mark] = T11 * db0 + T21 * db1 - z0;
        g_zb[b_mark + 1] = T22 * db1 - z1;
        a11 += db0 * z0; a21 += db1 * z0; a22 += db1 * z1;
        b_mark += 2;
      }
      const double c = tau * sc * P.disp;
      const double trace = 2.0 * c * c;
      const double z1v = P.zeta[(size_t) zeta_mark], z2v = P.zeta[(size_t) zeta_mark + 1], zs = z1v + z2v;
      const double pi1 = z1v / zs, pi2 = z2v / zs;
      const double sd1 = std::sqrt(pi1 * trace), sd2 = std::sqrt(pi2 * trace);
      const double rho = P.rho[(size_t) rho_mark];
      const double r = 2.0 * rho - 1.0, sq = std::sqrt(1.0 - r * r);
      adj_c = (a11 * T11 + a21 * T21 + a22 * T22) / c;        // every entry of T is linear in c
      const double a_sd1 = a11, a_sd2 = a21 * r + a22 * sq;
      const double a_pi1 = a_sd1 * sd1 / (2.0 * pi1), a_pi2 = a_sd2 * sd2 / (2.0 * pi2);
      const double dot = a_pi1 * pi1 + a_pi2 * pi2;
      double a_z1 = (a_pi1 - dot) / zs, a_z2 = (a_pi2 - dot) / zs;
      double a_rho = 2.0 * (a21 * sd2 - a22 * sd2 * r / sq);
      const double nu = regularization_[(size_t) pos_reg++] + 0.5 * (p_[(size_t) i] - 2);
      a_rho += (nu - 1.0) / rho - (nu - 1.0) / (1.0 - rho);
      g_rho[rho_mark] = a_rho * rho * (1.0 - rho) + (1.0 - 2.0 * rho);
      a_z1 += (delta_[(size_t) zeta_mark] - 1.0) / z1v - 1.0;
      a_z2 += (delta_[(size_t) zeta_mark + 1] - 1.0) / z2v - 1.0;
      g_zeta[zeta_mark] = a_z1 * z1v + 1.0;
      g_zeta[zeta_mark + 1] = a_z2 * z2v + 1.0;
      zeta_mark += 2; rho_mark += 1;
    }
    double a_tau = adj_c * sc * P.disp + (shape_[(size_t) i] - 1.0) / tau - 1.0;
    adj_disp += adj_c * tau * sc;
    g_tau[i] = a_tau * tau + 1.0;
  }
  if (has_aux_) {
    const double a_aux = adj_disp + S / (sigma * sigma * sigma) - N / sigma;
    const double a_au = a_aux * (prior_dist_for_aux_ == 0 ? 1.0 : prior_scale_for_aux_) + d_au_prior;
    grad[pos] = a_au * P.aux_unscaled + 1.0;
  }
  *lp_out = lp;
  bool bad = !std::isfinite(lp);
  for (int i = 0; i < num_params_; ++i) if (!std::isfinite(grad[i])) bad = true;
  return bad ? 1 : 0;
}

void GlmmModel::write_array(const double* q, double* out) const
{
  Params P; transform(q, P);
  int pos = 0;
  for (int k = 0; k < len_zbeta_; ++k) out[pos++] = P.z_beta[k];
  for (int k = 0; k < len_extra_; ++k) out[pos++] = P.extra[(size_t) k];
  for (int k = 0; k < q_; ++k) out[pos++] = P.z_b[k];
  for (int k = 0; k < len_z_T_; ++k) out[pos++] = P.z_T[k];
  for (double v : P.rho) out[pos++] = v;
  for (double v : P.zeta) out[pos++] = v;
  for (double v : P.tau) out[pos++] = v;
  if (has_aux_) { out[pos++] = P.aux_unscaled; out[pos++] = P.aux; }
  for (double v : P.beta) out[pos++] = v;
  for (double v : P.b) out[pos++] = v;
  for (double v : P.theta_L) out[pos++] = v;
}

}  // namespace s4b
