// stan4bart_b200/csrc/leaf_stats.cuh
// Stand-alone leaf sufficient statistics of one tree (the "leaf-stat" half of BASELINE.json's metric): per bottom
// node (n, sum, sum of squares) of the partial residual r_i + mu_leaf(i), in ONE launch.
//
//   * grid = a multiple of the SM count; every thread streams quads of 4 rows: the residuals as 2 x 16-byte loads, the binned
//     predictors as one 32-bit word per rule and quad (coalesced: column major, 4 rows per word); four quads per iteration (~150 bytes per thread in flight, 1024 threads per SM);
//   * the tree's rules sit in shared memory; a row's rule outcomes form a bit pattern that indexes a 256-entry table of leaf
//     slots (trees with more than 8 rules walk node by node);
//   * accumulation into lane-private shared-memory bins (no atomics), then a fixed-order reduction: threads -> warp (shuffles)
//     -> CTA -> one partial row per CTA -> the last CTA to finish sums the rows in CTA order.  Deterministic run to run.
// Algorithmic bytes per row: 8 (residual) + 2 (node id) + 1 (split column) = 11; actually read: 8 + the number of
// rules of the tree.
#pragma once

#include "s4b_common.cuh"

namespace s4b {

constexpr int kLeafBlock = 256;
constexpr int kLeafSlots = 8;              // bottom nodes handled by the fast kernel (more: the generic per-tree pass); 46 KB of bins => 4 CTAs per SM

struct LeafSmem {
  uint32_t irec[S4B_NODE_CAP];             // rule i (internal nodes in index order): var << 8 | cut
  uint8_t table[256];                      // rule pattern -> slot (n_int <= 8)
  uint32_t trav[S4B_NODE_CAP];             // node walk (n_int > 8): var << 16 | cut << 8 | right ; 0xFFFF.. for a bottom node
  uint8_t slot[S4B_NODE_CAP];
  double val[kLeafSlots + 1];
  int n_int, n_leaves, nn, fits;
  int last;
};

template <int V> struct LeafInt { static constexpr int value = V; };

__device__ __forceinline__ double leaf_wsum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// warp 0 of the CTA: the tree's rules in internal-node order, its bottom nodes' slots in node order, the rule-pattern -> slot table.
// The node count and the first 32 nodes are read in the same round trip (the node array has S4B_NODE_CAP >= 32 entries, so the read
// is in bounds whatever the count is), and the table is walked in shared memory: one L2 latency instead of one per dependent read --
// at n = 1 M the pass itself takes ~2 us, so this set-up is a visible share of the launch.
__device__ __forceinline__ void leaf_setup(LeafSmem& S, const DTree& t, int lane)
{
  static_assert(S4B_NODE_CAP >= 32, "the first pass reads 32 nodes unconditionally");
  const uint2 first = *reinterpret_cast<const uint2*>(&t.nodes[lane]);          // var | cut << 16, right | parent << 16
  const double first_mu = t.nodes[lane].mu;
  const int nn = t.num_nodes;
  int n_int = 0, n_leaf = 0;
  for (int base = 0; base < nn; base += 32) {
    const int k = base + lane;
    uint2 rec = first; double mu = first_mu;
    if (base > 0 && k < nn) { rec = *reinterpret_cast<const uint2*>(&t.nodes[k]); mu = t.nodes[k].mu; }
    const int var = (int) (int16_t) (rec.x & 0xFFFFu), cut = (int) (int16_t) (rec.x >> 16), right = (int) (int16_t) (rec.y & 0xFFFFu);
    const bool in = k < nn && var >= 0, lf = k < nn && var < 0;
    const unsigned mi = __ballot_sync(0xffffffffu, in), ml = __ballot_sync(0xffffffffu, lf);
    const unsigned below = (1u << lane) - 1u;
    if (in) S.irec[n_int + __popc(mi & below)] = ((uint32_t) var << 8) | (uint32_t) (cut & 0xFF);
    if (k < nn) {
      S.slot[k] = lf ? (uint8_t) min(n_leaf + __popc(ml & below), kLeafSlots) : (uint8_t) 255;
      S.trav[k] = lf ? 0xFFFFFFFFu : (((uint32_t) var << 16) | ((uint32_t) (cut & 0xFF) << 8) | (uint32_t) (right & 0xFF));
      if (lf && n_leaf + __popc(ml & below) < kLeafSlots) S.val[n_leaf + __popc(ml & below)] = mu;
    }
    n_int += __popc(mi); n_leaf += __popc(ml);
  }
  __syncwarp();
  if (lane == 0) { S.n_int = n_int; S.n_leaves = n_leaf; S.nn = nn; S.fits = n_leaf <= kLeafSlots ? 1 : 0; S.val[kLeafSlots] = 0.0; }
  if (n_int <= 8 && nn <= 32) {
    // internal-node mask of the (<= 32-node) tree, then every pattern's bottom node
    const unsigned imask = __ballot_sync(0xffffffffu, lane < nn && S.trav[lane < nn ? lane : 0] != 0xFFFFFFFFu);
    for (int e = lane; e < (1 << n_int); e += 32) {
      int node = 0;
      while ((imask >> node) & 1u) { const int id = __popc(imask & ((1u << node) - 1u)); node = ((e >> id) & 1) ? node + 1 : (int) (S.trav[node] & 0xFFu); }
      S.table[e] = S.slot[node];
    }
  }
}

// every thread's bins sit in shared memory ([slot][thread]): CTA reduction, one partial row per CTA, the last CTA sums the rows
__device__ __forceinline__ void leaf_epilogue(LeafSmem& S, const double2* bins, const int* cnts, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                              double* __restrict__ out, int* __restrict__ fits_out)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  // ---- CTA reduction in a fixed order: warp w takes slots w, w + 8, ...; one partial row (n, sum, sum of squares) per slot and CTA ----
  const int L = S.n_leaves, G = gridDim.x;
  for (int s = warp; s < L; s += kLeafBlock / 32) {
    double a = 0.0, b = 0.0; int c = 0;
#pragma unroll
    for (int i = 0; i < kLeafBlock / 32; ++i) { const double2 v = bins[s * kLeafBlock + i * 32 + lane]; a += v.x; b += v.y; c += cnts[s * kLeafBlock + i * 32 + lane]; }
    a = leaf_wsum(a); b = leaf_wsum(b); c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) { partials[(size_t) (3 * s) * G + blockIdx.x] = (double) c; partials[(size_t) (3 * s + 1) * G + blockIdx.x] = a; partials[(size_t) (3 * s + 2) * G + blockIdx.x] = b; }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) S.last = atomicAdd(ticket, 1u) == (unsigned) (G - 1) ? 1 : 0;
  __syncthreads();
  if (!S.last) return;
  __threadfence();
  // ---- the last CTA sums the partial rows of all CTAs in CTA order (lane-strided, then a shuffle tree: the same order every run) ----
  for (int v = warp; v < 3 * L; v += kLeafBlock / 32) {
    const double* src = partials + (size_t) v * G;
    double acc = 0.0;
    for (int b = lane; b < G; b += 32) acc += __ldcg(src + b);
    acc = leaf_wsum(acc);
    if (lane == 0) out[v] = acc;
  }
  if (tid == 0) { *ticket = 0u; *fits_out = 1; }
}

// out: [3 * slot + {0, 1, 2}] = n, sum, sum of squares; *fits_out = 0 when the tree has more than kLeafSlots bottom nodes
__global__ void __launch_bounds__(kLeafBlock) k_leaf_stats(long long n, long long npad, const uint8_t* __restrict__ xt, const double* __restrict__ R,
                                                           const DTree* __restrict__ trees, int tree_index, double* __restrict__ partials,
                                                           unsigned int* __restrict__ ticket, double* __restrict__ out, int* __restrict__ fits_out)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LeafSmem& S = *reinterpret_cast<LeafSmem*>(smem_raw);
  double2* bins = reinterpret_cast<double2*>(smem_raw + ((sizeof(LeafSmem) + 15) / 16) * 16);        // [kLeafSlots + 1][kLeafBlock]: (sum, sum of squares)
  int* cnts = reinterpret_cast<int*>(bins + (kLeafSlots + 1) * kLeafBlock);                            // [kLeafSlots + 1][kLeafBlock]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) leaf_setup(S, trees[tree_index], lane);
  for (int k = tid; k < (kLeafSlots + 1) * kLeafBlock; k += kLeafBlock) { bins[k] = make_double2(0.0, 0.0); cnts[k] = 0; }
  __syncthreads();
  if (!S.fits) { if (blockIdx.x == 0 && tid == 0) *fits_out = 0; return; }
  const int n_int = S.n_int;
  const bool bitmap = n_int <= 8 && S.nn <= 32;
  const uint32_t* xt32 = reinterpret_cast<const uint32_t*>(xt);
  const long long col_words = npad >> 2, nquad = (n + 3) >> 2;
  const long long stride = (long long) gridDim.x * kLeafBlock;

  auto slots_of = [&](long long q) -> uint32_t {
    if (bitmap) {
      uint32_t pat = 0u;
      for (int i = 0; i < n_int; ++i) {
        const uint32_t rec = S.irec[i];
        pat |= __vsetleu4(__ldg(xt32 + (long long) (rec >> 8) * col_words + q), (rec & 0xFFu) * 0x01010101u) << i;
      }
      return (uint32_t) S.table[pat & 0xFFu] | ((uint32_t) S.table[(pat >> 8) & 0xFFu] << 8) | ((uint32_t) S.table[(pat >> 16) & 0xFFu] << 16) |
             ((uint32_t) S.table[pat >> 24] << 24);
    }
    uint32_t res = 0u;
    for (int o = 0; o < 4; ++o) {
      int node = 0;
      uint32_t tr = S.trav[0];
      while (tr != 0xFFFFFFFFu) {
        const uint32_t w = __ldg(xt32 + (long long) (tr >> 16) * col_words + q);
        node = ((w >> (8 * o)) & 0xFFu) <= ((tr >> 8) & 0xFFu) ? node + 1 : (int) (tr & 0xFFu);
        tr = S.trav[node];
      }
      res |= (uint32_t) S.slot[node] << (8 * o);
    }
    return res;
  };
  auto add_quad = [&](long long q, double2 a, double2 b, uint32_t sl) {
    const double r[4] = { a.x, a.y, b.x, b.y };
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      int s = (sl >> (8 * o)) & 0xFF;
      if (4 * q + o >= n) s = kLeafSlots;                       // padding rows of the last quad: trash row
      const double pr = r[o] + S.val[s];
      double2 v = bins[s * kLeafBlock + tid];
      v.x += pr; v.y = fma(pr, pr, v.y);
      bins[s * kLeafBlock + tid] = v;
      cnts[s * kLeafBlock + tid] += 1;
    }
  };
  for (long long q0 = (long long) blockIdx.x * kLeafBlock + tid; q0 < nquad; q0 += 4 * stride) {
    // all global loads of four quads are issued before the first is used
    double2 a[4], b[4]; uint32_t sl[4]; bool live[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long q = q0 + k * stride;
      live[k] = q < nquad;
      a[k] = make_double2(0.0, 0.0); b[k] = a[k];
      if (live[k]) { a[k] = __ldg(reinterpret_cast<const double2*>(R + 4 * q)); b[k] = __ldg(reinterpret_cast<const double2*>(R + 4 * q + 2)); }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) sl[k] = live[k] ? slots_of(q0 + k * stride) : 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (live[k]) add_quad(q0 + k * stride, a[k], b[k], sl[k]);
  }
  leaf_epilogue(S, bins, cnts, partials, ticket, out, fits_out);
}

// ---------------------------------------------------------------------------------------
// The same pass, software-pipelined, templated on the number of bottom nodes it holds (LM; BART trees average 2-3), as ncu asked for it (profiles/README.md, round 2:
// ncu_k_leaf_stats_r2_*): with k_leaf_stats a row costs ~12 shared-memory wavefronts per warp and the L1 / shared-memory pipe was the
// busiest unit (75 %) at 3.3 TB/s; and every warp of the SM issued its loads, waited and computed in step, so memory and arithmetic did
// not overlap (long scoreboard 6 per issue at 46 % issue utilisation).  Here
//   * software pipelining: the loads of the next QB quads are in flight while the current QB quads are accumulated;
//   * the rule outcomes of a row come from byte compares on the prefetched predictor words and a 4-bit-per-pattern table held in one
//     register (<= 3 rules => 8 patterns): no table look-up in shared memory;
//   * MODE 1 (<= 2 bottom nodes): the thread's bins are registers, and a row is added to every bin - as +0.0 / fma(0, pr, .) to the bins
//     it does not belong to, which leaves them bit for bit unchanged - so the loop has no branch and no shared-memory access at all;
//   * MODE 2 (3-4 bottom nodes, where the selects of MODE 1 cost more than they save): (sum, sum of squares) in the thread's
//     shared-memory bins, the counts in registers.
//   * 5-8 bottom nodes (LM = 8): the same pipelining and compares, the pattern table and the counts in shared memory.
// Per-thread sums are formed in the same order as in k_leaf_stats; results differ only through the grid size (the order of the final
// sums); counts are exact.  32 M rows: 59 us (<= 2 bottom nodes) / 80 us (3-4) against 86 / 88-109 us with k_leaf_stats.
// ---------------------------------------------------------------------------------------
template <int LM, int QB, int MODE>
__global__ void __launch_bounds__(kLeafBlock, 2) k_leaf_stats_small(long long n, long long npad, const uint8_t* __restrict__ xt, const double* __restrict__ R,
                                                                    const DTree* __restrict__ trees, int tree_index, double* __restrict__ partials,
                                                                    unsigned int* __restrict__ ticket, double* __restrict__ out, int* __restrict__ fits_out)
{
  static_assert(LM >= 1 && LM <= kLeafSlots, "bins for kLeafSlots bottom nodes");
  static_assert(MODE == 2 || (MODE == 1 && LM == 2), "register bins: two bottom nodes");
  constexpr bool kRegTable = LM <= 4;        // <= 3 rules: 8 patterns x 4 bits in one register; more: the table in shared memory
  constexpr int NR = LM > 1 ? LM - 1 : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LeafSmem& S = *reinterpret_cast<LeafSmem*>(smem_raw);
  double2* bins = reinterpret_cast<double2*>(smem_raw + ((sizeof(LeafSmem) + 15) / 16) * 16);
  int* cnts = reinterpret_cast<int*>(bins + (kLeafSlots + 1) * kLeafBlock);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long nquad = n >> 2;                                // full quads only in the loop: a ragged last quad is added at the end
  const long long stride = (long long) gridDim.x * kLeafBlock;
  struct Buf { double2 a[QB], b[QB]; uint32_t w[QB][NR]; };
  auto prefetch_r = [&](Buf& B, long long q0) {
#pragma unroll
    for (int k = 0; k < QB; ++k) {
      const long long q = q0 + k * stride;
      if (q < nquad) { B.a[k] = __ldg(reinterpret_cast<const double2*>(R + 4 * q)); B.b[k] = __ldg(reinterpret_cast<const double2*>(R + 4 * q + 2)); }
    }
  };
  Buf A, B;
  // the residuals of the first quads do not depend on the tree: their loads are in flight while warp 0 reads and prepares the tree
  prefetch_r(A, (long long) blockIdx.x * kLeafBlock + tid);
  if (warp == 0) leaf_setup(S, trees[tree_index], lane);
  if (MODE == 2) {
#pragma unroll
    for (int j = 0; j < LM; ++j) { bins[j * kLeafBlock + tid] = make_double2(0.0, 0.0); if (!kRegTable) cnts[j * kLeafBlock + tid] = 0; }
  }
  __syncthreads();
  if (S.n_leaves > LM) { if (blockIdx.x == 0 && tid == 0) *fits_out = 0; return; }       // (the host picks the kernel by the tree's size)
  const int n_int = S.n_int;
  const uint32_t* xt32 = reinterpret_cast<const uint32_t*>(xt);
  const long long col_words = npad >> 2;
  // rules in registers; a rule the tree does not have reads nothing and never sets its bit.
  // x <= cut for the four rows of a quad at once, in 16-bit lanes: (cut + 256) - x has bit 8 set iff x <= cut
  const uint32_t* col[NR]; uint32_t cutk[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const uint32_t rec = i < n_int ? S.irec[i] : 0u;
    col[i] = xt32 + (long long) (rec >> 8) * col_words;
    cutk[i] = i < n_int ? ((rec & 0xFFu) | 0x100u) * 0x00010001u : 0x00FF00FFu;       // (no such rule: 255 - 0, bit 8 never set)
  }
  // pattern -> slot, 4 bits per pattern; the pattern arrives multiplied by 4 (rule i sets bit i + 2)
  uint32_t tbl = 0u;
#pragma unroll
  for (int e = 0; e < 8; ++e) tbl |= ((uint32_t) S.table[e & ((1 << n_int) - 1)] & 0xFu) << (4 * e);
  double sum[LM], sq[LM], val[LM];
  int cnt1 = 0, rows_done = 0;             // MODE 1 (two bins): rows of bin 1 and all rows
  unsigned long long cntp = 0ull;          // MODE 2: four 16-bit counts (the host keeps rows per thread below 65536)
#pragma unroll
  for (int j = 0; j < LM; ++j) { sum[j] = 0.0; sq[j] = 0.0; val[j] = S.val[j]; }

  auto add_row = [&](double r, int s) {
    if (MODE == 1) {
      // branch free: a row is added to every bin, as +0.0 / fma(0, pr, .) to the bins it does not belong to (bit for bit the same sums)
#pragma unroll
      for (int j = 0; j < LM; ++j) {
        const double pr = r + val[j];
        const double pm = s == j ? pr : 0.0;
        sum[j] += pm; sq[j] = fma(pm, pr, sq[j]);
      }
      cnt1 += s == 1 ? 1 : 0; rows_done += 1;
    } else {
      // (sum, sum of squares) in the thread's shared-memory bin, the counts packed in a register pair
      const double pr = r + S.val[s];
      double2 v = bins[s * kLeafBlock + tid];
      v.x += pr; v.y = fma(pr, pr, v.y);
      bins[s * kLeafBlock + tid] = v;
      if (kRegTable) cntp += 1ull << (16 * s);
      else cnts[s * kLeafBlock + tid] += 1;
    }
  };

  auto prefetch_w = [&](Buf& B, long long q0) {
#pragma unroll
    for (int k = 0; k < QB; ++k) {
      const long long q = q0 + k * stride;
      if (q < nquad) {
#pragma unroll
        for (int i = 0; i < NR; ++i) B.w[k][i] = i < n_int ? __ldg(col[i] + q) : 0u;
      }
    }
  };
  auto prefetch = [&](Buf& B, long long q0) { prefetch_r(B, q0); prefetch_w(B, q0); };
  auto accumulate = [&](const Buf& B, long long q0) {
#pragma unroll
    for (int k = 0; k < QB; ++k) {
      if (q0 + k * stride >= nquad) break;
      // rows 0 and 2 of the quad in the 16-bit lanes of pe, rows 1 and 3 in those of po: 4 x (pattern of the row)
      uint32_t pe = 0u, po = 0u;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const uint32_t w = B.w[k][i];
        const uint32_t te = cutk[i] - (w & 0x00FF00FFu), to = cutk[i] - ((w >> 8) & 0x00FF00FFu);
        pe |= (te >> (6 - i)) & (0x00010001u << (i + 2));
        po |= (to >> (6 - i)) & (0x00010001u << (i + 2));
      }
      if (kRegTable) {
        add_row(B.a[k].x, (int) ((tbl >> (pe & 0xFFFFu)) & 0xFu));
        add_row(B.a[k].y, (int) ((tbl >> (po & 0xFFFFu)) & 0xFu));
        add_row(B.b[k].x, (int) ((tbl >> (pe >> 16)) & 0xFu));
        add_row(B.b[k].y, (int) ((tbl >> (po >> 16)) & 0xFu));
      } else {
        add_row(B.a[k].x, (int) S.table[(pe & 0xFFFFu) >> 2]);
        add_row(B.a[k].y, (int) S.table[(po & 0xFFFFu) >> 2]);
        add_row(B.b[k].x, (int) S.table[pe >> 18]);
        add_row(B.b[k].y, (int) S.table[po >> 18]);
      }
    }
  };
  {
    long long q0 = (long long) blockIdx.x * kLeafBlock + tid;
    const long long step = QB * stride;
    prefetch_w(A, q0);
    while (q0 < nquad) {
      prefetch(B, q0 + step);
      accumulate(A, q0);
      q0 += step;
      if (q0 >= nquad) break;
      prefetch(A, q0 + step);
      accumulate(B, q0);
      q0 += step;
    }
  }
  if ((n & 3) != 0 && blockIdx.x == 0 && tid == 0) {
    // the ragged last quad, row by row
    for (long long row = 4 * nquad; row < n; ++row) {
      int pat = 0;
      for (int i = 0; i < n_int; ++i) { const uint32_t rec = S.irec[i]; pat |= ((uint32_t) xt[(long long) (rec >> 8) * npad + row] <= (rec & 0xFFu) ? 1 : 0) << i; }
      add_row(R[row], (int) S.table[pat]);
    }
  }
#pragma unroll
  for (int j = 0; j < LM; ++j) {
    if (MODE == 1) { bins[j * kLeafBlock + tid] = make_double2(sum[j], sq[j]); cnts[j * kLeafBlock + tid] = j == 0 ? rows_done - cnt1 : j == 1 ? cnt1 : 0; }
    else if (kRegTable) cnts[j * kLeafBlock + tid] = (int) ((cntp >> (16 * j)) & 0xFFFFull);
  }
  leaf_epilogue(S, bins, cnts, partials, ticket, out, fits_out);
}

}  // namespace s4b
