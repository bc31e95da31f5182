// stan4bart_b200/csrc/sweep_pipe.cuh
// Software-pipelined variant of the persistent BART sweep (sweep_kernel.cuh): the production path of an unweighted,
// unsharded chain whose trees are small (every tree <= 30 nodes and <= 8 statistic slots this sweep; k_prepare_sweep decides
// per sweep, otherwise the synchronous kernel runs).
//
// In the synchronous kernel a tree step is ONE serial chain: accumulate -> CTA reduce -> grid barrier -> 148-row reduce ->
// Metropolis decision -> residual update -> next accumulate (12.9 k cycles, no phase long).  Here the workers and the
// controller run one step apart:
//
//   workers   : ... | U(t-2) W(t) A(t) reduce, ARRIVE(t) | U(t-1) W(t+1) A(t+1) reduce, ARRIVE(t+1) | ...     (never wait for a decision)
//   controller: ...        | wait(t-1) rows(t-1) correct D(t-1) | wait(t) rows(t) correct D(t) | ...
//
// A(t) needs the residuals after update t-1, which the controller is still deciding.  But update t-1 adds one constant per
// CELL of step t-1's partition (cell = bottom node under either outcome of the proposal), so
//     sum_{i in slot s of t} r_i(after t-1) = sum_{i in s} r_i(after t-2) + sum_c n[s][c] * delta_{t-1}[c],
// where n[s][c] counts the rows in slot s of step t and cell c of step t-1.  The workers accumulate the first sum and the
// integer cross table n (exact, order independent); the controller adds the correction once delta_{t-1} is known (it is its
// own previous output).  Every CTA's controller does the same arithmetic in the same order on the same global partial rows,
// so all CTAs hold bitwise identical statistics and take identical decisions, as before.
//
// Cross-CTA reduction: per step every CTA adds its partial results into ONE small accumulator row with integer atomics
// (red.global.add.u64) before it arrives at the barrier: the cross-table counts as packed integers, the per-slot sums as exact
// two-limb fixed point (value = hi 2^-20 + lo 2^-72: every double of magnitude < 2^43 is represented exactly, integer addition
// is associative), so the totals do not depend on the order in which the CTAs arrive -- bitwise reproducible run to run, and
// identical in every CTA because all read the same words.  Every accumulator word also counts its contributions in its top
// byte (each CTA adds 1 << 56 together with its data), so a word tells by itself when it is complete: there is no separate
// barrier counter, no release fence on the workers' side and no flag-then-data double round trip on the controller's --
// it polls the very words it needs, one L2 round trip per step instead of reading 148 partial rows.
//
// Drained steps: a step whose cross table would not fit the counters -- (its slots) x (the previous step's cells) above the capacity
// the launch was given -- waits for the previous decision as well and applies it before it accumulates: its rows then have no
// update pending (one cell, no correction), at the price of one synchronous step; every thread derives the same flag from the step
// descriptors, so nothing is communicated.  The workers add the plain residuals: the leaf values of the current tree enter as
// count x value on the controller's side (one multiply per slot instead of one add per row).
//
// Rings: accumulator rows and barrier counters are 4 deep (a CTA can run at most two steps ahead of another CTA's controller
// reading rows); descriptors 3 deep; update tables 2 deep.  Decisions, draws and tree updates are the synchronous kernel's
// (w_plan / w_decide_fast), so both kernels produce the same chain up to the rounding of the slot sums.
#pragma once

#include "sweep_kernel.cuh"

namespace s4b {

constexpr int kPipeAcc = 2 * kPipeSlots + kPipeCross / 2;     // per step: (hi, lo) per slot sum, then one word per pair of cross-table entries
constexpr int kPipePacked = 24;            // cross tables up to this many entries are counted in two packed registers per thread (5-bit fields)
struct PipeSmem {
  StepDesc sd[kPipeDescs];
  PipeInfo info[kPipeDescs];
  DTree tree[2];
  UpdateDesc upd[2];
  unsigned long long accprev[kPipeRing][kPipeAcc];   // accumulator rows as last read (a row is reused every kPipeRing steps and never zeroed)
  double2 draws[2][32];                     // decision draws of the step being decided / the next one
  double dcell[2][kPipeCells];              // delta of step parity: mu_old - mu_new per cell
  int ncnt[4 * kPipeSlots];                  // staging of the (hi, lo) limbs of the slot sums
  int cross[kPipeCross + 2];                 // the cross table of the step being decided
  CtlScratch csd;
  LeafStat st[S4B_MAX_SLOTS];
  FastPlanSmem plan;
  double inv_sigsq;
  int fail;
  RngState rng;
  BartParams prm;
};

// 1 << sh for sh < 64, 0 otherwise (PTX shl clamps the shift amount; an amount that wrapped below zero is a large unsigned number)
__device__ __forceinline__ unsigned long long shl64_clamp(unsigned sh)
{
  unsigned long long r;
  asm("shl.b64 %0, 1, %1;" : "=l"(r) : "r"(sh));
  return r;
}

__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// Asynchronous fetch of a step descriptor (the part the sweep reads: from b_tree on) and its cell table into the shared-memory
// ring: one warp issues cp.async copies and goes on with its work -- no register staging, no stall on the L2 round trip.  The
// descriptors in global memory were laid out by w_copy_desc in k_prepare_sweep, so a plain copy gives the same contents.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc)
{
  const unsigned d = (unsigned) __cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ inline void w_fetch_desc_async(StepDesc& dst, const StepDesc& src, PipeInfo& idst, const PipeInfo& isrc, int lane)
{
  constexpr size_t kFrom = offsetof(StepDesc, b_tree);
  static_assert(kFrom % 8 == 0 && sizeof(StepDesc) % 8 == 0 && sizeof(PipeInfo) % 8 == 0, "8-byte copies");
  static_assert(offsetof(PipeInfo, stab) % 4 == 0 && offsetof(PipeInfo, ptab) % 4 == 0, "the pattern tables are read as 32-bit words");
  const char* s = reinterpret_cast<const char*>(&src) + kFrom;
  char* d = reinterpret_cast<char*>(&dst) + kFrom;
  for (int i = lane; i < (int) ((sizeof(StepDesc) - kFrom) / 8); i += 32) cp_async8(d + 8 * i, s + 8 * i);
  for (int i = lane; i < (int) (sizeof(PipeInfo) / 8); i += 32) cp_async8(reinterpret_cast<char*>(&idst) + 8 * i, reinterpret_cast<const char*>(&isrc) + 8 * i);
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ inline void w_copy_info(PipeInfo& dst, const PipeInfo& src, int lane)
{
  for (int i = lane; i < (int) (sizeof(PipeInfo) / 8); i += 32) reinterpret_cast<unsigned long long*>(&dst)[i] = reinterpret_cast<const unsigned long long*>(&src)[i];
  __syncwarp();
}

// rule patterns of the owned quads -> statistic slots (one table look-up per row); a change / swap step also under the proposed
// rules; a birth step moves the rows of the node to split to the two new slots L + side
template <int NQ>
__device__ __forceinline__ void pipe_walk(const StepDesc& sd, const PipeInfo& pi, const uint32_t* __restrict__ tile, int tile_stride, int tid,
                                          uint32_t (&sp)[NQ], uint32_t (&pp)[NQ], int nq /* quads this warp owns (warp uniform): the others are skipped */)
{
  const int kind = sd.b_kind, n_int = sd.b_cur.n_int;
  uint32_t pat[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) pat[j] = 0u;
#pragma unroll 1
  for (int i = 0; i < n_int; ++i) {
    const uint32_t rec = sd.b_cur.irec[i];
    const uint32_t* col = tile + (rec >> 8) * tile_stride + tid;
    const uint32_t cut4 = (rec & 0xFFu) * 0x01010101u;
#pragma unroll
    for (int j = 0; j < NQ; ++j) if (j < nq) pat[j] |= __vsetleu4(col[j * kWorkers], cut4) << i;
  }
  // pattern -> slot.  Up to three rules (eight patterns) the table is two registers and the four look-ups of a quad are two byte
  // permutes: the pattern bytes are folded into the four selector nibbles, PRMT then picks the four slot bytes at once
  const bool small = n_int <= 3;
  if (small) {
    const uint32_t tlo = *reinterpret_cast<const uint32_t*>(pi.stab), thi = *reinterpret_cast<const uint32_t*>(pi.stab + 4);
#pragma unroll
    for (int j = 0; j < NQ; ++j) if (j < nq) { const uint32_t x = pat[j] | (pat[j] >> 4); sp[j] = __byte_perm(tlo, thi, __byte_perm(x, 0u, 0x4420u)); }
  } else {
#pragma unroll
    for (int j = 0; j < NQ; ++j) if (j < nq)
      sp[j] = (uint32_t) pi.stab[pat[j] & 0xFFu] | ((uint32_t) pi.stab[(pat[j] >> 8) & 0xFFu] << 8) | ((uint32_t) pi.stab[(pat[j] >> 16) & 0xFFu] << 16) |
              ((uint32_t) pi.stab[pat[j] >> 24] << 24);
  }
  if (kind == 2 || kind == 3) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) pat[j] = 0u;
#pragma unroll 1
    for (int i = 0; i < n_int; ++i) {
      const uint32_t rec = sd.b_prop.irec[i];
      const uint32_t* col = tile + (rec >> 8) * tile_stride + tid;
      const uint32_t cut4 = (rec & 0xFFu) * 0x01010101u;
#pragma unroll
      for (int j = 0; j < NQ; ++j) if (j < nq) pat[j] |= __vsetleu4(col[j * kWorkers], cut4) << i;
    }
    if (small) {
      const uint32_t tlo = *reinterpret_cast<const uint32_t*>(pi.ptab), thi = *reinterpret_cast<const uint32_t*>(pi.ptab + 4);
#pragma unroll
      for (int j = 0; j < NQ; ++j) if (j < nq) { const uint32_t x = pat[j] | (pat[j] >> 4); pp[j] = __byte_perm(tlo, thi, __byte_perm(x, 0u, 0x4420u)); }
    } else {
#pragma unroll
      for (int j = 0; j < NQ; ++j) if (j < nq)
        pp[j] = (uint32_t) pi.ptab[pat[j] & 0xFFu] | ((uint32_t) pi.ptab[(pat[j] >> 8) & 0xFFu] << 8) | ((uint32_t) pi.ptab[(pat[j] >> 16) & 0xFFu] << 16) |
                ((uint32_t) pi.ptab[pat[j] >> 24] << 24);
    }
  } else if (kind == 0) {
    const uint32_t* col = tile + sd.b_var * tile_stride + tid;
    const uint32_t cut4 = (uint32_t) sd.b_cut * 0x01010101u;
    const uint32_t sb4 = (uint32_t) pi.slot_b * 0x01010101u, l4 = (uint32_t) sd.b_num_leaves * 0x01010101u;
#pragma unroll
    for (int j = 0; j < NQ; ++j) if (j < nq) {
      const uint32_t side = __vcmpgtu4(col[j * kWorkers], cut4) & 0x01010101u;
      const uint32_t at_b = __vcmpeq4(sp[j], sb4);                 // 0xFF in the bytes of rows that sit in the node to split
      sp[j] = (sp[j] & ~at_b) | ((l4 + side) & at_b);
    }
  }
}

// acc_ring: kPipeRing accumulator rows of kPipeAcc 64-bit words (zeroed by the host before every launch)
template <int NQ, bool PROF = false>
__global__ void __launch_bounds__(kSweepBlock, 1) k_sweep_pipe(BartDev dv, unsigned long long* acc_rings, int launch_parity,
                                                                const StepDesc* __restrict__ descs, const PipeInfo* __restrict__ infos,
                                                                const double2* __restrict__ draws, const int* __restrict__ pos_in, int* __restrict__ pos_out,
                                                                int count_entries, unsigned long long* __restrict__ ran, unsigned long long* __restrict__ prof)
{
  // this launch takes the run of consecutive steps that fit, starting at *pos_in; the synchronous kernel (launched next) takes
  // the step that stopped it.  Every CTA scans the same flags, so all agree on the range without talking to each other.
  // two accumulator rings alternate by launch: this launch adds into one (zero on entry) and clears the other for the next
  // launch, so no host-side memset sits between the launches of a sweep
  unsigned long long* acc_ring = acc_rings + (size_t) (launch_parity & 1) * kPipeRing * kPipeAcc;
  if (blockIdx.x == 0) {
    unsigned long long* other = acc_rings + (size_t) ((launch_parity + 1) & 1) * kPipeRing * kPipeAcc;
    for (int i = threadIdx.x; i < kPipeRing * kPipeAcc; i += blockDim.x) other[i] = 0ull;
  }
  const int T_all = dv.params->num_trees;
  const int t_begin = *pos_in;
  // a step fits when its trees are small (k_prepare_sweep's flag); when its cross table -- (its slots) x (the previous step's cells) --
  // does not fit the shared-memory counters it runs drained (below); the first step of a run has no predecessor in flight (one cell)
  // (every warp scans by itself, 8 x 32 flags per round with all loads issued before the first is used: a thread-serial scan costs one L2
  // round trip per step, ~0.1 ms for a 200-tree sweep)
  int t_end = T_all;
  {
    const int ln = threadIdx.x & 31;
    for (int base = t_begin; base < T_all && t_end == T_all; base += 256) {
      int okv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { const int t = base + 32 * k + ln; okv[k] = t < T_all ? __ldg(&infos[t].ok) : 0; }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const unsigned bad = __ballot_sync(0xffffffffu, okv[k] == 0);
        if (bad != 0u && t_end == T_all) t_end = min(T_all, base + 32 * k + __ffs(bad) - 1);
      }
    }
  }
  if (t_end == t_begin) { if (blockIdx.x == 0 && threadIdx.x == 0) *pos_out = t_begin; return; }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PipeSmem& S = *reinterpret_cast<PipeSmem*>(smem_raw);
  double* bins = reinterpret_cast<double*>(smem_raw + ((sizeof(PipeSmem) + 15) / 16) * 16);                    // [kPipeSlots + 1][kWorkers]
  unsigned long long* pkw = reinterpret_cast<unsigned long long*>(bins + (kPipeSlots + 1) * kWorkers);            // [2][kWorkers]: packed counters of small cross tables
  uint8_t* cnt = reinterpret_cast<uint8_t*>(pkw + 2 * kWorkers);                                                   // [count_entries + 1][kWorkers] bytes
  uint32_t* tile = reinterpret_cast<uint32_t*>(cnt + (size_t) (count_entries + 1) * kWorkers);                    // [p][NQ * kWorkers]
  constexpr int tile_stride = NQ * kWorkers;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_worker = tid < kWorkers;
  const int G = gridDim.x, cta = blockIdx.x;
  const long long n = dv.n, npad = dv.npad;
  const long long nquad = (n + 3) >> 2;
  const long long q_lo = nquad * cta / G, q_hi = nquad * (cta + 1) / G;

  double R[NQ][4];
  unsigned valid_mask = 0, obs_mask = 0;
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const long long q = q_lo + (long long) j * kWorkers + tid;
    if (is_worker && q < q_hi) {
      valid_mask |= 1u << j;
      for (int o = 0; o < 4; ++o) if (4 * q + o < n) obs_mask |= 1u << (4 * j + o);
      double2 a = *reinterpret_cast<const double2*>(dv.R + 4 * q), b = *reinterpret_cast<const double2*>(dv.R + 4 * q + 2);
      R[j][0] = a.x; R[j][1] = a.y; R[j][2] = b.x; R[j][3] = b.y;
    } else { R[j][0] = R[j][1] = R[j][2] = R[j][3] = 0.0; }
  }
  if (tid == 0) {
    const double sg = dv.params->sigma; S.inv_sigsq = 1.0 / (sg * sg);
    S.fail = 0; S.prm = *dv.params; S.rng = *dv.rng; S.csd.draws_total = 0; S.csd.prof_on = 0;
  }
  for (int i = tid; i < (int) (2 * sizeof(UpdateDesc) / sizeof(uint32_t)); i += kSweepBlock) reinterpret_cast<uint32_t*>(S.upd)[i] = 0u;
  for (int i = tid; i < 3 * S4B_MAX_SLOTS; i += kSweepBlock) reinterpret_cast<double*>(S.st)[i] = 0.0;
  if (tid < 2 * kPipeCells) (&S.dcell[0][0])[tid] = 0.0;
  for (int i = tid; i < kPipeRing * kPipeAcc; i += kSweepBlock) (&S.accprev[0][0])[i] = 0ull;
  __syncthreads();
  const int p = S.prm.p, T = S.prm.num_trees;
  const unsigned long long step0 = S.prm.step_id;
  const uint32_t* xt32 = reinterpret_cast<const uint32_t*>(dv.xt);
  const int col_words = (int) (npad >> 2);
  if (is_worker) {
    for (int v = 0; v < p; ++v)
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        const long long q = q_lo + (long long) j * kWorkers + tid;
        tile[v * tile_stride + j * kWorkers + tid] = ((valid_mask >> j) & 1u) ? __ldg(xt32 + (long long) v * col_words + q) : 0u;
      }
  } else {
    const DTree& g = dv.trees[t_begin];
    const int nn = g.num_nodes;
    if (lane == 0) { S.tree[t_begin & 1].num_nodes = nn; S.tree[t_begin & 1].pad = 0; }
    for (int i = lane; i < nn * (int) (sizeof(DNode) / 4); i += 32) reinterpret_cast<uint32_t*>(S.tree[t_begin & 1].nodes)[i] = reinterpret_cast<const uint32_t*>(g.nodes)[i];
    S.draws[t_begin & 1][lane] = __ldcg(draws + t_begin * 32 + lane);
    w_copy_desc(S.sd[t_begin % kPipeDescs], descs[t_begin], lane);
    w_copy_info(S.info[t_begin % kPipeDescs], infos[t_begin], lane);
  }
  __syncthreads();

  if (is_worker) {
    // =====================================================================================  workers
    uint32_t sp[NQ], pp[NQ], cprev[NQ], cprev2[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) { cprev[j] = 0u; cprev2[j] = 0u; pp[j] = 0u; }
    // rows beyond the data (the tail of the last quad, quads beyond q_hi) are parked in the trash slot kPipeSlots at every step
    constexpr unsigned kAllObs = NQ == 8 ? 0xFFFFFFFFu : ((1u << (4 * NQ)) - 1u);
    const bool ragged = obs_mask != kAllObs;
    // quads that hold data in at least one lane of this warp: the rows are dealt out quad-major (q_lo + j * 480 + tid), so at n = 1 M
    // the last quad is empty for the upper half of the warps -- they skip it altogether instead of computing on trash rows
    const int nq = __reduce_max_sync(0xffffffffu, 32 - __clz(valid_mask));

    int C = 1;                                 // cells of the previous step as the cross table sees them (1 at the start and after a drained step)
    int na = t_begin;                          // the next decision whose update the residuals have not seen yet
    // cycle counters (thread 0 of CTA 0, only when asked for): [0] wait for decisions, [1] U, [2] W, [3] A, [4] CTA reduce + arrive, [6] drained steps
    long long wp0 = 0, wp1 = 0, wp2 = 0, wp3 = 0, wp4 = 0, wp6 = 0;
    const bool wprof = PROF && prof != nullptr && cta == 0 && tid == 0;
    const int e_trash = count_entries;
    uint32_t* cntw = reinterpret_cast<uint32_t*>(cnt);
    for (int k = 0; k <= kPipeSlots; ++k) bins[k * kWorkers + tid] = 0.0;
    for (int i = tid; i < (count_entries + 1) * (kWorkers / 4); i += kWorkers) cntw[i] = 0u;
    named_bar_sync(1, kWorkers);
    for (int t = t_begin; t < t_end; ++t) {
      const StepDesc& sd = S.sd[t % kPipeDescs];
      const PipeInfo& pi = S.info[t % kPipeDescs];
      const long long k0 = PROF ? clock64() : 0;
      const int kind = sd.b_kind, L = sd.b_num_leaves, nslots = sd.b_nslots;
      // a cross table beyond the counters' capacity: this step waits for the previous decision too and starts from updated residuals
      const bool drained = t > t_begin && nslots * C > count_entries;
      // ---- U: apply the decisions up to t-2 (t-1 as well when drained): per-cell deltas ----
      long long kw = 0;
      for (const int need = drained ? t - 1 : t - 2; na <= need; ++na) {
        const long long w0 = PROF ? clock64() : 0;
        named_bar_sync(2 + (na & 1), kSweepBlock);
        if (PROF) kw += clock64() - w0;
        const double* dc = S.dcell[na & 1];
        const bool newest = na == t - 1;
#pragma unroll
        for (int j = 0; j < NQ; ++j) if (j < nq) {
          const uint32_t cw = newest ? cprev[j] : cprev2[j];
#pragma unroll
          for (int o = 0; o < 4; ++o) R[j][o] += dc[(cw >> (8 * o)) & 0xFF];
        }
      }
      if (drained) {
        C = 1;
#pragma unroll
        for (int j = 0; j < NQ; ++j) cprev[j] = 0u;
      }
      const long long k2 = PROF ? clock64() : 0;
      // one warp fetches the next step's descriptor (its ring slot held step t-2, which has been decided)
      if (warp == kWorkerWarps - 1 && t + 1 < t_end)
        w_fetch_desc_async(S.sd[(t + 1) % kPipeDescs], descs[t + 1], S.info[(t + 1) % kPipeDescs], infos[t + 1], lane);
      // ---- W(t): slots of every owned row ----
      pipe_walk<NQ>(sd, pi, tile, tile_stride, tid, sp, pp, nq);
      if (ragged) {
#pragma unroll
        for (int j = 0; j < NQ; ++j) if (j < nq) {
          const uint32_t nib = (~obs_mask >> (4 * j)) & 0xFu;                                   // rows beyond the data
          const uint32_t tm = ((nib * 0x00204081u) & 0x01010101u) * 0xFFu;                      // 0xFF in their bytes
          sp[j] = (sp[j] & ~tm) | (((uint32_t) kPipeSlots * 0x01010101u) & tm); pp[j] |= tm;
        }
      }
      const long long k3 = PROF ? clock64() : 0;
      const bool two_trees = (kind == 2 || kind == 3);
      const int E = nslots * C;
      const int C_next = pi.ncells;
      // small cross tables (the usual case) are counted in two packed registers per thread, 12 five-bit fields each (a thread owns at
      // most 24 rows and an entry sees a row at most once); larger ones in the shared-memory byte counters
      const bool packed = E <= kPipePacked;
      unsigned long long pk0 = 0ull, pk1 = 0ull;
      // the previous step's reduction tasks have read the bins and the count table
      // (which have also put them back to zero: every bin row and counter row is zero between steps)
      if (t > t_begin) named_bar_sync(1, kWorkers);
      // ---- A(t): per-slot sums of the residuals (after t-2, or after t-1 when drained), and the cross table (slot of t) x (cell of t-1);
      //      the leaf values of the current tree are added by the controller (count x value); rows of a proposed slot come from
      //      several current leaves, so their sums carry the leaf values themselves ----
      uint32_t ccur[NQ];
#pragma unroll
      for (int j = 0; j < NQ; ++j) ccur[j] = 0u;
      if (!two_trees) {
#pragma unroll
        for (int j = 0; j < NQ; ++j) if (j < nq) {
          int s[4], e[4];
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            s[o] = (sp[j] >> (8 * o)) & 0xFF;
            const int cp = (cprev[j] >> (8 * o)) & 0xFF;
            e[o] = s[o] * C + cp;                                    // rows in the trash slot: beyond E
          }
#pragma unroll
          for (int o = 0; o < 4; ++o) bins[s[o] * kWorkers + tid] += R[j][o];
          if (packed) {
#pragma unroll
            for (int o = 0; o < 4; ++o) { const unsigned sh = (unsigned) e[o] * 5u; pk0 += shl64_clamp(sh); pk1 += shl64_clamp(sh - 60u); }
          } else {
#pragma unroll
            for (int o = 0; o < 4; ++o) cnt[(s[o] >= kPipeSlots ? e_trash : e[o]) * kWorkers + tid] += 1;
          }
          ccur[j] = sp[j];                                         // cells of this step = its slots
        }
      } else {
#pragma unroll
        for (int j = 0; j < NQ; ++j) if (j < nq) {
          double pr[4]; int s[4], q[4], e[4], e2[4];
          uint32_t cc = 0u;
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            s[o] = (sp[j] >> (8 * o)) & 0xFF;
            const int pq = (pp[j] >> (8 * o)) & 0xFF;
            const int cp = (cprev[j] >> (8 * o)) & 0xFF;
            const bool in = pq != 255;
            pr[o] = R[j][o] + pi.vs[s[o]];
            q[o] = in ? pq : kPipeSlots;
            e[o] = s[o] * C + cp;
            e2[o] = in ? pq * C + cp : 0xFFFF;
            const int cell = s[o] >= kPipeSlots ? kPipeSlots : (int) pi.cellbase[s[o]] + (in ? pq - L : 0);
            cc |= (uint32_t) cell << (8 * o);
          }
#pragma unroll
          for (int o = 0; o < 4; ++o) { bins[s[o] * kWorkers + tid] += R[j][o]; bins[q[o] * kWorkers + tid] += pr[o]; }
          if (packed) {
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              const unsigned sh = (unsigned) e[o] * 5u, sh2 = (unsigned) e2[o] * 5u;
              pk0 += shl64_clamp(sh) + shl64_clamp(sh2); pk1 += shl64_clamp(sh - 60u) + shl64_clamp(sh2 - 60u);
            }
          } else {
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              cnt[(s[o] >= kPipeSlots ? e_trash : e[o]) * kWorkers + tid] += 1;
              cnt[(e2[o] == 0xFFFF ? e_trash : e2[o]) * kWorkers + tid] += 1;
            }
          }
          ccur[j] = cc;
        }
      }
      if (packed) { pkw[tid] = pk0; pkw[kWorkers + tid] = pk1; }
#pragma unroll
      for (int j = 0; j < NQ; ++j) { cprev2[j] = cprev[j]; cprev[j] = ccur[j]; }
      const long long k4 = PROF ? clock64() : 0;
      if (warp == kWorkerWarps - 1) cp_async_wait_all();          // the next step's descriptor has landed (issued a whole step ago)
      named_bar_sync(1, kWorkers);
      // ---- CTA reduction, then integer atomics into the step's accumulator row ----
      unsigned long long* acc = acc_ring + (size_t) ((t - t_begin) & (kPipeRing - 1)) * kPipeAcc;
      const int npairs = (E + 1) >> 1;
      for (int task = warp; task < nslots + npairs; task += kWorkerWarps) {
        if (task < nslots) {
          double a = 0.0;
#pragma unroll
          for (int i = 0; i < kWorkerWarps; ++i) { a += bins[task * kWorkers + i * 32 + lane]; bins[task * kWorkers + i * 32 + lane] = 0.0; }
          a = w_sum(a);
          if (lane == 0) {
            // two-limb fixed point: hi = floor(a 2^20) (biased by 2^47 to stay non-negative), lo = (a 2^20 - hi) 2^48; bits 56..63 count
            // the contributions.  |a| < 2^27 per CTA; 148 CTAs x 2^48 stay below the count field.
            const double xs = a * 1048576.0, fl = floor(xs);
            const unsigned long long hi = (unsigned long long) ((long long) fl + (1ll << 47)), lo = (unsigned long long) ((xs - fl) * 281474976710656.0);
            atomicAdd(acc + 2 * task, (1ull << 56) | hi);
            atomicAdd(acc + 2 * task + 1, (1ull << 56) | lo);
          }
        } else if (packed) {
          // two cross-table entries per task (both in the same packed word: 12 fields per word)
          const int e0 = 2 * (task - nslots);
          const unsigned long long* wsrc = pkw + (e0 >= 12 ? kWorkers : 0);
          const int f0 = 5 * (e0 >= 12 ? e0 - 12 : e0);
          int c0 = 0, c1 = 0;
#pragma unroll
          for (int i = 0; i < kWorkerWarps; ++i) { const unsigned long long w = wsrc[i * 32 + lane] >> f0; c0 += (int) (w & 31ull); c1 += (int) ((w >> 5) & 31ull); }
          c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1);
          if (e0 + 1 >= E) c1 = 0;
          if (lane == 0) atomicAdd(acc + 2 * kPipeSlots + (task - nslots), (1ull << 56) | (unsigned long long) (unsigned) c0 | ((unsigned long long) (unsigned) c1 << 28));
        } else {
          // two cross-table entries per task: 480 byte counters each, four at a time with dp4a
          const int e0 = 2 * (task - nslots), e1 = e0 + 1;
          int c0 = 0, c1 = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int w = i * 32 + lane;
            if (w < kWorkers / 4) {
              c0 = __dp4a(cntw[e0 * (kWorkers / 4) + w], 0x01010101u, (unsigned) c0); cntw[e0 * (kWorkers / 4) + w] = 0u;
              if (e1 < E) { c1 = __dp4a(cntw[e1 * (kWorkers / 4) + w], 0x01010101u, (unsigned) c1); cntw[e1 * (kWorkers / 4) + w] = 0u; }
            }
          }
          c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1);
          if (lane == 0) atomicAdd(acc + 2 * kPipeSlots + (task - nslots), (1ull << 56) | (unsigned long long) (unsigned) c0 | ((unsigned long long) (unsigned) c1 << 28));
        }
      }
      // (the bins, the packed words and the count table are rewritten by the next step only after the next named barrier)
      C = C_next;
      if (wprof) { const long long k5 = clock64(); wp0 += kw; wp1 += k2 - k0 - kw; wp2 += k3 - k2; wp3 += k4 - k3; wp4 += k5 - k4; wp6 += drained ? 1 : 0; }
    }
    if (wprof) { prof[0] += (unsigned long long) wp0; prof[1] += (unsigned long long) wp1; prof[2] += (unsigned long long) wp2; prof[3] += (unsigned long long) wp3; prof[4] += (unsigned long long) wp4; prof[5] += (unsigned long long) (t_end - t_begin); prof[6] += (unsigned long long) wp6; }
    // ---- drain: the updates not yet applied ----
    for (; na < t_end; ++na) {
      named_bar_sync(2 + (na & 1), kSweepBlock);
      const double* dc = S.dcell[na & 1];
      const bool newest = na == t_end - 1;
#pragma unroll
      for (int j = 0; j < NQ; ++j) if (j < nq) {
        const uint32_t cw = newest ? cprev[j] : cprev2[j];
#pragma unroll
        for (int o = 0; o < 4; ++o) R[j][o] += dc[(cw >> (8 * o)) & 0xFF];
      }
    }
#pragma unroll
    for (int j = 0; j < NQ; ++j) if ((valid_mask >> j) & 1u) {
      const long long q = q_lo + (long long) j * kWorkers + tid;
      *reinterpret_cast<double2*>(dv.R + 4 * q) = make_double2(R[j][0], R[j][1]);
      *reinterpret_cast<double2*>(dv.R + 4 * q + 2) = make_double2(R[j][2], R[j][3]);
    }
  } else {
    // =====================================================================================  controller warp
    WarpRng rngd; rngd.g = &S.rng; rngd.cs = &S.csd; rngd.lane = lane; rngd.writer = cta == 0;
    int C = 1;                                 // cells of step u - 1 (kept here: the workers recycle that step's descriptor slot while this step is decided)
    // cycle counters (lane 0 of CTA 0): [8] plan + tree fetch + draws, [9] wait for the rows, [10] row reduction + correction, [11] decision, [12] deltas + arrive
    long long cp0 = 0, cp1 = 0, cp2 = 0, cp3 = 0, cp4 = 0;
    const bool cprof = PROF && prof != nullptr && cta == 0 && lane == 0;
    for (int u = t_begin; u < t_end; ++u) {
      const long long h0 = PROF ? clock64() : 0;
      StepDesc& sd = S.sd[u % kPipeDescs];
      const PipeInfo& pi = S.info[u % kPipeDescs];
      DTree& tree = S.tree[u & 1];
      // ---- before the barrier: plan the decision, fetch the next tree, adopt this step's pre-computed draws ----
      { const FastPlan pl = w_plan(tree, sd, S.upd[u & 1], S.csd, lane); plan_store(S.plan, pl, lane); }
      // the next step's tree (a tree fits 32 nodes here: 8 + 32 x 24 bytes) and decision draws come in asynchronously: needed one step from now
      if (u + 1 < t_end) {
        const char* g = reinterpret_cast<const char*>(&dv.trees[u + 1]);
        char* tn = reinterpret_cast<char*>(&S.tree[(u + 1) & 1]);
        for (int i = lane; i < (int) ((8 + 32 * sizeof(DNode)) / 8); i += 32) cp_async8(tn + 8 * i, g + 8 * i);
        cp_async8(reinterpret_cast<char*>(&S.draws[(u + 1) & 1][lane]), reinterpret_cast<const char*>(draws + (u + 1) * 32 + lane));
        cp_async8(reinterpret_cast<char*>(&S.draws[(u + 1) & 1][lane]) + 8, reinterpret_cast<const char*>(draws + (u + 1) * 32 + lane) + 8);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      rngd.enter(step0 + (unsigned long long) u, 1u);
      { const double2 dz = S.draws[u & 1][lane]; S.csd.ubuf[lane] = dz.x; S.csd.zbuf[lane] = dz.y; }
      rngd.adopt();
      const long long h1 = PROF ? clock64() : 0;
      // (no separate barrier: the accumulator words themselves tell when every CTA has contributed, see below)
      __syncwarp();
      const long long h2 = PROF ? clock64() : 0;
      // ---- reduce the partial rows of all CTAs (fixed order), correct the sums with the previous step's deltas ----
      const int nslots = sd.b_nslots;
      // a pair of steps whose cross table exceeds the counters: the workers ran this step drained (update u-1 applied first), one cell
      const bool drained = u > t_begin && nslots * C > count_entries;
      if (drained) C = 1;
      const int npairs = (nslots * C + 1) >> 1;
      {
        // One load per lane and 32 values: this step's totals = accumulator row now - the row as read kPipeRing steps ago; a word
        // is complete when its top byte has advanced by the number of CTAs.  (Shared-memory data written by this CTA's workers
        // before they issued their atomics -- the descriptor ring -- is read only after this loop has seen those atomics.)
        const int slot_b = (u - t_begin) & (kPipeRing - 1);
        const unsigned long long* acc = acc_ring + (size_t) slot_b * kPipeAcc;
        const int nvals = 2 * kPipeSlots + npairs;
        const unsigned long long want = (unsigned long long) (G & 0xFF);
        for (int i0 = 0; i0 < nvals; i0 += 32) {
          const int i = i0 + lane;
          const bool used = i < nvals && (i >= 2 * kPipeSlots || i < 2 * nslots);
          unsigned long long v = 0ull;
          const unsigned long long prev = used ? S.accprev[slot_b][i] : 0ull;
          const long long w0 = clock64();
          bool done = !used;
          for (;;) {
            if (!done) {
              unsigned long long now;
              asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(now) : "l"(acc + i) : "memory");
              v = now - prev;
              if ((v >> 56) == want) { done = true; S.accprev[slot_b][i] = now; }
            }
            if (__all_sync(0xffffffffu, done)) break;
            if (S.fail || clock64() - w0 > 4000000000LL) { S.fail = 1; break; }
          }
          v &= (1ull << 56) - 1ull;
          if (i < 2 * kPipeSlots) reinterpret_cast<unsigned long long*>(S.ncnt)[i] = v;       // staged: (hi, lo) pairs, combined below
          else if (i < nvals) {
            const int k = i - 2 * kPipeSlots;
            S.cross[2 * k] = (int) (v & 0xFFFFFFFull); S.cross[2 * k + 1] = (int) ((v >> 28) & 0xFFFFFFFull);
          }
        }
        __syncwarp();
        if (lane < nslots) {
          const long long hi = (long long) reinterpret_cast<unsigned long long*>(S.ncnt)[2 * lane] - (long long) G * (1ll << 47);
          const long long lo = (long long) reinterpret_cast<unsigned long long*>(S.ncnt)[2 * lane + 1];
          S.st[lane].sum = (double) hi * 9.5367431640625e-07 + (double) lo * 3.3881317890172014e-21;       // 2^-20, 2^-68
        }
      }
      __syncwarp();
      if (lane < nslots) {
        const double* dprev = S.dcell[(u + 1) & 1];          // deltas of step u-1 (zeros before the first step; already applied when drained)
        int cnt_s = 0; double corr = 0.0;
        for (int cidx = 0; cidx < C; ++cidx) { const int m = S.cross[lane * C + cidx]; cnt_s += m; if (m != 0 && !drained) corr += (double) m * dprev[cidx]; }   // (an empty cell's delta is undefined)
        // the workers summed the plain residuals of the current slots: the partial residual adds the slot's current leaf value per row
        // (the rows of a proposed slot of a change / swap step sit in several current leaves: those sums carry the values already)
        const bool two = sd.b_kind == 2 || sd.b_kind == 3;
        if (!two || lane < sd.b_num_leaves) corr += (double) cnt_s * pi.vs[lane];
        S.st[lane].n = (double) cnt_s;
        S.st[lane].sum += corr;
      }
      __syncwarp();
      const long long h3 = PROF ? clock64() : 0;
      // ---- Metropolis decision + leaf draws (same code as the synchronous kernel) ----
      {
        const FastPlan plan = plan_load(S.plan, lane);
        w_decide_fast<false>(plan, tree, S.prm, rngd, sd, S.st, S.upd[u & 1], S.csd, nullptr, lane, S.inv_sigsq, sd.accept_thr);
      }
      rngd.commit();
      const long long h4 = PROF ? clock64() : 0;
      // ---- per-cell deltas of this step: what the workers add to the residuals, and the next step's correction ----
      {
        const UpdateDesc& upd = S.upd[u & 1];
        const bool accepted = upd.mode != 0;
        for (int c = lane; c < kPipeCells; c += 32) {
          double dlt = 0.0;
          if (c < pi.ncells) { const int a = pi.cell_a[c]; const int f = accepted ? (int) pi.cell_f[c] : a; dlt = upd.val_old[a] - upd.val_new[f]; }
          S.dcell[u & 1][c] = dlt;
        }
      }
      __syncwarp();
      C = pi.ncells;
      named_bar_arrive(2 + (u & 1), kSweepBlock);            // decision u is done: the workers may apply it
      cp_async_wait_all(); __syncwarp();                     // next tree and draws are in shared memory
      if (cprof) { const long long h5 = clock64(); cp0 += h1 - h0; cp1 += h2 - h1; cp2 += h3 - h2; cp3 += h4 - h3; cp4 += h5 - h4; }
      if (cta == 0) {
        DTree& g = dv.trees[u];
        const int nn = tree.num_nodes;
        if (lane == 0) g.num_nodes = nn;
        for (int i = lane; i < nn * (int) (sizeof(DNode) / 4); i += 32) reinterpret_cast<uint32_t*>(g.nodes)[i] = reinterpret_cast<const uint32_t*>(tree.nodes)[i];
      }
      __syncwarp();
    }
    if (cprof) { prof[8] += (unsigned long long) cp0; prof[9] += (unsigned long long) cp1; prof[10] += (unsigned long long) cp2; prof[11] += (unsigned long long) cp3; prof[12] += (unsigned long long) cp4; }
    if (cta == 0 && lane == 0) {
      RngState out = S.rng;
      out.counter += (unsigned long long) S.csd.draws_total;
      *dv.rng = out;
      if (S.fail) dv.params->error_flag |= 8u;
      if (t_end == T) dv.params->step_id = step0 + (unsigned long long) T;
      dv.desc->a_valid = 0;
      *pos_out = t_end;
      *ran += (unsigned long long) (t_end - t_begin);
    }
  }
}

}  // namespace s4b
