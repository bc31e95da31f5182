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
// Rings: partial rows and barrier counters are 4 deep (a CTA can run at most two steps ahead of another CTA's controller
// reading rows); descriptors 3 deep; update tables 2 deep.  Decisions, draws and tree updates are the synchronous kernel's
// (w_plan / w_decide_fast), so both kernels produce the same chain up to the rounding of the slot sums.
#pragma once

#include "sweep_kernel.cuh"

namespace s4b {

struct PipeSmem {
  StepDesc sd[kPipeDescs];
  PipeInfo info[kPipeDescs];
  DTree tree[2];
  UpdateDesc upd[2];
  double dcell[2][kPipeCells];              // delta of step parity: mu_old - mu_new per cell
  int ncnt[kBinSlots * kPipeCells];         // the reduced cross table of the step being decided
  CtlScratch csd;
  LeafStat st[S4B_MAX_SLOTS];
  FastPlanSmem plan;
  double inv_sigsq;
  int fail;
  RngState rng;
  BartParams prm;
};

__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ inline void w_copy_info(PipeInfo& dst, const PipeInfo& src, int lane)
{
  if (lane == 0) { dst.ncells = src.ncells; dst.ok = src.ok; }
  dst.cellbase[lane] = src.cellbase[lane];
  if (lane < kPipeCells) { dst.cell_a[lane] = src.cell_a[lane]; dst.cell_f[lane] = src.cell_f[lane]; }
  __syncwarp();
}

// partial rows of one step in the ring: [0, kBinSlots) per-slot sums (double), then `count_words` rows of packed counts
// (4 x 16 bit per CTA: a CTA holds at most 480 x 24 rows); row r of CTA c at ring + r * G + c
template <int NQ>
__global__ void __launch_bounds__(kSweepBlock, 1) k_sweep_pipe(BartDev dv, unsigned int* counters, double* ring, int ring_stride,
                                                                const StepDesc* __restrict__ descs, const PipeInfo* __restrict__ infos,
                                                                const double2* __restrict__ draws, const unsigned int* __restrict__ not_ok, int count_words,
                                                                unsigned long long* __restrict__ ran)
{
  if (*not_ok != 0u) return;               // some tree of this sweep does not fit: the synchronous kernel (launched next) runs instead
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PipeSmem& S = *reinterpret_cast<PipeSmem*>(smem_raw);
  double* bins = reinterpret_cast<double*>(smem_raw + ((sizeof(PipeSmem) + 15) / 16) * 16);                    // [kBinSlots + 1][kWorkers]
  uint32_t* cnt = reinterpret_cast<uint32_t*>(bins + (kBinSlots + 1) * kWorkers);                               // [count_words + 1][kWorkers]
  uint32_t* tile = cnt + (size_t) (count_words + 1) * kWorkers;                                                   // [p][NQ * kWorkers]
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
    const DTree& g = dv.trees[0];
    const int nn = g.num_nodes;
    if (lane == 0) { S.tree[0].num_nodes = nn; S.tree[0].pad = 0; }
    for (int i = lane; i < nn * (int) (sizeof(DNode) / 4); i += 32) reinterpret_cast<uint32_t*>(S.tree[0].nodes)[i] = reinterpret_cast<const uint32_t*>(g.nodes)[i];
    w_copy_desc(S.sd[0], descs[0], lane);
    w_copy_info(S.info[0], infos[0], lane);
  }
  __syncthreads();

  if (is_worker) {
    // =====================================================================================  workers
    uint32_t leaf_pack[NQ], aux_pack[NQ], cprev[NQ], cprev2[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) { cprev[j] = 0u; cprev2[j] = 0u; }
    for (int t = 0; t < T; ++t) {
      const StepDesc& sd = S.sd[t % kPipeDescs];
      const PipeInfo& pi = S.info[t % kPipeDescs];
      if (t >= 2) {
        // ---- U(t-2): wait for decision t-2, add its per-cell deltas ----
        named_bar_sync(2 + (t & 1), kSweepBlock);
        const double* dc = S.dcell[t & 1];
#pragma unroll
        for (int j = 0; j < NQ; ++j)
#pragma unroll
          for (int o = 0; o < 4; ++o) R[j][o] += dc[(cprev2[j] >> (8 * o)) & 0xFF];
      }
      // one warp fetches the next step's descriptor (its ring slot held step t-2, which has been decided)
      if (warp == kWorkerWarps - 1 && t + 1 < T) { w_copy_desc(S.sd[(t + 1) % kPipeDescs], descs[t + 1], lane); w_copy_info(S.info[(t + 1) % kPipeDescs], infos[t + 1], lane); }
      // ---- W(t) ----
      walk_step<NQ>(sd, tile, tile_stride, tid, valid_mask, leaf_pack, aux_pack);
      const int kind = sd.b_kind, L = sd.b_num_leaves, nslots = sd.b_nslots;
      const bool two_trees = (kind == 2 || kind == 3);
      const int birth_node = kind == 0 ? sd.b_node : -1;
      const int C = t > 0 ? S.info[(t - 1) % kPipeDescs].ncells : 1;
      const int nwords = (nslots * C + 3) >> 2;
      for (int k = 0; k < nslots; ++k) bins[k * kWorkers + tid] = 0.0;
      for (int k = 0; k < nwords; ++k) cnt[k * kWorkers + tid] = 0u;
      // ---- A(t): per-slot sums of (residual after t-2) + mu_t, and the cross table (slot of t) x (cell of t-1) ----
      uint32_t ccur[NQ];
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        double pr[4]; int row[4], row2[4], ent[4], ent2[4];
        uint32_t cc = 0u;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int leaf = (leaf_pack[j] >> (8 * o)) & 0xFF;
          const int aux = (aux_pack[j] >> (8 * o)) & 0xFF;
          const int cp = (cprev[j] >> (8 * o)) & 0xFF;
          const bool ok = (obs_mask >> (4 * j + o)) & 1u;
          pr[o] = R[j][o] + sd.b_cur.val[leaf];
          int sa, sb = 255, cell;
          if (two_trees) {
            sa = sd.b_cur.slot[leaf]; sb = sd.b_prop.slot[aux];
            cell = (int) pi.cellbase[leaf] + (sb != 255 ? sb - L : 0);
          } else {
            sa = leaf == birth_node ? L + aux : (int) sd.b_cur.slot[leaf];
            cell = sa;
          }
          cc |= (uint32_t) (ok ? cell : 0) << (8 * o);
          row[o] = ok ? sa : kBinSlots;
          ent[o] = ok ? sa * C + cp : 4 * count_words;
          const bool second = two_trees && sb != 255 && ok;
          row2[o] = second ? sb : kBinSlots;
          ent2[o] = second ? sb * C + cp : 4 * count_words;
        }
        ccur[j] = cc;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          bins[row[o] * kWorkers + tid] += pr[o];
          cnt[(ent[o] >> 2) * kWorkers + tid] += 1u << (8 * (ent[o] & 3));
          if (two_trees) {
            bins[row2[o] * kWorkers + tid] += pr[o];
            cnt[(ent2[o] >> 2) * kWorkers + tid] += 1u << (8 * (ent2[o] & 3));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NQ; ++j) { cprev2[j] = cprev[j]; cprev[j] = ccur[j]; }
      named_bar_sync(1, kWorkers);
      // ---- CTA reduction: one partial row per slot / count word ----
      double* rows = ring + (size_t) (t & (kPipeRing - 1)) * ring_stride;
      for (int task = warp; task < nslots + nwords; task += kWorkerWarps) {
        if (task < nslots) {
          double a = 0.0;
#pragma unroll
          for (int i = 0; i < kWorkerWarps; ++i) a += bins[task * kWorkers + i * 32 + lane];
          a = w_sum(a);
          if (lane == 0) rows[(size_t) task * G + cta] = a;
        } else {
          const int w = task - nslots;
          int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
          for (int i = 0; i < kWorkerWarps; ++i) {
            const uint32_t v = cnt[w * kWorkers + i * 32 + lane];
            c0 += (int) (v & 0xFFu); c1 += (int) ((v >> 8) & 0xFFu); c2 += (int) ((v >> 16) & 0xFFu); c3 += (int) (v >> 24);
          }
          c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1);
          c2 = __reduce_add_sync(0xffffffffu, c2); c3 = __reduce_add_sync(0xffffffffu, c3);
          if (lane == 0) {
            const unsigned long long packed = (unsigned long long) c0 | ((unsigned long long) c1 << 16) | ((unsigned long long) c2 << 32) | ((unsigned long long) c3 << 48);
            reinterpret_cast<unsigned long long*>(rows)[(size_t) (kBinSlots + w) * G + cta] = packed;
          }
        }
      }
      // every row of this CTA is stored before thread 0 publishes them; nobody waits here
      named_bar_sync(1, kWorkers);
      if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counters + (t & (kPipeRing - 1))) : "memory");
    }
    // ---- drain: the last two updates ----
    for (int u = T - 2; u < T; ++u) {
      if (u < 0) continue;
      named_bar_sync(2 + (u & 1), kSweepBlock);
      const double* dc = S.dcell[u & 1];
#pragma unroll
      for (int j = 0; j < NQ; ++j)
#pragma unroll
        for (int o = 0; o < 4; ++o) R[j][o] += dc[((u == T - 1 ? cprev[j] : cprev2[j]) >> (8 * o)) & 0xFF];
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
    for (int u = 0; u < T; ++u) {
      StepDesc& sd = S.sd[u % kPipeDescs];
      const PipeInfo& pi = S.info[u % kPipeDescs];
      DTree& tree = S.tree[u & 1];
      // ---- before the barrier: plan the decision, fetch the next tree, adopt this step's pre-computed draws ----
      { const FastPlan pl = w_plan(tree, sd, S.upd[u & 1], S.csd, lane); plan_store(S.plan, pl, lane); }
      if (u + 1 < T) {
        DTree& tn = S.tree[(u + 1) & 1];
        const DTree& g = dv.trees[u + 1];
        const int nn = g.num_nodes;
        if (lane == 0) { tn.num_nodes = nn; tn.pad = 0; }
        for (int i = lane; i < nn * (int) (sizeof(DNode) / 4); i += 32) reinterpret_cast<uint32_t*>(tn.nodes)[i] = reinterpret_cast<const uint32_t*>(g.nodes)[i];
      }
      rngd.enter(step0 + (unsigned long long) u, 1u);
      { const double2 dz = __ldcg(draws + u * 32 + lane); S.csd.ubuf[lane] = dz.x; S.csd.zbuf[lane] = dz.y; }
      rngd.adopt();
      // ---- wait until every CTA has published its rows of step u ----
      if (lane == 0) {
        const unsigned int target = (unsigned int) (u / kPipeRing + 1) * (unsigned int) G;
        const unsigned int* ctr = counters + (u & (kPipeRing - 1));
        unsigned int v;
        const long long w0 = clock64();
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
          if (v < target && (S.fail || clock64() - w0 > 4000000000LL)) { S.fail = 1; break; }
        } while (v < target);
      }
      __syncwarp();
      // ---- reduce the partial rows of all CTAs (fixed order), correct the sums with the previous step's deltas ----
      const int nslots = sd.b_nslots;
      const int C = u > 0 ? S.info[(u - 1) % kPipeDescs].ncells : 1;
      const int nwords = (nslots * C + 3) >> 2;
      const double* rows = ring + (size_t) (u & (kPipeRing - 1)) * ring_stride;
      for (int r0 = 0; r0 < nslots; r0 += 4) {
        double acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double* src = rows + (size_t) (r0 + k) * G;
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
          if (r0 + k < nslots) {
            { int b = lane;       if (b < G) a0 = __ldcg(src + b); }
            { int b = lane + 32;  if (b < G) a1 = __ldcg(src + b); }
            { int b = lane + 64;  if (b < G) a2 = __ldcg(src + b); }
            { int b = lane + 96;  if (b < G) a3 = __ldcg(src + b); }
            { int b = lane + 128; if (b < G) a4 = __ldcg(src + b); }
            for (int b = lane + 160; b < G; b += 32) a4 += __ldcg(src + b);
          }
          acc[k] = (((a0 + a1) + a2) + a3) + a4;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { const double a = w_sum(acc[k]); if (lane == 0 && r0 + k < nslots) S.st[r0 + k].sum = a; }
      }
      for (int w0 = 0; w0 < nwords; w0 += 4) {
        int c[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned long long* src = reinterpret_cast<const unsigned long long*>(rows) + (size_t) (kBinSlots + w0 + k) * G;
          c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0;
          if (w0 + k < nwords)
            for (int b = lane; b < G; b += 32) {
              const unsigned long long v = __ldcg(src + b);
              c[k][0] += (int) (v & 0xFFFFull); c[k][1] += (int) ((v >> 16) & 0xFFFFull); c[k][2] += (int) ((v >> 32) & 0xFFFFull); c[k][3] += (int) (v >> 48);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const int tot = __reduce_add_sync(0xffffffffu, c[k][f]);
            if (lane == 0 && w0 + k < nwords) S.ncnt[4 * (w0 + k) + f] = tot;
          }
      }
      __syncwarp();
      if (lane < nslots) {
        const double* dprev = S.dcell[(u + 1) & 1];          // deltas of step u-1 (zeros before the first step)
        int cnt_s = 0; double corr = 0.0;
        for (int cidx = 0; cidx < C; ++cidx) { const int m = S.ncnt[lane * C + cidx]; cnt_s += m; if (m != 0) corr += (double) m * dprev[cidx]; }   // (an empty cell's delta is undefined)
        S.st[lane].n = (double) cnt_s;
        S.st[lane].sum += corr;
      }
      __syncwarp();
      // ---- Metropolis decision + leaf draws (same code as the synchronous kernel) ----
      {
        const FastPlan plan = plan_load(S.plan, lane);
        w_decide_fast<false>(plan, tree, S.prm, rngd, sd, S.st, S.upd[u & 1], S.csd, nullptr, lane, S.inv_sigsq, sd.accept_thr);
      }
      rngd.commit();
      // ---- per-cell deltas of this step: what the workers add to the residuals, and the next step's correction ----
      {
        const UpdateDesc& upd = S.upd[u & 1];
        const bool accepted = upd.mode != 0;
        if (lane < kPipeCells) {
          double dlt = 0.0;
          if (lane < pi.ncells) { const int a = pi.cell_a[lane]; const int f = accepted ? (int) pi.cell_f[lane] : a; dlt = upd.val_old[a] - upd.val_new[f]; }
          S.dcell[u & 1][lane] = dlt;
        }
      }
      __syncwarp();
      named_bar_arrive(2 + (u & 1), kSweepBlock);            // decision u is done: the workers may apply it
      if (cta == 0) {
        DTree& g = dv.trees[u];
        const int nn = tree.num_nodes;
        if (lane == 0) g.num_nodes = nn;
        for (int i = lane; i < nn * (int) (sizeof(DNode) / 4); i += 32) reinterpret_cast<uint32_t*>(g.nodes)[i] = reinterpret_cast<const uint32_t*>(tree.nodes)[i];
      }
      __syncwarp();
    }
    if (cta == 0 && lane == 0) {
      RngState out = S.rng;
      out.counter += (unsigned long long) S.csd.draws_total;
      *dv.rng = out;
      if (S.fail) dv.params->error_flag |= 8u;
      dv.params->step_id = step0 + (unsigned long long) T;
      dv.desc->a_valid = 0;
      *ran += 1ull;
    }
  }
}

}  // namespace s4b
