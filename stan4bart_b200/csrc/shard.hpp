// stan4bart_b200/csrc/shard.hpp
// Observation-sharded mode (BASELINE config E): rows are split over the GPUs of one NVSwitch box, trees /
// RNG / NUTS state are replicated, and the only data exchanged are the tiny per-node sufficient statistics of every tree
// step and the (1 + K + q) GLMM reductions.  The exchange does not go through NCCL: every rank owns a "mailbox" in its
// HBM that all peers map through CUDA IPC; a rank stores its contribution directly into every peer's mailbox over NVLink
// (one-shot all-gather), then each rank sums the contributions in rank order, so all ranks obtain bitwise identical
// results and take identical Metropolis decisions.  The per-tree exchange happens INSIDE the persistent sweep kernel.
#pragma once

#include "s4b_common.cuh"

#include <stdexcept>

namespace s4b {

constexpr int kMaxRanks = 8;
constexpr int kMailVec = 1024;     // capacity of the generic small all-reduce

struct Mailbox {
  // per-tree-step statistics, double buffered by step parity.  Low-latency layout: every double travels as two 8-byte
  // words (flag32 << 32 | half32) so that data and flag arrive in the same store -- no fence and no second hop between
  // "data written" and "flag visible"; the flag is the low 32 bits of the step's sequence number
  ulonglong2 step_ll[2][kMaxRanks][3 * S4B_MAX_SLOTS];
  // number of per-tree-step exchanges this rank has completed: the next exchange uses kseq + 1.  Kept on the device so that
  // graph-captured launches need no per-launch argument; identical on all ranks because they make identical call sequences
  unsigned long long kseq;
  // generic small vectors (GLMM reductions, min / max for the rescale, cut-point ranges)
  unsigned long long vec_flag[2][kMaxRanks];
  double vec_data[2][kMaxRanks][kMailVec];
};

struct ShardDev {
  int rank, world;
  long long obs_offset;            // global index of this rank's first observation (keys the latent draws)
  Mailbox* mail[kMaxRanks];        // mail[r] = rank r's mailbox as mapped into this process (mail[rank] is local)
};

enum ReduceOp { kOpSum = 0, kOpMax = 1 };

class ShardContext {
 public:
  ShardContext(int rank, int world);
  ~ShardContext();
  ShardContext(const ShardContext&) = delete;
  ShardContext& operator=(const ShardContext&) = delete;
  void ipc_handle(void* out64) const;                       // cudaIpcMemHandle_t of the local mailbox
  void attach(const void* handles64_by_rank);               // world handles, 64 bytes each
  bool attached() const { return attached_; }
  int rank() const { return dev_.rank; }
  int world() const { return dev_.world; }
  // this rank's rows are [first_obs, first_obs + n_local) of total_obs
  void set_obs_range(long long first_obs, long long total_obs) { dev_.obs_offset = first_obs; total_obs_ = total_obs; }
  long long obs_offset() const { return dev_.obs_offset; }
  long long total_obs() const { return total_obs_; }
  void check_error();
  const ShardDev& dev() const { return dev_; }
  // in-place all-reduce of a small device vector (n <= kMailVec), identical result on every rank
  void allreduce(double* d_vec, int n, ReduceOp op, cudaStream_t stream);
  // Reference collective: the same small all-reduces through ncclAllReduce over NVLink instead of the
  // peer-mapped mailboxes.  NCCL is loaded at run time (libnccl.so.2); every rank calls nccl_init with the 128-byte unique id
  // that rank 0 obtained from nccl_unique_id (hand it round with any transport).  use_nccl(true) routes allreduce() through it.
  static void nccl_unique_id(void* out128);
  void nccl_init(const void* id128);
  void use_nccl(bool on);
  bool using_nccl() const { return use_nccl_; }
  // host convenience (any length; chunks of kMailVec)
  void allreduce_host(double* h_vec, long long n, ReduceOp op, cudaStream_t stream);

 private:
  ShardDev dev_;
  Mailbox* local_ = nullptr;
  bool attached_ = false;
  unsigned long long vec_seq_ = 0;
  long long total_obs_ = 0;
  double* d_tmp_ = nullptr;
  unsigned int* d_err_ = nullptr;
  void* nccl_comm_ = nullptr;
  bool use_nccl_ = false;
};

// device side of the generic all-reduce, callable from any single CTA
__device__ inline unsigned long long global_timer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until the flag reaches `seq`; gives up after kPeerTimeoutNs (a peer died or never launched) so that a broken
// job fails with an error instead of hanging the GPU
constexpr unsigned long long kPeerTimeoutNs = 30ull * 1000ull * 1000ull * 1000ull;
__device__ inline bool mailbox_wait(const unsigned long long* flag, unsigned long long seq)
{
  unsigned long long v;
  unsigned long long t0 = 0;
  unsigned int spins = 0;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= seq) return true;
    if ((++spins & 1023u) == 0u) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kPeerTimeoutNs) return false;
    }
  }
}
// flag-in-word transport of one double (see Mailbox::step_ll)
__device__ inline void mailbox_send_ll(ulonglong2* slot, double v, unsigned int seq32)
{
  const unsigned long long bits = (unsigned long long) __double_as_longlong(v);
  const unsigned long long f = (unsigned long long) seq32 << 32;
  const unsigned long long lo = f | (bits & 0xFFFFFFFFull), hi = f | (bits >> 32);
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(lo), "l"(hi) : "memory");
}
__device__ inline bool mailbox_recv_ll(const ulonglong2* slot, unsigned int seq32, double* out)
{
  unsigned long long lo, hi;
  unsigned long long t0 = 0;
  unsigned int spins = 0;
  for (;;) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(slot) : "memory");
    if ((unsigned int) (lo >> 32) == seq32 && (unsigned int) (hi >> 32) == seq32) break;
    if ((++spins & 1023u) == 0u) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kPeerTimeoutNs) return false;
    }
  }
  *out = __longlong_as_double((long long) ((hi << 32) | (lo & 0xFFFFFFFFull)));
  return true;
}
__device__ inline void mailbox_post(unsigned long long* flag, unsigned long long seq)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
}
__device__ inline double mailbox_load(const double* p)
{
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

}  // namespace s4b
