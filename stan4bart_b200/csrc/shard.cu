// stan4bart_b200/csrc/shard.cu -- see shard.hpp
#include "shard.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace s4b {

// one CTA: publish my vector into every rank's mailbox, wait for all ranks, reduce in rank order
__global__ void __launch_bounds__(256) k_allreduce_small(ShardDev sh, double* __restrict__ vec, int n, int op, unsigned long long seq,
                                                         unsigned int* __restrict__ err)
{
  const int tid = threadIdx.x;
  const int par = (int) (seq & 1ull);
  for (int dst = 0; dst < sh.world; ++dst)
    for (int i = tid; i < n; i += blockDim.x) sh.mail[dst]->vec_data[par][sh.rank][i] = vec[i];
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    for (int dst = 0; dst < sh.world; ++dst) mailbox_post(&sh.mail[dst]->vec_flag[par][sh.rank], seq);
  }
  __shared__ int ok;
  if (tid == 0) ok = 1;
  __syncthreads();
  if (tid < sh.world) { if (!mailbox_wait(&sh.mail[sh.rank]->vec_flag[par][tid], seq)) ok = 0; }
  __syncthreads();
  if (!ok) { if (tid == 0) *err = 1u; return; }
  const Mailbox* mine = sh.mail[sh.rank];
  for (int i = tid; i < n; i += blockDim.x) {
    double acc = mailbox_load(&mine->vec_data[par][0][i]);
    for (int src = 1; src < sh.world; ++src) {
      double v = mailbox_load(&mine->vec_data[par][src][i]);
      acc = op == kOpSum ? acc + v : fmax(acc, v);
    }
    vec[i] = acc;
  }
}

ShardContext::ShardContext(int rank, int world)
{
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) throw std::invalid_argument("shard context: bad rank / world");
  std::memset(&dev_, 0, sizeof dev_);
  dev_.rank = rank; dev_.world = world; dev_.obs_offset = 0;
  S4B_CUDA(cudaMalloc(&local_, sizeof(Mailbox)));
  S4B_CUDA(cudaMemset(local_, 0, sizeof(Mailbox)));
  S4B_CUDA(cudaMalloc(&d_tmp_, sizeof(double) * kMailVec));
  S4B_CUDA(cudaMalloc(&d_err_, sizeof(unsigned int)));
  S4B_CUDA(cudaMemset(d_err_, 0, sizeof(unsigned int)));
  dev_.mail[rank] = local_;
  if (world == 1) attached_ = true;
  S4B_CUDA(cudaDeviceSynchronize());
}

ShardContext::~ShardContext()
{
  for (int r = 0; r < dev_.world; ++r) if (r != dev_.rank && dev_.mail[r] != nullptr) cudaIpcCloseMemHandle(dev_.mail[r]);
  cudaFree(local_); cudaFree(d_tmp_); cudaFree(d_err_);
}

void ShardContext::ipc_handle(void* out64) const
{
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
  cudaIpcMemHandle_t h;
  S4B_CUDA(cudaIpcGetMemHandle(&h, local_));
  std::memcpy(out64, &h, 64);
}

void ShardContext::attach(const void* handles64_by_rank)
{
  const unsigned char* p = static_cast<const unsigned char*>(handles64_by_rank);
  for (int r = 0; r < dev_.world; ++r) {
    if (r == dev_.rank) continue;
    cudaIpcMemHandle_t h; std::memcpy(&h, p + 64 * (size_t) r, 64);
    void* ptr = nullptr;
    S4B_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    dev_.mail[r] = static_cast<Mailbox*>(ptr);
  }
  attached_ = true;
}

void ShardContext::allreduce(double* d_vec, int n, ReduceOp op, cudaStream_t stream)
{
  if (dev_.world == 1) return;
  if (!attached_) throw std::runtime_error("shard context: peers not attached");
  if (n < 0 || n > kMailVec) throw std::invalid_argument("shard all-reduce: vector too long");
  ++vec_seq_;
  static const bool debug = getenv("S4B_SHARD_DEBUG") != nullptr;
  if (debug) fprintf(stderr, "[s4b shard] rank %d allreduce seq %llu n %d op %d\n", dev_.rank, vec_seq_, n, (int) op);
  k_allreduce_small<<<1, 256, 0, stream>>>(dev_, d_vec, n, (int) op, vec_seq_, d_err_);
  S4B_CUDA(cudaGetLastError());
}

void ShardContext::allreduce_host(double* h_vec, long long n, ReduceOp op, cudaStream_t stream)
{
  if (dev_.world == 1) return;
  for (long long off = 0; off < n; off += kMailVec) {
    int m = (int) std::min<long long>(kMailVec, n - off);
    S4B_CUDA(cudaMemcpyAsync(d_tmp_, h_vec + off, sizeof(double) * (size_t) m, cudaMemcpyHostToDevice, stream));
    allreduce(d_tmp_, m, op, stream);
    S4B_CUDA(cudaMemcpyAsync(h_vec + off, d_tmp_, sizeof(double) * (size_t) m, cudaMemcpyDeviceToHost, stream));
    S4B_CUDA(cudaStreamSynchronize(stream));
  }
  check_error();
}

void ShardContext::check_error()
{
  unsigned int e = 0;
  S4B_CUDA(cudaMemcpy(&e, d_err_, sizeof e, cudaMemcpyDeviceToHost));
  if (e) throw std::runtime_error("shard exchange timed out waiting for a peer");
}

}  // namespace s4b
