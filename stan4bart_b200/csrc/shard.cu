// stan4bart_b200/csrc/shard.cu -- see shard.hpp
#include "shard.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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

// ---- NCCL, bound at run time: the library is the reference collective, not a link-time dependency of the product path ----
namespace {
struct NcclId { char internal[128]; };
struct NcclApi {
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi& nccl_api()
{
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (h != nullptr) {
      api.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(h, "ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclId, int)>(dlsym(h, "ncclCommInitRank"));
      api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(h, "ncclAllReduce"));
      api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
      api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
      api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
    }
  }
  if (!api.ok) throw std::runtime_error("NCCL (libnccl.so.2) could not be loaded");
  return api;
}
void nccl_check(int rc, const char* what) { if (rc != 0) throw std::runtime_error(std::string(what) + ": " + nccl_api().GetErrorString(rc)); }
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;       // nccl.h: ncclFloat64, ncclSum, ncclMax
}  // namespace

void ShardContext::nccl_unique_id(void* out128)
{
  NcclId id;
  nccl_check(nccl_api().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out128, &id, sizeof id);
}

void ShardContext::nccl_init(const void* id128)
{
  if (nccl_comm_ != nullptr) return;
  NcclId id; std::memcpy(&id, id128, sizeof id);
  nccl_check(nccl_api().CommInitRank(&nccl_comm_, dev_.world, id, dev_.rank), "ncclCommInitRank");
}

void ShardContext::use_nccl(bool on)
{
  if (on && nccl_comm_ == nullptr) throw std::runtime_error("shard context: nccl_init first");
  use_nccl_ = on;
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
  if (nccl_comm_ != nullptr) nccl_api().CommDestroy(nccl_comm_);
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
  if (dev_.world == 1 && !use_nccl_) return;         // (a one-rank NCCL communicator still goes through ncclAllReduce: the cross-check on one GPU)
  if (!attached_) throw std::runtime_error("shard context: peers not attached");
  if (n < 0 || n > kMailVec) throw std::invalid_argument("shard all-reduce: vector too long");
  if (use_nccl_) {        // the reference collective: same payload, same ranks, NCCL's ring / tree over NVLink
    nccl_check(nccl_api().AllReduce(d_vec, d_vec, (size_t) n, kNcclFloat64, op == kOpSum ? kNcclSum : kNcclMax, nccl_comm_, stream), "ncclAllReduce");
    return;
  }
  ++vec_seq_;
  static const bool debug = getenv("S4B_SHARD_DEBUG") != nullptr;
  if (debug) fprintf(stderr, "[s4b shard] rank %d allreduce seq %llu n %d op %d\n", dev_.rank, vec_seq_, n, (int) op);
  k_allreduce_small<<<1, 256, 0, stream>>>(dev_, d_vec, n, (int) op, vec_seq_, d_err_);
  S4B_CUDA(cudaGetLastError());
}

void ShardContext::allreduce_host(double* h_vec, long long n, ReduceOp op, cudaStream_t stream)
{
  if (dev_.world == 1 && !use_nccl_) return;
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
