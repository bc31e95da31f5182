// Dependent-chain latencies of the fp64 operations on the sweep kernel's controller path (one warp, one SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_latency tools/ubench/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHAIN 64
template <int OP>
__global__ void k(double* out, long long* cyc, double seed, double c2)
{
  __shared__ double sm[64];
  sm[threadIdx.x] = seed + threadIdx.x;
  __syncwarp();
  double x = seed + 1e-3 * threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < CHAIN; ++i) {
    if (OP == 0) x = fma(x, c2, 0.25);
    else if (OP == 1) x = x + c2;
    else if (OP == 2) x = 1.0 / (x + c2);
    else if (OP == 3) x = sqrt(x + c2);
    else if (OP == 4) x = log(x + 3.0);
    else if (OP == 5) x = exp(x * 0.1);
    else if (OP == 6) x = sm[(threadIdx.x + (int) x) & 31] + c2;
    else if (OP == 7) x = __shfl_xor_sync(0xffffffffu, x, 1) + c2;
    else if (OP == 8) { __syncwarp(); x += c2; }
    else if (OP == 9) x = (double) ((float) x * 1.0001f);
    else if (OP == 10) x = log1p(x * 0.5);
    else if (OP == 11) x = x / (c2 + (double) i);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name)
{
  double* d; long long* c; cudaMalloc(&d, 256 * 8); cudaMalloc(&c, 8);
  k<OP><<<1, 32>>>(d, c, 1.5, 0.75);
  k<OP><<<1, 32>>>(d, c, 1.5, 0.75);
  long long h = 0; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %8.1f cycles / op\n", name, (double) h / CHAIN);
  cudaFree(d); cudaFree(c);
}

int main()
{
  run<0>("dfma");
  run<1>("dadd");
  run<2>("1/x (+add)");
  run<3>("sqrt (+add)");
  run<4>("log (+add)");
  run<5>("exp (+mul)");
  run<6>("smem load dependent (+add)");
  run<7>("shfl (+add)");
  run<8>("syncwarp (+add)");
  run<9>("f64->f32->f64 mul");
  run<10>("log1p (+mul)");
  run<11>("x/y");
  return 0;
}
