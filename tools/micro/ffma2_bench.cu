// Microbenchmark: latency and throughput of FFMA vs packed FFMA2 (sm_100a) to decide whether pairing the
// left/right leg arithmetic into f32x2 instructions pays. nvcc -arch=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool PACKED>
__global__ void k(float* out, int iters, float a, float b) {
  float2 x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (PACKED) {
        x[i] = __ffma2_rn(x[i], aa, bb);
      } else {
        x[i].x = fmaf(x[i].x, aa.x, bb.x);
        x[i].y = fmaf(x[i].y, aa.y, bb.y);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = float(t1 - t0);
}

template <int ILP, bool PACKED>
void run(const char* name, int blocks, int threads) {
  float* d;
  cudaMalloc(&d, blocks * threads * sizeof(float));
  const int iters = 20000;
  k<ILP, PACKED><<<blocks, threads>>>(d, iters, 1.0001f, 1e-6f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ILP, PACKED><<<blocks, threads>>>(d, iters, 1.0001f, 1e-6f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
  const double fma = double(blocks) * threads * iters * ILP * 2.0;
  printf("%-22s blocks=%4d thr=%4d ILP=%d: %.3f ms, %.1f cycles/iter (warp 0), %.2f T-fma/s\n", name, blocks, threads, ILP, ms,
         cyc / iters, fma / ms / 1e9);
  cudaFree(d);
}

int main() {
  // latency: one warp, one dependent chain (pair)
  run<1, false>("FFMA x2 (scalar)", 1, 32);
  run<1, true>("FFMA2 (packed)", 1, 32);
  run<4, false>("FFMA x2 (scalar)", 1, 32);
  run<4, true>("FFMA2 (packed)", 1, 32);
  // throughput: full chip, 8 warps/SM like the step kernel, ILP 2
  run<2, false>("FFMA x2 (scalar)", 148, 256);
  run<2, true>("FFMA2 (packed)", 148, 256);
  // throughput: saturated
  run<8, false>("FFMA x2 (scalar)", 148 * 4, 512);
  run<8, true>("FFMA2 (packed)", 148 * 4, 512);
  return 0;
}
