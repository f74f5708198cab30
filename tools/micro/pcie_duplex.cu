// SPDX-License-Identifier: Apache-2.0
// Developer microbenchmark: what the host<->device leg of one env step can reach on this box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_duplex pcie_duplex.cu && ./pcie_duplex
// Measures pinned H2D (9.4 MB = 65536 x 144 B) and D2H (8.3 MB = 65536 x 126 B) alone, concurrently on two streams
// (full duplex), chunked, and through zero-copy (SM loads/stores on mapped pinned memory).
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void zc_read(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) dst[i] = src[i];
}
// one thread = one env: 9 x float4 in (144 B), then 30 floats out as 15 float2 (like k_step)
__global__ void zc_env(const float4* __restrict__ act, float2* __restrict__ obs, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { const float4 v = act[size_t(i) * 9 + k]; s += v.x + v.y + v.z + v.w; }
#pragma unroll
  for (int k = 0; k < 15; ++k) obs[size_t(i) * 15 + k] = make_float2(s, s + k);
}

template <typename F>
double wall_ms(F f, int reps) {
  for (int i = 0; i < 3; ++i) f();
  cudaDeviceSynchronize();
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) f();
  cudaDeviceSynchronize();
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
}

int main() {
  const int n = 65536;
  const size_t ab = size_t(n) * 144, ob = size_t(n) * 126;
  char *ha, *ho, *da, *dob;
  CK(cudaHostAlloc(&ha, ab, cudaHostAllocMapped));
  CK(cudaHostAlloc(&ho, ob, cudaHostAllocMapped));
  CK(cudaMalloc(&da, ab));
  CK(cudaMalloc(&dob, ob));
  cudaStream_t s1, s2, s3;
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
  const int R = 100;
  double t;
  t = wall_ms([&] { cudaMemcpyAsync(da, ha, ab, cudaMemcpyHostToDevice, s1); cudaStreamSynchronize(s1); }, R);
  printf("H2D %.2f MB alone + sync: %.4f ms (%.1f GB/s)\n", ab / 1e6, t, ab / t / 1e6);
  t = wall_ms([&] { cudaMemcpyAsync(ho, dob, ob, cudaMemcpyDeviceToHost, s2); cudaStreamSynchronize(s2); }, R);
  printf("D2H %.2f MB alone + sync: %.4f ms (%.1f GB/s)\n", ob / 1e6, t, ob / t / 1e6);
  t = wall_ms([&] {
    cudaMemcpyAsync(da, ha, ab, cudaMemcpyHostToDevice, s1);
    cudaMemcpyAsync(ho, dob, ob, cudaMemcpyDeviceToHost, s2);
    cudaStreamSynchronize(s1); cudaStreamSynchronize(s2);
  }, R);
  printf("H2D + D2H concurrently (2 streams): %.4f ms\n", t);
  t = wall_ms([&] {
    cudaMemcpyAsync(da, ha, ab, cudaMemcpyHostToDevice, s1);
    cudaMemcpyAsync(ho, dob, ob, cudaMemcpyDeviceToHost, s1);
    cudaStreamSynchronize(s1);
  }, R);
  printf("H2D then D2H (1 stream): %.4f ms\n", t);
  for (int c : {2, 4, 8, 16}) {
    t = wall_ms([&] {
      for (int k = 0; k < c; ++k) cudaMemcpyAsync(da + k * (ab / c), ha + k * (ab / c), ab / c, cudaMemcpyHostToDevice, s1);
      cudaStreamSynchronize(s1);
    }, R);
    printf("H2D in %d chunks (1 stream): %.4f ms\n", c, t);
    t = wall_ms([&] {
      for (int k = 0; k < c; ++k) {
        cudaMemcpyAsync(da + k * (ab / c), ha + k * (ab / c), ab / c, cudaMemcpyHostToDevice, s1);
        cudaMemcpyAsync(ho + k * (ob / c), dob + k * (ob / c), ob / c, cudaMemcpyDeviceToHost, s2);
      }
      cudaStreamSynchronize(s1); cudaStreamSynchronize(s2);
    }, R);
    printf("H2D + D2H concurrently in %d chunks each: %.4f ms\n", c, t);
  }
  // zero-copy
  float4 *zha; float2* zho;
  CK(cudaHostGetDevicePointer(&zha, ha, 0));
  CK(cudaHostGetDevicePointer(&zho, ho, 0));
  for (int grid : {148, 296, 592, 1184}) {
    t = wall_ms([&] { zc_read<<<grid, 256, 0, s1>>>(zha, reinterpret_cast<float4*>(da), ab / 16); cudaStreamSynchronize(s1); }, R);
    printf("zero-copy read 9.4 MB, grid %d x 256: %.4f ms (%.1f GB/s)\n", grid, t, ab / t / 1e6);
    t = wall_ms([&] { zc_read<<<grid, 256, 0, s1>>>(reinterpret_cast<const float4*>(dob), reinterpret_cast<float4*>(ho), ob / 16); cudaStreamSynchronize(s1); }, R);
    printf("zero-copy write 8.3 MB, grid %d x 256: %.4f ms (%.1f GB/s)\n", grid, t, ob / t / 1e6);
  }
  t = wall_ms([&] { zc_env<<<n / 224 + 1, 224, 0, s1>>>(zha, zho, n); cudaStreamSynchronize(s1); }, R);
  printf("zero-copy env-shaped kernel (144 B in, 120 B out per thread, 65536 threads): %.4f ms\n", t);
  t = wall_ms([&] { zc_env<<<n / 224 + 1, 224, 0, s1>>>(reinterpret_cast<float4*>(da), reinterpret_cast<float2*>(dob), n); cudaStreamSynchronize(s1); }, R);
  printf("same kernel on device memory: %.4f ms\n", t);
  // launch + sync floor
  t = wall_ms([&] { zc_env<<<1, 32, 0, s1>>>(reinterpret_cast<float4*>(da), reinterpret_cast<float2*>(dob), 32); cudaStreamSynchronize(s1); }, R);
  printf("empty launch + stream sync: %.4f ms\n", t);
  return 0;
}
