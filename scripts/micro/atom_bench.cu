// Micro-benchmark: throughput of the atomic flavours the group-by kernel could use (B200).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

// mode 0: ATOMS.ADD u32 x1   1: ATOMS.ADD u32 x2 (count + lo)   2: shared atomicAdd u64   3: global RED u64 x1
// mode 4: global RED u64 x2  5: global RED u32 x2  6: shared u32 count + global RED u64 sum
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(unsigned long long* gtab, uint32_t slots, uint32_t iters, unsigned long long* sink) {
  extern __shared__ unsigned long long sm[];
  uint32_t* sm32 = reinterpret_cast<uint32_t*>(sm);
  for (uint32_t i = threadIdx.x; i < slots * 2; i += blockDim.x) sm32[i] = 0;
  __syncthreads();
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  unsigned long long acc = 0;
  for (uint32_t i = 0; i < iters; i++) {
    const uint32_t slot = rng(s) % slots;
    const uint32_t v = s & 1023u;
    if (MODE == 0) atomicAdd(sm32 + slot, v);
    if (MODE == 1) { atomicAdd(sm32 + 2 * slot, 1u); uint32_t old = atomicAdd(sm32 + 2 * slot + 1, v); if (old + v < old) acc++; }
    if (MODE == 2) atomicAdd(sm + slot, (1ull << 32) | v);
    if (MODE == 3) atomicAdd(gtab + slot, (unsigned long long)v);
    if (MODE == 4) { atomicAdd(gtab + 2 * slot, 1ull); atomicAdd(gtab + 2 * slot + 1, (unsigned long long)v); }
    if (MODE == 5) { atomicAdd(reinterpret_cast<unsigned int*>(gtab) + 2 * slot, 1u); atomicAdd(reinterpret_cast<unsigned int*>(gtab) + 2 * slot + 1, v); }
    if (MODE == 6) { atomicAdd(sm32 + slot, 1u); atomicAdd(gtab + slot, (unsigned long long)v); }
    if (MODE == 7) { unsigned long long old = atomicAdd(sm + slot, (1ull << 32) | v); if (uint32_t(old) + v < uint32_t(old)) acc++; }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < slots; i += blockDim.x) acc += sm[i];
  if (acc == 0xdeadbeefull) *sink = acc;
}

template <int MODE>
void run(const char* name, uint32_t slots, int threads, int per_sm) {
  unsigned long long *gtab, *sink;
  cudaMalloc(&gtab, size_t(slots) * 16);
  cudaMemset(gtab, 0, size_t(slots) * 16);
  cudaMalloc(&sink, 8);
  const uint32_t iters = 4096;
  const size_t smem = size_t(slots) * 8;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * per_sm;
  k<MODE><<<grid, threads, smem>>>(gtab, slots, 64, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<grid, threads, smem>>>(gtab, slots, iters, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  const double rows = double(grid) * threads * iters;
  printf("%-44s slots %6u thr %4d x%d: %8.3f ms  %8.2f Grows/s  %6.2f cyc/lane/SM  %s\n", name, slots, threads, per_sm, ms, rows / ms / 1e6,
         ms * 1e-3 * 1.9e9 / (rows / 148), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(gtab); cudaFree(sink);
}

int main() {
  for (uint32_t slots : {17u, 1024u, 16705u}) {
    for (int thr : {512, 1024}) {
      run<0>("shared u32 x1", slots, thr, 1);
      run<1>("shared u32 x2 (count, lo+carry)", slots, thr, 1);
      run<2>("shared u64 x1 (packed count|lo)", slots, thr, 1);
      run<7>("shared u64 x1 + carry check", slots, thr, 1);
      run<3>("global RED u64 x1", slots, thr, 1);
      run<4>("global RED u64 x2", slots, thr, 1);
      run<5>("global RED u32 x2", slots, thr, 1);
      run<6>("shared u32 count + global RED u64", slots, thr, 1);
    }
  }
  run<3>("global RED u64 x1", 1u << 20, 1024, 1);
  run<4>("global RED u64 x2", 1u << 20, 1024, 1);
  run<3>("global RED u64 x1", 1u << 22, 1024, 1);
  return 0;
}
