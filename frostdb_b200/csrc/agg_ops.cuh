// Aggregate reducers shared by the scan kernels: the per-value combine, the identity a partial starts
// from, and the atomic that folds a partial into a table cell (aggregate.go:734-935 semantics).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace fgpu {

__device__ __forceinline__ void atomic_min_f64(long long* addr, double v) {
  // Go's `if v < minV` (aggregate.go:847-857): NaN never replaces, ties keep the stored value.
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (v < __longlong_as_double((long long)old)) {
    unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}
__device__ __forceinline__ void atomic_max_f64(long long* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (v > __longlong_as_double((long long)old)) {
    unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}

__device__ __forceinline__ void apply_agg(uint8_t func, bool is_float, long long* cell, long long bits) {
  if (func == 1 /*sum*/) {
    if (is_float) atomicAdd(reinterpret_cast<double*>(cell), __longlong_as_double(bits));
    else atomicAdd(reinterpret_cast<unsigned long long*>(cell), (unsigned long long)bits);
  } else if (func == 2 /*min*/) {
    if (is_float) atomic_min_f64(cell, __longlong_as_double(bits));
    else atomicMin(cell, bits);
  } else if (func == 3 /*max*/) {
    if (is_float) atomic_max_f64(cell, __longlong_as_double(bits));
    else atomicMax(cell, bits);
  }
}

__device__ __forceinline__ long long agg_identity(uint8_t func, bool is_float) {
  if (func == 2) return is_float ? 0x7ff0000000000000ll : 0x7fffffffffffffffll;
  if (func == 3) return is_float ? (long long)0xfff0000000000000ull : (long long)0x8000000000000000ull;
  return 0;  // sum: +0 (int) / +0.0 (double)
}

__device__ __forceinline__ long long agg_combine(uint8_t func, bool is_float, long long a, long long b) {
  if (func == 1) {
    if (is_float) return __double_as_longlong(__longlong_as_double(a) + __longlong_as_double(b));
    return (long long)((unsigned long long)a + (unsigned long long)b);
  }
  if (func == 2) {
    if (is_float) return (__longlong_as_double(b) < __longlong_as_double(a)) ? b : a;
    return (b < a) ? b : a;
  }
  if (is_float) return (__longlong_as_double(b) > __longlong_as_double(a)) ? b : a;
  return (b > a) ? b : a;
}


}  // namespace fgpu
