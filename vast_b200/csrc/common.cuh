// Shared host/device helpers for the vast_b200 C-ABI library.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/vast_b200.h"

namespace vast {

void set_last_error(const char* fmt, ...);

#define VAST_REQUIRE(cond, code, ...)    \
  do {                                   \
    if (!(cond)) {                       \
      ::vast::set_last_error(__VA_ARGS__); \
      return (code);                     \
    }                                    \
  } while (0)

#define VAST_CUDA_OK(expr)                                                                         \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::vast::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VAST_ERR_CUDA;                                                                        \
    }                                                                                              \
  } while (0)

// cudaGetLastError after a launch (does not synchronise).
#define VAST_LAUNCH_OK(name)                                                                \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      ::vast::set_last_error("launch of %s failed: %s", name, cudaGetErrorString(_e));      \
      return VAST_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int device_sm_count();

// Optional per-kernel timing (bench.py's roofline leg): when enabled through vast_timing_enable,
// every launch wrapped in VAST_TIMED is bracketed by CUDA events on its own stream.
int timing_begin(const char* name, cudaStream_t stream);
void timing_end(int slot, cudaStream_t stream);
#define VAST_TIMED(stream, name, ...)                        \
  do {                                                       \
    const int _ts = ::vast::timing_begin(name, stream);      \
    __VA_ARGS__;                                             \
    ::vast::timing_end(_ts, stream);                         \
  } while (0)

// Programmatic dependent launch: a kernel launched with the attribute may begin (barrier / TMEM set-up, descriptor
// prefetch) while its predecessor in the stream drains; it must execute pdl_wait() before touching anything a
// predecessor wrote.  pdl_trigger() lets the successor start launching once every CTA of this grid has reached it.
// 256-bit global store (sm_100: STG.E.ENL2.256): a thread that owns a row segment writes a whole 32-byte sector per
// instruction.  The address must be 32-byte aligned.
__device__ __forceinline__ void st_global_v8_b32(void* p, const uint4& lo, const uint4& hi) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
               "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}
__device__ __forceinline__ void st_global_v8_f32(void* p, const float (&g)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(g[0]), "f"(g[1]), "f"(g[2]), "f"(g[3]),
               "f"(g[4]), "f"(g[5]), "f"(g[6]), "f"(g[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_v8_b32(const void* p, uint4& lo, uint4& hi) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <class... KArgs, class... Args>
cudaError_t launch_ex(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, int cluster,
                      Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = static_cast<unsigned>(n);
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Bump allocator over the caller-provided workspace (the library never allocates).
struct Workspace {
  char* base;
  size_t cap;
  size_t off;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), cap(n), off(0) {}
  template <class T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return base != nullptr ? off <= cap : off == 0; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// order-preserving map float -> uint32 (larger float <=> larger uint; 0 is below every float): atomicMax on floats
__device__ __forceinline__ uint32_t f32_orderable(float s) {
  const uint32_t b = __float_as_uint(s);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// 16-byte streaming load / store (read-once, write-once data).
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
  const uint4 r = ld_stream16(p);
  return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

}  // namespace vast
