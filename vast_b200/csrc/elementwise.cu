// HBM-bound kernels of the hot path: pool + concat (+ backward), L2 normalise (+ backward),
// all-gather send-buffer packing, negative-row gather + 3-way concat.  All coalesced, 128-bit
// vectorised, warp-shuffle reductions; grids sized so every SM has several CTAs in flight.
#include <type_traits>

#include "common.cuh"

namespace vast {

// ------------------------------------------------------------------ vector helpers
template <class T>
struct Vec;  // 16-byte vector of T <-> floats
template <>
struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&f)[4]) {
    const uint4 u = ld_stream16(p);
    f[0] = __uint_as_float(u.x);
    f[1] = __uint_as_float(u.y);
    f[2] = __uint_as_float(u.z);
    f[3] = __uint_as_float(u.w);
  }
  __device__ static void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 u = ld_stream16(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 u;
    uint32_t* w = &u.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <>
struct Vec<__half> {
  static constexpr int N = 8;
  __device__ static void load(const __half* p, float (&f)[8]) {
    const uint4 u = ld_stream16(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ static void store(__half* p, const float (&f)[8]) {
    uint4 u;
    uint32_t* w = &u.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = u;
  }
};

template <class T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <class T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// ------------------------------------------------------------------ pool + concat
struct PoolSeg {
  const void* src;  // [bs, n, tok, c]
  int64_t n, tok, c;
  int mode;         // 0 = token 0 only, 1 = mean over tokens
  int64_t out_off;  // column offset in the concat row
};
struct PoolParams {
  PoolSeg seg[3];
  int64_t bs;
  void* out;
  int64_t ldo;
};

constexpr int POOL_THREADS = 512;

// grid = (bs, 3).  Thread = (vector column, token lane); token lanes stride the (frame, token)
// list, partial sums are combined through shared memory in a fixed order (deterministic).
template <class TI, class TO>
__global__ void __launch_bounds__(POOL_THREADS) pool_concat_kernel(const PoolParams p) {
  const PoolSeg s = p.seg[blockIdx.y];
  if (s.src == nullptr) return;
  constexpr int V = Vec<TI>::N;
  const int64_t b = blockIdx.x;
  const int64_t used = s.mode == 0 ? 1 : s.tok;  // tokens that contribute per frame
  const int64_t count = s.n * used;
  const bool vec_ok = (s.c % V) == 0;
  __shared__ float red[POOL_THREADS * 8];
  const TI* base = static_cast<const TI*>(s.src) + b * s.n * s.tok * s.c;
  TO* out = static_cast<TO*>(p.out) + b * p.ldo + s.out_off;
  const float inv = 1.0f / static_cast<float>(count);
  if (vec_ok) {
    const int nvec = static_cast<int>(s.c / V);
    for (int v0 = 0; v0 < nvec; v0 += POOL_THREADS) {  // column super-chunks (c > 512*V only)
      const int ncol = min(nvec - v0, POOL_THREADS);
      const int lanes = POOL_THREADS / ncol;  // token lanes
      const int vc = threadIdx.x % ncol;
      const int tl = threadIdx.x / ncol;
      float acc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.f;
      if (tl < lanes) {
        // (frame, token) advance incrementally: a 64-bit division per 16-byte load made this loop ALU-bound
        const int iused = static_cast<int>(used), icount = static_cast<int>(count);
        int f = tl / iused, t = tl - f * iused;
        const TI* col = base + static_cast<int64_t>(v0 + vc) * V;
        const int64_t tok_stride = s.c, frame_stride = s.tok * s.c;
#pragma unroll 8
        for (int j = tl; j < icount; j += lanes) {
          float x[V];
          Vec<TI>::load(col + f * frame_stride + t * tok_stride, x);
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] += x[i];
          t += lanes;
          while (t >= iused) {
            t -= iused;
            ++f;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) red[threadIdx.x * V + i] = acc[i];
      __syncthreads();
      if (tl == 0 && vc < ncol) {
        for (int l = 1; l < lanes; ++l)
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] += red[(l * ncol + vc) * V + i];
#pragma unroll
        for (int i = 0; i < V; ++i) out[static_cast<int64_t>(v0 + vc) * V + i] = from_f<TO>(acc[i] * inv);
      }
      __syncthreads();
    }
  } else {  // scalar fallback for odd channel counts
    for (int64_t c = threadIdx.x; c < s.c; c += POOL_THREADS) {
      float a = 0.f;
      for (int64_t j = 0; j < count; ++j) {
        const int64_t f = j / used, t = j - f * used;
        a += to_f<TI>(base[(f * s.tok + t) * s.c + c]);
      }
      out[c] = from_f<TO>(a * inv);
    }
  }
}

// Backward: dense gradient of one encoder output, every element written.
template <class T>
__global__ void pool_bwd_kernel(const float* __restrict__ g, int64_t ldg, int64_t col_off, T* __restrict__ dst,
                                int64_t bs, int64_t n, int64_t tok, int64_t c, int mode) {
  const int64_t total = bs * n * tok * c;
  const float scale = 1.0f / static_cast<float>(mode == 0 ? n : n * tok);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t ch = i % c;
    const int64_t t = (i / c) % tok;
    const int64_t b = i / (c * tok * n);
    float v = 0.f;
    if (mode == 1 || t == 0) v = g[b * ldg + col_off + ch] * scale;
    dst[i] = from_f<T>(v);
  }
}

// ------------------------------------------------------------------ L2 normalise
// One warp per row, two passes over the row (second pass hits L1/L2).
template <class TI>
__global__ void __launch_bounds__(128) l2norm_kernel(const TI* __restrict__ x, int64_t rows, int64_t dim, int64_t ldx,
                                                     float eps, float* __restrict__ y32, int64_t ldy,
                                                     __nv_bfloat16* __restrict__ y16, int64_t ld16,
                                                     float* __restrict__ inv_norm) {
  const int64_t row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const TI* xr = x + row * ldx;
  float ss = 0.f;
  for (int64_t c = lane; c < dim; c += 32) {
    const float v = to_f<TI>(xr[c]);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float denom = fmaxf(sqrtf(ss), eps);
  if (lane == 0 && inv_norm) inv_norm[row] = 1.0f / denom;
  for (int64_t c = lane; c < dim; c += 32) {
    const float v = to_f<TI>(xr[c]) / denom;
    if (y32) y32[row * ldy + c] = v;
    if (y16) y16[row * ld16 + c] = __float2bfloat16_rn(v);
  }
}

// Vectorised variant (dim, ldx multiples of the 16-byte vector; outputs 16-byte aligned): 128-bit loads, float4 /
// 16-byte bf16 stores; the second pass re-reads the row from L1.
template <class TI>
__global__ void __launch_bounds__(128) l2norm_vec_kernel(const TI* __restrict__ x, int64_t rows, int dim, int64_t ldx, float eps,
                                                         float* __restrict__ y32, int64_t ldy,
                                                         __nv_bfloat16* __restrict__ y16, int64_t ld16,
                                                         float* __restrict__ inv_norm) {
  constexpr int V = Vec<TI>::N;
  const int64_t row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const TI* xr = x + row * ldx;
  float ss = 0.f;
#pragma unroll 4
  for (int c = lane * V; c < dim; c += 32 * V) {
    float v[V];
    const uint4 u = *reinterpret_cast<const uint4*>(xr + c);  // cached: read again below
    if constexpr (sizeof(TI) == 4) {
      v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    } else {
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if constexpr (std::is_same<TI, __nv_bfloat16>::value) {
          v[2 * i] = __uint_as_float(w[i] << 16);
          v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        } else {
          const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
          v[2 * i] = t.x;
          v[2 * i + 1] = t.y;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) ss = fmaf(v[i], v[i], ss);
  }
  ss = warp_sum(ss);
  const float denom = fmaxf(sqrtf(ss), eps);
  if (lane == 0 && inv_norm) inv_norm[row] = 1.0f / denom;
#pragma unroll 4
  for (int c = lane * V; c < dim; c += 32 * V) {
    float v[V];
    Vec<TI>::load(xr + c, v);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = v[i] / denom;
    if (y32) {
#pragma unroll
      for (int i = 0; i < V; i += 4)
        *reinterpret_cast<float4*>(y32 + row * ldy + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    if (y16) {
      uint32_t w[V / 2];
#pragma unroll
      for (int i = 0; i < V / 2; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
      }
      if constexpr (V == 8)
        *reinterpret_cast<uint4*>(y16 + row * ld16 + c) = make_uint4(w[0], w[1], w[2], w[3]);
      else
        *reinterpret_cast<uint2*>(y16 + row * ld16 + c) = make_uint2(w[0], w[1]);
    }
  }
}

__global__ void __launch_bounds__(128) l2norm_bwd_kernel(const float* __restrict__ g, int64_t ldg,
                                                         const float* __restrict__ y, int64_t ldy,
                                                         const float* __restrict__ inv_norm, int64_t rows, int64_t dim,
                                                         float eps, float* __restrict__ gx, int64_t ldgx) {
  const int64_t row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float inv = inv_norm[row];
  float dot = 0.f;
  for (int64_t c = lane; c < dim; c += 32) dot = fmaf(g[row * ldg + c], y[row * ldy + c], dot);
  dot = warp_sum(dot);
  const bool clamped = inv >= 1.0f / eps;  // ||x|| <= eps: y = x / eps, no projection term
  for (int64_t c = lane; c < dim; c += 32) {
    const float gv = g[row * ldg + c];
    gx[row * ldgx + c] = clamped ? gv * inv : (gv - y[row * ldy + c] * dot) * inv;
  }
}

// ------------------------------------------------------------------ all-gather send-buffer packing
template <class TI>
__global__ void pack_pair_kernel(const TI* __restrict__ ft, const TI* __restrict__ fc, int64_t bs, int64_t dim,
                                 int64_t ld, __nv_bfloat16* __restrict__ pack) {
  const int64_t total = bs * dim;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = i / dim, c = i - b * dim;
    pack[b * 2 * dim + c] = __float2bfloat16_rn(to_f<TI>(ft[b * ld + c]));
    pack[b * 2 * dim + dim + c] = __float2bfloat16_rn(to_f<TI>(fc[b * ld + c]));
  }
}

// Vectorised variant (dim % 8 == 0, ld % 8 == 0, 16-byte aligned pointers): one thread converts 8
// consecutive elements of one of the two inputs -> one 16-byte store.
template <class TI>
__global__ void __launch_bounds__(256) pack_pair_vec_kernel(const TI* __restrict__ ft, const TI* __restrict__ fc, int64_t bs,
                                                           int dim8, int64_t ld, __nv_bfloat16* __restrict__ pack) {
  const int64_t total = bs * 2 * dim8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = i / (2 * dim8);
    const int v = static_cast<int>(i - b * 2 * dim8);  // 16-byte vector inside the packed row
    const bool second = v >= dim8;
    const TI* src = (second ? fc : ft) + b * ld + static_cast<int64_t>(second ? v - dim8 : v) * 8;
    uint4 o;
    if constexpr (sizeof(TI) == 4) {
      const float4 a = ld_stream_f4(src), c = ld_stream_f4(src + 4);
      __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(c.x, c.y), h3 = __floats2bfloat162_rn(c.z, c.w);
      o.x = *reinterpret_cast<uint32_t*>(&h0);
      o.y = *reinterpret_cast<uint32_t*>(&h1);
      o.z = *reinterpret_cast<uint32_t*>(&h2);
      o.w = *reinterpret_cast<uint32_t*>(&h3);
    } else if constexpr (std::is_same<TI, __nv_bfloat16>::value) {
      o = ld_stream16(src);
    } else {
      const uint4 raw = ld_stream16(src);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
      uint32_t* ow = &o.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        __nv_bfloat162 r = __floats2bfloat162_rn(f.x, f.y);
        ow[j] = *reinterpret_cast<uint32_t*>(&r);
      }
    }
    *reinterpret_cast<uint4*>(pack + b * 16 * dim8 + static_cast<int64_t>(v) * 8) = o;
  }
}

// ------------------------------------------------------------------ fused pack + all-gather over NVLink
// The all-gather of model/vast.py:395,404 done by the packing kernel itself: every 16-byte vector of this rank's
// packed rows is stored straight into the gathered buffer of EVERY rank of the node -- one `multimem.st` to the
// NVSwitch multicast address of the symmetric buffer (the switch replicates it), or, without multicast support,
// one peer store per rank.  No intermediate send buffer and no NCCL call; a cross-rank barrier after the kernel
// (issued by the host side on the same stream) publishes the buffer.
constexpr int PUSH_MAX_PEERS = 16;
struct PushDst {
  void* mc;                    // multicast address of the gathered buffer, or nullptr
  void* peer[PUSH_MAX_PEERS];  // unicast address of the gathered buffer on every rank (used when mc == nullptr)
  int world;
};
__device__ __forceinline__ void multimem_st16(void* p, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}
// Arrival flags (vast_pack_pair_push_signal): instead of a cross-rank barrier AFTER the kernel, the kernel itself tells
// every rank "rank `me`'s rows of this step are in your buffer": every thread fences its stores at system scope, the
// last block to finish (ticket) bumps this rank's push epoch and release-stores it into slot `me` of every rank's flag
// array.  Consumers (vast_wait_arrivals, or the S GEMM's TMA producer) acquire-load their own array.  The kernel also
// lets its successor in the stream start launching at once (griddepcontrol): what follows waits on the flags, not on
// this kernel's end-of-grid flush.
struct PushSignal {
  unsigned* flag[PUSH_MAX_PEERS];  // every rank's flag array [world] (this rank's slot: index `me`)
  int* state;                      // local: [0] ticket, [1] push epoch, [2] consumer epoch
  int me;
};
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <class TI, bool SIGNAL>
__global__ void __launch_bounds__(256) pack_pair_push_kernel(const TI* __restrict__ ft, const TI* __restrict__ fc, int64_t bs,
                                                            int dim8, int64_t ld, int64_t row_offset, const PushDst dst,
                                                            const PushSignal sig) {
  if constexpr (SIGNAL) pdl_trigger();
  const int64_t total = bs * 2 * dim8;
  // (the signalling form runs a grid-stride loop over few, fat blocks: one system-scope fence per block)
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
  const int64_t b = i / (2 * dim8);
  const int v = static_cast<int>(i - b * 2 * dim8);  // 16-byte vector inside the packed row
  const bool second = v >= dim8;
  const TI* src = (second ? fc : ft) + b * ld + static_cast<int64_t>(second ? v - dim8 : v) * 8;
  uint4 o;
  if constexpr (sizeof(TI) == 4) {
    const float4 a = ld_stream_f4(src), c = ld_stream_f4(src + 4);
    __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(c.x, c.y), h3 = __floats2bfloat162_rn(c.z, c.w);
    o.x = *reinterpret_cast<uint32_t*>(&h0);
    o.y = *reinterpret_cast<uint32_t*>(&h1);
    o.z = *reinterpret_cast<uint32_t*>(&h2);
    o.w = *reinterpret_cast<uint32_t*>(&h3);
  } else if constexpr (std::is_same<TI, __nv_bfloat16>::value) {
    o = ld_stream16(src);
  } else {
    const uint4 raw = ld_stream16(src);
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
    uint32_t* ow = &o.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      __nv_bfloat162 r = __floats2bfloat162_rn(f.x, f.y);
      ow[j] = *reinterpret_cast<uint32_t*>(&r);
    }
  }
  const int64_t off = ((row_offset + b) * 2 * dim8 + v) * 16;  // byte offset inside the gathered [N, 2D] bf16 buffer
  if (dst.mc != nullptr) {
    multimem_st16(static_cast<char*>(dst.mc) + off, o);
  } else {
    for (int r = 0; r < dst.world; ++r) *reinterpret_cast<uint4*>(static_cast<char*>(dst.peer[r]) + off) = o;
  }
  }
  if constexpr (SIGNAL) {
    __syncthreads();  // the block's stores happen before thread 0's fence, which is cumulative over them
    if (threadIdx.x == 0) {
      __threadfence_system();  // ... and are performed (remote ones included) before the block takes its ticket
      if (atomicAdd(&sig.state[0], 1) == static_cast<int>(gridDim.x) - 1) {  // every block's stores are fenced
        __threadfence_system();
        const unsigned e = static_cast<unsigned>(sig.state[1]) + 1u;
        sig.state[1] = static_cast<int>(e);
        sig.state[0] = 0;
        for (int r = 0; r < dst.world; ++r) st_release_sys_u32(sig.flag[r] + sig.me, e);
      }
    }
  }
}

// One warp: lane r waits until rank r's rows of this step have arrived (flag >= consumer epoch + 1), then the consumer
// epoch advances.  Launched as a programmatic dependent of the push kernel: it spins while that kernel's stores drain.
__global__ void __launch_bounds__(32) wait_arrivals_kernel(const unsigned* __restrict__ flags, int world, int* __restrict__ state) {
  pdl_trigger();
  const unsigned want = static_cast<unsigned>(state[2]) + 1u;
  const int r = threadIdx.x;
  if (r < world) {
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_sys_u32(flags + r) - want) < 0) {
      if (clock64() - t0 > 6000000000LL) {  // ~3 s: a rank never pushed -- trap instead of hanging the GPU
        printf("vast_b200: arrival watchdog (rank slot %d, want %u, have %u)\n", r, want, ld_acquire_sys_u32(flags + r));
        __trap();
      }
    }
  }
  __syncwarp();
  if (r == 0) state[2] = static_cast<int>(want);
}

// ------------------------------------------------------------------ negative gather + 3-way concat
// Source row r in [0, 2bs): r < bs -> cond_local[r], written to out rows r and r + 2bs (read once,
// written twice); r >= bs -> cond_all[neg_cond[r - bs]] written to out row r.  Compulsory traffic:
// 2 reads + 3 writes of a [bs, S*H] block.
constexpr int GATHER_THREADS = 256;
constexpr int GATHER_UNROLL = 4;
__global__ void __launch_bounds__(GATHER_THREADS) gather_cond3_kernel(const uint4* __restrict__ cond_local,
                                                                     const uint4* __restrict__ cond_all,
                                                                     const int64_t* __restrict__ neg_cond, int64_t bs,
                                                                     int64_t n_total, int64_t row_vecs,
                                                                     uint4* __restrict__ out) {
  const int64_t r = blockIdx.y;
  const uint4* src;
  uint4* dst0 = out + r * row_vecs;
  uint4* dst1 = nullptr;
  if (r < bs) {
    src = cond_local + r * row_vecs;
    dst1 = out + (r + 2 * bs) * row_vecs;
  } else {
    int64_t j = neg_cond[r - bs];
    j = j < 0 ? 0 : (j >= n_total ? n_total - 1 : j);
    src = cond_all + j * row_vecs;
  }
  const int64_t chunk = static_cast<int64_t>(GATHER_THREADS) * GATHER_UNROLL;
  for (int64_t base = blockIdx.x * chunk; base < row_vecs; base += gridDim.x * chunk) {
    uint4 v[GATHER_UNROLL];
#pragma unroll
    for (int u = 0; u < GATHER_UNROLL; ++u) {
      const int64_t i = base + u * GATHER_THREADS + threadIdx.x;
      if (i < row_vecs) v[u] = ld_stream16(src + i);
    }
#pragma unroll
    for (int u = 0; u < GATHER_UNROLL; ++u) {
      const int64_t i = base + u * GATHER_THREADS + threadIdx.x;
      if (i < row_vecs) {
        st_stream16(dst0 + i, v[u]);
        if (dst1) st_stream16(dst1 + i, v[u]);
      }
    }
  }
}

__global__ void gather_ids3_kernel(const int64_t* __restrict__ ids_local, const int64_t* __restrict__ mask_local,
                                   const int64_t* __restrict__ ids_all, const int64_t* __restrict__ mask_all,
                                   const int64_t* __restrict__ neg_text, int64_t bs, int64_t n_total, int64_t L,
                                   int64_t* __restrict__ ids_out, int64_t* __restrict__ mask_out) {
  const int64_t total = 3 * bs * L;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / L, c = i - r * L;
    if (r < 2 * bs) {
      const int64_t s = r < bs ? r : r - bs;
      ids_out[i] = ids_local[s * L + c];
      mask_out[i] = mask_local[s * L + c];
    } else {
      int64_t j = neg_text[r - 2 * bs];
      j = j < 0 ? 0 : (j >= n_total ? n_total - 1 : j);
      ids_out[i] = ids_all[j * L + c];
      mask_out[i] = mask_all[j * L + c];
    }
  }
}

template <class TI>
static int launch_pool(const PoolParams& p, int out_dtype, cudaStream_t stream) {
  dim3 grid(static_cast<unsigned>(p.bs), 3);
  if (out_dtype == VAST_F32)
    pool_concat_kernel<TI, float><<<grid, POOL_THREADS, 0, stream>>>(p);
  else if (out_dtype == VAST_BF16)
    pool_concat_kernel<TI, __nv_bfloat16><<<grid, POOL_THREADS, 0, stream>>>(p);
  else
    pool_concat_kernel<TI, __half><<<grid, POOL_THREADS, 0, stream>>>(p);
  VAST_LAUNCH_OK("pool_concat");
  return VAST_OK;
}

static inline unsigned grid_for(int64_t total, int threads) {
  const int64_t want = ceil_div64(total, threads);
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  return static_cast<unsigned>(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace vast

using namespace vast;

extern "C" int vast_pool_concat(const void* vision, int64_t n_v, int64_t tok_v, int64_t c_v, int vision_mode,
                                const void* audio, int64_t n_a, int64_t tok_a, int64_t c_a, int audio_mode,
                                const void* subtitle, int64_t tok_s, int64_t c_s, int dtype, int64_t bs, void* out,
                                int out_dtype, int64_t ldo, vast_stream_t stream) {
  VAST_REQUIRE(out != nullptr && bs >= 0, VAST_ERR_INVALID, "pool_concat: null output");
  VAST_REQUIRE(vision || audio || subtitle, VAST_ERR_INVALID, "pool_concat: no modality given");
  VAST_REQUIRE(dtype >= VAST_F32 && dtype <= VAST_F16 && out_dtype >= VAST_F32 && out_dtype <= VAST_F16,
               VAST_ERR_UNSUPPORTED, "pool_concat: bad dtype");
  if (bs == 0) return VAST_OK;
  PoolParams p;
  memset(&p, 0, sizeof(p));
  int64_t off = 0;
  if (vision) {
    VAST_REQUIRE(n_v > 0 && tok_v > 0 && c_v > 0, VAST_ERR_INVALID, "pool_concat: bad vision shape");
    p.seg[0] = {vision, n_v, tok_v, c_v, vision_mode, off};
    off += c_v;
  }
  if (audio) {
    VAST_REQUIRE(n_a > 0 && tok_a > 0 && c_a > 0, VAST_ERR_INVALID, "pool_concat: bad audio shape");
    p.seg[1] = {audio, n_a, tok_a, c_a, audio_mode, off};
    off += c_a;
  }
  if (subtitle) {
    VAST_REQUIRE(tok_s > 0 && c_s > 0, VAST_ERR_INVALID, "pool_concat: bad subtitle shape");
    p.seg[2] = {subtitle, 1, tok_s, c_s, 0, off};
    off += c_s;
  }
  VAST_REQUIRE(ldo >= off, VAST_ERR_INVALID, "pool_concat: ldo < concat width");
  // vector loads need 16-byte aligned rows
  for (int i = 0; i < 3; ++i)
    if (p.seg[i].src)
      VAST_REQUIRE((reinterpret_cast<uintptr_t>(p.seg[i].src) & 15) == 0, VAST_ERR_INVALID,
                   "pool_concat: inputs must be 16-byte aligned");
  p.bs = bs;
  p.out = out;
  p.ldo = ldo;
  if (dtype == VAST_F32) return launch_pool<float>(p, out_dtype, stream);
  if (dtype == VAST_BF16) return launch_pool<__nv_bfloat16>(p, out_dtype, stream);
  return launch_pool<__half>(p, out_dtype, stream);
}

template <class T>
static int launch_pool_bwd(const float* g, int64_t ldg, int64_t off, void* dst, int64_t bs, int64_t n, int64_t tok,
                           int64_t c, int mode, cudaStream_t stream) {
  const int64_t total = bs * n * tok * c;
  if (total == 0) return VAST_OK;
  pool_bwd_kernel<T><<<grid_for(total, 256), 256, 0, stream>>>(g, ldg, off, static_cast<T*>(dst), bs, n, tok, c, mode);
  VAST_LAUNCH_OK("pool_concat_bwd");
  return VAST_OK;
}

extern "C" int vast_pool_concat_bwd(const float* grad_out, int64_t ldg, int64_t bs, void* grad_vision, int64_t n_v,
                                    int64_t tok_v, int64_t c_v, int vision_mode, void* grad_audio, int64_t n_a,
                                    int64_t tok_a, int64_t c_a, int audio_mode, void* grad_subtitle, int64_t tok_s,
                                    int64_t c_s, int dtype, vast_stream_t stream) {
  VAST_REQUIRE(grad_out != nullptr, VAST_ERR_INVALID, "pool_concat_bwd: null grad");
  VAST_REQUIRE(dtype >= VAST_F32 && dtype <= VAST_F16, VAST_ERR_UNSUPPORTED, "pool_concat_bwd: bad dtype");
  int64_t off = 0;
  struct S {
    void* dst;
    int64_t n, tok, c;
    int mode;
  } segs[3] = {{grad_vision, n_v, tok_v, c_v, vision_mode},
               {grad_audio, n_a, tok_a, c_a, audio_mode},
               {grad_subtitle, 1, tok_s, c_s, 0}};
  for (auto& s : segs) {
    if (!s.dst) {
      continue;
    }
    int rc;
    if (dtype == VAST_F32)
      rc = launch_pool_bwd<float>(grad_out, ldg, off, s.dst, bs, s.n, s.tok, s.c, s.mode, stream);
    else if (dtype == VAST_BF16)
      rc = launch_pool_bwd<__nv_bfloat16>(grad_out, ldg, off, s.dst, bs, s.n, s.tok, s.c, s.mode, stream);
    else
      rc = launch_pool_bwd<__half>(grad_out, ldg, off, s.dst, bs, s.n, s.tok, s.c, s.mode, stream);
    if (rc) return rc;
    off += s.c;
  }
  return VAST_OK;
}

extern "C" int vast_l2norm(const void* x, int x_dtype, int64_t rows, int64_t dim, int64_t ldx, float eps, float* y_f32,
                           int64_t ldy, void* y_16, int64_t ld16, float* inv_norm, vast_stream_t stream) {
  VAST_REQUIRE(x != nullptr && rows >= 0 && dim > 0, VAST_ERR_INVALID, "l2norm: bad arguments");
  VAST_REQUIRE(ldx >= dim && (!y_f32 || ldy >= dim) && (!y_16 || ld16 >= dim), VAST_ERR_INVALID, "l2norm: bad ld");
  if (rows == 0) return VAST_OK;
  const unsigned grid = static_cast<unsigned>(ceil_div64(rows, 4));
  auto* y16 = static_cast<__nv_bfloat16*>(y_16);
  const int64_t vin = x_dtype == VAST_F32 ? 4 : 8;  // elements per 16-byte input vector
  const bool vec = dim % vin == 0 && ldx % vin == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && dim < (1ll << 30) &&
                   (!y_f32 || (ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(y_f32) & 15) == 0)) &&
                   (!y_16 || (ld16 % vin == 0 && (reinterpret_cast<uintptr_t>(y_16) & (vin == 8 ? 15 : 7)) == 0));
  if (vec && (x_dtype == VAST_F32 || x_dtype == VAST_BF16 || x_dtype == VAST_F16)) {
    const int d = static_cast<int>(dim);
    if (x_dtype == VAST_F32)
      l2norm_vec_kernel<float><<<grid, 128, 0, stream>>>(static_cast<const float*>(x), rows, d, ldx, eps, y_f32, ldy, y16, ld16, inv_norm);
    else if (x_dtype == VAST_BF16)
      l2norm_vec_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), rows, d, ldx, eps, y_f32, ldy, y16, ld16, inv_norm);
    else
      l2norm_vec_kernel<__half><<<grid, 128, 0, stream>>>(static_cast<const __half*>(x), rows, d, ldx, eps, y_f32, ldy, y16, ld16, inv_norm);
    VAST_LAUNCH_OK("l2norm");
    return VAST_OK;
  }
  if (x_dtype == VAST_F32)
    l2norm_kernel<float><<<grid, 128, 0, stream>>>(static_cast<const float*>(x), rows, dim, ldx, eps, y_f32, ldy, y16, ld16, inv_norm);
  else if (x_dtype == VAST_BF16)
    l2norm_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), rows, dim, ldx, eps, y_f32, ldy, y16, ld16, inv_norm);
  else if (x_dtype == VAST_F16)
    l2norm_kernel<__half><<<grid, 128, 0, stream>>>(static_cast<const __half*>(x), rows, dim, ldx, eps, y_f32, ldy, y16, ld16, inv_norm);
  else
    VAST_REQUIRE(false, VAST_ERR_UNSUPPORTED, "l2norm: bad dtype");
  VAST_LAUNCH_OK("l2norm");
  return VAST_OK;
}

extern "C" int vast_l2norm_bwd(const float* grad_y, int64_t ldg, const float* y, int64_t ldy, const float* inv_norm,
                               int64_t rows, int64_t dim, float eps, float* grad_x, int64_t ldgx, vast_stream_t stream) {
  VAST_REQUIRE(grad_y && y && inv_norm && grad_x && dim > 0, VAST_ERR_INVALID, "l2norm_bwd: null pointer");
  if (rows == 0) return VAST_OK;
  l2norm_bwd_kernel<<<static_cast<unsigned>(ceil_div64(rows, 4)), 128, 0, stream>>>(grad_y, ldg, y, ldy, inv_norm, rows,
                                                                                   dim, eps, grad_x, ldgx);
  VAST_LAUNCH_OK("l2norm_bwd");
  return VAST_OK;
}

extern "C" int vast_pack_pair(const void* feat_t, const void* feat_cond, int dtype, int64_t bs, int64_t dim,
                              int64_t ld_in, void* pack_bf16, vast_stream_t stream) {
  VAST_REQUIRE(feat_t && feat_cond && pack_bf16 && dim > 0 && ld_in >= dim, VAST_ERR_INVALID, "pack_pair: bad arguments");
  if (bs == 0) return VAST_OK;
  auto* out = static_cast<__nv_bfloat16*>(pack_bf16);
  const bool vec = dim % 8 == 0 && ld_in % 8 == 0 &&
                   ((reinterpret_cast<uintptr_t>(feat_t) | reinterpret_cast<uintptr_t>(feat_cond) | reinterpret_cast<uintptr_t>(pack_bf16)) & 15) == 0;
  if (vec && (dtype == VAST_F32 || dtype == VAST_BF16 || dtype == VAST_F16)) {
    const int dim8 = static_cast<int>(dim / 8);
    const unsigned g = static_cast<unsigned>(ceil_div64(bs * 2 * dim8, 256));  // one 16-byte vector per thread
    if (dtype == VAST_F32)
      VAST_TIMED(stream, "pack_pair", (pack_pair_vec_kernel<float><<<g, 256, 0, stream>>>(static_cast<const float*>(feat_t), static_cast<const float*>(feat_cond), bs, dim8, ld_in, out)));
    else if (dtype == VAST_BF16)
      VAST_TIMED(stream, "pack_pair", (pack_pair_vec_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(feat_t), static_cast<const __nv_bfloat16*>(feat_cond), bs, dim8, ld_in, out)));
    else
      VAST_TIMED(stream, "pack_pair", (pack_pair_vec_kernel<__half><<<g, 256, 0, stream>>>(static_cast<const __half*>(feat_t), static_cast<const __half*>(feat_cond), bs, dim8, ld_in, out)));
    VAST_LAUNCH_OK("pack_pair");
    return VAST_OK;
  }
  const unsigned grid = grid_for(bs * dim, 256);
  if (dtype == VAST_F32)
    pack_pair_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(feat_t), static_cast<const float*>(feat_cond), bs, dim, ld_in, out);
  else if (dtype == VAST_BF16)
    pack_pair_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(feat_t), static_cast<const __nv_bfloat16*>(feat_cond), bs, dim, ld_in, out);
  else if (dtype == VAST_F16)
    pack_pair_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(feat_t), static_cast<const __half*>(feat_cond), bs, dim, ld_in, out);
  else
    VAST_REQUIRE(false, VAST_ERR_UNSUPPORTED, "pack_pair: bad dtype");
  VAST_LAUNCH_OK("pack_pair");
  return VAST_OK;
}

static int pack_pair_push_impl(const void* feat_t, const void* feat_cond, int dtype, int64_t bs, int64_t dim, int64_t ld_in,
                               int64_t row_offset, void* multicast_ptr, void* const* peer_ptrs, int world,
                               const PushSignal* sig, vast_stream_t stream) {
  VAST_REQUIRE(feat_t && feat_cond && dim > 0 && ld_in >= dim && bs >= 0 && row_offset >= 0, VAST_ERR_INVALID,
               "pack_pair_push: bad arguments");
  VAST_REQUIRE(world >= 1 && world <= PUSH_MAX_PEERS && (multicast_ptr != nullptr || peer_ptrs != nullptr), VAST_ERR_INVALID,
               "pack_pair_push: need a multicast pointer or 1..%d peer pointers", PUSH_MAX_PEERS);
  VAST_REQUIRE(dim % 8 == 0 && ld_in % 8 == 0 &&
                   ((reinterpret_cast<uintptr_t>(feat_t) | reinterpret_cast<uintptr_t>(feat_cond) |
                     reinterpret_cast<uintptr_t>(multicast_ptr)) & 15) == 0,
               VAST_ERR_UNSUPPORTED, "pack_pair_push: dim and ld must be multiples of 8, pointers 16-byte aligned");
  VAST_REQUIRE(bs > 0 || sig == nullptr, VAST_ERR_INVALID, "pack_pair_push_signal: an empty block cannot signal");
  if (bs == 0) return VAST_OK;
  PushDst dst;
  memset(&dst, 0, sizeof(dst));
  dst.mc = multicast_ptr;
  dst.world = world;
  if (multicast_ptr == nullptr)
    for (int r = 0; r < world; ++r) {
      VAST_REQUIRE(peer_ptrs[r] != nullptr && (reinterpret_cast<uintptr_t>(peer_ptrs[r]) & 15) == 0, VAST_ERR_INVALID,
                   "pack_pair_push: bad peer pointer %d", r);
      dst.peer[r] = peer_ptrs[r];
    }
  PushSignal none;
  memset(&none, 0, sizeof(none));
  const PushSignal& sg = sig ? *sig : none;
  const int dim8 = static_cast<int>(dim / 8);
  unsigned g = static_cast<unsigned>(ceil_div64(bs * 2 * dim8, 256));
  if (sig != nullptr && g > static_cast<unsigned>(2 * device_sm_count())) g = static_cast<unsigned>(2 * device_sm_count());
#define VAST_PUSH(T)                                                                                                           \
  do {                                                                                                                         \
    if (sig)                                                                                                                   \
      VAST_TIMED(stream, "pack_pair_push", (pack_pair_push_kernel<T, true><<<g, 256, 0, stream>>>(                              \
                                               static_cast<const T*>(feat_t), static_cast<const T*>(feat_cond), bs, dim8, ld_in, row_offset, dst, sg))); \
    else                                                                                                                       \
      VAST_TIMED(stream, "pack_pair_push", (pack_pair_push_kernel<T, false><<<g, 256, 0, stream>>>(                             \
                                               static_cast<const T*>(feat_t), static_cast<const T*>(feat_cond), bs, dim8, ld_in, row_offset, dst, sg))); \
  } while (0)
  if (dtype == VAST_F32)
    VAST_PUSH(float);
  else if (dtype == VAST_BF16)
    VAST_PUSH(__nv_bfloat16);
  else if (dtype == VAST_F16)
    VAST_PUSH(__half);
  else
    VAST_REQUIRE(false, VAST_ERR_UNSUPPORTED, "pack_pair_push: bad dtype");
#undef VAST_PUSH
  VAST_LAUNCH_OK("pack_pair_push");
  return VAST_OK;
}

extern "C" int vast_pack_pair_push(const void* feat_t, const void* feat_cond, int dtype, int64_t bs, int64_t dim, int64_t ld_in,
                                   int64_t row_offset, void* multicast_ptr, void* const* peer_ptrs, int world,
                                   vast_stream_t stream) {
  return pack_pair_push_impl(feat_t, feat_cond, dtype, bs, dim, ld_in, row_offset, multicast_ptr, peer_ptrs, world, nullptr, stream);
}

extern "C" int vast_pack_pair_push_signal(const void* feat_t, const void* feat_cond, int dtype, int64_t bs, int64_t dim,
                                          int64_t ld_in, int64_t row_offset, void* multicast_ptr, void* const* peer_ptrs,
                                          int world, void* const* flag_ptrs, int my_rank, int* sync_state,
                                          vast_stream_t stream) {
  VAST_REQUIRE(peer_ptrs != nullptr || multicast_ptr != nullptr, VAST_ERR_INVALID, "pack_pair_push_signal: no destination");
  VAST_REQUIRE(flag_ptrs && sync_state && world >= 1 && world <= PUSH_MAX_PEERS && my_rank >= 0 && my_rank < world,
               VAST_ERR_INVALID, "pack_pair_push_signal: bad flag arguments");
  PushSignal sig;
  memset(&sig, 0, sizeof(sig));
  for (int r = 0; r < world; ++r) {
    VAST_REQUIRE(flag_ptrs[r] != nullptr && (reinterpret_cast<uintptr_t>(flag_ptrs[r]) & 3) == 0, VAST_ERR_INVALID,
                 "pack_pair_push_signal: bad flag pointer %d", r);
    sig.flag[r] = static_cast<unsigned*>(flag_ptrs[r]);
  }
  sig.state = sync_state;
  sig.me = my_rank;
  return pack_pair_push_impl(feat_t, feat_cond, dtype, bs, dim, ld_in, row_offset, multicast_ptr, peer_ptrs, world, &sig, stream);
}

extern "C" int vast_wait_arrivals(const void* flags, int world, int* sync_state, vast_stream_t stream) {
  VAST_REQUIRE(flags && sync_state && world >= 1 && world <= 32, VAST_ERR_INVALID, "wait_arrivals: bad arguments");
  VAST_TIMED(stream, "wait_arrivals",
             (launch_ex(wait_arrivals_kernel, 1, 32, 0, stream, 1, static_cast<const unsigned*>(flags), world, sync_state)));
  VAST_LAUNCH_OK("wait_arrivals");
  return VAST_OK;
}

// ------------------------------------------------------------------ negative rows straight from their owners
// The sampled negative of local row b is global row j = neg_cond[b] of the rank-ordered batch: rank j / rows_per_rank
// holds it at local index j % rows_per_rank.  Every rank keeps its [rows_per_rank, S, H] block in symmetric memory
// (mapped into all ranks of the node), so the middle third of the [3bs, S, H] ITM input is read over NVLink from the
// owner's block directly into its final place: no all-gather of the whole [N, S, H] tensor with gradient
// (model/vast.py:422), no index exchange, no intermediate buffer, no host synchronisation.
struct PeerSrc {
  const uint4* base[PUSH_MAX_PEERS];
};
__global__ void __launch_bounds__(GATHER_THREADS) gather_cond3_peer_kernel(const uint4* __restrict__ cond_local, PeerSrc peers,
                                                                          const int64_t* __restrict__ neg_cond, int64_t bs,
                                                                          int64_t rows_per_rank, int world, int64_t row_vecs,
                                                                          uint4* __restrict__ out) {
  const int64_t r = blockIdx.y;
  const uint4* src;
  uint4* dst0 = out + r * row_vecs;
  uint4* dst1 = nullptr;
  if (r < bs) {
    src = cond_local + r * row_vecs;
    dst1 = out + (r + 2 * bs) * row_vecs;
  } else {
    int64_t j = neg_cond[r - bs];
    const int64_t n_total = rows_per_rank * world;
    j = j < 0 ? 0 : (j >= n_total ? n_total - 1 : j);
    const int owner = static_cast<int>(j / rows_per_rank);
    src = peers.base[owner] + (j - owner * rows_per_rank) * row_vecs;
  }
  const int64_t chunk = static_cast<int64_t>(GATHER_THREADS) * GATHER_UNROLL;
  for (int64_t base = blockIdx.x * chunk; base < row_vecs; base += gridDim.x * chunk) {
    uint4 v[GATHER_UNROLL];
#pragma unroll
    for (int u = 0; u < GATHER_UNROLL; ++u) {
      const int64_t i = base + u * GATHER_THREADS + threadIdx.x;
      if (i < row_vecs) v[u] = ld_stream16(src + i);
    }
#pragma unroll
    for (int u = 0; u < GATHER_UNROLL; ++u) {
      const int64_t i = base + u * GATHER_THREADS + threadIdx.x;
      if (i < row_vecs) {
        st_stream16(dst0 + i, v[u]);
        if (dst1) st_stream16(dst1 + i, v[u]);
      }
    }
  }
}

// Backward of the peer gather: the gradient of this rank's rows = sum of the gradient rows of every request that named
// them.  requests [world * bs] = the all-gathered neg_cond of all ranks (request e = rank e / bs, row e % bs); the
// requesters' gradient blocks [bs, S, H] sit in symmetric memory.  Kernel 1 lists, per local row, the requests that
// target it in ASCENDING order (so the fp32 sums below run in a fixed order: deterministic); kernel 2 pulls and adds.
constexpr int PULL_CAP = 16;
__global__ void __launch_bounds__(128) pull_list_kernel(const int64_t* __restrict__ requests, int64_t n_req, int64_t row0,
                                                       int64_t rows, int* __restrict__ counts, int* __restrict__ lists) {
  const int64_t l = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (l >= rows) return;
  const int lane = threadIdx.x & 31;
  int cnt = 0;
  for (int64_t e0 = 0; e0 < n_req; e0 += 32) {
    const int64_t e = e0 + lane;
    const bool hit = e < n_req && requests[e] == row0 + l;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int at = cnt + __popc(m & ((1u << lane) - 1u));
      if (at < PULL_CAP) lists[l * PULL_CAP + at] = static_cast<int>(e);
    }
    cnt += __popc(m);
  }
  if (lane == 0) counts[l] = cnt;
}
template <class T>
__device__ __forceinline__ void acc8(float (&a)[8], const uint4& v);
template <>
__device__ __forceinline__ void acc8<__nv_bfloat16>(float (&a)[8], const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a[2 * j] += __uint_as_float(w[j] << 16);
    a[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
  }
}
template <>
__device__ __forceinline__ void acc8<__half>(float (&a)[8], const uint4& v) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    a[2 * j] += f.x;
    a[2 * j + 1] += f.y;
  }
}
// T = 16-bit element (8 per 16-byte vector) or float (4 per vector); out = base (optional, same dtype) + pulled sum
template <class T>
__global__ void __launch_bounds__(GATHER_THREADS) pull_rows_kernel(PeerSrc peers, const int64_t* __restrict__ requests,
                                                                  int64_t n_req, int64_t bs, int64_t row0,
                                                                  const int* __restrict__ counts, const int* __restrict__ lists,
                                                                  const uint4* __restrict__ base_grad, int64_t row_vecs,
                                                                  uint4* __restrict__ out) {
  const int64_t l = blockIdx.y;
  const int cnt = counts[l];
  for (int64_t i = blockIdx.x * static_cast<int64_t>(GATHER_THREADS) + threadIdx.x; i < row_vecs;
       i += static_cast<int64_t>(gridDim.x) * GATHER_THREADS) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (base_grad != nullptr) {
      const uint4 v = ld_stream16(base_grad + l * row_vecs + i);
      if constexpr (sizeof(T) == 4) {
        a[0] = __uint_as_float(v.x); a[1] = __uint_as_float(v.y); a[2] = __uint_as_float(v.z); a[3] = __uint_as_float(v.w);
      } else {
        acc8<T>(a, v);
      }
    }
    auto add = [&](int e) {
      const uint4 v = ld_stream16(peers.base[e / bs] + (e % bs) * row_vecs + i);
      if constexpr (sizeof(T) == 4) {
        a[0] += __uint_as_float(v.x); a[1] += __uint_as_float(v.y); a[2] += __uint_as_float(v.z); a[3] += __uint_as_float(v.w);
      } else {
        acc8<T>(a, v);
      }
    };
    const int listed = cnt < PULL_CAP ? cnt : PULL_CAP;
    for (int q = 0; q < listed; ++q) add(lists[l * PULL_CAP + q]);
    if (cnt > PULL_CAP) {  // more requests than the list holds (a row sampled > 16 times): rescan, still ascending
      int seen = 0;
      for (int64_t e = 0; e < n_req; ++e)
        if (requests[e] == row0 + l && seen++ >= PULL_CAP) add(static_cast<int>(e));
    }
    uint4 o;
    if constexpr (sizeof(T) == 4) {
      o = make_uint4(__float_as_uint(a[0]), __float_as_uint(a[1]), __float_as_uint(a[2]), __float_as_uint(a[3]));
    } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
      uint32_t* w = &o.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a[2 * j], a[2 * j + 1]);
        w[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
    } else {
      uint32_t* w = &o.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2 h = __floats2half2_rn(a[2 * j], a[2 * j + 1]);
        w[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
    st_stream16(out + l * row_vecs + i, o);
  }
}

extern "C" int vast_gather_rows_concat3_peer(const int64_t* ids_local, const int64_t* mask_local, const int64_t* ids_all,
                                             const int64_t* mask_all, int64_t L, const void* cond_local,
                                             void* const* cond_peers, int world, int64_t rows_per_rank,
                                             int64_t row_bytes_cond, const int64_t* neg_text, const int64_t* neg_cond,
                                             int64_t bs, int64_t* ids_out, int64_t* mask_out, void* cond_out,
                                             vast_stream_t stream) {
  VAST_REQUIRE(bs >= 0 && rows_per_rank >= bs && world >= 1 && world <= PUSH_MAX_PEERS && cond_peers != nullptr, VAST_ERR_INVALID,
               "gather_rows_concat3_peer: bad arguments");
  if (bs == 0) return VAST_OK;
  VAST_REQUIRE(cond_local && cond_out && neg_cond, VAST_ERR_INVALID, "gather_rows_concat3_peer: null cond pointer");
  VAST_REQUIRE(row_bytes_cond > 0 && row_bytes_cond % 16 == 0, VAST_ERR_UNSUPPORTED,
               "gather_rows_concat3_peer: S*H*elem_bytes (%lld) must be a multiple of 16", (long long)row_bytes_cond);
  VAST_REQUIRE(2 * bs <= 65535, VAST_ERR_UNSUPPORTED, "gather_rows_concat3_peer: bs too large");
  PeerSrc ps;
  for (int r = 0; r < PUSH_MAX_PEERS; ++r) ps.base[r] = nullptr;
  uintptr_t bits = reinterpret_cast<uintptr_t>(cond_local) | reinterpret_cast<uintptr_t>(cond_out);
  for (int r = 0; r < world; ++r) {
    VAST_REQUIRE(cond_peers[r] != nullptr, VAST_ERR_INVALID, "gather_rows_concat3_peer: null peer pointer %d", r);
    ps.base[r] = static_cast<const uint4*>(cond_peers[r]);
    bits |= reinterpret_cast<uintptr_t>(cond_peers[r]);
  }
  VAST_REQUIRE((bits & 15) == 0, VAST_ERR_INVALID, "gather_rows_concat3_peer: cond pointers must be 16-byte aligned");
  const int64_t row_vecs = row_bytes_cond / 16;
  const int64_t chunk = GATHER_THREADS * GATHER_UNROLL;
  int64_t gx = ceil_div64(row_vecs, chunk);
  if (gx > 64) gx = 64;
  dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(2 * bs));
  gather_cond3_peer_kernel<<<grid, GATHER_THREADS, 0, stream>>>(static_cast<const uint4*>(cond_local), ps, neg_cond, bs, rows_per_rank,
                                                                world, row_vecs, static_cast<uint4*>(cond_out));
  VAST_LAUNCH_OK("gather_cond3_peer");
  if (ids_out || mask_out) {
    VAST_REQUIRE(ids_out && mask_out && ids_local && mask_local && ids_all && mask_all && neg_text && L > 0,
                 VAST_ERR_INVALID, "gather_rows_concat3_peer: null ids/mask pointer");
    gather_ids3_kernel<<<grid_for(3 * bs * L, 256), 256, 0, stream>>>(ids_local, mask_local, ids_all, mask_all, neg_text, bs,
                                                                      rows_per_rank * world, L, ids_out, mask_out);
    VAST_LAUNCH_OK("gather_ids3");
  }
  return VAST_OK;
}

extern "C" size_t vast_pull_row_grads_workspace_bytes(int64_t bs) {
  return bs > 0 ? align_up(sizeof(int) * bs, 256) + align_up(sizeof(int) * bs * PULL_CAP, 256) : 0;
}

extern "C" int vast_pull_row_grads(const int64_t* requests, void* const* grad_peers, int world, int64_t bs, int64_t row0,
                                   int64_t row_bytes, int dtype, const void* base_grad, void* out, void* workspace,
                                   size_t workspace_bytes, vast_stream_t stream) {
  VAST_REQUIRE(requests && grad_peers && out && workspace && world >= 1 && world <= PUSH_MAX_PEERS && bs > 0, VAST_ERR_INVALID,
               "pull_row_grads: bad arguments");
  VAST_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0 && bs <= 65535, VAST_ERR_UNSUPPORTED, "pull_row_grads: row bytes must be a multiple of 16");
  VAST_REQUIRE(workspace_bytes >= vast_pull_row_grads_workspace_bytes(bs), VAST_ERR_WORKSPACE, "pull_row_grads: workspace too small");
  PeerSrc ps;
  for (int r = 0; r < PUSH_MAX_PEERS; ++r) ps.base[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    VAST_REQUIRE(grad_peers[r] != nullptr && (reinterpret_cast<uintptr_t>(grad_peers[r]) & 15) == 0, VAST_ERR_INVALID,
                 "pull_row_grads: bad peer pointer %d", r);
    ps.base[r] = static_cast<const uint4*>(grad_peers[r]);
  }
  int* counts = static_cast<int*>(workspace);
  int* lists = reinterpret_cast<int*>(static_cast<char*>(workspace) + align_up(sizeof(int) * bs, 256));
  const int64_t n_req = bs * world;
  pull_list_kernel<<<static_cast<unsigned>(ceil_div64(bs, 4)), 128, 0, stream>>>(requests, n_req, row0, bs, counts, lists);
  VAST_LAUNCH_OK("pull_list");
  const int64_t row_vecs = row_bytes / 16;
  int64_t gx = ceil_div64(row_vecs, GATHER_THREADS * 4);
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(bs));
#define VAST_PULL(T)                                                                                                      \
  pull_rows_kernel<T><<<grid, GATHER_THREADS, 0, stream>>>(ps, requests, n_req, bs, row0, counts, lists,                  \
                                                          static_cast<const uint4*>(base_grad), row_vecs, static_cast<uint4*>(out))
  if (dtype == VAST_F32)
    VAST_PULL(float);
  else if (dtype == VAST_BF16)
    VAST_PULL(__nv_bfloat16);
  else if (dtype == VAST_F16)
    VAST_PULL(__half);
  else
    VAST_REQUIRE(false, VAST_ERR_UNSUPPORTED, "pull_row_grads: bad dtype");
#undef VAST_PULL
  VAST_LAUNCH_OK("pull_rows");
  return VAST_OK;
}

extern "C" int vast_gather_rows_concat3(const int64_t* ids_local, const int64_t* mask_local, const int64_t* ids_all,
                                        const int64_t* mask_all, int64_t L, const void* cond_local, const void* cond_all,
                                        int64_t row_bytes_cond, const int64_t* neg_text, const int64_t* neg_cond,
                                        int64_t bs, int64_t n_total, int64_t* ids_out, int64_t* mask_out, void* cond_out,
                                        vast_stream_t stream) {
  VAST_REQUIRE(bs >= 0 && n_total >= bs, VAST_ERR_INVALID, "gather_rows_concat3: bad sizes");
  if (bs == 0) return VAST_OK;
  if (cond_out) {
    VAST_REQUIRE(cond_local && cond_all && neg_cond, VAST_ERR_INVALID, "gather_rows_concat3: null cond pointer");
    VAST_REQUIRE(row_bytes_cond > 0 && row_bytes_cond % 16 == 0, VAST_ERR_UNSUPPORTED,
                 "gather_rows_concat3: S*H*elem_bytes (%lld) must be a multiple of 16", (long long)row_bytes_cond);
    VAST_REQUIRE(((reinterpret_cast<uintptr_t>(cond_local) | reinterpret_cast<uintptr_t>(cond_all) |
                   reinterpret_cast<uintptr_t>(cond_out)) & 15) == 0,
                 VAST_ERR_INVALID, "gather_rows_concat3: cond pointers must be 16-byte aligned");
    VAST_REQUIRE(2 * bs <= 65535, VAST_ERR_UNSUPPORTED, "gather_rows_concat3: bs too large");
    const int64_t row_vecs = row_bytes_cond / 16;
    const int64_t chunk = GATHER_THREADS * GATHER_UNROLL;
    int64_t gx = ceil_div64(row_vecs, chunk);
    if (gx > 64) gx = 64;
    dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(2 * bs));
    gather_cond3_kernel<<<grid, GATHER_THREADS, 0, stream>>>(static_cast<const uint4*>(cond_local),
                                                             static_cast<const uint4*>(cond_all), neg_cond, bs, n_total,
                                                             row_vecs, static_cast<uint4*>(cond_out));
    VAST_LAUNCH_OK("gather_cond3");
  }
  if (ids_out || mask_out) {
    VAST_REQUIRE(ids_out && mask_out && ids_local && mask_local && ids_all && mask_all && neg_text && L > 0,
                 VAST_ERR_INVALID, "gather_rows_concat3: null ids/mask pointer");
    gather_ids3_kernel<<<grid_for(3 * bs * L, 256), 256, 0, stream>>>(ids_local, mask_local, ids_all, mask_all, neg_text,
                                                                      bs, n_total, L, ids_out, mask_out);
    VAST_LAUNCH_OK("gather_ids3");
  }
  return VAST_OK;
}
