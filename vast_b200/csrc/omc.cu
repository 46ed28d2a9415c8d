// OMC / ITC contrastive step (model/vast.py:405-440 + autograd) as tensor-core GEMMs with fused
// epilogues.  Two problems are batched in every launch:
//   problem 0 "cond2t": rows = local feat_cond, columns = all feat_t
//   problem 1 "t2cond": rows = local feat_t,    columns = all feat_cond
//
//   K1 transpose_ksum     : B operands -> fp16 transposes (dQ GEMM operand) + column sums
//   K2 stats GEMM         : S tile in TMEM -> online (max, sum-exp, sum z, z_target) per row   [pass 1]
//   K3 stats_finalize     : merge partials -> lse, per-row CE terms, ksum
//   K4 prob GEMM          : S tile recomputed -> p = exp(z - lse): fp16 P tile (L2-resident
//                           workspace), sum p*z (for d temp), exponential-race hard negatives [pass 2]
//   K5 sample_finalize    : merge race partials -> negative indices, d temp row terms
//   K6 dQ GEMM            : dQ = P . K  (fp16 x fp16 -> fp32, split-K partials)
//   K7 grad_finalize      : dQ = (1/(2 bs tau)) (P.K - (eps/N) sum_j K_j - (1-eps) K_target)
//   K8 final_reduce       : loss, d temp (single block, fixed order -> deterministic)
//
// Why dQ is not fused FlashAttention-style into K4: the dQ accumulator of a 128-row block is
// 128 x D fp32; at D = 1024 that is 512 KB, twice the 256 KB of TMEM (512 columns x 128 lanes),
// so an output-stationary fused backward would have to recompute S once per 256-column slice of
// D (4x the S work).  Staging P in fp16 through L2 costs one extra write+read of bs x N x 2 B
// (32 MB per direction at bs = N = 4096, L2-resident on a 126 MB L2) and keeps S at exactly two
// evaluations.  The fp32 logits / log-softmax / gradient matrices are never materialised.
#include "common.cuh"
#include "gemm_tc.cuh"

namespace vast {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Philox4x32-10 (Salmon et al., SC'11); same generator family as torch's CUDA RNG.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// Exp(1) variate from a 32-bit word: v = (x + 0.5) 2^-32, E = -log1p(-v)  (small E, which decides
// the race, keeps full relative precision).
__device__ __forceinline__ float expo_from_bits(uint32_t x) {
  float v = fmaf(__uint2float_rn(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  v = fminf(v, 0.99999994f);
  const float small = v * fmaf(0.5f, v, 1.0f);
  const float big = -kLn2 * lg2_approx(1.0f - v);
  return v < 9.765625e-4f ? small : big;
}

// ------------------------------------------------------------------ pass 1 epilogue: row statistics
struct EpiStats {
  struct Params {
    float4* partial;  // [2][M][num_slots] (m2, l, sum s, s_target)
    int num_slots;
    float scale2;  // log2(e) / tau (used when temp_dev is null)
    const float* temp_dev;
    int tgt_offset;
  };
  static size_t smem_bytes() { return 0; }
  const Params& p;
  float m2, l, sz, zt, scale2;
  __device__ EpiStats(const Params& p_, uint8_t*) : p(p_) { scale2 = p.temp_dev ? kLog2e / __ldg(p.temp_dev) : p.scale2; }
  __device__ __forceinline__ void item_begin(const tc::ItemCtx&) {
    m2 = -INFINITY;
    l = 0.f;
    sz = 0.f;
    zt = 0.f;
  }
  __device__ __forceinline__ void chunk(const tc::ItemCtx& c, const uint32_t (&v)[32], int col0) {
    if (col0 >= c.N) return;
    const int nvalid = c.N - col0;  // >= 1; >= 32 for a full chunk
    float cm = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) cm = fmaxf(cm, __uint_as_float(v[i]));
    cm *= scale2;
    if (cm > m2) {
      l *= ex2_approx(m2 - cm);
      m2 = cm;
    }
    float ls = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i < nvalid) {
        const float s = __uint_as_float(v[i]);
        ls += ex2_approx(fmaf(s, scale2, -m2));
        ss += s;
      }
    }
    l += ls;
    sz += ss;
    const int t = p.tgt_offset + c.row - col0;
    if (static_cast<unsigned>(t) < 32u) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i == t) zt = __uint_as_float(v[i]);
    }
  }
  __device__ __forceinline__ void item_end(const tc::ItemCtx& c) {
    if (c.row_valid)
      p.partial[(static_cast<int64_t>(c.prob) * c.M + c.row) * p.num_slots + c.slot] = make_float4(m2, l, sz, zt);
  }
};

// ------------------------------------------------------------------ pass 2 epilogue: probabilities + race
struct EpiProb {
  struct Params {
    const float* lse2;  // [2][M] log2-domain row log-sum-exp
    __half* P;          // [2][M][ldp] softmax probabilities (dQ GEMM A operand) or nullptr
    int64_t ldp;
    float4* partial;  // [2][M][num_slots] (w_best, e_best, idx_best, sum p*s)
    int num_slots;
    float scale2;
    const float* temp_dev;
    float floor;
    int tgt_offset;
    int row_offset;  // global row of local row 0 (decorrelates ranks that share a seed)
    uint32_t seed_lo, seed_hi, off_lo, off_hi;
    const float* noise;  // [2][M][N] caller-supplied Exp(1) noise (debug) or nullptr
    int do_sample;
  };
  static size_t smem_bytes() { return 0; }
  const Params& p;
  float lse2, bw, be, pz, scale2;
  int bidx, tcol;
  __device__ EpiProb(const Params& p_, uint8_t*) : p(p_) { scale2 = p.temp_dev ? kLog2e / __ldg(p.temp_dev) : p.scale2; }
  __device__ __forceinline__ void item_begin(const tc::ItemCtx& c) {
    lse2 = c.row_valid ? p.lse2[static_cast<int64_t>(c.prob) * c.M + c.row] : 0.f;
    bw = -1.f;
    be = 1.f;
    bidx = -1;
    pz = 0.f;
    tcol = p.tgt_offset + c.row;
  }
  __device__ __forceinline__ void chunk(const tc::ItemCtx& c, const uint32_t (&v)[32], int col0) {
    if (col0 >= c.N) return;
    const int nvalid = c.N - col0;
    float pr[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float s = __uint_as_float(v[i]);
      const float q = (i < nvalid) ? ex2_approx(fmaf(s, scale2, -lse2)) : 0.f;
      pr[i] = q;
      pz = fmaf(q, s, pz);
    }
    if (p.P != nullptr && c.row_valid) {
      __half* dst = p.P + (static_cast<int64_t>(c.prob) * c.M + c.row) * p.ldp + col0;
      // The target column is excluded from the fp16 P matrix: its coefficient (p_iy - (1 - eps)) is a
      // cancellation and is applied in fp32 by the gradient finalize kernel.
      const int tq = tcol - col0;
      if (nvalid >= 32) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          uint32_t* w = &u.x;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __half2 h = __floats2half2_rn((i + 2 * j == tq) ? 0.f : pr[i + 2 * j],
                                                (i + 2 * j + 1 == tq) ? 0.f : pr[i + 2 * j + 1]);
            w[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(dst + i) = u;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nvalid) dst[i] = __float2half_rn(i == tq ? 0.f : pr[i]);
      }
    }
    if (p.do_sample) {
      const float* nz =
          p.noise ? p.noise + (static_cast<int64_t>(c.prob) * c.M + (c.row_valid ? c.row : 0)) * c.N + col0 : nullptr;
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        uint4 r = make_uint4(0, 0, 0, 0);
        if (nz == nullptr)
          r = philox4x32_10(make_uint4(static_cast<uint32_t>((col0 >> 2) + g4), static_cast<uint32_t>(p.row_offset + c.row),
                                       p.off_lo, (p.off_hi << 1) | static_cast<uint32_t>(c.prob)),
                            make_uint2(p.seed_lo, p.seed_hi));
        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = g4 * 4 + j;
          const int col = col0 + i;
          float e;
          if (nz != nullptr)
            e = (i < nvalid) ? nz[i] : 1.f;
          else
            e = expo_from_bits(rw[j]);
          float w = pr[i] + p.floor;
          if (col == tcol) w = 0.f;
          // argmax of w / e without the division; strict > keeps the first (lowest) index on ties
          if (i < nvalid && w * be > bw * e) {
            bw = w;
            be = e;
            bidx = col;
          }
        }
      }
    }
  }
  __device__ __forceinline__ void item_end(const tc::ItemCtx& c) {
    if (c.row_valid)
      p.partial[(static_cast<int64_t>(c.prob) * c.M + c.row) * p.num_slots + c.slot] =
          make_float4(bw, be, __int_as_float(bidx), pz);
  }
};

// ------------------------------------------------------------------ K1: transposes + column sums
// pack [N, 2D] bf16 -> KT [2][D][Npad] fp16 with KT[0] = feat_t_all^T, KT[1] = feat_cond_all^T, and
// ksum_partial [2][nrb][D] = per-256-row-block column sums (fixed order).
constexpr int TR_ROWS = 256;
__global__ void __launch_bounds__(256) transpose_ksum_kernel(const __nv_bfloat16* __restrict__ pack, int64_t n_total,
                                                            int64_t dim, int64_t npad, __half* __restrict__ kt,
                                                            float* __restrict__ ksum_partial, int nrb) {
  __shared__ __half tile[64][66];
  __shared__ float csum[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  // blockIdx.x = prob * ceil(D/64) + column block inside that problem's D columns (any D)
  const int nblk = static_cast<int>((dim + 63) / 64);
  const int prob = static_cast<int>(blockIdx.x) / nblk;
  const int64_t d0 = static_cast<int64_t>(static_cast<int>(blockIdx.x) - prob * nblk) * 64;
  const int64_t c0 = prob * dim + d0;  // column in [0, 2D) of the packed row
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * TR_ROWS;
  float s0 = 0.f, s1 = 0.f;
  for (int sub = 0; sub < TR_ROWS / 64; ++sub) {
    const int64_t rb = r0 + sub * 64;
    if (rb >= n_total) break;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = ty * 8 + i;
      float a = 0.f, b = 0.f;
      if (rb + r < n_total) {
        const int64_t col = c0 + 2 * tx;
        const __nv_bfloat16* src = pack + (rb + r) * 2 * dim + col;
        if (d0 + 2 * tx < dim) a = __bfloat162float(src[0]);
        if (d0 + 2 * tx + 1 < dim) b = __bfloat162float(src[1]);
      }
      s0 += a;
      s1 += b;
      tile[r][2 * tx] = __float2half_rn(a);
      tile[r][2 * tx + 1] = __float2half_rn(b);
    }
    __syncthreads();
    // write 64 d-rows x 64 n: each thread one (d, pair of n)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int d = ty * 8 + i;
      if (d0 + d < dim) {
        const int64_t n = rb + 2 * tx;
        __half* dst = kt + (static_cast<int64_t>(prob) * dim + d0 + d) * npad + n;
        if (n + 1 < npad)
          *reinterpret_cast<__half2*>(dst) = __halves2half2(tile[2 * tx][d], tile[2 * tx + 1][d]);
        else if (n < npad)
          dst[0] = tile[2 * tx][d];
      }
    }
    __syncthreads();
  }
  csum[ty][2 * tx] = s0;
  csum[ty][2 * tx + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += csum[i][threadIdx.x];
    if (d0 + threadIdx.x < dim)
      ksum_partial[(static_cast<int64_t>(prob) * nrb + blockIdx.y) * dim + d0 + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------ K3: merge pass-1 partials
__global__ void __launch_bounds__(128) omc_stats_finalize_kernel(const float4* __restrict__ partial, int slots, int M,
                                                                int N, float inv_tau, const float* __restrict__ temp_dev,
                                                                float eps_ls,
                                                                float* __restrict__ lse2, float* __restrict__ zt,
                                                                float* __restrict__ sz, float* __restrict__ rowce,
                                                                float* __restrict__ lse_out, int row_blocks,
                                                                const float* __restrict__ ksum_partial, int nrb, int D,
                                                                float* __restrict__ ksum) {
  if (static_cast<int>(blockIdx.x) >= row_blocks) {  // trailing blocks: ksum[p][d] = sum of row-block partials
    const int idx = (blockIdx.x - row_blocks) * 128 + threadIdx.x;
    if (idx < 2 * D) {
      const int p = idx / D, d = idx - p * D;
      float s = 0.f;
      for (int b = 0; b < nrb; ++b) s += ksum_partial[(static_cast<int64_t>(p) * nrb + b) * D + d];
      ksum[idx] = s;
    }
    return;
  }
  const int r = blockIdx.x * 128 + threadIdx.x;
  if (r >= 2 * M) return;
  if (temp_dev) inv_tau = 1.0f / __ldg(temp_dev);
  const float4* pp = partial + static_cast<int64_t>(r) * slots;
  float mm = -INFINITY;
  for (int s = 0; s < slots; ++s) mm = fmaxf(mm, pp[s].x);
  float l = 0.f, ssum = 0.f, t = 0.f;
  for (int s = 0; s < slots; ++s) {
    const float4 q = pp[s];
    if (q.y > 0.f) l += q.y * exp2f(q.x - mm);
    ssum += q.z;
    t += q.w;
  }
  const float l2 = mm + log2f(l);
  lse2[r] = l2;
  zt[r] = t;
  sz[r] = ssum;
  const float lse = l2 * kLn2;
  if (lse_out) lse_out[r] = lse;
  rowce[r] = lse - (1.f - eps_ls) * inv_tau * t - (eps_ls / static_cast<float>(N)) * inv_tau * ssum;
}

// ------------------------------------------------------------------ K5: merge race partials
__global__ void __launch_bounds__(128) omc_sample_finalize_kernel(const float4* __restrict__ partial, int slots, int M,
                                                                 int N, float inv_tau, const float* __restrict__ temp_dev,
                                                                 float eps_ls,
                                                                 const float* __restrict__ zt,
                                                                 const float* __restrict__ sz,
                                                                 int64_t* __restrict__ neg_idx,
                                                                 float* __restrict__ rowdt) {
  const int r = blockIdx.x * 128 + threadIdx.x;
  if (r >= 2 * M) return;
  if (temp_dev) inv_tau = 1.0f / __ldg(temp_dev);
  const float4* pp = partial + static_cast<int64_t>(r) * slots;
  float bw = -1.f, be = 1.f, pz = 0.f;
  int bidx = -1;
  for (int s = 0; s < slots; ++s) {
    const float4 q = pp[s];
    pz += q.w;
    const int idx = __float_as_int(q.z);
    if (idx < 0) continue;
    const float lhs = q.x * be, rhs = bw * q.y;
    if (lhs > rhs || (lhs == rhs && idx < bidx)) {
      bw = q.x;
      be = q.y;
      bidx = idx;
    }
  }
  if (neg_idx) neg_idx[r] = bidx;
  // d loss / d tau row term: -(1/tau) * sum_j (p_ij - y_ij) z_ij   (scaled by 1/(2 bs) in final_reduce)
  rowdt[r] = -inv_tau * inv_tau * (pz - (1.f - eps_ls) * zt[r] - (eps_ls / static_cast<float>(N)) * sz[r]);
}

// ------------------------------------------------------------------ K7: gradient assembly
__global__ void __launch_bounds__(256) omc_grad_finalize_kernel(const float* __restrict__ part, int ksplits,
                                                               int64_t split_stride, int M, int D,
                                                               const float* __restrict__ ksum,
                                                               const __nv_bfloat16* __restrict__ pack, int row_offset,
                                                               int N, float eps_ls, float inv_tau,
                                                               const float* __restrict__ temp_dev,
                                                               const float* __restrict__ lse2,
                                                               const float* __restrict__ zt,
                                                               float* __restrict__ grad_cond, float* __restrict__ grad_t) {
  const int64_t total = 2LL * M * D;
  if (temp_dev) inv_tau = 1.0f / __ldg(temp_dev);
  const float gs = inv_tau / (2.0f * M);
  const float scale2 = kLog2e * inv_tau;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const int row = static_cast<int>((i / D) % M);
    const int p = static_cast<int>(i / (static_cast<int64_t>(D) * M));
    float v = 0.f;
    for (int k = 0; k < ksplits; ++k) v += part[k * split_stride + i];
    v -= (eps_ls / static_cast<float>(N)) * ksum[p * D + d];
    // target row of the gathered operand: problem 0 -> feat_t_all, problem 1 -> feat_cond_all
    const float kt = __bfloat162float(pack[static_cast<int64_t>(row_offset + row) * 2 * D + p * D + d]);
    // target column in fp32: coefficient p_iy - (1 - eps)  (P holds 0 there)
    const float pt = exp2f(fmaf(zt[p * M + row], scale2, -lse2[p * M + row]));
    v += (pt - (1.f - eps_ls)) * kt;
    (p == 0 ? grad_cond : grad_t)[static_cast<int64_t>(row) * D + d] = gs * v;
  }
}

// ------------------------------------------------------------------ K8: scalar reductions (one block)
__global__ void __launch_bounds__(1024) omc_final_reduce_kernel(const float* __restrict__ rowce,
                                                               const float* __restrict__ rowdt, int rows2, float scale,
                                                               float* __restrict__ loss, float* __restrict__ grad_temp) {
  __shared__ float red[2][1024];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < rows2; i += 1024) {
    a += rowce[i];
    if (rowdt) b += rowdt[i];
  }
  red[0][threadIdx.x] = a;
  red[1][threadIdx.x] = b;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (static_cast<int>(threadIdx.x) < s) {
      red[0][threadIdx.x] += red[0][threadIdx.x + s];
      red[1][threadIdx.x] += red[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (loss) loss[0] = red[0][0] * scale;
    if (grad_temp) grad_temp[0] = red[1][0] * scale;
  }
}

// ------------------------------------------------------------------ host orchestration
struct OmcPlan {
  tc::GemmShape g_s;   // S GEMMs (pass 1 and pass 2)
  tc::GemmShape g_dq;  // dQ GEMM
  int bn_dq;
  int slots;
  int nrb;
  int64_t npad;
  // workspace offsets (bytes)
  size_t off_partial, off_lse2, off_zt, off_sz, off_rowce, off_rowdt, off_P, off_KT, off_ksump, off_ksum, off_dq, total;
};

static void omc_plan(OmcPlan* pl, int64_t bs, int64_t n_total, int64_t dim, int need_sample, int need_grad) {
  const int sms = device_sm_count();
  tc::fill_shape(&pl->g_s, 2, (int)bs, (int)n_total, (int)dim, 256, 1);
  tc::choose_splits(&pl->g_s, sms, 32, 1);
  pl->slots = pl->g_s.n_splits * 2;  // NE = 8 -> two column halves per split
  pl->npad = static_cast<int64_t>(align_up(static_cast<size_t>(n_total), 8));
  pl->nrb = ceil_div((int)n_total, TR_ROWS);
  pl->bn_dq = dim >= 256 ? 256 : 128;
  tc::fill_shape(&pl->g_dq, 2, (int)bs, (int)dim, (int)n_total, pl->bn_dq, 0);
  tc::choose_splits(&pl->g_dq, sms, 64, 4);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = align_up(off, 256);
    const size_t r = off;
    off += bytes;
    return r;
  };
  pl->off_partial = take(sizeof(float4) * 2 * bs * pl->slots);
  pl->off_lse2 = take(sizeof(float) * 2 * bs);
  pl->off_zt = take(sizeof(float) * 2 * bs);
  pl->off_sz = take(sizeof(float) * 2 * bs);
  pl->off_rowce = take(sizeof(float) * 2 * bs);
  pl->off_rowdt = take(sizeof(float) * 2 * bs);
  pl->off_P = pl->off_KT = pl->off_ksump = pl->off_ksum = pl->off_dq = 0;
  if (need_grad) {
    pl->off_P = take(sizeof(__half) * 2 * bs * pl->npad);
    pl->off_KT = take(sizeof(__half) * 2 * dim * pl->npad);
    pl->off_ksump = take(sizeof(float) * 2 * pl->nrb * dim);
    pl->off_ksum = take(sizeof(float) * 2 * dim);
    pl->off_dq = take(sizeof(float) * pl->g_dq.k_splits * 2 * bs * dim);
  }
  (void)need_sample;
  pl->total = align_up(off, 256);
}

template <int BN>
static int launch_dq(tc::KernelParams<tc::EpiStore::Params>& P, cudaStream_t stream) {
  return tc::launch_gemm<tc::EpiStore, BN, 4, 4>(P, stream, "omc_dq_gemm");
}

}  // namespace vast

using namespace vast;

extern "C" size_t vast_omc_workspace_bytes(int64_t bs, int64_t n_total, int64_t dim, int need_sample, int need_grad) {
  if (bs <= 0 || n_total <= 0 || dim <= 0) return 0;
  OmcPlan pl;
  omc_plan(&pl, bs, n_total, dim, need_sample, need_grad);
  return pl.total;
}

extern "C" int vast_omc_step(const void* pack, int64_t bs, int64_t n_total, int64_t dim, int64_t row_offset,
                             float contra_temp, const float* contra_temp_dev, float label_smoothing,
                             float weight_floor, uint64_t seed,
                             uint64_t offset, const float* debug_noise, float* loss, int64_t* neg_idx,
                             float* grad_cond, float* grad_t, float* grad_temp, float* lse, void* workspace,
                             size_t workspace_bytes, vast_stream_t stream) {
  VAST_REQUIRE(pack && loss && workspace, VAST_ERR_INVALID, "omc_step: null pointer");
  VAST_REQUIRE(bs > 0 && n_total >= bs && dim > 0, VAST_ERR_INVALID, "omc_step: bad sizes");
  VAST_REQUIRE(bs < (1 << 24) && n_total < (1 << 30) && dim <= 16384, VAST_ERR_UNSUPPORTED, "omc_step: sizes too large");
  VAST_REQUIRE(dim % 8 == 0, VAST_ERR_UNSUPPORTED, "omc_step: dim must be a multiple of 8 (got %lld)", (long long)dim);
  VAST_REQUIRE(row_offset >= 0 && row_offset + bs <= n_total, VAST_ERR_INVALID, "omc_step: local rows outside [0, n_total)");
  VAST_REQUIRE(contra_temp_dev != nullptr || contra_temp > 0.f, VAST_ERR_INVALID, "omc_step: contra_temp must be positive");
  const bool need_grad = grad_cond || grad_t || grad_temp;
  VAST_REQUIRE(!need_grad || (grad_cond && grad_t && grad_temp), VAST_ERR_INVALID,
               "omc_step: give all of grad_cond, grad_t, grad_temp or none");
  const bool need_sample = neg_idx != nullptr;
  OmcPlan pl;
  omc_plan(&pl, bs, n_total, dim, need_sample, need_grad);
  VAST_REQUIRE(workspace_bytes >= pl.total, VAST_ERR_WORKSPACE, "omc_step: workspace %zu < required %zu", workspace_bytes, pl.total);
  VAST_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, VAST_ERR_INVALID, "omc_step: workspace must be 256-byte aligned");

  char* ws = static_cast<char*>(workspace);
  float4* partial = reinterpret_cast<float4*>(ws + pl.off_partial);
  float* lse2 = reinterpret_cast<float*>(ws + pl.off_lse2);
  float* zt = reinterpret_cast<float*>(ws + pl.off_zt);
  float* sz = reinterpret_cast<float*>(ws + pl.off_sz);
  float* rowce = reinterpret_cast<float*>(ws + pl.off_rowce);
  float* rowdt = reinterpret_cast<float*>(ws + pl.off_rowdt);
  __half* Pbuf = need_grad ? reinterpret_cast<__half*>(ws + pl.off_P) : nullptr;
  __half* KT = need_grad ? reinterpret_cast<__half*>(ws + pl.off_KT) : nullptr;
  float* ksump = need_grad ? reinterpret_cast<float*>(ws + pl.off_ksump) : nullptr;
  float* ksum = need_grad ? reinterpret_cast<float*>(ws + pl.off_ksum) : nullptr;
  float* dqpart = need_grad ? reinterpret_cast<float*>(ws + pl.off_dq) : nullptr;

  const auto* pk = static_cast<const __nv_bfloat16*>(pack);
  const float inv_tau = contra_temp_dev ? 0.f : 1.0f / contra_temp;  // device pointer wins (no host sync)
  const int M = static_cast<int>(bs), N = static_cast<int>(n_total), D = static_cast<int>(dim);
  const int row_blocks = ceil_div(2 * M, 128);
  int rc;

  // K1
  if (need_grad) {
    dim3 grid(static_cast<unsigned>(2 * ceil_div(D, 64)), static_cast<unsigned>(pl.nrb));
    VAST_TIMED(stream, "transpose_ksum", (transpose_ksum_kernel<<<grid, 256, 0, stream>>>(pk, n_total, dim, pl.npad, KT, ksump, pl.nrb)));
    VAST_LAUNCH_OK("transpose_ksum");
  }

  // tensor maps of the S GEMMs: A = local rows, B = all rows, both strided views of `pack`
  CUtensorMap tmA[2], tmB[2];
  rc = tc::make_tmap_2d(&tmA[0], pk + row_offset * 2 * dim + dim, VAST_BF16, bs, dim, 2 * dim, tc::BM);  // local cond
  if (rc) return rc;
  rc = tc::make_tmap_2d(&tmB[0], pk, VAST_BF16, n_total, dim, 2 * dim, 256);  // all t
  if (rc) return rc;
  rc = tc::make_tmap_2d(&tmA[1], pk + row_offset * 2 * dim, VAST_BF16, bs, dim, 2 * dim, tc::BM);  // local t
  if (rc) return rc;
  rc = tc::make_tmap_2d(&tmB[1], pk + dim, VAST_BF16, n_total, dim, 2 * dim, 256);  // all cond
  if (rc) return rc;

  // K2: pass 1
  {
    tc::KernelParams<EpiStats::Params> P;
    memset(&P, 0, sizeof(P));
    P.g = pl.g_s;
    for (int i = 0; i < 2; ++i) {
      P.tmA[i] = tmA[i];
      P.tmB[i] = tmB[i];
    }
    P.epi = {partial, pl.slots, kLog2e * inv_tau, contra_temp_dev, static_cast<int>(row_offset)};
    rc = tc::launch_gemm<EpiStats, 256, 4, 8>(P, stream, "omc_stats_gemm");
    if (rc) return rc;
  }
  // K3
  {
    const int kblocks = need_grad ? ceil_div(2 * D, 128) : 0;
    VAST_TIMED(stream, "omc_stats_finalize",
               (omc_stats_finalize_kernel<<<row_blocks + kblocks, 128, 0, stream>>>(
                   partial, pl.slots, M, N, inv_tau, contra_temp_dev, label_smoothing, lse2, zt, sz, rowce, lse, row_blocks,
                   ksump, pl.nrb, D, ksum)));
    VAST_LAUNCH_OK("omc_stats_finalize");
  }
  // K4 + K5: pass 2
  if (need_sample || need_grad) {
    tc::KernelParams<EpiProb::Params> P;
    memset(&P, 0, sizeof(P));
    P.g = pl.g_s;
    for (int i = 0; i < 2; ++i) {
      P.tmA[i] = tmA[i];
      P.tmB[i] = tmB[i];
    }
    P.epi.lse2 = lse2;
    P.epi.P = Pbuf;
    P.epi.ldp = pl.npad;
    P.epi.partial = partial;
    P.epi.num_slots = pl.slots;
    P.epi.scale2 = kLog2e * inv_tau;
    P.epi.temp_dev = contra_temp_dev;
    P.epi.floor = weight_floor;
    P.epi.tgt_offset = static_cast<int>(row_offset);
    P.epi.row_offset = static_cast<int>(row_offset);
    P.epi.seed_lo = static_cast<uint32_t>(seed);
    P.epi.seed_hi = static_cast<uint32_t>(seed >> 32);
    P.epi.off_lo = static_cast<uint32_t>(offset);
    P.epi.off_hi = static_cast<uint32_t>(offset >> 32);
    P.epi.noise = debug_noise;
    P.epi.do_sample = need_sample ? 1 : 0;
    rc = tc::launch_gemm<EpiProb, 256, 4, 8>(P, stream, "omc_prob_gemm");
    if (rc) return rc;
    VAST_TIMED(stream, "omc_sample_finalize",
               (omc_sample_finalize_kernel<<<row_blocks, 128, 0, stream>>>(partial, pl.slots, M, N, inv_tau, contra_temp_dev,
                                                                           label_smoothing, zt, sz, neg_idx, rowdt)));
    VAST_LAUNCH_OK("omc_sample_finalize");
  }
  // K6 + K7: dQ
  if (need_grad) {
    tc::KernelParams<tc::EpiStore::Params> P;
    memset(&P, 0, sizeof(P));
    P.g = pl.g_dq;
    for (int i = 0; i < 2; ++i) {
      rc = tc::make_tmap_2d(&P.tmA[i], Pbuf + static_cast<int64_t>(i) * bs * pl.npad, VAST_F16, bs, n_total, pl.npad, tc::BM);
      if (rc) return rc;
      rc = tc::make_tmap_2d(&P.tmB[i], KT + static_cast<int64_t>(i) * dim * pl.npad, VAST_F16, dim, n_total, pl.npad, pl.bn_dq);
      if (rc) return rc;
    }
    P.epi = {dqpart, dim, bs * dim, 2 * bs * dim, 1.0f};
    rc = pl.bn_dq == 256 ? launch_dq<256>(P, stream) : launch_dq<128>(P, stream);
    if (rc) return rc;
    const int64_t total = 2LL * M * D;
    int64_t gb = ceil_div64(total, 256 * 4);
    const int64_t cap = static_cast<int64_t>(device_sm_count()) * 8;
    if (gb > cap) gb = cap;
    VAST_TIMED(stream, "omc_grad_finalize",
               (omc_grad_finalize_kernel<<<static_cast<unsigned>(gb), 256, 0, stream>>>(
                   dqpart, pl.g_dq.k_splits, 2 * bs * dim, M, D, ksum, pk, static_cast<int>(row_offset), N, label_smoothing,
                   inv_tau, contra_temp_dev, lse2, zt, grad_cond, grad_t)));
    VAST_LAUNCH_OK("omc_grad_finalize");
  }
  // K8
  VAST_TIMED(stream, "omc_final_reduce",
             (omc_final_reduce_kernel<<<1, 1024, 0, stream>>>(rowce, need_grad ? rowdt : nullptr, 2 * M, 1.0f / (2.0f * M), loss,
                                                             need_grad ? grad_temp : nullptr)));
  VAST_LAUNCH_OK("omc_final_reduce");
  return VAST_OK;
}
