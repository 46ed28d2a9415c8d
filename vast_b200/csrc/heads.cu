// The two small heads either side of the contrastive path, as tcgen05 GEMMs with fused epilogues:
//
//  * vast_project_normalize -- Contra_head / fusion Linear + F.normalize (model/vast.py:221-279):
//      y = x . W^T + b ;  feat = y / max(|y|_2, eps)
//    The epilogue reads the accumulator tile TWICE from TMEM: pass 1 sums the squares of y = acc + b over its 128 rows x
//    256 columns; the column tiles of a row block exchange those partial sums through a ticket (they run side by
//    side); pass 2 writes the normalised feature once -- fp32, and the bf16 copy straight into the all-gather slot.
//    y never reaches HBM unnormalised: no separate normalise kernel, no cuBLAS call, no second launch.
//
//  * vast_match_head -- Match_head + softmax[:, 1] (model/general_module.py:34-42, model/vast.py:378):
//      h = GELU(cls . W1^T + b1) ;  z = W2 . LayerNorm(h) + b2 ;  score = softmax(z)[1]
//    LayerNorm followed by a 2-row Linear needs only four sums per row -- sum h, sum h^2, sum u0 h, sum u1 h with
//    u_c = W2[c] * gamma -- so the hidden activations never leave the accumulator tile:
//      z_c = (sum u_c h - mean sum u_c) / sqrt(var + eps) + (W2[c] . beta + b2[c]),  score = 1 / (1 + exp(z_0 - z_1)).
//
// Both take PACKED 16-bit operands (vast_sim_pack_operand: a plain bf16 cast, or bf16 splits of fp32 values whose
// leading cross products give fp32-grade results), so "bf16-in / fp32-accumulate" and "fp32" are the same kernels.
#include <cstdlib>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace vast {

// ------------------------------------------------------------------ projection + L2 normalise
struct EpiProj {
  struct Params {
    const float* bias;   // [N] or nullptr
    float* y;            // [M][ldy] out: normalised features (fp32)
    int64_t ldy;
    __nv_bfloat16* y16;  // [M][ld16] bf16 copy (e.g. the all-gather slot) or nullptr
    int64_t ld16;
    float* inv_norm;     // [M] 1 / max(|y|, eps) or nullptr
    float* sumsq;        // [M][slots] per-item partial sums of squares
    int slots;           // n_splits * 2
    int* tickets;        // [m_blocks] zero-initialised by the host before every launch
    int items_per_block; // work items covering one 128-row block (= column tiles)
    float eps;
    int N;
  };
  static constexpr bool kUnrollChunks = false;
  static constexpr int kAuxWarps = 0;
  static constexpr bool kSecondPass = true;  // pass 1 over the accumulator: sum of squares; pass 2: normalise + store
  const Params& p;
  float ss, inv;
  __device__ EpiProj(const Params& p_, uint8_t*) : p(p_) {}
  __device__ __forceinline__ void item_begin(const tc::ItemCtx&) {
    ss = 0.f;
    inv = 0.f;
  }
  __device__ __forceinline__ void prefetch(const tc::ItemCtx&, int) {}
  __device__ __forceinline__ void advance(const tc::ItemCtx&, int, bool) {}
  // pass 1: y = acc + bias, only its squares are kept
  __device__ __forceinline__ void chunk(const tc::ItemCtx& c, const uint32_t (&v)[32], int col0) {
    if (!c.row_valid || col0 >= c.N) return;
    if (col0 + 32 <= c.N && (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias + col0) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {  // the bias as 16-byte broadcast loads: 8 per chunk instead of 32
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
        const float o0 = __uint_as_float(v[i]) + b.x, o1 = __uint_as_float(v[i + 1]) + b.y;
        const float o2 = __uint_as_float(v[i + 2]) + b.z, o3 = __uint_as_float(v[i + 3]) + b.w;
        ss = fmaf(o0, o0, ss);
        ss = fmaf(o1, o1, ss);
        ss = fmaf(o2, o2, ss);
        ss = fmaf(o3, o3, ss);
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (col0 + i < c.N) {
        const float o = __uint_as_float(v[i]) + (p.bias != nullptr ? __ldg(p.bias + col0 + i) : 0.f);
        ss = fmaf(o, o, ss);
      }
  }
  // between the passes (one 256-column tile per item): publish this thread's partial sum of squares, take a ticket,
  // wait until all column tiles of the row block have published theirs (they run side by side on neighbouring CTA
  // pairs; a straggler of the next round never waits on this one), add the partials in slot order (the same sum in
  // every CTA: deterministic).  Only 4-byte partials cross CTAs; y itself never leaves the accumulator unnormalised.
  __device__ __forceinline__ void between(const tc::ItemCtx& c) {
    const bool owns = c.m_blk * tc::BM < c.M;  // (the odd CTA of the last pair may own no rows at all)
    if (c.row_valid) p.sumsq[static_cast<int64_t>(c.row) * p.slots + c.slot] = ss;
    __threadfence();
    asm volatile("bar.sync 2, 256;" ::: "memory");  // the 8 epilogue warps
    const int ew = (static_cast<int>(threadIdx.x) >> 5) - 2;
    if (ew == 0 && c.lane == 0 && owns) atomicAdd(p.tickets + c.m_blk, 1);
    if (!c.row_valid) return;
    int v;
    const int* t = p.tickets + c.m_blk;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(t) : "memory");
    const long long t0 = clock64();
    while (v < p.items_per_block) {
      __nanosleep(100);
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(t) : "memory");
      if (clock64() - t0 > 6000000000LL) __trap();
    }
    float tot = 0.f;
    for (int sl = 0; sl < p.slots; ++sl) tot += __ldcg(p.sumsq + static_cast<int64_t>(c.row) * p.slots + sl);
    inv = 1.0f / fmaxf(sqrtf(tot), p.eps);
  }
  // pass 2: the normalised feature, fp32 and bf16, written once
  __device__ __forceinline__ void chunk2(const tc::ItemCtx& c, const uint32_t (&v)[32], int col0) {
    if (!c.row_valid || col0 >= c.N) return;
    float* dst = p.y + static_cast<int64_t>(c.row) * p.ldy + col0;
    __nv_bfloat16* d16 = p.y16 != nullptr ? p.y16 + static_cast<int64_t>(c.row) * p.ld16 + col0 : nullptr;
    const bool vec = (col0 + 32 <= c.N) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
                     (d16 == nullptr || (reinterpret_cast<uintptr_t>(d16) & 15) == 0) &&
                     (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias + col0) & 15) == 0);
    const bool wide = ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) && (d16 == nullptr || (reinterpret_cast<uintptr_t>(d16) & 31) == 0);
    if (vec) {
      uint4 u16[4];
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        float o[8];
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (p.bias != nullptr) {
          b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
          b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i + 4));
        }
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (__uint_as_float(v[i + j]) + bb[j]) * inv;
        if (wide) {  // whole 32-byte sectors per instruction (row-per-thread stores)
          st_global_v8_f32(dst + i, o);
        } else {
          *reinterpret_cast<float4*>(dst + i) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(dst + i + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
        if (d16 != nullptr) {
          uint32_t* w = &u16[i >> 3].x;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
            w[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          if (!wide) *reinterpret_cast<uint4*>(d16 + i) = u16[i >> 3];
        }
      }
      if (wide && d16 != nullptr) {
        st_global_v8_b32(d16, u16[0], u16[1]);
        st_global_v8_b32(d16 + 16, u16[2], u16[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < c.N) {
          const float o = (__uint_as_float(v[i]) + (p.bias != nullptr ? __ldg(p.bias + col0 + i) : 0.f)) * inv;
          dst[i] = o;
          if (d16 != nullptr) d16[i] = __float2bfloat16_rn(o);
        }
    }
  }
  __device__ __forceinline__ void item_end(const tc::ItemCtx& c) {
    if (c.row_valid && c.n_split == 0 && c.half == 0 && p.inv_norm != nullptr) p.inv_norm[c.row] = inv;
  }
};

// ------------------------------------------------------------------ Match_head + softmax[:, 1]
struct EpiMatch {
  struct Params {
    const float* b1;  // [H]
    const float* u0;  // [H] W2[0] * gamma
    const float* u1;  // [H] W2[1] * gamma
    float U0, U1;     // sum u_c
    float v0, v1;     // W2[c] . beta + b2[c]
    float eps;
    float* score;     // [M] softmax(z)[1]
    float* logits;    // [M][2] or nullptr
  };
  static constexpr bool kUnrollChunks = false;
  static constexpr int kAuxWarps = 0;
  const Params& p;
  float s1, s2, t0, t1;
  __device__ EpiMatch(const Params& p_, uint8_t*) : p(p_) {}
  __device__ __forceinline__ void item_begin(const tc::ItemCtx&) { s1 = s2 = t0 = t1 = 0.f; }
  __device__ __forceinline__ void prefetch(const tc::ItemCtx&, int) {}
  __device__ __forceinline__ void advance(const tc::ItemCtx&, int, bool) {}
  __device__ __forceinline__ void chunk(const tc::ItemCtx& c, const uint32_t (&v)[32], int col0) {
    if (col0 >= c.N) return;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int col = col0 + i;
      if (col < c.N) {
        const float h = __uint_as_float(v[i]) + __ldg(p.b1 + col);
        const float g = h * 0.5f * (1.0f + erff(h * 0.70710678118654752f));  // the reference's erf GELU (general_module.py:13-17)
        s1 += g;
        s2 = fmaf(g, g, s2);
        t0 = fmaf(__ldg(p.u0 + col), g, t0);
        t1 = fmaf(__ldg(p.u1 + col), g, t1);
      }
    }
  }
  __device__ __forceinline__ void item_end(const tc::ItemCtx& c) {
    if (!c.row_valid) return;
    const float inv_h = 1.0f / static_cast<float>(c.N);
    const float mu = s1 * inv_h;
    const float var = fmaxf(s2 * inv_h - mu * mu, 0.f);
    const float rs = rsqrtf(var + p.eps);
    const float z0 = (t0 - mu * p.U0) * rs + p.v0, z1 = (t1 - mu * p.U1) * rs + p.v1;
    p.score[c.row] = 1.0f / (1.0f + expf(z0 - z1));
    if (p.logits != nullptr) {
      p.logits[2 * c.row] = z0;
      p.logits[2 * c.row + 1] = z1;
    }
  }
};

}  // namespace vast

using namespace vast;

extern "C" size_t vast_project_normalize_workspace_bytes(int64_t rows, int64_t dim_out) {
  if (rows <= 0 || dim_out <= 0) return 0;
  const int64_t slots = 2 * ceil_div64(dim_out, 256);
  return align_up(sizeof(float) * rows * slots, 256) + align_up(sizeof(int) * (ceil_div64(rows, tc::BM) + 1), 256);
}

extern "C" int vast_project_normalize(const void* x_op, const void* w_op, int64_t rows, int64_t dim_out, int64_t cols,
                                      const float* bias, float eps, float* y, int64_t ldy, void* y16, int64_t ld16,
                                      float* inv_norm, void* workspace, size_t workspace_bytes, vast_stream_t stream) {
  VAST_REQUIRE(x_op && w_op && y && workspace, VAST_ERR_INVALID, "project_normalize: null pointer");
  VAST_REQUIRE(rows > 0 && dim_out > 0 && cols > 0 && rows < (1 << 30) && dim_out <= 16384, VAST_ERR_INVALID,
               "project_normalize: bad sizes");
  VAST_REQUIRE(cols % 8 == 0, VAST_ERR_UNSUPPORTED, "project_normalize: operand width must be a multiple of 8");
  VAST_REQUIRE(ldy >= dim_out && (y16 == nullptr || ld16 >= dim_out), VAST_ERR_INVALID, "project_normalize: bad leading dimension");
  VAST_REQUIRE(workspace_bytes >= vast_project_normalize_workspace_bytes(rows, dim_out), VAST_ERR_WORKSPACE,
               "project_normalize: workspace too small");
  using Epi = EpiProj;
  tc::KernelParams<Epi::Params> P;
  memset(&P, 0, sizeof(P));
  const int cl = tc::pick_cluster((int)rows);
  tc::fill_shape(&P.g, 1, (int)rows, (int)dim_out, (int)cols, 256, 1, 1, false, cl);
  // one work item per (row-block group, 256-column tile): as many items as the shape offers, no K split
  P.g.n_splits = P.g.n_tiles;
  P.g.tiles_per_split = 1;
  P.g.num_items = P.g.m_groups * P.g.n_splits;
  int rc = tc::make_tmap_2d(&P.tmA[0], x_op, VAST_BF16, rows, cols, cols, tc::BM);
  if (rc) return rc;
  rc = tc::make_tmap_2d(&P.tmB[0], w_op, VAST_BF16, dim_out, cols, cols, 256 / cl);
  if (rc) return rc;
  const int slots = 2 * P.g.n_splits;
  char* ws = static_cast<char*>(workspace);
  float* sumsq = reinterpret_cast<float*>(ws);
  int* tickets = reinterpret_cast<int*>(ws + align_up(sizeof(float) * rows * slots, 256));
  VAST_CUDA_OK(cudaMemsetAsync(tickets, 0, sizeof(int) * P.g.m_blocks, stream));
  P.epi = {bias, y, ldy, static_cast<__nv_bfloat16*>(y16), ld16, inv_norm, sumsq, slots, tickets, P.g.n_splits, eps, (int)dim_out};
  return tc::launch_gemm<Epi, 256, 4, 8>(P, stream, "project_normalize_gemm", 16);
}

extern "C" int vast_match_head(const void* cls_op, const void* w1_op, int64_t rows, int64_t hidden, int64_t cols,
                               const float* b1, const float* u0, const float* u1, float sum_u0, float sum_u1, float v0,
                               float v1, float eps, float* score, float* logits, vast_stream_t stream) {
  VAST_REQUIRE(cls_op && w1_op && b1 && u0 && u1 && score, VAST_ERR_INVALID, "match_head: null pointer");
  VAST_REQUIRE(rows > 0 && hidden > 0 && cols > 0 && rows < (1 << 30) && hidden <= 16384, VAST_ERR_INVALID, "match_head: bad sizes");
  VAST_REQUIRE(cols % 8 == 0, VAST_ERR_UNSUPPORTED, "match_head: operand width must be a multiple of 8");
  using Epi = EpiMatch;
  tc::KernelParams<Epi::Params> P;
  memset(&P, 0, sizeof(P));
  const int cl = tc::pick_cluster((int)rows);
  tc::fill_shape(&P.g, 1, (int)rows, (int)hidden, (int)cols, 256, 1, 1, false, cl);  // one item per row-block group: all tiles
  int rc = tc::make_tmap_2d(&P.tmA[0], cls_op, VAST_BF16, rows, cols, cols, tc::BM);
  if (rc) return rc;
  rc = tc::make_tmap_2d(&P.tmB[0], w1_op, VAST_BF16, hidden, cols, cols, 256 / cl);
  if (rc) return rc;
  P.epi = {b1, u0, u1, sum_u0, sum_u1, v0, v1, eps, score, logits};
  return tc::launch_gemm<Epi, 256, 4, 4>(P, stream, "match_head_gemm");
}
