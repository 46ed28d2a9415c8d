// vast_gemm_nt: the tensor-core mainloop with the plain-store epilogue, exposed for unit tests
// (it validates TMA + UMMA descriptors + TMEM plumbing against a reference matmul) and used as
// the building block of the dQ GEMM.
#include "common.cuh"
#include "gemm_tc.cuh"

namespace vast {

__global__ void reduce_ksplit_kernel(const float* __restrict__ part, int64_t split_stride, int splits,
                                     float* __restrict__ out, int64_t ldc, int64_t M, int64_t N) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= M * N) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[k * split_stride + i];  // fixed order: deterministic
  out[(i / N) * ldc + (i % N)] = s;
}

}  // namespace vast

using namespace vast;

extern "C" size_t vast_gemm_nt_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  tc::GemmShape g;
  tc::fill_shape(&g, 1, (int)M, (int)N, (int)K, 256, 1, 1, false, tc::pick_cluster((int)M));
  tc::choose_splits(&g, device_sm_count(), 64, 8);
  return g.k_splits > 1 ? static_cast<size_t>(g.k_splits) * M * N * sizeof(float) + 256 : 0;
}

template <bool B_MN>
static int gemm_impl(const void* A, int64_t lda, int dtype_a, const void* B, int64_t ldb, int dtype_b, int64_t M, int64_t N,
                     int64_t K, float alpha, float* C, int64_t ldc, void* workspace, size_t workspace_bytes,
                     vast_stream_t stream, const char* name) {
  VAST_REQUIRE(A && B && C, VAST_ERR_INVALID, "%s: null pointer", name);
  VAST_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1 << 30) && N < (1 << 30) && K < (1 << 30), VAST_ERR_INVALID,
               "%s: bad sizes", name);
  VAST_REQUIRE((dtype_a == VAST_BF16 || dtype_a == VAST_F16) && dtype_b == dtype_a, VAST_ERR_UNSUPPORTED,
               "%s: both operands bf16 or both f16 (tcgen05 kind::f16 faults on mixed formats)", name);
  using Epi = tc::EpiStore;
  tc::KernelParams<Epi::Params> P;
  memset(&P, 0, sizeof(P));
  const int cl = tc::pick_cluster((int)M);
  tc::fill_shape(&P.g, 1, (int)M, (int)N, (int)K, 256, dtype_a == VAST_BF16 ? 1 : 0, dtype_b == VAST_BF16 ? 1 : 0, B_MN, cl);
  tc::choose_splits(&P.g, device_sm_count(), 64, 8);
  int rc = tc::make_tmap_2d(&P.tmA[0], A, dtype_a, M, K, lda, tc::BM);
  if (rc) return rc;
  rc = B_MN ? tc::make_tmap_2d(&P.tmB[0], B, dtype_b, K, N, ldb, tc::BK) : tc::make_tmap_2d(&P.tmB[0], B, dtype_b, N, K, ldb, 256 / cl);
  if (rc) return rc;
  if (const char* e = getenv("VAST_GEMM_PROBE")) {
    if (atoi(e) == 7 && cl == 2) {  // cluster split-K: one single-tile item per cluster of two CTA pairs
      tc::fill_shape(&P.g, 1, (int)M, (int)N, (int)K, 256, dtype_a == VAST_BF16 ? 1 : 0, dtype_b == VAST_BF16 ? 1 : 0, B_MN, 2);
      P.g.n_splits = P.g.n_tiles;
      P.g.tiles_per_split = 1;
      P.g.num_items = P.g.m_groups * P.g.n_tiles;
      P.epi = {C, ldc, 0, 0, alpha};
      return tc::launch_gemm_cl<Epi, 256, 6, 4, B_MN, 2, 0, 1, 0, 2>(P, stream, name, 0);
    }
  }
  if (P.g.k_splits > 1) {
    Workspace ws(workspace, workspace_bytes);
    float* part = ws.take<float>(static_cast<size_t>(P.g.k_splits) * M * N);
    VAST_REQUIRE(workspace && ws.ok(), VAST_ERR_WORKSPACE, "%s: workspace too small", name);
    P.epi = {part, N, 0, M * N, alpha};
    rc = tc::launch_gemm<Epi, 256, 4, 4, B_MN>(P, stream, name);
    if (rc) return rc;
    const int64_t total = M * N;
    reduce_ksplit_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(part, M * N, P.g.k_splits, C, ldc, M, N);
    VAST_LAUNCH_OK("reduce_ksplit");
    return VAST_OK;
  }
  P.epi = {C, ldc, 0, 0, alpha};
  // diagnostics (scripts/gemm_probe.py): VAST_GEMM_PROBE = 1 tensor pipe alone, 2 operand ingest alone, 3 clusters of two
  // pairs sharing A, 4 = 3 + 1, 5 = 3 + 2; 7 = cluster split-K (single wave of single-tile items)
  if (const char* e = getenv("VAST_GEMM_PROBE")) {
    const int mode = atoi(e);
    if (cl == 2 && mode >= 1 && mode <= 6) {
      if (mode >= 3 && mode <= 5) {
        VAST_REQUIRE(P.g.n_tiles % 2 == 0, VAST_ERR_INVALID, "%s: probe needs an even tile count", name);
        P.g.n_splits = P.g.n_tiles;  // one tile per item
        P.g.tiles_per_split = 1;
        P.g.num_items = P.g.m_groups * P.g.n_tiles;
        rc = tc::make_tmap_2d(&P.tmA[0], A, dtype_a, M, K, lda, 64);
        if (rc) return rc;
      }
      switch (mode) {
        case 1: return tc::launch_gemm_cl<Epi, 256, 6, 4, B_MN, 2, 0, 1, 1>(P, stream, name, 0);
        case 2: return tc::launch_gemm_cl<Epi, 256, 6, 4, B_MN, 2, 0, 1, 2>(P, stream, name, 0);
        case 3: return tc::launch_gemm_cl<Epi, 256, 6, 4, B_MN, 2, 0, 2, 0>(P, stream, name, 0);
        case 4: return tc::launch_gemm_cl<Epi, 256, 6, 4, B_MN, 2, 0, 2, 1>(P, stream, name, 0);
        case 5: return tc::launch_gemm_cl<Epi, 256, 6, 4, B_MN, 2, 0, 2, 2>(P, stream, name, 0);
        default: break;
      }
    }
  }
  return tc::launch_gemm<Epi, 256, 4, 4, B_MN>(P, stream, name);
}

// C[m, n] = alpha * sum_k A[m, k] * B[n, k]   (both operands K-major)
extern "C" int vast_gemm_nt(const void* A, int64_t lda, const void* B, int64_t ldb, int dtype, int64_t M, int64_t N,
                            int64_t K, float alpha, float* C, int64_t ldc, void* workspace, size_t workspace_bytes,
                            vast_stream_t stream) {
  return gemm_impl<false>(A, lda, dtype, B, ldb, dtype, M, N, K, alpha, C, ldc, workspace, workspace_bytes, stream, "gemm_nt");
}

// C[m, n] = alpha * sum_k A[m, k] * B[k, n]   (B row-major = MN-major operand, consumed without a transpose;
// A and B share one 16-bit format)
extern "C" int vast_gemm_nn(const void* A, int64_t lda, int dtype_a, const void* B, int64_t ldb, int dtype_b, int64_t M,
                            int64_t N, int64_t K, float alpha, float* C, int64_t ldc, void* workspace,
                            size_t workspace_bytes, vast_stream_t stream) {
  return gemm_impl<true>(A, lda, dtype_a, B, ldb, dtype_b, M, N, K, alpha, C, ldc, workspace, workspace_bytes, stream, "gemm_nn");
}
