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
  tc::fill_shape(&g, 1, (int)M, (int)N, (int)K, 256, 1);
  tc::choose_splits(&g, device_sm_count(), 64, 8);
  return g.k_splits > 1 ? static_cast<size_t>(g.k_splits) * M * N * sizeof(float) + 256 : 0;
}

extern "C" int vast_gemm_nt(const void* A, int64_t lda, const void* B, int64_t ldb, int dtype, int64_t M, int64_t N,
                            int64_t K, float alpha, float* C, int64_t ldc, void* workspace, size_t workspace_bytes,
                            vast_stream_t stream) {
  VAST_REQUIRE(A && B && C, VAST_ERR_INVALID, "gemm_nt: null pointer");
  VAST_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1 << 30) && N < (1 << 30) && K < (1 << 30), VAST_ERR_INVALID,
               "gemm_nt: bad sizes");
  VAST_REQUIRE(dtype == VAST_BF16 || dtype == VAST_F16, VAST_ERR_UNSUPPORTED, "gemm_nt: bf16/f16 only");
  using Epi = tc::EpiStore;
  tc::KernelParams<Epi::Params> P;
  memset(&P, 0, sizeof(P));
  tc::fill_shape(&P.g, 1, (int)M, (int)N, (int)K, 256, dtype == VAST_BF16 ? 1 : 0);
  tc::choose_splits(&P.g, device_sm_count(), 64, 8);
  int rc = tc::make_tmap_2d(&P.tmA[0], A, dtype, M, K, lda, tc::BM);
  if (rc) return rc;
  rc = tc::make_tmap_2d(&P.tmB[0], B, dtype, N, K, ldb, 256);
  if (rc) return rc;
  if (P.g.k_splits > 1) {
    Workspace ws(workspace, workspace_bytes);
    float* part = ws.take<float>(static_cast<size_t>(P.g.k_splits) * M * N);
    VAST_REQUIRE(workspace && ws.ok(), VAST_ERR_WORKSPACE, "gemm_nt: workspace too small");
    P.epi = {part, N, 0, M * N, alpha};
    rc = tc::launch_gemm<Epi, 256, 4, 4>(P, stream, "gemm_nt");
    if (rc) return rc;
    const int64_t total = M * N;
    reduce_ksplit_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(part, M * N, P.g.k_splits, C, ldc, M, N);
    VAST_LAUNCH_OK("reduce_ksplit");
    return VAST_OK;
  }
  P.epi = {C, ldc, 0, 0, alpha};
  return tc::launch_gemm<Epi, 256, 4, 4>(P, stream, "gemm_nt");
}
