// Host-side runtime pieces shared by every op: last-error string, SM count, TMA tensor-map
// encoding (driver entry point fetched through the runtime, no link-time libcuda dependency),
// work-split heuristics.
#include <cstdarg>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace vast {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VAST_NO_PDL");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

struct KTimer {
  static constexpr int MAX = 8192;
  bool enabled = false, created = false;
  cudaEvent_t ev[MAX][2];
  char names[MAX][48];
  int n = 0;
};
static KTimer g_timer;

int timing_begin(const char* name, cudaStream_t stream) {
  KTimer& t = g_timer;
  if (!t.enabled || t.n >= KTimer::MAX) return -1;
  const int slot = t.n++;
  strncpy(t.names[slot], name, 47);
  t.names[slot][47] = 0;
  cudaEventRecord(t.ev[slot][0], stream);
  return slot;
}
void timing_end(int slot, cudaStream_t stream) {
  if (slot >= 0) cudaEventRecord(g_timer.ev[slot][1], stream);
}

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* map, const void* ptr, int dtype, int64_t rows, int64_t cols, int64_t ld_elems,
                 int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  VAST_REQUIRE(fn != nullptr, VAST_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  VAST_REQUIRE(dtype == VAST_BF16 || dtype == VAST_F16, VAST_ERR_UNSUPPORTED, "tensor map: 16-bit operands only");
  VAST_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, VAST_ERR_INVALID, "operand pointer must be 16-byte aligned");
  VAST_REQUIRE(ld_elems % 8 == 0 && ld_elems >= cols, VAST_ERR_INVALID,
               "operand leading dimension (%lld) must be a multiple of 8 and >= cols (%lld)", (long long)ld_elems,
               (long long)cols);
  VAST_REQUIRE(rows > 0 && cols > 0 && box_rows > 0 && box_rows <= 256, VAST_ERR_INVALID, "bad tensor map shape");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype == VAST_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VAST_REQUIRE(r == CUDA_SUCCESS, VAST_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VAST_OK;
}

void choose_splits(GemmShape* g, int sm_count, int max_n_splits, int max_k_splits, double item_overhead) {
  const int row_blocks = g->num_problems * g->m_groups;
  sm_count = sm_count / (g->cl > 0 ? g->cl : 1);  // persistent grid = one cluster per `cl` SMs
  // cost model: waves * (work per item) with a small per-item flush overhead (in k-block units)
  double best = 1e300;
  int best_ns = 1, best_ks = 1;
  const int ns_hi = g->n_tiles < max_n_splits ? g->n_tiles : max_n_splits;
  const int ks_hi = g->k_blocks < max_k_splits ? g->k_blocks : max_k_splits;
  for (int ks = 1; ks <= ks_hi; ++ks) {
    const int kbps = ceil_div(g->k_blocks, ks);
    const int ks_eff = ceil_div(g->k_blocks, kbps);
    if (ks_eff != ks) continue;
    for (int ns = 1; ns <= ns_hi; ++ns) {
      const int tps = ceil_div(g->n_tiles, ns);
      const int ns_eff = ceil_div(g->n_tiles, tps);
      if (ns_eff != ns) continue;
      const long long items = 1LL * row_blocks * ns * ks;
      const long long waves = (items + sm_count - 1) / sm_count;
      const double per_item = static_cast<double>(tps) * kbps + item_overhead + (ks > 1 ? 4.0 : 0.0);
      const double cost = waves * per_item;
      if (cost < best * 0.999) {
        best = cost;
        best_ns = ns;
        best_ks = ks;
      }
    }
  }
  g->n_splits = best_ns;
  g->tiles_per_split = ceil_div(g->n_tiles, best_ns);
  g->k_splits = best_ks;
  g->kb_per_split = ceil_div(g->k_blocks, best_ks);
  g->num_items = row_blocks * best_ns * best_ks;
}

}  // namespace tc
}  // namespace vast

extern "C" {

const char* vast_last_error_string(void) { return vast::g_last_error; }

int vast_version(void) { return VAST_B200_VERSION; }

int vast_sm_count(void) { return vast::device_sm_count(); }

int vast_timing_enable(int on) {
  vast::KTimer& t = vast::g_timer;
  if (on && !t.created) {
    for (int i = 0; i < vast::KTimer::MAX; ++i)
      for (int j = 0; j < 2; ++j)
        if (cudaEventCreate(&t.ev[i][j]) != cudaSuccess) {
          vast::set_last_error("timing_enable: cudaEventCreate failed");
          return VAST_ERR_CUDA;
        }
    t.created = true;
  }
  t.enabled = on != 0;
  t.n = 0;
  return VAST_OK;
}

int vast_timing_read(float* ms_out, char* names_out, int max_entries) {
  vast::KTimer& t = vast::g_timer;
  int n = t.n < max_entries ? t.n : max_entries;
  for (int i = 0; i < n; ++i) {
    if (cudaEventSynchronize(t.ev[i][1]) != cudaSuccess || cudaEventElapsedTime(&ms_out[i], t.ev[i][0], t.ev[i][1]) != cudaSuccess) {
      vast::set_last_error("timing_read: event query failed");
      return VAST_ERR_CUDA;
    }
    if (names_out) memcpy(names_out + 48 * i, t.names[i], 48);
  }
  t.n = 0;
  return n;
}

}  // extern "C"
