// placeholder (replaced next): retrieval entry points
#include "common.cuh"
using namespace vast;
#define NOTYET(name) do { set_last_error(name ": not implemented yet"); return VAST_ERR_UNSUPPORTED; } while (0)
extern "C" {
int64_t vast_sim_operand_cols(int64_t dim, int mode) { return mode == VAST_SIM_FP32X3 ? 6 * dim : dim; }
int vast_sim_pack_operand(const void*, int, int64_t, int64_t, int64_t, int, int, void*, vast_stream_t) { NOTYET("sim_pack_operand"); }
size_t vast_sim_topk_workspace_bytes(int64_t, int64_t, int64_t, int64_t) { return 0; }
int vast_sim_topk(const void*, const void*, int64_t, int64_t, int64_t, int64_t, int64_t, uint64_t*, void*, size_t, vast_stream_t) { NOTYET("sim_topk"); }
int vast_topk_merge(const uint64_t*, int64_t, int64_t, int64_t, int64_t, uint64_t*, vast_stream_t) { NOTYET("topk_merge"); }
int vast_topk_unpack(const uint64_t*, int64_t, float*, int32_t*, vast_stream_t) { NOTYET("topk_unpack"); }
int vast_rescore_f64(const float*, int64_t, const float*, int64_t, int64_t, int64_t, int32_t*, int64_t, int64_t, double*, vast_stream_t) { NOTYET("rescore_f64"); }
int vast_exact_topk_rows(const float*, int64_t, const float*, int64_t, int64_t, int64_t, const int32_t*, int64_t, int64_t, int64_t, int32_t*, double*, vast_stream_t) { NOTYET("exact_topk_rows"); }
int vast_dense_topk(const float*, int64_t, int64_t, int64_t, int64_t, int, int32_t*, float*, vast_stream_t) { NOTYET("dense_topk"); }
int vast_dense_rank_of_gt(const float*, int64_t, int64_t, int64_t, int, const int32_t*, int64_t, int32_t*, vast_stream_t) { NOTYET("dense_rank_of_gt"); }
int vast_bucket_by_video(const int32_t*, const int32_t*, int64_t, int64_t, int32_t*, int32_t*, void*, size_t, vast_stream_t) { NOTYET("bucket_by_video"); }
int vast_scatter_scores(const int32_t*, const int32_t*, const float*, int64_t, float*, int64_t, vast_stream_t) { NOTYET("scatter_scores"); }
}
