// cross_encoder.cu -- K5 placeholder (replaced below in this round)
#include "common.cuh"
struct CeModel { int dummy; };
void ce_model_free(CeModel* m) { delete m; }
extern "C" {
int sb_ce_load(sb_ctx*, const float*, int64_t, const sb_ce_config*) { sb_set_error("cross-encoder not built"); return SB_ERR_UNSUPPORTED; }
int sb_ce_score(sb_ctx*, const int32_t*, const int32_t*, const int32_t*, int32_t, int32_t, float*, float*) { sb_set_error("cross-encoder not built"); return SB_ERR_UNSUPPORTED; }
int sb_ce_score_dev(sb_ctx*, const int32_t*, const int32_t*, const int32_t*, int32_t, int32_t, float*, float*, void*) { sb_set_error("cross-encoder not built"); return SB_ERR_UNSUPPORTED; }
}
