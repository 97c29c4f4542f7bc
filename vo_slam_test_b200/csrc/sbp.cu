// sbp.cu -- placeholder, replaced by the grid / projection-search kernels.
#include "common.cuh"
using namespace orbx;
extern "C" {
int orbx_grid_build(const orbx_keypoint*, int, float, float, float, float, int32_t*, int32_t*, int) { set_error("not built yet"); return ORBX_ERR_ARG; }
int orbx_search_by_projection_frame(const orbx_frame_view*, const orbx_sbp_frame_points*, float, float, int, int, int, int32_t*, int*, int) { set_error("not built yet"); return ORBX_ERR_ARG; }
int orbx_search_by_projection_local(const orbx_frame_view*, const orbx_sbp_local_points*, float, float, int32_t*, int*, int) { set_error("not built yet"); return ORBX_ERR_ARG; }
}
