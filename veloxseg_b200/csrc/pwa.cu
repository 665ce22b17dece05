// placeholder until the PWA kernels land
#include "vx_kernels.h"
using namespace vx;
extern "C" int vx_pwa_saved_layout(const vx_pwa_desc*, vx_pwa_saved*) { set_error("pwa: not built"); return VX_ERR_UNSUPPORTED; }
extern "C" size_t vx_pwa_workspace(const vx_pwa_desc*) { return 0; }
extern "C" int vx_pwa_block_fwd(const vx_pwa_desc*, const void* const*, void* const*, void*, size_t, vx_stream_t) { set_error("pwa: not built"); return VX_ERR_UNSUPPORTED; }
extern "C" int vx_pwa_block_bwd(const vx_pwa_desc*, const void* const*, void* const*, void*, size_t, vx_stream_t) { set_error("pwa: not built"); return VX_ERR_UNSUPPORTED; }
extern "C" int vx_pwa_gather(const vx_pwa_desc*, int32_t, const void*, void*, void*, vx_stream_t) { set_error("pwa: not built"); return VX_ERR_UNSUPPORTED; }
