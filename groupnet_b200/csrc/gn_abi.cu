// C-ABI entry points of libgroupnet_b200.so that are not tied to one kernel file.
#include "gn_common.cuh"

namespace gn {
int stage_fwd_simt(const gn_stage_cfg* c, const gn_stage_weights* w, const float* h,
                   const float* H, const float* U, float* node_out, float* dist_out,
                   void* ws, size_t ws_bytes, cudaStream_t st, bool skip_edge_mlp);
size_t stage_workspace_bytes_simt(const gn_stage_cfg* c);
}  // namespace gn

extern "C" int gn_abi_version(void) { return GN_ABI_VERSION; }

extern "C" const char* gn_error_string(int code) {
  switch (code) {
    case GN_OK: return "ok";
    case GN_E_NULL: return "required pointer is NULL";
    case GN_E_SHAPE: return "unsupported or inconsistent shape";
    case GN_E_SCALE: return "selected index k out of range";
    case GN_E_WORKSPACE: return "workspace too small";
    case GN_E_PRECISION: return "unknown precision or path not built";
    case GN_E_ALIGN: return "pointer not 16-byte aligned";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown error";
}

extern "C" size_t gn_stage_workspace_bytes(const gn_stage_cfg* cfg) {
  if (!cfg) return 0;
  return gn::stage_workspace_bytes_simt(cfg);
}

extern "C" int gn_stage_launch_count(const gn_stage_cfg* cfg) {
  if (!cfg) return 0;
  return cfg->pairwise ? 5 : 6;
}

extern "C" int gn_stage_fwd(const gn_stage_cfg* cfg, const gn_stage_weights* w,
                            const float* h_in, const float* H, const float* U,
                            float* node_out, float* dist_out,
                            void* workspace, size_t workspace_bytes, gn_stream_t stream) {
  if (!cfg || !w || !h_in || !node_out || !workspace) return GN_E_NULL;
  if ((reinterpret_cast<uintptr_t>(h_in) | reinterpret_cast<uintptr_t>(workspace)) & 15) return GN_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (cfg->precision) {
    case GN_FP32:
      return gn::stage_fwd_simt(cfg, w, h_in, H, U, node_out, dist_out, workspace, workspace_bytes, st, false);
    default:
      return GN_E_PRECISION;
  }
}
