// C-ABI entry points of libgroupnet_b200.so that are not tied to one kernel file.
#include "gn_common.cuh"

#include "gn_stage.h"

// ---------------------------------------------------------------------------
// profiling hook: brackets every kernel launch with CUDA events on the launch
// stream while enabled; bench.py uses it for the live per-kernel roofline.
// ---------------------------------------------------------------------------
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
namespace gn {
struct ProfRec { const char* name; cudaEvent_t a, b; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

ProfScope::ProfScope(const char* name, cudaStream_t st) : name_(name), st_(st), rec_(nullptr) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec* r = new ProfRec{name, nullptr, nullptr};
  cudaEventCreate(&r->a);
  cudaEventCreate(&r->b);
  cudaEventRecord(r->a, st);
  rec_ = r;
}
ProfScope::~ProfScope() {
  if (!rec_) return;
  ProfRec* r = static_cast<ProfRec*>(rec_);
  cudaEventRecord(r->b, st_);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(*r);
  delete r;
}
}  // namespace gn

extern "C" void gn_profile_enable(int on) { gn::g_prof_on.store(on ? 1 : 0); }

extern "C" int gn_profile_collect(char* names, int names_len, float* total_ms, int* counts, int max_entries) {
  std::lock_guard<std::mutex> lk(gn::g_prof_mu);
  std::vector<std::string> keys;
  std::vector<float> ms;
  std::vector<int> cnt;
  for (auto& r : gn::g_prof) {
    cudaEventSynchronize(r.b);
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    size_t i = 0;
    for (; i < keys.size(); ++i) if (keys[i] == r.name) break;
    if (i == keys.size()) { keys.push_back(r.name); ms.push_back(0.f); cnt.push_back(0); }
    ms[i] += t; cnt[i] += 1;
  }
  gn::g_prof.clear();
  int n = static_cast<int>(keys.size()) < max_entries ? static_cast<int>(keys.size()) : max_entries;
  std::string joined;
  for (int i = 0; i < n; ++i) {
    if (total_ms) total_ms[i] = ms[i];
    if (counts) counts[i] = cnt[i];
    joined += keys[i];
    joined += ';';
  }
  if (names && names_len > 0) {
    std::strncpy(names, joined.c_str(), names_len - 1);
    names[names_len - 1] = 0;
  }
  return n;
}

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
  return gn::stage_workspace_bytes(cfg);
}

extern "C" int gn_stage_launch_count(const gn_stage_cfg* cfg) {
  if (!cfg) return 0;
  return gn::stage_launch_count(cfg);
}

extern "C" int gn_stage_fwd(const gn_stage_cfg* cfg, const gn_stage_weights* w,
                            const float* h_in, const float* H, const float* U,
                            float* node_out, float* dist_out,
                            void* workspace, size_t workspace_bytes, gn_stream_t stream) {
  if (!cfg || !w || !h_in || !node_out || !workspace) return GN_E_NULL;
  if ((reinterpret_cast<uintptr_t>(h_in) | reinterpret_cast<uintptr_t>(workspace)) & 15) return GN_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return gn::stage_fwd(cfg, w, h_in, H, U, node_out, dist_out, workspace, workspace_bytes, st);

}

extern "C" int gn_stage_saved_offsets(const gn_stage_cfg* cfg, size_t* out5) {
  if (!cfg || !out5) return GN_E_NULL;
  return gn::stage_saved_offsets(cfg, out5);
}

extern "C" size_t gn_stage_bwd_workspace_bytes(const gn_stage_cfg* cfg) {
  if (!cfg) return 0;
  return gn::stage_bwd_workspace_bytes(cfg);
}

extern "C" int gn_stage_bwd(const gn_stage_cfg* cfg, const gn_train_params* params,
                            const float* h_in, const float* H, const void* fwd_workspace,
                            const float* d_node_out, int64_t ld_dout, const float* d_dist, float* d_h,
                            void* workspace, size_t workspace_bytes, gn_stream_t stream) {
  if (!cfg || !params || !h_in || !fwd_workspace || !d_node_out || !d_h || !workspace) return GN_E_NULL;
  if (cfg->precision != GN_FP32) return GN_E_PRECISION;
  if (!cfg->pairwise && !H) return GN_E_NULL;
  size_t off[5];
  int rc = gn::stage_saved_offsets(cfg, off);
  if (rc != GN_OK) return rc;
  const char* fw = static_cast<const char*>(fwd_workspace);
  auto F = [&](size_t o) { return reinterpret_cast<const float*>(fw + o); };
  return gn::stage_bwd(cfg, params, h_in, H, F(off[0]), F(off[1]), F(off[2]), F(off[3]), F(off[4]),
                       d_node_out, ld_dout, d_dist, d_h, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

namespace gn { extern unsigned long long* g_trace_buffer; }
extern "C" void gn_profile_set_trace(unsigned long long* device_buffer) { gn::g_trace_buffer = device_buffer; }
