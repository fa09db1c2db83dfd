/*
 * cyg_emu.cpp -- host build of the DEVICE source cygym_b200/csrc/cyg_core.cuh (TEST INFRASTRUCTURE).
 *
 * The per-env transition the CUDA kernels run is written as portable integer C++.  This file
 * compiles exactly that source with g++ and drives it one env at a time over canonical arrays,
 * so the CPU-only build container can replay the golden trajectories through the device logic
 * (tests/test_emu_golden.py).  It is never loaded by the product package: cygym_b200 has no
 * CPU path and fails loudly when its CUDA library is missing.
 */
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <math.h>
#include "../../cygym_b200/csrc/cyg_core.cuh"
#include "../../cygym_b200/csrc/cyg_tables.h"

using namespace cyg;

struct Emu {
  TableBlob blob;
  Net net;
  uint32_t* logs = nullptr;            /* [B][log_cap] hop-log rings */
  const uint32_t* det_slots = nullptr; /* uploaded detectors */
  const int32_t* det_of_env = nullptr;
};

static std::string g_err;

template <int W>
static void import_env(const Net& n, uint32_t* rec, const uint32_t* dev, const uint32_t* blocked, const uint32_t* extra,
                       const uint32_t* scal, uint32_t* ckpt) {
  memset(rec, 0, sizeof(uint32_t) * (size_t)n.S);
  for (int i = 0; i < CYG_NSCAL; i++) rec[i] = scal[i];
  for (int d = 0; d < n.M; d++) {
    import_device<W>(&n, rec, d, dev[d]);
    if (ckpt) ckpt[d] = (ckpt[d] & ~CYG_CKI_REMOVED) | ((dev[d] & CYG_DEV_REMOVED) ? CYG_CKI_REMOVED : 0u);
  }
  for (int e = 0; e < n.E; e++) {
    if (!((blocked[e >> 5] >> (e & 31)) & 1u)) continue;
    rec[n.off_aux]++;
    pair_units(&n, e, [&](int wi, uint32_t m) { rec[n.off_inc + wi] |= m; });
  }
}
template <int W>
static void export_env(const Net& n, const uint32_t* rec, uint32_t* dev, uint32_t* blocked, uint32_t* extra, uint32_t* scal,
                       uint32_t* ckpt) {
  for (int i = 0; i < CYG_NSCAL; i++) scal[i] = rec[i];
  for (int d = 0; d < n.M; d++) {
    dev[d] = export_device<W>(&n, rec, d, ckpt ? ckpt[d] : ((dev[d] & CYG_DEV_REMOVED) ? CYG_CKI_REMOVED : 0u));
    if (ckpt) ckpt[d] &= ~CYG_CKI_REMOVED;
  }
  for (int i = 0; i < n.EW; i++) blocked[i] = 0;
  for (int e = 0; e < n.E; e++) if (pair_blocked(&n, rec, e)) blocked[e >> 5] |= 1u << (e & 31);
}

template <int W>
static void step_w(Emu* em, int B, int env_id0, uint32_t* dev, uint32_t* ckpt, uint32_t* blocked, uint32_t* extra,
                   uint32_t* scal, const uint32_t* hdr, const uint32_t* mask, const uint16_t* order, int order_stride, int G,
                   uint32_t flags, float* raw, float* shaped, int32_t* done, int32_t* exec_atype, uint32_t* pre_masks) {
  const Net& n = em->net;
  std::vector<uint32_t> rec(n.S);
  for (int b = 0; b < B; b++) {
    uint32_t* dv = dev + (size_t)b * n.M;
    import_env<W>(n, rec.data(), dv, blocked + (size_t)b * n.EW, extra + (size_t)b * n.cfg.xcap, scal + (size_t)b * CYG_NSCAL,
                  ckpt + (size_t)b * n.M);
    const uint32_t* h = hdr + (size_t)b * 4;
    const uint32_t* m = mask + (size_t)b * n.Wm;
    const uint16_t* o = order ? order + (size_t)b * order_stride : nullptr;
    Env<W> e(&n, rec.data(), ckpt + (size_t)b * n.M, extra + (size_t)b * n.cfg.xcap, (uint32_t)(env_id0 + b));
    if (em->logs && n.cfg.log_cap > 0) e.logs = em->logs + (size_t)b * n.cfg.log_cap;
    if (em->det_slots && em->det_of_env && em->det_of_env[b] >= 0) e.det = em->det_slots + (size_t)em->det_of_env[b] * CYG_DET_WORDS;
    int atype = e.step(h, m, o, (size_t)B * 4, (size_t)B * n.Wm, (size_t)B * order_stride, G, flags, raw + b, shaped + b,
                       done + b, pre_masks ? pre_masks + (size_t)b * 3 * n.Wm : nullptr);
    if (exec_atype) exec_atype[b] = atype;
    export_env<W>(n, rec.data(), dv, blocked + (size_t)b * n.EW, extra + (size_t)b * n.cfg.xcap, scal + (size_t)b * CYG_NSCAL,
                  ckpt + (size_t)b * n.M);
  }
}

template <int W>
static void randomize_w(Emu* em, int B, int env_id0, uint32_t* dev, uint32_t* blocked, uint32_t* extra, uint32_t* scal,
                        const uint8_t* env_mask) {
  const Net& n = em->net;
  std::vector<uint32_t> rec(n.S);
  for (int b = 0; b < B; b++) {
    if (env_mask && !env_mask[b]) continue;
    uint32_t* dv = dev + (size_t)b * n.M;
    import_env<W>(n, rec.data(), dv, blocked + (size_t)b * n.EW, extra + (size_t)b * n.cfg.xcap, scal + (size_t)b * CYG_NSCAL, nullptr);
    Env<W> e(&n, rec.data(), nullptr, extra + (size_t)b * n.cfg.xcap, (uint32_t)(env_id0 + b));
    e.randomize();
    export_env<W>(n, rec.data(), dv, blocked + (size_t)b * n.EW, extra + (size_t)b * n.cfg.xcap, scal + (size_t)b * CYG_NSCAL, nullptr);
  }
}

template <int W>
static void rebuild_w(Emu* em, int B, int env_id0, uint32_t* dev, uint32_t* blocked, uint32_t* extra, uint32_t* scal) {
  const Net& n = em->net;
  std::vector<uint32_t> rec(n.S);
  for (int b = 0; b < B; b++) {
    uint32_t* dv = dev + (size_t)b * n.M;
    import_env<W>(n, rec.data(), dv, blocked + (size_t)b * n.EW, extra + (size_t)b * n.cfg.xcap, scal + (size_t)b * CYG_NSCAL, nullptr);
    Env<W> e(&n, rec.data(), nullptr, extra + (size_t)b * n.cfg.xcap, (uint32_t)(env_id0 + b));
    e.rebuild_cache();
    export_env<W>(n, rec.data(), dv, blocked + (size_t)b * n.EW, extra + (size_t)b * n.cfg.xcap, scal + (size_t)b * CYG_NSCAL, nullptr);
  }
}

template <int W>
static void sample_w(Emu* em, int B, int env_id0, uint32_t* scal, int mode, uint32_t* hdr, uint32_t* mask, uint16_t* order, int order_stride) {
  const Net& n = em->net;
  std::vector<uint32_t> rec(n.S);
  for (int b = 0; b < B; b++) {
    for (int i = 0; i < CYG_NSCAL; i++) rec[i] = scal[(size_t)b * CYG_NSCAL + i];
    Env<W> e(&n, rec.data(), nullptr, nullptr, (uint32_t)(env_id0 + b));
    e.sample_action(mode, hdr + (size_t)b * 4, mask + (size_t)b * n.Wm, order ? order + (size_t)b * order_stride : nullptr);
    for (int i = 0; i < CYG_NSCAL; i++) scal[(size_t)b * CYG_NSCAL + i] = rec[i];
  }
}

template <int W>
static void observe_w(Emu* em, int B, const uint32_t* dev, int obs_mode, float* obs) {
  const Net& n = em->net;
  std::vector<uint32_t> rec(n.S);
  int dim = obs_mode == 2 ? 4 * n.M + n.cfg.X : 6 * n.M;
  for (int b = 0; b < B; b++) {
    memset(rec.data(), 0, sizeof(uint32_t) * (size_t)n.S);
    for (int d = 0; d < n.M; d++) import_device<W>(&n, rec.data(), d, dev[(size_t)b * n.M + d]);
    for (int j = 0; j < dim; j++) obs[(size_t)b * dim + j] = observe_elem<W>(&n, rec.data(), obs_mode, j);
  }
}

#define DISPATCH_W(fn, ...)                         \
  switch (em->net.W) {                              \
    case 1: fn<1>(__VA_ARGS__); break;              \
    case 2: fn<2>(__VA_ARGS__); break;              \
    case 3: fn<3>(__VA_ARGS__); break;              \
    case 4: fn<4>(__VA_ARGS__); break;              \
    default: fn<CYG_BIG_W>(__VA_ARGS__); break;     \
  }

extern "C" {
void* emu_create(const cyg_config* cfg, const int32_t* row_ptr, const int32_t* col, const uint8_t* mult,
                 const uint32_t* dev_static, const float* os_val, const float* ver_val) {
  Emu* em = new Emu();
  cyg_network hn = {row_ptr, col, mult, dev_static, os_val, ver_val};
  g_err = build_tables(*cfg, hn, em->blob);
  if (!g_err.empty()) { delete em; return nullptr; }
  relocate(em->blob, em->blob.words.data(), em->net);
  return em;
}
const char* emu_last_error(void) { return g_err.c_str(); }
void emu_destroy(void* h) { delete (Emu*)h; }
void emu_set_base_line(void* h, int32_t bl) { ((Emu*)h)->net.cfg.base_line = bl; }
void emu_set_aux(void* h, uint32_t* logs, const uint32_t* det_slots, const int32_t* det_of_env) {
  Emu* em = (Emu*)h;
  em->logs = logs; em->det_slots = det_slots; em->det_of_env = det_of_env;
}
int emu_pyset_order(const int* vals, int n, int* out) { return Env<1>::pyset_order(vals, n, out); }
int emu_record_words(void* h) { return ((Emu*)h)->net.S; }
int emu_step(void* h, int B, int env_id0, uint32_t* dev, uint32_t* ckpt, uint32_t* blocked, uint32_t* extra, uint32_t* scal,
             const uint32_t* hdr, const uint32_t* mask, const uint16_t* order, int order_stride, int G, uint32_t flags,
             float* raw, float* shaped, int32_t* done, int32_t* exec_atype, uint32_t* pre_masks) {
  Emu* em = (Emu*)h;
  DISPATCH_W(step_w, em, B, env_id0, dev, ckpt, blocked, extra, scal, hdr, mask, order, order_stride, G, flags, raw, shaped,
             done, exec_atype, pre_masks);
  return 0;
}
int emu_randomize(void* h, int B, int env_id0, uint32_t* dev, uint32_t* blocked, uint32_t* extra, uint32_t* scal,
                  const uint8_t* env_mask) {
  Emu* em = (Emu*)h;
  DISPATCH_W(randomize_w, em, B, env_id0, dev, blocked, extra, scal, env_mask);
  return 0;
}
int emu_rebuild(void* h, int B, int env_id0, uint32_t* dev, uint32_t* blocked, uint32_t* extra, uint32_t* scal) {
  Emu* em = (Emu*)h;
  DISPATCH_W(rebuild_w, em, B, env_id0, dev, blocked, extra, scal);
  return 0;
}
int emu_sample_actions(void* h, int B, int env_id0, uint32_t* scal, int mode, uint32_t* hdr, uint32_t* mask, uint16_t* order, int order_stride) {
  Emu* em = (Emu*)h;
  DISPATCH_W(sample_w, em, B, env_id0, scal, mode, hdr, mask, order, order_stride);
  return 0;
}
int emu_observe(void* h, int B, const uint32_t* dev, int obs_mode, float* obs) {
  Emu* em = (Emu*)h;
  DISPATCH_W(observe_w, em, B, dev, obs_mode, obs);
  return 0;
}
}
