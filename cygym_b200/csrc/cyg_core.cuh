/*
 * cyg_core.cuh -- per-env transition of the CyGym step path on the INTERNAL bit-plane record.
 *
 * One thread owns one env.  The env's whole dynamic state is a record of S uint32 words
 * that the step kernel stages in shared memory (TMA bulk copy in, bulk copy out):
 *
 *   [0, 16)                          the 16 scalars of include/cygym_b200.h (CYG_S_*)
 *   [16 + p*W, 16 + (p+1)*W)         bit-plane p, W = ceil(M/32) words, bit d = device d
 *   [off_blocked, off_blocked + EW)  blocked-edge bitset over base CSR edge ids (out-list order)
 *   [off_blocked_in, ... + EW)       the same bits permuted into in-list order, so that the in-edges of a
 *                                    device are a contiguous bit range too (block / unblock, volt:485-511)
 * (the rarely used per-env extra attacker hub-star edges and the per-device checkpoint words
 * live in side arrays in global memory and are touched only by the actions that need them)
 *
 * Bit-planes turn every O(M) loop of the reference (busy tick volt_typhoon_env.py:904-908,
 * workload advance :1242-1261, _count_comp :563-572, candidate lists of
 * CDSimulator.py:244-348, evolve_network's active/inactive sets CyberDefenseEnv.py:654-659)
 * into W word-wide boolean operations, and the lateral-movement scan (:1148-1185) into
 * "first set bit of adj[src] & candidates".  Multi-bit fields (busy_time, processing_time,
 * compromised_by) are bit-sliced across planes.
 *
 * Every function is written once, as plain integer C++ with no warp intrinsics, so the very
 * same source also compiles for the host: tests/emu builds it with g++ and replays the golden
 * trajectories through it on the CPU-only build container (logic check of the device source;
 * it is not shipped and the product library has no CPU path).
 *
 * Statement order and quirks follow the reference; the line numbers cited are the reference's.
 */
#ifndef CYG_CORE_CUH
#define CYG_CORE_CUH

#include <math.h>
#include <stdint.h>

#include "../../include/cygym_b200.h"

#if defined(__CUDACC__)
#define CYG_HD __host__ __device__ __forceinline__
#define CYG_HDN __host__ __device__ __noinline__
#else
#define CYG_HD inline
#define CYG_HDN inline
#endif

namespace cyg {

/* draw sites: oracle/draws.py (each is one RNG call site of the reference) */
enum {
  SITE_STALL = 1, SITE_BLOCK = 2, SITE_UNBLOCK = 3, SITE_ZDAY = 4, SITE_PROBE = 5, SITE_WL_SAMPLE = 6,
  SITE_WL_TRI = 7, SITE_WL_LAZY = 8, SITE_EV_POISSON = 9, SITE_EV_ADD = 10, SITE_EV_PICK = 11,
  SITE_EV_ATT = 12, SITE_SHUFFLE = 13, SITE_DETECT = 14, SITE_SA_TYPE = 15, SITE_SA_NDEV = 16,
  SITE_SA_DEVS = 17, SITE_SA_EXP = 18, SITE_SA_APP = 19
};

/* bit-plane ids */
enum {
  P_COMP = 0, P_KNOWN = 1, P_NYA = 2, P_OWNED = 3, P_HASWL = 4, P_BUSYSET = 5, P_ACTSET = 6,
  P_PT0 = 7, P_BUSY0 = 10, P_CBY0 = 14
};
#define CYG_REC_PLANES 16 /* word offset of plane 0 inside a record */
#define CYG_CKI_REMOVED 0x40000000u /* internal checkpoint word, spare bit: Device.removed_before (never read by step) */
#define CYG_MAX_W 4       /* M <= 128 for the bit-matrix kernels */

/* ---- shared network tables + derived sizes (device pointers on the GPU) ---- */
struct Net {
  cyg_config cfg;
  int M, W, E, EW, NP, S, ncby, off_blocked, off_blocked_in;
  const uint32_t* adj;        /* [M][W] out-neighbour bit rows (unique pairs; _outnbrs, volt:456-473) */
  const uint32_t* adjT;       /* [M][W] in-neighbour bit rows (_innbrs) */
  const uint32_t* mlo;        /* [M][W] bit v of row u: (mult(u,v)-1) & 1 */
  const uint32_t* mhi;        /* [M][W] bit v of row u: (mult(u,v)-1) & 2 */
  const uint32_t* mloT;       /* transposes of mlo / mhi */
  const uint32_t* mhiT;
  const int32_t* row_ptr;     /* [M+1] */
  const uint16_t* col;        /* [E] */
  const int32_t* in_ptr;      /* [M+1] */
  const uint16_t* in_eid;     /* [E] base edge id of the j-th in-edge (ascending source) */
  const uint16_t* out2in;     /* [E] inverse of in_eid */
  const uint32_t* e_mlo;      /* [EW] bit e: (mult(e)-1) & 1, out-list order; e_mhi: & 2 */
  const uint32_t* e_mhi;
  const uint32_t* ei_mlo;     /* the same in in-list order */
  const uint32_t* ei_mhi;
  const uint32_t* dev_static; /* [M] CYG_ST_* */
  const uint32_t* m_dc;       /* [W] masks over devices */
  const uint32_t* m_server;
  const uint32_t* m_reach;
  const uint32_t* m_valid;    /* bits < M */
  const uint32_t* m_rowmulti; /* device has an out-pair with multiplicity > 1 */
  const uint32_t* m_incmulti; /* device has an incident (out or in) pair with multiplicity > 1 */
  const uint32_t* m_napps;    /* [8][W] bit-planes of len(device.apps) */
  const uint32_t* m_vuln;     /* [X][W] */
  const float* os_val;        /* [M] */
  const float* ver_val;       /* [M] */
  const uint32_t* blob;       /* base of the table blob; [blob, blob + hot_words) is what a CTA stages in smem */
  uint32_t hot_words;
};

CYG_HD int popc(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
CYG_HD int ctz(uint32_t x) { /* x != 0 */
#ifdef __CUDA_ARCH__
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
CYG_HD uint32_t below(uint32_t x, uint32_t n) { /* floor(x*n/2^32): oracle/draws.py:below */
#ifdef __CUDA_ARCH__
  return __umulhi(x, n);
#else
  return (uint32_t)(((uint64_t)x * n) >> 32);
#endif
}
CYG_HD uint32_t lowmask(int nbits) { return nbits >= 32 ? 0xFFFFFFFFu : ((1u << nbits) - 1u); }
CYG_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; __builtin_memcpy(&f, &u, 4); return f;
#endif
}
CYG_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; __builtin_memcpy(&u, &f, 4); return u;
#endif
}

/* position of the r-th (0-based) set bit of x; r < popc(x) */
CYG_HD int select_in_word(uint32_t x, int r) {
  int pos = 0;
#pragma unroll
  for (int sh = 16; sh >= 1; sh >>= 1) {
    int c = popc((x >> pos) & ((1u << sh) - 1u));
    if (r >= c) { r -= c; pos += sh; }
  }
  return pos;
}

/* ---- Philox4x32-10 and the addressable draw contract (oracle/draws.py) ---- */
CYG_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t o[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

struct Rng { uint32_t k0, k1, env, epoch; };

/* sequential reader of one site's draws k = 0, 1, 2, ... inside the current epoch */
struct Stream {
  uint32_t k, b0, b1, b2, b3;
  int site;
  CYG_HD explicit Stream(int s) : k(0), b0(0), b1(0), b2(0), b3(0), site(s) {}
  CYG_HD uint32_t next(const Rng& r) {
    if ((k & 3u) == 0u) {
      uint32_t o[4];
      philox4x32_10(r.env, r.epoch, (uint32_t)site, k >> 2, r.k0, r.k1, o);
      b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3];
    }
    uint32_t j = k & 3u;
    k++;
    return j == 0 ? b0 : j == 1 ? b1 : j == 2 ? b2 : b3;
  }
  /* continue at draw index k + n (n draws are skipped unread) */
  CYG_HD void skip(const Rng& r, uint32_t nskip) {
    if (nskip == 0) return;
    uint32_t old_blk = (k + 3u) >> 2; /* first block not loaded yet */
    k += nskip;
    if ((k & 3u) != 0u && (k >> 2) >= old_blk) {
      uint32_t o[4];
      philox4x32_10(r.env, r.epoch, (uint32_t)site, k >> 2, r.k0, r.k1, o);
      b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3];
    }
  }
};

/* ---- per-env view over an internal record --------------------------------- */
template <int W>
struct Env {
  const Net* n;
  uint32_t* rec;   /* the record (shared memory on the GPU) */
  uint32_t* ckpt;  /* canonical per-device checkpoint words of this env [M] (global memory) */
  uint32_t* xtra;  /* extra (attacker hub-star) edges of this env [xcap] (global memory) */
  Rng rng;
  Stream stall;    /* SITE_STALL is shared by every group of a grouped step */
  double defcost, cleancost;

  CYG_HD Env(const Net* net, uint32_t* record, uint32_t* ck, uint32_t* xt, uint32_t env_id)
      : n(net), rec(record), ckpt(ck), xtra(xt), stall(SITE_STALL), defcost(0.0), cleancost(0.0) {
    rng.k0 = (uint32_t)net->cfg.seed;
    rng.k1 = (uint32_t)(net->cfg.seed >> 32);
    rng.env = env_id;
    rng.epoch = 0;
  }
  CYG_HD uint32_t& scal(int i) { return rec[i]; }
  CYG_HD uint32_t& pl(int p, int w) { return rec[CYG_REC_PLANES + p * W + w]; }
  CYG_HD uint32_t* blocked() { return rec + n->off_blocked; }
  CYG_HD uint32_t* blocked_in() { return rec + n->off_blocked_in; }
  CYG_HD uint32_t* extra() { return xtra; }
  CYG_HD int n_extra() { return (int)(rec[CYG_S_PREV_X] >> 16); }

  /* open a draw epoch (one per step / randomize / sample_action call) */
  CYG_HD void begin_epoch() {
    rng.epoch = scal(CYG_S_EPOCH);
    scal(CYG_S_EPOCH) = rng.epoch + 1u;
    stall = Stream(SITE_STALL);
  }

  /* ---- single-device accessors ---- */
  CYG_HD bool bit(int p, int d) { return (pl(p, d >> 5) >> (d & 31)) & 1u; }
  CYG_HD void setb(int p, int d) { pl(p, d >> 5) |= 1u << (d & 31); }
  CYG_HD void clrb(int p, int d) { pl(p, d >> 5) &= ~(1u << (d & 31)); }
  CYG_HD uint32_t field(int p0, int nb, int d) {
    uint32_t v = 0;
    for (int k = 0; k < nb; k++) v |= ((pl(p0 + k, d >> 5) >> (d & 31)) & 1u) << k;
    return v;
  }
  CYG_HD void set_field(int p0, int nb, int d, uint32_t v) {
    uint32_t m = 1u << (d & 31);
    int w = d >> 5;
    for (int k = 0; k < nb; k++) {
      uint32_t x = pl(p0 + k, w) & ~m;
      if ((v >> k) & 1u) x |= m;
      pl(p0 + k, w) = x;
    }
  }
  CYG_HD uint32_t busy(int d) { return field(P_BUSY0, 4, d); }
  CYG_HD void set_busy(int d, uint32_t b) {
    if (b > CYG_BUSY_MAX) { b = CYG_BUSY_MAX; scal(CYG_S_FLAGS) |= CYG_FL_ERR_BUSY; }
    set_field(P_BUSY0, 4, d, b);
  }
  CYG_HD void drop_wl(int d) { clrb(P_HASWL, d); set_field(P_PT0, 3, d, 0); }
  CYG_HD uint32_t cby(int d) { return field(P_CBY0, n->ncby, d); }
  CYG_HD void clr_cby(int d) { set_field(P_CBY0, n->ncby, d, 0); }
  CYG_HD uint32_t stall_draw(int low, int high) { /* _stall: random.randint(low, high) (volt:135-138) */
    return (uint32_t)low + below(stall.next(rng), (uint32_t)(high - low + 1));
  }

  /* ---- word-wide helpers ---- */
  CYG_HD uint32_t busy_nz(int w) { return pl(P_BUSY0, w) | pl(P_BUSY0 + 1, w) | pl(P_BUSY0 + 2, w) | pl(P_BUSY0 + 3, w); }
  CYG_HD int count(int p) {
    int c = 0;
    for (int w = 0; w < W; w++) c += popc(pl(p, w));
    return c;
  }
  /* busy_time += 1 for every device in mask m[] (saturating at CYG_BUSY_MAX, flagged) */
  CYG_HD void busy_inc(const uint32_t* m) {
    for (int w = 0; w < W; w++) {
      uint32_t c = m[w];
      for (int k = 0; k < 4; k++) {
        uint32_t b = pl(P_BUSY0 + k, w);
        pl(P_BUSY0 + k, w) = b ^ c;
        c &= b;
      }
      if (c) { /* 15 + 1: saturate */
        for (int k = 0; k < 4; k++) pl(P_BUSY0 + k, w) |= c;
        scal(CYG_S_FLAGS) |= CYG_FL_ERR_BUSY;
      }
    }
  }
  /* busy_time -= 1 for every device in mask m[] that has busy_time > 0 */
  CYG_HD void busy_dec(const uint32_t* m) {
    for (int w = 0; w < W; w++) {
      uint32_t c = m[w] & busy_nz(w);
      for (int k = 0; k < 4; k++) {
        uint32_t b = pl(P_BUSY0 + k, w);
        pl(P_BUSY0 + k, w) = b ^ c;
        c &= ~b;
      }
    }
  }
  /* r-th (0-based, ascending id) member of mask m[]; r < total popcount */
  CYG_HD int select_nth(const uint32_t* m, int r) {
    for (int w = 0; w < W; w++) {
      int c = popc(m[w]);
      if (r < c) return w * 32 + select_in_word(m[w], r);
      r -= c;
    }
    return -1;
  }

  /* ---- topology: base bit rows + blocked bitset + extra edges ---------------- */
  CYG_HD bool any_blocked() {
    uint32_t o = 0;
    const uint32_t* b = blocked();
    for (int i = 0; i < n->EW; i++) o |= b[i];
    int nx = n_extra();
    const uint32_t* x = extra();
    for (int j = 0; j < nx; j++) o |= x[j] & CYG_X_BLOCKED;
    return o != 0;
  }
  /* out[] = out-neighbours of u whose edge has blocked-state == want_blocked (unique ids; base + extra) */
  CYG_HD void out_row(int u, bool has_blk, bool want_blocked, uint32_t* out) {
    uint32_t bl[W];
    for (int w = 0; w < W; w++) bl[w] = 0;
    if (has_blk) {
      const uint32_t* b = blocked();
      int a = n->row_ptr[u], z = n->row_ptr[u + 1];
      for (int wi = a >> 5; wi <= (z - 1) >> 5 && a < z; wi++) {
        uint32_t x = b[wi];
        if (wi == (a >> 5)) x &= ~lowmask(a & 31);
        if (wi == ((z - 1) >> 5)) x &= lowmask(((z - 1) & 31) + 1);
        while (x) {
          int e = wi * 32 + ctz(x);
          x &= x - 1;
          int v = n->col[e];
          for (int w = 0; w < W; w++) if (w == (v >> 5)) bl[w] |= 1u << (v & 31);
        }
      }
    }
    for (int w = 0; w < W; w++) {
      uint32_t r = n->adj[u * W + w];
      out[w] = want_blocked ? (r & bl[w]) : (r & ~bl[w]);
    }
    int nx = n_extra();
    const uint32_t* x = extra();
    for (int j = 0; j < nx; j++) {
      uint32_t xe = x[j];
      if ((int)(xe & CYG_X_IDMASK) != u) continue;
      if (((xe & CYG_X_BLOCKED) != 0) != want_blocked) continue;
      int v = (int)((xe >> CYG_X_V_SHIFT) & CYG_X_IDMASK);
      for (int w = 0; w < W; w++) if (w == (v >> 5)) out[w] |= 1u << (v & 31);
    }
  }
  /* in[] = in-neighbours (sources) of u whose edge has blocked-state == want_blocked */
  CYG_HD void in_row(int u, bool has_blk, bool want_blocked, uint32_t* in) {
    uint32_t bl[W];
    for (int w = 0; w < W; w++) bl[w] = 0;
    if (has_blk) {
      const uint32_t* b = blocked();
      int a = n->in_ptr[u], z = n->in_ptr[u + 1];
      int j = a;
      for (int w = 0; w < W; w++) { /* the j-th in-edge belongs to the j-th set bit of adjT[u] */
        uint32_t r = n->adjT[u * W + w];
        while (r) {
          int s = ctz(r);
          r &= r - 1;
          int e = n->in_eid[j++];
          if ((b[e >> 5] >> (e & 31)) & 1u) bl[w] |= 1u << s;
        }
      }
      (void)z;
    }
    for (int w = 0; w < W; w++) {
      uint32_t r = n->adjT[u * W + w];
      in[w] = want_blocked ? (r & bl[w]) : (r & ~bl[w]);
    }
    int nx = n_extra();
    const uint32_t* x = extra();
    for (int j = 0; j < nx; j++) {
      uint32_t xe = x[j];
      if ((int)((xe >> CYG_X_V_SHIFT) & CYG_X_IDMASK) != u) continue;
      if (((xe & CYG_X_BLOCKED) != 0) != want_blocked) continue;
      int s = (int)(xe & CYG_X_IDMASK);
      for (int w = 0; w < W; w++) if (w == (s >> 5)) in[w] |= 1u << (s & 31);
    }
  }
  CYG_HD int base_eid(int u, int v) { /* edge id of the base pair (u, v); the bit must be set in adj[u] */
    int e = n->row_ptr[u];
    for (int w = 0; w < W; w++) {
      uint32_t r = n->adj[u * W + w];
      if (w < (v >> 5)) e += popc(r);
      else if (w == (v >> 5)) e += popc(r & lowmask(v & 31));
    }
    return e;
  }
  CYG_HD bool has_edge(int u, int v) { /* g.get_eid(u, v) != -1 (CyberDefenseEnv.py:752-770) */
    if ((n->adj[u * W + (v >> 5)] >> (v & 31)) & 1u) return true;
    int nx = n_extra();
    const uint32_t* x = extra();
    uint32_t key = (uint32_t)u | ((uint32_t)v << CYG_X_V_SHIFT);
    for (int j = 0; j < nx; j++) if ((x[j] & 0xFFFFFFu) == key) return true;
    return false;
  }
  CYG_HD void set_base_blocked(int e, bool b) { /* both orders of the bitset */
    int j = n->out2in[e];
    if (b) { blocked()[e >> 5] |= 1u << (e & 31); blocked_in()[j >> 5] |= 1u << (j & 31); }
    else { blocked()[e >> 5] &= ~(1u << (e & 31)); blocked_in()[j >> 5] &= ~(1u << (j & 31)); }
  }
  /* flip the blocked flag of edge (u, v) */
  CYG_HD void set_edge_blocked(int u, int v, bool b) {
    if ((n->adj[u * W + (v >> 5)] >> (v & 31)) & 1u) {
      set_base_blocked(base_eid(u, v), b);
      return;
    }
    int nx = n_extra();
    uint32_t* x = extra();
    uint32_t key = (uint32_t)u | ((uint32_t)v << CYG_X_V_SHIFT);
    for (int j = 0; j < nx; j++)
      if ((x[j] & 0xFFFFFFu) == key) { if (b) x[j] |= CYG_X_BLOCKED; else x[j] &= ~CYG_X_BLOCKED; return; }
  }
  /* _rebuild_graph_cache (volt:456-483) forgets every block (:476) */
  CYG_HD void rebuild_cache() {
    uint32_t *b = blocked(), *bi = blocked_in();
    for (int i = 0; i < n->EW; i++) { b[i] = 0; bi[i] = 0; }
    int nx = n_extra();
    uint32_t* x = extra();
    for (int j = 0; j < nx; j++) x[j] &= ~CYG_X_BLOCKED;
  }
  /* multiplicity-weighted size of row r[] restricted to ids < lim (lim = 32*W: all) */
  CYG_HD int weight_below(const uint32_t* r, const uint32_t* lo, const uint32_t* hi, bool multi, int lim) {
    int c = 0;
    for (int w = 0; w < W; w++) {
      uint32_t m = r[w];
      int d = lim - 32 * w;
      if (d <= 0) m = 0; else if (d < 32) m &= lowmask(d);
      c += popc(m);
      if (multi) c += popc(m & lo[w]) + 2 * popc(m & hi[w]);
    }
    return c;
  }
  /* id holding the r-th unit of weight of row r[] */
  CYG_HD int weighted_select(const uint32_t* r, const uint32_t* lo, const uint32_t* hi, bool multi, int rank) {
    for (int w = 0; w < W; w++) {
      uint32_t m = r[w];
      if (!multi) {
        int c = popc(m);
        if (rank < c) return w * 32 + select_in_word(m, rank);
        rank -= c;
      } else {
        while (m) {
          int s = ctz(m);
          m &= m - 1;
          int wt = 1 + (int)((lo[w] >> s) & 1u) + 2 * (int)((hi[w] >> s) & 1u);
          if (rank < wt) return w * 32 + s;
          rank -= wt;
        }
      }
    }
    return -1;
  }

  /* ---- busy tick over _busy_devices (volt:904-908) ---- */
  CYG_HD void tick_busyset() {
    uint32_t m[W];
    for (int w = 0; w < W; w++) m[w] = pl(P_BUSYSET, w);
    busy_dec(m);
  }
  CYG_HD void tick_all() { /* _tick_busy_time_once (volt:607-610) */
    uint32_t m[W];
    for (int w = 0; w < W; w++) m[w] = 0xFFFFFFFFu;
    busy_dec(m);
  }

  /* ---- action decoding (include/cygym_b200.h "actions") ---- */
  struct Act {
    int mode, atype, n_ex, n_dev, app_index;
    uint32_t exw;
    const uint32_t* mask;
    const uint16_t* order;
    CYG_HD int ex(int i) const { return (int)(int8_t)((exw >> (8 * i)) & 0xFFu); }
  };
  CYG_HD static void decode(const uint32_t* hdr, const uint32_t* mask, const uint16_t* order, Act& a) {
    uint32_t h0 = hdr[0];
    int at = (int)(h0 & 0xFFu);
    a.atype = at == (int)CYG_ATYPE_NONE ? -1000 : (int)(int8_t)at;
    a.mode = (int)((h0 >> 8) & 1u);
    a.n_ex = (int)((h0 >> 16) & 0xFFu);
    if (a.n_ex > 4) a.n_ex = 4;
    a.exw = hdr[1];
    a.n_dev = (int)hdr[2];
    a.app_index = (int)hdr[3];
    a.mask = mask;
    a.order = order;
  }
  /* iterator over device_indices: explicit order array, or ascending bits of the mask */
  struct DevIter {
    int i, cur;
    CYG_HD DevIter() : i(0), cur(0) {}
  };
  CYG_HD int next_dev(const Act& a, DevIter& it) {
    if (a.order) return (int)a.order[it.i++];
    it.i++;
    int d = it.cur, M = n->M;
    while (d < M) {
      uint32_t x = a.mask[d >> 5] >> (d & 31);
      if (x) { d += ctz(x); break; }
      d = (d | 31) + 1;
    }
    it.cur = d + 1;
    return d < M ? d : -1;
  }
  CYG_HD int first_dev(const Act& a) {
    DevIter it;
    return a.n_dev > 0 ? next_dev(a, it) : -1;
  }

  /* ---- defender actions ---- */
  /* busy_time = randint(0, high) for every device of mask a[] in ascending id order (one _stall draw each) */
  CYG_HD void stall_deposit(const uint32_t* a, int low, int high) {
    uint32_t range = (uint32_t)(high - low + 1);
    for (int w = 0; w < W; w++) {
      uint32_t bits = a[w];
      if (!bits) continue;
      uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
      while (bits) {
        int s = ctz(bits);
        bits &= bits - 1;
        uint32_t v = (uint32_t)low + below(stall.next(rng), range);
        if (v > CYG_BUSY_MAX) { v = CYG_BUSY_MAX; scal(CYG_S_FLAGS) |= CYG_FL_ERR_BUSY; }
        r0 |= (v & 1u) << s; r1 |= ((v >> 1) & 1u) << s; r2 |= ((v >> 2) & 1u) << s; r3 |= ((v >> 3) & 1u) << s;
      }
      uint32_t keep = ~a[w];
      pl(P_BUSY0, w) = (pl(P_BUSY0, w) & keep) | r0;
      pl(P_BUSY0 + 1, w) = (pl(P_BUSY0 + 1, w) & keep) | r1;
      pl(P_BUSY0 + 2, w) = (pl(P_BUSY0 + 2, w) & keep) | r2;
      pl(P_BUSY0 + 3, w) = (pl(P_BUSY0 + 3, w) & keep) | r3;
    }
  }
  /* clear isCompromised, compromised_by and the workload of every device in a[] */
  CYG_HD void wipe(const uint32_t* a, bool comp_too) {
    for (int w = 0; w < W; w++) {
      uint32_t keep = ~a[w];
      if (comp_too) {
        pl(P_COMP, w) &= keep;
        for (int k = 0; k < n->ncby; k++) pl(P_CBY0 + k, w) &= keep;
      }
      pl(P_HASWL, w) &= keep;
      pl(P_PT0, w) &= keep; pl(P_PT0 + 1, w) &= keep; pl(P_PT0 + 2, w) &= keep;
    }
  }
  /* device_indices as a set: the first n_dev bits of the mask (what the reference loop visits) */
  CYG_HD void listed(const Act& a, uint32_t* l) {
    int nl = 0;
    for (int w = 0; w < W; w++) { l[w] = a.mask[w] & n->m_valid[w]; nl += popc(l[w]); }
    if (a.n_dev >= nl) return;
    int keep = a.n_dev < 0 ? 0 : a.n_dev, seen = 0; /* inconsistent header: fewer entries than mask bits */
    for (int w = 0; w < W; w++) {
      uint32_t lm = l[w], out = 0;
      while (lm) { uint32_t lb = lm & (0u - lm); lm ^= lb; if (seen++ < keep) out |= lb; }
      l[w] = out;
    }
  }

  /* clean every listed device: volt_typhoon_env.py:996-1011 (and :676-690 in the grouped path), set form */
  CYG_HD void clean_set(const Act& act, double ds, double& cost) {
    uint32_t a[W];
    int nc = 0, nu = 0;
    uint32_t disc = 0;
    listed(act, a);
    for (int w = 0; w < W; w++) {
      a[w] &= ~pl(P_NYA, w) & ~pl(P_OWNED, w);
      nc += popc(a[w] & pl(P_COMP, w));
      nu += popc(a[w] & ~pl(P_COMP, w));
      for (int k = 0; k < n->ncby; k++) if (pl(P_CBY0 + k, w) & a[w]) disc |= 1u << k;
    }
    cost += (nc * 0.3 - nu * 0.01) * ds;
    cleancost += (nc * 0.3 + nu * 0.01) * ds;
    defcost += (nc * 0.3 + nu * 0.01) * ds;
    scal(CYG_S_FLAGS) |= disc << CYG_FL_DISC_SHIFT; /* exp.discovered = True */
    wipe(a, true);
    stall_deposit(a, 0, n->cfg.default_high);
  }
  CYG_HD void clean_device(int d, double ds, double& cost) { /* same, one device (explicit order form) */
    if (bit(P_OWNED, d)) return;
    bool comp = bit(P_COMP, d);
    cost += (comp ? 0.3 : -0.01) * ds;
    cleancost += (comp ? 0.3 : 0.01) * ds;
    defcost += (comp ? 0.3 : 0.01) * ds;
    scal(CYG_S_FLAGS) |= cby(d) << CYG_FL_DISC_SHIFT;
    clr_cby(d);
    clrb(P_COMP, d);
    set_busy(d, stall_draw(0, n->cfg.default_high));
    drop_wl(d);
  }

  /* meta actions shared by step (volt:918-976) and _step_apply_only (volt:627-668) */
  CYG_HD void defender_meta(const Act& a, int atype, bool grouped, double& cost, bool& dirty) {
    const cyg_config& c = n->cfg;
    double ds = (double)c.def_scale;
    if (atype == 2) {
      scal(CYG_S_CKPT)++;
      scal(CYG_S_FLAGS) |= CYG_FL_HAS_CKPT; /* checkpoint_variables stores an alias, not a copy */
      cost += -0.5 * a.n_dev * ds;
      defcost += 0.5 * a.n_dev * ds;
      uint32_t m[W];
      for (int w = 0; w < W; w++) m[w] = busy_nz(w);
      busy_inc(m);
    } else if (atype == 3) {
      scal(CYG_S_REVERT)++;
      if (scal(CYG_S_FLAGS) & CYG_FL_HAS_CKPT) { /* every device: busy = randint(0, high), workload dropped */
        uint32_t all[W];
        for (int w = 0; w < W; w++) all[w] = n->m_valid[w];
        stall_deposit(all, 0, c.default_high);
        for (int w = 0; w < W; w++) { pl(P_HASWL, w) = 0; pl(P_PT0, w) = 0; pl(P_PT0 + 1, w) = 0; pl(P_PT0 + 2, w) = 0; }
        cost += -1.0 * a.n_dev * ds;
        dirty = true;
      }
    } else if (atype == 10) {
      if (!grouped) { /* volt:946-953; the grouped variant has no busy bump (volt:650-659) */
        if (a.n_dev > 0) {
          int d = first_dev(a);
          if (d >= 0 && d < n->M) set_busy(d, busy(d) + 1);
        } else {
          uint32_t m[W];
          for (int w = 0; w < W; w++) m[w] = busy_nz(w);
          busy_inc(m);
        }
      }
      cost += -1.0 * ds;
      if (scal(CYG_S_LOGS) > 0) scal(CYG_S_FLAGS) |= CYG_FL_DET_TRAINED; /* sklearn fit: host territory */
    } else if (atype == 11) {
      int d = first_dev(a); /* the host raises ValueError when n_dev == 0 (volt:965-966) */
      if (d >= 0 && d < n->M) {
        uint32_t k = CYG_CK_VALID;
        if (bit(P_COMP, d)) k |= CYG_CK_COMP;
        if (bit(P_KNOWN, d)) k |= CYG_CK_KNOWN;
        if (bit(P_NYA, d)) k |= CYG_CK_NYA;
        if (n->dev_static[d] & CYG_ST_REACH) k |= CYG_CK_REACH;
        if (bit(P_HASWL, d)) k |= CYG_CK_HASWL | (field(P_PT0, 3, d) << CYG_DEV_PT_SHIFT);
        k |= busy(d) << CYG_DEV_BUSY_SHIFT;
        k |= cby(d) << CYG_DEV_CBY_SHIFT;
        ckpt[d] = k | (ckpt[d] & CYG_CKI_REMOVED);
        scal(CYG_S_CKPT)++;
        cost += -0.1 * ds;
        defcost += 0.1 * ds;
      }
    }
  }

  CYG_HD void restore_device(int d) { /* _apply_device_state (volt:430-437) */
    uint32_t k = ckpt[d];
    if (k & CYG_CK_COMP) setb(P_COMP, d); else clrb(P_COMP, d);
    if (k & CYG_CK_KNOWN) setb(P_KNOWN, d); else clrb(P_KNOWN, d);
    if (k & CYG_CK_NYA) setb(P_NYA, d); else clrb(P_NYA, d);
    drop_wl(d);
    if (k & CYG_CK_HASWL) { setb(P_HASWL, d); set_field(P_PT0, 3, d, (k >> CYG_DEV_PT_SHIFT) & CYG_DEV_PT_MASK); }
    set_busy(d, (k >> CYG_DEV_BUSY_SHIFT) & CYG_DEV_BUSY_MASK);
    set_field(P_CBY0, n->ncby, d, (k >> CYG_DEV_CBY_SHIFT) & CYG_DEV_CBY_MASK);
  }

  /* block / unblock one incident edge of d (volt:1071-1080, :1091-1100, :485-511):
   * pool = out-edges then in-edges whose blocked flag == want, each repeated `multiplicity` times.
   * General form in neighbour-id space (extra edges, multi-edges). */
  CYG_HD bool flip_incident_general(int d, bool want, Stream& st) {
    bool has_blk = any_blocked();
    if (want && !has_blk) return false;
    uint32_t o[W], in[W];
    out_row(d, has_blk, want, o);
    in_row(d, has_blk, want, in);
    bool multi = ((n->m_incmulti[d >> 5] >> (d & 31)) & 1u) != 0;
    const uint32_t *lo = n->mlo + d * W, *hi = n->mhi + d * W, *loT = n->mloT + d * W, *hiT = n->mhiT + d * W;
    int to = weight_below(o, lo, hi, multi, 32 * W);
    int ti = weight_below(in, loT, hiT, multi, 32 * W);
    int total = to + ti;
    if (total == 0) return false;
    int r = (int)below(st.next(rng), (uint32_t)total);
    if (r < to) {
      int v = weighted_select(o, lo, hi, multi, r);
      set_edge_blocked(d, v, !want);
    } else {
      int s = weighted_select(in, loT, hiT, multi, r - to);
      set_edge_blocked(s, d, !want);
    }
    return true;
  }
  /* Same pick in EDGE-ID space when the env has no extra edges: the out-edges of d are the contiguous bits
   * [row_ptr[d], row_ptr[d+1]) of the blocked bitset (ascending neighbour id == pool order) and its in-edges the
   * bits [in_ptr[d], in_ptr[d+1]) of the in-order copy (ascending source id). */
  CYG_HD int range_weight(const uint32_t* b, uint32_t flipw, int a, int z, bool multi, const uint32_t* lo, const uint32_t* hi) {
    int c = 0;
    if (a >= z) return 0;
    for (int wi = a >> 5; wi <= (z - 1) >> 5; wi++) {
      uint32_t x = b[wi] ^ flipw;
      if (wi == (a >> 5)) x &= ~lowmask(a & 31);
      if (wi == ((z - 1) >> 5)) x &= lowmask(((z - 1) & 31) + 1);
      c += popc(x);
      if (multi) c += popc(x & lo[wi]) + 2 * popc(x & hi[wi]);
    }
    return c;
  }
  CYG_HD int range_select(const uint32_t* b, uint32_t flipw, int a, int z, bool multi, const uint32_t* lo, const uint32_t* hi, int r) {
    for (int wi = a >> 5; wi <= (z - 1) >> 5; wi++) {
      uint32_t x = b[wi] ^ flipw;
      if (wi == (a >> 5)) x &= ~lowmask(a & 31);
      if (wi == ((z - 1) >> 5)) x &= lowmask(((z - 1) & 31) + 1);
      int cnt = popc(x);
      if (multi) cnt += popc(x & lo[wi]) + 2 * popc(x & hi[wi]);
      if (r < cnt) {
        if (!multi || ((x & (lo[wi] | hi[wi])) == 0)) return wi * 32 + select_in_word(x, r);
        while (x) {
          int sb = ctz(x);
          x &= x - 1;
          int wt = 1 + (int)((lo[wi] >> sb) & 1u) + 2 * (int)((hi[wi] >> sb) & 1u);
          if (r < wt) return wi * 32 + sb;
          r -= wt;
        }
      }
      r -= cnt;
    }
    return -1;
  }
  CYG_HD bool flip_incident(int d, bool want, Stream& st) {
    if (n_extra() > 0) return flip_incident_general(d, want, st);
    const uint32_t flipw = want ? 0u : 0xFFFFFFFFu; /* pool bits = blocked bits XOR flipw */
    const bool multi = ((n->m_incmulti[d >> 5] >> (d & 31)) & 1u) != 0;
    const int a = n->row_ptr[d], z = n->row_ptr[d + 1], c0 = n->in_ptr[d], c1 = n->in_ptr[d + 1];
    int to = range_weight(blocked(), flipw, a, z, multi, n->e_mlo, n->e_mhi);
    int ti = range_weight(blocked_in(), flipw, c0, c1, multi, n->ei_mlo, n->ei_mhi);
    int total = to + ti;
    if (total == 0) return false;
    int r = (int)below(st.next(rng), (uint32_t)total);
    int e;
    if (r < to) e = range_select(blocked(), flipw, a, z, multi, n->e_mlo, n->e_mhi, r);
    else e = n->in_eid[range_select(blocked_in(), flipw, c0, c1, multi, n->ei_mlo, n->ei_mhi, r - to)];
    set_base_blocked(e, !want);
    return true;
  }

  /* per-device defender actions when device_indices is a SET (ascending, duplicate-free): the loop of
   * volt:989-1123 collapses to word-wide operations for every type but block / unblock */
  CYG_HD void defender_per_device_set(const Act& a, int atype, double& cost, bool& dirty) {
    const cyg_config& c = n->cfg;
    double ds = (double)c.def_scale;
    if (atype == 1) { clean_set(a, ds, cost); return; }
    uint32_t act[W];
    int na = 0;
    listed(a, act);
    for (int w = 0; w < W; w++) { act[w] &= ~pl(P_NYA, w); na += popc(act[w]); }
    if (na == 0) return;
    switch (atype) {
      case 4: { /* volt:1013-1018 */
        cost += -1.0 * ds * na;
        if (a.app_index >= 0 && a.app_index < 256) { /* devices with app_index < len(apps): bit-sliced compare */
          uint32_t up[W];
          for (int w = 0; w < W; w++) {
            uint32_t gt = 0, eq = 0xFFFFFFFFu;
            for (int bb = 7; bb >= 0; bb--) {
              uint32_t nb = n->m_napps[bb * W + w], ab = ((a.app_index >> bb) & 1) ? 0xFFFFFFFFu : 0u;
              gt |= eq & nb & ~ab;
              eq &= ~(nb ^ ab);
            }
            up[w] = act[w] & gt;
          }
          stall_deposit(up, 0, c.default_high);
        }
      } break;
      case 5: /* volt:1020-1069 with an untrained detector: every prediction is "D" */
        scal(CYG_S_SCAN) += (uint32_t)na;
        if (scal(CYG_S_LOGS) > 0) {
          if (scal(CYG_S_FLAGS) & CYG_FL_DET_TRAINED) scal(CYG_S_FLAGS) |= CYG_FL_ERR_DETECTOR;
          cost += -0.5 * ds * na;
          defcost += 0.5 * ds * na;
        }
        break;
      case 6: case 9: {
        Stream st(atype == 6 ? SITE_BLOCK : SITE_UNBLOCK);
        cost += -0.5 * ds * na;
        defcost += 0.5 * ds * na;
        uint32_t cnt = 0;
        for (int w = 0; w < W; w++) {
          uint32_t bits = act[w];
          while (bits) {
            int d = w * 32 + ctz(bits);
            bits &= bits - 1;
            if (flip_incident(d, atype == 9, st)) cnt++;
          }
        }
        if (cnt) { scal(atype == 6 ? CYG_S_EBLK : CYG_S_EADD) += cnt; dirty = true; }
      } break;
      case 7: /* volt:1082-1089 */
        cost += -0.5 * ds * na;
        for (int w = 0; w < W; w++) pl(P_NYA, w) |= act[w];
        wipe(act, true);
        dirty = true;
        break;
      case 12: { /* acts on device_indices[0] once per listed active device (volt:1102-1109) */
        int dev0 = first_dev(a);
        if (ckpt[dev0] & CYG_CK_VALID) {
          restore_device(dev0);
          cost += -1.0 * ds * na;
          defcost += 1.0 * ds * na;
        }
      } break;
      case 13: { /* volt:1111-1123: only the last of the `na` _stall draws survives */
        int dev0 = first_dev(a);
        clrb(P_COMP, dev0);
        clr_cby(dev0);
        drop_wl(dev0);
        stall.skip(rng, (uint32_t)(na - 1));
        set_busy(dev0, stall_draw(3, c.default_high + 3));
        cost += -3.0 * ds * na;
        cleancost += 3.0 * ds * na;
        defcost += 3.0 * ds * na;
      } break;
      default: break;
    }
  }

  /* per-device defender actions in listed order (volt:989-1123): explicit order form (may repeat devices) */
  CYG_HD void defender_per_device_seq(const Act& a, int atype, double& cost, bool& dirty) {
    const cyg_config& c = n->cfg;
    double ds = (double)c.def_scale;
    int dev0 = first_dev(a);
    DevIter it;
    Stream sblk(SITE_BLOCK), sunb(SITE_UNBLOCK);
    for (int i = 0; i < a.n_dev; i++) {
      int d = next_dev(a, it);
      if (d < 0 || d >= n->M) break;
      if (bit(P_NYA, d)) continue;
      switch (atype) {
        case 1: clean_device(d, ds, cost); break;
        case 4:
          cost += -1.0 * ds;
          if (a.app_index >= 0 && a.app_index < (int)((n->dev_static[d] >> CYG_ST_NAPPS_SHIFT) & 0xFFu))
            set_busy(d, stall_draw(0, c.default_high));
          break;
        case 5:
          scal(CYG_S_SCAN)++;
          if (scal(CYG_S_LOGS) > 0) {
            if (scal(CYG_S_FLAGS) & CYG_FL_DET_TRAINED) scal(CYG_S_FLAGS) |= CYG_FL_ERR_DETECTOR;
            cost += -0.5 * ds;
            defcost += 0.5 * ds;
          }
          break;
        case 6:
          cost += -0.5 * ds;
          defcost += 0.5 * ds;
          if (flip_incident(d, false, sblk)) { scal(CYG_S_EBLK)++; dirty = true; }
          break;
        case 7:
          cost += -0.5 * ds;
          setb(P_NYA, d);
          clrb(P_COMP, d);
          clr_cby(d);
          drop_wl(d);
          dirty = true;
          break;
        case 9:
          cost += -0.5 * ds;
          defcost += 0.5 * ds;
          if (flip_incident(d, true, sunb)) { scal(CYG_S_EADD)++; dirty = true; }
          break;
        case 12:
          if (ckpt[dev0] & CYG_CK_VALID) {
            restore_device(dev0);
            cost += -1.0 * ds;
            defcost += 1.0 * ds;
          }
          break;
        case 13:
          clrb(P_COMP, dev0);
          clr_cby(dev0);
          drop_wl(dev0);
          set_busy(dev0, stall_draw(3, c.default_high + 3));
          cost += -3.0 * ds;
          cleancost += 3.0 * ds;
          defcost += 3.0 * ds;
          break;
        default: break;
      }
    }
  }
  CYG_HD void defender_per_device(const Act& a, int atype, double& cost, bool& dirty) {
    if (a.order) defender_per_device_seq(a, atype, cost, dirty);
    else defender_per_device_set(a, atype, cost, dirty);
  }

  /* ---- attacker actions (volt:1126-1202) ---- */
  /* unblocked out-neighbours of s (base row minus blocked pairs, plus unblocked extra edges) */
  CYG_HD void live_row(int s, bool has_blk, int nx, uint32_t* row) {
    for (int w = 0; w < W; w++) row[w] = n->adj[s * W + w];
    if (has_blk) {
      const uint32_t* b = blocked();
      int a = n->row_ptr[s], z = n->row_ptr[s + 1];
      if (a < z) {
        for (int wi = a >> 5; wi <= (z - 1) >> 5; wi++) {
          uint32_t x = b[wi];
          if (wi == (a >> 5)) x &= ~lowmask(a & 31);
          if (wi == ((z - 1) >> 5)) x &= lowmask(((z - 1) & 31) + 1);
          while (x) {
            int e = wi * 32 + ctz(x);
            x &= x - 1;
            int v = n->col[e];
            for (int w = 0; w < W; w++) if (w == (v >> 5)) row[w] &= ~(1u << (v & 31));
          }
        }
      }
    }
    if (nx > 0) {
      const uint32_t* x = extra();
      for (int j = 0; j < nx; j++) {
        uint32_t xe = x[j];
        if ((int)(xe & CYG_X_IDMASK) != s || (xe & CYG_X_BLOCKED)) continue;
        int v = (int)((xe >> CYG_X_V_SHIFT) & CYG_X_IDMASK);
        for (int w = 0; w < W; w++) if (w == (v >> 5)) row[w] |= 1u << (v & 31);
      }
    }
  }
  CYG_HD void attacker_act(const Act& a, int atype, double& cost) {
    const cyg_config& c = n->cfg;
    if (c.base_line == CYG_BL_NO_ATTACK) return;
    if (atype != 1 && atype != 2) return;
    uint32_t src[W]; /* snapshot of compromised-or-owned devices, taken before the loop (volt:1127-1128) */
    int ns = 0;
    for (int w = 0; w < W; w++) { src[w] = pl(P_COMP, w) | pl(P_OWNED, w); ns += popc(src[w]); }
    const bool has_blk = any_blocked();
    const int nx = n_extra();
    if (atype == 1) {
      Stream zday(SITE_ZDAY);
      uint32_t logs = scal(CYG_S_LOGS);
      uint32_t comp[W], known[W];
      for (int w = 0; w < W; w++) { comp[w] = pl(P_COMP, w); known[w] = pl(P_KNOWN, w); }
      for (int xi = 0; xi < a.n_ex; xi++) {
        int raw = a.ex(xi);
        if (c.zero_day && !(raw >= 0 && raw < 32 && ((c.zero_day_mask >> raw) & 1u))) { /* volt:1135-1136 */
          int cnt = popc(c.zero_day_mask);
          raw = select_in_word(c.zero_day_mask, (int)below(zday.next(rng), (uint32_t)cnt));
        }
        if (!(raw >= 0 && raw < c.n_exploits)) continue; /* ids are strings: an int never matches (volt:1141) */
        uint32_t kv[W], dcby[W]; /* known & vulnerable to this exploit; compromised_by additions */
        for (int w = 0; w < W; w++) { kv[w] = known[w] & n->m_vuln[raw * W + w]; dcby[w] = 0; }
        for (int sw = 0; sw < W; sw++) {
          uint32_t sbits = src[sw];
          const uint32_t dcw = n->m_dc[sw], mlw = n->m_rowmulti[sw];
          while (sbits) {
            const int sb = ctz(sbits);
            const int s = sw * 32 + sb;
            sbits &= sbits - 1;
            uint32_t row[W];
            live_row(s, has_blk, nx, row);
            const bool is_dc = ((dcw >> sb) & 1u) != 0;
            /* first neighbour that is hit: DC source -> any; reachable_by_attacker; or not yet compromised,
               known and vulnerable (volt:1163-1183) */
            int vw = -1;
            uint32_t cand_w = 0;
            for (int w = W - 1; w >= 0; w--) {
              uint32_t cand = is_dc ? row[w] : (row[w] & (n->m_reach[w] | (~comp[w] & kv[w])));
              if (cand) { vw = w; cand_w = cand; }
            }
            /* log_communication once per hop walked (volt:1161): all repeats of the neighbours before the hit, +1 */
            uint32_t hitbit = cand_w & (0u - cand_w);
            int cnt = 0;
            const bool multi = ((mlw >> sb) & 1u) != 0;
            for (int w = 0; w < W; w++) {
              uint32_t m = vw < 0 ? row[w] : (w < vw ? row[w] : (w == vw ? (row[w] & (hitbit - 1u)) : 0u));
              cnt += popc(m);
              if (multi) cnt += popc(m & n->mlo[s * W + w]) + 2 * popc(m & n->mhi[s * W + w]);
            }
            logs += (uint32_t)cnt + (vw >= 0 ? 1u : 0u);
            if (vw >= 0) {
              for (int w = 0; w < W; w++) if (w == vw) { comp[w] |= hitbit; if (is_dc) dcby[w] |= hitbit; }
            }
          }
        }
        for (int w = 0; w < W; w++) if (dcby[w]) pl(P_CBY0 + raw, w) |= dcby[w];
      }
      for (int w = 0; w < W; w++) pl(P_COMP, w) = comp[w];
      scal(CYG_S_LOGS) = logs;
    } else { /* probe (volt:1187-1202) */
      if (ns > 0) {
        Stream sp(SITE_PROBE);
        int s = select_nth(src, (int)below(sp.next(rng), (uint32_t)ns));
        uint32_t row[W];
        live_row(s, has_blk, nx, row);
        for (int w = 0; w < W; w++) {
          uint32_t cand = row[w] & ~pl(P_KNOWN, w);
          if (cand) { pl(P_KNOWN, w) |= cand & (0u - cand); cost += 0.1; break; }
        }
      }
    }
  }

  /* ---- workload advance (volt:1242-1261); returns current_work ---- */
  CYG_HD int workload_advance() {
    int cur = 0;
    for (int w = 0; w < W; w++) {
      uint32_t p0 = pl(P_PT0, w), p1 = pl(P_PT0 + 1, w), p2 = pl(P_PT0 + 2, w);
      uint32_t m = ~busy_nz(w) & ~pl(P_NYA, w) & pl(P_HASWL, w) & (p0 | p1 | p2);
      uint32_t b = m; /* processing_time -= 1 on m */
      uint32_t q0 = p0 ^ b; b &= ~p0;
      uint32_t q1 = p1 ^ b; b &= ~p1;
      uint32_t q2 = p2 ^ b;
      uint32_t fin = m & ~(q0 | q1 | q2);
      pl(P_PT0, w) = q0; pl(P_PT0 + 1, w) = q1; pl(P_PT0 + 2, w) = q2;
      pl(P_HASWL, w) &= ~fin;
      cur += popc(fin);
    }
    scal(CYG_S_WORK) += (uint32_t)cur;
    return cur;
  }

  /* ---- arrivals (volt:575-596, :141-145, :184-191, :266-293; CDSimulator.py:244-348) ---- */
  CYG_HD void generate_workloads(int num_loads, bool server, int n_active, Stream& ssamp, Stream& stri) {
    const cyg_config& c = n->cfg;
    if (n_active <= 0) return;                                  /* volt:205-207 */
    if (c.wl_cap >= 0 && num_loads > c.wl_cap) num_loads = c.wl_cap; /* volt:210-211 */
    if (num_loads > n_active) num_loads = n_active;             /* volt:234 */
    if (num_loads <= 0) return;
    uint32_t cand[W];
    int nc = 0;
    for (int w = 0; w < W; w++) {
      uint32_t t = n->m_valid[w] & ~pl(P_NYA, w) & ~pl(P_HASWL, w) & ~busy_nz(w);
      t &= server ? n->m_server[w] : ~n->m_server[w];
      cand[w] = t;
      nc += popc(t);
    }
    int k = num_loads < nc ? num_loads : nc;
    for (int j = 0; j < k; j++) { /* random.sample: pop the r-th remaining candidate (CDSimulator.py:298) */
      int r = (int)below(ssamp.next(rng), (uint32_t)(nc - j));
      int did = select_nth(cand, r);
      for (int w = 0; w < W; w++) if (w == (did >> 5)) cand[w] &= ~(1u << (did & 31));
      uint32_t xt = stri.next(rng); /* CDSimulator.py:308 */
      int pt = 1;
      for (int v = 0; v < 8; v++) pt += xt >= c.tri_tab[v];
      if (pt > c.tri_high) pt = c.tri_high;
      setb(P_HASWL, did);
      set_field(P_PT0, 3, did, (uint32_t)pt);
    }
  }
  CYG_HD void arrivals_if_due() {
    const cyg_config& c = n->cfg;
    int n_active = 0, idle = 0, free_s = 0;
    for (int w = 0; w < W; w++) {
      uint32_t act = n->m_valid[w] & ~pl(P_NYA, w);
      uint32_t idl = act & ~busy_nz(w) & ~pl(P_HASWL, w);
      n_active += popc(act);
      idle += popc(idl);
      free_s += popc(idl & n->m_server[w]);
    }
    int free_c = idle - free_s;
    /* _arrival_period (volt:141-145) */
    int period = (int)(c.wl_period_base + 0.5 * sqrt((double)(n_active > 1 ? n_active : 1)));
    if (period < 10) period = 10;
    if (period > c.wl_period_max) period = c.wl_period_max;
    if (scal(CYG_S_STEP) % (uint32_t)period != 0) return;
    if (n_active == 0 || 10 * idle < n_active) return; /* _idle_fraction() < 0.10 (volt:580) */
    int nC = 100, nS = 10;
    if (c.scaling_vulnerability) { /* _scaled_numloads (volt:266-293) */
      int req_c = 2 * n_active;
      int req_s = (2 * n_active + 5) / 10;
      if (req_c < 1) req_c = 1;
      if (req_s < 1) req_s = 1;
      int cap_c = free_c > 1 ? free_c : 1, cap_s = free_s > 1 ? free_s : 1;
      nC = req_c < cap_c ? req_c : cap_c;
      nS = req_s < cap_s ? req_s : cap_s;
    }
    if (c.wl_cap > 0) { /* volt:588-593 */
      int total = nC + nS;
      if (total > c.wl_cap) {
        double ratio = (double)c.wl_cap / (double)total;
        nC = (int)(nC * ratio); if (nC < 0) nC = 0;
        nS = (int)(nS * ratio); if (nS < 0) nS = 0;
      }
    }
    Stream ssamp(SITE_WL_SAMPLE), stri(SITE_WL_TRI);
    generate_workloads(nC, false, n_active, ssamp, stri);
    generate_workloads(nS, true, n_active, ssamp, stri);
  }

  /* ---- evolve_network (CyberDefenseEnv.py:583-875) ---- */
  CYG_HD void evolve_network() {
    const cyg_config& c = n->cfg;
    if (!(scal(CYG_S_FLAGS) & CYG_FL_SETS_INIT)) { /* :654-659 */
      for (int w = 0; w < W; w++) pl(P_ACTSET, w) = n->m_valid[w] & ~pl(P_NYA, w);
      scal(CYG_S_FLAGS) |= CYG_FL_SETS_INIT;
    }
    int n_act = count(P_ACTSET);
    Stream sp(SITE_EV_POISSON), sadd(SITE_EV_ADD), spick(SITE_EV_PICK), satt(SITE_EV_ATT);
    uint32_t xp = sp.next(rng); /* :668 */
    int num_events = 0;
    for (int j = 0; j < 16; j++) num_events += xp >= c.poisson_tab[j];
    int floor_n = c.num_of_device > c.min_network_size ? c.num_of_device : c.min_network_size;
    for (int ev = 0; ev < num_events; ev++) {
      uint32_t xa = sadd.next(rng); /* :679 */
      if ((uint64_t)xa < c.thr_p_add) {
        int n_inact = n->M - n_act;
        if (n_inact > 0) {
          uint32_t m[W];
          for (int w = 0; w < W; w++) m[w] = n->m_valid[w] & ~pl(P_ACTSET, w);
          int node = select_nth(m, (int)below(spick.next(rng), (uint32_t)n_inact)); /* :675 */
          clrb(P_NYA, node);
          setb(P_ACTSET, node);
          n_act++;
          uint32_t xt = satt.next(rng); /* :690: the draw is consumed even when p_attacker == 0 */
          if ((uint64_t)xt < c.thr_p_attacker) { setb(P_COMP, node); setb(P_OWNED, node); setb(P_KNOWN, node); }
        }
      } else if (n_act > floor_n) { /* :701-712 */
        uint32_t m[W];
        for (int w = 0; w < W; w++) m[w] = pl(P_ACTSET, w);
        int node = select_nth(m, (int)below(spick.next(rng), (uint32_t)n_act));
        setb(P_NYA, node);
        ckpt[node] |= CYG_CKI_REMOVED; /* removed_before */
        drop_wl(node);
        set_field(P_BUSY0, 4, node, 0);
        clrb(P_ACTSET, node);
        n_act--;
      }
    }
    /* bidirectional hub-star among active attacker-owned devices, hub = lowest id (:738-774) */
    int hub = -1;
    bool changed = false;
    for (int w = 0; w < W; w++) {
      uint32_t oa = pl(P_OWNED, w) & pl(P_ACTSET, w);
      while (oa) {
        int i = w * 32 + ctz(oa);
        oa &= oa - 1;
        if (hub < 0) { hub = i; continue; }
        for (int dir = 0; dir < 2; dir++) {
          int u = dir ? i : hub, v = dir ? hub : i;
          if (has_edge(u, v)) continue;
          int nx = n_extra();
          if (nx >= c.xcap) { scal(CYG_S_FLAGS) |= CYG_FL_ERR_XCAP; continue; }
          extra()[nx] = (uint32_t)u | ((uint32_t)v << CYG_X_V_SHIFT);
          scal(CYG_S_PREV_X) = (scal(CYG_S_PREV_X) & 0xFFFFu) | ((uint32_t)(nx + 1) << 16);
          changed = true;
        }
      }
    }
    /* the preferential-attachment repair (:776-843) needs a degree-0 vertex: dead on these graphs */
    if (changed) rebuild_cache();
  }

  CYG_HD void count_comp(int& n_comp, int& n_comp_dc) { /* _count_comp (volt:563-572) */
    int a = 0, b = 0;
    for (int w = 0; w < W; w++) {
      uint32_t m = pl(P_COMP, w) & ~pl(P_NYA, w) & ~pl(P_OWNED, w);
      a += popc(m);
      b += popc(m & n->m_dc[w]);
    }
    n_comp = a; n_comp_dc = b;
  }

  /* ---- the step (volt_typhoon_env.py:818-1333; grouped: :694-779) ---- */
  /* the action type step() ends up executing: None fill (volt:847-874), clamp into the action space
   * (:879-884), defender forced to the no-op unless base_line == "Nash" (:913-914).  Needs no env state,
   * so the kernel can sort a block's envs by it before their records arrive. */
  CYG_HD static int exec_type(const cyg_config& c, uint32_t h0) {
    int at = (int)(h0 & 0xFFu);
    int mode = (int)((h0 >> 8) & 1u);
    int atype = at == (int)CYG_ATYPE_NONE ? -1000 : (int)(int8_t)at;
    if (atype == -1000) {
      if (mode == CYG_MODE_DEFENDER) atype = (c.base_line == CYG_BL_NO_DEFENSE) ? 8 : 7;
      else atype = (c.base_line == CYG_BL_NO_ATTACK) ? 3 : 2;
    }
    if (mode == CYG_MODE_DEFENDER) { if (!(atype >= 0 && atype < c.def_space_n)) atype = 8; }
    else { if (!(atype >= 0 && atype < c.att_space_n)) atype = 3; }
    if (mode == CYG_MODE_DEFENDER && c.base_line != CYG_BL_NASH) atype = 8;
    return atype;
  }

  CYG_HD int step(const uint32_t* hdr, const uint32_t* mask, const uint16_t* order, size_t hdr_gs, size_t mask_gs,
                  size_t order_gs, int G, uint32_t flags, float* raw_out, float* shaped_out, int32_t* done_out,
                  uint32_t* pre_masks) {
    const cyg_config& c = n->cfg;
    const bool grouped = (flags & CYG_STEP_GROUPED) != 0;
    const bool skip_work = (flags & CYG_STEP_SKIP_WORK) != 0;
    begin_epoch();
    double cost = 0.0;
    bool dirty = false;
    defcost = (double)u2f(scal(CYG_S_DEFCOST));
    cleancost = (double)u2f(scal(CYG_S_CLEANCOST));
    Act a;
    decode(hdr, mask, order, a);
    const int mode = a.mode;
    int atype = 0;
    if (!grouped) {
      atype = exec_type(c, hdr[0]);
      if (a.atype == -1000) { a.n_dev = 0; a.n_ex = 1; a.exw = 0; a.app_index = 0; }
      tick_busyset(); /* volt:904-908 */
      if (mode == CYG_MODE_DEFENDER) {
        defender_meta(a, atype, false, cost, dirty);
        if (atype == 1 || atype == 4 || atype == 5 || atype == 6 || atype == 7 || atype == 9 || atype == 12 || atype == 13)
          defender_per_device(a, atype, cost, dirty);
      } else {
        attacker_act(a, atype, cost);
      }
    } else {
      for (int g = 0; g < G; g++) { /* _step_apply_only (volt:612-692) */
        Act ga;
        decode(hdr + g * hdr_gs, mask + g * mask_gs, order ? order + g * order_gs : (const uint16_t*)0, ga);
        int gt = ga.atype == -1000 ? 0 : ga.atype;
        if (gt == 0) gt = (mode == CYG_MODE_DEFENDER) ? 8 : 3;
        if (mode == CYG_MODE_DEFENDER) {
          if (c.base_line != CYG_BL_NASH) gt = 8;
          defender_meta(ga, gt, true, cost, dirty);
          if (gt == 1) {
            double ds = (double)c.def_scale;
            if (!ga.order) {
              clean_set(ga, ds, cost);
            } else {
              DevIter it;
              for (int i = 0; i < ga.n_dev; i++) {
                int d = next_dev(ga, it);
                if (d < 0 || d >= n->M) break;
                if (bit(P_NYA, d)) continue;
                clean_device(d, ds, cost);
              }
            }
          }
        }
        atype = gt;
      }
      tick_all(); /* _tick_busy_time_once (volt:607-610) */
    }
    scal(CYG_S_DEFCOST) = f2u((float)defcost);
    scal(CYG_S_CLEANCOST) = f2u((float)cleancost);

    /* work, arrivals, reward, counters, evolve (volt:1207-1333) */
    int cur_work = 0;
    if (!skip_work || grouped) {
      cur_work = workload_advance();
      arrivals_if_due();
    }
    double def_work = (double)c.work_scale * cur_work;
    int n_comp, n_comp_dc;
    count_comp(n_comp, n_comp_dc);
    if (!grouped) scal(CYG_S_COMPCNT) += (uint32_t)n_comp; /* volt:1267-1270; absent from step_grouped */
    double raw, shaped;
    if (mode == CYG_MODE_DEFENDER) {
      raw = cost + def_work - n_comp * (double)c.comp_scale;
      shaped = raw;
    } else {
      raw = cost + (double)c.comp_scale * (n_comp + 10 * n_comp_dc);
      double Md = (double)n->M;
      double phi = (double)n_comp / Md;
      double gam = (double)c.gamma;
      uint32_t pn = scal(CYG_S_PREV_X) & 0xFFFFu;
      double prev = (pn == 0xFFFFu) ? phi : gam * ((double)pn / Md);
      double bonus = 0.1 * (gam * phi - prev);
      scal(CYG_S_PREV_X) = (scal(CYG_S_PREV_X) & 0xFFFF0000u) | (uint32_t)n_comp;
      shaped = raw + bonus;
    }
    if (pre_masks) { /* the `state` step() returns is the pre-evolve view (volt:1306) */
      for (int w = 0; w < W; w++) {
        pre_masks[0 * W + w] = pl(P_COMP, w);
        pre_masks[1 * W + w] = pl(P_KNOWN, w);
        pre_masks[2 * W + w] = pl(P_NYA, w);
      }
    }
    if (!skip_work || grouped) {
      scal(CYG_S_STEP)++;
      if (mode == CYG_MODE_ATTACKER) scal(CYG_S_ATT_STEP)++; else scal(CYG_S_DEF_STEP)++;
    }
    int done = scal(CYG_S_STEP) > 1000u; /* _check_done (CyberDefenseEnv.py:547-552) */
    bool periodic = (scal(CYG_S_STEP) % (uint32_t)c.evolve_period) == 0;
    if (dirty || periodic) evolve_network();
    if (!grouped) { /* volt:1330 */
      for (int w = 0; w < W; w++) pl(P_BUSYSET, w) = busy_nz(w);
    }
    *raw_out = (float)raw; *shaped_out = (float)shaped; *done_out = done;
    return atype;
  }

  /* ---- randomize_compromise_and_ownership (volt:330-383) ---- */
  CYG_HD void randomize() {
    begin_epoch();
    uint32_t pool[W];
    int np = 0, k_owned = 0, k_comp = 0;
    for (int w = 0; w < W; w++) {
      uint32_t p = n->m_valid[w] & ~pl(P_NYA, w) & ~n->m_dc[w];
      pool[w] = p;
      np += popc(p);
      k_owned += popc(p & pl(P_OWNED, w));
      k_comp += popc(p & pl(P_COMP, w));
    }
    if (np == 0 || (k_owned == 0 && k_comp == 0)) return;
    int extra_c = k_comp - k_owned;
    if (extra_c < 0) extra_c = 0;
    for (int w = 0; w < W; w++) { pl(P_OWNED, w) &= ~pool[w]; pl(P_COMP, w) &= ~pool[w]; pl(P_KNOWN, w) &= ~pool[w]; }
    Stream ss(SITE_SHUFFLE);
    int rem = np;
    for (int j = 0; j < k_owned + extra_c && j < np; j++) { /* shuffle == successive uniform picks */
      int r = rem > 1 ? (int)below(ss.next(rng), (uint32_t)rem) : 0;
      int d = select_nth(pool, r);
      for (int w = 0; w < W; w++) if (w == (d >> 5)) pool[w] &= ~(1u << (d & 31));
      rem--;
      if (j < k_owned) setb(P_OWNED, d);
      setb(P_COMP, d);
      setb(P_KNOWN, d);
    }
  }

  /* ---- sample_action (CyberDefenseEnv.py:555-578): device_indices as a set ---- */
  CYG_HD void sample_action(int mode, uint32_t* hdr, uint32_t* mask) {
    const cyg_config& c = n->cfg;
    begin_epoch();
    Stream st(SITE_SA_TYPE), sn(SITE_SA_NDEV), sd(SITE_SA_DEVS), sx(SITE_SA_EXP), sa(SITE_SA_APP);
    int space = mode == CYG_MODE_DEFENDER ? c.def_space_n : c.att_space_n;
    int atype = (int)below(st.next(rng), (uint32_t)space);
    int ndev = 1 + (int)below(sn.next(rng), (uint32_t)c.num_of_device);
    uint32_t pool[W], pick[W];
    for (int w = 0; w < W; w++) { pool[w] = n->m_valid[w]; pick[w] = 0; }
    int rem = n->M;
    for (int j = 0; j < ndev; j++) {
      int d = select_nth(pool, (int)below(sd.next(rng), (uint32_t)rem));
      for (int w = 0; w < W; w++) if (w == (d >> 5)) { pool[w] &= ~(1u << (d & 31)); pick[w] |= 1u << (d & 31); }
      rem--;
    }
    int ex = (int)below(sx.next(rng), (uint32_t)c.X);
    int app = c.n_app_ids > 0 ? (int)below(sa.next(rng), (uint32_t)c.n_app_ids) : 0;
    hdr[0] = (uint32_t)(atype & 0xFF) | ((uint32_t)mode << 8) | (1u << 16);
    hdr[1] = (uint32_t)(ex & 0xFF);
    hdr[2] = (uint32_t)ndev;
    hdr[3] = (uint32_t)app;
    for (int w = 0; w < W; w++) mask[w] = pick[w];
  }
};

/* ---- canonical device word <-> bit-planes (include/cygym_b200.h) ------------ */
template <int W>
CYG_HD void import_device(const Net* n, uint32_t* rec, int d, uint32_t w) {
  uint32_t m = 1u << (d & 31);
  int wi = d >> 5;
  uint32_t* pl = rec + CYG_REC_PLANES;
  auto put = [&](int p, bool v) { if (v) pl[p * W + wi] |= m; else pl[p * W + wi] &= ~m; };
  put(P_COMP, w & CYG_DEV_COMP); put(P_KNOWN, w & CYG_DEV_KNOWN); put(P_NYA, w & CYG_DEV_NYA);
  put(P_OWNED, w & CYG_DEV_OWNED); put(P_HASWL, w & CYG_DEV_HASWL);
  put(P_BUSYSET, w & CYG_DEV_BUSYSET); put(P_ACTSET, w & CYG_DEV_ACTSET);
  uint32_t pt = (w >> CYG_DEV_PT_SHIFT) & CYG_DEV_PT_MASK;
  for (int k = 0; k < 3; k++) put(P_PT0 + k, (pt >> k) & 1u);
  uint32_t b = (w >> CYG_DEV_BUSY_SHIFT) & CYG_DEV_BUSY_MASK;
  if (b > CYG_BUSY_MAX) { b = CYG_BUSY_MAX; rec[CYG_S_FLAGS] |= CYG_FL_ERR_BUSY; }
  for (int k = 0; k < 4; k++) put(P_BUSY0 + k, (b >> k) & 1u);
  uint32_t cb = (w >> CYG_DEV_CBY_SHIFT) & CYG_DEV_CBY_MASK;
  for (int k = 0; k < n->ncby; k++) put(P_CBY0 + k, (cb >> k) & 1u);
}
template <int W>
CYG_HD uint32_t export_device(const Net* n, const uint32_t* rec, int d, uint32_t ckpt_internal) {
  int wi = d >> 5, s = d & 31;
  const uint32_t* pl = rec + CYG_REC_PLANES;
  auto get = [&](int p) -> uint32_t { return (pl[p * W + wi] >> s) & 1u; };
  uint32_t w = 0;
  if (get(P_COMP)) w |= CYG_DEV_COMP;
  if (get(P_KNOWN)) w |= CYG_DEV_KNOWN;
  if (get(P_NYA)) w |= CYG_DEV_NYA;
  if (get(P_OWNED)) w |= CYG_DEV_OWNED;
  if (ckpt_internal & CYG_CKI_REMOVED) w |= CYG_DEV_REMOVED;
  if (get(P_HASWL)) w |= CYG_DEV_HASWL;
  if (get(P_BUSYSET)) w |= CYG_DEV_BUSYSET;
  if (get(P_ACTSET)) w |= CYG_DEV_ACTSET;
  for (int k = 0; k < 3; k++) w |= get(P_PT0 + k) << (CYG_DEV_PT_SHIFT + k);
  for (int k = 0; k < 4; k++) w |= get(P_BUSY0 + k) << (CYG_DEV_BUSY_SHIFT + k);
  for (int k = 0; k < n->ncby; k++) w |= get(P_CBY0 + k) << (CYG_DEV_CBY_SHIFT + k);
  return w;
}

/* one element of an observation row (CyberDefenseEnv.py:146-257); obs_mode as in cyg_step_out */
template <int W>
CYG_HD float observe_elem(const Net* n, const uint32_t* rec, int obs_mode, int j) {
  int M = n->M;
  const uint32_t* pl = rec + CYG_REC_PLANES;
  if (obs_mode == 2) {
    if (j >= 4 * M) return (j - 4 * M) < n->cfg.n_exploits ? 1.f : 0.f;
    int d = j >> 2, k = j & 3, wi = d >> 5, s = d & 31;
    bool known = (pl[P_KNOWN * W + wi] >> s) & 1u, nya = (pl[P_NYA * W + wi] >> s) & 1u, owned = (pl[P_OWNED * W + wi] >> s) & 1u;
    if (!known || nya || !owned) return -1.f;
    if (k == 0) return n->os_val[d];
    if (k == 1) return n->ver_val[d];
    if (k == 2) return ((pl[P_COMP * W + wi] >> s) & 1u) ? 1.f : 0.f;
    return 1.f; /* known */
  }
  int d = j / 6, k = j - 6 * d, wi = d >> 5, s = d & 31;
  bool known = (pl[P_KNOWN * W + wi] >> s) & 1u, nya = (pl[P_NYA * W + wi] >> s) & 1u, owned = (pl[P_OWNED * W + wi] >> s) & 1u;
  bool comp = (pl[P_COMP * W + wi] >> s) & 1u;
  if (obs_mode == 1 && (nya || !owned)) return -1.f;
  switch (k) {
    case 0: return n->os_val[d];
    case 1: return n->ver_val[d];
    case 2: return obs_mode == 1 ? -1.f : (comp ? 1.f : 0.f);
    case 3: return 0.f; /* anomaly_score stays 0 under fast_scan (volt:46) */
    case 4: return known ? 1.f : 0.f;
    default: return nya ? 1.f : 0.f;
  }
}

}  // namespace cyg
#endif /* CYG_CORE_CUH */
