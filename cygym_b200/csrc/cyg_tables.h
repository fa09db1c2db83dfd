/*
 * cyg_tables.h -- host-side construction of the shared network tables (plain C++).
 *
 * Turns the caller's CSR description of the base graph (cyg_network: the flattened
 * _outnbrs cache of volt_typhoon_env.py:456-473 plus per-device statics) into the bit
 * matrices and masks the kernels read (cyg::Net in cyg_core.cuh), all packed in ONE blob so
 * that the library uploads it with a single copy and a CTA stages its hot part with a single
 * bulk copy.
 */
#ifndef CYG_TABLES_H
#define CYG_TABLES_H

#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "cyg_core.cuh"

namespace cyg {

struct TableBlob {
  std::vector<uint32_t> words; /* everything, 16-byte aligned sections */
  Net net;                     /* pointers are OFFSETS (in words) until relocate() */
  size_t hot_words;            /* prefix that the step kernel stages in shared memory */
  size_t o_adj, o_dc, o_server, o_reach, o_valid, o_napps, o_vuln, o_dinfo, o_unit, o_omulti, o_static, o_pair2unit, o_os, o_ver;
};

inline size_t tb_alloc(TableBlob& b, size_t nwords) {
  size_t off = b.words.size();
  size_t padded = (nwords + 3) & ~(size_t)3;
  b.words.resize(off + padded, 0u);
  return off;
}

/* Returns "" on success, else an error message. */
inline std::string build_tables(const cyg_config& cfg, const cyg_network& hn, TableBlob& b) {
  const int M = cfg.M, E = cfg.E, X = cfg.X;
  if (M < 1 || M > 32 * CYG_BIG_W) return "M must be in 1..2048";
  if (E < 0 || E > 65535) return "E out of range";
  if (X < 1 || X > 6) return "X (MaxExploits) must be in 1..6";
  if (cfg.n_exploits < 0 || cfg.n_exploits > X) return "n_exploits must be in 0..X";
  if (cfg.xcap < 0 || cfg.xcap > 4095) return "xcap out of range";
  if (cfg.evolve_period < 1) return "evolve_period must be >= 1";
  if (cfg.log_cap < 0 || cfg.log_cap > 65536) return "log_cap must be in 0..65536";
  if (cfg.wl_period_max < 1) return "wl_period_max must be >= 1";
  const int W = M <= 32 * CYG_MAX_W ? (M + 31) / 32 : CYG_BIG_W; /* large networks: planes padded to 64 words */
  const int EW = E > 0 ? (E + 31) / 32 : 1;
  if (hn.row_ptr[0] != 0 || hn.row_ptr[M] != E) return "row_ptr must start at 0 and end at E";
  /* first pass: validate, degrees in UNITS (a pair of multiplicity m is m units on either side) */
  std::vector<int> outu(M, 0), inu(M, 0), indeg(M, 0);
  for (int u = 0; u < M; u++) {
    if (hn.row_ptr[u + 1] < hn.row_ptr[u]) return "row_ptr not monotone";
    int prev = -1;
    for (int e = hn.row_ptr[u]; e < hn.row_ptr[u + 1]; e++) {
      int v = hn.col[e];
      if (v < 0 || v >= M) return "col out of range";
      if (v <= prev) return "neighbour lists must be strictly ascending (unique pairs; use mult[] for multi-edges)";
      prev = v;
      int mu = hn.mult ? hn.mult[e] : 1;
      if (mu < 1 || mu > 4) return "edge multiplicity must be in 1..4";
      outu[u] += mu; inu[v] += mu; indeg[v]++;
    }
  }
  /* evolve_network's preferential-attachment repair (CyberDefenseEnv.py:776-843) fires for a newly activated vertex of
   * total degree 0 and draws random.uniform; the kernels leave that branch out, so such a network is refused loudly */
  for (int i = 0; i < M; i++)
    if (indeg[i] == 0 && hn.row_ptr[i + 1] == hn.row_ptr[i] && M > 1)
      return "device " + std::to_string(i) + " has no incident edge: the preferential-attachment repair of evolve_network "
             "(CyberDefenseEnv.py:776-843) is not part of the kernels";
  std::vector<int> ip(M + 1, 0);
  for (int i = 0; i < M; i++) ip[i + 1] = ip[i] + outu[i] + inu[i];
  const int U2 = ip[M];
  if (U2 > 65535) return "more than 65535 incidence units (2 x sum of edge multiplicities)";
  for (int i = 0; i < M; i++) if (outu[i] > 65535) return "out-degree out of range";
  const int UW = U2 > 0 ? (U2 + 31) / 32 : 1;
  Net& n = b.net;
  memset(&n, 0, sizeof(n));
  n.cfg = cfg;
  n.M = M; n.W = W; n.E = E; n.EW = EW; n.Wm = (M + 31) / 32; n.U2 = U2; n.UW = UW;
  n.ncby = cfg.n_exploits > 0 ? cfg.n_exploits : 1;
  n.NP = P_CBY0 + n.ncby;
  n.off_inc = CYG_REC_PLANES + n.NP * W;
  n.off_aux = n.off_inc + UW;
  int S = n.off_aux + 1;
  if ((S & 1) == 0) S++; /* odd stride: thread-per-env accesses to shared memory are bank-conflict free */
  n.S = S;
  b.words.clear();
  /* hot section (staged in shared memory by the step kernels).  The adjacency bit rows come last: for the large
   * networks (W = 64: 8 KB per device row set, 512 KB at M = 2000) they stay in global memory and the hot prefix ends
   * in front of them -- the large-network kernel walks the unit table (CSR) instead. */
  b.o_dc = tb_alloc(b, W);
  b.o_server = tb_alloc(b, W);
  b.o_reach = tb_alloc(b, W);
  b.o_valid = tb_alloc(b, W);
  b.o_napps = tb_alloc(b, (size_t)8 * W);
  b.o_vuln = tb_alloc(b, (size_t)X * W);
  b.o_dinfo = tb_alloc(b, M + 1);
  b.o_unit = tb_alloc(b, (size_t)U2 + 1);
  b.o_omulti = tb_alloc(b, M);
  b.o_static = tb_alloc(b, M);
  if (W > CYG_MAX_W) b.hot_words = b.words.size();
  b.o_adj = tb_alloc(b, (size_t)M * W);
  if (W <= CYG_MAX_W) b.hot_words = b.words.size();
  /* cold section: read through L1/L2 (canonical pair <-> unit map for import / export, observation rows) */
  b.o_pair2unit = tb_alloc(b, (size_t)E + 1);
  b.o_os = tb_alloc(b, M);
  b.o_ver = tb_alloc(b, M);
  uint32_t* w = b.words.data();
  for (int i = 0; i <= M; i++) w[b.o_dinfo + i] = (uint32_t)ip[i] | ((uint32_t)(i < M ? outu[i] : 0) << 16);
  /* incidence units: device d owns [ip[d], ip[d+1]): its out-units (ascending neighbour id, a pair of multiplicity m
   * as m adjacent units -- the order of _outnbrs, volt:456-473), then its in-units (ascending source id).  Entry of a
   * unit: twin run start | far endpoint << 16 | (m - 1) << 28 | offset inside its own run << 30. */
  std::vector<int> ofill(M, 0), ifill(M, 0);
  for (int u = 0; u < M; u++) { /* ascending u => in-lists ascending by source */
    int pairs = 0, multi_cnt = 0;
    uint32_t dm = 0xFFu | (0xFFu << 10);
    for (int e = hn.row_ptr[u]; e < hn.row_ptr[u + 1]; e++, pairs++) {
      const int v = hn.col[e];
      const int mu = hn.mult ? hn.mult[e] : 1;
      const int so = ip[u] + ofill[u], si = ip[v] + outu[v] + ifill[v];
      ofill[u] += mu; ifill[v] += mu;
      for (int k = 0; k < mu; k++) {
        w[b.o_unit + so + k] = (uint32_t)si | ((uint32_t)v << 16) | ((uint32_t)(mu - 1) << 28) | ((uint32_t)k << 30);
        w[b.o_unit + si + k] = (uint32_t)so | ((uint32_t)u << 16) | ((uint32_t)(mu - 1) << 28) | ((uint32_t)k << 30);
      }
      w[b.o_pair2unit + e] = (uint32_t)so;
      w[b.o_adj + (size_t)u * W + (v >> 5)] |= 1u << (v & 31);
      if (mu > 1) { /* packed multi-edge runs of the out list: pair rank | (m - 1) << 8, two entries; bit 31: more */
        if (multi_cnt >= 2 || pairs >= 0xFF) dm |= 0x80000000u;
        else {
          dm &= ~(0x3FFu << (10 * multi_cnt));
          dm |= ((uint32_t)pairs | ((uint32_t)(mu - 1) << 8)) << (10 * multi_cnt);
          multi_cnt++;
        }
      }
    }
    w[b.o_omulti + u] = dm;
  }
  for (int i = 0; i < M; i++) {
    uint32_t st = hn.dev_static[i];
    w[b.o_static + i] = st;
    uint32_t bit = 1u << (i & 31);
    int wi = i >> 5;
    w[b.o_valid + wi] |= bit;
    if (st & CYG_ST_DC) w[b.o_dc + wi] |= bit;
    if (st & CYG_ST_SERVER) w[b.o_server + wi] |= bit;
    if (st & CYG_ST_REACH) w[b.o_reach + wi] |= bit;
    for (int k = 0; k < 8; k++)
      if ((st >> (CYG_ST_NAPPS_SHIFT + k)) & 1u) w[b.o_napps + (size_t)k * W + wi] |= bit;
    for (int e = 0; e < X; e++)
      if ((st >> (CYG_ST_VULN_SHIFT + e)) & 1u) w[b.o_vuln + (size_t)e * W + wi] |= bit;
    float osv = hn.os_val ? hn.os_val[i] : (float)i, vv = hn.ver_val ? hn.ver_val[i] : 0.f;
    memcpy(&w[b.o_os + i], &osv, 4);
    memcpy(&w[b.o_ver + i], &vv, 4);
  }
  return "";
}

/* point the Net at a copy of the blob living at `base` (host or device address) */
inline void relocate(const TableBlob& b, const uint32_t* base, Net& n) {
  n = b.net;
  n.blob = base;
  n.hot_words = (uint32_t)b.hot_words;
  n.inv_M = 1.0 / (double)n.M;
  n.o_adj = (uint32_t)b.o_adj;
  n.o_dc = (uint32_t)b.o_dc; n.o_server = (uint32_t)b.o_server; n.o_reach = (uint32_t)b.o_reach; n.o_valid = (uint32_t)b.o_valid;
  n.o_napps = (uint32_t)b.o_napps; n.o_vuln = (uint32_t)b.o_vuln;
  n.o_dinfo = (uint32_t)b.o_dinfo; n.o_unit = (uint32_t)b.o_unit; n.o_omulti = (uint32_t)b.o_omulti; n.o_static = (uint32_t)b.o_static;
  n.o_pair2unit = (uint32_t)b.o_pair2unit;
  n.o_os = (uint32_t)b.o_os; n.o_ver = (uint32_t)b.o_ver;
}

}  // namespace cyg
#endif
