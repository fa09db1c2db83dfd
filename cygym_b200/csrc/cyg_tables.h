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
  size_t o_adj, o_adjT, o_mlo, o_mhi, o_mloT, o_mhiT, o_row_ptr, o_col, o_in_ptr, o_in_eid, o_static, o_dc, o_server,
      o_reach, o_valid, o_rowmulti, o_incmulti, o_napps, o_vuln, o_os, o_ver, o_out2in, o_in_src, o_emlo, o_emhi, o_eimlo, o_eimhi, o_dmulti;
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
  if (cfg.wl_period_max < 1) return "wl_period_max must be >= 1";
  const int W = M <= 32 * CYG_MAX_W ? (M + 31) / 32 : CYG_BIG_W; /* large networks: planes padded to 64 words */
  const int EW = E > 0 ? (E + 31) / 32 : 1;
  if (hn.row_ptr[0] != 0 || hn.row_ptr[M] != E) return "row_ptr must start at 0 and end at E";
  Net& n = b.net;
  memset(&n, 0, sizeof(n));
  n.cfg = cfg;
  n.M = M; n.W = W; n.E = E; n.EW = EW; n.Wm = (M + 31) / 32;
  n.ncby = cfg.n_exploits > 0 ? cfg.n_exploits : 1;
  n.NP = P_CBY0 + n.ncby;
  n.off_blocked = CYG_REC_PLANES + n.NP * W;
  n.off_blocked_in = n.off_blocked + EW;
  n.off_aux = n.off_blocked_in + EW;
  int S = n.off_aux + 1;
  if ((S & 1) == 0) S++; /* odd stride: thread-per-env accesses to shared memory are bank-conflict free */
  n.S = S;
  b.words.clear();
  /* hot section (staged in shared memory by the step kernel) */
  b.o_adj = tb_alloc(b, (size_t)M * W);
  b.o_dc = tb_alloc(b, W);
  b.o_server = tb_alloc(b, W);
  b.o_reach = tb_alloc(b, W);
  b.o_valid = tb_alloc(b, W);
  b.o_rowmulti = tb_alloc(b, W);
  b.o_incmulti = tb_alloc(b, W);
  b.o_napps = tb_alloc(b, (size_t)8 * W);
  b.o_vuln = tb_alloc(b, (size_t)X * W);
  b.o_row_ptr = tb_alloc(b, M + 1);
  b.o_col = tb_alloc(b, (E + 1) / 2 + 1);
  b.o_in_ptr = tb_alloc(b, M + 1);
  b.o_in_eid = tb_alloc(b, (E + 1) / 2 + 1);
  b.o_static = tb_alloc(b, M);
  b.o_out2in = tb_alloc(b, (E + 1) / 2 + 1);
  b.o_in_src = tb_alloc(b, (E + 1) / 2 + 1);
  b.o_emlo = tb_alloc(b, EW);
  b.o_emhi = tb_alloc(b, EW);
  b.o_eimlo = tb_alloc(b, EW);
  b.o_eimhi = tb_alloc(b, EW);
  b.o_dmulti = tb_alloc(b, (size_t)2 * M);
  b.hot_words = b.words.size();
  /* cold section: read through L1/L2 (multi-edge weights, in-rows for envs with extra edges, observation rows) */
  b.o_mlo = tb_alloc(b, (size_t)M * W);
  b.o_mhi = tb_alloc(b, (size_t)M * W);
  b.o_adjT = tb_alloc(b, (size_t)M * W);
  b.o_mloT = tb_alloc(b, (size_t)M * W);
  b.o_mhiT = tb_alloc(b, (size_t)M * W);
  b.o_os = tb_alloc(b, M);
  b.o_ver = tb_alloc(b, M);
  uint32_t* w = b.words.data();
  uint16_t* col16 = (uint16_t*)(w + b.o_col);
  uint16_t* ineid16 = (uint16_t*)(w + b.o_in_eid);
  int32_t* rp = (int32_t*)(w + b.o_row_ptr);
  int32_t* ip = (int32_t*)(w + b.o_in_ptr);
  for (int i = 0; i <= M; i++) rp[i] = hn.row_ptr[i];
  std::vector<int> indeg(M + 1, 0);
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
      col16[e] = (uint16_t)v;
      w[b.o_adj + (size_t)u * W + (v >> 5)] |= 1u << (v & 31);
      w[b.o_adjT + (size_t)v * W + (u >> 5)] |= 1u << (u & 31);
      if ((mu - 1) & 1) { w[b.o_mlo + (size_t)u * W + (v >> 5)] |= 1u << (v & 31); w[b.o_mloT + (size_t)v * W + (u >> 5)] |= 1u << (u & 31); }
      if ((mu - 1) & 2) { w[b.o_mhi + (size_t)u * W + (v >> 5)] |= 1u << (v & 31); w[b.o_mhiT + (size_t)v * W + (u >> 5)] |= 1u << (u & 31); }
      if (mu > 1) {
        w[b.o_rowmulti + (u >> 5)] |= 1u << (u & 31);
        w[b.o_incmulti + (u >> 5)] |= 1u << (u & 31);
        w[b.o_incmulti + (v >> 5)] |= 1u << (v & 31);
      }
      indeg[v + 1]++;
    }
  }
  for (int i = 0; i < M; i++) indeg[i + 1] += indeg[i];
  for (int i = 0; i <= M; i++) ip[i] = indeg[i];
  std::vector<int> fill(M, 0);
  for (int u = 0; u < M; u++) /* ascending u => in-lists ascending by source */
    for (int e = hn.row_ptr[u]; e < hn.row_ptr[u + 1]; e++) {
      int v = hn.col[e];
      int j = indeg[v] + fill[v]++;
      ineid16[j] = (uint16_t)e;
      ((uint16_t*)(w + b.o_out2in))[e] = (uint16_t)j;
      ((uint16_t*)(w + b.o_in_src))[j] = (uint16_t)u;
      int mu = hn.mult ? hn.mult[e] : 1;
      if ((mu - 1) & 1) { w[b.o_emlo + (e >> 5)] |= 1u << (e & 31); w[b.o_eimlo + (j >> 5)] |= 1u << (j & 31); }
      if ((mu - 1) & 2) { w[b.o_emhi + (e >> 5)] |= 1u << (e & 31); w[b.o_eimhi + (j >> 5)] |= 1u << (j & 31); }
    }
  /* packed multi-edge entries of every device's out list and in list */
  for (int i = 0; i < M; i++) {
    for (int side = 0; side < 2; side++) {
      int lo = side ? ip[i] : rp[i], hi = side ? ip[i + 1] : rp[i + 1];
      uint32_t dm = 0xFFu | (0xFFu << 10);
      int cnt = 0;
      for (int p = lo; p < hi; p++) {
        int e = side ? ineid16[p] : p;
        int mu = hn.mult ? hn.mult[e] : 1;
        if (mu <= 1) continue;
        if (cnt >= 2 || p - lo >= 0xFF) { dm |= 0x80000000u; continue; }
        dm &= ~(0x3FFu << (10 * cnt));
        dm |= ((uint32_t)(p - lo) | ((uint32_t)(mu - 1) << 8)) << (10 * cnt);
        cnt++;
      }
      w[b.o_dmulti + 2 * i + side] = dm;
    }
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
  n.o_adj = (uint32_t)b.o_adj; n.o_adjT = (uint32_t)b.o_adjT;
  n.o_mlo = (uint32_t)b.o_mlo; n.o_mhi = (uint32_t)b.o_mhi; n.o_mloT = (uint32_t)b.o_mloT; n.o_mhiT = (uint32_t)b.o_mhiT;
  n.o_row_ptr = (uint32_t)b.o_row_ptr; n.o_col = (uint32_t)b.o_col; n.o_in_ptr = (uint32_t)b.o_in_ptr;
  n.o_in_eid = (uint32_t)b.o_in_eid; n.o_out2in = (uint32_t)b.o_out2in; n.o_in_src = (uint32_t)b.o_in_src; n.o_static = (uint32_t)b.o_static;
  n.o_dc = (uint32_t)b.o_dc; n.o_server = (uint32_t)b.o_server; n.o_reach = (uint32_t)b.o_reach; n.o_valid = (uint32_t)b.o_valid;
  n.o_rowmulti = (uint32_t)b.o_rowmulti; n.o_incmulti = (uint32_t)b.o_incmulti; n.o_napps = (uint32_t)b.o_napps;
  n.o_vuln = (uint32_t)b.o_vuln;
  n.o_dmulti = (uint32_t)b.o_dmulti;
  n.o_emlo = (uint32_t)b.o_emlo; n.o_emhi = (uint32_t)b.o_emhi; n.o_eimlo = (uint32_t)b.o_eimlo; n.o_eimhi = (uint32_t)b.o_eimhi;
  n.o_os = (uint32_t)b.o_os; n.o_ver = (uint32_t)b.o_ver;
}

}  // namespace cyg
#endif
