/*
 * cyg_coop.cuh -- warp-per-env forms of the draw-heavy defender actions (device only).
 *
 * Thread-per-env is the right mapping for the word-wide parts of a step, but four defender actions walk the
 * listed devices one by one, each with its own random draw: clean (volt_typhoon_env.py:996-1011), revert
 * (:928-943), upgrade (:1013-1018) and block / unblock (:1071-1100 -> :485-511).  One thread doing that for ~90
 * devices is a ~10^5-cycle dependent chain that the rest of the CTA waits for.  Here a whole warp takes ONE env
 * and its 32 lanes take 32 consecutive listed devices.
 *
 * Draws are addressable (oracle/draws.py: x = Philox(seed; env, epoch, site, k)), so lane i simply computes the
 * draw of ITS device, k = (draws consumed so far) + (rank of the device among those that draw).
 *
 *  - clean / revert / upgrade: every listed device draws, nothing a device does depends on another device, so
 *    the result is one ballot per busy_time bit-plane.
 *  - block / unblock is sequential in the reference: the pool of device i (its incident edges with the wanted
 *    blocked flag) shrinks when an EARLIER device flips an edge that also touches i.  The lanes pick
 *    speculatively against the state at the start of the round; a pick is invalid only if a lower lane picked an
 *    edge whose far endpoint is that lane's device.  The round commits the lanes below the first such lane and
 *    restarts from it, so the outcome is bit-identical to the sequential walk.
 *
 * All lanes call these functions with identical arguments; only lane 0 touches per-env scalars and cost sums.
 */
#ifndef CYG_COOP_CUH
#define CYG_COOP_CUH

#include "cyg_core.cuh"

namespace cyg {

#define CYG_FULL 0xFFFFFFFFu

__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31u); }
__device__ __forceinline__ uint32_t lanes_below(int lane) { return (1u << lane) - 1u; }

/* k-th draw of `site` in the current epoch */
__device__ __forceinline__ uint32_t draw_at(const Rng& r, int site, uint32_t k) {
  uint32_t o[4];
  philox4x32_10(r.env, r.epoch, (uint32_t)site, k >> 2, r.k0, r.k1, o);
  uint32_t j = k & 3u;
  return j == 0 ? o[0] : j == 1 ? o[1] : j == 2 ? o[2] : o[3];
}

template <int W>
struct Coop {
  typedef Env<W, true> E;

  /* busy_time = low + below(draw, range) for the devices of A[] (lane-uniform), draws k_base, k_base+1, ... in
   * ascending device order (the set form of _stall, volt:135-138) */
  static __device__ __forceinline__ void deposit(E& e, const uint32_t* A, int low, int high, uint32_t k_base) {
    const int lane = lane_id();
    const uint32_t range = (uint32_t)(high - low + 1);
    uint32_t kw = k_base;
    bool over = false;
    /* one Philox block per lane = draws [4 (k_base/4 + lane), +4): 128 draws cover every device of a W <= 4 env, so a
     * device fetches its draw from lane (k - k_base_aligned) / 4 by shuffle instead of running its own Philox */
    uint32_t blk[4];
    philox4x32_10(e.rng.env, e.rng.epoch, (uint32_t)SITE_STALL, (k_base >> 2) + (uint32_t)lane, e.rng.k0, e.rng.k1, blk);
    const uint32_t k_al = k_base & ~3u;
#pragma unroll
    for (int w = 0; w < W; w++) {
      const uint32_t aw = A[w];
      if (aw == 0) continue; /* uniform */
      const bool on = ((aw >> lane) & 1u) != 0;
      const uint32_t k = kw + (uint32_t)popc(aw & lanes_below(lane));
      const uint32_t rel = k - k_al; /* < 128 + 3 */
      const int from = (int)(rel >> 2) & 31;
      const uint32_t x0 = __shfl_sync(CYG_FULL, blk[0], from), x1 = __shfl_sync(CYG_FULL, blk[1], from);
      const uint32_t x2 = __shfl_sync(CYG_FULL, blk[2], from), x3 = __shfl_sync(CYG_FULL, blk[3], from);
      const uint32_t j = rel & 3u;
      uint32_t x = j == 0 ? x0 : j == 1 ? x1 : j == 2 ? x2 : x3;
      uint32_t v = 0;
      if (on) {
        if (rel >= 128u) x = draw_at(e.rng, SITE_STALL, k); /* only when k_base is not a multiple of 4 */
        v = (uint32_t)low + below(x, range);
        if (v > CYG_BUSY_MAX) { v = CYG_BUSY_MAX; over = true; }
      }
      const uint32_t r0 = __ballot_sync(CYG_FULL, v & 1u), r1 = __ballot_sync(CYG_FULL, v & 2u);
      const uint32_t r2 = __ballot_sync(CYG_FULL, v & 4u), r3 = __ballot_sync(CYG_FULL, v & 8u);
      if (lane == 0) {
        const uint32_t keep = ~aw;
        e.pl(P_BUSY0, w) = (e.pl(P_BUSY0, w) & keep) | r0;
        e.pl(P_BUSY0 + 1, w) = (e.pl(P_BUSY0 + 1, w) & keep) | r1;
        e.pl(P_BUSY0 + 2, w) = (e.pl(P_BUSY0 + 2, w) & keep) | r2;
        e.pl(P_BUSY0 + 3, w) = (e.pl(P_BUSY0 + 3, w) & keep) | r3;
      }
      kw += (uint32_t)popc(aw);
    }
    if (__any_sync(CYG_FULL, over) && lane == 0) e.scal(CYG_S_FLAGS) |= CYG_FL_ERR_BUSY;
    __syncwarp();
  }

  static __device__ __forceinline__ void set_blocked_atomic(E& e, int eid, bool b) {
    const int j = e.out2in(eid);
    uint32_t* bo = e.blocked() + (eid >> 5);
    uint32_t* bi = e.blocked_in() + (j >> 5);
    if (b) { atomicOr(bo, 1u << (eid & 31)); atomicOr(bi, 1u << (j & 31)); atomicAdd(&e.nblk(), 1u); }
    else { atomicAnd(bo, ~(1u << (eid & 31))); atomicAnd(bi, ~(1u << (j & 31))); atomicSub(&e.nblk(), 1u); }
  }

  /* block (6) / unblock (9) one incident edge per listed active device, in listed order.
   * A group of G lanes owns the env (the kernel uses G = 32; the first invalid pick of a round is typically 4-6
   * lanes in, but sub-warp groups diverge from each other and were measured slower). */
  template <int G>
  static __device__ __forceinline__ void flip(E& e, const typename E::Act& a, int atype, double& cost, bool& dirty) {
    const int lane = lane_id(), lg = lane % G, gbase = lane - lg;
    const uint32_t gm = (G == 32) ? CYG_FULL : (((1u << (G & 31)) - 1u) << gbase);
    const bool want = atype == 9;
    const int site = want ? SITE_UNBLOCK : SITE_BLOCK;
    const double ds = (double)e.n->cfg.def_scale;
    uint32_t act[W];
    const int na = e.listed_active(a, act);
    if (na == 0) return;
    if (lg == 0) { cost += -0.5 * ds * na; e.defcost += 0.5 * ds * na; }
    uint32_t cnt = 0;
    if (e.n_extra() > 0) {
      /* envs with extra (hub-star) edges (every env once randomize_compromise_and_ownership moved the owned set): the
       * same fixed-point passes with the pools in NEIGHBOUR-ID space, as flip_incident_general builds them (out_row /
       * in_row: base bit rows minus / intersected with the blocked pairs, plus the extra list; weights from the
       * multiplicity planes).  A pick is the directed pair (u, v); a lower lane's pick touches this lane's device d
       * as its out-edge (u == d: neighbour v leaves the out pool) or its in-edge (v == d: u leaves the in pool). */
      int pos = 0;
      uint32_t kbase = 0;
      while (pos < na) { /* uniform */
        const int idx = pos + lg;
        const bool valid = idx < na;
        const int d = valid ? e.select_nth(act, idx) : 0;
        const bool has_blk = e.any_blocked();
        uint32_t O0[W], I0[W], O[W], I[W];
        e.out_row(d, has_blk, want, O0);
        e.in_row(d, has_blk, want, I0);
#pragma unroll
        for (int q = 0; q < W; q++) { O0[q] = valid ? O0[q] : 0u; I0[q] = valid ? I0[q] : 0u; O[q] = O0[q]; I[q] = I0[q]; }
        const bool multi = e.devbit(e.n->o_incmulti, d);
        const uint32_t *lo = e.tc + e.n->o_mlo + d * W, *hi = e.tc + e.n->o_mhi + d * W;
        const uint32_t *loT = e.tc + e.n->o_mloT + d * W, *hiT = e.tc + e.n->o_mhiT + d * W;
        const uint32_t xd = draw_at(e.rng, site, kbase + (uint32_t)lg);
        int key = -2; /* u | v << 16 of the pick, -1: empty pool */
        uint32_t nem = 0;
        bool had_att = false;
        for (;;) {
          const int to = e.weight_below(O, lo, hi, multi, 32 * W), ti = e.weight_below(I, loT, hiT, multi, 32 * W);
          const int total = to + ti;
          const bool nonempty = total > 0;
          nem = __ballot_sync(gm, nonempty);
          const uint32_t x = __shfl_sync(gm, xd, popc(nem & lanes_below(lg)));
          int nkey = -1, other = -1;
          if (nonempty) {
            const int r = (int)below(x, (uint32_t)total);
            if (r < to) { other = e.weighted_select(O, lo, hi, multi, r); nkey = d | (other << 16); }
            else { other = e.weighted_select(I, loT, hiT, multi, r - to); nkey = other | (d << 16); }
          }
          const bool changed = __any_sync(gm, nkey != key);
          key = nkey;
          if (!changed) break;
          int vl = -1;
          if (nonempty) {
            const int ow = other >> 5;
            const uint32_t ob = 1u << (other & 31);
            uint32_t in = 0;
            int rank = 0;
#pragma unroll
            for (int w = 0; w < W; w++) {
              in |= act[w] & ob & eqmask(w, ow);
              rank += popc(act[w] & (w < ow ? 0xFFFFFFFFu : (ob - 1u) & eqmask(w, ow)));
            }
            const int l = rank - pos;
            if (in != 0 && l > lg && l < 32) vl = l;
          }
          const uint32_t bv = __ballot_sync(gm, vl >= 0);
          if (bv == 0 && !had_att) break;
          uint32_t att = bv;
#pragma unroll
          for (int k = 0; k < 5; k++) {
            const uint32_t bk = __ballot_sync(gm, vl >= 0 && ((vl >> k) & 1));
            att &= ((lg >> k) & 1) ? bk : ~bk;
          }
          had_att = bv != 0;
#pragma unroll
          for (int q = 0; q < W; q++) { O[q] = O0[q]; I[q] = I0[q]; }
          while (__any_sync(gm, att != 0)) {
            const int from = att ? (__ffs((int)att) - 1) : lg;
            const int kj = __shfl_sync(gm, key, from);
            if (att) {
              const int uj = kj & 0xFFFF, vj = kj >> 16;
              const bool mine_out = uj == d; /* else vj == d */
              const int nb_id = mine_out ? vj : uj;
              const uint32_t bit = 1u << (nb_id & 31);
#pragma unroll
              for (int q = 0; q < W; q++) {
                const uint32_t m = bit & eqmask(q, nb_id >> 5);
                O[q] &= ~(mine_out ? m : 0u);
                I[q] &= ~(mine_out ? 0u : m);
              }
              att &= att - 1u;
            }
          }
        }
        if (key >= 0) { /* commit: base pairs through both orders of the bitset, extra edges in their list */
          const int u = key & 0xFFFF, v = key >> 16;
          if ((e.adj(u, v >> 5) >> (v & 31)) & 1u) {
            set_blocked_atomic(e, e.base_eid(u, v), !want);
          } else {
            const int nx = e.n_extra();
            uint32_t* x = e.extra();
            const uint32_t k24 = (uint32_t)u | ((uint32_t)v << CYG_X_V_SHIFT);
            for (int j = 0; j < nx; j++)
              if ((x[j] & 0xFFFFFFu) == k24) { if (!want) atomicOr(&x[j], CYG_X_BLOCKED); else atomicAnd(&x[j], ~CYG_X_BLOCKED); break; }
          }
        }
        const uint32_t done = (uint32_t)popc(nem);
        cnt += done; kbase += done; pos += 32;
        __threadfence_block();
        __syncwarp(gm);
      }
    } else {
      /* Windows of 32 listed devices.  Lane i speculates the pick of device pos + i on the state at the start of the
       * window; a pick is wrong only if a LOWER lane flips an edge that ends at this lane's device (that edge leaves
       * the pool).  Each pass tells every lane which lower lanes currently do that ("attackers"), the lane drops those
       * edges from its start-of-window pool and picks again; the picks are a function of the lower lanes' picks only,
       * so the pass where nothing changes is the sequential walk's outcome (lane 0 is right after pass 1, lane i
       * after pass i + 1 at the latest; ~3 passes in practice) and all 32 flips commit at once.  The draw index of a
       * device = draws consumed + non-empty pools below it: lane l holds draw kbase + l and lanes fetch theirs by
       * shuffle, so a pool running empty mid-window only shifts the fetch. */
      int pos = 0;
      uint32_t kbase = 0;
      while (pos < na) { /* uniform */
#ifdef CYG_COUNT_ROUNDS
        e.dbg_rounds++;
#endif
        const int idx = pos + lg;
        const bool valid = idx < na;
        const int d = valid ? e.select_nth(act, idx) : 0;
        typename E::Pool P0, P;
        e.flip_windows(d, want, P0);
        if (!valid) {
#pragma unroll
          for (int q = 0; q < W; q++) { P0.xo[q] = 0; P0.xi[q] = 0; }
        }
        P = P0;
        const uint32_t xd = draw_at(e.rng, site, kbase + (uint32_t)lg);
        int eid = -2, other = -1;
        uint32_t nem = 0;
        bool had_att = false; /* some lane's pool of the current pass had edges removed (uniform) */
        for (;;) {
#ifdef CYG_COUNT_ROUNDS
          e.dbg_rounds += 0x10000;
#endif
          const int total = e.flip_weigh(P);
          const bool nonempty = total > 0;
          nem = __ballot_sync(gm, nonempty);
          const uint32_t x = __shfl_sync(gm, xd, popc(nem & lanes_below(lg)));
          int neid = -1, nother = -1;
          if (nonempty) neid = e.flip_pick(P, (int)below(x, (uint32_t)total), nother);
          const bool changed = __any_sync(gm, neid != eid);
          eid = neid; other = nother;
          if (!changed) break;
          /* victim of this lane's pick: the lane (above this one) whose device is the far endpoint */
          int vl = -1;
          if (nonempty) {
            const int ow = other >> 5;
            const uint32_t ob = 1u << (other & 31);
            uint32_t in = 0;
            int rank = 0;
#pragma unroll
            for (int w = 0; w < W; w++) {
              in |= act[w] & ob & eqmask(w, ow);
              rank += popc(act[w] & (w < ow ? 0xFFFFFFFFu : (ob - 1u) & eqmask(w, ow)));
            }
            const int l = rank - pos;
            if (in != 0 && l > lg && l < 32) vl = l;
          }
          const uint32_t bv = __ballot_sync(gm, vl >= 0);
          if (bv == 0 && !had_att) break; /* nobody is touched and the picks came from untouched pools: final */
          uint32_t att = bv;
#pragma unroll
          for (int k = 0; k < 5; k++) {
            const uint32_t bk = __ballot_sync(gm, vl >= 0 && ((vl >> k) & 1));
            att &= ((lg >> k) & 1) ? bk : ~bk;
          }
          had_att = bv != 0;
          P = P0;
          while (__any_sync(gm, att != 0)) {
            const int from = att ? (__ffs((int)att) - 1) : lg;
            const int ej = __shfl_sync(gm, eid, from);
            if (att) { e.pool_remove(P, ej); att &= att - 1u; }
          }
        }
        if (eid >= 0) set_blocked_atomic(e, eid, !want);
        const uint32_t done = (uint32_t)popc(nem);
        cnt += done; kbase += done; pos += 32;
        __syncwarp(gm);
      }
    }
    if (lg == 0 && cnt) {
      e.scal(want ? CYG_S_EADD : CYG_S_EBLK) += cnt;
      dirty = true;
    }
    __syncwarp(gm);
  }

  /* attacker action 1, exploit + lateral movement (volt:1126-1185), one warp per env: the 32 lanes take 32
   * consecutive sources of the snapshot.  A source's scan depends on earlier sources only through isCompromised
   * of its own hit (a "not yet compromised" hit is invalid if a lower lane hit the same device), so a round
   * commits the lanes below the first such lane and restarts from it: bit-identical to the sequential loop. */
  static __device__ __forceinline__ void attack(E& e, const typename E::Act& a) {
    const int lane = lane_id();
    uint32_t src[W];
    int ns = 0;
#pragma unroll
    for (int w = 0; w < W; w++) { src[w] = e.pl(P_COMP, w) | e.pl(P_OWNED, w); ns += popc(src[w]); }
    const bool has_blk = e.any_blocked();
    const int nx = e.n_extra();
    __syncwarp();
    /* envs with extra (hub-star) edges (nx > 0: every env after randomize_compromise_and_ownership moved the owned
     * set) take attack_source's general form, which materialises the unblocked row of a source from the base bit row,
     * the blocked pairs and the extra list; nx is per env, so the lanes of the warp stay on one form */
    uint32_t logs_add = 0, zk = 0;
    for (int xi = 0; xi < a.n_ex; xi++) { /* uniform */
      int raw = a.ex(xi);
      uint32_t zx = 0;
      if (e.needs_zday_draw(raw)) zx = draw_at(e.rng, SITE_ZDAY, zk++);
      raw = e.resolve_exploit(raw, zx);
      if (raw < 0) continue;
      uint32_t kv[W];
#pragma unroll
      for (int w = 0; w < W; w++) kv[w] = e.pl(P_KNOWN, w) & e.m_vuln(raw, w);
      int pos = 0;
      while (pos < ns) { /* uniform */
        const int idx = pos + lane;
        const bool valid = idx < ns;
        const int s = valid ? e.select_nth(src, idx) : 0;
        uint32_t comp[W];
#pragma unroll
        for (int w = 0; w < W; w++) comp[w] = e.pl(P_COMP, w);
        int cnt = 0, v = -1;
        bool rule3 = false;
        if (valid) v = e.attack_source(s, comp, kv, has_blk, nx, cnt, rule3);
        const uint32_t same = __match_any_sync(CYG_FULL, v >= 0 ? v : (0x1000 + lane));
        const bool conflict = valid && rule3 && (same & lanes_below(lane)) != 0;
        const uint32_t conf = __ballot_sync(CYG_FULL, conflict);
        int c = conf ? (__ffs((int)conf) - 1) : 32;
        if (c > ns - pos) c = ns - pos;
        if (valid && lane < c) {
          logs_add += (uint32_t)cnt + (v >= 0 ? 1u : 0u);
          if (v >= 0) {
            atomicOr(&e.pl(P_COMP, v >> 5), 1u << (v & 31));
            if (e.devbit(e.n->o_dc, s)) atomicOr(&e.pl(P_CBY0 + raw, v >> 5), 1u << (v & 31));
          }
        }
        pos += c;
        __syncwarp();
      }
    }
    logs_add = __reduce_add_sync(CYG_FULL, logs_add);
    if (lane == 0) e.scal(CYG_S_LOGS) += logs_add;
    __syncwarp();
  }

  static __device__ __forceinline__ bool is_heavy(int mode, int atype) {
    if (mode == CYG_MODE_ATTACKER) return atype == 1;
    return atype == 1 || atype == 3 || atype == 4 || atype == 6 || atype == 9;
  }

  /* the deposit-type heavy defender actions (clean / revert / upgrade) of a plain set-form step, one warp per env;
   * lane 0 owns cost / dirty / scalars */
  static __device__ __forceinline__ void defender(E& e, const typename E::Act& a, int atype, double& cost, bool& dirty) {
    const int lane = lane_id();
    const cyg_config& c = e.n->cfg;
    const double ds = (double)c.def_scale;
    if (atype == 1) {
      uint32_t A[W];
      e.clean_mask(a, A);
      __syncwarp();
      /* clean_scalar (rewards, discovery flags, state wipe; volt:996-1011) with one lane per (plane, word) */
      const int ncby = e.n->ncby, items = (5 + ncby) * W;
      int nc = 0, nu = 0;
      uint32_t disc = 0;
      for (int t = lane; t < items; t += 32) {
        const int pi = t / W, w = t - pi * W;
        uint32_t aw = 0;
#pragma unroll
        for (int q = 0; q < W; q++) aw |= A[q] & eqmask(q, w);
        const int p = pi == 0 ? P_COMP : pi == 1 ? P_HASWL : pi < 5 ? P_PT0 + (pi - 2) : P_CBY0 + (pi - 5);
        const uint32_t old = e.pl(p, w);
        e.pl(p, w) = old & ~aw;
        if (pi == 0) { nc += popc(aw & old); nu += popc(aw & ~old); }
        if (pi >= 5 && (old & aw) != 0) disc |= 1u << (pi - 5);
      }
      nc = __reduce_add_sync(CYG_FULL, nc);
      nu = __reduce_add_sync(CYG_FULL, nu);
      disc = __reduce_or_sync(CYG_FULL, disc);
      if (lane == 0) {
        cost += (nc * 0.3 - nu * 0.01) * ds;
        e.cleancost += (nc * 0.3 + nu * 0.01) * ds;
        e.defcost += (nc * 0.3 + nu * 0.01) * ds;
        e.scal(CYG_S_FLAGS) |= disc << CYG_FL_DISC_SHIFT; /* exp.discovered = True */
      }
      __syncwarp();
      deposit(e, A, 0, c.default_high, 0);
    } else if (atype == 3) {
      const bool has = (e.scal(CYG_S_FLAGS) & CYG_FL_HAS_CKPT) != 0;
      __syncwarp();
      if (lane == 0) e.scal(CYG_S_REVERT)++;
      if (has) {
        uint32_t A[W];
#pragma unroll
        for (int w = 0; w < W; w++) A[w] = e.m_valid(w);
        deposit(e, A, 0, c.default_high, 0);
        if (lane == 0) {
#pragma unroll
          for (int w = 0; w < W; w++) { e.pl(P_HASWL, w) = 0; e.pl(P_PT0, w) = 0; e.pl(P_PT0 + 1, w) = 0; e.pl(P_PT0 + 2, w) = 0; }
          cost += -1.0 * a.n_dev * ds;
          dirty = true;
        }
      }
    } else if (atype == 4) {
      uint32_t act[W], up[W];
      const int na = e.listed_active(a, act);
      if (na > 0) {
        e.upgrade_mask(act, a.app_index, up);
        if (lane == 0) cost += -1.0 * ds * na;
        deposit(e, up, 0, c.default_high, 0);
      }
    }
    __syncwarp();
  }
};

}  // namespace cyg
#endif
