/*
 * cyg_oracle.c -- CPU restatement of the CyGym step path (TEST INFRASTRUCTURE).
 *
 * This file is the checker, not the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  It restates,
 * one env at a time and in the reference's own statement order, what
 *   volt_typhoon_env.py:818-1333  (step)            :694-779 (step_grouped)
 *   volt_typhoon_env.py:612-692   (_step_apply_only) :330-383 (randomize_...)
 *   volt_typhoon_env.py:575-596, :141-145, :184-191, :193-245, :266-293 (arrivals)
 *   CDSimulator.py:244-348        (generate_workloads)
 *   CyberDefenseEnv.py:583-875    (evolve_network)   :555-578 (sample_action)
 *   CyberDefenseEnv.py:146-257    (observations)
 * compute, on the canonical struct-of-arrays layout of include/cygym_b200.h and
 * with the counter-based draw contract of oracle/draws.py.  It is pinned against
 * the UNMODIFIED reference by tests/test_oracle_vs_reference.py (live, in the
 * build container) and by the golden trajectories in tests/golden/ (everywhere).
 * Rewards are computed in double like the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/cygym_b200.h"

enum { /* oracle/draws.py site ids */
  SITE_STALL = 1, SITE_BLOCK = 2, SITE_UNBLOCK = 3, SITE_ZDAY = 4, SITE_PROBE = 5, SITE_WL_SAMPLE = 6,
  SITE_WL_TRI = 7, SITE_WL_LAZY = 8, SITE_EV_POISSON = 9, SITE_EV_ADD = 10, SITE_EV_PICK = 11,
  SITE_EV_ATT = 12, SITE_SHUFFLE = 13, SITE_DETECT = 14, SITE_SA_TYPE = 15, SITE_SA_NDEV = 16,
  SITE_SA_DEVS = 17, SITE_SA_EXP = 18, SITE_SA_APP = 19, N_SITES = 24
};

typedef struct {
  cyg_config cfg;
  int M, E, W, EW;
  int32_t *row_ptr, *col;
  uint8_t* mult;
  int32_t *in_ptr, *in_src, *in_eid; /* transpose of the base CSR (for _innbrs, volt:473) */
  uint32_t* dev_static;
  float *os_val, *ver_val;
  /* optional per-env side arrays set by cyo_set_aux(): hop-log ring [B][log_cap], detector slots + slot of env */
  uint32_t* logs;
  const uint32_t* det_slots;
  const int32_t* det_of_env;
} cyo_t;

/* ---- Philox4x32-10 (oracle/draws.py:philox4x32_10) ---------------------- */
static void philox(const uint32_t c[4], const uint32_t k[2], uint32_t o[4]) {
  uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], k0 = k[0], k1 = k[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

typedef struct {
  uint64_t seed;
  uint32_t env, epoch;
  uint32_t cnt[N_SITES];
} rng_t;

static uint32_t draw(rng_t* r, int site) {
  uint32_t k = r->cnt[site]++;
  uint32_t c[4] = {r->env, r->epoch, (uint32_t)site, k >> 2};
  uint32_t key[2] = {(uint32_t)r->seed, (uint32_t)(r->seed >> 32)};
  uint32_t o[4];
  philox(c, key, o);
  return o[k & 3];
}
static inline uint32_t below(uint32_t x, uint32_t n) { return (uint32_t)(((uint64_t)x * n) >> 32); }

void cyo_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox(ctr, key, out); }

/* ---- per-env view ------------------------------------------------------- */
typedef struct {
  const cyo_t* n;
  uint32_t *dev, *ckpt, *blocked, *extra, *scal;
  uint32_t* logs;      /* hop-log ring of this env (NULL: length only) */
  const uint32_t* det; /* detector slot of this env (NULL: none uploaded) */
  rng_t rng;
  double defcost, cleancost; /* float32 in HBM; accumulate in double per step then store */
} env_t;

#define BUSY(w) (((w) >> CYG_DEV_BUSY_SHIFT) & CYG_DEV_BUSY_MASK)
#define PT(w) (((w) >> CYG_DEV_PT_SHIFT) & CYG_DEV_PT_MASK)
#define CBY(w) (((w) >> CYG_DEV_CBY_SHIFT) & CYG_DEV_CBY_MASK)
static inline uint32_t set_busy(uint32_t w, uint32_t b) {
  if (b > CYG_DEV_BUSY_MASK) b = CYG_DEV_BUSY_MASK;
  return (w & ~(CYG_DEV_BUSY_MASK << CYG_DEV_BUSY_SHIFT)) | (b << CYG_DEV_BUSY_SHIFT);
}
static inline uint32_t drop_wl(uint32_t w) {
  return w & ~(CYG_DEV_HASWL | (CYG_DEV_PT_MASK << CYG_DEV_PT_SHIFT));
}
static inline uint32_t set_pt(uint32_t w, uint32_t pt) {
  return (w & ~(CYG_DEV_PT_MASK << CYG_DEV_PT_SHIFT)) | (pt << CYG_DEV_PT_SHIFT);
}
static inline uint32_t clr_cby(uint32_t w) { return w & ~(CYG_DEV_CBY_MASK << CYG_DEV_CBY_SHIFT); }
static inline int n_extra(const env_t* e) { return (int)(e->scal[CYG_S_PREV_X] >> 16); }

static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* _stall (volt_typhoon_env.py:135-138): random.randint(low, high) */
static uint32_t stall(env_t* e, int low, int high) {
  return (uint32_t)low + below(draw(&e->rng, SITE_STALL), (uint32_t)(high - low + 1));
}

/* ---- merged neighbour lists (base CSR + per-env extra edges) ------------ */
typedef struct { int v; int mult; int eid; /* base edge id, or -(1+extra index) */ } nbr_t;

static int out_nbrs(const env_t* e, int u, nbr_t* out) {
  const cyo_t* n = e->n;
  int cnt = 0;
  for (int i = n->row_ptr[u]; i < n->row_ptr[u + 1]; i++) { out[cnt].v = n->col[i]; out[cnt].mult = n->mult[i]; out[cnt].eid = i; cnt++; }
  int nx = n_extra(e);
  for (int j = 0; j < nx; j++) {
    uint32_t x = e->extra[j];
    if ((int)(x & CYG_X_IDMASK) != u) continue;
    int v = (int)((x >> CYG_X_V_SHIFT) & CYG_X_IDMASK);
    int p = cnt++;
    while (p > 0 && out[p - 1].v > v) { out[p] = out[p - 1]; p--; }
    out[p].v = v; out[p].mult = 1; out[p].eid = -(1 + j);
  }
  return cnt;
}
static int in_nbrs(const env_t* e, int u, nbr_t* out) {
  const cyo_t* n = e->n;
  int cnt = 0;
  for (int i = n->in_ptr[u]; i < n->in_ptr[u + 1]; i++) { out[cnt].v = n->in_src[i]; out[cnt].mult = n->mult[n->in_eid[i]]; out[cnt].eid = n->in_eid[i]; cnt++; }
  int nx = n_extra(e);
  for (int j = 0; j < nx; j++) {
    uint32_t x = e->extra[j];
    if ((int)((x >> CYG_X_V_SHIFT) & CYG_X_IDMASK) != u) continue;
    int s = (int)(x & CYG_X_IDMASK);
    int p = cnt++;
    while (p > 0 && out[p - 1].v > s) { out[p] = out[p - 1]; p--; }
    out[p].v = s; out[p].mult = 1; out[p].eid = -(1 + j);
  }
  return cnt;
}
static inline int is_blocked(const env_t* e, int eid) {
  if (eid >= 0) return (e->blocked[eid >> 5] >> (eid & 31)) & 1;
  return (e->extra[-eid - 1] & CYG_X_BLOCKED) != 0;
}
static inline void set_blocked(env_t* e, int eid, int b) {
  if (eid >= 0) {
    if (b) e->blocked[eid >> 5] |= 1u << (eid & 31); else e->blocked[eid >> 5] &= ~(1u << (eid & 31));
  } else {
    if (b) e->extra[-eid - 1] |= CYG_X_BLOCKED; else e->extra[-eid - 1] &= ~CYG_X_BLOCKED;
  }
}
static int has_edge(const env_t* e, int u, int v) { /* g.get_eid(u, v) != -1 */
  const cyo_t* n = e->n;
  for (int i = n->row_ptr[u]; i < n->row_ptr[u + 1]; i++) if (n->col[i] == v) return 1;
  int nx = n_extra(e);
  uint32_t key = (uint32_t)u | ((uint32_t)v << CYG_X_V_SHIFT);
  for (int j = 0; j < nx; j++) if ((e->extra[j] & 0xFFFFFFu) == key) return 1;
  return 0;
}

/* _rebuild_graph_cache (volt_typhoon_env.py:456-483): forgets every block (:476) */
static void rebuild_cache(env_t* e) {
  memset(e->blocked, 0, sizeof(uint32_t) * (size_t)e->n->EW);
  int nx = n_extra(e);
  for (int j = 0; j < nx; j++) e->extra[j] &= ~CYG_X_BLOCKED;
}

/* ---- evolve_network (CyberDefenseEnv.py:583-875) ------------------------ */
static int nth_member(const env_t* e, int want_active, uint32_t r) { /* r-th id (ascending) of the set */
  int M = e->n->M;
  for (int i = 0; i < M; i++) {
    int a = (e->dev[i] & CYG_DEV_ACTSET) != 0;
    if (a == want_active) { if (r == 0) return i; r--; }
  }
  return -1;
}
static void evolve_network(env_t* e) {
  const cyo_t* n = e->n;
  const cyg_config* c = &n->cfg;
  int M = n->M;
  uint32_t* scal = e->scal;
  if (!(scal[CYG_S_FLAGS] & CYG_FL_SETS_INIT)) { /* :654-659 */
    for (int i = 0; i < M; i++) {
      if (e->dev[i] & CYG_DEV_NYA) e->dev[i] &= ~CYG_DEV_ACTSET; else e->dev[i] |= CYG_DEV_ACTSET;
    }
    scal[CYG_S_FLAGS] |= CYG_FL_SETS_INIT;
  }
  int n_act = 0;
  for (int i = 0; i < M; i++) n_act += (e->dev[i] & CYG_DEV_ACTSET) != 0;
  uint32_t xp = draw(&e->rng, SITE_EV_POISSON); /* :668 */
  int num_events = 0;
  for (int j = 0; j < 16; j++) num_events += xp >= c->poisson_tab[j];
  int floor_n = c->num_of_device > c->min_network_size ? c->num_of_device : c->min_network_size;
  for (int ev = 0; ev < num_events; ev++) {
    uint32_t xa = draw(&e->rng, SITE_EV_ADD); /* :679 */
    if ((uint64_t)xa < c->thr_p_add) {
      int n_inact = M - n_act;
      if (n_inact > 0) {
        int node = nth_member(e, 0, below(draw(&e->rng, SITE_EV_PICK), (uint32_t)n_inact)); /* :675 */
        e->dev[node] &= ~CYG_DEV_NYA;
        e->dev[node] |= CYG_DEV_ACTSET;
        n_act++;
        uint32_t xt = draw(&e->rng, SITE_EV_ATT); /* :690 */
        if ((uint64_t)xt < c->thr_p_attacker) e->dev[node] |= CYG_DEV_COMP | CYG_DEV_OWNED | CYG_DEV_KNOWN;
      }
    } else if (n_act > floor_n) { /* :701-712 */
      int node = nth_member(e, 1, below(draw(&e->rng, SITE_EV_PICK), (uint32_t)n_act));
      uint32_t w = e->dev[node];
      w |= CYG_DEV_NYA | CYG_DEV_REMOVED;
      w = drop_wl(w);
      w = set_busy(w, 0);
      w &= ~CYG_DEV_ACTSET;
      e->dev[node] = w;
      n_act--;
    }
  }
  /* star among active attacker-owned devices (:738-774) */
  int hub = -1, changed = 0;
  for (int i = 0; i < M; i++) {
    if (!((e->dev[i] & CYG_DEV_OWNED) && (e->dev[i] & CYG_DEV_ACTSET))) continue;
    if (hub < 0) { hub = i; continue; }
    for (int dir = 0; dir < 2; dir++) {
      int u = dir ? i : hub, v = dir ? hub : i;
      if (has_edge(e, u, v)) continue;
      int nx = n_extra(e);
      if (nx >= c->xcap) { scal[CYG_S_FLAGS] |= CYG_FL_ERR_XCAP; continue; }
      e->extra[nx] = (uint32_t)u | ((uint32_t)v << CYG_X_V_SHIFT);
      scal[CYG_S_PREV_X] = (scal[CYG_S_PREV_X] & 0xFFFFu) | ((uint32_t)(nx + 1) << 16);
      changed = 1;
    }
  }
  /* the preferential-attachment repair (:776-843) needs a degree-0 vertex; networks
     are required to have min total degree >= 1 so it is dead. */
  if (changed) rebuild_cache(e);
}

/* ---- arrivals (volt_typhoon_env.py:575-596 and helpers) ------------------ */
static void generate_workloads(env_t* e, int num_loads, int server) { /* CDSimulator.py:244-348 */
  const cyo_t* n = e->n;
  int M = n->M;
  int n_active = 0;
  for (int i = 0; i < M; i++) n_active += !(e->dev[i] & CYG_DEV_NYA);
  if (n_active <= 0) return; /* volt:205-207 */
  if (n->cfg.wl_cap >= 0 && num_loads > n->cfg.wl_cap) num_loads = n->cfg.wl_cap; /* volt:210-211 */
  if (n->cfg.turbo) { /* cap + ramp (volt:219-231) */
    const cyg_config* c = &n->cfg;
    int frac_cap = (int)((server ? c->turbo_frac_servers : c->turbo_frac_clients) * (double)n_active);
    if (frac_cap < 1) frac_cap = 1;
    int hard_cap = server ? c->turbo_max_servers : c->turbo_max_clients;
    double ramp = (double)e->scal[CYG_S_STEP] / (double)(c->turbo_ramp_steps > 1 ? c->turbo_ramp_steps : 1);
    if (ramp > 1.0) ramp = 1.0;
    int base = frac_cap < hard_cap ? frac_cap : hard_cap;
    int turbo_cap = (int)rint((double)base * ramp); /* Python round(): half to even */
    if (turbo_cap < 1) turbo_cap = 1;
    if (num_loads > turbo_cap) num_loads = turbo_cap;
  }
  if (num_loads > n_active) num_loads = n_active; /* volt:234 */
  if (num_loads <= 0) return;
  int* cand = (int*)malloc(sizeof(int) * (size_t)M);
  int nc = 0;
  for (int i = 0; i < M; i++) {
    uint32_t w = e->dev[i];
    if (w & CYG_DEV_NYA) continue;
    if (w & CYG_DEV_HASWL) continue;
    if (BUSY(w) > 0) continue;
    int is_server = (n->dev_static[i] & CYG_ST_SERVER) != 0;
    if (is_server != server) continue;
    cand[nc++] = i;
  }
  int k = num_loads < nc ? num_loads : nc;
  for (int j = 0; j < k; j++) { /* random.sample: pop the r-th remaining candidate */
    uint32_t r = below(draw(&e->rng, SITE_WL_SAMPLE), (uint32_t)(nc - j));
    int did = cand[r];
    memmove(cand + r, cand + r + 1, sizeof(int) * (size_t)(nc - j - 1 - (int)r));
    uint32_t xt = draw(&e->rng, SITE_WL_TRI); /* CDSimulator.py:308 */
    int pt = 1;
    for (int v = 0; v < 8; v++) pt += xt >= n->cfg.tri_tab[v];
    if (pt > n->cfg.tri_high) pt = n->cfg.tri_high;
    e->dev[did] = set_pt(e->dev[did] | CYG_DEV_HASWL, (uint32_t)pt);
  }
  free(cand);
}

static void arrivals_if_due(env_t* e) {
  const cyo_t* n = e->n;
  const cyg_config* c = &n->cfg;
  int M = n->M;
  int n_active = 0, idle = 0, free_c = 0, free_s = 0;
  for (int i = 0; i < M; i++) {
    uint32_t w = e->dev[i];
    if (w & CYG_DEV_NYA) continue;
    n_active++;
    if (BUSY(w) == 0 && !(w & CYG_DEV_HASWL)) {
      idle++;
      if (n->dev_static[i] & CYG_ST_SERVER) free_s++; else free_c++;
    }
  }
  /* _arrival_period (volt:141-145) */
  int period = (int)(c->wl_period_base + 0.5 * sqrt((double)(n_active > 1 ? n_active : 1)));
  if (period < 10) period = 10;
  if (period > c->wl_period_max) period = c->wl_period_max;
  if (e->scal[CYG_S_STEP] % (uint32_t)period != 0) return;
  /* _idle_fraction() < 0.10 (volt:580): idle/active < 0.1  <=>  10*idle < active */
  if (n_active == 0 || 10 * idle < n_active) return;
  int nC = 100, nS = 10;
  if (c->scaling_vulnerability) { /* _scaled_numloads (volt:266-293) */
    int req_c = 2 * n_active;            /* round(100 * n_active/50) */
    int req_s = (2 * n_active + 5) / 10; /* round(10 * n_active/50): never a tie */
    if (req_c < 1) req_c = 1;
    if (req_s < 1) req_s = 1;
    int cap_c = free_c > 1 ? free_c : 1, cap_s = free_s > 1 ? free_s : 1;
    nC = req_c < cap_c ? req_c : cap_c;
    nS = req_s < cap_s ? req_s : cap_s;
  }
  if (c->wl_cap > 0) { /* volt:588-593 */
    int total = nC + nS;
    if (total > c->wl_cap) {
      double ratio = (double)c->wl_cap / (double)total;
      nC = (int)(nC * ratio); if (nC < 0) nC = 0;
      nS = (int)(nS * ratio); if (nS < 0) nS = 0;
    }
  }
  generate_workloads(e, nC, 0);
  generate_workloads(e, nS, 1);
}

/* ---- action decoding ----------------------------------------------------- */
typedef struct {
  int mode, atype, n_ex, ex[4], n_dev, app_index;
  int first; /* device_indices[0] when the set form carries it (hdr[2] >> 16, minus 1), else -1 */
  const uint32_t* mask;
  const uint16_t* order;
} act_t;

static void decode(const uint32_t* hdr, const uint32_t* mask, const uint16_t* order, act_t* a) {
  int at = (int)(hdr[0] & 0xFF);
  a->atype = at == (int)CYG_ATYPE_NONE ? -1000 : (int)(int8_t)at;
  a->mode = (int)((hdr[0] >> 8) & 1);
  a->n_ex = (int)((hdr[0] >> 16) & 0xFF);
  if (a->n_ex > 4) a->n_ex = 4;
  for (int i = 0; i < 4; i++) a->ex[i] = (int)(int8_t)((hdr[1] >> (8 * i)) & 0xFF);
  a->n_dev = (int)(hdr[2] & 0xFFFFu);
  a->first = (int)(hdr[2] >> 16) - 1;
  a->app_index = (int)hdr[3];
  a->mask = mask;
  a->order = order;
}
/* i-th entry of device_indices */
static int dev_at(const cyo_t* n, const act_t* a, int i, int* cursor) {
  if (a->order) return (int)a->order[i];
  int M = n->M;
  int d = *cursor;
  while (d < M && !((a->mask[d >> 5] >> (d & 31)) & 1)) d++;
  *cursor = d + 1;
  return d < M ? d : -1;
}

/* device_indices[0]: the order array's first entry, the header's first-drawn device (sample_action keeps the draw
 * order of random.sample only for this entry; CyberDefenseEnv.py:565), else the lowest listed id */
static int dev_first(const cyo_t* n, const act_t* a) {
  int cur = 0;
  if (!a->order && a->first >= 0 && a->first < n->M) return a->first;
  return dev_at(n, a, 0, &cur);
}

typedef struct { double cost; int dirty; } acc_t;

static void add_defcost(env_t* e, double d) { e->defcost += d; }
static void add_cleancost(env_t* e, double d) { e->cleancost += d; }

/* clean one device: volt_typhoon_env.py:996-1011 (and :676-690 in the grouped path) */
static void clean_device(env_t* e, int d, acc_t* acc) {
  const cyg_config* c = &e->n->cfg;
  uint32_t w = e->dev[d];
  if (w & CYG_DEV_OWNED) return;
  double ds = c->def_scale;
  int comp = (w & CYG_DEV_COMP) != 0;
  acc->cost += (comp ? 0.3 : -0.01) * ds;
  add_cleancost(e, (comp ? 0.3 : 0.01) * ds);
  add_defcost(e, (comp ? 0.3 : 0.01) * ds);
  e->scal[CYG_S_FLAGS] |= CBY(w) << CYG_FL_DISC_SHIFT; /* exp.discovered = True */
  w = clr_cby(w) & ~CYG_DEV_COMP;
  w = set_busy(w, stall(e, 0, c->default_high));
  w = drop_wl(w);
  e->dev[d] = w;
}

/* pool = out-edges then in-edges of `d` whose blocked flag == want (volt:501-511) */
static int pick_incident(env_t* e, int d, int want, int site, nbr_t* tmp) {
  int total = 0;
  int no = out_nbrs(e, d, tmp);
  for (int i = 0; i < no; i++) if (is_blocked(e, tmp[i].eid) == want) total += tmp[i].mult;
  int ni = in_nbrs(e, d, tmp + no);
  for (int i = 0; i < ni; i++) if (is_blocked(e, tmp[no + i].eid) == want) total += tmp[no + i].mult;
  if (total == 0) return 0x7FFFFFFF;
  int r = (int)below(draw(&e->rng, site), (uint32_t)total);
  for (int i = 0; i < no + ni; i++) {
    if (is_blocked(e, tmp[i].eid) != want) continue;
    if (r < tmp[i].mult) return tmp[i].eid;
    r -= tmp[i].mult;
  }
  return 0x7FFFFFFF;
}

/* meta actions shared by step (volt:918-976) and _step_apply_only (volt:627-668) */
static void defender_meta(env_t* e, const act_t* a, int atype, int grouped, acc_t* acc) {
  const cyo_t* n = e->n;
  const cyg_config* c = &n->cfg;
  int M = n->M;
  double ds = c->def_scale;
  uint32_t* scal = e->scal;
  if (atype == 2) {
    scal[CYG_S_CKPT]++;
    scal[CYG_S_FLAGS] |= CYG_FL_HAS_CKPT; /* checkpoint_variables: an alias, not a copy */
    acc->cost += -0.5 * a->n_dev * ds;
    add_defcost(e, 0.5 * a->n_dev * ds);
    for (int i = 0; i < M; i++) if (BUSY(e->dev[i]) > 0) e->dev[i] = set_busy(e->dev[i], BUSY(e->dev[i]) + 1);
  } else if (atype == 3) {
    scal[CYG_S_REVERT]++;
    if (scal[CYG_S_FLAGS] & CYG_FL_HAS_CKPT) {
      for (int i = 0; i < M; i++) {
        uint32_t w = set_busy(e->dev[i], stall(e, 0, c->default_high));
        e->dev[i] = drop_wl(w);
      }
      acc->cost += -1.0 * a->n_dev * ds;
      acc->dirty = 1;
    }
  } else if (atype == 10) {
    if (!grouped) { /* volt:946-953; the grouped variant has no busy bump (volt:650-659) */
      if (a->n_dev > 0) {
        int d = dev_first(n, a);
        e->dev[d] = set_busy(e->dev[d], BUSY(e->dev[d]) + 1);
      } else {
        for (int i = 0; i < M; i++) if (BUSY(e->dev[i]) > 0) e->dev[i] = set_busy(e->dev[i], BUSY(e->dev[i]) + 1);
      }
    }
    acc->cost += -1.0 * ds;
    if (scal[CYG_S_LOGS] > 0) { /* detector.train(last <= 2000 logs): the fit itself is scikit-learn's, on the host */
      scal[CYG_S_FLAGS] |= CYG_FL_DET_TRAINED;
      if (!c->turbo) scal[CYG_S_FLAGS] |= CYG_FL_DET_PENDING;
    }
  } else if (atype == 11) {
    int d = dev_first(n, a); /* host raises ValueError when n_dev == 0 (volt:965-966) */
    uint32_t w = e->dev[d];
    uint32_t k = CYG_CK_VALID;
    if (w & CYG_DEV_COMP) k |= CYG_CK_COMP;
    if (w & CYG_DEV_KNOWN) k |= CYG_CK_KNOWN;
    if (w & CYG_DEV_NYA) k |= CYG_CK_NYA;
    if (n->dev_static[d] & CYG_ST_REACH) k |= CYG_CK_REACH;
    if (w & CYG_DEV_HASWL) k |= CYG_CK_HASWL | (PT(w) << CYG_DEV_PT_SHIFT);
    k |= BUSY(w) << CYG_DEV_BUSY_SHIFT;
    k |= CBY(w) << CYG_DEV_CBY_SHIFT;
    e->ckpt[d] = k;
    scal[CYG_S_CKPT]++;
    acc->cost += -0.1 * ds;
    add_defcost(e, 0.1 * ds);
  }
}

static void restore_device(env_t* e, int d) { /* _apply_device_state (volt:430-437) */
  uint32_t k = e->ckpt[d], w = e->dev[d];
  w &= ~(CYG_DEV_COMP | CYG_DEV_KNOWN | CYG_DEV_NYA);
  if (k & CYG_CK_COMP) w |= CYG_DEV_COMP;
  if (k & CYG_CK_KNOWN) w |= CYG_DEV_KNOWN;
  if (k & CYG_CK_NYA) w |= CYG_DEV_NYA;
  w = drop_wl(w);
  if (k & CYG_CK_HASWL) w = set_pt(w | CYG_DEV_HASWL, PT(k));
  w = set_busy(w, BUSY(k));
  w = clr_cby(w) | (CBY(k) << CYG_DEV_CBY_SHIFT);
  e->dev[d] = w;
}

/* per-device defender actions (volt:989-1123) */
/* ---- trained detector (CDSimulator.py:714-723: batch_predict == IsolationForest.predict) -------------------------
 * The forest is fitted by scikit-learn on the host; a slot (include/cygym_b200.h, CYG_DET_*) holds its two trees and
 * the forest's verdict for every pair of leaves.  predict = two tree walks (sklearn: X[:, feature] <= threshold goes
 * left, X as float32 -- device ids are exact) + one table bit. */
static int det_leaf(const uint32_t* tree, double from, double to) {
  int node = 0;
  for (;;) {
    const uint32_t* nd = tree + 4 * node;
    int feat = (int)(nd[3] & 0xFFFFu);
    if (feat >= 2) return (int)(nd[3] >> 16);
    double thr;
    uint64_t bits = (uint64_t)nd[0] | ((uint64_t)nd[1] << 32);
    memcpy(&thr, &bits, 8);
    double x = feat == 0 ? from : to;
    node = (x <= thr) ? (int)(nd[2] & 0xFFFFu) : (int)(nd[2] >> 16);
  }
}
static int det_is_anomaly(const uint32_t* slot, int from, int to) {
  int l0 = det_leaf(slot + CYG_DET_TREE0, (double)from, (double)to);
  int l1 = det_leaf(slot + CYG_DET_TREE0 + CYG_DET_TREE_STRIDE, (double)from, (double)to);
  uint32_t idx = (uint32_t)l0 * slot[0] + (uint32_t)l1;
  return (int)((slot[CYG_DET_TABLE + (idx >> 5)] >> (idx & 31)) & 1u);
}
/* Iteration order of the CPython set {v_0, v_1, ...} built by inserting small non-negative ints in that order
 * (hash(v) == v): the reference walks `flagged_senders`, a set, and hands out one _stall draw per member
 * (volt:1062-1069).  Restates Objects/setobject.c (3.12): open addressing, 9 linear probes, perturb >>= 5,
 * i = i*5 + 1 + perturb, growth to the first power of two > 4*used when fill*5 >= mask*3.  Returns the member count. */
static int pyset_order(const int* vals, int n, int* out) {
  int table[256], tmp[256];
  int mask = 7, fill = 0;
  for (int i = 0; i <= mask; i++) table[i] = -1;
  for (int k = 0; k < n; k++) {
    const int key = vals[k];
    unsigned long long perturb = (unsigned long long)key;
    unsigned i = (unsigned)key & (unsigned)mask;
    int placed = 0;
    while (!placed) {
      int probes = (i + 9 <= (unsigned)mask) ? 9 : 0;
      unsigned j = i;
      do {
        if (table[j] < 0) { table[j] = key; fill++; placed = 1; break; }
        if (table[j] == key) { placed = 2; break; }
        j++;
      } while (probes--);
      if (placed) break;
      perturb >>= 5;
      i = (unsigned)((i * 5ull + 1ull + perturb) & (unsigned long long)mask);
    }
    if (placed == 1 && !((unsigned long long)fill * 5ull < (unsigned long long)mask * 3ull)) {
      int newsize = 8;
      while (newsize <= fill * 4) newsize <<= 1;
      const int newmask = newsize - 1;
      for (int q = 0; q <= newmask; q++) tmp[q] = -1;
      for (int q = 0; q <= mask; q++) {
        if (table[q] < 0) continue;
        const int kk = table[q];
        unsigned long long pb = (unsigned long long)kk;
        unsigned ii = (unsigned)kk & (unsigned)newmask;
        for (;;) {
          int probes = (ii + 9 <= (unsigned)newmask) ? 9 : 0;
          unsigned jj = ii;
          int ok = 0;
          do { if (tmp[jj] < 0) { tmp[jj] = kk; ok = 1; break; } jj++; } while (probes--);
          if (ok) break;
          pb >>= 5;
          ii = (unsigned)((ii * 5ull + 1ull + pb) & (unsigned long long)newmask);
        }
      }
      mask = newmask;
      for (int q = 0; q <= mask; q++) table[q] = tmp[q];
    }
  }
  int cnt = 0;
  for (int q = 0; q <= mask; q++) if (table[q] >= 0) out[cnt++] = table[q];
  return cnt;
}
int cyo_pyset_order(const int* vals, int n, int* out) { return pyset_order(vals, n, out); }

/* one iteration of defender action 5 with a TRAINED detector (volt:1052-1069): the last 30 log records are scored;
 * with a majority of anomalies every flagged sender is un-compromised and stalled (one _stall draw each, set order) */
static void scan_trained_once(env_t* e, acc_t* acc) {
  const cyg_config* c = &e->n->cfg;
  uint32_t nlogs = e->scal[CYG_S_LOGS];
  if (!e->det || !e->logs || (e->scal[CYG_S_FLAGS] & CYG_FL_DET_PENDING) || (c->log_cap < 30 && (uint32_t)c->log_cap < nlogs)) {
    e->scal[CYG_S_FLAGS] |= CYG_FL_ERR_DETECTOR;
    return;
  }
  int nw = nlogs < 30 ? (int)nlogs : 30, n_anom = 0, senders[30], ns = 0, order[30];
  for (int k = 0; k < nw; k++) {
    uint32_t r = e->logs[(nlogs - (uint32_t)nw + (uint32_t)k) % (uint32_t)c->log_cap];
    int from = (int)(r & 0xFFFFu), to = (int)(r >> 16);
    if (det_is_anomaly(e->det, from, to)) { n_anom++; senders[ns++] = from; }
  }
  if (n_anom >= nw / 2 + 1) {
    int cnt = pyset_order(senders, ns, order);
    for (int k = 0; k < cnt; k++) {
      int d = order[k];
      e->dev[d] &= ~CYG_DEV_COMP;
      e->dev[d] = set_busy(e->dev[d], stall(e, 0, c->default_high));
    }
  }
  (void)acc;
}

static void defender_per_device(env_t* e, const act_t* a, int atype, acc_t* acc, nbr_t* tmp) {
  const cyo_t* n = e->n;
  const cyg_config* c = &n->cfg;
  double ds = c->def_scale;
  uint32_t* scal = e->scal;
  int cur = 0;
  int dev0 = a->n_dev > 0 ? dev_first(n, a) : -1;
  for (int i = 0; i < a->n_dev; i++) {
    int d = dev_at(n, a, i, &cur);
    if (d < 0) break;
    if (e->dev[d] & CYG_DEV_NYA) continue;
    switch (atype) {
      case 1: clean_device(e, d, acc); break;
      case 4:
        acc->cost += -1.0 * ds;
        if (a->app_index >= 0 && a->app_index < (int)((n->dev_static[d] >> CYG_ST_NAPPS_SHIFT) & 0xFF))
          e->dev[d] = set_busy(e->dev[d], stall(e, 0, c->default_high));
        break;
      case 5:
        scal[CYG_S_SCAN]++;
        if (scal[CYG_S_LOGS] > 0) { /* window non-empty (volt:1052-1059); untrained detector -> all "D" */
          if ((scal[CYG_S_FLAGS] & CYG_FL_DET_TRAINED) && !c->turbo) scan_trained_once(e, acc); /* turbo: predictions = [] (volt:1055) */
          acc->cost += -0.5 * ds;
          add_defcost(e, 0.5 * ds);
        }
        break;
      case 6: {
        acc->cost += -0.5 * ds;
        add_defcost(e, 0.5 * ds);
        int eid = pick_incident(e, d, 0, SITE_BLOCK, tmp);
        if (eid != 0x7FFFFFFF) { set_blocked(e, eid, 1); scal[CYG_S_EBLK]++; acc->dirty = 1; }
      } break;
      case 7: {
        acc->cost += -0.5 * ds;
        uint32_t w = e->dev[d];
        w |= CYG_DEV_NYA;
        w &= ~CYG_DEV_COMP;
        w = clr_cby(w);
        w = drop_wl(w);
        e->dev[d] = w;
        acc->dirty = 1;
      } break;
      case 9: {
        acc->cost += -0.5 * ds;
        add_defcost(e, 0.5 * ds);
        int eid = pick_incident(e, d, 1, SITE_UNBLOCK, tmp);
        if (eid != 0x7FFFFFFF) { set_blocked(e, eid, 0); scal[CYG_S_EADD]++; acc->dirty = 1; }
      } break;
      case 12:
        if (e->ckpt[dev0] & CYG_CK_VALID) {
          restore_device(e, dev0);
          acc->cost += -1.0 * ds;
          add_defcost(e, 1.0 * ds);
        }
        break;
      case 13: {
        uint32_t w = e->dev[dev0];
        w &= ~CYG_DEV_COMP;
        w = clr_cby(w);
        w = drop_wl(w);
        w = set_busy(w, stall(e, 3, c->default_high + 3));
        e->dev[dev0] = w;
        acc->cost += -3.0 * ds;
        add_cleancost(e, 3.0 * ds);
        add_defcost(e, 3.0 * ds);
      } break;
      default: break;
    }
  }
}

/* attacker actions (volt:1126-1202) */
static void attacker_act(env_t* e, const act_t* a, int atype, acc_t* acc, nbr_t* tmp, int* sources) {
  const cyo_t* n = e->n;
  const cyg_config* c = &n->cfg;
  int M = n->M;
  int ns = 0;
  for (int i = 0; i < M; i++) if (e->dev[i] & (CYG_DEV_COMP | CYG_DEV_OWNED)) sources[ns++] = i;
  if (c->base_line == CYG_BL_NO_ATTACK) return;
  if (atype == 1) {
    for (int xi = 0; xi < a->n_ex; xi++) {
      int raw = a->ex[xi];
      if (c->zero_day && !(raw >= 0 && raw < 32 && ((c->zero_day_mask >> raw) & 1))) { /* volt:1135-1136 */
        int cnt = __builtin_popcount(c->zero_day_mask);
        int r = (int)below(draw(&e->rng, SITE_ZDAY), (uint32_t)cnt);
        for (int b = 0; b < 32; b++) if ((c->zero_day_mask >> b) & 1) { if (r == 0) { raw = b; break; } r--; }
      }
      if (!(raw >= 0 && raw < c->n_exploits)) continue; /* ids are strings: an int never matches (volt:1141) */
      for (int si = 0; si < ns; si++) {
        int s = sources[si];
        int is_dc = (n->dev_static[s] & CYG_ST_DC) != 0;
        int nn = out_nbrs(e, s, tmp);
        int done = 0;
        for (int j = 0; j < nn && !done; j++) {
          if (is_blocked(e, tmp[j].eid)) continue;
          int v = tmp[j].v;
          for (int rep = 0; rep < tmp[j].mult && !done; rep++) {
            if (e->logs) e->logs[e->scal[CYG_S_LOGS] % (uint32_t)c->log_cap] = (uint32_t)s | ((uint32_t)v << 16);
            e->scal[CYG_S_LOGS]++; /* log_communication (volt:1161 -> CDSimulator.py:667-673) */
            uint32_t w = e->dev[v];
            if (is_dc) {
              e->dev[v] = w | CYG_DEV_COMP | ((1u << raw) << CYG_DEV_CBY_SHIFT);
              done = 1;
            } else if (n->dev_static[v] & CYG_ST_REACH) {
              e->dev[v] = w | CYG_DEV_COMP;
              done = 1;
            } else if (!(w & CYG_DEV_COMP) && (w & CYG_DEV_KNOWN) && ((n->dev_static[v] >> (CYG_ST_VULN_SHIFT + raw)) & 1)) {
              e->dev[v] = w | CYG_DEV_COMP;
              done = 1;
            }
          }
        }
      }
    }
  } else if (atype == 2) {
    if (ns > 0) {
      int s = sources[below(draw(&e->rng, SITE_PROBE), (uint32_t)ns)];
      int nn = out_nbrs(e, s, tmp);
      for (int j = 0; j < nn; j++) {
        if (is_blocked(e, tmp[j].eid)) continue;
        int v = tmp[j].v;
        if (!(e->dev[v] & CYG_DEV_KNOWN)) { e->dev[v] |= CYG_DEV_KNOWN; acc->cost += 0.1; break; }
      }
    }
  }
}

static int workload_advance(env_t* e) { /* volt:1242-1261 */
  int M = e->n->M, cur = 0;
  for (int i = 0; i < M; i++) {
    uint32_t w = e->dev[i];
    if (BUSY(w) == 0 && !(w & CYG_DEV_NYA) && (w & CYG_DEV_HASWL) && PT(w) > 0) {
      uint32_t pt = PT(w) - 1;
      if (pt == 0) { w = drop_wl(w); e->scal[CYG_S_WORK]++; cur++; } else w = set_pt(w, pt);
      e->dev[i] = w;
    }
  }
  return cur;
}

static void count_comp(const env_t* e, int* n_comp, int* n_comp_dc) { /* _count_comp (volt:563-572) */
  int M = e->n->M, a = 0, b = 0;
  for (int i = 0; i < M; i++) {
    uint32_t w = e->dev[i];
    if ((w & CYG_DEV_COMP) && !(w & CYG_DEV_NYA) && !(w & CYG_DEV_OWNED)) { a++; if (e->n->dev_static[i] & CYG_ST_DC) b++; }
  }
  *n_comp = a; *n_comp_dc = b;
}

static void step_env(env_t* e, const uint32_t* hdr, const uint32_t* mask, const uint16_t* order, size_t hdr_gstride,
                     size_t mask_gstride, size_t order_gstride, int G, uint32_t flags, double* raw_out,
                     double* shaped_out, int32_t* done_out, int32_t* exec_atype, uint32_t* pre_masks) {
  const cyo_t* n = e->n;
  const cyg_config* c = &n->cfg;
  int M = n->M, W = n->W;
  uint32_t* scal = e->scal;
  memset(e->rng.cnt, 0, sizeof(e->rng.cnt));
  e->rng.epoch = scal[CYG_S_EPOCH];
  scal[CYG_S_EPOCH]++;
  e->defcost = (double)u2f(scal[CYG_S_DEFCOST]);
  e->cleancost = (double)u2f(scal[CYG_S_CLEANCOST]);
  nbr_t* tmp = (nbr_t*)malloc(sizeof(nbr_t) * (size_t)(2 * (n->E + 2 * c->xcap) + 4 * M + 16));
  int* sources = (int*)malloc(sizeof(int) * (size_t)M);
  acc_t acc = {0.0, 0};
  int grouped = (flags & CYG_STEP_GROUPED) != 0;
  int skip_work = (flags & CYG_STEP_SKIP_WORK) != 0;
  act_t a;
  decode(hdr, mask, order, &a);
  int mode = a.mode, atype = 0;
  if (!grouped) {
    atype = a.atype;
    if (atype == -1000) { /* action is None (volt:847-874) */
      if (mode == CYG_MODE_DEFENDER) { atype = (c->base_line == CYG_BL_NO_DEFENSE) ? 8 : 7; }
      else { atype = (c->base_line == CYG_BL_NO_ATTACK) ? 3 : 2; }
      a.n_dev = 0; a.first = -1; a.n_ex = 1; a.ex[0] = 0; a.app_index = 0;
    }
    if (mode == CYG_MODE_DEFENDER) { if (!(atype >= 0 && atype < c->def_space_n)) atype = 8; }
    else { if (!(atype >= 0 && atype < c->att_space_n)) atype = 3; }
    /* busy tick over _busy_devices (volt:904-908) */
    for (int i = 0; i < M; i++) {
      uint32_t w = e->dev[i];
      if ((w & CYG_DEV_BUSYSET) && BUSY(w) > 0) e->dev[i] = set_busy(w, BUSY(w) - 1);
    }
    if (mode == CYG_MODE_DEFENDER) {
      if (c->base_line != CYG_BL_NASH) atype = 8; /* volt:913-914 */
      defender_meta(e, &a, atype, 0, &acc);
      if (atype == 1 || atype == 4 || atype == 5 || atype == 6 || atype == 7 || atype == 9 || atype == 12 || atype == 13)
        defender_per_device(e, &a, atype, &acc, tmp);
    } else {
      attacker_act(e, &a, atype, &acc, tmp, sources);
    }
  } else {
    for (int g = 0; g < G; g++) { /* _step_apply_only (volt:612-692) */
      act_t ga;
      decode(hdr + g * hdr_gstride, mask + g * mask_gstride, order ? order + g * order_gstride : NULL, &ga);
      int gt = ga.atype == -1000 ? 0 : ga.atype;
      if (gt == 0) gt = (mode == CYG_MODE_DEFENDER) ? 8 : 3;
      if (mode == CYG_MODE_DEFENDER) {
        if (c->base_line != CYG_BL_NASH) gt = 8;
        defender_meta(e, &ga, gt, 1, &acc);
        if (gt == 1) {
          int cur = 0;
          for (int i = 0; i < ga.n_dev; i++) {
            int d = dev_at(n, &ga, i, &cur);
            if (d < 0) break;
            if (e->dev[d] & CYG_DEV_NYA) continue;
            clean_device(e, d, &acc);
          }
        }
      }
      atype = gt;
    }
    for (int i = 0; i < M; i++) { /* _tick_busy_time_once (volt:607-610) */
      uint32_t w = e->dev[i];
      if (BUSY(w) > 0) e->dev[i] = set_busy(w, BUSY(w) - 1);
    }
  }
  int cur_work = 0;
  if (!skip_work || grouped) {
    cur_work = workload_advance(e);
    arrivals_if_due(e);
  }
  double def_work = (double)c->work_scale * cur_work;
  int n_comp, n_comp_dc;
  count_comp(e, &n_comp, &n_comp_dc);
  if (!grouped) scal[CYG_S_COMPCNT] += (uint32_t)n_comp; /* volt:1267-1270; absent from step_grouped */
  double raw, shaped;
  if (mode == CYG_MODE_DEFENDER) {
    raw = acc.cost + def_work - n_comp * (double)c->comp_scale;
    shaped = raw;
  } else {
    raw = acc.cost + (double)c->comp_scale * (n_comp + 10 * n_comp_dc);
    double phi = (double)n_comp / (double)M;
    double gam = (double)c->gamma;
    uint32_t pn = scal[CYG_S_PREV_X] & 0xFFFFu;
    double prev = (pn == 0xFFFFu) ? phi : gam * ((double)pn / (double)M);
    double bonus = 0.1 * (gam * phi - prev);
    scal[CYG_S_PREV_X] = (scal[CYG_S_PREV_X] & 0xFFFF0000u) | (uint32_t)n_comp;
    shaped = raw + bonus;
  }
  if (pre_masks) {
    memset(pre_masks, 0, sizeof(uint32_t) * 3 * (size_t)W);
    for (int i = 0; i < M; i++) {
      uint32_t w = e->dev[i];
      if (w & CYG_DEV_COMP) pre_masks[0 * W + (i >> 5)] |= 1u << (i & 31);
      if (w & CYG_DEV_KNOWN) pre_masks[1 * W + (i >> 5)] |= 1u << (i & 31);
      if (w & CYG_DEV_NYA) pre_masks[2 * W + (i >> 5)] |= 1u << (i & 31);
    }
  }
  if (!skip_work || grouped) {
    scal[CYG_S_STEP]++;
    if (mode == CYG_MODE_ATTACKER) scal[CYG_S_ATT_STEP]++; else scal[CYG_S_DEF_STEP]++;
  }
  int done = scal[CYG_S_STEP] > 1000;
  int periodic = (scal[CYG_S_STEP] % (uint32_t)c->evolve_period) == 0;
  if (acc.dirty || periodic) evolve_network(e);
  if (!grouped) { /* volt:1330 */
    for (int i = 0; i < M; i++) {
      if (BUSY(e->dev[i]) > 0) e->dev[i] |= CYG_DEV_BUSYSET; else e->dev[i] &= ~CYG_DEV_BUSYSET;
    }
  }
  scal[CYG_S_DEFCOST] = f2u((float)e->defcost);
  scal[CYG_S_CLEANCOST] = f2u((float)e->cleancost);
  *raw_out = raw; *shaped_out = shaped; *done_out = done;
  if (exec_atype) *exec_atype = atype;
  free(tmp); free(sources);
}

/* randomize_compromise_and_ownership (volt_typhoon_env.py:330-383) */
static void randomize_env(env_t* e) {
  const cyo_t* n = e->n;
  int M = n->M;
  uint32_t* scal = e->scal;
  memset(e->rng.cnt, 0, sizeof(e->rng.cnt));
  e->rng.epoch = scal[CYG_S_EPOCH];
  scal[CYG_S_EPOCH]++;
  int* pool = (int*)malloc(sizeof(int) * (size_t)M);
  int np = 0, k_owned = 0, k_comp = 0;
  for (int i = 0; i < M; i++) {
    uint32_t w = e->dev[i];
    if ((w & CYG_DEV_NYA) || (n->dev_static[i] & CYG_ST_DC)) continue;
    pool[np++] = i;
    k_owned += (w & CYG_DEV_OWNED) != 0;
    k_comp += (w & CYG_DEV_COMP) != 0;
  }
  if (np == 0 || (k_owned == 0 && k_comp == 0)) { free(pool); return; }
  int extra = k_comp - k_owned; if (extra < 0) extra = 0;
  for (int i = 0; i < np; i++) e->dev[pool[i]] &= ~(CYG_DEV_OWNED | CYG_DEV_COMP | CYG_DEV_KNOWN);
  int rem = np;
  for (int j = 0; j < k_owned + extra && j < np; j++) { /* shuffle == successive uniform picks */
    uint32_t r = (rem > 1) ? below(draw(&e->rng, SITE_SHUFFLE), (uint32_t)rem) : 0;
    int d = pool[r];
    memmove(pool + r, pool + r + 1, sizeof(int) * (size_t)(rem - 1 - (int)r));
    rem--;
    if (j < k_owned) e->dev[d] |= CYG_DEV_OWNED | CYG_DEV_COMP | CYG_DEV_KNOWN;
    else e->dev[d] |= CYG_DEV_COMP | CYG_DEV_KNOWN;
  }
  free(pool);
}

/* sample_action (CyberDefenseEnv.py:555-578): device_indices as a set + ascending order */
static void sample_action_env(env_t* e, int mode, uint32_t* hdr, uint32_t* mask, uint16_t* order) {
  const cyo_t* n = e->n;
  const cyg_config* c = &n->cfg;
  int M = n->M, W = n->W;
  uint32_t* scal = e->scal;
  memset(e->rng.cnt, 0, sizeof(e->rng.cnt));
  e->rng.epoch = scal[CYG_S_EPOCH];
  scal[CYG_S_EPOCH]++;
  int space = mode == CYG_MODE_DEFENDER ? c->def_space_n : c->att_space_n;
  int atype = (int)below(draw(&e->rng, SITE_SA_TYPE), (uint32_t)space);
  int ndev = 1 + (int)below(draw(&e->rng, SITE_SA_NDEV), (uint32_t)c->num_of_device);
  int* pool = (int*)malloc(sizeof(int) * (size_t)M);
  for (int i = 0; i < M; i++) pool[i] = i;
  memset(mask, 0, sizeof(uint32_t) * (size_t)W);
  int rem = M, first = 0;
  for (int j = 0; j < ndev; j++) {
    uint32_t r = below(draw(&e->rng, SITE_SA_DEVS), (uint32_t)rem);
    int d = pool[r];
    memmove(pool + r, pool + r + 1, sizeof(int) * (size_t)(rem - 1 - (int)r));
    rem--;
    mask[d >> 5] |= 1u << (d & 31);
    if (order) order[j] = (uint16_t)d;
    if (j == 0) first = d;
  }
  free(pool);
  int ex = (int)below(draw(&e->rng, SITE_SA_EXP), (uint32_t)c->X);
  int app = c->n_app_ids > 0 ? (int)below(draw(&e->rng, SITE_SA_APP), (uint32_t)c->n_app_ids) : 0;
  hdr[0] = (uint32_t)(atype & 0xFF) | ((uint32_t)mode << 8) | (1u << 16);
  hdr[1] = (uint32_t)(ex & 0xFF);
  hdr[2] = (uint32_t)ndev | ((uint32_t)(first + 1) << 16); /* device_indices[0] = the first device drawn */
  hdr[3] = (uint32_t)app;
}

/* observations (CyberDefenseEnv.py:146-257) */
static void observe_env(const env_t* e, int obs_mode, float* out) {
  const cyo_t* n = e->n;
  int M = n->M, X = n->cfg.X;
  for (int i = 0; i < M; i++) {
    uint32_t w = e->dev[i];
    float os = n->os_val[i], ver = n->ver_val[i];
    float comp = (w & CYG_DEV_COMP) ? 1.f : 0.f, known = (w & CYG_DEV_KNOWN) ? 1.f : 0.f, nya = (w & CYG_DEV_NYA) ? 1.f : 0.f;
    int owned = (w & CYG_DEV_OWNED) != 0;
    if (obs_mode == 3) {
      float* r = out + 6 * i;
      r[0] = os; r[1] = ver; r[2] = comp; r[3] = 0.f; r[4] = known; r[5] = nya;
    } else if (obs_mode == 1) {
      float* r = out + 6 * i;
      if ((w & CYG_DEV_NYA) || !owned) { for (int k = 0; k < 6; k++) r[k] = -1.f; }
      else { r[0] = os; r[1] = ver; r[2] = -1.f; r[3] = 0.f; r[4] = known; r[5] = nya; }
    } else {
      float* r = out + 4 * i;
      if (!(w & CYG_DEV_KNOWN) || (w & CYG_DEV_NYA) || !owned) { for (int k = 0; k < 4; k++) r[k] = -1.f; }
      else { r[0] = os; r[1] = ver; r[2] = comp; r[3] = known; }
    }
  }
  if (obs_mode == 2) for (int k = 0; k < X; k++) out[4 * M + k] = k < n->cfg.n_exploits ? 1.f : 0.f;
}

/* ---- public (ctypes) surface --------------------------------------------- */
void* cyo_create(const cyg_config* cfg, const int32_t* row_ptr, const int32_t* col, const uint8_t* mult,
                 const uint32_t* dev_static, const float* os_val, const float* ver_val) {
  cyo_t* n = (cyo_t*)calloc(1, sizeof(cyo_t));
  n->cfg = *cfg;
  int M = n->M = cfg->M, E = n->E = cfg->E;
  n->W = (M + 31) / 32;
  n->EW = (E + 31) / 32;
  n->row_ptr = (int32_t*)malloc(sizeof(int32_t) * (size_t)(M + 1));
  memcpy(n->row_ptr, row_ptr, sizeof(int32_t) * (size_t)(M + 1));
  n->col = (int32_t*)malloc(sizeof(int32_t) * (size_t)(E + 1));
  memcpy(n->col, col, sizeof(int32_t) * (size_t)E);
  n->mult = (uint8_t*)malloc((size_t)E + 1);
  memcpy(n->mult, mult, (size_t)E);
  n->dev_static = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)M);
  memcpy(n->dev_static, dev_static, sizeof(uint32_t) * (size_t)M);
  n->os_val = (float*)malloc(sizeof(float) * (size_t)M);
  memcpy(n->os_val, os_val, sizeof(float) * (size_t)M);
  n->ver_val = (float*)malloc(sizeof(float) * (size_t)M);
  memcpy(n->ver_val, ver_val, sizeof(float) * (size_t)M);
  n->in_ptr = (int32_t*)calloc((size_t)M + 2, sizeof(int32_t));
  n->in_src = (int32_t*)malloc(sizeof(int32_t) * (size_t)(E + 1));
  n->in_eid = (int32_t*)malloc(sizeof(int32_t) * (size_t)(E + 1));
  for (int e = 0; e < E; e++) n->in_ptr[col[e] + 1]++;
  for (int i = 0; i < M; i++) n->in_ptr[i + 1] += n->in_ptr[i];
  int32_t* fill = (int32_t*)calloc((size_t)M + 1, sizeof(int32_t));
  for (int u = 0; u < M; u++)
    for (int e = row_ptr[u]; e < row_ptr[u + 1]; e++) {
      int v = col[e];
      int p = n->in_ptr[v] + fill[v]++;
      n->in_src[p] = u; n->in_eid[p] = e;
    }
  free(fill);
  return n;
}
void cyo_destroy(void* h) {
  cyo_t* n = (cyo_t*)h;
  if (!n) return;
  free(n->row_ptr); free(n->col); free(n->mult); free(n->dev_static); free(n->os_val); free(n->ver_val);
  free(n->in_ptr); free(n->in_src); free(n->in_eid); free(n);
}
void cyo_set_base_line(void* h, int32_t bl) { ((cyo_t*)h)->cfg.base_line = bl; }
void cyo_set_aux(void* h, uint32_t* logs, const uint32_t* det_slots, const int32_t* det_of_env) {
  cyo_t* n = (cyo_t*)h;
  n->logs = logs; n->det_slots = det_slots; n->det_of_env = det_of_env;
}

static void bind_env(env_t* e, const cyo_t* n, int b, int env_id0, uint32_t* dev, uint32_t* ckpt, uint32_t* blocked,
                     uint32_t* extra, uint32_t* scal) {
  e->n = n;
  e->dev = dev + (size_t)b * n->M;
  e->ckpt = ckpt ? ckpt + (size_t)b * n->M : NULL;
  e->blocked = blocked + (size_t)b * n->EW;
  e->extra = extra + (size_t)b * n->cfg.xcap;
  e->scal = scal + (size_t)b * CYG_NSCAL;
  e->logs = (n->logs && n->cfg.log_cap > 0) ? n->logs + (size_t)b * n->cfg.log_cap : NULL;
  e->det = (n->det_slots && n->det_of_env && n->det_of_env[b] >= 0) ? n->det_slots + (size_t)n->det_of_env[b] * CYG_DET_WORDS : NULL;
  e->rng.seed = n->cfg.seed;
  e->rng.env = (uint32_t)(env_id0 + b);
}

int cyo_step(void* h, int B, int env_id0, uint32_t* dev, uint32_t* ckpt, uint32_t* blocked, uint32_t* extra,
             uint32_t* scal, const uint32_t* hdr, const uint32_t* mask, const uint16_t* order, int order_stride, int G,
             uint32_t flags, double* raw, double* shaped, int32_t* done, int32_t* exec_atype, uint32_t* pre_masks,
             int n_threads) {
  const cyo_t* n = (const cyo_t*)h;
  int W = n->W;
#ifdef _OPENMP
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for num_threads(n_threads) schedule(static) if (n_threads > 1)
#endif
  for (int b = 0; b < B; b++) {
    env_t e;
    bind_env(&e, n, b, env_id0, dev, ckpt, blocked, extra, scal);
    step_env(&e, hdr + (size_t)b * 4, mask + (size_t)b * W, order ? order + (size_t)b * order_stride : NULL,
             (size_t)B * 4, (size_t)B * W, (size_t)B * order_stride, G, flags, raw + b, shaped + b, done + b,
             exec_atype ? exec_atype + b : NULL, pre_masks ? pre_masks + (size_t)b * 3 * W : NULL);
  }
  return 0;
}
int cyo_randomize(void* h, int B, int env_id0, uint32_t* dev, uint32_t* blocked, uint32_t* extra, uint32_t* scal,
                  const uint8_t* env_mask) {
  const cyo_t* n = (const cyo_t*)h;
  for (int b = 0; b < B; b++) {
    if (env_mask && !env_mask[b]) continue;
    env_t e;
    bind_env(&e, n, b, env_id0, dev, NULL, blocked, extra, scal);
    randomize_env(&e);
  }
  return 0;
}
int cyo_sample_actions(void* h, int B, int env_id0, uint32_t* scal, int mode, uint32_t* hdr, uint32_t* mask,
                       uint16_t* order, int order_stride) {
  const cyo_t* n = (const cyo_t*)h;
  static uint32_t dummy[1];
  for (int b = 0; b < B; b++) {
    env_t e;
    bind_env(&e, n, b, env_id0, dummy, NULL, dummy, dummy, scal);
    sample_action_env(&e, mode, hdr + (size_t)b * 4, mask + (size_t)b * n->W,
                      order ? order + (size_t)b * order_stride : NULL);
  }
  return 0;
}
int cyo_observe(void* h, int B, const uint32_t* dev, int obs_mode, float* obs) {
  const cyo_t* n = (const cyo_t*)h;
  int dim = obs_mode == 2 ? 4 * n->M + n->cfg.X : 6 * n->M;
  for (int b = 0; b < B; b++) {
    env_t e;
    memset(&e, 0, sizeof(e));
    e.n = n;
    e.dev = (uint32_t*)dev + (size_t)b * n->M;
    observe_env(&e, obs_mode, obs + (size_t)b * dim);
  }
  return 0;
}
int cyo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
