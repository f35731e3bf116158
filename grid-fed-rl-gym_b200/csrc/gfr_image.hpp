// gfr_image.hpp - host-side check of a gfr_feeder_desc and its packing into the "image" the
// kernels read (ints first, then doubles; offsets recorded in gfr::Layout).
#pragma once
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gfr_b200.h"
#include "gfr_device.cuh"

namespace gfr {

struct ImageBuilder {
  std::vector<int32_t> ints;
  std::vector<double> dbls;
  int add_i(const int32_t* p, int count) {
    int off = (int)ints.size();
    if (count > 0) ints.insert(ints.end(), p, p + count);
    return off;
  }
  int add_d(const double* p, int count) {
    int off = (int)dbls.size();
    if (count > 0) dbls.insert(dbls.end(), p, p + count);
    return off;
  }
};

struct FeederImage {
  Layout lay{};
  bool has_pv = false;
  bool root_is_slack = false;
  int center_depth = 0;            // levels of the tree rooted at its center (ceil(diameter / 2) + 1), uncapped
  std::vector<unsigned char> img;
  std::vector<double> load_pq;     // [2L] static active / reactive power (observation)
};

// returns an empty string on success, the complaint otherwise
// flat_start_factors: include the factorisation of the flat-start Jacobian (96 B per bus)
inline std::string build_feeder_image(const gfr_feeder_desc* d, FeederImage* out, bool flat_start_factors = true) {
  const int n = d->n_bus, nl = d->n_levels, L = d->n_load, G = d->n_gen, Bt = d->n_bat;
  if (n < 1 || nl < 1 || L < 0 || G < 0 || Bt < 0) return "bad feeder dimensions";
  if (!(d->s_base > 0.0)) return "s_base must be > 0";
  if (!d->order || !d->parent || !d->level_ptr || !d->child_ptr || (n > 1 && !d->child_idx) || !d->bus_type || !d->vm_set || !d->g ||
      !d->b || !d->gdiag || !d->bdiag || !d->r || !d->x || !d->line_of || !d->from_is_parent ||
      !d->rating || !d->load_profile)
    return "missing topology array";
  if ((L && (!d->load_bus || !d->load_base || !d->load_p || !d->load_q)) ||
      (G && (!d->gen_type || !d->gen_bus || !d->gen_cap || !d->gen_p0 || !d->gen_p1 || !d->gen_p2)) ||
      (Bt && (!d->bat_bus || !d->bat_cap || !d->bat_rating || !d->bat_eff || !d->bat_soc0)))
    return "missing component array";
  // ---- structure checks: root at k = 0, parents before children, one slack bus, child lists
  if (d->parent[0] != -1) return "k = 0 must be the root (parent -1)";
  if (d->level_ptr[0] != 0 || d->level_ptr[nl] != n) return "level_ptr must span [0, n]";
  if (d->level_ptr[1] != 1) return "level 0 must hold the root only";
  std::vector<int32_t> seen_ref(n, 0), seen_line(n > 1 ? n - 1 : 0, 0), seen_child(n, 0);
  std::vector<int> level(n, 0);
  for (int l = 0; l < nl; ++l) {
    if (d->level_ptr[l + 1] <= d->level_ptr[l]) return "empty level";
    for (int k = d->level_ptr[l]; k < d->level_ptr[l + 1]; ++k) level[k] = l;
  }
  int n_slack = 0;
  if (d->child_ptr[0] != 0 || d->child_ptr[n] != n - 1) return "child_ptr must span [0, n-1]";
  for (int k = 0; k < n; ++k) {
    if (d->order[k] < 0 || d->order[k] >= n || seen_ref[d->order[k]]++) return "order is not a permutation";
    if (d->child_ptr[k + 1] < d->child_ptr[k]) return "child_ptr must be non-decreasing";
    if (d->bus_type[k] != GFR_BUS_SLACK && d->bus_type[k] != GFR_BUS_PV && d->bus_type[k] != GFR_BUS_PQ)
      return "unknown bus_type";
    n_slack += d->bus_type[k] == GFR_BUS_SLACK;
    for (int q = d->child_ptr[k]; q < d->child_ptr[k + 1]; ++q) {
      const int c = d->child_idx[q];
      if (c <= k || c >= n || d->parent[c] != k || seen_child[c]++) return "child_idx does not match parent";
    }
    if (k > 0) {
      const int p = d->parent[k];
      if (p < 0 || p >= k) return "parent must precede its child in level order";
      if (level[p] >= level[k]) return "a bus must sit in a later level than its parent";
      const int li = d->line_of[k];
      if (li < 0 || li >= n - 1 || seen_line[li]++) return "line_of is not a permutation of the lines";
      if (!(d->g[k] == d->g[k]) || !(d->b[k] == d->b[k]) || (d->g[k] == 0.0 && d->b[k] == 0.0))
        return "branch with zero / NaN admittance";
    }
  }
  for (int k = 1; k < n; ++k) if (!seen_child[k]) return "child_idx does not list every bus";
  if (n_slack != 1) return "exactly one slack bus is required";
  for (int l = 0; l < L; ++l) if (d->load_bus[l] < 0 || d->load_bus[l] >= n) return "load_bus out of range";
  for (int g = 0; g < G; ++g) {
    if (d->gen_bus[g] < 0 || d->gen_bus[g] >= n) return "gen_bus out of range";
    if (d->gen_type[g] != GFR_GEN_SOLAR && d->gen_type[g] != GFR_GEN_WIND) return "unknown gen_type";
  }
  for (int b = 0; b < Bt; ++b) if (d->bat_bus[b] < 0 || d->bat_bus[b] >= n) return "bat_bus out of range";

  FeederImage* f = out;
  {
    // depth of the tree rooted at its center: the longest path (in edges) is found leaf -> root from the two
    // deepest subtrees of every bus (children come after their parent in level order)
    std::vector<int> h(n, 0);
    int diam = 0;
    for (int k = n - 1; k > 0; --k) {
      const int p = d->parent[k];
      diam = std::max(diam, h[p] + h[k] + 1);
      h[p] = std::max(h[p], h[k] + 1);
    }
    f->center_depth = (diam + 1) / 2 + 1;
  }
  Layout& lay = f->lay;
  lay.n = n; lay.nl = nl; lay.L = L; lay.G = G; lay.Bt = Bt; lay.A = Bt + G; lay.m = n - 1;
  lay.D = 2 * n + 2 * (n - 1) + 1 + 2 * L + G + 2 * Bt;
  lay.n_src = L + G + Bt; lay.R = R_BAT + 2 * Bt; lay.n_noise = 4 + L;
  lay.s_base = d->s_base;
  lay.inv_s_base = 1.0 / d->s_base;
  double lp = 0.0;
  for (int l = 0; l < L; ++l) lp = lp + d->load_p[l];   // sequential, as grid_env.py:744
  lay.load_p_sum = lp;

  ImageBuilder ib;
  std::vector<int32_t> flags(n), rank(n), bol(n > 1 ? n - 1 : 0);
  for (int k = 0; k < n; ++k) {
    int fl = 0;
    if (d->bus_type[k] == GFR_BUS_PQ) fl |= FL_PQ;
    else fl |= FL_FIXED_VM;
    if (d->bus_type[k] != GFR_BUS_SLACK) fl |= FL_THETA;
    if (d->bus_type[k] == GFR_BUS_PV) f->has_pv = true;
    if (d->from_is_parent[k]) fl |= FL_FROM_IS_PARENT;
    flags[k] = fl;
    rank[d->order[k]] = k;
    if (k > 0) bol[d->line_of[k]] = k;
  }
  // injection sources per bus: loads, then generators, then batteries (reference accumulation order)
  std::vector<int32_t> inj_ptr(n + 1, 0), inj_idx(lay.n_src);
  {
    std::vector<int> cnt(n, 0);
    for (int l = 0; l < L; ++l) cnt[d->load_bus[l]]++;
    for (int g = 0; g < G; ++g) cnt[d->gen_bus[g]]++;
    for (int b = 0; b < Bt; ++b) cnt[d->bat_bus[b]]++;
    for (int k = 0; k < n; ++k) inj_ptr[k + 1] = inj_ptr[k] + cnt[k];
    std::vector<int> fill(inj_ptr.begin(), inj_ptr.end() - 1);
    for (int l = 0; l < L; ++l) inj_idx[fill[d->load_bus[l]]++] = l;
    for (int g = 0; g < G; ++g) inj_idx[fill[d->gen_bus[g]]++] = L + g;
    for (int b = 0; b < Bt; ++b) inj_idx[fill[d->bat_bus[b]]++] = L + G + b;
  }
  // the slack bus and the path from it to the root of the elimination tree (sweep)
  {
    int ks = 0;
    for (int k = 0; k < n; ++k) if (d->bus_type[k] == GFR_BUS_SLACK) ks = k;
    lay.k_slack = ks;
    for (int k = ks; k > 0; k = d->parent[k]) flags[k] |= FL_SLACK_PATH;
  }
  // pool plan: given by the caller (checked by replaying the schedule) or one slot per bus
  std::vector<int32_t> pool_slot(n);
  if (d->pool_slot) {
    if (d->n_pool < 1 || d->n_pool > n) return "n_pool must be in [1, n]";
    std::vector<int> owner(d->n_pool, -1);        // which bus's contribution a slot holds
    for (int l = nl - 1; l >= 0; --l) {
      // a bus may take over a slot of one of its own children (it reads them before it writes);
      // any other slot it writes must have been free before this level started
      std::vector<int> before(owner);
      for (int k = d->level_ptr[l]; k < d->level_ptr[l + 1]; ++k) {
        const int sl = d->pool_slot[k];
        if (sl < 0 || sl >= d->n_pool) return "pool_slot out of range";
        const int prev = before[sl];
        if (prev >= 0 && d->parent[prev] != k) return "pool_slot reuses a slot that is still live";
        if (owner[sl] >= 0 && owner[sl] != prev) return "two buses of one level share a pool slot";
        owner[sl] = k;
        pool_slot[k] = sl;
      }
      for (int k = d->level_ptr[l]; k < d->level_ptr[l + 1]; ++k)
        for (int q = d->child_ptr[k]; q < d->child_ptr[k + 1]; ++q) {
          const int cs = d->pool_slot[d->child_idx[q]];
          if (owner[cs] == d->child_idx[q]) owner[cs] = -1;
        }
    }
    lay.n_pool = d->n_pool;
  } else {
    for (int k = 0; k < n; ++k) pool_slot[k] = k;
    lay.n_pool = n;
  }
  // per-bus topology records first (16-byte aligned at the image base)
  std::vector<int32_t> topo(4 * (size_t)n);
  for (int k = 0; k < n; ++k) {
    topo[4 * k + 0] = k > 0 ? d->parent[k] : 0;
    topo[4 * k + 1] = d->child_ptr[k];
    topo[4 * k + 2] = d->child_ptr[k + 1];
    topo[4 * k + 3] = flags[k];      // | pool slot << FL_POOL_SHIFT, added below
  }
  if (lay.n_pool > FL_POOL_MASK) return "more than 4095 pool slots";
  // The correction a bus needs on the way down (root -> leaf) arrives in field 0 of a pool slot.
  // Normally its parent scatters it into the slot of every child.  When the parent's first child c1
  // shares the parent's slot (inheritance) and is a leaf, nothing overwrites that slot between the
  // parent's turn and c1's, so every child that is handled no later than c1 simply reads the
  // parent's slot, and a parent all of whose other children do so skips the scatter loop.
  for (int k = 0; k < n; ++k) {
    int xs = pool_slot[k];
    if (k > 0) {
      const int p = d->parent[k];
      const int c1 = d->child_idx[d->child_ptr[p]];
      const bool c1_leaf = d->child_ptr[c1 + 1] == d->child_ptr[c1];
      if (c1_leaf && pool_slot[c1] == pool_slot[p] && level[k] <= level[c1]) xs = pool_slot[p];
    }
    topo[4 * k + 3] |= (pool_slot[k] << FL_POOL_SHIFT) | (int32_t)((uint32_t)xs << FL_XSLOT_SHIFT);
  }
  for (int k = 0; k < n; ++k) {
    const int q0 = d->child_ptr[k], q1 = d->child_ptr[k + 1];
    if (q1 == q0) continue;
    if (pool_slot[d->child_idx[q0]] == pool_slot[k]) topo[4 * k + 3] |= FL_INHERIT;
    bool all_read_mine = (topo[4 * k + 3] & FL_INHERIT) != 0;
    for (int q = q0 + 1; q < q1 && all_read_mine; ++q)
      all_read_mine = x_slot_of(topo[4 * d->child_idx[q] + 3]) == pool_slot[k] && pool_slot[d->child_idx[q]] != pool_slot[k];
    if (all_read_mine && q1 - q0 > 1) topo[4 * k + 3] |= FL_NO_SCATTER;
  }
  lay.o_topo = ib.add_i(topo.data(), 4 * n);
  lay.o_child_idx = ib.add_i(d->child_idx, n - 1);
  std::vector<int32_t> child_pool(n > 1 ? n - 1 : 0);
  for (int q = 0; q < n - 1; ++q) child_pool[q] = pool_slot[d->child_idx[q]];
  lay.o_child_pool = ib.add_i(child_pool.data(), n - 1);
  lay.o_level_ptr = ib.add_i(d->level_ptr, nl + 1);
  lay.o_rank = ib.add_i(rank.data(), n);
  lay.o_branch_of_line = ib.add_i(bol.data(), n - 1);
  lay.o_inj_ptr = ib.add_i(inj_ptr.data(), n + 1);
  lay.o_inj_idx = ib.add_i(inj_idx.data(), lay.n_src);
  lay.o_gen_type = ib.add_i(d->gen_type, G);
  const int n_int_padded = ((int)ib.ints.size() + 3) / 4 * 4;      // keep the doubles 16-byte aligned
  const int dbase = n_int_padded / 2;
  // paired arrays (one 128-bit load each): branch (g, b), diagonal (Re, Im Y_kk), branch (r, x)
  std::vector<double> gb(2 * (size_t)n), gbd(2 * (size_t)n), rx(2 * (size_t)n);
  for (int k = 0; k < n; ++k) {
    gb[2 * k] = k > 0 ? d->g[k] : 0.0; gb[2 * k + 1] = k > 0 ? d->b[k] : 0.0;   // the root has no branch
    gbd[2 * k] = d->gdiag[k]; gbd[2 * k + 1] = d->bdiag[k];
    rx[2 * k] = d->r[k]; rx[2 * k + 1] = d->x[k];
  }
  lay.o_gb = dbase + ib.add_d(gb.data(), 2 * n);
  lay.o_gbd = dbase + ib.add_d(gbd.data(), 2 * n);
  lay.o_rx = dbase + ib.add_d(rx.data(), 2 * n);
  // Factorisation of the flat-start Jacobian (the first Newton iteration of every instance):
  // the same leaf -> root block elimination the kernel runs, done once here.
  {
    std::vector<double> f0(12 * (size_t)n, 0.0), C(4 * (size_t)n, 0.0);
    bool ok = flat_start_factors;
    for (int l = nl - 1; l >= 0 && ok; --l) {
      for (int k = d->level_ptr[l]; k < d->level_ptr[l + 1]; ++k) {
        const bool th = d->bus_type[k] != GFR_BUS_SLACK, pq = d->bus_type[k] == GFR_BUS_PQ;
        const double vk = pq ? 1.0 : d->vm_set[k];
        auto vm = [&](int j) { return d->bus_type[j] == GFR_BUS_PQ ? 1.0 : d->vm_set[j]; };
        const double v2 = vk * vk;
        double P = d->gdiag[k] * v2, Q = -d->bdiag[k] * v2;
        if (k > 0) { const double a = vk * vm(d->parent[k]); P += -d->g[k] * a; Q += d->b[k] * a; }
        for (int q = d->child_ptr[k]; q < d->child_ptr[k + 1]; ++q) {
          const int c = d->child_idx[q];
          const double a = vk * vm(c);
          P += -d->g[c] * a; Q += d->b[c] * a;
        }
        double d00 = -Q - d->bdiag[k] * v2, d01 = P + d->gdiag[k] * v2, d10 = P - d->gdiag[k] * v2,
               d11 = Q - d->bdiag[k] * v2;
        for (int q = d->child_ptr[k]; q < d->child_ptr[k + 1]; ++q) {
          const double* cc = &C[4 * (size_t)d->child_idx[q]];
          d00 -= cc[0]; d01 -= cc[1]; d10 -= cc[2]; d11 -= cc[3];
        }
        const double a = k > 0 ? vk * vm(d->parent[k]) : 0.0;
        const double gk = k > 0 ? d->g[k] : 0.0, bk = k > 0 ? d->b[k] : 0.0;
        const double ga = -gk * a, al = bk * a, gl = -gk * a, ll = bk * a;      // sin = 0 at a flat start
        double u00 = al, u01 = ga, u10 = -ga, u11 = al;
        if (!th) { d00 = 1.0; d01 = 0.0; u00 = u01 = 0.0; }
        if (!pq) { d10 = 0.0; d11 = 1.0; u10 = u11 = 0.0; }
        const double det = d00 * d11 - d01 * d10;
        if (det == 0.0 || !(det == det)) { ok = false; break; }
        const double inv = 1.0 / det;
        const double i00 = d11 * inv, i01 = -d01 * inv, i10 = -d10 * inv, i11 = d00 * inv;
        const double m00 = i00 * u00 + i01 * u10, m01 = i00 * u01 + i01 * u11,
                     m10 = i10 * u00 + i11 * u10, m11 = i10 * u01 + i11 * u11;
        // six 16-byte fields per bus, field-major (field f of bus k at f0[2 * (f * n + k)]): the lanes
        // of a level read consecutive 16-byte units
        auto put = [&](int fld, double a0, double a1) {
          f0[2 * ((size_t)fld * n + k)] = a0; f0[2 * ((size_t)fld * n + k) + 1] = a1;
        };
        put(0, i00, i01); put(1, i10, i11);
        put(2, m00, m01); put(3, m10, m11);
        put(4, ll, gl);
        put(5, P, Q);                                  // calculated injections of the flat profile
        double* cc = &C[4 * (size_t)k];
        cc[0] = ll * m00 + gl * m10; cc[1] = ll * m01 + gl * m11;
        cc[2] = -gl * m00 + ll * m10; cc[3] = -gl * m01 + ll * m11;
      }
    }
    lay.o_f0 = ok ? dbase + ib.add_d(f0.data(), 12 * n) : -1;     // singular at the flat start: no shortcut
  }
  lay.o_rating = dbase + ib.add_d(d->rating, n);
  lay.o_vm_set = dbase + ib.add_d(d->vm_set, n);
  lay.o_load_base = dbase + ib.add_d(d->load_base, L);
  lay.o_gen_cap = dbase + ib.add_d(d->gen_cap, G);
  lay.o_gen_p0 = dbase + ib.add_d(d->gen_p0, G);
  lay.o_gen_p1 = dbase + ib.add_d(d->gen_p1, G);
  lay.o_gen_p2 = dbase + ib.add_d(d->gen_p2, G);
  lay.o_bat_cap = dbase + ib.add_d(d->bat_cap, Bt);
  lay.o_bat_rating = dbase + ib.add_d(d->bat_rating, Bt);
  lay.o_bat_eff = dbase + ib.add_d(d->bat_eff, Bt);
  lay.o_profile = dbase + ib.add_d(d->load_profile, 24);
  const size_t img_bytes = ((size_t)n_int_padded * 4 + ib.dbls.size() * 8 + 15) / 16 * 16;
  lay.img_bytes = (int)img_bytes;
  std::vector<unsigned char> img(img_bytes, 0);
  std::memcpy(img.data(), ib.ints.data(), ib.ints.size() * 4);
  std::memcpy(img.data() + (size_t)n_int_padded * 4, ib.dbls.data(), ib.dbls.size() * 8);

  std::vector<double> load_pq(2 * (size_t)L + 1, 0.0);
  for (int l = 0; l < L; ++l) { load_pq[2 * l] = d->load_p[l]; load_pq[2 * l + 1] = d->load_q[l]; }

  f->root_is_slack = d->bus_type[0] == GFR_BUS_SLACK;
  f->img.swap(img);
  f->load_pq.swap(load_pq);
  return std::string();
}

}  // namespace gfr
