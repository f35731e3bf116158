// gfr_device.cuh - per-instance math of the batched GridEnvironment.step path.
//
// One "group" of LANES threads (1, 2, 4, 8, 16 or 32, always inside one warp) owns one
// feeder instance at a time.  The instance's working set lives in shared memory, the
// compiled feeder (the "image") sits next to it, staged once per CTA with one bulk
// (TMA) copy.  Buses are numbered in LEVEL order (breadth-first from the slack bus, k = 0);
// lane `k % LANES` owns bus k in every phase, so a lane only ever needs a group barrier
// when it reads another bus's slots.
//
// The functions are __host__ __device__ so that tests/host_emu can run the LANES = 1
// instantiation on a CPU to debug control flow without a GPU.  That harness is test
// infrastructure; the shipped library only launches the __global__ kernels.
//
// Reference (paths under /root/reference/grid_fed_rl/):
//   Newton-Raphson        environments/power_flow.py:89-211  (polar; flat start; check-then-update)
//   Jacobian              environments/power_flow.py:213-295 (+ deviation D2)
//   update                environments/power_flow.py:297-327
//   line flows / losses   environments/power_flow.py:329-358, :199-200
//   env step              environments/grid_env.py:410-619 and the dynamics it calls
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GFR_HD __host__ __device__ __forceinline__
#else
#define GFR_HD inline
#endif

namespace gfr {

enum { SOLVER_SWEEP = 0, SOLVER_NEWTON = 1 };
enum { BUS_SLACK = 0, BUS_PV = 1, BUS_PQ = 2 };
enum { GEN_SOLAR = 0, GEN_WIND = 1 };
// per-bus slots of an instance's working set (field index)
enum { F_E = 0, F_F = 1, F_P = 2, F_M0 = 3, F_M1 = 4, F_M2 = 5, F_M3 = 6, F_V0 = 7, F_V1 = 8 };
enum { NF_NEWTON = 9, NF_SWEEP = 5, F_SCRATCH = 3 };
// bus flag bits
enum { FL_PQ = 1, FL_FROM_IS_PARENT = 2, FL_FIXED_VM = 4 };
// record (persistent per-instance state) slots, in doubles
enum { R_TIME = 0, R_FREQ, R_WIND, R_TEMP, R_CLOUD, R_TOTAL_LOSSES, R_EPISODE_REWARD, R_SEED,
       R_DRAWS, R_COUNTS, R_BAT };   // soc[Bt] then bpow[Bt] from R_BAT on

// Where everything is inside the feeder image (ints / doubles counted from the image base)
// plus the sizes; passed as a kernel parameter (constant bank).
struct Layout {
  int n, nl, L, G, Bt, A, D, m, n_src, R, img_bytes, n_noise;
  int o_parent, o_child_ptr, o_level_ptr, o_flags, o_order, o_rank, o_line_of,
      o_branch_of_line, o_inj_ptr, o_inj_idx, o_gen_type;
  int o_g, o_b, o_gdiag, o_bdiag, o_r, o_x, o_rating, o_vm_set, o_load_base, o_gen_cap,
      o_gen_p0, o_gen_p1, o_gen_p2, o_bat_cap, o_bat_rating, o_bat_eff, o_profile;
  double s_base, load_p_sum;
};

struct EnvCfg {
  double dt, v_min, v_max, f_min, f_max, penalty, load_noise, tol, accel;
  int episode_length, stochastic_loads, weather_variation, max_it;
};

struct SolveStat {
  double max_mismatch;
  int iterations;
  int converged;
};

// ----------------------------------------------------------------------------- group ops

template <int LANES>
struct Grp {
  int lane;        // lane inside the group
  int e;           // instance slot inside the CTA
  int E;           // instance slots per CTA
  int FS;          // field stride in doubles = ceil(n / LANES) * LANES * E
  unsigned mask;   // the group's lanes inside its warp
  double* st;      // CTA working set (shared memory)

  GFR_HD int sidx(int k) const {
    unsigned u = (unsigned)k;
    return (int)(((u / LANES) * (unsigned)E + (unsigned)e) * LANES + (u % LANES));
  }
  GFR_HD double& at(int field, int k) const { return st[field * FS + sidx(k)]; }
  GFR_HD double& scr(int j) const { return st[F_SCRATCH * FS + sidx(j)]; }
  // first index >= k0 owned by this lane
  GFR_HD int first(int k0) const { return k0 + ((lane - k0) & (LANES - 1)); }

  GFR_HD void sync() const {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) __syncwarp(mask);
#endif
  }
  GFR_HD double gmax_nan(double v) const {   // NaN-propagating max (numpy semantics)
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
      double w = __shfl_xor_sync(mask, v, o);
      v = (w > v || w != w) ? w : v;
    }
#endif
    return v;
  }
  GFR_HD double gmax(double v) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o));
#endif
    return v;
  }
  GFR_HD double gmin(double v) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(mask, v, o));
#endif
    return v;
  }
  GFR_HD double gsum(double v) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
#endif
    return v;
  }
  GFR_HD int gsum(int v) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
#endif
    return v;
  }
  GFR_HD int gor(int v) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(mask, v, o);
#endif
    return v;
  }
};

GFR_HD void sincos_small(double x, double* s, double* c) {
  // Newton corrections are small angles; a short Taylor pair is exact to < 1 ulp of 1.0
  // for |x| <= 0.25 (next terms: x^17/17! < 2e-25, x^18/18! < 1e-26)
  if (fabs(x) <= 0.25) {
    double x2 = x * x;
    double ps = -1.0 / 1307674368000.0;                 // -1/15!
    ps = ps * x2 + 1.0 / 6227020800.0;                  //  1/13!
    ps = ps * x2 - 1.0 / 39916800.0;                    // -1/11!
    ps = ps * x2 + 1.0 / 362880.0;                      //  1/9!
    ps = ps * x2 - 1.0 / 5040.0;                        // -1/7!
    ps = ps * x2 + 1.0 / 120.0;                         //  1/5!
    ps = ps * x2 - 1.0 / 6.0;                           // -1/3!
    *s = x + x * x2 * ps;
    double pc = 1.0 / 20922789888000.0;                 //  1/16!
    pc = pc * x2 - 1.0 / 87178291200.0;                 // -1/14!
    pc = pc * x2 + 1.0 / 479001600.0;                   //  1/12!
    pc = pc * x2 - 1.0 / 3628800.0;                     // -1/10!
    pc = pc * x2 + 1.0 / 40320.0;                       //  1/8!
    pc = pc * x2 - 1.0 / 720.0;                         // -1/6!
    pc = pc * x2 + 1.0 / 24.0;                          //  1/4!
    pc = pc * x2 - 0.5;
    *c = 1.0 + x2 * pc;
  } else {
    sincos(x, s, c);
  }
}

// ----------------------------------------------------------------------------- Philox4x32-10

struct U4 { uint32_t x, y, z, w; };

GFR_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

GFR_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    U4 t;
    t.x = hi1 ^ c.y ^ k0; t.y = lo1; t.z = hi0 ^ c.w ^ k1; t.w = lo0;
    c = t;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c;
}

GFR_HD double u53(uint32_t hi, uint32_t lo) {   // (0,1) on a 2^-53 grid, never 0 or 1
  uint64_t k = ((uint64_t)(hi >> 5) << 26) + (uint64_t)(lo >> 6);
  return ((double)k + 0.5) * (1.0 / 9007199254740992.0);
}

// block q of the noise row keyed (seed, draw): q = 0 -> uniform in *a; q >= 1 -> two normals
GFR_HD void noise_block(uint64_t seed, uint64_t draw, uint32_t q, double* a, double* b) {
  U4 c; c.x = (uint32_t)draw; c.y = (uint32_t)(draw >> 32); c.z = q; c.w = 0u;
  U4 w = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  double u1 = u53(w.x, w.y), u2 = u53(w.z, w.w);
  if (q == 0u) { *a = u1; *b = u2; return; }
  double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincos(6.283185307179586 * u2, &sn, &cs);
  *a = rad * cs; *b = rad * sn;
}

// normal for noise slot s >= 1 (slot 0 is the uniform)
GFR_HD double noise_slot(uint64_t seed, uint64_t draw, int s) {
  double a, b;
  if (s == 0) { noise_block(seed, draw, 0u, &a, &b); return a; }
  int j = s - 1;
  noise_block(seed, draw, (uint32_t)(1 + j / 2), &a, &b);
  return (j & 1) ? b : a;
}

// ----------------------------------------------------------------------------- solvers

#define GFR_II(name) (simg[lay.name])          /* int array base inside the image */
#define GFR_DI(name) (dimg[lay.name])          /* double array base */

// Flat start (power_flow.py:103, :131): 1.0 at 0 rad, slack / PV buses at their set magnitude.
template <int LANES>
GFR_HD void flat_start(const Grp<LANES>& g, const Layout& lay, const int* simg, const double* dimg,
                       int nf) {
  for (int k = g.lane; k < lay.n; k += LANES) {
    int fl = simg[lay.o_flags + k];
    g.at(F_E, k) = (fl & FL_FIXED_VM) ? dimg[lay.o_vm_set + k] : 1.0;
    g.at(F_F, k) = 0.0;
  }
  if (g.lane == 0 && nf > F_V1) { g.at(F_V0, 0) = 0.0; g.at(F_V1, 0) = 0.0; }   // slack correction = 0
  g.sync();
}

// Polar Newton-Raphson on a radial feeder.  Same iterates as the reference's dense solve:
// the Jacobian of a tree has one 2x2 block per bus and two per branch, so eliminating buses
// leaf -> root has no fill-in.  Unknowns per bus are (d theta, d|V| / |V|): scaling the
// |V| column removes every division from the assembly and leaves the update unchanged.
//   off-diagonal block J[i,j] = [[ al, ga ], [ -ga, al ]],
//        al = |Vi||Vj| (G_ij sin th_ij - B_ij cos th_ij), ga = |Vi||Vj| (G_ij cos th_ij + B_ij sin th_ij)
//   diagonal block  J[i,i]   = [[ -Q_i - B_ii |Vi|^2, P_i + G_ii |Vi|^2 ], [ P_i - G_ii |Vi|^2, Q_i - B_ii |Vi|^2 ]]
// with G_ij + jB_ij = -(g + jb) of the branch and everything written on e + jf = |V| e^{j theta}:
//   |Vi||Vj| cos th_ij = ei ej + fi fj,  |Vi||Vj| sin th_ij = fi ej - ei fj   (no trigonometry).
template <int LANES>
GFR_HD void newton_solve(const Grp<LANES>& g, const Layout& lay, const int* simg,
                         const double* dimg, double tol, int max_it, double accel,
                         SolveStat* out) {
  const int n = lay.n, nl = lay.nl;
  const int* parent = simg + lay.o_parent;
  const int* child_ptr = simg + lay.o_child_ptr;
  const int* level_ptr = simg + lay.o_level_ptr;
  const int* flags = simg + lay.o_flags;
  const double* bg = dimg + lay.o_g;
  const double* bb = dimg + lay.o_b;
  const double* gdiag = dimg + lay.o_gdiag;
  const double* bdiag = dimg + lay.o_bdiag;

  out->converged = 0;
  out->iterations = max_it;
  out->max_mismatch = INFINITY;

  for (int it = 0; it < max_it; ++it) {
    // ---- mismatch + diagonal blocks, every bus independently (power_flow.py:150-166, 213-295)
    double mm = 0.0;
    for (int k = g.first(1); k < n; k += LANES) {
      double ek = g.at(F_E, k), fk = g.at(F_F, k);
      double v2 = ek * ek + fk * fk;
      double gd = gdiag[k], bd = bdiag[k];
      double P = gd * v2, Q = -bd * v2;
      {
        int p = parent[k];
        double ep = g.at(F_E, p), fp = g.at(F_F, p);
        double a = ek * ep + fk * fp, s = fk * ep - ek * fp;
        P += -bg[k] * a - bb[k] * s;
        Q += -bg[k] * s + bb[k] * a;
      }
      for (int c = child_ptr[k]; c < child_ptr[k + 1]; ++c) {
        double ec = g.at(F_E, c), fc = g.at(F_F, c);
        double a = ek * ec + fk * fc, s = fk * ec - ek * fc;
        P += -bg[c] * a - bb[c] * s;
        Q += -bg[c] * s + bb[c] * a;
      }
      int pq = flags[k] & FL_PQ;
      double dP = g.at(F_P, k) - P;
      double dQ = pq ? (0.0 - Q) : 0.0;
      double aP = fabs(dP), aQ = fabs(dQ);
      double loc = (aQ > aP || aQ != aQ) ? aQ : aP;
      mm = (loc > mm || loc != loc) ? loc : mm;
      g.at(F_M0, k) = -Q - bd * v2;
      g.at(F_M1, k) = P + gd * v2;
      g.at(F_M2, k) = pq ? (P - gd * v2) : 0.0;
      g.at(F_M3, k) = pq ? (Q - bd * v2) : 1.0;
      g.at(F_V0, k) = dP;
      g.at(F_V1, k) = dQ;
    }
    mm = g.gmax_nan(mm);
    out->max_mismatch = mm;
    if (mm < tol) {                                   // checked before the update (:168-171)
      out->converged = 1;
      out->iterations = it + 1;
      break;
    }
    // ---- eliminate leaf -> root
    int singular = 0;
    for (int l = nl - 1; l >= 1; --l) {
      const int k1 = level_ptr[l + 1];
      for (int k = g.first(level_ptr[l]); k < k1; k += LANES) {
        double ek = g.at(F_E, k), fk = g.at(F_F, k);
        double d00 = g.at(F_M0, k), d01 = g.at(F_M1, k), d10 = g.at(F_M2, k), d11 = g.at(F_M3, k);
        double r0 = g.at(F_V0, k), r1 = g.at(F_V1, k);
        int pq = flags[k] & FL_PQ;
        for (int c = child_ptr[k]; c < child_ptr[k + 1]; ++c) {
          double ec = g.at(F_E, c), fc = g.at(F_F, c);
          double a = ek * ec + fk * fc, s = fk * ec - ek * fc;
          double ga = -bg[c] * a - bb[c] * s, al = -bg[c] * s + bb[c] * a;   // J[k,c]
          double m00 = g.at(F_M0, c), m01 = g.at(F_M1, c), m10 = g.at(F_M2, c), m11 = g.at(F_M3, c);
          double v0 = g.at(F_V0, c), v1 = g.at(F_V1, c);
          d00 -= al * m00 + ga * m10;
          d01 -= al * m01 + ga * m11;
          r0 -= al * v0 + ga * v1;
          if (pq) {
            d10 -= -ga * m00 + al * m10;
            d11 -= -ga * m01 + al * m11;
            r1 -= -ga * v0 + al * v1;
          }
        }
        int p = parent[k];
        double ep = g.at(F_E, p), fp = g.at(F_F, p);
        double a = ek * ep + fk * fp, s = fk * ep - ek * fp;
        double ga = -bg[k] * a - bb[k] * s, al = -bg[k] * s + bb[k] * a;     // J[k,p]
        double u00 = al, u01 = ga, u10 = pq ? -ga : 0.0, u11 = pq ? al : 0.0;
        double det = d00 * d11 - d01 * d10;
        if (det == 0.0) singular = 1;                 // dgesv's exact-zero pivot (:188-190)
        double inv = 1.0 / det;
        double i00 = d11 * inv, i01 = -d01 * inv, i10 = -d10 * inv, i11 = d00 * inv;
        g.at(F_M0, k) = i00 * u00 + i01 * u10;
        g.at(F_M1, k) = i00 * u01 + i01 * u11;
        g.at(F_M2, k) = i10 * u00 + i11 * u10;
        g.at(F_M3, k) = i10 * u01 + i11 * u11;
        g.at(F_V0, k) = i00 * r0 + i01 * r1;
        g.at(F_V1, k) = i10 * r0 + i11 * r1;
      }
      g.sync();
    }
    if (g.gor(singular)) {
      out->iterations = it + 1;
      break;
    }
    // ---- back-substitute root -> leaf: x_k = v_k - M_k x_parent  (x_slack = 0)
    for (int l = 1; l < nl; ++l) {
      const int k1 = level_ptr[l + 1];
      for (int k = g.first(level_ptr[l]); k < k1; k += LANES) {
        int p = parent[k];
        double x0 = g.at(F_V0, p), x1 = g.at(F_V1, p);
        g.at(F_V0, k) -= g.at(F_M0, k) * x0 + g.at(F_M1, k) * x1;
        g.at(F_V1, k) -= g.at(F_M2, k) * x0 + g.at(F_M3, k) * x1;
      }
      g.sync();
    }
    // ---- polar update, every bus independently (:297-327):
    //      theta += a dtheta, |V| += a d|V|  <=>  V *= (1 + a x1) e^{j a x0}
    for (int k = g.first(1); k < n; k += LANES) {
      double sn, cs;
      sincos_small(accel * g.at(F_V0, k), &sn, &cs);
      double sc = 1.0 + accel * g.at(F_V1, k);
      double ek = g.at(F_E, k), fk = g.at(F_F, k);
      g.at(F_E, k) = sc * (ek * cs - fk * sn);
      g.at(F_F, k) = sc * (ek * sn + fk * cs);
    }
    g.sync();
  }
}

// Backward / forward sweep on the same tree (no counterpart in the reference, SURVEY F6;
// compared with the reference's Newton-Raphson at tight tolerance).  Constant-power
// injections P + j0; convergence on max(|de|, |df|) over buses.
template <int LANES>
GFR_HD void sweep_solve(const Grp<LANES>& g, const Layout& lay, const int* simg,
                        const double* dimg, double tol, int max_it, SolveStat* out) {
  const int n = lay.n, nl = lay.nl;
  const int* parent = simg + lay.o_parent;
  const int* child_ptr = simg + lay.o_child_ptr;
  const int* level_ptr = simg + lay.o_level_ptr;
  const double* br = dimg + lay.o_r;
  const double* bx = dimg + lay.o_x;
  (void)n;
  out->converged = 0;
  out->iterations = max_it;
  out->max_mismatch = INFINITY;
  for (int it = 0; it < max_it; ++it) {
    // backward: branch current into bus k = its own draw plus its children's
    for (int l = nl - 1; l >= 1; --l) {
      const int k1 = level_ptr[l + 1];
      for (int k = g.first(level_ptr[l]); k < k1; k += LANES) {
        double ek = g.at(F_E, k), fk = g.at(F_F, k);
        double w = g.at(F_P, k) / (ek * ek + fk * fk);      // injected current = conj(S / V) = P V / |V|^2
        double jr = -w * ek, ji = -w * fk;
        for (int c = child_ptr[k]; c < child_ptr[k + 1]; ++c) {
          jr += g.at(F_M0, c);
          ji += g.at(F_M1, c);
        }
        g.at(F_M0, k) = jr;
        g.at(F_M1, k) = ji;
      }
      g.sync();
    }
    // forward: V_k = V_parent - z_k J_k
    double mm = 0.0;
    for (int l = 1; l < nl; ++l) {
      const int k1 = level_ptr[l + 1];
      for (int k = g.first(level_ptr[l]); k < k1; k += LANES) {
        int p = parent[k];
        double jr = g.at(F_M0, k), ji = g.at(F_M1, k);
        double en = g.at(F_E, p) - (br[k] * jr - bx[k] * ji);
        double fn = g.at(F_F, p) - (br[k] * ji + bx[k] * jr);
        double de = fabs(en - g.at(F_E, k)), df = fabs(fn - g.at(F_F, k));
        double loc = (df > de || df != df) ? df : de;
        mm = (loc > mm || loc != loc) ? loc : mm;
        g.at(F_E, k) = en;
        g.at(F_F, k) = fn;
      }
      g.sync();
    }
    mm = g.gmax_nan(mm);
    out->max_mismatch = mm;
    if (mm < tol) {
      out->converged = 1;
      out->iterations = it + 1;
      break;
    }
  }
}

// From -> to flow of the branch above bus k (power_flow.py:329-358): P (pu), |S| (pu), series loss (pu)
template <int LANES>
GFR_HD void branch_flow(const Grp<LANES>& g, const Layout& lay, const int* simg, const double* dimg,
                        int k, double* p_ft, double* s_abs, double* loss) {
  int p = simg[lay.o_parent + k];
  double ek = g.at(F_E, k), fk = g.at(F_F, k), ep = g.at(F_E, p), fp = g.at(F_F, p);
  double gg = dimg[lay.o_g + k], bb = dimg[lay.o_b + k];
  double de = ep - ek, df = fp - fk;                 // V_parent - V_k
  double ir = gg * de - bb * df, ii = gg * df + bb * de;   // current parent -> k
  double P, Q;
  if (simg[lay.o_flags + k] & FL_FROM_IS_PARENT) {
    P = ep * ir + fp * ii; Q = fp * ir - ep * ii;    // V_p conj(I)
  } else {
    P = -(ek * ir + fk * ii); Q = -(fk * ir - ek * ii);    // V_k conj(-I)
  }
  *p_ft = P;
  *s_abs = sqrt(P * P + Q * Q);
  *loss = gg * (de * de + df * df);                  // Re sum_i V_i conj((YV)_i), branch by branch
}

// ----------------------------------------------------------------------------- solver entry (gfr_solve)

struct SolOut {
  uint8_t* converged; int32_t* iterations; double* bus_voltages; double* bus_angles;
  double* line_flows; double* line_loadings; double* losses; double* max_mismatch;
};

template <int LANES, int SOLVER>
GFR_HD void solve_instance(const Grp<LANES>& g, const Layout& lay, const int* simg,
                           const double* dimg, const EnvCfg& cfg, int nf, long long env,
                           const double* p_inj, const SolOut& o) {
  const int n = lay.n, m = lay.m;
  const int* rank = simg + lay.o_rank;
  const double* pin = p_inj + env * n;
  for (int i = g.lane; i < n; i += LANES) g.at(F_P, rank[i]) = pin[i];
  flat_start(g, lay, simg, dimg, nf);
  SolveStat st;
  if (SOLVER == SOLVER_NEWTON) newton_solve(g, lay, simg, dimg, cfg.tol, cfg.max_it, cfg.accel, &st);
  else sweep_solve(g, lay, simg, dimg, cfg.tol, cfg.max_it, &st);
  for (int i = g.lane; i < n; i += LANES) {
    int k = rank[i];
    double e = g.at(F_E, k), f = g.at(F_F, k);
    if (o.bus_voltages) o.bus_voltages[env * n + i] = sqrt(e * e + f * f);
    if (o.bus_angles) o.bus_angles[env * n + i] = atan2(f, e);
  }
  double loss = 0.0;
  const int* bol = simg + lay.o_branch_of_line;
  for (int li = g.lane; li < m; li += LANES) {
    int k = bol[li];
    double P, S, ls;
    branch_flow(g, lay, simg, dimg, k, &P, &S, &ls);
    loss += ls;
    double rating = dimg[lay.o_rating + k];
    if (o.line_flows) o.line_flows[env * m + li] = P;
    if (o.line_loadings) o.line_loadings[env * m + li] = rating > 0.0 ? S * lay.s_base / rating : 0.0;
  }
  loss = g.gsum(loss);
  if (g.lane == 0) {
    if (o.losses) o.losses[env] = loss;
    if (o.max_mismatch) o.max_mismatch[env] = st.max_mismatch;
    if (o.converged) o.converged[env] = (uint8_t)st.converged;
    if (o.iterations) o.iterations[env] = st.iterations;
  }
  g.sync();
}

// ----------------------------------------------------------------------------- environment

struct StepOut {
  double* reward; uint8_t* terminated; uint8_t* truncated; uint8_t* error; uint8_t* converged;
  int32_t* iterations; double* max_voltage; double* min_voltage; double* losses;
  double* max_mismatch; uint8_t* violations; int32_t* violation_count; int32_t* current_step;
  double* episode_reward; double* noise_used;
};

GFR_HD bool finite_d(double x) { return fabs(x) <= 1.7976931348623157e308; }   // false for NaN / Inf

GFR_HD double clampd(double x, double lo, double hi) {    // np.maximum(lo, np.minimum(hi, x)), NaN stays NaN
  x = (x > hi) ? hi : x;
  return (x < lo) ? lo : x;
}

GFR_HD double hour_of(double t) { return fmod(t / 3600.0, 24.0); }   // (t / 3600) % 24, t >= 0

// SolarPVModel.get_power / WindTurbineModel.get_power (dynamics.py:120-142, :158-170)
GFR_HD double renewable_power(const Layout& lay, const int* simg, const double* dimg, int gi,
                              double hour, double wind, double temp, double cloud) {
  double cap = dimg[lay.o_gen_cap + gi];
  if (simg[lay.o_gen_type + gi] == GEN_SOLAR) {
    double area = dimg[lay.o_gen_p0 + gi], eff = dimg[lay.o_gen_p1 + gi];
    bool day = (hour >= 6.0) && (hour <= 18.0);
    double sun = day ? sin(3.141592653589793 * (hour - 6.0) / 12.0) : 0.0;
    double actual = (1000.0 * sun) * (1.0 - 0.8 * cloud);
    double dtc = temp - 25.0;
    double tf = 1.0 - 0.004 * (dtc > 0.0 ? dtc : 0.0);
    double p = actual * area * eff * tf;
    return p < cap ? p : cap;
  }
  double ci = dimg[lay.o_gen_p0 + gi], vr = dimg[lay.o_gen_p1 + gi], co = dimg[lay.o_gen_p2 + gi];
  if (wind < ci || wind > co) return 0.0;
  if (wind <= vr) {
    double q = (wind - ci) / (vr - ci);
    return cap * (q * q * q);
  }
  return cap;
}

// _update_weather (grid_env.py:653-681); u, z1..z3 are the four draws
GFR_HD void update_weather(double hour, double u, double z1, double z2, double z3, double* wind,
                           double* temp, double* cloud) {
  (void)u;   // irradiance is drawn but never read downstream
  *wind = clampd(*wind + (0.0 + z1 * 0.5), 0.0, 30.0);
  *temp = (25.0 + 10.0 * sin(2.0 * 3.141592653589793 * (hour - 12.0) / 24.0)) + (0.0 + z2 * 2.0);
  *cloud = clampd(*cloud + (0.0 + z3 * 0.1), 0.0, 1.0);
}

template <int LANES, int SOLVER>
GFR_HD void step_instance(const Grp<LANES>& g, const Layout& lay, const int* simg,
                          const double* dimg, const EnvCfg& cfg, int nf, long long env,
                          double* state, double* obs, const double* actions, const double* noise,
                          const StepOut& o) {
  const int n = lay.n, m = lay.m, L = lay.L, G = lay.G, Bt = lay.Bt, A = lay.A, D = lay.D;
  double* rec = state + env * lay.R;
  double* ob = obs + env * D;
  const double* act = actions + env * A;
  const int o_line = 2 * n, o_freq = 2 * n + 2 * m, o_gen = o_freq + 1 + 2 * L, o_bat = o_gen + G;

  // ---- action check (grid_env.py:424-428, 454-467): NaN / Inf -> -2 penalty, terminated,
  //      nothing advances.  A one-element action is replaced by 0.0 instead.
  int bad = 0;
  for (int a = g.lane; a < A; a += LANES) bad |= !finite_d(act[a]);
  bad = g.gor(bad);
  const bool zero_action = bad && (A == 1);
  if (zero_action) bad = 0;

  int step = ((const int32_t*)(rec + R_COUNTS))[0];
  int viol_count = ((const int32_t*)(rec + R_COUNTS))[1];
  double episode_reward = rec[R_EPISODE_REWARD];

  if (bad) {
    double vmax = -INFINITY, vmin = INFINITY;
    for (int i = g.lane; i < n; i += LANES) {
      double v = ob[2 * i];
      vmax = fmax(vmax, v); vmin = fmin(vmin, v);
    }
    vmax = g.gmax(vmax); vmin = g.gmin(vmin);
    {
      // get_observation() recomputes the renewable outputs from the clock and the weather
      // (grid_env.py:772-776); every other entry is state this path leaves alone
      const double hour = hour_of(rec[R_TIME]);
      for (int gi = g.lane; gi < G; gi += LANES)
        ob[o_gen + gi] = renewable_power(lay, simg, dimg, gi, hour, rec[R_WIND], rec[R_TEMP], rec[R_CLOUD]);
    }
    if (g.lane == 0) {
      if (o.reward) o.reward[env] = -cfg.penalty * 2.0;
      if (o.terminated) o.terminated[env] = 1;
      if (o.truncated) o.truncated[env] = 0;
      if (o.error) o.error[env] = 1;
      if (o.converged) o.converged[env] = 0;
      if (o.iterations) o.iterations[env] = 0;
      if (o.max_voltage) o.max_voltage[env] = vmax;
      if (o.min_voltage) o.min_voltage[env] = vmin;
      if (o.losses) o.losses[env] = 0.0;
      if (o.max_mismatch) o.max_mismatch[env] = 0.0;
      if (o.violations) { for (int q = 0; q < 4; ++q) o.violations[env * 4 + q] = 0; }
      if (o.violation_count) o.violation_count[env] = viol_count;
      if (o.current_step) o.current_step[env] = step;
      if (o.episode_reward) o.episode_reward[env] = episode_reward;
    }
    if (o.noise_used)
      for (int s = g.lane; s < lay.n_noise; s += LANES) o.noise_used[env * lay.n_noise + s] = 0.0;
    return;
  }

  double t = rec[R_TIME], freq = rec[R_FREQ], wind = rec[R_WIND], temp = rec[R_TEMP],
         cloud = rec[R_CLOUD], total_losses = rec[R_TOTAL_LOSSES];
  const uint64_t seed = ((const uint64_t*)rec)[R_SEED];
  const uint64_t draw = ((const uint64_t*)rec)[R_DRAWS];
  const double dt = cfg.dt;
  const double* nz = noise ? noise + env * lay.n_noise : nullptr;

  // ---- batteries (grid_env.py:629-641; dynamics.py:189-220, 304-324)
  double soc_reward = 0.0;
  for (int b = g.lane; b < Bt; b += LANES) {
    double rating = dimg[lay.o_bat_rating + b], cap = dimg[lay.o_bat_cap + b],
           eff = dimg[lay.o_bat_eff + b];
    double soc = rec[R_BAT + b], cur = rec[R_BAT + Bt + b];
    double cmd = (zero_action ? 0.0 : act[b]) * rating;
    if (cmd > 0.0) {
      double lim = cmd < rating ? cmd : rating;
      double en = lim * dt / 3600.0, room = soc * cap * eff;
      en = en < room ? en : room;
      cur = en * 3600.0 / dt;
      soc = soc - en / (cap * eff);
    } else if (cmd < 0.0) {
      double lim = -cmd < rating ? -cmd : rating;
      double en = lim * dt / 3600.0, room = ((1.0 - soc) * cap) / eff;
      en = en < room ? en : room;
      cur = -(en * 3600.0 / dt);
      soc = soc + en * eff / cap;
    }
    rec[R_BAT + b] = soc; rec[R_BAT + Bt + b] = cur;
    ob[o_bat + 2 * b] = soc; ob[o_bat + 2 * b + 1] = cur;
    g.scr(L + G + b) = cur;
    soc_reward += (soc >= 0.2 && soc <= 0.8) ? 1.0 : -5.0;
  }
  // ---- clock, weather (grid_env.py:470-471, 653-681)
  t += dt;
  step += 1;
  const double hour = hour_of(t);
  double z_load0 = 0.0;
  if (cfg.weather_variation || (cfg.stochastic_loads && !nz)) {
    double u, z1, z2, z3, dummy;
    if (nz) { u = nz[0]; z1 = nz[1]; z2 = nz[2]; z3 = nz[3]; }
    else {
      noise_block(seed, draw, 0u, &u, &dummy);
      noise_block(seed, draw, 1u, &z1, &z2);
      noise_block(seed, draw, 2u, &z3, &z_load0);
    }
    if (cfg.weather_variation) update_weather(hour, u, z1, z2, z3, &wind, &temp, &cloud);
    if (o.noise_used && g.lane == 0) {
      double* nu = o.noise_used + env * lay.n_noise;
      nu[0] = u; nu[1] = z1; nu[2] = z2; nu[3] = z3;
    }
  } else if (o.noise_used && g.lane == 0) {
    double* nu = o.noise_used + env * lay.n_noise;
    nu[0] = nu[1] = nu[2] = nu[3] = 0.0;
  }
  // ---- renewables (dynamics.py:120-170) and curtailment (grid_env.py:643-651)
  double tot_ren = 0.0, tot_used = 0.0;
  for (int gi = g.lane; gi < G; gi += LANES) {
    double p = renewable_power(lay, simg, dimg, gi, hour, wind, temp, cloud);
    double curtail = ((zero_action ? 0.0 : act[Bt + gi]) + 1.0) / 2.0;
    ob[o_gen + gi] = p;
    g.scr(L + gi) = p * curtail;
    tot_ren += p;
    tot_used += p - p * (1.0 - curtail);
  }
  // ---- loads (dynamics.py:54-75)
  if (cfg.stochastic_loads) {
    int hi = (int)hour;
    int nx = (hi + 1) % 24;
    double frac = hour - (double)hi;
    double mult = dimg[lay.o_profile + hi] * (1.0 - frac) + dimg[lay.o_profile + nx] * frac;
    if (nz) {
      for (int l = g.lane; l < L; l += LANES) {
        double z = nz[4 + l];
        double pl = dimg[lay.o_load_base + l] * (mult * (1.0 + (0.0 + cfg.load_noise * z))) * 1.0;
        g.scr(l) = pl > 0.0 ? pl : 0.0;
        if (o.noise_used) o.noise_used[env * lay.n_noise + 4 + l] = z;
      }
    } else {
      // slot 4 + l is normal j = 3 + l: block 1 + j / 2, component j & 1
      if (g.lane == 0 && L > 0) {
        double pl = dimg[lay.o_load_base] * (mult * (1.0 + (0.0 + cfg.load_noise * z_load0))) * 1.0;
        g.scr(0) = pl > 0.0 ? pl : 0.0;
        if (o.noise_used) o.noise_used[env * lay.n_noise + 4] = z_load0;
      }
      for (int q = 3 + g.lane; 2 * q - 5 < L; q += LANES) {
        double za, zb;
        noise_block(seed, draw, (uint32_t)q, &za, &zb);
        int l = 2 * q - 5;
        double pl = dimg[lay.o_load_base + l] * (mult * (1.0 + (0.0 + cfg.load_noise * za))) * 1.0;
        g.scr(l) = pl > 0.0 ? pl : 0.0;
        if (o.noise_used) o.noise_used[env * lay.n_noise + 4 + l] = za;
        if (l + 1 < L) {
          pl = dimg[lay.o_load_base + l + 1] * (mult * (1.0 + (0.0 + cfg.load_noise * zb))) * 1.0;
          g.scr(l + 1) = pl > 0.0 ? pl : 0.0;
          if (o.noise_used) o.noise_used[env * lay.n_noise + 4 + l + 1] = zb;
        }
      }
    }
  } else {
    for (int l = g.lane; l < L; l += LANES) {
      g.scr(l) = dimg[lay.o_load_base + l];
      if (o.noise_used) o.noise_used[env * lay.n_noise + 4 + l] = 0.0;
    }
  }
  g.sync();
  // ---- injections per bus (grid_env.py:683-720 with D3; power_flow.py:105-121 with D1)
  {
    const int* inj_ptr = simg + lay.o_inj_ptr;
    const int* inj_idx = simg + lay.o_inj_idx;
    for (int k = g.lane; k < n; k += LANES) {
      double ld = 0.0, gn = 0.0;
      for (int q = inj_ptr[k]; q < inj_ptr[k + 1]; ++q) {
        int j = inj_idx[q];
        double v = g.scr(j);
        if (j < L) ld += v;
        else if (j < L + G) gn += v;
        else if (v > 0.0) gn += v;
        else if (v < 0.0) ld += fabs(v);
      }
      // write after every lane has read its sources: F_P is not part of the scratch region
      g.at(F_P, k) = (0.0 - ld / lay.s_base) + gn / lay.s_base;
    }
  }
  g.sync();
  flat_start(g, lay, simg, dimg, nf);
  SolveStat st;
  if (SOLVER == SOLVER_NEWTON) newton_solve(g, lay, simg, dimg, cfg.tol, cfg.max_it, cfg.accel, &st);
  else sweep_solve(g, lay, simg, dimg, cfg.tol, cfg.max_it, &st);

  // ---- bus state -> observation (grid_env.py:722-731, 753-765), reductions for reward / constraints
  double dev = 0.0, vmax = -INFINITY, vmin = INFINITY;
  int v_hi = 0, v_lo = 0;
  {
    const int* rank = simg + lay.o_rank;
    for (int i = g.lane; i < n; i += LANES) {
      int k = rank[i];
      double e = g.at(F_E, k), f = g.at(F_F, k);
      double vm = sqrt(e * e + f * f);
      ob[2 * i] = vm;
      ob[2 * i + 1] = atan2(f, e);
      dev += fabs(vm - 1.0);
      vmax = (vm > vmax || vm != vm) ? vm : vmax;
      vmin = (vm < vmin || vm != vm) ? vm : vmin;
      v_hi |= vm > cfg.v_max;
      v_lo |= vm < cfg.v_min;
    }
  }
  double loss_pu = 0.0;
  int over80 = 0;
  {
    const int* bol = simg + lay.o_branch_of_line;
    for (int li = g.lane; li < m; li += LANES) {
      int k = bol[li];
      double P, S, ls;
      branch_flow(g, lay, simg, dimg, k, &P, &S, &ls);
      loss_pu += ls;
      double pw = P * lay.s_base;
      double rating = dimg[lay.o_rating + k];
      double loading = rating > 0.0 ? fabs(pw) / rating : 0.0;     // Line.update_state, base.py:261-264
      ob[o_line + 2 * li] = pw;
      ob[o_line + 2 * li + 1] = loading;
      over80 += loading > 0.8;
    }
  }
  dev = g.gsum(dev);
  loss_pu = g.gsum(loss_pu);
  over80 = g.gsum(over80);
  vmax = g.gmax_nan(vmax);
  {
    // NaN-propagating min
    double v = vmin;
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) {
      double w = __shfl_xor_sync(g.mask, v, off);
      v = (w < v || w != w) ? w : v;
    }
#endif
    vmin = v;
  }
  v_hi = g.gor(v_hi); v_lo = g.gor(v_lo);
  tot_ren = g.gsum(tot_ren); tot_used = g.gsum(tot_used); soc_reward = g.gsum(soc_reward);

  // ---- losses, frequency (grid_env.py:733-751; dynamics.py:260-273)
  const double losses_w = loss_pu * lay.s_base;
  total_losses += losses_w * dt / 3600.0;
  {
    double imb = (tot_ren - lay.load_p_sum - losses_w) / 1e6;
    double df = (imb - 1.0 * (freq - 60.0)) / (2.0 * 5.0 * 60.0);
    freq = clampd(freq + df * dt, 55.0, 65.0);
  }
  // ---- reward (grid_env.py:785-826)
  double reward = 0.0;
  reward -= dev * 10.0;
  reward -= fabs(freq - 60.0) * 20.0;
  reward -= (double)over80 * 50.0;
  reward -= total_losses * 0.1;
  reward += tot_used * 1e-5;
  reward += soc_reward;
  // ---- done / constraints (base.py:140-167; grid_env.py:563-608)
  const int terminated = step >= cfg.episode_length;
  const int f_hi = freq > cfg.f_max, f_lo = freq < cfg.f_min;
  const int anyv = v_hi | v_lo | f_hi | f_lo;
  viol_count += anyv;
  const int truncated = anyv && viol_count > 10;
  if (truncated) reward -= cfg.penalty;
  episode_reward += reward;

  if (g.lane == 0) {
    ob[o_freq] = freq;
    rec[R_TIME] = t; rec[R_FREQ] = freq; rec[R_WIND] = wind; rec[R_TEMP] = temp; rec[R_CLOUD] = cloud;
    rec[R_TOTAL_LOSSES] = total_losses; rec[R_EPISODE_REWARD] = episode_reward;
    ((uint64_t*)rec)[R_DRAWS] = draw + 1ull;
    ((int32_t*)(rec + R_COUNTS))[0] = step;
    ((int32_t*)(rec + R_COUNTS))[1] = viol_count;
    if (o.reward) o.reward[env] = reward;
    if (o.terminated) o.terminated[env] = (uint8_t)terminated;
    if (o.truncated) o.truncated[env] = (uint8_t)truncated;
    if (o.error) o.error[env] = 0;
    if (o.converged) o.converged[env] = (uint8_t)st.converged;
    if (o.iterations) o.iterations[env] = st.iterations;
    if (o.max_voltage) o.max_voltage[env] = vmax;
    if (o.min_voltage) o.min_voltage[env] = vmin;
    if (o.losses) o.losses[env] = losses_w;
    if (o.max_mismatch) o.max_mismatch[env] = st.max_mismatch;
    if (o.violations) {
      o.violations[env * 4 + 0] = (uint8_t)v_hi; o.violations[env * 4 + 1] = (uint8_t)v_lo;
      o.violations[env * 4 + 2] = (uint8_t)f_hi; o.violations[env * 4 + 3] = (uint8_t)f_lo;
    }
    if (o.violation_count) o.violation_count[env] = viol_count;
    if (o.current_step) o.current_step[env] = step;
    if (o.episode_reward) o.episode_reward[env] = episode_reward;
  }
  g.sync();   // the working set is reused by the group's next instance
}

// GridEnvironment.reset (grid_env.py:360-408): counters to zero, buses at 1.0 / 0, lines idle,
// 60 Hz, batteries at their initial state of charge, one weather update at t = 0, full observation.
template <int LANES>
GFR_HD void reset_instance(const Grp<LANES>& g, const Layout& lay, const int* simg,
                           const double* dimg, const EnvCfg& cfg, long long env, double* state,
                           double* obs, const double* load_pq, const double* bat_soc0,
                           const uint64_t* seeds, const double* noise, double start_time,
                           bool construct) {
  const int n = lay.n, m = lay.m, L = lay.L, G = lay.G, Bt = lay.Bt, D = lay.D;
  double* rec = state + env * lay.R;
  double* ob = obs + env * D;
  const int o_line = 2 * n, o_freq = 2 * n + 2 * m, o_load = o_freq + 1, o_gen = o_load + 2 * L,
            o_bat = o_gen + G;
  uint64_t seed = seeds ? seeds[env] : ((const uint64_t*)rec)[R_SEED];
  uint64_t draw = (seeds || construct) ? 0ull : ((const uint64_t*)rec)[R_DRAWS];
  // wind / temperature / cloud survive a reset (grid_env.py:673-681); construction sets 5 / 25 / 0.3
  double wind = construct ? 5.0 : rec[R_WIND], temp = construct ? 25.0 : rec[R_TEMP],
         cloud = construct ? 0.3 : rec[R_CLOUD];
  if (construct && !seeds) seed = (uint64_t)env;
  const bool draws = cfg.weather_variation && !construct;
  if (draws) {
    double u, z1, z2, z3, dummy;
    if (noise) { const double* nz = noise + env * 4; u = nz[0]; z1 = nz[1]; z2 = nz[2]; z3 = nz[3]; }
    else {
      noise_block(seed, draw, 0u, &u, &dummy);
      noise_block(seed, draw, 1u, &z1, &z2);
      noise_block(seed, draw, 2u, &z3, &dummy);
    }
    update_weather(hour_of(0.0), u, z1, z2, z3, &wind, &temp, &cloud);
  }
  for (int i = g.lane; i < n; i += LANES) { ob[2 * i] = 1.0; ob[2 * i + 1] = 0.0; }
  for (int li = g.lane; li < m; li += LANES) { ob[o_line + 2 * li] = 0.0; ob[o_line + 2 * li + 1] = 0.0; }
  for (int l = g.lane; l < 2 * L; l += LANES) ob[o_load + l] = load_pq[l];
  for (int gi = g.lane; gi < G; gi += LANES)
    ob[o_gen + gi] = renewable_power(lay, simg, dimg, gi, hour_of(0.0), wind, temp, cloud);
  for (int b = g.lane; b < Bt; b += LANES) {
    rec[R_BAT + b] = bat_soc0[b]; rec[R_BAT + Bt + b] = 0.0;
    ob[o_bat + 2 * b] = bat_soc0[b]; ob[o_bat + 2 * b + 1] = 0.0;
  }
  if (g.lane == 0) {
    ob[o_freq] = 60.0;
    rec[R_TIME] = start_time; rec[R_FREQ] = 60.0; rec[R_WIND] = wind; rec[R_TEMP] = temp;
    rec[R_CLOUD] = cloud; rec[R_TOTAL_LOSSES] = 0.0; rec[R_EPISODE_REWARD] = 0.0;
    ((uint64_t*)rec)[R_SEED] = seed;
    ((uint64_t*)rec)[R_DRAWS] = draws ? draw + 1ull : draw;
    ((int32_t*)(rec + R_COUNTS))[0] = 0;
    ((int32_t*)(rec + R_COUNTS))[1] = 0;
  }
}

}  // namespace gfr
