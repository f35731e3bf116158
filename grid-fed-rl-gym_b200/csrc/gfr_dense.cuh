// gfr_dense.cuh - Newton-Raphson load flow on an ARBITRARY (meshed) network: one CTA per
// instance, the dense polar Jacobian assembled in shared memory and factorised there by Gaussian
// elimination with partial pivoting.  This is the reference's algorithm as written - dense
// (2(n-1))^2 system, np.linalg.solve = LAPACK dgesv - for the networks the tree-ordered kernels of
// gfr_device.cuh cannot take (cycles: SyntheticFeeder(connectivity > 0), the shipped IEEE feeders
// with their loop-closing lines kept).  Size limit: the Jacobian has to fit the shared memory of an
// SM, N = (#non-slack) + (#PQ) <= ~165.
//
// Reference (paths under /root/reference/grid_fed_rl/):
//   Ybus                  environments/power_flow.py:48-73   (|z| <= 1e-12 -> open line)
//   Newton-Raphson        environments/power_flow.py:89-211  (flat start; check, then update)
//   Jacobian              environments/power_flow.py:213-295 (+ deviation D2); unknowns
//                         ordered as there: angles of the non-slack buses, then |V| of the PQ buses
//   solve                 environments/power_flow.py:187     (np.linalg.solve -> dgesv: LU, partial pivoting)
//   update                environments/power_flow.py:297-327
//   line flows / losses   environments/power_flow.py:329-358, :199-200
#pragma once
#include "gfr_device.cuh"

namespace gfr {

// Device-resident network (bus / line order = the caller's, no renumbering)
struct NetDev {
  int n, m, N, n_theta;           // buses, lines, unknowns, angle unknowns (= non-slack buses)
  double s_base;
  const int* bus_type;            // [n]
  const int* col_theta;           // [n] column (= row) of the bus's angle unknown, -1 for the slack
  const int* col_vm;              // [n] column (= row) of the bus's |V| unknown, -1 unless PQ
  const int* adj_ptr;             // [n + 1] neighbours of a bus, parallel lines merged
  const int* adj_idx;             // [nnz]
  const D2* adj_y;                // [nnz] (G_ij, B_ij) = -(sum of the series admittances between i and j)
  const D2* ydiag;                // [n]   (G_ii, B_ii)
  const double* vm_set;           // [n]
  const int* line_from;           // [m]
  const int* line_to;             // [m]
  const D2* line_y;               // [m] series g + jb (0 for an open line)
  const double* line_rating;      // [m]
};

GFR_HD size_t dense_smem_bytes(int n, int N) {
  const size_t ld = (size_t)(N | 1);                        // odd leading dimension: conflict-free column walks
  return ld * (size_t)N * 8 + (size_t)N * 8 + (size_t)n * 32 + 64 * 8;
}

#if defined(__CUDACC__)

// NaN-propagating maximum over the CTA (numpy's max); every thread gets the result
__device__ __forceinline__ double block_max_nan(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = (w > v || w != w) ? w : v;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { const double x = red[w]; v = (x > v || x != x) ? x : v; }
  return v;
}

__global__ void __launch_bounds__(256)
dense_solve_kernel(const NetDev net, const double tol, const int max_it, const double accel,
                   const double* __restrict__ p_inj, const SolOut o, const long long B) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int n = net.n, m = net.m, N = net.N;
  const int ld = N | 1;
  double* J = reinterpret_cast<double*>(smem);               // column-major, J[r + c * ld]
  double* rhs = J + (size_t)ld * N;                          // [N] mismatch, then the corrections
  D2* ef = reinterpret_cast<D2*>(rhs + N);                   // [n] e + jf   (16-byte aligned: ld is odd, so (ld + 1) N doubles is even)
  D2* pq = ef + n;                                           // [n] calculated (P, Q)
  double* red = reinterpret_cast<double*>(pq + n);           // [64] reductions, pivot search
  int* redi = reinterpret_cast<int*>(red + 32);
  const int tid = threadIdx.x, nt = blockDim.x;

  for (long long env = blockIdx.x; env < B; env += gridDim.x) {
    const double* pspec = p_inj + env * n;
    // flat start (:103, :131): 1.0 at 0 rad, slack / PV buses at their set magnitude
    for (int i = tid; i < n; i += nt) {
      D2 v;
      v.x = net.bus_type[i] == BUS_PQ ? 1.0 : net.vm_set[i];
      v.y = 0.0;
      ef[i] = v;
    }
    __syncthreads();
    int converged = 0, iterations = max_it;
    double max_mismatch = INFINITY;
    for (int it = 0; it < max_it; ++it) {
      // ---- calculated injections and mismatch (:150-166)
      double mm = 0.0;
      for (int i = tid; i < n; i += nt) {
        const D2 vi = ef[i];
        const D2 yd = net.ydiag[i];
        const double v2 = fma(vi.x, vi.x, vi.y * vi.y);
        double P = yd.x * v2, Q = -yd.y * v2;
        for (int q = net.adj_ptr[i]; q < net.adj_ptr[i + 1]; ++q) {
          const D2 vj = ef[net.adj_idx[q]];
          const D2 y = net.adj_y[q];
          const double a = fma(vi.x, vj.x, vi.y * vj.y), s = fma(vi.y, vj.x, -vi.x * vj.y);
          P = fma(y.x, a, fma(y.y, s, P));
          Q = fma(y.x, s, fma(-y.y, a, Q));
        }
        D2 c; c.x = P; c.y = Q;
        pq[i] = c;
        const int rt = net.col_theta[i], rv = net.col_vm[i];
        double aP = 0.0, aQ = 0.0;
        if (rt >= 0) { const double d = pspec[i] - P; rhs[rt] = d; aP = fabs(d); }
        if (rv >= 0) { const double d = 0.0 - Q; rhs[rv] = d; aQ = fabs(d); }
        const double loc = (aQ > aP || aQ != aQ) ? aQ : aP;
        mm = (loc > mm || loc != loc) ? loc : mm;
      }
      mm = block_max_nan(mm, red);
      max_mismatch = mm;
      if (mm < tol) { converged = 1; iterations = it + 1; break; }      // checked before the update (:168-171)
      // ---- Jacobian (:213-295 with D2), |V| columns scaled by |V| (the update below undoes it)
      for (int q = tid; q < ld * N; q += nt) J[q] = 0.0;
      __syncthreads();
      for (int i = tid; i < n; i += nt) {
        const int rt = net.col_theta[i], rv = net.col_vm[i];
        if (rt < 0) continue;
        const D2 vi = ef[i];
        const D2 yd = net.ydiag[i];
        const D2 c = pq[i];
        const double v2 = fma(vi.x, vi.x, vi.y * vi.y);
        J[rt + rt * ld] = fma(-yd.y, v2, -c.y);                        // dP/dtheta:    -Q - B v2
        if (rv >= 0) {
          J[rt + rv * ld] = fma(yd.x, v2, c.x);                        // V dP/dV:       P + G v2
          J[rv + rt * ld] = fma(-yd.x, v2, c.x);                       // dQ/dtheta:     P - G v2
          J[rv + rv * ld] = fma(-yd.y, v2, c.y);                       // V dQ/dV:       Q - B v2
        }
        for (int q = net.adj_ptr[i]; q < net.adj_ptr[i + 1]; ++q) {
          const int j = net.adj_idx[q];
          const int ct = net.col_theta[j], cv = net.col_vm[j];
          if (ct < 0) continue;                                          // the slack has no unknowns
          const D2 vj = ef[j];
          const D2 y = net.adj_y[q];
          const double a = fma(vi.x, vj.x, vi.y * vj.y), s = fma(vi.y, vj.x, -vi.x * vj.y);
          const double al = fma(y.x, s, -y.y * a);                       // |Vi||Vj| (G sin - B cos)
          const double ga = fma(y.x, a, y.y * s);                        // |Vi||Vj| (G cos + B sin)
          J[rt + ct * ld] = al;
          if (cv >= 0) J[rt + cv * ld] = ga;
          if (rv >= 0) {
            J[rv + ct * ld] = -ga;
            if (cv >= 0) J[rv + cv * ld] = al;
          }
        }
      }
      __syncthreads();
      // ---- Gaussian elimination with partial pivoting on [J | rhs] (dgesv, :187).  Two CTA barriers per
      //      pivot column: the multipliers are formed on the fly (nothing reuses L: the right-hand side is
      //      eliminated in the same sweep), and the warp that updates column k + 1 finds that column's
      //      pivot - largest |entry|, the first among equals (idamax) - before the barrier that ends step k.
      int singular = 0;
      {
        // pivot of column 0 by warp 0
        if (tid < 32) {
          double best = -1.0;
          int bi = 0;
          for (int i = tid; i < N; i += 32) {
            const double a = fabs(J[i]);
            if ((a > best || a != a) && !(best != best)) { best = a; bi = i; }
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            const bool take = (ob != ob && !(best != best)) || (!(best != best) && (ob > best || (ob == best && oi < bi)));
            if (take) { best = ob; bi = oi; }
          }
          if (tid == 0) { red[0] = best; redi[0] = bi; }
        }
        __syncthreads();
      }
      for (int k = 0; k < N; ++k) {
        const double best = red[0];
        const int bi = redi[0];
        if (!(best > 0.0) && !(best != best)) { singular = 1; break; }   // exact-zero pivot (:188-190); NaN runs on
        __syncthreads();                                                 // red is rewritten at the end of this step
        if (bi != k) {
          for (int c = k + tid; c <= N; c += nt) {
            double* pa = c < N ? &J[k + c * ld] : &rhs[k];
            double* pb = c < N ? &J[bi + c * ld] : &rhs[bi];
            const double t = *pa; *pa = *pb; *pb = t;
          }
          __syncthreads();
        }
        const double rp = 1.0 / J[k + k * ld];
        {
          const int tx = tid & 31, ty = tid >> 5, ny = nt >> 5;
          for (int c = k + 1 + ty; c <= N; c += ny) {
            double* col = c < N ? &J[c * ld] : rhs;
            const double u = col[k];
            double nbest = -1.0;
            int nbi = k + 1;
            for (int i = k + 1 + tx; i < N; i += 32) {
              const double v = fma(-(J[i + k * ld] * rp), u, col[i]);    // multiplier = entry x (1 / pivot), as dgetf2
              col[i] = v;
              const double a = fabs(v);
              if ((a > nbest || a != a) && !(nbest != nbest)) { nbest = a; nbi = i; }
            }
            if (c == k + 1 && c < N) {                                   // warp 0, first column of its sweep: the next pivot
#pragma unroll
              for (int off = 16; off > 0; off >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, nbest, off);
                const int oi = __shfl_xor_sync(0xffffffffu, nbi, off);
                const bool take = (ob != ob && !(nbest != nbest)) ||
                                  (!(nbest != nbest) && (ob > nbest || (ob == nbest && oi < nbi)));
                if (take) { nbest = ob; nbi = oi; }
              }
              if (tx == 0) { red[0] = nbest; redi[0] = nbi; }
            }
          }
        }
        __syncthreads();
      }
      if (singular) { iterations = it + 1; break; }
      // ---- back substitution by warp 0 alone (warp barriers instead of 2 N CTA barriers):
      //      x_k = rhs_k / U_kk, then rhs_i -= U_ik x_k above it
      if (tid < 32) {
        for (int k = N - 1; k >= 0; --k) {
          const double xk = rhs[k] / J[k + k * ld];
          __syncwarp();
          if (tid == 0) rhs[k] = xk;
          for (int i = tid; i < k; i += 32) rhs[i] = fma(-J[i + k * ld], xk, rhs[i]);
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- polar update (:297-327): theta += a dtheta, |V| += a d|V|  <=>  V *= (1 + a x_v) e^{j a x_theta}
      for (int i = tid; i < n; i += nt) {
        const int ct = net.col_theta[i], cv = net.col_vm[i];
        if (ct < 0) continue;
        double sn, cs;
        sincos_small(accel * rhs[ct], &sn, &cs);
        const double sc = cv >= 0 ? fma(accel, rhs[cv], 1.0) : 1.0;
        const D2 v = ef[i];
        D2 w;
        w.x = sc * fma(v.x, cs, -v.y * sn);
        w.y = sc * fma(v.x, sn, v.y * cs);
        ef[i] = w;
      }
      __syncthreads();
    }
    // ---- results (ref order = the caller's order)
    for (int i = tid; i < n; i += nt) {
      const D2 v = ef[i];
      if (o.bus_voltages) o.bus_voltages[env * n + i] = sqrt(v.x * v.x + v.y * v.y);
      if (o.bus_angles) o.bus_angles[env * n + i] = atan2_bus(v.y, v.x);
    }
    double loss = 0.0;
    for (int li = tid; li < m; li += nt) {
      const D2 vf = ef[net.line_from[li]], vt = ef[net.line_to[li]];
      const D2 y = net.line_y[li];
      const double de = vf.x - vt.x, df = vf.y - vt.y;
      const double ir = y.x * de - y.y * df, ii = y.x * df + y.y * de;      // I = y (V_from - V_to)
      const double P = vf.x * ir + vf.y * ii, Q = vf.y * ir - vf.x * ii;    // V_from conj(I)
      loss += y.x * (de * de + df * df);                                    // Re sum_i V_i conj((YV)_i), line by line
      const double rating = net.line_rating[li];
      if (o.line_flows) o.line_flows[env * m + li] = P;
      if (o.line_loadings) o.line_loadings[env * m + li] = rating > 0.0 ? sqrt(P * P + Q * Q) * net.s_base / rating : 0.0;
    }
    {   // deterministic sum over the CTA
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, off);
      __syncthreads();
      if ((tid & 31) == 0) red[tid >> 5] = loss;
      __syncthreads();
      loss = 0.0;
      for (int w = 0; w < (nt >> 5); ++w) loss += red[w];
    }
    if (tid == 0) {
      if (o.losses) o.losses[env] = loss;
      if (o.max_mismatch) o.max_mismatch[env] = max_mismatch;
      if (o.converged) o.converged[env] = (uint8_t)converged;
      if (o.iterations) o.iterations[env] = iterations;
    }
    __syncthreads();
  }
}

#endif  // __CUDACC__

}  // namespace gfr
