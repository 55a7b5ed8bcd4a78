// gpc_traj.cuh -- candidate generation on the device: primitive chains -> way-points -> trajectory
// points -> fidelity-labelled candidate rows, one CTA per candidate path.
//
// replaces the host loops that run right before the information-gain operators
//   GraceRIGV3.py:235-294  evaluateTraj              (sequential scan over <= numLegs primitives)
//   GraceRIGV3.py:373-394  edgePointsToTrajPoints    (np.arange / np.interp resampling at measRate)
//   GraceRIGV3.py:396-427  pathToTrajPoints          (concatenate edges, np.unique(round(., 4)) keeping order)
//   GraceRIGV3.py:508-512 / 529-533                  (fidelity label from the localisation variance)
// Quirks are kept: t_off advances by the LAST ENTRY of the last way-point (the variance when
// withVar, GraceRIGV3.py:423); the variance only grows while the swimmer is under water and is
// reset whenever depth <= 0.  Arithmetic mirrors NumPy's (no FMA contraction where NumPy rounds twice).
#pragma once
#include "gpc_common.cuh"

#define GPC_TRAJ_MAXRAW 512   // raw points per path before de-duplication
#define GPC_TRAJ_MAXEDGE 32   // edges per path

// edge e: way-points wp[(prim_off[e] + e) .. +n_prims], 4 doubles each: distance, depth, time, variance
__device__ __forceinline__ void traj_scan_edge(const double* __restrict__ prims, long p0, long p1, double vrate,
                                               int with_var, double* __restrict__ wp) {
  double t = 0.0, dist = 0.0, var = 0.0, depth = 0.0;
  bool uw = false, restart = false;
  wp[0] = dist; wp[1] = depth; wp[2] = t; wp[3] = var;
  for (long p = p0; p < p1; ++p) {
    const double* pr = prims + p * 4;
    const int kind = (int)pr[0];
    if (kind == 0) {            // spiral (dz, _, speed)
      const double dt = fabs(__ddiv_rn(pr[1], pr[3]));
      t = __dadd_rn(t, dt);
      var = __dadd_rn(var, __dmul_rn(vrate, dt));
      depth = __dadd_rn(depth, pr[1]);
    } else if (kind == 1) {     // glide (gp, dz, speed)
      const double dt = fabs(__ddiv_rn(pr[2], pr[3]));
      t = __dadd_rn(t, dt);
      var = __dadd_rn(var, __dmul_rn(vrate, dt));
      dist = __dadd_rn(dist, __ddiv_rn(pr[2], tan(pr[1])));
      depth = __dadd_rn(depth, pr[2]);
    } else if (kind == 2) {     // swim (dist, speed)
      const double dt = __ddiv_rn(pr[1], pr[2]);
      t = __dadd_rn(t, dt);
      var = __dadd_rn(var, __dmul_rn(__dmul_rn(vrate, uw ? 1.0 : 0.0), dt));
      dist = __dadd_rn(dist, pr[1]);
    } else if (kind == 3) {     // flat dive (dz, speed)
      const double dt = fabs(__ddiv_rn(pr[1], pr[2]));
      t = __dadd_rn(t, dt);
      var = __dadd_rn(var, __dmul_rn(vrate, dt));
      depth = __dadd_rn(depth, pr[1]);
    }
    if (depth > 0.0) uw = restart = true;
    else if (depth <= 0.1 && restart) uw = restart = false;
    if (depth <= 0.0) var = 0.0;
    double* w = wp + (p - p0 + 1) * 4;
    w[0] = dist; w[1] = depth; w[2] = t; w[3] = with_var ? var : 0.0;
  }
}

// np.interp(x, xp, fp) for increasing xp (n >= 1)
__device__ __forceinline__ double traj_interp(double x, const double* __restrict__ wp, int n, double toff, int col) {
  const double x0 = __dadd_rn(wp[2], toff), xn = __dadd_rn(wp[(n - 1) * 4 + 2], toff);
  if (x < x0) return wp[col];                      // left of the table: fp[0]
  if (x >= xn) return wp[(n - 1) * 4 + col];       // right of / on the last knot: fp[-1]
  int j = 0;
  for (int q = 1; q < n - 1; ++q)
    if (__dadd_rn(wp[q * 4 + 2], toff) <= x) j = q;
  const double xj = __dadd_rn(wp[j * 4 + 2], toff), xj1 = __dadd_rn(wp[(j + 1) * 4 + 2], toff);
  const double fj = wp[j * 4 + col], fj1 = wp[(j + 1) * 4 + col];
  if (xj == x) return fj;
  const double slope = __ddiv_rn(__dsub_rn(fj1, fj), __dsub_rn(xj1, xj));
  return __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xj)), fj);
}

// grid = C paths, 128 threads.  pts [C][max_pts][5] (x, y, z, t, var), fid [C][max_pts], counts [C]
// (counts[c] = number of points the path produces; > max_pts means the output was truncated).
__global__ void __launch_bounds__(128) k_traj_points(const long* __restrict__ edge_off, const double* __restrict__ edge_xy,
                                                     const long* __restrict__ prim_off, const double* __restrict__ prims,
                                                     double* __restrict__ wpws, double vrate, double meas_rate, int dense,
                                                     int with_var, double t_off0, double fl0, double fl1, int have_fl,
                                                     int max_pts, double* __restrict__ pts, double* __restrict__ fid,
                                                     long* __restrict__ counts) {
  __shared__ double raw[GPC_TRAJ_MAXRAW][5];
  __shared__ double toff_e[GPC_TRAJ_MAXEDGE];
  __shared__ int start_e[GPC_TRAJ_MAXEDGE + 1];
  __shared__ unsigned char keep[GPC_TRAJ_MAXRAW];
  __shared__ int outpos[GPC_TRAJ_MAXRAW];
  __shared__ int total_s;
  const int c = blockIdx.x, tid = threadIdx.x;
  const long e0 = edge_off[c], e1 = edge_off[c + 1];
  const int ne = (int)(e1 - e0) < GPC_TRAJ_MAXEDGE ? (int)(e1 - e0) : GPC_TRAJ_MAXEDGE;
  // (A) way-points of every edge (one thread per edge)
  for (int e = tid; e < ne; e += 128)
    traj_scan_edge(prims, prim_off[e0 + e], prim_off[e0 + e + 1], vrate, with_var, wpws + (prim_off[e0 + e] + e0 + e) * 4);
  __syncthreads();
  // (B) time offsets and point counts per edge
  const double step = __ddiv_rn(1.0, meas_rate);
  if (tid == 0) {
    double toff = t_off0;
    int tot = 0;
    for (int e = 0; e < ne; ++e) {
      const long p0 = prim_off[e0 + e], np_ = prim_off[e0 + e + 1] - p0;
      const double* wp = wpws + (p0 + e0 + e) * 4;
      const double* last = wp + np_ * 4;
      toff_e[e] = toff;
      start_e[e] = tot;
      int n = (int)np_ + 1;
      if (dense) n = (int)ceil(__ddiv_rn(last[2], step));  // len(np.arange(0, T, step))
      if (n < 0) n = 0;
      tot += n;
      toff = __dadd_rn(toff, with_var ? last[3] : last[2]);  // `t_off += wpnts[-1][-1]`
    }
    start_e[ne] = tot;
    total_s = tot;
  }
  __syncthreads();
  const int total = total_s < GPC_TRAJ_MAXRAW ? total_s : GPC_TRAJ_MAXRAW;
  // (C) raw rows x, y, z, t, var
  for (int i = tid; i < total; i += 128) {
    int e = 0;
    while (e + 1 < ne && start_e[e + 1] <= i) ++e;
    const int li = i - start_e[e];
    const long p0 = prim_off[e0 + e];
    const int nwp = (int)(prim_off[e0 + e + 1] - p0) + 1;
    const double* wp = wpws + (p0 + e0 + e) * 4;
    const double* xy = edge_xy + (e0 + e) * 4;
    const double b = atan2(__dsub_rn(xy[3], xy[1]), __dsub_rn(xy[2], xy[0]));
    const double cb = cos(b), sb = sin(b);
    double d, z, t, v;
    if (dense) {
      t = __dadd_rn(__dmul_rn((double)li, step), toff_e[e]);
      d = traj_interp(t, wp, nwp, toff_e[e], 0);
      z = traj_interp(t, wp, nwp, toff_e[e], 1);
      v = with_var ? traj_interp(t, wp, nwp, toff_e[e], 3) : 0.0;
    } else {
      d = wp[li * 4]; z = wp[li * 4 + 1]; t = __dadd_rn(wp[li * 4 + 2], toff_e[e]); v = wp[li * 4 + 3];
    }
    raw[i][0] = __dadd_rn(xy[0], __dmul_rn(d, cb));
    raw[i][1] = __dadd_rn(xy[1], __dmul_rn(d, sb));
    raw[i][2] = z;
    raw[i][3] = t;
    raw[i][4] = v;
  }
  __syncthreads();
  // (D) np.unique(np.round(rows, 4), axis=0) keeping first occurrences in order
  for (int i = tid; i < total; i += 128) {
    double key[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) key[q] = rint(__dmul_rn(raw[i][q], 1e4));
    bool dup = false;
    for (int j = 0; j < i && !dup; ++j) {
      bool same = true;
#pragma unroll
      for (int q = 0; q < 5; ++q) same = same && (rint(__dmul_rn(raw[j][q], 1e4)) == key[q]);
      dup = same;
    }
    keep[i] = dup ? 0 : 1;
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int i = 0; i < total; ++i) { outpos[i] = n; n += keep[i]; }
    counts[c] = (total_s > GPC_TRAJ_MAXRAW) ? (long)total_s : (long)n;
    total_s = n;
  }
  __syncthreads();
  for (int i = tid; i < total; i += 128) {
    if (!keep[i] || outpos[i] >= max_pts) continue;
    double* o = pts + ((long)c * max_pts + outpos[i]) * 5;
#pragma unroll
    for (int q = 0; q < 5; ++q) o[q] = raw[i][q];
    if (fid) {
      const double v = raw[i][4];
      double f = 0.0;
      if (have_fl) f = (v < fl0) ? 2.0 : ((v > fl0 && v < fl1) ? 1.0 : 0.0);
      fid[(long)c * max_pts + outpos[i]] = f;
    }
  }
}
