// Per-point fp64 math of the map-consistency path, usable from device code and (for the
// host-side unit tests in tests/test_host_math.py) from plain C++.  No reference code is
// copied: the reference calls LAPACK (depth_cloud.py:386) and pytorch3d (transform.py:73).
#pragma once
#include "dc_common.cuh"

// ---------------------------------------------------------------------------------------
// Symmetric 3x3 eigen-decomposition, closed form (trigonometric) + Rayleigh polish.
// Replaces torch.linalg.eigh on [N,3,3] (depth_cloud.py:376-399).  Eigenvalues ascending.
// ---------------------------------------------------------------------------------------
struct dc_sym3 {
  double xx, xy, xz, yy, yz, zz;
};

DC_HD void dc_cross(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

DC_HD double dc_dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Unit eigenvector of the (scaled) matrix m for eigenvalue lam: the best conditioned cross
// product of two rows of (m - lam I).  Falls back to an arbitrary unit vector orthogonal to the
// dominant row when (m - lam I) has rank <= 1 (repeated eigenvalue).
DC_HD void dc_sym3_eigvec(const dc_sym3& m, double lam, double v[3]) {
  const double r0[3] = {m.xx - lam, m.xy, m.xz};
  const double r1[3] = {m.xy, m.yy - lam, m.yz};
  const double r2[3] = {m.xz, m.yz, m.zz - lam};
  double c0[3], c1[3], c2[3];
  dc_cross(r0, r1, c0);
  dc_cross(r1, r2, c1);
  dc_cross(r2, r0, c2);
  const double n0 = dc_dot3(c0, c0), n1 = dc_dot3(c1, c1), n2 = dc_dot3(c2, c2);
  const double* c = c0;
  double n = n0;
  if (n1 > n) { c = c1; n = n1; }
  if (n2 > n) { c = c2; n = n2; }
  const double s0 = dc_dot3(r0, r0), s1 = dc_dot3(r1, r1), s2 = dc_dot3(r2, r2);
  double smax = s0 > s1 ? s0 : s1;
  smax = smax > s2 ? smax : s2;
  if (n > 1e-28 * smax * smax && n > 0.0) {
    const double inv = 1.0 / sqrt(n);
    v[0] = c[0] * inv; v[1] = c[1] * inv; v[2] = c[2] * inv;
    return;
  }
  // rank <= 1: pick u orthogonal to the dominant row
  const double* r = r0;
  if (s1 >= s0 && s1 >= s2) r = r1;
  if (s2 >= s0 && s2 >= s1) r = r2;
  if (smax <= 0.0) { v[0] = 1.0; v[1] = 0.0; v[2] = 0.0; return; }
  // e = axis of the smallest |component| of r; v = normalize(r x e)
  const double ax = fabs(r[0]), ay = fabs(r[1]), az = fabs(r[2]);
  double e[3] = {0.0, 0.0, 0.0};
  if (ax <= ay && ax <= az) e[0] = 1.0; else if (ay <= az) e[1] = 1.0; else e[2] = 1.0;
  double u[3];
  dc_cross(r, e, u);
  const double inv = 1.0 / sqrt(dc_dot3(u, u));
  v[0] = u[0] * inv; v[1] = u[1] * inv; v[2] = u[2] * inv;
}

DC_HD double dc_sym3_rayleigh(const dc_sym3& m, const double v[3]) {
  const double ax = m.xx * v[0] + m.xy * v[1] + m.xz * v[2];
  const double ay = m.xy * v[0] + m.yy * v[1] + m.yz * v[2];
  const double az = m.xz * v[0] + m.yz * v[1] + m.zz * v[2];
  return v[0] * ax + v[1] * ay + v[2] * az;
}

// Orthonormal pair (u, v) spanning the plane orthogonal to the unit vector w, with u x v = w.
DC_HD void dc_complement(const double w[3], double u[3], double v[3]) {
  if (fabs(w[0]) > fabs(w[1])) {
    const double inv = 1.0 / sqrt(w[0] * w[0] + w[2] * w[2]);
    u[0] = -w[2] * inv; u[1] = 0.0; u[2] = w[0] * inv;
  } else {
    const double inv = 1.0 / sqrt(w[1] * w[1] + w[2] * w[2]);
    u[0] = 0.0; u[1] = w[2] * inv; u[2] = -w[1] * inv;
  }
  dc_cross(w, u, v);
}

DC_HD void dc_sym3_apply(const dc_sym3& m, const double x[3], double y[3]) {
  y[0] = m.xx * x[0] + m.xy * x[1] + m.xz * x[2];
  y[1] = m.xy * x[0] + m.yy * x[1] + m.yz * x[2];
  y[2] = m.xz * x[0] + m.yz * x[1] + m.zz * x[2];
}

// The two remaining eigenpairs of m in the plane orthogonal to the unit eigenvector w: the 2x2
// problem [[u'mu, u'mv], [u'mv, v'mv]] is solved in closed form, which stays accurate when the two
// eigenvalues are close to each other or tiny compared to the one already found.
// Outputs: lo <= hi and the unit eigenvector of lo (elo); the eigenvector of hi is w x elo / elo x w.
DC_HD void dc_sym3_deflate(const dc_sym3& m, const double w[3], double& lo, double& hi, double elo[3]) {
  double u[3], v[3], mu[3], mv[3];
  dc_complement(w, u, v);
  dc_sym3_apply(m, u, mu);
  dc_sym3_apply(m, v, mv);
  const double a = dc_dot3(u, mu), b = dc_dot3(u, mv), c = dc_dot3(v, mv);
  const double hd = 0.5 * (a - c), mid = 0.5 * (a + c);
  const double rad = sqrt(hd * hd + b * b);
  lo = mid - rad;
  hi = mid + rad;
  // eigenvector of lo in (u, v) coordinates: (b, lo - a) or (lo - c, b), whichever is longer
  double x0 = b, y0 = lo - a, x1 = lo - c, y1 = b;
  if (x1 * x1 + y1 * y1 > x0 * x0 + y0 * y0) { x0 = x1; y0 = y1; }
  const double n2 = x0 * x0 + y0 * y0;
  if (n2 > 0.0) {
    const double inv = 1.0 / sqrt(n2);
    x0 *= inv; y0 *= inv;
  } else {
    x0 = 1.0; y0 = 0.0;   // a == c, b == 0: every direction of the plane is an eigenvector
  }
  elo[0] = x0 * u[0] + y0 * v[0]; elo[1] = x0 * u[1] + y0 * v[1]; elo[2] = x0 * u[2] + y0 * v[2];
}

// lam[3] ascending.  want_vecs: 0 = eigenvalues only, 1 = V[0..2] = eigenvector of lam[0],
// 3 = V[3*j + i] = component i of eigenvector j (orthonormal, right-handed).
// Closed form in three steps: trigonometric root for the best separated eigenvalue (largest if
// det(B) >= 0, else smallest), its eigenvector from cross products + Rayleigh quotient, then the
// deflated 2x2 problem for the other pair.  Eigenvalue errors are O(eps * |A|) like LAPACK's.
// Returns false when the input is not finite (outputs are NaN, like LAPACK would propagate).
DC_HD bool dc_sym3_eig(const dc_sym3& a, double lam[3], double* V, int want_vecs) {
  double s = fabs(a.xx);
  s = fmax(s, fabs(a.xy)); s = fmax(s, fabs(a.xz));
  s = fmax(s, fabs(a.yy)); s = fmax(s, fabs(a.yz)); s = fmax(s, fabs(a.zz));
  if (!(s < INFINITY)) {   // inf or NaN
    const double nan = NAN;
    lam[0] = lam[1] = lam[2] = nan;
    for (int i = 0; i < 3 * want_vecs; ++i) V[i] = nan;
    return false;
  }
  const double inv = s > 0.0 ? 1.0 / s : 0.0;
  const dc_sym3 m = {a.xx * inv, a.xy * inv, a.xz * inv, a.yy * inv, a.yz * inv, a.zz * inv};
  const double q = (m.xx + m.yy + m.zz) * (1.0 / 3.0);
  const double bxx = m.xx - q, byy = m.yy - q, bzz = m.zz - q;
  const double p1 = m.xy * m.xy + m.xz * m.xz + m.yz * m.yz;
  const double p2 = bxx * bxx + byy * byy + bzz * bzz + 2.0 * p1;
  if (p2 <= 0.0) {         // multiple of the identity (including the zero matrix)
    lam[0] = lam[1] = lam[2] = q * s;
    if (want_vecs >= 1) { V[0] = 1; V[1] = 0; V[2] = 0; }
    if (want_vecs >= 3) { V[3] = 0; V[4] = 1; V[5] = 0; V[6] = 0; V[7] = 0; V[8] = 1; }
    return true;
  }
  const double p = sqrt(p2 * (1.0 / 6.0));
  const double ip = 1.0 / p;
  const double cxx = bxx * ip, cyy = byy * ip, czz = bzz * ip, cxy = m.xy * ip, cxz = m.xz * ip, cyz = m.yz * ip;
  const double det = cxx * (cyy * czz - cyz * cyz) - cxy * (cxy * czz - cyz * cxz) + cxz * (cxy * cyz - cyy * cxz);
  double r = 0.5 * det;
  r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
  const double phi = acos(r) * (1.0 / 3.0);
  double l0, l1, l2, w[3], e[3];
  if (r >= 0.0) {
    // the largest eigenvalue is the isolated one (linear / generic neighbourhoods)
    l2 = q + 2.0 * p * cos(phi);
    dc_sym3_eigvec(m, l2, w);
    l2 = dc_sym3_rayleigh(m, w);
    dc_sym3_deflate(m, w, l0, l1, e);
    if (want_vecs >= 1) { V[0] = e[0]; V[1] = e[1]; V[2] = e[2]; }
    if (want_vecs >= 3) {
      double v1[3];
      dc_cross(w, e, v1);
      V[3] = v1[0]; V[4] = v1[1]; V[5] = v1[2];
      V[6] = w[0]; V[7] = w[1]; V[8] = w[2];
    }
  } else {
    // the smallest eigenvalue is the isolated one (planar neighbourhoods, the common case)
    l0 = q + 2.0 * p * cos(phi + 2.0943951023931954923);   // + 2 pi / 3
    dc_sym3_eigvec(m, l0, w);
    l0 = dc_sym3_rayleigh(m, w);
    if (want_vecs >= 1) { V[0] = w[0]; V[1] = w[1]; V[2] = w[2]; }
    dc_sym3_deflate(m, w, l1, l2, e);
    if (want_vecs >= 3) {
      double v2[3];
      dc_cross(w, e, v2);
      V[3] = e[0]; V[4] = e[1]; V[5] = e[2];
      V[6] = v2[0]; V[7] = v2[1]; V[8] = v2[2];
    }
  }
  lam[0] = l0 * s; lam[1] = l1 * s; lam[2] = l2 * s;
  return true;
}

// ---------------------------------------------------------------------------------------
// Polynomial bias terms: pw[k] = gamma ^ e[k]  (model.py:243-248, torch.pow in fp64).
// Small integer exponents (the reference's [2,4], [4], [6] ...) use exact repeated squaring.
// ---------------------------------------------------------------------------------------
DC_HD double dc_pow_exp(double g, double e) {
  const double ei = floor(e);
  if (ei == e && e >= 0.0 && e <= 16.0) {
    int n = (int)e;
    double r = 1.0, b = g;
    while (n) { if (n & 1) r *= b; b *= b; n >>= 1; }
    return r;
  }
  return pow(g, e);
}

// d(gamma^e)/de = gamma^e ln(gamma), with torch's convention 0 at gamma == 0, e >= 0.
DC_HD double dc_pow_exp_dlog(double g, double e, double pw) {
  if (g == 0.0 && e >= 0.0) return 0.0;
  return pw * log(g);
}

// ---------------------------------------------------------------------------------------
// SE(3) correction: T = P * [R(omega) t; 0 1]   (eval.py:68-82, transform.py:68-78;
// pytorch3d axis_angle -> quaternion -> matrix).  Row-major 3x4 outputs.
// ---------------------------------------------------------------------------------------
struct dc_aa_cache {
  double r, i, j, k, ts, theta, s;
};

DC_HD void dc_axis_angle_matrix(const double w[3], double R[9], dc_aa_cache* cache) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  const double half = 0.5 * th;
  const double s = (th < 1e-6) ? (0.5 - th2 * (1.0 / 48.0)) : (sin(half) / th);
  const double r = cos(half), i = w[0] * s, j = w[1] * s, k = w[2] * s;
  const double ts = 2.0 / (r * r + i * i + j * j + k * k);
  R[0] = 1.0 - ts * (j * j + k * k); R[1] = ts * (i * j - k * r); R[2] = ts * (i * k + j * r);
  R[3] = ts * (i * j + k * r); R[4] = 1.0 - ts * (i * i + k * k); R[5] = ts * (j * k - i * r);
  R[6] = ts * (i * k - j * r); R[7] = ts * (j * k + i * r); R[8] = 1.0 - ts * (i * i + j * j);
  if (cache) { cache->r = r; cache->i = i; cache->j = j; cache->k = k; cache->ts = ts; cache->theta = th; cache->s = s; }
}

// Reverse mode of dc_axis_angle_matrix: gR[9] -> gw[3].
DC_HD void dc_axis_angle_matrix_bwd(const double w[3], const double gR[9], double gw[3]) {
  double R[9];
  dc_aa_cache c;
  dc_axis_angle_matrix(w, R, &c);
  const double r = c.r, i = c.i, j = c.j, k = c.k, ts = c.ts;
  const double A = j * j + k * k, B = i * j - k * r, C = i * k + j * r, D = i * j + k * r, E = i * i + k * k,
               F = j * k - i * r, G = i * k - j * r, H = j * k + i * r, M = i * i + j * j;
  const double g_ts = -gR[0] * A + gR[1] * B + gR[2] * C + gR[3] * D - gR[4] * E + gR[5] * F + gR[6] * G + gR[7] * H - gR[8] * M;
  const double gA = -ts * gR[0], gB = ts * gR[1], gC = ts * gR[2], gD = ts * gR[3], gE = -ts * gR[4],
               gF = ts * gR[5], gG = ts * gR[6], gH = ts * gR[7], gM = -ts * gR[8];
  double gi = gB * j + gC * k + gD * j + gE * 2.0 * i - gF * r + gG * k + gH * r + gM * 2.0 * i;
  double gj = gA * 2.0 * j + gB * i + gC * r + gD * i + gF * k - gG * r + gH * k + gM * 2.0 * j;
  double gk = gA * 2.0 * k - gB * r + gC * i + gD * r + gE * 2.0 * k + gF * j + gG * i + gH * j;
  double gr = -gB * k + gC * j + gD * k - gF * i - gG * j + gH * i;
  const double n = 2.0 / ts;
  const double gn = -2.0 / (n * n) * g_ts;
  gr += 2.0 * r * gn; gi += 2.0 * i * gn; gj += 2.0 * j * gn; gk += 2.0 * k * gn;
  const double gs = gi * w[0] + gj * w[1] + gk * w[2];
  gw[0] = c.s * gi; gw[1] = c.s * gj; gw[2] = c.s * gk;
  const double th = c.theta, half = 0.5 * th;
  double ds;
  if (th < 1e-6) ds = -th * (1.0 / 24.0);
  else ds = cos(half) / (2.0 * th) - sin(half) / (th * th);
  const double gth = gr * (-0.5 * sin(half)) + gs * ds;
  if (th > 0.0) {
    const double f = gth / th;
    gw[0] += f * w[0]; gw[1] += f * w[1]; gw[2] += f * w[2];
  }
}

// T[12] (row-major 3x4) = P[16] (row-major 4x4) * Delta(delta[6] = xyz, axis-angle)
DC_HD void dc_pose_compose(const double P[16], const double delta[6], double T[12]) {
  double Rd[9];
  dc_axis_angle_matrix(delta + 3, Rd, nullptr);
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < 3; ++b)
      T[4 * a + b] = P[4 * a + 0] * Rd[b] + P[4 * a + 1] * Rd[3 + b] + P[4 * a + 2] * Rd[6 + b];
    T[4 * a + 3] = P[4 * a + 0] * delta[0] + P[4 * a + 1] * delta[1] + P[4 * a + 2] * delta[2] + P[4 * a + 3];
  }
}

// gT[12] -> gdelta[6]
DC_HD void dc_pose_compose_bwd(const double P[16], const double delta[6], const double gT[12], double gdelta[6]) {
  double gRd[9];
  for (int c = 0; c < 3; ++c) {
    // dL/dt_delta[c] = sum_a P[a][c] gT[a][3]
    gdelta[c] = P[0 + c] * gT[3] + P[4 + c] * gT[7] + P[8 + c] * gT[11];
    for (int b = 0; b < 3; ++b)
      gRd[3 * c + b] = P[0 + c] * gT[0 + b] + P[4 + c] * gT[4 + b] + P[8 + c] * gT[8 + b];
  }
  dc_axis_angle_matrix_bwd(delta + 3, gRd, gdelta + 3);
}
