// sb_tables.cpp -- host-side table generation: cubic B-spline weights and banded Cholesky
// factors, Chebyshev matrices, Bluestein ring plans.  Pure C++ (no CUDA).
//
// Algorithm provenance: Ooyama (2002) cubic-spline transform and the FFTW R2HC / REDFT00
// conventions, as constrained by the Scythe.jl call sites (SURVEY.md App. A.2/A.3;
// /root/reference/src/spectralGrid.jl:20-45 for the dimensions).
#include "sb_internal.hpp"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace sb {

static const double kPi = 3.14159265358979323846264338327950288;
static const long double kPiL = 3.14159265358979323846264338327950288L;

// ------------------------------------------------------------------ splines
static double bspline(double delta, int deriv, double DXr) {
  // delta = (x - x_m)/DX ; returns d^deriv/dx^deriv of the cubic B-spline centred on x_m
  double z = std::fabs(delta);
  if (z >= 2.0) return 0.0;
  double sgn = (delta > 0) ? -1.0 : 1.0;
  double z2 = 2.0 - z;
  double z1 = (z < 1.0) ? 1.0 - z : 0.0;
  switch (deriv) {
    case 0: return (z2 * z2 * z2 - 4.0 * z1 * z1 * z1) / 6.0;
    case 1: return sgn * 3.0 * DXr * (z2 * z2 - 4.0 * z1 * z1) / 6.0;
    case 2: return DXr * DXr * (z2 - 4.0 * z1);
    default: return sgn * DXr * DXr * DXr * ((z > 1.0) ? 1.0 : ((z < 1.0) ? -3.0 : 0.0));
  }
}

static const double kGauss[3] = {-0.7745966692414834 /* -sqrt(3/5) */, 0.0, 0.7745966692414834};
static const double kGaussW[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};

static void spline_weights4(double DX, double phi[4][3][4], double wq[3]) {
  const double DXr = 1.0 / DX;
  for (int mu = 0; mu < 3; ++mu) {
    double t = 0.5 + 0.5 * kGauss[mu];  // position inside the cell in [0,1]
    wq[mu] = DX * kGaussW[mu];
    for (int j = 0; j < 4; ++j) {
      double delta = t + 1.0 - j;       // (x - x_{cell-1+j}) / DX
      for (int d = 0; d < 4; ++d) phi[d][mu][j] = bspline(delta, d, DXr);
    }
  }
}

void spline_weights(double DX, double phi[3][3][4], double wq[3]) {
  double p4[4][3][4];
  spline_weights4(DX, p4, wq);
  std::memcpy(phi, p4, sizeof(double) * 3 * 3 * 4);
}

void spline_mish_points(double xmin, double DX, int num_cells, std::vector<double>& r) {
  r.resize((size_t)num_cells * 3);
  for (int c = 0; c < num_cells; ++c)
    for (int mu = 0; mu < 3; ++mu) r[(size_t)c * 3 + mu] = xmin + (c + 0.5) * DX + 0.5 * DX * kGauss[mu];
}

static int bc_rank(int bc) {
  switch (bc) {
    case 1: case 2: case 3: return 1;  // R1T0, R1T1, R1T2
    case 4: case 5: return 2;          // R2T10, R2T20
    case 6: return 3;                  // R3
    default: return 0;
  }
}
static void bc_coeffs(int bc, double c[2]) {
  c[0] = c[1] = 0.0;
  switch (bc) {
    case 1: c[0] = -4.0; c[1] = -1.0; break;  // R1T0: a_-1 = -4 a_0 - a_1
    case 2: c[0] = 0.0; c[1] = 1.0; break;    // R1T1: a_-1 = a_1
    case 3: c[0] = 2.0; c[1] = -1.0; break;   // R1T2: a_-1 = 2 a_0 - a_1
    case 4: c[0] = 1.0; c[1] = -0.5; break;   // R2T10: a_-1 = a_1, a_0 = -a_1/2
    case 5: c[0] = -1.0; c[1] = 0.0; break;   // R2T20: a_-1 = -a_1, a_0 = 0
    default: break;
  }
}

// nonzeros of row i of Gamma: (column, value) pairs
static int gamma_row(const SplineFactor& f, int i, int cols[3], double vals[3]) {
  int n = 0;
  cols[n] = i + f.rL; vals[n] = 1.0; ++n;
  if (f.rL == 1 && i < 2) { cols[n] = 0; vals[n] = f.foldL[i]; ++n; }
  if (f.rL == 2 && i == 0) { cols[n] = 0; vals[n] = f.foldL[0]; ++n; cols[n] = 1; vals[n] = f.foldL[1]; ++n; }
  if (f.rR == 1 && i >= f.nfree - 2) { cols[n] = f.M - 1; vals[n] = f.foldR[f.nfree - 1 - i]; ++n; }
  if (f.rR == 2 && i == f.nfree - 1) { cols[n] = f.M - 1; vals[n] = f.foldR[0]; ++n; cols[n] = f.M - 2; vals[n] = f.foldR[1]; ++n; }
  return n;
}

SplineFactor make_spline_factor(int num_cells, double DX, double l_q, int bcl, int bcr) {
  SplineFactor f;
  f.M = num_cells + 3;
  const int M = f.M;
  double phi[4][3][4], wq[3];
  spline_weights4(DX, phi, wq);
  const double eps_q = std::pow(l_q * DX / (2.0 * kPi), 6.0);
  // P+Q in symmetric band storage pq[m][d] = PQ(m, m+d), d = 0..3
  std::vector<double> pq((size_t)M * 4, 0.0);
  for (int c = 0; c < num_cells; ++c)
    for (int mu = 0; mu < 3; ++mu)
      for (int j = 0; j < 4; ++j)
        for (int jj = j; jj < 4; ++jj)
          pq[(size_t)(c + j) * 4 + (jj - j)] +=
              wq[mu] * (phi[0][mu][j] * phi[0][mu][jj] + eps_q * phi[3][mu][j] * phi[3][mu][jj]);
  auto PQ = [&](int a, int b) -> double {
    if (a > b) std::swap(a, b);
    return (b - a > 3) ? 0.0 : pq[(size_t)a * 4 + (b - a)];
  };
  if (bcl == 7 || bcr == 7) {
    if (!(bcl == 7 && bcr == 7)) throw std::invalid_argument("PERIODIC must be set on both ends");
    f.periodic = true;
    const int nc = num_cells;
    f.nfree = nc;
    // G = Gamma PQ Gamma^T with Gamma[m mod nc, m+1] = 1, m = -1..nc+1
    std::vector<double> G((size_t)nc * nc, 0.0);
    auto wrap = [&](int col) { int m = col - 1; return ((m % nc) + nc) % nc; };
    for (int a = 0; a < M; ++a)
      for (int b = std::max(0, a - 3); b <= std::min(M - 1, a + 3); ++b) G[(size_t)wrap(a) * nc + wrap(b)] += PQ(a, b);
    if (!invert(G, nc)) throw std::runtime_error("periodic spline matrix is singular");
    f.dense.assign((size_t)M * M, 0.0);
    for (int a = 0; a < M; ++a)
      for (int b = 0; b < M; ++b) f.dense[(size_t)a * M + b] = G[(size_t)wrap(a) * nc + wrap(b)];
    return f;
  }
  f.rL = bc_rank(bcl);
  f.rR = bc_rank(bcr);
  bc_coeffs(bcl, f.foldL);
  bc_coeffs(bcr, f.foldR);
  f.nfree = M - f.rL - f.rR;
  const int n = f.nfree;
  if (n < 4) throw std::invalid_argument("too few cells for the requested boundary conditions");
  // folded band matrix g[i][d] = G(i, i+d)
  std::vector<double> g((size_t)n * 4, 0.0);
  for (int i = 0; i < n; ++i) {
    int ci[3]; double vi[3];
    int ni = gamma_row(f, i, ci, vi);
    for (int d = 0; d < 4 && i + d < n; ++d) {
      int cj[3]; double vj[3];
      int nj = gamma_row(f, i + d, cj, vj);
      double s = 0.0;
      for (int a = 0; a < ni; ++a)
        for (int b = 0; b < nj; ++b) s += vi[a] * vj[b] * PQ(ci[a], cj[b]);
      g[(size_t)i * 4 + d] = s;
    }
  }
  // banded Cholesky, lower factor l[i][d] = L(i, i-d)
  std::vector<double> l((size_t)n * 4, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int d = std::min(3, i); d >= 0; --d) {
      int j = i - d;  // column
      double s = g[(size_t)j * 4 + d];
      for (int k = std::max(0, i - 3); k < j; ++k) {
        if (j - k > 3) continue;
        s -= l[(size_t)i * 4 + (i - k)] * l[(size_t)j * 4 + (j - k)];
      }
      if (d == 0) {
        if (s <= 0.0) throw std::runtime_error("spline matrix not positive definite");
        l[(size_t)i * 4] = std::sqrt(s);
      } else {
        l[(size_t)i * 4 + d] = s / l[(size_t)j * 4];
      }
    }
  }
  f.chol.resize((size_t)n * 4);
  for (int i = 0; i < n; ++i) {
    f.chol[(size_t)i * 4] = 1.0 / l[(size_t)i * 4];
    for (int d = 1; d < 4; ++d) f.chol[(size_t)i * 4 + d] = l[(size_t)i * 4 + d];
  }
  return f;
}

// ------------------------------------------------------------------ dense helpers
void matmul(const double* A, const double* B, double* C, int n, int k, int m) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) {
      long double s = 0.0L;
      for (int q = 0; q < k; ++q) s += (long double)A[(size_t)i * k + q] * B[(size_t)q * m + j];
      C[(size_t)i * m + j] = (double)s;
    }
}

bool lu_factor(std::vector<double>& A, std::vector<int>& piv, int n) {
  piv.resize(n);
  for (int c = 0; c < n; ++c) {
    int p = c;
    double best = std::fabs(A[(size_t)c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(A[(size_t)r * n + c]) > best) { best = std::fabs(A[(size_t)r * n + c]); p = r; }
    if (best == 0.0) return false;
    piv[c] = p;
    if (p != c)
      for (int k = 0; k < n; ++k) std::swap(A[(size_t)c * n + k], A[(size_t)p * n + k]);
    double inv = 1.0 / A[(size_t)c * n + c];
    for (int r = c + 1; r < n; ++r) {
      double m = A[(size_t)r * n + c] * inv;
      A[(size_t)r * n + c] = m;
      if (m != 0.0)
        for (int k = c + 1; k < n; ++k) A[(size_t)r * n + k] -= m * A[(size_t)c * n + k];
    }
  }
  return true;
}

void lu_solve(const std::vector<double>& LU, const std::vector<int>& piv, int n, double* b) {
  // full rows (incl. the L part) were swapped during factorisation: permute b first, then solve
  for (int c = 0; c < n; ++c)
    if (piv[c] != c) std::swap(b[c], b[piv[c]]);
  for (int c = 0; c < n; ++c)
    for (int r = c + 1; r < n; ++r) b[r] -= LU[(size_t)r * n + c] * b[c];
  for (int r = n - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < n; ++k) s -= LU[(size_t)r * n + k] * b[k];
    b[r] = s / LU[(size_t)r * n + r];
  }
}

bool invert(std::vector<double>& A, int n) {
  std::vector<double> LU = A;
  std::vector<int> piv;
  if (!lu_factor(LU, piv, n)) return false;
  std::vector<double> col(n);
  for (int j = 0; j < n; ++j) {
    std::fill(col.begin(), col.end(), 0.0);
    col[j] = 1.0;
    lu_solve(LU, piv, n, col.data());
    for (int i = 0; i < n; ++i) A[(size_t)i * n + j] = col[i];
  }
  return true;
}

// ------------------------------------------------------------------ Chebyshev
ChebTables make_cheb_tables(int nz, int bz, double zmin, double zmax) {
  ChebTables t;
  t.nz = nz;
  t.bz = bz;
  const double L = zmax - zmin;
  t.z.resize(nz);
  t.fwd.assign((size_t)bz * nz, 0.0);
  t.T0.assign((size_t)nz * nz, 0.0);
  t.T1 = t.T0; t.T2 = t.T0; t.Tint = t.T0;
  const int N1 = nz - 1;
  for (int j = 0; j < nz; ++j) t.z[j] = zmin + 0.5 * L * (1.0 - (double)cosl(kPiL * j / N1));
  // CB: b_k = [u_0 + (-1)^k u_{N-1} + 2 sum u_j cos(pi j k/(N-1))] / (2 (N-1))
  for (int k = 0; k < bz; ++k)
    for (int j = 0; j < nz; ++j) {
      long double w = (j == 0 || j == N1) ? 1.0L : 2.0L;
      long long jk = ((long long)j * k) % (2LL * N1);
      t.fwd[(size_t)k * nz + j] = (double)(w * cosl(kPiL * (long double)jk / N1) / (2.0L * N1));
    }
  // synthesis matrices via three-term recurrences in xi = cos(pi j/(N-1)) (xi=+1 at the bottom)
  for (int j = 0; j < nz; ++j) {
    long double xi = cosl(kPiL * j / N1);
    std::vector<long double> T(nz + 2), D1(nz + 2), D2(nz + 2);
    T[0] = 1; D1[0] = 0; D2[0] = 0;
    T[1] = xi; D1[1] = 1; D2[1] = 0;
    for (int k = 1; k <= nz; ++k) {
      T[k + 1] = 2 * xi * T[k] - T[k - 1];
      D1[k + 1] = 2 * T[k] + 2 * xi * D1[k] - D1[k - 1];
      D2[k + 1] = 4 * D1[k] + 2 * xi * D2[k] - D2[k - 1];
    }
    for (int k = 0; k < nz; ++k) {
      long double s = (k == 0 || k == N1) ? 1.0L : 2.0L;
      // exact endpoint/cosine value for T to avoid recurrence drift
      long long jk = ((long long)j * k) % (2LL * N1);
      long double Tk = cosl(kPiL * (long double)jk / N1);
      t.T0[(size_t)j * nz + k] = (double)(s * Tk);
      t.T1[(size_t)j * nz + k] = (double)(s * (-2.0L / L) * D1[k]);
      t.T2[(size_t)j * nz + k] = (double)(s * (4.0L / ((long double)L * L)) * D2[k]);
      // antiderivative I_k(xi) with int_{zmin}^{z} = (L/2) (I_k(1) - I_k(xi))
      auto I = [&](long double x, const std::vector<long double>& Tx) -> long double {
        if (k == 0) return x;
        if (k == 1) return 0.25L * (2 * x * x - 1);  // T_2/4
        return 0.5L * (Tx[k + 1] / (k + 1) - Tx[k - 1] / (k - 1));
      };
      std::vector<long double> T1v(nz + 2, 1.0L);  // T_k(1) = 1
      t.Tint[(size_t)j * nz + k] = (double)(s * 0.5L * L * (I(1.0L, T1v) - I(xi, T)));
    }
  }
  return t;
}

std::vector<double> cheb_bc_matrix(const ChebTables& t, int bcb, int bct) {
  const int bz = t.bz, nz = t.nz;
  std::vector<double> IG((size_t)bz * bz, 0.0);
  for (int i = 0; i < bz; ++i) IG[(size_t)i * bz + i] = 1.0;
  std::vector<std::vector<double>> rows;
  auto pick = [&](int bc, int lev) {
    if (bc == 0) return;
    const std::vector<double>& T = (bc == 1) ? t.T0 : (bc == 2) ? t.T1 : t.T2;
    rows.emplace_back(T.begin() + (size_t)lev * nz, T.begin() + (size_t)lev * nz + bz);
  };
  pick(bcb, 0);
  pick(bct, nz - 1);
  const int nb = (int)rows.size();
  if (nb == 0) return IG;
  // I - C^T (C C^T)^-1 C
  std::vector<double> CCt((size_t)nb * nb);
  for (int a = 0; a < nb; ++a)
    for (int b = 0; b < nb; ++b) {
      long double s = 0;
      for (int k = 0; k < bz; ++k) s += (long double)rows[a][k] * rows[b][k];
      CCt[(size_t)a * nb + b] = (double)s;
    }
  if (!invert(CCt, nb)) throw std::runtime_error("degenerate vertical boundary conditions");
  for (int i = 0; i < bz; ++i)
    for (int j = 0; j < bz; ++j) {
      long double s = 0;
      for (int a = 0; a < nb; ++a)
        for (int b = 0; b < nb; ++b) s += (long double)rows[a][i] * CCt[(size_t)a * nb + b] * rows[b][j];
      IG[(size_t)i * bz + j] -= (double)s;
    }
  return IG;
}

// ------------------------------------------------------------------ FFT plans
static void cmul(double& xr, double& xi, double wr, double wi) {
  double r = xr * wr - xi * wi, i = xr * wi + xi * wr;
  xr = r; xi = i;
}

void host_fft_dif(double* x, int L, const double* tw) {
  int Ns = L;
  while (Ns >= 4) {
    const int Nq = Ns / 4, step = L / Ns;
    for (int b = 0; b < L; b += Ns)
      for (int j = 0; j < Nq; ++j) {
        double* p0 = x + 2 * (b + j); double* p1 = p0 + 2 * Nq; double* p2 = p1 + 2 * Nq; double* p3 = p2 + 2 * Nq;
        double t0r = p0[0] + p2[0], t0i = p0[1] + p2[1], t1r = p0[0] - p2[0], t1i = p0[1] - p2[1];
        double t2r = p1[0] + p3[0], t2i = p1[1] + p3[1];
        double t3r = p1[1] - p3[1], t3i = -(p1[0] - p3[0]);  // (a1-a3)*(-i)
        double y0r = t0r + t2r, y0i = t0i + t2i, y1r = t1r + t3r, y1i = t1i + t3i;
        double y2r = t0r - t2r, y2i = t0i - t2i, y3r = t1r - t3r, y3i = t1i - t3i;
        cmul(y1r, y1i, tw[2 * (j * step)], tw[2 * (j * step) + 1]);
        cmul(y2r, y2i, tw[2 * (2 * j * step)], tw[2 * (2 * j * step) + 1]);
        cmul(y3r, y3i, tw[2 * (3 * j * step)], tw[2 * (3 * j * step) + 1]);
        p0[0] = y0r; p0[1] = y0i; p1[0] = y1r; p1[1] = y1i; p2[0] = y2r; p2[1] = y2i; p3[0] = y3r; p3[1] = y3i;
      }
    Ns = Nq;
  }
  if (Ns == 2)
    for (int b = 0; b < L; b += 2) {
      double* p = x + 2 * b;
      double ar = p[0], ai = p[1], br = p[2], bi = p[3];
      p[0] = ar + br; p[1] = ai + bi; p[2] = ar - br; p[3] = ai - bi;
    }
}

void host_fft_dit(double* x, int L, const double* tw) {
  int log2L = 0;
  while ((1 << log2L) < L) ++log2L;
  int Ns = (log2L & 1) ? 2 : 4;
  if (Ns == 2) {
    for (int b = 0; b < L; b += 2) {
      double* p = x + 2 * b;
      double ar = p[0], ai = p[1], br = p[2], bi = p[3];
      p[0] = ar + br; p[1] = ai + bi; p[2] = ar - br; p[3] = ai - bi;
    }
    Ns = 8;
  }
  for (; Ns <= L; Ns *= 4) {
    const int Nq = Ns / 4, step = L / Ns;
    for (int b = 0; b < L; b += Ns)
      for (int j = 0; j < Nq; ++j) {
        double* p0 = x + 2 * (b + j); double* p1 = p0 + 2 * Nq; double* p2 = p1 + 2 * Nq; double* p3 = p2 + 2 * Nq;
        double a1r = p1[0], a1i = p1[1], a2r = p2[0], a2i = p2[1], a3r = p3[0], a3i = p3[1];
        cmul(a1r, a1i, tw[2 * (j * step)], -tw[2 * (j * step) + 1]);
        cmul(a2r, a2i, tw[2 * (2 * j * step)], -tw[2 * (2 * j * step) + 1]);
        cmul(a3r, a3i, tw[2 * (3 * j * step)], -tw[2 * (3 * j * step) + 1]);
        double t0r = p0[0] + a2r, t0i = p0[1] + a2i, t1r = p0[0] - a2r, t1i = p0[1] - a2i;
        double t2r = a1r + a3r, t2i = a1i + a3i;
        double t3r = -(a1i - a3i), t3i = (a1r - a3r);  // (a1-a3)*(+i)
        p0[0] = t0r + t2r; p0[1] = t0i + t2i; p1[0] = t1r + t3r; p1[1] = t1i + t3i;
        p2[0] = t0r - t2r; p2[1] = t0i - t2i; p3[0] = t1r - t3r; p3[1] = t1i - t3i;
      }
  }
}

void build_ring_plans(int min_fast_L, const std::vector<int>& ring_ri, std::vector<FftClass>& classes,
                      std::vector<RingPlan>& plans, std::vector<double>& blob) {
  classes.clear();
  plans.assign(ring_ri.size(), RingPlan());
  blob.clear();
  auto class_of = [&](int L) -> int {
    for (size_t i = 0; i < classes.size(); ++i)
      if (classes[i].L == L) return (int)i;
    FftClass c;
    c.L = L;
    if (L % 3 == 0) {          // composite: three sub-FFTs of length L/3 (sb_ringfft2.cu k_*_l3)
      c.R = 3; c.L2 = L / 3;
      while ((1 << c.log2L) < c.L2) ++c.log2L;
      c.fast = true;
      fast_class_twiddles(c.L2, c.twp, c.twoff);
      std::vector<double> extra;
      fft3_class_tables(c.L2, extra);
      c.twp.insert(c.twp.end(), extra.begin(), extra.end());
      classes.push_back(std::move(c));
      return (int)classes.size() - 1;
    }
    c.L2 = L;
    while ((1 << c.log2L) < L) ++c.log2L;
    c.tw.resize((size_t)2 * L);
    for (int t = 0; t < L; ++t) {
      c.tw[2 * t] = (double)cosl(-2.0L * kPiL * t / L);
      c.tw[2 * t + 1] = (double)sinl(-2.0L * kPiL * t / L);
    }
    c.fast = fast_class_supported(L) && L >= min_fast_L;
    if (c.fast) fast_class_twiddles(L, c.twp, c.twoff);
    classes.push_back(std::move(c));
    return (int)classes.size() - 1;
  };
  for (size_t r = 0; r < ring_ri.size(); ++r) {
    RingPlan& p = plans[r];
    const int ri = ring_ri[r];
    p.n = 4 + 4 * ri;
    p.m = p.n / 4;
    int L = 4;
    while (L < 2 * p.m - 1) L *= 2;
    // smooth lengths: 3/4 of the power of two is enough when 2m-1 <= 3 * 2^(a-2).  Measured on B200: 6144 beats 8192
    // by 1.4x and 3072 beats 4096 slightly; 1536 only ties 2048 (three short sub-FFTs lose the pruning and
    // the radix-8 final pass), so the default threshold is L >= 4096 (SB_FFT3_MINL overrides).
    // Since the v4 kernels (sb_ringfft4.cu: Tensor-Memory tables, bulk-staged rows; L <= 4096) the power of two wins up to
    // 4096 (N = 2 weak-scaling outer tile, m <= 1417: inv_l 11.9 ms with L = 3072 composite against 7.1 ms for the inner
    // tile's 2048 class), so the composite classes start where v4 stops: L = 6144 instead of 8192.
    static const int minL = std::getenv("SB_FFT3_MINL") ? std::atoi(std::getenv("SB_FFT3_MINL")) : (fft4_supported(4096, false) ? 8192 : 4096);
    if (fft3_enabled() && fast_class_supported(L) && L >= minL && L >= 2048 && L <= 8192 && 2 * p.m - 1 <= 3 * (L / 4)) L = 3 * (L / 4);
    p.L = L;
    p.cls = class_of(L);
    p.off = (long long)blob.size();
    const int m = p.m, n = p.n;
    blob.resize(blob.size() + (size_t)6 * m + (size_t)2 * L);
    double* chirp = blob.data() + p.off;
    double* wk = chirp + 2 * m;
    double* ph = wk + 2 * m;
    double* FH = ph + 2 * m;
    for (int a = 0; a < m; ++a) {
      long long q = ((long long)a * a) % (2LL * m);
      long double ang = -kPiL * (long double)q / m;          // exp(-i pi a^2 / m)
      chirp[2 * a] = (double)cosl(ang); chirp[2 * a + 1] = (double)sinl(ang);
      long double w = -2.0L * kPiL * a / n;                   // omega^a = exp(-2 pi i a / n)
      wk[2 * a] = (double)cosl(w); wk[2 * a + 1] = (double)sinl(w);
      long long q2 = ((long long)a * (ri - 1)) % (2LL * n);   // exp(-i a ymin), ymin = pi (ri-1)/n
      long double pa = -kPiL * (long double)q2 / n;
      ph[2 * a] = (double)cosl(pa); ph[2 * a + 1] = (double)sinl(pa);
    }
    // FH = DIF( h ), h_j = conj(chirp_j) at j and L-j, scaled by 1/L, kept in DIF (digit-reversed) order
    std::vector<double> h((size_t)2 * L, 0.0);
    for (int j = 0; j < m; ++j) {
      h[2 * j] = chirp[2 * j]; h[2 * j + 1] = -chirp[2 * j + 1];
      if (j > 0) { h[2 * (L - j)] = chirp[2 * j]; h[2 * (L - j) + 1] = -chirp[2 * j + 1]; }
    }
    const double invL = 1.0 / L;
    if (classes[p.cls].R == 3) {
      // H[3k+r] = DFT_{L2}( (h[j] + w3^r h[j+L2] + w3^2r h[j+2 L2]) W_L^{jr} )[k]; each r kept in DIF16 order, [r][e][tl]
      const int L2 = L / 3, T = L2 / 16;
      typedef std::complex<long double> cld;
      std::vector<double> hs((size_t)2 * L2);
      for (int rr = 0; rr < 3; ++rr) {
        for (int j = 0; j < L2; ++j) {
          cld acc(0.0L, 0.0L);
          for (int q = 0; q < 3; ++q) {
            const cld hv((long double)h[2 * (j + q * L2)], (long double)h[2 * (j + q * L2) + 1]);
            acc += hv * std::polar(1.0L, -2.0L * kPiL * (long double)((rr * q) % 3) / 3.0L);
          }
          acc *= std::polar(1.0L, -2.0L * kPiL * (long double)(((long long)j * rr) % L) / (long double)L);
          hs[2 * j] = (double)acc.real(); hs[2 * j + 1] = (double)acc.imag();
        }
        host_fft_dif16(hs.data(), L2);
        for (int tl = 0; tl < T; ++tl)
          for (int e = 0; e < 16; ++e) {
            FH[2 * ((size_t)rr * L2 + e * T + tl)] = hs[2 * (tl * 16 + e)] * invL;
            FH[2 * ((size_t)rr * L2 + e * T + tl) + 1] = hs[2 * (tl * 16 + e) + 1] * invL;
          }
      }
    } else if (classes[p.cls].fast) {
      // fast kernel: DIF16 order, laid out [e][tl] so that thread tl reads element tl*16+e coalesced
      host_fft_dif16(h.data(), L);
      const int T = L / 16;
      for (int tl = 0; tl < T; ++tl)
        for (int e = 0; e < 16; ++e) {
          FH[2 * (e * T + tl)] = h[2 * (tl * 16 + e)] * invL;
          FH[2 * (e * T + tl) + 1] = h[2 * (tl * 16 + e) + 1] * invL;
        }
    } else {
      host_fft_dif(h.data(), L, classes[p.cls].tw.data());
      for (int i = 0; i < 2 * L; ++i) FH[i] = h[i] * invL;
    }
      if (classes[p.cls].fast) {
      // v2 pre-combined tables (long double): see sb_internal.hpp for the formulas they implement
      typedef std::complex<long double> cld;
      const cld I(0.0L, 1.0L);
      auto chirpL = [&](int a) { long long q = ((long long)a * a) % (2LL * m); return std::polar(1.0L, -kPiL * (long double)q / m); };
      auto wL = [&](int a) { return std::polar(1.0L, -2.0L * kPiL * (long double)a / n); };
      auto phL = [&](int a) { long long q2 = ((long long)a * (ri - 1)) % (2LL * n); return std::polar(1.0L, -kPiL * (long double)q2 / n); };
      p.off2 = (long long)blob.size();
      blob.resize(blob.size() + (size_t)16 * m);
      double* PQ = blob.data() + p.off2;
      double* AF = PQ + 8 * m;
      for (int k = 0; k < m; ++k) {
        const int km = k ? m - k : 0;
        const cld ek = k ? 2.0L * std::conj(phL(k)) : cld(1.0L), ekm = km ? 2.0L * std::conj(phL(km)) : cld(1.0L);
        const cld wk_ = wL(k), wkm = wL(km), ck = chirpL(k), ckm = chirpL(km);
        for (int h = 0; h < 2; ++h) {
          const cld sk = h ? std::conj(wk_ * wk_) : cld(1.0L), skm = h ? std::conj(wkm * wkm) : cld(1.0L);
          const cld alpha = 0.5L * ek * sk * (cld(1.0L) + I * std::conj(wk_));
          const cld beta = 0.5L * std::conj(ekm * skm) * (cld(1.0L) + I * wkm);
          const cld P = std::conj(alpha) * ck, Q = std::conj(beta) * ck;
          PQ[(size_t)(2 * h) * 2 * m + 2 * k] = (double)P.real(); PQ[(size_t)(2 * h) * 2 * m + 2 * k + 1] = (double)P.imag();
          PQ[(size_t)(2 * h + 1) * 2 * m + 2 * k] = (double)Q.real(); PQ[(size_t)(2 * h + 1) * 2 * m + 2 * k + 1] = (double)Q.imag();
        }
        const cld pn = phL(k) / (long double)n, w1 = wk_, w2 = w1 * w1, w3 = w2 * w1;
        const cld A[4] = {pn * 0.5L * (cld(1.0L) - I * w1) * ck, pn * 0.5L * (cld(1.0L) + I * w1) * std::conj(ckm),
                          pn * 0.5L * (w2 - I * w3) * ck, pn * 0.5L * (w2 + I * w3) * std::conj(ckm)};
        for (int j = 0; j < 4; ++j) { AF[(size_t)j * 2 * m + 2 * k] = (double)A[j].real(); AF[(size_t)j * 2 * m + 2 * k + 1] = (double)A[j].imag(); }
      }
    }
  }
}

}  // namespace sb
