// host_tables.cpp — host-side helpers of the product library (no GPU needed):
//   crystallography tables (FCC {111}<110>, HCP prismatic/basal/pyramidal<c+a>/twins) and the
//   integer-exact periodic Voronoi generator used for the synthetic polycrystals of
//   BASELINE.json "configs" (SURVEY.md §8(d)).
// Reference counterpart: absent (/root/reference holds only LICENSE); the tables follow the
// standard crystallography of the VPSC/EVPFFT literature cited in SURVEY.md §0.
#include "../../include/evpfft.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

const double kPi = 3.14159265358979323846;

void zero_phase(evp_phase *p) { std::memset(p, 0, sizeof(*p)); }

void set_sys(evp_phase *p, int s, const double b[3], const double n[3], int mode) {
  double bl = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
  double nl = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  for (int k = 0; k < 3; ++k) {
    p->b[s][k] = b[k] / bl;
    p->n[s][k] = n[k] / nl;
  }
  p->mode[s] = mode;
}

// Miller-Bravais -> Cartesian (a1 || x, c || z)
void hcp_dir(const int uvtw[4], double ca, double out[3]) {
  const double u = uvtw[0], v = uvtw[1], t = uvtw[2], w = uvtw[3];
  out[0] = u - 0.5 * v - 0.5 * t;
  out[1] = (v - t) * std::sqrt(3.0) * 0.5;
  out[2] = w * ca;
}
void hcp_plane(const int hkil[4], double ca, double out[3]) {
  const double h = hkil[0], k = hkil[1], l = hkil[3];
  out[0] = h;
  out[1] = (h + 2.0 * k) / std::sqrt(3.0);
  out[2] = l / ca;
}
void rotz(const double in[3], int k60, double out[3]) {
  const double a = k60 * kPi / 3.0, c = std::cos(a), s = std::sin(a);
  out[0] = c * in[0] - s * in[1];
  out[1] = s * in[0] + c * in[1];
  out[2] = in[2];
}

inline uint64_t splitmix64(uint64_t &state) {
  state += 0x9E3779B97F4A7C15ull;
  uint64_t z = state;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline double u01(uint64_t &state) { return ((double)(splitmix64(state) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

}  // namespace

extern "C" {

int evp_phase_fcc(evp_phase *out, double c11, double c12, double c44, double gamma0, double nrate,
                  double tau0, double tau1, double theta0, double theta1) {
  if (!out) return EVP_ERR_ARG;
  zero_phase(out);
  out->nsys = 12;
  out->nmodes = 1;
  double *c = out->c_voigt;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[6 * i + j] = (i == j) ? c11 : c12;
  for (int i = 3; i < 6; ++i) c[6 * i + i] = c44;
  static const double N[4][3] = {{1, 1, 1}, {-1, 1, 1}, {1, -1, 1}, {1, 1, -1}};
  static const double B[12][3] = {{0, 1, -1}, {1, 0, -1}, {1, -1, 0}, {0, 1, -1}, {1, 0, 1}, {1, 1, 0},
                                  {0, 1, 1},  {1, 0, -1}, {1, 1, 0},  {0, 1, 1},  {1, 0, 1}, {1, -1, 0}};
  for (int s = 0; s < 12; ++s) set_sys(out, s, B[s], N[s / 3], 0);
  out->twin[0] = 0;
  out->gamma0[0] = gamma0;
  out->nrate[0] = nrate;
  out->tau0[0] = tau0;
  out->tau1[0] = tau1;
  out->theta0[0] = theta0;
  out->theta1[0] = theta1;
  out->hlat[0][0] = 1.0;
  return EVP_OK;
}

// modes: 0 prismatic<a> (3), 1 basal<a> (3), 2 pyramidal<c+a> 1st order (12), 3 {10-12} tensile twin (6)
// with_twin: 0 = slip only (18 systems), 1 = + tensile twins (24), 2 = + {11-22} compressive twins (30, mode 4)
int evp_phase_hcp(evp_phase *out, double ca, const double c5[5], int32_t with_twin, double gamma0, double nrate,
                  const double tau0_mode[4], const double voce_mode[4][3]) {
  if (!out || !c5 || !tau0_mode || !(ca > 0)) return EVP_ERR_ARG;
  zero_phase(out);
  const double c11 = c5[0], c12 = c5[1], c13 = c5[2], c33 = c5[3], c44 = c5[4];
  double *c = out->c_voigt;
  c[0] = c[7] = c11; c[1] = c[6] = c12; c[2] = c[12] = c[8] = c[13] = c13; c[14] = c33;
  c[21] = c[28] = c44; c[35] = 0.5 * (c11 - c12);
  int s = 0;
  double b0[3], n0[3], b[3], n[3];
  {  // prismatic (10-10)[-12-10]
    const int pl[4] = {1, 0, -1, 0}, dr[4] = {-1, 2, -1, 0};
    hcp_plane(pl, ca, n0); hcp_dir(dr, ca, b0);
    for (int k = 0; k < 3; ++k) { rotz(b0, k, b); rotz(n0, k, n); set_sys(out, s++, b, n, 0); }
  }
  {  // basal (0001)[2-1-10]
    const int pl[4] = {0, 0, 0, 1}, dr[4] = {2, -1, -1, 0};
    hcp_plane(pl, ca, n0); hcp_dir(dr, ca, b0);
    for (int k = 0; k < 3; ++k) { rotz(b0, k, b); rotz(n0, k, n); set_sys(out, s++, b, n, 1); }
  }
  {  // pyramidal <c+a> (10-11)[-1-123] and (10-11)[-2113]
    const int pl[4] = {1, 0, -1, 1}, d1[4] = {-1, -1, 2, 3}, d2[4] = {-2, 1, 1, 3};
    hcp_plane(pl, ca, n0);
    for (int which = 0; which < 2; ++which) {
      hcp_dir(which ? d2 : d1, ca, b0);
      for (int k = 0; k < 6; ++k) { rotz(b0, k, b); rotz(n0, k, n); set_sys(out, s++, b, n, 2); }
    }
  }
  int nmodes = 3;
  if (with_twin >= 1) {  // {10-12} twin: shear along [-1011] (extension of c) for c/a < sqrt 3, reversed ([10-1-1]) above (Zn, Cd)
    const int pl[4] = {1, 0, -1, 2}, dr[4] = {-1, 0, 1, 1};
    hcp_plane(pl, ca, n0); hcp_dir(dr, ca, b0);
    if (ca * ca > 3.0) for (double &v : b0) v = -v;
    for (int k = 0; k < 6; ++k) { rotz(b0, k, b); rotz(n0, k, n); set_sys(out, s++, b, n, 3); }
    out->twin[3] = 1;
    out->twin_shear[3] = std::fabs(ca * ca - 3.0) / (std::sqrt(3.0) * ca);
    nmodes = 4;
  }
  if (with_twin >= 2) {  // compressive twin (11-22)[11-2-3]
    const int pl[4] = {1, 1, -2, 2}, dr[4] = {1, 1, -2, -3};
    hcp_plane(pl, ca, n0); hcp_dir(dr, ca, b0);
    for (int k = 0; k < 6; ++k) { rotz(b0, k, b); rotz(n0, k, n); set_sys(out, s++, b, n, 4); }
    out->twin[4] = 1;
    out->twin_shear[4] = 2.0 * (ca * ca - 2.0) / (3.0 * ca);
    nmodes = 5;
  }
  out->nsys = s;
  out->nmodes = nmodes;
  for (int m = 0; m < nmodes; ++m) {
    const int mm = std::min(m, 3);
    out->gamma0[m] = gamma0;
    out->nrate[m] = nrate;
    out->tau0[m] = tau0_mode[mm];
    if (voce_mode) {
      out->tau1[m] = voce_mode[mm][0];
      out->theta0[m] = voce_mode[mm][1];
      out->theta1[m] = voce_mode[mm][2];
    }
    for (int m2 = 0; m2 < nmodes; ++m2) out->hlat[m][m2] = 1.0;
  }
  out->twin_thr1 = 0.1;
  out->twin_thr2 = 0.5;
  return EVP_OK;
}

// Periodic Voronoi tessellation in integer arithmetic on the 2x refined lattice:
// voxel centre (2i+1), seed s in [0, 2n); squared periodic distance; ties -> lowest grain id.
// Seeds and orientations come from one splitmix64 stream: first 3*ngrains position draws, then
// per grain 4 normals (Box-Muller) -> unit quaternion -> rotation matrix (crystal -> sample).
int evp_voronoi(const evp_grid *g, int32_t ngrains, uint64_t seed, int32_t z0, int32_t nzl, int32_t *grain_out,
                double *rot_out) {
  if (!g || ngrains < 1 || g->nx < 1 || g->ny < 1 || g->nz < 1) return EVP_ERR_ARG;
  if (z0 < 0 || nzl < 0 || z0 + nzl > g->nz) return EVP_ERR_ARG;
  const int nx = g->nx, ny = g->ny, nz = g->nz;
  const int64_t LX = 2 * (int64_t)nx, LY = 2 * (int64_t)ny, LZ = 2 * (int64_t)nz;
  uint64_t st = seed;
  std::vector<int64_t> sx(ngrains), sy(ngrains), sz(ngrains);
  for (int i = 0; i < ngrains; ++i) {
    sx[i] = (int64_t)(splitmix64(st) % (uint64_t)LX);
    sy[i] = (int64_t)(splitmix64(st) % (uint64_t)LY);
    sz[i] = (int64_t)(splitmix64(st) % (uint64_t)LZ);
  }
  if (rot_out) {
    for (int i = 0; i < ngrains; ++i) {
      double q[4];
      for (int k = 0; k < 4; k += 2) {
        const double u1 = u01(st), u2 = u01(st);
        const double r = std::sqrt(-2.0 * std::log(u1));
        q[k] = r * std::cos(2.0 * kPi * u2);
        q[k + 1] = r * std::sin(2.0 * kPi * u2);
      }
      const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      const double w = q[0] / nq, x = q[1] / nq, y = q[2] / nq, z = q[3] / nq;
      double *R = rot_out + 9 * (size_t)i;
      R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
      R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
      R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
    }
  }
  if (!grain_out || nzl == 0) return EVP_OK;

  // cell list on the refined lattice: ~2 seeds per cell
  int nc = (int)std::floor(std::cbrt((double)ngrains / 2.0));
  nc = std::max(1, std::min(nc, 64));
  const int ncx = std::max(1, std::min(nc, nx)), ncy = std::max(1, std::min(nc, ny)), ncz = std::max(1, std::min(nc, nz));
  // cell edge (refined units), cells cover [0,L) with the last one possibly larger
  const int64_t cx = LX / ncx, cy = LY / ncy, cz = LZ / ncz;
  auto cell_of = [](int64_t p, int64_t c, int n) { return (int)std::min<int64_t>(p / c, n - 1); };
  std::vector<std::vector<int>> cells((size_t)ncx * ncy * ncz);
  for (int i = 0; i < ngrains; ++i)
    cells[((size_t)cell_of(sz[i], cz, ncz) * ncy + cell_of(sy[i], cy, ncy)) * ncx + cell_of(sx[i], cx, ncx)].push_back(i);
  const int64_t cmin = std::min(cx, std::min(cy, cz));
  const int rmax = std::max(ncx, std::max(ncy, ncz)) / 2 + 1;

#pragma omp parallel for collapse(2) schedule(static)
  for (int z = z0; z < z0 + nzl; ++z)
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x) {
        const int64_t px = 2 * x + 1, py = 2 * y + 1, pz = 2 * z + 1;
        const int ccx = cell_of(px, cx, ncx), ccy = cell_of(py, cy, ncy), ccz = cell_of(pz, cz, ncz);
        int64_t best = INT64_MAX;
        int bid = -1;
        for (int r = 0; r <= rmax; ++r) {
          // scan the shell of cells at Chebyshev offset r (periodic wrap; a cell may be seen twice on
          // tiny cell grids, which is harmless)
          for (int dz = -r; dz <= r; ++dz)
            for (int dy = -r; dy <= r; ++dy)
              for (int dx = -r; dx <= r; ++dx) {
                if (std::max(std::abs(dx), std::max(std::abs(dy), std::abs(dz))) != r) continue;
                const int gx = ((ccx + dx) % ncx + ncx) % ncx, gy = ((ccy + dy) % ncy + ncy) % ncy, gz = ((ccz + dz) % ncz + ncz) % ncz;
                for (int id : cells[((size_t)gz * ncy + gy) * ncx + gx]) {
                  int64_t ddx = std::llabs(px - sx[id]); ddx = std::min(ddx, LX - ddx);
                  int64_t ddy = std::llabs(py - sy[id]); ddy = std::min(ddy, LY - ddy);
                  int64_t ddz = std::llabs(pz - sz[id]); ddz = std::min(ddz, LZ - ddz);
                  const int64_t d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                  if (d2 < best || (d2 == best && id < bid)) { best = d2; bid = id; }
                }
              }
          // any seed image in a cell at offset > r is strictly farther than r*cmin along one axis
          const int64_t lb = (int64_t)r * cmin;
          if (bid >= 0 && lb * lb >= best) break;
        }
        grain_out[((size_t)(z - z0) * ny + y) * nx + x] = bid;
      }
  return EVP_OK;
}


// ---------------------------------------------------------------------------------------------
// Per-voxel text microstructure: "phi1 Phi phi2 i j k grain phase", Bunge angles in degrees,
// 1-based voxel indices in any order, 1-based phase ids.  The common exchange format of the EVPFFT / VPSC family
// [literature; unverified for LApx, whose reader is not in the mount].
// ---------------------------------------------------------------------------------------------
static void bunge_to_rot(double p1, double P, double p2, double *R /* crystal -> sample, row major */) {
  const double d2r = kPi / 180.0;
  const double c1 = std::cos(p1 * d2r), s1 = std::sin(p1 * d2r), c = std::cos(P * d2r), s = std::sin(P * d2r), c2 = std::cos(p2 * d2r),
               s2 = std::sin(p2 * d2r);
  // g (sample -> crystal) = Rz(phi2) Rx(Phi) Rz(phi1); crystal -> sample = g^T
  const double g[9] = {c1 * c2 - s1 * s2 * c, s1 * c2 + c1 * s2 * c, s2 * s, -c1 * s2 - s1 * c2 * c, -s1 * s2 + c1 * c2 * c, c2 * s, s1 * s, -c1 * s, c};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = g[3 * j + i];
}
static void rot_to_bunge(const double *R, double *p1, double *P, double *p2) {
  // g = R^T;  g[2][2] = cos Phi, g[2][0] = s1 s, g[2][1] = -c1 s, g[0][2] = s2 s, g[1][2] = c2 s
  const double r2d = 180.0 / kPi;
  const double sP = std::sqrt(R[2] * R[2] + R[5] * R[5]);
  const double Phi = std::atan2(sP, R[8]);
  if (sP > 1e-9) {
    *p1 = std::atan2(R[2], -R[5]) * r2d;    // g[2][0] = R[0][2], g[2][1] = R[1][2]
    *p2 = std::atan2(R[6], R[7]) * r2d;     // g[0][2] = R[2][0], g[1][2] = R[2][1]
  } else {                                  // Phi = 0 or 180: only phi1 +- phi2 is defined; put it all into phi1
    *p1 = std::atan2(R[3], R[0]) * r2d;     // g[0][1] = R[1][0], g[0][0] = R[0][0]
    *p2 = 0.0;
  }
  *P = Phi * r2d;
}

int evp_read_microstructure_txt(const char *path, const evp_grid *g, int32_t *grain, int32_t *phase, double *rot9) {
  if (!path || !g || !grain || !rot9) return EVP_ERR_ARG;
  std::FILE *f = std::fopen(path, "r");
  if (!f) return EVP_ERR_ARG;
  const size_t N = (size_t)g->nx * g->ny * g->nz;
  std::vector<char> seen(N, 0);
  double p1, P, p2;
  long i, j, k;
  int gr, ph;
  size_t n = 0;
  int rc = EVP_OK;
  while (std::fscanf(f, "%lf %lf %lf %ld %ld %ld %d %d", &p1, &P, &p2, &i, &j, &k, &gr, &ph) == 8) {
    if (i < 1 || j < 1 || k < 1 || i > g->nx || j > g->ny || k > g->nz) { rc = EVP_ERR_ARG; break; }
    const size_t v = ((size_t)(k - 1) * g->ny + (j - 1)) * g->nx + (i - 1);
    if (seen[v]) { rc = EVP_ERR_ARG; break; }   // a voxel listed twice
    seen[v] = 1;
    double R[9];
    bunge_to_rot(p1, P, p2, R);
    for (int c = 0; c < 9; ++c) rot9[c * N + v] = R[c];
    grain[v] = gr;
    if (phase) phase[v] = ph >= 1 ? ph - 1 : 0;
    ++n;
  }
  std::fclose(f);
  if (rc == EVP_OK && n != N) rc = EVP_ERR_ARG;   // missing voxels / trailing garbage
  return rc;
}

int evp_write_microstructure_txt(const char *path, const evp_grid *g, const int32_t *grain, const int32_t *phase, const double *rot9) {
  if (!path || !g || !grain || !rot9) return EVP_ERR_ARG;
  std::FILE *f = std::fopen(path, "w");
  if (!f) return EVP_ERR_ARG;
  const size_t N = (size_t)g->nx * g->ny * g->nz;
  for (int k = 0; k < g->nz; ++k)
    for (int j = 0; j < g->ny; ++j)
      for (int i = 0; i < g->nx; ++i) {
        const size_t v = ((size_t)k * g->ny + j) * g->nx + i;
        double R[9], p1, P, p2;
        for (int c = 0; c < 9; ++c) R[c] = rot9[c * N + v];
        rot_to_bunge(R, &p1, &P, &p2);
        std::fprintf(f, "%.17g %.17g %.17g %d %d %d %d %d\n", p1, P, p2, i + 1, j + 1, k + 1, grain[v], phase ? phase[v] + 1 : 1);
      }
  return std::fclose(f) == 0 ? EVP_OK : EVP_ERR_ARG;
}

}  // extern "C"
