// evpfft_driver — host driver: deck reader, increment loop, stress-strain curve and field dumps.
// C++17 above the C ABI of include/evpfft.h; it links only against libevpfft_b200.so (no CPU fallback).
// The reference's deck format is unknown (mount holds only LICENSE): the keyed text deck below is ours; the reader sits
// behind read_deck() so that a genuine LApx deck reader can replace it.
//
//   evpfft_driver deck.txt            run
//   evpfft_driver --check deck.txt    parse and print the deck, do not touch the GPU
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/evpfft.h"

struct Deck {
  evp_grid grid{32, 32, 32, 1, 1, 1};
  std::string micro_kind = "voronoi", micro_file;
  int ngrains = 50;
  uint64_t seed = 0;
  std::string crystal = "fcc";
  double c3[3] = {168400, 121400, 75400}, covera = 1.594, c5[5] = {143500, 72500, 65400, 164900, 32100};
  double gamma0 = 1.0, nrate = 10, tau0 = 16, tau1 = 0, theta0 = 0, theta1 = 0;
  double hcp_tau0[4] = {20, 100, 160, 80}, hcp_voce[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  int with_twin = 1;
  bool c0_average = true;
  double c0[36]{};
  int iudot[9] = {1, 1, 1, 1, 1, 1, 1, 1, 1}, iscau[6] = {0, 0, 0, 0, 0, 0};
  double udot[9]{}, scau[6]{};
  double dt = 1e-4;
  int increments = 1;
  evp_ctrl ctrl{1e-5, 1e-5, 100, 1, 1e-6, 100, 0, 0};
  std::string curve = "curve.txt";
  std::vector<std::pair<std::string, std::string>> field_out;
};

static bool fail(const std::string &m) { std::cerr << "evpfft_driver: " << m << "\n"; return false; }

bool read_deck(const std::string &path, Deck &d) {
  std::ifstream in(path);
  if (!in) return fail("cannot open deck " + path);
  std::string line;
  int ln = 0;
  while (std::getline(in, line)) {
    ++ln;
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line.resize(hash);
    std::istringstream ss(line);
    std::string key;
    if (!(ss >> key)) continue;
    auto bad = [&]() { return fail(path + ":" + std::to_string(ln) + ": bad '" + key + "' line"); };
    if (key == "grid") { if (!(ss >> d.grid.nx >> d.grid.ny >> d.grid.nz)) return bad(); }
    else if (key == "spacing") { if (!(ss >> d.grid.dx >> d.grid.dy >> d.grid.dz)) return bad(); }
    else if (key == "microstructure") {
      if (!(ss >> d.micro_kind)) return bad();
      if (d.micro_kind == "voronoi") { if (!(ss >> d.ngrains >> d.seed)) return bad(); }
      else if (d.micro_kind == "file") { if (!(ss >> d.micro_file)) return bad(); }
      else return bad();
    } else if (key == "phase") {
      if (!(ss >> d.crystal)) return bad();
      if (d.crystal == "fcc") { if (!(ss >> d.c3[0] >> d.c3[1] >> d.c3[2])) return bad(); }
      else if (d.crystal == "hcp") { if (!(ss >> d.covera >> d.c5[0] >> d.c5[1] >> d.c5[2] >> d.c5[3] >> d.c5[4] >> d.with_twin)) return bad(); }
      else return bad();
    } else if (key == "rate") { if (!(ss >> d.gamma0 >> d.nrate)) return bad(); }
    else if (key == "voce") { if (!(ss >> d.tau0 >> d.tau1 >> d.theta0 >> d.theta1)) return bad(); }
    else if (key == "voce_mode") {
      int m;
      if (!(ss >> m) || m < 0 || m > 3 || !(ss >> d.hcp_tau0[m] >> d.hcp_voce[m][0] >> d.hcp_voce[m][1] >> d.hcp_voce[m][2])) return bad();
    } else if (key == "reference_medium") {
      std::string k;
      if (!(ss >> k)) return bad();
      d.c0_average = (k == "average");
      if (!d.c0_average) { if (k != "voigt") return bad(); for (double &v : d.c0) if (!(ss >> v)) return bad(); }
    } else if (key == "loading") {
      std::string k;
      if (!(ss >> k)) return bad();
      for (int i = 0; i < 9; ++i) { d.iudot[i] = 1; d.udot[i] = 0; }
      for (int i = 0; i < 6; ++i) { d.iscau[i] = 0; d.scau[i] = 0; }
      if (k == "uniaxial_tension") {
        int ax; double rate;
        if (!(ss >> ax >> rate) || ax < 1 || ax > 3) return bad();
        for (int a = 0; a < 3; ++a) if (a != ax - 1) { d.iudot[4 * a] = 0; d.iscau[a] = 1; }
        d.udot[4 * (ax - 1)] = rate;
      } else if (k == "strain_rate") { for (double &v : d.udot) if (!(ss >> v)) return bad(); }
      else if (k == "plane_strain_compression") {
        double rate;
        if (!(ss >> rate)) return bad();
        d.iudot[0] = 0; d.iscau[0] = 1; d.udot[8] = -rate;
      } else if (k == "mixed") {
        for (int &v : d.iudot) if (!(ss >> v)) return bad();
        for (double &v : d.udot) if (!(ss >> v)) return bad();
        for (int &v : d.iscau) if (!(ss >> v)) return bad();
        for (double &v : d.scau) if (!(ss >> v)) return bad();
      } else return bad();
    } else if (key == "dt") { if (!(ss >> d.dt)) return bad(); }
    else if (key == "increments") { if (!(ss >> d.increments)) return bad(); }
    else if (key == "tol") { if (!(ss >> d.ctrl.tol_stress >> d.ctrl.tol_strain)) return bad(); }
    else if (key == "itmax") { if (!(ss >> d.ctrl.itmax)) return bad(); }
    else if (key == "tol_newton") { if (!(ss >> d.ctrl.tol_newton)) return bad(); }
    else if (key == "update_texture") { if (!(ss >> d.ctrl.update_texture)) return bad(); }
    else if (key == "update_twinning") { if (!(ss >> d.ctrl.update_twinning)) return bad(); }
    else if (key == "output_curve") { if (!(ss >> d.curve)) return bad(); }
    else if (key == "output_field") { std::string f, p; if (!(ss >> f >> p)) return bad(); d.field_out.emplace_back(f, p); }
    else return fail(path + ":" + std::to_string(ln) + ": unknown key '" + key + "'");
  }
  return true;
}

// per-voxel text file "phi1 Phi phi2 i j k grain phase" (SURVEY.md §8(f).3): the library's reader, shared with the Python mirror
bool read_micro_file(const Deck &d, std::vector<int32_t> &grain, std::vector<int32_t> &phase, std::vector<double> &rot9) {
  const size_t N = (size_t)d.grid.nx * d.grid.ny * d.grid.nz;
  grain.assign(N, -1); phase.assign(N, 0); rot9.assign(9 * N, 0.0);
  if (evp_read_microstructure_txt(d.micro_file.c_str(), &d.grid, grain.data(), phase.data(), rot9.data()) != 0)
    return fail("cannot read microstructure file " + d.micro_file + " (missing file, voxel index out of range, duplicate or missing voxels)");
  return true;
}

void print_deck(const Deck &d) {
  std::printf("grid %d %d %d  spacing %g %g %g\n", d.grid.nx, d.grid.ny, d.grid.nz, d.grid.dx, d.grid.dy, d.grid.dz);
  if (d.micro_kind == "voronoi") std::printf("microstructure voronoi %d grains seed %llu\n", d.ngrains, (unsigned long long)d.seed);
  else std::printf("microstructure file %s\n", d.micro_file.c_str());
  std::printf("phase %s  rate gamma0 %g n %g\n", d.crystal.c_str(), d.gamma0, d.nrate);
  std::printf("loading iudot"); for (int v : d.iudot) std::printf(" %d", v);
  std::printf(" udot"); for (double v : d.udot) std::printf(" %g", v);
  std::printf(" iscau"); for (int v : d.iscau) std::printf(" %d", v);
  std::printf("\ndt %g increments %d tol %g %g itmax %d tol_newton %g texture %d twinning %d\n", d.dt, d.increments, d.ctrl.tol_stress,
              d.ctrl.tol_strain, d.ctrl.itmax, d.ctrl.tol_newton, d.ctrl.update_texture, d.ctrl.update_twinning);
}

#define EVP(call)                                                                                     \
  do {                                                                                                \
    const int rc_ = (call);                                                                           \
    if (rc_ != 0) { std::cerr << "evpfft_driver: " #call " failed (" << rc_ << "): " << evp_last_error(h) << "\n"; return 2; } \
  } while (0)

int main(int argc, char **argv) {
  bool check = false;
  std::string path;
  for (int a = 1; a < argc; ++a) { if (std::string(argv[a]) == "--check") check = true; else path = argv[a]; }
  if (path.empty()) { std::cerr << "usage: evpfft_driver [--check] deck.txt\n"; return 1; }
  Deck d;
  if (!read_deck(path, d)) return 1;
  print_deck(d);
  if (check) return 0;

  evp_phase ph;
  if (d.crystal == "fcc") evp_phase_fcc(&ph, d.c3[0], d.c3[1], d.c3[2], d.gamma0, d.nrate, d.tau0, d.tau1, d.theta0, d.theta1);
  else evp_phase_hcp(&ph, d.covera, d.c5, d.with_twin, d.gamma0, d.nrate, d.hcp_tau0, d.hcp_voce);
  evp_handle h = nullptr;
  EVP(evp_create(&d.grid, &ph, 1, nullptr, &h));
  const size_t N = (size_t)d.grid.nx * d.grid.ny * d.grid.nz;
  std::vector<int32_t> grain, phase;
  std::vector<double> rot9;
  if (d.micro_kind == "voronoi") {
    grain.resize(N); rot9.resize(9 * N);
    std::vector<double> grot((size_t)9 * d.ngrains);
    if (evp_voronoi(&d.grid, d.ngrains, d.seed, 0, d.grid.nz, grain.data(), grot.data()) != 0) { fail("evp_voronoi failed"); return 2; }
    for (size_t v = 0; v < N; ++v)
      for (int c = 0; c < 9; ++c) rot9[c * N + v] = grot[(size_t)9 * grain[v] + c];
  } else if (!read_micro_file(d, grain, phase, rot9)) return 1;
  EVP(evp_set_microstructure(h, grain.data(), phase.empty() ? nullptr : phase.data(), rot9.data()));
  EVP(evp_set_reference_medium(h, d.c0_average ? nullptr : d.c0));
  EVP(evp_set_control(h, &d.ctrl));
  EVP(evp_set_loading(h, d.iudot, d.udot, d.iscau, d.scau));
  std::FILE *fc = std::fopen(d.curve.c_str(), "w");
  if (!fc) { fail("cannot write " + d.curve); return 1; }
  std::fprintf(fc, "# inc iters converged err_stress err_strain E11 E22 E33 E23 E13 E12 S11 S22 S33 S23 S13 S12 EP11 EP22 EP33 EP23 EP13 EP12 seconds\n");
  for (int inc = 1; inc <= d.increments; ++inc) {
    evp_step_report r{};
    EVP(evp_step(h, d.dt, &r));
    std::fprintf(fc, "%d %d %d %.6e %.6e", inc, r.iters, r.converged, r.err_stress, r.err_strain);
    for (double v : r.emacro) std::fprintf(fc, " %.12e", v);
    for (double v : r.savg) std::fprintf(fc, " %.12e", v);
    for (double v : r.epavg) std::fprintf(fc, " %.12e", v);
    std::fprintf(fc, " %.4f\n", r.seconds);
    std::fflush(fc);
    std::printf("increment %3d: %3d iterations%s  err %.2e %.2e  E33 %.5e  S33 %.6f  (%.3f s)\n", inc, r.iters, r.converged ? "" : " (NOT converged)",
                r.err_stress, r.err_strain, r.emacro[2], r.savg[2], r.seconds);
  }
  std::fclose(fc);
  for (auto &fo : d.field_out) {
    int id = -1, nc = 6;
    if (fo.first == "stress") id = EVP_FIELD_STRESS; else if (fo.first == "strain") id = EVP_FIELD_STRAIN;
    else if (fo.first == "plastic_strain") id = EVP_FIELD_PLASTIC_STRAIN; else if (fo.first == "rotation") { id = EVP_FIELD_ROTATION; nc = 9; }
    else { fail("unknown output field " + fo.first); return 1; }
    std::vector<double> buf((size_t)nc * N);
    EVP(evp_get_field(h, (evp_field)id, buf.data(), buf.size() * sizeof(double)));
    std::ofstream out(fo.second, std::ios::binary);
    out.write(reinterpret_cast<const char *>(buf.data()), (std::streamsize)(buf.size() * sizeof(double)));
    std::printf("wrote %s (%s, %d x %zu fp64, SoA [comp][z][y][x])\n", fo.second.c_str(), fo.first.c_str(), nc, N);
  }
  evp_destroy(h);
  return 0;
}
