// evpfft_microstructure — standalone microstructure tool, C++17 above the C ABI host helpers.
//
//   evpfft_microstructure voronoi NX NY NZ NGRAINS SEED OUT.txt   periodic Voronoi polycrystal (integer exact, one random
//                                                                 orientation per grain) as a per-voxel text file
//   evpfft_microstructure info FILE.txt NX NY NZ                  read a per-voxel text file, print grain / phase statistics
//
// File format: one line per voxel, "phi1 Phi phi2 i j k grain phase" (Bunge angles in degrees, 1-based indices, 1-based
// phase).  No GPU needed: only host-side entry points of libevpfft_b200.so are used.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "../../../include/evpfft.h"

int main(int argc, char **argv) {
  if (argc >= 8 && std::string(argv[1]) == "voronoi") {
    evp_grid g{atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), 1.0, 1.0, 1.0};
    const int ng = atoi(argv[5]);
    const unsigned long long seed = strtoull(argv[6], nullptr, 10);
    if (g.nx < 1 || g.ny < 1 || g.nz < 1 || ng < 1) { std::fprintf(stderr, "bad grid / grain count\n"); return 1; }
    const size_t N = (size_t)g.nx * g.ny * g.nz;
    std::vector<int32_t> grain(N);
    std::vector<double> grot((size_t)9 * ng), rot9(9 * N);
    if (evp_voronoi(&g, ng, seed, 0, g.nz, grain.data(), grot.data()) != 0) { std::fprintf(stderr, "evp_voronoi failed\n"); return 2; }
    for (size_t v = 0; v < N; ++v)
      for (int c = 0; c < 9; ++c) rot9[c * N + v] = grot[(size_t)9 * grain[v] + c];
    if (evp_write_microstructure_txt(argv[7], &g, grain.data(), nullptr, rot9.data()) != 0) { std::fprintf(stderr, "cannot write %s\n", argv[7]); return 2; }
    std::printf("wrote %s: %d x %d x %d voxels, %d grains, seed %llu\n", argv[7], g.nx, g.ny, g.nz, ng, seed);
    return 0;
  }
  if (argc >= 6 && std::string(argv[1]) == "info") {
    evp_grid g{atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), 1.0, 1.0, 1.0};
    const size_t N = (size_t)g.nx * g.ny * g.nz;
    std::vector<int32_t> grain(N), phase(N);
    std::vector<double> rot9(9 * N);
    if (evp_read_microstructure_txt(argv[2], &g, grain.data(), phase.data(), rot9.data()) != 0) { std::fprintf(stderr, "cannot read %s for this grid\n", argv[2]); return 2; }
    std::set<int32_t> gs(grain.begin(), grain.end()), ps(phase.begin(), phase.end());
    std::printf("%s: %zu voxels, %zu grains, %zu phases\n", argv[2], N, gs.size(), ps.size());
    return 0;
  }
  std::fprintf(stderr, "usage: evpfft_microstructure voronoi NX NY NZ NGRAINS SEED OUT.txt | info FILE.txt NX NY NZ\n");
  return 1;
}
