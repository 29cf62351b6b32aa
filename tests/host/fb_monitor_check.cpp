// Host-only check program for the C++ FullBodyMPPI ZMP monitors (csrc/host/controllers.hpp): reads the cycles of
// tests/golden/fb_estimator.dat (raw: int32 n, n x 30 doubles), writes n x 5 doubles {zmp_x, zmp_y, true_ZMP xyz}.
// No GPU: the controller object is never init()-ed.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../ccv_mppi_path_tracker_b200/csrc/host/controllers.hpp"

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 1;
  int32_t n = 0;
  if (fread(&n, 4, 1, f) != 1) return 1;
  std::vector<double> in((size_t)n * 30);
  if (fread(in.data(), 8, in.size(), f) != in.size()) return 1;
  fclose(f);
  mppi::FullBodyMPPI fb;
  FILE *o = fopen(argv[2], "wb");
  for (int c = 0; c < n; ++c) {
    const double *v = in.data() + (size_t)c * 30;
    fb.dt_ = v[0];
    double forces[6][3];
    for (int k = 0; k < 6; ++k)
      for (int j = 0; j < 3; ++j) forces[k][j] = v[8 + 3 * k + j];
    fb.calc_true_ZMP(forces);
    fb.update_model_zmp(v[1], v[2], v[3], v[4], v + 5);
    double out[5] = {fb.zmp_x_, fb.zmp_y_, fb.true_ZMP[0], fb.true_ZMP[1], fb.true_ZMP[2]};
    fwrite(out, 8, 5, o);
  }
  fclose(o);
  return 0;
}
