// Host-only check program for cmd_vel() / cmd_pos() of the C++ controller classes (csrc/host/controllers.hpp): reads
// the cases of tests/golden/cmd_golden.dat (raw: int32 n, n x 8 doubles {v0, w0, steer0 / direction0, roll_v0, roll of
// the current state, dt, steer_off, roll_off}) and writes, for the model given as argv[1] (dd | sd | fb), n x 7 doubles
// {cmd_vel.linear_x, angular_z, cmd_pos.steer_l, steer_r, fore, rear, roll} -- what the reference's publish_CmdVel() /
// publish_CmdPos() publish (DD:248-263, SD:266-296, FB:238-275).  No GPU: the controller is never init()-ed.
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../../ccv_mppi_path_tracker_b200/csrc/host/controllers.hpp"

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  const std::string model = argv[1];
  FILE *f = fopen(argv[2], "rb");
  if (!f) return 1;
  int32_t n = 0;
  if (fread(&n, 4, 1, f) != 1) return 1;
  std::vector<double> in((size_t)n * 8);
  if (fread(in.data(), 8, in.size(), f) != in.size()) return 1;
  fclose(f);
  mppi::DiffDriveMPPI dd;
  mppi::SteeringDiffDriveMPPI sd;
  mppi::FullBodyMPPI fb;
  mppi::MPPIBase *ctl = model == "dd" ? (mppi::MPPIBase *)&dd : (model == "sd" ? (mppi::MPPIBase *)&sd : (mppi::MPPIBase *)&fb);
  ctl->horizon_ = 3;
  ctl->resize_host();
  FILE *o = fopen(argv[3], "wb");
  for (int c = 0; c < n; ++c) {
    const double *v = in.data() + (size_t)c * 8;
    ctl->optimal_solution.at(0, 0) = v[0];
    ctl->optimal_solution.at(0, 1) = v[1];
    if (model != "dd") ctl->optimal_solution.at(0, 2) = v[2];
    ctl->dt_ = v[5];
    if (model == "fb") {
      fb.optimal_solution.at(0, 3) = v[3];
      fb.set_state(0.0, 0.0, 0.0, v[4], 0.0);
      fb.steer_off_ = v[6] != 0.0;
      fb.roll_off_ = v[7] != 0.0;
    }
    const mppi::CmdVel cv = ctl->cmd_vel();
    const mppi::CmdPos cp = ctl->cmd_pos();
    const double out[7] = {cv.linear_x, cv.angular_z, cp.steer_l, cp.steer_r, cp.fore, cp.rear, cp.roll};
    fwrite(out, 8, 7, o);
  }
  fclose(o);
  return 0;
}
