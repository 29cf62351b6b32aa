// mppi_harness.cpp -- ROS-free C++ harness around the controller classes: closed-loop path tracking on the
// kinematic model (the role Gazebo plays for the reference nodes), one MPPI solve per 10 Hz tick on the B200.
//
//   mppi_harness --model dd|sd|fb [--K n] [--T n] [--cycles n] [--path file.csv | --sin L A1 omega1] [--launch]
//                [--param name=value ...] [--seed s] [--graph] [--split] [--quiet]
//                [--state x,y,yaw[,roll,pitch]] [--noise eps.f32] [--u0 u.f64] [--dump out.bin]
// Prints one line per cycle (pose, cmd_vel, cmd_pos) and a JSON summary with the solve latency percentiles and
// the cross-track RMSE against the path (the metric of the reference's src/calc_e_rmse.py:30-49).
// Parity mode (tests/test_gpu_parity.py): --noise feeds a standard-normal tensor [T-1][K][U] (float32, the reference's
// draw order) instead of the internal generator, --u0 a warm start [T-1][U] (float64), --state the initial pose;
// --dump writes, after the FIRST cycle, the mppi_params the class handed to the C ABI (raw struct), the new
// optimal_solution [T-1][U] (float64) and the per-sample costs [K] (float32).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <memory>

#include "controllers.hpp"

using namespace mppi;

static double nearest_path_distance(const std::vector<double> &xy, double x, double y) {
  double best = 1e300;
  for (size_t k = 0; k + 1 < xy.size(); k += 2) best = std::min(best, hypot(x - xy[k], y - xy[k + 1]));
  return best;
}

int main(int argc, char **argv) {
  std::string model = "dd", path_file, noise_file, u0_file, dump_file, state_arg;
  int K = -1, T = -1, cycles = 50;
  bool launch = false, graph = false, split = false, quiet = false;
  double L = 10.0, A1 = 1.0, om1 = 0.25;
  uint64_t seed = 0x5EED0000ull;
  std::vector<std::pair<std::string, double>> overrides;
  for (int a = 1; a < argc; ++a) {
    std::string s = argv[a];
    auto next = [&]() { return a + 1 < argc ? argv[++a] : (char *)"0"; };
    if (s == "--model") model = next();
    else if (s == "--K") K = atoi(next());
    else if (s == "--T") T = atoi(next());
    else if (s == "--cycles") cycles = atoi(next());
    else if (s == "--path") path_file = next();
    else if (s == "--sin") { L = atof(next()); A1 = atof(next()); om1 = atof(next()); }
    else if (s == "--launch") launch = true;
    else if (s == "--graph") graph = true;
    else if (s == "--split") split = true;
    else if (s == "--quiet") quiet = true;
    else if (s == "--seed") seed = strtoull(next(), nullptr, 0);
    else if (s == "--noise") noise_file = next();
    else if (s == "--u0") u0_file = next();
    else if (s == "--dump") dump_file = next();
    else if (s == "--state") state_arg = next();
    else if (s == "--param") {
      std::string kv = next();
      size_t eq = kv.find('=');
      if (eq != std::string::npos) overrides.emplace_back(kv.substr(0, eq), atof(kv.substr(eq + 1).c_str()));
    } else { fprintf(stderr, "unknown argument %s\n", s.c_str()); return 2; }
  }
  try {
    std::unique_ptr<MPPIBase> ctl;
    DiffDriveMPPI *dd = nullptr;
    SteeringDiffDriveMPPI *sd = nullptr;
    FullBodyMPPI *fb = nullptr;
    if (model == "dd") ctl.reset(dd = new DiffDriveMPPI());
    else if (model == "sd") ctl.reset(sd = new SteeringDiffDriveMPPI());
    else if (model == "fb") ctl.reset(fb = new FullBodyMPPI());
    else { fprintf(stderr, "--model dd|sd|fb\n"); return 2; }
    if (launch) {  // launch/{diff_drive,steering_diff_drive,full_body}_mppi.launch
      ctl->set_param("v_max", 2.0);
      ctl->set_param("path_weight", 10.0);
      ctl->set_param("v_weight", 1.0);  // ignored by DD/SD exactly like the nodes ignore it
      ctl->set_param("v_ref", fb ? 2.0 : 1.2);
      if (sd) ctl->set_param("num_samples", 1000);
      if (fb) {
        ctl->set_param("zmp_weight", 10.0); ctl->set_param("roll_v_weight", 0.5); ctl->set_param("yaw_weight", 2.0);
        ctl->set_param("roll_off", 1); ctl->set_param("use_gazebo_pose", 0);
        L = 20.0; A1 = 1.5; om1 = 0.127;
      }
    }
    for (auto &kv : overrides)
      if (!ctl->set_param(kv.first, kv.second)) fprintf(stderr, "warning: unknown parameter %s\n", kv.first.c_str());
    if (K > 0) ctl->set_param("num_samples", K);
    if (T > 0) ctl->set_param("horizon", T);
    ctl->init(0);
    ctl->set_seed(seed);
    ctl->use_graph(graph);
    std::vector<double> path = path_file.empty() ? make_sin_path(L, 0.1, A1, om1, 0.0, 0, 0, 0.0, 0, 0, 0.0) : load_path_csv(path_file);
    if (path.empty()) throw Error(MPPI_ERR_INVALID, "empty path");
    ctl->pathCallback(path);

    double x = path[0], y = path[1], yaw = 0.0, roll = 0.0, pitch = 0.0;
    if (!state_arg.empty()) {
      double v[5] = {x, y, 0, 0, 0};
      int k = 0;
      std::stringstream ss(state_arg);
      std::string tok;
      while (k < 5 && std::getline(ss, tok, ',')) v[k++] = atof(tok.c_str());
      x = v[0]; y = v[1]; yaw = v[2]; roll = v[3]; pitch = v[4];
    }
    auto read_file = [](const std::string &name, size_t bytes, void *dst) {
      FILE *f = fopen(name.c_str(), "rb");
      if (!f || fread(dst, 1, bytes, f) != bytes) throw Error(MPPI_ERR_INVALID, "cannot read " + name);
      fclose(f);
    };
    const size_t n_u = (size_t)(ctl->horizon_ - 1) * ctl->num_controls();
    if (!noise_file.empty()) {
      std::vector<float> eps(n_u * (size_t)ctl->num_samples_);
      read_file(noise_file, eps.size() * sizeof(float), eps.data());
      int rc = mppi_set_noise(ctl->handle(), eps.data());
      if (rc != MPPI_OK) throw Error(rc, mppi_last_error(ctl->handle()));
    }
    if (!u0_file.empty()) read_file(u0_file, n_u * sizeof(double), ctl->optimal_solution.u.data());
    std::vector<double> lat_us;
    double se = 0.0, emax = 0.0;
    for (int c = 0; c < cycles; ++c) {
      if (dd) dd->set_pose(x, y, yaw);
      if (sd) sd->set_pose(x, y, yaw);
      if (fb) fb->set_state(x, y, yaw, roll, pitch);
      auto t0 = std::chrono::steady_clock::now();
      if (split) { ctl->sampling(); ctl->predict_States(); ctl->calc_Weights(); ctl->determine_OptimalSolution(); }
      else ctl->solve();
      auto t1 = std::chrono::steady_clock::now();
      lat_us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
      if (c == 0 && !dump_file.empty()) {
        FILE *o = fopen(dump_file.c_str(), "wb");
        if (!o) throw Error(MPPI_ERR_INVALID, "cannot write " + dump_file);
        const mppi_params p = ctl->params_for_abi();
        const std::vector<float> cost = ctl->costs();
        fwrite(&p, sizeof p, 1, o);
        fwrite(ctl->optimal_solution.u.data(), sizeof(double), ctl->optimal_solution.u.size(), o);
        fwrite(cost.data(), sizeof(float), cost.size(), o);
        fclose(o);
      }
      const CmdVel cv = ctl->cmd_vel();
      const CmdPos cp = ctl->cmd_pos();
      const ControlSequence &u = ctl->optimal_solution;
      // plant: the same kinematics the controller predicts with (predict_NextState), one dt
      const double heading = dd ? yaw : yaw + u.at(0, 2);
      x += cv.linear_x * cos(heading) * ctl->dt_;
      y += cv.linear_x * sin(heading) * ctl->dt_;
      yaw += cv.angular_z * ctl->dt_;
      if (fb) { roll += u.at(0, 3) * ctl->dt_; pitch += u.at(0, 4) * ctl->dt_; }
      const double e = nearest_path_distance(path, x, y);
      se += e * e;
      emax = std::max(emax, e);
      if (!quiet)
        printf("cycle %3d pose %.3f %.3f %.3f  cmd_vel %.4f %.4f  cmd_pos %.4f %.4f %.4f  e %.4f  solve %.1f us\n", c, x, y,
               yaw, cv.linear_x, cv.angular_z, cp.steer_l, cp.steer_r, cp.roll, e, lat_us.back());
    }
    std::vector<double> s = lat_us;
    std::sort(s.begin(), s.end());
    double st[3];
    ctl->stats(st);
    printf("{\"model\": \"%s\", \"K\": %d, \"T\": %d, \"cycles\": %d, \"solve_p50_us\": %.1f, \"solve_p99_us\": %.1f, "
           "\"rmse_m\": %.4f, \"max_error_m\": %.4f, \"final_x\": %.3f, \"ess\": %.2f}\n",
           model.c_str(), (int)ctl->num_samples_, ctl->horizon_, cycles, s[s.size() / 2], s[(s.size() * 99) / 100],
           sqrt(se / cycles), emax, x, st[2]);
  } catch (const Error &e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
