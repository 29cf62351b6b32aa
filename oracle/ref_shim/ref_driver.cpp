// ref_driver.cpp -- runs the UNMODIFIED reference node translation unit for one MPPI cycle, without ROS.
// TEST INFRASTRUCTURE ONLY.  Built by oracle/ref_shim/Makefile into oracle/_ref/ref_{dd,sd,fb}; never shipped.
//
// The reference source file is #included where it lies under /root/reference (REF_SRC), compiled against the stub
// headers in oracle/ref_shim/include, with `main` renamed and `private` opened so that this driver can
//   * fill path_, the current pose / state, dt_ and the warm start optimal_solution,
//   * replace ONLY the random draw of sampling() (std::mt19937 seeded from random_device is not reproducible):
//     sample = eps * control_noise_ + mean, i.e. what std::normal_distribution(mean, sigma) returns for the
//     standard normal eps, followed by the reference's own clamp() and steer_off rule,
//   * call the reference's predict_States(), calc_Weights(), determine_OptimalSolution() and calc_Cost() as is.
// The reference reads/writes one element past its control vectors (diff_drive_mppi.cpp:204, :230-235): a global
// operator new with zeroed padding makes that read return 0.0 deterministically instead of heap garbage.
//
// I/O: argv[1] = input file, argv[2] = output file (raw little-endian, layout in oracle/ref_runner.py).
// Timing mode (bench.py --impl reference / cpu_baseline): `--time <input file> <n_solves>` runs the reference's OWN
// cycle body -- sampling(); predict_States(); calc_Weights(); determine_OptimalSolution(); (diff_drive_mppi.cpp:
// 352-358), its own std::mt19937 draw included -- n_solves times on the calling thread (the node is single
// threaded, diff_drive_mppi.cpp:336-368) and prints the seconds of each solve.  Built with -O2 as *_time.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <new>
#include <queue>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

void *operator new(size_t n) {
  void *p = calloc(1, n + 64);
  if (!p) throw std::bad_alloc();
  return p;
}
void *operator new[](size_t n) { return operator new(n); }
void operator delete(void *p) noexcept { free(p); }
void operator delete[](void *p) noexcept { free(p); }
void operator delete(void *p, size_t) noexcept { free(p); }
void operator delete[](void *p, size_t) noexcept { free(p); }

#define private public
#define main ref_node_main
#include REF_SRC
#undef main
#undef private

#if REF_NODE == 1
typedef DiffDriveMPPI Node;
static const int U = 2, S = 3;
#elif REF_NODE == 2
typedef SteeringDiffDriveMPPI Node;
static const int U = 3, S = 3;
#else
typedef FullBodyMPPI Node;
static const int U = 5, S = 5;
#endif

static std::vector<double> *control(RobotStates &r, int u) {
#if REF_NODE == 1
  return u == 0 ? &r.v_ : &r.w_;
#elif REF_NODE == 2
  return u == 0 ? &r.v_ : (u == 1 ? &r.w_ : &r.steer_);
#else
  switch (u) {
    case 0: return &r.v_;
    case 1: return &r.w_;
    case 2: return &r.direction_;
    case 3: return &r.roll_v_;
    default: return &r.pitch_v_;
  }
#endif
}

template <class T>
static void rd(FILE *f, T *dst, size_t n) {
  if (fread(dst, sizeof(T), n, f) != n) {
    fprintf(stderr, "ref_driver: short read\n");
    exit(2);
  }
}

#if REF_NODE == 3
// estimator mode: argv = --estimator in out.  in: int32 n_cycles, then per cycle 30 doubles
// {dt, imu_roll, imu_pitch, accel_x, accel_y, omega[3], forces[6][3], pose x, y, yaw, pad}; out: per cycle 5 doubles
// {zmp_x, zmp_y, true_ZMP x, y, z} after the reference's own calc_true_ZMP() and get_CurrentState() (FB:528-596).
static int run_estimator(const char *fin, const char *fout) {
  FILE *f = fopen(fin, "rb");
  if (!f) return 1;
  int32_t n;
  rd(f, &n, 1);
  std::vector<double> in((size_t)n * 30);
  rd(f, in.data(), in.size());
  fclose(f);
  std::cout.setstate(std::ios_base::failbit);
  ref_shim::param_table()["horizon"] = 3;
  ref_shim::param_table()["num_samples"] = 1;
  Node node;
  static const char *topics[6] = {"/left_force_sensor/raw",       "/right_force_sensor/raw",     "/front_left_force_sensor/raw",
                                  "/front_right_force_sensor/raw", "/back_left_force_sensor/raw", "/back_right_force_sensor/raw"};
  FILE *o = fopen(fout, "wb");
  for (int c = 0; c < n; ++c) {
    const double *v = in.data() + (size_t)c * 30;
    node.dt_ = v[0];
    node.imu_roll_ = v[1];
    node.imu_pitch_ = v[2];
    node.accel_x = v[3];
    node.accel_y = v[4];
    node.filterd_imu_.angular_velocity.x = v[5];
    node.filterd_imu_.angular_velocity.y = v[6];
    node.filterd_imu_.angular_velocity.z = v[7];
    for (int k = 0; k < 6; ++k) {
      node.force_sensor_data_[topics[k]].wrench.force.x = v[8 + 3 * k];
      node.force_sensor_data_[topics[k]].wrench.force.y = v[9 + 3 * k];
      node.force_sensor_data_[topics[k]].wrench.force.z = v[10 + 3 * k];
    }
    node.use_gazebo_pose_ = true;
    node.gazebo_pose_.pose.position.x = v[26];
    node.gazebo_pose_.pose.position.y = v[27];
    node.gazebo_pose_.pose.orientation = tf::createQuaternionMsgFromYaw(v[28]);
    node.calc_true_ZMP();
    node.get_CurrentState();
    double out[5] = {node.current_state_.zmp_x_[0], node.current_state_.zmp_y_[0], node.true_ZMP.x(), node.true_ZMP.y(),
                     node.true_ZMP.z()};
    fwrite(out, 8, 5, o);
  }
  fclose(o);
  return 0;
}
#endif

// command mode: argv = --cmd in out.  in: int32 n, then per case 8 doubles {v0, w0, steer0 (SD) / direction0 (FB),
// roll_v0, roll of the current state, dt, steer_off, roll_off}; out: per case 7 doubles {cmd_vel.linear.x,
// cmd_vel.angular.z, cmd_pos.steer_l, steer_r, fore, rear, roll} as the reference's own publish_CmdVel() and
// publish_CmdPos() leave them in the message members (DD:248-263, SD:266-296, FB:238-275).
static int run_cmd(const char *fin, const char *fout) {
  FILE *f = fopen(fin, "rb");
  if (!f) return 1;
  int32_t n;
  rd(f, &n, 1);
  std::vector<double> in((size_t)n * 8);
  rd(f, in.data(), in.size());
  fclose(f);
  std::cout.setstate(std::ios_base::failbit);
  ref_shim::param_table()["horizon"] = 3;
  ref_shim::param_table()["num_samples"] = 1;
  Node node;
  FILE *o = fopen(fout, "wb");
  for (int c = 0; c < n; ++c) {
    const double *v = in.data() + (size_t)c * 8;
#if REF_NODE == 3
    RobotStates &opt = node.optimal_solution_;
    opt.direction_[0] = v[2];
    opt.roll_v_[0] = v[3];
    node.current_state_.roll_[0] = v[4];
    node.steer_off_ = v[6] != 0.0;
    node.roll_off_ = v[7] != 0.0;
#else
    RobotStates &opt = node.optimal_solution;
#if REF_NODE == 2
    opt.steer_[0] = v[2];
#endif
#endif
    opt.v_[0] = v[0];
    opt.w_[0] = v[1];
    node.dt_ = v[5];
    node.publish_CmdVel();
    node.publish_CmdPos();
    double out[7] = {node.cmd_vel_.linear.x, node.cmd_vel_.angular.z, node.cmd_pos_.steer_l, node.cmd_pos_.steer_r,
                     node.cmd_pos_.fore,     node.cmd_pos_.rear,      node.cmd_pos_.roll};
    fwrite(out, 8, 7, o);
  }
  fclose(o);
  return 0;
}

#include <chrono>

int main(int argc, char **argv) {
  if (argc >= 4 && std::string(argv[1]) == "--cmd") return run_cmd(argv[2], argv[3]);
#if REF_NODE == 3
  if (argc >= 4 && std::string(argv[1]) == "--estimator") return run_estimator(argv[2], argv[3]);
#endif
  int time_solves = 0;
  if (argc >= 4 && std::string(argv[1]) == "--time") {
    time_solves = atoi(argv[3]);
    argv[1] = argv[2];
    if (time_solves < 1) return 1;
  }
  if (argc < 3) return 1;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 1;
  int32_t hdr[4];
  rd(f, hdr, 4);
  const int K = hdr[0], T = hdr[1], n_path = hdr[2], n_params = hdr[3];
  for (int k = 0; k < n_params; ++k) {
    char name[32];
    double v;
    rd(f, name, 32);
    rd(f, &v, 1);
    name[31] = 0;
    ref_shim::param_table()[name] = v;
  }
  double state[5], dt;
  rd(f, state, 5);
  rd(f, &dt, 1);
  std::vector<double> path(2 * (size_t)n_path), u0((size_t)(T - 1) * U);
  std::vector<float> eps((size_t)(T - 1) * K * U);
  rd(f, path.data(), path.size());
  rd(f, u0.data(), u0.size());
  rd(f, eps.data(), eps.size());
  fclose(f);
  std::cout.setstate(std::ios_base::failbit);  // the nodes print every cycle

  Node node;  // constructor reads the parameters (incl. horizon, num_samples) from the table
  node.dt_ = dt;
  const bool timing = time_solves > 0;
  node.path_.poses.resize(n_path);
  for (int k = 0; k < n_path; ++k) {
    node.path_.poses[k].pose.position.x = path[2 * k];
    node.path_.poses[k].pose.position.y = path[2 * k + 1];
  }
  double yaw_used = state[2];
#if REF_NODE == 3
  node.current_state_.x_[0] = state[0];
  node.current_state_.y_[0] = state[1];
  node.current_state_.yaw_[0] = state[2];
  node.current_state_.roll_[0] = state[3];
  node.current_state_.pitch_[0] = state[4];
  RobotStates &opt = node.optimal_solution_;
#else
  node.current_pose_.pose.position.x = state[0];
  node.current_pose_.pose.position.y = state[1];
  node.current_pose_.pose.orientation = tf::createQuaternionMsgFromYaw(state[2]);
  yaw_used = tf::getYaw(node.current_pose_.pose.orientation);  // what predict_States() will read
  RobotStates &opt = node.optimal_solution;
#endif
  for (int t = 0; t < T - 1; ++t)
    for (int u = 0; u < U; ++u) (*control(opt, u))[t] = u0[(size_t)t * U + u];

  if (timing) {  // the reference's own cycle body, untouched (its own RNG: results are not reproducible)
    for (int n = 0; n < time_solves; ++n) {
      const auto t0 = std::chrono::steady_clock::now();
      node.sampling();
      node.predict_States();
      node.calc_Weights();
      node.determine_OptimalSolution();
      const auto t1 = std::chrono::steady_clock::now();
      printf("%.9f\n", std::chrono::duration<double>(t1 - t0).count());
      fflush(stdout);
    }
    return 0;
  }

  // sampling() with the supplied standard normals: same loop order, the reference's clamp()
  const double lo[5] = {node.v_min_, node.w_min_,
#if REF_NODE == 1
                        0, 0, 0};
  const double hi[5] = {node.v_max_, node.w_max_, 0, 0, 0};
#elif REF_NODE == 2
                        node.steer_min_, 0, 0};
  const double hi[5] = {node.v_max_, node.w_max_, node.steer_max_, 0, 0};
#else
                        node.steer_min_, node.roll_v_min_, node.pitch_v_min_};
  const double hi[5] = {node.v_max_, node.w_max_, node.steer_max_, node.roll_v_max_, node.pitch_v_max_};
#endif
  for (int t = 0; t < T - 1; ++t)
    for (int i = 0; i < K; ++i) {
      for (int u = 0; u < U; ++u) {
        std::vector<double> &dst = *control(node.sample[i], u);
        dst[t] = (double)eps[((size_t)t * K + i) * U + u] * node.control_noise_ + (*control(opt, u))[t];
      }
      for (int u = 0; u < U; ++u) node.clamp((*control(node.sample[i], u))[t], lo[u], hi[u]);
#if REF_NODE == 3
      if (node.steer_off_) node.sample[i].direction_[t] = 0.0;
#endif
    }

  node.predict_States();
  node.calc_Weights();
  std::vector<double> cost(K);
  for (int i = 0; i < K; ++i) cost[i] = node.calc_Cost(node.sample[i]);
  node.determine_OptimalSolution();

  FILE *o = fopen(argv[2], "wb");
  if (!o) return 1;
  fwrite(&yaw_used, 8, 1, o);
  int32_t cur = node.current_index_;
  for (int t = 0; t < T; ++t) {
    double w[3] = {node.x_ref_[t], node.y_ref_[t], node.yaw_ref_[t]};
    fwrite(w, 8, 3, o);
  }
  fwrite(cost.data(), 8, K, o);
  fwrite(node.weights_.data(), 8, K, o);
  for (int t = 0; t < T - 1; ++t)
    for (int u = 0; u < U; ++u) fwrite(&(*control(opt, u))[t], 8, 1, o);
  for (int i = 0; i < K; ++i)
    for (int t = 0; t < T; ++t) {
      double s[5] = {node.sample[i].x_[t], node.sample[i].y_[t], node.sample[i].yaw_[t], 0, 0};
#if REF_NODE == 3
      s[3] = node.sample[i].roll_[t];
      s[4] = node.sample[i].pitch_[t];
#endif
      fwrite(s, 8, S, o);
    }
#if REF_NODE == 3
  for (int i = 0; i < K; ++i)
    for (int t = 0; t < T - 2; ++t) {
      double z[2] = {node.sample[i].zmp_x_[t], node.sample[i].zmp_y_[t]};
      fwrite(z, 8, 2, o);
    }
#endif
  fwrite(&cur, 4, 1, o);
  fclose(o);
  return 0;
}
