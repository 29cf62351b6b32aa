// controllers.hpp -- ROS-free C++ host classes that keep the class and parameter surface of the reference nodes
// and run the optimisation step on the B200 through the C ABI (include/mppi_b200.h).  No CPU fallback: every
// compute call goes to libmppi_b200.so and throws mppi::Error when that fails.
//
// Mirrors (paths relative to /root/reference):
//   class DiffDriveMPPI          include/ccv_mppi_path_tracker/diff_drive_mppi.h:52,          src/diff_drive_mppi.cpp
//   class SteeringDiffDriveMPPI  include/ccv_mppi_path_tracker/steering_diff_drive_mppi.h:56, src/steering_diff_drive_mppi.cpp
//   class FullBodyMPPI           include/ccv_mppi_path_tracker/full_body_mppi.h:68,           src/full_body_mppi.cpp
// Same member names for the parameters (horizon_, num_samples_, control_noise_, lambda_, v_max_, ... path_weight_,
// v_weight_, ...), same defaults (constructor `nh_.param` calls), same per-cycle method names.  ROS I/O becomes
// plain calls: pathCallback(xy) takes the path, set_pose()/set_state() stand for get_Transform()/get_CurrentState(),
// cmd_vel()/cmd_pos() return what publish_CmdVel()/publish_CmdPos() would publish.
//
// The reference's cycle is `sampling(); predict_States(); calc_Weights(); determine_OptimalSolution();` over shared
// member arrays (diff_drive_mppi.cpp:352-358).  Here the four names remain callable in that order:
//   sampling()                    stages pose / window / warm start and starts the H2D copy      (mppi_upload)
//   predict_States(), calc_Weights()  enqueue the fused kernels once (the first of the two calls) (mppi_enqueue)
//   determine_OptimalSolution()   copies the new optimal_solution back                             (mppi_download)
// solve() does all of it in one ABI call (mppi_solve).
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/mppi_b200.h"

namespace mppi {

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error("mppi error " + std::to_string(c) + ": " + m), code(c) {}
};

// optimal_solution of the reference (RobotStates, controls only): [T-1][U] row-major
struct ControlSequence {
  int horizon = 0, U = 0;
  std::vector<double> u;
  double &at(int t, int k) { return u[(size_t)t * U + k]; }
  double at(int t, int k) const { return u[(size_t)t * U + k]; }
};

struct CmdVel { double linear_x = 0, angular_z = 0; };                          // geometry_msgs/Twist fields used
struct CmdPos { double steer_l = 0, steer_r = 0, fore = 0, rear = 0, roll = 0; };  // ccv_dynamixel_msgs/CmdPoseByRadian

// ---- path sources (reference_path_creator.cpp:37-56, data/data.csv rows `x,y,`) --------------------------------
inline std::vector<double> make_sin_path(double course_length = 10.0, double resolution = 0.1, double A1 = 0.0,
                                         double omega1 = 0.0, double delta1 = 1.57, double A2 = 0.0, double omega2 = 0.0,
                                         double delta2 = 1.57, double A3 = 0.0, double omega3 = 0.0, double delta3 = 1.57,
                                         double init_x = 0.0, double init_y = 0.0) {
  std::vector<double> xy;
  for (double s = 0.0; s < course_length; s += resolution) {
    double y = A1 * cos(2 * M_PI * omega1 * s + delta1) + A2 * cos(2 * M_PI * omega2 * s + delta2) +
               A3 * cos(2 * M_PI * omega3 * s + delta3) + init_y;
    y -= A1 + A2 + A3;
    xy.push_back(init_x + s);
    xy.push_back(y);
  }
  return xy;
}
inline std::vector<double> load_path_csv(const std::string &file) {
  std::ifstream in(file);
  if (!in) throw Error(MPPI_ERR_INVALID, "cannot open " + file);
  std::vector<double> xy;
  std::string line;
  while (std::getline(in, line)) {
    std::stringstream ss(line);
    std::string a, b;
    if (std::getline(ss, a, ',') && std::getline(ss, b, ',') && !a.empty() && !b.empty()) {
      xy.push_back(atof(a.c_str()));
      xy.push_back(atof(b.c_str()));
    }
  }
  return xy;
}

class MPPIBase {
 public:
  virtual ~MPPIBase() { if (h_) mppi_destroy(h_); }
  MPPIBase(const MPPIBase &) = delete;
  MPPIBase &operator=(const MPPIBase &) = delete;

  // parameters common to the three nodes, reference member names (diff_drive_mppi.h:87-96)
  int horizon_ = 15;
  double num_samples_ = 1000.0;  // a double in DD/SD (diff_drive_mppi.h:88), int in FB
  double dt_ = 0.1, control_noise_ = 0.5, exploration_noise_ = 0.5, lambda_ = 1.0;
  double v_max_ = 1.2, w_max_ = 2.0, v_min_ = -1.2, w_min_ = -2.0;
  double v_ref_ = 0.8, resolution_ = 0.1, path_weight_ = 1.0, v_weight_ = 1.0, pitch_offset_ = 0.0;
  double tread_ = 0.501, wheel_radius_ = 0.1435;
  ControlSequence optimal_solution;

  // nh_.param(name, member, default) replacement: set a parameter by its ROS name before init()
  virtual bool set_param(const std::string &name, double v) {
    static const std::map<std::string, double MPPIBase::*> tab = {
        {"dt", &MPPIBase::dt_}, {"num_samples", &MPPIBase::num_samples_}, {"control_noise", &MPPIBase::control_noise_},
        {"lambda", &MPPIBase::lambda_}, {"v_max", &MPPIBase::v_max_}, {"w_max", &MPPIBase::w_max_},
        {"v_min", &MPPIBase::v_min_}, {"w_min", &MPPIBase::w_min_}, {"pitch_offset", &MPPIBase::pitch_offset_},
        {"v_ref", &MPPIBase::v_ref_}, {"resolution", &MPPIBase::resolution_},
        {"exploration_noise", &MPPIBase::exploration_noise_}, {"path_weight", &MPPIBase::path_weight_}};
    if (name == "horizon") { horizon_ = (int)v; return true; }
    auto it = tab.find(name);
    if (it == tab.end()) return false;
    this->*(it->second) = v;
    return true;
  }
  // `key=value` lines (a launch file's <param> list flattened); unknown keys are ignored like unread ROS params
  void load_params(const std::string &file) {
    std::ifstream in(file);
    std::string line;
    while (std::getline(in, line)) {
      size_t eq = line.find('=');
      if (eq == std::string::npos || line[0] == '#') continue;
      set_param(line.substr(0, eq), atof(line.substr(eq + 1).c_str()));
    }
  }

  // allocate the device side (the reference constructor's sample.resize / init, DD:36-46)
  void init(int device = 0, int n_robots = 1) {
    if (h_) throw Error(MPPI_ERR_STATE, "init() called twice");
    mppi_params p = abi_params();
    int rc = mppi_create(&h_, model_, &p, (int)num_samples_, horizon_, n_robots, device);
    if (rc != MPPI_OK) throw Error(rc, mppi_last_error(nullptr));
    resize_host();
  }
  // the host-side members alone (optimal_solution zeroed like RobotStates::init, current state): what cmd_vel(),
  // cmd_pos() and optimal_path() work on.  init() calls it; host-only tools call it instead of init().
  void resize_host() {
    optimal_solution.horizon = horizon_;
    optimal_solution.U = U_;
    optimal_solution.u.assign((size_t)(horizon_ - 1) * U_, 0.0);
    state_.assign(S_, 0.0);
  }
  // the mppi_params the class hands to the C ABI (parity checks)
  mppi_params params_for_abi() const { return abi_params(); }
  void update_params() { mppi_params p = abi_params(); check(mppi_set_params(h_, &p)); }

  // pathCallback (DD:48-52)
  void pathCallback(const std::vector<double> &path_xy) {
    check(mppi_set_path(h_, 0, path_xy.data(), (int)(path_xy.size() / 2)));
    path_received_ = true;
  }
  void set_seed(uint64_t seed, uint64_t counter = 0) { check(mppi_set_seed(h_, seed, counter)); }
  void use_graph(bool on) { check(mppi_use_graph(h_, on ? 1 : 0)); }

  // one control cycle in one ABI call
  const ControlSequence &solve() {
    check(mppi_solve(h_, state_.data(), dt_, optimal_solution.u.data()));
    return optimal_solution;
  }
  // the reference's four calls (DD:352-358)
  void sampling() { check(mppi_upload(h_, state_.data(), dt_, optimal_solution.u.data())); enqueued_ = false; }
  void predict_States() { enqueue_once(); }
  void calc_Weights() { enqueue_once(); }
  void determine_OptimalSolution() { enqueue_once(); check(mppi_download(h_, optimal_solution.u.data())); }

  CmdVel cmd_vel() const { return CmdVel{optimal_solution.at(0, 0), optimal_solution.at(0, 1)}; }  // DD:248-253
  // publish_OptimalPath (DD:295-312): optimal_solution re-rolled from the current pose with predict_NextState, in
  // double on the host like the reference; returns T x {x, y, yaw}
  std::vector<double> optimal_path() const {
    std::vector<double> st((size_t)horizon_ * 3, 0.0);
    st[0] = state_[0]; st[1] = state_[1]; st[2] = state_[2];
    for (int t = 0; t + 1 < horizon_; ++t) {
      const double heading = U_ == 2 ? st[3 * t + 2] : st[3 * t + 2] + optimal_solution.at(t, 2);
      st[3 * (t + 1)] = st[3 * t] + optimal_solution.at(t, 0) * cos(heading) * dt_;
      st[3 * (t + 1) + 1] = st[3 * t + 1] + optimal_solution.at(t, 0) * sin(heading) * dt_;
      st[3 * (t + 1) + 2] = st[3 * t + 2] + optimal_solution.at(t, 1) * dt_;
    }
    return st;
  }
  virtual CmdPos cmd_pos() const = 0;

  std::vector<float> costs() const {
    std::vector<float> c((size_t)num_samples_);
    check(mppi_get_costs(h_, 0, c.data()));
    return c;
  }
  void stats(double out[3]) const { check(mppi_get_stats(h_, 0, out)); }
  mppi_handle handle() const { return h_; }
  int num_controls() const { return U_; }
  int num_states() const { return S_; }
  bool path_received_ = false;

 protected:
  MPPIBase(int model, int U, int S) : model_(model), U_(U), S_(S) {}
  virtual mppi_params abi_params() const = 0;
  mppi_params common_params() const {
    mppi_params p;
    memset(&p, 0, sizeof p);
    p.control_noise = control_noise_;
    p.lambda = lambda_;
    p.v_ref = v_ref_;
    p.resolution = resolution_;
    p.u_min[0] = v_min_; p.u_max[0] = v_max_;
    p.u_min[1] = w_min_; p.u_max[1] = w_max_;
    p.path_weight = path_weight_;
    p.v_weight = v_weight_;
    return p;
  }
  void check(int rc) const { if (rc != MPPI_OK) throw Error(rc, mppi_last_error(h_)); }
  void enqueue_once() { if (!enqueued_) { check(mppi_enqueue(h_)); enqueued_ = true; } }
  // steering angles of the inner / outer wheel (SD:273-296, FB:246-262); R = |v/w| (inf when w = 0, as the reference)
  void steer_in_out(double v, double w, double delta, double &steer_l, double &steer_r) const {
    const double R = fabs(v / w);
    const double in = atan2(R * sin(delta), R * cos(delta) - tread_ / 2.0);
    const double out = atan2(R * sin(delta), R * cos(delta) + tread_ / 2.0);
    if (w > 0.0) { steer_l = in; steer_r = out; } else { steer_l = out; steer_r = in; }
  }
  int model_, U_, S_;
  mppi_handle h_ = nullptr;
  std::vector<double> state_;
  bool enqueued_ = false;
};

// ---- diff_drive_mppi: unicycle, controls (v, w) ---------------------------------------------------------------
class DiffDriveMPPI : public MPPIBase {
 public:
  DiffDriveMPPI() : MPPIBase(MPPI_MODEL_DIFF_DRIVE, 2, 3) { pitch_offset_ = 3.0 * M_PI / 180.0; }  // DD:17-34
  bool set_param(const std::string &name, double v) override {
    if (name == "control_weight") { v_weight_ = v; return true; }  // DD:34 reads "control_weight"; "v_weight" is ignored
    return MPPIBase::set_param(name, v);
  }
  // get_Transform (DD:314-330): pose of base_link in odom
  void set_pose(double x, double y, double yaw) { state_[0] = x; state_[1] = y; state_[2] = yaw; }
  CmdPos cmd_pos() const override { return CmdPos{0.0, 0.0, pitch_offset_, pitch_offset_, 0.0}; }  // DD:255-263
 protected:
  mppi_params abi_params() const override { return common_params(); }
};

// ---- steering_diff_drive_mppi: controls (v, w, steer) ---------------------------------------------------------
class SteeringDiffDriveMPPI : public MPPIBase {
 public:
  double steer_max_ = 30.0 * M_PI / 180.0, steer_min_ = -30.0 * M_PI / 180.0;
  SteeringDiffDriveMPPI() : MPPIBase(MPPI_MODEL_STEERING, 3, 3) {  // SD:18-36
    num_samples_ = 10000.0;
    w_max_ = 1.0;
    w_min_ = -1.0;
    exploration_noise_ = 0.1;
    pitch_offset_ = 3.0 * M_PI / 180.0;
  }
  bool set_param(const std::string &name, double v) override {
    if (name == "control_weight") { v_weight_ = v; return true; }
    if (name == "steer_max") { steer_max_ = v; return true; }
    if (name == "steer_min") { steer_min_ = v; return true; }
    return MPPIBase::set_param(name, v);
  }
  void set_pose(double x, double y, double yaw) { state_[0] = x; state_[1] = y; state_[2] = yaw; }
  CmdPos cmd_pos() const override {  // SD:273-296
    CmdPos c;
    steer_in_out(optimal_solution.at(0, 0), optimal_solution.at(0, 1), optimal_solution.at(0, 2), c.steer_l, c.steer_r);
    c.fore = c.rear = pitch_offset_;
    c.roll = 0.0;
    return c;
  }
 protected:
  mppi_params abi_params() const override {
    mppi_params p = common_params();
    p.u_min[2] = steer_min_;
    p.u_max[2] = steer_max_;
    return p;
  }
};

// ---- full_body_mppi: controls (v, w, direction, roll_v, pitch_v), states + roll, pitch -------------------------
class FullBodyMPPI : public MPPIBase {
 public:
  double steer_max_ = 30.0 * M_PI / 180.0, steer_min_ = -30.0 * M_PI / 180.0;
  double roll_max_ = 30.0 * M_PI / 180.0, roll_min_ = -30.0 * M_PI / 180.0;
  double pitch_max_ = 15.0 * M_PI / 180.0, pitch_min_ = -15.0 * M_PI / 180.0;
  double roll_v_max_ = 30.0 * M_PI / 180.0, roll_v_min_ = -30.0 * M_PI / 180.0;
  double pitch_v_max_ = 15.0 * M_PI / 180.0, pitch_v_min_ = -15.0 * M_PI / 180.0;
  double zmp_weight_ = 1.0, roll_v_weight_ = 1.0, back_weight_ = 1.0, yaw_weight_ = 1.0;
  bool roll_off_ = false, steer_off_ = false, use_gazebo_pose_ = true;
  FullBodyMPPI() : MPPIBase(MPPI_MODEL_FULL_BODY, 5, 5) {  // FB:8-46
    num_samples_ = 10000;
    w_max_ = 1.0;
    w_min_ = -1.0;
    v_min_ = -3.0;
    v_ref_ = 1.2;
    exploration_noise_ = 0.1;
    pitch_offset_ = 0.0;
  }
  bool set_param(const std::string &name, double v) override {
    static const std::map<std::string, double FullBodyMPPI::*> tab = {
        {"steer_max", &FullBodyMPPI::steer_max_}, {"steer_min", &FullBodyMPPI::steer_min_},
        {"roll_max", &FullBodyMPPI::roll_max_}, {"roll_min", &FullBodyMPPI::roll_min_},
        {"pitch_max", &FullBodyMPPI::pitch_max_}, {"pitch_min", &FullBodyMPPI::pitch_min_},
        {"roll_v_max", &FullBodyMPPI::roll_v_max_}, {"roll_v_min", &FullBodyMPPI::roll_v_min_},
        {"pitch_v_max", &FullBodyMPPI::pitch_v_max_}, {"pitch_v_min", &FullBodyMPPI::pitch_v_min_},
        {"zmp_weight", &FullBodyMPPI::zmp_weight_}, {"roll_v_weight", &FullBodyMPPI::roll_v_weight_},
        {"back_weight", &FullBodyMPPI::back_weight_}, {"yaw_weight", &FullBodyMPPI::yaw_weight_}};
    if (name == "v_weight") { v_weight_ = v; return true; }  // FB:35 reads "v_weight"
    if (name == "roll_off") { roll_off_ = v != 0.0; return true; }
    if (name == "steer_off") { steer_off_ = v != 0.0; return true; }
    if (name == "use_gazebo_pose") { use_gazebo_pose_ = v != 0.0; return true; }
    auto it = tab.find(name);
    if (it != tab.end()) { this->*(it->second) = v; return true; }
    return MPPIBase::set_param(name, v);
  }
  // get_CurrentState (FB:528-566): pose + IMU roll / pitch
  void set_state(double x, double y, double yaw, double roll, double pitch) {
    state_[0] = x; state_[1] = y; state_[2] = yaw; state_[3] = roll; state_[4] = pitch;
  }

  // ---- the node's ZMP monitors (published on zmp_y / true_zmp, FB:628-633; not inputs of the solve) -------------
  // constants: FBh:213-216 body box + mass, FB:86-91 base2CoM = height / 2 and I_O, FBh:30 gravity_, FBh:218 alpha
  double mass = 60.0, upper_body_height = 0.8075, upper_body_depth = 0.208, upper_body_width = 0.208, alpha = 0.3;
  double zmp_x_ = 0.0, zmp_y_ = 0.0;        // current_state_.zmp_x_[0], zmp_y_[0] (low-passed model ZMP)
  double true_ZMP[3] = {0.0, 0.0, 0.0};     // force-sensor ZMP (low-passed)
  double last_HG[3] = {0.0, 0.0, 0.0};

  // computeZMPfromModel (FB:597-603)
  void computeZMPfromModel(const double CoM[3], const double accel[3], const double HGdot[3], double out[3]) const {
    const double g[3] = {0.0, 0.0, -9.8};
    double mg[3], ma[3], c1[3], c2[3], MO[3];
    for (int k = 0; k < 3; ++k) { mg[k] = mass * g[k]; ma[k] = mass * accel[k]; }
    cross(CoM, mg, c1);
    cross(CoM, ma, c2);
    for (int k = 0; k < 3; ++k) MO[k] = c1[k] - c2[k] - HGdot[k];
    const double z[3] = {0.0, 0.0, 1.0};
    double zx[3];
    cross(z, MO, zx);
    const double denom = mass * ((g[0] - accel[0]) * z[0] + (g[1] - accel[1]) * z[1] + (g[2] - accel[2]) * z[2]);
    for (int k = 0; k < 3; ++k) out[k] = zx[k] / denom;
  }
  // the ZMP part of get_CurrentState (FB:551-566): IMU attitude, base-frame acceleration and angular velocity in,
  // low-passed model ZMP out.  Call once per cycle after set_state().
  void update_model_zmp(double imu_roll, double imu_pitch, double accel_x, double accel_y, const double omega[3]) {
    const double b = upper_body_height / 2;
    const double I[3] = {(mass * (upper_body_width * upper_body_width + upper_body_height * upper_body_height)) / 12 + mass * b * b,
                         (mass * (upper_body_height * upper_body_height + upper_body_depth * upper_body_depth)) / 12 + mass * b * b,
                         (mass * (upper_body_depth * upper_body_depth + upper_body_width * upper_body_width)) / 12};
    const double CoM[3] = {b * sin(imu_pitch), -b * sin(imu_roll), b * cos(imu_pitch) * cos(imu_roll)};
    const double accel[3] = {accel_x, accel_y, 0.0};
    double HG[3], HGdot[3], Z[3];
    for (int k = 0; k < 3; ++k) {
      HG[k] = I[k] * omega[k];
      HGdot[k] = (HG[k] - last_HG[k]) / dt_;
      last_HG[k] = HG[k];
    }
    computeZMPfromModel(CoM, accel, HGdot, Z);
    zmp_x_ = alpha * Z[0] + (1 - alpha) * zmp_x_;
    zmp_y_ = alpha * Z[1] + (1 - alpha) * zmp_y_;
  }
  // calc_true_ZMP (FB:568-596): six contact forces (wheels l, r, casters fl, fr, bl, br) in the base frame
  void calc_true_ZMP(const double forces[6][3]) {
    static const double pos[6][3] = {{0.0, 0.225, 0.075},     {0.0, -0.225, 0.075},   {0.245, 0.167, -0.003},
                                     {0.245, -0.167, -0.004}, {-0.245, -0.167, -0.004}, {-0.245, 0.167, -0.003}};  // FB:58-64
    double sumF[3] = {0, 0, 0}, sumM[3] = {0, 0, 0};
    for (int i = 0; i < 6; ++i)
      if (forces[i][2] > 0.0) {
        double m[3];
        cross(pos[i], forces[i], m);
        for (int k = 0; k < 3; ++k) { sumF[k] += forces[i][k]; sumM[k] += m[k]; }
      }
    const double n[3] = {0.0, 0.0, 1.0};
    const double denom = sumF[0] * n[0] + sumF[1] * n[1] + sumF[2] * n[2];
    if (fabs(denom) < 1e-6) return;  // FB:588-592
    double num[3];
    cross(n, sumM, num);
    for (int k = 0; k < 3; ++k) true_ZMP[k] = alpha * (num[k] / denom) + (1 - alpha) * true_ZMP[k];
  }
  static void cross(const double a[3], const double b[3], double out[3]) {
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
  }
  CmdPos cmd_pos() const override {  // FB:246-275
    CmdPos c;
    if (!steer_off_) steer_in_out(optimal_solution.at(0, 0), optimal_solution.at(0, 1), optimal_solution.at(0, 2), c.steer_l, c.steer_r);
    c.roll = state_[3] + optimal_solution.at(0, 3) * dt_;
    if (c.roll > roll_max_) c.roll = roll_max_;
    else if (c.roll < roll_min_) c.roll = roll_min_;
    if (roll_off_) c.roll = 0.0;
    c.fore = c.rear = pitch_offset_;
    return c;
  }
 protected:
  mppi_params abi_params() const override {
    mppi_params p = common_params();
    p.u_min[2] = steer_min_; p.u_max[2] = steer_max_;
    p.u_min[3] = roll_v_min_; p.u_max[3] = roll_v_max_;
    p.u_min[4] = pitch_v_min_; p.u_max[4] = pitch_v_max_;
    p.zmp_weight = roll_off_ ? 0.0 : zmp_weight_;        // FB:43-46
    p.roll_v_weight = roll_off_ ? 0.0 : roll_v_weight_;
    p.back_weight = back_weight_;
    p.yaw_weight = yaw_weight_;
    p.steer_off = steer_off_ ? 1 : 0;
    return p;
  }
};

}  // namespace mppi
