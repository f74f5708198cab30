// Developer tool (round 2): how often the body-ground contact gate fires on the bench workload (BASELINE configs[2]:
// random torques, termination at |pitch| > 1 or base below 0.15 m, next-step auto-reset), on the HOST build of the
// kernel arithmetic. Prints, per 50 ticks, the fraction of robot-substeps whose torso holds contact rows and the
// fraction of 32-robot groups ("warps") that would take general_contact_solve() in a substep.
//   python -c "import sys; sys.path.insert(0,'.'); from upkie_b200 import _abi; from upkie_b200.model import Model; import bench; \
//     open('/tmp/pgs_model.bin','wb').write(bytes(Model.standard_upkie().to_struct()) + bytes(bench.servos_config()))"
//   g++ -O2 -std=c++17 -o /tmp/body_gate_stats tools/r02/body_gate_stats.cpp && /tmp/body_gate_stats 4096 300
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "../../upkie_b200/csrc/params.h"
using namespace upkie_b200;
static bool any_fn(bool p) { return p; }
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 4096, ticks = argc > 2 ? atoi(argv[2]) : 300;
  UpkieModel m; UpkieSimConfig c;
  FILE* f = fopen("/tmp/pgs_model.bin", "rb");
  if (!f || fread(&m, sizeof(m), 1, f) != 1 || fread(&c, sizeof(c), 1, f) != 1) { fprintf(stderr, "no /tmp/pgs_model.bin\n"); return 1; }
  fclose(f);
  SimParams P; std::memset(&P, 0, sizeof(P)); std::string err;
  if (make_sim_params(m, c, P, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> U(-1.f, 1.f), U01(0.f, 1.f);
  std::vector<RobotState> S(n);
  std::vector<float> eps(size_t(n) * 6), mu(n);
  std::vector<uint32_t> episode(n, 0);
  std::vector<uint8_t> done(n, 0);
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 6; ++k) eps[size_t(i) * 6 + k] = 0.2f * U(rng);
    mu[i] = 0.5f + 0.7f * U01(rng);
    std::memset(&S[i], 0, sizeof(RobotState));
    float init[UPKIE_INIT_DIM];
    sample_init_state(P, 2025, i, ++episode[i], init);
    reset_robot(P, S[i], init, &eps[size_t(i) * 6], mu[i], any_fn, P.joint_limits);
  }
  std::vector<uint8_t> hit(size_t(n) * 5);
  long robots = 0, robots_hit = 0, warps = 0, warps_hit = 0;
  for (int t = 0; t < ticks; ++t) {
    for (int i = 0; i < n; ++i) {
      RobotState& s = S[i];
      float a[UPKIE_ACT_DIM];
      for (int j = 0; j < 6; ++j) {
        a[j * 6 + 0] = nanf(""); a[j * 6 + 1] = 0.f; a[j * 6 + 2] = U(rng) * P.tau_max[j];
        a[j * 6 + 3] = 0.f; a[j * 6 + 4] = 0.f; a[j * 6 + 5] = P.tau_max[j];
      }
      const bool resetting = done[i] != 0;
      int nsub = P.nb_substeps;
      if (resetting) {
        float init[UPKIE_INIT_DIM];
        sample_init_state(P, 2025, i, ++episode[i], init);
        reset_pose(s, init);
        nsub = 1;
      } else {
        clamp_servo_action(P, a);
      }
      for (int sub = 0; sub < 5; ++sub) {
        float rec[UPKIE_BODY_REC_DIM] = {0.f};
        if (sub < nsub) servo_substep(P, s, a, resetting, &eps[size_t(i) * 6], mu[i], any_fn, NoSync(), nullptr, sub, nullptr,
                                      P.joint_limits, BodyRecOut{rec, 1});
        hit[size_t(i) * 5 + sub] = rec[0] != 0.f;
      }
      observe_update(P, s);
      bool term = (fabsf(base_pitch(s)) > P.fall_pitch) || (s.pos[2] < P.min_base_height);
      if (resetting) { reset_wrapper_state(s); term = false; }
      done[i] = term ? 1 : 0;
    }
    for (int sub = 0; sub < 5; ++sub)
      for (int w = 0; w < n / 32; ++w) {
        int any = 0;
        for (int l = 0; l < 32; ++l) { any |= hit[size_t(w * 32 + l) * 5 + sub]; robots_hit += hit[size_t(w * 32 + l) * 5 + sub]; }
        robots += 32; warps += 1; warps_hit += any;
      }
    if ((t + 1) % 50 == 0) {
      printf("ticks %4d-%4d: robot-substeps with torso rows %.5f, warp-substeps taking the general solver %.5f\n", t - 48, t + 1,
             double(robots_hit) / robots, double(warps_hit) / warps);
      robots = robots_hit = warps = warps_hit = 0;
    }
  }
  return 0;
}
