// Developer tool (round 2): per-robot PGS sweep statistics of the bench workload on the HOST build of the kernel
// arithmetic, to size what warp divergence in the sweep loop costs and what grouping robots by class would save.
//   g++ -O2 -std=c++17 -DUPKIE_PGS_STATS -o /tmp/pgs_stats tools/r02/pgs_stats.cpp && /tmp/pgs_stats 4096 200 > /tmp/pgs.bin
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>
static thread_local int g_last = 0;  // 0: contact phase skipped; s: six-row sweeps; 100 + s: ten-row sweeps
inline void upkie_pgs_stats(int s) { g_last = s; }
#include "../../upkie_b200/csrc/params.h"
using namespace upkie_b200;
static bool any_fn(bool p) { return p; }
extern "C" int upkie_b200_dummy;
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 4096, ticks = argc > 2 ? atoi(argv[2]) : 200;
  const int limits = argc > 3 ? atoi(argv[3]) : 3;
  // model + config from a dump written by python (tools/r02/pgs_stats.py)
  UpkieModel m; UpkieSimConfig c;
  FILE* f = fopen("/tmp/pgs_model.bin", "rb");
  if (!f || fread(&m, sizeof(m), 1, f) != 1 || fread(&c, sizeof(c), 1, f) != 1) { fprintf(stderr, "no /tmp/pgs_model.bin\n"); return 1; }
  fclose(f);
  c.joint_limits = limits;
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
  std::vector<uint8_t> rec(size_t(n) * 5);
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
        g_last = 0;
        if (sub < nsub) servo_substep(P, s, a, resetting, &eps[size_t(i) * 6], mu[i], any_fn, NoSync(), nullptr, sub, nullptr, P.joint_limits);
        else g_last = 255;  // lane idle (reset tick)
        rec[size_t(i) * 5 + sub] = uint8_t(g_last);
      }
      observe_update(P, s);
      bool term = (fabsf(base_pitch(s)) > P.fall_pitch) || (s.pos[2] < P.min_base_height);
      if (resetting) { reset_wrapper_state(s); term = false; }
      done[i] = term ? 1 : 0;
    }
    fwrite(rec.data(), 1, rec.size(), stdout);
  }
  return 0;
}
