// fuzz_main.cpp — TEST INFRASTRUCTURE ONLY.  Drives the host-compiled device functions (hostsim.cpp) with random and
// adversarial inputs under AddressSanitizer + UBSan: NaN / infinite / huge poses and scan points, particles outside
// the map, zero-length rays, degenerate quaternions, single-point maps.  On the GPU an out-of-bounds read in these
// functions would be silent; here it aborts.  Exit code 0 = no finding.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/mcl3dl_b200.h"

extern "C" int hostsim_measure(const mcl3dl_point* map, size_t n, const mcl3dl_lik_params* lp, const mcl3dl_beam_params* bp,
                               float cell_factor, const mcl3dl_pose* poses, size_t P, const mcl3dl_point* lik_pts,
                               size_t n_lik, const mcl3dl_point* beam_pts, size_t n_beam, const float* origins_xyz,
                               size_t n_origins, mcl3dl_result* out, uint8_t* status);

int main()
{
  std::mt19937 rng(2024);
  std::uniform_real_distribution<float> u(-1.f, 1.f);
  const float specials[] = {0.f, -0.f, 1e-30f, 1e30f, -1e30f, NAN, INFINITY, -INFINITY, 3.4e38f, 1e6f, -1e6f};
  auto weird = [&](float scale) { return (rng() % 9 == 0) ? specials[rng() % (sizeof(specials) / sizeof(float))] : u(rng) * scale; };
  long checks = 0;
  for (int trial = 0; trial < 300; ++trial)
  {
    const size_t n = trial % 7 == 0 ? 1 : 1 + rng() % 400;
    std::vector<mcl3dl_point> map(n);
    for (auto& p : map)
    {
      p.x = u(rng) * 3.f;
      p.y = u(rng) * 3.f;
      p.z = trial % 5 == 0 ? 0.f : u(rng);  // flat maps: one z layer
      p.label = rng() % 3;
    }
    mcl3dl_lik_params lp{5.0f, trial % 3 ? 0.2f : 0.45f, 0.05f, {1.f, 1.f, trial % 2 ? 5.f : 1.f}};
    mcl3dl_beam_params bp;
    std::memset(&bp, 0, sizeof(bp));
    bp.map_grid_size[0] = bp.map_grid_size[1] = bp.map_grid_size[2] = 0.1f;
    bp.dda_grid_size = trial % 4 ? 0.2f : 0.05f;
    bp.ray_angle_half = 0.25 * M_PI / 180.0;
    bp.hit_tolerance = 0.3f;
    bp.hit_range_sq = 0.09f;
    bp.sin_total_ref = 0.5f;
    bp.beam_likelihood = 0.9f;
    bp.beam_likelihood_min = 0.2f;
    bp.filter_label_max = trial % 3 ? 0xFFFFFFFFu : 1u;
    bp.add_penalty_short_only_mode = trial % 2;
    bp.use_raycast_using_dda = trial % 2;
    const size_t P = 1 + rng() % 6, nl = rng() % 12, nb = rng() % 12;
    std::vector<mcl3dl_pose> poses(P);
    for (auto& q : poses)
    {
      q.px = weird(4.f);
      q.py = weird(4.f);
      q.pz = weird(1.f);
      q._pad = 0;
      q.qx = weird(1.f);
      q.qy = weird(1.f);
      q.qz = weird(1.f);
      q.qw = (rng() % 11 == 0) ? 0.f : weird(1.f);
    }
    std::vector<mcl3dl_point> lik(nl), beam(nb);
    for (auto& p : lik) p = mcl3dl_point{weird(5.f), weird(5.f), weird(2.f), 0};
    for (auto& p : beam) p = mcl3dl_point{weird(5.f), weird(5.f), weird(2.f), static_cast<uint32_t>(rng() % 2)};
    if (nb && trial % 13 == 0) beam[0] = mcl3dl_point{0.f, 0.f, 0.f, 0};  // zero-length ray for origin 0 = (0,0,0)
    const float origins[6] = {0.f, 0.f, 0.f, weird(0.3f), weird(0.3f), weird(0.3f)};
    std::vector<mcl3dl_result> out(P);
    std::vector<uint8_t> status(P * (nb ? nb : 1));
    if (hostsim_measure(map.data(), n, &lp, &bp, 1.0f, poses.data(), P, lik.data(), nl, beam.data(), nb, origins, 2, out.data(),
                        status.data()) != 0)
    {
      std::fprintf(stderr, "hostsim_measure rejected trial %d\n", trial);
      return 2;
    }
    for (size_t i = 0; i < P; ++i)
    {
      if (out[i].n_short + out[i].n_hit + out[i].n_long > nb || out[i].match_cnt > nl)
      {
        std::fprintf(stderr, "impossible tallies in trial %d\n", trial);
        return 3;
      }
      ++checks;
    }
  }
  std::printf("fuzz ok: %ld particle records\n", checks);
  return 0;
}
