// World-size-2 test of the sharded C ABI from a C++ caller (include/nmpc_b200/c_api.h, "several GPUs, one box"):
// a batch of cart-pole DDP solves sharded over two devices (GPU 0 and GPU 1; twice GPU 0 when the box has one) must give
// exactly what one handle on one GPU gives for the same instances, in the same order, for a ragged batch, through the
// host gather and through the direct device-to-device gather.  Prints "key value" lines for tests/test_sharding.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime_api.h>

#include <nmpc_b200/c_api.h>
#include <nmpc_b200/models/cartpole.h>
#include <nmpc_ddp/ShardedDDPSolver.h>

#define CHECK(call)                                                            \
  do                                                                           \
  {                                                                            \
    int rc_ = (call);                                                          \
    if(rc_ != 0)                                                               \
    {                                                                          \
      std::printf("error %d at %s: %s\n", rc_, #call, nmpc_b200_last_error()); \
      return 1;                                                                \
    }                                                                          \
  } while(0)

int main()
{
  const int n_dev = nmpc_b200_device_count();
  if(n_dev <= 0)
  {
    nmpc_b200_ddp_sharded * h = nullptr;
    double p[16] = {0};
    int rc = nmpc_b200_ddp_create_sharded("cartpole", p, 10, nullptr, 8, nullptr, 0, &h);
    std::printf("no_device_error %d %s\n", rc, nmpc_b200_last_error());
    return 0;
  }
  const int B = 37, N = 50, NX = 4, NU = 1; // 37 = 19 + 18: ragged
  int nx, nu, ng, np;
  CHECK(nmpc_b200_model_dims("cartpole", &nx, &nu, &ng, &np));
  std::vector<double> params(np);
  CHECK(nmpc_b200_model_default_params("cartpole", params.data()));
  nmpc_b200_ddp_config cfg;
  nmpc_b200_ddp_config_default(&cfg);
  cfg.horizon_steps = N;
  cfg.max_iter = 12;
  std::vector<double> x0(B * NX), u_init((size_t)B * N * NU, 0.0);
  for(int b = 0; b < B; b++)
  {
    x0[b * NX + 0] = 0.1 * std::sin(1.0 + b);
    x0[b * NX + 1] = 3.14159265358979323846 + 0.3 * std::cos(2.0 * b);
    x0[b * NX + 2] = 0.05 * b / B;
    x0[b * NX + 3] = -0.2 * std::sin(0.7 * b);
  }

  // one handle on device 0
  nmpc_b200_ddp * one = nullptr;
  CHECK(nmpc_b200_ddp_create("cartpole", params.data(), np, &cfg, B, 0, &one));
  CHECK(nmpc_b200_ddp_solve(one, B, 0.0, x0.data(), u_init.data(), N, 0, nullptr));
  std::vector<double> u_one((size_t)B * N * NU), cost_one(B), u0_one(B * NU);
  std::vector<int> iters_one(B);
  CHECK(nmpc_b200_ddp_get(one, NMPC_B200_DDP_U, u_one.data(), u_one.size() * 8, 0, nullptr));
  CHECK(nmpc_b200_ddp_get(one, NMPC_B200_DDP_COST, cost_one.data(), cost_one.size() * 8, 0, nullptr));
  CHECK(nmpc_b200_ddp_get(one, NMPC_B200_DDP_U0, u0_one.data(), u0_one.size() * 8, 0, nullptr));
  CHECK(nmpc_b200_ddp_get(one, NMPC_B200_DDP_ITERS, iters_one.data(), iters_one.size() * 4, 0, nullptr));

  // two shards
  const int devices[2] = {0, n_dev > 1 ? 1 : 0};
  nmpc_b200_ddp_sharded * sh = nullptr;
  CHECK(nmpc_b200_ddp_create_sharded("cartpole", params.data(), np, &cfg, B, devices, 2, &sh));
  std::printf("devices %d %d\n", devices[0], devices[1]);
  std::printf("num_shards %d\n", nmpc_b200_ddp_sharded_num_shards(sh));
  for(int s = 0; s < 2; s++)
  {
    int b, e, d;
    CHECK(nmpc_b200_ddp_sharded_range(sh, B, s, &b, &e, &d));
    std::printf("range%d %d %d %d\n", s, b, e, d);
  }
  CHECK(nmpc_b200_ddp_sharded_solve(sh, B, 0.0, x0.data(), u_init.data(), N));
  std::vector<double> u_sh((size_t)B * N * NU), cost_sh(B), u0_sh(B * NU);
  std::vector<int> iters_sh(B);
  CHECK(nmpc_b200_ddp_sharded_get(sh, NMPC_B200_DDP_U, u_sh.data(), u_sh.size() * 8, -1));
  CHECK(nmpc_b200_ddp_sharded_get(sh, NMPC_B200_DDP_COST, cost_sh.data(), cost_sh.size() * 8, -1));
  CHECK(nmpc_b200_ddp_sharded_get(sh, NMPC_B200_DDP_ITERS, iters_sh.data(), iters_sh.size() * 4, -1));
  // first-step controls of all shards gathered by direct stores into ONE buffer on device 0
  double * d_u0 = nullptr;
  cudaSetDevice(0);
  cudaMalloc(reinterpret_cast<void **>(&d_u0), sizeof(double) * B * NU);
  cudaMemset(d_u0, 0, sizeof(double) * B * NU);
  CHECK(nmpc_b200_ddp_sharded_get(sh, NMPC_B200_DDP_U0, d_u0, sizeof(double) * B * NU, 0));
  cudaMemcpy(u0_sh.data(), d_u0, sizeof(double) * B * NU, cudaMemcpyDeviceToHost);
  cudaFree(d_u0);

  double worst = 0;
  for(size_t i = 0; i < u_one.size(); i++) worst = std::fmax(worst, std::fabs(u_one[i] - u_sh[i]) / (1 + std::fabs(u_one[i])));
  std::printf("u_max_rel_diff %.3e\n", worst);
  std::printf("u_identical %d\n", (int)(std::memcmp(u_one.data(), u_sh.data(), u_one.size() * 8) == 0));
  std::printf("cost_identical %d\n", (int)(std::memcmp(cost_one.data(), cost_sh.data(), cost_one.size() * 8) == 0));
  std::printf("iters_identical %d\n", (int)(std::memcmp(iters_one.data(), iters_sh.data(), iters_one.size() * 4) == 0));
  std::printf("u0_peer_identical %d\n", (int)(std::memcmp(u0_one.data(), u0_sh.data(), u0_one.size() * 8) == 0));
  int max_it = 0;
  for(int b = 0; b < B; b++) max_it = iters_sh[b] > max_it ? iters_sh[b] : max_it;
  std::printf("max_iters %d\n", max_it);

  // a second, smaller solve on the same handle; errors
  CHECK(nmpc_b200_ddp_sharded_solve(sh, 3, 0.0, x0.data(), u_init.data(), N));
  std::vector<double> c3(3);
  CHECK(nmpc_b200_ddp_sharded_get(sh, NMPC_B200_DDP_COST, c3.data(), 24, -1));
  std::printf("small_identical %d\n", (int)(std::memcmp(c3.data(), cost_one.data(), 24) == 0));
  int rc = nmpc_b200_ddp_sharded_solve(sh, B + 1, 0.0, x0.data(), u_init.data(), N);
  std::printf("too_large %d\n", rc);
  rc = nmpc_b200_ddp_sharded_solve(sh, B, 0.0, x0.data(), u_init.data(), N - 1);
  std::printf("bad_horizon %d %s\n", rc, nmpc_b200_last_error());
  rc = nmpc_b200_ddp_sharded_get(sh, NMPC_B200_DDP_COST, c3.data(), 8, -1);
  std::printf("short_dst %d\n", rc);
  CHECK(nmpc_b200_ddp_sharded_destroy(sh));

  // the same through the C++ facade (include/nmpc_ddp/ShardedDDPSolver.h)
  {
    using Problem = nmpc_ddp::FunctorProblem<nmpc_b200::models::CartPole<double>>;
    auto problem = std::make_shared<Problem>("cartpole");
    nmpc_ddp::ShardedDDPSolver<4, 1> solver(problem, B, {devices[0], devices[1]});
    solver.config().horizon_steps = N;
    solver.config().max_iter = 12;
    const std::vector<bool> converged = solver.solveBatch(B, 0.0, x0.data(), u_init.data(), N);
    const auto u0 = solver.firstInputs();
    bool same = (int)u0.size() == B && solver.numShards() == 2;
    for(int b = 0; b < B && same; b++) same = u0[b][0] == u0_one[b];
    std::printf("facade_identical %d\n", (int)same);
    int n_conv = 0;
    for(int b = 0; b < B; b++) n_conv += converged[b] ? 1 : 0;
    std::printf("facade_converged %d\n", n_conv);
  }
  CHECK(nmpc_b200_ddp_destroy(one));
  std::printf("done 1\n");
  return 0;
}
