// Development tool: section cycle counts of the TF32x3 rollout kernel (-DTAMPC_PROFILE). Random weights.
#include <cstdio>
#include <vector>
#define TAMPC_PROFILE
#include "../tampc_b200/csrc/mma_rollout.cuh"
using namespace tampc;
using Model = ModelRexMma<PolicyTF32x3, 5, 3, 16, 32, 5, 32, 32, 5, 16, 32>;
int main(int argc, char** argv) {
  int K = argc > 1 ? atoi(argv[1]) : 8192, NM = argc > 2 ? atoi(argv[2]) : 1, block = argc > 3 ? atoi(argv[3]) : 32;
  const int T = 40;
  std::vector<unsigned char> w(Model::kBytes);
  for (size_t i = 0; i < w.size() / 4; ++i) ((float*)w.data())[i] = 0.01f * ((int)(i * 2654435761u % 200) - 100) / 100.f;
  void* dw; cudaMalloc(&dw, w.size() + 16384); cudaMemset(dw, 0, w.size() + 16384); cudaMemcpy(dw, w.data(), w.size(), cudaMemcpyHostToDevice);
  tampc_cost_desc c{}; c.kind = 0; c.nx = 5; c.nu = 3; c.ang_mask = 4; c.usim_mask = 5; c.use_trap_cost = 1; c.n_traps = 3; c.trap_weight = 1; c.has_terminal = 1; c.terminal_multiplier = 50;
  for (int i = 0; i < 5; ++i) { c.dist_scale[i] = i < 2 ? 1 : 0.1; c.Q[i * 8 + i] = 10; c.goal[i] = 0.5; for (int p = 0; p < 3; ++p) c.trap_x[p][i] = 0.3 * p; }
  for (int i = 0; i < 3; ++i) { c.R[i * 4 + i] = 0.01; for (int p = 0; p < 3; ++p) c.trap_u[p][i] = 0.5; }
  tampc_cost_desc* dc; cudaMalloc(&dc, sizeof c); cudaMemcpy(dc, &c, sizeof c, cudaMemcpyHostToDevice);
  SamplerD s{}; for (int i = 0; i < 3; ++i) { s.chol[i * 4 + i] = 0.5; s.lam_sigma_inv[i * 4 + i] = 0.02; s.u_min[i] = -1; s.u_max[i] = 1; } s.lambda_inv = 100; s.has_bounds = 1;
  SamplerD* ds; cudaMalloc(&ds, sizeof s); cudaMemcpy(ds, &s, sizeof s, cudaMemcpyHostToDevice);
  double* dU; cudaMalloc(&dU, T * 4 * 8); cudaMemset(dU, 0, T * 4 * 8);
  double hx[8] = {-0.4, 0.23, -0.6, 0, 0}; double* dx; cudaMalloc(&dx, 64); cudaMemcpy(dx, hx, 64, cudaMemcpyHostToDevice);
  double* dcost; cudaMalloc(&dcost, K * 8);
  RolloutArgs a{}; a.image = dw; /* prof tool: image = weights only; cost/sampler parts zero */ a.slope = 0.01; a.cost = dc; a.sampler = ds; a.U = dU; a.shift = 1; memcpy(a.state, hx, sizeof hx);
  a.src.mode = 0; a.src.seed = 1; a.src.call = 1; a.src.k_global = K; a.src.T = T; a.K_local = K; a.cost_out = dcost; a.wrap_mask = 4; a.has_guard = 1; a.bad_norm = 1e5;
  for (int rep = 0; rep < 2; ++rep) {
    unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_prof, z, sizeof z);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0);
    const bool split = argc > 4 && argv[4][0] == 's';
    const bool split3 = split && argv[4][1] == '3';
    cudaError_t e = split3 ? launch_rollout_mma_split<Model, 3>(a, 192, 0) : split ? launch_rollout_mma_split<Model, 2>(a, 128, 0) : (argc > 4 && argv[4][0] == 'p') ? launch_rollout_mma_pipe<Model>(a, 64, 0) : (argc > 4 ? (NM == 1 ? launch_rollout_mma<Model, float, float, 1, true>(a, block, 0) : launch_rollout_mma<Model, float, float, 2, false>(a, block, 0)) : (NM == 1 ? launch_rollout_mma<Model, double, float, 1, true>(a, block, 0) : launch_rollout_mma<Model, double, float, 2, false>(a, block, 0)));
    cudaEventRecord(e1); cudaDeviceSynchronize(); float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long g[16]; cudaMemcpyFromSymbol(g, g_prof, sizeof g);
    const char* names_1[] = {"sample_action", "guard", "stage_in", "enc", "dyn", "dec", "stage_out+dx", "wrap", "cost", "publish"};
    const char* names_s[] = {"A wait_c", "A load", "A enc", "A dyn", "A dx+publish", "B pert", "B dec", "B wait_x", "-", "B cost"};
    const char** names = split ? names_s : names_1;
    int SW = NM * 16; double warps = (double)((K + SW - 1) / SW);
    if (rep) { printf("K=%d NM=%d block=%d  %.1f us  err=%d\n", K, NM, block, ms * 1e3, (int)e); double tot = 0;
      for (int i = 0; i < 10; ++i) { double c = g[i] / warps / T; tot += c; printf("  %-14s %8.0f cycles/step/warp\n", names[i], c); } printf("  %-14s %8.0f\n", "sum", tot); }
  }
  return 0;
}
