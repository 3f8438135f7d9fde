// warp_slots.cu — which hardware warp slot (%warpid; sub-partition = %warpid % 4) do the warps of co-resident CTAs get?
// Sizes the role layout of rollout_mma_split_kernel (six-warp CTAs, up to three per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_slots warp_slots.cu && ./warp_slots [threads] [ctas] [smem_kb]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void probe(int* out, long long spin) {
  extern __shared__ unsigned char sm[];
  unsigned smid, wid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  const long long t0 = clock64();
  while (clock64() - t0 < spin) {}  // keep every CTA resident until the whole grid has been placed
  if ((threadIdx.x & 31) == 0) {
    int* o = out + (blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 2;
    o[0] = (int)smid;
    o[1] = (int)wid;
  }
  if (sm[0] == 255 && spin < 0) out[0] = sm[1];
}

int main(int argc, char** argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 192, ctas = argc > 2 ? atoi(argv[2]) : 256, smem = (argc > 3 ? atoi(argv[3]) : 60) * 1024;
  const int wpc = threads / 32;
  int* d;
  cudaMalloc(&d, sizeof(int) * 2 * ctas * wpc);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<ctas, threads, smem>>>(d, 2000000);
  int* h = (int*)malloc(sizeof(int) * 2 * ctas * wpc);
  cudaMemcpy(h, d, sizeof(int) * 2 * ctas * wpc, cudaMemcpyDeviceToHost);
  if (cudaGetLastError() != cudaSuccess) return 1;
  // print the CTAs of the first three SMs that hold more than one CTA, and a histogram of (warp index in CTA -> sub-partition)
  int shown = 0;
  for (int sm = 0; sm < 148 && shown < 3; ++sm) {
    int n = 0;
    for (int c = 0; c < ctas; ++c) n += h[c * wpc * 2] == sm;
    if (n < 2) continue;
    ++shown;
    printf("{\"smid\": %d, \"ctas\": [", sm);
    bool first = true;
    for (int c = 0; c < ctas; ++c) {
      if (h[c * wpc * 2] != sm) continue;
      printf("%s{\"cta\": %d, \"warpid\": [", first ? "" : ", ", c);
      for (int w = 0; w < wpc; ++w) printf("%s%d", w ? ", " : "", h[(c * wpc + w) * 2 + 1]);
      printf("]}");
      first = false;
    }
    printf("]}\n");
  }
  long hist[32][4] = {};
  for (int c = 0; c < ctas; ++c)
    for (int w = 0; w < wpc; ++w) hist[w][h[(c * wpc + w) * 2 + 1] & 3]++;
  for (int w = 0; w < wpc; ++w)
    printf("{\"warp_in_cta\": %d, \"subpartition_hist\": [%ld, %ld, %ld, %ld]}\n", w, hist[w][0], hist[w][1], hist[w][2], hist[w][3]);
  return 0;
}
