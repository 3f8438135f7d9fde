// mma_stream.cuh — 3xTF32 rollout for WIDE plain-MLP dynamics (hidden width 256: the top of the synthetic sweep,
// BASELINE.json configs[3] / SURVEY.md §8d config 4-5).
//
// (nx+nu) -> H -> H -> nx with H = 256 no longer fits the register-chained kernels of mma_rollout.cuh: the H x H layer
// alone is 512 KB of split weights (more than shared memory) and its 256-wide input AND output would need 256 registers
// per lane.  Here
//   * layer 1 and layer 3 (16 KB each) stay resident in shared memory,
//   * the H x H layer is STREAMED from L2 in 16 pieces of 32 KB (8 k-blocks x 8 n-tiles) through a double buffer filled
//     with cp.async one piece ahead; all 8 warps of the CTA (128 samples) consume the same piece, so every byte fetched
//     feeds 128 samples (weight traffic: 544 KB per CTA per step against ~28 us of tensor work),
//   * its output is produced in chunks of 8 n-tiles (32 accumulator registers) and each finished chunk is consumed at
//     once as 8 k-blocks of layer 3 — the 256-wide hidden output never exists in full, nothing is spilled.
// Accumulation orders equal mma_layer's (k-blocks in order, lo*hi + hi*lo + hi*hi per block; four interleaved
// accumulators for the one-tile output layer), so the arithmetic is the same 3xTF32 scheme as the narrower shapes.
#pragma once
#include "mma_rollout.cuh"

namespace tampc {

template <int NX_, int NU_, int H>
struct ModelMlpStream {
  static_assert(H % 64 == 0 && NX_ + NU_ <= 8 && NX_ <= 8, "hidden width in chunks of 8 n-tiles");
  using P = PolicyTF32x3;
  using T = float;
  using Net = Mlp3Mma<P, 1, H / 8, H / 8, 1>;  // packed layout shared with the resident-weight kernels
  static constexpr int NX = NX_, NU = NU_, H8 = H / 8;
  static constexpr int kBytes = Net::kBytes;
  static constexpr int kL1Bytes = Net::kOff1, kL3Bytes = Net::kBytes - Net::kOff2;
  static constexpr int kPieceBytes = 64 * P::kBlockBytes;  // 8 k-blocks x 8 n-tiles
  static constexpr int kPieces = (H8 / 8) * (H8 / 8);
  static constexpr int kInStride = 8, kOutStride = 10;
  static constexpr int kStageElems = 32 * kInStride + 32 * kOutStride;
  static constexpr int kMacs = (NX_ + NU_) * H + H * H + H * NX_;
};

template <class Model>
struct StreamSmem {
  static constexpr int WARPS = 8;
  // global image: packed networks | CostS | SamplerS<float> | SamplerS<double>   (mppi_api.cu set_model)
  static constexpr size_t gCost = align16<int>(Model::kBytes);
  static constexpr size_t gSampler = gCost + align16<int>(sizeof(CostS<float, float>));
  static constexpr size_t gEnd = gSampler + align16<int>(sizeof(SamplerS<float>));
  // shared: consts | L1 | L3 | Ush | per-warp stage | two weight pieces
  static constexpr size_t kCost = 0;
  static constexpr size_t kSampler = gSampler - gCost;
  static constexpr size_t kL1 = gEnd - gCost;
  static constexpr size_t kL3 = kL1 + align16<int>(Model::kL1Bytes);
  static constexpr size_t kU = kL3 + align16<int>(Model::kL3Bytes);
  static __host__ __device__ constexpr size_t stage_off(int T_) { return kU + align16<int>(sizeof(float) * (size_t)T_ * kMaxNu); }
  static __host__ __device__ constexpr size_t piece_off(int T_) { return (stage_off(T_) + WARPS * align16<int>(sizeof(float) * Model::kStageElems) + 127) & ~size_t(127); }
  static __host__ __device__ constexpr size_t bytes(int T_) { return piece_off(T_) + 2 * (size_t)Model::kPieceBytes; }
};

template <class Model>
__global__ void __launch_bounds__(256, 1) rollout_mma_stream_kernel(const RolloutArgs a) {
  using P = PolicyTF32x3;
  using L = StreamSmem<Model>;
  constexpr int NX = Model::NX, NU = Model::NU, H8 = Model::H8, SW = P::MT, NCH = H8 / 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  CostS<float, float>& cs = *reinterpret_cast<CostS<float, float>*>(smem_raw + L::kCost);
  SamplerS<float>& sp = *reinterpret_cast<SamplerS<float>*>(smem_raw + L::kSampler);
  const unsigned char* w1 = smem_raw + L::kL1;
  const unsigned char* w3 = smem_raw + L::kL3;
  float* Ush = reinterpret_cast<float*>(smem_raw + L::kU);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = a.src.T;
  float* stage = reinterpret_cast<float*>(smem_raw + L::stage_off(T) + (size_t)warp * align16<int>(sizeof(float) * Model::kStageElems));
  unsigned char* pieces = smem_raw + L::piece_off(T);
  const unsigned char* gimg = reinterpret_cast<const unsigned char*>(a.image);
  const unsigned char* g2 = gimg + Model::Net::kOff1;            // layer 2: bias[H] | blocks (kb * H8 + o)
  const float* bias2 = reinterpret_cast<const float*>(g2);
  const unsigned char* g2blocks = g2 + H8 * 8 * sizeof(float);

  // piece p = (oc, kg): k-blocks kg*8.., n-tiles oc*8..: eight 4 KB runs of the packed layer
  auto fetch_piece = [&](int p) {
    const int oc = p / NCH, kg = p % NCH;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(pieces + (size_t)(p & 1) * Model::kPieceBytes);
    for (int i = tid; i < Model::kPieceBytes / 16; i += 256) {
      const int n = i >> 8, off = (i & 255) * 16;  // 4096 / 16 = 256 requests per run
      const unsigned char* src = g2blocks + ((size_t)((kg * 8 + n) * H8 + oc * 8) * P::kBlockBytes) + off;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + n * 4096 + off), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  {
    copy_image_async(smem_raw + L::kCost, gimg + L::gCost, (int)(L::gEnd - L::gCost), tid, 256);
    copy_image_async(smem_raw + L::kL1, gimg, Model::kL1Bytes, tid, 256);
    copy_image_async(smem_raw + L::kL3, gimg + Model::Net::kOff2, Model::kL3Bytes, tid, 256);
    for (int i = tid; i < T * NU; i += 256) {
      int t = i / NU, d = i % NU, ts = t + a.shift;
      Ush[t * kMaxNu + d] = (float)(ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d]);
    }
    copy_image_wait();
  }
  __syncthreads();
  fetch_piece(0);

  const int row = lane % SW, part = lane / SW;
  const int kbase = (blockIdx.x * L::WARPS + warp) * SW;
  const bool live = kbase + row < a.K_local;
  const int k = min(kbase + row, a.K_local - 1);  // idle rows / warps recompute the last sample and never write
  float x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = (float)a.state[i];
  double cost = 0.0;
  float pert = 0.f;
  const float slope = (float)a.slope;
  const float bad2 = (float)(a.bad_norm * a.bad_norm);
  const int t4 = lane & 3;
  float* sin = stage;
  float* sout = stage + 32 * Model::kInStride;

#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    float u[NU], npost[NU];
    pert += sample_action<float, NU>(a.src, sp, Ush + t * kMaxNu, k, t, u, npost);
    bool bad = false;
    if (a.has_guard) {
      float n2 = 0.f;
#pragma unroll
      for (int i = 0; i < NX; ++i) n2 = __fmaf_rn(x[i], x[i], n2);
      bad = n2 > bad2;
      if (bad) {
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = 0.f;
#pragma unroll
        for (int i = 0; i < NU; ++i) u[i] = 0.f;
      }
    }
    if (lane == row) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sin[row * Model::kInStride + i] = i < NX ? x[i] : (i < NX + NU ? u[i - NX] : 0.f);
    }
    __syncwarp();
    float in[1][1][P::R];
    load_blocks<P, 1, 1, Model::kInStride>(sin, in, lane);
    // ---- layer 1 (resident): the whole 256-wide hidden vector of the tile in registers
    float h1[1][H8][P::R];
    mma_layer<P, 1, 1, H8, true>(in, h1, w1, slope, lane);
    // ---- layer 2 (streamed) fused with layer 3 (resident)
    const float* bias3 = reinterpret_cast<const float*>(w3);
    const unsigned char* b3 = w3 + 8 * sizeof(float);
    float acc3[4][P::R];
    {
      const float c0 = bias3[2 * t4], c1 = bias3[2 * t4 + 1];
      acc3[0][0] = c0; acc3[0][1] = c1; acc3[0][2] = c0; acc3[0][3] = c1;
#pragma unroll
      for (int s = 1; s < 4; ++s)
#pragma unroll
        for (int r = 0; r < P::R; ++r) acc3[s][r] = 0.f;
    }
#pragma unroll 1
    for (int oc = 0; oc < NCH; ++oc) {
      float acc[8][P::R];
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const float c0 = __ldg(bias2 + (oc * 8 + o) * 8 + 2 * t4), c1 = __ldg(bias2 + (oc * 8 + o) * 8 + 2 * t4 + 1);
        acc[o][0] = c0; acc[o][1] = c1; acc[o][2] = c0; acc[o][3] = c1;
      }
#pragma unroll
      for (int kg = 0; kg < NCH; ++kg) {
        const int p = oc * NCH + kg;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                       // piece p has landed for everyone; everyone is done with piece p-1
        fetch_piece((p + 1) % Model::kPieces);  // (the last one prefetches the next step's first piece)
        const unsigned char* wb = pieces + (size_t)(p & 1) * Model::kPieceBytes;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const float av[4] = {h1[0][kg * 8 + n][0], h1[0][kg * 8 + n][2], h1[0][kg * 8 + n][1], h1[0][kg * 8 + n][3]};
          uint32_t ahi[4], alo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ahi[i] = __float_as_uint(av[i]) & 0xffffe000u;
            alo[i] = __float_as_uint(av[i] - __uint_as_float(ahi[i]));
          }
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const uint4 wv = *reinterpret_cast<const uint4*>(wb + ((n * 8 + o) * 32 + lane) * 16);
            mma1688_tf32(acc[o], alo, wv.x, wv.y, acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
            mma1688_tf32(acc[o], ahi, wv.z, wv.w, acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
            mma1688_tf32(acc[o], ahi, wv.x, wv.y, acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
          }
        }
      }
      // chunk oc of the hidden output is complete: activation, then straight into layer 3 as k-blocks oc*8..oc*8+7
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        float hv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) hv[r] = fmaxf(acc[n][r], slope * acc[n][r]);
        const float av[4] = {hv[0], hv[2], hv[1], hv[3]};
        uint32_t ahi[4], alo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ahi[i] = __float_as_uint(av[i]) & 0xffffe000u;
          alo[i] = __float_as_uint(av[i] - __uint_as_float(ahi[i]));
        }
        const uint4 wv = *reinterpret_cast<const uint4*>(b3 + ((oc * 8 + n) * 32 + lane) * 16);
        // products q = 3 * kb + j are dealt over four accumulators (q % 4), as mma_layer does for a one-tile output;
        // 3 * (oc * 8 + n) % 4 depends on n only (24 % 4 == 0)
        constexpr int dummy = 0;
        (void)dummy;
        const int q0 = (3 * n) & 3;
        mma1688_tf32(acc3[q0], alo, wv.x, wv.y, acc3[q0][0], acc3[q0][1], acc3[q0][2], acc3[q0][3]);
        mma1688_tf32(acc3[(q0 + 1) & 3], ahi, wv.z, wv.w, acc3[(q0 + 1) & 3][0], acc3[(q0 + 1) & 3][1], acc3[(q0 + 1) & 3][2], acc3[(q0 + 1) & 3][3]);
        mma1688_tf32(acc3[(q0 + 2) & 3], ahi, wv.x, wv.y, acc3[(q0 + 2) & 3][0], acc3[(q0 + 2) & 3][1], acc3[(q0 + 2) & 3][2], acc3[(q0 + 2) & 3][3]);
      }
    }
    float dxf[1][1][P::R];
#pragma unroll
    for (int r = 0; r < P::R; ++r) dxf[0][0][r] = ((acc3[0][r] + acc3[1][r]) + acc3[2][r]) + acc3[3][r];
    store_blocks<P, 1, 1, Model::kOutStride>(sout, dxf, 0, lane);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] += sout[row * Model::kOutStride + i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      if ((a.wrap_mask >> i) & 1u) x[i] = angle_normalize<float>(x[i]);
      if (bad) x[i] = 0.f;
    }
    cost += (double)running_cost<float, float, NX, NU>(cs, x, u, part, 2);
    if (a.states_out != nullptr && live && part == 0) {
      double* so = a.states_out + ((size_t)k * T + t) * NX;
#pragma unroll
      for (int i = 0; i < NX; ++i) so[i] = (double)x[i];
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // the prefetch issued by the last piece
  const double mine = cost;
  cost += __shfl_sync(0xffffffffu, mine, (row + SW) & 31);
  double total = cost + (double)terminal_cost<float, float, NX>(cs, x);
  total += (double)pert;
  if (live && part == 0) a.cost_out[k] = total;
}

template <class Model>
cudaError_t launch_rollout_mma_stream(const RolloutArgs& a, int /*block*/, cudaStream_t stream) {
  auto kern = rollout_mma_stream_kernel<Model>;
  using L = StreamSmem<Model>;
  const size_t smem = L::bytes(a.src.T);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int per_block = L::WARPS * PolicyTF32x3::MT;
  const int grid = (a.K_local + per_block - 1) / per_block;
  kern<<<grid, 256, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace tampc
