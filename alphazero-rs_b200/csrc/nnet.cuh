// The network behind NNet::predict (src/nnet.rs:40-44) — this build's own definition (the
// reference ships only a non-functional TF1 file):
//
//   in [B,2,6,7] -> conv3x3(2->C)+ReLU -> R x { conv3x3 C->C, ReLU, conv3x3 C->C, +skip, ReLU }
//   policy: conv1x1(C->2)+ReLU -> FC(84->7) -> softmax        value: conv1x1(C->1)+ReLU ->
//   FC(42->64)+ReLU -> FC(64->1) -> tanh.        C = 128, R = 6; BatchNorm folded into conv bias.
//
// Output contract = NNet::predict: probabilities pi[B,7] and value v[B].
// This file: parameter layout + the fp32 reference path (one CTA per position, CUDA cores) that
// pins the network's numerics (vs a torch fp32 restatement, 1e-4).  The bf16 tensor-core path
// (tcgen05/TMEM, csrc/nnet_tc.cuh) computes the tower; both share the head kernel.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace azb {

constexpr int kNetC = 128;     // tower width
constexpr int kCells = 42;

// Flat fp32 parameter vector, in this order (all row-major, the last index fastest):
//   stem_w [9][2][C]      (tap = (dy+1)*3 + (dx+1), input plane, out channel)   stem_b [C]
//   tower_w[2R][9][C][C]  (layer, tap, in channel, out channel)                 tower_b[2R][C]
//   pol_w  [C][2], pol_b[2], pol_fc_w[84][7] (input = plane*42 + cell), pol_fc_b[7]
//   val_w  [C],    val_b[1], val_fc1_w[42][64], val_fc1_b[64], val_fc2_w[64], val_fc2_b[1]
struct NetLayout {
  int R;
  size_t stem_w, stem_b, tower_w, tower_b, pol_w, pol_b, pol_fc_w, pol_fc_b, val_w, val_b, val_fc1_w,
      val_fc1_b, val_fc2_w, val_fc2_b, total;
};
inline NetLayout net_layout(int R) {
  NetLayout L{};
  L.R = R;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += n; return r; };
  L.stem_w = take(9 * 2 * kNetC);
  L.stem_b = take(kNetC);
  L.tower_w = take(static_cast<size_t>(2 * R) * 9 * kNetC * kNetC);
  L.tower_b = take(static_cast<size_t>(2 * R) * kNetC);
  L.pol_w = take(kNetC * 2);
  L.pol_b = take(2);
  L.pol_fc_w = take(84 * 7);
  L.pol_fc_b = take(7);
  L.val_w = take(kNetC);
  L.val_b = take(1);
  L.val_fc1_w = take(42 * 64);
  L.val_fc1_b = take(64);
  L.val_fc2_w = take(64);
  L.val_fc2_b = take(1);
  L.total = o;
  return L;
}

// Policy + value heads on the tower output of one position (act[cell][C] in shared memory).
// 128 threads.  Writes pi[0..6] (pi[7] = 0) and v.
template <int STRIDE>
__device__ __forceinline__ void heads_from_smem(const float* __restrict__ prm, const NetLayout& L,
                                                const float (*act)[STRIDE], float* scratch /* >= 256 floats */,
                                                float* pi_out, float* v_out) {
  const int tid = threadIdx.x;
  float* pol = scratch;        // [84] plane*42 + cell
  float* val = scratch + 84;   // [42]
  float* h1 = scratch + 128;   // [64]
  float* logit = scratch + 192;  // [7]
  for (int o = tid; o < 126; o += 128) {
    if (o < 84) {
      const int pl = o / 42, cell = o % 42;
      float s = prm[L.pol_b + pl];
      for (int ci = 0; ci < kNetC; ++ci) s = fmaf(act[cell][ci], prm[L.pol_w + ci * 2 + pl], s);
      pol[o] = fmaxf(s, 0.0f);
    } else {
      const int cell = o - 84;
      float s = prm[L.val_b];
      for (int ci = 0; ci < kNetC; ++ci) s = fmaf(act[cell][ci], prm[L.val_w + ci], s);
      val[cell] = fmaxf(s, 0.0f);
    }
  }
  __syncthreads();
  if (tid < 7) {
    float s = prm[L.pol_fc_b + tid];
    for (int i = 0; i < 84; ++i) s = fmaf(pol[i], prm[L.pol_fc_w + i * 7 + tid], s);
    logit[tid] = s;
  } else if (tid >= 64) {
    const int j = tid - 64;
    float s = prm[L.val_fc1_b + j];
    for (int i = 0; i < 42; ++i) s = fmaf(val[i], prm[L.val_fc1_w + i * 64 + j], s);
    h1[j] = fmaxf(s, 0.0f);
  }
  __syncthreads();
  if (tid == 0) {
    float m = logit[0];
    for (int a = 1; a < 7; ++a) m = fmaxf(m, logit[a]);
    float e[7], sum = 0.0f;
    for (int a = 0; a < 7; ++a) { e[a] = expf(logit[a] - m); sum += e[a]; }
    for (int a = 0; a < 7; ++a) pi_out[a] = e[a] / sum;
    pi_out[7] = 0.0f;
  } else if (tid == 32) {
    float s = prm[L.val_fc2_b];
    for (int j = 0; j < 64; ++j) s = fmaf(h1[j], prm[L.val_fc2_w + j], s);
    *v_out = tanhf(s);
  }
}

// 3x3 "same" convolution of one position held in shared memory, thread = output channel.
// Accumulation order: bias, then tap-major / input-channel-minor.
template <int CIN>
__device__ __forceinline__ void conv3x3_smem(const float* __restrict__ w /* [9][CIN][C] */,
                                             const float* __restrict__ b, const float (*in)[kNetC],
                                             float acc[kCells]) {
  const int co = threadIdx.x;
  const float bias = b[co];
#pragma unroll
  for (int c = 0; c < kCells; ++c) acc[c] = bias;
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    for (int ci = 0; ci < CIN; ++ci) {
      const float wv = w[(tap * CIN + ci) * kNetC + co];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const int rr = r + dy;
        if (rr < 0 || rr >= 6) continue;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          const int cc = c + dx;
          if (cc < 0 || cc >= 7) continue;
          acc[r * 7 + c] = fmaf(in[rr * 7 + cc][ci], wv, acc[r * 7 + c]);
        }
      }
    }
  }
}

// fp32 reference forward: one CTA (128 threads) per position.  states[i] = {cur.lo, cur.hi, opp.lo,
// opp.hi}: plane 0 = stones of the side to move, plane 1 = the other side, bit = row*7 + col.
__global__ void __launch_bounds__(128)
k_nnet_fp32(const float* __restrict__ prm, NetLayout L, const uint4* __restrict__ states,
            const uint32_t* __restrict__ count, uint32_t max_batch, float* __restrict__ pi_out,
            float* __restrict__ v_out) {
  extern __shared__ float smem[];
  float (*a0)[kNetC] = reinterpret_cast<float (*)[kNetC]>(smem);
  float (*a1)[kNetC] = reinterpret_cast<float (*)[kNetC]>(smem + kCells * kNetC);
  float* scratch = smem + 2 * kCells * kNetC;
  const uint32_t n = count ? min(*count, max_batch) : max_batch;
  const int tid = threadIdx.x;
  float acc[kCells];
  for (uint32_t pos = blockIdx.x; pos < n; pos += gridDim.x) {
    const uint4 st = states[pos];
    const uint64_t cur = (static_cast<uint64_t>(st.y) << 32) | st.x, opp = (static_cast<uint64_t>(st.w) << 32) | st.z;
    __syncthreads();
    for (int i = tid; i < kCells * 2; i += 128) {
      const int cell = i >> 1, pl = i & 1;
      a0[cell][pl] = (((pl ? opp : cur) >> cell) & 1ull) ? 1.0f : 0.0f;
    }
    __syncthreads();
    conv3x3_smem<2>(prm + L.stem_w, prm + L.stem_b, a0, acc);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kCells; ++c) a0[c][tid] = fmaxf(acc[c], 0.0f);
    __syncthreads();
    for (int blk = 0; blk < L.R; ++blk) {
      const float* w1 = prm + L.tower_w + static_cast<size_t>(2 * blk) * 9 * kNetC * kNetC;
      const float* w2 = w1 + 9 * kNetC * kNetC;
      conv3x3_smem<kNetC>(w1, prm + L.tower_b + (2 * blk) * kNetC, a0, acc);
#pragma unroll
      for (int c = 0; c < kCells; ++c) a1[c][tid] = fmaxf(acc[c], 0.0f);
      __syncthreads();
      conv3x3_smem<kNetC>(w2, prm + L.tower_b + (2 * blk + 1) * kNetC, a1, acc);
      __syncthreads();
#pragma unroll
      for (int c = 0; c < kCells; ++c) a0[c][tid] = fmaxf(acc[c] + a0[c][tid], 0.0f);
      __syncthreads();
    }
    heads_from_smem<kNetC>(prm, L, a0, scratch, pi_out + static_cast<size_t>(pos) * 8u, v_out + pos);
  }
}

// features f32 [B,2,6,7] (0/1 planes, Game::to_features) -> bitboard pairs
__global__ void k_features_to_bb(const float* __restrict__ feat, uint32_t n, uint4* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t cur = 0, opp = 0;
  for (int c = 0; c < kCells; ++c) {
    if (feat[static_cast<size_t>(i) * 84 + c] > 0.5f) cur |= 1ull << c;
    if (feat[static_cast<size_t>(i) * 84 + 42 + c] > 0.5f) opp |= 1ull << c;
  }
  out[i] = make_uint4(static_cast<uint32_t>(cur), static_cast<uint32_t>(cur >> 32), static_cast<uint32_t>(opp),
                      static_cast<uint32_t>(opp >> 32));
}

}  // namespace azb
