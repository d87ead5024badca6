// The small kernels of the training step (NNet::train, src/nnet.rs:38; loss and optimiser as the reference's
// connect_four_net.py:102-112: softmax cross-entropy on pi + mean squared error on v, Adam) around the tensor-core
// convolution kernels of nnet_tc.cuh (forward, backward data, backward weights):
//   k_heads_backward  loss + gradients of the two heads, dL/d(tower output) gated by its ReLU
//   k_colsum_bf16     bias gradients (column sums of a pre-activation gradient)
//   k_stem_backward   gradients of the 2 -> 128 stem convolution from the binary input planes
//   k_adam            Adam on the fp32 master parameters
//   k_build_tiles / k_build_stem_table  the bf16 operand tiles (forward and backward) and the stem table from the
//                     updated parameters
// Mixed precision: fp32 master parameters and gradients, bf16 tower weights / activations / activation gradients
// with fp32 accumulation, fp32 heads.
#pragma once
#include "nnet_tc.cuh"

namespace azb {

// ---- heads: forward recomputed, loss, backward ------------------------------------------------------
// One CTA of 128 threads walks positions pos = blockIdx.x, + gridDim.x, ...; head-parameter gradients are summed in
// shared memory and added to `grad` (layout = the parameter vector) once per CTA.
constexpr int kHeadGradFloats = kNetC * 2 + 2 + 84 * 7 + 7 + kNetC + 1 + 42 * 64 + 64 + 64 + 1;  // pol_w .. val_fc2_b: contiguous in NetLayout
__global__ void __launch_bounds__(128)
k_heads_backward(const float* __restrict__ prm, NetLayout L, const __nv_bfloat16* __restrict__ act, uint32_t n_pos,
                 const float* __restrict__ pis, const float* __restrict__ vs, float inv_batch, ActLayout lay,
                 __nv_bfloat16* __restrict__ d_act, float* __restrict__ grad, float* __restrict__ loss /*[2]*/) {
  extern __shared__ float sm[];
  float (*a0)[kNetC + 1] = reinterpret_cast<float (*)[kNetC + 1]>(sm);  // [42][129]
  float* pol_pre = sm + kCells * (kNetC + 1);  // [84]
  float* val_pre = pol_pre + 84;               // [42]
  float* h1 = val_pre + 42;                    // [64] (post-ReLU)
  float* logit = h1 + 64;                      // [8]
  float* dlogit = logit + 8;                   // [8]
  float* dh1 = dlogit + 8;                     // [64]
  float* dpol_pre = dh1 + 64;                  // [84]
  float* dval_pre = dpol_pre + 84;             // [42]
  float* misc = dval_pre + 42;                 // [4]: ds, loss_pi, loss_v
  float* gacc = misc + 4;                      // [kHeadGradFloats], same order as the parameters from L.pol_w on
  const int tid = threadIdx.x;
  for (int i = tid; i < kHeadGradFloats; i += 128) gacc[i] = 0.0f;
  float loss_pi = 0.0f, loss_v = 0.0f;  // thread 0 / thread 32
  const size_t o_pol_w = 0, o_pol_b = L.pol_b - L.pol_w, o_pol_fc_w = L.pol_fc_w - L.pol_w, o_pol_fc_b = L.pol_fc_b - L.pol_w,
               o_val_w = L.val_w - L.pol_w, o_val_b = L.val_b - L.pol_w, o_fc1_w = L.val_fc1_w - L.pol_w,
               o_fc1_b = L.val_fc1_b - L.pol_w, o_fc2_w = L.val_fc2_w - L.pol_w, o_fc2_b = L.val_fc2_b - L.pol_w;
  for (uint32_t pos = blockIdx.x; pos < n_pos; pos += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < kCells * kNetC / 2; i += 128) {
      const int cell = (2 * i) / kNetC, c = (2 * i) % kNetC;
      const uint32_t wd = *reinterpret_cast<const uint32_t*>(act + lay.row(pos, cell / 7, cell % 7) * kNetC + c);
      a0[cell][c] = __uint_as_float(wd << 16);
      a0[cell][c + 1] = __uint_as_float(wd & 0xFFFF0000u);
    }
    __syncthreads();
    // forward, as heads_from_smem
    for (int o = tid; o < 126; o += 128) {
      if (o < 84) {
        const int pl = o / 42, cell = o % 42;
        float s = prm[L.pol_b + pl];
        for (int ci = 0; ci < kNetC; ++ci) s = fmaf(a0[cell][ci], prm[L.pol_w + ci * 2 + pl], s);
        pol_pre[o] = s;
      } else {
        const int cell = o - 84;
        float s = prm[L.val_b];
        for (int ci = 0; ci < kNetC; ++ci) s = fmaf(a0[cell][ci], prm[L.val_w + ci], s);
        val_pre[cell] = s;
      }
    }
    __syncthreads();
    if (tid < 7) {
      float s = prm[L.pol_fc_b + tid];
      for (int i = 0; i < 84; ++i) s = fmaf(fmaxf(pol_pre[i], 0.0f), prm[L.pol_fc_w + i * 7 + tid], s);
      logit[tid] = s;
    } else if (tid >= 64) {
      const int j = tid - 64;
      float s = prm[L.val_fc1_b + j];
      for (int i = 0; i < 42; ++i) s = fmaf(fmaxf(val_pre[i], 0.0f), prm[L.val_fc1_w + i * 64 + j], s);
      h1[j] = fmaxf(s, 0.0f);
    }
    __syncthreads();
    if (tid == 0) {  // softmax cross-entropy: -sum pi log p, d/dlogit = p * sum(pi) - pi
      float m = logit[0];
      for (int a = 1; a < 7; ++a) m = fmaxf(m, logit[a]);
      float e[7], sum = 0.0f, spi = 0.0f;
      for (int a = 0; a < 7; ++a) { e[a] = expf(logit[a] - m); sum += e[a]; }
      const float lse = m + logf(sum);
      for (int a = 0; a < 7; ++a) {
        const float t = pis[static_cast<size_t>(pos) * 7 + a];
        spi += t;
        loss_pi += t * (lse - logit[a]);
      }
      for (int a = 0; a < 7; ++a) dlogit[a] = (e[a] / sum * spi - pis[static_cast<size_t>(pos) * 7 + a]) * inv_batch;
    } else if (tid == 32) {  // (v - z)^2 with v = tanh(s)
      float s = prm[L.val_fc2_b];
      for (int j = 0; j < 64; ++j) s = fmaf(h1[j], prm[L.val_fc2_w + j], s);
      const float v = tanhf(s), z = vs[pos];
      loss_v += (v - z) * (v - z);
      misc[0] = 2.0f * (v - z) * (1.0f - v * v) * inv_batch;
    }
    __syncthreads();
    const float ds = misc[0];
    if (tid < 64) {
      gacc[o_fc2_w + tid] += ds * h1[tid];
      const float d = h1[tid] > 0.0f ? ds * prm[L.val_fc2_w + tid] : 0.0f;
      dh1[tid] = d;
      gacc[o_fc1_b + tid] += d;
    } else if (tid < 71) {
      gacc[o_pol_fc_b + (tid - 64)] += dlogit[tid - 64];
    } else if (tid == 71) {
      gacc[o_fc2_b] += ds;
    }
    __syncthreads();
    for (int o = tid; o < 126; o += 128) {
      if (o < 84) {
        float s = 0.0f;
        for (int j = 0; j < 7; ++j) s = fmaf(dlogit[j], prm[L.pol_fc_w + o * 7 + j], s);
        dpol_pre[o] = pol_pre[o] > 0.0f ? s : 0.0f;
      } else {
        const int i = o - 84;
        float s = 0.0f;
        for (int j = 0; j < 64; ++j) s = fmaf(dh1[j], prm[L.val_fc1_w + i * 64 + j], s);
        dval_pre[i] = val_pre[i] > 0.0f ? s : 0.0f;
      }
    }
    for (int i = tid; i < 42 * 64; i += 128) gacc[o_fc1_w + i] += fmaxf(val_pre[i / 64], 0.0f) * dh1[i % 64];
    for (int i = tid; i < 84 * 7; i += 128) gacc[o_pol_fc_w + i] += fmaxf(pol_pre[i / 7], 0.0f) * dlogit[i % 7];
    __syncthreads();
    {  // 1x1 convolutions: thread = channel
      const int c = tid;
      const float w0 = prm[L.pol_w + c * 2], w1 = prm[L.pol_w + c * 2 + 1], wv = prm[L.val_w + c];
      float g0 = 0.0f, g1 = 0.0f, gv = 0.0f;
      for (int cell = 0; cell < kCells; ++cell) {
        const float a = a0[cell][c], d0 = dpol_pre[cell], d1 = dpol_pre[42 + cell], dv = dval_pre[cell];
        g0 = fmaf(a, d0, g0);
        g1 = fmaf(a, d1, g1);
        gv = fmaf(a, dv, gv);
        const float da = a > 0.0f ? fmaf(d0, w0, fmaf(d1, w1, dv * wv)) : 0.0f;  // gated by the tower output's ReLU
        d_act[lay.row(pos, cell / 7, cell % 7) * kNetC + c] = __float2bfloat16_rn(da);
      }
      gacc[o_pol_w + c * 2] += g0;
      gacc[o_pol_w + c * 2 + 1] += g1;
      gacc[o_val_w + c] += gv;
      if (c < 3) {
        float s = 0.0f;
        for (int cell = 0; cell < kCells; ++cell) s += c == 0 ? dpol_pre[cell] : (c == 1 ? dpol_pre[42 + cell] : dval_pre[cell]);
        if (c < 2) gacc[o_pol_b + c] += s; else gacc[o_val_b] += s;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < kHeadGradFloats; i += 128) atomicAdd(grad + L.pol_w + i, gacc[i]);
  if (tid == 0) atomicAdd(loss + 0, loss_pi * inv_batch);
  if (tid == 32) atomicAdd(loss + 1, loss_v * inv_batch);
}
constexpr size_t kHeadsBwdSmem = (kCells * (kNetC + 1) + 84 + 42 + 64 + 8 + 8 + 64 + 84 + 42 + 4 + kHeadGradFloats) * sizeof(float);

// ---- bias gradient: out[c] += sum over rows of dz[row][c] (padding rows are zero) -------------------------------
__global__ void __launch_bounds__(256) k_colsum_bf16(const __nv_bfloat16* __restrict__ dz, uint32_t rows, float* __restrict__ out) {
  // 16 threads read one 256-byte row as 16-byte vectors, a CTA takes 16 rows per step; 8 fp32 partial sums per thread
  __shared__ float part[16][kNetC];
  const int cg = threadIdx.x & 15, rr = threadIdx.x >> 4;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (uint32_t r = blockIdx.x * 16u + rr; r < rows; r += gridDim.x * 16u) {
    const uint4 x = *reinterpret_cast<const uint4*>(dz + static_cast<size_t>(r) * kNetC + cg * 8);
    const uint32_t wd[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s[2 * k] += __uint_as_float(wd[k] << 16);
      s[2 * k + 1] += __uint_as_float(wd[k] & 0xFFFF0000u);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) part[rr][cg * 8 + k] = s[k];
  __syncthreads();
  if (threadIdx.x < kNetC) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += part[i][threadIdx.x];
    atomicAdd(out + threadIdx.x, t);
  }
}

// ---- stem: dW[tap][plane][c] += x_shift[row][plane] * dz[row][c], db[c] += dz[row][c]; thread = channel, 18 + 1
// register accumulators, the bit tests are warp-uniform.  512 threads = 4 position groups x 128 channels (a full SM of
// warps: the kernel is latency-bound on the dz loads, seven of them in flight per board row), the groups' sums are
// combined in shared memory and added to the gradient once per CTA. ------------------------------------------------
constexpr int kStemBwdGroups = 4;
__global__ void __launch_bounds__(128 * kStemBwdGroups)
k_stem_backward(const uint4* __restrict__ states, const __nv_bfloat16* __restrict__ dz, uint32_t n_pos, ActLayout lay,
                float* __restrict__ d_w /*[9][2][128]*/, float* __restrict__ d_b /*[128]*/) {
  __shared__ float part[kStemBwdGroups - 1][19][kNetC];
  const int c = threadIdx.x & (kNetC - 1), grp = threadIdx.x >> 7;
  float acc[18], accb = 0.0f;
#pragma unroll
  for (int i = 0; i < 18; ++i) acc[i] = 0.0f;
  for (uint32_t pos = blockIdx.x * kStemBwdGroups + grp; pos < n_pos; pos += gridDim.x * kStemBwdGroups) {
    const uint4 st = states[pos];
    const uint64_t cur = (static_cast<uint64_t>(st.y) << 32) | st.x, opp = (static_cast<uint64_t>(st.w) << 32) | st.z;
#pragma unroll 1
    for (int r = 0; r < 6; ++r) {
      float d[7];
#pragma unroll
      for (int x = 0; x < 7; ++x) d[x] = __bfloat162float(dz[lay.row(pos, r, x) * kNetC + c]);
      // the three board rows the taps of this output row look at, as 7-bit masks (0 outside the board)
      uint32_t mc[3], mo[3];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int rr = r + dy - 1;
        const bool in = rr >= 0 && rr < 6;
        mc[dy] = in ? static_cast<uint32_t>(cur >> (rr * 7)) & 0x7Fu : 0u;
        mo[dy] = in ? static_cast<uint32_t>(opp >> (rr * 7)) & 0x7Fu : 0u;
      }
#pragma unroll
      for (int x = 0; x < 7; ++x) {
        accb += d[x];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int cc = x + tap % 3 - 1;
          if (cc < 0 || cc >= 7) continue;
          if ((mc[tap / 3] >> cc) & 1u) acc[tap * 2] += d[x];
          if ((mo[tap / 3] >> cc) & 1u) acc[tap * 2 + 1] += d[x];
        }
      }
    }
  }
  if (grp > 0) {
#pragma unroll
    for (int i = 0; i < 18; ++i) part[grp - 1][i][c] = acc[i];
    part[grp - 1][18][c] = accb;
  }
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      float t = acc[i];
      for (int g2 = 0; g2 < kStemBwdGroups - 1; ++g2) t += part[g2][i][c];
      atomicAdd(d_w + i * kNetC + c, t);
    }
    float t = accb;
    for (int g2 = 0; g2 < kStemBwdGroups - 1; ++g2) t += part[g2][18][c];
    atomicAdd(d_b + c, t);
  }
}

// ---- Adam (connect_four_net.py:112: AdamOptimizer(lr)); t = 1, 2, ... -----------------------------------------
__global__ void k_adam(float* __restrict__ prm, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v, size_t n,
                       float lr, float b1, float b2, float eps, float c1 /*1 - b1^t*/, float c2 /*1 - b2^t*/) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = grad[i];
  const float mi = b1 * m[i] + (1.0f - b1) * g, vi = b2 * v[i] + (1.0f - b2) * g * g;
  m[i] = mi;
  v[i] = vi;
  prm[i] -= lr * (mi / c1) / (sqrtf(vi / c2) + eps);
}

// ---- operand tiles from the fp32 parameters (what azb_nnet::upload does on the host) ---------------------------
__global__ void k_build_tiles(const float* __restrict__ prm, NetLayout L, uint16_t* __restrict__ fwd, uint16_t* __restrict__ bwd) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // (layer, kb, n, k)
  const size_t total = static_cast<size_t>(2 * L.R) * kTcKBlocks * 128 * 64;
  if (i >= total) return;
  const uint32_t k = i % 64, n = (i / 64) % 128, kb = (i / (64 * 128)) % kTcKBlocks, layer = static_cast<uint32_t>(i / (64 * 128 * kTcKBlocks));
  const uint32_t tap = kb >> 1, half = kb & 1;
  const size_t byte = (n / 8) * 1024 + (n % 8) * 128 + (((k / 8) ^ (n % 8)) * 16) + (k % 8) * 2;
  const size_t tile = (static_cast<size_t>(layer) * kTcKBlocks + kb) * (kTcTileBytes / 2) + byte / 2;
  const float* w = prm + L.tower_w + static_cast<size_t>(layer) * 9 * kNetC * kNetC;
  const __nv_bfloat16 f = __float2bfloat16_rn(w[(static_cast<size_t>(tap) * kNetC + half * 64 + k) * kNetC + n]);
  const __nv_bfloat16 b = __float2bfloat16_rn(w[(static_cast<size_t>(8 - tap) * kNetC + n) * kNetC + half * 64 + k]);
  fwd[tile] = *reinterpret_cast<const uint16_t*>(&f);
  bwd[tile] = *reinterpret_cast<const uint16_t*>(&b);
}
__global__ void k_build_stem_table(const float* __restrict__ prm, NetLayout L, float* __restrict__ tab) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3u * 64u * kNetC) return;
  const uint32_t c = e % kNetC, bits = (e / kNetC) % 64u, dyi = e / (kNetC * 64u);
  float acc = 0.0f;
  for (uint32_t dxi = 0; dxi < 3u; ++dxi)
    for (uint32_t pl = 0; pl < 2u; ++pl)
      if ((bits >> (dxi + 3u * pl)) & 1u) acc += prm[L.stem_w + ((dyi * 3u + dxi) * 2u + pl) * kNetC + c];
  tab[e] = acc;
}

}  // namespace azb
