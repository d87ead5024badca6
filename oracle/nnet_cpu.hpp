// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/quirks.h).
//
// Plain C++ fp32 forward pass of the evaluator network behind NNet::predict
// (/root/reference/src/nnet.rs:40-44), for the CPU-baseline leg of BASELINE configs 1 and 3
// (bench.py `cpu_baseline` / --impl reference) and as a third witness of the network's numerics.
// The reference itself has no working network (connect_four_net.py is a non-functional
// TF1 file); the architecture is this build's definition (alphazero-rs_b200/csrc/nnet.cuh):
//
//   in [2,6,7] -> conv3x3(2->C)+ReLU -> R x { conv3x3 C->C, ReLU, conv3x3 C->C, +skip, ReLU }
//   policy: conv1x1(C->2)+ReLU -> FC(84->7) -> softmax      value: conv1x1(C->1)+ReLU -> FC(42->64)+ReLU
//   -> FC(64->1) -> tanh;  C = 128.
//
// Flat parameter vector (the layout azb_nnet_get_params returns):
//   stem_w[9][2][C] stem_b[C] tower_w[2R][9][C][C] tower_b[2R][C] pol_w[C][2] pol_b[2] pol_fc_w[84][7]
//   pol_fc_b[7] val_w[C] val_b[1] val_fc1_w[42][64] val_fc1_b[64] val_fc2_w[64] val_fc2_b[1]
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

#include "mcts.hpp"

namespace azo {

struct CpuNet {
  static constexpr int C = 128, CELLS = 42;
  int R = 0;
  std::vector<float> prm;
  size_t stem_w, stem_b, tower_w, tower_b, pol_w, pol_b, pol_fc_w, pol_fc_b, val_w, val_b, val_fc1_w, val_fc1_b,
      val_fc2_w, val_fc2_b, total;

  static size_t num_params(int R) {
    return 9 * 2 * C + C + static_cast<size_t>(2 * R) * 9 * C * C + static_cast<size_t>(2 * R) * C + C * 2 + 2 +
           84 * 7 + 7 + C + 1 + 42 * 64 + 64 + 64 + 1;
  }
  CpuNet(int blocks, const float* params) : R(blocks) {
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += n; return r; };
    stem_w = take(9 * 2 * C); stem_b = take(C);
    tower_w = take(static_cast<size_t>(2 * R) * 9 * C * C); tower_b = take(static_cast<size_t>(2 * R) * C);
    pol_w = take(C * 2); pol_b = take(2); pol_fc_w = take(84 * 7); pol_fc_b = take(7);
    val_w = take(C); val_b = take(1); val_fc1_w = take(42 * 64); val_fc1_b = take(64);
    val_fc2_w = take(64); val_fc2_b = take(1);
    total = o;
    prm.assign(params, params + total);
  }

  // out[cell][co] = bias[co] + sum_tap sum_ci in[cell + tap][ci] * w[tap][ci][co]   ("same" padding)
  void conv3x3(const float* w, const float* b, int cin, const float* in /*[42][cin]*/, float* out /*[42][C]*/) const {
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 7; ++c) {
        float* o = out + (r * 7 + c) * C;
        for (int co = 0; co < C; ++co) o[co] = b[co];
        for (int tap = 0; tap < 9; ++tap) {
          const int rr = r + tap / 3 - 1, cc = c + tap % 3 - 1;
          if (rr < 0 || rr >= 6 || cc < 0 || cc >= 7) continue;
          const float* x = in + (rr * 7 + cc) * cin;
          const float* wt = w + static_cast<size_t>(tap) * cin * C;
          for (int ci = 0; ci < cin; ++ci) {
            const float xv = x[ci];
            if (xv == 0.0f) continue;  // post-ReLU activations are sparse; skipping zeros changes no sum
            const float* wr = wt + static_cast<size_t>(ci) * C;
            for (int co = 0; co < C; ++co) o[co] += xv * wr[co];
          }
        }
      }
  }

  // boards: [2][6][7] f32 planes (Game::to_features); pi[7], v
  void predict_one(const float* board, float* pi, float* v) const {
    float in0[CELLS * 2];
    for (int cell = 0; cell < CELLS; ++cell) {
      in0[cell * 2 + 0] = board[cell];
      in0[cell * 2 + 1] = board[42 + cell];
    }
    std::vector<float> a0(CELLS * C), a1(CELLS * C), a2(CELLS * C);
    const float* P = prm.data();
    conv3x3(P + stem_w, P + stem_b, 2, in0, a0.data());
    for (auto& x : a0) x = x > 0.0f ? x : 0.0f;
    for (int blk = 0; blk < R; ++blk) {
      const float* w1 = P + tower_w + static_cast<size_t>(2 * blk) * 9 * C * C;
      const float* w2 = w1 + static_cast<size_t>(9) * C * C;
      conv3x3(w1, P + tower_b + (2 * blk) * C, C, a0.data(), a1.data());
      for (auto& x : a1) x = x > 0.0f ? x : 0.0f;
      conv3x3(w2, P + tower_b + (2 * blk + 1) * C, C, a1.data(), a2.data());
      for (size_t i = 0; i < a0.size(); ++i) {
        const float s = a2[i] + a0[i];
        a0[i] = s > 0.0f ? s : 0.0f;
      }
    }
    float pol[84], val[42], h1[64], logit[7];
    for (int cell = 0; cell < CELLS; ++cell) {
      float s0 = P[pol_b + 0], s1 = P[pol_b + 1], sv = P[val_b];
      const float* x = a0.data() + cell * C;
      for (int ci = 0; ci < C; ++ci) {
        s0 += x[ci] * P[pol_w + ci * 2 + 0];
        s1 += x[ci] * P[pol_w + ci * 2 + 1];
        sv += x[ci] * P[val_w + ci];
      }
      pol[cell] = s0 > 0.0f ? s0 : 0.0f;
      pol[42 + cell] = s1 > 0.0f ? s1 : 0.0f;
      val[cell] = sv > 0.0f ? sv : 0.0f;
    }
    for (int a = 0; a < 7; ++a) {
      float s = P[pol_fc_b + a];
      for (int i = 0; i < 84; ++i) s += pol[i] * P[pol_fc_w + i * 7 + a];
      logit[a] = s;
    }
    for (int j = 0; j < 64; ++j) {
      float s = P[val_fc1_b + j];
      for (int i = 0; i < 42; ++i) s += val[i] * P[val_fc1_w + i * 64 + j];
      h1[j] = s > 0.0f ? s : 0.0f;
    }
    float m = logit[0];
    for (int a = 1; a < 7; ++a) m = logit[a] > m ? logit[a] : m;
    float e[7], sum = 0.0f;
    for (int a = 0; a < 7; ++a) { e[a] = std::exp(logit[a] - m); sum += e[a]; }
    for (int a = 0; a < 7; ++a) pi[a] = e[a] / sum;
    float s = P[val_fc2_b];
    for (int j = 0; j < 64; ++j) s += h1[j] * P[val_fc2_w + j];
    *v = std::tanh(s);
  }
};

// NNet::predict over the CPU network (batch rows one after the other, as the reference's batch-1 inference thread does)
struct CpuNetEvaluator : Evaluator {
  const CpuNet* net;
  explicit CpuNetEvaluator(const CpuNet* n) : net(n) {}
  void predict(const float* boards, size_t batch, size_t, size_t, float* pi, float* v) override {
    for (size_t i = 0; i < batch; ++i) net->predict_one(boards + 84 * i, pi + 7 * i, v + i);
  }
};

}  // namespace azo
