// learn.cuh — the host side of Coach::learn (/root/reference/src/coach.rs:169-396), the sample history
// with its on-disk form (coach.rs:55-81,159-167) and the weight checkpoints.  Included at the end of
// engine.cu: everything here is bookkeeping around the three device phases (self-play rounds, training
// steps, arena rounds), which are the entry points defined above.
#pragma once

#include <dirent.h>
#include <sys/stat.h>

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

namespace {

// ---- bincode 1.3.1 default options: little-endian, fixed-width ints, u64 sequence lengths ---------------
constexpr uint64_t kSampleBytes = (1 + 8 + 3 * 8 + 8 + 84 * 4) + (1 + 8 + 8 + 7 * 4) + 4;  // 426

struct ByteSink {
  std::FILE* f;
  bool ok = true;
  void put(const void* p, size_t n) {
    if (ok && std::fwrite(p, 1, n, f) != n) ok = false;
  }
  void u8(uint8_t v) { put(&v, 1); }
  void u64(uint64_t v) { put(&v, 8); }  // the build targets little-endian hosts only (x86-64 / aarch64)
};

void encode_sample(uint8_t* out, const float* board, const float* pi, float v) {
  auto u64 = [&](uint64_t x) {
    std::memcpy(out, &x, 8);
    out += 8;
  };
  // board: ndarray ArrayD<f32> {v, dim, data}: dim of IxDyn is a sequence (length prefix)
  *out++ = 1;
  u64(3);
  u64(2);
  u64(6);
  u64(7);
  u64(84);
  std::memcpy(out, board, 84 * 4);
  out += 84 * 4;
  // pi: Array1<f32>: dim of Ix1 is [usize; 1] = a tuple, no length prefix
  *out++ = 1;
  u64(7);
  u64(7);
  std::memcpy(out, pi, 7 * 4);
  out += 7 * 4;
  std::memcpy(out, &v, 4);
}

// reads one sample at p (end = one past the last readable byte); returns bytes consumed or 0 on a format error
size_t decode_sample(const uint8_t* p, const uint8_t* end, float* board, float* pi, float* v) {
  const uint8_t* p0 = p;
  auto need = [&](size_t n) { return static_cast<size_t>(end - p) >= n; };
  auto u64 = [&](uint64_t* x) {
    if (!need(8)) return false;
    std::memcpy(x, p, 8);
    p += 8;
    return true;
  };
  uint64_t nd = 0, d = 0, len = 0, dims[3] = {0, 0, 0};
  if (!need(1) || *p++ != 1) return 0;
  if (!u64(&nd) || nd != 3) return 0;
  for (int i = 0; i < 3; ++i)
    if (!u64(&dims[i])) return 0;
  // two layouts exist: [2,6,7] (C,H,W: the declared get_feature_shape, connect_four_game.rs:87, and what this engine
  // writes) and [6,7,2] (H,W,C: what the reference's literal to_features produces, :219-237, defect F11).  The engine's
  // planes are channel-first, so a channel-last sample is transposed here; any other shape is not a connect-four sample.
  const bool chw = dims[0] == 2 && dims[1] == 6 && dims[2] == 7;
  const bool hwc = dims[0] == 6 && dims[1] == 7 && dims[2] == 2;
  if (!(chw || hwc) || !u64(&len) || len != 84 || !need(84 * 4)) return 0;
  if (board) {
    if (chw) {
      std::memcpy(board, p, 84 * 4);
    } else {
      for (int cell = 0; cell < 42; ++cell)
        for (int c = 0; c < 2; ++c) std::memcpy(board + c * 42 + cell, p + (cell * 2 + c) * 4, 4);
    }
  }
  p += 84 * 4;
  if (!need(1) || *p++ != 1) return 0;
  if (!u64(&d) || d != 7 || !u64(&len) || len != 7 || !need(7 * 4 + 4)) return 0;
  if (pi) std::memcpy(pi, p, 7 * 4);
  p += 7 * 4;
  if (v) std::memcpy(v, p, 4);
  p += 4;
  return static_cast<size_t>(p - p0);
}

int write_examples(const char* path, uint64_t n_iters, const uint64_t* counts, const float* boards, const float* pis,
                   const float* vs, const std::deque<SampleBlock>* hist) {
  std::FILE* f = std::fopen(path, "wb");
  if (!f) return fail(AZB_ERR_INVALID, std::string("cannot write ") + path + ": " + std::strerror(errno));
  ByteSink s{f};
  s.u64(n_iters);
  std::vector<uint8_t> buf(kSampleBytes * 1024);
  uint64_t base = 0;
  for (uint64_t it = 0; it < n_iters; ++it) {
    const uint64_t n = hist ? (*hist)[it].size() : counts[it];
    const float* b = hist ? (*hist)[it].boards.data() : boards + base * 84;
    const float* p = hist ? (*hist)[it].pis.data() : pis + base * 7;
    const float* v = hist ? (*hist)[it].vs.data() : vs + base;
    s.u64(n);
    for (uint64_t i = 0; i < n; i += 1024) {
      const uint64_t m = std::min<uint64_t>(1024, n - i);
      for (uint64_t j = 0; j < m; ++j) encode_sample(buf.data() + j * kSampleBytes, b + (i + j) * 84, p + (i + j) * 7, v[i + j]);
      s.put(buf.data(), m * kSampleBytes);
    }
    base += n;
  }
  const bool ok = s.ok && std::fclose(f) == 0;
  if (!ok) return fail(AZB_ERR_INVALID, std::string("short write to ") + path);
  return AZB_OK;
}

// whole-file reader: samples land in out (may be null: sizes only), per-entry sizes in counts
int read_examples(const char* path, std::deque<SampleBlock>* out, std::vector<uint64_t>* counts) {
  std::FILE* f = std::fopen(path, "rb");
  if (!f) return fail(AZB_ERR_INVALID, std::string("cannot read ") + path + ": " + std::strerror(errno));
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf(static_cast<size_t>(sz > 0 ? sz : 0));
  const bool rd = sz >= 0 && std::fread(buf.data(), 1, buf.size(), f) == buf.size();
  std::fclose(f);
  if (!rd) return fail(AZB_ERR_INVALID, std::string("short read from ") + path);
  const uint8_t *p = buf.data(), *end = p + buf.size();
  auto u64 = [&](uint64_t* x) {
    if (static_cast<size_t>(end - p) < 8) return false;
    std::memcpy(x, p, 8);
    p += 8;
    return true;
  };
  const std::string bad = std::string(path) + " is not a bincode VecDeque<VecDeque<TrainingSample>> of connect-four samples";
  uint64_t n_iters = 0;
  if (!u64(&n_iters) || n_iters > buf.size()) return fail(AZB_ERR_INVALID, bad);
  if (out) out->clear();
  if (counts) counts->clear();
  for (uint64_t it = 0; it < n_iters; ++it) {
    uint64_t n = 0;
    if (!u64(&n) || n > buf.size()) return fail(AZB_ERR_INVALID, bad);
    SampleBlock blk;
    if (out) {
      blk.boards.resize(n * 84);
      blk.pis.resize(n * 7);
      blk.vs.resize(n);
    }
    for (uint64_t i = 0; i < n; ++i) {
      const size_t used = decode_sample(p, end, out ? blk.boards.data() + i * 84 : nullptr, out ? blk.pis.data() + i * 7 : nullptr,
                                        out ? blk.vs.data() + i : nullptr);
      if (!used) return fail(AZB_ERR_INVALID, bad);
      p += used;
    }
    if (out) out->push_back(std::move(blk));
    if (counts) counts->push_back(n);
  }
  if (p != end) return fail(AZB_ERR_INVALID, bad + " (trailing bytes)");
  return AZB_OK;
}

// coach.rs:55-72: the entry of the directory with the largest numeric stem (only *.examples are considered here)
int latest_examples(const char* dir, uint64_t* iteration, bool* dir_exists) {
  DIR* d = opendir(dir);
  if (dir_exists) *dir_exists = d != nullptr;
  if (!d) return fail(AZB_ERR_INVALID, std::string("cannot open directory ") + dir);
  bool found = false;
  uint64_t best = 0;
  while (dirent* e = readdir(d)) {
    const std::string name = e->d_name;
    const std::string ext = ".examples";
    if (name.size() <= ext.size() || name.compare(name.size() - ext.size(), ext.size(), ext) != 0) continue;
    const std::string stem = name.substr(0, name.size() - ext.size());
    if (stem.empty() || stem.size() > 18 || stem.find_first_not_of("0123456789") != std::string::npos) continue;
    const uint64_t v = std::stoull(stem);
    if (!found || v > best) best = v;
    found = true;
  }
  closedir(d);
  if (!found) return fail(AZB_ERR_INVALID, std::string("no <n>.examples in ") + dir);
  *iteration = best;
  return AZB_OK;
}

// ---- host Philox-4x32-10 (the same generator the device draws actions with, mcts.cuh) ---------------------
void philox_host(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3], k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0, n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1, n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// coach.rs:296-297 `all_train_examples.shuffle(rng)`: rand 0.7's walk `for i in (1..len).rev() swap(i, gen_range(0, i+1))`
// with the draw taken from Philox: key (lo32 seed, lo32 iteration), counter (lo32 i, purpose 2, hi32 seed, hi32 i),
// j = high 64 bits of x * (i+1), x = out[0] | out[1] << 32.
constexpr uint32_t kPurposeShuffle = 2;
void shuffle_perm(uint64_t seed, uint64_t iteration, uint64_t n, uint64_t* perm) {
  for (uint64_t i = 0; i < n; ++i) perm[i] = i;
  for (uint64_t i = n; i-- > 1;) {
    const uint32_t ctr[4] = {static_cast<uint32_t>(i), kPurposeShuffle, static_cast<uint32_t>(seed >> 32), static_cast<uint32_t>(i >> 32)};
    const uint32_t key[2] = {static_cast<uint32_t>(seed), static_cast<uint32_t>(iteration)};
    uint32_t o[4];
    philox_host(ctr, key, o);
    const uint64_t x = static_cast<uint64_t>(o[0]) | (static_cast<uint64_t>(o[1]) << 32);
    const uint64_t j = static_cast<uint64_t>((static_cast<unsigned __int128>(x) * (i + 1)) >> 64);
    std::swap(perm[i], perm[j]);
  }
}

// coach.rs:383-390
bool accept_new_model(uint64_t nwins, uint64_t pwins, float update_threshold) {
  if (pwins + nwins == 0) return false;
  return !(static_cast<float>(nwins) / static_cast<float>(pwins + nwins) < update_threshold);
}

struct NetDeleter {
  void operator()(azb_nnet* n) const { azb_nnet_destroy(n); }
};

double wall_ms(std::chrono::steady_clock::time_point a) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
}

}  // namespace

// Coach::setup, coach.rs:55-81: the newest `<n>.examples` of the checkpoint directory becomes the history.  A missing
// directory means an empty history; it is created when the first file is written (the reference creates it at once).
static int coach_resume_history(azb_coach* c) {
  uint64_t it = 0;
  bool exists = false;
  if (latest_examples(c->checkpoint_dir.c_str(), &it, &exists) != AZB_OK) return AZB_OK;  // no directory / no file yet
  const std::string path = c->checkpoint_dir + "/" + std::to_string(it) + ".examples";
  const int rc = read_examples(path.c_str(), &c->history.entries, nullptr);
  // a resumed coach numbers its iterations after the file it resumed from: game ids (the Philox streams), shuffles and
  // `<n>.examples` names continue instead of replaying iteration 0 into a window that already holds it
  if (rc == AZB_OK) c->resume_base = it + 1;
  return rc;
}

__global__ void k_scale_f32(float* __restrict__ p, size_t n, float s) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] *= s;
}

static int ensure_dir(const std::string& dir) {
  struct stat st;
  if (stat(dir.c_str(), &st) == 0) return S_ISDIR(st.st_mode) ? AZB_OK : fail(AZB_ERR_INVALID, dir + " is not a directory");
  if (mkdir(dir.c_str(), 0777) != 0) return fail(AZB_ERR_INVALID, "cannot create " + dir + ": " + std::strerror(errno));
  return AZB_OK;
}

extern "C" {

int azb_examples_write(const char* path, uint64_t n_iters, const uint64_t* counts, const float* boards, const float* pis,
                       const float* vs) {
  if (!path || (n_iters && !counts)) return fail(AZB_ERR_INVALID, "NULL argument");
  uint64_t total = 0;
  for (uint64_t i = 0; i < n_iters; ++i) total += counts[i];
  if (total && (!boards || !pis || !vs)) return fail(AZB_ERR_INVALID, "NULL argument");
  return write_examples(path, n_iters, counts, boards, pis, vs, nullptr);
}

int azb_examples_stat(const char* path, uint64_t* n_iters, uint64_t* counts, uint64_t cap_iters, uint64_t* n_samples) {
  if (!path || !n_iters) return fail(AZB_ERR_INVALID, "NULL argument");
  std::vector<uint64_t> c;
  const int rc = read_examples(path, nullptr, &c);
  if (rc) return rc;
  *n_iters = c.size();
  uint64_t total = 0;
  for (size_t i = 0; i < c.size(); ++i) {
    total += c[i];
    if (counts && i < cap_iters) counts[i] = c[i];
  }
  if (n_samples) *n_samples = total;
  return AZB_OK;
}

int azb_examples_read(const char* path, float* boards, float* pis, float* vs, uint64_t cap_samples) {
  if (!path || !boards || !pis || !vs) return fail(AZB_ERR_INVALID, "NULL argument");
  std::deque<SampleBlock> h;
  const int rc = read_examples(path, &h, nullptr);
  if (rc) return rc;
  uint64_t total = 0;
  for (auto& b : h) total += b.size();
  if (total > cap_samples) return fail(AZB_ERR_CAPACITY, "sample buffer too small");
  uint64_t at = 0;
  for (auto& b : h) {
    std::memcpy(boards + at * 84, b.boards.data(), b.boards.size() * 4);
    std::memcpy(pis + at * 7, b.pis.data(), b.pis.size() * 4);
    std::memcpy(vs + at, b.vs.data(), b.vs.size() * 4);
    at += b.size();
  }
  return AZB_OK;
}

int azb_examples_latest(const char* checkpoint_directory, uint64_t* iteration) {
  if (!checkpoint_directory || !iteration) return fail(AZB_ERR_INVALID, "NULL argument");
  return latest_examples(checkpoint_directory, iteration, nullptr);
}

int azb_learn_accept(uint64_t nwins, uint64_t pwins, float update_threshold) {
  return accept_new_model(nwins, pwins, update_threshold) ? 1 : 0;
}

int azb_learn_shuffle_perm(uint64_t seed, uint64_t iteration, uint64_t n, uint64_t* perm) {
  if (n && !perm) return fail(AZB_ERR_INVALID, "NULL argument");
  shuffle_perm(seed, iteration, n, perm);
  return AZB_OK;
}

void azb_learn_config_default(azb_learn_config* lc) {
  if (!lc) return;
  std::memset(lc, 0, sizeof(*lc));
  lc->epochs = 10;      // connect_four_net.py:13
  lc->batch_size = 64;  // connect_four_net.py:14
  // the reference's Adam runs at 1e-3 (connect_four_net.py:21) on a network with BatchNorm in front of every ReLU; this
  // network has none (folded away for inference) and at 1e-3 its 2-channel policy head dies within ~100 steps
  // (profiles/r1_train.md): the default step is 1e-4
  lc->adam = azb_train_config{1e-4f, 0.9f, 0.999f, 1e-8f};
  lc->save_files = 1;
  // gating games: 4 random opening plies (see azb_learn_config.arena_k_open for why 0 is not the default here)
  lc->arena_k_open = 4;
}

// ---- weight checkpoints ------------------------------------------------------------------------------------
int azb_nnet_save(azb_nnet* n, const char* path) {
  if (!n || !path) return fail(AZB_ERR_INVALID, "NULL argument");
  const uint64_t N = n->L.total;
  const int rcs = n->sync_host();
  if (rcs) return rcs;
  const bool adam = n->d_adam_m.bytes >= N * 4 && n->adam_t > 0;
  std::vector<float> m, v;
  if (adam) {
    AZB_CUDA(cudaSetDevice(n->cfg.device));
    m.resize(N);
    v.resize(N);
    AZB_CUDA(cudaMemcpy(m.data(), n->d_adam_m.p, N * 4, cudaMemcpyDeviceToHost));
    AZB_CUDA(cudaMemcpy(v.data(), n->d_adam_v.p, N * 4, cudaMemcpyDeviceToHost));
  }
  std::FILE* f = std::fopen(path, "wb");
  if (!f) return fail(AZB_ERR_INVALID, std::string("cannot write ") + path + ": " + std::strerror(errno));
  ByteSink s{f};
  const uint32_t hdr[3] = {1u, static_cast<uint32_t>(n->cfg.blocks), adam ? 1u : 0u};
  s.put("AZBW", 4);
  s.put(hdr, sizeof(hdr));
  s.u64(N);
  s.u64(adam ? n->adam_t : 0);
  s.put(n->h_params.data(), N * 4);
  if (adam) {
    s.put(m.data(), N * 4);
    s.put(v.data(), N * 4);
  }
  const bool ok = s.ok && std::fclose(f) == 0;
  if (!ok) return fail(AZB_ERR_INVALID, std::string("short write to ") + path);
  return AZB_OK;
}

int azb_nnet_load(azb_nnet* n, const char* path) {
  if (!n || !path) return fail(AZB_ERR_INVALID, "NULL argument");
  std::FILE* f = std::fopen(path, "rb");
  if (!f) return fail(AZB_ERR_INVALID, std::string("cannot read ") + path + ": " + std::strerror(errno));
  char magic[4];
  uint32_t hdr[3];
  uint64_t N = 0, t = 0;
  bool ok = std::fread(magic, 1, 4, f) == 4 && std::fread(hdr, 4, 3, f) == 3 && std::fread(&N, 8, 1, f) == 1 && std::fread(&t, 8, 1, f) == 1;
  ok = ok && std::memcmp(magic, "AZBW", 4) == 0 && hdr[0] == 1;
  if (!ok) {
    std::fclose(f);
    return fail(AZB_ERR_INVALID, std::string(path) + " is not an AZBW version-1 weight file");
  }
  if (hdr[1] != static_cast<uint32_t>(n->cfg.blocks) || N != n->L.total) {
    std::fclose(f);
    return fail(AZB_ERR_INVALID, std::string(path) + ": architecture differs from the network handle");
  }
  std::vector<float> w(N), m, v;
  ok = std::fread(w.data(), 4, N, f) == N;
  if (ok && hdr[2]) {
    m.resize(N);
    v.resize(N);
    ok = std::fread(m.data(), 4, N, f) == N && std::fread(v.data(), 4, N, f) == N;
  }
  std::fclose(f);
  if (!ok) return fail(AZB_ERR_INVALID, std::string("short read from ") + path);
  int rc = azb_nnet_set_params(n, w.data(), N);
  if (rc) return rc;
  if (hdr[2]) {
    AZB_CUDA(n->d_adam_m.ensure(N * 4));
    AZB_CUDA(n->d_adam_v.ensure(N * 4));
    AZB_CUDA(cudaMemcpy(n->d_adam_m.p, m.data(), N * 4, cudaMemcpyHostToDevice));
    AZB_CUDA(cudaMemcpy(n->d_adam_v.p, v.data(), N * 4, cudaMemcpyHostToDevice));
    n->adam_t = t;
  } else {
    n->d_adam_m.release();
    n->d_adam_v.release();
    n->adam_t = 0;
  }
  return AZB_OK;
}

int azb_nnet_copy(azb_nnet* dst, azb_nnet* src) {
  if (!dst || !src) return fail(AZB_ERR_INVALID, "NULL argument");
  if (dst == src) return AZB_OK;
  if (dst->cfg.blocks != src->cfg.blocks || dst->cfg.precision != src->cfg.precision || dst->cfg.device != src->cfg.device)
    return fail(AZB_ERR_INVALID, "networks differ in architecture, precision or device");
  const uint64_t N = src->L.total;
  AZB_CUDA(cudaSetDevice(src->cfg.device));
  // everything the forward / backward kernels read is copied device to device (no host-side tile build): the fp32 master
  // parameters, the forward and backward bf16 weight tiles, the stem table; the head weights travel as a host struct
  auto d2d = [](DevBuf& d, const DevBuf& s_) -> cudaError_t {
    if (!s_.p) return cudaSuccess;
    cudaError_t e = d.ensure(s_.bytes);
    return e != cudaSuccess ? e : cudaMemcpy(d.p, s_.p, s_.bytes, cudaMemcpyDeviceToDevice);
  };
  AZB_CUDA(d2d(dst->d_params, src->d_params));
  AZB_CUDA(d2d(dst->d_wtiles, src->d_wtiles));
  AZB_CUDA(d2d(dst->d_wtiles_bwd, src->d_wtiles_bwd));
  AZB_CUDA(d2d(dst->d_stem_tab, src->d_stem_tab));
  dst->wtile_copy_bytes = src->wtile_copy_bytes;
  dst->head_w = src->head_w;
  dst->h_params = src->h_params;
  dst->host_stale = src->host_stale;
  if (src->d_adam_m.bytes >= N * 4 && src->adam_t > 0) {
    AZB_CUDA(dst->d_adam_m.ensure(N * 4));
    AZB_CUDA(dst->d_adam_v.ensure(N * 4));
    AZB_CUDA(cudaMemcpy(dst->d_adam_m.p, src->d_adam_m.p, N * 4, cudaMemcpyDeviceToDevice));
    AZB_CUDA(cudaMemcpy(dst->d_adam_v.p, src->d_adam_v.p, N * 4, cudaMemcpyDeviceToDevice));
    dst->adam_t = src->adam_t;
  } else {
    dst->d_adam_m.release();
    dst->d_adam_v.release();
    dst->adam_t = 0;
  }
  dst->grads_ready = false;
  return AZB_OK;
}

// ---- history -----------------------------------------------------------------------------------------------
int azb_coach_history_stat(azb_coach* c, uint64_t* n_iters, uint64_t* counts, uint64_t cap_iters, uint64_t* n_samples) {
  if (!c || !n_iters) return fail(AZB_ERR_INVALID, "NULL argument");
  const auto& h = c->history.entries;
  *n_iters = h.size();
  uint64_t total = 0;
  for (size_t i = 0; i < h.size(); ++i) {
    total += h[i].size();
    if (counts && i < cap_iters) counts[i] = h[i].size();
  }
  if (n_samples) *n_samples = total;
  return AZB_OK;
}

int azb_coach_history_export(azb_coach* c, float* boards, float* pis, float* vs, uint64_t cap_samples) {
  if (!c || !boards || !pis || !vs) return fail(AZB_ERR_INVALID, "NULL argument");
  uint64_t total = 0;
  for (auto& b : c->history.entries) total += b.size();
  if (total > cap_samples) return fail(AZB_ERR_CAPACITY, "sample buffer too small");
  uint64_t at = 0;
  for (auto& b : c->history.entries) {
    std::memcpy(boards + at * 84, b.boards.data(), b.boards.size() * 4);
    std::memcpy(pis + at * 7, b.pis.data(), b.pis.size() * 4);
    std::memcpy(vs + at, b.vs.data(), b.vs.size() * 4);
    at += b.size();
  }
  return AZB_OK;
}

int azb_coach_save_train_examples(azb_coach* c, uint64_t iteration, const char* checkpoint_directory) {
  if (!c || !checkpoint_directory) return fail(AZB_ERR_INVALID, "NULL argument");
  const int rc = ensure_dir(checkpoint_directory);
  if (rc) return rc;
  const std::string path = std::string(checkpoint_directory) + "/" + std::to_string(iteration) + ".examples";
  return write_examples(path.c_str(), c->history.entries.size(), nullptr, nullptr, nullptr, nullptr, &c->history.entries);
}

int azb_coach_load_train_examples(azb_coach* c, const char* path) {
  if (!c || !path) return fail(AZB_ERR_INVALID, "NULL argument");
  std::deque<SampleBlock> h;
  const int rc = read_examples(path, &h, nullptr);
  if (rc) return rc;
  c->history.entries = std::move(h);
  return AZB_OK;
}

// ---- Coach::learn — coach.rs:169-396 -----------------------------------------------------------------------
// With `dist` (world > 1) the iteration is data parallel over one process per GPU: every rank self-plays a contiguous share
// of the num_eps games and keeps their samples (no collective on the self-play path), the training steps average the
// gradients through the caller's all-reduce (NCCL over NVLink in place on the device buffer), the gating games are split
// over the ranks and their three counters summed, so that every rank takes the same accept decision on identical models.
int azb_coach_learn_dist(azb_coach* c, const azb_nnet_config* net_cfg, const azb_learn_config* lc_in, const azb_dist* dist,
                         azb_learn_report* reports, uint64_t cap_reports, uint64_t* n_reports, azb_nnet** final_net) {
  if (!c || !net_cfg) return fail(AZB_ERR_INVALID, "NULL argument");
  const uint32_t world = dist ? dist->world : 1u, rank = dist ? dist->rank : 0u;
  if (dist && (world == 0 || rank >= world || !dist->allreduce_sum_f32_device || !dist->allreduce_sum_u64_host))
    return fail(AZB_ERR_INVALID, "azb_dist needs rank < world and both all-reduce callbacks");
  auto sum_u64 = [&](uint64_t* v, uint64_t n) -> int {
    if (world == 1) return AZB_OK;
    return dist->allreduce_sum_u64_host(v, n, dist->user) ? fail(AZB_ERR_INVALID, "allreduce_sum_u64_host failed") : AZB_OK;
  };
  // Every rank must leave a collective phase together: a rank that failed locally (capacity, file I/O, a CUDA error)
  // tells the others at the next phase boundary instead of returning while they wait in an all-reduce forever.
  auto agree = [&](int rc) -> int {
    if (world == 1) return rc;
    const std::string mine = rc ? g_err : std::string();
    uint64_t bad = rc ? 1u : 0u;
    if (dist->allreduce_sum_u64_host(&bad, 1, dist->user)) return fail(AZB_ERR_INVALID, "allreduce_sum_u64_host failed");
    if (rc) return fail(rc, mine);
    if (bad) return fail(AZB_ERR_INVALID, "another rank failed in this phase of Coach::learn (its own call reports why)");
    return AZB_OK;
  };
  auto split = [&](uint64_t total, uint64_t* first, uint64_t* n) {  // contiguous shares, the remainder over the first ranks
    const uint64_t base = total / world, rem = total % world;
    *n = base + (rank < rem ? 1 : 0);
    *first = rank * base + std::min<uint64_t>(rank, rem);
  };
  if (n_reports) *n_reports = 0;
  if (final_net) *final_net = nullptr;
  azb_learn_config lc;
  if (lc_in) lc = *lc_in;
  else azb_learn_config_default(&lc);
  const azb_config& cfg = c->cfg;
  if (cfg.evaluator != AZB_EVAL_NNET) return fail(AZB_ERR_INVALID, "Coach::learn needs evaluator AZB_EVAL_NNET");
  if (net_cfg->precision != AZB_NNET_BF16_TC) return fail(AZB_ERR_UNSUPPORTED, "training needs the tensor-core tower");
  if (lc.batch_size == 0 || lc.batch_size > (1u << 20)) return fail(AZB_ERR_INVALID, "batch_size out of range");
  if (cfg.num_eps < world || cfg.max_history_length == 0) return fail(AZB_ERR_INVALID, "num_eps (>= ranks) and max_history_length must be positive");
  if (lc.batch_size < world) return fail(AZB_ERR_INVALID, "batch_size is the global batch: it must be >= the number of ranks");
  const bool files = lc.save_files != 0;
  if (files && !cfg.checkpoint_directory) return fail(AZB_ERR_INVALID, "save_files needs checkpoint_directory");
  const std::string dir = cfg.checkpoint_directory ? cfg.checkpoint_directory : "";

  // two device-resident models: nets[cur] is `model_id`, nets[cur ^ 1] is the candidate `model_id + 1`
  std::unique_ptr<azb_nnet, NetDeleter> nets[2];
  for (int k = 0; k < 2; ++k) {
    azb_nnet* n = nullptr;
    const int rc = azb_nnet_create(net_cfg, &n);
    if (rc) return rc;
    nets[k].reset(n);
  }
  int cur = 0;
  uint64_t model_id = 0;  // coach.rs:177
  if (files) {
    const int rcd = ensure_dir(dir);
    if (rcd) return rcd;
    const std::string w0 = dir + "/0.azbw";
    struct stat st;
    if (stat(w0.c_str(), &st) == 0) {
      const int rc = azb_nnet_load(nets[0].get(), w0.c_str());
      if (rc) return rc;
    } else {
      const int rc = azb_nnet_save(nets[0].get(), w0.c_str());
      if (rc) return rc;
    }
  }
  azb_nnet* const caller_net = c->net;
  struct Restore {  // the coach's evaluator is borrowed for the duration of the call
    azb_coach* c;
    azb_nnet* n;
    ~Restore() { c->net = n; }
  } restore{c, caller_net};

  auto& hist = c->history.entries;
  std::vector<float> sb, sp, sv;  // shuffled structure-of-arrays window
  std::vector<uint64_t> perm;
  for (uint64_t iteration = 0; iteration < cfg.num_iters; ++iteration) {
    azb_learn_report rep{};
    rep.iteration = iteration;
    rep.model_id_before = model_id;
    SampleBlock blk;
    auto t0 = std::chrono::steady_clock::now();
    const uint64_t it_id = c->resume_base + iteration;  // numbering continues after a resumed `<n>.examples`
    if (!lc.skip_first_play || iteration > 0) {  // coach.rs:240
      int rc = azb_coach_set_nnet(c, nets[cur].get());
      azb_selfplay_stats st{};
      uint64_t ep_first = 0, ep_n = 0;
      split(cfg.num_eps, &ep_first, &ep_n);
      if (!rc) rc = azb_coach_self_play(c, ep_n, it_id * cfg.num_eps + ep_first, &st);  // coach.rs:241-272
      uint64_t n = 0;
      if (!rc) {
        azb_coach_num_samples(c, &n);
        blk.boards.resize(n * 84);
        blk.pis.resize(n * 7);
        blk.vs.resize(n);
        rc = azb_coach_export_samples(c, blk.boards.data(), blk.pis.data(), blk.vs.data(), n, nullptr);
      }
      rc = agree(rc);
      if (rc) return rc;
      rep.games = st.games;
      rep.samples_played = n;
      const uint64_t queue_cap = (cfg.max_queue_length + world - 1) / world;  // this rank's share of the queue
      if (n > queue_cap) blk.drop_front(n - queue_cap);  // coach.rs:274-277
    }
    rep.samples_kept = blk.size();
    rep.selfplay_ms = wall_ms(t0);
    hist.push_back(std::move(blk));                                    // coach.rs:284
    if (hist.size() > cfg.max_history_length) hist.pop_front();        // coach.rs:286-289
    {                                                                  // coach.rs:291-293
      const int rc = agree(files ? azb_coach_save_train_examples(c, it_id, dir.c_str()) : AZB_OK);
      if (rc) return rc;
    }
    uint64_t total = 0;
    for (auto& b : hist) total += b.size();
    rep.history_iterations = hist.size();
    rep.history_samples = total;
    uint64_t glob[2] = {total, total == 0 ? 1u : 0u};  // {samples over all ranks, ranks without samples}
    {
      const int rcg = sum_u64(glob, 2);
      if (rcg) return rcg;
    }
    if (glob[1]) return fail(AZB_ERR_INVALID, "no training samples (coach.rs:304 assert!(num_samples > 0))");

    // coach.rs:295-327: flatten, shuffle, AOS -> SOA
    t0 = std::chrono::steady_clock::now();
    perm.resize(total);
    shuffle_perm(cfg.seed + rank, it_id, total, perm.data());
    sb.resize(total * 84);
    sp.resize(total * 7);
    sv.resize(total);
    {
      std::vector<const SampleBlock*> owner;
      std::vector<uint64_t> start;
      uint64_t acc = 0;
      for (auto& b : hist) {
        owner.push_back(&b);
        start.push_back(acc);
        acc += b.size();
      }
      for (uint64_t i = 0; i < total; ++i) {
        const uint64_t src = perm[i];
        size_t e = static_cast<size_t>(std::upper_bound(start.begin(), start.end(), src) - start.begin()) - 1;
        const uint64_t k = src - start[e];
        std::memcpy(&sb[i * 84], &owner[e]->boards[k * 84], 84 * 4);
        std::memcpy(&sp[i * 7], &owner[e]->pis[k * 7], 7 * 4);
        sv[i] = owner[e]->vs[k];
      }
    }
    // coach.rs:329-331: train(samples, model_id, model_id + 1)
    azb_nnet* cand = nets[cur ^ 1].get();
    int rc = agree(azb_nnet_copy(cand, nets[cur].get()));
    if (rc) return rc;
    // batch_size is the global batch: every rank contributes batch_size / world samples per step; the step count is the
    // same on all ranks (a pass = the global window once)
    const uint64_t bs = std::min<uint64_t>(std::max<uint64_t>(lc.batch_size / world, 1), total);
    const uint64_t steps = lc.epochs ? lc.epochs : (glob[0] + static_cast<uint64_t>(lc.batch_size) - 1) / lc.batch_size;
    std::vector<float> wb, wp, wv;  // a batch that wraps around the end of the list
    for (uint64_t s = 0; s < steps; ++s) {
      const uint64_t at = (s * bs) % total;
      const float *pb = &sb[at * 84], *pp = &sp[at * 7], *pv = &sv[at];
      if (at + bs > total) {
        wb.resize(bs * 84);
        wp.resize(bs * 7);
        wv.resize(bs);
        for (uint64_t i = 0; i < bs; ++i) {
          const uint64_t k = (at + i) % total;
          std::memcpy(&wb[i * 84], &sb[k * 84], 84 * 4);
          std::memcpy(&wp[i * 7], &sp[k * 7], 7 * 4);
          wv[i] = sv[k];
        }
        pb = wb.data();
        pp = wp.data();
        pv = wv.data();
      }
      float loss[2] = {0.0f, 0.0f};
      rc = agree(azb_nnet_train_begin(cand, pb, pp, pv, bs, loss));  // (before the gradient all-reduce: nobody waits alone)
      if (rc) return rc;
      if (world > 1) {  // mean over ranks of the per-rank mean gradients = gradient of the mean loss over the global batch
        AZB_CUDA(cudaDeviceSynchronize());
        if (dist->allreduce_sum_f32_device(cand->d_grad.p, cand->L.total, dist->user)) return fail(AZB_ERR_INVALID, "allreduce_sum_f32_device failed");
        k_scale_f32<<<static_cast<unsigned>((cand->L.total + 255) / 256), 256>>>(cand->d_grad.as<float>(), cand->L.total, 1.0f / static_cast<float>(world));
        AZB_CUDA(cudaGetLastError());
      }
      rc = azb_nnet_train_apply(cand, &lc.adam);
      if (rc) return rc;
      if (s == 0) std::memcpy(rep.loss_first, loss, 8);
      std::memcpy(rep.loss_last, loss, 8);
    }
    rep.train_steps = steps;
    rep.train_ms = wall_ms(t0);
    {  // python_nnet.rs:76-79 save_checkpoint(model, model_id, checkpoint)
      const std::string wpath = dir + "/" + std::to_string(model_id + 1) + ".azbw";
      rc = agree(files ? azb_nnet_save(cand, wpath.c_str()) : AZB_OK);
      if (rc) return rc;
    }

    // coach.rs:333-375: the candidate (player A) against the current model, temp 0
    t0 = std::chrono::steady_clock::now();
    uint64_t counts[3] = {0, 0, 0};
    {
      uint64_t pair_first = 0, pairs = 0;  // arena.rs:83: num / 2 games per seat order; the pairs are split over the ranks
      split(cfg.num_arena_games / 2, &pair_first, &pairs);
      azb_config acfg = cfg;
      // the arena keys its opening streams by (seed, game id): ids are distinct over iterations, ranks and games
      azb_arena_opts ao{};
      ao.k_open = lc.arena_k_open;
      ao.shared_trees = lc.arena_shared_trees;
      ao.first_game_id = it_id * cfg.num_arena_games + 2 * pair_first;
      rc = agree(azb_arena_play_games_ex(&acfg, 2 * pairs, AZB_EVAL_NNET, AZB_EVAL_NNET, cand, nets[cur].get(), &ao, counts,
                                         nullptr, nullptr, nullptr, nullptr, nullptr));
      if (rc) return rc;
      rc = sum_u64(counts, 3);
      if (rc) return rc;
    }
    rep.arena_ms = wall_ms(t0);
    rep.nwins = counts[0];  // coach.rs:377-379
    rep.pwins = counts[1];
    rep.draws = counts[2];
    rep.accepted = accept_new_model(rep.nwins, rep.pwins, cfg.update_threshold) ? 1 : 0;  // coach.rs:383-390
    if (rep.accepted) {
      model_id += 1;
      cur ^= 1;
    }
    rep.model_id_after = model_id;
    if (reports && iteration < cap_reports) reports[iteration] = rep;
    if (n_reports) *n_reports = iteration + 1;
  }
  if (final_net) *final_net = nets[cur].release();
  return AZB_OK;
}

int azb_coach_learn(azb_coach* c, const azb_nnet_config* net_cfg, const azb_learn_config* lc, azb_learn_report* reports,
                    uint64_t cap_reports, uint64_t* n_reports, azb_nnet** final_net) {
  return azb_coach_learn_dist(c, net_cfg, lc, nullptr, reports, cap_reports, n_reports, final_net);
}

}  // extern "C"
