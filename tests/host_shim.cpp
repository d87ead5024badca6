// Test-only: compiles the product's __host__ __device__ arithmetic (c4_bitboard.cuh, tree.cuh)
// with g++ so that the bitboard logic, the packed-counter arithmetic and PUCT can be checked
// against the oracle on the CPU box (the GPU tests then check the same code on the device).
// Built with -ffp-contract=off to match the device's separately rounded __f*_rn intrinsics.
#include <cstdint>
#include "../alphazero-rs_b200/csrc/tree.cuh"

using namespace azb;
extern "C" {
int shim_game_ended_code(uint64_t cur, uint64_t opp, uint32_t quirks) { return game_ended_code(BB{cur, opp}, quirks); }
uint64_t shim_state_key(uint64_t cur, uint64_t opp) { return state_key(BB{cur, opp}); }
void shim_play_canonical(uint64_t cur, uint64_t opp, int a, uint64_t* out) {
  BB n = play_canonical(BB{cur, opp}, a);
  out[0] = n.cur; out[1] = n.opp;
}
uint32_t shim_valid_mask(uint64_t occ) { return valid_mask(occ); }
uint64_t shim_mirror(uint64_t b) { return mirror(b); }
uint64_t shim_unvisit(uint64_t c, float v, uint32_t quirks) { return counter_unvisit(c, v, quirks); }
float shim_puct(uint64_t child, float prior, uint32_t parent_n, int cpuct) {
  float sq = sqrtf(static_cast<float>(parent_n) + kEps);
  return puct_u(child, prior, sq, static_cast<float>(cpuct));
}
float shim_terminal_e(uint32_t code) { return terminal_e(code); }
}
