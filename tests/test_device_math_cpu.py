"""The product's __host__ __device__ arithmetic (bitboards, packed counter, PUCT) compiled for
the host (tests/host_shim.cpp) and compared with the oracle's array-based restatement."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "shim.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out,
                    os.path.join(ROOT, "tests", "host_shim.cpp")], check=True)
    L = C.CDLL(out)
    L.shim_game_ended_code.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    L.shim_state_key.restype = C.c_uint64
    L.shim_state_key.argtypes = [C.c_uint64, C.c_uint64]
    L.shim_play_canonical.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    L.shim_valid_mask.restype = C.c_uint32
    L.shim_valid_mask.argtypes = [C.c_uint64]
    L.shim_mirror.restype = C.c_uint64
    L.shim_mirror.argtypes = [C.c_uint64]
    L.shim_unvisit.restype = C.c_uint64
    L.shim_unvisit.argtypes = [C.c_uint64, C.c_float, C.c_uint32]
    L.shim_puct.restype = C.c_float
    L.shim_puct.argtypes = [C.c_uint64, C.c_float, C.c_uint32, C.c_int]
    return L


def bits(cells, val):
    b = 0
    for r in range(6):
        for c in range(7):
            if cells[r][c] == val:
                b |= 1 << (r * 7 + c)
    return b


def random_positions(orc, n_games, seed, stop_at_win_quirks=None):
    """Random playouts; yields every intermediate (state, player).  Play continues past missed
    wins (like the reference under Q1) until the board is full, so both colours can own lines."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_games):
        s = orc.init_board(1)
        p = np.array([1], np.int8)
        for _ply in range(42):
            v = orc.valid_moves(s)[0]
            if not v.any():
                break
            a = rng.choice(np.flatnonzero(v))
            s, p = orc.next_state(s, p, a)
            out.append((s.copy(), int(p[0])))
    return out


def test_bitboard_vs_oracle(oracle, shim):
    pos = random_positions(oracle, 60, 1234)
    states = np.concatenate([s for s, _ in pos])
    players = np.array([p for _, p in pos], np.int8)
    canon = oracle.canonical_form(states, players)
    keys = oracle.state_key(canon)
    valid = oracle.valid_moves(canon)
    code2val = {0: 0.0, 1: 1.0, 2: -1.0, 3: np.float32(1e-4)}
    for q in (0, 1):
        ended = oracle.game_ended(canon, 1, q)
        for i in range(len(canon)):
            cur, opp = bits(canon[i]["s"], 1), bits(canon[i]["s"], -1)
            assert code2val[shim.shim_game_ended_code(cur, opp, q)] == ended[i]
    assert len(set(keys.tolist())) == len({c["s"].tobytes() for c in canon})  # key is injective
    for i in range(len(canon)):
        cur, opp = bits(canon[i]["s"], 1), bits(canon[i]["s"], -1)
        assert shim.shim_state_key(cur, opp) == keys[i]
        vm = shim.shim_valid_mask(cur | opp)
        assert [(vm >> a) & 1 for a in range(7)] == valid[i].tolist()
        assert shim.shim_mirror(cur) == bits(canon[i]["s"][:, ::-1], 1)
        for a in np.flatnonzero(valid[i]):
            nxt, npl = oracle.next_state(canon[i:i + 1], 1, a)
            nc = oracle.canonical_form(nxt, npl)
            out = (C.c_uint64 * 2)()
            shim.shim_play_canonical(cur, opp, int(a), out)
            assert out[0] == bits(nc[0]["s"], 1) and out[1] == bits(nc[0]["s"], -1)


def test_counter_and_puct_vs_oracle(oracle, shim):
    rng = np.random.default_rng(5)
    c = oracle.L.azo_counter_init()
    vals = [1.0, -1.0, 0.0, -0.0, 1e-4, -1e-4, 0.5, -0.33, 0.999, -0.004, 1e9, -1e9, float("nan")]
    for q in (0, 4):
        cc = c
        for i in range(400):
            cc = oracle.L.azo_counter_visit(cc)
            v = vals[i % len(vals)] if i < 100 else float(np.float32(rng.uniform(-1, 1)))
            a = oracle.L.azo_counter_unvisit(cc, v, 100.0, q)
            b = shim.shim_unvisit(cc, v, q)
            assert a == b, (hex(cc), v, q)
            cc = a
            w, n, vl, qv = oracle.counter_read(cc)
            prior = float(np.float32(rng.uniform(0, 1)))
            pn = int(rng.integers(0, 40000))
            u_ref = np.float32(qv) + np.float32(np.float32(np.float32(1.0) * np.float32(prior)) *
                                                np.sqrt(np.float32(pn) + np.float32(1e-6), dtype=np.float32)) / np.float32(np.uint16(1 + n))
            assert np.float32(shim.shim_puct(cc, prior, pn, 1)).view(np.uint32) == np.float32(u_ref).view(np.uint32)
