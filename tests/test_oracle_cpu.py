"""The oracle against every golden vector available for the path (CPU only).

1. the reference's own 14 unit tests, restated (oracle/test_reference_units.cpp);
2. SURVEY App. D vectors from an independent emulation (tests/golden/survey_vectors.json);
3. the oracle's own committed episode fixtures (tests/golden/oracle_episodes.json) — guards
   against silent drift of the oracle between rounds.
"""
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "survey_vectors.json")))


def play(orc, plies):
    s = orc.init_board(1)
    p = np.array([1], np.int8)
    for a in plies:
        s, p = orc.next_state(s, p, a)
    return s, p


def test_reference_unit_tests_restated(oracle):
    r = subprocess.run([oracle.UNIT_BIN], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "OK 14" in r.stdout


@pytest.mark.parametrize("w", GOLD["game_ended_witnesses"])
def test_game_ended_witnesses(oracle, w):
    s, _ = play(oracle, w["plies"])
    assert oracle.game_ended(s, w["player"], w["quirks"])[0] == np.float32(w["expect"])


@pytest.mark.parametrize("v", GOLD["root_counts"])
def test_survey_root_counts(oracle, v):
    m = oracle.Mcts(num_sims=v["sims"], quirks=v["quirks"])
    root = oracle.init_board(1)
    counts, _ = m.get_action_prob(root, 1.0)
    assert counts.tolist() == v["counts"]
    if "root_counter" in v:
        assert m.counter_of(root) == int(v["root_counter"], 16)
    st = m.stats()
    if "slots" in v:
        assert st[6] == v["slots"]
    if "dup_links" in v:
        assert st[4] == v["dup_links"]
    if "terminal_hits" in v:
        assert st[3] == v["terminal_hits"]
    if "evals" in v:
        assert st[5] == v["evals"]


@pytest.mark.parametrize("g", GOLD["temp0_games"])
def test_survey_temp0_games(oracle, g):
    # temp = 0 on every ply (temp_threshold = 0 => episode_step < 0 never holds)
    ep = oracle.execute_episode(num_sims=g["sims"], quirks=g["quirks"], temp_threshold=0)
    assert ep["actions"][: ep["plies"]].tolist() == g["actions"]
    assert ep["final_r"] == np.float32(g["result"])


def test_survey_inherited_counts(oracle):
    ep = oracle.execute_episode(num_sims=25, quirks=5, temp_threshold=0)
    assert ep["counts"][:3].tolist() == GOLD["alternate25_first_counts"]


def test_counter_arithmetic(oracle):
    # SURVEY App. B.2 witnesses at WIN_SCALE = 100
    c0 = oracle.L.azo_counter_init()
    assert c0 == 0x7FFFFFFF00000000
    c = oracle.L.azo_counter_visit(c0)
    assert c == 0x7FFFFFFF00010001
    assert oracle.L.azo_counter_unvisit(c, -1e-4, 100.0, 15) == 0x7FFFFFFF00010000  # incr = 0, v < 0
    assert oracle.L.azo_counter_unvisit(c, 0.0, 100.0, 15) == 0x8000000000010000    # +1 slip
    assert oracle.L.azo_counter_unvisit(c, 0.0, 100.0, 0) == 0x7FFFFFFF00010000
    assert oracle.L.azo_counter_unvisit(c, 1.0, 100.0, 15) == 0x7FFFFFFF00010000 + (101 << 32)
    assert oracle.L.azo_counter_unvisit(c, -1.0, 100.0, 15) == 0x7FFFFFFF00010000 - (100 << 32)
    w, n, vl, q = oracle.counter_read(0x7FFFFFFF00010000 - (100 << 32))
    assert (w, n, vl, q) == (-1.0, 1, 0, -1.0)


def test_symmetries_and_features(oracle):
    s, p = play(oracle, [0, 1, 1, 2])
    c = oracle.canonical_form(s, -1)
    assert (c["s"] == -s["s"]).all() and c["me"][0] == 1
    pi = np.arange(7, dtype=np.float32)[None]
    ss, pp = oracle.symmetries(c, pi)
    assert (ss[0, 1]["s"] == c[0]["s"][:, ::-1]).all()
    assert pp[0, 1].tolist() == pi[0, ::-1].tolist()
    f = oracle.to_features(c)
    assert f.shape == (1, 2, 6, 7)
    assert (f[0, 0] == (c[0]["s"] == 1)).all() and (f[0, 1] == (c[0]["s"] == -1)).all()


def test_oracle_episode_fixture(oracle):
    fx = json.load(open(os.path.join(HERE, "golden", "oracle_episodes.json")))
    for e in fx["episodes"]:
        ep = oracle.execute_episode(num_sims=e["num_sims"], quirks=e["quirks"], seed=e["seed"],
                                    episode_id=e["episode_id"], evaluator=e["evaluator"],
                                    num_sim_threads=e.get("num_sim_threads", 1))
        n = ep["plies"]
        assert ep["actions"][:n].tolist() == e["actions"]
        assert ep["counts"][:n].tolist() == e["counts"]
        assert ep["vs"].tolist() == e["vs"]
        assert [int(x) for x in ep["stats"]] == e["stats"]
        assert np.float32(ep["final_r"]) == np.float32(e["final_r"]) and ep["final_player"] == e["final_player"]


def test_cpu_network_matches_torch_restatement(azb, oracle):
    """oracle/nnet_cpu.hpp (the plain C++ fp32 forward behind the CPU-baseline legs of BASELINE configs 1 and 3) against an
    independent torch float64 restatement of the same architecture on the same random parameters: 1e-4, the tolerance the
    north star states for fp32.  azb.param_layout is host-only Python (no device call)."""
    import torch
    import torch.nn.functional as F
    blocks = 2
    L = azb.param_layout(blocks)
    rng = np.random.default_rng(5)
    params = (rng.standard_normal(L["total"]) * 0.05).astype(np.float32)
    boards = (rng.random((6, 2, 6, 7)) < 0.3).astype(np.float32)
    boards[:, 1] *= 1.0 - boards[:, 0]

    def get(name):
        o, shape = L[name]
        return torch.from_numpy(params[o:o + int(np.prod(shape))].reshape(shape).copy()).double()

    x = torch.from_numpy(boards).double()
    x = F.relu(F.conv2d(x, get("stem_w").reshape(3, 3, 2, 128).permute(3, 2, 0, 1), get("stem_b"), padding=1))
    tw, tb = get("tower_w"), get("tower_b")
    for b in range(blocks):
        y = F.relu(F.conv2d(x, tw[2 * b].reshape(3, 3, 128, 128).permute(3, 2, 0, 1), tb[2 * b], padding=1))
        y = F.conv2d(y, tw[2 * b + 1].reshape(3, 3, 128, 128).permute(3, 2, 0, 1), tb[2 * b + 1], padding=1)
        x = F.relu(x + y)
    pol = F.relu(torch.einsum("bchw,cp->bphw", x, get("pol_w")) + get("pol_b").view(1, 2, 1, 1)).reshape(len(boards), 84)
    rpi = torch.softmax(pol @ get("pol_fc_w") + get("pol_fc_b"), dim=1).numpy()
    val = F.relu(torch.einsum("bchw,c->bhw", x, get("val_w")) + get("val_b")).reshape(len(boards), 42)
    rv = torch.tanh(F.relu(val @ get("val_fc1_w") + get("val_fc1_b")) @ get("val_fc2_w") + get("val_fc2_b")).numpy()
    pi, v = oracle.cpu_net_predict(params, blocks, boards)
    assert np.abs(pi - rpi).max() < 1e-4 and np.abs(v - rv).max() < 1e-4
    assert np.abs(pi - pi[0]).max() > 1e-4  # different positions, different outputs
    # the timing leg built on it runs and counts what it did
    r = oracle.bench_selfplay_net(params, blocks, 2, 2, num_sims=6, max_plies=2, seed=3)
    assert r["sims"] == 2 * 2 * 6 and r["plies"] == 4 and r["evals"] > 0


def test_wave_mode_invariants(oracle):
    """Tree-parallel search (num_sim_threads = K, oracle/mcts.hpp search_wave): every wave leaves no virtual loss behind
    (VL field 0, N = visits), the root is visited once per simulation, K = 1 is the deterministic search itself, and
    num_sims % K != 0 is refused (async_mcts.rs:192)."""
    root = oracle.init_board(1)
    for k in (1, 2, 4, 8):
        m = oracle.Mcts(num_sims=64, quirks=0, evaluator=oracle.EVAL_HASH, num_sim_threads=k)
        counts, _ = m.get_action_prob(root, 1.0)
        c = m.counter_of(root)
        assert c & 0xFFFF == 0 and (c >> 16) & 0xFFFF == 64          # VL back to 0, N(root) = 64 visits
        keys, counters, e, p, hp = m.dump()
        assert (counters & np.uint64(0xFFFF) == 0).all()
        assert int(counts.sum()) <= 64 - 1
        if k == 1:
            ref = oracle.Mcts(num_sims=64, quirks=0, evaluator=oracle.EVAL_HASH)
            rc, _ = ref.get_action_prob(root, 1.0)
            assert counts.tolist() == rc.tolist()
    with pytest.raises(RuntimeError):
        oracle.Mcts(num_sims=10, quirks=0, evaluator=0, num_sim_threads=4).get_action_prob(root, 1.0)
