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
                                    episode_id=e["episode_id"], evaluator=e["evaluator"])
        n = ep["plies"]
        assert ep["actions"][:n].tolist() == e["actions"]
        assert ep["counts"][:n].tolist() == e["counts"]
        assert ep["vs"].tolist() == e["vs"]
        assert [int(x) for x in ep["stats"]] == e["stats"]
        assert np.float32(ep["final_r"]) == np.float32(e["final_r"]) and ep["final_player"] == e["final_player"]
