"""K5/K7 parity: whole self-play games on the device (Coach::execute_episode, coach.rs:104-157)
vs the oracle: action sequences, per-ply root counts, exported SOA samples and labels."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 2], ids=["persistent", "rounds"])
def schedule(request):
    """1 = one persistent kernel (a warp plays a whole game), 2 = lock-step rounds."""
    return request.param


def check_games(azb, orc, coach, games, first_game_id, **okw):
    st = coach.self_play(games, first_game_id)
    tr = coach.traces()
    boards, pis, vs = coach.export_samples()
    assert st["games"] == games and st["samples"] == 2 * st["plies"] == len(vs)
    offs = np.concatenate([[0], np.cumsum(tr["plies"].astype(np.int64))])
    tot = np.zeros(6, np.uint64)
    check = range(games) if games <= 64 else list(range(0, games, max(1, games // 24)))
    for g in check:
        o = orc.execute_episode(episode_id=first_game_id + g, **okw)
        n = o["plies"]
        assert tr["plies"][g] == n, (g, tr["plies"][g], n)
        assert tr["actions"][g, :n].tolist() == o["actions"][:n].tolist(), g
        assert (tr["actions"][g, n:] == 0xFF).all()
        assert np.array_equal(tr["counts"][g, :n], o["counts"][:n]), g
        assert np.float32(tr["final_r"][g]) == np.float32(o["final_r"]) and tr["final_player"][g] == o["final_player"]
        a, b = 2 * offs[g], 2 * offs[g + 1]
        assert np.array_equal(boards[a:b], o["boards"]), g
        assert np.array_equal(pis[a:b].view(np.uint32), o["pis"].view(np.uint32)), g
        assert np.array_equal(vs[a:b].view(np.uint32), o["vs"].view(np.uint32)), g
        tot += o["stats"]
    if len(check) == games:
        assert [st[k] for k in ("sims", "levels", "expansions", "terminal_hits", "dup_links", "evals")] == tot.tolist()
    return st


@pytest.mark.parametrize("quirks", [0, 15])
@pytest.mark.parametrize("evaluator", [0, 1])
def test_example_config_25_sims(azb, oracle, quirks, evaluator, schedule):
    # examples/connect_four.rs:55-71: 25 sims/move, cpuct 1, temp_threshold 15, max_depth 1000
    coach = azb.Coach.setup("./checkpoint", 1000000, 0.6, 15, 20, 200000, 1, 1, 40, 1, 1, 25, 1, 1000, 1,
                            quirks=quirks, evaluator=evaluator, seed=1, schedule=schedule)
    check_games(azb, oracle, coach, 32, 0, num_sims=25, quirks=quirks, seed=1, evaluator=evaluator)


def test_baseline_config1_50_sims(azb, oracle, schedule):
    coach = azb.Coach(num_sims=50, seed=1, schedule=schedule)
    check_games(azb, oracle, coach, 1, 0, num_sims=50, quirks=0, seed=1, evaluator=0)
    b, p, v = coach.execute_episode(3)
    o = oracle.execute_episode(num_sims=50, seed=1, episode_id=3)
    assert np.array_equal(b, o["boards"]) and np.array_equal(v, o["vs"])


def test_more_games_than_resident_trees(azb, oracle, schedule):
    # 40 games on 8 resident trees: trees are recycled (table cleared) between games
    coach = azb.Coach(num_sims=60, seed=5, evaluator=1, max_concurrent_games=8, schedule=schedule)
    check_games(azb, oracle, coach, 40, 100, num_sims=60, quirks=0, seed=5, evaluator=1)


def test_sampled_800_sims(azb, oracle, schedule):
    # BASELINE config 2 parameters on a small batch: 800 sims/move, seed 0xA1FA0
    coach = azb.Coach(num_sims=800, seed=0xA1FA0, evaluator=0, schedule=schedule)
    st = check_games(azb, oracle, coach, 256, 0, num_sims=800, quirks=0, seed=0xA1FA0, evaluator=0)
    assert st["sims"] == 800 * st["plies"]


def test_oracle_fixture_on_device(azb, schedule):
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_episodes.json")))
    for e in fx["episodes"]:
        coach = azb.Coach(num_sims=e["num_sims"], quirks=e["quirks"], seed=e["seed"], evaluator=e["evaluator"], schedule=schedule,
                          num_sim_threads=e.get("num_sim_threads", 1))
        st = coach.self_play(1, e["episode_id"])
        tr = coach.traces()
        n = int(tr["plies"][0])
        assert tr["actions"][0, :n].tolist() == e["actions"]
        assert tr["counts"][0, :n].tolist() == e["counts"]
        _, _, vs = coach.export_samples()
        assert vs.tolist() == e["vs"]
        assert [st[k] for k in ("sims", "levels", "expansions", "terminal_hits", "dup_links", "evals")] == e["stats"]
        assert st["owners_max"] == e["seen_len"]


def test_full_size_properties(azb, schedule):
    """BASELINE config 2 at full width (4096 games) with fewer sims: size-independent
    properties — every game ends legally, samples are consistent, and the run is
    deterministic (same seed => identical traces)."""
    coach = azb.Coach(num_sims=48, seed=0xA1FA0, evaluator=1, schedule=schedule)
    st = coach.self_play(4096, 0)
    tr = coach.traces()
    boards, pis, vs = coach.export_samples()
    assert st["sims"] == 48 * st["plies"]
    assert (tr["plies"] >= 7).all() and (tr["plies"] <= 42).all()
    assert np.allclose(pis.sum(axis=1), 1.0, atol=1e-5)
    stones = boards.sum(axis=(1, 2, 3))
    # sample 2k and 2k+1 are mirror images with reversed pi
    assert np.array_equal(boards[0::2], boards[1::2][:, :, :, ::-1])
    assert np.array_equal(pis[0::2], pis[1::2][:, ::-1])
    offs = np.concatenate([[0], np.cumsum(tr["plies"].astype(np.int64))])
    assert np.array_equal(stones[2 * offs[:-1]], np.zeros(4096))  # every game starts on the empty board
    assert set(np.unique(np.abs(vs)).tolist()) <= {1.0, np.float32(1e-4)}
    st2 = coach.self_play(4096, 0)
    tr2 = coach.traces()
    assert np.array_equal(tr["actions"], tr2["actions"]) and st2["levels"] == st["levels"]


def test_pipelined_calls_equal_blocking_calls(azb, oracle):
    """azb_coach_self_play_begin / _end on two coaches used in turn (consecutive batches overlap on the device) give, batch by
    batch, what the blocking call gives; one batch is replayed by the oracle."""
    G, sims = 96, 60
    cs = [azb.Coach(num_sims=sims, seed=21, evaluator=azb.EVAL_HASH) for _ in range(2)]
    ref = azb.Coach(num_sims=sims, seed=21, evaluator=azb.EVAL_HASH)
    cs[0].self_play_begin(G, 0)
    for k in range(4):
        if k + 1 < 4:
            cs[(k + 1) % 2].self_play_begin(G, (k + 1) * G)
        st = cs[k % 2].self_play_end()
        tr = cs[k % 2].traces()
        b, p, v = cs[k % 2].export_samples()
        rst = ref.self_play(G, k * G)
        rtr = ref.traces()
        rb, rp, rv = ref.export_samples()
        for key in ("plies", "sims", "levels", "expansions", "terminal_hits", "dup_links", "evals"):
            assert st[key] == rst[key], (k, key)
        assert np.array_equal(tr["actions"], rtr["actions"]) and np.array_equal(tr["counts"], rtr["counts"])
        assert np.array_equal(b, rb) and np.array_equal(p, rp) and np.array_equal(v, rv)
    o = oracle.execute_episode(num_sims=sims, seed=21, episode_id=3 * G + 5, evaluator=oracle.EVAL_HASH)
    n = o["plies"]
    assert tr["plies"][5] == n and tr["actions"][5, :n].tolist() == o["actions"][:n].tolist()
    with pytest.raises(azb.AzbError):
        cs[0].self_play_end()  # nothing in flight
    cs[0].self_play_begin(G, 0)
    with pytest.raises(azb.AzbError):
        cs[0].self_play_begin(G, G)  # one call in flight per coach
    cs[0].self_play_end()
