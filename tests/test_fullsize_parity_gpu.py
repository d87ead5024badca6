"""Parity at the BASELINE configurations' REAL sizes ("what's weak" 1-2 and "missing" 1):

* config 2: 4096 games x 800 sims on the persistent kernel, sampled games replayed by the oracle;
* config 3: 8192 games x 400 sims with the ResNet-6x128 tensor-core evaluator, in-round de-duplication and the evaluation
  cache ON, sampled games replayed by the oracle with the same network as its predict callback, and the whole run hashed
  against a run with both switched off (what scripts/dedup_ab.sh did outside pytest);
* config 4 as specified: two DIFFERENT networks head to head, per-game results, action traces and root counts against the
  oracle's arena (arena.rs:7-99, coach.rs:356-372) with the networks as predict callbacks.

Size-dependent races are the realistic risk here (one was found by luck in round 1), so these run the full widths."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def replay(oracle, tr, games, **okw):
    for g in games:
        o = oracle.execute_episode(episode_id=g, **okw)
        n = o["plies"]
        assert tr["plies"][g] == n, (g, tr["plies"][g], n)
        assert tr["actions"][g, :n].tolist() == o["actions"][:n].tolist(), g
        assert np.array_equal(tr["counts"][g, :n], o["counts"][:n]), g


def test_config2_full_size_sampled_games(azb, oracle):
    """4096 x 800, seed 0xA1FA0, uniform evaluator: 24 games spread over the batch (first, last, every 178th) and the
    longest game of the batch, replayed bit for bit; every game's sims = 800 x plies."""
    coach = azb.Coach(num_sims=800, seed=0xA1FA0, evaluator=azb.EVAL_UNIFORM)
    st = coach.self_play(4096, 0)
    tr = coach.traces()
    assert st["games"] == 4096 and st["sims"] == 800 * st["plies"] and st["trees_resident"] == 4096
    games = sorted(set(list(range(0, 4096, 178)) + [4095, int(np.argmax(tr["plies"])), int(np.argmin(tr["plies"]))]))
    assert len(games) >= 24
    replay(oracle, tr, games, num_sims=800, quirks=0, seed=0xA1FA0, evaluator=oracle.EVAL_UNIFORM)
    # the exported samples of the sampled games
    boards, pis, vs = coach.export_samples()
    offs = np.concatenate([[0], np.cumsum(tr["plies"].astype(np.int64))])
    for g in games[:6]:
        o = oracle.execute_episode(episode_id=g, num_sims=800, quirks=0, seed=0xA1FA0, evaluator=oracle.EVAL_UNIFORM)
        a, b = 2 * offs[g], 2 * offs[g + 1]
        assert np.array_equal(boards[a:b], o["boards"]) and np.array_equal(vs[a:b], o["vs"])
        assert np.array_equal(pis[a:b].view(np.uint32), o["pis"].view(np.uint32))


_C3_SCRIPT = (
    "import importlib, sys, hashlib, numpy as np\n"
    f"sys.path.insert(0, {ROOT!r})\n"
    "azb = importlib.import_module('alphazero-rs_b200')\n"
    "net = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC)\n"
    "coach = azb.Coach(nnet=net, num_sims=400, seed=0xA1FA0, evaluator=azb.EVAL_NNET)\n"
    "st = coach.self_play(8192, 0)\n"
    "tr = coach.traces()\n"
    "h = hashlib.sha256()\n"
    "for k in ('actions', 'counts', 'plies'): h.update(np.ascontiguousarray(tr[k]).tobytes())\n"
    "print('HASH', h.hexdigest(), st['plies'], st['sims'], st['levels'], st['expansions'], st['evals'], st['nn_positions'], st['nn_cache_hits'])\n")


def test_config3_full_size_sampled_games_and_dedup_ab(azb, oracle):
    """8192 x 400 with the 6-block tensor-core network, de-duplication + cache on (the defaults): 8 sampled games replayed
    by the oracle searching with net.predict; then the same call in a child process with both mechanisms OFF must hash
    to the same traces (actions, root counts, plies of all 8192 games) and the same search statistics."""
    net = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC)
    coach = azb.Coach(nnet=net, num_sims=400, seed=0xA1FA0, evaluator=azb.EVAL_NNET)
    st = coach.self_play(8192, 0)
    tr = coach.traces()
    assert st["games"] == 8192 and st["sims"] == 400 * st["plies"]
    assert st["nn_positions"] + st["nn_cache_hits"] <= st["evals"] and st["nn_cache_hits"] > 0
    games = [0, 1171, 2342, 3513, 4684, 5855, 7026, 8191]
    replay(oracle, tr, games, num_sims=400, quirks=0, seed=0xA1FA0, evaluator=oracle.EVAL_CALLBACK, callback=net.predict)
    h = hashlib.sha256()
    for k in ("actions", "counts", "plies"):
        h.update(np.ascontiguousarray(tr[k]).tobytes())
    env = dict(os.environ, AZB200_LEAF_DEDUP="0", AZB200_EVAL_CACHE="0")
    r = subprocess.run([sys.executable, "-c", _C3_SCRIPT], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("HASH")][0].split()
    assert line[1] == h.hexdigest()
    assert [int(x) for x in line[2:7]] == [st[k] for k in ("plies", "sims", "levels", "expansions", "evals")]
    assert int(line[7]) == st["evals"] and int(line[8]) == 0  # without either, every evaluation is a network row


@pytest.mark.parametrize("k_open", [0, 3])
def test_config4_two_networks_match_oracle(azb, oracle, k_open):
    """arena::play_games with two different 6-block networks (seeds 7 and 8, the BASELINE config 4 players) on the device
    vs the oracle's arena with the two networks as predict callbacks: per-game results, plies, action traces and the root
    counts of every searched ply, 32 games (16 per seat order), Win/Loss/Draw tally."""
    a = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC)
    b = azb.NNet(seed=8, blocks=6, precision=azb.NNET_BF16_TC)
    num, sims = 32, 48
    counts, res, st, tr = azb.arena_play_games_traced(num, azb.EVAL_NNET, azb.EVAL_NNET, a, b, k_open=k_open, num_sims=sims,
                                                      seed=11)
    oc, ores, otr = oracle.arena_play_games_traced(num, oracle.EVAL_CALLBACK, oracle.EVAL_CALLBACK, a.predict, b.predict,
                                                   num_sims=sims, seed=11, k_open=k_open, shared_trees=0)
    assert res.tolist() == ores.tolist()
    assert list(counts) == [int(x) for x in oc] and sum(counts) == num
    assert tr["plies"].tolist() == otr["plies"].tolist()
    assert np.array_equal(tr["actions"], otr["actions"])
    assert np.array_equal(tr["counts"], otr["counts"])
    if k_open:
        assert len({tuple(r) for r in tr["actions"].tolist()}) > 4  # the openings diversify the games


def test_arena_shared_trees_match_oracle(azb, oracle):
    """The reference's own layout (coach.rs:333-372): pmcts / nmcts are created ONCE and keep growing through all games
    of the match, which are played strictly one after the other.  shared_trees=1 on the device vs the oracle's shared
    mode: results, traces and root counts (later games see the earlier games' statistics, so the games differ)."""
    for ea, eb in ((1, 0), (1, 1)):
        counts, res, st, tr = azb.arena_play_games_traced(10, ea, eb, shared_trees=1, num_sims=40, seed=3,
                                                          mcts_reserve_size=400000)
        oc, ores, otr = oracle.arena_play_games_traced(10, ea, eb, num_sims=40, seed=3, shared_trees=1, reserve=400000)
        assert res.tolist() == ores.tolist()
        assert list(counts) == [int(x) for x in oc]
        assert np.array_equal(tr["actions"], otr["actions"]) and np.array_equal(tr["counts"], otr["counts"])
    a = azb.NNet(seed=7, blocks=1, precision=azb.NNET_BF16_TC)
    b = azb.NNet(seed=8, blocks=1, precision=azb.NNET_BF16_TC)
    counts, res, st, tr = azb.arena_play_games_traced(6, azb.EVAL_NNET, azb.EVAL_NNET, a, b, shared_trees=1, num_sims=25,
                                                      seed=4, mcts_reserve_size=400000)
    oc, ores, otr = oracle.arena_play_games_traced(6, oracle.EVAL_CALLBACK, oracle.EVAL_CALLBACK, a.predict, b.predict,
                                                   num_sims=25, seed=4, shared_trees=1, reserve=400000)
    assert res.tolist() == ores.tolist() and np.array_equal(tr["actions"], otr["actions"])
    assert np.array_equal(tr["counts"], otr["counts"])
