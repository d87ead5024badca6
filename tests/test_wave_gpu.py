"""Tree-parallel search with virtual loss, num_sim_threads = K > 1 (src/async_mcts.rs:191-217, src/node.rs:77-92,359-365).

The reference runs K OS threads whose interleaving is the scheduler's: its own results are not reproducible.  The device
(csrc/mcts.cuh wave_*) and the oracle (oracle/mcts.hpp search_wave) run the K threads in the SAME fixed interleaving —
waves of K walks, then the wave's evaluations, then K backups — so parity is bit-exact here too: root counts, raw
counters, the whole tree, whole games, with fused evaluators and with the batched network (up to K leaves per tree and
round).  Deterministic mode (K = 1) stays what every other parity test pins."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def compare_tree(m, o, tree=0):
    ka, ca, ea, pa, ha = m.dump(tree)
    kb, cb, eb, pb, hb = o.dump()
    assert np.array_equal(ka, kb)
    assert np.array_equal(ca, cb), [(hex(k), hex(x), hex(y)) for k, x, y in zip(ka, ca, cb) if x != y][:5]
    assert np.array_equal(ea.view(np.uint32), eb.view(np.uint32))
    assert np.array_equal(ha, hb) and np.array_equal(pa.view(np.uint32), pb.view(np.uint32))


@pytest.mark.parametrize("quirks", [0, 15, 2])
@pytest.mark.parametrize("evaluator", [0, 1])
@pytest.mark.parametrize("k,sims", [(2, 2), (2, 50), (4, 8), (4, 200), (8, 8), (8, 800)])
def test_root_search_waves_match_oracle(azb, oracle, quirks, evaluator, k, sims):
    root = oracle.init_board(1)
    m = azb.AsyncMcts(1, num_sims=sims, quirks=quirks, evaluator=evaluator, mcts_reserve_size=400000, num_sim_threads=k)
    o = oracle.Mcts(num_sims=sims, quirks=quirks, evaluator=evaluator, num_sim_threads=k)
    for temp in (1.0, 0.0):  # two consecutive searches on the same tree (the first wave of the first shares the F1 root)
        ca, pa = m.get_action_prob(root, temp)
        cb, pb = o.get_action_prob(root, temp)
        assert ca[0].tolist() == cb.tolist()
        assert int(m.counter_of(root)[0]) == o.counter_of(root)
    sa, sb = m.stats()[0], o.stats()
    assert sa[:6].tolist() == sb[:6].tolist() and sa[7] == sb[7]
    compare_tree(m, o)


def test_waves_differ_from_deterministic_mode_but_visit_the_same_total(azb, oracle):
    root = oracle.init_board(1)
    c1, _ = azb.AsyncMcts(1, num_sims=400, evaluator=1, mcts_reserve_size=400000).get_action_prob(root, 1.0)
    c4, _ = azb.AsyncMcts(1, num_sims=400, evaluator=1, mcts_reserve_size=400000, num_sim_threads=4).get_action_prob(root, 1.0)
    assert c1[0].tolist() != c4[0].tolist()           # virtual loss spreads the walks of a wave
    assert abs(int(c1[0].sum()) - int(c4[0].sum())) <= 4  # sims minus the root evaluations
    with pytest.raises(azb.AzbError):
        azb.AsyncMcts(1, num_sims=10, num_sim_threads=4)   # async_mcts.rs:192: num_sims % num_threads == 0
    with pytest.raises(azb.AzbError):
        azb.AsyncMcts(1, num_sims=32, num_sim_threads=16)  # at most 8 walks per wave


@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("k", [2, 4, 8])
def test_selfplay_waves_match_oracle(azb, oracle, k, schedule):
    coach = azb.Coach(num_sims=48, seed=5, evaluator=azb.EVAL_HASH, num_sim_threads=k, schedule=schedule)
    st = coach.self_play(24, 7)
    tr = coach.traces()
    boards, pis, vs = coach.export_samples()
    offs = np.concatenate([[0], np.cumsum(tr["plies"].astype(np.int64))])
    tot = np.zeros(6, np.uint64)
    for g in range(24):
        o = oracle.execute_episode(num_sims=48, seed=5, episode_id=7 + g, evaluator=oracle.EVAL_HASH, num_sim_threads=k)
        n = o["plies"]
        assert tr["plies"][g] == n and tr["actions"][g, :n].tolist() == o["actions"][:n].tolist(), g
        assert np.array_equal(tr["counts"][g, :n], o["counts"][:n]), g
        a, b = 2 * offs[g], 2 * offs[g + 1]
        assert np.array_equal(pis[a:b].view(np.uint32), o["pis"].view(np.uint32)) and np.array_equal(vs[a:b], o["vs"])
        tot += o["stats"]
    assert [st[x] for x in ("sims", "levels", "expansions", "terminal_hits", "dup_links", "evals")] == tot.tolist()


@pytest.mark.parametrize("k", [2, 4])
def test_network_selfplay_waves_match_oracle(azb, oracle, k):
    """The batched evaluator with K leaves per tree and round (de-duplication and the evaluation cache on): games replayed
    by the oracle's wave mode with the same network as its predict callback."""
    net = azb.NNet(seed=7, blocks=2, precision=azb.NNET_BF16_TC)
    coach = azb.Coach(nnet=net, num_sims=32, seed=3, evaluator=azb.EVAL_NNET, num_sim_threads=k)
    st = coach.self_play(12, 0)
    tr = coach.traces()
    assert st["games"] == 12 and st["sims"] == 32 * st["plies"]
    for g in (0, 5, 11):
        o = oracle.execute_episode(num_sims=32, seed=3, episode_id=g, evaluator=oracle.EVAL_CALLBACK, callback=net.predict,
                                   num_sim_threads=k)
        n = o["plies"]
        assert tr["plies"][g] == n and tr["actions"][g, :n].tolist() == o["actions"][:n].tolist(), g
        assert np.array_equal(tr["counts"][g, :n], o["counts"][:n]), g
    # fewer rounds than deterministic mode needs for the same number of simulations
    c1 = azb.Coach(nnet=net, num_sims=32, seed=3, evaluator=azb.EVAL_NNET)
    s1 = c1.self_play(12, 0)
    assert st["launches"] < s1["launches"]


def test_arena_waves_match_oracle(azb, oracle):
    counts, res, st, tr = azb.arena_play_games_traced(8, 1, 0, k_open=2, num_sims=40, seed=9, num_sim_threads=4)
    oc, ores, otr = oracle.arena_play_games_traced(8, 1, 0, num_sims=40, seed=9, k_open=2, num_sim_threads=4)
    assert res.tolist() == ores.tolist() and np.array_equal(tr["actions"], otr["actions"])
    assert np.array_equal(tr["counts"], otr["counts"])
