"""arena::play_games (src/arena.rs:62-99) on the device vs the oracle: per-game results and the
Win/Loss/Draw tally of player A, bit-exact (two MCTS players, temp 0, a fresh tree pair per game)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("quirks", [0, 15])
@pytest.mark.parametrize("evals", [(0, 1), (1, 0), (1, 1)])
def test_arena_matches_oracle(azb, oracle, quirks, evals):
    ea, eb = evals
    num, sims = 12, 40
    for k_open in (0, 3):
        counts, res, st = azb.arena_play_games(num, ea, eb, k_open=k_open, num_sims=sims, quirks=quirks, seed=9)
        ocounts, ores = oracle.arena_play_games(num, ea, eb, num_sims=sims, quirks=quirks, seed=9, shared_trees=0, k_open=k_open)
        assert res.tolist() == ores.tolist(), (k_open, res, ores)
        assert list(counts) == [int(x) for x in ocounts]
        assert st["games"] == num and sum(counts) == num


def test_arena_example_size_and_odd_num(azb, oracle):
    # examples/connect_four.rs:64: 40 arena games; arena.rs:83 plays num/2 per seat order
    counts, res, st = azb.arena_play_games(41, 1, 0, num_sims=25, seed=1)
    ocounts, ores = oracle.arena_play_games(41, 1, 0, num_sims=25, seed=1, shared_trees=0)
    assert len(res) == 40 and res.tolist() == ores.tolist()
    assert list(counts) == [int(x) for x in ocounts]
    # with k_open = 0 and deterministic players every game of a seat order is identical
    assert len(set(res[:20].tolist())) == 1 and len(set(res[20:].tolist())) == 1


def test_arena_recycles_slots(azb, oracle):
    counts, res, st = azb.arena_play_games(24, 1, 0, k_open=4, num_sims=30, seed=3, max_concurrent_games=5)
    ocounts, ores = oracle.arena_play_games(24, 1, 0, num_sims=30, seed=3, shared_trees=0, k_open=4)
    assert res.tolist() == ores.tolist()
