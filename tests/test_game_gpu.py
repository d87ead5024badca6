"""K1 parity: the batched bitboard kernels behind the Game trait vs the oracle, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def all_positions_upto(orc, depth):
    """Exhaustive: every action sequence of length <= depth from the empty board."""
    states = [orc.init_board(1)]
    players = [np.array([1], np.int8)]
    frontier_s, frontier_p = states[0], players[0]
    for _ in range(depth):
        n = len(frontier_s)
        s = np.repeat(frontier_s, 7)
        p = np.repeat(frontier_p, 7)
        a = np.tile(np.arange(7, dtype=np.uint8), n)
        ok = orc.valid_moves(s)[np.arange(len(s)), a] == 1
        s, p, a = s[ok], p[ok], a[ok]
        frontier_s, frontier_p = orc.next_state(s, p, a)
        states.append(frontier_s)
        players.append(frontier_p)
    return np.concatenate(states), np.concatenate(players)


def random_playout_positions(orc, n_games, seed):
    rng = np.random.default_rng(seed)
    s = orc.init_board(n_games)
    p = np.ones(n_games, np.int8)
    out_s, out_p = [s.copy()], [p.copy()]
    for _ in range(42):
        v = orc.valid_moves(s)
        alive = v.any(axis=1)
        if not alive.any():
            break
        s, p = s[alive], p[alive]
        v = v[alive]
        a = np.array([rng.choice(np.flatnonzero(r)) for r in v], np.uint8)
        s, p = orc.next_state(s, p, a)
        out_s.append(s.copy())
        out_p.append(p.copy())
    return np.concatenate(out_s), np.concatenate(out_p)


def check_all_ops(azb, orc, s, p, seed=0):
    G = azb.ConnectFourGame
    rng = np.random.default_rng(seed)
    n = len(s)
    assert np.array_equal(G.get_valid_moves(s), orc.valid_moves(s))
    for q in (0, 1):
        for pl in (p, -p, np.ones(n, np.int8)):
            a = G.get_game_ended(s, pl, q)
            b = orc.game_ended(s, pl, q)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    v = orc.valid_moves(s)
    has = v.any(axis=1)
    act = np.array([rng.choice(np.flatnonzero(r)) if r.any() else 0 for r in v], np.uint8)
    ns_a, np_a = G.get_next_state(s[has], p[has], act[has])
    ns_b, np_b = orc.next_state(s[has], p[has], act[has])
    assert ns_a.tobytes() == ns_b.tobytes() and np.array_equal(np_a, np_b)
    ca, cb = G.get_canonical_form(s, p), orc.canonical_form(s, p)
    assert ca.tobytes() == cb.tobytes()
    pi = rng.random((n, 7), dtype=np.float32)
    sa, pa = G.get_symmetries(ca, pi)
    sb, pb = orc.symmetries(cb, pi)
    assert sa.tobytes() == sb.tobytes() and np.array_equal(pa.view(np.uint32), pb.view(np.uint32))
    for st in (s, ca, sa.reshape(-1)):
        assert np.array_equal(G.to_features(st), orc.to_features(st))
    s2 = s.copy()
    s2["me"] = -1
    assert np.array_equal(G.to_features(s2), orc.to_features(s2))
    assert (G.eval_heuristic(s) == 0).all()


def test_exhaustive_upto_6_plies(azb, oracle):
    s, p = all_positions_upto(oracle, 6)
    assert len(s) == sum(7 ** k for k in range(7))
    check_all_ops(azb, oracle, s, p)


def test_random_playouts_to_full_board(azb, oracle):
    # playouts run past (missed) wins until the board is full: both colours own lines, which
    # exercises the "first line in scan order decides" semantics
    s, p = random_playout_positions(oracle, 3000, 99)
    assert len(s) > 100000
    check_all_ops(azb, oracle, s, p, seed=1)


def test_reference_diagonal_kat(azb):
    # connect_four_game.rs:244-264
    G = azb.ConnectFourGame
    s, p = G.get_init_board(1), np.array([1], np.int8)
    for a in [0, 1, 1, 2, 0, 2, 2, 3, 3, 3, 3]:
        s, p = G.get_next_state(s, p, a)
    for q in (0, 1):
        assert G.get_game_ended(s, 1, q)[0] == 1.0


def test_empty_and_invalid_inputs(azb):
    G = azb.ConnectFourGame
    assert G.get_valid_moves(np.zeros(0, azb.STATE_DTYPE)).shape == (0, 7)
    s, p = G.get_init_board(1), np.array([1], np.int8)
    for _ in range(6):
        s, p = G.get_next_state(s, p, 3)
    assert G.get_valid_moves(s)[0].tolist() == [1, 1, 1, 0, 1, 1, 1]
    with pytest.raises(azb.AzbError) as e:  # the reference panics on a full column (:95-99)
        G.get_next_state(s, p, 3)
    assert e.value.code == azb.ERR_INVALID
