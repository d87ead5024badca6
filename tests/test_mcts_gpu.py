"""K2-K4 parity: warp-per-tree search vs the oracle's AsyncMcts, bit-exact visit counts, raw
64-bit counters and the whole tree (every unique state's counter, terminal value and priors)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PROFILES = [0, 15, 5, 2, 1 | 4]


def compare_tree(m, o, tree=0):
    ka, ca, ea, pa, ha = m.dump(tree)
    kb, cb, eb, pb, hb = o.dump()
    assert len(ka) == len(kb)
    assert np.array_equal(ka, kb)
    assert np.array_equal(ca, cb), [(hex(k), hex(x), hex(y)) for k, x, y in zip(ka, ca, cb) if x != y][:5]
    assert np.array_equal(ea.view(np.uint32), eb.view(np.uint32))
    assert np.array_equal(ha, hb)
    assert np.array_equal(pa.view(np.uint32), pb.view(np.uint32))


@pytest.mark.parametrize("quirks", PROFILES)
@pytest.mark.parametrize("evaluator", [0, 1])
@pytest.mark.parametrize("sims", [1, 2, 25, 50, 800])
def test_root_search_matches_oracle(azb, oracle, quirks, evaluator, sims):
    root = oracle.init_board(1)
    m = azb.AsyncMcts(1, num_sims=sims, quirks=quirks, evaluator=evaluator, mcts_reserve_size=400000)
    o = oracle.Mcts(num_sims=sims, quirks=quirks, evaluator=evaluator)
    for temp in (1.0, 0.0):  # two consecutive searches on the same tree
        ca, pa = m.get_action_prob(root, temp)
        cb, pb = o.get_action_prob(root, temp)
        assert ca[0].tolist() == cb.tolist()
        if sims > 1:
            assert np.array_equal(pa[0].view(np.uint32), pb.view(np.uint32))
        assert int(m.counter_of(root)[0]) == o.counter_of(root)
    sa, sb = m.stats()[0], o.stats()
    assert sa[:6].tolist() == sb[:6].tolist()   # sims, levels, expansions, terminal hits, dup links, evals
    assert sa[7] == sb[7]                       # NodeStore.seen.len()
    compare_tree(m, o)


@pytest.mark.parametrize("temp", [0.5, 2.0, 0.25])
def test_generic_temperature_within_tolerance(azb, oracle, temp):
    """get_action_prob with a temperature other than 0 / 1 (async_mcts.rs:108-113: counts^(1/temp), normalised).  The
    coach only ever passes 1 or 0 (coach.rs:122-126); this form goes through powf on both sides (CUDA's vs libm's), so it
    is pinned by tolerance, not by bits: 2e-6 relative on every probability, the counts themselves stay bit-exact, and the
    vector against a float64 restatement from those counts."""
    root = oracle.init_board(1)
    m = azb.AsyncMcts(1, num_sims=200, quirks=0, evaluator=1, mcts_reserve_size=100000)
    o = oracle.Mcts(num_sims=200, quirks=0, evaluator=1)
    ca, pa = m.get_action_prob(root, temp)
    cb, pb = o.get_action_prob(root, temp)
    assert ca[0].tolist() == cb.tolist()
    want = ca[0].astype(np.float64) ** (1.0 / temp)
    want /= want.sum()
    assert np.abs(pa[0] - pb).max() <= 2e-6 * max(1.0, float(pb.max()))
    assert np.abs(pa[0] - want).max() <= 2e-6
    assert abs(float(pa[0].sum()) - 1.0) < 1e-5


def test_survey_golden_on_device(azb, oracle):
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_vectors.json")))
    root = oracle.init_board(1)
    for v in gold["root_counts"]:
        m = azb.AsyncMcts(1, num_sims=v["sims"], quirks=v["quirks"], mcts_reserve_size=400000)
        counts, _ = m.get_action_prob(root, 1.0)
        assert counts[0].tolist() == v["counts"]
        if "root_counter" in v:
            assert int(m.counter_of(root)[0]) == int(v["root_counter"], 16)
        if "dup_links" in v:
            assert m.stats()[0][4] == v["dup_links"]


@pytest.mark.parametrize("quirks", [0, 15])
def test_many_trees_walk_a_game(azb, oracle, quirks):
    """64 trees, each following its own random game: at every ply all trees search from their
    current canonical state (tree reuse, transposition links, F12 roots for unseen states)."""
    n, sims = 64, 120
    rng = np.random.default_rng(7)
    m = azb.AsyncMcts(n, num_sims=sims, quirks=quirks, evaluator=1, mcts_reserve_size=400000)
    os_ = [oracle.Mcts(num_sims=sims, quirks=quirks, evaluator=1) for _ in range(n)]
    s = oracle.init_board(n)
    p = np.ones(n, np.int8)
    done = np.zeros(n, bool)
    for ply in range(14):
        canon = oracle.canonical_form(s, p)
        ca, pa = m.get_action_prob(canon, 1.0)
        for i in range(n):
            cb, pb = os_[i].get_action_prob(canon[i:i + 1], 1.0)
            assert ca[i].tolist() == cb.tolist(), (ply, i)
            assert np.array_equal(pa[i].view(np.uint32), pb.view(np.uint32))
        v = oracle.valid_moves(s)
        # every third ply tree i plays a RANDOM legal move (often unvisited => new root, F12),
        # otherwise the most visited one
        a = np.array([rng.choice(np.flatnonzero(v[i])) if (ply + i) % 3 == 0 else int(np.argmax(ca[i]))
                      for i in range(n)], np.uint8)
        s, p = oracle.next_state(s, p, a)
        done |= oracle.game_ended(s, p, quirks) != 0
        if done.any():
            break
    roots = oracle.canonical_form(s, p)
    ctr = m.counter_of(roots)
    st = m.stats()
    for i in range(0, n, 7):
        assert int(ctr[i]) == os_[i].counter_of(roots[i:i + 1])
        assert st[i][:6].tolist() == os_[i].stats()[:6].tolist()
        compare_tree(m, os_[i], tree=i)


def test_pool_overflow_is_an_error(azb, oracle):
    m = azb.AsyncMcts(1, num_sims=400, mcts_reserve_size=70)  # 12 blocks
    with pytest.raises(azb.AzbError) as e:
        m.get_action_prob(oracle.init_board(1), 1.0)
    assert e.value.code == azb.ERR_CAPACITY


def test_fast_arithmetic_is_ieee_exact(azb):
    # the level loop's division / sqrt without slow-path calls vs __frcp_rn/__fsqrt_rn/__fdiv_rn
    assert azb.selftest_arith() == [0, 0, 0, 0]


@pytest.mark.parametrize("quirks", [0, 15])
@pytest.mark.parametrize("max_depth", [0, 2, 7])
def test_depth_limit_uses_the_generic_walk(azb, oracle, quirks, max_depth):
    """max_depth < 43 routes every simulation to the GENERIC variant of one_sim_impl (depth check at every
    level, __fdiv_rn, no speculative prefix): async_mcts.rs:241-244 with repair F6."""
    root = oracle.init_board(1)
    m = azb.AsyncMcts(1, num_sims=300, quirks=quirks, evaluator=1, max_depth=max_depth, mcts_reserve_size=400000)
    o = oracle.Mcts(num_sims=300, quirks=quirks, evaluator=1, max_depth=max_depth)
    for temp in (1.0, 0.0):
        ca, pa = m.get_action_prob(root, temp)
        cb, pb = o.get_action_prob(root, temp)
        assert ca[0].tolist() == cb.tolist()
        assert np.array_equal(pa[0].view(np.uint32), pb.view(np.uint32))
    assert m.stats()[0][:6].tolist() == o.stats()[:6].tolist()
    compare_tree(m, o)


def test_visit_counts_past_the_16_bit_wrap(azb, oracle):
    """Quirk Q6: N is 16 bits and wraps into W (node.rs:17).  Trees whose visit counts may reach the wrap switch
    from the hot variant (speculative prefix, FMA division) to the GENERIC one in the middle of their life: the
    result must stay bit-identical with the oracle across the switch and across the wrap itself."""
    root = oracle.init_board(1)
    sims = 22000
    m = azb.AsyncMcts(1, num_sims=sims, quirks=0, evaluator=0, mcts_reserve_size=1000000)
    o = oracle.Mcts(num_sims=sims, quirks=0, evaluator=0, reserve=1000000)
    for call in range(4):  # 88 000 simulations through the same root: its N wraps at 65 536
        ca, pa = m.get_action_prob(root, 1.0)
        cb, pb = o.get_action_prob(root, 1.0)
        assert ca[0].tolist() == cb.tolist(), call
        assert int(m.counter_of(root)[0]) == o.counter_of(root), call
    assert m.stats()[0][:6].tolist() == o.stats()[:6].tolist()
    compare_tree(m, o)
