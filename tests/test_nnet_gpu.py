"""NNet::predict (src/nnet.rs:40-44) and network-evaluated self-play.

* fp32 path vs a torch fp32 restatement of the same architecture with the same weights: 1e-4
  (the tolerance the north star states for fp32).
* bf16 tensor-core path vs the fp32 path: stated bf16 tolerance.
* self-play / arena with evaluator NNET (lock-step rounds, batched leaves) vs the oracle whose
  evaluator callback is this very network: bit-exact visit counts and trajectories.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def random_features(oracle, n_games=40, seed=3):
    rng = np.random.default_rng(seed)
    feats = []
    for _ in range(n_games):
        s = oracle.init_board(1)
        p = np.array([1], np.int8)
        for _ply in range(int(rng.integers(0, 43))):
            v = oracle.valid_moves(s)[0]
            if not v.any():
                break
            s, p = oracle.next_state(s, p, rng.choice(np.flatnonzero(v)))
            feats.append(oracle.to_features(oracle.canonical_form(s, p))[0])
    feats.append(np.zeros((2, 6, 7), np.float32))
    return np.stack(feats)


def torch_forward(azb, params, blocks, boards):
    import torch
    import torch.nn.functional as F
    L = azb.param_layout(blocks)

    def get(name):
        o, shape = L[name]
        return torch.from_numpy(params[o:o + int(np.prod(shape))].reshape(shape).copy()).double()

    x = torch.from_numpy(boards).double()
    # stem_w [9][2][C] -> conv weight [C][2][3][3]
    w = get("stem_w").reshape(3, 3, 2, 128).permute(3, 2, 0, 1)
    x = F.relu(F.conv2d(x, w, get("stem_b"), padding=1))
    tw, tb = get("tower_w"), get("tower_b")
    for b in range(blocks):
        w1 = tw[2 * b].reshape(3, 3, 128, 128).permute(3, 2, 0, 1)
        w2 = tw[2 * b + 1].reshape(3, 3, 128, 128).permute(3, 2, 0, 1)
        y = F.relu(F.conv2d(x, w1, tb[2 * b], padding=1))
        y = F.conv2d(y, w2, tb[2 * b + 1], padding=1)
        x = F.relu(x + y)
    pol = F.relu(torch.einsum("bchw,cp->bphw", x, get("pol_w")) + get("pol_b").view(1, 2, 1, 1)).reshape(len(boards), 84)
    logits = pol @ get("pol_fc_w") + get("pol_fc_b")
    pi = torch.softmax(logits, dim=1)
    val = F.relu(torch.einsum("bchw,c->bhw", x, get("val_w")) + get("val_b")).reshape(len(boards), 42)
    h = F.relu(val @ get("val_fc1_w") + get("val_fc1_b"))
    v = torch.tanh(h @ get("val_fc2_w") + get("val_fc2_b"))
    return pi.numpy(), v.numpy()


@pytest.mark.parametrize("blocks", [1, 6])
def test_fp32_path_matches_torch(azb, oracle, blocks):
    net = azb.NNet(seed=7, blocks=blocks, precision=azb.NNET_FP32)
    feats = random_features(oracle)
    pi, v = net.predict(feats)
    rpi, rv = torch_forward(azb, net.get_params(), blocks, feats)
    assert np.abs(pi - rpi).max() < 1e-4, np.abs(pi - rpi).max()
    assert np.abs(v - rv).max() < 1e-4, np.abs(v - rv).max()
    assert np.allclose(pi.sum(1), 1.0, atol=1e-5)
    # non-degenerate: different positions give different outputs
    assert np.abs(pi - pi[0]).max() > 1e-3 and np.ptp(v) > 1e-3


def test_predict_is_batch_independent(azb, oracle):
    net = azb.NNet(seed=8, blocks=2, precision=azb.NNET_FP32)
    feats = random_features(oracle, 10)
    pi, v = net.predict(feats)
    for i in (0, 5, len(feats) - 1):
        p1, v1 = net.predict(feats[i:i + 1])
        assert np.array_equal(p1[0].view(np.uint32), pi[i].view(np.uint32)) and v1[0] == v[i]


def test_set_params_roundtrip(azb, oracle):
    net = azb.NNet(seed=1, blocks=1, precision=azb.NNET_FP32)
    w = net.get_params()
    assert len(w) == azb.param_layout(1)["total"]
    feats = random_features(oracle, 3)
    before = net.predict(feats)
    w2 = w.copy()
    o, shape = azb.param_layout(1)["val_fc2_b"]
    w2[o] += 0.5
    net.set_params(w2)
    after = net.predict(feats)
    assert np.array_equal(before[0], after[0]) and not np.array_equal(before[1], after[1])


@pytest.mark.parametrize("quirks", [0, 15])
def test_selfplay_with_network_matches_oracle(azb, oracle, quirks):
    net = azb.NNet(seed=7, blocks=2, precision=azb.NNET_FP32)
    coach = azb.Coach(nnet=net, num_sims=30, seed=4, quirks=quirks, evaluator=azb.EVAL_NNET)
    st = coach.self_play(6, 10)
    tr = coach.traces()
    boards, pis, vs = coach.export_samples()
    offs = np.concatenate([[0], np.cumsum(tr["plies"].astype(np.int64))])
    for g in range(6):
        o = oracle.execute_episode(num_sims=30, quirks=quirks, seed=4, episode_id=10 + g,
                                   evaluator=oracle.EVAL_CALLBACK, callback=net.predict)
        n = o["plies"]
        assert tr["plies"][g] == n
        assert tr["actions"][g, :n].tolist() == o["actions"][:n].tolist()
        assert np.array_equal(tr["counts"][g, :n], o["counts"][:n])
        a, b = 2 * offs[g], 2 * offs[g + 1]
        assert np.array_equal(pis[a:b].view(np.uint32), o["pis"].view(np.uint32))
        assert np.array_equal(vs[a:b], o["vs"])
    assert st["evals"] > 0 and st["launches"] > 3 * st["evals"] / 6 / 2


def torch_forward_bf16(azb, params, blocks, boards):
    """The bf16 path's numerics restated in torch: fp32 stem -> bf16; tower convs with bf16 weights and
    bf16 activations, fp32 accumulation, bias/residual/ReLU in fp32, bf16 store; heads in fp32."""
    import torch
    import torch.nn.functional as F
    L = azb.param_layout(blocks)

    def get(name):
        o, shape = L[name]
        return torch.from_numpy(params[o:o + int(np.prod(shape))].reshape(shape).copy())

    def q(t):
        return t.to(torch.bfloat16).to(torch.float64)

    x = torch.from_numpy(boards).double()
    w = get("stem_w").double().reshape(3, 3, 2, 128).permute(3, 2, 0, 1)
    x = q(F.relu(F.conv2d(x, w, get("stem_b").double(), padding=1)).float())
    tw, tb = get("tower_w"), get("tower_b").double()
    for b in range(blocks):
        w1 = q(tw[2 * b]).reshape(3, 3, 128, 128).permute(3, 2, 0, 1)
        w2 = q(tw[2 * b + 1]).reshape(3, 3, 128, 128).permute(3, 2, 0, 1)
        y = q(F.relu(F.conv2d(x, w1, tb[2 * b], padding=1)).float())
        x = q(F.relu(F.conv2d(y, w2, tb[2 * b + 1], padding=1) + x).float())
    pol = F.relu(torch.einsum("bchw,cp->bphw", x, get("pol_w").double()) + get("pol_b").double().view(1, 2, 1, 1)).reshape(len(boards), 84)
    pi = torch.softmax(pol @ get("pol_fc_w").double() + get("pol_fc_b").double(), dim=1)
    val = F.relu(torch.einsum("bchw,c->bhw", x, get("val_w").double()) + get("val_b").double()).reshape(len(boards), 42)
    h = F.relu(val @ get("val_fc1_w").double() + get("val_fc1_b").double())
    v = torch.tanh(h @ get("val_fc2_w").double() + get("val_fc2_b").double())
    return pi.numpy(), v.numpy()


@pytest.mark.parametrize("blocks", [1, 6])
def test_bf16_tensor_core_path(azb, oracle, blocks):
    net = azb.NNet(seed=7, blocks=blocks, precision=azb.NNET_BF16_TC)
    feats = random_features(oracle, 60)          # ~1200 positions: many M tiles, ragged last tile
    pi, v = net.predict(feats)
    rpi, rv = torch_forward_bf16(azb, net.get_params(), blocks, feats)
    # same quantisation points, only the fp32 accumulation order differs.  One layer pair: 5e-3.
    # Through 12 layers a borderline bf16 rounding that flips one ulp (0.4 %) is amplified by the
    # random-init tower, so the worst element gets the bf16 tolerance (5e-2) and the MEAN stays tight.
    tol = 5e-3 if blocks == 1 else 5e-2
    assert np.abs(pi - rpi).max() < tol, np.abs(pi - rpi).max()
    assert np.abs(v - rv).max() < tol, np.abs(v - rv).max()
    assert np.abs(pi - rpi).mean() < 2e-3 and np.abs(v - rv).mean() < 4e-3, (np.abs(pi - rpi).mean(), np.abs(v - rv).mean())
    # against the fp32 network: stated bf16 tolerance 5e-2 (13 layers of bf16 rounding)
    net32 = azb.NNet(seed=7, blocks=blocks, precision=azb.NNET_FP32)
    p32, v32 = net32.predict(feats)
    assert np.abs(pi - p32).max() < 5e-2 and np.abs(v - v32).max() < 5e-2
    assert np.allclose(pi.sum(1), 1.0, atol=1e-5)


def test_bf16_path_is_batch_independent(azb, oracle):
    net = azb.NNet(seed=3, blocks=2, precision=azb.NNET_BF16_TC)
    feats = random_features(oracle, 30)
    pi, v = net.predict(feats)
    for lo, hi in ((0, 1), (5, 6), (100, 300), (len(feats) - 1, len(feats))):
        p1, v1 = net.predict(feats[lo:hi])
        assert np.array_equal(p1.view(np.uint32), pi[lo:hi].view(np.uint32)) and np.array_equal(v1, v[lo:hi])


def test_selfplay_with_tensor_core_network_matches_oracle(azb, oracle):
    net = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC)
    coach = azb.Coach(nnet=net, num_sims=25, seed=2, evaluator=azb.EVAL_NNET)
    st = coach.self_play(4, 0)
    tr = coach.traces()
    for g in range(4):
        o = oracle.execute_episode(num_sims=25, quirks=0, seed=2, episode_id=g, evaluator=oracle.EVAL_CALLBACK,
                                   callback=net.predict)
        n = o["plies"]
        assert tr["plies"][g] == n
        assert tr["actions"][g, :n].tolist() == o["actions"][:n].tolist()
        assert np.array_equal(tr["counts"][g, :n], o["counts"][:n])


def test_network_games_do_not_depend_on_the_slot_count(azb):
    """The lock-step rounds run min(games, 8192, memory) slots; a game's moves, root counts and samples must not depend on how
    many games are in flight beside it (the round a leaf is evaluated in changes, the batch it shares changes, the evaluation
    cache fills in another order: the per-position result does not).  200 games on 16 slots (recycled 12 times), on 64, and
    all at once."""
    net = azb.NNet(seed=5, blocks=2, precision=azb.NNET_BF16_TC)
    ref = None
    for slots in (0, 64, 16):
        coach = azb.Coach(nnet=net, num_sims=40, seed=9, evaluator=azb.EVAL_NNET, max_concurrent_games=slots)
        st = coach.self_play(200, 1000)
        tr = coach.traces()
        b, p, v = coach.export_samples()
        got = (tr["plies"].tolist(), tr["actions"].tobytes(), tr["counts"].tobytes(), b.tobytes(), p.tobytes(), v.tobytes(), st["sims"])
        if ref is None:
            ref = got
        assert got == ref, slots


def test_arena_two_networks(azb, oracle):
    a = azb.NNet(seed=7, blocks=1, precision=azb.NNET_BF16_TC)
    b = azb.NNet(seed=8, blocks=1, precision=azb.NNET_BF16_TC)
    counts, res, st = azb.arena_play_games(8, azb.EVAL_NNET, azb.EVAL_NNET, a, b, k_open=2, num_sims=20, seed=5)
    assert sum(counts) == 8 and st["evals"] > 0
    counts2, res2, _ = azb.arena_play_games(8, azb.EVAL_NNET, azb.EVAL_NNET, a, b, k_open=2, num_sims=20, seed=5)
    assert res.tolist() == res2.tolist()


@pytest.mark.parametrize("n_pos", [9, 25, 297, 301, 1185, 1500, 2500, 3071, 3500])
def test_tower_implementations_agree_bit_for_bit(azb, oracle, tmp_path, n_pos):
    """Three implementations of the 2R-convolution tower accumulate every output in the same K order in fp32 and must agree
    bit for bit: k_tower_tc3 (one launch, position-aligned tiles, a CTA pair takes its tiles through all layers; the
    default up to ~3 k positions), k_conv3x3_tc3 launched layer by layer (AZB200_TOWER=0; also what 3500 positions use by
    default) and k_conv3x3_tc<1> (one CTA, cp.async gather, dense layout, streamed weights: AZB200_TC_PAIR=0).  The kernels
    are chosen once per process, so the other two run in child processes.  Sizes: one tile, one two-tile unit's worth, the last size with one tile
    per CTA pair and the first with two, sizes that take the 9-position two-tile units (ragged last units), the largest
    caller-sized batch that takes the tower and one above its size limit."""
    import os, subprocess, sys
    rng = np.random.default_rng(n_pos)  # disjoint random stone sets (not necessarily reachable positions)
    feats = (rng.random((n_pos, 2, 6, 7)) < 0.3).astype(np.float32)
    feats[:, 1] *= 1.0 - feats[:, 0]
    np.save(tmp_path / "feats.npy", feats)
    net = azb.NNet(seed=11, blocks=3, precision=azb.NNET_BF16_TC)
    pi, v = net.predict(feats)
    pi2, v2 = net.predict(feats)  # and the same bits twice (the tower's hand-over between layers is a race if it is wrong)
    assert np.array_equal(pi.view(np.uint32), pi2.view(np.uint32)) and np.array_equal(v.view(np.uint32), v2.view(np.uint32))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import importlib, sys, numpy as np\n"
        f"sys.path.insert(0, {root!r})\n"
        "azb = importlib.import_module('alphazero-rs_b200')\n"
        f"feats = np.load({str(tmp_path / 'feats.npy')!r})\n"
        "net = azb.NNet(seed=11, blocks=3, precision=azb.NNET_BF16_TC)\n"
        "pi, v = net.predict(feats)\n"
        f"np.save({str(tmp_path / 'pi.npy')!r}, pi); np.save({str(tmp_path / 'v.npy')!r}, v)\n"
    )
    for var in ({"AZB200_TOWER": "0"}, {"AZB200_TC_PAIR": "0"}):
        subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, **var), timeout=300)
        pi1, v1 = np.load(tmp_path / "pi.npy"), np.load(tmp_path / "v.npy")
        assert np.array_equal(pi.view(np.uint32), pi1.view(np.uint32)), var
        assert np.array_equal(v.view(np.uint32), v1.view(np.uint32)), var


@pytest.mark.parametrize("n_pos", [300, 1500, 2500])
def test_tower_hand_over_is_not_a_race(azb, n_pos):
    """k_tower_tc3 hands a tile from the epilogue warps' generic stores to the CTA pair's TMA reads of the next layer through a
    publisher warp, two flags in shared memory and a reader-side proxy fence (csrc/nnet_tc.cuh).  If that ordering were wrong,
    a pass would now and then read a stale row: 1000 passes over the same positions must give the same bits every time (one
    tile per pair, several tiles per pair with two-tile units, ragged units)."""
    rng = np.random.default_rng(100 + n_pos)
    feats = (rng.random((n_pos, 2, 6, 7)) < 0.3).astype(np.float32)
    feats[:, 1] *= 1.0 - feats[:, 0]
    net = azb.NNet(seed=3, blocks=6, precision=azb.NNET_BF16_TC)
    pi0, v0 = net.predict(feats)
    for _ in range(1000):
        pi, v = net.predict(feats)
        assert np.array_equal(pi.view(np.uint32), pi0.view(np.uint32)) and np.array_equal(v.view(np.uint32), v0.view(np.uint32))


def test_config3_parameters_sampled_games(azb, oracle):
    """BASELINE config 3 parameters (400 sims/move, ResNet-6x128 bf16 on the tcgen05 tower, seed 0xA1FA0) on a
    512-game batch — many M tiles per layer, rounds with thousands of pending leaves — and two of its games replayed
    by the oracle with the same network as its predict callback: actions and per-ply root counts bit for bit."""
    net = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC)
    coach = azb.Coach(nnet=net, num_sims=400, seed=0xA1FA0, evaluator=azb.EVAL_NNET)
    st = coach.self_play(512, 0)
    tr = coach.traces()
    assert st["games"] == 512 and st["sims"] == 400 * st["plies"]
    for g in (3, 509):
        o = oracle.execute_episode(num_sims=400, quirks=0, seed=0xA1FA0, episode_id=g, evaluator=oracle.EVAL_CALLBACK,
                                   callback=net.predict)
        n = o["plies"]
        assert tr["plies"][g] == n
        assert tr["actions"][g, :n].tolist() == o["actions"][:n].tolist()
        assert np.array_equal(tr["counts"][g, :n], o["counts"][:n])


def test_leaf_dedup_is_invisible_and_saves_evaluations(azb, tmp_path):
    """A position several trees ask for in the same round goes through the network once (rounds.cuh LeafBufs): games
    with the same move history grow identical trees, so during the first move of a call every game wants the same
    leaves; and a position the network evaluated in an earlier round of the call is answered from the call's evaluation
    cache.  The games themselves must not change: the run without either (child process, AZB200_LEAF_DEDUP=0
    AZB200_EVAL_CACHE=0) produces the same actions, root counts and samples; only the number of network rows differs."""
    import os, subprocess, sys
    script = (
        "import importlib, sys, numpy as np\n"
        f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
        "azb = importlib.import_module('alphazero-rs_b200')\n"
        "net = azb.NNet(seed=7, blocks=2)\n"
        "coach = azb.Coach(nnet=net, num_sims=40, seed=9, evaluator=azb.EVAL_NNET, temp_threshold=6)\n"
        "st = coach.self_play(96, 0)\n"
        "tr = coach.traces(); b, p, v = coach.export_samples()\n"
        "np.savez(sys.argv[1], actions=tr['actions'], counts=tr['counts'], plies=tr['plies'], boards=b, pis=p, vs=v,\n"
        "         evals=st['evals'], nn=st['nn_positions'], hits=st['nn_cache_hits'])\n"
        "a = azb.NNet(seed=7, blocks=1); bnet = azb.NNet(seed=8, blocks=1)\n"
        "c0, r0, s0 = azb.arena_play_games(16, azb.EVAL_NNET, azb.EVAL_NNET, a, bnet, k_open=0, num_sims=20, seed=5)\n"
        "c2, r2, s2 = azb.arena_play_games(16, azb.EVAL_NNET, azb.EVAL_NNET, a, bnet, k_open=3, num_sims=20, seed=5)\n"
        "np.savez(sys.argv[2], r0=r0, r2=r2, e0=s0['evals'], n0=s0['nn_positions'], e2=s2['evals'], n2=s2['nn_positions'])\n")
    out = {}
    for tag, val in (("on", "1"), ("off", "0")):
        env = dict(os.environ, AZB200_LEAF_DEDUP=val, AZB200_EVAL_CACHE=val)
        r = subprocess.run([sys.executable, "-c", script, str(tmp_path / f"sp_{tag}.npz"), str(tmp_path / f"ar_{tag}.npz")],
                           env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        out[tag] = (np.load(tmp_path / f"sp_{tag}.npz"), np.load(tmp_path / f"ar_{tag}.npz"))
    (sp_on, ar_on), (sp_off, ar_off) = out["on"], out["off"]
    for k in ("actions", "counts", "plies", "boards", "pis", "vs"):
        assert np.array_equal(sp_on[k], sp_off[k]), k
    assert int(sp_on["evals"]) == int(sp_off["evals"]) == int(sp_off["nn"])  # without it every evaluation is a network row
    assert int(sp_on["nn"]) < int(sp_on["evals"])                            # the shared first moves are evaluated once
    assert int(sp_on["hits"]) > 0 and int(sp_off["hits"]) == 0               # ... and later games find them in the cache
    assert int(sp_on["nn"]) + int(sp_on["hits"]) <= int(sp_on["evals"])
    assert np.array_equal(ar_on["r0"], ar_off["r0"]) and np.array_equal(ar_on["r2"], ar_off["r2"])
    # an arena without opening plies plays the same game 8 times per seat order: 2 distinct games' worth of rows
    assert int(ar_on["n0"]) * 6 <= int(ar_on["e0"]) and int(ar_off["n0"]) == int(ar_off["e0"])
    assert int(ar_on["n2"]) <= int(ar_on["e2"])
