"""Coach::learn (coach.rs:169-396) on the device: self-play -> queue trim -> history window -> `<n>.examples` ->
shuffle -> train model_id -> model_id+1 -> arena -> accept rule, with double-buffered networks; plus the weight
checkpoints and the resume in Coach::setup (coach.rs:55-81).  The device phases themselves (self-play rounds, training
step, arena) have their own parity tests; here the loop's bookkeeping is checked against the oracle's restatement
(oracle/learn.hpp) and against the same phases run one by one through the public calls."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG = dict(num_iters=3, num_eps=24, num_sims=16, num_arena_games=8, max_queue_length=700, max_history_length=2,
           update_threshold=0.5, seed=11, temp_threshold=8)
BLOCKS, NET_SEED, EPOCHS, BATCH = 2, 21, 3, 128


def run_learn(azb, ckpt, **over):
    cfg = dict(CFG, **over)
    coach = azb.Coach(checkpoint_directory=str(ckpt).encode(), evaluator=azb.EVAL_NNET, **cfg)
    reports, net = coach.learn(epochs=EPOCHS, batch_size=BATCH, blocks=BLOCKS, seed=NET_SEED, arena_k_open=2)
    return coach, reports, net


def test_learn_loop_bookkeeping(azb, oracle, tmp_path):
    ckpt = tmp_path / "ckpt"
    coach, reports, net = run_learn(azb, ckpt)
    assert [r["iteration"] for r in reports] == [0, 1, 2]
    # coach.rs:274-289 through the oracle's restatement
    sizes, dropped = oracle.learn_window([r["samples_played"] for r in reports], CFG["max_queue_length"],
                                         CFG["max_history_length"])
    model_id = 0
    for i, r in enumerate(reports):
        assert r["games"] == CFG["num_eps"] and r["samples_played"] > CFG["max_queue_length"]  # the trim is exercised
        assert r["samples_kept"] == r["samples_played"] - dropped[i]
        assert r["history_iterations"] == int((sizes[i] > 0).sum())
        assert r["history_samples"] == int(sizes[i].sum())
        assert r["train_steps"] == EPOCHS
        assert r["nwins"] + r["pwins"] + r["draws"] == CFG["num_arena_games"]
        assert bool(r["accepted"]) == oracle.learn_accept(r["nwins"], r["pwins"], CFG["update_threshold"])  # coach.rs:383-390
        assert r["model_id_before"] == model_id
        model_id += r["accepted"]
        assert r["model_id_after"] == model_id
        assert all(np.isfinite(r["loss_first"])) and all(np.isfinite(r["loss_last"]))
        # `<iteration>.examples` holds the window of that iteration, byte for byte what the oracle encodes
        c, b, p, v = azb.examples_read(ckpt / f"{i}.examples")
        assert c.tolist() == [int(x) for x in sizes[i] if x > 0]
        assert (ckpt / f"{i}.examples").read_bytes() == oracle.examples_encode(c, b, p, v)
        assert os.path.exists(ckpt / f"{r['model_id_before'] + 1}.azbw")
    # the history in the coach is the last file
    hc, hb, hp, hv = coach.history()
    c, b, p, v = azb.examples_read(ckpt / "2.examples")
    assert hc.tolist() == c.tolist() and (hb == b).all() and (hp == p).all() and (hv == v).all()
    # model 0 on disk = the random init of net_cfg; the returned network = the accepted model's file
    ref0 = azb.NNet(seed=NET_SEED, blocks=BLOCKS)
    m0 = azb.NNet(seed=1, blocks=BLOCKS)
    m0.load(ckpt / "0.azbw")
    assert (m0.get_params() == ref0.get_params()).all()
    mf = azb.NNet(seed=1, blocks=BLOCKS)
    mf.load(ckpt / f"{model_id}.azbw")
    if model_id > 0 and reports[-1]["accepted"]:
        assert (mf.get_params() == net.get_params()).all()
    # Coach::setup resumes from the newest file (coach.rs:55-77)
    again = azb.Coach(checkpoint_directory=str(ckpt).encode(), evaluator=azb.EVAL_NNET, **CFG)
    rc, rb, rp, rv = again.history()
    assert rc.tolist() == hc.tolist() and (rb == hb).all() and (rp == hp).all() and (rv == hv).all()


def test_learn_first_iteration_equals_the_phases_run_by_hand(azb, oracle, tmp_path):
    """Iteration 0 replayed through the public calls: self-play with model 0, keep the newest max_queue_length samples,
    shuffle with the documented permutation, the first Adam step's losses, and (when the candidate is accepted or not)
    the arena result of candidate vs model 0."""
    ckpt = tmp_path / "ckpt"
    coach, reports, net = run_learn(azb, ckpt, num_iters=1)
    r = reports[0]
    net0 = azb.NNet(seed=NET_SEED, blocks=BLOCKS)
    c2 = azb.Coach(nnet=net0, evaluator=azb.EVAL_NNET, checkpoint_directory=str(tmp_path / "none").encode(), **CFG)
    c2.self_play(CFG["num_eps"], 0)
    boards, pis, vs = c2.export_samples()
    assert len(vs) == r["samples_played"]
    keep = slice(len(vs) - CFG["max_queue_length"], None)
    hc, hb, hp, hv = coach.history()
    assert hc.tolist() == [CFG["max_queue_length"]]
    assert (hb == boards[keep]).all() and (hp == pis[keep]).all() and (hv == vs[keep]).all()
    perm = oracle.learn_shuffle_perm(CFG["seed"], 0, len(hv)).astype(np.int64)
    cand = azb.NNet(seed=NET_SEED, blocks=BLOCKS)
    losses = []
    for s in range(EPOCHS):
        idx = perm[(np.arange(BATCH) + s * BATCH) % len(perm)]
        losses.append(cand.train((hb[idx], hp[idx], hv[idx]), lr=1e-4))  # azb_learn_config_default's step
    assert np.allclose(losses[0], r["loss_first"], rtol=1e-4, atol=1e-5)
    assert np.allclose(losses[-1], r["loss_last"], rtol=2e-2, atol=1e-3)  # fp32 atomics in the weight-gradient reduction
    counts, _, _ = azb.arena_play_games(CFG["num_arena_games"], azb.EVAL_NNET, azb.EVAL_NNET, cand, net0, k_open=2,
                                        **{k: CFG[k] for k in ("num_sims", "seed", "temp_threshold")})
    assert sum(counts) == CFG["num_arena_games"]


def test_weight_checkpoint_round_trip_with_adam_state(azb, tmp_path):
    rng = np.random.default_rng(3)
    boards = (rng.random((64, 2, 6, 7)) < 0.3).astype(np.float32)
    pis = rng.random((64, 7)).astype(np.float32)
    pis /= pis.sum(1, keepdims=True)
    vs = rng.choice(np.array([-1.0, 1.0], np.float32), 64)
    a = azb.NNet(seed=5, blocks=2)
    a.train((boards, pis, vs))
    a.train((boards, pis, vs))
    a.save(tmp_path / "7.azbw")
    b = azb.NNet(seed=6, blocks=2)
    b.load(tmp_path / "7.azbw")
    assert (a.get_params() == b.get_params()).all()
    c = azb.NNet(seed=9, blocks=2)
    c.copy_from(a)
    assert (a.get_params() == c.get_params()).all()
    pa = a.predict(boards)
    assert all((x == y).all() for x, y in zip(pa, b.predict(boards))) and all((x == y).all() for x, y in zip(pa, c.predict(boards)))
    # the Adam moments and step count travel: a third step moves all three the same way
    la, lb, lc = a.train((boards, pis, vs)), b.train((boards, pis, vs)), c.train((boards, pis, vs))
    assert np.allclose(la, lb, rtol=1e-5) and np.allclose(la, lc, rtol=1e-5)
    da = a.get_params()
    for other in (b, c):
        d = other.get_params()
        assert np.abs(da - d).max() <= 2e-3 * 1.0 and np.corrcoef(da, d)[0, 1] > 0.999999
    with pytest.raises(azb.AzbError):
        azb.NNet(seed=1, blocks=3).load(tmp_path / "7.azbw")  # architecture mismatch is an error, not a reshape
    (tmp_path / "bad.azbw").write_bytes(b"AZBX" + b"\0" * 64)
    with pytest.raises(azb.AzbError):
        b.load(tmp_path / "bad.azbw")


def test_learn_skip_first_play_and_errors(azb, tmp_path):
    # skip_first_play with an empty history: coach.rs:304 assert!(num_samples > 0)
    coach = azb.Coach(checkpoint_directory=str(tmp_path / "a").encode(), evaluator=azb.EVAL_NNET, **CFG)
    with pytest.raises(azb.AzbError):
        coach.learn(skip_first_play=True, epochs=1, batch_size=32, blocks=1)
    # with a resumed history the first iteration trains on it without playing
    ck = tmp_path / "b"
    os.makedirs(ck)
    rng = np.random.default_rng(1)
    boards = (rng.random((300, 2, 6, 7)) < 0.3).astype(np.float32)
    pis = np.full((300, 7), 1 / 7, np.float32)
    vs = rng.choice(np.array([-1.0, 1.0], np.float32), 300)
    azb.examples_write(ck / "4.examples", [100, 200], boards, pis, vs)
    coach = azb.Coach(checkpoint_directory=str(ck).encode(), evaluator=azb.EVAL_NNET, **dict(CFG, num_iters=1, max_history_length=5))
    reports, net = coach.learn(skip_first_play=True, epochs=2, batch_size=32, blocks=1)
    r = reports[0]
    assert r["games"] == 0 and r["samples_kept"] == 0
    assert r["history_iterations"] == 3 and r["history_samples"] == 300  # the empty entry is pushed (coach.rs:284)
    # learn needs the network evaluator
    uni = azb.Coach(checkpoint_directory=str(tmp_path / "c").encode(), **CFG)
    with pytest.raises(azb.AzbError):
        uni.learn(epochs=1, batch_size=32, blocks=1)
