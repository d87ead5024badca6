"""The library's own communicator and multi-GPU fan-out (csrc/dist.cuh): NCCL loaded with dlopen, one process
per GPU, no torch.  One GPU is what the test box has: a world of one rank exercises init / all-reduce / the azb_dist
callbacks inside azb_coach_learn_dist, and azb_coach_self_play_multi runs its per-device worker threads on device 0 (twice
the same device = two workers sharing one GPU).  The N = 2..8 path is run by bench.py --gpus N."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_comm_world_of_one(azb, tmp_path):
    comm = azb.Comm(0, 1, 0, str(tmp_path / "nccl.id"))
    assert comm.reduce(3.5, "SUM") == 3.5 and comm.reduce(-2.0, "MAX") == -2.0 and comm.reduce(7.0, "MIN") == 7.0
    comm.barrier()
    d = comm.dist()
    assert d.rank == 0 and d.world == 1
    # the u64 host callback sums through a device buffer
    a = (np.arange(3, dtype=np.uint64) + 5)
    import ctypes as C
    assert d.allreduce_sum_u64_host(a.ctypes.data_as(C.POINTER(C.c_uint64)), 3, d.user) == 0
    assert a.tolist() == [5, 6, 7]
    # the f32 device callback: ncclAllReduce in place on the gradient vector where it lives (a sum over one rank = identity)
    net = azb.NNet(seed=3, blocks=1)
    rng = np.random.default_rng(0)
    boards = (rng.random((8, 2, 6, 7)) < 0.2).astype(np.float32)
    pis = np.full((8, 7), 1 / 7, np.float32)
    net.train_begin(boards, pis, rng.uniform(-1, 1, 8).astype(np.float32))
    g0 = net.grads()
    ptr, n = C.c_void_p(), C.c_uint64()
    assert azb.lib.azb_nnet_grads_device(net._h, C.byref(ptr), C.byref(n)) == 0 and n.value == len(g0)
    assert d.allreduce_sum_f32_device(ptr, n.value, d.user) == 0
    assert np.array_equal(net.grads(), g0) and np.abs(g0).sum() > 0
    comm.close()


def test_learn_through_the_library_communicator(azb, tmp_path):
    """azb_coach_learn_dist with the callbacks of azb_dist_make (ncclAllReduce on the library's communicator) must give what
    azb_coach_learn gives on one rank: same reports, bit-identical parameters."""
    def run(dist, sub):
        coach = azb.Coach(checkpoint_directory=str(tmp_path / sub).encode(), num_eps=48, num_sims=20, num_arena_games=8,
                          num_iters=1, seed=5, evaluator=azb.EVAL_NNET, max_queue_length=10 ** 9)
        reports, net = coach.learn(epochs=3, batch_size=64, lr=1e-4, seed=7, blocks=1, dist=dist, save_files=False)
        return reports, net.get_params()
    comm = azb.Comm(0, 1, 0, str(tmp_path / "nccl2.id"))
    # world == 1: learn_dist takes the single-rank path, but the communicator's callbacks are wired and callable
    r1, p1 = run(comm, "a")
    r0, p0 = run(None, "b")
    assert np.array_equal(p0, p1)
    for k in ("games", "samples_played", "train_steps", "nwins", "pwins", "draws", "accepted"):
        assert r0[0][k] == r1[0][k], k
    comm.close()


def test_self_play_multi_matches_single_calls(azb, oracle):
    """One call, one worker thread per listed device; shard d plays games [first + d * n, ...).  Two workers on device 0:
    the union of their games equals two separate self-play calls, and sampled games replay on the oracle."""
    stats, wall = azb.self_play_multi([0, 0], 24, first_game_id=100, num_sims=30, seed=9, evaluator=azb.EVAL_HASH)
    assert len(stats) == 2 and all(s["games"] == 24 for s in stats) and wall > 0
    for d in range(2):
        coach = azb.Coach(num_sims=30, seed=9, evaluator=azb.EVAL_HASH)
        st = coach.self_play(24, 100 + 24 * d)
        for k in ("plies", "sims", "levels", "expansions", "terminal_hits", "dup_links", "evals"):
            assert st[k] == stats[d][k], (d, k)
    # with the network evaluator every worker creates its own replica from the same seed
    nc = azb.NnetConfig(0, 1, azb.NNET_BF16_TC, 0, 7)
    stats, _ = azb.self_play_multi([0], 6, first_game_id=0, net_cfg=nc, num_sims=20, seed=2, evaluator=azb.EVAL_NNET)
    net = azb.NNet(seed=7, blocks=1)
    coach = azb.Coach(nnet=net, num_sims=20, seed=2, evaluator=azb.EVAL_NNET)
    st = coach.self_play(6, 0)
    assert st["plies"] == stats[0]["plies"] and st["levels"] == stats[0]["levels"] and st["evals"] == stats[0]["evals"]
