"""azb_coach_learn_dist: Coach::learn data parallel over ranks (SURVEY 8e / BASELINE config 5).  Two ranks share the one
GPU of the test box and reduce over gloo (torch's gloo takes CUDA tensors); the same script runs one rank per GPU over
NCCL (scripts/learn_dist.py under torchrun).  Checked: the ranks' self-play shares are exactly the games a single process
plays (bit for bit), every rank reports the same arena counters / decision, the replicas stay bit-identical."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_learn_like_one(azb, tmp_path):
    env = dict(os.environ, AZB_DIST_BACKEND="gloo", AZB_DIST_ONE_GPU="1", AZB_DIST_OUT=str(tmp_path), AZB_DIST_QUEUE="100000",
               AZB_DIST_TEMP_THRESHOLD="8")
    iters, eps, sims, batch, arena, epochs, lr, blocks = 2, 24, 16, 128, 8, 3, 1e-4, 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "scripts", "learn_dist.py")] + \
          [str(x) for x in (iters, eps, sims, batch, arena, epochs, lr, blocks)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    assert summary["replicas_identical"] and summary["n_gpus"] == 2
    rep = [json.load(open(tmp_path / f"rank{k}.json")) for k in range(2)]
    dat = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    assert rep[0]["crc"] == rep[1]["crc"]
    for a, b in zip(rep[0]["reports"], rep[1]["reports"]):
        assert a["games"] == b["games"] == eps // 2
        for k in ("nwins", "pwins", "draws", "accepted", "model_id_after", "train_steps"):
            assert a[k] == b[k], k
        assert a["nwins"] + a["pwins"] + a["draws"] == arena
        assert a["train_steps"] == epochs
    # iteration 0: rank r played games [12 r, 12 r + 12) of the 24 a single process plays with the same model 0
    net0 = azb.NNet(seed=7, blocks=blocks)
    c = azb.Coach(nnet=net0, evaluator=azb.EVAL_NNET, checkpoint_directory=str(tmp_path / "none").encode(), num_sims=sims,
                  seed=0xA1FA0, temp_threshold=8)
    c.self_play(eps, 0)
    boards, pis, vs = c.export_samples()
    plies = c.traces()["plies"].astype(np.int64)
    cut = 2 * int(plies[: eps // 2].sum())
    assert len(dat[0]["vs"]) == cut and len(dat[1]["vs"]) == len(vs) - cut
    assert (dat[0]["boards"] == boards[:cut]).all() and (dat[0]["pis"] == pis[:cut]).all() and (dat[0]["vs"] == vs[:cut]).all()
    assert (dat[1]["boards"] == boards[cut:]).all() and (dat[1]["pis"] == pis[cut:]).all() and (dat[1]["vs"] == vs[cut:]).all()
    assert (dat[0]["params"] == dat[1]["params"]).all()
