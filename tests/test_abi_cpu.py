"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/azb200.h declares, and refuses (loudly) to compute without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "azb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(azb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(azb):
    names = declared_symbols()
    assert len(names) >= 20
    lib = C.CDLL(azb.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/azb200.h but not exported"


def test_binding_covers_header(azb):
    assert set(declared_symbols()) == set(azb.ABI_SYMBOLS)


def test_no_torch_in_library(azb):
    out = subprocess.run(["ldd", azb.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out


def test_config_default_mirrors_example(azb):
    cfg = azb.default_config()
    # examples/connect_four.rs:55-71
    assert (cfg.mcts_reserve_size, cfg.temp_threshold, cfg.max_history_length, cfg.max_queue_length) == (1000000, 15, 20, 200000)
    assert (cfg.inference_batch_size, cfg.num_episode_threads, cfg.num_arena_games, cfg.num_iters, cfg.num_eps) == (1, 1, 40, 1, 1)
    assert (cfg.num_sims, cfg.num_sim_threads, cfg.max_depth, cfg.cpuct) == (25, 1, 1000, 1)
    assert abs(cfg.update_threshold - 0.6) < 1e-7


def test_host_only_calls(azb):
    assert azb.ConnectFourGame.get_feature_shape() == [2, 6, 7]
    b = azb.ConnectFourGame.get_init_board(3)
    assert (b["s"] == 0).all() and (b["me"] == 1).all()


def test_invalid_config_rejected(azb):
    with pytest.raises(azb.AzbError) as e:
        azb.AsyncMcts(1, num_sims=32, num_sim_threads=16)  # waves hold at most 8 walks
    assert e.value.code == azb.ERR_UNSUPPORTED
    with pytest.raises(azb.AzbError) as e:
        azb.AsyncMcts(1, num_sims=25, num_sim_threads=4)   # async_mcts.rs:192: num_sims % num_threads == 0
    assert e.value.code == azb.ERR_INVALID
    with pytest.raises(azb.AzbError) as e:
        azb.Coach(num_sims=0)
    assert e.value.code == azb.ERR_INVALID


def test_fails_loudly_without_gpu(azb):
    if azb.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(azb.AzbError) as e:
        azb.ConnectFourGame.get_valid_moves(azb.ConnectFourGame.get_init_board(1))
    assert e.value.code == azb.ERR_CUDA
    with pytest.raises(azb.AzbError) as e:
        azb.Coach(num_sims=25)
    assert e.value.code == azb.ERR_CUDA
    with pytest.raises(azb.AzbError):
        azb.AsyncMcts(1, num_sims=25)


def test_product_does_not_reference_oracle():
    """The product path must not import, link or open anything under oracle/."""
    pkg = os.path.join(ROOT, "alphazero-rs_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "azoracle" not in txt and "oracle_api" not in txt and "oracle/" not in txt.replace("the oracle", ""), f
