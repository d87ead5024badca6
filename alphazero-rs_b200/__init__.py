"""azb200 — host-side mirror (Python, ctypes) of the reference's self-play API over libazb200.so.

The product is the C-ABI library (include/azb200.h).  This module is the thin binding the
tests and bench.py use; its names follow the reference crate:

  ConnectFourGame  — trait Game           (src/game.rs:10-28, connect_four_game.rs)
  AsyncMcts        — src/async_mcts.rs    (private in the reference; test hook here)
  Coach            — src/coach.rs         (setup / execute_episode / self-play fan-out)

There is no CPU fallback: if libazb200.so is missing the import fails, and every compute
call fails with AZB_ERR_CUDA when no device is visible.
"""
import ctypes as C
import importlib.util
import os

import numpy as np

from . import sharding  # noqa: F401  (multi-GPU game sharding helpers)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AZB200_LIB", os.path.join(_HERE, "libazb200.so"))  # override: kernel-variant sweeps

Q1_WIN_RANGE_LITERAL = 1
Q2_BACKUP_NO_ALTERNATE = 2
Q3_POS_BACKUP_PLUS_ONE = 4
Q4_VLABEL_LITERAL = 8
PROFILE_REFERENCE = 15
PROFILE_SANE = 0
EVAL_UNIFORM = 0
EVAL_HASH = 1
EVAL_NNET = 2

AZB_OK = 0
ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED = -1, -2, -3, -4

STATE_DTYPE = np.dtype([("s", np.int8, (6, 7)), ("me", np.int8)])  # 43 bytes, packed
assert STATE_DTYPE.itemsize == 43


class AzbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"azb200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    """azb_config: the 15 positional parameters of Coach::setup (coach.rs:38-54) + engine fields."""

    _fields_ = [
        ("checkpoint_directory", C.c_char_p),
        ("mcts_reserve_size", C.c_uint64),
        ("update_threshold", C.c_float),
        ("temp_threshold", C.c_uint64),
        ("max_history_length", C.c_uint64),
        ("max_queue_length", C.c_uint64),
        ("inference_batch_size", C.c_uint64),
        ("num_episode_threads", C.c_uint64),
        ("num_arena_games", C.c_uint64),
        ("num_iters", C.c_uint64),
        ("num_eps", C.c_uint64),
        ("num_sims", C.c_uint64),
        ("num_sim_threads", C.c_uint64),
        ("max_depth", C.c_uint64),
        ("cpuct", C.c_int32),
        ("quirks", C.c_uint32),
        ("seed", C.c_uint64),
        ("evaluator", C.c_int32),
        ("device", C.c_int32),
        ("max_concurrent_games", C.c_uint64),
        ("schedule", C.c_uint32),
        ("plies_per_launch", C.c_uint32),
    ]


class SelfPlayStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "games", "plies", "samples", "sims", "levels", "expansions", "terminal_hits",
        "dup_links", "evals", "blocks_used_max", "owners_max")] + [("device_ms", C.c_double), ("launches", C.c_uint64),
                                                                ("trees_resident", C.c_uint64), ("nn_positions", C.c_uint64),
                                                                ("nn_cache_hits", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ArenaOpts(C.Structure):
    """azb_arena_opts"""
    _fields_ = [("k_open", C.c_uint32), ("shared_trees", C.c_uint32), ("first_game_id", C.c_uint64)]


class NnetConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("blocks", C.c_int32), ("precision", C.c_int32),
                ("reserved", C.c_int32), ("seed", C.c_uint64)]


class TrainConfig(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float)]


class LearnConfig(C.Structure):
    """azb_learn_config: the training schedule of an iteration (connect_four_net.py:13-21) + Coach::learn's flags."""
    _fields_ = [("epochs", C.c_uint32), ("batch_size", C.c_uint32), ("adam", TrainConfig), ("arena_k_open", C.c_uint32),
                ("skip_first_play", C.c_uint32), ("save_files", C.c_uint32), ("arena_shared_trees", C.c_uint32)]


class LearnReport(C.Structure):
    _fields_ = ([(n, C.c_uint64) for n in ("iteration", "model_id_before", "model_id_after", "games", "samples_played",
                                           "samples_kept", "history_iterations", "history_samples", "train_steps")] +
                [("loss_first", C.c_float * 2), ("loss_last", C.c_float * 2)] +
                [(n, C.c_uint64) for n in ("nwins", "pwins", "draws")] +
                [("accepted", C.c_int32), ("reserved", C.c_int32),
                 ("selfplay_ms", C.c_double), ("train_ms", C.c_double), ("arena_ms", C.c_double)])

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}
        d["loss_first"] = [float(x) for x in self.loss_first]
        d["loss_last"] = [float(x) for x in self.loss_last]
        return d


ALLREDUCE_F32_DEVICE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_void_p)
ALLREDUCE_U64_HOST = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint64), C.c_uint64, C.c_void_p)


class Dist(C.Structure):
    """azb_dist: rank, world and the caller's two all-reduce callbacks (the library owns no communicator)."""
    _fields_ = [("rank", C.c_uint32), ("world", C.c_uint32), ("allreduce_sum_f32_device", ALLREDUCE_F32_DEVICE),
                ("allreduce_sum_u64_host", ALLREDUCE_U64_HOST), ("user", C.c_void_p)]


def make_dist(dist, device):
    """azb_dist over an initialised torch.distributed: NCCL reduces the device buffer in place over NVLink; gloo (CPU
    tests, or two ranks sharing one GPU) takes CUDA tensors too and stages them through the host."""
    import torch

    def ar_f32(ptr, count, _user):
        try:
            class _Dev:
                __cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
            t = torch.as_tensor(_Dev(), device=f"cuda:{device}")
            dist.all_reduce(t)
            torch.cuda.synchronize(device)
            return 0
        except Exception as e:  # never let an exception cross the C frame
            print("azb200 allreduce_sum_f32_device:", e)
            return 1

    def ar_u64(ptr, count, _user):
        try:
            a = np.ctypeslib.as_array(ptr, shape=(int(count),))
            t = torch.from_numpy(a.astype(np.int64))
            if dist.get_backend() == "nccl":
                t = t.to(f"cuda:{device}")
            dist.all_reduce(t)
            a[:] = t.cpu().numpy().astype(np.uint64)
            return 0
        except Exception as e:
            print("azb200 allreduce_sum_u64_host:", e)
            return 1

    d = Dist(dist.get_rank(), dist.get_world_size(), ALLREDUCE_F32_DEVICE(ar_f32), ALLREDUCE_U64_HOST(ar_u64), None)
    return d


NNET_BF16_TC, NNET_FP32 = 0, 1


class Comm:
    """The library's own NCCL communicator (azb_dist_*): one process per GPU, no torch.  The 128-byte unique id travels
    from rank 0 to the others through a file next to the rendezvous port (every rank of a box sees /tmp)."""
    OPS = {"SUM": 0, "MAX": 1, "MIN": 2}

    @staticmethod
    def _pick_nccl():
        """The library loads NCCL with dlopen("libnccl.so.2") unless AZB200_NCCL_LIB names a file.  In a Python process
        that may import torch LATER, the system library must not be the one loaded first: the loader would hand that copy
        (same soname, older version) to torch, whose libtorch_cuda then misses symbols.  So when the torch wheel's own
        NCCL (site-packages/nvidia/nccl) is installed, point the library at it — found by path, torch is not imported."""
        if os.environ.get("AZB200_NCCL_LIB"):
            return
        spec = importlib.util.find_spec("nvidia")
        for root in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(root, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["AZB200_NCCL_LIB"] = cand
                return

    def __init__(self, rank, world, device, id_path, timeout_s=120.0):
        import time
        self._pick_nccl()
        self.rank, self.world, self.device = rank, world, device
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            _check(lib.azb_dist_unique_id(buf))
            tmp = f"{id_path}.tmp{os.getpid()}"
            with open(tmp, "wb") as f:
                f.write(bytes(buf))
            os.replace(tmp, id_path)  # atomic: the other ranks never see a partial id
            raw = bytes(buf)
        else:
            t0 = time.time()
            while not (os.path.exists(id_path) and os.path.getsize(id_path) == 128):
                if time.time() - t0 > timeout_s:
                    raise TimeoutError(f"no NCCL unique id at {id_path}")
                time.sleep(0.01)
            raw = open(id_path, "rb").read()
        self._h = C.c_void_p()
        _check(lib.azb_dist_init((C.c_uint8 * 128).from_buffer_copy(raw), rank, world, device, C.byref(self._h)))

    def dist(self):
        """azb_dist for azb_coach_learn_dist (callbacks backed by this communicator)."""
        d = Dist()
        _check(lib.azb_dist_make(self._h, C.byref(d)))
        return d

    def reduce(self, value, op):
        a = np.array([float(value)], np.float64)
        _check(lib.azb_dist_allreduce_f64(self._h, _ptr(a), 1, self.OPS[op]))
        return float(a[0])

    def barrier(self):
        self.reduce(0.0, "SUM")

    def close(self):
        if self._h and lib is not None:
            lib.azb_dist_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close


def self_play_multi(devices, games_per_device, first_game_id=0, net_cfg=None, **cfg):
    """azb_coach_self_play_multi: one call, one host thread per device.  Returns (per-device stats dicts, wall ms)."""
    config = cfg.pop("config", None) or default_config(**cfg)
    dev = np.ascontiguousarray(devices, np.int32)
    stats = (SelfPlayStats * len(dev))()
    wall = C.c_double()
    _check(lib.azb_coach_self_play_multi(C.byref(config), C.byref(net_cfg) if net_cfg is not None else None, _ptr(dev), len(dev),
                                         games_per_device, first_game_id, stats, C.byref(wall)))
    return [s.as_dict() for s in stats], wall.value


def build_module():
    spec = importlib.util.spec_from_file_location("azb200_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python alphazero-rs_b200/build.py` "
            "(the engine has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.azb_last_error.restype = C.c_char_p
    lib.azb_config_default.argtypes = [C.POINTER(Config)]
    lib.azb_config_default.restype = None
    vp, sz, u64, u32, f32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_float
    sigs = {
        "azb_device_count": [],
        "azb_release_caches": [],
        "azb_c4_init": [vp, sz],
        "azb_c4_feature_shape": [vp],
        "azb_c4_next_state": [vp, vp, vp, sz, vp, vp],
        "azb_c4_valid_moves": [vp, sz, vp],
        "azb_c4_game_ended": [vp, vp, sz, u32, vp],
        "azb_c4_canonical_form": [vp, vp, sz, vp],
        "azb_c4_symmetries": [vp, vp, sz, vp, vp],
        "azb_c4_eval_heuristic": [vp, sz, vp],
        "azb_c4_to_features": [vp, sz, vp],
        "azb_coach_setup": [C.POINTER(Config), C.POINTER(vp)],
        "azb_coach_destroy": [vp],
        "azb_coach_self_play": [vp, u64, u64, C.POINTER(SelfPlayStats)],
        "azb_coach_self_play_begin": [vp, u64, u64],
        "azb_coach_self_play_end": [vp, C.POINTER(SelfPlayStats)],
        "azb_coach_span_mark": [vp],
        "azb_coach_span_ms": [vp, vp, C.POINTER(C.c_double)],
        "azb_coach_traces": [vp, vp, vp, vp, vp, vp],
        "azb_coach_num_samples": [vp, C.POINTER(u64)],
        "azb_coach_ply_times": [vp, vp],
        "azb_coach_export_samples": [vp, vp, vp, vp, u64, C.POINTER(u64)],
        "azb_mcts_create": [C.POINTER(Config), u64, C.POINTER(vp)],
        "azb_mcts_destroy": [vp],
        "azb_mcts_get_action_prob": [vp, vp, f32, vp, vp],
        "azb_mcts_counter_of": [vp, vp, vp],
        "azb_mcts_stats": [vp, vp],
        "azb_mcts_dump": [vp, u64, u64, vp, vp, vp, vp, vp, C.POINTER(u64)],
        "azb_selftest_arith": [vp],
        "azb_host_alloc": [sz, C.POINTER(vp)],
        "azb_host_free": [vp],
        "azb_nnet_create": [C.POINTER(NnetConfig), C.POINTER(vp)],
        "azb_nnet_destroy": [vp],
        "azb_nnet_predict": [vp, vp, sz, sz, vp, vp],
        "azb_nnet_num_params": [vp, C.POINTER(u64)],
        "azb_nnet_get_params": [vp, vp, u64],
        "azb_nnet_set_params": [vp, vp, u64],
        "azb_coach_set_nnet": [vp, vp],
        "azb_nnet_benchmark": [vp, u64, u32, C.POINTER(C.c_double)],
        "azb_nnet_train_begin": [vp, vp, vp, vp, u64, vp],
        "azb_nnet_grads": [vp, vp, u64],
        "azb_nnet_set_grads": [vp, vp, u64],
        "azb_nnet_grads_device": [vp, C.POINTER(C.c_void_p), C.POINTER(u64)],
        "azb_nnet_train_apply": [vp, vp],
        "azb_nnet_train": [vp, vp, vp, vp, u64, vp, vp],
        "azb_nnet_conv_hook": [vp, C.c_int32, C.c_int32, vp, vp, vp, u64, vp],
        "azb_nnet_wgrad_hook": [vp, vp, vp, u64, vp],
        "azb_arena_play_games": [C.POINTER(Config), u64, C.c_int32, C.c_int32, vp, vp, u32, vp, vp,
                                 C.POINTER(SelfPlayStats)],
        "azb_arena_play_games_ex": [C.POINTER(Config), u64, C.c_int32, C.c_int32, vp, vp, C.POINTER(ArenaOpts), vp, vp, vp, vp, vp,
                                    C.POINTER(SelfPlayStats)],
        "azb_examples_write": [C.c_char_p, u64, vp, vp, vp, vp],
        "azb_examples_stat": [C.c_char_p, C.POINTER(u64), vp, u64, C.POINTER(u64)],
        "azb_examples_read": [C.c_char_p, vp, vp, vp, u64],
        "azb_examples_latest": [C.c_char_p, C.POINTER(u64)],
        "azb_nnet_save": [vp, C.c_char_p],
        "azb_nnet_load": [vp, C.c_char_p],
        "azb_nnet_copy": [vp, vp],
        "azb_coach_learn": [vp, C.POINTER(NnetConfig), C.POINTER(LearnConfig), C.POINTER(LearnReport), u64, C.POINTER(u64),
                            C.POINTER(vp)],
        "azb_coach_learn_dist": [vp, C.POINTER(NnetConfig), C.POINTER(LearnConfig), C.POINTER(Dist), C.POINTER(LearnReport), u64,
                                 C.POINTER(u64), C.POINTER(vp)],
        "azb_dist_unique_id": [vp],
        "azb_dist_init": [vp, u32, u32, C.c_int32, C.POINTER(vp)],
        "azb_dist_destroy": [vp],
        "azb_dist_make": [vp, C.POINTER(Dist)],
        "azb_dist_allreduce_f64": [vp, vp, u64, C.c_int32],
        "azb_coach_self_play_multi": [C.POINTER(Config), C.POINTER(NnetConfig), vp, u32, u64, u64, vp, C.POINTER(C.c_double)],
        "azb_coach_history_stat": [vp, C.POINTER(u64), vp, u64, C.POINTER(u64)],
        "azb_coach_history_export": [vp, vp, vp, vp, u64],
        "azb_coach_save_train_examples": [vp, u64, C.c_char_p],
        "azb_coach_load_train_examples": [vp, C.c_char_p],
        "azb_learn_accept": [u64, u64, f32],
        "azb_learn_shuffle_perm": [u64, u64, u64, vp],
    }
    lib.azb_learn_config_default.argtypes = [C.POINTER(LearnConfig)]
    lib.azb_learn_config_default.restype = None
    for name, argtypes in sigs.items():
        if "AZB200_LIB" in os.environ and not hasattr(lib, name):
            continue  # an older build loaded for a kernel-variant A/B (scripts/ab.sh): calls to it fail at use
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.argtypes = argtypes
        fn.restype = C.c_int
    return lib, sorted(sigs) + ["azb_last_error", "azb_config_default", "azb_learn_config_default"]


lib, ABI_SYMBOLS = _load()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc != AZB_OK:
        raise AzbError(rc, lib.azb_last_error().decode())


def device_count():
    return lib.azb_device_count()


def release_caches():
    """Free the calling thread's evaluation cache (~2.7 GB of device memory kept between network runs)."""
    _check(lib.azb_release_caches())


class PinnedArray:
    """numpy view over page-locked host memory (azb_host_alloc)."""

    def __init__(self, shape, dtype=np.float32):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = C.c_void_p()
        _check(lib.azb_host_alloc(n, C.byref(self._p)))
        buf = (C.c_char * max(n, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        if self._p and lib is not None:
            self.array = None
            lib.azb_host_free(self._p)
            self._p = C.c_void_p()

    __del__ = close


def selftest_arith():
    """{reciprocal, sqrt, division} mismatches of the fast f32 helpers vs the IEEE intrinsics."""
    out = np.zeros(4, np.uint64)
    _check(lib.azb_selftest_arith(_ptr(out)))
    return out.tolist()


def default_config(**kw):
    cfg = Config()
    lib.azb_config_default(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def _states(a):
    a = np.ascontiguousarray(a, dtype=STATE_DTYPE)
    return a.reshape(-1)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class ConnectFourGame:
    """Batched `impl Game for ConnectFourGame` (connect_four_game.rs:81-238) on the device.

    Every method takes/returns numpy arrays of STATE_DTYPE (the reference struct without the
    redundant `heights`), n states per call.
    """

    HEIGHT, WIDTH, ACTIONS = 6, 7, 7

    @staticmethod
    def get_init_board(n=1):
        out = np.zeros(n, STATE_DTYPE)
        _check(lib.azb_c4_init(_ptr(out), n))
        return out

    @staticmethod
    def get_feature_shape():
        out = (C.c_size_t * 3)()
        _check(lib.azb_c4_feature_shape(out))
        return list(out)

    @staticmethod
    def get_next_state(states, player, action):
        s = _states(states)
        n = len(s)
        player = np.ascontiguousarray(np.broadcast_to(np.asarray(player, np.int8), (n,)))
        action = np.ascontiguousarray(np.broadcast_to(np.asarray(action, np.uint8), (n,)))
        out = np.zeros(n, STATE_DTYPE)
        nxt = np.zeros(n, np.int8)
        _check(lib.azb_c4_next_state(_ptr(s), _ptr(player), _ptr(action), n, _ptr(out), _ptr(nxt)))
        return out, nxt

    @staticmethod
    def get_valid_moves(states, player=1):
        s = _states(states)
        out = np.zeros((len(s), 7), np.uint8)
        _check(lib.azb_c4_valid_moves(_ptr(s), len(s), _ptr(out)))
        return out

    @staticmethod
    def get_game_ended(states, player, quirks=PROFILE_SANE):
        s = _states(states)
        n = len(s)
        player = np.ascontiguousarray(np.broadcast_to(np.asarray(player, np.int8), (n,)))
        out = np.zeros(n, np.float32)
        _check(lib.azb_c4_game_ended(_ptr(s), _ptr(player), n, quirks, _ptr(out)))
        return out

    @staticmethod
    def get_canonical_form(states, player):
        s = _states(states)
        n = len(s)
        player = np.ascontiguousarray(np.broadcast_to(np.asarray(player, np.int8), (n,)))
        out = np.zeros(n, STATE_DTYPE)
        _check(lib.azb_c4_canonical_form(_ptr(s), _ptr(player), n, _ptr(out)))
        return out

    @staticmethod
    def get_symmetries(states, pi):
        s = _states(states)
        n = len(s)
        pi = np.ascontiguousarray(pi, np.float32).reshape(n, 7)
        out_s = np.zeros((n, 2), STATE_DTYPE)
        out_pi = np.zeros((n, 2, 7), np.float32)
        _check(lib.azb_c4_symmetries(_ptr(s), _ptr(pi), n, _ptr(out_s), _ptr(out_pi)))
        return out_s, out_pi

    @staticmethod
    def eval_heuristic(states):
        s = _states(states)
        out = np.zeros(len(s), np.float32)
        _check(lib.azb_c4_eval_heuristic(_ptr(s), len(s), _ptr(out)))
        return out

    @staticmethod
    def to_features(states):
        s = _states(states)
        out = np.zeros((len(s), 2, 6, 7), np.float32)
        _check(lib.azb_c4_to_features(_ptr(s), len(s), _ptr(out)))
        return out


class AsyncMcts:
    """n_trees independent search trees on the device (AsyncMcts::default, async_mcts.rs:27-48)."""

    def __init__(self, n_trees=1, **cfg):
        self.cfg = cfg.pop("config", None) or default_config(**cfg)
        self.n_trees = n_trees
        self._h = C.c_void_p()
        _check(lib.azb_mcts_create(C.byref(self.cfg), n_trees, C.byref(self._h)))

    def close(self):
        if self._h and lib is not None:
            lib.azb_mcts_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def get_action_prob(self, states, temp):
        """async_mcts.rs:74-115 for every tree: returns (counts[n,7] u16, pi[n,7] f32)."""
        s = _states(states)
        assert len(s) == self.n_trees
        counts = np.zeros((self.n_trees, 7), np.uint16)
        pi = np.zeros((self.n_trees, 7), np.float32)
        _check(lib.azb_mcts_get_action_prob(self._h, _ptr(s), C.c_float(temp), _ptr(counts), _ptr(pi)))
        return counts, pi

    def counter_of(self, states):
        s = _states(states)
        assert len(s) == self.n_trees
        out = np.zeros(self.n_trees, np.uint64)
        _check(lib.azb_mcts_counter_of(self._h, _ptr(s), _ptr(out)))
        return out

    def stats(self):
        out = np.zeros((self.n_trees, 8), np.uint64)
        _check(lib.azb_mcts_stats(self._h, _ptr(out)))
        return out

    def dump(self, tree=0, cap=1 << 20):
        keys = np.zeros(cap, np.uint64)
        counters = np.zeros(cap, np.uint64)
        e = np.zeros(cap, np.float32)
        p = np.zeros((cap, 7), np.float32)
        hp = np.zeros(cap, np.uint8)
        n = C.c_uint64()
        _check(lib.azb_mcts_dump(self._h, tree, cap, _ptr(keys), _ptr(counters), _ptr(e), _ptr(p), _ptr(hp), C.byref(n)))
        n = min(n.value, cap)
        order = np.argsort(keys[:n])
        return keys[:n][order], counters[:n][order], e[:n][order], p[:n][order], hp[:n][order]


class Coach:
    """Coach::setup + the self-play half of Coach::learn (coach.rs:38-157, 241-272)."""

    def __init__(self, nnet=None, **cfg):
        self.cfg = cfg.pop("config", None) or default_config(**cfg)
        self._h = C.c_void_p()
        _check(lib.azb_coach_setup(C.byref(self.cfg), C.byref(self._h)))
        self.n_games = 0
        self.nnet = nnet
        if nnet is not None:
            _check(lib.azb_coach_set_nnet(self._h, nnet._h))

    @classmethod
    def setup(cls, checkpoint_directory, mcts_reserve_size, update_threshold, temp_threshold,
              max_history_length, max_queue_length, inference_batch_size, num_episode_threads,
              num_arena_games, num_iters, num_eps, num_sims, num_sim_threads, max_depth, cpuct, **extra):
        """Same 15 positional parameters as Coach::setup (coach.rs:38-54)."""
        return cls(checkpoint_directory=str(checkpoint_directory).encode(),
                   mcts_reserve_size=mcts_reserve_size, update_threshold=update_threshold,
                   temp_threshold=temp_threshold, max_history_length=max_history_length,
                   max_queue_length=max_queue_length, inference_batch_size=inference_batch_size,
                   num_episode_threads=num_episode_threads, num_arena_games=num_arena_games,
                   num_iters=num_iters, num_eps=num_eps, num_sims=num_sims,
                   num_sim_threads=num_sim_threads, max_depth=max_depth, cpuct=cpuct, **extra)

    def close(self):
        if self._h and lib is not None:
            lib.azb_coach_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def self_play(self, n_games, first_game_id=0):
        """n_games concurrent execute_episode calls (coach.rs:104-157); returns the stats dict."""
        st = SelfPlayStats()
        _check(lib.azb_coach_self_play(self._h, n_games, first_game_id, C.byref(st)))
        self.n_games = n_games
        return st.as_dict()

    def self_play_begin(self, n_games, first_game_id=0):
        """Launch a self-play call and return (azb_coach_self_play_begin); self_play_end() collects it.  Two coaches used in
        turn overlap one batch's tail with the next batch's head."""
        _check(lib.azb_coach_self_play_begin(self._h, n_games, first_game_id))
        self._begun = n_games

    def self_play_end(self):
        st = SelfPlayStats()
        _check(lib.azb_coach_self_play_end(self._h, C.byref(st)))
        self.n_games = self._begun
        return st.as_dict()

    def span_mark(self):
        _check(lib.azb_coach_span_mark(self._h))

    def span_ms(self, last):
        """Device milliseconds (CUDA events) from this coach's span_mark() to the end of `last`'s most recent call."""
        ms = C.c_double()
        _check(lib.azb_coach_span_ms(self._h, last._h, C.byref(ms)))
        return ms.value

    def execute_episode(self, episode_id=0):
        """One game; returns its SOA samples like the reference's VecDeque<TrainingSample>."""
        self.self_play(1, episode_id)
        return self.export_samples()

    def traces(self):
        g = self.n_games
        actions = np.zeros((g, 64), np.uint8)
        counts = np.zeros((g, 64, 7), np.uint16)
        plies = np.zeros(g, np.uint32)
        final_r = np.zeros(g, np.float32)
        final_player = np.zeros(g, np.int8)
        _check(lib.azb_coach_traces(self._h, _ptr(actions), _ptr(counts), _ptr(plies), _ptr(final_r), _ptr(final_player)))
        return dict(actions=actions, counts=counts, plies=plies, final_r=final_r, final_player=final_player)

    def ply_times(self):
        out = np.zeros((self.n_games, 64), np.uint64)
        _check(lib.azb_coach_ply_times(self._h, _ptr(out)))
        return out

    def num_samples(self):
        n = C.c_uint64()
        _check(lib.azb_coach_num_samples(self._h, C.byref(n)))
        return n.value

    def export_samples(self, out=None):
        """SOATrainingSamples (nnet.rs:33): (boards[n,2,6,7], pis[n,7], vs[n])."""
        n = self.num_samples()
        if out is None:
            out = (np.zeros((n, 2, 6, 7), np.float32), np.zeros((n, 7), np.float32), np.zeros(n, np.float32))
        boards, pis, vs = out
        w = C.c_uint64()
        _check(lib.azb_coach_export_samples(self._h, _ptr(boards), _ptr(pis), _ptr(vs), len(vs), C.byref(w)))
        return boards[: w.value], pis[: w.value], vs[: w.value]


    # ---- Coach::learn and the sample history (coach.rs:55-81,159-396) ----
    def learn(self, net_cfg=None, skip_first_play=False, epochs=10, batch_size=64, lr=1e-4, arena_k_open=None,
              save_files=True, seed=7, blocks=6, dist=None, arena_shared_trees=False):
        """Coach::learn(checkpoint, skip_first_play, ...) — coach.rs:169-396.  Returns (reports, accepted NNet).
        With `dist` (an initialised torch.distributed, one process per GPU) the iteration is data parallel
        (azb_coach_learn_dist): self-play and arena games sharded over the ranks, gradients averaged by all-reduce."""
        if net_cfg is None:
            net_cfg = NnetConfig(self.cfg.device, blocks, NNET_BF16_TC, 0, seed)
        lc = LearnConfig()
        lib.azb_learn_config_default(C.byref(lc))
        lc.epochs, lc.batch_size = epochs, batch_size
        if arena_k_open is not None:
            lc.arena_k_open = arena_k_open
        lc.arena_shared_trees = int(arena_shared_trees)
        lc.adam.lr = lr
        lc.skip_first_play, lc.save_files = int(skip_first_play), int(save_files)
        n_it = int(self.cfg.num_iters)
        reports = (LearnReport * max(1, n_it))()
        n = C.c_uint64()
        h = C.c_void_p()
        if isinstance(dist, Comm):
            d = dist.dist()
            _check(lib.azb_coach_learn_dist(self._h, C.byref(net_cfg), C.byref(lc), C.byref(d), reports, n_it, C.byref(n), C.byref(h)))
        elif dist is not None and dist.get_world_size() > 1:
            d = make_dist(dist, self.cfg.device)
            _check(lib.azb_coach_learn_dist(self._h, C.byref(net_cfg), C.byref(lc), C.byref(d), reports, n_it, C.byref(n), C.byref(h)))
        else:
            _check(lib.azb_coach_learn(self._h, C.byref(net_cfg), C.byref(lc), reports, n_it, C.byref(n), C.byref(h)))
        net = NNet.__new__(NNet)
        net.cfg, net._h = net_cfg, h
        return [reports[i].as_dict() for i in range(n.value)], net

    def history(self):
        """struct Coach.history (coach.rs:19): (counts per entry, boards, pis, vs) over all entries, oldest first."""
        n_it, n_s = C.c_uint64(), C.c_uint64()
        _check(lib.azb_coach_history_stat(self._h, C.byref(n_it), None, 0, C.byref(n_s)))
        counts = np.zeros(max(1, n_it.value), np.uint64)
        _check(lib.azb_coach_history_stat(self._h, C.byref(n_it), _ptr(counts), len(counts), C.byref(n_s)))
        n = n_s.value
        boards, pis, vs = np.zeros((n, 2, 6, 7), np.float32), np.zeros((n, 7), np.float32), np.zeros(n, np.float32)
        if n:
            _check(lib.azb_coach_history_export(self._h, _ptr(boards), _ptr(pis), _ptr(vs), n))
        return counts[: n_it.value], boards, pis, vs

    def save_train_examples(self, iteration, checkpoint):
        """Coach::save_train_examples — coach.rs:159-167."""
        _check(lib.azb_coach_save_train_examples(self._h, iteration, str(checkpoint).encode()))

    def load_train_examples(self, path):
        _check(lib.azb_coach_load_train_examples(self._h, str(path).encode()))


def examples_write(path, counts, boards, pis, vs):
    """`<iteration>.examples`: bincode of VecDeque<VecDeque<TrainingSample>> (coach.rs:159-167; layout in azb200.h)."""
    counts = np.ascontiguousarray(counts, np.uint64)
    boards = np.ascontiguousarray(boards, np.float32)
    pis = np.ascontiguousarray(pis, np.float32)
    vs = np.ascontiguousarray(vs, np.float32)
    assert int(counts.sum()) == len(vs) == len(pis.reshape(-1, 7)) == len(boards.reshape(-1, 84))
    _check(lib.azb_examples_write(str(path).encode(), len(counts), _ptr(counts), _ptr(boards), _ptr(pis), _ptr(vs)))


def examples_read(path):
    """-> (counts per history entry, boards[n,2,6,7], pis[n,7], vs[n])  (coach.rs:55-77)"""
    n_it, n_s = C.c_uint64(), C.c_uint64()
    _check(lib.azb_examples_stat(str(path).encode(), C.byref(n_it), None, 0, C.byref(n_s)))
    counts = np.zeros(max(1, n_it.value), np.uint64)
    _check(lib.azb_examples_stat(str(path).encode(), C.byref(n_it), _ptr(counts), len(counts), C.byref(n_s)))
    n = n_s.value
    boards, pis, vs = np.zeros((max(n, 1), 2, 6, 7), np.float32), np.zeros((max(n, 1), 7), np.float32), np.zeros(max(n, 1), np.float32)
    _check(lib.azb_examples_read(str(path).encode(), _ptr(boards), _ptr(pis), _ptr(vs), n))
    return counts[: n_it.value], boards[:n], pis[:n], vs[:n]


def examples_latest(checkpoint_directory):
    it = C.c_uint64()
    _check(lib.azb_examples_latest(str(checkpoint_directory).encode(), C.byref(it)))
    return it.value


def learn_accept(nwins, pwins, update_threshold):
    """The accept rule of coach.rs:383-390."""
    return bool(lib.azb_learn_accept(nwins, pwins, C.c_float(update_threshold)))


def learn_shuffle_perm(seed, iteration, n):
    perm = np.zeros(max(n, 1), np.uint64)
    _check(lib.azb_learn_shuffle_perm(seed, iteration, n, _ptr(perm)))
    return perm[:n]


class NNet:
    """trait NNet (src/nnet.rs:35-45): new / predict (+ parameter access).  One handle = one model."""

    def __init__(self, seed=7, blocks=6, precision=NNET_BF16_TC, device=0):
        self.cfg = NnetConfig(device, blocks, precision, 0, seed)
        self._h = C.c_void_p()
        _check(lib.azb_nnet_create(C.byref(self.cfg), C.byref(self._h)))

    def close(self):
        if self._h and lib is not None:  # (lib is None while the interpreter shuts down)
            lib.azb_nnet_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def predict(self, boards, model_id=0):
        """boards[B,2,6,7] f32 -> (pi[B,7], v[B])  (nnet.rs:40-44)"""
        boards = np.ascontiguousarray(boards, np.float32).reshape(-1, 2, 6, 7)
        n = len(boards)
        pi = np.zeros((n, 7), np.float32)
        v = np.zeros(n, np.float32)
        _check(lib.azb_nnet_predict(self._h, _ptr(boards), n, model_id, _ptr(pi), _ptr(v)))
        return pi, v

    def train_begin(self, boards, pis, vs):
        """Forward + loss + backward of NNet::train (nnet.rs:38); returns (policy loss, value loss); gradients stay on the device."""
        boards = np.ascontiguousarray(boards, np.float32).reshape(-1, 2, 6, 7)
        pis = np.ascontiguousarray(pis, np.float32).reshape(-1, 7)
        vs = np.ascontiguousarray(vs, np.float32).reshape(-1)
        assert len(boards) == len(pis) == len(vs)
        loss = np.zeros(2, np.float32)
        _check(lib.azb_nnet_train_begin(self._h, _ptr(boards), _ptr(pis), _ptr(vs), len(vs), _ptr(loss)))
        return float(loss[0]), float(loss[1])

    def grads(self):
        out = np.zeros(self.num_params(), np.float32)
        _check(lib.azb_nnet_grads(self._h, _ptr(out), len(out)))
        return out

    def grads_tensor(self):
        """The device gradient buffer as a torch tensor (no copy; valid until the next train_begin)."""
        import torch
        p, n = C.c_void_p(), C.c_uint64()
        _check(lib.azb_nnet_grads_device(self._h, C.byref(p), C.byref(n)))

        class _Dev:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "<f4", "data": (int(p.value), False), "version": 2}
        return torch.as_tensor(_Dev(), device=f"cuda:{self.cfg.device}")

    def set_grads(self, g):
        g = np.ascontiguousarray(g, np.float32)
        _check(lib.azb_nnet_set_grads(self._h, _ptr(g), len(g)))

    def train_apply(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        cfg = TrainConfig(lr, beta1, beta2, eps)
        _check(lib.azb_nnet_train_apply(self._h, C.byref(cfg)))

    def train(self, samples, lr=1e-3, dist=None, **adam):
        """NNet::train on SOATrainingSamples (boards, pis, vs).  With `dist` (an initialised torch.distributed) the step is
        data parallel: every rank trains on its own shard, the gradients are averaged with all_reduce before Adam."""
        loss = self.train_begin(*samples)
        if dist is not None and dist.get_world_size() > 1:
            if dist.get_backend() == "nccl":  # in place on the library's device buffer: NCCL over NVLink, no host round trip
                g = self.grads_tensor()
                dist.all_reduce(g)
                g /= dist.get_world_size()
            else:
                self.set_grads(sharding.average_gradients(dist, self.grads()))
        self.train_apply(lr=lr, **adam)
        return loss

    def conv_hook(self, layer, mode, x, residual=None, mask=None):
        """Tower convolution `layer` on the tensor cores: mode 0 forward, mode 1 backward data (see azb200.h).
        x / residual / mask: f32 [n, 42, 128] (cell-major, channel-minor); returns the same shape."""
        x = np.ascontiguousarray(x, np.float32)
        res = None if residual is None else np.ascontiguousarray(residual, np.float32)
        msk = None if mask is None else np.ascontiguousarray(mask, np.float32)
        out = np.zeros_like(x)
        _check(lib.azb_nnet_conv_hook(self._h, layer, mode, _ptr(x), _ptr(res) if res is not None else None,
                                      _ptr(msk) if msk is not None else None, len(x), _ptr(out)))
        return out

    def wgrad_hook(self, x, dz):
        """dW[9, 128 ci, 128 co] of a tower convolution from its input x and the gradient dz of its pre-activation."""
        x = np.ascontiguousarray(x, np.float32)
        dz = np.ascontiguousarray(dz, np.float32)
        dw = np.zeros((9, 128, 128), np.float32)
        _check(lib.azb_nnet_wgrad_hook(self._h, _ptr(x), _ptr(dz), len(x), _ptr(dw)))
        return dw

    def benchmark(self, batch, iters=10):
        """Device-only mean milliseconds per forward pass over `batch` resident positions."""
        ms = C.c_double()
        _check(lib.azb_nnet_benchmark(self._h, batch, iters, C.byref(ms)))
        return ms.value

    def num_params(self):
        n = C.c_uint64()
        _check(lib.azb_nnet_num_params(self._h, C.byref(n)))
        return n.value

    def get_params(self):
        out = np.zeros(self.num_params(), np.float32)
        _check(lib.azb_nnet_get_params(self._h, _ptr(out), len(out)))
        return out

    def set_params(self, w):
        w = np.ascontiguousarray(w, np.float32)
        _check(lib.azb_nnet_set_params(self._h, _ptr(w), len(w)))

    def save(self, path):
        """`<model_id>.azbw` weight checkpoint (python_nnet.rs:76-79 save_checkpoint)."""
        _check(lib.azb_nnet_save(self._h, str(path).encode()))

    def load(self, path):
        _check(lib.azb_nnet_load(self._h, str(path).encode()))

    def copy_from(self, other):
        _check(lib.azb_nnet_copy(self._h, other._h))


def param_layout(blocks=6, channels=128):
    """Offsets/shapes of the flat parameter vector (csrc/nnet.cuh NetLayout)."""
    c = channels
    spec = [("stem_w", (9, 2, c)), ("stem_b", (c,)), ("tower_w", (2 * blocks, 9, c, c)), ("tower_b", (2 * blocks, c)),
            ("pol_w", (c, 2)), ("pol_b", (2,)), ("pol_fc_w", (84, 7)), ("pol_fc_b", (7,)), ("val_w", (c,)),
            ("val_b", (1,)), ("val_fc1_w", (42, 64)), ("val_fc1_b", (64,)), ("val_fc2_w", (64,)), ("val_fc2_b", (1,))]
    out, o = {}, 0
    for name, shape in spec:
        n = int(np.prod(shape))
        out[name] = (o, shape)
        o += n
    out["total"] = o
    return out


def arena_play_games(num, eval_a, eval_b, net_a=None, net_b=None, k_open=0, **cfg):
    """arena::play_games (src/arena.rs:62-99) with two MCTS players; returns ((win, loss, draw), results, stats)."""
    config = cfg.pop("config", None) or default_config(**cfg)
    counts = np.zeros(3, np.uint64)
    results = np.zeros(max(1, 2 * (num // 2)), np.int8)
    st = SelfPlayStats()
    _check(lib.azb_arena_play_games(C.byref(config), num, eval_a, eval_b, net_a._h if net_a else None,
                                    net_b._h if net_b else None, k_open, _ptr(counts), _ptr(results), C.byref(st)))
    return tuple(int(x) for x in counts), results[: 2 * (num // 2)], st.as_dict()


def arena_play_games_traced(num, eval_a, eval_b, net_a=None, net_b=None, k_open=0, shared_trees=0, first_game_id=0, **cfg):
    """azb_arena_play_games_ex: the match with its options (shared_trees = the reference's persistent tree pair,
    coach.rs:333-354) and the per-game traces.  Returns ((win, loss, draw), results, stats,
    dict(actions[G, 64], counts[G, 64, 7], plies[G]))."""
    config = cfg.pop("config", None) or default_config(**cfg)
    g = max(1, 2 * (num // 2))
    counts = np.zeros(3, np.uint64)
    results = np.zeros(g, np.int8)
    actions = np.full((g, 64), 0xFF, np.uint8)
    rc = np.zeros((g, 64, 7), np.uint16)
    plies = np.zeros(g, np.uint32)
    st = SelfPlayStats()
    opts = ArenaOpts(k_open, shared_trees, first_game_id)
    _check(lib.azb_arena_play_games_ex(C.byref(config), num, eval_a, eval_b, net_a._h if net_a else None,
                                       net_b._h if net_b else None, C.byref(opts), _ptr(counts), _ptr(results),
                                       _ptr(actions), _ptr(rc), _ptr(plies), C.byref(st)))
    n = 2 * (num // 2)
    return (tuple(int(x) for x in counts), results[:n], st.as_dict(),
            dict(actions=actions[:n], counts=rc[:n], plies=plies[:n]))
