"""ctypes binding of the CPU oracle (oracle/_build/libazoracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_build", "libazoracle.so")
UNIT_BIN = os.path.join(ROOT, "oracle", "_build", "test_reference_units")

STATE_DTYPE = np.dtype([("s", np.int8, (6, 7)), ("me", np.int8)])
EVAL_UNIFORM, EVAL_HASH, EVAL_CALLBACK = 0, 1, 2
PREDICT_FN = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)


class Params(C.Structure):
    _fields_ = [("mcts_reserve_size", C.c_uint64), ("temp_threshold", C.c_uint64),
                ("num_sims", C.c_uint64), ("max_depth", C.c_uint64), ("cpuct", C.c_int32),
                ("quirks", C.c_uint32), ("seed", C.c_uint64), ("num_sim_threads", C.c_uint64)]


def params(num_sims=25, quirks=0, seed=1, temp_threshold=15, max_depth=1000, cpuct=1, reserve=None, num_sim_threads=1):
    if reserve is None:
        reserve = max(4096, int(num_sims) * 8 * 64)
    return Params(reserve, temp_threshold, num_sims, max_depth, cpuct, quirks, seed, num_sim_threads)


L = C.CDLL(LIB_PATH)
vp = C.c_void_p
L.azo_last_error.restype = C.c_char_p
L.azo_counter_init.restype = C.c_uint64
L.azo_counter_visit.restype = C.c_uint64
L.azo_counter_visit.argtypes = [C.c_uint64]
L.azo_counter_unvisit.restype = C.c_uint64
L.azo_counter_unvisit.argtypes = [C.c_uint64, C.c_float, C.c_float, C.c_uint32]
L.azo_counter_read.argtypes = [C.c_uint64, C.c_float, vp, vp, vp, vp]
L.azo_uniform01.restype = C.c_float
L.azo_uniform01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
L.azo_choose_weighted.argtypes = [vp, C.c_size_t, C.c_float]
L.azo_c4_next_state.argtypes = [vp, vp, vp, C.c_size_t, vp, vp]
L.azo_c4_valid_moves.argtypes = [vp, C.c_size_t, vp]
L.azo_c4_game_ended.argtypes = [vp, vp, C.c_size_t, C.c_uint32, vp]
L.azo_c4_canonical_form.argtypes = [vp, vp, C.c_size_t, vp]
L.azo_c4_symmetries.argtypes = [vp, vp, C.c_size_t, vp, vp]
L.azo_c4_to_features.argtypes = [vp, C.c_size_t, vp]
L.azo_c4_key.argtypes = [vp, C.c_size_t, vp]
L.azo_mcts_create.restype = vp
L.azo_mcts_create.argtypes = [vp, C.POINTER(Params), C.c_int, vp, vp]
L.azo_mcts_destroy.argtypes = [vp]
L.azo_mcts_get_action_prob.argtypes = [vp, vp, C.c_float, vp, vp]
L.azo_mcts_set_num_sims.argtypes = [vp, C.c_uint64]
L.azo_mcts_len.restype = C.c_uint64
L.azo_mcts_len.argtypes = [vp]
L.azo_mcts_seen_len.restype = C.c_uint64
L.azo_mcts_seen_len.argtypes = [vp]
L.azo_mcts_stats.argtypes = [vp, vp]
L.azo_mcts_counter_of.restype = C.c_uint64
L.azo_mcts_counter_of.argtypes = [vp, vp]
L.azo_mcts_dump.restype = C.c_uint64
L.azo_mcts_dump.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, vp]
L.azo_execute_episode.argtypes = [C.POINTER(Params), C.c_uint64, C.c_int, vp, vp] + [vp] * 11
L.azo_arena_play_games.argtypes = [C.POINTER(Params), C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint32, vp, vp]
L.azo_bench_selfplay.argtypes = [C.POINTER(Params), C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp]
L.azo_arena_play_games_cb.argtypes = [C.POINTER(Params), C.c_uint64, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, C.c_uint32,
                                      C.c_uint64, vp, vp, vp, vp, vp]
L.azo_bench_selfplay_net.argtypes = [C.POINTER(Params), C.c_int, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                     C.c_uint64, vp, vp, vp, vp]
L.azo_cpu_net_predict.argtypes = [C.c_int, vp, C.c_uint64, vp, C.c_uint64, vp, vp]


def _p(a):
    return a.ctypes.data_as(vp)


def _states(a):
    return np.ascontiguousarray(a, dtype=STATE_DTYPE).reshape(-1)


def _bcast(x, n, dt):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(x, dt), (n,)))


def init_board(n=1):
    s = np.zeros(n, STATE_DTYPE)
    s["me"] = 1
    return s


def next_state(states, player, action):
    s = _states(states); n = len(s)
    out = np.zeros(n, STATE_DTYPE); nxt = np.zeros(n, np.int8)
    L.azo_c4_next_state(_p(s), _p(_bcast(player, n, np.int8)), _p(_bcast(action, n, np.uint8)), n, _p(out), _p(nxt))
    return out, nxt


def valid_moves(states):
    s = _states(states); out = np.zeros((len(s), 7), np.uint8)
    L.azo_c4_valid_moves(_p(s), len(s), _p(out))
    return out


def game_ended(states, player, quirks=0):
    s = _states(states); n = len(s); out = np.zeros(n, np.float32)
    L.azo_c4_game_ended(_p(s), _p(_bcast(player, n, np.int8)), n, quirks, _p(out))
    return out


def canonical_form(states, player):
    s = _states(states); n = len(s); out = np.zeros(n, STATE_DTYPE)
    L.azo_c4_canonical_form(_p(s), _p(_bcast(player, n, np.int8)), n, _p(out))
    return out


def symmetries(states, pi):
    s = _states(states); n = len(s)
    pi = np.ascontiguousarray(pi, np.float32).reshape(n, 7)
    os_ = np.zeros((n, 2), STATE_DTYPE); op = np.zeros((n, 2, 7), np.float32)
    L.azo_c4_symmetries(_p(s), _p(pi), n, _p(os_), _p(op))
    return os_, op


def to_features(states):
    s = _states(states); out = np.zeros((len(s), 2, 6, 7), np.float32)
    L.azo_c4_to_features(_p(s), len(s), _p(out))
    return out


def state_key(states):
    s = _states(states); out = np.zeros(len(s), np.uint64)
    L.azo_c4_key(_p(s), len(s), _p(out))
    return out


def counter_read(c, scale=100.0):
    w = C.c_float(); n = C.c_uint16(); vl = C.c_uint16(); q = C.c_float()
    L.azo_counter_read(c, scale, C.byref(w), C.byref(n), C.byref(vl), C.byref(q))
    return w.value, n.value, vl.value, q.value


class Mcts:
    """AsyncMcts<C4> of the oracle (deterministic mode)."""

    def __init__(self, num_sims=25, quirks=0, evaluator=EVAL_UNIFORM, root=None, callback=None, **kw):
        self.p = params(num_sims=num_sims, quirks=quirks, **kw)
        self._cb = make_callback(callback) if callback else None
        fn = C.cast(self._cb, vp) if self._cb else None
        r = _p(_states(root)) if root is not None else None
        self._h = L.azo_mcts_create(r, C.byref(self.p), evaluator, fn, None)
        if not self._h:
            raise RuntimeError(L.azo_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            L.azo_mcts_destroy(self._h)
            self._h = None

    __del__ = close

    def get_action_prob(self, state, temp):
        s = _states(state)
        counts = np.zeros(7, np.uint16); pi = np.zeros(7, np.float32)
        if L.azo_mcts_get_action_prob(self._h, _p(s), C.c_float(temp), _p(counts), _p(pi)) != 0:
            raise RuntimeError(L.azo_last_error().decode())
        return counts, pi

    def counter_of(self, state):
        return int(L.azo_mcts_counter_of(self._h, _p(_states(state))))

    def stats(self):
        out = np.zeros(8, np.uint64)
        L.azo_mcts_stats(self._h, _p(out))
        out[6] = L.azo_mcts_len(self._h)
        out[7] = L.azo_mcts_seen_len(self._h)
        return out

    def dump(self):
        cap = int(L.azo_mcts_seen_len(self._h))
        keys = np.zeros(cap, np.uint64); counters = np.zeros(cap, np.uint64)
        e = np.zeros(cap, np.float32); p = np.zeros((cap, 7), np.float32); hp = np.zeros(cap, np.uint8)
        n = L.azo_mcts_dump(self._h, cap, _p(keys), _p(counters), _p(e), _p(p), _p(hp))
        assert n == cap
        order = np.argsort(keys)
        return keys[order], counters[order], e[order], p[order], hp[order]


def make_callback(predict):
    """Wrap predict(boards[B,2,6,7]) -> (pi[B,7], v[B]) as the oracle's evaluator callback
    (mirrors trait NNet::predict, nnet.rs:40-44)."""
    def cb(boards, batch, pi, v, user):
        b = np.ctypeslib.as_array(boards, shape=(batch, 2, 6, 7)).copy()
        p_, v_ = predict(b)
        np.ctypeslib.as_array(pi, shape=(batch, 7))[:] = p_
        np.ctypeslib.as_array(v, shape=(batch,))[:] = v_
    return PREDICT_FN(cb)


def execute_episode(num_sims=25, quirks=0, seed=1, episode_id=0, evaluator=EVAL_UNIFORM, callback=None, **kw):
    """Coach::execute_episode of the oracle; returns a dict with the full trace."""
    p = params(num_sims=num_sims, quirks=quirks, seed=seed, **kw)
    cb = make_callback(callback) if callback else None
    fn = C.cast(cb, vp) if cb else None
    actions = np.full(64, 0xFF, np.uint8); counts = np.zeros((64, 7), np.uint16)
    boards = np.zeros((128, 2, 6, 7), np.float32); pis = np.zeros((128, 7), np.float32); vs = np.zeros(128, np.float32)
    ns = C.c_uint64(); fr = C.c_float(); fp = C.c_int8(); st = np.zeros(6, np.uint64)
    nl = C.c_uint64(); sl = C.c_uint64()
    plies = L.azo_execute_episode(C.byref(p), episode_id, evaluator, fn, None, _p(actions), _p(counts), _p(boards),
                                  _p(pis), _p(vs), C.addressof(ns), C.addressof(fr), C.addressof(fp), _p(st),
                                  C.addressof(nl), C.addressof(sl))
    if plies < 0:
        raise RuntimeError(L.azo_last_error().decode())
    n = ns.value
    return dict(plies=plies, actions=actions, counts=counts, boards=boards[:n], pis=pis[:n], vs=vs[:n],
                final_r=fr.value, final_player=fp.value, stats=st, nodes_len=nl.value, seen_len=sl.value)


def arena_play_games(num, eval_a, eval_b, num_sims=25, quirks=0, seed=1, shared_trees=0, k_open=0, **kw):
    p = params(num_sims=num_sims, quirks=quirks, seed=seed, **kw)
    out = np.zeros(3, np.uint64); res = np.zeros(num, np.int8)
    if L.azo_arena_play_games(C.byref(p), num, eval_a, eval_b, shared_trees, k_open, _p(out), _p(res)) != 0:
        raise RuntimeError(L.azo_last_error().decode())
    return out, res[: 2 * (num // 2)]


def arena_play_games_traced(num, eval_a, eval_b, callback_a=None, callback_b=None, num_sims=25, quirks=0, seed=1,
                            shared_trees=0, k_open=0, first_game_id=0, **kw):
    """arena::play_games of the oracle with per-game traces; EVAL_CALLBACK players search with callback_a / callback_b
    as NNet::predict.  Returns (counts[3], results, dict(actions[G, 64], counts[G, 64, 7], plies[G]))."""
    p = params(num_sims=num_sims, quirks=quirks, seed=seed, **kw)
    G = 2 * (num // 2)
    out = np.zeros(3, np.uint64); res = np.zeros(max(G, 1), np.int8)
    actions = np.full((max(G, 1), 64), 0xFF, np.uint8); rc = np.zeros((max(G, 1), 64, 7), np.uint16)
    plies = np.zeros(max(G, 1), np.uint32)
    cba = make_callback(callback_a) if callback_a else None
    cbb = make_callback(callback_b) if callback_b else None
    fa = C.cast(cba, vp) if cba else None
    fb = C.cast(cbb, vp) if cbb else None
    if L.azo_arena_play_games_cb(C.byref(p), num, eval_a, fa, None, eval_b, fb, None, shared_trees, k_open,
                                 first_game_id, _p(out), _p(res), _p(actions), _p(rc), _p(plies)) != 0:
        raise RuntimeError(L.azo_last_error().decode())
    return out, res[:G], dict(actions=actions[:G], counts=rc[:G], plies=plies[:G])


def bench_selfplay(n_games, n_threads, num_sims=800, quirks=0, seed=0xA1FA0, evaluator=EVAL_UNIFORM, first_game_id=0, **kw):
    p = params(num_sims=num_sims, quirks=quirks, seed=seed, **kw)
    sims = C.c_uint64(); plies = C.c_uint64(); secs = C.c_double(); lv = C.c_uint64(); ex = C.c_uint64()
    rc = L.azo_bench_selfplay(C.byref(p), evaluator, n_games, n_threads, first_game_id, C.addressof(sims),
                              C.addressof(plies), C.addressof(secs), C.addressof(lv), C.addressof(ex))
    if rc != 0:
        raise RuntimeError(L.azo_last_error().decode())
    return dict(sims=sims.value, plies=plies.value, seconds=secs.value, levels=lv.value, expansions=ex.value)


def bench_selfplay_net(params_vec, blocks, n_games, n_threads, num_sims=400, quirks=0, seed=0xA1FA0, first_game_id=0, max_plies=0, **kw):
    """Oracle self-play with the network evaluated on the CPU (oracle/nnet_cpu.hpp), one game per thread, every game cut
    after max_plies plies (0 = whole games): the CPU leg of BASELINE configs 1 and 3."""
    p = params(num_sims=num_sims, quirks=quirks, seed=seed, **kw)
    w = np.ascontiguousarray(params_vec, np.float32)
    sims = C.c_uint64(); plies = C.c_uint64(); secs = C.c_double(); ev = C.c_uint64()
    rc = L.azo_bench_selfplay_net(C.byref(p), blocks, _p(w), len(w), n_games, n_threads, first_game_id, max_plies,
                                  C.addressof(sims), C.addressof(plies), C.addressof(secs), C.addressof(ev))
    if rc != 0:
        raise RuntimeError(L.azo_last_error().decode())
    return dict(sims=sims.value, plies=plies.value, seconds=secs.value, evals=ev.value)


def cpu_net_predict(params_vec, blocks, boards):
    w = np.ascontiguousarray(params_vec, np.float32)
    b = np.ascontiguousarray(boards, np.float32).reshape(-1, 2, 6, 7)
    pi = np.zeros((len(b), 7), np.float32); v = np.zeros(len(b), np.float32)
    if L.azo_cpu_net_predict(blocks, _p(w), len(w), _p(b), len(b), _p(pi), _p(v)) != 0:
        raise RuntimeError(L.azo_last_error().decode())
    return pi, v


# ---- Coach::learn host decisions + the .examples encoding (oracle/learn.hpp) ----
L.azo_examples_encode.restype = C.c_uint64
L.azo_examples_encode.argtypes = [C.c_uint64, vp, vp, vp, vp, vp, C.c_uint64]
L.azo_learn_accept.argtypes = [C.c_uint64, C.c_uint64, C.c_float]
L.azo_learn_shuffle_perm.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, vp]
L.azo_learn_shuffle_perm.restype = None
L.azo_learn_window.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp]
L.azo_learn_window.restype = None


def examples_encode(counts, boards, pis, vs):
    """bincode::serialize(&history) (coach.rs:163) -> bytes"""
    counts = np.ascontiguousarray(counts, np.uint64)
    boards = np.ascontiguousarray(boards, np.float32)
    pis = np.ascontiguousarray(pis, np.float32)
    vs = np.ascontiguousarray(vs, np.float32)
    p = lambda a: a.ctypes.data_as(vp)
    n = L.azo_examples_encode(len(counts), p(counts), p(boards), p(pis), p(vs), None, 0)
    out = np.zeros(n, np.uint8)
    L.azo_examples_encode(len(counts), p(counts), p(boards), p(pis), p(vs), p(out), n)
    return out.tobytes()


def learn_accept(nwins, pwins, thr):
    return bool(L.azo_learn_accept(nwins, pwins, C.c_float(thr)))


def learn_shuffle_perm(seed, iteration, n):
    perm = np.zeros(max(n, 1), np.uint64)
    L.azo_learn_shuffle_perm(seed, iteration, n, perm.ctypes.data_as(vp))
    return perm[:n]


def learn_window(played, max_queue, max_hist):
    """-> (sizes[n][max_hist] of the history after each iteration, dropped[n])  (coach.rs:274-289)"""
    played = np.ascontiguousarray(played, np.uint64)
    sizes = np.zeros((len(played), max_hist), np.uint64)
    dropped = np.zeros(len(played), np.uint64)
    L.azo_learn_window(played.ctypes.data_as(vp), len(played), max_queue, max_hist, sizes.ctypes.data_as(vp),
                       dropped.ctypes.data_as(vp))
    return sizes, dropped
