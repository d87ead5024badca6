"""Host-side pieces of Coach::learn (coach.rs:169-396) that need no device: the `<n>.examples` file format
(coach.rs:55-81,159-167: bincode of VecDeque<VecDeque<TrainingSample>>), the accept rule (coach.rs:383-390) and the
shuffle permutation — product (libazb200.so) against the oracle (oracle/learn.hpp), an independent struct-level parser
written here from the format description, and the committed fixture tests/golden/tiny_examples.json."""
import json
import os
import struct

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def parse_examples(blob):
    """Independent reader: walks the bytes the way serde would drive bincode 1.x + ndarray 0.13."""
    at = 0

    def take(fmt):
        nonlocal at
        v = struct.unpack_from("<" + fmt, blob, at)
        at += struct.calcsize("<" + fmt)
        return v

    out = []
    (n_it,) = take("Q")
    for _ in range(n_it):
        (n,) = take("Q")
        entry = []
        for _ in range(n):
            (ver,) = take("B")
            assert ver == 1
            (nd,) = take("Q")
            shape = take(f"{nd}Q")
            (ln,) = take("Q")
            assert ln == int(np.prod(shape))
            board = np.array(take(f"{ln}f"), np.float32).reshape(shape)
            (ver,) = take("B")
            assert ver == 1
            (d0,) = take("Q")
            (ln,) = take("Q")
            assert d0 == ln
            pi = np.array(take(f"{ln}f"), np.float32)
            (v,) = take("f")
            entry.append((board, pi, np.float32(v)))
        out.append(entry)
    assert at == len(blob)
    return out


def random_history(rng, counts):
    n = int(sum(counts))
    boards = (rng.random((n, 2, 6, 7)) < 0.4).astype(np.float32)
    pis = rng.random((n, 7)).astype(np.float32)
    vs = rng.choice(np.array([-1.0, 1.0, 1e-4], np.float32), n)
    return boards, pis, vs


@pytest.mark.parametrize("counts", [[], [0], [1], [3, 0, 5], [7, 2049, 1, 0, 33]])
def test_examples_bytes_match_oracle_and_parser(azb, oracle, tmp_path, counts):
    rng = np.random.default_rng(len(counts) + 17)
    boards, pis, vs = random_history(rng, counts)
    path = tmp_path / "0.examples"
    azb.examples_write(path, counts, boards, pis, vs)
    blob = path.read_bytes()
    assert len(blob) == 8 + 8 * len(counts) + 426 * int(sum(counts))
    assert blob == oracle.examples_encode(counts, boards, pis, vs)  # bit-exact
    parsed = parse_examples(blob)
    assert [len(e) for e in parsed] == list(counts)
    k = 0
    for e in parsed:
        for board, pi, v in e:
            assert board.shape == (2, 6, 7)
            assert (board == boards[k]).all() and (pi == pis[k]).all() and v == vs[k]
            k += 1
    c2, b2, p2, v2 = azb.examples_read(path)
    assert c2.tolist() == list(counts)
    assert (b2 == boards).all() and (p2 == pis).all() and (v2 == vs).all()


def test_examples_golden_fixture(azb, tmp_path):
    g = json.load(open(os.path.join(HERE, "golden", "tiny_examples.json")))
    blob = bytes.fromhex(g["bincode_hex"])
    path = tmp_path / "3.examples"
    path.write_bytes(blob)
    counts, boards, pis, vs = azb.examples_read(path)
    assert counts.tolist() == g["counts"]
    assert (boards.reshape(len(vs), -1) == np.array(g["boards"], np.float32)).all()
    assert (pis == np.array(g["pis"], np.float32)).all() and (vs == np.array(g["vs"], np.float32)).all()
    out = tmp_path / "4.examples"
    azb.examples_write(out, counts, boards, pis, vs)
    assert out.read_bytes() == blob


def test_examples_reader_edge_cases(azb, tmp_path):
    rng = np.random.default_rng(5)
    boards, pis, vs = random_history(rng, [2])
    path = tmp_path / "0.examples"
    azb.examples_write(path, [2], boards, pis, vs)
    blob = path.read_bytes()
    # the literal Game::to_features writes [6,7,2] (H,W,C — F11) where the declared shape is [2,6,7]: such a file is what a
    # real reference run would leave behind.  The reader transposes it into the engine's channel-first planes (it used to
    # copy the 84 floats verbatim, interleaving the two planes per cell — ADVICE r1).
    hwc = np.ascontiguousarray(boards.reshape(2, 2, 6, 7).transpose(0, 2, 3, 1))  # the same samples, channel-last
    lit = bytearray(blob)
    for i in range(2):
        off = 8 + 8 + i * 426 + 1 + 8
        lit[off:off + 24] = struct.pack("<3Q", 6, 7, 2)
        lit[off + 24 + 8:off + 24 + 8 + 336] = hwc[i].tobytes()
    (tmp_path / "1.examples").write_bytes(bytes(lit))
    c, b, p, v = azb.examples_read(tmp_path / "1.examples")
    assert c.tolist() == [2] and np.array_equal(b.reshape(2, 2, 6, 7), boards.reshape(2, 2, 6, 7))
    assert np.array_equal(p, pis) and np.array_equal(v, vs)
    # any other shape is not a connect-four sample, even when it has 84 elements
    for dims in ((7, 6, 2), (2, 7, 6), (1, 1, 84), (3, 4, 7)):
        odd = bytearray(blob)
        off = 8 + 8 + 1 + 8
        odd[off:off + 24] = struct.pack("<3Q", *dims)
        (tmp_path / "2.examples").write_bytes(bytes(odd))
        with pytest.raises(azb.AzbError) as e:
            azb.examples_read(tmp_path / "2.examples")
        assert e.value.code == azb.ERR_INVALID
    for bad in (blob[:-1], blob + b"\0", blob[:8] + struct.pack("<Q", 3) + blob[16:], b"", blob[:100]):
        (tmp_path / "2.examples").write_bytes(bad)
        with pytest.raises(azb.AzbError) as e:
            azb.examples_read(tmp_path / "2.examples")
        assert e.value.code == azb.ERR_INVALID
    with pytest.raises(azb.AzbError):
        azb.examples_read(tmp_path / "missing.examples")


def test_examples_latest(azb, tmp_path):
    """Coach::setup picks the largest numeric stem (coach.rs:58-72); weight files may share the directory."""
    with pytest.raises(azb.AzbError):
        azb.examples_latest(tmp_path / "nope")
    with pytest.raises(azb.AzbError):
        azb.examples_latest(tmp_path)
    for name in ("0.examples", "9.examples", "10.examples", "3.azbw", "77.azbw", "x.examples", "notes.txt"):
        (tmp_path / name).write_bytes(b"")
    assert azb.examples_latest(tmp_path) == 10


def test_accept_rule_matches_oracle(azb, oracle):
    for thr in (0.0, 0.5, 0.55, 0.6, 1.0):
        for n in range(0, 12):
            for p in range(0, 12):
                assert azb.learn_accept(n, p, thr) == oracle.learn_accept(n, p, thr), (n, p, thr)
    assert not azb.learn_accept(0, 0, 0.0)      # coach.rs:383: no decisive game -> reject
    assert azb.learn_accept(6, 4, 0.6)          # 0.6 < 0.6 is false -> accept
    assert not azb.learn_accept(5, 4, 0.6)


def test_shuffle_perm_matches_oracle(azb, oracle):
    for seed, it, n in [(1, 0, 0), (1, 0, 1), (1, 0, 2), (7, 3, 1000), (0xA1FA0, 11, 4097), (2**40 + 5, 2, 333)]:
        a, b = azb.learn_shuffle_perm(seed, it, n), oracle.learn_shuffle_perm(seed, it, n)
        assert (a == b).all()
        assert sorted(a.tolist()) == list(range(n))
    assert (azb.learn_shuffle_perm(1, 0, 500) != azb.learn_shuffle_perm(1, 1, 500)).any()
    assert (azb.learn_shuffle_perm(1, 0, 500) != azb.learn_shuffle_perm(2, 0, 500)).any()


def test_window_restatement(oracle):
    """coach.rs:274-289 on counts: queue trim from the front, history window."""
    sizes, dropped = oracle.learn_window([10, 50, 0, 7], max_queue=20, max_hist=2)
    assert dropped.tolist() == [0, 30, 0, 0]
    assert sizes.tolist() == [[10, 0], [10, 20], [20, 0], [0, 7]]
