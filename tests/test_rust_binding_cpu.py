"""rust/azb200-sys is uncompiled source (no Rust toolchain in the image).  These checks keep it aligned with the C ABI:
its extern block declares exactly the header's symbols with the same parameter counts, and its #[repr(C)] structs list
the same fields, in the same order and with the same widths, as the ctypes mirror the GPU tests call through."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RS = open(os.path.join(ROOT, "rust", "azb200-sys", "src", "lib.rs")).read()


def header_functions():
    src = open(os.path.join(ROOT, "include", "azb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(azb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        args = args.strip()
        out[name] = 0 if args in ("", "void") else len(args.split(","))
    return out


def rust_functions():
    block = RS[RS.index('extern "C" {'):]
    out = {}
    for name, args in re.findall(r"pub fn (azb_[a-z0-9_]+)\s*\(([^)]*)\)", block):
        out[name] = 0 if not args.strip() else len([a for a in args.split(",") if a.strip()])
    return out


def rust_struct(name):
    body = re.search(r"pub struct %s \{(.*?)\n\}" % name, RS, flags=re.S).group(1)
    return re.findall(r"pub ([a-z0-9_]+): ([^,\n]+),", body)


RUST_SIZES = {"u64": 8, "i32": 4, "u32": 4, "f32": 4, "f64": 8, "[f32; 2]": 8, "*const c_char": 8, "azb_train_config": 16, "i8": 1}


def test_extern_block_matches_header():
    h, r = header_functions(), rust_functions()
    assert set(h) == set(r), (sorted(set(h) - set(r)), sorted(set(r) - set(h)))
    assert len(h) >= 60
    for name in h:
        assert h[name] == r[name], (name, h[name], r[name])


def test_structs_match_ctypes_mirror(azb):
    pairs = [("azb_config", azb.Config), ("azb_selfplay_stats", azb.SelfPlayStats), ("azb_nnet_config", azb.NnetConfig),
             ("azb_train_config", azb.TrainConfig), ("azb_learn_config", azb.LearnConfig), ("azb_learn_report", azb.LearnReport),
             ("azb_arena_opts", azb.ArenaOpts)]
    for rname, cls in pairs:
        rf = rust_struct(rname)
        assert [n for n, _ in rf] == [n for n, _ in cls._fields_], rname
        for (n, rt), (_, ct) in zip(rf, cls._fields_):
            assert RUST_SIZES[rt.strip()] == C.sizeof(ct), (rname, n, rt)
    st = rust_struct("azb_c4_state")
    assert st == [("s", "[[i8; 7]; 6]"), ("me", "i8")] and "#[repr(C, packed)]" in RS


def test_safe_wrapper_uses_only_declared_symbols():
    src = open(os.path.join(ROOT, "rust", "azb200", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(azb_[a-z0-9_]+)\s*\(", src))
    assert used and used <= set(rust_functions())
