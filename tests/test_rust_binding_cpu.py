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


def test_integration_excerpt_matches_crate():
    """INTEGRATION.md shows struct and fn declarations a maintainer may copy: every one of them must be the crate's own
    text (a struct 16 bytes short would make the library write past it — ADVICE r1).  scripts/gen_integration.py rewrites
    the block."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_integration", os.path.join(ROOT, "scripts", "gen_integration.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    a = md.index("<!-- BEGIN sys-excerpt -->") + len("<!-- BEGIN sys-excerpt -->\n")
    b = md.index("\n<!-- END sys-excerpt -->")
    assert md[a:b] == gen.excerpt(), "INTEGRATION.md drifted from rust/azb200-sys: run python scripts/gen_integration.py"
    # and no other Rust struct declaration hides elsewhere in the document
    for name, body in re.findall(r"pub struct (azb_[a-z0-9_]+) \{(.*?)\n\}", md, flags=re.S):
        crate = re.search(r"pub struct %s \{(.*?)\n\}" % name, RS, flags=re.S)
        assert crate and crate.group(1) == body, name


def test_wrapper_implements_the_reference_traits():
    """rust/azb200 restates `trait Game` (src/game.rs:10-28) and `trait NNet` (src/nnet.rs:35-45) and implements them for
    ConnectFourGame / B200Net: every method of the two traits appears in the matching `impl ... for ...` block with the
    trait's own signature (text compare, whitespace-insensitive)."""
    src = open(os.path.join(ROOT, "rust", "azb200", "src", "lib.rs")).read()

    def block(header):
        i = src.index(header)
        depth, j = 0, src.index("{", i)
        for k in range(j, len(src)):
            depth += src[k] == "{"
            depth -= src[k] == "}"
            if depth == 0:
                return src[j + 1:k]
        raise AssertionError(header)

    def sigs(text):
        return {re.sub(r"\s+", " ", m).strip() for m in re.findall(r"fn [a-z_]+(?:<[^>]*>)?\([^)]*\)(?: -> [^;{]+)?", text)}

    for trait, impl in (("pub trait Game:", "impl Game for ConnectFourGame"), ("pub trait NNet", "impl NNet for B200Net")):
        want, have = sigs(block(trait)), sigs(block(impl))
        assert want and want == have, (trait, sorted(want ^ have))
    game = sigs(block("pub trait Game:"))
    assert len(game) == 9 and "fn get_next_state(&self, player: i8, action: u8) -> (Self, i8)" in game
    # the closure API of src/arena.rs:7-11,62-67
    assert re.search(r"pub fn play_game<G: Game>\(player_actions: &\[&dyn Fn\(&G\) -> u8\], board: &Option<G>, verbose: bool\) -> i8", src)
    assert re.search(r"pub fn play_games<G: Game>\(num: usize, player_actions: Vec<&dyn Fn\(&G\) -> u8>, board: Option<G>, verbose: bool\)", src)
