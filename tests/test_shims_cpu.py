"""The Rust shim crates (shims/) cannot be compiled here (no rustc / cargo in this image), so what CAN be checked without
a compiler is checked: every file has balanced delimiters outside strings / comments / char literals, every `sys::omk_*`
call the shims make is declared in the sys crate's `extern "C"` block with the number of arguments the call passes, and
the public items the reference's callers use are all present with the reference's names."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "shims")


def strip_rust(text):
    """Blank out comments, string literals and char literals (keeps offsets irrelevant, structure intact)."""
    out, i, n = [], 0, len(text)
    while i < n:
        c = text[i]
        if text.startswith("//", i):
            j = text.find("\n", i)
            i = n if j < 0 else j
        elif text.startswith("/*", i):
            j = text.find("*/", i + 2)
            i = n if j < 0 else j + 2
        elif c == '"':
            j = i + 1
            while j < n and text[j] != '"':
                j += 2 if text[j] == "\\" else 1
            out.append('""')
            i = j + 1
        elif c == "'" and re.match(r"'(\\.|[^\\'])'", text[i:i + 4]):
            m = re.match(r"'(\\.|[^\\'])'", text[i:i + 4])
            out.append("' '")
            i += m.end()
        else:
            out.append(c)
            i += 1
    return "".join(out)


def rust_files():
    for d, _, files in os.walk(SHIMS):
        for f in files:
            if f.endswith(".rs"):
                yield os.path.join(d, f)


def split_args(s):
    depth, cur, out = 0, "", []
    for ch in s:
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def test_delimiters_balance_in_every_shim_file():
    pairs = {")": "(", "]": "[", "}": "{"}
    for path in rust_files():
        stack = []
        for ch in strip_rust(open(path).read()):
            if ch in "([{":
                stack.append(ch)
            elif ch in pairs:
                assert stack and stack[-1] == pairs[ch], f"{path}: unbalanced {ch}"
                stack.pop()
        assert not stack, f"{path}: unclosed {stack[-1]}"


def test_every_sys_call_is_declared_with_the_same_number_of_arguments():
    sys_src = strip_rust(open(os.path.join(SHIMS, "omok-b200-sys", "src", "lib.rs")).read())
    declared = {m.group(1): len(split_args(m.group(2))) for m in re.finditer(r"pub fn (omk_\w+)\s*\(([^;]*?)\)\s*(?:->[^;]*)?;", sys_src, re.S)}
    assert len(declared) >= 40
    calls = 0
    for crate in ("environment", "alpha-zero"):
        src = strip_rust(open(os.path.join(SHIMS, crate, "src", "lib.rs")).read())
        for m in re.finditer(r"sys::(omk_\w+)\s*\(", src):
            name, i, depth = m.group(1), m.end(), 1
            j = i
            while depth:
                depth += {"(": 1, ")": -1}.get(src[j], 0)
                j += 1
            args = split_args(src[i:j - 1])
            assert name in declared, f"{crate}: sys::{name} is not declared in omok-b200-sys"
            assert len(args) == declared[name], f"{crate}: sys::{name} called with {len(args)} arguments, declared with {declared[name]}"
            calls += 1
    assert calls >= 15


def test_the_reference_callers_surface_is_present():
    """the items src/trainer.rs, benchmark/ and gui/ import from the two crates."""
    env = open(os.path.join(SHIMS, "environment", "src", "lib.rs")).read()
    az = open(os.path.join(SHIMS, "alpha-zero", "src", "lib.rs")).read()
    for item in ("pub enum Stone", "pub enum Turn", "pub enum GameStatus", "pub struct Environment", "pub fn new(", "pub fn place_stone(",
                 "pub fn encode_board", "pub fn opponent(", "pub fn is_terminal("):
        assert item in env, item
    for item in ("pub struct Agent", "pub enum ActionSamplingMode", "pub struct AgentModel", "pub struct ModelIO", "pub enum EnvTurnMode",
                 "pub fn encode_nn_input", "pub fn encode_nn_targets", "pub struct MCTSExecutor", "pub struct ParallelMCTSExecutor",
                 "pub fn sample_action(", "pub fn play_action(", "pub fn ensure_action_exists(", "pub fn compute_policy(", "pub fn evaluate_p(",
                 "pub fn evaluate_pv(", "pub fn train(", "pub fn execute(", "pub fn run(", "pub fn save(", "pub fn load("):
        assert item in az, item
    assert "omk_train_step" in az and "Err(Status::from_message(\"AgentModel::train is not part" not in az
