"""The oracle is test infrastructure: the product package must never import, call or execute it,
and must have no torch / CPU fallback for its ops."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "semi_pd_b200")


def _sources():
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                yield os.path.join(d, f)


def test_product_never_touches_oracle():
    for path in _sources():
        text = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
        assert "llama_oracle" not in text, path


def _code_only(path):
    """Python source with string literals (docstrings) and comments removed."""
    import io
    import tokenize
    out = []
    for tok in tokenize.generate_tokens(io.StringIO(open(path).read()).readline):
        if tok.type not in (tokenize.STRING, tokenize.COMMENT):
            out.append(tok.string)
    return " ".join(out)


def test_ops_have_no_eager_fallback():
    text = _code_only(os.path.join(PKG, "ops.py"))
    for banned in ("F . linear", "torch . matmul", "scaled_dot_product_attention", "F . silu", "functional"):
        assert banned not in text, banned
    layers = _code_only(os.path.join(PKG, "layers.py"))
    assert "forward_native" not in layers


def test_no_triton_or_compile_in_product():
    for path in _sources():
        text = _code_only(path) if path.endswith(".py") else open(path).read()
        assert "import triton" not in text and "torch . compile" not in text, path
