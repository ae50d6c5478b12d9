"""The product package must never import, call or execute anything under oracle/ (nor fall back to torch ops)."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "prodiff_b200")


def _py_files():
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(d, f)


def test_product_never_imports_oracle():
    for path in _py_files():
        tree = ast.parse(open(path).read(), path)
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for n in names:
                assert not n.split(".")[0] == "oracle", f"{path} imports {n}"
        assert "oracle" not in open(path).read().replace("# oracle", ""), f"{path} mentions the oracle"


def test_product_forward_has_no_torch_compute_fallback():
    """wavenet.py / diffusion.py may not call conv / linear / activation ops: all compute is in the CUDA library."""
    banned = ("F.conv1d", "F.linear", "torch.sigmoid", "torch.tanh", "F.relu", "softplus", "torch.matmul", "@ ",
              "self.input_projection(", "self.mlp(", "layer(x", "self.skip_projection(", "self.output_projection(")
    for name in ("wavenet.py", "engine.py", "encoder.py", "vocoder.py"):
        src = open(os.path.join(PKG, name)).read()
        for b in banned:
            assert b not in src, f"{name} contains {b}"


def test_csrc_has_no_library_gemm_or_triton():
    for d, _, files in os.walk(os.path.join(PKG, "csrc")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".py")):
                src = open(os.path.join(d, f)).read()
                for b in ("cublas", "cudnn", "cutlass", "cute::", "triton"):
                    assert b not in src.lower() or f == "pd_ptx.cuh" and b in ("cutlass", "cute::"), f"{f} mentions {b}"
