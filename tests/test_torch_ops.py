"""`torch.ops` registration over the C-ABI (semi_pd_b200/torch_ops.py): schemas equal to the
reference's `sgl-kernel/csrc/torch_extension.cc:49-73`, a second namespace can be added so that the
reference's own wrappers (`sgl_kernel/elementwise.py`) bind with zero edits, no CPU kernel."""
import pytest
import torch

from semi_pd_b200 import torch_ops


def test_ops_are_registered_with_the_reference_schemas():
    for name in ("rmsnorm", "fused_add_rmsnorm", "silu_and_mul", "apply_rope_pos_ids_cos_sin_cache"):
        op = getattr(torch.ops.spd, name)
        schema = str(op.default._schema)
        assert schema.startswith(f"spd::{name}(")
    s = str(torch.ops.spd.rmsnorm.default._schema)
    assert "! -> ) output" in s and "float eps" in s and "int cuda_stream" in s  # output is written
    s = str(torch.ops.spd.fused_add_rmsnorm.default._schema)
    assert "! -> ) input" in s and "! -> ) residual" in s
    s = str(torch.ops.spd.apply_rope_pos_ids_cos_sin_cache.default._schema)
    assert "bool interleave" in s and "! -> ) q_rope" in s and "Tensor q," in s


def test_second_namespace_and_no_cpu_kernel():
    torch_ops.register("sgl_kernel_spd_test")
    torch_ops.register("sgl_kernel_spd_test")  # idempotent
    assert hasattr(torch.ops.sgl_kernel_spd_test, "silu_and_mul")
    x = torch.zeros(2, 256, dtype=torch.bfloat16)
    out = torch.zeros(2, 128, dtype=torch.bfloat16)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.spd.silu_and_mul(out, x, 0)  # host tensors: no CPU path


@pytest.mark.gpu
def test_registered_ops_match_eager_math():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    x = torch.randn(37, 512, device=dev).bfloat16()
    w = (torch.randn(512, device=dev) * 0.1 + 1).bfloat16()
    out = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    torch.ops.spd.rmsnorm(out, x, w, 1e-5, st)
    ref = (x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-5)).bfloat16().float() * w.float()
    assert (out.float() - ref).abs().max().item() < 3e-2
    res = torch.randn_like(x)
    x2, r2 = x.clone(), res.clone()
    torch.ops.spd.fused_add_rmsnorm(x2, r2, w, 1e-5)
    assert torch.equal(r2, (x.float() + res.float()).bfloat16())
    g = torch.randn(9, 1024, device=dev).bfloat16()
    o = torch.empty(9, 512, device=dev, dtype=torch.bfloat16)
    torch.ops.spd.silu_and_mul(o, g, st)
    ref = torch.nn.functional.silu(g[:, :512].float()) * g[:, 512:].float()
    assert (o.float() - ref).abs().max().item() < 3e-2
    # rope through the wrapper shape the reference uses (views of one storage, in place)
    T, H, D = 11, 4, 128
    q = torch.randn(T, H * D, device=dev).bfloat16()
    k = torch.randn(T, 2 * D, device=dev).bfloat16()
    pos = torch.arange(T, device=dev)
    inv = 1.0 / (10000 ** (torch.arange(0, D, 2, device=dev).float() / D))
    fr = torch.outer(torch.arange(64, device=dev).float(), inv)
    cache = torch.cat([fr.cos(), fr.sin()], -1)
    q0, k0 = q.clone(), k.clone()
    torch.ops.spd.apply_rope_pos_ids_cos_sin_cache(q.view(T, -1, D), k.view(T, -1, D), q.view(T, -1, D),
                                                  k.view(T, -1, D), cache, pos, False, st)

    def neox(x):
        x = x.float().view(T, -1, D)
        c, s = cache[pos][:, None, : D // 2], cache[pos][:, None, D // 2:]
        x1, x2_ = x[..., : D // 2], x[..., D // 2:]
        return torch.cat([x1 * c - x2_ * s, x2_ * c + x1 * s], -1).view(T, -1)

    assert (q.float() - neox(q0)).abs().max().item() < 3e-2
    assert (k.float() - neox(k0)).abs().max().item() < 3e-2
