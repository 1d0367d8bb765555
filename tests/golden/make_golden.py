"""Generates tests/golden/*.pt by running the REFERENCE'S OWN modules (imported from
/root/reference in the build container; the GPU box has no /root/reference, which is why the
outputs are committed).  Run:  python tests/golden/make_golden.py

Reference modules exercised (python/sglang/srt/…):
  layers/attention/torch_native_backend.py  TorchNativeAttnBackend.forward_extend / forward_decode
  layers/radix_attention.py                 RadixAttention
  layers/rotary_embedding.py                RotaryEmbedding.forward_native, Llama3RotaryEmbedding
  model_executor/forward_batch_info.py      compute_position_torch, ForwardMode
They need three import shims because of version drift in this image: an `IPython`
stub, a `semi_pd_ipc` stub and three `triton.runtime.cache` names.
"""
import os
import sys
import types

import torch

REF = "/root/reference/python"
OUT = os.path.dirname(os.path.abspath(__file__))


def _shim():
    sys.path.insert(0, REF)
    ip = types.ModuleType("IPython")
    ip.display = types.ModuleType("IPython.display")
    ip.display.HTML = ip.display.display = lambda *a, **k: None
    sys.modules["IPython"] = ip
    sys.modules["IPython.display"] = ip.display
    sys.modules["semi_pd_ipc"] = types.ModuleType("semi_pd_ipc")
    import triton.runtime.cache as trc
    for n in ("default_cache_dir", "default_dump_dir", "default_override_dir"):
        if not hasattr(trc, n):
            setattr(trc, n, lambda: "/tmp/triton")


class _Pool:
    """Duck-typed KV pool (the reference's MHATokenToKVPool cannot be constructed without CUDA,
    memory_pool.py:229)."""

    def __init__(self, slots, H, D, dtype):
        self.k = torch.zeros(slots, H, D, dtype=dtype)
        self.v = torch.zeros(slots, H, D, dtype=dtype)

    def set_kv_buffer(self, layer, loc, k, v):
        self.k[loc] = k
        self.v[loc] = v

    def get_key_buffer(self, layer_id):
        return self.k

    def get_value_buffer(self, layer_id):
        return self.v


def main():
    _shim()
    from sglang.srt.layers.attention.torch_native_backend import TorchNativeAttnBackend
    from sglang.srt.layers.radix_attention import RadixAttention
    from sglang.srt.layers.rotary_embedding import Llama3RotaryEmbedding, RotaryEmbedding
    from sglang.srt.model_executor.forward_batch_info import ForwardMode, compute_position_torch

    torch.manual_seed(20261017)
    golden = {}

    # ---- positions -------------------------------------------------------------------------
    cases = []
    for pre, ext in (([0], [1]), ([0, 5, 100, 0], [7, 1, 300, 64]), ([3] * 37, list(range(1, 38)))):
        p, s = compute_position_torch(torch.tensor(pre, dtype=torch.int32), torch.tensor(ext, dtype=torch.int32))
        cases.append(dict(prefix=pre, extend=ext, positions=p, start=s))
    golden["positions"] = cases

    # ---- RoPE (forward_native, fp32 inputs so that the cache dtype cast is a no-op) -----------
    rope = []
    for (D, rot, maxpos, base, neox, T, Hq, Hk) in ((128, 128, 2048, 10000, True, 33, 8, 2),
                                                    (128, 128, 4096, 500000, True, 17, 32, 8),
                                                    (64, 64, 1024, 10000, False, 9, 4, 4),
                                                    (128, 64, 2048, 10000, True, 5, 4, 1)):
        r = RotaryEmbedding(D, rot, maxpos, base, neox, torch.float32)
        pos = torch.randint(0, maxpos, (T,))
        q, k = torch.randn(T, Hq * D), torch.randn(T, Hk * D)
        qo, ko = r.forward_native(pos, q.clone(), k.clone())
        rope.append(dict(D=D, rot=rot, maxpos=maxpos, base=base, neox=neox, pos=pos, q=q, k=k,
                         q_out=qo, k_out=ko, cache_rows=r.cos_sin_cache[pos]))
    sc = dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
              original_max_position_embeddings=8192)
    r3 = Llama3RotaryEmbedding(128, 128, 16384, 500000, True, torch.float32, sc["factor"],
                               sc["low_freq_factor"], sc["high_freq_factor"],
                               sc["original_max_position_embeddings"])
    golden["rope"] = rope
    golden["llama3_cache"] = dict(scaling=sc, rows=torch.tensor([0, 1, 17, 4095, 8191, 16383]),
                                  values=r3.cos_sin_cache[torch.tensor([0, 1, 17, 4095, 8191, 16383])])

    # ---- attention through RadixAttention -> TorchNativeAttnBackend ---------------------------
    class _MR:
        device = "cpu"

    class _FB:
        pass

    attn = []
    for dtype in (torch.float32, torch.bfloat16):
        for (Hq, Hk, D, pres, exts) in ((8, 2, 64, [0], [5]), (8, 2, 64, [0, 0, 0], [1, 17, 24]),
                                        (32, 8, 128, [7, 0, 30], [9, 2, 16]), (4, 4, 128, [64], [1])):
            bs = len(pres)
            seqs = [a + b for a, b in zip(pres, exts)]
            slots = sum(seqs) + 9
            pool = _Pool(slots, Hk, D, dtype)
            pool.k[:] = torch.randn(slots, Hk, D).to(dtype)
            pool.v[:] = torch.randn(slots, Hk, D).to(dtype)
            rtt = torch.zeros(bs + 2, max(seqs) + 4, dtype=torch.int32)
            rpi = torch.randperm(bs + 2)[:bs]
            perm = torch.randperm(slots - 1)[: sum(seqs)] + 1
            o = 0
            for i in range(bs):
                rtt[rpi[i], : seqs[i]] = perm[o:o + seqs[i]].int()
                o += seqs[i]
            T = sum(exts)
            q = torch.randn(T, Hq * D).to(dtype)
            k = torch.randn(T, Hk * D).to(dtype)
            v = torch.randn(T, Hk * D).to(dtype)
            loc = torch.cat([rtt[rpi[i], pres[i]:seqs[i]] for i in range(bs)]).long()
            layer = RadixAttention(Hq, D, D ** -0.5, Hk, layer_id=0)
            fb = _FB()
            fb.forward_mode = ForwardMode.EXTEND
            fb.token_to_kv_pool = pool

            class _R2T:
                req_to_token = rtt
            fb.req_to_token_pool = _R2T()
            fb.req_pool_indices = rpi
            fb.seq_lens = torch.tensor(seqs)
            fb.extend_prefix_lens = torch.tensor(pres)
            fb.extend_seq_lens = torch.tensor(exts)
            fb.out_cache_loc = loc
            fb.attn_backend = TorchNativeAttnBackend(_MR())
            k_pool0, v_pool0 = pool.k.clone(), pool.v.clone()
            out_ext = layer(q.clone(), k.clone(), v.clone(), fb)
            # decode step on top: one more token per request
            loc_d = (torch.arange(bs) + slots - bs - 0).long()  # unused high slots
            loc_d = torch.tensor([s for s in range(1, slots) if s not in set(perm.tolist())][:bs]).long()
            for i in range(bs):
                rtt[rpi[i], seqs[i]] = int(loc_d[i])
            qd = torch.randn(bs, Hq * D).to(dtype)
            kd = torch.randn(bs, Hk * D).to(dtype)
            vd = torch.randn(bs, Hk * D).to(dtype)
            fb.forward_mode = ForwardMode.DECODE
            fb.seq_lens = torch.tensor([s + 1 for s in seqs])
            fb.out_cache_loc = loc_d
            out_dec = layer(qd.clone(), kd.clone(), vd.clone(), fb)
            attn.append(dict(dtype=str(dtype), Hq=Hq, Hk=Hk, D=D, prefix=pres, extend=exts,
                             req_to_token=rtt.clone(), req_pool_indices=rpi, k_pool0=k_pool0,
                             v_pool0=v_pool0, q=q, k=k, v=v, out_cache_loc=loc, out_extend=out_ext,
                             qd=qd, kd=kd, vd=vd, out_cache_loc_d=loc_d, out_decode=out_dec))
    golden["attention"] = attn
    torch.save(golden, os.path.join(OUT, "reference_ops.pt"))
    print("wrote", os.path.join(OUT, "reference_ops.pt"),
          os.path.getsize(os.path.join(OUT, "reference_ops.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
