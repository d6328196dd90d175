"""MSDeformAttn nn.Module drop-ins (CUDA) vs outputs of the reference's own modules (tests/golden/module_golden.npz)."""
import os

import numpy as np
import pytest
import torch

from vnext_b200.synth import fill_state_dict

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(12, 16), (6, 8), (3, 4), (2, 2)]


def inputs(seed, N, nf, Lq):
    g = torch.Generator().manual_seed(seed)
    shapes = torch.as_tensor(SHAPES, dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    return dict(shapes=shapes.cuda(), lsi=lsi.cuda(), S=S,
                src=torch.randn(N, nf, S, 256, generator=g).cuda(), query=torch.randn(N, nf, Lq, 256, generator=g).cuda(),
                ref2=torch.rand(N, nf, Lq, 4, 2, generator=g).cuda(),
                ref4=(torch.rand(N, nf, Lq, 4, 4, generator=g) * 0.5 + 0.2).cuda(),
                mask=(torch.rand(N, nf, S, generator=g) < 0.1).cuda())


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(ROOT, "tests", "golden", "module_golden.npz"))


def _close(a, key, G, tol=1e-4):
    b = torch.from_numpy(G[key])
    assert a.shape == b.shape, (a.shape, b.shape)
    assert (a.detach().cpu() - b).abs().max().item() < tol * max(1.0, b.abs().max().item()), key


def _load(m, seed):
    m.load_state_dict(fill_state_dict({k: v.shape for k, v in m.state_dict().items()}, seed))
    return m.cuda().eval()


def test_idol_module(G):
    from vnext_b200.modules import MSDeformAttn
    torch.backends.cuda.matmul.allow_tf32 = False
    m = _load(MSDeformAttn(256, 4, 8, 4), 11)
    d = inputs(1, 2, 1, 50)
    with torch.no_grad():
        for tag, ref in (("r2", d["ref2"][:, 0]), ("r4", d["ref4"][:, 0])):
            o, loc, aw = m(d["query"][:, 0], ref, d["src"][:, 0], d["shapes"], d["lsi"], d["mask"][:, 0])
            _close(o, f"idol_{tag}_out", G); _close(loc, f"idol_{tag}_loc", G); _close(aw, f"idol_{tag}_attn", G)
    # trains like the reference: gradients flow to every parameter through the CUDA backward
    m.train()
    q = d["query"][:, 0].clone().requires_grad_(True)
    o, _, _ = m(q, d["ref2"][:, 0], d["src"][:, 0], d["shapes"], d["lsi"], d["mask"][:, 0])
    o.square().mean().backward()
    assert q.grad is not None and all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    assert sorted(m.state_dict()) == sorted(["sampling_offsets.weight", "sampling_offsets.bias", "attention_weights.weight",
                                             "attention_weights.bias", "value_proj.weight", "value_proj.bias",
                                             "output_proj.weight", "output_proj.bias"])


def test_seqformer_modules(G):
    from vnext_b200 import _cabi
    from vnext_b200.modules import MSDeformAttnSeqFormer
    torch.backends.cuda.matmul.allow_tf32 = False
    d = inputs(2, 1, 3, 40)
    me = _load(MSDeformAttnSeqFormer(256, 4, 8, 4, "encode"), 12)
    qe = torch.randn(1, 3, d["S"], 256, generator=torch.Generator().manual_seed(5)).cuda()
    refe = torch.rand(1, d["S"], 4, 2, generator=torch.Generator().manual_seed(6)).cuda()
    with torch.no_grad():
        o = me(qe, None, refe, d["src"], d["shapes"], d["lsi"], d["mask"])
        # no-grad path = one fused kernel for all frames (self-attention shape: the cache-blocked kernel by default)
        assert _cabi.last_variant() in ("fwd_cached_fused_f32", "fwd_fused_f32_d32_tma")
        _close(o, "seq_enc_out", G)
    o2 = me(qe, None, refe, d["src"], d["shapes"], d["lsi"], d["mask"])    # autograd path, same numbers
    assert (o2 - o).abs().max().item() < 1e-4
    md = _load(MSDeformAttnSeqFormer(256, 4, 8, 4, "decode"), 13)
    with torch.no_grad():
        o, ob, loc, aw = md(d["query"][:, 0], d["query"][:, 0], d["ref2"], d["src"], d["shapes"], d["lsi"], d["mask"])
        _close(o, "seq_dec0_out", G); _close(ob, "seq_dec0_box", G); _close(loc, "seq_dec0_loc", G); _close(aw, "seq_dec0_attn", G)
        o, ob, loc, aw = md(d["query"][:, 0], d["query"], d["ref4"], d["src"], d["shapes"], d["lsi"], d["mask"])
        _close(o, "seq_dec1_out", G); _close(ob, "seq_dec1_box", G); _close(loc, "seq_dec1_loc", G); _close(aw, "seq_dec1_attn", G)
    assert "output_proj_box.weight" in md.state_dict()
