"""GPU numerics of the small fused decoder kernels (csrc/small_attention.cu, csrc/decoder_small.cu) against plain PyTorch
fp32 formulations of what they replace in the reference's decoder layer
(projects/SeqFormer/seqformer/models/deformable_transformer.py:280-317, 366-377)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


@pytest.mark.parametrize("B,n,heads", [(1, 300, 8), (5, 300, 8), (2, 37, 4), (3, 384, 8), (1, 1, 1), (2, 64, 2)])
def test_small_mha_matches_multihead_attention(B, n, heads):
    from vnext_b200 import fused
    C = heads * 32
    g = torch.Generator().manual_seed(B * 1000 + n)
    mha = torch.nn.MultiheadAttention(C, heads, dropout=0.0, batch_first=True).cuda().eval()
    x = torch.randn(B, n, C, generator=g).cuda()
    pos = torch.randn(1, n, C, generator=g).cuda()
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        want = mha(x + pos, x + pos, x, need_weights=False)[0]
        # the kernel's contract: everything after the in-projection and before the out-projection
        qk = F.linear(x + pos, mha.in_proj_weight[:2 * C], mha.in_proj_bias[:2 * C])
        v = F.linear(x, mha.in_proj_weight[2 * C:], mha.in_proj_bias[2 * C:])
        wide = torch.zeros(B, n, 3 * C + 64, device="cuda")            # strided rows, as out of a wider buffer
        wide[..., :2 * C] = qk
        wide[..., 2 * C:3 * C] = v
        o = fused.small_mha(wide[..., :3 * C], heads)
        got = F.linear(o, mha.out_proj.weight, mha.out_proj.bias)
    assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("T,nq,C", [(5, 300, 256), (1, 300, 256), (3, 17, 64), (64, 5, 256)])
def test_time_pool_ln(T, nq, C):
    from vnext_b200 import fused
    g = torch.Generator().manual_seed(T + nq)
    out = torch.randn(1, nq, C, generator=g).cuda()
    t2 = torch.randn(T, nq, C, generator=g).cuda()
    logit = (torch.randn(T, nq, 1, generator=g) * 3).cuda()
    gamma, beta = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
    want = F.layer_norm(out + (t2 * torch.softmax(logit, 0)).sum(0, keepdim=True), (C,), gamma, beta, 1e-5)
    got = fused.time_pool_ln(out, t2, logit, gamma, beta)
    assert torch.allclose(got, want, rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_box_refine(ref_dim):
    from vnext_b200 import fused
    from vnext_b200.seqformer.engine import inverse_sigmoid
    g = torch.Generator().manual_seed(ref_dim)
    delta = (torch.randn(5, 300, 4, generator=g) * 2).cuda()
    ref = torch.rand(5, 300, ref_dim, generator=g).cuda()
    ref[0, 0] = 0.0
    ref[0, 1] = 1.0                                   # the clamps of inverse_sigmoid
    ref[0, 2, 0] = 1.5
    if ref_dim == 4:
        want = (delta + inverse_sigmoid(ref)).sigmoid()
    else:
        want = torch.cat([delta[..., :2] + inverse_sigmoid(ref), delta[..., 2:]], -1).sigmoid()
    got = fused.box_refine(delta, ref.expand(5, 300, ref_dim))
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_ref_points_px(ref_dim):
    from vnext_b200 import fused
    from vnext_b200.seqformer.engine import inverse_sigmoid
    ref = torch.rand(5, 300, ref_dim, generator=torch.Generator().manual_seed(1)).cuda()
    ref[0, 0] = 0.0
    ref[0, 1] = 1.0
    want = inverse_sigmoid(ref)[..., :2].sigmoid() * torch.tensor([1280.0, 720.0], device="cuda")
    got = fused.ref_points_px(ref, 1280.0, 720.0)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-3)
