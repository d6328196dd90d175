"""GPU numerics of the single-pass NHWC kernels vs plain PyTorch fp32 references of the same ops."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


def test_bias_act():
    from vnext_b200 import fused
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 17, 23, 64, generator=g).cuda()
    b = torch.randn(64, generator=g).cuda()
    r = torch.randn(3, 17, 23, 64, generator=g).cuda()
    assert torch.equal(fused.bias_act(x, b, relu=True), F.relu(x + b))
    assert torch.equal(fused.bias_act(x, b, relu=False, residual=r), x + b + r)
    y = x.clone()
    fused.bias_act(y, b, relu=True, residual=r, out=y)
    assert torch.equal(y, F.relu(x + b + r))


@pytest.mark.parametrize("H,W", [(368, 640), (7, 9), (48, 64)])
def test_stem_pool(H, W):
    from vnext_b200 import fused
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, H, W, 64, generator=g).cuda()
    b = torch.randn(64, generator=g).cuda()
    want = F.max_pool2d(F.relu(x.permute(0, 3, 1, 2) + b.view(1, -1, 1, 1)), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(fused.stem_pool(x, b), want.contiguous())


@pytest.mark.parametrize("N,HW", [(5, 14720), (2, 12), (3, 240)])
def test_groupnorm_nhwc_into_strided_output(N, HW):
    from vnext_b200 import fused
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(N, HW, 256, generator=g) * 3 + 1).cuda()
    gamma, beta = (torch.rand(256, generator=g) + 0.5).cuda(), torch.randn(256, generator=g).cuda()
    want = F.group_norm(x.double().permute(0, 2, 1), 32, gamma.double(), beta.double(), 1e-5).permute(0, 2, 1)
    big = torch.zeros(N, HW + 100, 256, device="cuda")
    fused.groupnorm_nhwc(x, gamma, beta, 32, 1e-5, out=big[:, 50:50 + HW, :])
    assert (big[:, 50:50 + HW].double() - want).abs().max().item() < 1e-4
    assert (big[:, :50] == 0).all() and (big[:, 50 + HW:] == 0).all()


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
def test_preprocess(dtype):
    from vnext_b200 import fused
    g = torch.Generator().manual_seed(3)
    fr = torch.randint(0, 256, (3, 3, 80, 112), generator=g).to(dtype).cuda()
    mean, std = (123.675, 116.280, 103.530), (58.395, 57.120, 57.375)
    out = fused.preprocess(fr, mean, std, 96, 128)
    m = torch.tensor(mean, device="cuda").view(1, 3, 1, 1)
    s = torch.tensor(std, device="cuda").view(1, 3, 1, 1)
    want = F.pad((fr.float() - m) / s, (0, 16, 0, 16)).permute(0, 2, 3, 1)
    assert torch.allclose(out, want, rtol=1e-6, atol=1e-6)
    assert (out[:, 80:] == 0).all() and (out[:, :, 112:] == 0).all()


@pytest.mark.parametrize("per_frame", [False, True])
def test_dynamic_mask_head_kernel(per_frame):
    from oracle import seqformer_oracle as SO
    from vnext_b200 import fused
    g = torch.Generator().manual_seed(7)
    T, n, H, W = 3, 13, 23, 40
    feats = torch.randn(T, H, W, 8, generator=g)
    params = torch.randn((T, n, 169) if per_frame else (n, 169), generator=g) * 0.3
    ref = torch.rand(T, n, 2, generator=g) * torch.tensor([W * 8.0, H * 8.0])
    got = fused.dynamic_mask_head(feats.cuda(), params.cuda(), ref.cuda(), 8).cpu()          # [n, T, H, W]
    for t in range(T):
        prm = params[t] if per_frame else params
        want = SO.dynamic_mask_head(feats[t:t + 1].permute(0, 3, 1, 2), ref[t:t + 1], prm[None], out_stride=8)  # no upsample
        assert torch.allclose(got[:, t], want[0], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("H,W,OH,OW", [(720, 1280, 720, 1280), (80, 112, 80, 112), (96, 128, 60, 80), (90, 120, 180, 250)])
def test_mask_finalize_kernel(H, W, OH, OW):
    from oracle import seqformer_oracle as SO
    from vnext_b200 import fused
    g = torch.Generator().manual_seed(H + OW)
    Hp, Wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    l8 = torch.randn(4, 2, Hp // 8, Wp // 8, generator=g) * 3
    got = fused.mask_finalize(l8.cuda(), (H, W), (OH, OW), 4).cpu()
    m4 = SO.aligned_bilinear(l8.reshape(8, 1, Hp // 8, Wp // 8), 2).view(4, 2, Hp // 4, Wp // 4)
    pm = F.interpolate(m4, size=(Hp, Wp), mode="bilinear", align_corners=False).sigmoid()[:, :, :H, :W]
    want = F.interpolate(pm, size=(OH, OW), mode="nearest") > 0.5
    assert got.dtype == torch.bool and got.shape == want.shape
    assert (got == want).float().mean().item() > 0.9995


@pytest.mark.parametrize("T,H,W,u8", [(2, 96, 128, True), (1, 70, 100, False), (5, 480, 864, True), (1, 33, 260, True)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_stem_conv7x7_on_tensor_cores(T, H, W, u8, out_dtype):
    """Normalise + pad + 7x7/s2/p3 stem as one implicit GEMM (overlapping TMA view of a zero-bordered 4-channel image) vs
    F.conv2d of the reference preprocessing (S/seqformer.py:413-429 + S/util/misc.py:288-316; resnet.py:355-359)."""
    import torch.nn.functional as F
    from vnext_b200 import fused
    g = torch.Generator().manual_seed(T * 7 + H)
    fr = torch.randint(0, 256, (T, 3, H, W), generator=g)
    fr = fr.to(torch.uint8) if u8 else fr.float() + 0.25
    w = (torch.randn(64, 3, 7, 7, generator=g) / 12).cuda()
    mean, std = (123.675, 116.28, 103.53), (58.395, 57.12, 57.375)
    Hp, Wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    x = (fr.double().cuda() - torch.tensor(mean, dtype=torch.double, device="cuda").view(1, 3, 1, 1)) / \
        torch.tensor(std, dtype=torch.double, device="cuda").view(1, 3, 1, 1)
    x = F.pad(x, (0, Wp - W, 0, Hp - H))
    want = F.conv2d(x, w.double(), None, stride=2, padding=3).permute(0, 2, 3, 1)
    got = fused.stem_conv7x7(fr.cuda(), fused.pack_stem_weight(w), mean, std, Hp, Wp, out_dtype=out_dtype)
    assert got.shape == want.shape and got.dtype == out_dtype
    tol = (2e-3 if out_dtype == torch.float32 else 8e-3) * want.abs().max().item()
    assert (got.double() - want).abs().max().item() < tol
