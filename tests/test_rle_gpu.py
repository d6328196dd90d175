"""GPU parity of the device RLE encoder (csrc/rle_encode.cu behind vnext_rle_encode) — bit-exact against the oracle
(oracle/rle_oracle.py, pinned by the reference's own RLE strings): golden strings, edge shapes, dense / noisy / degenerate masks,
buffer overflow reporting, the full BASELINE size (50 masks of 720x1280), and the SeqFormer facade's mask_format="rle"."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rle_oracle as R

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rle_golden.json")))


def _blobs(n, h, w, seed, smooth=9):
    """Smooth random fields thresholded at 0: blob-shaped masks with a few runs per column, like real segmentations."""
    g = torch.Generator().manual_seed(seed)
    f = torch.randn((n, 1, h // smooth + 2, w // smooth + 2), generator=g)
    f = torch.nn.functional.interpolate(f, size=(h, w), mode="bilinear", align_corners=False)[:, 0]
    return f > 0.3


def _check(masks):
    from vnext_b200 import rle
    got = rle.encode(masks.cuda())
    flat = masks.reshape(-1, *masks.shape[-2:]).numpy()
    assert len(got) == flat.shape[0]
    for g_, m in zip(got, flat):
        want = R.encode(m)
        assert g_["size"] == want["size"]
        assert g_["counts"] == want["counts"].decode("ascii")
    return got


def test_golden_masks_give_the_reference_strings():
    from vnext_b200 import rle
    for case in GOLDEN:
        h, w = case["size"]
        mask = torch.from_numpy(R.decode(R.from_string(case["counts"]), h, w)).bool()
        out = rle.encode(mask.cuda()[None])
        assert out == [{"size": [h, w], "counts": case["counts"]}]


@pytest.mark.parametrize("h,w", [(1, 1), (1, 300), (300, 1), (7, 9), (8, 128), (9, 129), (64, 127), (90, 160), (426, 640), (720, 1280)])
def test_shapes_against_the_oracle(h, w):
    g = torch.Generator().manual_seed(h * 1000 + w)
    masks = torch.stack([torch.rand((h, w), generator=g) < p for p in (0.5, 0.03, 0.97)] +
                        [torch.zeros(h, w, dtype=torch.bool), torch.ones(h, w, dtype=torch.bool)])
    if h >= 64 and w >= 64:
        masks = torch.cat((masks, _blobs(3, h, w, 5)))
    _check(masks)


def test_first_and_last_pixels_and_column_wraps():
    h, w = 33, 47
    cases = []
    for pix in [(0, 0), (h - 1, w - 1), (h - 1, 0), (0, 1), (h - 1, 5)]:
        m = torch.zeros(h, w, dtype=torch.bool)
        m[pix] = True
        cases += [m, ~m]
    col = torch.zeros(h, w, dtype=torch.bool)
    col[:, 3] = True                         # a run that starts and ends exactly at column boundaries
    col[h - 1, 7] = col[0, 8] = True         # a run that crosses a column boundary
    cases += [col, ~col]
    _check(torch.stack(cases))


def test_uint8_masks_with_values_other_than_one():
    from vnext_b200 import rle
    g = torch.Generator().manual_seed(3)
    m = (torch.rand((2, 40, 50), generator=g) < 0.4)
    u8 = m.to(torch.uint8) * torch.randint(1, 255, m.shape, generator=g, dtype=torch.uint8)
    assert rle.encode(u8.cuda()) == rle.encode(m.cuda())
    # a base address that is not 4-byte aligned (W % 4 == 0 otherwise takes the 4-columns-per-thread path)
    m4 = torch.rand((3, 40, 48), generator=g) < 0.4
    buf = torch.zeros(m4.numel() + 8, dtype=torch.uint8, device="cuda")
    for shift in (0, 1, 2, 3):
        view = buf[shift:shift + m4.numel()].view(3, 40, 48)
        view.copy_(m4.to(torch.uint8))
        got = rle.encode(view)
        assert [g_["counts"].encode() for g_ in got] == [R.encode(x.numpy())["counts"] for x in m4]


def test_checkerboard_overflows_the_default_buffer_and_reports_the_needed_size():
    from vnext_b200.rle import RleEncoder, split_strings
    h, w = 64, 96
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    masks = torch.stack([(yy + xx) % 2 == 0, (yy + xx) % 2 == 1]).cuda()
    enc = RleEncoder(2, h, w, "cuda", capacity=100)
    offs, chars = enc.encode(masks)
    guard = chars.clone()
    total = int(offs[-1])
    want = [R.encode(m.cpu().numpy())["counts"] for m in masks]
    assert total == sum(len(s) for s in want) > 100
    assert bytes(guard.cpu().numpy()) == (want[0] + want[1])[:100]            # the part that fits is right, nothing written past it
    enc.grow(total)
    offs, chars = enc.encode(masks)
    got = split_strings(offs.cpu(), chars.cpu(), h, w)
    assert [g_["counts"].encode() for g_ in got] == want


def test_argument_errors_and_empty_batch():
    from vnext_b200 import _cabi
    lib = _cabi.lib()
    offs = torch.full((1,), 7, dtype=torch.int64, device="cuda")
    assert lib.vnext_rle_encode(None, 0, 8, 8, None, 0, offs.data_ptr(), None, 0, None) == 0
    torch.cuda.synchronize()
    assert int(offs[0]) == 0
    assert lib.vnext_rle_encode(None, 1, 8, 8, None, 0, offs.data_ptr(), None, 0, None) != 0
    assert b"vnext_rle_encode" in lib.vnext_last_error()
    assert lib.vnext_rle_encode(None, 1, 0, 8, None, 0, offs.data_ptr(), None, 0, None) != 0
    with pytest.raises(RuntimeError):
        from vnext_b200.rle import RleEncoder
        RleEncoder(1, 8, 8, "cuda").encode(torch.zeros(8, 8, dtype=torch.bool))          # host tensor: no CPU fallback


def test_full_size_clip_50_masks_720p_round_trip_and_oracle():
    """BASELINE geometry: top-10 instances x 5 frames of 720x1280.  Every string decodes back to its mask (size-independent
    property), the first masks also equal the oracle's strings, and the encoder works inside a CUDA graph."""
    from vnext_b200.rle import RleEncoder, split_strings
    masks = _blobs(50, 720, 1280, 11, smooth=40).view(10, 5, 720, 1280)
    masks[3, 2] = False
    masks[4, 0] = True
    dm = masks.cuda()
    enc = RleEncoder(50, 720, 1280, "cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        enc.encode(dm)
        s.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            enc.encode(dm)
        enc.chars.zero_()
        graph.replay()
    s.synchronize()
    total = int(enc.char_offsets[-1])
    assert 0 < total < enc.capacity
    got = split_strings(enc.char_offsets.cpu(), enc.chars.cpu(), 720, 1280)
    flat = masks.view(50, 720, 1280).numpy()
    for i, (g_, m) in enumerate(zip(got, flat)):
        cnts = R.from_string(g_["counts"])
        assert (R.decode(cnts, 720, 1280) == m).all(), i
        if i < 6 or i in (17, 20):
            assert g_["counts"] == R.encode(m)["counts"].decode("ascii")


def test_seqformer_facade_rle_output_equals_its_boolean_masks():
    from vnext_b200.seqformer.meta_arch import SeqFormer
    from vnext_b200.seqformer.ytvis_eval import instances_to_coco_json_video
    from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
    from vnext_b200.synth import fill_state_dict, synthetic_clip
    sd = fill_state_dict(seqformer_r50_state_spec(), 0)
    frames = synthetic_clip(3, 96, 128, 0)
    inputs = [{"image": frames, "height": 120, "width": 160, "video_id": 7, "length": 3}]
    model = SeqFormer(state_dict=sd, precision="fp32").eval()
    ref = model(inputs)
    model_rle = SeqFormer(state_dict=sd, precision="fp32", mask_format="rle").eval()
    out = model_rle(inputs)
    assert out["pred_scores"] == ref["pred_scores"] and out["pred_labels"] == ref["pred_labels"]
    assert len(out["pred_masks"]) == len(ref["pred_masks"]) > 0
    for rles, m in zip(out["pred_masks"], ref["pred_masks"]):
        assert len(rles) == m.shape[0]
        for r, frame in zip(rles, m):
            assert r["size"] == [120, 160]
            assert (R.decode(R.from_string(r["counts"]), 120, 160) == frame.numpy()).all()
    # the evaluator's record: identical whether the masks arrive encoded or as the reference's boolean tensors
    a, b = instances_to_coco_json_video(inputs, out), instances_to_coco_json_video(inputs, ref)
    assert a == b and a[0]["video_id"] == 7 and len(a[0]["segmentations"]) == 3
    # stream(): same output, runners keyed by the mask format
    outs = list(model_rle.stream([inputs, inputs]))
    assert outs[0] == out and outs[1] == out
