"""SeqFormer training on the B200: dense forward/backward with the sm_100a MSDA kernels against the reference's own
losses / assignments / gradients, the facade's training forward, and one trainer step."""
import numpy as np
import pytest
import torch

from test_seqformer_train_host import B, GOLD, H, NF, SEED, W, fake_batched_inputs, run
from vnext_b200 import _cabi
from vnext_b200.seqformer import train as ST
from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
from vnext_b200.synth import fill_state_dict, synthetic_clip, synthetic_clip_targets

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.fixture()
def strict_fp32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_train_step_matches_reference(strict_fp32):
    n0 = _cabi.lib().vnext_launch_count()
    worst, _ = run("cuda", 2e-3, 1e-2, 1e-2)
    assert _cabi.lib().vnext_launch_count() - n0 >= 24
    print("worst grad-norm rel err vs reference:", worst)


def test_facade_training_forward_and_trainer_step(strict_fp32):
    from vnext_b200.seqformer.meta_arch import SeqFormer
    gold = dict(np.load(GOLD))
    batched, _ = fake_batched_inputs()
    m = SeqFormer(state_dict=fill_state_dict(seqformer_r50_state_spec(), SEED), device="cuda", precision="fp32", dropout=0.0)
    m.train()
    losses = m(batched)
    assert len(losses) == 5 * 6
    np.testing.assert_allclose(float(sum(losses.values()).detach()), float(gold["total"]), rtol=2e-3)
    model = m.train_model(NF)
    trainer = ST.SeqFormerTrainer(model)
    frames = torch.stack(synthetic_clip(B * NF, H, W, SEED)).cuda()
    targets = [{k: v.cuda() for k, v in t.items()} for t in synthetic_clip_targets(B, NF, H, W, SEED)]
    w0 = model.flat.w.clone()
    for _ in range(2):
        _, total, _ = trainer.step(frames, targets)
        assert torch.isfinite(total)
    assert trainer.step_count == 2 and not torch.equal(w0, model.flat.w)
