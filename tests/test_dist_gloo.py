"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: clip sharding + max-over-ranks timing reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vnext_b200.dist import max_over_ranks, shard, sum_over_ranks
    clips = list(range(7))
    mine = shard(clips, rank, world)
    ms = 10.0 + 5.0 * rank                       # rank 1 is slower
    worst = max_over_ranks(ms)
    total = sum_over_ranks(len(mine))
    dist.barrier()
    q.put((rank, mine, worst, total))
    dist.destroy_process_group()


def test_shard_and_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, w0, t0), (r1, m1, w1, t1) = res
    assert sorted(m0 + m1) == list(range(7)) and not set(m0) & set(m1)     # every clip exactly once
    assert w0 == w1 == 15.0                                                # everyone reports the slowest rank
    assert t0 == t1 == 7.0


def _grad_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vnext_b200.idol.train import FlatParams, IDOLTrainer

    class Stub:                                    # the trainer only needs .flat and .criterion of the model
        class criterion:
            all_reduce_fn = None
    g = torch.Generator().manual_seed(0)
    tensors = {"detr.detr.backbone.0.backbone.res3.0.conv1.weight": torch.randn(8, 4, 1, 1, generator=g),
               "detr.detr.class_embed.0.weight": torch.randn(5, 7, generator=g),
               "detr.detr.class_embed.0.bias": torch.randn(5, generator=g)}
    model = Stub()
    model.flat = FlatParams(tensors, "cpu")
    trainer = IDOLTrainer(model, clip_value=0.5)
    loss = sum(((p * (rank + 1)) ** 2).sum() for p in model.flat.params.values())     # rank-dependent gradients
    loss.backward()
    local = model.flat.g.clone()
    norm = trainer.reduce_gradients()                                                 # ONE all-reduce of the flat buffer
    nb = torch.tensor([3.0 + rank])
    model.criterion.all_reduce_fn(nb)                                                 # the 4-byte num_boxes exchange
    q.put((rank, local.tolist(), model.flat.g.tolist(), float(norm), float(trainer.scale), float(nb),
           model.flat.groups(), {k: v[0] for k, v in model.flat.offsets.items()}))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, s0, n0, c0, nb0, groups, offs), (_, l1, s1, n1, c1, nb1, _, _) = res
    l0, l1, s0, s1 = (torch.tensor(x) for x in (l0, l1, s0, s1))          # plain lists cross the process boundary
    assert torch.equal(s0, s1) and torch.allclose(s0, l0 + l1)            # both ranks hold the summed gradient
    mean_norm = float(((l0 + l1) / 2).norm())
    assert abs(n0 - mean_norm) < 1e-5 and n0 == n1
    want = min(1.0, 0.5 / (mean_norm + 1e-6)) / 2                         # clip coefficient x 1/world, applied in AdamW
    assert abs(c0 - want) < 1e-7 and c0 == c1
    assert nb0 == nb1 == 7.0
    # layout: backbone group first (its own learning rate), 64-element aligned segments
    assert groups[0][0] == 0 and groups[0][1] == groups[1][0] == 64 and groups[1][1] == 192
    assert offs["detr.detr.backbone.0.backbone.res3.0.conv1.weight"] == 0 and all(o % 64 == 0 for o in offs.values())


def _numboxes_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vnext_b200.seqformer import criterion as crit
    c = crit.SetCriterion(5, crit.default_weight_dict(), 2, world_size_fn=dist.get_world_size,
                          all_reduce_fn=lambda t: dist.all_reduce(t))
    g = torch.Generator().manual_seed(7)                       # identical predictions on both ranks ...
    out = {"pred_logits": torch.randn(1, 6, 5, generator=g), "pred_boxes": torch.rand(1, 2, 6, 4, generator=g) * 0.5 + 0.2}
    n = 1 + 2 * rank                                           # ... but rank 0 has 1 ground truth, rank 1 has 3
    tg = [{"labels": torch.arange(n) % 5, "boxes": torch.rand(n, 2, 4, generator=g) * 0.4 + 0.3,
           "masks": torch.zeros(n, 2, 32, 32, dtype=torch.bool), "size": torch.tensor([32, 32])}]
    ind = c.matcher(out, tg)
    out["pred_masks"] = [torch.zeros(1, n, 2, 8, 8)]
    losses = c(out, tg, [ind])
    alone = crit.SetCriterion(5, crit.default_weight_dict(), 2)(out, tg, [ind])
    q.put((rank, float(losses["loss_bbox"]), float(alone["loss_bbox"]), n))
    dist.destroy_process_group()


def test_seqformer_num_boxes_is_the_cross_rank_mean_world2():
    """SeqFormer normalises every loss by the MEAN ground-truth count over ranks (deformable_detr.py:415-420)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_numboxes_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, dist_loss, solo_loss, n in res:
        assert abs(dist_loss - solo_loss * n / 2.0) < 1e-6 * max(1.0, abs(solo_loss))       # mean count = (1 + 3) / 2
