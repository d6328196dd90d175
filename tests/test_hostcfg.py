"""Host-side configuration helpers (no GPU): the CPU quota the process may use and the thread-pool cap derived from it, and
the host copy of the MSDA geometry tensors."""
import numpy as np
import torch


def test_cpu_quota_and_thread_cap():
    from vnext_b200.hostcfg import cpu_quota, tune_host_threads
    q = cpu_quota()
    assert 1 <= q <= (torch.get_num_threads() if False else 4096)
    before = torch.get_num_threads()
    try:
        n = tune_host_threads(ranks_on_host=1, cap=4)
        assert 1 <= n <= max(1, min(4, q)) or n <= before          # never raises the pool, caps it at min(cap, quota)
        n8 = tune_host_threads(ranks_on_host=8, cap=4)
        assert 1 <= n8 <= n
    finally:
        torch.set_num_threads(before)


def test_host_geometry_is_tied_to_the_tensor_object():
    """A recycled device address must never serve another tensor's geometry: the host copy lives on the shapes tensor and is
    invalidated by in-place writes (CPU tensors stand in for device tensors here: the cache logic is device-agnostic)."""
    from vnext_b200 import msda
    shapes = torch.tensor([[6, 4], [3, 2]])
    lsi = torch.tensor([0, 24])
    hs, hl = msda.host_geometry(shapes, lsi)
    assert hs.tolist() == [[6, 4], [3, 2]] and hl.tolist() == [0, 24]
    assert msda.host_geometry(shapes, lsi)[0] is hs                # cached on the object
    shapes[0, 0] = 8                                               # in-place change bumps the version counter
    hs2, _ = msda.host_geometry(shapes, lsi)
    assert hs2.tolist() == [[8, 4], [3, 2]]
    other = torch.tensor([[5, 5], [2, 2]])
    assert msda.host_geometry(other, torch.tensor([0, 25]))[0].tolist() == [[5, 5], [2, 2]]
    # a caller that knows the geometry on the host registers it without any device round trip
    s3, l3 = torch.tensor([[4, 4], [2, 2]]), torch.tensor([0, 16])
    msda.attach_host_geometry(s3, l3, [(4, 4), (2, 2)])
    a, b = msda.host_geometry(s3, l3)
    assert a.tolist() == [[4, 4], [2, 2]] and b.tolist() == [0, 16] and a.dtype == np.int64
    assert msda._pyramid_in_order(a, b, 20) and not msda._pyramid_in_order(a, b, 21)
