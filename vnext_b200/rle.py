"""COCO RLE of predicted masks on the device (include/vnext_b200.h: vnext_rle_encode) — the evaluator's wire format
(projects/IDOL/idol/data/ytvis_eval.py:174-214: `mask_util.encode(np.array(mask[:, :, None], order="F"))[0]`, then
`counts.decode("utf-8")`), produced before the masks ever leave the GPU: the host receives the strings (KBs per clip for
real masks) instead of H*W bytes per mask.
"""
import torch

from . import _cabi


class RleEncoder:
    """Fixed-shape encoder: n_masks masks of [H, W]; owns the workspace and the output buffers (graph-capturable: every
    launch has a fixed grid, the buffers never move).  `capacity` = bytes of the string buffer (default H*W/2 per mask — a
    string only gets that long for masks that flip every other pixel; `encode` reports the needed size so the caller can
    retry with `grow()`)."""

    def __init__(self, n_masks, H, W, device, capacity=None):
        self.n_masks, self.H, self.W = int(n_masks), int(H), int(W)
        self.device = torch.device(device)
        need = int(_cabi.lib().vnext_rle_workspace_bytes(self.n_masks, self.H, self.W))
        self.workspace = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
        self.char_offsets = torch.zeros(self.n_masks + 1, dtype=torch.int64, device=self.device)
        self.capacity = int(capacity if capacity is not None else max(4096, self.n_masks * self.H * self.W // 2))
        self.chars = torch.empty(self.capacity, dtype=torch.uint8, device=self.device)

    def grow(self, capacity):
        """A larger string buffer (a NEW allocation: never call this on an encoder whose encode() was captured in a CUDA graph —
        the graph keeps writing the old buffer; make a second encoder instead, as ClipRunner does)."""
        self.capacity = int(capacity)
        self.chars = torch.empty(self.capacity, dtype=torch.uint8, device=self.device)

    def encode(self, masks):
        """masks: bool / uint8 tensor of n_masks*H*W elements on the device (any leading shape) -> (char_offsets, chars),
        both device tensors owned by this encoder (overwritten by the next call)."""
        if not masks.is_cuda:
            raise RuntimeError("vnext_b200.rle: CUDA tensors only (no CPU fallback)")
        if masks.dtype not in (torch.bool, torch.uint8) or not masks.is_contiguous() or \
                masks.numel() != self.n_masks * self.H * self.W:
            raise RuntimeError(f"vnext_b200.rle: expected a contiguous bool/uint8 tensor of {self.n_masks}x{self.H}x{self.W} "
                               f"elements, got {masks.dtype} {tuple(masks.shape)}")
        info = {"bytes": 2.0 * masks.numel()}
        with torch.cuda.device(self.device):
            rc = _cabi.timed_call("rle_encode", info, masks, lambda: _cabi.lib().vnext_rle_encode(
                masks.data_ptr(), self.n_masks, self.H, self.W, self.workspace.data_ptr(), self.workspace.numel(),
                self.char_offsets.data_ptr(), self.chars.data_ptr(), self.capacity,
                torch.cuda.current_stream(self.device).cuda_stream))
        _cabi.check(rc, "vnext_rle_encode")
        return self.char_offsets, self.chars


def split_strings(char_offsets, chars, H, W):
    """Host side: char_offsets [n+1] (host int64 tensor / array), chars (host uint8 tensor / bytes holding at least
    char_offsets[-1] bytes) -> n dicts {"size": [H, W], "counts": str}, the entries of a COCO "segmentation" /
    YTVIS "segmentations" list (ytvis_eval.py:196-204)."""
    offs = [int(v) for v in (char_offsets.tolist() if hasattr(char_offsets, "tolist") else char_offsets)]
    raw = bytes(chars) if isinstance(chars, (bytes, bytearray, memoryview)) else chars[:offs[-1]].numpy().tobytes()
    text = raw.decode("ascii")
    return [{"size": [int(H), int(W)], "counts": text[a:b]} for a, b in zip(offs[:-1], offs[1:])]


def encode(masks):
    """One-off convenience: device masks [..., H, W] -> list (flattened leading dims) of RLE dicts.  Synchronises."""
    H, W = masks.shape[-2:]
    n = masks.numel() // (H * W) if H * W else 0
    enc = RleEncoder(n, H, W, masks.device)
    offs, chars = enc.encode(masks.contiguous())
    total = int(offs[-1].item())
    if total > enc.capacity:
        enc.grow(total)
        offs, chars = enc.encode(masks.contiguous())
    return split_strings(offs.cpu(), chars[:total].cpu(), H, W)
