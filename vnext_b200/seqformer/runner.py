"""Static-shape execution of the SeqFormer engine: CUDA-graph capture + pinned host staging.

One `ClipRunner` serves clips of a fixed (T, H, W): the whole device part of
`SeqFormer.forward(eval)` (preprocess -> backbone -> encoder -> decoder -> heads -> top-k masks ->
upsample/threshold) is captured once into a CUDA graph and replayed per clip, so the ~10^3 kernel
launches of a clip cost one graph launch ("CUDA streams and graphs instead of a tracing compiler").
Host<->device traffic per clip: the uint8 frames in (pinned), the top-k scores and boolean masks out — or, with
mask_format="rle", the masks' COCO RLE strings (vnext_b200/rle.py: the evaluator's wire format, ytvis_eval.py:174-214,
encoded on the device inside the same graph), KBs instead of H*W bytes per mask.
"""
import torch


_warned_threads = False


def _warn_if_oversubscribed():
    global _warned_threads
    if _warned_threads:
        return
    _warned_threads = True
    from ..hostcfg import cpu_quota
    if torch.get_num_threads() > 2 * cpu_quota():
        import warnings
        warnings.warn(f"vnext_b200: torch uses {torch.get_num_threads()} intra-op threads but this process may use "
                      f"{cpu_quota()} CPUs (cgroup quota): the host staging copies will be throttled and clips will stall; call "
                      "vnext_b200.hostcfg.tune_host_threads() once at start-up")


class ClipRunner:
    RLE_FIRST_BYTES = 1 << 20          # string bytes copied with the offsets in one go (the rest, if any, in a second copy)

    def __init__(self, engine, T, H, W, ori_size=None, topk=10, use_graph=True, mask_format="bool"):
        _warn_if_oversubscribed()
        if mask_format not in ("bool", "rle"):
            raise ValueError(f"mask_format must be 'bool' or 'rle', got {mask_format!r}")
        self.engine, self.T, self.H, self.W = engine, T, H, W
        self.ori_size = tuple(ori_size) if ori_size is not None else (H, W)
        self.topk = topk
        self.mask_format = mask_format
        self.rle = self._rle_big = None
        self.direct_h2d = False        # measured slower than one staged copy in round 1 (five small copies); kept selectable
        dev = engine.device
        self.frames_dev = torch.zeros(T, 3, H, W, dtype=torch.uint8, device=dev)
        self.frames_pin = torch.zeros(T, 3, H, W, dtype=torch.uint8).pin_memory()
        self.stream = torch.cuda.Stream(device=dev)
        # non-blocking stream: order it after whatever stream packed the engine's weights / built the ClipPlan
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        self.graph = self.graph_head = None
        self._pending = None
        with torch.cuda.device(dev):
            with torch.cuda.stream(self.stream):
                for _ in range(2):                    # builds the ClipPlan, warms cuDNN/cuBLAS heuristics
                    self.scores, self.masks = self._device_step()
            self.stream.synchronize()
            if mask_format == "rle":
                from ..rle import RleEncoder
                k, t, oh, ow = self.masks.shape
                self.rle = RleEncoder(k * t, oh, ow, dev)
                first = min(self.rle.capacity, self.RLE_FIRST_BYTES)
                self._off_pin = torch.empty(k * t + 1, dtype=torch.int64).pin_memory()
                self._chars_pin = torch.empty(first, dtype=torch.uint8).pin_memory()
                with torch.cuda.stream(self.stream):
                    self.rle.encode(self.masks)            # warm-up outside the capture
                self.stream.synchronize()
            if use_graph:
                # two graphs sharing one memory pool: the small head (preprocess + backbone) is submitted in tens of
                # microseconds, so the GPU is busy while the host submits the large tail — a single ~1000-node graph
                # kept the GPU idle for the ~1.5 ms its launch takes (profiles/r02_stream_stress_and_e2e_breakdown.txt)
                self.graph_head = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_head, stream=self.stream):
                    self._feats = engine.infer_video_stage1(self.frames_dev)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=self.stream, pool=self.graph_head.pool()):
                    self.scores, self.masks = engine.infer_video_stage2(self._feats, self.frames_dev.shape, self.ori_size,
                                                                        self.topk)
                    if self.rle is not None:
                        self.rle.encode(self.masks)
        self.h2d_bytes = self.frames_pin.numel()
        self.d2h_bytes = self.scores.numel() * 4 + self.masks.numel()      # rle: updated per clip by collect()

    def _device_step(self):
        return self.engine.infer_video_device(self.frames_dev, self.ori_size, self.topk)   # uint8 in, fused normalise

    def run_device(self):
        """Inputs already resident in HBM (self.frames_dev); enqueues one clip on self.stream."""
        with torch.cuda.stream(self.stream):          # graph.replay() launches on the *current* stream
            if self.graph is not None:
                self.graph_head.replay()
                self.graph.replay()
            else:
                self.scores, self.masks = self._device_step()
                if self.rle is not None:
                    self.rle.encode(self.masks)

    def run_host(self, frames_u8=None, frames_list=None):
        """End to end: host frames -> pinned staging -> device -> graph -> pinned host outputs (+ sync).  `frames_list`:
        the caller's per-frame tensors (uint8-valued, float or uint8).  Copying already-pinned caller frames to the
        device one by one was measured slower (290 vs 313-328 frames/s end to end) than one staged copy, so everything
        goes through the staging buffer.  The outputs are FRESH pinned tensors every call (PyTorch's caching host
        allocator recycles a block only after the caller dropped it), so results of earlier clips are never
        overwritten — the reference returns new tensors too."""
        self.submit(frames_u8, frames_list)
        return self.collect()

    def submit(self, frames_u8=None, frames_list=None):
        """Stage the frames and enqueue H2D + graph + D2H on this runner's stream without waiting (see collect())."""
        direct = (frames_list is not None and self.direct_h2d and
                  all(f.dtype == torch.uint8 and f.is_contiguous() and f.is_pinned() for f in frames_list))
        block = None if direct else self._pinned_block(frames_list)
        if direct or block is not None:
            pass                                      # caller's frames are pinned uint8: no host staging copy
        elif frames_list is not None:
            for i, f in enumerate(frames_list):
                self.frames_pin[i].copy_(f)
        elif frames_u8 is not None:
            self.frames_pin.copy_(frames_u8)
        scores_pin = torch.empty(self.scores.shape, dtype=self.scores.dtype, pin_memory=True)
        masks_pin = None if self.rle is not None else torch.empty(self.masks.shape, dtype=self.masks.dtype, pin_memory=True)
        self.stream.wait_stream(torch.cuda.current_stream(self.frames_dev.device))
        with torch.cuda.stream(self.stream):
            if block is not None:                     # one DMA straight from the caller's pinned memory
                self.frames_dev.copy_(block, non_blocking=True)
            elif direct:
                for i, f in enumerate(frames_list):
                    self.frames_dev[i].copy_(f, non_blocking=True)
            else:
                self.frames_dev.copy_(self.frames_pin, non_blocking=True)
            self.run_device()
            scores_pin.copy_(self.scores, non_blocking=True)
            if self.rle is not None:       # offsets + the first bytes of the strings in one round trip
                self._off_pin.copy_(self.rle.char_offsets, non_blocking=True)
                self._chars_pin.copy_(self.rle.chars[:self._chars_pin.numel()], non_blocking=True)
            else:
                masks_pin.copy_(self.masks, non_blocking=True)
        self._pending = (scores_pin, masks_pin)

    def _pinned_block(self, frames_list):
        """The caller's frames as ONE [T,3,H,W] pinned uint8 tensor when they are adjacent views of one pinned allocation
        (e.g. a decoded clip held in a single buffer): a single H2D copy from the caller's memory, no staging copy."""
        if not frames_list or len(frames_list) != self.T:
            return None
        f0 = frames_list[0]
        if f0.dtype != torch.uint8 or tuple(f0.shape) != (3, self.H, self.W) or not f0.is_pinned():
            return None
        n = f0.numel()
        for i, f in enumerate(frames_list):
            if (f.dtype != torch.uint8 or f.shape != f0.shape or not f.is_contiguous() or
                    f.untyped_storage().data_ptr() != f0.untyped_storage().data_ptr() or
                    f.storage_offset() != f0.storage_offset() + i * n):
                return None
        return torch.as_strided(f0, (self.T, 3, self.H, self.W), (n, self.H * self.W, self.W, 1))

    def collect(self):
        """Wait for the clip enqueued by submit() and return its (scores, masks): pinned host tensors, or with
        mask_format="rle" masks = [k][T] RLE dicts {"size": [H, W], "counts": str}."""
        self.stream.synchronize()
        out, self._pending = self._pending, None
        if self.rle is None:
            return out
        return out[0], self._collect_rle()

    def _collect_rle(self):
        from ..rle import split_strings
        total = int(self._off_pin[-1])
        first = self._chars_pin.numel()
        chars = self._chars_pin
        if total > first:                  # long strings (noisy masks): fetch all of them in a second copy
            with torch.cuda.stream(self.stream):
                src = self.rle.chars
                if total > self.rle.capacity:
                    # the strings outgrew the buffer the captured graph writes (its address and capacity are baked into the
                    # graph, so that encoder is left alone): encode again, eagerly, into a larger buffer — same masks, same
                    # offsets; this runner has one clip in flight, self.masks is still this clip's
                    from ..rle import RleEncoder
                    if self._rle_big is None or self._rle_big.capacity < total:
                        k, t, oh, ow = self.masks.shape
                        self._rle_big = RleEncoder(k * t, oh, ow, self.masks.device, capacity=total + total // 4)
                    self._rle_big.encode(self.masks)
                    src = self._rle_big.chars
                chars = torch.empty(total, dtype=torch.uint8, pin_memory=True)
                chars.copy_(src[:total], non_blocking=True)
            self.stream.synchronize()
        self.d2h_bytes = self.scores.numel() * 4 + self._off_pin.numel() * 8 + max(total, first)
        k, t, oh, ow = self.masks.shape
        flat = split_strings(self._off_pin, chars, oh, ow)
        return [flat[i * t:(i + 1) * t] for i in range(k)]
