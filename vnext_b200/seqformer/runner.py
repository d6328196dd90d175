"""Static-shape execution of the SeqFormer engine: CUDA-graph capture + pinned host staging.

One `ClipRunner` serves clips of a fixed (T, H, W): the whole device part of
`SeqFormer.forward(eval)` (preprocess -> backbone -> encoder -> decoder -> heads -> top-k masks ->
upsample/threshold) is captured once into a CUDA graph and replayed per clip, so the ~10^3 kernel
launches of a clip cost one graph launch ("CUDA streams and graphs instead of a tracing compiler").
Host<->device traffic per clip: the uint8 frames in (pinned), the top-k scores and boolean masks out.
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
    def __init__(self, engine, T, H, W, ori_size=None, topk=10, use_graph=True):
        _warn_if_oversubscribed()
        self.engine, self.T, self.H, self.W = engine, T, H, W
        self.ori_size = tuple(ori_size) if ori_size is not None else (H, W)
        self.topk = topk
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
        self.h2d_bytes = self.frames_pin.numel()
        self.d2h_bytes = self.scores.numel() * 4 + self.masks.numel()

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
        if direct:
            pass                                      # caller's frames are pinned uint8: no host staging copy
        elif frames_list is not None:
            for i, f in enumerate(frames_list):
                self.frames_pin[i].copy_(f)
        elif frames_u8 is not None:
            self.frames_pin.copy_(frames_u8)
        scores_pin = torch.empty(self.scores.shape, dtype=self.scores.dtype, pin_memory=True)
        masks_pin = torch.empty(self.masks.shape, dtype=self.masks.dtype, pin_memory=True)
        self.stream.wait_stream(torch.cuda.current_stream(self.frames_dev.device))
        with torch.cuda.stream(self.stream):
            if direct:
                for i, f in enumerate(frames_list):
                    self.frames_dev[i].copy_(f, non_blocking=True)
            else:
                self.frames_dev.copy_(self.frames_pin, non_blocking=True)
            self.run_device()
            scores_pin.copy_(self.scores, non_blocking=True)
            masks_pin.copy_(self.masks, non_blocking=True)
        self._pending = (scores_pin, masks_pin)

    def collect(self):
        """Wait for the clip enqueued by submit() and return its (scores, masks) pinned host tensors."""
        self.stream.synchronize()
        out, self._pending = self._pending, None
        return out
