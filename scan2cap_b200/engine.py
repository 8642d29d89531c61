"""Training-step engine: the whole CapNet step (zero grads -> forward -> loss -> backward -> [all-reduce] -> Adam)
as ONE CUDA graph replay.

The reference's step is ~3 500 kernel launches issued one by one from Python (with CUDA_LAUNCH_BLOCKING=1 in
its own scripts, scripts/train.py:354), so on a B200 it is bound by launch latency, not by the GPU.  The product
path is written without host synchronisation or data-dependent shapes, which makes the whole step capturable:
the GPU then runs back-to-back kernels with no host in the loop.  Graphs are cached per input signature
(shapes / dtypes / number of teacher-forced words); inputs are copied into static device buffers, from pinned
host memory when the caller passes host tensors.

With more than one rank the step is two graphs with the NCCL all-reduce of the flat gradient buffer between them.
"""
import torch
import torch.distributed as dist

from .distributed import FlatGradients
from .lib.loss_helper import get_scene_cap_loss


class TrainStep(object):
    def __init__(self, model, dataset_config, lr=1e-3, weight_decay=1e-5, detection=True, caption=True,
                 orientation=False, distance=False, use_cuda_graph=True, loss_fn=None):
        self.model = model
        self.DC = dataset_config
        self.flags = dict(detection=detection, caption=caption, orientation=orientation, distance=distance)
        self.device = next(model.parameters()).device
        self.flat = FlatGradients(model)
        self.use_graph = use_cuda_graph and self.device.type == "cuda"
        self.opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay, capturable=self.use_graph)
        self.loss_fn = loss_fn or get_scene_cap_loss
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._graphs = {}
        self._stage = {}           # signature -> staging copies of the static input buffers (prefetch target)
        self._prefetched = None    # (data_dict object, signature) whose inputs are in flight / in the staging buffers
        self._copy_stream = None
        self._stage_free = None    # event: the staging buffers have been copied into the static inputs
        self.kernels_per_step = None
        self.last = None  # data_dict of the last step (outputs live in graph-owned memory when graphed)

    # ---- the step, eager -------------------------------------------------------------------------------
    def _fwd_bwd(self, data):
        self.flat.zero_()
        out = self.loss_fn(self.model(data), self.device, self.DC, None, **self.flags)
        out["loss"].backward()
        return out

    def _step_eager(self, data):
        out = self._fwd_bwd(data)
        self.flat.all_reduce_mean()
        self.opt.step()
        return out

    # ---- graph capture -------------------------------------------------------------------------------
    @staticmethod
    def _signature(data):
        return tuple(sorted((k, tuple(v.shape), str(v.dtype)) if isinstance(v, torch.Tensor) else (k, v)
                            for k, v in data.items()))

    def _capture(self, data):
        static = {k: (torch.empty(v.shape, dtype=v.dtype, device=self.device) if isinstance(v, torch.Tensor) else v)
                  for k, v in data.items()}
        self._load(static, data)
        # warm-up on a side stream (allocator / cuBLAS workspaces / lazy kernel loading), as capture requires
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._step_eager(dict(static))
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        from . import _lib
        n0 = _lib.LAUNCH_COUNT
        g1 = torch.cuda.CUDAGraph()
        g2 = None
        if self.world == 1:
            with torch.cuda.graph(g1):
                out = self._fwd_bwd(dict(static))
                self.opt.step()
        else:
            with torch.cuda.graph(g1):
                out = self._fwd_bwd(dict(static))
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, pool=g1.pool()):
                self.opt.step()
        self.kernels_per_step = _lib.LAUNCH_COUNT - n0  # libs2c entry points captured in one step
        return static, g1, g2, out

    def run_eager(self, data_dict):
        """The same step issued kernel by kernel (used by bench.py to time single kernels with CUDA events)."""
        data = {k: (v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v)
                for k, v in data_dict.items()}
        self.last = self._step_eager(data)
        return self.last["loss"]

    def _load(self, static, data):
        for k, v in data.items():
            if isinstance(v, torch.Tensor):
                static[k].copy_(v, non_blocking=True)

    # ---- public ----------------------------------------------------------------------------------------------
    def prefetch(self, data_dict):
        """Start copying the NEXT step's inputs (pinned host tensors) to the device on a copy stream, so the transfer
        overlaps the step that is currently running; pass the SAME dict object to run() afterwards (and do not modify
        its tensors in between).  A no-op until the graph for this input signature exists, or without CUDA graphs."""
        if not self.use_graph:
            return
        if "num_words" not in data_dict:
            data_dict["num_words"] = int(data_dict["lang_len"].max().item())
        sig = self._signature(data_dict)
        if sig not in self._graphs:
            return
        static = self._graphs[sig][0]
        if sig not in self._stage:
            self._stage[sig] = {k: torch.empty_like(v) for k, v in static.items() if isinstance(v, torch.Tensor)}
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        if self._stage_free is not None:
            # wait only for the staging -> static copies of the step that consumed the staging buffers last (an event
            # recorded before that step's graph replay), NOT for the step itself: the transfer overlaps its compute
            self._copy_stream.wait_event(self._stage_free)
        with torch.cuda.stream(self._copy_stream):
            for k, v in data_dict.items():
                if isinstance(v, torch.Tensor):
                    self._stage[sig][k].copy_(v, non_blocking=True)
        self._prefetched = (data_dict, sig)

    def run(self, data_dict):
        """One training step on `data_dict` (host or device tensors; include the Python int "num_words" =
        lang_len.max() to avoid a device->host read).  Returns the (device) scalar loss."""
        if "num_words" not in data_dict:
            data_dict = dict(data_dict)
            data_dict["num_words"] = int(data_dict["lang_len"].max().item())
        if not self.use_graph:
            data = {k: (v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v)
                    for k, v in data_dict.items()}
            self.last = self._step_eager(data)
            return self.last["loss"]
        sig = self._signature(data_dict)
        if sig not in self._graphs:
            self._graphs[sig] = self._capture(data_dict)
        static, g1, g2, out = self._graphs[sig]
        if self._prefetched is not None and self._prefetched[0] is data_dict and self._prefetched[1] == sig:
            # inputs already on the device (prefetch): wait for the copy stream, then staging -> static (device copies)
            torch.cuda.current_stream(self.device).wait_stream(self._copy_stream)
            for k, v in self._stage[sig].items():
                static[k].copy_(v, non_blocking=True)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream(self.device))
        else:
            self._load(static, data_dict)
        self._prefetched = None
        g1.replay()
        if g2 is not None:
            self.flat.all_reduce_mean()
            g2.replay()
        self.last = out
        return out["loss"]
