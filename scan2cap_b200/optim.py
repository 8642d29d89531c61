"""Adam on flat buffers: the optimizer.step() of the reference's training loop (lib/solver.py:293-300; Adam(lr, weight_decay)
created at scripts/train.py:134) as ONE streaming kernel (s2c_adam_step, csrc/adam.cu) instead of the framework's
multi-tensor walk over CapNet's 144 parameter tensors (~27 launches, 0.43 ms per step).

FlatAdam is a torch.optim.Adam: same constructor arguments, param_groups, state layout ({"step", "exp_avg", "exp_avg_sq"}
per parameter) and state_dict()/load_state_dict() format, so checkpoints written by the reference's solver load unchanged.
What differs is where the tensors live: parameters, gradients (distributed.FlatGradients) and both moments are views into
one contiguous fp32 buffer each, in the same order.  Nothing on the host is read per step (no step counter, no learning
rate baked into a captured graph: the hyper-parameters sit in a small device tensor refreshed by sync_hyper())."""
import torch

from ._lib import call


class FlatAdam(torch.optim.Adam):
    def __init__(self, flat_grads, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = list(flat_grads.params)
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self._flat_grads = flat_grads
        dev = flat_grads.flat.device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam runs on the CUDA kernels of libs2c only (no CPU path)")
        npad = flat_grads.flat.numel()
        self._params = params
        self.flat_param = torch.zeros(npad, dtype=torch.float32, device=dev)
        self.flat_exp_avg = torch.zeros(npad, dtype=torch.float32, device=dev)
        self.flat_exp_avg_sq = torch.zeros(npad, dtype=torch.float32, device=dev)
        self.steps = torch.zeros(len(params), dtype=torch.float32, device=dev)
        self._coef = torch.zeros(2, dtype=torch.float32, device=dev)
        self._hyper = torch.zeros(5, dtype=torch.float32, device=dev)
        self._hyper_host = None
        with torch.no_grad():
            for p, off in zip(params, flat_grads.offsets):
                view = self.flat_param[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view           # the module keeps its Parameter objects; only their storage moves
        self._rehome_state(copy_from=None)
        self.sync_hyper()

    # ---- state: views into the flat moment buffers, in the layout torch.optim.Adam exposes ------------------------------
    def _rehome_state(self, copy_from):
        with torch.no_grad():
            for i, (p, off) in enumerate(zip(self._params, self._flat_grads.offsets)):
                n = p.numel()
                views = {"step": self.steps[i], "exp_avg": self.flat_exp_avg[off:off + n].view_as(p),
                         "exp_avg_sq": self.flat_exp_avg_sq[off:off + n].view_as(p)}
                old = copy_from.get(p) if copy_from is not None else None
                if old:
                    for k, v in views.items():
                        if k in old:
                            v.copy_(torch.as_tensor(old[k]).to(v.device, v.dtype))
                self.state[p] = views

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)           # casts and copies: the loaded tensors replace the views ...
        loaded = {p: dict(st) for p, st in self.state.items()}
        self._rehome_state(copy_from=loaded)           # ... so copy them back into the flat buffers
        self.sync_hyper()

    def zero_grad(self, set_to_none=False):
        self._flat_grads.zero_()                       # the .grad views must stay: never set them to None

    def add_param_group(self, param_group):
        if getattr(self, "_params", None) is not None:
            raise RuntimeError("FlatAdam: the flat buffers are laid out at construction; one parameter group only")
        super().add_param_group(param_group)

    # ---- hyper-parameters on the device --------------------------------------------------------------------------------
    def sync_hyper(self):
        """Upload lr / betas / eps / weight_decay when they changed (call before replaying a captured step; never
        inside a capture)."""
        g = self.param_groups[0]
        if g.get("amsgrad") or g.get("maximize"):
            raise RuntimeError("FlatAdam: amsgrad / maximize are not implemented")
        cur = (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]))
        if cur != self._hyper_host:
            self._hyper.copy_(torch.tensor(cur, dtype=torch.float32))
            self._hyper_host = cur

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        stream = torch.cuda.current_stream(self.flat_param.device).cuda_stream
        with torch.cuda.device(self.flat_param.device):
            call("s2c_adam_step", self.flat_param.data_ptr(), self._flat_grads.flat.data_ptr(),
                 self.flat_exp_avg.data_ptr(), self.flat_exp_avg_sq.data_ptr(), self.flat_param.numel(),
                 self._hyper.data_ptr(), self.steps.data_ptr(), self.steps.numel(), self._coef.data_ptr(),
                 float(grad_scale), stream)
        return loss
