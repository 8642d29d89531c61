"""Data parallelism for CapNet: one process per GPU, scenes sharded over the batch axis, ONE NCCL all-reduce
per step over a single flat gradient buffer (the reference itself is single-GPU).

Every parameter's ``.grad`` is a view into one contiguous fp32 buffer, so autograd accumulates straight into
it and the all-reduce needs no flatten / unflatten copies.  Parameters that receive no gradient in a step (e.g.
the graph head when a scene has no edge) simply stay zero, which keeps the buffer layout static.
BatchNorm statistics stay per-GPU, exactly as N independent copies of the reference would behave."""
import torch
import torch.distributed as dist


ALIGN = 64  # floats


class FlatGradients(object):
    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        dev = self.params[0].device
        # every parameter's slice starts on a 256-byte boundary (the kernels read weights with 16-byte vector loads, and
        # optim.FlatAdam lays the parameters themselves out the same way); the gaps stay zero.  The total is a multiple
        # of 4 floats: the fused Adam kernel walks the buffer as float4.
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)]
        for p, v in zip(self.params, self.views):
            p.grad = v
        self.numel = sum(p.numel() for p in self.params)

    def zero_(self):
        self.flat.zero_()

    # ---- one step without per-parameter accumulation kernels ---------------------------------------------------------------
    # With .grad pointing at a zeroed view, autograd's AccumulateGrad issues one `grad += new` kernel per parameter
    # (106 launches per CapNet step).  release() drops the views before backward, so AccumulateGrad simply keeps the
    # tensor each backward function returned; collect() then moves all of them into the flat buffer with ONE multi-tensor
    # copy, zero-fills the slices of parameters that received nothing, and re-installs the views as .grad.
    def release(self):
        for p in self.params:
            p.grad = None

    def collect(self):
        srcs, dsts, strided, missing = [], [], [], []
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None:
                missing.append(v)
            elif g is not v and g.data_ptr() != v.data_ptr():
                if g.is_contiguous() and g.dtype == v.dtype and g.shape == v.shape:
                    srcs.append(g)
                    dsts.append(v)
                else:
                    strided.append((v, g))
            p.grad = v
        with torch.no_grad():
            if missing:
                torch._foreach_zero_(missing)
            if srcs:
                torch._foreach_copy_(dsts, srcs)
            for v, g in strided:
                v.copy_(g)

    def all_reduce_mean(self, group=None):
        """Mean over ranks in ONE collective: ncclAllReduce(avg) (the 1/world scale is applied inside NCCL's reduction,
        no extra kernel); backends without AVG (gloo, used by the CPU tests) sum and scale."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self.flat.is_cuda and dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.mul_(1.0 / dist.get_world_size(group))


def shard_batch(data_dict, rank, world_size):
    """Rank r's contiguous slice of every per-scene tensor (batch axis 0)."""
    out = {}
    for k, v in data_dict.items():
        if isinstance(v, torch.Tensor) and v.dim() > 0:
            B = v.shape[0]
            assert B % world_size == 0, "global batch %d not divisible by world size %d" % (B, world_size)
            per = B // world_size
            out[k] = v[rank * per:(rank + 1) * per]
        else:
            out[k] = v
    return out
