"""Training loop with the call surface of the reference's Solver (lib/solver.py:78-232, _feed :366-468), built on
engine.TrainStep: the whole iteration is one CUDA-graph replay, the NEXT batch's host->device copies and FPS indices run
on a side stream meanwhile (prefetch), and nothing reads a device value back per iteration -- the reference's
``.item()`` calls per loss key (:424-440) and ``CUDA_LAUNCH_BLOCKING=1`` (scripts/train.py:354) make every iteration
host-bound.  Loss scalars are kept on the device and fetched in one transfer every `verbose` iterations.

    solver = Solver(model, device, DC, dataset, dataloader, optimizer=None, stamp="run", detection=True, caption=True,
                    orientation=True)
    solver(epoch=1, verbose=10)

Differences from the reference, on purpose: the optimiser is created inside TrainStep (Adam, lr / weight decay of
scripts/train.py:134; a capturable optimiser is needed for the graph) -- an `optimizer` argument is accepted only to
read lr / weight_decay from; evaluation (capeval, METEOR via java) and tensorboard logging are out of scope; the
`val_step` hook calls an optional user callback instead."""
import time

import torch

from ..engine import TrainStep

LOG_KEYS = ("loss", "cap_loss", "ori_loss", "dist_loss", "objectness_loss", "vote_loss", "box_loss", "cap_acc",
            "ori_acc", "obj_acc", "pred_ious", "pos_ratio", "neg_ratio")


class Solver(object):
    def __init__(self, model, device, config, dataset, dataloader, optimizer=None, stamp="", val_step=10,
                 detection=True, caption=True, orientation=False, distance=False, use_tf=True, lr_decay_step=None,
                 lr_decay_rate=None, bn_decay_step=None, bn_decay_rate=None, criterion="meteor", checkpoint_best=None,
                 use_cuda_graph=True, on_validation=None):
        self.model, self.device, self.config = model, device, config
        self.dataset, self.dataloader, self.stamp, self.val_step = dataset, dataloader, stamp, val_step
        lr, wd = 1e-3, 1e-5
        if optimizer is not None:
            lr = optimizer.param_groups[0].get("lr", lr)
            wd = optimizer.param_groups[0].get("weight_decay", wd)
        assert use_tf, "training uses teacher forcing (lib/solver.py:294)"
        self.engine = TrainStep(model, config, lr=lr, weight_decay=wd, detection=detection, caption=caption,
                                orientation=orientation, distance=distance, use_cuda_graph=use_cuda_graph)
        self.on_validation = on_validation
        self.log = {"train": {k: [] for k in LOG_KEYS + ("iter_time",)}}
        self._global_iter_id = 0
        self.epoch = self.verbose = 0

    def __call__(self, epoch, verbose):
        self.epoch, self.verbose = epoch, verbose
        for epoch_id in range(epoch):
            self._feed(self.dataloader["train"], "train", epoch_id)
        return self.log

    def _flush(self, pending, t0):
        """One device->host transfer for the loss scalars of the iterations since the last report."""
        if not pending:
            return
        stacked = torch.stack([torch.stack([p[k].detach().float().reshape(()) for k in LOG_KEYS]) for p in pending])
        host = stacked.cpu()   # the only synchronisation of the loop
        dt = (time.time() - t0) / len(pending)
        for row in host:
            for k, v in zip(LOG_KEYS, row.tolist()):
                self.log["train"][k].append(v)
            self.log["train"]["iter_time"].append(dt)
        if self.verbose:
            last = self.log["train"]
            print("[train] iter %d  loss %.4f  cap %.4f  vote %.4f  obj %.4f  box %.4f  (%.1f ms / iter)" % (
                self._global_iter_id, last["loss"][-1], last["cap_loss"][-1], last["vote_loss"][-1],
                last["objectness_loss"][-1], last["box_loss"][-1], 1e3 * dt), flush=True)

    def _feed(self, dataloader, phase, epoch_id):
        assert phase == "train"
        self.model.train()
        it = iter(dataloader)
        nxt = next(it, None)
        if nxt is not None:
            self.engine.prefetch(nxt)
        pending, t0 = [], time.time()
        while nxt is not None:
            cur = nxt
            self.engine.run(cur)
            nxt = next(it, None)
            if nxt is not None:
                self.engine.prefetch(nxt)   # copies + FPS indices of the next batch overlap the running step
            out = self.engine.last
            # the graph's outputs live in graph-owned memory that the next replay overwrites: clone the few scalars
            pending.append({k: out[k].clone() if k in out else torch.zeros((), device=self.device) for k in LOG_KEYS})
            self._global_iter_id += 1
            if self.verbose and self._global_iter_id % self.verbose == 0:
                self._flush(pending, t0)
                pending, t0 = [], time.time()
            if self.on_validation is not None and self._global_iter_id % self.val_step == 0:
                self.on_validation(self, epoch_id)
        self._flush(pending, t0)
