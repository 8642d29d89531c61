"""Flip-aware parity machinery shared by the full-network GPU tests.

The product (scan2cap_b200, libs2c kernels) and the oracle (oracle/ref_model.py over the reference's own CUDA kernels)
evaluate the same fp32 network in different summation orders, so activations differ by ~1e-6 relative.  ReLU and
max-pool are discontinuous: among ~10^8 activations a handful sit within 1e-6 of zero (or of the runner-up) and
take DIFFERENT branches in the two implementations; each such flip changes the gradients by O(1 / sqrt(#rows)),
which is what made round 1 use a loose global bar.  Here the decisions of both sides are captured and compared:

  * per shared-MLP call (SA1..SA4, FP1, FP2, vote aggregation): ReLU masks of every layer and the max-pool routing;
  * a GROUP (one pooled output vector; one point for the FP modules) in which any decision differs is excluded from
    BOTH backward passes (its output gradient is zeroed by a tensor hook), the number of such groups is reported;
  * the two small point-wise heads (voting module, proposal head: Conv1d -> BatchNorm -> ReLU twice) are handled by
    margin instead: a row (seed / proposal) with any pre-ReLU activation of the ORACLE within HEAD_MARGIN x mean|y| of
    zero could flip, and is excluded from both backward passes the same way (hooks on the heads' output tensors);
  * EdgeConv message MLPs (graph module): the ReLU masks of every edge are compared exactly; a flipped edge's message
    gradient is zeroed on both sides.  Caption decoder: the ReLU masks of map_topdown / map_lang are compared per
    (scene, word); from the first flipped word of a scene on, that scene's logits are excluded from both backward
    passes (the recurrence carries the flip forward);
  * all remaining groups took identical decisions, so every parameter's gradient must agree to fp32 accuracy:
    relative L2 <= GRAD_RTOL per parameter, measured against max(|g|, 1e-2 x the largest gradient norm of the same
    sub-module) (bias-like parameters in front of a BatchNorm have an analytically zero gradient).
"""
import os

import torch

GRAD_RTOL = 1e-3
HEAD_MARGIN = 5e-4
HEAD_KEYS = {   # head -> (oracle BatchNorm modules in front of its ReLUs, data_dict outputs that carry its rows)
    "vgen": (("vgen.bn1", "vgen.bn2"), ("vote_xyz", "vote_features")),
    "proposal.proposal": (("proposal.proposal.1", "proposal.proposal.4"),
                          ("objectness_scores", "center", "heading_scores", "heading_residuals_normalized",
                           "size_scores", "size_residuals_normalized", "sem_cls_scores")),
}
MODULES = ["backbone_net.sa1", "backbone_net.sa2", "backbone_net.sa3", "backbone_net.sa4", "backbone_net.fp1",
           "backbone_net.fp2", "proposal.vote_aggregation"]


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _get(model, dotted):
    m = model
    for part in dotted.split("."):
        m = getattr(m, part)
    return m


class DecisionTracker(object):
    """Captures the ReLU / max-pool decisions of the seven shared-MLP calls on both sides, and (after `resolve`)
    zeroes the output gradient of every group whose decisions differ -- identically in both backward passes."""

    def __init__(self, ours, ref):
        from scan2cap_b200.lib.pointnet2 import fused_mlp
        self.fused_mlp = fused_mlp
        from scan2cap_b200.lib import caption_decoder
        from scan2cap_b200.models import graph_module
        self.caption_decoder, self.graph_module = caption_decoder, graph_module
        self.edge_names = [n for n in ("graph.gc_layers.0", "graph.gc_layers.1", "graph.edge_layer")
                           if self._has(ref, n)]
        self.ref_edge_hidden = {n: [] for n in self.edge_names}   # per call (= per scene): post-ReLU (E_b, 128)
        self.edge_keep = {}                                       # (name, call) -> (E_b) float, oracle side
        self.ref_cap = {"u": [], "lang": []}                      # per word: post-ReLU (B, emb)
        self.ours, self.ref = ours, ref
        self.keep = {}        # module name -> (B, M) float 0/1, filled by resolve()
        self.flips = {}       # module name -> (#groups with a differing decision, #groups)
        self.ref_masks = {name: [] for name in MODULES}
        self.handles = []
        self.head_risk = {}   # head -> (B, rows) bool: some pre-ReLU activation of the oracle is within the margin of 0
        for head, (bns, _) in HEAD_KEYS.items():
            for bn in bns:
                self.handles.append(_get(ref, bn).register_forward_hook(self._head_hook(head)))
        for n in self.edge_names:
            conv = _get(ref, n)
            self.handles.append(conv.map_edge[1].register_forward_hook(
                lambda mod, inp, out, n=n: self.ref_edge_hidden[n].append(out.detach())))
            self.handles.append(conv.map_edge.register_forward_hook(self._ref_edge_out_hook(n)))
        if hasattr(ref, "caption"):
            self.handles.append(ref.caption.map_topdown.register_forward_hook(
                lambda mod, inp, out: self.ref_cap["u"].append(out.detach())))
            self.handles.append(ref.caption.map_lang.register_forward_hook(
                lambda mod, inp, out: self.ref_cap["lang"].append(out.detach())))
        for name in MODULES:
            mo, mr = _get(ours, name), _get(ref, name)
            mlp_r = mr.mlp_module if hasattr(mr, "mlp_module") else mr.mlp
            for layer in mlp_r.children():
                self.handles.append(layer.register_forward_hook(self._ref_layer_hook(name)))
            self.handles.append(mo.register_forward_hook(self._out_hook(name)))
            self.handles.append(mr.register_forward_hook(self._out_hook(name)))

    @staticmethod
    def _has(model, dotted):
        try:
            _get(model, dotted)
            return True
        except (AttributeError, IndexError):
            return False

    def _ref_edge_out_hook(self, name):
        def hook(mod, inp, out):
            call = len(self.ref_edge_hidden[name]) - 1   # the ReLU hook of this call has already fired
            if out.requires_grad:
                out.register_hook(lambda g, key=(name, call): g * self.edge_keep[key].to(g.dtype).unsqueeze(-1))
        return hook

    def resolve_graph_and_caption(self, o, r):
        """Exact ReLU decisions of the EdgeConv message MLPs (per edge) and of the caption decoder (per scene, word)."""
        gcap, ccap = self.graph_module.CAPTURE, self.caption_decoder.CAPTURE
        self.graph_module.CAPTURE = self.caption_decoder.CAPTURE = None
        if self.edge_names:
            emask = gcap[0]["edge_mask"]                 # (B, K*L) bool: slot (i, t) is an edge of the compacted graph
            B = emask.shape[0]
            layers = gcap[1:]
            assert len(layers) == len(self.edge_names), (len(layers), self.edge_names)
            nflip = nedge = 0
            for name, cap in zip(self.edge_names, layers):
                hid = cap["hidden"].view(B, emask.shape[1], -1) > 0
                keep = torch.ones(emask.shape, dtype=torch.float32, device=emask.device)
                scenes = [b for b in range(B) if int(emask[b].sum()) > 0] if name.endswith("edge_layer") else list(range(B))
                assert len(scenes) == len(self.ref_edge_hidden[name]), (name, len(scenes), len(self.ref_edge_hidden[name]))
                for call, b in enumerate(scenes):
                    theirs = self.ref_edge_hidden[name][call] > 0            # (E_b, 128), same edge order
                    mine = hid[b][emask[b]]
                    assert mine.shape == theirs.shape, (name, b, mine.shape, theirs.shape)
                    same = (mine == theirs).all(-1)
                    self.edge_keep[(name, call)] = same.float()
                    keep[b][emask[b]] = same.float()
                    nflip += int((~same).sum())
                    nedge += same.numel()
                cap["message"].register_hook(lambda g, k=keep.reshape(-1, 1): g * k.to(g.dtype))
            self.flips["graph edges"] = (nflip, nedge)
        if ccap:
            assert len(ccap) == 1
            u_o, l_o = ccap[0]["u"] > 0, ccap[0]["lang"] > 0              # (T, B, emb)
            u_r = torch.stack(self.ref_cap["u"], 0) > 0
            l_r = torch.stack(self.ref_cap["lang"], 0) > 0
            assert u_o.shape == u_r.shape, (u_o.shape, u_r.shape)
            flip = ((u_o != u_r).any(-1) | (l_o != l_r).any(-1)).t()       # (B, T)
            keep = (torch.cumsum(flip.long(), 1) == 0).float()            # words before the scene's first flip
            self.flips["caption words"] = (int((keep == 0).sum()), keep.numel())
            for d in (o, r):
                d["lang_cap"].register_hook(lambda g, k=keep.unsqueeze(-1): g * k.to(g.dtype))

    # -- capture ---------------------------------------------------------------------------------------------
    def _ref_layer_hook(self, name):
        def hook(mod, inp, out):
            self.ref_masks[name].append(out.detach())   # post-ReLU activations (B,C,M,ns); > 0 <=> pre-activation > 0
        return hook

    def _head_hook(self, head):
        def hook(mod, inp, out):   # out: (B, C, rows) pre-ReLU
            y = out.detach()
            risk = (y.abs() < HEAD_MARGIN * y.abs().mean()).any(1)
            self.head_risk[head] = risk if head not in self.head_risk else (self.head_risk[head] | risk)
        return hook

    def exclude_head_rows(self, o, r):
        """Zero the gradient of the at-risk rows of the two heads at their output tensors, in both data_dicts."""
        for head, (_, keys) in HEAD_KEYS.items():
            keep = (~self.head_risk[head]).float()    # (B, rows)
            self.flips[head] = (int(self.head_risk[head].sum()), keep.numel())
            for d in (o, r):
                # (the product's fused loss kernel reads the packed head outputs instead of their slices)
                for k in keys + (("_head_outputs",) if head == "proposal.proposal" and "_head_outputs" in d else ()):
                    t = d[k]
                    if not t.requires_grad:
                        continue
                    if t.shape[1] == keep.shape[1]:     # (B, rows, ...)
                        m = keep.view(keep.shape + (1,) * (t.dim() - 2))
                    else:                                # (B, C, rows)
                        m = keep.unsqueeze(1)
                    t.register_hook(lambda g, m=m: g * m.to(g.dtype))

    def _out_hook(self, name):
        def hook(mod, inp, out):
            feats = out[1] if isinstance(out, tuple) else out    # (B,C,M) pooled features / (B,C,n) FP output
            if feats.requires_grad:
                feats.register_hook(lambda g: g * self.keep[name].to(g.dtype).unsqueeze(1))
        return hook

    def __enter__(self):
        self.fused_mlp.CAPTURE = []
        self.graph_module.CAPTURE = []
        self.caption_decoder.CAPTURE = []
        self.head_risk = {}
        return self

    def __exit__(self, *a):
        self.fused_mlp.CAPTURE = None
        self.graph_module.CAPTURE = self.caption_decoder.CAPTURE = None
        for h in self.handles:
            h.remove()

    def ours_captured(self):
        cap = self.fused_mlp.CAPTURE
        self.fused_mlp.CAPTURE = None   # stop capturing (later forward calls would append)
        assert len(cap) == len(MODULES), "expected %d shared-MLP calls, saw %d" % (len(MODULES), len(cap))
        return dict(zip(MODULES, cap))

    # -- compare -------------------------------------------------------------------------------------------
    def resolve(self, o=None, r=None):
        ours = self.ours_captured()
        total = 0
        membership = None
        if o is not None:
            # vote aggregation groups votes whose COORDINATES differ by ~1e-6 between the two sides: a vote within that
            # distance of a ball's surface is a member on one side only (the other six grouping stages work on
            # bit-identical coordinates).  Same kernel on both sides' tensors -> per-group "same neighbour list".
            from scan2cap_b200.lib.pointnet2 import _ext
            va = self.ours.proposal.vote_aggregation
            io = _ext.ball_query(o["aggregated_vote_xyz"].detach().contiguous(), o["vote_xyz"].detach().contiguous(),
                                 va.radius, va.nsample)
            ir = _ext.ball_query(r["aggregated_vote_xyz"].detach().contiguous(), r["vote_xyz"].detach().contiguous(),
                                 va.radius, va.nsample)
            membership = (io == ir).all(-1)
        for name in MODULES:
            o, r = ours[name], self.ref_masks[name]
            L, ns = len(o["Ys"]), o["ns"]
            assert len(r) == L, (name, len(r), L)
            B, _, M, ns_r = r[0].shape
            assert ns_r == ns and B * M == o["G"], (name, r[0].shape, o["G"], ns)
            same = torch.ones((B, M), dtype=torch.bool, device=r[0].device)
            for l in range(L - 1):
                scale, shift = o["affine"][l]
                mine = (torch.addcmul(shift, o["Ys"][l], scale) > 0).view(B, M, ns, -1)
                theirs = (r[l] > 0).permute(0, 2, 3, 1)
                same &= (mine == theirs).flatten(2).all(-1)
            # last layer + pooling: the pooled value is positive on both sides or on neither, and each side's arg-max
            # is a maximiser on the other side too (exact ties -- padded / duplicated neighbours -- are not flips)
            scale, shift = o["affine"][L - 1]
            act_o = torch.relu(torch.addcmul(shift, o["Ys"][L - 1], scale)).view(B, M, ns, -1)   # (B,M,ns,C)
            act_r = r[L - 1].permute(0, 2, 3, 1)
            max_o, max_r = act_o.amax(2), act_r.amax(2)                                        # (B,M,C)
            am_o = o["argmax"].view(B, M, -1).long()
            am_r = act_r.argmax(2)
            pos_same = (max_o > 0) == (max_r > 0)
            o_in_r = torch.gather(act_r, 2, am_o.unsqueeze(2)).squeeze(2) == max_r
            r_in_o = torch.gather(act_o, 2, am_r.unsqueeze(2)).squeeze(2) == max_o
            dead = (max_o <= 0) & (max_r <= 0)     # no gradient flows through an all-rectified channel
            same &= (pos_same & ((o_in_r & r_in_o) | dead)).all(-1)
            if name == "proposal.vote_aggregation" and membership is not None:
                same &= membership
            self.keep[name] = same.float()
            nflip = int((~same).sum())
            self.flips[name] = (nflip, B * M)
            total += nflip
        self.ref_masks = {name: [] for name in MODULES}   # free
        return total

    def report(self):
        return ", ".join("%s %d/%d" % (n.split(".")[-1], f, g) for n, (f, g) in self.flips.items())


def _group_of(name):
    for m in MODULES:
        if name.startswith(m + "."):
            return m
    return name.split(".")[0] + ("." + name.split(".")[1] if name.startswith("proposal.") else "")


def check_grads_per_parameter(ours, ref, rtol=GRAD_RTOL, floor=1e-2, label=""):
    """Every parameter: |g_ours - g_ref|_2 <= rtol * max(|g_ref|_2, floor * max |g|_2 over its sub-module)."""
    go = {n: p.grad for n, p in ours.named_parameters() if p.grad is not None}
    gr = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    assert set(go) == set(gr), "parameters with a gradient differ: %s" % sorted(set(go) ^ set(gr))
    groups = {}
    for n in gr:
        groups.setdefault(_group_of(n), []).append(n)
    rows = []
    for g, names in groups.items():
        gmax = max(float(gr[n].double().norm()) for n in names)
        for n in names:
            a, b = go[n].double(), gr[n].double()
            scale = max(float(b.norm()), floor * gmax, 1e-30)
            rows.append((float((a - b).norm()) / scale, n, float(b.norm()), gmax))
    rows.sort(reverse=True)
    head = "; ".join("%s %.2e" % (n, e) for e, n, _, _ in rows[:5])
    print("%sper-parameter gradient error (rel. L2), worst five: %s" % (label, head))
    bad = [(n, e) for e, n, _, _ in rows if not e < rtol]
    assert not bad, "gradients beyond %g (relative L2 per parameter): %s" % (rtol, bad[:12])
    return rows[0][0]


def checkpoint_path(name):
    """A shipped VoteNet checkpoint (git-ignored copy under baseline/_ref/pretrained, see baseline/install_ref.py)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = os.path.join(root, "baseline", "_ref", "pretrained", name, "model.pth")
    return p if os.path.exists(p) else None
