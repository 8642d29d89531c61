"""Teacher-forced top-down decoder recurrence on the cluster kernels of libs2c.so (csrc/caption.cu), as one
autograd Function: ONE launch for the T forward steps, ONE for the T backward steps, then the weight gradients as
GEMMs over the (T*B)-row stacks the backward kernel emitted.

Arithmetic = the per-word step of the reference (models/caption_module.py:250-292) inside the loop of
forward_sample_batch (:428-500); see TopDownSceneCaptionModule._forward_sample_batch for what is hoisted out of it.
"""
import ctypes

import torch
from torch.autograd import Function

from .._lib import CaptionParams, call
from .pointnet2._ext import _guard, _stream


DEBUG_TS = None  # set to a list to collect (T, 8) globaltimer stamps of every forward call
CAPTURE = None   # test hook (tests/parity_utils.py): a list that receives the rectified u / lang stacks (T,B,E) of each call
# Kernel choice (fixed policy, both are libs2c kernels): the forward recurrence runs on the persistent cooperative grid
# (csrc/caption_grid.cu, taken by the library when B <= 8 and a barrier counter is passed), the backward recurrence on
# the cluster kernel (csrc/caption.cu): it measured faster there (1.33 vs 1.64 ms at B=8, T=26) -- its per-word chain
# has more dependent stages, and 16 SMs streaming from L2 beat 128 SMs paying a grid barrier per stage.  The tests
# flip these two constants to cover the other two kernels.
USE_GRID = True
USE_GRID_BWD = False


def supported(pre_word, mapped, obj):
    E, H, F = pre_word.shape[2], mapped.shape[2], obj.shape[2]
    return (pre_word.is_cuda and pre_word.dtype == torch.float32 and E % 4 == 0 and F % 4 == 0 and H % 64 == 0
            and E >= 4 and F >= 4 and mapped.shape[1] >= 1)


def require_supported(pre_word, mapped, obj):
    if not supported(pre_word, mapped, obj):
        raise RuntimeError("caption decoder kernels: unsupported input (emb %d, hidden %d, feat %d): need CUDA fp32, emb / "
                           "feat multiples of 4, hidden a multiple of 64 (a shape whose shared-memory plan does not fit "
                           "is reported by the library itself)" % (pre_word.shape[2], mapped.shape[2], obj.shape[2]))


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class _TopDownDecode(Function):
    @staticmethod
    def forward(ctx, pre_word, pre_tgt, mapped, obj, valid, w_tdh, w_ih1, w_hh1, b_ih1, b_hh1, w_hidd, w_att, w_lang,
                b_lang, w_ih2, w_hh2, b_ih2, b_hh2):
        ctx.set_materialize_grads(False)
        B, T, E = pre_word.shape
        K, H = mapped.shape[1], mapped.shape[2]
        F = obj.shape[2]
        dev = pre_word.device
        pre_word, pre_tgt, mapped, obj, valid = _c(pre_word), _c(pre_tgt), _c(mapped), _c(obj), _c(valid)
        if w_tdh.stride(1) != 1 or w_tdh.stride(0) % 4 != 0 or w_tdh.data_ptr() % 16 != 0:
            w_tdh = w_tdh.contiguous()
        ws = [_c(w) for w in (w_ih1, w_hh1, b_ih1, b_hh1, w_hidd, w_att, w_lang, b_lang, w_ih2, w_hh2, b_ih2, b_hh2)]
        # per-step tensors, (T,B,.): one allocation, sliced
        widths = dict(u=E, h1=H, r1=H, z1=H, n1=H, hn1=H, q=H, probs=K, att=F, lang=E, r2=H, z2=H, n2=H, hn2=H, h2=H)
        buf = torch.empty((T * B * sum(widths.values()),), dtype=torch.float32, device=dev)
        saved, off = {}, 0
        for name, w in widths.items():
            saved[name] = buf[off:off + T * B * w].view(T, B, w)
            off += T * B * w
        P = CaptionParams()
        P.B, P.T, P.K, P.E, P.H, P.F = B, T, K, E, H, F
        P.ld_tdh = w_tdh.stride(0)
        P.pre_word, P.pre_tgt, P.mapped, P.obj, P.valid = (t.data_ptr() for t in (pre_word, pre_tgt, mapped, obj, valid))
        P.w_tdh = w_tdh.data_ptr()
        for name, w in zip("w_ih1 w_hh1 b_ih1 b_hh1 w_hidd w_att w_lang b_lang w_ih2 w_hh2 b_ih2 b_hh2".split(), ws):
            setattr(P, name, w.data_ptr())
        for name, t in saved.items():
            setattr(P, name, t.data_ptr())
        scores = torch.empty((T, B, K), dtype=torch.float32, device=dev)  # scratch of the forward kernel
        P.scores = scores.data_ptr()
        bar = torch.zeros((1,), dtype=torch.int32, device=dev)  # grid-barrier counter of the persistent-grid variant
        P.grid_bar = bar.data_ptr() if USE_GRID else None
        if DEBUG_TS is not None:  # profiling aid (tools/caption_probe.py): per-word stage time stamps
            DEBUG_TS.append(torch.zeros((T, 8), dtype=torch.int64, device=dev))
            P.dbg_ts = DEBUG_TS[-1].data_ptr()
        with _guard(pre_word):
            call("s2c_caption_decode_fwd", ctypes.byref(P), _stream(pre_word))
        if CAPTURE is not None:
            CAPTURE.append({"u": saved["u"], "lang": saved["lang"]})
        ctx.save_for_backward(pre_word, pre_tgt, mapped, obj, valid, w_tdh, *ws, buf)
        ctx.dims = (B, T, K, E, H, F)
        ctx.widths = widths
        hiddens = saved["h2"].transpose(0, 1)      # (B,T,H) view
        attn = saved["probs"].permute(1, 2, 0)     # (B,K,T) view
        return hiddens, attn

    @staticmethod
    def backward(ctx, d_hiddens, d_attn):
        B, T, K, E, H, F = ctx.dims
        (pre_word, pre_tgt, mapped, obj, valid, w_tdh, w_ih1, w_hh1, b_ih1, b_hh1, w_hidd, w_att, w_lang, b_lang,
         w_ih2, w_hh2, b_ih2, b_hh2, buf) = ctx.saved_tensors
        dev = buf.device
        saved, off = {}, 0
        for name, w in ctx.widths.items():
            saved[name] = buf[off:off + T * B * w].view(T, B, w)
            off += T * B * w
        d_h2 = (torch.zeros((T, B, H), dtype=torch.float32, device=dev) if d_hiddens is None
                else d_hiddens.transpose(0, 1).contiguous())
        d_probs = None if d_attn is None else d_attn.permute(2, 0, 1).contiguous()
        wt = dict(wt_tdh=w_tdh.t().contiguous(), wt_ih1=w_ih1.t().contiguous(), wt_hh1=w_hh1.t().contiguous(),
                  wt_hidd=w_hidd.t().contiguous(), wt_lang=w_lang.t().contiguous(), wt_ih2=w_ih2.t().contiguous(),
                  wt_hh2=w_hh2.t().contiguous())
        gw = dict(dgi2=3 * H, dgh2=3 * H, dlang=E, datt=F, dq=H, dgi1=3 * H, dgh1=3 * H, du=E)
        gbuf = torch.empty((T * B * sum(gw.values()),), dtype=torch.float32, device=dev)
        g, off = {}, 0
        for name, w in gw.items():
            g[name] = gbuf[off:off + T * B * w].view(T * B, w)
            off += T * B * w
        d_mapped = torch.zeros_like(mapped)
        d_obj = torch.zeros_like(obj)
        d_watt = torch.zeros(((B + 7) // 8, H), dtype=torch.float32, device=dev)
        P = CaptionParams()
        P.B, P.T, P.K, P.E, P.H, P.F = B, T, K, E, H, F
        P.ld_tdh = w_tdh.stride(0)
        P.mapped, P.obj, P.valid, P.w_att = mapped.data_ptr(), obj.data_ptr(), valid.data_ptr(), w_att.data_ptr()
        for name, t in saved.items():
            setattr(P, name, t.data_ptr())
        for name, t in wt.items():
            setattr(P, name, t.data_ptr())
        for name, t in g.items():
            setattr(P, name, t.data_ptr())
        P.d_h2 = d_h2.data_ptr()
        P.d_probs = d_probs.data_ptr() if d_probs is not None else None
        P.d_mapped, P.d_obj, P.d_watt = d_mapped.data_ptr(), d_obj.data_ptr(), d_watt.data_ptr()
        bar = torch.zeros((1,), dtype=torch.int32, device=dev)
        P.grid_bar = bar.data_ptr() if (USE_GRID and USE_GRID_BWD) else None
        with _guard(buf):
            call("s2c_caption_decode_bwd", ctypes.byref(P), _stream(buf))

        # weight gradients: GEMMs over the (T*B)-row stacks
        def rows(name):
            return saved[name].reshape(T * B, -1)

        def prev(name):  # the state each step started from: zeros, then the previous step's output
            h = saved[name]
            return torch.cat([torch.zeros_like(h[:1]), h[:-1]], 0).reshape(T * B, -1)
        h1p, h2p = prev("h1"), prev("h2")
        # outputs of the 7 short-reduction GEMMs (+ bias gradients as column sums) in one buffer
        shapes = dict(w_ih2=(3 * H, E), w_hh2=(3 * H, H), w_lang=(E, F + H), w_hidd=(H, H), w_ih1=(3 * H, E),
                      w_hh1=(3 * H, H), w_tdh=(E, H), b_ih2=(3 * H,), b_hh2=(3 * H,), b_lang=(E,), b_ih1=(3 * H,),
                      b_hh1=(3 * H,), b_td=(E,))
        sizes = {k: (v[0] * v[1] if len(v) == 2 else v[0]) for k, v in shapes.items()}
        wbuf = torch.empty((sum(sizes.values()),), dtype=torch.float32, device=dev)
        d, off = {}, 0
        for k, v in shapes.items():
            d[k] = wbuf[off:off + sizes[k]].view(*v)
            off += sizes[k]
        lang_in = torch.cat([rows("att"), rows("h1")], 1)

        def gemm_tn(A, X, out, colsum=None):
            call("s2c_gemm_tn", A.data_ptr(), A.stride(0), X.data_ptr(), X.stride(0), A.shape[0], A.shape[1], X.shape[1],
                 out.data_ptr(), out.stride(0), colsum.data_ptr() if colsum is not None else None, _stream(buf))
        with _guard(buf):
            gemm_tn(g["dgi2"], rows("lang"), d["w_ih2"], d["b_ih2"])
            gemm_tn(g["dgh2"], h2p, d["w_hh2"], d["b_hh2"])
            gemm_tn(g["dlang"], lang_in, d["w_lang"], d["b_lang"])
            gemm_tn(g["dq"], rows("h1"), d["w_hidd"])
            gemm_tn(g["dgi1"], rows("u"), d["w_ih1"], d["b_ih1"])
            gemm_tn(g["dgh1"], h1p, d["w_hh1"], d["b_hh1"])
            gemm_tn(g["du"], h2p, d["w_tdh"], d["b_td"])          # b_td = sum over (t, b) of du: also d pre_tgt's total
        du3 = g["du"].view(T, B, E)
        return (du3.transpose(0, 1), du3.sum(0), d_mapped, d_obj, None, d["w_tdh"],
                d["w_ih1"], d["w_hh1"], d["b_ih1"], d["b_hh1"], d["w_hidd"], d_watt.sum(0).view_as(w_att),
                d["w_lang"], d["b_lang"], d["w_ih2"], d["w_hh2"], d["b_ih2"], d["b_hh2"])


def topdown_decode(pre_word, pre_tgt, mapped, obj, valid, w_tdh, cell1, map_hidd, attend, map_lang, cell2):
    """-> hiddens (B,T,H) [h2 after every word], attn (B,K,T) [softmax over the valid proposals].
    cell1 / cell2: nn.GRUCell; map_hidd / attend: bias-free nn.Linear; map_lang: nn.Linear (its ReLU is applied here)."""
    return _TopDownDecode.apply(pre_word, pre_tgt, mapped, obj, valid, w_tdh,
                                cell1.weight_ih, cell1.weight_hh, cell1.bias_ih, cell1.bias_hh,
                                map_hidd.weight, attend.weight, map_lang.weight, map_lang.bias,
                                cell2.weight_ih, cell2.weight_hh, cell2.bias_ih, cell2.bias_hh)
