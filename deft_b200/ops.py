"""Tensor-level wrappers over the C ABI (device pointers in, device tensors out).

PyTorch is plumbing here: it owns device memory and the current stream.  Every function requires
CUDA tensors and raises otherwise (no CPU path).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, int_array, lib, ptr_array, stream_ptr

FLAG_RELU, FLAG_TENSORCORE, FLAG_TF32X1 = _lib.FLAG_RELU, _lib.FLAG_TENSORCORE, _lib.FLAG_TF32X1


class NHWC:
    """A channels-last activation: memory [N,H,W,pitch] fp32, logical channels C <= pitch."""
    __slots__ = ("t", "N", "H", "W", "C", "pitch", "ptr")

    def __init__(self, t, C=None):
        assert t.dim() == 4 and t.dtype == torch.float32 and t.is_contiguous()
        self.t = t
        self.N, self.H, self.W, self.pitch = t.shape
        self.C = self.pitch if C is None else C
        self.ptr = t.data_ptr()

    @staticmethod
    def empty(N, H, W, C, device):
        return NHWC(torch.empty((N, H, W, C), dtype=torch.float32, device=device))

    def nchw_view(self):
        """Zero-copy [N,C,H,W]-shaped (channels_last-strided) view, what DLASeg returns."""
        return self.t[..., : self.C].permute(0, 3, 1, 2)


def as_nhwc(x):
    """Accept a [N,C,H,W] tensor in either memory format; returns an NHWC handle (zero-copy when the
    tensor already is channels-last, otherwise one deft_nchw_to_nhwc launch)."""
    _lib.require_device(x)
    assert x.dim() == 4 and x.dtype == torch.float32
    N, Cc, H, W = x.shape
    st = x.stride()
    if (W > 1 and H > 1 and (st[1] == 1 or Cc == 1) and st[3] % 4 == 0 and st[3] >= Cc and st[2] == W * st[3]
            and (N == 1 or st[0] == H * W * st[3]) and x.data_ptr() % 16 == 0):
        h = NHWC.__new__(NHWC)
        h.t, h.N, h.H, h.W, h.C, h.pitch, h.ptr = x, N, H, W, Cc, st[3], x.data_ptr()
        return h
    xc = x.contiguous()
    out = NHWC.empty(N, H, W, Cc, x.device)
    check(lib().deft_nchw_to_nhwc(xc.data_ptr(), out.ptr, out.pitch, N, Cc, H, W, stream_ptr()), "nchw_to_nhwc")
    return out


def to_nchw(a: NHWC):
    out = torch.empty((a.N, a.C, a.H, a.W), dtype=torch.float32, device=a.t.device)
    check(lib().deft_nhwc_to_nchw(a.ptr, a.pitch, out.data_ptr(), a.N, a.C, a.H, a.W, stream_ptr()), "nhwc_to_nchw")
    return out


def _p(t):
    return None if t is None else t.data_ptr()


KIND_CONV, KIND_DCN, KIND_HEADS = 0, 1, 2


def tc_pack(wp, ldw, K, M, kind=KIND_CONV, taps=1):
    """Packed weight Wp[K][ldw] -> (tcgen05 tile-image blob, tile N) for an M-row problem; one kernel launch,
    done at load time.  `taps` = KH*KW of the conv that will consume it (its K order is channel-block-major when
    the channel count allows).  The returned bn must accompany the blob in every call."""
    bn = lib().deft_tc_tile_n(ldw, M, kind)
    n = lib().deft_tc_blob_bytes(ldw, K, bn)
    blob = torch.empty(n // 4, dtype=torch.float32, device=wp.device)
    check(lib().deft_tc_pack_weights(wp.data_ptr(), ldw, K, bn, taps, blob.data_ptr(), stream_ptr()), "tc_pack_weights")
    return blob, bn


def tc_scratch(ldw, K, M, kind, device):
    n = lib().deft_tc_scratch_bytes(ldw, K, M, kind)
    return torch.empty(n, dtype=torch.uint8, device=device) if n else None


def conv2d_bn_act(srcs, wp, ldw, scale, shift, cout, kh, kw, stride, pad, relu, residual=None, out=None, flags=0, w_tc=None, tc_bn=0,
                  ws=None):
    s0 = srcs[0]
    OH = (s0.H + 2 * pad - kh) // stride + 1
    OW = (s0.W + 2 * pad - kw) // stride + 1
    if out is None:
        out = NHWC.empty(s0.N, OH, OW, cout, s0.t.device)
    fl = flags | (FLAG_RELU if relu else 0)
    if (flags & FLAG_TENSORCORE) and w_tc is None:
        M = s0.N * OH * OW
        w_tc, tc_bn = tc_pack(wp, ldw, wp.shape[0], M, taps=kh * kw)
        if ws is None:
            ws = tc_scratch(ldw, wp.shape[0], M, KIND_CONV, s0.t.device)
    check(lib().deft_conv2d_bn_act(ptr_array([s.ptr for s in srcs]), int_array([s.C for s in srcs]),
                                   int_array([s.pitch for s in srcs]), len(srcs), s0.N, s0.H, s0.W, wp.data_ptr(), ldw,
                                   _p(w_tc), tc_bn, _p(scale), _p(shift), None if residual is None else residual.ptr,
                                   0 if residual is None else residual.pitch, out.ptr, out.pitch, cout, kh, kw, stride, pad,
                                   fl, _p(ws), 0 if ws is None else ws.numel(), stream_ptr()), "conv2d_bn_act")
    return out


def stem_conv7x7(x_nchw, wp, scale, shift, out=None):
    N, _, H, W = x_nchw.shape
    if out is None:
        out = NHWC.empty(N, H, W, 16, x_nchw.device)
    check(lib().deft_stem_conv7x7(x_nchw.data_ptr(), N, H, W, wp.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.ptr,
                                  out.pitch, stream_ptr()), "stem_conv7x7")
    return out


def maxpool2x2(a: NHWC, out=None):
    if out is None:
        out = NHWC.empty(a.N, a.H // 2, a.W // 2, a.C, a.t.device)
    check(lib().deft_maxpool2x2(a.ptr, a.pitch, out.ptr, out.pitch, a.N, a.H, a.W, a.C, stream_ptr()), "maxpool2x2")
    return out


def dcn_workspace(N, H, W, cin, cout, device):
    n = lib().deft_dcnv2_workspace_bytes(N, H, W, cin, cout)
    return torch.empty(n, dtype=torch.uint8, device=device)


def dcnv2_forward(a: NHWC, pk, relu, ws=None, out=None, flags=0):
    if ws is None:
        ws = dcn_workspace(a.N, a.H, a.W, pk["cin"], pk["cout"], a.t.device)
    if out is None:
        out = NHWC.empty(a.N, a.H, a.W, pk["cout"], a.t.device)
    fl = flags | (FLAG_RELU if relu else 0)
    tc = dcn_tc_blobs(pk, a.N * a.H * a.W) if (flags & FLAG_TENSORCORE) else (None, 0, None, 0)
    check(lib().deft_dcnv2_forward(a.ptr, a.pitch, a.N, a.H, a.W, pk["cin"], pk["w_om"].data_ptr(), pk["b_om"].data_ptr(),
                                   _p(tc[2]), tc[3], pk["w"].data_ptr(), pk["ldw"], _p(tc[0]), tc[1], pk["scale"].data_ptr(),
                                   pk["shift"].data_ptr(), out.ptr, out.pitch, pk["cout"], fl, ws.data_ptr(), ws.numel(), stream_ptr()),
          "dcnv2_forward")
    return out


def dcn_tc_blobs(pk, M):
    """(w_tc, tc_bn, w_om_tc, om_bn) for an M-pixel problem, packed once and cached in the pack dict."""
    cache = pk.setdefault("_tc", {})
    if M not in cache:
        w_tc, bn = tc_pack(pk["w"], pk["ldw"], 9 * pk["cin"], M, KIND_DCN, taps=9)
        w_om_tc, om_bn = tc_pack(pk["w_om"], 32, 9 * pk["cin"], M, KIND_CONV, taps=9)
        cache[M] = (w_tc, bn, w_om_tc, om_bn)
    return cache[M]


def upsample_add(a: NHWC, wp, f, skip=None, out=None):
    if out is None:
        out = NHWC.empty(a.N, a.H * f, a.W * f, a.C, a.t.device)
    check(lib().deft_upsample_add(a.ptr, a.pitch, wp.data_ptr(), None if skip is None else skip.ptr,
                                  0 if skip is None else skip.pitch, out.ptr, out.pitch, a.N, a.H, a.W, a.C, f, stream_ptr()),
          "upsample_add")
    return out


def heads_forward(feat: NHWC, hp, outs=None, ws=None, flags=0):
    nh = len(hp["head_c"])
    if ws is None:
        ws = torch.empty(lib().deft_heads_workspace_bytes(feat.N, feat.H, feat.W, nh, hp["head_conv"]), dtype=torch.uint8,
                         device=feat.t.device)
    if outs is None:
        outs = [torch.empty((feat.N, c, feat.H, feat.W), dtype=torch.float32, device=feat.t.device) for c in hp["head_c"]]
    if (flags & FLAG_TENSORCORE) and "w1_tc" not in hp:
        hp["w1_tc"], _ = tc_pack(hp["w1"], hp["w1"].shape[1], hp["w1"].shape[0], feat.N * feat.H * feat.W, KIND_HEADS, taps=9)
    check(lib().deft_heads_forward(feat.ptr, feat.pitch, feat.N, feat.H, feat.W, feat.C, hp["w1"].data_ptr(), _p(hp.get("w1_tc")), hp["b1"].data_ptr(),
                                   hp["head_conv"], hp["w2"].data_ptr(), hp["b2"].data_ptr(), int_array(hp["head_c"]), nh,
                                   ptr_array([o.data_ptr() for o in outs]), flags, ws.data_ptr(), ws.numel(), stream_ptr()),
          "heads_forward")
    return outs


def sigmoid_output_(hm, dep=None, depth_scale=1.0):
    check(lib().deft_sigmoid_output(_p(hm), 0 if hm is None else hm.numel(), _p(dep), 0 if dep is None else dep.numel(),
                                    float(depth_scale), stream_ptr()), "sigmoid_output")


def decode(hm, K, reg=None, wh=None, ltrb_amodal=None, ws=None):
    _lib.require_device(hm)
    hm = hm.contiguous()
    B, Cc, H, W = hm.shape
    dev = hm.device
    if ws is None:
        ws = torch.empty(lib().deft_decode_workspace_bytes(B, Cc, H, W, K), dtype=torch.uint8, device=dev)
    scores = torch.empty((B, K), dtype=torch.float32, device=dev)
    inds = torch.empty((B, K), dtype=torch.int64, device=dev)
    clses = torch.empty((B, K), dtype=torch.int32, device=dev)
    ys = torch.empty((B, K), dtype=torch.float32, device=dev)
    xs = torch.empty((B, K), dtype=torch.float32, device=dev)
    bboxes = torch.empty((B, K, 4), dtype=torch.float32, device=dev) if (wh is not None or ltrb_amodal is not None) else None
    bba = torch.empty((B, K, 4), dtype=torch.float32, device=dev) if ltrb_amodal is not None else None
    cont = lambda t: None if t is None else t.contiguous()
    reg, wh, ltrb_amodal = cont(reg), cont(wh), cont(ltrb_amodal)
    check(lib().deft_decode(hm.data_ptr(), B, Cc, H, W, K, _p(reg), _p(wh), _p(ltrb_amodal), scores.data_ptr(), inds.data_ptr(),
                            clses.data_ptr(), ys.data_ptr(), xs.data_ptr(), _p(bboxes), _p(bba), ws.data_ptr(), ws.numel(),
                            stream_ptr()), "decode")
    return dict(scores=scores, inds=inds, clses=clses, ys=ys, xs=xs, bboxes=bboxes, bboxes_amodal=bba)


def gather_feat(feat, inds):
    feat = feat.contiguous()
    B, Cc, H, W = feat.shape
    K = inds.shape[1]
    out = torch.empty((B, K, Cc), dtype=torch.float32, device=feat.device)
    check(lib().deft_gather_feat(feat.data_ptr(), B, Cc, H * W, inds.data_ptr(), K, out.data_ptr(), stream_ptr()), "gather_feat")
    return out


def boxes_to_centres(boxes, W, H, out=None):
    """boxes [N,4] device (x1,y1,x2,y2) in a WxH frame -> [N,2] centres in [-1,1] (image.py:391-412)."""
    N = boxes.shape[0]
    if out is None:
        out = torch.empty((N, 2), dtype=torch.float32, device=boxes.device)
    check(lib().deft_boxes_to_centres(boxes.data_ptr(), N, C.c_float(W), C.c_float(H), out.data_ptr(), stream_ptr()), "boxes_to_centres")
    return out


def afe_extract(maps, sel_w, sel_b, sel_co, centres, align_corners=False, out=None, ws=None):
    """maps: 13 NHWC handles (batch 1); centres [N,2] device tensor -> emb [N, Cemb]."""
    N = centres.shape[0]
    cemb = sum(sel_co)
    if out is None:
        out = torch.empty((N, cemb), dtype=torch.float32, device=centres.device)
    if ws is None:
        ws = torch.empty(lib().deft_afe_extract_workspace_bytes(N), dtype=torch.uint8, device=centres.device)
    check(lib().deft_afe_extract(ptr_array([m.ptr for m in maps]), int_array([m.pitch for m in maps]),
                                 int_array([m.C for m in maps]), int_array([m.H for m in maps]), int_array([m.W for m in maps]),
                                 ptr_array([w.data_ptr() for w in sel_w]), ptr_array([b.data_ptr() for b in sel_b]),
                                 int_array(sel_co), centres.data_ptr(), N, out.data_ptr(), 1 if align_corners else 0,
                                 ws.data_ptr(), ws.numel(), stream_ptr()), "afe_extract")
    return out


def afe_affinity(xp, counts, xn, m, P, pk, out=None, ws=None, flags=0):
    """xp [F,P,Cemb] device slab, counts: list[int] (len F), xn [>=m, Cemb]; -> out [F,P,P+1]."""
    F = xp.shape[0]
    cemb = pk["cemb"]
    dev = xp.device
    if ws is None:
        ws = torch.empty(lib().deft_afe_affinity_workspace_bytes(F, P, cemb), dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.zeros((F, P, P + 1), dtype=torch.float32, device=dev)
    tc, tcbn = None, None
    if flags & FLAG_TENSORCORE:
        cache = pk.setdefault("_tc", {})
        if (F, P) not in cache:       # any valid tile works for any M; pick tiles for the nominal (full) problem sizes
            nominal = {"w1a": F * P, "w1b": P, "w2": F * P * P, "w3": F * P * P, "w4": F * P * P}
            cache[(F, P)] = [tc_pack(pk[k], pk[k].shape[1], pk[k].shape[0], nominal[k]) for k in ("w1a", "w1b", "w2", "w3", "w4")]
        blobs = cache[(F, P)]
        tc = ptr_array([b.data_ptr() for b, _ in blobs])
        tcbn = int_array([bn for _, bn in blobs])
    check(lib().deft_afe_affinity(xp.data_ptr(), int_array(list(counts)), F, xn.data_ptr(), m, P, cemb, pk["w1a"].data_ptr(),
                                  pk["w1b"].data_ptr(), pk["sh1"].data_ptr(), pk["w2"].data_ptr(), pk["sc2"].data_ptr(),
                                  pk["sh2"].data_ptr(), pk["w3"].data_ptr(), pk["sc3"].data_ptr(), pk["sh3"].data_ptr(),
                                  pk["w4"].data_ptr(), pk["b4"].data_ptr(), pk["w5"].data_ptr(), C.c_float(pk["b5"]), tc, tcbn,
                                  out.data_ptr(), flags, ws.data_ptr(), ws.numel(), stream_ptr()), "afe_affinity")
    return out


def lstm_step(x, h, c, w, nout, h_out=None, c_out=None, pred=None):
    T, F = x.shape
    dev = x.device
    h_out = torch.empty_like(h) if h_out is None else h_out
    c_out = torch.empty_like(c) if c_out is None else c_out
    pred = torch.empty((T, nout), dtype=torch.float32, device=dev) if pred is None else pred
    check(lib().deft_lstm_step(x.data_ptr(), h.data_ptr(), c.data_ptr(), w["w_ih"].data_ptr(), w["w_hh"].data_ptr(),
                               w["b_ih"].data_ptr(), w["b_hh"].data_ptr(), w["w1"].data_ptr(), w["b1"].data_ptr(),
                               w["w2"].data_ptr(), w["b2"].data_ptr(), T, F, nout, h_out.data_ptr(), c_out.data_ptr(),
                               pred.data_ptr(), stream_ptr()), "lstm_step")
    return h_out, c_out, pred
