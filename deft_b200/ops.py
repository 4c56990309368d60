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


def _device_of(args):
    for a in args:
        if isinstance(a, torch.Tensor):
            return a.device
        if isinstance(a, NHWC):
            return a.t.device
        if isinstance(a, (list, tuple)) and a:
            d = _device_of(a)
            if d is not None:
                return d
    return None


def _on_input_device(fn):
    """Run the wrapped C-ABI call with the device of its first tensor argument current, so that the stream handed to
    the library (stream_ptr()), the kernels' per-device attributes and the TMA descriptors all belong to the device
    that owns the pointers -- a model on cuda:1 works while cuda:0 is current, like a torch nn.Module does."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kw):
        d = _device_of(args) or _device_of(tuple(kw.values()))
        if d is None or d.type != "cuda" or d.index is None or d.index == torch.cuda.current_device():
            return fn(*args, **kw)
        with _lib.on_device(d):
            return fn(*args, **kw)
    return wrapper


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


@_on_input_device
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


@_on_input_device
def to_nchw(a: NHWC):
    out = torch.empty((a.N, a.C, a.H, a.W), dtype=torch.float32, device=a.t.device)
    check(lib().deft_nhwc_to_nchw(a.ptr, a.pitch, out.data_ptr(), a.N, a.C, a.H, a.W, stream_ptr()), "nhwc_to_nchw")
    return out


def _p(t):
    return None if t is None else t.data_ptr()


KIND_CONV, KIND_DCN, KIND_HEADS = 0, 1, 2


@_on_input_device
def tc_pack(wp, ldw, K, M, kind=KIND_CONV, taps=1):
    """Packed weight Wp[K][ldw] -> (tcgen05 tile-image blob, tile N) for an M-row problem; one kernel launch,
    done at load time.  `taps` = KH*KW of the conv that will consume it (its K order is channel-block-major when
    the channel count allows).  The returned bn must accompany the blob in every call."""
    bn = lib().deft_tc_tile_n(ldw, M, kind)
    n = lib().deft_tc_blob_bytes(ldw, K, bn)
    blob = torch.empty(n // 4, dtype=torch.float32, device=wp.device)
    check(lib().deft_tc_pack_weights(wp.data_ptr(), ldw, K, bn, taps, blob.data_ptr(), stream_ptr()), "tc_pack_weights")
    return blob, bn


def conv_tc_pack(wp, ldw, cin_total, kh, kw, M):
    """tcgen05 weight image of a conv whose packed weight is wp [(ky,kx,ci)][ldw].  The 16-channel 3x3 layers use the
    tap-major pseudo-channel producer (TsC16AP): their image is packed from a copy zero-padded to K = 192 with taps = 1."""
    if cin_total == 16 and kh == 3 and kw == 3:
        pad = torch.zeros((192, wp.shape[1]), dtype=wp.dtype, device=wp.device)
        pad[: wp.shape[0]] = wp
        return tc_pack(pad, ldw, 192, M, taps=1)
    return tc_pack(wp, ldw, wp.shape[0], M, taps=kh * kw)


def tc_scratch(ldw, K, M, kind, device):
    n = lib().deft_tc_scratch_bytes(ldw, K, M, kind)
    return torch.empty(n, dtype=torch.uint8, device=device) if n else None


@_on_input_device
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
        w_tc, tc_bn = conv_tc_pack(wp, ldw, sum(s.C for s in srcs), kh, kw, M)
        if ws is None:
            ws = tc_scratch(ldw, wp.shape[0], M, KIND_CONV, s0.t.device)
    check(lib().deft_conv2d_bn_act(ptr_array([s.ptr for s in srcs]), int_array([s.C for s in srcs]),
                                   int_array([s.pitch for s in srcs]), len(srcs), s0.N, s0.H, s0.W, wp.data_ptr(), ldw,
                                   _p(w_tc), tc_bn, _p(scale), _p(shift), None if residual is None else residual.ptr,
                                   0 if residual is None else residual.pitch, out.ptr, out.pitch, cout, kh, kw, stride, pad,
                                   fl, _p(ws), 0 if ws is None else ws.numel(), stream_ptr()), "conv2d_bn_act")
    return out


@_on_input_device
def stem_conv7x7(x_nchw, wp, scale, shift, out=None):
    N, _, H, W = x_nchw.shape
    if out is None:
        out = NHWC.empty(N, H, W, 16, x_nchw.device)
    check(lib().deft_stem_conv7x7(x_nchw.data_ptr(), N, H, W, wp.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.ptr,
                                  out.pitch, stream_ptr()), "stem_conv7x7")
    return out


@_on_input_device
def maxpool2x2(a: NHWC, out=None):
    if out is None:
        out = NHWC.empty(a.N, a.H // 2, a.W // 2, a.C, a.t.device)
    check(lib().deft_maxpool2x2(a.ptr, a.pitch, out.ptr, out.pitch, a.N, a.H, a.W, a.C, stream_ptr()), "maxpool2x2")
    return out


def dcn_workspace(N, H, W, cin, cout, device):
    n = lib().deft_dcnv2_workspace_bytes(N, H, W, cin, cout)
    return torch.empty(n, dtype=torch.uint8, device=device)


@_on_input_device
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


@_on_input_device
def upsample_add(a: NHWC, wp, f, skip=None, out=None):
    if out is None:
        out = NHWC.empty(a.N, a.H * f, a.W * f, a.C, a.t.device)
    check(lib().deft_upsample_add(a.ptr, a.pitch, wp.data_ptr(), None if skip is None else skip.ptr,
                                  0 if skip is None else skip.pitch, out.ptr, out.pitch, a.N, a.H, a.W, a.C, f, stream_ptr()),
          "upsample_add")
    return out


@_on_input_device
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


@_on_input_device
def sigmoid_output_(hm, dep=None, depth_scale=1.0):
    check(lib().deft_sigmoid_output(_p(hm), 0 if hm is None else hm.numel(), _p(dep), 0 if dep is None else dep.numel(),
                                    float(depth_scale), stream_ptr()), "sigmoid_output")


@_on_input_device
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


@_on_input_device
def gather_feat(feat, inds):
    feat = feat.contiguous()
    B, Cc, H, W = feat.shape
    K = inds.shape[1]
    out = torch.empty((B, K, Cc), dtype=torch.float32, device=feat.device)
    check(lib().deft_gather_feat(feat.data_ptr(), B, Cc, H * W, inds.data_ptr(), K, out.data_ptr(), stream_ptr()), "gather_feat")
    return out


@_on_input_device
def boxes_to_centres(boxes, W, H, out=None):
    """boxes [N,4] device (x1,y1,x2,y2) in a WxH frame -> [N,2] centres in [-1,1] (image.py:391-412)."""
    N = boxes.shape[0]
    if out is None:
        out = torch.empty((N, 2), dtype=torch.float32, device=boxes.device)
    check(lib().deft_boxes_to_centres(boxes.data_ptr(), N, C.c_float(W), C.c_float(H), out.data_ptr(), stream_ptr()), "boxes_to_centres")
    return out


@_on_input_device
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


@_on_input_device
def afe_affinity(xp, counts, xn, m, P, pk, out=None, ws=None, flags=0):
    """xp [F,P,Cemb] device slab, counts: list[int] (len F), xn [>=m, Cemb]; -> out [F,P,P+1]."""
    F = xp.shape[0]
    cemb = pk["cemb"]
    dev = xp.device
    if ws is None:
        ws = torch.empty(lib().deft_afe_affinity_workspace_bytes(F, P, cemb), dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.zeros((F, P, P + 1), dtype=torch.float32, device=dev)
    tc, tcbn = _affinity_tc(pk, F, P, flags)
    check(lib().deft_afe_affinity(xp.data_ptr(), int_array(list(counts)), F, xn.data_ptr(), m, P, cemb, pk["w1a"].data_ptr(),
                                  pk["w1b"].data_ptr(), pk["sh1"].data_ptr(), pk["w2"].data_ptr(), pk["sc2"].data_ptr(),
                                  pk["sh2"].data_ptr(), pk["w3"].data_ptr(), pk["sc3"].data_ptr(), pk["sh3"].data_ptr(),
                                  pk["w4"].data_ptr(), pk["b4"].data_ptr(), pk["w5"].data_ptr(), C.c_float(pk["b5"]), tc, tcbn,
                                  out.data_ptr(), flags, ws.data_ptr(), ws.numel(), stream_ptr()), "afe_affinity")
    return out


@_on_input_device
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


# ---------------------------------------------------------------------------------------------- a8 cached / a9 / f1 / f2
def _affinity_tc(pk, F, P, flags):
    """tcgen05 tile images of the five MLP weights (cached in the pack dict) as ctypes arrays, or (None, None)."""
    if not (flags & FLAG_TENSORCORE):
        return None, None
    cache = pk.setdefault("_tc", {})
    if (F, P) not in cache:       # any valid tile works for any M; pick tiles for the nominal (full) problem sizes
        nominal = {"w1a": F * P, "w1b": P, "w2": F * P * P, "w3": F * P * P, "w4": F * P * P}
        cache[(F, P)] = [tc_pack(pk[k], pk[k].shape[1], pk[k].shape[0], nominal[k]) for k in ("w1a", "w1b", "w2", "w3", "w4")]
    blobs = cache[(F, P)]
    return ptr_array([b.data_ptr() for b, _ in blobs]), int_array([bn for _, bn in blobs])


def _mlp_args(pk):
    return (pk["w2"].data_ptr(), pk["sc2"].data_ptr(), pk["sh2"].data_ptr(), pk["w3"].data_ptr(), pk["sc3"].data_ptr(), pk["sh3"].data_ptr(),
            pk["w4"].data_ptr(), pk["b4"].data_ptr(), pk["w5"].data_ptr(), C.c_float(pk["b5"]))


@_on_input_device
def afe_project(emb, pk, au=None, bv=None, want_au=True, want_bv=True, flags=0, tcF=1, tcP=None):
    """emb [rows, Cemb] -> (Au' [rows,512] with the folded bias/BN shift, Bv' [rows,512]) of the decomposed first MLP layer."""
    rows = emb.shape[0]
    dev = emb.device
    if want_au and au is None:
        au = torch.empty((rows, 512), dtype=torch.float32, device=dev)
    if want_bv and bv is None:
        bv = torch.empty((rows, 512), dtype=torch.float32, device=dev)
    tc, tcbn = _affinity_tc(pk, tcF, tcP or rows, flags)
    check(lib().deft_afe_project(emb.data_ptr(), rows, pk["cemb"], pk["w1a"].data_ptr(), pk["w1b"].data_ptr(), pk["sh1"].data_ptr(), tc, tcbn,
                                 _p(au) if want_au else None, _p(bv) if want_bv else None, flags, stream_ptr()), "afe_project")
    return au, bv


@_on_input_device
def afe_affinity_cached(au_mem, n_dev, bv, m_max, P, pk, rows=None, m_dev=None, out=None, ws=None, flags=0):
    """au_mem [F,P,512] cached Au' of the memory ring, n_dev int32[F] live rows (device), bv [>=m_max,512] -> out [F,P,P+1]."""
    F = au_mem.shape[0]
    rows = P if rows is None else rows
    dev = au_mem.device
    need = lib().deft_afe_pairs_workspace_bytes(F, rows, m_max)
    if ws is None:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.empty((F, P, P + 1), dtype=torch.float32, device=dev)
    tc, tcbn = _affinity_tc(pk, F, P, flags)
    check(lib().deft_afe_affinity_cached(au_mem.data_ptr(), F, P, rows, n_dev.data_ptr(), bv.data_ptr(), m_max, _p(m_dev), *_mlp_args(pk), tc, tcbn,
                                         out.data_ptr(), flags, ws.data_ptr(), ws.numel(), stream_ptr()), "afe_affinity_cached")
    return out


@_on_input_device
def afe_forward_scores(au, bv, pk, false_constant=1.0, out=None, ws=None, flags=0):
    """au, bv [B,P,512] -> [B,1,P+1,P+1] (AFE.forward after the extractor + add_unmatched_dim)."""
    B, P, _ = au.shape
    dev = au.device
    if ws is None:
        ws = torch.empty(lib().deft_afe_pairs_workspace_bytes(B, P, P), dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.empty((B, 1, P + 1, P + 1), dtype=torch.float32, device=dev)
    tc, tcbn = _affinity_tc(pk, B, P, flags)
    check(lib().deft_afe_forward_scores(au.data_ptr(), bv.data_ptr(), B, P, *_mlp_args(pk), tc, tcbn, C.c_float(false_constant), out.data_ptr(), flags,
                                        ws.data_ptr(), ws.numel(), stream_ptr()), "afe_forward_scores")
    return out


@_on_input_device
def afe_loss(inp, target, mask0, mask1):
    """AFE_module.loss forward; returns (out7 float32[7] device, indexes_pre int64 [B,1,P] device)."""
    B, _, Q, _ = inp.shape
    P = Q - 1
    dev = inp.device
    out7 = torch.empty(7, dtype=torch.float32, device=dev)
    idx = torch.empty((B, 1, P), dtype=torch.int64, device=dev)
    ws = torch.empty(lib().deft_afe_loss_workspace_bytes(B, P), dtype=torch.uint8, device=dev)
    check(lib().deft_afe_loss(inp.data_ptr(), target.data_ptr(), mask0.data_ptr(), mask1.data_ptr(), B, P, out7.data_ptr(), idx.data_ptr(),
                              ws.data_ptr(), ws.numel(), stream_ptr()), "afe_loss")
    return out7, idx


@_on_input_device
def memory_push(emb, au, rows, C_, F, P, clock, slot_frame, counts, memory, au_mem, m_dev=None):
    check(lib().deft_memory_push(emb.data_ptr(), _p(au), rows, _p(m_dev), C_, F, P, clock.data_ptr(), slot_frame.data_ptr(), counts.data_ptr(),
                                 memory.data_ptr(), _p(au_mem), stream_ptr()), "memory_push")


@_on_input_device
def track_similarity(aff, clock, slot_frame, counts, node_frame, node_id, node_cnt, m, mm, m_frame, sim, cost=None, m_dev=None, max_track_node=50,
                     decay=1.0, decay2=0.01):
    F, P, _ = aff.shape
    T = node_cnt.shape[0]
    check(lib().deft_track_similarity(aff.data_ptr(), F, P, clock.data_ptr(), slot_frame.data_ptr(), counts.data_ptr(), node_frame.data_ptr(),
                                      node_id.data_ptr(), node_cnt.data_ptr(), T, m, _p(m_dev), mm, m_frame, max_track_node, C.c_double(decay),
                                      C.c_double(decay2), sim.data_ptr(), _p(cost), stream_ptr()), "track_similarity")
    return sim, cost


@_on_input_device
def track_append(match, clock, node_frame, node_id, node_cnt):
    check(lib().deft_track_append(match.data_ptr(), clock.data_ptr(), node_cnt.shape[0], node_frame.data_ptr(), node_id.data_ptr(),
                                  node_cnt.data_ptr(), stream_ptr()), "track_append")


@_on_input_device
def count_above(scores, thresh, out):
    check(lib().deft_count_above(scores.data_ptr(), scores.numel(), C.c_float(thresh), out.data_ptr(), stream_ptr()), "count_above")
    return out


@_on_input_device
def stem_conv7x7_u8(img_u8, mean, std, wp, scale, shift, out=None, x_nchw=None):
    """img_u8 [N,H,W,3] uint8 (device) -> stem output NHWC [N,H,W,16]; normalisation fused (detector.py:408-409)."""
    N, H, W, _ = img_u8.shape
    if out is None:
        out = NHWC.empty(N, H, W, 16, img_u8.device)
    check(lib().deft_stem_conv7x7_u8(img_u8.data_ptr(), N, H, W, _lib.float_array(mean), _lib.float_array(std), wp.data_ptr(), scale.data_ptr(),
                                     shift.data_ptr(), out.ptr, out.pitch, _p(x_nchw), stream_ptr()), "stem_conv7x7_u8")
    return out


@_on_input_device
def preprocess_warp(img_u8, inv_affine, mean, std, H, W, out=None):
    """img_u8 [h,w,3] uint8 device; inv_affine: 6 floats mapping network-input (x,y) -> source (x,y) -> fp32 [1,3,H,W]."""
    sh, sw, _ = img_u8.shape
    if out is None:
        out = torch.empty((1, 3, H, W), dtype=torch.float32, device=img_u8.device)
    check(lib().deft_preprocess_warp(img_u8.data_ptr(), sh, sw, _lib.float_array(inv_affine), _lib.float_array(mean), _lib.float_array(std),
                                     out.data_ptr(), H, W, stream_ptr()), "preprocess_warp")
    return out


@_on_input_device
def dcnv2_backward(a: NHWC, pk, grad_y: NHWC):
    """Backward of the bare DCN module (no BN).  a: input NHWC, grad_y: NHWC [N,H,W,Cout] -> (grad_x NHWC, grad_w [9Cin][ldw],
    grad_b [Cout], grad_w_om [9Cin][32], grad_b_om [32]) in the forward's packed layouts."""
    dev = a.t.device
    cin, cout, ldw = pk["cin"], pk["cout"], pk["ldw"]
    gx = NHWC.empty(a.N, a.H, a.W, cin, dev)
    gw = torch.empty((9 * cin, ldw), dtype=torch.float32, device=dev)
    gb = torch.empty(cout, dtype=torch.float32, device=dev)
    gwo = torch.empty((9 * cin, 32), dtype=torch.float32, device=dev)
    gbo = torch.empty(32, dtype=torch.float32, device=dev)
    ws = torch.empty(lib().deft_dcnv2_backward_workspace_bytes(a.N, a.H, a.W, cin, cout), dtype=torch.uint8, device=dev)
    check(lib().deft_dcnv2_backward(a.ptr, a.pitch, a.N, a.H, a.W, cin, pk["w_om"].data_ptr(), pk["b_om"].data_ptr(), pk["w"].data_ptr(), ldw, cout,
                                    grad_y.ptr, grad_y.pitch, gx.ptr, gx.pitch, gwo.data_ptr(), gbo.data_ptr(), gw.data_ptr(), gb.data_ptr(),
                                    ws.data_ptr(), ws.numel(), stream_ptr()), "dcnv2_backward")
    return gx, gw, gb, gwo, gbo


@_on_input_device
def fuse_and_assign(cost, m, cost_limit, lambda_=0.9, pred_xy=None, n_obs=None, det_xy=None, m_dev=None, match_track=None, match_det=None, ws=None):
    """cost [T, ldc>=m] fp32 device -> (match_track int32 [T], match_det int32 [m]); the reference's fuse_motion (LSTM
    branch) + linear_assignment(thresh=cost_limit), exact."""
    T, ldc = cost.shape
    dev = cost.device
    if match_track is None:
        match_track = torch.empty(T, dtype=torch.int32, device=dev)
    if match_det is None:
        match_det = torch.empty(m, dtype=torch.int32, device=dev)
    if ws is None:
        ws = torch.empty(lib().deft_assign_workspace_bytes(T, m), dtype=torch.uint8, device=dev)
    check(lib().deft_fuse_and_assign(cost.data_ptr(), ldc, T, m, _p(m_dev), _p(pred_xy), _p(n_obs), _p(det_xy), C.c_double(lambda_),
                                     C.c_double(cost_limit), match_track.data_ptr(), match_det.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()),
          "fuse_and_assign")
    return match_track, match_det
