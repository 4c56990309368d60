"""Frame plan: the whole DLASeg forward (a1-a5) as a fixed list of C-ABI launches over a
pre-allocated NHWC arena, optionally frozen into a CUDA graph.

Built once per (input shape, device, weights version); `run()` is ~130 back-to-back kernel launches
with no host synchronisation, no allocation and no Python tensor ops in between.  Weight packing
(BN folding, im2col-major weight layout) happens here, once.
Reference structure being executed: dla.py:400-411 (DLA.forward), 271-284 (Tree), 693-699 (IDAUp),
728-735 (DLAUp), 789-802 (DLASeg.img2feats), base_model.py:111-132 (heads).
"""
from __future__ import annotations

import torch

from . import _lib, ops, packing
from ._lib import int_array, lib, ptr_array
from .ops import NHWC, FLAG_RELU


class FramePlan:
    def __init__(self, sd, heads, N, H, W, device, flags=0, u8_norm=None):
        """u8_norm = (mean[3], std[3]): the plan's input slot is a uint8 HWC frame [N,H,W,3] and the normalisation
        (x/255 - mean)/std of detector.py:408-409 is fused into the stem's load; None: the
        reference's pre-normalised fp32 NCHW image."""
        if H % 32 or W % 32:
            raise ValueError(f"DLASeg input must be a multiple of 32 (got {H}x{W}); pad like detector.py:370-371 does")
        self.sd = {k: v.detach().to(device=device, dtype=torch.float32) if v.is_floating_point() else v for k, v in sd.items()}
        self.heads = dict(heads)
        self.N, self.H, self.W, self.device, self.flags = N, H, W, device, flags
        self.calls = []          # (cfunc, args-without-stream)
        self.keep = []           # packed weights / buffers kept alive
        self.graph = None
        self.graph_pdl = None
        self.u8_norm = None if u8_norm is None else (tuple(float(v) for v in u8_norm[0]), tuple(float(v) for v in u8_norm[1]))
        if self.u8_norm is None:
            self.x_in = torch.empty((N, 3, H, W), dtype=torch.float32, device=device)
        else:
            self.x_in = torch.empty((N, H, W, 3), dtype=torch.uint8, device=device)
        self._dcn_ws = None
        self._dcn_ws_bytes = 0
        self._build()

    # ------------------------------------------------------------------ recording helpers
    def _rec(self, name, *args):
        self.calls.append((getattr(lib(), name), args, name))

    def _buf(self, H, W, C):
        b = NHWC.empty(self.N, H, W, C, self.device)
        self.keep.append(b)
        return b

    def _conv(self, srcs, wkey, bnkey, stride=1, pad=None, relu=True, residual=None):
        w = self.sd[wkey]
        co, _, kh, kw = w.shape
        pad = (kh // 2) if pad is None else pad
        wp, ldw = packing.pack_conv(w)
        scale, shift = packing.fold_bn(self.sd, bnkey)
        s0 = srcs[0]
        out = self._buf((s0.H + 2 * pad - kh) // stride + 1, (s0.W + 2 * pad - kw) // stride + 1, co)
        M = self.N * out.H * out.W
        w_tc, tc_bn = (None, 0)
        if self.flags & ops.FLAG_TENSORCORE:
            w_tc, tc_bn = ops.conv_tc_pack(wp, ldw, sum(s.C for s in srcs), kh, kw, M)
            self._dcn_ws_bytes = max(self._dcn_ws_bytes, lib().deft_tc_scratch_bytes(ldw, wp.shape[0], M, ops.KIND_CONV))
        self.keep += [wp, scale, shift, w_tc]
        self._rec("deft_conv2d_bn_act", ptr_array([s.ptr for s in srcs]), int_array([s.C for s in srcs]),
                  int_array([s.pitch for s in srcs]), len(srcs), self.N, s0.H, s0.W, wp.data_ptr(), ldw,
                  None if w_tc is None else w_tc.data_ptr(), tc_bn, scale.data_ptr(), shift.data_ptr(),
                  None if residual is None else residual.ptr, 0 if residual is None else residual.pitch,
                  out.ptr, out.pitch, co, kh, kw, stride, pad, self.flags | (FLAG_RELU if relu else 0), "DCN_WS", "DCN_WS_BYTES")
        return out

    def _maxpool(self, a):
        out = self._buf(a.H // 2, a.W // 2, a.C)
        self._rec("deft_maxpool2x2", a.ptr, a.pitch, out.ptr, out.pitch, self.N, a.H, a.W, a.C)
        return out

    def _deform(self, p, a):
        """DeformConv (dla.py:646-665): DCN + folded BN + ReLU."""
        pk = packing.pack_dcn(self.sd, p + ".conv", p + ".actf.0")
        tc = ops.dcn_tc_blobs(pk, self.N * a.H * a.W) if self.flags & ops.FLAG_TENSORCORE else (None, 0, None, 0)
        tcp = lambda t: None if t is None else t.data_ptr()
        self.keep.append(pk)
        out = self._buf(a.H, a.W, pk["cout"])
        need = lib().deft_dcnv2_workspace_bytes(self.N, a.H, a.W, pk["cin"], pk["cout"])
        self._dcn_ws_bytes = max(self._dcn_ws_bytes, need)
        self._rec("deft_dcnv2_forward", a.ptr, a.pitch, self.N, a.H, a.W, pk["cin"], pk["w_om"].data_ptr(), pk["b_om"].data_ptr(),
                  tcp(tc[2]), tc[3], pk["w"].data_ptr(), pk["ldw"], tcp(tc[0]), tc[1], pk["scale"].data_ptr(), pk["shift"].data_ptr(),
                  out.ptr, out.pitch, pk["cout"], self.flags | FLAG_RELU, "DCN_WS", "DCN_WS_BYTES")
        return out

    def _up_add(self, p, a, skip):
        wp, f = packing.pack_deconv(self.sd[p + ".weight"])
        self.keep.append(wp)
        out = self._buf(a.H * f, a.W * f, a.C)
        self._rec("deft_upsample_add", a.ptr, a.pitch, wp.data_ptr(), skip.ptr, skip.pitch, out.ptr, out.pitch, self.N, a.H, a.W, a.C, f)
        return out

    # ------------------------------------------------------------------ network structure
    def _block(self, p, x, residual, stride):
        t = self._conv([x], p + ".conv1.weight", p + ".bn1", stride=stride, relu=True)
        return self._conv([t], p + ".conv2.weight", p + ".bn2", relu=True, residual=residual)

    def _tree1(self, p, x, stride, level_root, children):
        """levels == 1 Tree (dla.py:271-284)."""
        children = list(children)
        bottom = self._maxpool(x) if stride > 1 else x
        if level_root:
            children.append(bottom)
        if (p + ".project.0.weight") in self.sd:
            residual = self._conv([bottom], p + ".project.0.weight", p + ".project.1", relu=False)
        else:
            residual = bottom
        x1 = self._block(p + ".tree1", x, residual, stride)
        x2 = self._block(p + ".tree2", x1, x1, 1)
        return self._conv([x2, x1] + children, p + ".root.conv.weight", p + ".root.bn", relu=True)

    def _tree2(self, p, x):
        """levels == 2, stride 2, level_root Tree (level3/level4).  The outer `project` is computed by the
        reference and discarded (every callee recomputes its residual), so it is skipped here."""
        bottom = self._maxpool(x)
        x1 = self._tree1(p + ".tree1", x, 2, False, [])
        return self._tree1(p + ".tree2", x1, 1, False, [bottom, x1])

    def _ida(self, p, layers, startp, endp):
        for i in range(startp + 1, endp):
            j = i - startp
            t = self._deform(f"{p}.proj_{j}", layers[i])
            u = self._up_add(f"{p}.up_{j}", t, layers[i - 1])
            layers[i] = self._deform(f"{p}.node_{j}", u)

    def _build(self):
        sd, N, H, W = self.sd, self.N, self.H, self.W
        # stem (7x7 on the NCHW image)
        wp, _ = packing.pack_conv(sd["base.base_layer.0.weight"], 16)
        sc, sh = packing.fold_bn(sd, "base.base_layer.1")
        self.keep += [wp, sc, sh]
        stem = self._buf(H, W, 16)
        if self.u8_norm is None:
            self._rec("deft_stem_conv7x7", self.x_in.data_ptr(), N, H, W, wp.data_ptr(), sc.data_ptr(), sh.data_ptr(), stem.ptr, stem.pitch)
        else:
            mean_a, std_a = _lib.float_array(self.u8_norm[0]), _lib.float_array(self.u8_norm[1])
            self.keep += [mean_a, std_a]
            self._rec("deft_stem_conv7x7_u8", self.x_in.data_ptr(), N, H, W, mean_a, std_a, wp.data_ptr(), sc.data_ptr(), sh.data_ptr(), stem.ptr,
                      stem.pitch, None)
        l0 = self._conv([stem], "base.level0.0.weight", "base.level0.1")
        l1 = self._conv([l0], "base.level1.0.weight", "base.level1.1", stride=2)
        l2 = self._tree1("base.level2", l1, 2, False, [])
        l3 = self._tree2("base.level3", l2)
        l4 = self._tree2("base.level4", l3)
        l5 = self._tree1("base.level5", l4, 2, True, [])
        base = [l0, l1, l2, l3, l4, l5]
        # DLAUp (dla.py:728-735), startp = 2
        layers = list(base)
        ups = [layers[-1]]
        for i in range(len(layers) - 2 - 1):
            self._ida(f"dla_up.ida_{i}", layers, len(layers) - i - 2, len(layers))
            ups.insert(0, layers[-1])
        # ida_up over clones of ups[0..2] (dla.py:796-799); y[0] is never written, so it aliases ups[0]
        y = [ups[0], ups[1], ups[2]]
        self._ida("ida_up", y, 0, 3)
        self.feature_maps = base + ups + y
        # heads
        hp = packing.pack_heads(sd, self.heads)
        feat = y[-1]
        hh, hw = feat.H, feat.W
        fused = bool(self.flags & ops.FLAG_TENSORCORE) and max(hp["head_c"]) <= 16
        if fused:     # one tcgen05 kernel: 3x3 conv + bias + ReLU + the 1x1 conv; the hidden map never reaches HBM
            hp["w1_tc"], _ = ops.tc_pack(hp["w1"], hp["w1"].shape[1], hp["w1"].shape[0], N * hh * hw, ops.KIND_HEADS, taps=9)
        self.keep.append(hp)
        self.head_out = {h: torch.empty((N, c, hh, hw), dtype=torch.float32, device=self.device) for h, c in self.heads.items()}
        ws_bytes = 256 if fused else lib().deft_heads_workspace_bytes(N, hh, hw, len(self.heads), hp["head_conv"])
        self.heads_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        self._rec("deft_heads_forward", feat.ptr, feat.pitch, N, hh, hw, feat.C, hp["w1"].data_ptr(),
                  hp["w1_tc"].data_ptr() if "w1_tc" in hp else None, hp["b1"].data_ptr(),
                  hp["head_conv"], hp["w2"].data_ptr(), hp["b2"].data_ptr(), int_array(hp["head_c"]), len(self.heads),
                  ptr_array([self.head_out[h].data_ptr() for h in hp["names"]]), self.flags, self.heads_ws.data_ptr(), ws_bytes)
        # shared DCN workspace (layers run back to back on one stream)
        self._dcn_ws = torch.empty(max(self._dcn_ws_bytes, 256), dtype=torch.uint8, device=self.device)
        fixed = []
        for fn, args, name in self.calls:
            args = tuple(self._dcn_ws.data_ptr() if a == "DCN_WS" else (self._dcn_ws.numel() if a == "DCN_WS_BYTES" else a)
                         for a in args)
            fixed.append((fn, args, name))
        self.calls = fixed
        self.fm_views = [m.nchw_view() for m in self.feature_maps]

    # ------------------------------------------------------------------ execution
    def launch(self):
        with _lib.on_device(self.device) as s:          # the plan's device current; s = its current stream
            for fn, args, name in self.calls:
                rc = fn(*args, s)
                if rc != 0:
                    _lib.check(rc, name)

    def launch_profiled(self):
        """Eager launch with a CUDA-event pair around every C-ABI call (on the launching stream).
        Returns [(entry point, milliseconds)], used by bench.py for the per-kernel roofline."""
        evs = []
        with _lib.on_device(self.device) as s:
            for fn, args, name in self.calls:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*args, s)
                e1.record()
                if rc != 0:
                    _lib.check(rc, name)
                evs.append((name, args, e0, e1))
        torch.cuda.synchronize(self.device)
        return [(name, args, e0.elapsed_time(e1)) for name, args, e0, e1 in evs]

    def run(self, x):
        if x.data_ptr() != self.x_in.data_ptr():
            self.x_in.copy_(x, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.launch()
        return self.head_out, self.fm_views

    def capture(self):
        """Freeze the launch list into a CUDA graph (after one eager warm-up so that lazy
        cudaFuncSetAttribute calls are done)."""
        with _lib.on_device(self.device):
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self.launch()
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.launch()
            self.graph = g
            self.graph_pdl = bool(_lib.lib().deft_get_pdl())     # the PDL mode the launches were captured under
        return self

    @property
    def n_launches(self):
        """Kernel launches per frame (DCN = 2 kernels, heads = 2 kernels)."""
        n = 0
        for _, _, name in self.calls:
            n += 2 if name in ("deft_dcnv2_forward", "deft_heads_forward") else 1
        return n
