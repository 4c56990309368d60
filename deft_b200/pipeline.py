"""Per-stream frame pipeline: the public "one call per frame" API of deft_b200.

Chains, on one CUDA stream with no host synchronisation in between, exactly what the reference does
per frame between `images.to(device)` and the tracker's cost matrix:

    DLASeg.forward -> _sigmoid_output -> generic_decode(K) -> convert_detection -> AFE extractor (N=K)
      -> affinity of the K new embeddings against the F-frame embedding memory (FeatureRecorder, tracker.py:46-90)
      [-> LSTM motion step for T tracks when configured]

and keeps the embedding memory (a ring of F frames x P rows) resident in HBM.  One instance per video
stream / GPU; streams are independent (no collective), which is how the path shards across GPUs.  Within a
stream, consecutive frames are software-pipelined two deep (dense part of frame t+1 over the tail of frame t).
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .model.decode import _REG_HEADS

class FramePipeline:
    """Software pipelining across the frames of ONE video stream.  `depth` = activation arenas = frames in flight,
    `dense_streams` = CUDA streams the dense parts (DLASeg: backbone + heads) rotate over.
      depth=1                   : everything of a frame on the caller's stream (the reference's order).
      depth=2, dense_streams=1  : the small-grid tail of frame t (sigmoid, decode, embedding extractor, affinity,
                                  LSTM) runs on a second stream while frame t+1's dense part runs on the caller's.
      depth=3, dense_streams=2  : (default for depth >= 3: dense_streams = depth - 1) additionally two dense parts overlap (the backbone of frame t+1 does not depend on
                                  frame t), which fills the SMs the tile tails and small layers of one frame leave idle.
    The tails always run in frame order on one stream and are the only users of the embedding memory, so every mode
    returns bit-identical results; a frame's results are complete once the device has been synchronised
    (`synchronize()`), or after `join()` for work queued behind it on the caller's stream."""

    def __init__(self, model, opt, height, width, mem_frames=5, lstm=None, lstm_tracks=0, device=None, use_cuda_graph=True, depth=3,
                 dense_streams=None):
        self.model, self.opt = model, opt
        self.device = device or next(model.parameters()).device
        self.H, self.W = height, width
        self.K = opt.K
        self.P = opt.max_object
        if self.K > self.P:
            raise ValueError("K must be <= max_object (forward_stacker_features contract, AFE.py:110-122)")
        if dense_streams is None:
            dense_streams = max(1, depth - 1)
        if not (1 <= depth <= 4 and 1 <= dense_streams <= depth):
            raise ValueError("need 1 <= dense_streams <= depth <= 4")
        self.depth = depth
        self.S = dense_streams
        # PDL shortens an isolated frame but early-launched CTAs hold SMs the other frames in flight could use
        # (measured at depth 4: 497 fps with, 522 fps without); the plans below are captured with this setting
        _lib.lib().deft_set_pdl(1 if depth == 1 else 0)
        self.F = mem_frames
        self.cemb = sum(model.AFE.selector_output_channel)
        self.lstm, self.T = lstm, lstm_tracks
        dev = self.device
        probe = torch.empty((1, 3, height, width), dtype=torch.float32, device=dev)
        self.memory = torch.zeros((self.F, self.P, self.cemb), dtype=torch.float32, device=dev)  # embedding ring
        self.counts = [self.K] * self.F          # live rows per memory frame (steady state: K detections/frame)
        self.head = 0
        self.aff = [torch.zeros((self.F, self.P, self.P + 1), dtype=torch.float32, device=dev) for _ in range(depth)]
        self.aff_ws = torch.empty(_lib.lib().deft_afe_affinity_workspace_bytes(self.F, self.P, self.cemb), dtype=torch.uint8, device=dev)
        self.ext_ws = torch.empty(_lib.lib().deft_afe_extract_workspace_bytes(self.K), dtype=torch.uint8, device=dev)
        self.dec_ws = None
        if lstm is not None:
            fin = lstm.model.fin
            self.lstm_h = torch.zeros((self.T, 128), device=dev)
            self.lstm_c = torch.zeros((self.T, 128), device=dev)
            self.lstm_x = torch.zeros((self.T, fin), device=dev)
        self.plans = [model.plan_for(probe, slot=i) for i in range(depth)]
        self.plan = self.plans[0]
        self.xs = [p.x_in for p in self.plans]    # the plans' static input slots: no extra device copy per frame
        self.x = self.xs[0]
        self.tail_graph = [None] * depth
        self.out = [None] * depth
        self.t = 0                                # frames submitted
        self.tail_stream = torch.cuda.Stream(device=dev) if depth > 1 else None
        self.dense_streams = [torch.cuda.Stream(device=dev) for _ in range(self.S)] if self.S > 1 else None
        self._ev_submit = [torch.cuda.Event() for _ in range(depth)]
        self._last_dense = None
        self._ev_dense = [torch.cuda.Event() for _ in range(depth)]
        self._ev_tail = [torch.cuda.Event() for _ in range(depth)]
        self._copy_stream, self._stage_idx = None, 0
        self.use_cuda_graph = use_cuda_graph

    # ------------------------------------------------------------------------------------------
    def _tail(self, slot):
        """Everything after DLASeg.forward for the frame whose activations sit in arena `slot`."""
        plan = self.plans[slot]
        heads = plan.head_out
        hm = heads["hm"]
        ops.sigmoid_output_(hm, heads.get("dep"), getattr(self.opt, "depth_scale", 1.0))
        if self.dec_ws is None:
            B, C, h, w = hm.shape
            self.dec_ws = torch.empty(_lib.lib().deft_decode_workspace_bytes(B, C, h, w, self.K), dtype=torch.uint8, device=self.device)
        d = ops.decode(hm, self.K, reg=heads.get("reg"), wh=heads.get("wh"), ltrb_amodal=heads.get("ltrb_amodal"), ws=self.dec_ws)
        res = {"scores": d["scores"], "clses": d["clses"], "bboxes": d["bboxes"], "xs": d["xs"], "ys": d["ys"]}
        for h in _REG_HEADS:
            if h in heads:
                res[h] = ops.gather_feat(heads[h], d["inds"])
        oh, ow = hm.shape[2], hm.shape[3]
        centres = ops.boxes_to_centres(d["bboxes"].view(-1, 4), float(ow), float(oh))
        pk = self.model.AFE.packed()
        emb = ops.afe_extract(plan.feature_maps, pk["sel_w"], pk["sel_b"], pk["sel_co"], centres, self.model.AFE.align_corners,
                              ws=self.ext_ws)
        ops.afe_affinity(self.memory, self.counts, emb, self.K, self.P, pk["mlp"], out=self.aff[slot], ws=self.aff_ws, flags=self.model.AFE._flags)
        res["embeddings"] = emb
        res["affinity"] = self.aff[slot]
        if self.lstm is not None:
            hn, cn, pred = self.lstm.predict_batch(self.lstm_h, self.lstm_c, self.lstm_x)
            self.lstm_h.copy_(hn)
            self.lstm_c.copy_(cn)
            res["lstm_pred"] = pred
        return res

    def _body(self, slot=0):
        """One whole frame, eagerly, on the current stream (image already in self.xs[slot])."""
        self.plans[slot].run(self.xs[slot])
        return self._tail(slot)

    def _push_memory(self, emb):
        self.memory[self.head, : self.K].copy_(emb)
        self.head = (self.head + 1) % self.F

    def warmup(self):
        """One eager frame per arena, then (use_cuda_graph) the tails are frozen into CUDA graphs; the dense part is
        the plan's own graph."""
        res = None
        for slot in range(self.depth):
            if slot > 0:
                self.xs[slot].copy_(self.xs[0])
            r = self._body(slot)
            res = res or r
        torch.cuda.synchronize(self.device)
        if self.use_cuda_graph and self.tail_graph[0] is None:
            for slot in range(self.depth):
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    self._tail(slot)
                torch.cuda.current_stream(self.device).wait_stream(side)
                with torch.cuda.graph(g):
                    self.out[slot] = self._tail(slot)
                self.tail_graph[slot] = g
            torch.cuda.synchronize(self.device)
        return res

    def _run_tail(self, slot):
        if self.tail_graph[slot] is not None:
            self.tail_graph[slot].replay()
            res = self.out[slot]
        else:
            res = self._tail(slot)
        self._push_memory(res["embeddings"])
        return res

    def step_device(self, image_dev=None):
        """One frame with the image already in HBM (copied into the next arena's input slot if given).  Returns the
        frame's result tensors (device; complete after synchronize(), or join() + stream order)."""
        slot = self.t % self.depth
        cur = torch.cuda.current_stream(self.device)
        ds = cur
        if self.dense_streams is not None:
            ds = self.dense_streams[self.t % self.S]
            self._ev_submit[slot].record(cur)             # the image (and whatever produced it) is ordered before the frame
        self.t += 1
        self._last_dense = ds
        x = self.xs[slot]
        with torch.cuda.stream(ds):
            if ds is not cur:
                ds.wait_event(self._ev_submit[slot])
            if self.depth > 1:
                ds.wait_event(self._ev_tail[slot])        # the frame that last used this arena has finished its tail
            if image_dev is not None and image_dev.data_ptr() != x.data_ptr():
                x.copy_(image_dev, non_blocking=True)
            self.plans[slot].run(x)
            if self.depth == 1:
                return self._run_tail(slot)
            self._ev_dense[slot].record(ds)
        with torch.cuda.stream(self.tail_stream):
            self.tail_stream.wait_event(self._ev_dense[slot])
            res = self._run_tail(slot)
            self._ev_tail[slot].record(self.tail_stream)
        return res

    def join(self):
        """Make the caller's stream wait (on the GPU, not the host) for every submitted frame's tail."""
        cur = torch.cuda.current_stream(self.device)
        for st in (self.dense_streams or []):
            cur.wait_stream(st)
        if self.tail_stream is not None:
            cur.wait_stream(self.tail_stream)

    def synchronize(self):
        """Block the host until every submitted frame is complete."""
        torch.cuda.current_stream(self.device).synchronize()
        for st in (self.dense_streams or []):
            st.synchronize()
        if self.tail_stream is not None:
            self.tail_stream.synchronize()

    def step(self, image_host, out_host=None):
        """End-to-end frame through the public API: pinned host image in, host results out
        (H2D of the frame, D2H of detections + affinity -- the copies the reference makes at
        detector.py:150,545-547 and AFE.py:154-158)."""
        # The upload runs on its own stream into a staging ring (one slot per frame in flight), so the H2D of frame t+1 overlaps the
        # kernels of frame t whenever the caller does not synchronise between steps (results land in pinned
        # memory asynchronously; synchronise the device / call synchronize() before reading them).
        cur = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            ns = max(2, self.depth)                              # staging ring: one slot per frame in flight
            self._stage = [torch.empty_like(self.x) for _ in range(ns)]
            self._stage_free = [torch.cuda.Event() for _ in range(ns)]
            self._stage_ready = [torch.cuda.Event() for _ in range(ns)]
            for e in self._stage_free:
                e.record(cur)
        j = self._stage_idx
        self._stage_idx = (j + 1) % len(self._stage)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free[j])
            self._stage[j].copy_(image_host, non_blocking=True)
            self._stage_ready[j].record(self._copy_stream)
        cur.wait_event(self._stage_ready[j])
        res = self.step_device(self._stage[j])                 # device-to-device copy into the arena's slot, ~3 us
        self._stage_free[j].record(self._last_dense)           # ... on the stream that ran it
        if out_host is None:
            out_host = {k: torch.empty(res[k].shape, dtype=res[k].dtype, pin_memory=True) for k in ("scores", "clses", "bboxes", "affinity")}
        with torch.cuda.stream(self.tail_stream if self.tail_stream is not None else cur):
            for k, v in out_host.items():
                v.copy_(res[k], non_blocking=True)
        return out_host

    def count_launches(self):
        """Kernel launches of one frame (our kernels only): the library counts every launch it makes
        (deft_launch_count); one eager frame is run between two readings."""
        torch.cuda.synchronize(self.device)
        before = _lib.lib().deft_launch_count()
        self.plans[0].launch()
        self._tail(0)
        torch.cuda.synchronize(self.device)
        return int(_lib.lib().deft_launch_count() - before)
