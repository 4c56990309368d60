"""Per-stream frame pipeline: the public "one call per frame" API of deft_b200.

Chains, with no host synchronisation in between, exactly what the reference does per frame between
`images.to(device)` and the tracker's cost matrix:

    [uint8 frame -> normalise (fused into the stem)]                                      detector.py:408-409
    DLASeg.forward -> _sigmoid_output -> generic_decode(K) -> convert_detection -> AFE extractor (N=K)
      -> affinity of the K new embeddings against the F-frame embedding memory            tracker.py:46-90 (FeatureRecorder)
      -> [per-track decayed median over the last <= 4 nodes -> [tracks, dets] cost]       tracker.py:219-252, 886-894
      -> [LSTM motion step for T tracks when configured]
      -> recorder update (ring slot = frame % F)                                          tracker.py:59-74

The recorder lives on the device: embedding ring [F,P,Cemb], the cached first-layer products Au' of every memory frame
[F,P,512] (the pairwise MLP's first layer is linear in the concatenated pair, so the memory side is computed once
per frame, not once per frame per memory frame), per-slot frame index and live-row count, a frame clock, and the
per-track node lists.  One instance per video stream; streams are independent (no collective), which is how the path
shards across GPUs.  Within a stream, consecutive frames are software-pipelined `depth` deep.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .model.decode import _REG_HEADS

import os

M_FRAME = {"kitti_tracking": 5, "nuscenes": 3}       # tracker.py:77-81 (default 10)
# DEFT_NVTX=1: NVTX ranges around the two halves of every frame ("deft.dense[slot]", "deft.tail[slot]") for ncu --nvtx / timeline tools
_NVTX = os.environ.get("DEFT_NVTX", "0") not in ("", "0")


class _Range:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


class FramePipeline:
    """Software pipelining across the frames of ONE video stream.  `depth` = activation arenas = frames in flight,
    `dense_streams` = CUDA streams the dense parts (DLASeg: backbone + heads) rotate over.
      depth=1                   : everything of a frame on the caller's stream (the reference's order).
      depth=2, dense_streams=1  : the small-grid tail of frame t (sigmoid, decode, embedding extractor, affinity,
                                  LSTM) runs on a second stream while frame t+1's dense part runs on the caller's.
      depth>=3                  : (default dense_streams = depth - 1) additionally several dense parts overlap (the
                                  backbone of frame t+1 does not depend on frame t), which fills the SMs the tile
                                  tails and small layers of one frame leave idle.
    The tails always run in frame order on one stream and are the only users of the recorder state, so every mode
    returns bit-identical results; a frame's results are complete once the device has been synchronised
    (`synchronize()`), or after `join()` for work queued behind it on the caller's stream.

    mem_rows   rows of the pair grid per memory frame = the most detections a frame may put into the recorder
               (default K).  tracks > 0 adds the per-track similarity reduction (row f1) for that many track slots.
    uint8_norm (mean[3], std[3]): frames are uint8 HWC [1,H,W,3] and are normalised inside the stem (row f2)."""

    def __init__(self, model, opt, height, width, mem_frames=5, lstm=None, lstm_tracks=0, device=None, use_cuda_graph=True, depth=3,
                 dense_streams=None, mem_rows=None, tracks=0, uint8_norm=None, score_thresh=None, assign_thresh=None):
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
        self.F = mem_frames
        self.rows = min(self.P, self.K if mem_rows is None else mem_rows)
        self.T_tracks = tracks
        self.score_thresh = score_thresh
        self.assign_thresh = assign_thresh          # row f3: linear_assignment(cost, thresh) on the device + node append (tracker.py:915)
        if assign_thresh is not None and not tracks:
            raise ValueError("assign_thresh needs tracks > 0 (the assignment consumes the [tracks, dets] cost matrix)")
        self.uint8_norm = uint8_norm
        self.use_cuda_graph = use_cuda_graph
        self.cemb = sum(model.AFE.selector_output_channel)
        self.lstm, self.T = lstm, lstm_tracks
        dataset = getattr(opt, "dataset", "mot")
        self.mm = 2 if dataset == "nuscenes" else 4             # tracker.py:232-235
        self.m_frame = M_FRAME.get(dataset, 10)
        dev = self.device
        with _lib.on_device(dev):
            # PDL shortens an isolated frame but early-launched CTAs hold SMs the other frames in flight could use
            # (measured at depth 4: 497 fps with, 522 fps without); the plans below are captured with this setting
            _lib.lib().deft_set_pdl(1 if depth == 1 else 0)
            F, P = self.F, self.P
            # ---- recorder state (device-resident)
            self.memory = torch.zeros((F, P, self.cemb), dtype=torch.float32, device=dev)
            self.au_mem = torch.zeros((F, P, 512), dtype=torch.float32, device=dev)
            self.clock = torch.zeros(2, dtype=torch.int32, device=dev)
            self.slot_frame = torch.full((F,), -1, dtype=torch.int32, device=dev)
            self.counts = torch.zeros(F, dtype=torch.int32, device=dev)
            self.m_dev = torch.zeros(1, dtype=torch.int32, device=dev) if score_thresh is not None else None
            self.aff = [torch.zeros((F, P, P + 1), dtype=torch.float32, device=dev) for _ in range(depth)]
            self.aff_ws = torch.empty(_lib.lib().deft_afe_pairs_workspace_bytes(F, self.rows, self.K), dtype=torch.uint8, device=dev)
            self.ext_ws = torch.empty(_lib.lib().deft_afe_extract_workspace_bytes(self.K), dtype=torch.uint8, device=dev)
            self.dec_ws = None
            if tracks:
                L = _lib.lib().deft_track_nodes_per_track()
                self.node_frame = torch.full((tracks, L), -1, dtype=torch.int32, device=dev)
                self.node_id = torch.zeros((tracks, L), dtype=torch.int32, device=dev)
                self.node_cnt = torch.zeros(tracks, dtype=torch.int32, device=dev)
                self.sim = [torch.zeros((tracks, P + 1), dtype=torch.float32, device=dev) for _ in range(depth)]
                self.cost = [torch.zeros((tracks, P), dtype=torch.float32, device=dev) for _ in range(depth)]
                if assign_thresh is not None:
                    self.match_track = [torch.full((tracks,), -1, dtype=torch.int32, device=dev) for _ in range(depth)]
                    self.match_det = [torch.full((self.K,), -1, dtype=torch.int32, device=dev) for _ in range(depth)]
                    self.assign_ws = torch.empty(_lib.lib().deft_assign_workspace_bytes(tracks, self.K), dtype=torch.uint8, device=dev)
            if lstm is not None:
                fin = lstm.model.fin
                self.lstm_h = torch.zeros((self.T, 128), device=dev)
                self.lstm_c = torch.zeros((self.T, 128), device=dev)
                self.lstm_x = torch.zeros((self.T, fin), device=dev)
            self.plans = [model.plan_for_shape(1, height, width, dev, slot=i, u8_norm=uint8_norm) for i in range(depth)]
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

    # ------------------------------------------------------------------------------------------ recorder state
    def reset_memory(self):
        """Empty recorder, frame clock 0, no track nodes (a new video)."""
        self.memory.zero_(); self.au_mem.zero_(); self.clock.zero_(); self.counts.zero_(); self.slot_frame.fill_(-1)
        if self.T_tracks:
            self.node_frame.fill_(-1); self.node_id.zero_(); self.node_cnt.zero_()

    def prefill_memory(self, embeddings, rows=None):
        """Steady state for benchmarks / tests: F recorder frames (frame indices 0..F-1) holding `embeddings`
        [F, rows, Cemb]; the clock is left at F (the next frame to be associated)."""
        F, r, _ = embeddings.shape
        assert F == self.F and r <= self.P
        rows = r if rows is None else rows
        self.reset_memory()
        self.memory[:, :r].copy_(embeddings)
        pk = self.model.AFE.packed()["mlp"]
        au, _ = ops.afe_project(self.memory.view(F * self.P, self.cemb), pk, want_bv=False, flags=self.model.AFE._flags, tcF=F, tcP=self.P)
        self.au_mem.copy_(au.view(F, self.P, 512))
        self.counts.fill_(rows)
        self.slot_frame.copy_(torch.arange(F, dtype=torch.int32))
        self.clock[0] = F

    def set_tracks(self, node_frame, node_id, node_cnt):
        """Install per-track node lists (host or device int tensors [T,L], [T,L], [T]; oldest node first)."""
        self.node_frame.copy_(node_frame.to(torch.int32)); self.node_id.copy_(node_id.to(torch.int32)); self.node_cnt.copy_(node_cnt.to(torch.int32))

    def append_matches(self, match):
        """After association: match[t] = current-frame detection of track t or -1 (device or host int tensor).  Enqueued on
        the tail stream, i.e. ordered after the frame whose cost matrix produced it and before the next frame's reduction.
        Must be called BEFORE the next step (the node's frame index is the clock of the frame just associated, so the
        append reads clock - 1 through a shifted clock view)."""
        m = match.to(device=self.device, dtype=torch.int32)
        st = self.tail_stream if self.tail_stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.stream(st):
            self._prev_clock = getattr(self, "_prev_clock", None)
            if self._prev_clock is None:
                self._prev_clock = torch.zeros(1, dtype=torch.int32, device=self.device)
            torch.sub(self.clock[:1], 1, out=self._prev_clock)          # the recorder has already advanced past that frame
            ops.track_append(m, self._prev_clock, self.node_frame, self.node_id, self.node_cnt)

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
        if self.m_dev is not None:                  # live detections = scores above the threshold, counted on the device
            ops.count_above(d["scores"], self.score_thresh, self.m_dev)
            res["n_dets"] = self.m_dev
        oh, ow = hm.shape[2], hm.shape[3]
        centres = ops.boxes_to_centres(d["bboxes"].view(-1, 4), float(ow), float(oh))
        pk = self.model.AFE.packed()
        fl = self.model.AFE._flags
        emb = ops.afe_extract(plan.feature_maps, pk["sel_w"], pk["sel_b"], pk["sel_co"], centres, self.model.AFE.align_corners,
                              ws=self.ext_ws)
        # first MLP layer of the new frame: Bv' for this frame's affinity, Au' for when it is a memory frame
        au_new, bv = ops.afe_project(emb, pk["mlp"], flags=fl, tcF=self.F, tcP=self.P)
        ops.afe_affinity_cached(self.au_mem, self.counts, bv, self.K, self.P, pk["mlp"], rows=self.rows, m_dev=self.m_dev, out=self.aff[slot],
                                ws=self.aff_ws, flags=fl)
        res["embeddings"] = emb
        res["affinity"] = self.aff[slot]
        if self.T_tracks:
            ops.track_similarity(self.aff[slot], self.clock, self.slot_frame, self.counts, self.node_frame, self.node_id, self.node_cnt, self.K,
                                 self.mm, self.m_frame, self.sim[slot], self.cost[slot], m_dev=self.m_dev)
            res["similarity"] = self.sim[slot]
            res["cost"] = self.cost[slot]
            if self.assign_thresh is not None:
                # association on the device: appearance cost -> exact linear assignment -> the matched tracks get this frame's node
                # (motion fusion is available through ops.fuse_and_assign for callers that keep the tracks' predictions on the device)
                ops.fuse_and_assign(self.cost[slot], self.K, self.assign_thresh, m_dev=self.m_dev, match_track=self.match_track[slot],
                                    match_det=self.match_det[slot], ws=self.assign_ws)
                ops.track_append(self.match_track[slot], self.clock, self.node_frame, self.node_id, self.node_cnt)
                res["match_track"] = self.match_track[slot]
                res["match_det"] = self.match_det[slot]
        if self.lstm is not None:
            hn, cn, pred = self.lstm.predict_batch(self.lstm_h, self.lstm_c, self.lstm_x)
            self.lstm_h.copy_(hn)
            self.lstm_c.copy_(cn)
            res["lstm_pred"] = pred
        # FeatureRecorder.update: this frame becomes a memory frame (ring slot clock % F), the clock advances
        ops.memory_push(emb, au_new, min(self.K, self.rows), self.cemb, self.F, self.P, self.clock, self.slot_frame, self.counts, self.memory,
                        self.au_mem, m_dev=self.m_dev)
        return res

    def _body(self, slot=0):
        """One whole frame, eagerly, on the current stream (image already in self.xs[slot])."""
        self.plans[slot].run(self.xs[slot])
        return self._tail(slot)

    def warmup(self):
        """One eager frame per arena, then (use_cuda_graph) both halves of every arena are frozen into CUDA graphs: the
        dense plan (111 launches) and the tail.  The recorder state is restored afterwards: warm-up frames are not part
        of the stream."""
        with _lib.on_device(self.device):
            saved = [self.memory, self.au_mem, self.clock, self.slot_frame, self.counts]
            if self.T_tracks:
                saved += [self.node_frame, self.node_id, self.node_cnt]
            state = [t.clone() for t in saved]
            lstm_state = [self.lstm_h.clone(), self.lstm_c.clone()] if self.lstm is not None else None
            res = None
            for slot in range(self.depth):
                if slot > 0:
                    self.xs[slot].copy_(self.xs[0])
                r = self._body(slot)
                res = res or r
            torch.cuda.synchronize(self.device)
            if self.use_cuda_graph:
                for p in self.plans:
                    if p.graph is None:
                        p.capture()
                if self.tail_graph[0] is None:
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
            for t, s in zip(saved, state):
                t.copy_(s)
            if lstm_state is not None:
                self.lstm_h.copy_(lstm_state[0]); self.lstm_c.copy_(lstm_state[1])
            torch.cuda.synchronize(self.device)
        return res

    def _run_tail(self, slot):
        if self.tail_graph[slot] is not None:
            self.tail_graph[slot].replay()
            return self.out[slot]
        return self._tail(slot)

    def step_device(self, image_dev=None):
        """One frame with the image already in HBM (copied into the next arena's input slot if given).  Returns the
        frame's result tensors (device; complete after synchronize(), or join() + stream order)."""
        with _lib.on_device(self.device):
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
                with _Range(f"deft.dense[{slot}]"):
                    self.plans[slot].run(x)
                if self.depth == 1:
                    with _Range(f"deft.tail[{slot}]"):
                        return self._run_tail(slot)
                self._ev_dense[slot].record(ds)
            with torch.cuda.stream(self.tail_stream):
                self.tail_stream.wait_event(self._ev_dense[slot])
                with _Range(f"deft.tail[{slot}]"):
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

    def host_keys(self):
        """What step() brings back to the host: the detections and ONE matrix -- the [tracks, dets] cost when the
        track reduction is on (row f1), otherwise the raw per-memory-frame affinity block the reference copies."""
        if self.T_tracks and self.assign_thresh is not None:
            return ("scores", "clses", "bboxes", "match_det")          # association done on the device: only the matches travel
        return ("scores", "clses", "bboxes", "cost") if self.T_tracks else ("scores", "clses", "bboxes", "affinity")

    def step(self, image_host, out_host=None):
        """End-to-end frame through the public API: pinned host image in, host results out
        (H2D of the frame, D2H of detections + affinity/cost -- the copies the reference makes at
        detector.py:150,545-547 and AFE.py:154-158)."""
        # The upload runs on its own stream into a staging ring (one slot per frame in flight), so the H2D of frame t+1 overlaps the
        # kernels of frame t whenever the caller does not synchronise between steps (results land in pinned
        # memory asynchronously; synchronise the device / call synchronize() before reading them).
        with _lib.on_device(self.device):
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
                out_host = {k: torch.empty(res[k].shape, dtype=res[k].dtype, pin_memory=True) for k in self.host_keys()}
            with torch.cuda.stream(self.tail_stream if self.tail_stream is not None else cur):
                for k, v in out_host.items():
                    v.copy_(res[k], non_blocking=True)
            return out_host

    def count_launches(self):
        """Kernel launches of one frame (our kernels only): the library counts every launch it makes
        (deft_launch_count); one eager frame is run between two readings (recorder state restored)."""
        with _lib.on_device(self.device):
            torch.cuda.synchronize(self.device)
            saved = [self.memory, self.au_mem, self.clock, self.slot_frame, self.counts] + ([self.node_frame, self.node_id, self.node_cnt] if self.T_tracks else [])
            state = [t.clone() for t in saved]
            before = _lib.lib().deft_launch_count()
            self.plans[0].launch()
            self._tail(0)
            torch.cuda.synchronize(self.device)
            n = int(_lib.lib().deft_launch_count() - before)
            for t, s in zip(saved, state):
                t.copy_(s)
            torch.cuda.synchronize(self.device)
            return n
