"""Drop-in for src/lib/model/decode.py ``generic_decode(output, K=100, opt=None)`` (decode.py:102-219)
and ``Detector._sigmoid_output`` (detector.py:486-494).

NMS + exact two-stage top-K + all gathers/box arithmetic run in three kernels (deft_decode); the
result dict has the reference's keys, dtypes and shapes.  Indices/scores/boxes are bit-exact w.r.t.
the reference on the same (sigmoided) heat-map; among exactly tied scores the lowest flat index wins.
Heads DEFT never enables (hps / hm_hp keypoints, ltrb, cat-specific wh) are not implemented.
"""
import torch

from .. import _lib, ops

_REG_HEADS = ["tracking", "dep", "rot", "dim", "amodel_offset", "nuscenes_att", "velocity"]


def sigmoid_output(output, opt=None):
    """In place, like the reference: hm -> sigmoid; dep -> 1/(sigmoid(dep)+1e-6)-1, times depth_scale."""
    hm = output.get("hm")
    dep = output.get("dep")
    if hm is not None:
        _lib.require_device(hm)
    if dep is not None and not dep.is_contiguous():
        dep = output["dep"] = dep.contiguous()
    ops.sigmoid_output_(hm, dep, getattr(opt, "depth_scale", 1.0) if opt is not None else 1.0)
    return output


def generic_decode(output, K=100, opt=None):
    if "hm" not in output:
        return {}
    for bad in ("hps", "hm_hp", "ltrb"):
        if bad in output:
            raise NotImplementedError(f"generic_decode: head '{bad}' is not used by DEFT and not implemented")
    heat = output["hm"]
    _lib.require_device(heat)
    if opt is not None and getattr(opt, "zero_tracking", False):
        output["tracking"] *= 0
    batch, cat = heat.shape[:2]
    if "wh" in output and output["wh"].shape[1] != 2:
        raise NotImplementedError("generic_decode: cat-specific wh is not used by DEFT")
    d = ops.decode(heat, K, reg=output.get("reg"), wh=output.get("wh"), ltrb_amodal=output.get("ltrb_amodal"))
    inds = d["inds"]
    ret = {"scores": d["scores"], "clses": d["clses"].float(), "xs": d["xs"], "ys": d["ys"],
           "cts": torch.stack([d["xs"], d["ys"]], dim=2)}
    if d["bboxes"] is not None:
        ret["bboxes"] = d["bboxes"]
    for head in _REG_HEADS:
        if head in output:
            ret[head] = ops.gather_feat(output[head], inds)
    if d["bboxes_amodal"] is not None:
        ret["bboxes_amodal"] = d["bboxes_amodal"]
        ret["bboxes"] = d["bboxes_amodal"]
    return ret
