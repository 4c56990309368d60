"""The dozen `opt` fields the hot path reads (reference src/lib/opts.py) as a plain
dataclass -- no argv parsing at import (the reference's tracker.py:139 does that)."""
from dataclasses import dataclass, field
from typing import Dict, List


@dataclass
class Opt:
    dataset: str = "mot"            # AFE.py:23 ("nuscenes" selects the 704-d embedding / 18-d LSTM)
    max_object: int = 100           # AFE.py:63 (opts.py:339)
    head_kernel: int = 3            # base_model.py:26-30
    prior_bias: float = -4.6        # base_model.py:92 (opts.py:151)
    dla_node: str = "dcn"           # dla.py:766
    load_model: str = "synthetic"   # dla.py:770 (non-empty: never download ImageNet weights)
    pre_img: bool = False           # dla.py:346
    pre_hm: bool = False            # dla.py:354
    model_output_list: bool = False  # base_model.py:119
    zero_tracking: bool = False     # decode.py:106
    K: int = 100                    # detector.py:544 (opts.py:235)
    depth_scale: float = 1.0        # detector.py:493
    lstm: bool = False              # kalman_filter_lstm.py:52
    load_model_traj: str = ""       # kalman_filter_lstm.py:53
    heads: Dict[str, int] = field(default_factory=lambda: dict(hm=1, reg=2, wh=2, tracking=2, ltrb_amodal=4))
    head_conv: Dict[str, List[int]] = field(default_factory=dict)
    # deft_b200 extensions
    grid_sample_align_corners: bool = False   # torch>=1.3 default; the reference's torch 1.2 used True
    tensor_core: bool = True        # tcgen05 engine (3xTF32, fp32-accurate); False = SIMT fp32 anchor
    tf32x1: bool = False            # single-pass TF32 (outside the 1e-3 parity budget; benchmarking only)
    clone_outputs: bool = False     # True: return fresh tensors instead of views into the frame arena
