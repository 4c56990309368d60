"""Import the REAL reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box; callers must use
``available()`` first.  Follows the verified recipe of SURVEY.md section 8(c): never import
``model.model`` (broken on modern torchvision), give DLASeg a non-empty ``load_model`` so it does
not try to download ImageNet weights (dla.py:769-771).
"""
import os
import sys
import warnings
from types import SimpleNamespace

REF_LIB = "/root/reference/src/lib"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available():
    return os.path.isdir(REF_LIB)


def make_opt(dataset="mot", max_object=100):
    return SimpleNamespace(head_kernel=3, prior_bias=-4.6, dataset=dataset, max_object=max_object,
                           dla_node="dcn", load_model="synthetic", pre_img=False, pre_hm=False,
                           model_output_list=False, zero_tracking=False)


def load():
    """Returns (dla_module, AFE_module_mod, decode_module)."""
    sys.dont_write_bytecode = True          # the reference tree is read-only
    for p in (_SHIM, REF_LIB):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", category=UserWarning)
    import model.networks.dla as dla        # picks up the dcn_v2 shim through dla.py:26
    import model.networks.AFE as AFE
    import model.decode as decode
    return dla, AFE, decode


def build_dlaseg(heads, dataset="mot", max_object=100, head_conv=256):
    dla, _, _ = load()
    opt = make_opt(dataset, max_object)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = dla.DLASeg(34, dict(heads), {h: [head_conv] for h in heads}, opt).eval()
    return m, opt
