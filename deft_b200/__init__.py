"""deft_b200 -- B200-native (sm_100a) implementation of DEFT's per-frame dense hot path.

Drop-in surface (same names / signatures / state_dict keys as the reference's src/lib/model):
    deft_b200.model.model        create_model, load_model, save_model
    deft_b200.model.networks.dla DLASeg
    deft_b200.model.networks.AFE AFE_module
    deft_b200.model.decode       generic_decode
    deft_b200.dcn_v2             DCN
    deft_b200.kalman_filter_lstm DecoderRNN, KalmanFilterLSTM
All arithmetic runs in hand-written CUDA kernels behind the C ABI of include/deft_b200.h
(libdeft_b200.so); there is no CPU or PyTorch-eager fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
