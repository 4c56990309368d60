"""Drop-in for src/lib/utils/tracking_utils/kalman_filter_lstm.py: ``DecoderRNN`` (9-29) and
``KalmanFilterLSTM.predict`` (65-78).  One fused kernel does the LSTM step + out1 + out2 for all
tracks; ``predict`` keeps the reference's one-track signature, ``predict_batch`` is the B200 way."""
import torch
from torch import nn

from . import _lib, ops


class DecoderRNN(nn.Module):
    def __init__(self, num_hidden, dataset):
        super().__init__()
        if num_hidden != 128:
            raise NotImplementedError("deft_b200 DecoderRNN: num_hidden must be 128 (the reference's only value)")
        self.num_hidden = num_hidden
        fin, fout = (18, 16) if dataset == "nuscenes" else (11, 20)
        self.lstm = nn.LSTM(fin, num_hidden)          # parameter holder: weight_ih_l0, weight_hh_l0, bias_*_l0
        self.out1 = nn.Linear(num_hidden, 64)
        self.out2 = nn.Linear(64, fout)
        self.fin, self.fout = fin, fout

    def weights(self, device=None):
        """Contiguous fp32 parameter tensors for deft_lstm_step.  `device` = the device of the call's inputs: the
        kernel dereferences raw pointers, so parameters living anywhere else (e.g. a model left on the CPU) must be
        refused here -- there is no host fallback."""
        if device is not None:
            for name, prm in self.named_parameters():
                if prm.device != device:
                    raise _lib.DeftError(f"DecoderRNN parameter '{name}' is on {prm.device} but the inputs are on {device}: "
                                         "move the model first (KalmanFilterLSTM(opt) does, like the reference: kalman_filter_lstm.py:52-57)")
        g = lambda t: t.detach().float().contiguous()
        return dict(w_ih=g(self.lstm.weight_ih_l0), w_hh=g(self.lstm.weight_hh_l0), b_ih=g(self.lstm.bias_ih_l0),
                    b_hh=g(self.lstm.bias_hh_l0), w1=g(self.out1.weight), b1=g(self.out1.bias), w2=g(self.out2.weight),
                    b2=g(self.out2.bias))

    def step(self, x, h, c):
        """x [T,F], h,c [T,128] -> (h', c', out [T,fout]) in one launch."""
        _lib.require_device(x)
        if h.device != x.device or c.device != x.device:
            raise _lib.DeftError("DecoderRNN.step: x, h and c must live on the same device")
        return ops.lstm_step(x.contiguous(), h.contiguous(), c.contiguous(), self.weights(x.device), self.fout)

    def forward(self, input_traj):
        """[B,T,F] -> [B,fout]: the reference unrolls nn.LSTM over T from zero state, then out2(out1(h_T))."""
        _lib.require_device(input_traj)
        B, T, _ = input_traj.shape
        h = torch.zeros(B, 128, device=input_traj.device)
        c = torch.zeros(B, 128, device=input_traj.device)
        out = None
        w = self.weights(input_traj.device)
        for t in range(T):
            h, c, out = ops.lstm_step(input_traj[:, t].contiguous(), h, c, w, self.fout)
        return out


class KalmanFilterLSTM(object):
    def __init__(self, opt, model=None):
        """Reference constructor (kalman_filter_lstm.py:50-62): build DecoderRNN(128, opt.dataset); when
        opt.load_model_traj is set, load that checkpoint through load_model, move the model to opt.device and switch to
        eval().  `model=` (deft_b200 extension) injects an already prepared DecoderRNN instead."""
        self.opt = opt
        if model is None:
            model = DecoderRNN(128, opt.dataset)
            if getattr(opt, "load_model_traj", "") != "":
                from .model.model import load_model
                model = load_model(model, opt.load_model_traj, opt)
            dev = getattr(opt, "device", None)
            if dev is None:                                    # the reference derives it from opt.gpus (line 56)
                gpus = getattr(opt, "gpus", [0])
                dev = torch.device("cuda", gpus[0]) if gpus and gpus[0] >= 0 else torch.device("cuda")
                try:
                    opt.device = dev
                except Exception:
                    pass
            model = model.to(dev).eval()
        self.model = model
        self.MAX_dis_fut = 4 if opt.dataset == "nuscenes" else 5

    def predict(self, h0, c0, new_features):
        """h0,c0 [1,1,128], new_features [1,1,F] -> (hn, cn, {1..K: np[4]})  (one D2H, like the reference)."""
        hn, cn, y = self.model.step(new_features.reshape(1, -1), h0.reshape(1, 128), c0.reshape(1, 128))
        x = y.view(self.MAX_dis_fut, -1).cpu().numpy()
        return hn.view(1, 1, 128), cn.view(1, 1, 128), {1 + i: x[i] for i in range(self.MAX_dis_fut)}

    def predict_batch(self, h, c, feats):
        """All T tracks in one launch, everything stays on the device: -> (h', c', pred [T,K,4])."""
        hn, cn, y = self.model.step(feats, h, c)
        return hn, cn, y.view(feats.shape[0], self.MAX_dis_fut, -1)
