"""Seeded synthetic weights for benchmarking without checkpoints (there is no network for real ones).

PyTorch's default init leaves BatchNorm at identity and the upstream DCN init zeroes conv_offset_mask, which
would degenerate the deformable gather into a plain conv; this gives every BN non-trivial running stats and
the offset convs ~1 px offsets so the benchmark exercises the real data-dependent path."""
import math

import torch


def randomize_(model, seed=317):
    g = torch.Generator()
    g.manual_seed(seed)
    with torch.no_grad():
        for name, p in model.state_dict().items():
            if name.endswith("num_batches_tracked"):
                continue
            if name.endswith("running_var") or (name.endswith(".weight") and p.dim() == 1):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif name.endswith("running_mean") or name.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif "conv_offset_mask.weight" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.01)
            elif ".up_" in name:
                p.add_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.05)
            elif p.dim() == 4:
                fan_in = p.shape[1] * p.shape[2] * p.shape[3]
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * math.sqrt(4.5 / fan_in))
            elif p.dim() == 2:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) / math.sqrt(p.shape[1]))
        for name, p in model.state_dict().items():
            if "conv_offset_mask.bias" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.75)
            if name == "hm.2.bias":
                p.fill_(-4.6)
    if hasattr(model, "invalidate"):
        model.invalidate()
    return model
