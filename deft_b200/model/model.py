"""Drop-in for src/lib/model/model.py: ``create_model`` / ``load_model`` / ``save_model`` with the reference's
signatures and checkpoint format (model.py:32-121), so DEFT's detector / trainer entry points and released
``.pth`` files work unchanged.

Checkpoint contract kept from the reference: a dict with ``state_dict`` (optionally ``epoch`` / ``optimizer``), keys
possibly prefixed with ``module.`` by nn.DataParallel; loading is tolerant (model.py:56-91) -- a tensor whose shape
does not fit, or a key the model does not have, is reported and skipped, and the model keeps its own value.
"""
import torch

from .networks.dla import DLASeg

_BACKBONES = {"dla": DLASeg}


def _split_arch(arch):
    """'dla_34' -> ('dla', 34); a bare name means 0 layers (model.py:33-34)."""
    name, sep, depth = arch.partition("_")
    return name, int(depth) if sep else 0


def create_model(arch, head, head_conv, opt=None):
    name, depth = _split_arch(arch)
    try:
        backbone = _BACKBONES[name]
    except KeyError:
        raise NotImplementedError(f"deft_b200 implements arch 'dla_34' only (got '{arch}'): the other reference backbones "
                                  "return no FeatureMaps and cannot drive DEFT's tracker (SURVEY.md section 2.1 #15)") from None
    return backbone(depth, heads=head, head_convs=head_conv, opt=opt)


def _without_dataparallel_prefix(tensors):
    """nn.DataParallel saves 'module.<key>'; 'module_list...' keys are real names and stay (model.py:49-53)."""
    out = {}
    for key, value in tensors.items():
        wrapped = key.startswith("module") and not key.startswith("module_list")
        out[key[len("module."):] if wrapped else key] = value
    return out


def _reconcile(loaded, own, reset_hm=False, reuse_hm=False):
    """What load_state_dict(strict=False) should see: every key of the model, taken from the checkpoint when it is
    there with the right shape, otherwise the model's current tensor.  Returns (merged, report lines).
    reset_hm / reuse_hm (model.py:58-77): with reset_hm, `hm*` tensors whose leading dimension is 80 (COCO) or 1
    (CrowdHuman) are treated like a shape mismatch even when the shapes agree; a mismatched tensor is then either
    skipped or, with reuse_hm, its leading slice is reused (the reference's `a < a` comparison at model.py:70 is
    always False, so the executed branch is `loaded[:required]`, which is what is restated here)."""
    merged, report = {}, []
    for key, value in loaded.items():
        if key not in own:
            report.append(f"Drop parameter {key}.")
            merged[key] = value                  # harmless under strict=False; kept so the behaviour matches
            continue
        mismatch = tuple(value.shape) != tuple(own[key].shape)
        if not mismatch and reset_hm and key.startswith("hm") and value.dim() > 0 and value.shape[0] in (80, 1):
            mismatch = True
        if not mismatch:
            merged[key] = value
        elif reuse_hm:
            report.append(f"Reusing parameter {key}, required shape{own[key].shape}, loaded shape{value.shape}.")
            merged[key] = value[: own[key].shape[0]]
        else:
            report.append(f"Skip loading parameter {key}, required shape{own[key].shape}, loaded shape{value.shape}.")
            merged[key] = own[key]
    for key in own:
        if key not in loaded:
            report.append(f"No param {key}.")
            merged[key] = own[key]
    return merged, report


def _resumed_lr(opt, epoch):
    lr = opt.lr
    for milestone in opt.lr_step:                # model.py:98-103: one decade per passed milestone
        if epoch >= milestone:
            lr *= 0.1
    return lr


def load_model(model, model_path, opt, optimizer=None):
    ckpt = torch.load(model_path, map_location="cpu", weights_only=False)
    print("loaded {}, epoch {}".format(model_path, ckpt.get("epoch", "?") if isinstance(ckpt, dict) else "?"))
    tensors = ckpt["state_dict"] if isinstance(ckpt, dict) and "state_dict" in ckpt else ckpt
    merged, report = _reconcile(_without_dataparallel_prefix(tensors), model.state_dict(), reset_hm=bool(getattr(opt, "reset_hm", False)),
                                reuse_hm=bool(getattr(opt, "reuse_hm", False)))
    for line in report:
        print(line)
    model.load_state_dict(merged, strict=False)
    if optimizer is None:
        return model
    # with an optimizer the reference always answers (model, optimizer, start_epoch) (model.py:94-110); the learning
    # rate is re-derived from the schedule, the optimizer state itself is not restored there either
    epoch = 0
    if getattr(opt, "resume", False):
        if isinstance(ckpt, dict) and "optimizer" in ckpt:
            epoch = ckpt["epoch"]
            lr = _resumed_lr(opt, epoch)
            for group in optimizer.param_groups:
                group["lr"] = lr
            print("Resumed optimizer with start lr", lr)
        else:
            print("No optimizer parameters in checkpoint.")
    return model, optimizer, epoch


def save_model(path, epoch, model, optimizer=None):
    net = model.module if isinstance(model, torch.nn.DataParallel) else model
    payload = {"epoch": epoch, "state_dict": net.state_dict()}
    if optimizer is not None:
        payload["optimizer"] = optimizer.state_dict()
    torch.save(payload, path)
