"""Drop-in for src/lib/model/model.py: create_model / load_model / save_model (model.py:32-121)."""
import torch

from .networks.dla import DLASeg

_network_factory = {"dla": DLASeg}


def create_model(arch, head, head_conv, opt=None):
    num_layers = int(arch[arch.find("_") + 1:]) if "_" in arch else 0
    arch = arch[: arch.find("_")] if "_" in arch else arch
    if arch not in _network_factory:
        raise NotImplementedError(f"deft_b200 implements arch 'dla_34' only (got '{arch}'): the other reference backbones "
                                  "return no FeatureMaps and cannot drive DEFT's tracker (SURVEY.md section 2.1 #15)")
    return _network_factory[arch](num_layers, heads=head, head_convs=head_conv, opt=opt)


def load_model(model, model_path, opt, optimizer=None):
    """Same tolerant semantics as model.py:40-110: strip 'module.' prefixes, skip shape mismatches,
    fill missing keys from the model, load non-strict; optionally replay the LR schedule."""
    checkpoint = torch.load(model_path, map_location=lambda storage, loc: storage, weights_only=False)
    print("loaded {}, epoch {}".format(model_path, checkpoint.get("epoch", "?")))
    state_dict_ = checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint
    state_dict = {}
    for k in state_dict_:
        if k.startswith("module") and not k.startswith("module_list"):
            state_dict[k[7:]] = state_dict_[k]
        else:
            state_dict[k] = state_dict_[k]
    model_state_dict = model.state_dict()
    for k in list(state_dict.keys()):
        if k in model_state_dict:
            if state_dict[k].shape != model_state_dict[k].shape:
                print("Skip loading parameter {}, required shape{}, loaded shape{}.".format(
                    k, model_state_dict[k].shape, state_dict[k].shape))
                state_dict[k] = model_state_dict[k]
        else:
            print("Drop parameter {}.".format(k))
    for k in model_state_dict:
        if k not in state_dict:
            print("No param {}.".format(k))
            state_dict[k] = model_state_dict[k]
    model.load_state_dict(state_dict, strict=False)
    if optimizer is not None and getattr(opt, "resume", False) and "optimizer" in checkpoint:
        start_epoch = checkpoint["epoch"]
        start_lr = opt.lr
        for step in opt.lr_step:
            if start_epoch >= step:
                start_lr *= 0.1
        for param_group in optimizer.param_groups:
            param_group["lr"] = start_lr
        print("Resumed optimizer with start lr", start_lr)
        return model, optimizer, start_epoch
    return model


def save_model(path, epoch, model, optimizer=None):
    if isinstance(model, torch.nn.DataParallel):
        state_dict = model.module.state_dict()
    else:
        state_dict = model.state_dict()
    data = {"epoch": epoch, "state_dict": state_dict}
    if optimizer is not None:
        data["optimizer"] = optimizer.state_dict()
    torch.save(data, path)
