"""The hot-path pieces of the reference's src/utils.py under the same import path: Dotdict (needed to unpickle
shipped checkpoints), get_optimizers (:36-59) and load_checkpoint (:276-313).  Data loading, wandb
logging and the train/val loops of the reference file are out of scope."""
import torch

from maskedsst_b200.optim import FusedAdam


class Dotdict(object):
    def __init__(self, data):
        self.__dict__.update(data)


def get_optimizers(model, config, fused=True):
    """AdamW / Adam + ReduceLROnPlateau(.9, 5) / CosineAnnealingLR(50), reference src/utils.py:36-59; the update
    itself is the fused multi-tensor kernel (msst_adam_step) unless fused=False."""
    if config.optimizer not in ("Adam", "AdamW"):
        raise ValueError(config.optimizer)
    decoupled = config.optimizer == "AdamW"
    if fused and next(model.parameters()).is_cuda:
        optimizer = FusedAdam(model.parameters(), lr=config.lr, weight_decay=config.weight_decay, decoupled=decoupled,
                              clamp=1.0 if getattr(config, "clip_grad_norm", False) else 0.0)
    else:
        cls = torch.optim.AdamW if decoupled else torch.optim.Adam
        optimizer = cls(model.parameters(), lr=config.lr, weight_decay=config.weight_decay)
    if config.scheduler == "ReduceLROnPlateau":
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, factor=0.9, patience=5)
    elif config.scheduler == "cosine":
        scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=50, eta_min=0, last_epoch=-1)
    else:
        raise ValueError(config.scheduler)
    return optimizer, scheduler


def load_checkpoint(config, model, classifier_name, device):
    """Loads a SimMIM pre-training checkpoint into a fresh encoder: keeps only 'encoder.*' tensors (prefix
    stripped), optionally crops pos_embed, keeps the model's freshly initialised classifier, strict load."""
    checkpoint = torch.load(config.checkpoint_path, map_location=device, weights_only=False)
    weights = {k[len("encoder."):]: v for k, v in checkpoint["model_state_dict"].items() if k.startswith("encoder.")}
    linear_idx = 2 if model.pixelwise else 1
    head = model.mlp_head[linear_idx]
    if getattr(config, "patch_sub", 0) != 0 and weights.get("pos_embed") is not None:
        assert model.pos_embed.shape[1] == (config.image_size - config.patch_sub) ** 2
        weights["pos_embed"] = weights["pos_embed"][:, : model.pos_embed.shape[1], :]
    weights.pop(f"{classifier_name}.1.bias", None)
    weights.pop(f"{classifier_name}.1.weight", None)
    weights[f"{classifier_name}.{linear_idx}.bias"] = head.bias
    weights[f"{classifier_name}.{linear_idx}.weight"] = head.weight
    print(model.load_state_dict(weights))
    return model


def save_pretrain_checkpoint(path, model, config, losses, lr_current, img):
    """Writes a SimMIM pre-training checkpoint in the reference's pickle layout (pretrain.py:135-148): a dict with
    `losses` (1-D tensor), `config` (the Dotdict instance itself), `model_state_dict` (the SimMIM wrapper's state_dict,
    'encoder.'-prefixed keys + mask_token + to_pixels.*), `lr_current`, and the last batch under both `input` and
    `transformer_input` -- so the reference's load_checkpoint (src/utils.py:276-313) and ours read it unchanged."""
    stats = {
        "losses": torch.as_tensor(losses),
        "config": config,
        "model_state_dict": model.state_dict(),
        "lr_current": lr_current,
        "input": img.detach(),
        "transformer_input": img,
    }
    torch.save(stats, path)
    return stats


def save_finetune_checkpoint(path, model, config, lr_current, epoch):
    """Fine-tuning checkpoint in the reference layout (src/utils.py:584-601): `config` as a plain dict, the encoder's
    state_dict (no prefix), `lr_current`, `epoch`."""
    stats = {
        "config": dict(config.__dict__) if not isinstance(config, dict) else config,
        "model_state_dict": model.state_dict(),
        "lr_current": lr_current,
        "epoch": epoch,
    }
    torch.save(stats, path)
    return stats
