"""Checkpoint files in the reference's layout: one `torch.save`d dict per model version at `weights/cp-<version>`,
mapping a fixed key per network to its `state_dict()` (apgs/gmm/apg_training.py:101-115, apgs/dmm/apg_training.py:103-116,
apgs/bmnist/apg_training.py:95-104).  The encoders / decoders of this package keep the reference's parameter names, so a
file written here loads in the reference and vice versa (`init_*_models(load_version=...)`).  Like the reference, no
optimizer or RNG state is stored: the particle population lives only inside one apg_objective call."""
import os

import torch

# task -> checkpoint key of every TRAINED network, in the order of the task's `models` tuple (None: not stored)
LAYOUT = {
    "gmm": ("enc-rws-eta", "enc-apg-z", "enc-apg-eta", None),
    "dmm": ("enc-rws-mu", "enc-apg-local", "enc-apg-mu", "dec"),
    "bmnist": ("enc-coor", None, "enc-digit", "dec-digit"),
}


def path_of(version, root="weights"):
    return os.path.join(root, "cp-%s" % version)


def save(task, models, version, root="weights"):
    keys = LAYOUT[task]
    if len(models) != len(keys):
        raise ValueError("%s expects %d models, got %d" % (task, len(keys), len(models)))
    ckpt = {k: m.state_dict() for k, m in zip(keys, models) if k is not None}
    os.makedirs(root, exist_ok=True)
    torch.save(ckpt, path_of(version, root))
    return path_of(version, root)


def load(task, models, version, root="weights", map_location=None):
    weights = torch.load(path_of(version, root), map_location=map_location)
    for k, m in zip(LAYOUT[task], models):
        if k is not None:
            m.load_state_dict(weights[k])
    return models
