"""Config / checkpoint plumbing compatible with the reference ``basemodel.py`` (one
``np.savez`` file per network in a directory + a JSON ``config``; reference basemodel.py:17-182).
Host I/O only."""
import json
import os

import numpy as np
import torch


class Config(object):
    def __init__(self, **params):
        super().__setattr__("memo", [])
        for k, v in params.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if name not in self.memo:
            self.memo.append(name)
        super().__setattr__(name, value)

    def __delattr__(self, name):
        self.memo.remove(name)
        super().__delattr__(name)

    def __getitem__(self, k):
        assert k in self.memo, f"{k} not found, try {self.memo}"
        return getattr(self, k)

    def __contains__(self, k):
        return k in self.memo

    def __repr__(self):
        return "class Config containing: " + str({k: getattr(self, k) for k in self.memo})

    def save(self, path):
        with open(path, "w") as f:
            json.dump({k: getattr(self, k) for k in self.memo}, f)

    def load(self, path):
        for k in list(self.memo):
            delattr(self, k)
        with open(path) as f:
            for k, v in json.load(f).items():
                setattr(self, k, v)


def ckpt_save(ckpt, folder):
    assert isinstance(ckpt, dict)
    assert not os.path.exists(folder), folder + " already exists"
    os.mkdir(folder)
    for key, val in ckpt.items():
        path = os.path.join(folder, key)
        if key == "config":
            val.save(path)
        else:
            with open(path, "wb") as f:
                np.savez(f, **{k: v.detach().cpu().numpy() for k, v in val.items()})


def ckpt_load(folder):
    if os.path.isfile(folder):
        return torch.load(folder)
    ckpt = {}
    for key in os.listdir(folder):
        path = os.path.join(folder, key)
        if key == "config":
            ckpt[key] = Config()
            ckpt[key].load(path)
        else:
            z = np.load(path)
            ckpt[key] = {k: torch.from_numpy(z[k]) for k in z.files}
    return ckpt


class BaseModel(object):
    def __init__(self, cfg=None, ckpt=None, objects=None):
        if ckpt is not None:
            self.load(cfg=cfg, ckpt=ckpt, objects=objects)
        else:
            self.build(cfg)
        self.training = True

    def build(self, cfg):
        self.cfg = cfg

    def _modules(self):
        return {k: v for k, v in self.__dict__.items() if isinstance(v, torch.nn.Module)}

    def to(self, device):
        for m in self._modules().values():
            m.to(device)
        return self

    def train(self, mode=True):
        for m in self._modules().values():
            m.train(mode)
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def get_saveable(self):
        return self._modules()

    def save(self, ckpt, objects=None):
        saveable = self.get_saveable()
        objects = {k: saveable[k].state_dict() for k in (objects or saveable.keys())}
        objects["config"] = self.cfg
        ckpt_save(objects, ckpt)

    def load(self, ckpt, cfg=None, objects=None):
        ckpt = ckpt_load(ckpt)
        if cfg is None:
            cfg = ckpt.pop("config")
        self.build(cfg)
        saveable = self.get_saveable()
        for k in (objects or saveable.keys()):
            # a requested network that is not in the checkpoint is an error, like the reference's ckpt[k]
            # (basemodel.py:178-181): a typo in --load_nets must not leave a network at random init
            if k not in ckpt:
                raise KeyError(f"{k!r} requested but not in the checkpoint (has {sorted(ckpt)})")
            if k not in saveable:
                raise KeyError(f"{k!r} is not a network of this model (has {sorted(saveable)})")
            saveable[k].load_state_dict(ckpt[k])
