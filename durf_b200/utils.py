"""Data contract of the hot path: the reference's Rays/BoxRays tuples and the Config fields the path reads
(internal/utils.py:77-144).  gin is not required: `Config` is a plain dataclass carrying configs/carla_dyn.gin's
values as defaults; `load_gin` parses the simple `Name.field = value` lines of the reference's .gin files."""
from __future__ import annotations

import ast
import collections
from dataclasses import dataclass, fields

Rays = collections.namedtuple('Rays', ('origins', 'directions', 'viewdirs', 'radii', 'lossmult', 'near', 'far'))
BoxRays = Rays   # internal/utils.py:84-86


def namedtuple_map(fn, tup):
    """internal/utils.py:162-165."""
    return type(tup)(*map(fn, tup))


@dataclass
class Config:
    """internal/utils.py:89-144 (fields used by the train step), defaults = configs/carla_dyn.gin."""
    batch_size: int = 512
    near: float = 0.0
    far: float = 200.0
    timesteps: int = 5
    lr_init: float = 5e-4
    lr_final: float = 5e-6
    lr_delay_steps: int = 2500
    lr_delay_mult: float = 0.01
    max_steps: int = 200000
    eps_init: float = 3.0
    eps_final: float = 0.2
    eps_max_steps: int = 200000
    eps_delay_steps: int = 0
    alpha_init: float = 10.0
    alpha_final: float = 10.0
    alpha_delay_steps: int = 0
    alpha_max_steps: int = 1
    coarse_loss_mult: float = 0.1
    box_loss_mult: float = 0.0
    tv_loss_mult: float = 0.0
    depth_loss_mult: float = 0.0001
    near_loss_mult: float = 0.01
    empty_loss_mult: float = 1.0
    sky_loss_mult: float = 1.0
    weight_decay_mult: float = 0.0
    grad_max_norm: float = 1.0
    grad_max_val: float = 0.1
    disable_multiscale_loss: bool = False
    randomized: bool = True
    white_bkgd: bool = False
    rand_bkgd: bool = False
    # dataset loader (internal/utils.py:93-104, values of configs/carla_dyn.gin)
    dataset_loader: str = 'carla_dyn'
    batching: str = 'timestep'
    factor: int = 4
    spherify: bool = True
    centering: bool = True
    random_box: bool = False
    random_yaw: bool = False
    box_noise: float = 0.5
    yaw_noise: float = 5.0
    render_path: bool = False
    llffhold: int = 11


def load_gin(path: str):
    """Parse `Config.x = v` / `MipNerfModel.x = v` / `MLP.x = v` lines -> (config_kwargs, model_kwargs)."""
    cfg, model = {}, {}
    cfg_fields = {f.name for f in fields(Config)}
    for line in open(path):
        line = line.split('#')[0].strip()
        if '=' not in line:
            continue
        key, val = [s.strip() for s in line.split('=', 1)]
        scope, name = key.split('.', 1)
        try:
            v = ast.literal_eval(val)
        except Exception:
            v = val
        if scope == 'Config' and name in cfg_fields:
            cfg[name] = v
        elif scope == 'MipNerfModel':
            model[name] = v
        elif scope == 'MLP':
            model[{'net_width': 'net_width', 'net_depth': 'net_depth', 'net_width_condition': 'net_width_condition',
                   'net_depth_condition': '_net_depth_condition'}.get(name, '_' + name)] = v
    model = {k: v for k, v in model.items() if not k.startswith('_')}
    return cfg, model
