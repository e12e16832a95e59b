"""Shared test helpers: golden-fixture loading and config reconstruction."""
import ast
import os

import numpy as np
import torch

import neurad_studio_b200 as nsb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    meta = ast.literal_eval(str(z["__meta__"]))
    groups = {}
    for k in z.files:
        if k == "__meta__":
            continue
        g, rest = k.split("/", 1)
        groups.setdefault(g, {})[rest] = torch.from_numpy(z[k])
    return meta, groups


def cfg_from_meta(meta):
    return nsb.small_config(
        n_actors=meta["n_actors"], log2_main=meta["log2_main"], log2_prop=meta["log2_prop"],
        static_scale=meta["static_scale"], duration=meta["duration"], num_sensors=meta["num_sensors"],
    )
