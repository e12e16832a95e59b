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


def load_config1():
    """BASELINE config 1 fixture (tests/golden/config1.npz, written by oracle/make_golden_config1.py from the real
    reference): parameters (the 64 MB hash table is re-created from its deterministic generator and spot-checked
    against the committed sub-sample), rays and the reference's outputs."""
    from oracle import simple_oracle as S

    meta, g = load_golden("config1.npz")
    p = dict(g["param"])
    n_rows = p["scalings"].shape[0] * (1 << meta["log2_hashmap_size"])
    p["hash_table"] = S.synthetic_table(n_rows, 2, meta["table_scale"])
    assert torch.equal(p["hash_table"][::4099], p.pop("hash_table_sub"))
    return meta, p, g["ray"], g["ref"]
