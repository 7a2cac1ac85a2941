"""Shared test helpers: rebuild the synthetic case behind a golden fixture."""
import json
import os

import numpy as np
import torch

from alignsdf_b200 import synthetic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_index():
    with open(os.path.join(GOLD, "index.json")) as f:
        return json.load(f)


def field_cases():
    return [k for k in golden_index() if not k.startswith("legacy")]


def load_case(name):
    meta = golden_index()[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    dec = synthetic.make_decoder(meta["seed"], meta["kind"], meta["latent_size"], meta["pf"],
                                 meta["style"], meta["network_specs"],
                                 use_classifier=meta.get("use_classifier", False),
                                 init=meta.get("init", "engineered"), out_gain=meta.get("out_gain", 1.0),
                                 bias_shift=meta.get("bias_shift"))
    sample = synthetic.make_sample(meta["seed"], meta["latent_size"], meta["pf"], meta["style"],
                                   pixel_align=tuple(meta["pixel_align"]) if meta.get("pixel_align") else None)
    return meta, g, dec, sample


def to_cuda(sample):
    return sample.to(torch.device("cuda"))
