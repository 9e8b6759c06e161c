"""Readers for the committed fixtures under tests/golden/ (written by tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def unpack_w(arr, tag):
    """Stored weight -> float32 array holding the exact values (bf16 is stored as raw uint16 bits)."""
    if tag == "bf16":
        return (arr.astype(np.uint32) << 16).view(np.float32)
    return arr.astype(np.float32)


def layer(npz, key):
    tag = str(npz[f"{key}|tag"])
    d = {"tag": tag, "W_before": unpack_w(npz[f"{key}|W_before"], tag)}
    C = d["W_before"].shape[1]
    d["mask"] = np.unpackbits(npz[f"{key}|mask"], axis=1)[:, :C].astype(bool)
    d["W_after"] = unpack_w(npz[f"{key}|W_after"], tag) if f"{key}|W_after" in npz.files \
        else np.where(d["mask"], d["W_before"], np.float32(0))
    for k in npz.files:
        if k.startswith(key + "|") and k.split("|")[1] not in ("tag", "W_before", "W_after", "mask"):
            d[k.split("|")[1]] = npz[k]
    return d


def to_torch(w32, tag, device="cpu"):
    import torch
    dt = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[tag]
    return torch.from_numpy(np.ascontiguousarray(w32)).to(dt).to(device)


DSNOT_CASES = {   # golden case -> oracle / kernel keyword arguments (tests/golden/make_golden.py gen_dsnot_toy)
    "shipped_unstr60": dict(),
    "upstream_unstr60": dict(ref_fixup=False),
    "upstream_unstr60_samesign": dict(ref_fixup=False, without_same_sign=False),
    "upstream_unstr60_magnitude": dict(ref_fixup=False, initial_method="magnitude"),
    "shipped_2of4": dict(prune_n=2, prune_m=4),
    "shipped_4of8": dict(prune_n=4, prune_m=8),
}


def dsnot_layer(npz, case, key):
    """(W float32, dtype tag, stats dict, reference keep mask) of one linear of one DSnoT golden case."""
    tag = str(npz[f"{key}|tag"])
    W = unpack_w(npz[f"{key}|W_before"], tag)
    keep = np.unpackbits(npz[f"{case}|{key}|mask"], axis=1)[:, :W.shape[1]].astype(bool)
    stats = {s: npz[f"{key}|{s}"] for s in ("scaler_row", "sum_metric_row", "var")}
    return W, tag, stats, keep
