"""Helpers shared by the parity tests: load the golden fixtures made by tests/golden/make_golden.py."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from _clips import fuzz_clip
from screenpressor_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digests():
    with open(os.path.join(GOLDEN_DIR, "ref_digests.json")) as f:
        return json.load(f)


def digests_long():
    """full-length BASELINE configs: per-frame [type, size, md5] written by the unmodified reference (make_golden.py)"""
    with open(os.path.join(GOLDEN_DIR, "ref_digests_long.json")) as f:
        return json.load(f)


def streams():
    return np.load(os.path.join(GOLDEN_DIR, "ref_streams.npz"))


def load_case(entry):
    """-> (clip, keyflags, width, height, bits_per_pixel)"""
    if entry["kind"] == "synth":
        name, n, interval = entry["args"]
        cfg = synth.CONFIGS[name]
        return synth.make_clip(cfg, n), synth.keyframe_flags(n, interval), cfg.width, cfg.height, cfg.bpp
    w, h, n, seed, bpp, levels = entry["args"]
    clip, keys = fuzz_clip(w, h, n, seed, bpp, levels)
    return clip, keys, w, h, bpp


def check_frames(name, rows, produced):
    """produced: list of (bytes, ftype).  Compare with the reference's [ftype, size, md5] rows."""
    assert len(produced) == len(rows)
    for i, ((data, ft), (gft, gsz, gmd5)) in enumerate(zip(produced, rows)):
        assert ft == gft, f"{name} frame {i}: frame type {ft} != reference {gft}"
        assert len(data) == gsz, f"{name} frame {i}: {len(data)} bytes != reference {gsz}"
        assert hashlib.md5(data).hexdigest() == gmd5, f"{name} frame {i}: bytes differ from the reference"
