"""Shared helpers for the parity tests."""
import os

import numpy as np

from pycusdr_b200.config import loadModularJson
from pycusdr_b200.protocol import loadProtocol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RADIO = "UHF-H"


def load_conf(rel):
    return loadModularJson(os.path.join(ROOT, "config", rel))


def conf_variant(rel, blockSize=None, doppCarrierSteps=None, **radio_over):
    conf = load_conf(rel)
    if blockSize is not None:
        conf["GPU"]["UHF"]["blockSize"] = blockSize
    if doppCarrierSteps is not None:
        conf["Radios"]["Rx"][RADIO]["doppCarrierSteps"] = doppCarrierSteps
    conf["Radios"]["Rx"][RADIO].update(radio_over)
    return conf


def protocol_for(conf):
    name = conf["Main"]["protocols"]["UHF"]
    return loadProtocol(name)(conf=conf)


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    wide = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a, b = a.astype(wide), b.astype(wide)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
