"""Helpers shared by the oracle-pin tests and the GPU parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def sd_from(rec, prefix="sd.", dtype=torch.float64):
    out = {}
    for k, v in rec.items():
        if k.startswith(prefix):
            t = torch.from_numpy(v)
            out[k[len(prefix):]] = t.to(dtype) if t.is_floating_point() else t
    return out


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    d = (a - b).norm().item()
    n = b.norm().item()
    return d / n if n > 0 else d


def build_port(name, rec, dtype=torch.float64, prefix="sd.", dropout=0.0):
    """Instantiate the oracle port for golden case ``name`` with the recorded weights."""
    from oracle import port
    V, E, H, F, A, D = (int(rec["meta." + k]) for k in ("V", "E", "H", "F", "A", "D"))
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        if name == "factored":
            m = port.DecoderFactoredLSTM(E, H, F, V, 1, dropout=dropout, max_seq_length=12)
        elif name == "factored_att":
            m = port.DecoderFactoredLSTMAtt(A, E, H, F, V, 1, feature_size=D, dropout=dropout, max_seq_length=12)
        elif name == "nic":
            m = port.DecoderRNN(E, H, V, 1, dropout=dropout, max_seq_length=12)
        elif name == "nic_att":
            m = port.DecoderRNNAtt(A, E, H, V, 1, feature_size=D, dropout=dropout, max_seq_length=12)
        else:
            raise KeyError(name)
    finally:
        torch.set_default_dtype(old)
    m.load_state_dict(sd_from(rec, prefix, dtype))
    return m


CASES = {
    # name: (attention?, modes)
    "factored": (False, ["factual", "happy"]),
    "factored_att": (True, ["factual", "angry"]),
    "nic": (False, [None]),
    "nic_att": (True, [None]),
}
