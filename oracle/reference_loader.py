"""Load the UNMODIFIED reference model files from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  The GPU box has no /root/reference: nothing that runs there
may call this.  Used by oracle/make_golden.py (fixture generation) and by the
``not gpu`` pin tests (skipped when the reference tree is absent).

One source-text shim is applied at load time:
``top_k_words / self.vocab_size`` -> ``//`` so that ``sample()`` indexes with integers on
torch >= 1.5 (stylenet/model.py:249, model_att.py:381, nic/model.py:162,
nic/model_att.py:261, app/backend/model.py:170,442).  Nothing else is changed.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("SN_REFERENCE_ROOT", "/root/reference")

_FILES = {
    "stylenet": "stylenet/model.py",
    "stylenet_att": "stylenet/model_att.py",
    "nic": "nic/model.py",
    "nic_att": "nic/model_att.py",
    "app": "app/backend/model.py",
    "app_att": "app/backend/model_att.py",
    "stylenet_utils": "stylenet/utils.py",
    "seq2seq": "seq2seq/model.py",
}
_cache = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, _FILES["stylenet"]))


def load(name: str) -> types.ModuleType:
    """Return the reference module ``name`` (one of _FILES) with the ``//`` shim."""
    if name in _cache:
        return _cache[name]
    path = os.path.join(REF_ROOT, _FILES[name])
    with open(path, "r") as fh:
        src = fh.read()
    src = src.replace("top_k_words / self.vocab_size", "top_k_words // self.vocab_size")
    mod = types.ModuleType("sn_reference_" + name)
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    _cache[name] = mod
    return mod
