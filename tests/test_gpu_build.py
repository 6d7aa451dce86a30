"""The native library is built FROM SOURCE on the GPU box (nvcc for sm_100a, no cached objects: _obj/ does not
travel) and must behave exactly like the prebuilt libastc_b200.so the other tests load."""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]

_PROBE = r"""
import hashlib, sys, torch
sys.path.insert(0, %r)
import astc_encoder_b200 as A
from astc_encoder_b200 import synth
h = hashlib.sha256()
for dim, kw in ((4, {}), (4, dict(has_alpha=True, srgb=True)), (6, dict(has_alpha=True, srgb=True)), (6, dict(is_normal_map=True)),
                (4, dict(axis_method=1))):
    img = (synth.synth_normal if kw.get("is_normal_map") else synth.synth_rgba)(517, 389, 77 + dim).cuda()
    out = A.encode_astc(img, A.encode_option(is4x4=dim == 4, is6x6=dim == 6, **kw))
    torch.cuda.synchronize()
    h.update(out.cpu().numpy().tobytes())
print("LIB", A.lib()._name)
print("SHA", h.hexdigest())
"""


def _probe(lib_path=None):
    env = dict(os.environ)
    if lib_path is not None:
        env["ASTC_B200_LIB"] = str(lib_path)
    else:
        env.pop("ASTC_B200_LIB", None)
    res = subprocess.run([sys.executable, "-c", _PROBE % str(ROOT)], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    fields = dict(l.split(" ", 1) for l in res.stdout.strip().splitlines() if " " in l)
    return fields["LIB"], fields["SHA"]


def test_library_builds_from_source_on_this_box():
    from astc_encoder_b200 import build as B
    try:
        B._nvcc()
    except RuntimeError as e:                              # a GPU box without the CUDA toolkit: nothing to prove here
        pytest.skip(str(e))
    lib = B.build_variant("boxbuild", [])               # every .cu / .cpp through nvcc -gencode arch=compute_100a,code=sm_100a
    try:
        assert lib.exists() and lib.stat().st_size > 1_000_000
        used, sha_fresh = _probe(lib)
        assert used.endswith("libastc_b200_boxbuild.so")
        _, sha_shipped = _probe(None)
        assert sha_fresh == sha_shipped
    finally:
        lib.unlink(missing_ok=True)
