"""CPU: the C-ABI library loads and exports every symbol include/uahn.h declares; host-side logic."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cuahn_vio_b200 import build
    build.build()
    from cuahn_vio_b200 import api
    return api.load_library()


def test_header_symbols_exported(lib):
    from cuahn_vio_b200 import api
    declared = set()
    for name, symbols in (("uahn.h", api.EXPORTED_SYMBOLS), ("uahn_ekf.h", api.EKF_SYMBOLS),
                          ("uahn_preproc.h", api.PREPROC_SYMBOLS)):
        hdr = open(os.path.join(ROOT, "include", name)).read()
        found = set(re.findall(r"\b(uahn_[a-z0-9_]+)\s*\(", hdr))
        assert found == set(symbols), (name, found ^ set(symbols))
        declared |= found
    for name in declared:
        assert hasattr(lib, name), name


def test_create_fails_loudly_without_gpu_or_weights(lib):
    from cuahn_vio_b200 import api
    with pytest.raises(api.UahnError):
        api.Uahn("/nonexistent/weights.bin")


def test_philox_masks_host(lib):
    from cuahn_vio_b200 import api
    a = api.philox_keep_masks(7, 3)
    b = api.philox_keep_masks(7, 3)
    c = api.philox_keep_masks(7, 4)
    assert a.shape == (2, 16, 5376) and np.array_equal(a, b) and not np.array_equal(a, c)
    keep = a.mean()
    assert abs(keep - 0.95) < 0.003          # Bernoulli(0.95) keep, 172k draws
    assert set(np.unique(a)) == {0, 1}


def test_weight_export_roundtrip(tmp_path, synth_sd):
    from cuahn_vio_b200 import weights, synthetic
    p = str(tmp_path / "w.bin")
    weights.export_state_dict(synth_sd, p)
    raw = open(p, "rb").read()
    assert raw[:8] == b"UAHNWTS1" and struct.unpack("<I", raw[8:12])[0] == 54
    total = sum(int(np.prod(s)) for s in synthetic.state_dict_schema().values())
    assert total == 6541312
    assert len(raw) > total * 4
    bad = dict(synth_sd)
    bad.pop(next(iter(bad)))
    with pytest.raises(KeyError):
        weights.export_state_dict(bad, p)
