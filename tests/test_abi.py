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


def test_philox_masks_known_answer():
    """Pins the mask generator's definition (Philox4x32-7 counter layout + alias table): a recorded run stays replayable
    across versions of the library.  64-bit seed and pair index exercise the high counter words."""
    import hashlib
    from cuahn_vio_b200 import api
    for (seed, pair), (digest, kept) in {
        (7, 3): ("ea53fb60db8ab3ba7650fb29efb7c9034d78439220bf23631dffb5baaac1801c", 163511),
        (2 ** 40 + 5, 2 ** 33 + 1): ("57a19ba8654dc13860f7881105e8cc5e072703d0600462a54cbb9691df5b5004", 163313),
    }.items():
        m = api.philox_keep_masks(seed, pair)
        assert int(m.sum()) == kept
        assert hashlib.sha256(np.packbits(m).tobytes()).hexdigest() == digest


def test_philox_masks_are_iid_bernoulli():
    """The keep bytes come from an alias table over the 256 byte patterns (one 32-bit uniform per 8 units): the units must
    still be independent Bernoulli(0.95) — keep rate, drops per byte ~ Binomial(8, 0.05), uniform drop position."""
    from math import comb
    from cuahn_vio_b200 import api
    m = np.stack([api.philox_keep_masks(11, i) for i in range(20)])[..., :5120]       # first dropout, reference order
    kp = np.arange(5120)
    ref = (kp & 255) * 20 + (kp >> 8)                                                  # kernel order -> reference order
    drop = 1 - m.reshape(-1, 5120)[:, ref].astype(np.float64)                          # [640 rows, 5120] in kernel order
    n = drop.size
    assert abs(drop.mean() - 0.05) < 4 * (0.05 * 0.95 / n) ** 0.5
    per_byte = drop.reshape(-1, 8).sum(1).astype(int)
    nb = per_byte.size
    freq = np.bincount(per_byte, minlength=9) / nb
    for k in range(5):
        e = comb(8, k) * 0.05 ** k * 0.95 ** (8 - k)
        assert abs(freq[k] - e) < 5 * (e * (1 - e) / nb) ** 0.5 + 1e-7, (k, freq[k], e)
    single = drop.reshape(-1, 8)[per_byte == 1].mean(0)
    assert np.abs(single - 0.125).max() < 0.004
    has = drop.reshape(drop.shape[0], 640, 8).sum(2) > 0                               # neighbouring bytes independent
    assert abs((has[:, :-1] & has[:, 1:]).mean() - has.mean() ** 2) < 0.003


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
