"""Codec oracle against the reference's own quantisers (tests/golden/codecs.npz)."""
import os

import numpy as np

from oracle import codecs


def _g(golden_dir):
    with np.load(os.path.join(golden_dir, "codecs.npz")) as z:
        return {k: z[k] for k in z.files}


def test_quantisers_equal_reference(golden_dir):
    g = _g(golden_dir)
    x = g["x"]
    assert (codecs.quant16(x).astype(np.int64) == g["q16"]).all()
    assert (codecs.quant8(x).astype(np.int64) == g["q8"]).all()
    assert (codecs.quant4_codes(x).astype(np.int64) == g["q4"]).all()


def test_dequantisers_match_reference_python_within_rounding(golden_dir):
    """The C++ decoders (what the oracle restates) and reduce_precision.py's Python decoders
    are different expressions of the same map: equal to the script's printed precision."""
    g = _g(golden_dir)
    # 16-bit: script rounds to 6 decimals and uses (v/65000)*1.3-0.65; C++ uses v*0.00002-0.65
    assert np.abs(codecs.dequant16(g["q16"]).astype(np.float64) - g["d16"]).max() < 2e-6
    # 8-bit: script rounds to 3 decimals
    assert np.abs(codecs.dequant8(g["q8"]).astype(np.float64) - g["d8"]).max() < 6e-4
    # 4-bit: same table
    assert (codecs.LUT4[np.minimum(g["q4"], 14)] == g["d4"].astype(np.float32)).all()


def test_round_trip_error_bounds():
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.65, 0.65, 20000).astype(np.float32)
    assert np.abs(codecs.dequant16(codecs.quant16(x)) - x).max() <= 2.1e-5
    assert np.abs(codecs.dequant8(codecs.quant8(x)) - x).max() <= 1.0 / 254 + 1e-6
    t = x.reshape(-1, 16)
    p = codecs.quantize_table(t, 4)
    assert p.shape == (t.shape[0], 8)
    d = codecs.dequantize_rows(p, 4)
    assert d.shape == t.shape and set(np.unique(d)).issubset(set(codecs.LUT4.tolist()))
    assert (np.sign(d) == np.sign(np.where(np.abs(t) < 0.00025, np.sign(d), t))).all()


def test_product_host_codecs_equal_the_reference_on_the_golden_vector(golden_dir):
    """ev-store-dlrm_b200/codecs.py (what writes the backing-store rows the GPU decodes) against the reference's own
    quantisers on the golden vector -- dense range, out-of-range values, bucket edges, signed zeros -- and against the
    oracle's dequantisers on every code the quantisers can emit."""
    import importlib
    import os

    import numpy as np
    pc = importlib.import_module("ev-store-dlrm_b200").codecs
    from oracle import codecs as oc
    with np.load(os.path.join(golden_dir, "codecs.npz")) as z:
        g = {k: z[k] for k in z.files}
    x = g["x"].astype(np.float32)
    n = len(x) - len(x) % 2
    t = x[:n].reshape(-1, 2)
    assert np.array_equal(pc.encode_table(t, 16).reshape(-1).astype(np.int64), g["q16"][:n])
    assert np.array_equal(pc.encode_table(t, 8).reshape(-1).astype(np.int64), g["q8"][:n])
    packed = pc.encode_table(t, 4).reshape(-1)
    assert np.array_equal(np.stack([packed >> 4, packed & 15], axis=1).reshape(-1).astype(np.int64), np.minimum(g["q4"][:n], 15))
    for prec in (16, 8, 4):
        raw = pc.encode_table(t, prec)
        assert np.array_equal(raw, oc.quantize_table(t, prec))
        assert np.array_equal(pc.decode_rows(raw, prec), oc.dequantize_rows(raw, prec))
    # every 16-bit and 8-bit code, every nibble pair
    all16 = np.arange(65536, dtype=np.uint16).reshape(-1, 2)
    assert np.array_equal(pc.decode_rows(all16, 16), oc.dequantize_rows(all16, 16))
    all8 = np.arange(256, dtype=np.uint8).reshape(-1, 2)
    assert np.array_equal(pc.decode_rows(all8, 8), oc.dequantize_rows(all8, 8))
    assert np.array_equal(pc.decode_rows(all8, 4), oc.dequantize_rows(all8, 4))
