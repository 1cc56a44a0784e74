"""API-compat tests: Pyfhel 2.3.1 surface, pickle/context re-attachment,
and the notebook's cell-3 sequence end to end through the 19 FLPyfhelin names."""
import os
import pickle

import numpy as np
import pytest

from hefl_b200.compat import PyCtxt, Pyfhel
from hefl_b200.he import oracle


def test_fractional_encoder_matches_oracle():
    import torch

    from hefl_b200.he.bfv import BFVFracContext

    ctx = BFVFracContext(m=1024)
    vals = [0.0, 1.0, -1.0, 0.5, -0.375, 3.1415926, -123.456, 1e-5, 2 ** 20 + 0.25]
    msg = ctx.encode(torch.tensor(vals, dtype=torch.float64))
    for v, row in zip(vals, msg.tolist()):
        assert row == oracle.frac_encode(v, 1024)
        assert abs(oracle.frac_decode(row) - v) <= 2 ** -32 * max(1.0, abs(v)) + 2 ** -32
    back = ctx.decode(msg)
    assert np.allclose(back.numpy(), vals, atol=2 ** -31)


def test_pyfhel_231_surface_and_repr():
    HE = Pyfhel()
    HE.contextGen(p=65537, sec=128, m=1024)
    HE.keyGen(seed=1)
    rep = repr(HE)
    assert "pk:Y, sk:Y, rtk:-, rlk:-" in rep
    assert "contx(p=65537, m=1024, base=2, sec=128, dig=64i.32f, batch=False)" in rep
    c = HE.encryptFrac(0.123456)
    assert isinstance(c, PyCtxt)
    assert abs(HE.decryptFrac(c) - 0.123456) < 1e-8
    d = HE.encryptFrac(np.float32(-2.5))
    s = c + d
    assert abs(HE.decryptFrac(s) - (0.123456 - 2.5)) < 1e-6
    assert abs(HE.decryptFrac(c + 0) - 0.123456) < 1e-8            # first-client path (:380-381)
    assert abs(HE.decryptFrac(0 + c) - 0.123456) < 1e-8
    assert abs(HE.decryptFrac(s * 0.5) - (0.123456 - 2.5) / 2) < 1e-6  # ct * float (:385)
    assert HE.noiseLevel(c) > 0


def test_pyfhel_3x_keyword_and_more_clients_need_m2048():
    HE = Pyfhel()
    HE.contextGen(p=65537, sec=128, n=2048)        # README R:7: `m` became `n`
    HE.keyGen(seed=2)
    vals = [0.3, -0.7, 1.25, 0.01, 2.0]
    cts = [HE.encryptFrac(v) for v in vals]
    acc = cts[0] + 0
    for c in cts[1:]:
        acc = c + acc
    avg = acc * (1 / len(vals))
    assert abs(HE.decryptFrac(avg) - np.mean(vals)) < 1e-6


def test_pickle_drops_context_and_keys_then_rehydrates(tmp_path):
    HE = Pyfhel()
    HE.contextGen(p=65537, sec=128, m=1024)
    HE.keyGen(seed=3)
    con, pk, sk = HE.to_bytes_context(), HE.to_bytes_publicKey(), HE.to_bytes_secretKey()
    ct = HE.encryptFrac(4.75)
    blob = pickle.dumps({"key": HE, "val": np.array([ct], dtype=object)}, protocol=pickle.HIGHEST_PROTOCOL)
    assert sk[1:50] not in blob                    # secret key never travels with the object (Q11)
    back = pickle.loads(blob)
    HE2, ct2 = back["key"], back["val"][0]
    assert "pk:-, sk:-" in repr(HE2) and "contx(-)" in repr(HE2)
    assert ct2._pyfhel is None
    with pytest.raises(RuntimeError):
        _ = ct2 + ct2                              # detached ciphertexts cannot compute
    HE2.from_bytes_context(con)
    HE2.from_bytes_publicKey(pk)
    ct2._pyfhel = HE2                              # FLPyfhelin.py:320-321
    twice = ct2 + ct2
    with pytest.raises(RuntimeError):
        HE2.decryptFrac(twice)                     # public context cannot decrypt
    HE2.from_bytes_secretKey(sk)
    assert abs(HE2.decryptFrac(twice) - 9.5) < 1e-6


def test_ckks_mode_for_config0():
    HE = Pyfhel()
    HE.contextGen(scheme="CKKS", n=4096, scale_bits=40, qi_sizes=[36, 36, 37])
    HE.keyGen(seed=4)
    HE.relinKeyGen(bitCount=12)
    a = np.linspace(-1, 1, 1000)
    b = np.linspace(0.5, 1.5, 1000)
    ca, cb = HE.encryptFrac(a), HE.encryptFrac(b)
    assert np.allclose(HE.decryptFrac(ca + cb), a + b, atol=1e-5)
    assert np.allclose(HE.decryptFrac(ca * 0.25), a * 0.25, atol=1e-4)
    assert np.allclose(HE.decryptFrac(ca * cb), a * b, atol=1e-2)


def test_notebook_cell3_sequence_end_to_end(tmp_path, monkeypatch):
    """2 clients, 2-layer CNN, synthetic 28x28 images, every stage through the reference's names."""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "examples"))
    import encrypted_fl_main as drv

    monkeypatch.setattr(sys, "argv", ["encrypted_fl_main.py", "--workdir", str(tmp_path), "--synthetic", "28",
                                      "--clients", "2", "--epochs", "1", "--m", "2048",
                                      "--train-images", "64", "--test-images", "16"])
    cwd = os.getcwd()
    try:
        metrics, times = drv.main()
    finally:
        os.chdir(cwd)
    for f in ["publickey.pickle", "privatekey.pickle", "main_model.hdf5", "agg_model.hdf5", "weights/weights1.npy",
              "weights/weights2.npy", "weights/client_1.pickle", "weights/client_2.pickle",
              "weights/aggregated.pickle", "plainweights.pickle"]:
        assert os.path.exists(os.path.join(tmp_path, f)), f
    assert metrics.shape == (4, 1) and 0.0 <= float(metrics.loc["accuracy"].iloc[0]) <= 1.0
    # encrypted FedAvg == plaintext FedAvg of the two clients' weight files
    w1 = np.load(os.path.join(tmp_path, "weights/weights1.npy"), allow_pickle=True)
    w2 = np.load(os.path.join(tmp_path, "weights/weights2.npy"), allow_pickle=True)
    os.chdir(tmp_path)
    try:
        from hefl_b200.compat import FLPyfhelin as FL

        agg = FL.load_model("agg_model.hdf5").get_weights()
        with open("weights/client_1.pickle", "rb") as h:
            d = pickle.load(h)
        assert set(d) == {"key", "val"} and list(d["val"])[0] == "c_0_0"
        assert d["val"]["c_0_0"].shape == (3, 3, 3, 8) and d["val"]["c_0_0"].dtype == object
    finally:
        os.chdir(cwd)
    for a, x, y in zip(agg, w1, w2):
        assert np.allclose(a, (x + y) / 2, atol=1e-6)


def test_a_real_keras_hdf5_file_is_recognised_and_reported(tmp_path):
    """Files that carry the HDF5 signature go to the h5py reader; without h5py the error says what is missing
    instead of failing inside torch.load."""
    import importlib.util

    import pytest

    from hefl_b200.compat import keras_like

    p = tmp_path / "main_model.hdf5"
    p.write_bytes(keras_like.HDF5_MAGIC + b"\x00" * 64)
    if importlib.util.find_spec("h5py") is None:
        with pytest.raises(RuntimeError, match="h5py"):
            keras_like.load_model(str(p))
