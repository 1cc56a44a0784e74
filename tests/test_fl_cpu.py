"""CPU tests of the FL runtime: config 0 of BASELINE.json (2 clients, 2-layer CNN, 28x28,
CKKS on CPU) end to end; encrypted FedAvg == plaintext FedAvg; fault injection; resume."""
import os

import torch

from hefl_b200.config import FLConfig
from hefl_b200.fl import FederatedRunner, simulate_clients
from hefl_b200.models import ParamPack, create_model


def _cfg(**kw):
    base = dict(model="cnn2", image_size=28, in_channels=1, num_classes=10, batch_size=8,
                local_epochs=2, steps_per_epoch=3, val_steps=1, clients=2, he_preset="n2048_l1",
                device="cpu", transport="loopback", dtype="fp32", nn_backend="cudnn")
    base.update(kw)
    return FLConfig(**base)


def test_two_client_fedavg_matches_plaintext():
    r = simulate_clients(_cfg(), rounds=1)
    assert r["contributors"] == 2
    assert r["max_abs_err"] < 1e-4
    assert torch.isfinite(r["encrypted_avg"]).all()


def test_packed_n4096_three_limbs_and_coeff_packing():
    r = simulate_clients(_cfg(he_preset="n4096_l3", packing="coeff"), rounds=1)
    assert r["max_abs_err"] < 1e-6


def test_dropped_client_uses_participation_mask():
    r = simulate_clients(_cfg(clients=3), rounds=1, drop_client=1)
    assert r["contributors"] == 2
    assert r["max_abs_err"] < 1e-4


def test_sequential_clients_quirk_flag_changes_result():
    a = simulate_clients(_cfg(), rounds=1)["encrypted_avg"]
    b = simulate_clients(_cfg(compat_sequential_clients=True), rounds=1)["encrypted_avg"]
    assert (a - b).abs().max() > 1e-6


def test_runner_round_and_resume(tmp_path):
    cfg = _cfg(rounds=1)
    run = FederatedRunner(cfg, device=torch.device("cpu"))
    rec = run.run_round(check=True)
    assert set(rec["stage_ms"]) == {"train", "encrypt", "aggregate", "decrypt"}
    ck = os.path.join(tmp_path, "ck.pt")
    run.save_checkpoint(ck)
    flat = run.pack.flat.clone()
    run2 = FederatedRunner(cfg, device=torch.device("cpu"))
    run2.load_checkpoint(ck)
    assert run2.round == 1 and torch.equal(run2.pack.flat, flat)


def test_keras_dict_roundtrip():
    m = create_model("medcnn", 3, 2, 256)
    p = ParamPack(m)
    d = p.to_keras_dict()
    assert list(d)[:4] == ["c_0_0", "c_0_1", "c_2_0", "c_2_1"] and "c_15_1" in d and len(d) == 18
    assert d["c_0_0"].shape == (3, 3, 3, 32) and d["c_13_0"].shape == (512, 128)
    before = p.flat.clone()
    p.flat.zero_()
    p.from_keras_dict(d)
    assert torch.equal(p.flat, before)


def test_debug_poison_between_rounds_keeps_results_identical():
    """Poisoning the ciphertext buffers after every round (SURVEY.md §5.2 debug mode) must not change
    any result: every round rewrites all the words it reads."""
    import torch
    from hefl_b200.config import FLConfig
    from hefl_b200.fl import FederatedRunner

    def run(poison):
        cfg = _cfg(local_epochs=1, steps_per_epoch=2, clients=1, seed=3, debug_poison=poison, deterministic_crypto=True)
        r = FederatedRunner(cfg, device=torch.device("cpu"))
        for _ in range(2):
            r.run_round(check=True)
        return r.pack.flat.clone(), getattr(r, "poisoned", 0), r

    a, na, _ = run(False)
    b, nb, rb = run(True)
    assert na == 0 and nb == 2
    assert torch.equal(a, b)


def test_fp8_flag_is_per_model_not_process_wide():
    """FLConfig(dtype="fp8") marks the 1x1 convolutions of THAT model; another trainer in the same process
    (bf16) is not affected, and on CPU the flag stays off (the fp8 GEMM path is CUDA-only)."""
    import torch
    from hefl_b200.fl.trainer import LocalTrainer
    from hefl_b200.models import ParamPack, create_model
    from hefl_b200.ops import fp8

    m = create_model("resnet50", num_classes=4)
    assert fp8.set_model_fp8(m, True) >= 30 and all(c.use_fp8 for c in m.modules() if isinstance(c, fp8.Conv1x1))
    cfg = _cfg(model="resnet18", image_size=32, in_channels=3, num_classes=4, dtype="fp32")
    net = create_model("resnet18", in_channels=3, num_classes=4)
    LocalTrainer(net, ParamPack(net), cfg, torch.device("cpu"))
    assert all(c.use_fp8 is False for c in net.modules() if isinstance(c, fp8.Conv1x1))
    assert fp8.ENABLE is False
    x = torch.randn(2, 3, 32, 32)
    assert net(x).shape == (2, 4)                 # fallback path is a plain convolution


def test_round_record_reports_ckks_precision_when_asked():
    import torch
    from hefl_b200.fl import FederatedRunner
    cfg = _cfg(local_epochs=1, steps_per_epoch=2, clients=1, transport="gloo", debug_precision=True)
    r = FederatedRunner(cfg, device=torch.device("cpu"))
    rec = r.run_round(check=True)
    assert rec["ckks_max_abs_err"] < 1e-6 and rec["ckks_precision_bits"] > 20
    cfg2 = _cfg(local_epochs=1, steps_per_epoch=2, clients=1, transport="gloo")
    rec2 = FederatedRunner(cfg2, device=torch.device("cpu")).run_round()
    assert "ckks_max_abs_err" not in rec2
