"""Randomness, key-handling and masking properties (constant encryption seeds, keys derived
from the public config seed). CPU only."""
import pickle

import torch

from hefl_b200 import _ext
from hefl_b200.config import FLConfig
from hefl_b200.fl import FederatedRunner
from hefl_b200.he.context import CKKSContext

ops = _ext.ops()


def _cfg(**kw):
    base = dict(model="cnn2", image_size=28, in_channels=1, num_classes=10, batch_size=8, local_epochs=1,
                steps_per_epoch=1, val_steps=0, clients=1, he_preset="n2048_l1", nn_backend="cudnn", dtype="fp32",
                transport="loopback", device="cpu")
    base.update(kw)
    return FLConfig(**base)


def test_two_rehydrated_public_key_holders_never_share_encryption_randomness(tmp_path, monkeypatch):
    """What the reference's clients do: unpickle publickey.pickle, from_bytes_context / from_bytes_publicKey, encrypt.
    Two such instances must produce different ciphertexts for the same value (fresh OS-entropy seed per instance),
    and c1 must differ between two plaintexts."""
    from hefl_b200.compat.pyfhel_shim import Pyfhel

    he = Pyfhel()
    he.contextGen(p=65537, sec=128, m=2048)
    he.keyGen()
    blob = pickle.dumps({"HE": he, "con": he.to_bytes_context(), "pk": he.to_bytes_publicKey()})

    def client():
        k = pickle.loads(blob)
        h = k["HE"]
        h.from_bytes_context(k["con"])
        h.from_bytes_publicKey(k["pk"])
        return h

    a, b = client(), client()
    ca, cb = a.encryptFrac(0.25), b.encryptFrac(0.25)
    assert not torch.equal(ca._data, cb._data)
    c2 = a.encryptFrac(0.5)
    assert not torch.equal(ca._data[..., 1, :, :] if ca._data.dim() > 3 else ca._data[1], c2._data[..., 1, :, :] if c2._data.dim() > 3 else c2._data[1])
    assert abs(he.decryptFrac(ca) - 0.25) < 1e-6 and abs(he.decryptFrac(cb) - 0.25) < 1e-6
    # two independent key generations differ as well (no global-RNG or constant seed behind keyGen)
    h2 = Pyfhel()
    h2.contextGen(p=65537, sec=128, m=2048)
    h2.keyGen()
    assert he.to_bytes_publicKey() != h2.to_bytes_publicKey()


def test_runner_keys_and_encryption_seeds_come_from_entropy_unless_asked_otherwise():
    dev = torch.device("cpu")
    r1, r2 = FederatedRunner(_cfg(), device=dev), FederatedRunner(_cfg(), device=dev)
    assert not torch.equal(r1.pk, r2.pk) and not torch.equal(r1.sk, r2.sk)       # same cfg.seed, different keys
    assert r1._encrypt_seed() != r1._encrypt_seed()                              # fresh per call, even in the same round
    d1, d2 = (FederatedRunner(_cfg(deterministic_crypto=True), device=dev) for _ in range(2))
    assert torch.equal(d1.pk, d2.pk) and d1._encrypt_seed() == d2._encrypt_seed()


def test_pairwise_masks_cancel_in_the_sum_and_hide_each_addend():
    ctx = CKKSContext(2048, prime_bits=(54, 54), scale_bits=40, enforce_security=False)
    P, C = 4, 3
    gen = torch.Generator().manual_seed(0)
    data = [torch.stack([torch.randint(0, q, (C, 2, 2048), generator=gen, dtype=torch.int64) for q in ctx.primes], dim=2).contiguous()
            for _ in range(P)]
    seeds = {(i, j): 1000003 * (i + 1) + 97 * (j + 1) for i in range(P) for j in range(i + 1, P)}
    masked = []
    for i in range(P):
        peers = [j for j in range(P) if j != i]
        m = data[i].clone()
        ops.pairwise_mask_(m, [seeds[(min(i, j), max(i, j))] for j in peers], [1 if j > i else -1 for j in peers], 7,
                           ctx.L, ctx.logn, ctx.consts_cpu)
        assert (m != data[i]).float().mean() > 0.99            # every word is re-randomised
        for l, q in enumerate(ctx.primes):
            assert int(m[:, :, l].max()) < q and int(m[:, :, l].min()) >= 0
        masked.append(m)
    plain = torch.empty_like(data[0])
    ops.local_sum_modq(data, plain, ctx.L, ctx.logn, ctx.consts)
    hidden = torch.empty_like(data[0])
    ops.local_sum_modq(masked, hidden, ctx.L, ctx.logn, ctx.consts)
    assert torch.equal(plain, hidden)
    # a different round re-keys the stream
    m2 = data[0].clone()
    ops.pairwise_mask_(m2, [seeds[(0, j)] for j in (1, 2, 3)], [1, 1, 1], 8, ctx.L, ctx.logn, ctx.consts_cpu)
    assert not torch.equal(m2, masked[0])
