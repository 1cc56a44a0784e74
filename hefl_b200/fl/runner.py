"""Federated round loop: one client per rank (one process per GPU), encrypted FedAvg.

One round = what the reference's notebook cell 3 does once (N:233-272):
``train_clients`` (FLPyfhelin.py:179-198) -> ``export_encrypted_clients_weights`` (:242) ->
``aggregate_encrypted_weights`` (:366) -> ``decrypt_import_weights`` (:263), but

* clients are concurrent ranks, not iterations of a loop, and start every round from the same
  global model (true FedAvg; ``compat_sequential_clients`` reproduces quirk Q1 in the
  single-process simulation),
* there is an outer round loop (the reference runs exactly one, Q2),
* the pickle-file "network" and the server loop are replaced by the ciphertext all-reduce,
* averaging is a sum in ciphertext space with 1/K folded into the decode scale.

Roles: every rank holds the public context; the secret key is generated (OS entropy) and kept by
the key-holder rank alone (default rank 0), which owns no chunk of the fused all-reduce, decrypts
the aggregate and broadcasts the averaged plaintext model.
"""
from __future__ import annotations

import hashlib
import json
import math
import os
import secrets
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from ..config import FLConfig
from ..he.context import CKKSContext, CtBatch
from ..models import ParamPack, create_model
from ..parallel import LoopbackTransport, make_transport
from ..utils import DeviceTimer, JsonlLogger, StageTimes
from .data import BatchFeeder, SyntheticImageDataset, shard_range, split_train_val
from .trainer import LocalTrainer


class FederatedRunner:
    def __init__(self, cfg: FLConfig, rank: int = 0, world: int = 1, group=None,
                 device: Optional[torch.device] = None, samples_per_client: Optional[int] = None, dataset=None):
        self.cfg = cfg
        self.rank, self.world, self.group = rank, world, group
        self.device = device or torch.device(cfg.device if torch.cuda.is_available() or cfg.device == "cpu" else "cpu")
        torch.manual_seed(cfg.seed)                     # the reference seeds nothing (Q12)
        self.model = create_model(cfg.model, cfg.in_channels, cfg.num_classes, cfg.image_size).to(self.device)
        self.pack = ParamPack(self.model)
        self.trainer = LocalTrainer(self.model, self.pack, cfg, self.device)
        hp = cfg.he_params()
        self.ctx = CKKSContext(hp["n"], prime_bits=hp["prime_bits"], scale_bits=hp["scale_bits"],
                               device=self.device, sec=cfg.sec)
        # Trust model (the reference's roles: the aggregator works with get_pk only, FLPyfhelin.py:370; get_sk is
        # opened in decrypt alone, :284). The key holder generates the pair from OS entropy, keeps sk, and every
        # other rank only ever receives pk. It owns no chunk of the fused all-reduce, so it never loads a peer's
        # un-aggregated ciphertext; it decrypts the aggregate and broadcasts the averaged plaintext model.
        self.key_holder = int(cfg.key_holder) if world > 1 else -1
        if self.key_holder >= world:
            raise ValueError(f"key_holder {self.key_holder} is not a rank of this {world}-client federation")
        self.has_sk = world == 1 or self.key_holder < 0 or rank == self.key_holder
        self._setup_keys()
        self.n_ct = self.ctx.num_ct(self.pack.numel, cfg.packing)
        self.ct_numel = self.n_ct * 2 * self.ctx.L * self.ctx.n
        kind = cfg.transport
        if self.device.type == "cpu" and kind == "fused":
            kind = "gloo" if world > 1 else "loopback"
        self.transport = make_transport(kind, self.ctx, self.ct_numel, world=world, group=group,
                                        **({"algo": cfg.allreduce_algo, "timeout_s": cfg.timeout_s,
                                            "no_owner": self.key_holder} if kind == "fused" else {}))
        self._pair_seeds = self._agree_pair_seeds() if (cfg.pairwise_masks and world > 1) else None
        # data: IID contiguous shard of a synthetic set (FLPyfhelin.py:75-78), 90/10 split (:85)
        # or, with ``dataset`` (e.g. ImageFolderDataset of this client's shard), real images: 10 % validate
        if dataset is not None:
            self.dataset = dataset
            per = len(dataset)
            nval = int(per * 0.1)
        else:
            per = samples_per_client or (cfg.steps_per_epoch * cfg.batch_size + cfg.val_steps * cfg.batch_size)
            self.dataset = SyntheticImageDataset(per, cfg.image_size, cfg.in_channels, cfg.num_classes,
                                                 seed=cfg.seed + 17 * rank)
            nval = cfg.val_steps * cfg.batch_size
        self.train_feed = BatchFeeder(self.dataset, range(nval, per), cfg.batch_size, self.device,
                                      shuffle=True, seed=cfg.seed + rank)
        self.val_feed = BatchFeeder(self.dataset, range(0, nval), cfg.batch_size, self.device,
                                    shuffle=True, seed=cfg.seed + rank) if nval else None
        self.round = 0
        self.participating = True        # set False to sit a round out (needs cfg.allow_dropouts on every rank)
        self._active = None              # contributors of the current round when dropouts are allowed
        self.timer = DeviceTimer(self.device)
        self.log = JsonlLogger(cfg.log_jsonl, rank)
        self.global_flat = self.pack.flat.clone()
        self.history: List[Dict] = []

    # ------------------------------------------------------------------ keys and randomness
    def _setup_keys(self) -> None:
        cfg = self.cfg
        if cfg.deterministic_crypto:
            # bit-reproducible runs for tests: everything from the PUBLIC cfg.seed -- never in production
            sk, self.pk = self.ctx.keygen(seed=cfg.seed)
            self.sk = sk if self.has_sk else None
            return
        gen_rank = self.key_holder if self.key_holder >= 0 else 0
        if self.world == 1 or self.rank == gen_rank:
            sk, pk = self.ctx.keygen(seed=secrets.randbits(63))
        else:
            sk = torch.empty(self.ctx.L, self.ctx.n, dtype=torch.int64, device=self.device)
            pk = torch.empty(2, self.ctx.L, self.ctx.n, dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.broadcast(pk, src=gen_rank, group=self.group)
            if self.key_holder < 0:          # "every client can decrypt" mode: sk is handed out explicitly
                dist.broadcast(sk, src=gen_rank, group=self.group)
        self.pk = pk
        self.sk = sk if self.has_sk else None

    def _encrypt_seed(self) -> int:
        """Fresh randomness for every encryption call. Re-running a round (resume, retry) must never reuse
        (u, e0, e1) on different weights, and a seed derived from public values lets anybody strip the mask."""
        if self.cfg.deterministic_crypto:
            return (self.cfg.seed * 1_000_003 + self.round * 1009 + self.rank) & 0x7FFFFFFFFFFFFFFF
        return secrets.randbits(63)

    def _agree_pair_seeds(self):
        """One 63-bit seed per peer. X25519 key agreement over the process group: only the two ends of a
        pair can derive their seed. Returns (seeds, signs) with sign +1 towards higher ranks, -1 towards lower."""
        peers = [j for j in range(self.world) if j != self.rank]
        if self.cfg.deterministic_crypto:
            def ps(i, j):
                h = hashlib.sha256(f"hefl-pair-{self.cfg.seed}-{min(i, j)}-{max(i, j)}".encode()).digest()
                return int.from_bytes(h[:8], "little") >> 1
            seeds = [ps(self.rank, j) for j in peers]
        else:
            from cryptography.hazmat.primitives import serialization
            from cryptography.hazmat.primitives.asymmetric.x25519 import X25519PrivateKey, X25519PublicKey

            priv = X25519PrivateKey.generate()
            pub = priv.public_key().public_bytes(serialization.Encoding.Raw, serialization.PublicFormat.Raw)
            mine = torch.tensor(list(pub), dtype=torch.uint8, device=self.device)
            allp = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allp, mine, group=self.group)
            seeds = []
            for j in peers:
                shared = priv.exchange(X25519PublicKey.from_public_bytes(bytes(allp[j].cpu().tolist())))
                seeds.append(int.from_bytes(hashlib.sha256(b"hefl-pair" + shared).digest()[:8], "little") >> 1)
        signs = [1 if j > self.rank else -1 for j in peers]
        return seeds, signs

    def _mask(self, data: torch.Tensor, chunk: int = 0) -> None:
        if self._pair_seeds is None:
            return
        seeds, signs = self._pair_seeds
        self.ctx.ops.pairwise_mask_(data, seeds, signs, (self.round * 4096 + chunk) & 0xFFFFFFFF, self.ctx.L,
                                    self.ctx.logn, self.ctx.consts_cpu)

    # ------------------------------------------------------------------ stages
    def local_train(self, early_stopping: Optional[int] = None):
        with self.timer.stage("train"):
            return self.trainer.fit(self.train_feed, self.val_feed, self.cfg.local_epochs,
                                    early_stopping=early_stopping)

    def _contributors(self) -> float:
        """K of this round: the world size, or — with ``allow_dropouts`` — the number of ranks that take part
        (participation mask). A client that sits out still joins the collective with an
        encryption of zeros, so the fused kernel and its barriers are unchanged."""
        if self._active is not None:
            return float(self._active)
        return float(self.transport.contributors())

    def _agree_on_participants(self) -> None:
        self._active = None
        if not self.cfg.allow_dropouts or isinstance(self.transport, LoopbackTransport):
            return
        t = torch.tensor([1.0 if self.participating else 0.0], dtype=torch.float32, device=self.device)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        self._active = int(round(float(t)))

    def _my_update(self) -> torch.Tensor:
        return self.pack.flat if (self.participating or not self.cfg.allow_dropouts) else torch.zeros_like(self.pack.flat)

    def encrypt_update(self) -> CtBatch:
        self._agree_on_participants()
        with self.timer.stage("encrypt"):
            buf = self.transport.buffer(self.ct_numel)
            ct = self.ctx.encrypt(self._my_update(), self.pk, seed=self._encrypt_seed(), packing=self.cfg.packing,
                                  out=buf)
            self._mask(ct.data)
            return ct

    def aggregate(self, ct: CtBatch) -> CtBatch:
        with self.timer.stage("aggregate"):
            data = self.transport.allreduce(ct.data)
            return CtBatch(data, ct.scale, ct.nvals, ct.packing)

    def decrypt_apply(self, agg: CtBatch) -> None:
        with self.timer.stage("decrypt"):
            k = self._contributors()
            if k == 0:                                       # nobody took part: keep the current global model
                return
            if self.has_sk:
                avg = self.ctx.decrypt(agg, self.sk, divide_by=float(k))
            else:
                avg = torch.empty_like(self.pack.flat)
            if self.key_holder >= 0:
                avg = avg.to(self.pack.flat.dtype).contiguous()
                dist.broadcast(avg, src=self.key_holder, group=self.group)
            self.pack.load_flat(avg)
            if self.trainer.engine is not None:
                self.trainer.engine.after_restore()

    # ------------------------------------------------------------------ chunked, overlapped FedAvg
    def fedavg_pipelined(self, chunk_cts: int = 256) -> None:
        """encode+encrypt of chunk i+1, the all-reduce of chunk i and decrypt+decode of chunk i-1
        run concurrently on three streams (the scaling axis of this workload is
        ciphertext volume; ResNet-18 is 1.5 GB of ciphertext per client). Result is identical to
        the unchunked path: chunks are whole ciphertexts and every chunk uses its own seed offset."""
        ctx, dev = self.ctx, self.device
        vpc = ctx.values_per_ct(self.cfg.packing)
        n_ct = self.n_ct
        per_ct = 2 * ctx.L * ctx.n
        self._agree_on_participants()
        flat = self._my_update()
        out = torch.empty_like(flat)
        k = self._contributors()
        if k == 0:
            return
        seed = self._encrypt_seed()
        # the three streams live as long as the runner: the caching allocator keeps one pool per stream, so
        # fresh streams every round meant fresh cudaMallocs (and cudaFree stalls) every round
        if getattr(self, "_pipe_streams", None) is None:
            self._pipe_streams = tuple(torch.cuda.Stream(dev) for _ in range(3))
        s_enc, s_comm, s_dec = self._pipe_streams
        cur = torch.cuda.current_stream(dev)
        for st in (s_enc, s_comm, s_dec):
            st.wait_stream(cur)
        buf = self.transport.buffer(self.ct_numel)
        # The three stages only overlap if they fit on the chip together: encrypt / decrypt are persistent
        # one-CTA-per-SM kernels and the all-reduce is a copy engine that saturates NVLink from a few SMs, so the
        # SMs are partitioned for the duration of the pipeline (measured on ResNet-18, 8 GPUs: see profiles/).
        split = None
        if self.world > 1 and dev.type == "cuda" and self.transport.name == "fused":
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            env = os.environ.get("HEFL_PIPE_SPLIT")
            if env:
                split = tuple(int(v) for v in env.split(","))
            else:
                ar = 24
                dec = 28 if self.has_sk else 0
                split = (sms - ar - dec, dec, ar)
            if min(split[0], split[2]) >= 2 * ctx.L and split[0] + split[1] + split[2] <= sms:
                ctx.ops.set_he_cta_limits(0, split[0], split[1])
                saved_blocks, self.transport.blocks = self.transport.blocks, split[2]
            else:
                split = None
        timeline = [] if os.environ.get("HEFL_TIMELINE") else None

        def mark(stream):
            e = torch.cuda.Event(enable_timing=timeline is not None)
            e.record(stream)
            return e

        with self.timer.stage("fedavg_pipelined"):
            ev_enc, ev_comm = [], []
            chunks = [(c0, min(n_ct, c0 + chunk_cts)) for c0 in range(0, n_ct, chunk_cts)]
            t_origin = mark(cur) if timeline is not None else None
            for ci, (c0, c1) in enumerate(chunks):
                vals = flat[c0 * vpc: min(flat.numel(), c1 * vpc)]
                view = buf[c0 * per_ct: c1 * per_ct]
                rec = {"chunk": ci}
                with torch.cuda.stream(s_enc):
                    if timeline is not None:
                        rec["enc0"] = mark(s_enc)
                    ct = ctx.encrypt(vals, self.pk, seed=seed, packing=self.cfg.packing, out=view, ct_offset=c0)
                    self._mask(ct.data, chunk=ci)
                    ev_enc.append(mark(s_enc))
                with torch.cuda.stream(s_comm):
                    s_comm.wait_event(ev_enc[ci])
                    if timeline is not None:
                        rec["ar0"] = mark(s_comm)
                    data = self.transport.allreduce(ct.data)
                    ev_comm.append(mark(s_comm))
                with torch.cuda.stream(s_dec):
                    s_dec.wait_event(ev_comm[ci])
                    if timeline is not None:
                        rec["dec0"] = mark(s_dec)
                    if self.has_sk:
                        avg = ctx.decrypt(CtBatch(data, ct.scale, ct.nvals, ct.packing), self.sk, divide_by=k)
                        out[c0 * vpc: c0 * vpc + avg.numel()].copy_(avg)
                    if timeline is not None:
                        rec["dec1"] = mark(s_dec)
                if timeline is not None:
                    rec["enc1"], rec["ar1"] = ev_enc[ci], ev_comm[ci]
                    timeline.append(rec)
            cur.wait_stream(s_dec)
            cur.wait_stream(s_comm)
            cur.wait_stream(s_enc)
            if split is not None:
                ctx.ops.set_he_cta_limits(0, 0, 0)
                self.transport.blocks = saved_blocks
            if timeline is not None:
                torch.cuda.synchronize(dev)
                rows = [{"chunk": r["chunk"], **{k: round(t_origin.elapsed_time(r[k]), 4) for k in r if k != "chunk"}}
                        for r in timeline]
                if self.rank == 0:
                    with open(os.environ["HEFL_TIMELINE"], "w") as f:
                        json.dump({"world": self.world, "n_ct": n_ct, "chunk_cts": chunk_cts, "sm_split_enc_dec_ar": split,
                                   "unit": "ms since the start of the FedAvg stage", "chunks": rows}, f, indent=1)
            if self.key_holder >= 0:
                dist.broadcast(out, src=self.key_holder, group=self.group)
            self.pack.load_flat(out)

    def guard_finite(self) -> None:
        """NaN/Inf guard on decoded weights (failure detection)."""
        if not bool(torch.isfinite(self.pack.flat).all()):
            raise FloatingPointError(f"round {self.round}: decrypted weights are not finite")

    # ------------------------------------------------------------------ round
    def run_round(self, check: bool = False, pipelined: Optional[bool] = None) -> Dict:
        hist = self.local_train()
        plain_mean = None
        if self.cfg.debug_precision and not isinstance(self.transport, LoopbackTransport):
            # plaintext FedAvg oracle (SURVEY.md §5.5 "CKKS precision: max abs error vs plaintext FedAvg")
            # the oracle averages what the encrypted path averages: sitting-out clients contribute zeros and
            # K is the agreed participant count
            self._agree_on_participants()
            plain_mean = self._my_update().detach().clone().double()
            if self.world > 1:
                dist.all_reduce(plain_mean, op=dist.ReduceOp.SUM, group=self.group)
            kk = self._contributors()
            plain_mean = plain_mean / kk if kk > 0 else None
        if pipelined is None:
            pipelined = self.device.type == "cuda" and self.n_ct > 512 and self.transport.name in ("fused", "nccl", "gloo")
        if pipelined:
            self.fedavg_pipelined()
            if self.trainer.engine is not None:
                self.trainer.engine.after_restore()
        else:
            ct = self.encrypt_update()
            agg = self.aggregate(ct)
            self.decrypt_apply(agg)
        if self.cfg.debug_poison:
            self.poison_buffers()
        times = self.timer.resolve()
        if hasattr(self.transport, "check_status"):
            self.transport.check_status()
        if check:
            self.guard_finite()
        rec = {"round": self.round, "rank": self.rank, "stage_ms": times,
               "loss": hist[-1].loss if hist else None, "accuracy": hist[-1].accuracy if hist else None,
               "transport": self.transport.name, "n_ct": self.n_ct, "ct_bytes": self.ct_numel * 8}
        if plain_mean is not None:
            err = float((self.pack.flat.double() - plain_mean).abs().max())
            rec["ckks_max_abs_err"] = err
            rec["ckks_precision_bits"] = float(-math.log2(err)) if err > 0 else float("inf")
        self.history.append(rec)
        self.round += 1
        return rec

    POISON = 0x7FF8DEADBEEF7FF8          # > every q_l: a stale word can never pass for a residue

    def poison_buffers(self) -> None:
        """Debug mode (``debug_poison`` / HEFL_DEBUG_POISON=1): after the round's result is
        installed, every buffer that carried ciphertext words is overwritten, so a later round that reads
        something it did not write this round (stale tile, missed barrier) decodes to garbage and trips
        ``guard_finite`` / the plaintext cross-check instead of silently reusing last round's data."""
        buf = self.transport.buffer(self.ct_numel)
        buf.fill_(self.POISON)
        out = getattr(self.transport, "out", None)
        if isinstance(out, torch.Tensor):
            out.fill_(self.POISON)
        self.poisoned = getattr(self, "poisoned", 0) + 1

    def run(self, rounds: Optional[int] = None) -> List[Dict]:
        out = []
        for _ in range(rounds or self.cfg.rounds):
            rec = self.run_round(check=True)
            rec["stage_ms_max"] = StageTimes.max_over_ranks(rec["stage_ms"], self.device, self.group)
            self.log.write(rec)
            out.append(rec)
        return out

    # ------------------------------------------------------------------ evaluation (notebook N:262-270)
    @torch.no_grad()
    def evaluate(self, dataset, batch_size: Optional[int] = None) -> Dict[str, float]:
        """Weighted precision / recall / F1 and accuracy of the current global model on ``dataset`` — the four
        numbers the reference's notebook reports after aggregation (sklearn ``average='weighted'``, N:267-270).
        Runs the plain PyTorch forward of the model (any backend / device); not a hot path."""
        bs = batch_size or self.cfg.batch_size
        self.model.eval()
        preds = []
        for i in range(0, len(dataset), bs):
            x = dataset.images[i:i + bs].to(self.device, non_blocking=True)
            x = x.permute(0, 3, 1, 2).float() * (1.0 / 255.0)
            preds.append(self.model(x).float().argmax(1).cpu())
        self.model.train()
        return classification_metrics(dataset.labels.cpu(), torch.cat(preds), int(getattr(dataset, "classes", self.cfg.num_classes)))

    # ------------------------------------------------------------------ checkpoint / resume
    def save_checkpoint(self, path: str) -> None:
        """Round index + global model (rank 0 -> ``path``) and, PER RANK, optimiser moments, LR state and RNG
        (``path.rank{r}``) so that a resumed multi-round run equals an uninterrupted one."""
        if self.rank == 0:
            torch.save({"round": self.round, "flat": self.pack.flat.detach().cpu(), "config": self.cfg.to_json()}, path)
        torch.save({"round": self.round, "opt": self.trainer.state_dict(), "rng": torch.get_rng_state(),
                    "cuda_rng": torch.cuda.get_rng_state(self.device) if self.device.type == "cuda" else None},
                   f"{path}.rank{self.rank}")

    def load_checkpoint(self, path: str) -> None:
        ck = torch.load(path, map_location="cpu", weights_only=True)      # tensors and plain containers only
        self.round = int(ck["round"])
        self.pack.load_flat(ck["flat"])
        mine = f"{path}.rank{self.rank}"
        if os.path.exists(mine):
            st = torch.load(mine, map_location="cpu", weights_only=True)
            self.trainer.load_state_dict(st["opt"])
            torch.set_rng_state(st["rng"])
            if st.get("cuda_rng") is not None and self.device.type == "cuda":
                torch.cuda.set_rng_state(st["cuda_rng"], self.device)
        else:                                       # no state of this rank (e.g. a different world size): start it fresh
            self.trainer.reset_optimizer()
        if self.trainer.engine is not None:
            self.trainer.engine.after_restore()


def classification_metrics(y_true: torch.Tensor, y_pred: torch.Tensor, num_classes: int) -> Dict[str, float]:
    """precision / recall / f1 (support-weighted, zero where undefined — sklearn's ``average='weighted',
    zero_division=0``) and accuracy."""
    y_true, y_pred = y_true.long().flatten(), y_pred.long().flatten()
    n = max(1, y_true.numel())
    num_classes = max(num_classes, int(max(y_true.max(), y_pred.max())) + 1) if y_true.numel() else num_classes
    p_w = r_w = f_w = 0.0
    for c in range(num_classes):
        tp = float(((y_pred == c) & (y_true == c)).sum())
        fp = float(((y_pred == c) & (y_true != c)).sum())
        fn = float(((y_pred != c) & (y_true == c)).sum())
        prec = tp / (tp + fp) if tp + fp > 0 else 0.0
        rec = tp / (tp + fn) if tp + fn > 0 else 0.0
        f1 = 2 * prec * rec / (prec + rec) if prec + rec > 0 else 0.0
        w = (tp + fn) / n
        p_w, r_w, f_w = p_w + w * prec, r_w + w * rec, f_w + w * f1
    return {"precision": p_w, "recall": r_w, "f1": f_w, "accuracy": float((y_true == y_pred).sum()) / n}


def simulate_clients(cfg: FLConfig, device: Optional[torch.device] = None, rounds: int = 1,
                     drop_client: Optional[int] = None) -> Dict:
    """Single-process simulation of ``cfg.clients`` clients through the loopback transport — the
    reference's own structure (clients = loop iterations, FLPyfhelin.py:184) and our fake
    backend for tests. Returns the encrypted-FedAvg result and the plaintext FedAvg oracle."""
    device = device or torch.device("cpu")
    torch.manual_seed(cfg.seed)
    hp = cfg.he_params()
    ctx = CKKSContext(hp["n"], prime_bits=hp["prime_bits"], scale_bits=hp["scale_bits"], device=device,
                      sec=cfg.sec)
    sk, pk = ctx.keygen(seed=cfg.seed)
    model = create_model(cfg.model, cfg.in_channels, cfg.num_classes, cfg.image_size).to(device)
    pack = ParamPack(model)
    trainer = LocalTrainer(model, pack, cfg, device, backend="cudnn", use_graph=False)
    K = cfg.clients
    total = K * (cfg.steps_per_epoch + cfg.val_steps) * cfg.batch_size
    ds = SyntheticImageDataset(total, cfg.image_size, cfg.in_channels, cfg.num_classes, seed=cfg.seed, pin=False)
    lb = LoopbackTransport(ctx, K)
    global_flat = pack.flat.clone()
    result = {}
    for rnd in range(rounds):
        lb.reset()
        plain = []
        for i in range(K):
            if not cfg.compat_sequential_clients:
                pack.load_flat(global_flat)          # true FedAvg: same starting point
                trainer.reset_optimizer()
            s, e = shard_range(total, i, K)
            tr, va = split_train_val(s, e)
            tf = BatchFeeder(ds, tr, cfg.batch_size, device, seed=cfg.seed + i)
            vf = BatchFeeder(ds, va, cfg.batch_size, device, seed=cfg.seed + i) if len(va) else None
            trainer.fit(tf, vf, cfg.local_epochs)
            plain.append(pack.flat.clone())
            ct = ctx.encrypt(pack.flat, pk, seed=cfg.seed * 7919 + rnd * 131 + i, packing=cfg.packing)
            if drop_client is not None and i == drop_client:
                lb.drop(i)
            else:
                lb.contribute(i, ct.data)
        agg = CtBatch(lb.reduce(), ctx.scale, pack.numel, cfg.packing)
        k = lb.contributors()
        avg = ctx.decrypt(agg, sk, divide_by=float(k))
        kept = [p for i, p in enumerate(plain) if not (drop_client is not None and i == drop_client)]
        oracle = torch.stack(kept).mean(0)
        global_flat = avg.clone()
        pack.load_flat(global_flat)
        result = {"encrypted_avg": avg, "plain_avg": oracle, "contributors": k,
                  "max_abs_err": float((avg - oracle).abs().max())}
    result["model"] = model
    result["pack"] = pack
    result["ctx"] = ctx
    return result
