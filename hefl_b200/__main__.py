"""Command-line launcher of the product path: multi-round encrypted federated training, one client per
process (one process per GPU under ``torchrun``; a single process on CPU or one GPU is a 1-client federation).

    # 8 clients on one box, 3 rounds of the reference's medical-CNN recipe (10 local epochs each)
    torchrun --nproc-per-node 8 -m hefl_b200 --rounds 3 --log-jsonl run.jsonl --checkpoint ckpt.pt
    # CPU plumbing config (BASELINE.json configs[0])
    python -m hefl_b200 --device cpu --model cnn2 --image-size 28 --in-channels 1 --num-classes 10 \\
        --he-preset n2048_l1 --nn-backend cudnn --dtype fp32 --rounds 2

Every ``FLConfig`` field is a flag (``--local-epochs 10``) or an environment variable
(``HEFL_LOCAL_EPOCHS=10``); the reference hard-codes all of them (FLPyfhelin.py:31-36, notebook N:24-32).
Extra flags: ``--data-dir DIR`` (the reference's ``image/Train`` layout ``DIR/<label>/<file>``: every rank decodes
its own IID contiguous shard, 10 % of it validates; without it a synthetic set of the configured shape is used),
``--test-dir DIR`` (same layout; after the last round rank 0 prints the reference's four metrics — weighted
precision / recall / F1 and accuracy, notebook N:267-270 — of the aggregated model),
``--checkpoint PATH`` (written by rank 0 after every round, resumed from if it exists),
``--simulate`` (all ``--clients`` in this one process through the loopback transport: the reference's own
structure, clients as loop iterations, FLPyfhelin.py:184).
"""
from __future__ import annotations

import argparse
import json
import os
import sys


def main(argv=None) -> int:
    import torch
    import torch.distributed as dist

    from .config import FLConfig
    from .fl import FederatedRunner, simulate_clients
    from .utils import StageTimes

    extra = argparse.ArgumentParser(add_help=False)
    extra.add_argument("--checkpoint", default=None)
    extra.add_argument("--data-dir", default=None)
    extra.add_argument("--test-dir", default=None)
    extra.add_argument("--stream-decode", action="store_true",
                       help="decode images per batch (flow_from_dataframe semantics) instead of the whole shard up front")
    extra.add_argument("--simulate", action="store_true")
    extra.add_argument("-h", "--help", action="store_true")
    ns, rest = extra.parse_known_args(argv)
    if ns.help:
        print(__doc__)
        print("FLConfig fields:", ", ".join(f.name for f in __import__("dataclasses").fields(FLConfig)))
        return 0
    cfg = FLConfig.from_args(rest)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = cfg.device == "cuda" and torch.cuda.is_available()
    if cfg.device == "cuda" and not use_cuda:
        print("hefl_b200: no CUDA device, falling back to --device cpu --nn-backend cudnn --dtype fp32", file=sys.stderr)
        cfg.device, cfg.nn_backend, cfg.dtype = "cpu", "cudnn", "fp32"
    device = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)

    if ns.simulate:
        res = simulate_clients(cfg, device=device, rounds=cfg.rounds)
        err = float(res["max_abs_err"]) if "max_abs_err" in res else None
        print(json.dumps({"mode": "simulate", "clients": cfg.clients, "rounds": cfg.rounds, "max_abs_err_vs_plaintext": err}))
        return 0

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if use_cuda:
            dist.init_process_group("nccl", device_id=device)
        else:
            dist.init_process_group("gloo")
    cfg.clients = world
    dataset = None
    if ns.data_dir:
        from .fl.data import ImageFolderDataset

        dataset = ImageFolderDataset(ns.data_dir, cfg.image_size, cfg.in_channels, index=rank, num_clients=world,
                                     shuffle_seed=cfg.seed, stream=ns.stream_decode)
        cfg.num_classes = max(cfg.num_classes, dataset.classes)
    run = FederatedRunner(cfg, rank=rank, world=world, device=device, dataset=dataset)
    if ns.checkpoint and os.path.exists(ns.checkpoint):
        run.load_checkpoint(ns.checkpoint)
        if rank == 0:
            print(f"hefl_b200: resumed from {ns.checkpoint} at round {run.round}", file=sys.stderr)
    while run.round < cfg.rounds:
        rec = run.run_round(check=True)
        rec["stage_ms_max"] = StageTimes.max_over_ranks(rec["stage_ms"], device, None)
        run.log.write(rec)
        if rank == 0:
            print(json.dumps({"round": rec["round"], "loss": rec["loss"], "accuracy": rec["accuracy"],
                              "stage_ms_max": rec["stage_ms_max"], "clients": world, "transport": rec["transport"]}))
        if ns.checkpoint:
            run.save_checkpoint(ns.checkpoint)       # rank 0: global model; every rank: its own optimiser / RNG state
        if world > 1:
            dist.barrier()
    if ns.test_dir and rank == 0:
        from .fl.data import ImageFolderDataset

        test = ImageFolderDataset(ns.test_dir, cfg.image_size, cfg.in_channels, shuffle_seed=None, pin=False)
        print(json.dumps({"test_images": len(test), **run.evaluate(test)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
