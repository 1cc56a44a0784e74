"""Multi-process worker: one full encrypted-FedAvg round through FederatedRunner on every rank;
checks that the decrypted aggregate equals the plaintext mean of the ranks' locally trained weights
and that all ranks end with identical models. gloo (CPU) or nccl (one GPU per rank)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hefl_b200.config import FLConfig  # noqa: E402
from hefl_b200.fl import FederatedRunner  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--model", default="cnn2")
    ap.add_argument("--transport", default=None)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--key-holder", type=int, default=-1)
    ap.add_argument("--drop-rank", type=int, default=-1)
    ap.add_argument("--masks", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    gpu = args.backend == "nccl"
    if gpu:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        device = torch.device("cuda", torch.cuda.current_device())
        dist.init_process_group("nccl", device_id=device)
    else:
        device = torch.device("cpu")
        dist.init_process_group("gloo")
    if args.model == "medcnn":
        cfg = FLConfig(model="medcnn", local_epochs=1, steps_per_epoch=2, val_steps=1, clients=world,
                       nn_backend="tcgen05" if gpu else "cudnn", transport=args.transport or "fused",
                       device="cuda" if gpu else "cpu", debug_poison=True)
    else:
        cfg = FLConfig(model="cnn2", image_size=28, in_channels=1, num_classes=10, batch_size=8, local_epochs=1,
                       steps_per_epoch=2, val_steps=1, clients=world, he_preset="n4096_l3", nn_backend="cudnn",
                       dtype="bf16" if gpu else "fp32", transport=args.transport or ("fused" if gpu else "gloo"),
                       device="cuda" if gpu else "cpu", debug_poison=True)   # stale words would break the cross-check
    cfg.key_holder = args.key_holder
    cfg.allow_dropouts = args.drop_rank >= 0
    cfg.pairwise_masks = args.masks
    run = FederatedRunner(cfg, rank=rank, world=world, device=device)
    if args.key_holder >= 0:
        assert (run.sk is not None) == (rank == args.key_holder), "only the key holder may keep the secret key"
    # every rank must hold the SAME public key (generated once, from OS entropy, by the key holder / rank 0)
    pk0 = run.pk.clone()
    dist.broadcast(pk0, src=0)
    assert torch.equal(pk0, run.pk), "public keys differ between ranks"
    worst = 0.0
    for rnd in range(args.rounds):
        run.local_train()
        mine = run.pack.flat.clone()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        if args.drop_rank >= 0:                       # that client sits every round out: mean over the others
            run.participating = rank != args.drop_rank
            gathered = [g for r, g in enumerate(gathered) if r != args.drop_rank]
        plain_mean = torch.stack(gathered).mean(0)
        ct = run.encrypt_update()
        if hasattr(run.transport, "peer_load_steps"):
            run.transport.peer_load_steps()
        agg = run.aggregate(ct)
        if hasattr(run.transport, "peer_load_steps") and args.key_holder >= 0 and world > 1:
            steps = run.transport.peer_load_steps()
            # the key holder owns no chunk: it must not have loaded a single word of a peer's un-aggregated data
            assert (steps == 0) == (rank == args.key_holder), f"rank {rank}: {steps} peer-load steps"
        run.decrypt_apply(agg)
        run.guard_finite()
        if hasattr(run.transport, "check_status"):
            run.transport.check_status()
        err = float((run.pack.flat - plain_mean).abs().max())
        worst = max(worst, err)
        # every rank must hold the same global model after the round
        ref = run.pack.flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, run.pack.flat), "ranks diverged"
        run.round += 1
    t = torch.tensor([worst], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("MP_REPORT " + json.dumps({"world": world, "backend": args.backend, "transport": run.transport.name,
                                         "rounds": args.rounds, "max_abs_err": float(t), "n_ct": run.n_ct}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
