"""Multi-process worker: checks every ciphertext all-reduce transport bit-exactly against a
locally computed oracle, stresses flag reuse over many rounds, and (on GPUs) reports
bandwidth. Launched by torchrun from test_multiproc.py / bench/allreduce_sweep.py."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hefl_b200.he.context import CKKSContext  # noqa: E402
from hefl_b200.parallel import CollectiveTransport, FusedTransport  # noqa: E402


def make_input(ctx, rank, rnd, C):
    gen = torch.Generator().manual_seed(1000 * rnd + rank)
    return torch.stack([torch.randint(0, q, (C, 2, ctx.n), generator=gen, dtype=torch.int64)
                        for q in ctx.primes], dim=2).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--cts", type=int, default=9)
    ap.add_argument("--algos", default="two_shot,one_shot")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    gpu = args.backend == "nccl"
    if gpu:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group(args.backend)
    device = torch.device("cuda", torch.cuda.current_device()) if gpu else torch.device("cpu")
    ctx = CKKSContext(4096, prime_bits=(36, 36, 37), scale_bits=40, device=device)
    C = args.cts
    numel = C * 2 * ctx.L * ctx.n
    report = {"world": world, "backend": args.backend, "checks": 0}

    def expected(rnd):
        acc = sum(make_input(ctx, r, rnd, C) for r in range(world))
        for l, q in enumerate(ctx.primes):
            acc[:, :, l] %= q
        return acc

    coll = CollectiveTransport(ctx, numel)
    for rnd in range(args.rounds):
        x = coll.buffer(numel).view(C, 2, ctx.L, ctx.n)
        x.copy_(make_input(ctx, rank, rnd, C))
        got = coll.allreduce(x)
        assert torch.equal(got.cpu(), expected(rnd)), f"collective mismatch round {rnd}"
        report["checks"] += 1

    if gpu:
        ft = FusedTransport(ctx, numel, timeout_s=10.0)
        report["symm_backend"] = ft.sym.backend
        report["multicast"] = bool(ft.sym.mc_ptr)
        algos = args.algos.split(",")
        if ft.sym.mc_ptr and "multimem" not in algos and os.environ.get("HEFL_TEST_MULTIMEM", "1") == "1":
            algos.append("multimem")
        for algo in algos:
            ft.algo = algo
            for rnd in range(args.rounds):
                x = ft.buffer(numel).view(C, 2, ctx.L, ctx.n)
                x.copy_(make_input(ctx, rank, rnd, C))
                got = ft.allreduce(x)
                torch.cuda.synchronize()
                ft.check_status()
                assert torch.equal(got.cpu(), expected(rnd)), f"fused {algo} mismatch round {rnd} rank {rank}"
                report["checks"] += 1
                dist.barrier()
            # back-to-back launches without host sync: flag reuse stress
            x = ft.buffer(numel).view(C, 2, ctx.L, ctx.n)
            for rnd in range(20):
                x.copy_(make_input(ctx, rank, 0, C))
                got = ft.allreduce(x)
            torch.cuda.synchronize()
            ft.check_status()
            assert torch.equal(got.cpu(), expected(0)), f"fused {algo} stress mismatch"
            report["checks"] += 1
            dist.barrier()
        report["algos"] = algos
    dist.barrier()
    if rank == 0:
        print("MP_REPORT " + json.dumps(report))
        if args.out:
            json.dump(report, open(args.out, "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
