"""S4 (logistic regression, p=256, 1024 chains) with DATA ROWS sharded over the ranks and one NCCL
all-reduce of [C, P+1] per evaluation (BASELINE.json config 4). Launch with torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      tools/bench_s4_rowshard.py --rows-per-gpu 1250000 --steps 5

Every rank evaluates ALL chains on its rows (rank > 0 with LSLB_SKIP_PRIOR); value = chains /
(max over ranks of the per-call time). Also checks the reduced result against a single-rank
evaluation of the concatenated rows on a few chains (only when --check and the data fit one GPU).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from liesel_b200 import dist as ld  # noqa: E402
from liesel_b200 import synth  # noqa: E402
from liesel_b200.plan import Plan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-gpu", type=int, default=1_250_000)
ap.add_argument("--chains", type=int, default=1024)
ap.add_argument("--p", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--check", action="store_true")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

n_total = args.rows_per_gpu * world
spec, centre = synth.s4_logistic(n=n_total, p=args.p, rank=rank, world_size=world)
plan = Plan(spec, device=dev)
th, aux = synth.evaluation_points(spec, centre, args.chains)
t = torch.as_tensor(th, device=dev)
rs = ld.RowSharded(lambda prm, ax, flags: plan.logp_grad(prm, ax, flags=flags))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

for _ in range(3):
    flush.zero_()
    lp, g = rs.logp_grad(t)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
for a, b in ev:
    flush.zero_()
    a.record()
    lp, g = rs.logp_grad(t)
    b.record()
torch.cuda.synchronize()
ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / args.steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ok = None
if args.check:
    # all ranks hold the same reduced result; rank 0 re-evaluates every shard itself on 4 chains
    sel = [0, 1, args.chains // 2, args.chains - 1]
    if rank == 0:
        acc_lp, acc_g = 0, 0
        for r in range(world):
            s_r, _ = synth.s4_logistic(n=n_total, p=args.p, rank=r, world_size=world)
            p_r = Plan(s_r, device=dev)
            l_r, g_r = p_r.logp_grad(t[sel].contiguous(), None, flags=0 if r == 0 else ld.SKIP_PRIOR)
            acc_lp, acc_g = acc_lp + l_r, acc_g + g_r
            del p_r
        ok = bool(torch.allclose(acc_lp, lp[sel], rtol=2e-6) and
                  torch.allclose(acc_g, g[sel], rtol=1e-4, atol=1e-3 * float(acc_g.abs().max())))
if rank == 0:
    print(json.dumps({
        "config": "S4 row-sharded", "n_total": n_total, "rows_per_gpu": args.rows_per_gpu, "p": args.p,
        "chains": args.chains, "n_gpus": world, "ms_per_eval_call": float(ms.item()),
        "evals_per_s": args.chains / (float(ms.item()) * 1e-3),
        "allreduce_bytes": args.chains * (args.p + 1) * 4, "variant": plan.variant, "check_ok": ok,
    }), flush=True)
if world > 1:
    dist.destroy_process_group()
