"""Row sharding with the sum over ranks INSIDE the C call: peer memory (lslb_*_peer_*) and NCCL
(lslb_*_nccl_*) against the torch.distributed all-reduce of the partials and against a single-rank
evaluation of all rows. Launch with torchrun on >= 2 GPUs of one box:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      tools/rowshard_check.py [--bench]

Checks (exit code 1 on failure, one JSON line on rank 0):
  * logp/grad of a spline location-scale model (fp32 and fp64) and of the dense logistic model that
    takes the tcgen05 path, IWLS score/information/Cholesky of a spline block;
  * peer results are BITWISE identical on all ranks and equal to the rank-ordered sum of partials;
  * 40 back-to-back sharded calls with changing parameters (slot alternation, no host sync);
  * --bench: per-call time of one S4 shard per rank (1.25e6 rows x 256, 1024 chains) per transport.
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
from liesel_b200 import splines, synth  # noqa: E402
from liesel_b200.plan import Plan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bench", action="store_true")
ap.add_argument("--rows-per-gpu", type=int, default=1_250_000)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

report: dict = {"n_gpus": world, "checks": {}}
fails: list[str] = []


def same_on_all_ranks(t: torch.Tensor) -> bool:
    if world == 1:
        return True
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return all(torch.equal(parts[0].view(torch.uint8), p.view(torch.uint8)) for p in parts[1:])


def torch_sum(t: torch.Tensor) -> torch.Tensor:
    """Rank-ordered sum of per-rank tensors (what the peer kernel must reproduce bit for bit)."""
    if world == 1:
        return t.clone()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    acc = parts[0].clone()
    for p in parts[1:]:
        acc += p
    return acc


def fix_bands(spec):
    """Bands are a function of (x, knots) per row; fix them on the full data before slicing rows."""
    for blk in spec.blocks:
        if blk.kind == 1 and blk.start is None:
            from liesel_b200 import _lib

            x = torch.as_tensor(blk.x, device=dev)
            kn = torch.as_tensor(blk.knots, device=dev)
            s, w = _lib.spline_basis(x, kn)
            blk.start, blk.weights = s.cpu().numpy(), w.cpu().numpy()
    return spec


def check_model(tag, spec, centre, C, iwls_block=None, rtol=2e-5):
    spec = fix_bands(spec)
    th, aux = synth.evaluation_points(spec, centre, C)
    t = torch.as_tensor(th, device=dev)
    ax = torch.as_tensor(aux, device=dev) if spec.num_aux else None
    gidx = next((b.idx for b in spec.blocks if b.idx is not None), None)
    a, b = ld.row_shard(spec.n, rank, world, gidx)
    full = Plan(spec, device=dev)
    part = Plan(spec.rows(a, b), device=dev)
    lp_ref, g_ref = full.logp_grad(t, ax)
    flags = 0 if rank == 0 else ld.SKIP_PRIOR
    lp_loc, g_loc = part.logp_grad(t, ax, flags=flags)
    lp_sum, g_sum = torch_sum(lp_loc), torch_sum(g_loc)
    p_elems = max(C * (spec.num_params + 1), C * 64 * 65 if iwls_block else 0)
    peer = ld.PeerGroup(p_elems, dev)
    nccl = ld.NcclComm(dev)
    res = {}
    for name, sh in (("peer", peer), ("nccl", nccl)):
        lp, g = part.logp_grad(t, ax, shard=sh)
        lp_v, _ = part.logp_grad(t, ax, want_grad=False, shard=sh)
        torch.cuda.synchronize()
        scale = float(g_ref.abs().max())
        ok = (torch.allclose(lp, lp_ref, rtol=rtol, atol=0) and
              torch.allclose(g, g_ref, rtol=rtol, atol=rtol * scale) and
              torch.allclose(lp_v, lp, rtol=1e-6, atol=0))
        if name == "peer":
            ok = ok and torch.equal(lp, lp_sum) and torch.equal(g, g_sum)
            ok = ok and same_on_all_ranks(lp) and same_on_all_ranks(g)
        res[name] = bool(ok)
        if iwls_block is not None:
            s_ref, i_ref, c_ref = full.iwls(iwls_block, t, ax)
            s, i, c = part.iwls(iwls_block, t, ax, shard=sh)
            torch.cuda.synchronize()
            ok_i = (torch.allclose(s, s_ref, rtol=rtol, atol=rtol * float(s_ref.abs().max())) and
                    torch.allclose(i, i_ref, rtol=rtol, atol=rtol * float(i_ref.abs().max())) and
                    torch.allclose(c, c_ref, rtol=50 * rtol, atol=50 * rtol * float(c_ref.abs().max())))
            if name == "peer":
                ok_i = ok_i and same_on_all_ranks(i) and same_on_all_ranks(c)
            res[name + "_iwls"] = bool(ok_i)
    # back-to-back calls with changing parameters, no host synchronisation in between
    outs = []
    for k in range(40):
        tk = (t * (1.0 + 0.001 * k)).contiguous()
        outs.append((tk, part.logp_grad(tk, ax, shard=peer)))
    torch.cuda.synchronize()
    ok_seq = True
    for tk, (lp, g) in outs[::13]:
        l2, g2 = part.logp_grad(tk, ax, flags=flags)
        ok_seq = ok_seq and torch.equal(lp, torch_sum(l2)) and torch.equal(g, torch_sum(g2))
    res["peer_sequence"] = bool(ok_seq)
    res["peer_timed_out"] = peer.timed_out()
    if world > 1 and iwls_block is not None and spec.dtype == np.float32:
        # failure is fatal for the whole group: the LAST rank aborts instead of evaluating; every
        # other rank's call must come back all-NaN with its error word set (never a partial sum),
        # and a collective reset restores bitwise identical results
        if rank == world - 1:
            peer.abort()
            torch.cuda.synchronize()
            saw = peer.timed_out()
        else:
            lp, g = part.logp_grad(t, ax, shard=peer)
            torch.cuda.synchronize()
            saw = bool(torch.isnan(lp).all() and torch.isnan(g).all()) and peer.timed_out()
        peer.reset()
        lp, g = part.logp_grad(t, ax, shard=peer)
        torch.cuda.synchronize()
        res["peer_poison_then_reset"] = bool(saw and not peer.timed_out() and torch.equal(lp, lp_sum)
                                             and torch.equal(g, g_sum))
    res["variant"] = part.variant
    peer.close()
    nccl.close()
    flat = torch.tensor([int(all(v for k_, v in res.items() if k_ not in ("variant", "peer_timed_out"))
                             and not res["peer_timed_out"])], device=dev)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.MIN)
    if not int(flat.item()):
        fails.append(tag)
    report["checks"][tag] = res


K = splines.pspline_penalty(20, 2)
for dtype in (np.float32, np.float64):
    spec, centre = synth.s3_gamlss(n=60_000, dtype=dtype)
    check_model(f"s3_{np.dtype(dtype).name}", spec, centre, 96, iwls_block="loc_np0",
                rtol=2e-5 if dtype == np.float32 else 1e-10)
spec, centre = synth.s4_logistic(n=70_000, p=256, dtype=np.float32)
check_model("s4_tc", spec, centre, 256)
spec, centre = synth.s5_poisson_ri(n=40_000, G=900, dtype=np.float32, balanced=False)
check_model("s5_groups", spec, centre, 64)

if args.bench:
    n_total = args.rows_per_gpu * world
    spec, centre = synth.s4_logistic(n=n_total, p=256, rank=rank, world_size=world)
    plan = Plan(spec, device=dev)
    C = 1024
    th, _ = synth.evaluation_points(spec, centre, C)
    t = torch.as_tensor(th, device=dev)
    peer = ld.PeerGroup(C * 257, dev)
    nccl = ld.NcclComm(dev)
    rs = ld.RowSharded(lambda prm, ax, flags: plan.logp_grad(prm, ax, flags=flags))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    arms = {
        "local_only": lambda: plan.logp_grad(t, None, flags=0 if rank == 0 else ld.SKIP_PRIOR),
        "torch_allreduce": lambda: rs.logp_grad(t),
        "nccl_in_c": lambda: plan.logp_grad(t, None, shard=nccl),
        "peer_memory": lambda: plan.logp_grad(t, None, shard=peer),
    }
    timing = {}
    for name, fn in arms.items():
        for _ in range(3):
            flush.zero_()
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
        for a_, b_ in ev:
            flush.zero_()
            a_.record()
            fn()
            b_.record()
        torch.cuda.synchronize()
        ms = torch.tensor([sum(a_.elapsed_time(b_) for a_, b_ in ev) / args.steps], device=dev,
                          dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        timing[name] = float(ms.item())
    report["bench"] = {"config": "S4 row-sharded", "rows_per_gpu": args.rows_per_gpu, "p": 256,
                       "chains": C, "ms_per_call_max_over_ranks": timing,
                       "exchange_bytes": C * 257 * 4, "variant": plan.variant,
                       "peer_timed_out": peer.timed_out()}
    peer.close()
    nccl.close()

if rank == 0:
    report["ok"] = not fails
    report["failed"] = fails
    print(json.dumps(report), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(1 if fails else 0)
