"""
Host-side logic of the N > 1 paths with world_size = 2 on CPU (`gloo`). The per-rank evaluation is
played by the oracle (the CUDA library cannot run here); what is under test is liesel_b200.dist:
row/chain partition arithmetic, prior-once semantics and the all-reduce / all-gather plumbing.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from liesel_b200 import dist as ld
from liesel_b200 import synth
from liesel_b200.spec import ModelSpec
from oracle import sam


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_arithmetic():
    for n, w in [(10, 3), (7, 8), (1024, 8), (5, 1)]:
        r = [ld.chain_shard(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    g = np.repeat(np.arange(7), [3, 5, 1, 4, 6, 2, 9])
    r = [ld.row_shard(len(g), k, 4, g) for k in range(4)]
    assert r[0][0] == 0 and r[-1][1] == len(g)
    for (a, b) in r[:-1]:
        assert b == len(g) or g[b] != g[b - 1]  # boundaries sit on group starts
    with pytest.raises(ValueError):
        ld.chain_shard(4, 4, 4)


def _slice_spec(spec, a, b):
    """Rows [a, b) of an S4-style (dense, logistic) spec."""
    out = ModelSpec(dtype=spec.dtype)
    out.add_response(spec.y[a:b], "bernoulli").add_predictor("logits")
    blk = spec.blocks[0]
    out.add_p_smooth(blk.X[a:b], blk.prior.m, blk.prior.s, "logits", blk.name)
    return out


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec, centre = synth.s4_logistic(n=600, p=5, dtype=np.float64)
        th, aux = synth.evaluation_points(spec, centre, 6)
        a, b = ld.row_shard(spec.n, rank, world)
        local = _slice_spec(spec, a, b)

        def evaluate_local(params, aux_, flags):
            p = params.numpy()
            lp = sam.log_prob(local, p)
            g = sam.grad(local, p)
            if flags & ld.SKIP_PRIOR:  # likelihood part only on ranks > 0
                parts = sam.log_prob_parts(local, p)
                lp = parts["response"]
                pr = local.blocks[0].prior
                g = g + (p - pr.m) / (pr.s * pr.s)
            return torch.from_numpy(np.ascontiguousarray(lp)), torch.from_numpy(np.ascontiguousarray(g))

        rs = ld.RowSharded(evaluate_local)
        lp, g = rs.logp_grad(torch.from_numpy(th.astype(np.float64)))
        lp_v, none = rs.logp_grad(torch.from_numpy(th.astype(np.float64)), want_grad=False)
        ref_lp = sam.log_prob(spec, th)
        ref_g = sam.grad(spec, th)
        ok = (np.allclose(lp.numpy(), ref_lp, rtol=1e-12) and np.allclose(g.numpy(), ref_g, rtol=1e-10, atol=1e-10)
              and none is None and np.allclose(lp_v.numpy(), ref_lp, rtol=1e-12))
        # chain sharding: every rank evaluates its own chains, gather restores the order
        c0, c1 = ld.chain_shard(th.shape[0], rank, world)
        mine = torch.from_numpy(sam.log_prob(spec, th[c0:c1]))
        allc = ld.gather_chains(mine, th.shape[0])
        ok = ok and np.allclose(allc.numpy(), ref_lp, rtol=1e-12)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_row_and_chain_sharding_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


@pytest.mark.parametrize("cfg", ["s3c", "s5", "s2"])
def test_spec_rows_partials_add_up(cfg):
    """``ModelSpec.rows`` keeps layout/priors/knots: likelihood parts of the shards add up (fp64 oracle)."""
    from oracle import splines as OS

    if cfg == "s3c":
        spec, centre = synth.s3_gamlss(n=3000, dtype=np.float64)
    elif cfg == "s5":
        spec, centre = synth.s5_poisson_ri(n=3000, G=90, dtype=np.float64, balanced=False)
    else:
        spec, centre = synth.s2_pspline_gam(n=2000, dtype=np.float64)
    for blk in spec.blocks:  # bands and column means are fixed on the full data before sharding
        if blk.kind == 1 and blk.start is None:
            blk.start, blk.weights = OS.banded_basis(blk.x, blk.knots, blk.order)
        if blk.kind == 1 and blk.center and blk.colmean is None:
            blk.colmean = OS.band_to_dense(blk.start, np.asarray(blk.weights, dtype=np.float64), blk.ncoef).mean(axis=0)
    th, aux = synth.evaluation_points(spec, centre, 3)
    ref_lik = sam.log_prob_parts(spec, th, aux)["response"]
    gidx = next((b.idx for b in spec.blocks if b.idx is not None), None)
    acc = 0.0
    for rank in range(3):
        a, b = ld.row_shard(spec.n, rank, 3, gidx)
        part = spec.rows(a, b)
        assert part.num_params == spec.num_params and part.position_keys() == spec.position_keys()
        acc = acc + sam.log_prob_parts(part, th, aux)["response"]
    np.testing.assert_allclose(acc, ref_lik, rtol=1e-11)
    with pytest.raises(ValueError):
        spec.rows(5, spec.n + 1)
