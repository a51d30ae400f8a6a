"""
Row sharding across the GPUs of one box with the sum over ranks inside the C call
(``lslb_*_peer_*`` over NVLink peer memory, ``lslb_*_nccl_*``). Needs >= 2 GPUs: skipped on the
single-GPU box, exercised by ``gpurun --gpus 2``. The world-size-1 degenerate case runs everywhere.
"""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from liesel_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_entry_points_world1():
    """One rank: the peer / NCCL variants return exactly what the plain call returns."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from liesel_b200 import dist as ld
    from liesel_b200.plan import Plan

    spec, centre = synth.s3_gamlss(n=30_000)
    plan = Plan(spec)
    C = 40
    th, aux = synth.evaluation_points(spec, centre, C)
    t, ax = torch.as_tensor(th, device=plan.device), torch.as_tensor(aux, device=plan.device)
    lp0, g0 = plan.logp_grad(t, ax)
    s0, i0, c0 = plan.iwls("loc_np0", t, ax)
    peer = ld.PeerGroup(C * 20 * 21, plan.device)
    nccl = ld.NcclComm(plan.device)
    for sh in (peer, nccl):
        for _ in range(3):  # both slots
            lp, g = plan.logp_grad(t, ax, shard=sh)
            assert torch.equal(lp, lp0) and torch.equal(g, g0)
        lpv, none = plan.logp_grad(t, ax, want_grad=False, shard=sh)
        assert none is None
        np.testing.assert_allclose(lpv.cpu().numpy(), lp0.cpu().numpy(), rtol=1e-6)
        s, i, c = plan.iwls("loc_np0", t, ax, shard=sh)
        assert torch.equal(s, s0) and torch.equal(i, i0) and torch.equal(c, c0)
    assert not peer.timed_out()
    # the reference-facing interface takes the same transport objects
    from liesel_b200.interface import B200Interface

    iface = B200Interface(plan, shard=peer)
    state = iface.unpack(t, ax)
    assert torch.equal(iface.log_prob_and_grad(state)[0], lp0)
    assert torch.equal(iface.chol_info_fn("loc_np0_beta")(state), c0)
    # a result larger than the exchange buffer is refused with a status code, not a crash
    from liesel_b200 import _lib

    small = ld.PeerGroup(8, plan.device)
    with pytest.raises(_lib.LieselB200Error):
        plan.logp_grad(t, ax, shard=small)
    for h in (peer, nccl, small):
        h.close()


def test_peer_group_failure_is_loud_and_recoverable():
    """A failed peer group never returns a partial sum: after ``abort`` (what a timed-out wait or a
    failed launch does) the host refuses further calls with a status code, the device-side error
    word is set, and ``reset`` brings the group back to bitwise identical results."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from liesel_b200 import _lib
    from liesel_b200 import dist as ld
    from liesel_b200.plan import Plan

    spec, centre = synth.s3_gamlss(n=20_000)
    plan = Plan(spec)
    C = 40
    th, aux = synth.evaluation_points(spec, centre, C)
    t, ax = torch.as_tensor(th, device=plan.device), torch.as_tensor(aux, device=plan.device)
    lp0, g0 = plan.logp_grad(t, ax)
    peer = ld.PeerGroup(C * 43, plan.device)
    lp, g = plan.logp_grad(t, ax, shard=peer)
    assert torch.equal(lp, lp0) and not peer.timed_out()
    peer.abort()
    torch.cuda.synchronize()
    assert peer.timed_out()
    with pytest.raises(_lib.LieselB200Error):
        plan.logp_grad(t, ax, shard=peer)
    peer.reset()
    assert not peer.timed_out()
    for _ in range(3):
        lp, g = plan.logp_grad(t, ax, shard=peer)
        assert torch.equal(lp, lp0) and torch.equal(g, g0)
    peer.close()


def test_row_sharded_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on one box")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "rowshard_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
    rep = json.loads(lines[-1])
    assert rep["ok"], rep
