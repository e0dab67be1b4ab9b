"""world_size-2 gloo test (CPU) of the data-parallel host logic: sharding, the rendezvous payload broadcast, metric
averaging, and the identity the design rests on -- the SUM of per-rank gradients scaled by 1/N equals the gradient of
the global batch (the oracle stands in for the GPU on each rank)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from midi_vae_b200 import dist as mdist
    from oracle import midivae_oracle as O
    from tests import util
    ecfg, ocfg = util.make_cfgs(T=8, H=16, L=8, feedback="teacher_forced", max_batch=8)
    w = util.make_weights(ecfg)
    p = util.to_torch(w)
    r, hist, eps, _ = util.make_batch(ecfg, 8, seed=11)
    # full-batch reference (identical on both ranks)
    X, I, V, C, th, te, _ = util.oracle_inputs(ocfg, r, hist, eps, None)
    m_full, g_full, _ = O.loss_and_grads(ocfg, p, X, I, V, C, th, te)
    # this rank's shard
    rs = mdist.shard_rolls(r, rank, world)
    Xs, Is, Vs, Cs, ths, tes, _ = util.oracle_inputs(ocfg, rs, mdist.shard_array(hist, rank, world), mdist.shard_array(eps, rank, world), None)
    m_loc, g_loc, _ = O.loss_and_grads(ocfg, p, Xs, Is, Vs, Cs, ths, tes)
    worst = 0.0
    for k in sorted(g_loc):
        t = g_loc[k].clone()
        dist.all_reduce(t)                      # what ncclAllReduce(sum) does on the flat arena
        t /= world                              # grad_scale = 1/world inside Adam
        worst = max(worst, float((t - g_full[k]).abs().max() / (g_full[k].abs().max() + 1e-30)))
    m_avg = mdist.average_metrics(m_loc)
    uid = mdist.broadcast_bytes(bytes(range(128)) if rank == 0 else None)
    ok = worst < 1e-10 and abs(m_avg["loss"] - m_full["loss"]) < 1e-10 and uid == bytes(range(128)) and mdist.shard_bounds(8, rank, world) == (4 * rank, 4 * rank + 4)
    out[rank] = (ok, worst, m_avg["loss"], m_full["loss"])
    dist.destroy_process_group()


def test_data_parallel_identities_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        ok, worst, la, lf = out[rank]
        assert ok, (rank, worst, la, lf)


def test_shard_helpers():
    sys.path.insert(0, ROOT)
    from midi_vae_b200 import dist as mdist, synth
    r = synth.make_batch(6, 8)
    assert len(mdist.shard_rolls(r, 1, 3)) == 2
    assert np.array_equal(mdist.shard_rolls(r, 2, 3).pitch, r.pitch[4:6])
    with pytest.raises(ValueError):
        mdist.shard_bounds(7, 0, 2)
    assert mdist.shard_array(None, 0, 2) is None
