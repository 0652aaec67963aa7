"""Host logic of the multi-GPU drivers at world size 2 under ``gloo`` on CPU (no GPU needed).

The drivers contain no arithmetic, so they are exercised with an oracle-backed backend
(``tests/seq_oracle_backend.py``) and must reproduce the oracle's single-process, unchunked forward."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avse_challenge_b200 import CONFIGS, init_state_dicts, synth_mixture
from avse_challenge_b200.parallel import (SequenceParallelSeparator, ShardedSeparator, make_seq_plan, shard_slices)
from oracle import restate
from tests.helpers import rel_max
from tests.seq_oracle_backend import OracleSeqBackend


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, fn, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, q, args)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):              # drain before join: results are plain numpy (no torch shared-memory handles)
        r, v = q.get()
        out[r] = _from_np(v)
    for p in procs:
        p.join(300)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    return out


def _to_np(v):
    if isinstance(v, torch.Tensor):
        return ("t", v.detach().cpu().numpy())
    if isinstance(v, (tuple, list)):
        return ("l", [_to_np(x) for x in v])
    return ("o", v)


def _from_np(v):
    kind, x = v
    if kind == "t":
        return torch.from_numpy(x)
    if kind == "l":
        return [_from_np(y) for y in x]
    return x


def _entry(rank, world, port, fn, q, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = ("o", None)
        try:
            res = _to_np(fn(rank, world, *args))
        except Exception as e:  # report instead of hanging the parent on q.get()
            import traceback
            res = ("o", "ERROR: " + traceback.format_exc())
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- plans
def test_shard_slices_cover_and_balance():
    for n in (0, 1, 5, 32, 257):
        for w in (1, 2, 3, 8):
            sl = shard_slices(n, w)
            assert sl[0][0] == 0 and sl[-1][1] == n and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("L,world,sub", [(101, 2, 3), (3999, 8, 64), (1199999, 8, 64), (24, 8, 4), (50, 1, 7),
                                         (1199999, 8, 74), (1199999, 1, 74), (3999, 1, 16)])
def test_seq_plan_is_a_partition(L, world, sub):
    plan = make_seq_plan(L, world, sub)
    assert plan.ranges[0][0] == 0 and plan.ranges[-1][1] == L
    for (a, b), c, last in zip(plan.ranges, plan.chunks, plan.last_len):
        assert b - a >= 3 and 1 <= c <= plan.cmax and 1 <= last <= plan.Ls
        assert (c - 1) * plan.Ls + last == b - a
    with pytest.raises(ValueError):
        make_seq_plan(5, 2, 1)


# ----------------------------------------------------------------------------------------------- batch sharding
def _sharded_job(rank, world, B, T):
    hp = CONFIGS["tiny"]
    sds = init_state_dicts(hp, 1234)
    mix, _ = synth_mixture(B, T, seed=3)
    calls = []

    def local(m):
        calls.append(m.shape[0])
        with torch.no_grad():
            return restate.separate(m, sds, hp.n_mamba, scan_impl="c")

    est = ShardedSeparator(local, n_spk=hp.n_spk)(mix)
    return est, calls


def test_batch_sharding_world2_matches_single_process():
    B, T = 3, 808
    hp = CONFIGS["tiny"]
    sds = init_state_dicts(hp, 1234)
    mix, _ = synth_mixture(B, T, seed=3)
    with torch.no_grad():
        ref = restate.separate(mix, sds, hp.n_mamba, scan_impl="c")
    out = _run(2, _sharded_job, B, T)
    assert out[0][1] == [2] and out[1][1] == [1]          # uneven contiguous shards, one local call each
    for r in (0, 1):
        # utterances are independent; CPU matmuls block differently for 1-, 2- and 3-row batches, hence a tolerance
        assert out[r][0].shape == ref.shape and rel_max(out[r][0], ref) < 1e-5


# ----------------------------------------------------------------------------------------------- sequence parallel
def _seqpar_job(rank, world, T, sub, exchange):
    hp = CONFIGS["tiny"]
    sds = init_state_dicts(hp, 1234)
    mix, _ = synth_mixture(1, T, seed=11)
    sp = SequenceParallelSeparator(hp, sub_chunks=sub, exchange=exchange, backend=OracleSeqBackend(hp, sds))
    return sp(mix)


@pytest.mark.parametrize("world,T,sub,exchange", [(2, 816, 3, "allgather"), (2, 816, 3, "sendrecv"),
                                                  (2, 1003, 1, "allgather"), (3, 1500, 2, "sendrecv")])
def test_sequence_parallel_matches_unchunked_oracle(world, T, sub, exchange):
    hp = CONFIGS["tiny"]
    sds = init_state_dicts(hp, 1234)
    mix, _ = synth_mixture(1, T, seed=11)
    with torch.no_grad():
        ref = restate.separate(mix, sds, hp.n_mamba, scan_impl="c")
    out = _run(world, _seqpar_job, T, sub, exchange)
    for r in range(world):
        assert out[r].shape == ref.shape
        assert rel_max(out[r], ref) < 2e-5, (r, rel_max(out[r], ref))   # fp32 re-association across chunk seams only
    assert torch.equal(out[0], out[world - 1])


def test_sequence_parallel_single_rank_many_chunks():
    hp = CONFIGS["tiny"]
    sds = init_state_dicts(hp, 1234)
    mix, _ = synth_mixture(1, 1600, seed=5)
    with torch.no_grad():
        ref = restate.separate(mix, sds, hp.n_mamba, scan_impl="c")
    est = SequenceParallelSeparator(hp, sub_chunks=7, backend=OracleSeqBackend(hp, sds))(mix)
    assert rel_max(est, ref) < 2e-5


def test_sequence_parallel_defaults_and_rejections():
    """Host logic of the driver: default sub-chunk count (whole waves of a 148-SM GPU), the default plan on the
    oracle backend (no graph without the CUDA backend), causal stacks and bad arguments rejected at construction."""
    import inspect
    import avse_challenge_b200 as mtn
    assert mtn.SequenceParallelSeparator is SequenceParallelSeparator and mtn.ShardedSeparator is ShardedSeparator
    assert inspect.signature(SequenceParallelSeparator.__init__).parameters["sub_chunks"].default == 74
    hp = CONFIGS["tiny"]
    sds = init_state_dicts(hp, 1234)
    sp = SequenceParallelSeparator(hp, backend=OracleSeqBackend(hp, sds))
    assert sp.sub_chunks == 74 and sp.use_graph is False
    mix, _ = synth_mixture(1, 1600, seed=9)          # 199 frames -> 74 chunks of 3 frames (the conv halo's minimum)
    with torch.no_grad():
        ref = restate.separate(mix, sds, hp.n_mamba, scan_impl="c")
    assert rel_max(sp(mix), ref) < 2e-5
    with pytest.raises(NotImplementedError):
        SequenceParallelSeparator(hp.causal(), backend=OracleSeqBackend(hp, sds))
    with pytest.raises(ValueError):
        SequenceParallelSeparator(hp, exchange="ring", backend=OracleSeqBackend(hp, sds))
    with pytest.raises(ValueError):
        sp(torch.zeros(2, 1600))
