"""The N>1 path on CPU: world_size-2 gloo processes exercise the channel sharding and the long-stream segment
plan (host logic only; no collective is on the data path -- gloo carries just the barrier / max-over-ranks /
gather that bench.py and a multi-GPU driver use)."""
import os
import socket
from fractions import Fraction

import numpy as np
import pytest

import multirate_b200 as mr
import multirate_oracle as mo


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_channels, n_samples, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import multirate_b200 as mr_
    h = np.random.default_rng(1).random(3528).astype(np.float32)
    ratio = Fraction(147, 160)
    # (1) channel shard: every rank owns a block; the blocks tile the batch exactly
    lo, hi = mr_.channel_shard(n_channels, world, rank)
    blocks = [None] * world
    dist.all_gather_object(blocks, (lo, hi))
    assert blocks[0][0] == 0 and blocks[-1][1] == n_channels
    assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    # (2) every rank runs the SAME state machine on its shard: counts and carried state agree without talking
    f = mr_.FIRFilter(h, ratio, nchannels=hi - lo, sample_dtype=np.complex64, device=-1)
    import ctypes as C
    counts = []
    for n in (65536, 1, 65536, 777):
        c = C.c_int64(); mr_._ffi.check(mr_._ffi.lib().mrb_advance(f._handle, n, C.byref(c))); counts.append(c.value)
    s = f._get_state()
    t = torch.tensor(counts + [s.phi_idx, s.input_deficit], dtype=torch.int64)
    ts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(ts, t)
    assert all(torch.equal(ts[0], x) for x in ts)
    # (3) long-stream split: this rank's segment starts at the closed-form state; the segments tile the outputs
    plan = mr_.segment_plan(mr_.FIRFilter(h, ratio), n_samples, world, align=160)
    n0, n1, k0, cnt = plan[rank]
    g = [None] * world
    dist.all_gather_object(g, (k0, cnt))
    assert g[0][0] == 0 and all(g[i][0] + g[i][1] == g[i + 1][0] for i in range(world - 1))
    # (4) timing reduction used by bench.py: max over ranks
    tt = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    assert tt.item() == float(world)
    dist.barrier()
    if rank == 0:
        ret["total"] = g[-1][0] + g[-1][1]
    dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as tmp
    world, n_channels, n_samples = 2, 65536 + 3, 2 ** 31
    mgr = tmp.Manager()
    ret = mgr.dict()
    tmp.spawn(_worker, args=(world, _free_port(), n_channels, n_samples, ret), nprocs=world, join=True)
    # the segments' outputs add up to the single-stream count (reference outputlength, src/Filters.jl:352-357)
    assert ret["total"] == mo.outputlength_ratio(n_samples, Fraction(147, 160), 1)


def test_segment_plan_against_oracle_stream():
    """Small stream, 3 segments: outputs of segment r are exactly outputs [k0_r, k0_r + count_r) of the stream."""
    rng = np.random.default_rng(3)
    h = rng.random(61)
    for ratio in (Fraction(3, 17), Fraction(147, 160), Fraction(1, 8), Fraction(4, 1), Fraction(1, 1)):
        n = 5000
        x = rng.random(n)
        whole = mo.filt(h, x, ratio)
        plan = mr.segment_plan(mr.FIRFilter(h, ratio), n, 3, align=ratio.denominator)
        assert plan[0][2] == 0 and plan[-1][2] + plan[-1][3] == len(whole)
        H = mo.FIRFilter(h, ratio).historyLen
        for n0, n1, k0, cnt in plan:
            # oracle twin of mrb_seek: run a fresh filter over the halo-extended segment and drop the warm-up
            o = mo.FIRFilter(h, ratio)
            o.filt(x[:n0])                       # brings state AND history to n0 (what seek + halo provide)
            y = o.filt(x[n0:n1])
            assert len(y) == cnt and np.allclose(y, whole[k0:k0 + cnt], rtol=0, atol=1e-12 * np.abs(whole).max())


@pytest.mark.parametrize("polyorder", [None, 3])
def test_segment_plan_table_kinds_against_oracle_stream(polyorder):
    """SURVEY 8f rank 4 on the host: the plan of an arbitrary-rate / Farrow stream split (start states by exact replay)
    tiles the single-stream output exactly, and a fresh oracle filter brought to n0 reproduces each segment."""
    rng = np.random.default_rng(5)
    N = 16
    h = rng.random(N * 6)
    for rate in (0.918734, 2.31, 1 / 3.7):
        args = (h, rate, N) if polyorder is None else (h, rate, N, polyorder)
        n = 6000
        x = rng.random(n)
        whole = mo.FIRFilter(*args).filt(x)
        plan = mr.segment_plan(mr.FIRFilter(*args), n, 4)
        assert plan[0][2] == 0 and plan[-1][2] + plan[-1][3] == len(whole)
        assert all(a[2] + a[3] == b[2] for a, b in zip(plan, plan[1:]))
        for n0, n1, k0, cnt in plan:
            o = mo.FIRFilter(*args)
            o.filt(x[:n0])
            y = o.filt(x[n0:n1])
            assert len(y) == cnt and np.array_equal(y, whole[k0:k0 + cnt])


@pytest.mark.parametrize("n,w", [(10, 3), (65536, 8), (7, 8), (0, 2)])
def test_channel_shard_tiles(n, w):
    blocks = [mr.channel_shard(n, w, r) for r in range(w)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
