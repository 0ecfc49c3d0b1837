"""World-size-2 gloo run (CPU) of the multi-GPU host logic: disjoint batch shards, DistributedSampler wiring of
the experiment loader, sync_dist metric reduction and bench.py's max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opticalflowdiffusion_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = (torch.arange(10).float().view(10, 1), torch.arange(10).view(10, 1) * 2)
        mine = parallel.shard_batch(batch, rank, world)
        red = parallel.reduce_metrics({"val/loss": float(rank + 1), "val/mse": 2.0 * rank})
        tmax = parallel.max_over_ranks(0.5 + rank)
        from opticalflowdiffusion_b200.config import compose
        from opticalflowdiffusion_b200.experiments.exp_matrix_flow import MatrixFlowExperiment
        cfg = compose(["algorithm.target=flow", "dataset.height=16", "dataset.width=24", "dataset.length=8",
                       "experiment.validation.data.batch_size=2"])
        exp = MatrixFlowExperiment.__new__(MatrixFlowExperiment)       # loader wiring only: no model, no GPU
        exp.cfg, exp.rank, exp.world = cfg, rank, world
        seen = []
        for img, tgt, flow in exp._loader("validation", cfg.experiment.validation):
            assert img.shape == (2, 3, 16, 24) and flow.shape == (2, 2, 16, 24)
            seen.append(float(img.sum()))
        # training's exchange step (optim.allreduce_gradients): SUM all-reduce of the flat gradient, mean folded into
        # the optimiser's grad_scale, and the next step() must consume exactly the reduced buffer
        from opticalflowdiffusion_b200 import optim

        class FlatStandIn:            # the slice of FusedAdam that allreduce_gradients touches (no CUDA here)
            def __init__(self, g):
                self.g, self.grad_scale, self._reduced = g, 1.0, None

            def _still_flat(self):
                return True

            def flat_gradient(self):
                return self.g

        fs = FlatStandIn(torch.full((5,), float(rank + 1)))
        optim.allreduce_gradients(fs)
        mean = optim.allreduce_flat(torch.tensor([float(rank), 10.0]))
        # GradSync: the bucketed exchange the UNet backward drives (bucket_ready per contiguous range, then finish);
        # fp32 and bf16 wire formats; afterwards allreduce_gradients must NOT exchange the same gradient again
        synced = []
        for dt in (None, torch.bfloat16):
            fs2 = FlatStandIn(torch.arange(12, dtype=torch.float32) * (rank + 1))
            gs = optim.GradSync(fs2, comm_dtype=dt)
            for lo, hi in ((8, 12), (4, 8), (0, 4)):          # reverse layer order
                gs.bucket_ready(fs2.g, lo, hi)
            gs.finish(fs2.g)
            pre = fs2._presynced
            optim.allreduce_gradients(fs2)
            synced.append((fs2.g.tolist(), fs2.grad_scale, pre, fs2._presynced, fs2._reduced is None, gs.buckets_last_backward,
                           gs.bytes_last_backward))
        # start-up: identical replicas (DDP's parameter broadcast)
        torch.manual_seed(100 + rank)
        lin = torch.nn.Linear(3, 2)
        lin.register_buffer("buf", torch.full((2,), float(rank)))
        nsent = parallel.sync_module_from_rank0(lin)
        q.put((rank, mine[0].flatten().tolist(), red, tmax, seen, (fs._reduced.tolist(), fs.grad_scale, mean.tolist()),
               synced, (nsent, lin.weight.detach().flatten().tolist(), lin.buf.tolist())))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, red0, t0, seen0, g0, sy0, b0), (r1, s1, red1, t1, seen1, g1, sy1, b1) = out
    want = [float(3 * i) for i in range(12)]                                        # rank 0: i, rank 1: 2i
    assert sy0 == sy1 and sy0[0] == (want, 0.5, True, False, True, 3, 48) and sy0[1] == (want, 0.5, True, False, True, 3, 24)
    assert b0 == b1 and b0[0] == 3 and b0[2] == [0.0, 0.0]                           # rank 0's init and buffer everywhere
    assert g0 == g1 == ([3.0] * 5, 0.5, [0.5, 10.0])                                  # SUM all-reduce, 1/world in grad_scale
    assert s0 == [0.0, 1.0, 2.0, 3.0, 4.0] and s1 == [5.0, 6.0, 7.0, 8.0, 9.0]       # disjoint, complete
    assert red0 == red1 == {"val/loss": 1.5, "val/mse": 1.0}
    assert t0 == t1 == 1.5
    assert len(seen0) == len(seen1) == 2 and set(seen0).isdisjoint(seen1)             # 8 items / 2 ranks / batch 2


def test_shard_range_balanced():
    for n, world in ((8, 2), (10, 4), (3, 8)):
        parts = [list(parallel.shard_range(n, r, world)) for r in range(world)]
        assert sum(parts, []) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_allreduce_gradients_refuses_silent_independent_replicas(monkeypatch):
    """WORLD_SIZE > 1 without a process group used to return silently: N independent models."""
    from opticalflowdiffusion_b200 import optim
    monkeypatch.setenv("WORLD_SIZE", "2")
    assert not dist.is_initialized()
    with pytest.raises(RuntimeError, match="not initialised"):
        optim.allreduce_gradients(object())
    monkeypatch.setenv("WORLD_SIZE", "1")
    optim.allreduce_gradients(object())          # single process: nothing to exchange
