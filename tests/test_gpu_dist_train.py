"""-m gpu: data-parallel training through the experiment runner: two ranks (both on cuda:0, gloo
backend, since one box of the test pool has a single GPU and NCCL refuses two ranks per device) run
``MatrixFlowExperiment.train(max_steps=2)`` from DIFFERENT random initialisations.  The runner must create the process
group, broadcast rank 0's parameters, shard the data, exchange the gradients bucket by bucket during the backward
(``optim.GradSync``) and leave both ranks with bit-identical parameters that differ from a single-process run on one shard."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

OV = ["algorithm.target=flow", "algorithm.gpu_augment=true", "algorithm.lr=1e-3", "dataset.height=32", "dataset.width=32",
      "dataset.length=8", "experiment.training.data.batch_size=2", "experiment.training.data.shuffle=false"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0",
                      FD_DIST_BACKEND="gloo")
    import random
    import torch.distributed as dist
    from opticalflowdiffusion_b200.config import compose
    from opticalflowdiffusion_b200.experiments import build_experiment
    random.seed(0)
    torch.manual_seed(1000 + rank)                      # different init per rank: the runner has to broadcast rank 0's
    exp = build_experiment(compose(OV), None, None)
    exp.algo.preprocess = (lambda f: (lambda batch, aug=True: f(batch, aug=False)))(exp.algo.preprocess)
    w0 = exp.algo.unet.final_conv.weight.detach().clone()
    torch.manual_seed(7)                                # same t / noise draws on both ranks
    out = exp.train(max_steps=2)
    sync = exp.algo.unet.grad_sync
    import hashlib
    # plain Python values only: tensors sent through a multiprocessing queue live in shared memory owned by the sender
    sd = {k: hashlib.sha1(v.detach().cpu().numpy().tobytes()).hexdigest() for k, v in exp.algo.unet.state_dict().items()}
    n_params = sum(v.numel() for v in exp.algo.unet.state_dict().values())
    q.put((rank, float(w0.double().sum()), sd, n_params, out["steps"], sync.buckets_last_backward, sync.bytes_last_backward,
           float(exp.algo.optimizers.grad_scale)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_training_keeps_replicas_identical():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=600) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    (_, init0, sd0, n_params, steps0, nb0, bytes0, gs0), (_, init1, sd1, _, steps1, nb1, bytes1, gs1) = out
    assert init0 != init1                                       # the ranks really started from different weights
    assert steps0 == steps1 == 2 and gs0 == gs1 == 0.5
    assert nb0 == nb1 == 8                                      # unet_train.GRAD_GROUPS: every bucket went through GradSync
    assert bytes0 >= 4 * n_params                               # the whole fp32 gradient was exchanged
    for k in sd0:
        assert sd0[k] == sd1[k], k                              # identical replicas after two optimiser steps (SHA-1 per tensor)
