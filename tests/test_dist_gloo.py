"""World-size-2 gloo checks (CPU) of the data-parallel host logic: sample sharding and the single
flat gradient all-reduce a training caller issues per optimizer step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adaptiveisp_b200.dist import allreduce_grads, shard_range


def test_shard_range_partitions_exactly():
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))   # same init on all ranks
    unused = torch.nn.Linear(3, 2)                                            # grad stays None (like fc_mask)
    zero = torch.nn.Linear(4, 4)                                              # exact-zero grad (unselected filter)
    x_all = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10.0
    lo, hi = shard_range(8, rank, world)
    net(x_all[lo:hi]).sum().backward()
    for p in zero.parameters():
        p.grad = torch.zeros_like(p)
    params = list(net.parameters()) + list(unused.parameters()) + list(zero.parameters())
    nbytes = allreduce_grads(params, average=False)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    ref.load_state_dict(net.state_dict())
    ref(x_all).sum().backward()
    ok = all(torch.allclose(a.grad, b.grad, atol=1e-5) for a, b in zip(net.parameters(), ref.parameters()))
    ok = ok and all(p.grad is None for p in unused.parameters())
    ok = ok and all(float(p.grad.abs().max()) == 0.0 for p in zero.parameters())
    ok = ok and nbytes == 4 * sum(p.numel() for p in list(net.parameters()) + list(zero.parameters()))
    # the grads now live in ONE persistent flat buffer: a second backward accumulates into it in place and
    # the next all-reduce needs no gather / scatter copies
    bucket = params[0]._aisp_grad_bucket
    ok = ok and bucket.attached() and bucket.nbytes == nbytes
    ptrs = [p.grad.data_ptr() for p in bucket.params]
    bucket.zero()
    net(x_all[lo:hi]).sum().backward()
    ok = ok and [p.grad.data_ptr() for p in bucket.params] == ptrs
    ok = ok and allreduce_grads(params, average=False) == nbytes and params[0]._aisp_grad_bucket is bucket
    ok = ok and all(torch.allclose(a.grad, b.grad, atol=1e-5) for a, b in zip(net.parameters(), ref.parameters()))
    total = bucket.clip_grad_norm_(1e-5)
    ok = ok and float(total) > 0 and float(torch.linalg.vector_norm(bucket.flat)) <= 1.001e-5
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_numa_helpers_are_safe_without_a_gpu():
    """The pinned-buffer NUMA placement is best effort: without sysfs information (no GPU, VM) it
    reports node -1 and leaves the affinity mask alone."""
    import os
    from adaptiveisp_b200 import dist as D
    assert D._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert D._parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    node, prev = D.bind_to_gpu_numa_node(0)
    assert prev == before
    if node < 0:
        assert os.sched_getaffinity(0) == before
    else:
        os.sched_setaffinity(0, prev)
