"""world_size-2 gloo test of the ray-shard + tile-gather host logic (CPU; the render function is the oracle)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hyperreel_b200.ray_shard import render_sharded, shard_range


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.hyperreel_oracle import HyperReelOracle
    from tests.cases import build_case
    torch.set_num_threads(1)
    case = build_case("shiny_tiny", n=n)
    orc = HyperReelOracle(case.model_cfg_plain, case.dataset, case.state_dict)
    out = render_sharded(case.rays, lambda r, **kw: {"rgb": orc.render(r.clone())})
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [301, 64])
def test_sharded_render_equals_single_process(n):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    from oracle.hyperreel_oracle import HyperReelOracle
    from tests.cases import build_case
    case = build_case("shiny_tiny", n=n)
    # per-shard renders must equal the corresponding rows of a one-shot render (rays are independent)
    ref = HyperReelOracle(case.model_cfg_plain, case.dataset, case.state_dict).render(case.rays.clone())
    assert out.shape == (n, 3)
    assert (out - ref).abs().max() <= 2e-6
