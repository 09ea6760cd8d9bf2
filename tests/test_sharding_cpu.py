"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch by image with no data-path
collective, run the per-rank work (here the CPU oracle on a tiny model, standing in for the engine), and
reduce timings with max-over-ranks exactly as bench.py does."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from m2trans_b200.sharding import gather_counts, max_over_ranks, shard_range


def test_shard_range_partitions_exactly():
    for total in (1, 7, 16, 32, 64, 65):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import m2trans_oracle as O
    from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict
    torch.set_num_threads(1)
    sd = synthetic_state_dict(2, 0, n_blocks=1)
    x = synthetic_input(3, 32, 32, seed=9)                       # the whole job: 3 frames
    s, e = shard_range(x.shape[0], rank, world)
    with torch.no_grad():
        y = O.forward(sd, x[s:e], n_blocks=1)                    # this rank's frames only; no exchange
    torch.save({"span": (s, e), "y": y}, os.path.join(out_dir, f"rank{rank}.pt"))
    ms = max_over_ranks([10.0 + rank, 5.0 - rank])               # per-rank timings -> max
    counts = gather_counts(e - s)
    dist.barrier()
    if rank == 0:
        torch.save({"ms": ms, "counts": counts}, os.path.join(out_dir, "reduce.pt"))
    dist.destroy_process_group()


def test_two_rank_image_sharding_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import m2trans_oracle as O
    from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict
    x = synthetic_input(3, 32, 32, seed=9)
    with torch.no_grad():
        want = O.forward(synthetic_state_dict(2, 0, n_blocks=1), x, n_blocks=1)
    parts = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    assert [p["span"] for p in parts] == [(0, 2), (2, 3)]
    got = torch.cat([p["y"] for p in parts], 0)
    assert torch.allclose(got, want, atol=1e-6)                  # sharding by image is exact
    red = torch.load(os.path.join(tmp_path, "reduce.pt"))
    assert red["ms"] == [11.0, 5.0] and red["counts"] == [2, 1]
