"""Host-side multi-GPU logic on CPU: contiguous sharding + accept-bitmap gather over torch.distributed (gloo,
world_size 2).  The per-rank "verifier" here is the oracle (test infrastructure), standing in for the GPU shard."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def test_shard_ranges_cover_and_align():
    import stark_symphony_b200 as S
    from importlib import import_module

    sh = import_module("stark_symphony_b200.sharding")
    for n in (0, 1, 31, 32, 33, 1000, 1024, 65536, (1 << 20) + 5):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                b, e = sh.shard_range(n, r, world)
                assert b == covered or b == e == n
                assert b % 32 == 0 or b == n
                covered = max(covered, e)
            assert covered == n


def _worker(rank, world, port, n, tmp):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module

    import stark_symphony_b200  # noqa: F401
    sh = import_module("stark_symphony_b200.sharding")
    from oracle import oracle as O
    from oracle import witparse as W

    orc = O.Oracle()
    cfg = O.make_config("testing", O.MODE_PROVER_CONSISTENT)
    packed, _ = W.pack_stwo(W.load_wit(open(os.path.join(GOLDEN, "stwo_proof_testing.wit")).read()), 1, 2, 4)
    stride = len(packed)
    batch = np.tile(packed, n).reshape(n, stride)
    bad = list(range(3, n, 11))
    for r in bad:
        batch[r, 24] ^= 1  # corrupt an OODS sample
    b, e = sh.shard_range(n, rank, world)
    accept, _, _ = orc.stwo_verify_batch(cfg, batch[b:e].reshape(-1), e - b)
    local = torch.from_numpy(accept.view(np.int32).copy())
    full = sh.gather_accept_bitmaps(local, n, world)
    bits = np.unpackbits(full.numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
    expect = np.ones(n, dtype=bool)
    expect[bad] = False
    ok = bool((bits == expect).all())
    open(os.path.join(tmp, f"rank{rank}.ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [70, 64])
def test_two_rank_gather_matches_single_pass(tmp_path, n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "rank0.ok").read() == "1" and open(tmp_path / "rank1.ok").read() == "1"
