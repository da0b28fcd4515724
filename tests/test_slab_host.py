"""Host-side logic of the slab decomposition (N > 1 path) on CPU: the strip partition, and a world_size-2 gloo run
that plays the rendezvous bench.py / a Java host performs (rank 0 creates the 128-byte communicator id, broadcasts it,
every rank builds its ifl_config).  The device side (NCCL + CUDA IPC) is exercised on GPUs by tools/slab_check.py."""
import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("h,world", [(16384, 1), (16384, 2), (16384, 8), (1024, 4), (1000, 3), (96, 3), (33, 2), (192, 4), (288, 8)])
def test_slab_partition_covers_every_row_once(ifl, h, world):
    parts = ifl.slab_partition(h, world)
    assert len(parts) == world
    rows = []
    for (s0, s1, r0, r1) in parts:
        assert 0 <= s0 < s1 and r0 == min(32 * s0, h)          # no rank is ever empty (192 rows / 4 ranks, 288 / 8)
        rows += list(range(r0, r1))
    assert rows == list(range(h))
    # balanced: the strip counts differ by at most one, the larger ranges first
    counts = [p[1] - p[0] for p in parts]
    assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)


def test_slab_partition_rejects_more_ranks_than_strips(ifl):
    with pytest.raises(ValueError):
        ifl.slab_partition(32, 2)


WORKER = r'''
import os, sys, ctypes as C
sys.path.insert(0, os.environ["IFL_ROOT"])
import torch.distributed as dist
import incfluid_b200 as ifl
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ids = [bytes(range(128)) if rank == 0 else None]          # stands in for ifl_comm_unique_id() (needs NCCL + a GPU)
dist.broadcast_object_list(ids, src=0)
cfg = ifl.abi.default_config(ifl.abi.F3_CONJUGATE_GRADIENTS, 256, 256, rank=rank, world_size=world)
C.memmove(cfg.comm_id, ids[0], 128)
assert bytes(cfg.comm_id) == bytes(range(128)) and cfg.rank == rank and cfg.world_size == world
parts = ifl.slab_partition(256, world)
mine = [None] * world
dist.all_gather_object(mine, parts[rank])
assert mine == parts
# without a GPU the library must refuse loudly instead of falling back
lib, fn = ifl.load()
h = C.c_void_p()
rc = fn["ifl_create"](C.byref(cfg), C.byref(h))
import torch
if not torch.cuda.is_available():
    assert rc == ifl.abi.ERR_CUDA, rc
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank%d-ok\n" % rank); sys.stdout.flush()
'''


def test_two_rank_rendezvous_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, IFL_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                         env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "rank0-ok" in out.stdout and "rank1-ok" in out.stdout


@pytest.mark.parametrize("variant,accepted", [("F1_MATRIXLESS", False), ("F2_BETTER_ADVECTION", False), ("F3_CONJUGATE_GRADIENTS", True),
                                              ("F4_SOLID_BOUNDARIES", True), ("F5_CURVED_BOUNDARIES", True), ("F6_HEAT", True),
                                              ("F7_VARIABLE_DENSITY", True), ("F8_FLIP", False)])
def test_which_variants_run_on_several_gpus(ifl, variant, accepted):
    """ifl_create validates rank / world_size before it touches a device: the MIC(0)-PCG variants 3 (slab decomposition)
    and 4-7 (slab-decomposed solver loops) are accepted -- and then fail loudly here for want of a GPU -- while the
    Gauss-Seidel variants and FLIP are refused as unsupported."""
    import torch
    if accepted and torch.cuda.is_available():
        pytest.skip("with a GPU present ifl_create would go on to an NCCL rendezvous; tools/slab_check.py covers that")
    abi = ifl.abi
    cfg = abi.default_config(getattr(abi, variant), 256, 256, rank=1, world_size=2)
    lib, fn = ifl.load()
    h = C.c_void_p()
    rc = fn["ifl_create"](C.byref(cfg), C.byref(h))
    assert rc == (abi.ERR_CUDA if accepted else abi.ERR_UNSUPPORTED), rc
