"""N > 1 host logic on CPU: two processes over gloo (127.0.0.1).  What the ranks of a multi-GPU run
must agree on without talking about it — the decomposition plan of a shared body — and what they
exchange — shards of an ensemble, 64-byte mailbox handles — exactly as bench.py does it."""
import ctypes as C
import hashlib
import os
import socket
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import ctypes as C, hashlib, importlib, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["SBS_ROOT"])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
sc = importlib.import_module("soft-body-simulator_b200.scenes")

# (1) ensemble sharding as in bench.make_scene(config4): contiguous chunks of bodies, no overlap
total = 8
per = total // world
shard = sc.config4(per, W=3, H=3, D=4, first=rank * per)
mine = [hashlib.sha1(it.x.tobytes()).hexdigest() for it in shard.items if isinstance(it, sc.TetBody)]
everyone = [None] * world
dist.all_gather_object(everyone, mine)
full = sc.config4(total, W=3, H=3, D=4)
want = [hashlib.sha1(it.x.tobytes()).hexdigest() for it in full.items if isinstance(it, sc.TetBody)]
assert sum(everyone, []) == want, "shards must tile the ensemble in rank order"

# (2) the decomposition plan of ONE body is computed independently by every rank: it must be identical
L = C.CDLL(os.environ["SBS_HOSTSCENE"])
u32p, dp, i32p = C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.POINTER(C.c_int32)
L.hs_partition.argtypes = [C.c_int64, C.c_int64, u32p, dp, C.c_int, C.c_int, i32p, i32p, C.c_int]
pos, tets = sc.bar_model(9, 9, 41)
tets = np.ascontiguousarray(tets, np.uint32)
x0 = np.ascontiguousarray(pos, np.float64)
treg = np.empty(len(tets), np.int32)
rrank = np.full(1024, -1, np.int32)
n = L.hs_partition(len(pos), len(tets), tets.ctypes.data_as(u32p), x0.ctypes.data_as(dp), 148, world,
                   treg.ctypes.data_as(i32p), rrank.ctypes.data_as(i32p), 1024)
assert n > 0 and n % world == 0, n
digest = hashlib.sha1(treg.tobytes() + rrank[:n].tobytes()).hexdigest()
digests = [None] * world
dist.all_gather_object(digests, digest)
assert len(set(digests)) == 1, "ranks disagree on the plan"
# every rank runs an equal, contiguous block of regions, and every tet has exactly one rank
assert np.array_equal(rrank[:n], np.repeat(np.arange(world), n // world))
tet_rank = rrank[treg]
counts = np.bincount(tet_rank, minlength=world)
assert counts.sum() == len(tets) and counts.min() > 0.8 * counts.max(), counts

# (3) mailbox handles travel like this in bench.py: 64 opaque bytes per rank, gathered in rank order
handle = bytes([rank]) * 64
handles = [None] * world
dist.all_gather_object(handles, handle)
assert handles == [bytes([r]) * 64 for r in range(world)]
dist.barrier()
dist.destroy_process_group()
print("rank %d ok" % rank)
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_over_gloo(tmp_path):
    build = os.path.join(HERE, "_build")
    os.makedirs(build, exist_ok=True)
    lib = os.path.join(build, "libhostscene.so")
    src = [os.path.join(HERE, "host_scene.cpp"), os.path.join(ROOT, "soft-body-simulator_b200/csrc/scene_build.cpp")]
    hdr = os.path.join(ROOT, "soft-body-simulator_b200/csrc/scene_build.h")
    if not os.path.exists(lib) or any(os.path.getmtime(lib) < os.path.getmtime(f) for f in src + [hdr]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", lib] + src)
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   SBS_ROOT=ROOT, SBS_HOSTSCENE=lib, GLOO_SOCKET_IFNAME="lo")
        procs.append(subprocess.Popen([sys.executable, str(worker)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out)
        assert "rank %d ok" % rank in out
