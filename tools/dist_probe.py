"""Per-round anatomy of ONE index-sharded sumcheck on rank 0 (torchrun --nproc-per-node N tools/dist_probe.py [log_n])."""
import ctypes, os, sys, time, pathlib
import numpy as np
import torch
import torch.distributed as dist
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import jolt_b200
from jolt_b200 import Polynomial
from jolt_b200.api import _p
from jolt_b200.dist import init_comm, prove_sharded, sharded_claim

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sess = jolt_b200.Session(local, cuda_stream=stream.cuda_stream)
init_comm(sess, dist)
n = 1 << lg
g = torch.Generator(device="cuda").manual_seed(0xB200 + 16 * rank)
base = []
for j in range(2):
    t = torch.randint(0, 2 ** 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, 3] &= (1 << 60) - 1
    base.append(t)
claim = sharded_claim(sess, [Polynomial.wrap_device(sess, b.data_ptr(), n) for b in base], dist)
best = None
for rep in range(6):
    bufs = [b.clone() for b in base]
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    prove_sharded(sess, [Polynomial.wrap_device(sess, b.data_ptr(), n) for b in bufs], claim, 7, raw=True)
    dt = time.perf_counter() - t0
    log = np.zeros((64, 8), dtype=np.uint64)
    cnt = ctypes.c_size_t()
    sess.check(sess.lib.jb_ctx_run_log(sess.h, _p(log), 64, ctypes.byref(cnt)))
    if best is None or dt < best[0]:
        best = (dt, log[: cnt.value].astype(np.int64))
if rank == 0:
    dt, log = best
    print(f"world {dist.get_world_size()} 2^{lg}/rank: {dt * 1e6:.1f} us wall, last run had {len(log)} commands")
    print(" cmd  pass_us  dev_idle_us  host_wait_us")
    for k in range(len(log)):
        idle = (log[k + 1, 0] - log[k, 1]) / 1e3 if k + 1 < len(log) else float("nan")
        print(f"{k:4d} {(log[k, 1] - log[k, 0]) / 1e3:8.2f} {idle:12.2f} {(log[k, 3] - log[k, 2]) / 1e3:13.2f}")
dist.barrier()
dist.destroy_process_group()
