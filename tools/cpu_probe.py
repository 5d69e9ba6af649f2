import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--torch" in sys.argv:
    import torch  # noqa
from bench import cpu_sumcheck_times
from oracle import coracle as C
thr = int(os.environ.get("PROBE_THREADS", C.max_threads()))
ts = cpu_sumcheck_times(22, 2, 1, thr, 3)
print("torch" if "--torch" in sys.argv else "plain", "threads", thr, "max", C.max_threads(), "cpus", os.cpu_count(),
      "affinity", len(os.sched_getaffinity(0)), [round(t, 3) for t in ts], flush=True)
