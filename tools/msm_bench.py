"""G1 MSM throughput on one B200 (BASELINE config 3): terms/s at 2^16..2^24, synthetic bases
(i+1)*G generated on the device, uniform 253-bit scalars resident in HBM; the result is checked
against the closed form at every size. CPU = the C restatement's Pippenger on all host cores."""
import json, sys, time, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import jolt_b200
from jolt_b200 import G1Bases, Polynomial, g1_jacobian_to_affine
from oracle import bn254 as O
from oracle import coracle as C
from test_gpu_msm import _weighted_sum, G

sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [16, 20, 22, 24]
cpu = "--cpu" in sys.argv
sess = jolt_b200.Session(0)
out = []
for lg in sizes:
    n = 1 << lg
    t0 = time.perf_counter()
    bases = G1Bases.generate_multiples(sess, G, n)
    gen_s = time.perf_counter() - t0
    pre_s = 0.0
    if "--pre" in sys.argv:
        t0 = time.perf_counter()
        bases.precompute()
        pre_s = time.perf_counter() - t0
    sc = C.rand_limbs(0x5CA1A2, n)
    tab = Polynomial.new(sess, sc)
    sess.synchronize()
    sess.timing_enable(True, 0)
    res = bases.msm(tab)            # warm-up + correctness
    ok = g1_jacobian_to_affine(res) == O.g1_scalar_mul(O.G1_GEN, _weighted_sum(sc))
    sess.timing_collect()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        bases.msm(tab)
        ts.append(time.perf_counter() - t0)
    acc = [t["ms"] for t in sess.timing_collect() if t["kind"] == "msm_accumulate"]
    sess.timing_enable(False)
    t0 = time.perf_counter()
    bases.msm(sc)                   # host scalars: + H2D of 32 B/term
    e2e = time.perf_counter() - t0
    rec = dict(kind="msm", log_n=lg, ok=bool(ok), ms=round(min(ts) * 1e3, 3), mterms_per_s=round(n / min(ts) / 1e6, 1),
               accumulate_ms=round(sum(acc) / max(len(acc), 1), 3), e2e_ms=round(e2e * 1e3, 3), srs_generate_s=round(gen_s, 2), precompute_s=round(pre_s, 2))
    if cpu and lg <= 20:
        xy = bases.affine()
        t0 = time.perf_counter()
        C.g1_msm_pippenger(xy, sc, 0, C.max_threads())
        rec["cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        rec["cpu_cores"] = C.max_threads()
    print(json.dumps(rec), flush=True)
    out.append(rec)
    bases.free(); tab.free()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
with open(ROOT / "gpurun_out" / "msm_bench.jsonl", "w") as f:
    for r in out:
        f.write(json.dumps(r) + "\n")
