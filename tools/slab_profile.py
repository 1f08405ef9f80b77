"""One rank of the periodic box workload (BASELINE configs[4] share of one GPU: the slab exchanges ghosts with itself through
the peer-to-peer path) for an ncu launch list of the slab kernels:

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file out.csv python tools/slab_profile.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bench_multi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--per-gpu", type=int, default=160)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
args = argparse.Namespace(per_gpu=a.per_gpu, steps=a.steps, warmup=a.warmup, dr=bench.DR_10M, workload="box", no_box=True)
os.environ.setdefault("SP_BENCH_NO_SAMPLER", "1")
r = bench_multi.run_slab_workload(args, "box", bench.UNIT, bench.ClockSampler, e2e=False, breakdown=False)
print({k: r[k] for k in ("workload", "value", "ms_per_step", "particles", "ghosts_per_gpu", "gpu_launches")})
