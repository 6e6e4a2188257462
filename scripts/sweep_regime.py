"""Development aid: which tile heuristic (L2-resident vs streaming) wins at n = 21..24."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import bench

CFG = {
    "l2res": dict(AQC_TILE_BITS_GRAD="10", AQC_TILE_BITS_APPLY="11", AQC_TILE_LOW_BITS="2", AQC_TILE_LOW_BITS_APPLY="1"),
    "large": dict(AQC_TILE_BITS_GRAD="11", AQC_TILE_BITS_APPLY="12", AQC_TILE_LOW_BITS="4", AQC_TILE_LOW_BITS_APPLY="3"),
    "default": {},
}
for n in [int(a) for a in sys.argv[1:]] or [21, 22, 23, 24]:
    for name, env in CFG.items():
        for k in ("AQC_TILE_BITS_GRAD", "AQC_TILE_BITS_APPLY", "AQC_TILE_LOW_BITS", "AQC_TILE_LOW_BITS_APPLY"):
            os.environ.pop(k, None)
        os.environ.update(env)
        r = bench.measure_gpu(n, 2, 5, 3, 0, True)
        ms = float(np.mean(r["step_ms"]))
        print(f"n={n} {name:8s} {1e3 / ms:9.2f} evals/s  obj {np.mean(r['obj_ms']):.3f} ms grad {np.mean(r['grad_ms']):.3f} ms "
              f"passes {r['passes_grad']}/{r['passes_dag']}", flush=True)
