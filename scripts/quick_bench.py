"""Quick device-time probe of objective + gradient at a few sizes (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aqc_research_b200 import circuit_structures as cs, utils
from aqc_research_b200.engine import SvWorkspace
from aqc_research_b200.parametric_circuit import TrotterAnsatz

sizes = [(int(a.split(":")[0]), int(a.split(":")[1])) for a in sys.argv[1:]] or [(20, 2), (24, 4), (28, 4)]
for n, L in sizes:
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, L), True)
    np.random.seed(n)
    th = utils.rand_thetas(circ.num_thetas)
    ws = SvWorkspace(circ, num_slots=4)
    ws.fill_random(0, 1)
    idx = np.array([0] + [1 << q for q in range(n)], dtype=np.int64)
    P = (n - 1) * L + n // 2
    for it in range(3):
        t0 = time.perf_counter()
        hs = ws.objective(th, 0, 1, idx)
        ms_o = ws.last_kernel_ms
        g = ws.grad(th, x_basis=0, z0=1, w=2, z=3)
        ms_g = ws.last_kernel_ms
        wall = time.perf_counter() - t0
    tot = (ms_o + ms_g) * 1e-3
    bytes_eval = 96.0 * 2**n * P
    flops = (104.0 * (circ.num_blocks + circ.half_layer_num_blocks) + 78.0 * n) * 2**n
    print(f"n={n} L={L} T={circ.num_thetas} passes(grad,dag)=({ws.num_passes(0)},{ws.num_passes(2)}) "
          f"obj {ms_o:.3f} ms grad {ms_g:.3f} ms wall {wall*1e3:.3f} ms -> {1/tot:.3f} evals/s; "
          f"pair-run GB/s {bytes_eval/tot/1e9:.0f} ({bytes_eval/tot/6540.8e9:.2f} of measured HBM); "
          f"{flops/tot/1e12:.2f} TFLOP/s fp64", flush=True)
    ws.close()
