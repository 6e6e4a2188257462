"""Quick device-time probe of the MPS objective + gradient (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200.mps_engine import MpsWorkspace
from aqc_research_b200.parametric_circuit import TrotterAnsatz

cases = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(12, 16, 4), (30, 32, 6), (50, 64, 20)]
for n, chi, L in cases:
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, L), True)
    rng = np.random.RandomState(n)
    th_t = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    ws = MpsWorkspace(circ, num_slots=4, chi_max=chi, trunc_thr=1e-6)
    neel = sum(1 << q for q in range(0, n, 2))
    ws.set_product(0, neel)
    t0 = time.perf_counter()
    ws.apply(th_t, 0, 0, dagger=False)       # target: random-circuit state truncated to chi
    t_target = time.perf_counter() - t0
    dims = [l.size for l in ws.download(0)[1]]
    idx = np.array([neel] + [neel ^ (1 << q) for q in range(n)], dtype=np.int64)
    for it in range(2):
        t0 = time.perf_counter()
        hs = ws.objective(th, 0, 1, idx)
        ms_o, l_o = ws.last_kernel_ms, ws.last_num_launches
        g = ws.grad(th, x_basis=neel, z0=1, w=2, z=3)
        ms_g, l_g = ws.last_kernel_ms, ws.last_num_launches
        wall = time.perf_counter() - t0
    wd = [l.size for l in ws.download(2)[1]]
    print(f"n={n} chi={chi} L={L} T={circ.num_thetas} target bonds max {max(dims)} (build {t_target*1e3:.1f} ms) "
          f"w bonds max {max(wd)} | obj {ms_o:.2f} ms ({l_o} launches) grad {ms_g:.2f} ms ({l_g} launches) "
          f"wall {wall*1e3:.1f} ms -> {1/wall:.3f} evals/s; |g|={np.linalg.norm(g):.3e} norm z0={abs(ws.dot(1,1)):.6f}", flush=True)
    import ctypes as ct
    from aqc_research_b200 import _lib
    buf = np.zeros(256, dtype=np.int32)
    cnt = _lib.load().aqc_mps_debug_sweeps(ws.handle, buf.ctypes.data_as(_lib.c_int32_p), buf.size)
    print('   Jacobi sweeps of the last step:', sorted(set(buf[:cnt].tolist())), flush=True)
    ws.close()
