"""Development aid: MPS objective + gradient on a PHYSICAL workload (Trotter-evolved Neel target,
ansatz angles near the Trotter point), the regime the time-evolution driver runs in."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200.mps_engine import MpsWorkspace
from aqc_research_b200.parametric_circuit import TrotterAnsatz
from aqc_research_b200.model_sp_lhs.trotter import trotter as trotop

n, chi, L, T = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else (50, 64, 20, 4.0)
circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, L), True)
th_t = trotop.init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=T, delta=1.0)
rng = np.random.RandomState(1)
th = th_t + 0.01 * np.pi * (2 * rng.rand(circ.num_thetas) - 1)
ws = MpsWorkspace(circ, num_slots=4, chi_max=chi, trunc_thr=1e-6)
neel = sum(1 << q for q in range(0, n, 2))
ws.set_product(0, neel)
ws.apply(th_t, 0, 0, dagger=False)
dims = [l.size for l in ws.download(0)[1]]
idx = np.array([neel] + [neel ^ (1 << q) for q in range(n)], dtype=np.int64)
for it in range(2):
    t0 = time.perf_counter()
    hs = ws.objective(th, 0, 1, idx)
    ms_o = ws.last_kernel_ms
    g = ws.grad(th, x_basis=neel, z0=1, w=2, z=3)
    ms_g = ws.last_kernel_ms
    wall = time.perf_counter() - t0
print(f"phys n={n} chi={chi} L={L} T={T}: target bonds max {max(dims)} | obj {ms_o:.2f} ms grad {ms_g:.2f} ms wall {wall*1e3:.1f} ms "
      f"-> {1/wall:.3f} evals/s; fidelity {abs(np.ravel(hs)[0])**2:.6f} |g|={np.linalg.norm(g):.3e}", flush=True)
import ctypes as ct
from aqc_research_b200 import _lib
buf = np.zeros(256, dtype=np.int32)
cnt = _lib.load().aqc_mps_debug_sweeps(ws.handle, buf.ctypes.data_as(_lib.c_int32_p), buf.size)
print('   Jacobi sweeps of the last step:', sorted(set(buf[:cnt].tolist())), flush=True)
ws.close()
