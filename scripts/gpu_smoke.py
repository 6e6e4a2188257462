"""Quick GPU smoke of the engine on small states (used under a short `timeout` before longer GPU jobs)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200.engine import SvWorkspace
from aqc_research_b200.parametric_circuit import TrotterAnsatz
from oracle import sv_oracle as O

ns = [int(a) for a in sys.argv[1:]] or [6, 10, 13]
for n in ns:
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    rng = np.random.RandomState(n)
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    y = rng.rand(2**n) + 1j * rng.rand(2**n)
    y /= np.linalg.norm(y)
    t0 = time.time()
    ws = SvWorkspace(circ, num_slots=4)
    ws.upload(0, y)
    idx = O.basis_state_indices(n)
    hs = ws.objective(th, 0, 1, idx)[0]
    z0 = O.apply_v(circ, th, y, dagger=True)
    e1 = np.linalg.norm(hs - z0[idx]) / np.linalg.norm(z0[idx])
    g = ws.grad(th, x_basis=int(idx[1]), z0=1, w=2, z=3)[0]
    e = np.zeros(2**n, dtype=complex); e[idx[1]] = 1
    ref = O.grad_sweep(circ, th, e, z0)
    e2 = np.linalg.norm(g - ref) / np.linalg.norm(ref)
    print(f"n={n} hs err {e1:.2e} grad err {e2:.2e} launches {ws.last_num_launches} ({time.time()-t0:.2f}s)", flush=True)
    ws.close()
    assert e1 < 1e-10 and e2 < 1e-10
print("smoke ok")
