"""L-BFGS end-to-end extra of bench.py alone (GPU)."""
import json, sys
sys.path.insert(0, ".")
import bench
for n, l in ((5, 2), (12, 2), (12, 6)):
    print(json.dumps(bench.measure_lbfgs(n=n, layers=l)))
