"""development aid: sketching generator timings (bench.measure_sketch alone)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for n, m in ((10, 32), (12, 64)):
    print(json.dumps(bench.measure_sketch(n=n, m=m, with_cpu=True)))
