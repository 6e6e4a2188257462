"""development aid: coordinate-descent throughput (bench.measure_cd7 alone)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print(json.dumps(bench.measure_cd7(with_cpu=False)))
