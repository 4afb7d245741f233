import sys, os, json
sys.path.insert(0, "/root/repo")
import bench
from optimization_b200.device import Context
ctx = Context(0)
print(json.dumps(bench.so3_c5(ctx)))
