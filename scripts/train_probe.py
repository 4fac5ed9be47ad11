import json, sys
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.train_block(bench.load_peaks()), indent=1))
