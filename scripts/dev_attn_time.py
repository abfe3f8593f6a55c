"""Dev helper: time the fused reference attention at the 64x64 level (bench.kernel_rooflines leg (2))."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

peaks = bench.load_peaks() if hasattr(bench, "load_peaks") else {}
r = bench.kernel_rooflines(torch.device("cuda:0"), peaks)
a = r["ref_attention"]
print("V5=%s: %s" % (os.environ.get("AP_ATTENTION_V5", "-"), json.dumps({k: a[k] for k in a if k != "kernel"})))
