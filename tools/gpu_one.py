"""One launch of the batched kernel on config-2 LPs (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dantzig_b200 import generate, Template, Batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
tpr = int(sys.argv[2]) if len(sys.argv) > 2 else 0
w = generate.config2(B)
t = Template(w.structure)
b = Batch(t, w.B, worker_warps=tpr)
b.upload(w.theta); b.solve(); r = b.download(light=True)
print("B", B, b.launch_info(), "ms %.2f" % b.kernel_ms(), "optimal", int((r.status == 0).sum()), "pivots", int(r.pivots.sum()))
