"""Host-side phases of a proof (ZKB_HOST_PROFILE=1): begin / enqueue / wait / assemble, for a tiny and a large shape."""
import sys, os, time
os.environ["ZKB_HOST_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_stark_project_b200 as Z
from zk_stark_project_b200 import lib as L
from tests import common as T

ctx = L.Context(0)
p = T.aggregation_prover(16, Z.ProofOptions.reference())
tr = p.build_trace()
air = ctx.prepare(p.describe(tr))
data = np.ascontiguousarray(tr.data)
for i in range(6):
    t0 = time.perf_counter()
    ctx.prove_host(air, data.ctypes.data)
    print("agg wall us %.1f" % ((time.perf_counter() - t0) * 1e6), {k: round(v * 1e3, 1) for k, v in ctx.stage_times().items() if v > 0}, flush=True)
n = 1 << 16
d2 = T.random_felts(240 * n, 5).reshape(240, n, 2)
air2 = ctx.prepare(T.synthetic_training_air(n, Z.ProofOptions.reference(), d2))
pin = L.PinnedBuffer(d2.nbytes)
pin.view()[:] = d2.reshape(-1).view(np.uint8)
dptr = ctx.upload_trace(pin.ptr, 240, n)
for i in range(4):
    t0 = time.perf_counter()
    ctx.prove_device(air2, dptr)
    print("train wall us %.1f" % ((time.perf_counter() - t0) * 1e6), flush=True)
