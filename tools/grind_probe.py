"""Stage times of a tiny proof at several grinding factors (diagnostic for the device-side PoW search)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_stark_project_b200 as Z
from zk_stark_project_b200 import lib as L
from tests import common as T

ctx = L.Context(0)
for bits in (0, 8, 12, 16, 18, 20, 21):
    p = T.mimc_prover(4, 256, Z.ProofOptions(40, 8, bits, Z.FieldExtension.NONE, 16, 7))
    tr = p.build_trace()
    air = p.describe(tr)
    data = np.ascontiguousarray(tr.data)
    ctx.prove_host(air, data.ctypes.data)
    t0 = time.time()
    proof, ts = ctx.prove_host(air, data.ctypes.data)
    dt = time.time() - t0
    st = ctx.stage_times()
    print(bits, "nonce", ts.pow_nonce, "wall ms %.3f" % (dt * 1e3), {k: round(v, 3) for k, v in st.items() if v > 0}, flush=True)
