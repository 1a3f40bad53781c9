"""A/B timing of the delta kernels (development aid): python scripts/bench_delta.py c2|c4 [steps]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from graal_b200 import _lib
from graal_b200.sampler import sampler, CUR

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
variants = [dict(a.split("=") for a in v.split(",")) for v in sys.argv[3:]] or [dict(GRAAL_DELTA_REL="0", GRAAL_BAND_FAST="0"), dict()]
if cfg == "c4":
    from graal_b200.level import synthetic_roofline_level
    inp, lists, tables, info = synthetic_roofline_level(device="cuda")
    mk = lambda: sampler.from_inputs(inp, device=0, rng=np.random.RandomState(1), device_contact_lists=lists, proposal_tables=tables)
    params = ([1.0, 9.6, -1.5, 3.0, 800.0], info["d_max_kb"])
else:
    import bench as B
    pyr, inp, name = B.build_level("c2", 1)
    mk = lambda: sampler.from_inputs(inp, device=0, rng=np.random.RandomState(1))
    params = B.model_params(pyr)
n = int(inp.n_new_frags)
frags = np.random.RandomState(4242).permutation(n)[:steps + 5]
ref = None
for v in variants:
    for k in list(os.environ):
        if k.startswith("GRAAL_"):
            os.environ.pop(k)
    os.environ.update(v)
    g = mk()
    g.set_parameters(*params)
    g.modify_gl_cuda_buffer()
    fA = int(frags[0]); nb = g.return_neighbours(fA, 3); nb.sort()
    g.score_neighbours(fA, nb)
    first = g._fetch()[16:16 + 13 * len(nb)].copy()
    traj = []
    for it in range(5):
        traj.append(g.step_max_likelihood(int(frags[it]), 3)[5:7])
    _lib.check(g.lib.graal_profile_enable(g.ctx, 1))
    for it in range(5, 5 + steps):
        traj.append(g.step_max_likelihood(int(frags[it]), 3)[5:7])
    out = {}
    for kname, kid in _lib.KERNELS.items():
        tot, cnt = C.c_double(), C.c_longlong()
        _lib.check(g.lib.graal_profile_read(g.ctx, kid, C.byref(tot), C.byref(cnt), 1))
        out[kname] = tot.value / max(1, cnt.value)
    _lib.check(g.lib.graal_profile_enable(g.ctx, 0))
    t0 = time.time()
    torch.cuda.synchronize()
    g.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(g.stream)
    for it in range(5, 5 + steps):
        g.step_max_likelihood(int(frags[it % len(frags)]), 3)
    ev1.record(g.stream); g.sync(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    if ref is None:
        ref = (first, traj)
    d = np.abs(first - ref[0]).max() / np.abs(ref[0]).max()
    print(v, "e2e step %.3f ms | " % ms + "  ".join("%s %.4f" % (k, x) for k, x in out.items()),
          "| first-proposal max diff / max |delta| %.2e  same trajectory %s" % (d, traj == ref[1]), flush=True)
    g.free_gpu()
