"""A/B timing of the full-likelihood contact pass variants (development aid).
python scripts/bench_full.py c2|c4 [reps]  -- every variant in one process (environment read at context creation)."""
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
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
def _v(spec):
    """unroll:minb:sub[:stab]  (stab = GRAAL_WIN_STAB: law table of the windowed pass from shared memory 1 / through L1 0)"""
    u, m, sb = spec.split(":")[:3]
    d = dict(GRAAL_FULL_WIN="1", GRAAL_WIN_UNROLL=u, GRAAL_WIN_MINB=m, GRAAL_WIN_SUB=sb)
    if len(spec.split(":")) > 3:
        d["GRAAL_WIN_STAB"] = spec.split(":")[3]
    return d


variants = [dict(GRAAL_FULL_WIN="0")] + [_v(a) for a in (sys.argv[3:] or ["8:4:2", "8:4:4"])]
if cfg == "c4":
    from graal_b200.level import synthetic_roofline_level
    inp, lists, tables, info = synthetic_roofline_level(device="cuda")
    mk = lambda: sampler.from_inputs(inp, device=0, rng=np.random.RandomState(1), device_contact_lists=lists, proposal_tables=tables)
    params = ([1.0, 9.6, -1.5, 3.0, 800.0], info["d_max_kb"])
else:
    sys.path.insert(0, ROOT)
    import bench as B
    pyr, inp, name = B.build_level("c2", 1)
    mk = lambda: sampler.from_inputs(inp, device=0, rng=np.random.RandomState(1))
    params = B.model_params(pyr)
ref = None
sched = None
for v in variants:
    for k in ("GRAAL_FULL_WIN", "GRAAL_WIN_UNROLL", "GRAAL_WIN_MINB", "GRAAL_WIN_SUB", "GRAAL_WIN_STAB"):
        os.environ.pop(k, None)
    os.environ.update(v)
    g = mk()
    g.set_parameters(*params)
    vals = [g.eval_likelihood()]
    # evolve the genome with a few real MCMC steps (same draws for every variant), then evaluate again
    rng = np.random.RandomState(7)
    for it in range(12):
        g.step_max_likelihood(int(rng.randint(int(g.n_new_frags))), 3)
    g.modify_gl_cuda_buffer()
    vals.append(g.eval_likelihood())
    _lib.check(g.lib.graal_profile_enable(g.ctx, 1))
    for it in range(reps):
        g.eval_likelihood()
    out = {}
    for kname, kid in (("FULL_CONTACTS", 0), ("FULL_BAND", 1), ("FULL_WINDOWS", 6)):
        tot, cnt = C.c_double(), C.c_longlong()
        _lib.check(g.lib.graal_profile_read(g.ctx, kid, C.byref(tot), C.byref(cnt), 1))
        out[kname] = tot.value / max(1, cnt.value)
    _lib.check(g.lib.graal_profile_enable(g.ctx, 0))
    E, W = g.n_contacts, int(g.init_n_sub_frags)
    gbs = (8 * E + 40 * W + 8) / (out["FULL_CONTACTS"] * 1e-3) / 1e9
    if ref is None:
        ref = vals
    rel = [abs(a - b) / abs(b) for a, b in zip(vals, ref)]
    print(v, "contacts %.4f ms (%.0f GB/s, %.3f of 6539.9)  windows %.4f ms  band %.4f ms  values %r rel diff vs first %r" %
          (out["FULL_CONTACTS"], gbs, gbs / 6539.9, out["FULL_WINDOWS"], out["FULL_BAND"], vals, rel), flush=True)
    g.free_gpu()
