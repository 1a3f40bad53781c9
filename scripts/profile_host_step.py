"""cProfile of the host side of sampler.step_max_likelihood on the C2 level (run on a GPU box):
where the time between two device round trips goes (profiles/README.md)."""
import cProfile, pstats, io, sys, time
import numpy as np
sys.argv = ["bench.py"]
sys.path.insert(0, "/root/repo")
import bench
from graal_b200.sampler import sampler
import torch
pyr, inp, name = bench.build_level("c2", 1)
g = sampler.from_inputs(inp, device=0, rng=np.random.RandomState(1000))
p, d_max = bench.model_params(pyr)
g.set_parameters(p, d_max)
g.init_likelihood()
rs = np.random.RandomState(4242)
frags = rs.randint(0, int(g.n_new_frags), size=400)
for it in range(50):
    g.step_max_likelihood(int(frags[it]), 3)
torch.cuda.synchronize()
pr = cProfile.Profile()
t = time.time()
pr.enable()
for it in range(50, 350):
    g.step_max_likelihood(int(frags[it]), 3)
pr.disable()
torch.cuda.synchronize()
print("ms/step", (time.time() - t) / 300 * 1e3)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
print(s.getvalue()[:4500])
