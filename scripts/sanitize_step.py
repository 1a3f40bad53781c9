"""compute-sanitizer target: BASELINE config C1 level 2 (563 bins), lanes + CUDA graphs on -- full likelihood, four
step_max_likelihood steps (3 proposals x 13 candidates each, commit), one nuisance-parameter step, one MH step, the validation
step, the older step, local_flip.  GRAAL_WIN_STAB=1 forces the shared-memory law table of the full pass."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from graal_b200.level import yeast_shaped_pyramid, prepare_sampler_inputs
from graal_b200.sampler import sampler
import bench as B

pyr = yeast_shaped_pyramid(n_levels=3)
inp = prepare_sampler_inputs(pyr, 2)
g = sampler.from_inputs(inp, rng=np.random.RandomState(4))
p, dm = B.model_params(pyr)
g.set_parameters(p, dm)
g.bins = np.arange(10.0, 510.0, 10.0)
g.init_likelihood()
for fA in (5, 77, 300, 412):
    print(g.step_max_likelihood(fA, 3)[:7])
print(g.step_nuisance_parameters()[:7])
g.set_jumping_distributions_parameters(3)
print(g.step_mtm(40)[:3])
# the rest of the class surface (variants.py): validation step, older proposal rule, local_flip
print(g.debug_step_max_likelihood(123, 2)[:3])
print(g.step_max_likelihood_4_visu(200, 2)[:3])
g.local_flip(250, 13, int(g.modify_gl_cuda_buffer()))
print("launches", g.gpu_launches)
g.free_gpu()
