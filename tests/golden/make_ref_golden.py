"""Golden vectors computed by the REFERENCE'S OWN KERNELS (kernels3.cu compiled for the host by
oracle/ref_emu, see oracle/ref_emu/build.py) -> tests/golden/ref_kernels.npz.

Run where /root/reference exists:   python tests/golden/make_ref_golden.py
The fixture travels to machines without the reference (the GPU box); tests/test_reference_kernels.py checks the
NumPy oracle -- and, on a GPU, the device path -- against it."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import helpers as H                                        # noqa: E402
from oracle import mutations as M, likelihood as L, ref_emu as R   # noqa: E402
from graal_b200.level import prepare_sampler_inputs, build_synthetic_pyramid   # noqa: E402

PYR = dict(contig_bp=[300_000, 200_000, 150_000, 90_000, 40_000, 6_000], n_frags0=600, n_levels=3, seed=11,
           cis_rowsum=300.0, v_inter=0.05)
LEVEL = 2


def pyramid():
    return build_synthetic_pyramid(PYR["contig_bp"], PYR["n_frags0"], PYR["n_levels"], seed=PYR["seed"],
                                   cis_rowsum=PYR["cis_rowsum"], v_inter=PYR["v_inter"])


def move_cases(o, rng, n_rounds):
    """(state, call, reference result) for every mutation kernel from scrambled states."""
    n = o.n_new_frags
    src, spec, out, side = [], [], [], []
    for rnd in range(n_rounds):
        H.scramble(o, rng, 15)
        max_id = int(o.modify_gl_cuda_buffer())
        if rnd >= 2:                                    # close a contig into a circle: paste its two ends
            for c in np.unique(o.cur["id_c"]):
                bins = np.nonzero(o.cur["id_c"] == c)[0]
                if bins.size >= 4 and o.cur["circ"][bins[0]] == 0:
                    head = int(bins[np.argmin(o.cur["pos"][bins])]); tail = int(bins[np.argmax(o.cur["pos"][bins])])
                    new = M.copy_slot(o.cur)
                    M.paste_contigs(new, o.cur, tail, head, max_id)
                    if new["circ"][head] == 1:
                        for k in M.FIELDS:
                            o.cur[k][:] = new[k]
                        break
            max_id = int(o.modify_gl_cuda_buffer())
        fA, fB = (int(x) for x in rng.choice(n, 2, replace=False))
        if rnd >= 2 and rnd % 2 == 0:                   # a proposal inside the circle
            cb = np.nonzero(o.cur["circ"] == 1)[0]
            if cb.size >= 2:
                fA, fB = int(cb[0]), int(cb[-1])
        sentinel = M.copy_slot(o.cur)
        sentinel["pos"][:] = 12345

        def run(op, s, a=0, b=0, aux=0, mx=0, with_ids=False):
            d = M.copy_slot(sentinel)
            ids = np.zeros(n, np.int32) if with_ids else None
            R.move(op, d, s, a, b, aux=aux, max_id=mx, ids=ids)
            src.append(R.pack(s)); spec.append((R.OPS[op], a, b, aux, mx)); out.append(R.pack(d))
            side.append(ids if with_ids else np.zeros(n, np.int32))
            return d, ids

        run("flip", o.cur, fA)
        run("swap_activity", o.cur, fA, mx=max_id)
        run("simple_copy", o.cur)
        for up in (0, 1):
            run("split", o.cur, fA, aux=up, mx=max_id, with_ids=True)
        run("paste", o.cur, fA, fB, mx=max_id)
        pop, pid = run("pop_out", o.cur, fA, mx=max_id, with_ids=True)
        mx2 = int(pid.max())
        for k in (1, 2, 3, 4):
            for ori in (1, -1):
                run("pop_in_%d" % k, pop, fA, fB, aux=ori, mx=mx2)
    return np.array(src), np.array(spec, dtype=np.int32), np.array(out), np.array(side), M.copy_slot(sentinel)


def likelihood_cases(o, rng, n_states, n_props):
    n = o.n_new_frags
    states, fulls, props, deltas = [], [], [], []
    for st in range(n_states):
        H.scramble(o, rng, 12)
        max_id = o.modify_gl_cuda_buffer()
        cur = R.evaluate_likelihood(o.cur, o.lv, o.param_simu)
        states.append(R.pack(o.cur)); fulls.append(cur)
        for it in range(n_props):
            fA, fB = (int(x) for x in rng.choice(n, 2, replace=False))
            M.perform_modifications(o.ws, o.cur, fA, fB, max_id)
            no_rep, rep = o.candidate_index_sets(fA, fB)
            d = [R.sub_compute_likelihood(o.ws.collector[j], o.lv, o.param_simu, cur, no_rep, rep, o.uniq_frags) for j in range(13)]
            props.append((st, fA, fB, int(max_id))); deltas.append(d)
    return np.array(states), np.array(fulls), np.array(props, dtype=np.int32), np.array(deltas)


def main():
    pyr = pyramid()
    out = {}
    o = H.make_oracle(prepare_sampler_inputs(pyr, LEVEL), pyr)
    src, spec, res, side, sentinel = move_cases(o, np.random.RandomState(21), 5)
    out.update(mv_src=src, mv_spec=spec, mv_out=res, mv_side=side, mv_sentinel_pos=np.int32(12345))
    for tag, kw in (("u", {}), ("r", {"allow_repeats": True})):
        o = H.make_oracle(prepare_sampler_inputs(pyr, LEVEL, **kw), pyr)
        states, fulls, props, deltas = likelihood_cases(o, np.random.RandomState(6), 3, 3)
        out.update({"ll_%s_states" % tag: states, "ll_%s_full" % tag: fulls, "ll_%s_props" % tag: props, "ll_%s_deltas" % tag: deltas})
    s = np.array([0.0, 0.3, 1.0, 7.7, 55.5, 300.0, 999.0, 1500.0], dtype=np.float32)
    out.update(sc_s=s, sc_rippe=R.rippe_contacts(s, o.param_simu), sc_rippe_circ=R.rippe_contacts_circ(s, np.full_like(s, 2000.0), o.param_simu),
               sc_params=R.params8(o.param_simu))
    path = os.path.join(HERE, "ref_kernels.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(spec), "mutation calls,", sum(len(out["ll_%s_deltas" % t]) for t in "ur"), "proposals")


if __name__ == "__main__":
    main()
