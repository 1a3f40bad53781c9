#!/usr/bin/env python
"""bench.py -- GRAAL MCMC move-scoring hot path on B200.

Metric (BASELINE.json): MCMC move log-likelihood evaluations / s.  One "step" is one
``step_max_likelihood`` of the reference sampler (cuda_lib_gl.py:1793-1980) for one bin: relabel,
full log-likelihood, ``n_neighbours`` proposals x 13 candidate structures built and scored, one
candidate committed -> 13 * n_neighbours move evaluations.

Workload at N = 1: BASELINE config C2 -- synthetic T. reesei-shaped pyramid (77 contigs, 33 Mb,
100,000 level-0 fragments, factor 3), single chain, run at level 1 (33k bins, 100k sub-frags,
23.7 M stored contact entries = 190 MB of contact lists, larger than the 126 MB L2, re-streamed every
step).  N > 1: ``--chains-per-gpu`` independent replica chains per GPU (weak scaling) with a replica-exchange
all_gather of (loglik, temperature index) every --exchange-every steps.

    value     device-resident throughput: the recorded proposal schedule replayed with no host round trip
    e2e       the same schedule through the public sampler API (host RNG draws, D2H of scores / state)
    roofline  the contact-list pass of the full likelihood, CUDA events inside the library; ``roofline_delta``: the contact
              pass of the move deltas; ``roofline_c4``: the same two kernels on BASELINE config C4 (200 k bins, 237 M contacts)
    parity    the first timed step scored by the NumPy oracle on the host (39 deltas + the full likelihood) against the
              device's numbers for the same step; the run FAILS (exit code 1) outside the tolerance
    cpu_baseline / --impl reference: the NumPy oracle (sparse formulation) on all host cores, same schedule
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TMP = 13


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c4", "c5", "tiny"])
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--neighbours", type=int, default=3)
    ap.add_argument("--chains-per-gpu", type=int, default=1)
    ap.add_argument("--exchange-every", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the roofline_c4 block (C4 level generated and timed in the same run)")
    ap.add_argument("--no-original", action="store_true", help="skip the original_kernels block (kernels3.cu compiled for sm_100a, C1 level 1)")
    ap.add_argument("--incremental", action="store_true",
                    help="carry the likelihood of the current state from the committed candidate (full pass every 256 steps "
                         "only) instead of recomputing it every step as the reference does; NOT the default workload")
    ap.add_argument("--profile-only", action="store_true", help="short replay for ncu (no CPU baseline, no e2e)")
    return ap.parse_args()


def build_level(cfg, level):
    from graal_b200.level import yeast_shaped_pyramid, treesei_shaped_pyramid, build_synthetic_pyramid, prepare_sampler_inputs
    if cfg == "c2":
        pyr = treesei_shaped_pyramid(n_levels=max(2, level + 1))
        name = "C2: synthetic T. reesei-shaped pyramid (77 contigs, 33 Mb, 100k level-0 frags, factor 3), level %d, single chain" % level
    elif cfg == "c1":
        pyr = yeast_shaped_pyramid(n_levels=max(2, level + 1))
        name = "C1: synthetic S. cerevisiae-shaped pyramid (16 chr, 12 Mb, 5k frags, factor 3), level %d" % level
    else:
        pyr = build_synthetic_pyramid([300_000, 200_000, 150_000, 90_000, 40_000, 6_000], 600, max(2, level + 1),
                                      seed=11, cis_rowsum=300.0, v_inter=0.05)
        name = "tiny debug pyramid, level %d" % level
    return pyr, prepare_sampler_inputs(pyr, level), name


def model_params(pyr):
    from graal_b200.level import rippe_law
    law, A = pyr.spec["law"], pyr.spec["amplitude"]
    grid = np.linspace(1.0, 5000.0, 50000)
    above = grid[A * rippe_law(grid) > pyr.spec["v_inter"]]
    d_max = float(above[-1]) if above.size else 100.0
    return [law["kuhn"], law["lm"], law["slope"], law["d"], A], d_max


def make_config(name, n_bins, W, E, k_nb, n_contigs, chains, world, exchange_every, incremental=False):
    """The workload description, identical for both arms (the driver compares the two `config` objects)."""
    return {"workload": name, "bins": int(n_bins), "sub_frags": int(W), "contact_entries": int(E),
            "contact_list_MB": round(8 * E / 1e6, 1), "neighbours_per_step": int(k_nb), "candidates_per_neighbour": N_TMP,
            "state": "assembled genome (%d contigs)" % n_contigs,
            "schedule": "bins: RandomState(4242).permutation; neighbours: the reference's proposal rule on RandomState(1000 + chain)"
                        + ("; resident replay: rank 0's recorded schedule on every rank" if world > 1 else ""),
            "l2": "inputs larger than L2: the %.0f MB contact list is re-streamed every step" % (8 * E / 1e6),
            "full_likelihood": ("incremental: carried from the committed candidate, full pass every 256 steps (NOT the reference's schedule)"
                                if incremental else "recomputed every step, as the reference does"),
            "parallelism": "%d replica chain%s per GPU" % (chains, "" if chains == 1 else "s")
                           + (", replica-exchange all_gather every %d steps" % exchange_every if world * chains > 1 else "")}


def bin_schedule(n, total):
    sched_rng = np.random.RandomState(4242)          # the same bins on every rank: replicas do statistically identical work
    frags = sched_rng.permutation(n)[:total]
    if frags.size < total:
        frags = np.concatenate([frags, sched_rng.randint(0, n, size=total - frags.size)])
    return frags


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            if not any(a <= t <= b + 0.2 for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# the oracle on the host cores: one step = the 13 x k candidate deltas + the full likelihood of the current state
# ------------------------------------------------------------------------------------------------------------------
_W = {}


def _w_delta(task):
    """(fA, fB, j) -> (delta, mass) of candidate j of the proposal; the 13 candidates of a proposal are built once per worker."""
    from oracle import sparse as S, mutations as M
    fA, fB, j = task
    c = _W
    if c.get("prop") != (fA, fB):
        if "ws" not in c:
            c["ws"] = M.Workspace(c["n"])
        M.perform_modifications(c["ws"], c["cur"], fA, fB, c["max_id"])
        cur = c["cur"]
        c["bins_u"] = np.nonzero((cur["id_c"] == cur["id_c"][fA]) | (cur["id_c"] == cur["id_c"][fB]))[0]
        c["prop"] = (fA, fB)
    return S.sparse_delta(c["ws"].collector[j], c["cur"], c["lv"], c["par"], c["bins_u"], return_mass=True)


def _w_full(task):
    """One share of the full likelihood: a slice of the stored contacts and a group of contigs of the band mass."""
    from oracle import sparse as S
    c = _W
    if "geo" not in c:
        c["geo"] = S.Geo(c["cur"], c["lv"])
    kind, arg = task
    if kind == "contacts":
        t, _ = S.contact_terms(c["geo"], c["lv"], c["par"], np.arange(arg[0], arg[1]))
        return float(t.sum())
    return -S.band_mass(c["geo"], c["lv"], c["par"], arg)


def _w_noop(j):
    return j


class CpuStep:
    """The oracle (NumPy port of the reference semantics, sparse formulation) scoring whole steps of a level on the host:
    per step the 13 candidates of every proposal and the full likelihood of the current state, spread over `workers` forked
    processes (the candidates are independent, as on the reference's 13 streams; the workers inherit the read-only level
    and are started before any timed region).  The genome stays the initial one (nothing is committed): a bounded sample of
    the chain's work, same schedule."""

    def __init__(self, inp, pyr, workers):
        from oracle import mutations as M, sparse as S, likelihood as L
        import multiprocessing as mp
        self.S = S
        p, d_max = model_params(pyr)
        self.par = L.make_params(p[0], p[1], p[2], p[3], p[4], d_max, inp.mean_value_trans)
        self.lv = S.SparseLevel(inp.n_frags, inp.np_sub_frags_id, inp.np_sub_frags_len_bp, inp.np_sub_frags_accu,
                                inp.mean_squared_frags_per_bin, *inp.sub_coo)
        self.cur = {k: np.array(inp.S_o_A_frags[k], dtype=np.int32) for k in M.FIELDS}
        self.cur["ori"][:] = 1
        self.max_id = M.relabel_contigs(self.cur)
        self.workers = max(1, int(workers))
        _W.clear()
        _W.update(cur=self.cur, lv=self.lv, par=self.par, n=inp.n_new_frags, max_id=self.max_id)
        self.pool = mp.get_context("fork").Pool(self.workers)
        self.pool.map(_w_noop, range(self.workers))
        # shares of the full likelihood: contact slices of equal size, contigs grouped by band work (~ size)
        E = int(self.lv.rows.size)
        n_sl = 2 * self.workers
        cuts = np.linspace(0, E, n_sl + 1).astype(np.int64)
        geo = S.Geo(self.cur, self.lv)
        ids, cnt = np.unique(geo.id_c, return_counts=True)
        order = np.argsort(-cnt)
        groups = [[] for _ in range(min(n_sl, ids.size))]
        load = np.zeros(len(groups))
        for k in order:
            g = int(np.argmin(load)); groups[g].append(ids[k]); load[g] += cnt[k]
        self.full_tasks = [("contacts", (int(a), int(b))) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        self.full_tasks += [("band", np.nonzero(np.isin(geo.id_c, g))[0]) for g in groups if g]
        self.const = -(S.g0_mass(self.lv, self.par) + S.quirk_mass(self.cur, self.lv, self.par))

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool = None
        _W.clear()

    def step(self, fA, neighbours, with_full=True):
        """-> (deltas [(delta, mass)] in proposal-major order, full likelihood or None, seconds)."""
        t0 = time.time()
        tasks = [(int(fA), int(fB), j) for fB in neighbours for j in range(N_TMP)]
        r_d = self.pool.map_async(_w_delta, tasks, chunksize=1)
        r_f = self.pool.map_async(_w_full, self.full_tasks, chunksize=1) if with_full else None
        deltas = r_d.get()
        full = (sum(r_f.get()) + self.const) if with_full else None
        return deltas, full, time.time() - t0


class HostProposals:
    """The reference's proposal rule (return_neighbours, cuda_lib_gl.py:2295-2331) and the one uniform draw of the candidate
    choice on the host, consuming RandomState(1000 + chain) exactly as graal_b200.sampler does -- the reference arm follows
    the same (bin, neighbours) schedule without a GPU."""

    def __init__(self, inp, seed):
        from graal_b200.sampler import neighbour_tables
        self.rng = np.random.RandomState(seed)
        self.xk, self.pk = neighbour_tables(inp.level_coo, int(inp.n_frags), [], 10)
        self.id_d = np.asarray(inp.S_o_A_frags["id_d"])

    def neighbours(self, fA, delta):
        ori = int(self.id_d[fA])
        distri = self.pk[ori]
        n_max = min(min(10, delta), int(np.count_nonzero(distri)))
        nb = sorted(int(e) for e in self.rng.choice(self.xk[ori], n_max, p=distri, replace=False))
        self.rng.random_sample()                  # the draw of the sampled candidate
        return nb


def parity_block(deltas_cpu, full_cpu, deltas_gpu, full_gpu):
    """Device vs oracle on one step: the asserted bound is tests/helpers.delta_check's."""
    worst_mass, worst_rel, ok = 0.0, 0.0, True
    for (ref, mass), got in zip(deltas_cpu, deltas_gpu):
        err = abs(got - ref)
        worst_mass = max(worst_mass, err / max(mass, 1e-300))
        if abs(ref) > 0:
            worst_rel = max(worst_rel, err / abs(ref))
        ok &= err <= 1e-6 * abs(ref) + 1e-7 * mass + 1e-9
    full_rel = abs(full_gpu - full_cpu) / abs(full_cpu) if full_cpu else None
    ok &= full_rel is not None and full_rel <= 1e-7
    rank_ok = int(np.argmax([d[0] for d in deltas_cpu])) == int(np.argmax(deltas_gpu))
    return {"n_deltas": len(deltas_gpu), "max_err_over_mass": worst_mass, "max_rel": worst_rel, "full_rel": full_rel,
            "best_candidate_agrees": bool(rank_ok), "tolerance": "|err| <= 1e-6 |delta| + 1e-7 mass; full 1e-7 relative", "ok": bool(ok)}


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path = the NumPy oracle port (the reference is
    Python 2 + PyCUDA and cannot run here; its kernels need a GPU), on all host cores, same schedule, same step content."""
    pyr, inp, name = build_level(args.config, args.level)
    k_nb = args.neighbours
    workers = max(1, min(32, os.cpu_count() or 1))
    port = CpuStep(inp, pyr, workers)
    props = HostProposals(inp, 1000)
    total = args.warmup + args.steps
    frags = bin_schedule(int(inp.n_new_frags), total)
    vals, t_all, n_ev = [], [], 0
    t_budget, t_start = 240.0, time.time()                   # the whole run ends within a few minutes
    for step in range(total):
        fA = int(frags[step])
        nb = props.neighbours(fA, k_nb)
        if not nb:
            continue
        if step < args.warmup and time.time() - t_start > 30.0:
            continue                                          # bounded warm-up
        _, _, dt = port.step(fA, nb, with_full=True)
        if step >= args.warmup:
            vals.append(N_TMP * len(nb) / dt); t_all.append(dt); n_ev += N_TMP * len(nb)
        if time.time() - t_start > t_budget and vals:
            break
    port.close()
    v = float(n_ev / sum(t_all)) if t_all else 0.0
    n_contigs = len(np.unique(inp.S_o_A_frags["id_c"]))
    E = int(port.lv.rows.size)
    line = {"impl": "reference", "metric": "mcmc_move_loglik_evals_per_s", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(t_all)) * 1e3 if t_all else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 expected contacts, f64 log-likelihood accumulation",
            "data": "synthetic",
            "config": make_config(name, inp.n_frags, inp.init_n_sub_frags, E, k_nb, n_contigs, max(1, args.chains_per_gpu),
                                  int(os.environ.get("WORLD_SIZE", "1")), args.exchange_every),
            "steps_measured": len(vals),
            "cpu_baseline": {"value": v, "unit": "evals/s", "cores": workers, "kind": "port",
                             "sample": "per step: the 13 candidate deltas of each of the %d proposals AND the full likelihood of the current state "
                                       "(all 23.7 M contacts + band mass), NumPy oracle (sparse formulation) on %d forked workers; the genome "
                                       "stays the initial one (no commit); %d of %d steps measured inside the 240 s budget"
                                       % (k_nb, workers, len(vals), args.steps)},
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def profile_kernels(g, _lib):
    import ctypes as C
    kern = {}
    for kname, kid in _lib.KERNELS.items():
        tot, cnt = C.c_double(), C.c_longlong()
        _lib.check(g.lib.graal_profile_read(g.ctx, kid, C.byref(tot), C.byref(cnt), 1))
        kern[kname] = {"ms_total": tot.value, "launches": cnt.value, "ms_avg": tot.value / cnt.value if cnt.value else None}
    cnts = (C.c_longlong * 4)()
    _lib.check(g.lib.graal_profile_counters(g.ctx, cnts, 1))
    return kern, [int(x) for x in cnts]


def roofline_blocks(kern, counters, E, W, peak, peak_src, traffic):
    """`roofline` of the contact-list pass of the full likelihood and `roofline_delta` of the delta contact pass
    (8 B per entry visited, 24 B per row, 16 B per distinct sub-frag, 8 B per candidate)."""
    fc = kern["FULL_CONTACTS"]
    alg = 8 * E + 24 * W + 16 * W + 8
    ach = alg / (fc["ms_avg"] * 1e-3) / 1e9 if fc["ms_avg"] else None
    roof = {"bound": "hbm", "kernel": "k_full_contacts_win (contact-list pass of the per-step full likelihood)",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if ach else None,
            "traffic": traffic.get("full_contacts"), "algorithmic_bytes_per_launch": alg, "avg_launch_ms": fc["ms_avg"],
            "windows_ms": kern.get("FULL_WINDOWS", {}).get("ms_avg"),
            "peak_source": peak_src, "contacts_per_s": E / (fc["ms_avg"] * 1e-3) if fc["ms_avg"] else None}
    dc = kern["DELTA_CONTACTS"]
    e_u, r_u, b_u, n_p = counters
    roof_d = None
    if dc["ms_total"] and n_p:
        alg_d = 8 * e_u + 24 * r_u + 16 * r_u + 8 * N_TMP * n_p
        ach_d = alg_d / (dc["ms_total"] * 1e-3) / 1e9
        roof_d = {"bound": "hbm", "kernel": "delta contact pass (rows of contig(fA) + contig(fB), 13 candidates per launch)",
                  "achieved": ach_d, "peak": peak, "unit": "GB/s", "frac": ach_d / peak,
                  "traffic": traffic.get("delta_contacts"), "algorithmic_bytes_per_launch": alg_d / n_p, "avg_launch_ms": dc["ms_avg"],
                  "proposals": n_p, "mean_entries_in_U": e_u / n_p, "mean_bins_in_U": b_u / n_p,
                  "candidate_contacts_per_s": N_TMP * e_u / (dc["ms_total"] * 1e-3),
                  "limiter": "not HBM: dependent instruction chains per entry (14 evaluations) at 16 warps per SM -- ncu: issue "
                             "26-30 %, 5-10 warps per issue waiting on memory, L2 / L1 throughput 33-35 % (DESIGN.md section 11)"}
    return roof, roof_d


def original_kernels_block(k_nb, n_steps=8):
    """Second baseline: the reference's ORIGINAL kernels (kernels3.cu compiled for sm_100a, baseline/ref_gpu.py) scoring the
    same steps as the device path on BASELINE config C1 level 1 (1,672 bins, W = 5,000: their dense formulation needs
    N < 4,609) -- their own launch sequence per step (full likelihood, per neighbour 22 mutation launches + 17 blocking
    max round trips + index sets + 13 delta launches on 13 streams) against ours (one fused call + one fetch)."""
    import torch
    from baseline import ref_gpu
    from graal_b200.level import yeast_shaped_pyramid, prepare_sampler_inputs
    from graal_b200.sampler import sampler, CUR
    if not ref_gpu.available():
        return {"unavailable": "baseline/_ref/kernels3_sm100a.cubin not built"}
    pyr = yeast_shaped_pyramid(n_levels=2)
    inp = prepare_sampler_inputs(pyr, 1)
    g = sampler.from_inputs(inp, rng=np.random.RandomState(1000))
    p, d_max = model_params(pyr)
    g.set_parameters(p, d_max)
    r = ref_gpu.RefGPU(inp, np.array(list(g.param_simu[0]), dtype=np.float32))
    max_id = int(g.modify_gl_cuda_buffer())
    r.slot_from_host(ref_gpu.CUR, g.slot_to_host(CUR))
    n = int(g.n_new_frags)
    frags = bin_schedule(n, n_steps + 2)
    steps = []
    for fA in frags:
        nb = g.return_neighbours(int(fA), k_nb); nb.sort()
        if nb:
            steps.append((int(fA), nb))
    worst, t_ref, t_dev, n_ev = 0.0, 0.0, 0.0, 0
    for it, (fA, nb) in enumerate(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        full_r, d_r = r.score_step(fA, nb, max_id)
        t1 = time.perf_counter()
        g.score_neighbours(fA, nb)
        _lib_check(g.lib.graal_full_loglik(g.ctx, CUR, None, g._ptr(g.d_out, 0)))
        out = g._fetch()
        t2 = time.perf_counter()
        d_g, full_g = np.array(out[16:16 + N_TMP * len(nb)]), float(out[0])
        if it >= 2:                                               # two warm-up steps
            t_ref += t1 - t0; t_dev += t2 - t1; n_ev += N_TMP * len(nb)
        scale = max(1.0, float(np.abs(d_r).max()))
        worst = max(worst, float(np.abs(d_g - d_r).max()) / scale, abs(full_g - full_r) / abs(full_r))
    launches = r.launches
    g.free_gpu()
    return {"workload": "C1 level 1 (1,672 bins, 5,000 sub-frags, dense 100 MB matrix), %d steps x %d neighbours x 13 candidates, static genome" % (len(steps) - 2, k_nb),
            "original_kernels_evals_per_s": n_ev / t_ref if t_ref else None, "ours_evals_per_s": n_ev / t_dev if t_dev else None,
            "speedup": (t_ref / t_dev) if t_dev else None, "original_ms_per_step": 1e3 * t_ref / max(1, len(steps) - 2),
            "ours_ms_per_step": 1e3 * t_dev / max(1, len(steps) - 2), "original_launches": launches,
            "max_rel_difference": worst, "timing": "host wall clock around each step (both sides end with a device round trip)"}


def _lib_check(rc):
    from graal_b200 import _lib
    _lib.check(rc)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    k_nb = args.neighbours

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from graal_b200 import _lib
    from graal_b200.sampler import sampler, CUR
    from graal_b200 import replica as R

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n_ch = max(1, args.chains_per_gpu)

    def c4_sampler(seed, big=False, share=None, inp4=None):
        from graal_b200.level import synthetic_roofline_level
        if share is None:
            kw = dict(n_bins=1_000_000, n_contigs=64, n_draws=1_200_000_000) if big else {}
            inp4, lists, tables, info = synthetic_roofline_level(device=dev, **kw)
            g4 = sampler.from_inputs(inp4, device=local_rank, rng=np.random.RandomState(seed), device_contact_lists=lists, proposal_tables=tables)
            g4._d_max_kb = info["d_max_kb"]
        else:
            g4 = sampler.from_inputs(inp4, device=local_rank, rng=np.random.RandomState(seed), share_level_with=share)
            g4._d_max_kb = share._d_max_kb
        g4.set_parameters([1.0, 9.6, -1.5, 3.0, 800.0], g4._d_max_kb)
        return inp4, g4

    chains = []
    if args.config in ("c4", "c5"):
        # BASELINE configs C4 (200k bins / ~237 M stored contacts) and C5 (1 M bins / ~1 B contacts, the chains of a GPU share
        # one copy of the level), generated on the GPU (no CPU baseline: the NumPy oracle does not fit these sizes in the time
        # budget; tests/test_gpu_bench_configs.py checks a row sample of C4)
        big = args.config == "c5"
        inp, g = c4_sampler(1000 + rank * n_ch, big=big)
        chains = [g]
        for ch in range(1, n_ch):
            chains.append(c4_sampler(1000 + rank * n_ch + ch, big=big, share=g, inp4=inp)[1])
        name = ("C5: synthetic 1M-bin / ~1B-contact level (64 contigs), %d chains per GPU sharing one copy of the level" % n_ch) if big else \
               "C4: synthetic 200k-bin / ~200M-contact level (24 contigs, offsets ~ s^-1.5 truncated at d_max = 1000 kb, 5% trans), single chain"
        pyr = None
        args.no_cpu_baseline = True
    else:
        pyr, inp, name = build_level(args.config, args.level)
        p, d_max = model_params(pyr)
        for ch in range(n_ch):
            share = chains[0] if chains else None
            g = sampler.from_inputs(inp, device=local_rank, rng=np.random.RandomState(1000 + rank * n_ch + ch), share_level_with=share)
            g.set_parameters(p, d_max)
            chains.append(g)
    g = chains[0]
    for gc in chains:
        gc.incremental_likelihood = bool(args.incremental)
    rex = None
    if world * n_ch > 1:
        rex = R.ReplicaExchange(n_ch, R.temperature_ladder(world * n_ch), exchange_every=args.exchange_every, seed=20141217, device=dev)
        for ch, gc in enumerate(chains):
            R.attach(gc, rex, ch)
    n = int(g.n_new_frags)
    init_state = g.slot_to_host(CUR)
    total = args.warmup + args.steps
    frags = bin_schedule(n, total)

    def barrier():
        if world > 1:
            dist.barrier()
        for gc in chains:
            gc.sync()
        torch.cuda.synchronize(dev)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    windows = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_mark = [0.0]

    # ---------------- pass 1: end to end through the public API (records the schedule) --------------
    schedules = [[] for _ in chains]
    e2e_ms, h2d, d2h = None, 0, 0
    if rex is not None:
        rex.warm_up()                      # NCCL channel set-up for the exchange's all_gather, outside the timed region
    if not args.profile_only:
        for it in range(total):
            if it == args.warmup:
                barrier()
                t_mark[0] = time.time()
                ev0.record(g.stream)
            fA = int(frags[it])
            if n_ch == 1:
                out = g.step_max_likelihood(fA, k_nb)
                schedules[0].append((fA, list(g.id_neighbours) if out[5] >= 0 else [], int(out[6]), int(out[5])))
            else:
                outs = R.step_chains(chains, fA, k_nb)
                for ch, (gc, out) in enumerate(zip(chains, outs)):
                    schedules[ch].append((fA, list(gc.id_neighbours) if out[5] >= 0 else [], int(out[6]), int(out[5])))
            if rex is not None:
                rex.maybe_exchange(it + 1, [gc.likelihood_t for gc in chains])
        for gc in chains[1:]:
            gc.sync()
        ev1.record(g.stream)
        barrier()
        windows.append((t_mark[0], time.time()))
        e2e_ms = ev0.elapsed_time(ev1)
        d2h = g.h_out.numel() * 8 * n_ch     # one fetch of the pinned output block per chain-step: scores, stats, candidate distances
        h2d = 0                              # proposal ids travel as kernel arguments
    else:
        for it in range(total):
            fA = int(frags[it])
            for ch, gc in enumerate(chains):
                nb = gc.return_neighbours(fA, k_nb); nb.sort()
                schedules[ch].append((fA, nb, nb[0] if nb else fA, 6 if nb else -1))

    if world > 1:
        # the resident replay measures the machine, not the luck of a chain: every rank replays RANK 0's recorded schedule
        # (same proposals, same committed moves: identical work per GPU); the end-to-end pass above ran the ranks' own chains
        box = [schedules]
        dist.broadcast_object_list(box, src=0)
        schedules = box[0]
    # ---------------- pass 2: device resident replay (no host round trip inside the timed region) ------
    def replay(profile=False):
        for gc in chains:
            gc.slot_from_host(CUR, init_state)
        if rex is not None:
            rex.reset()
        l0 = sum(gc.gpu_launches for gc in chains)
        for it in range(total):
            if it == args.warmup:
                barrier()
                if profile:
                    for gc in chains:
                        _lib.check(gc.lib.graal_profile_enable(gc.ctx, 1))
                t_mark[0] = time.time()
                l0 = sum(gc.gpu_launches for gc in chains)
                ev0.record(g.stream)
            for gc, sched in zip(chains, schedules):
                fA, nb, fB, op = sched[it]
                gc.step_device(fA, nb, fB, op, full=(not args.incremental) or it % 256 == 0)
            if rex is not None and not profile:
                rex.maybe_exchange_device(it + 1, chains)
        for gc in chains[1:]:
            gc.sync()
        ev1.record(g.stream)
        barrier()
        return ev0.elapsed_time(ev1), sum(gc.gpu_launches for gc in chains) - l0

    ms, launches = replay(False)
    windows.append((t_mark[0], time.time()))
    if args.profile_only:
        print(json.dumps({"profile_only": True, "ms_per_step": ms / max(1, args.steps), "launches": launches}))
        return
    # ---------------- pass 3: same replay with the library's per-kernel event timers ---------------------
    replay(True)
    kern, counters = profile_kernels(g, _lib)
    for gc in chains:
        _lib.check(gc.lib.graal_profile_enable(gc.ctx, 0))

    # ---------------- aggregate over ranks -------------------------------------------------------------
    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    n_eval = sum(N_TMP * len(nb) for sched in schedules for (_, nb, _, op) in sched[args.warmup:] if op >= 0)
    tot_eval = reduce_ranks(float(n_eval), dist.ReduceOp.SUM)
    ms_max, ms_min = reduce_ranks(ms, dist.ReduceOp.MAX), reduce_ranks(ms, dist.ReduceOp.MIN)
    e2e_max, e2e_min = reduce_ranks(e2e_ms, dist.ReduceOp.MAX), reduce_ranks(e2e_ms, dist.ReduceOp.MIN)
    value = tot_eval / (ms_max * 1e-3)
    e2e_value = tot_eval / (e2e_max * 1e-3)
    exch = rex.stats() if rex is not None else None

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    clk = clocks.stop(windows)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    traffic_all = {}
    try:
        traffic_all = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    E, W = g.n_contacts, int(g.init_n_sub_frags)
    cfg_key = args.config if args.config != "c2" or args.level == 1 else "c2_l%d" % args.level
    roof, roof_d = roofline_blocks(kern, counters, E, W, peak, peak_src, traffic_all.get(cfg_key, {}))
    line = {
        "metric": "mcmc_move_loglik_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 expected contacts, f64 log-likelihood accumulation",
        "data": "synthetic",
        "config": make_config(name, g.n_frags, W, E, k_nb, len(np.unique(init_state["id_c"])), n_ch, world, args.exchange_every, args.incremental),
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_max / args.steps},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof,
        "roofline_delta": roof_d,
        "kernels": kern,
        "contacts_scored_per_s": {"streamed_full_pass": roof["contacts_per_s"],
                                  "candidate_contacts_delta_pass": roof_d["candidate_contacts_per_s"] if roof_d else None},
        "ranks": {"ms_min": ms_min / args.steps, "ms_max": ms_max / args.steps, "e2e_ms_min": e2e_min / args.steps,
                  "e2e_ms_max": e2e_max / args.steps, "exchange": exch},
    }
    rc = 0
    if world == 1 and n_ch == 1 and not args.no_cpu_baseline:
        # the first timed step with neighbours, on the INITIAL genome, by the oracle on the host cores and by the device
        it0 = next((i for i in range(args.warmup, total) if schedules[0][i][1]), None)
        if it0 is not None:
            fA, nb, _, _ = schedules[0][it0]
            workers = max(1, min(32, os.cpu_count() or 1))
            port = CpuStep(inp, pyr, workers)
            port.step(fA, nb[:1], with_full=False)                     # workers warm (imports, first-touch of the level)
            d_cpu, full_cpu, dt = port.step(fA, nb, with_full=True)
            port.close()
            g.slot_from_host(CUR, init_state)
            g.modify_gl_cuda_buffer()
            full_gpu = float(g.eval_likelihood())
            g.score_neighbours(fA, nb)
            d_gpu = [float(x) for x in g._fetch()[16:16 + N_TMP * len(nb)]]
            line["parity"] = parity_block(d_cpu, full_cpu, d_gpu, full_gpu)
            line["cpu_baseline"] = {"value": N_TMP * len(nb) / dt, "unit": "evals/s", "cores": workers, "kind": "port",
                                    "host_cores_visible": os.cpu_count(),
                                    "sample": "one whole step (the first timed one, on the initial genome): the %d candidate deltas of its %d "
                                              "proposals and the full likelihood of the state (all %d contacts + band mass), NumPy oracle "
                                              "(sparse formulation) on %d forked workers, %.1f s" % (N_TMP * len(nb), len(nb), E, workers, dt)}
            if not line["parity"]["ok"]:
                rc = 1
    if world == 1 and n_ch == 1 and not args.no_original and args.config == "c2":
        try:
            line["original_kernels"] = original_kernels_block(k_nb)
        except Exception as e:                                    # the cubin is an optional artefact (built where /root/reference exists)
            line["original_kernels"] = {"unavailable": repr(e)[:200]}
    if world == 1 and not args.no_c4 and args.config == "c2" and n_ch == 1:
        # BASELINE config C4 in the same run: the HBM roofline configuration (the contact list is 1.9 GB)
        for gc in chains:
            gc.free_gpu()
        del chains, g
        torch.cuda.empty_cache()
        inp4, g4 = c4_sampler(1000)
        n4 = int(g4.n_new_frags)
        fr4 = bin_schedule(n4, 13)
        for it in range(3):
            g4.step_max_likelihood(int(fr4[it]), k_nb)
        _lib.check(g4.lib.graal_profile_enable(g4.ctx, 1))
        t0 = time.time()
        for it in range(3, 13):
            g4.step_max_likelihood(int(fr4[it]), k_nb)
        g4.sync()
        step_ms = (time.time() - t0) / 10 * 1e3
        kern4, cnt4 = profile_kernels(g4, _lib)
        _lib.check(g4.lib.graal_profile_enable(g4.ctx, 0))
        r4, r4d = roofline_blocks(kern4, cnt4, g4.n_contacts, int(g4.init_n_sub_frags), peak, peak_src, traffic_all.get("c4", {}))
        r4["workload"] = "C4: synthetic 200k-bin / 237M-contact level (24 contigs), 10 steps of the same chain with the per-kernel timers on"
        r4["ms_per_step_with_timers"] = step_ms
        line["roofline_c4"] = r4
        line["roofline_delta_c4"] = r4d
        g4.free_gpu()
    print(json.dumps(line))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    if rc:
        sys.exit(rc)


if __name__ == "__main__":
    main()
