#!/usr/bin/env python
"""bench.py -- GRAAL MCMC move-scoring hot path on B200.

Metric (BASELINE.json): MCMC move log-likelihood evaluations / s.  One "step" is one
``step_max_likelihood`` of the reference sampler (cuda_lib_gl.py:1793-1980) for one bin: relabel,
full log-likelihood, ``n_neighbours`` proposals x 13 candidate structures built and scored, one
candidate committed -> 13 * n_neighbours move evaluations.

Workload at N = 1: BASELINE config C2 -- synthetic T. reesei-shaped pyramid (77 contigs, 33 Mb,
100,000 level-0 fragments, factor 3), single chain, run at level 1 (33k bins, 100k sub-frags,
~30 M stored contact entries = 240 MB of contact lists, larger than the 126 MB L2, re-streamed every
step).  N > 1: one independent replica chain per GPU (weak scaling) with a replica-exchange
all_gather of (loglik, temperature index) every --exchange-every steps.

    value     device-resident throughput: the recorded proposal schedule replayed with no host round trip
    e2e       the same schedule through the public sampler API (host RNG draws, D2H of scores / state)
    roofline  the contact-list pass of the full likelihood (k_full_contacts), CUDA events inside the library
    cpu_baseline / --impl reference: the NumPy oracle (sparse formulation) on the host cores
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
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c4", "tiny"])
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--neighbours", type=int, default=3)
    ap.add_argument("--exchange-every", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200",
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


_POOL_CTX = {}


def _pool_delta(args):
    from oracle import sparse as S
    cand, bins_u = args
    c = _POOL_CTX
    return S.sparse_delta(cand, c["cur"], c["lv"], c["par"], bins_u)


def _pool_noop(j):
    return j


class CpuPort:
    """The oracle (NumPy port of the reference semantics, sparse formulation) scoring proposals of a level on
    the host: the level structures are built once; per proposal the 13 candidates are built and scored, by
    `workers` forked processes when workers > 1 (the candidates are independent, as on the reference's 13
    streams; the processes inherit the read-only level and are started before any timed region)."""

    def __init__(self, inp, pyr, workers=1):
        from oracle import mutations as M, sparse as S, likelihood as L
        self.M, self.S = M, S
        p, d_max = model_params(pyr)
        self.par = L.make_params(p[0], p[1], p[2], p[3], p[4], d_max, inp.mean_value_trans)
        r, c, v = inp.sub_coo
        t0 = time.time()
        self.lv = S.SparseLevel(inp.n_frags, inp.np_sub_frags_id, inp.np_sub_frags_len_bp, inp.np_sub_frags_accu,
                                inp.mean_squared_frags_per_bin, r, c, v)
        self.cur = {k: np.array(inp.S_o_A_frags[k], dtype=np.int32) for k in M.FIELDS}
        self.cur["ori"][:] = 1
        self.max_id = M.relabel_contigs(self.cur)
        self.t_setup = time.time() - t0
        self.ws = M.Workspace(inp.n_new_frags)
        self.workers, self.pool = int(workers), None
        if self.workers > 1:
            import multiprocessing as mp
            _POOL_CTX.update(cur=self.cur, lv=self.lv, par=self.par)
            self.pool = mp.get_context("fork").Pool(self.workers)
            self.pool.map(_pool_noop, range(self.workers))

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool = None
        _POOL_CTX.clear()

    def score(self, fA, fB, max_seconds=45.0):
        M, S, cur, lv = self.M, self.S, self.cur, self.lv
        M.perform_modifications(self.ws, cur, fA, fB, self.max_id)
        in_u = (cur["id_c"] == cur["id_c"][fA]) | (cur["id_c"] == cur["id_c"][fB])
        bins_u = np.nonzero(in_u)[0]
        n_eval = 0
        t0 = time.time()
        if self.pool is not None:
            self.pool.map(_pool_delta, [(self.ws.collector[j], bins_u) for j in range(N_TMP)], chunksize=1)
            n_eval = N_TMP
        else:
            for j in range(N_TMP):
                S.sparse_delta(self.ws.collector[j], cur, lv, self.par, bins_u)
                n_eval += 1
                if time.time() - t0 > max_seconds:
                    break
        return dict(n_eval=n_eval, t_delta=time.time() - t0, contig_bins=int(bins_u.size))

    def full_sample(self, fA):
        """The per-step full likelihood, timed on a row sample of the contact list and of the band pairs."""
        S, lv = self.S, self.lv
        t0 = time.time()
        geo = S.Geo(self.cur, lv)
        sel = np.nonzero(lv.rows % 20 == 0)[0]
        S.contact_terms(geo, lv, self.par, sel)
        sub_sample = np.nonzero(geo.id_c == geo.id_c[lv.sub_id[fA, 0]])[0]
        S.band_mass(geo, lv, self.par, sub_sample)
        return dict(t_full_sample=time.time() - t0, sample_rows=int(sel.size), sample_subs=int(sub_sample.size))


def cpu_port_sample(inp, pyr, fA, fB, max_seconds=45.0):
    """One proposal on one thread + the full-likelihood sample (the `cpu_baseline` of the default run)."""
    port = CpuPort(inp, pyr, workers=1)
    r = port.score(fA, fB, max_seconds)
    r.update(port.full_sample(fA))
    r.update(t_setup=port.t_setup, n_contacts=int(port.lv.rows.size), W=int(port.lv.W))
    return r


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    k_nb = args.neighbours

    if args.impl == "reference":
        # the reference's own CPU implementation of the path = the NumPy oracle port (the reference is
        # Python 2 + PyCUDA and cannot run here; its kernels need a GPU).  Rank 0 only.
        if rank != 0:
            return
        pyr, inp, name = build_level(args.config, args.level)
        rng = np.random.RandomState(1000)
        s = inp.S_o_A_frags
        vals, t_all = [], []
        workers = max(1, min(N_TMP, os.cpu_count() or 1))        # all the host threads the 13 candidates can use
        port = CpuPort(inp, pyr, workers=workers)
        t_budget, t_start = 240.0, time.time()                   # the whole run ends within a few minutes
        for step in range(args.warmup + args.steps):
            fA = int(rng.randint(inp.n_new_frags))
            fB = int(s["next"][fA]) if s["next"][fA] >= 0 else int(s["prev"][fA])
            r = port.score(fA, fB, max_seconds=20.0)
            if step >= args.warmup:
                vals.append(r["n_eval"] / r["t_delta"]); t_all.append(r["t_delta"])
            if time.time() - t_start > t_budget and vals:
                break
        port.close()
        v = float(np.mean(vals)) if vals else 0.0
        line = {"impl": "reference", "metric": "mcmc_move_loglik_evals_per_s", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(t_all)) * 1e3 if t_all else None,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 expected / f64 log-likelihood",
                "data": "synthetic", "config": {"workload": name, "neighbours": k_nb, "steps_measured": len(vals)},
                "cpu_baseline": {"value": v, "unit": "evals/s", "cores": workers, "kind": "port",
                                 "sample": "per step: the 13 candidate deltas of ONE proposal of the same level, NumPy oracle (sparse formulation), "
                                           "one forked worker per candidate up to the host core count; the per-step full likelihood is not included"},
                "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from graal_b200 import _lib
    from graal_b200.sampler import sampler, CUR, FRAG_FIELDS
    from graal_b200 import replica as R

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    if args.config == "c4":
        # BASELINE config C4: 200k bins / ~200 M stored contacts, generated on the GPU (no CPU baseline: the
        # NumPy oracle does not fit this size in the time budget)
        from graal_b200.level import synthetic_roofline_level
        inp, lists, tables, info = synthetic_roofline_level(device=dev)
        name = "C4: synthetic 200k-bin / ~200M-contact level (24 contigs, offsets ~ s^-1.5 truncated at d_max = 1000 kb, 5% trans), single chain"
        pyr = None
        g = sampler.from_inputs(inp, device=local_rank, rng=np.random.RandomState(1000 + rank),
                                device_contact_lists=lists, proposal_tables=tables)
        g.set_parameters([1.0, 9.6, -1.5, 3.0, 800.0], info["d_max_kb"])
        args.no_cpu_baseline = True
    else:
        pyr, inp, name = build_level(args.config, args.level)
        g = sampler.from_inputs(inp, device=local_rank, rng=np.random.RandomState(1000 + rank))
        p, d_max = model_params(pyr)
        g.set_parameters(p, d_max)
    g.incremental_likelihood = bool(args.incremental)
    rex = None
    if world > 1:
        rex = R.ReplicaExchange(1, R.temperature_ladder(world), exchange_every=args.exchange_every, seed=20141217, device=dev)
        R.attach(g, rex, 0)
    n = int(g.n_new_frags)
    init_state = g.slot_to_host(CUR)
    total = args.warmup + args.steps
    sched_rng = np.random.RandomState(4242)          # the same bins on every rank: replicas do statistically identical work
    frags = sched_rng.permutation(n)[:total]
    if frags.size < total:
        frags = np.concatenate([frags, sched_rng.randint(0, n, size=total - frags.size)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    windows = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---------------- pass 1: end to end through the public API (records the schedule) --------------
    schedule = []
    e2e_ms, h2d, d2h = None, 0, 0
    if rex is not None:
        rex.warm_up()                      # NCCL channel set-up for the exchange's all_gather, outside the timed region
    if not args.profile_only:
        for it in range(total):
            if it == args.warmup:
                barrier()
                t_w0 = time.time()
                ev0.record(g.stream)
            fA = int(frags[it])
            out = g.step_max_likelihood(fA, k_nb)
            schedule.append((fA, list(g.id_neighbours), int(out[6]), int(out[5])))
            if rex is not None:
                rex.maybe_exchange(it + 1, [g.likelihood_t])
        ev1.record(g.stream)
        barrier()
        windows.append((t_w0, time.time()))
        e2e_ms = ev0.elapsed_time(ev1)
        # per step: D2H of the pinned output block, twice; H2D: none
        # (proposal ids travel as kernel arguments)
        d2h = g.h_out.numel() * 8            # one fetch of the pinned output block per step: scores, stats, candidate distances
        h2d = 0
    else:
        for it in range(total):
            fA = int(frags[it])
            nb = g.return_neighbours(fA, k_nb); nb.sort()
            schedule.append((fA, nb, nb[0] if nb else fA, 6))

    # ---------------- pass 2: device resident replay (no host round trip inside the timed region) ------
    def replay(profile=False):
        g.slot_from_host(CUR, init_state)
        if rex is not None:
            rex.temp_index = np.arange(world, dtype=np.int64); rex.round_id = 0
        l0 = g.gpu_launches
        for it, (fA, nb, fB, op) in enumerate(schedule):
            if it == args.warmup:
                barrier()
                if profile:
                    _lib.check(g.lib.graal_profile_enable(g.ctx, 1))
                nonlocal_t[0] = time.time()
                l0 = g.gpu_launches
                ev0.record(g.stream)
            g.step_device(fA, nb, fB, op, full=(not args.incremental) or it % 256 == 0)
            if rex is not None and not profile and (it + 1) % args.exchange_every == 0:
                like = g._fetch()[0]
                rex.maybe_exchange(it + 1, [like])
        ev1.record(g.stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        return ms, g.gpu_launches - l0

    nonlocal_t = [0.0]
    ms, launches = replay(False)
    windows.append((nonlocal_t[0], time.time()))
    if args.profile_only:
        print(json.dumps({"profile_only": True, "ms_per_step": ms / max(1, args.steps), "launches": launches}))
        return
    # ---------------- pass 3: same replay with the library's per-kernel event timers ---------------------
    prof_ms, _ = replay(True)
    kern = {}
    for kname, kid in _lib.KERNELS.items():
        import ctypes as C
        tot, cnt = C.c_double(), C.c_longlong()
        _lib.check(g.lib.graal_profile_read(g.ctx, kid, C.byref(tot), C.byref(cnt), 1))
        kern[kname] = {"ms_total": tot.value, "launches": cnt.value, "ms_avg": tot.value / cnt.value if cnt.value else None}
    _lib.check(g.lib.graal_profile_enable(g.ctx, 0))

    # ---------------- aggregate over ranks -------------------------------------------------------------
    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_eval = sum(N_TMP * len(nb) for (_, nb, _, _) in schedule[args.warmup:])
    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
    tot_eval = sum_over_ranks(float(n_eval))
    ms_max = max_over_ranks(ms)
    e2e_max = max_over_ranks(e2e_ms)
    value = tot_eval / (ms_max * 1e-3)
    e2e_value = tot_eval / (e2e_max * 1e-3)

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
    E, W = g.n_contacts, int(g.init_n_sub_frags)
    alg_bytes = 8 * E + 24 * W + 16 * W + 8
    fc = kern["FULL_CONTACTS"]
    achieved = alg_bytes / (fc["ms_avg"] * 1e-3) / 1e9 if fc["ms_avg"] else None
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        try:
            traffic = json.load(open(tr_path)).get("k_full_contacts_dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": "mcmc_move_loglik_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 expected contacts, f64 log-likelihood accumulation",
        "data": "synthetic",
        "config": {"workload": name, "bins": int(g.n_frags), "sub_frags": W, "contact_entries": E,
                   "contact_list_MB": round(8 * E / 1e6, 1), "neighbours_per_step": k_nb, "candidates_per_neighbour": N_TMP,
                   "state": "assembled genome (%d contigs)" % len(np.unique(init_state["id_c"])),
                   "l2": "inputs larger than L2: the %.0f MB contact list is re-streamed every step" % (8 * E / 1e6),
                   "full_likelihood": ("incremental: carried from the committed candidate, full pass every 256 steps (NOT the reference's schedule)"
                                       if args.incremental else "recomputed every step, as the reference does"),
                   "parallelism": "1 replica chain per GPU" + (", replica-exchange all_gather every %d steps" % args.exchange_every if world > 1 else "")},
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_max / args.steps},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "k_full_contacts (contact-list pass of the per-step full likelihood)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                     "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": fc["ms_avg"],
                     "peak_source": peak_src, "contacts_per_s": E / (fc["ms_avg"] * 1e-3) if fc["ms_avg"] else None},
        "kernels": kern,
        "contacts_scored_per_s": {"streamed_full_pass": E / (fc["ms_avg"] * 1e-3) if fc["ms_avg"] else None},
    }
    if world == 1 and not args.no_cpu_baseline:
        fA, nb, _, _ = schedule[args.warmup]
        r = cpu_port_sample(inp, pyr, fA, nb[0] if nb else fA, max_seconds=30.0)
        line["cpu_baseline"] = {"value": r["n_eval"] / r["t_delta"], "unit": "evals/s", "cores": 1, "kind": "port",
                                "host_cores_visible": os.cpu_count(),
                                "sample": "the %d candidate deltas of the first timed proposal (contigs of %d bins), NumPy oracle (sparse "
                                          "formulation), %.1f s; the per-step full likelihood is NOT included (a %d-row / %d-sub-frag sample "
                                          "of it took %.1f s)" % (r["n_eval"], r["contig_bins"], r["t_delta"], r["sample_rows"],
                                                                    r["sample_subs"], r["t_full_sample"])}
    print(json.dumps(line))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
