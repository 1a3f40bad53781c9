"""Replica chains / replica-exchange temperatures across GPUs (NEW relative to the reference).

The reference is single-GPU; its only hook is ``sampler.temperature()`` which returns the constant 1.0
(cuda_lib_gl.py:2590-2603) and is applied to the candidate weights (:1932-1934) and to the
nuisance-parameter acceptance ratio (:2091-2092).  The MCMC path does not shard inside one chain
(each step touches one or two contigs), so the only axis that shards is the CHAIN: one process per
GPU (torchrun), every rank holds a full read-only copy of the level and its own chain(s).

Exchange protocol (the only collective on the path): every ``exchange_every`` steps each rank
contributes (log-likelihood, temperature index) per local chain to one ``all_gather`` (16 B per
chain, NCCL over NVLink on GPUs, gloo on CPU); every rank then runs the SAME deterministic even/odd
neighbour-swap decision from a shared seed and swaps temperature LABELS, never states.

The exchange never blocks a chain: ``post`` copies each chain's log-likelihood out of its device output block with a
tiny device copy on the chain's stream (or takes the host values the sampler already holds), runs ONE
``all_gather_into_tensor`` on a side stream and an asynchronous copy of the gathered table into pinned memory; the
decision is taken from that table at the NEXT step boundary (``consume``), identically on every rank.

Note on the acceptance rule: the reference tempers its candidate draw as (normalised shifted linear score)**(1/T)
(cuda_lib_gl.py:1932-1934; restated in sampler._sample), not as exp(L / T), so a tempered chain does not target
pi**(1/T) and the Boltzmann swap rule below is a HEURISTIC coupling of the chains, not an exact replica-exchange
scheme with detailed balance.  The T = 1 chain follows the reference's own transition rule between swaps.
"""
import time

import numpy as np


def temperature_ladder(n, ratio=1.25):
    """T_k = ratio**k, k = 0..n-1."""
    return np.power(float(ratio), np.arange(n, dtype=np.float64))


def swap_decisions(logliks, temp_index, temperatures, round_id, seed):
    """Deterministic replica-exchange round.  ``logliks[c]`` / ``temp_index[c]`` for every GLOBAL chain c.
    Pairs (k, k+1) with k = round_id mod 2, +2, ... are proposed; a swap between the chains holding
    T_k and T_{k+1} is accepted with probability min(1, exp((1/T_k - 1/T_{k+1}) * (L_{k+1} - L_k))).
    Returns (new temp_index, list of (k, accepted, chain_k, chain_k1))."""
    temp_index = np.array(temp_index, dtype=np.int64).copy()
    logliks = np.asarray(logliks, dtype=np.float64)
    n_t = len(temperatures)
    holder = np.full(n_t, -1, dtype=np.int64)
    holder[temp_index] = np.arange(len(temp_index))
    rng = np.random.RandomState((int(seed) + 7919 * int(round_id)) % (2 ** 31 - 1))
    log = []
    for k in range(int(round_id) % 2, n_t - 1, 2):
        a, b = holder[k], holder[k + 1]
        u = rng.random_sample()
        if a < 0 or b < 0:
            continue
        la, lb = logliks[a], logliks[b]
        acc = False
        if np.isfinite(la) and np.isfinite(lb):
            x = (1.0 / temperatures[k] - 1.0 / temperatures[k + 1]) * (lb - la)
            acc = bool(x >= 0 or u < np.exp(x))
        if acc:
            temp_index[a], temp_index[b] = k + 1, k
            holder[k], holder[k + 1] = b, a
        log.append((k, acc, int(a), int(b)))
    return temp_index, log


class ReplicaExchange:
    """Temperature bookkeeping of the chains of this rank + the periodic label exchange."""

    def __init__(self, n_local_chains, temperatures, exchange_every=100, seed=0, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if self.distributed else 0
        self.world = dist.get_world_size() if self.distributed else 1
        self.n_local = int(n_local_chains)
        self.n_chains = self.n_local * self.world
        self.temperatures = np.asarray(temperatures, dtype=np.float64)
        if len(self.temperatures) != self.n_chains:
            raise ValueError("need one temperature per chain (%d != %d)" % (len(self.temperatures), self.n_chains))
        self.exchange_every = int(exchange_every)
        self.seed = int(seed)
        self.device = device if device is not None else torch.device("cpu")
        self.temp_index = np.arange(self.n_chains, dtype=np.int64)      # chain c starts at T_c
        self.round_id = 0
        self.history = []
        self.n_collectives = 0

    def global_chain(self, local):
        return self.rank * self.n_local + local

    def temperature(self, local=0):
        return float(self.temperatures[self.temp_index[self.global_chain(local)]])

    def exchange(self, local_logliks):
        """all_gather (loglik, temperature index) of every chain, then the shared swap decision."""
        torch = self.torch
        mine = np.zeros((self.n_local, 2), dtype=np.float64)
        mine[:, 0] = np.asarray(local_logliks, dtype=np.float64)
        mine[:, 1] = self.temp_index[self.rank * self.n_local:(self.rank + 1) * self.n_local]
        if self.distributed:
            t = torch.from_numpy(mine).to(self.device)
            out = [torch.empty_like(t) for _ in range(self.world)]
            self.dist.all_gather(out, t)
            self.n_collectives += 1
            allv = torch.stack(out).reshape(-1, 2).cpu().numpy()
        else:
            allv = mine
        gathered_index = allv[:, 1].astype(np.int64)
        if not np.array_equal(gathered_index, self.temp_index):
            raise RuntimeError("replica temperature labels diverged between ranks")
        self.temp_index, log = swap_decisions(allv[:, 0], gathered_index, self.temperatures, self.round_id, self.seed)
        self.history.append(log)
        self.round_id += 1
        return log

    # ---- non-blocking exchange ------------------------------------------------------------------------------------
    def _buffers(self):
        if getattr(self, "_send", None) is not None:
            return
        torch = self.torch
        self._cuda = self.device.type == "cuda"
        self._send = torch.zeros((self.n_local, 2), dtype=torch.float64, device=self.device)
        self._recv = torch.zeros((self.n_chains, 2), dtype=torch.float64, device=self.device)
        self._h_send = torch.zeros((self.n_local, 2), dtype=torch.float64)
        self._host = torch.zeros((self.n_chains, 2), dtype=torch.float64)
        if self._cuda:
            self._h_send, self._host = self._h_send.pin_memory(), self._host.pin_memory()
            self._side = torch.cuda.Stream(device=self.device)
            self._done = torch.cuda.Event()
        self._pending = False
        self.t_post = self.t_consume = 0.0
        self.n_posted = 0

    def post(self, local_logliks=None, chains=None):
        """Start an exchange.  ``chains``: samplers whose device output block holds the log-likelihood of the current state
        in d_out[0] (device-resident replay); else ``local_logliks`` are host values.  Returns at once."""
        self._buffers()
        torch, t0 = self.torch, time.perf_counter()
        lo = self.rank * self.n_local
        self._h_send[:, 1] = torch.from_numpy(self.temp_index[lo:lo + self.n_local].astype(np.float64))
        if chains is None:
            self._h_send[:, 0] = torch.from_numpy(np.asarray(local_logliks, dtype=np.float64))
        if self._cuda:
            if chains is not None:
                for ch, g in enumerate(chains):
                    with torch.cuda.stream(g.stream):
                        self._send[ch, 0:1].copy_(g.d_out[0:1])              # device copy on the chain's stream
                    self._side.wait_stream(g.stream)
            with torch.cuda.stream(self._side):
                if chains is not None:
                    self._send[:, 1].copy_(self._h_send[:, 1], non_blocking=True)
                else:
                    self._send.copy_(self._h_send, non_blocking=True)
                if self.distributed:
                    self.dist.all_gather_into_tensor(self._recv, self._send)
                    self.n_collectives += 1
                else:
                    self._recv.copy_(self._send)
                self._host.copy_(self._recv, non_blocking=True)
                self._done.record(self._side)
        else:
            if chains is not None:
                self._h_send[:, 0] = torch.stack([g.d_out[0] for g in chains]).cpu()
            if self.distributed:
                out = [torch.empty_like(self._h_send) for _ in range(self.world)]
                self.dist.all_gather(out, self._h_send)
                self.n_collectives += 1
                self._host.copy_(torch.stack(out).reshape(-1, 2))
            else:
                self._host.copy_(self._h_send)
        self._pending = True
        self.n_posted += 1
        self.t_post += time.perf_counter() - t0

    def consume(self):
        """Apply the exchange posted earlier (no-op if none): the shared swap decision from the gathered table."""
        if not getattr(self, "_pending", False):
            return None
        t0 = time.perf_counter()
        if self._cuda:
            self._done.synchronize()
        allv = self._host.numpy().copy()
        self._pending = False
        gathered_index = allv[:, 1].astype(np.int64)
        if not np.array_equal(gathered_index, self.temp_index):
            raise RuntimeError("replica temperature labels diverged between ranks")
        self.temp_index, log = swap_decisions(allv[:, 0], gathered_index, self.temperatures, self.round_id, self.seed)
        self.history.append(log)
        self.round_id += 1
        self.t_consume += time.perf_counter() - t0
        return log

    def reset(self):
        self.consume()
        self.temp_index = np.arange(self.n_chains, dtype=np.int64)
        self.round_id = 0

    def stats(self):
        self._buffers()
        return {"posted": self.n_posted, "post_ms_total": self.t_post * 1e3, "consume_ms_total": self.t_consume * 1e3,
                "every": self.exchange_every, "blocking": False}

    def warm_up(self):
        """One exchange of the real shape that changes nothing: pays the lazy communicator / channel set-up of the
        backend outside any timed or latency-critical region."""
        self._buffers()
        keep = (self.temp_index.copy(), self.round_id, len(self.history), self.n_posted, self.t_post, self.t_consume)
        self.post(local_logliks=np.full(self.n_local, np.nan))      # non-finite likelihoods never swap
        self.consume()
        self.temp_index, self.round_id = keep[0], keep[1]
        del self.history[keep[2]:]
        self.n_posted, self.t_post, self.t_consume = keep[3], keep[4], keep[5]

    def maybe_exchange(self, step, local_logliks):
        """Host-side likelihoods (the values step_max_likelihood returned): the exchange posted at a boundary is applied at
        the next step boundary; nothing blocks."""
        log = self.consume()
        if self.exchange_every > 0 and step > 0 and step % self.exchange_every == 0:
            self.post(local_logliks=local_logliks)
        return log

    def maybe_exchange_device(self, step, chains):
        """Device-resident likelihoods (d_out[0] of every chain): no host round trip."""
        log = self.consume()
        if self.exchange_every > 0 and step > 0 and step % self.exchange_every == 0:
            self.post(chains=chains)
        return log


def step_chains(chains, id_fA, delta, t=0, n_step=1):
    """One step of several chains that share a GPU: every chain's work is enqueued first (its kernels run on its own
    streams and overlap with the other chains'), then each chain's round trip / draw / commit is finished in turn."""
    # (chains that draw from ONE generator keep the host draw: the device draw hands the next uniform over before it knows
    # whether it is consumed, which only one pending step per generator can do)
    own_rng = len({id(g.rng) for g in chains}) == len(chains)
    for g in chains:
        g.step_begin(id_fA, delta, t, n_step, device_draw=own_rng)
    return [g.step_end(t, n_step) for g in chains]


def attach(sampler_obj, rex, local=0):
    """Make ``sampler.temperature()`` follow the replica's current temperature."""
    sampler_obj.temperature = lambda t=0, n_step=1: rex.temperature(local)
    return sampler_obj
