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
"""
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

    def warm_up(self):
        """One all_gather of the exchange's shape that changes nothing: pays the lazy communicator / channel
        set-up of the backend outside any timed or latency-critical region."""
        if self.distributed:
            t = self.torch.zeros((self.n_local, 2), dtype=self.torch.float64, device=self.device)
            out = [self.torch.empty_like(t) for _ in range(self.world)]
            self.dist.all_gather(out, t)
            self.torch.stack(out).cpu()

    def maybe_exchange(self, step, local_logliks):
        if self.exchange_every > 0 and step > 0 and step % self.exchange_every == 0:
            return self.exchange(local_logliks)
        return None


def attach(sampler_obj, rex, local=0):
    """Make ``sampler.temperature()`` follow the replica's current temperature."""
    sampler_obj.temperature = lambda t=0, n_step=1: rex.temperature(local)
    return sampler_obj
