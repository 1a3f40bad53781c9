"""The reference's original kernels (baseline/_ref/kernels3_sm100a.cubin) on the B200: loaded with the CUDA driver API
(cuda-python), launched with the reference's own grid / block shapes and in the reference's own sequences
(cuda_lib_gl.py: evaluate_likelihood :546-569, new_perform_modificationS :841-954 / 1045-1048, stream_likelihood
:2441-2533).  torch owns the device buffers; the kernels see raw pointers and the 14-pointer `frag` struct
(kernels3.cu:9-24), exactly what PyCUDA / GPUStruct handed them.  Dense level matrix: N < 4,609 only.

Used (a) as a cross-check oracle on the GPU box (tests/test_gpu_original_kernels.py: integer state bit-exact, likelihood
to the float32-libm tolerance) and (b) as the second baseline of bench.py (their step time next to ours)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CUBIN = os.path.join(HERE, "_ref", "kernels3_sm100a.cubin")
FIELDS = ("pos", "id_c", "start_bp", "len_bp", "circ", "id", "prev", "next", "l_cont", "l_cont_bp", "ori", "rep", "activ", "id_d")
N_TMP = 13
# slots: 0 current, 1..13 collectors, 14 pop, 15 trans1, 16 trans2
CUR, CAND0, POP, TRANS1, TRANS2, N_SLOTS = 0, 1, 14, 15, 16, 17


def available():
    return os.path.exists(CUBIN)


def _check(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError("CUDA driver error %s" % str(err))
    return res[1] if len(res) == 2 else res[1:]


class RefGPU:
    def __init__(self, inp, params8, device=0, size_block=128):
        """``inp``: graal_b200.level.SamplerInputs (no repeats, no blacklist: the comparison levels); ``params8``: the 8 float32
        of param_simu.  ``size_block``: the block size main_gl.py passes to the mutation kernels (EM: 128)."""
        import torch
        from cuda.bindings import driver
        self.torch, self.drv = torch, driver
        self.dev = torch.device("cuda", device)
        torch.cuda.set_device(self.dev)
        torch.zeros(1, device=self.dev)                      # primary context current
        data = open(CUBIN, "rb").read()
        self.module = _check(driver.cuModuleLoadData(data))
        self.fn = {}
        for name in ("flip_frag", "swap_activity_frag", "pop_out_frag", "pop_in_frag_1", "pop_in_frag_2", "pop_in_frag_3", "pop_in_frag_4",
                     "split_contig", "paste_contigs", "simple_copy", "copy_struct", "fill_sub_index_fA", "fill_sub_index_fB",
                     "evaluate_likelihood", "sub_compute_likelihood"):
            self.fn[name] = _check(driver.cuModuleGetFunction(self.module, name.encode()))
        self.size_block = int(size_block)
        n = self.n = int(inp.n_new_frags)
        self.N = int(inp.n_frags)
        self.W = W = int(inp.init_n_sub_frags)
        if self.N >= 4609:
            raise ValueError("the reference's pixel index is float32-exact below 4,609 bins only (SURVEY F2)")
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.dev)
        host = np.zeros((N_SLOTS, 14, n), dtype=np.int32)
        for fi, k in enumerate(FIELDS):
            host[CUR, fi] = np.ones(n, dtype=np.int32) if k == "ori" else np.asarray(inp.S_o_A_frags[k], dtype=np.int32)
        host[1:, FIELDS.index("ori")] = 1
        host[1:, FIELDS.index("activ")] = 1
        self.slots = t(host, np.int32)
        base = self.slots.data_ptr()
        ptrs = np.array([[base + ((s * 14 + f) * n) * 4 for f in range(14)] for s in range(N_SLOTS)], dtype=np.int64)
        self.frag_structs = t(ptrs, np.int64)                # one 14-pointer struct per slot
        self.id_contigs = torch.zeros((3, n), dtype=torch.int32, device=self.dev)     # pop / trans1 / trans2 side arrays
        self.sub_index = torch.zeros(n, dtype=torch.int32, device=self.dev)
        # level: dense symmetric zero-diagonal sub-level matrix (cuda_lib_gl.py:153-172)
        r, c, v = (np.asarray(a) for a in inp.sub_coo)
        keep = r != c
        r, c, v = t(r[keep], np.int64), t(c[keep], np.int64), t(v[keep], np.float32)
        obs = torch.zeros((W, W), dtype=torch.float32, device=self.dev)
        obs.index_put_((r, c), v, accumulate=True)
        obs.index_put_((c, r), v, accumulate=True)
        self.obs = obs
        self.collector = t(inp.collector_id_repeats, np.int32)
        self.dispatcher = t(inp.frag_dispatcher, np.int32)
        self.sub_id = t(inp.np_sub_frags_id, np.int32)
        self.rep_sub = torch.zeros_like(self.sub_id)
        self.sub_len = t(inp.np_sub_frags_len_bp, np.float32)
        self.sub_accu = t(inp.np_sub_frags_accu, np.int32)
        self.uniq = t(np.arange(self.N), np.int32)
        self.list_rep = t(np.array([-1]), np.int32)
        self.nfpb = float(inp.mean_squared_frags_per_bin)
        self.params = t(np.asarray(params8, dtype=np.float32), np.float32)
        self.n_pix = self.N * (self.N - 1) // 2 + self.N
        self.curr_likelihood = torch.zeros(self.n_pix, dtype=torch.float64, device=self.dev)
        self.out = torch.zeros(16 * N_TMP, dtype=torch.float64, device=self.dev)
        self.stream = torch.cuda.current_stream(self.dev).cuda_stream
        # the 13 streams of stream_likelihood (cuda_lib_gl.py:2441-2448): PyCUDA streams are BLOCKING with respect to the
        # legacy default stream, which orders them after the candidate kernels without explicit events
        self.streams = [_check(driver.cuStreamCreate(0)) for _ in range(N_TMP)]
        self.launches = 0

    # ---------------------------------------------------------------- plumbing
    def _frag(self, slot):
        return self.frag_structs.data_ptr() + slot * 14 * 8

    def _launch(self, name, grid, block, args, types, smem=0, stream=None):
        self.launches += 1
        _check(self.drv.cuLaunchKernel(self.fn[name], int(grid), 1, 1, int(block), 1, 1, int(smem), self.stream if stream is None else stream,
                                       (tuple(args), tuple(types)), 0))

    def slot_to_host(self, slot):
        a = self.slots[slot].cpu().numpy()
        return {k: a[i].copy() for i, k in enumerate(FIELDS)}

    def slot_from_host(self, slot, arrays):
        for i, k in enumerate(FIELDS):
            self.slots[slot, i] = self.torch.from_numpy(np.ascontiguousarray(arrays[k], dtype=np.int32)).to(self.dev)

    # ---------------------------------------------------------------- mutation kernels (grid n // block + 1)
    def move(self, op, dst, src, id_a=0, id_b=0, aux=0, max_id=0, ids=0):
        P, I = C.c_void_p, C.c_int
        n, blk = self.n, self.size_block
        g = n // blk + 1
        d, s = self._frag(dst), self._frag(src)
        side = self.id_contigs[ids].data_ptr()
        if op == "flip":
            self._launch("flip_frag", g, blk, (d, s, id_a, n), (P, P, I, I))
        elif op == "swap_activity":
            self._launch("swap_activity_frag", g, blk, (d, s, id_a, max_id, n), (P, P, I, I, I))
        elif op == "pop_out":
            self._launch("pop_out_frag", g, blk, (d, s, side, id_a, max_id, n), (P, P, P, I, I, I))
        elif op in ("pop_in_1", "pop_in_2", "pop_in_3", "pop_in_4"):
            self._launch("pop_in_frag_" + op[-1], g, blk, (d, s, id_a, id_b, max_id, aux, n), (P, P, I, I, I, I, I))
        elif op == "split":
            self._launch("split_contig", g, blk, (d, s, side, id_a, aux, max_id, n), (P, P, P, I, I, I, I))
        elif op == "paste":
            self._launch("paste_contigs", g, blk, (d, s, id_a, id_b, max_id, n), (P, P, I, I, I, I))
        elif op == "simple_copy":
            self._launch("simple_copy", g, blk, (d, s, n), (P, P, I))
        elif op == "copy_struct":
            self._launch("copy_struct", g, blk, (d, s, side, n), (P, P, P, I))
        else:
            raise ValueError(op)

    def _max(self, ids):
        """ga.max(...) + .get() of the reference (cuda_lib_gl.py:857,934,943): a reduction and a blocking round trip."""
        return int(self.id_contigs[ids].max().item())

    def perform_modifications(self, id_fA, id_fB, max_id):
        """new_perform_modificationS (cuda_lib_gl.py:841-954, 1045-1048): the 13 candidates, 22 launches + 17 max round trips."""
        for mode in range(9):
            self.move("pop_out", POP, CUR, id_fA, max_id=max_id, ids=0)
            m2 = self._max(0)
            dst = CAND0 + mode
            if mode == 0:
                self.move("simple_copy", dst, POP)
            elif mode == 1:
                self.move("flip", dst, CUR, id_fA)
            elif mode in (2, 3):
                self.move("pop_in_1", dst, POP, id_fA, id_fB, 1 if mode == 2 else -1, m2)
            elif mode in (4, 5):
                self.move("pop_in_2", dst, POP, id_fA, id_fB, 1 if mode == 4 else -1, m2)
            elif mode in (6, 7):
                self.move("pop_in_3", dst, POP, id_fA, id_fB, 1 if mode == 6 else -1, m2)
            else:
                self.move("swap_activity", dst, POP, id_fA, max_id=m2)
        mode = 0
        for up_a in (0, 1):
            self.move("split", TRANS1, CUR, id_fA, aux=up_a, max_id=max_id, ids=1)
            for up_b in (0, 1):
                m1 = self._max(1)
                self.move("split", TRANS2, TRANS1, id_fB, aux=up_b, max_id=m1, ids=2)
                m2 = self._max(2)
                self.move("paste", CAND0 + 9 + mode, TRANS2, id_fA, id_fB, max_id=m2)
                mode += 1

    # ---------------------------------------------------------------- likelihood kernels
    def evaluate_likelihood(self, slot=CUR, block=512, stride=50):
        """cuda_lib_gl.py:546-569, 629: per-pixel log-likelihood into curr_likelihood, then ga.sum."""
        P, I, F = C.c_void_p, C.c_int, C.c_float
        triu = self.N * (self.N - 1) // 2
        grid = max(1, int((self.n_pix // block + 1) / stride))
        self._launch("evaluate_likelihood", grid, block,
                     (self.obs.data_ptr(), self._frag(slot), self.collector.data_ptr(), self.dispatcher.data_ptr(), self.sub_id.data_ptr(),
                      self.rep_sub.data_ptr(), self.sub_len.data_ptr(), self.sub_accu.data_ptr(), self.curr_likelihood.data_ptr(),
                      self.params.data_ptr(), triu, self.n_pix, self.N, self.W, self.nfpb),
                     (P, P, P, P, P, P, P, P, P, P, I, I, I, I, F))
        return float(self.curr_likelihood.sum().item())

    def index_sets(self, id_fA, id_fB):
        """fill_sub_index_fA / fB + the host set operations of stream_likelihood (cuda_lib_gl.py:2441-2470)."""
        P, I = C.c_void_p, C.c_int
        cur = self.slots[CUR]
        head = cur[:, [id_fA, id_fB]].cpu().numpy()              # id_c, l_cont of the two bins (the reference reads the host copy)
        cA, cB = int(head[1, 0]), int(head[1, 1])
        lA, lB = int(head[8, 0]), int(head[8, 1])
        self._launch("fill_sub_index_fA", self.n // 1024 + 1, 1024, (self._frag(CUR), self.sub_index.data_ptr(), cA, self.n), (P, P, I, I))
        size = lA
        if cB != cA:
            self._launch("fill_sub_index_fB", self.n // 512 + 1, 512, (self._frag(CUR), self.sub_index.data_ptr(), cB, lA, self.n), (P, P, I, I, I))
            size = lA + lB
        init = self.sub_index[:size].cpu().numpy()
        no_rep = np.setdiff1d(init, np.zeros(0, dtype=np.int32))
        return self.torch.from_numpy(no_rep.astype(np.int32)).to(self.dev)

    def sub_compute_likelihood(self, slot, sub_index_no_rep, out_index, block=512, stride=50, stream=None):
        """cuda_lib_gl.py:2477-2533: one launch (4 KiB dynamic shared memory), atomicAdd into out[out_index]."""
        P, I, F = C.c_void_p, C.c_int, C.c_float
        n_u = int(sub_index_no_rep.shape[0])
        n_no_rep = n_u * (n_u - 1) // 2
        n_values = n_no_rep                                      # no repeats on the comparison levels
        grid = (n_values // block + 1) // stride + 1
        self._launch("sub_compute_likelihood", grid, block,
                     (self.obs.data_ptr(), self._frag(slot), sub_index_no_rep.data_ptr(), self.list_rep.data_ptr(), self.uniq.data_ptr(),
                      self.collector.data_ptr(), self.dispatcher.data_ptr(), self.sub_id.data_ptr(), self.sub_len.data_ptr(), self.sub_accu.data_ptr(),
                      self.out.data_ptr() + 8 * out_index, self.curr_likelihood.data_ptr(), self.params.data_ptr(),
                      n_no_rep, n_no_rep, n_no_rep, n_values, self.N, 0, self.W, self.N, self.nfpb),
                     (P, P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, F), smem=8 * block, stream=stream)

    def stream_likelihood(self, id_fA, id_fB, id_x, max_id):
        """cuda_lib_gl.py:2392-2546 for one neighbour: candidates, index sets, 13 delta launches -> out[13 * id_x + j]."""
        self.perform_modifications(id_fA, id_fB, max_id)
        sub = self.index_sets(id_fA, id_fB)
        self.out[N_TMP * id_x:N_TMP * (id_x + 1)] = 0
        for j in range(N_TMP):
            self.sub_compute_likelihood(CAND0 + j, sub, N_TMP * id_x + j, stream=self.streams[j])
        for j in range(N_TMP):                                    # cuda_lib_gl.py:2535-2546: wait for every stream
            _check(self.drv.cuStreamSynchronize(self.streams[j]))

    def score_step(self, id_fA, neighbours, max_id):
        """The scoring part of step_max_likelihood (cuda_lib_gl.py:1828-1897): full likelihood, then every neighbour."""
        full = self.evaluate_likelihood(CUR)
        for x, fB in enumerate(neighbours):
            self.stream_likelihood(id_fA, fB, x, max_id)
        self.torch.cuda.synchronize(self.dev)
        return full, self.out[:N_TMP * len(neighbours)].cpu().numpy()
