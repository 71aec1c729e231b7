"""Device executor: leaf arena build + plan execution through the C ABI.

PyTorch is used only for device memory, streams, CUDA graphs and (in
``dist.py``) the NCCL process group; every arithmetic kernel is ours.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import runtime as rt
from .core.network import K_DENSE, K_ONES, K_VEC, Network
from .planner import ARENA, OUT, WS, Plan, Step, plan_from_ssa_path, plan_network


def _torch():
    import torch
    return torch


def require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise rt.B2Error(
            "optyx_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def leaf_table(net: Network, plan: Plan, varying=None) -> Tuple[np.ndarray, np.ndarray]:
    """Descriptor table (+ the unit leaf) and the flat params vector of one eval.
    ``varying``: indices of the host-valued leaves whose values differ per eval of a batch
    (they index the per-eval params buffer); the others are flagged ``shared`` and index
    one vector for the whole batch.  ``None``: every host-valued leaf is per-eval, and the
    returned vector holds them all."""
    if varying is None:
        # everything in the table follows from the network STRUCTURE (kinds, shapes, integer
        # parameters -- Network.structure_key, which is also what the plan is cached by), so
        # it is built once per plan; per eval only the host values are gathered (cfg4: 3588
        # leaves, 25 ms of per-row Python on every eval before)
        cached = getattr(plan, "_leaf_table_static", None)
        if cached is not None and len(cached[0]) == len(net.leaves) + 1:
            table, host_leaves = cached
            ps = [np.asarray(net.leaves[k].params, dtype=np.complex128).reshape(-1) for k in host_leaves]
            return table, (np.concatenate(ps) if ps else np.zeros(1, dtype=np.complex128))
    n = len(net.leaves)
    table = np.zeros(n + 1, dtype=rt.LEAF_DESC_DTYPE)
    params: List[np.ndarray] = []
    poff = soff = 0
    for k, leaf in enumerate(net.leaves):
        row = table[k]
        row["kind"], row["ipar"] = leaf.kind, leaf.ipar
        row["offset"], row["nelem"] = plan.leaf_offsets[k], leaf.size
        if leaf.kind in (K_VEC, K_DENSE):
            # host-valued leaves are copied element by element: flat on the device, any
            # rank (the planner alone knows their shape and indices)
            row["ndim"] = 1
            row["dims"][0] = min(leaf.size, 2 ** 31 - 1)
            if varying is None or k in varying:
                row["poff"] = poff
                if varying is None:
                    params.append(np.asarray(leaf.params, dtype=np.complex128).reshape(-1))
                poff += leaf.size
            else:
                row["shared"], row["poff"] = 1, soff
                soff += leaf.size
        else:
            row["ndim"] = len(leaf.shape)
            row["dims"][:len(leaf.shape)] = leaf.shape
    unit = table[n]
    unit["kind"], unit["ndim"] = K_ONES, 1
    unit["dims"][0] = 1
    unit["offset"], unit["nelem"] = plan.unit_offset, 1
    pvec = np.concatenate(params) if params else np.zeros(1, dtype=np.complex128)
    if varying is None:
        table.setflags(write=False)
        plan._leaf_table_static = (table, [k for k, l in enumerate(net.leaves)
                                           if l.kind in (K_VEC, K_DENSE)])
    return table, pvec


def params_vector(net: Network) -> np.ndarray:
    ps = [np.asarray(l.params, dtype=np.complex128).reshape(-1)
          for l in net.leaves if l.kind in (K_VEC, K_DENSE)]
    return np.concatenate(ps) if ps else np.zeros(1, dtype=np.complex128)


BRANCH_STREAMS = 32      # capture streams = concurrent branches of the contraction tree
                         # (cfg2's 394 small steps: 0.23 ms with 8, 0.16 with 16, 0.13-0.17 with 32)


class Session:
    """Device buffers for one plan (arena, workspace, output) and its launcher."""

    def __init__(self, plan: Plan, device=None):
        torch = require_cuda()
        self.torch = torch
        self.lib = rt.load()
        self.plan = plan
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.tdtype = torch.complex128 if plan.dtype == np.complex128 else torch.complex64
        self.code = rt.dtype_code(plan.dtype)
        self.esize = plan.dtype.itemsize
        b = plan.batch
        self.arena = torch.empty(b * plan.arena_elems, dtype=self.tdtype, device=self.device)
        self.ws = torch.empty(max(1, plan.ws_elems), dtype=self.tdtype, device=self.device)
        self.out = torch.empty((b,) + tuple(plan.out_shape), dtype=self.tdtype,
                               device=self.device)
        self.desc_dev = None
        self.params_dev = None
        self.params_host = None
        self.shared_dev = None
        self._params_copied = None
        self.n_leaves = 0
        self.params_stride = 0
        self.graph = None
        self.launches = 0
        self._init_small_program()
        self._init_step_table()

    # ---- small-step program ----------------------------------------------------
    def _init_small_program(self) -> None:
        """Device tables of the persistent small-step launch (``plan.small``)."""
        plan, torch = self.plan, self.torch
        sp = plan.small
        self.small_args = None
        if sp is None:
            return

        def up(arr):
            raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
            if raw.size == 0:
                raw = np.zeros(16, dtype=np.uint8)
            return torch.from_numpy(raw.copy()).to(self.device)
        self._sp_bufs = [up(sp.blob), up(sp.blob_off), up(sp.out_tab), up(sp.red_tab)]
        # [ticket, finished CTAs, epoch, pad] + one done-flag per (eval, task); zeroed once
        self._sp_sync = torch.zeros(4 + plan.batch * sp.n_tasks, dtype=torch.int32,
                                    device=self.device)
        args = rt.SmallProgramArgs()
        args.blob_dev, args.blob_off_dev, args.out_tab_dev, args.red_tab_dev = \
            [b.data_ptr() for b in self._sp_bufs]
        args.base_dev[ARENA] = self.arena.data_ptr()
        args.base_dev[WS] = self.ws.data_ptr()
        args.base_dev[OUT] = self.out.data_ptr()
        args.sync_dev = self._sp_sync.data_ptr()
        args.n_tasks, args.batch = sp.n_tasks, plan.batch
        args.smem_bytes = sp.smem_bytes(self.esize)
        self.small_args = args

    def _init_step_table(self) -> None:
        """Host table ``b2_run_plan`` walks: every step that is not part of the
        small-step program, in plan order; no pointers (buffers are bound per call)."""
        plan = self.plan
        if len(plan.sliced) > rt.MAX_SLICED:
            raise ValueError(f"{len(plan.sliced)} sliced indices: the step table holds {rt.MAX_SLICED}")
        rows = [st for st in plan.steps if not st.grouped]
        table = (rt.PlanStep * max(1, len(rows)))()
        for row, st in zip(table, rows):
            row.kind = rt.STEP_GEMM if st.kernel == "gemm" else rt.STEP_DIRECT
            row.per_slice, row.final = int(st.per_slice), int(st.final)
            row.a_space, row.b_space, row.c_space = st.a.space, st.b.space, st.c.space
            row.a_off, row.b_off, row.c_off = st.a.offset, st.b.offset, st.c.offset
            row.c_elems = st.c.size * plan.batch
            row.n_a_slice, row.n_b_slice = len(st.a_slice), len(st.b_slice)
            for i, (pos, stride) in enumerate(st.a_slice):
                row.a_slice_pos[i], row.a_slice_stride[i] = pos, stride
            for i, (pos, stride) in enumerate(st.b_slice):
                row.b_slice_pos[i], row.b_slice_stride[i] = pos, stride
            if st.kernel == "gemm":
                row.desc.gemm = st.desc
            else:
                row.desc.direct = st.desc
        self.step_table, self.n_table = table, len(rows)
        self.sliced_dims = (C.c_int32 * max(1, len(plan.sliced)))(*plan.sliced_dims)
        self.n_invariant = sum(1 for st in rows if not st.per_slice)
        self._launch_count = C.c_int32(0)

    def run_plan(self, what: int, scalar: complex, slice_range=(0, 0), stream=None) -> None:
        """ONE call into the C plan interpreter (``b2_run_plan``)."""
        plan = self.plan
        if stream is None:
            stream = self.torch.cuda.current_stream().cuda_stream
        scalar = complex(scalar)
        small = C.byref(self.small_args) if self.small_args is not None else None
        rt.check(self.lib.b2_run_plan(
            self.step_table, self.n_table, self.sliced_dims, len(plan.sliced), small,
            self.arena.data_ptr(), self.ws.data_ptr(), self.out.data_ptr(), self.out.numel(),
            what, slice_range[0], slice_range[1], scalar.real, scalar.imag, self.code, stream,
            C.byref(self._launch_count)), "b2_run_plan")
        self.launches += self._launch_count.value

    def _launch_small_groups(self, stream) -> None:
        if self.small_args is not None:
            rt.check(self.lib.b2_run_small_program(C.byref(self.small_args), self.code, stream),
                     "b2_run_small_program")
            self.launches += 1

    # ---- leaves -------------------------------------------------------------
    def set_leaves(self, nets: Sequence[Network]) -> None:
        """Upload the descriptor table (structure of nets[0]) and every eval's
        host-computed params (pinned staging buffer, async copy)."""
        torch = self.torch
        plan = self.plan
        assert len(nets) == plan.batch
        if (plan.batch == 1 and self.desc_dev is not None and getattr(self, "_varying", None) is None
                and getattr(self, "_uploaded_net", None) is nets[0]):
            # the very same (cached, immutable) Network object as the last upload: its values
            # are already on the device -- a re-built identical diagram maps to it through
            # the fingerprint cache
            return
        table, p0 = leaf_table(nets[0], plan)
        if self.desc_dev is None or getattr(self, "_varying", None) is not None:
            self._upload_table(table)
            self._varying = None
            self.shared_dev = None
            self.graph = None            # a captured graph holds the old buffers
            self.params_stride = len(p0)
            self.params_host = torch.empty((plan.batch, len(p0)), dtype=torch.complex128,
                                           pin_memory=True)
            self.params_dev = torch.empty((plan.batch, len(p0)), dtype=torch.complex128,
                                          device=self.device)
        if self._params_copied is not None:
            # the previous async H2D copy may still be reading the pinned buffer
            self._params_copied.synchronize()
        ph = self.params_host.numpy()
        ph[0] = p0
        for e in range(1, plan.batch):
            ph[e] = params_vector(nets[e])
        self.params_dev.copy_(self.params_host, non_blocking=True)
        self.h2d_bytes = int(self.params_host.numel() * 16 + self.desc_dev.numel())
        if self._params_copied is None:
            self._params_copied = torch.cuda.Event()
        self._params_copied.record()
        self._uploaded_net = nets[0] if plan.batch == 1 else None

    def set_leaves_batched(self, net: Network) -> None:
        """The same for ONE network whose parameters are array-valued (``net.batch``
        evals compiled in one pass).  Only the leaves that really vary get a per-eval
        column block ``[batch, P_var]``; every other host value is uploaded once, in one
        vector the array-build kernel shares between the evals."""
        torch, plan = self.torch, self.plan
        batch = net.batch or 1
        assert batch == plan.batch, (batch, plan.batch)
        self._uploaded_net = None
        varying = frozenset(k for k, l in enumerate(net.leaves)
                            if l.params is not None and l.params.ndim == 2)
        if self.desc_dev is None or getattr(self, "_varying", None) != varying:
            table, _ = leaf_table(net, plan, varying=varying)
            self._upload_table(table)
            self._varying = varying
            self.graph = None            # a captured graph holds the old buffers
            p_var = max(1, sum(net.leaves[k].size for k in varying))
            p_sh = max(1, sum(l.size for k, l in enumerate(net.leaves)
                              if l.params is not None and k not in varying))
            self.params_stride = p_var
            self.params_host = torch.empty((batch, p_var), dtype=torch.complex128, pin_memory=True)
            self.params_dev = torch.empty((batch, p_var), dtype=torch.complex128, device=self.device)
            self.shared_host = torch.empty(p_sh, dtype=torch.complex128, pin_memory=True)
            self.shared_dev = torch.empty(p_sh, dtype=torch.complex128, device=self.device)
        if self._params_copied is not None:
            self._params_copied.synchronize()
        ph, sh = self.params_host.numpy(), self.shared_host.numpy()
        pv = ps = 0
        for k, l in enumerate(net.leaves):
            if l.params is None:
                continue
            if k in varying:
                ph[:, pv:pv + l.size] = l.params.T
                pv += l.size
            else:
                sh[ps:ps + l.size] = l.params
                ps += l.size
        self.params_dev.copy_(self.params_host, non_blocking=True)
        self.shared_dev.copy_(self.shared_host, non_blocking=True)
        self.h2d_bytes = int((self.params_host.numel() + self.shared_host.numel()) * 16
                             + self.desc_dev.numel())
        if self._params_copied is None:
            self._params_copied = torch.cuda.Event()
        self._params_copied.record()

    def _upload_table(self, table: np.ndarray) -> None:
        torch, plan = self.torch, self.plan
        self.n_leaves = len(table)
        self.desc_dev = torch.from_numpy(table.view(np.uint8).copy()).to(self.device)
        # leaf holding the first element of every 256-element block of the arena
        firsts = np.arange(0, plan.arena_elems, 256, dtype=np.int64)
        firsts = np.append(firsts, plan.arena_elems - 1)
        block_leaf = np.searchsorted(table["offset"], firsts, side="right") - 1
        self.block_leaf_dev = torch.from_numpy(block_leaf.astype(np.int32)).to(self.device)

    def build_leaves(self) -> None:
        plan = self.plan
        stream = self.torch.cuda.current_stream().cuda_stream
        rt.check(self.lib.b2_build_leaves_indexed(
            self.desc_dev.data_ptr(), self.n_leaves, self.block_leaf_dev.data_ptr(), plan.arena_elems,
            self.params_dev.data_ptr(), self.params_stride,
            self.shared_dev.data_ptr() if self.shared_dev is not None else None,
            self.arena.data_ptr(), plan.arena_elems, plan.batch, self.code, stream), "b2_build_leaves")
        self.launches += 1

    # ---- steps --------------------------------------------------------------
    def _ptr(self, lay, slice_vals, terms) -> int:
        base = {ARENA: self.arena, WS: self.ws, OUT: self.out}[lay.space].data_ptr()
        off = lay.offset
        for pos, stride in terms:
            off += slice_vals[pos] * stride
        return base + off * self.esize

    def _launch(self, st: Step, slice_vals, alpha: complex, accumulate: int, stream) -> None:
        a = self._ptr(st.a, slice_vals, st.a_slice)
        b = self._ptr(st.b, slice_vals, st.b_slice)
        c = self._ptr(st.c, slice_vals, ())
        if st.kernel == "gemm":
            rt.check(self.lib.b2_contract_gemm(C.byref(st.desc), a, b, c, alpha.real, alpha.imag,
                                               accumulate, self.code, stream), "b2_contract_gemm")
        else:
            rt.check(self.lib.b2_contract_direct(C.byref(st.desc), a, b, c, alpha.real,
                                                 alpha.imag, accumulate, self.code, stream),
                     "b2_contract_direct")
        self.launches += 1

    def contract(self, scalar: complex = 1.0, slice_range: Optional[Tuple[int, int]] = None,
                 zero_out: bool = True) -> None:
        """Run the plan: invariant steps once, per-slice steps for every slice id
        in ``slice_range`` (default all), accumulating into ``self.out``."""
        plan = self.plan
        s0, s1 = slice_range if slice_range is not None else (0, plan.nslices)
        zero = rt.RUN_ZERO if plan.needs_zero_out and zero_out else 0
        slices = rt.RUN_SLICES if plan.sliced else 0
        if plan.sliced and self.n_invariant > 32 and not self.torch.cuda.is_current_stream_capturing():
            # many slice-invariant launches left outside the small-step program: one CUDA
            # graph (captured on first use) instead of one launch each
            if getattr(self, "inv_graph", None) is None:
                self._capture_invariant()
            if zero:
                self.run_plan(rt.RUN_ZERO, scalar)
            self.inv_graph.replay()
            self.run_plan(slices, scalar, (s0, s1))
        else:
            self.run_plan(zero | rt.RUN_INVARIANT | slices, scalar, (s0, s1))

    def _capture_invariant(self) -> None:
        """CUDA graph of the slice-invariant steps (none of them is final: alpha = 1;
        their buffers are private, ``planner._reallocate_for_slices``)."""
        torch = self.torch

        def body():
            self.run_plan(rt.RUN_INVARIANT, 1.0 + 0j)
        launches = self.launches
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()                                   # warm-up outside the capture
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            body()
        self.launches = launches
        self.inv_graph = graph

    def take_out(self):
        """Hand the output buffer to the caller (a result object that may read it much
        later): the session gets a fresh one before its next run.  With a captured graph
        the buffer address is frozen into the graph, so the caller gets a copy."""
        out = self.out
        if self.graph is not None:
            return out.clone()
        self.out = self.torch.empty_like(out)
        if self.small_args is not None:
            self.small_args.base_dev[OUT] = self.out.data_ptr()
        return out

    def run(self, nets: Sequence[Network], scalar: Optional[complex] = None):
        """Full device path for a batch of structurally identical networks."""
        self.set_leaves(nets)
        self.build_leaves()
        self.contract(nets[0].scalar if scalar is None else scalar)
        return self.out

    # ---- CUDA graph ---------------------------------------------------------
    def contract_branches(self, scalar: complex, n_streams: Optional[int] = None) -> None:
        """Unsliced plan with independent subtrees issued on different streams
        (fork after the leaf build, join before the end).  Under stream capture
        the event waits become graph edges, so the captured graph is the
        contraction TREE, not a chain: tiny steps of different branches overlap
        instead of paying one launch latency each."""
        torch, plan = self.torch, self.plan
        assert plan.parallel_safe and not plan.sliced
        n_streams = BRANCH_STREAMS if n_streams is None else n_streams
        origin = torch.cuda.current_stream()
        if plan.needs_zero_out:
            rt.check(self.lib.b2_fill_zero(self.out.numel(), self.out.data_ptr(), self.code,
                                           origin.cuda_stream), "b2_fill_zero")
        if not hasattr(self, "_pool") or len(self._pool) != n_streams:
            self._pool = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        pool = self._pool
        self._launch_small_groups(origin.cuda_stream)        # tiny subtrees: one launch, before the fork
        fork = torch.cuda.Event()
        fork.record(origin)
        for s in pool:
            s.wait_event(fork)
        producer = {}                       # id(layout) -> step index
        for k, st in enumerate(plan.steps):
            if not st.grouped:
                producer[id(st.c)] = k
        where = {}                          # step index -> stream index
        consumers_elsewhere = {}            # step index -> event (recorded after launch)
        # first pass: stream assignment (inherit the larger operand's stream)
        rr = 0
        for k, st in enumerate(plan.steps):
            if st.grouped:
                continue
            pa, pb = producer.get(id(st.a)), producer.get(id(st.b))
            if pa is not None:
                where[k] = where[pa]
            elif pb is not None:
                where[k] = where[pb]
            else:
                where[k] = rr % n_streams
                rr += 1
        need_event = set()
        for k, st in enumerate(plan.steps):
            if st.grouped:
                continue
            for p in (producer.get(id(st.a)), producer.get(id(st.b))):
                if p is not None and where[p] != where[k]:
                    need_event.add(p)
        events = {}
        zeros = [0] * len(plan.sliced)
        scalar = complex(scalar)
        for k, st in enumerate(plan.steps):
            if st.grouped:
                continue
            stream = pool[where[k]]
            for p in (producer.get(id(st.a)), producer.get(id(st.b))):
                if p is not None and where[p] != where[k]:
                    stream.wait_event(events[p])
            self._launch(st, zeros, scalar if st.final else 1.0 + 0j, 0, stream.cuda_stream)
            if k in need_event:
                ev = torch.cuda.Event()
                ev.record(stream)
                events[k] = ev
        for s in pool:
            join = torch.cuda.Event()
            join.record(s)
            origin.wait_event(join)

    def capture(self, scalar: complex = 1.0, branches: bool = True) -> None:
        """Capture leaf build + all contraction launches into one CUDA graph
        (launch-bound small networks: hundreds of tiny steps per eval)."""
        torch = self.torch
        assert self.desc_dev is not None, "set_leaves first"
        use_branches = branches and self.plan.parallel_safe and not self.plan.sliced

        def body():
            self.build_leaves()
            if use_branches:
                self.contract_branches(scalar)
            else:
                self.contract(scalar)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()
        self.graph = graph

    def replay(self) -> None:
        self.graph.replay()


_PLAN_CACHE: Dict[tuple, Plan] = {}


def get_plan(net: Network, dtype=np.complex128, batch: int = 1, target_size=None,
             trials: int = 8, seed: int = 0, gemm_threshold=(32, 16, 8, 4.0),
             ssa_path=None, sliced_inds=(), min_slices: int = 1) -> Plan:
    """Plan cache keyed by network *structure* (the counterpart of cotengra's
    ReusableHyperOptimizer the reference accepts, backends.py:49-56, 436-455).
    ``ssa_path`` (+ ``sliced_inds``): execute that contraction tree instead of
    searching for one (``planner.plan_from_ssa_path``)."""
    tree_key = None if ssa_path is None else (tuple(tuple(int(i) for i in e) for e in ssa_path),
                                              tuple(int(q) for q in sliced_inds))
    key = (net.structure_key(), np.dtype(dtype).str, batch, target_size, trials, seed,
           tuple(gemm_threshold), tree_key, min_slices)
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if ssa_path is not None:
            plan = plan_from_ssa_path(net, ssa_path, sliced_inds, dtype=dtype, batch=batch,
                                      gemm_threshold=gemm_threshold)
        else:
            plan = plan_network(net, dtype=dtype, batch=batch, trials=trials, seed=seed,
                                target_size=target_size, gemm_threshold=gemm_threshold,
                                min_slices=min_slices)
        _PLAN_CACHE[key] = plan
    return plan


def evaluate(net: Network, dtype=np.complex128, target_size=None, **kw):
    """One-shot convenience: plan, build, contract; returns the device tensor
    of shape ``net.output_shape``."""
    plan = get_plan(net, dtype=dtype, target_size=target_size, **kw)
    sess = Session(plan)
    out = sess.run([net])
    return out[0]
