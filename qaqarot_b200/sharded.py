"""A statevector sharded over the GPUs of one box: one process per GPU, ``torch.distributed`` for the plumbing.

The reference has no distributed code at all; this is new.  The 2^n amplitudes are split on
the top ``log2(P)`` index bits: rank r holds the amplitudes whose *global* bits equal r, as an ordinary
``n_local = n - log2(P)``-qubit device state whose ``index_offset`` tells the kernels which shard it is
(include/b200sv.h: b200_state_set_index_offset).  Consequences:

* gates whose non-diagonal target lives on a local bit run with no communication; controls, phases, fans and
  Pauli-Z signs on global bits cost nothing either (per-shard constants);
* a non-diagonal gate on a global wire makes the planner emit an EXCHANGE op (planner.normalize/bring_home):
  the global bit swaps places with the top local bit, i.e. every rank trades one *contiguous* half of its shard
  with the partner rank whose shard number differs in that bit -- ``isend``/``irecv`` over NCCL (NVLink), through
  a bounded bounce buffer because a 36-qubit shard (128 GiB) leaves no room for a second copy;
* the wire -> index-bit layout is tracked (``layout``) and *not* undone after a program: measurement, sampling
  and expectation values translate wire numbers instead, and only ``download`` (small registers) un-permutes;
* scalars (norms, conditional probabilities, expectation partial sums) are all-reduced.

The local state is pluggable so that the same logic runs on CPU under ``gloo`` with the numpy program
interpreter standing in for the device (tests/test_sharded.py); on a GPU box it is a ``DeviceState`` wrapping a
torch CUDA tensor, so that NCCL can address the amplitudes directly.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .program import Program


def _torch():
    import torch
    import torch.distributed as dist
    return torch, dist


class _DevicePointer:
    """Exposes device memory owned by the C library to torch (``__cuda_array_interface__``), zero copy."""

    def __init__(self, ptr: int, n_doubles: int):
        self.__cuda_array_interface__ = {"shape": (n_doubles,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class _CudaLocal:
    """n_local-qubit DeviceState (memory owned by libb200sv) plus a torch view of it for torch.distributed."""

    def __init__(self, n_local: int, index_offset: int, device: int):
        torch, _ = _torch()
        from .device import DeviceState
        torch.cuda.set_device(device)
        self.state = DeviceState(n_local, device)
        self.state.set_index_offset(index_offset)
        self.tensor = torch.as_tensor(_DevicePointer(self.state.device_ptr, 2 << n_local), device=f"cuda:{device}")
        self.device = device
        self.peers = {}                        # partner rank -> mapped device pointer of its shard

    def buffer(self):
        return self.tensor

    def before_comm(self):
        self.state.sync()                      # kernels run on the state's own stream

    def after_comm(self):
        torch, _ = _torch()
        torch.cuda.synchronize(self.device)

    def close(self):
        for p in self.peers.values():
            try:
                self.state.ipc_close(p)
            except Exception:
                pass
        self.peers = {}
        self.tensor = None
        self.state.close()


class _HostLocal:
    """CPU stand-in: the numpy interpreter of the op program (test infrastructure injects the class)."""

    def __init__(self, state):
        torch, _ = _torch()
        self.state = state
        self.tensor = torch.from_numpy(state.psi.view(np.float64))

    def buffer(self):
        return self.tensor

    def before_comm(self):
        pass

    def after_comm(self):
        pass

    def close(self):
        pass


class ShardedState:
    """Same method surface as ``device.DeviceState``, in *wire* coordinates, collective across the process group.
    Every rank must make the same calls in the same order."""

    BOUNCE_AMPS = 1 << 24           # 256 MiB bounce buffer for exchanges
    MIN_LOCAL = 9                   # smallest tile csrc/tile.cu is built for (planner.PlanConfig.kernel_tiles)

    def __init__(self, n_qubits: int, device: Optional[int] = None, group=None, host_state_factory=None):
        torch, dist = _torch()
        if not dist.is_initialized():
            raise RuntimeError("ShardedState needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        g = self.world.bit_length() - 1
        if 1 << g != self.world:
            raise ValueError(f"world size {self.world} is not a power of two")
        if n_qubits - g < 1:
            raise ValueError(f"{n_qubits} qubits cannot be sharded over {self.world} ranks")
        if host_state_factory is None and n_qubits - g < self.MIN_LOCAL:
            raise ValueError(f"{n_qubits} qubits over {self.world} GPUs leaves {n_qubits - g} local qubits; the fused tile "
                             f"pass needs at least {self.MIN_LOCAL} (use B200Backend on one GPU for registers this small)")
        self.n_qubits = n_qubits
        self.n_global = g
        self.n_local = n_qubits - g
        self.layout: List[int] = list(range(n_qubits))       # wire -> index bit
        offset = self.rank << self.n_local
        if host_state_factory is not None:
            self.local = _HostLocal(host_state_factory(self.n_local, offset))
        else:
            self.local = _CudaLocal(self.n_local, offset, 0 if device is None else device)
        self.device = device
        self._bounce = None
        self.exchange_bytes = 0          # bytes this rank has sent (= received) in exchanges
        self.exchange_ms = 0.0           # wall time spent in exchanges (synchronised), for reporting
        self.n_exchanges = 0             # global wires brought home
        self.n_remaps = 0                # ... in this many steps of two barriers each (exchange_many)
        self.profile = None              # list of (kinds, ms) per executed segment while profiling is on
        self.use_p2p = True              # exchanges by the in-place peer-to-peer kernel (GPUs only)

    def set_profiling(self, on: bool) -> None:
        self.st.set_profiling(on)
        self.profile = [] if on else None

    # -- helpers ----------------------------------------------------------------------------------
    @property
    def st(self):
        return self.local.state

    @property
    def dim(self) -> int:
        return 1 << self.n_qubits

    def _allreduce(self, values: np.ndarray) -> np.ndarray:
        torch, dist = _torch()
        t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64).copy())
        if dist.get_backend(self.group) == "nccl":
            t = t.cuda(self.local.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def _my_bit(self, index_bit: int) -> int:
        return (self.rank >> (index_bit - self.n_local)) & 1

    def _phys_index(self, wire_index: int) -> int:
        p = 0
        for w in range(self.n_qubits):
            if (wire_index >> w) & 1:
                p |= 1 << self.layout[w]
        return p

    def _wire_at(self) -> List[int]:
        inv = [0] * self.n_qubits
        for w, p in enumerate(self.layout):
            inv[p] = w
        return inv

    # -- lifetime / contents ----------------------------------------------------------------------
    def close(self) -> None:
        if self.local is not None:
            self.local.close()
            self.local = None
        self._bounce = None

    def set_basis(self, index: int = 0) -> None:
        self.layout = list(range(self.n_qubits))
        owner, local_index = index >> self.n_local, index & ((1 << self.n_local) - 1)
        if owner == self.rank:
            self.st.set_basis(local_index)
        else:
            self.st.set_basis(0)
            self.st.scale(0.0)

    def upload(self, vec: np.ndarray, offset: int = 0) -> None:
        """`vec`: the whole register in wire order (every rank passes the same array)."""
        assert offset == 0 and vec.size == self.dim
        self.layout = list(range(self.n_qubits))
        lo = self.rank << self.n_local
        self.st.upload(np.ascontiguousarray(vec[lo:lo + (1 << self.n_local)], dtype=np.complex128))

    def download_shard(self) -> np.ndarray:
        """This rank's amplitudes in physical order (index bit k of the local index = index bit k)."""
        return self.st.download()

    def download(self, offset: int = 0, count: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """The whole register in wire order, on every rank.  Small registers only (2^n amplitudes on the host)."""
        torch, dist = _torch()
        if self.n_qubits > 30:
            raise ValueError("download() of a sharded register above 30 qubits: use download_shard() and `layout`")
        mine = torch.from_numpy(self.st.download().view(np.float64).copy())
        nccl = dist.get_backend(self.group) == "nccl"
        if nccl:
            mine = mine.cuda(self.local.device)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        phys = np.concatenate([p.cpu().numpy() for p in parts]).view(np.complex128)
        if self.layout != list(range(self.n_qubits)):
            idx = np.arange(self.dim, dtype=np.int64)
            src = np.zeros(self.dim, dtype=np.int64)
            for w, p in enumerate(self.layout):
                src |= ((idx >> w) & 1) << p
            phys = phys[src]
        res = phys if offset == 0 and count is None else phys[offset:offset + (count or self.dim - offset)]
        if out is not None:
            out[:] = res
            return out
        return np.ascontiguousarray(res)

    # -- checkpoints (a 33-36 qubit state cannot return through one ndarray) -------------------
    CHECKPOINT_FORMAT = "b200sv-sharded-v1"

    CHECKPOINT_CHUNK = 1 << 24       # amplitudes per transfer (256 MiB): a 128 GiB shard never sits in host memory

    def _all_ok(self, ok: bool, what: str) -> None:
        """Collective agreement before anything irreversible: every rank raises together or none does."""
        bad = self._allreduce(np.array([0.0 if ok else 1.0]))[0]
        if bad:
            raise ValueError(f"{what} ({int(bad)} of {self.world} ranks failed" + ("" if ok else ", this one included") + ")")

    def save(self, directory: str) -> None:
        """Every rank streams its shard as raw little-endian complex128 (`shard_<rank>.c128`, physical index order) in
        256 MiB pieces; rank 0 then writes `header.json` -- register size, world size, wire -> index-bit layout, without
        which the shards cannot be interpreted (swap gates and exchanges only relabel wires) -- to a temporary name and
        renames it into place last, so a directory with a header is a complete checkpoint."""
        import json
        import os
        _, dist = _torch()
        err = None
        try:
            if self.rank == 0:
                os.makedirs(directory, exist_ok=True)
                try:
                    os.remove(os.path.join(directory, "header.json"))         # an old header must not bless new shards
                except FileNotFoundError:
                    pass
        except OSError as e:
            err = e
        dist.barrier(group=self.group)
        if err is None:
            try:
                total, chunk = 1 << self.n_local, self.CHECKPOINT_CHUNK
                with open(os.path.join(directory, f"shard_{self.rank}.c128"), "wb") as f:
                    for off in range(0, total, chunk):
                        self.st.download(off, min(chunk, total - off)).astype("<c16", copy=False).tofile(f)
                    f.flush()
                    os.fsync(f.fileno())
            except OSError as e:
                err = e
        self._all_ok(err is None, f"checkpoint {directory}: writing a shard failed: {err}")
        if self.rank == 0:
            tmp = os.path.join(directory, "header.json.tmp")
            with open(tmp, "w") as f:
                json.dump({"format": self.CHECKPOINT_FORMAT, "n_qubits": self.n_qubits, "world": self.world,
                           "n_local": self.n_local, "dtype": "complex128", "byteorder": "little",
                           "layout_wire_to_index_bit": list(self.layout)}, f)
                f.flush()
                os.fsync(f.fileno())
            os.replace(tmp, os.path.join(directory, "header.json"))
        dist.barrier(group=self.group)

    def load(self, directory: str) -> None:
        """Restore a checkpoint written by `save` on the same world size (shards are not re-cut).  Header and shard
        sizes are validated on every rank and agreed on collectively before the state is touched."""
        import json
        import os
        problem, h = None, {}
        try:
            with open(os.path.join(directory, "header.json")) as f:
                h = json.load(f)
            layout = [int(b) for b in h.get("layout_wire_to_index_bit", [])]
            if h.get("format") != self.CHECKPOINT_FORMAT:
                problem = f"unknown format {h.get('format')!r}"
            elif (h.get("n_qubits"), h.get("world"), h.get("n_local")) != (self.n_qubits, self.world, self.n_local):
                problem = (f"written for {h.get('n_qubits')} qubits on {h.get('world')} ranks ({h.get('n_local')} local), this "
                           f"register has {self.n_qubits} on {self.world} ({self.n_local} local)")
            elif h.get("dtype") != "complex128" or h.get("byteorder") != "little":
                problem = f"dtype {h.get('dtype')!r} / byteorder {h.get('byteorder')!r} not supported"
            elif sorted(layout) != list(range(self.n_qubits)):
                problem = "layout_wire_to_index_bit is not a permutation of the index bits"
            else:
                size = os.path.getsize(os.path.join(directory, f"shard_{self.rank}.c128"))
                if size != 16 << self.n_local:
                    problem = f"shard_{self.rank}.c128 has {size} bytes, expected {16 << self.n_local}"
        except (OSError, ValueError, TypeError) as e:
            problem = str(e)
        self._all_ok(problem is None, f"checkpoint {directory} cannot be loaded: {problem}")
        total, chunk = 1 << self.n_local, self.CHECKPOINT_CHUNK
        with open(os.path.join(directory, f"shard_{self.rank}.c128"), "rb") as f:
            for off in range(0, total, chunk):
                part = np.fromfile(f, dtype="<c16", count=min(chunk, total - off))
                self.st.upload(part.astype(np.complex128, copy=False), off)
        self.layout = layout

    def copy_from(self, other: "ShardedState") -> None:
        self.st.copy_from(other.st)
        self.layout = list(other.layout)

    def clone(self) -> "ShardedState":
        """A second register of the same size (the caller must know it fits: snapshots for mid-circuit
        measurement with several shots, save_cache)."""
        factory = None
        if isinstance(self.local, _HostLocal):
            proto = self.local.state
            factory = lambda nl, off: type(proto)(nl, 0, off)
        c = ShardedState(self.n_qubits, self.device, self.group, factory)
        c.copy_from(self)
        return c

    # -- evolution --------------------------------------------------------------------------------
    def apply(self, program: Program, init_basis: Optional[int] = None) -> None:
        if init_basis is not None:
            assert init_basis == 0, "only |0...0> is layout independent"
            # |0...0> looks the same in every wire layout: adopt the one the program was planned from
            self.layout = list(program.initial_pos) if program.initial_pos else list(range(self.n_qubits))
            segs = program.segments()
            if not segs or not isinstance(segs[0], Program):
                layout = self.layout
                self.set_basis(init_basis)
                self.layout = layout
                init_basis = None
        if len(program) == 0 and not program.final_pos:
            return
        if program.initial_pos and list(program.initial_pos) != self.layout:
            raise ValueError("program was planned for a different wire layout than the state is in")
        segs = program.segments()
        k = 0
        while k < len(segs):
            seg = segs[k]
            if isinstance(seg, Program):
                self.st.apply(seg, init_basis=init_basis if k == 0 else None)
                if self.profile is not None:
                    ms, kinds = self.st.last_op_times()
                    self.profile.append((kinds.copy(), ms.copy()))
                k += 1
                continue
            # consecutive exchanges on distinct bits are ONE remap of several global wires
            run = [seg]
            while (k + len(run) < len(segs) and not isinstance(segs[k + len(run)], Program)
                   and segs[k + len(run)][0] not in [l for l, _ in run] and segs[k + len(run)][1] not in [g for _, g in run]
                   and len(run) < 8):
                run.append(segs[k + len(run)])
            self.exchange_many(run)
            k += len(run)
        if program.final_pos:
            self.layout = list(program.final_pos)

    def exchange(self, lbit: int, gpos: int) -> None:
        """Swap local index bit `lbit` with global bit `gpos`: this rank trades the half of its shard whose bit
        `lbit` differs from its own bit `gpos` with the partner rank.  The wire layout is the caller's job.

        On GPUs the two partners swap in place over NVLink with one kernel each (csrc/permute.cu k_peer_swap on
        the partner's shard mapped through CUDA IPC), each taking one half of the pairs; ``use_p2p = False`` or a
        host stand-in falls back to NCCL / gloo send-recv of the contiguous top half through a bounce buffer
        (with a local transposition before and after when `lbit` is not the top bit)."""
        torch, dist = _torch()
        import time
        assert 0 <= lbit < self.n_local <= gpos < self.n_qubits
        partner = self.rank ^ (1 << (gpos - self.n_local))
        half = 1 << (self.n_local - 1)                          # amplitudes that leave (and arrive)
        self.local.before_comm()
        dist.barrier(group=self.group)                          # both shards are quiescent
        t0 = time.perf_counter()
        if self.use_p2p and isinstance(self.local, _CudaLocal):
            peer = self._peer_pointer(partner)
            if half == 1:
                lo, cnt = 0, (1 if self.rank < partner else 0)
            else:
                lo, cnt = (0 if self.rank < partner else half // 2), half // 2     # each partner moves half the pairs
            self.st.peer_swap(peer, lbit, self._my_bit(gpos), lo, cnt)
        else:
            top = self.n_local - 1
            if lbit != top:
                self._local_transpose(lbit, top)
            self._exchange_sendrecv(partner, half, gpos)
            if lbit != top:
                self._local_transpose(lbit, top)
        self.local.after_comm()
        dist.barrier(group=self.group)
        self.exchange_ms += (time.perf_counter() - t0) * 1e3
        self.exchange_bytes += 16 * half
        self.n_exchanges += 1
        self.n_remaps += 1

    def exchange_many(self, pairs: Sequence[Tuple[int, int]]) -> None:
        """Swap k local index bits with k global bits in one step, pairs = [(local bit, global bit), ...] all distinct:
        the amplitude at (shard G, local bits L) moves to (shard L, local bits G).  Each rank trades one block of
        2^(n_local - k) amplitudes with each of the 2^k - 1 ranks that differ from it in a subset of those global bits,
        so (1 - 2^-k) of the shard leaves -- k single exchanges would move k / 2 of it, behind 2k barriers instead of 2.
        GPUs: one k_peer_swap_bits launch per partner, in place over CUDA-IPC peer memory.  Host stand-in (gloo tests):
        the same blocks by send/recv.  NCCL without peer access: the single-bit exchanges one after the other."""
        torch, dist = _torch()
        import time
        pairs = [(int(l), int(g)) for l, g in pairs]
        if len(pairs) == 1:
            return self.exchange(*pairs[0])
        p2p = self.use_p2p and isinstance(self.local, _CudaLocal)
        if not p2p and not isinstance(self.local, _HostLocal):
            for l, g in pairs:
                self.exchange(l, g)
            return
        pairs.sort()
        k = len(pairs)
        lbits, gposs = [l for l, _ in pairs], [g for _, g in pairs]
        assert len(set(lbits)) == k and len(set(gposs)) == k and all(0 <= l < self.n_local <= g < self.n_qubits for l, g in pairs)
        mine_g = [self._my_bit(g) for g in gposs]

        def pattern(bits):
            return sum(b << l for b, l in zip(bits, lbits))
        block = 1 << (self.n_local - k)
        self.local.before_comm()
        dist.barrier(group=self.group)                          # every shard is quiescent
        t0 = time.perf_counter()
        for d in range(1, 1 << k):
            diff = [(d >> j) & 1 for j in range(k)]
            partner = self.rank ^ sum(diff[j] << (gposs[j] - self.n_local) for j in range(k))
            mine_pat, peer_pat = pattern([a ^ b for a, b in zip(mine_g, diff)]), pattern(mine_g)
            if p2p:
                lo, cnt = (0, block // 2) if self.rank < partner else (block // 2, block - block // 2)
                self.st.peer_swap_bits(self._peer_pointer(partner, any_partner=True), lbits, mine_pat, peer_pat, lo, cnt)
            else:
                psi = self.local.state.psi
                idx = np.arange(block, dtype=np.int64)
                for l in lbits:                                   # insert a zero at every exchanged bit (ascending)
                    idx = ((idx >> l) << (l + 1)) | (idx & ((1 << l) - 1))
                idx |= mine_pat
                out = torch.from_numpy(psi[idx].view(np.float64).copy())
                inc = torch.empty_like(out)
                ops = [dist.P2POp(dist.isend, out, partner, self.group), dist.P2POp(dist.irecv, inc, partner, self.group)]
                if self.rank > partner:
                    ops.reverse()
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
                psi[idx] = inc.numpy().view(np.complex128)
        self.local.before_comm()                                # the swaps were queued on the state's stream
        self.local.after_comm()
        dist.barrier(group=self.group)
        self.exchange_ms += (time.perf_counter() - t0) * 1e3
        self.exchange_bytes += 16 * block * ((1 << k) - 1)
        self.n_exchanges += k
        self.n_remaps += 1

    def _local_transpose(self, a: int, b: int) -> None:
        prog = Program()
        woff = prog.add_words([min(a, b) | (max(a, b) << 8)])
        prog.add_op(_lib.OP_PERMUTE, woff=woff, wlen=1)
        self.st.apply(prog)
        self.local.before_comm()

    def _peer_pointer(self, partner: int, any_partner: bool = False) -> int:
        """The partner's shard mapped into this process (CUDA IPC).  The handles are published once, collectively (every
        rank reaches its first exchange at the same point of the program) and every other shard of the box is mapped:
        a remap of several global wires pairs this rank with ranks that differ in more than one bit."""
        torch, dist = _torch()
        if not self.local.peers:
            handles = [None] * self.world
            dist.all_gather_object(handles, self.st.ipc_export(), group=self.group)
            for r, h in enumerate(handles):
                if r != self.rank:
                    self.local.peers[r] = self.st.ipc_open(h)
        return self.local.peers[partner]

    def _exchange_sendrecv(self, partner: int, half: int, gpos: int) -> None:
        torch, dist = _torch()
        start = half if self._my_bit(gpos) == 0 else 0
        buf = self.local.buffer()                               # float64 view, 2 doubles per amplitude
        chunk = min(half, self.BOUNCE_AMPS)
        if self._bounce is None or self._bounce.numel() < 2 * chunk:
            self._bounce = torch.empty(2 * chunk, dtype=torch.float64, device=buf.device)
        for off in range(0, half, chunk):
            region = buf[2 * (start + off): 2 * (start + off + chunk)]
            tmp = self._bounce[:2 * chunk]
            ops = [dist.P2POp(dist.isend, region, partner, self.group), dist.P2POp(dist.irecv, tmp, partner, self.group)]
            if self.rank > partner:
                ops.reverse()
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            region.copy_(tmp)

    def _make_local(self, wires: Sequence[int]) -> None:
        """Bring every wire in `wires` onto a local index bit (used by X / Y expectation values)."""
        need = [w for w in wires if self.layout[w] >= self.n_local]
        keep = set(self.layout[w] for w in wires)
        if len(set(wires)) > self.n_local:
            raise ValueError(f"a Pauli term with X/Y on {len(set(wires))} qubits does not fit the {self.n_local} local qubits "
                             f"of a shard")
        for w in need:
            inv = self._wire_at()
            victim = max(p for p in range(self.n_local) if p not in keep)
            gpos = self.layout[w]
            self.exchange(victim, gpos)
            self.layout[w], self.layout[inv[victim]] = victim, gpos
            keep.add(victim)

    # -- measurement ------------------------------------------------------------------------------
    def norm2(self) -> float:
        return float(self._allreduce(np.array([self.st.norm2()]))[0])

    def prob_zero(self, target: int) -> float:
        p = self.layout[target]
        if p < self.n_local:
            mine = self.st.prob_zero(p)
        else:
            mine = self.st.norm2() if self._my_bit(p) == 0 else 0.0
        return float(self._allreduce(np.array([mine]))[0])

    def collapse(self, target: int, outcome: int, p: float, reset: bool = False) -> None:
        """Never changes the layout (the programs that follow were planned for it)."""
        torch, dist = _torch()
        pb = self.layout[target]
        if pb < self.n_local:
            self.st.collapse(pb, outcome, p, reset=reset)
            return
        keep = self._my_bit(pb) == outcome
        self.st.scale(1.0 / math.sqrt(p) if keep else 0.0)
        if reset and outcome == 1:
            # move the surviving |1> shards onto their |0> partners, whole shards at a time
            partner = self.rank ^ (1 << (pb - self.n_local))
            buf = self.local.buffer()
            self.local.before_comm()
            chunk = 2 * self.BOUNCE_AMPS
            for off in range(0, buf.numel(), chunk):
                part = buf[off:off + chunk]
                if keep:
                    dist.send(part, partner, group=self.group)
                else:
                    dist.recv(part, partner, group=self.group)
            self.local.after_comm()
            if keep:
                self.st.scale(0.0)

    def scale(self, z: complex) -> None:
        self.st.scale(z)

    def first_above(self, eps: float) -> Optional[Tuple[int, complex]]:
        """(wire-order index, amplitude) of the first amplitude above eps, without gathering the register: every shard
        searches its own amplitudes for the smallest wire-order index (b200_first_above_wire), the minimum over the
        ranks wins and its owner publishes the amplitude."""
        key = self.st.first_above_wire(eps, self._wire_at())
        none = float(1 << 62)
        keys = self._allgather_scalar(float(key) if key is not None else none)      # exact: keys are below 2^53
        best = min(keys)
        if best >= none:
            return None
        best = int(best)
        phys = self._phys_index(best)
        val = np.zeros(2)
        if phys >> self.n_local == self.rank:
            a = self.st.download(phys & ((1 << self.n_local) - 1), 1)[0]
            val[:] = (a.real, a.imag)
        val = self._allreduce(val)
        return best, complex(val[0], val[1])

    def _allgather_scalar(self, x: float) -> List[float]:
        v = np.zeros(self.world)
        v[self.rank] = x
        return [float(t) for t in self._allreduce(v)]

    def sample(self, order: Sequence[int], uniforms: np.ndarray) -> np.ndarray:
        """The descent of csrc/sample.cu with the per-level sums added over the shards."""
        order = [int(q) for q in order]
        m = len(order)
        uniforms = np.asarray(uniforms, dtype=np.float64).reshape(-1, max(1, m))
        n_shots = uniforms.shape[0]
        out = np.zeros(n_shots, dtype=np.uint64)
        prefix = np.zeros(n_shots, dtype=np.uint64)           # outcomes so far, as physical index bits
        fixed_mask = 0
        lmask = (1 << self.n_local) - 1
        mine = self.rank << self.n_local
        # first K levels from one reduction (see csrc/sample.cu): marginal table of K distinct qubits.  When those
        # qubits sit exactly on the local index bits 0..K-1 (K up to 13) the table is one CONTIGUOUS read of every shard
        # (b200_marginal_low) instead of 2^(K-1) strided groups -- the strided form ran at a quarter of the bandwidth
        # and made up most of the 0.4 s the 36-qubit descent took.
        j0 = 0
        K = 0
        if hasattr(self.st, "marginal_low"):
            lo, hi = self.st.MARGINAL_LOW_BITS
            seen = 0
            for t in range(min(m, hi, self.n_local - 2)):
                b = self.layout[order[t]]
                if b >= self.n_local or (seen >> b) & 1:
                    break
                seen |= 1 << b
                if seen == (1 << (t + 1)) - 1:
                    K = t + 1
            if K < max(lo, 11):
                K = 0
        if K:
            bits = [self.layout[w] for w in order[:K]]
            table = self._allreduce(self.st.marginal_low(K))
            v = np.zeros(1 << K, dtype=np.int64)              # bit t of p = outcome of order[t] = index bit bits[t]
            pidx = np.arange(1 << K, dtype=np.int64)
            for t in range(K):
                v |= ((pidx >> t) & 1) << bits[t]
            node = [None] * (K + 1)
            node[K] = table[v]
            for j in range(K - 1, -1, -1):
                node[j] = node[j + 1][: 1 << j] + node[j + 1][1 << j:]
            p = np.zeros(n_shots, dtype=np.int64)
            for j in range(K):
                s0, s1 = node[j + 1][p], node[j + 1][p | (1 << j)]
                bit = ~(uniforms[:, j] < s0 / (s0 + s1))
                p |= bit.astype(np.int64) << j
                out |= bit.astype(np.uint64) << np.uint64(j)
                prefix |= np.where(bit, np.uint64(1 << bits[j]), np.uint64(0))
            fixed_mask = (1 << K) - 1
            j0 = K
        K0 = K
        K = 0
        while not K0 and K < m and K < 10 and order[K] not in order[:K]:
            K += 1
        if K >= 2:
            bits = [self.layout[w] for w in order[:K]]
            mask0 = sum(1 << b for b in bits[:-1])
            n_groups = 1 << (K - 1)
            gb = np.zeros(n_groups, dtype=np.uint64)
            for t in range(K - 1):
                gb |= ((np.arange(n_groups, dtype=np.uint64) >> np.uint64(t)) & np.uint64(1)) << np.uint64(bits[t])
            sums = np.zeros((n_groups, 2))
            gmask = mask0 & ~lmask
            here = np.nonzero((gb & np.uint64(~lmask & ((1 << self.n_qubits) - 1))) == np.uint64(mine & gmask))[0]
            if here.size:
                local_bits = gb[here] & np.uint64(lmask)
                pb = bits[-1]
                if pb < self.n_local:
                    sums[here] = self.st.cond_probs(mask0 & lmask, pb, local_bits)
                else:
                    sums[here, self._my_bit(pb)] = self.st.cond_probs(mask0 & lmask, -1, local_bits)[:, 0]
            sums = self._allreduce(sums.reshape(-1)).reshape(-1, 2)
            node = [None] * (K + 1)
            node[K] = np.concatenate([sums[:, 0], sums[:, 1]])
            for j in range(K - 1, -1, -1):
                node[j] = node[j + 1][: 1 << j] + node[j + 1][1 << j:]
            p = np.zeros(n_shots, dtype=np.int64)
            for j in range(K):
                s0, s1 = node[j + 1][p], node[j + 1][p | (1 << j)]
                bit = ~(uniforms[:, j] < s0 / (s0 + s1))
                p |= bit.astype(np.int64) << j
                out |= bit.astype(np.uint64) << np.uint64(j)
                prefix |= np.where(bit, np.uint64(1 << bits[j]), np.uint64(0))
            fixed_mask = mask0 | (1 << bits[-1])
            j0 = K
        for j, wire in enumerate(order):
            if j < j0:
                continue
            pb = self.layout[wire]
            tbit = 1 << pb
            if fixed_mask & tbit:
                ones = (prefix & np.uint64(tbit)) != 0
                p0 = np.where(ones, 0.0, 1.0)
                bit = ~(uniforms[:, j] < p0)
                out |= bit.astype(np.uint64) << np.uint64(j)
                continue
            groups, inverse = np.unique(prefix, return_inverse=True)
            sums = np.zeros((groups.size, 2))
            gmask = fixed_mask & ~lmask
            here = [i for i, gb in enumerate(groups) if (int(gb) & ~lmask) == (mine & gmask)]
            if here:
                local_bits = np.array([int(groups[i]) & lmask for i in here], dtype=np.uint64)
                if pb < self.n_local:
                    part = self.st.cond_probs(fixed_mask & lmask, pb, local_bits)
                    sums[here] = part
                else:
                    part = self.st.cond_probs(fixed_mask & lmask, -1, local_bits)
                    sums[here, self._my_bit(pb)] = part[:, 0]
            sums = self._allreduce(sums.reshape(-1)).reshape(-1, 2)
            p0 = sums[inverse, 0] / (sums[inverse, 0] + sums[inverse, 1])
            bit = ~(uniforms[:, j] < p0)
            out |= bit.astype(np.uint64) << np.uint64(j)
            prefix |= np.where(bit, np.uint64(tbit), np.uint64(0))
            fixed_mask |= tbit
        return out

    def marginal(self, wires: Sequence[int]) -> np.ndarray:
        """Probabilities of the 2^k outcomes of the distinct `wires` (bit j of the index = wires[j]): each rank
        reduces the outcomes compatible with its shard number, the partial tables are added over the ranks."""
        bits = [self.layout[int(w)] for w in wires]
        lmask = (1 << self.n_local) - 1
        idx = np.arange(1 << len(bits), dtype=np.uint64)
        gb = np.zeros_like(idx)
        for j, b in enumerate(bits):
            gb |= ((idx >> np.uint64(j)) & np.uint64(1)) << np.uint64(b)
        mask = sum(1 << b for b in bits)
        mine = (self.rank << self.n_local) & mask
        here = np.nonzero((gb & np.uint64(mask & ~lmask)) == np.uint64(mine))[0]
        table = np.zeros(idx.size)
        if here.size:
            table[here] = self.st.cond_probs(mask & lmask, -1, gb[here] & np.uint64(lmask))[:, 0]
        return self._allreduce(table)

    def expect_group(self, xmask: int, zymasks: Sequence[int], coeffs: Sequence[complex]) -> complex:
        wires = [w for w in range(self.n_qubits) if (xmask >> w) & 1]
        self._make_local(wires)                                 # X / Y pair amplitudes inside one shard only
        def phys(mask: int) -> int:
            return self._phys_index(int(mask))
        val = self.st.expect_group(phys(xmask), [phys(z) for z in zymasks], coeffs)
        tot = self._allreduce(np.array([val.real, val.imag]))
        return complex(tot[0], tot[1])

    # -- bookkeeping ------------------------------------------------------------------------------
    def launch_count(self) -> int:
        return self.st.launch_count()

    def last_program_ms(self) -> float:
        return self.st.last_program_ms()

    def sync(self) -> None:
        self.st.sync()
