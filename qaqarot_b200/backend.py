"""``B200Backend`` -- the drop-in for blueqat's ``NumPyBackend`` on one B200.

Same plugin interface and keyword arguments as the reference backend
(``run(gates, n_qubits, shots=None, initial=None, returns=None, ignore_global=False, save_cache=False)``,
blueqat/backends/numpy_backend.py:572-630; ``statevector / shots / oneshot / samples / make_cache``, :632-685;
``copy``, backendbase.py:27-33), plus ``hamiltonian=`` following the convention of the reference's only
consumer of that keyword (blueqat/backends/quimb.py:17-72: a float, taking precedence over shots).

The host side only plans: operations are lowered to primitives (lowering.py), grouped into fused passes
(planner.py) and handed to the CUDA library through the C ABI (device.py -> include/b200sv.h).  Randomness
comes from Python's global ``random.random()`` in the reference's order -- one draw per measured / reset
qubit per shot (numpy_backend.py:456, 488) -- so seeded runs reproduce NumPyBackend's Counters.

There is no CPU path: constructing the first device state raises ``B200LibraryError`` / ``B200Error`` when
the library or a GPU is missing.
"""
from __future__ import annotations

import math
import random
import warnings
from collections import Counter
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .lowering import Lowered, Prim, lower
from .pauli import canonical_terms
from . import param_replay as R
from .program import Program, unfused_program

try:                                                   # subclass blueqat's base class when it is importable
    from blueqat.backends.backendbase import Backend as _Base
except Exception:                                      # blueqat absent (e.g. on the GPU box)
    class _Base:                                       # minimal stand-in for backendbase.Backend
        def copy(self):
            raise NotImplementedError

        def make_cache(self, gates, n_qubits):
            return None

DEFAULT_SHOTS = 1024                                   # numpy_backend.py:33
RETURNS = ("statevector", "shots", "statevector_and_shots", "_inner_ctx", "samples")


def _default_state_factory(n_qubits: int, device: int):
    from .device import DeviceState
    return DeviceState(n_qubits, device)


class RunContext:
    """What ``returns="_inner_ctx"`` hands back (backend-private in the reference, :521)."""

    def __init__(self, n_qubits):
        self.n_qubits = n_qubits
        self.qubits: Optional[np.ndarray] = None
        self.cregs: List[int] = [0] * n_qubits
        self.shots_result: Counter = Counter()
        self.samples: List[Dict[str, List[int]]] = []
        self.device_state = None


class _Compiled:
    """Lowered + planned form of gates[start:], cached on the backend between runs of the same circuit."""

    def __init__(self, gates: Sequence, n_qubits: int, start: int, planner: Callable[..., Program],
                 initial_layout: Optional[Sequence[int]] = None, from_zero: bool = False,
                 on_pass: Optional[Callable[[Program], None]] = None):
        self.gates = tuple(gates)
        self.n_qubits = n_qubits
        self.start = start
        self.planner = planner
        self.from_zero = from_zero          # the prefix starts from |0...0>: its initial wire layout is free
        self.initial_layout = list(initial_layout) if initial_layout is not None else None
        low = lower(self.gates[start:], n_qubits, start)
        self.n_gates = low.n_gates
        prims = low.prims
        user_planner = planner
        layout = [self.initial_layout]

        first = [True]

        def planner(seg, n):            # a sharded planner leaves wires where they are: thread the layout along
            free = from_zero and first[0] and getattr(user_planner, "accepts_free_layout", False)
            kw = {"free_initial_layout": True} if free else {}
            if first[0] and on_pass is not None and getattr(user_planner, "accepts_on_pass", False):
                kw["on_pass"] = on_pass          # the prefix is launched pass by pass while it is being planned
            first[0] = False
            if getattr(user_planner, "threads_layout", False):
                prog = user_planner(seg, n, layout[0], **kw)
                layout[0] = list(prog.final_pos)
                return prog
            return user_planner(seg, n, **kw)

        k = low.first_nonunitary if low.first_nonunitary is not None else len(prims)
        self.prefix_prims = prims[:k]
        self.prefix = planner(self.prefix_prims, n_qubits)
        self.first_nonunitary_op = prims[k].op_index if k < len(prims) else None
        # the rest: alternating unitary segments (as programs) and measure/reset primitives
        self.rest: List[Any] = []
        seg: List[Prim] = []
        for p in prims[k:]:
            if p.unitary:
                seg.append(p)
            else:
                if seg:
                    self.rest.append(planner(seg, n_qubits))
                    seg = []
                self.rest.append(p)
        if seg:
            self.rest.append(planner(seg, n_qubits))
        self.terminal = bool(self.rest) and all(isinstance(r, Prim) and r.kind == "measure" for r in self.rest)

    def matches(self, gates: Sequence, n_qubits: int, start: int, initial_layout=None, from_zero=False) -> bool:
        """The cached plan serves a gate list that is the same BY VALUE (operation kind, targets, parameters, options):
        an ansatz rebuilt with the angles it already had (vqe.py:67-83 builds a fresh circuit per call, once for the
        sampler and again for the energies) does not re-plan.  New angles do: the plan's tables are functions of them."""
        return (n_qubits == self.n_qubits and start == self.start and len(gates) == len(self.gates)
                and from_zero == self.from_zero
                and (list(initial_layout) if initial_layout is not None else None) == self.initial_layout
                and all(a is b or _same_op(a, b) for a, b in zip(gates, self.gates)))


class _Streamer:
    """Launches the ops of a program as the planner appends them (planner.plan's on_pass), without waiting for the
    device: with a new angle vector every iteration (vqe.py:67-83) the plan cannot be reused, but the device can run pass
    k while the host encodes pass k + 1."""

    def __init__(self, st, init_basis: Optional[int]):
        self.st, self.init_basis = st, init_basis
        self.ops = self.words = self.reals = 0
        self.first = True

    def __call__(self, prog: Program) -> None:
        n = len(prog._ops)
        if n == self.ops:
            return
        self.st.apply_piece(prog, self.ops, n, self.words, self.reals, init_basis=self.init_basis if self.first else None,
                            first=self.first)
        self.ops, self.words, self.reals, self.first = n, len(prog._words), len(prog._reals), False

    def finish(self, prog: Program) -> None:
        self(prog)
        if self.first and self.init_basis is not None:       # an empty program: only the preparation is left
            self.st.set_basis(self.init_basis & (self.st.dim - 1))


def _same_op(a, b) -> bool:
    if type(a) is not type(b) or getattr(a, "lowername", None) != getattr(b, "lowername", None):
        return False
    try:
        if a.targets != b.targets or len(a.params) != len(b.params):
            return False
        if not all(type(x) is type(y) and bool(np.all(x == y)) for x, y in zip(a.params, b.params)):
            return False
        return (getattr(a, "options", None) == getattr(b, "options", None)
                and getattr(a, "key", None) == getattr(b, "key", None)
                and getattr(a, "duplicated", None) == getattr(b, "duplicated", None))
    except Exception:
        return False


def _check_initial(initial, n_qubits: int) -> np.ndarray:
    """numpy_backend.py:500-512."""
    if not isinstance(initial, np.ndarray):
        raise ValueError(f"`initial` must be a np.ndarray, but {type(initial)}")
    if initial.shape != (2 ** n_qubits,):
        raise ValueError(f"`initial.shape` is not matched. Expected: {(2**n_qubits,)}, Actual: {initial.shape}")
    if initial.dtype != complex:
        initial = initial.astype(complex)
    return initial


class B200Backend(_Base):
    """Statevector simulator backend running on one NVIDIA B200 through libb200sv.so."""

    def __init__(self, device: int = 0, fusion: bool = True, state_factory=None, planner=None, plan_config=None):
        self.device = device
        self.fusion = fusion
        self.plan_config = plan_config
        self.cache = None                    # device snapshot taken before the first Measurement/Reset
        self.cache_idx = -1
        self._state_factory = state_factory or _default_state_factory
        if planner is None:
            if fusion:
                from .planner import DEFAULT, plan as _plan
                _cfg = plan_config or DEFAULT

                def planner(prims, n, free_initial_layout=False, on_pass=None):
                    return _plan(prims, n, _cfg, free_initial_layout=free_initial_layout, on_pass=on_pass)
                planner.accepts_free_layout = True
                planner.accepts_on_pass = True
                import dataclasses
                _cfg_exact = dataclasses.replace(_cfg, exact_rotations=True)

                def planner_exact(prims, n, free_initial_layout=False, on_pass=None):
                    return _plan(prims, n, _cfg_exact, free_initial_layout=free_initial_layout, on_pass=on_pass)
                planner_exact.accepts_free_layout = True
                planner_exact.accepts_on_pass = True
                planner_exact.min_qubits = _cfg.min_qubits
                self._planner_exact = planner_exact
            else:
                planner = lambda prims, n: unfused_program(prims, n)
        self._planner = planner
        if not hasattr(self, "_planner_exact"):
            self._planner_exact = None
        self.stream_plans = True             # launch a newly planned circuit pass by pass (b200_apply_program_ex)
        # Objective loops (param_replay): the same gate structure with new angles is launched at once with the payload a
        # model of the planner predicts, planned for real while the device runs, and compared.
        self.param_replay = True
        self._models: Dict[Any, Any] = {}    # structure key -> ParamModel, or None when the structure is not modelable
        self._last_key = None
        self._last_angles = None
        self._repeats = 0
        self.replay_stats = {"built": 0, "hits": 0, "misses": 0, "unmodelable": 0}
        self._work = None                    # reusable work state
        self._compiled: Optional[_Compiled] = None
        self.last_stats: Dict[str, Any] = {}
        # optional caller-provided host destination for returned statevectors (e.g. a page-locked buffer,
        # device.PinnedBuffer): like the reference, which returns its live context array, the result is then
        # only valid until the next run on this backend.
        self.result_buffer: Optional[np.ndarray] = None

    # -- Backend protocol -------------------------------------------------------------------------
    def copy(self) -> "B200Backend":
        """Device buffers cannot be deep-copied; clone the snapshot explicitly (backendbase.py:27-33)."""
        other = B200Backend(self.device, self.fusion, self._state_factory, self._planner)
        if self.cache is not None:
            other.cache = self.cache.clone()
            other.cache_idx = self.cache_idx
        return other

    def __deepcopy__(self, memo):
        return self.copy()

    def close(self) -> None:
        for st in (self._work, self.cache):
            if st is not None:
                st.close()
        self._work = self.cache = None
        self.cache_idx = -1

    def _clear_cache(self) -> None:
        if self.cache is not None:
            self.cache.close()
        self.cache, self.cache_idx = None, -1

    def _state(self, n_qubits: int):
        if self._work is None or self._work.n_qubits != n_qubits:
            if self._work is not None:
                self._work.close()
            self._work = self._state_factory(n_qubits, self.device)
        return self._work

    def _compile(self, gates: Sequence, n_qubits: int, start: int, initial_layout=None, from_zero=False,
                 on_pass=None, planner=None) -> _Compiled:
        """`on_pass`: only used when the circuit has to be planned (no cached plan); the caller checks
        `on_pass.ops` to see whether the prefix has already been launched."""
        c = self._compiled
        planner = planner or self._planner
        if c is None or c.planner is not planner or not c.matches(gates, n_qubits, start, initial_layout, from_zero):
            c = self._compiled = _Compiled(gates, n_qubits, start, planner, initial_layout, from_zero, on_pass)
        return c

    # -- objective loops: speculative launch from a model of the planner (param_replay.py) -------------------
    def _replay_model(self, gates, n_qubits: int):
        """The ParamModel of this gate structure, built when the structure shows up for the third time in a row with
        other angles; None when there is none (yet) or the structure is not modelable."""
        key = R.structure_key(gates, n_qubits)
        angles = R.angle_values(gates)
        last_key, last_angles = self._last_key, self._last_angles
        self._last_key, self._last_angles = key, angles
        generation = 0
        if key in self._models:
            model = self._models[key]
            # a model describes the planner around the angles it was built at (every rotation within a quarter turn of
            # there): a loop that has walked out of that range, and stays out, gets a new one around where it is now
            if (model is None or model.declined < R.RECENTRE_AFTER or model.generation >= R.MAX_RECENTRES
                    or key != last_key or np.array_equal(angles, last_angles)):
                return model, angles
            generation = model.generation + 1
        elif key != last_key or angles.size == 0 or np.array_equal(angles, last_angles):
            self._repeats = 0
            return None, angles
        else:
            # the model costs one plan per distinct angle: only a loop pays that back, so wait for the third call in a row
            self._repeats += 1
            if self._repeats < 2:
                return None, angles
        planner = self._planner_exact

        def plan_prefix(g):
            low = lower(tuple(g), n_qubits, 0)
            k = low.first_nonunitary if low.first_nonunitary is not None else len(low.prims)
            return planner(low.prims[:k], n_qubits, free_initial_layout=True)
        try:
            model = R.ParamModel(list(gates), last_angles, plan_prefix)
            model.generation = generation
            self.replay_stats["built"] += 1
        except Exception:                                    # NotModelable, or a probe the planner itself rejects: the
            model = None                                     # loop simply goes on without a model for this structure
            self.replay_stats["unmodelable"] += 1
        if len(self._models) > 16:
            self._models.clear()
        self._models[key] = model
        return model, angles

    def _download(self, st) -> np.ndarray:
        buf = self.result_buffer
        if buf is not None and buf.size == st.dim and buf.dtype == np.complex128 and buf.flags.c_contiguous:
            return st.download(out=buf)
        return st.download()

    # -- the run loop -----------------------------------------------------------------------------
    def _execute(self, gates, n_qubits, shots: int, initial, save_cache: bool, want_final_state: bool,
                 use_backend_cache: bool = True) -> RunContext:
        """Counterpart of the shot loop + _run_inner (numpy_backend.py:545-570, 618-626)."""
        ctx = RunContext(n_qubits)
        cache, cache_idx = (self.cache, self.cache_idx) if use_backend_cache else (None, -1)
        start = cache_idx + 1 if cache is not None else 0
        st = self._state(n_qubits)
        ctx.device_state = st
        if cache is not None:
            st.copy_from(cache)
        elif initial is not None:
            st.upload(initial)
        # |0...0> is synthesised by the first pass (B200_OP_INIT)
        init_basis = 0 if cache is None and initial is None else None
        # a circuit that has to be planned is launched pass by pass while the planner works (see _Streamer); a cached
        # plan (the same gates by value) goes to the device in one call
        streamer = _Streamer(st, init_basis) if self.stream_plans and hasattr(st, "apply_piece") else None
        layout = getattr(cache, "layout", None) if cache is not None else None
        from_zero = cache is None and initial is None
        planner, model, spec = None, None, None
        c = self._compiled
        if (self.param_replay and from_zero and start == 0 and streamer is not None and self._planner_exact is not None
                and n_qubits >= self._planner_exact.min_qubits
                and not (c is not None and c.matches(gates, n_qubits, start, layout, from_zero))):
            model, angles = self._replay_model(gates, n_qubits)
            if model is not None:
                planner = self._planner_exact            # (the model describes plans with exact rotations)
                spec = model.evaluate(angles)
                model.declined = 0 if spec is not None else model.declined + 1
                if spec is not None:
                    # launch with the predicted payload now; the real plan is made while the device runs
                    prog = model.program(spec)
                    st.apply_piece(prog, 0, len(prog._ops), 0, 0, init_basis=init_basis, first=True)
                    streamer = None
        comp = self._compile(gates, n_qubits, start, layout, from_zero=from_zero, on_pass=streamer, planner=planner)
        if spec is not None:
            real = comp.prefix
            if (model.matches(real) and len(real._reals) == spec.size
                    and float(np.abs(np.asarray(real._reals) - spec).max()) <= R.TOL):
                self.replay_stats["hits"] += 1
            else:
                # a branch of the planner went the other way for these angles: run the real plan, forget the model
                self.replay_stats["misses"] += 1
                self._models[self._last_key] = None
                st.apply(real, init_basis=init_basis)
        elif streamer is not None and (streamer.ops > 0 or streamer.words > 0):
            streamer.finish(comp.prefix)
        else:
            st.apply(comp.prefix, init_basis=init_basis)
        self.last_stats = {"n_gates": comp.n_gates, "prefix_passes": comp.prefix.n_passes,
                           "prefix_prims": comp.prefix.n_prims}
        if not comp.rest:
            # no measurement or reset: every shot leaves cregs = 0...0 (numpy_backend.py:622-623)
            if n_qubits > 0:
                ctx.shots_result["0" * n_qubits] = shots
            ctx.samples = [{} for _ in range(shots)]
            return ctx

        # a snapshot is only needed to re-run measurements shot by shot (or to keep as the cache): terminal
        # measurements are sampled without collapsing, and a second copy of a large register may not even fit
        snap = None
        if (shots > 1 and not comp.terminal) or save_cache:
            snap = st.clone()
        if save_cache:
            self._clear_cache()
            self.cache, self.cache_idx = snap, comp.first_nonunitary_op - 1

        if comp.terminal:
            self._sample_terminal(ctx, comp, st, shots, want_final_state)
        else:
            for s in range(shots):
                if s > 0:
                    st.copy_from(snap)
                self._trajectory(ctx, comp, st)
        if snap is not None and not save_cache:
            snap.close()
        return ctx

    def _trajectory(self, ctx: RunContext, comp: _Compiled, st) -> None:
        """One shot through measure/reset one qubit at a time (gate_measure / gate_reset, :447-497)."""
        cregs = [0] * ctx.n_qubits
        sample: Dict[str, List[int]] = {}
        pending: Optional[Tuple[Any, List[int]]] = None

        def flush():
            nonlocal pending
            if pending is not None:
                _store_key(sample, pending[0], pending[1])
                pending = None

        for item in comp.rest:
            if isinstance(item, Program):
                flush()
                st.apply(item)
                continue
            if item.kind == "measure":
                if pending is None or pending[0][2] != item.group[2]:
                    flush()
                    pending = (item.group, [])
                if item.target < 0:
                    continue
            else:
                flush()
            p0 = st.prob_zero(item.target)
            if random.random() < p0:
                st.collapse(item.target, 0, p0, reset=item.kind == "reset")
                bit = 0
            else:
                st.collapse(item.target, 1, 1.0 - p0, reset=item.kind == "reset")
                bit = 1
            if item.kind == "measure":
                cregs[item.target] = bit
                pending[1].append(bit)
        flush()
        ctx.cregs = cregs
        ctx.samples.append(sample)
        if cregs:
            ctx.shots_result["".join(map(str, cregs))] += 1

    def _sample_terminal(self, ctx: RunContext, comp: _Compiled, st, shots: int, want_final_state: bool) -> None:
        """All shots of a circuit ending in measurements, in one batched descent (csrc/sample.cu)."""
        meas = [p for p in comp.rest if p.target >= 0]
        order = [p.target for p in meas]
        m = len(order)
        # the reference draws shot by shot, one uniform per measured qubit in target_iter order
        uniforms = np.array([random.random() for _ in range(shots * m)], dtype=np.float64).reshape(shots, max(m, 1))
        bits_all = np.zeros(shots, dtype=object)
        for lo in range(0, m, 64):                      # the device call takes <= 64 qubits at a time
            hi = min(m, lo + 64)
            if lo == 0:
                chunk = st.sample(order[lo:hi], uniforms[:, lo:hi])
                for s in range(shots):
                    bits_all[s] = int(chunk[s])
            else:                                       # pragma: no cover - needs > 64 measured qubits
                raise ValueError("more than 64 measured qubits in a terminal measurement are not supported")
        n = ctx.n_qubits
        for s in range(shots):
            b = bits_all[s]
            cregs = [0] * n
            sample: Dict[str, List[int]] = {}
            j = 0
            groups: List[Tuple[Any, List[int]]] = []
            for p in comp.rest:
                if not groups or groups[-1][0][2] != p.group[2]:
                    groups.append((p.group, []))
                if p.target < 0:
                    continue
                bit = (b >> j) & 1
                j += 1
                cregs[p.target] = bit
                groups[-1][1].append(bit)
            for grp, vals in groups:
                _store_key(sample, grp, vals)
            ctx.samples.append(sample)
            ctx.shots_result["".join(map(str, cregs))] += 1
            ctx.cregs = cregs
        if want_final_state and shots > 0:
            b = bits_all[shots - 1]
            for j, q in enumerate(order):
                bit = (b >> j) & 1
                p0 = st.prob_zero(q)
                st.collapse(q, bit, (1.0 - p0) if bit else p0)

    # -- public entry points ----------------------------------------------------------------------
    def run(self, gates, n_qubits, shots: Optional[int] = None, initial: Optional[np.ndarray] = None,
            returns: Optional[str] = None, ignore_global: bool = False, save_cache: bool = False,
            hamiltonian=None, **kwargs) -> Any:
        if hamiltonian is not None:
            if kwargs:
                warnings.warn(f"Unknown run arguments: {kwargs}")
            return self.expectation(gates, n_qubits, hamiltonian, initial=initial)
        if returns is None:
            returns = "statevector" if shots is None else "shots"
        if returns not in RETURNS:
            raise ValueError(f"Unknown returns type '{returns}'")
        if shots is None:
            shots = 1 if returns in ("statevector", "_inner_ctx") else DEFAULT_SHOTS
        if returns == "statevector" and shots > 1:
            shots = 1
            warnings.warn("When `returns` = 'statevector', `shots` is ignored.")
        if kwargs:
            warnings.warn(f"Unknown run arguments: {kwargs}")
        if returns == "samples":
            return self.samples(gates, n_qubits, shots, initial)

        if initial is not None:
            initial = _check_initial(initial, n_qubits)
            if save_cache:
                warnings.warn("When initial is not None, saving cache is disabled.")
                save_cache = False
            self._clear_cache()
        elif self.cache is not None and self.cache.n_qubits != n_qubits:
            self._clear_cache()

        want_state = returns != "shots"
        ctx = self._execute(gates, n_qubits, shots, initial, save_cache, want_state)
        if returns == "shots":
            return ctx.shots_result
        st = ctx.device_state
        if ignore_global:
            _ignore_global_phase(st)
        ctx.qubits = self._download(st)
        if returns == "statevector":
            return ctx.qubits
        if returns == "statevector_and_shots":
            return ctx.qubits, ctx.shots_result
        return ctx

    def statevector(self, gates, n_qubits, initial: Optional[np.ndarray] = None) -> np.ndarray:
        if initial is not None:
            initial = _check_initial(initial, n_qubits)
        ctx = self._execute(gates, n_qubits, 1, initial, False, True, use_backend_cache=False)
        return ctx.device_state.download()

    def shots(self, gates, n_qubits, shots: int, initial: Optional[np.ndarray] = None):
        if initial is not None:
            initial = _check_initial(initial, n_qubits)
        return self._execute(gates, n_qubits, shots, initial, False, False, use_backend_cache=False).shots_result

    def oneshot(self, gates, n_qubits, initial: Optional[np.ndarray] = None):
        if initial is not None:
            initial = _check_initial(initial, n_qubits)
        ctx = self._execute(gates, n_qubits, 1, initial, False, True, use_backend_cache=False)
        return ctx.device_state.download(), ctx.shots_result.most_common()[0][0]

    def samples(self, gates, n_qubits, shots: int, initial: Optional[np.ndarray] = None):
        """Experimental in the reference (numpy_backend.py:669-682)."""
        if initial is not None:
            initial = _check_initial(initial, n_qubits)
        return self._execute(gates, n_qubits, shots, initial, False, False, use_backend_cache=False).samples

    def make_cache(self, gates, n_qubits) -> None:
        self.run(gates, n_qubits, save_cache=True)

    def expectation(self, gates, n_qubits, hamiltonian, initial: Optional[np.ndarray] = None) -> float:
        """<psi| H |psi> for a Pauli expression, evaluated on the device without copying the state."""
        terms = canonical_terms(hamiltonian)
        for _, letters in terms:
            for _, q in letters:
                if not 0 <= q < n_qubits:
                    raise ValueError(f"hamiltonian acts on qubit {q}, circuit has {n_qubits} qubits")
        if initial is not None:
            initial = _check_initial(initial, n_qubits)
        ctx = self._execute(gates, n_qubits, 1, initial, False, True, use_backend_cache=initial is None)
        return expectation_on_state(ctx.device_state, terms)


    def marginal(self, gates, n_qubits, meas, initial: Optional[np.ndarray] = None) -> Dict[Tuple[int, ...], float]:
        """{outcome of the qubits `meas`: probability} of the final state, reduced on the device (what
        ``vqe.expect(circuit.run(), meas)`` computes on the host, vqe.py:231-259; see samplers.py)."""
        from .samplers import expect_on_state
        for q in meas:
            if not 0 <= int(q) < n_qubits:
                raise ValueError(f"measured qubit {q} outside the {n_qubits}-qubit register")
        if initial is not None:
            initial = _check_initial(initial, n_qubits)
        ctx = self._execute(gates, n_qubits, 1, initial, False, True, use_backend_cache=initial is None)
        return expect_on_state(ctx.device_state, meas)


def expectation_on_state(st, terms) -> float:
    groups: Dict[int, Tuple[List[int], List[complex]]] = {}
    for coeff, letters in terms:
        xmask = zy = ny = 0
        for letter, q in letters:
            if letter in "XY":
                xmask |= 1 << q
            if letter in "YZ":
                zy |= 1 << q
            ny += letter == "Y"
        g = groups.setdefault(xmask, ([], []))
        g[0].append(zy)
        g[1].append(complex(coeff) * (1j) ** ny)
    total = 0.0 + 0.0j
    for xmask in sorted(groups):
        zys, coeffs = groups[xmask]
        total += st.expect_group(xmask, zys, coeffs)
    return float(total.real)


def _store_key(sample: Dict[str, List[int]], group, measured: List[int]) -> None:
    """Measurement-key bookkeeping of gate_measure (numpy_backend.py:467-476)."""
    key, duplicated, _ = group
    if key is None:
        return
    if key in sample:
        if duplicated == "replace":
            sample[key] = measured
        elif duplicated == "append":
            sample[key] = sample[key] + measured
        else:
            raise ValueError("Measurement key is duplicated.")
    else:
        sample[key] = measured


def _ignore_global_phase(st) -> None:
    """utils.ignore_global_phase (utils.py:130-144) on the device-resident state."""
    hit = st.first_above(0.0000001)
    if hit is not None:
        q = hit[1]
        st.scale(abs(q) / q)


class ShardedB200Backend(B200Backend):
    """The same backend with the register sharded over the ranks of a torch.distributed process group (one
    process per GPU, launched with torchrun).  Every rank runs the same circuit and gets the same results:
    statevectors are gathered (small registers), Counters and expectation values are identical everywhere --
    seed Python's ``random`` identically on all ranks for shots.  See qaqarot_b200/sharded.py."""

    def __init__(self, device: Optional[int] = None, plan_config=None, group=None, host_state_factory=None):
        import os
        import torch.distributed as dist
        from .planner import DEFAULT, plan
        from .sharded import ShardedState
        if not dist.is_initialized():
            raise RuntimeError("ShardedB200Backend needs torch.distributed.init_process_group() first")
        world = dist.get_world_size(group)
        g = world.bit_length() - 1
        cfg = plan_config or DEFAULT
        if device is None and host_state_factory is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))

        def planner(prims, n, layout=None, free_initial_layout=False):
            return plan(prims, n, cfg, n_local=n - g, initial_pos=layout, free_initial_layout=free_initial_layout)
        planner.threads_layout = True
        planner.accepts_free_layout = True

        def factory(n, dev):
            return ShardedState(n, device=dev, group=group, host_state_factory=host_state_factory)
        super().__init__(device=device if device is not None else 0, fusion=True, state_factory=factory, planner=planner,
                         plan_config=plan_config)
        self.group = group
        self._host_state_factory = host_state_factory

    GATHER_LIMIT = 30               # ShardedState.download gathers 2^n amplitudes on every host

    def run(self, gates, n_qubits, shots=None, initial=None, returns=None, **kwargs):
        # refuse BEFORE the circuit runs what could only fail after it: a register this large does not come back as
        # one ndarray (use shots=, hamiltonian=, returns="_inner_ctx" + device_state.download_shard()/save())
        wants_vector = (returns in ("statevector", "statevector_and_shots")
                        or (returns is None and shots is None and kwargs.get("hamiltonian") is None))
        if n_qubits > self.GATHER_LIMIT and wants_vector:
            raise ValueError(f"a sharded {n_qubits}-qubit statevector is not gathered on the host (limit "
                             f"{self.GATHER_LIMIT}): ask for shots=, hamiltonian=, or returns='_inner_ctx' and read "
                             f"device_state.download_shard() / device_state.save(directory)")
        return super().run(gates, n_qubits, shots=shots, initial=initial, returns=returns, **kwargs)

    def _download(self, st):
        if st.n_qubits > self.GATHER_LIMIT:          # returns="_inner_ctx": the context carries the device state only
            return None
        return super()._download(st)

    def copy(self) -> "ShardedB200Backend":
        other = ShardedB200Backend(self.device, self.plan_config, self.group, self._host_state_factory)
        if self.cache is not None:
            other.cache = self.cache.clone()
            other.cache_idx = self.cache_idx
        return other
