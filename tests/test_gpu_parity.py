"""Parity of the CUDA path (through the C ABI) against the reference fixtures and the CPU oracle.

Run on the B200 box: ``python -m pytest tests -m gpu``.  Tolerance: 1e-12 max-abs on amplitudes
(BASELINE.json north_star); shot Counters must be equal for the same seeded ``random.random()`` stream.
"""
import random
import warnings
from collections import Counter

import numpy as np
import pytest

import cases as C
from conftest import case_names
from helpers import TOL, maxabs, run_case
from oracle import sv_oracle as O
from qaqarot_b200 import B200Backend, Circuit, X, Y, Z

pytestmark = pytest.mark.gpu
MODES = [False, True]


@pytest.fixture(scope="module", params=MODES, ids=["unfused", "fused"])
def backend(request):
    be = B200Backend(fusion=request.param)
    yield be
    be.close()


@pytest.mark.parametrize("name", case_names("statevector"))
def test_statevector_cases(name, backend, gold):
    assert maxabs(run_case(gold.cases[name], backend), gold.sv(name)) <= TOL


@pytest.mark.parametrize("name", case_names("shots"))
def test_shot_cases(name, backend, gold):
    assert dict(run_case(gold.cases[name], backend)) == gold.meta["cases"][name]["shots"]


def test_statevector_and_shots(backend, gold):
    sv, cnt = run_case(gold.cases["sv_and_shots"], backend)
    assert dict(cnt) == gold.meta["cases"]["sv_and_shots"]["shots"]
    assert maxabs(sv, gold.sv("sv_and_shots")) <= TOL


def test_samples(backend, gold):
    assert run_case(gold.cases["samples_keys"], backend) == gold.meta["cases"]["samples_keys"]["samples"]


def test_hamiltonian(backend, gold):
    letter = {"X": X, "Y": Y, "Z": Z}
    for name, recs in gold.meta["expect"].items():
        for rec in recs:
            h = 0
            for coeff, letters in rec["terms"]:
                t = coeff
                for ch, q in letters:
                    t = t * letter[ch][q]
                h = h + t
            circ = C.build(Circuit, gold.cases[name])
            init = C.initial_state(gold.cases[name])
            kw = {} if init is None else {"initial": init}
            assert abs(circ.run(backend=backend, hamiltonian=h, **kw) - rec["value"]) < 1e-10


def _oracle(circ, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return O.OracleBackend().run(circ.ops, circ.n_qubits, **kw)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 11, 14, 17, 20])
def test_every_target_position_against_oracle(n, backend):
    """Each single-gate kernel on every qubit position (low strides through high strides)."""
    rng = np.random.default_rng(n)
    init = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    init /= np.linalg.norm(init)
    c = Circuit(n)
    for t in range(n):
        c.h[t].rx(0.3 + t)[t].t[t].rz(0.1 * t + 0.2)[t].u(0.4, 0.5 + t, 0.6, 0.7)[t].x[t].y[t]
        if n > 1:
            o = (t + 1 + int(rng.integers(n - 1))) % n
            c.cx[o, t].cz[t, o].cu(0.3, 0.2, 0.1, 0.5)[o, t].crz(0.9)[t, o].cphase(0.7)[o, t].swap[o, t]
        if n > 2:
            a, b = [q for q in rng.permutation(n) if q != t][:2]
            c.ccx[int(a), int(b), t].ccz[t, int(a), int(b)]
    assert maxabs(c.run(backend=backend, initial=init), _oracle(c, initial=init)) <= TOL


@pytest.mark.parametrize("n,depth", [(16, 20), (20, 20)])
def test_c1_random_layers(n, depth, backend):
    """BASELINE configs[0]: Circuit(20) random H/RZ/CX layers depth 20, statevector + shots=1000."""
    case = {"name": "c1", "n": n, "ops": C.random_layer_ops(n, depth), "initial_seed": None,
            "run": {"shots": None, "seed": None, "returns": None}}
    circ = C.build(Circuit, case)
    ref = _oracle(circ)
    assert maxabs(circ.run(backend=backend), ref) <= TOL
    circ.m[:]
    shots = 1000
    random.seed(7)
    want = _oracle(circ, shots=shots)
    random.seed(7)
    assert circ.run(backend=backend, shots=shots) == want


@pytest.mark.parametrize("n", [12, 18, 21])
def test_c2_qft(n, backend):
    """BASELINE configs[1] at oracle-sized n: against the oracle and the analytic answer."""
    x = 123456789 % (1 << n)
    case = {"name": "qft", "n": n, "ops": C.qft_ops(n, x), "initial_seed": None,
            "run": {"shots": None, "seed": None, "returns": None}}
    circ = C.build(Circuit, case)
    got = circ.run(backend=backend)
    assert maxabs(got, _oracle(circ)) <= TOL
    y = np.arange(1 << n, dtype=np.int64)
    analytic = np.exp(2j * np.pi * ((x * y) % (1 << n)) / (1 << n)) / np.sqrt(1 << n)
    assert maxabs(got, analytic) <= 1e-11


@pytest.mark.parametrize("n", [14, 20])
def test_c3_qaoa_energy(n, backend):
    ops, edges = C.qaoa_ops(n)
    case = {"name": "qaoa", "n": n, "ops": ops, "initial_seed": None,
            "run": {"shots": None, "seed": None, "returns": None}}
    circ = C.build(Circuit, case)
    h = sum((Z[i] * Z[j] for i, j in edges[1:]), Z[edges[0][0]] * Z[edges[0][1]])
    psi = _oracle(circ)
    want = O.expectation([(1.0, [("Z", i), ("Z", j)]) for i, j in edges], psi)
    assert abs(circ.run(backend=backend, hamiltonian=h) - want) < 1e-10
    assert maxabs(circ.run(backend=backend), psi) <= TOL


def test_c3_qaoa_28_qubits_energy_against_the_statevector_route():
    """BASELINE configs[2] at full size: run(hamiltonian=...) -- reductions on the device-resident state --
    against the energy computed on the host from the downloaded statevector, term by term."""
    from qaqarot_b200 import workloads as W
    n = 28
    circ, ham, edges = W.qaoa_maxcut(n)
    be = B200Backend()
    got = circ.run(backend=be, hamiltonian=ham)
    psi = circ.run(backend=be)
    prob = (psi.real ** 2 + psi.imag ** 2).reshape([2] * n)            # axis k = qubit n-1-k
    del psi
    assert abs(float(prob.sum()) - 1.0) < 1e-10
    want = 0.0
    for i, j in edges:
        keep = (n - 1 - i, n - 1 - j)
        m = prob.sum(axis=tuple(a for a in range(n) if a not in keep))
        want += float(m[0, 0] + m[1, 1] - m[0, 1] - m[1, 0])
    be.close()
    assert abs(got - want) < 1e-10


@pytest.mark.parametrize("n", [16, 28])
def test_objective_loop_new_angles_every_call(n):
    """SURVEY 8(f) rank 1 (vqe.py:67-83): the optimiser rebuilds the ansatz with new angles on every call.  State and
    work buffers stay on the device, nothing but the energy returns; the first call is launched pass by pass while the
    planner is still working (b200_apply_program_ex), the later ones at once from param_replay's model of the planner,
    confirmed by the real plan made while the device runs.  Five random angle vectors against a backend that plans first
    and launches afterwards, and against the oracle at 16 qubits."""
    from qaqarot_b200 import workloads as W
    streamed, plain = B200Backend(), B200Backend()
    plain.stream_plans = False
    plain.param_replay = False
    worst = 0.0
    for seed in range(5):
        circ, ham, edges = W.qaoa_maxcut(n, p=4, seed=100 + seed)
        e1 = circ.run(backend=streamed, hamiltonian=ham)
        e2 = circ.run(backend=plain, hamiltonian=ham)
        worst = max(worst, abs(e1 - e2))
        if n <= 16:
            psi = _oracle(circ)
            want = O.expectation([(1.0, [("Z", i), ("Z", j)]) for i, j in edges], psi)
            assert abs(e1 - want) < 1e-10
            assert maxabs(circ.run(backend=streamed), psi) <= TOL        # (same angles: the cached plan, one call)
    # from the third call on the circuit was launched with the payload param_replay's model of the planner predicted,
    # and the real plan confirmed every one of them
    assert streamed.replay_stats == {"built": 1, "hits": 3, "misses": 0, "unmodelable": 0}
    streamed.close()
    plain.close()
    assert worst <= 1e-12


def test_runs_are_bit_reproducible():
    """The same program three times on 2^24 amplitudes: bit-identical statevectors.  (compute-sanitizer is not available on
    the GPU pool; a race between the TMA pipeline, the group barriers and the register tile would show up here as a
    difference in some run.)  Layers + QFT: layer records, tables with variants, fans."""
    from qaqarot_b200 import workloads as W
    n = 24
    be = B200Backend()
    for circ in (W.random_layers(n, 10), W.qft(n, 987654321)):
        first = circ.run(backend=be).copy()
        for _ in range(2):
            be._compiled = None
            assert np.array_equal(circ.run(backend=be), first)
    be.close()


def test_readme_hamiltonian_example(backend):
    """/root/reference/README.md:83-89: Circuit(4).x[:].run(hamiltonian=Z[0] + Z[1]) == -2.0."""
    assert Circuit(4).x[:].run(backend=backend, hamiltonian=1 * Z[0] + 1 * Z[1]) == -2.0


def test_ignore_global_phase(backend):
    """utils.ignore_global_phase (utils.py:130-144) on the device: k_first_above + k_scale."""
    rng = np.random.default_rng(3)
    for n in (1, 4, 11, 16):
        case = {"name": "ig", "n": n, "ops": C.random_ops(rng, n, 30 + 2 * n, True), "initial_seed": None,
                "run": {"shots": None, "seed": None, "returns": None}}
        circ = C.build(Circuit, case)
        assert maxabs(circ.run(backend=backend, ignore_global=True), _oracle(circ, ignore_global=True)) <= TOL
    v = Circuit().x[0].rz(1.0)[0].run(backend=backend, ignore_global=True)
    assert np.allclose(v, [0, 1], atol=1e-15)
    # the first amplitudes are (numerically) zero: the search has to pass them
    circ = Circuit(12).x[11].h[0].rz(0.9)[11].cphase(0.3)[0, 11]
    assert maxabs(circ.run(backend=backend, ignore_global=True), _oracle(circ, ignore_global=True)) <= TOL


def test_cache_copy_and_append_on_the_device(backend):
    """make_cache -> run -> append -> run, Circuit.copy() with its backend, oneshot (numpy_backend.py:564-568,
    610-685; backendbase.py:27-33; tests/test_circuit.py test_cache_then_append) on device snapshots."""
    be = backend.copy()
    n = 10
    c = Circuit(n).h[:].rz(0.4)[3].cx[3, 9].cx[0, 5].m[:]
    c._backends["b200"] = be
    assert be.cache is None and be.cache_idx == -1
    c.make_cache(backend="b200")
    assert be.cache is not None and be.cache_idx == len(c.ops) - 2
    random.seed(3)
    want = _oracle(c, shots=100)
    random.seed(3)
    assert c.run(backend="b200", shots=100) == want
    c.x[1].h[2].m[1]
    random.seed(4)
    want = _oracle(c, shots=50)
    random.seed(4)
    assert c.run(backend="b200", shots=50) == want and be.cache is not None
    c2 = c.copy()
    be2 = c2._backends["b200"]
    assert be2 is not be and be2.cache is not None and be2.cache is not be.cache
    random.seed(5)
    want = _oracle(c2, shots=30)
    random.seed(5)
    assert c2.run(backend="b200", shots=30) == want
    random.seed(6)
    wsv, wkey = O.OracleBackend().oneshot(c.ops, n) if hasattr(O.OracleBackend, "oneshot") else (None, None)
    if wsv is not None:
        random.seed(6)
        gsv, gkey = c.oneshot(backend="b200")
        assert gkey == wkey and maxabs(gsv, wsv) <= TOL
    d = Circuit().x[0].m[0]                                   # a different qubit count invalidates the snapshot
    d._backends["b200"] = be
    assert d.run(backend="b200", shots=3) == Counter({"1": 3}) and be.cache is None
    be.close()
    be2.close()


@pytest.mark.parametrize("name", ["c1", "qft", "qaoa", "random"])
def test_fused_circuit_from_the_transpiler_backend_runs_on_the_device(name, backend):
    """run(backend="b200_fuse") returns the planner's normal form as a circuit (transpile.py): executed on the device it
    gives the original circuit's statevector."""
    n = 13
    rng = np.random.default_rng(8)
    ops = {"c1": lambda: C.random_layer_ops(n, 8), "qft": lambda: C.qft_ops(n, 4321), "qaoa": lambda: C.qaoa_ops(n - 1)[0],
           "random": lambda: C.random_ops(rng, n, 120, True)}[name]()
    nq = n - 1 if name == "qaoa" else n
    circ = C.build(Circuit, {"name": name, "n": nq, "ops": ops})
    fused = circ.run(backend="b200_fuse")
    assert isinstance(fused, Circuit) and fused.n_qubits == circ.n_qubits
    want = _oracle(circ)
    assert maxabs(fused.run(backend=backend), want) <= TOL
    assert maxabs(circ.run(backend=backend), want) <= TOL


def test_midcircuit_and_reset_against_oracle(backend):
    rng = np.random.default_rng(99)
    for n in (3, 6, 10):
        ops = C.random_ops(rng, n, 40, False) + [C._op("m", 0)] + C.random_ops(rng, n, 20, False) + \
            [C._op("reset", n - 1)] + C.random_ops(rng, n, 10, False) + [C._op("m", slice(None))]
        case = {"name": "mid", "n": n, "ops": ops, "initial_seed": None,
                "run": {"shots": 60, "seed": 5, "returns": "statevector_and_shots"}}
        circ = C.build(Circuit, case)
        random.seed(5)
        rsv, rcnt = _oracle(circ, shots=60, returns="statevector_and_shots")
        gsv, gcnt = run_case(case, backend)
        assert gcnt == rcnt
        assert maxabs(gsv, rsv) <= TOL


def test_size_independent_properties_at_scale(backend):
    """26 qubits (1 GiB): norm preservation, U then U^dagger returns |0..0>, GHZ shots."""
    n = 26
    case = {"name": "big", "n": n, "ops": C.random_layer_ops(n, 4), "initial_seed": None,
            "run": {"shots": None, "seed": None, "returns": None}}
    circ = C.build(Circuit, case)
    both = circ + circ.dagger()
    ctx = both.run(backend=backend, returns="_inner_ctx")
    st = ctx.device_state
    assert abs(st.norm2() - 1.0) < 1e-10
    idx, amp = st.first_above(0.5)
    assert idx == 0 and abs(amp - 1.0) < 1e-10
    ghz = Circuit(n).h[0]
    for q in range(n - 1):
        ghz.cx[q, q + 1]
    random.seed(1)
    cnt = ghz.m[:].run(backend=backend, shots=50)
    assert set(cnt) <= {"0" * n, "1" * n} and sum(cnt.values()) == 50


def test_backend_requires_the_cuda_library(monkeypatch):
    """No silent CPU fallback: a missing library is a hard error."""
    from qaqarot_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb200sv.so")
    with pytest.raises(_lib.B200LibraryError):
        B200Backend().run(Circuit().h[0].ops, 1)


@pytest.mark.parametrize("a", [9, 10, 11, 12])
@pytest.mark.parametrize("extra", [0, 1, 3, 6])
def test_tile_kernel_sizes_against_oracle(a, extra):
    """Every tile size the fused kernel is instantiated for, on random circuits with slices, 3-qubit gates and
    controls / phases on bits outside the tile."""
    from qaqarot_b200.planner import PlanConfig
    n = a + extra
    rng = np.random.default_rng(100 * a + extra)
    init = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    init /= np.linalg.norm(init)
    circ = C.build(Circuit, {"n": n, "ops": C.random_ops(rng, n, 150, True)})
    be = B200Backend(plan_config=PlanConfig(a=a))
    try:
        got = circ.run(backend=be, initial=init)
        kinds = [o[0] for o in be._compiled.prefix._ops]
        assert 16 in kinds                                   # the fused pass really ran
    finally:
        be.close()
    assert maxabs(got, _oracle(circ, initial=init)) <= TOL


def _tile_runs(words, woff):
    """Runs of contiguous tile bits above index bit 2 in the TILE descriptor at `woff` (csrc/tile.cu build_tile_map:
    each run is one dimension of the TMA tensor; more than four are enumerated by separate copies)."""
    a = words[woff] & 0xFF
    bits = [((words[woff + 1 + i // 8]) >> (8 * (i % 8))) & 0xFF for i in range(a)]
    runs, prev = 0, None
    for q in bits[3:]:
        if prev is None or q != prev + 1:
            runs += 1
        prev = q
    return runs


@pytest.mark.parametrize("n,stride", [(20, 2), (24, 3), (27, 3)])
def test_tile_kernel_scattered_tiles(n, stride):
    """Targets on isolated index bits: the tile is a box of more than five runs of bits, so one tile moves as several
    TMA copies (2, 4, ... 32 per tile) that land side by side in the shared-memory buffer."""
    rng = np.random.default_rng(500 + n)
    circ = Circuit(n)
    targets = list(range(5, n, stride))
    for rep in range(3):
        for q in targets:
            circ.h[q].rz(float(rng.uniform(0, 6.28)))[q]
        for q0, q1 in zip(targets, targets[1:]):
            circ.cx[q0, q1]
        circ.ry(0.3 + rep)[0].rx(0.2)[3]
    be = B200Backend()
    try:
        got = circ.run(backend=be)
        prog = be._compiled.prefix
        runs = [_tile_runs(prog._words, o[5]) for o in prog._ops if o[0] == 16]
    finally:
        be.close()
    assert max(runs) > 4, runs                  # some pass really needed sub-copies
    if n <= 24:
        assert maxabs(got, _oracle(circ)) <= TOL
    else:
        # size-independent properties at 2 GiB: norm, and the inverse circuit brings |0..0> back
        assert abs(np.vdot(got, got).real - 1.0) < 1e-12
        be = B200Backend()
        try:
            back = (circ + circ.dagger()).run(backend=be)
        finally:
            be.close()
        assert abs(abs(back[0]) - 1.0) < 1e-12


def test_tile_kernel_matches_descriptor_interpreter():
    """Same encoded program on the CUDA executor and on the numpy interpreter of the descriptor."""
    from oracle.program_interp import OracleState
    from qaqarot_b200.device import DeviceState
    from qaqarot_b200.lowering import lower
    from qaqarot_b200.planner import PlanConfig, plan
    from qaqarot_b200 import workloads as W
    for n, circ in ((15, W.qft(15)), (16, W.random_layers(16, 10)), (14, W.qaoa_maxcut(14)[0])):
        for a in (10, 11, 12):
            prog = plan(lower(circ.ops, n).prims, n, PlanConfig(a=a))
            dev, ref = DeviceState(n), OracleState(n)
            dev.apply(prog)
            ref.apply(prog)
            got = dev.download()
            dev.close()
            assert maxabs(got, ref.psi) <= 1e-14


@pytest.mark.parametrize("a", [9, 12])
def test_layer_record_with_table_variants_matches_interpreter(a):
    """A hand-made round: four rotations in one LAYER record whose table carries a register-local CZ and phase, and
    three selector variants (a thread bit x register bit phase, a fixed-bit target with a register entry, a 1-qubit
    phase on a thread bit); then a general 2x2 + fan record, so both kernel instantiations (with / without the general
    target stages) run.  CUDA executor against the numpy interpreter on a dense random state."""
    from oracle.program_interp import OracleState
    from qaqarot_b200.device import DeviceState
    from qaqarot_b200 import planner as P
    from qaqarot_b200.program import Program
    n = a + 3
    rng = np.random.default_rng(90 + a)
    reg = [4, 5, 6, 7]

    def rot(q, t):
        return P.shear_item(q, float(((t + np.pi / 2) % np.pi) - np.pi / 2))

    w = [np.exp(1j * x) for x in (0.4, 1.1, -0.8, 2.0)]
    layer = [rot(q, 0.3 + q) for q in reg] + [
        P.Item("fan", t=4, w=1.0, entries={0: w[0]}), P.Item("fan", t=a + 1, w=1.0, entries={6: w[1]}),
        P.Item("fan", t=1, w=w[2], entries={}), P.Item("fan", t=5, w=w[3], entries={6: -1.0 + 0j}),
        rot(4, -1.2), rot(7, 0.9)]
    general = [P.Item("mat", t=5, m=np.array([0.6, 0.8j, 0.8j, 0.6], dtype=complex)),
               P.Item("fan", t=5, w=w[0], entries={2: w[1], a + 2: w[2]}), P.Item("x", t=6, ctrls=(a,))]
    for items, want_full in ((layer, False), (layer + general, True)):
        prog = Program()
        P.encode_pass(prog, P.Pass(list(range(a)), [P.Round(reg, items)]), n, 4, 4, False, True)
        kinds = set()
        pos = prog._words[3] >> 32
        for g in range(prog._words[pos] & 0xFFFF):
            kinds.add(prog._words[pos + 3 + 3 * g] & 0xFF)
        assert (bool(kinds & {P.T_MATC, P.T_MATR, P.T_X})) == want_full and P.T_LAYER in kinds
        psi = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
        dev, ref = DeviceState(n), OracleState(n)
        dev.upload(psi)
        ref.upload(psi)
        dev.apply(prog)
        ref.apply(prog)
        got = dev.download()
        dev.close()
        assert maxabs(got, ref.psi) <= 1e-13


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 10, 11, 14, 18, 21])
def test_permute_kernel_against_interpreter(n):
    """In-place index-bit involutions (relabelled swap gates): random disjoint transpositions, bit reversal,
    pairs among the low bits, the empty involution.  Pure data movement: bit exact."""
    from oracle.program_interp import OracleState
    from qaqarot_b200 import _lib
    from qaqarot_b200.device import DeviceState
    from qaqarot_b200.program import Program
    rng = np.random.default_rng(n)
    init = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    cases = [[(q, n - 1 - q) for q in range(n // 2)], []]
    if n >= 4:
        cases.append([(0, 1), (2, 3)])
    for _ in range(6):
        perm = [int(x) for x in rng.permutation(n)]
        k = int(rng.integers(0, n // 2 + 1))
        cases.append([(min(perm[2 * i], perm[2 * i + 1]), max(perm[2 * i], perm[2 * i + 1])) for i in range(k)])
    for pairs in cases:
        prog = Program()
        woff = prog.add_words([p | (q << 8) for p, q in pairs])
        prog.add_op(_lib.OP_PERMUTE, woff=woff, wlen=len(pairs))
        dev, ref = DeviceState(n), OracleState(n)
        dev.upload(init)
        ref.upload(init)
        dev.apply(prog)
        ref.apply(prog)
        got = dev.download()
        dev.close()
        assert np.array_equal(got, ref.psi), pairs


def test_c2_qft_30_qubits_against_the_analytic_answer():
    """BASELINE configs[1] at full size (16 GiB state): windows of the result against
    exp(2 pi i x y / N) / sqrt(N), the norm, and the known first amplitude."""
    from qaqarot_b200 import workloads as W
    n, x = 30, 123456789
    be = B200Backend()
    try:
        ctx = W.qft(n).run(backend=be, returns="_inner_ctx")
        st = ctx.device_state
        assert abs(st.norm2() - 1.0) < 1e-10
        rng = np.random.default_rng(30)
        starts = [0, (1 << n) - 4096] + [int(s) for s in rng.integers(0, (1 << n) - 4096, 6)]
        for s in starts:
            got = st.download(offset=s, count=4096)
            y = np.arange(s, s + 4096, dtype=np.int64)
            want = W.qft_analytic_amplitude(n, x, y)
            assert maxabs(got, want) <= 1e-14        # |amp| = 2^-15, so this is 3e-10 relative
    finally:
        be.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [10, 14])
def test_expect_group_with_shard_offset_many_terms_and_complex_coefficients(n):
    """b200_expect_group as a shard of a larger register sees it (index offset above n), with more terms than one table
    launch holds (k_expect_tab: 56) and complex coefficients; below 2^11 amplitudes the loop kernel.  Against the
    definition, sum conj(a[i ^ x]) a[i] sum_k c_k (-1)^popc((i | offset) & zy_k), in numpy."""
    from qaqarot_b200.device import DeviceState
    rng = np.random.default_rng(4 + n)
    dim = 1 << n
    psi = rng.standard_normal(dim) + 1j * rng.standard_normal(dim)
    psi /= np.linalg.norm(psi)
    st = DeviceState(n)
    st.upload(psi)
    idx = np.arange(dim, dtype=np.uint64)
    for offset_bits, n_terms, cplx, xmask in ((0, 7, False, 0), (3, 60, False, 0), (2, 130, True, 0b1011 << 3), (1, 56, True, 5)):
        offset = offset_bits << n
        st.set_index_offset(offset)
        zy = [int(v) for v in rng.integers(0, 1 << (n + 2), n_terms)]
        c = rng.standard_normal(n_terms) + (1j * rng.standard_normal(n_terms) if cplx else 0)
        w = np.zeros(dim, dtype=complex)
        for m, ck in zip(zy, c):
            v = (idx | np.uint64(offset)) & np.uint64(m)
            par = np.zeros(dim, dtype=np.uint64)
            for b in range(n + 2):
                par ^= (v >> np.uint64(b)) & np.uint64(1)
            w += np.where(par == 1, -ck, ck)
        b = psi[(idx ^ np.uint64(xmask)).astype(np.int64)]
        want = complex(np.sum(np.conj(b) * psi * w))
        got = st.expect_group(xmask, zy, list(c))
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (offset_bits, n_terms, got, want)
    st.close()
