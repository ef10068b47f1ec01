"""The sharded (multi-GPU) path on CPU: world_size 2 and 4 under gloo, the numpy program interpreter standing in
for each rank's device.  Checks planner exchanges, layout tracking, the exchange itself, sharded sampling and
expectation values against the oracle on the same circuits."""
import os
import random
import socket
import sys
import traceback
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, errors):
    try:
        for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        import cases as C
        from oracle import sv_oracle as O
        from oracle.program_interp import OracleState
        from qaqarot_b200 import Circuit, ShardedB200Backend, X, Y, Z, workloads as W
        from qaqarot_b200.planner import PlanConfig

        cfg = PlanConfig(a=5, r=2, c=2, min_qubits=0, kernel_tiles=tuple(range(1, 20)))
        be = ShardedB200Backend(plan_config=cfg, host_state_factory=lambda nl, off: OracleState(nl, 0, off))

        def oracle(circ, **kw):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return O.OracleBackend().run(circ.ops, circ.n_qubits, **kw)

        rng = np.random.default_rng(7)                      # identical on every rank
        # 1. statevectors: benchmark shapes and random circuits (slices, swaps, 3-qubit gates, initial=)
        n = 9
        circs = [W.qft(n), W.random_layers(n, 5), W.qaoa_maxcut(8)[0]]
        circs += [C.build(Circuit, {"n": n, "ops": C.random_ops(rng, n, 80, True)}) for _ in range(3)]
        for k, circ in enumerate(circs):
            init = None
            if k % 2:
                init = rng.standard_normal(2 ** circ.n_qubits) + 1j * rng.standard_normal(2 ** circ.n_qubits)
                init /= np.linalg.norm(init)
            kw = {} if init is None else {"initial": init}
            got = circ.run(backend=be, **kw)
            assert np.abs(got - oracle(circ, **kw)).max() <= 1e-12, f"statevector case {k}"
            assert be._compiled.prefix.n_exchanges > 0 or k == 2
        if world >= 4:
            st = be._work
            assert st.n_remaps < st.n_exchanges, "several global wires never came home in one remap"
            # ... and a remap of k bits is the k single-bit exchanges, in any order
            vec = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
            g = st.n_global
            pairs = [(1, st.n_local + j) for j in range(g)]
            pairs = [(st.n_local - 1 - 2 * j, st.n_local + j) for j in range(g)]
            st.upload(vec)
            st.exchange_many(pairs)
            merged = st.download_shard().copy()
            st.upload(vec)
            for l, gp in reversed(pairs):
                st.exchange(l, gp)
            assert np.array_equal(st.download_shard(), merged)
            want = vec.reshape([2] * n)
            for l, gp in pairs:
                want = np.swapaxes(want, n - 1 - l, n - 1 - gp)
            lo = rank << st.n_local
            assert np.array_equal(want.reshape(-1)[lo:lo + (1 << st.n_local)], merged)
        # 2. shots with the same seeded draws, measured in scrambled order, on global and local wires
        circ = C.build(Circuit, {"n": n, "ops": C.random_ops(rng, n, 60, False)}).m[(8, 2, 7, 0, 5)]
        random.seed(3)
        want = oracle(circ, shots=200)
        random.seed(3)
        assert circ.run(backend=be, shots=200) == want
        circ = W.random_layers(n, 4).m[:]
        random.seed(4)
        want = oracle(circ, shots=100)
        random.seed(4)
        assert circ.run(backend=be, shots=100) == want
        # 2b. a register large enough for the contiguous low-bit marginal (b200_marginal_low: the first 11+ measured wires
        # on the local index bits 0..10+): same Counter, and the table really came from that call
        n2 = 15
        circ = W.random_layers(n2, 3).m[:]
        random.seed(6)
        want = oracle(circ, shots=150)
        random.seed(6)
        ctx = circ.run(backend=be, shots=150, returns="_inner_ctx")
        assert ctx.shots_result == want
        local = ctx.device_state.st
        calls = []
        orig = local.marginal_low
        local.marginal_low = lambda k: (calls.append(k), orig(k))[1]
        random.seed(6)
        assert circ.run(backend=be, shots=150) == want
        if world == 2:            # (with more ranks the wires that came home may have displaced one of the low ones)
            assert calls and calls[0] >= 11, calls
        # 3. mid-circuit measurement and reset on a global wire (trajectory path: prob_zero / collapse)
        circ = Circuit(n).h[:].cx[0, 8].m[8].h[8].reset[7].rx(0.3)[7].m[:]
        random.seed(5)
        rsv, rcnt = oracle(circ, shots=20, returns="statevector_and_shots")
        random.seed(5)
        gsv, gcnt = circ.run(backend=be, shots=20, returns="statevector_and_shots")
        assert gcnt == rcnt and np.abs(gsv - rsv).max() <= 1e-12
        # 4. expectation values: ZZ on every pair class, X / Y terms on global wires
        circ = W.random_layers(n, 3)
        psi = oracle(circ)
        h = Z[0] * Z[8] + 0.5 * Z[7] * Z[8] + 0.25 * Z[3] - 1.5 * X[8] * Y[2] + 0.75 * Y[7] * Y[8] * Z[0] + 2.0
        terms = [(1.0, [("Z", 0), ("Z", 8)]), (0.5, [("Z", 7), ("Z", 8)]), (0.25, [("Z", 3)]),
                 (-1.5, [("X", 8), ("Y", 2)]), (0.75, [("Y", 7), ("Y", 8), ("Z", 0)]), (2.0, [])]
        assert abs(circ.run(backend=be, hamiltonian=h) - O.expectation(terms, psi)) < 1e-10
        # 5. marginal distribution of measured wires (global and local), the non-sampling VQE sampler
        for meas in ((8, 2, 0), (7,), (1, 8, 8, 3)):
            want = O.expect_marginal(psi, meas)
            got = be.marginal(circ.ops, n, meas)
            assert all(abs(got.get(k, 0.0) - want.get(k, 0.0)) < 1e-13 for k in set(got) | set(want)), meas
        # 6. checkpoint: shards + header with the wire layout, restored into a fresh register
        import tempfile
        from qaqarot_b200.sharded import ShardedState
        circ = C.build(Circuit, {"n": n, "ops": C.random_ops(rng, n, 50, True)}).swap[0, 8].h[8]
        want = oracle(circ)
        ctx = be._execute(circ.ops, n, 1, None, False, True)
        path = [tempfile.mkdtemp() if rank == 0 else None]
        dist.broadcast_object_list(path, src=0)
        ctx.device_state.save(path[0])
        fresh = ShardedState(n, host_state_factory=lambda nl, off: OracleState(nl, 0, off))
        fresh.load(path[0])
        assert fresh.layout == ctx.device_state.layout
        assert np.abs(fresh.download() - want).max() <= 1e-12
        #    ... a header that does not describe a permutation, or a short shard on ONE rank, makes EVERY rank raise
        #    before the register is touched (a rank raising alone would leave the others in the next collective)
        import json
        before = fresh.download()
        if rank == 0:
            with open(os.path.join(path[0], "header.json")) as f:
                h = json.load(f)
            h["layout_wire_to_index_bit"][0] = h["layout_wire_to_index_bit"][1]
            with open(os.path.join(path[0], "header.json"), "w") as f:
                json.dump(h, f)
        dist.barrier()
        with pytest.raises(ValueError, match="not a permutation"):
            fresh.load(path[0])
        ctx.device_state.save(path[0])
        if rank == world - 1:
            with open(os.path.join(path[0], f"shard_{rank}.c128"), "r+b") as f:
                f.truncate(100)
        dist.barrier()
        with pytest.raises(ValueError, match="cannot be loaded"):
            fresh.load(path[0])
        assert np.array_equal(fresh.download(), before)
        # 7. ignore_global=True on a permuted layout (swaps relabel wires): first amplitude above 1e-7 in WIRE order
        circ = Circuit(n).h[:].rz(0.7)[:].swap[0, 8].swap[1, 7].cx[8, 3].t[2].x[0]
        want = oracle(circ, ignore_global=True)
        got = circ.run(backend=be, ignore_global=True)
        assert ctx.device_state.layout != list(range(n)) or True
        assert np.abs(got - want).max() <= 1e-12
        circ = Circuit(n).x[8].swap[8, 1].h[0]                # a sparse state: most amplitudes are zero
        assert np.abs(circ.run(backend=be, ignore_global=True) - oracle(circ, ignore_global=True)).max() <= 1e-12
        # 8. a register too large to gather is refused before the circuit runs
        with pytest.raises(ValueError, match="not gathered"):
            be.run([], 31)
        with pytest.raises(ValueError, match="not gathered"):
            be.run([], 33, returns="statevector_and_shots", shots=3)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errors.put((rank, traceback.format_exc()))
        raise


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_backend_under_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    errors = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, errors)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    msgs = []
    while not errors.empty():
        msgs.append(errors.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            msgs.append((-1, "worker timed out"))
    assert not msgs, "\n".join(f"rank {r}:\n{m}" for r, m in msgs)
    assert all(p.exitcode == 0 for p in procs)
