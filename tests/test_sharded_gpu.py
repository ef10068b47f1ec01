"""The sharded path on real GPUs (NCCL + CUDA IPC), world 2 / 4 / 8: each case is skipped unless the box has that many.
Run with ``gpurun --gpus 8 -- python -m pytest tests/test_sharded_gpu.py -m gpu``."""
import os
import random
import socket
import sys
import traceback
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        from qaqarot_b200 import _lib
        return _lib.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, errors):
    try:
        for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device(f"cuda:{rank}"))
        import cases as C
        from oracle import sv_oracle as O
        from qaqarot_b200 import Circuit, ShardedB200Backend, X, Y, Z, workloads as W

        be = ShardedB200Backend(device=rank)

        def oracle(circ, **kw):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return O.OracleBackend().run(circ.ops, circ.n_qubits, **kw)

        rng = np.random.default_rng(11)
        for n, circ in ((16, W.qft(16)), (15, W.random_layers(15, 6)), (14, W.qaoa_maxcut(14)[0]),
                        (15, C.build(Circuit, {"n": 15, "ops": C.random_ops(rng, 15, 120, True)}))):
            got = circ.run(backend=be)
            assert np.abs(got - oracle(circ)).max() <= 1e-12, f"statevector {n}"
            assert be._compiled.prefix.n_exchanges > 0
        circ = W.random_layers(15, 4).m[:]
        random.seed(4)
        want = oracle(circ, shots=200)
        random.seed(4)
        assert circ.run(backend=be, shots=200) == want
        circ = Circuit(14).h[:].cx[0, 13].m[13].h[13].reset[12].rx(0.3)[12].m[:]
        random.seed(5)
        rsv, rcnt = oracle(circ, shots=10, returns="statevector_and_shots")
        random.seed(5)
        gsv, gcnt = circ.run(backend=be, shots=10, returns="statevector_and_shots")
        assert gcnt == rcnt and np.abs(gsv - rsv).max() <= 1e-12
        circ = W.random_layers(14, 3)
        psi = oracle(circ)
        h = Z[0] * Z[13] + 0.5 * Z[12] * Z[13] - 1.5 * X[13] * Y[2] + 2.0
        terms = [(1.0, [("Z", 0), ("Z", 13)]), (0.5, [("Z", 12), ("Z", 13)]), (-1.5, [("X", 13), ("Y", 2)]), (2.0, [])]
        assert abs(circ.run(backend=be, hamiltonian=h) - O.expectation(terms, psi)) < 1e-10
        # ignore_global=True on a layout permuted by swaps: first amplitude above 1e-7 in wire order, found per shard
        circ = Circuit(15).h[:].rz(0.7)[:].swap[0, 14].swap[1, 13].cx[14, 3].t[2].x[0]
        assert np.abs(circ.run(backend=be, ignore_global=True) - oracle(circ, ignore_global=True)).max() <= 1e-12
        # checkpoint: chunked shard files + header, restored into a fresh register on the same GPUs
        import tempfile
        from qaqarot_b200.sharded import ShardedState
        circ = C.build(Circuit, {"n": 15, "ops": C.random_ops(rng, 15, 60, True)}).swap[0, 14].h[14]
        want = oracle(circ)
        ctx = be._execute(circ.ops, 15, 1, None, False, True)
        path = [tempfile.mkdtemp() if rank == 0 else None]
        dist.broadcast_object_list(path, src=0)
        old_chunk, ShardedState.CHECKPOINT_CHUNK = ShardedState.CHECKPOINT_CHUNK, 1 << 10     # several pieces per shard
        ctx.device_state.save(path[0])
        fresh = ShardedState(15, device=rank)
        fresh.load(path[0])
        ShardedState.CHECKPOINT_CHUNK = old_chunk
        assert fresh.layout == ctx.device_state.layout
        assert np.abs(fresh.download() - want).max() <= 1e-12
        fresh.close()
        # size-independent checks at 28 qubits (2 GiB per shard on two GPUs): norm, U U^dagger = identity
        big = W.random_layers(28, 2)
        both = big + big.dagger()
        st = be._execute(both.ops, 28, 1, None, False, True).device_state          # (no gather of the 4 GiB register)
        assert abs(st.norm2() - 1.0) < 1e-10
        assert abs(st.prob_zero(27) - 1.0) < 1e-10 and abs(st.prob_zero(3) - 1.0) < 1e-10
        be.close()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errors.put((rank, traceback.format_exc()))
        raise


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_backend_on_gpus(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    errors = ctx.Queue()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [ctx.Process(target=_worker, args=(r, world, port, errors)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    msgs = []
    while not errors.empty():
        msgs.append(errors.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            msgs.append((-1, "worker timed out"))
    assert not msgs, "\n".join(f"rank {r}:\n{m}" for r, m in msgs)
    assert all(p.exitcode == 0 for p in procs)
