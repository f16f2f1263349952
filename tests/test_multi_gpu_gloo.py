"""N>1 host path on CPU: two processes over gloo shard the sample passes (multigpu.shard_passes), each renders its share with
the oracle standing in for the GPU context (same render_samples(first, n, stride) contract), and a reduce(SUM) of the
per-rank running sums into a scratch buffer reproduces the single-process image — also on a SECOND readback after more passes
(the running sums are not modified by a readback).  The GPU version of the same test is tests/test_gpu_render.py::
test_sample_stride_sharding_sums_to_single; the NCCL reduce itself is exercised by `bench.py --gpus N`."""
import os, subprocess, sys, textwrap
import numpy as np
from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
    import numpy as np, torch, torch.distributed as dist
    from conftest import scene_at
    from glsl_pathtracer_b200 import multigpu
    from oracle import binding as ob

    class OracleCtx:                      # oracle behind the Context.render_samples contract
        def __init__(self, sc):
            self.o = ob.Oracle(sc); w, h = self.o.size; self.acc = np.zeros((h, w, 4), np.float32)
        def render_samples(self, first, n, stride):
            for i in range(n):
                self.o.render(first + i * stride, 1, accum=self.acc)

    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{port}', rank=int(sys.argv[1]), world_size=2)
    rank = dist.get_rank()
    sc = scene_at('cornell_box_orig', 48, 32, 24, 16, 3)
    ctx = OracleCtx(sc)
    drv = multigpu.ShardedRenderer(ctx, rank, 2)
    t = torch.from_numpy(ctx.acc)
    drv.render(5)
    first = drv.reduced(t).numpy().copy()     # progressive readback after the first batch ...
    drv.render(4)                             # ... rendering goes on: 9 passes in total, odd split
    second = drv.reduced(t).numpy().copy()    # ... and a second readback: must not count the first batch twice
    if rank == 0:
        o = ob.Oracle(sc)
        ref5 = o.render(1, 5).copy(); ref9 = o.render(6, 4, accum=ref5.copy())
        assert drv.samples_total() == 9
        np.testing.assert_allclose(first, ref5, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(second, ref9, rtol=1e-5, atol=1e-5)
        own = ob.Oracle(sc); mine = np.zeros_like(ref9)
        for p in multigpu.passes_of(1, 5, 0, 2) + multigpu.passes_of(6, 4, 0, 2): own.render(p, 1, accum=mine)
        np.testing.assert_array_equal(ctx.acc, mine)      # rank 0's running sum still holds only its own passes
        print('GLOO_OK')
    dist.barrier(); dist.destroy_process_group()
""")


def test_two_rank_sample_sharding_and_reduce(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GLOO_OK" in outs[0]
