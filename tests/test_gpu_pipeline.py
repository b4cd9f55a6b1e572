"""GPU: engine.HostPipeline - the double-buffered host -> device -> host driver gives exactly what a direct call of the
same pipeline gives, slot after slot, with different inputs in flight in the two slots."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_graph", [True, False])
def test_host_pipeline_matches_direct_evaluation(use_graph):
    import torch

    from eftpipe_b200 import engine, plan, synthetic

    B = 40  # not a multiple of 32: padded lanes inside
    dp = engine.DevicePlan(plan.build_tracer_plan(Nl=3))
    batches = [synthetic.make_batch(B, 0.7, seed=100 + i) for i in range(3)]

    def fn(plin, f):
        pm, _ = dp.eval_terms(plin, f)
        return pm, pm.sum(dim=(1, 2, 3))

    pipe = engine.HostPipeline(fn, dict(plin=(B, 200), f=(B,)), nslots=2, use_graph=use_graph)
    direct = []
    for b in batches:
        pm, _ = dp.eval_terms(b.plin, b.f)
        direct.append(pm.cpu().numpy())
    got = []
    for i, b in enumerate(batches):  # slot 0, 1, 0: the third submit reuses slot 0 after its read-back
        slot = i % 2
        if i >= 2:
            pipe.wait(slot)
        pipe.host_in(slot)["plin"].copy_(torch.as_tensor(b.plin))
        pipe.host_in(slot)["f"].copy_(torch.as_tensor(b.f))
        pipe.submit(slot)
        if i == 1:  # read slot 0's result while slot 1 is in flight
            got.append(pipe.wait(0)[0].numpy().copy())
    got.append(pipe.wait(1)[0].numpy().copy())
    got.append(pipe.wait(0)[0].numpy().copy())
    for g, d in zip(got, direct):
        assert np.array_equal(g, d)
    assert pipe.h2d_bytes == (B * 200 + B) * 8 and pipe.d2h_bytes == (direct[0].size + B) * 8
