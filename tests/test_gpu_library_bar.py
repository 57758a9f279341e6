"""Library bar: the same forward as eager PyTorch on the same GPU (fp16 autocast: cuBLAS tensor-core GEMMs,
ATen elementwise kernels, every head + pre-training heads as the reference executes) next to the engine, batch 64, 36 regions
x 30 tokens.
Not a parity test: it records both throughputs in the parity log and only asserts that the engine is the faster one."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_engine_beats_torch_eager_autocast(full_oracle, parity_log):
    import copy
    import vilbert_b200 as vb
    from oracle import vilbert_ref as R
    B = 64
    inp = R.make_inputs(B, 30, 36, seed=99, full_masks=True)
    dev = [t.cuda() for t in inp]
    eager = copy.deepcopy(full_oracle).cuda().eval()

    def time_ms(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        t_eager = time_ms(lambda: eager(*dev, compute_pretraining_heads=True), 5)
    del eager
    cfg = vb.BertConfig.from_dict(full_oracle.config.to_dict())
    eng = vb.VILBertForVLTasks.from_pretrained(full_oracle.state_dict(), config=cfg, num_labels=full_oracle.num_labels).eval().cuda(0)
    t_eng = time_ms(lambda: eng(*dev), 20)
    eng.close()
    parity_log(test="library_bar_B64", torch_eager_fp16_autocast_ms=t_eager, engine_ms=t_eng, torch_eager_pairs_per_s=B / t_eager * 1e3,
               engine_pairs_per_s=B / t_eng * 1e3)
    assert t_eng < t_eager
