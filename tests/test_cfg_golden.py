"""CFG arithmetic of the Step1X / Qwen families against tests/golden/cfg.pt, which
oracle/make_golden.make_cfg produced by exec'ing the reference's own lines (Step1XEdit/inplace.py:401-410,
QwenImageEdit/inplace.py:401-405): the oracle on the CPU bit for bit, the CUDA kernels through the C ABI within one bf16
rounding (the per-token norms are reduced in a different order on the device)."""
import os

import pytest
import torch


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "cfg.pt"), weights_only=False)


def test_oracle_cfg_functions_match_reference_lines(golden):
    from oracle.qwen import cfg_norm_rescaled
    from oracle.step1x import cfg_norm_processed
    from standins.step1x import Step1XEditPipeline
    pos, neg = golden["pos"], golden["neg"]
    assert len(golden["step1x"]) == 3 and len(golden["qwen"]) == 2
    for c in golden["step1x"]:
        got = cfg_norm_processed(pos, neg, c["scale"], torch.tensor(c["t"]), c["truncate"],
                                 Step1XEditPipeline.process_diff_norm, c["k"])
        assert got.dtype == torch.bfloat16 and torch.equal(got, c["out"])
    for c in golden["qwen"]:
        assert torch.equal(cfg_norm_rescaled(pos, neg, c["scale"]), c["out"])


@pytest.mark.gpu
def test_cuda_cfg_kernels_match_reference_lines(golden):
    from regione_b200 import ops
    from standins.step1x import Step1XEditPipeline
    pos, neg = golden["pos"][0].cuda(), golden["neg"][0].cuda()

    def close(got, want):
        got, want = got.float().cpu(), want[0].float()
        rel = float((got - want).norm() / want.norm())
        same = float((got == want).float().mean())
        assert rel <= 4e-3 and same > 0.9, (rel, same)

    for c in golden["step1x"]:
        if c["t"] > c["truncate"]:
            denom = Step1XEditPipeline.process_diff_norm(ops.cfg_diff_norm(pos, neg), c["k"])
            close(ops.cfg_combine(pos, neg, c["scale"], denom), c["out"])
        else:
            got = ops.cfg_combine(pos, neg, c["scale"])
            assert torch.equal(got.cpu(), c["out"][0])          # no reduction involved: bit-exact
    for c in golden["qwen"]:
        close(ops.cfg_rescale(pos, neg, c["scale"]), c["out"])
