"""CPU tests of the Qwen-Image-Edit host side: dispatch on the pipeline class name, patching / restoring, the stand-in
rotary table's contract, and the not-yet-built families failing loudly."""
import pytest
import torch

from regione_b200 import RegionEHelper, params
from standins import standin
from regione_b200 import qwen_image_edit as qw


def _pipe(cls=standin.QwenImageEditPipeline):
    tr = standin.QwenImageTransformer2DModel(dim=256, heads=2, n_blocks=2, mlp_ratio=4, in_channels=64, ctx_dim=64)
    return cls(tr)


def test_qwen_enable_disable():
    """RegionE/QwenImageEdit/inplace.py:53-71."""
    pipe = _pipe()
    base_cls, base_sched = pipe.__class__, pipe.scheduler.__class__
    h = RegionEHelper(pipe)
    assert h.config["threshold"] == 0.80 and h.config["cache_threshold"] == 0.03       # RegionE.py:5
    h.enable()
    assert pipe.__class__.__name__ == "RegionEQwenImageEditPipeline" and isinstance(pipe, base_cls)
    assert pipe.scheduler.__class__.__name__ == "RegionEFlowMatchEulerDiscreteScheduler"
    assert pipe.scheduler._regione_manager is qw.MANAGER
    assert all(b.attn.processor is not None for b in pipe.transformer.transformer_blocks)
    assert qw.gamma == params.GAMMA["QwenImageEditPipeline"]
    h.disable()
    assert pipe.__class__ is base_cls and pipe.scheduler.__class__ is base_sched
    assert "forward" not in pipe.transformer.__dict__


def test_qwen_plus_uses_its_own_gamma():
    class QwenImageEditPlusPipeline(standin.QwenImageEditPipeline):
        pass
    pipe = _pipe(QwenImageEditPlusPipeline)
    h = RegionEHelper(pipe)
    h.enable()
    try:
        assert qw.gamma == params.GAMMA["QwenImageEditPlusPipeline"] != params.GAMMA["QwenImageEditPipeline"]
    finally:
        h.disable()


def test_qwen_rope_table_contract():
    """pos_embed(img_shapes, txt_seq_lens) -> unit-modulus complex [L+C, 64] / [T, 64]; noise and condition images of
    equal shape differ only in the frame axis (first 8 frequencies)."""
    r = standin.QwenEmbedRope()
    img, txt = r([[(1, 6, 10), (1, 6, 10)]], [7])
    assert img.shape == (120, 64) and txt.shape == (7, 64) and img.dtype == torch.complex64
    assert torch.allclose(img.abs(), torch.ones(120, 64), atol=1e-5)
    assert torch.allclose(img[:60, 8:], img[60:, 8:]) and not torch.allclose(img[:60, :8], img[60:, :8])


def test_step1x_enable_disable_both_versions():
    from standins import step1x as sx
    from regione_b200 import step1x_edit as s1
    tr = sx.Step1XEditTransformer2DModel(dim=256, heads=2, n_double=1, n_single=1, ctx_dim=64, vec_dim=32)
    pipe = sx.Step1XEditPipeline(tr)
    h = RegionEHelper(pipe)
    assert h.config["threshold"] == 0.88 and h.config["cache_threshold"] == 0.02     # RegionE.py:3
    h.enable()
    assert pipe.__class__.__name__ == "RegionEStep1XEditPipeline" and pipe.scheduler._regione_manager is s1.MANAGER
    assert [b.attn.processor.single for b in list(tr.transformer_blocks) + list(tr.single_transformer_blocks)] == \
        [False, True]
    h.disable()
    assert pipe.__class__ is sx.Step1XEditPipeline

    class Step1XEditPipelineV1P2(sx.Step1XEditPipeline):
        pass
    from regione_b200 import step1x_edit_v1p2 as s2
    pipe2 = Step1XEditPipelineV1P2(tr)
    h2 = RegionEHelper(pipe2)
    h2.enable()
    assert pipe2.__class__.__name__ == "RegionEStep1XEditPipelineV1P2"
    assert pipe2.scheduler._regione_manager is s2.MANAGER and s2.gamma == params.GAMMA["Step1XEditPipelineV1P2"]
    h2.disable()
    assert pipe2.__class__ is Step1XEditPipelineV1P2


def test_config2_step_count_and_gamma_are_inputs():
    """BASELINE configs[2] (Qwen-Image-Edit, 20 steps, cache_threshold 0.02): the reference refuses step counts other
    than 28 (RegionE.py:44, utils.py:391) because its tables have 27 fitted entries. Here the table is an input of
    `set_params`; the host planner must then equal the oracle's AVDC rule evaluated with the same table (parity is
    oracle-with-same-table only: unpinned w.r.t. the reference)."""
    from oracle.schedule import avdc_plan, flow_match_sigmas
    from regione_b200.manager import plan_steps
    table = params.resample_gamma(params.GAMMA["QwenImageEditPipeline"], 20)
    assert len(table) == 19 and table[0] == params.GAMMA["QwenImageEditPipeline"][0]
    assert table[-1] == params.GAMMA["QwenImageEditPipeline"][-1]
    assert params.resample_gamma(params.GAMMA["QwenImageEditPipeline"], 28) == params.GAMMA["QwenImageEditPipeline"]
    pipe = _pipe()
    h = RegionEHelper(pipe)
    saved = dict(h.config)
    try:
        with pytest.raises(AssertionError):
            h.set_params(num_inference_steps=20)                       # no table: the reference's assertion stands
        with pytest.raises(AssertionError):
            h.set_params(num_inference_steps=20, gamma=table[:-1])     # wrong length
        h.set_params(num_inference_steps=20, cache_threshold=0.02, gamma=table)
        h.enable()
        try:
            M = qw.MANAGER
            assert M.inference_step == 20 and M.gamma == table and M.refresh_step == [16, 19]
            _, ts = flow_match_sigmas(20, 2304)
            got = plan_steps(ts, qw.gamma, M)
            want = avdc_plan(ts, table, 6, 2, "16", 0.02, inference_step=20)
            assert [s for s, _ in got] == [w["mode"] == "SKIP" for w in want]
            for (_, r), w in zip(got, want):
                assert (r is None) == (w["ratio"] is None) and (r is None or float(r) == w["ratio"])
            modes = [w["mode"] for w in want]
            assert modes[:6] == ["FULL"] * 6 and modes[15] == "FULL" and modes[18:] == ["FULL", "FULL"]
        finally:
            h.disable()
        h.set_params()                                                 # back to the reference's 28-step defaults
        assert h.config["num_inference_steps"] == 28 and "gamma" not in h.config
    finally:
        h.config.clear()
        h.config.update(saved)
