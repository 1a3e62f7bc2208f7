"""The oracle restatement against fixtures produced by the reference's own code (oracle/make_golden.py)."""
import os

import torch

from oracle import region_ops as ro
from oracle import schedule as sc
from oracle.make_golden import synthetic_partition_inputs


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_token_selector_matches_reference(golden_dir):
    g = _load(golden_dir, "region_ops.pt")
    assert len(g["selector"]) >= 6
    for c in g["selector"]:
        if "estimate" in c:
            est, cond = c["estimate"], c["condition"]
        else:
            est, cond = synthetic_partition_inputs(c["seed"], c["gh"], c["gw"], c["frac"])
        e, u, raw, final, sim = ro.select_tokens(est, cond, c["threshold"], c["gh"], c["gw"], c["erosion_dilation"])
        assert torch.equal(e.to(torch.int32), c["edited"]), f"edited ids differ for seed {c['seed']}"
        assert torch.equal(u.to(torch.int32), c["unedited"])
        assert e.shape[1] + u.shape[1] == c["gh"] * c["gw"]
        assert bool((e[:, 1:] > e[:, :-1]).all()) and bool((u[:, 1:] > u[:, :-1]).all())  # ascending


def test_empty_edited_set_edge_case(golden_dir):
    c = [c for c in _load(golden_dir, "region_ops.pt")["selector"] if c["frac"] == 0.0][0]
    assert c["edited"].shape == (1, 0) and c["unedited"].shape[1] == c["gh"] * c["gw"]


def test_morphology_matches_reference(golden_dir):
    for m in _load(golden_dir, "region_ops.pt")["morphology"]:
        assert torch.equal(ro.clean_mask(m["mask"].float()).to(torch.uint8), m["out"])
    # zero padding: an all-ones grid loses its border in the erosion
    er = ro.erode_cross3(torch.ones(6, 6))
    assert er[0].sum() == 0 and er[:, 0].sum() == 0 and er[1:-1, 1:-1].all()


def test_gather_scatter_match_reference(golden_dir):
    g = _load(golden_dir, "region_ops.pt")["gather"]
    ids = g["ids"].long()
    got = ro.gather_rows(g["latent"], ids)
    assert torch.equal(got, g["gathered"])
    assert torch.equal(ro.scatter_rows(got, ids, torch.zeros_like(g["latent"])), g["scattered"])


def test_gamma_tables_match_reference(golden_dir):
    g = _load(golden_dir, "schedule.pt")
    assert torch.equal(torch.tensor(sc.GAMMA["FluxKontext"], dtype=torch.float16), g["gamma"])


def test_avdc_plan_matches_reference_rules(golden_dir):
    g = _load(golden_dir, "schedule.pt")
    for p in g["plans"]:
        plan = sc.avdc_plan(p["timesteps"], sc.GAMMA["FluxKontext"], **p["params"])
        assert [s["mode"] for s in plan] == [s["mode"] for s in p["steps"]], p["params"]
        assert [s["write_cache"] for s in plan] == [s["write_cache"] for s in p["steps"]]
        for a, b in zip(plan, p["steps"]):
            if a["mode"] == "SKIP":
                assert a["ratio"] == b["ratio"]


def test_default_schedule_is_survey_appendix_a(golden_dir):
    sig, ts = sc.flow_match_sigmas(28, 4096)
    modes = [s["mode"] for s in sc.avdc_plan(ts, sc.GAMMA["FluxKontext"], cache_threshold=0.04)]
    assert [i for i, m in enumerate(modes) if m == "REGION"] == [6, 8, 11, 13, 23]
    assert [i for i, m in enumerate(modes) if m == "FULL"] == [0, 1, 2, 3, 4, 5, 15, 26, 27]
    modes = [s["mode"] for s in sc.avdc_plan(ts, sc.GAMMA["FluxKontext"], cache_threshold=0.01)]
    assert [i for i, m in enumerate(modes) if m == "REGION"] == [6, 8, 11, 13, 19, 21, 23, 25]
    modes = [s["mode"] for s in sc.avdc_plan(ts, sc.GAMMA["Step1XEdit"], cache_threshold=0.02)]
    assert [i for i, m in enumerate(modes) if m == "REGION"] == [6, 14, 19, 22, 24]
    assert abs(float(ts[5]) - 935.6) < 0.05 and abs(float(ts[15]) - 732.4) < 0.05


def test_region_state_parameter_validation():
    st = ro.RegionState()
    base = dict(num_inference_steps=28, warmup_step=6, post_step=2, threshold=0.88, cache_threshold=0.04,
                erosion_dilation=True)
    st.set_parameters(dict(base, refresh_step="16"))
    assert st.refresh_step == [16, 27]
    for bad in ("7", "26", "10,11"):
        try:
            st.set_parameters(dict(base, refresh_step=bad))
        except AssertionError:
            continue
        raise AssertionError(f"refresh_step={bad} should be rejected")
