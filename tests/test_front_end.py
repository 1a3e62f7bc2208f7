"""Host functions either side of the loop against fixtures produced by the reference's own
code (oracle/make_golden.make_front_end -> tests/golden/front_end.json)."""
import argparse
import json
import os

import numpy as np
import pytest
import torch

from regione_b200 import cli, params, schedule
from standins.diffusers_like import FlowMatchEulerDiscreteScheduler


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "front_end.json")) as f:
        return json.load(f)


def test_calculate_shift_bit_exact(golden):
    from oracle import schedule as osc
    for n, hx in golden["calculate_shift"].items():
        want = float.fromhex(hx)
        assert schedule.calculate_shift(int(n)) == want, n
        assert osc.calculate_shift(int(n)) == want, n
    assert schedule.calculate_shift(4050, 256, 8192, 0.5, 0.9) == float.fromhex(golden["calculate_shift_custom"])


def test_gamma_tables_and_defaults_of_every_family(golden):
    assert set(golden["gamma"]) == set(params.GAMMA) == set(params.DEFAULTS)
    for name, table in golden["gamma"].items():
        ours = torch.tensor(params.GAMMA[name], dtype=torch.float16)        # the reference keeps gamma in fp16
        assert torch.equal(ours, torch.tensor(table, dtype=torch.float16)), name
        assert len(table) == 27
    assert params.DEFAULTS == golden["defaults"]


@pytest.mark.parametrize("family", list(cli.FAMILIES))
def test_cli_flags_match_reference_main(golden, family):
    want = golden["cli"][family]
    got = {}
    for a in cli.build_parser(family)._actions:
        if a.dest != "help":
            got[a.dest] = dict(default=a.default, type=getattr(a.type, "__name__", None),
                               flag=isinstance(a, argparse._StoreTrueAction))
    for dest, spec in want.items():
        assert dest in got, f"{family}: missing --{dest}"
        assert got[dest] == spec, f"{family}: --{dest} {got[dest]} != {spec}"
    assert set(got) - set(want) == {"grid", "txt_len", "rho", "no_warmup"}     # synthetic-source extras only


def test_cli_refuses_cpu_and_unknown_family(capsys):
    assert cli.main([]) == 2
    assert cli.main(["NoSuchFamily"]) == 2
    if not torch.cuda.is_available():
        with pytest.raises(SystemExit, match="CUDA-only"):
            cli.main(["FluxKontext", "--use_regione", "--model_path", "synthetic:tiny"])


def test_retrieve_timesteps_semantics():
    sch = FlowMatchEulerDiscreteScheduler()
    sig = np.linspace(1.0, 1 / 28, 28)
    mu = schedule.calculate_shift(4096)
    ts, n = schedule.retrieve_timesteps(sch, 28, "cpu", sigmas=sig, mu=mu)
    assert n == 28 and ts is sch.timesteps and sch.sigmas.shape[0] == 29 and float(sch.sigmas[-1]) == 0.0
    from oracle.schedule import flow_match_sigmas
    o_sig, o_ts = flow_match_sigmas(28, 4096)
    assert torch.equal(sch.sigmas, o_sig) and torch.equal(ts, o_ts)
    with pytest.raises(ValueError, match="Only one of"):
        schedule.retrieve_timesteps(sch, 28, "cpu", timesteps=[1, 2], sigmas=sig, mu=mu)

    class NoCustom:
        def set_timesteps(self, num_inference_steps, device=None):
            self.timesteps = torch.arange(num_inference_steps)

    with pytest.raises(ValueError, match="custom sigmas"):
        schedule.retrieve_timesteps(NoCustom(), 28, "cpu", sigmas=sig)
    with pytest.raises(ValueError, match="custom timestep"):
        schedule.retrieve_timesteps(NoCustom(), None, "cpu", timesteps=[3, 2, 1])
    ts, n = schedule.retrieve_timesteps(NoCustom(), 5, "cpu")
    assert n == 5 and len(ts) == 5


def test_patched_scheduler_keeps_an_inspectable_signature():
    """retrieve_timesteps looks for `sigmas` in set_timesteps' signature (utils.py:91): the RegionE scheduler class
    built by warp_modules must not hide it behind *args."""
    import inspect
    from regione_b200.flux_kontext import RegionESchedulerMixin
    cls = type("S", (RegionESchedulerMixin, FlowMatchEulerDiscreteScheduler), {})
    s = cls.from_config(FlowMatchEulerDiscreteScheduler().config)
    assert {"sigmas", "mu", "timesteps"} <= set(inspect.signature(s.set_timesteps).parameters)
    schedule.retrieve_timesteps(s, 28, "cpu", sigmas=np.linspace(1.0, 1 / 28, 28), mu=1.15)
    assert s.sigmas.shape[0] == 29


def test_avdc_plans_of_every_family_match_their_own_loop_code(golden_dir):
    """tests/golden/family_schedules.json: each family's AVDC block + gamma table exec'd from its inplace.py
    (oracle/make_golden.make_family_schedules). The product's host planner and the oracle's must take the same
    compute / skip decisions and reuse ratios, bit for bit, for all five pipelines."""
    from oracle.schedule import avdc_plan, flow_match_sigmas
    from regione_b200.manager import RegionManager, plan_steps
    with open(os.path.join(golden_dir, "family_schedules.json")) as f:
        golden = json.load(f)
    assert set(golden) == set(params.GAMMA)
    _, ts = flow_match_sigmas(28, 4096)
    for name, plans in golden.items():
        assert len(plans) == 3
        for p in plans:
            m = RegionManager()
            m.set_parameters(p["params"])
            plan = plan_steps(ts, params.GAMMA[name], m)
            assert [bool(s) for s, _ in plan] == [st["skip"] for st in p["steps"]], (name, p["params"])
            for (skip, ratio), st in zip(plan, p["steps"]):
                if skip:
                    assert float(ratio) == st["ratio"], (name, p["params"])
            prm = p["params"]
            o = avdc_plan(ts, params.GAMMA[name], warmup_step=prm["warmup_step"], post_step=prm["post_step"],
                          refresh_step=prm["refresh_step"], cache_threshold=prm["cache_threshold"])
            assert [s["mode"] == "SKIP" for s in o] == [st["skip"] for st in p["steps"]], (name, p["params"])


def test_cli_sharding_and_rho_sweep(monkeypatch):
    items = list(range(10))
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    assert cli._shard(items) == items
    monkeypatch.setenv("WORLD_SIZE", "4")
    got = []
    for r in range(4):
        monkeypatch.setenv("RANK", str(r))
        part = cli._shard(items)
        assert part == items[r::4]
        got += part
    assert sorted(got) == items                                   # every item exactly once, no collective involved
    a = cli.build_parser("Step1X-Edit-v1p2").parse_args(["--rho", "sweep"])
    assert {cli._rho(a, s) for s in range(50)} == set(cli.RHO_SWEEP)
    assert cli._rho(cli.build_parser("FluxKontext").parse_args(["--rho", "0.4"]), 3) == 0.4
