"""The reference's CLI / timing harness (src/<Family>/main.py) over the B200 hot path with the synthetic model source:
demo mode and evaluation mode for every family, file layout and time_consuming.json format as the reference's."""
import json
import os

import pytest
import torch

from regione_b200 import cli

pytestmark = pytest.mark.gpu


def _demo_list(tmp_path, n=2):
    p = tmp_path / "data.jsonl"
    with open(p, "w") as f:
        for i in range(n):
            f.write(json.dumps({"instruction": f"edit {i}", "key": f"assets/demo_{i}"}) + "\n")
    return str(p)


@pytest.mark.parametrize("family", list(cli.FAMILIES))
def test_demo_mode_every_family(tmp_path, family, capsys):
    out = tmp_path / "out"
    rc = cli.main([family, "--use_regione", "--erosion_dilation", "--model_path", "synthetic:tiny",
                   "--image_path", _demo_list(tmp_path), "--output_dir", str(out)])
    assert rc == 0
    text = capsys.readouterr().out
    assert "Warmup..." in text and text.count("Time consuming:") == 2
    for i in range(2):
        lat = torch.load(out / f"demo_{i}.pt")
        assert lat.shape == (1, 256, 64) and lat.dtype == torch.bfloat16 and bool(torch.isfinite(lat.float()).all())
    assert not torch.equal(torch.load(out / "demo_0.pt"), torch.load(out / "demo_1.pt"))


def test_evaluation_mode_layout(tmp_path):
    bench = tmp_path / "bench"
    for task in ("color_alter", "text_change"):
        os.makedirs(bench / task / "img")
        with open(bench / task / "metadata.jsonl", "w") as f:
            for i in range(3):
                f.write(json.dumps({"key": f"{task}_{i}", "instruction": f"do {i}"}) + "\n")
    out = tmp_path / "result"
    rc = cli.main(["FluxKontext", "--use_regione", "--erosion_dilation", "--evaluation", "--no_warmup",
                   "--model_path", "synthetic:tiny", "--threshold", "0.88", "--image_path", str(bench),
                   "--output_dir", str(out)])
    assert rc == 0
    for task in ("color_alter", "text_change"):
        tc = json.load(open(out / task / "time_consuming.json"))
        assert set(tc) == {"num_item", "ave_time_consuming", "time_consuming_list"}      # metric_merge.py:36-62
        assert tc["num_item"] == 3 and len(tc["time_consuming_list"]) == 3
        assert abs(tc["ave_time_consuming"] - sum(tc["time_consuming_list"]) / 3) < 1e-9
        meta = json.load(open(out / task / "metadata.json"))
        assert meta == {f"{task}_{i}": f"do {i}" for i in range(3)}
        assert sorted(os.listdir(out / task / "generation")) == [f"{task}_{i}.pt" for i in range(3)]


def test_vanilla_loop_is_refused_for_the_synthetic_source(tmp_path):
    with pytest.raises(SystemExit, match="use_regione"):
        cli.main(["FluxKontext", "--model_path", "synthetic:tiny", "--image_path", _demo_list(tmp_path)])
