"""bench.py's output contract, checked without a GPU: the reference arm's JSON line (with the CPU timing stubbed), the
clock-sample reduction and the peak lookup."""
import argparse
import io
import json
import os
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e"}


def _args(**kw):
    d = dict(workload="sdxl", gpus=1, steps=3, warmup=3, impl="reference", batch=1, denoise_steps=50, guidance_scale=7.5,
             no_cpu_baseline=False, no_config3=False, no_reference_cuda=False, cpu_sample=False)
    d.update(kw)
    return argparse.Namespace(**d)


def test_reference_arm_line(monkeypatch):
    stub = dict(value=3.2e-4, unit="images/s", cores=64, kind="reference", sample="stub", dtype="bf16", t_step_s=6.0, t_step_runs_s=[6.1, 5.9],
                run_spread=0.03, setup_s=1.0)
    monkeypatch.setattr(bench, "cpu_reference", lambda args, full=True: stub)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(_args(gpus=2), rank=1, world=2)  # only rank 0 reports
    assert buf.getvalue() == ""
    with redirect_stdout(buf):
        bench.run_reference(_args(gpus=2), rank=0, world=2)
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= d.keys()
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"].startswith("images/sec") and d["unit"] == "images/s"
    assert d["config"]["workload"] == bench.sdxl_workload(_args()) == "sdxl_unet_1024_50step_cfg7.5_b1_per_gpu+vae_decode"
    assert {k: d["cpu_baseline"][k] for k in ("value", "unit", "cores", "kind", "sample")} == {k: stub[k] for k in ("value", "unit", "cores", "kind", "sample")}
    assert d["dtype"] == "bf16"  # the arm reports the dtype it actually ran
    assert d["e2e"] == dict(value=stub["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d["gpu_launches"] == 0 and abs(d["ms_per_step"] - 1000.0 / stub["value"]) < 1e-6


def test_clock_sampler_summary():
    cs = bench.ClockSampler(0)
    cs.rows = [["1965", "1965", "900.1", "Not Active", "Not Active", "Not Active", "Active"],
               ["1575", "1965", "990.0", "Not Active", "Not Active", "Not Active", "Not Active"],
               ["1800", "1965", "950.0", "Not Active", "Not Active", "Not Active", "Not Active"],
               ["garbage"]]
    s = cs.summary()
    assert s["sm_mhz"] == 1800.0 and s["sm_max_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"] and s["samples"] == 3
    assert bench.ClockSampler(0).summary() == dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)


def test_peaks_source(tmp_path, monkeypatch):
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    p = bench.peaks()
    assert p["source"] == "fallback" and p["tflops"] > 1000 and p["hbm_gbs"] > 5000
    with open(tmp_path / "MEASURED_PEAKS.json", "w") as f:
        json.dump(dict(hbm_gbs=6500.0, bf16_tflops=1600.0, bf16_tflops_sustained=1400.0), f)
    p = bench.peaks()
    assert p == dict(hbm_gbs=6500.0, tflops=1400.0, burst=1600.0, source="measured")


def test_flop_constants_match_baseline():
    # 2 x 50 x 6.7612 + 10.470 TFLOP per image; 74.385 TFLOP per Flux forward
    assert abs((2 * 50 * bench.UNET_FLOP_PER_SAMPLE + bench.VAE_FLOP_PER_IMAGE) / 1e12 - 686.6) < 0.1
    assert abs(28 * bench.FLUX_FLOP_PER_FORWARD / 1e12 - 2082.8) < 0.1


def test_one_numa_node_physical_cores_is_a_subset_of_the_affinity_mask():
    cores = bench.one_numa_node_physical_cores()
    assert cores and set(cores) <= set(os.sched_getaffinity(0)) and len(set(cores)) == len(cores)


def test_context_parallel_section_runs_last_and_lands_in_the_flux_section(monkeypatch):
    """N > 1: the Flux context-parallel measurement is deferred to the end of run_b200 (under its watchdog) and its result - or
    its failure string - ends up in line["flux"]["context_parallel"] of the one JSON line; ranks other than 0 print nothing."""
    class FakeCtx:
        def __init__(self, args, rank, world, local_rank):
            self.args, self.rank, self.world, self.gpu_dead, self.flux_state = args, rank, world, False, None

    order = []

    def fake_flux(cx, standalone=False):
        order.append("flux")
        cx.flux_state = ("pipe", "tr", "res", "call", 28, 1900.0)
        return dict(metric="latents/sec", value=1.0)

    def fake_cp(cx, pipe, tr, res, call, nsteps, replica_ms):
        order.append("cp")
        assert (pipe, tr, res, call, nsteps, replica_ms) == ("pipe", "tr", "res", "call", 28, 1900.0)
        return dict(ulysses_degree=cx.world, value=0.8)

    monkeypatch.setattr(bench, "Ctx", FakeCtx)
    monkeypatch.setattr(bench, "sdxl_section", lambda cx: order.append("sdxl") or dict(metric="images/sec", value=2.0))
    monkeypatch.setattr(bench, "flux_section", fake_flux)
    monkeypatch.setattr(bench, "flux_context_parallel", fake_cp)
    out = []
    monkeypatch.setattr(bench, "emit", lambda line: out.append(json.loads(json.dumps(line))))
    a = _args(workload="all", gpus=2, impl="b200", no_context_parallel=False)
    bench.run_b200(a, 0, 2, 0)
    assert order == ["sdxl", "flux", "cp"] and len(out) == 1
    assert out[0]["flux"]["context_parallel"] == dict(ulysses_degree=2, value=0.8) and out[0]["value"] == 2.0
    order.clear()
    bench.run_b200(a, 1, 2, 1)
    assert order == ["sdxl", "flux", "cp"] and len(out) == 1  # rank 1 measures too, prints nothing
    order.clear()
    bench.run_b200(_args(workload="flux", gpus=2, impl="b200", no_context_parallel=False), 0, 2, 0)
    assert order == ["flux", "cp"] and out[1]["context_parallel"]["value"] == 0.8
