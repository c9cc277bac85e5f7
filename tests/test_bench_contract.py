"""bench.py contract pieces that do not need a GPU: workload naming, algorithmic bytes per launch class and the
roofline object assembled from per-op CUDA-event times (keys the driver reads)."""
import argparse
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("pcb_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def _args(**kw):
    d = dict(mode=None, config="default", batch=0, volume=0, steps=10, warmup=3, gpus=1, sw_batch=4)
    d.update(kw)
    return argparse.Namespace(**d)


def test_algorithmic_bytes_per_launch_class():
    v, c = 4096000, 32
    assert bench._op_bytes("mlp_fwd", f"mlp_fwd:m0C{c}H64Co32V{v}", 4) == 4 * v * 2 * (c + 2 * 32)       # y + residual in, out
    assert bench._op_bytes("mlp_bwd_fused", f"mlp_bwd_fused:m0C{c}H64Co32V{v}", 4) == 4 * v * 2 * (2 * c + 32)
    assert bench._op_bytes("dwconv_fwd", f"dwconv_fwd:m0C{c}V{v}", 1) == v * 2 * 2 * c
    assert bench._op_bytes("dw_bwd_data", f"dw_bwd_data:m0C{c}V{v}", 1) == v * 2 * 3 * c                 # + fused residual gradient
    assert bench._op_bytes("dw_bwd_data", f"dw_bwd_data:m2C64V{v}", 1) == v * 2 * 2 * 64
    assert bench._op_bytes("gn_bwd", f"gn_bwd:C{c}V{v}", 2) == 2 * v * 2 * 3 * c
    assert bench._op_bytes("tn_gemm", "tn_gemm:M32N64V512000", 1) is None


def test_roofline_object_has_the_contract_keys():
    a = _args()
    key = "mlp_bwd_fused:m0C32H64Co32V4096000"
    prof = {key: [3.0] * 12, "dwconv_fwd:m0C32V4096000": [1.4] * 12, "tn_gemm:M32N64V512000": [1.0] * 3,
            "gn_bwd:C32V4096000": [0.6] * 12}
    roof = bench.build_roofline(prof, 4, 6554.9, True, 100.0, 3)
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "launch_class", "others", "peak_source"):
        assert k in roof
    assert roof["launch_class"] == key and roof["bound"] == "hbm" and roof["unit"] == "GB/s"
    nbytes = 4 * 4096000 * 2 * (2 * 32 + 32)
    assert abs(roof["achieved"] - nbytes / 3.0e-3 / 1e9) < 1e-6 and abs(roof["frac"] - roof["achieved"] / 6554.9) < 1e-12
    assert roof["traffic"] == bench.NCU_TRAFFIC[key] * 4                      # per-sample ncu capture x batch
    assert [r["launch_class"].split(":")[0] for r in roof["others"]] == ["dwconv_fwd", "gn_bwd", "tn_gemm"]
    assert bench.build_roofline({}, 4, 6554.9, True, 100.0, 3) is None
    # inference leg: samples per launch = the engine's sw_batch (round-1 bug: a hard-coded 2 understated frac 2x)
    r2 = bench.build_roofline({"mlp_fwd:m0C32H64Co32V4096000": [1.2] * 10}, a.sw_batch, 6554.9, True, 50.0, 1)
    assert r2["samples_per_launch"] == 4 and r2["algorithmic_bytes"] == 4 * 4096000 * 2 * (32 + 2 * 32)


def test_workload_config_names_the_workload():
    c = bench.train_config(_args(), "c2", 2)
    assert c["global_batch"] == 8 and c["parallelism"] == "dp2" and "BASELINE configs[1]" in c["workload"]
    c3 = bench.train_config(_args(), "c3", 8)
    assert "3-channel" in c3["workload"] and c3["global_batch"] == 32 and c3["crop"] == [160] * 3
    c4 = bench.train_config(_args(), "c4", 8)
    assert "MedNeXt-L" in c4["workload"] and c4["crop"] == [224] * 3 and c4["global_batch"] == 8
    ci = bench.infer_config(_args(), 2)
    assert ci["volume"] == [1280, 640, 640] and "z-slab" in ci["parallelism"] and ci["sw_batch"] == 4
    c5 = bench.infer_config(_args(config="c5"), 8)
    assert c5["volume"] == [2048] * 3 and "BASELINE configs[4]" in c5["workload"]
    assert all("model" not in d for d in (c, c3, c4, ci, c5))
    assert bench.tiles_per_axis(480) == 5 and bench.tiles_per_axis(2048) == 25 and bench.tiles_per_axis(160) == 1
    assert bench.tiles_per_axis(640) == 7 and bench.tiles_per_axis(200) == 2


def test_reference_arm_prints_one_json_line():
    # the reference arm runs on host cores only: one REAL 160^3 tile forward per step (the inference config; the training
    # config's real 160^3 fp32 step takes minutes on the 8 build-container cores and is exercised on the GPU box)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c5",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC_INFER and d["unit"] == "Mvox/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert "160^3" in d["cpu_baseline"]["sample"] and "extrapol" not in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_cpu_train_sample_is_a_real_crop_of_the_config():
    # small side only to keep the CPU suite fast; the default (side=0) is the config's own crop size (160 / 224)
    r = bench.cpu_train_rate(1, 0, "c3", side=32)
    assert r["unit"] == "sub-volumes/s" and r["value"] > 0 and "1x32^3" in r["sample"] and "no size extrapolation" in r["sample"]


def test_step_roofline_arithmetic():
    # 4 sub-volumes per step at the measured peaks of this pod: 3 x 5.49 GB / 6554.9 GB/s = 2.51 ms each (HBM-bound)
    r = bench.step_roofline("train", 4, 91.4, 6554.9, 1366.2)
    assert r["bound"] == "hbm" and abs(r["roofline_ms_per_step"] - 4 * 3 * 5.49e9 / 6554.9e9 * 1e3) < 1e-9
    assert abs(r["frac"] - r["roofline_ms_per_step"] / 91.4) < 1e-12 and 0.10 < r["frac"] < 0.12
    ri = bench.step_roofline("infer", 125, 750.0, 6554.9, 1366.2)       # 480^3 volume: 5^3 tiles
    assert abs(ri["roofline_ms_per_step"] - 125 * 5.49e9 / 6554.9e9 * 1e3) < 1e-9
    rl = bench.step_roofline("train", 1, 100.0, 6554.9, 1366.2, bench.FWD_WORK[("L", 224)], "224^3")   # C4: tensor-bound
    assert rl["bound"] == "tensor" and abs(rl["roofline_ms_per_step"] - 3 * 5416e9 / 1366.2e12 * 1e3) < 1e-9
