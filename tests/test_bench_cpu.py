"""bench.py host logic that needs no GPU: the roofline's `traffic` comes from the committed ncu launch list of the same forward
(profiles/r02_traffic_*.json, made by tools/gpu_launchlists.sh + tools/dram_summary.py), keyed by workload and kernel."""
import importlib.util
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec = importlib.util.spec_from_file_location("fmi_bench_under_test", ROOT / "bench.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.argv = argv


def test_ncu_traffic_lookup():
    b = _bench()
    t = b.ncu_traffic("picnet_ref", "conv_gemm")
    assert t and t["bytes_per_step"] == t["read_bytes_per_step"] + t["write_bytes_per_step"] > 1e9 and t["launches_per_step"] == 79
    assert (ROOT / t["source"]).is_file()
    ir, dec = b.ncu_traffic("refpsp", "conv_gemm_ir"), b.ncu_traffic("refpsp", "conv_gemm")
    assert ir and dec and ir["launches_per_step"] == 71 and dec["launches_per_step"] == 35       # trunk + heads vs StyleGAN2 decoder
    assert b.ncu_traffic("train_picnet", "attn_bwd") is None and b.ncu_traffic("picnet_ref", "no_such_kernel") is None


def test_committed_bench_record_is_on_the_metric():
    """The committed closing record: headline = configs[0] (PICNet-ref 256^2, batch 4) with roofline + traffic + e2e + cpu_baseline,
    the other three configs as full records."""
    line = [ln for ln in (ROOT / "profiles" / "r02_bench_n1_final3.json").read_text().splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["config"]["name"] == "picnet_ref" and d["config"]["per_gpu_batch"] == 4 and d["unit"] == "img/s" and d["n_gpus"] == 1
    rf = d["roofline"]
    assert rf["bound"] in ("hbm", "tensor") and 0 < rf["frac"] < 1 and rf["traffic"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert [r["config"]["name"] for r in d["records"]] == ["refpsp", "train_picnet", "train_psp"]
    assert not d["clocks"]["reasons"]
