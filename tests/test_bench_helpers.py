"""CPU checks of the arithmetic behind bench.py's derived figures (no GPU, no timing): the helpers are fed the recorded
round-2 bench line (profiles/r2_bench_n1.json) and must reproduce / bound what it states."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_prove_roofline_on_the_recorded_line():
    import bench_extras
    line = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_n1.json")))
    pr = line["prove"]
    roof, pipe = bench_extras.prove_roofline(pr["op_counts"], 20, 22, int(pr["msm_tables"]), pr["value"], line["int_pipe"]["peak"],
                                             line["roofline"]["peak"], line["roofline"]["peak_source"])
    n, m = 1 << 20, 1 << 22
    cnt = pr["op_counts"]
    assert roof["algorithmic_bytes"] == cnt["msm_n"] * n * 96 + cnt["ifft_n"] * n * 64 + (cnt["coset_fft_4n"] + cnt["ifft_4n"]) * m * 64
    assert roof["bound"] == "hbm" and 0 < roof["frac"] < 0.05            # far from the HBM roof: the path is integer-pipe bound
    assert abs(roof["achieved"] - roof["algorithmic_bytes"] / pr["value"] / 1e9) < 1e-9
    # 15 windows of 17 bits, 6 products per addition; (m / 2) log2 m butterflies per transform
    msm = cnt["msm_n"] * n * 15 * 6
    ntt = cnt["ifft_n"] * (n // 2) * 20 + (cnt["coset_fft_4n"] + cnt["ifft_4n"]) * (m // 2) * 22
    assert abs(pipe["achieved"] - (msm + ntt) / pr["value"] / 1e9) < 1e-6
    assert 0 < pipe["frac"] < 1.0 and "UPPER" in pipe["note"]


def test_numa_placement_picks_the_gpu_local_cores(monkeypatch):
    """bench.py at N > 1 runs each rank on the cores NVML reports as local to its GPU: the mask-to-CPU-set logic with a fake NVML
    (half of the allowed cores), the refusals (no NVML, a mask equal to what is already allowed, the switch), affinity restored."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    allowed = sorted(os.sched_getaffinity(0))
    if len(allowed) < 4:
        import pytest
        pytest.skip("needs four cores to split")
    half = allowed[:len(allowed) // 2]
    words = (max(os.cpu_count() or 1, 1) + 63) // 64

    def mask_of(cpus):
        out = [0] * words
        for c in cpus:
            out[c // 64] |= 1 << (c % 64)
        return out

    state = {"mask": mask_of(half)}
    fake = types.SimpleNamespace(nvmlInit=lambda: None, nvmlDeviceGetHandleByPciBusId=lambda s: ("h", s), nvmlDeviceGetHandleByIndex=lambda i: ("h", i),
                                 nvmlDeviceGetCpuAffinity=lambda h, n: state["mask"][:n])
    props = types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1b, pci_device_id=0)
    torch = types.SimpleNamespace(cuda=types.SimpleNamespace(get_device_properties=lambda i: props))
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    try:
        assert bench.bind_to_gpu_numa(torch, 0) == half and sorted(os.sched_getaffinity(0)) == half
        os.sched_setaffinity(0, allowed)
        state["mask"] = mask_of(allowed)                      # one memory node: nothing to do
        assert bench.bind_to_gpu_numa(torch, 0) is None and sorted(os.sched_getaffinity(0)) == allowed
        state["mask"] = mask_of(half)
        monkeypatch.setenv("H2A_BENCH_NUMA", "0")
        assert bench.bind_to_gpu_numa(torch, 0) is None
        monkeypatch.delenv("H2A_BENCH_NUMA")
        fake.nvmlInit = lambda: (_ for _ in ()).throw(RuntimeError("no NVML"))
        assert bench.bind_to_gpu_numa(torch, 0) is None and sorted(os.sched_getaffinity(0)) == allowed
    finally:
        os.sched_setaffinity(0, allowed)
