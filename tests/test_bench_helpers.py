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
