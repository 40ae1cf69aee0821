"""In-situ drop-in test (VERDICT r1, row b'): the UNMODIFIED reference -- `minisgl.llm.LLM` ->
`Scheduler` -> `Engine` -> `GraphRunner` -> `CacheManager` / radix cache, pip-installed into the
git-ignored oracle/_ref/minisgl_site -- drives `--attn b200` and, in separate processes, its own `fi` and
`trtllm` backends on the same requests (Qwen3-0.6B shape, dummy weights, greedy, teacher-forced),
see tools/insitu.py.  Skipped when that directory is absent (the reference cannot travel in git).

Gates:
  * structure: chunked prefill, radix-hit extends and padded CUDA-graph replays all happened, through
    the real scheduler, with the b200 kernels launched (launch counter of libb200attn.so);
  * indexing / append side effect, bit exact: layer-0 K and V rows of every appended token
    (prefill, extend and graph-replayed decode appends) equal the reference run's rows bit for bit
    -- those rows do not depend on attention, so any difference is an append / out_loc error;
  * logits: per row max|a-b| / max|b| (tools/insitu.logits_rel_err).  north_star's 1e-3 is stated
    for identical inputs to ONE attention call; here the perturbation passes 28 layers of a
    random-weight bf16 network, and the reference's own two backends (fi vs trtllm) differ by
    `trtllm_vs_fi` on the same run.  Gate: b200-vs-fi <= max(LOGITS_TOL, 2 x trtllm-vs-fi)
    (the amplification is chaotic: the same order of magnitude as the reference's own spread is what can be asked).
  * `patch_minisgl_layers` + `patch_minisgl_kernels` (our RMSNorm / RoPE / row-gather kernels inside
    the reference's model, bound before graph capture) and `install_into_minisgl` (the one-call form,
    under the reference's own fi backend): same logits gate, layer-0 K / V rows within one bf16 ulp of
    the FlashInfer-kernel runs (the kernels are bit-identical to FlashInfer on the golden shapes; at
    other row counts FlashInfer's CuTe-DSL norm uses a different fp32 reduction order).
"""
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
import insitu  # noqa: E402

pytestmark = pytest.mark.gpu

LOGITS_TOL = 1e-3


@pytest.fixture(scope="module")
def summary():
    if not insitu.reference_available():
        pytest.skip("reference not installed (pip install --no-deps --target oracle/_ref/minisgl_site <reference>)")
    import argparse

    out = ROOT / "gpurun_out" / "insitu_summary.json"
    return insitu.run(argparse.Namespace(out=str(out), layers=28))


def test_reference_scheduler_paths_were_exercised(summary):
    assert "errors" not in summary, summary.get("errors")
    s = summary["b200_vs_fi_page64"]
    assert s["chunked_reqs"] >= 2, s
    assert s["extend_reqs_with_cache_hit"] >= 2, s
    assert s["graph_replays"] >= 10, s
    assert s["b200_launches"] and s["b200_launches"] > 28 * s["forwards"] // 2, s
    # with the layers / kernels patched in, every norm, rope and row gather of the model is ours as well
    assert summary["b200_patched_vs_fi_page64"]["b200_launches"] > 3 * s["b200_launches"]


@pytest.mark.parametrize("pair", ["b200_vs_fi_page1", "b200_vs_fi_page64", "b200_vs_trtllm_page64"])
def test_appended_rows_bit_exact(summary, pair):
    """Layer-0 K / V rows do not depend on attention: with the reference's own norm / RoPE kernels in the
    model they must be bit-identical between the b200 run and the reference runs (append / out_loc parity)."""
    s = summary[pair]
    assert s["k_first_bit_exact"] and s["v_first_bit_exact"], s


@pytest.mark.parametrize("pair", ["b200_patched_vs_fi_page64", "fi_with_b200_layers_vs_fi_page64"])
def test_patched_layers_rows_within_one_ulp(summary, pair):
    """With OUR RMSNorm / RoPE / row-gather kernels inside the reference's model the layer-0 rows may differ
    from FlashInfer's by the fp32 reduction order only: one bf16 ulp (2^-7 relative)."""
    s = summary[pair]
    assert s["k_first_rel"] <= 2.0**-7 and s["v_first_rel"] <= 2.0**-7, s


@pytest.mark.parametrize("pair", ["b200_vs_fi_page1", "b200_vs_fi_page64", "b200_patched_vs_fi_page64",
                                  "fi_with_b200_layers_vs_fi_page64"])
def test_logits_match_reference_flashinfer_path(summary, pair):
    ref_spread = summary["trtllm_vs_fi_page64"]["logits_rel_worst"]
    got = summary[pair]["logits_rel_worst"]
    bound = max(LOGITS_TOL, 2.0 * ref_spread)
    print(json.dumps({"pair": pair, "rel": got, "reference_fi_vs_trtllm": ref_spread, "bound": bound}))
    assert got <= bound, (pair, got, bound)
