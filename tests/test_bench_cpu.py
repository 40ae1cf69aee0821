"""CPU checks of bench.py's workload bookkeeping: the schedule must be BASELINE.json's cfg1 (the
reference's benchmark/offline/bench.py RNG replayed), the algorithmic byte / flop formulas are
SURVEY.md 8(d)'s, and the `--impl reference` leg prints the contract's JSON line without a GPU."""
import json
import subprocess
import sys
from pathlib import Path

import bench

ROOT = Path(__file__).resolve().parent.parent


def test_schedule_is_cfg1():
    s = bench.Schedule()
    # SURVEY.md 8(d) "derived exactly (replayed RNG)"
    assert len(s.in_lens) == 256 and len(s.out_lens) == 256
    assert sum(s.in_lens) == 142827 and min(s.in_lens) == 107 and max(s.in_lens) == 1024
    assert sum(s.out_lens) == 133966 and min(s.out_lens) == 103 and max(s.out_lens) == 1024
    assert s.n_iters == 1023
    assert s.decode_token_steps == 133710
    assert s.sum_kv == 120795204
    # every request decodes out-1 tokens (the first output token comes from the prefill)
    assert s.decode_token_steps == sum(o - 1 for o in s.out_lens)
    assert sum(len(s.live(i)) for i in range(s.n_iters)) == s.decode_token_steps
    assert sum(d for i in range(s.n_iters) for (_, _, d) in s.live(i)) == s.sum_kv


def test_live_sets_shrink_and_lengths_grow():
    s = bench.Schedule()
    first, mid, last = s.live(0), s.live(500), s.live(s.n_iters - 1)
    assert len(first) == 256 and len(first) >= len(mid) >= len(last) >= 1
    for (t, c, d) in mid:
        assert d == s.in_lens[t] + 501 and c == d - 1 and 500 < s.out_lens[t] - 1
    its = s.sample_iters(40)
    assert len(its) == 40 and its == sorted(its) and 0 <= its[0] and its[-1] < s.n_iters


def test_prefill_batches_respect_the_token_budget():
    """Greedy admission with chunk splitting (reference scheduler/prefill.py:64-90): every batch spends at
    most max_extend_tokens, a prompt straddling the budget continues as a cached_len > 0 chunk in the
    next batch, and the chunks of a request tile [0, in_len) exactly."""
    for name in ("cfg1", "cfg2", "cfg4"):
        wl = bench.WORKLOADS[name]
        s = bench.Schedule(wl)
        batches = s.prefill_batches()
        covered = {}
        for b in batches:
            assert 0 < sum(d - c for (_, c, d) in b) <= wl.max_extend
            for (t, c, d) in b:
                start = min(wl.shared_prefix, (s.in_lens[t] - 1) // 64 * 64, (s.in_lens[0] - 1) // 64 * 64) if wl.shared_prefix else 0
                assert c == covered.get(t, start) and c < d <= s.in_lens[t]
                covered[t] = d
        assert covered == {t: n for t, n in enumerate(s.in_lens)}
        assert [len(b) for b in batches[:-1]] and all(sum(d - c for (_, c, d) in b) == wl.max_extend for b in batches[:-1])
    s = bench.Schedule(bench.WORKLOADS["cfg4"])
    assert s.sum_kv == 68935680 and s.n_iters == 255


def test_algorithmic_work_formulas():
    tr = [(0, 99, 100), (1, 511, 512), (2, 0, 1)]
    kv = 100 + 512 + 1
    row = 2 * 8 * 128 * 2  # K and V rows of one token, 8 kv heads, bf16
    assert bench.decode_bytes_per_layer(tr, 16, 8) == kv * row + 3 * row + 3 * 2 * 16 * 128 * 2
    # tp=8 shard: one kv head, two q heads
    assert bench.decode_bytes_per_layer(tr, 2, 1) == kv * 512 + 3 * 512 + 3 * 2 * 2 * 128 * 2
    # causal flops: 4*Hq*D*sum(q*cached + q(q+1)/2)
    ptr = [(0, 0, 150), (1, 64, 200)]
    want = 4 * 16 * 128 * (150 * 151 // 2 + 136 * 64 + 136 * 137 // 2)
    assert bench.prefill_flops_per_layer(ptr, 16) == want
    nnz, cached = 150 + 136, 64
    assert bench.prefill_bytes_per_layer(ptr, 16, 8) == nnz * 2 * 16 * 256 + nnz * 2 * 2 * 8 * 256 + cached * 2 * 8 * 256
    s = bench.Schedule()
    total = sum(bench.prefill_flops_per_layer(b, 16) for b in s.prefill_batches())
    assert abs(total - 4.03e11) / 4.03e11 < 5e-3  # "cfg1 no-cache = 4.03e11 / layer"
    # whole decode job: sum_kv * 4096 B per layer dominates
    assert s.sum_kv * 4096 * bench.L == 120795204 * 114688


def test_graph_batch_sizes_match_the_reference_engine():
    bs = bench.graph_bs_list(256)  # engine/graph.py:67: [1, 2, 4] + range(8, max+1, 8)
    assert bs[:4] == [1, 2, 4, 8] and bs[-1] == 256 and all(b % 8 == 0 for b in bs[3:])
    assert bench.effective_cpus() >= 1


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-reqs", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # nothing extrapolated: the reported step time is the measured one, tokens = requests actually attended
    assert abs(d["value"] - 1 / (d["ms_per_step"] * 1e-3)) / d["value"] < 0.02
    assert d["config"]["page_size"] == 64 and d["config"]["layers"] == 28 and "ALL 28 layers" in d["config"]["sample"]
