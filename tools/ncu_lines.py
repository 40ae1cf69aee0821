#!/usr/bin/env python
"""Attribute the warp-stall samples of an Nsight Compute report to CUDA source lines, offline.

`ncu --page source --csv` only prints the SASS view; this joins it, instruction by instruction, with
`nvdisasm -gi` of the same kernel from the built library (compiled with -lineinfo), follows the
"inlined at" chain up to the kernel's own file, and prints the hottest lines with their top stall
reasons.  The library must be the build that was profiled.

    python tools/ncu_lines.py gpurun_out/prefill.ncu-rep attn_prefill_tc.cu 'attn_prefill_tc_kernelI13__nv_bfloat16' [--top 40]
"""
import argparse
import collections
import csv
import io
import re
import subprocess
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def disassemble(lib: Path, src_name: str, mangled_substr: str):
    """[(sass_text, [(file, line) innermost .. outermost])] for the kernel, in address order."""
    with tempfile.TemporaryDirectory() as td:
        if str(lib).endswith(".cubin"):  # e.g. an older revision compiled with `nvcc -cubin -lineinfo`
            cubin = lib
        else:
            subprocess.run(["cuobjdump", "-xelf", "all", str(lib)], cwd=td, check=True, capture_output=True)
            stem = src_name.rsplit(".", 1)[0]
            cubin = next(p for p in Path(td).glob("*.cubin") if p.name.startswith(stem + "."))
        text = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], check=True, capture_output=True, text=True).stdout
    lines = text.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and mangled_substr in l)
    out, chain = [], []
    fresh = True  # the next "//## File" line starts a new chain
    for l in lines[start + 1:]:
        if l.startswith("\t.section") or l.startswith(".text."):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain = []
                fresh = False
            chain.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m:
            out.append((m.group(1).strip(), list(chain)))
            fresh = True
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("source", help="kernel source file name, e.g. attn_prefill_tc.cu")
    ap.add_argument("kernel", help="substring of the mangled kernel name")
    ap.add_argument("--lib", default=str(ROOT / "mini-sglang_b200" / "libb200attn.so"))
    ap.add_argument("--src", default=None, help="source file to quote lines from (default: csrc/<source>)")
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    sass = disassemble(Path(args.lib), args.source, args.kernel)
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    i_src, i_samp = hdr.index("Source"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[2:] if len(r) > i_samp and r[i_samp].isdigit()]
    if len(data) != len(sass):
        raise SystemExit(f"instruction count differs (report {len(data)}, library {len(sass)}): not the profiled build")
    total = sum(int(r[i_samp]) for r in data)
    by_line = collections.Counter()
    reasons = collections.defaultdict(collections.Counter)
    for r, (_, chain) in zip(data, sass):
        own = [c for c in chain if c[0] == args.source]
        key = own[-1] if own else (chain[-1] if chain else ("?", 0))
        inner = chain[0] if chain and chain[0] != key else None
        by_line[(key, inner)] += int(r[i_samp])
        for c in stall_cols:
            v = int(r[c]) if r[c].isdigit() else 0
            if v:
                reasons[(key, inner)][hdr[c][6:]] += v
    src = Path(args.src or (ROOT / "mini-sglang_b200" / "csrc" / args.source)).read_text().splitlines()
    print(f"total samples {total}, {len(data)} instructions")
    for (key, inner), s in by_line.most_common(args.top):
        f, ln = key
        text = src[ln - 1].strip()[:80] if f == args.source and 0 < ln <= len(src) else ""
        why = ", ".join(f"{k}:{v}" for k, v in reasons[(key, inner)].most_common(3))
        via = f" via {inner[0]}:{inner[1]}" if inner else ""
        print(f"{s:6d} {s / total:6.3f}  {f}:{ln}{via}  [{why}]  {text}")


if __name__ == "__main__":
    main()
