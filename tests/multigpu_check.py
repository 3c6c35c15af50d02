"""N > 1 on hardware: the product CLI with `--gpus N` (one process, N
contexts, chunk k -> GPU k mod N, `mgk_allreduce_counters` = ncclCommInitAll + ncclAllReduce,
the reference's ReportManager::update src/report_manager.rs:377-435 / src/lib.rs:819-827)
must write the very files one GPU writes: every FASTQ (decompressed) and every report
equal to the goldens the reference itself produced.  Used by tests/test_multi_gpu.py
and, so that the scaling run executes it too, by bench.py's rank 0 at N = 2."""
from __future__ import annotations

import json
import os
import shutil
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def cli_matches_goldens(binary: str, n_gpus: int) -> dict:
    """Returns {case: files compared}; raises AssertionError on the first difference."""
    import lanes
    import ref_matrix as M
    from golden.make_golden import synthetic_variants
    done = {}
    with open(os.path.join(HERE, "golden", "synthetic_reference.json")) as f:
        gold = json.load(f)
    name, lane, extra, n = [v for v in synthetic_variants() if v[0] == "cfg5"][0]
    d = tempfile.mkdtemp(prefix="mgk_ngpu_")
    try:
        # BASELINE config 5 (384 samples): 20 000 pairs in chunks of 1 MiB -> ~8 chunks per read file, round-robin
        r1, r2, sheet = lanes.write_lane(lane, n, d)
        out = os.path.join(d, "out")
        lanes.run_binary(binary, lanes.demux_args(lane, r1, r2, sheet, out, list(extra) + ["--gpus", str(n_gpus), "--chunk-size", "1048576"]))
        got, want = lanes.digest_dir(out), gold[name]["files"]
        assert sorted(got) == sorted(want), f"{name} on {n_gpus} GPUs: file sets differ: {sorted(set(got) ^ set(want))[:6]}"
        diff = [k for k in want if got[k] != want[k]]
        assert not diff, f"{name} on {n_gpus} GPUs: {len(diff)} files differ from the reference's, e.g. {diff[:4]}"
        done[name] = len(want)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    # the reference's own large_ds fixture (60 000 pairs) in 256 KiB chunks: ~90 chunks alternate between the GPUs
    with open(os.path.join(HERE, "golden", "reference_matrix.json")) as f:
        matrix = json.load(f)
    case = [c for c in M.all_cases() if c.name == "large-pe"][0]
    out = tempfile.mkdtemp(prefix="mgk_ngpu_")
    shutil.rmtree(out)
    try:
        p = M.run_case(binary, case, os.path.join(HERE, "golden", "testing_data"), out, ("--gpus", str(n_gpus), "--chunk-size", "262144"))
        assert p.returncode == 0, f"large-pe on {n_gpus} GPUs: exit {p.returncode}\n{p.stderr[-600:]}"
        entry = matrix[case.name]
        M.compare_outputs(case, {k: tuple(v) for k, v in entry["files"].items()}, out, entry.get("stats_override", ""))
        done[case.name] = len(entry["files"])
    finally:
        shutil.rmtree(out, ignore_errors=True)
    return done
