"""bench.py, the parts that run without a GPU: the reference arm prints the contract's JSON line, the file-level leg
(tools/file_bench.py) runs end to end with the oracle-bound CLI standing in for the product."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "mgikit")


def test_reference_arm_prints_the_contract_line(oracle_cli):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-pairs", "4000"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "read_pairs_per_second_demultiplexed" and d["unit"] == "read pairs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference binary not available")
def test_file_level_leg_runs_with_the_oracle_cli(oracle_cli):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import file_bench
    from mgikit_b200 import synth
    old = file_bench.BLOCK_PAIRS
    file_bench.BLOCK_PAIRS = 10000
    try:
        r = file_bench.run(synth.CONFIG2, 30000, oracle_cli, REF_BIN, prefix_pairs=20000)
    finally:
        file_bench.BLOCK_PAIRS = old
    assert r["pairs"] == 30000 and r["ours"]["pairs_per_s"] > 0 and r["reference"]["pairs_per_s"] > 0
    assert "byte-identical" in r["parity"]["reports_full_run"] and "identical" in r["parity"]["fastq_prefix"]
    assert "error" not in r["reformat"] and "identical" in r["reformat"]["parity"]
    assert "error" not in r["ours_single_member_input"]


def test_gzip_digest_is_streamed_and_equal_to_the_whole_file_digest(tmp_path):
    import gzip
    import hashlib
    import zlib
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import file_bench
    parts = [os.urandom(3000) + bytes([65 + k % 4]) * 500 for k in range(300)]
    blob = b""
    for b in parts:
        c = zlib.compressobj(1, zlib.DEFLATED, 31)
        blob += c.compress(b) + c.flush()
    p = tmp_path / "many.gz"
    p.write_bytes(blob)
    assert file_bench._gz_digest(str(p)) == hashlib.sha256(b"".join(parts)).hexdigest() == hashlib.sha256(gzip.decompress(blob)).hexdigest()
    (tmp_path / "empty.gz").write_bytes(b"")
    assert file_bench._gz_digest(str(tmp_path / "empty.gz")) == hashlib.sha256(b"").hexdigest()
