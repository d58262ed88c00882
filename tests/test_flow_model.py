"""The streaming form of the Ion Torrent flow model that the device runs (dwgsim_b200/csrc/flow_model.h, host/device shared
source) against the oracle's restatement of generate_errors_flows (src/dwgsim.c:246-417): same reads, same Philox FLOW
draws, on random flow orders, error rates from 0 to 0.5, both strands, homopolymer-rich reads.  No GPU needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_streaming_flow_model_equals_oracle(oracle, tmp_path):
    exe = str(tmp_path / "flow_model_check")
    cmd = ["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "flow_model_check.cpp"),
           "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "150000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "cases ok" in r.stdout
