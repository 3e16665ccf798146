"""bf_motion_compensator without a GPU: the flag surface of the reference's tool (bf_motion_compensator.cpp:64-130)
plus the additions SURVEY 8b asks for, and the no-CPU-fallback rule (the tool must fail loudly, not compute on the host)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "better_flow_b200", "bf_motion_compensator")

REFERENCE_FLAGS = ["--refresh-time=", "--refresh-event-count=", "-i/--interactive", "-G", "--stm-disable", "--img", "--img-prefix",
                   "--video", "--video-name", "--video-fps=", "--bufferize-file", "-o <name>/--outfile=", "--version"]
ADDED_FLAGS = ["--max-iter=", "--scale=", "--slice-time=", "--max-events=", "--sensor=", "--flow-out=", "--batch=", "--gpus=",
               "--optimizer=", "--device=", "--quiet"]


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "better_flow_b200"), "all"])
    return CLI


def test_help_lists_the_reference_flags_and_the_additions(cli):
    r = subprocess.run([cli, "--help"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0
    text = r.stdout + r.stderr
    for f in REFERENCE_FLAGS + ADDED_FLAGS:
        assert f in text, f
    # the reference's defaults (bf_motion_compensator.cpp:6-7,44-45): 50000 events, 200 ms, 33 ms, 20000 events
    assert "50000 events" in text and "0.200000 seconds" in text
    assert "default = 0.033000" in text and "default = 20000" in text


def test_version_and_missing_file(cli):
    r = subprocess.run([cli, "--version"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "better flow" in r.stdout
    r = subprocess.run([cli], capture_output=True, text=True, timeout=60)
    assert "usage" in (r.stdout + r.stderr)


def test_no_cpu_fallback(cli, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    f = tmp_path / "ev.txt"
    with open(f, "w") as o:
        for k in range(3000):
            o.write("%.9f %d %d %d\n" % (1.0 + k * 1e-5, k % 240, (7 * k) % 180, k & 1))
    r = subprocess.run([cli, "--quiet", "--refresh-event-count=1000", str(f)], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
