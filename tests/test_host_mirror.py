"""The C++ host mirror (frankenpaxos_b200/host/frankenpaxos_host.hpp: Actor / Transport /
Chan / FakeTransport / GpuProxyLeader / GpuAcceptor) above the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "host", "fake_transport_parity")


def build():
    from frankenpaxos_b200 import build as B
    from oracle import fpx_oracle_py as O
    lib = B.build()
    O.build()
    src = os.path.join(ROOT, "tests", "host", "fake_transport_parity.cc")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", src, "-o", BIN,
           "-L" + os.path.dirname(lib), "-lfpx", "-L" + os.path.join(ROOT, "oracle"), "-lfpx_oracle",
           "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    subprocess.check_call(cmd)
    return BIN


def test_host_mirror_compiles_and_links():
    assert os.path.exists(build())


def test_batching_shim_on_the_oracle_backend():
    """No GPU: the host logic alone (GpuProxyLeader / GpuAcceptor buffering a delivery burst, one flush,
    replies routed back in order) against one-handler-per-message, both over the oracle backend --
    BASELINE cfg1 on FakeTransport with a leader change, three seeds."""
    r = subprocess.run([build(), "128", "cpu"], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("PARITY OK") == 3


@pytest.mark.gpu
def test_fake_transport_run_cfg1_parity():
    """BASELINE cfg1 (MultiPaxos f=1, 3 acceptors, 128 slots on FakeTransport) with a
    leader change: batched GPU actors vs per-message oracle actors, identical transcripts."""
    r = subprocess.run([build(), "128"], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("PARITY OK") == 3
