"""liboimgpu as a `struct spdk_copy_engine` (integration/spdk/copy_engine_oimgpu.c) under the
reference's UNMODIFIED bdev layer and Malloc bdev.  The driver (integration/spdk/bdevio_oimgpu.c) runs the
data-integrity cases of S/test/bdev/bdevio/bdevio.c:388-800 through spdk_bdev_writev/readv/write_zeroes/unmap."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "integration", "spdk", "_build", "bdevio_oimgpu")


def _build():
    if os.path.isdir("/root/reference/vendor/github.com/spdk/spdk"):
        from oim_b200 import build
        build.build()
        r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "integration", "spdk")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    if not os.path.exists(BIN):
        pytest.skip("integration/spdk/_build/bdevio_oimgpu not built (needs /root/reference at build time)")


def _run(env_extra):
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=240, env={**os.environ, **env_extra})
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_bdevio_cases_on_the_references_memcpy_engine():
    """control: the same driver with the engine switched off - the cases themselves are sound"""
    _build()
    out = _run({"OIMGPU_COPY_ENGINE": "off"})
    assert out["engine"] == "memcpy" and out["failed"] == 0 and out["cases"] >= 50


@pytest.mark.gpu
def test_bdevio_cases_on_the_oimgpu_copy_engine():
    """the registered engine wins (copy_create_cb, S/lib/copy/copy_engine.c:186-203): every byte bdev_malloc.c moves goes
    through oim_copy_kernel / oim_fill_kernel, asynchronously, completions reaped by the channel's poller"""
    _build()
    out = _run({})
    assert out["engine"] == "oimgpu", out
    assert out["failed"] == 0 and out["cases"] >= 50
    assert out["engine_ops"] >= 40 and out["engine_bytes"] > 10 << 20
