"""Checkpoint / restart (shamrock_b200/csrc/dump.cu; container of shamrock/src/io/ShamrockDump.cpp:25-274):
a model restarted from a dump continues bit-identically, patch list and scheduler state included."""
import json
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from shamrock_b200 import _capi  # noqa: E402
from tests import scenarios as S  # noqa: E402

FIELDS = [nm for nm, _ in _capi.HOST_FIELDS]


def read_headers(path):
    out, off = [], 0
    raw = open(path, "rb").read()
    for _ in range(3):
        (n,) = struct.unpack_from("<Q", raw, off)
        out.append(json.loads(raw[off + 8: off + 8 + n].decode()))
        off += 8 + n
    return out, off, len(raw)


@pytest.mark.parametrize("scenario", ["periodic_split", "disc"])
def test_restart_continues_bit_identically(tmp_path, scenario):
    if scenario == "disc":
        sc = S.disc(4000, "M4", grid=(2, 2, 1))
    else:
        sc = S.periodic_box(10000, "M4", "cd10", jitter=0.2, grid=(2, 1, 1))
    a = S.make_cuda(sc, keep_step_data=False)
    a.init_scheduler(10**9, 0, step_freq=0)
    for _ in range(2):
        a.evolve_once()
    if scenario == "periodic_split":
        a.split_patch(1)  # the patch list of the dump is not the initial grid
        a.evolve_once()
    f = tmp_path / "state.sham"
    a.dump(f)
    user, pmeta, table = read_headers(f)[0]
    assert user["format"] == "shamb200-1" and user["time"] == a.state()["time"]
    assert [p["id_patch"] for p in pmeta["patchlist"]] == [a.patch_info(ip)["id"] for ip in range(a.patch_count)]
    assert table["pids"] == [p["id_patch"] for p in pmeta["patchlist"]]
    hdr, off, size = read_headers(f)
    assert off + table["offsets"][-1] + table["bytecounts"][-1] == size
    # a fresh model with ANOTHER configuration: the dump replaces it
    other = S.periodic_box(100, "M6", "constant")
    b = S.make_cuda(other, keep_step_data=False)
    b.load_dump(f)
    assert b.patch_count == a.patch_count
    sa, sb = a.state(), b.state()
    for k in ("time", "dt", "cfl_multiplier"):
        assert sa[k] == sb[k]
    for ip in range(a.patch_count):
        assert a.patch_info(ip) == b.patch_info(ip) and a.patch_size(ip) == b.patch_size(ip)
        for nm in FIELDS:
            assert np.array_equal(a.get(ip, nm), b.get(ip, nm)), (ip, nm)
    for _ in range(2):
        sa, sb = a.evolve_once(), b.evolve_once()
        assert sa["dt"] == sb["dt"] and sa["npart"] == sb["npart"]
    for ip in range(a.patch_count):
        for nm in FIELDS:
            assert np.array_equal(a.get(ip, nm), b.get(ip, nm)), (ip, nm)


def test_not_a_dump(tmp_path):
    f = tmp_path / "junk"
    f.write_bytes(b"\x05\x00\x00\x00\x00\x00\x00\x00hello" * 4)
    m = S.make_cuda(S.periodic_box(100, "M4", "cd10"))
    with pytest.raises(_capi.ShamB200Error):
        m.load_dump(f)
