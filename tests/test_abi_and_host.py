"""CPU-only checks: the C-ABI library loads and exports every symbol include/dmfb_b200.h declares, the
ctypes mirrors match the C struct layout, cfg_init reproduces the reference's constructor checks and the
python-computed direction table (dmfb.py:442-454), and the env sharding helper partitions exactly."""
import ctypes as C
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def P():
    return importlib.import_module("marl-dmfb_b200")


def test_library_exports_every_declared_symbol(P):
    hdr = open(os.path.join(ROOT, "include", "dmfb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:dmfb|meda)_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"dmfb_cfg_t", "dmfb_state_t"}
    assert len(declared) >= 15
    lib = P._native.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"symbols declared in include/dmfb_b200.h but not exported: {missing}"
    assert set(P._native.EXPORTS) <= declared
    assert lib.dmfb_abi_version() == 4


def test_ctypes_mirrors_match_c_layout(P, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dmfb_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(dmfb_cfg_t),sizeof(dmfb_state_t),sizeof(dmfb_out_t),offsetof(dmfb_cfg_t,dir_x),'
                   'offsetof(dmfb_cfg_t,l2_row),sizeof(meda_cfg_t),sizeof(meda_state_t),sizeof(meda_out_t));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    n = P._native
    want = [C.sizeof(n.DmfbCfg), C.sizeof(n.DmfbState), C.sizeof(n.DmfbOut), n.DmfbCfg.dir_x.offset,
            n.DmfbCfg.l2_row.offset, C.sizeof(n.MedaCfg), C.sizeof(n.MedaState), C.sizeof(n.MedaOut)]
    assert got == want


def test_step_flags_match_the_header(P, tmp_path):
    """The Python layer's step flags are the header's macros (compiled, not parsed)."""
    src = tmp_path / "fl.c"
    src.write_text('#include <stdio.h>\n#include <stdint.h>\n#include "dmfb_b200.h"\nint main(){printf("%u %u %u %u %u\\n",'
                   'DMFB_STEP_RECORD_USAGE,DMFB_STEP_FREEZE_TERM,DMFB_STEP_AUTO_RESET,DMFB_STEP_SKIP_TASK_SEARCH,'
                   'DMFB_STEP_SEARCH_SHARE(5));return 0;}\n')
    exe = tmp_path / "fl"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    n = P._native
    assert got == [n.STEP_RECORD_USAGE, n.STEP_FREEZE_TERM, n.STEP_AUTO_RESET, n.STEP_SKIP_TASK_SEARCH, n.step_search_share(5)]
    assert len({*got[:4]}) == 4 and not (got[4] & sum(got[:4])) and got[4] < (1 << 31)     # disjoint bits, bit 31 free


def _py_dir(d, dim, fov):
    hf = fov // 2
    if abs(d) > hf:  # the reference expression, evaluated by python itself (banker's round on float64)
        if d > 0:
            return round((d - hf) / ((dim - hf) / (10 - hf))) + hf
        return round((d + hf) / ((dim - hf) / (10 - hf))) - hf
    return d


@pytest.mark.parametrize("W,L,fov", [(10, 10, 9), (20, 20, 9), (50, 50, 9), (50, 50, 5), (12, 15, 7), (16, 11, 8),
                                     (30, 30, 11), (40, 40, 19), (64, 33, 13), (128, 128, 9), (50, 50, 3)])
def test_cfg_tables_match_python_round(P, W, L, fov):
    cfg = P._native.DmfbCfg()
    assert P._native.load().dmfb_cfg_init(C.byref(cfg), W, L, 2, 0, fov, 1, 0, 0.1) == 0
    assert (cfg.max_step, cfg.obs_dim, cfg.n_actions) == (2 * (W + L), 3 * fov * fov + 2, 5)
    for d in range(-(W - 1), W):
        assert cfg.dir_x[d + W - 1] == _py_dir(d, W, fov), (d, W, fov)
    for d in range(-(L - 1), L):
        assert cfg.dir_y[d + L - 1] == _py_dir(d, L, fov), (d, L, fov)
    hf = fov // 2
    for c in range(2 * hf + 1):
        row = np.zeros((fov, fov), bool)
        if 0 < c <= hf:
            row[:c, :] = True
        elif c > hf:
            row[fov - (c - hf):, :] = True
        bits = np.array([(cfg.l2_row[c][q >> 5] >> (q & 31)) & 1 for q in range(fov * fov)], bool).reshape(fov, fov)
        assert np.array_equal(bits, row)
        bits = np.array([(cfg.l2_col[c][q >> 5] >> (q & 31)) & 1 for q in range(fov * fov)], bool).reshape(fov, fov)
        assert np.array_equal(bits, row.T)


def test_cfg_init_rejects_like_the_reference(P):
    lib, cfg = P._native.load(), P._native.DmfbCfg()
    f = lambda *a: lib.dmfb_cfg_init(C.byref(cfg), *a)  # noqa: E731
    assert f(8, 8, 2, 0, 9, 1, 0, 0.1) == 1     # RuntimeError('Fov is too large')         dmfb.py:139-140
    assert f(5, 5, 5, 0, 5, 1, 0, 0.1) == 2     # TypeError('Too many droplets for DMFB')  dmfb.py:144-146
    assert f(4, 10, 2, 0, 3, 1, 0, 0.1) == 6    # assert width >= 5 and length >= 5        dmfb.py:489
    assert f(10, 10, 0, 0, 5, 1, 0, 0.1) == 3   # assert n_agents > 0                      dmfb.py:490
    assert f(10, 10, 4, 0, 9, 1, 0, 0.1) == 0
    with pytest.raises(RuntimeError, match="Fov is too large"):
        P._native.check(1)
    with pytest.raises(TypeError, match="Too many droplets"):
        P._native.check(2)


def test_missing_library_is_a_loud_error(P, monkeypatch):
    n = P._native
    monkeypatch.setattr(n, "_lib", None)
    monkeypatch.setattr(n._build, "LIB", "/nonexistent/libdmfb_b200.so")
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        n.load()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "marl-dmfb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("# noqa", ""), f"{f} mentions the oracle"


def test_cpu_device_is_refused(P):
    with pytest.raises((RuntimeError, AssertionError)):
        P.BatchedDMFB(4, 10, 10, 4, fov=9, device="cpu")


@pytest.mark.parametrize("n,world", [(65536, 8), (262144, 8), (1000, 3), (7, 8), (0, 2)])
def test_shard_range_partitions(P, n, world):
    spans = [P.shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("cells,n,threads", [(243, 1000, 4), (324, 257, 3), (75, 33, 1), (7, 5, 8), (1083, 64, 2)])
def test_host_unpack_expands_packed_records(P, cells, n, threads):
    """Host half of the packed observation transfer (dmfb_host_set_transfer): 4-bit cells -> int8, direction bytes
    unchanged.  Runs without a GPU."""
    import ctypes as C
    lib = P._native.load()
    rng = np.random.default_rng(cells + n)
    obs = rng.integers(0, 16, (n, cells + 2)).astype(np.int8)
    obs[:, -2:] = rng.integers(-10, 11, (n, 2))
    nb = (cells + 1) // 2
    stride = (nb + 2 + 3) & ~3
    packed = np.zeros((n, stride), np.uint8)
    padded = np.zeros((n, 2 * nb), np.uint8)
    padded[:, :cells] = obs[:, :cells]
    packed[:, :nb] = padded[:, 0::2] | (padded[:, 1::2] << 4)
    packed[:, nb:nb + 2] = obs[:, -2:].view(np.uint8)
    out = np.full((n, cells + 2), 99, np.int8)
    rc = lib.dmfb_host_unpack_records(packed.ctypes.data_as(C.c_void_p), stride, out.ctypes.data_as(C.c_void_p), cells, n,
                                      threads)
    assert rc == 0
    np.testing.assert_array_equal(out, obs)



def test_sub_batch_ranges_cover_the_batch_on_aligned_boundaries():
    """pipeline.sub_batch_ranges: the sub-batches of BatchedDMFB / BatchedMEDA(sub_batches=K) partition [0, N), start on
    multiples of 64 envs (16-byte aligned observation rows for any row length, whole step-kernel tiles) and never
    outnumber the 64-env units."""
    import importlib
    pipe = importlib.import_module("marl-dmfb_b200.pipeline")
    for n in (1, 63, 64, 65, 1000, 4096, 65536, 65537):
        for k in (1, 2, 3, 4, 8, 5000):
            r = pipe.sub_batch_ranges(n, k)
            assert r[0][0] == 0 and r[-1][1] == n and len(r) <= max(1, min(k, (n + 63) // 64))
            assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(hi > lo for lo, hi in r)
            assert all(lo % 64 == 0 for lo, _ in r)
    assert pipe.sub_batch_ranges(65536, 4) == [(0, 16384), (16384, 32768), (32768, 49152), (49152, 65536)]


def test_gpu_cpu_affinity_parses_nvidia_smi_topo():
    """host.gpu_cpu_affinity: the "CPU Affinity" column of `nvidia-smi topo -m` (ANSI-underlined header, tab separated),
    used by bench.py to keep each rank's pinned buffers and unpack threads on its GPU's NUMA node."""
    import importlib
    host = importlib.import_module("marl-dmfb_b200.host")
    one = "\x1b[4m\tGPU0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\x1b[0m\nGPU0\t X \t0-15\t0\t\tN/A\n\nLegend:\n\n  X    = Self\n"
    assert host.gpu_cpu_affinity(0, one) == set(range(16)) and host.gpu_cpu_affinity(1, one) is None
    two = ("\x1b[4m\tGPU0\tGPU1\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\x1b[0m\n"
           "GPU0\t X \tNV18\tPIX\t0-55,112-167\t0\t\tN/A\nGPU1\tNV18\t X \tSYS\t56-111,168-223\t1\t\tN/A\nNIC0\tPIX\tSYS\t X \n")
    assert host.gpu_cpu_affinity(0, two) == set(range(0, 56)) | set(range(112, 168))
    assert min(host.gpu_cpu_affinity(1, two)) == 56 and len(host.gpu_cpu_affinity(1, two)) == 112
    assert host.gpu_cpu_affinity(0, "no such table") is None and host.gpu_cpu_affinity(0, "") is None
